"""Generate tests/golden/*.pt from the REAL reference (imported from /root/reference).

TEST INFRASTRUCTURE.  Run in the build container only:  python -m oracle.make_golden
The fixtures are what pins the oracle restatement (oracle/restate.py) and the
CUDA path on the GPU box, where the reference tree does not exist.
All inputs are seeded; tensors are kept small (a few hundred KB per file).
"""
from __future__ import annotations

import contextlib
import io
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_loader  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def cell_inputs(regime: str, B, NH, S, DH, seed, dtype=torch.float64):
    """Synthetic cell inputs in the regimes of SURVEY.md 8d config 2."""
    g = torch.Generator().manual_seed(seed)
    # values are drawn in float32 (so fixtures store them losslessly as float32) and widened
    rn = lambda *s: torch.randn(*s, generator=g, dtype=torch.float32).to(dtype)
    if regime == "bottleneck":          # observed statistics at the real bottleneck (post init_weights)
        q, k, v = 0.06 * rn(B, NH, S, DH), 0.06 * rn(B, NH, S, DH), 0.12 * rn(B, NH, S, DH)
        ig, fg = -0.67 + 0.48 * rn(B, NH, S, 1), 0.41 + 1.03 * rn(B, NH, S, 1)
    elif regime == "randn_f4":
        q, k, v = rn(B, NH, S, DH), rn(B, NH, S, DH), rn(B, NH, S, DH)
        ig, fg = rn(B, NH, S, 1), 4.0 + rn(B, NH, S, 1)
    elif regime == "randn_f0":
        q, k, v = rn(B, NH, S, DH), rn(B, NH, S, DH), rn(B, NH, S, DH)
        ig, fg = rn(B, NH, S, 1), rn(B, NH, S, 1)
    elif regime == "randn_fm2":
        q, k, v = rn(B, NH, S, DH), rn(B, NH, S, DH), rn(B, NH, S, DH)
        ig, fg = rn(B, NH, S, 1), -2.0 + rn(B, NH, S, 1)
    else:
        raise ValueError(regime)
    # round-trip through float32 so the stored float32 copy is exact
    return tuple(t.float().to(dtype) for t in (q, k, v, ig, fg))


def gen_cell(ns):
    vl = ns.vision_lstm
    cases = {}
    for name, (regime, B, NH, S, DH, seed) in {
        "bottleneck_s320_dh16": ("bottleneck", 1, 4, 320, 16, 0),
        "randn_f4_s200_dh16": ("randn_f4", 2, 2, 200, 16, 1),     # ragged vs a 128 chunk
        "randn_f0_s256_dh32": ("randn_f0", 1, 2, 256, 32, 2),
        "randn_fm2_s130_dh8": ("randn_fm2", 1, 4, 130, 8, 3),
        "randn_f4_s256_dh64": ("randn_f4", 1, 1, 256, 64, 4),
    }.items():
        q, k, v, ig, fg = [t.requires_grad_() for t in cell_inputs(regime, B, NH, S, DH, seed)]
        h = vl.parallel_stabilized_simple(q, k, v, ig, fg)
        dh = torch.randn(h.shape, generator=torch.Generator().manual_seed(100 + seed), dtype=torch.float32).to(h.dtype)
        grads = torch.autograd.grad(h, [q, k, v, ig, fg], dh)
        h32 = vl.parallel_stabilized_simple(*[t.detach().float() for t in (q, k, v, ig, fg)])
        cases[name] = dict(
            regime=regime, q=q.detach().float(), k=k.detach().float(), v=v.detach().float(),
            ig=ig.detach().float(), fg=fg.detach().float(), h=h.detach(), dh=dh.float(),
            dq=grads[0].float(), dk=grads[1].float(), dv=grads[2].float(), dig=grads[3].float(),
            dfg=grads[4].float(), h_fp32_ref=h32)
    torch.save(cases, os.path.join(OUT, "cell.pt"))


def randomise_like_init_weights(mod, ns, seed):
    """utils.init_weights (utils.py:191-215) re-initialises every nn.Linear
    xavier-normal with N(0,1) biases; additionally perturb the LayerNorm /
    outnorm / skip / conv parameters a little so that every parameter matters."""
    torch.manual_seed(seed)
    mod.apply(ns.utils.init_weights)
    with torch.no_grad():
        for n, p in mod.named_parameters():
            if n.endswith("norm.weight") or n.endswith("outnorm.weight"):
                p.add_(0.1 * torch.randn_like(p))
            if n.endswith("learnable_skip"):
                p.add_(0.1 * torch.randn_like(p))


BLOCK_CASES = {
    "dim32_s200_fwd": (32, 200, 2, False, 10),
    "dim32_s200_rev": (32, 200, 2, True, 11),
    "dim16_s150_fwd": (16, 150, 1, False, 12),      # 32^3-stage dims: E=32, DH=8
    "dim64_s140_rev": (64, 140, 1, True, 13),       # class default f_maps=8: DH=32
}
WIDE_BLOCK_CASES = {
    "dim128_s160_rev": (128, 160, 1, True, 14),     # f_maps = 16: E = 256, DH = 64
}


def gen_block(ns, table=None, fname="vil_block.pt"):
    vl = ns.vision_lstm
    cases = {}
    for name, (dim, S, B, rev, seed) in (table or BLOCK_CASES).items():
        direction = vl.SequenceTraversal.ROWWISE_FROM_BOT_RIGHT if rev else vl.SequenceTraversal.ROWWISE_FROM_TOP_LEFT
        blk = vl.ViLBlock(dim=dim, direction=direction).double()
        randomise_like_init_weights(blk, ns, seed)
        blk = blk.double()
        x = torch.randn(B, S, dim, dtype=torch.float64, generator=torch.Generator().manual_seed(seed)).requires_grad_()
        y = blk(x)
        dy = torch.randn(y.shape, dtype=torch.float64, generator=torch.Generator().manual_seed(seed + 50))
        params = dict(blk.named_parameters())
        grads = torch.autograd.grad(y, [x] + list(params.values()), dy)
        cases[name] = dict(
            dim=dim, reverse=rev, x=x.detach(), y=y.detach(), dy=dy, dx=grads[0],
            state_dict={k_: v_.detach().clone() for k_, v_ in blk.state_dict().items()},
            param_grads={n: g for n, g in zip(params.keys(), grads[1:])})
    torch.save(cases, os.path.join(OUT, fname))


def gen_block_wide(ns):
    gen_block(ns, WIDE_BLOCK_CASES, "vil_block_wide.pt")


def gen_wrapper(ns):
    enc = ns.UxLSTMEnc_3d
    wrap = enc.ViLLayer(dim=32)
    randomise_like_init_weights(wrap, ns, 20)
    x = torch.randn(2, 32, 4, 6, 8, generator=torch.Generator().manual_seed(20))
    with torch.no_grad():
        y32 = wrap(x)
        y64 = wrap.double()(x.double())
    torch.save(dict(x=x, y_fp32=y32, y=y64, state_dict={k: v.detach().clone() for k, v in wrap.state_dict().items()}),
               os.path.join(OUT, "vil_wrapper.pt"))


def gen_poe(ns):
    bb, RA, loss = ns.buildingblocks, ns.RA_HVED, ns.loss
    g = torch.Generator().manual_seed(30)
    B, C, d = 2, 2, 6
    mod_mu = 1.3 * torch.randn(4, B, C, d, d, d, generator=g, dtype=torch.float64)
    mod_lv = 1.4 * torch.randn(4, B, C, d, d, d, generator=g, dtype=torch.float64)
    mod_lv[0, 0, 0, 0, 0, :3] = torch.tensor([80.0, -80.0, 50.0], dtype=torch.float64)   # exercise clip
    # assemble (5,B,...) exactly as RA_HVED.py:576-580 does (prior first, clip on logvar)
    mu = torch.cat([torch.zeros(1, B, C, d, d, d, dtype=torch.float64), mod_mu], 0)
    lv = torch.cat([torch.zeros(1, B, C, d, d, d, dtype=torch.float64), RA.clip(mod_lv)], 0)
    experts = bb.ProductOfExperts()
    out = dict(mod_mu=mod_mu, mod_logvar=mod_lv, subsets=[list(s) for s in RA.SUBSETS_MODALITIES])
    pd_mu, pd_lv = [], []
    for subset in RA.SUBSETS_MODALITIES:
        a, b = experts(mu, lv, subset)
        pd_mu.append(a)
        pd_lv.append(b)
    out["pd_mu"], out["pd_logvar"] = torch.stack(pd_mu), torch.stack(pd_lv)
    # gradients through clip + PoE for two subsets
    for idx in (14, 5):
        mm = mod_mu.clone().requires_grad_()
        ll = mod_lv.clone().requires_grad_()
        mu5 = torch.cat([torch.zeros(1, B, C, d, d, d, dtype=torch.float64), mm], 0)
        lv5 = torch.cat([torch.zeros(1, B, C, d, d, d, dtype=torch.float64), RA.clip(ll)], 0)
        a, b = experts(mu5, lv5, RA.SUBSETS_MODALITIES[idx])
        ga = torch.randn(a.shape, generator=g, dtype=torch.float64)
        gb = torch.randn(b.shape, generator=g, dtype=torch.float64)
        gm, gl = torch.autograd.grad([a, b], [mm, ll], [ga, gb])
        out[f"grad_{idx}"] = dict(g_mu=ga, g_logvar=gb, d_mod_mu=gm, d_mod_logvar=gl)
    # per-sample drop (ProductOfExperts2, buildingblocks.py:875-886); mutates mu in place
    drop = torch.tensor([[False, True, False, True], [True, False, False, False]])
    mu_mut = mu.clone()
    a, b = bb.ProductOfExperts2()(mu_mut, lv.clone(), drop)
    out["drop"], out["drop_pd_mu"], out["drop_pd_logvar"], out["drop_mu_after"] = drop, a, b, mu_mut
    # reparametrize (RA_HVED.py:741-747) with the global generator seeded
    torch.manual_seed(31)
    z = RA.reparametrize(out["pd_mu"][14].float(), out["pd_logvar"][14].float(), False)
    torch.manual_seed(31)
    noise = torch.empty_like(z).normal_()
    out["reparam_seed"], out["reparam_z"], out["reparam_noise"] = 31, z, noise
    # compute_KLD (loss.py:85-115) on the (B,5,...) view the model returns
    out["kld_14"] = loss.compute_KLD(mu.transpose(1, 0), lv.transpose(1, 0), [14])
    out["kld_3_7_12"] = loss.compute_KLD(mu.transpose(1, 0), lv.transpose(1, 0), [3, 7, 12])
    torch.save(out, os.path.join(OUT, "poe.pt"))


def gen_smvae_extras(ns):
    """Round-2 fixtures from the real reference: compute_KLD with a NON-standard prior in slab 0 (loss.py:95-97, 113) and its
    gradients w.r.t. all five slabs, clip's gradient mask (RA_HVED.py:749-753) through the fusion, ZeroLayerF fwd + bwd
    (buildingblocks.py:308-323)."""
    RA, bb, loss = ns.RA_HVED, ns.buildingblocks, ns.loss
    g = torch.Generator().manual_seed(77)
    B, C, d = 2, 2, 6
    out = {}
    mu = (1.3 * torch.randn(B, 5, C, d, d, d, generator=g, dtype=torch.float64)).requires_grad_()
    lv = (0.9 * torch.randn(B, 5, C, d, d, d, generator=g, dtype=torch.float64)).requires_grad_()
    k = loss.compute_KLD(mu, lv, [14, 2, 9])
    gm, gl = torch.autograd.grad(k, [mu, lv])
    out["kld_prior"] = dict(mu=mu.detach(), logvar=lv.detach(), subsets=[14, 2, 9], kld=k.detach(), d_mu=gm, d_logvar=gl)
    # clip + PoE: raw logvars with entries beyond +-50, gradient must vanish there (torch.clamp's autograd)
    mm = 1.3 * torch.randn(4, B, C, d, d, d, generator=g, dtype=torch.float64)
    raw = 30.0 * torch.randn(4, B, C, d, d, d, generator=g, dtype=torch.float64)
    mm_r, raw_r = mm.clone().requires_grad_(), raw.clone().requires_grad_()
    mu5 = torch.cat([torch.zeros(1, B, C, d, d, d, dtype=torch.float64), mm_r], 0)
    lv5 = torch.cat([torch.zeros(1, B, C, d, d, d, dtype=torch.float64), RA.clip(raw_r)], 0)
    res = {}
    for idx in (14, 6):
        a, b = bb.ProductOfExperts()(mu5, lv5, RA.SUBSETS_MODALITIES[idx])
        ga = torch.randn(a.shape, generator=g, dtype=torch.float64)
        gb = torch.randn(b.shape, generator=g, dtype=torch.float64)
        dm, dl = torch.autograd.grad([a, b], [mm_r, raw_r], [ga, gb], retain_graph=True)
        res[idx] = dict(pd_mu=a.detach(), pd_logvar=b.detach(), g_mu=ga, g_logvar=gb, d_mod_mu=dm, d_raw_logvar=dl)
    out["clip_poe"] = dict(mod_mu=mm, raw_logvar=raw, clipped=RA.clip(raw), cases=res)
    # ZeroLayerF
    x = torch.randn(3, 4, 5, 5, 5, generator=g, dtype=torch.float64).requires_grad_()
    alpha = torch.tensor([True, False, True])
    y = bb.ZeroLayerF.apply(x, alpha)
    gy = torch.randn(y.shape, generator=g, dtype=torch.float64)
    (dx,) = torch.autograd.grad(y, x, gy)
    out["zero_layer"] = dict(x=x.detach(), alpha=alpha, y=y.detach(), gy=gy, dx=dx)
    torch.save(out, os.path.join(OUT, "smvae_extras.pt"))


def gen_losses(ns):
    """DiceLoss (loss.py:188-209, 257-301) of the real reference on sigmoid-like probabilities and a binary mask: value,
    per-channel coefficients and the gradient w.r.t. the probabilities; one channel is empty on both sides (clamp active)."""
    import contextlib
    import io
    loss = ns.loss
    g = torch.Generator().manual_seed(5)
    p = torch.rand(2, 4, 6, 7, 5, generator=g, dtype=torch.float64)
    t = (torch.rand(2, 4, 6, 7, 5, generator=g) > 0.5).double()
    p[:, 3] = 0
    t[:, 3] = 0
    pr = p.clone().requires_grad_()
    with contextlib.redirect_stdout(io.StringIO()):
        val = loss.DiceLoss()(pr, t)
        per = loss.compute_per_channel_dice(p, t)
    (dp,) = torch.autograd.grad(val, pr)
    torch.save(dict(p=p, t=t, loss=val.detach(), per_channel=per, dp=dp), os.path.join(OUT, "losses.pt"))


def gen_conv_norm(ns):
    """The normalisation + activation layers of the convolution path, run through the REAL reference modules on CPU in fp64:
    SingleConv order 'ilc' (buildingblocks.py:444-462), BasicConv (buildingblocks.py:11-31) and DuSEAttention with its two
    BatchNorm3d layers (modules/DuSFE.py:87-154) in train and eval mode.  Forward values and gradients."""
    import importlib
    bb = ns.buildingblocks
    dusfe = importlib.import_module("modules.DuSFE")
    g = torch.Generator().manual_seed(11)
    rn = lambda *s: torch.randn(*s, generator=g, dtype=torch.float32).double()
    out = {}

    torch.manual_seed(3)
    sc = bb.SingleConv(4, 6, kernel_size=3, order="ilc", padding=1).double()
    x = (rn(2, 4, 9, 10, 12) * 1.7 + 0.3).requires_grad_()                 # spatial 1080: not a multiple of a 16-byte vector of fp16
    mid = sc.LeakyReLU(sc.instancenorm(x))                                  # in-place on the norm output, as in the Sequential
    gm = rn(*mid.shape)
    (dx_mid,) = torch.autograd.grad(mid, x, gm, retain_graph=False)
    x2 = x.detach().clone().requires_grad_()
    y = sc(x2)
    gy = rn(*y.shape)
    grads = torch.autograd.grad(y, [x2, sc.conv.weight, sc.conv.bias], gy)
    out["single_conv_ilc"] = dict(x=x.detach(), mid=mid.detach(), gm=gm, dx_mid=dx_mid, y=y.detach(), gy=gy, dx=grads[0],
                                  dconv_weight=grads[1], dconv_bias=grads[2], slope=sc.LeakyReLU.negative_slope,
                                  state_dict={k: v.detach().clone() for k, v in sc.state_dict().items()})

    torch.manual_seed(4)
    bc = bb.BasicConv(3, 5, 3, padding=1).double()
    x = rn(1, 3, 8, 8, 16).requires_grad_()
    y = bc(x)
    gy = rn(*y.shape)
    grads = torch.autograd.grad(y, [x, bc.conv.weight], gy)
    out["basic_conv"] = dict(x=x.detach(), y=y.detach(), gy=gy, dx=grads[0], dconv_weight=grads[1], slope=bc.relu.negative_slope,
                             state_dict={k: v.detach().clone() for k, v in bc.state_dict().items()})

    torch.manual_seed(5)
    att = dusfe.DuSEAttention(4).double()
    with torch.no_grad():
        for bn in (att.bn_fuse_ch1, att.bn_fuse_ch2):                       # non-trivial affine parameters and running statistics
            bn.weight.copy_(1 + 0.3 * rn(4)), bn.bias.copy_(0.2 * rn(4))
            bn.running_mean.copy_(0.1 * rn(4)), bn.running_var.copy_(1 + 0.2 * torch.rand(4, generator=g).double())
    sd0 = {k: v.detach().clone() for k, v in att.state_dict().items()}
    a, b = (rn(3, 4, 6, 8, 8) + 0.5).requires_grad_(), (0.7 * rn(3, 4, 6, 8, 8)).requires_grad_()
    rec = {}
    def record(m, i, o):
        rec["train" if m.training else "eval"] = (i[0].detach().clone(), o.detach().clone())

    hook = att.bn_fuse_ch1.register_forward_hook(record)
    att.train()
    y1, y2 = att(a, b)
    g1, g2 = rn(*y1.shape), rn(*y2.shape)
    params = [att.bn_fuse_ch1.weight, att.bn_fuse_ch1.bias, att.bn_fuse_ch2.weight, att.bn_fuse_ch2.bias]
    grads = torch.autograd.grad([y1, y2], [a, b] + params, [g1, g2])
    sd1 = {k: v.detach().clone() for k, v in att.state_dict().items()}
    att.eval()
    e1, e2 = att(a, b)
    egrads = torch.autograd.grad([e1, e2], [a, b] + params, [g1, g2])
    hook.remove()
    out["duse_attention"] = dict(a=a.detach(), b=b.detach(), g1=g1, g2=g2, state_dict_before=sd0, state_dict_after_train=sd1,
                                 train_y1=y1.detach(), train_y2=y2.detach(), train_grads=[t.detach() for t in grads],
                                 eval_y1=e1.detach(), eval_y2=e2.detach(), eval_grads=[t.detach() for t in egrads],
                                 bn1_train_in=rec["train"][0], bn1_train_out=rec["train"][1], bn1_eval_in=rec["eval"][0],
                                 bn1_eval_out=rec["eval"][1])
    torch.save(out, os.path.join(OUT, "conv_norm.pt"))


def gen_atten(ns):
    """AttenModule2 (buildingblocks.py:259-301) of the real reference on CPU in fp64: output and the gradients w.r.t. both inputs and
    all eight parameters.  Spatial size (9, 11, 37): ragged against the kernel's 8 x 8 x 32 tiles in every dimension."""
    bb = ns.buildingblocks
    g = torch.Generator().manual_seed(13)
    rn = lambda *s: torch.randn(*s, generator=g, dtype=torch.float32).double()
    torch.manual_seed(6)
    att = bb.AttenModule2(8, 4).double()
    with torch.no_grad():
        for q in att.parameters():                     # larger than the default init so that the gates leave the linear regime
            q.mul_(4.0)
    seg_x, enc_x = rn(2, 3, 9, 11, 37).requires_grad_(), rn(2, 5, 9, 11, 37).requires_grad_()
    y = att(seg_x, enc_x)
    gy = rn(*y.shape)
    names = [n for n, _ in att.named_parameters()]
    grads = torch.autograd.grad(y, [seg_x, enc_x] + list(att.parameters()), gy)
    torch.save(dict(seg_x=seg_x.detach(), enc_x=enc_x.detach(), y=y.detach(), gy=gy, d_seg_x=grads[0], d_enc_x=grads[1],
                    param_grads=dict(zip(names, grads[2:])), state_dict={k: v.detach().clone() for k, v in att.state_dict().items()}),
               os.path.join(OUT, "atten_module2.pt"))


def gen_dwconv3(ns):
    """BasicConv(C, C, 3, padding=1, groups=C) of the real reference (the latent levels' conv_block, RA_HVED.py:406) on CPU in fp64:
    the depthwise convolution's output, the block's output and the gradients; spatial size ragged against 8 x 8 x 32 tiles."""
    bb = ns.buildingblocks
    g = torch.Generator().manual_seed(17)
    rn = lambda *s: torch.randn(*s, generator=g, dtype=torch.float32).double()
    torch.manual_seed(7)
    bc = bb.BasicConv(4, 4, 3, padding=1, groups=4).double()
    x = rn(2, 4, 9, 11, 37).requires_grad_()
    conv_out = bc.conv(x)
    y = bc(x)
    gy = rn(*y.shape)
    grads = torch.autograd.grad(y, [x, bc.conv.weight], gy)
    gc = rn(*conv_out.shape)
    cgrads = torch.autograd.grad(conv_out, [x, bc.conv.weight], gc)
    torch.save(dict(x=x.detach(), conv_out=conv_out.detach(), y=y.detach(), gy=gy, dx=grads[0], dconv_weight=grads[1], gc=gc,
                    conv_dx=cgrads[0], conv_dweight=cgrads[1], state_dict={k: v.detach().clone() for k, v in bc.state_dict().items()}),
               os.path.join(OUT, "dwconv3.pt"))


def gen_model_boundary(ns):
    """Run the full XLSTM_HVED on a small seeded volume and record the tensors
    that cross the hot-path boundary (RA_HVED.py:588-597 and 623-626)."""
    model = ref_loader.build_model(f_maps=4, seed=1).eval()
    rec = dict(poe=[], vil=None)
    orig_experts = model.experts.forward

    def experts_hook(mu, logvar, subset, eps=1e-8):
        a, b = orig_experts(mu, logvar, subset, eps)
        rec["poe"].append(dict(mu=mu.detach().clone(), logvar=logvar.detach().clone(), subset=list(subset),
                               pd_mu=a.detach().clone(), pd_logvar=b.detach().clone()))
        return a, b

    model.experts.forward = experts_hook
    orig_vil = model.mViL.forward

    def vil_hook(x):
        y = orig_vil(x)
        rec["vil"] = dict(x=x.detach().clone(), y=y.detach().clone())
        return y

    model.mViL.forward = vil_hook
    torch.manual_seed(2)
    x = torch.rand(1, 4, 48, 48, 48)
    x[:, 1] = 0                                   # subset 12 = (0,2,3): modality 1 missing
    with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
        seg, _ = model(x, [12], valid=True)
    out = dict(subset_idx=12, input_seed=2, input_shape=list(x.shape), poe=rec["poe"], vil=rec["vil"],
               vil_state_dict={k: v.detach().clone() for k, v in model.mViL.state_dict().items()},
               seg_mean=seg.mean().item(), seg_frac_pos=(seg > 0.5).float().mean().item())
    torch.save(out, os.path.join(OUT, "model_boundary.pt"))


def main():
    ns = ref_loader.load_reference()
    assert ns is not None, "reference tree not found"
    os.makedirs(OUT, exist_ok=True)
    only = sys.argv[1:]            # e.g. `python oracle/make_golden.py smvae_extras` regenerates one file
    gens = dict(cell=gen_cell, vil_block=gen_block, vil_block_wide=gen_block_wide, vil_wrapper=gen_wrapper, poe=gen_poe, smvae_extras=gen_smvae_extras, losses=gen_losses,
                conv_norm=gen_conv_norm, atten_module2=gen_atten, dwconv3=gen_dwconv3, model_boundary=gen_model_boundary)
    for name, fn in gens.items():
        if not only or name in only:
            fn(ns)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
