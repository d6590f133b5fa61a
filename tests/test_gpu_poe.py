"""GPU parity of the fused S-MVAE PoE kernel against the reference fixtures and the oracle."""
import pytest
import torch

from conftest import load_golden, rel_l2, rel_linf
from oracle import restate

pytestmark = pytest.mark.gpu

TOL = 1e-3   # BASELINE.json north_star: PoE mu/logvar within 1e-3 relative


def _mu5(c):
    z = torch.zeros(1, *c["mod_mu"].shape[1:], dtype=torch.float64)
    return torch.cat([z, c["mod_mu"]], 0), torch.cat([z, restate.clip_logvar(c["mod_logvar"])], 0)


def test_poe_all_15_subsets_one_launch_golden():
    from xlstm_hved_b200 import ops
    c = load_golden("poe.pt")
    mu, lv = _mu5(c)
    pm, pl, _, kld = ops.poe_fwd(mu.float().cuda(), lv.float().cuda(), restate.SUBSETS_MODALITIES, want_kld=True)
    assert rel_linf(pm, c["pd_mu"]) < TOL and rel_linf(pl, c["pd_logvar"]) < TOL
    assert rel_l2(pm, c["pd_mu"]) < 1e-5 and rel_l2(pl, c["pd_logvar"]) < 1e-5
    n = mu[0].numel()
    assert abs(0.5 * kld[14].item() / n - c["kld_14"].item()) < 1e-4 * abs(c["kld_14"].item())
    k3 = 0.5 * (kld[3] + kld[7] + kld[12]).item() / n / 3
    assert abs(k3 - c["kld_3_7_12"].item()) < 1e-4 * abs(c["kld_3_7_12"].item())


def test_poe_drop_mask_and_reparam_golden():
    from xlstm_hved_b200 import ops
    c = load_golden("poe.pt")
    mu, lv = _mu5(c)
    pm, pl, _, _ = ops.poe_fwd(mu.float().cuda(), lv.float().cuda(), [(0, 1, 2, 3)], drop=c["drop"].cuda())
    assert rel_linf(pm[0], c["drop_pd_mu"]) < TOL and rel_linf(pl[0], c["drop_pd_logvar"]) < TOL
    z = ops.reparam_fwd(c["pd_mu"][14].float().cuda(), c["pd_logvar"][14].float().cuda(), c["reparam_noise"].cuda())
    assert rel_linf(z, c["reparam_z"]) < 1e-5
    # fused sampling inside the PoE launch
    pm, pl, zf, _ = ops.poe_fwd(mu.float().cuda(), lv.float().cuda(), [(0, 1, 2, 3)], noise=c["reparam_noise"].cuda()[None])
    assert rel_linf(zf[0], c["reparam_z"]) < 1e-4


def test_poe_backward_golden():
    from xlstm_hved_b200 import ops
    c = load_golden("poe.pt")
    mu, lv = _mu5(c)
    for idx in (14, 5):
        g = c[f"grad_{idx}"]
        dmu, dlv = ops.poe_bwd(mu.float().cuda(), lv.float().cuda(), [restate.SUBSETS_MODALITIES[idx]],
                               g_mu=g["g_mu"].float().cuda()[None], g_logvar=g["g_logvar"].float().cuda()[None])
        dlv = dlv[1:].cpu().double() * (c["mod_logvar"].abs() < 50)      # clip's mask is applied by the caller (torch.clamp)
        assert rel_l2(dmu[1:], g["d_mod_mu"]) < 1e-5 and rel_l2(dlv, g["d_mod_logvar"]) < 1e-5


@pytest.mark.parametrize("B", [1, 3])
def test_poe_model_level_shapes_vs_oracle(B):
    """The four latent levels of a 128^3 volume, all 15 subsets."""
    from xlstm_hved_b200 import ops
    g = torch.Generator().manual_seed(B)
    for C, d in ((1, 64), (2, 32), (4, 16), (8, 8)):
        mu = torch.cat([torch.zeros(1, B, C, d, d, d), 1.3 * torch.randn(4, B, C, d, d, d, generator=g)])
        lv = torch.cat([torch.zeros(1, B, C, d, d, d), (1.4 * torch.randn(4, B, C, d, d, d, generator=g)).clamp(-50, 50)])
        pm, pl, _, _ = ops.poe_fwd(mu.cuda(), lv.cuda(), restate.SUBSETS_MODALITIES)
        for i, s in enumerate(restate.SUBSETS_MODALITIES):
            a, b = restate.poe(mu.double(), lv.double(), s)
            assert rel_linf(pm[i], a) < TOL and rel_linf(pl[i], b) < TOL


def test_poe_full_backward_with_sampling_and_kld_vs_autograd():
    from xlstm_hved_b200 import ops
    g = torch.Generator().manual_seed(7)
    shape = (2, 2, 6, 6, 6)
    mu = torch.cat([torch.zeros(1, *shape), torch.randn(4, *shape, generator=g)]).double().requires_grad_()
    lv = torch.cat([torch.zeros(1, *shape), torch.randn(4, *shape, generator=g)]).double().requires_grad_()
    subsets = [(0, 1, 2, 3), (1, 3)]
    noise = torch.randn(2, *shape, generator=g).double()
    gz = torch.randn(2, *shape, generator=g).double()
    ks = [0.3, -0.2]
    loss = 0
    for i, s in enumerate(subsets):
        a, b = restate.poe(mu, lv, s)
        z = restate.reparametrize(a, b, noise[i])
        # KL against the prior the reference hands over: slab 0 (loss.py:95-97, 113) -- it gets a direct gradient too
        loss = loss + (z * gz[i]).sum() + ks[i] * (-1.0 + lv[0] - b + (b.exp() + (a - mu[0]) ** 2) / (lv[0].exp() + 1e-8)).sum()
    rmu, rlv = torch.autograd.grad(loss, [mu, lv])
    dmu, dlv = ops.poe_bwd(mu.detach().float().cuda(), lv.detach().float().cuda(), subsets, noise=noise.float().cuda(),
                           g_z=gz.float().cuda(), kld_scale=ks)
    assert rel_l2(dmu, rmu) < 1e-5 and rel_l2(dlv, rlv) < 1e-5


def test_module_level_dropins_forward_and_gradients_vs_oracle():
    """The drop-in callables of xlstm_hved_b200.modules (ProductOfExperts, ProductOfExperts2, reparametrize, compute_KLD)
    with autograd, against autograd through the oracle restatement."""
    import xlstm_hved_b200 as xh
    g = torch.Generator().manual_seed(21)
    shape = (3, 2, 4, 4, 4)
    mu0 = torch.cat([torch.zeros(1, *shape), 1.3 * torch.randn(4, *shape, generator=g)])
    lv0 = torch.cat([torch.zeros(1, *shape), 1.4 * torch.randn(4, *shape, generator=g)])
    drop = torch.tensor([[False, True, False, False], [True, False, False, True], [False, False, False, False]])
    w = torch.randn(shape, generator=g)

    def loss_fn(poe, poe_drop, reparam, kld, mu, lv, noise):
        a, b = poe(mu, lv, (0, 2))
        c, d = poe_drop(mu.clone(), lv, drop.to(mu.device))
        z = reparam(a, b, noise)
        return (z * w.to(mu)).sum() + (c * d).sum() + 0.3 * kld(mu.transpose(1, 0), lv.transpose(1, 0), [14, 3])

    # oracle (fp64, CPU)
    mu64, lv64 = mu0.double().requires_grad_(), lv0.double().requires_grad_()
    torch.manual_seed(5)
    noise = torch.empty(shape).normal_()
    ref = loss_fn(lambda m, l, s: restate.poe(m, l, s), lambda m, l, dr: restate.poe_drop(m, l, dr)[:2],
                  lambda a, b, nz: restate.reparametrize(a, b, nz.double()), restate.compute_kld, mu64, lv64, noise)
    rmu, rlv = torch.autograd.grad(ref, [mu64, lv64])
    # kernels (reparametrize draws its own noise from the global CUDA generator: feed the same noise through the op)
    muc, lvc = mu0.cuda().requires_grad_(), lv0.cuda().requires_grad_()
    from xlstm_hved_b200.modules import _ReparamFunction
    got = loss_fn(xh.ProductOfExperts(), xh.ProductOfExperts2(),
                  lambda a, b, nz: _ReparamFunction.apply(a.contiguous(), b.contiguous(), nz.cuda()), xh.compute_KLD, muc, lvc, noise)
    gmu, glv = torch.autograd.grad(got, [muc, lvc])
    assert abs(got.item() - ref.item()) < 1e-4 * abs(ref.item())
    assert rel_l2(gmu, rmu) < 1e-4 and rel_l2(glv, rlv) < 1e-4


def test_product_of_experts2_mutates_mu_like_the_reference():
    import xlstm_hved_b200 as xh
    c = load_golden("poe.pt")
    mu, lv = _mu5(c)
    mu_c = mu.float().cuda()
    a, b = xh.ProductOfExperts2()(mu_c, lv.float().cuda(), c["drop"].cuda())
    assert rel_linf(a, c["drop_pd_mu"]) < TOL and rel_linf(b, c["drop_pd_logvar"]) < TOL
    assert torch.equal(mu_c.cpu().double(), c["drop_mu_after"].float().double())      # buildingblocks.py:879-881


@pytest.mark.parametrize("n_extra", [0, 3])          # 128-bit and scalar code paths
def test_poe_standard_prior_flag_equals_explicit_zero_prior(n_extra):
    """XHVED_POE_STANDARD_PRIOR: the constant prior (mu = 0, logvar = 0, RA_HVED.py:576-580) is declared instead of read.
    The slab is filled with NaN here to prove that it is not touched; results must equal the explicit-zero-prior call."""
    from xlstm_hved_b200 import ops
    g = torch.Generator().manual_seed(5)
    n = 4 * 1000 + n_extra
    mod_mu, mod_lv = 1.3 * torch.randn(4, n, generator=g), (1.4 * torch.randn(4, n, generator=g)).clamp(-50, 50)
    zeros, nans = torch.zeros(1, n), torch.full((1, n), float("nan"))
    subsets = [(0, 1, 2, 3), (1, 3), (2,)]
    noise, gz = torch.randn(3, n, generator=g).cuda(), torch.randn(3, n, generator=g).cuda()
    ks = [0.3, -0.2, 0.1]
    ref_mu, ref_lv = torch.cat([zeros, mod_mu]).cuda(), torch.cat([zeros, mod_lv]).cuda()
    nan_mu, nan_lv = torch.cat([nans, mod_mu]).cuda(), torch.cat([nans, mod_lv]).cuda()
    a = ops.poe_fwd(ref_mu, ref_lv, subsets, noise=noise, want_kld=True)
    b = ops.poe_fwd(nan_mu, nan_lv, subsets, noise=noise, want_kld=True, standard_prior=True)
    for x, y in zip(a[:3], b[:3]):       # two template instantiations: equal up to FMA contraction
        assert torch.allclose(x, y, rtol=1e-6, atol=1e-7)
    assert torch.allclose(a[3], b[3], rtol=1e-5)          # KL sums: block atomics, summation order is not fixed
    dmu, dlv = ops.poe_bwd(ref_mu, ref_lv, subsets, noise=noise, g_z=gz, kld_scale=ks)
    dmu4, dlv4 = ops.poe_bwd(nan_mu, nan_lv, subsets, noise=noise, g_z=gz, kld_scale=ks, standard_prior=True)
    # the two template instantiations (prior read / prior declared) may contract their FMAs differently: equal to rounding
    close = lambda a, b: torch.allclose(a, b, rtol=1e-5, atol=1e-6)
    assert dmu4.shape == (4, n) and close(dmu[1:], dmu4) and close(dlv[1:], dlv4)


def test_poe_levels_one_launch_equals_per_level_launches():
    """xhved_poe_fwd_levels / _bwd_levels: the four latent levels of a volume in one launch (shapes of SURVEY appendix A,
    plus a ragged level that forces the scalar path) must give exactly what four separate launches give."""
    from xlstm_hved_b200 import ops
    g = torch.Generator().manual_seed(9)
    subsets = [(0, 1, 2, 3), (0, 2)]
    for shapes in ([(2, 1, 16, 16, 16), (2, 2, 8, 8, 8), (2, 4, 4, 4, 4), (2, 8, 2, 2, 2)], [(1, 3, 5, 7), (2, 4, 4, 4)]):
        levels, noises, gzs, scales = [], [], [], []
        for sh in shapes:
            mu = torch.cat([torch.zeros(1, *sh), 1.3 * torch.randn(4, *sh, generator=g)]).cuda()
            lv = torch.cat([torch.zeros(1, *sh), (1.4 * torch.randn(4, *sh, generator=g)).clamp(-50, 50)]).cuda()
            levels.append((mu, lv))
            noises.append(torch.randn(2, *sh, generator=g).cuda())
            gzs.append(torch.randn(2, *sh, generator=g).cuda())
            scales.append([0.3, -0.1])
        for sp in (False, True):
            kld = torch.zeros(len(shapes), 2, device="cuda")
            outs = ops.poe_fwd_levels(levels, subsets, noises=noises, kld_out=kld, standard_prior=sp)
            grads = ops.poe_bwd_levels(levels, subsets, noises=noises, g_zs=gzs, kld_scales=scales, standard_prior=sp)
            for l, (mu, lv) in enumerate(levels):
                pm, pl, z, k1 = ops.poe_fwd(mu, lv, subsets, noise=noises[l], want_kld=True, standard_prior=sp)
                for got, want in zip(outs[l], (pm, pl, z)):      # (a ragged level runs the scalar instantiation)
                    assert torch.allclose(got, want, rtol=1e-6, atol=1e-7)
                assert torch.allclose(kld[l], k1, rtol=1e-5)
                dmu, dlv = ops.poe_bwd(mu, lv, subsets, noise=noises[l], g_z=gzs[l], kld_scale=scales[l], standard_prior=sp)
                # a ragged level forces the scalar instantiation for the whole fused launch: equal up to FMA contraction
                assert torch.allclose(grads[l][0], dmu, rtol=1e-5, atol=1e-6) and torch.allclose(grads[l][1], dlv, rtol=1e-5, atol=1e-6)


def test_poe_full_bench_size_properties():
    """Full bench size (B=32 volumes, 4 levels, 11.1 M latent elements): (1) the fused KL sum equals the KL recomputed from
    the kernel's own outputs; (2) with one modality the product of that expert with the standard prior is reproduced in
    closed form: T = 1/(e^lv + eps) + 1/(1 + eps), mu^ = mu T_m / T; (3) a dropped batch == the complementary subset."""
    from xlstm_hved_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(1)
    B = 32
    for C, d in ((1, 64), (8, 8)):
        sh = (B, C, d, d, d)
        mu = torch.cat([torch.zeros(1, *sh, device="cuda"), 1.3 * torch.randn(4, *sh, device="cuda", generator=g)])
        lv = torch.cat([torch.zeros(1, *sh, device="cuda"), (1.4 * torch.randn(4, *sh, device="cuda", generator=g)).clamp(-50, 50)])
        pm, pl, _, kld = ops.poe_fwd(mu, lv, [(0, 1, 2, 3), (2,)], want_kld=True, standard_prior=True)
        ref_kld = (-1.0 - pl.double() + (pl.double().exp() + pm.double() ** 2)).flatten(1).sum(1)
        assert torch.allclose(kld.double(), ref_kld, rtol=2e-4)
        T = 1.0 / (lv[3].double().exp() + 1e-8) + 1.0 / (1.0 + 1e-8)
        assert rel_linf(pm[1], mu[3].double() / (lv[3].double().exp() + 1e-8) / T) < 1e-5
        assert rel_linf(pl[1], -T.log()) < 1e-5
        drop = torch.zeros(B, 4, dtype=torch.bool, device="cuda")
        drop[:, 0] = True
        drop[:, 3] = True
        dm, dl, _, _ = ops.poe_fwd(mu, lv, [(0, 1, 2, 3)], drop=drop)
        cm, cl, _, _ = ops.poe_fwd(mu, lv, [(1, 2)])
        assert torch.equal(dm, cm) and torch.equal(dl, cl)


def test_compute_kld_reads_the_prior_slab_golden():
    """compute_KLD hands slab 0 of its inputs to KL_divergence as the prior (loss.py:95-97, 113): a NON-standard prior must
    change the result and receive its own gradient (fixture generated from the reference, oracle/make_golden.py)."""
    import xlstm_hved_b200 as xh
    k = load_golden("smvae_extras.pt")["kld_prior"]
    mu = k["mu"].float().cuda().requires_grad_()
    lv = k["logvar"].float().cuda().requires_grad_()
    val = xh.compute_KLD(mu, lv, k["subsets"])
    gm, gl = torch.autograd.grad(val, [mu, lv])
    assert abs(val.item() - k["kld"].item()) < 1e-4 * abs(k["kld"].item())
    assert rel_l2(gm, k["d_mu"]) < 1e-4 and rel_l2(gl, k["d_logvar"]) < 1e-4
    assert rel_linf(gm[:, 0], k["d_mu"][:, 0]) < 1e-3 and rel_linf(gl[:, 0], k["d_logvar"][:, 0]) < 1e-3     # the prior's own gradient
    # and the standard prior still agrees with the declared-constant fast path
    from xlstm_hved_b200 import ops
    mu5 = k["mu"].float().transpose(1, 0).contiguous().cuda()
    lv5 = k["logvar"].float().transpose(1, 0).contiguous().cuda()
    mu5[0] = 0
    lv5[0] = 0
    a = ops.poe_fwd(mu5, lv5, [(0, 1, 2, 3), (1,)], want_kld=True)[3]
    b = ops.poe_fwd(mu5, lv5, [(0, 1, 2, 3), (1,)], want_kld=True, standard_prior=True)[3]
    assert torch.allclose(a, b, rtol=1e-5)


def test_clip_standalone_and_fused_into_the_fusion_golden():
    """clip (RA_HVED.py:749-753) as its own kernel and fused into the PoE launch (XHVED_POE_CLIP): raw logvars with entries
    beyond +-50 give the reference's outputs, and d_logvar vanishes exactly where the reference's clamp gradient does."""
    import xlstm_hved_b200 as xh
    from xlstm_hved_b200 import ops
    cp = load_golden("smvae_extras.pt")["clip_poe"]
    raw = cp["raw_logvar"].float().cuda().requires_grad_()
    y = xh.clip(raw)
    assert torch.equal(y.detach().cpu(), cp["clipped"].float())
    g = torch.randn_like(y)
    (dx,) = torch.autograd.grad(y, raw, g)
    assert torch.equal(dx, g * (raw.detach().abs() <= 50))
    assert torch.isnan(xh.clip(torch.tensor([float("nan"), 60.0, -70.0], device="cuda"))[0])       # NaN passes, like torch.clamp
    sh = cp["mod_mu"].shape[1:]
    z = torch.zeros(1, *sh)
    mu5 = torch.cat([z, cp["mod_mu"].float()]).cuda()
    lv5_raw = torch.cat([z, cp["raw_logvar"].float()]).cuda()
    idxs = sorted(cp["cases"])
    subsets = [restate.SUBSETS_MODALITIES[i] for i in idxs]
    for sp in (False, True):
        (pm, pl, _), = ops.poe_fwd_levels([(mu5, lv5_raw)], subsets, clip=(-50.0, 50.0), standard_prior=sp)
        g_mu = torch.stack([cp["cases"][i]["g_mu"].float() for i in idxs]).cuda()
        g_lv = torch.stack([cp["cases"][i]["g_logvar"].float() for i in idxs]).cuda()
        for j, i in enumerate(idxs):
            r = cp["cases"][i]
            assert rel_linf(pm[j], r["pd_mu"]) < TOL and rel_linf(pl[j], r["pd_logvar"]) < TOL
            one = lambda t: t[j:j + 1].contiguous()
            (dm, dl), = ops.poe_bwd_levels([(mu5, lv5_raw)], [subsets[j]], g_mus=[one(g_mu)], g_logvars=[one(g_lv)], clip=(-50.0, 50.0),
                                           standard_prior=sp)
            dm, dl = (dm, dl) if sp else (dm[1:], dl[1:])
            assert rel_l2(dm, r["d_mod_mu"]) < 1e-5 and rel_l2(dl, r["d_raw_logvar"]) < 1e-5
            outside = (cp["raw_logvar"].abs() > 50)
            assert (dl.cpu()[outside] == 0).all()


def test_zero_layer_golden_and_patch_target():
    """ZeroLayerF (buildingblocks.py:308-323; call sites RA_HVED.py:559, U_Hemis.py:42): forward and backward."""
    import xlstm_hved_b200 as xh
    zl = load_golden("smvae_extras.pt")["zero_layer"]
    x = zl["x"].float().cuda().requires_grad_()
    y = xh.modules.ZeroLayerF.apply(x, zl["alpha"].cuda())
    (dx,) = torch.autograd.grad(y, x, zl["gy"].float().cuda())
    assert torch.equal(y.detach().cpu(), zl["y"].float()) and torch.equal(dx.cpu(), zl["dx"].float())
    # ragged row length (scalar path) and a mask held on the host
    x2 = torch.randn(4, 7, device="cuda")
    m2 = torch.tensor([False, True, True, False])
    assert torch.equal(xh.modules.ZeroLayerF.apply(x2, m2), restate.zero_rows(x2.cpu(), m2).cuda())


def test_dice_loss_fused_golden_and_large():
    """compute_per_channel_dice / DiceLoss (loss.py:188-209, 257-301) as one fused pass: the reference fixture (value, per-channel
    coefficients, gradient, a channel with the clamp active), fp16 inputs as under the reference's autocast, and a full-size
    segmentation output against the oracle."""
    import xlstm_hved_b200 as xh
    c = load_golden("losses.pt")
    p = c["p"].float().cuda().requires_grad_()
    t = c["t"].float().cuda()
    val = xh.DiceLoss()(p, t)
    (dp,) = torch.autograd.grad(val, p)
    assert abs(val.item() - c["loss"].item()) < 1e-5
    assert rel_linf(xh.modules.compute_per_channel_dice(p.detach(), t), c["per_channel"]) < 1e-5 and rel_l2(dp, c["dp"]) < 1e-5
    assert dp[:, 3].abs().max().item() == 0.0
    w = torch.tensor([0.5, 1.0, 2.0, 1.0], device="cuda")
    assert rel_linf(xh.modules.compute_per_channel_dice(p.detach(), t, weight=w), restate.dice_per_channel(c["p"], c["t"], weight=w.cpu().double())) < 1e-5
    h = xh.modules.compute_per_channel_dice(p.detach().half(), t.half())
    assert h.dtype == torch.float16 and rel_linf(h.float(), c["per_channel"]) < 2e-3
    g = torch.Generator(device="cuda").manual_seed(2)
    P = torch.rand(2, 3, 128, 128, 128, device="cuda", generator=g).requires_grad_()
    T = (torch.rand(2, 3, 128, 128, 128, device="cuda", generator=g) > 0.5).float()
    v = xh.DiceLoss()(P, T)
    (dP,) = torch.autograd.grad(v, P)
    P64 = P.detach().double().requires_grad_()
    v64 = restate.dice_loss(P64, T.double())
    (dP64,) = torch.autograd.grad(v64, P64)
    assert abs(v.item() - v64.item()) < 1e-5 and rel_l2(dP, dP64) < 1e-5
