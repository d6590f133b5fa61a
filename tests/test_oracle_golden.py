"""The oracle restatement (oracle/restate.py) against fixtures produced by the REAL
reference (oracle/make_golden.py).  CPU only; this is what pins the oracle."""
import pytest
import torch

from conftest import load_golden, rel_l2, rel_linf
from oracle import restate


@pytest.fixture(scope="module")
def cell():
    return load_golden("cell.pt")


@pytest.mark.parametrize("name", ["bottleneck_s320_dh16", "randn_f4_s200_dh16", "randn_f0_s256_dh32",
                                  "randn_fm2_s130_dh8", "randn_f4_s256_dh64"])
def test_cell_parallel_and_chunkwise_match_reference(cell, name):
    c = cell[name]
    q, k, v, ig, fg = [c[n].double() for n in ("q", "k", "v", "ig", "fg")]
    assert rel_linf(restate.mlstm_parallel(q, k, v, ig, fg), c["h"]) < 1e-12
    for L in (32, 128):
        assert rel_linf(restate.mlstm_chunkwise(q, k, v, ig, fg, chunk=L), c["h"]) < 1e-10
    h32 = restate.mlstm_parallel_reference_cost(*[t.float() for t in (q, k, v, ig, fg)])
    assert rel_linf(h32, c["h_fp32_ref"]) < 1e-5


@pytest.mark.parametrize("name", ["bottleneck_s320_dh16", "randn_f4_s200_dh16", "randn_fm2_s130_dh8"])
def test_cell_backward_matches_reference_autograd(cell, name):
    c = cell[name]
    q, k, v, ig, fg, dh = [c[n].double() for n in ("q", "k", "v", "ig", "fg", "dh")]
    got = restate.mlstm_backward(q, k, v, ig, fg, dh, through_max=True)
    for g, n in zip(got, ("dq", "dk", "dv", "dig", "dfg")):
        assert rel_linf(g, c[n]) < 1e-6, n          # fixtures store grads as float32
    # dropping the gradient through the row max (what the kernels do) is a ~1e-5 effect here
    got = restate.mlstm_backward(q, k, v, ig, fg, dh, through_max=False)
    for g, n in zip(got, ("dq", "dk", "dv", "dig", "dfg")):
        assert rel_linf(g, c[n]) < 1e-3, n
    # autograd through the chunkwise restatement gives the same gradients
    leaves = [t.clone().requires_grad_() for t in (q, k, v, ig, fg)]
    h = restate.mlstm_chunkwise(*leaves, chunk=64)
    for g, n in zip(torch.autograd.grad(h, leaves, dh), ("dq", "dk", "dv", "dig", "dfg")):
        assert rel_linf(g, c[n]) < 1e-6, n


def test_stabiliser_scan_is_row_max(cell):
    c = cell["randn_f0_s256_dh32"]
    q, k, v, ig, fg = [c[n].double() for n in ("q", "k", "v", "ig", "fg")]
    _, m, _ = restate.mlstm_parallel(q, k, v, ig, fg, return_aux=True)
    assert (restate.mlstm_stabiliser_scan(ig, fg) - m).abs().max() < 1e-12


@pytest.mark.parametrize("name", ["dim32_s200_fwd", "dim32_s200_rev", "dim16_s150_fwd", "dim64_s140_rev", "dim128_s160_rev"])
def test_vil_block_matches_reference(name):
    c = load_golden("vil_block_wide.pt" if name.startswith("dim128") else "vil_block.pt")[name]
    p = {k: v.double() for k, v in c["state_dict"].items()}
    x = c["x"].clone().requires_grad_()
    leaves = {k: v.clone().requires_grad_() for k, v in p.items()}
    for cellfn in (restate.mlstm_parallel, lambda *a: restate.mlstm_chunkwise(*a, chunk=64)):
        y = restate.vil_block(x, leaves, reverse=c["reverse"], cell=cellfn)
        assert rel_linf(y, c["y"]) < 1e-10
    grads = torch.autograd.grad(y, [x] + [leaves[n] for n in c["param_grads"]], c["dy"])
    assert rel_linf(grads[0], c["dx"]) < 1e-8
    for g, n in zip(grads[1:], c["param_grads"]):
        assert rel_linf(g, c["param_grads"][n]) < 1e-8, n


def test_vil_wrapper_matches_reference():
    c = load_golden("vil_wrapper.pt")
    p = {k: v.double() for k, v in c["state_dict"].items()}
    y = restate.vil_wrapper(c["x"].double(), p)
    assert rel_linf(y, c["y"]) < 1e-10
    assert y.shape == c["x"].shape


def _mu5(c):
    z = torch.zeros(1, *c["mod_mu"].shape[1:], dtype=c["mod_mu"].dtype)
    return torch.cat([z, c["mod_mu"]], 0), torch.cat([z, restate.clip_logvar(c["mod_logvar"])], 0)


def test_poe_all_subsets_match_reference():
    c = load_golden("poe.pt")
    mu, lv = _mu5(c)
    assert [list(s) for s in restate.SUBSETS_MODALITIES] == c["subsets"]
    for i, subset in enumerate(restate.SUBSETS_MODALITIES):
        a, b = restate.poe(mu, lv, subset)
        assert rel_linf(a, c["pd_mu"][i]) < 1e-13 and rel_linf(b, c["pd_logvar"][i]) < 1e-13


def test_poe_backward_drop_reparam_kld_match_reference():
    c = load_golden("poe.pt")
    mu, lv = _mu5(c)
    for idx in (14, 5):
        g = c[f"grad_{idx}"]
        dmu, dlv = restate.poe_backward(mu, lv, restate.SUBSETS_MODALITIES[idx], g["g_mu"], g["g_logvar"])
        dlv = dlv * (c["mod_logvar"].abs() < 50)         # clip's gradient mask (RA_HVED.py:580, 749-753)
        assert rel_linf(dmu, g["d_mod_mu"]) < 1e-12 and rel_linf(dlv, g["d_mod_logvar"]) < 1e-12
    a, b, mu_after = restate.poe_drop(mu, lv, c["drop"])
    assert rel_linf(a, c["drop_pd_mu"]) < 1e-13 and rel_linf(b, c["drop_pd_logvar"]) < 1e-13
    assert torch.equal(mu_after, c["drop_mu_after"])
    z = restate.reparametrize(c["pd_mu"][14].float(), c["pd_logvar"][14].float(), c["reparam_noise"])
    assert rel_linf(z, c["reparam_z"]) < 1e-6
    mu_b5, lv_b5 = mu.transpose(1, 0), lv.transpose(1, 0)
    assert abs(restate.compute_kld(mu_b5, lv_b5, [14]).item() - c["kld_14"].item()) < 1e-12
    assert abs(restate.compute_kld(mu_b5, lv_b5, [3, 7, 12]).item() - c["kld_3_7_12"].item()) < 1e-12


def test_model_boundary_fixture_is_consistent():
    c = load_golden("model_boundary.pt")
    assert len(c["poe"]) == 4
    for rec in c["poe"]:
        a, b = restate.poe(rec["mu"].double(), rec["logvar"].double(), tuple(rec["subset"]))
        assert rel_linf(a, rec["pd_mu"]) < 1e-5 and rel_linf(b, rec["pd_logvar"]) < 1e-5
    p = {k: v.double() for k, v in c["vil_state_dict"].items()}
    y = restate.vil_wrapper(c["vil"]["x"].double(), p)
    assert rel_linf(y, c["vil"]["y"]) < 1e-4     # the fixture is the reference's fp32 run


def test_smvae_extras_general_prior_clip_zero_layer_match_reference():
    """Round-2 fixtures: compute_KLD with a non-standard prior in slab 0 (loss.py:95-97, 113) and its gradients, clip's
    gradient mask through the fusion (RA_HVED.py:580, 749-753), ZeroLayerF forward and backward (buildingblocks.py:308-323)."""
    c = load_golden("smvae_extras.pt")
    k = c["kld_prior"]
    mu, lv = k["mu"].clone().requires_grad_(), k["logvar"].clone().requires_grad_()
    val = restate.compute_kld(mu, lv, k["subsets"])
    gm, gl = torch.autograd.grad(val, [mu, lv])
    assert abs(val.item() - k["kld"].item()) < 1e-12 * abs(k["kld"].item())
    assert rel_linf(gm, k["d_mu"]) < 1e-12 and rel_linf(gl, k["d_logvar"]) < 1e-12
    cp = c["clip_poe"]
    assert torch.equal(restate.clip_logvar(cp["raw_logvar"]), cp["clipped"]) and (cp["raw_logvar"].abs() > 50).sum() > 10
    z = torch.zeros(1, *cp["mod_mu"].shape[1:], dtype=torch.float64)
    mu5, lv5 = torch.cat([z, cp["mod_mu"]]), torch.cat([z, cp["clipped"]])
    for idx, r in cp["cases"].items():
        subset = restate.SUBSETS_MODALITIES[idx]
        a, b = restate.poe(mu5, lv5, subset)
        assert rel_linf(a, r["pd_mu"]) < 1e-13 and rel_linf(b, r["pd_logvar"]) < 1e-13
        dmu, dlv = restate.poe_backward(mu5, lv5, subset, r["g_mu"], r["g_logvar"])
        dlv = dlv * (cp["raw_logvar"].abs() <= 50)
        assert rel_linf(dmu, r["d_mod_mu"]) < 1e-12 and rel_linf(dlv, r["d_raw_logvar"]) < 1e-12
    zl = c["zero_layer"]
    assert torch.equal(restate.zero_rows(zl["x"], zl["alpha"]), zl["y"]) and torch.equal(restate.zero_rows(zl["gy"], zl["alpha"]), zl["dx"])


def test_dice_loss_matches_reference():
    """DiceLoss / compute_per_channel_dice (loss.py:188-209, 257-301), value and gradient, incl. a channel where the clamp is active."""
    c = load_golden("losses.pt")
    p = c["p"].clone().requires_grad_()
    val = restate.dice_loss(p, c["t"])
    (dp,) = torch.autograd.grad(val, p)
    assert abs(val.item() - c["loss"].item()) < 1e-14
    assert rel_linf(restate.dice_per_channel(c["p"], c["t"]), c["per_channel"]) < 1e-14 and rel_linf(dp, c["dp"]) < 1e-13
    assert c["per_channel"][3].item() == 0.0 and c["dp"][:, 3].abs().max().item() == 0.0


# ------------------------------------------------------------------ conv path: InstanceNorm3d / BatchNorm3d + LeakyReLU (K6)
@pytest.fixture(scope="module")
def conv_norm():
    return load_golden("conv_norm.pt")


def test_instance_norm_lrelu_matches_reference_single_conv(conv_norm):
    c = conv_norm["single_conv_ilc"]
    x = c["x"].clone().requires_grad_()
    mid = restate.instance_norm_lrelu(x, slope=c["slope"])
    assert rel_linf(mid, c["mid"]) < 1e-12
    (dx,) = torch.autograd.grad(mid, x, c["gm"])
    assert rel_linf(dx, c["dx_mid"]) < 1e-10


def test_instance_norm_lrelu_matches_reference_basic_conv(conv_norm):
    c = conv_norm["basic_conv"]
    w = c["state_dict"]["conv.weight"]
    y = restate.instance_norm_lrelu(torch.nn.functional.conv3d(c["x"], w, padding=1), slope=c["slope"])
    assert rel_linf(y, c["y"]) < 1e-12


def test_batch_norm_matches_reference_duse_attention(conv_norm):
    c = conv_norm["duse_attention"]
    sd0, sd1 = c["state_dict_before"], c["state_dict_after_train"]
    w, b = sd0["bn_fuse_ch1.weight"], sd0["bn_fuse_ch1.bias"]
    y, rm, rv = restate.batch_norm_lrelu(c["bn1_train_in"], w, b, sd0["bn_fuse_ch1.running_mean"], sd0["bn_fuse_ch1.running_var"],
                                         training=True)
    assert rel_linf(y, c["bn1_train_out"]) < 1e-12
    assert rel_linf(rm, sd1["bn_fuse_ch1.running_mean"]) < 1e-12 and rel_linf(rv, sd1["bn_fuse_ch1.running_var"]) < 1e-12
    y, _, _ = restate.batch_norm_lrelu(c["bn1_eval_in"], w, b, sd1["bn_fuse_ch1.running_mean"], sd1["bn_fuse_ch1.running_var"],
                                       training=False)
    assert rel_linf(y, c["bn1_eval_out"]) < 1e-12


def test_atten_module2_matches_reference():
    c = load_golden("atten_module2.pt")
    seg_x, enc_x = c["seg_x"].clone().requires_grad_(), c["enc_x"].clone().requires_grad_()
    p = {k: v.clone().requires_grad_() for k, v in c["state_dict"].items()}
    y = restate.atten_module2(seg_x, enc_x, p)
    assert rel_linf(y, c["y"]) < 1e-12
    grads = torch.autograd.grad(y, [seg_x, enc_x] + [p[k] for k in c["param_grads"]], c["gy"])
    assert rel_linf(grads[0], c["d_seg_x"]) < 1e-10 and rel_linf(grads[1], c["d_enc_x"]) < 1e-10
    for g, k in zip(grads[2:], c["param_grads"]):
        assert rel_linf(g, c["param_grads"][k]) < 1e-10, k


def test_depthwise_basic_conv_matches_reference():
    c = load_golden("dwconv3.pt")
    x, w = c["x"].clone().requires_grad_(), c["state_dict"]["conv.weight"].clone().requires_grad_()
    y, conv_out = restate.basic_conv_depthwise(x, w)
    assert rel_linf(conv_out, c["conv_out"]) < 1e-12 and rel_linf(y, c["y"]) < 1e-11
    dx, dw = torch.autograd.grad(y, [x, w], c["gy"])
    assert rel_linf(dx, c["dx"]) < 1e-9 and rel_linf(dw, c["dconv_weight"]) < 1e-9
