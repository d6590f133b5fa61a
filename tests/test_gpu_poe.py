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
        loss = loss + (z * gz[i]).sum() + ks[i] * (-1.0 - b + (b.exp() + a * a) / (1 + 1e-8)).sum()
    rmu, rlv = torch.autograd.grad(loss, [mu, lv])
    dmu, dlv = ops.poe_bwd(mu.detach().float().cuda(), lv.detach().float().cuda(), subsets, noise=noise.float().cuda(),
                           g_z=gz.float().cuda(), kld_scale=ks)
    assert rel_l2(dmu, rmu) < 1e-5 and rel_l2(dlv, rlv) < 1e-5
