"""End-to-end drop-in parity: the UNMODIFIED reference model (baseline/_ref or /root/reference, when present) next
to the same model with patch_model() applied, on the GPU.  North-star bar: identical (p>0.5) mask per channel and
identical channel argmax on >= 99.9% of voxels (the model outputs sigmoid probabilities of 3 nested regions,
RA_HVED.py:483-484); PoE mu/logvar within 1e-3.  Skipped where no reference tree exists; the boundary fixtures
(tests/golden/model_boundary.pt) cover the same tensors without it."""
import contextlib
import io

import pytest
import torch

from conftest import load_golden, rel_l2, rel_linf
from oracle import ref_loader

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _fp32_reference():
    """The bar is measured against the reference computing in fp32.  PyTorch runs cuDNN convolutions in TF32 by default; at that
    precision the STOCK model disagrees with its own fp64 evaluation on 0.27 % of the voxels (argmax, 64x96x64, subset 7;
    0.04 % in fp32: tools/diag_conv_norm.py), which would make "99.9 % identical to stock" a statement about that noise."""
    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def test_model_boundary_fixture_vil_and_poe():
    """Tensors recorded at the hot-path boundary of a real reference forward (48^3 volume, subset 12)."""
    import xlstm_hved_b200 as xh
    c = load_golden("model_boundary.pt")
    wrap = xh.ViLLayer3D(dim=32).cuda()
    wrap.load_state_dict(c["vil_state_dict"], strict=True)
    with torch.no_grad():
        y = wrap(c["vil"]["x"].cuda())
    br, br_ref = y.cpu() - c["vil"]["x"], c["vil"]["y"] - c["vil"]["x"]
    assert rel_l2(br, br_ref) < 2e-2
    poe = xh.ProductOfExperts()
    for rec in c["poe"]:
        a, b = poe(rec["mu"].cuda(), rec["logvar"].cuda(), tuple(rec["subset"]))
        assert rel_linf(a, rec["pd_mu"]) < 1e-3 and rel_linf(b, rec["pd_logvar"]) < 1e-3


@pytest.fixture(scope="module")
def model():
    if ref_loader.find_reference() is None:
        pytest.skip("no reference tree on this machine (baseline/_ref absent)")
    m = ref_loader.build_model(f_maps=4, seed=1)
    return m.cuda().eval()


@pytest.mark.parametrize("size,subset", [((128, 128, 128), 14), ((64, 96, 64), 7)])
def test_full_model_segmentation_parity(model, size, subset):
    import xlstm_hved_b200 as xh
    ns = ref_loader.load_reference()
    torch.manual_seed(5)
    x = torch.rand(1, 4, *size, device="cuda")
    present = ns.RA_HVED.SUBSETS_MODALITIES[subset]
    for m in range(4):
        if m not in present:
            x[:, m] = 0                                     # evaluation.py:306-307
    with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
        seg_ref, _ = model(x, [subset], valid=True)
        counts = xh.patch_model(model)
        try:
            seg_new, _ = model(x, [subset], valid=True)
        finally:
            xh.unpatch_model(model)
    assert counts["ViLBlock"] == 1 and counts["ProductOfExperts"] == 1
    same_mask = ((seg_ref > 0.5) == (seg_new > 0.5)).float().mean().item()
    same_arg = (seg_ref.argmax(1) == seg_new.argmax(1)).float().mean().item()
    print(size, subset, "mask agreement", same_mask, "argmax agreement", same_arg, "max |dp|", (seg_ref - seg_new).abs().max().item())
    assert same_mask >= 0.999 and same_arg >= 0.999
    assert counts["InstanceNorm3d"] == 80 and counts["BatchNorm3d"] == 18 and counts["fused_LeakyReLU"] == 80
    if size[0] <= 64:
        # against the truth: the stock model evaluated in fp64.  The patched model (ViL / S-MVAE kernels + the conv path's fused
        # normalisation) must agree with it at least as well as the stock fp32 model does.
        import copy
        m64 = copy.deepcopy(model).double()
        with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
            seg64, _ = m64(x.double(), [subset], valid=True)
        del m64
        agree = lambda p: (((p.double() > 0.5) == (seg64 > 0.5)).double().mean().item(),
                           (p.argmax(1) == seg64.argmax(1)).double().mean().item())
        (m_ref, a_ref), (m_new, a_new) = agree(seg_ref), agree(seg_new)
        print("vs fp64: stock", m_ref, a_ref, "patched", m_new, a_new)
        assert m_new >= m_ref - 3e-4 and a_new >= a_ref - 3e-4
        assert m_new >= 0.999 and a_new >= 0.999


def test_full_model_training_gradients_parity(model):
    """seg + recon + KL losses, sampling on: parameter gradients of the patched model vs the stock one."""
    import xlstm_hved_b200 as xh
    ns = ref_loader.load_reference()
    model.train()
    try:
        x = torch.rand(1, 4, 64, 64, 64, device="cuda", generator=torch.Generator("cuda").manual_seed(9))

        def run():
            model.zero_grad(set_to_none=True)
            torch.manual_seed(11)                          # same epsilon draws on both paths
            with contextlib.redirect_stdout(io.StringIO()):
                seg, (mu_list, lv_list), recon = model(x, [14], recon=True)
                kld = sum(ns.loss.compute_KLD(mu_list[l], lv_list[l], [14]) for l in range(4)) / 4
            loss = seg.mean() + 0.2 * ((recon[0] - x) ** 2).mean() + 0.2 * kld
            loss.backward()
            return loss.item(), {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}

        loss_ref, g_ref = run()
        _, g_ref2 = run()                                  # run-to-run noise of the reference itself (atomics)
        xh.patch_model(model)
        try:
            loss_new, g_new = run()
        finally:
            xh.unpatch_model(model)
        assert abs(loss_ref - loss_new) < 1e-3 * abs(loss_ref)
        assert set(g_ref) == set(g_new)
        worst, checked, num, den = ("", 0.0), 0, 0.0, 0.0
        gmax = max(g.norm().item() for g in g_ref.values())
        for n in g_ref:
            # gradients that are mathematically zero (conv biases in front of InstanceNorm ...) are pure rounding
            # noise even between two runs of the stock model: only parameters the reference reproduces are compared
            # (conv biases in front of InstanceNorm have |g| ~ 1e-9 of the largest tensor: rounding residue, skipped)
            if g_ref[n].norm().item() < 1e-5 * gmax or rel_l2(g_ref2[n], g_ref[n]) > 1e-3:
                continue
            e = rel_l2(g_new[n], g_ref[n])
            if e > worst[1]:
                worst = (n, e)
            checked += 1
            num += (g_new[n].double() - g_ref[n].double()).pow(2).sum().item()
            den += g_ref[n].double().pow(2).sum().item()
            # single bias-like tensors are sums with heavy cancellation: bound them loosely, the whole gradient tightly
            assert e < 1.5e-1, (n, e)
        total = (num / den) ** 0.5
        vil = [n for n in g_ref if n.startswith("mViL.vil.")]
        print("loss", loss_ref, loss_new, "parameters checked", checked, "whole-gradient rel_l2", total, "worst tensor", worst)
        assert len(vil) == 14 and checked > 100
        assert total < 2e-2
        for n in vil:
            assert rel_l2(g_new[n], g_ref[n]) < 5e-2, n
    finally:
        model.eval()


def test_graphed_step_replays_forward_and_backward():
    """xh.GraphedStep: a ViLBlock forward+backward captured once must give, on refilled static inputs, what eager calls give."""
    import xlstm_hved_b200 as xh
    torch.manual_seed(4)
    blk = xh.ViLBlock(32, xh.SequenceTraversal.ROWWISE_FROM_BOT_RIGHT).cuda()
    with torch.no_grad():
        for p in blk.parameters():
            p.add_(0.05 * torch.randn_like(p))
    x_static = torch.zeros(2, 300, 32, device="cuda")
    gy = torch.randn(2, 300, 32, device="cuda")

    def step():
        x = x_static.detach().requires_grad_()
        for p in blk.parameters():
            p.grad = None
        y = blk(x)
        y.backward(gy)
        return y.detach(), x.grad, blk.layer.proj_up.weight.grad

    graphed = xh.GraphedStep(step)
    for seed in (1, 2):
        x_static.copy_(torch.randn(2, 300, 32, generator=torch.Generator().manual_seed(seed)))
        y_g, dx_g, dw_g = (t.clone() for t in graphed())
        y_e, dx_e, dw_e = step()
        assert torch.equal(y_g, y_e) and torch.equal(dx_g, dx_e)
        assert rel_l2(dw_g, dw_e) < 1e-5          # parameter gradients: atomics, order not fixed


def test_all_subsets_in_one_batch_matches_15_forwards(model):
    """xh.all_subsets_forward: the 15 masked copies of a volume as ONE batch through the patched model (per-sample drop mask,
    ProductOfExperts2 path) against the reference's own evaluation loop on the STOCK model (15 forwards, test.py:78-102)."""
    import xlstm_hved_b200 as xh
    ns = ref_loader.load_reference()
    torch.manual_seed(9)
    x = torch.rand(1, 4, 64, 64, 64, device="cuda")
    ref = []
    with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
        for idx, present in enumerate(ns.RA_HVED.SUBSETS_MODALITIES):
            xm = x.clone()
            for m in range(4):
                if m not in present:
                    xm[:, m] = 0
            seg, _ = model(xm, [idx], valid=True)
            ref.append(seg)
    ref = torch.stack(ref)
    xh.patch_model(model)
    try:
        got = xh.all_subsets_forward(model, x)
        got8 = xh.all_subsets_forward(model, x, max_batch=8)
    finally:
        xh.unpatch_model(model)
    assert got.shape == ref.shape == (15, 1, 3, 64, 64, 64)
    same_mask = ((ref > 0.5) == (got > 0.5)).float().mean().item()
    same_arg = (ref.argmax(2) == got.argmax(2)).float().mean().item()
    print("all subsets in one batch: same (p>0.5)", same_mask, "same argmax", same_arg, "max abs diff", (ref - got).abs().max().item())
    assert same_mask >= 0.999 and same_arg >= 0.999
    assert ((got8 > 0.5) == (got > 0.5)).float().mean().item() >= 0.999          # (cuDNN picks algorithms per batch size)


def test_graphed_subsets_forward_replays_the_patched_model(model):
    """xh.GraphedSubsetsForward: the patched evaluation forward recorded once per shape and replayed on refilled static buffers must
    give what the eager batched forward gives -- first use (recording), a second input (pure replay), another subset list (a second
    graph) -- and the stock subset-index forward's masks."""
    import xlstm_hved_b200 as xh
    ns = ref_loader.load_reference()
    g = torch.Generator(device="cuda").manual_seed(21)
    xs = [torch.rand(1, 4, 64, 64, 64, device="cuda", generator=g) for _ in range(2)]
    with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
        x7 = xs[1].clone()
        for m in range(4):
            if m not in ns.RA_HVED.SUBSETS_MODALITIES[7]:
                x7[:, m] = 0
        stock7, _ = model(x7, [7], valid=True)
    xh.patch_model(model)
    try:
        fwd = xh.GraphedSubsetsForward(model)
        for x in xs:
            got = fwd(x, subsets=[14, 7])
            ref = xh.all_subsets_forward(model, x, subsets=[14, 7])
            assert got.shape == ref.shape == (2, 1, 3, 64, 64, 64)
            assert (got - ref).abs().max().item() < 1e-4
        assert len(fwd._graphs) == 1
        assert ((got[1] > 0.5) == (stock7 > 0.5)).float().mean().item() >= 0.999
        one = fwd(xs[0], subsets=[3])
        assert len(fwd._graphs) == 2 and (one - xh.all_subsets_forward(model, xs[0], subsets=[3])).abs().max().item() < 1e-4
    finally:
        xh.unpatch_model(model)
