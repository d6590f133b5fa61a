"""GPU parity of K6 (InstanceNorm3d / BatchNorm3d fused with LeakyReLU, csrc/norm_act.cu) through the C ABI: against fixtures the
REAL reference modules produced (tests/golden/conv_norm.pt: SingleConv 'ilc', BasicConv, DuSEAttention) and against the fp64
oracle (oracle/restate.py) on seeded inputs: ragged / unaligned planes, one-chunk and many-chunk planes, fp16 / bf16."""
import pytest
import torch
import torch.nn as nn

from conftest import load_golden, rel_l2, rel_linf
from oracle import restate

pytestmark = pytest.mark.gpu

TOL_F32 = 2e-5          # fp32 kernel vs the fp64 reference values, relative to the largest magnitude
TOL_G32 = 1e-4


@pytest.fixture(scope="module")
def golden():
    return load_golden("conv_norm.pt")


@pytest.fixture(autouse=True)
def _exact_convs():
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False          # the convolutions around the kernel stay on PyTorch: keep them fp32-exact
    yield
    torch.backends.cudnn.allow_tf32 = old


def test_instance_norm_lrelu_golden(golden):
    from xlstm_hved_b200 import modules
    c = golden["single_conv_ilc"]
    x = c["x"].float().cuda().requires_grad_()
    y = modules.instance_norm_act(x, slope=c["slope"])
    assert rel_linf(y, c["mid"]) < TOL_F32
    (dx,) = torch.autograd.grad(y, x, c["gm"].float().cuda())
    assert rel_linf(dx, c["dx_mid"]) < TOL_G32


class BasicConv(nn.Module):
    """Structure of buildingblocks.BasicConv (buildingblocks.py:11-31): the patch recognises it by class name."""

    def __init__(self, i, o, k, padding):
        super().__init__()
        self.conv = nn.Conv3d(i, o, k, padding=padding, bias=False)
        self.norm = nn.InstanceNorm3d(o)
        self.relu = nn.LeakyReLU(negative_slope=1e-2, inplace=True)

    def forward(self, x):
        return self.relu(self.norm(self.conv(x)))


def test_patched_single_conv_and_basic_conv_golden(golden):
    """nn.Sequential(instancenorm, LeakyReLU, conv) and BasicConv with the reference's weights after patch_model: the norm runs on
    the kernel, the LeakyReLU is fused away, values and gradients match the reference; unpatch restores the classes."""
    import xlstm_hved_b200 as xh
    c = golden["single_conv_ilc"]
    sc = nn.Sequential()
    sc.add_module("instancenorm", nn.InstanceNorm3d(4))
    sc.add_module("LeakyReLU", nn.LeakyReLU(negative_slope=1e-2, inplace=True))
    sc.add_module("conv", nn.Conv3d(4, 6, 3, padding=1))
    sc.load_state_dict({k: v.float() for k, v in c["state_dict"].items()}, strict=True)
    sc.cuda()
    counts = xh.patch_model(sc)
    assert counts["InstanceNorm3d"] == 1 and counts["fused_LeakyReLU"] == 1
    assert isinstance(sc.LeakyReLU, nn.LeakyReLU) and sc.instancenorm.fused_slope == pytest.approx(0.01)
    x = c["x"].float().cuda().requires_grad_()
    y = sc(x)
    assert rel_linf(y, c["y"]) < 5e-5
    grads = torch.autograd.grad(y, [x, sc.conv.weight, sc.conv.bias], c["gy"].float().cuda())
    for g, n in zip(grads, ("dx", "dconv_weight", "dconv_bias")):
        assert rel_linf(g, c[n]) < 2e-4, n
    xh.unpatch_model(sc)
    assert type(sc.instancenorm) is nn.InstanceNorm3d and type(sc.LeakyReLU) is nn.LeakyReLU
    assert "fused_slope" not in sc.instancenorm.__dict__
    assert rel_linf(sc(x), c["y"]) < 5e-5                       # and the stock path agrees too

    c = golden["basic_conv"]
    bc = BasicConv(3, 5, 3, 1)
    bc.load_state_dict({k: v.float() for k, v in c["state_dict"].items()}, strict=True)
    bc.cuda()
    counts = xh.patch_model(bc)
    assert counts["InstanceNorm3d"] == 1 and counts["fused_LeakyReLU"] == 1
    x = c["x"].float().cuda().requires_grad_()
    y = bc(x)
    assert rel_linf(y, c["y"]) < 5e-5
    grads = torch.autograd.grad(y, [x, bc.conv.weight], c["gy"].float().cuda())
    assert rel_linf(grads[0], c["dx"]) < 5e-4 and rel_linf(grads[1], c["dconv_weight"]) < 5e-4


def test_batch_norm_golden_train_eval_and_running_stats(golden):
    """The first BatchNorm3d of the reference's DuSEAttention: train mode (batch statistics, running statistics moved) and
    eval mode (frozen statistics), on the tensors the real module saw."""
    import xlstm_hved_b200 as xh
    c = golden["duse_attention"]
    sd0, sd1 = c["state_dict_before"], c["state_dict_after_train"]
    bn = xh.modules.BatchNorm3d(4)
    bn.load_state_dict({k.split(".", 1)[1]: v.float() for k, v in sd0.items() if k.startswith("bn_fuse_ch1.")}, strict=True)
    bn.cuda().train()
    y = bn(c["bn1_train_in"].float().cuda())
    assert rel_linf(y, c["bn1_train_out"]) < TOL_F32
    assert rel_linf(bn.running_mean, sd1["bn_fuse_ch1.running_mean"]) < 1e-5
    assert rel_linf(bn.running_var, sd1["bn_fuse_ch1.running_var"]) < 1e-5
    assert int(bn.num_batches_tracked) == int(sd1["bn_fuse_ch1.num_batches_tracked"])
    bn.eval()
    y = bn(c["bn1_eval_in"].float().cuda())
    assert rel_linf(y, c["bn1_eval_out"]) < TOL_F32


@pytest.mark.parametrize("training", [True, False])
@pytest.mark.parametrize("slope", [1.0, 0.2])
def test_batch_norm_backward_vs_oracle(training, slope):
    """dx, dgamma, dbeta in both modes (frozen statistics: dx takes no correction terms), several samples, ragged planes."""
    from xlstm_hved_b200 import modules
    g = torch.Generator().manual_seed(21)
    x = (torch.randn(3, 5, 7, 9, 11, generator=g) * 2 + 1).double()
    w, b = (1 + 0.3 * torch.randn(5, generator=g)).double(), (0.2 * torch.randn(5, generator=g)).double()
    rm, rv = (0.5 * torch.randn(5, generator=g)).double(), (1 + torch.rand(5, generator=g)).double()
    gy = torch.randn(x.shape, generator=g).double()
    leaves = [t.clone().requires_grad_() for t in (x, w, b)]
    y_ref, _, _ = restate.batch_norm_lrelu(leaves[0], leaves[1], leaves[2], rm, rv, training=training, slope=slope)
    ref = torch.autograd.grad(y_ref, leaves, gy)
    cu = [t.float().cuda().requires_grad_() for t in (x, w, b)]
    y = modules.batch_norm_act(cu[0], cu[1], cu[2], rm.float().cuda(), rv.float().cuda(), training, 0.1, 1e-5, slope)
    assert rel_linf(y, y_ref) < TOL_F32
    got = torch.autograd.grad(y, cu, gy.float().cuda())
    for a, r, n in zip(got, ref, ("dx", "dgamma", "dbeta")):
        assert rel_linf(a, r) < TOL_G32, n


@pytest.mark.parametrize("shape", [(1, 4, 16, 16, 16),        # one chunk per plane: single launch
                                   (2, 3, 5, 7, 3),           # tiny, unaligned planes (105 elements)
                                   (2, 3, 37, 41, 33),        # many chunks, odd plane size: scalar path, ragged last chunk
                                   (1, 12, 32, 32, 32),       # aligned, 8 chunks per plane
                                   (1, 1, 1, 1, 1)])          # a single element: variance 0, output beta
@pytest.mark.parametrize("affine", [False, True])
def test_instance_norm_shapes_vs_oracle(shape, affine):
    from xlstm_hved_b200 import modules
    g = torch.Generator().manual_seed(sum(shape))
    x = (torch.randn(shape, generator=g) * 3 - 1).double()
    C = shape[1]
    w = (1 + 0.3 * torch.randn(C, generator=g)).double() if affine else None
    b = (0.2 * torch.randn(C, generator=g)).double() if affine else None
    gy = torch.randn(shape, generator=g).double()
    leaves = [x.clone().requires_grad_()] + ([w.clone().requires_grad_(), b.clone().requires_grad_()] if affine else [])
    y_ref = restate.instance_norm_lrelu(leaves[0], *(leaves[1:] if affine else (None, None)), slope=0.01)
    ref = torch.autograd.grad(y_ref, leaves, gy)
    cu = [t.float().cuda().requires_grad_() for t in leaves]
    y = modules.instance_norm_act(cu[0], *(cu[1:] if affine else (None, None)), slope=0.01)
    scale = max(y_ref.abs().max().item(), 1e-3)
    assert (y.double().cpu() - y_ref).abs().max().item() < TOL_F32 * scale + 1e-6
    got = torch.autograd.grad(y, cu, gy.float().cuda())
    for a, r in zip(got, ref):
        assert (a.double().cpu() - r).abs().max().item() < TOL_G32 * max(r.abs().max().item(), 1.0)


@pytest.mark.parametrize("shape", [(2, 4, 16, 16, 16), (1, 4, 64, 64, 48), (2, 3, 37, 41, 33)])
def test_constant_planes_normalise_to_exactly_zero(shape):
    """The encoder of a missing modality (input zeroed, evaluation.py:306-307) feeds constant planes (the conv bias) into the
    norm.  The exact answer is 0 (x - mean == 0); an ulp of error in the mean is multiplied by 1 / sqrt(eps) = 316 and blown up to
    unit variance by the next layer's norm (PyTorch's fp32 kernel returns 0 on power-of-two planes and ~1e-5 on others:
    tools/diag_conv_norm.py), so the kernel is exact by construction (sums around a pivot element) -- per plane (instance), over
    the batch (batch statistics), with and without the activation."""
    from xlstm_hved_b200 import modules
    N, C = shape[:2]
    vals = torch.tensor([0.7310586, -3.3000002, 1.2345e-3, 0.0, 171.25, -1e-5])
    x = vals[torch.arange(N * C) % 6].reshape(N, C, 1, 1, 1).expand(shape).contiguous().cuda()
    assert modules.instance_norm_act(x, slope=0.01).abs().max().item() == 0.0
    xb = vals[torch.arange(C) % 6].reshape(1, C, 1, 1, 1).expand(shape).contiguous().cuda()
    w, b = torch.full((C,), 1.5, device="cuda"), torch.full((C,), 0.25, device="cuda")
    y = modules.batch_norm_act(xb, w, b, None, None, True, 0.1)
    assert (y - 0.25).abs().max().item() == 0.0
    if shape[2:] == (16, 16, 16):                                   # PyTorch is exact too where its Welford merge tree is even
        assert torch.nn.functional.instance_norm(x).abs().max().item() == 0.0


@pytest.mark.parametrize("dtype,tol", [(torch.float16, 2e-3), (torch.bfloat16, 1.6e-2)])
def test_instance_norm_half_precision_io(dtype, tol):
    """Under autocast (train.py:207) the convolutions hand over fp16: the kernel reads / writes the 16-bit type, computes in fp32."""
    from xlstm_hved_b200 import modules
    g = torch.Generator().manual_seed(8)
    x = (torch.randn(2, 4, 24, 20, 26, generator=g) * 2 + 0.5).to(dtype)          # 12,480 elements per plane: two chunks
    gy = torch.randn(x.shape, generator=g).to(dtype)
    xr = x.double().requires_grad_()
    y_ref = restate.instance_norm_lrelu(xr, slope=0.01)
    (dx_ref,) = torch.autograd.grad(y_ref, xr, gy.double())
    xc = x.cuda().requires_grad_()
    y = modules.instance_norm_act(xc, slope=0.01)
    assert y.dtype == dtype
    (dx,) = torch.autograd.grad(y, xc, gy.cuda())
    assert dx.dtype == dtype
    assert rel_linf(y, y_ref) < tol and rel_linf(dx, dx_ref) < tol
    assert rel_l2(y, y_ref) < tol / 2 and rel_l2(dx, dx_ref) < tol / 2


def test_instance_norm_full_size_properties_and_torch_agreement():
    """BASELINE size: the (1, 4, 128^3) tensor the first encoder level normalises.  Size-independent properties (zero mean, unit
    variance per plane before the activation; invariance to a per-plane shift and scale of the input) and agreement with PyTorch's
    own kernels on the same device (the path the patch replaces), forward and backward."""
    from xlstm_hved_b200 import modules
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.rand(1, 4, 128, 128, 128, device="cuda", generator=g) * 3 - 0.2
    y = modules.instance_norm_act(x)                                      # slope 1: plain normalisation
    flat = y.double().reshape(4, -1)
    assert flat.mean(-1).abs().max().item() < 1e-5
    assert (flat.var(-1, unbiased=False) - 1).abs().max().item() < 1e-4
    scale = torch.tensor([0.5, 2.0, 7.0, 1.0], device="cuda").reshape(1, 4, 1, 1, 1)
    y2 = modules.instance_norm_act(x * scale + 4.0)
    assert (y2 - y).abs().max().item() < 2e-4
    xr = x.clone().requires_grad_()
    ref = torch.nn.functional.leaky_relu(torch.nn.functional.instance_norm(xr), 0.01)
    xc = x.clone().requires_grad_()
    got = modules.instance_norm_act(xc, slope=0.01)
    assert (got - ref).abs().max().item() < 1e-5
    gy = torch.randn(x.shape, device="cuda", generator=g)
    (dr,) = torch.autograd.grad(ref, xr, gy)
    (dg,) = torch.autograd.grad(got, xc, gy)
    assert (dg - dr).abs().max().item() < 1e-4 * dr.abs().max().item()


def test_patched_norms_pickle_as_reference_classes_and_track_running_stats_is_left_alone():
    import io
    import xlstm_hved_b200 as xh
    m = nn.Sequential(nn.InstanceNorm3d(3), nn.LeakyReLU(0.2, inplace=True), nn.InstanceNorm3d(3, track_running_stats=True),
                      nn.BatchNorm3d(3)).cuda()
    counts = xh.patch_model(m)
    assert counts["InstanceNorm3d"] == 1 and counts["BatchNorm3d"] == 1 and counts["fused_LeakyReLU"] == 1
    assert type(m[2]) is nn.InstanceNorm3d                         # not taken over
    buf = io.BytesIO()
    torch.save(m, buf)
    buf.seek(0)
    back = torch.load(buf, weights_only=False)
    assert type(back[0]) is nn.InstanceNorm3d and type(back[1]) is nn.LeakyReLU and "fused_slope" not in back[0].__dict__
    x = torch.randn(2, 3, 6, 6, 6, device="cuda")
    ref = back.train()(x.clone())
    got = m.train()(x.clone())
    assert (ref - got).abs().max().item() < 1e-5
    xh.unpatch_model(m)
