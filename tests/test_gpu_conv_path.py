"""GPU parity of the conv-path kernels K7 - K10.  K7 (the spatial gate of AttenModule2: depthwise 7^3 conv + 1x1x1 conv + sigmoid as one dense G -> 1 convolution
kernel, csrc/gate7.cu) through the C ABI: against the fixture the REAL reference module produced (tests/golden/atten_module2.pt),
the fp64 oracle on ragged shapes, and PyTorch's own convolution kernels at the model's full size."""
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from conftest import load_golden, rel_linf
from oracle import restate

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _exact_convs():
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32 = old


class AttenModule2(nn.Module):
    """Parameters and structure of buildingblocks.AttenModule2 (buildingblocks.py:263-274); the patch recognises it by class name and
    replaces the forward, so this stand-in has none of its own."""

    def __init__(self):
        super().__init__()
        self.compress = _ChannelPool()
        self.enc_spatial = nn.Conv3d(4, 16, 7, stride=1, padding=3, groups=4)
        self.enc_spatial2 = nn.Conv3d(16, 1, 1, stride=1)
        self.seg_spatial = nn.Conv3d(2, 8, 7, stride=1, padding=3, groups=2)
        self.seg_spatial2 = nn.Conv3d(8, 1, 1, stride=1)

    def forward(self, seg_x, enc_x, recon_x=None):
        raise AssertionError("the patched forward must run")


class _ChannelPool(nn.Module):
    def forward(self, x):
        return restate.channel_pool(x)


def test_patched_atten_module2_golden():
    import xlstm_hved_b200 as xh
    c = load_golden("atten_module2.pt")
    att = AttenModule2()
    att.load_state_dict({k: v.float() for k, v in c["state_dict"].items()}, strict=True)
    att.cuda()
    counts = xh.patch_model(att)
    assert counts["AttenModule2"] == 1
    seg_x, enc_x = c["seg_x"].float().cuda().requires_grad_(), c["enc_x"].float().cuda().requires_grad_()
    y = att(seg_x, enc_x)
    assert rel_linf(y, c["y"]) < 1e-5
    names = list(c["param_grads"])
    params = dict(att.named_parameters())
    grads = torch.autograd.grad(y, [seg_x, enc_x] + [params[n] for n in names], c["gy"].float().cuda())
    assert rel_linf(grads[0], c["d_seg_x"]) < 1e-4 and rel_linf(grads[1], c["d_enc_x"]) < 1e-4
    for g, n in zip(grads[2:], names):
        assert rel_linf(g, c["param_grads"][n]) < 1e-4, n
    xh.unpatch_model(att)
    assert type(att) is AttenModule2


@pytest.mark.parametrize("shape", [(1, 4, 16, 16, 32), (2, 2, 5, 7, 3), (1, 4, 24, 17, 70), (3, 1, 8, 8, 33)])
def test_gate7_shapes_vs_oracle(shape):
    """Volumes smaller than one tile, ragged in every dimension, several samples, 1 / 2 / 4 channels; forward and all gradients
    (x, the depthwise weights and bias, the pointwise weights and bias) against fp64 autograd through the two convolutions."""
    from xlstm_hved_b200 import modules
    N, G = shape[:2]
    torch.manual_seed(sum(shape))
    dw, pw = nn.Conv3d(G, 4 * G, 7, padding=3, groups=G), nn.Conv3d(4 * G, 1, 1)
    with torch.no_grad():
        for q in list(dw.parameters()) + list(pw.parameters()):
            q.mul_(3.0)
    x = torch.randn(shape)
    gy = torch.randn(N, 1, *shape[2:])
    dw64, pw64 = nn.Conv3d(G, 4 * G, 7, padding=3, groups=G).double(), nn.Conv3d(4 * G, 1, 1).double()
    dw64.load_state_dict(dw.state_dict()), pw64.load_state_dict(pw.state_dict())
    x64 = x.double().requires_grad_()
    ref = torch.sigmoid(pw64(dw64(x64)))
    ref_grads = torch.autograd.grad(ref, [x64] + list(dw64.parameters()) + list(pw64.parameters()), gy.double())
    dw.cuda(), pw.cuda()
    xc = x.cuda().requires_grad_()
    got = modules.spatial_gate(xc, dw, pw)
    assert (got.double().cpu() - ref).abs().max().item() < 2e-6                    # a sigmoid: absolute
    grads = torch.autograd.grad(got, [xc] + list(dw.parameters()) + list(pw.parameters()), gy.cuda())
    for a, r in zip(grads, ref_grads):
        assert rel_linf(a, r) < 1e-4


def test_gate7_full_size_vs_pytorch_kernels():
    """The (1, 4, 128^3) gate of the model's last attention stage against PyTorch's conv_depthwise3d + cuDNN path on the same
    device (the path the patch replaces), forward and backward; linearity of the pre-activation in x as a size-independent check."""
    from xlstm_hved_b200 import modules, ops
    torch.manual_seed(0)
    dw, pw = nn.Conv3d(4, 16, 7, padding=3, groups=4).cuda(), nn.Conv3d(16, 1, 1).cuda()
    x = torch.randn(1, 4, 128, 128, 128, device="cuda")
    gy = torch.randn(1, 1, 128, 128, 128, device="cuda")
    xr = x.clone().requires_grad_()
    ref = torch.sigmoid(pw(dw(xr)))
    ref_grads = torch.autograd.grad(ref, [xr] + list(dw.parameters()) + list(pw.parameters()), gy)
    xc = x.clone().requires_grad_()
    got = modules.spatial_gate(xc, dw, pw)
    assert (got - ref).abs().max().item() < 5e-6
    grads = torch.autograd.grad(got, [xc] + list(dw.parameters()) + list(pw.parameters()), gy)
    for a, r in zip(grads, ref_grads):
        assert rel_linf(a, r) < 2e-4
    # logit(gate(a x1 + b x2)) = a logit(gate(x1)) + b logit(gate(x2)) - (a + b - 1) bias
    w = torch.randn(4, 343, device="cuda") * 0.02
    x2 = torch.randn_like(x)
    logit = lambda t: torch.log(t) - torch.log1p(-t)
    l1, l2, l12 = (logit(ops.gate7_fwd(t, w).double()) for t in (x, x2, 0.5 * x - 1.5 * x2))
    assert (l12 - (0.5 * l1 - 1.5 * l2)).abs().max().item() < 1e-4


# ------------------------------------------------------------------ K8: depthwise 3x3x3 convolution (csrc/dwconv3.cu)
class BasicConv(nn.Module):
    """Structure of buildingblocks.BasicConv (buildingblocks.py:11-31) as RA_HVED.py:406 instantiates it."""

    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv3d(c, c, 3, padding=1, groups=c, bias=False)
        self.norm = nn.InstanceNorm3d(c)
        self.relu = nn.LeakyReLU(negative_slope=1e-2, inplace=True)

    def forward(self, x):
        return self.relu(self.norm(self.conv(x)))


def test_patched_depthwise_basic_conv_golden():
    """The reference's BasicConv(4, 4, 3, padding=1, groups=4): conv on K8, norm + LeakyReLU on K6, against the reference's values."""
    import xlstm_hved_b200 as xh
    from xlstm_hved_b200 import ops
    c = load_golden("dwconv3.pt")
    bc = BasicConv(4)
    bc.load_state_dict({k: v.float() for k, v in c["state_dict"].items()}, strict=True)
    bc.cuda()
    counts = xh.patch_model(bc)
    assert counts["DepthwiseConv3d"] == 1 and counts["InstanceNorm3d"] == 1 and counts["fused_LeakyReLU"] == 1
    x = c["x"].float().cuda().requires_grad_()
    conv_out = bc.conv(x)
    assert rel_linf(conv_out, c["conv_out"]) < 1e-5
    dx, dw = torch.autograd.grad(conv_out, [x, bc.conv.weight], c["gc"].float().cuda())
    assert rel_linf(dx, c["conv_dx"]) < 1e-5 and rel_linf(dw, c["conv_dweight"]) < 1e-4
    y = bc(x)
    assert rel_linf(y, c["y"]) < 5e-5
    dx, dw = torch.autograd.grad(y, [x, bc.conv.weight], c["gy"].float().cuda())
    assert rel_linf(dx, c["dx"]) < 5e-4 and rel_linf(dw, c["dconv_weight"]) < 5e-4
    xh.unpatch_model(bc)
    assert type(bc.conv) is nn.Conv3d


@pytest.mark.parametrize("shape,bias", [((1, 4, 16, 16, 32), False), ((2, 3, 5, 7, 3), True), ((1, 8, 24, 17, 70), True),
                                        ((1, 1, 9, 8, 33), True), ((1, 4, 128, 128, 128), False)])
def test_dwconv3_vs_pytorch_fp64(shape, bias):
    """Forward, input gradient, weight and bias gradients against F.conv3d in fp64 (on the GPU for the full-size case)."""
    from xlstm_hved_b200 import modules
    C = shape[1]
    torch.manual_seed(sum(shape))
    conv = nn.Conv3d(C, C, 3, padding=1, groups=C, bias=bias).cuda()
    x = torch.randn(shape, device="cuda")
    gy = torch.randn(shape, device="cuda")
    params = list(conv.parameters())
    x64 = x.double().requires_grad_()
    p64 = [p.detach().double().requires_grad_() for p in params]
    ref = F.conv3d(x64, p64[0], p64[1] if bias else None, padding=1, groups=C)
    ref_grads = torch.autograd.grad(ref, [x64] + p64, gy.double())
    xc = x.clone().requires_grad_()
    assert modules.dwconv3_supported(conv)
    got = modules.depthwise_conv3_forward(conv, xc)
    assert rel_linf(got, ref) < 1e-5
    grads = torch.autograd.grad(got, [xc] + params, gy)
    for a, r in zip(grads, ref_grads):
        assert rel_linf(a, r) < 1e-4


# ------------------------------------------------------------------ K9: 1x1x1 convolution (csrc/pwconv.cu)
@pytest.mark.parametrize("cin,cout,shape,bias", [(4, 4, (16, 16, 32), True), (4, 1, (5, 7, 3), True), (4, 3, (9, 11, 37), True),
                                                 (8, 32, (8, 8, 8), False), (16, 16, (6, 5, 7), True), (32, 32, (4, 4, 4), False),
                                                 (8, 8, (12, 10, 18), True), (16, 1, (8, 8, 8), True), (4, 4, (128, 128, 128), True)])
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-5), (torch.float16, 3e-3)])
def test_pwconv_vs_fp64(cin, cout, shape, bias, dtype, tol):
    """Forward, input gradient, weight and bias gradients of the 1x1x1 convolution against an fp64 einsum (on the GPU): register
    tiles of 1 / 4 / 8 / 16 / 32 output channels, both weight-gradient strategies, aligned and ragged volumes, fp32 and fp16 I/O."""
    from xlstm_hved_b200 import modules
    N = 1 if shape[0] == 128 else 2
    torch.manual_seed(cin * 100 + cout + sum(shape))
    conv = nn.Conv3d(cin, cout, 1, bias=bias).cuda()
    x = torch.randn(N, cin, *shape, device="cuda").to(dtype)
    gy = torch.randn(N, cout, *shape, device="cuda").to(dtype)
    params = list(conv.parameters())
    x64 = x.double().requires_grad_()
    p64 = [p.detach().double().requires_grad_() for p in params]
    ref = torch.einsum("oi,nidhw->nodhw", p64[0].reshape(cout, cin), x64)
    if bias:
        ref = ref + p64[1].reshape(1, cout, 1, 1, 1)
    ref_grads = torch.autograd.grad(ref, [x64] + p64, gy.double())
    xc = x.clone().requires_grad_()
    assert modules.pwconv_supported(conv)
    got = modules.pointwise_conv_forward(conv, xc)
    assert got.dtype == dtype and rel_linf(got, ref) < tol
    grads = torch.autograd.grad(got, [xc] + params, gy)
    assert grads[0].dtype == dtype and grads[1].dtype == torch.float32
    for a, r in zip(grads, ref_grads):
        assert rel_linf(a, r) < (tol if dtype == torch.float16 else 1e-4)


def test_patched_pointwise_conv_under_autocast_matches_pytorch():
    """Inside an fp16 autocast region (train.py:207) the patched 1x1x1 layer takes an fp32 input as fp16 and returns fp16, like
    PyTorch's convolution; values and gradients agree with the stock layer to fp16 accuracy."""
    import copy
    import xlstm_hved_b200 as xh
    torch.manual_seed(1)
    stock = nn.Sequential(nn.Conv3d(4, 4, 1), nn.Conv3d(4, 1, 1)).cuda()
    mine = copy.deepcopy(stock)
    counts = xh.patch_model(mine)
    assert counts["PointwiseConv3d"] == 2
    x = torch.randn(1, 4, 32, 32, 32, device="cuda")
    outs = []
    for m in (stock, mine):
        xin = x.clone().requires_grad_()
        with torch.autocast("cuda", dtype=torch.float16):
            y = m(xin)
        assert y.dtype == torch.float16
        g = torch.autograd.grad(y.float().square().sum(), [xin] + list(m.parameters()))
        outs.append((y, g))
    assert rel_linf(outs[1][0], outs[0][0]) < 5e-3
    for a, r in zip(outs[1][1], outs[0][1]):
        assert a.dtype == r.dtype and rel_linf(a, r) < 1e-2
    xh.unpatch_model(mine)


# ------------------------------------------------------------------ K10: dense 3x3x3 convolution with few channels (csrc/conv3.cu)
@pytest.mark.parametrize("cin,cout,shape,bias", [(4, 4, (16, 16, 32), True), (12, 4, (9, 11, 37), True), (4, 8, (8, 8, 33), True),
                                                 (24, 8, (12, 10, 18), False), (48, 16, (8, 9, 10), True), (16, 32, (5, 7, 3), True),
                                                 (3, 5, (40, 8, 32), True), (4, 4, (128, 128, 128), True)])
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-5), (torch.float16, 4e-3)])
def test_conv3_vs_pytorch_fp64(cin, cout, shape, bias, dtype, tol):
    """Forward, input gradient, weight and bias gradients against F.conv3d in fp64 (on the GPU): both register tiles, several
    output-channel groups, ragged volumes, more than four tiles along d (two weight-gradient super-tiles), fp32 and fp16 I/O."""
    from xlstm_hved_b200 import modules
    N = 1 if shape[0] == 128 else 2
    torch.manual_seed(cin * 100 + cout + sum(shape))
    conv = nn.Conv3d(cin, cout, 3, padding=1, bias=bias).cuda()
    x = torch.randn(N, cin, *shape, device="cuda").to(dtype)
    gy = torch.randn(N, cout, *shape, device="cuda").to(dtype)
    params = list(conv.parameters())
    x64 = x.double().requires_grad_()
    p64 = [p.detach().double().requires_grad_() for p in params]
    ref = F.conv3d(x64, p64[0], p64[1] if bias else None, padding=1)
    ref_grads = torch.autograd.grad(ref, [x64] + p64, gy.double())
    xc = x.clone().requires_grad_()
    assert modules.conv3_supported(conv)
    got = modules.dense_conv3_forward(conv, xc)
    assert got.dtype == dtype and rel_linf(got, ref) < tol
    grads = torch.autograd.grad(got, [xc] + params, gy)
    assert grads[0].dtype == dtype and grads[1].dtype == torch.float32
    for a, r in zip(grads, ref_grads):
        assert rel_linf(a, r) < (tol if dtype == torch.float16 else 1e-4)


# ------------------------------------------------------------------ K9 + K6 inside the reference's DuSEAttention (fixture from the real module)
class DuSEAttention(nn.Module):
    """Parameters and forward of modules/DuSFE.py:87-154 (dual squeeze-and-excitation with two BatchNorm3d layers), restated so that the
    fixture recorded from the REAL module can be replayed on the GPU box, where the reference tree does not exist."""

    def __init__(self, c):
        super().__init__()
        self.avg_pool_ch1, self.avg_pool_ch2 = nn.AdaptiveAvgPool3d((1, 1, 1)), nn.AdaptiveAvgPool3d((1, 1, 1))
        self.fc_comb, self.fc_ch1, self.fc_ch2 = nn.Linear(2 * c, c), nn.Linear(c, c), nn.Linear(c, c)
        self.conv_squeeze_ch1, self.conv_squeeze_ch2 = nn.Conv3d(c, 1, 1), nn.Conv3d(c, 1, 1)
        self.conv_comb = nn.Conv3d(2, 1, 1)
        self.conv_adjust_ch1, self.conv_adjust_ch2 = nn.Conv3d(1, 1, 3, padding=1), nn.Conv3d(1, 1, 3, padding=1)
        self.conv_fuse_ch1, self.bn_fuse_ch1 = nn.Conv3d(3 * c, c, 3, padding=1), nn.BatchNorm3d(c)
        self.conv_fuse_ch2, self.bn_fuse_ch2 = nn.Conv3d(3 * c, c, 3, padding=1), nn.BatchNorm3d(c)

    def forward(self, a, b):
        B, C, D, H, W = a.size()
        s = torch.cat((self.avg_pool_ch1(a).view(B, C), self.avg_pool_ch2(b).view(B, C)), 1)
        f = self.fc_comb(s)
        f1, f2 = torch.sigmoid(self.fc_ch1(f)), torch.sigmoid(self.fc_ch2(f))
        a_sc, b_sc = a * f1.view(B, C, 1, 1, 1), b * f2.view(B, C, 1, 1, 1)
        comb = self.conv_comb(torch.cat((self.conv_squeeze_ch1(a), self.conv_squeeze_ch2(b)), 1))
        g1, g2 = torch.sigmoid(self.conv_adjust_ch1(comb)), torch.sigmoid(self.conv_adjust_ch2(comb))
        a_cs, b_cs = a * g1.view(B, 1, D, H, W), b * g2.view(B, 1, D, H, W)
        return self.bn_fuse_ch1(a + a_sc + a_cs), self.bn_fuse_ch2(b + b_sc + b_cs)


def test_patched_duse_attention_golden():
    """The whole DuSEAttention of the reference with its 1x1x1 convolutions on K9, the 1 -> 1 3x3x3 ones on K8 and both BatchNorm3d
    layers on K6: train mode (values, gradients w.r.t. both inputs and the BatchNorm parameters, running statistics) and eval mode."""
    import xlstm_hved_b200 as xh
    c = load_golden("conv_norm.pt")["duse_attention"]
    att = DuSEAttention(4)
    att.load_state_dict({k: v.float() for k, v in c["state_dict_before"].items()}, strict=True)
    att.cuda()
    counts = xh.patch_model(att)
    assert counts["BatchNorm3d"] == 2 and counts["PointwiseConv3d"] == 3 and counts["DepthwiseConv3d"] == 2
    a, b = c["a"].float().cuda().requires_grad_(), c["b"].float().cuda().requires_grad_()
    g1, g2 = c["g1"].float().cuda(), c["g2"].float().cuda()
    params = [att.bn_fuse_ch1.weight, att.bn_fuse_ch1.bias, att.bn_fuse_ch2.weight, att.bn_fuse_ch2.bias]
    att.train()
    y1, y2 = att(a, b)
    assert rel_linf(y1, c["train_y1"]) < 5e-5 and rel_linf(y2, c["train_y2"]) < 5e-5
    grads = torch.autograd.grad([y1, y2], [a, b] + params, [g1, g2])
    for got, ref in zip(grads, c["train_grads"]):
        assert rel_linf(got, ref) < 5e-4
    after = c["state_dict_after_train"]
    for name in ("bn_fuse_ch1", "bn_fuse_ch2"):
        bn = getattr(att, name)
        assert rel_linf(bn.running_mean, after[name + ".running_mean"]) < 1e-5 and rel_linf(bn.running_var, after[name + ".running_var"]) < 1e-5
    att.eval()
    e1, e2 = att(a, b)
    assert rel_linf(e1, c["eval_y1"]) < 5e-5 and rel_linf(e2, c["eval_y2"]) < 5e-5
    egrads = torch.autograd.grad([e1, e2], [a, b] + params, [g1, g2])
    for got, ref in zip(egrads, c["eval_grads"]):
        assert rel_linf(got, ref) < 5e-4
    xh.unpatch_model(att)
