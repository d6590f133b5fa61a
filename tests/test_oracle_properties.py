"""Size-independent properties of the oracle itself (CPU, fp64): the facts of SURVEY 8a-note the kernels are built on.
They pin the restatement beyond the fixed fixtures: any chunk length gives the parallel form, the backward restatement is
autograd of the forward, the flipped direction is the anti-causal scan, PoE gradients are autograd of PoE."""
import pytest
import torch

from conftest import rel_linf
from oracle import restate


def _cell_inputs(B, NH, S, DH, fmean, seed):
    g = torch.Generator().manual_seed(seed)
    q, k, v = (torch.randn(B, NH, S, DH, generator=g, dtype=torch.float64) for _ in range(3))
    ig = torch.randn(B, NH, S, 1, generator=g, dtype=torch.float64)
    fg = fmean + torch.randn(B, NH, S, 1, generator=g, dtype=torch.float64)
    return q, k, v, ig, fg


@pytest.mark.parametrize("S,chunk", [(1, 128), (5, 2), (97, 16), (128, 128), (129, 128), (300, 64), (300, 300), (513, 128)])
@pytest.mark.parametrize("fmean", [-2.0, 0.4, 4.0])
def test_chunkwise_equals_parallel_for_any_chunk_length(S, chunk, fmean):
    q, k, v, ig, fg = _cell_inputs(2, 3, S, 8, fmean, S * 7 + chunk)
    ref = restate.mlstm_parallel(q, k, v, ig, fg)
    got = restate.mlstm_chunkwise(q, k, v, ig, fg, chunk=chunk)
    assert (got - ref).abs().max() <= 1e-10 * (1 + ref.abs().max())


@pytest.mark.parametrize("S", [3, 64, 200])
def test_backward_restatement_is_autograd_of_the_forward(S):
    q, k, v, ig, fg = _cell_inputs(1, 2, S, 8, 1.0, S)
    leaves = [t.clone().requires_grad_() for t in (q, k, v, ig, fg)]
    dh = torch.randn(1, 2, S, 8, generator=torch.Generator().manual_seed(1), dtype=torch.float64)
    ref = torch.autograd.grad(restate.mlstm_parallel(*leaves), leaves, dh)
    got = restate.mlstm_backward(q, k, v, ig, fg, dh, through_max=True)
    for a, b in zip(got, ref):
        assert (a - b).abs().max() <= 1e-9 * (1 + b.abs().max())
    # dropping the gradient through the row maximum (what the kernels do) moves the gate gradients by ~eps only
    approx = restate.mlstm_backward(q, k, v, ig, fg, dh, through_max=False)
    for a, b in zip(approx[3:], ref[3:]):
        assert (a - b).norm() <= 1e-4 * (1e-12 + b.norm())


def test_stabiliser_is_a_one_dimensional_scan():
    q, k, v, ig, fg = _cell_inputs(2, 2, 150, 4, 0.0, 3)
    _, m_ref, _ = restate.mlstm_parallel(q, k, v, ig, fg, return_aux=True)
    m_scan = restate.mlstm_stabiliser_scan(ig, fg)
    assert (m_scan.squeeze(-1) - m_ref.squeeze(-1)).abs().max() < 1e-10


def test_reverse_block_is_the_forward_block_on_the_flipped_sequence():
    from conftest import load_golden
    c = load_golden("vil_block.pt")["dim32_s200_fwd"]
    p = {k: v.double() for k, v in c["state_dict"].items()}
    x = torch.randn(2, 77, 32, generator=torch.Generator().manual_seed(5), dtype=torch.float64)
    fwd_on_flipped = restate.vil_block(x.flip(1), p, reverse=False).flip(1)
    rev = restate.vil_block(x, p, reverse=True)
    assert (fwd_on_flipped - rev).abs().max() < 1e-12


@pytest.mark.parametrize("subset", [(0,), (1, 3), (0, 1, 2, 3)])
def test_poe_backward_restatement_is_autograd(subset):
    g = torch.Generator().manual_seed(len(subset))
    sh = (2, 3, 4, 4, 4)
    mu = torch.cat([torch.zeros(1, *sh), torch.randn(4, *sh, generator=g)]).double().requires_grad_()
    lv = torch.cat([torch.zeros(1, *sh), torch.randn(4, *sh, generator=g)]).double().requires_grad_()
    g_mu, g_lv = torch.randn(sh, generator=g).double(), torch.randn(sh, generator=g).double()
    a, b = restate.poe(mu, lv, subset)
    ref = torch.autograd.grad((a * g_mu).sum() + (b * g_lv).sum(), [mu, lv])
    got = restate.poe_backward(mu.detach(), lv.detach(), subset, g_mu, g_lv)
    for x, y in zip(got, ref):           # the restatement returns the 4 modality slabs (the prior is a constant)
        assert (x - y[1:]).abs().max() <= 1e-12 * (1 + y.abs().max())


def test_dropped_modalities_equal_the_complementary_subset():
    g = torch.Generator().manual_seed(2)
    sh = (3, 2, 4, 4, 4)
    mu = torch.cat([torch.zeros(1, *sh), torch.randn(4, *sh, generator=g)]).double()
    lv = torch.cat([torch.zeros(1, *sh), torch.randn(4, *sh, generator=g)]).double()
    drop = torch.zeros(3, 4, dtype=torch.bool)
    drop[:, 1] = True
    a, b, mu_after = restate.poe_drop(mu.clone(), lv, drop)
    c, d = restate.poe(mu, lv, (0, 2, 3))
    assert (a - c).abs().max() < 1e-12 and (b - d).abs().max() < 1e-12
    assert (mu_after[2] == 0).all() and torch.equal(mu_after[1], mu[1])      # the reference zeroes mu[m+1] of dropped modalities


# ------------------------------------------------------------------ conv path (K6 - K10): identities the kernels rely on
def test_spatial_gate_composition_identity():
    """K7 composes AttenModule2's depthwise 7^3 convolution (expansion 4) and the 1x1x1 convolution that follows into ONE dense G -> 1
    convolution: W[g] = sum_j w2[4g + j] W1[4g + j], b = sum_o w2[o] b1[o] + b2 (csrc/gate7.cu, modules.spatial_gate).  Checked in fp64
    against the two convolutions run one after the other, as the reference does (buildingblocks.py:283-285)."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(0)
    G, e = 4, 4
    x = torch.randn(2, G, 9, 8, 11, generator=g, dtype=torch.float64)
    w1, b1 = torch.randn(G * e, 1, 7, 7, 7, generator=g, dtype=torch.float64), torch.randn(G * e, generator=g, dtype=torch.float64)
    w2, b2 = torch.randn(1, G * e, 1, 1, 1, generator=g, dtype=torch.float64), torch.randn(1, generator=g, dtype=torch.float64)
    ref = F.conv3d(F.conv3d(x, w1, b1, padding=3, groups=G), w2, b2)
    w = (w1.reshape(G, e, 343) * w2.reshape(G, e, 1)).sum(1).reshape(1, G, 7, 7, 7)
    b = (w2.reshape(-1) * b1).sum() + b2
    assert rel_linf(F.conv3d(x, w, b, padding=3), ref) < 1e-12


def test_instance_norm_restatement_properties():
    """Invariance of the normalisation to a per-plane shift and positive scale of its input; batch statistics over one sample equal
    instance statistics; the LeakyReLU commutes with nothing but is applied after the affine map (slope 1 = identity)."""
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 3, 5, 6, 7, generator=g, dtype=torch.float64) * 3 + 1
    y = restate.instance_norm_lrelu(x, eps=0.0)
    scale = torch.tensor([0.5, 2.0, 7.0], dtype=torch.float64).reshape(1, 3, 1, 1, 1)
    assert rel_linf(restate.instance_norm_lrelu(x * scale + 4.0, eps=0.0), y) < 1e-12
    assert y.reshape(6, -1).mean(-1).abs().max() < 1e-12 and (y.reshape(6, -1).var(-1, unbiased=False) - 1).abs().max() < 1e-12
    w, b = torch.ones(3, dtype=torch.float64), torch.zeros(3, dtype=torch.float64)
    one = x[:1]
    yb, _, _ = restate.batch_norm_lrelu(one, w, b, training=True, slope=0.2)
    assert rel_linf(yb, restate.instance_norm_lrelu(one, slope=0.2)) < 1e-12
    pre = restate.instance_norm_lrelu(x)
    act = restate.instance_norm_lrelu(x, slope=0.01)
    assert torch.equal(act, torch.where(pre > 0, pre, pre * 0.01))


def test_depthwise_restatement_is_a_grouped_convolution():
    """restate.basic_conv_depthwise writes the 27 taps out; it must equal F.conv3d with groups = channels (what the reference runs)."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(2)
    x = torch.randn(2, 4, 5, 6, 7, generator=g, dtype=torch.float64)
    w = torch.randn(4, 1, 3, 3, 3, generator=g, dtype=torch.float64)
    _, conv = restate.basic_conv_depthwise(x, w)
    assert rel_linf(conv, F.conv3d(x, w, padding=1, groups=4)) < 1e-12
