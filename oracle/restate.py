"""CPU restatement of the XLSTM-HVED hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this module.  Nothing under
``xlstm_hved_b200/`` imports it; the product path is CUDA-only and fails loudly
when its shared library is missing.

Every function restates one piece of the reference (paths relative to the
reference root) in plain torch, dtype-generic (run it in float64 for the oracle,
float32 for the "what the reference computes" CPU baseline).

Parity pinning: the reference holds NO tests / golden vectors for this path
(SURVEY.md section 4), so these restatements are pinned against outputs of the
reference itself, imported and executed in the build container by
``oracle/make_golden.py``; the resulting fixtures live in ``tests/golden/`` and
``tests/test_oracle_golden.py`` re-checks the restatement against them on every
run (CPU, no reference needed).
"""
from __future__ import annotations

import math
from itertools import chain, combinations

import torch
import torch.nn.functional as F

# RA_HVED.py:733-738 -- subset index -> tuple of present modalities
SUBSETS_MODALITIES = list(chain(*map(lambda r: combinations([0, 1, 2, 3], r), range(1, 5))))


# --------------------------------------------------------------------------
# K1: stabilised mLSTM cell
# --------------------------------------------------------------------------
def mlstm_parallel(q, k, v, ig, fg, eps: float = 1e-6, return_aux: bool = False):
    """Parallel (O(S^2)) stabilised mLSTM -- vision_lstm.py:48-130.

    q,k,v: (B,NH,S,DH); ig,fg: (B,NH,S,1) gate pre-activations.
    Row-wise stabiliser (stabilize_rowwise=True) and eps=1e-6 as used by the
    reference's only call site (vision_lstm.py:327-334).
    """
    B, NH, S, DH = q.shape
    lf = F.logsigmoid(fg)                                   # :82
    c = torch.cumsum(lf, dim=-2)                            # :89-95 (the leading 0 row/col is cut at :105)
    logD = c - c.transpose(-2, -1) + ig.transpose(-2, -1)   # :99-108  c_t - c_s + i_s
    tri = torch.tril(torch.ones(S, S, dtype=torch.bool, device=q.device))
    logD = torch.where(tri, logD, torch.full_like(logD, -float("inf")))
    m = logD.max(dim=-1, keepdim=True).values               # :111
    Dm = torch.exp(logD - m)                                # :115-116
    Sm = q @ (k / math.sqrt(DH)).transpose(-2, -1)          # :118-121
    C = Sm * Dm                                             # :122
    den = C.sum(dim=-1, keepdim=True)
    n = torch.maximum(den.abs(), torch.exp(-m))             # :123
    h = (C / (n + eps)) @ v                                 # :125-128
    if return_aux:
        return h, m.squeeze(-1), den.squeeze(-1)
    return h


def mlstm_parallel_reference_cost(q, k, v, ig, fg, eps: float = 1e-6):
    """Same function, keeping the reference's op sequence and temporaries
    (repeat of the (S+1)x(S+1) cumsum matrix, transpose-subtract, where, ...)
    so that timing it on CPU is a fair stand-in for the reference's cost
    (vision_lstm.py:82-128).  Used only by bench.py's CPU legs."""
    B, NH, S, DH = q.shape
    lf = F.logsigmoid(fg)
    tri = torch.tril(torch.ones((S, S), dtype=torch.bool, device=q.device))
    c0 = torch.cat([torch.zeros((B, NH, 1, 1), dtype=q.dtype, device=q.device), torch.cumsum(lf, dim=-2)], dim=-2)
    rep = c0.repeat(1, 1, 1, S + 1)
    full = rep - rep.transpose(-2, -1)
    logfg = torch.where(tri, full[:, :, 1:, 1:], -float("inf"))
    logD = logfg + ig.transpose(-2, -1)
    m, _ = torch.max(logD, dim=-1, keepdim=True)
    Dm = torch.exp(logD - m)
    ks = k / math.sqrt(DH)
    C = (q @ ks.transpose(-2, -1)) * Dm
    n = torch.maximum(C.sum(dim=-1, keepdim=True).abs(), torch.exp(-m))
    return (C / (n + eps)) @ v


def mlstm_chunkwise(q, k, v, ig, fg, chunk: int = 128, eps: float = 1e-6, return_aux: bool = False):
    """Chunkwise form, exactly equal to :func:`mlstm_parallel` (SURVEY.md 8a-note,
    derived from vision_lstm.py:82-128).  Linear in S; carried state per (b,head)
    is (Cst[DH,DH], nst[DH], mst)."""
    B, NH, S, DH = q.shape
    dt, dev = q.dtype, q.device
    scale = 1.0 / math.sqrt(DH)
    lf = F.logsigmoid(fg).squeeze(-1)            # (B,NH,S)
    ii = ig.squeeze(-1)
    Cst = torch.zeros(B, NH, DH, DH, dtype=dt, device=dev)
    nst = torch.zeros(B, NH, DH, dtype=dt, device=dev)
    mst = torch.full((B, NH), -float("inf"), dtype=dt, device=dev)
    hs, ms, dens = [], [], []
    for s0 in range(0, S, chunk):
        s1 = min(S, s0 + chunk)
        L = s1 - s0
        qc, kc, vc = q[:, :, s0:s1], k[:, :, s0:s1] * scale, v[:, :, s0:s1]
        b = torch.cumsum(lf[:, :, s0:s1], dim=-1)                    # chunk-local inclusive cumsum
        ic = ii[:, :, s0:s1]
        logD = b[..., :, None] - b[..., None, :] + ic[..., None, :]
        tri = torch.tril(torch.ones(L, L, dtype=torch.bool, device=dev))
        logD = torch.where(tri, logD, torch.full_like(logD, -float("inf")))
        m_intra = logD.max(dim=-1).values
        m_inter = b + mst[..., None]
        m = torch.maximum(m_intra, m_inter)
        Sm = (qc @ kc.transpose(-2, -1)) * torch.exp(logD - m[..., None])
        w = torch.exp(m_inter - m)                                    # 0 for the first chunk (mst=-inf)
        num = Sm @ vc + w[..., None] * (qc @ Cst)
        den = Sm.sum(-1) + w * (qc * nst[..., None, :]).sum(-1)
        n = torch.maximum(den.abs(), torch.exp(-m))
        hs.append(num / (n[..., None] + eps))
        ms.append(m)
        dens.append(den)
        # state update
        g = b[..., -1]
        a = g[..., None] - b + ic
        m_new = torch.maximum(g + mst, a.max(dim=-1).values)
        decay = torch.exp(g + mst - m_new)
        wk = torch.exp(a - m_new[..., None])[..., None] * kc          # (B,NH,L,DH)
        Cst = decay[..., None, None] * Cst + wk.transpose(-2, -1) @ vc
        nst = decay[..., None] * nst + wk.sum(dim=-2)
        mst = m_new
    h = torch.cat(hs, dim=2)
    if return_aux:
        return h, torch.cat(ms, dim=2), torch.cat(dens, dim=2)
    return h


def mlstm_stabiliser_scan(ig, fg):
    """m_t = c_t + cummax_{s<=t}(i_s - c_s): the row max of vision_lstm.py:111
    written as a 1-D scan."""
    lf = F.logsigmoid(fg).squeeze(-1)
    c = torch.cumsum(lf, dim=-1)
    return c + torch.cummax(ig.squeeze(-1) - c, dim=-1).values


def mlstm_backward(q, k, v, ig, fg, dh, eps: float = 1e-6, through_max: bool = True):
    """Manual backward of :func:`mlstm_parallel` (checked
    against autograd of vision_lstm.py:48-130 in tests).  ``through_max=False``
    drops the (tiny) gradient routed through the row-max stabiliser, which is
    what the CUDA kernels do."""
    B, NH, S, DH = q.shape
    scale = 1.0 / math.sqrt(DH)
    lf = F.logsigmoid(fg)
    c = torch.cumsum(lf, dim=-2)
    logD = c - c.transpose(-2, -1) + ig.transpose(-2, -1)
    tri = torch.tril(torch.ones(S, S, dtype=torch.bool, device=q.device))
    logD = torch.where(tri, logD, torch.full_like(logD, -float("inf")))
    m, arg = logD.max(dim=-1, keepdim=True)
    Dm = torch.exp(logD - m)
    Sm = (q @ k.transpose(-2, -1)) * scale
    C = Sm * Dm
    den = C.sum(-1, keepdim=True)
    floor = torch.exp(-m)
    N = torch.maximum(den.abs(), floor) + eps
    h = (C / N) @ v
    dhh = (dh * h).sum(-1, keepdim=True)
    dn = -dhh / N
    active = den.abs() > floor
    db = torch.where(active, dn * torch.sign(den), torch.zeros_like(dn))
    dC = (dh @ v.transpose(-2, -1)) / N + db
    dS = dC * Dm
    dv = (C / N).transpose(-2, -1) @ dh
    dq = (dS @ k) * scale
    dk = (dS.transpose(-2, -1) @ q) * scale
    G = dC * C                                   # d/dlogD (with m held fixed)
    di = G.sum(-2)                               # column sums  == k . dk
    dc = G.sum(-1) - G.sum(-2)                   # == q.dq - k.dk
    if through_max:
        # N depends on m through the floor branch; C depends on m through Dm.
        dm = torch.where(active, torch.zeros_like(dn), -dn * floor) - G.sum(-1, keepdim=True)
        onehot = torch.zeros_like(logD).scatter_(-1, arg, 1.0)
        Gm = onehot * dm
        di = di + Gm.sum(-2)
        dc = dc + Gm.sum(-1) - Gm.sum(-2)
    dlf = torch.flip(torch.cumsum(torch.flip(dc, dims=[-1]), dim=-1), dims=[-1])
    dfg = dlf.unsqueeze(-1) * torch.sigmoid(-fg)
    return dq, dk, dv, di.unsqueeze(-1), dfg


# --------------------------------------------------------------------------
# K2/K3: the ViL block around the cell
# --------------------------------------------------------------------------
def headwise_linear(x, w):
    """LinearHeadwiseExpand.forward, no bias -- vision_lstm.py:158-168.
    w: (nblocks, out_d, d); x: (..., nblocks*d)."""
    nb, od, d = w.shape
    xb = x.reshape(*x.shape[:-1], nb, d)
    y = torch.einsum("...nd,nod->...no", xb, w)
    return y.reshape(*x.shape[:-1], nb * od)


def causal_conv1d(x, w, b):
    """CausalConv1d.forward -- vision_lstm.py:213-221.  x: (B,S,E); w: (E,1,K); b: (E,).
    y_t = b + sum_j w_j x_{t-(K-1)+j}, zeros before t=0."""
    K = w.shape[-1]
    xp = F.pad(x.transpose(1, 2), (K - 1, 0))
    y = F.conv1d(xp, w, b, groups=x.shape[-1])
    return y.transpose(1, 2)


def multihead_layernorm(h, w, eps: float = 1e-5):
    """MultiHeadLayerNorm.forward -- vision_lstm.py:271-287.  h: (B,NH,S,DH);
    per token and head normalise over DH; weight (NH*DH) applied as 1+w, no bias."""
    B, NH, S, DH = h.shape
    mu = h.mean(-1, keepdim=True)
    var = h.var(-1, unbiased=False, keepdim=True)
    y = (h - mu) / torch.sqrt(var + eps)
    y = y * (1.0 + w).reshape(1, NH, 1, DH)
    return y


def vil_block(x, p: dict, reverse: bool = False, cell=mlstm_parallel, return_intermediates: bool = False):
    """ViLBlock.forward = x + layer(norm(x)) -- vision_lstm.py:494-502 with
    DropPath(0) (vision_lstm_util.py:168-175), LayerNorm weight 1+w, eps 1e-5
    (224-268), inner ViLLayer.forward (415-453), MatrixLSTMCell.forward (302-339).

    x: (B,S,C).  ``p`` uses the reference's state_dict keys relative to the
    ViLBlock: norm.weight, layer.proj_up.weight, layer.{q,k,v}_proj.weight,
    layer.conv1d.conv.{weight,bias}, layer.mlstm_cell.{igate,fgate}.{weight,bias},
    layer.mlstm_cell.outnorm.weight, layer.learnable_skip, layer.proj_down.weight.
    ``reverse`` = SequenceTraversal.ROWWISE_FROM_BOT_RIGHT (419-424, 446-451).
    """
    B, S, Cdim = x.shape
    xn = F.layer_norm(x, (Cdim,), weight=1.0 + p["norm.weight"], bias=None, eps=1e-5)
    if reverse:
        xn = xn.flip(dims=[1])
    inner = xn @ p["layer.proj_up.weight"].t()
    E = inner.shape[-1] // 2
    x_m, z = inner[..., :E], inner[..., E:]
    conv = causal_conv1d(x_m, p["layer.conv1d.conv.weight"], p["layer.conv1d.conv.bias"])
    act = F.silu(conv)
    qf = headwise_linear(act, p["layer.q_proj.weight"])
    kf = headwise_linear(act, p["layer.k_proj.weight"])
    vf = headwise_linear(x_m, p["layer.v_proj.weight"])
    NH = p["layer.mlstm_cell.igate.weight"].shape[0]
    DH = E // NH
    gate_in = torch.cat([qf, kf, vf], dim=-1)
    ig = gate_in @ p["layer.mlstm_cell.igate.weight"].t() + p["layer.mlstm_cell.igate.bias"]
    fg = gate_in @ p["layer.mlstm_cell.fgate.weight"].t() + p["layer.mlstm_cell.fgate.bias"]
    to_heads = lambda t: t.reshape(B, S, NH, DH).transpose(1, 2)
    q, k, v = to_heads(qf), to_heads(kf), to_heads(vf)
    ig4 = ig.transpose(-1, -2).unsqueeze(-1)
    fg4 = fg.transpose(-1, -2).unsqueeze(-1)
    h = cell(q, k, v, ig4, fg4)
    hn = multihead_layernorm(h, p["layer.mlstm_cell.outnorm.weight"])
    hn = hn.transpose(1, 2).reshape(B, S, E)
    hs = hn + p["layer.learnable_skip"] * act
    out = (hs * F.silu(z)) @ p["layer.proj_down.weight"].t()
    if reverse:
        out = out.flip(dims=[1])
    y = x + out
    if return_intermediates:
        return y, dict(q=q, k=k, v=v, ig=ig4, fg=fg4, h=h, act=act, z=z)
    return y


def vil_wrapper(x, p: dict, cell=mlstm_parallel):
    """Outer 3-D ViLLayer.forward / forward_patch_token -- UxLSTMEnc_3d.py:54-63,77-87.
    x: (B,C,*spatial) -> same.  ``p`` keys are relative to the wrapper
    (vil.norm.weight, vil.layer....); its own ``norm.*`` parameters are dead."""
    B, Cdim = x.shape[:2]
    sp = x.shape[2:]
    tok = x.reshape(B, Cdim, -1).transpose(-1, -2)
    pp = {kk[len("vil."):]: vv for kk, vv in p.items() if kk.startswith("vil.")}
    y = vil_block(tok.to(torch.float32) if x.dtype == torch.float16 else tok, pp, reverse=False, cell=cell)
    return y.transpose(-1, -2).reshape(B, Cdim, *sp)


# --------------------------------------------------------------------------
# K4: S-MVAE product of experts, clip, reparametrisation, KL
# --------------------------------------------------------------------------
def clip_logvar(x):
    """clip -- RA_HVED.py:749-753."""
    return torch.clamp(x, min=-50.0, max=50.0)


def poe(mu, logvar, mod_list, eps: float = 1e-8):
    """ProductOfExperts.forward -- buildingblocks.py:853-866 (dup loss.py:49-63).
    mu, logvar: (5,B,...) with index 0 the prior; mod_list: tuple of ints in 0..3."""
    sel = [m + 1 for m in mod_list] + [0]
    lv = torch.stack([logvar[i] for i in sel], 0)
    mm = torch.stack([mu[i] for i in sel], 0)
    T = 1.0 / (torch.exp(lv) + eps)
    Tsum = T.sum(0)
    return (mm * T).sum(0) / Tsum, torch.log(1.0 / Tsum)


def poe_drop(mu, logvar, drop, eps: float = 1e-8):
    """ProductOfExperts2.forward -- buildingblocks.py:875-886 (ZeroLayerF 308-323).
    drop: (B,4) bool, True = modality missing for that sample.  Returns
    (pd_mu, pd_logvar, mu_after) where mu_after is ``mu`` as the reference leaves
    it (it overwrites mu[m+1] in place with the zeroed copy)."""
    T = 1.0 / (torch.exp(logvar) + eps)
    mu = mu.clone()
    T = T.clone()
    for m in range(drop.shape[1]):
        mu[m + 1][drop[:, m]] = 0
        T[m + 1][drop[:, m]] = 0
    Tsum = T.sum(0)
    return (mu * T).sum(0) / Tsum, torch.log(1.0 / Tsum), mu


def reparametrize(mu, logvar, noise=None, valid: bool = False):
    """reparametrize -- RA_HVED.py:741-747.  ``noise`` is the N(0,1) draw the
    reference takes from the global generator at :744."""
    if valid:
        return mu
    return noise * torch.exp(0.5 * logvar) + mu


def kl_to_prior(mu1, logvar1):
    """KL_divergence(mu1, logvar1, prior) -- loss.py:29-40 as called from
    compute_KLD (loss.py:110): mu2=0, logvar2=0 passed explicitly, so eps=1e-8."""
    return 0.5 * torch.mean(-1.0 - logvar1 + (logvar1.exp() + mu1 * mu1) / (1.0 + 1e-8))


def kl_divergence(mu1, logvar1, mu2, logvar2, eps: float = 1e-8):
    """KL_divergence(mu1, logvar1, mu2, logvar2) -- loss.py:29-40 with an explicit second distribution (eps = 1e-8)."""
    return 0.5 * torch.mean(-1.0 + logvar2 - logvar1 + (logvar1.exp() + (mu1 - mu2) ** 2) / (logvar2.exp() + eps))


def compute_kld(mu_b5, logvar_b5, subset_index_list=(14,)):
    """compute_KLD -- loss.py:85-115.  Inputs are (B,5,...) as returned in the
    model's mu_list/logvar_list (RA_HVED.py:582-583).  The prior handed to KL_divergence is slab 0 of the inputs
    (loss.py:95-97, 113) -- the standard normal for XLSTM_HVED, but whatever the caller put there in general."""
    mu = mu_b5.transpose(1, 0)
    lv = logvar_b5.transpose(1, 0)
    tot, cnt = 0.0, 0
    for idx, subset in enumerate(SUBSETS_MODALITIES):
        if idx in subset_index_list:
            cnt += 1
            smu, slv = poe(mu, lv, subset)
            tot = tot + kl_divergence(smu, slv, mu[0], lv[0])
    return tot / cnt


def dice_per_channel(inp, target, epsilon: float = 1e-6, weight=None):
    """compute_per_channel_dice -- loss.py:257-285 (flatten: 287-301): per channel over batch and voxels,
    2 sum(p t) / clamp(sum p^2 + sum t^2, eps), the optional weight on the intersection."""
    C = inp.shape[1]
    p = inp.transpose(0, 1).reshape(C, -1)
    t = target.transpose(0, 1).reshape(C, -1).to(p.dtype)
    inter = (p * t).sum(-1)
    if weight is not None:
        inter = weight * inter
    return 2 * (inter / ((p * p).sum(-1) + (t * t).sum(-1)).clamp(min=epsilon))


def dice_loss(inp, target, weight=None):
    """DiceLoss.forward -- loss.py:201-209: 1 - mean over channels (the input is NOT normalised: line 203 is commented out)."""
    return 1.0 - torch.mean(dice_per_channel(inp, target, weight=weight))


def zero_rows(x, alpha):
    """ZeroLayerF.forward (and .backward applied to the gradient) -- buildingblocks.py:308-323."""
    y = x.clone()
    y[alpha] = 0
    return y


def instance_norm_lrelu(x, weight=None, bias=None, eps: float = 1e-5, slope: float = 1.0):
    """nn.InstanceNorm3d followed by nn.LeakyReLU(slope) -- SingleConv order 'ilc' (buildingblocks.py:414-416, 430-431; the
    Sequential runs them in order, 462) and BasicConv.forward (buildingblocks.py:23-31).  The arithmetic lives in PyTorch
    (torch.nn.functional.instance_norm, torch 2.11: affine=False, track_running_stats=False by default): per (sample, channel)
    plane, y = (x - mean) / sqrt(biased var + eps) [* weight + bias]; LeakyReLU: v > 0 ? v : slope v.  slope = 1: no activation."""
    N, C = x.shape[:2]
    flat = x.reshape(N, C, -1)
    mean = flat.mean(-1, keepdim=True)
    var = ((flat - mean) ** 2).mean(-1, keepdim=True)
    y = (flat - mean) / torch.sqrt(var + eps)
    if weight is not None:
        y = y * weight.reshape(1, C, 1)
    if bias is not None:
        y = y + bias.reshape(1, C, 1)
    y = torch.where(y > 0, y, y * slope)
    return y.reshape(x.shape)


def batch_norm_lrelu(x, weight, bias, running_mean=None, running_var=None, training: bool = True, momentum: float = 0.1,
                     eps: float = 1e-5, slope: float = 1.0):
    """nn.BatchNorm3d (modules/DuSFE.py:108-110 used at 151-152; 17-36, 187) [+ LeakyReLU].  torch.nn.functional.batch_norm:
    training = statistics per channel over (N, *spatial) with the biased variance, running statistics moved by ``momentum``
    towards the batch mean / UNBIASED variance; eval = running statistics.  Returns (y, new_running_mean, new_running_var)."""
    C = x.shape[1]
    flat = x.transpose(0, 1).reshape(C, -1)
    if training:
        mean = flat.mean(-1)
        var = ((flat - mean[:, None]) ** 2).mean(-1)
        n = flat.shape[1]
        if running_mean is not None:
            running_mean = (1 - momentum) * running_mean + momentum * mean.detach()
            running_var = (1 - momentum) * running_var + momentum * var.detach() * n / max(n - 1, 1)
    else:
        mean, var = running_mean, running_var
    shape = [1, C] + [1] * (x.dim() - 2)
    y = (x - mean.reshape(shape)) / torch.sqrt(var.reshape(shape) + eps)
    y = y * weight.reshape(shape) + bias.reshape(shape)
    y = torch.where(y > 0, y, y * slope)
    return y, running_mean, running_var


def basic_conv_depthwise(x, weight, slope: float = 0.01):
    """BasicConv(C, C, 3, padding=1, groups=C).forward -- buildingblocks.py:23-31 as instantiated at RA_HVED.py:406: depthwise
    3x3x3 convolution without bias (zero padding 1), InstanceNorm3d, LeakyReLU(0.01).  Written out tap by tap."""
    import torch.nn.functional as F
    C = x.shape[1]
    xp = F.pad(x, (1, 1, 1, 1, 1, 1))
    D, H, W = x.shape[2:]
    y = torch.zeros_like(x)
    for kd in range(3):
        for kh in range(3):
            for kw in range(3):
                y = y + xp[:, :, kd:kd + D, kh:kh + H, kw:kw + W] * weight[:, 0, kd, kh, kw].reshape(1, C, 1, 1, 1)
    return instance_norm_lrelu(y, slope=slope), y


def channel_pool(x):
    """ChannelPool.forward -- buildingblocks.py:136-138: (max over channels, mean over channels)."""
    return torch.cat((x.max(1)[0].unsqueeze(1), x.mean(1).unsqueeze(1)), dim=1)


def atten_module2(seg_x, enc_x, p: dict):
    """AttenModule2.forward -- buildingblocks.py:277-301 (recon_x = None, the only way the model calls it): two spatial gates,
    each a depthwise 7x7x7 convolution (4 output channels per input channel, buildingblocks.py:271, 273) followed by a 1x1x1
    convolution to one channel (272, 274) and a sigmoid; p = the module's state_dict."""
    import torch.nn.functional as F
    spa = channel_pool(seg_x)
    enc_spa = torch.cat([spa, channel_pool(enc_x)], 1)
    e = F.conv3d(enc_spa, p["enc_spatial.weight"], p["enc_spatial.bias"], padding=3, groups=enc_spa.shape[1])
    e = torch.sigmoid(F.conv3d(e, p["enc_spatial2.weight"], p["enc_spatial2.bias"]))
    s_enc = enc_x + enc_x * e
    g = F.conv3d(spa, p["seg_spatial.weight"], p["seg_spatial.bias"], padding=3, groups=spa.shape[1])
    g = torch.sigmoid(F.conv3d(g, p["seg_spatial2.weight"], p["seg_spatial2.bias"]))
    return torch.cat([seg_x * (1 + g), s_enc], 1)


def poe_backward(mu, logvar, mod_list, g_mu, g_lv, eps: float = 1e-8):
    """Manual backward of :func:`poe` w.r.t. the 4 modality experts (SURVEY.md
    8a-note).  Returns (dmu, dlogvar) of shape (4,...) -- zeros for experts not
    in mod_list."""
    sel = [m + 1 for m in mod_list] + [0]
    T = {i: 1.0 / (torch.exp(logvar[i]) + eps) for i in sel}
    Ssum = sum(T.values())
    mu_hat = sum(mu[i] * T[i] for i in sel) / Ssum
    dmu = torch.zeros_like(mu[1:])
    dlv = torch.zeros_like(mu[1:])
    for m in mod_list:
        i = m + 1
        dmu[m] = g_mu * T[i] / Ssum
        dT = g_mu * (mu[i] - mu_hat) / Ssum - g_lv / Ssum
        dlv[m] = -dT * T[i] * T[i] * torch.exp(logvar[i])
    return dmu, dlv


# --------------------------------------------------------------------------
# Precision model of the CUDA cell (what "ideal bf16 operands" costs)
# --------------------------------------------------------------------------
def _bf16(x):
    return x.to(torch.bfloat16).to(x.dtype)


def mlstm_forward_backward_bf16_operands(q, k, v, ig, fg, dh, eps: float = 1e-6, den_from_rounded_p: bool = False):
    """vision_lstm.py:48-130 and its gradient with exactly the operands the sm_100a kernels feed to the
    tensor cores rounded to bf16 (q, k, v, P = S o D', h, dh/N, db, dS) and everything else exact.
    Used by the tests to separate kernel bugs (kernel != this) from the unavoidable cost of bf16
    operands in ill-conditioned regimes (this != fp64 reference).
    den_from_rounded_p: the warp-specialised forward takes the normaliser input from the ones column of the P [V|1]
    product, i.e. it sums the bf16-ROUNDED P -- consistent with the numerator -- instead of the fp32 values."""
    B, NH, S, DH = q.shape
    scale = 1.0 / math.sqrt(DH)
    q, k, v, dh = _bf16(q), _bf16(k), _bf16(v), _bf16(dh)
    lf = F.logsigmoid(fg)
    c = torch.cumsum(lf, dim=-2)
    logD = c - c.transpose(-2, -1) + ig.transpose(-2, -1)
    tri = torch.tril(torch.ones(S, S, dtype=torch.bool, device=q.device))
    logD = torch.where(tri, logD, torch.full_like(logD, -float("inf")))
    m = logD.max(dim=-1, keepdim=True).values
    Dm = torch.exp(logD - m) * scale
    C = (q @ k.transpose(-2, -1)) * Dm
    Cb = _bf16(C)
    den = (Cb if den_from_rounded_p else C).sum(-1, keepdim=True)
    floor = torch.exp(-m)
    N = torch.maximum(den.abs(), floor) + eps
    h = _bf16((Cb @ v) / N)
    dhh = (dh * h).sum(-1, keepdim=True)
    dn = -dhh / N
    db = _bf16(torch.where(den.abs() > floor, dn * torch.sign(den), torch.zeros_like(dn)))
    Gv = _bf16(dh / N)
    dS = _bf16((Gv @ v.transpose(-2, -1) + db) * Dm)
    dq, dk, dv = dS @ k, dS.transpose(-2, -1) @ q, Cb.transpose(-2, -1) @ Gv
    di = (k * dk).sum(-1)
    dc = (q * dq).sum(-1) - di
    dlf = torch.flip(torch.cumsum(torch.flip(dc, dims=[-1]), dim=-1), dims=[-1])
    return h, (dq, dk, dv, di.unsqueeze(-1), dlf.unsqueeze(-1) * torch.sigmoid(-fg))
