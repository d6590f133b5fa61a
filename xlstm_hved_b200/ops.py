"""Host-side ops over the C ABI (include/xhved.h): allocation with torch, launch through ctypes.

Nothing here computes on the CPU or with torch math: every op launches the
sm_100a kernels of libxhved.so on the current CUDA stream.
"""
from __future__ import annotations

import collections
import ctypes
import os
import threading
from ctypes import c_float, c_uint32

import torch

from . import _lib
from ._lib import check, ptr, stream

CHUNK = 128
# copies of the parameter-gradient block the backward kernels scatter their atomics over (XHVED_GRAD_REPLICAS to tune)
GRAD_REPLICAS = max(1, int(os.environ.get("XHVED_GRAD_REPLICAS", "32")))

# RA_HVED.py:733-738 -- subset index -> modalities present
SUBSETS_MODALITIES = [(0,), (1,), (2,), (3,), (0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3),
                      (0, 1, 2), (0, 1, 3), (0, 2, 3), (1, 2, 3), (0, 1, 2, 3)]


def padded_dh(dh: int) -> int:
    for p in (16, 32, 64, 128):
        if dh <= p:
            return p
    raise RuntimeError(f"head dim {dh} > 128 is not supported by the sm_100a mLSTM kernels")


def _f32c(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


# ----------------------------------------------------------------------------- S-MVAE
def _masks(subsets):
    arr = (c_uint32 * len(subsets))()
    for i, s in enumerate(subsets):
        m = 0
        for mod in s:
            if not 0 <= int(mod) <= 3:
                raise ValueError(f"modality index {mod} outside 0..3")
            m |= 1 << int(mod)
        arr[i] = m
    return arr


POE_STANDARD_PRIOR = 1        # include/xhved.h: XHVED_POE_STANDARD_PRIOR


def poe_fwd(mu5: torch.Tensor, logvar5: torch.Tensor, subsets, drop=None, noise=None, want_kld=False, eps: float = 1e-8,
            kld_out=None, standard_prior: bool = False):
    """All requested subsets in one launch.  mu5/logvar5: (5, ...) fp32 contiguous (prior first).
    Returns (pd_mu, pd_logvar, z or None, kld_sums or None) each of shape (len(subsets), ...).
    kld_out: optional pre-zeroed fp32 buffer of len(subsets) elements the KL sums are accumulated into (lets a caller
    collect the sums of several latent levels in one tensor).
    standard_prior: the caller guarantees that slab 0 is the model's constant prior (mu = 0, logvar = 0,
    RA_HVED.py:576-580); the kernel then does not read it (20 % fewer input bytes)."""
    lib = _lib.load_library()
    assert mu5.shape == logvar5.shape and mu5.shape[0] == 5
    mu5, logvar5 = _f32c(mu5), _f32c(logvar5)
    n = mu5[0].numel()
    ns = len(subsets)
    out_mu = torch.empty((ns, *mu5.shape[1:]), device=mu5.device, dtype=torch.float32)
    out_lv = torch.empty_like(out_mu)
    z = torch.empty_like(out_mu) if noise is not None else None
    if kld_out is not None:
        assert kld_out.dtype == torch.float32 and kld_out.is_contiguous() and kld_out.numel() == ns and kld_out.device == mu5.device
        kld = kld_out
    else:
        kld = torch.zeros(ns, device=mu5.device, dtype=torch.float32) if want_kld else None
    per_sample = 0
    if drop is not None:
        drop = drop.to(device=mu5.device, dtype=torch.uint8).contiguous()
        per_sample = n // drop.shape[0]
    if noise is not None:
        noise = _f32c(noise)
        assert noise.numel() == ns * n
    check(lib.xhved_poe_fwd(ptr(mu5), ptr(logvar5), n, n, _masks(subsets), ns, ptr(drop), per_sample, eps, ptr(out_mu),
                            ptr(out_lv), ptr(noise), ptr(z), ptr(kld), POE_STANDARD_PRIOR if standard_prior else 0, stream()),
          "xhved_poe_fwd")
    return out_mu, out_lv, z, kld


def poe_bwd(mu5, logvar5, subsets, g_mu=None, g_logvar=None, noise=None, g_z=None, kld_scale=None, drop=None, eps: float = 1e-8,
            standard_prior: bool = False):
    """Gradients of poe_fwd w.r.t. the expert slabs.  Returns (d_mu, d_logvar) of shape (5, ...); with standard_prior=True
    the constant prior slab is neither read nor given a gradient and the results have shape (4, ...) (modalities only)."""
    lib = _lib.load_library()
    mu5, logvar5 = _f32c(mu5), _f32c(logvar5)
    n = mu5[0].numel()
    ns = len(subsets)
    d_mu = torch.empty_like(mu5)
    d_lv = torch.empty_like(mu5)
    per_sample = 0
    if drop is not None:
        drop = drop.to(device=mu5.device, dtype=torch.uint8).contiguous()
        per_sample = n // drop.shape[0]
    ks = None
    if kld_scale is not None:
        ks = (c_float * ns)(*[float(v) for v in kld_scale])
    g_mu = _f32c(g_mu) if g_mu is not None else None
    g_logvar = _f32c(g_logvar) if g_logvar is not None else None
    g_z = _f32c(g_z) if g_z is not None else None
    noise = _f32c(noise) if noise is not None else None
    check(lib.xhved_poe_bwd(ptr(mu5), ptr(logvar5), n, n, _masks(subsets), ns, ptr(drop), per_sample, eps, ptr(g_mu), ptr(g_logvar),
                            ptr(noise), ptr(g_z), ks, ptr(d_mu), ptr(d_lv), POE_STANDARD_PRIOR if standard_prior else 0, stream()),
          "xhved_poe_bwd")
    # with a standard prior slab 0 is never written: hand back the modality slabs only
    return (d_mu[1:], d_lv[1:]) if standard_prior else (d_mu, d_lv)


def _dp(t):
    return t.data_ptr() if t is not None else None


POE_CLIP = 2                  # include/xhved.h: XHVED_POE_CLIP


def _poe_opts(eps, standard_prior, clip):
    """xhved_poe_opts: clip = None or (lo, hi) -- the model's clamp of the modality logvars (RA_HVED.py:580, 749-753) fused
    into the launch (the slabs then hold RAW logvars; the backward zeroes d_logvar outside the interval)."""
    o = _lib.PoeOpts()
    o.eps = eps
    o.flags = (POE_STANDARD_PRIOR if standard_prior else 0) | (POE_CLIP if clip is not None else 0)
    o.clip_lo, o.clip_hi = (float(clip[0]), float(clip[1])) if clip is not None else (0.0, 0.0)
    return o


def poe_fwd_levels(levels, subsets, noises=None, kld_out=None, eps: float = 1e-8, standard_prior: bool = False, clip=None,
                   drops=None):
    """poe_fwd for several latent levels in ONE launch (the four latent resolutions of a volume; the small ones are
    latency-bound on their own).  levels: list of (mu5, logvar5) with shape (5, ...) each; noises: optional list of
    (len(subsets), ...) tensors; kld_out: optional pre-zeroed fp32 (n_levels, len(subsets)) buffer; clip: None or
    (lo, hi), see _poe_opts; drops: optional per-level (B, 4) uint8 missing flags (ProductOfExperts2).
    Returns a list of (pd_mu, pd_logvar, z or None) per level."""
    lib = _lib.load_library()
    ns, nl = len(subsets), len(levels)
    arr = (_lib.PoeLevel * nl)()
    keep, outs = [], []
    for l, (mu5, lv5) in enumerate(levels):
        assert mu5.shape == lv5.shape and mu5.shape[0] == 5
        mu5, lv5 = _f32c(mu5), _f32c(lv5)
        n = mu5[0].numel()
        out_mu = torch.empty((ns, *mu5.shape[1:]), device=mu5.device, dtype=torch.float32)
        out_lv = torch.empty_like(out_mu)
        noise = _f32c(noises[l]) if noises is not None else None
        z = torch.empty_like(out_mu) if noise is not None else None
        if noise is not None:
            assert noise.numel() == ns * n
        kl = kld_out[l] if kld_out is not None else None
        if kl is not None:
            assert kl.dtype == torch.float32 and kl.is_contiguous() and kl.numel() == ns
        drop = drops[l].to(device=mu5.device, dtype=torch.uint8).contiguous() if drops is not None and drops[l] is not None else None
        e = arr[l]
        e.mu, e.logvar, e.n, e.expert_stride = mu5.data_ptr(), lv5.data_ptr(), n, n
        e.drop, e.per_sample = _dp(drop), (n // drop.shape[0] if drop is not None else 0)
        e.out_mu, e.out_logvar, e.noise, e.out_z, e.kld_out = out_mu.data_ptr(), out_lv.data_ptr(), _dp(noise), _dp(z), _dp(kl)
        keep.append((mu5, lv5, noise, drop))
        outs.append((out_mu, out_lv, z))
    opts = _poe_opts(eps, standard_prior, clip)
    check(lib.xhved_poe_fwd_levels_opts(arr, nl, _masks(subsets), ns, ctypes.byref(opts), stream()), "xhved_poe_fwd_levels_opts")
    return outs


def poe_bwd_levels(levels, subsets, noises=None, g_zs=None, kld_scales=None, eps: float = 1e-8, standard_prior: bool = False,
                   clip=None, g_mus=None, g_logvars=None, drops=None):
    """Backward of poe_fwd_levels through z, the KL term and (optionally) the fused outputs themselves, one launch.
    kld_scales: per level a list of len(subsets) floats; g_mus / g_logvars: optional per-level (len(subsets), ...) upstream
    gradients of pd_mu / pd_logvar.  Returns a list of (d_mu, d_logvar) per level ((4, ...) modality slabs with
    standard_prior, else (5, ...))."""
    lib = _lib.load_library()
    ns, nl = len(subsets), len(levels)
    arr = (_lib.PoeLevelGrad * nl)()
    keep, outs = [], []
    for l, (mu5, lv5) in enumerate(levels):
        mu5, lv5 = _f32c(mu5), _f32c(lv5)
        n = mu5[0].numel()
        d_mu, d_lv = torch.empty_like(mu5), torch.empty_like(mu5)
        noise = _f32c(noises[l]) if noises is not None else None
        g_z = _f32c(g_zs[l]) if g_zs is not None else None
        ks = (c_float * ns)(*[float(v) for v in kld_scales[l]]) if kld_scales is not None else None
        g_mu = _f32c(g_mus[l]) if g_mus is not None and g_mus[l] is not None else None
        g_lv = _f32c(g_logvars[l]) if g_logvars is not None and g_logvars[l] is not None else None
        drop = drops[l].to(device=mu5.device, dtype=torch.uint8).contiguous() if drops is not None and drops[l] is not None else None
        e = arr[l]
        e.mu, e.logvar, e.n, e.expert_stride = mu5.data_ptr(), lv5.data_ptr(), n, n
        e.drop, e.per_sample = _dp(drop), (n // drop.shape[0] if drop is not None else 0)
        e.g_mu, e.g_logvar, e.noise, e.g_z = _dp(g_mu), _dp(g_lv), _dp(noise), _dp(g_z)
        if ks is not None:
            e.kld_scale = ks
        e.d_mu, e.d_logvar = d_mu.data_ptr(), d_lv.data_ptr()
        keep.append((mu5, lv5, noise, g_z, ks, g_mu, g_lv, drop))
        outs.append((d_mu[1:], d_lv[1:]) if standard_prior else (d_mu, d_lv))
    opts = _poe_opts(eps, standard_prior, clip)
    check(lib.xhved_poe_bwd_levels_opts(arr, nl, _masks(subsets), ns, ctypes.byref(opts), stream()), "xhved_poe_bwd_levels_opts")
    return outs


def clip_fwd(x, lo: float = -50.0, hi: float = 50.0):
    """clip (RA_HVED.py:749-753): clamp(x, lo, hi) on the device."""
    lib = _lib.load_library()
    x = _f32c(x)
    y = torch.empty_like(x)
    if x.numel():
        check(lib.xhved_clip_fwd(ptr(x), x.numel(), lo, hi, ptr(y), stream()), "xhved_clip_fwd")
    return y


def clip_bwd(x, g, lo: float = -50.0, hi: float = 50.0):
    lib = _lib.load_library()
    x, g = _f32c(x), _f32c(g)
    dx = torch.empty_like(x)
    if x.numel():
        check(lib.xhved_clip_bwd(ptr(x), ptr(g), x.numel(), lo, hi, ptr(dx), stream()), "xhved_clip_bwd")
    return dx


def zero_rows(x, mask):
    """ZeroLayerF (buildingblocks.py:308-323): a copy of x whose rows x[b] with mask[b] set are zero.  The reference's
    backward is the same map applied to the gradient."""
    lib = _lib.load_library()
    x = _f32c(x)
    mask = mask.to(device=x.device, dtype=torch.uint8).contiguous()
    if mask.dim() != 1 or mask.shape[0] != x.shape[0]:
        raise ValueError("zero_rows: mask must be a (B,) boolean vector over the leading dimension of x")
    y = torch.empty_like(x)
    if x.numel():
        check(lib.xhved_zero_rows(ptr(x), ptr(mask), x.shape[0], x.numel() // x.shape[0], ptr(y), stream()), "xhved_zero_rows")
    return y


def dice_sums(p, t):
    """{sum p t, sum p^2, sum t^2} per channel of (N, C, *spatial) fp32 tensors, one pass.  Returns a (C, 3) tensor."""
    lib = _lib.load_library()
    p, t = _f32c(p), _f32c(t)
    N, C = p.shape[:2]
    spatial = p[0, 0].numel()
    sums = torch.zeros(C, 3, device=p.device, dtype=torch.float32)
    check(lib.xhved_dice_sums(ptr(p), ptr(t), N, C, spatial, ptr(sums), stream()), "xhved_dice_sums")
    return sums


def dice_bwd(p, t, sums, g_dice, eps: float = 1e-6):
    lib = _lib.load_library()
    p, t = _f32c(p), _f32c(t)
    N, C = p.shape[:2]
    dp = torch.empty_like(p)
    check(lib.xhved_dice_bwd(ptr(p), ptr(t), ptr(sums), ptr(_f32c(g_dice)), N, C, p[0, 0].numel(), eps, ptr(dp), stream()), "xhved_dice_bwd")
    return dp


def reparam_fwd(mu, logvar, noise):
    lib = _lib.load_library()
    mu, logvar, noise = _f32c(mu), _f32c(logvar), _f32c(noise)
    z = torch.empty_like(mu)
    check(lib.xhved_reparam_fwd(ptr(mu), ptr(logvar), ptr(noise), mu.numel(), ptr(z), stream()), "xhved_reparam_fwd")
    return z


def reparam_bwd(logvar, noise, g_z):
    lib = _lib.load_library()
    logvar, noise, g_z = _f32c(logvar), _f32c(noise), _f32c(g_z)
    d_mu, d_lv = torch.empty_like(g_z), torch.empty_like(g_z)
    check(lib.xhved_reparam_bwd(ptr(logvar), ptr(noise), ptr(g_z), g_z.numel(), ptr(d_mu), ptr(d_lv), stream()), "xhved_reparam_bwd")
    return d_mu, d_lv


# ----------------------------------------------------------------------------- InstanceNorm3d / BatchNorm3d + LeakyReLU (K6)
NORM_INSTANCE, NORM_BATCH, NORM_FROZEN = 0, 1, 2
_NORM_DTYPES = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}
_NORM_PLANS = {}         # (N, C, spatial, dtype, mode, eps, slope) -> NormPlan: the shape struct and the host-only workspace query, cached


class NormPlan:
    """Everything about one normalisation call that depends on the shape only: the C struct (never modified afterwards, so shared
    between calls and threads), the number of statistics groups, the size of the partials scratch in floats."""
    __slots__ = ("sh", "ref", "groups", "ws_floats", "C", "mode")

    def __init__(self, N, C, spatial, dtype, mode, eps, slope):
        sh = _lib.NormShape()
        sh.N, sh.C, sh.spatial, sh.mode, sh.dtype, sh.eps, sh.slope = N, C, spatial, mode, _NORM_DTYPES[dtype], eps, slope
        self.sh, self.ref, self.C, self.mode = sh, ctypes.byref(sh), C, mode
        self.groups = N * C if mode == NORM_INSTANCE else C
        self.ws_floats = (_lib.load_library().xhved_norm_act_workspace(N, C, spatial, sh.dtype) + 3) // 4


def norm_plan(x, mode, eps, slope) -> NormPlan:
    shape = x.shape
    spatial = 1
    for d in shape[2:]:
        spatial *= d
    key = (shape[0], shape[1], spatial, x.dtype, mode, eps, slope)
    plan = _NORM_PLANS.get(key)
    if plan is None:
        if x.dtype not in _NORM_DTYPES:
            raise RuntimeError(f"xlstm_hved_b200 norm_act: unsupported dtype {x.dtype}")
        if len(_NORM_PLANS) > 4096:
            _NORM_PLANS.clear()
        plan = _NORM_PLANS[key] = NormPlan(shape[0], shape[1], spatial, x.dtype, mode, eps, slope)
    return plan


def norm_act_fwd_raw(x, gamma, beta, plan: NormPlan, mean=None, rstd=None):
    """The lean form behind the autograd function: x contiguous CUDA; gamma / beta fp32 contiguous or None.  Returns (y, stats):
    ONE fp32 buffer holding mean [groups] | rstd [groups] | the partials scratch (frozen statistics: stats is None, mean / rstd are
    the given tensors)."""
    lib = _lib.load_library()
    y = torch.empty_like(x)
    g = plan.groups
    if plan.mode == NORM_FROZEN:
        stats, pm, pr, pp = None, mean.data_ptr(), rstd.data_ptr(), None
    else:
        stats = torch.empty(2 * g + plan.ws_floats, device=x.device, dtype=torch.float32)
        pm = stats.data_ptr()
        pr, pp = pm + 4 * g, pm + 8 * g
    check(lib.xhved_norm_act_fwd(x.data_ptr(), gamma.data_ptr() if gamma is not None else None, beta.data_ptr() if beta is not None else None,
                                 plan.ref, pm, pr, pp, y.data_ptr(), stream()), "xhved_norm_act_fwd")
    return y, stats


def norm_act_bwd_raw(x, dy, gamma, beta, plan: NormPlan, stats=None, mean=None, rstd=None, want_param_grads: bool = False):
    """Backward of norm_act_fwd_raw: `stats` as returned there (its partials region is reused as this call's scratch), or mean / rstd
    tensors for frozen statistics.  Returns (dx, dgb): dgb = fp32 (2, C) = (dgamma, dbeta) or None."""
    lib = _lib.load_library()
    dx = torch.empty_like(x)
    g = plan.groups
    if stats is not None:
        pm = stats.data_ptr()
        pr, pp, scratch = pm + 4 * g, pm + 8 * g, None
    else:
        scratch = torch.empty(max(plan.ws_floats, 1), device=x.device, dtype=torch.float32)
        pm, pr, pp = mean.data_ptr(), rstd.data_ptr(), scratch.data_ptr()
    dgb = torch.zeros(2, plan.C, device=x.device, dtype=torch.float32) if want_param_grads else None
    pg = dgb.data_ptr() if want_param_grads else None
    check(lib.xhved_norm_act_bwd(x.data_ptr(), dy.data_ptr(), gamma.data_ptr() if gamma is not None else None,
                                 beta.data_ptr() if beta is not None else None, pm, pr, plan.ref, pp, dx.data_ptr(), pg,
                                 pg + 4 * plan.C if want_param_grads else None, stream()), "xhved_norm_act_bwd")
    return dx, dgb


def norm_act_fwd(x, gamma=None, beta=None, mode: int = NORM_INSTANCE, eps: float = 1e-5, slope: float = 1.0, mean=None, rstd=None):
    """x: (N, C, *spatial) fp32 / fp16 / bf16.  Returns (y, mean, rstd): the statistics are fp32 of N*C (instance) or C (batch)
    entries; with mode = NORM_FROZEN they are inputs (C entries).  slope = 1: no activation."""
    if not x.is_cuda:
        raise RuntimeError("xlstm_hved_b200 has no CPU path")
    x = x.contiguous()
    plan = norm_plan(x, mode, eps, slope)
    gamma, beta = (_f32c(gamma) if gamma is not None else None), (_f32c(beta) if beta is not None else None)
    if mode == NORM_FROZEN:
        mean, rstd = _f32c(mean), _f32c(rstd)
        return norm_act_fwd_raw(x, gamma, beta, plan, mean, rstd)[0], mean, rstd
    y, stats = norm_act_fwd_raw(x, gamma, beta, plan)
    return y, stats[:plan.groups], stats[plan.groups:2 * plan.groups]


def norm_act_bwd(x, dy, mean, rstd, gamma=None, beta=None, mode: int = NORM_INSTANCE, eps: float = 1e-5, slope: float = 1.0,
                 want_param_grads: bool = False):
    """Backward of norm_act_fwd.  Returns (dx, dgamma, dbeta); the last two are None unless want_param_grads."""
    x = x.contiguous()
    dy = dy.to(x.dtype).contiguous()
    plan = norm_plan(x, mode, eps, slope)
    gamma, beta = (_f32c(gamma) if gamma is not None else None), (_f32c(beta) if beta is not None else None)
    dx, dgb = norm_act_bwd_raw(x, dy, gamma, beta, plan, None, _f32c(mean), _f32c(rstd), want_param_grads)
    return (dx, dgb[0], dgb[1]) if want_param_grads else (dx, None, None)


_WS_CACHE = {}           # (entry point name, *shape) -> bytes: the workspace queries are host-only functions of the shape


def _workspace(name, *shape):
    key = (name, *shape)
    n = _WS_CACHE.get(key)
    if n is None:
        if len(_WS_CACHE) > 4096:
            _WS_CACHE.clear()
        n = _WS_CACHE[key] = getattr(_lib.load_library(), name)(*shape)
    return n


# ----------------------------------------------------------------------------- spatial gate of AttenModule2 (K7)
def gate7_fwd(x, w, bias=None):
    """x: (N, G, D, H, W) fp32, w: (G, 343) composed weights, bias: one-element tensor or None.  Returns sigmoid(conv + bias),
    (N, 1, D, H, W)."""
    lib = _lib.load_library()
    if not x.is_cuda:
        raise RuntimeError("xlstm_hved_b200 has no CPU path")
    x, w = _f32c(x), _f32c(w)
    N, G, D, H, W = x.shape
    if w.numel() != G * 343:
        raise RuntimeError(f"gate7: weights of {w.numel()} elements for {G} channels")
    gate = torch.empty(N, 1, D, H, W, device=x.device, dtype=torch.float32)
    check(lib.xhved_gate7_fwd(ptr(x), ptr(w), ptr(_f32c(bias)) if bias is not None else None, N, G, D, H, W, ptr(gate), stream()),
          "xhved_gate7_fwd")
    return gate


def gate7_bwd(x, w, gate, dgate, want_dx: bool = True, want_dw: bool = True):
    """Returns (dx, dw, dbias); dx is None unless want_dx, dw / dbias are None unless want_dw."""
    lib = _lib.load_library()
    x, w, gate, dgate = _f32c(x), _f32c(w), _f32c(gate), _f32c(dgate)
    N, G, D, H, W = x.shape
    dx = torch.empty_like(x) if want_dx else None
    dw = db = part = None
    if want_dw:
        dw = torch.empty(G, 343, device=x.device, dtype=torch.float32)
        db = torch.empty(1, device=x.device, dtype=torch.float32)
        part = torch.empty(_workspace("xhved_gate7_workspace", N, G, D, H, W), device=x.device, dtype=torch.uint8)
    check(lib.xhved_gate7_bwd(ptr(x), ptr(w), ptr(gate), ptr(dgate), N, G, D, H, W, ptr(part), ptr(dx), ptr(dw), ptr(db), stream()),
          "xhved_gate7_bwd")
    return dx, dw, db


# ----------------------------------------------------------------------------- depthwise 3x3x3 convolution (K8)
def dwconv3_fwd(x, w, bias=None):
    """x: (N, C, D, H, W) fp32; w: (C, 1, 3, 3, 3) or (C, 27); bias: (C) or None."""
    lib = _lib.load_library()
    if not x.is_cuda:
        raise RuntimeError("xlstm_hved_b200 has no CPU path")
    x, w = _f32c(x), _f32c(w)
    N, C, D, H, W = x.shape
    if w.numel() != C * 27:
        raise RuntimeError(f"dwconv3: weights of {w.numel()} elements for {C} channels")
    y = torch.empty_like(x)
    check(lib.xhved_dwconv3_fwd(ptr(x), ptr(w), ptr(_f32c(bias)) if bias is not None else None, N, C, D, H, W, ptr(y), stream()),
          "xhved_dwconv3_fwd")
    return y


def dwconv3_bwd(x, w, dy, want_dx: bool = True, want_dw: bool = True, want_db: bool = False):
    """Returns (dx, dw, dbias) -- None where not wanted; dw has the shape of w."""
    lib = _lib.load_library()
    x, wc, dy = _f32c(x), _f32c(w), _f32c(dy)
    N, C, D, H, W = x.shape
    dx = torch.empty_like(x) if want_dx else None
    dw = db = part = None
    if want_dw or want_db:
        dw = torch.empty(w.shape, device=x.device, dtype=torch.float32)
        db = torch.empty(C, device=x.device, dtype=torch.float32) if want_db else None
        part = torch.empty(_workspace("xhved_dwconv3_workspace", N, C, D, H, W), device=x.device, dtype=torch.uint8)
    check(lib.xhved_dwconv3_bwd(ptr(x), ptr(wc), ptr(dy), N, C, D, H, W, ptr(part), ptr(dx), ptr(dw), ptr(db), stream()), "xhved_dwconv3_bwd")
    return dx, (dw if want_dw else None), db


# ----------------------------------------------------------------------------- 1x1x1 convolution (K9)
def pwconv_fwd(x, w, bias=None):
    """x: (N, Cin, *spatial) fp32 / fp16 / bf16; w: (Cout, Cin[, 1, 1, 1]); bias: (Cout) or None.  Output in x's type."""
    lib = _lib.load_library()
    if not x.is_cuda:
        raise RuntimeError("xlstm_hved_b200 has no CPU path")
    if x.dtype not in _NORM_DTYPES:
        raise RuntimeError(f"xlstm_hved_b200 pwconv: unsupported dtype {x.dtype}")
    x, w = x.contiguous(), _f32c(w)
    N, Cin = x.shape[:2]
    Cout, vol = w.shape[0], x.numel() // (N * Cin)
    y = torch.empty((N, Cout) + tuple(x.shape[2:]), device=x.device, dtype=x.dtype)
    check(lib.xhved_pwconv_fwd(ptr(x), ptr(w), ptr(_f32c(bias)) if bias is not None else None, N, Cin, Cout, vol, _NORM_DTYPES[x.dtype],
                               ptr(y), stream()), "xhved_pwconv_fwd")
    return y


def pwconv_bwd(x, w, dy, want_dx: bool = True, want_dw: bool = True, want_db: bool = False):
    """Returns (dx in x's type, dw fp32 in w's shape, dbias fp32) -- None where not wanted."""
    lib = _lib.load_library()
    x, wc = x.contiguous(), _f32c(w)
    dy = dy.to(x.dtype).contiguous()
    N, Cin = x.shape[:2]
    Cout, vol = w.shape[0], x.numel() // (N * Cin)
    dx = torch.empty_like(x) if want_dx else None
    dw = torch.empty(w.shape, device=x.device, dtype=torch.float32) if want_dw else None
    db = torch.empty(Cout, device=x.device, dtype=torch.float32) if want_db else None
    part = None
    if (want_dw or want_db) and Cin <= 8 and Cout <= 8:
        part = torch.empty(_workspace("xhved_pwconv_workspace", N, Cin, Cout, vol), device=x.device, dtype=torch.uint8)
    check(lib.xhved_pwconv_bwd(ptr(x), ptr(wc), ptr(dy), N, Cin, Cout, vol, _NORM_DTYPES[x.dtype], ptr(part), ptr(dx), ptr(dw), ptr(db),
                               stream()), "xhved_pwconv_bwd")
    return dx, dw, db


# ----------------------------------------------------------------------------- dense 3x3x3 convolution, few channels (K10)
def conv3_fwd(x, w, bias=None):
    """x: (N, Cin, D, H, W) fp32 / fp16 / bf16; w: (Cout, Cin, 3, 3, 3); bias: (Cout) or None.  Output in x's type."""
    lib = _lib.load_library()
    if not x.is_cuda:
        raise RuntimeError("xlstm_hved_b200 has no CPU path")
    if x.dtype not in _NORM_DTYPES:
        raise RuntimeError(f"xlstm_hved_b200 conv3: unsupported dtype {x.dtype}")
    x, w = x.contiguous(), _f32c(w)
    N, Cin, D, H, W = x.shape
    Cout = w.shape[0]
    if w.numel() != Cout * Cin * 27:
        raise RuntimeError(f"conv3: weight {tuple(w.shape)} does not fit {Cin} input channels")
    y = torch.empty(N, Cout, D, H, W, device=x.device, dtype=x.dtype)
    check(lib.xhved_conv3_fwd(ptr(x), ptr(w), ptr(_f32c(bias)) if bias is not None else None, N, Cin, Cout, D, H, W, _NORM_DTYPES[x.dtype],
                              ptr(y), stream()), "xhved_conv3_fwd")
    return y


def conv3_bwd(x, w, dy, want_dx: bool = True, want_dw: bool = True, want_db: bool = False):
    """Returns (dx in x's type, dw fp32 in w's shape, dbias fp32) -- None where not wanted."""
    lib = _lib.load_library()
    x, wc = x.contiguous(), _f32c(w)
    dy = dy.to(x.dtype).contiguous()
    N, Cin, D, H, W = x.shape
    Cout = w.shape[0]
    dx = torch.empty_like(x) if want_dx else None
    dw = torch.empty(w.shape, device=x.device, dtype=torch.float32) if want_dw else None
    db = torch.empty(Cout, device=x.device, dtype=torch.float32) if want_db else None
    part = None
    if want_dw or want_db:
        part = torch.empty(_workspace("xhved_conv3_workspace", N, Cin, Cout, D, H, W), device=x.device, dtype=torch.uint8)
    check(lib.xhved_conv3_bwd(ptr(x), ptr(wc), ptr(dy), N, Cin, Cout, D, H, W, _NORM_DTYPES[x.dtype], ptr(part), ptr(dx), ptr(dw), ptr(db),
                              stream()), "xhved_conv3_bwd")
    return dx, dw, db


# ----------------------------------------------------------------------------- mLSTM cell
class CellBuffers:
    """Device buffers of one chunkwise cell invocation (tiles + saved-for-backward state)."""

    def __init__(self, BH: int, S: int, dh: int, device):
        self.BH, self.S, self.dh = BH, S, dh
        sizes = _lib.MlstmWorkspace()
        check(_lib.load_library().xhved_mlstm_workspace_query(BH, S, dh, ctypes.byref(sizes)), "xhved_mlstm_workspace_query")
        self.dhp, self.nc = sizes.dhp, sizes.nc
        self.Sp = self.nc * CHUNK
        ne = self.dhp + 16
        bf, f32 = torch.bfloat16, torch.float32
        nt = BH * self.nc
        # the shapes below restate what the library reports (tests/test_host_cpu.py checks the two against each other)
        assert sizes.tile_bytes == nt * CHUNK * self.dhp * 2 and sizes.dstate_bytes == nt * self.dhp * ne * 4
        e = lambda *s, dtype=f32: torch.empty(*s, device=device, dtype=dtype)
        self.q = e(nt, CHUNK * self.dhp, dtype=bf)
        self.k = e(nt, CHUNK * self.dhp, dtype=bf)
        self.v = e(nt, CHUNK * self.dhp, dtype=bf)
        self.h = e(nt, CHUNK * self.dhp, dtype=bf)
        self.ig = e(BH, self.Sp)
        self.fg = e(BH, self.Sp)
        self.m = e(BH, self.Sp)
        self.den = e(BH, self.Sp)
        self.ws_dstate = e(nt, self.dhp * ne)
        self.ws_g = e(nt)
        self.ws_amax = e(nt)
        self.states = e(nt, 2 * self.dhp * ne, dtype=bf)     # hi/lo pair
        self.m_prev = e(nt)


def mlstm_fwd_tiles(buf: CellBuffers, eps: float = 1e-6):
    lib = _lib.load_library()
    check(lib.xhved_mlstm_fwd(ptr(buf.q), ptr(buf.k), ptr(buf.v), ptr(buf.ig), ptr(buf.fg), buf.BH, buf.nc, buf.dh, buf.dhp, eps,
                              ptr(buf.h), ptr(buf.m), ptr(buf.den), ptr(buf.ws_dstate), ptr(buf.ws_g), ptr(buf.ws_amax),
                              ptr(buf.states), ptr(buf.m_prev), stream()), "xhved_mlstm_fwd")


def mlstm_pack_inputs(q, k, v, ig, fg) -> CellBuffers:
    """(B,NH,S,DH) fp32 q,k,v and (B,NH,S,1) gate pre-activations -> tile-native device buffers."""
    lib = _lib.load_library()
    B, NH, S, DH = q.shape
    buf = CellBuffers(B * NH, S, DH, q.device)
    # contiguous copies of strided views (the reference hands over transposed head views) stay referenced until their
    # kernel has been enqueued: a temporary freed earlier could be handed to the next allocation of this function
    for src, dst in ((q, buf.q), (k, buf.k), (v, buf.v)):
        src_c = _f32c(src)
        check(lib.xhved_mlstm_pack(ptr(src_c), buf.BH, S, DH, buf.dhp, ptr(dst), stream()), "xhved_mlstm_pack")
    ig_c, fg_c = _f32c(ig), _f32c(fg)
    check(lib.xhved_mlstm_pack_gates(ptr(ig_c), ptr(fg_c), buf.BH, S, ptr(buf.ig), ptr(buf.fg), stream()), "xhved_mlstm_pack_gates")
    return buf


def mlstm_unpack_h(buf: CellBuffers, B: int, NH: int) -> torch.Tensor:
    lib = _lib.load_library()
    h = torch.empty(B, NH, buf.S, buf.dh, device=buf.h.device, dtype=torch.float32)
    check(lib.xhved_mlstm_unpack(ptr(buf.h), buf.BH, buf.S, buf.dh, buf.dhp, ptr(h), stream()), "xhved_mlstm_unpack")
    return h


def umma_selftest(a_tile: torch.Tensor, b_tile: torch.Tensor, N: int, K: int, a_mn: bool, b_mn: bool) -> torch.Tensor:
    lib = _lib.load_library()
    d = torch.empty(128, N, device=a_tile.device, dtype=torch.float32)
    check(lib.xhved_umma_selftest(ptr(a_tile), ptr(b_tile), N, K, int(a_mn), int(b_mn), ptr(d), stream()), "xhved_umma_selftest")
    return d


def to_tile_native(mat: torch.Tensor) -> torch.Tensor:
    """[R, C] -> bf16 tile-native bytes ((c/8)*R + r, c%8).  Layout helper for tests / diagnostics."""
    R, C = mat.shape
    return mat.to(torch.bfloat16).reshape(R, C // 8, 8).permute(1, 0, 2).contiguous()


class CellGradBuffers:
    def __init__(self, buf: CellBuffers):
        dev = buf.q.device
        nt = buf.BH * buf.nc
        ne = buf.dhp + 16
        e = lambda *s, dtype=torch.float32: torch.empty(*s, device=dev, dtype=dtype)
        # bf16 tiles in the layout of q / k / v (xhved.h: xhved_mlstm_bwd)
        self.dq = torch.empty_like(buf.q)
        self.dk = torch.empty_like(buf.q)
        self.dv = torch.empty_like(buf.q)
        self.dig = e(buf.BH, buf.Sp)
        self.dfg = e(buf.BH, buf.Sp)
        self.rstates = e(nt, 2 * buf.dhp * ne, dtype=torch.bfloat16)
        self.mu_next = e(nt)
        self.ws_dc = e(buf.BH, buf.Sp)


def mlstm_bwd_tiles(buf: CellBuffers, dh_tiles: torch.Tensor, eps: float = 1e-6) -> CellGradBuffers:
    """Backward of mlstm_fwd_tiles; dh_tiles: bf16 tile-native gradient of h."""
    lib = _lib.load_library()
    gb = CellGradBuffers(buf)
    check(lib.xhved_mlstm_bwd(ptr(buf.q), ptr(buf.k), ptr(buf.v), ptr(buf.ig), ptr(buf.fg), ptr(buf.h), ptr(dh_tiles), ptr(buf.m),
                              ptr(buf.den), ptr(buf.states), ptr(buf.m_prev), buf.BH, buf.nc, buf.dh, buf.dhp, eps, ptr(gb.dq),
                              ptr(gb.dk), ptr(gb.dv), ptr(gb.dig), ptr(gb.dfg), ptr(buf.ws_dstate), ptr(buf.ws_g), ptr(buf.ws_amax),
                              ptr(gb.rstates), ptr(gb.mu_next), ptr(gb.ws_dc), stream()), "xhved_mlstm_bwd")
    return gb


def _unpack_grad(tiles, BH, S, dh, dhp, shape):
    """bf16 gradient tiles -> fp32 (B, NH, S, DH) (the stand-alone cell entry point returns what autograd expects)."""
    lib = _lib.load_library()
    dst = torch.empty(shape, device=tiles.device, dtype=torch.float32)
    check(lib.xhved_mlstm_unpack(ptr(tiles), BH, S, dh, dhp, ptr(dst), stream()), "xhved_mlstm_unpack")
    return dst


class MLSTMCellFunction(torch.autograd.Function):
    """parallel_stabilized_simple (vision_lstm.py:48-130) on the chunkwise tcgen05 kernels."""

    @staticmethod
    @_lib.on_device
    def forward(ctx, q, k, v, ig, fg, eps):
        B, NH, S, DH = q.shape
        buf = mlstm_pack_inputs(q, k, v, ig, fg)
        mlstm_fwd_tiles(buf, eps)
        ctx.buf, ctx.eps, ctx.shape = buf, eps, (B, NH, S, DH)
        return mlstm_unpack_h(buf, B, NH)

    @staticmethod
    @_lib.on_device
    def backward(ctx, dh):
        lib = _lib.load_library()
        buf = ctx.buf
        B, NH, S, DH = ctx.shape
        dh_tiles = torch.empty_like(buf.h)
        dh_c = _f32c(dh)
        check(lib.xhved_mlstm_pack(ptr(dh_c), buf.BH, S, DH, buf.dhp, ptr(dh_tiles), stream()), "xhved_mlstm_pack")
        gb = mlstm_bwd_tiles(buf, dh_tiles, ctx.eps)
        dq = _unpack_grad(gb.dq, buf.BH, S, DH, buf.dhp, (B, NH, S, DH))
        dk = _unpack_grad(gb.dk, buf.BH, S, DH, buf.dhp, (B, NH, S, DH))
        dv = _unpack_grad(gb.dv, buf.BH, S, DH, buf.dhp, (B, NH, S, DH))
        dig = gb.dig.view(B, NH, buf.Sp)[:, :, :S].unsqueeze(-1).contiguous()
        dfg = gb.dfg.view(B, NH, buf.Sp)[:, :, :S].unsqueeze(-1).contiguous()
        return dq, dk, dv, dig, dfg, None


def parallel_stabilized_simple(queries, keys, values, igate_preact, fgate_preact, lower_triangular_matrix=None,
                               stabilize_rowwise: bool = True, eps: float = 1e-6):
    """Drop-in for vision_lstm.py:48-57 (same signature).  The causal mask argument is accepted and ignored
    (the kernels are causal by construction); only the row-wise stabiliser the reference uses is built."""
    if not stabilize_rowwise:
        raise NotImplementedError("stabilize_rowwise=False is never used by XLSTM-HVED and is not built")
    if not queries.is_cuda:
        raise RuntimeError("xlstm_hved_b200 has no CPU path")
    out_dtype = queries.dtype
    h = MLSTMCellFunction.apply(queries.float(), keys.float(), values.float(), igate_preact.float(), fgate_preact.float(), eps)
    return h.to(out_dtype)


# ----------------------------------------------------------------------------- ViL block (K2 + K1 + K3)
VIL_PARAM_KEYS = ["norm.weight", "layer.proj_up.weight", "layer.conv1d.conv.weight", "layer.conv1d.conv.bias",
                  "layer.q_proj.weight", "layer.k_proj.weight", "layer.v_proj.weight",
                  "layer.mlstm_cell.igate.weight", "layer.mlstm_cell.igate.bias",
                  "layer.mlstm_cell.fgate.weight", "layer.mlstm_cell.fgate.bias",
                  "layer.mlstm_cell.outnorm.weight", "layer.learnable_skip", "layer.proj_down.weight"]


def _param_struct(tensors, cls):
    st = cls()
    for name, t in zip(_lib._PARAM_FIELDS, tensors):
        setattr(st, name, t.data_ptr() if t is not None else None)
    return st


def _token_strides(x_tok: torch.Tensor):
    """x_tok: a (B,S,C) view (any strides, e.g. the transpose of an NCDHW feature)."""
    return x_tok.stride(0), x_tok.stride(1), x_tok.stride(2)


def _dense_view(t: torch.Tensor) -> torch.Tensor:
    """The kernels address (B,S,C) views through their strides and allocate outputs with the same strides, which is only
    sound when distinct indices map to distinct addresses.  Autograd hands over EXPANDED gradients (``y.sum().backward()``
    arrives with strides (0,0,0)) and callers may pass overlapping views: those are materialised here; non-overlapping views
    of any memory format (e.g. the transpose of an NCDHW feature, slices) pass through untouched."""
    if t.numel() <= 1 or t.is_contiguous():
        return t
    extent = 1                                       # elements spanned by the dimensions with smaller strides
    for st, sz in sorted((st, sz) for sz, st in zip(t.shape, t.stride()) if sz > 1):
        if st < extent:                              # stride 0 (expanded) or overlapping windows
            return t.contiguous()
        extent = st * (sz - 1) + extent
    return t


class VilSaved:
    """What one forward of the fused ViL block leaves for its backward: ONE device blob (tiles, gates, stabiliser, carried
    states, act / z / xm; carved up inside xhved_vil_block_fwd / _bwd) and the sizes the library reported for this shape."""
    __slots__ = ("blob", "scratch_bytes", "n_param_grads", "stride")

    def __init__(self, blob, scratch_bytes, n_param_grads, stride):
        self.blob, self.scratch_bytes, self.n_param_grads, self.stride = blob, scratch_bytes, n_param_grads, stride


_BLOCK_SIZES = {}        # (B, S, C, replicas) -> (saved_bytes, scratch_bytes, n_param_grads, stride): host-only queries, cached


def _block_sizes(B, S, C):
    key = (B, S, C, GRAD_REPLICAS)
    if key not in _BLOCK_SIZES:
        lib = _lib.load_library()
        sv, sc, npg = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
        check(lib.xhved_vil_block_workspace(B, S, C, GRAD_REPLICAS, ctypes.byref(sv), ctypes.byref(sc), ctypes.byref(npg)),
              "xhved_vil_block_workspace")
        sizes = _lib.VilWorkspaceSizes()
        check(lib.xhved_vil_workspace_query(B, S, C, ctypes.byref(sizes)), "xhved_vil_workspace_query")
        _BLOCK_SIZES[key] = (sv.value, sc.value, npg.value, sizes.grad_replica_stride)
    return _BLOCK_SIZES[key]


def _shape_struct(x_tok, y_tok, reverse):
    B, S, C = x_tok.shape
    sh = _lib.VilShape()
    sh.B, sh.S, sh.C, sh.NH, sh.QB, sh.reverse = B, S, C, 4, 4, int(bool(reverse))
    sh.x_stride_b, sh.x_stride_n, sh.x_stride_c = _token_strides(x_tok)
    sh.y_stride_b, sh.y_stride_n, sh.y_stride_c = _token_strides(y_tok)
    return sh


def vil_block_fwd(x_tok: torch.Tensor, params, reverse: bool, eps: float = 1e-6):
    """x_tok: (B,S,C) fp32 view; params: the 14 tensors in VIL_PARAM_KEYS order.  Returns (y_tok, saved).
    ONE call into the library (xhved_vil_block_fwd: K2 -> cell -> K3) and two allocations (y, the saved blob)."""
    lib = _lib.load_library()
    x_tok = _dense_view(x_tok)
    B, S, C = x_tok.shape
    saved_bytes, scratch_bytes, npg, stride = _block_sizes(B, S, C)
    blob = torch.empty(saved_bytes, device=x_tok.device, dtype=torch.uint8)
    # output takes the memory format of the input (NCDHW-backed token view stays NCDHW-backed)
    y = torch.empty_strided(x_tok.shape, x_tok.stride(), device=x_tok.device, dtype=torch.float32)
    ps = _param_struct(params, _lib.VilParams)
    sh = _shape_struct(x_tok, y, reverse)
    check(lib.xhved_vil_block_fwd(ptr(x_tok), ctypes.byref(ps), ctypes.byref(sh), eps, ptr(blob), ptr(y), stream()), "xhved_vil_block_fwd")
    return y, VilSaved(blob, scratch_bytes, npg, stride)


def vil_block_bwd(x_tok, dy_tok, params, reverse, ws: VilSaved, eps: float = 1e-6):
    """Backward of vil_block_fwd.  Returns (dx_tok, [14 parameter gradients in VIL_PARAM_KEYS order]).
    ONE call into the library (xhved_vil_block_bwd) and three allocations (scratch blob, dx, the flat parameter gradients)."""
    lib = _lib.load_library()
    dev = x_tok.device
    if dy_tok.dtype != torch.float32:
        dy_tok = dy_tok.float()
    dy_tok = _dense_view(dy_tok)
    assert sum(p.numel() for p in params) == ws.n_param_grads
    scratch = torch.empty(ws.scratch_bytes, device=dev, dtype=torch.uint8)
    out = torch.empty(ws.n_param_grads, device=dev, dtype=torch.float32)
    dx = torch.empty_strided(dy_tok.shape, dy_tok.stride(), device=dev, dtype=torch.float32)
    ps = _param_struct(params, _lib.VilParams)
    sh = _shape_struct(x_tok, dy_tok, reverse)          # y_* strides describe dy and dx
    # parameter gradients are accumulated with global atomics into GRAD_REPLICAS zero-filled copies (CTA i -> copy i % R) to
    # spread the traffic over L2 slices; the call sums the copies at the end
    sh.grad_replicas, sh.grad_replica_stride = GRAD_REPLICAS, ws.stride
    check(lib.xhved_vil_block_bwd(ptr(x_tok), ptr(dy_tok), ctypes.byref(ps), ctypes.byref(sh), eps, ptr(ws.blob), ptr(scratch), ptr(dx),
                                  ptr(out), stream()), "xhved_vil_block_bwd")
    return dx, _split_param_grads(out, params)


def _split_param_grads(flat, params):
    grads, off = [], 0
    for p in params:
        grads.append(flat[off:off + p.numel()].view(p.shape))
        off += p.numel()
    return grads


class VilBlockFunction(torch.autograd.Function):
    """ViLBlock.forward (vision_lstm.py:494-502) as three fused launches + the chunkwise cell, with backward."""

    @staticmethod
    @_lib.on_device
    def forward(ctx, x_tok, reverse, eps, *params):
        params = [p.detach() if p.dtype == torch.float32 and p.is_contiguous() else p.detach().float().contiguous() for p in params]
        x_tok = _dense_view(x_tok)
        y, ws = vil_block_fwd(x_tok, params, reverse, eps)
        ctx.reverse, ctx.eps, ctx.ws = reverse, eps, ws
        ctx.save_for_backward(x_tok, *params)
        return y

    @staticmethod
    @_lib.on_device
    def backward(ctx, dy):
        x_tok, *params = ctx.saved_tensors
        dx, grads = vil_block_bwd(x_tok, dy, params, ctx.reverse, ctx.ws, ctx.eps)
        return (dx, None, None, *grads)


# ----------------------------------------------------------------------------- per-block CUDA graphs (small batches)
# At the reference's real batch (one volume, train.py:50) a block is ~14 launches of 5-30 us per direction: the kernels take
# ~0.12 ms, the host side -- Python, ctypes structs, allocations, the launches themselves -- 0.5 ms.  For small blocks the
# forward and the backward of a block are therefore captured ONCE into CUDA graphs over static buffers (x / dy are copied in,
# y / dx / the flat parameter gradients cloned out: three small copies instead of ~28 launches) and replayed afterwards.
#   * a "slot" = the static buffers + the two graphs of one (device, shape, strides, direction, parameter storage); a slot is
#     busy from its forward until its backward (or until autograd frees the node), so the two forwards of a training step
#     (train.py:222-225) use two slots; at most _SLOT_CAP slots per key, then the plain path runs;
#   * nothing is graphed while the caller itself captures a stream (bench.py replays the whole step), for parameters that are
#     not fp32 leaf tensors (nn.DataParallel replicas are re-broadcast every step), or above _GRAPH_MAX_TOKENS tokens (the
#     kernels then outlast the host side anyway).
# XHVED_BLOCK_GRAPHS=0 disables, =1 forces graphs at every size; default "auto".
_GRAPH_MODE = os.environ.get("XHVED_BLOCK_GRAPHS", "auto")
_GRAPH_MAX_TOKENS = 65536
_SLOT_CAP = 4
_MAX_KEYS = 16
_GRAPH_LOCK = threading.Lock()
_GRAPH_POOLS = collections.OrderedDict()      # key -> [slots], least recently used first


def set_block_graphs(mode) -> None:
    """"auto" (default), True / "1" (always), False / "0" (never): per-block CUDA graphs, see above."""
    global _GRAPH_MODE
    _GRAPH_MODE = {True: "1", False: "0"}.get(mode, str(mode))


class _BlockGraphSlot:
    def __init__(self, x_like, params, reverse, eps):
        dev = x_like.device
        self.params, self.reverse, self.eps = params, reverse, eps
        self.busy, self.gen, self.bwd = False, 0, None
        self.x_in = torch.empty_strided(x_like.shape, x_like.stride(), device=dev, dtype=torch.float32)
        self.x_in.copy_(x_like)
        self._warm(lambda: vil_block_fwd(self.x_in, params, reverse, eps))
        # captures are serialised by _GRAPH_LOCK (the caller holds it) and run on a stream of their own: torch.cuda.graph's default
        # capture stream is shared by all threads
        self.cap_stream = torch.cuda.Stream()
        self.fwd = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.fwd, stream=self.cap_stream, capture_error_mode="thread_local"):
            self.y, self.ws = vil_block_fwd(self.x_in, params, reverse, eps)

    @staticmethod
    def _warm(fn):
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            fn()
        cur.wait_stream(side)

    def capture_bwd(self, dy):
        with _GRAPH_LOCK:
            self.dy_in = torch.empty_strided(self.y.shape, self.y.stride(), device=self.y.device, dtype=torch.float32)
            self.dy_in.copy_(dy)
            self._warm(lambda: self._bwd_raw())
            bwd = torch.cuda.CUDAGraph()
            with torch.cuda.graph(bwd, stream=self.cap_stream, capture_error_mode="thread_local"):
                self.dx, self.flat = self._bwd_raw()
            self.bwd = bwd

    def _bwd_raw(self):
        dx, grads = vil_block_bwd(self.x_in, self.dy_in, self.params, self.reverse, self.ws, self.eps)
        return dx, grads[0]._base if grads[0]._base is not None else grads[0]


class _SlotRelease:
    """Frees the slot when autograd drops the node (a forward whose backward never runs)."""

    def __init__(self, slot, gen):
        self.slot, self.gen = slot, gen

    def __del__(self):
        if self.slot.gen == self.gen:
            self.slot.busy = False


def _acquire_slot(x_tok, params, reverse, eps):
    if _GRAPH_MODE == "0" or torch.cuda.is_current_stream_capturing():
        return None
    if _GRAPH_MODE != "1" and x_tok.shape[0] * x_tok.shape[1] > _GRAPH_MAX_TOKENS:
        return None
    if any(p.dtype != torch.float32 or not p.is_contiguous() or not p.is_leaf for p in params):
        return None
    key = (x_tok.device.index, tuple(x_tok.shape), tuple(x_tok.stride()), bool(reverse), float(eps), tuple(p.data_ptr() for p in params))
    with _GRAPH_LOCK:
        pool = _GRAPH_POOLS.get(key)
        if pool is None:
            while len(_GRAPH_POOLS) >= _MAX_KEYS:                  # forget the least recently used shape / parameter set
                _GRAPH_POOLS.popitem(last=False)
            pool = _GRAPH_POOLS[key] = []
        else:
            _GRAPH_POOLS.move_to_end(key)
        slot = next((s for s in pool if not s.busy), None)
        if slot is None:
            if len(pool) >= _SLOT_CAP:
                return None
            slot = _BlockGraphSlot(x_tok, [p.detach() for p in params], bool(reverse), eps)
            pool.append(slot)
        slot.busy = True
        slot.gen += 1
        return slot


class _GraphedVilBlockFunction(torch.autograd.Function):
    """VilBlockFunction replayed from the CUDA graphs of a slot."""

    @staticmethod
    @_lib.on_device
    def forward(ctx, x_tok, slot, will_backward, *params):
        slot.x_in.copy_(x_tok)
        slot.fwd.replay()
        y = slot.y.clone()
        ctx.slot, ctx.gen = slot, slot.gen
        if will_backward:
            ctx.release = _SlotRelease(slot, slot.gen)
        else:
            slot.busy = False               # no graph is recorded (no_grad / nothing requires grad): the saved state is not needed
        return y

    @staticmethod
    @_lib.on_device
    def backward(ctx, dy):
        slot = ctx.slot
        if slot.gen != ctx.gen:
            raise RuntimeError("xlstm_hved_b200: the saved state of this graphed ViL block has been reused by a later forward "
                               "(backward twice through the same node?); set XHVED_BLOCK_GRAPHS=0 for that pattern")
        if slot.bwd is None:
            slot.capture_bwd(dy)
        else:
            slot.dy_in.copy_(dy)
        slot.bwd.replay()
        dx, flat = slot.dx.clone(), slot.flat.clone()
        slot.busy = False
        return (dx, None, None, *_split_param_grads(flat, slot.params))


def vil_block(x_tok: torch.Tensor, params, reverse: bool = False, eps: float = 1e-6) -> torch.Tensor:
    """x_tok: (B,S,C) view of fp32 CUDA memory (any strides); params in VIL_PARAM_KEYS order."""
    if not x_tok.is_cuda:
        raise RuntimeError("xlstm_hved_b200 has no CPU path")
    if x_tok.dtype != torch.float32:
        x_tok = x_tok.float()
    x_tok = _dense_view(x_tok)
    with torch.cuda.device(x_tok.device):
        slot = _acquire_slot(x_tok, params, reverse, eps)
    if slot is not None:
        will_backward = torch.is_grad_enabled() and (x_tok.requires_grad or any(p.requires_grad for p in params))
        return _GraphedVilBlockFunction.apply(x_tok, slot, will_backward, *params)
    return VilBlockFunction.apply(x_tok, bool(reverse), eps, *params)


# ----------------------------------------------------------------------------- ViL blocks wider than the fused kernels
WIDE_DIMS = (128, 256)        # f_maps 16 / 32: E = 256 / 512, head dim 64 / 128 (csrc/vil_wide.cu)


class _WideCoreFunction(torch.autograd.Function):
    """Everything of a wide ViL block between proj_up and proj_down (vision_lstm.py:428-440): conv, SiLU, block-diagonal q / k / v,
    gates, the cell, per-head norm, skip and z gate -- four fused glue kernels (csrc/vil_wide.cu) around the tcgen05 cell kernels.
    up: (B, S, 2E) fp32 contiguous in natural token order; returns hg (B, S, E)."""

    @staticmethod
    @_lib.on_device
    def forward(ctx, up, reverse, eps, conv_w, conv_b, qw, kw, vw, igw, igb, fgw, fgb, ow, sk):
        lib = _lib.load_library()
        B, S, E2 = up.shape
        E = E2 // 2
        prm = [_f32c(t.detach()) for t in (conv_w, conv_b, qw, kw, vw, igw, igb, fgw, fgb, ow, sk)]
        buf = CellBuffers(4 * B, S, E // 4, up.device)
        act = torch.empty(B, S, E, device=up.device, dtype=torch.float32)
        check(lib.xhved_vil_wide_pre_fwd(ptr(up), *[ptr(t) for t in prm[:9]], B, S, E, int(reverse), ptr(buf.q), ptr(buf.k), ptr(buf.v),
                                         ptr(buf.ig), ptr(buf.fg), ptr(act), stream()), "xhved_vil_wide_pre_fwd")
        mlstm_fwd_tiles(buf, eps)
        hg = torch.empty(B, S, E, device=up.device, dtype=torch.float32)
        check(lib.xhved_vil_wide_post_fwd(ptr(buf.h), ptr(act), ptr(up), ptr(prm[9]), ptr(prm[10]), B, S, E, int(reverse), ptr(hg), stream()),
              "xhved_vil_wide_post_fwd")
        ctx.buf, ctx.reverse, ctx.eps = buf, bool(reverse), eps
        ctx.save_for_backward(up, act, *prm)
        return hg

    @staticmethod
    @_lib.on_device
    def backward(ctx, dhg):
        lib = _lib.load_library()
        up, act, *prm = ctx.saved_tensors
        conv_w, conv_b, qw, kw, vw, igw, igb, fgw, fgb, ow, sk = prm
        buf = ctx.buf
        B, S, E2 = up.shape
        E = E2 // 2
        dev = up.device
        dhg = _f32c(dhg)
        z = lambda *shape: torch.zeros(*shape, device=dev, dtype=torch.float32)
        d_up, d_act, dh_tiles = torch.empty_like(up), torch.empty_like(act), torch.empty_like(buf.h)
        g_ow, g_sk = z(E), z(E)
        check(lib.xhved_vil_wide_post_bwd(ptr(dhg), ptr(buf.h), ptr(act), ptr(up), ptr(ow), ptr(sk), B, S, E, int(ctx.reverse), ptr(dh_tiles),
                                          ptr(d_act), ptr(d_up), ptr(g_ow), ptr(g_sk), stream()), "xhved_vil_wide_post_bwd")
        gb = mlstm_bwd_tiles(buf, dh_tiles, ctx.eps)
        dg_rm = torch.empty(B, S, 8, device=dev, dtype=torch.float32)
        g_cw, g_cb, g_qw, g_kw, g_vw = z(*conv_w.shape), z(E), z(*qw.shape), z(*kw.shape), z(*vw.shape)
        check(lib.xhved_vil_wide_pre_bwd(ptr(up), ptr(conv_w), ptr(conv_b), ptr(qw), ptr(kw), ptr(vw), ptr(igw), ptr(fgw), B, S, E,
                                         int(ctx.reverse), ptr(gb.dq), ptr(gb.dk), ptr(gb.dv), ptr(gb.dig), ptr(gb.dfg), ptr(d_act), ptr(d_up),
                                         ptr(dg_rm), ptr(g_cw), ptr(g_cb), ptr(g_qw), ptr(g_kw), ptr(g_vw), stream()), "xhved_vil_wide_pre_bwd")
        # gate Linear(3E -> 4) x 2: d bias = sum over tokens; d weight THROUGH the 4x4 projections (csrc/vil_pre.cu):
        # [dig|dfg]^T q = ([dig|dfg]^T act) Wq^T (k alike; v with x_mlstm and Wv): two plain (8 x T)(T x E) GEMMs
        dg2 = dg_rm.view(-1, 8)
        r_a = (dg2.t() @ act.view(-1, E)).view(8, E // 4, 4)
        r_x = (dg2.t() @ up.view(-1, E2)[:, :E]).view(8, E // 4, 4)
        g_w = torch.cat([torch.einsum("gbd,bod->gbo", r_a, qw).reshape(8, E), torch.einsum("gbd,bod->gbo", r_a, kw).reshape(8, E),
                         torch.einsum("gbd,bod->gbo", r_x, vw).reshape(8, E)], dim=1)
        g_b = dg2.sum(0)
        return (d_up, None, None, g_cw, g_cb, g_qw, g_kw, g_vw, g_w[:4].contiguous(), g_b[:4].contiguous(), g_w[4:].contiguous(),
                g_b[4:].contiguous(), g_ow, g_sk)


def _split_hilo_cat(x2d: torch.Tensor, b_side: bool = False) -> torch.Tensor:
    """fp32 (rows, cols) -> bf16 (rows, 3 cols) = [hi | lo | hi] (or [hi | hi | lo] for the b side), see xhved_split_hilo_cat."""
    lib = _lib.load_library()
    x2d = _f32c(x2d)
    rows, cols = x2d.shape
    out = torch.empty(rows, 3 * cols, device=x2d.device, dtype=torch.bfloat16)
    check(lib.xhved_split_hilo_cat(ptr(x2d), rows, cols, int(b_side), ptr(out), stream()), "xhved_split_hilo_cat")
    return out


class _HiLoLinearFunction(torch.autograd.Function):
    """y = x W^T (nn.Linear without bias) as bf16 tensor-core GEMMs with fp32-class accuracy: the operands are split into bf16
    hi + lo and ONE library GEMM over the tripled contraction dimension gives hi*hi + lo*hi + hi*lo with fp32 accumulation
    (the 3-product of csrc/umma.cuh, ~16 mantissa bits).  x: (T, K) fp32, W: (N, K) fp32 -> (T, N) fp32."""

    @staticmethod
    @_lib.on_device
    def forward(ctx, x, w):
        xc = _split_hilo_cat(x)                                   # (T, 3K)  [hi | lo | hi]
        wc = _split_hilo_cat(w, b_side=True)                      # (N, 3K)  [hi | hi | lo]
        ctx.save_for_backward(xc, w)
        return torch.mm(xc, wc.t(), out_dtype=torch.float32)

    @staticmethod
    @_lib.on_device
    def backward(ctx, dy):
        xc, w = ctx.saved_tensors
        n, k = w.shape
        dyc = _split_hilo_cat(dy)                                 # (T, 3N)  [hi | lo | hi]
        # dx = dy W: contraction over N, the b side stacked along it ([hi ; hi ; lo] rows)
        wr = _split_hilo_cat(w.t().contiguous(), b_side=True).t().contiguous()        # (3N, K)
        dx = torch.mm(dyc, wr, out_dtype=torch.float32)
        # dW = dy^T x: contraction over the tokens, the three products separately (their outputs are small)
        dyh, dyl, xh, xl = dyc[:, :n], dyc[:, n:2 * n], xc[:, :k], xc[:, k:2 * k]
        dw = (torch.mm(dyh.t(), xh, out_dtype=torch.float32) + torch.mm(dyl.t(), xh, out_dtype=torch.float32) +
              torch.mm(dyh.t(), xl, out_dtype=torch.float32))
        return dx, dw


def vil_block_wide(x_tok: torch.Tensor, params, reverse: bool = False, eps: float = 1e-6) -> torch.Tensor:
    """ViLBlock.forward (vision_lstm.py:499-502) at dim 128 / 256: LayerNorm + proj_up and proj_down + residual are plain library
    GEMMs / torch ops (per token, so the direction flip does not touch them), everything between them is _WideCoreFunction."""
    if not x_tok.is_cuda:
        raise RuntimeError("xlstm_hved_b200 has no CPU path")
    if x_tok.shape[-1] not in WIDE_DIMS:
        raise RuntimeError(f"vil_block_wide covers dims {WIDE_DIMS}")
    if x_tok.dtype != torch.float32:
        x_tok = x_tok.float()
    norm_w, up_w, conv_w, conv_b, qw, kw, vw, igw, igb, fgw, fgb, ow, sk, down_w = params
    C = x_tok.shape[-1]
    xn = torch.nn.functional.layer_norm(x_tok, (C,), weight=1.0 + norm_w, bias=None, eps=1e-5)
    Bq, Sq = x_tok.shape[:2]
    up = _HiLoLinearFunction.apply(xn.reshape(-1, C), up_w).view(Bq, Sq, -1)
    hg = _WideCoreFunction.apply(up, bool(reverse), eps, conv_w, conv_b, qw, kw, vw, igw, igb, fgw, fgb, ow, sk)
    return x_tok + _HiLoLinearFunction.apply(hg.reshape(-1, hg.shape[-1]), down_w).view(Bq, Sq, C)
