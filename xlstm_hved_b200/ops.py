"""Host-side ops over the C ABI (include/xhved.h): allocation with torch, launch through ctypes.

Nothing here computes on the CPU or with torch math: every op launches the
sm_100a kernels of libxhved.so on the current CUDA stream.
"""
from __future__ import annotations

import ctypes
from ctypes import c_float, c_uint32

import torch

from . import _lib
from ._lib import check, ptr, stream

CHUNK = 128

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


def poe_fwd(mu5: torch.Tensor, logvar5: torch.Tensor, subsets, drop=None, noise=None, want_kld=False, eps: float = 1e-8):
    """All requested subsets in one launch.  mu5/logvar5: (5, ...) fp32 contiguous (prior first).
    Returns (pd_mu, pd_logvar, z or None, kld_sums or None) each of shape (len(subsets), ...)."""
    lib = _lib.load_library()
    assert mu5.shape == logvar5.shape and mu5.shape[0] == 5
    mu5, logvar5 = _f32c(mu5), _f32c(logvar5)
    n = mu5[0].numel()
    ns = len(subsets)
    out_mu = torch.empty((ns, *mu5.shape[1:]), device=mu5.device, dtype=torch.float32)
    out_lv = torch.empty_like(out_mu)
    z = torch.empty_like(out_mu) if noise is not None else None
    kld = torch.zeros(ns, device=mu5.device, dtype=torch.float32) if want_kld else None
    per_sample = 0
    if drop is not None:
        drop = drop.to(device=mu5.device, dtype=torch.uint8).contiguous()
        per_sample = n // drop.shape[0]
    if noise is not None:
        noise = _f32c(noise)
        assert noise.numel() == ns * n
    check(lib.xhved_poe_fwd(ptr(mu5), ptr(logvar5), n, n, _masks(subsets), ns, ptr(drop), per_sample, eps, ptr(out_mu),
                            ptr(out_lv), ptr(noise), ptr(z), ptr(kld), stream()), "xhved_poe_fwd")
    return out_mu, out_lv, z, kld


def poe_bwd(mu5, logvar5, subsets, g_mu=None, g_logvar=None, noise=None, g_z=None, kld_scale=None, drop=None, eps: float = 1e-8):
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
                            ptr(noise), ptr(g_z), ks, ptr(d_mu), ptr(d_lv), stream()), "xhved_poe_bwd")
    return d_mu, d_lv


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


# ----------------------------------------------------------------------------- mLSTM cell
class CellBuffers:
    """Device buffers of one chunkwise cell invocation (tiles + saved-for-backward state)."""

    def __init__(self, BH: int, S: int, dh: int, device):
        self.BH, self.S, self.dh = BH, S, dh
        self.dhp = padded_dh(dh)
        self.nc = (S + CHUNK - 1) // CHUNK
        self.Sp = self.nc * CHUNK
        ne = self.dhp + 16
        bf, f32 = torch.bfloat16, torch.float32
        nt = BH * self.nc
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
        self.states = e(nt, self.dhp * ne, dtype=bf)
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
    for src, dst in ((q, buf.q), (k, buf.k), (v, buf.v)):
        check(lib.xhved_mlstm_pack(ptr(_f32c(src)), buf.BH, S, DH, buf.dhp, ptr(dst), stream()), "xhved_mlstm_pack")
    check(lib.xhved_mlstm_pack_gates(ptr(_f32c(ig)), ptr(_f32c(fg)), buf.BH, S, ptr(buf.ig), ptr(buf.fg), stream()),
          "xhved_mlstm_pack_gates")
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
