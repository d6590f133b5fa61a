#!/usr/bin/env python
"""Developer tool: per-phase clock64 totals of mlstm_chunk_grad_ws (library built with XHVED_NVCC_EXTRA=-DXHVED_TRACE)."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from xlstm_hved_b200 import _lib, ops  # noqa: E402

B, NH, S, DH = [int(a) for a in sys.argv[1:5]] if len(sys.argv) >= 5 else (32, 4, 4096, 16)
lib = _lib.load_library()
g = torch.Generator(device="cuda").manual_seed(0)
q, k, v = [0.3 * torch.randn(B, NH, S, DH, device="cuda", generator=g) for _ in range(3)]
ig = torch.randn(B, NH, S, 1, device="cuda", generator=g)
fg = 2.0 + torch.randn(B, NH, S, 1, device="cuda", generator=g)
buf = ops.mlstm_pack_inputs(q, k, v, ig, fg)
dh_tiles = torch.randn_like(buf.h.float()).to(torch.bfloat16)
ops.mlstm_fwd_tiles(buf)
for _ in range(2):
    ops.mlstm_bwd_tiles(buf, dh_tiles)
fn = lib.xhved_debug_trace
fn.argtypes = [ctypes.POINTER(ctypes.c_ulonglong), ctypes.c_int]
out = (ctypes.c_ulonglong * 36)()
fn(out, 1)
iters = 5
for _ in range(iters):
    ops.mlstm_bwd_tiles(buf, dh_tiles)
fn(out, 0)
ntiles = B * NH * ((S + 127) // 128)
names = {0: ["loop", "wait dP", "convert dS", "scans(next)", "wait full+ma", "epilogue dQ dK", "rcumsum"],
         1: ["loop", "wait prep", "wait S^T", "convert P^T", "G(next)", "wait mb", "epilogue dV"],
         2: ["loop", "wait mb/tfree(prev)", "wait full", "issue S^T", "wait tfree(prev)", "issue load", "wait prep", "issue dP", "wait c0",
             "issue dQ dK", "wait c1", "issue dV"]}
for who in range(3):
    row = [out[who * 12 + i] / (iters * ntiles) for i in range(len(names[who]))]
    print(("group0 (rows)", "group1 (cols)", "control")[who], " | ".join(f"{n}={c:.0f}" for n, c in zip(names[who], row)), "| total", f"{sum(row):.0f}", "cycles/tile")
