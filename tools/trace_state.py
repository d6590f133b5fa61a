#!/usr/bin/env python
"""Developer tool: clock64 phase trace of CTA 0 of the persistent chunk_state kernel (library built with -DXHVED_TRACE)."""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from xlstm_hved_b200 import _lib, ops
B, NH, S, DH = 32, 4, 4096, 16
g = torch.Generator(device="cuda").manual_seed(0)
q, k, v = [0.3 * torch.randn(B, NH, S, DH, device="cuda", generator=g) for _ in range(3)]
ig = torch.randn(B, NH, S, 1, device="cuda", generator=g)
fg = 2.0 + torch.randn(B, NH, S, 1, device="cuda", generator=g)
buf = ops.mlstm_pack_inputs(q, k, v, ig, fg)
for _ in range(3):
    ops.mlstm_fwd_tiles(buf)
torch.cuda.synchronize()
lib = _lib.load_library()
out = (ctypes.c_longlong * 256)()
lib.xhved_debug_trace.argtypes = [ctypes.c_void_p]
assert lib.xhved_debug_trace(out) == 0
t = [[out[r * 32 + i] for i in range(32)] for r in range(8)]
t0 = t[7][0]
names = ["prod_issue", "mma_ready", "mma_issued", "epi_mdone", "epi_done", "cons_full", "cons_arrive"]
print("kernel body cycles:", t[7][1] - t0)
for i in range(8):
    print(i, " ".join(f"{names[r]}={t[r][i] - t0:6d}" for r in range(7)))
