#!/usr/bin/env python
"""Developer tool: a few forward + backward passes of the cell alone (for ncu captures).  argv: B NH S DH"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from xlstm_hved_b200 import ops  # noqa: E402

B, NH, S, DH = [int(a) for a in sys.argv[1:5]] if len(sys.argv) >= 5 else (32, 4, 4096, 16)
g = torch.Generator(device="cuda").manual_seed(0)
q, k, v = [0.3 * torch.randn(B, NH, S, DH, device="cuda", generator=g) for _ in range(3)]
ig = torch.randn(B, NH, S, 1, device="cuda", generator=g)
fg = 2.0 + torch.randn(B, NH, S, 1, device="cuda", generator=g)
buf = ops.mlstm_pack_inputs(q, k, v, ig, fg)
dh_tiles = torch.randn_like(buf.h.float()).to(torch.bfloat16)
for _ in range(3):
    ops.mlstm_fwd_tiles(buf)
    ops.mlstm_bwd_tiles(buf, dh_tiles)
torch.cuda.synchronize()
