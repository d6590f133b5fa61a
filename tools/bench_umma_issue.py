#!/usr/bin/env python
"""Diagnostic: cycles per tcgen05.mma (M = 128, K = 16, bf16) by N, accumulator count and operand placement."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from xlstm_hved_b200 import _lib  # noqa: E402

lib = _lib.load_library()
out = torch.zeros(2, dtype=torch.int64, device="cuda")
rows = []
for N in (16, 32, 64, 128):
    for n_acc in (1, 2, 4):
        if n_acc * max(N, 32) > 256:
            continue
        for a_tmem in (0, 1):
            for mn in (0, 2, 3):      # 2 = whole-warp issue, lane-predicated; 3 = whole-warp issue, uniform operands + elect.sync
                for reps in (8, 64):
                    _lib.check(lib.xhved_umma_issue_bench(N, reps, n_acc, a_tmem, mn, _lib.ptr(out), _lib.stream()), "bench")
                    _lib.check(lib.xhved_umma_issue_bench(N, reps, n_acc, a_tmem, mn, _lib.ptr(out), _lib.stream()), "bench")
                    torch.cuda.synchronize()
                    i, t = out.tolist()
                    rows.append(dict(N=N, n_acc=n_acc, a_in_tmem=a_tmem, mn_major=mn, reps=reps, issue_cycles=i, total_cycles=t,
                                     issue_per_mma=round(i / reps, 1), total_per_mma=round(t / reps, 1)))
                    print(json.dumps(rows[-1]))
