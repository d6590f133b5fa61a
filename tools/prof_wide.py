#!/usr/bin/env python
"""Developer tool: torch-profiler table of a wide (unfused) ViLBlock forward + backward."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import xlstm_hved_b200 as xh  # noqa: E402

dim = int(sys.argv[1]) if len(sys.argv) > 1 else 128
blk = xh.ViLBlock(dim, xh.SequenceTraversal.ROWWISE_FROM_TOP_LEFT).cuda()
x = torch.randn(16, 4096, dim, device="cuda", requires_grad=True)
gy = torch.randn_like(x)
for _ in range(2):
    blk(x).backward(gy)
torch.cuda.synchronize()
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA, torch.profiler.ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        blk(x).backward(gy)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=18, max_name_column_width=60))
