#!/usr/bin/env python
"""Developer tool: torch-profiler table of one wide (not fused) ViL block pair forward + backward.  argv: dim B"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import xlstm_hved_b200 as xh
dim = int(sys.argv[1]) if len(sys.argv) > 1 else 128
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
S = 4096
dev = torch.device("cuda", 0)
blocks = [xh.ViLBlock(dim, xh.SequenceTraversal.ROWWISE_FROM_TOP_LEFT).to(dev), xh.ViLBlock(dim, xh.SequenceTraversal.ROWWISE_FROM_BOT_RIGHT).to(dev)]
x = torch.randn(B, S, dim, device=dev)
gy = torch.randn(B, S, dim, device=dev)
def step():
    xt = x.detach().requires_grad_()
    y = blocks[1](blocks[0](xt))
    y.backward(gy)
for _ in range(3):
    step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=70))
