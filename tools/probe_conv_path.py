#!/usr/bin/env python
"""Developer probe: how much the reference's conv path gains from CUDA-graph replay, channels_last_3d and
fp16 / bf16 autocast at inference, and what it costs in segmentation parity.  One 128^3 volume, batch 1, patched model."""
import contextlib, io, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import xlstm_hved_b200 as xh            # noqa: E402
from oracle import ref_loader           # noqa: E402


def timeit(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


model = ref_loader.build_model(f_maps=4, seed=1).cuda().eval()
xh.patch_model(model)
x = torch.rand(1, 4, 128, 128, 128, device="cuda")
drop = torch.zeros(1, 4, dtype=torch.bool, device="cuda")


def fwd(inp=x, **kw):
    with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
        return model(inp, [14], instance_missing=True, drop=drop, valid=True)[0]


out = {}
ref = fwd().float()
out["eager_fp32_ms"] = round(timeit(fwd), 2)
for name, dt in (("fp16", torch.float16), ("bf16", torch.bfloat16)):
    def f():
        with torch.autocast("cuda", dtype=dt):
            return fwd()
    y = f().float()
    out[f"autocast_{name}_ms"] = round(timeit(f), 2)
    out[f"autocast_{name}_same_mask"] = ((y > 0.5) == (ref > 0.5)).float().mean().item()
try:
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fwd()
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(g):
        yg = fwd()
    out["graph_fp32_ms"] = round(timeit(g.replay), 2)
    out["graph_same"] = torch.equal(yg.float(), ref) or ((yg > 0.5) == (ref > 0.5)).float().mean().item()
except Exception as e:
    out["graph_error"] = f"{type(e).__name__}: {str(e)[:200]}"
try:
    model_cl = model.to(memory_format=torch.channels_last_3d)
    xc = x.contiguous(memory_format=torch.channels_last_3d)
    y = fwd(xc).float()
    out["channels_last_fp32_ms"] = round(timeit(lambda: fwd(xc)), 2)
    out["channels_last_same_mask"] = ((y > 0.5) == (ref > 0.5)).float().mean().item()
except Exception as e:
    out["channels_last_error"] = f"{type(e).__name__}: {str(e)[:200]}"
print(json.dumps(out))
