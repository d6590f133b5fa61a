"""Kernel-level breakdown of the reference model's inference forward / training step on the GPU (SURVEY 8f rank 1 evidence).
Test / dev tooling: needs a reference tree (baseline/_ref)."""
import contextlib, io, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_loader           # noqa: E402
from torch.profiler import profile, ProfilerActivity

model = ref_loader.build_model(f_maps=4, seed=1).cuda().eval()
if "--patched" in sys.argv:
    import xlstm_hved_b200 as xh
    print(xh.patch_model(model))
x = torch.rand(1, 4, 128, 128, 128, device="cuda")

def infer():
    with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
        model(x, [14], valid=True)

if "--train" in sys.argv:
    model.train()

    def infer():                                     # training-like forward + backward (as tools/bench_model.py)
        model.zero_grad(set_to_none=True)
        with contextlib.redirect_stdout(io.StringIO()):
            seg, (mu_list, logvar_list), recon = model(x, [14], recon=True, valid=False)
        loss = seg.float().mean() + sum(r.float().mean() for r in recon) + sum(m.float().pow(2).mean() for m in mu_list)
        loss.backward()

for _ in range(2):
    infer()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
    infer()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=35, max_name_column_width=70))
print(prof.key_averages(group_by_input_shape=True).table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=40, max_shapes_column_width=90))

if "--train" in sys.argv:
    sys.exit(0)
# per top-level module wall time (synchronised hooks)
times = {}
def mk(name):
    ev = {}
    def pre(m, a):
        torch.cuda.synchronize(); ev["t"] = torch.cuda.Event(enable_timing=True); ev["t"].record()
    def post(m, a, o):
        e = torch.cuda.Event(enable_timing=True); e.record(); torch.cuda.synchronize()
        times[name] = times.get(name, 0.0) + ev["t"].elapsed_time(e)
    return pre, post
for name, mod in model.named_children():
    subs = list(mod.named_children()) if isinstance(mod, torch.nn.ModuleList) else []
    pre, post = mk(name); mod.register_forward_pre_hook(pre); mod.register_forward_hook(post)
    for n2, m2 in subs:
        pre, post = mk(f"{name}.{n2}"); m2.register_forward_pre_hook(pre); m2.register_forward_hook(post)
infer()
print(json.dumps({k: round(v, 2) for k, v in sorted(times.items(), key=lambda kv: -kv[1])}))
