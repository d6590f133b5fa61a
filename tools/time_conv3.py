"""Device time of the K10 / K8 kernels at the model's shapes (dev tooling)."""
import json, sys, torch
sys.path.insert(0, '/root/repo')
from xlstm_hved_b200 import ops

def t(fn, iters=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return round(e0.elapsed_time(e1) / iters * 1e3, 1)

out = {}
for (cin, cout, res, dt) in ((4, 4, 128, torch.float32), (4, 4, 128, torch.float16), (12, 4, 128, torch.float16), (8, 8, 64, torch.float16),
                             (16, 16, 32, torch.float16), (48, 16, 32, torch.float16)):
    x = torch.randn(1, cin, res, res, res, device="cuda").to(dt)
    dy = torch.randn(1, cout, res, res, res, device="cuda").to(dt)
    w = torch.randn(cout, cin, 3, 3, 3, device="cuda") * 0.1
    key = f"{cin}->{cout}@{res}^3 {str(dt)[6:]}"
    out[key] = {"fwd_us": t(lambda: ops.conv3_fwd(x, w)), "dgrad_us": t(lambda: ops.conv3_bwd(x, w, dy, want_dw=False)),
                "wgrad_us": t(lambda: ops.conv3_bwd(x, w, dy, want_dx=False, want_db=True))}
print(json.dumps(out))
