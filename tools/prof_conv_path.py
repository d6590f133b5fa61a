"""A few launches of the conv-path kernels (K6, K7, K8) at the model's first-level shape for ncu (dev tooling)."""
import sys, torch
sys.path.insert(0, '/root/repo')
from xlstm_hved_b200 import ops
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
x = torch.rand(n, 4, 128, 128, 128, device="cuda")
dy = torch.randn_like(x)
w3, w7 = torch.randn(4, 27, device="cuda") * 0.2, torch.randn(4, 343, device="cuda") * 0.02
for _ in range(3):
    y, mean, rstd = ops.norm_act_fwd(x, slope=0.01)
    ops.norm_act_bwd(x, dy, mean, rstd, slope=0.01)
    ops.dwconv3_fwd(x, w3)
    ops.dwconv3_bwd(x, w3, dy)
    gate = ops.gate7_fwd(x, w7)
    ops.gate7_bwd(x, w7, gate, dy[:, :1].contiguous())
    w1 = w3[:, :4].contiguous()
    ops.pwconv_fwd(x, w1)
    ops.pwconv_bwd(x, w1, dy, want_db=True)
torch.cuda.synchronize()
