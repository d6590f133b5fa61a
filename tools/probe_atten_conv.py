"""AttenModule2's spatial gate (buildingblocks.py:259-301): depthwise 7^3 conv (expansion 4) + 1x1x1 conv to one channel, stock
vs the algebraically composed dense Cin -> 1 7^3 convolution on cuDNN.  fwd and fwd+bwd times (dev tooling)."""
import json, sys, torch, torch.nn as nn, torch.nn.functional as F

def timeit(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

for tf32 in (True, False):
    torch.backends.cudnn.allow_tf32 = tf32
    for G, res in ((4, 128), (2, 128), (4, 64)):
        dw = nn.Conv3d(G, 4 * G, 7, padding=3, groups=G).cuda()
        pw = nn.Conv3d(4 * G, 1, 1).cuda()
        x = torch.randn(1, G, res, res, res, device="cuda", requires_grad=True)
        def stock(): return torch.sigmoid(pw(dw(x)))
        def composed():
            w = (dw.weight.view(G, 4, 7, 7, 7) * pw.weight.view(G, 4, 1, 1, 1)).sum(1).unsqueeze(0)
            b = (pw.weight.view(-1) * dw.bias).sum() + pw.bias
            return torch.sigmoid(F.conv3d(x, w, b, padding=3))
        out = {"tf32": tf32, "G": G, "res": res, "max_diff": (stock() - composed()).abs().max().item()}
        for name, fn in (("stock", stock), ("composed", composed)):
            with torch.no_grad():
                out[name + "_fwd_ms"] = round(timeit(fn), 3)
            def fb():
                y = fn(); y.sum().backward()
            out[name + "_fwd_bwd_ms"] = round(timeit(fb, iters=5, warm=2), 3)
        print(json.dumps(out))
