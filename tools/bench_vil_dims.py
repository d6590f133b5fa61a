"""ViLBlock pair (TOP_LEFT + BOT_RIGHT) forward+backward at the model widths dim 16 / 32 / 64 (f_maps 2 / 4 / 8: fused K2 / K3
kernels) and dim 128 / 256 (f_maps 16 / 32: the cell on the tcgen05 kernels at head dim 64 / 128, the per-token glue as torch
ops), per-kernel times from the library's CUDA-event profiler.  Prints one JSON line per width."""
import ctypes, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import xlstm_hved_b200 as xh
from xlstm_hved_b200 import _lib


def run(dim, B=32, S=4096, iters=10):
    dev = torch.device("cuda", 0)
    blocks = [xh.ViLBlock(dim, xh.SequenceTraversal.ROWWISE_FROM_TOP_LEFT).to(dev), xh.ViLBlock(dim, xh.SequenceTraversal.ROWWISE_FROM_BOT_RIGHT).to(dev)]
    torch.manual_seed(dim)
    side = round(S ** (1 / 3))
    x = torch.randn(B, dim, side, side, side, device=dev)
    gy = torch.randn(B, S, dim, device=dev)

    def step():
        xt = x.detach().requires_grad_()
        y = blocks[1](blocks[0](xt.reshape(B, dim, -1).transpose(-1, -2)))
        y.backward(gy)

    for _ in range(3):
        step()
    lib = _lib.load_library()
    nk = lib.xhved_profile_kernel_count()
    names = [lib.xhved_profile_kernel_name(i).decode() for i in range(nk)]
    ms, cnt = (ctypes.c_float * nk)(), (ctypes.c_int * nk)()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        step()
    e1.record()
    torch.cuda.synchronize()
    lib.xhved_profile_enable(1)
    lib.xhved_profile_read(ms, cnt, nk)
    for _ in range(iters):
        step()
    lib.xhved_profile_read(ms, cnt, nk)
    lib.xhved_profile_enable(0)
    kern = {names[i]: round(ms[i] / iters, 4) for i in range(nk) if cnt[i]}
    return {"dim": dim, "B": B, "S": S, "path": "fused K2 / cell / K3" if dim in xh.modules.FUSED_DIMS else "library GEMMs + fused glue kernels + cell kernels",
            "eager_ms_per_pair_fwd_bwd": round(e0.elapsed_time(e1) / iters, 4),
            "kernel_ms": dict(sorted(kern.items(), key=lambda kv: -kv[1])), "kernel_ms_total": round(sum(kern.values()), 4)}


if __name__ == "__main__":
    for dim in (16, 32, 64):
        print(json.dumps(run(dim)))
    for dim in (128, 256):                      # roofline shapes B = 16, S = 4096
        print(json.dumps(run(dim, B=16)))
