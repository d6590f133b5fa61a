"""Where the hot path sits inside the whole reference model ("the callers either side of the path").

Times the UNMODIFIED reference XLSTM_HVED (f_maps=4, 128^3, B=1) on the GPU with its stock PyTorch ViL / PoE path and with
xh.patch_model applied: inference forward (valid=True, subset 14) and a training-like forward+backward.  Also times the
hot-path modules alone inside the stock model (forward hooks would perturb it, so the ViL wrapper is timed stand-alone on a
tensor of the bottleneck shape).  Needs a reference tree (baseline/_ref or $XHVED_REFERENCE); test/dev tooling, not product.
"""
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


def main():
    if ref_loader.find_reference() is None:
        print(json.dumps({"unavailable": "no reference tree"}))
        return
    model = ref_loader.build_model(f_maps=4, seed=1).cuda()
    x = torch.rand(1, 4, 128, 128, 128, device="cuda")
    out = {}

    def infer():
        with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
            model(x, [14], valid=True)

    def train_step():
        model.zero_grad(set_to_none=True)
        with contextlib.redirect_stdout(io.StringIO()):
            seg, (mu_list, logvar_list), recon = model(x, [14], recon=True, valid=False)
        loss = seg.float().mean() + sum(r.float().mean() for r in recon) + sum(m.float().pow(2).mean() for m in mu_list)
        loss.backward()

    feat = torch.randn(1, 32, 16, 16, 16, device="cuda", requires_grad=True)

    def vil_alone():
        y = model.mViL(feat)
        y.sum().backward()

    for tag in ("stock", "patched", "patched_conv_norm"):
        if tag == "patched":
            xh.patch_model(model, conv_path=False)          # the ViL / S-MVAE hot path only
        elif tag == "patched_conv_norm":
            out["patch_counts"] = {k: v for k, v in xh.patch_model(model).items() if isinstance(v, int)}    # + K6
        model.eval()
        out[f"{tag}_infer_ms"] = round(timeit(infer), 2)
        model.train()
        out[f"{tag}_train_fwd_bwd_ms"] = round(timeit(train_step, iters=3, warm=1), 2)
        out[f"{tag}_mViL_fwd_bwd_ms"] = round(timeit(vil_alone, iters=10), 3)
    # the patched inference forward replayed as ONE CUDA graph (eager it is bound by ~900 launches from Python)
    model.eval()
    static_x = x.clone()
    # per-sample drop mask handed over as a device tensor: the subset-index form builds it on the host and copies it inside the
    # forward (RA_HVED.py:515-520), which a stream capture refuses
    drop = torch.zeros(1, 4, dtype=torch.bool, device="cuda")
    with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                model(static_x, [14], instance_missing=True, drop=drop, valid=True)
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            seg_static, _ = model(static_x, [14], instance_missing=True, drop=drop, valid=True)
        eager_seg, _ = model(x, [14], valid=True)
    graph.replay()
    out["patched_conv_norm_graph_infer_same_mask"] = ((seg_static > 0.5) == (eager_seg > 0.5)).float().mean().item()
    out["patched_conv_norm_graph_infer_ms"] = round(timeit(graph.replay, iters=10), 2)
    xh.unpatch_model(model)
    out["note"] = "reference XLSTM_HVED f_maps=4, one 128^3 volume, fp32 eager on one B200; mViL = the bottleneck ViL wrapper alone"
    print(json.dumps(out))


if __name__ == "__main__":
    main()
