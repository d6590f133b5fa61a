"""Is the patched conv path as accurate as the stock one?  (dev tooling; needs baseline/_ref)
(1) per InstanceNorm3d / BatchNorm3d call of a real forward: error of PyTorch's fp32 kernel and of K6 against an fp64 evaluation
    of the same input; (2) whole model: stock fp32, patched fp32 (hot path only / + conv path) against the stock model in fp64."""
import contextlib, copy, io, json, os, sys
import torch
import torch.nn as nn
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import xlstm_hved_b200 as xh            # noqa: E402
from oracle import ref_loader           # noqa: E402

torch.backends.cudnn.allow_tf32 = "--tf32" in sys.argv
torch.backends.cuda.matmul.allow_tf32 = "--tf32" in sys.argv
ns = ref_loader.load_reference()
model = ref_loader.build_model(f_maps=4, seed=1).cuda().eval()
size, subset = (64, 96, 64), 7
torch.manual_seed(5)
x = torch.rand(1, 4, *size, device="cuda")
for m in range(4):
    if m not in ns.RA_HVED.SUBSETS_MODALITIES[subset]:
        x[:, m] = 0

worst = {"torch": (0, ""), "k6": (0, "")}
rows = []
def hook(name):
    def fn(mod, inp, out):
        a = inp[0].detach()
        with torch.no_grad():
            if isinstance(mod, nn.InstanceNorm3d):
                ref = torch.nn.functional.instance_norm(a.double(), eps=mod.eps)
                mine = xh.modules.instance_norm_act(a.clone(), eps=mod.eps)
            else:
                ref = torch.nn.functional.batch_norm(a.double(), mod.running_mean.double(), mod.running_var.double(), mod.weight.double(),
                                                     mod.bias.double(), False, 0.0, mod.eps)
                mine = xh.modules.batch_norm_act(a.clone(), mod.weight, mod.bias, mod.running_mean, mod.running_var, False, 0.0, mod.eps)
            et = (out.detach().double() - ref).abs().max().item()      # hooks see the output before the in-place LeakyReLU runs
            em = (mine.double() - ref).abs().max().item()
            flat = a.double().reshape(a.shape[0] * a.shape[1], -1)
            ratio = (flat.mean(-1).abs() / flat.std(-1, unbiased=False).clamp_min(1e-30)).max().item()
        rows.append((name, tuple(a.shape), et, em, ratio))
    return fn
hs = [m.register_forward_hook(hook(n)) for n, m in model.named_modules() if type(m) in (nn.InstanceNorm3d, nn.BatchNorm3d)]
with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
    seg32, _ = model(x, [subset], valid=True)
for h in hs:
    h.remove()
rows.sort(key=lambda r: -max(r[2], r[3]))
print("layer, shape, |torch32 - fp64|max, |k6 - fp64|max, max |mean|/std of the input planes")
for r in rows[:12]:
    print(r)
print("layers where k6 is worse than torch by > 2x:", sum(1 for r in rows if r[3] > 2 * r[2] + 1e-7), "of", len(rows))

m64 = copy.deepcopy(model).double()
with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
    seg64, _ = m64(x.double(), [subset], valid=True)
    xh.patch_model(model, conv_path=False)
    seg_hot, _ = model(x, [subset], valid=True)
    xh.unpatch_model(model)
    xh.patch_model(model)
    seg_all, _ = model(x, [subset], valid=True)
    xh.unpatch_model(model)
def agree(p, q):
    p, q = p.double(), q.double()
    return dict(mask=((p > 0.5) == (q > 0.5)).double().mean().item(), argmax=(p.argmax(1) == q.argmax(1)).double().mean().item(),
                max_dp=(p - q).abs().max().item())
print(json.dumps({"tf32": torch.backends.cudnn.allow_tf32, "stock32_vs_fp64": agree(seg32, seg64), "hotpath_vs_fp64": agree(seg_hot, seg64),
                  "hot+conv_vs_fp64": agree(seg_all, seg64), "hot+conv_vs_stock32": agree(seg_all, seg32), "hotpath_vs_stock32": agree(seg_hot, seg32)}))
