import cProfile, pstats, sys, torch, io
sys.path.insert(0, '/root/repo')
import xlstm_hved_b200 as xh
wrap = xh.ViLLayer3D(dim=32).cuda()
feat = torch.randn(1, 32, 16, 16, 16, device="cuda", requires_grad=True)
gy = torch.randn(1, 32, 16, 16, 16, device="cuda")
def step():
    y = wrap(feat)
    y.backward(gy)
for _ in range(20): step()
torch.cuda.synchronize()
import time
t=time.perf_counter()
for _ in range(200): step()
torch.cuda.synchronize()
print("ms per fwd+bwd", (time.perf_counter()-t)/200*1e3)
pr = cProfile.Profile(); pr.enable()
for _ in range(200): step()
torch.cuda.synchronize(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28); print(s.getvalue()[:6000])
