import sys, torch, time
sys.path.insert(0,'.')
from xlstm_hved_b200 import ops
B=32
dev='cuda'
for C,d in ((1,64),(2,32)):
    mu=torch.randn(5,B,C,d,d,d,device=dev); lv=torch.randn(5,B,C,d,d,d,device=dev)
    noise=torch.randn(1,B,C,d,d,d,device=dev); gz=torch.randn(1,B,C,d,d,d,device=dev)
    n=mu[0].numel()
    for name,fn,byt in (('fwd',lambda: ops.poe_fwd(mu,lv,[(0,1,2,3)],noise=noise,want_kld=True),56),('bwd',lambda: ops.poe_bwd(mu,lv,[(0,1,2,3)],noise=noise,g_z=gz,kld_scale=[0.1]),88)):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): fn()
        e1.record(); torch.cuda.synchronize()
        ms=e0.elapsed_time(e1)/20
        print(f'level C={C} d={d} {name}: {ms*1000:.1f} us  {n*byt/ms/1e6:.0f} GB/s')
