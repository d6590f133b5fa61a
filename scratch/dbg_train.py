import contextlib, io, sys, torch
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
from oracle import ref_loader
import xlstm_hved_b200 as xh
ns = ref_loader.load_reference()
model = ref_loader.build_model(f_maps=4, seed=1).cuda().train()
x = torch.rand(1,4,64,64,64, device='cuda', generator=torch.Generator('cuda').manual_seed(9))
rec = {}
def run(tag):
    model.zero_grad(set_to_none=True)
    torch.manual_seed(11)
    zs = []
    orig = ns.RA_HVED.reparametrize
    def hook(mu, logvar, valid=False):
        z = orig(mu, logvar, valid); zs.append(z.detach().clone()); return z
    ns.RA_HVED.reparametrize = hook
    with contextlib.redirect_stdout(io.StringIO()):
        seg,(mu_list,lv_list),recon = model(x,[14],recon=True)
        kld = sum(ns.loss.compute_KLD(mu_list[l], lv_list[l],[14]) for l in range(4))/4
    ns.RA_HVED.reparametrize = orig
    loss = seg.mean()+0.2*((recon[0]-x)**2).mean()+0.2*kld
    loss.backward()
    return loss.item(), zs, {n:p.grad.detach().clone() for n,p in model.named_parameters() if p.grad is not None}, kld.item()
l0,z0,g0,k0 = run('ref')
l0b,z0b,g0b,k0b = run('ref2')
xh.patch_model(model)
l1,z1,g1,k1 = run('new')
xh.unpatch_model(model)
print('loss',l0,l0b,l1,'kld',k0,k1)
for a,b,c in zip(z0,z0b,z1): print('z ref-vs-ref', (a-b).abs().max().item(), 'ref-vs-new', (a-c).abs().max().item(), a.abs().max().item())
rel=lambda a,b: ((a-b).norm()/b.norm().clamp_min(1e-30)).item()
errs = sorted([(rel(g1[n],g0[n]), rel(g0b[n],g0[n]), n, g0[n].norm().item()) for n in g0], reverse=True)
for e in errs[:25]: print('%.4f  (ref-vs-ref %.4f)  %s  |g|=%.3e'%e)
