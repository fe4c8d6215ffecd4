import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import geoopt_math as ogm
from coskad_b200 import gmath
K = torch.tensor(-1.)
gen = torch.Generator().manual_seed(9)
z = (torch.randn(512, 16, generator=gen) * torch.logspace(-2, 0.6, 512)[:, None])
c = torch.randn(16, generator=gen) * 0.1
w = torch.rand(512, generator=gen)
for with_project in (True, False):
    res = {}
    for dt in (torch.float32, torch.float64):
        zz = z.to(dt).clone().detach().requires_grad_(True)
        x = ogm.expmap0(zz, k=K.to(dt))
        if with_project:
            x = ogm.project(x, k=K.to(dt), eps=4e-3)
        (ogm.dist(c.to(dt), x, k=K.to(dt)) * w.to(dt)).sum().backward()
        res[dt] = zz.grad.clone()
    got = gmath.poincare_score_bwd(z.cuda(), c.cuda(), w.cuda(), with_project).cpu().double()
    r32, r64 = res[torch.float32].double(), res[torch.float64]
    nz = z.norm(dim=-1)
    for lo, hi in ((0, 0.5), (0.5, 1.5), (1.5, 3.1), (3.1, 15), (15, 100)):
        m = (nz >= lo) & (nz < hi)
        if m.sum() == 0: continue
        sc = r64[m].abs().amax(dim=-1, keepdim=True) + 1e-30
        e_ours = ((got[m] - r64[m]).abs() / sc).max().item()
        e_ref = ((r32[m] - r64[m]).abs() / sc).max().item()
        print(f'proj={with_project} |z| in [{lo},{hi}) n={int(m.sum())}: ours vs f64 {e_ours:.2e}; torch f32 vs f64 {e_ref:.2e}; grad scale {sc.max().item():.2e}')
