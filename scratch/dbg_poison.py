import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import stsgcn as onet
from tests.helpers import make_pair
m, sd = make_pair('stse', 16, seed=0)
m.train()
x = onet.synth_windows(64, seed=7)
dt = torch.float64
params = {k: (v.to(dt).clone().requires_grad_(True) if v.is_floating_point() and 'running' not in k and k != 'c' else (v.to(dt) if v.is_floating_point() else v.clone())) for k, v in sd.items()}
z = onet.stse_forward(x.to(dt), params, training=True, new_stats={})
(z ** 2).mean().backward()
xc = x.cuda()
def poison(val):
    t = torch.full((256 * 1024 * 1024 // 4,), val, device='cuda')   # 256 MB
    u = [torch.full((n,), val, device='cuda') for n in (64, 256, 1024, 4096, 65536, 1 << 20) for _ in range(8)]
    del t, u
for it, val in enumerate([0.0, float('nan'), 1e30, 0.0, float('nan')]):
    poison(val)
    for p in m.parameters(): p.grad = None
    zc = m(xc)
    (zc ** 2).mean().backward()
    errs = {}
    for k in ('encoder.model.3.tcn.1.bias', 'encoder.model.3.tcn.1.weight', 'encoder.model.3.gcn.A', 'encoder.model.0.gcn.A', 'btlnk.weight'):
        g = dict(m.named_parameters())[k].grad.cpu().double(); r = params[k].grad
        errs[k.replace('encoder.model.', 'L')] = float((g - r).abs().max() / r.abs().max())
    print(f'iter {it} poison={val}: ' + '  '.join(f'{k} {v:.1e}' for k, v in errs.items()))
