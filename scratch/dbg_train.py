import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import stsgcn as onet
from tests.helpers import make_pair
m, sd = make_pair('stse', 16, seed=0)
m.train()
x = onet.synth_windows(64, seed=7)
res = {}
for dt in (torch.float32, torch.float64):
    params = {k: (v.to(dt).clone().requires_grad_(True) if v.is_floating_point() and 'running' not in k and k != 'c' else (v.to(dt) if v.is_floating_point() else v.clone())) for k, v in sd.items()}
    z = onet.stse_forward(x.to(dt), params, training=True, new_stats={})
    (z ** 2).mean().backward()
    res[dt] = ({k: p.grad.double() for k, p in params.items() if p.is_floating_point() and p.requires_grad}, z.detach().double())
z = m(x.cuda())
(z ** 2).mean().backward()
g64, z64 = res[torch.float64]; g32, z32 = res[torch.float32]
print('z: ours vs f64 %.2e ; torch f32 vs f64 %.2e' % (float((z.detach().cpu().double() - z64).abs().max() / z64.abs().max()), float((z32 - z64).abs().max() / z64.abs().max())))
for k, p in m.named_parameters():
    if k.endswith('.0.bias') and ('tcn' in k or 'residual' in k): continue
    sc = g64[k].abs().max() + 1e-30
    eo = float((p.grad.cpu().double() - g64[k]).abs().max() / sc); er = float((g32[k] - g64[k]).abs().max() / sc)
    flag = '  <<<' if eo > 3 * er + 1e-6 else ''
    print(f'{k:40s} ours {eo:.2e}  torch-f32 {er:.2e}{flag}')
