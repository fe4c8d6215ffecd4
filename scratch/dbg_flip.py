import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import stsgcn as onet
from tests.helpers import make_pair
from coskad_b200 import train
m, sd = make_pair('stse', 16, seed=0)
m.train()
x = onet.synth_windows(64, seed=7)
dt = torch.float64
params = {k: (v.to(dt).clone().requires_grad_(True) if v.is_floating_point() and 'running' not in k and k != 'c' else (v.to(dt) if v.is_floating_point() else v.clone())) for k, v in sd.items()}
h, acts = onet.layer_stack(x.to(dt), params, 'encoder', training=True, new_stats={}, return_all=True)
for a in acts: a.retain_grad()
z = torch.nn.functional.linear(h.reshape(64, -1), params['btlnk.weight'], params['btlnk.bias'])
(z ** 2).mean().backward()
for it in range(4):
    for p in m.parameters(): p.grad = None
    hs = []; hc = x.cuda()
    for layer in m.encoder.model:
        hc = train.layer_forward(layer, hc, True); hc.retain_grad(); hs.append(hc)
    zc = train.linear_reduce(hc.reshape(64, -1), m.btlnk.weight, m.btlnk.bias)
    (zc ** 2).mean().backward()
    msg = []
    for i in (3, 2):
        o = hs[i].detach().cpu().double(); r = acts[i].detach()
        flip = (o >= 0) != (r >= 0)
        slope = float(params[f'encoder.model.{i}.prelu.weight'].detach())
        dref = acts[i].grad
        # effect of the flipped masks on d beta (sum ds) per channel
        delta = torch.where(flip, (1 - slope) * dref * torch.where(o >= 0, 1.0, -1.0), torch.zeros_like(dref)).sum(dim=(0, 2, 3))
        gb = dict(m.named_parameters())[f'encoder.model.{i}.tcn.1.bias'].grad.cpu().double()
        rb = params[f'encoder.model.{i}.tcn.1.bias'].grad
        msg.append(f'L{i}: flips {int(flip.sum())}, |pre| at flips max {float(r[flip].abs().max()) if flip.any() else 0:.1e}, dbeta err {float((gb-rb).abs().max()/rb.abs().max()):.1e}, predicted from flips {float(delta.abs().max()/rb.abs().max()):.1e}, residual {float((gb-rb-delta).abs().max()/rb.abs().max()):.1e}')
    print(f'iter {it}: ' + ' | '.join(msg))
