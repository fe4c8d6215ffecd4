import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import stsgcn as onet
sd = onet.init_state_dict('stse', latent_dim=16, seed=0)
x = onet.synth_windows(64, seed=7)
dt = torch.float64
params = {k: (v.to(dt).clone().requires_grad_(True) if v.is_floating_point() and 'running' not in k and k != 'c' else (v.to(dt) if v.is_floating_point() else v.clone())) for k, v in sd.items()}
h, acts = onet.layer_stack(x.to(dt), params, 'encoder', training=True, new_stats={}, return_all=True)
for a in acts: a.retain_grad()
z = torch.nn.functional.linear(h.reshape(64, -1), params['btlnk.weight'], params['btlnk.bias'])
(z ** 2).mean().backward()
dout = acts[3].grad  # [64,64,12,17]
out = acts[3].detach()
slope = float(params['encoder.model.3.prelu.weight'])
ds = torch.where(out >= 0, dout, slope * dout)
S = ds.sum(dim=(0, 2, 3)); SA = ds.abs().sum(dim=(0, 2, 3))
print('cancellation ratio sum|ds|/|sum ds| per channel: median %.1f max %.1f' % (float((SA / S.abs()).median()), float((SA / S.abs()).max())))
print('max|S| %.3e, typical sum|ds| %.3e' % (float(S.abs().max()), float(SA.mean())))
# float32 sequential accumulate emulation error relative to max|S|
ds32 = ds.float()
S32 = ds32.permute(1, 0, 2, 3).reshape(64, -1).cumsum(dim=1)[:, -1].double()
print('float32 cumsum error / max|S|: %.2e' % float((S32 - S).abs().max() / S.abs().max()))
