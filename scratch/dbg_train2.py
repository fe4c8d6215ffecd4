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
# ours
hs = []
hc = x.cuda()
for layer in m.encoder.model:
    hc = train.layer_forward(layer, hc, True)
    hc.retain_grad()
    hs.append(hc)
zc = train.linear_reduce(hc.reshape(64, -1), m.btlnk.weight, m.btlnk.bias)
(zc ** 2).mean().backward()
torch.cuda.synchronize()
for rep in range(2):
  for i in range(4):
    a, b = hs[i].detach().cpu().double(), acts[i].detach()
    ga, gb = hs[i].grad.cpu().double(), acts[i].grad
    print(rep, f'layer {i}: act err {float((a-b).abs().max()/b.abs().max()):.2e}  act-grad err {float((ga-gb).abs().max()/gb.abs().max()):.2e}')
# isolate layer 3 backward with the oracle's exact dout
layer = m.encoder.model[3]
xin = acts[2].detach().float().cuda().requires_grad_(True)
out = train.layer_forward(layer, xin, True)
for p in layer.parameters(): p.grad = None
out.backward(acts[3].grad.float().cuda())
print('layer3 isolated: dX err %.2e' % float((xin.grad.cpu().double() - acts[2].grad).abs().max() / acts[2].grad.abs().max()))
for k, p in layer.named_parameters():
    r = params['encoder.model.3.' + k].grad
    print(f'   {k:22s} err {float((p.grad.cpu().double()-r).abs().max()/(r.abs().max()+1e-30)):.2e}  refmax {float(r.abs().max()):.2e}')
print('--- (a) isolated layer3 with OUR hs[2] and OUR dout')
xin = hs[2].detach().clone().requires_grad_(True)
out = train.layer_forward(layer, xin, True)
out.backward(hs[3].grad.clone())
print('   dX err vs oracle %.2e ; vs chained ours %.2e' % (float((xin.grad.cpu().double() - acts[2].grad).abs().max() / acts[2].grad.abs().max()), float((xin.grad - hs[2].grad).abs().max() / hs[2].grad.abs().max())))
print('--- (b) chain layers 2->3 from oracle acts[1]')
x1 = acts[1].detach().float().cuda().requires_grad_(True)
h2 = train.layer_forward(m.encoder.model[2], x1, True); h2.retain_grad()
h3 = train.layer_forward(m.encoder.model[3], h2, True)
h3.backward(acts[3].grad.float().cuda())
print('   h2 grad err %.2e ; x1 grad err %.2e' % (float((h2.grad.cpu().double() - acts[2].grad).abs().max() / acts[2].grad.abs().max()), float((x1.grad.cpu().double() - acts[1].grad).abs().max() / acts[1].grad.abs().max())))
print('--- (c) same as (b) but clone between layers')
x1 = acts[1].detach().float().cuda().requires_grad_(True)
h2 = train.layer_forward(m.encoder.model[2], x1, True); h2c = h2.clone(); h2c.retain_grad()
h3 = train.layer_forward(m.encoder.model[3], h2c, True)
h3.backward(acts[3].grad.float().cuda())
print('   h2 grad err %.2e ; x1 grad err %.2e' % (float((h2c.grad.cpu().double() - acts[2].grad).abs().max() / acts[2].grad.abs().max()), float((x1.grad.cpu().double() - acts[1].grad).abs().max() / acts[1].grad.abs().max())))
