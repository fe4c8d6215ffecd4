import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import stsgcn as onet
from tests.helpers import make_pair
from coskad_b200 import train
m, sd = make_pair('stse', 16, seed=0)
m.train()
x = onet.synth_windows(64, seed=7).cuda()
def run(mode):
    hs = []
    hc = x
    for layer in m.encoder.model:
        hc = train.layer_forward(layer, hc, True); hc.retain_grad(); hs.append(hc)
    zc = train.linear_reduce(hc.reshape(64, -1), m.btlnk.weight, m.btlnk.bias)
    (zc ** 2).mean().backward()
    if mode == 'sleep': time.sleep(1.0)
    if mode == 'streamsync': torch.cuda.current_stream().synchronize()
    if mode == 'devsync': torch.cuda.synchronize()
    g = [h.grad.cpu().clone() for h in hs]
    torch.cuda.synchronize()
    g2 = [h.grad.cpu().clone() for h in hs]
    print(mode, ['%.2e' % float((a - b).abs().max() / b.abs().max()) for a, b in zip(g, g2)])
for mode in ('none', 'sleep', 'streamsync', 'devsync', 'none'):
    run(mode)
