import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from coskad_b200 import synth
dev = torch.device('cuda', 0)
g = torch.Generator(device=dev).manual_seed(999)
x = torch.empty(131072, 2, 12, 17, device=dev); synth.synth_windows_(x, g)
ae = synth.make_model('stsae', 8, seed=0, device=dev)
c = torch.zeros(8, device=dev)
for _ in range(4):
    ae.autoencode_score(x, center=c, want_xhat=False)
torch.cuda.synchronize()
