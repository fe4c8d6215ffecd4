"""throughput of the other fused eval variants: auto-encoder (decoder + reconstruction score, latent 8) and D=8 encoder"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from coskad_b200 import synth, _lib
dev = torch.device('cuda', 0)
B = 262144
g = torch.Generator(device=dev).manual_seed(999)
x = torch.empty(B, 2, 12, 17, device=dev); synth.synth_windows_(x, g)
def timeit(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
ae = synth.make_model('stsae', 8, seed=0, device=dev)
c = torch.zeros(8, device=dev)
ms = timeit(lambda: ae.autoencode_score(x, center=c, want_xhat=False))
print(f'STSAE D=8 encode+decode+rec score: {ms:.2f} ms -> {B/ms*1e3/1e6:.2f} M windows/s')
e8 = synth.make_model('stse', 8, seed=0, device=dev)
ms = timeit(lambda: e8.encode_score(x, _lib.SCORE_EUCLID, center=c, want_latent=False))
print(f'STSE D=8 euclid score: {ms:.2f} ms -> {B/ms*1e3/1e6:.2f} M windows/s')
e16 = synth.make_model('stse', 16, seed=0, device=dev)
c16 = torch.zeros(16, device=dev)
ms = timeit(lambda: e16.encode_score(x, _lib.SCORE_POINCARE, center=c16, want_latent=False))
print(f'STSE D=16 poincare score: {ms:.2f} ms -> {B/ms*1e3/1e6:.2f} M windows/s')
# spherical VAE (use_vae): fc_mean + fc_var as a 9-row head (NDQ = 3), cosine score to the mean direction
from coskad_b200 import spherical
torch.manual_seed(0)
sv = spherical.STSVAE(input_dim=2, layer_channels=[32, 16, 32], hidden_dimension=64, latent_dim=8, n_frames=12, n_joints=17)
synth.randomize_bn_(sv, 0)
sv = sv.to(dev).eval()
mv = torch.nn.functional.normalize(torch.randn(1, 8, device=dev), dim=-1)
ms = timeit(lambda: sv.cosine_scores(x, mean_vector=mv, sample=False))
print(f'STSVAE D=8 (9-row head) cosine score on the mean direction: {ms:.2f} ms -> {B/ms*1e3/1e6:.2f} M windows/s')
# trajectory front end, device-resident: stride-1 windows of 300-frame persons x 5 transforms
import math
from coskad_b200 import _lib as L
plen, per = 300, 300 - 12 + 1
persons = B // (5 * per)
base = torch.arange(per, device=dev).repeat(persons) + torch.arange(persons, device=dev).repeat_interleave(per) * plen
rows = base.repeat(5).contiguous()
tr = torch.arange(5, device=dev, dtype=torch.int32).repeat_interleave(base.numel()).contiguous()
traj = (torch.randn(persons * plen, 34, device=dev) * 0.4).contiguous()
c45 = math.cos(math.radians(45.0))
mats = torch.tensor([[[1, 0, 0], [0, 1, 0]], [[-1, 0, 0], [0, 1, 0]], [[0, -1, 0], [1, 0, 0]], [[0, 1, 0], [1, 0, 0]],
                     [[c45, -c45, 0], [c45, c45, 0]]], dtype=torch.float32, device=dev)
ms = timeit(lambda: e16.encode_score_traj(traj, rows, tr, mats, flavour=1, center=c16, want_latent=False))
print(f'STSE D=16 from trajectories ({rows.numel()} windows, 5 transforms, device-resident): {ms:.2f} ms -> {rows.numel()/ms*1e3/1e6:.2f} M windows/s')
