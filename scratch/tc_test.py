import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from coskad_b200 import _lib
ctx = _lib.context(0)
def pack(W):  # W [N,K] -> canonical K-major image [k/4][n/8][n%8][k%4]
    N, K = W.shape
    return W.view(N // 8, 8, K // 4, 4).permute(2, 0, 1, 3).contiguous().view(-1)
def trunc(x):
    return (x.view(torch.int32) & -8192).view(torch.float32)
torch.manual_seed(0)
for K, N in ((16, 16), (32, 32), (64, 64), (32, 16), (64, 32)):
    A = torch.randn(128, K, device='cuda')
    W = torch.randn(N, K, device='cuda') / K ** 0.5
    Wh = trunc(W); Wl = trunc(W - Wh)
    ref = (A.double() @ W.double().t())
    for swap in (0, 1):
        out = torch.zeros(128, N, device='cuda')
        ph, pl = pack(Wh), pack(Wl)
        torch.cuda.synchronize(); print('launch', K, N, swap, ph.shape, ph.dtype, flush=True)
        rc = ctx.lib.coskad_debug_tc_mix(ctx.h, A.data_ptr(), ph.data_ptr(), pl.data_ptr(), K, N, swap, out.data_ptr(), 0)
        ctx.check(rc, 'tc')
        torch.cuda.synchronize()
        err = float((out.double() - ref).abs().max() / ref.abs().max())
        tf32 = (trunc(A).double() @ Wh.double().t())
        err1 = float((tf32 - ref).abs().max() / ref.abs().max())
        print(f'K={K} N={N} swap={swap}: max rel err {err:.3e}   (single-pass tf32 would be {err1:.1e})')
