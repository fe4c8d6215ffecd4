import sys, os, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from coskad_b200 import _lib
print('main thread', threading.get_ident(), 'stream', torch.cuda.current_stream().cuda_stream, 'dev', torch.cuda.current_device())
class F(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        print('fwd thread', threading.get_ident(), 'stream', torch.cuda.current_stream(x.device).cuda_stream, _lib.stream_ptr(x.device))
        return x * 2
    @staticmethod
    def backward(ctx, g):
        print('bwd thread', threading.get_ident(), 'stream', torch.cuda.current_stream(g.device).cuda_stream, _lib.stream_ptr(g.device), 'current dev', torch.cuda.current_device())
        return g * 2
x = torch.randn(4, device='cuda', requires_grad=True)
F.apply(x).sum().backward()
