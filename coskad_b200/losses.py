"""Training losses of the COSKAD task modules (host glue over the CUDA ops).

calc_reg_loss mirrors utils/model_utils.py:90-105: 0.5 * sum ||p||^2 over the tensors whose name does
not contain 'bias', divided by the NUMBER of such tensors (it includes BatchNorm weights, PReLU slopes,
A and T).  The reference walks the ~42 tensors one by one (three tiny kernels each, and as many again in
the backward); here the value comes from one multi-tensor norm and the gradient (alpha/n * p per tensor)
from one multi-tensor scale -- same value up to fp32 summation order.
"""
from __future__ import annotations

from typing import List

import torch


class _L2Reg(torch.autograd.Function):
    """0.5 * sum_i ||p_i||_2^2 / n   (utils/model_utils.py:92-102: first tensor 0.5*sum(p^2), the rest 0.5*||p||_2^2)"""

    @staticmethod
    def forward(ctx, n_avg: float, *params: torch.Tensor) -> torch.Tensor:
        ctx.n_avg = n_avg
        ctx.params = params                       # the leaves themselves (backward looks at their .grad)
        ctx.save_for_backward(*params)
        norms = torch._foreach_norm([p.detach() for p in params], 2)
        return 0.5 * torch.stack(norms).square().sum() / n_avg

    @staticmethod
    def backward(ctx, g: torch.Tensor):
        params = ctx.saved_tensors
        scale = g / ctx.n_avg
        grads = torch._foreach_mul([p.detach() for p in params], scale)
        # parameters whose .grad is a view of an attached dist.FlatGradBucket: one multi-tensor add into the bucket instead
        # of one AccumulateGrad add_ launch per tensor (42 of them per step); the others go back to autograd
        direct = [getattr(p, 'coskad_direct_grad', False) and p.grad is not None for p in ctx.params]
        if any(direct):
            with torch.no_grad():
                torch._foreach_add_([p.grad for p, d in zip(ctx.params, direct) if d], [g_ for g_, d in zip(grads, direct) if d])
        return (None, *[None if d else g_ for g_, d in zip(grads, direct)])


def calc_reg_loss(model, reg_type: str = 'l2', avg: bool = True):
    parameters: List[torch.Tensor] = [param for name, param in model.named_parameters() if 'bias' not in name]
    if reg_type.lower() == 'l2':
        return _L2Reg.apply(float(len(parameters)) if avg else 1.0, *parameters)
    return torch.tensor(0.0, device=next(model.parameters()).device)
