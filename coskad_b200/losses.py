"""Training losses of the COSKAD task modules (host glue over the CUDA ops).

calc_reg_loss mirrors utils/model_utils.py:90-105: 0.5 * sum ||p||^2 over the tensors whose name does
not contain 'bias', divided by the NUMBER of such tensors (it includes BatchNorm weights, PReLU slopes,
A and T).  It is a reduction over 240 k parameters per step -- kept in PyTorch (tiny).
"""
from __future__ import annotations

import torch


def calc_reg_loss(model, reg_type: str = 'l2', avg: bool = True):
    reg_loss = None
    parameters = list(param for name, param in model.named_parameters() if 'bias' not in name)
    num_params = len(parameters)
    if reg_type.lower() == 'l2':
        for param in parameters:
            if reg_loss is None:
                reg_loss = 0.5 * torch.sum(param ** 2)
            else:
                reg_loss = reg_loss + 0.5 * param.norm(2) ** 2
        if avg:
            reg_loss /= num_params
        return reg_loss
    return torch.tensor(0.0, device=next(model.parameters()).device)
