"""Synthetic inputs and weights for benchmarks and smoke runs (no datasets or checkpoints offline).

Shapes and distributions follow SURVEY.md 8(d): UBnormal-shape windows are robust-scaled coordinates
~N(0, 0.4^2) clamped to +-3 with 2 % of the (t, v) joints zeroed in both coordinates (missing joints,
utils/data.py:374-383); STC-shape windows are per-window centres U(-1,1) + sigma 0.1 clamped to [-1,1]
(utils/dataset_utils.py:36-42).  BatchNorm running statistics are randomised so that the eval-mode
fold is non-trivial.
"""
from __future__ import annotations

import torch


def synth_windows_(x: torch.Tensor, generator: torch.Generator, shape: str = 'ubnormal', chunk: int = 1 << 20) -> torch.Tensor:
    """fill x [N,2,12,17] in place, on its own device, chunk by chunk"""
    N = x.shape[0]
    for lo in range(0, N, chunk):
        xi = x[lo:lo + chunk]
        if shape == 'ubnormal':
            xi.normal_(0.0, 0.4, generator=generator).clamp_(-3, 3)
            drop = torch.rand((xi.shape[0], 1) + tuple(xi.shape[2:]), device=x.device, generator=generator) < 0.02
            xi.masked_fill_(drop, 0.0)
        elif shape == 'stc':
            ctr = torch.rand((xi.shape[0], xi.shape[1], 1, 1), device=x.device, generator=generator) * 2 - 1
            xi.normal_(0.0, 0.1, generator=generator).add_(ctr).clamp_(-1, 1)
        else:
            raise ValueError(shape)
    return x


@torch.no_grad()
def randomize_bn_(model: torch.nn.Module, seed: int = 0) -> torch.nn.Module:
    g = torch.Generator().manual_seed(seed)
    for m in model.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.weight.copy_(torch.rand(m.weight.shape, generator=g) + 0.5)
            m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)
            m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.2)
            m.running_var.copy_(torch.rand(m.running_var.shape, generator=g) + 0.5)
    return model


def make_model(kind: str = 'stse', latent_dim: int = 16, seed: int = 0, device='cuda'):
    """random-init network of the reference architecture (config/UBnormal/*.yaml shapes)"""
    from . import sts
    torch.manual_seed(seed)
    if kind == 'stsvae':
        from .spherical import STSVAE as cls
    else:
        cls = {'stse': sts.STSE, 'stsae': sts.STSAE}[kind]
    m = cls(input_dim=2, layer_channels=[32, 16, 32], hidden_dimension=64, latent_dim=latent_dim, n_frames=12,
            n_joints=17, encoder_type='sts_gcn', projector='linear', distance='euclidean', dropout=0.0)
    randomize_bn_(m, seed)
    return m.to(device).eval()
