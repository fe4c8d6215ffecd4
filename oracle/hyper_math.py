"""Oracle (TEST INFRASTRUCTURE): the ``utils/hyper_math.py`` flavour of the Poincare-ball
arithmetic (curvature convention c = +1, hyptorch constants).

The reference never imports utils/hyper_math.py (dead code), but BASELINE.json's north star
names it, and -- unlike geoopt -- it IS in the tree, so this flavour is PINNED:
``oracle/gen_golden.py`` runs the real module and commits tests/golden/geometry_hyper_math.npz.
It differs from geoopt only in constants (norm clamp 1e-5, artanh clamp 1e-5, +1e-5 on the
Mobius denominator, project eps 1e-3) and in the mean (Klein-model Einstein midpoint).
"""
from __future__ import annotations

import torch

Tensor = torch.Tensor


def tanh(x: Tensor, clamp: float = 15) -> Tensor:                  # utils/hyper_math.py:13-14
    return x.clamp(-clamp, clamp).tanh()


def artanh(x: Tensor) -> Tensor:                                   # :18-24
    x = x.clamp(-1 + 1e-5, 1 - 1e-5)
    return (torch.log(1 + x).sub(torch.log(1 - x))).mul(0.5)


def project(x: Tensor, c: float = 1.0) -> Tensor:                  # :100-105
    norm = torch.clamp_min(x.norm(dim=-1, keepdim=True, p=2), 1e-5)
    maxnorm = (1 - 1e-3) / (c ** 0.5)
    return torch.where(norm > maxnorm, x / norm * maxnorm, x)


def mobius_add(x: Tensor, y: Tensor, c: float = 1.0) -> Tensor:    # :173-179
    x2 = x.pow(2).sum(dim=-1, keepdim=True)
    y2 = y.pow(2).sum(dim=-1, keepdim=True)
    xy = (x * y).sum(dim=-1, keepdim=True)
    num = (1 + 2 * c * xy + c * y2) * x + (1 - c * x2) * y
    denom = 1 + 2 * c * xy + c ** 2 * x2 * y2
    return num / (denom + 1e-5)


def dist(x: Tensor, y: Tensor, c: float = 1.0) -> Tensor:          # :207-210
    sqrt_c = c ** 0.5
    return artanh(sqrt_c * mobius_add(-x, y, c).norm(dim=-1, p=2)) * 2 / sqrt_c


def expmap0(u: Tensor, c: float = 1.0) -> Tensor:                  # :302-306
    sqrt_c = c ** 0.5
    u_norm = torch.clamp_min(u.norm(dim=-1, p=2, keepdim=True), 1e-5)
    return tanh(sqrt_c * u_norm) * u / (sqrt_c * u_norm)


def poincare_mean(x: Tensor, dim: int = 0, c: float = 1.0) -> Tensor:   # :440-477
    xk = 2 * x / (1 + c * x.pow(2).sum(-1, keepdim=True))                # p2k
    lamb = 1 / torch.sqrt(1 - c * xk.pow(2).sum(dim=-1, keepdim=True))   # lorenz_factor
    mean = torch.sum(lamb * xk, dim=dim, keepdim=True) / torch.sum(lamb, dim=dim, keepdim=True)
    mean = mean / (1 + torch.sqrt(1 - c * mean.pow(2).sum(-1, keepdim=True)))   # k2p
    return mean.squeeze(dim)
