"""Oracle (TEST INFRASTRUCTURE): restatement of geoopt==0.5.0
``geoopt/manifolds/stereographic/math.py`` for the functions COSKAD calls with k = -1.

PARITY UNPINNED: geoopt is a third-party dependency pinned in the reference's
environment.yml:247, it is not vendored under /root/reference and cannot be installed
offline, and the reference has no golden vectors for it.  The formulas and constants below
follow the published 0.5.0 source; the call sites that anchor them are
models/hyperbolic_encoder.py:110,122,147,157,179,181,266, utils/eval_utils.py:67 and
eval_COSKAD.py:195.  The independent, fully pinned cross-check is ``oracle.hyper_math``.

All functions take torch tensors; ``k`` is a 0-dim tensor (the reference passes
``torch.tensor(-1.)``, hyperbolic_encoder.py:70).  Only k < 0 is implemented.
"""
from __future__ import annotations

import torch

Tensor = torch.Tensor
MIN_NORM = 1e-15
PROJ_EPS = {torch.float32: 4e-3, torch.float64: 1e-5}


def _k(k, ref: Tensor) -> Tensor:
    k = torch.as_tensor(k, dtype=ref.dtype, device=ref.device)
    assert bool((k < 0).all()), 'oracle.geoopt_math restates the k<0 (Poincare ball) branch only'
    return k


def tanh(x: Tensor) -> Tensor:
    return x.clamp(-15, 15).tanh()


def artanh(x: Tensor) -> Tensor:
    x = x.clamp(-1 + 1e-7, 1 - 1e-7)
    return (torch.log(1 + x).sub(torch.log(1 - x))).mul(0.5)


def sabs(x: Tensor, eps: float = 1e-15) -> Tensor:
    return x.abs().add(eps)


def clamp_abs(x: Tensor, eps: float = 1e-15) -> Tensor:
    s = torch.sign(x)
    s = torch.where(s == 0, torch.ones_like(s), s)          # sign(0) := +1
    return s * sabs(x, eps)


def tan_k(x: Tensor, k: Tensor) -> Tensor:
    k_sqrt = sabs(k).sqrt()
    return k_sqrt.reciprocal() * tanh(x * k_sqrt)


def artan_k(x: Tensor, k: Tensor) -> Tensor:
    k_sqrt = sabs(k).sqrt()
    return k_sqrt.reciprocal() * artanh(x * k_sqrt)


def expmap0(u: Tensor, *, k, dim: int = -1) -> Tensor:
    k = _k(k, u)
    u_norm = u.norm(dim=dim, p=2, keepdim=True).clamp_min(MIN_NORM)
    return tan_k(u_norm, k) * (u / u_norm)


def project(x: Tensor, *, k, dim: int = -1, eps: float = -1.0) -> Tensor:
    k = _k(k, x)
    if eps < 0:
        eps = PROJ_EPS[x.dtype]
    maxnorm = (1 - eps) / (sabs(k) ** 0.5)
    norm = x.norm(dim=dim, keepdim=True, p=2).clamp_min(MIN_NORM)
    cond = norm > maxnorm
    projected = x / norm * maxnorm
    return torch.where(cond, projected, x)


def mobius_add(x: Tensor, y: Tensor, *, k, dim: int = -1) -> Tensor:
    k = _k(k, x)
    x2 = x.pow(2).sum(dim=dim, keepdim=True)
    y2 = y.pow(2).sum(dim=dim, keepdim=True)
    xy = (x * y).sum(dim=dim, keepdim=True)
    num = (1 - 2 * k * xy - k * y2) * x + (1 + k * x2) * y
    denom = 1 - 2 * k * xy + k ** 2 * x2 * y2
    return num / denom.clamp_min(MIN_NORM)


def dist(x: Tensor, y: Tensor, *, k, keepdim: bool = False, dim: int = -1) -> Tensor:
    k = _k(k, x)
    return 2.0 * artan_k(mobius_add(-x, y, k=k, dim=dim).norm(dim=dim, p=2, keepdim=keepdim), k)


def dist0(x: Tensor, *, k, keepdim: bool = False, dim: int = -1) -> Tensor:
    k = _k(k, x)
    return 2.0 * artan_k(x.norm(dim=dim, p=2, keepdim=keepdim), k)


def lambda_x(x: Tensor, *, k, keepdim: bool = False, dim: int = -1) -> Tensor:
    k = _k(k, x)
    return 2 / (1 + k * x.pow(2).sum(dim=dim, keepdim=keepdim)).clamp_min(MIN_NORM)


def mobius_scalar_mul(r, x: Tensor, *, k, dim: int = -1) -> Tensor:
    k = _k(k, x)
    x_norm = x.norm(dim=dim, keepdim=True, p=2).clamp_min(MIN_NORM)
    return tan_k(r * artan_k(x_norm, k), k) * (x / x_norm)


def weighted_midpoint(xs: Tensor, *, k, dim: int = -1) -> Tensor:
    """weights=None, reducedim = every dim but ``dim``, keepdim=False (the only form COSKAD uses:
    hyperbolic_encoder.py:122,179 on a [N,D] tensor)."""
    k = _k(k, xs)
    reducedim = [d for d in range(xs.dim()) if d != (dim % xs.dim())]
    gamma = lambda_x(xs, k=k, dim=dim, keepdim=True)
    denominator = (gamma - 1).sum(reducedim, keepdim=True)
    nominator = (gamma * xs).sum(reducedim, keepdim=True)
    two_mean = nominator / clamp_abs(denominator, 1e-10)
    a_mean = mobius_scalar_mul(torch.tensor(0.5, dtype=xs.dtype), two_mean, k=k, dim=dim)
    for d in sorted(reducedim, reverse=True):
        a_mean = a_mean.squeeze(d)
    return a_mean
