"""Oracle (TEST INFRASTRUCTURE): restatement of the parts of nicola-decao/power_spherical that COSKAD uses
(models/sts/vae.py:7,110-111,129; KL at models/spherical_vae.py:92).

PARITY UNPINNED: power_spherical is an un-vendored, un-pinned third-party package (reference .gitignore:22);
it cannot be installed offline and the reference has no fixtures for it.  Formulas follow the published source
(SURVEY.md appendix B).  Sampling is expressed as a deterministic function of explicit noise
(t ~ 2 Beta(alpha, beta) - 1, v ~ uniform on S^{d-2}) so that the CUDA kernel can be compared exactly.
"""
from __future__ import annotations

import math

import torch

Tensor = torch.Tensor
_EPS = 1e-7


def ps_alpha_beta(kappa: Tensor, d: int):
    return (d - 1) / 2 + kappa, torch.full_like(kappa, (d - 1) / 2)


def t_transform(t: Tensor, v: Tensor) -> Tensor:
    """_TTransform: y = [t, sqrt(clamp(1 - t^2, eps)) * v]   (t [B,1], v [B,d-1] unit vectors)"""
    return torch.cat((t, v * torch.sqrt(torch.clamp(1 - t ** 2, _EPS))), -1)


def householder(y: Tensor, loc: Tensor) -> Tensor:
    """_HouseholderRotationTransform: reflect e1 onto loc"""
    u = torch.zeros_like(loc)
    u[..., 0] = 1.0
    u = u - loc
    u = u / (u.norm(dim=-1, keepdim=True) + 1e-5)
    return y - 2 * (y * u).sum(-1, keepdim=True) * u


def rsample_from_noise(loc: Tensor, t: Tensor, v: Tensor) -> Tensor:
    return householder(t_transform(t, v), loc)


def draw_noise(kappa: Tensor, d: int, generator=None):
    """t = 2 Beta(alpha, beta) - 1 [B,1];  v = normalised Gaussian [B,d-1]"""
    a, b = ps_alpha_beta(kappa, d)
    t = 2 * torch.distributions.Beta(a, b).sample() - 1
    g = torch.randn(kappa.shape + (d - 1,), generator=generator, dtype=kappa.dtype)
    return t.unsqueeze(-1), g / g.norm(dim=-1, keepdim=True)


def ps_entropy(kappa: Tensor, d: int) -> Tensor:
    a, b = ps_alpha_beta(kappa, d)
    log_norm = -((a + b) * math.log(2) + torch.lgamma(a) - torch.lgamma(a + b) + b * math.log(math.pi))
    return -(log_norm + kappa * (math.log(2) + torch.digamma(a) - torch.digamma(a + b)))


def hu_entropy(d: int) -> float:
    """HypersphericalUniform(dim = d - 1).entropy()"""
    return math.log(2) + (d / 2) * math.log(math.pi) - math.lgamma(d / 2)


def kl_ps_uniform(kappa: Tensor, d: int) -> Tensor:
    """KL(PowerSpherical(mu, kappa) || HypersphericalUniform) = -H(PS) + H(U)   (per sample)"""
    return -ps_entropy(kappa, d) + hu_entropy(d)
