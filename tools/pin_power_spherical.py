#!/usr/bin/env python
"""Pin the PowerSpherical restatement the day the `power_spherical` package is importable (companion of tools/pin_geoopt.py).

nicola-decao/power_spherical is un-vendored and unpinned upstream (.gitignore:22 of the reference; imported at
models/sts/vae.py:7) and absent from this image, so oracle/power_spherical.py restates the pieces COSKAD uses from the published
source: the Beta parameters of the marginal t, the Householder sampling transform of `rsample`, `entropy`, the entropy of
HypersphericalUniform and the registered KL.  This script imports the REAL package, evaluates those on seeded inputs, diffs them
against the restatement and exits 0 (identical within the stated tolerance) or 1; without the package it exits 2 and says so.

    python tools/pin_power_spherical.py [--rtol 1e-6 --atol 1e-7]
"""
from __future__ import annotations

import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument('--rtol', type=float, default=1e-6)
    ap.add_argument('--atol', type=float, default=1e-7)
    a = ap.parse_args()
    try:
        from power_spherical.distributions import HypersphericalUniform, PowerSpherical
    except Exception as exc:
        print(f'power_spherical is not importable here ({type(exc).__name__}: {exc}); the restatement stays UNPINNED')
        return 2
    from oracle import power_spherical as ours
    g = torch.Generator().manual_seed(11)
    d = 8                                                         # latent_dim of config/UBnormal/spherical_vae.yaml:35
    loc = torch.nn.functional.normalize(torch.randn(512, d, generator=g), dim=-1)
    loc[0] = torch.eye(d)[0]                                      # loc == e1: the reflection degenerates
    kappa = torch.rand(512, generator=g) * 60 + 1                 # softplus(.) + 1 >= 1 (models/sts/vae.py:85)
    q = PowerSpherical(loc, kappa)
    p = HypersphericalUniform(d - 1)
    bad = 0

    def check(name, got, ref):
        nonlocal bad
        got, ref = torch.as_tensor(got, dtype=torch.float64), torch.as_tensor(ref, dtype=torch.float64)
        ok = torch.allclose(got, ref, rtol=a.rtol, atol=a.atol)
        print(f'{"ok  " if ok else "DIFF"} {name}: max abs diff {float((got - ref).abs().max()):.3e}')
        bad += 0 if ok else 1

    def attempt(name, fn):
        # the attribute layout of the installed package version is not known here: a missing attribute is reported, not fatal
        try:
            got, ref = fn()
            check(name, got, ref)
        except AttributeError as exc:
            print(f'skip {name}: {exc}')

    al, be = ours.ps_alpha_beta(kappa, d)
    attempt('marginal Beta alpha', lambda: (al, q.base_dist.marginal_t.base_dist.concentration1))
    attempt('marginal Beta beta', lambda: (be, q.base_dist.marginal_t.base_dist.concentration0))
    attempt('entropy', lambda: (ours.ps_entropy(kappa, d), q.entropy()))
    attempt('uniform entropy', lambda: (ours.hu_entropy(d), p.entropy()))
    attempt('KL(PS || U)', lambda: (ours.kl_ps_uniform(kappa, d), torch.distributions.kl.kl_divergence(q, p)))
    # the sampling transform on the package's own noise: t from the marginal, v uniform on S^{d-2}
    torch.manual_seed(5)
    t, v = ours.draw_noise(kappa, d, generator=torch.Generator().manual_seed(5))
    y = torch.cat((t.unsqueeze(-1), v * torch.sqrt(torch.clamp(1 - t.unsqueeze(-1) ** 2, 1e-7))), -1)
    real_z = q.transforms[-1](y) if hasattr(q, 'transforms') and q.transforms else None
    if real_z is not None:
        check('Householder transform of (t, v)', ours.rsample_from_noise(loc, t, v), real_z)
    else:
        print('skip Householder transform: the installed package exposes no transform list')
    print('PINNED' if bad == 0 else f'{bad} function(s) differ from the real power_spherical')
    return 0 if bad == 0 else 1


if __name__ == '__main__':
    sys.exit(main())
