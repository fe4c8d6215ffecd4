#!/usr/bin/env python
"""Pin the geoopt restatement the day a geoopt wheel exists.

geoopt==0.5.0 (environment.yml:247 of the reference) is absent from /root/reference and from this image, so
oracle/geoopt_math.py restates `geoopt/manifolds/stereographic/math.py` from its published source and the parity of the
hyperbolic tail is "unpinned".  This script is the one command that lifts the cap: it imports the REAL
`geoopt.manifolds.stereographic.math`, evaluates every function COSKAD calls (expmap0, project, dist in both argument
orders, dist0, weighted_midpoint, lambda_x, mobius_add, mobius_scalar_mul) on the inputs of the committed golden file
tests/golden/geometry_geoopt_restated.npz plus boundary cases (norms near 0, at the projection radius, beyond the tanh /
artanh clamps), diffs them against the restatement AND against the golden outputs, and exits 0 (identical to the stated
tolerance) or 1.  Without geoopt it exits 2 and says so.

    python tools/pin_geoopt.py [--atol 0 --rtol 1e-7] [--write-golden]
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def cases(g):
    u = torch.from_numpy(g['u'])
    c = torch.from_numpy(g['center'])
    gen = torch.Generator().manual_seed(7)
    d = u.shape[-1]
    dirs = torch.nn.functional.normalize(torch.randn(64, d, generator=gen), dim=-1)
    radii = torch.tensor([0.0, 1e-20, 1e-15, 1e-8, 1e-3, 0.5, 0.99, 0.995, 0.996, 0.9961, 0.999, 1.0, 3.0, 14.9, 15.0, 15.1, 40.0])
    edge = (dirs[:, None, :] * radii[None, :, None]).reshape(-1, d)
    return torch.cat([u, edge]), c


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument('--rtol', type=float, default=1e-7)
    ap.add_argument('--atol', type=float, default=0.0)
    ap.add_argument('--write-golden', action='store_true', help='rewrite the golden file from the REAL geoopt outputs')
    a = ap.parse_args()
    try:
        import geoopt
        import geoopt.manifolds.stereographic.math as real
    except Exception as exc:
        print(f'geoopt is not importable here ({type(exc).__name__}: {exc}); the restatement stays UNPINNED')
        return 2
    from oracle import geoopt_math as ours
    path = os.path.join(ROOT, 'tests', 'golden', 'geometry_geoopt_restated.npz')
    g = np.load(path)
    u, c = cases(g)
    k = torch.tensor(-1.)
    bad = 0

    def check(name, got, ref):
        nonlocal bad
        ok = torch.allclose(got, ref, rtol=a.rtol, atol=a.atol, equal_nan=True)
        err = float((got - ref).abs().max())
        print(f'{"ok  " if ok else "DIFF"} {name:28s} max abs diff {err:.3e}')
        bad += 0 if ok else 1

    x_r = real.expmap0(u, k=k)
    check('expmap0', ours.expmap0(u, k=k), x_r)
    p_r = real.project(x_r, k=k)
    check('project', ours.project(x_r, k=k), p_r)
    check('dist(x, c)', ours.dist(p_r, c, k=k), real.dist(p_r, c, k=k))
    check('dist(c, x)', ours.dist(c, p_r, k=k), real.dist(c, p_r, k=k))
    check('dist0', ours.dist0(p_r, k=k), real.dist0(p_r, k=k))
    check('weighted_midpoint', ours.weighted_midpoint(p_r[: g['u'].shape[0]], k=k), real.weighted_midpoint(p_r[: g['u'].shape[0]], k=k))
    for fn in ('lambda_x', 'mobius_add', 'mobius_scalar_mul'):
        if hasattr(ours, fn) and hasattr(real, fn):
            if fn == 'lambda_x':
                check(fn, ours.lambda_x(p_r, k=k), real.lambda_x(p_r, k=k))
            elif fn == 'mobius_add':
                check(fn, ours.mobius_add(-p_r, c.expand_as(p_r), k=k), real.mobius_add(-p_r, c.expand_as(p_r), k=k))
            else:
                r = torch.tensor(0.5)
                check(fn, ours.mobius_scalar_mul(r, p_r, k=k), real.mobius_scalar_mul(r, p_r, k=k))
    n = g['u'].shape[0]
    for key, val in (('expmap0', x_r[:n]), ('project', p_r[:n]), ('dist', real.dist(p_r[:n], c, k=k)),
                     ('dist_cx', real.dist(c, p_r[:n], k=k)), ('dist0', real.dist0(p_r[:n], k=k)),
                     ('midpoint', real.weighted_midpoint(p_r[:n], k=k))):
        check(f'golden[{key}] vs real geoopt', torch.from_numpy(g[key]), val)
    if a.write_golden:
        out = {f: g[f] for f in g.files}
        out.update(expmap0=x_r[:n].numpy(), project=p_r[:n].numpy(), dist=real.dist(p_r[:n], c, k=k).numpy(),
                   dist_cx=real.dist(c, p_r[:n], k=k).numpy(), dist0=real.dist0(p_r[:n], k=k).numpy(),
                   midpoint=real.weighted_midpoint(p_r[:n], k=k).numpy())
        np.savez(path, **out)
        print(f'rewrote {path} from geoopt {getattr(geoopt, "__version__", "?")}')
    print(f'geoopt {getattr(geoopt, "__version__", "?")}: ' + ('restatement PINNED (all functions agree)' if bad == 0 else f'{bad} function(s) differ'))
    return 0 if bad == 0 else 1


if __name__ == '__main__':
    sys.exit(main())
