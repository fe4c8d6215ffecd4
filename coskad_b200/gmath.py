"""``gmath``-compatible geometry namespace backed by the sm_100a kernels.

Mirrors the functions COSKAD calls on ``geoopt.manifolds.stereographic.math`` with ``k = -1``
(models/hyperbolic_encoder.py:110,122,147,157,179,181,266; utils/eval_utils.py:67;
eval_COSKAD.py:195): ``expmap0, project, dist, dist0, weighted_midpoint``; plus the
``utils/hyper_math.py`` flavour (``hm_*``), the sharded center update and the fused, differentiable
training score ``poincare_score``.  The element-wise ops are differentiable (analytic
vector-Jacobian kernels), so the reference's own ``training_step`` runs on them unchanged.  Tensors must live on a CUDA
device; there is no CPU path.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib


def _k_is_minus_one(k) -> None:
    kv = float(k) if not isinstance(k, (int, float)) else k
    if abs(kv + 1.0) > 1e-12:
        raise NotImplementedError(f'coskad_b200.gmath implements curvature k = -1 (the only value COSKAD uses), got {kv}')


def _prep(x: torch.Tensor, what: str, dim: int = -1):
    if not isinstance(x, torch.Tensor):
        raise TypeError(f'{what} must be a tensor')
    if not x.is_cuda:
        raise _lib.CoskadError(f'{what} is not on a CUDA device: coskad_b200.gmath has no CPU fallback')
    if dim not in (-1, x.dim() - 1):
        raise NotImplementedError('only dim=-1 is implemented')
    x2 = x.detach().to(torch.float32).contiguous()
    D = x2.shape[-1]
    return x2.view(-1, D), D


def _needs_grad(*ts) -> bool:
    return torch.is_grad_enabled() and any(torch.is_tensor(t) and t.requires_grad for t in ts)


def _ctx(x: torch.Tensor) -> _lib.Context:
    return _lib.context(x.device.index if x.device.index is not None else torch.cuda.current_device())


def _map_fwd(x: torch.Tensor, op: int, what: str, dim: int = -1) -> torch.Tensor:
    x2, D = _prep(x, what, dim)
    out = torch.empty_like(x2)
    ctx = _ctx(x2)
    ctx.check(ctx.lib.coskad_geom_map(ctx.h, op, x2.data_ptr(), x2.shape[0], D, out.data_ptr(), _lib.stream_ptr(x2.device)),
              'coskad_geom_map')
    return out.view(x.shape)


_MAP_HAS_BWD = (_lib.MAP_EXPMAP0, _lib.MAP_PROJECT, _lib.MAP_EXPMAP0_PROJECT, _lib.MAP_L2NORMALIZE)


class _MapFn(torch.autograd.Function):
    """an element-wise gmath map with its analytic vector-Jacobian product (coskad_geom_map_bwd): what autograd does through
    geoopt at models/hyperbolic_encoder.py:147"""

    @staticmethod
    def forward(ctx, x, op, what):
        ctx.save_for_backward(x.detach())
        ctx.op = op
        return _map_fwd(x, op, what)

    @staticmethod
    def backward(ctx, gout):
        (x,) = ctx.saved_tensors
        x2, D = _prep(x, 'x')
        g2 = gout.detach().to(torch.float32).contiguous().view(-1, D)
        gin = torch.empty_like(x2)
        c = _ctx(x2)
        c.check(c.lib.coskad_geom_map_bwd(c.h, ctx.op, x2.data_ptr(), g2.data_ptr(), x2.shape[0], D, gin.data_ptr(),
                                          _lib.stream_ptr(x2.device)), 'coskad_geom_map_bwd')
        return gin.view(x.shape), None, None


def _map(x: torch.Tensor, op: int, what: str, dim: int = -1) -> torch.Tensor:
    if _needs_grad(x):
        if op not in _MAP_HAS_BWD:
            raise NotImplementedError('the utils/hyper_math.py flavour is forward-only (a pinned cross-check, not a training path)')
        if dim not in (-1, x.dim() - 1):
            raise NotImplementedError('only dim=-1 is implemented')
        return _MapFn.apply(x, op, what)
    return _map_fwd(x, op, what, dim)


def expmap0(u: torch.Tensor, *, k=-1.0, dim: int = -1) -> torch.Tensor:
    _k_is_minus_one(k)
    return _map(u, _lib.MAP_EXPMAP0, 'u', dim)


def project(x: torch.Tensor, *, k=-1.0, dim: int = -1, eps: float = -1.0) -> torch.Tensor:
    _k_is_minus_one(k)
    if eps >= 0 and abs(eps - 4e-3) > 1e-12:
        raise NotImplementedError('project: only the float32 default eps = 4e-3 is implemented')
    return _map(x, _lib.MAP_PROJECT, 'x', dim)


def expmap0_project(u: torch.Tensor, *, k=-1.0) -> torch.Tensor:
    """project(expmap0(u)) in one launch (models/hyperbolic_encoder.py:110,147)."""
    _k_is_minus_one(k)
    return _map(u, _lib.MAP_EXPMAP0_PROJECT, 'u')


def hm_expmap0(u: torch.Tensor, c: float = 1.0) -> torch.Tensor:     # utils/hyper_math.py:302-306
    assert c == 1.0
    return _map(u, _lib.MAP_EXPMAP0_HM, 'u')


def hm_project(x: torch.Tensor, c: float = 1.0) -> torch.Tensor:     # utils/hyper_math.py:100-105
    assert c == 1.0
    return _map(x, _lib.MAP_PROJECT_HM, 'x')


def l2_normalize(z: torch.Tensor) -> torch.Tensor:                   # models/sts/vae.py:81
    return _map(z, _lib.MAP_L2NORMALIZE, 'z')


def _pair_fwd(flavour: int, a: torch.Tensor, b: torch.Tensor, keepdim: bool = False) -> torch.Tensor:
    # broadcast the 1-D operand like torch does; the kernel computes f(a_row, b_row)
    if a.dim() == 1 and b.dim() > 1:
        a = a.expand_as(b)
    a2, D = _prep(a, 'x')
    if b.dim() == 1 or b.numel() == D:
        b2, bc = b.detach().to(device=a2.device, dtype=torch.float32).contiguous().view(-1), 1
        if b2.numel() != D:
            raise ValueError('dimension mismatch')
    else:
        b2, D2 = _prep(b, 'y')
        bc = 0
        if b2.shape != a2.shape:
            raise ValueError(f'shape mismatch {tuple(a.shape)} vs {tuple(b.shape)}')
    out = torch.empty(a2.shape[0], device=a2.device, dtype=torch.float32)
    ctx = _ctx(a2)
    ctx.check(ctx.lib.coskad_dist(ctx.h, flavour, a2.data_ptr(), b2.data_ptr(), bc, a2.shape[0], D, out.data_ptr(),
                                  _lib.stream_ptr(a2.device)), 'coskad_dist')
    out = out.view(a.shape[:-1])
    return out.unsqueeze(-1) if keepdim else out


_PAIR_HAS_BWD = (_lib.SCORE_POINCARE, _lib.SCORE_POINCARE_NOPROJ, _lib.SCORE_EUCLID, _lib.SCORE_COSINE)


class _PairFn(torch.autograd.Function):
    """f(a, b) per row with gradients to both operands (coskad_dist_bwd); a [D] operand is broadcast and its gradient is the
    sum of the per-row gradients.  What autograd does through gmath.dist at models/hyperbolic_encoder.py:157."""

    @staticmethod
    def forward(ctx, a, b, flavour, keepdim):
        ctx.save_for_backward(a.detach(), b.detach())
        ctx.flavour, ctx.keepdim = flavour, keepdim
        return _pair_fwd(flavour, a, b, keepdim)

    @staticmethod
    def backward(ctx, gs):
        a, b = ctx.saved_tensors
        rows = b if a.dim() == 1 and b.dim() > 1 else a
        D = rows.shape[-1]
        a2 = a.to(device=rows.device, dtype=torch.float32).expand_as(rows).contiguous().view(-1, D)
        bc = int(b.dim() == 1 or b.numel() == D) if rows is a else 0
        b2 = b.to(device=rows.device, dtype=torch.float32).contiguous().view(-1) if bc else \
            b.to(torch.float32).contiguous().view(-1, D)
        g2 = gs.detach().to(torch.float32).contiguous().view(-1)
        need_a, need_b = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        ga = torch.empty_like(a2) if need_a else None
        gb = torch.empty_like(a2) if need_b else None
        c = _ctx(a2)
        c.check(c.lib.coskad_dist_bwd(c.h, ctx.flavour, a2.data_ptr(), b2.data_ptr(), bc, g2.data_ptr(), a2.shape[0], D,
                                      _lib._ptr(ga), _lib._ptr(gb), _lib.stream_ptr(a2.device)), 'coskad_dist_bwd')

        def shape_like(g, t):
            if g is None:
                return None
            g = g.view(rows.shape)
            return g.reshape(-1, D).sum(0).view(t.shape) if t.numel() == D and rows.numel() != D else g.view(t.shape)
        return shape_like(ga, a), shape_like(gb, b), None, None


def _pair(flavour: int, a: torch.Tensor, b: torch.Tensor, keepdim: bool = False) -> torch.Tensor:
    if _needs_grad(a, b):
        if flavour not in _PAIR_HAS_BWD:
            raise NotImplementedError('the utils/hyper_math.py flavour is forward-only (a pinned cross-check, not a training path)')
        return _PairFn.apply(a, b.to(a.device) if b.device != a.device else b, flavour, keepdim)
    return _pair_fwd(flavour, a, b, keepdim)


def dist(x: torch.Tensor, y: torch.Tensor, *, k=-1.0, keepdim: bool = False, dim: int = -1) -> torch.Tensor:
    """2 artanh(|(-x) (+) y|); either operand may be a single point [D] (the center)."""
    _k_is_minus_one(k)
    if x.dim() == 1 and y.dim() > 1 and x.device != y.device:
        x = x.to(y.device)
    return _pair(_lib.SCORE_POINCARE, x, y, keepdim)


def hm_dist(x: torch.Tensor, y: torch.Tensor, c: float = 1.0) -> torch.Tensor:   # utils/hyper_math.py:207-210
    assert c == 1.0
    return _pair(_lib.SCORE_POINCARE_HM, x, y)


def euclid_score(z: torch.Tensor, c: torch.Tensor) -> torch.Tensor:
    """mean_d (c_d - z_d)^2   (utils/eval_utils.py:61-64)"""
    return _pair(_lib.SCORE_EUCLID, z, c)


def cosine_score(z: torch.Tensor, c: torch.Tensor) -> torch.Tensor:
    """1 - cos(c, z)   (eval_COSKAD.py:81)"""
    return _pair(_lib.SCORE_COSINE, z, c)


def dist0(x: torch.Tensor, *, k=-1.0, keepdim: bool = False, dim: int = -1) -> torch.Tensor:
    _k_is_minus_one(k)
    x2, D = _prep(x, 'x', dim)
    out = torch.empty(x2.shape[0], device=x2.device, dtype=torch.float32)
    ctx = _ctx(x2)
    ctx.check(ctx.lib.coskad_dist0(ctx.h, x2.data_ptr(), x2.shape[0], D, out.data_ptr(), _lib.stream_ptr(x2.device)),
              'coskad_dist0')
    out = out.view(x.shape[:-1])
    return out.unsqueeze(-1) if keepdim else out


# ---- Mahalanobis distance to the center (distance: 'mahalanobis') -----------------------------
class _MahalanobisFn(torch.autograd.Function):
    """sqrt((u - c)^T VI (u - c)) per row, differentiable w.r.t. u (the loss of models/euclidean_encoder_staticCenter.py:185)"""

    @staticmethod
    def forward(ctx, u, c, VI):
        u2, D = _prep(u, 'u')
        c2 = c.detach().to(device=u2.device, dtype=torch.float32).contiguous().view(-1)
        vi = VI.detach().to(device=u2.device, dtype=torch.float32).contiguous()
        if c2.numel() != D or tuple(vi.shape) != (D, D):
            raise _lib.CoskadError(f'mahalanobis: u [.., {D}] needs a center [{D}] and VI [{D}, {D}], got {tuple(c.shape)} / {tuple(VI.shape)}')
        out = torch.empty(u2.shape[0], device=u2.device, dtype=torch.float32)
        cx = _ctx(u2)
        cx.check(cx.lib.coskad_mahalanobis(cx.h, u2.data_ptr(), c2.data_ptr(), vi.data_ptr(), u2.shape[0], D, out.data_ptr(),
                                           _lib.stream_ptr(u2.device)), 'coskad_mahalanobis')
        ctx.save_for_backward(u2, c2, vi)
        ctx.shape = u.shape
        return out

    @staticmethod
    def backward(ctx, gs):
        u2, c2, vi = ctx.saved_tensors
        gs = gs.detach().to(torch.float32).contiguous()
        gz = torch.empty_like(u2)
        cx = _ctx(u2)
        cx.check(cx.lib.coskad_mahalanobis_bwd(cx.h, u2.data_ptr(), c2.data_ptr(), vi.data_ptr(), gs.data_ptr(), u2.shape[0],
                                               u2.shape[1], gz.data_ptr(), _lib.stream_ptr(u2.device)), 'coskad_mahalanobis_bwd')
        return gz.view(ctx.shape), None, None


def mahalanobis_score(u: torch.Tensor, c: torch.Tensor, VI: torch.Tensor) -> torch.Tensor:
    """per-row distance [B]: ``mahalanobis(u, c, VI, reduce='none')`` of utils/eval_utils.py:28-38 without the two trailing
    singleton dimensions"""
    return _MahalanobisFn.apply(u, c, VI)


def mahalanobis(u: torch.Tensor, v: torch.Tensor, VI: torch.Tensor, reduce: str = 'mean') -> torch.Tensor:
    """utils/eval_utils.py:28-38, same signature: ``v`` is the center [D] (or [1, D]); returns the mean distance, or the
    per-row distances shaped [B, 1, 1] like the reference's matmul result for ``reduce != 'mean'``"""
    d = mahalanobis_score(u, v.reshape(-1), VI)
    return d.mean() if reduce == 'mean' else d.view(-1, 1, 1)


def cov_accumulator(D: int, device) -> torch.Tensor:
    """zeroed [D*D + 1] float64 accumulator: the scatter matrix (row-major) and the row count"""
    return torch.zeros(D * D + 1, dtype=torch.float64, device=device)


def cov_partial(z: torch.Tensor, mu: torch.Tensor, acc: torch.Tensor) -> None:
    """acc += sum_b (z_b - mu)(z_b - mu)^T  (batch_cov_mat_step, models/euclidean_encoder_staticCenter.py:40-46); partial sums
    of batches and ranks add"""
    z2, D = _prep(z, 'z')
    mu2 = mu.detach().to(device=z2.device, dtype=torch.float32).contiguous().view(-1)
    assert acc.dtype == torch.float64 and acc.numel() == D * D + 1 and acc.is_cuda and mu2.numel() == D
    cx = _ctx(z2)
    cx.check(cx.lib.coskad_cov_partial(cx.h, z2.data_ptr(), mu2.data_ptr(), z2.shape[0], D, acc.data_ptr(),
                                       _lib.stream_ptr(z2.device)), 'coskad_cov_partial')


def inv_cov_finalize(acc: torch.Tensor, D: int) -> torch.Tensor:
    """inverse of the sample covariance ``scatter / (n - 1)`` (compute_inv_cov_mat, models/euclidean_encoder_staticCenter.py:133-142):
    a D x D inverse once per epoch -- parameter preparation, done by torch.linalg in float32 like the reference"""
    n = acc[D * D]
    cov = (acc[:D * D] / (n - 1.0)).to(torch.float32).view(D, D)
    return torch.inverse(cov)


# ---- center update: per-shard partial sums (float64) -> [all-reduce] -> finalize --------------
def center_accumulator(D: int, device) -> torch.Tensor:
    """zeroed [D+2] float64 accumulator: D sums, the gamma-1 sum, the window count"""
    return torch.zeros(D + 2, dtype=torch.float64, device=device)


def center_partial(z: torch.Tensor, acc: torch.Tensor, flavour: int = _lib.SCORE_POINCARE) -> None:
    z2, D = _prep(z, 'z')
    assert acc.dtype == torch.float64 and acc.numel() == D + 2 and acc.is_cuda
    ctx = _ctx(z2)
    ctx.check(ctx.lib.coskad_center_partial(ctx.h, flavour, z2.data_ptr(), z2.shape[0], D, acc.data_ptr(),
                                            _lib.stream_ptr(z2.device)), 'coskad_center_partial')


def center_finalize(acc: torch.Tensor, D: int, flavour: int = _lib.SCORE_POINCARE, eps: float = 0.0) -> torch.Tensor:
    out = torch.empty(D, dtype=torch.float32, device=acc.device)
    ctx = _ctx(acc)
    ctx.check(ctx.lib.coskad_center_finalize(ctx.h, flavour, acc.data_ptr(), D, float(eps), out.data_ptr(),
                                             _lib.stream_ptr(acc.device)), 'coskad_center_finalize')
    return out


def weighted_midpoint(xs: torch.Tensor, *, k=-1.0, dim: int = -1, **kw) -> torch.Tensor:
    """gyro-midpoint of the rows of xs [N, D] (weights=None form, models/hyperbolic_encoder.py:122,179)"""
    _k_is_minus_one(k)
    if kw.get('weights') is not None:
        raise NotImplementedError('weighted_midpoint: only weights=None is used by COSKAD')
    x2, D = _prep(xs, 'xs', dim)
    acc = center_accumulator(D, x2.device)
    center_partial(x2, acc, _lib.SCORE_POINCARE)
    return center_finalize(acc, D, _lib.SCORE_POINCARE)


# ---- fused differentiable training score -------------------------------------------------------
def poincare_score_bwd(z: torch.Tensor, c: torch.Tensor, dscore: torch.Tensor, with_project: bool = True) -> torch.Tensor:
    z2, D = _prep(z, 'z')
    c2 = c.detach().to(device=z2.device, dtype=torch.float32).contiguous().view(-1)
    g2 = dscore.detach().to(device=z2.device, dtype=torch.float32).contiguous().view(-1)
    dz = torch.empty_like(z2)
    ctx = _ctx(z2)
    ctx.check(ctx.lib.coskad_poincare_score_bwd(ctx.h, z2.data_ptr(), c2.data_ptr(), g2.data_ptr(), z2.shape[0], D,
                                                int(with_project), dz.data_ptr(), _lib.stream_ptr(z2.device)),
              'coskad_poincare_score_bwd')
    return dz.view(z.shape)


class _PoincareScore(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, c, with_project):
        zd = z.detach()
        x = expmap0_project(zd) if with_project else expmap0(zd)
        ctx.save_for_backward(zd, c.detach())
        ctx.with_project = with_project
        ctx.mark_non_differentiable(x)
        return dist(c.detach().to(zd.device).expand_as(x).contiguous(), x), x

    @staticmethod
    def backward(ctx, dscore, _dx):
        z, c = ctx.saved_tensors
        return poincare_score_bwd(z, c, dscore.contiguous(), ctx.with_project), None, None


def poincare_score(z: torch.Tensor, c: torch.Tensor, with_project: bool = True):
    """(dist(c, x), x) with x = project(expmap0(z)); differentiable w.r.t. z (the center is a constant,
    models/hyperbolic_encoder.py:147-157)."""
    return _PoincareScore.apply(z, c, with_project)
