"""GPU parity of the geometry / center / aggregation kernels against the oracle and the fixtures."""
import os

import numpy as np
import pytest
import torch

from oracle import aggregate as oagg
from oracle import geoopt_math as ogm
from oracle import hyper_math as ohm

pytestmark = pytest.mark.gpu
K = torch.tensor(-1.)


def _close(got, ref, rtol=1e-4, atol=1e-6):
    got, ref = got.detach().cpu().double(), torch.as_tensor(ref).double()
    assert got.shape == ref.shape, (got.shape, ref.shape)
    err = (got - ref).abs()
    assert bool((err <= rtol * ref.abs() + atol).all()), f'max abs err {float(err.max()):.3e}'


def test_gmath_fixture_geoopt(golden_dir):
    from coskad_b200 import gmath
    g = np.load(os.path.join(golden_dir, 'geometry_geoopt_restated.npz'))
    u, c = torch.from_numpy(g['u']).cuda(), torch.from_numpy(g['center']).cuda()
    e = gmath.expmap0(u, k=K)
    p = gmath.project(e, k=K)
    _close(e, g['expmap0'], 1e-5, 1e-8)
    _close(p, g['project'], 1e-5, 1e-8)
    pf = torch.from_numpy(g['project']).cuda()
    _close(gmath.dist(pf, c, k=K), g['dist'], 1e-4, 1e-6)
    _close(gmath.dist(c, pf, k=K), g['dist_cx'], 1e-4, 1e-6)
    _close(gmath.dist0(pf, k=K), g['dist0'], 1e-4, 1e-6)
    _close(gmath.weighted_midpoint(pf, k=K), g['midpoint'], 1e-4, 1e-6)


def test_hyper_math_fixture_pinned(golden_dir):
    from coskad_b200 import gmath
    g = np.load(os.path.join(golden_dir, 'geometry_hyper_math.npz'))
    u, c = torch.from_numpy(g['u']).cuda(), torch.from_numpy(g['center']).cuda()
    e = gmath.hm_expmap0(u)
    _close(e, g['expmap0'], 1e-5, 1e-8)
    _close(gmath.hm_project(e), g['project'], 1e-5, 1e-8)
    _close(gmath.hm_dist(torch.from_numpy(g['project']).cuda(), c), g['dist'], 1e-4, 1e-6)


@pytest.mark.parametrize('D', [8, 16, 24, 64])
def test_geometry_random(D):
    from coskad_b200 import gmath
    gen = torch.Generator().manual_seed(D)
    u = torch.randn(1000, D, generator=gen) * torch.logspace(-3, 0.5, 1000)[:, None]
    y = torch.randn(1000, D, generator=gen) * 0.1
    x = ogm.project(ogm.expmap0(u, k=K), k=K)
    _close(gmath.project(gmath.expmap0(u.cuda(), k=K), k=K), x, 1e-5, 1e-8)
    xs = x * 0.8
    _close(gmath.dist(xs.cuda(), y.cuda(), k=K), ogm.dist(xs, y, k=K))
    _close(gmath.dist(xs.cuda(), y[0].cuda(), k=K), ogm.dist(xs, y[0], k=K))
    _close(gmath.weighted_midpoint(xs.cuda(), k=K), ogm.weighted_midpoint(xs, k=K), 1e-4, 1e-6)
    # edge cases: zero vector, tiny and huge norms
    e = torch.zeros(4, D)
    e[1, 0] = 1e-20
    e[2] = 1e4
    e[3, 1] = 16.0
    _close(gmath.expmap0(e.cuda(), k=K), ogm.expmap0(e, k=K), 1e-5, 1e-12)
    _close(gmath.project(gmath.expmap0(e.cuda(), k=K), k=K), ogm.project(ogm.expmap0(e, k=K), k=K), 1e-5, 1e-12)


def test_center_partials_add_like_shards():
    """the midpoint of the union equals finalize(sum of per-shard partials) -- what the all-reduce relies on"""
    from coskad_b200 import gmath
    gen = torch.Generator().manual_seed(3)
    x = ogm.project(ogm.expmap0(torch.randn(5000, 16, generator=gen) * 0.5, k=K), k=K)
    ref = ogm.weighted_midpoint(x, k=K)
    acc = gmath.center_accumulator(16, 'cuda')
    for part in x.split(777):
        gmath.center_partial(part.cuda(), acc, flavour=1)
    _close(gmath.center_finalize(acc, 16, flavour=1), ref, 1e-5, 1e-7)
    assert float(acc[17]) == 5000.0
    # euclidean mean + tolerance clamp (euclidean_encoder_staticCenter.py:118-123)
    z = torch.randn(3000, 16, generator=gen) * 0.01
    z[:, 3] = -1e-5
    c = z.sum(0) / 3000
    eps = 1e-3
    c[(abs(c) < eps) & (c < 0)] = -eps
    c[(abs(c) < eps) & (c > 0)] = eps
    acc = gmath.center_accumulator(16, 'cuda')
    gmath.center_partial(z.cuda(), acc, flavour=3)
    _close(gmath.center_finalize(acc, 16, flavour=3, eps=eps), c, 1e-5, 1e-8)


def test_poincare_score_backward_matches_autograd():
    """fp32 kernel vs float64 autograd through the restated geoopt formulas (the truth); the fp32
    autograd of the same formulas is measured alongside: the kernel must not be worse than 3x it."""
    from coskad_b200 import gmath
    gen = torch.Generator().manual_seed(9)
    z = torch.randn(512, 16, generator=gen) * torch.logspace(-2, 0.6, 512)[:, None]
    c = torch.randn(16, generator=gen) * 0.1
    w = torch.rand(512, generator=gen)
    for with_project in (True, False):
        ref = {}
        for dt in (torch.float32, torch.float64):
            zz = z.to(dt).clone().requires_grad_(True)
            x = ogm.expmap0(zz, k=K.to(dt))
            if with_project:
                x = ogm.project(x, k=K.to(dt), eps=4e-3)
            (ogm.dist(c.to(dt), x, k=K.to(dt)) * w.to(dt)).sum().backward()
            ref[dt] = zz.grad.double()
        got = gmath.poincare_score_bwd(z.cuda(), c.cuda(), w.cuda(), with_project).cpu().double()
        # without project() tanh saturates to 1.0f for |z| > ~3 and the reference's own fp32 gradient is garbage
        rows = torch.ones(512, dtype=torch.bool) if with_project else (z.norm(dim=-1) < 3.0)
        scale = ref[torch.float64][rows].abs().amax(dim=-1, keepdim=True) + 1e-30
        e_ours = float(((got[rows] - ref[torch.float64][rows]).abs() / scale).max())
        e_f32 = float(((ref[torch.float32][rows] - ref[torch.float64][rows]).abs() / scale).max())
        assert e_ours < 5e-4 and e_ours < 3 * e_f32 + 1e-6, (with_project, e_ours, e_f32)


def test_poincare_score_autograd_function():
    from coskad_b200 import gmath
    gen = torch.Generator().manual_seed(2)
    z = (torch.randn(300, 16, generator=gen) * 0.7)
    c = torch.randn(16, generator=gen) * 0.1
    zr = z.clone().requires_grad_(True)
    loss_ref = ogm.dist(c, ogm.project(ogm.expmap0(zr, k=K), k=K), k=K).mean()
    loss_ref.backward()
    zc = z.cuda().requires_grad_(True)
    s, x = gmath.poincare_score(zc, c.cuda(), True)
    s.mean().backward()
    _close(s.mean(), loss_ref.detach(), 1e-5, 1e-7)
    _close(zc.grad, zr.grad, 1e-3, 1e-6 * float(zr.grad.abs().max()) + 1e-8)
    _close(x, ogm.project(ogm.expmap0(z, k=K), k=K), 1e-5, 1e-8)


def test_frame_aggregation_bit_exact(golden_dir):
    from coskad_b200 import aggregate
    g = np.load(os.path.join(golden_dir, 'aggregate_ref.npz'))
    clips = [tuple(int(v) for v in r) for r in g['clips']]
    out = aggregate.score_and_aggregate(torch.from_numpy(g['scores']).cuda(), g['trans'], g['meta'], g['frames'], clips,
                                        num_transform=2, smooth=True)
    got = np.concatenate([c for t in range(2) for c in out[t]])
    assert got.dtype == np.float64
    assert np.array_equal(got, g['curves'])     # bit exact vs the reference's own functions


@pytest.mark.parametrize('seed', [0, 1, 2])
def test_frame_aggregation_random_datasets(seed):
    from coskad_b200 import aggregate
    trans, meta, frames, clips, gts = oagg.synth_dataset(n_clips=9, seed=seed, num_transform=3, max_persons=6)
    rng = np.random.default_rng(seed)
    scores = rng.random(len(trans)).astype(np.float32) * 3
    scores[rng.random(len(trans)) < 0.05] = 0.0
    ref = oagg.aggregate_dataset(scores, trans, meta, frames, clips, 3, smooth=False)
    out = aggregate.score_and_aggregate(torch.from_numpy(scores).cuda(), trans, meta, frames, clips, num_transform=3,
                                        smooth=False)
    for t in range(3):
        for a, b in zip(out[t], ref[t]):
            assert np.array_equal(a, b)
    # with pad_scores + smoothing + AUC (host side, same scipy/sklearn calls)
    ref = oagg.aggregate_dataset(scores, trans, meta, frames, clips, 3, pad_size=5, gts=gts)
    out = aggregate.score_and_aggregate(torch.from_numpy(scores).cuda(), trans, meta, frames, clips, num_transform=3,
                                        pad_size=5, gts=gts)
    for t in range(3):
        for a, b in zip(out[t], ref[t]):
            assert np.array_equal(a, b)


# ---- distance: 'mahalanobis' (utils/eval_utils.py:28-55, models/euclidean_encoder_staticCenter.py:40-46,133-142) -----------
@pytest.mark.parametrize('D', [8, 16])
def test_mahalanobis_matches_reference_fixture(golden_dir, D):
    """tests/golden/mahalanobis_ref.npz holds outputs of the reference's own mahalanobis / windows_based_loss_mahalanobis /
    batch_cov_mat_step (oracle/gen_golden.py)"""
    from coskad_b200 import aggregate, gmath
    g = np.load(os.path.join(golden_dir, 'mahalanobis_ref.npz'))
    z, mu, VI = (torch.from_numpy(g[f'{k}{D}']).cuda() for k in ('z', 'mu', 'VI'))
    zq = torch.from_numpy(g[f'zq{D}']).cuda()
    ref = torch.from_numpy(g[f'dist{D}'])
    got = gmath.mahalanobis_score(zq, mu, VI)
    # the last row IS the center: the reference's two matmuls give exactly 0 there and so does the kernel
    assert float(got[-1]) == 0.0 and float(ref[-1]) == 0.0
    _close(got[:-1], ref[:-1], 1e-4, 0.0)                                   # pure relative, like the score gate
    _close(gmath.mahalanobis(zq, mu, VI), ref.mean(), 1e-5, 0.0)
    assert gmath.mahalanobis(zq, mu, VI, reduce='none').shape == (zq.shape[0], 1, 1)
    # scatter matrix: shard-additive float64 sums of float32 products vs the reference's float32 matmul + sum
    acc = gmath.cov_accumulator(D, z.device)
    for i in range(0, z.shape[0], 256):
        gmath.cov_partial(z[i:i + 256], mu, acc)
    assert float(acc[D * D]) == z.shape[0]
    scat = torch.from_numpy(g[f'scatter{D}']).double()
    got_s = acc[:D * D].view(D, D).cpu()
    assert float((got_s - scat).abs().max()) <= 2e-5 * float(scat.abs().max())
    whole = gmath.cov_accumulator(D, z.device)
    gmath.cov_partial(z, mu, whole)
    assert float((whole - acc).abs().max()) <= 1e-9 * float(acc.abs().max())          # partial sums of shards add
    # the host-side inverse of an SPD 8x8 / 16x16 matrix amplifies input rounding by its condition number
    vi = gmath.inv_cov_finalize(acc, D).cpu()
    assert float((vi - VI.cpu()).abs().max()) <= 2e-3 * float(VI.abs().max())
    # compat surface: same float64 [w, n_frames] matrix as the reference's per-window loop
    pose = aggregate.windows_based_loss_mahalanobis(mu, g[f'zq{D}'][:40], VI, g[f'frames{D}'], 60)
    ref_pose = g[f'pose{D}']
    assert pose.dtype == np.float64 and pose.shape == ref_pose.shape and np.array_equal(pose != 0, ref_pose != 0)
    assert np.allclose(pose, ref_pose, rtol=1e-4, atol=0.0)


def test_mahalanobis_backward_matches_oracle_autograd(golden_dir):
    from coskad_b200 import gmath
    from oracle import mahalanobis as omah
    g = np.load(os.path.join(golden_dir, 'mahalanobis_ref.npz'))
    z = torch.from_numpy(g['z16'][:300]).double().requires_grad_(True)
    mu, VI = torch.from_numpy(g['mu16']).double(), torch.from_numpy(g['VI16']).double()
    w = torch.linspace(0.5, 1.5, 300, dtype=torch.float64)
    (omah.mahalanobis(z, mu, VI, reduce='none').view(-1) * w).sum().backward()
    zc = torch.from_numpy(g['z16'][:300]).cuda().requires_grad_(True)
    (gmath.mahalanobis_score(zc, mu.float().cuda(), VI.float().cuda()) * w.float().cuda()).sum().backward()
    err = (zc.grad.cpu().double() - z.grad).abs()
    assert float(err.max()) <= 1e-4 * float(z.grad.abs().max()), float(err.max())
    # a non-symmetric VI: the gradient is (VI + VI^T) d / (2 dist), not VI d / dist
    VIa = VI + 0.05 * torch.randn(16, 16, dtype=torch.float64, generator=torch.Generator().manual_seed(3)).triu(1)
    z2 = torch.from_numpy(g['z16'][:64]).double().requires_grad_(True)
    omah.mahalanobis(z2, mu, VIa).backward()
    zc2 = torch.from_numpy(g['z16'][:64]).cuda().requires_grad_(True)
    gmath.mahalanobis(zc2, mu.float().cuda(), VIa.float().cuda()).backward()
    assert float((zc2.grad.cpu().double() - z2.grad).abs().max()) <= 1e-4 * float(z2.grad.abs().max())
