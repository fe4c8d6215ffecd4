"""GPU parity of the trajectory front end (coskad_encode_score_traj_fwd): sliding-window construction and the
test-time affine transforms inside the fused kernel vs the materialised windows of the reference
(fixture windows_ref.npz: outputs of utils.preprocessing / utils.dataset_utils run by oracle/gen_golden.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import geoopt_math as ogm
from oracle import stsgcn as onet
from oracle import windows as owin
from tests.helpers import make_pair

pytestmark = pytest.mark.gpu
K = torch.tensor(-1.)
RTOL = 1e-4


def _close(got, ref, what, atol_scale=1e-5):
    got, ref = got.detach().cpu().double(), ref.detach().cpu().double()
    atol = atol_scale * float(ref.abs().max())
    bad = (got - ref).abs() > RTOL * ref.abs() + atol
    assert not bool(bad.any()), f'{what}: {int(bad.sum())}/{bad.numel()} outside rtol {RTOL}; max abs err {float((got - ref).abs().max()):.3e}'


def test_windows_from_trajectory_bit_identical(golden_dir):
    """no transform: the gathered windows take exactly the arithmetic of the materialised ones"""
    g = np.load(os.path.join(golden_dir, 'windows_ref.npz'))
    m, _ = make_pair('stse', 16, seed=0)
    traj = torch.from_numpy(g['traj']).cuda()
    starts = torch.from_numpy(g['starts']).cuda()
    c = torch.zeros(16); c[0] = 0.05
    z_t, s_t = m.encode_score_traj(traj, starts, flavour=1, center=c)
    z_w, s_w = m.encode_score(torch.from_numpy(g['windows']).cuda(), 1, center=c)
    assert torch.equal(z_t, z_w) and torch.equal(s_t, s_w)


def test_transforms_match_reference_windows(golden_dir):
    """5 x N windows in the reference's dataset order (index = trans * N + sample, utils/dataset.py:65-74): scores of the
    in-kernel transform vs the oracle network on the REFERENCE-transformed windows"""
    g = np.load(os.path.join(golden_dir, 'windows_ref.npz'))
    m, sd = make_pair('stse', 16, seed=0)
    N, T = len(g['starts']), len(g['mats'])
    traj = torch.from_numpy(g['traj']).cuda()
    win_row = torch.from_numpy(np.tile(g['starts'], T)).cuda()
    trans = torch.arange(T).repeat_interleave(N).cuda()
    mats = torch.from_numpy(g['mats'][:, :2, :]).cuda()
    xr = torch.from_numpy(g['transformed'].reshape(T * N, 2, 12, 17))
    with torch.no_grad():
        zr = onet.stse_forward(xr, sd)
        pr = ogm.project(ogm.expmap0(zr, k=K), k=K)
        c = ogm.weighted_midpoint(pr, k=K)
        sr = ogm.dist(pr, c, k=K)
    z, s = m.encode_score_traj(traj, win_row, trans, mats, flavour=1, center=c)
    _close(z, zr, 'latent (trajectory + transform front end)')
    _close(s, sr, 'poincare score (trajectory + transform front end)')
    # and against the same kernel fed with the materialised reference windows
    z2, s2 = m.encode_score(xr.cuda(), 1, center=c)
    _close(z, z2, 'latent vs materialised windows', atol_scale=1e-6)


@pytest.mark.parametrize('N', [1, 2, 4, 7, 500])
def test_ragged_and_multi_person(N):
    """several persons concatenated in one trajectory buffer, window counts that are not multiples of the tile"""
    rng = np.random.default_rng(N)
    lens = [12, 30, 13, 57]
    traj = (rng.standard_normal((sum(lens), 34)) * 0.4).astype(np.float32)
    offs = np.cumsum([0] + lens[:-1])
    rows = np.concatenate([o + owin.sliding_starts(l, 12, 1) for o, l in zip(offs, lens)])
    rows = rows[rng.integers(0, len(rows), size=N)]
    tr = rng.integers(0, 5, size=N)
    mats = owin.ae_trans_mats()
    m, sd = make_pair('stse', 16, seed=0)
    x = owin.windows_from_rows(traj, rows)
    x = np.stack([owin.apply_pose_transform(w, mats[t]) for w, t in zip(x, tr)], 0).astype(np.float32)
    with torch.no_grad():
        zr = onet.stse_forward(torch.from_numpy(x), sd)
    z, s = m.encode_score_traj(torch.from_numpy(traj).cuda(), torch.from_numpy(rows).cuda(), torch.from_numpy(tr).cuda(),
                               torch.from_numpy(mats[:, :2]).cuda())
    assert s is None and z.shape == (N, 16)
    _close(z, zr, f'latent, N={N}')


def test_front_end_argument_errors():
    from coskad_b200 import _lib
    m, _ = make_pair('stse', 16, seed=0)
    traj = torch.zeros(20, 34, device='cuda')
    rows = torch.zeros(3, dtype=torch.int64, device='cuda')
    with pytest.raises(ValueError):
        m.encode_score_traj(traj, rows, trans=torch.zeros(3, dtype=torch.int32, device='cuda'))      # trans without mats
    with pytest.raises(ValueError):
        m.encode_score_traj(torch.zeros(20, 33, device='cuda'), rows)                                 # wrong row width
    with pytest.raises(_lib.CoskadError):
        m.encode_score_traj(torch.zeros(5, 34, device='cuda'), rows)                                  # shorter than a window
    # out-of-range rows are clamped into the buffer, never read outside it
    z, _ = m.encode_score_traj(traj, torch.tensor([-5, 100, 8], device='cuda'))
    assert bool(torch.isfinite(z).all())


@pytest.mark.parametrize('n', [1, 777, 40_001])
def test_host_scorers_equal_direct_calls(n):
    """pipeline.HostScorer / TrajectoryScorer (chunked H2D on a copy stream, scores back to pinned host memory) return exactly
    what one direct device call returns, for sizes that are not multiples of the chunk"""
    from coskad_b200.pipeline import HostScorer, TrajectoryScorer
    m, _ = make_pair('stse', 16, seed=0)
    rng = np.random.default_rng(n)
    c = torch.full((16,), 0.01, device='cuda')
    # windows
    xh = torch.from_numpy((rng.standard_normal((n, 2, 12, 17)) * 0.4).astype(np.float32)).pin_memory()
    got = HostScorer(m, 1, chunk=4096).score(xh, center=c)
    _, ref = m.encode_score(xh.cuda(), 1, center=c)
    assert torch.equal(got, ref.cpu())
    # trajectories: one long person, n windows at random starts, 5 transforms
    traj = torch.from_numpy((rng.standard_normal((500, 34)) * 0.4).astype(np.float32)).pin_memory()
    rows = torch.from_numpy(rng.integers(0, 500 - 12 + 1, size=n)).pin_memory()
    tr = torch.from_numpy(rng.integers(0, 5, size=n).astype(np.int32)).pin_memory()
    mats = torch.from_numpy(owin.ae_trans_mats()[:, :2].copy())
    got = TrajectoryScorer(m, 1, chunk=8192).score(traj, rows, tr, mats, center=c)
    _, ref = m.encode_score_traj(traj.cuda(), rows.cuda(), tr.cuda(), mats.cuda(), flavour=1, center=c)
    assert torch.equal(got, ref.cpu())
