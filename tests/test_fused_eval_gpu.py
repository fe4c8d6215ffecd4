"""GPU parity: the fused sm_100a eval kernel (through the C ABI) against the CPU oracle.

Tolerances: latents / scores within 1e-4 relative in fp32 (BASELINE.json north_star); the fixtures
in tests/golden were produced by the reference's own modules (oracle/gen_golden.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import geoopt_math as ogm
from oracle import hyper_math as ohm
from oracle import stsgcn as onet
from tests.helpers import make_pair, rel_err, stage_report

pytestmark = pytest.mark.gpu
K = torch.tensor(-1.)
RTOL = 1e-4


def _assert_close(got, ref, what, atol_scale=1e-5, report=None):
    got, ref = got.detach().cpu().double(), ref.detach().cpu().double()
    atol = atol_scale * float(ref.abs().max())
    bad = (got - ref).abs() > RTOL * ref.abs() + atol
    if bool(bad.any()):
        msg = f'{what}: {int(bad.sum())}/{bad.numel()} outside rtol {RTOL}; max abs err {float((got - ref).abs().max()):.3e}'
        if report is not None:
            msg += '\n' + report()
        raise AssertionError(msg)


def test_golden_reference_vectors(golden_dir):
    """the reference's own outputs on 10 windows (fixture made by models.sts.ae.STSE itself)"""
    g = np.load(os.path.join(golden_dir, 'stse_ref.npz'))
    m, sd = make_pair('stse', 16, seed=0)
    x = torch.from_numpy(g['x'])
    z, _ = m.encode_score(x.cuda())
    _assert_close(z, torch.from_numpy(g['z']), 'latent vs reference fixture', report=lambda: stage_report(m, sd, x))


@pytest.mark.parametrize('impl', [0, 1])
@pytest.mark.parametrize('B', [1, 2, 3, 4, 5, 7, 448, 1001])
def test_ragged_batches(B, impl):
    """impl 1: tcgen05 (3xTF32) channel mixing; impl 0: all-FP32 CUDA-core kernel"""
    m, sd = make_pair('stse', 16, seed=0)
    m.fused_impl = impl
    x = onet.synth_windows(B, seed=B)
    with torch.no_grad():
        ref = onet.stse_forward(x, sd)
    z = m(x.cuda())
    assert z.shape == (B, 16)
    _assert_close(z, ref, f'latent B={B}', report=lambda: stage_report(m, sd, x))


@pytest.mark.parametrize('impl', [0, 1])
@pytest.mark.parametrize('shape', ['ubnormal', 'stc'])
def test_config0_4096_windows(shape, impl):
    """BASELINE.json configs[0]: 4096 synthetic 17-joint windows, hyperbolic static-center scoring"""
    m, sd = make_pair('stse', 16, seed=0)
    m.fused_impl = impl
    x = onet.synth_windows(4096, seed=999, shape=shape)
    with torch.no_grad():
        zr = onet.stse_forward(x, sd)
        xr = ogm.project(ogm.expmap0(zr, k=K), k=K)
        c = ogm.weighted_midpoint(xr, k=K)
        sr = ogm.dist(xr, c, k=K)
    z, s = m.encode_score(x.cuda(), 1, center=c.cuda())
    _assert_close(z, zr, 'latent')                      # signed components cross zero: 1e-4 relative + 1e-5 of the largest
    # the north-star gate: per-window SCORES within 1e-4 RELATIVE, no absolute slack (scores are bounded away from zero)
    _assert_close(s, sr, 'poincare score', atol_scale=0.0)
    # the score recomputed by the oracle FROM THE KERNEL'S latent isolates the geometry arithmetic
    with torch.no_grad():
        s2 = ogm.dist(ogm.project(ogm.expmap0(z.cpu(), k=K), k=K), c, k=K)
    assert rel_err(s, s2, atol=1e-6) < 2e-5


@pytest.mark.parametrize('flavour', [1, 2, 3, 4, 5])
def test_score_flavours(flavour):
    m, sd = make_pair('stse', 16, seed=2)
    x = onet.synth_windows(600, seed=5)
    g = torch.Generator().manual_seed(0)
    c = torch.randn(16, generator=g) * 0.05
    with torch.no_grad():
        zr = onet.stse_forward(x, sd)
        if flavour == 1:
            sr = ogm.dist(ogm.project(ogm.expmap0(zr, k=K), k=K), c, k=K)
        elif flavour == 2:
            sr = ogm.dist(ogm.expmap0(zr, k=K), c, k=K)
        elif flavour == 3:
            sr = torch.mean(torch.nn.MSELoss(reduction='none')(c.expand_as(zr), zr), dim=-1)     # eval_utils.py:63-64
        elif flavour == 4:
            sr = 1 - torch.nn.functional.cosine_similarity(c.expand_as(zr), zr)                  # eval_COSKAD.py:81
        else:
            sr = ohm.dist(ohm.project(ohm.expmap0(zr)), c.expand_as(zr))
    _, s = m.encode_score(x.cuda(), flavour, center=c.cuda())
    _assert_close(s, sr, f'score flavour {flavour}')


def test_large_latents_hit_the_ball_boundary():
    """scale the bottleneck so |z| is large: project() must clip to 1 - 4e-3 exactly like the oracle"""
    m, sd = make_pair('stse', 16, seed=0)
    sd = {k: v.clone() for k, v in sd.items()}
    sd['btlnk.weight'] *= 40.0
    m.load_state_dict(sd)
    x = onet.synth_windows(300, seed=3)
    with torch.no_grad():
        zr = onet.stse_forward(x, sd)
        xr = ogm.project(ogm.expmap0(zr, k=K), k=K)
        assert float((xr.norm(dim=-1) > 0.995).float().mean()) > 0.5
    z, s = m.encode_score(x.cuda(), 1, center=torch.zeros(16).cuda())
    with torch.no_grad():
        s2 = ogm.dist(ogm.project(ogm.expmap0(z.cpu(), k=K), k=K), torch.zeros(16), k=K)
    # at the boundary artanh' = 125: compare the geometry on identical latents
    assert rel_err(s, s2, atol=1e-6) < 1e-4


def test_weights_resync_after_update():
    m, sd = make_pair('stse', 16, seed=0)
    x = onet.synth_windows(33, seed=1)
    z0 = m(x.cuda()).clone()
    with torch.no_grad():
        m.encoder.model[2].prelu.weight.fill_(0.1)
        m.encoder.model[1].tcn[1].running_var.mul_(1.7)
    sd2 = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    with torch.no_grad():
        ref = onet.stse_forward(x, sd2)
    z1 = m(x.cuda())
    assert float((z1 - z0).abs().max()) > 1e-4
    _assert_close(z1, ref, 'latent after in-place weight update')


def test_autoencoder_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, 'stsae_ref.npz'))
    m, sd = make_pair('stsae', 8, seed=1)
    x = torch.from_numpy(g['x'])
    z, xh = m(x.cuda())
    _assert_close(z, torch.from_numpy(g['z']), 'AE latent vs reference fixture')
    _assert_close(xh, torch.from_numpy(g['xhat']), 'AE reconstruction vs reference fixture', atol_scale=1e-4)


@pytest.mark.parametrize('impl', [0, 1])
@pytest.mark.parametrize('B', [1, 5, 1000])
def test_autoencoder_scores(B, impl):
    """impl 1: tensor-core encoder + FP32 decoder stages in fused_eval_tc_kernel<true>; impl 0: the all-FP32 kernel"""
    m, sd = make_pair('stsae', 8, seed=1)
    m.fused_impl = impl
    x = onet.synth_windows(B, seed=10 + B)
    c = torch.full((8,), 0.02)
    with torch.no_grad():
        zr, xr = onet.stsae_forward(x, sd)
        rec = torch.mean((x - xr).permute(0, 2, 3, 1).reshape(B, -1) ** 2, dim=-1)     # eval_utils.py:81-87
        lat = torch.mean((c - zr) ** 2, dim=-1)
    z, xh, rs, ls = m.autoencode_score(x.cuda(), center=c.cuda())
    _assert_close(z, zr, 'AE latent')
    _assert_close(xh, xr, 'AE reconstruction', atol_scale=1e-4)
    _assert_close(rs, rec, 'reconstruction score')
    _assert_close(ls, lat, 'latent score')


def test_full_size_properties():
    """size-independent properties at a size the oracle cannot reach: scoring is per-window, so a
    permutation of the windows permutes the scores, and chunked calls equal one big call bit for bit."""
    m, _ = make_pair('stse', 16, seed=0)
    B = 300_000
    g = torch.Generator(device='cuda').manual_seed(1)
    x = (torch.randn(B, 2, 12, 17, device='cuda', generator=g) * 0.4).clamp_(-3, 3)
    c = torch.full((16,), 0.01, device='cuda')
    _, s = m.encode_score(x, 1, center=c)
    perm = torch.randperm(B, device='cuda', generator=g)
    _, sp = m.encode_score(x[perm].contiguous(), 1, center=c)
    assert torch.equal(sp, s[perm])
    _, s1 = m.encode_score(x[:100_001].contiguous(), 1, center=c)
    _, s2 = m.encode_score(x[100_001:].contiguous(), 1, center=c)
    assert torch.equal(torch.cat([s1, s2]), s)
    assert bool(torch.isfinite(s).all()) and float(s.min()) >= 0


def test_empty_batch_and_bad_arguments():
    """B = 0 is a no-op that returns empty tensors; wrong shapes / missing center fail loudly (no silent fallback)"""
    from coskad_b200 import _lib
    m, _ = make_pair('stse', 16, seed=0)
    z, s = m.encode_score(torch.empty(0, 2, 12, 17, device='cuda'), 1, center=torch.zeros(16, device='cuda'))
    assert z.shape == (0, 16) and s.shape == (0,)
    ae, _ = make_pair('stsae', 8, seed=1)
    z, xh, rs, ls = ae.autoencode_score(torch.empty(0, 2, 12, 17, device='cuda'), center=torch.zeros(8, device='cuda'))
    assert z.shape == (0, 8) and xh.shape == (0, 2, 12, 17) and rs.shape == (0,) and ls.shape == (0,)
    zt, st = m.encode_score_traj(torch.zeros(12, 34, device='cuda'), torch.zeros(0, dtype=torch.int64, device='cuda'),
                                 flavour=1, center=torch.zeros(16, device='cuda'))
    assert zt.shape == (0, 16) and st.shape == (0,)
    with pytest.raises((ValueError, _lib.CoskadError)):
        m.encode_score(torch.zeros(4, 2, 12, 16, device='cuda'))               # 16 joints: the kernels are built for 17
    with pytest.raises((ValueError, _lib.CoskadError)):
        m.encode_score(torch.zeros(4, 3, 12, 17, device='cuda'))               # 3 coordinates
    with pytest.raises((ValueError, _lib.CoskadError, RuntimeError, TypeError)):
        m.encode_score(torch.zeros(4, 2, 12, 17))                               # host tensor: there is no CPU path


def test_scores_invariant_to_launch_geometry():
    """a window's score does not depend on which tile / CTA / position inside the tile it lands in"""
    m, _ = make_pair('stse', 16, seed=0)
    g = torch.Generator(device='cuda').manual_seed(7)
    x = (torch.randn(1000, 2, 12, 17, device='cuda', generator=g) * 0.4).clamp_(-3, 3)
    c = torch.full((16,), 0.01, device='cuda')
    _, s = m.encode_score(x, 1, center=c)
    for off in (1, 2, 5):
        _, s2 = m.encode_score(x[off:].contiguous(), 1, center=c)
        assert torch.equal(s2, s[off:])
    _, s1 = m.encode_score(x[:1].contiguous(), 1, center=c)
    assert torch.equal(s1, s[:1])


@pytest.mark.parametrize('kind', ['stse16', 'stse8', 'stsae8'])
def test_repeat_launches_are_bit_identical(kind):
    """the kernel software-pipelines tiles (deferred score, layer 1 one tile ahead, aliased TMEM buffers, barrier parities that
    run across tiles): any race or stale buffer shows up as run-to-run differences, so 10 launches must agree bit for bit"""
    g = torch.Generator(device='cuda').manual_seed(11)
    x = (torch.randn(50_001, 2, 12, 17, device='cuda', generator=g) * 0.4).clamp_(-3, 3)
    if kind == 'stsae8':
        m, _ = make_pair('stsae', 8, seed=1)
        c = torch.full((8,), 0.02, device='cuda')
        run = lambda: m.autoencode_score(x, center=c)
    else:
        d = 16 if kind == 'stse16' else 8
        m, _ = make_pair('stse', d, seed=0)
        c = torch.full((d,), 0.01, device='cuda')
        run = lambda: m.encode_score(x, 1 if d == 16 else 3, center=c)
    ref = [t.clone() for t in run()]
    for _ in range(9):
        out = run()
        for a, b in zip(out, ref):
            assert torch.equal(a, b)
    assert all(bool(torch.isfinite(t).all()) for t in ref)
