"""CPU: the oracle restatements against the fixtures generated from the REAL reference
(oracle/gen_golden.py ran the reference's own modules; see its header for what is pinned)."""
import os

import numpy as np
import torch

from oracle import aggregate as oagg
from oracle import geoopt_math as ogm
from oracle import hyper_math as ohm
from oracle import stsgcn as onet


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def test_stse_oracle_matches_reference_vectors(golden_dir):
    g = _load(golden_dir, 'stse_ref.npz')
    sd = onet.init_state_dict('stse', latent_dim=16, seed=0)
    chk = sum(float(v.double().sum()) for k, v in sd.items() if v.is_floating_point())
    assert abs(chk - float(g['sd_checksum'])) < 1e-9, 'seeded state dict differs from the one the fixture was made with'
    x = torch.from_numpy(g['x'])
    with torch.no_grad():
        z = onet.stse_forward(x, sd)
        h1 = onet.st_gcnn_layer(x, sd, 'encoder.model.0')
    assert torch.allclose(z, torch.from_numpy(g['z']), rtol=1e-5, atol=1e-6)
    assert torch.allclose(h1, torch.from_numpy(g['h1']), rtol=1e-5, atol=1e-6)


def test_synth_windows_reproducible(golden_dir):
    g = _load(golden_dir, 'stse_ref.npz')
    assert np.array_equal(onet.synth_windows(10, seed=999).numpy(), g['x'])


def test_stsae_oracle_matches_reference_vectors(golden_dir):
    g = _load(golden_dir, 'stsae_ref.npz')
    sd = onet.init_state_dict('stsae', latent_dim=8, seed=1)
    with torch.no_grad():
        z, xh = onet.stsae_forward(torch.from_numpy(g['x']), sd)
    assert torch.allclose(z, torch.from_numpy(g['z']), rtol=1e-5, atol=1e-6)
    assert torch.allclose(xh, torch.from_numpy(g['xhat']), rtol=1e-5, atol=1e-6)


def test_train_mode_oracle_matches_reference(golden_dir):
    g = _load(golden_dir, 'stse_train_ref.npz')
    sd = onet.init_state_dict('stse', latent_dim=16, seed=0)
    params = {k: v.clone().requires_grad_(v.is_floating_point() and 'running' not in k and k != 'c') for k, v in sd.items()}
    stats = {}
    z = onet.stse_forward(torch.from_numpy(g['x']), params, training=True, new_stats=stats)
    assert torch.allclose(z, torch.from_numpy(g['z']), rtol=1e-4, atol=1e-5)
    (z ** 2).mean().backward()
    for k in g.files:
        if k.startswith('grad.'):
            ref = torch.from_numpy(g[k])
            got = params[k[5:]].grad
            assert torch.allclose(got, ref, rtol=1e-3, atol=1e-6 + 1e-4 * float(ref.abs().max())), k
        if k.startswith('stat.'):
            assert torch.allclose(stats[k[5:]], torch.from_numpy(g[k]), rtol=1e-5, atol=1e-6), k


def test_fold_is_exact_algebra():
    sd = onet.init_state_dict('stse', seed=3)
    x = onet.synth_windows(5, seed=1)
    for i in range(4):
        pfx = f'encoder.model.{i}'
        with torch.no_grad():
            ref = onet.st_gcnn_layer(x, sd, pfx)
            w1, w2, b = onet.fold_layer_eval(sd, pfx)
            g = onet.graph_contract(x, sd[pfx + '.gcn.A'], sd[pfx + '.gcn.T'])
            pre = torch.einsum('oc,nctv->notv', w1, g) + torch.einsum('oc,nctv->notv', w2, x) + b[None, :, None, None]
            got = torch.nn.functional.prelu(pre, sd[pfx + '.prelu.weight'])
        assert float((got - ref).abs().max()) < 1e-5 * max(1.0, float(ref.abs().max()))
        x = ref


def test_hyper_math_flavour_pinned(golden_dir):
    g = _load(golden_dir, 'geometry_hyper_math.npz')
    u, c = torch.from_numpy(g['u']), torch.from_numpy(g['center'])
    e = ohm.expmap0(u)
    p = ohm.project(e)
    assert torch.allclose(e, torch.from_numpy(g['expmap0']), rtol=1e-6, atol=1e-8)
    assert torch.allclose(p, torch.from_numpy(g['project']), rtol=1e-6, atol=1e-8)
    assert torch.allclose(ohm.dist(p, c.expand_as(p)), torch.from_numpy(g['dist']), rtol=1e-5, atol=1e-7)
    assert torch.allclose(ohm.poincare_mean(p[:40] * 0.5), torch.from_numpy(g['mean']), rtol=1e-5, atol=1e-7)


def test_geoopt_restatement_properties():
    """geoopt is not installable here (parity unpinned): check the restatement's defining properties."""
    k = torch.tensor(-1.)
    g = torch.Generator().manual_seed(1)
    u = torch.randn(256, 16, generator=g) * torch.logspace(-2, 1, 256)[:, None]
    x = ogm.project(ogm.expmap0(u, k=k), k=k)
    assert float(x.norm(dim=-1).max()) <= 1 - 4e-3 + 1e-6                     # ball containment
    y = ogm.project(ogm.expmap0(torch.randn(256, 16, generator=g), k=k), k=k)
    xs_, ys_ = x * 0.7, y * 0.7      # away from the boundary: artanh' = 1/(1-r^2) amplifies fp32 rounding there
    assert torch.allclose(ogm.dist(xs_, ys_, k=k), ogm.dist(ys_, xs_, k=k), rtol=1e-4, atol=1e-5)   # symmetry
    assert torch.allclose(ogm.dist(x, torch.zeros(16), k=k), ogm.dist0(x, k=k), rtol=1e-5, atol=1e-6)
    # closed form: d(0, x) = 2 artanh |x|
    xs = x[x.norm(dim=-1) < 0.9]
    assert torch.allclose(ogm.dist0(xs, k=k), 2 * torch.atanh(xs.norm(dim=-1)), rtol=1e-4, atol=1e-6)
    # midpoint is permutation invariant and of one point is the point
    m1 = ogm.weighted_midpoint(x[:64] * 0.5, k=k)
    m2 = ogm.weighted_midpoint((x[:64] * 0.5)[torch.randperm(64, generator=g)], k=k)
    assert torch.allclose(m1, m2, rtol=1e-4, atol=1e-6)
    one = x[3:4] * 0.3
    assert torch.allclose(ogm.weighted_midpoint(one, k=k), one[0], rtol=1e-4, atol=1e-6)
    # agrees with the hyper_math flavour away from the boundary, to the size of the constants
    xi = x[x.norm(dim=-1) < 0.5]
    assert torch.allclose(ogm.dist(xi, y[:len(xi)] * 0.3, k=k), ohm.dist(xi, y[:len(xi)] * 0.3), rtol=1e-3, atol=1e-4)


def test_aggregation_oracle_matches_reference_functions(golden_dir):
    g = _load(golden_dir, 'aggregate_ref.npz')
    clips = [tuple(int(v) for v in r) for r in g['clips']]
    agg = oagg.aggregate_dataset(g['scores'], g['trans'], g['meta'], g['frames'], clips, 2)
    got = np.concatenate([c for t in range(2) for c in agg[t]])
    assert np.array_equal(got, g['curves'])           # bit exact float64
    assert np.array_equal(oagg.pad_scores(g['pad_in'].copy(), np.zeros(20), 2), g['pad_out'])


def test_frame_id_zero_wraps_and_zero_scores_are_absent():
    frames = np.array([[0, 1, 2], [1, 2, 3]])
    loss = np.array([0.5, 0.0], dtype=np.float32)
    pose = oagg.scatter_windows(loss, frames, 6)
    assert pose[0, -1] == 0.5 and pose[0, 0] == 0.5 and pose[0, 1] == 0.5
    cur = oagg.person_curve(loss, frames, 6)
    assert np.array_equal(cur, np.array([0.5, 0.5, 0, 0, 0, 0.5]))


def test_power_spherical_restatement_properties():
    """power_spherical is un-vendored upstream (parity unpinned): check defining properties of the restatement"""
    from oracle import power_spherical as ops
    g = torch.Generator().manual_seed(0)
    mu = torch.nn.functional.normalize(torch.randn(2000, 8, generator=g), dim=-1)
    kappa = torch.full((2000,), 30.0)
    t, v = ops.draw_noise(kappa, 8, generator=g)
    z = ops.rsample_from_noise(mu, t, v)
    assert torch.allclose(z.norm(dim=-1), torch.ones(2000), atol=1e-3)            # samples live on the sphere
    assert float((z * mu).sum(-1).mean()) > 0.8                                     # concentrated around loc
    assert torch.allclose((z * mu).sum(-1), t.squeeze(-1), atol=2e-3)               # <z, mu> = t (Householder maps e1 -> mu)
    # entropy decreases with concentration and tends to the uniform entropy as kappa -> 0
    k = torch.tensor([1e-4, 1.0, 10.0, 100.0])
    h = ops.ps_entropy(k, 8)
    assert bool((h[1:] < h[:-1]).all()) and abs(float(h[0]) - ops.hu_entropy(8)) < 1e-3
    assert bool((ops.kl_ps_uniform(k, 8) >= -1e-6).all())


def test_windows_oracle_matches_reference_vectors(golden_dir):
    """window construction + test-time transforms vs outputs of the reference's own utils.preprocessing /
    utils.dataset_utils (fixture written by oracle/gen_golden.py)"""
    from oracle import windows as owin
    g = _load(golden_dir, 'windows_ref.npz')
    assert np.array_equal(owin.ae_trans_mats(), g['mats'])
    assert np.array_equal(owin.sliding_starts(g['traj'].shape[0], 12, 1), g['starts'])
    w = owin.windows_from_rows(g['traj'], g['starts'])
    assert np.array_equal(w, g['windows'])
    for t, m in enumerate(g['mats']):
        tr = np.stack([owin.apply_pose_transform(x, m) for x in w], 0).astype(np.float32)
        assert np.array_equal(tr, g['transformed'][t])
    # identity transform first, x-flip second (utils/dataset_utils.py:304-306)
    assert np.array_equal(g['transformed'][0], w)
    assert np.array_equal(g['transformed'][1][:, 0], -w[:, 0]) and np.array_equal(g['transformed'][1][:, 1], w[:, 1])


def test_fused_reg_loss_matches_reference_formula():
    """coskad_b200.losses.calc_reg_loss (one multi-tensor norm) vs the per-tensor loop of utils/model_utils.py:90-105"""
    from coskad_b200.losses import calc_reg_loss
    torch.manual_seed(3)
    m = torch.nn.Sequential(torch.nn.Conv2d(2, 8, 1), torch.nn.BatchNorm2d(8), torch.nn.PReLU(), torch.nn.Linear(17, 5))
    params = [p for n, p in m.named_parameters() if 'bias' not in n]
    ref = None
    for p in params:
        ref = 0.5 * torch.sum(p ** 2) if ref is None else ref + 0.5 * p.norm(2) ** 2
    ref = ref / len(params)
    got = calc_reg_loss(m)
    assert abs(float(got.detach()) - float(ref.detach())) <= 1e-6 * abs(float(ref.detach()))
    g_ref = torch.autograd.grad(ref * 2.5, params)
    g_got = torch.autograd.grad(got * 2.5, params)
    for a, b in zip(g_got, g_ref):
        assert torch.allclose(a, b, rtol=1e-6, atol=1e-8)
    assert all(p.grad is None for n, p in m.named_parameters() if 'bias' in n)


def test_geoopt_restatement_satisfies_the_poincare_ball_identities():
    """geoopt itself cannot be imported here (parity unpinned, oracle/geoopt_math.py header); what CAN be checked without it is
    that the restated formulas are the Poincare-ball operations they claim to be: closed forms in float64."""
    torch.manual_seed(0)
    k = torch.tensor(-1., dtype=torch.float64)
    u = torch.randn(256, 16, dtype=torch.float64) * 0.3
    v = torch.randn(256, 16, dtype=torch.float64) * 0.3
    x, y = ogm.expmap0(u, k=k), ogm.expmap0(v, k=k)
    # expmap0 lands inside the unit ball with ||x|| = tanh(||u||); dist0 is its inverse: d(0, x) = 2 ||u||
    assert bool((x.norm(dim=-1) < 1).all())
    assert torch.allclose(x.norm(dim=-1), torch.tanh(u.norm(dim=-1)), rtol=1e-12)
    assert torch.allclose(ogm.dist0(x, k=k), 2 * u.norm(dim=-1), rtol=1e-9)
    # distance: symmetric, zero on the diagonal, equal to the closed form arcosh(1 + 2|x-y|^2 / ((1-|x|^2)(1-|y|^2)))
    d = ogm.dist(x, y, k=k)
    assert torch.allclose(d, ogm.dist(y, x, k=k), rtol=1e-10)
    assert float(ogm.dist(x, x, k=k).abs().max()) < 1e-6
    closed = torch.acosh(1 + 2 * (x - y).pow(2).sum(-1) / ((1 - x.pow(2).sum(-1)) * (1 - y.pow(2).sum(-1))))
    assert torch.allclose(d, closed, rtol=1e-8, atol=1e-10)
    assert torch.allclose(ogm.dist0(x, k=k), ogm.dist(torch.zeros_like(x), x, k=k), rtol=1e-9)
    # Moebius addition: left identity, left inverse, and it is an isometry: d(a + x, a + y) = d(x, y)
    a = ogm.expmap0(torch.randn(256, 16, dtype=torch.float64) * 0.2, k=k)
    assert torch.allclose(ogm.mobius_add(torch.zeros_like(x), x, k=k), x, atol=1e-14)
    assert float(ogm.mobius_add(-x, x, k=k).abs().max()) < 1e-12
    assert torch.allclose(ogm.dist(ogm.mobius_add(a, x, k=k), ogm.mobius_add(a, y, k=k), k=k), d, rtol=1e-7, atol=1e-9)
    # gyro-midpoint: of one repeated point it is that point; of two points it is equidistant and halves the distance
    p = x[:1]
    assert torch.allclose(ogm.weighted_midpoint(p.expand(7, -1), k=k), p[0], atol=1e-12)
    for i in range(8):
        m = ogm.weighted_midpoint(torch.stack([x[i], y[i]]), k=k)
        dx, dy = ogm.dist(m, x[i], k=k), ogm.dist(m, y[i], k=k)
        assert abs(float(dx - dy)) < 1e-9 and abs(float(dx + dy - d[i])) < 1e-8
    # project only touches points outside the (1 - eps) ball and puts them on its boundary along the same ray
    far = x / x.norm(dim=-1, keepdim=True) * 0.9999
    pr = ogm.project(far.float(), k=torch.tensor(-1.))
    assert torch.allclose(pr.norm(dim=-1), torch.full((256,), 1 - 4e-3), rtol=1e-6)
    assert torch.allclose(pr / pr.norm(dim=-1, keepdim=True), (far / far.norm(dim=-1, keepdim=True)).float(), atol=1e-6)
    assert torch.equal(ogm.project(x.float(), k=torch.tensor(-1.)), x.float())


def test_power_spherical_restatement_is_the_distribution_it_claims():
    """power_spherical cannot be imported (parity unpinned); the restatement is checked against what defines the distribution:
    density p(x) ~ (1 + mu.x)^kappa on S^{d-1}  =>  marginal of t = mu.x on [-1, 1] ~ (1+t)^(kappa + (d-3)/2) (1-t)^((d-3)/2),
    i.e. t = 2 Beta(alpha, beta) - 1 with the restated (alpha, beta); entropy / KL against numerical integration."""
    import math
    from oracle import power_spherical as ops
    d = 8
    for kappa in (0.5, 3.0, 25.0):
        kap = torch.tensor([kappa], dtype=torch.float64)
        a, b = ops.ps_alpha_beta(kap, d)
        assert abs(float(a) - ((d - 1) / 2 + kappa)) < 1e-12 and abs(float(b) - (d - 1) / 2) < 1e-12
        # numerical entropy of the density on the sphere: p(x) = C (1 + t)^kappa, surface element ~ |S^{d-2}| (1 - t^2)^((d-3)/2) dt
        t = torch.linspace(-1, 1, 2_000_001, dtype=torch.float64)[1:-1]
        area = 2 * math.pi ** ((d - 1) / 2) / math.gamma((d - 1) / 2)              # |S^{d-2}|
        w = area * (1 - t * t) ** ((d - 3) / 2)
        un = (1 + t) ** kappa
        Z = torch.trapezoid(un * w, t)
        p = un / Z
        H = -torch.trapezoid(p * torch.log(p) * w, t)
        assert abs(float(H) - float(ops.ps_entropy(kap, d))) < 1e-5 * max(1.0, abs(float(H)))
        sphere = 2 * math.pi ** (d / 2) / math.gamma(d / 2)                          # |S^{d-1}|
        assert abs(ops.hu_entropy(d) - math.log(sphere)) < 1e-12
        assert float(ops.kl_ps_uniform(kap, d)) > 0
    # sampling transform: unit norm, the t-coordinate is the cosine to the mean direction, t = 1 gives the mean itself
    g = torch.Generator().manual_seed(0)
    loc = torch.randn(64, d, generator=g, dtype=torch.float64)
    loc = loc / loc.norm(dim=-1, keepdim=True)
    tt, v = ops.draw_noise(torch.full((64,), 4.0, dtype=torch.float64), d, generator=g)
    z = ops.rsample_from_noise(loc, tt, v)
    assert torch.allclose(z.norm(dim=-1), torch.ones(64, dtype=torch.float64), atol=1e-4)
    assert torch.allclose((z * loc).sum(-1), tt.squeeze(-1), atol=1e-4)
    z1 = ops.rsample_from_noise(loc, torch.ones(64, 1, dtype=torch.float64), v)
    assert torch.allclose(z1, loc, atol=2e-3)        # sqrt(clamp(1 - t^2, 1e-7)) leaves a 3e-4 tangential component upstream too


def test_pin_geoopt_script_plumbing(tmp_path):
    """tools/pin_geoopt.py: exits 2 without geoopt (the state of this image: parity of the hyperbolic tail UNPINNED), and
    diffs every function when a module of that name is importable -- exercised here with a stand-in package that re-exports
    the restatement, which must then come out as 'agree' (exit 0)"""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = os.path.join(root, 'tools', 'pin_geoopt.py')
    try:
        import geoopt  # noqa: F401
        have = True
    except Exception:
        have = False
    if not have:
        res = subprocess.run([sys.executable, script], capture_output=True, text=True, cwd=str(tmp_path))
        assert res.returncode == 2 and 'UNPINNED' in res.stdout, res.stdout + res.stderr
    pkg = tmp_path / 'geoopt' / 'manifolds' / 'stereographic'
    pkg.mkdir(parents=True)
    for d in (tmp_path / 'geoopt', tmp_path / 'geoopt' / 'manifolds', pkg):
        (d / '__init__.py').write_text('__version__ = "stand-in"\n')
    (pkg / 'math.py').write_text('from oracle.geoopt_math import *  # noqa\n')
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([str(tmp_path), root]))
    res = subprocess.run([sys.executable, script], capture_output=True, text=True, cwd=str(tmp_path), env=env)
    assert res.returncode == 0 and 'PINNED' in res.stdout and 'DIFF' not in res.stdout, res.stdout + res.stderr


def test_mahalanobis_oracle_matches_reference_functions(golden_dir):
    """tests/golden/mahalanobis_ref.npz: outputs of the reference's own mahalanobis / windows_based_loss_mahalanobis /
    batch_cov_mat_step (utils/eval_utils.py:28-55, models/euclidean_encoder_staticCenter.py:40-46,133-142)"""
    from oracle import mahalanobis as omah
    g = _load(golden_dir, 'mahalanobis_ref.npz')
    for D in (8, 16):
        z, mu, VI = (torch.from_numpy(g[f'{k}{D}']) for k in ('z', 'mu', 'VI'))
        batches = [z[i:i + 256] for i in range(0, z.shape[0], 256)]
        assert torch.equal(sum(omah.batch_cov_mat_step(b, mu) for b in batches), torch.from_numpy(g[f'scatter{D}']))
        assert torch.equal(omah.inv_cov(batches, mu), VI)
        zq = torch.from_numpy(g[f'zq{D}'])
        assert torch.equal(omah.mahalanobis(zq, mu, VI, reduce='none').reshape(-1), torch.from_numpy(g[f'dist{D}']))
        pose = omah.windows_based_loss_mahalanobis(mu, g[f'zq{D}'][:40], VI, g[f'frames{D}'], 60)
        assert np.array_equal(pose, g[f'pose{D}'])


def test_stsvae_oracle_matches_reference_class(golden_dir):
    """tests/golden/stsvae_ref.npz: outputs of the reference's own models/sts/vae.py STSVAE (encode of 'ps' and 'normal', the
    whole 'normal' forward with the reference's noise draw)"""
    g = _load(golden_dir, 'stsvae_ref.npz')
    x = torch.from_numpy(g['x'])
    for dist in ('ps', 'normal'):
        sd = onet.init_state_dict('stsvae', latent_dim=8, seed=2, distribution=dist)
        with torch.no_grad():
            zm, zv = onet.stsvae_encode(x, sd, distribution=dist)
        assert torch.allclose(zm, torch.from_numpy(g[f'z_mean_{dist}']), rtol=1e-5, atol=1e-6)
        assert torch.allclose(zv, torch.from_numpy(g[f'z_var_{dist}']), rtol=1e-5, atol=1e-6)
    with torch.no_grad():
        Z = zm + zv * torch.from_numpy(g['eps_normal'])
        xh = onet.stsae_decode(Z, sd, x.shape)
    assert torch.allclose(Z, torch.from_numpy(g['z_normal']), rtol=1e-5, atol=1e-6)
    assert torch.allclose(xh, torch.from_numpy(g['xhat_normal']), rtol=1e-4, atol=1e-5)
    q = torch.distributions.Normal(zm, zv)
    kl = torch.distributions.kl.kl_divergence(q, torch.distributions.Normal(torch.zeros_like(zm), torch.ones_like(zv))).sum(-1).mean()
    assert abs(float(kl) - float(g['kl_normal'])) < 1e-5 * abs(float(g['kl_normal']))
