"""Pin the oracle against the REAL reference and write the golden fixtures under tests/golden/.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):
    python oracle/gen_golden.py
What it does
  1. imports the reference's own modules (models.sts.ae.STSE / STSAE, utils.hyper_math,
     utils.model_utils, utils.eval_utils) from /root/reference -- third-party imports that are
     missing here (matplotlib, geoopt) are stubbed with empty modules, and ``Tensor.cuda`` is made
     a no-op, ONLY so that the reference's own aggregation / scoring functions can run on CPU;
  2. checks every oracle restatement against them (asserts);
  3. stores small input/output vectors of the REFERENCE (not of the oracle) as .npz fixtures.
The reference ships no tests or golden vectors of its own, so these
reference-generated vectors are the pin.  geoopt and power_spherical are not importable and not
vendored: oracle.geoopt_math / oracle.power_spherical stay "parity unpinned".
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = '/root/reference'
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
GOLD = os.path.join(ROOT, 'tests', 'golden')


def _stub(name: str) -> types.ModuleType:
    mod = types.ModuleType(name)
    sys.modules[name] = mod
    return mod


_ONLY = ''


def _savez(path: str, **arrays) -> None:
    """np.savez, restricted to ``<only>_ref.npz`` under --only (the other fixtures stay byte-identical on disk)"""
    if _ONLY and os.path.basename(path) != f'{_ONLY}_ref.npz':
        return
    np.savez(path, **arrays)


def _ref_function(path: str, name: str):
    """one top-level function of a reference file that cannot be imported as a module here (its module imports Lightning):
    compiled from the reference's own source at generation time -- nothing is copied into this repository"""
    import ast
    src = open(path).read()
    node = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == name)
    ns = {'torch': torch, 'np': np}
    exec(compile(ast.Module(body=[node], type_ignores=[]), path, 'exec'), ns)
    return ns[name]


def mahalanobis_fixture(ref_eu) -> None:
    """distance 'mahalanobis': utils/eval_utils.py:28-55 + models/euclidean_encoder_staticCenter.py:40-46,133-142"""
    from oracle import mahalanobis as omah
    ref_cov_step = _ref_function(os.path.join(REF, 'models', 'euclidean_encoder_staticCenter.py'), 'batch_cov_mat_step')
    g = torch.Generator().manual_seed(21)
    out = {}
    for D in (8, 16):
        A = torch.randn(D, D, generator=g) * 0.3 + torch.eye(D)
        z = torch.randn(700, D, generator=g) @ A.T * 0.2 + 0.05          # correlated latents
        z[5] = z[6]                                                       # a repeated row
        mu = z.mean(0)
        scat_ref = sum(ref_cov_step(z[i:i + 256], mu) for i in range(0, 700, 256))
        scat_or = sum(omah.batch_cov_mat_step(z[i:i + 256], mu) for i in range(0, 700, 256))
        assert torch.equal(scat_ref, scat_or), 'oracle.mahalanobis.batch_cov_mat_step differs from the reference'
        VI = torch.inverse(scat_ref / (700 - 1))                          # compute_inv_cov_mat :142
        assert torch.equal(VI, omah.inv_cov([z[i:i + 256] for i in range(0, 700, 256)], mu))
        zq = torch.cat([z[:200], mu.view(1, -1)])                         # the center itself: distance 0
        d_ref = ref_eu.mahalanobis(zq, mu, VI, reduce='none')
        assert torch.equal(d_ref, omah.mahalanobis(zq, mu, VI, reduce='none'))
        assert torch.equal(ref_eu.mahalanobis(zq, mu, VI), omah.mahalanobis(zq, mu, VI))
        frames = (np.arange(12)[None, :] + np.arange(1, 41)[:, None]).astype(np.int64)
        pose_ref = ref_eu.windows_based_loss_mahalanobis(mu, zq[:40].numpy(), VI, frames, 60)
        assert np.array_equal(pose_ref, omah.windows_based_loss_mahalanobis(mu, zq[:40].numpy(), VI, frames, 60))
        out.update({f'z{D}': z.numpy(), f'mu{D}': mu.numpy(), f'scatter{D}': scat_ref.numpy(), f'VI{D}': VI.numpy(),
                    f'zq{D}': zq.numpy(), f'dist{D}': d_ref.reshape(-1).numpy(), f'pose{D}': pose_ref, f'frames{D}': frames})
    _savez(os.path.join(GOLD, 'mahalanobis_ref.npz'), **out)


def vae_fixture() -> None:
    """the REAL models/sts/vae.py STSVAE with ``power_spherical`` stubbed (un-vendored, only needed by reparameterize of 'ps'):
    encode() of both distributions (fc_mean / normalisation / softplus + 1, vae.py:63-91) and, for 'normal', the whole
    forward (Normal rsample, decode, vae.py:107-132) plus the KL term of models/spherical_vae.py:89-90"""
    from oracle import stsgcn as onet
    ps = _stub('power_spherical')
    psd = _stub('power_spherical.distributions')
    ps.__path__ = []
    ps.distributions = psd
    psd.PowerSpherical = psd.HypersphericalUniform = type('Unavailable', (), {})
    import models.sts.vae as ref_vae          # noqa: E402  (reference)
    x = onet.synth_windows(12, seed=999)
    out = {'x': x.numpy()}
    for dist in ('ps', 'normal'):
        sd = onet.init_state_dict('stsvae', latent_dim=8, seed=2, distribution=dist)
        ref = ref_vae.STSVAE(input_dim=2, layer_channels=[32, 16, 32], hidden_dimension=64, latent_dim=8, n_frames=12,
                             n_joints=17, encoder_type='sts_gcn', projector='linear', distance='euclidean', dropout=0.0,
                             # upstream defect: STSAE hands (device, bias) to STSE in swapped order (models/sts/ae.py:196), so
                             # STSVAE.build_model calls torch.tensor(0, device=<the bias flag>) (vae.py:60) and cannot be
                             # constructed with the defaults; a torch.device in the ``bias`` slot ends up as STSE.device and the
                             # truthy 'cpu' string in the ``device`` slot as the Linear layers' bias flag -- the arithmetic is
                             # the unmodified reference's
                             bias=torch.device('cpu'), device='cpu', distribution=dist)
        ref.load_state_dict(sd, strict=True)
        ref.eval()
        with torch.no_grad():
            zm, zv = ref.encode(x)
            zm_o, zv_o = onet.stsvae_encode(x, sd, distribution=dist)
            assert torch.equal(zm, zm_o) and torch.equal(zv, zv_o), f'oracle stsvae_encode differs from the reference ({dist})'
            out[f'z_mean_{dist}'], out[f'z_var_{dist}'] = zm.numpy(), zv.numpy()
            if dist == 'normal':
                torch.manual_seed(7)
                Z, Xh, (q, p, zv2) = ref(x)
                out['eps_normal'] = ((Z - zm) / zv).numpy()
                out['z_normal'], out['xhat_normal'] = Z.numpy(), Xh.numpy()
                out['kl_normal'] = torch.distributions.kl.kl_divergence(q, p).sum(-1).mean().numpy()
    _savez(os.path.join(GOLD, 'stsvae_ref.npz'), **out)


def main(only: str = '') -> None:
    global _ONLY
    _ONLY = only
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(4)
    from oracle import aggregate as oagg
    from oracle import geoopt_math as ogm
    from oracle import hyper_math as ohm
    from oracle import stsgcn as onet

    # ---- the real reference modules -------------------------------------------------------------
    import models.sts.ae as ref_ae            # noqa: E402  (reference)
    import utils.hyper_math as ref_hm         # noqa: E402
    import utils.model_utils as ref_mu        # noqa: E402

    # -- STSE (hyperbolic / euclidean encoder configs: latent 16)
    sd = onet.init_state_dict('stse', latent_dim=16, seed=0)
    ref = ref_ae.STSE(input_dim=2, layer_channels=[32, 16, 32], hidden_dimension=64, latent_dim=16, n_frames=12,
                      n_joints=17, encoder_type='sts_gcn', projector='linear', distance='euclidean', dropout=0.0)
    missing = ref.load_state_dict(sd, strict=True)
    ref.eval()
    x = onet.synth_windows(10, seed=999)
    with torch.no_grad():
        z_ref = ref(x)
        h1_ref = ref.encoder.model[0](x)
        z_or = onet.stse_forward(x, sd)
        h1_or = onet.st_gcnn_layer(x, sd, 'encoder.model.0')
    assert torch.equal(z_ref, z_or) or (z_ref - z_or).abs().max() < 1e-6, (z_ref - z_or).abs().max()
    assert (h1_ref - h1_or).abs().max() < 1e-6
    # eval-mode fold used by the CUDA path
    w1, w2, b = onet.fold_layer_eval(sd, 'encoder.model.0')
    g = onet.graph_contract(x, sd['encoder.model.0.gcn.A'], sd['encoder.model.0.gcn.T'])
    pre = torch.einsum('oc,nctv->notv', w1, g) + torch.einsum('oc,nctv->notv', w2, x) + b[None, :, None, None]
    h1_fold = torch.nn.functional.prelu(pre, sd['encoder.model.0.prelu.weight'])
    assert (h1_fold - h1_ref).abs().max() < 2e-5
    _savez(os.path.join(GOLD, 'stse_ref.npz'), x=x.numpy(), z=z_ref.numpy(), h1=h1_ref.numpy(),
             sd_checksum=np.float64(sum(float(v.double().sum()) for k, v in sd.items() if v.is_floating_point())))

    # training-mode forward + backward of the reference (loss = mean of z^2 as a stand-in upstream grad)
    ref.train()
    xt = onet.synth_windows(64, seed=7)
    zt = ref(xt)
    loss = (zt ** 2).mean()
    loss.backward()
    grads = {k: p.grad.detach().clone() for k, p in ref.named_parameters()}
    new_stats = {}
    zt_or = onet.stse_forward(xt, sd, training=True, new_stats=new_stats)
    assert (zt - zt_or).abs().max() < 1e-5
    ref_sd_after = ref.state_dict()
    for k, v in new_stats.items():
        assert (ref_sd_after[k] - v).abs().max() < 1e-6, k
    _savez(os.path.join(GOLD, 'stse_train_ref.npz'), x=xt.numpy(), z=zt.detach().numpy(),
             **{'grad.' + k: v.numpy() for k, v in grads.items() if k in (
                 'encoder.model.0.gcn.A', 'encoder.model.0.gcn.T', 'encoder.model.3.tcn.0.weight',
                 'encoder.model.1.tcn.1.weight', 'encoder.model.2.prelu.weight', 'btlnk.bias',
                 'encoder.model.1.residual.0.bias')},
             **{'stat.' + k: ref_sd_after[k].numpy() for k in ('encoder.model.0.tcn.1.running_mean',
                                                               'encoder.model.3.residual.1.running_var')})
    # calc_reg_loss
    reg_ref = ref_mu.calc_reg_loss(ref)
    reg_or = onet.calc_reg_loss(list(ref.named_parameters()))
    assert abs(float(reg_ref) - float(reg_or)) < 1e-6 * abs(float(reg_ref))

    # -- STSAE (euclidean auto-encoder config: latent 8)
    sd_ae = onet.init_state_dict('stsae', latent_dim=8, seed=1)
    ref_aem = ref_ae.STSAE(input_dim=2, layer_channels=[32, 16, 32], hidden_dimension=64, latent_dim=8, n_frames=12,
                           n_joints=17, encoder_type='sts_gcn', projector='linear', distance='euclidean', dropout=0.0)
    ref_aem.load_state_dict(sd_ae, strict=True)
    ref_aem.eval()
    with torch.no_grad():
        z_ae, xh_ae = ref_aem(x)
        z_ae_or, xh_ae_or = onet.stsae_forward(x, sd_ae)
    assert (z_ae - z_ae_or).abs().max() < 1e-6 and (xh_ae - xh_ae_or).abs().max() < 1e-6
    _savez(os.path.join(GOLD, 'stsae_ref.npz'), x=x.numpy(), z=z_ae.numpy(), xhat=xh_ae.numpy())

    # ---- utils/hyper_math.py (the pinned geometry flavour) ----------------------------------------
    g_ = torch.Generator().manual_seed(5)
    u = torch.randn(64, 16, generator=g_) * torch.logspace(-3, 0.7, 64)[:, None]
    cen = torch.randn(16, generator=g_) * 0.1
    e_ref = ref_hm.expmap0(u, c=1.0)
    p_ref = ref_hm.project(e_ref, c=1.0)
    d_ref = ref_hm.dist(p_ref, cen.expand_as(p_ref), c=1.0)
    m_ref = ref_hm.poincare_mean(p_ref[:40] * 0.5, dim=0, c=1.0)
    assert torch.allclose(e_ref, ohm.expmap0(u)) and torch.allclose(p_ref, ohm.project(e_ref))
    assert torch.allclose(d_ref, ohm.dist(p_ref, cen.expand_as(p_ref)), rtol=1e-6, atol=1e-7)
    assert torch.allclose(m_ref, ohm.poincare_mean(p_ref[:40] * 0.5), rtol=1e-6, atol=1e-7)
    _savez(os.path.join(GOLD, 'geometry_hyper_math.npz'), u=u.numpy(), center=cen.numpy(), expmap0=e_ref.numpy(),
             project=p_ref.numpy(), dist=d_ref.numpy(), mean=m_ref.numpy())
    # restated geoopt (UNPINNED) -- stored so the CUDA tests have fixed vectors; these are oracle outputs
    k = torch.tensor(-1.)
    e_g = ogm.expmap0(u, k=k)
    p_g = ogm.project(e_g, k=k)
    _savez(os.path.join(GOLD, 'geometry_geoopt_restated.npz'), u=u.numpy(), center=cen.numpy(), expmap0=e_g.numpy(),
             project=p_g.numpy(), dist=ogm.dist(p_g, cen, k=k).numpy(), dist_cx=ogm.dist(cen, p_g, k=k).numpy(),
             dist0=ogm.dist0(p_g, k=k).numpy(), midpoint=ogm.weighted_midpoint(p_g, k=k).numpy())

    # ---- utils/eval_utils.py: the real aggregation functions --------------------------------------
    plt = _stub('matplotlib.pyplot')
    _stub('matplotlib').pyplot = plt
    geo = _stub('geoopt'); man = _stub('geoopt.manifolds'); ste = _stub('geoopt.manifolds.stereographic')
    gm = _stub('geoopt.manifolds.stereographic.math')
    geo.manifolds = man; man.stereographic = ste; ste.math = gm
    for name in ('expmap0', 'project', 'dist', 'dist0', 'weighted_midpoint'):
        setattr(gm, name, getattr(ogm, name))          # restated geoopt stands in for the missing package
    torch.Tensor.cuda = lambda self, *a, **kw: self    # CPU container: make .cuda() a no-op for the reference
    import utils.eval_utils as ref_eu                  # noqa: E402

    trans, meta, frames, clips, gts = oagg.synth_dataset(n_clips=4, seed=3, num_transform=2)
    rng = np.random.default_rng(11)
    hidden = (rng.standard_normal((len(trans), 16)) * 0.3).astype(np.float32)
    hidden[rng.random(len(trans)) < 0.02] = 0.0          # exact-zero scores ("absent" windows)
    c_t = torch.zeros(16)
    loss_fn = torch.nn.MSELoss(reduction='none')
    # per-person matrices from the real windows_based_loss_hy (euclidean branch), then the eval loop
    ref_curves, or_curves = [], []
    for t in range(2):
        ct = trans == t
        for scene, clip, F in clips:
            cc = ct & (meta[:, 0] == scene) & (meta[:, 1] == clip)
            per_person = []
            for fig in sorted(set(meta[cc][:, 2])):
                cf = cc & (meta[:, 2] == fig)
                lm = ref_eu.windows_based_loss_hy(c_t, hidden[cf], frames[cf], F, loss_fn, False)
                lm = np.where(lm == 0.0, np.nan, lm)
                with np.errstate(all='ignore'):
                    import warnings
                    with warnings.catch_warnings():
                        warnings.simplefilter('ignore')
                        fl = np.nanmean(lm, 0)
                per_person.append(np.where(np.isnan(fl), 0, fl))
            clip_score = np.amax(np.stack(per_person, 0), 0)
            ref_curves.append(ref_eu.score_process(clip_score.copy()))
    scores = ((hidden.astype(np.float32) - 0.0) ** 2)
    scores = torch.mean(loss_fn(c_t, torch.from_numpy(hidden)), dim=-1).numpy()
    agg = oagg.aggregate_dataset(scores, trans, meta, frames, clips, 2)
    or_curves = [c for t in range(2) for c in agg[t]]
    for a, b_ in zip(ref_curves, or_curves):
        assert np.array_equal(a, b_), 'oracle.aggregate differs from the reference functions'
    # pad_scores
    fr = np.array([0, 0, 0, .5, .4, 0, 0, 0, 0, 0, .2, .1, 0, 0, 0, 0, 0, 0, 0, 0.])
    assert np.array_equal(ref_eu.pad_scores(fr.copy(), np.zeros(20), 2), oagg.pad_scores(fr.copy(), np.zeros(20), 2))
    _savez(os.path.join(GOLD, 'aggregate_ref.npz'), scores=scores, trans=trans, meta=meta, frames=frames,
             clips=np.asarray(clips, dtype=np.int64), curves=np.concatenate(ref_curves),
             pad_in=fr, pad_out=ref_eu.pad_scores(fr.copy(), np.zeros(20), 2))
    mahalanobis_fixture(ref_eu)
    vae_fixture()
    if only in ('mahalanobis', 'stsvae'):
        return
    # ---- window construction + test-time transforms (utils/dataset_utils.py, utils/preprocessing.py, utils/dataset.py) ----
    from oracle import windows as owin
    import utils.dataset_utils as ref_du      # noqa: E402  (reference)
    import utils.preprocessing as ref_pp      # noqa: E402
    ref_mats = np.stack([t.trans_mat.numpy() for t in ref_du.ae_trans_list], 0)
    assert np.array_equal(ref_mats, owin.ae_trans_mats()), 'oracle.windows affine matrices differ from ae_trans_list'
    rng = np.random.default_rng(5)
    traj = (rng.standard_normal((40, 34)) * 0.4).astype(np.float32)
    traj[rng.random((40, 17)).repeat(2, 1) < 0.05] = 0.0                     # missing joints stay (0, 0)
    Xw, _ = ref_pp._aggregate_rnn_autoencoder_data(traj, input_length=12, input_gap=0, pred_length=0)
    starts = owin.sliding_starts(traj.shape[0], 12, 1)
    assert Xw.shape[0] == len(starts)
    # PoseDatasetRobust.gen_dataset (utils/dataset.py:253-273): reshape (N,T,17,2), conf channel 1.0, transpose to (N,C,T,V)
    segs = np.empty((*Xw.shape[:2], 17, 3)); segs[..., :2] = Xw.reshape(*Xw.shape[:2], 17, 2); segs[..., 2] = 1.0
    segs = np.transpose(segs, (0, 3, 1, 2)).astype(np.float32)
    assert np.array_equal(segs[:, :2], owin.windows_from_rows(traj, starts)), 'oracle.windows window layout differs'
    ref_tr = np.stack([np.stack([t(np.array(w)) for w in segs], 0) for t in ref_du.ae_trans_list], 0)    # [5, N, 3, 12, 17]
    or_tr = np.stack([np.stack([owin.apply_pose_transform(w, m) for w in segs], 0) for m in owin.ae_trans_mats()], 0)
    assert np.array_equal(ref_tr, or_tr), 'oracle.windows apply_pose_transform differs from the reference'
    _savez(os.path.join(GOLD, 'windows_ref.npz'), traj=traj, starts=starts, mats=ref_mats, windows=segs[:, :2],
             transformed=ref_tr[:, :, :2].astype(np.float32))
    print('golden fixtures written to', GOLD)
    for fn in sorted(os.listdir(GOLD)):
        print('  ', fn, os.path.getsize(os.path.join(GOLD, fn)), 'bytes')


if __name__ == '__main__':
    # --only mahalanobis: (re)generate tests/golden/mahalanobis_ref.npz and stop (the other fixtures stay byte-identical)
    main(sys.argv[sys.argv.index('--only') + 1] if '--only' in sys.argv else '')
