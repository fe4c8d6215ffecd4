"""The drop-in boundary on the B200: the reference's OWN ``LitEncoder`` (models/hyperbolic_encoder.py, the unmodified file
from baseline/_ref) runs setup / training_step / forward on the CUDA hot path through coskad_b200.compat -- shim STSE at
the missing import path, coskad_b200.gmath under the geoopt name -- and agrees with (i) the re-hosted
coskad_b200.tasks.LitEncoder and (ii) the reference's own torch network (models/sts/ae.py STSE on the CPU) followed by
the restated geoopt math."""
import argparse
import copy
import importlib
import os

import pytest
import torch
import yaml

from oracle import geoopt_math as ogm

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, 'baseline', '_ref')
K = torch.tensor(-1.)


def _args(static_center: bool) -> argparse.Namespace:
    with open(os.path.join(REF, 'config/UBnormal/hyperbolic_encoder.yaml')) as f:
        d = yaml.load(f, Loader=yaml.FullLoader)
    d.update(projector='linear', device='cuda', static_center=static_center, dataset_batch_size=256)
    return argparse.Namespace(**d)


class _FakeTrainer:
    """what LitEncoder.setup reads: trainer._data_connector._train_dataloader_source.dataloader() (hyperbolic_encoder.py:95)"""

    def __init__(self, batches):
        from coskad_b200.trainer import _DataConnector
        self._data_connector = _DataConnector(batches)
        self.device = torch.device('cuda', 0)


def _batches(n_batches=3, bsz=256, seed=3):
    from oracle import stsgcn as onet
    x = onet.synth_windows(n_batches * bsz, seed=seed)
    out = []
    for i in range(n_batches):
        xi = x[i * bsz:(i + 1) * bsz]
        out.append([xi, torch.zeros(bsz, dtype=torch.int64), torch.zeros(bsz, 4, dtype=torch.int64),
                    torch.arange(12).repeat(bsz, 1) + 1])
    return out


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'models')), reason='baseline/_ref absent (oracle/install_ref.py)')
@pytest.mark.parametrize('static_center', [False, True])
def test_reference_litencoder_runs_on_the_cuda_path(static_center):
    import coskad_b200.compat as compat
    from coskad_b200 import sts, tasks
    from coskad_b200.synth import randomize_bn_
    compat.install(reference=REF)
    ref_mod = importlib.import_module('models.hyperbolic_encoder')
    assert os.path.realpath(ref_mod.__file__).startswith(os.path.realpath(REF))
    ref_net = importlib.import_module('models.sts.ae')            # the reference's own torch network (CPU leg)

    args = _args(static_center)
    torch.manual_seed(0)
    ref_lit = ref_mod.LitEncoder(copy.copy(args))                 # the reference's class ...
    assert isinstance(ref_lit.model, sts.STSE)                    # ... on the CUDA-backed network
    randomize_bn_(ref_lit.model, 1)
    our_lit = tasks.LitEncoder(copy.copy(args))
    our_lit.load_state_dict(ref_lit.state_dict())
    cpu_net = ref_net.STSE(input_dim=2, layer_channels=[32, 16, 32], hidden_dimension=64, latent_dim=16, n_frames=12,
                           n_joints=17, encoder_type='sts_gcn', projector='linear', distance='euclidean', dropout=0.0)
    cpu_net.load_state_dict({k[len('model.'):]: v.clone() for k, v in ref_lit.state_dict().items()})
    ref_lit.cuda()
    our_lit.cuda()

    batches = _batches()
    ref_lit.trainer = _FakeTrainer(batches)
    our_lit.trainer = _FakeTrainer(batches)

    # ---- setup('fit'): center of the projected embeddings (hyperbolic_encoder.py:85-135) -------------------------------
    ref_lit.setup('fit')
    our_lit.setup('fit')
    cpu_net.eval()
    with torch.no_grad():
        zs = torch.cat([cpu_net(b[0]) for b in batches])
        c_cpu = ogm.weighted_midpoint(ogm.project(ogm.expmap0(zs, k=K), k=K), k=K)
    torch.testing.assert_close(ref_lit.model.c.cpu(), c_cpu, rtol=1e-4, atol=1e-7)
    torch.testing.assert_close(ref_lit.model.c, our_lit.model.c, rtol=1e-6, atol=1e-8)

    # ---- forward (predict / validation, :72-80), eval mode ---------------------------------------------------------------
    ref_lit.eval(); our_lit.eval()
    dev_batch = [t.cuda() for t in batches[1]]
    with torch.no_grad():
        h_ref, tr, meta, fr = ref_lit(dev_batch)
        h_our = our_lit(dev_batch)[0]
        h_cpu = cpu_net(batches[1][0])
    assert torch.equal(h_ref, h_our)                               # same kernel, same weights
    torch.testing.assert_close(h_ref.cpu(), h_cpu, rtol=1e-4, atol=1e-5 * float(h_cpu.abs().max()))
    assert tr is dev_batch[1] and meta is dev_batch[2] and fr is dev_batch[3]

    # ---- training_step (:137-172): the reference differentiates through gmath.expmap0 / project / dist one call at a time
    ref_lit.train(); our_lit.train(); cpu_net.train()
    loss_ref = ref_lit.training_step(dev_batch, 0)
    loss_our = our_lit.training_step(dev_batch, 0)
    loss_ref.backward()
    loss_our.backward()
    calc_reg_loss = importlib.import_module('utils.model_utils').calc_reg_loss      # the reference's own function
    z = cpu_net(batches[1][0])
    x = ogm.project(ogm.expmap0(z, k=K), k=K)
    loss_cpu = ogm.dist(c_cpu, x, k=K).mean() + args.alpha * calc_reg_loss(cpu_net)
    loss_cpu.backward()
    assert abs(float(loss_ref) - float(loss_cpu)) <= 1e-5 * abs(float(loss_cpu)), (float(loss_ref), float(loss_cpu))
    assert abs(float(loss_ref) - float(loss_our)) <= 1e-6 * abs(float(loss_our)), (float(loss_ref), float(loss_our))
    if not static_center:                                          # the reference keeps the projected latents in cumt (:149-153)
        assert ref_lit.cumt.shape == (256, 16) and float(ref_lit.cumt.norm(dim=-1).max()) <= 1 - 4e-3 + 1e-6
    g_ref = dict(ref_lit.model.named_parameters())
    g_our = dict(our_lit.model.named_parameters())
    g_cpu = dict(cpu_net.named_parameters())
    for name in ('btlnk.weight', 'btlnk.bias', 'encoder.model.3.tcn.0.weight', 'encoder.model.0.gcn.A', 'encoder.model.1.gcn.T',
                 'encoder.model.2.prelu.weight', 'encoder.model.0.residual.0.weight'):
        a, b, c_ = g_ref[name].grad, g_our[name].grad, g_cpu[name].grad
        # separate expmap0 / project / dist VJP kernels vs the fused poincare_score backward: fp32 reassociation only
        rel_ab = float((a - b).norm() / b.norm())
        assert rel_ab < 1e-4, f'{name} (reference class vs tasks.LitEncoder): relative Frobenius error {rel_ab:.2e}'
        # vs the reference's torch network on the CPU: a pre-activation within fp32 rounding of zero may take the other PReLU
        # branch in either implementation (tests/test_train_gpu.py pins the branch pattern for the element-wise check), so
        # this cross-implementation check is norm-wise
        rel = float((a.cpu() - c_).norm() / c_.norm())
        assert rel < 1e-3, f'{name} (CUDA vs reference torch): relative Frobenius error {rel:.2e}'

    # ---- training_epoch_end (:175-188): dynamic center update from cumt --------------------------------------------------
    if not static_center:
        ref_lit.training_epoch_end([loss_ref])
        with torch.no_grad():
            c2 = ogm.weighted_midpoint(ref_lit.cumt.cpu(), k=K)
        torch.testing.assert_close(ref_lit.temp.cpu(), c2, rtol=1e-4, atol=1e-7)


def test_gmath_elementwise_vjp_matches_autograd_of_the_restated_formulas():
    """expmap0 / project / dist / cosine / euclid VJP kernels (coskad_geom_map_bwd, coskad_dist_bwd) vs float64 autograd of
    oracle/geoopt_math.py, inside and outside the projection radius"""
    from coskad_b200 import gmath
    g = torch.Generator().manual_seed(11)
    for scale in (0.3, 1.5, 4.0):                                  # 4.0: most rows clipped by project
        u = (torch.randn(300, 16, generator=g) * scale / 4).float()
        cen = (torch.randn(16, generator=g) * 0.05).float()
        w = torch.randn(300, generator=g).float()
        ud = u.clone().cuda().requires_grad_(True)
        x = gmath.project(gmath.expmap0(ud, k=K), k=K)
        (gmath.dist(cen.cuda(), x, k=K) * w.cuda()).sum().backward()
        u64 = u.double().requires_grad_(True)
        x64 = ogm.project(ogm.expmap0(u64, k=K), k=K, eps=4e-3)
        (ogm.dist(cen.double(), x64, k=K) * w.double()).sum().backward()
        ref = u64.grad
        err = (ud.grad.cpu().double() - ref).abs().max() / ref.abs().max()
        assert float(err) < 2e-4, (scale, float(err))
    # two row operands, both requiring grad; and a broadcast operand requiring grad
    a = (torch.randn(64, 16, generator=g) * 0.1).float()
    b = (torch.randn(64, 16, generator=g) * 0.1).float()
    for fn, ref_fn in ((lambda p, q: gmath.dist(p, q, k=K), lambda p, q: ogm.dist(p, q, k=K)),
                       (gmath.euclid_score, lambda p, q: ((q - p) ** 2).mean(-1)),
                       (gmath.cosine_score, lambda p, q: 1 - torch.nn.functional.cosine_similarity(q, p, dim=-1))):
        ad, bd = a.clone().cuda().requires_grad_(True), b.clone().cuda().requires_grad_(True)
        fn(ad, bd).sum().backward()
        a64, b64 = a.double().requires_grad_(True), b.double().requires_grad_(True)
        ref_fn(a64, b64).sum().backward()
        for got, ref in ((ad.grad, a64.grad), (bd.grad, b64.grad)):
            assert float((got.cpu().double() - ref).abs().max() / ref.abs().max()) < 1e-4
        cd = b[0].clone().cuda().requires_grad_(True)
        fn(a.cuda(), cd).sum().backward()
        c64 = b[0].double().requires_grad_(True)
        ref_fn(a.double(), c64).sum().backward()
        assert float((cd.grad.cpu().double() - c64.grad).abs().max() / c64.grad.abs().max()) < 1e-4
    z = torch.randn(50, 8, generator=g).float()
    zd = z.clone().cuda().requires_grad_(True)
    (gmath.l2_normalize(zd) * torch.arange(8.0).cuda()).sum().backward()
    z64 = z.double().requires_grad_(True)
    ((z64 / z64.norm(dim=-1, keepdim=True)) * torch.arange(8.0).double()).sum().backward()
    assert float((zd.grad.cpu().double() - z64.grad).abs().max() / z64.grad.abs().max()) < 1e-4
