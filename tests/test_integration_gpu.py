"""GPU integration: task modules + trainer + entry-point flow on a synthetic dataset; AUC parity to 4 decimals
against the oracle pipeline (oracle geometry + oracle aggregation + the same sklearn call) on the same latents."""
import argparse
import os

import numpy as np
import pytest
import torch

from oracle import aggregate as oagg
from oracle import geoopt_math as ogm
from oracle import stsgcn as onet

pytestmark = pytest.mark.gpu
K = torch.tensor(-1.)


def _args(tmp_path, **kw):
    from coskad_b200 import config as ccfg
    ns = argparse.Namespace(dataset_choice='synthetic', exp_dir=str(tmp_path), dir_name='run', hyperbolic=True,
                            static_center=False, use_decoder=False, use_vae=False, latent_dim=16, ae_epochs=3, opt_lr=1e-3,
                            dataset_batch_size=256, dataset_num_transform=2, projector='linear', validation=True, seed=5,
                            dataset_synthetic_clips=6)
    for k, v in kw.items():
        setattr(ns, k, v)
    return ccfg.init_sub_args(ns)


def _oracle_auc(args, model, ds):
    """reference pipeline on the CPU: oracle STSE eval forward -> geoopt score -> reference aggregation -> AUC"""
    from sklearn.metrics import roc_auc_score
    sd = {k[len('model.'):]: v.detach().cpu() for k, v in model.state_dict().items() if k.startswith('model.')}
    with torch.no_grad():
        z = onet.stse_forward(ds.x, sd)
        c = sd['c']
        if args.hyperbolic:
            s = ogm.dist(ogm.project(ogm.expmap0(z, k=K), k=K), c, k=K)
        else:
            s = torch.mean((c - z) ** 2, dim=-1)
    nt = args.dataset_num_transform
    curves = oagg.aggregate_dataset(s.numpy(), ds.trans.numpy(), ds.meta.numpy(), ds.frames.numpy(), ds.clips, nt)
    gt = np.concatenate([ds.gts[(s_, c_)] for s_, c_, _ in ds.clips])
    pds = np.mean(np.stack([np.concatenate(curves[t]) for t in range(nt)], 0), 0)
    return float(roc_auc_score(gt, pds)), s


@pytest.mark.parametrize('hyperbolic,static_center', [(True, False), (True, True), (False, False)])
def test_train_eval_auc_parity(tmp_path, hyperbolic, static_center):
    from coskad_b200 import tasks
    from coskad_b200.data import get_dataset_and_loader
    from coskad_b200.trainer import Trainer, load_checkpoint
    import eval_COSKAD
    torch.manual_seed(0)
    args, ae_args, *_ = _args(tmp_path, hyperbolic=hyperbolic, static_center=static_center)
    train_ds, train_loader = get_dataset_and_loader(ae_args, 'train')
    test_ds, test_loader = get_dataset_and_loader(ae_args, 'test')
    args.gt_table = (test_ds.clips, test_ds.gts)
    model = tasks.select_task(args)(args)
    trainer = Trainer(max_epochs=args.ae_epochs, ckpt_dir=args.ckpt_dir, monitor='validation_auc', mode='max', save_top_k=2)
    trainer.fit(model, train_loader, test_loader)
    h = trainer.history
    assert len(h) == 3 and all(np.isfinite(e['train_loss_mean']) for e in h), h
    assert h[-1]['loss'] < h[0]['loss'], h          # last-step loss of the epoch goes down under Adam
    assert all(np.isfinite(e['validation_auc']) for e in h)
    # center: single-process reference semantics on the union of the data
    assert model.model.c.shape == (16,) and bool(torch.isfinite(model.model.c).all())
    if not static_center:
        assert len(model.centers) == 4 if hyperbolic else True
    # checkpoints: top-2, reference key names
    cks = trainer.best_checkpoints
    assert 1 <= len(cks) <= 2
    sd = torch.load(cks[0], weights_only=False)['state_dict']
    assert 'model.encoder.model.0.gcn.A' in sd and 'model.btlnk.weight' in sd and 'model.c' in sd
    # eval entry-point flow on a fresh module from the checkpoint
    model2 = tasks.select_task(args)(args)
    auc, per_t, curves = eval_COSKAD.evaluate(args, model2, (test_ds, test_loader), ckpt_path=cks[0])
    ref_auc, ref_scores = _oracle_auc(args, model2, test_ds)
    assert round(auc, 4) == round(ref_auc, 4), (auc, ref_auc)
    assert set(per_t) == {0, 1}


def test_mahalanobis_distance_task(tmp_path):
    """distance: 'mahalanobis' on the Euclidean static-center encoder (models/euclidean_encoder_staticCenter.py:125-142,182-185,
    268-270): inv_cov_matrix after setup / every epoch = inverse sample covariance of the training latents around the center,
    the loss falls, and the eval scores are the reference's Mahalanobis distances (AUC to 4 decimals)"""
    from sklearn.metrics import roc_auc_score
    from coskad_b200 import tasks
    from coskad_b200.data import get_dataset_and_loader
    from coskad_b200.trainer import Trainer
    from oracle import mahalanobis as omah
    torch.manual_seed(0)
    args, ae_args, *_ = _args(tmp_path, hyperbolic=False, static_center=True, distance='mahalanobis', ae_epochs=0)
    train_ds, train_loader = get_dataset_and_loader(ae_args, 'train')
    test_ds, test_loader = get_dataset_and_loader(ae_args, 'test')
    args.gt_table = (test_ds.clips, test_ds.gts)
    model = tasks.select_task(args)(args)
    assert model.distance == 'mahalanobis' and 'model.inv_cov_matrix' in model.state_dict()
    Trainer(max_epochs=0, verbose=False).fit(model, train_loader)              # setup('fit') only
    sd = {k[len('model.'):]: v.detach().cpu() for k, v in model.state_dict().items() if k.startswith('model.')}
    with torch.no_grad():
        z_tr = onet.stse_forward(train_ds.x, sd)
    vi_ref = omah.inv_cov([z_tr[i:i + 256] for i in range(0, z_tr.shape[0], 256)], sd['c'])
    vi = model.model.inv_cov_matrix.cpu()
    assert float((vi - vi_ref).abs().max()) <= 5e-3 * float(vi_ref.abs().max()), (vi, vi_ref)
    # training: three epochs, the Mahalanobis loss goes down and the matrix is refreshed from the epoch's latents
    args.ae_epochs = 3
    trainer = Trainer(max_epochs=3, verbose=False)
    trainer.fit(model, train_loader, test_loader)
    h = trainer.history
    assert len(h) == 3 and all(np.isfinite(e['train_loss_mean']) for e in h) and h[-1]['loss'] < h[0]['loss'], h
    assert not torch.equal(model.model.inv_cov_matrix.cpu(), vi)
    # eval scores vs the oracle pipeline on the same weights / center / matrix
    sd = {k[len('model.'):]: v.detach().cpu() for k, v in model.state_dict().items() if k.startswith('model.')}
    model.eval()
    with torch.no_grad():
        z_te = onet.stse_forward(test_ds.x, sd)
        s_ref = omah.mahalanobis(z_te, sd['c'], sd['inv_cov_matrix'], reduce='none').view(-1)
        s = model.window_scores(model.model(test_ds.x.cuda()))
    assert float(((s.cpu() - s_ref).abs() / s_ref.abs()).max()) <= 1e-4
    nt = args.dataset_num_transform
    auc = model.post_processing(model.model(test_ds.x.cuda()), test_ds.trans, test_ds.meta, test_ds.frames, validation=False)
    curves = oagg.aggregate_dataset(s_ref.numpy(), test_ds.trans.numpy(), test_ds.meta.numpy(), test_ds.frames.numpy(),
                                    test_ds.clips, nt)
    gt = np.concatenate([test_ds.gts[(s_, c_)] for s_, c_, _ in test_ds.clips])
    pds = np.mean(np.stack([np.concatenate(curves[t]) for t in range(nt)], 0), 0)
    assert round(auc, 4) == round(float(roc_auc_score(gt, pds)), 4)


def test_center_init_matches_reference_semantics(tmp_path):
    """setup('fit'): c = weighted_midpoint(project(expmap0(z))) over ALL training windows (hyperbolic_encoder.py:101-123)"""
    from coskad_b200 import tasks
    from coskad_b200.data import get_dataset_and_loader
    from coskad_b200.trainer import Trainer
    torch.manual_seed(1)
    args, ae_args, *_ = _args(tmp_path, ae_epochs=0)
    ds, loader = get_dataset_and_loader(ae_args, 'train')
    model = tasks.LitEncoder(args)
    Trainer(max_epochs=0, verbose=False).fit(model, loader)
    sd = {k[len('model.'):]: v.detach().cpu() for k, v in model.state_dict().items() if k.startswith('model.')}
    with torch.no_grad():
        z = onet.stse_forward(ds.x, sd)
        c_ref = ogm.weighted_midpoint(ogm.project(ogm.expmap0(z, k=K), k=K), k=K)
    assert torch.allclose(model.model.c.cpu(), c_ref, rtol=1e-4, atol=1e-6), (model.model.c.cpu(), c_ref)


def test_autoencoder_task_runs(tmp_path):
    from coskad_b200 import tasks
    from coskad_b200.data import get_dataset_and_loader
    from coskad_b200.trainer import Trainer
    torch.manual_seed(2)
    args, ae_args, *_ = _args(tmp_path, hyperbolic=False, use_decoder=True, latent_dim=8, ae_epochs=2)
    _, train_loader = get_dataset_and_loader(ae_args, 'train')
    test_ds, test_loader = get_dataset_and_loader(ae_args, 'test')
    args.gt_table = (test_ds.clips, test_ds.gts)
    model = tasks.select_task(args)(args)
    assert isinstance(model, tasks.LitAutoEncoder)
    tr = Trainer(max_epochs=2, verbose=False).fit(model, train_loader, test_loader)
    assert all(np.isfinite(e['train_loss_mean']) for e in tr.history)
    assert 0.0 <= tr.history[-1]['validation_auc'] <= 1.0


@pytest.mark.parametrize('use_decoder', [False, True])
def test_cuda_graph_training_matches_eager(tmp_path, use_decoder):
    """Trainer(cuda_graph=True): the replayed step (training_step + backward + Adam, dynamic center, LR schedule) follows
    the eager trajectory EXACTLY: the training path has no floating-point atomics (fixed-order two-stage reductions) and
    both modes run the capturable optimizer, so the loss trajectories are bit-identical"""
    from coskad_b200 import tasks
    from coskad_b200.data import get_dataset_and_loader
    from coskad_b200.trainer import Trainer
    res = []
    for graph in (False, True):
        torch.manual_seed(3)
        kw = dict(hyperbolic=False, use_decoder=True, latent_dim=8) if use_decoder else {}
        args, ae_args, *_ = _args(tmp_path, ae_epochs=3, validation=False, dataset_batch_size=48, **kw)
        _, loader = get_dataset_and_loader(ae_args, 'train')
        model = tasks.select_task(args)(args)
        tr = Trainer(max_epochs=3, verbose=False, cuda_graph=graph).fit(model, loader)
        sizes = [int(b[0].shape[0]) for b in loader]
        expected = n_eager = 0
        captured = False
        for _ in range(3):                       # every full-size batch after 3 eager warm-up steps is a replay
            for sz in sizes:
                captured = captured or (n_eager >= 3 and sz == loader.batch_size)
                if captured and sz == loader.batch_size:
                    expected += 1
                else:
                    n_eager += 1
        assert expected >= 6 and sizes[-1] != loader.batch_size, sizes      # the ragged last batch stays eager
        assert tr.graph_replays == (expected if graph else 0), (tr.graph_replays, expected, sizes)
        res.append((model, [e['train_loss_mean'] for e in tr.history], [e['loss'] for e in tr.history]))
    (m0, mean0, last0), (m1, mean1, last1) = res
    assert mean1 == mean0, (mean1, mean0)
    assert last1 == last0, (last1, last0)
    assert torch.equal(m1.model.c, m0.model.c)
    if not use_decoder:
        assert len(m1.centers) == len(m0.centers) == 4
        for a, b in zip(m1.centers, m0.centers):
            assert torch.equal(a, b)
    # parameters: compared through the function they define (a conv bias in front of train-mode BatchNorm has an exactly
    # zero gradient, so Adam normalises rounding noise and bias / running_mean drift together without changing the output)
    for (k, a), (_, b) in zip(m1.state_dict().items(), m0.state_dict().items()):
        assert torch.equal(a, b), k
    x = next(iter(loader))[0].cuda()
    with torch.no_grad():
        from coskad_b200 import _lib
        (z1, s1), (z0, s0) = (m.model.eval().encode_score(x, _lib.SCORE_EUCLID) for m in (m1, m0))
    assert torch.equal(z1, z0) and torch.equal(s1, s0), (float((z1 - z0).abs().max()), float((s1 - s0).abs().max()))
