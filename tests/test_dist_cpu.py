"""CPU, gloo, world_size 2: the host-side multi-rank plumbing (sharding, center all-reduce, flat gradient
bucket, score gather).  The CUDA kernels are not involved -- these are the collectives around them."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from coskad_b200.pipeline import shard_range


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from coskad_b200 import dist as cdist
    out = {}
    # center accumulators add across shards
    n = 1001
    lo, hi = shard_range(n, rank, world)
    g = torch.Generator().manual_seed(0)
    z = torch.randn(n, 16, generator=g, dtype=torch.float64)
    acc = torch.zeros(18, dtype=torch.float64)
    acc[:16] = z[lo:hi].sum(0)
    acc[17] = hi - lo
    cdist.allreduce_center_acc(acc)
    out['acc_ok'] = bool(torch.allclose(acc[:16], z.sum(0))) and float(acc[17]) == n
    # flat gradient bucket == mean of the per-rank gradients
    torch.manual_seed(1)
    lin = torch.nn.Sequential(torch.nn.Linear(5, 3), torch.nn.PReLU())
    cdist.broadcast_module_(lin)
    x = torch.randn(8, 5, generator=torch.Generator().manual_seed(10 + rank))
    lin(x).pow(2).sum().backward()
    local = [p.grad.clone() for p in lin.parameters()]
    gathered = [None] * world
    dist.all_gather_object(gathered, local)
    cdist.FlatGradBucket(lin.parameters()).allreduce_()
    mean = [sum(gr[i] for gr in gathered) / world for i in range(len(local))]
    out['grad_ok'] = all(torch.allclose(p.grad, m) for p, m in zip(lin.parameters(), mean))
    w0 = [p.detach().clone() for p in lin.parameters()]
    gathered = [None] * world
    dist.all_gather_object(gathered, w0)
    out['bcast_ok'] = all(torch.equal(a, b) for a, b in zip(gathered[0], gathered[1]))
    # score gather restores dataset order, ragged last shard
    scores = torch.arange(n, dtype=torch.float32)
    got = cdist.gather_rows(scores[lo:hi].clone(), n)
    out['gather_ok'] = bool(torch.equal(got, scores))
    lat = torch.arange(n * 3, dtype=torch.float32).view(n, 3)
    out['gather2d_ok'] = bool(torch.equal(cdist.gather_rows(lat[lo:hi].clone(), n), lat))
    q.put((rank, out))
    dist.destroy_process_group()


def test_two_rank_plumbing_gloo():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in range(2):
        assert all(res[r].values()), res


@pytest.mark.parametrize('n,world', [(0, 2), (1, 2), (7, 2), (8, 4), (1001, 8), (16777216, 8)])
def test_shard_range_partitions(n, world):
    spans = [shard_range(n, r, world) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == n
    for (a, b), (c, d) in zip(spans[:-1], spans[1:]):
        assert b == c and a <= b
    assert sum(b - a for a, b in spans) == n


def test_config_prefix_namespaces(tmp_path):
    import argparse
    from coskad_b200 import config as ccfg
    ns = argparse.Namespace(dataset_seg_len=12, dataset_batch_size=64, opt_lr=1e-3, ae_epochs=50, debug=True,
                            exp_dir=str(tmp_path), dataset_choice='synthetic', dir_name='x')
    args, ae_args, ae2, res, opt = ccfg.init_sub_args(ns)
    assert ae_args.seg_len == 12 and ae_args.batch_size == 64 and opt.lr == 1e-3
    assert args.ae_epochs == 10                   # debug forces 10 epochs (utils/argparser.py:11-12)
    assert os.path.isdir(args.ckpt_dir)


def test_compat_shims_register_reference_import_paths():
    import coskad_b200.compat as compat
    compat.install()
    from models.stse.stse_hidden_hypersphere import STSE
    from models.stsae.stsae_hidden_hypersphere import STSAE
    m = STSE(c_in=2, h_dim=64, latent_dim=16, n_frames=12, dropout=0.0, n_joints=17, channels=[32, 16, 32],
             projector='linear', encoder_type='STS_GCN')
    assert m.latent_dim == 16 and 'encoder.model.0.gcn.A' in m.state_dict() and m.c.shape == (16,)
    a = STSAE(c_in=2, h_dim=64, latent_dim=8, n_frames=12, dropout=0.0, n_joints=17, channels=[32, 16, 32])
    assert 'rev_btlnk.weight' in a.state_dict()
    import geoopt.manifolds.stereographic.math as gmath
    assert all(hasattr(gmath, f) for f in ('expmap0', 'project', 'dist', 'dist0', 'weighted_midpoint'))


class _TinyLit(torch.nn.Module):
    """a LightningModule-shaped toy (torch ops only) to drive trainer.TrainStep on the CPU"""

    def __init__(self):
        super().__init__()
        torch.manual_seed(5)
        self.net = torch.nn.Sequential(torch.nn.Linear(6, 4), torch.nn.PReLU(), torch.nn.Linear(4, 1))

    def training_step(self, batch, batch_idx):
        x, y = batch
        return torch.nn.functional.mse_loss(self.net(x), y)


def _step_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from coskad_b200 import dist as cdist
    from coskad_b200.trainer import TrainStep
    out = {}
    # sharded loader: rank r takes batches r, r + world, ..; every rank the same number (the ragged tail is dropped)
    loader = [[torch.full((2, 6), float(i)), torch.full((2, 1), float(i))] for i in range(7)]
    mine = [int(b[0][0, 0]) for b in cdist.ShardedLoader(loader)]
    out['shard_ok'] = mine == [i for i in range(6) if i % world == rank] and len(cdist.ShardedLoader(loader)) == 3
    # data-parallel eager step through the attached flat bucket == single-process step on the averaged gradient
    g = torch.Generator().manual_seed(100)
    xs, ys = torch.randn(world, 8, 6, generator=g), torch.randn(world, 8, 1, generator=g)
    lit = _TinyLit()
    cdist.broadcast_module_(lit)
    opt = torch.optim.SGD(lit.parameters(), lr=0.1)
    bucket = cdist.FlatGradBucket(lit.parameters()).attach()
    out['views_ok'] = bucket.attached() and all(p.grad.data_ptr() >= bucket.flat.data_ptr() for p in lit.parameters())
    ts = TrainStep(lit, opt, bucket, torch.device('cpu'))
    for _ in range(3):
        ts.eager([xs[rank], ys[rank]], 0)
    out['still_attached'] = bucket.attached()          # zero_() instead of zero_grad(set_to_none=True) keeps the views
    ref = _TinyLit()
    ropt = torch.optim.SGD(ref.parameters(), lr=0.1)
    for _ in range(3):
        ropt.zero_grad()
        sum(ref.training_step([xs[r], ys[r]], 0) for r in range(world)).div(world).backward()
        ropt.step()
    out['step_ok'] = all(torch.allclose(a, b, rtol=1e-5, atol=1e-7) for a, b in zip(lit.parameters(), ref.parameters()))
    flat = torch.cat([p.detach().reshape(-1) for p in lit.parameters()])
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    out['replicas_identical'] = all(torch.equal(gathered[0], t) for t in gathered[1:])
    q.put((rank, out))
    dist.destroy_process_group()


def test_data_parallel_trainstep_gloo():
    """world_size 2 on the CPU: the sharded loader, the attached flat gradient bucket (in-place all-reduce, DDP averaging) and
    trainer.TrainStep's eager step -- the host logic of BASELINE configs[4]"""
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_step_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in range(2):
        assert all(res[r].values()), res
