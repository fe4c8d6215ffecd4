"""Multi-GPU check (run under torchrun on N B200s; tests/test_multi_gpu.py spawns it with 2 ranks under pytest -m gpu):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 \
        tests/multi_gpu_check.py

(1) window-sharded eval: all-gathered scores are bit-identical to the single-GPU scores of the whole set;
(2) center: finalize(all-reduce of per-shard float64 partial sums) == single-GPU center of the union;
(3) data-parallel training step: after the flat NCCL all-reduce every rank holds the mean of the per-rank
    gradients (per-rank BatchNorm statistics, like the reference's DDP without SyncBN);
(4) the graph-replayed data-parallel step (trainer.TrainStep: graph A -> in-place ncclAllReduce -> graph B) follows the
    eager data-parallel step: same losses, and the replicas stay identical across ranks.
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from coskad_b200 import _lib, dist as cdist, gmath            # noqa: E402
from coskad_b200.losses import calc_reg_loss                  # noqa: E402
from coskad_b200.pipeline import shard_range                  # noqa: E402
from coskad_b200.synth import make_model, synth_windows_      # noqa: E402


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    model = make_model('stse', 16, seed=0, device=dev)
    cdist.broadcast_module_(model)
    N = 100_003
    x = torch.empty((N, 2, 12, 17), device=dev)
    synth_windows_(x, torch.Generator(device=dev).manual_seed(7))     # same seed on every rank -> same global set
    lo, hi = shard_range(N, rank, world)
    # (2) center from shard partials
    z_loc, _ = model.encode_score(x[lo:hi].contiguous())
    acc = gmath.center_accumulator(16, dev)
    gmath.center_partial(gmath.expmap0_project(z_loc), acc, _lib.SCORE_POINCARE)
    cdist.allreduce_center_acc(acc)
    c = gmath.center_finalize(acc, 16, _lib.SCORE_POINCARE)
    z_all, _ = model.encode_score(x)
    c_single = gmath.weighted_midpoint(gmath.expmap0_project(z_all))
    err_c = float((c - c_single).abs().max() / c_single.abs().max())
    # (1) sharded scores
    _, s_loc = model.encode_score(x[lo:hi].contiguous(), _lib.SCORE_POINCARE, center=c_single)
    s_gather = cdist.gather_rows(s_loc, N)
    _, s_single = model.encode_score(x, _lib.SCORE_POINCARE, center=c_single)
    same = bool(torch.equal(s_gather, s_single))
    # (3) one data-parallel training step
    model.train()
    B = 2048
    xb = x[rank * B:(rank + 1) * B].contiguous()
    z = model(xb)
    s, _ = gmath.poincare_score(z, c_single, True)
    loss = s.mean() + 1e-6 * calc_reg_loss(model)
    loss.backward()
    local_grads = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
    gl = [torch.empty_like(local_grads) for _ in range(world)]
    dist.all_gather(gl, local_grads)
    cdist.FlatGradBucket(model.parameters()).allreduce_()
    after = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
    mean = torch.stack(gl).mean(0)
    err_g = float((after - mean).abs().max() / mean.abs().max())
    # (4) graph-replayed data-parallel steps vs eager ones, from the same initial state
    import argparse
    import copy
    from coskad_b200 import config as ccfg, tasks
    from coskad_b200.trainer import TrainStep, _make_capturable
    ns = argparse.Namespace(hyperbolic=True, static_center=False, latent_dim=16, dataset_batch_size=B, projector='linear',
                            ae_epochs=10, opt_lr=1e-3, validation=False)
    args, *_ = ccfg.init_sub_args(ns, make_dirs=False)
    torch.manual_seed(3)
    lit0 = tasks.LitEncoder(args).to(dev)
    cdist.broadcast_module_(lit0)
    lit0.model.c = c_single.clone()
    batches = [[x[(i * world + rank) * B:(i * world + rank + 1) * B].contiguous(), torch.zeros(B, device=dev)] for i in range(7)]
    losses = {}
    finals = {}
    for mode in ('eager', 'graph'):
        lit = copy.deepcopy(lit0)
        lit.temp = c_single.clone()
        lit.train()
        opt = lit.configure_optimizers()['optimizer']
        if mode == 'graph':
            _make_capturable(opt, dev)
        ts = TrainStep(lit, opt, cdist.FlatGradBucket(lit.parameters()).attach(), dev)
        ls = []
        for i, b in enumerate(batches):
            if mode == 'graph' and i == 3:
                ts.capture(b, i)
            ls.append(float(ts.replay(b) if ts.captured else ts.eager(b, i)))
        losses[mode] = ls
        finals[mode] = torch.cat([p.detach().reshape(-1) for p in lit.parameters()])
    err_l = max(abs(a - b) / abs(a) for a, b in zip(losses['eager'], losses['graph']))
    fl = [torch.empty_like(finals['graph']) for _ in range(world)]
    dist.all_gather(fl, finals['graph'])
    replicas_equal = all(bool(torch.equal(fl[0], f)) for f in fl[1:])
    ok = same and err_c < 1e-5 and err_g < 1e-6 and err_l < 5e-3 and replicas_equal
    res = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(res, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f'multi_gpu_check world={world}: scores_bit_identical={same} center_rel_err={err_c:.2e} '
              f'grad_allreduce_rel_err={err_g:.2e} graph_vs_eager_loss_rel_err={err_l:.2e} replicas_identical={replicas_equal} '
              f'-> {"OK" if float(res) == 1.0 else "FAIL"}')
    dist.destroy_process_group()
    sys.exit(0 if float(res) == 1.0 else 1)


if __name__ == '__main__':
    main()
