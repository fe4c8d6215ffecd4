"""Multi-GPU check (run under torchrun on N B200s; not collected by pytest):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 \
        tests/multi_gpu_check.py

(1) window-sharded eval: all-gathered scores are bit-identical to the single-GPU scores of the whole set;
(2) center: finalize(all-reduce of per-shard float64 partial sums) == single-GPU center of the union;
(3) data-parallel training step: after the flat NCCL all-reduce every rank holds the mean of the per-rank
    gradients (per-rank BatchNorm statistics, like the reference's DDP without SyncBN).
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
    ok = same and err_c < 1e-5 and err_g < 1e-6
    res = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(res, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f'multi_gpu_check world={world}: scores_bit_identical={same} center_rel_err={err_c:.2e} '
              f'grad_allreduce_rel_err={err_g:.2e} -> {"OK" if float(res) == 1.0 else "FAIL"}')
    dist.destroy_process_group()
    sys.exit(0 if float(res) == 1.0 else 1)


if __name__ == '__main__':
    main()
