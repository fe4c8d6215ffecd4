"""torch.distributed plumbing of the hot path (NCCL on B200s, gloo in the CPU tests).

Only three exchanges exist: the center partial sums (D+2 float64), one flat gradient
bucket per step (the reference gets this from Lightning DDP, train_COSKAD.py:78), and the final score
gather.  Windows are sharded in contiguous blocks (pipeline.shard_range); there is no data-path
collective in the encoder.  All messages are latency-bound on NVLink 5 / NVSwitch.
"""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch
import torch.distributed as dist


def is_dist() -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def world() -> int:
    return dist.get_world_size() if is_dist() else 1


def rank() -> int:
    return dist.get_rank() if is_dist() else 0


def allreduce_center_acc(acc: torch.Tensor) -> torch.Tensor:
    """sum the [D+2] float64 center accumulators of all ranks in place (one 136-byte all-reduce for D=16)"""
    if is_dist():
        dist.all_reduce(acc, op=dist.ReduceOp.SUM)
    return acc


class FlatGradBucket:
    """One flat float32 bucket holding every parameter gradient: a single all-reduce (0.96 MB for the STSE
    encoder) enqueued right behind the last backward kernel, then averaged like DDP.

    ``attach()`` makes every ``p.grad`` a VIEW of the flat buffer (DDP's ``gradient_as_bucket_view``): autograd accumulates
    straight into the bucket, the all-reduce runs in place, and the optimizer reads the averaged gradients from the same
    memory -- no gather / scatter copies, and the addresses are stable, which is what lets the training step be replayed
    as CUDA graphs around the collective (trainer.TrainStep).  Without ``attach()`` the bucket gathers the gradients into
    a temporary flat tensor and scatters the average back."""

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        self.numel = sum(p.numel() for p in self.params)
        self.flat: Optional[torch.Tensor] = None
        self.views: List[torch.Tensor] = []

    def attached(self) -> bool:
        return self.flat is not None and all(p.grad is v for p, v in zip(self.params, self.views))

    def attach(self) -> 'FlatGradBucket':
        if not self.params:
            return self
        if self.flat is None or self.flat.device != self.params[0].device:
            # padded to a multiple of 4 floats: the flat Adam kernel (optim.FlatAdam) walks the buffers as float4
            self.flat = torch.zeros((self.numel + 3) // 4 * 4, device=self.params[0].device, dtype=torch.float32)
            self.views, off = [], 0
            for p in self.params:
                self.views.append(self.flat[off:off + p.numel()].view_as(p))
                off += p.numel()
        for p, v in zip(self.params, self.views):
            if p.grad is not v:
                if p.grad is not None:
                    v.copy_(p.grad)
                p.grad = v
            # the training kernels may accumulate into the (zeroed) views themselves instead of handing the gradient to
            # autograd's AccumulateGrad: one add_ launch per parameter and step less (train._direct, losses._L2Reg)
            p.coskad_direct_grad = True
        return self

    def detach(self) -> None:
        """give the gradients back to autograd (``p.grad = None``): after this the kernels return their gradients"""
        for p in self.params:
            p.grad = None
            p.coskad_direct_grad = False

    def zero_(self) -> None:
        """replaces optimizer.zero_grad(set_to_none=True), which would detach the views"""
        self.attach()
        if self.flat is not None:
            self.flat.zero_()

    def allreduce_(self) -> None:
        if not is_dist() or not self.params:
            return
        if self.attached():
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            self.flat.div_(world())
            return
        # one concatenation kernel in, one multi-tensor copy out (instead of two tiny copies per parameter)
        for p in self.params:
            if p.grad is None:
                p.grad = torch.zeros_like(p)
        grads = [p.grad for p in self.params]
        flat = torch.cat([g.reshape(-1) for g in grads])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat.div_(world())
        views = [v.view_as(g) for v, g in zip(torch.split(flat, [g.numel() for g in grads]), grads)]
        torch._foreach_copy_(grads, views)


class ShardedLoader:
    """every rank takes the batches ``i % world == rank`` of the wrapped loader and all ranks take the same number of them
    (the collectives of a data-parallel step must pair up): what Lightning's DDP strategy gets from a DistributedSampler
    (train_COSKAD.py:75-85).  ``set_epoch`` is forwarded to a sampler that has it."""

    def __init__(self, loader, rank_: Optional[int] = None, world_: Optional[int] = None):
        self.loader = loader
        self.rank = rank() if rank_ is None else rank_
        self.world = world() if world_ is None else world_
        self.batch_size = getattr(loader, 'batch_size', None)

    def set_epoch(self, epoch: int) -> None:
        sampler = getattr(self.loader, 'sampler', None)
        if hasattr(sampler, 'set_epoch'):
            sampler.set_epoch(epoch)

    def __len__(self) -> int:
        return len(self.loader) // self.world

    def __iter__(self):
        usable = (len(self.loader) // self.world) * self.world
        for i, batch in enumerate(self.loader):
            if i >= usable:
                break
            if i % self.world == self.rank:
                yield batch


def gather_rows(local: torch.Tensor, n_total: int) -> torch.Tensor:
    """all-gather per-window rows (scores [n] or latents [n, D]) of contiguous shards back into dataset
    order; shards are ceil(n_total/world) long, the last ones may be short or empty."""
    if not is_dist():
        return local
    w = world()
    per = (n_total + w - 1) // w
    pad_shape = (per,) + tuple(local.shape[1:])
    buf = torch.zeros(pad_shape, dtype=local.dtype, device=local.device)
    buf[: local.shape[0]] = local
    out = torch.empty((w * per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, buf)
    return out[:n_total]


def broadcast_module_(module: torch.nn.Module, src: int = 0) -> None:
    """identical initial weights on every rank (what DDP does at wrap time)"""
    if not is_dist():
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src=src)
