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
    encoder) enqueued right behind the last backward kernel, then averaged like DDP."""

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        self.numel = sum(p.numel() for p in self.params)
        self.flat: Optional[torch.Tensor] = None

    def allreduce_(self) -> None:
        if not is_dist() or not self.params:
            return
        # one concatenation kernel in, one multi-tensor copy out (instead of two tiny copies per parameter)
        for p in self.params:
            if p.grad is None:
                p.grad = torch.zeros_like(p)
        grads = [p.grad for p in self.params]
        self.flat = torch.cat([g.reshape(-1) for g in grads])
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
        self.flat.div_(world())
        views = [v.view_as(g) for v, g in zip(torch.split(self.flat, [g.numel() for g in grads]), grads)]
        torch._foreach_copy_(grads, views)


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
