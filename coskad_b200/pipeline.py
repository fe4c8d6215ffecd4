"""Public end-to-end scoring API: host pose windows in, anomaly scores out.

``score_windows_host`` is the call a user of the drop-in makes with data that lives on the host
(what eval_COSKAD.py:115-120 does with a DataLoader + ``predict`` + ``light_processing_data``):
pinned host windows are streamed to the B200 in chunks on a copy stream while the fused kernel
scores the previous chunk on the compute stream, and the scores come back to pinned host memory.
``shard_range`` is the window partition used by every multi-GPU entry point (contiguous blocks in
dataset order).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """contiguous block of ceil(n/world) windows per rank, dataset order (last ranks may be short/empty)"""
    per = (n + world - 1) // world
    lo = min(rank * per, n)
    return lo, min(lo + per, n)


class HostScorer:
    """Double-buffered H2D -> fused kernel -> D2H pipeline around ``STSE.encode_score``.

    ``chunk``: windows per H2D copy / kernel launch.  Only the first chunk's copy is exposed, so small chunks win as long
    as a launch still fills the 148 persistent CTAs evenly: 16 384 windows = 36.9 tiles per CTA (measured on one B200:
    12.99 M windows/s end to end vs 12.36 M with 131 072-window chunks, device-resident rate 13.1 M)."""

    def __init__(self, model, flavour: int = _lib.SCORE_POINCARE, chunk: int = 16384, device: Optional[int] = None):
        self.model, self.flavour, self.chunk = model, flavour, int(chunk)
        self.device = torch.device('cuda', torch.cuda.current_device() if device is None else device)
        shape = (self.chunk, model.input_dim, model.n_frames, model.n_joints)
        self.dbuf = [torch.empty(shape, device=self.device, dtype=torch.float32) for _ in range(2)]
        self.copy_stream = torch.cuda.Stream(self.device)
        self.ready = [torch.cuda.Event() for _ in range(2)]     # H2D of buffer i done
        self.free = [torch.cuda.Event() for _ in range(2)]      # kernel reading buffer i done
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    @torch.no_grad()
    def score(self, x_host: torch.Tensor, out_host: Optional[torch.Tensor] = None,
              center: Optional[torch.Tensor] = None, on_device_scores=None) -> torch.Tensor:
        """x_host [N,C,T,V] float32 (pinned for full overlap) -> scores [N] float32 on the host.
        ``on_device_scores(dscore)`` (optional) runs stream-ordered between the last kernel and the D2H copy -- the hook of
        the multi-GPU score all-gather."""
        assert not x_host.is_cuda and x_host.dtype == torch.float32 and x_host.is_contiguous()
        N = x_host.shape[0]
        if out_host is None:
            out_host = torch.empty(N, dtype=torch.float32).pin_memory()
        comp = torch.cuda.current_stream(self.device)
        dscore = torch.empty(N, device=self.device, dtype=torch.float32)
        nchunks = (N + self.chunk - 1) // self.chunk
        for b in range(2):
            self.free[b].record(comp)
        for i in range(nchunks):
            b = i & 1
            lo, hi = i * self.chunk, min((i + 1) * self.chunk, N)
            with torch.cuda.stream(self.copy_stream):
                self.copy_stream.wait_event(self.free[b])
                self.dbuf[b][: hi - lo].copy_(x_host[lo:hi], non_blocking=True)
                self.ready[b].record(self.copy_stream)
            comp.wait_event(self.ready[b])
            self.model.encode_score(self.dbuf[b][: hi - lo], self.flavour, center=center, want_latent=False,
                                    score_out=dscore[lo:hi])
            self.free[b].record(comp)
            self.h2d_bytes += (hi - lo) * x_host[0].numel() * 4
        if on_device_scores is not None:
            on_device_scores(dscore)
        out_host.copy_(dscore, non_blocking=True)
        self.d2h_bytes += N * 4
        comp.synchronize()
        return out_host


def score_windows_host(model, x_host: torch.Tensor, flavour: int = _lib.SCORE_POINCARE, chunk: int = 16384,
                       center: Optional[torch.Tensor] = None) -> torch.Tensor:
    return HostScorer(model, flavour, chunk).score(x_host, center=center)


class TrajectoryScorer:
    """Host trajectories in, anomaly scores out, through the trajectory front end of the fused kernel.

    What eval_COSKAD.py:107-120 does with ``get_dataset_and_loader`` (sliding windows materialised on the host, one copy
    per test-time transform) + ``predict``: here only the scaled trajectory rows (136 B per frame) and two index arrays
    cross PCIe -- with stride-1 windows and ``num_transform`` = 5 that is ~40x fewer bytes than the window tensor -- and
    the windows are assembled and transformed inside the kernel's input stage (``coskad_encode_score_traj_fwd``)."""

    def __init__(self, model, flavour: int = _lib.SCORE_POINCARE, device: Optional[int] = None, chunk: int = 65536):
        self.model, self.flavour, self.chunk = model, flavour, int(chunk)
        self.device = torch.device('cuda', torch.cuda.current_device() if device is None else device)
        self.copy_stream = torch.cuda.Stream(self.device)
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    @torch.no_grad()
    def score(self, traj_host: torch.Tensor, win_row_host: torch.Tensor, trans_host: Optional[torch.Tensor] = None,
              mats: Optional[torch.Tensor] = None, out_host: Optional[torch.Tensor] = None,
              center: Optional[torch.Tensor] = None, keep_on_device: bool = False) -> torch.Tensor:
        """traj_host [rows, 2V] f32, win_row_host [N] i64, trans_host [N] i32 (optional, with mats [n,2,3]) -> scores [N].
        The trajectory buffer goes over first; the per-window index arrays follow in chunks on a copy stream while the
        kernel scores the previous chunk."""
        assert not traj_host.is_cuda and traj_host.dtype == torch.float32 and traj_host.is_contiguous()
        N = win_row_host.numel()
        if out_host is None and not keep_on_device:
            out_host = torch.empty(N, dtype=torch.float32).pin_memory()
        comp = torch.cuda.current_stream(self.device)
        traj = traj_host.to(self.device, non_blocking=True)
        self.h2d_bytes += traj_host.numel() * 4
        mt = None
        if trans_host is not None:
            trans_host = trans_host.to(torch.int32)
            mt = mats.to(self.device, dtype=torch.float32)
            self.h2d_bytes += mt.numel() * 4
        dscore = torch.empty(N, device=self.device, dtype=torch.float32)
        rows_d = torch.empty(N, device=self.device, dtype=torch.int64)
        tr_d = torch.empty(N, device=self.device, dtype=torch.int32) if trans_host is not None else None
        nchunks = (N + self.chunk - 1) // self.chunk
        ready = [torch.cuda.Event() for _ in range(nchunks)]
        self.copy_stream.wait_stream(comp)
        with torch.cuda.stream(self.copy_stream):
            for i in range(nchunks):
                lo, hi = i * self.chunk, min((i + 1) * self.chunk, N)
                rows_d[lo:hi].copy_(win_row_host[lo:hi], non_blocking=True)
                if tr_d is not None:
                    tr_d[lo:hi].copy_(trans_host[lo:hi], non_blocking=True)
                ready[i].record(self.copy_stream)
        for i in range(nchunks):
            lo, hi = i * self.chunk, min((i + 1) * self.chunk, N)
            comp.wait_event(ready[i])
            self.model.encode_score_traj(traj, rows_d[lo:hi], None if tr_d is None else tr_d[lo:hi], mt, flavour=self.flavour,
                                         center=center, want_latent=False, score_out=dscore[lo:hi])
        self.h2d_bytes += N * 8 + (N * 4 if tr_d is not None else 0)
        if keep_on_device:           # the aggregation / AUC tail consumes the scores where they are (stream-ordered)
            return dscore
        out_host.copy_(dscore, non_blocking=True)
        self.d2h_bytes += N * 4
        comp.synchronize()
        return out_host
