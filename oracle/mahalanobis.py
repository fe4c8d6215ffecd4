"""CPU restatement of the Mahalanobis scoring of ``distance: 'mahalanobis'`` (TEST INFRASTRUCTURE ONLY).

Follows utils/eval_utils.py:28-55 (``mahalanobis``, ``windows_based_loss_mahalanobis``) and
models/euclidean_encoder_staticCenter.py:40-46,133-142 (``batch_cov_mat_step``, ``compute_inv_cov_mat``).
Pinned: ``oracle/gen_golden.py`` runs the reference's own functions on the same inputs, asserts equality with this
restatement and stores the REFERENCE outputs in tests/golden/mahalanobis_ref.npz.
"""
from __future__ import annotations

import numpy as np
import torch


def mahalanobis(u: torch.Tensor, v: torch.Tensor, VI: torch.Tensor, reduce: str = 'mean') -> torch.Tensor:
    """utils/eval_utils.py:28-38: sqrt((u - v)^T VI (u - v)) per row, [B, 1, 1] unless reduced"""
    if u.dim() < 3:
        u = u.reshape(*u.shape, 1)
    if v.dim() < 3:
        v = v.reshape(*v.shape, 1)
    d = torch.sqrt(torch.matmul(torch.matmul(torch.transpose(u - v, 1, 2), VI), u - v))
    return d.mean() if reduce == 'mean' else d


def batch_cov_mat_step(X: torch.Tensor, mu: torch.Tensor) -> torch.Tensor:
    """models/euclidean_encoder_staticCenter.py:40-46: sum over the batch of (x - mu)(x - mu)^T"""
    if X.dim() < 3:
        X = X.reshape(*X.shape, 1)
    if mu.dim() < 3:
        mu = mu.reshape(*mu.shape, 1)
    return torch.sum(torch.matmul(X - mu, torch.transpose(X - mu, 1, 2)), dim=0)


def inv_cov(batches, mu: torch.Tensor) -> torch.Tensor:
    """models/euclidean_encoder_staticCenter.py:133-142 compute_inv_cov_mat over cached latent batches"""
    s = torch.zeros(mu.numel(), mu.numel())
    n = 0
    for h in batches:
        s += batch_cov_mat_step(h, mu)
        n += h.shape[0]
    return torch.inverse(s / (n - 1))


def windows_based_loss_mahalanobis(hidden_c, hidden_out_fig, VI, frames_fig, n_frames) -> np.ndarray:
    """utils/eval_utils.py:41-55 (float64 [w, n_frames]; pose[n, frames - 1] = distance of window n)"""
    w = hidden_out_fig.shape[0]
    loss = mahalanobis(torch.from_numpy(hidden_out_fig), hidden_c, VI, reduce='none')
    loss = torch.mean(loss, dim=-1)
    pose = np.zeros(shape=(w, n_frames))
    for n in range(w):
        pose[n, frames_fig[n] - 1] = loss[n]
    return pose
