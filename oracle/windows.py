"""ORACLE (test infrastructure only): CPU restatement of the reference's window construction and test-time transforms.

    sliding windows over trajectory rows    utils/preprocessing.py:58-89 (_aggregate_rnn_autoencoder_data, input_gap 0)
    [T, 2V] rows -> [C, T, V] windows       utils/get_robust_data.py (reshape(-1, seg_len, V, C) -> transpose to (C, T, V))
    affine matrices / application           utils/dataset_utils.py:255-284 (get_aff_trans_mat, apply_pose_transform)
    the five evaluation transforms          utils/dataset_utils.py:304-310 (ae_trans_list)
    dataset index -> (sample, transform)    utils/dataset.py:65-74 (index % num_samples, index // num_samples)

Pinned: tests/golden/windows_ref.npz holds outputs of the reference's own get_aff_trans_mat / apply_pose_transform /
ae_trans_list imported from /root/reference (oracle/gen_golden.py).  Only tests/ may import this module.
"""
from __future__ import annotations

import math

import numpy as np


def get_aff_trans_mat(sx=1, sy=1, tx=0, ty=0, rot=0, flip=False) -> np.ndarray:
    """utils/dataset_utils.py:255-268 (float32 matrices, flip @ (rot @ trans_scale))"""
    cos_r, sin_r = math.cos(math.radians(rot)), math.sin(math.radians(rot))
    flip_mat = np.eye(3, dtype=np.float32)
    if flip:
        flip_mat[0, 0] = -1.0
    trans_scale = np.array([[sx, 0, tx], [0, sy, ty], [0, 0, 1]], dtype=np.float32)
    rot_mat = np.array([[cos_r, -sin_r, 0], [sin_r, cos_r, 0], [0, 0, 1]], dtype=np.float32)
    return (flip_mat @ (rot_mat @ trans_scale)).astype(np.float32)


# utils/dataset_utils.py:304-310
AE_TRANS = [dict(rot=0, flip=False), dict(rot=0, flip=True), dict(rot=90, flip=False), dict(rot=90, flip=True),
            dict(rot=45, flip=False)]


def ae_trans_mats() -> np.ndarray:
    return np.stack([get_aff_trans_mat(**kw) for kw in AE_TRANS], 0)


def apply_pose_transform(pose: np.ndarray, trans_mat: np.ndarray) -> np.ndarray:
    """utils/dataset_utils.py:271-284 for a [C>=2, T, V] window: (x, y, 1) -> trans_mat rows 0,1; extra channels kept"""
    ones = np.ones_like(pose[:1])
    pw1 = np.concatenate([pose[:2], ones], axis=0)
    out = np.einsum('ktv,ck->ctv', pw1, trans_mat)
    return np.concatenate([out[:2], pose[2:]], axis=0)


def windows_from_rows(traj: np.ndarray, win_row: np.ndarray, seg_len: int = 12) -> np.ndarray:
    """traj [rows, 2V] (x0,y0,x1,y1,..) -> windows [N, 2, seg_len, V]: x[c][t][v] = traj[row+t][2v+c]"""
    V = traj.shape[1] // 2
    out = np.empty((len(win_row), 2, seg_len, V), dtype=traj.dtype)
    for i, r in enumerate(win_row):
        w = traj[r:r + seg_len].reshape(seg_len, V, 2)
        out[i] = w.transpose(2, 0, 1)
    return out


def sliding_starts(n_rows: int, seg_len: int = 12, stride: int = 1) -> np.ndarray:
    """start rows of the windows of one trajectory (utils/preprocessing.py:66-78: range(0, n - len + 1, stride))"""
    return np.arange(0, n_rows - seg_len + 1, stride, dtype=np.int64)
