"""Synthetic pose-window datasets in the reference's batch-tuple format.

The reference's loader yields ``[x f32 [C,T,V], trans_idx, meta [4]=(scene,clip,person,start), frames [T]]``
(utils/dataset.py:87-95, utils/preprocessing.py:18-55).  Dataset loading / preprocessing itself is CPU
data prep and out of scope; this generator stands in for it offline: per-person
smooth trajectories, stride-1 sliding windows, ``num_transform`` copies, Bernoulli frame labels with a
planted perturbation on anomalous frames so that a trained scorer separates them.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np
import torch
from torch.utils.data import DataLoader, Dataset


class SyntheticPoseDataset(Dataset):
    def __init__(self, n_clips: int = 8, seed: int = 0, num_transform: int = 1, seg_len: int = 12, n_joints: int = 17,
                 n_coords: int = 2, max_persons: int = 3, frame_range=(80, 200), anomaly_rate: float = 0.15,
                 anomaly_scale: float = 1.5, train: bool = False):
        rng = np.random.default_rng(seed)
        xs, trans, meta, frames = [], [], [], []
        self.clips: List[Tuple[int, int, int]] = []
        self.gts: Dict[Tuple[int, int], np.ndarray] = {}
        base_pose = rng.normal(0, 0.4, size=(n_coords, n_joints)).astype(np.float32)
        for ci in range(n_clips):
            scene, clip = 1 + ci // 4, 1 + ci % 4
            F = int(rng.integers(frame_range[0], frame_range[1]))
            gt = np.zeros(F, dtype=np.int64)
            if not train:
                nseg = max(1, int(anomaly_rate * F / 15))
                for _ in range(nseg):
                    a = int(rng.integers(0, F - 15))
                    gt[a:a + 15] = 1
            self.clips.append((scene, clip, F))
            self.gts[(scene, clip)] = gt
            for person in range(1, 1 + int(rng.integers(1, max_persons + 1))):
                first = int(rng.integers(1, max(2, F // 4)))
                last = int(rng.integers(min(F - 1, first + seg_len + 5), F + 1))
                t = np.arange(first, last)
                phase = rng.uniform(0, 2 * np.pi, size=(n_coords, n_joints, 1))
                traj = base_pose[:, :, None] + 0.15 * np.sin(0.2 * t[None, None, :] + phase) + \
                    rng.normal(0, 0.03, size=(n_coords, n_joints, len(t)))
                an = gt[np.clip(t - 1, 0, F - 1)] == 1          # frame ids are 1-based
                traj = traj + an[None, None, :] * rng.normal(0, 0.25 * anomaly_scale, size=traj.shape)
                traj = np.clip(traj, -3, 3).astype(np.float32)
                for tr in range(num_transform):
                    sgn = -1.0 if tr % 2 else 1.0                 # stand-in for the affine test-time transforms
                    for s in range(0, len(t) - seg_len + 1):
                        w = traj[:, :, s:s + seg_len].transpose(0, 2, 1).copy()    # [C, T, V]
                        w[0] *= sgn
                        xs.append(w)
                        trans.append(tr)
                        meta.append((scene, clip, person, int(t[s])))
                        frames.append(t[s:s + seg_len])
        self.x = torch.from_numpy(np.stack(xs)).contiguous()
        self.trans = torch.tensor(trans, dtype=torch.int64)
        self.meta = torch.tensor(meta, dtype=torch.int64)
        self.frames = torch.from_numpy(np.stack(frames).astype(np.int64))
        self.num_transform = num_transform

    def __len__(self) -> int:
        return self.x.shape[0]

    def __getitem__(self, i):
        return [self.x[i], self.trans[i], self.meta[i], self.frames[i]]


def get_dataset_and_loader(args, split: str = 'train', validation: bool = False):
    """same call shape as utils/dataset.py:284-327 for the synthetic stand-in"""
    train = split == 'train'
    ds = SyntheticPoseDataset(n_clips=getattr(args, 'synthetic_clips', 8), seed=getattr(args, 'seed', 999) + (0 if train else 1),
                              num_transform=1 if train else max(1, getattr(args, 'num_transform', 1)),
                              seg_len=getattr(args, 'seg_len', 12), train=train)
    loader = DataLoader(ds, batch_size=getattr(args, 'batch_size', 2048), shuffle=train, drop_last=False, num_workers=0)
    return ds, loader
