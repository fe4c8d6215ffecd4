"""Oracle (TEST INFRASTRUCTURE): numpy transcription of the reference's frame-level aggregation.

Follows utils/eval_utils.py:57-106 (scatter of per-window scores to frames, index ``frames - 1``),
eval_COSKAD.py:140-253 (transformation x clip x person loops, ``== 0 -> NaN``, nanmean, amax,
score_process, AUC) and utils/eval_utils.py:200-207,232-248 (score_process, pad_scores).
The per-window scores are an INPUT here (they come from the network + geometry oracles); the
``.cuda()`` / ``.cpu()`` hops of the reference are dropped, the arithmetic is unchanged.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
from scipy.ndimage import gaussian_filter1d


def scatter_windows(loss: np.ndarray, frames_fig: np.ndarray, n_frames: int) -> np.ndarray:
    """utils/eval_utils.py:69-74: pose[n, frames_fig[n] - 1] = loss[n]  (float64 [w, n_frames])."""
    w = loss.shape[0]
    pose = np.zeros(shape=(w, n_frames))
    for n in range(pose.shape[0]):
        pose[n, frames_fig[n] - 1] = loss[n]   # added -1 (upstream comment)
    return pose


def person_curve(loss: np.ndarray, frames_fig: np.ndarray, n_frames: int) -> np.ndarray:
    """eval_COSKAD.py:201-203."""
    loss_matrix = scatter_windows(loss, frames_fig, n_frames)
    loss_matrix = np.where(loss_matrix == 0.0, np.nan, loss_matrix)
    with np.errstate(all='ignore'):
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            fig = np.nanmean(loss_matrix, 0)
    return np.where(np.isnan(fig), 0, fig)


def ranges(nums):                                      # utils/eval_utils.py:210-214
    nums = sorted(set(nums))
    gaps = [[s, e] for s, e in zip(nums, nums[1:]) if s + 1 < e]
    edges = iter(nums[:1] + sum(gaps, []) + nums[-1:])
    return list(zip(edges, edges))


def pad_scores(fig_reconstruction_loss: np.ndarray, gt: np.ndarray, pad_size: int) -> np.ndarray:
    """utils/eval_utils.py:232-248 (in place, like upstream)."""
    zero_interval = set(list(range(len(gt) - 1))) - set(np.nonzero(fig_reconstruction_loss)[0])
    non_presence_intervals = ranges(zero_interval)
    nope = []
    for _, interval in enumerate(non_presence_intervals):
        start, end = interval
        if start == 0 and end == len(gt) - 2:
            continue
        elif start == 0 and end != len(gt) - 2:
            nope.append((start, min(end + pad_size, len(gt))))
        elif start != 0 and end == len(gt) - 2:
            nope.append((max(start - pad_size, 0), end))
        elif start != 0 and end != len(gt) - 2:
            nope.append((max(start - pad_size, 0), min(end + pad_size, len(gt))))
    for interval in nope:
        fig_reconstruction_loss[range(interval[0], interval[1])] = 0
    return fig_reconstruction_loss


def score_process(score: np.ndarray) -> np.ndarray:
    """utils/eval_utils.py:200-207 (win_size / dataname / use_scaler are ignored upstream)."""
    scores_shifted = np.zeros_like(score)
    shift = 8 + (8 // 2) - 1
    scores_shifted[shift:] = score[:-shift]
    return gaussian_filter1d(scores_shifted, 30)


def aggregate_dataset(scores: np.ndarray, trans: np.ndarray, meta: np.ndarray, frames: np.ndarray,
                      clips: Sequence[Tuple[int, int, int]], num_transform: int, pad_size: int = -1,
                      gts: Optional[Dict[Tuple[int, int], np.ndarray]] = None,
                      smooth: bool = True) -> Dict[int, List[np.ndarray]]:
    """eval_COSKAD.py:140-220 with the per-window score already computed.

    clips: (scene, clip, n_frames) in the sorted gt-file order.  Returns, per transformation, the
    list of per-clip score curves (after score_process when ``smooth``)."""
    out: Dict[int, List[np.ndarray]] = {}
    for transformation in range(num_transform):
        cond_transform = (trans == transformation)
        s_t, meta_t, frames_t = scores[cond_transform], meta[cond_transform], frames[cond_transform]
        model_scores = []
        for scene_idx, clip_idx, n_frames in clips:
            cond = (meta_t[:, 0] == scene_idx) & (meta_t[:, 1] == clip_idx)
            s_c, meta_c, frames_c = s_t[cond], meta_t[cond], frames_t[cond]
            figs_ids = sorted(list(set(meta_c[:, 2])))
            error_per_person = []
            for fig in figs_ids:
                cond_fig = (meta_c[:, 2] == fig)
                fig_loss = person_curve(s_c[cond_fig], frames_c[cond_fig], n_frames)
                if pad_size != -1:
                    gt = gts[(scene_idx, clip_idx)] if gts is not None else np.zeros(n_frames)
                    fig_loss = pad_scores(fig_loss, gt, pad_size)
                error_per_person.append(fig_loss)
            clip_score = np.amax(np.stack(error_per_person, axis=0), axis=0)
            if smooth:
                clip_score = score_process(clip_score)
            model_scores.append(clip_score)
        out[transformation] = model_scores
    return out


def synth_dataset(n_clips: int = 6, seed: int = 0, num_transform: int = 2, seg_len: int = 12,
                  max_persons: int = 4, frame_range=(60, 160)):
    """Synthetic window metadata in the reference's batch-tuple format (utils/dataset.py:87-95,
    utils/preprocessing.py:18-55): trans [N], meta [N,4]=(scene,clip,person,start), frames [N,seg_len]
    (1-based ids like upstream, some tracks starting at frame id 0 to exercise the -1 wrap), plus the
    clip table and Bernoulli ground truth."""
    rng = np.random.default_rng(seed)
    trans, meta, frames, clips, gts = [], [], [], [], {}
    for ci in range(n_clips):
        scene, clip = 1 + ci // 3, 1 + ci % 3
        F = int(rng.integers(frame_range[0], frame_range[1]))
        clips.append((scene, clip, F))
        gt = (rng.random(F) < 0.15).astype(np.int64)
        gt[:3] = [0, 1, 0]
        gts[(scene, clip)] = gt
        for person in range(1, 1 + int(rng.integers(1, max_persons + 1))):
            first = int(rng.integers(0, max(1, F // 3)))          # frame id of the first window
            last = int(rng.integers(first + seg_len, F + 1))
            starts = list(range(first, last - seg_len + 1))
            if len(starts) > 4 and rng.random() < 0.5:            # a gap in the track
                del starts[len(starts) // 2: len(starts) // 2 + 3]
            for t in range(num_transform):
                for s in starts:
                    trans.append(t)
                    meta.append((scene, clip, person, s))
                    frames.append(np.arange(s, s + seg_len))
    order = rng.permutation(len(trans))                           # the loader does not sort by person
    trans = np.asarray(trans, dtype=np.int64)[order]
    meta = np.asarray(meta, dtype=np.int64)[order]
    frames = np.asarray(frames, dtype=np.int64)[order]
    return trans, meta, frames, clips, gts
