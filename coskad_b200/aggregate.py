"""Frame-level score aggregation: the batched replacement of the Python triple loop in
eval_COSKAD.py:140-220 (transformation x clip x person, one ``.cpu()`` sync per window upstream)
and of utils/eval_utils.py:57-106.

``score_and_aggregate`` groups the windows by (transformation, clip, person) once on the host
(index arithmetic only), then one CUDA launch builds every person's per-frame curve (float64,
window order = dataset order, bit-exact with numpy's nanmean) and a second one takes the max over
each clip's persons.  ``pad_scores`` / ``score_process`` / AUC stay on the host with the very same
scipy / sklearn calls the reference makes (negligible cost, parity "AUC to 4
decimals").  The reference-named helpers at the bottom keep ``utils.eval_utils``' signatures.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib


def _runs(mask: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """(first, last) index of every maximal run of True in a boolean vector"""
    edge = np.diff(np.concatenate(([0], mask.astype(np.int8), [0])))
    return np.nonzero(edge == 1)[0], np.nonzero(edge == -1)[0] - 1


def ranges(nums) -> List[Tuple[int, int]]:
    """maximal runs of consecutive integers in ``nums`` as inclusive (first, last) pairs (utils/eval_utils.py:210-214)"""
    vals = np.unique(np.asarray(list(nums), dtype=np.int64))
    if vals.size == 0:
        return []
    cut = np.nonzero(np.diff(vals) > 1)[0]
    first = np.concatenate(([vals[0]], vals[cut + 1]))
    last = np.concatenate((vals[cut], [vals[-1]]))
    return [(int(a), int(b)) for a, b in zip(first, last)]


def pad_scores(fig_reconstruction_loss, gt, pad_size):
    """Widen every gap in which the person is absent by ``pad_size`` frames (in place, like utils/eval_utils.py:232-248).

    Interval arithmetic over the first L-1 frames (L = len(gt)): a run [s, e] of exact-zero scores is widened to
    [s - pad, e + pad) clipped to [0, L) -- except that a run touching frame 0 keeps its start, a run touching frame L-2
    keeps e as its (exclusive) end, and a run covering all of [0, L-2] (person never present) is left alone.  The widened
    intervals are merged with a difference array and zeroed in one assignment."""
    L = len(gt)
    curve = fig_reconstruction_loss
    if L < 2:
        return curve
    absent = np.ones(L - 1, dtype=bool)
    present = np.nonzero(curve)[0]
    absent[present[present < L - 1]] = False
    s, e = _runs(absent)
    keep = ~((s == 0) & (e == L - 2))
    s, e = s[keep], e[keep]
    lo = np.where(s == 0, 0, np.maximum(s - pad_size, 0))
    hi = np.where(e == L - 2, e, np.minimum(e + pad_size, L))
    ok = hi > lo
    mark = np.zeros(L + 1, dtype=np.int64)
    np.add.at(mark, lo[ok], 1)
    np.add.at(mark, hi[ok], -1)
    curve[np.cumsum(mark[:L]) > 0] = 0
    return curve


SHIFT_FRAMES = 8 + (8 // 2) - 1       # utils/eval_utils.py:203
SMOOTH_SIGMA = 30                     # utils/eval_utils.py:205


def score_process(score, win_size=50, dataname='STC', use_scaler=False):
    """delay the curve by SHIFT_FRAMES (zeros shifted in) and smooth it with scipy's gaussian_filter1d(sigma 30); the other
    arguments are accepted and ignored, like upstream (utils/eval_utils.py:200-207)"""
    from scipy.ndimage import gaussian_filter1d
    delayed = np.zeros_like(score)
    n = len(score) - SHIFT_FRAMES
    if n > 0:
        delayed[SHIFT_FRAMES:] = score[:n]
    return gaussian_filter1d(delayed, SMOOTH_SIGMA)


def filter_vectors_by_cond(vecs, cond):
    """boolean-mask every array of ``vecs`` (utils/eval_utils.py:168-172)"""
    return [np.asarray(v)[cond] if not torch.is_tensor(v) else v[cond] for v in vecs]


class GroupIndex:
    """CSR grouping of windows by (transformation, clip, person) -- host index arithmetic only."""

    def __init__(self, trans: np.ndarray, meta: np.ndarray, clips: Sequence[Tuple[int, int, int]], num_transform: int):
        trans = np.asarray(trans).astype(np.int64).reshape(-1)
        meta = np.asarray(meta).astype(np.int64)
        n_clip = len(clips)
        nfr = np.asarray([int(f) for _, _, f in clips], dtype=np.int64)
        mult = int(meta[:, 1].max(initial=0)) + 1 + max((c for _, c, _ in clips), default=0)
        sc = meta[:, 0] * mult + meta[:, 1]
        if n_clip:
            keys = np.asarray([s * mult + c for s, c, _ in clips], dtype=np.int64)
            order = np.argsort(keys)
            pos = np.clip(np.searchsorted(keys[order], sc), 0, n_clip - 1)
            cidx = np.where(keys[order][pos] == sc, order[pos], -1)     # (scene, clip) -> clip index, -1 if unknown
        else:
            cidx = np.full(len(sc), -1, dtype=np.int64)
        sel = np.nonzero((cidx >= 0) & (trans >= 0) & (trans < num_transform))[0]
        gclip = trans[sel] * n_clip + cidx[sel]                       # global clip id: transformation major
        person = meta[sel, 2]
        srt = np.lexsort((sel, person, gclip))                        # stable in dataset order inside a person
        self.win_idx = sel[srt]
        g, p = gclip[srt], person[srt]
        new = np.ones(len(srt), dtype=bool)
        new[1:] = (g[1:] != g[:-1]) | (p[1:] != p[:-1])
        starts = np.nonzero(new)[0]
        self.n_persons = len(starts)
        self.person_off = np.concatenate([starts, [len(srt)]]).astype(np.int64)
        self.person_clip = g[starts].astype(np.int32)
        self.person_id = p[starts]
        self.n_clips = num_transform * n_clip
        self.clip_frames = np.tile(nfr, num_transform)
        self.clip_off = np.concatenate([[0], np.cumsum(self.clip_frames)]).astype(np.int64)
        self.clip_person_off = np.searchsorted(self.person_clip, np.arange(self.n_clips + 1)).astype(np.int64)
        pf = self.clip_frames[self.person_clip] if self.n_persons else np.zeros(0, dtype=np.int64)
        self.person_out_off = np.concatenate([[0], np.cumsum(pf)]).astype(np.int64)
        self.n_clip_per_transform = n_clip


class DeviceGroupIndex:
    """The same CSR grouping built where the scores live: no O(N log N) host work and no D2H of the per-window metadata.

    Windows are ordered by one sort of the packed key ((transformation * n_clips + clip) * P + person) * N + window --
    unique per window, so the order inside a person is the dataset order the bit-exact float64 mean needs -- and the
    person / clip boundaries come from neighbour comparisons, prefix sums and a binary search.  The sort, scan and search
    are torch (CUB) primitives: index plumbing; the arithmetic on scores is in ``coskad_frame_aggregate``.  Fields are
    tensors on ``device`` with the meaning of :class:`GroupIndex` (equality is tested on the CPU in
    tests/test_group_index_cpu.py)."""

    def __init__(self, trans, meta, clips: Sequence[Tuple[int, int, int]], num_transform: int, device=None):
        as_t = lambda a: a if torch.is_tensor(a) else torch.as_tensor(np.ascontiguousarray(a))
        trans = as_t(trans).reshape(-1)
        meta = as_t(meta)
        dev = torch.device(device) if device is not None else meta.device
        trans = trans.to(device=dev, dtype=torch.int64)
        meta = meta.to(device=dev, dtype=torch.int64)
        N = int(trans.numel())
        n_clip = len(clips)
        i64 = dict(dtype=torch.int64, device=dev)
        nfr = torch.tensor([int(f) for _, _, f in clips], **i64)
        mult = (int(meta[:, 1].max()) if N else 0) + 1 + max((c for _, c, _ in clips), default=0)
        sc = meta[:, 0] * mult + meta[:, 1]
        if n_clip and N:
            keys = torch.tensor([s * mult + c for s, c, _ in clips], **i64)
            skeys, order = torch.sort(keys)
            pos = torch.searchsorted(skeys, sc).clamp_(0, n_clip - 1)
            cidx = torch.where(skeys[pos] == sc, order[pos], torch.full_like(sc, -1))
        else:
            cidx = torch.full((N,), -1, **i64)
        ok = (cidx >= 0) & (trans >= 0) & (trans < num_transform)
        person = meta[:, 2] if N else torch.zeros(0, **i64)
        P = (int(person.max()) if N else 0) + 1
        gp = (trans * n_clip + cidx) * P + person                         # (global clip, person), transformation major
        big = num_transform * max(n_clip, 1) * P                          # rejected windows sort behind every real key
        if (big + 1) * max(N, 1) >= 2 ** 62:
            raise ValueError('aggregation key does not fit 62 bits: too many clips x persons x windows')
        packed = torch.where(ok, gp, torch.full_like(gp, big)) * max(N, 1) + torch.arange(N, **i64)
        srt = torch.sort(packed).values
        n_sel = int(ok.sum()) if N else 0
        srt = srt[:n_sel]
        self.win_idx = srt % max(N, 1)
        g = srt // max(N, 1)
        new = torch.ones(n_sel, dtype=torch.bool, device=dev)
        if n_sel > 1:
            new[1:] = g[1:] != g[:-1]
        starts = torch.nonzero(new).view(-1)
        self.n_persons = int(starts.numel())
        self.person_off = torch.cat([starts, torch.tensor([n_sel], **i64)])
        self.person_clip = (g[starts] // P).to(torch.int32)
        self.person_id = g[starts] % P
        self.n_clips = num_transform * n_clip
        self.clip_frames = nfr.repeat(num_transform)
        zero = torch.zeros(1, **i64)
        self.clip_off = torch.cat([zero, torch.cumsum(self.clip_frames, 0)])
        self.clip_person_off = torch.searchsorted(self.person_clip.to(torch.int64), torch.arange(self.n_clips + 1, **i64))
        pf = self.clip_frames[self.person_clip.to(torch.int64)]
        self.person_out_off = torch.cat([zero, torch.cumsum(pf, 0)])
        self.n_clip_per_transform = n_clip
        self.total_person_frames = int(self.person_out_off[-1])
        self.total_clip_frames = int(self.clip_off[-1])
        self.max_clip_frames = int(nfr.max()) if n_clip else 0


def aggregate_curves(score: torch.Tensor, frames, gi) -> Tuple[torch.Tensor, torch.Tensor]:
    """(per-person curves, per-clip max curves) as float64 device tensors; ``gi`` is a GroupIndex (host arrays, uploaded
    here) or a DeviceGroupIndex (already on the device)"""
    if not score.is_cuda:
        raise _lib.CoskadError('scores must be a CUDA tensor: the aggregation runs on the B200 (no CPU fallback)')
    dev = score.device
    score = score.detach().to(torch.float32).contiguous().view(-1)
    frames_d = torch.as_tensor(np.ascontiguousarray(frames), dtype=torch.int64).to(dev) if not torch.is_tensor(frames) \
        else frames.to(device=dev, dtype=torch.int64).contiguous()
    T = int(frames_d.shape[1])
    t = lambda a, dt: (a.to(device=dev, dtype=dt) if torch.is_tensor(a) else
                       torch.as_tensor(np.ascontiguousarray(a), dtype=dt).to(dev)).contiguous()
    win_idx, person_off = t(gi.win_idx, torch.int64), t(gi.person_off, torch.int64)
    person_clip, person_out_off = t(gi.person_clip, torch.int32), t(gi.person_out_off, torch.int64)
    clip_person_off, clip_off = t(gi.clip_person_off, torch.int64), t(gi.clip_off, torch.int64)
    if isinstance(gi, DeviceGroupIndex):
        total_pf, total_cf, max_cf = gi.total_person_frames, gi.total_clip_frames, gi.max_clip_frames
    else:
        total_pf, total_cf = int(gi.person_out_off[-1]), int(gi.clip_off[-1])
        max_cf = int(gi.clip_frames.max(initial=0))
    person_out = torch.empty(max(total_pf, 1), dtype=torch.float64, device=dev)
    out = torch.zeros(max(total_cf, 1), dtype=torch.float64, device=dev)
    ctx = _lib.context(dev.index if dev.index is not None else torch.cuda.current_device())
    rc = ctx.lib.coskad_frame_aggregate(ctx.h, score.data_ptr(), frames_d.data_ptr(), T, win_idx.data_ptr(),
                                        person_off.data_ptr(), person_clip.data_ptr(), person_out_off.data_ptr(),
                                        gi.n_persons, clip_person_off.data_ptr(), clip_off.data_ptr(), gi.n_clips,
                                        total_pf, max_cf, person_out.data_ptr(), out.data_ptr(), _lib.stream_ptr(dev))
    ctx.check(rc, 'coskad_frame_aggregate')
    return person_out, out


def score_and_aggregate(score: torch.Tensor, trans, meta, frames, clips: Sequence[Tuple[int, int, int]],
                        num_transform: int, pad_size: int = -1, gts: Optional[Dict[Tuple[int, int], np.ndarray]] = None,
                        smooth: bool = True, masks: Optional[Dict[Tuple[int, int], np.ndarray]] = None
                        ) -> Dict[int, List[np.ndarray]]:
    """Per transformation, the list of per-clip frame-score curves (float64), identical to what the
    loops of eval_COSKAD.py:140-220 produce from the same per-window scores.

    clips: (scene, clip, n_frames) in the sorted gt-file order.  masks: optional boolean frame masks
    (HR subsets, eval_COSKAD.py:213-215) applied before score_process."""
    gi = GroupIndex(trans, meta, clips, num_transform)
    person_out, out = aggregate_curves(score, frames, gi)
    ncl = gi.n_clip_per_transform
    res: Dict[int, List[np.ndarray]] = {}
    if pad_size != -1:
        pc = person_out.cpu().numpy()
    oc = out.cpu().numpy()
    for tr in range(num_transform):
        curves = []
        for ci, (scene, clip, F) in enumerate(clips):
            gc = tr * ncl + ci
            if pad_size != -1:
                p0, p1 = int(gi.clip_person_off[gc]), int(gi.clip_person_off[gc + 1])
                gt = gts[(scene, clip)] if gts is not None else np.zeros(F)
                per = [pad_scores(pc[gi.person_out_off[p]: gi.person_out_off[p + 1]].copy(), gt, pad_size)
                       for p in range(p0, p1)]
                cs = np.amax(np.stack(per, axis=0), axis=0)            # eval_COSKAD.py:211
            else:
                cs = oc[gi.clip_off[gc]: gi.clip_off[gc + 1]].copy()
            if masks is not None and (scene, clip) in masks:
                cs = cs[masks[(scene, clip)]]
            if smooth:
                cs = score_process(cs)
            curves.append(cs)
        res[tr] = curves
    return res


# ---- device post-processing: smoothing + AUC without leaving the GPU -------------------------------
def gaussian_weights(sigma: float = 30.0, truncate: float = 4.0) -> np.ndarray:
    """the normalised kernel scipy.ndimage.gaussian_filter1d(order 0) builds (scipy/ndimage/_filters.py _gaussian_kernel1d)"""
    radius = int(truncate * float(sigma) + 0.5)
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / (sigma * sigma) * x ** 2)
    return phi / phi.sum()


def score_process_device(curves: torch.Tensor, curve_off: torch.Tensor, sigma: float = SMOOTH_SIGMA, shift: int = SHIFT_FRAMES
                         ) -> torch.Tensor:
    """utils/eval_utils.py:200-207 for all curves at once on the device (float64, scipy's summation order)"""
    if not curves.is_cuda:
        raise _lib.CoskadError('curves must be a CUDA tensor (no CPU fallback; use score_process on the host)')
    dev = curves.device
    curves = curves.to(torch.float64).contiguous()
    curve_off = curve_off.to(device=dev, dtype=torch.int64).contiguous()
    w = torch.from_numpy(gaussian_weights(sigma)).to(dev)
    out = torch.empty_like(curves)
    ctx = _lib.context(dev.index if dev.index is not None else torch.cuda.current_device())
    rc = ctx.lib.coskad_score_process(ctx.h, curves.data_ptr(), curve_off.data_ptr(), curve_off.numel() - 1, int(shift),
                                      w.data_ptr(), (w.numel() - 1) // 2, out.data_ptr(), _lib.stream_ptr(dev))
    ctx.check(rc, 'coskad_score_process')
    return out


def auc_device(scores: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
    """sklearn.metrics.roc_auc_score(labels, scores) for binary labels as a 0-dim float64 device tensor: ROC points at the
    distinct score thresholds (descending), trapezoid rule (sklearn/metrics/_ranking.py _binary_clf_curve + auc)"""
    s = scores.to(torch.float64).view(-1)
    y = (labels.to(device=s.device).view(-1) != 0).to(torch.float64)
    order = torch.argsort(s, descending=True, stable=True)
    s, y = s[order], y[order]
    last = torch.ones_like(s, dtype=torch.bool)
    last[:-1] = s[1:] != s[:-1]
    idx = torch.nonzero(last).view(-1)
    tps = torch.cumsum(y, 0)[idx]
    fps = (idx + 1).to(torch.float64) - tps
    zero = torch.zeros(1, dtype=torch.float64, device=s.device)
    tpr = torch.cat([zero, tps]) / tps[-1]
    fpr = torch.cat([zero, fps]) / fps[-1]
    return torch.trapezoid(tpr, fpr)


def score_auc_device(score: torch.Tensor, trans, meta, frames, clips: Sequence[Tuple[int, int, int]], num_transform: int,
                     gts: Dict[Tuple[int, int], np.ndarray]) -> Tuple[float, Dict[int, float]]:
    """eval_COSKAD.py:140-253 without pad_scores / HR masks, entirely on the device: frame aggregation -> shift + Gaussian
    smoothing -> per-transformation AUC on the concatenated clips and the AUC of the mean curve.  One D2H copy of
    num_transform + 1 doubles at the end."""
    gi = DeviceGroupIndex(trans, meta, clips, num_transform, device=score.device)
    _, out = aggregate_curves(score, frames, gi)
    dev = out.device
    total = gi.total_clip_frames
    sm = score_process_device(out[:total], gi.clip_off)
    per = total // num_transform                                  # every transformation covers the same clips
    gt = torch.from_numpy(np.concatenate([np.asarray(gts[(s, c)]).reshape(-1) for (s, c, _f) in clips])).to(dev)
    st = sm.view(num_transform, per)
    aucs = torch.stack([auc_device(st[t], gt) for t in range(num_transform)] + [auc_device(st.mean(0), gt)]).cpu()
    return float(aucs[-1]), {t: float(aucs[t]) for t in range(num_transform)}


# ---- reference-named compat surface (utils/eval_utils.py:41-106) ------------------------------------
def _scatter(loss: torch.Tensor, frames_fig, n_frames: int) -> np.ndarray:
    """pose[n, frames_fig[n]-1] = loss[n] as a float64 [w, n_frames] matrix, one D2H copy instead of w."""
    l = loss.detach().to(torch.float32).cpu().numpy()
    w = l.shape[0]
    pose = np.zeros(shape=(w, n_frames))
    fr = np.asarray(frames_fig)
    rows = np.repeat(np.arange(w), fr.shape[1])
    pose[rows, (fr - 1).reshape(-1)] = np.repeat(l, fr.shape[1])
    return pose


def _classify_loss_fn(loss_fn) -> str:
    """'mse' or 'cosine': the two per-window losses the reference passes (eval_COSKAD.py:65 ``nn.MSELoss(reduction='none')``,
    :81 ``lambda x, y: 1 - F.cosine_similarity(x, y)``).  A callable is identified by what it computes on a CPU probe;
    anything else is rejected instead of being scored with the wrong kernel."""
    if loss_fn is None or isinstance(loss_fn, torch.nn.MSELoss):
        return 'mse'
    g = torch.Generator().manual_seed(0)
    a, b = torch.randn(3, 5, generator=g), torch.randn(3, 5, generator=g)
    try:
        got = loss_fn(a, b)
    except Exception as exc:
        raise NotImplementedError(f'loss_fn {loss_fn!r} is not callable on tensors: {exc}') from exc
    if torch.is_tensor(got) and got.shape == (3,) and torch.allclose(got, 1 - torch.nn.functional.cosine_similarity(a, b), atol=1e-6):
        return 'cosine'
    if torch.is_tensor(got) and got.shape == (3, 5) and torch.allclose(got, (a - b) ** 2, atol=1e-6):
        return 'mse'
    raise NotImplementedError('windows_based_loss_hy implements the two losses COSKAD uses (element-wise MSE, 1 - cosine '
                              f'similarity); {loss_fn!r} computes something else')


def windows_based_loss_hy(hidden_c, hidden_out_fig, frames_fig, n_frames, loss_fn=None, hyperbolic=False):
    """utils/eval_utils.py:57-74.  hyperbolic: gmath.dist(latents, c); else mean_d MSE(c, z) or the cosine score
    (eval_COSKAD.py:81); any other ``loss_fn`` raises NotImplementedError."""
    from . import gmath
    z = torch.as_tensor(hidden_out_fig).cuda() if not torch.is_tensor(hidden_out_fig) else hidden_out_fig.cuda()
    c = torch.as_tensor(hidden_c).cuda().view(-1)
    if hyperbolic:
        loss = gmath.dist(z, c, k=-1.0)
    elif _classify_loss_fn(loss_fn) == 'mse':
        loss = gmath.euclid_score(z, c)
    else:
        loss = gmath.cosine_score(z, c)
    return _scatter(loss, frames_fig, n_frames)


def windows_based_loss_mahalanobis(hidden_c, hidden_out_fig, VI, frames_fig, n_frames):
    """utils/eval_utils.py:41-55: per-window Mahalanobis distance to the center with the inverse covariance matrix ``VI``
    (``coskad_mahalanobis``), scattered to the windows' frames like windows_based_loss_hy"""
    from . import gmath
    z = torch.as_tensor(hidden_out_fig, dtype=torch.float32).cuda()
    c = torch.as_tensor(hidden_c, dtype=torch.float32).cuda().view(-1)
    vi = torch.as_tensor(VI, dtype=torch.float32).cuda()
    return _scatter(gmath.mahalanobis_score(z, c, vi), frames_fig, n_frames)


def windows_based_loss_rec_and_hy(gt_fig, out_fig, hidden_c, hidden_out_fig, frames_fig, n_frames, loss_fn=None,
                                  rec_loss_weight=0.2, loss_type='rec'):
    """utils/eval_utils.py:77-106."""
    from . import gmath
    assert len(gt_fig.shape) == 4
    z = torch.as_tensor(hidden_out_fig).cuda()
    c = torch.as_tensor(hidden_c).cuda().view(-1)
    w = gt_fig.shape[0]
    g = torch.as_tensor(gt_fig).cuda().reshape(w, -1)
    o = torch.as_tensor(out_fig).cuda().reshape(w, -1)
    loss = gmath.euclid_score(o, g)               # mean over (c,t,v) of (gt - out)^2 -- order independent
    loss_h = gmath.euclid_score(z, c)
    if loss_type == 'rec+hyp':
        loss = loss / rec_loss_weight + loss_h
    elif loss_type == 'hyp':
        loss = loss_h
    return _scatter(loss, frames_fig, n_frames)
