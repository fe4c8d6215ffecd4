"""GPU parity of the device post-processing: shift + Gaussian smoothing bit-exact with
scipy.ndimage.gaussian_filter1d (the call of utils/eval_utils.py:200-207), rank AUC equal to sklearn's roc_auc_score, and the
all-device eval tail (aggregate -> smooth -> AUC) equal to the host tail of eval_COSKAD.py:213-253."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref_score_process(score):                       # utils/eval_utils.py:200-207
    from scipy.ndimage import gaussian_filter1d
    shifted = np.zeros_like(score)
    shift = 8 + (8 // 2) - 1
    shifted[shift:] = score[:-shift]
    return gaussian_filter1d(shifted, 30)


def test_smoothing_bit_exact_with_scipy():
    from coskad_b200 import aggregate
    rng = np.random.default_rng(0)
    lens = [1, 5, 11, 12, 13, 64, 119, 120, 121, 241, 300, 777, 1200, 2500]
    curves = [rng.random(n) * (rng.random(n) > 0.3) for n in lens]           # exact zeros = frames without a person
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    got = aggregate.score_process_device(torch.from_numpy(np.concatenate(curves)).cuda(), torch.from_numpy(off)).cpu().numpy()
    for i, c in enumerate(curves):
        ref = _ref_score_process(c)
        assert np.array_equal(got[off[i]:off[i + 1]], ref), f'curve of {lens[i]} frames: max abs diff {np.abs(got[off[i]:off[i + 1]] - ref).max():.3e}'


@pytest.mark.parametrize('ties', [False, True])
def test_auc_matches_sklearn(ties):
    from sklearn.metrics import roc_auc_score
    from coskad_b200 import aggregate
    rng = np.random.default_rng(1)
    y = (rng.random(20000) < 0.1).astype(np.int64)
    s = rng.random(20000) + 0.3 * y
    if ties:
        s = np.round(s, 2)
    got = float(aggregate.auc_device(torch.from_numpy(s).cuda(), torch.from_numpy(y).cuda()))
    assert abs(got - roc_auc_score(y, s)) < 1e-12


def test_device_eval_tail_equals_host_tail():
    from coskad_b200 import aggregate, tasks
    from coskad_b200.data import SyntheticPoseDataset
    ds = SyntheticPoseDataset(n_clips=7, seed=3, num_transform=5)
    g = torch.Generator().manual_seed(0)
    scores = (torch.rand(len(ds.x), generator=g) + 0.05).cuda()
    nt = 5
    curves = aggregate.score_and_aggregate(scores, ds.trans.numpy(), ds.meta.numpy(), ds.frames.numpy(), ds.clips, nt)
    auc_h, per_h = tasks.auc_from_curves(curves, ds.clips, ds.gts)
    auc_d, per_d = aggregate.score_auc_device(scores, ds.trans.numpy(), ds.meta.numpy(), ds.frames.numpy(), ds.clips, nt, ds.gts)
    assert abs(auc_h - auc_d) < 1e-12
    for t in range(nt):
        assert abs(per_h[t] - per_d[t]) < 1e-12
