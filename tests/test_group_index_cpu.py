"""CPU: the host-side grouping that feeds the aggregation kernel reproduces the reference's
boolean-mask filtering order (eval_COSKAD.py:146-191)."""
import numpy as np

from coskad_b200.aggregate import GroupIndex, _scatter, pad_scores, score_process
from oracle import aggregate as oagg


def test_group_index_matches_reference_filtering():
    trans, meta, frames, clips, _ = oagg.synth_dataset(n_clips=7, seed=4, num_transform=3, max_persons=5)
    # windows of a clip that is not in the gt list must be ignored, like upstream
    meta = meta.copy()
    meta[::37, 1] = 99
    gi = GroupIndex(trans, meta, clips, 3)
    p = 0
    for t in range(3):
        ct = trans == t
        for ci, (scene, clip, F) in enumerate(clips):
            cc = ct & (meta[:, 0] == scene) & (meta[:, 1] == clip)
            gc = t * len(clips) + ci
            assert gi.clip_off[gc + 1] - gi.clip_off[gc] == F
            persons = sorted(set(meta[cc][:, 2]))
            assert gi.clip_person_off[gc + 1] - gi.clip_person_off[gc] == len(persons)
            for fig in persons:
                idx = np.nonzero(cc & (meta[:, 2] == fig))[0]          # dataset order
                got = gi.win_idx[gi.person_off[p]: gi.person_off[p + 1]]
                assert np.array_equal(got, idx)
                assert gi.person_clip[p] == gc and gi.person_id[p] == fig
                assert gi.person_out_off[p + 1] - gi.person_out_off[p] == F
                p += 1
    assert p == gi.n_persons


def test_empty_and_unknown():
    gi = GroupIndex(np.zeros(0, dtype=np.int64), np.zeros((0, 4), dtype=np.int64), [(1, 1, 10)], 2)
    assert gi.n_persons == 0 and gi.n_clips == 2 and gi.clip_off[-1] == 20


def test_host_postprocessing_equals_oracle():
    import torch
    rng = np.random.default_rng(0)
    s = rng.random(300)
    s[40:90] = 0
    assert np.array_equal(score_process(s.copy()), oagg.score_process(s.copy()))
    gt = np.zeros(300)
    assert np.array_equal(pad_scores(s.copy(), gt, 7), oagg.pad_scores(s.copy(), gt, 7))
    fr = np.stack([np.arange(a, a + 12) for a in (0, 1, 5, 30)])
    l = torch.tensor([0.1, 0.0, 0.3, 0.4])
    assert np.array_equal(_scatter(l, fr, 50), oagg.scatter_windows(l.numpy(), fr, 50))


def test_hr_mask_loader(tmp_path):
    """utils/model_utils.py:149-161 restated: {scene}_{clip}.npy boolean masks keyed by (scene, clip)"""
    from coskad_b200.tasks import hr_ubnormal
    m = np.array([True, False, True])
    np.save(tmp_path / '3_17.npy', m)
    np.save(tmp_path / '12_4.npy', ~m)
    got = hr_ubnormal(str(tmp_path / '*.npy'))
    assert sorted(got) == [(3, 17), (12, 4)]
    assert np.array_equal(got[(3, 17)], m) and np.array_equal(got[(12, 4)], ~m)


def test_device_group_index_equals_host_group_index():
    """the torch (device) construction yields the host CSR field by field -- run here on CPU tensors"""
    import torch
    from coskad_b200.aggregate import DeviceGroupIndex
    for seed, ntr in ((4, 3), (9, 1), (12, 5)):
        trans, meta, frames, clips, _ = oagg.synth_dataset(n_clips=6, seed=seed, num_transform=ntr, max_persons=5)
        meta = meta.copy()
        meta[::29, 1] = 77                       # windows of an unknown clip are dropped
        trans = trans.copy()
        trans[5::41] = ntr + 2                   # ... and of an out-of-range transformation
        perm = np.random.default_rng(seed).permutation(len(trans))      # dataset order is arbitrary
        trans, meta = trans[perm], meta[perm]
        h = GroupIndex(trans, meta, clips, ntr)
        d = DeviceGroupIndex(torch.from_numpy(trans), torch.from_numpy(meta), clips, ntr, device='cpu')
        assert d.n_persons == h.n_persons and d.n_clips == h.n_clips
        for f in ('win_idx', 'person_off', 'person_clip', 'person_id', 'clip_frames', 'clip_off', 'clip_person_off',
                  'person_out_off'):
            assert np.array_equal(getattr(d, f).numpy(), np.asarray(getattr(h, f))), f
        assert d.total_person_frames == int(h.person_out_off[-1]) and d.total_clip_frames == int(h.clip_off[-1])
    e = DeviceGroupIndex(torch.zeros(0, dtype=torch.int64), torch.zeros((0, 4), dtype=torch.int64), [(1, 1, 10)], 2, device='cpu')
    assert e.n_persons == 0 and e.n_clips == 2 and e.total_clip_frames == 20


def test_pad_scores_interval_form_equals_the_reference_loops():
    rng = np.random.default_rng(1)
    for trial in range(200):
        L = int(rng.integers(2, 120))
        s = rng.random(L)
        for _ in range(int(rng.integers(0, 5))):
            a = int(rng.integers(0, L)); b = int(rng.integers(a, L + 1))
            s[a:b] = 0
        if trial % 17 == 0:
            s[:] = 0
        pad = int(rng.integers(0, 9))
        gt = np.zeros(L)
        assert np.array_equal(pad_scores(s.copy(), gt, pad), oagg.pad_scores(s.copy(), gt, pad)), (trial, L, pad)
    from coskad_b200.aggregate import ranges
    for nums in ([], [3], [0, 1, 2, 5, 6, 9], [4, 2, 3, 10]):
        assert ranges(set(nums)) == oagg.ranges(set(nums))
    for n in (0, 5, 11, 12, 40):
        x = rng.random(n)
        assert np.array_equal(score_process(x.copy()), oagg.score_process(x.copy())) if n else True


def test_loss_fn_is_identified_by_what_it_computes():
    import pytest
    import torch
    import torch.nn.functional as F
    from coskad_b200.aggregate import _classify_loss_fn
    assert _classify_loss_fn(None) == 'mse' and _classify_loss_fn(torch.nn.MSELoss(reduction='none')) == 'mse'
    assert _classify_loss_fn(lambda x, y: 1 - F.cosine_similarity(x, y)) == 'cosine'        # eval_COSKAD.py:81
    assert _classify_loss_fn(lambda x, y: (x - y) ** 2) == 'mse'
    with pytest.raises(NotImplementedError):
        _classify_loss_fn(torch.nn.L1Loss(reduction='none'))
