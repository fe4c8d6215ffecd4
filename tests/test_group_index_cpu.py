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
