"""GPU parity: the product MultiTracker (device Kalman + cost + LSA cascade) replayed on the synthetic
sequences with KLT bypassed (the reference's own klt boxes / homography injected) must reproduce the
reference's visible track IDs and boxes EXACTLY, frame by frame (SURVEY.md §8c tier T3-bypassed)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def _dets(tlbr, labels, conf):
    dt = np.dtype([('tlbr', float, 4), ('label', int), ('conf', float)], align=True)
    arr = np.zeros(len(tlbr), dt)
    arr['tlbr'], arr['label'], arr['conf'] = tlbr, labels, conf
    return arr.view(np.recarray)


def tracker_cfg():
    """cfg/mot.json:44-96 of the reference (tracker_cfg)."""
    from types import SimpleNamespace as NS
    return dict(max_age=6, age_penalty=2, motion_weight=0.2, max_assoc_cost=0.8, max_reid_cost=0.6, iou_thresh=0.4,
                duplicate_thresh=0.8, occlusion_thresh=0.7, conf_thresh=0.5, confirm_hits=1, history_size=50,
                kalman_filter_cfg=NS(std_factor_acc=2.25, std_offset_acc=78.5, std_factor_det=(0.08, 0.08),
                                     std_factor_klt=(0.14, 0.14), min_std_det=(4.0, 4.0), min_std_klt=(5.0, 5.0),
                                     init_pos_weight=5, init_vel_weight=12, vel_coupling=0.6, vel_half_life=2))


@pytest.mark.parametrize("name", ["seq_T64.npz", "seq_T200.npz", "seq_T70_overlap.npz"])
def test_sequence_ids_and_boxes_exact(name):
    from fastmot_b200 import MultiTracker
    from fastmot_b200.synth import SyntheticScene
    g = np.load(os.path.join(GOLDEN, name))
    scene = SyntheticScene(**eval(str(g['scene_kw'])))
    n_frames, skip = int(g['n_frames']), int(g['frame_skip'])
    trk = MultiTracker(scene.size, str(g['metric']), **tracker_cfg())
    trk.reset(1 / 30)
    for t in range(n_frames):
        if t == 0:
            tlbr, labels, conf, ids = scene.detections(0)
            trk.init(None, _dets(tlbr, labels, conf))
        else:
            H = g[f'H_{t}']
            klt = {int(k): b for k, b in zip(g[f'klt_ids_{t}'], g[f'klt_tlbr_{t}'])}
            rat = {int(k): float(r) for k, r in zip(g[f'klt_ids_{t}'], g[f'klt_ratio_{t}'])}
            trk.inject_flow(klt, None if H.size == 0 else H, rat)
            trk.compute_flow(None)
            trk.apply_kalman()
            ids_k = np.array(list(trk.tracks.keys()), np.int64)
            assert np.array_equal(ids_k, g[f'kal_ids_{t}']), f"frame {t}: track set after kalman"
            got = np.array([trk.tracks[k].tlbr for k in ids_k]).reshape(-1, 4)
            assert np.array_equal(got, g[f'kal_tlbr_{t}']), f"frame {t}: boxes after kalman"
            if t % skip == 0:
                tlbr, labels, conf, ids = scene.detections(t)
                trk.update(t, _dets(tlbr, labels, conf), scene.embeddings(ids, t))
                ids_u = np.array(list(trk.tracks.keys()), np.int64)
                assert np.array_equal(ids_u, g[f'upd_ids_{t}']), f"frame {t}: ids after update"
                assert np.array_equal([trk.tracks[k].age for k in ids_u], g[f'upd_age_{t}'])
                assert np.array_equal([trk.tracks[k].hits for k in ids_u], g[f'upd_hits_{t}'])
                assert np.array_equal(np.array(list(trk.hist_tracks.keys()), np.int64), g[f'upd_hist_{t}'])
        vis = [(k, v.tlbr) for k, v in trk.tracks.items() if v.confirmed and v.active]
        assert np.array_equal(np.array([k for k, _ in vis], np.int64), g[f'vis_ids_{t}']), f"frame {t}"
        assert np.array_equal(np.array([b for _, b in vis]).reshape(-1, 4), g[f'vis_tlbr_{t}']), f"frame {t}"
    ids = g['final_ids']
    mean = np.array([trk.tracks[int(k)].state[0] for k in ids])
    np.testing.assert_allclose(mean, g['final_mean'], atol=1e-6)
    cov = np.array([trk.tracks[int(k)].state[1] for k in ids])
    np.testing.assert_allclose(cov, g['final_cov'], rtol=1e-7, atol=1e-6)
    for k, cnt, avg in zip(ids, g['final_cnt'], g['final_avg']):
        assert trk.tracks[int(k)].avg_feat.count == cnt
        if cnt:
            np.testing.assert_allclose(trk.tracks[int(k)].avg_feat(), avg, atol=1e-5)


@pytest.mark.parametrize("name", ["seq_T64.npz", "seq_T200.npz", "seq_T70_overlap.npz"])
def test_sequence_ids_and_boxes_exact_per_stage_cascade(name, monkeypatch):
    """Same check with the fused cascade kernel off (one cost + assignment launch and one D2H per stage: the path
    frames with more than 256 detections take)."""
    monkeypatch.setenv("FM_FUSE_CASCADE", "0")
    test_sequence_ids_and_boxes_exact(name)


@pytest.mark.parametrize("name", ["seq_T64.npz", "seq_T200.npz", "seq_T70_overlap.npz"])
def test_sequence_ids_and_boxes_exact_always_fused_cascade(name, monkeypatch):
    """... and with fm_assoc_cascade forced for every update (by default single-stage frames keep the per-stage path)."""
    monkeypatch.setenv("FM_FUSE_CASCADE", "2")
    test_sequence_ids_and_boxes_exact(name)
