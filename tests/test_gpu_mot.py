"""GPU: the MOT.step scheduler (shared frame upload, detector stream, ReID batch, tracker) on a golden sequence.
Detector and ReID outputs are replaced by the scripted ones (random weights cannot detect) AFTER both networks ran,
so the visible tracks must match the reference golden (same IDs, boxes within +-1 px)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def test_mot_step_schedule_matches_reference_golden():
    from types import SimpleNamespace as NS
    from fastmot_b200 import MOT, DET_DTYPE
    from fastmot_b200.synth import SyntheticScene
    from oracle.run import default_tracker_cfg
    g = np.load(os.path.join(GOLDEN, "seq_T64.npz"))
    scene = SyntheticScene(**eval(str(g['scene_kw'])))

    def dets(t):
        tl, lb, cf, _ = scene.detections(t)
        d = np.zeros(len(tl), DET_DTYPE)
        d['tlbr'], d['label'], d['conf'] = tl, lb, cf
        return d.view(np.recarray)

    def embs(t, d):
        return scene.embeddings(scene.detections(t)[3], t)

    mot = MOT(scene.size, detector_frame_skip=5, class_ids=(0,),
              yolo_detector_cfg=NS(model='YOLOv4Tiny'), feature_extractor_cfgs=(NS(model='OSNet025'),),
              tracker_cfg=NS(**default_tracker_cfg()), detections_override=dets, embeddings_override=embs)
    mot.reset(1 / 30)
    for t in range(17):
        frame = scene.frame(t)
        mot.step(frame)
        assert mot.frame_count == t + 1
        vis = {trk.trk_id: trk.tlbr for trk in mot.visible_tracks()}
        want = dict(zip(g[f'vis_ids_{t}'].tolist(), g[f'vis_tlbr_{t}']))
        assert set(vis) == set(want), (t, set(vis) ^ set(want))
        for k in vis:
            assert np.abs(vis[k] - want[k]).max() <= 1.0, (t, k)
    MOT.print_timing_info()
    assert mot.detector.last_num_candidates >= 0


def test_mot_real_dataflow_runs_and_tracks():
    """No embedding override: OSNet embeddings of the real crops drive the association (ids must stay stable)."""
    from types import SimpleNamespace as NS
    from fastmot_b200 import MOT, DET_DTYPE
    from fastmot_b200.synth import SyntheticScene
    from oracle.run import default_tracker_cfg
    scene = SyntheticScene(40, seed=12, label=0, dropout_frames=())

    def dets(t):
        tl, lb, cf, _ = scene.detections(t)
        d = np.zeros(len(tl), DET_DTYPE)
        d['tlbr'], d['label'], d['conf'] = tl, lb, cf
        return d.view(np.recarray)

    mot = MOT(scene.size, detector_frame_skip=5, class_ids=(0,), yolo_detector_cfg=NS(model='YOLOv4Tiny'),
              feature_extractor_cfgs=(NS(model='OSNet025'),), tracker_cfg=NS(**default_tracker_cfg()),
              detections_override=dets)
    mot.reset(1 / 30)
    ids = None
    for t in range(16):
        mot.step(scene.frame(t))
        if t >= 5:
            cur = sorted(trk.trk_id for trk in mot.visible_tracks())
            assert len(cur) == 40
            ids = ids or cur
            assert cur == ids
