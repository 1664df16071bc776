"""TEST INFRASTRUCTURE ONLY. Drives the *imported reference* MultiTracker (container only).

Mirrors the schedule of the reference MOT.step (fastmot/mot.py:134-164) with scripted detections
and embeddings, because MOT itself needs TensorRT (SURVEY.md §8c gotcha 4).
"""
import numpy as np

from .refshim import load_reference, reference_config


def make_dets(fm, tlbr, labels, conf):
    arr = np.zeros(len(tlbr), fm.detector.DET_DTYPE)
    arr['tlbr'] = tlbr
    arr['label'] = labels
    arr['conf'] = conf
    return arr.view(np.recarray)


def run_reference_tracker(scene, n_frames, frame_skip=5, metric='cosine', emb_noise=0.0,
                          capture=None, tracker_kwargs=None):
    """Returns list (per frame) of dict(ids=int64[n], tlbr=f64[n,4]) of confirmed+active tracks,
    plus the tracker object. `capture(t, trk, phase)` is an optional hook."""
    fm = load_reference()
    cfg = reference_config()
    kw = vars(cfg.mot_cfg.tracker_cfg).copy()
    if tracker_kwargs:
        kw.update(tracker_kwargs)
    trk = fm.MultiTracker(scene.size, metric, **kw)
    trk.reset(1.0 / 30)
    out = []
    for t in range(n_frames):
        frame = scene.frame(t)
        if t == 0:
            tlbr, labels, conf, ids = scene.detections(t)
            trk.init(frame, make_dets(fm, tlbr, labels, conf))
        elif t % frame_skip == 0:
            trk.compute_flow(frame)
            if capture:
                capture(t, trk, 'flow')
            trk.apply_kalman()
            if capture:
                capture(t, trk, 'kalman')
            tlbr, labels, conf, ids = scene.detections(t)
            emb = scene.embeddings(ids, t, emb_noise)
            trk.update(t, make_dets(fm, tlbr, labels, conf), emb)
            if capture:
                capture(t, trk, 'update')
        else:
            trk.compute_flow(frame)
            if capture:
                capture(t, trk, 'flow')
            trk.apply_kalman()
            if capture:
                capture(t, trk, 'kalman')
        vis = [(k, v.tlbr.copy()) for k, v in trk.tracks.items() if v.confirmed and v.active]
        out.append(dict(ids=np.array([k for k, _ in vis], np.int64),
                        tlbr=np.array([b for _, b in vis], np.float64).reshape(-1, 4)))
    return out, trk
