"""TEST INFRASTRUCTURE ONLY — CPU baseline of the whole per-frame path (bench.py `cpu_baseline` / `--impl reference`).

Mirrors MOT.step (fastmot/mot.py:125-168) on the host: letterbox -> conv stack (fp32 PyTorch CPU; the reference's
TensorRT engines cannot run here) -> head decode -> filter + DIoU-NMS -> ReID crops (cv2) -> OSNet (PyTorch CPU) ->
OracleTracker (cv2 KLT, batched numpy Kalman, SciPy-equivalent assignment).  Uses every host thread the libraries
take by default (torch intra-op pool, OpenCV parallel_for).
"""
import time

import numpy as np
import torch

from fastmot_b200.models import darknet, osnet
from fastmot_b200 import models
from . import detect, nets
from .run import default_tracker_cfg
from .tracker import OracleTracker


class OraclePipeline:
    def __init__(self, size, yolo='YOLOv4CSP', reid='OSNet10', frame_skip=5, class_ids=(0,), head_obj_bias=-5.0,
                 detections_override=None, run_nets=True, embeddings_override=None, scene_ids=None, klt=True):
        # run_nets=False: the conv stacks are skipped (the reference runs them in TensorRT, never on its CPU); the
        # detector output comes from detections_override and the embeddings from embeddings_override(t, ids).
        # klt=False: KLT bypassed (flow result = the tracks' current boxes, identity homography): BASELINE configs[0].
        self.run_nets, self.embeddings_override, self.scene_ids, self.klt = run_nets, embeddings_override, scene_ids, klt
        self.size = size
        self.frame_skip = frame_skip
        self.ym = models.YOLO.get_model(yolo)
        self.rm = models.ReID.get_model(reid)
        self.layers = darknet.BUILDERS[self.ym.CFG](num_classes=self.ym.NUM_CLASSES,
                                                     anchors_per_head=len(self.ym.ANCHORS[0]) // 2)
        self.yw = darknet.synthetic_weights(self.layers, 3, head_obj_bias=head_obj_bias,
                                            num_classes=self.ym.NUM_CLASSES)
        self.ops = osnet.build_osnet(self.rm.ARCH[1], self.rm.OUTPUT_LAYOUT)
        self.rw = osnet.synthetic_weights(self.ops)
        self.label_mask = np.zeros(self.ym.NUM_CLASSES, bool)
        self.label_mask[list(class_ids)] = True
        c, h, w = self.ym.INPUT_SHAPE
        self.in_wh = (w, h)
        self.roi, self.up, self.off = detect.letterbox_geometry(size, self.in_wh, self.ym.LETTERBOX)
        self.tracker = OracleTracker(size, self.rm.METRIC, **default_tracker_cfg())
        self.tracker.reset(1 / 30.)
        self.frame_count = 0
        self.detections_override = detections_override
        self.stage_s = {}

    def _t(self, name, t0):
        self.stage_s[name] = self.stage_s.get(name, 0.0) + time.perf_counter() - t0

    def _detect(self, frame):
        if frame is None:          # association-only configuration: scripted detections, no image work
            return self.detections_override(self.frame_count)
        t0 = time.perf_counter()
        x = torch.as_tensor(detect.letterbox(frame, self.in_wh, self.roi))[None]
        self._t('preproc', t0)
        if self.run_nets:
            t0 = time.perf_counter()
            with torch.no_grad():
                heads = nets.run_darknet(self.layers, self.yw, x)
            self._t('yolo', t0)
            t0 = time.perf_counter()
            dec = [detect.yolo_decode(h.numpy(), a, s, self.in_wh, self.ym.NUM_CLASSES, self.ym.NEW_COORDS)
                   for h, a, s in zip(heads, self.ym.ANCHORS, self.ym.SCALES)]
            out = detect.filter_dets(np.concatenate(dec), self.up, self.off, self.label_mask, 0.25, 0.5, 800000, 1.2)
            self._t('nms', t0)
        else:
            # the reference's CPU share of the detector: class / score filter + DIoU-NMS over ALL K0 candidate rows
            # (detector.py:322-365).  Candidates: the scripted boxes, 5 jittered copies each above the threshold,
            # embedded in K0 low-score rows (K0 = the candidate count of this detector's heads).
            cand = self._synthetic_candidates(self.frame_count)
            t0 = time.perf_counter()
            out = detect.filter_dets(cand, self.up, self.off, self.label_mask, 0.25, 0.5, 800000, 1.2)
            self._t('nms', t0)
        if self.detections_override is not None:
            out = self.detections_override(self.frame_count)
        return out

    def _synthetic_candidates(self, t):
        k0 = sum(len(a) // 2 * (self.in_wh[0] // f) * (self.in_wh[1] // f)
                 for a, f in zip(self.ym.ANCHORS, self.ym.LAYER_FACTORS))
        rng = np.random.default_rng(t)
        cand = np.zeros((k0, 7), np.float32)
        cand[:, 4] = 0.01
        cand[:, 6] = 0.5
        tl, lb, cf = self.detections_override(t)
        up, off = np.asarray(self.up, np.float64), np.asarray(self.off, np.float64)
        rows = rng.choice(k0, 5 * len(tl), replace=False)
        for j in range(5):
            jit = rng.normal(0, 1.0, (len(tl), 2))
            x = (tl[:, 0] + off[0] + jit[:, 0]) / up[0]
            y = (tl[:, 1] + off[1] + jit[:, 1]) / up[1]
            w = (tl[:, 2] - tl[:, 0] + 1) / up[0]
            h = (tl[:, 3] - tl[:, 1] + 1) / up[1]
            r = rows[j * len(tl):(j + 1) * len(tl)]
            cand[r, 0], cand[r, 1], cand[r, 2], cand[r, 3] = x, y, w, h
            cand[r, 4] = 0.9 - 0.05 * j
            cand[r, 5] = lb
            cand[r, 6] = 1.0
        return cand

    def _embed(self, frame, tlbr):
        if len(tlbr) == 0:
            return np.zeros((0, self.rm.OUTPUT_LAYOUT), np.float32)
        if frame is None:
            return self.embeddings_override(self.frame_count, self.scene_ids(self.frame_count))
        t0 = time.perf_counter()
        crops = torch.as_tensor(detect.roi_preprocess(frame, tlbr))
        self._t('crops', t0)
        if not self.run_nets:
            ids = self.scene_ids(self.frame_count)
            return self.embeddings_override(self.frame_count, ids)
        t0 = time.perf_counter()
        with torch.no_grad():
            emb = nets.run_osnet(self.ops, self.rw, crops).numpy()
        self._t('osnet', t0)
        return emb

    def time_nets(self, frame, tlbr):
        """Seconds of one detector frame's conv stacks in fp32 PyTorch-CPU (NOT something the reference does)."""
        x = torch.as_tensor(detect.letterbox(frame, self.in_wh, self.roi))[None]
        crops = torch.as_tensor(detect.roi_preprocess(frame, tlbr))
        t0 = time.perf_counter()
        with torch.no_grad():
            nets.run_darknet(self.layers, self.yw, x)
        t1 = time.perf_counter()
        with torch.no_grad():
            nets.run_osnet(self.ops, self.rw, crops)
        t2 = time.perf_counter()
        return {"yolo": round(t1 - t0, 3), "osnet": round(t2 - t1, 3)}

    def _flow(self, frame):
        trk = self.tracker
        if self.klt:
            trk.compute_flow(frame)
        else:
            klt = {k: v.tlbr for k, v in trk.tracks.items()}
            trk.compute_flow(frame, injected=(klt, np.eye(3), {k: 1.0 for k in klt}))

    def step(self, frame):
        trk = self.tracker
        if self.frame_count == 0:
            tlbr, labels, conf = self._detect(frame)
            trk.init(frame, tlbr, labels)
        elif self.frame_count % self.frame_skip == 0:
            dets = self._detect(frame)
            t0 = time.perf_counter()
            self._flow(frame)
            self._t('flow', t0)
            emb = self._embed(frame, dets[0])
            t0 = time.perf_counter()
            trk.apply_kalman()
            self._t('kalman', t0)
            t0 = time.perf_counter()
            trk.update(self.frame_count, dets[0], dets[1], dets[2], emb)
            self._t('assoc', t0)
        else:
            t0 = time.perf_counter()
            self._flow(frame)
            self._t('flow', t0)
            t0 = time.perf_counter()
            trk.apply_kalman()
            self._t('kalman', t0)
        self.frame_count += 1

    def visible(self):
        return self.tracker.visible()
