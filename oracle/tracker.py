"""TEST INFRASTRUCTURE ONLY (checker + CPU baseline; never shipped, never imported by fastmot_b200).

CPU restatement of the reference tracker path: OracleFlow (fastmot/flow.py:135-264, on top of OpenCV 4.13 — the
reference's own third-party arithmetic, flow.py:95,129-131,153-154,171-173,187-190,205-207,223-226,245-248) and
OracleTracker (fastmot/tracker.py:121-401 + fastmot/track.py:91-225) with a batched Kalman filter
(oracle/kalman.py) and the association primitives of oracle/assoc.py.

Pinned against the unmodified reference frame by frame in tests/test_oracle_vs_reference.py (container only) and
against tests/golden/seq_*.npz everywhere.  Container orders (dict / OrderedDict / set / sort stability) are kept
identical to the reference because they are observable.
"""
from collections import OrderedDict
import itertools

import numpy as np

from . import assoc
from .kalman import KalmanOracle, FLOW, DETECTOR


class OTrack:
    _count = 0

    def __init__(self, frame_id, tlbr, state, label, confirm_hits=1):
        OTrack._count += 1
        self.trk_id = OTrack._count
        self.start_frame = frame_id
        self.end_frame = frame_id
        self.tlbr = np.asarray(tlbr, np.float64)
        self.mean, self.cov = state
        self.label = label
        self.confirm_hits = confirm_hits
        self.age = 0
        self.hits = 0
        self.f_sum = None
        self.f_avg = None
        self.f_count = 0
        self.inlier_ratio = 1.
        self.keypoints = np.empty((0, 2), np.float32)
        self.prev_keypoints = np.empty((0, 2), np.float32)

    def __lt__(self, other):
        return (self.tlbr[-1], -self.age) < (other.tlbr[-1], -other.age)

    @property
    def active(self):
        return self.age < 2

    @property
    def confirmed(self):
        return self.hits >= self.confirm_hits

    def feat_update(self, vec, merged_count=None):
        """AverageFeature.update / merge arithmetic (track.py:100-126), float32 like the reference."""
        if merged_count is None:
            self.f_count += 1
        else:
            self.f_count = merged_count
        if self.f_sum is None:
            self.f_sum = vec.copy()
            self.f_avg = vec.copy()
        else:
            self.f_sum = self.f_sum + vec
            avg = (self.f_sum.astype(np.float64) * (1. / self.f_count)).astype(self.f_sum.dtype)
            nrm = np.linalg.norm(avg)
            self.f_avg = (avg.astype(np.float64) * (1. / float(nrm))).astype(self.f_sum.dtype)


class OracleFlow:
    def __init__(self, size, bg_feat_scale_factor=(0.1, 0.1), opt_flow_scale_factor=(0.5, 0.5), feat_density=0.005,
                 feat_dist_factor=0.06, ransac_max_iter=500, ransac_conf=0.99, max_error=100, inlier_thresh=4,
                 bg_feat_thresh=10, obj_feat_params=None, opt_flow_params=None):
        import cv2
        self.cv2 = cv2
        self.size = size
        self.bg_scale = bg_feat_scale_factor
        self.of_scale = opt_flow_scale_factor
        self.feat_density = feat_density
        self.feat_dist_factor = feat_dist_factor
        self.ransac_max_iter = ransac_max_iter
        self.ransac_conf = ransac_conf
        self.max_error = max_error
        self.inlier_thresh = inlier_thresh
        self.obj_feat_params = dict(maxCorners=1000, qualityLevel=0.06, blockSize=3)
        if obj_feat_params is not None:
            self.obj_feat_params.update(vars(obj_feat_params))
        self.lk_params = dict(winSize=(5, 5), maxLevel=5, criteria=(3, 10, 0.03))   # see flow.py:85-93
        self.fast = cv2.FastFeatureDetector_create(threshold=bg_feat_thresh)
        W, H = size
        self.of_sz = (round(self.of_scale[0] * W), round(self.of_scale[1] * H))
        self.bg_sz = (round(self.bg_scale[0] * W), round(self.bg_scale[1] * H))
        self.frame_rect = np.array([0., 0., W - 1., H - 1.])
        self.prev_gray = None
        self.prev_small = None
        self.bg_keypoints = np.empty((0, 2), np.float32)
        self.prev_bg_keypoints = np.empty((0, 2), np.float32)
        self.debug = {}

    def _prep(self, frame):
        cv2 = self.cv2
        gray = cv2.cvtColor(frame, cv2.COLOR_BGR2GRAY)
        return gray, cv2.resize(gray, self.of_sz)

    def init(self, frame):
        self.prev_gray, self.prev_small = self._prep(frame)
        self.bg_keypoints = np.empty((0, 2), np.float32)
        self.prev_bg_keypoints = np.empty((0, 2), np.float32)

    @staticmethod
    def _clip(tlbr, rect):
        out = np.array([max(tlbr[0], rect[0]), max(tlbr[1], rect[1]), min(tlbr[2], rect[2]), min(tlbr[3], rect[3])])
        return None if (out[2] < out[0] or out[3] < out[1]) else out

    @staticmethod
    def _view(img, tlbr):
        x0, y0, x1, y1 = (max(int(v), 0) for v in tlbr)
        return img[y0:y1 + 1, x0:x1 + 1]

    def predict(self, frame, tracks):
        cv2 = self.cv2
        W, H = self.size
        gray, small = self._prep(frame)
        tracks.sort(reverse=True)
        fg = np.full((H, W), 255, np.uint8)
        chunks = []
        for trk in tracks:
            inside = self._clip(trk.tlbr, self.frame_rect)
            tmask = self._view(fg, inside)
            area = int(np.count_nonzero(tmask))
            kps = trk.keypoints
            if len(kps):
                pi = np.rint(kps).astype(np.int32)
                ok = (pi[:, 0] >= inside[0]) & (pi[:, 0] <= inside[2]) & (pi[:, 1] >= inside[1]) & (pi[:, 1] <= inside[3])
                kps, pi = kps[ok], pi[ok]
                kps = kps[fg[pi[:, 1], pi[:, 0]] == 255] if len(kps) else kps
            if len(kps) < self.feat_density * area:
                img = self._view(self.prev_gray, inside)
                md = max(int(round(np.sqrt(area) * self.feat_dist_factor)), 1)
                found = cv2.goodFeaturesToTrack(img, mask=tmask, minDistance=md, **self.obj_feat_params)
                if found is None:
                    kps = np.empty((0, 2), np.float32)
                else:
                    pts = found.reshape(-1, 2) + np.asarray(inside[:2], np.float32)
                    c = np.array([(trk.tlbr[0] + trk.tlbr[2]) / 2, (trk.tlbr[1] + trk.tlbr[3]) / 2])
                    ax = np.array([trk.tlbr[2] - trk.tlbr[0] + 1, trk.tlbr[3] - trk.tlbr[1] + 1]) * 0.5
                    kps = pts[np.sum(((pts - c) / ax) ** 2, axis=1) <= 1.]
            chunks.append(kps.astype(np.float32).reshape(-1, 2))
            tmask[:] = 0
        ends = list(itertools.accumulate(len(c) for c in chunks)) if chunks else [0]
        begins = [0] + ends[:-1]

        def swap():
            self.prev_gray, self.prev_small = gray, small

        bg_img = cv2.resize(self.prev_gray, self.bg_sz)
        bg_mask = cv2.resize(fg, self.bg_sz, interpolation=cv2.INTER_NEAREST)
        kp = self.fast.detect(bg_img, mask=bg_mask)
        if len(kp) == 0:
            self.bg_keypoints = np.empty((0, 2), np.float32)
            swap()
            return {}, None
        bg_pts = np.float32([k.pt for k in kp]) * (np.float32(1) / np.asarray(self.bg_scale, np.float32))
        bg_begin = ends[-1]
        allp = np.concatenate(chunks + [bg_pts]).astype(np.float32)
        scaled = (allp * np.asarray(self.of_scale, np.float32)).reshape(-1, 1, 2)
        cur, st, err = cv2.calcOpticalFlowPyrLK(self.prev_small, small, scaled, None, **self.lk_params)
        st = st.ravel().astype(bool) & (err.ravel() < self.max_error)
        cur = cur.reshape(-1, 2)
        cur[st] = cur[st] * (np.float32(1) / np.asarray(self.of_scale, np.float32))
        self.debug = dict(all_prev=allp, all_cur=cur.copy(), status=st.copy(), err=err.ravel().copy(),
                          begins=list(begins), ends=list(ends), bg_begin=bg_begin)
        swap()

        sel = np.nonzero(st[bg_begin:-1])[0]
        pb, mb = allp[bg_begin:-1][sel], cur[bg_begin:-1][sel]
        if len(mb) < 4:
            self.bg_keypoints = np.empty((0, 2), np.float32)
            return {}, None
        Hm, mask = cv2.findHomography(pb, mb, method=cv2.RANSAC, maxIters=self.ransac_max_iter,
                                      confidence=self.ransac_conf)
        if Hm is None:
            self.bg_keypoints = np.empty((0, 2), np.float32)
            return {}, None
        inl = mask.ravel().astype(bool)
        self.prev_bg_keypoints, self.bg_keypoints = pb[inl], mb[inl]
        if len(self.bg_keypoints) < self.inlier_thresh:
            self.bg_keypoints = np.empty((0, 2), np.float32)
            return {}, None

        out = {}
        fg[:] = 255
        for b, e, trk in zip(begins, ends, tracks):
            sel = np.nonzero(st[b:e])[0]
            pp, mp = allp[b:e][sel], cur[b:e][sel]
            if len(mp):
                pi = np.rint(mp).astype(np.int32)
                ok = (pi[:, 0] >= 0) & (pi[:, 1] >= 0) & (pi[:, 0] < W) & (pi[:, 1] < H)
                pp, mp, pi = pp[ok], mp[ok], pi[ok]
                keep = fg[pi[:, 1], pi[:, 0]] == 255 if len(pi) else np.zeros(0, bool)
                pp, mp = pp[keep], mp[keep]
            if len(mp) < 3:
                trk.keypoints = np.empty((0, 2), np.float32)
                continue
            A, mask = cv2.estimateAffinePartial2D(pp, mp, method=cv2.RANSAC, maxIters=self.ransac_max_iter,
                                                  confidence=self.ransac_conf)
            if A is None:
                trk.keypoints = np.empty((0, 2), np.float32)
                continue
            tl = A @ np.array([trk.tlbr[0], trk.tlbr[1], 1.])
            sc = np.linalg.norm(A[:2, 0])
            sc = 1. if sc < 0.9 or sc > 1.1 else sc
            w, h = trk.tlbr[2] - trk.tlbr[0] + 1, trk.tlbr[3] - trk.tlbr[1] + 1
            est = np.rint([tl[0], tl[1], tl[0] + w * sc - 1., tl[1] + h * sc - 1.])
            inl = mask.ravel().astype(bool)
            trk.prev_keypoints, trk.keypoints = pp[inl], mp[inl]
            if self._clip(est, self.frame_rect) is None or len(trk.keypoints) < self.inlier_thresh:
                trk.keypoints = np.empty((0, 2), np.float32)
                continue
            out[trk.trk_id] = est
            trk.inlier_ratio = len(trk.keypoints) / len(mp)
            self._view(fg, est)[:] = 0
        return out, Hm


class OracleTracker:
    def __init__(self, size, metric, max_age=6, age_penalty=2, motion_weight=0.2, max_assoc_cost=0.9,
                 max_reid_cost=0.45, iou_thresh=0.4, duplicate_thresh=0.8, occlusion_thresh=0.7, conf_thresh=0.5,
                 confirm_hits=1, history_size=50, kalman_filter_cfg=None, flow_cfg=None, use_flow=True):
        self.size = size
        self.metric = metric.lower()
        self.max_age, self.age_penalty, self.motion_weight = max_age, age_penalty, motion_weight
        self.max_assoc_cost, self.max_reid_cost = max_assoc_cost, max_reid_cost
        self.iou_thresh, self.duplicate_thresh, self.occlusion_thresh = iou_thresh, duplicate_thresh, occlusion_thresh
        self.conf_thresh, self.confirm_hits, self.history_size = conf_thresh, confirm_hits, history_size
        self.kf = KalmanOracle(**(vars(kalman_filter_cfg) if kalman_filter_cfg is not None else {}))
        self.flow = OracleFlow(size, **(vars(flow_cfg) if flow_cfg is not None else {})) if use_flow else None
        self.tracks = {}
        self.hist_tracks = OrderedDict()
        self.frame_rect = np.array([0., 0., size[0] - 1., size[1] - 1.])
        self.klt_bboxes = {}
        self.homography = None
        self.trace = []

    def reset(self, dt):
        self.kf.reset_dt(dt)
        self.hist_tracks.clear()
        OTrack._count = 0

    def _spawn(self, frame_id, tlbr, label):
        m, c = self.kf.create(tlbr[None])
        t = OTrack(frame_id, tlbr, (m[0], c[0]), label, self.confirm_hits)
        self.tracks[t.trk_id] = t

    def init(self, frame, tlbr, labels):
        self.tracks.clear()
        if self.flow is not None and frame is not None:     # frame None: KLT-bypassed association runs (no image)
            self.flow.init(frame)
        for b, l in zip(tlbr, labels):
            self._spawn(0, np.asarray(b, np.float64), int(l))

    def compute_flow(self, frame, injected=None):
        if injected is not None:
            self.klt_bboxes, self.homography, ratios = injected
            for k, r in ratios.items():
                if k in self.tracks:
                    self.tracks[k].inlier_ratio = r
        else:
            active = [t for t in self.tracks.values() if t.active]
            self.klt_bboxes, self.homography = self.flow.predict(frame, active)
        if self.homography is None:
            self.tracks.clear()

    def apply_kalman(self):
        items = list(self.tracks.items())
        if not items:
            return
        mean = np.array([t.mean for _, t in items])
        cov = np.array([t.cov for _, t in items])
        mean, cov = self.kf.warp(mean, cov, self.homography)
        mean, cov = self.kf.predict(mean, cov)
        has = np.array([k in self.klt_bboxes for k, _ in items])
        if has.any():
            z = np.array([self.klt_bboxes[k] for k, _ in items if k in self.klt_bboxes])
            mult = np.array([max(self.age_penalty * t.age, 1) / t.inlier_ratio for k, t in items if k in self.klt_bboxes])
            m2, c2 = self.kf.update(mean[has], cov[has], z, FLOW, mult)
            mean[has], cov[has] = m2, c2
        boxes = np.rint(mean[:, :4])
        lost = assoc.ios(boxes, self.frame_rect) < 0.5
        for i, (k, t) in enumerate(items):
            t.mean, t.cov, t.tlbr = mean[i], cov[i], boxes[i]
            if lost[i]:
                self._mark_lost(k)

    def _mark_lost(self, k):
        t = self.tracks.pop(k)
        if t.confirmed:
            self.hist_tracks[k] = t
            if len(self.hist_tracks) > self.history_size:
                self.hist_tracks.popitem(last=False)

    def _cost_feat(self, ids, det_tlbr, det_label, emb, occ):
        n, m = len(ids), len(det_tlbr)
        if n == 0 or m == 0:
            return np.empty((n, m))
        trks = [self.tracks[i] for i in ids]
        feats = np.zeros((n, emb.shape[1]))
        invalid = np.zeros(n, bool)
        for i, t in enumerate(trks):
            if t.f_count > 0:
                feats[i] = t.f_avg
            else:
                invalid[i] = True
        fill = min(self.max_assoc_cost + 0.1, 1.)
        with np.errstate(all='ignore'):
            c = assoc.cdist(feats, emb, self.metric, invalid[:, None] | occ[None, :], fill)
        md = self.kf.motion_distance(np.array([t.mean for t in trks]), np.array([t.cov for t in trks]), det_tlbr)
        c = assoc.fuse_motion(c, md, self.motion_weight)
        return assoc.gate_cost(c, [t.label for t in trks], det_label, self.max_assoc_cost)

    def _cost_iou(self, ids, det_tlbr, det_label):
        n, m = len(ids), len(det_tlbr)
        if n == 0 or m == 0:
            return np.empty((n, m))
        trks = [self.tracks[i] for i in ids]
        c = assoc.iou_dist(np.array([t.tlbr for t in trks]), det_tlbr)
        return assoc.gate_cost(c, [t.label for t in trks], det_label, 1. - self.iou_thresh)

    def update(self, frame_id, det_tlbr, det_label, det_conf, emb):
        det_tlbr = np.asarray(det_tlbr, np.float64).reshape(-1, 4)
        det_label = np.asarray(det_label).reshape(-1)
        emb = np.asarray(emb)
        occ = assoc.find_occluded(det_tlbr, self.occlusion_thresh)
        n_depth = (self.max_age + 2) // 2
        by_depth = [[] for _ in range(n_depth)]
        unconfirmed = []
        for k, t in self.tracks.items():
            (by_depth[t.age // 2] if t.confirmed else unconfirmed).append(k)
        m1, u1 = [], []
        u_det = list(range(len(det_tlbr)))
        for depth, ids in enumerate(by_depth):
            if not u_det:
                u1.extend(itertools.chain.from_iterable(by_depth[depth:]))
                break
            if not ids:
                continue
            c = self._cost_feat(ids, det_tlbr[u_det], det_label[u_det], emb[u_det], occ[u_det])
            self.trace.append(('feat', frame_id, c.copy(), list(ids), list(u_det)))
            m, ut, u_det = assoc.linear_assignment(c, ids, u_det)
            m1 += m
            u1 += ut
        active = [k for k in u1 if self.tracks[k].active]
        u1 = [k for k in u1 if not self.tracks[k].active]
        c = self._cost_iou(active, det_tlbr[u_det], det_label[u_det])
        m2, u2, u_det = assoc.linear_assignment(c, active, u_det)
        c = self._cost_iou(unconfirmed, det_tlbr[u_det], det_label[u_det])
        m3, u3, u_det = assoc.linear_assignment(c, unconfirmed, u_det)

        hist_ids = [k for k, t in self.hist_tracks.items() if t.f_count >= 2]
        u_det = [d for d in u_det if det_conf[d] >= self.conf_thresh]
        valid = [d for d in u_det if not occ[d]]
        invalid = [d for d in u_det if occ[d]]
        if hist_ids and valid:
            feats = np.array([self.hist_tracks[k].f_avg for k in hist_ids], np.float64)
            with np.errstate(all='ignore'):
                c = assoc.cdist(feats, emb[valid], self.metric)
            labels = list(itertools.islice((t.label for t in self.hist_tracks.values()), len(hist_ids)))
            c = assoc.gate_cost(c, labels, det_label[valid])
        else:
            c = np.empty((len(hist_ids), len(valid)))
        reid, _, reid_u = assoc.greedy_match(c, hist_ids, valid, self.max_reid_cost)

        matches = set(itertools.chain(m1, m2, m3))
        u_trk = set(itertools.chain(u1, u2, u3))
        matches, u_trk = self._rectify(matches, u_trk, det_tlbr)

        for k, d in reid:
            t = self.hist_tracks.pop(k)
            m, cv = self.kf.create(det_tlbr[d][None])
            t.mean, t.cov = m[0], cv[0]
            t.start_frame = t.end_frame = frame_id
            t.tlbr = det_tlbr[d].copy()
            t.feat_update(emb[d])
            t.age = 0
            t.keypoints = np.empty((0, 2), np.float32)
            t.prev_keypoints = np.empty((0, 2), np.float32)
            self.tracks[k] = t
        ml = list(matches)
        if ml:
            mean = np.array([self.tracks[k].mean for k, _ in ml])
            cov = np.array([self.tracks[k].cov for k, _ in ml])
            mean, cov = self.kf.update(mean, cov, det_tlbr[[d for _, d in ml]], DETECTOR)
            boxes = np.rint(mean[:, :4])
            lost = assoc.ios(boxes, self.frame_rect) < 0.5
            for i, (k, d) in enumerate(ml):
                t = self.tracks[k]
                ok = not occ[d]
                if lost[i]:
                    ok = False
                    self._mark_lost(k)
                t.end_frame = frame_id
                t.tlbr, t.mean, t.cov = boxes[i], mean[i], cov[i]
                if ok:
                    t.feat_update(emb[d])
                t.age = 0
                t.hits += 1
        for k in u_trk:
            t = self.tracks[k]
            t.age += 1
            if not t.confirmed:
                del self.tracks[k]
                continue
            if t.age > self.max_age:
                self._mark_lost(k)
        for d in itertools.chain(invalid, reid_u):
            self._spawn(frame_id, det_tlbr[d].copy(), int(det_label[d]))

    def _rectify(self, matches, u_trk, det_tlbr):
        inactive = [m for m in matches if not self.tracks[m[0]].active]
        u_active = [k for k in u_trk if self.tracks[k].confirmed and self.tracks[k].active]
        if not inactive or not u_active:
            return matches, u_trk
        m_in, d_ids = zip(*inactive)
        c = assoc.iou_dist(np.array([self.tracks[k].tlbr for k in u_active]), det_tlbr[list(d_ids)])
        dup, _, _ = assoc.greedy_match(c, u_active, list(range(len(inactive))), 1. - self.duplicate_thresh)
        for uk, col in dup:
            mk, d = m_in[col], d_ids[col]
            tu, tm = self.tracks[uk], self.tracks[mk]
            if tm.end_frame < tu.start_frame:
                tm.end_frame = tu.end_frame
                tm.tlbr, tm.mean, tm.cov, tm.age = tu.tlbr, tu.mean, tu.cov, tu.age
                tm.hits += tu.hits
                tm.keypoints, tm.prev_keypoints = tu.keypoints, tu.prev_keypoints
                total = tm.f_count + tu.f_count
                if tm.f_sum is None:
                    tm.f_sum, tm.f_avg, tm.f_count = tu.f_sum, tu.f_avg, total
                elif tu.f_sum is not None:
                    tm.feat_update(tu.f_sum, merged_count=total)
                else:
                    tm.f_count = total
                u_trk.remove(uk)
                del self.tracks[uk]
            else:
                u_trk.remove(uk)
                u_trk.add(mk)
                matches.remove((mk, d))
                matches.add((uk, d))
        return matches, u_trk

    def visible(self):
        return [(k, t.tlbr.copy()) for k, t in self.tracks.items() if t.confirmed and t.active]
