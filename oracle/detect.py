"""TEST INFRASTRUCTURE ONLY (checker, never shipped / never imported by fastmot_b200).

Numpy restatement of the reference detector pre/post-processing:
  yolo_decode      fastmot/plugins/yolo_layer.cu:115-230 (CalDetection, CalDetection_NewCoords)
  filter_dets      fastmot/detector.py:322-365 (+ find_split_indices numba.py:55-64)
  diou_nms         fastmot/utils/rect.py:198-244
  letterbox        fastmot/detector.py:289-320 (CuPy zoom order=1 mode='opencv' grid_mode=True; CuPy is absent in
                   this image, so this one is a restatement of the documented semantics: "parity unpinned")
  roi_preprocess   fastmot/feature_extractor.py:84-98 + fastmot/utils/rect.py:92-97 (cv2.resize is the
                   reference's third-party arithmetic: OpenCV >= 3.3, here 4.13.0)
"""
import numpy as np


def _sig(x):
    return (1.0 / (1.0 + np.exp(-x.astype(np.float32)))).astype(np.float32)


def yolo_decode(head, anchors, scale_x_y, input_wh, num_classes, new_coords):
    """head: [(5+C)*A, H, W] float32 -> (A*H*W, 7) float32 [x, y, w, h, box_conf, class_id, class_prob]."""
    head = np.asarray(head, np.float32)
    A = len(anchors) // 2
    C = num_classes
    _, H, W = head.shape
    t = head.reshape(A, 5 + C, H, W)
    cls = t[:, 5:]
    class_id = np.argmax(cls, axis=1)                        # first max, like the strict '>' scan
    best = np.max(cls, axis=1)
    col = np.arange(W, dtype=np.float32)[None, None, :]
    row = np.arange(H, dtype=np.float32)[None, :, None]
    s = np.float32(scale_x_y)
    aw = np.asarray(anchors[0::2], np.float32)[:, None, None]
    ah = np.asarray(anchors[1::2], np.float32)[:, None, None]
    if new_coords:
        cls_prob, box_prob = best, t[:, 4]
        bx = (col + (s * t[:, 0] - (s - 1) * np.float32(0.5))) / np.float32(W)
        by = (row + (s * t[:, 1] - (s - 1) * np.float32(0.5))) / np.float32(H)
        bw = t[:, 2] * t[:, 2] * 4 * aw / np.float32(input_wh[0])
        bh = t[:, 3] * t[:, 3] * 4 * ah / np.float32(input_wh[1])
    else:
        cls_prob, box_prob = _sig(best), _sig(t[:, 4])
        bx = (col + (s * _sig(t[:, 0]) - (s - 1) * np.float32(0.5))) / np.float32(W)
        by = (row + (s * _sig(t[:, 1]) - (s - 1) * np.float32(0.5))) / np.float32(H)
        bw = np.exp(t[:, 2]) * aw / np.float32(input_wh[0])
        bh = np.exp(t[:, 3]) * ah / np.float32(input_wh[1])
    bx = bx - bw / 2
    by = by - bh / 2
    out = np.stack([bx, by, bw, bh, box_prob, class_id.astype(np.float32), cls_prob], -1).astype(np.float32)
    return out.reshape(-1, 7)


def diou_nms(tlwhs, scores, nms_thresh, beta=0.6):
    """Greedy DIoU-NMS; f32-valued inputs, f64 arithmetic where Numba promotes (see csrc/detect.cu)."""
    tlwhs = np.asarray(tlwhs, np.float32)
    areas = (tlwhs[:, 2] * tlwhs[:, 3]).astype(np.float32)
    order = np.argsort(-np.asarray(scores, np.float32), kind='stable')
    tls = tlwhs[:, :2]
    brs = (tlwhs[:, :2] + tlwhs[:, 2:]).astype(np.float64) - 1
    centers = (tls + brs) / 2
    keep = []
    while order.size > 0:
        i = order[0]
        keep.append(i)
        o = order[1:]
        ix0 = np.maximum(tls[i, 0], tls[o, 0]); iy0 = np.maximum(tls[i, 1], tls[o, 1])
        ix1 = np.minimum(brs[i, 0], brs[o, 0]); iy1 = np.minimum(brs[i, 1], brs[o, 1])
        iw = np.maximum(0, ix1 - ix0 + 1); ih = np.maximum(0, iy1 - iy0 + 1)
        inter = iw * ih
        union = (areas[i] + areas[o]).astype(np.float32) - inter
        with np.errstate(divide='ignore', invalid='ignore'):
            iou = inter / union
        ex0 = np.minimum(tls[i, 0], tls[o, 0]); ey0 = np.minimum(tls[i, 1], tls[o, 1])
        ex1 = np.maximum(brs[i, 0], brs[o, 0]); ey1 = np.maximum(brs[i, 1], brs[o, 1])
        c = (ex1 - ex0 + 1) ** 2 + (ey1 - ey0 + 1) ** 2
        d = np.sum((centers[i] - centers[o]) ** 2, axis=1)
        with np.errstate(divide='ignore', invalid='ignore'):
            diou = iou - (d / c) ** beta
        order = o[diou <= nms_thresh]
    return np.array(keep, np.int64)


def filter_dets(det_out, size, offset, label_mask, conf_thresh, nms_thresh, max_area, min_ar):
    """Returns (tlbr f64 (D,4), label i64 (D,), conf f64 (D,)) — class asc, objectness desc."""
    det_out = np.asarray(det_out, np.float32).reshape(-1, 7)
    cls = det_out[:, 5].astype(np.int64)
    score = (det_out[:, 4] * det_out[:, 6]).astype(np.float32)
    keep = label_mask[cls] & (score.astype(np.float64) >= conf_thresh)
    d = det_out[keep].copy()
    if len(d) == 0:
        return np.zeros((0, 4)), np.zeros(0, np.int64), np.zeros(0)
    sz = np.asarray(size, np.float64)
    d[:, :4] = (d[:, :4].astype(np.float64) * np.concatenate([sz, sz])).astype(np.float32)
    d[:, :2] = (d[:, :2].astype(np.float64) - np.asarray(offset, np.float64)).astype(np.float32)
    tl, lb, cf = [], [], []
    for c in np.unique(d[:, 5]):
        cd = d[d[:, 5] == c]
        for i in diou_nms(cd[:, :4], cd[:, 4], nms_thresh):
            # to_tlbr (rect.py:48-57) under Numba: float(f32) stays f32, so x + w is an f32 add; the `- 1.`
            # literal promotes to f64 (pinned empirically: half-way cases round like this, not like f64 sums)
            x, y, w, h = cd[i, :4]
            box = np.rint([float(x), float(y), float(np.float32(x + w)) - 1., float(np.float32(y + h)) - 1.])
            bw, bh = box[2] - box[0] + 1, box[3] - box[1] + 1
            area = 0. if (bw <= 0 or bh <= 0) else bw * bh
            ar = bh / bw if bw > 0 else 0.
            if 0 < area <= max_area and ar >= min_ar:
                tl.append(box); lb.append(int(c)); cf.append(float(np.float32(cd[i, 4] * cd[i, 6])))
    return (np.array(tl, np.float64).reshape(-1, 4), np.array(lb, np.int64), np.array(cf, np.float64))


def letterbox_geometry(src_wh, dst_wh, letterbox):
    """detector.py:302-320 -> (roi_x, roi_y, roi_w, roi_h), upscaled_sz, bbox_offset."""
    src = np.array(src_wh)
    dst = np.array(dst_wh)
    if letterbox:
        scale = min(dst / src)
        scaled = np.rint(src * scale).astype(int)
        off = (dst - scaled) / 2
        roi = (int(off[0]), int(off[1]), int(scaled[0]), int(scaled[1]))
        upscaled = np.rint(dst / scale).astype(int)
        bbox_offset = (upscaled - src) / 2
    else:
        roi = (0, 0, int(dst[0]), int(dst[1]))
        upscaled = src
        bbox_offset = np.zeros(2)
    return roi, upscaled, bbox_offset


def letterbox(frame, dst_wh, roi):
    """-> float32 CHW RGB in [0,1] with 0.5 padding."""
    H, W = frame.shape[:2]
    rx, ry, rw, rh = roi
    out = np.full((3, dst_wh[1], dst_wh[0]), 0.5, np.float32)
    sx = np.clip((np.arange(rw) + 0.5) * (W / rw) - 0.5, 0, W - 1)
    sy = np.clip((np.arange(rh) + 0.5) * (H / rh) - 0.5, 0, H - 1)
    x0 = np.floor(sx).astype(int); y0 = np.floor(sy).astype(int)
    x1 = np.minimum(x0 + 1, W - 1); y1 = np.minimum(y0 + 1, H - 1)
    fx = (sx - x0)[None, :, None]; fy = (sy - y0)[:, None, None]
    f = frame.astype(np.float64)
    top = f[y0][:, x0] * (1 - fx) + f[y0][:, x1] * fx
    bot = f[y1][:, x0] * (1 - fx) + f[y1][:, x1] * fx
    small = np.rint(top * (1 - fy) + bot * fy)                  # uint8 in the reference
    chw = small[..., ::-1].transpose(2, 0, 1)
    out[:, ry:ry + rh, rx:rx + rw] = (chw * (1 / 255.)).astype(np.float32)
    return out


def roi_preprocess(frame, tlbrs, out_wh=(128, 256)):
    """-> float32 (N,3,H,W): crop (int truncation, clamp>=0), cv2.resize INTER_LINEAR, ImageNet normalise."""
    import cv2
    t = np.maximum(np.asarray(tlbrs).astype(np.int_), 0)
    out = np.empty((len(t), 3, out_wh[1], out_wh[0]), np.float32)
    mean = np.array([0.485, 0.456, 0.406]); std = np.array([0.229, 0.224, 0.225])
    for i in range(len(t)):
        img = frame[t[i, 1]:t[i, 3] + 1, t[i, 0]:t[i, 2] + 1]
        img = cv2.resize(img, out_wh)
        chw = img[..., ::-1].transpose(2, 0, 1)
        out[i] = ((chw / 255. - mean[:, None, None]) / std[:, None, None]).astype(np.float32)
    return out


def roi_preprocess_fixedpoint(frame, tlbrs, out_wh=(128, 256)):
    """Same as roi_preprocess but with OpenCV's 8-bit INTER_LINEAR fixed-point formula restated in numpy
    (what csrc/preproc.cu implements); used to show the kernel's formula == cv2 on this build."""
    t = np.maximum(np.asarray(tlbrs).astype(np.int_), 0)
    H, W = frame.shape[:2]
    out = np.empty((len(t), 3, out_wh[1], out_wh[0]), np.float32)
    mean = np.array([0.485, 0.456, 0.406]); std = np.array([0.229, 0.224, 0.225])

    def coef(n_out, n_in):
        f = ((np.arange(n_out) + 0.5) * (n_in / n_out) - 0.5).astype(np.float32)
        s = np.floor(f).astype(int)
        f = f - s
        lo = s < 0
        f[lo] = 0; s[lo] = 0
        hi = s >= n_in - 1
        f[hi] = 0; s[hi] = n_in - 1
        a0 = np.rint((np.float32(1) - f) * np.float32(2048)).astype(np.int64)
        a1 = np.rint(f * np.float32(2048)).astype(np.int64)
        return s, np.minimum(s + 1, n_in - 1), a0, a1

    for i in range(len(t)):
        img = frame[t[i, 1]:min(t[i, 3], H - 1) + 1, t[i, 0]:min(t[i, 2], W - 1) + 1].astype(np.int64)
        ch, cw = img.shape[:2]
        sx, sx1, a0, a1 = coef(out_wh[0], cw)
        sy, sy1, b0, b1 = coef(out_wh[1], ch)
        hrow = img[:, sx] * a0[None, :, None] + img[:, sx1] * a1[None, :, None]
        px = (((b0[:, None, None] * (hrow[sy] >> 4)) >> 16) + ((b1[:, None, None] * (hrow[sy1] >> 4)) >> 16) + 2) >> 2
        px = np.clip(px, 0, 255)
        chw = px[..., ::-1].transpose(2, 0, 1)
        out[i] = ((chw / 255. - mean[:, None, None]) / std[:, None, None]).astype(np.float32)
    return out
