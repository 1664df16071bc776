"""CPU: detector-side oracle vs the reference-generated golden (filter + DIoU-NMS + rounding, exact) and vs
OpenCV for the ROI resize formula the kernel implements."""
import os

import numpy as np

from conftest import GOLDEN
from oracle import detect


def test_filter_dets_golden_exact():
    g = np.load(os.path.join(GOLDEN, "detect_filter.npz"))
    for k in range(int(g['n'])):
        t, l, c = detect.filter_dets(g[f'det_{k}'], g[f'size_{k}'], g[f'off_{k}'], g[f'lm_{k}'], 0.25, 0.5,
                                     800000, 1.2)
        assert np.array_equal(t, g[f'tlbr_{k}'])
        assert np.array_equal(l, g[f'label_{k}'])
        np.testing.assert_allclose(c, g[f'conf_{k}'], atol=1e-7)


def test_roi_fixedpoint_formula_matches_cv2_within_1lsb():
    from fastmot_b200.synth import SyntheticScene
    sc = SyntheticScene(40, seed=1)
    fr = sc.frame(0)
    tl = sc.detections(0)[0]
    tl = np.concatenate([tl, [[-5.5, 10.2, 40.7, 90.9], [1890, 1000, 1950, 1100], [100, 100, 400, 700]]])
    a = detect.roi_preprocess(fr, tl)
    b = detect.roi_preprocess_fixedpoint(fr, tl)
    lsb = np.abs(a - b) * 255 * 0.229
    assert lsb.max() <= 1.03   # (per-channel std differs by 2 %)
    assert (lsb < 1e-3).mean() > 0.99


def test_letterbox_geometry_matches_survey_numbers():
    roi, up, off = detect.letterbox_geometry((1920, 1080), (640, 640), True)
    assert roi == (0, 140, 640, 360) and tuple(up) == (1920, 1920) and tuple(off) == (0., 420.)
    from fastmot_b200.detector import letterbox_geometry
    roi2, up2, off2 = letterbox_geometry((1920, 1080), (640, 640), True)
    assert roi2 == roi and tuple(up2) == tuple(up) and tuple(off2) == tuple(off)
