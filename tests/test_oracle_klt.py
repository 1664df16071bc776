"""CPU: the pure-Python restatement of OpenCV's pyramidal LK (oracle/lk_restate.py, the arithmetic csrc/klt_lk.cu
implements) is pinned bit-for-bit to cv2.calcOpticalFlowPyrLK with the reference's parameters
(fastmot/flow.py:85-89, :203-209): pyramid levels, Scharr derivatives, tracked points, status and error."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")


def _textured(h, w, seed, shift=(0.0, 0.0)):
    rng = np.random.default_rng(seed)
    low = rng.integers(0, 256, (h // 8 + 3, w // 8 + 3)).astype(np.float32)
    big = cv2.resize(low, (w + 16, h + 16), interpolation=cv2.INTER_CUBIC)
    m = np.float32([[1, 0, 8 + shift[0]], [0, 1, 8 + shift[1]]])
    img = cv2.warpAffine(big, m, (w, h), flags=cv2.INTER_LINEAR | cv2.WARP_INVERSE_MAP)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def test_pyramid_and_scharr_bit_exact():
    from oracle import lk_restate
    img = _textured(135, 240, 1)
    levels, derivs = lk_restate.build_pyramid(img, (5, 5), 5)
    n, pyr = cv2.buildOpticalFlowPyramid(img, (5, 5), 5, withDerivatives=True, pyrBorder=cv2.BORDER_REFLECT_101,
                                         derivBorder=cv2.BORDER_CONSTANT, tryReuseInputImage=False)
    assert len(levels) == n + 1
    for l in range(n + 1):
        assert np.array_equal(levels[l], pyr[2 * l]), l
        assert np.array_equal(derivs[l], pyr[2 * l + 1].reshape(derivs[l].shape)), l


@pytest.mark.parametrize("seed,shift", [(3, (1.3, -0.7)), (4, (-2.2, 1.6))])
def test_lk_points_status_error_bit_exact(seed, shift):
    from oracle import lk_restate
    prev = _textured(135, 240, seed)
    cur = _textured(135, 240, seed, shift)
    rng = np.random.default_rng(seed)
    pts = np.stack([rng.uniform(-2, 242, 40), rng.uniform(-2, 137, 40)], 1).astype(np.float32)
    want, st, err = cv2.calcOpticalFlowPyrLK(prev, cur, pts.reshape(-1, 1, 2), None, winSize=(5, 5), maxLevel=5,
                                             criteria=(3, 10, 0.03))
    got, gst, gerr = lk_restate.lk_track(prev, cur, pts)
    st = st.reshape(-1).astype(bool)
    assert np.array_equal(gst, st)
    assert st.sum() > 20
    assert np.array_equal(got[st], want.reshape(-1, 2)[st])
    assert np.array_equal(gerr[st], err.reshape(-1)[st])
