"""CPU, build container only: the oracle restatements against the UNMODIFIED reference imported from
/root/reference (skipped where the reference tree is absent, e.g. on the GPU box)."""
import numpy as np
import pytest

from oracle.refshim import reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason="reference tree not present")


@pytest.mark.parametrize("scene_kw,n_frames", [
    (dict(n_objects=40, seed=9), 17),
    (dict(n_objects=30, seed=4, overlap=True), 12),
    (dict(n_objects=64, seed=1), 22),                     # config-1 size; scripted drop-outs on frames 10 / 15 -> aging,
                                                          # IoU stage and re-identification from the history
    (dict(n_objects=50, seed=7, bounce_radius=8), 31),    # direction reversals (benchmark scene motion model)
    (dict(n_objects=30, seed=12, dropout_frames=(10,), dropout_every=1), 22),       # a detector frame with NO detections
    (dict(n_objects=30, seed=13, dropout_frames=(5, 10, 15), dropout_every=2), 22),  # half the objects never confirm
])
def test_oracle_tracker_full_pipeline_identical(scene_kw, n_frames):
    """OracleTracker + OracleFlow (cv2) vs reference MultiTracker + Flow: identical ids and boxes per frame."""
    from fastmot_b200.synth import SyntheticScene
    from oracle.ref_run import run_reference_tracker
    from oracle.run import run_oracle_tracker
    scene = SyntheticScene(**scene_kw)
    ref, rtrk = run_reference_tracker(scene, n_frames)
    got, otrk = run_oracle_tracker(SyntheticScene(**scene_kw), n_frames)
    for t in range(n_frames):
        assert np.array_equal(ref[t]['ids'], got[t]['ids']), t
        assert np.array_equal(ref[t]['tlbr'], got[t]['tlbr']), t
    assert list(rtrk.tracks.keys()) == list(otrk.tracks.keys())
    np.testing.assert_allclose(rtrk.homography, otrk.homography, atol=1e-12)
    for k in rtrk.tracks:
        np.testing.assert_allclose(rtrk.tracks[k].state[0], otrk.tracks[k].mean, atol=1e-6)


def test_detect_oracle_against_reference_functions():
    from oracle.refshim import load_reference
    from oracle import detect
    fm = load_reference()
    rng = np.random.default_rng(3)
    for trial in range(6):
        n = int(rng.integers(5, 150))
        tlwh = np.concatenate([rng.uniform(0, 300, (n, 2)), rng.uniform(10, 120, (n, 2))], 1).astype(np.float32)
        sc = rng.uniform(0.3, 1, n).astype(np.float32)
        assert np.array_equal(fm.utils.rect.diou_nms(tlwh, sc, 0.5), detect.diou_nms(tlwh, sc, 0.5))
