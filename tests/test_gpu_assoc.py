"""GPU parity: association-block kernels (csrc/kalman.cu, csrc/assoc.cu) through the C-ABI against the
reference-generated goldens and the oracle.  Tolerances: fp64 state <= 1e-8 (north_star allows 1e-3),
assignment indices bit-exact."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def prim():
    return np.load(os.path.join(GOLDEN, "assoc_primitives.npz"))


def test_device_is_b200(lib):
    assert lib.fm_device_ok() == 1, lib.fm_last_error()


def test_kalman_batched_chain(prim, lib):
    from gpu_util import dev, host, kalman_params
    from fastmot_b200 import _lib
    from fastmot_b200.devmem import ptr, stream_ptr
    from fastmot_b200.kalman_filter import FM_KF_WARP, FM_KF_PREDICT, FM_KF_UPDATE, FM_KF_MEAS_DET
    kf = kalman_params()
    n = len(prim['kf_tlbr'])
    cap = n + 7
    mean = torch.zeros(cap, 8, dtype=torch.float64, device="cuda")
    cov = torch.zeros(cap, 64, dtype=torch.float64, device="cuda")
    tl = torch.zeros(cap, 4, dtype=torch.float64, device="cuda")
    perm = np.random.default_rng(0).permutation(cap)[:n].astype(np.int32)   # scattered slots
    slots = dev(perm)
    z0 = dev(prim['kf_tlbr'])
    kf.create_batched(mean, cov, tl, ptr(slots), ptr(z0), None, n)
    np.testing.assert_allclose(host(mean)[perm], prim['kf_m0'], atol=1e-12)
    np.testing.assert_allclose(host(cov)[perm].reshape(n, 8, 8), prim['kf_c0'], atol=1e-9)
    H = dev(prim['kf_H'].reshape(9))
    zf = dev(prim['kf_zflow'])
    mult = dev(prim['kf_mult'])
    out = torch.zeros(n, 4, dtype=torch.float64, device="cuda")
    lost = torch.zeros(n, dtype=torch.uint8, device="cuda")
    kf.step_batched(mean, cov, tl, ptr(slots), n, FM_KF_WARP | FM_KF_PREDICT | FM_KF_UPDATE, homography=ptr(H),
                    meas=ptr(zf), mult_num=ptr(mult), frame_size=(1920, 1080), out_tlbr=ptr(out), out_lost=ptr(lost))
    m1 = host(mean)[perm]
    np.testing.assert_allclose(m1, prim['kf_m1'], atol=1e-8)
    np.testing.assert_allclose(host(cov)[perm].reshape(n, 8, 8), prim['kf_c1'], rtol=1e-9, atol=1e-8)
    np.testing.assert_array_equal(host(out), np.rint(m1[:, :4]))
    np.testing.assert_array_equal(host(tl)[perm], np.rint(m1[:, :4]))
    from oracle.assoc import ios
    np.testing.assert_array_equal(host(lost).astype(bool), ios(np.rint(m1[:, :4]), [0, 0, 1919, 1079]) < 0.5)
    # Mahalanobis
    zd = dev(prim['kf_zdet'])
    md = torch.zeros(n, n, dtype=torch.float64, device="cuda")
    _lib.check(lib.fm_motion_distance(ptr(mean), ptr(cov), ptr(slots), n, ptr(zd), n, kf.params, ptr(md),
                                      stream_ptr()), "maha")
    np.testing.assert_allclose(host(md), prim['kf_maha'], rtol=1e-8, atol=1e-8)
    kf.step_batched(mean, cov, tl, ptr(slots), n, FM_KF_UPDATE | FM_KF_MEAS_DET, meas=ptr(zd),
                    frame_size=(1920, 1080), out_tlbr=ptr(out), out_lost=ptr(lost))
    np.testing.assert_allclose(host(mean)[perm], prim['kf_m2'], atol=1e-8)
    np.testing.assert_allclose(host(cov)[perm].reshape(n, 8, 8), prim['kf_c2'], rtol=1e-9, atol=1e-8)


def test_kalman_numpy_api_matches_oracle(prim):
    from fastmot_b200 import KalmanFilter, MeasType
    from oracle.kalman import KalmanOracle, FLOW
    kf, ko = KalmanFilter(), KalmanOracle(1 / 30)
    kf.reset_dt(1 / 30)
    m, c = kf.create(prim['kf_tlbr'][3])
    mo, co = ko.create(prim['kf_tlbr'][3:4])
    np.testing.assert_allclose(m, mo[0], atol=1e-12)
    m, c = kf.warp(m, c, prim['kf_H'])
    m, c = kf.predict(m, c)
    m, c = kf.update(m, c, prim['kf_zflow'][3], MeasType.FLOW, 1.7)
    mo, co = ko.warp(mo, co, prim['kf_H'])
    mo, co = ko.predict(mo, co)
    mo, co = ko.update(mo, co, prim['kf_zflow'][3:4], FLOW, 1.7)
    np.testing.assert_allclose(m, mo[0], atol=1e-8)
    np.testing.assert_allclose(c, co[0], rtol=1e-9, atol=1e-8)
    d = kf.motion_distance(m, c, prim['kf_zdet'][:9])
    np.testing.assert_allclose(d, ko.motion_distance(mo, co, prim['kf_zdet'][:9])[0], rtol=1e-8)


@pytest.mark.parametrize("metric", ["cosine", "euclidean"])
def test_matching_cost_fused(prim, lib, metric):
    from gpu_util import dev, host, kalman_params, Keep
    from fastmot_b200 import _lib
    from fastmot_b200.devmem import ptr, stream_ptr
    from oracle import assoc
    from oracle.kalman import KalmanOracle
    kf = kalman_params()
    ko = KalmanOracle(1 / 30)
    XA, XB = prim['cd_XA'], prim['cd_XB']
    nt, nd = len(XA), len(XB)
    mean, cov = prim['kf_m1'][:nt], prim['kf_c1'][:nt]
    rng = np.random.default_rng(1)
    det_tlbr = np.rint(prim['kf_tlbr'][rng.permutation(64)[:nd]] + rng.normal(0, 4, (nd, 4)))
    valid = (rng.uniform(size=nt) > 0.1)
    occ = rng.uniform(size=nd) < 0.15
    tl = rng.integers(0, 2, nt).astype(np.int64)
    dl = rng.integers(0, 2, nd).astype(np.int64)
    sel = rng.permutation(nd)[:37].astype(np.int32)
    # oracle
    c = assoc.cdist(XA, XB[sel], metric, (~valid)[:, None] | occ[sel][None, :], 0.9)
    md = ko.motion_distance(mean, cov, det_tlbr[sel])
    c = assoc.fuse_motion(c, md, 0.2)
    c = assoc.gate_cost(c, tl, dl[sel], 0.8)
    out = torch.zeros(nt, len(sel), dtype=torch.float64, device="cuda")
    slots = dev(np.arange(nt, dtype=np.int32))
    K = Keep()
    rc = lib.fm_matching_cost(K(XA), K(valid.astype(np.uint8)), K(mean),
                              K(cov.reshape(nt, 64)), ptr(slots), K(tl), nt, K(XB),
                              K(det_tlbr), K(dl), K(occ.astype(np.uint8)), K(sel),
                              len(sel), 512, 1 if metric == 'cosine' else 0, 0.9, 0.2, 0.8, kf.params, ptr(out),
                              stream_ptr())
    _lib.check(rc, "fm_matching_cost")
    got = host(out)
    assert np.array_equal(got >= 1e5, c >= 1e5)
    np.testing.assert_allclose(got, c, atol=1e-6)
    assert (c < 1e5).sum() > 5


def test_cdist_golden(prim, lib):
    """cdist alone (motion off, gate off) against the reference's own cdist output."""
    from gpu_util import dev, host, kalman_params, Keep
    from fastmot_b200 import _lib
    from fastmot_b200.devmem import ptr, stream_ptr
    kf = kalman_params()
    K = Keep()
    XA, XB, mask = prim['cd_XA'], prim['cd_XB'], prim['cd_mask']
    nt, nd = len(XA), len(XB)
    out = torch.zeros(nt, nd, dtype=torch.float64, device="cuda")
    z = torch.zeros(nt, 64, dtype=torch.float64, device="cuda")
    for metric, key in ((1, 'cd_cos'), (0, 'cd_euc')):
        rc = lib.fm_matching_cost(K(XA), None, ptr(z), ptr(z), K(np.arange(nt, dtype=np.int32)), None,
                                  nt, K(XB), ptr(z), None, None, None, nd, 512, metric, 0.9, -1.0, -1.0,
                                  kf.params, ptr(out), stream_ptr())
        _lib.check(rc, "fm_matching_cost")
        got = host(out)
        np.testing.assert_allclose(got[~mask], prim[key][~mask], atol=1e-6)


def test_iou_and_occlusion(prim, lib):
    from gpu_util import dev, host, Keep
    from fastmot_b200 import _lib
    from fastmot_b200.devmem import ptr, stream_ptr
    K = Keep()
    a, b = prim['iou_a'], prim['iou_b']
    out = torch.zeros(len(a), len(b), dtype=torch.float64, device="cuda")
    _lib.check(lib.fm_iou_cost(K(a), None, None, len(a), K(b), None, None, len(b), -1.0, ptr(out),
                               stream_ptr()), "iou")
    np.testing.assert_allclose(host(out), prim['iou_dist'], atol=1e-12)
    boxes = prim['occ_in']
    occ = torch.zeros(len(boxes), dtype=torch.uint8, device="cuda")
    _lib.check(lib.fm_find_occluded(K(boxes), len(boxes), float(prim['occ_thresh']), ptr(occ),
                                    stream_ptr()), "occ")
    np.testing.assert_array_equal(host(occ).astype(bool), prim['occ_out'])


def _check_lsa(cost, rid, cid, want_m, want_ur, want_uc):
    from gpu_util import run_lsa
    from fastmot_b200.tracker import MultiTracker
    c4r, st = run_lsa(cost)
    assert st == 0
    m, ur, uc = MultiTracker._split(None, c4r, cost.shape[0], cost.shape[1], rid, cid)
    assert np.array_equal(np.array(m, np.int64).reshape(-1, 2), want_m)
    assert ur == want_ur and uc == want_uc


def test_lsa_golden_bit_exact(prim):
    for k in range(int(prim['n_la'])):
        _check_lsa(prim[f'la_cost_{k}'], prim[f'la_rid_{k}'].tolist(), prim[f'la_cid_{k}'].tolist(),
                   prim[f'la_m_{k}'], prim[f'la_ur_{k}'].tolist(), prim[f'la_uc_{k}'].tolist())


@pytest.mark.parametrize("name", ["seq_T64.npz", "seq_T200.npz", "seq_T70_overlap.npz"])
def test_lsa_sequence_vectors(name):
    g = np.load(os.path.join(GOLDEN, name))
    for i in range(int(g['n_lsa'])):
        _check_lsa(g[f'lsa_cost_{i}'], g[f'lsa_rid_{i}'].tolist(), g[f'lsa_cid_{i}'].tolist(),
                   g[f'lsa_matches_{i}'], g[f'lsa_urow_{i}'].tolist(), g[f'lsa_ucol_{i}'].tolist())


def test_lsa_random_vs_oracle_including_ties_and_large():
    from gpu_util import run_lsa
    from oracle import assoc
    rng = np.random.default_rng(5)
    shapes = [(1, 1), (1, 9), (9, 1), (200, 177), (177, 200), (23, 23), (300, 310), (257, 64), (1500, 40)]
    for (nr, nc) in shapes:
        for mode in range(3):
            if mode == 0:
                C = rng.uniform(0, 1, (nr, nc))
            elif mode == 1:
                C = rng.integers(0, 3, (nr, nc)).astype(float)
            else:
                C = np.where(rng.uniform(size=(nr, nc)) < 0.6, 1e5, np.round(rng.uniform(0, 1, (nr, nc)), 2))
            if nr * nc > 100000 and mode == 1:
                continue
            rows, cols = assoc.lsa(C)
            want = np.full(nr, -1, np.int64)
            want[rows] = cols
            c4r, st = run_lsa(C)
            assert st == 0
            got = np.where(c4r <= -2, -2 - c4r, c4r)
            assert np.array_equal(got, want), (nr, nc, mode)
            dem = c4r <= -2
            assert np.array_equal(dem, (want >= 0) & (C[np.arange(nr), np.maximum(want, 0)] >= 1e5))


def test_lsa_degenerate_shapes():
    from gpu_util import run_lsa
    c4r, st = run_lsa(np.zeros((5, 0)))
    assert st == 0 and np.array_equal(c4r, -np.ones(5, np.int32))
    c4r, st = run_lsa(np.zeros((0, 4)))
    assert len(c4r) == 0
    c4r, st = run_lsa(np.full((3, 3), np.inf))
    assert st == 1


def test_greedy_golden(prim):
    from gpu_util import run_greedy
    from fastmot_b200.tracker import MultiTracker
    for k in range(int(prim['n_la'])):
        C = prim[f'la_cost_{k}']
        c4r, order = run_greedy(C, 0.5)
        m, ur, uc = MultiTracker._split_greedy(c4r, order, C.shape[0], C.shape[1], prim[f'la_rid_{k}'].tolist(),
                                               prim[f'la_cid_{k}'].tolist())
        assert np.array_equal(np.array(m, np.int64).reshape(-1, 2), prim[f'gr_m_{k}'])
        assert ur == prim[f'gr_ur_{k}'].tolist() and uc == prim[f'gr_uc_{k}'].tolist()


def test_feature_update_matches_reference_semantics(lib):
    from gpu_util import dev, host, Keep
    from fastmot_b200 import _lib
    from fastmot_b200.devmem import ptr, stream_ptr
    K = Keep()
    rng = np.random.default_rng(2)
    E = 512
    vec = rng.normal(size=(6, E)).astype(np.float32)
    vec /= np.linalg.norm(vec, axis=1, keepdims=True)
    cap = 5
    s = torch.zeros(cap, E, device="cuda")
    a = torch.zeros(cap, E, device="cuda")
    last = torch.zeros(cap, E, device="cuda")
    v = torch.zeros(cap, dtype=torch.uint8, device="cuda")
    # host model of track.py:100-126
    hs, ha, cnt = {}, {}, {}
    seq = [(2, 0), (4, 1), (2, 2), (2, 3), (4, 4), (0, 5)]
    for slot, vi in seq:
        cnt[slot] = cnt.get(slot, 0) + 1
        if cnt[slot] == 1:
            hs[slot] = vec[vi].copy()
            ha[slot] = vec[vi].copy()
        else:
            hs[slot] = hs[slot] + vec[vi]
            av = (hs[slot].astype(np.float64) * (1. / cnt[slot])).astype(np.float32)
            ha[slot] = (av.astype(np.float64) * (1. / np.linalg.norm(av))).astype(np.float32)
        _lib.check(lib.fm_feature_update(ptr(s), ptr(a), ptr(last), ptr(v), K(np.array([slot], np.int32)),
                                         K(vec), K(np.array([vi], np.int32)),
                                         K(np.array([cnt[slot]], np.int32)), 1, E, stream_ptr()), "feat")
    for slot in hs:
        np.testing.assert_allclose(host(s)[slot], hs[slot], atol=1e-6)
        np.testing.assert_allclose(host(a)[slot], ha[slot], atol=1e-6)
    assert host(v).tolist() == [1, 0, 1, 0, 1]


def _cascade_emulated(F, I, R, goff, active, conf, occ, conf_thresh, max_reid):
    """tracker.py:199-233 on given cost matrices with the oracle's SciPy / Numba replays (oracle/assoc.py)."""
    from oracle import assoc
    n_conf, n_det = F.shape[0], F.shape[1]
    u_det = list(range(n_det))
    m1, u1 = [], []
    for g in range(len(goff) - 1):
        rows = list(range(goff[g], goff[g + 1]))
        if not rows:
            continue
        if not u_det:
            u1 += rows
            continue
        m, ur, u_det = assoc.linear_assignment(F[np.ix_(rows, u_det)], rows, u_det)
        m1 += m
        u1 += ur
    act = [r for r in u1 if active[r]]
    u1 = [r for r in u1 if not active[r]]

    def stage(rows, u_det):
        if not rows or not u_det:
            return [], list(rows), list(u_det)
        return assoc.linear_assignment(I[np.ix_(rows, u_det)], rows, u_det)
    m2, u2, u_det = stage(act, u_det)
    m3, u3, u_det = stage(list(range(n_conf, I.shape[0])), u_det)
    u_det = [d for d in u_det if conf[d] >= conf_thresh]
    valid = [d for d in u_det if not occ[d]]
    invalid = [d for d in u_det if occ[d]]
    hist = list(range(R.shape[0]))
    if hist and valid:
        reid, _, reid_u = assoc.greedy_match(R[np.ix_(hist, valid)], hist, valid, max_reid)
    else:
        reid, reid_u = [], valid
    return m1, u1, m2, u2, m3, u3, reid, invalid, reid_u


@pytest.mark.parametrize("seed", range(12))
def test_assoc_cascade_vs_per_stage_emulation(seed):
    """fm_assoc_cascade (all stages in one launch, lists on the device) == the stage-by-stage cascade built from the
    oracle's SciPy LSA replay, Numba set order and greedy match, on random cost matrices with INF gates, exact ties
    (quantised costs), empty groups, occluded / low-confidence detections."""
    import ctypes as C
    from fastmot_b200 import _lib
    from fastmot_b200.devmem import ptr, stream_ptr
    from gpu_util import dev, host
    lib = _lib.require_device()
    rng = np.random.default_rng(100 + seed)
    n_det = int(rng.integers(1, 200)) if seed else 200
    sizes = [int(rng.integers(0, 90)) if rng.random() > 0.25 else 0 for _ in range(4)] if seed else [200, 0, 0, 0]
    while sum(sizes) > 256:
        sizes[int(np.argmax(sizes))] //= 2
    goff = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    n_conf = int(goff[-1])
    n_unconf = int(rng.integers(0, 40))
    n_hist = int(rng.integers(0, 50))

    def costs(nr, quant):
        c = rng.uniform(0, 1, (nr, n_det))
        if quant:
            c = np.round(c * quant) / quant          # exact ties
        c[rng.random((nr, n_det)) < 0.35] = 1e5       # gated pairs
        return np.ascontiguousarray(c)
    F = costs(n_conf, 16 if seed % 2 else 0)
    I = costs(n_conf + n_unconf, 8 if seed % 3 == 0 else 0)
    R = costs(n_hist, 0)
    active = (rng.random(n_conf) < 0.6).astype(np.uint8)
    conf = rng.uniform(0.3, 1.0, n_det)
    occ = (rng.random(n_det) < 0.2).astype(np.uint8)
    want = _cascade_emulated(F, I, R, goff, active, conf, occ, 0.5, 0.6)

    cap = max(n_conf + n_unconf, n_det, n_hist, 1)
    n_out = int(lib.fm_assoc_cascade_out_ints(cap))
    keep = [dev(a) for a in (goff, active if n_conf else np.zeros(1, np.uint8), F.reshape(-1) if F.size else np.zeros(1),
                             I.reshape(-1) if I.size else np.zeros(1), R.reshape(-1) if R.size else np.zeros(1), conf, occ)]
    sub = torch.zeros(256 * 256, dtype=torch.float64, device="cuda")
    out = torch.full((n_out,), -9, dtype=torch.int32, device="cuda")
    d = _lib.FmCascadeDesc()
    d.n_det, d.n_conf, d.n_groups, d.n_unconf, d.n_hist, d.cap = n_det, n_conf, 4, n_unconf, n_hist, cap
    d.goff, d.conf_active, d.feat_cost, d.iou_cost, d.reid_cost, d.det_conf, d.det_occluded = (ptr(t) for t in keep)
    d.sub, d.out = ptr(sub), ptr(out)
    d.conf_thresh, d.max_reid_cost = 0.5, 0.6
    _lib.check(lib.fm_assoc_cascade(C.byref(d), stream_ptr()), "fm_assoc_cascade")
    o = host(out)
    hdr = o[:16]
    assert hdr[0] == 0
    arr = [o[16 + k * cap: 16 + (k + 1) * cap] for k in range(14)]
    n = [int(v) for v in hdr[1:10]]

    def pairs(a, b, k):
        return list(zip(a[:k].tolist(), b[:k].tolist()))
    got = (pairs(arr[0], arr[1], n[0]), arr[6][:n[3]].tolist(), pairs(arr[2], arr[3], n[1]), arr[7][:n[4]].tolist(),
           pairs(arr[4], arr[5], n[2]), arr[8][:n[5]].tolist(), pairs(arr[9], arr[10], n[6]), arr[11][:n[7]].tolist(),
           arr[12][:n[8]].tolist())
    names = ("matches1", "u_trk1", "matches2", "u_trk2", "matches3", "u_trk3", "reid", "invalid", "reid_u")
    for nm, g_, w_ in zip(names, got, want):
        assert [tuple(x) if isinstance(x, (tuple, list)) else x for x in g_] == \
               [tuple(x) if isinstance(x, (tuple, list)) else x for x in w_], (seed, nm, g_, w_)
    assert np.array_equal(arr[13][:n_det], occ)
