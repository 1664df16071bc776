"""CPU: the oracle restatements reproduce the golden vectors generated from the unmodified reference
(oracle/make_goldens.py).  These pin the oracle everywhere, including on the GPU box."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import assoc
from oracle.kalman import KalmanOracle, FLOW, DETECTOR


@pytest.fixture(scope="module")
def prim():
    return np.load(os.path.join(GOLDEN, "assoc_primitives.npz"))


def test_kalman_chain(prim):
    kf = KalmanOracle(1 / 30)
    m, c = kf.create(prim['kf_tlbr'])
    np.testing.assert_allclose(m, prim['kf_m0'], atol=1e-9)
    np.testing.assert_allclose(c, prim['kf_c0'], atol=1e-9)
    m, c = kf.warp(m, c, prim['kf_H'])
    m, c = kf.predict(m, c)
    m, c = kf.update(m, c, prim['kf_zflow'], FLOW, prim['kf_mult'])
    np.testing.assert_allclose(m, prim['kf_m1'], atol=1e-8)
    np.testing.assert_allclose(c, prim['kf_c1'], rtol=1e-9, atol=1e-8)
    md = kf.motion_distance(m, c, prim['kf_zdet'])
    np.testing.assert_allclose(md, prim['kf_maha'], rtol=1e-8, atol=1e-8)
    m, c = kf.update(m, c, prim['kf_zdet'], DETECTOR)
    np.testing.assert_allclose(m, prim['kf_m2'], atol=1e-8)
    np.testing.assert_allclose(c, prim['kf_c2'], rtol=1e-9, atol=1e-8)


def test_distances(prim):
    np.testing.assert_allclose(assoc.cdist(prim['cd_XA'], prim['cd_XB'], 'cosine', prim['cd_mask'], 0.9),
                               prim['cd_cos'], atol=1e-6)
    np.testing.assert_allclose(assoc.cdist(prim['cd_XA'], prim['cd_XB'], 'euclidean', prim['cd_mask'], 0.9),
                               prim['cd_euc'], atol=1e-6)
    np.testing.assert_allclose(assoc.iou_dist(prim['iou_a'], prim['iou_b']), prim['iou_dist'], atol=1e-12)
    assert np.array_equal(assoc.find_occluded(prim['occ_in'], float(prim['occ_thresh'])), prim['occ_out'])


def test_assignment_bit_exact(prim):
    for k in range(int(prim['n_la'])):
        C = prim[f'la_cost_{k}']
        rid, cid = prim[f'la_rid_{k}'].tolist(), prim[f'la_cid_{k}'].tolist()
        m, ur, uc = assoc.linear_assignment(C, rid, cid)
        assert np.array_equal(np.array(m, np.int64).reshape(-1, 2), prim[f'la_m_{k}'])
        assert ur == prim[f'la_ur_{k}'].tolist()
        assert uc == prim[f'la_uc_{k}'].tolist()
        m, ur, uc = assoc.greedy_match(C, rid, cid, 0.5)
        assert np.array_equal(np.array(m, np.int64).reshape(-1, 2), prim[f'gr_m_{k}'])
        assert ur == prim[f'gr_ur_{k}'].tolist()
        assert uc == prim[f'gr_uc_{k}'].tolist()


@pytest.mark.parametrize("name", ["seq_T64.npz", "seq_T200.npz", "seq_T70_overlap.npz"])
def test_sequence_lsa_vectors(name):
    g = np.load(os.path.join(GOLDEN, name))
    for i in range(int(g['n_lsa'])):
        m, ur, uc = assoc.linear_assignment(g[f'lsa_cost_{i}'], g[f'lsa_rid_{i}'].tolist(),
                                            g[f'lsa_cid_{i}'].tolist())
        assert np.array_equal(np.array(m, np.int64).reshape(-1, 2), g[f'lsa_matches_{i}'])
        assert ur == g[f'lsa_urow_{i}'].tolist()
        assert uc == g[f'lsa_ucol_{i}'].tolist()


def test_numba_set_order_product_copy_matches_oracle():
    from fastmot_b200.utils.numba_compat import set_difference_order
    rng = np.random.default_rng(0)
    for _ in range(300):
        n = int(rng.integers(1, 700))
        k = int(rng.integers(0, n + 1))
        removed = rng.permutation(n)[:k]
        assert set_difference_order(n, removed) == assoc.numba_set_difference_order(n, removed)
