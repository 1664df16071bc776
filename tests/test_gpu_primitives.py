"""GPU: the tcgen05 / TMEM / TMA building blocks of the fused kernels against numpy (fm_probe_umma)."""
import numpy as np
import pytest
import torch

from fastmot_b200 import _lib
from fastmot_b200.devmem import ptr, stream_ptr
from fastmot_b200.packing import pack_b_sw128
from gpu_util import dev, host

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n1,n2,rows,row0", [(64, 64, 128, 0), (64, 96, 512, 128), (96, 96, 200, 96),
                                             (128, 256, 128, 0), (64, 128, 100, 0)])
def test_probe_tma_ss_ts(n1, n2, rows, row0):
    lib = _lib.require_device()
    rng = np.random.default_rng(n1 + n2)
    a = rng.normal(0, 1, (rows, 64)).astype(np.float16)
    b1 = rng.normal(0, 0.2, (n1, 64)).astype(np.float16)
    b2 = rng.normal(0, 0.2, (n2, n1)).astype(np.float16)
    a_d, b1_d, b2_d = dev(a), dev(pack_b_sw128(b1)), dev(pack_b_sw128(b2))
    o0 = torch.zeros(128, n1, dtype=torch.float32, device="cuda")
    o1 = torch.zeros(128, n2, dtype=torch.float32, device="cuda")
    _lib.check(lib.fm_probe_umma(ptr(a_d), rows, row0, ptr(b1_d), ptr(b2_d), n1, n2, ptr(o0), ptr(o1), stream_ptr()),
               "fm_probe_umma")
    at = np.zeros((128, 64), np.float32)
    m = max(0, min(128, rows - row0))
    at[:m] = a[row0:row0 + m].astype(np.float32)
    want0 = at @ b1.astype(np.float32).T
    got0 = host(o0)
    assert np.abs(got0 - want0).max() < 1e-3, ("ss", np.abs(got0 - want0).max())
    want1 = got0.astype(np.float16).astype(np.float32) @ b2.astype(np.float32).T
    got1 = host(o1)
    assert np.abs(got1 - want1).max() < 2e-3, ("ts", np.abs(got1 - want1).max())
