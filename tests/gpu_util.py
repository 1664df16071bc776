"""Helpers for -m gpu tests: call the C-ABI with torch tensors as device-memory containers."""
import ctypes as C

import numpy as np
import torch

from fastmot_b200 import _lib
from fastmot_b200.devmem import ptr, stream_ptr

DEV = "cuda"


def dev(a, dtype=None):
    a = np.ascontiguousarray(a if dtype is None else np.asarray(a, dtype))
    return torch.as_tensor(a).to(DEV)


def host(t):
    torch.cuda.synchronize()
    return t.cpu().numpy()


def kalman_params(dt=1 / 30., **kw):
    from fastmot_b200.kalman_filter import KalmanFilter
    kf = KalmanFilter(**kw)
    kf.reset_dt(dt)
    return kf


def run_lsa(cost):
    lib = _lib.load()
    cost = np.ascontiguousarray(cost, np.float64)
    nr, nc = cost.shape
    c = dev(cost.reshape(-1) if cost.size else np.zeros(1))
    out = torch.full((max(nr, 1),), -7, dtype=torch.int32, device=DEV)
    st = torch.zeros(1, dtype=torch.int32, device=DEV)
    ws = torch.empty(max(int(lib.fm_lsa_workspace_bytes(nr, nc)), 16), dtype=torch.uint8, device=DEV)
    _lib.check(lib.fm_lsa(ptr(c), nr, nc, ptr(out), ptr(st), ptr(ws), stream_ptr()), "fm_lsa")
    return host(out)[:nr], int(host(st)[0])


def run_greedy(cost, max_cost):
    lib = _lib.load()
    cost = np.ascontiguousarray(cost, np.float64)
    nr, nc = cost.shape
    c = dev(cost.reshape(-1) if cost.size else np.zeros(1))
    out = torch.full((max(nr, 1),), -7, dtype=torch.int32, device=DEV)
    order = torch.full((max(nr, 1),), -7, dtype=torch.int32, device=DEV)
    _lib.check(lib.fm_greedy_match(ptr(c), nr, nc, float(max_cost), ptr(out), ptr(order), stream_ptr()), "greedy")
    return host(out)[:nr], host(order)[:nr]


class Keep:
    """Uploads arrays and keeps the tensors alive until the object dies (a bare ptr(dev(x)) would free
    the tensor before the kernel runs)."""

    def __init__(self):
        self.held = []

    def __call__(self, a, dtype=None):
        t = dev(a, dtype)
        self.held.append(t)
        return ptr(t)
