"""Kalman filter front-end with the reference's constructor and method names
(fastmot/kalman_filter.py:14-24, 86-292) over the batched sm_100a kernels in csrc/kalman.cu.

`MultiTracker` uses the batched device entry points (`step_batched`, `create_batched`); the per-track
numpy methods (`create/predict/project/update/motion_distance/warp`) are kept for drop-in compatibility
and run the same kernels on a one-element batch.
"""
from enum import Enum

import numpy as np
import torch

from . import _lib
from .devmem import ptr, stream_ptr

FM_KF_WARP, FM_KF_PREDICT, FM_KF_UPDATE, FM_KF_MEAS_DET, FM_KF_MEAS_BY_SLOT = 1, 2, 4, 8, 16


class MeasType(Enum):
    FLOW = 0
    DETECTOR = 1


class KalmanFilter:
    def __init__(self,
                 std_factor_acc=2.25,
                 std_offset_acc=78.5,
                 std_factor_det=(0.08, 0.08),
                 std_factor_klt=(0.14, 0.14),
                 min_std_det=(4.0, 4.0),
                 min_std_klt=(5.0, 5.0),
                 init_pos_weight=5,
                 init_vel_weight=12,
                 vel_coupling=0.6,
                 vel_half_life=2):
        assert std_factor_acc >= 0
        assert std_factor_det[0] >= 0 and std_factor_det[1] >= 0
        assert std_factor_klt[0] >= 0 and std_factor_klt[1] >= 0
        assert min_std_det[0] >= 0 and min_std_det[1] >= 0
        assert min_std_klt[0] >= 0 and min_std_klt[1] >= 0
        assert init_pos_weight >= 0
        assert init_vel_weight >= 0
        assert 0 <= vel_coupling <= 1
        assert vel_half_life > 0
        self.std_factor_acc = std_factor_acc
        self.std_offset_acc = std_offset_acc
        self.std_factor_det = std_factor_det
        self.std_factor_klt = std_factor_klt
        self.min_std_det = min_std_det
        self.min_std_klt = min_std_klt
        self.init_pos_weight = init_pos_weight
        self.init_vel_weight = init_vel_weight
        self.vel_coupling = vel_coupling
        self.vel_half_life = vel_half_life
        self._lib = _lib.load()
        self.reset_dt(1 / 30.)

    # ------------------------------------------------------------------ matrices (kalman_filter.py:294-306)
    def _init_mat(self, dt):
        acc_cov = np.diag([0.25 * dt**4] * 4 + [dt**2] * 4)
        acc_cov[4:, :4] = np.eye(4) * (0.5 * dt**3)
        acc_cov[:4, 4:] = np.eye(4) * (0.5 * dt**3)
        meas_mat = np.eye(4, 8)
        trans_mat = np.eye(8)
        for i in range(4):
            trans_mat[i, i + 4] = self.vel_coupling * dt
            trans_mat[i, (i + 2) % 4 + 4] = (1. - self.vel_coupling) * dt
            trans_mat[i + 4, i + 4] = 0.5**(dt / self.vel_half_life)
        return acc_cov, meas_mat, trans_mat

    def reset_dt(self, dt):
        self.acc_cov, self.meas_mat, self.trans_mat = self._init_mat(dt)
        p = _lib.FmKalmanParams()
        p.trans_mat[:] = self.trans_mat.ravel().tolist()
        p.acc_cov[:] = self.acc_cov.ravel().tolist()
        p.std_factor_acc = self.std_factor_acc
        p.std_offset_acc = self.std_offset_acc
        p.std_factor_det[:] = list(map(float, self.std_factor_det))
        p.std_factor_klt[:] = list(map(float, self.std_factor_klt))
        p.min_std_det[:] = list(map(float, self.min_std_det))
        p.min_std_klt[:] = list(map(float, self.min_std_klt))
        p.init_pos_weight = self.init_pos_weight
        p.init_vel_weight = self.init_vel_weight
        self.params = p

    # ------------------------------------------------------------------ batched device entry points
    def step_batched(self, mean_pool, cov_pool, tlbr_pool, slots_ptr, n, flags, homography=None, h_ok=None,
                     meas=None, has_meas=None, mult_num=None, mult_den_pool=None, frame_size=(0, 0),
                     out_tlbr=None, out_lost=None, hold=None):
        """All pointer arguments are ctypes c_void_p (or None)."""
        rc = self._lib.fm_kalman_step_batched(ptr(mean_pool), ptr(cov_pool), ptr(tlbr_pool), slots_ptr, n, flags,
                                              homography, h_ok, hold, meas, has_meas, mult_num, mult_den_pool,
                                              self.params, float(frame_size[0]), float(frame_size[1]),
                                              out_tlbr, out_lost, stream_ptr())
        _lib.check(rc, "fm_kalman_step_batched")

    def create_batched(self, mean_pool, cov_pool, tlbr_pool, slots_ptr, tlbr_ptr, idx_ptr, n):
        rc = self._lib.fm_kalman_create_batched(ptr(mean_pool), ptr(cov_pool), ptr(tlbr_pool), slots_ptr,
                                                tlbr_ptr, idx_ptr, n, self.params, stream_ptr())
        _lib.check(rc, "fm_kalman_create_batched")

    # ------------------------------------------------------------------ per-track numpy API (drop-in)
    def _run1(self, mean, cov, flags, H=None, meas=None, mult=1.0):
        _lib.require_device()
        dev = torch.device("cuda")
        m = torch.as_tensor(np.asarray(mean, np.float64).reshape(1, 8)).to(dev)
        c = torch.as_tensor(np.asarray(cov, np.float64).reshape(1, 64)).to(dev)
        slots = torch.zeros(1, dtype=torch.int32, device=dev)
        Hd = None if H is None else torch.as_tensor(np.asarray(H, np.float64).reshape(9)).to(dev)
        z = None if meas is None else torch.as_tensor(np.asarray(meas, np.float64).reshape(1, 4)).to(dev)
        mu = torch.full((1,), float(mult), dtype=torch.float64, device=dev)
        self.step_batched(m, c, None, ptr(slots), 1, flags, ptr(Hd), None, ptr(z), None, ptr(mu), None)
        return m.cpu().numpy().reshape(8), c.cpu().numpy().reshape(8, 8)

    def create(self, det_meas):
        _lib.require_device()
        dev = torch.device("cuda")
        m = torch.zeros(1, 8, dtype=torch.float64, device=dev)
        c = torch.zeros(1, 64, dtype=torch.float64, device=dev)
        slots = torch.zeros(1, dtype=torch.int32, device=dev)
        z = torch.as_tensor(np.asarray(det_meas, np.float64).reshape(1, 4)).to(dev)
        self.create_batched(m, c, None, ptr(slots), ptr(z), None, 1)
        return m.cpu().numpy().reshape(8), c.cpu().numpy().reshape(8, 8)

    def predict(self, mean, covariance):
        return self._run1(mean, covariance, FM_KF_PREDICT)

    def warp(self, mean, covariance, H):
        return self._run1(mean, covariance, FM_KF_WARP, H=H)

    def update(self, mean, covariance, measurement, meas_type, multiplier=1.):
        flags = FM_KF_UPDATE | (FM_KF_MEAS_DET if meas_type == MeasType.DETECTOR else 0)
        return self._run1(mean, covariance, flags, meas=measurement, mult=multiplier)

    def project(self, mean, covariance, meas_type, multiplier=1.):
        """Host-side helper (cheap, 4x4): kalman_filter.py:149-178."""
        mean = np.asarray(mean, np.float64)
        covariance = np.asarray(covariance, np.float64)
        if meas_type == MeasType.FLOW:
            fac, mn = self.std_factor_klt, self.min_std_klt
        elif meas_type == MeasType.DETECTOR:
            fac, mn = self.std_factor_det, self.min_std_det
        else:
            raise ValueError('Invalid measurement type')
        w, h = mean[2] - mean[0] + 1, mean[3] - mean[1] + 1
        std = np.array([max(fac[0] * w, mn[0]), max(fac[1] * h, mn[1])] * 2) * multiplier
        return mean[:4].copy(), covariance[:4, :4] + np.diag(std**2)

    def motion_distance(self, mean, covariance, measurements):
        """Squared Mahalanobis distances (kalman_filter.py:206-225)."""
        _lib.require_device()
        dev = torch.device("cuda")
        z = np.asarray(measurements, np.float64).reshape(-1, 4)
        n = len(z)
        m = torch.as_tensor(np.asarray(mean, np.float64).reshape(1, 8)).to(dev)
        c = torch.as_tensor(np.asarray(covariance, np.float64).reshape(1, 64)).to(dev)
        zd = torch.as_tensor(z).to(dev)
        out = torch.empty(1, max(n, 1), dtype=torch.float64, device=dev)
        rc = self._lib.fm_motion_distance(ptr(m), ptr(c), None, 1, ptr(zd), n, self.params, ptr(out), stream_ptr())
        _lib.check(rc, "fm_motion_distance")
        return out.cpu().numpy().reshape(-1)[:n]
