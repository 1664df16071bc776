"""TEST INFRASTRUCTURE ONLY (checker, never shipped / never imported by fastmot_b200).

Numpy restatement of the reference Kalman filter, batched over tracks.
Follows fastmot/kalman_filter.py: create :96-126, warp :227-292, _init_mat :294-306,
_predict :308-319, _project :321-336, _update :338-345, _maha_distance :347-353.
Pinned against the imported reference in tests/test_oracle_vs_reference.py (container only) and
against tests/golden/assoc_*.npz everywhere.
"""
import numpy as np

FLOW, DETECTOR = 0, 1

DEFAULTS = dict(std_factor_acc=2.25, std_offset_acc=78.5, std_factor_det=(0.08, 0.08),
                std_factor_klt=(0.14, 0.14), min_std_det=(4.0, 4.0), min_std_klt=(5.0, 5.0),
                init_pos_weight=5, init_vel_weight=12, vel_coupling=0.6, vel_half_life=2)


class KalmanOracle:
    def __init__(self, dt=1 / 30., **kw):
        p = dict(DEFAULTS)
        p.update(kw)
        self.p = p
        self.reset_dt(dt)

    def reset_dt(self, dt):
        p = self.p
        q = np.zeros((8, 8))
        q[np.arange(4), np.arange(4)] = 0.25 * dt ** 4
        q[np.arange(4, 8), np.arange(4, 8)] = dt ** 2
        q[np.arange(4, 8), np.arange(4)] = 0.5 * dt ** 3
        q[np.arange(4), np.arange(4, 8)] = 0.5 * dt ** 3
        a = np.eye(8)
        for i in range(4):
            a[i, i + 4] = p['vel_coupling'] * dt
            a[i, (i + 2) % 4 + 4] = (1. - p['vel_coupling']) * dt
            a[i + 4, i + 4] = 0.5 ** (dt / p['vel_half_life'])
        self.acc_cov, self.trans_mat = q, a

    # ---- batched primitives: mean (n,8), cov (n,8,8) ------------------------------------------
    def create(self, tlbr):
        tlbr = np.asarray(tlbr, np.float64).reshape(-1, 4)
        p = self.p
        n = len(tlbr)
        mean = np.concatenate([tlbr, np.zeros((n, 4))], 1)
        wh = np.stack([tlbr[:, 2] - tlbr[:, 0] + 1, tlbr[:, 3] - tlbr[:, 1] + 1], 1)
        fac = np.asarray(p['std_factor_det'], float)
        mn = np.asarray(p['min_std_det'], float)
        pos = np.maximum(p['init_pos_weight'] * fac * wh, mn)
        vel = np.maximum(p['init_vel_weight'] * fac * wh, mn)
        std = np.concatenate([pos, pos, vel, vel], 1)
        cov = np.zeros((n, 8, 8))
        cov[:, np.arange(8), np.arange(8)] = std ** 2
        return mean, cov

    def warp(self, mean, cov, H):
        H1, h2, h3 = H[:2, :2], H[:2, 2], H[2, :2]
        n = len(mean)
        out = np.empty_like(mean)
        F = np.zeros((n, 8, 8))
        for c in range(2):
            ip, iv = slice(2 * c, 2 * c + 2), slice(4 + 2 * c, 6 + 2 * c)
            p, v = mean[:, ip], mean[:, iv]
            u = p @ H1.T + h2
            w = v @ H1.T
            a = p @ h3 + 1.0
            b = v @ h3
            out[:, ip] = u / a[:, None]
            out[:, iv] = w / a[:, None] - (b / a ** 2)[:, None] * u
            d = H1[None] / a[:, None, None] - u[:, :, None] * h3[None, None, :] / (a ** 2)[:, None, None]
            F[:, ip, ip] = d
            F[:, iv, iv] = d
            F[:, iv, ip] = (-(w[:, :, None] * h3[None, None, :] + b[:, None, None] * H1[None]) /
                            (a ** 2)[:, None, None] +
                            2 * b[:, None, None] * u[:, :, None] * h3[None, None, :] / (a ** 3)[:, None, None])
        cov = F @ cov @ F.transpose(0, 2, 1)
        return out, cov

    def predict(self, mean, cov):
        p = self.p
        size = np.maximum(mean[:, 2] - mean[:, 0] + 1, mean[:, 3] - mean[:, 1] + 1)
        std = p['std_factor_acc'] * size + p['std_offset_acc']
        A = self.trans_mat
        mean = mean @ A.T
        cov = A[None] @ cov @ A.T[None] + self.acc_cov[None] * (std ** 2)[:, None, None]
        cov = 0.5 * (cov + cov.transpose(0, 2, 1))
        return mean, cov

    def project(self, mean, cov, meas_type, multiplier=1.0):
        p = self.p
        fac = np.asarray(p['std_factor_klt'] if meas_type == FLOW else p['std_factor_det'], float)
        mn = np.asarray(p['min_std_klt'] if meas_type == FLOW else p['min_std_det'], float)
        wh = np.stack([mean[:, 2] - mean[:, 0] + 1, mean[:, 3] - mean[:, 1] + 1], 1)
        std = np.maximum(fac * wh, mn)
        std = np.concatenate([std, std], 1) * np.asarray(multiplier, float).reshape(-1, 1)
        S = cov[:, :4, :4].copy()
        S[:, np.arange(4), np.arange(4)] += std ** 2
        return mean[:, :4].copy(), S

    def update(self, mean, cov, meas, meas_type, multiplier=1.0):
        pm, S = self.project(mean, cov, meas_type, multiplier)
        PHt = cov[:, :, :4]
        K = np.linalg.solve(S, PHt.transpose(0, 2, 1)).transpose(0, 2, 1)
        y = np.asarray(meas, float).reshape(-1, 4) - pm
        mean = mean + np.einsum('nij,nj->ni', K, y)
        cov = cov - K @ S @ K.transpose(0, 2, 1)
        return mean, cov

    def motion_distance(self, mean, cov, meas):
        """(n_trk, n_det) squared Mahalanobis distances."""
        pm, S = self.project(mean, cov, DETECTOR)
        L = np.linalg.cholesky(S)
        diff = np.asarray(meas, float)[None, :, :] - pm[:, None, :]
        y = np.linalg.solve(L, diff.transpose(0, 2, 1))
        return np.sum(y ** 2, axis=1)
