"""Slot-indexed device pools holding per-track state (Kalman mean/cov, box, features, keypoints)."""
import numpy as np
import torch


class TrackPool:
    def __init__(self, capacity=2048, feat_dim=512, max_kp=1024, device="cuda"):
        self.capacity = capacity
        self.feat_dim = feat_dim
        self.max_kp = max_kp
        dev = torch.device(device)
        f64, f32 = torch.float64, torch.float32
        self.mean = torch.zeros(capacity, 8, dtype=f64, device=dev)
        self.cov = torch.zeros(capacity, 64, dtype=f64, device=dev)
        self.tlbr = torch.zeros(capacity, 4, dtype=f64, device=dev)
        self.inlier_ratio = torch.ones(capacity, dtype=f64, device=dev)
        self.klt_tlbr = torch.zeros(capacity, 4, dtype=f64, device=dev)
        self.klt_ok = torch.zeros(capacity, dtype=torch.uint8, device=dev)
        self.feat_sum = torch.zeros(capacity, feat_dim, dtype=f32, device=dev)
        self.feat_avg = torch.zeros(capacity, feat_dim, dtype=f32, device=dev)
        self.feat_last = torch.zeros(capacity, feat_dim, dtype=f32, device=dev)
        self.feat_valid = torch.zeros(capacity, dtype=torch.uint8, device=dev)
        self.kp = torch.zeros(capacity, max_kp, 2, dtype=f32, device=dev)
        self.kp_prev = torch.zeros(capacity, max_kp, 2, dtype=f32, device=dev)
        self.kp_count = torch.zeros(capacity, dtype=torch.int32, device=dev)
        self._free = list(range(capacity - 1, -1, -1))

    def acquire(self):
        if not self._free:
            raise MemoryError("TrackPool exhausted (raise capacity)")
        slot = self._free.pop()
        return slot

    def release(self, slot):
        self._free.append(slot)

    def reset_slots(self, slots):
        """Fresh-track defaults for newly acquired slots (track.py:141-148)."""
        if len(slots) == 0:
            return
        idx = torch.as_tensor(np.asarray(slots, np.int64), device=self.mean.device)
        self.feat_valid[idx] = 0
        self.kp_count[idx] = 0
        self.inlier_ratio[idx] = 1.0
        self.klt_ok[idx] = 0

    # ---- lazy host views (synchronising; API compatibility / tests only) ----
    def fetch_state(self, slot):
        return (self.mean[slot].cpu().numpy().copy(), self.cov[slot].cpu().numpy().reshape(8, 8).copy())

    def fetch_feature(self, slot, avg=True):
        return (self.feat_avg if avg else self.feat_sum)[slot].cpu().numpy().copy()

    def fetch_last_feat(self, slot):
        return self.feat_last[slot].cpu().numpy().copy()

    def fetch_scalar(self, name, slot):
        return float(getattr(self, name)[slot].item())

    def fetch_keypoints(self, slot, prev=False):
        n = int(self.kp_count[slot].item())
        src = self.kp_prev if prev else self.kp
        return src[slot, :n].cpu().numpy().copy()

    def copy_slot(self, dst, src, fields):
        for f in fields:
            t = getattr(self, f)
            t[dst] = t[src]
