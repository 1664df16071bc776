"""Deterministic synthetic 1080p scene used by tests, smoke() and bench.py.

Pure numpy (no OpenCV): band-limited background panned 1 px/frame, T textured rectangles on a
grid with small random velocities, exact inclusive-pixel detections with scripted drop-outs and
fixed random unit embeddings per identity.  The layout follows SURVEY.md §8(d); the reference has
no data generator of its own (it reads video files, fastmot/videoio.py).
"""
import numpy as np


def _cubic_kernel(t, a=-0.5):
    t = np.abs(t)
    out = np.zeros_like(t)
    m1 = t <= 1
    m2 = (t > 1) & (t < 2)
    out[m1] = (a + 2) * t[m1] ** 3 - (a + 3) * t[m1] ** 2 + 1
    out[m2] = a * t[m2] ** 3 - 5 * a * t[m2] ** 2 + 8 * a * t[m2] - 4 * a
    return out


def _upsample_axis(img, out_len, axis):
    """Separable Catmull-Rom upsampling along `axis` (half-pixel centres, edge clamp)."""
    in_len = img.shape[axis]
    scale = in_len / out_len
    x = (np.arange(out_len) + 0.5) * scale - 0.5
    x0 = np.floor(x).astype(np.int64)
    acc = 0
    for k in range(-1, 3):
        idx = np.clip(x0 + k, 0, in_len - 1)
        w = _cubic_kernel(x - (x0 + k))
        shape = [1] * img.ndim
        shape[axis] = out_len
        acc = acc + np.take(img, idx, axis=axis) * w.reshape(shape)
    return acc


def smooth_texture(rng, h, w, factor=10, channels=3):
    """Band-limited u8 texture: low-res uniform noise upsampled by `factor` (cubic)."""
    lh, lw = max(h // factor, 2) + 1, max(w // factor, 2) + 1
    low = rng.integers(0, 256, size=(lh, lw, channels)).astype(np.float64)
    up = _upsample_axis(_upsample_axis(low, h, 0), w, 1)
    return np.clip(np.rint(up), 0, 255).astype(np.uint8)


class SyntheticScene:
    """1920x1080 stream with `n_objects` moving textured rectangles.

    Attributes after `frame(t)`: none cached; everything is a pure function of (seed, t).
    """

    def __init__(self, n_objects=200, size=(1920, 1080), seed=0, label=1, emb_dim=512,
                 pan=1, overlap=False, dropout_frames=(10, 15), dropout_every=9, bounce_radius=None):
        self.size = size
        # bounce_radius (px): fold each object's linear displacement back into [-r, r] (triangle wave), so that a
        # long stream keeps every object inside its grid cell -- constant track count for benchmarks of any length.
        # None (the golden sequences) = unbounded linear motion.
        self.bounce_radius = bounce_radius
        self.n = n_objects
        self.label = label
        self.pan = pan
        self.dropout_frames = tuple(dropout_frames)
        self.dropout_every = dropout_every
        rng = np.random.default_rng(seed)
        W, H = size
        self.bg = smooth_texture(rng, H, W, 10)
        cols, rows = 20, 10
        assert n_objects <= cols * rows
        pitch_x, pitch_y = 92, 104
        if overlap:
            pitch_x, pitch_y = 60, 104
        idx = np.arange(n_objects)
        if n_objects < cols * rows:
            # spread the objects over the grid deterministically
            idx = np.sort(rng.choice(cols * rows, n_objects, replace=False))
        self.w = rng.integers(36, 56, n_objects)
        self.h = rng.integers(70, 96, n_objects)
        self.x0 = 40.0 + (idx % cols) * pitch_x
        self.y0 = 20.0 + (idx // cols) * pitch_y
        self.vel = rng.normal(0.0, 0.5, size=(n_objects, 2))
        self.tex = [smooth_texture(rng, int(self.h[i]), int(self.w[i]), 6) for i in range(n_objects)]
        emb = rng.normal(size=(n_objects, emb_dim))
        emb /= np.linalg.norm(emb, axis=1, keepdims=True)
        self.emb = emb.astype(np.float32)

    def positions(self, t):
        dx, dy = self.vel[:, 0] * t, self.vel[:, 1] * t
        if self.bounce_radius is not None:
            r = float(self.bounce_radius)
            dx = r - np.abs((dx + r) % (4 * r) - 2 * r)      # triangle wave: identity on [-r, r], period 4r
            dy = r - np.abs((dy + r) % (4 * r) - 2 * r)
        x = np.rint(self.x0 + dx).astype(np.int64)
        y = np.rint(self.y0 + dy).astype(np.int64)
        return x, y

    def frame(self, t):
        """BGR u8 HxWx3 frame number t."""
        img = np.roll(self.bg, self.pan * t, axis=1).copy()
        W, H = self.size
        x, y = self.positions(t)
        for i in range(self.n):
            xa, ya = int(x[i]), int(y[i])
            xb, yb = xa + int(self.w[i]), ya + int(self.h[i])
            cxa, cya, cxb, cyb = max(xa, 0), max(ya, 0), min(xb, W), min(yb, H)
            if cxb <= cxa or cyb <= cya:
                continue
            img[cya:cyb, cxa:cxb] = self.tex[i][cya - ya:cyb - ya, cxa - xa:cxb - xa]
        return img

    def detections(self, t, conf=0.9):
        """(tlbr f64 (D,4), labels i64 (D,), conf f64 (D,), identity index (D,))."""
        x, y = self.positions(t)
        keep = np.ones(self.n, bool)
        if t in self.dropout_frames:
            keep[::self.dropout_every] = False
        W, H = self.size
        tlbr = np.stack([x, y, x + self.w - 1, y + self.h - 1], 1).astype(np.float64)
        inside = (tlbr[:, 2] >= 0) & (tlbr[:, 3] >= 0) & (tlbr[:, 0] < W) & (tlbr[:, 1] < H)
        keep &= inside
        ids = np.nonzero(keep)[0]
        return (tlbr[ids], np.full(len(ids), self.label, np.int64),
                np.full(len(ids), conf, np.float64), ids)

    def embeddings(self, ids, t=0, noise=0.0):
        e = self.emb[ids].copy()
        if noise > 0:
            rng = np.random.default_rng(100003 * (t + 1))
            e = e + rng.normal(0, noise, e.shape).astype(np.float32)
            e /= np.linalg.norm(e, axis=1, keepdims=True)
        return e.astype(np.float32)
