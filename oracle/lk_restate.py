"""TEST INFRASTRUCTURE ONLY.  Pure-Python restatement of OpenCV's LKTrackerInvoker (lkpyramid.cpp, OpenCV 4.13)
for small point sets — documents exactly what csrc/klt_lk.cu implements and is itself pinned against
cv2.calcOpticalFlowPyrLK in tests/test_oracle_klt.py."""
import numpy as np


def _refl(p, n):
    if n == 1:
        return 0
    while p < 0 or p >= n:
        p = -p if p < 0 else 2 * n - 2 - p
    return p


def pyr_down(src):
    h, w = src.shape
    dh, dw = (h + 1) // 2, (w + 1) // 2
    s = src.astype(np.int64)
    g = [1, 4, 6, 4, 1]
    ys, xs = np.arange(dh) * 2, np.arange(dw) * 2
    rf = np.vectorize(_refl)
    acc = np.zeros((dh, dw), np.int64)
    for j in range(5):
        ry = rf(ys + j - 2, h)
        ra = np.zeros((dh, dw), np.int64)
        for i in range(5):
            ra += g[i] * s[ry][:, rf(xs + i - 2, w)]
        acc += g[j] * ra
    return ((acc + 128) >> 8).astype(np.uint8)


def scharr(src):
    h, w = src.shape
    s = src.astype(np.int64)
    rf = np.vectorize(_refl)
    ym, yp = rf(np.arange(h) - 1, h), rf(np.arange(h) + 1, h)
    xm, xp = rf(np.arange(w) - 1, w), rf(np.arange(w) + 1, w)
    sm = (s[ym] + s[yp]) * 3 + s * 10
    df = s[yp] - s[ym]
    return np.stack([sm[:, xp] - sm[:, xm], (df[:, xp] + df[:, xm]) * 3 + df * 10], -1).astype(np.int16)


def build_pyramid(img, win=(5, 5), max_level=5):
    levels = [img]
    for _ in range(max_level):
        nxt = pyr_down(levels[-1])
        if nxt.shape[1] <= win[0] or nxt.shape[0] <= win[1]:
            break
        levels.append(nxt)
    return levels, [scharr(l) for l in levels]


def lk_track(prev, cur, pts, win=(5, 5), max_level=5, max_count=10, eps=0.03, min_eig=1e-4):
    f32 = np.float32
    I_l, dI_l = build_pyramid(prev, win, max_level)
    J_l, _ = build_pyramid(cur, win, max_level)
    L = len(I_l)
    ww, wh = win
    half = (f32((ww - 1) * 0.5), f32((wh - 1) * 0.5))
    eps2 = f32(eps * eps)
    FS = f32(1.0 / (1 << 20))
    out = np.zeros((len(pts), 2), f32)
    status = np.ones(len(pts), bool)
    err = np.zeros(len(pts), f32)

    def px(img, x, y):
        return int(img[_refl(y, img.shape[0]), _refl(x, img.shape[1])])

    def dv(d, x, y):
        if x < 0 or y < 0 or x >= d.shape[1] or y >= d.shape[0]:
            return (0, 0)
        return (int(d[y, x, 0]), int(d[y, x, 1]))

    def weights(a, b):
        w00 = int(np.rint(f32(f32(f32(1) - a) * f32(f32(1) - b)) * f32(16384)))
        w01 = int(np.rint(f32(a * f32(f32(1) - b)) * f32(16384)))
        w10 = int(np.rint(f32(f32(f32(1) - a) * b) * f32(16384)))
        return w00, w01, w10, 16384 - w00 - w01 - w10

    for pi, p in enumerate(pts):
        nx = ny = f32(0)
        for level in range(L - 1, -1, -1):
            I, dI, J = I_l[level], dI_l[level], J_l[level]
            H, W = I.shape
            sc = f32(1. / (1 << level))
            ppx, ppy = f32(p[0] * sc), f32(p[1] * sc)
            if level == L - 1:
                nx, ny = ppx, ppy
            else:
                nx, ny = f32(nx * 2), f32(ny * 2)
            ppx, ppy = f32(ppx - half[0]), f32(ppy - half[1])
            ix, iy = int(np.floor(ppx)), int(np.floor(ppy))
            if ix < -ww or ix >= W or iy < -wh or iy >= H:
                if level == 0:
                    status[pi] = False
                    err[pi] = 0
                continue
            w00, w01, w10, w11 = weights(f32(ppx - ix), f32(ppy - iy))
            Iv = np.zeros((wh, ww), np.int64); Ix = np.zeros_like(Iv); Iy = np.zeros_like(Iv)
            A11 = A12 = A22 = f32(0)
            for y in range(wh):
                for x in range(ww):
                    xx, yy = ix + x, iy + y
                    Iv[y, x] = (px(I, xx, yy) * w00 + px(I, xx + 1, yy) * w01 + px(I, xx, yy + 1) * w10 +
                                px(I, xx + 1, yy + 1) * w11 + 256) >> 9
                    d00, d01, d10, d11 = dv(dI, xx, yy), dv(dI, xx + 1, yy), dv(dI, xx, yy + 1), dv(dI, xx + 1, yy + 1)
                    Ix[y, x] = (d00[0] * w00 + d01[0] * w01 + d10[0] * w10 + d11[0] * w11 + 8192) >> 14
                    Iy[y, x] = (d00[1] * w00 + d01[1] * w01 + d10[1] * w10 + d11[1] * w11 + 8192) >> 14
                    A11 = f32(A11 + f32(Ix[y, x] * Ix[y, x])); A12 = f32(A12 + f32(Ix[y, x] * Iy[y, x]))
                    A22 = f32(A22 + f32(Iy[y, x] * Iy[y, x]))
            A11, A12, A22 = f32(A11 * FS), f32(A12 * FS), f32(A22 * FS)
            D = f32(f32(A11 * A22) - f32(A12 * A12))
            me = f32(f32(A22 + A11 - np.sqrt(f32(f32((A11 - A22) * (A11 - A22)) + f32(f32(4) * A12 * A12)))) /
                     f32(2 * ww * wh))
            if me < min_eig or D < np.finfo(f32).eps:
                if level == 0:
                    status[pi] = False
                continue
            D = f32(1) / D
            cx, cy = f32(nx - half[0]), f32(ny - half[1])
            pdx = pdy = f32(0)
            for j in range(max_count):
                jx, jy = int(np.floor(cx)), int(np.floor(cy))
                if jx < -ww or jx >= W or jy < -wh or jy >= H:
                    if level == 0:
                        status[pi] = False
                    break
                v00, v01, v10, v11 = weights(f32(cx - jx), f32(cy - jy))
                b1 = b2 = f32(0)
                for y in range(wh):
                    for x in range(ww):
                        xx, yy = jx + x, jy + y
                        jv = (px(J, xx, yy) * v00 + px(J, xx + 1, yy) * v01 + px(J, xx, yy + 1) * v10 +
                              px(J, xx + 1, yy + 1) * v11 + 256) >> 9
                        diff = jv - Iv[y, x]
                        b1 = f32(b1 + f32(diff * Ix[y, x])); b2 = f32(b2 + f32(diff * Iy[y, x]))
                b1, b2 = f32(b1 * FS), f32(b2 * FS)
                dx = f32(f32(f32(A12 * b2) - f32(A22 * b1)) * D)
                dy = f32(f32(f32(A12 * b1) - f32(A11 * b2)) * D)
                cx, cy = f32(cx + dx), f32(cy + dy)
                nx, ny = f32(cx + half[0]), f32(cy + half[1])
                if f32(dx * dx + dy * dy) <= eps2:
                    break
                if j > 0 and abs(dx + pdx) < 0.01 and abs(dy + pdy) < 0.01:
                    nx, ny = f32(nx - dx * f32(0.5)), f32(ny - dy * f32(0.5))
                    break
                pdx, pdy = dx, dy
            if level == 0 and status[pi]:
                ex, ey = f32(nx - half[0]), f32(ny - half[1])
                jx, jy = int(np.floor(ex)), int(np.floor(ey))
                if jx < -ww or jx >= W or jy < -wh or jy >= H:
                    status[pi] = False
                else:
                    v00, v01, v10, v11 = weights(f32(ex - jx), f32(ey - jy))
                    ev = 0
                    for y in range(wh):
                        for x in range(ww):
                            xx, yy = jx + x, jy + y
                            jv = (px(J, xx, yy) * v00 + px(J, xx + 1, yy) * v01 + px(J, xx, yy + 1) * v10 +
                                  px(J, xx + 1, yy + 1) * v11 + 256) >> 9
                            ev += abs(jv - Iv[y, x])
                    err[pi] = f32(ev) / f32(32 * ww * wh)
        out[pi] = (nx, ny)
    return out, status, err
