// Pyramidal Lucas-Kanade sparse optical flow, 8 lanes per point (one lane per window row, warp shuffles for the
// window sums), all pyramid levels inside one launch.
//
// Restates cv2.calcOpticalFlowPyrLK as called at fastmot/flow.py:203-209 (winSize 5x5, maxLevel 5,
// criteria (COUNT|EPS, 10, 0.03), flags 0, minEigThreshold 1e-4): OpenCV lkpyramid.cpp LKTrackerInvoker —
// Q14 bilinear weights, int16 Scharr derivatives, patch values descaled to 5 fractional bits, window sums in
// fp32 IN OPENCV'S ROW-MAJOR ORDER (group_seq_sum), err = mean |I - J| over the window / 32.  The pure-Python
// restatement of the same formulas (oracle/lk_restate.py) is bit-identical to cv2 4.13 on this image.
#include "common.cuh"
#include "../../include/fastmot_b200.h"

namespace {

#define LK_W_BITS 14
#define LK_MAX_WIN 8

__device__ __forceinline__ int refl101(int p, int n) {
    if (n == 1) return 0;
    while (p < 0 || p >= n) p = p < 0 ? -p : 2 * n - 2 - p;
    return p;
}

__device__ __forceinline__ float group_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    return v;
}

// Sum of the per-lane row arrays in OpenCV's sequential row-major order (bit-exact fp32 accumulation):
// lane r continues the running sum handed over by lane r-1.
// WW / WH: compile-time window size (0 = use the runtime win_w / win_h); the reference always runs 5x5, and with
// constant bounds the per-element predicates and loop overhead disappear (the profile had them at ~19 % of the
// kernel's instructions).
template <int NQ, int WW, int WH>
__device__ __forceinline__ void group_seq_sum(float (*vals)[LK_MAX_WIN], int win_w_rt, int win_h_rt, int lane8,
                                              float* out) {
    const int win_w = WW ? WW : win_w_rt, win_h = WH ? WH : win_h_rt;
    const int lane = threadIdx.x & 31, gbase = lane & ~7;
    float acc[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) acc[q] = 0.f;
#pragma unroll
    for (int r = 0; r < (WH ? WH : LK_MAX_WIN); ++r) {
        if (r >= win_h) break;
        float in[NQ];
#pragma unroll
        for (int q = 0; q < NQ; ++q) in[q] = __shfl_sync(0xffffffffu, acc[q], gbase | (r > 0 ? r - 1 : 0));
        if (lane8 == r) {
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                float a = r > 0 ? in[q] : 0.f;
#pragma unroll
                for (int x = 0; x < LK_MAX_WIN; ++x)
                    if (x < win_w) a += vals[q][x];
                acc[q] = a;
            }
        }
    }
#pragma unroll
    for (int q = 0; q < NQ; ++q) out[q] = __shfl_sync(0xffffffffu, acc[q], gbase | (win_h - 1));
}

__device__ __forceinline__ int px(const unsigned char* img, int w, int h, int x, int y) {
    return img[(size_t)refl101(y, h) * w + refl101(x, w)];
}

__device__ __forceinline__ short2 dv(const short2* d, int w, int h, int x, int y) {
    if (x < 0 || y < 0 || x >= w || y >= h) return make_short2(0, 0);  // BORDER_CONSTANT 0 on derivatives
    return d[(size_t)y * w + x];
}

// Rows y and y+1, columns x0 .. x0+ww of an 8-bit level (BORDER_REFLECT_101 outside).  The bilinear taps of
// neighbouring window columns share pixels, so a lane fetches 2 x (ww+1) values once instead of 4 per column, and
// windows that lie inside the image (nearly all of them) skip the reflection arithmetic, which was > 50 % of the
// kernel's executed instructions.
template <int WW>
__device__ __forceinline__ void fetch_px_rows(const unsigned char* __restrict__ img, int W, int H, int x0, int y,
                                              int win_w_rt, int* r0, int* r1) {
    const int ww = WW ? WW : win_w_rt;
    if (x0 >= 0 && x0 + ww < W && y >= 0 && y + 1 < H) {
        const unsigned char* p = img + (size_t)y * W + x0;
#pragma unroll
        for (int x = 0; x <= (WW ? WW : LK_MAX_WIN); ++x) {
            if (x > ww) break;
            r0[x] = p[x];
            r1[x] = p[W + x];
        }
    } else {
        const unsigned char* pa = img + (size_t)refl101(y, H) * W;
        const unsigned char* pb = img + (size_t)refl101(y + 1, H) * W;
#pragma unroll
        for (int x = 0; x <= (WW ? WW : LK_MAX_WIN); ++x) {
            if (x > ww) break;
            const int xr = refl101(x0 + x, W);
            r0[x] = pa[xr];
            r1[x] = pb[xr];
        }
    }
}

// same for the int16 Scharr derivative level (BORDER_CONSTANT 0 outside)
template <int WW>
__device__ __forceinline__ void fetch_dv_rows(const short2* __restrict__ d, int W, int H, int x0, int y, int win_w_rt,
                                              short2* d0, short2* d1) {
    const int ww = WW ? WW : win_w_rt;
    if (x0 >= 0 && x0 + ww < W && y >= 0 && y + 1 < H) {
        const short2* p = d + (size_t)y * W + x0;
#pragma unroll
        for (int x = 0; x <= (WW ? WW : LK_MAX_WIN); ++x) {
            if (x > ww) break;
            d0[x] = p[x];
            d1[x] = p[W + x];
        }
    } else {
#pragma unroll
        for (int x = 0; x <= (WW ? WW : LK_MAX_WIN); ++x) {
            if (x > ww) break;
            d0[x] = dv(d, W, H, x0 + x, y);
            d1[x] = dv(d, W, H, x0 + x, y + 1);
        }
    }
}

template <int WW, int WH>
__global__ void __launch_bounds__(128) lk_kernel(FmPyramid prev, FmPyramid cur, const float* __restrict__ pts_full,
                                                  const int* __restrict__ meta, float pt_scale_x, float pt_scale_y,
                                                  int win_w_rt, int win_h_rt, int max_count, double eps_sq,
                                                  float min_eig_thr,
                                                  float max_error, float unscale_x, float unscale_y,
                                                  float* __restrict__ out_pts, unsigned char* __restrict__ out_status,
                                                  float* __restrict__ out_err) {
    const int win_w = WW ? WW : win_w_rt, win_h = WH ? WH : win_h_rt;
    const int n_pts = meta[1];
    const int lane8 = threadIdx.x & 7;
    const int groups_per_block = blockDim.x >> 3;
    const int n_levels = prev.n_levels;
    const float half_w = (win_w - 1) * 0.5f, half_h = (win_h - 1) * 0.5f;
    const bool row_on = lane8 < win_h;
    for (int base = blockIdx.x * groups_per_block; base < n_pts; base += gridDim.x * groups_per_block) {
        const int p = base + (threadIdx.x >> 3);
        const bool valid = p < n_pts;
        // _scale_pts (flow.py:327-331): full-res -> optical-flow resolution, f32
        const float ptx = valid ? pts_full[2 * p] * pt_scale_x : 0.f;
        const float pty = valid ? pts_full[2 * p + 1] * pt_scale_y : 0.f;
        float nx = 0.f, ny = 0.f;  // nextPts[ptidx]
        bool status = valid;
        float err = 0.f;
        for (int level = n_levels - 1; level >= 0; --level) {
            const int W = prev.w[level], H = prev.h[level];
            const unsigned char* I = prev.img[level];
            const unsigned char* J = cur.img[level];
            const short2* dI = (const short2*)prev.deriv[level];
            const float sc = (float)(1. / (1 << level));
            float ppx = ptx * sc, ppy = pty * sc;
            if (level == n_levels - 1) { nx = ppx; ny = ppy; } else { nx *= 2.f; ny *= 2.f; }
            ppx -= half_w; ppy -= half_h;
            int ix = (int)floorf(ppx), iy = (int)floorf(ppy);
            bool lvl_on = valid;
            if (ix < -win_w || ix >= W || iy < -win_h || iy >= H) {
                if (level == 0) { status = false; err = 0.f; }
                lvl_on = false;
            }
            float a = ppx - ix, b = ppy - iy;
            int iw00 = (int)rintf((1.f - a) * (1.f - b) * (1 << LK_W_BITS));
            int iw01 = (int)rintf(a * (1.f - b) * (1 << LK_W_BITS));
            int iw10 = (int)rintf((1.f - a) * b * (1 << LK_W_BITS));
            int iw11 = (1 << LK_W_BITS) - iw00 - iw01 - iw10;
            short Iv[LK_MAX_WIN], Ix[LK_MAX_WIN], Iy[LK_MAX_WIN];
            float pa[3][LK_MAX_WIN];
#pragma unroll
            for (int x = 0; x < LK_MAX_WIN; ++x) { pa[0][x] = 0.f; pa[1][x] = 0.f; pa[2][x] = 0.f; }
            if (lvl_on && row_on) {
                const int y = iy + lane8;
                int r0[LK_MAX_WIN + 1], r1[LK_MAX_WIN + 1];
                short2 e0[LK_MAX_WIN + 1], e1[LK_MAX_WIN + 1];
                fetch_px_rows<WW>(I, W, H, ix, y, win_w, r0, r1);
                fetch_dv_rows<WW>(dI, W, H, ix, y, win_w, e0, e1);
#pragma unroll
                for (int x = 0; x < (WW ? WW : LK_MAX_WIN); ++x) {
                    if (x >= win_w) break;
                    const int ival = (r0[x] * iw00 + r0[x + 1] * iw01 + r1[x] * iw10 + r1[x + 1] * iw11 +
                                      (1 << (LK_W_BITS - 5 - 1))) >> (LK_W_BITS - 5);
                    const short2 d00 = e0[x], d01 = e0[x + 1];
                    const short2 d10 = e1[x], d11 = e1[x + 1];
                    const int ixv = (d00.x * iw00 + d01.x * iw01 + d10.x * iw10 + d11.x * iw11 +
                                     (1 << (LK_W_BITS - 1))) >> LK_W_BITS;
                    const int iyv = (d00.y * iw00 + d01.y * iw01 + d10.y * iw10 + d11.y * iw11 +
                                     (1 << (LK_W_BITS - 1))) >> LK_W_BITS;
                    Iv[x] = (short)ival; Ix[x] = (short)ixv; Iy[x] = (short)iyv;
                    pa[0][x] = (float)(ixv * ixv); pa[1][x] = (float)(ixv * iyv); pa[2][x] = (float)(iyv * iyv);
                }
            }
            float asum[3];
            group_seq_sum<3, WW, WH>(pa, win_w, win_h, lane8, asum);
            float A11 = asum[0], A12 = asum[1], A22 = asum[2];
            const float FLT_SCALE = 1.f / (1 << 20);
            A11 *= FLT_SCALE; A12 *= FLT_SCALE; A22 *= FLT_SCALE;
            float D = A11 * A22 - A12 * A12;
            const float min_eig = (A22 + A11 - sqrtf((A11 - A22) * (A11 - A22) + 4.f * A12 * A12)) /
                                  (2 * win_w * win_h);
            if (lvl_on && (min_eig < min_eig_thr || D < 1.1920929e-07f)) {
                if (level == 0) status = false;
                lvl_on = false;
            }
            D = 1.f / D;
            float cx = nx - half_w, cy = ny - half_h;  // nextPt -= halfWin
            float pdx = 0.f, pdy = 0.f;
            bool iter_on = lvl_on;
            for (int j = 0; j < max_count; ++j) {
                if (!__any_sync(0xffffffffu, iter_on)) break;
                const int jx = (int)floorf(cx), jy = (int)floorf(cy);
                if (iter_on && (jx < -win_w || jx >= W || jy < -win_h || jy >= H)) {
                    if (level == 0) status = false;
                    iter_on = false;
                }
                const float a2 = cx - jx, b2 = cy - jy;
                const int jw00 = (int)rintf((1.f - a2) * (1.f - b2) * (1 << LK_W_BITS));
                const int jw01 = (int)rintf(a2 * (1.f - b2) * (1 << LK_W_BITS));
                const int jw10 = (int)rintf((1.f - a2) * b2 * (1 << LK_W_BITS));
                const int jw11 = (1 << LK_W_BITS) - jw00 - jw01 - jw10;
                float pb[2][LK_MAX_WIN];
#pragma unroll
                for (int x = 0; x < LK_MAX_WIN; ++x) { pb[0][x] = 0.f; pb[1][x] = 0.f; }
                if (iter_on && row_on) {
                    const int y = jy + lane8;
                    int r0[LK_MAX_WIN + 1], r1[LK_MAX_WIN + 1];
                    fetch_px_rows<WW>(J, W, H, jx, y, win_w, r0, r1);
#pragma unroll
                    for (int x = 0; x < (WW ? WW : LK_MAX_WIN); ++x) {
                        if (x >= win_w) break;
                        const int jval = (r0[x] * jw00 + r0[x + 1] * jw01 + r1[x] * jw10 + r1[x + 1] * jw11 +
                                          (1 << (LK_W_BITS - 5 - 1))) >> (LK_W_BITS - 5);
                        const int diff = jval - Iv[x];
                        pb[0][x] = (float)(diff * Ix[x]);
                        pb[1][x] = (float)(diff * Iy[x]);
                    }
                }
                float bsum[2];
                group_seq_sum<2, WW, WH>(pb, win_w, win_h, lane8, bsum);
                const float b1 = bsum[0] * FLT_SCALE, b2s = bsum[1] * FLT_SCALE;
                if (iter_on) {
                    const float dx = (A12 * b2s - A22 * b1) * D, dy = (A12 * b1 - A11 * b2s) * D;
                    cx += dx; cy += dy;
                    nx = cx + half_w; ny = cy + half_h;
                    if ((double)dx * (double)dx + (double)dy * (double)dy <= eps_sq) {
                        iter_on = false;
                    } else if (j > 0 && fabsf(dx + pdx) < 0.01f && fabsf(dy + pdy) < 0.01f) {
                        nx -= dx * 0.5f; ny -= dy * 0.5f;
                        iter_on = false;
                    }
                    pdx = dx; pdy = dy;
                }
            }
            if (level == 0) {
                // err = mean |J - I| over the window / 32 at the final position
                const float ex = nx - half_w, ey = ny - half_h;
                const int jx = (int)floorf(ex), jy = (int)floorf(ey);
                bool err_on = status;
                if (err_on && (jx < -win_w || jx >= W || jy < -win_h || jy >= H)) { status = false; err_on = false; }
                const float a2 = ex - jx, b2 = ey - jy;
                const int jw00 = (int)rintf((1.f - a2) * (1.f - b2) * (1 << LK_W_BITS));
                const int jw01 = (int)rintf(a2 * (1.f - b2) * (1 << LK_W_BITS));
                const int jw10 = (int)rintf((1.f - a2) * b2 * (1 << LK_W_BITS));
                const int jw11 = (1 << LK_W_BITS) - jw00 - jw01 - jw10;
                float ev = 0.f;
                if (err_on && row_on) {
                    const int y = jy + lane8;
                    int r0[LK_MAX_WIN + 1], r1[LK_MAX_WIN + 1];
                    fetch_px_rows<WW>(J, W, H, jx, y, win_w, r0, r1);
#pragma unroll
                    for (int x = 0; x < (WW ? WW : LK_MAX_WIN); ++x) {
                        if (x >= win_w) break;
                        const int jval = (r0[x] * jw00 + r0[x + 1] * jw01 + r1[x] * jw10 + r1[x + 1] * jw11 +
                                          (1 << (LK_W_BITS - 5 - 1))) >> (LK_W_BITS - 5);
                        ev += (float)abs(jval - Iv[x]);
                    }
                }
                ev = group_sum(ev);
                if (err_on) err = ev / (float)(32 * win_w * win_h);   // OpenCV divides (1-ulp difference from a reciprocal)
            }
        }
        if (valid && lane8 == 0) {
            // _get_status (flow.py:348-349) and _unscale_pts with mask (flow.py:335-344)
            const bool good = status && (err < max_error);
            out_status[p] = good ? 1 : 0;
            out_err[p] = err;
            out_pts[2 * p] = good ? nx * unscale_x : nx;
            out_pts[2 * p + 1] = good ? ny * unscale_y : ny;
        }
    }
}

}  // namespace

extern "C" int fm_lk_track(const FmPyramid* prev, const FmPyramid* cur, const float* pts_full, const int* meta,
                           float pt_scale_x, float pt_scale_y, int win_w, int win_h, int max_count, float epsilon,
                           float min_eig_thr, float max_error, float* out_pts, unsigned char* out_status,
                           float* out_err, void* stream) {
    FM_REQUIRE(prev && cur, "fm_lk_track: pyramids are NULL");
    FM_REQUIRE(win_w >= 1 && win_w <= LK_MAX_WIN && win_h >= 1 && win_h <= LK_MAX_WIN, "fm_lk_track: window > 8");
    FM_REQUIRE(prev->n_levels == cur->n_levels && prev->n_levels <= FM_MAX_PYR_LEVELS, "fm_lk_track: levels");
    // OpenCV clamps the criteria: maxCount in [0,100], epsilon in [0,10], then squares epsilon
    max_count = max_count < 0 ? 0 : (max_count > 100 ? 100 : max_count);
    double e = epsilon < 0 ? 0 : (epsilon > 10 ? 10 : epsilon);
    if (win_w == 5 && win_h == 5)       // the reference's (only) configuration: compile-time window
        lk_kernel<5, 5><<<FM_NUM_SMS * 8, 128, 0, (cudaStream_t)stream>>>(
            *prev, *cur, pts_full, meta, pt_scale_x, pt_scale_y, win_w, win_h, max_count, e * e, min_eig_thr, max_error,
            1.0f / pt_scale_x, 1.0f / pt_scale_y, out_pts, out_status, out_err);
    else
        lk_kernel<0, 0><<<FM_NUM_SMS * 8, 128, 0, (cudaStream_t)stream>>>(
            *prev, *cur, pts_full, meta, pt_scale_x, pt_scale_y, win_w, win_h, max_count, e * e, min_eig_thr, max_error,
            1.0f / pt_scale_x, 1.0f / pt_scale_y, out_pts, out_status, out_err);
    FM_CHECK_LAUNCH("fm_lk_track");
    return FM_OK;
}
