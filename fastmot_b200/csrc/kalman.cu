// Batched 8-state Kalman filter for bounding boxes (fp64), one 64-thread CTA per track.
//
// Replaces the per-track Python loop of the reference:
//   fastmot/tracker.py:164-183 (apply_kalman)  -> warp + predict + update(FLOW) + round + out-of-frame test
//   fastmot/tracker.py:262-274 (update, matched) -> update(DETECTOR) + round + out-of-frame test
//   fastmot/kalman_filter.py:96-126 (create), :227-292 (warp), :308-345 (_predict/_project/_update)
// Arithmetic is fp64 like the reference; association order inside 8x8 products differs (<=1e-12 rel).
#include "common.cuh"
#include "../../include/fastmot_b200.h"

namespace {

__device__ __forceinline__ void mat8_mul(const double* A, const double* B, double* C, int t, bool transB) {
    // C = A * B (or A * B^T); 64 threads, thread t owns C[r][c]
    int r = t >> 3, c = t & 7;
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) acc += A[r * 8 + k] * (transB ? B[c * 8 + k] : B[k * 8 + c]);
    C[t] = acc;
}

// 4x4 inverse by Gauss-Jordan with partial pivoting (S is SPD in practice).
__device__ void inv4(const double* S, double* Si) {
    double a[4][8];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            a[i][j] = S[i * 4 + j];
            a[i][j + 4] = (i == j) ? 1.0 : 0.0;
        }
    for (int col = 0; col < 4; ++col) {
        int piv = col;
        double best = fabs(a[col][col]);
        for (int r = col + 1; r < 4; ++r)
            if (fabs(a[r][col]) > best) { best = fabs(a[r][col]); piv = r; }
        if (piv != col)
            for (int j = 0; j < 8; ++j) { double tmp = a[col][j]; a[col][j] = a[piv][j]; a[piv][j] = tmp; }
        double d = 1.0 / a[col][col];
        for (int j = 0; j < 8; ++j) a[col][j] *= d;
        for (int r = 0; r < 4; ++r) {
            if (r == col) continue;
            double f = a[r][col];
            for (int j = 0; j < 8; ++j) a[r][j] -= f * a[col][j];
        }
    }
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) Si[i * 4 + j] = a[i][j + 4];
}

// Jacobian + transformed mean of one corner (kalman_filter.py:247-281, written in block form).
// p: corner index 0 (tl) or 1 (br).
__device__ void warp_corner(const double* Hm, const double* x, int p, double* xo, double* F) {
    const int ip = 2 * p, iv = 4 + 2 * p;
    const double h11 = Hm[0], h12 = Hm[1], h21 = Hm[3], h22 = Hm[4];
    const double t1 = Hm[2], t2 = Hm[5], g1 = Hm[6], g2 = Hm[7];
    const double px = x[ip], py = x[ip + 1], vx = x[iv], vy = x[iv + 1];
    const double u1 = h11 * px + h12 * py + t1, u2 = h21 * px + h22 * py + t2;  // H1 p + h2
    const double w1 = h11 * vx + h12 * vy, w2 = h21 * vx + h22 * vy;            // H1 v
    const double a = g1 * px + g2 * py + 1.0;
    const double b = g1 * vx + g2 * vy;
    const double ia = 1.0 / a, ia2 = ia * ia, ia3 = ia2 * ia;
    xo[ip] = u1 * ia;
    xo[ip + 1] = u2 * ia;
    xo[iv] = w1 * ia - b * u1 * ia2;
    xo[iv + 1] = w2 * ia - b * u2 * ia2;
    const double g[2] = {g1, g2};
    const double H1[4] = {h11, h12, h21, h22};
    const double u[2] = {u1, u2};
    const double w[2] = {w1, w2};
    for (int r = 0; r < 2; ++r)
        for (int c = 0; c < 2; ++c) {
            double d = H1[r * 2 + c] * ia - u[r] * g[c] * ia2;
            F[(ip + r) * 8 + (ip + c)] = d;
            F[(iv + r) * 8 + (iv + c)] = d;
            F[(iv + r) * 8 + (ip + c)] =
                -(w[r] * g[c] + b * H1[r * 2 + c]) * ia2 + 2.0 * b * u[r] * g[c] * ia3;
        }
}

__global__ void __launch_bounds__(64) kalman_kernel(double* __restrict__ mean_pool, double* __restrict__ cov_pool,
                                                     double* __restrict__ tlbr_pool,
                                                     const int* __restrict__ slots, int n, int flags,
                                                     const double* __restrict__ Hm,
                                                     const int* __restrict__ h_ok,
                                                     const int* __restrict__ hold,
                                                     const double* __restrict__ meas,
                                                     const unsigned char* __restrict__ has_meas,
                                                     const double* __restrict__ mult_num,
                                                     const double* __restrict__ mult_den_pool,
                                                     FmKalmanParams prm, double frame_w, double frame_h,
                                                     double* __restrict__ out_tlbr,
                                                     unsigned char* __restrict__ out_lost) {
    __shared__ double P[64], F[64], T[64], x[8], xn[8], S[16], Si[16], K[32], KS[32], y[4];
    const int i = blockIdx.x;
    if (i >= n) return;
    if (h_ok && *h_ok == 0) return;  // camera motion estimation failed: caller clears all tracks
    if (hold && *hold != 0) return;  // the KLT box rounds have not reached their fixed point: caller reruns
    const int t = threadIdx.x;
    const int slot = slots[i];
    P[t] = cov_pool[(size_t)slot * 64 + t];
    if (t < 8) x[t] = mean_pool[(size_t)slot * 8 + t];
    __syncthreads();

    if (flags & FM_KF_WARP) {
        F[t] = 0.0;
        __syncthreads();
        if (t < 2) warp_corner(Hm, x, t, xn, F);
        __syncthreads();
        if (t < 8) x[t] = xn[t];
        mat8_mul(F, P, T, t, false);
        __syncthreads();
        mat8_mul(T, F, P, t, true);
        __syncthreads();
    }
    if (flags & FM_KF_PREDICT) {
        // kalman_filter.py:308-319; process noise scaled by the pre-predict box size
        double w = x[2] - x[0] + 1.0, h = x[3] - x[1] + 1.0;
        double sz = w > h ? w : h;
        double sd = prm.std_factor_acc * sz + prm.std_offset_acc;
        double sd2 = sd * sd;
        if (t < 8) {
            double acc = 0.0;
#pragma unroll
            for (int k = 0; k < 8; ++k) acc += prm.trans_mat[t * 8 + k] * x[k];
            xn[t] = acc;
        }
        mat8_mul(prm.trans_mat, P, T, t, false);
        __syncthreads();
        mat8_mul(T, prm.trans_mat, F, t, true);
        F[t] += prm.acc_cov[t] * sd2;
        __syncthreads();
        int r = t >> 3, c = t & 7;
        P[t] = 0.5 * (F[t] + F[c * 8 + r]);
        if (t < 8) x[t] = xn[t];
        __syncthreads();
    }
    const int mi = (flags & FM_KF_MEAS_BY_SLOT) ? slot : i;
    const bool do_meas = (flags & FM_KF_UPDATE) && (has_meas == nullptr || has_meas[mi]);
    if (do_meas) {
        // kalman_filter.py:321-345
        const bool det = flags & FM_KF_MEAS_DET;
        double w = x[2] - x[0] + 1.0, h = x[3] - x[1] + 1.0;
        double m = 1.0;
        if (mult_num) m = mult_num[i] / (mult_den_pool ? mult_den_pool[slot] : 1.0);
        if (t < 16) {
            int r = t >> 2, c = t & 3;
            double v = P[r * 8 + c];
            if (r == c) {
                double fac = det ? prm.std_factor_det[r & 1] : prm.std_factor_klt[r & 1];
                double mn = det ? prm.min_std_det[r & 1] : prm.min_std_klt[r & 1];
                double sd = fmax(fac * ((r & 1) ? h : w), mn) * m;
                v += sd * sd;
            }
            S[t] = v;
        }
        if (t < 4) y[t] = meas[(size_t)mi * 4 + t] - x[t];
        __syncthreads();
        if (t == 0) inv4(S, Si);
        __syncthreads();
        if (t < 32) {  // K = P[:, :4] * S^-1   (8x4)
            int r = t >> 2, c = t & 3;
            double acc = 0.0;
#pragma unroll
            for (int k = 0; k < 4; ++k) acc += P[r * 8 + k] * Si[k * 4 + c];
            K[t] = acc;
        }
        __syncthreads();
        if (t < 32) {  // KS = K * S
            int r = t >> 2, c = t & 3;
            double acc = 0.0;
#pragma unroll
            for (int k = 0; k < 4; ++k) acc += K[r * 4 + k] * S[k * 4 + c];
            KS[t] = acc;
        }
        if (t < 8) {
            double acc = x[t];
#pragma unroll
            for (int k = 0; k < 4; ++k) acc += K[t * 4 + k] * y[k];
            xn[t] = acc;
        }
        __syncthreads();
        {
            int r = t >> 3, c = t & 7;
            double acc = 0.0;
#pragma unroll
            for (int k = 0; k < 4; ++k) acc += KS[r * 4 + k] * K[c * 4 + k];
            P[t] -= acc;
        }
        if (t < 8) x[t] = xn[t];
        __syncthreads();
    }
    cov_pool[(size_t)slot * 64 + t] = P[t];
    if (t < 8) mean_pool[(size_t)slot * 8 + t] = x[t];
    if (t < 4 && out_tlbr) out_tlbr[(size_t)i * 4 + t] = rint(x[t]);
    if (t < 4 && tlbr_pool) tlbr_pool[(size_t)slot * 4 + t] = rint(x[t]);
    if (t == 0 && out_lost) {
        // ios(next_tlbr, frame_rect) < 0.5  (rect.py:100-109; frame_rect = [0,0,W-1,H-1])
        double x1 = rint(x[0]), y1 = rint(x[1]), x2 = rint(x[2]), y2 = rint(x[3]);
        double iw = fmin(x2, frame_w - 1.0) - fmax(x1, 0.0) + 1.0;
        double ih = fmin(y2, frame_h - 1.0) - fmax(y1, 0.0) + 1.0;
        double v = 0.0;
        if (iw > 0 && ih > 0) {
            double bw = x2 - x1 + 1.0, bh = y2 - y1 + 1.0;
            double area = (bw <= 0 || bh <= 0) ? 0.0 : bw * bh;
            v = iw * ih / area;
        }
        out_lost[i] = (v < 0.5) ? 1 : 0;
    }
}

__global__ void kalman_create_kernel(double* __restrict__ mean_pool, double* __restrict__ cov_pool,
                                     double* __restrict__ tlbr_pool, const int* __restrict__ slots,
                                     const double* __restrict__ tlbr, const int* __restrict__ tlbr_idx, int n,
                                     FmKalmanParams prm) {
    // kalman_filter.py:96-126
    int i = blockIdx.x;
    int t = threadIdx.x;
    if (i >= n) return;
    int slot = slots[i];
    const double* b = tlbr + (size_t)(tlbr_idx ? tlbr_idx[i] : i) * 4;
    double w = b[2] - b[0] + 1.0, h = b[3] - b[1] + 1.0;
    int r = t >> 3, c = t & 7;
    if (t < 4 && tlbr_pool) tlbr_pool[(size_t)slot * 4 + t] = b[t];
    double v = 0.0;
    if (r == c) {
        double wt = (r < 4) ? prm.init_pos_weight : prm.init_vel_weight;
        double sd = fmax(wt * prm.std_factor_det[r & 1] * ((r & 1) ? h : w), prm.min_std_det[r & 1]);
        v = sd * sd;
    }
    cov_pool[(size_t)slot * 64 + t] = v;
    if (t < 8) mean_pool[(size_t)slot * 8 + t] = (t < 4) ? b[t] : 0.0;
}

// kalman_filter.py:206-225 — squared Mahalanobis distance of every detection to every track's projected state.
__global__ void motion_distance_kernel(const double* __restrict__ mean_pool, const double* __restrict__ cov_pool,
                                       const int* __restrict__ slots, int n_trk, const double* __restrict__ det_tlbr,
                                       int n_det, FmKalmanParams prm, double* __restrict__ out) {
    __shared__ double L[16], pm[4];
    const int i = blockIdx.x;
    const int slot = slots ? slots[i] : i;
    if (threadIdx.x == 0) {
        const double* x = mean_pool + (size_t)slot * 8;
        const double* P = cov_pool + (size_t)slot * 64;
        double w = x[2] - x[0] + 1.0, h = x[3] - x[1] + 1.0, S[16];
        for (int r = 0; r < 4; ++r)
            for (int c = 0; c < 4; ++c) { S[r * 4 + c] = P[r * 8 + c]; L[r * 4 + c] = 0.0; }
        for (int r = 0; r < 4; ++r) {
            double sd = fmax(prm.std_factor_det[r & 1] * ((r & 1) ? h : w), prm.min_std_det[r & 1]);
            S[r * 4 + r] += sd * sd;
            pm[r] = x[r];
        }
        for (int c = 0; c < 4; ++c) {
            double d = S[c * 4 + c];
            for (int k = 0; k < c; ++k) d -= L[c * 4 + k] * L[c * 4 + k];
            d = sqrt(d);
            L[c * 4 + c] = d;
            for (int r = c + 1; r < 4; ++r) {
                double v = S[r * 4 + c];
                for (int k = 0; k < c; ++k) v -= L[r * 4 + k] * L[c * 4 + k];
                L[r * 4 + c] = v / d;
            }
        }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < n_det; j += blockDim.x) {
        const double* z = det_tlbr + (size_t)j * 4;
        double y[4];
        for (int r = 0; r < 4; ++r) {
            double v = z[r] - pm[r];
            for (int k = 0; k < r; ++k) v -= L[r * 4 + k] * y[k];
            y[r] = v / L[r * 4 + r];
        }
        out[(size_t)i * n_det + j] = y[0] * y[0] + y[1] * y[1] + y[2] * y[2] + y[3] * y[3];
    }
}

}  // namespace

extern "C" int fm_motion_distance(const double* mean_pool, const double* cov_pool, const int* slots, int n_trk,
                                  const double* det_tlbr, int n_det, const FmKalmanParams* params, double* out,
                                  void* stream) {
    FM_REQUIRE(params != nullptr, "fm_motion_distance: params is NULL");
    if (n_trk <= 0 || n_det <= 0) return FM_OK;
    motion_distance_kernel<<<n_trk, 128, 0, (cudaStream_t)stream>>>(mean_pool, cov_pool, slots, n_trk, det_tlbr,
                                                                    n_det, *params, out);
    FM_CHECK_LAUNCH("fm_motion_distance");
    return FM_OK;
}

extern "C" int fm_kalman_step_batched(double* mean_pool, double* cov_pool, double* tlbr_pool, const int* slots, int n,
                                      int flags,
                                      const double* homography, const int* h_ok, const int* hold, const double* meas,
                                      const unsigned char* has_meas, const double* mult_num,
                                      const double* mult_den_pool, const FmKalmanParams* params, double frame_w,
                                      double frame_h, double* out_tlbr, unsigned char* out_lost, void* stream) {
    FM_REQUIRE(params != nullptr, "fm_kalman_step_batched: params is NULL");
    FM_REQUIRE(!(flags & FM_KF_WARP) || homography, "fm_kalman_step_batched: WARP needs a homography");
    FM_REQUIRE(!(flags & FM_KF_UPDATE) || meas, "fm_kalman_step_batched: UPDATE needs measurements");
    if (n <= 0) return FM_OK;
    kalman_kernel<<<n, 64, 0, (cudaStream_t)stream>>>(mean_pool, cov_pool, tlbr_pool, slots, n, flags, homography, h_ok,
                                                      hold, meas,
                                                      has_meas, mult_num, mult_den_pool, *params, frame_w, frame_h,
                                                      out_tlbr, out_lost);
    FM_CHECK_LAUNCH("fm_kalman_step_batched");
    return FM_OK;
}

extern "C" int fm_kalman_create_batched(double* mean_pool, double* cov_pool, double* tlbr_pool, const int* slots,
                                        const double* tlbr, const int* tlbr_idx, int n,
                                        const FmKalmanParams* params, void* stream) {
    FM_REQUIRE(params != nullptr, "fm_kalman_create_batched: params is NULL");
    if (n <= 0) return FM_OK;
    kalman_create_kernel<<<n, 64, 0, (cudaStream_t)stream>>>(mean_pool, cov_pool, tlbr_pool, slots, tlbr, tlbr_idx,
                                                             n, *params);
    FM_CHECK_LAUNCH("fm_kalman_create_batched");
    return FM_OK;
}
