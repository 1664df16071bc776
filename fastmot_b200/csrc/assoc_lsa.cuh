// Single-warp replay of SciPy's rectangular_lsap shortest-augmenting-path solver (device function shared by
// lsa_warp_kernel in assoc_lsa_block.cu and the fused association cascade in assoc_cascade.cu).
// Reference call site: fastmot/utils/matching.py:27 (scipy.optimize.linear_sum_assignment).
#pragma once
#include "common.cuh"
#include "../../include/fastmot_b200.h"

namespace lsa {

__device__ __forceinline__ unsigned long long dkey(double d) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(d);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}


// -----------------------------------------------------------------------------------------------------------------
// ONE warp, K columns per lane, every per-column quantity in registers, no barrier at all: a step is K reduced-cost
// updates per lane, two redux.sync for the 64-bit minimum, two for the tie positions, a ballot to find the owner of the
// selected column.  Same SciPy replay as the block kernels (positions in `remaining`, swap-with-last, tie rule, fp64
// operation order) -> bit-exact.  The walk along the augmenting path is done by lane 0 from the (column, path) pairs
// recorded at every selection.
// -----------------------------------------------------------------------------------------------------------------
struct LsaWarpSmem {
    double u[256], vmv[256];
    int col4row[256], vrow[256], selcol[256], selpath[256], r4c[256], linkcol[256];
};

// One warp solves the nr0 x nc0 problem (max(nr0, nc0) <= 32 K); every lane of the warp must call it.
template <int K>
__device__ void lsa_warp_solve(const double* __restrict__ cost, int nr0, int nc0, int* __restrict__ out_col4row,
                               int* __restrict__ status, LsaWarpSmem& sm) {
    double* s_u = sm.u;
    double* s_vmv = sm.vmv;
    int* s_col4row = sm.col4row;
    int* s_vrow = sm.vrow;
    int* s_selcol = sm.selcol;
    int* s_selpath = sm.selpath;
    int* s_r4c = sm.r4c;
    int* s_linkcol = sm.linkcol;
    const int lane = threadIdx.x & 31;
    const bool transpose = nc0 < nr0;
    const int nr = transpose ? nc0 : nr0;
    const int nc = transpose ? nr0 : nc0;
    const size_t si = transpose ? (size_t)1 : (size_t)nc0;
    const unsigned long long KINF = dkey(INFINITY);
    double v[K], spc[K], c_i[K], nxt[K];
    int r4c[K], path[K], pos[K];
    size_t sj[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int j = lane + 32 * k;
        sj[k] = transpose ? (size_t)j * nc0 : (size_t)j;
        v[k] = 0.0; r4c[k] = -1; path[k] = -1;
        nxt[k] = j < nc ? cost[sj[k]] : 0.0;
    }
    for (int i = lane; i < nr; i += 32) { s_u[i] = 0.0; s_col4row[i] = -1; }
    __syncwarp();
    bool infeasible = false;
    for (int cur = 0; cur < nr; ++cur) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int j = lane + 32 * k;
            spc[k] = INFINITY;
            pos[k] = j < nc ? nc - 1 - j : -1;
            c_i[k] = nxt[k];
            if (j < nc && cur + 1 < nr) nxt[k] = cost[sj[k] + (size_t)(cur + 1) * si];
        }
        int num = nc, i = cur, sink = -1, nvis = 0, nsel = 0;
        unsigned scmask = 0;                     // bit k: column lane + 32 k was selected in this row
        double minVal = 0.0;
        while (true) {
            if (lane == 0) { s_vrow[nvis] = i; s_vmv[nvis] = minVal; }
            ++nvis;
            const double ui = s_u[i];
            unsigned long long lkey = KINF;
            unsigned long long key[K];
#pragma unroll
            for (int k = 0; k < K; ++k) {
                key[k] = KINF;
                if (pos[k] >= 0) {
                    const double r = __dsub_rn(__dsub_rn(__dadd_rn(minVal, c_i[k]), ui), v[k]);
                    if (r < spc[k]) { spc[k] = r; path[k] = i; }
                    key[k] = dkey(spc[k] + 0.0);
                    lkey = min(lkey, key[k]);
                }
            }
            const unsigned hi = (unsigned)(lkey >> 32);
            const unsigned mhi = __reduce_min_sync(0xffffffffu, hi);
            const unsigned lo = hi == mhi ? (unsigned)lkey : 0xffffffffu;
            const unsigned mlo = __reduce_min_sync(0xffffffffu, lo);
            const unsigned long long gkey = ((unsigned long long)mhi << 32) | mlo;
            if (gkey == KINF) { infeasible = true; break; }
            int lU = -1, lF = 0x7fffffff;
#pragma unroll
            for (int k = 0; k < K; ++k)
                if (pos[k] >= 0 && key[k] == gkey) {
                    if (r4c[k] == -1) lU = max(lU, pos[k]);
                    lF = min(lF, pos[k]);
                }
            const int gU = __reduce_max_sync(0xffffffffu, lU);
            const int gF = __reduce_min_sync(0xffffffffu, lF);
            const int idx = gU >= 0 ? gU : gF;
            // owner of position idx -> broadcast (column, row4col, path)
            int mycol = -1, myr4c = -1, mypath = -1;
#pragma unroll
            for (int k = 0; k < K; ++k)
                if (pos[k] == idx) { mycol = lane + 32 * k; myr4c = r4c[k]; mypath = path[k]; }
            const unsigned own = __ballot_sync(0xffffffffu, mycol >= 0);
            const int ol = __ffs(own) - 1;
            const int jsel = __shfl_sync(0xffffffffu, mycol, ol);
            const int rsel = __shfl_sync(0xffffffffu, myr4c, ol);
            const int psel = __shfl_sync(0xffffffffu, mypath, ol);
            minVal = __longlong_as_double((long long)((gkey >> 63) ? (gkey & 0x7fffffffffffffffull) : ~gkey));
            if (lane == 0) { s_selcol[nsel] = jsel; s_selpath[nsel] = psel; }
            ++nsel;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                if (pos[k] == idx) { pos[k] = -1; scmask |= 1u << k; }
                else if (pos[k] == num - 1) pos[k] = idx;
            }
            --num;
            if (rsel == -1) { sink = jsel; break; }
            i = rsel;
#pragma unroll
            for (int k = 0; k < K; ++k)
                if (pos[k] >= 0) c_i[k] = cost[sj[k] + (size_t)i * si];
        }
        if (infeasible) break;
        __syncwarp();        // lane 0's visit records (s_vrow / s_vmv) -> the lanes that apply them (racecheck, r02)
        // dual updates
        if (lane == 0) s_u[cur] = __dadd_rn(s_u[cur], minVal);
        for (int t = 1 + lane; t < nvis; t += 32) {
            const int r = s_vrow[t];
            s_u[r] = __dadd_rn(s_u[r], __dsub_rn(minVal, s_vmv[t]));
        }
#pragma unroll
        for (int k = 0; k < K; ++k)
            if (scmask & (1u << k)) v[k] = __dsub_rn(v[k], __dsub_rn(minVal, spc[k]));
        // augmenting path: lane 0 walks it through the recorded (column -> path) pairs
        int nlinks = 0;
        if (lane == 0) {
            int jj = sink;
            while (true) {
                int ii = -1;
                for (int t = nsel - 1; t >= 0; --t)
                    if (s_selcol[t] == jj) { ii = s_selpath[t]; break; }
                s_linkcol[nlinks] = jj;
                s_r4c[nlinks] = ii;
                ++nlinks;
                const int tmp = s_col4row[ii];
                s_col4row[ii] = jj;
                jj = tmp;
                if (ii == cur) break;
            }
        }
        nlinks = __shfl_sync(0xffffffffu, nlinks, 0);
        __syncwarp();
        for (int t = 0; t < nlinks; ++t) {
            const int jj = s_linkcol[t], ii = s_r4c[t];
#pragma unroll
            for (int k = 0; k < K; ++k)
                if (jj == lane + 32 * k) r4c[k] = ii;
        }
        __syncwarp();
    }
    if (infeasible) {
        if (lane == 0) status[0] = 1;
        for (int k = lane; k < nr0; k += 32) out_col4row[k] = -1;
        return;
    }
    if (lane == 0) status[0] = 0;
#pragma unroll
    for (int k = 0; k < K; ++k)
        if (lane + 32 * k < nc) s_r4c[lane + 32 * k] = r4c[k];
    __syncwarp();
    for (int k = lane; k < nr0; k += 32) {
        int c = transpose ? s_r4c[k] : s_col4row[k];
        if (c >= 0 && cost[(size_t)k * nc0 + c] >= FM_INF_COST) c = -2 - c;
        out_col4row[k] = c;
    }
}

// dispatch on the problem size (both dimensions <= 256); all 32 lanes call it
__device__ inline void lsa_warp_solve_any(const double* cost, int nr, int nc, int* out_col4row, int* status, LsaWarpSmem& sm) {
    const int big = nr > nc ? nr : nc;
    if (big <= 32) lsa_warp_solve<1>(cost, nr, nc, out_col4row, status, sm);
    else if (big <= 64) lsa_warp_solve<2>(cost, nr, nc, out_col4row, status, sm);
    else if (big <= 128) lsa_warp_solve<4>(cost, nr, nc, out_col4row, status, sm);
    else lsa_warp_solve<8>(cost, nr, nc, out_col4row, status, sm);
}

}  // namespace lsa
