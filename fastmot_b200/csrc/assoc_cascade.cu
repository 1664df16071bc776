// Fused association cascade: every assignment stage of MultiTracker.update (fastmot/tracker.py:185-247) in ONE
// launch with the id lists kept on the device -- no host round trip between the stages.
//
//   stage 1  per age-depth group of confirmed tracks: fused appearance + motion cost -> LSA      (tracker.py:205-214)
//   stage 2  still-active leftovers of stage 1: IoU cost -> LSA                                   (:216-220)
//   stage 3  unconfirmed tracks: IoU cost -> LSA                                                  (:222-223)
//   re-id    lost-track history vs confident, non-occluded leftovers: greedy on the ReID cost     (:225-233)
//
// A cost entry depends only on its (track, detection) pair, so the three cost matrices are computed ONCE for all rows
// x all detections by the grid-wide kernels of assoc.cu (fm_matching_cost / fm_iou_cost, identity selections); this
// kernel (one CTA) then, per stage, gathers the sub-matrix of the rows and the still-unmatched detections, solves it
// with the single-warp SciPy replay (assoc_lsa.cuh), and rebuilds the lists exactly as the reference does:
// matches in row order, unmatched rows / columns in the iteration order of Numba's typed set
// (fastmot/utils/matching.py:57-70, restated on the host in fastmot_b200/utils/numba_compat.py), demoted pairs
// (cost >= INF_COST) appended.  The order matters: it decides LSA ties in later stages and the order new track ids are
// handed out.  Every dimension (group size, detections, unconfirmed, history) must be <= 256; larger frames take
// the per-stage host-driven path of fastmot_b200/tracker.py.
#include "assoc_lsa.cuh"

namespace {

constexpr int CAS_MAX = 256;
constexpr int CAS_TABLE = 2048;        // Numba set table for n <= 256: 16 -> 512 -> x4

struct CasSmem {
    lsa::LsaWarpSmem lsa;
    int rows[CAS_MAX], cols[CAS_MAX], c4r[CAS_MAX], order[CAS_MAX];
    int u_det[CAS_MAX], u_trk1[CAS_MAX], tmp[CAS_MAX];
    int table[CAS_TABLE];
    unsigned char removed[CAS_MAX], rdead[CAS_MAX], cdead[CAS_MAX];
    double red_val[32];
    int red_idx[32];
    int n_u_det, n_u_trk1, status, stop;
};

// Ordered compaction over [0, n): emit(i, position) for every i with pred(i), positions base, base + 1, ... in index
// order.  Block-wide (every thread calls it); returns base + count.
template <typename Pred, typename Emit>
__device__ int compact(CasSmem& sm, int n, int base, Pred pred, Emit emit) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;
    int total = base;
    for (int c0 = 0; c0 < n; c0 += blockDim.x) {
        const int i = c0 + tid;
        const bool f = i < n && pred(i);
        const unsigned bal = __ballot_sync(0xffffffffu, f);
        if (lane == 0) sm.red_idx[wid] = __popc(bal);
        __syncthreads();
        int before = 0, all = 0;
        for (int w = 0; w < nw; ++w) {
            const int c = sm.red_idx[w];
            all += c;
            if (w < wid) before += c;
        }
        if (f) emit(i, total + before + __popc(bal & ((1u << lane) - 1)));
        __syncthreads();
        total += all;
    }
    return total;
}

// list(set(range(n)) - set(removed)) under Numba's typed set (numba_compat.set_difference_order), n <= 256.
// out[0 .. return) = the surviving indices in the set's iteration order.  Block-wide.  Only the re-insertion of the
// survivors into the shrunk table is serial (it is the hash-probe order that defines the result); the table is
// cleared and read back by the whole CTA.
__device__ int numba_set_order(CasSmem& sm, int n, const unsigned char* removed, int* out) {
    const int tid = threadIdx.x;
    int* table = sm.table;
    const int nk = compact(sm, n, 0, [&](int v) { return !removed[v]; }, [&](int v, int pos) { out[pos] = v; });
    int size = 16;
    while (size < 2 * n) size <<= 1;
    if (2 * n >= size) size <<= 2;
    const int min_entries = 2 * nk > 16 ? 2 * nk : 16;
    if (!(size >= 4 * min_entries && size > 16)) return nk;       // no shrink: slot == value, ascending
    int new_size = size;
    while ((new_size >> 1) >= min_entries) new_size >>= 1;
    const int m = new_size - 1;
    for (int i = tid; i < new_size; i += blockDim.x) table[i] = -1;
    __syncthreads();
    if (tid == 0) {
        for (int k = 0; k < nk; ++k) {
            const int v = out[k];
            int i = v & m;
            bool found = false;
            for (int t = 0; t < 3; ++t) {
                if (table[i] < 0) { found = true; break; }
                i = (i + 1) & m;
            }
            if (!found) {
                unsigned long long perturb = (unsigned long long)v;
                while (table[i] >= 0) {
                    perturb >>= 5;
                    i = (int)(((unsigned long long)i * 5ull + 1ull + perturb) & (unsigned long long)m);
                }
            }
            table[i] = v;
        }
    }
    __syncthreads();
    return compact(sm, new_size, 0, [&](int i) { return table[i] >= 0; }, [&](int i, int pos) { out[pos] = table[i]; });
}

struct OutLists {
    int* hdr;
    int *m_row[3], *m_det[3], *u[3], *reid_row, *reid_det, *invalid, *reid_u, *occ;
};

__device__ OutLists carve_out(int* out, int cap) {
    OutLists o;
    o.hdr = out;
    int* p = out + 16;
    for (int s = 0; s < 3; ++s) { o.m_row[s] = p; p += cap; o.m_det[s] = p; p += cap; }
    for (int s = 0; s < 3; ++s) { o.u[s] = p; p += cap; }
    o.reid_row = p; p += cap;
    o.reid_det = p; p += cap;
    o.invalid = p; p += cap;
    o.reid_u = p; p += cap;
    o.occ = p;
    return o;
}

// One LSA stage: rows sm.rows[0..nr) (global row ids), columns sm.u_det; appends matches / unmatched rows to the given
// lists and replaces sm.u_det by the unmatched columns.  Block-wide call (all threads).
__device__ void lsa_stage(const FmCascadeDesc& d, CasSmem& sm, const double* cost, int nr, int* m_row, int* m_det,
                          int& n_m, int* u_rows, int& n_u) {
    const int tid = threadIdx.x;
    const int nc = sm.n_u_det;
    if (nr == 0) return;
    if (nc == 0) {                                   // `_solve`: every row stays unmatched, in order
        if (tid == 0) {
            for (int r = 0; r < nr; ++r) u_rows[n_u + r] = sm.rows[r];
        }
        __syncthreads();
        n_u += nr;
        return;
    }
    for (int e = tid; e < nr * nc; e += blockDim.x) {
        const int i = e / nc, j = e - i * nc;
        d.sub[e] = cost[(size_t)sm.rows[i] * d.n_det + sm.u_det[j]];
    }
    __syncthreads();
    if (tid < 32) lsa::lsa_warp_solve_any(d.sub, nr, nc, sm.c4r, &sm.status, sm.lsa);
    __syncthreads();
    if (sm.status != 0) return;
    // ---- `_get_assignment_matches` (matching.py:57-70), every list built by ordered block-wide compaction ----
    const int* c4r = sm.c4r;
    n_m = compact(sm, nr, n_m, [&](int r) { return c4r[r] >= 0; },
                  [&](int r, int pos) { m_row[pos] = sm.rows[r]; m_det[pos] = sm.u_det[c4r[r]]; });
    // unmatched rows: Numba set order of the unassigned, then the demoted (cost >= INF) in row order
    for (int r = tid; r < nr; r += blockDim.x) sm.removed[r] = c4r[r] != -1;
    __syncthreads();
    const int n_keep = numba_set_order(sm, nr, sm.removed, sm.order);
    for (int k = tid; k < n_keep; k += blockDim.x) u_rows[n_u + k] = sm.rows[sm.order[k]];
    __syncthreads();
    n_u = compact(sm, nr, n_u + n_keep, [&](int r) { return c4r[r] <= -2; }, [&](int r, int pos) { u_rows[pos] = sm.rows[r]; });
    // unmatched columns, same rule
    for (int c = tid; c < nc; c += blockDim.x) sm.removed[c] = 0;
    __syncthreads();
    for (int r = tid; r < nr; r += blockDim.x) {
        const int c = c4r[r];
        if (c != -1) sm.removed[c >= 0 ? c : -2 - c] = 1;
    }
    __syncthreads();
    const int n_keepc = numba_set_order(sm, nc, sm.removed, sm.order);
    for (int k = tid; k < n_keepc; k += blockDim.x) sm.tmp[k] = sm.u_det[sm.order[k]];
    __syncthreads();
    const int n_ucol = compact(sm, nr, n_keepc, [&](int r) { return c4r[r] <= -2; },
                               [&](int r, int pos) { sm.tmp[pos] = sm.u_det[-2 - c4r[r]]; });
    __syncthreads();
    for (int k = tid; k < n_ucol; k += blockDim.x) sm.u_det[k] = sm.tmp[k];
    if (tid == 0) sm.n_u_det = n_ucol;
    __syncthreads();
}

// greedy_match (matching.py:73-97) on the gathered sub-matrix; block-wide
__device__ void greedy_stage(const FmCascadeDesc& d, CasSmem& sm, int nr, int nc, double max_cost) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;
    for (int k = tid; k < nr; k += blockDim.x) { sm.rdead[k] = 0; sm.c4r[k] = -1; sm.order[k] = -1; }
    for (int k = tid; k < nc; k += blockDim.x) sm.cdead[k] = 0;
    if (tid == 0) sm.stop = 0;
    __syncthreads();
    const int total = nr * nc;
    const int max_iter = nr < nc ? nr : nc;
    for (int iter = 0; iter < max_iter; ++iter) {
        double best = INFINITY;
        int bidx = 0x7fffffff;
        for (int e = tid; e < total; e += blockDim.x) {
            const int r = e / nc, c = e - r * nc;
            if (sm.rdead[r] || sm.cdead[c]) continue;
            const double v = d.sub[e];
            if (v < best || (v == best && e < bidx)) { best = v; bidx = e; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
            if (ov < best || (ov == best && oi < bidx)) { best = ov; bidx = oi; }
        }
        if (lane == 0) { sm.red_val[wid] = best; sm.red_idx[wid] = bidx; }
        __syncthreads();
        if (wid == 0) {
            best = lane < nw ? sm.red_val[lane] : INFINITY;
            bidx = lane < nw ? sm.red_idx[lane] : 0x7fffffff;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ov = __shfl_xor_sync(0xffffffffu, best, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
                if (ov < best || (ov == best && oi < bidx)) { best = ov; bidx = oi; }
            }
            if (lane == 0) {
                if (bidx != 0x7fffffff && best <= max_cost) {
                    const int r = bidx / nc, c = bidx - r * nc;
                    sm.rdead[r] = 1; sm.cdead[c] = 1;
                    sm.c4r[r] = c; sm.order[r] = iter;
                } else {
                    sm.stop = 1;
                }
            }
        }
        __syncthreads();
        if (sm.stop) break;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(1024) assoc_cascade_kernel(FmCascadeDesc d) {
    extern __shared__ __align__(16) unsigned char cas_raw[];
    CasSmem& sm = *reinterpret_cast<CasSmem*>(cas_raw);
    const int tid = threadIdx.x;
    OutLists o = carve_out(d.out, d.cap);
    if (tid == 0) { sm.status = 0; sm.n_u_trk1 = 0; sm.n_u_det = d.n_det; }
    for (int j = tid; j < d.n_det; j += blockDim.x) {
        sm.u_det[j] = j;
        o.occ[j] = d.det_occluded[j];
    }
    __syncthreads();
    int n_m[3] = {0, 0, 0}, n_u[3] = {0, 0, 0};
    // ---- stage 1: confirmed tracks, youngest depth group first ----
    for (int g = 0; g < d.n_groups && sm.status == 0; ++g) {
        const int r0 = d.goff[g], nr = d.goff[g + 1] - r0;
        if (nr == 0) continue;
        for (int k = tid; k < nr; k += blockDim.x) sm.rows[k] = r0 + k;
        __syncthreads();
        lsa_stage(d, sm, d.feat_cost, nr, o.m_row[0], o.m_det[0], n_m[0], sm.u_trk1, n_u[0]);
    }
    // ---- stage 2: active leftovers vs IoU ----
    int n_u1_inactive = 0;
    if (sm.status == 0) {
        if (tid == 0) {
            int na = 0, ni = 0;
            for (int k = 0; k < n_u[0]; ++k) {
                const int r = sm.u_trk1[k];
                if (d.conf_active[r]) sm.rows[na++] = r; else o.u[0][ni++] = r;
            }
            sm.cols[2] = na;
            sm.cols[3] = ni;
        }
        __syncthreads();
        const int na = sm.cols[2];
        n_u1_inactive = sm.cols[3];
        __syncthreads();
        lsa_stage(d, sm, d.iou_cost, na, o.m_row[1], o.m_det[1], n_m[1], o.u[1], n_u[1]);
    }
    // ---- stage 3: unconfirmed tracks vs IoU (rows n_conf .. n_conf + n_unconf of the IoU matrix) ----
    if (sm.status == 0) {
        for (int k = tid; k < d.n_unconf; k += blockDim.x) sm.rows[k] = d.n_conf + k;
        __syncthreads();
        lsa_stage(d, sm, d.iou_cost, d.n_unconf, o.m_row[2], o.m_det[2], n_m[2], o.u[2], n_u[2]);
    }
    // ---- re-identification ----
    int n_reid = 0, n_invalid = 0, n_reid_u = 0;
    if (sm.status == 0) {
        if (tid == 0) {
            int nv = 0, ni = 0;
            for (int k = 0; k < sm.n_u_det; ++k) {
                const int dj = sm.u_det[k];
                if (!(d.det_conf[dj] >= d.conf_thresh)) continue;
                if (d.det_occluded[dj]) o.invalid[ni++] = dj; else sm.cols[nv++] = dj;
            }
            sm.tmp[0] = nv;
            sm.tmp[1] = ni;
        }
        __syncthreads();
        const int nv = sm.tmp[0];
        n_invalid = sm.tmp[1];
        __syncthreads();
        const int nh = d.n_hist;
        if (nh == 0 || nv == 0) {
            for (int k = tid; k < nv; k += blockDim.x) o.reid_u[k] = sm.cols[k];
            n_reid_u = nv;
        } else {
            for (int e = tid; e < nh * nv; e += blockDim.x) {
                const int i = e / nv, j = e - i * nv;
                d.sub[e] = d.reid_cost[(size_t)i * d.n_det + sm.cols[j]];
            }
            __syncthreads();
            greedy_stage(d, sm, nh, nv, d.max_reid_cost);
            if (tid == 0) {
                // `_get_greedy_matches`: matches in discovery order, leftover columns in index order
                int nm = 0;
                const int iters = nh < nv ? nh : nv;
                for (int it = 0; it < iters; ++it)
                    for (int r = 0; r < nh; ++r)
                        if (sm.order[r] == it) { o.reid_row[nm] = r; o.reid_det[nm] = sm.cols[sm.c4r[r]]; ++nm; }
                int nu = 0;
                for (int c = 0; c < nv; ++c)
                    if (!sm.cdead[c]) o.reid_u[nu++] = sm.cols[c];
                sm.tmp[0] = nm;
                sm.tmp[1] = nu;
            }
            __syncthreads();
            n_reid = sm.tmp[0];
            n_reid_u = sm.tmp[1];
        }
    }
    if (tid == 0) {
        o.hdr[0] = sm.status;
        o.hdr[1] = n_m[0]; o.hdr[2] = n_m[1]; o.hdr[3] = n_m[2];
        o.hdr[4] = n_u1_inactive; o.hdr[5] = n_u[1]; o.hdr[6] = n_u[2];
        o.hdr[7] = n_reid; o.hdr[8] = n_invalid; o.hdr[9] = n_reid_u;
    }
}

}  // namespace

extern "C" long long fm_assoc_cascade_out_ints(int cap) { return 16 + 14LL * cap; }

extern "C" int fm_assoc_cascade(const FmCascadeDesc* d, void* stream) {
    FM_REQUIRE(d != nullptr, "fm_assoc_cascade: desc is NULL");
    FM_REQUIRE(d->n_det >= 0 && d->n_det <= CAS_MAX && d->n_conf >= 0 && d->n_conf <= CAS_MAX && d->n_unconf >= 0 &&
                   d->n_unconf <= CAS_MAX && d->n_hist >= 0 && d->n_hist <= CAS_MAX,
               "fm_assoc_cascade: every dimension must be <= 256 (larger frames: per-stage path)");
    FM_REQUIRE(d->cap >= d->n_det && d->cap >= d->n_conf + d->n_unconf && d->cap >= d->n_hist, "fm_assoc_cascade: cap");
    FM_REQUIRE(d->out != nullptr && d->sub != nullptr, "fm_assoc_cascade: out / sub is NULL");
    FM_REQUIRE(d->n_groups >= 0 && (d->n_groups == 0 || d->goff != nullptr), "fm_assoc_cascade: goff");
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(assoc_cascade_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(CasSmem));
        attr = true;
    }
    assoc_cascade_kernel<<<1, 1024, sizeof(CasSmem), (cudaStream_t)stream>>>(*d);
    FM_CHECK_LAUNCH("fm_assoc_cascade");
    return FM_OK;
}
