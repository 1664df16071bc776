// Blocked greedy NMS scan (one warp) + final detection filters with ordered, parallel output compaction.
// Replaces the serial candidate-by-candidate loop: 64 candidates per step are resolved in registers from the
// diagonal mask words, then the mask rows of the survivors are OR-ed into the tail of the removed-bitmap with
// independent loads.  Semantics identical to fastmot/utils/rect.py:198-244 + fastmot/detector.py:357-365.
#include "common.cuh"
#include "../../include/fastmot_b200.h"

namespace {

__global__ void __launch_bounds__(32) nms_scan_blocked_kernel(const unsigned long long* __restrict__ keys,
                                                               const float* __restrict__ dense,
                                                               const int* __restrict__ counter, int key_cap,
                                                               const unsigned long long* __restrict__ mask,
                                                               int mask_words, double max_area, double min_ar,
                                                               int max_out, double* __restrict__ out_tlbr,
                                                               long long* __restrict__ out_label,
                                                               double* __restrict__ out_conf,
                                                               int* __restrict__ out_count) {
    extern __shared__ unsigned long long sm[];   // removed[nw] | keep[nw]
    const int lane = threadIdx.x;
    const int n = min(*counter, key_cap);
    const int nw = (n + 63) >> 6;
    unsigned long long* removed = sm;
    unsigned long long* keep = sm + nw;
    for (int w = lane; w < nw; w += 32) { removed[w] = 0; keep[w] = 0; }
    __syncwarp();
    for (int blk = 0; blk < nw; ++blk) {
        const int base = blk << 6;
        // diagonal words of the 64 rows of this block (bit b of row r set => r suppresses base+b, b > r)
        const int r0 = base + lane, r1 = base + 32 + lane;
        const unsigned long long d0 = r0 < n ? mask[(size_t)r0 * mask_words + blk] : 0ull;
        const unsigned long long d1 = r1 < n ? mask[(size_t)r1 * mask_words + blk] : 0ull;
        unsigned long long rem = removed[blk];
        unsigned long long kept = 0;
        const int lim = min(64, n - base);
        for (int b = 0; b < lim; ++b) {
            const unsigned long long row = __shfl_sync(0xffffffffu, b < 32 ? d0 : d1, b & 31);
            if (!((rem >> b) & 1ull)) { kept |= 1ull << b; rem |= row; }
        }
        if (lane == 0) keep[blk] = kept;
        // survivors of this block suppress later blocks: lanes read consecutive words of a survivor's mask row
        // (coalesced), four independent row loads in flight per step
        for (int w0 = blk + 1; w0 < nw; w0 += 32) {
            const int w = w0 + lane;
            const bool wok = w < nw;
            unsigned long long acc = 0;
            unsigned long long k2 = kept;
            while (k2) {
                int b[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    b[q] = k2 ? __ffsll((long long)k2) - 1 : -1;
                    if (k2) k2 &= k2 - 1;
                }
                unsigned long long v[4];
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    v[q] = (b[q] >= 0 && wok) ? mask[(size_t)(base + b[q]) * mask_words + w] : 0ull;
                acc |= (v[0] | v[1]) | (v[2] | v[3]);
            }
            if (wok) removed[w] |= acc;
        }
        __syncwarp();
    }
    // final filters + ordered compaction (detector.py:357-365), 32 candidates per round
    int nout = 0;
    for (int i0 = 0; i0 < n; i0 += 32) {
        const int i = i0 + lane;
        bool ok = false;
        double x1 = 0, y1 = 0, x2 = 0, y2 = 0, conf = 0;
        long long label = 0;
        if (i < n && ((keep[i >> 6] >> (i & 63)) & 1ull)) {
            const float* d = dense + (size_t)(keys[i] & 0xffffff) * 8;
            x1 = rint((double)d[0]); y1 = rint((double)d[1]);
            // to_tlbr under Numba: x + w is an f32 add, the `- 1.` literal promotes to f64 (oracle/detect.py)
            x2 = rint((double)(d[0] + d[2]) - 1.0); y2 = rint((double)(d[1] + d[3]) - 1.0);
            const double w = x2 - x1 + 1.0, h = y2 - y1 + 1.0;
            const double area = (w <= 0 || h <= 0) ? 0.0 : w * h;
            const double ar = w > 0 ? h / w : 0.0;
            ok = area > 0 && area <= max_area && ar >= min_ar;
            label = (long long)d[5];
            conf = (double)__fmul_rn(d[4], d[6]);
        }
        const unsigned bal = __ballot_sync(0xffffffffu, ok);
        const int pos = nout + __popc(bal & ((1u << lane) - 1));
        if (ok && pos < max_out) {
            out_tlbr[pos * 4 + 0] = x1; out_tlbr[pos * 4 + 1] = y1; out_tlbr[pos * 4 + 2] = x2; out_tlbr[pos * 4 + 3] = y2;
            out_label[pos] = label;
            out_conf[pos] = conf;
        }
        nout += __popc(bal);
    }
    if (lane == 0) *out_count = min(nout, max_out);
}

}  // namespace

int fm_launch_nms_scan(const unsigned long long* keys, const float* dense, const int* counter, int key_cap,
                       const unsigned long long* mask, int words, double max_area, double min_ar, int max_out,
                       double* out_tlbr, long long* out_label, double* out_conf, int* out_count, cudaStream_t s) {
    nms_scan_blocked_kernel<<<1, 32, (size_t)(2 * words + 2) * 8, s>>>(keys, dense, counter, key_cap, mask, words,
                                                                     max_area, min_ar, max_out, out_tlbr, out_label,
                                                                     out_conf, out_count);
    return 0;
}
