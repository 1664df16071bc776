// YOLO detector post-processing on the device: head decode fused with score threshold + compaction,
// (class, objectness) sort, per-class DIoU-NMS with a bit-mask, final box rounding and area/aspect filters.
//
// Reference: fastmot/plugins/yolo_layer.cu:127-230 (CalDetection, CalDetection_NewCoords),
//            fastmot/detector.py:322-365 (_filter_dets), fastmot/utils/rect.py:198-244 (diou_nms),
//            rect.py:48-57 (to_tlbr), :21-32 (aspect_ratio, area).
// The reference copies all K0 decoded candidates to the host and filters there; here only the D survivors
// (48 B each) leave the device.
#include "common.cuh"
#include "../../include/fastmot_b200.h"

int fm_launch_nms_scan(const unsigned long long* keys, const float* dense, const int* counter, int key_cap,
                       const unsigned long long* mask, int words, double max_area, double min_ar, int max_out,
                       double* out_tlbr, long long* out_label, double* out_conf, int* out_count, int* status,
                       cudaStream_t s);

namespace {

__device__ __forceinline__ float sigmoidf_fast(float x) { return 1.0f / (1.0f + __expf(-x)); }

// One thread per (anchor, cell).  Input NCHW-style head tensor [(5+C)*A, H, W] in fp32 or fp16.
template <typename T>
__global__ void yolo_decode_filter_kernel(const T* __restrict__ in, int yolo_w, int yolo_h, int num_anchors,
                                          FmYoloHead head, int num_classes, int input_w, int input_h, int new_coords,
                                          int nhwc,
                                          int cand_base, const unsigned char* __restrict__ label_mask,
                                          double conf_thresh, float size_w, float size_h, float off_x, float off_y,
                                          float* __restrict__ dense, unsigned long long* __restrict__ keys,
                                          int* __restrict__ counter, int key_cap) {
    const int total = yolo_w * yolo_h;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total * num_anchors) return;
    const int info_len = 5 + num_classes;
    const int anchor = idx / total, cell = idx - anchor * total;
    // planar [(5+C)*A, H, W] (the TensorRT plugin's input) or channels-last [H, W, (5+C)*A] (our conv engine)
    const size_t as = nhwc ? 1 : (size_t)total;
    const T* cur = nhwc ? in + (size_t)cell * info_len * num_anchors + (size_t)anchor * info_len
                        : in + (size_t)anchor * info_len * total + cell;
    int class_id = 0;
    float best = -INFINITY;
    for (int i = 5; i < info_len; ++i) {
        float l = (float)cur[(size_t)i * as];
        if (l > best) { best = l; class_id = i - 5; }
    }
    const float t0 = (float)cur[0], t1 = (float)cur[as], t2 = (float)cur[2 * as], t3 = (float)cur[3 * as],
                t4 = (float)cur[4 * as];
    const int row = cell / yolo_w, col = cell - row * yolo_w;
    const float s = head.scale_x_y;
    float cls_prob, box_prob, bx, by, bw, bh;
    // explicit _rn intrinsics: no FMA contraction, so the fp32 results equal the numpy oracle bit for bit
    const float half_sm1 = __fmul_rn(s - 1.0f, 0.5f);
    float ex, ey;
    if (new_coords) {
        cls_prob = best;
        box_prob = t4;
        ex = t0; ey = t1;
        bw = __fdiv_rn(__fmul_rn(__fmul_rn(__fmul_rn(t2, t2), 4.0f), head.anchors[2 * anchor]), (float)input_w);
        bh = __fdiv_rn(__fmul_rn(__fmul_rn(__fmul_rn(t3, t3), 4.0f), head.anchors[2 * anchor + 1]), (float)input_h);
    } else {
        cls_prob = sigmoidf_fast(best);
        box_prob = sigmoidf_fast(t4);
        ex = sigmoidf_fast(t0); ey = sigmoidf_fast(t1);
        bw = __fdiv_rn(__fmul_rn(__expf(t2), head.anchors[2 * anchor]), (float)input_w);
        bh = __fdiv_rn(__fmul_rn(__expf(t3), head.anchors[2 * anchor + 1]), (float)input_h);
    }
    bx = __fdiv_rn(__fadd_rn((float)col, __fsub_rn(__fmul_rn(s, ex), half_sm1)), (float)yolo_w);
    by = __fdiv_rn(__fadd_rn((float)row, __fsub_rn(__fmul_rn(s, ey), half_sm1)), (float)yolo_h);
    bx = __fsub_rn(bx, __fdiv_rn(bw, 2.0f));
    by = __fsub_rn(by, __fdiv_rn(bh, 2.0f));
    // detector.py:331-336: class mask and score threshold
    if (!label_mask[class_id]) return;
    const float score = __fmul_rn(box_prob, cls_prob);
    if (!((double)score >= conf_thresh)) return;
    // detector.py:339-341: scale to pixels (f32 <- f64 product), subtract letterbox offset
    float px = (float)((double)bx * (double)size_w);
    float py = (float)((double)by * (double)size_h);
    float pw = (float)((double)bw * (double)size_w);
    float ph = (float)((double)bh * (double)size_h);
    px = (float)((double)px - (double)off_x);
    py = (float)((double)py - (double)off_y);
    const int gidx = cand_base + idx;
    float* d = dense + (size_t)gidx * 8;
    d[0] = px; d[1] = py; d[2] = pw; d[3] = ph; d[4] = box_prob; d[5] = (float)class_id; d[6] = cls_prob;
    const int slot = atomicAdd(counter, 1);
    if (slot < key_cap) {
        // ascending u64 order == class asc, objectness desc, candidate index asc
        unsigned int sb = ~__float_as_uint(box_prob);   // objectness is >= 0
        keys[slot] = ((unsigned long long)(unsigned)class_id << 56) | ((unsigned long long)sb << 24) |
                     (unsigned long long)(gidx & 0xffffff);
    }
}

// Single-CTA bitonic sort of up to 16384 keys in shared memory.
__global__ void __launch_bounds__(1024) sort_keys_kernel(unsigned long long* __restrict__ keys,
                                                          const int* __restrict__ counter, int key_cap,
                                                          int* __restrict__ status) {
    extern __shared__ unsigned long long sk[];
    int n = *counter;
    if (n > key_cap) {
        if (threadIdx.x == 0) status[0] = 1;   // overflow: host raises
        n = key_cap;
    }
    if (n <= 1) return;
    int np2 = 1;
    while (np2 < n) np2 <<= 1;
    for (int i = threadIdx.x; i < np2; i += blockDim.x) sk[i] = i < n ? keys[i] : ~0ull;
    __syncthreads();
    for (int k = 2; k <= np2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < (np2 >> 1); t += blockDim.x) {
                int lo = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                int hi = lo | j;
                bool asc = (lo & k) == 0;
                unsigned long long a = sk[lo], b = sk[hi];
                if ((a > b) == asc) { sk[lo] = b; sk[hi] = a; }
            }
            __syncthreads();
        }
    }
    for (int i = threadIdx.x; i < n; i += blockDim.x) keys[i] = sk[i];
}

// rect.py:198-244 in fp64 on fp32-valued inputs (Numba promotes `tl + wh - 1` with an int literal to f64).
__device__ __forceinline__ bool diou_suppresses(const float* a, const float* b, double thresh) {
    const double ax = a[0], ay = a[1], bx = b[0], by = b[1];
    const double abx = (double)(a[0] + a[2]) - 1.0, aby = (double)(a[1] + a[3]) - 1.0;
    const double bbx = (double)(b[0] + b[2]) - 1.0, bby = (double)(b[1] + b[3]) - 1.0;
    const double iw = fmax(0.0, fmin(abx, bbx) - fmax(ax, bx) + 1.0);
    const double ih = fmax(0.0, fmin(aby, bby) - fmax(ay, by) + 1.0);
    const double inter = iw * ih;
    const double area_a = (double)(a[2] * a[3]), area_b = (double)(b[2] * b[3]);
    const double iou = inter / ((double)(float)(area_a + area_b) - inter);
    if (!(iou > thresh)) return false;  // DIoU <= IoU
    const double ew = fmax(abx, bbx) - fmin(ax, bx) + 1.0, eh = fmax(aby, bby) - fmin(ay, by) + 1.0;
    const double c = ew * ew + eh * eh;
    const double dx = (ax + abx) / 2 - (bx + bbx) / 2, dy = (ay + aby) / 2 - (by + bby) / 2;
    const double d = dx * dx + dy * dy;
    return iou - pow(d / c, 0.6) > thresh;
}

// mask[i][w] bit b set  <=>  sorted candidate j = 64 w + b (j > i, same class) is suppressed by i.
__global__ void __launch_bounds__(64) nms_mask_kernel(const unsigned long long* __restrict__ keys,
                                                       const float* __restrict__ dense,
                                                       const int* __restrict__ counter, int key_cap, double thresh,
                                                       unsigned long long* __restrict__ mask, int mask_words) {
    __shared__ float sb[64][4];
    __shared__ int scls[64];
    int n = min(*counter, key_cap);
    const int nb = (n + 63) >> 6;
    const int ntiles = nb * nb;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int rb = tile / nb, cb = tile - rb * nb;
        if (cb < rb) continue;  // only j > i matters
        __syncthreads();
        const int j = cb * 64 + threadIdx.x;
        if (j < n) {
            unsigned long long kj = keys[j];
            const float* d = dense + (size_t)(kj & 0xffffff) * 8;
            sb[threadIdx.x][0] = d[0]; sb[threadIdx.x][1] = d[1]; sb[threadIdx.x][2] = d[2]; sb[threadIdx.x][3] = d[3];
            scls[threadIdx.x] = (int)(kj >> 56);
        }
        __syncthreads();
        const int i = rb * 64 + threadIdx.x;
        if (i < n) {
            unsigned long long ki = keys[i];
            const float* a = dense + (size_t)(ki & 0xffffff) * 8;
            float av[4] = {a[0], a[1], a[2], a[3]};
            const int ci = (int)(ki >> 56);
            unsigned long long bits = 0;
            const int jn = min(64, n - cb * 64);
            for (int b = 0; b < jn; ++b) {
                int jj = cb * 64 + b;
                if (jj > i && scls[b] == ci && diou_suppresses(av, sb[b], thresh)) bits |= 1ull << b;
            }
            mask[(size_t)i * mask_words + cb] = bits;
        }
    }
}

// Serial greedy scan (one warp) + final filters + ordered output (detector.py:357-365).
__global__ void __launch_bounds__(32) nms_scan_kernel(const unsigned long long* __restrict__ keys,
                                                       const float* __restrict__ dense,
                                                       const int* __restrict__ counter, int key_cap,
                                                       const unsigned long long* __restrict__ mask, int mask_words,
                                                       double max_area, double min_ar, int max_out,
                                                       double* __restrict__ out_tlbr, long long* __restrict__ out_label,
                                                       double* __restrict__ out_conf, int* __restrict__ out_count) {
    extern __shared__ unsigned long long removed[];
    const int lane = threadIdx.x;
    int n = min(*counter, key_cap);
    const int nw = (n + 63) >> 6;
    for (int w = lane; w < nw; w += 32) removed[w] = 0;
    __syncwarp();
    int nout = 0;
    for (int i = 0; i < n; ++i) {
        const bool dead = (removed[i >> 6] >> (i & 63)) & 1ull;
        if (dead) continue;  // warp-uniform (shared state)
        for (int w = (i >> 6) + lane; w < nw; w += 32) removed[w] |= mask[(size_t)i * mask_words + w];
        __syncwarp();
        if (lane == 0) {
            const float* d = dense + (size_t)(keys[i] & 0xffffff) * 8;
            const double xmin = (double)d[0], ymin = (double)d[1];
            const double x1 = rint(xmin), y1 = rint(ymin);
            // to_tlbr under Numba: x + w is an f32 add, the `- 1.` literal promotes to f64 (oracle/detect.py)
            const double x2 = rint((double)(d[0] + d[2]) - 1.0), y2 = rint((double)(d[1] + d[3]) - 1.0);
            const double w = x2 - x1 + 1.0, h = y2 - y1 + 1.0;
            const double area = (w <= 0 || h <= 0) ? 0.0 : w * h;
            const double ar = w > 0 ? h / w : 0.0;
            if (area > 0 && area <= max_area && ar >= min_ar && nout < max_out) {
                out_tlbr[nout * 4 + 0] = x1; out_tlbr[nout * 4 + 1] = y1;
                out_tlbr[nout * 4 + 2] = x2; out_tlbr[nout * 4 + 3] = y2;
                out_label[nout] = (long long)d[5];
                out_conf[nout] = (double)__fmul_rn(d[4], d[6]);
                removed[nw] = 1;  // scratch flag: accepted
            } else {
                removed[nw] = 0;
            }
        }
        __syncwarp();
        nout += (int)removed[nw];
        __syncwarp();
    }
    if (lane == 0) *out_count = nout;
}

}  // namespace

extern "C" int fm_yolo_decode_filter(const void* head_out, int is_fp16, int nhwc, int yolo_w, int yolo_h, int num_anchors,
                                     const FmYoloHead* head, int num_classes, int input_w, int input_h,
                                     int new_coords, int cand_base, const unsigned char* label_mask,
                                     double conf_thresh, float size_w, float size_h, float off_x, float off_y,
                                     float* dense, unsigned long long* keys, int* counter, int key_cap, void* stream) {
    FM_REQUIRE(head != nullptr, "fm_yolo_decode_filter: head is NULL");
    FM_REQUIRE(num_anchors <= FM_MAX_ANCHORS, "fm_yolo_decode_filter: too many anchors");
    FM_REQUIRE(cand_base + yolo_w * yolo_h * num_anchors <= (1 << 24), "fm_yolo_decode_filter: > 2^24 candidates");
    int total = yolo_w * yolo_h * num_anchors;
    if (total <= 0) return FM_OK;
    dim3 grid(fm_cdiv(total, 128));
    if (is_fp16)
        yolo_decode_filter_kernel<__half><<<grid, 128, 0, (cudaStream_t)stream>>>(
            (const __half*)head_out, yolo_w, yolo_h, num_anchors, *head, num_classes, input_w, input_h, new_coords,
            nhwc, cand_base, label_mask, conf_thresh, size_w, size_h, off_x, off_y, dense, keys, counter, key_cap);
    else
        yolo_decode_filter_kernel<float><<<grid, 128, 0, (cudaStream_t)stream>>>(
            (const float*)head_out, yolo_w, yolo_h, num_anchors, *head, num_classes, input_w, input_h, new_coords,
            nhwc, cand_base, label_mask, conf_thresh, size_w, size_h, off_x, off_y, dense, keys, counter, key_cap);
    FM_CHECK_LAUNCH("fm_yolo_decode_filter");
    return FM_OK;
}

extern "C" long long fm_nms_mask_bytes(int key_cap) {
    long long words = (key_cap + 63) / 64;
    return (long long)key_cap * words * 8;
}

extern "C" int fm_diou_nms_filter(unsigned long long* keys, const float* dense, const int* counter, int key_cap,
                                  double nms_thresh, double max_area, double min_aspect_ratio,
                                  unsigned long long* mask, int max_out, double* out_tlbr, long long* out_label,
                                  double* out_conf, int* out_count, int* status, void* stream) {
    FM_REQUIRE(key_cap > 0 && key_cap <= 16384, "fm_diou_nms_filter: key_cap must be in (0, 16384]");
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(sort_keys_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 8);
        attr_set = true;
    }
    cudaStream_t s = (cudaStream_t)stream;
    int np2 = 1;
    while (np2 < key_cap) np2 <<= 1;
    cudaMemsetAsync(status, 0, sizeof(int), s);
    sort_keys_kernel<<<1, 1024, (size_t)np2 * 8, s>>>(keys, counter, key_cap, status);
    FM_CHECK_LAUNCH("sort_keys_kernel");
    const int words = (key_cap + 63) / 64;
    nms_mask_kernel<<<FM_NUM_SMS * 8, 64, 0, s>>>(keys, dense, counter, key_cap, nms_thresh, mask, words);
    FM_CHECK_LAUNCH("nms_mask_kernel");
    fm_launch_nms_scan(keys, dense, counter, key_cap, mask, words, max_area, min_aspect_ratio, max_out, out_tlbr,
                       out_label, out_conf, out_count, status, s);   // blocked scan, detect_nms.cu
    FM_CHECK_LAUNCH("nms_scan_kernel");
    return FM_OK;
}
