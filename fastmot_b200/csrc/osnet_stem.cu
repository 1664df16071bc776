// OSNet stem in one launch: conv 7x7 / stride 2 / pad 3 (3 -> 64) + bias + ReLU + max-pool 3x3 / stride 2 / pad 1
// (torchreid OSNet.conv1 + maxpool; first layers of the TensorRT engine behind fastmot/utils/inference.py:106-117).
//
// Input: fp16 [n][H + 8][W + 8][4] -- RGB + one zero channel per pixel, the crop at (+4, +4) inside a zero border
// (fm_roi_resize_norm layout 2).  With 8 bytes per pixel a kernel row of an output pixel is ONE 64-byte run
// (8 pixels x 4 channels, the first pixel gets zero weights), so the implicit-GEMM A tile needs no gather:
//     A[m = (oy, ox)][k = r * 32 + j * 4 + c] = in[2 oy + 1 + r][2 ox + j][c]          r = 0..7, j = 0..7
// is a 5-D TMA tensor (32 elements | kernel row r | 64 ox, stride 16 B | oy, stride 2 rows | crop) whose box
// {32, 1, 64, 2, 1} lands as a [128 pixels][32 K] tile of 64-byte rows, 64-byte swizzled (a 64-byte inner box cannot
// share a 128-byte swizzle row with a second kernel row: TMA pads it): seven 8 KB TMA loads per 128 output pixels
// (two conv rows), K = 7 x 32 (j = 0 carries zero weights).  tcgen05.mma 128 x 64 x 16 on SWIZZLE_64B descriptors,
// accumulators double buffered in TMEM; the weights (28 KB image) stay in shared memory for the whole CTA.
// One CTA = NP pooled rows of one crop: it walks the conv-row pairs ("tiles") top to bottom, keeps the last three
// conv rows (ReLU'd fp16) in shared memory and emits one pooled row per tile.  Post-ReLU values are >= 0, so the
// max-pool padding is a plain zero.
#include "tc_common.cuh"
#include "../../include/fastmot_b200.h"

namespace {

using namespace tc;

constexpr int STEM_NS = 6;             // TMA ring depth (8 KB stages, one kernel row each)
constexpr int STEM_KR = 7;             // kernel rows
constexpr int STEM_ROWP = 64 * 128 + 64 * 16;   // one conv row: 64 px x (128 B + 16 B pad)

__device__ float* g_stem_dbg = nullptr;        // debugging aid (scripts/debug_stem.py): raw accumulators + A slices of CTA 0, tile 0

struct StemArgs {
    int n, np, bands;            // crops, pooled rows per CTA, CTAs per crop
    const uint8_t* wimg;         // 7 slices (kernel rows) of [64 x 64 B], SWIZZLE_64B
    const float* bias;           // [64]
    __half* out;                 // [n][64][32][64]
};

__global__ void __launch_bounds__(288, 2) osnet_stem_kernel(const __grid_constant__ CUtensorMap map_x, StemArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* s_w = smem;                                  // 7 x 4 KB weights (32 KB reserved)
    uint8_t* s_ring = smem + 32768;                       // STEM_NS x 8 KB A slices
    uint8_t* s_rows = s_ring + STEM_NS * 8192;            // 3 conv rows, pixel-major, 144 B per pixel
    __shared__ uint64_t w_full, ring_full[STEM_NS], ring_empty[STEM_NS], acc_full[2], acc_empty[2];
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    fm_pdl_trigger();
    const int crop = blockIdx.x / a.bands, band = blockIdx.x - crop * a.bands;
    const int p0 = band * a.np;                           // first pooled row
    const int t0 = p0 > 0 ? p0 - 1 : 0;                   // first tile (conv rows 2t, 2t + 1)
    const int t1 = p0 + a.np;                             // one past the last tile
    const int ntiles = t1 - t0;
    if (tid == 0) {
        if (smem_u32(smem) & 1023u) __trap();
        mbar_init(&w_full, 1);
        for (int i = 0; i < STEM_NS; ++i) { mbar_init(&ring_full[i], 1); mbar_init(&ring_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 8); }
        mbar_fence_init();
    }
    if (warp == 8) {
        tmem_alloc<128>(&s_tmem);
        if (lane == 0) tma_prefetch_desc(&map_x);
    }
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = s_tmem;

    if (warp == 8) {
        const bool leader = lane == 0;
        if (leader) {
            mbar_expect_tx(&w_full, STEM_KR * 4096);
            bulk_load(s_w, a.wimg, STEM_KR * 4096, &w_full);
        }
        fm_pdl_wait();
        const uint32_t idesc = idesc_f16(64);
        const int iters = ntiles * STEM_KR;
        auto issue = [&](int i) {
            const int st = i % STEM_NS, t = t0 + i / STEM_KR, kr = i % STEM_KR;
            if (i >= STEM_NS) mbar_wait(&ring_empty[st], (uint32_t)((i / STEM_NS - 1) & 1));
            if (leader) {
                mbar_expect_tx(&ring_full[st], 8192);
                tma_load_5d(s_ring + st * 8192, &map_x, &ring_full[st], 0, kr, 0, 2 * t, crop);
            }
        };
        for (int i = 0; i < STEM_NS - 1 && i < iters; ++i) issue(i);
        mbar_wait(&w_full, 0);
        for (int i = 0; i < iters; ++i) {
            if (i + STEM_NS - 1 < iters) issue(i + STEM_NS - 1);
            const int st = i % STEM_NS, tl = i / STEM_KR, ks = i % STEM_KR, ai = tl & 1;
            if (ks == 0 && tl >= 2) mbar_wait(&acc_empty[ai], (uint32_t)(((tl >> 1) - 1) & 1));
            mbar_wait(&ring_full[st], (uint32_t)((i / STEM_NS) & 1));
            fence_after();
            if (g_stem_dbg && blockIdx.x == 0 && tl == 0) {
                const uint32_t* src = reinterpret_cast<const uint32_t*>(s_ring + st * 8192);
                uint32_t* dstw = reinterpret_cast<uint32_t*>(g_stem_dbg + 128 * 64) + ks * 2048;
                for (int w = lane; w < 2048; w += 32) dstw[w] = src[w];
                __syncwarp();
            }
            const uint32_t sa = smem_u32(s_ring + st * 8192), sb = smem_u32(s_w) + ks * 4096;
            if (leader) {
#pragma unroll
                for (int k = 0; k < 2; ++k)
                    mma_ss(tmem + ai * 64, smem_desc_sw64(sa + k * 32), smem_desc_sw64(sb + k * 32), idesc,
                           (ks > 0 || k > 0) ? 1u : 0u);
                commit(&ring_empty[st]);
                if (ks == STEM_KR - 1) commit(&acc_full[ai]);
            }
            __syncwarp();
        }
    } else {
        // 8 epilogue / pooling warps: lane quarter q = warp & 3 (conv row q >> 1, ox half q & 1), channel half warp >> 2
        fm_pdl_wait();
        const int q = warp & 3, hsel = warp >> 2;
        const int oyl = q >> 1, ox = (q & 1) * 32 + lane;
        const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
        float b[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) b[i] = __ldg(a.bias + hsel * 32 + i);
        for (int tl = 0; tl < ntiles; ++tl) {
            const int t = t0 + tl, ai = tl & 1;
            mbar_wait_sleep(&acc_full[ai], (uint32_t)((tl >> 1) & 1));
            fence_after();
            uint32_t r[32];
            tmem_ld32(lane_base + ai * 64 + hsel * 32, r);
            tmem_ld_wait();
            if (g_stem_dbg && blockIdx.x == 0 && tl == 0)
                for (int e = 0; e < 32; ++e) g_stem_dbg[(q * 32 + lane) * 64 + hsel * 32 + e] = __uint_as_float(r[e]);
            fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[ai]);
            // conv row 2t + oyl -> ring slot (2t + oyl) % 3
            uint8_t* dst = s_rows + ((2 * t + oyl) % 3) * STEM_ROWP + ox * 144 + hsel * 64;
#pragma unroll
            for (int e = 0; e < 4; ++e)
                *reinterpret_cast<uint4*>(dst + e * 16) = make_uint4(
                    pack_h2(fmaxf(__uint_as_float(r[e * 8 + 0]) + b[e * 8 + 0], 0.f), fmaxf(__uint_as_float(r[e * 8 + 1]) + b[e * 8 + 1], 0.f)),
                    pack_h2(fmaxf(__uint_as_float(r[e * 8 + 2]) + b[e * 8 + 2], 0.f), fmaxf(__uint_as_float(r[e * 8 + 3]) + b[e * 8 + 3], 0.f)),
                    pack_h2(fmaxf(__uint_as_float(r[e * 8 + 4]) + b[e * 8 + 4], 0.f), fmaxf(__uint_as_float(r[e * 8 + 5]) + b[e * 8 + 5], 0.f)),
                    pack_h2(fmaxf(__uint_as_float(r[e * 8 + 6]) + b[e * 8 + 6], 0.f), fmaxf(__uint_as_float(r[e * 8 + 7]) + b[e * 8 + 7], 0.f)));
            asm volatile("bar.sync 1, 256;" ::: "memory");
            // pooled row py = t: conv rows 2t - 1, 2t, 2t + 1 (row -1 does not exist: values are >= 0, use 0)
            if (t >= p0) {
                const bool has_prev = t > 0;
                for (int it = tid; it < 32 * 8; it += 256) {
                    const int px = it >> 3, c8 = it & 7;
                    __half2 m[4];
                    const __half2 z = __float2half2_rn(0.f);
                    m[0] = m[1] = m[2] = m[3] = z;
#pragma unroll
                    for (int dy = -1; dy <= 1; ++dy) {
                        if (dy < 0 && !has_prev) continue;
                        const uint8_t* rowp = s_rows + ((2 * t + dy + 3) % 3) * STEM_ROWP + c8 * 16;
#pragma unroll
                        for (int dx = -1; dx <= 1; ++dx) {
                            const int cx = 2 * px + dx;
                            if (cx < 0) continue;
                            const uint4 v = *reinterpret_cast<const uint4*>(rowp + cx * 144);
                            const __half2* h = reinterpret_cast<const __half2*>(&v);
                            m[0] = __hmax2(m[0], h[0]); m[1] = __hmax2(m[1], h[1]);
                            m[2] = __hmax2(m[2], h[2]); m[3] = __hmax2(m[3], h[3]);
                        }
                    }
                    *reinterpret_cast<uint4*>(a.out + (((size_t)crop * 64 + t) * 32 + px) * 64 + c8 * 8) =
                        make_uint4(*reinterpret_cast<uint32_t*>(&m[0]), *reinterpret_cast<uint32_t*>(&m[1]),
                                   *reinterpret_cast<uint32_t*>(&m[2]), *reinterpret_cast<uint32_t*>(&m[3]));
                }
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 8) tmem_dealloc<128>(tmem);
}

}  // namespace

extern "C" int fm_osnet_stem_set_debug(void* p) {      // not part of the public header
    float* q = (float*)p;
    cudaMemcpyToSymbol(g_stem_dbg, &q, sizeof(q));
    return FM_OK;
}

// x: fp16 [n][264][136][4] (fm_roi_resize_norm layout 2, zero border); wimg: pack_b_sw64 image of the [64][224]
// weight matrix W[cout][r * 32 + j * 4 + c] = w[cout][r][j - 1][c] (zero for j == 0 and c == 3); bias fp32 [64];
// out: fp16 [n][64][32][64] NHWC (after ReLU and the 3x3 / 2 max-pool).
extern "C" int fm_osnet_stem(const void* x, int n, const void* wimg, const float* bias, void* out, void* stream) {
    if (n <= 0) return FM_OK;
    fm_encode_tiled_fn enc = fm_get_encode_tiled();
    FM_REQUIRE(enc != nullptr, "fm_osnet_stem: cuTensorMapEncodeTiled not available");
    const uint64_t P = 136 * 4;                           // elements per padded input row
    CUtensorMap map;
    cuuint64_t dims[5] = {32, 8, 64, 128, (cuuint64_t)n};
    cuuint64_t strides[4] = {P * 2, 16, 2 * P * 2, 264 * P * 2};    // bytes, dims 1..4
    cuuint32_t box[5] = {32, 1, 64, 2, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    void* base = (void*)((const __half*)x + P);           // first conv row reads padded row 1 (= image row -3)
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, base, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char buf[128];
        snprintf(buf, sizeof buf, "fm_osnet_stem: cuTensorMapEncodeTiled failed (%d)", (int)r);
        fm_set_last_error(buf);
        return FM_ERR_CUDA;
    }
    StemArgs a;
    a.n = n; a.np = 8; a.bands = 64 / a.np;
    a.wimg = (const uint8_t*)wimg; a.bias = bias; a.out = (__half*)out;
    const int smem = 32768 + STEM_NS * 8192 + 3 * STEM_ROWP;
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(osnet_stem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        attr = true;
    }
    cudaError_t e = fm_launch_pdl(osnet_stem_kernel, dim3(n * a.bands), dim3(288), (size_t)smem, (cudaStream_t)stream, map, a);
    if (e != cudaSuccess) { fm_set_last_error(cudaGetErrorString(e)); return FM_ERR_CUDA; }
    FM_CHECK_LAUNCH("fm_osnet_stem");
    return FM_OK;
}
