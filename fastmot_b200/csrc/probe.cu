// Self-test of the tensor-core primitives the fused kernels are built from (tests/test_gpu_primitives.py):
//   TMA tensor-map load into 128-byte-swizzled shared memory  ->  tcgen05.mma (A, B from shared memory)
//   -> tcgen05.ld -> fp16 -> tcgen05.st (activation tile kept in TMEM)  ->  tcgen05.mma with A from TMEM
//   -> tcgen05.ld -> global.
// One CTA, 128 threads; A is [128 x 64] fp16, B1 [n1 x 64] and B2 [n2 x n1] are pre-swizzled K-slice images
// (fm_pack_b_sw128 layout, moved by cp.async.bulk).
#include "tc_common.cuh"
#include "../../include/fastmot_b200.h"

static fm_encode_tiled_fn g_encode = nullptr;

fm_encode_tiled_fn fm_get_encode_tiled() {
    if (g_encode) return g_encode;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
        return nullptr;
    g_encode = (fm_encode_tiled_fn)fn;
    return g_encode;
}

int fm_make_tmap_f16_3d(CUtensorMap* out, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1,
                        uint64_t stride2, uint32_t b0, uint32_t b1, uint32_t b2) {
    fm_encode_tiled_fn enc = fm_get_encode_tiled();
    if (!enc) { fm_set_last_error("cuTensorMapEncodeTiled not available"); return FM_ERR_CUDA; }
    cuuint64_t dims[3] = {d0, d1, d2};
    cuuint64_t strides[2] = {stride1 * 2, stride2 * 2};     // bytes, dims 1..2
    cuuint32_t box[3] = {b0, b1, b2};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char buf[128];
        snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled failed (%d)", (int)r);
        fm_set_last_error(buf);
        return FM_ERR_CUDA;
    }
    return FM_OK;
}

int fm_make_tmap_f16_nd(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides,
                        const uint32_t* box) {
    fm_encode_tiled_fn enc = fm_get_encode_tiled();
    if (!enc) { fm_set_last_error("cuTensorMapEncodeTiled not available"); return FM_ERR_CUDA; }
    if (rank < 1 || rank > 5) { fm_set_last_error("fm_make_tmap_f16_nd: rank"); return FM_ERR_ARG; }
    cuuint64_t d[5], st[4];
    cuuint32_t bx[5], estr[5] = {1, 1, 1, 1, 1};
    for (int i = 0; i < rank; ++i) { d[i] = dims[i]; bx[i] = box[i]; }
    for (int i = 0; i + 1 < rank; ++i) st[i] = strides[i] * 2;      // bytes
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, const_cast<void*>(base), d, st, bx, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char buf[128];
        snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled (rank %d) failed (%d)", rank, (int)r);
        fm_set_last_error(buf);
        return FM_ERR_CUDA;
    }
    return FM_OK;
}

namespace {

__global__ void __launch_bounds__(128) probe_kernel(const __grid_constant__ CUtensorMap map_a, int row0,
                                                     const uint8_t* __restrict__ b1, const uint8_t* __restrict__ b2,
                                                     int n1, int n2, float* __restrict__ out0, float* __restrict__ out1) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sA = smem;                       // 128 x 128 B
    uint8_t* sB1 = sA + 16384;                // n1 x 128 B
    uint8_t* sB2 = sB1 + 128 * 128;           // ceil(n1 / 64) slices of n2 x 128 B
    __shared__ uint64_t bar_load, bar_mma;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nsl2 = (n1 + 63) / 64;
    if (tid == 0) {
        tc::mbar_init(&bar_load, 1);
        tc::mbar_init(&bar_mma, 1);
        tc::mbar_fence_init();
    }
    if (warp == 0) tc::tmem_alloc<512>(&s_tmem);
    tc::fence_before();
    __syncthreads();
    tc::fence_after();
    const uint32_t tmem = s_tmem;
    if (tid == 0) {
        const uint32_t bytes = 16384 + n1 * 128 + nsl2 * n2 * 128;
        tc::mbar_expect_tx(&bar_load, bytes);
        // four boxes of 32 rows each (the fused kernels load one image row per box)
        for (int r = 0; r < 4; ++r) tc::tma_load_3d(sA + r * 4096, &map_a, &bar_load, 0, row0 + r * 32, 0);
        tc::bulk_load(sB1, b1, n1 * 128, &bar_load);
        tc::bulk_load(sB2, b2, nsl2 * n2 * 128, &bar_load);
    }
    tc::mbar_wait(&bar_load, 0);
    if (tid == 0) {
        tc::fence_after();
        const uint32_t idesc = tc::idesc_f16(n1);
#pragma unroll
        for (int k = 0; k < 4; ++k)
            tc::mma_ss(tmem + 256, tc::smem_desc_sw128(tc::smem_u32(sA) + k * 32),
                       tc::smem_desc_sw128(tc::smem_u32(sB1) + k * 32), idesc, k > 0);
        tc::commit(&bar_mma);
    }
    tc::mbar_wait(&bar_mma, 0);
    tc::fence_after();
    // acc0 -> global (fp32) and -> fp16 activation tile in TMEM columns [0, n1 / 2)
    const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
    for (int c0 = 0; c0 < n1; c0 += 8) {
        uint32_t r[8];
        tc::tmem_ld8(lane_base + 256 + c0, r);
        tc::tmem_ld_wait();
        uint32_t p[4];
#pragma unroll
        for (int q = 0; q < 8; ++q) out0[(size_t)tid * n1 + c0 + q] = __uint_as_float(r[q]);
#pragma unroll
        for (int q = 0; q < 4; ++q) p[q] = tc::pack_h2(__uint_as_float(r[2 * q]), __uint_as_float(r[2 * q + 1]));
        tc::tmem_st4(lane_base + c0 / 2, p);
    }
    tc::tmem_st_wait();
    tc::fence_before();
    __syncthreads();
    if (tid == 0) {
        tc::fence_after();
        const uint32_t idesc = tc::idesc_f16(n2);
        for (int k = 0; k < n1 / 16; ++k) {
            const int ks = k >> 2, kk = k & 3;
            tc::mma_ts(tmem + 256, tmem + k * 8,
                       tc::smem_desc_sw128(tc::smem_u32(sB2) + ks * n2 * 128 + kk * 32), idesc, k > 0);
        }
        tc::commit(&bar_mma);
    }
    tc::mbar_wait(&bar_mma, 1);
    tc::fence_after();
    for (int c0 = 0; c0 < n2; c0 += 8) {
        uint32_t r[8];
        tc::tmem_ld8(lane_base + 256 + c0, r);
        tc::tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 8; ++q) out1[(size_t)tid * n2 + c0 + q] = __uint_as_float(r[q]);
    }
    tc::fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<512>(tmem);
    (void)lane;
}

}  // namespace

// a: fp16 [rows][64] row-major in device memory (rows >= row0 + 128 is NOT required: rows past the end read as zero);
// b1: packed [n1 x 64], b2: packed [n2 x n1]; out0 fp32 [128][n1], out1 fp32 [128][n2].
extern "C" int fm_probe_umma(const void* a, int rows, int row0, const void* b1, const void* b2, int n1, int n2,
                             float* out0, float* out1, void* stream) {
    FM_REQUIRE(n1 % 16 == 0 && n1 >= 16 && n1 <= 128 && n2 % 16 == 0 && n2 >= 16 && n2 <= 256, "fm_probe_umma: n1/n2");
    CUtensorMap map;
    int rc = fm_make_tmap_f16_3d(&map, a, 64, (uint64_t)rows, 1, 64, (uint64_t)rows * 64, 64, 32, 1);
    if (rc) return rc;
    const int smem = 16384 + 128 * 128 + 2 * 256 * 128 + 1024;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    probe_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(map, row0, (const uint8_t*)b1, (const uint8_t*)b2, n1, n2, out0,
                                                        out1);
    FM_CHECK_LAUNCH("fm_probe_umma");
    return FM_OK;
}
