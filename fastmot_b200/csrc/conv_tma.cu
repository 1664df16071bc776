// Warp-specialised, TMA-fed tcgen05 convolution for sm_100a with the split-K reduction done inside a thread-block
// cluster (distributed shared memory) -- the batch-1 detector layers of the YOLO stack (role of the TensorRT conv
// tactics behind fastmot/utils/inference.py:106-117; layer semantics of scripts/yolo2onnx.py:558-700).
//
//   D[128 output pixels, BN filters] = sum over K slices of  A_slice[128 x 64] * W_slice[BN x 64]^T
//
// Supported layers: 1x1 and 3x3, stride 1, "same" padding, cin % 64 == 0, 8-channel aligned views (everything else
// stays on conv_tc.cu).  A K slice is (filter tap, 64 input channels).
//   * A operand: one TMA box per slice.  The 128 tile rows are a TW x TH rectangle of output pixels; for tap (r, s)
//     the box is the same rectangle shifted by (r - pad, s - pad) in a [C, W, H] tensor map, so the zero padding is the
//     TMA out-of-bounds fill and no thread ever computes an im2col address.  1x1 layers use the flattened
//     [C, N*H*W, 1] view (128 consecutive pixels).
//   * B operand: one TMA box [64 x BN] of the K-major weight matrix [cout][kh*kw*cin].
//   * both land 128-byte swizzled in an NS-stage mbarrier ring; one thread issues the copies and the four
//     tcgen05.mma (128 x BN x 16) per slice, accumulators in TMEM; weight boxes of the first stages are requested before
//     griddepcontrol.wait (they do not depend on the previous layer).
//   * split K: gridDim.z = S CTAs of one cluster share an output tile, each owns nk / S slices.  Partial tiles are
//     written as fp32 to the CTA's own (now idle) ring memory; after a cluster barrier CTA z sums rows z, z+S, ... of
//     all S partial tiles through ld.shared::cluster and finishes them (bias, activation, residual, fp16 NHWC store,
//     lanes along the channels).  No fp32 workspace in HBM, no second kernel.
//   * S == 1: the tile goes TMEM -> fp16 staging -> coalesced rows like conv_tc.cu.
#include "tc_common.cuh"
#include "conv_act.cuh"
#include <stdlib.h>
#include <string.h>

namespace {

using namespace tc;

struct ConvTmaArgs {
    int W, H, TW, TH, tiles_w;      // output plane, tile rectangle, tiles per plane row
    int kc, kw, pad;                // cin / 64, filter width, padding
    int nk, sps;                    // K slices in total / per cluster rank
    int cout, cout_stride, cout_offset, res_stride, res_offset, act;
    const float* bias;
    const __half* residual;
    __half* out;
};

__device__ __forceinline__ void cl_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cl_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t cl_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ float4 ld_cluster_f4(uint32_t local_saddr, uint32_t rank) {
    uint32_t raddr;
    float4 v;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(local_saddr), "r"(rank));
    asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "r"(raddr)
                 : "memory");
    return v;
}
__device__ __forceinline__ void bar_sync_128() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// bias + activation (+ residual) of 8 consecutive channels of one pixel, fp16 store
__device__ __forceinline__ void finish8(float (&x)[8], const ConvTmaArgs& a, size_t pix, int n, int act, bool res_first) {
    if (a.bias) {
        const float4 ba = __ldg(reinterpret_cast<const float4*>(a.bias + n));
        const float4 bb = __ldg(reinterpret_cast<const float4*>(a.bias + n + 4));
        x[0] += ba.x; x[1] += ba.y; x[2] += ba.z; x[3] += ba.w;
        x[4] += bb.x; x[5] += bb.y; x[6] += bb.z; x[7] += bb.w;
    }
    if (!res_first) tc_act8(x, act);
    if (a.residual) {
        const uint4 rv = *reinterpret_cast<const uint4*>(a.residual + pix * a.res_stride + a.res_offset + n);
        const __half2* rh = reinterpret_cast<const __half2*>(&rv);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float2 f = __half22float2(rh[q]);
            x[2 * q] += f.x;
            x[2 * q + 1] += f.y;
        }
    }
    if (res_first) tc_act8(x, act);
    *reinterpret_cast<uint4*>(a.out + pix * a.cout_stride + a.cout_offset + n) =
        make_uint4(pack_h2(x[0], x[1]), pack_h2(x[2], x[3]), pack_h2(x[4], x[5]), pack_h2(x[6], x[7]));
}

template <int BN, int NS>
__global__ void __launch_bounds__(160) conv_tma_kernel(const __grid_constant__ CUtensorMap map_a,
                                                        const __grid_constant__ CUtensorMap map_b, ConvTmaArgs a) {
    constexpr int A_BYTES = 16384, B_BYTES = BN * 128, STAGE = A_BYTES + B_BYTES;
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t full[NS], empty[NS], acc_full;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    fm_pdl_trigger();
    const int S = (int)gridDim.z;
    const int z = S > 1 ? (int)cl_rank() : 0;
    const int th = (int)blockIdx.x / a.tiles_w, tw = (int)blockIdx.x - th * a.tiles_w;
    const int w0 = tw * a.TW, h0 = th * a.TH;
    const int n0 = (int)blockIdx.y * BN;
    const int rows_used = a.TW * a.TH;
    const int k0 = z * a.sps;
    const int iters = min(a.nk - k0, a.sps);

    if (tid == 0) {
        if (smem_u32(smem) & 1023u) __trap();
#pragma unroll
        for (int i = 0; i < NS; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(&acc_full, 1);
        mbar_fence_init();
    }
    if (warp == 4) {
        tmem_alloc<BN>(&s_tmem);
        if (lane == 0) { tma_prefetch_desc(&map_a); tma_prefetch_desc(&map_b); }
    }
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = s_tmem;

    if (warp == 4) {
        // ------------------------------------------- control warp --------------------------------------------
        const bool leader = lane == 0;
        const uint32_t idesc = idesc_f16(BN);
        const uint32_t tx = (uint32_t)rows_used * 128u + (uint32_t)B_BYTES;
        auto load_b = [&](int i) {
            const int st = i % NS;
            mbar_expect_tx(&full[st], tx);
            tma_load_3d(smem + (size_t)st * STAGE + A_BYTES, &map_b, &full[st], (k0 + i) * 64, n0, 0);
        };
        auto load_a = [&](int i) {
            const int st = i % NS;
            const int g = k0 + i;
            const int tap = g / a.kc, c = g - tap * a.kc;
            const int fr = tap / a.kw, fs = tap - fr * a.kw;
            tma_load_3d(smem + (size_t)st * STAGE, &map_a, &full[st], c * 64, w0 + fs - a.pad, h0 + fr - a.pad);
        };
        const int pre = min(NS - 1, iters);
        if (leader)
            for (int i = 0; i < pre; ++i) load_b(i);           // weights: independent of the previous layer
        fm_pdl_wait();
        if (leader)
            for (int i = 0; i < pre; ++i) load_a(i);
        for (int i = 0; i < iters; ++i) {
            const int nx = i + NS - 1;
            if (nx < iters) {
                const int sn = nx % NS;
                if (nx >= NS) mbar_wait(&empty[sn], (uint32_t)((nx / NS - 1) & 1));
                if (leader) { load_b(nx); load_a(nx); }
            }
            const int st = i % NS;
            mbar_wait(&full[st], (uint32_t)((i / NS) & 1));
            fence_after();
            if (leader) {
                const uint32_t sa = smem_u32(smem + (size_t)st * STAGE), sb = sa + A_BYTES;
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    mma_ss(tmem, smem_desc_sw128(sa + k * 32), smem_desc_sw128(sb + k * 32), idesc, (i > 0 || k > 0) ? 1u : 0u);
                commit(&empty[st]);
                if (i == iters - 1) commit(&acc_full);
            }
            __syncwarp();
        }
    } else {
        // ------------------------------------------- epilogue warps ------------------------------------------
        fm_pdl_wait();
        mbar_wait_sleep(&acc_full, 0);
        fence_after();
        const int row = warp * 32 + lane;
        const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
        if (S == 1) {
            constexpr int PITCH = BN * 2 + 16;
#pragma unroll 1
            for (int j0 = 0; j0 < BN; j0 += 32) {
                uint32_t r[32];
                tmem_ld32(lane_base + j0, r);
                tmem_ld_wait();
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    *reinterpret_cast<uint4*>(smem + row * PITCH + (j0 + e * 8) * 2) = make_uint4(
                        pack_h2(__uint_as_float(r[e * 8 + 0]), __uint_as_float(r[e * 8 + 1])),
                        pack_h2(__uint_as_float(r[e * 8 + 2]), __uint_as_float(r[e * 8 + 3])),
                        pack_h2(__uint_as_float(r[e * 8 + 4]), __uint_as_float(r[e * 8 + 5])),
                        pack_h2(__uint_as_float(r[e * 8 + 6]), __uint_as_float(r[e * 8 + 7])));
            }
            fence_before();
            bar_sync_128();
            constexpr int CPR = BN / 8;
            const int act = a.act & 0xff;
            const bool res_first = (a.act & FM_ACT_AFTER_RESIDUAL) != 0;
            for (int i = tid; i < 128 * CPR; i += 128) {
                const int rr = i / CPR, ch = i - rr * CPR;
                const int n = n0 + ch * 8;
                const int hl = rr / a.TW, wl = rr - hl * a.TW;
                const int h = h0 + hl, w = w0 + wl;
                if (rr >= rows_used || h >= a.H || w >= a.W || n >= a.cout) continue;
                const uint4 pk = *reinterpret_cast<const uint4*>(smem + rr * PITCH + ch * 16);
                const __half2* ph = reinterpret_cast<const __half2*>(&pk);
                float x[8];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float2 f = __half22float2(ph[q]);
                    x[2 * q] = f.x;
                    x[2 * q + 1] = f.y;
                }
                finish8(x, a, (size_t)h * a.W + w, n, act, res_first);
            }
        } else {
            constexpr int PITCH = BN * 4 + 16;
#pragma unroll 1
            for (int j0 = 0; j0 < BN; j0 += 32) {
                uint32_t r[32];
                tmem_ld32(lane_base + j0, r);
                tmem_ld_wait();
#pragma unroll
                for (int q = 0; q < 32; q += 4)
                    *reinterpret_cast<uint4*>(smem + row * PITCH + (j0 + q) * 4) = make_uint4(r[q], r[q + 1], r[q + 2], r[q + 3]);
            }
            fence_before();
        }
    }
    if (S > 1) {
        // every thread of every CTA of the cluster: partial tiles are in place -> reduce -> nobody leaves early
        cl_arrive();
        cl_wait();
        if (warp < 4) {
            constexpr int PITCH = BN * 4 + 16;
            constexpr int CPR = BN / 8;
            const int act = a.act & 0xff;
            const bool res_first = (a.act & FM_ACT_AFTER_RESIDUAL) != 0;
            const int nrows = (128 - z + S - 1) / S;
            const uint32_t base = smem_u32(smem);
            for (int i = tid; i < nrows * CPR; i += 128) {
                const int ri = i / CPR, ch = i - ri * CPR;
                const int rr = z + ri * S;
                const int n = n0 + ch * 8;
                const int hl = rr / a.TW, wl = rr - hl * a.TW;
                const int h = h0 + hl, w = w0 + wl;
                if (rr >= rows_used || h >= a.H || w >= a.W || n >= a.cout) continue;
                const uint32_t off = base + (uint32_t)(rr * PITCH + ch * 32);
                float x[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) x[q] = 0.f;
                for (int p = 0; p < S; ++p) {
                    const float4 v0 = ld_cluster_f4(off, (uint32_t)p);
                    const float4 v1 = ld_cluster_f4(off + 16, (uint32_t)p);
                    x[0] += v0.x; x[1] += v0.y; x[2] += v0.z; x[3] += v0.w;
                    x[4] += v1.x; x[5] += v1.y; x[6] += v1.z; x[7] += v1.w;
                }
                finish8(x, a, (size_t)h * a.W + w, n, act, res_first);
            }
        }
        cl_arrive();
        cl_wait();
    }
    fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc<BN>(tmem);
}

struct Plan {
    int W, H, TW, TH, tiles_w, tiles;
};

Plan plan_tiles(const FmConvDesc* d) {
    Plan p;
    if (d->kh == 1) {
        p.W = d->n * d->ho * d->wo; p.H = 1; p.TW = 128; p.TH = 1;
        p.tiles_w = fm_cdiv(p.W, 128); p.tiles = p.tiles_w;
        return p;
    }
    p.W = d->wo; p.H = d->ho;
    int best = 1 << 30, btw = 8;
    for (int tw = 4; tw <= 128 && tw <= p.W; ++tw) {
        const int th = 128 / tw;
        if (th < 1) break;
        const int t = fm_cdiv(p.W, tw) * fm_cdiv(p.H, th);
        // fewer tiles first; then the wider rectangle (longer contiguous runs for the TMA box and the stores)
        if (t < best || (t == best && tw > btw)) { best = t; btw = tw; }
    }
    p.TW = btw; p.TH = 128 / btw;
    p.tiles_w = fm_cdiv(p.W, p.TW);
    p.tiles = best;
    return p;
}

template <int BN, int NS>
int max_clusters(int s) {
    static int cache[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (cache[s]) return cache[s];
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(1, 1, s);
    cfg.blockDim = dim3(160);
    cfg.dynamicSmemBytes = NS * (16384 + BN * 128);
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = s;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, conv_tma_kernel<BN, NS>, &cfg) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        n = s == 1 ? 2 * FM_NUM_SMS : -1;
    }
    cache[s] = n;
    return n;
}

template <int BN, int NS>
int launch_tma(const FmConvDesc* d, const Plan& p, const void* in, const void* wgt, const float* bias, const void* residual,
               void* out, cudaStream_t st) {
    constexpr int SMEM = NS * (16384 + BN * 128);
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(conv_tma_kernel<BN, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
        attr = true;
    }
    const int kc = d->cin / 64, taps = d->kh * d->kw, nk = taps * kc;
    const int ncol = fm_cdiv(d->cout, BN);
    const int tiles = p.tiles * ncol;
    static int smax_env = -1;            // FM_CONV_TMA_SPLIT=<n>: upper bound of the cluster size (1 = never split K)
    if (smax_env < 0) { const char* e = getenv("FM_CONV_TMA_SPLIT"); smax_env = e ? atoi(e) : 8; if (smax_env < 1) smax_env = 1; if (smax_env > 8) smax_env = 8; }
    int S = 1;
    if (tiles <= 100) {
        int smax = nk / 2 < smax_env ? nk / 2 : smax_env;
        for (int s = smax; s >= 2; --s) {
            const int cap = max_clusters<BN, NS>(s);
            if (cap >= tiles) { S = s; break; }
        }
    }
    int sps = fm_cdiv(nk, S);
    S = fm_cdiv(nk, sps);
    ConvTmaArgs a;
    a.W = p.W; a.H = p.H; a.TW = p.TW; a.TH = p.TH; a.tiles_w = p.tiles_w;
    a.kc = kc; a.kw = d->kw; a.pad = d->pad; a.nk = nk; a.sps = sps;
    a.cout = d->cout; a.cout_stride = d->cout_stride; a.cout_offset = d->cout_offset;
    a.res_stride = d->res_stride; a.res_offset = d->res_offset; a.act = d->act;
    a.bias = bias; a.residual = (const __half*)residual; a.out = (__half*)out;
    CUtensorMap map_a, map_b;
    const __half* base = (const __half*)in + d->cin_offset;
    int rc;
    if (d->kh == 1)
        rc = fm_make_tmap_f16_3d(&map_a, base, (uint64_t)d->cin, (uint64_t)p.W, 1, (uint64_t)d->cin_stride,
                                 (uint64_t)p.W * d->cin_stride, 64, 128, 1);
    else
        rc = fm_make_tmap_f16_3d(&map_a, base, (uint64_t)d->cin, (uint64_t)p.W, (uint64_t)p.H, (uint64_t)d->cin_stride,
                                 (uint64_t)p.W * d->cin_stride, 64, (uint32_t)p.TW, (uint32_t)p.TH);
    if (rc) return rc;
    const uint64_t ktot = (uint64_t)taps * d->cin;
    rc = fm_make_tmap_f16_3d(&map_b, wgt, ktot, (uint64_t)d->cout, 1, ktot, ktot * d->cout, 64, BN, 1);
    if (rc) return rc;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(p.tiles, ncol, S);
    cfg.blockDim = dim3(160);
    cfg.dynamicSmemBytes = SMEM;
    cfg.stream = st;
    cudaLaunchAttribute attrs[2];
    int na = 0;
    if (fm_pdl_enabled()) {
        attrs[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attrs[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    if (S > 1) {
        attrs[na].id = cudaLaunchAttributeClusterDimension;
        attrs[na].val.clusterDim.x = 1; attrs[na].val.clusterDim.y = 1; attrs[na].val.clusterDim.z = S;
        ++na;
    }
    cfg.attrs = attrs;
    cfg.numAttrs = na;
    cudaError_t e = cudaLaunchKernelEx(&cfg, conv_tma_kernel<BN, NS>, map_a, map_b, a);
    if (e != cudaSuccess) { fm_set_last_error(cudaGetErrorString(e)); return FM_ERR_CUDA; }
    return FM_OK;
}

}  // namespace

extern "C" int fm_conv2d_tma_supported(const FmConvDesc* d) {
    if (!d) return 0;
    if (d->kh != d->kw || (d->kh != 1 && d->kh != 3)) return 0;
    if (d->stride != 1 || d->pad != d->kh / 2 || d->hi != d->ho || d->wi != d->wo) return 0;
    if (d->cin < 64 || d->cin % 64 || d->cin_stride % 8 || d->cin_offset % 8) return 0;
    if (d->cout < 32 || d->cout % 8 || d->cout_stride % 8 || d->cout_offset % 8) return 0;
    if (d->res_stride % 8 || d->res_offset % 8) return 0;
    if (d->kh == 3 && (d->n != 1 || d->wo < 4)) return 0;       // one image per [C, W, H] tensor map
    if ((long long)d->n * d->ho * d->wo <= 0) return 0;
    return 1;
}

extern "C" int fm_conv2d_tma(const FmConvDesc* d, const void* in, const void* wgt, const float* bias, const void* residual,
                             void* out, void* stream) {
    FM_REQUIRE(d != nullptr, "fm_conv2d_tma: desc is NULL");
    FM_REQUIRE(fm_conv2d_tma_supported(d), "fm_conv2d_tma: layer not supported by the TMA path (1x1 / 3x3, stride 1, "
                                           "same padding, cin % 64 == 0, 8-channel aligned views)");
    FM_REQUIRE((((uintptr_t)in | (uintptr_t)wgt | (uintptr_t)out | (uintptr_t)residual) & 15) == 0,
               "fm_conv2d_tma: tensors must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const Plan p = plan_tiles(d);
    const int nk = d->kh * d->kw * (d->cin / 64);
    // 64-wide filter tiles when 128-wide ones (with the deepest K split the layer allows) would leave half the GPU idle
    static int force_bn = -1;            // FM_CONV_TMA_BN=64|128 (experiments)
    if (force_bn < 0) { const char* e = getenv("FM_CONV_TMA_BN"); force_bn = e ? atoi(e) : 0; }
    int bn = d->cout >= 128 ? 128 : 64;
    if (bn == 128) {
        const int smax = nk / 2 < 8 ? (nk / 2 < 1 ? 1 : nk / 2) : 8;
        if ((long long)p.tiles * fm_cdiv(d->cout, 128) * smax <= FM_NUM_SMS / 2) bn = 64;
    }
    if (force_bn == 64 || force_bn == 128) bn = force_bn;
    int rc = bn == 128 ? launch_tma<128, 3>(d, p, in, wgt, bias, residual, out, st)
                       : launch_tma<64, 4>(d, p, in, wgt, bias, residual, out, st);
    if (rc) return rc;
    FM_CHECK_LAUNCH("fm_conv2d_tma");
    return FM_OK;
}
