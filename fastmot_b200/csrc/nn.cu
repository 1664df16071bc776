// Neural-network layer kernels for the detector / ReID conv stacks, NHWC fp16 activations, fp32 accumulation.
// This file holds the generic SIMT implicit-GEMM convolution (any kernel size / stride / channel count; used for
// stems and as the on-device cross-check of the tcgen05 path in conv_tc.cu) and the bandwidth-bound layers:
// max/avg pooling, nearest upsample, channel-slice copy (route), residual add, depthwise 3x3, global average pool,
// OSNet channel gate, fully-connected head.
//
// Layer semantics follow the Darknet->ONNX converter the reference ships (scripts/yolo2onnx.py:558-870): conv
// `SAME_LOWER`-style symmetric padding (pad = k/2), BN folded into weight+bias, leaky 0.1, mish, swish, logistic,
// maxpool SAME_UPPER, route concat / channel-group split, nearest upsample x2; OSNet ops per SURVEY.md Appendix D.
#include "common.cuh"
#include "../../include/fastmot_b200.h"

// 16-byte vectorised fast paths (nn_vec.cu); each returns 1 if it took the call
int fm_vec_dwconv3(const void*, const void*, const float*, void*, int, int, int, int, int, cudaStream_t);
int fm_vec_add_act(const void*, const void*, void*, long long, int, cudaStream_t);
int fm_vec_add_act_strided(const void*, int, int, const void*, int, int, void*, int, int, long long, int, int, cudaStream_t);
int fm_vec_avgpool2(const void*, void*, int, int, int, int, cudaStream_t);
int fm_vec_maxpool(const void*, void*, int, int, int, int, int, int, int, int, int, int, int, int, int, int, cudaStream_t);
int fm_vec_upsample_copy(const void*, void*, int, int, int, int, int, int, int, int, int, cudaStream_t);
int fm_vec_gap(const void*, float*, int, int, int, cudaStream_t);
int fm_vec_gate_apply(const void*, const float*, void*, size_t, int, size_t, int, cudaStream_t);

namespace {

__device__ __forceinline__ float apply_act(float v, int act) {
    switch (act) {
        case FM_ACT_LEAKY: return v > 0.f ? v : 0.1f * v;
        case FM_ACT_RELU: return v > 0.f ? v : 0.f;
        case FM_ACT_MISH: {
            // x * tanh(softplus(x)); softplus with the usual overflow guard
            float sp = v > 20.f ? v : log1pf(__expf(v));
            return v * tanhf(sp);
        }
        case FM_ACT_SWISH: return v / (1.f + __expf(-v));
        case FM_ACT_LOGISTIC: return 1.f / (1.f + __expf(-v));
        default: return v;
    }
}

// ---------------------------------------------------------------------------------------------------------
// Generic implicit-GEMM conv: M = N*Ho*Wo pixels, Ncol = Cout, K = kh*kw*Cin.  64x64 tile, 16-deep K slices.
// ---------------------------------------------------------------------------------------------------------
#define CT_M 64
#define CT_N 64
#define CT_K 16
__global__ void __launch_bounds__(256) conv_simt_kernel(FmConvDesc d, const __half* __restrict__ in,
                                                         const __half* __restrict__ wgt,
                                                         const float* __restrict__ bias,
                                                         const __half* __restrict__ residual,
                                                         __half* __restrict__ out) {
    __shared__ __half sA[CT_K][CT_M + 2];
    __shared__ __half sB[CT_K][CT_N + 2];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.x * CT_M, n0 = blockIdx.y * CT_N;
    const int M = d.n * d.ho * d.wo;
    const int K = d.kh * d.kw * d.cin;
    const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads, 4x4 micro-tile each
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int k0 = 0; k0 < K; k0 += CT_K) {
        // A tile: CT_M pixels x CT_K reduction indices (gather with padding)
        for (int e = tid; e < CT_M * CT_K; e += 256) {
            const int kk = e % CT_K, mm = e / CT_K;
            const int m = m0 + mm, k = k0 + kk;
            __half v = __float2half(0.f);
            if (m < M && k < K) {
                const int c = k % d.cin, rs = k / d.cin, s = rs % d.kw, r = rs / d.kw;
                const int wo = m % d.wo, t = m / d.wo, ho = t % d.ho, nb = t / d.ho;
                const int hi = ho * d.stride - d.pad + r, wi = wo * d.stride - d.pad + s;
                if (hi >= 0 && hi < d.hi && wi >= 0 && wi < d.wi)
                    v = in[(((size_t)nb * d.hi + hi) * d.wi + wi) * d.cin_stride + d.cin_offset + c];
            }
            sA[kk][mm] = v;
        }
        for (int e = tid; e < CT_N * CT_K; e += 256) {
            const int kk = e % CT_K, nn = e / CT_K;
            const int n = n0 + nn, k = k0 + kk;
            sB[kk][nn] = (n < d.cout && k < K) ? wgt[(size_t)n * K + k] : __float2half(0.f);
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < CT_K; ++kk) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = __half2float(sA[kk][ty * 4 + i]);
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = __half2float(sB[kk][tx * 4 + j]);
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] += a[i] * b[j];
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= d.cout) continue;
            float v = acc[i][j] + (bias ? bias[n] : 0.f);
            const bool res_first = (d.act & FM_ACT_AFTER_RESIDUAL) != 0;
            if (!res_first) v = apply_act(v, d.act & 0xff);
            if (residual) v += __half2float(residual[(size_t)m * d.res_stride + d.res_offset + n]);
            if (res_first) v = apply_act(v, d.act & 0xff);
            out[(size_t)m * d.cout_stride + d.cout_offset + n] = __float2half(v);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Bandwidth-bound layers.  All take channel strides/offsets so routes (concat / group split) need no copies.
// ---------------------------------------------------------------------------------------------------------
// maxpool, Darknet/ONNX SAME_UPPER: out = ceil(in / stride), pad_total = (out-1)*stride + k - in, pad_lo = total/2
__global__ void maxpool_kernel(const __half* __restrict__ in, __half* __restrict__ out, int n, int hi, int wi, int c,
                               int cin_stride, int cin_off, int ho, int wo, int cout_stride, int cout_off, int k,
                               int stride, int pad_lo_h, int pad_lo_w) {
    const size_t total = (size_t)n * ho * wo * c;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int ch = idx % c;
        size_t t = idx / c;
        const int x = t % wo; t /= wo;
        const int y = t % ho;
        const int b = t / ho;
        float best = -INFINITY;
        for (int r = 0; r < k; ++r) {
            const int yy = y * stride - pad_lo_h + r;
            if (yy < 0 || yy >= hi) continue;
            for (int s = 0; s < k; ++s) {
                const int xx = x * stride - pad_lo_w + s;
                if (xx < 0 || xx >= wi) continue;
                best = fmaxf(best, __half2float(in[(((size_t)b * hi + yy) * wi + xx) * cin_stride + cin_off + ch]));
            }
        }
        out[(((size_t)b * ho + y) * wo + x) * cout_stride + cout_off + ch] = __float2half(best);
    }
}

__global__ void avgpool2_kernel(const __half* __restrict__ in, __half* __restrict__ out, int n, int hi, int wi, int c) {
    const int ho = hi / 2, wo = wi / 2;
    const size_t total = (size_t)n * ho * wo * c;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int ch = idx % c;
        size_t t = idx / c;
        const int x = t % wo; t /= wo;
        const int y = t % ho;
        const int b = t / ho;
        const __half* p = in + (((size_t)b * hi + 2 * y) * wi + 2 * x) * c + ch;
        const float v = __half2float(p[0]) + __half2float(p[c]) + __half2float(p[(size_t)wi * c]) +
                        __half2float(p[(size_t)wi * c + c]);
        out[idx] = __float2half(0.25f * v);
    }
}

// nearest upsample x`s` and/or channel-slice copy: out[b, y, x, cout_off + ch] = in[b, y/s, x/s, cin_off + ch]
__global__ void upsample_copy_kernel(const __half* __restrict__ in, __half* __restrict__ out, int n, int hi, int wi,
                                     int c, int cin_stride, int cin_off, int s, int cout_stride, int cout_off) {
    const int ho = hi * s, wo = wi * s;
    const size_t total = (size_t)n * ho * wo * c;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int ch = idx % c;
        size_t t = idx / c;
        const int x = t % wo; t /= wo;
        const int y = t % ho;
        const int b = t / ho;
        out[(((size_t)b * ho + y) * wo + x) * cout_stride + cout_off + ch] =
            in[(((size_t)b * hi + y / s) * wi + x / s) * cin_stride + cin_off + ch];
    }
}

// out = act(a + b)   (Darknet shortcut; OSNet residual + ReLU)
__global__ void add_act_kernel(const __half* __restrict__ a, const __half* __restrict__ b, __half* __restrict__ out,
                               size_t n, int act) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        out[i] = __float2half(apply_act(__half2float(a[i]) + __half2float(b[i]), act));
}

// depthwise 3x3, stride 1, pad 1, + bias (folded BN) + activation
__global__ void dwconv3_kernel(const __half* __restrict__ in, const __half* __restrict__ w /* [9][c] */,
                               const float* __restrict__ bias, __half* __restrict__ out, int n, int h, int wd, int c,
                               int act) {
    const size_t total = (size_t)n * h * wd * c;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int ch = idx % c;
        size_t t = idx / c;
        const int x = t % wd; t /= wd;
        const int y = t % h;
        const int b = t / h;
        float acc = bias ? bias[ch] : 0.f;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const int yy = y + r - 1;
            if (yy < 0 || yy >= h) continue;
#pragma unroll
            for (int s = 0; s < 3; ++s) {
                const int xx = x + s - 1;
                if (xx < 0 || xx >= wd) continue;
                acc += __half2float(in[(((size_t)b * h + yy) * wd + xx) * c + ch]) * __half2float(w[(r * 3 + s) * c + ch]);
            }
        }
        out[idx] = __float2half(apply_act(acc, act));
    }
}

// global average pool: out[b][c] (fp32) = mean over h*w
__global__ void __launch_bounds__(256) gap_kernel(const __half* __restrict__ in, float* __restrict__ out, int hw, int c) {
    const int b = blockIdx.x;
    for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
        float acc = 0.f;
        const __half* p = in + (size_t)b * hw * c + ch;
        for (int i = 0; i < hw; ++i) acc += __half2float(p[(size_t)i * c]);
        out[(size_t)b * c + ch] = acc / hw;
    }
}

// OSNet ChannelGate on pooled features: g = sigmoid(W2 relu(W1 p + b1) + b2); one CTA per sample.
__global__ void __launch_bounds__(128) gate_fc_kernel(const float* __restrict__ pooled, const float* __restrict__ w1,
                                                       const float* __restrict__ b1, const float* __restrict__ w2,
                                                       const float* __restrict__ b2, float* __restrict__ gate, int c,
                                                       int cr) {
    extern __shared__ float sh[];  // c + cr
    float* sp = sh;
    float* sh1 = sh + c;
    const int b = blockIdx.x;
    for (int i = threadIdx.x; i < c; i += blockDim.x) sp[i] = pooled[(size_t)b * c + i];
    __syncthreads();
    for (int j = threadIdx.x; j < cr; j += blockDim.x) {
        float a = b1[j];
        for (int i = 0; i < c; ++i) a += w1[(size_t)j * c + i] * sp[i];
        sh1[j] = a > 0.f ? a : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < c; i += blockDim.x) {
        float a = b2[i];
        for (int j = 0; j < cr; ++j) a += w2[(size_t)i * cr + j] * sh1[j];
        gate[(size_t)b * c + i] = 1.f / (1.f + __expf(-a));
    }
}

// acc (+)= gate[b][c] * x[b, :, :, c]
__global__ void gate_apply_kernel(const __half* __restrict__ x, const float* __restrict__ gate, __half* __restrict__ acc,
                                  size_t per_sample, int c, size_t total, int accumulate) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int ch = i % c;
        const size_t b = i / per_sample;
        float v = __half2float(x[i]) * gate[b * c + ch];
        if (accumulate) v += __half2float(acc[i]);
        acc[i] = __float2half(v);
    }
}

// fully connected + folded BN + ReLU, then row L2 normalisation (feature_extractor.py:73): one CTA per sample
__global__ void __launch_bounds__(256) fc_norm_kernel(const float* __restrict__ in, const float* __restrict__ w,
                                                       const float* __restrict__ bias, float* __restrict__ out, int cin,
                                                       int cout, int relu, int normalize) {
    extern __shared__ float sin_[];  // cin + 8
    __shared__ float s_part[8];
    const int b = blockIdx.x;
    for (int i = threadIdx.x; i < cin; i += blockDim.x) sin_[i] = in[(size_t)b * cin + i];
    __syncthreads();
    float sq = 0.f;
    for (int j = threadIdx.x; j < cout; j += blockDim.x) {
        float a = bias ? bias[j] : 0.f;
        const float* wr = w + (size_t)j * cin;
        for (int i = 0; i < cin; ++i) a += wr[i] * sin_[i];
        if (relu) a = a > 0.f ? a : 0.f;
        out[(size_t)b * cout + j] = a;
        sq += a * a;
    }
    if (!normalize) return;
    sq = warp_sum(sq);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = sq;
    __syncthreads();
    float tot = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) tot += s_part[i];
    const float inv = 1.f / sqrtf(tot);
    for (int j = threadIdx.x; j < cout; j += blockDim.x) out[(size_t)b * cout + j] *= inv;
}

// Same op, S samples per CTA and one warp per output feature: lanes stride the weight row with 128-bit loads
// (coalesced; the row-per-thread kernel above reads 32 different rows per load instruction), the weight row is
// reused for the S samples, and the L2 norm is taken from shared memory before a single coalesced store.
template <int S>
__global__ void __launch_bounds__(256) fc_norm_warp_kernel(const float* __restrict__ in, const float* __restrict__ w,
                                                            const float* __restrict__ bias, float* __restrict__ out,
                                                            int n, int cin, int cout, int relu, int normalize) {
    extern __shared__ float fsm[];   // x[S][cin] | y[S][cout]
    __shared__ float s_inv[S];
    float* sx = fsm;
    float* sy = fsm + (size_t)S * cin;
    const int b0 = blockIdx.x * S;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < S * cin; i += blockDim.x) {
        const int s = i / cin, b = b0 + s;
        sx[i] = b < n ? in[(size_t)b * cin + (i - s * cin)] : 0.f;
    }
    __syncthreads();
    const int c4 = cin >> 2;
    for (int j = warp; j < cout; j += 8) {
        const float4* wr = reinterpret_cast<const float4*>(w + (size_t)j * cin);
        float acc[S];
#pragma unroll
        for (int s = 0; s < S; ++s) acc[s] = 0.f;
        for (int i = lane; i < c4; i += 32) {
            const float4 wv = wr[i];
#pragma unroll
            for (int s = 0; s < S; ++s) {
                const float4 xv = reinterpret_cast<const float4*>(sx + (size_t)s * cin)[i];
                acc[s] += wv.x * xv.x + wv.y * xv.y + wv.z * xv.z + wv.w * xv.w;
            }
        }
#pragma unroll
        for (int s = 0; s < S; ++s) {
            float a = warp_sum(acc[s]);
            if (lane == 0) {
                a += bias ? bias[j] : 0.f;
                if (relu) a = fmaxf(a, 0.f);
                sy[(size_t)s * cout + j] = a;
            }
        }
    }
    __syncthreads();
    if (warp < S) {
        float sq = 0.f;
        for (int j = lane; j < cout; j += 32) { const float a = sy[(size_t)warp * cout + j]; sq += a * a; }
        sq = warp_sum(sq);
        if (lane == 0) s_inv[warp] = normalize ? 1.f / sqrtf(sq) : 1.f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < S * cout; i += blockDim.x) {
        const int s = i / cout, b = b0 + s;
        if (b < n) out[(size_t)b * cout + (i - s * cout)] = sy[i] * s_inv[s];
    }
}

// Same op for the ReID head (200 x 512 -> 512, feature_extractor.py:62-74) spread over the whole GPU: a cluster of 8 CTAs
// shares S samples, CTA r computes output features [r * cout / 8, (r + 1) * cout / 8) (a warp per feature, lanes along
// the weight row), the per-sample sums of squares are exchanged through distributed shared memory and added in rank
// order (deterministic), and every CTA normalises and stores its own slice.  The two-samples-per-CTA kernel above made
// 100 CTAs each stream the whole 1 MB weight matrix (76 us); this one reads it 25 times with 200 CTAs.
template <int S>
__global__ void __cluster_dims__(8, 1, 1) __launch_bounds__(256)
fc_norm_cluster_kernel(const float* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias,
                       float* __restrict__ out, int n, int cin, int cout, int relu, int normalize) {
    extern __shared__ float fsm[];   // x[S][cin]
    __shared__ float sy[S][128];     // this CTA's slice of the outputs (cout / 8 <= 128)
    __shared__ float s_ssq[S], s_inv[S];
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const int ob = cout >> 3, j0 = (int)rank * ob;
    const int b0 = (int)(blockIdx.x >> 3) * S;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < S * cin; i += blockDim.x) {
        const int s = i / cin, b = b0 + s;
        fsm[i] = b < n ? in[(size_t)b * cin + (i - s * cin)] : 0.f;
    }
    __syncthreads();
    const int c4 = cin >> 2;
    for (int jj = warp; jj < ob; jj += 8) {
        const float4* wr = reinterpret_cast<const float4*>(w + (size_t)(j0 + jj) * cin);
        float acc[S];
#pragma unroll
        for (int s = 0; s < S; ++s) acc[s] = 0.f;
        for (int i = lane; i < c4; i += 32) {
            const float4 wv = __ldg(wr + i);
#pragma unroll
            for (int s = 0; s < S; ++s) {
                const float4 xv = reinterpret_cast<const float4*>(fsm + (size_t)s * cin)[i];
                acc[s] += wv.x * xv.x + wv.y * xv.y + wv.z * xv.z + wv.w * xv.w;
            }
        }
#pragma unroll
        for (int s = 0; s < S; ++s) {
            float a = warp_sum(acc[s]);
            if (lane == 0) {
                a += bias ? bias[j0 + jj] : 0.f;
                if (relu) a = fmaxf(a, 0.f);
                sy[s][jj] = a;
            }
        }
    }
    __syncthreads();
    if (warp < S) {
        float sq = 0.f;
        for (int jj = lane; jj < ob; jj += 32) { const float a = sy[warp][jj]; sq += a * a; }
        sq = warp_sum(sq);
        if (lane == 0) s_ssq[warp] = sq;
    }
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (threadIdx.x < S) {
        float tot = 0.f;
        const uint32_t laddr = (uint32_t)__cvta_generic_to_shared(&s_ssq[threadIdx.x]);
#pragma unroll
        for (uint32_t r = 0; r < 8; ++r) {
            uint32_t raddr;
            float v;
            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(laddr), "r"(r));
            asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(raddr));
            tot += v;
        }
        s_inv[threadIdx.x] = normalize ? 1.f / sqrtf(tot) : 1.f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < S * ob; i += blockDim.x) {
        const int s = i / ob, jj = i - s * ob, b = b0 + s;
        if (b < n) out[(size_t)b * cout + j0 + jj] = sy[s][jj] * s_inv[s];
    }
    // nobody leaves while a peer may still read its sums
    asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// strided variant: operands are channel slices of wider NHWC buffers
__global__ void add_act_strided_kernel(const __half* __restrict__ a, int a_stride, int a_off,
                                       const __half* __restrict__ b, int b_stride, int b_off, __half* __restrict__ out,
                                       int o_stride, int o_off, size_t pixels, int c, int act) {
    const size_t total = pixels * c;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t p = i / c;
        const int ch = i - p * c;
        const float v = __half2float(a[p * a_stride + a_off + ch]) + __half2float(b[p * b_stride + b_off + ch]);
        out[p * o_stride + o_off + ch] = __float2half(apply_act(v, act));
    }
}

inline int grid_for(size_t total, int block = 256) {
    size_t g = (total + block - 1) / block;
    size_t cap = (size_t)FM_NUM_SMS * 16;
    return (int)(g < cap ? (g ? g : 1) : cap);
}

}  // namespace

extern "C" int fm_conv2d_simt(const FmConvDesc* d, const void* in, const void* wgt, const float* bias,
                              const void* residual, void* out, void* stream) {
    FM_REQUIRE(d != nullptr, "fm_conv2d_simt: desc is NULL");
    const int M = d->n * d->ho * d->wo;
    if (M <= 0) return FM_OK;
    dim3 grid(fm_cdiv(M, CT_M), fm_cdiv(d->cout, CT_N));
    conv_simt_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(*d, (const __half*)in, (const __half*)wgt, bias,
                                                             (const __half*)residual, (__half*)out);
    FM_CHECK_LAUNCH("fm_conv2d_simt");
    return FM_OK;
}

extern "C" int fm_maxpool(const void* in, void* out, int n, int hi, int wi, int c, int cin_stride, int cin_off, int k,
                          int stride, int cout_stride, int cout_off, void* stream) {
    const int ho = (hi + stride - 1) / stride, wo = (wi + stride - 1) / stride;
    const int ph = max((ho - 1) * stride + k - hi, 0), pw = max((wo - 1) * stride + k - wi, 0);
    const size_t total = (size_t)n * ho * wo * c;
    if (!total) return FM_OK;
    if (fm_vec_maxpool(in, out, n, hi, wi, c, cin_stride, cin_off, ho, wo, cout_stride, cout_off, k, stride, ph / 2,
                       pw / 2, (cudaStream_t)stream)) {
        FM_CHECK_LAUNCH("fm_maxpool");
        return FM_OK;
    }
    maxpool_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>((const __half*)in, (__half*)out, n, hi, wi, c,
                                                                      cin_stride, cin_off, ho, wo, cout_stride,
                                                                      cout_off, k, stride, ph / 2, pw / 2);
    FM_CHECK_LAUNCH("fm_maxpool");
    return FM_OK;
}

extern "C" int fm_maxpool_pad(const void* in, void* out, int n, int hi, int wi, int c, int k, int stride, int pad,
                              void* stream) {
    // PyTorch-style MaxPool2d(k, stride, padding=pad), floor mode (OSNet stem)
    const int ho = (hi + 2 * pad - k) / stride + 1, wo = (wi + 2 * pad - k) / stride + 1;
    const size_t total = (size_t)n * ho * wo * c;
    if (!total) return FM_OK;
    if (fm_vec_maxpool(in, out, n, hi, wi, c, c, 0, ho, wo, c, 0, k, stride, pad, pad, (cudaStream_t)stream)) {
        FM_CHECK_LAUNCH("fm_maxpool_pad");
        return FM_OK;
    }
    maxpool_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>((const __half*)in, (__half*)out, n, hi, wi, c, c,
                                                                      0, ho, wo, c, 0, k, stride, pad, pad);
    FM_CHECK_LAUNCH("fm_maxpool_pad");
    return FM_OK;
}

extern "C" int fm_avgpool2(const void* in, void* out, int n, int hi, int wi, int c, void* stream) {
    const size_t total = (size_t)n * (hi / 2) * (wi / 2) * c;
    if (!total) return FM_OK;
    if (fm_vec_avgpool2(in, out, n, hi, wi, c, (cudaStream_t)stream)) { FM_CHECK_LAUNCH("fm_avgpool2"); return FM_OK; }
    avgpool2_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>((const __half*)in, (__half*)out, n, hi, wi, c);
    FM_CHECK_LAUNCH("fm_avgpool2");
    return FM_OK;
}

extern "C" int fm_upsample_copy(const void* in, void* out, int n, int hi, int wi, int c, int cin_stride, int cin_off,
                                int scale, int cout_stride, int cout_off, void* stream) {
    const size_t total = (size_t)n * hi * scale * wi * scale * c;
    if (!total) return FM_OK;
    if (fm_vec_upsample_copy(in, out, n, hi, wi, c, cin_stride, cin_off, scale, cout_stride, cout_off,
                             (cudaStream_t)stream)) {
        FM_CHECK_LAUNCH("fm_upsample_copy");
        return FM_OK;
    }
    upsample_copy_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>((const __half*)in, (__half*)out, n, hi, wi,
                                                                            c, cin_stride, cin_off, scale, cout_stride,
                                                                            cout_off);
    FM_CHECK_LAUNCH("fm_upsample_copy");
    return FM_OK;
}

extern "C" int fm_add_act(const void* a, const void* b, void* out, long long n, int act, void* stream) {
    if (n <= 0) return FM_OK;
    if (fm_vec_add_act(a, b, out, n, act, (cudaStream_t)stream)) { FM_CHECK_LAUNCH("fm_add_act"); return FM_OK; }
    add_act_kernel<<<grid_for((size_t)n), 256, 0, (cudaStream_t)stream>>>((const __half*)a, (const __half*)b,
                                                                          (__half*)out, (size_t)n, act);
    FM_CHECK_LAUNCH("fm_add_act");
    return FM_OK;
}

extern "C" int fm_dwconv3(const void* in, const void* w, const float* bias, void* out, int n, int h, int wd, int c,
                          int act, void* stream) {
    const size_t total = (size_t)n * h * wd * c;
    if (!total) return FM_OK;
    if (fm_vec_dwconv3(in, w, bias, out, n, h, wd, c, act, (cudaStream_t)stream)) { FM_CHECK_LAUNCH("fm_dwconv3"); return FM_OK; }
    dwconv3_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>((const __half*)in, (const __half*)w, bias,
                                                                      (__half*)out, n, h, wd, c, act);
    FM_CHECK_LAUNCH("fm_dwconv3");
    return FM_OK;
}

extern "C" int fm_global_avgpool(const void* in, float* out, int n, int hw, int c, void* stream) {
    if (n <= 0) return FM_OK;
    if (fm_vec_gap(in, out, n, hw, c, (cudaStream_t)stream)) { FM_CHECK_LAUNCH("fm_global_avgpool"); return FM_OK; }
    gap_kernel<<<n, 256, 0, (cudaStream_t)stream>>>((const __half*)in, out, hw, c);
    FM_CHECK_LAUNCH("fm_global_avgpool");
    return FM_OK;
}

extern "C" int fm_channel_gate(const void* x, float* pooled, float* gate, const float* w1, const float* b1,
                               const float* w2, const float* b2, void* acc, int n, int hw, int c, int cr,
                               int accumulate, void* stream) {
    if (n <= 0) return FM_OK;
    cudaStream_t s = (cudaStream_t)stream;
    if (!fm_vec_gap(x, pooled, n, hw, c, s)) gap_kernel<<<n, 256, 0, s>>>((const __half*)x, pooled, hw, c);
    gate_fc_kernel<<<n, 128, (c + cr) * sizeof(float), s>>>(pooled, w1, b1, w2, b2, gate, c, cr);
    const size_t total = (size_t)n * hw * c;
    if (!fm_vec_gate_apply(x, gate, acc, (size_t)hw * c, c, total, accumulate, s))
        gate_apply_kernel<<<grid_for(total), 256, 0, s>>>((const __half*)x, gate, (__half*)acc, (size_t)hw * c, c, total,
                                                      accumulate);
    FM_CHECK_LAUNCH("fm_channel_gate");
    return FM_OK;
}

extern "C" int fm_fc_norm(const float* in, const float* w, const float* bias, float* out, int n, int cin, int cout,
                          int relu, int normalize, void* stream) {
    if (n <= 0) return FM_OK;
    if ((cin & 3) == 0 && cin <= 2048 && (cout & 7) == 0 && cout <= 1024 && n >= 16) {
        constexpr int S = 8;
        static bool attr = false;
        if (!attr) {
            cudaFuncSetAttribute(fc_norm_cluster_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, S * 2048 * 4);
            attr = true;
        }
        const size_t smem = (size_t)S * cin * sizeof(float);
        fc_norm_cluster_kernel<S><<<8 * ((n + S - 1) / S), 256, smem, (cudaStream_t)stream>>>(in, w, bias, out, n, cin, cout,
                                                                                            relu, normalize);
        FM_CHECK_LAUNCH("fm_fc_norm");
        return FM_OK;
    }
    if ((cin & 3) == 0 && cin <= 4096 && cout <= 4096) {
        constexpr int S = 2;
        const size_t smem = (size_t)S * (cin + cout) * sizeof(float);
        fc_norm_warp_kernel<S><<<(n + S - 1) / S, 256, smem, (cudaStream_t)stream>>>(in, w, bias, out, n, cin, cout,
                                                                                    relu, normalize);
        FM_CHECK_LAUNCH("fm_fc_norm");
        return FM_OK;
    }
    fc_norm_kernel<<<n, 256, (cin + 8) * sizeof(float), (cudaStream_t)stream>>>(in, w, bias, out, cin, cout, relu,
                                                                               normalize);
    FM_CHECK_LAUNCH("fm_fc_norm");
    return FM_OK;
}

extern "C" int fm_add_act_strided(const void* a, int a_stride, int a_off, const void* b, int b_stride, int b_off,
                                  void* out, int o_stride, int o_off, long long pixels, int c, int act, void* stream) {
    if (pixels <= 0 || c <= 0) return FM_OK;
    if (fm_vec_add_act_strided(a, a_stride, a_off, b, b_stride, b_off, out, o_stride, o_off, pixels, c, act,
                               (cudaStream_t)stream)) {
        FM_CHECK_LAUNCH("fm_add_act_strided");
        return FM_OK;
    }
    add_act_strided_kernel<<<grid_for((size_t)pixels * c), 256, 0, (cudaStream_t)stream>>>(
        (const __half*)a, a_stride, a_off, (const __half*)b, b_stride, b_off, (__half*)out, o_stride, o_off,
        (size_t)pixels, c, act);
    FM_CHECK_LAUNCH("fm_add_act_strided");
    return FM_OK;
}
