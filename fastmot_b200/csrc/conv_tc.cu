// tcgen05 / TMEM implicit-GEMM convolution for sm_100a (the dense contraction of the detector and ReID stacks).
//
//   D[M = N*Ho*Wo pixels, Cout] = im2col(X)[M, K = kh*kw*Cin] * W^T[K, Cout],  fp16 operands, fp32 accumulate in TMEM.
//
// conv_tc_kernel<BN, STAGES, SPLITK>: one CTA computes a 128 x BN output tile.  The K loop walks 64-element slices:
// all 128 threads gather the A slice (128 pixels x 64 reduction elements, zero-filled at the image border; 1x1 /
// stride-1 layers skip the im2col index arithmetic) and the B slice (BN filters x 64) with 16-byte cp.async (LDGSTS)
// copies straight into the 128-byte-swizzled K-major layout the tensor core reads; one thread issues four
// tcgen05.mma (UMMA 128 x BN x 16) per slice and commits them to an mbarrier.  The smem ring is STAGES deep with
// STAGES-1 slices of copies in flight; a stage is refilled only after the commit barrier of the MMAs that read it
// fired.  STAGES is chosen against the wave capacity of the layer (launch code at the bottom): deep rings for
// one-wave layers, shallow rings (more resident CTAs) for many-wave ones.
// Epilogue: each warp reads its 32 TMEM lanes with tcgen05.ld (one output row per thread), stages the fp16 tile in
// the now idle ring buffers and writes it out with lanes running along the channels (row-coalesced 16-byte
// stores), fusing bias + activation (+ residual, + channel-slice offsets, so route/concat layers need no copy).
// Under-filled grids split K over blockIdx.z (as many splits as fit in ONE wave of resident CTAs); the partial
// tiles go through the same smem transpose to an fp32 workspace and splitk_reduce_kernel finishes the layer.
// All launches are programmatic-dependent-launch chains (common.cuh).
//
// The split-K scratch belongs to the caller (FmConvDesc.ws): one per engine / stream, no library-global state.
//
// Replaces the TensorRT conv tactics behind fastmot/utils/inference.py:106-117.  Descriptor bit layouts follow the
// PTX ISA "tcgen05 matrix / instruction descriptor" tables (cross-checked against cute/arch/mma_sm100_desc.hpp).
#include "common.cuh"
#include "../../include/fastmot_b200.h"
#include "conv_act.cuh"
#include <stdlib.h>

namespace {

constexpr int TC_BM = 128;
constexpr int TC_BK = 64;           // fp16 elements per K slice = one 128-byte swizzle row

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity));
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor: start>>4 | LBO=1 | SBO=1024B | version 1 | layout 2
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// kind::f16 instruction descriptor: D=f32, A=B=f16, both K-major, M=128, N=BN
__device__ __forceinline__ uint32_t make_idesc(int bn) {
    uint32_t d = 0;
    d |= 1u << 4;                       // c_format = F32
    d |= 0u << 7;                       // a_format = F16
    d |= 0u << 10;                      // b_format = F16
    d |= (uint32_t)(bn >> 3) << 17;     // n_dim
    d |= (uint32_t)(TC_BM >> 4) << 24;  // m_dim
    return d;
}

__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,"
        "%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ unsigned long long* g_dbg_dev = nullptr;   // optional per-CTA phase timestamps (scripts/bench_conv.py)

__device__ __forceinline__ unsigned long long gtimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#define DBG_STAMP(k)                                                                             \
    do {                                                                                         \
        if (g_dbg_dev && tid == 0) {                                                             \
            const int cta = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);      \
            if (cta < 4096) g_dbg_dev[cta * 8 + (k)] = gtimer();                                 \
        }                                                                                        \
    } while (0)

// Epilogue for 32 consecutive output channels of one pixel row: bias + activation (+ residual) and fp16 NHWC store.
// Stores are 32 bytes per thread (st.global.v8.b32 = one full DRAM sector) whenever the slice is 32-byte aligned:
// 16-byte stores to rows that are hundreds of bytes apart are partial-sector writes and were measured to make the
// epilogue 20-60 us per CTA (profiles/r01_conv_phase_timing.md).
__device__ __forceinline__ void epilogue_store32(float (&v32)[32], size_t m, int n_base, const FmConvDesc& d,
                                                 const float* __restrict__ bias, const __half* __restrict__ residual,
                                                 __half* __restrict__ out, int act, bool res_first) {
    // every index into v32 is a compile-time constant after unrolling: the accumulators must stay in registers
    // (a first version indexed them dynamically and the epilogue ran out of local memory)
    const bool al16 = ((d.cout_stride | d.cout_offset) & 15) == 0;
    const bool res16 = residual != nullptr && ((d.res_stride | d.res_offset) & 15) == 0;
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
        const int n = n_base + hh * 16;
        if (n < d.cout) {
            const bool full = n + 16 <= d.cout;
            float v[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const float x = v32[hh * 16 + q] + ((bias && (full || n + q < d.cout)) ? bias[n + q] : 0.f);
                v[q] = res_first ? x : tc_act(x, act);
            }
            __half* op = out + m * d.cout_stride + d.cout_offset + n;
            if (full && al16) {
                if (residual) {
                    const __half* rp = residual + m * d.res_stride + d.res_offset + n;
                    if (res16) {
                        uint32_t r[8];
                        asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
                                       "=r"(r[7])
                                     : "l"(rp));
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&r[q]));
                            v[2 * q] += f.x; v[2 * q + 1] += f.y;
                        }
                    } else {
#pragma unroll
                        for (int q = 0; q < 16; ++q) v[q] += __half2float(rp[q]);
                    }
                }
                uint32_t w[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float a = res_first ? tc_act(v[2 * q], act) : v[2 * q];
                    const float b = res_first ? tc_act(v[2 * q + 1], act) : v[2 * q + 1];
                    const __half2 h = __floats2half2_rn(a, b);
                    w[q] = *reinterpret_cast<const uint32_t*>(&h);
                }
                asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(op), "r"(w[0]), "r"(w[1]),
                             "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
                             : "memory");
            } else {
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    if (n + q < d.cout) {
                        float x = v[q];
                        if (residual) x += __half2float(residual[m * d.res_stride + d.res_offset + n + q]);
                        op[q] = __float2half(res_first ? tc_act(x, act) : x);
                    }
                }
            }
        }
    }
}


// Staged epilogue of one warp's 32 x BN accumulator slab.  A thread owns one output ROW in TMEM; writing it straight
// out makes every warp store touch 32 different lines (measured: 15-60 us of LSU replays per CTA).  The slab goes
// through shared memory instead and is written back with lanes running along the channels, so each store
// instruction covers whole rows.  Needs 8-channel alignment of the output / residual views.
template <int BN>
__device__ __forceinline__ void epilogue_staged(uint8_t* stg, uint32_t lane_addr, int lane, int m_warp0, int n0, int M,
                                                const FmConvDesc& d, const float* __restrict__ bias,
                                                const __half* __restrict__ residual, __half* __restrict__ out,
                                                int act, bool res_first) {
    constexpr int PITCH = BN * 2 + 16;           // bytes; +16 keeps the per-row 16-byte writes conflict-free
    // phase 1: raw accumulators -> fp16 -> smem (row per lane).  Bias / activation wait for phase 2, where a lane
    // keeps the same 8 channels for every row and so loads its bias once.
#pragma unroll 1
    for (int j0 = 0; j0 < BN; j0 += 32) {
        float v32[32];
        tmem_ld32(lane_addr + j0, v32);
#pragma unroll
        for (int q0 = 0; q0 < 32; q0 += 8) {
            uint32_t w[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const __half2 h = __floats2half2_rn(v32[q0 + 2 * q], v32[q0 + 2 * q + 1]);
                w[q] = *reinterpret_cast<const uint32_t*>(&h);
            }
            *reinterpret_cast<uint4*>(stg + lane * PITCH + (j0 + q0) * 2) = make_uint4(w[0], w[1], w[2], w[3]);
        }
    }
    __syncwarp();
    constexpr int CPR = BN / 8;                  // 16-byte chunks per row
    constexpr int RPI = 32 / CPR;                // rows per store instruction
    const int chunk = lane % CPR, rsub = lane / CPR;
    const int n = n0 + chunk * 8;
    if (n < d.cout) {                            // cout % 8 == 0 (staged_ok): the whole chunk is in range
        float b8[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) b8[q] = bias ? bias[n + q] : 0.f;
        const __half* rp = residual ? residual + d.res_offset + n : nullptr;
        __half* op = out + d.cout_offset + n;
#pragma unroll 2
        for (int r0 = 0; r0 < 32; r0 += RPI) {
            const int row = r0 + rsub;
            const int mm = m_warp0 + row;
            if (mm >= M) break;
            const uint4 pk = *reinterpret_cast<const uint4*>(stg + row * PITCH + chunk * 16);
            const __half2* ph = reinterpret_cast<const __half2*>(&pk);
            float x[8];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float2 f = __half22float2(ph[q]);
                x[2 * q] = f.x + b8[2 * q];
                x[2 * q + 1] = f.y + b8[2 * q + 1];
            }
            if (!res_first) tc_act8(x, act);
            if (rp) {
                const uint4 rv = *reinterpret_cast<const uint4*>(rp + (size_t)mm * d.res_stride);
                const __half2* rh = reinterpret_cast<const __half2*>(&rv);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float2 f = __half22float2(rh[q]);
                    x[2 * q] += f.x;
                    x[2 * q + 1] += f.y;
                }
            }
            if (res_first) tc_act8(x, act);
            uint32_t w[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const __half2 h = __floats2half2_rn(x[2 * q], x[2 * q + 1]);
                w[q] = *reinterpret_cast<const uint32_t*>(&h);
            }
            *reinterpret_cast<uint4*>(op + (size_t)mm * d.cout_stride) = make_uint4(w[0], w[1], w[2], w[3]);
        }
    }
    __syncwarp();
}

__device__ __forceinline__ bool staged_ok(const FmConvDesc& d, const void* residual) {
    return ((d.cout_stride | d.cout_offset | d.cout) & 7) == 0 &&
           (residual == nullptr || ((d.res_stride | d.res_offset) & 7) == 0);
}

// SPLITK instantiations carry the partial-sum path and the last-CTA reduction (they need ~135 registers; the plain
// ones stay at 72 so that 7-8 CTAs fit on an SM for the memory-bound 1x1 layers).
template <int BN, int STAGES, bool SPLITK>
__global__ void __launch_bounds__(128) conv_tc_kernel(FmConvDesc d, const __half* __restrict__ in,
                                                       const __half* __restrict__ wgt, const float* __restrict__ bias,
                                                       const __half* __restrict__ residual, __half* __restrict__ out,
                                                       float* ws, int slices_per_split) {
    // 1024-byte alignment for the 128B swizzle atoms, by declaration (rounding the pointer through an integer makes the
    // compiler fall back to generic LD / ST for the staged epilogue)
    extern __shared__ __align__(1024) uint8_t smem[];
    constexpr int A_BYTES = TC_BM * 128, B_BYTES = BN * 128, STAGE_BYTES = A_BYTES + B_BYTES;
    __shared__ uint64_t bar_stage[STAGES];
    __shared__ uint64_t bar_done;
    __shared__ uint32_t s_tmem;

    const int tid = threadIdx.x, warp = tid >> 5;
    fm_pdl_trigger();
    DBG_STAMP(0);
    const int m0 = blockIdx.x * TC_BM, n0 = blockIdx.y * BN;
    const int M = d.n * d.ho * d.wo;
    const int Ktot = d.kh * d.kw * d.cin;
    const int nk_total = (Ktot + TC_BK - 1) / TC_BK;
    // split-K: blockIdx.z owns K slices [kb0, kb0 + nk); partial sums go to the fp32 workspace
    const int kb0 = blockIdx.z * slices_per_split;
    const int nk = min(nk_total - kb0, slices_per_split);

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) mbar_init(&bar_stage[s], 1);
        mbar_init(&bar_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)),
                     "r"((uint32_t)(BN < 32 ? 32 : BN)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem;
    DBG_STAMP(1);

    // ---- per-thread gather plan: chunk column c (16 B = 8 channels), rows (tid>>3) + 16*i ----
    const int c = tid & 7;
    const int rbase = tid >> 3;
    int pn[8], ph[8], pw[8];
    // 1x1 / stride 1 / no padding: input pixel == output pixel, no index arithmetic at all
    const bool pointwise = d.kh == 1 && d.kw == 1 && d.stride == 1 && d.pad == 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + rbase + 16 * i;
        if (pointwise) {
            pn[i] = m < M ? 0 : -1; ph[i] = 0; pw[i] = 0;
        } else if (m < M) {
            const int wo = m % d.wo, t = m / d.wo, ho = t % d.ho;
            pn[i] = t / d.ho;
            ph[i] = ho * d.stride - d.pad;
            pw[i] = wo * d.stride - d.pad;
        } else {
            pn[i] = -1; ph[i] = 0; pw[i] = 0;
        }
    }
    const uint32_t idesc = make_idesc(BN);

    // cp.async (LDGSTS) gather of K-slice `kb` into ring stage `st`; out-of-image / out-of-range chunks are
    // zero-filled by passing src-size 0.
    auto issue_loads = [&](int kb, int st) {
        uint8_t* sA = smem + (size_t)st * STAGE_BYTES;
        uint8_t* sB = sA + A_BYTES;
        const int kelem = (kb0 + kb) * TC_BK + c * 8;
        const bool kvalid = kelem < Ktot;
        if (pointwise) {
            const __half* base = in + d.cin_offset + (kvalid ? kelem : 0);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int r = rbase + 16 * i;
                const bool ok = kvalid && pn[i] >= 0;
                const __half* src = ok ? base + (size_t)(m0 + r) * d.cin_stride : in;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(sA + r * 128 + ((c ^ (r & 7)) << 4))),
                             "l"(src), "r"(ok ? 16u : 0u));
            }
        } else {
        const int tap = kvalid ? kelem / d.cin : 0;
        const int cch = kvalid ? kelem - tap * d.cin : 0;
        const int fr = tap / d.kw, fs = tap - fr * d.kw;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int r = rbase + 16 * i;
            const __half* src = in;
            uint32_t bytes = 0;
            if (kvalid && pn[i] >= 0) {
                const int hi = ph[i] + fr, wi = pw[i] + fs;
                if (hi >= 0 && hi < d.hi && wi >= 0 && wi < d.wi) {
                    src = in + (((size_t)pn[i] * d.hi + hi) * d.wi + wi) * d.cin_stride + d.cin_offset + cch;
                    bytes = 16;
                }
            }
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(sA + r * 128 + ((c ^ (r & 7)) << 4))),
                         "l"(src), "r"(bytes));
        }
        }
#pragma unroll
        for (int i = 0; i < BN / 16; ++i) {
            const int r = rbase + 16 * i;
            const int n = n0 + r;
            const bool ok = kvalid && n < d.cout;
            const __half* src = ok ? wgt + (size_t)n * Ktot + kelem : wgt;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(sB + r * 128 + ((c ^ (r & 7)) << 4))),
                         "l"(src), "r"(ok ? 16u : 0u));
        }
    };

    // everything above is independent of other kernels' output; the activations (and the split-K workspace /
    // output buffers, which an earlier kernel may still be reading) are not
    fm_pdl_wait();
#pragma unroll
    for (int p = 0; p < STAGES - 1; ++p) {
        if (p < nk) issue_loads(p, p);
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    DBG_STAMP(2);
    for (int kb = 0; kb < nk; ++kb) {
        const int s = kb % STAGES;
        const int kn = kb + STAGES - 1;
        if (kn < nk) {
            const int sn = kn % STAGES;
            // the MMAs that last read stage `sn` (slice kn - STAGES) must have retired before it is overwritten
            if (kn >= STAGES) mbar_wait(&bar_stage[sn], (uint32_t)(((kn / STAGES) - 1) & 1));
            issue_loads(kn, sn);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group %0;" ::"n"(STAGES - 1) : "memory");   // slice kb has landed (this thread)
        fence_async_smem();      // make the smem writes visible to the tensor-core (async) proxy
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            const uint32_t a_addr = smem_u32(smem + (size_t)s * STAGE_BYTES), b_addr = a_addr + A_BYTES;
#pragma unroll
            for (int k = 0; k < TC_BK / 16; ++k) {
                // advance 16 elements (32 bytes) along K inside the swizzle atom: +2 in the encoded start address
                const uint64_t adesc = make_smem_desc(a_addr + k * 32);
                const uint64_t bdesc = make_smem_desc(b_addr + k * 32);
                mma_f16(tmem_base, adesc, bdesc, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            }
            tc_commit(&bar_stage[s]);    // fires when the MMAs reading this stage are done
        }
    }
    DBG_STAMP(3);
    if (tid == 0) tc_commit(&bar_done);
    mbar_wait(&bar_done, 0);
    tc_fence_after();
    DBG_STAMP(4);

    // ---- epilogue: TMEM lane = tile row ----
    const int m = m0 + tid;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
    const int act = d.act & 0xff;
    const bool res_first = (d.act & FM_ACT_AFTER_RESIDUAL) != 0;
    const bool staged = !SPLITK && staged_ok(d, residual);
    if (staged) {
        // the ring buffers are idle now: reuse them as the staging tile
        epilogue_staged<BN>(smem + (size_t)warp * 32 * (BN * 2 + 16), lane_addr, tid & 31, m0 + warp * 32, n0, M, d,
                            bias, residual, out, act, res_first);
    } else if (SPLITK && (d.cout & 3) == 0) {
        // raw fp32 partials through the same smem transpose: whole rows per store instruction (the last-CTA reduction
        // below fences on these stores, so their latency is on the critical path)
        constexpr int PITCH = BN * 4 + 16;
        uint8_t* stg = smem + (size_t)warp * 32 * PITCH;
        const int lane = tid & 31;
#pragma unroll 1
        for (int j0 = 0; j0 < BN; j0 += 32) {
            float v32[32];
            tmem_ld32(lane_addr + j0, v32);
#pragma unroll
            for (int q = 0; q < 32; q += 4)
                *reinterpret_cast<float4*>(stg + lane * PITCH + (j0 + q) * 4) =
                    make_float4(v32[q], v32[q + 1], v32[q + 2], v32[q + 3]);
        }
        __syncwarp();
        constexpr int LPR = BN / 4, RPI = 32 / LPR;
        const int n = n0 + (lane % LPR) * 4;
        float* wz = ws + (size_t)blockIdx.z * M * d.cout;
#pragma unroll 4
        for (int r0 = 0; r0 < 32; r0 += RPI) {
            const int row = r0 + lane / LPR;
            const int mm = m0 + warp * 32 + row;
            if (mm < M && n < d.cout)
                *reinterpret_cast<float4*>(wz + (size_t)mm * d.cout + n) =
                    *reinterpret_cast<const float4*>(stg + row * PITCH + (lane % LPR) * 16);
        }
    } else {
#pragma unroll 1
        for (int j0 = 0; j0 < BN; j0 += 32) {
            float v32[32];
            tmem_ld32(lane_addr + j0, v32);   // warp-collective: every lane executes it
            if (j0 == 0) DBG_STAMP(7);
            if (m >= M) continue;
            if (SPLITK) {                     // raw fp32 partials; bias / activation happen in splitk_reduce_kernel
                float* wp = ws + ((size_t)blockIdx.z * M + m) * d.cout + n0 + j0;
                if (n0 + j0 + 32 <= d.cout && (d.cout & 7) == 0) {      // full-sector (32-byte) stores
#pragma unroll
                    for (int q = 0; q < 32; q += 8)
                        asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(wp + q),
                                     "r"(__float_as_uint(v32[q])), "r"(__float_as_uint(v32[q + 1])),
                                     "r"(__float_as_uint(v32[q + 2])), "r"(__float_as_uint(v32[q + 3])),
                                     "r"(__float_as_uint(v32[q + 4])), "r"(__float_as_uint(v32[q + 5])),
                                     "r"(__float_as_uint(v32[q + 6])), "r"(__float_as_uint(v32[q + 7]))
                                     : "memory");
                } else {
#pragma unroll
                    for (int q = 0; q < 32; ++q)
                        if (n0 + j0 + q < d.cout) wp[q] = v32[q];
                }
                continue;
            }
            epilogue_store32(v32, (size_t)m, n0 + j0, d, bias, residual, out, act, res_first);
        }
    }
    DBG_STAMP(5);
    tc_fence_before();
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                     "r"((uint32_t)(BN < 32 ? 32 : BN)));
    DBG_STAMP(6);
}

__global__ void __launch_bounds__(256) splitk_reduce_kernel(FmConvDesc d, const float* __restrict__ ws, int splits,
                                                             const float* __restrict__ bias,
                                                             const __half* __restrict__ residual,
                                                             __half* __restrict__ out) {
    fm_pdl_trigger();
    fm_pdl_wait();
    const int M = d.n * d.ho * d.wo;
    const size_t total = (size_t)M * d.cout;
    const int act = d.act & 0xff;
    const bool res_first = (d.act & FM_ACT_AFTER_RESIDUAL) != 0;
    if (staged_ok(d, residual)) {
        // 8 channels per thread: two 128-bit partial-sum loads per split, one 128-bit store
        const int cg = d.cout >> 3;
        const size_t total8 = (size_t)M * cg;
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total8;
             i += (size_t)gridDim.x * blockDim.x) {
            const size_t m = i / cg;
            const int n = (int)(i - m * cg) * 8;
            float x[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) x[q] = bias ? bias[n + q] : 0.f;
            const float* wp = ws + m * d.cout + n;
            for (int z = 0; z < splits; ++z) {
                const float4 a = *reinterpret_cast<const float4*>(wp + (size_t)z * total);
                const float4 b2 = *reinterpret_cast<const float4*>(wp + (size_t)z * total + 4);
                x[0] += a.x; x[1] += a.y; x[2] += a.z; x[3] += a.w;
                x[4] += b2.x; x[5] += b2.y; x[6] += b2.z; x[7] += b2.w;
            }
            if (!res_first) tc_act8(x, act);
            if (residual) {
                const uint4 rv = *reinterpret_cast<const uint4*>(residual + m * d.res_stride + d.res_offset + n);
                const __half2* rh = reinterpret_cast<const __half2*>(&rv);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float2 f = __half22float2(rh[q]);
                    x[2 * q] += f.x;
                    x[2 * q + 1] += f.y;
                }
            }
            if (res_first) tc_act8(x, act);
            uint32_t w[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const __half2 h = __floats2half2_rn(x[2 * q], x[2 * q + 1]);
                w[q] = *reinterpret_cast<const uint32_t*>(&h);
            }
            *reinterpret_cast<uint4*>(out + m * d.cout_stride + d.cout_offset + n) = make_uint4(w[0], w[1], w[2], w[3]);
        }
        return;
    }
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t m = i / d.cout;
        const int n = (int)(i - m * d.cout);
        float acc = 0.f;
        for (int z = 0; z < splits; ++z) acc += ws[(size_t)z * total + i];
        float v = acc + (bias ? bias[n] : 0.f);
        if (!res_first) v = tc_act(v, act);
        if (residual) v += __half2float(residual[m * d.res_stride + d.res_offset + n]);
        if (res_first) v = tc_act(v, act);
        out[m * d.cout_stride + d.cout_offset + n] = __float2half(v);
    }
}

template <int BN, int STAGES>
int launch_tc(const FmConvDesc* d, const void* in, const void* wgt, const float* bias, const void* residual, void* out,
              cudaStream_t s) {
    // the ring doubles as the epilogue's staging tile (4 warps x 32 rows x (BN*2+16) bytes)
    constexpr int ring = STAGES * (TC_BM * 128 + BN * 128), stg = 4 * 32 * (BN * 2 + 16);
    constexpr int smem = (ring > stg ? ring : stg) + 1024;
    constexpr int stg32 = 4 * 32 * (BN * 4 + 16);       // fp32 staging of the split-K partial tile
    constexpr int smem_split = (ring > stg32 ? ring : stg32) + 1024;
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(conv_tc_kernel<BN, STAGES, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cudaFuncSetAttribute(conv_tc_kernel<BN, STAGES, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_split);
        attr = true;
    }
    const int M = d->n * d->ho * d->wo;
    const int nk = (d->kh * d->kw * d->cin + TC_BK - 1) / TC_BK;
    dim3 grid(fm_cdiv(M, TC_BM), fm_cdiv(d->cout, BN), 1);
    int sps = nk;
    // split-K when the output tiling alone cannot fill the 148 SMs (batch-1 deep layers)
    const int tiles = grid.x * grid.y;
    static int split_max_tiles = -1;     // FM_CONV_SPLIT_MAXTILES overrides the largest tile count that still splits K
    if (split_max_tiles < 0) {
        const char* e = getenv("FM_CONV_SPLIT_MAXTILES");
        split_max_tiles = e ? atoi(e) : FM_NUM_SMS / 2;
    }
    float* ws = (float*)d->ws;
    if (ws && tiles <= split_max_tiles && nk >= 8) {
        // as many K splits as still fit in ONE wave of resident CTAs: a 149th CTA on 148 single-CTA SMs runs after the
        // others and doubles the layer time (seen on the 40x40 and 20x20 YOLO layers: 156 / 160 CTAs, 2 waves)
        constexpr int per_sm = (227 * 1024) / (smem_split + 1024) > 0 ? (227 * 1024) / (smem_split + 1024) : 1;
        int want = FM_NUM_SMS * per_sm / tiles;
        if (want > nk / 4) want = nk / 4;
        if (want > 1) {
            sps = (nk + want - 1) / want;
            const int splits = (nk + sps - 1) / sps;
            if ((long long)splits * M * d->cout * 4 <= d->ws_bytes) grid.z = splits; else sps = nk;
        }
    }
    if (grid.z > 1)
        fm_launch_pdl(conv_tc_kernel<BN, STAGES, true>, grid, dim3(128), (size_t)smem_split, s, *d, (const __half*)in,
                      (const __half*)wgt, bias, (const __half*)residual, (__half*)out, ws, sps);
    else
        fm_launch_pdl(conv_tc_kernel<BN, STAGES, false>, grid, dim3(128), (size_t)smem, s, *d, (const __half*)in,
                      (const __half*)wgt, bias, (const __half*)residual, (__half*)out, ws, sps);
    if (grid.z > 1) {
        const size_t total = (size_t)M * d->cout;
        const int blocks = (int)((total + 255) / 256 < (size_t)FM_NUM_SMS * 8 ? (total + 255) / 256 : FM_NUM_SMS * 8);
        fm_launch_pdl(splitk_reduce_kernel, dim3(blocks), dim3(256), (size_t)0, s, *d, (const float*)ws, (int)grid.z,
                      bias, (const __half*)residual, (__half*)out);
        fm_count_launches(1);
    }
    return 0;
}

}  // namespace

extern "C" int fm_conv_set_debug(void* dbg) {
    unsigned long long* p = (unsigned long long*)dbg;
    cudaMemcpyToSymbol(g_dbg_dev, &p, sizeof(p));
    return FM_OK;
}

extern "C" int fm_conv2d_tc_supported(const FmConvDesc* d) {
    if (!d) return 0;
    if (d->cin % 8 || d->cin_stride % 8 || d->cin_offset % 8) return 0;   // 16-byte operand chunks
    if ((long long)d->kh * d->kw * d->cin < 32) return 0;                  // not worth a tensor-core tile
    if (d->n * d->ho * d->wo <= 0 || d->cout <= 0) return 0;
    return 1;
}

extern "C" int fm_conv2d_tc(const FmConvDesc* d, const void* in, const void* wgt, const float* bias,
                            const void* residual, void* out, void* stream) {
    FM_REQUIRE(d != nullptr, "fm_conv2d_tc: desc is NULL");
    FM_REQUIRE(fm_conv2d_tc_supported(d), "fm_conv2d_tc: shape not supported by the tcgen05 path");
    cudaStream_t s = (cudaStream_t)stream;
    const int nk = (d->kh * d->kw * d->cin + TC_BK - 1) / TC_BK;
    const int m_tiles_all = (d->n * d->ho * d->wo + TC_BM - 1) / TC_BM;
    // ring depth follows the K extent: short reductions (OSNet 1x1) want many co-resident CTAs, long ones (3x3 on
    // wide layers) want many slices of copies in flight
    static int force_bn = -1;      // FM_CONV_BN=32|64|128 overrides the tile width (experiments only)
    if (force_bn < 0) { const char* e = getenv("FM_CONV_BN"); force_bn = e ? atoi(e) : 0; }
    int bn = d->cout <= 32 ? 32 : d->cout <= 64 ? 64 : 128;
    if (bn == 128) {
        // between half a wave and two waves of 128-wide tiles, 64-wide tiles fill the 148 SMs better (measured on
        // 128->128 @ 224x16x8 and the 80x80 YOLO 3x3); fewer tiles than that go to split-K instead
        const int tiles128 = m_tiles_all * ((d->cout + 127) / 128);
        if (tiles128 > FM_NUM_SMS / 2 && tiles128 < 2 * FM_NUM_SMS) bn = 64;
    }
    if (force_bn) bn = force_bn;
    // Ring depth: deep rings hide the gather latency of a long K loop, but they cost residency (6 x 32 KB = one CTA
    // per SM).  What decides is how the tile count sits against one wave of resident CTAs:
    //   * the layer fits in one wave at the deep setting          -> deep ring (and split-K fills the idle SMs),
    //   * it fits in one wave only with a shallower ring          -> that ring (a second, mostly empty wave doubles
    //                                                                the layer time),
    //   * many waves either way (OSNet stem, first YOLO layers)   -> shallow ring, residency hides the latency.
    static int rules = -1;          // FM_CONV_RULES=0 restores the K-only rule (A/B timing)
    if (rules < 0) { const char* e = getenv("FM_CONV_RULES"); rules = (e && e[0] == '0') ? 0 : 1; }
    const int tiles = m_tiles_all * ((d->cout + bn - 1) / bn);
    if (bn == 32) {
        if (nk == 1) launch_tc<32, 1>(d, in, wgt, bias, residual, out, s);
        else if (nk <= 2 || (rules && tiles > 5 * FM_NUM_SMS)) launch_tc<32, 2>(d, in, wgt, bias, residual, out, s);
        else launch_tc<32, 4>(d, in, wgt, bias, residual, out, s);
    } else if (bn == 64) {
        // 64-wide: 4 stages = 96 KB (2 CTAs / SM), 2 stages = 48 KB (4 / SM)
        if (nk == 1) launch_tc<64, 1>(d, in, wgt, bias, residual, out, s);
        else if (nk <= 2 || (rules ? tiles > 2 * FM_NUM_SMS : m_tiles_all >= 8 * FM_NUM_SMS))
            launch_tc<64, 2>(d, in, wgt, bias, residual, out, s);   // OSNet 7x7 stem: 727 -> 457 us measured
        else launch_tc<64, 4>(d, in, wgt, bias, residual, out, s);
    } else {
        // 128-wide: 6 stages = 192 KB (1 CTA / SM), 3 stages = 96 KB (2 / SM), 2 stages = 64 KB (3 / SM)
        if (nk == 1) launch_tc<128, 1>(d, in, wgt, bias, residual, out, s);
        else if (nk <= 2 || (rules && tiles > 2 * FM_NUM_SMS)) launch_tc<128, 2>(d, in, wgt, bias, residual, out, s);
        else if (nk < 6 || (rules && tiles > FM_NUM_SMS)) launch_tc<128, 3>(d, in, wgt, bias, residual, out, s);
        else launch_tc<128, 6>(d, in, wgt, bias, residual, out, s);
    }
    FM_CHECK_LAUNCH("fm_conv2d_tc");
    return FM_OK;
}
