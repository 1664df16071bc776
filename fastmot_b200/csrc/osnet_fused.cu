// Fused OSNet OSBlock kernels for sm_100a (replace the per-layer launches of the ReID stack; role of the TensorRT
// OSNet engine behind fastmot/utils/inference.py:106-117 + fastmot/feature_extractor.py:48-74).
//
// osb_streams_kernel<W, MID, T, NACC, NS, NW>  ("kernel S")
//   One CTA owns a strip of T x 128 pixels of one crop (all W columns, SR = 128 T / W rows) and computes, without
//   leaving the SM,   x1 = relu(conv1x1(x) + b1)   and the four Lite-3x3 streams of the block
//       stream s:  (1x1 linear conv -> depthwise 3x3 + bias + ReLU)  x (s + 1)
//   writing only the four stream outputs ("tails", fp16 NHWC) and their per-channel sums (for the channel gate).
//   * conv1: A tiles arrive by TMA (one box per image row, 128-byte swizzle, zero fill outside the image), weight
//     slices by cp.async.bulk, through an NS-stage mbarrier ring; tcgen05.mma with both operands in shared memory.
//   * activations never touch shared memory as MMA operands: the conv1 result (x1) and the running stream activation
//     live in TMEM as fp16 A operands (lane = pixel, two channels per column), written with tcgen05.st by the
//     thread that owns the lane; the 1x1 convs of the streams are tcgen05.mma with A from TMEM.
//   * tile t holds the image rows y = t (mod T) of the strip, so the T pixels a thread owns (one lane in every
//     tile) are vertically adjacent: the depthwise 3x3 loads (T + 2) x 3 neighbours for T outputs.  Its input (the
//     pointwise output, fp16) is the only activation in shared memory: chunk-planar [8 channels][row][x], with a
//     zero row above and below.
//   * strips of stage 1 carry a 4-row halo that is recomputed (4 = the deepest stream); rows outside the image are
//     forced to zero after every pointwise conv (= the zero padding of the depthwise conv).
//   Warp roles: NW (8 or 16) compute warps (TMEM lane quarter = warp & 3, channel group = warp >> 2), 1 control warp
//   (one thread issues TMA, bulk copies and every tcgen05.mma).
//
// Layouts: activations NHWC fp16; weight images are packed on the host (fastmot_b200/packing.py).
#include "tc_common.cuh"
#include "../../include/fastmot_b200.h"
#include <string.h>

namespace {

using namespace tc;


// optional phase stamps (scripts/osb_phases.py): clock64 of one CTA at the phase boundaries of every level
__device__ long long* g_osb_dbg = nullptr;
#define OSB_STAMP(slot)                                                                        \
    do {                                                                                       \
        if (g_osb_dbg && blockIdx.x == gridDim.x / 2) g_osb_dbg[(slot)] = clock64();           \
    } while (0)

struct OsbStreamsArgs {
    int H, n_crops, cin, R, halo, strips;
    const uint8_t* w1;      // conv1 weight image: cin/64 slices of [MID x 128 B]
    const float* b1;        // [MID]
    const uint8_t* pw;      // 10 pointwise images, each PW_BYTES
    const uint8_t* dw;      // 10 blobs, each DW_BYTES: [9][MID] fp16 | pw bias f32[MID] | dw bias f32[MID]
    __half* tails[4];       // chunk-planar [n][MID / 8][H][W][8] each
    float* gap_part;        // [n][strips][4][MID]
};

// ---- thread-block cluster helpers (DSMEM halo exchange between the strips of one crop) ----
__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void st_cluster_v4(uint32_t local_saddr, uint32_t rank, uint4 v) {
    uint32_t raddr;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(local_saddr), "r"(rank));
    asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(raddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
                 : "memory");
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

template <int N>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t* r);
template <>
__device__ __forceinline__ void tmem_ld_cols<16>(uint32_t taddr, uint32_t* r) { tmem_ld16(taddr, r); }
template <>
__device__ __forceinline__ void tmem_ld_cols<24>(uint32_t taddr, uint32_t* r) {
    tmem_ld16(taddr, r);
    tmem_ld8(taddr + 16, r + 16);
}
template <>
__device__ __forceinline__ void tmem_ld_cols<32>(uint32_t taddr, uint32_t* r) { tmem_ld32(taddr, r); }
template <>
__device__ __forceinline__ void tmem_ld_cols<48>(uint32_t taddr, uint32_t* r) {
    tmem_ld32(taddr, r);
    tmem_ld16(taddr + 32, r + 32);
}
template <>
__device__ __forceinline__ void tmem_ld_cols<64>(uint32_t taddr, uint32_t* r) {
    tmem_ld32(taddr, r);
    tmem_ld32(taddr + 32, r + 32);
}

// level -> (stream, depth in stream)
__device__ __forceinline__ void level_sj(int lvl, int& s, int& j) {
    if (lvl < 1) { s = 0; j = lvl; }
    else if (lvl < 3) { s = 1; j = lvl - 1; }
    else if (lvl < 6) { s = 2; j = lvl - 3; }
    else { s = 3; j = lvl - 6; }
}

template <int W, int MID, int T, int NACC, int NS, int NW, int CL = 1>
struct SCfg {
    static constexpr int kThreads = NW * 32 + 32;
    static constexpr int SR = 128 * T / W;                  // strip rows
    static constexpr int RQ = 32 / W;                       // image rows per TMEM lane quarter and tile
    static constexpr int TROWS = 128 / W;                   // image rows per tile
    static constexpr int CW = MID / (NW / 4);               // channels per compute warp
    static constexpr int NCH = MID / 8;                     // 16-byte chunks per pixel
    static constexpr int NSL = (MID + 63) / 64;             // K slices of the pointwise weights
    static constexpr int PW_BYTES = NSL * MID * 128;
    static constexpr int DW_BYTES = 9 * MID * 2 + 2 * MID * 4;
    static constexpr int PLANE = (SR + 2) * W * 16;         // bytes of one chunk plane of P
    static constexpr int P_BYTES = NCH * PLANE;
    static constexpr int STAGE = 16384 + MID * 128;
    static constexpr int RING = NS * STAGE;
    static constexpr int REGION = (P_BYTES > RING ? P_BYTES : RING);
    static constexpr int X1_COL = 0, ACT_COL = T * MID / 2, ACC_COL = T * MID;
    static constexpr int TMEM_NEED = ACC_COL + NACC * MID;
    static constexpr int TMEM_COLS = TMEM_NEED <= 256 ? 256 : 512;
    static constexpr int DWB = (DW_BYTES + 127) / 128 * 128;  // one depthwise / bias blob
    static constexpr int DW_AREA = (2 * DWB + 1023) / 1024 * 1024;   // keeps the swizzled region 1024-byte aligned
    static constexpr int SMEM = PW_BYTES + DW_AREA + REGION + 4 * NW * CW;
    static_assert(TMEM_NEED <= 512, "TMEM budget");
    static_assert(W == 8 || W == 16 || W == 32, "W");
    static_assert(MID % 32 == 0 && MID <= 128, "MID");
};

// CL > 1: the CL strips of a crop form a thread-block cluster; nothing is recomputed: after every pointwise conv the
// first / last row of a strip is also stored into the neighbour strip's halo row through distributed shared memory.
// Cluster barrier protocol (every thread of the cluster alternates arrive / wait):
//   arrive (conv1 ring dead)  |  per level:  wait -> pointwise epilogue (local + remote rows) -> arrive, wait ->
//   depthwise -> arrive  |  final wait.
template <int W, int MID, int T, int NACC, int NS, int NW, int CL, int MINB>
__global__ void __launch_bounds__(NW * 32 + 32, MINB)
osb_streams_kernel(const __grid_constant__ CUtensorMap map_x, OsbStreamsArgs a) {
    using C = SCfg<W, MID, T, NACC, NS, NW, CL>;
    constexpr int kComputeWarps = NW, kComputeThreads = NW * 32;
    // 1024-byte alignment comes from the declaration: rounding the pointer through an integer makes the compiler lose
    // the shared address space and emit generic LD/ST (seen in SASS: ~3x slower depthwise loop)
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* s_pw = smem;                                               // pointwise weight image (one level)
    uint8_t* s_dw0 = s_pw + C::PW_BYTES;                                // two depthwise / bias blobs
    constexpr int DWB = C::DWB;
    uint8_t* s_region = s_dw0 + C::DW_AREA;                             // conv1 ring (1024-aligned), later the P planes
    float* s_gap = reinterpret_cast<float*>(s_region + C::REGION);      // [NW warps][CW]
    __shared__ uint64_t ring_full[NS], ring_empty[NS], acc_full[NACC], acc_empty[NACC];
    __shared__ uint64_t pw_full, pw_empty, dw_full[2], act_ready;
    __shared__ uint32_t s_tmem;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    fm_pdl_trigger();
    const int crop = blockIdx.x / a.strips, strip = blockIdx.x - crop * a.strips;   // CL > 1: strip == cluster rank
    const int y0 = strip * a.R - a.halo;                                // image row of strip row 0
    const int nsl1 = a.cin >> 6;

    if (tid == 0) {
        if (smem_u32(smem) & 1023u) __trap();                           // swizzled tiles need the alignment
        for (int i = 0; i < NS; ++i) { mbar_init(&ring_full[i], 1); mbar_init(&ring_empty[i], 1); }
        for (int i = 0; i < NACC; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], kComputeWarps); }
        mbar_init(&pw_full, 1); mbar_init(&pw_empty, 1);
        mbar_init(&dw_full[0], 1); mbar_init(&dw_full[1], 1);
        mbar_init(&act_ready, kComputeWarps);
        mbar_fence_init();
    }
    if (warp == kComputeWarps) {
        tmem_alloc<C::TMEM_COLS>(&s_tmem);
        if (lane == 0) tma_prefetch_desc(&map_x);
    }
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = s_tmem;
    fm_pdl_wait();
    if (tid == 0) OSB_STAMP(0);

    if (warp == kComputeWarps) {
        // =========================================== control warp ==============================================
        // The whole warp runs this code (warp-uniform operands stay in uniform registers: a single diverged lane
        // needed a register-to-uniform move per MMA operand, ~100 cycles per tcgen05.mma); lane 0 issues.
        const bool leader = lane == 0;
        if (leader) {
            // weights of level 0 (constants: no dependence on earlier kernels)
            mbar_expect_tx(&pw_full, C::PW_BYTES);
            bulk_load(s_pw, a.pw, C::PW_BYTES, &pw_full);
            mbar_expect_tx(&dw_full[0], C::DW_BYTES);
            bulk_load(s_dw0, a.dw, C::DW_BYTES, &dw_full[0]);
        }
        const uint32_t idesc = idesc_f16(MID);
        uint32_t acc_use = 0;                                      // accumulator ring uses so far
        // ---- conv1: ring of (A tile slice by TMA, weight slice by bulk copy) ----
        const int iters = T * nsl1;
        auto issue = [&](int i) {
            const int s = i % NS, t = i / nsl1, ks = i - t * nsl1;
            if (i >= NS) mbar_wait(&ring_empty[s], (uint32_t)((i / NS - 1) & 1));
            uint8_t* sa = s_region + (size_t)s * C::STAGE;
            if (leader) {
                mbar_expect_tx(&ring_full[s], C::STAGE);
#pragma unroll
                for (int rr = 0; rr < C::TROWS; ++rr) {
                    const int y = y0 + rr * T + t;
                    tma_load_3d(sa + rr * W * 128, &map_x, &ring_full[s], ks * 64, y * W, crop);
                }
                bulk_load(sa + 16384, a.w1 + (size_t)ks * MID * 128, MID * 128, &ring_full[s]);
            }
        };
        for (int i = 0; i < NS - 1 && i < iters; ++i) issue(i);
        for (int i = 0; i < iters; ++i) {
            if (i + NS - 1 < iters) issue(i + NS - 1);
            const int s = i % NS, t = i / nsl1, ks = i - t * nsl1;
            const int ai = (int)(acc_use % NACC);
            if (ks == 0 && acc_use >= NACC) mbar_wait(&acc_empty[ai], (uint32_t)((acc_use / NACC - 1) & 1));
            mbar_wait(&ring_full[s], (uint32_t)((i / NS) & 1));
            fence_after();
            const uint32_t sa = smem_u32(s_region + (size_t)s * C::STAGE), sb = sa + 16384;
            if (leader) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    mma_ss(tmem + C::ACC_COL + ai * MID, smem_desc_sw128(sa + k * 32), smem_desc_sw128(sb + k * 32),
                           idesc, (ks > 0 || k > 0) ? 1u : 0u);
                commit(&ring_empty[s]);
                if (ks == nsl1 - 1) commit(&acc_full[ai]);
            }
            if (ks == nsl1 - 1) ++acc_use;
            __syncwarp();
        }
        OSB_STAMP(1);
        if (CL > 1) cluster_arrive();
        // ---- the ten pointwise convs: A from TMEM ----
        uint64_t bdesc_pw[MID / 16];
#pragma unroll
        for (int k = 0; k < MID / 16; ++k)
            bdesc_pw[k] = smem_desc_sw128(smem_u32(s_pw) + (k >> 2) * MID * 128 + (k & 3) * 32);
        for (int lvl = 0; lvl < 10; ++lvl) {
            int s, j;
            level_sj(lvl, s, j);
            mbar_wait(&pw_full, (uint32_t)(lvl & 1));
            if (leader) OSB_STAMP(16 + lvl * 16 + 0);
            mbar_wait(&act_ready, (uint32_t)(lvl & 1));
            if (leader) OSB_STAMP(16 + lvl * 16 + 1);
            fence_after();
            const uint32_t a_base = tmem + (j == 0 ? C::X1_COL : C::ACT_COL);
            for (int t = 0; t < T; ++t) {
                const int ai = (int)(acc_use % NACC);
                if (acc_use >= NACC) mbar_wait(&acc_empty[ai], (uint32_t)((acc_use / NACC - 1) & 1));
                fence_after();
                if (leader) {
#pragma unroll
                    for (int k = 0; k < MID / 16; ++k)
                        mma_ts(tmem + C::ACC_COL + ai * MID, a_base + t * (MID / 2) + k * 8, bdesc_pw[k], idesc,
                               k > 0 ? 1u : 0u);
                    commit(&acc_full[ai]);
                }
                ++acc_use;
                __syncwarp();
            }
            if (leader) {
                commit(&pw_empty);
                OSB_STAMP(16 + lvl * 16 + 2);
            }
            if (lvl + 1 < 10) {
                if (leader) {
                    // the depthwise blob of level lvl - 1 is dead (act_ready of this level was its last reader)
                    uint8_t* sd = s_dw0 + ((lvl + 1) & 1) * DWB;
                    mbar_expect_tx(&dw_full[(lvl + 1) & 1], C::DW_BYTES);
                    bulk_load(sd, a.dw + (size_t)(lvl + 1) * C::DW_BYTES, C::DW_BYTES, &dw_full[(lvl + 1) & 1]);
                }
                mbar_wait(&pw_empty, (uint32_t)(lvl & 1));           // this level's MMAs have read s_pw
                if (leader) {
                    OSB_STAMP(16 + lvl * 16 + 3);
                    mbar_expect_tx(&pw_full, C::PW_BYTES);
                    bulk_load(s_pw, a.pw + (size_t)(lvl + 1) * C::PW_BYTES, C::PW_BYTES, &pw_full);
                }
            }
            __syncwarp();
            if (CL > 1) { cluster_wait(); cluster_arrive(); cluster_wait(); cluster_arrive(); }
        }
        if (CL > 1) cluster_wait();
    } else {
        // =========================================== compute warps ============================================
        const int q = warp & 3, g = warp >> 2;
        const int yl = lane / W, x = lane % W;
        const int run = q * C::RQ + yl;                                 // vertical run of T rows owned by this thread
        const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
        const int c0 = g * C::CW;                                       // first channel of this warp
        uint32_t acc_use = 0;
        // ---- conv1 epilogue: relu(acc + b1) -> fp16 -> x1 in TMEM ----
        {
            float b[C::CW];
#pragma unroll
            for (int i = 0; i < C::CW; ++i) b[i] = __ldg(a.b1 + c0 + i);
            for (int t = 0; t < T; ++t) {
                const int ai = (int)(acc_use % NACC);
                mbar_wait_sleep(&acc_full[ai], (uint32_t)((acc_use / NACC) & 1));
                fence_after();
                uint32_t r[C::CW];
                tmem_ld_cols<C::CW>(lane_base + C::ACC_COL + ai * MID + c0, r);
                tmem_ld_wait();
                fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[ai]);
                ++acc_use;
#pragma unroll
                for (int i = 0; i < C::CW / 8; ++i) {
                    uint32_t p[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        p[e] = pack_h2(fmaxf(__uint_as_float(r[i * 8 + 2 * e]) + b[i * 8 + 2 * e], 0.f),
                                       fmaxf(__uint_as_float(r[i * 8 + 2 * e + 1]) + b[i * 8 + 2 * e + 1], 0.f));
                    tmem_st4(lane_base + C::X1_COL + t * (MID / 2) + c0 / 2 + i * 4, p);
                }
            }
            tmem_st_wait();
            fence_before();
            // conv1's ring is dead (its last MMA completed before acc_full fired): zero the border rows of P
            for (int i = tid; i < C::NCH * 2 * W; i += kComputeThreads) {
                const int ch = i / (2 * W), rem = i - ch * 2 * W, top = rem / W, xx = rem - top * W;
                // in a cluster the inner halo rows belong to the neighbour strips (they write them every level)
                if (CL > 1 && (top ? strip != CL - 1 : strip != 0)) continue;
                *reinterpret_cast<uint4*>(s_region + (size_t)ch * C::PLANE + ((top ? C::SR + 1 : 0) * W + xx) * 16) =
                    make_uint4(0u, 0u, 0u, 0u);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&act_ready);
            if (CL > 1) cluster_arrive();         // this CTA's conv1 ring (= the P planes) may now be written remotely
        }
        // ---- levels ----
        for (int lvl = 0; lvl < 10; ++lvl) {
            int s, j;
            level_sj(lvl, s, j);
            const bool tail = j == s;
            const uint8_t* sd = s_dw0 + (lvl & 1) * DWB;
            const __half* dww = reinterpret_cast<const __half*>(sd);                    // [9][MID]
            const float* bpw = reinterpret_cast<const float*>(sd + 9 * MID * 2);        // [MID]
            const float* bdw = bpw + MID;                                               // [MID]
            __half* tail_base = s == 0 ? a.tails[0] : s == 1 ? a.tails[1] : s == 2 ? a.tails[2] : a.tails[3];
            mbar_wait_sleep(&dw_full[lvl & 1], (uint32_t)((lvl >> 1) & 1));
            // pointwise epilogue: acc + bias -> fp16 (zero outside the image) -> P planes.  The TMEM load of tile
            // t + 1 is issued before tile t is converted (two register buffers).
            {
                uint32_t rbuf[2][C::CW];
                auto fetch = [&](int t, uint32_t* r) {
                    const int ai = (int)((acc_use + t) % NACC);
                    mbar_wait_sleep(&acc_full[ai], (uint32_t)(((acc_use + t) / NACC) & 1));
                    fence_after();
                    tmem_ld_cols<C::CW>(lane_base + C::ACC_COL + ai * MID + c0, r);
                };
                if (tid == 0) OSB_STAMP(16 + lvl * 16 + 4);
                if (CL > 1) cluster_wait();        // every strip of the crop is done reading its planes (previous level)
                fetch(0, rbuf[0]);
                if (tid == 0) OSB_STAMP(16 + lvl * 16 + 5);
#pragma unroll
                for (int t = 0; t < T; ++t) {
                    uint32_t* r = rbuf[t & 1];
                    tmem_ld_wait();
                    fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&acc_empty[(int)((acc_use + t) % NACC)]);
                    if (t + 1 < T && NACC > 1) fetch(t + 1, rbuf[(t + 1) & 1]);
                    const int yloc = run * T + t, y = y0 + yloc;
                    const bool inside = y >= 0 && y < a.H;
#pragma unroll
                    for (int i = 0; i < C::CW / 8; ++i) {
                        const float4 ba = *reinterpret_cast<const float4*>(bpw + c0 + i * 8);
                        const float4 bb = *reinterpret_cast<const float4*>(bpw + c0 + i * 8 + 4);
                        uint32_t p[4];
                        p[0] = pack_h2(__uint_as_float(r[i * 8 + 0]) + ba.x, __uint_as_float(r[i * 8 + 1]) + ba.y);
                        p[1] = pack_h2(__uint_as_float(r[i * 8 + 2]) + ba.z, __uint_as_float(r[i * 8 + 3]) + ba.w);
                        p[2] = pack_h2(__uint_as_float(r[i * 8 + 4]) + bb.x, __uint_as_float(r[i * 8 + 5]) + bb.y);
                        p[3] = pack_h2(__uint_as_float(r[i * 8 + 6]) + bb.z, __uint_as_float(r[i * 8 + 7]) + bb.w);
                        const uint4 pv = inside ? make_uint4(p[0], p[1], p[2], p[3]) : make_uint4(0u, 0u, 0u, 0u);
                        uint8_t* dst = s_region + (size_t)(c0 / 8 + i) * C::PLANE + ((yloc + 1) * W + x) * 16;
                        *reinterpret_cast<uint4*>(dst) = pv;
                        if (CL > 1) {
                            // first / last row of the strip -> halo row of the strip above / below
                            if (yloc == 0 && strip > 0)
                                st_cluster_v4(smem_u32(dst) + (uint32_t)(C::SR * W * 16), (uint32_t)(strip - 1), pv);
                            if (yloc == C::SR - 1 && strip < CL - 1)
                                st_cluster_v4(smem_u32(dst) - (uint32_t)(C::SR * W * 16), (uint32_t)(strip + 1), pv);
                        }
                    }
                    if (t + 1 < T && NACC == 1) fetch(t + 1, rbuf[(t + 1) & 1]);
                }
                acc_use += T;
            }
            if (tid == 0) OSB_STAMP(16 + lvl * 16 + 6);
            if (CL > 1) { cluster_arrive(); cluster_wait(); }      // all rows (own and halo) of the crop are in place
            else named_bar_sync(1, kComputeThreads);
            if (tid == 0) OSB_STAMP(16 + lvl * 16 + 7);
            // depthwise 3x3 + bias + ReLU over the vertical run of this thread
            float gsum[8];
#pragma unroll
            for (int i = 0; i < C::CW / 8; ++i) {
                const int c8 = c0 / 8 + i;
                const uint8_t* plane = s_region + (size_t)c8 * C::PLANE + (size_t)(run * T) * W * 16;
                __half2 wv[9][4];
#pragma unroll
                for (int k = 0; k < 9; ++k) {
                    const uint4 u = *reinterpret_cast<const uint4*>(dww + k * MID + c8 * 8);
                    wv[k][0] = *reinterpret_cast<const __half2*>(&u.x); wv[k][1] = *reinterpret_cast<const __half2*>(&u.y);
                    wv[k][2] = *reinterpret_cast<const __half2*>(&u.z); wv[k][3] = *reinterpret_cast<const __half2*>(&u.w);
                }
                __half2 bias2[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) bias2[e] = __floats2half2_rn(bdw[c8 * 8 + 2 * e], bdw[c8 * 8 + 2 * e + 1]);
                if (tail) {
#pragma unroll
                    for (int e = 0; e < 8; ++e) gsum[e] = 0.f;
                }
                uint4 win[4][3];
                auto load_row = [&](int r, uint4 (&dst)[3]) {
                    const uint8_t* rp = plane + (size_t)r * W * 16;
                    dst[1] = *reinterpret_cast<const uint4*>(rp + x * 16);
                    dst[0] = x > 0 ? *reinterpret_cast<const uint4*>(rp + (x - 1) * 16) : make_uint4(0u, 0u, 0u, 0u);
                    dst[2] = x < W - 1 ? *reinterpret_cast<const uint4*>(rp + (x + 1) * 16) : make_uint4(0u, 0u, 0u, 0u);
                };
                load_row(0, win[0]);
                load_row(1, win[1]);
                load_row(2, win[2]);
#pragma unroll
                for (int t = 0; t < T; ++t) {
                    if (t + 1 < T) load_row(t + 3, win[(t + 3) & 3]);      // next output's new row, before the math
                    __half2 o[4] = {bias2[0], bias2[1], bias2[2], bias2[3]};
#pragma unroll
                    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                        for (int dx = 0; dx < 3; ++dx) {
                            const uint4& v = win[(t + dy) & 3][dx];
                            o[0] = __hfma2(wv[dy * 3 + dx][0], *reinterpret_cast<const __half2*>(&v.x), o[0]);
                            o[1] = __hfma2(wv[dy * 3 + dx][1], *reinterpret_cast<const __half2*>(&v.y), o[1]);
                            o[2] = __hfma2(wv[dy * 3 + dx][2], *reinterpret_cast<const __half2*>(&v.z), o[2]);
                            o[3] = __hfma2(wv[dy * 3 + dx][3], *reinterpret_cast<const __half2*>(&v.w), o[3]);
                        }
                    const __half2 z = __float2half2_rn(0.f);
                    uint32_t p[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        o[e] = __hmax2(o[e], z);
                        p[e] = *reinterpret_cast<const uint32_t*>(&o[e]);
                    }
                    if (!tail) {
                        tmem_st4(lane_base + C::ACT_COL + t * (MID / 2) + c8 * 4, p);
                    } else {
                        const int yloc = run * T + t, y = y0 + yloc;
                        if (yloc >= a.halo && yloc < a.halo + a.R && y < a.H) {
                            // chunk-planar tail [crop][chunk][y][x][8]: a warp writes whole 16-byte-per-pixel rows
                            *reinterpret_cast<uint4*>(tail_base + ((((size_t)crop * C::NCH + c8) * a.H + y) * W + x) * 8) =
                                make_uint4(p[0], p[1], p[2], p[3]);
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float2 f = __half22float2(o[e]);
                                gsum[2 * e] += f.x;
                                gsum[2 * e + 1] += f.y;
                            }
                        }
                    }
                }
                if (tail) {
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const float v = warp_sum(gsum[e]);
                        if (lane == 0) s_gap[warp * C::CW + i * 8 + e] = v;
                    }
                }
            }
            if (tail) {
                named_bar_sync(1, kComputeThreads);
                if (tid < MID) {
                    // channel tid belongs to channel quarter g2; sum its four lane quarters in a fixed order
                    const int g2 = tid / C::CW, cc = tid - g2 * C::CW;
                    float v = 0.f;
#pragma unroll
                    for (int q2 = 0; q2 < 4; ++q2) v += s_gap[(g2 * 4 + q2) * C::CW + cc];
                    a.gap_part[(((size_t)crop * a.strips + strip) * 4 + s) * MID + tid] = v;
                }
            } else {
                tmem_st_wait();
            }
            if (tid == 0) OSB_STAMP(16 + lvl * 16 + 8);
            fence_before();
            __syncwarp();
            if (lane == 0 && lvl + 1 < 10) mbar_arrive(&act_ready);
            if (CL > 1) cluster_arrive();
        }
        if (CL > 1) cluster_wait();
    }
    fence_before();
    __syncthreads();
    if (warp == kComputeWarps) tmem_dealloc<C::TMEM_COLS>(tmem);
}

template <int W, int MID, int T, int NACC, int NS, int NW, int CL, int MINB = 1>
int launch_streams(const FmOsbStreams* d, cudaStream_t st) {
    using C = SCfg<W, MID, T, NACC, NS, NW, CL>;
    OsbStreamsArgs a;
    a.H = d->h; a.n_crops = d->n; a.cin = d->cin;
    if (CL > 1) { a.R = C::SR; a.halo = 0; a.strips = CL; }
    else if (d->h == C::SR) { a.R = C::SR; a.halo = 0; a.strips = 1; }
    else { a.halo = 4; a.R = C::SR - 8; a.strips = d->h / a.R; }
    if (a.R <= 0 || a.strips * a.R != d->h) { fm_set_last_error("fm_osb_streams: strip plan"); return FM_ERR_ARG; }
    a.w1 = (const uint8_t*)d->w1; a.b1 = d->b1; a.pw = (const uint8_t*)d->pw; a.dw = (const uint8_t*)d->dw;
    for (int i = 0; i < 4; ++i) a.tails[i] = (__half*)d->tails[i];
    a.gap_part = d->gap_part;
    CUtensorMap map;
    int rc = fm_make_tmap_f16_3d(&map, d->x, (uint64_t)d->cin, (uint64_t)d->h * W, (uint64_t)d->n, (uint64_t)d->cin,
                                 (uint64_t)d->h * W * d->cin, 64, W, 1);
    if (rc) return rc;
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(osb_streams_kernel<W, MID, T, NACC, NS, NW, CL, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             C::SMEM);
        attr = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(d->n * a.strips);
    cfg.blockDim = dim3(C::kThreads);
    cfg.dynamicSmemBytes = C::SMEM;
    cfg.stream = st;
    cudaLaunchAttribute attrs[2];
    int na = 0;
    if (fm_pdl_enabled()) {
        attrs[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attrs[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    if (CL > 1) {
        attrs[na].id = cudaLaunchAttributeClusterDimension;
        attrs[na].val.clusterDim.x = CL; attrs[na].val.clusterDim.y = 1; attrs[na].val.clusterDim.z = 1;
        ++na;
    }
    cfg.attrs = attrs;
    cfg.numAttrs = na;
    cudaError_t e = cudaLaunchKernelEx(&cfg, osb_streams_kernel<W, MID, T, NACC, NS, NW, CL, MINB>, map, a);
    if (e != cudaSuccess) { fm_set_last_error(cudaGetErrorString(e)); return FM_ERR_CUDA; }
    return FM_OK;
}

// =====================================================================================================================
// osb_merge_kernel<MID, NCTA>  ("kernel G"): the second half of an OSBlock in one launch
//     g_s   = sigmoid(W2 relu(W1 mean(tail_s) + b1) + b2)               (unified aggregation gate, s = 0..3)
//     u     = sum_s g_s * tail_s                                        (fp16, never leaves the SM)
//     out   = relu(conv3(u) + b3 + identity)        identity = x (cin == cout)  or  downsample(x) (1x1, cin != cout)
// One CTA = 128 consecutive pixels of one crop x NCTA output channels.  The downsample conv is folded into the same
// accumulation as extra K slices: x tiles arrive by TMA while the compute warps build u in shared memory (tails are
// chunk-planar, so lanes run along the pixels and land swizzled rows without bank conflicts); weight slices of
// [W_down | W_3] stream through the same ring by cp.async.bulk.  Epilogue: TMEM -> fp16 staging tile -> coalesced
// rows (+ identity rows read coalesced) -> ReLU -> NHWC.
// =====================================================================================================================
}  // namespace
int fm_gate_fc4_part(const float* gap_part, int strips, int n, int hw, const float* w1, const float* b1, const float* w2,
                     const float* b2, float* gate, int c, int cr, cudaStream_t s);    // nn_vec.cu
namespace {

struct OsbMergeArgs {
    int n, hw, cin, cout, strips, has_down, cr;
    const __half* tails[4];      // chunk-planar [n][MID / 8][hw][8]
    const float* gap_part;       // [n][strips][4][MID]
    const float* gates;          // [4][n][MID] sigmoid gates (gate_fc4_part_kernel)
    const uint8_t* wimg;         // per N range: (cin / 64 if has_down) + NSLU slices of [NCTA x 128 B]
    const float* bias;           // [cout] = b3 (+ b_down)
    const __half* res;           // identity [n][hw][cout] or NULL
    __half* out;                 // [n][hw][cout]
};

template <int MID, int NCTA>
struct GCfg {
    static constexpr int NSLU = (MID + 63) / 64;
    static constexpr int NS = 2;
    static constexpr int BSL = NCTA * 128;                  // one weight slice
    static constexpr int STAGE = 16384 + BSL;
    static constexpr int RING = NS * STAGE;
    static constexpr int PITCH = NCTA * 2 + 16;             // staging row
    static constexpr int STG = 128 * PITCH;
    static constexpr int REGION = RING > STG ? RING : STG;
    static constexpr int SMEM = REGION + 4 * MID * 4;
    static_assert(NSLU <= NS, "the u slices live in distinct ring stages");
    static constexpr int TMEM_COLS = NCTA <= 128 ? 128 : 256;
    static constexpr int kThreads = 8 * 32 + 32;
};

template <int MID, int NCTA>
__global__ void __launch_bounds__(288, 2)
osb_merge_kernel(const __grid_constant__ CUtensorMap map_x, OsbMergeArgs a) {
    using C = GCfg<MID, NCTA>;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* s_region = smem;                                    // ring (A slice | weight slice), later the staging tile
    float* s_gate = reinterpret_cast<float*>(smem + C::REGION);  // [4][MID]
    __shared__ uint64_t ring_full[C::NS], ring_empty[C::NS], u_ready, acc_full, u_free[C::NSLU];
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    fm_pdl_trigger();
    const int tiles_per = a.hw >> 7;
    const int crop = blockIdx.x / tiles_per, p0 = (blockIdx.x - crop * tiles_per) << 7;
    const int n0 = blockIdx.y * NCTA;
    const int nslx = a.has_down ? (a.cin >> 6) : 0;
    const int iters = nslx + C::NSLU;
    if (tid == 0) {
        if (smem_u32(smem) & 1023u) __trap();
        for (int i = 0; i < C::NS; ++i) { mbar_init(&ring_full[i], 1); mbar_init(&ring_empty[i], 1); }
        mbar_init(&u_ready, 8);
        mbar_init(&acc_full, 1);
        for (int i = 0; i < C::NSLU; ++i) mbar_init(&u_free[i], 1);
        mbar_fence_init();
    }
    if (warp == 8) {
        tmem_alloc<C::TMEM_COLS>(&s_tmem);
        if (lane == 0 && a.has_down) tma_prefetch_desc(&map_x);
    }
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = s_tmem;
    fm_pdl_wait();

    if (warp == 8) {
        // ------------------------------------------- control warp --------------------------------------------
        const bool leader = lane == 0;
        const uint32_t idesc = idesc_f16(NCTA);
        const uint8_t* wimg = a.wimg + (size_t)blockIdx.y * iters * C::BSL;
        auto issue = [&](int i) {
            const int st = i % C::NS;
            if (i >= C::NS) mbar_wait(&ring_empty[st], (uint32_t)((i / C::NS - 1) & 1));
            uint8_t* sa = s_region + (size_t)st * C::STAGE;
            if (leader) {
                const bool isx = i < nslx;
                // the stage's A half now belongs to the compute warps (they build slice i - nslx of u in place).  A
                // single-use barrier: a warp that is not in lock step with the ring cannot use its parity waits.
                if (!isx) mbar_arrive(&u_free[i - nslx]);
                mbar_expect_tx(&ring_full[st], (isx ? 16384u : 0u) + (uint32_t)C::BSL);
                if (isx) tma_load_3d(sa, &map_x, &ring_full[st], i * 64, p0, crop);
                bulk_load(sa + 16384, wimg + (size_t)i * C::BSL, C::BSL, &ring_full[st]);
            }
        };
        for (int i = 0; i < C::NS - 1 && i < iters; ++i) issue(i);
        for (int i = 0; i < iters; ++i) {
            if (i + C::NS - 1 < iters) issue(i + C::NS - 1);
            const int st = i % C::NS;
            mbar_wait(&ring_full[st], (uint32_t)((i / C::NS) & 1));
            const bool isx = i < nslx;
            if (!isx && i == nslx) mbar_wait(&u_ready, 0);
            fence_after();
            const uint32_t sa = smem_u32(s_region + (size_t)st * C::STAGE), sb = sa + 16384;   // u slices: built in place
            const int ksteps = isx ? 4 : min(4, (MID - (i - nslx) * 64) / 16);
            if (leader) {
                for (int k = 0; k < ksteps; ++k)
                    mma_ss(tmem, smem_desc_sw128(sa + k * 32), smem_desc_sw128(sb + k * 32), idesc, (i > 0 || k > 0) ? 1u : 0u);
                commit(&ring_empty[st]);
                if (i == iters - 1) commit(&acc_full);
            }
            __syncwarp();
        }
    } else {
        // ------------------------------------------- compute warps -------------------------------------------
        constexpr int NCH = MID / 8;
        // gates of this crop (computed once per crop by the small FC kernel launched before this one)
        for (int i = tid; i < 4 * MID; i += 256) {
            const int st = i / MID, c = i - st * MID;
            s_gate[i] = a.gates[((size_t)st * a.n + crop) * MID + c];
        }
        named_bar_sync(1, 256);
        // u tile: item = (chunk, pixel); lanes run along the pixels (coalesced planar reads, conflict-free swizzled
        // writes).  Slice k of u is the A operand of ring iteration nslx + k and is written straight into that stage.
#pragma unroll
        for (int k = 0; k < C::NSLU; ++k) mbar_wait_sleep(&u_free[k], 0);
        for (int i = tid; i < NCH * 128; i += 256) {
            const int c8 = i >> 7, px = i & 127;
            const size_t off = (((size_t)crop * NCH + c8) * a.hw + p0 + px) * 8;
            float o[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) o[q] = 0.f;
#pragma unroll
            for (int st = 0; st < 4; ++st) {
                const uint4 v = __ldg(reinterpret_cast<const uint4*>(a.tails[st] + off));
                const __half2* h = reinterpret_cast<const __half2*>(&v);
                const float* g = s_gate + st * MID + c8 * 8;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float2 f = __half22float2(h[q]);
                    o[2 * q] += f.x * g[2 * q];
                    o[2 * q + 1] += f.y * g[2 * q + 1];
                }
            }
            const int sl = c8 >> 3, cc = c8 & 7;
            *reinterpret_cast<uint4*>(s_region + (size_t)((nslx + sl) % C::NS) * C::STAGE + px * 128 + ((cc ^ (px & 7)) << 4)) =
                make_uint4(pack_h2(o[0], o[1]), pack_h2(o[2], o[3]), pack_h2(o[4], o[5]), pack_h2(o[6], o[7]));
        }
        fence_async_smem();            // generic-proxy writes -> visible to the tensor core
        __syncwarp();
        if (lane == 0) mbar_arrive(&u_ready);
        // epilogue
        mbar_wait_sleep(&acc_full, 0);
        fence_after();
        const int q = warp & 3, hsel = warp >> 2;
        const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
        const int row = q * 32 + lane;
        constexpr int HALF = NCTA / 2;
#pragma unroll 1
        for (int j0 = 0; j0 < HALF; j0 += 32) {
            uint32_t r[32];
            tmem_ld32(lane_base + hsel * HALF + j0, r);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 4; ++e)
                *reinterpret_cast<uint4*>(s_region + row * C::PITCH + (hsel * HALF + j0 + e * 8) * 2) = make_uint4(
                    pack_h2(__uint_as_float(r[e * 8 + 0]), __uint_as_float(r[e * 8 + 1])),
                    pack_h2(__uint_as_float(r[e * 8 + 2]), __uint_as_float(r[e * 8 + 3])),
                    pack_h2(__uint_as_float(r[e * 8 + 4]), __uint_as_float(r[e * 8 + 5])),
                    pack_h2(__uint_as_float(r[e * 8 + 6]), __uint_as_float(r[e * 8 + 7])));
        }
        fence_before();
        named_bar_sync(1, 256);
        // coalesced pass: lanes along the channels
        constexpr int CPR = NCTA / 8;                       // 16-byte chunks per row
        for (int i = tid; i < 128 * CPR; i += 256) {
            const int rr = i / CPR, ch = i - rr * CPR;
            const int n = n0 + ch * 8;
            const uint4 pk = *reinterpret_cast<const uint4*>(s_region + rr * C::PITCH + ch * 16);
            const __half2* ph = reinterpret_cast<const __half2*>(&pk);
            const size_t gofs = ((size_t)crop * a.hw + p0 + rr) * a.cout + n;
            const float4 ba = __ldg(reinterpret_cast<const float4*>(a.bias + n));
            const float4 bb = __ldg(reinterpret_cast<const float4*>(a.bias + n + 4));
            float xv[8];
            { const float2 f = __half22float2(ph[0]); xv[0] = f.x + ba.x; xv[1] = f.y + ba.y; }
            { const float2 f = __half22float2(ph[1]); xv[2] = f.x + ba.z; xv[3] = f.y + ba.w; }
            { const float2 f = __half22float2(ph[2]); xv[4] = f.x + bb.x; xv[5] = f.y + bb.y; }
            { const float2 f = __half22float2(ph[3]); xv[6] = f.x + bb.z; xv[7] = f.y + bb.w; }
            if (a.res) {
                const uint4 rv = __ldg(reinterpret_cast<const uint4*>(a.res + gofs));
                const __half2* rh = reinterpret_cast<const __half2*>(&rv);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float2 f = __half22float2(rh[e]);
                    xv[2 * e] += f.x;
                    xv[2 * e + 1] += f.y;
                }
            }
            *reinterpret_cast<uint4*>(a.out + gofs) =
                make_uint4(pack_h2(fmaxf(xv[0], 0.f), fmaxf(xv[1], 0.f)), pack_h2(fmaxf(xv[2], 0.f), fmaxf(xv[3], 0.f)),
                           pack_h2(fmaxf(xv[4], 0.f), fmaxf(xv[5], 0.f)), pack_h2(fmaxf(xv[6], 0.f), fmaxf(xv[7], 0.f)));
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 8) tmem_dealloc<C::TMEM_COLS>(tmem);
}

template <int MID, int NCTA>
int launch_merge(const FmOsbMerge* d, cudaStream_t st) {
    using C = GCfg<MID, NCTA>;
    OsbMergeArgs a;
    a.n = d->n; a.hw = d->hw; a.cin = d->cin; a.cout = d->cout; a.strips = d->strips; a.cr = d->cr;
    a.has_down = d->x != nullptr;
    for (int i = 0; i < 4; ++i) a.tails[i] = (const __half*)d->tails[i];
    a.gap_part = d->gap_part;
    a.gates = d->gate_scratch;
    a.wimg = (const uint8_t*)d->wimg; a.bias = d->bias; a.res = (const __half*)d->res; a.out = (__half*)d->out;
    CUtensorMap map;
    memset(&map, 0, sizeof map);
    if (a.has_down) {
        int rc = fm_make_tmap_f16_3d(&map, d->x, (uint64_t)d->cin, (uint64_t)d->hw, (uint64_t)d->n, (uint64_t)d->cin,
                                     (uint64_t)d->hw * d->cin, 64, 128, 1);
        if (rc) return rc;
    }
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(osb_merge_kernel<MID, NCTA>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
        attr = true;
    }
    dim3 grid(d->n * (d->hw >> 7), d->cout / NCTA);
    cudaError_t e = fm_launch_pdl(osb_merge_kernel<MID, NCTA>, grid, dim3(C::kThreads), (size_t)C::SMEM, st, map, a);
    if (e != cudaSuccess) { fm_set_last_error(cudaGetErrorString(e)); return FM_ERR_CUDA; }
    return FM_OK;
}

}  // namespace

extern "C" int fm_osb_merge_ncta(int mid, int cout) {
    if (mid == 64 && cout % 256 == 0) return 256;
    if (mid == 96 && cout % 192 == 0) return 192;
    if (mid == 128 && cout % 256 == 0) return 256;
    return 0;
}

extern "C" int fm_osb_merge(const FmOsbMerge* d, void* stream) {
    FM_REQUIRE(d != nullptr, "fm_osb_merge: desc is NULL");
    FM_REQUIRE(fm_osb_merge_ncta(d->mid, d->cout) > 0, "fm_osb_merge: unsupported (mid, cout)");
    FM_REQUIRE(d->hw % 128 == 0 && d->cr >= 1 && d->cr <= 8 && d->strips >= 1, "fm_osb_merge: hw / cr / strips");
    FM_REQUIRE((d->x != nullptr) != (d->res != nullptr), "fm_osb_merge: exactly one of x (downsample) and res (identity)");
    FM_REQUIRE(d->x == nullptr || (d->cin % 64 == 0 && d->cin >= 64), "fm_osb_merge: cin must be a multiple of 64");
    FM_REQUIRE(d->gate_scratch != nullptr, "fm_osb_merge: gate_scratch (4 * n * mid floats) is NULL");
    if (d->n <= 0) return FM_OK;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = fm_gate_fc4_part(d->gap_part, d->strips, d->n, d->hw, d->gw1, d->gb1, d->gw2, d->gb2, d->gate_scratch, d->mid,
                              d->cr, st);
    if (rc) return rc;
    if (d->mid == 64) rc = launch_merge<64, 256>(d, st);
    else if (d->mid == 96) rc = launch_merge<96, 192>(d, st);
    else rc = launch_merge<128, 256>(d, st);
    if (rc) return rc;
    FM_CHECK_LAUNCH("fm_osb_merge");
    return FM_OK;
}

namespace {
}  // namespace

extern "C" int fm_osb_set_debug(void* dbg) {     // debugging aid, not part of the public header
    long long* p = (long long*)dbg;
    cudaMemcpyToSymbol(g_osb_dbg, &p, sizeof(p));
    return FM_OK;
}

extern "C" int fm_osb_streams_strips(int h, int w, int mid) {
    if (w == 32 && mid == 64 && h == 64) {
        const char* e = getenv("FM_OSB_CLUSTER");
        if (!(e && e[0] == '0')) return 4;          // one 4-CTA cluster per crop
    }
    if (w == 32 && mid == 64) return h == 16 ? 1 : (h % 8 == 0 && h > 16 ? h / 8 : 0);   // 16 rows = one strip, no halo
    if (w == 16 && mid == 96 && h == 32) {
        const char* e = getenv("FM_OSB_CLUSTER");
        return (e && e[0] == '0') ? 1 : 2;          // two-CTA cluster per crop
    }
    if (w == 8 && mid == 128) return h == 16 ? 1 : 0;
    return 0;
}

extern "C" int fm_osb_streams(const FmOsbStreams* d, void* stream) {
    FM_REQUIRE(d != nullptr, "fm_osb_streams: desc is NULL");
    FM_REQUIRE(fm_osb_streams_strips(d->h, d->w, d->mid) > 0, "fm_osb_streams: unsupported stage geometry");
    FM_REQUIRE(d->cin % 64 == 0 && d->cin >= 64, "fm_osb_streams: cin must be a multiple of 64");
    if (d->n <= 0) return FM_OK;
    cudaStream_t st = (cudaStream_t)stream;
    int rc;
    static int nw = -1;            // FM_OSB_WARPS=16: sixteen compute warps for stage 1 (A/B timing)
    if (nw < 0) { const char* e = getenv("FM_OSB_WARPS"); nw = (e && atoi(e) == 16) ? 16 : 8; }
    static int cl = -1;            // FM_OSB_CLUSTER=0: strips with a recomputed 4-row halo instead of the 4-CTA cluster
    if (cl < 0) { const char* e = getenv("FM_OSB_CLUSTER"); cl = (e && e[0] == '0') ? 0 : 1; }
    static int s3b = -1;           // FM_OSB_S3_BLOCKS=1: one stage-3 CTA per SM (default two)
    if (s3b < 0) { const char* e = getenv("FM_OSB_S3_BLOCKS"); s3b = (e && e[0] == '1') ? 1 : 2; }
    if (d->w == 32 && d->h == 64 && cl && nw == 16) rc = launch_streams<32, 64, 4, 4, 3, 16, 4>(d, st);
    else if (d->w == 32 && d->h == 64 && cl) rc = launch_streams<32, 64, 4, 4, 3, 8, 4>(d, st);
    else if (d->w == 32) rc = launch_streams<32, 64, 4, 4, 3, 8, 1>(d, st);
    else if (d->w == 16 && cl) rc = launch_streams<16, 96, 2, 2, 3, 8, 2>(d, st);       // two 16-row strips per crop
    else if (d->w == 16) rc = launch_streams<16, 96, 4, 1, 3, 8, 1>(d, st);
    else if (s3b == 2) rc = launch_streams<8, 128, 1, 1, 2, 8, 1, 2>(d, st);
    else rc = launch_streams<8, 128, 1, 1, 2, 8, 1>(d, st);
    if (rc) return rc;
    FM_CHECK_LAUNCH("fm_osb_streams");
    return FM_OK;
}
