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

template <int W, int MID, int T, int NACC, int NS, int NW>
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

template <int W, int MID, int T, int NACC, int NS, int NW>
__global__ void __launch_bounds__(NW * 32 + 32, 1)
osb_streams_kernel(const __grid_constant__ CUtensorMap map_x, OsbStreamsArgs a) {
    using C = SCfg<W, MID, T, NACC, NS, NW>;
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
    const int crop = blockIdx.x / a.strips, strip = blockIdx.x - crop * a.strips;
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
        // =========================================== control thread ===========================================
        if (lane == 0) {
            // weights of level 0 (constants: no dependence on earlier kernels)
            mbar_expect_tx(&pw_full, C::PW_BYTES);
            bulk_load(s_pw, a.pw, C::PW_BYTES, &pw_full);
            mbar_expect_tx(&dw_full[0], C::DW_BYTES);
            bulk_load(s_dw0, a.dw, C::DW_BYTES, &dw_full[0]);
            const uint32_t idesc = idesc_f16(MID);
            uint32_t acc_use = 0;                                      // accumulator ring uses so far
            // ---- conv1: ring of (A tile slice by TMA, weight slice by bulk copy) ----
            const int iters = T * nsl1;
            auto issue = [&](int i) {
                const int s = i % NS, t = i / nsl1, ks = i - t * nsl1;
                if (i >= NS) mbar_wait(&ring_empty[s], (uint32_t)((i / NS - 1) & 1));
                uint8_t* sa = s_region + (size_t)s * C::STAGE;
                mbar_expect_tx(&ring_full[s], C::STAGE);
#pragma unroll
                for (int rr = 0; rr < C::TROWS; ++rr) {
                    const int y = y0 + rr * T + t;
                    tma_load_3d(sa + rr * W * 128, &map_x, &ring_full[s], ks * 64, y * W, crop);
                }
                bulk_load(sa + 16384, a.w1 + (size_t)ks * MID * 128, MID * 128, &ring_full[s]);
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
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    mma_ss(tmem + C::ACC_COL + ai * MID, smem_desc_sw128(sa + k * 32), smem_desc_sw128(sb + k * 32),
                           idesc, (ks > 0 || k > 0) ? 1u : 0u);
                commit(&ring_empty[s]);
                if (ks == nsl1 - 1) { commit(&acc_full[ai]); ++acc_use; }
            }
            OSB_STAMP(1);
            // ---- the ten pointwise convs: A from TMEM ----
            uint64_t bdesc_pw[MID / 16];
#pragma unroll
            for (int k = 0; k < MID / 16; ++k)
                bdesc_pw[k] = smem_desc_sw128(smem_u32(s_pw) + (k >> 2) * MID * 128 + (k & 3) * 32);
            for (int lvl = 0; lvl < 10; ++lvl) {
                int s, j;
                level_sj(lvl, s, j);
                mbar_wait(&pw_full, (uint32_t)(lvl & 1));
                OSB_STAMP(16 + lvl * 16 + 0);
                mbar_wait(&act_ready, (uint32_t)(lvl & 1));
                OSB_STAMP(16 + lvl * 16 + 1);
                fence_after();
                const uint32_t a_base = tmem + (j == 0 ? C::X1_COL : C::ACT_COL);
                for (int t = 0; t < T; ++t) {
                    const int ai = (int)(acc_use % NACC);
                    if (acc_use >= NACC) mbar_wait(&acc_empty[ai], (uint32_t)((acc_use / NACC - 1) & 1));
                    fence_after();
#pragma unroll
                    for (int k = 0; k < MID / 16; ++k)
                        mma_ts(tmem + C::ACC_COL + ai * MID, a_base + t * (MID / 2) + k * 8, bdesc_pw[k], idesc,
                               k > 0 ? 1u : 0u);
                    commit(&acc_full[ai]);
                    ++acc_use;
                }
                commit(&pw_empty);
                OSB_STAMP(16 + lvl * 16 + 2);
                if (lvl + 1 < 10) {
                    // the depthwise blob of level lvl - 1 is dead (act_ready of this level was its last reader)
                    uint8_t* sd = s_dw0 + ((lvl + 1) & 1) * DWB;
                    mbar_expect_tx(&dw_full[(lvl + 1) & 1], C::DW_BYTES);
                    bulk_load(sd, a.dw + (size_t)(lvl + 1) * C::DW_BYTES, C::DW_BYTES, &dw_full[(lvl + 1) & 1]);
                    mbar_wait(&pw_empty, (uint32_t)(lvl & 1));           // this level's MMAs have read s_pw
                    OSB_STAMP(16 + lvl * 16 + 3);
                    mbar_expect_tx(&pw_full, C::PW_BYTES);
                    bulk_load(s_pw, a.pw + (size_t)(lvl + 1) * C::PW_BYTES, C::PW_BYTES, &pw_full);
                }
            }
        }
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
                *reinterpret_cast<uint4*>(s_region + (size_t)ch * C::PLANE + ((top ? C::SR + 1 : 0) * W + xx) * 16) =
                    make_uint4(0u, 0u, 0u, 0u);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&act_ready);
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
                        *reinterpret_cast<uint4*>(s_region + (size_t)(c0 / 8 + i) * C::PLANE + ((yloc + 1) * W + x) * 16) =
                            inside ? make_uint4(p[0], p[1], p[2], p[3]) : make_uint4(0u, 0u, 0u, 0u);
                    }
                    if (t + 1 < T && NACC == 1) fetch(t + 1, rbuf[(t + 1) & 1]);
                }
                acc_use += T;
            }
            if (tid == 0) OSB_STAMP(16 + lvl * 16 + 6);
            named_bar_sync(1, kComputeThreads);
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
        }
    }
    fence_before();
    __syncthreads();
    if (warp == kComputeWarps) tmem_dealloc<C::TMEM_COLS>(tmem);
}

template <int W, int MID, int T, int NACC, int NS, int NW>
int launch_streams(const FmOsbStreams* d, cudaStream_t st) {
    using C = SCfg<W, MID, T, NACC, NS, NW>;
    OsbStreamsArgs a;
    a.H = d->h; a.n_crops = d->n; a.cin = d->cin;
    if (d->h == C::SR) { a.R = C::SR; a.halo = 0; a.strips = 1; }
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
        cudaFuncSetAttribute(osb_streams_kernel<W, MID, T, NACC, NS, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             C::SMEM);
        attr = true;
    }
    cudaError_t e = fm_launch_pdl(osb_streams_kernel<W, MID, T, NACC, NS, NW>, dim3(d->n * a.strips), dim3(C::kThreads),
                                  (size_t)C::SMEM, st, map, a);
    if (e != cudaSuccess) { fm_set_last_error(cudaGetErrorString(e)); return FM_ERR_CUDA; }
    return FM_OK;
}

}  // namespace

extern "C" int fm_osb_set_debug(void* dbg) {     // debugging aid, not part of the public header
    long long* p = (long long*)dbg;
    cudaMemcpyToSymbol(g_osb_dbg, &p, sizeof(p));
    return FM_OK;
}

extern "C" int fm_osb_streams_strips(int h, int w, int mid) {
    if (w == 32 && mid == 64) return h == 16 ? 1 : (h % 8 == 0 && h > 16 ? h / 8 : 0);   // 16 rows = one strip, no halo
    if (w == 16 && mid == 96) return h == 32 ? 1 : 0;
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
    if (d->w == 32) rc = launch_streams<32, 64, 4, 4, 3, 8>(d, st);
    else if (d->w == 16) rc = launch_streams<16, 96, 4, 1, 3, 8>(d, st);
    else rc = launch_streams<8, 128, 1, 1, 2, 8>(d, st);
    if (rc) return rc;
    FM_CHECK_LAUNCH("fm_osb_streams");
    return FM_OK;
}
