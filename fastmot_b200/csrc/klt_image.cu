// KLT image pyramid kernels (byte work, HBM/L2-bound): BGR->gray + 0.5x box mean in one pass, 5-tap Gaussian
// pyrDown, int16 Scharr derivatives, 0.1x background image + mask.
//
// Reference: fastmot/flow.py:121-133, 153-154, 187-189 (cv2.cvtColor / cv2.resize) and the pyramid that
// cv2.calcOpticalFlowPyrLK builds internally (flow.py:203-207; OpenCV lkpyramid.cpp: buildOpticalFlowPyramid,
// calcScharrDeriv).  Fixed-point formulas restated in SURVEY.md Appendix C.
#include "common.cuh"
#include "../../include/fastmot_b200.h"

namespace {

__device__ __forceinline__ int gray_of(const unsigned char* p) {
    // OpenCV 4.13 BGR2GRAY, 15-bit fixed point (pinned against cv2.cvtColor: 0 mismatches on 1M random pixels;
    // the 14-bit constants 1868/9617/4899 quoted in SURVEY.md Appendix C differ on 0.2 % of pixels)
    return (p[0] * 3735 + p[1] * 19235 + p[2] * 9798 + 16384) >> 15;
}

// One thread per 2x2 block of the full-resolution frame.
__global__ void __launch_bounds__(256) gray_half_kernel(const unsigned char* __restrict__ frame, int w, int h,
                                                         unsigned char* __restrict__ gray,
                                                         unsigned char* __restrict__ small, int sw, int sh) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;  // small coords
    const int y = blockIdx.y;
    if (x >= sw || y >= sh) return;
    const int x0 = 2 * x, y0 = 2 * y;
    const int x1 = min(x0 + 1, w - 1), y1 = min(y0 + 1, h - 1);
    const unsigned char* r0 = frame + (size_t)y0 * w * 3;
    const unsigned char* r1 = frame + (size_t)y1 * w * 3;
    const int a = gray_of(r0 + x0 * 3), b = gray_of(r0 + x1 * 3);
    const int c = gray_of(r1 + x0 * 3), d = gray_of(r1 + x1 * 3);
    gray[(size_t)y0 * w + x0] = a;
    gray[(size_t)y0 * w + x1] = b;
    gray[(size_t)y1 * w + x0] = c;
    gray[(size_t)y1 * w + x1] = d;
    small[(size_t)y * sw + x] = (a + b + c + d + 2) >> 2;  // cv2.resize INTER_LINEAR at exactly 0.5x
}

__device__ __forceinline__ int reflect101(int p, int n) {
    if (n == 1) return 0;
    while (p < 0 || p >= n) p = p < 0 ? -p : 2 * n - 2 - p;
    return p;
}

// pyrDown: dst(x,y) = (sum_{i,j} g[i] g[j] src(2x+i-2, 2y+j-2) + 128) >> 8, g = [1 4 6 4 1], BORDER_REFLECT_101
__global__ void __launch_bounds__(256) pyr_down_kernel(const unsigned char* __restrict__ src, int sw, int sh,
                                                        unsigned char* __restrict__ dst, int dw, int dh) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= dw || y >= dh) return;
    const int g[5] = {1, 4, 6, 4, 1};
    int acc = 0;
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        const unsigned char* row = src + (size_t)reflect101(2 * y + j - 2, sh) * sw;
        int racc = 0;
#pragma unroll
        for (int i = 0; i < 5; ++i) racc += g[i] * row[reflect101(2 * x + i - 2, sw)];
        acc += g[j] * racc;
    }
    dst[(size_t)y * dw + x] = (acc + 128) >> 8;
}

// calcScharrDeriv: vertical [3 10 3] smoothing / [-1 0 1] diff first, then horizontal; reflect-101 borders.
__global__ void __launch_bounds__(256) scharr_kernel(const unsigned char* __restrict__ src, int w, int h,
                                                      short2* __restrict__ deriv) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= w || y >= h) return;
    const unsigned char* r0 = src + (size_t)reflect101(y - 1, h) * w;
    const unsigned char* r1 = src + (size_t)y * w;
    const unsigned char* r2 = src + (size_t)reflect101(y + 1, h) * w;
    const int xm = reflect101(x - 1, w), xp = reflect101(x + 1, w);
    const int s_m = (r0[xm] + r2[xm]) * 3 + r1[xm] * 10, s_p = (r0[xp] + r2[xp]) * 3 + r1[xp] * 10;
    const int d_m = r2[xm] - r0[xm], d_c = r2[x] - r0[x], d_p = r2[xp] - r0[xp];
    deriv[(size_t)y * w + x] = make_short2((short)(s_p - s_m), (short)((d_p + d_m) * 3 + d_c * 10));
}

// 0.1x background image (cv2.resize INTER_LINEAR at 1/10: mean of the 2x2 block at (10x+4, 10y+4)) and
// nearest-neighbour mask (src pixel (10x, 10y)); mask source is the owner map (>= NO_OWNER means foreground-free).
__global__ void bg_small_kernel(const unsigned char* __restrict__ gray, const int* __restrict__ owner, int w, int h,
                                unsigned char* __restrict__ bg, unsigned char* __restrict__ bg_mask, int bw, int bh,
                                double inv_sx, double inv_sy) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= bw || y >= bh) return;
    // generic cv2 INTER_LINEAR 8-bit fixed point (2048-scale coefficients), exact 2x2 mean when scale = 10
    float fx = (float)((x + 0.5) * inv_sx - 0.5), fy = (float)((y + 0.5) * inv_sy - 0.5);
    int sx = (int)floorf(fx), sy = (int)floorf(fy);
    fx -= sx; fy -= sy;
    if (sx < 0) { fx = 0; sx = 0; }
    if (sx >= w - 1) { fx = 0; sx = w - 1; }
    if (sy < 0) { fy = 0; sy = 0; }
    if (sy >= h - 1) { fy = 0; sy = h - 1; }
    const int a0 = (int)rintf((1.f - fx) * 2048.f), a1 = (int)rintf(fx * 2048.f);
    const int b0 = (int)rintf((1.f - fy) * 2048.f), b1 = (int)rintf(fy * 2048.f);
    const int sx1 = min(sx + 1, w - 1), sy1 = min(sy + 1, h - 1);
    const unsigned char* r0 = gray + (size_t)sy * w;
    const unsigned char* r1 = gray + (size_t)sy1 * w;
    const int h0 = r0[sx] * a0 + r0[sx1] * a1, h1 = r1[sx] * a0 + r1[sx1] * a1;
    bg[(size_t)y * bw + x] = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
    const int nx = min((int)floor(x * inv_sx), w - 1), ny = min((int)floor(y * inv_sy), h - 1);
    bg_mask[(size_t)y * bw + x] = owner[(size_t)ny * w + nx] == FM_NO_OWNER ? 255 : 0;
}

}  // namespace

extern "C" int fm_gray_half(const unsigned char* frame, int w, int h, unsigned char* gray, unsigned char* small,
                            void* stream) {
    const int sw = (w + 1) / 2, sh = (h + 1) / 2;
    FM_REQUIRE(w % 2 == 0 && h % 2 == 0, "fm_gray_half: frame size must be even (0.5x resize = 2x2 mean)");
    dim3 grid(fm_cdiv(sw, 256), sh);
    gray_half_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(frame, w, h, gray, small, sw, sh);
    FM_CHECK_LAUNCH("fm_gray_half");
    return FM_OK;
}

extern "C" int fm_pyr_level(const unsigned char* src, int sw, int sh, unsigned char* dst, void* stream) {
    const int dw = (sw + 1) / 2, dh = (sh + 1) / 2;
    dim3 grid(fm_cdiv(dw, 256), dh);
    pyr_down_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, sw, sh, dst, dw, dh);
    FM_CHECK_LAUNCH("fm_pyr_level");
    return FM_OK;
}

extern "C" int fm_scharr(const unsigned char* src, int w, int h, short* deriv, void* stream) {
    dim3 grid(fm_cdiv(w, 256), h);
    scharr_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, w, h, (short2*)deriv);
    FM_CHECK_LAUNCH("fm_scharr");
    return FM_OK;
}

extern "C" int fm_bg_small(const unsigned char* gray, const int* owner, int w, int h, unsigned char* bg,
                           unsigned char* bg_mask, int bw, int bh, void* stream) {
    dim3 grid(fm_cdiv(bw, 128), bh);
    bg_small_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(gray, owner, w, h, bg, bg_mask, bw, bh,
                                                            (double)w / bw, (double)h / bh);
    FM_CHECK_LAUNCH("fm_bg_small");
    return FM_OK;
}
