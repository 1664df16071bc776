// Activation functions shared by the tcgen05 conv kernels (conv_tc.cu, conv_tma.cu).  Darknet semantics
// (scripts/yolo2onnx.py:404-470 in the reference): leaky 0.1, mish, swish, logistic, relu.
#pragma once
#include "common.cuh"
#include "../../include/fastmot_b200.h"

static __device__ __forceinline__ float tc_mish(float v) {
    // x * tanh(softplus(x)) with n = e^x:  tanh(log(1+n)) = n(n+2) / (n(n+2) + 2)  -- one ex2, one rcp, no cancellation
    // branch-free: for v >= 20 the clamp gives t = e^20 (e^20 + 2) ~ 2.4e17 and t / (t + 2) rounds to 1.0f, i.e. v.
    // (A `v > 20 ? v : ...` select compiled to a divergent branch per element: eight serial MUFU chains per pixel row.)
    const float n = __expf(fminf(v, 20.f));
    const float t = n * (n + 2.f);
    return v * __fdividef(t, t + 2.f);
}

static __device__ __forceinline__ float tc_act(float v, int act) {
    switch (act) {
        case FM_ACT_LEAKY: return v > 0.f ? v : 0.1f * v;
        case FM_ACT_RELU: return fmaxf(v, 0.f);
        case FM_ACT_MISH: return tc_mish(v);
        case FM_ACT_SWISH: return __fdividef(v, 1.f + __expf(-v));
        case FM_ACT_LOGISTIC: return __fdividef(1.f, 1.f + __expf(-v));
        default: return v;
    }
}

// activation of 8 values with the (warp-uniform) switch hoisted out of the element loop
static __device__ __forceinline__ void tc_act8(float (&v)[8], int act) {
    switch (act) {
        case FM_ACT_LEAKY:
#pragma unroll
            for (int q = 0; q < 8; ++q) v[q] = v[q] > 0.f ? v[q] : 0.1f * v[q];
            break;
        case FM_ACT_RELU:
#pragma unroll
            for (int q = 0; q < 8; ++q) v[q] = fmaxf(v[q], 0.f);
            break;
        case FM_ACT_MISH:
#pragma unroll
            for (int q = 0; q < 8; ++q) v[q] = tc_mish(v[q]);
            break;
        case FM_ACT_SWISH:
#pragma unroll
            for (int q = 0; q < 8; ++q) v[q] = __fdividef(v[q], 1.f + __expf(-v[q]));
            break;
        case FM_ACT_LOGISTIC:
#pragma unroll
            for (int q = 0; q < 8; ++q) v[q] = __fdividef(1.f, 1.f + __expf(-v[q]));
            break;
        default: break;
    }
}

