// Shared helpers for the fastmot_b200 C-ABI library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>

#define FM_OK 0
#define FM_ERR_CUDA 1
#define FM_ERR_ARG 2
#define FM_ERR_CAPACITY 3

extern "C" void fm_set_last_error(const char* msg);

extern "C" void fm_count_launches(int n);

#define FM_CHECK_LAUNCH(name)                                                      \
    do {                                                                           \
        fm_count_launches(1);                                                      \
        cudaError_t e__ = cudaGetLastError();                                      \
        if (e__ != cudaSuccess) {                                                  \
            char buf__[256];                                                       \
            snprintf(buf__, sizeof buf__, "%s: %s", name, cudaGetErrorString(e__)); \
            fm_set_last_error(buf__);                                              \
            return FM_ERR_CUDA;                                                    \
        }                                                                          \
    } while (0)

#define FM_REQUIRE(cond, msg)                  \
    do {                                       \
        if (!(cond)) {                         \
            fm_set_last_error(msg);            \
            return FM_ERR_ARG;                 \
        }                                      \
    } while (0)

static inline int fm_cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

#define FM_NUM_SMS 148

// round-half-to-even of a double to an integral double (matches Python/Numba round(x, 0)).
__host__ __device__ __forceinline__ double fm_rint(double x) {
#ifdef __CUDA_ARCH__
    return rint(x);
#else
    return nearbyint(x);
#endif
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
