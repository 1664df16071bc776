// tcgen05 / TMEM implicit-GEMM convolution (placeholder until the kernel lands in this round).
#include "common.cuh"
#include "../../include/fastmot_b200.h"

extern "C" int fm_conv2d_tc_supported(const FmConvDesc* d) { (void)d; return 0; }

extern "C" int fm_conv2d_tc(const FmConvDesc* d, const void* in, const void* wgt, const float* bias,
                            const void* residual, void* out, void* stream) {
    (void)d; (void)in; (void)wgt; (void)bias; (void)residual; (void)out; (void)stream;
    fm_set_last_error("fm_conv2d_tc: shape not supported by the tcgen05 path");
    return FM_ERR_ARG;
}
