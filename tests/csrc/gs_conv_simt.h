/* Test-only CUDA-core twins of gs_conv2d_{fwd,dgrad,wgrad} (include/gaiaseg_b200.h): one thread per output, same
 * arguments and math contract; dgrad reads w_krsc.  Exported by tests/libgaiaseg_simt.so only. */
#pragma once
#include "../../include/gaiaseg_b200.h"
#ifdef __cplusplus
extern "C" {
#endif
int gs_conv2d_fwd_simt(const gs_conv_geom* g, const void* x, const void* w_krsc, void* y, const float* scale,
                       const float* shift, const void* residual, int32_t res_ld, int32_t flags, double* stats,
                       void* stream);
int gs_conv2d_dgrad_simt(const gs_conv_geom* g, const void* dy, const void* w_krsc, void* dx, const void* residual,
                         int32_t res_ld, void* stream);
int gs_conv2d_wgrad_simt(const gs_conv_geom* g, const void* x, const void* dy, float* dw_krsc, void* stream);
#ifdef __cplusplus
}
#endif
