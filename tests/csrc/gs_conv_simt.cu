// gs_conv_simt.cu -- direct (one thread per output) convolution kernels on CUDA cores.
//
// TEST INFRASTRUCTURE, not the product: built into tests/libgaiaseg_simt.so (never into libgaiaseg_b200.so) and loaded
// only by tests/gs_checks.py / tools/gpu_diag.py to triage the tcgen05 implicit-GEMM kernels on the GPU box
// (same C-ABI arguments, suffix _simt: separates "descriptor / pipeline bug" from "host-side geometry bug").  Same math
// contract: bf16 operands, fp32 accumulation, identical epilogue order.
#include <cuda_bf16.h>

#include "../../include/gaiaseg_b200.h"
#include "../../gaia_seg_b200/csrc/gs_host.h"
#include "gs_conv_simt.h"

namespace gs {

__global__ void conv_fwd_simt_kernel(gs_conv_geom g, const __nv_bfloat16* __restrict__ x,
                                     const __nv_bfloat16* __restrict__ w, void* __restrict__ y,
                                     const float* __restrict__ scale, const float* __restrict__ shift,
                                     const __nv_bfloat16* __restrict__ res, int res_ld, int flags,
                                     double* __restrict__ stats) {
    const long long total = (long long)g.N * g.Ho * g.Wo * g.Co;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int co = (int)(i % g.Co);
        long long t = i / g.Co;
        const int wo = (int)(t % g.Wo); t /= g.Wo;
        const int ho = (int)(t % g.Ho);
        const int n = (int)(t / g.Ho);
        float acc = 0.f;
        for (int r = 0; r < g.kh; ++r) {
            const int h = ho * g.stride - g.pad + r * g.dil;
            if (h < 0 || h >= g.H) continue;
            for (int s = 0; s < g.kw; ++s) {
                const int ww = wo * g.stride - g.pad + s * g.dil;
                if (ww < 0 || ww >= g.W) continue;
                const __nv_bfloat16* xp = x + ((long long)(n * g.H + h) * g.W + ww) * g.x_ld;
                const __nv_bfloat16* wp = w + (((long long)co * g.kh + r) * g.kw + s) * g.Ci_max;
                for (int ci = 0; ci < g.Ci; ++ci) acc = fmaf(__bfloat162float(xp[ci]), __bfloat162float(wp[ci]), acc);
            }
        }
        const long long pix = (long long)(n * g.Ho + ho) * g.Wo + wo;
        if (scale) acc *= scale[co];
        if (shift) acc += shift[co];
        if (res) acc += __bfloat162float(res[pix * res_ld + co]);
        if (flags & GS_EPI_RELU) acc = fmaxf(acc, 0.f);
        if (flags & GS_EPI_OUT_F32) {
            reinterpret_cast<float*>(y)[pix * g.y_ld + co] = acc;
        } else {
            const __nv_bfloat16 o = __float2bfloat16_rn(acc);
            reinterpret_cast<__nv_bfloat16*>(y)[pix * g.y_ld + co] = o;
            if (stats) {
                const double v = (double)__bfloat162float(o);
                atomicAdd(stats + co, v);
                atomicAdd(stats + g.Co + co, v * v);
            }
        }
    }
}

__global__ void conv_dgrad_simt_kernel(gs_conv_geom g, const __nv_bfloat16* __restrict__ dy,
                                       const __nv_bfloat16* __restrict__ w, __nv_bfloat16* __restrict__ dx,
                                       const __nv_bfloat16* __restrict__ res, int res_ld) {
    const long long total = (long long)g.N * g.H * g.W * g.Ci;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int ci = (int)(i % g.Ci);
        long long t = i / g.Ci;
        const int ww = (int)(t % g.W); t /= g.W;
        const int h = (int)(t % g.H);
        const int n = (int)(t / g.H);
        float acc = 0.f;
        for (int r = 0; r < g.kh; ++r) {
            const int hn = h + g.pad - r * g.dil;
            if (hn < 0 || hn % g.stride) continue;
            const int ho = hn / g.stride;
            if (ho >= g.Ho) continue;
            for (int s = 0; s < g.kw; ++s) {
                const int wn = ww + g.pad - s * g.dil;
                if (wn < 0 || wn % g.stride) continue;
                const int wo = wn / g.stride;
                if (wo >= g.Wo) continue;
                const __nv_bfloat16* dp = dy + ((long long)(n * g.Ho + ho) * g.Wo + wo) * g.y_ld;
                for (int co = 0; co < g.Co; ++co)
                    acc = fmaf(__bfloat162float(dp[co]),
                               __bfloat162float(w[(((long long)co * g.kh + r) * g.kw + s) * g.Ci_max + ci]), acc);
            }
        }
        const long long pix = (long long)(n * g.H + h) * g.W + ww;
        if (res) acc += __bfloat162float(res[pix * res_ld + ci]);
        dx[pix * g.x_ld + ci] = __float2bfloat16_rn(acc);
    }
}

__global__ void conv_wgrad_simt_kernel(gs_conv_geom g, const __nv_bfloat16* __restrict__ x,
                                       const __nv_bfloat16* __restrict__ dy, float* __restrict__ dw) {
    const long long total = (long long)g.Co * g.kh * g.kw * g.Ci;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int ci = (int)(i % g.Ci);
        long long t = i / g.Ci;
        const int s = (int)(t % g.kw); t /= g.kw;
        const int r = (int)(t % g.kh);
        const int co = (int)(t / g.kh);
        float acc = 0.f;
        for (int n = 0; n < g.N; ++n)
            for (int ho = 0; ho < g.Ho; ++ho) {
                const int h = ho * g.stride - g.pad + r * g.dil;
                if (h < 0 || h >= g.H) continue;
                for (int wo = 0; wo < g.Wo; ++wo) {
                    const int ww = wo * g.stride - g.pad + s * g.dil;
                    if (ww < 0 || ww >= g.W) continue;
                    acc = fmaf(__bfloat162float(dy[((long long)(n * g.Ho + ho) * g.Wo + wo) * g.y_ld + co]),
                               __bfloat162float(x[((long long)(n * g.H + h) * g.W + ww) * g.x_ld + ci]), acc);
                }
            }
        dw[(((long long)co * g.kh + r) * g.kw + s) * g.Ci_max + ci] += acc;
    }
}

static inline int simt_grid(long long total) {
    long long gsz = (total + 127) / 128;
    if (gsz > 148 * 32) gsz = 148 * 32;
    if (gsz < 1) gsz = 1;
    return (int)gsz;
}

}  // namespace gs

using namespace gs;

extern "C" int gs_conv2d_fwd_simt(const gs_conv_geom* g, const void* x, const void* w_krsc, void* y,
                                  const float* scale, const float* shift, const void* residual, int32_t res_ld,
                                  int32_t flags, double* stats, void* stream) {
    GS_REQUIRE(g && x && w_krsc && y, "conv_fwd_simt: null pointer");
    const long long total = (long long)g->N * g->Ho * g->Wo * g->Co;
    conv_fwd_simt_kernel<<<simt_grid(total), 128, 0, static_cast<cudaStream_t>(stream)>>>(
        *g, reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<const __nv_bfloat16*>(w_krsc), y, scale, shift,
        reinterpret_cast<const __nv_bfloat16*>(residual), res_ld, flags, stats);
    GS_LAUNCHED();
    return 0;
}

extern "C" int gs_conv2d_dgrad_simt(const gs_conv_geom* g, const void* dy, const void* w_krsc, void* dx,
                                    const void* residual, int32_t res_ld, void* stream) {
    GS_REQUIRE(g && dy && w_krsc && dx, "conv_dgrad_simt: null pointer");
    const long long total = (long long)g->N * g->H * g->W * g->Ci;
    conv_dgrad_simt_kernel<<<simt_grid(total), 128, 0, static_cast<cudaStream_t>(stream)>>>(
        *g, reinterpret_cast<const __nv_bfloat16*>(dy), reinterpret_cast<const __nv_bfloat16*>(w_krsc),
        reinterpret_cast<__nv_bfloat16*>(dx), reinterpret_cast<const __nv_bfloat16*>(residual), res_ld);
    GS_LAUNCHED();
    return 0;
}

extern "C" int gs_conv2d_wgrad_simt(const gs_conv_geom* g, const void* x, const void* dy, float* dw_krsc,
                                    void* stream) {
    GS_REQUIRE(g && x && dy && dw_krsc, "conv_wgrad_simt: null pointer");
    const long long total = (long long)g->Co * g->kh * g->kw * g->Ci;
    conv_wgrad_simt_kernel<<<simt_grid(total), 128, 0, static_cast<cudaStream_t>(stream)>>>(
        *g, reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<const __nv_bfloat16*>(dy), dw_krsc);
    GS_LAUNCHED();
    return 0;
}
