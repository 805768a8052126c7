// gs_optim.cu -- fused SGD(momentum, weight decay) over the flat fp32 master buffer + refresh of the
// bf16 shadow weights the tcgen05 convolutions read through TMA.
// replaces torch.optim.SGD.step (configs/_dynamic_/models/pspnet_ar50to101v2_gsync.py:175-178).
//
// Master conv weights are stored [Co_max][kh][kw][Ci_max] (channels_last memory of the OIHW
// parameter), so the forward shadow `w_krsc` is an element-wise bf16 cast at the same flat index;
// dgrad reads the same buffer as an MN-major tcgen05 operand, so no transposed copy exists.
#include "../../include/gaiaseg_b200.h"
#include "gs_host.h"

#include <cuda_bf16.h>

namespace gs {

__global__ void __launch_bounds__(256) sgd_flat_kernel(float4* __restrict__ p, const float4* __restrict__ g,
                                                       float4* __restrict__ buf, long long n4,
                                                       const float* __restrict__ hyper, int first,
                                                       uint2* __restrict__ shadow) {
    pdl_sync();
    // hyper-parameters live in device memory so that a captured CUDA graph of the step follows the LR schedule
    const float lr = __ldg(hyper), mom = __ldg(hyper + 1), wd = __ldg(hyper + 2), gscale = __ldg(hyper + 3);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4;
         i += (long long)gridDim.x * blockDim.x) {
        float4 pv = p[i];
        const float4 gv = g[i];
        float4 bv;
        const float d0 = fmaf(wd, pv.x, gv.x * gscale), d1 = fmaf(wd, pv.y, gv.y * gscale);
        const float d2 = fmaf(wd, pv.z, gv.z * gscale), d3 = fmaf(wd, pv.w, gv.w * gscale);
        if (first) {
            bv = make_float4(d0, d1, d2, d3);
        } else {
            bv = buf[i];
            bv.x = fmaf(mom, bv.x, d0); bv.y = fmaf(mom, bv.y, d1);
            bv.z = fmaf(mom, bv.z, d2); bv.w = fmaf(mom, bv.w, d3);
        }
        buf[i] = bv;
        pv.x -= lr * bv.x; pv.y -= lr * bv.y; pv.z -= lr * bv.z; pv.w -= lr * bv.w;
        p[i] = pv;
        if (shadow) {
            __nv_bfloat162 a = __floats2bfloat162_rn(pv.x, pv.y), b = __floats2bfloat162_rn(pv.z, pv.w);
            uint2 o;
            o.x = *reinterpret_cast<uint32_t*>(&a);
            o.y = *reinterpret_cast<uint32_t*>(&b);
            shadow[i] = o;
        }
    }
}

}  // namespace gs

using namespace gs;

extern "C" int gs_sgd_flat(float* p, const float* g, float* momentum_buf, int64_t n, const float* hyper,
                           int32_t first_step, void* shadow_bf16, void* stream) {
    GS_REQUIRE(p && g && momentum_buf && hyper, "sgd_flat: null pointer");
    GS_REQUIRE(n % 4 == 0, "sgd_flat: n (%lld) must be a multiple of 4 (flat buffers are padded)", (long long)n);
    GS_REQUIRE(((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) |
                 reinterpret_cast<uintptr_t>(momentum_buf)) & 15) == 0, "sgd_flat: buffers must be 16-byte aligned");
    if (n == 0) return 0;
    const long long n4 = n / 4;
    long long grid = (n4 + 255) / 256;
    if (grid > 148 * 16) grid = 148 * 16;
    gs::launch(sgd_flat_kernel, dim3((int)grid), dim3(256), 0, static_cast<cudaStream_t>(stream), 
        reinterpret_cast<float4*>(p), reinterpret_cast<const float4*>(g), reinterpret_cast<float4*>(momentum_buf), n4,
        hyper, first_step, reinterpret_cast<uint2*>(shadow_bf16));
    GS_LAUNCHED();
    return 0;
}

