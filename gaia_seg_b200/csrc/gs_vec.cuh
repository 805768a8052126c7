// gs_vec.cuh -- shared helpers for the memory-bound NHWC bf16 kernels.
//
// Layout contract: an activation is [P pixels][ld elements] bf16 with C <= ld active channels,
// C % 8 == 0, ld % 8 == 0, base 16-byte aligned => every (pixel, 8-channel group) is one 16-byte
// vector.  Thread mapping ("column map"): a block is Vc x R threads; thread (cx, ry) owns the
// channel vectors cx, cx+Vc, ... and walks pixels ry, ry+R*grid, ... so that per-channel
// parameters (scale / shift / mean ...) live in registers across the whole pixel loop and a warp
// reads consecutive 16-byte vectors of one pixel row (fully coalesced 128-bit accesses).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace gs {

struct ColMap {
    int C8;       // channel vectors per pixel
    int Vc;       // thread columns
    int R;        // pixel rows per block
    int threads;  // Vc * R
};

static inline ColMap make_colmap(int C) {
    ColMap m;
    m.C8 = C / 8;
    const int nchunk = (m.C8 + 255) / 256;
    m.Vc = (m.C8 + nchunk - 1) / nchunk;
    m.R = 256 / m.Vc;
    if (m.R < 1) m.R = 1;
    m.threads = m.Vc * m.R;
    return m;
}

// grid for a streaming kernel: enough blocks for ~`ppt` pixels per thread row, capped to a few waves
static inline int colmap_grid(const ColMap& m, long long P, int ppt, int max_blocks) {
    long long g = (P + (long long)m.R * ppt - 1) / ((long long)m.R * ppt);
    if (g < 1) g = 1;
    if (g > max_blocks) g = max_blocks;
    return (int)g;
}

__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_stream(uint4* p, const uint4& v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
                 "r"(v.w)
                 : "memory");
}

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
    f[0] = __uint_as_float(u.x << 16); f[1] = __uint_as_float(u.x & 0xFFFF0000u);
    f[2] = __uint_as_float(u.y << 16); f[3] = __uint_as_float(u.y & 0xFFFF0000u);
    f[4] = __uint_as_float(u.z << 16); f[5] = __uint_as_float(u.z & 0xFFFF0000u);
    f[6] = __uint_as_float(u.w << 16); f[7] = __uint_as_float(u.w & 0xFFFF0000u);
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
    uint4 u;
    u.x = pack2(f[0], f[1]); u.y = pack2(f[2], f[3]); u.z = pack2(f[4], f[5]); u.w = pack2(f[6], f[7]);
    return u;
}
__device__ __forceinline__ void load8f(const float* p, float (&f)[8]) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p));
    const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

}  // namespace gs
