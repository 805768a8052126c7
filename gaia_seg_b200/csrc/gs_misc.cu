// gs_misc.cu -- pooling, layout and small streaming kernels (memory-bound, sm_100a).
#include "../../include/gaiaseg_b200.h"
#include "gs_host.h"
#include "gs_vec.cuh"

namespace gs {

static inline int flat_grid(long long total, int threads) {
    long long g = (total + threads - 1) / threads;
    const long long cap = 148LL * 16;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

// ------------------------------------------------------------------------------------------------
// MaxPool2d(3, 2, 1)   (gaiaseg/models/backbones/dynamic_resnet.py:302)
// First maximum in (r, s) scan order wins, as in ATen.  idx holds r*3+s per element for backward.
// ------------------------------------------------------------------------------------------------
__global__ void maxpool_fwd_kernel(const uint4* __restrict__ x, long long x_ld8, int N, int H, int W, int C8,
                                   uint4* __restrict__ y, long long y_ld8, uint2* __restrict__ idx, int Ho, int Wo) {
    pdl_sync();
    const long long total = (long long)N * Ho * Wo * C8;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int cv = (int)(i % C8);
        long long t = i / C8;
        const int wo = (int)(t % Wo); t /= Wo;
        const int ho = (int)(t % Ho);
        const int n = (int)(t / Ho);
        float best[8];
        uint32_t bi[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { best[k] = -INFINITY; bi[k] = 255; }
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const int h = ho * 2 - 1 + r;
            if (h < 0 || h >= H) continue;
#pragma unroll
            for (int s = 0; s < 3; ++s) {
                const int w = wo * 2 - 1 + s;
                if (w < 0 || w >= W) continue;
                float f[8];
                unpack8(__ldg(x + ((long long)(n * H + h) * W + w) * x_ld8 + cv), f);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    if (bi[k] == 255 || f[k] > best[k] || f[k] != f[k]) { best[k] = f[k]; bi[k] = r * 3 + s; }
                }
            }
        }
        const long long po = (long long)(n * Ho + ho) * Wo + wo;
        y[po * y_ld8 + cv] = pack8(best);
        if (idx) {
            uint2 o;
            o.x = bi[0] | (bi[1] << 8) | (bi[2] << 16) | (bi[3] << 24);
            o.y = bi[4] | (bi[5] << 8) | (bi[6] << 16) | (bi[7] << 24);
            idx[po * C8 + cv] = o;
        }
    }
}

__global__ void maxpool_bwd_kernel(const uint4* __restrict__ dy, long long dy_ld8, const uint2* __restrict__ idx, int N,
                                   int H, int W, int C8, int Ho, int Wo, uint4* __restrict__ dx, long long dx_ld8) {
    pdl_sync();
    const long long total = (long long)N * H * W * C8;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int cv = (int)(i % C8);
        long long t = i / C8;
        const int w = (int)(t % W); t /= W;
        const int h = (int)(t % H);
        const int n = (int)(t / H);
        float acc[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = 0.f;
        // windows ho with 2*ho-1 <= h <= 2*ho+1
        const int ho_lo = h >> 1;            // ceil((h-1)/2) for h >= 0
        const int ho_hi = (h + 1) >> 1;
        const int wo_lo = w >> 1;
        const int wo_hi = (w + 1) >> 1;
        for (int ho = ho_lo; ho <= ho_hi; ++ho) {
            if (ho >= Ho) continue;
            const int r = h - (2 * ho - 1);
            if (r < 0 || r > 2) continue;
            for (int wo = wo_lo; wo <= wo_hi; ++wo) {
                if (wo >= Wo) continue;
                const int s = w - (2 * wo - 1);
                if (s < 0 || s > 2) continue;
                const uint32_t key = r * 3 + s;
                const long long po = (long long)(n * Ho + ho) * Wo + wo;
                const uint2 id = __ldg(idx + po * C8 + cv);
                float g[8];
                unpack8(__ldg(dy + po * dy_ld8 + cv), g);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const uint32_t b = ((k < 4 ? id.x : id.y) >> ((k & 3) * 8)) & 0xFF;
                    if (b == key) acc[k] += g[k];
                }
            }
        }
        dx[((long long)(n * H + h) * W + w) * dx_ld8 + cv] = pack8(acc);
    }
}

// ------------------------------------------------------------------------------------------------
// AdaptiveAvgPool2d(s) for the PSP pyramid (gaiaseg/models/decode_heads/dynamic_psp_head.py:51)
// out[n, i, j, c] = mean over rows [floor(i*H/s), ceil((i+1)*H/s)) x cols likewise.  One block per
// (n, bin, channel-vector chunk); threads split the bin's pixels, smem tree reduce.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) adaptive_pool_fwd_kernel(const uint4* __restrict__ x, long long x_ld8, int H,
                                                                int W, int C8, int S, uint4* __restrict__ y,
                                                                long long y_ld8) {
    pdl_sync();
    __shared__ float red[256 * 8];
    const int bin = blockIdx.x % (S * S);
    const int n = blockIdx.x / (S * S);
    const int bi = bin / S, bj = bin % S;
    const int h0 = (bi * H) / S, h1 = ((bi + 1) * H + S - 1) / S;
    const int w0 = (bj * W) / S, w1 = ((bj + 1) * W + S - 1) / S;
    const int bw = w1 - w0, npx = (h1 - h0) * bw;
    const int Vc = C8 < 32 ? C8 : 32;
    const int R = 256 / Vc;
    const int cx = threadIdx.x % Vc, ry = threadIdx.x / Vc;
    for (int cv0 = blockIdx.y * Vc; cv0 < C8; cv0 += gridDim.y * Vc) {
        const int cv = cv0 + cx;
        float acc[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = 0.f;
        if (cv < C8 && ry < R) {
            for (int q = ry; q < npx; q += R) {
                const int h = h0 + q / bw, w = w0 + q % bw;
                float f[8];
                unpack8(__ldg(x + ((long long)(n * H + h) * W + w) * x_ld8 + cv), f);
#pragma unroll
                for (int k = 0; k < 8; ++k) acc[k] += f[k];
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 8; ++k) red[threadIdx.x * 8 + k] = acc[k];
        __syncthreads();
        if (ry == 0 && cv < C8) {
            float s[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) s[k] = 0.f;
            for (int r = 0; r < R; ++r)
#pragma unroll
                for (int k = 0; k < 8; ++k) s[k] += red[(r * Vc + cx) * 8 + k];
            const float inv = 1.f / (float)npx;
#pragma unroll
            for (int k = 0; k < 8; ++k) s[k] *= inv;
            y[((long long)n * S * S + bin) * y_ld8 + cv] = pack8(s);
        }
    }
}

// dx[n,h,w,c] (+)= sum over bins containing (h,w) of dy[n,bin,c] / npx(bin)
__global__ void adaptive_pool_bwd_kernel(const uint4* __restrict__ dy, long long dy_ld8, int N, int H, int W, int C8,
                                         int S, uint4* __restrict__ dx, long long dx_ld8, int accumulate) {
    pdl_sync();
    const long long total = (long long)N * H * W * C8;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int cv = (int)(i % C8);
        long long t = i / C8;
        const int w = (int)(t % W); t /= W;
        const int h = (int)(t % H);
        const int n = (int)(t / H);
        float acc[8];
        uint4* dst = dx + ((long long)(n * H + h) * W + w) * dx_ld8 + cv;
        if (accumulate) unpack8(*dst, acc);
        else {
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k] = 0.f;
        }
        for (int bi = 0; bi < S; ++bi) {
            const int h0 = (bi * H) / S, h1 = ((bi + 1) * H + S - 1) / S;
            if (h < h0 || h >= h1) continue;
            for (int bj = 0; bj < S; ++bj) {
                const int w0 = (bj * W) / S, w1 = ((bj + 1) * W + S - 1) / S;
                if (w < w0 || w >= w1) continue;
                const float inv = 1.f / (float)((h1 - h0) * (w1 - w0));
                float g[8];
                unpack8(__ldg(dy + ((long long)n * S * S + bi * S + bj) * dy_ld8 + cv), g);
#pragma unroll
                for (int k = 0; k < 8; ++k) acc[k] = fmaf(g[k], inv, acc[k]);
            }
        }
        *dst = pack8(acc);
    }
}

// ------------------------------------------------------------------------------------------------
// channel copies / adds (concat without torch.cat), dropout scale, casts, layout
// ------------------------------------------------------------------------------------------------
__global__ void copy_channels_kernel(const uint4* __restrict__ src, long long src_ld8, uint4* __restrict__ dst,
                                     long long dst_ld8, long long P, int C8) {
    pdl_sync();
    const long long total = P * C8;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long p = i / C8;
        const int cv = (int)(i - p * C8);
        stg_stream(dst + p * dst_ld8 + cv, ldg_stream(src + p * src_ld8 + cv));
    }
}

__global__ void add_channels_kernel(const uint4* __restrict__ src, long long src_ld8, uint4* __restrict__ dst,
                                    long long dst_ld8, long long P, int C8) {
    pdl_sync();
    const long long total = P * C8;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long p = i / C8;
        const int cv = (int)(i - p * C8);
        float a[8], b[8];
        unpack8(ldg_stream(src + p * src_ld8 + cv), a);
        uint4* d = dst + p * dst_ld8 + cv;
        unpack8(*d, b);
#pragma unroll
        for (int k = 0; k < 8; ++k) b[k] += a[k];
        *d = pack8(b);
    }
}

// y[n, p, c] = x[n, p, c] * scale_nc[n*C + c]      (Dropout2d mask / keep-prob)
__global__ void scale_nc_kernel(const uint4* __restrict__ x, long long x_ld8, const float* __restrict__ scale_nc,
                                uint4* __restrict__ y, long long y_ld8, int N, long long HW, int C8) {
    pdl_sync();
    const long long total = (long long)N * HW * C8;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int cv = (int)(i % C8);
        const long long p = i / C8;
        const int n = (int)(p / HW);
        float f[8], s[8];
        unpack8(ldg_stream(x + p * x_ld8 + cv), f);
        load8f(scale_nc + ((long long)n * C8 + cv) * 8, s);
#pragma unroll
        for (int k = 0; k < 8; ++k) f[k] *= s[k];
        stg_stream(y + p * y_ld8 + cv, pack8(f));
    }
}

// fp32 [P][src_ld] (C used) -> bf16 [P][dst_ld], columns C..dst_ld zero-filled
__global__ void cast_f32_bf16_kernel(const float* __restrict__ src, long long src_ld, __nv_bfloat16* __restrict__ dst,
                                     long long dst_ld, long long P, int C) {
    pdl_sync();
    const long long total = P * dst_ld;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long p = i / dst_ld;
        const int c = (int)(i - p * dst_ld);
        dst[i] = __float2bfloat16_rn(c < C ? src[p * src_ld + c] : 0.f);
    }
}

// out[c] (+)= sum_p src[p][c]    (fp32; bias gradient of conv_seg)
__global__ void __launch_bounds__(256) colsum_f32_kernel(const float* __restrict__ src, long long ld, long long P, int C,
                                                         float* __restrict__ out) {
    pdl_sync();
    // block handles a slab of pixels; thread t handles column t % Cc, row group t / Cc
    __shared__ float red[256];
    const int Cc = C < 256 ? C : 256;
    const int R = 256 / Cc;
    const int cx = threadIdx.x % Cc, ry = threadIdx.x / Cc;
    for (int c0 = 0; c0 < C; c0 += Cc) {
        const int c = c0 + cx;
        float s = 0.f;
        if (c < C && ry < R)
            for (long long p = (long long)blockIdx.x * R + ry; p < P; p += (long long)gridDim.x * R) s += src[p * ld + c];
        __syncthreads();
        red[threadIdx.x] = s;
        __syncthreads();
        if (ry == 0 && c < C) {
            float t = 0.f;
            for (int r = 0; r < R; ++r) t += red[r * Cc + cx];
            atomicAdd(out + c, t);
        }
    }
}

// fp32 NCHW -> bf16 NHWC (pitch ld, channels C..Cpad zero) via a 32x32 smem transpose per (n, c-tile, hw-tile)
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ src, int C, long long HW, __nv_bfloat16* __restrict__ dst,
                                    long long ld, int Cpad) {
    pdl_sync();
    __shared__ float tile[32][33];
    const int n = blockIdx.z;
    const long long hw0 = (long long)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int c = c0 + j;
        const long long hw = hw0 + threadIdx.x;
        tile[j][threadIdx.x] = (c < C && hw < HW) ? src[((long long)n * C + c) * HW + hw] : 0.f;
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const long long hw = hw0 + j;
        const int c = c0 + threadIdx.x;
        if (hw < HW && c < Cpad) dst[((long long)n * HW + hw) * ld + c] = __float2bfloat16_rn(tile[threadIdx.x][j]);
    }
}

__global__ void nhwc_to_nchw_kernel(const __nv_bfloat16* __restrict__ src, long long ld, int C, long long HW,
                                    float* __restrict__ dst) {
    pdl_sync();
    __shared__ float tile[32][33];
    const int n = blockIdx.z;
    const long long hw0 = (long long)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const long long hw = hw0 + j;
        const int c = c0 + threadIdx.x;
        tile[j][threadIdx.x] = (c < C && hw < HW) ? __bfloat162float(src[((long long)n * HW + hw) * ld + c]) : 0.f;
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int c = c0 + j;
        const long long hw = hw0 + threadIdx.x;
        if (c < C && hw < HW) dst[((long long)n * C + c) * HW + hw] = tile[threadIdx.x][j];
    }
}

// ------------------------------------------------------------------------------------------------
// im2col of the fp32 NCHW image for the first conv (Ci = 3): out[n, ho, wo, (r*kw + s)*C + c]
// ------------------------------------------------------------------------------------------------
__global__ void im2col_image_kernel(const float* __restrict__ img, int N, int C, int H, int W, int kh, int kw,
                                    int stride, int pad, int Ho, int Wo, int Kpad, __nv_bfloat16* __restrict__ out) {
    pdl_sync();
    const long long total = (long long)N * Ho * Wo * Kpad;
    const int K = kh * kw * C;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(i % Kpad);
        long long t = i / Kpad;
        const int wo = (int)(t % Wo); t /= Wo;
        const int ho = (int)(t % Ho);
        const int n = (int)(t / Ho);
        float v = 0.f;
        if (k < K) {
            const int c = k % C;
            const int rs = k / C;
            const int r = rs / kw, s = rs % kw;
            const int h = ho * stride - pad + r, w = wo * stride - pad + s;
            if (h >= 0 && h < H && w >= 0 && w < W) v = __ldg(img + ((long long)(n * C + c) * H + h) * W + w);
        }
        out[i] = __float2bfloat16_rn(v);
    }
}

// ------------------------------------------------------------------------------------------------
// bilinear resize of a bf16 NHWC map (align_corners = False) -- the PPM branches of the PSP head
// (resize at gaiaseg/models/decode_heads/dynamic_psp_head.py:67-71).  Source maps are tiny (s x s, s <= 6).
// forward: one thread per (output pixel, 8-channel vector).  backward: gather, one block per (n, source cell).
// ------------------------------------------------------------------------------------------------
struct Tap2 {
    int i0, i1;
    float l0, l1;
};
__device__ __forceinline__ Tap2 src_tap2(float scale, int dst, int in_size) {
    float src = scale * (static_cast<float>(dst) + 0.5f) - 0.5f;
    if (src < 0.f) src = 0.f;
    Tap2 t;
    t.i0 = static_cast<int>(src);
    if (t.i0 > in_size - 1) t.i0 = in_size - 1;
    t.i1 = t.i0 + ((t.i0 < in_size - 1) ? 1 : 0);
    t.l1 = src - static_cast<float>(t.i0);
    t.l1 = fminf(fmaxf(t.l1, 0.f), 1.f);
    t.l0 = 1.f - t.l1;
    return t;
}

__global__ void upsample_bf16_fwd_kernel(const uint4* __restrict__ src, long long src_ld8, int N, int h, int w, int C8,
                                         uint4* __restrict__ dst, long long dst_ld8, int H, int W, float rh, float rw) {
    pdl_sync();
    const long long total = (long long)N * H * W * C8;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int cv = (int)(i % C8);
        long long t = i / C8;
        const int x = (int)(t % W); t /= W;
        const int y = (int)(t % H);
        const int n = (int)(t / H);
        const Tap2 ty = src_tap2(rh, y, h), tx = src_tap2(rw, x, w);
        float a[8], b[8], c[8], d[8], o[8];
        unpack8(__ldg(src + ((long long)(n * h + ty.i0) * w + tx.i0) * src_ld8 + cv), a);
        unpack8(__ldg(src + ((long long)(n * h + ty.i0) * w + tx.i1) * src_ld8 + cv), b);
        unpack8(__ldg(src + ((long long)(n * h + ty.i1) * w + tx.i0) * src_ld8 + cv), c);
        unpack8(__ldg(src + ((long long)(n * h + ty.i1) * w + tx.i1) * src_ld8 + cv), d);
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = ty.l0 * (tx.l0 * a[k] + tx.l1 * b[k]) + ty.l1 * (tx.l0 * c[k] + tx.l1 * d[k]);
        stg_stream(dst + ((long long)(n * H + y) * W + x) * dst_ld8 + cv, pack8(o));
    }
}

__global__ void __launch_bounds__(256) upsample_bf16_bwd_kernel(const uint4* __restrict__ ddst, long long ddst_ld8, int H,
                                                                int W, int C8, uint4* __restrict__ dsrc,
                                                                long long dsrc_ld8, int h, int w, float rh, float rw) {
    pdl_sync();
    __shared__ float red[256 * 8];
    const int cell = blockIdx.x % (h * w);
    const int n = blockIdx.x / (h * w);
    const int ci = cell / w, cj = cell % w;
    const float inv_rh = 1.f / rh, inv_rw = 1.f / rw;
    int ylo = (int)floorf((ci - 0.5f) * inv_rh - 0.5f) - 1, yhi = (int)ceilf((ci + 1.5f) * inv_rh - 0.5f) + 1;
    int xlo = (int)floorf((cj - 0.5f) * inv_rw - 0.5f) - 1, xhi = (int)ceilf((cj + 1.5f) * inv_rw - 0.5f) + 1;
    ylo = ylo < 0 ? 0 : ylo; xlo = xlo < 0 ? 0 : xlo;
    yhi = yhi > H - 1 ? H - 1 : yhi; xhi = xhi > W - 1 ? W - 1 : xhi;
    const int bw = xhi - xlo + 1, npx = (yhi - ylo + 1) * bw;
    const int Vc = C8 < 32 ? C8 : 32;
    const int R = 256 / Vc;
    const int cx = threadIdx.x % Vc, ry = threadIdx.x / Vc;
    for (int cv0 = blockIdx.y * Vc; cv0 < C8; cv0 += gridDim.y * Vc) {
        const int cv = cv0 + cx;
        float acc[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = 0.f;
        if (cv < C8 && ry < R) {
            for (int q = ry; q < npx; q += R) {
                const int y = ylo + q / bw, x = xlo + q % bw;
                const Tap2 ty = src_tap2(rh, y, h), tx = src_tap2(rw, x, w);
                const float wy = (ty.i0 == ci ? ty.l0 : 0.f) + (ty.i1 == ci ? ty.l1 : 0.f);
                const float wx = (tx.i0 == cj ? tx.l0 : 0.f) + (tx.i1 == cj ? tx.l1 : 0.f);
                const float wgt = wy * wx;
                if (wgt == 0.f) continue;
                float g[8];
                unpack8(__ldg(ddst + ((long long)(n * H + y) * W + x) * ddst_ld8 + cv), g);
#pragma unroll
                for (int k = 0; k < 8; ++k) acc[k] = fmaf(wgt, g[k], acc[k]);
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 8; ++k) red[threadIdx.x * 8 + k] = acc[k];
        __syncthreads();
        if (ry == 0 && cv < C8) {
            float sum[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) sum[k] = 0.f;
            for (int r = 0; r < R; ++r)
#pragma unroll
                for (int k = 0; k < 8; ++k) sum[k] += red[(r * Vc + cx) * 8 + k];
            dsrc[((long long)n * h * w + cell) * dsrc_ld8 + cv] = pack8(sum);
        }
    }
}

// zero-fill channels [0, C) of P pixels (the gap of the segmented PSP concat)
__global__ void zero_channels_kernel(uint4* __restrict__ dst, long long dst_ld8, long long P, int C8) {
    pdl_sync();
    const long long total = P * C8;
    const uint4 z = make_uint4(0, 0, 0, 0);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long p = i / C8;
        stg_stream(dst + p * dst_ld8 + (int)(i - p * C8), z);
    }
}

static int check_act8(const void* p, long long ld, int C, const char* what) {
    GS_REQUIRE(p != nullptr, "%s: null pointer", what);
    GS_REQUIRE(C > 0 && C % 8 == 0, "%s: channels (%d) must be a positive multiple of 8", what, C);
    GS_REQUIRE(ld >= C && ld % 8 == 0, "%s: pitch (%lld) must be >= C and a multiple of 8", what, ld);
    GS_REQUIRE((reinterpret_cast<uintptr_t>(p) & 15) == 0, "%s: pointer must be 16-byte aligned", what);
    return 0;
}

}  // namespace gs

using namespace gs;

extern "C" int gs_maxpool3x3s2_fwd(const void* x, int32_t N, int32_t H, int32_t W, int32_t C, int32_t x_ld, void* y,
                                   int32_t Ho, int32_t Wo, int32_t y_ld, void* idx, void* stream) {
    if (check_act8(x, x_ld, C, "maxpool x") || check_act8(y, y_ld, C, "maxpool y")) return -1;
    GS_REQUIRE(Ho == (H + 2 - 3) / 2 + 1 && Wo == (W + 2 - 3) / 2 + 1, "maxpool: output size (%d,%d) inconsistent", Ho, Wo);
    const long long total = (long long)N * Ho * Wo * (C / 8);
    if (total <= 0) return 0;
    gs::launch(maxpool_fwd_kernel, dim3(flat_grid(total, 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), 
        reinterpret_cast<const uint4*>(x), x_ld / 8, N, H, W, C / 8, reinterpret_cast<uint4*>(y), y_ld / 8,
        reinterpret_cast<uint2*>(idx), Ho, Wo);
    GS_LAUNCHED();
    return 0;
}

extern "C" int gs_maxpool3x3s2_bwd(const void* dy, int32_t dy_ld, const void* idx, int32_t N, int32_t H, int32_t W,
                                   int32_t C, int32_t Ho, int32_t Wo, void* dx, int32_t dx_ld, void* stream) {
    if (check_act8(dy, dy_ld, C, "maxpool dy") || check_act8(dx, dx_ld, C, "maxpool dx")) return -1;
    GS_REQUIRE(idx != nullptr, "maxpool_bwd: null index tensor");
    const long long total = (long long)N * H * W * (C / 8);
    if (total <= 0) return 0;
    gs::launch(maxpool_bwd_kernel, dim3(flat_grid(total, 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), 
        reinterpret_cast<const uint4*>(dy), dy_ld / 8, reinterpret_cast<const uint2*>(idx), N, H, W, C / 8, Ho, Wo,
        reinterpret_cast<uint4*>(dx), dx_ld / 8);
    GS_LAUNCHED();
    return 0;
}

extern "C" int gs_adaptive_avgpool_fwd(const void* x, int32_t N, int32_t H, int32_t W, int32_t C, int32_t x_ld,
                                       int32_t S, void* y, int32_t y_ld, void* stream) {
    if (check_act8(x, x_ld, C, "adaptive_pool x") || check_act8(y, y_ld, C, "adaptive_pool y")) return -1;
    GS_REQUIRE(S >= 1 && S <= H && S <= W, "adaptive_pool: bins %d vs input %dx%d", S, H, W);
    const int C8 = C / 8;
    const int Vc = C8 < 32 ? C8 : 32;
    int gy = (C8 + Vc - 1) / Vc;
    dim3 grid(N * S * S, gy);
    gs::launch(adaptive_pool_fwd_kernel, dim3(grid), dim3(256), 0, static_cast<cudaStream_t>(stream), 
        reinterpret_cast<const uint4*>(x), x_ld / 8, H, W, C8, S, reinterpret_cast<uint4*>(y), y_ld / 8);
    GS_LAUNCHED();
    return 0;
}

extern "C" int gs_adaptive_avgpool_bwd(const void* dy, int32_t dy_ld, int32_t N, int32_t H, int32_t W, int32_t C,
                                       int32_t S, void* dx, int32_t dx_ld, int32_t accumulate, void* stream) {
    if (check_act8(dy, dy_ld, C, "adaptive_pool dy") || check_act8(dx, dx_ld, C, "adaptive_pool dx")) return -1;
    const long long total = (long long)N * H * W * (C / 8);
    if (total <= 0) return 0;
    gs::launch(adaptive_pool_bwd_kernel, dim3(flat_grid(total, 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), 
        reinterpret_cast<const uint4*>(dy), dy_ld / 8, N, H, W, C / 8, S, reinterpret_cast<uint4*>(dx), dx_ld / 8,
        accumulate);
    GS_LAUNCHED();
    return 0;
}

extern "C" int gs_upsample_bf16_fwd(const void* src, int32_t src_ld, int32_t N, int32_t h, int32_t w, int32_t C, void* dst,
                                    int32_t dst_ld, int32_t H, int32_t W, void* stream) {
    if (check_act8(src, src_ld, C, "upsample_bf16 src") || check_act8(dst, dst_ld, C, "upsample_bf16 dst")) return -1;
    const long long total = (long long)N * H * W * (C / 8);
    if (total <= 0) return 0;
    const float rh = static_cast<float>(h) / static_cast<float>(H), rw = static_cast<float>(w) / static_cast<float>(W);
    gs::launch(upsample_bf16_fwd_kernel, dim3(flat_grid(total, 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), 
        reinterpret_cast<const uint4*>(src), src_ld / 8, N, h, w, C / 8, reinterpret_cast<uint4*>(dst), dst_ld / 8, H, W,
        rh, rw);
    GS_LAUNCHED();
    return 0;
}

extern "C" int gs_upsample_bf16_bwd(const void* ddst, int32_t ddst_ld, int32_t N, int32_t H, int32_t W, int32_t C,
                                    void* dsrc, int32_t dsrc_ld, int32_t h, int32_t w, void* stream) {
    if (check_act8(ddst, ddst_ld, C, "upsample_bf16_bwd ddst") || check_act8(dsrc, dsrc_ld, C, "upsample_bf16_bwd dsrc"))
        return -1;
    GS_REQUIRE(N > 0 && h > 0 && w > 0 && H > 0 && W > 0, "upsample_bf16_bwd: bad shape");
    const float rh = static_cast<float>(h) / static_cast<float>(H), rw = static_cast<float>(w) / static_cast<float>(W);
    const int C8 = C / 8;
    const int Vc = C8 < 32 ? C8 : 32;
    dim3 grid(N * h * w, (C8 + Vc - 1) / Vc);
    gs::launch(upsample_bf16_bwd_kernel, dim3(grid), dim3(256), 0, static_cast<cudaStream_t>(stream), 
        reinterpret_cast<const uint4*>(ddst), ddst_ld / 8, H, W, C8, reinterpret_cast<uint4*>(dsrc), dsrc_ld / 8, h, w, rh,
        rw);
    GS_LAUNCHED();
    return 0;
}

extern "C" int gs_zero_channels(void* dst, int32_t dst_ld, int64_t P, int32_t C, void* stream) {
    if (check_act8(dst, dst_ld, C, "zero_channels dst")) return -1;
    if (P <= 0) return 0;
    gs::launch(zero_channels_kernel, dim3(flat_grid(P * (C / 8), 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), 
        reinterpret_cast<uint4*>(dst), dst_ld / 8, P, C / 8);
    GS_LAUNCHED();
    return 0;
}

extern "C" int gs_copy_channels(const void* src, int32_t src_ld, void* dst, int32_t dst_ld, int64_t P, int32_t C,
                                void* stream) {
    if (check_act8(src, src_ld, C, "copy_channels src") || check_act8(dst, dst_ld, C, "copy_channels dst")) return -1;
    if (P <= 0) return 0;
    gs::launch(copy_channels_kernel, dim3(flat_grid(P * (C / 8), 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), 
        reinterpret_cast<const uint4*>(src), src_ld / 8, reinterpret_cast<uint4*>(dst), dst_ld / 8, P, C / 8);
    GS_LAUNCHED();
    return 0;
}

extern "C" int gs_add_channels(const void* src, int32_t src_ld, void* dst, int32_t dst_ld, int64_t P, int32_t C,
                               void* stream) {
    if (check_act8(src, src_ld, C, "add_channels src") || check_act8(dst, dst_ld, C, "add_channels dst")) return -1;
    if (P <= 0) return 0;
    gs::launch(add_channels_kernel, dim3(flat_grid(P * (C / 8), 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), 
        reinterpret_cast<const uint4*>(src), src_ld / 8, reinterpret_cast<uint4*>(dst), dst_ld / 8, P, C / 8);
    GS_LAUNCHED();
    return 0;
}

extern "C" int gs_scale_nc(const void* x, int32_t x_ld, const float* scale_nc, void* y, int32_t y_ld, int32_t N,
                           int64_t HW, int32_t C, void* stream) {
    if (check_act8(x, x_ld, C, "scale_nc x") || check_act8(y, y_ld, C, "scale_nc y")) return -1;
    GS_REQUIRE(scale_nc != nullptr, "scale_nc: null scale");
    const long long total = (long long)N * HW * (C / 8);
    if (total <= 0) return 0;
    gs::launch(scale_nc_kernel, dim3(flat_grid(total, 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), 
        reinterpret_cast<const uint4*>(x), x_ld / 8, scale_nc, reinterpret_cast<uint4*>(y), y_ld / 8, N, HW, C / 8);
    GS_LAUNCHED();
    return 0;
}

extern "C" int gs_cast_f32_bf16(const float* src, int32_t src_ld, void* dst, int32_t dst_ld, int64_t P, int32_t C,
                                void* stream) {
    GS_REQUIRE(src && dst && C > 0 && src_ld >= C && dst_ld >= C, "cast_f32_bf16: bad arguments");
    if (P <= 0) return 0;
    gs::launch(cast_f32_bf16_kernel, dim3(flat_grid(P * dst_ld, 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), 
        src, src_ld, reinterpret_cast<__nv_bfloat16*>(dst), dst_ld, P, C);
    GS_LAUNCHED();
    return 0;
}

extern "C" int gs_colsum_f32(const float* src, int32_t ld, int64_t P, int32_t C, float* out, void* stream) {
    GS_REQUIRE(src && out && C > 0 && ld >= C, "colsum_f32: bad arguments");
    if (P <= 0) return 0;
    const int Cc = C < 256 ? C : 256;
    const int R = 256 / Cc;
    long long g = (P + (long long)R * 64 - 1) / ((long long)R * 64);
    if (g > 148 * 4) g = 148 * 4;
    if (g < 1) g = 1;
    gs::launch(colsum_f32_kernel, dim3((int)g), dim3(256), 0, static_cast<cudaStream_t>(stream), src, ld, P, C, out);
    GS_LAUNCHED();
    return 0;
}

extern "C" int gs_nchw_f32_to_nhwc_bf16(const float* src, int32_t N, int32_t C, int32_t H, int32_t W, void* dst,
                                        int32_t ld, int32_t Cpad, void* stream) {
    GS_REQUIRE(src && dst && N > 0 && C > 0 && Cpad >= C && ld >= Cpad, "nchw->nhwc: bad arguments");
    const long long HW = (long long)H * W;
    dim3 grid((unsigned)((HW + 31) / 32), (unsigned)((Cpad + 31) / 32), (unsigned)N);
    gs::launch(nchw_to_nhwc_kernel, dim3(grid), dim3(dim3(32, 8)), 0, static_cast<cudaStream_t>(stream), 
        src, C, HW, reinterpret_cast<__nv_bfloat16*>(dst), ld, Cpad);
    GS_LAUNCHED();
    return 0;
}

extern "C" int gs_nhwc_bf16_to_nchw_f32(const void* src, int32_t ld, int32_t N, int32_t C, int32_t H, int32_t W,
                                        float* dst, void* stream) {
    GS_REQUIRE(src && dst && N > 0 && C > 0 && ld >= C, "nhwc->nchw: bad arguments");
    const long long HW = (long long)H * W;
    dim3 grid((unsigned)((HW + 31) / 32), (unsigned)((C + 31) / 32), (unsigned)N);
    gs::launch(nhwc_to_nchw_kernel, dim3(grid), dim3(dim3(32, 8)), 0, static_cast<cudaStream_t>(stream), 
        reinterpret_cast<const __nv_bfloat16*>(src), ld, C, HW, dst);
    GS_LAUNCHED();
    return 0;
}

extern "C" int gs_im2col_image(const float* img_nchw, int32_t N, int32_t C, int32_t H, int32_t W, int32_t kh,
                               int32_t kw, int32_t stride, int32_t pad, int32_t Ho, int32_t Wo, int32_t Kpad, void* out,
                               void* stream) {
    GS_REQUIRE(img_nchw && out, "im2col: null pointer");
    GS_REQUIRE(Kpad >= kh * kw * C && Kpad % 8 == 0, "im2col: Kpad %d too small / unaligned for K=%d", Kpad, kh * kw * C);
    GS_REQUIRE(Ho == (H + 2 * pad - kh) / stride + 1 && Wo == (W + 2 * pad - kw) / stride + 1,
               "im2col: output size (%d,%d) inconsistent", Ho, Wo);
    const long long total = (long long)N * Ho * Wo * Kpad;
    if (total <= 0) return 0;
    gs::launch(im2col_image_kernel, dim3(flat_grid(total, 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), 
        img_nchw, N, C, H, W, kh, kw, stride, pad, Ho, Wo, Kpad, reinterpret_cast<__nv_bfloat16*>(out));
    GS_LAUNCHED();
    return 0;
}
