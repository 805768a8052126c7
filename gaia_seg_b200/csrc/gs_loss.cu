// gs_loss.cu -- fused bilinear-upsample + cross-entropy(ignore_index) + top-1 accuracy, its gradient,
// and fused upsample + argmax for inference.  The N*K*H*W up-sampled logit tensor of the reference
// (resize -> F.cross_entropy -> accuracy; gaiaseg/models/decode_heads/dynamic_fcn_head.py:137-159,
// CE restated at gaiaseg/models/losses/cross_entropy_loss.py:67-94, accuracy at accuracy.py:4-49)
// is never materialised:
//   forward : one thread per OUTPUT pixel, 4-tap lerp of the K low-res logits on the fly, online
//             soft-max (max / sum-exp / first arg-max in one sweep); writes an 8-byte record
//             (log-sum-exp, label or -1) per pixel for the backward pass.
//   backward: GATHER form, one thread per (low-res cell, class): walks the output pixels whose taps
//             touch the cell, re-interpolates its class logit, p = exp(v - lse) and accumulates
//             w * (p - [label == k]).  No atomics, deterministic.
// Source-index arithmetic is ATen's area_pixel_compute_source_index (align_corners = False), fp32.
//
// Algorithmic bytes per call: logits N*h*w*K*4 (read) + labels N*H*W*8 (read) + records N*H*W*8
// (write, read again by backward) + dlogits N*h*w*K*4 (write).
#include <math.h>

#include "../../include/gaiaseg_b200.h"
#include "gs_host.h"

namespace gs {

struct Tap {
    int i0, i1;
    float l0, l1;
};

__device__ __forceinline__ Tap src_tap(float scale, int dst, int in_size) {
    float src = scale * (static_cast<float>(dst) + 0.5f) - 0.5f;
    if (src < 0.f) src = 0.f;
    Tap t;
    t.i0 = static_cast<int>(src);
    if (t.i0 > in_size - 1) t.i0 = in_size - 1;
    t.i1 = t.i0 + ((t.i0 < in_size - 1) ? 1 : 0);
    t.l1 = src - static_cast<float>(t.i0);
    if (t.l1 < 0.f) t.l1 = 0.f;
    if (t.l1 > 1.f) t.l1 = 1.f;
    t.l0 = 1.f - t.l1;
    return t;
}

__device__ __forceinline__ float lerp4(const Tap& ty, const Tap& tx, float a, float b, float c, float d) {
    return ty.l0 * (tx.l0 * a + tx.l1 * b) + ty.l1 * (tx.l0 * c + tx.l1 * d);
}

struct PixRec {
    float lse;
    int label;  // -1: ignored
};

__global__ void __launch_bounds__(256) upsample_ce_fwd_kernel(const float* __restrict__ logits, int N, int h, int w,
                                                              int K, int ld, const long long* __restrict__ labels,
                                                              int H, int W, int ignore_index, float rh, float rw,
                                                              PixRec* __restrict__ rec, double* __restrict__ out_sum,
                                                              unsigned long long* __restrict__ out_counts) {
    double loss_acc = 0.0;
    unsigned int n_ign = 0, n_hit = 0;
    const long long total = (long long)N * H * W;
    for (long long pix = blockIdx.x * (long long)blockDim.x + threadIdx.x; pix < total;
         pix += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(pix % W);
        const long long t = pix / W;
        const int y = (int)(t % H);
        const int n = (int)(t / H);
        const Tap ty = src_tap(rh, y, h), tx = src_tap(rw, x, w);
        const float* p00 = logits + ((long long)(n * h + ty.i0) * w + tx.i0) * ld;
        const float* p01 = logits + ((long long)(n * h + ty.i0) * w + tx.i1) * ld;
        const float* p10 = logits + ((long long)(n * h + ty.i1) * w + tx.i0) * ld;
        const float* p11 = logits + ((long long)(n * h + ty.i1) * w + tx.i1) * ld;
        const long long lab = labels[pix];
        const bool valid = (lab != ignore_index) && lab >= 0 && lab < K;
        float m = -INFINITY, s = 0.f, vlab = 0.f, best = -INFINITY;
        int arg = 0;
        for (int k = 0; k < K; ++k) {
            const float v = lerp4(ty, tx, __ldg(p00 + k), __ldg(p01 + k), __ldg(p10 + k), __ldg(p11 + k));
            if (v > best || k == 0) { best = v; arg = k; }
            if (v > m) {
                s = s * expf(m - v) + 1.f;
                m = v;
            } else {
                s += expf(v - m);
            }
            if (k == (int)lab) vlab = v;
        }
        PixRec r;
        if (valid) {
            const float lse = m + logf(s);
            loss_acc += static_cast<double>(lse - vlab);
            r.lse = lse;
            r.label = (int)lab;
            if (arg == (int)lab) ++n_hit;
        } else {
            r.lse = 0.f;
            r.label = -1;
            ++n_ign;
        }
        if (rec) rec[pix] = r;
    }
    // block reduction -> one atomic triple per block
    __shared__ double s_loss[8];
    __shared__ unsigned int s_ign[8], s_hit[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        loss_acc += __shfl_xor_sync(0xffffffffu, loss_acc, o);
        n_ign += __shfl_xor_sync(0xffffffffu, n_ign, o);
        n_hit += __shfl_xor_sync(0xffffffffu, n_hit, o);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { s_loss[warp] = loss_acc; s_ign[warp] = n_ign; s_hit[warp] = n_hit; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0;
        unsigned long long b = 0, c = 0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { a += s_loss[i]; b += s_ign[i]; c += s_hit[i]; }
        atomicAdd(out_sum, a);
        atomicAdd(out_counts, b);
        atomicAdd(out_counts + 1, c);
    }
}

__global__ void __launch_bounds__(256) upsample_ce_bwd_kernel(const float* __restrict__ logits, int N, int h, int w,
                                                              int K, int ld, const PixRec* __restrict__ rec, int H,
                                                              int W, float rh, float rw, float grad_scale,
                                                              const float* __restrict__ grad_scale_dev,
                                                              float* __restrict__ dlogits, int dl_ld) {
    const long long total = (long long)N * h * w * K;
    if (grad_scale_dev) grad_scale *= __ldg(grad_scale_dev);
    const float inv_rh = 1.f / rh, inv_rw = 1.f / rw;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(idx % K);
        long long t = idx / K;
        const int j = (int)(t % w); t /= w;
        const int i = (int)(t % h);
        const int n = (int)(t / h);
        int ylo = (int)floorf((i - 0.5f) * inv_rh - 0.5f) - 1;
        int yhi = (int)ceilf((i + 1.5f) * inv_rh - 0.5f) + 1;
        int xlo = (int)floorf((j - 0.5f) * inv_rw - 0.5f) - 1;
        int xhi = (int)ceilf((j + 1.5f) * inv_rw - 0.5f) + 1;
        if (ylo < 0) ylo = 0;
        if (xlo < 0) xlo = 0;
        if (yhi > H - 1) yhi = H - 1;
        if (xhi > W - 1) xhi = W - 1;
        const float* base = logits + (long long)n * h * w * ld + k;
        float acc = 0.f;
        for (int y = ylo; y <= yhi; ++y) {
            const Tap ty = src_tap(rh, y, h);
            const float wy = (ty.i0 == i ? ty.l0 : 0.f) + (ty.i1 == i ? ty.l1 : 0.f);
            if (wy == 0.f) continue;
            const float* r0 = base + (long long)ty.i0 * w * ld;
            const float* r1 = base + (long long)ty.i1 * w * ld;
            const PixRec* rrow = rec + ((long long)n * H + y) * W;
            for (int x = xlo; x <= xhi; ++x) {
                const Tap tx = src_tap(rw, x, w);
                const float wx = (tx.i0 == j ? tx.l0 : 0.f) + (tx.i1 == j ? tx.l1 : 0.f);
                if (wx == 0.f) continue;
                const PixRec r = rrow[x];
                if (r.label < 0) continue;
                const float v = lerp4(ty, tx, __ldg(r0 + (long long)tx.i0 * ld), __ldg(r0 + (long long)tx.i1 * ld),
                                      __ldg(r1 + (long long)tx.i0 * ld), __ldg(r1 + (long long)tx.i1 * ld));
                const float g = __expf(v - r.lse) - (r.label == k ? 1.f : 0.f);
                acc = fmaf(wy * wx, g, acc);
            }
        }
        dlogits[((long long)(n * h + i) * w + j) * dl_ld + k] = acc * grad_scale;
    }
}

// fused upsample + argmax (first maximum wins); optional two-step resize is done by the caller with
// gs_upsample_bilinear_f32 first.
__global__ void __launch_bounds__(256) upsample_argmax_kernel(const float* __restrict__ logits, int N, int h, int w,
                                                              int K, int ld, int H, int W, float rh, float rw,
                                                              long long* __restrict__ out) {
    const long long total = (long long)N * H * W;
    for (long long pix = blockIdx.x * (long long)blockDim.x + threadIdx.x; pix < total;
         pix += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(pix % W);
        const long long t = pix / W;
        const int y = (int)(t % H);
        const int n = (int)(t / H);
        const Tap ty = src_tap(rh, y, h), tx = src_tap(rw, x, w);
        const float* p00 = logits + ((long long)(n * h + ty.i0) * w + tx.i0) * ld;
        const float* p01 = logits + ((long long)(n * h + ty.i0) * w + tx.i1) * ld;
        const float* p10 = logits + ((long long)(n * h + ty.i1) * w + tx.i0) * ld;
        const float* p11 = logits + ((long long)(n * h + ty.i1) * w + tx.i1) * ld;
        float best = -INFINITY;
        int arg = 0;
        for (int k = 0; k < K; ++k) {
            const float v = lerp4(ty, tx, __ldg(p00 + k), __ldg(p01 + k), __ldg(p10 + k), __ldg(p11 + k));
            if (v > best || k == 0) { best = v; arg = k; }
        }
        out[pix] = arg;
    }
}

__global__ void __launch_bounds__(256) upsample_bilinear_kernel(const float* __restrict__ src, int N, int h, int w, int K,
                                                                int ld, float* __restrict__ dst, int H, int W,
                                                                int dst_ld, float rh, float rw) {
    const long long total = (long long)N * H * W * K;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(idx % K);
        long long t = idx / K;
        const int x = (int)(t % W); t /= W;
        const int y = (int)(t % H);
        const int n = (int)(t / H);
        const Tap ty = src_tap(rh, y, h), tx = src_tap(rw, x, w);
        const float* b = src + (long long)n * h * w * ld + k;
        const float v = lerp4(ty, tx, __ldg(b + ((long long)ty.i0 * w + tx.i0) * ld),
                              __ldg(b + ((long long)ty.i0 * w + tx.i1) * ld),
                              __ldg(b + ((long long)ty.i1 * w + tx.i0) * ld),
                              __ldg(b + ((long long)ty.i1 * w + tx.i1) * ld));
        dst[((long long)(n * H + y) * W + x) * dst_ld + k] = v;
    }
}

static inline int loss_grid(long long total) {
    long long g = (total + 255) / 256;
    if (g > 148 * 8) g = 148 * 8;
    if (g < 1) g = 1;
    return (int)g;
}

}  // namespace gs

using namespace gs;

extern "C" int64_t gs_upsample_ce_record_bytes(int32_t N, int32_t H, int32_t W) {
    return (int64_t)N * H * W * (int64_t)sizeof(PixRec);
}

extern "C" int gs_upsample_ce_fwd(const float* logits, int32_t N, int32_t h, int32_t w, int32_t K, int32_t ld,
                                  const int64_t* labels, int32_t H, int32_t W, int32_t ignore_index, double* out_sum,
                                  int64_t* out_counts, void* pix_rec, void* stream) {
    GS_REQUIRE(logits && labels && out_sum && out_counts, "upsample_ce_fwd: null pointer");
    GS_REQUIRE(N > 0 && h > 0 && w > 0 && H > 0 && W > 0 && K > 0 && ld >= K, "upsample_ce_fwd: bad shape");
    const float rh = static_cast<float>(h) / static_cast<float>(H), rw = static_cast<float>(w) / static_cast<float>(W);
    upsample_ce_fwd_kernel<<<loss_grid((long long)N * H * W), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        logits, N, h, w, K, ld, reinterpret_cast<const long long*>(labels), H, W, ignore_index, rh, rw,
        reinterpret_cast<PixRec*>(pix_rec), out_sum, reinterpret_cast<unsigned long long*>(out_counts));
    GS_LAUNCHED();
    return 0;
}

extern "C" int gs_upsample_ce_bwd(const float* logits, int32_t N, int32_t h, int32_t w, int32_t K, int32_t ld,
                                  const void* pix_rec, int32_t H, int32_t W, float grad_scale,
                                  const float* grad_scale_dev, float* dlogits, int32_t dl_ld, void* stream) {
    GS_REQUIRE(logits && pix_rec && dlogits, "upsample_ce_bwd: null pointer");
    GS_REQUIRE(N > 0 && h > 0 && w > 0 && H > 0 && W > 0 && K > 0 && ld >= K && dl_ld >= K, "upsample_ce_bwd: bad shape");
    const float rh = static_cast<float>(h) / static_cast<float>(H), rw = static_cast<float>(w) / static_cast<float>(W);
    upsample_ce_bwd_kernel<<<loss_grid((long long)N * h * w * K), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        logits, N, h, w, K, ld, reinterpret_cast<const PixRec*>(pix_rec), H, W, rh, rw, grad_scale, grad_scale_dev, dlogits,
        dl_ld);
    GS_LAUNCHED();
    return 0;
}

extern "C" int gs_upsample_argmax(const float* logits, int32_t N, int32_t h, int32_t w, int32_t K, int32_t ld,
                                  int32_t H, int32_t W, int64_t* labels_out, void* stream) {
    GS_REQUIRE(logits && labels_out, "upsample_argmax: null pointer");
    GS_REQUIRE(N > 0 && h > 0 && w > 0 && H > 0 && W > 0 && K > 0 && ld >= K, "upsample_argmax: bad shape");
    const float rh = static_cast<float>(h) / static_cast<float>(H), rw = static_cast<float>(w) / static_cast<float>(W);
    upsample_argmax_kernel<<<loss_grid((long long)N * H * W), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        logits, N, h, w, K, ld, H, W, rh, rw, reinterpret_cast<long long*>(labels_out));
    GS_LAUNCHED();
    return 0;
}

extern "C" int gs_upsample_bilinear_f32(const float* src, int32_t N, int32_t h, int32_t w, int32_t K, int32_t ld,
                                        float* dst, int32_t H, int32_t W, int32_t dst_ld, void* stream) {
    GS_REQUIRE(src && dst, "upsample_bilinear: null pointer");
    GS_REQUIRE(N > 0 && h > 0 && w > 0 && H > 0 && W > 0 && K > 0 && ld >= K && dst_ld >= K, "upsample_bilinear: bad shape");
    const float rh = static_cast<float>(h) / static_cast<float>(H), rw = static_cast<float>(w) / static_cast<float>(W);
    upsample_bilinear_kernel<<<loss_grid((long long)N * H * W * K), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        src, N, h, w, K, ld, dst, H, W, dst_ld, rh, rw);
    GS_LAUNCHED();
    return 0;
}
