// gs_loss.cu -- fused bilinear-upsample + cross-entropy(ignore_index) + top-1 accuracy, its gradient,
// and fused upsample + argmax for inference.  The N*K*H*W up-sampled logit tensor of the reference
// (resize -> F.cross_entropy -> accuracy; gaiaseg/models/decode_heads/dynamic_fcn_head.py:137-159,
// CE restated at gaiaseg/models/losses/cross_entropy_loss.py:67-94, accuracy at accuracy.py:4-49)
// is never materialised:
//   forward : one thread per OUTPUT pixel, 4-tap lerp of the K low-res logits on the fly, online
//             soft-max (max / sum-exp / first arg-max in one sweep); writes an 8-byte record
//             (log-sum-exp, label or -1) per pixel for the backward pass.
//   backward: one CTA per TILE of low-res cells (all classes), the only writer of its dlogits: the
//             soft-max terms of every output pixel are computed once per CTA, the bilinear weights are
//             applied separably through shared memory, one coalesced store.  No atomics,
//             deterministic (details at the kernel below).
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
    pdl_sync();
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

// ------------------------------------------------------------------------------------------------
// backward.  dlogits[n, i, j, k] = grad_scale * sum over valid output pixels (y, x) of
//                                  wy(y, i) * wx(x, j) * ( exp(v_k(y, x) - lse(y, x)) - [label(y, x) == k] )
// One CTA owns a TILE of low-res cells (all classes) and is the only writer of its dlogits -> no atomics, fixed
// summation order (deterministic).  The tile's logits + a one-cell halo live in shared memory; every output pixel whose
// taps touch the tile has its soft-max terms computed ONCE per CTA (the round-1 kernel recomputed them per
// (cell, class) thread: ~4*19 times), and the bilinear weights are applied separably:
//   phase 1  thread <-> output pixel of a strip of R rows:  G[r][x][k] = exp(v_k - lse) - [label == k]
//   phase 2  thread <-> (row r, low-res column j, k):       T[r][j][k] = sum_x wx(x, j) * G[r][x][k]
//   phase 3  thread <-> (low-res cell i, j, k):             ACC[i][j][k] += sum_r wy(y_r, i) * T[r][j][k]
// then ONE coalesced store of ACC * grad_scale.  Classes are processed in chunks of <= kBwdKC (p_k only needs the
// recorded log-sum-exp, so chunks are independent): K = 150 fits the same shared-memory budget as K = 19.
// ------------------------------------------------------------------------------------------------
constexpr int kBwdTile = 8;        // low-res cells per tile edge
constexpr int kBwdKC = 32;         // classes per chunk
constexpr int kBwdThreads = 512;
constexpr int kBwdKSplit = 4;     // phase 1: threads per output pixel (classes interleaved) -> 4x the parallelism
constexpr int kBwdMaxRegion = 512; // output pixels per tile edge the tap tables can hold (scale factors up to ~50)

struct BwdSmem {
    // byte offsets into the dynamic shared memory block (all 16-byte aligned)
    int logits, G, T, acc, rec, xt_i, xt_w, yt_i, yt_w, xr, yr, total;
};

__host__ __device__ inline BwdSmem bwd_smem_layout(int K, int KC, int XW, int R) {
    BwdSmem L;
    const int kcp = KC | 1;   // odd pitch: phase 1 writes a pixel's classes with lane stride kcp -> conflict-free
    int off = 0;
    auto take = [&](int bytes) { const int o = off; off += (bytes + 15) & ~15; return o; };
    L.logits = take((kBwdTile + 2) * (kBwdTile + 2) * K * 4);
    L.G = take(R * XW * kcp * 4);
    L.T = take(R * kBwdTile * kcp * 4);
    L.acc = take(kBwdTile * kBwdTile * kcp * 4);
    L.rec = take(R * XW * 8);          // the strip's per-pixel records, staged with coalesced, independent loads
    L.xt_i = take(XW * 8);
    L.xt_w = take(XW * 8);
    L.yt_i = take(kBwdMaxRegion * 8);
    L.yt_w = take(kBwdMaxRegion * 8);
    L.xr = take(kBwdTile * 8);
    L.yr = take(kBwdTile * 8);
    L.total = off;
    return L;
}

__global__ void __launch_bounds__(kBwdThreads) upsample_ce_bwd_kernel(const float* __restrict__ logits, int N, int h, int w,
                                                                      int K, int ld, const PixRec* __restrict__ rec, int H,
                                                                      int W, float rh, float rw, float grad_scale,
                                                                      const float* __restrict__ grad_scale_dev,
                                                                      float* __restrict__ dlogits, int dl_ld, int XWmax,
                                                                      int R, int KC) {
    pdl_sync();
    extern __shared__ __align__(16) unsigned char smem_bwd[];
    const BwdSmem L = bwd_smem_layout(K, KC, XWmax, R);
    float* sL = reinterpret_cast<float*>(smem_bwd + L.logits);
    float* sG = reinterpret_cast<float*>(smem_bwd + L.G);
    float* sT = reinterpret_cast<float*>(smem_bwd + L.T);
    float* sA = reinterpret_cast<float*>(smem_bwd + L.acc);
    PixRec* sR = reinterpret_cast<PixRec*>(smem_bwd + L.rec);
    int2* xti = reinterpret_cast<int2*>(smem_bwd + L.xt_i);
    float2* xtw = reinterpret_cast<float2*>(smem_bwd + L.xt_w);
    int2* yti = reinterpret_cast<int2*>(smem_bwd + L.yt_i);
    float2* ytw = reinterpret_cast<float2*>(smem_bwd + L.yt_w);
    int2* xr = reinterpret_cast<int2*>(smem_bwd + L.xr);   // per tile column: [first, last] region index that touches it
    int2* yr = reinterpret_cast<int2*>(smem_bwd + L.yr);
    const int kcp = KC | 1;
    const int tid = threadIdx.x;
    if (grad_scale_dev) grad_scale *= __ldg(grad_scale_dev);

    const int tiles_w = (w + kBwdTile - 1) / kBwdTile, tiles_h = (h + kBwdTile - 1) / kBwdTile;
    int t = blockIdx.x;
    const int tj = t % tiles_w; t /= tiles_w;
    const int ti = t % tiles_h;
    const int n = t / tiles_h;
    const int i0 = ti * kBwdTile, j0 = tj * kBwdTile;
    const int th = min(kBwdTile, h - i0), tw = min(kBwdTile, w - j0);

    // output-pixel region whose taps can touch the tile (one extra pixel of margin on each side; pixels that turn out
    // not to touch it get zero weights below)
    const float inv_rh = 1.f / rh, inv_rw = 1.f / rw;
    int ylo = (int)floorf((i0 - 0.5f) * inv_rh - 0.5f) - 1;
    int yhi = (int)ceilf((i0 + th + 0.5f) * inv_rh - 0.5f) + 1;
    int xlo = (int)floorf((j0 - 0.5f) * inv_rw - 0.5f) - 1;
    int xhi = (int)ceilf((j0 + tw + 0.5f) * inv_rw - 0.5f) + 1;
    if (ylo < 0) ylo = 0;
    if (xlo < 0) xlo = 0;
    if (yhi > H - 1) yhi = H - 1;
    if (xhi > W - 1) xhi = W - 1;
    const int XW = xhi - xlo + 1, YH = yhi - ylo + 1;   // host guarantees XW <= XWmax, YH <= kBwdMaxRegion

    // tap tables + the tile's logits (with halo) -> shared memory
    for (int i = tid; i < XW; i += kBwdThreads) {
        const Tap tp = src_tap(rw, xlo + i, w);
        xti[i] = make_int2(tp.i0, tp.i1);
        xtw[i] = make_float2(tp.l0, tp.l1);
    }
    for (int i = tid; i < YH; i += kBwdThreads) {
        const Tap tp = src_tap(rh, ylo + i, h);
        yti[i] = make_int2(tp.i0, tp.i1);
        ytw[i] = make_float2(tp.l0, tp.l1);
    }
    constexpr int LW = kBwdTile + 2;
    for (int i = tid; i < LW * LW * K; i += kBwdThreads) {
        const int k = i % K;
        const int c = i / K;
        const int lj = c % LW, li = c / LW;
        const int gi = i0 - 1 + li, gj = j0 - 1 + lj;
        float v = 0.f;
        if (gi >= 0 && gi < h && gj >= 0 && gj < w) v = __ldg(logits + ((long long)(n * h + gi) * w + gj) * ld + k);
        sL[i] = v;
    }
    __syncthreads();
    if (tid < kBwdTile) {          // first / last region column touching tile column tid (taps are monotone in x)
        int a = XW, b = -1;
        for (int i = 0; i < XW; ++i)
            if (xti[i].x == j0 + tid || xti[i].y == j0 + tid) { if (i < a) a = i; b = i; }
        xr[tid] = make_int2(a, b);
    } else if (tid >= 32 && tid < 32 + kBwdTile) {
        const int c = tid - 32;
        int a = YH, b = -1;
        for (int i = 0; i < YH; ++i)
            if (yti[i].x == i0 + c || yti[i].y == i0 + c) { if (i < a) a = i; b = i; }
        yr[c] = make_int2(a, b);
    }
    __syncthreads();

    for (int kc = 0; kc < K; kc += KC) {
        const int kn = min(KC, K - kc);
        for (int i = tid; i < kBwdTile * kBwdTile * kcp; i += kBwdThreads) sA[i] = 0.f;
        for (int yb = 0; yb < YH; yb += R) {
            const int rows = min(R, YH - yb);
            // ---- phase 0: the strip's records -> shared memory (one independent load per thread: the round-trip to L2 is
            // paid once per strip instead of once per pixel of every thread's serial loop) ----
            if (kc == 0 || K > KC) {
                for (int p = tid; p < rows * XW; p += kBwdThreads) {
                    const int r = p / XW, xx = p - r * XW;
                    sR[p] = rec[((long long)n * H + (ylo + yb + r)) * W + (xlo + xx)];
                }
                __syncthreads();
            }
            // ---- phase 1: soft-max gradient terms of the strip, once per pixel (kBwdKSplit lanes share a pixel) ----
            for (int it = tid; it < rows * XW * kBwdKSplit; it += kBwdThreads) {
                const int q = it & (kBwdKSplit - 1);
                const int p = it / kBwdKSplit;
                const int r = p / XW, xx = p - r * XW;
                const int2 yi = yti[yb + r], xi = xti[xx];
                float* g = sG + (size_t)(r * XW + xx) * kcp;
                const PixRec pr = sR[p];
                const bool inside = yi.x >= i0 - 1 && yi.y <= i0 + th && xi.x >= j0 - 1 && xi.y <= j0 + tw;
                if (pr.label < 0 || !inside) {
                    for (int kk = q; kk < kn; kk += kBwdKSplit) g[kk] = 0.f;
                    continue;
                }
                const float2 wy = ytw[yb + r], wx = xtw[xx];
                Tap ty, tx;
                ty.l0 = wy.x; ty.l1 = wy.y; tx.l0 = wx.x; tx.l1 = wx.y;
                const float* a00 = sL + ((yi.x - i0 + 1) * LW + (xi.x - j0 + 1)) * K + kc;
                const float* a01 = sL + ((yi.x - i0 + 1) * LW + (xi.y - j0 + 1)) * K + kc;
                const float* a10 = sL + ((yi.y - i0 + 1) * LW + (xi.x - j0 + 1)) * K + kc;
                const float* a11 = sL + ((yi.y - i0 + 1) * LW + (xi.y - j0 + 1)) * K + kc;
                const int lab = pr.label - kc;
#pragma unroll 4
                for (int kk = q; kk < kn; kk += kBwdKSplit) {
                    const float v = lerp4(ty, tx, a00[kk], a01[kk], a10[kk], a11[kk]);
                    g[kk] = __expf(v - pr.lse) - (kk == lab ? 1.f : 0.f);
                }
            }
            __syncthreads();
            // ---- phase 2: horizontal taps: T[r][j][kk] = sum_x wx(x, j) * G[r][x][kk]  (x ascending: fixed order) ----
            for (int o = tid; o < rows * tw * kn; o += kBwdThreads) {
                const int kk = o % kn;
                const int c = o / kn;
                const int j = c % tw, r = c / tw;
                const int2 range = xr[j];
                float acc = 0.f;
                const float* g = sG + (size_t)(r * XW) * kcp + kk;
#pragma unroll 4
                for (int xx = range.x; xx <= range.y; ++xx) {
                    const int2 xi = xti[xx];
                    const float2 wx = xtw[xx];
                    const float wgt = (xi.x == j0 + j ? wx.x : 0.f) + (xi.y == j0 + j ? wx.y : 0.f);
                    acc = fmaf(wgt, g[(size_t)xx * kcp], acc);
                }
                sT[(r * kBwdTile + j) * kcp + kk] = acc;
            }
            __syncthreads();
            // ---- phase 3: vertical taps into the tile accumulators (rows ascending: fixed order) ----
            for (int o = tid; o < th * tw * kn; o += kBwdThreads) {
                const int kk = o % kn;
                const int c = o / kn;
                const int j = c % tw, i = c / tw;
                const int2 range = yr[i];
                const int ra = max(range.x - yb, 0), rb = min(range.y - yb, rows - 1);
                float acc = sA[(i * kBwdTile + j) * kcp + kk];
                for (int r = ra; r <= rb; ++r) {
                    const int2 yi = yti[yb + r];
                    const float2 wy = ytw[yb + r];
                    const float wgt = (yi.x == i0 + i ? wy.x : 0.f) + (yi.y == i0 + i ? wy.y : 0.f);
                    acc = fmaf(wgt, sT[(r * kBwdTile + j) * kcp + kk], acc);
                }
                sA[(i * kBwdTile + j) * kcp + kk] = acc;
            }
            __syncthreads();
        }
        // ---- store the chunk: consecutive threads -> consecutive classes of one cell ----
        for (int o = tid; o < th * tw * kn; o += kBwdThreads) {
            const int kk = o % kn;
            const int c = o / kn;
            const int j = c % tw, i = c / tw;
            dlogits[((long long)(n * h + i0 + i) * w + (j0 + j)) * dl_ld + kc + kk] = sA[(i * kBwdTile + j) * kcp + kk] * grad_scale;
        }
        __syncthreads();
    }
}

// fused upsample + argmax (first maximum wins); optional two-step resize is done by the caller with
// gs_upsample_bilinear_f32 first.
__global__ void __launch_bounds__(256) upsample_argmax_kernel(const float* __restrict__ logits, int N, int h, int w,
                                                              int K, int ld, int H, int W, float rh, float rw,
                                                              long long* __restrict__ out) {
    pdl_sync();
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
    pdl_sync();
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
    gs::launch(upsample_ce_fwd_kernel, dim3(loss_grid((long long)N * H * W)), dim3(256), 0, static_cast<cudaStream_t>(stream), 
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
    // widest output-pixel region a tile can touch (same arithmetic as the kernel, + slack), strip height R and class chunk
    // sized for <= ~110 KB of shared memory per CTA (2 CTAs of 512 threads per SM)
    const int xw_max = static_cast<int>(ceilf((kBwdTile + 1.5f) / rw)) + 6;
    const int yh_max = static_cast<int>(ceilf((kBwdTile + 1.5f) / rh)) + 6;
    GS_REQUIRE(xw_max <= kBwdMaxRegion && yh_max <= kBwdMaxRegion,
               "upsample_ce_bwd: scale factor %dx%d -> %dx%d too large for the tile tables", h, w, H, W);
    const int KC = K < kBwdKC ? K : kBwdKC;
    int R = (64 * 1024) / (xw_max * (KC | 1) * 4);
    if (R < 1) R = 1;
    if (R > 16) R = 16;
    const BwdSmem L = bwd_smem_layout(K, KC, xw_max, R);
    GS_REQUIRE(L.total <= 200 * 1024, "upsample_ce_bwd: %d classes need %d bytes of shared memory", K, L.total);
    static int attr_bytes = 0;
    if (L.total > attr_bytes) {
        GS_CUDA_OK(cudaFuncSetAttribute(upsample_ce_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));
        attr_bytes = L.total;
    }
    const int tiles = N * ((h + kBwdTile - 1) / kBwdTile) * ((w + kBwdTile - 1) / kBwdTile);
    gs::launch(upsample_ce_bwd_kernel, dim3(tiles), dim3(kBwdThreads), L.total, static_cast<cudaStream_t>(stream), 
        logits, N, h, w, K, ld, reinterpret_cast<const PixRec*>(pix_rec), H, W, rh, rw, grad_scale, grad_scale_dev, dlogits,
        dl_ld, xw_max, R, KC);
    GS_LAUNCHED();
    return 0;
}

extern "C" int gs_upsample_argmax(const float* logits, int32_t N, int32_t h, int32_t w, int32_t K, int32_t ld,
                                  int32_t H, int32_t W, int64_t* labels_out, void* stream) {
    GS_REQUIRE(logits && labels_out, "upsample_argmax: null pointer");
    GS_REQUIRE(N > 0 && h > 0 && w > 0 && H > 0 && W > 0 && K > 0 && ld >= K, "upsample_argmax: bad shape");
    const float rh = static_cast<float>(h) / static_cast<float>(H), rw = static_cast<float>(w) / static_cast<float>(W);
    gs::launch(upsample_argmax_kernel, dim3(loss_grid((long long)N * H * W)), dim3(256), 0, static_cast<cudaStream_t>(stream), 
        logits, N, h, w, K, ld, H, W, rh, rw, reinterpret_cast<long long*>(labels_out));
    GS_LAUNCHED();
    return 0;
}

extern "C" int gs_upsample_bilinear_f32(const float* src, int32_t N, int32_t h, int32_t w, int32_t K, int32_t ld,
                                        float* dst, int32_t H, int32_t W, int32_t dst_ld, void* stream) {
    GS_REQUIRE(src && dst, "upsample_bilinear: null pointer");
    GS_REQUIRE(N > 0 && h > 0 && w > 0 && H > 0 && W > 0 && K > 0 && ld >= K && dst_ld >= K, "upsample_bilinear: bad shape");
    const float rh = static_cast<float>(h) / static_cast<float>(H), rw = static_cast<float>(w) / static_cast<float>(W);
    gs::launch(upsample_bilinear_kernel, dim3(loss_grid((long long)N * H * W * K)), dim3(256), 0, static_cast<cudaStream_t>(stream), 
        src, N, h, w, K, ld, dst, H, W, dst_ld, rh, rw);
    GS_LAUNCHED();
    return 0;
}
