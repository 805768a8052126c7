// gs_norm.cu -- DynamicBatchNorm2d / DynamicSyncBatchNorm kernels (memory-bound, sm_100a).
//
// Replaces [EXT] gaiavision DynBN / DynSyncBN == F.batch_norm on the channel-prefix slice
// (sites: gaiaseg/models/backbones/dynamic_resnet.py:267-300, gaiaseg/models/utils/dynamic_res_layer.py:92).
// Statistics travel as fp64 [2C] = (sum, sum of squares): the conv epilogue or gs_bn_stats
// produces them, the host all-reduces them over the SyncBN group (packed, one message per layer),
// gs_bn_finalize turns them into mean / invstd / scale / shift and updates the running stats.
//
// Algorithmic bytes (E = P*C elements, bf16):  stats E*2 | apply 2E*2 (+E*2 residual)
//   bwd_reduce 3E*2 (dz, y, z) | bwd_apply 4E*2 (+E*2 when the residual gradient is written).
#include "../../include/gaiaseg_b200.h"
#include "gs_host.h"
#include "gs_vec.cuh"
#include "gs_comm.cuh"

#include <stdlib.h>

namespace gs {

// ------------------------------------------------------------------------------------------------
// block reduction of 16 per-thread partials over the R pixel rows of a ColMap block, then fp64 atomics
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void block_reduce16_atomic(float (&a)[8], float (&b)[8], float* red, int Vc, int R, int cv0,
                                                      int C8, int C, double* out) {
    const int tid = threadIdx.x;
    __syncthreads();  // previous chunk finished reading `red`
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        red[tid * 16 + i] = a[i];
        red[tid * 16 + 8 + i] = b[i];
    }
    __syncthreads();
    const int nthreads = Vc * R;
    for (int o = tid; o < 16 * Vc; o += nthreads) {
        const int cx = o >> 4, i = o & 15;
        const int cv = cv0 + cx;
        if (cv >= C8) continue;
        float s = 0.f;
        for (int r = 0; r < R; ++r) s += red[(r * Vc + cx) * 16 + i];
        const int c = cv * 8 + (i & 7);
        atomicAdd(out + (i < 8 ? c : C + c), static_cast<double>(s));
    }
}

// ------------------------------------------------------------------------------------------------
// stats: sum / sumsq
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bn_stats_kernel(const uint4* __restrict__ x, long long ld8, long long P, int C,
                                                       int C8, int Vc, int R, double* __restrict__ stats) {
    pdl_sync();
    __shared__ float red[256 * 16];
    const int cx = threadIdx.x % Vc, ry = threadIdx.x / Vc;
    const long long S = (long long)gridDim.x * R;
    for (int cv0 = 0; cv0 < C8; cv0 += Vc) {
        const int cv = cv0 + cx;
        float s[8], q[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { s[i] = 0.f; q[i] = 0.f; }
        if (cv < C8) {
            long long p = (long long)blockIdx.x * R + ry;
            for (; p + 3 * S < P; p += 4 * S) {
                uint4 v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) v[u] = ldg_stream(x + (p + u * S) * ld8 + cv);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    float f[8];
                    unpack8(v[u], f);
#pragma unroll
                    for (int i = 0; i < 8; ++i) { s[i] += f[i]; q[i] = fmaf(f[i], f[i], q[i]); }
                }
            }
            for (; p < P; p += S) {
                float f[8];
                unpack8(ldg_stream(x + p * ld8 + cv), f);
#pragma unroll
                for (int i = 0; i < 8; ++i) { s[i] += f[i]; q[i] = fmaf(f[i], f[i], q[i]); }
            }
        }
        block_reduce16_atomic(s, q, red, Vc, R, cv0, C8, C, stats);
    }
}

// ------------------------------------------------------------------------------------------------
// finalize / eval affine (tiny)
// ------------------------------------------------------------------------------------------------
__global__ void bn_finalize_kernel(const double* __restrict__ stats, double count, int C, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* __restrict__ rm, float* __restrict__ rv,
                                   float momentum, float eps, float* __restrict__ mean, float* __restrict__ invstd,
                                   float* __restrict__ scale, float* __restrict__ shift) {
    pdl_sync();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const double m = stats[c] / count;
    double var = stats[C + c] / count - m * m;
    if (var < 0.0) var = 0.0;
    const float mf = static_cast<float>(m);
    const float istd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
    const float g = gamma ? gamma[c] : 1.f;
    const float b = beta ? beta[c] : 0.f;
    const float sc = g * istd;
    if (mean) mean[c] = mf;
    if (invstd) invstd[c] = istd;
    scale[c] = sc;
    shift[c] = b - mf * sc;
    if (rm) rm[c] = (1.f - momentum) * rm[c] + momentum * mf;
    if (rv) {
        const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
        rv[c] = (1.f - momentum) * rv[c] + momentum * static_cast<float>(unbiased);
    }
}

__global__ void bn_eval_affine_kernel(int C, const float* __restrict__ gamma, const float* __restrict__ beta,
                                      const float* __restrict__ rm, const float* __restrict__ rv, float eps,
                                      float* __restrict__ scale, float* __restrict__ shift) {
    pdl_sync();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const float g = gamma ? gamma[c] : 1.f;
    const float b = beta ? beta[c] : 0.f;
    const float sc = g * rsqrtf(rv[c] + eps);
    scale[c] = sc;
    shift[c] = b - rm[c] * sc;
}

// ------------------------------------------------------------------------------------------------
// apply: z = relu?( y*scale + shift (+ res) )
// FROM_STATS: scale / shift are derived in the kernel prologue from the (all-reduced) fp64 sums -- the
// "finalize" step costs no launch; block 0 also stores mean / invstd / scale / shift for the backward pass and
// updates the running statistics of the channel prefix.
// ------------------------------------------------------------------------------------------------
struct BnTrainArgs {
    const double* stats;   // [2C] sum, sumsq
    double inv_count;      // 1 / elements per channel (over the SyncBN group)
    double unbias;         // count / (count - 1)
    const float* gamma;    // may be NULL
    const float* beta;     // may be NULL
    float* rm;             // running mean / var prefix, may be NULL
    float* rv;
    float momentum, eps;
    float* aff;            // [4][C]: mean, invstd, scale, shift
    int C;
    // several ranks: block 0 runs the SyncBN exchange of `stats` itself (in place) and then raises flag[0]; the other
    // blocks wait for it before they read the sums -- no separate exchange launch on the critical chain
    SyncArgs sync;
    unsigned long long* flag;   // zero-initialised device word (NULL when sync.world <= 1)
};

template <bool HAS_RES, bool FROM_STATS>
__global__ void __launch_bounds__(256) bn_apply_kernel(const uint4* __restrict__ y, long long y_ld8,
                                                       const float* __restrict__ scale, const float* __restrict__ shift,
                                                       const uint4* __restrict__ res, long long res_ld8, int relu,
                                                       uint4* __restrict__ z, long long z_ld8, long long P, int C8,
                                                       int Vc, int R, const __grid_constant__ BnTrainArgs t) {
    pdl_sync();
    const int cx = threadIdx.x % Vc, ry = threadIdx.x / Vc;
    const long long S = (long long)gridDim.x * R;
    if (FROM_STATS && t.sync.world > 1) {
        if (blockIdx.x == 0) {
            syncbn_exchange_block(const_cast<double*>(t.stats), 2 * t.C, t.sync.peers, t.sync.rank, t.sync.world,
                                  t.sync.seq_dev, nullptr, nullptr, t.sync.timeout_ns, threadIdx.x, blockDim.x,
                                  t.sync.phase != 2);   // phase 2: the conv kernel's last CTA has pushed the sums
            __threadfence();
            __syncthreads();
            if (threadIdx.x == 0) st_release_gpu(t.flag, 1ull);
        } else {
            if (threadIdx.x == 0) {
                while (ld_acquire_gpu(t.flag) == 0ull) {}
            }
            __syncthreads();
        }
    }
    for (int cv = cx; cv < C8; cv += Vc) {
        // software pipeline: the first batch of loads is issued BEFORE the per-channel prologue (it does not depend on the
        // statistics), and inside the loop the next batch is in flight while the current one is normalised and stored --
        // 8 x 16 bytes per tensor and thread in flight instead of 4, and the prologue's latency (fp64 sums -> scale / shift)
        // overlaps the first memory round trip
        long long p = (long long)blockIdx.x * R + ry;
        uint4 v[4], rr[4];
        bool have = p + 3 * S < P;
        if (have) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                v[u] = ldg_stream(y + (p + u * S) * y_ld8 + cv);
                if (HAS_RES) rr[u] = ldg_stream(res + (p + u * S) * res_ld8 + cv);
            }
        }
        float sc[8], sh[8];
        if (FROM_STATS) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int c = cv * 8 + i;
                // (several ranks: block 0 has just rewritten the sums -> read them through L2; single rank: plain cached loads --
                // every thread row of the block reads the same channels, and .cg made this prologue 17 % slower)
                const double s1 = t.sync.world > 1 ? __ldcg(t.stats + c) : t.stats[c];
                const double s2 = t.sync.world > 1 ? __ldcg(t.stats + t.C + c) : t.stats[t.C + c];
                const double m = s1 * t.inv_count;
                double var = s2 * t.inv_count - m * m;
                if (var < 0.0) var = 0.0;
                const float mf = static_cast<float>(m);
                const float istd = rsqrtf(static_cast<float>(var) + t.eps);   // fp64 only where cancellation can occur
                const float g = t.gamma ? __ldg(t.gamma + c) : 1.f;
                const float b = t.beta ? __ldg(t.beta + c) : 0.f;
                sc[i] = g * istd;
                sh[i] = b - mf * sc[i];
                if (blockIdx.x == 0 && ry == 0) {
                    t.aff[c] = mf;
                    t.aff[t.C + c] = istd;
                    t.aff[2 * t.C + c] = sc[i];
                    t.aff[3 * t.C + c] = sh[i];
                    if (t.rm) t.rm[c] = (1.f - t.momentum) * t.rm[c] + t.momentum * mf;
                    if (t.rv) t.rv[c] = (1.f - t.momentum) * t.rv[c] + t.momentum * static_cast<float>(var * t.unbias);
                }
            }
        } else {
            if (scale) load8f(scale + cv * 8, sc);
            else {
#pragma unroll
                for (int i = 0; i < 8; ++i) sc[i] = 1.f;
            }
            if (shift) load8f(shift + cv * 8, sh);
            else {
#pragma unroll
                for (int i = 0; i < 8; ++i) sh[i] = 0.f;
            }
        }
        while (have) {
            const long long pn = p + 4 * S;
            const bool hn = pn + 3 * S < P;
            uint4 v2[4], rr2[4];
            if (hn) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    v2[u] = ldg_stream(y + (pn + u * S) * y_ld8 + cv);
                    if (HAS_RES) rr2[u] = ldg_stream(res + (pn + u * S) * res_ld8 + cv);
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                float f[8], g[8];
                unpack8(v[u], f);
                if (HAS_RES) unpack8(rr[u], g);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float tt = fmaf(f[i], sc[i], sh[i]);
                    if (HAS_RES) tt += g[i];
                    f[i] = relu ? fmaxf(tt, 0.f) : tt;
                }
                stg_stream(z + (p + u * S) * z_ld8 + cv, pack8(f));
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                v[u] = v2[u];
                if (HAS_RES) rr[u] = rr2[u];
            }
            p = pn;
            have = hn;
        }
        for (; p < P; p += S) {
            float f[8], g[8];
            unpack8(ldg_stream(y + p * y_ld8 + cv), f);
            if (HAS_RES) unpack8(ldg_stream(res + p * res_ld8 + cv), g);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float tt = fmaf(f[i], sc[i], sh[i]);
                if (HAS_RES) tt += g[i];
                f[i] = relu ? fmaxf(tt, 0.f) : tt;
            }
            stg_stream(z + p * z_ld8 + cv, pack8(f));
        }
    }
}

// ------------------------------------------------------------------------------------------------
// SyncBN exchange halves of the two backward kernels (several ranks, GS_SYNCBN_FOLD=2).  Out of line on purpose: they run
// once per kernel, and inlined they changed the load scheduling of the streaming loops (fewer loads in flight: the
// reduction kernels got 20 % slower on a single rank).
// ------------------------------------------------------------------------------------------------
// send half: the last block to finish sees the rank's final sums and pushes them to the peers, so the NVLink flight
// overlaps this kernel's tail and the launch of gs_bn_bwd_apply (which polls)
__device__ __noinline__ void bn_reduce_push_tail(const SyncArgs& sync, const double* sums, int C, unsigned long long* ticket) {
    __shared__ unsigned int is_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(ticket, 1ull) + 1 == static_cast<unsigned long long>(gridDim.x)) ? 1u : 0u;
    __syncthreads();
    if (is_last) {
        __threadfence();
        syncbn_push_block(sums, 2 * C, sync.peers, sync.rank, sync.world, sync.seq_dev, threadIdx.x, blockDim.x);
    }
}
// receive half: block 0 adds the parameter gradients from the LOCAL sums, polls the peers' contributions, writes the group
// sums in place and releases the other blocks
__device__ __noinline__ void bn_apply_poll_head(const SyncArgs& sync, double* sums, int C, float* dgamma, float* dbeta,
                                                unsigned long long* flag) {
    if (blockIdx.x == 0) {
        syncbn_exchange_block(sums, 2 * C, sync.peers, sync.rank, sync.world, sync.seq_dev, dgamma, dbeta, sync.timeout_ns,
                              threadIdx.x, blockDim.x, sync.phase != 2);
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) st_release_gpu(flag, 1ull);
    } else {
        if (threadIdx.x == 0) {
            while (ld_acquire_gpu(flag) == 0ull) {}
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// backward pass 1: per-channel sums of g and g*xhat, g = dz * [z > 0]
// ------------------------------------------------------------------------------------------------
// MASK: 0 = no activation, 1 = ReLU mask from the stored output z (needed when a residual was added),
//       2 = ReLU mask recomputed from y: [fma(y, scale, shift) > 0] -- bit-identical to the forward's test and one
//           tensor read cheaper.
// SYNC: several ranks with the split exchange (the single-rank instantiation carries no exchange code at all)
template <int MASK, bool SYNC>
__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(const uint4* __restrict__ dz, long long dz_ld8,
                                                            const uint4* __restrict__ y, long long y_ld8,
                                                            const uint4* __restrict__ z, long long z_ld8,
                                                            const float* __restrict__ mean,
                                                            const float* __restrict__ invstd,
                                                            const float* __restrict__ scale,
                                                            const float* __restrict__ shift, long long P, int C,
                                                            int C8, int Vc, int R, double* __restrict__ sums,
                                                            const __grid_constant__ SyncArgs sync,
                                                            unsigned long long* __restrict__ ticket) {
    pdl_sync();
    constexpr bool HAS_Z = (MASK == 1);
    __shared__ float red[256 * 16];
    const int cx = threadIdx.x % Vc, ry = threadIdx.x / Vc;
    const long long S = (long long)gridDim.x * R;
    for (int cv0 = 0; cv0 < C8; cv0 += Vc) {
        const int cv = cv0 + cx;
        float sg[8], sx[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { sg[i] = 0.f; sx[i] = 0.f; }
        if (cv < C8) {
            float mu[8], is[8], sc[8], sh[8];
            load8f(mean + cv * 8, mu);
            load8f(invstd + cv * 8, is);
            if (MASK == 2) { load8f(scale + cv * 8, sc); load8f(shift + cv * 8, sh); }
            long long p = (long long)blockIdx.x * R + ry;
            for (; p + 3 * S < P; p += 4 * S) {
                uint4 a[4], b[4], c[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    a[u] = ldg_stream(dz + (p + u * S) * dz_ld8 + cv);
                    b[u] = ldg_stream(y + (p + u * S) * y_ld8 + cv);
                    if (HAS_Z) c[u] = ldg_stream(z + (p + u * S) * z_ld8 + cv);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    float g[8], yy[8], zz[8];
                    unpack8(a[u], g);
                    unpack8(b[u], yy);
                    if (HAS_Z) unpack8(c[u], zz);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        bool dead = HAS_Z && !(zz[i] > 0.f);
                        if (MASK == 2) dead = !(fmaf(yy[i], sc[i], sh[i]) > 0.f);
                        const float gi = dead ? 0.f : g[i];
                        sg[i] += gi;
                        sx[i] = fmaf(gi, (yy[i] - mu[i]) * is[i], sx[i]);
                    }
                }
            }
            for (; p < P; p += S) {
                float g[8], yy[8], zz[8];
                unpack8(ldg_stream(dz + p * dz_ld8 + cv), g);
                unpack8(ldg_stream(y + p * y_ld8 + cv), yy);
                if (HAS_Z) unpack8(ldg_stream(z + p * z_ld8 + cv), zz);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    bool dead = HAS_Z && !(zz[i] > 0.f);
                    if (MASK == 2) dead = !(fmaf(yy[i], sc[i], sh[i]) > 0.f);
                    const float gi = dead ? 0.f : g[i];
                    sg[i] += gi;
                    sx[i] = fmaf(gi, (yy[i] - mu[i]) * is[i], sx[i]);
                }
            }
        }
        block_reduce16_atomic(sg, sx, red, Vc, R, cv0, C8, C, sums);
    }
    if (SYNC) bn_reduce_push_tail(sync, sums, C, ticket);
}

// ------------------------------------------------------------------------------------------------
// backward pass 2: dy = gamma*invstd*( g - sum_g/count - xhat*sum_gx/count ); dres = g
// ------------------------------------------------------------------------------------------------
template <int MASK, bool HAS_DRES, bool SYNC>
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const uint4* __restrict__ dz, long long dz_ld8,
                                                           const uint4* __restrict__ y, long long y_ld8,
                                                           const uint4* __restrict__ z, long long z_ld8,
                                                           const float* __restrict__ mean,
                                                           const float* __restrict__ invstd,
                                                           const float* __restrict__ scale,
                                                           const float* __restrict__ shift,
                                                           const float* __restrict__ gamma,
                                                           const double* __restrict__ sums, double inv_count,
                                                           long long P, int C, int C8, int Vc, int R,
                                                           uint4* __restrict__ dy, long long dy_ld8,
                                                           uint4* __restrict__ dres, long long dres_ld8,
                                                           float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                           const __grid_constant__ SyncArgs sync,
                                                           unsigned long long* __restrict__ flag) {
    pdl_sync();
    constexpr bool HAS_Z = (MASK == 1);
    const int cx = threadIdx.x % Vc, ry = threadIdx.x / Vc;
    const long long S = (long long)gridDim.x * R;
    if (SYNC) {
        // several ranks: `sums` holds the LOCAL sums and gs_bn_bwd_reduce's last block has pushed them.  Block 0 adds
        // the parameter gradients from the local sums, polls the peers' contributions, writes the group sums in place
        // and releases the other blocks (no exchange launch between the two passes).
        bn_apply_poll_head(sync, const_cast<double*>(sums), C, dgamma, dbeta, flag);
        dgamma = nullptr;
        dbeta = nullptr;
    }
    for (int cv = cx; cv < C8; cv += Vc) {
        float mu[8], is[8], k0[8], k1[8], k2[8], sc[8], sh[8];
        load8f(mean + cv * 8, mu);
        load8f(invstd + cv * 8, is);
        if (MASK == 2) { load8f(scale + cv * 8, sc); load8f(shift + cv * 8, sh); }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int c = cv * 8 + i;
            const float g = gamma ? __ldg(gamma + c) : 1.f;
            k0[i] = g * is[i];                                               // gamma * invstd
            // (several ranks: block 0 has just rewritten the sums -> read them through L2)
            k1[i] = static_cast<float>((SYNC ? __ldcg(sums + c) : sums[c]) * inv_count);           // mean of g
            k2[i] = static_cast<float>((SYNC ? __ldcg(sums + C + c) : sums[C + c]) * inv_count);   // mean of g*xhat
            if (blockIdx.x == 0 && ry == 0) {   // parameter gradients (single-rank case: local sums == group sums)
                if (dgamma) dgamma[c] += static_cast<float>(sums[C + c]);
                if (dbeta) dbeta[c] += static_cast<float>(sums[c]);
            }
        }
        long long p = (long long)blockIdx.x * R + ry;
        for (; p + 3 * S < P; p += 4 * S) {
            uint4 va[4], vb[4], vc[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                va[u] = ldg_stream(dz + (p + u * S) * dz_ld8 + cv);
                vb[u] = ldg_stream(y + (p + u * S) * y_ld8 + cv);
                if (HAS_Z) vc[u] = ldg_stream(z + (p + u * S) * z_ld8 + cv);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                float g[8], yy[8], zz[8], o[8];
                unpack8(va[u], g);
                unpack8(vb[u], yy);
                if (HAS_Z) unpack8(vc[u], zz);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    bool dead = HAS_Z && !(zz[i] > 0.f);
                    if (MASK == 2) dead = !(fmaf(yy[i], sc[i], sh[i]) > 0.f);
                    const float gi = dead ? 0.f : g[i];
                    g[i] = gi;
                    const float xh = (yy[i] - mu[i]) * is[i];
                    o[i] = k0[i] * (gi - k1[i] - xh * k2[i]);
                }
                stg_stream(dy + (p + u * S) * dy_ld8 + cv, pack8(o));
                if (HAS_DRES) stg_stream(dres + (p + u * S) * dres_ld8 + cv, pack8(g));
            }
        }
        for (; p < P; p += S) {
            float g[8], yy[8], zz[8], o[8];
            unpack8(ldg_stream(dz + p * dz_ld8 + cv), g);
            unpack8(ldg_stream(y + p * y_ld8 + cv), yy);
            if (HAS_Z) unpack8(ldg_stream(z + p * z_ld8 + cv), zz);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                bool dead = HAS_Z && !(zz[i] > 0.f);
                if (MASK == 2) dead = !(fmaf(yy[i], sc[i], sh[i]) > 0.f);
                const float gi = dead ? 0.f : g[i];
                g[i] = gi;
                const float xh = (yy[i] - mu[i]) * is[i];
                o[i] = k0[i] * (gi - k1[i] - xh * k2[i]);
            }
            stg_stream(dy + p * dy_ld8 + cv, pack8(o));
            if (HAS_DRES) stg_stream(dres + p * dres_ld8 + cv, pack8(g));
        }
    }
}

// ------------------------------------------------------------------------------------------------
// fused backward: reduce -> grid barrier (-> SyncBN exchange by block 0) -> apply, ONE cooperative launch.
// Halves the BN-backward launches (2 x ~360 per sandwich cycle of 10-30 us kernels whose time is mostly launch ramp and
// tail), the second pass over dz / y comes from L2, and with several ranks the exchange needs no launch of its own.
// scratch: two zero-initialised 64-bit words: [0] arrival counter of the grid barrier, [1] "sums final" flag.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void block_reduce8_atomic(const float (&a)[8], float* red, int Vc, int R, int cv0, int C8,
                                                     double* out) {
    const int tid = threadIdx.x;
    __syncthreads();  // previous round finished reading `red`
#pragma unroll
    for (int i = 0; i < 8; ++i) red[tid * 8 + i] = a[i];
    __syncthreads();
    const int nthreads = Vc * R;
    for (int o = tid; o < 8 * Vc; o += nthreads) {
        const int cx = o >> 3, i = o & 7;
        const int cv = cv0 + cx;
        if (cv >= C8) continue;
        float s = 0.f;
        for (int r = 0; r < R; ++r) s += red[(r * Vc + cx) * 8 + i];
        atomicAdd(out + cv * 8 + i, static_cast<double>(s));
    }
}

template <int MASK, bool HAS_DRES>
__global__ void __launch_bounds__(256, 2) bn_bwd_fused_kernel(const uint4* __restrict__ dz, long long dz_ld8,
                                                           const uint4* __restrict__ y, long long y_ld8,
                                                           const uint4* __restrict__ z, long long z_ld8,
                                                           const float* __restrict__ mean,
                                                           const float* __restrict__ invstd,
                                                           const float* __restrict__ scale,
                                                           const float* __restrict__ shift,
                                                           const float* __restrict__ gamma, double* __restrict__ sums,
                                                           double inv_count, long long P, int C, int C8, int Vc, int R,
                                                           uint4* __restrict__ dy, long long dy_ld8,
                                                           uint4* __restrict__ dres, long long dres_ld8,
                                                           float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                           const __grid_constant__ SyncArgs sync,
                                                           unsigned long long* __restrict__ scratch) {
    pdl_sync();
    constexpr bool HAS_Z = (MASK == 1);
    __shared__ float red[256 * 8];
    const int cx = threadIdx.x % Vc, ry = threadIdx.x / Vc;
    const long long S = (long long)gridDim.x * R;
    // ---------------- phase 1: per-channel sums of g and g * xhat ----------------
    for (int cv0 = 0; cv0 < C8; cv0 += Vc) {
        const int cv = cv0 + cx;
        float sg[8], sx[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { sg[i] = 0.f; sx[i] = 0.f; }
        if (cv < C8) {
            float mu[8], is[8], sc[8], sh[8];
            load8f(mean + cv * 8, mu);
            load8f(invstd + cv * 8, is);
            if (MASK == 2) { load8f(scale + cv * 8, sc); load8f(shift + cv * 8, sh); }
            long long p = (long long)blockIdx.x * R + ry;
            for (; p + 3 * S < P; p += 4 * S) {
                uint4 a[4], b[4], c[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    a[u] = __ldg(dz + (p + u * S) * dz_ld8 + cv);      // (default caching: phase 2 re-reads these lines)
                    b[u] = __ldg(y + (p + u * S) * y_ld8 + cv);
                    if (HAS_Z) c[u] = __ldg(z + (p + u * S) * z_ld8 + cv);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    float g[8], yy[8], zz[8];
                    unpack8(a[u], g);
                    unpack8(b[u], yy);
                    if (HAS_Z) unpack8(c[u], zz);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        bool dead = HAS_Z && !(zz[i] > 0.f);
                        if (MASK == 2) dead = !(fmaf(yy[i], sc[i], sh[i]) > 0.f);
                        const float gi = dead ? 0.f : g[i];
                        sg[i] += gi;
                        sx[i] = fmaf(gi, (yy[i] - mu[i]) * is[i], sx[i]);
                    }
                }
            }
            for (; p < P; p += S) {
                float g[8], yy[8], zz[8];
                unpack8(__ldg(dz + p * dz_ld8 + cv), g);
                unpack8(__ldg(y + p * y_ld8 + cv), yy);
                if (HAS_Z) unpack8(__ldg(z + p * z_ld8 + cv), zz);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    bool dead = HAS_Z && !(zz[i] > 0.f);
                    if (MASK == 2) dead = !(fmaf(yy[i], sc[i], sh[i]) > 0.f);
                    const float gi = dead ? 0.f : g[i];
                    sg[i] += gi;
                    sx[i] = fmaf(gi, (yy[i] - mu[i]) * is[i], sx[i]);
                }
            }
        }
        block_reduce8_atomic(sg, red, Vc, R, cv0, C8, sums);
        block_reduce8_atomic(sx, red, Vc, R, cv0, C8, sums + C);
    }
    // ---------------- grid barrier; block 0 finalises the sums ----------------
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) atomicAdd(scratch, 1ull);
    if (blockIdx.x == 0) {
        if (threadIdx.x == 0) {
            const unsigned long long t0 = gtimer();
            unsigned int spins = 0;
            while (ld_acquire_gpu(scratch) < static_cast<unsigned long long>(gridDim.x)) {
                if ((++spins & 4095u) == 0 && gtimer() - t0 > 5000000000ull) {
                    printf("gaiaseg_b200: bn_bwd grid barrier timed out (%llu of %u blocks arrived)\n", ld_acquire_gpu(scratch),
                           gridDim.x);
                    __trap();
                }
            }
        }
        __syncthreads();
        if (sync.world > 1) {
            // LOCAL sums -> parameter gradients, then the exchange over NVLink peer memory (sums := sums over all ranks)
            syncbn_exchange_block(sums, 2 * C, sync.peers, sync.rank, sync.world, sync.seq_dev, dgamma, dbeta,
                                  sync.timeout_ns, threadIdx.x, blockDim.x);
        } else {
            for (int c = threadIdx.x; c < C; c += blockDim.x) {
                if (dgamma) dgamma[c] += static_cast<float>(__ldcg(sums + C + c));
                if (dbeta) dbeta[c] += static_cast<float>(__ldcg(sums + c));
            }
        }
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) st_release_gpu(scratch + 1, 1ull);
    } else {
        if (threadIdx.x == 0) {
            const unsigned long long t0 = gtimer();
            unsigned int spins = 0;
            while (ld_acquire_gpu(scratch + 1) == 0ull) {
                if ((++spins & 4095u) == 0 && gtimer() - t0 > (sync.world > 1 ? sync.timeout_ns + 5000000000ull : 5000000000ull)) {
                    printf("gaiaseg_b200: bn_bwd waited too long for block 0\n");
                    __trap();
                }
            }
        }
        __syncthreads();
    }
    // ---------------- phase 2: dy = gamma*invstd*( g - mean(g) - xhat*mean(g*xhat) ); dres = g ----------------
    for (int cv = cx; cv < C8; cv += Vc) {
        float mu[8], is[8], k0[8], k1[8], k2[8], sc[8], sh[8];
        load8f(mean + cv * 8, mu);
        load8f(invstd + cv * 8, is);
        if (MASK == 2) { load8f(scale + cv * 8, sc); load8f(shift + cv * 8, sh); }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int c = cv * 8 + i;
            const float g = gamma ? __ldg(gamma + c) : 1.f;
            k0[i] = g * is[i];
            k1[i] = static_cast<float>(__ldcg(sums + c) * inv_count);
            k2[i] = static_cast<float>(__ldcg(sums + C + c) * inv_count);
        }
        long long p = (long long)blockIdx.x * R + ry;
        for (; p + 3 * S < P; p += 4 * S) {
            uint4 va[4], vb[4], vc[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                va[u] = ldg_stream(dz + (p + u * S) * dz_ld8 + cv);
                vb[u] = ldg_stream(y + (p + u * S) * y_ld8 + cv);
                if (HAS_Z) vc[u] = ldg_stream(z + (p + u * S) * z_ld8 + cv);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                float g[8], yy[8], zz[8], o[8];
                unpack8(va[u], g);
                unpack8(vb[u], yy);
                if (HAS_Z) unpack8(vc[u], zz);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    bool dead = HAS_Z && !(zz[i] > 0.f);
                    if (MASK == 2) dead = !(fmaf(yy[i], sc[i], sh[i]) > 0.f);
                    const float gi = dead ? 0.f : g[i];
                    g[i] = gi;
                    const float xh = (yy[i] - mu[i]) * is[i];
                    o[i] = k0[i] * (gi - k1[i] - xh * k2[i]);
                }
                stg_stream(dy + (p + u * S) * dy_ld8 + cv, pack8(o));
                if (HAS_DRES) stg_stream(dres + (p + u * S) * dres_ld8 + cv, pack8(g));
            }
        }
        for (; p < P; p += S) {
            float g[8], yy[8], zz[8], o[8];
            unpack8(ldg_stream(dz + p * dz_ld8 + cv), g);
            unpack8(ldg_stream(y + p * y_ld8 + cv), yy);
            if (HAS_Z) unpack8(ldg_stream(z + p * z_ld8 + cv), zz);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                bool dead = HAS_Z && !(zz[i] > 0.f);
                if (MASK == 2) dead = !(fmaf(yy[i], sc[i], sh[i]) > 0.f);
                const float gi = dead ? 0.f : g[i];
                g[i] = gi;
                const float xh = (yy[i] - mu[i]) * is[i];
                o[i] = k0[i] * (gi - k1[i] - xh * k2[i]);
            }
            stg_stream(dy + p * dy_ld8 + cv, pack8(o));
            if (HAS_DRES) stg_stream(dres + p * dres_ld8 + cv, pack8(g));
        }
    }
}

// ------------------------------------------------------------------------------------------------
// one-pass backward, CHANNEL-PARTITIONED over thread-block clusters.
// The BN reductions are per channel, so nothing forces a grid-wide barrier: cluster k owns the channel vectors
// [k Vc, k Vc + Vc) for ALL pixels, CTA r of the cluster the pixel range [r Pc, r Pc + Pc).  Phase 1 streams dz / y (/ z)
// once and leaves per-CTA partial sums in shared memory; ONE hardware cluster barrier; every CTA sums the partials of its
// cluster over distributed shared memory in rank order (deterministic -- no atomics, no zero-filled accumulator);
// phase 2 re-reads the CTA's own pixels (L1 / L2 hits: the CTA touched exactly these lines microseconds ago) and writes
// dy (/ dres).  One launch instead of two, no co-residency requirement across clusters (no cooperative launch, so it runs
// beside the side-stream weight gradients), and with several ranks each cluster runs the SyncBN exchange for ITS channels
// (CTA 0 pushes, every CTA polls the rank's own inbox) -- no exchange launch and no single-block bottleneck.
// scratch: one zero-initialised 64-bit word (clusters that finished the exchange; the last one bumps the sequence).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cl_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cl_nctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cl_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ double cl_ld_f64(const double* local, uint32_t cta) {
    const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(local));
    uint32_t ra;
    double v;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(a), "r"(cta));
    asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(ra) : "memory");
    return v;
}

template <int MASK, bool HAS_DRES>
__global__ void __launch_bounds__(512) bn_bwd_cluster_kernel(const uint4* __restrict__ dz, long long dz_ld8,
                                                             const uint4* __restrict__ y, long long y_ld8,
                                                             const uint4* __restrict__ z, long long z_ld8,
                                                             const float* __restrict__ mean,
                                                             const float* __restrict__ invstd,
                                                             const float* __restrict__ scale,
                                                             const float* __restrict__ shift,
                                                             const float* __restrict__ gamma, double inv_count,
                                                             long long P, int C, int C8, int Vc, uint4* __restrict__ dy,
                                                             long long dy_ld8, uint4* __restrict__ dres,
                                                             long long dres_ld8, float* __restrict__ dgamma,
                                                             float* __restrict__ dbeta,
                                                             const __grid_constant__ SyncArgs sync,
                                                             unsigned long long* __restrict__ scratch) {
    pdl_sync();
    constexpr bool HAS_Z = (MASK == 1);
    __shared__ float wred[16 * 8 * 16];     // [warp][cx][16] warp totals (Vc <= 8, <= 16 warps)
    __shared__ double part[8 * 16];         // [cx][16] this CTA's partial sums (read by the whole cluster)
    __shared__ float kf[8 * 16];            // [cx][16] mean(g) | mean(g * xhat) of this cluster's channels
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int CL = static_cast<int>(cl_nctarank()), r = static_cast<int>(cl_ctarank());
    const int group = blockIdx.x / CL;
    const int cx = tid % Vc, ry = tid / Vc, R = blockDim.x / Vc;     // Vc is a power of two <= 8: lane % Vc == cx
    const int cv = group * Vc + cx;
    const bool act = cv < C8;
    const long long Pc = (P + CL - 1) / CL;
    const long long p0 = r * Pc, p1 = (p0 + Pc < P) ? p0 + Pc : P;
    float mu[8], is[8], sc[8], sh[8];
    if (act) {
        load8f(mean + cv * 8, mu);
        load8f(invstd + cv * 8, is);
        if (MASK == 2) { load8f(scale + cv * 8, sc); load8f(shift + cv * 8, sh); }
    }
    // the exchange's sequence number: read BEFORE the first cluster barrier -- the last cluster bumps the counter once every
    // cluster has finished its exchange, and every CTA of every cluster reads it ahead of its own barrier, so nobody can see
    // the bumped value
    const unsigned long long seq0 = sync.world > 1 ? __ldcg(sync.seq_dev) + 1 : 0ull;
    // ---------------- phase 1: partial sums of g and g * xhat over this CTA's pixels ----------------
    float sg[8], sx[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { sg[i] = 0.f; sx[i] = 0.f; }
    if (act) {
        long long p = p0 + ry;
        for (; p + 3 * R < p1; p += 4 * R) {
            uint4 a[4], b[4], c[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                a[u] = __ldg(dz + (p + u * R) * dz_ld8 + cv);      // (default caching: phase 2 re-reads these lines)
                b[u] = __ldg(y + (p + u * R) * y_ld8 + cv);
                if (HAS_Z) c[u] = __ldg(z + (p + u * R) * z_ld8 + cv);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                float g[8], yy[8], zz[8];
                unpack8(a[u], g);
                unpack8(b[u], yy);
                if (HAS_Z) unpack8(c[u], zz);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    bool dead = HAS_Z && !(zz[i] > 0.f);
                    if (MASK == 2) dead = !(fmaf(yy[i], sc[i], sh[i]) > 0.f);
                    const float gi = dead ? 0.f : g[i];
                    sg[i] += gi;
                    sx[i] = fmaf(gi, (yy[i] - mu[i]) * is[i], sx[i]);
                }
            }
        }
        for (; p < p1; p += R) {
            float g[8], yy[8], zz[8];
            unpack8(__ldg(dz + p * dz_ld8 + cv), g);
            unpack8(__ldg(y + p * y_ld8 + cv), yy);
            if (HAS_Z) unpack8(__ldg(z + p * z_ld8 + cv), zz);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                bool dead = HAS_Z && !(zz[i] > 0.f);
                if (MASK == 2) dead = !(fmaf(yy[i], sc[i], sh[i]) > 0.f);
                const float gi = dead ? 0.f : g[i];
                sg[i] += gi;
                sx[i] = fmaf(gi, (yy[i] - mu[i]) * is[i], sx[i]);
            }
        }
    }
    // warp: lanes with the same cx (lane % Vc) hold partials of the same channels
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        for (int o = 16; o >= Vc; o >>= 1) {
            sg[i] += __shfl_xor_sync(0xFFFFFFFFu, sg[i], o);
            sx[i] += __shfl_xor_sync(0xFFFFFFFFu, sx[i], o);
        }
    }
    if (lane < Vc) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            wred[(warp * 8 + lane) * 16 + i] = sg[i];
            wred[(warp * 8 + lane) * 16 + 8 + i] = sx[i];
        }
    }
    __syncthreads();
    const int nval = Vc * 16;                 // thread t < nval owns value (cx = t / 16, j = t % 16)
    if (tid < nval) {
        double s = 0.0;
        for (int w = 0; w < nwarps; ++w) s += static_cast<double>(wred[(w * 8 + (tid >> 4)) * 16 + (tid & 15)]);
        part[tid] = s;
    }
    cl_sync();                                // partials of every CTA of the cluster are visible
    double tot = 0.0;
    if (tid < nval)
        for (int q = 0; q < CL; ++q) tot += cl_ld_f64(part + tid, static_cast<uint32_t>(q));
    cl_sync();                                // nobody reads a peer's shared memory after this point (CTAs may exit)
    if (tid < nval) {
        const int vcv = group * Vc + (tid >> 4), j = tid & 15;
        const bool vact = vcv < C8;
        const int c = vcv * 8 + (j & 7);
        if (vact && r == 0) {                 // parameter gradients from the LOCAL sums
            if (j < 8) { if (dbeta) dbeta[c] += static_cast<float>(tot); }
            else if (dgamma) dgamma[c] += static_cast<float>(tot);
        }
        if (sync.world > 1) {
            const unsigned long long seq = seq0;
            const unsigned long long tag = (seq & 0xFFFFFFFFull) << 32;
            const int slot = static_cast<int>(seq % kCommSlots);
            const int gi = (j < 8 ? 0 : C) + c;                          // index in the [sum g | sum g*xhat] layout
            if (vact) {
                if (r == 0) {
                    const unsigned long long bb = static_cast<unsigned long long>(__double_as_longlong(tot));
                    const unsigned long long w0 = (bb & 0xFFFFFFFFull) | tag, w1 = (bb >> 32) | tag;
                    for (int q = 0; q < sync.world; ++q) {
                        if (q == sync.rank) continue;
                        unsigned long long* dst = reinterpret_cast<unsigned long long*>(
                            sync.peers.p[q] + (static_cast<size_t>(slot) * sync.world + sync.rank) * kCommSlotDoubles + gi);
                        st_u64_sys(dst, w0);
                        st_u64_sys(dst + 1, w1);
                    }
                }
                const ulonglong2* inbox = sync.peers.p[sync.rank] + static_cast<size_t>(slot) * sync.world * kCommSlotDoubles;
                const unsigned long long t0 = gtimer();
                double s = 0.0;
                for (int q = 0; q < sync.world; ++q) {
                    if (q == sync.rank) { s += tot; continue; }
                    const ulonglong2* src = inbox + static_cast<size_t>(q) * kCommSlotDoubles + gi;
                    ulonglong2 w = ld_v2_sys(src);
                    unsigned int spins = 0;
                    while ((w.x & 0xFFFFFFFF00000000ull) != tag || (w.y & 0xFFFFFFFF00000000ull) != tag) {
                        if ((++spins & 1023u) == 0 && gtimer() - t0 > sync.timeout_ns) {
                            printf("gaiaseg_b200: SyncBN peer exchange (bn_bwd cluster) timed out (rank %d waiting for rank %d, seq %llu)\n",
                                   sync.rank, q, seq);
                            __trap();
                        }
                        w = ld_v2_sys(src);
                    }
                    s += __longlong_as_double(static_cast<long long>((w.x & 0xFFFFFFFFull) | (w.y << 32)));
                }
                tot = s;
            }
        }
        kf[tid] = static_cast<float>(tot * inv_count);
    }
    __syncthreads();
    if (sync.world > 1 && r == 0 && tid == 0) {
        // every cluster read the sequence number before it pushed; the LAST cluster to finish its exchange bumps it
        __threadfence();
        const unsigned long long done = atomicAdd(scratch, 1ull);
        if (done + 1 == static_cast<unsigned long long>(gridDim.x / CL)) *sync.seq_dev = seq0;
    }
    if (!act) return;
    // ---------------- phase 2: dy = gamma*invstd*( g - mean(g) - xhat*mean(g*xhat) ); dres = g ----------------
    float k0[8], k1[8], k2[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float g = gamma ? __ldg(gamma + cv * 8 + i) : 1.f;
        k0[i] = g * is[i];
        k1[i] = kf[cx * 16 + i];
        k2[i] = kf[cx * 16 + 8 + i];
    }
    long long p = p0 + ry;
    for (; p + 3 * R < p1; p += 4 * R) {
        uint4 va[4], vb[4], vc[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            va[u] = __ldg(dz + (p + u * R) * dz_ld8 + cv);
            vb[u] = __ldg(y + (p + u * R) * y_ld8 + cv);
            if (HAS_Z) vc[u] = __ldg(z + (p + u * R) * z_ld8 + cv);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            float g[8], yy[8], zz[8], o[8];
            unpack8(va[u], g);
            unpack8(vb[u], yy);
            if (HAS_Z) unpack8(vc[u], zz);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                bool dead = HAS_Z && !(zz[i] > 0.f);
                if (MASK == 2) dead = !(fmaf(yy[i], sc[i], sh[i]) > 0.f);
                const float gi = dead ? 0.f : g[i];
                g[i] = gi;
                const float xh = (yy[i] - mu[i]) * is[i];
                o[i] = k0[i] * (gi - k1[i] - xh * k2[i]);
            }
            stg_stream(dy + (p + u * R) * dy_ld8 + cv, pack8(o));
            if (HAS_DRES) stg_stream(dres + (p + u * R) * dres_ld8 + cv, pack8(g));
        }
    }
    for (; p < p1; p += R) {
        float g[8], yy[8], zz[8], o[8];
        unpack8(__ldg(dz + p * dz_ld8 + cv), g);
        unpack8(__ldg(y + p * y_ld8 + cv), yy);
        if (HAS_Z) unpack8(__ldg(z + p * z_ld8 + cv), zz);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            bool dead = HAS_Z && !(zz[i] > 0.f);
            if (MASK == 2) dead = !(fmaf(yy[i], sc[i], sh[i]) > 0.f);
            const float gi = dead ? 0.f : g[i];
            g[i] = gi;
            const float xh = (yy[i] - mu[i]) * is[i];
            o[i] = k0[i] * (gi - k1[i] - xh * k2[i]);
        }
        stg_stream(dy + p * dy_ld8 + cv, pack8(o));
        if (HAS_DRES) stg_stream(dres + p * dres_ld8 + cv, pack8(g));
    }
}

// eval-mode / frozen BN backward: dy = scale * g (no statistics terms); dres = g
template <bool HAS_Z, bool HAS_DRES>
__global__ void __launch_bounds__(256) affine_bwd_kernel(const uint4* __restrict__ dz, long long dz_ld8,
                                                         const uint4* __restrict__ z, long long z_ld8,
                                                         const float* __restrict__ scale, long long P, int C8, int Vc,
                                                         int R, uint4* __restrict__ dy, long long dy_ld8,
                                                         uint4* __restrict__ dres, long long dres_ld8) {
    pdl_sync();
    const int cx = threadIdx.x % Vc, ry = threadIdx.x / Vc;
    const long long S = (long long)gridDim.x * R;
    for (int cv = cx; cv < C8; cv += Vc) {
        float sc[8];
        if (scale) load8f(scale + cv * 8, sc);
        else {
#pragma unroll
            for (int i = 0; i < 8; ++i) sc[i] = 1.f;
        }
        for (long long p = (long long)blockIdx.x * R + ry; p < P; p += S) {
            float g[8], zz[8], o[8];
            unpack8(ldg_stream(dz + p * dz_ld8 + cv), g);
            if (HAS_Z) unpack8(ldg_stream(z + p * z_ld8 + cv), zz);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float gi = (HAS_Z && !(zz[i] > 0.f)) ? 0.f : g[i];
                g[i] = gi;
                o[i] = sc[i] * gi;
            }
            stg_stream(dy + p * dy_ld8 + cv, pack8(o));
            if (HAS_DRES) stg_stream(dres + p * dres_ld8 + cv, pack8(g));
        }
    }
}

__global__ void bn_bwd_param_kernel(const double* __restrict__ sums, int C, float* __restrict__ dgamma,
                                    float* __restrict__ dbeta, int accumulate) {
    pdl_sync();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const float dg = static_cast<float>(sums[C + c]);
    const float db = static_cast<float>(sums[c]);
    if (dgamma) dgamma[c] = accumulate ? dgamma[c] + dg : dg;
    if (dbeta) dbeta[c] = accumulate ? dbeta[c] + db : db;
}

static int check_act(const void* p, long long ld, int C, const char* what) {
    GS_REQUIRE(p != nullptr, "%s: null pointer", what);
    GS_REQUIRE(C > 0 && C % 8 == 0, "%s: channels (%d) must be a positive multiple of 8", what, C);
    GS_REQUIRE(ld >= C && ld % 8 == 0, "%s: pitch (%lld) must be >= C and a multiple of 8", what, ld);
    GS_REQUIRE((reinterpret_cast<uintptr_t>(p) & 15) == 0, "%s: pointer must be 16-byte aligned", what);
    return 0;
}

}  // namespace gs

using namespace gs;

// Grid policy.  Every block of a REDUCING kernel ends with 2*C same-address fp64 atomics, and the L2 atomic units
// serialise those (~8-17 ns per op and address, measured): the number of blocks is the dominant cost for these
// 10-20 us kernels, so reductions use at most 2 blocks per SM with deep unrolling (>= 64 KB of loads in flight per SM).
// Streaming (apply) kernels use up to 3 blocks per SM and at least 12 pixels per thread row so the per-channel
// prologue is amortised.
static inline int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return v ? atoi(v) : dflt;
}
static inline int reduce_grid(const ColMap& m, long long P) {
    static const int bps = env_int("GS_BN_REDUCE_BLOCKS_PER_SM", 2), ppt = env_int("GS_BN_REDUCE_PPT", 8);
    const int per_sm = m.threads <= 192 ? bps + 1 : bps;   // narrow blocks (C8 = 160, 320): keep >= 480 threads per SM
    return colmap_grid(m, P, ppt, 148 * per_sm);
}
// gs_sync_desc (C ABI) -> SyncArgs (kernel argument); NULL / world <= 1 -> no exchange
static int make_sync(const gs_sync_desc* d, SyncArgs* out, const char* what) {
    *out = SyncArgs{};
    out->world = 1;
    if (d == nullptr || d->world <= 1) return 0;
    GS_REQUIRE(d->world <= kCommMaxWorld && d->rank >= 0 && d->rank < d->world, "%s: bad rank %d / world %d", what, d->rank,
               d->world);
    GS_REQUIRE(d->peer_inboxes != nullptr && d->seq_dev != nullptr, "%s: sync descriptor without inboxes / sequence counter",
               what);
    for (int r = 0; r < d->world; ++r) {
        GS_REQUIRE(d->peer_inboxes[r] != nullptr, "%s: inbox of rank %d is not mapped", what, r);
        out->peers.p[r] = reinterpret_cast<ulonglong2*>(const_cast<void*>(d->peer_inboxes[r]));
    }
    out->rank = d->rank;
    out->world = d->world;
    out->seq_dev = reinterpret_cast<unsigned long long*>(d->seq_dev);
    out->timeout_ns = comm_timeout_ns();
    out->phase = d->phase;
    return 0;
}

static inline int stream_grid(const ColMap& m, long long P) {
    static const int bps = env_int("GS_BN_STREAM_BLOCKS_PER_SM", 3), ppt = env_int("GS_BN_STREAM_PPT", 12);
    return colmap_grid(m, P, ppt, 148 * bps);
}

extern "C" int gs_bn_stats(const void* x, int64_t P, int32_t C, int32_t ld, double* stats, void* stream) {
    if (check_act(x, ld, C, "bn_stats x")) return -1;
    GS_REQUIRE(stats != nullptr, "bn_stats: null stats");
    if (P <= 0) return 0;
    const ColMap m = make_colmap(C);
    const int grid = reduce_grid(m, P);
    gs::launch(bn_stats_kernel, dim3(grid), dim3(m.threads), 0, static_cast<cudaStream_t>(stream), 
        reinterpret_cast<const uint4*>(x), ld / 8, P, C, m.C8, m.Vc, m.R, stats);
    GS_LAUNCHED();
    return 0;
}

extern "C" int gs_bn_finalize(const double* stats, double count, int32_t C, const float* gamma, const float* beta,
                              float* running_mean, float* running_var, float momentum, float eps, float* mean,
                              float* invstd, float* scale, float* shift, void* stream) {
    GS_REQUIRE(stats && scale && shift, "bn_finalize: null pointer");
    GS_REQUIRE(C > 0 && count > 0, "bn_finalize: empty reduction (C=%d count=%f)", C, count);
    gs::launch(bn_finalize_kernel, dim3((C + 127) / 128), dim3(128), 0, static_cast<cudaStream_t>(stream), 
        stats, count, C, gamma, beta, running_mean, running_var, momentum, eps, mean, invstd, scale, shift);
    GS_LAUNCHED();
    return 0;
}

extern "C" int gs_bn_eval_affine(int32_t C, const float* gamma, const float* beta, const float* running_mean,
                                 const float* running_var, float eps, float* scale, float* shift, void* stream) {
    GS_REQUIRE(running_mean && running_var && scale && shift, "bn_eval_affine: null pointer");
    GS_REQUIRE(C > 0, "bn_eval_affine: C=%d", C);
    gs::launch(bn_eval_affine_kernel, dim3((C + 127) / 128), dim3(128), 0, static_cast<cudaStream_t>(stream), 
        C, gamma, beta, running_mean, running_var, eps, scale, shift);
    GS_LAUNCHED();
    return 0;
}

extern "C" int gs_bn_apply(const void* y, int32_t y_ld, const float* scale, const float* shift, const void* residual,
                           int32_t res_ld, int32_t relu, void* z, int32_t z_ld, int64_t P, int32_t C, void* stream) {
    if (check_act(y, y_ld, C, "bn_apply y") || check_act(z, z_ld, C, "bn_apply z")) return -1;
    if (residual && check_act(residual, res_ld, C, "bn_apply residual")) return -1;
    if (P <= 0) return 0;
    const ColMap m = make_colmap(C);
    const int grid = stream_grid(m, P);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    BnTrainArgs t{};
    if (residual)
        gs::launch(bn_apply_kernel<true, false>, dim3(grid), dim3(m.threads), 0, st, 
            reinterpret_cast<const uint4*>(y), y_ld / 8, scale, shift, reinterpret_cast<const uint4*>(residual),
            res_ld / 8, relu, reinterpret_cast<uint4*>(z), z_ld / 8, P, m.C8, m.Vc, m.R, t);
    else
        gs::launch(bn_apply_kernel<false, false>, dim3(grid), dim3(m.threads), 0, st, 
            reinterpret_cast<const uint4*>(y), y_ld / 8, scale, shift, nullptr, 0, relu, reinterpret_cast<uint4*>(z),
            z_ld / 8, P, m.C8, m.Vc, m.R, t);
    GS_LAUNCHED();
    return 0;
}

extern "C" int gs_bn_apply_train(const void* y, int32_t y_ld, const double* stats, double count, const float* gamma,
                                 const float* beta, float* running_mean, float* running_var, float momentum, float eps,
                                 float* aff, const void* residual, int32_t res_ld, int32_t relu, void* z, int32_t z_ld,
                                 int64_t P, int32_t C, const gs_sync_desc* sync, void* stream) {
    if (check_act(y, y_ld, C, "bn_apply_train y") || check_act(z, z_ld, C, "bn_apply_train z")) return -1;
    if (residual && check_act(residual, res_ld, C, "bn_apply_train residual")) return -1;
    GS_REQUIRE(stats && aff && count > 0, "bn_apply_train: null stats / aff or empty count");
    if (P <= 0) return 0;
    const ColMap m = make_colmap(C);
    const int grid = stream_grid(m, P);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    BnTrainArgs t{};
    t.stats = stats; t.inv_count = 1.0 / count; t.unbias = count > 1.0 ? count / (count - 1.0) : 1.0;
    t.gamma = gamma; t.beta = beta; t.rm = running_mean; t.rv = running_var; t.momentum = momentum; t.eps = eps;
    t.aff = aff; t.C = C;
    if (make_sync(sync, &t.sync, "bn_apply_train")) return -1;
    if (t.sync.world > 1) {
        GS_REQUIRE(C <= kCommSlotDoubles / 2, "bn_apply_train: %d channels exceed the exchange slot", C);
        // the caller's stats buffer carries two zero-initialised scratch words behind the 2C sums
        t.flag = reinterpret_cast<unsigned long long*>(const_cast<double*>(stats) + 2 * C) + 1;
    }
    if (residual)
        gs::launch<16>(bn_apply_kernel<true, true>, dim3(grid), dim3(m.threads), 0, st, 
            reinterpret_cast<const uint4*>(y), y_ld / 8, nullptr, nullptr, reinterpret_cast<const uint4*>(residual),
            res_ld / 8, relu, reinterpret_cast<uint4*>(z), z_ld / 8, P, m.C8, m.Vc, m.R, t);
    else
        gs::launch<16>(bn_apply_kernel<false, true>, dim3(grid), dim3(m.threads), 0, st, 
            reinterpret_cast<const uint4*>(y), y_ld / 8, nullptr, nullptr, nullptr, 0, relu,
            reinterpret_cast<uint4*>(z), z_ld / 8, P, m.C8, m.Vc, m.R, t);
    GS_LAUNCHED();
    return 0;
}

extern "C" int gs_bn_bwd_reduce(const void* dz, int32_t dz_ld, const void* y, int32_t y_ld, const void* z,
                                int32_t z_ld, const float* mean, const float* invstd, const float* scale,
                                const float* shift, int32_t relu, int64_t P, int32_t C, double* sums,
                                const gs_sync_desc* sync, void* stream) {
    if (check_act(dz, dz_ld, C, "bn_bwd_reduce dz") || check_act(y, y_ld, C, "bn_bwd_reduce y")) return -1;
    if (z && check_act(z, z_ld, C, "bn_bwd_reduce z")) return -1;
    GS_REQUIRE(mean && invstd && sums, "bn_bwd_reduce: null pointer");
    GS_REQUIRE(!relu || z || (scale && shift), "bn_bwd_reduce: ReLU mask needs z or (scale, shift)");
    if (P <= 0) return 0;
    const ColMap m = make_colmap(C);
    const int grid = reduce_grid(m, P);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int mask = !relu ? 0 : (z ? 1 : 2);
    SyncArgs sa;
    if (make_sync(sync, &sa, "bn_bwd_reduce")) return -1;
    GS_REQUIRE(sa.world <= 1 || (sa.phase == 1 && C <= kCommSlotDoubles / 2),
               "bn_bwd_reduce: sync descriptor must have phase == 1 and C <= %d", kCommSlotDoubles / 2);
    unsigned long long* ticket = reinterpret_cast<unsigned long long*>(sums + 2 * C);   // first scratch word behind the sums
#define GS_BWD_REDUCE(MK)                                                                                            \
    gs::launch<32>(sa.world > 1 ? bn_bwd_reduce_kernel<MK, true> : bn_bwd_reduce_kernel<MK, false>, dim3(grid), dim3(m.threads), 0, st, \
        reinterpret_cast<const uint4*>(dz), dz_ld / 8, reinterpret_cast<const uint4*>(y), y_ld / 8,                  \
        reinterpret_cast<const uint4*>(z), z_ld / 8, mean, invstd, scale, shift, P, C, m.C8, m.Vc, m.R, sums, sa, ticket)
    if (mask == 0) GS_BWD_REDUCE(0);
    else if (mask == 1) GS_BWD_REDUCE(1);
    else GS_BWD_REDUCE(2);
#undef GS_BWD_REDUCE
    GS_LAUNCHED();
    return 0;
}

extern "C" int gs_bn_bwd_apply(const void* dz, int32_t dz_ld, const void* y, int32_t y_ld, const void* z, int32_t z_ld,
                               const float* mean, const float* invstd, const float* scale, const float* shift,
                               int32_t relu, const float* gamma, const double* sums, double count, int64_t P, int32_t C,
                               void* dy, int32_t dy_ld, void* dres, int32_t dres_ld, float* dgamma, float* dbeta,
                               const gs_sync_desc* sync, void* stream) {
    if (check_act(dz, dz_ld, C, "bn_bwd_apply dz") || check_act(y, y_ld, C, "bn_bwd_apply y") ||
        check_act(dy, dy_ld, C, "bn_bwd_apply dy"))
        return -1;
    if (z && check_act(z, z_ld, C, "bn_bwd_apply z")) return -1;
    if (dres && check_act(dres, dres_ld, C, "bn_bwd_apply dres")) return -1;
    GS_REQUIRE(mean && invstd && sums && count > 0, "bn_bwd_apply: null pointer / empty count");
    GS_REQUIRE(!relu || z || (scale && shift), "bn_bwd_apply: ReLU mask needs z or (scale, shift)");
    if (P <= 0) return 0;
    const ColMap m = make_colmap(C);
    const int grid = stream_grid(m, P);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const double inv_count = 1.0 / count;
    const int mask = !relu ? 0 : (z ? 1 : 2);
    SyncArgs sa;
    if (make_sync(sync, &sa, "bn_bwd_apply")) return -1;
    GS_REQUIRE(sa.world <= 1 || C <= kCommSlotDoubles / 2, "bn_bwd_apply: %d channels exceed the exchange slot", C);
    // second scratch word behind the sums: "group sums final" flag (zero on entry)
    unsigned long long* flag = reinterpret_cast<unsigned long long*>(const_cast<double*>(sums) + 2 * C) + 1;
#define GS_BWD_APPLY(MK, HD)                                                                                         \
    gs::launch<64>(sa.world > 1 ? bn_bwd_apply_kernel<MK, HD, true> : bn_bwd_apply_kernel<MK, HD, false>, dim3(grid), dim3(m.threads), 0, st, \
        reinterpret_cast<const uint4*>(dz), dz_ld / 8, reinterpret_cast<const uint4*>(y), y_ld / 8,                  \
        reinterpret_cast<const uint4*>(z), z_ld / 8, mean, invstd, scale, shift, gamma, sums, inv_count, P, C, m.C8, \
        m.Vc, m.R, reinterpret_cast<uint4*>(dy), dy_ld / 8, reinterpret_cast<uint4*>(dres), dres_ld / 8, dgamma, dbeta, \
        sa, flag)
    if (dres) {
        if (mask == 0) GS_BWD_APPLY(0, true);
        else if (mask == 1) GS_BWD_APPLY(1, true);
        else GS_BWD_APPLY(2, true);
    } else {
        if (mask == 0) GS_BWD_APPLY(0, false);
        else if (mask == 1) GS_BWD_APPLY(1, false);
        else GS_BWD_APPLY(2, false);
    }
#undef GS_BWD_APPLY
    GS_LAUNCHED();
    return 0;
}

extern "C" int gs_bn_bwd(const void* dz, int32_t dz_ld, const void* y, int32_t y_ld, const void* z, int32_t z_ld,
                         const float* mean, const float* invstd, const float* scale, const float* shift, int32_t relu,
                         const float* gamma, double* sums, double count, int64_t P, int32_t C, void* dy, int32_t dy_ld,
                         void* dres, int32_t dres_ld, float* dgamma, float* dbeta, const gs_sync_desc* sync,
                         void* stream) {
    if (check_act(dz, dz_ld, C, "bn_bwd dz") || check_act(y, y_ld, C, "bn_bwd y") || check_act(dy, dy_ld, C, "bn_bwd dy"))
        return -1;
    if (z && check_act(z, z_ld, C, "bn_bwd z")) return -1;
    if (dres && check_act(dres, dres_ld, C, "bn_bwd dres")) return -1;
    GS_REQUIRE(mean && invstd && sums && count > 0, "bn_bwd: null pointer / empty count");
    GS_REQUIRE(!relu || z || (scale && shift), "bn_bwd: ReLU mask needs z or (scale, shift)");
    if (P <= 0) return 0;
    SyncArgs sa;
    if (make_sync(sync, &sa, "bn_bwd")) return -1;
    GS_REQUIRE(sa.world <= 1 || C <= kCommSlotDoubles / 2, "bn_bwd: %d channels exceed the exchange slot", C);
    const ColMap m = make_colmap(C);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int mask = !relu ? 0 : (z ? 1 : 2);
    static const int kind = env_int("GS_BN_FUSED_KIND", 1);   // 1: channel-partitioned clusters (default), 0: grid barrier
    if (kind == 1) {
        // geometry: Vc channel vectors per cluster (16 Vc contiguous bytes per pixel), G = ceil(C8 / Vc) clusters of CL CTAs
        // (CL pixel ranges).  Widest Vc <= 4 that still gives ~one wave of CTAs with the largest cluster; then the largest
        // power-of-two CL with G * CL <= target.
        static const int cl_max = env_int("GS_BN_CL_MAX", 16), target = env_int("GS_BN_CL_TARGET", 296);
        static const int vc_max = env_int("GS_BN_CL_VC", 4), threads_env = env_int("GS_BN_CL_THREADS", 256);
        int Vc = vc_max;
        while (Vc > 1 && ((m.C8 + Vc - 1) / Vc) * cl_max < 128) Vc >>= 1;
        const int G = (m.C8 + Vc - 1) / Vc;
        int CL = 1;
        while (CL * 2 <= cl_max && G * CL * 2 <= target) CL *= 2;
        int threads = threads_env;
        while (threads > 64 && (long long)(threads / Vc) * CL * 4 > P) threads >>= 1;   // tiny maps: keep every row busy
        const double inv_count = 1.0 / count;
        unsigned long long* scratch = reinterpret_cast<unsigned long long*>(sums + 2 * C);   // zeroed word behind the sums
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(G * CL);
        cfg.blockDim = dim3(threads);
        cfg.dynamicSmemBytes = 0;
        cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
#define GS_BWD_CL(MK, HD)                                                                                               \
    do {                                                                                                                \
        static bool np = false;                                                                                         \
        if (!np) {                                                                                                      \
            GS_CUDA_OK(cudaFuncSetAttribute(bn_bwd_cluster_kernel<MK, HD>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1)); \
            np = true;                                                                                                  \
        }                                                                                                               \
        GS_CUDA_OK(cudaLaunchKernelEx(&cfg, bn_bwd_cluster_kernel<MK, HD>, reinterpret_cast<const uint4*>(dz),          \
                                      (long long)(dz_ld / 8), reinterpret_cast<const uint4*>(y), (long long)(y_ld / 8), \
                                      reinterpret_cast<const uint4*>(z), (long long)(z_ld / 8), mean, invstd, scale,    \
                                      shift, gamma, inv_count, (long long)P, (int)C, m.C8, Vc,                         \
                                      reinterpret_cast<uint4*>(dy), (long long)(dy_ld / 8),                            \
                                      reinterpret_cast<uint4*>(dres), (long long)(dres_ld / 8), dgamma, dbeta, sa,     \
                                      scratch));                                                                        \
    } while (0)
        if (dres) {
            if (mask == 0) GS_BWD_CL(0, true);
            else if (mask == 1) GS_BWD_CL(1, true);
            else GS_BWD_CL(2, true);
        } else {
            if (mask == 0) GS_BWD_CL(0, false);
            else if (mask == 1) GS_BWD_CL(1, false);
            else GS_BWD_CL(2, false);
        }
#undef GS_BWD_CL
        GS_LAUNCHED();
        return 0;
    }
    // the blocks wait on one another (grid barrier): COOPERATIVE launch, grid <= what is co-resident by construction
    static const int bps = env_int("GS_BN_FUSED_BLOCKS_PER_SM", 2), ppt = env_int("GS_BN_FUSED_PPT", 8);
    const int grid = colmap_grid(m, P, ppt, num_sms() * bps);
    unsigned long long* scratch = reinterpret_cast<unsigned long long*>(sums + 2 * C);   // two zeroed words behind the sums
    const double inv_count = 1.0 / count;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(m.threads);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeCooperative;
    at[0].val.cooperative = 1;
    cfg.attrs = at;
    static const int coop = env_int("GS_BN_FUSED_COOP", 1);
    cfg.numAttrs = coop ? 1 : 0;
#define GS_BWD_FUSED(MK, HD)                                                                                           \
    GS_CUDA_OK(cudaLaunchKernelEx(&cfg, bn_bwd_fused_kernel<MK, HD>, reinterpret_cast<const uint4*>(dz),              \
                                  (long long)(dz_ld / 8), reinterpret_cast<const uint4*>(y), (long long)(y_ld / 8),   \
                                  reinterpret_cast<const uint4*>(z), (long long)(z_ld / 8), mean, invstd, scale, shift, \
                                  gamma, sums, inv_count, (long long)P, (int)C, m.C8, m.Vc, m.R,                      \
                                  reinterpret_cast<uint4*>(dy), (long long)(dy_ld / 8), reinterpret_cast<uint4*>(dres), \
                                  (long long)(dres_ld / 8), dgamma, dbeta, sa, scratch))
    if (dres) {
        if (mask == 0) GS_BWD_FUSED(0, true);
        else if (mask == 1) GS_BWD_FUSED(1, true);
        else GS_BWD_FUSED(2, true);
    } else {
        if (mask == 0) GS_BWD_FUSED(0, false);
        else if (mask == 1) GS_BWD_FUSED(1, false);
        else GS_BWD_FUSED(2, false);
    }
#undef GS_BWD_FUSED
    GS_LAUNCHED();
    return 0;
}

extern "C" int gs_affine_bwd(const void* dz, int32_t dz_ld, const void* z, int32_t z_ld, const float* scale, int64_t P,
                             int32_t C, void* dy, int32_t dy_ld, void* dres, int32_t dres_ld, void* stream) {
    if (check_act(dz, dz_ld, C, "affine_bwd dz") || check_act(dy, dy_ld, C, "affine_bwd dy")) return -1;
    if (z && check_act(z, z_ld, C, "affine_bwd z")) return -1;
    if (dres && check_act(dres, dres_ld, C, "affine_bwd dres")) return -1;
    if (P <= 0) return 0;
    const ColMap m = make_colmap(C);
    const int grid = stream_grid(m, P);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
#define GS_AFF_BWD(HZ, HD)                                                                                      \
    gs::launch(affine_bwd_kernel<HZ, HD>, dim3(grid), dim3(m.threads), 0, st,                                                       \
        reinterpret_cast<const uint4*>(dz), dz_ld / 8, reinterpret_cast<const uint4*>(z), z_ld / 8, scale, P,   \
        m.C8, m.Vc, m.R, reinterpret_cast<uint4*>(dy), dy_ld / 8, reinterpret_cast<uint4*>(dres), dres_ld / 8)
    if (z && dres) GS_AFF_BWD(true, true);
    else if (z) GS_AFF_BWD(true, false);
    else if (dres) GS_AFF_BWD(false, true);
    else GS_AFF_BWD(false, false);
#undef GS_AFF_BWD
    GS_LAUNCHED();
    return 0;
}

extern "C" int gs_bn_bwd_param(const double* sums_local, int32_t C, float* dgamma, float* dbeta, int32_t accumulate,
                               void* stream) {
    GS_REQUIRE(sums_local != nullptr && C > 0, "bn_bwd_param: null pointer");
    gs::launch(bn_bwd_param_kernel, dim3((C + 127) / 128), dim3(128), 0, static_cast<cudaStream_t>(stream), sums_local, C, dgamma, dbeta,
                                                                                      accumulate);
    GS_LAUNCHED();
    return 0;
}
