// gs_conv_tc.cu -- elastic-width convolutions as tcgen05 / TMEM implicit GEMMs (sm_100a).
//
// Two kernels:
//   igemm_kernel : forward conv and data-gradient.  Persistent (one CTA per SM), warp-specialised:
//       warp 0   TMA producer   -- per (tap r,s ; 64-channel K chunk) one 4-D box of the NHWC
//                                  activation (shifted by the tap, OOB = zero padding) and one
//                                  box of the max-width weight; the tensor-map extents are the
//                                  ACTIVE prefix slice, the strides the MAX widths => the slice
//                                  is addressed in place, no copy.
//       warp 1   MMA issuer     -- tcgen05.mma kind::f16 (bf16 x bf16 -> fp32), M = 128 pixels,
//                                  N = up to 256 output channels, accumulators double-buffered
//                                  in TMEM (2 x 256 columns) so the epilogue of tile i overlaps
//                                  the main loop of tile i+1.
//       warps 2-17 epilogue     -- 4 sub-tile groups x 4 TMEM sub-partitions: tcgen05.ld -> affine /
//                                  residual / ReLU -> bf16 -> swizzled smem staging -> TMA store;
//                                  per-channel sum / sum^2 (DynBN batch statistics) from the staged
//                                  tile, kept in registers across the CTA's tiles -> fp64 atomics.
//       Instantiations: <1> one CTA per tile (short K loops), <2, false> CTA pairs (cta_group::2,
//       M = 256, half of the weight tile per CTA), <2, true> pairs with one 320 / 384-column
//       accumulator (wide n-tiles).  Launched with programmatic dependent launch (gs_host.h).
//   wgrad_kernel : weight gradient, stream-K over (item, 64-pixel chunk), both operands MN-major
//       (pixels are K), fp32 red.global.add into the active prefix of the flat gradient.
//
// GEMM view (forward):  Y[pix, co] = sum_{r,s,ci} X[pix shifted by (r,s), ci] * W[co, r, s, ci]
#include <cuda_bf16.h>
#include <stdlib.h>

#include "../../include/gaiaseg_b200.h"
#include "gs_host.h"
#include "gs_ptx.cuh"
#include "gs_comm.cuh"

namespace gs {

// ------------------------------------------------------------------------------------------------
// igemm (fwd / dgrad)
// ------------------------------------------------------------------------------------------------
constexpr int kIgABytes = 128 * 64 * 2;       // 128 pixels x 64 channels bf16
constexpr int kIgStagingBytes = 128 * 256 * 2;  // epilogue tile, 4 sub-tiles of [128][64] bf16
constexpr int kIgBarBytes = 256;
constexpr int kIgThreads = 576;                // TMA warp, MMA warp, 16 epilogue warps
// CG = 1: one CTA per 128-pixel x 256-channel tile.  CG = 2: a CTA PAIR (cluster of two SMs of one TPC, tcgen05
// cta_group::2) computes 256 pixels x 256 channels -- each CTA stages its own 128 pixels of A and only HALF of the weight
// tile (the tensor cores of both SMs read both halves), i.e. 32 KB instead of 48 KB of shared-memory fill per K chunk,
// which is what bounds the single-CTA main loop (profiles/r01_igemm_ncu_full.md).
template <int CG>
struct IgCfg {
    static constexpr int kStages = CG == 1 ? 3 : 5;
    static constexpr int kBBytes = (256 / CG) * 64 * 2;   // (half of) up to 256 out-channels x 64 channels
    static constexpr int kStageBytes = kIgABytes + kBBytes;
    static constexpr int kSmemBytes = 1024 + kStages * kStageBytes + kIgStagingBytes + kIgBarBytes;
};
// the epilogue hands an accumulator buffer back to the MMA thread (pair: the leader's barrier)
template <int CG>
__device__ __forceinline__ void ig_release_acc(uint64_t* bar) {
    if (CG == 2) mbar_arrive_cluster(bar, 0);
    else mbar_arrive(bar);
}

// Optional in-kernel trace (gs_debug_set_trace): CTA 0 writes %globaltimer stamps of its pipeline events.
//   [0] kernel start  [1] setup done  [16+i] producer issued stage i  [80+i] MMA saw stage i full
//   [144+4t+{0,1,2,3}] epilogue tile t: accumulator ready / TMEM drained + staged / store issued / tile done
__device__ unsigned long long* g_trace = nullptr;
__device__ __forceinline__ void trace(int slot) {
    if (g_trace != nullptr && blockIdx.x == 0) g_trace[slot] = global_timer_ns();
}

struct IgemmParams {
    int N, Ho, Wo;        // output pixels
    int TH, TW;           // spatial tile, TH*TW = 128
    int tiles_h, tiles_w;
    int m_tiles, n_tiles;
    int Cout;             // active GEMM-N (output channels)
    int Kc;               // active reduction channels per tap
    int kchunks;          // ceil(Kc / 64)
    int taps_h, taps_w;
    int in_mul, base, step;  // input coord of tap t for output o:  o*in_mul + base + t*step
    int b_box_rows;       // rows of the weight box (<= 256)            (K-major B, forward)
    int b_mn;             // 1: B is MN-major (dgrad reads W[co][r][s][ci] in place: K = co rows, N = ci contiguous)
    // n-tile geometry.  Normal: nt_w = 256, two accumulator buffers of 256 TMEM columns (the epilogue of tile i overlaps the
    // main loop of tile i+1).  WIDE (CTA pairs, long K loops, Cout a multiple of 320 / 384): nt_w = 320 or 384 in ONE
    // 384-column accumulator -- Cout = 320 is one balanced tile instead of 256 + 64 (the 64-column units cost the same
    // per K chunk: the main loop is bound by operand delivery, not by the MMA), the A tile is read once instead of twice.
    // Two MMAs per K step (N = 256, then N = nt_w - 256); each CTA of the pair stages 128 + h2 weight rows.
    int nt_w, wide, h2;
    int even;             // wide + K-major B: two MMAs of N = nt_w / 2 each (CTA r stages rows [hp r, hp r + hp) and
                          // [2 hp + hp r, ...), hp = nt_w / 4) instead of N = 256 followed by a small N = nt_w - 256 one
    int stages, stage_bytes;
    // epilogue
    int direct;           // 1: per-thread global stores (fp32 out or unaligned), 0: TMA store
    int out_f32;
    int relu;
    void* out;
    long long out_ld;
    const float* scale;
    const float* shift;
    const __nv_bfloat16* residual;
    long long res_ld;
    // non-NULL: per-channel sum / sum^2 of the stored (bf16-rounded) output, fp64 [2 * Cout] (DynBN forward statistics)
    double* stats;
    // SyncBN over several ranks (sync.world > 1, needs stats): the LAST CTA to flush its statistics pushes the rank's final
    // sums to every peer inbox over NVLink -- conv and the send half of the statistic all-reduce in one kernel; the BN
    // apply kernel polls.  sync_ticket: zero-initialised 64-bit word counting the CTAs that have flushed.
    SyncArgs sync;
    unsigned long long* sync_ticket;
    // Fused DynBN apply (training forward, gs_conv2d_fwd_bn): after its last tile every CTA flushes its statistics, the
    // persistent grid meets at ONE barrier (all CTAs are co-resident by construction: grid <= SM count, one CTA per SM),
    // and each CTA normalises the tiles IT has just written -- y comes back from L2, z = relu?(y * scale + shift (+ res))
    // is written once -- so the training forward of a conv + BN layer is one launch instead of two.
    struct BnTail {
        int on;
        int relu;
        double inv_count, unbias;
        const float* gamma;
        const float* beta;
        float* rm;
        float* rv;
        float momentum, eps;
        float* aff;                       // [4][C]: mean, invstd, scale, shift (for the backward pass)
        const __nv_bfloat16* res;         // residual added AFTER the normalisation (bn3 of a bottleneck), may be NULL
        long long res_ld;
        __nv_bfloat16* z;
        long long z_ld;
        unsigned long long* barrier;      // zero-initialised arrival counter (first scratch word behind the sums)
        unsigned long long timeout_ns;
        int tw_shift;                     // log2(TW)
    } bn;
};

__device__ __forceinline__ uint4 ld_cg_v4(const void* p) {
    uint4 r;
    asm volatile("ld.global.cg.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ uint4 ld_nc_v4(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_na_v4(void* p, const uint4& v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// The tail of gs_conv2d_fwd_bn, run by the 512 epilogue threads (e = 0..511) after the statistic flush.  Out of line: it
// runs once per kernel and must not cost the main loop any registers.
// Thread map: the n-tile has V = n_valid / 8 channel vectors; thread (cx = e % V, ry = e / V) owns vector cx (its scale /
// shift stay in registers) and the pixel rows ry, ry + R, ... of every tile of the CTA -- a warp touches whole pixel rows
// (V * 16 contiguous bytes).  The (tile, row) slots of a thread are walked as ONE stream through a rolling window of W
// register buffers: a slot is refilled as soon as it has been consumed, so W 16-byte loads per tensor stay in flight per
// thread across tile boundaries (with one CTA per SM nothing else hides the L2 latency), and the first window is issued
// BEFORE the grid barrier -- the CTA's own tiles are complete by then.
template <int CG, bool HAS_RES>
__device__ __noinline__ void ig_bn_tail(const IgemmParams& p, int mu0, int groups, int m_units, int rank, int nt,
                                        uint8_t* staging) {
    constexpr int W = HAS_RES ? 4 : 8;
    const int e = static_cast<int>(threadIdx.x) - 64;
    const int n0 = nt * p.nt_w;
    int n_valid = p.Cout - n0;
    if (n_valid > p.nt_w) n_valid = p.nt_w;
    const int V = n_valid >> 3;
    const int R = 512 / V;
    const int cx = e % V, ry = e / V;
    const int c0 = n0 + cx * 8;
    const int rpt = (128 + R - 1) / R;                                   // row slots per tile and thread
    const int my_tiles = mu0 < m_units ? (m_units - mu0 + groups - 1) / groups : 0;
    const int S = ry < R ? my_tiles * rpt : 0;                           // slots of this thread
    const int tiles_hw = p.tiles_h * p.tiles_w;
    const int tw_mask = p.TW - 1;
    const __nv_bfloat16* const ybase = reinterpret_cast<const __nv_bfloat16*>(p.out) + c0;
    const __nv_bfloat16* const rbase = p.bn.res + c0;
    __nv_bfloat16* const zbase = p.bn.z + c0;
    // slot cursor (slots are visited strictly in order)
    int it_t = 0, it_j = 0, it_img = p.N, it_h0 = 0, it_w0 = 0;
    auto setup_tile = [&](int t) {
        const int mu = mu0 + t * groups;
        it_img = p.N;
        if (mu >= m_units) return;
        const int mt = mu * CG + rank;
        const int img = mt / tiles_hw;
        const int rem = mt - img * tiles_hw;
        it_img = img;                                                    // (phantom tile of a pair: img == N -> no pixel)
        it_h0 = (rem / p.tiles_w) * p.TH;
        it_w0 = (rem % p.tiles_w) * p.TW;
    };
    auto next_pix = [&]() -> int {
        int pix = -1;
        const int row = ry + it_j * R;
        const int hh = it_h0 + (row >> p.bn.tw_shift), ww = it_w0 + (row & tw_mask);
        if (it_img < p.N && row < 128 && hh < p.Ho && ww < p.Wo) pix = (it_img * p.Ho + hh) * p.Wo + ww;
        if (++it_j == rpt) { it_j = 0; setup_tile(++it_t); }
        return pix;
    };
    uint4 buf[W], rbuf[W];
    int pixb[W];
    if (e == 0) trace(208);
    setup_tile(0);
#pragma unroll
    for (int u = 0; u < W; ++u) {
        pixb[u] = u < S ? next_pix() : -1;
        if (pixb[u] >= 0) {
            buf[u] = ld_cg_v4(ybase + static_cast<long long>(pixb[u]) * p.out_ld);
            if (HAS_RES) rbuf[u] = ld_nc_v4(rbase + static_cast<long long>(pixb[u]) * p.bn.res_ld);
        }
    }
    // ---- grid barrier: every CTA's atomics on `stats` are globally visible afterwards
    if (e == 0) trace(209);
    __threadfence();
    named_bar_sync(5, 512);
    if (e == 0) {
        trace(210);
        __threadfence();
        atomicAdd(p.bn.barrier, 1ull);
        const unsigned long long want = gridDim.x;
        const unsigned long long t0 = gtimer();
        unsigned int spins = 0;
        while (ld_acquire_gpu(p.bn.barrier) < want) {
            if ((++spins & 1023u) == 0 && gtimer() - t0 > p.bn.timeout_ns) {
                printf("gaiaseg_b200: conv + BN grid barrier timed out (CTA %d of %d)\n", static_cast<int>(blockIdx.x),
                       static_cast<int>(gridDim.x));
                __trap();
            }
        }
        __threadfence();
        trace(211);
    }
    named_bar_sync(5, 512);
    // ---- finalize: ONE thread per channel of the n-tile turns the sums into scale / shift (fp64 only where cancellation
    // can occur) and leaves them in the idle staging buffer; the CTA that owns the first pixel tile also stores them for
    // the backward pass and updates the running statistics.  (Every thread reading the sums of its own 8 channels made
    // 148 x 512 x 16 L2 reads of the same few lines: 13 us.)
    float* const sm = reinterpret_cast<float*>(staging);       // [0, 512): scale, [512, 1024): shift
    if (e < n_valid) {
        const int c = n0 + e;
        const double s1 = __ldcg(p.stats + c);
        const double s2 = __ldcg(p.stats + p.Cout + c);
        const double m = s1 * p.bn.inv_count;
        double var = s2 * p.bn.inv_count - m * m;
        if (var < 0.0) var = 0.0;
        const float mf = static_cast<float>(m);
        const float istd = rsqrtf(static_cast<float>(var) + p.bn.eps);
        const float g = p.bn.gamma ? __ldg(p.bn.gamma + c) : 1.f;
        const float b = p.bn.beta ? __ldg(p.bn.beta + c) : 0.f;
        const float scv = g * istd;
        const float shv = b - mf * scv;
        sm[e] = scv;
        sm[512 + e] = shv;
        if (mu0 == 0 && rank == 0) {
            p.bn.aff[c] = mf;
            p.bn.aff[p.Cout + c] = istd;
            p.bn.aff[2 * p.Cout + c] = scv;
            p.bn.aff[3 * p.Cout + c] = shv;
            if (p.bn.rm) p.bn.rm[c] = (1.f - p.bn.momentum) * p.bn.rm[c] + p.bn.momentum * mf;
            if (p.bn.rv) p.bn.rv[c] = (1.f - p.bn.momentum) * p.bn.rv[c] + p.bn.momentum * static_cast<float>(var * p.bn.unbias);
        }
    }
    named_bar_sync(5, 512);
    if (S == 0) return;                       // (no barrier below this point)
    float sc[8], sh[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        sc[i] = sm[cx * 8 + i];
        sh[i] = sm[512 + cx * 8 + i];
    }
    if (e == 0) trace(212);
    // ---- normalise: consume a slot, store z, refill the slot with the load W slots ahead
    for (int base = 0; base < S; base += W) {
#pragma unroll
        for (int u = 0; u < W; ++u) {
            if (pixb[u] >= 0) {
                const uint32_t* w4 = reinterpret_cast<const uint32_t*>(&buf[u]);
                const uint32_t* r4 = reinterpret_cast<const uint32_t*>(&rbuf[u]);
                uint4 o;
                uint32_t* o4 = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float a = fmaf(bf16_lo(w4[j]), sc[2 * j], sh[2 * j]);
                    float b = fmaf(bf16_hi(w4[j]), sc[2 * j + 1], sh[2 * j + 1]);
                    if (HAS_RES) { a += bf16_lo(r4[j]); b += bf16_hi(r4[j]); }
                    if (p.bn.relu) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); }
                    o4[j] = pack_bf16x2(a, b);
                }
                st_na_v4(zbase + static_cast<long long>(pixb[u]) * p.bn.z_ld, o);
            }
            pixb[u] = (base + W + u) < S ? next_pix() : -1;
            if (pixb[u] >= 0) {
                buf[u] = ld_cg_v4(ybase + static_cast<long long>(pixb[u]) * p.out_ld);
                if (HAS_RES) rbuf[u] = ld_nc_v4(rbase + static_cast<long long>(pixb[u]) * p.bn.res_ld);
            }
        }
    }
    if (e == 0) trace(213);
}

// Fused send half of the SyncBN all-reduce (several ranks), called by the 512 epilogue threads after their statistic
// atomics: every CTA fences and takes a ticket; the last one sees the rank's final sums in L2 and stores them into the
// peers' inboxes (flag-in-data protocol, gs_comm.cuh).  Out of line: it runs once per kernel and must not cost the main
// loop any registers.
__device__ __noinline__ void ig_syncbn_push(const IgemmParams& p, uint8_t* staging) {
    volatile uint32_t* is_last = reinterpret_cast<volatile uint32_t*>(staging + 16384);   // behind the `red` scratch
    __threadfence();
    named_bar_sync(5, 512);
    if (threadIdx.x == 64) {
        const unsigned long long ticket = atomicAdd(p.sync_ticket, 1ull);
        *is_last = (ticket + 1 == static_cast<unsigned long long>(gridDim.x)) ? 1u : 0u;
    }
    named_bar_sync(5, 512);
    if (*is_last) {
        __threadfence();
        syncbn_push_block(p.stats, 2 * p.Cout, p.sync.peers, p.sync.rank, p.sync.world, p.sync.seq_dev,
                          static_cast<int>(threadIdx.x) - 64, 512);
    }
}

template <int CG, bool WIDE>
__global__ void __launch_bounds__(kIgThreads, 1)
igemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
             const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR,
             const __grid_constant__ IgemmParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    constexpr int kIgStages = IgCfg<CG>::kStages;           // compile-time MAXIMUM (barrier arrays, shared-memory carve-up)
    // ring geometry: the pair ring's 5 x 32 KB, or (wide tiles) 4 x 40 KB -- compile-time, like everything `WIDE` selects
    constexpr int n_stages = WIDE ? 4 : kIgStages;
    constexpr int stage_bytes = WIDE ? (kIgABytes + 24 * 1024) : IgCfg<CG>::kStageBytes;
    uint8_t* stage_base = smem;
    uint8_t* staging = smem + kIgStages * IgCfg<CG>::kStageBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(staging + kIgStagingBytes);
    uint64_t* full_bar = bars;                  // [kIgStages]
    uint64_t* empty_bar = bars + kIgStages;     // [kIgStages]
    uint64_t* tfull_bar = bars + 2 * kIgStages; // [2]
    uint64_t* tempty_bar = tfull_bar + 2;       // [2]
    uint64_t* res_bar = tempty_bar + 2;          // [4] residual sub-tile landed in the staging ring
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(res_bar + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    pdl_trigger();                       // the next kernel's blocks may be scheduled; they wait for our completion
    if (threadIdx.x == 0) trace(0);

    if (warp == 0 && elect_one()) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        if (!p.direct) tma_prefetch_desc(&tmC);
        if (!p.direct && p.residual != nullptr) tma_prefetch_desc(&tmR);
        for (int i = 0; i < kIgStages; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull_bar[i], 1);
            mbar_init(&tempty_bar[i], 16 * CG);  // one arrive per epilogue warp (of both CTAs of a pair: leader's barrier)
        }
        for (int i = 0; i < 4; ++i) mbar_init(&res_bar[i], 1);
        fence_mbar_init();
    }
    if (warp == 1) {
        if (CG == 2) tmem_alloc_pair(tmem_ptr, 512);
        else tmem_alloc(tmem_ptr, 512);
    }
    tc_fence_before_sync();
    if (CG == 2) cluster_sync_all();   // the peer's barriers must be initialised before anything signals them
    else __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_ptr;
    // everything above (barrier init, TMEM allocation, descriptor prefetch) overlapped the tail of the previous kernel;
    // from here on global memory is touched: wait for the earlier kernels of the stream to complete
    pdl_wait();

    // Work unit = one CTA (CG 1) or one CTA pair (CG 2).  Unit u owns ONE n-tile (u % n_tiles) for the whole kernel and
    // walks m-units u / n_tiles + i * groups; CTA `rank` of the unit computes m-tile  m_unit * CG + rank  (a pair with
    // an odd tile count ends on a phantom tile: every load is out of bounds = zero, every store is clipped).
    const int rank = CG == 2 ? static_cast<int>(cluster_ctarank()) : 0;
    const int unit = blockIdx.x / CG;
    const int groups = (gridDim.x / CG) / p.n_tiles;   // host guarantees (gridDim.x / CG) % n_tiles == 0
    const int m_units = (p.m_tiles + CG - 1) / CG;
    const int nt = unit % p.n_tiles;
    const int mu0 = unit / p.n_tiles;
    const int tiles_hw = p.tiles_h * p.tiles_w;
    constexpr bool wide = WIDE;
    if (threadIdx.x == 0) trace(1);

    if (warp == 0) {
        // =========================== TMA producer ===========================
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            int tr_i = 0;
            // the n_tiles units that share an activation tile run side by side (L2 reuse), the weight tile of a unit
            // never changes, and the DynBN statistics of its output channels stay in registers until the end.
            for (int mu = mu0; mu < m_units; mu += groups) {
                const int mt = mu * CG + rank;
                const int img = mt / tiles_hw;
                const int rem = mt - img * tiles_hw;
                const int h0 = (rem / p.tiles_w) * p.TH;
                const int w0 = (rem % p.tiles_w) * p.TW;
                const int n0 = nt * p.nt_w;
                int n_valid = p.Cout - n0;
                if (n_valid > p.nt_w) n_valid = p.nt_w;
                // this CTA's share of the weight tile: all of it, or (pair) n_half channels from n0 + rank * n_half
                const int n_half = ((n_valid + 15) & ~15) / CG;
                const int nb0 = n0 + rank * (CG == 2 ? n_half : 0);
                const int b_boxes = ((CG == 2 ? n_half : n_valid) + 63) >> 6;   // MN-major B: [64 k-rows][64 n] boxes
                // (pair: the leader's barrier counts the bytes of BOTH CTAs; their shares have the same size)
                uint32_t tx_bytes = CG * (kIgABytes + (p.b_mn ? b_boxes * 8192 : p.b_box_rows * 128));
                // wide tile: accumulator column j <-> channel n0 + j needs CTA r to stage rows [128 r, 128 r + 128) for the
                // N = 256 MMA and rows [256 + h2 r, 256 + h2 r + h2) for the second one
                if (wide) tx_bytes = CG * (kIgABytes + (p.b_mn ? 3 * 8192 : (128 + p.h2) * 128));   // (even split: 2 hp = 128 + h2)
                for (int r = 0; r < p.taps_h; ++r) {
                    const int ih = h0 * p.in_mul + p.base + r * p.step;
                    for (int s = 0; s < p.taps_w; ++s) {
                        const int iw = w0 * p.in_mul + p.base + s * p.step;
                        for (int kc = 0; kc < p.kchunks; ++kc) {
                            mbar_wait(&empty_bar[stage], phase ^ 1);
                            if (tr_i < 64) trace(16 + tr_i++);
                            uint8_t* a_dst = stage_base + stage * stage_bytes;
                            uint8_t* b_dst = a_dst + kIgABytes;
                            if (CG == 1) {
                                mbar_arrive_expect_tx(&full_bar[stage], tx_bytes);
                                tma_load_4d(a_dst, &tmA, &full_bar[stage], kc * 64, iw, ih, img);
                                if (p.b_mn) {
                                    for (int j = 0; j < b_boxes; ++j)
                                        tma_load_4d(b_dst + j * 8192, &tmB, &full_bar[stage], n0 + j * 64, s, r, kc * 64);
                                } else {
                                    tma_load_4d(b_dst, &tmB, &full_bar[stage], kc * 64, s, r, n0);
                                }
                            } else {
                                if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], tx_bytes);
                                tma_load_4d_pair(a_dst, &tmA, &full_bar[stage], kc * 64, iw, ih, img);
                                if (wide) {
                                    const int c1 = n0 + rank * 128, c2 = n0 + 256 + rank * p.h2;
                                    if (p.b_mn) {
                                        tma_load_4d_pair(b_dst, &tmB, &full_bar[stage], c1, s, r, kc * 64);
                                        tma_load_4d_pair(b_dst + 8192, &tmB, &full_bar[stage], c1 + 64, s, r, kc * 64);
                                        tma_load_4d_pair(b_dst + 16384, &tmB, &full_bar[stage], c2, s, r, kc * 64);
                                    } else if (p.even) {   // two boxes of hp = nt_w / 4 rows (b_box_rows == hp)
                                        const int hp = p.nt_w >> 2;
                                        tma_load_4d_pair(b_dst, &tmB, &full_bar[stage], kc * 64, s, r, n0 + rank * hp);
                                        tma_load_4d_pair(b_dst + hp * 128, &tmB, &full_bar[stage], kc * 64, s, r,
                                                         n0 + 2 * hp + rank * hp);
                                    } else {      // weight boxes of 32 rows (b_box_rows == 32)
                                        for (int j = 0; j < 4; ++j)
                                            tma_load_4d_pair(b_dst + j * 4096, &tmB, &full_bar[stage], kc * 64, s, r, c1 + 32 * j);
                                        for (int j = 0; j < (p.h2 >> 5); ++j)
                                            tma_load_4d_pair(b_dst + (4 + j) * 4096, &tmB, &full_bar[stage], kc * 64, s, r,
                                                             c2 + 32 * j);
                                    }
                                } else if (p.b_mn) {
                                    for (int j = 0; j < b_boxes; ++j)
                                        tma_load_4d_pair(b_dst + j * 8192, &tmB, &full_bar[stage], nb0 + j * 64, s, r,
                                                         kc * 64);
                                } else {
                                    tma_load_4d_pair(b_dst, &tmB, &full_bar[stage], kc * 64, s, r, nb0);
                                }
                            }
                            if (++stage == n_stages) { stage = 0; phase ^= 1; }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // =========================== MMA issuer (pair: the leader CTA only) ===========================
        if (rank == 0 && elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            int tr_i = 0;
            const int iters = p.taps_h * p.taps_w * p.kchunks;
            for (int mu = mu0; mu < m_units; mu += groups) {
                int n_valid = p.Cout - nt * p.nt_w;
                if (n_valid > p.nt_w) n_valid = p.nt_w;
                const bool even = wide && p.even != 0;
                const uint32_t n1 = even ? static_cast<uint32_t>(p.nt_w >> 1) : 256u;          // columns of the first MMA
                const uint32_t umma_n = wide ? n1 : static_cast<uint32_t>((n_valid + 15) & ~15);
                const uint32_t idesc = make_idesc_bf16(128 * CG, umma_n, false, p.b_mn != 0);
                const uint32_t idesc2 = make_idesc_bf16(128 * CG, even ? n1 : static_cast<uint32_t>(2 * p.h2), false, p.b_mn != 0);
                const uint32_t b2_off = even ? static_cast<uint32_t>((p.nt_w >> 2) * 128) : 128u * 128u;   // K-major part 2
                if (CG == 2) mbar_wait_cluster(&tempty_bar[acc], acc_phase ^ 1);
                else mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
                tc_fence_after_sync();
                const uint32_t d_tmem = tmem_base + acc * 256;
                uint32_t accumulate = 0;
                int kc = 0;
                for (int it = 0; it < iters; ++it) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after_sync();
                    if (tr_i < 64) trace(80 + tr_i++);
                    const uint32_t a_addr = smem_u32(stage_base + stage * stage_bytes);
                    const uint32_t b_addr = a_addr + kIgABytes;
                    int krem = p.Kc - kc * 64;
                    const int ksteps = krem >= 64 ? 4 : (krem + 15) >> 4;
#pragma unroll 1
                    for (int k = 0; k < ksteps; ++k) {
                        const uint64_t adesc = make_smem_desc_sw128(a_addr + k * 32, 16, 1024);
                        const uint64_t bdesc = p.b_mn ? make_smem_desc_sw128(b_addr + k * 2048, 8192, 1024)
                                                      : make_smem_desc_sw128(b_addr + k * 32, 16, 1024);
                        if (CG == 2) umma_bf16_ss_pair(d_tmem, adesc, bdesc, idesc, accumulate);
                        else umma_bf16_ss(d_tmem, adesc, bdesc, idesc, accumulate);
                        if (wide) {   // second part of the wide tile: accumulator columns [256, 256 + 2 h2)
                            const uint64_t bdesc2 = p.b_mn ? make_smem_desc_sw128(b_addr + 16384 + k * 2048, 8192, 1024)
                                                           : make_smem_desc_sw128(b_addr + b2_off + k * 32, 16, 1024);
                            umma_bf16_ss_pair(d_tmem + n1, adesc, bdesc2, idesc2, accumulate);
                        }
                        accumulate = 1;
                    }
                    // frees the smem slot (of both CTAs) once these MMAs retire
                    if (CG == 2) umma_commit_pair(&empty_bar[stage]);
                    else umma_commit(&empty_bar[stage]);
                    if (++kc == p.kchunks) kc = 0;
                    if (++stage == n_stages) { stage = 0; phase ^= 1; }
                }
                if (CG == 2) umma_commit_pair(&tfull_bar[acc]);
                else umma_commit(&tfull_bar[acc]);
                if (wide) acc_phase ^= 1;           // ONE accumulator (up to 384 columns): the same buffer every tile
                else {
                    acc ^= 1;
                    if (acc == 0) acc_phase ^= 1;
                }
            }
        }
    } else {
        // =========================== epilogue (warps 2..17) ===========================
        // 16 warps = 4 sub-tile groups x 4 TMEM sub-partitions: group `sub` owns the 64-column sub-tiles sub, sub + 4 (the
        // second only exists in a wide tile of 320 / 384 columns) -- one TMA-store box each -- warp quarter q its 32 rows.
        // The sub-tiles of a round drain CONCURRENTLY -- the per-tile epilogue is instruction-issue bound, so it is spread
        // over all four schedulers (4 warps each) -- each group through its own 16 KB staging slot, named barrier and
        // TMA-store thread:   TMEM -> registers -> affine / residual / ReLU -> bf16 -> swizzled staging -> TMA store.
        // DynBN statistics: every lane re-reads the 32 rows its own warp just staged (column pair = lane,
        // conflict-free) and keeps sum / sum^2 in registers across ALL tiles of the CTA; fp64 atomics once at the end.
        const int q = warp & 3;            // TMEM sub-partition of this warp (hardware rule: warp id % 4)
        const int sub = (warp - 2) >> 2;   // sub-tile group 0..3
        const int row = q * 32 + lane;     // accumulator row == pixel within the tile
        const int g_tid = threadIdx.x - 64 - sub * 128;   // thread index inside the group (0 = its TMA thread)
        const int th = row / p.TW;
        const int tw = row - th * p.TW;
        int acc = 0;
        uint32_t acc_phase = 0;
        uint32_t res_phase = 0;
        int tr_t = 0;
        uint8_t* const slot = staging + sub * (128 * 128);
        const uint32_t slot_u32 = smem_u32(slot);
        // shared-space addresses (32 bit): the slot is 1024-aligned, so the 128B-swizzle XOR of 16-byte chunk k of this
        // thread's row is  wr ^ (k << 4)
        const uint32_t wr = (slot_u32 + row * 128) ^ (static_cast<uint32_t>(row & 7) << 4);
        const uint32_t st_lane = (static_cast<uint32_t>(lane >> 2) << 4) | (static_cast<uint32_t>(lane & 3) << 2);
        const uint32_t st_base = slot_u32 + q * (32 * 128);
        // statistics of this lane's column pair, packed (col, col + 1) fp32x2, one set per round (sub-tiles sub, sub + 4)
        uint64_t st_s2a = 0ull, st_q2a = 0ull, st_s2b = 0ull, st_q2b = 0ull;   // (scalars: a dynamically indexed array would live in local memory)
        const int n0 = nt * p.nt_w;
        int n_valid = p.Cout - n0;
        if (n_valid > p.nt_w) n_valid = p.nt_w;
        const int nsub = (n_valid + 63) >> 6;              // 64-column sub-tiles of this n-tile (<= 4, wide: 5 or 6)
        const int nchunks = (n_valid + 31) >> 5;
        const bool res_tma = (!p.direct) && (p.residual != nullptr);
        for (int mu = mu0; mu < m_units; mu += groups) {
            const int mt = mu * CG + rank;
            const int img = mt / tiles_hw;
            const int rem = mt - img * tiles_hw;
            const int h0 = (rem / p.tiles_w) * p.TH;
            const int w0 = (rem % p.tiles_w) * p.TW;
            const int h = h0 + th, w = w0 + tw;
            const bool valid = (img < p.N) && (h < p.Ho) && (w < p.Wo);
            const long long pix = (static_cast<long long>(img) * p.Ho + h) * p.Wo + w;

#pragma unroll 1
            for (int rd = 0; rd < (WIDE ? 2 : 1); ++rd) {
                const int st = sub + 4 * rd;                       // sub-tile of this round
                if (rd == 1 && st >= nsub) break;                  // (round 0 always runs: it hands the accumulator back)
                const bool has_cols = st < nsub;                   // does this group own any column in this round?
                const bool last_round = st + 4 >= nsub;            // the accumulator is released after the last drain
                if (!p.direct && has_cols) {
                    // the previous store must have finished READING the slot before anybody refills it; with a
                    // residual the refill is the TMA load of the residual tile IN PLACE (it overlaps the MMA main loop of
                    // this tile; the epilogue adds it at the very positions it overwrites)
                    if (g_tid == 0) {
                        tma_store_wait_read0();
                        if (res_tma) {
                            mbar_arrive_expect_tx(&res_bar[sub], 128 * 128);
                            tma_load_4d(slot, &tmR, &res_bar[sub], n0 + st * 64, w0, h0, img);
                        }
                    }
                    if (!res_tma) named_bar_sync(1 + sub, 128);
                }

                if (rd == 0) {
                    mbar_wait(&tfull_bar[acc], acc_phase);
                    tc_fence_after_sync();
                    if (g_tid == 0 && sub == 0 && tr_t < 16) trace(144 + 4 * tr_t);
                }
                const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * 256 + st * 64;

                if (!has_cols) {
                    // nothing to drain for this group: still hand the accumulator back
                    tc_fence_before_sync();
                    __syncwarp();
                    if (lane == 0) ig_release_acc<CG>(&tempty_bar[acc]);
                } else if (p.direct) {
#pragma unroll 1
                    for (int cc = 0; cc < 2; ++cc) {
                        const int c = st * 2 + cc;
                        if (c >= nchunks) break;
                        uint32_t raw[32];
                        tmem_ld_32x32b_x32(t_addr + cc * 32, raw);
                        tmem_ld_wait();
                        float v[32];
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[i]);
                        const int col0 = n0 + c * 32;
                        if (p.scale != nullptr) {
#pragma unroll
                            for (int i = 0; i < 32; ++i)
                                if (col0 + i < p.Cout) v[i] *= __ldg(p.scale + col0 + i);
                        }
                        if (p.shift != nullptr) {
#pragma unroll
                            for (int i = 0; i < 32; ++i)
                                if (col0 + i < p.Cout) v[i] += __ldg(p.shift + col0 + i);
                        }
                        if (p.residual != nullptr && valid) {
                            const __nv_bfloat16* rp = p.residual + pix * p.res_ld + col0;
#pragma unroll
                            for (int i = 0; i < 32; ++i)
                                if (col0 + i < p.Cout) v[i] += __bfloat162float(rp[i]);
                        }
                        if (p.relu) {
#pragma unroll
                            for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
                        }
                        if (valid) {
                            if (p.out_f32) {
                                float* op = reinterpret_cast<float*>(p.out) + pix * p.out_ld + col0;
#pragma unroll
                                for (int i = 0; i < 32; ++i)
                                    if (col0 + i < p.Cout) op[i] = v[i];
                            } else {
                                __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.out) + pix * p.out_ld + col0;
#pragma unroll
                                for (int i = 0; i < 32; ++i)
                                    if (col0 + i < p.Cout) op[i] = __float2bfloat16_rn(v[i]);
                            }
                        }
                    }
                    if (last_round) {
                        tc_fence_before_sync();
                        __syncwarp();
                        if (lane == 0) ig_release_acc<CG>(&tempty_bar[acc]);
                    }
                } else {
                    // the sub-tile drains in four 16-column groups through two alternating register sets: the TMEM load of
                    // group g+1 is in flight while group g is converted and staged (18 warps leave 96 registers per thread)
                    uint32_t ra[16], rb[16];
                    const int ngrp = min(4, (n_valid - st * 64 + 15) >> 4);   // 16-column groups with active columns
                    tmem_ld_32x32b_x16(t_addr, ra);
                    if (res_tma) mbar_wait(&res_bar[sub], res_phase);
                    tmem_ld_wait();
                    auto stage_group = [&](const uint32_t (&raw)[16], int g) {
                        float v[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(raw[i]);
                        const int col0 = n0 + st * 64 + g * 16;
                        if (p.scale != nullptr || p.shift != nullptr) {
                            if (col0 + 16 <= p.Cout) {       // whole group inside the active width: vector loads
#pragma unroll
                                for (int g4 = 0; g4 < 4; ++g4) {
                                    if (p.scale != nullptr) {
                                        const float4 s4 = __ldg(reinterpret_cast<const float4*>(p.scale + col0) + g4);
                                        v[g4 * 4 + 0] *= s4.x; v[g4 * 4 + 1] *= s4.y;
                                        v[g4 * 4 + 2] *= s4.z; v[g4 * 4 + 3] *= s4.w;
                                    }
                                    if (p.shift != nullptr) {
                                        const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.shift + col0) + g4);
                                        v[g4 * 4 + 0] += b4.x; v[g4 * 4 + 1] += b4.y;
                                        v[g4 * 4 + 2] += b4.z; v[g4 * 4 + 3] += b4.w;
                                    }
                                }
                            } else {
#pragma unroll
                                for (int i = 0; i < 16; ++i)
                                    if (col0 + i < p.Cout) {
                                        if (p.scale != nullptr) v[i] *= __ldg(p.scale + col0 + i);
                                        if (p.shift != nullptr) v[i] += __ldg(p.shift + col0 + i);
                                    }
                            }
                        }
                        if (res_tma) {
#pragma unroll
                            for (int g8 = 0; g8 < 2; ++g8) {
                                const uint4 u = lds128(wr ^ ((g * 2 + g8) << 4));
                                v[g8 * 8 + 0] += bf16_lo(u.x); v[g8 * 8 + 1] += bf16_hi(u.x);
                                v[g8 * 8 + 2] += bf16_lo(u.y); v[g8 * 8 + 3] += bf16_hi(u.y);
                                v[g8 * 8 + 4] += bf16_lo(u.z); v[g8 * 8 + 5] += bf16_hi(u.z);
                                v[g8 * 8 + 6] += bf16_lo(u.w); v[g8 * 8 + 7] += bf16_hi(u.w);
                            }
                        }
                        if (p.relu) {
#pragma unroll
                            for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
                        }
                        if (!valid) {
#pragma unroll
                            for (int i = 0; i < 16; ++i) v[i] = 0.f;  // keeps OOB pixels out of the statistics
                        }
#pragma unroll
                        for (int g8 = 0; g8 < 2; ++g8) {
                            uint4 u;
                            u.x = pack_bf16x2(v[g8 * 8 + 0], v[g8 * 8 + 1]);
                            u.y = pack_bf16x2(v[g8 * 8 + 2], v[g8 * 8 + 3]);
                            u.z = pack_bf16x2(v[g8 * 8 + 4], v[g8 * 8 + 5]);
                            u.w = pack_bf16x2(v[g8 * 8 + 6], v[g8 * 8 + 7]);
                            sts128(wr ^ ((g * 2 + g8) << 4), u);
                        }
                    };
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        if (g < ngrp) {
                            if (g + 1 < ngrp) {
                                if (g & 1) tmem_ld_32x32b_x16(t_addr + (g + 1) * 16, ra);
                                else tmem_ld_32x32b_x16(t_addr + (g + 1) * 16, rb);
                            }
                            if (g & 1) stage_group(rb, g);
                            else stage_group(ra, g);
                            if (g + 1 < ngrp) tmem_ld_wait();
                            if (last_round && g == (ngrp > 1 ? ngrp - 2 : 0)) {
                                // the last group of this warp's accumulator share is in registers -> hand the buffer back
                                tc_fence_before_sync();
                                __syncwarp();
                                if (lane == 0) ig_release_acc<CG>(&tempty_bar[acc]);
                                if (g_tid == 0 && sub == 0 && tr_t < 16) trace(145 + 4 * tr_t);
                            }
                        }
                    }
                    // make the staged rows visible to the TMA engine, then one thread of the group stores the box while
                    // every warp accumulates the statistics of the rows it staged
                    fence_proxy_async_smem();
                    named_bar_sync(1 + sub, 128);
                    if (g_tid == 0) {
                        tma_store_4d(&tmC, slot, n0 + st * 64, w0, h0, img);
                        tma_store_commit();
                        if (sub == 0 && tr_t < 16) trace(146 + 4 * tr_t);
                    }
                    if (p.stats != nullptr && st * 64 + lane * 2 < n_valid) {
                        // lane owns columns (2*lane, 2*lane+1) of the sub-tile over the 32 rows its warp staged: 32
                        // conflict-free LDS.32 (row r: chunk (lane>>2) ^ (r & 7), word lane & 3) -> packed fp32x2
                        // accumulation (FADD2 / FFMA2) of both columns at once
                        uint64_t s2 = rd == 0 ? st_s2a : st_s2b, q2 = rd == 0 ? st_q2a : st_q2b;
#pragma unroll
                        for (int r = 0; r < 32; ++r) {
                            const uint32_t wv = lds32(st_base + r * 128 + (st_lane ^ ((r & 7) << 4)));
                            const uint64_t pv = pack_f32x2(bf16_lo(wv), bf16_hi(wv));
                            s2 = add_f32x2(s2, pv);
                            q2 = fma_f32x2(pv, pv, q2);
                        }
                        if (rd == 0) { st_s2a = s2; st_q2a = q2; } else { st_s2b = s2; st_q2b = q2; }
                    }
                    if (res_tma) res_phase ^= 1;
                }
            }
            if (g_tid == 0 && sub == 0 && tr_t < 16) trace(147 + 4 * tr_t);
            ++tr_t;
            if (wide) acc_phase ^= 1;
            else {
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
        }
        if (!p.direct && g_tid == 0) tma_store_wait_all0();
        if (p.stats != nullptr) {
            // one flush per kernel: the 4 row-quarters are summed through the now idle staging buffer, then every column
            // goes out with two fp64 atomics.  red: [4 quarters][1024]: sums at 0..511, squares at 512..1023
            named_bar_sync(5, 512);
            float* red = reinterpret_cast<float*>(staging);
            {
                const int c = sub * 64 + lane * 2;
                red[q * 1024 + c] = f32x2_lo(st_s2a);
                red[q * 1024 + c + 1] = f32x2_hi(st_s2a);
                red[q * 1024 + 512 + c] = f32x2_lo(st_q2a);
                red[q * 1024 + 512 + c + 1] = f32x2_hi(st_q2a);
                const int c2 = c + 256;                                  // round 1: sub-tile sub + 4
                red[q * 1024 + c2] = f32x2_lo(st_s2b);
                red[q * 1024 + c2 + 1] = f32x2_hi(st_s2b);
                red[q * 1024 + 512 + c2] = f32x2_lo(st_q2b);
                red[q * 1024 + 512 + c2 + 1] = f32x2_hi(st_q2b);
            }
            named_bar_sync(5, 512);
            for (int c = threadIdx.x - 64; c < n_valid; c += 512) {
                const int col = n0 + c;
                const float su = red[c] + red[1024 + c] + red[2048 + c] + red[3072 + c];
                const float sq = red[512 + c] + red[1536 + c] + red[2560 + c] + red[3584 + c];
                atomicAdd(p.stats + col, static_cast<double>(su));
                atomicAdd(p.stats + p.Cout + col, static_cast<double>(sq));
            }
            if (p.sync.world > 1) ig_syncbn_push(p, staging);
            if (p.bn.on) {
                if (p.bn.res != nullptr) ig_bn_tail<CG, true>(p, mu0, groups, m_units, rank, nt, staging);
                else ig_bn_tail<CG, false>(p, mu0, groups, m_units, rank, nt, staging);
            }
        }
    }

    tc_fence_before_sync();
    if (CG == 2) cluster_sync_all();   // nobody leaves (or frees TMEM) while the peer may still signal or read
    else __syncthreads();
    if (warp == 1) {
        if (CG == 2) tmem_dealloc_pair(tmem_base, 512);
        else tmem_dealloc(tmem_base, 512);
    }
}

// Spatial tile of `total` (128 or 64) pixels: widest power-of-two row segment that fits.
static void choose_tile(int Wo, int total, int* TH, int* TW) {
    int tw = 1;
    while (tw < total && tw < Wo) tw <<= 1;
    *TW = tw;
    *TH = total / tw;
}

struct IgemmLaunch {
    // activation ("input" of the GEMM A operand)
    const void* a_ptr; int a_C; long long a_ld; int a_H, a_W;
    int a_estride;         // element stride of the spatial traversal (fwd conv stride), 1 or 2
    // weights: 4-D [rows_max][kh][kw][cols_max] with active extents rows (=Cout) x cols (=Kc)
    const void* b_ptr; int b_rows_max, b_cols_max;
    int b_mn;              // 1: b_ptr is W[K rows = Kc][kh][kw][N cols = Cout] with pitch b_cols_max (dgrad in place)
    int N, Ho, Wo, Cout, Kc, kh, kw;
    int in_mul, base, step;
    void* out; long long out_ld; int out_f32;
    const float* scale; const float* shift; const void* residual; long long res_ld; int relu; double* stats;
    const gs_bn_bwd_fuse* fuse;   // reserved, must be NULL
    const gs_sync_desc* sync;     // forward only: push the final statistics to the SyncBN peers (phase 1)
    const IgemmParams::BnTail* bn;   // forward only: fused DynBN apply (gs_conv2d_fwd_bn), NULL otherwise
};

static int launch_igemm(const IgemmLaunch& L, cudaStream_t stream) {
    GS_REQUIRE(L.a_ld % 8 == 0 && (reinterpret_cast<uintptr_t>(L.a_ptr) & 15) == 0,
               "conv: activation pitch (%lld) must be a multiple of 8 elements and 16-byte aligned", L.a_ld);
    GS_REQUIRE(L.b_cols_max % 8 == 0 && (reinterpret_cast<uintptr_t>(L.b_ptr) & 15) == 0,
               "conv: weight inner extent (%d) must be a multiple of 8", L.b_cols_max);
    GS_REQUIRE(L.a_estride >= 1 && L.a_estride <= 2, "conv: stride %d not supported (1 or 2)", L.a_estride);
    IgemmParams p{};
    p.N = L.N; p.Ho = L.Ho; p.Wo = L.Wo;
    choose_tile(L.Wo, 128, &p.TH, &p.TW);
    p.tiles_h = (int)gs_ceil_div(L.Ho, p.TH);
    p.tiles_w = (int)gs_ceil_div(L.Wo, p.TW);
    p.m_tiles = L.N * p.tiles_h * p.tiles_w;
    p.Cout = L.Cout; p.Kc = L.Kc; p.kchunks = (int)gs_ceil_div(L.Kc, 64);
    p.taps_h = L.kh; p.taps_w = L.kw;
    p.in_mul = L.in_mul; p.base = L.base; p.step = L.step;
    // CTA pairs when the K loop is long enough for the halved operand traffic to matter (short loops are bound by
    // the epilogue, where the lock-step of a pair costs a little); GS_IGEMM_CTA_GROUP=1 forces the single-CTA kernel.
    static const int cg_env = [] { const char* e = getenv("GS_IGEMM_CTA_GROUP"); return e ? atoi(e) : 2; }();
    static const int cg_min_iters = [] { const char* e = getenv("GS_IGEMM_PAIR_MIN_ITERS"); return e ? atoi(e) : 8; }();
    const int iters = L.kh * L.kw * p.kchunks;
    const int cg = (cg_env == 2 && p.m_tiles >= 2 && iters >= cg_min_iters) ? 2 : 1;
    const bool tma_store_ok = !L.out_f32 && (L.out_ld % 8 == 0) && (L.Cout % 8 == 0) &&
                              ((reinterpret_cast<uintptr_t>(L.out) & 15) == 0);
    // wide n-tiles (one 320 / 384-column accumulator instead of 256 + 64 / 256 + 128): pairs, long K loops (the epilogue
    // of a single-buffered accumulator does not overlap the next tile's main loop), widths that tile evenly
    static const int wide_env = [] { const char* e = getenv("GS_IGEMM_WIDE"); return e ? atoi(e) : 1; }();
    static const int wide_min_iters = [] { const char* e = getenv("GS_IGEMM_WIDE_MIN_ITERS"); return e ? atoi(e) : 16; }();
    p.nt_w = 256; p.wide = 0; p.h2 = 0;
    if (wide_env && cg == 2 && iters >= wide_min_iters && tma_store_ok && L.Cout > 256) {
        if (L.Cout % 320 == 0) p.nt_w = 320;
        else if (L.Cout % 384 == 0) p.nt_w = 384;
        if (p.nt_w != 256) { p.wide = 1; p.h2 = (p.nt_w - 256) / 2; }
    }
    static const int even_env = [] { const char* e = getenv("GS_IGEMM_WIDE_EVEN"); return e ? atoi(e) : 1; }();
    p.even = (p.wide && !L.b_mn && even_env) ? 1 : 0;
    p.n_tiles = (int)gs_ceil_div(L.Cout, p.nt_w);
    if (cg == 1) { p.stages = IgCfg<1>::kStages; p.stage_bytes = IgCfg<1>::kStageBytes; }
    else if (!p.wide) { p.stages = IgCfg<2>::kStages; p.stage_bytes = IgCfg<2>::kStageBytes; }
    else { p.stages = 4; p.stage_bytes = kIgABytes + 24 * 1024; }     // 4 x 40 KB = the pair ring's 5 x 32 KB
    p.b_box_rows = p.wide ? (p.even ? p.nt_w / 4 : 32) : (L.Cout >= 256 ? 256 : gs_round_up(L.Cout, 16)) / cg;
    p.b_mn = L.b_mn;
    p.out = L.out; p.out_ld = L.out_ld; p.out_f32 = L.out_f32; p.relu = L.relu;
    p.scale = L.scale; p.shift = L.shift;
    p.residual = reinterpret_cast<const __nv_bfloat16*>(L.residual); p.res_ld = L.res_ld;
    p.stats = L.stats;
    p.sync = SyncArgs{};
    p.sync.world = 1;
    p.sync_ticket = nullptr;
    if (L.sync != nullptr && L.sync->world > 1) {
        GS_REQUIRE(L.stats != nullptr, "conv: the SyncBN push needs the statistics accumulator");
        GS_REQUIRE(L.sync->phase == 1, "conv: sync descriptor must have phase == 1 (push only), got %d", L.sync->phase);
        GS_REQUIRE(L.sync->world <= kCommMaxWorld && L.sync->rank >= 0 && L.sync->rank < L.sync->world &&
                       L.sync->peer_inboxes != nullptr && L.sync->seq_dev != nullptr,
                   "conv: bad sync descriptor (rank %d / world %d)", L.sync->rank, L.sync->world);
        GS_REQUIRE(2 * L.Cout <= kCommSlotDoubles, "conv: %d channels exceed the exchange slot", L.Cout);
        for (int r = 0; r < L.sync->world; ++r)
            p.sync.peers.p[r] = reinterpret_cast<ulonglong2*>(const_cast<void*>(L.sync->peer_inboxes[r]));
        p.sync.rank = L.sync->rank;
        p.sync.world = L.sync->world;
        p.sync.seq_dev = reinterpret_cast<unsigned long long*>(L.sync->seq_dev);
        p.sync.phase = 1;
        p.sync_ticket = reinterpret_cast<unsigned long long*>(L.stats + 2 * L.Cout);   // first scratch word behind the sums
    }
    GS_REQUIRE(L.fuse == nullptr, "dgrad: the fused BN-backward reduction was removed (measured slower than gs_bn_bwd_reduce); "
                                  "pass fuse = NULL");
    p.direct = tma_store_ok ? 0 : 1;
    p.bn = IgemmParams::BnTail{};
    if (L.bn != nullptr) {
        GS_REQUIRE(tma_store_ok && L.stats != nullptr, "conv + BN: needs the bf16 TMA-store epilogue (Co %% 8 == 0) and a statistics buffer");
        GS_REQUIRE(p.sync.world <= 1, "conv + BN: several ranks are not supported by the fused path (use gs_conv2d_fwd_syncbn + gs_bn_apply_train)");
        p.bn = *L.bn;
        p.bn.on = 1;
        p.bn.barrier = reinterpret_cast<unsigned long long*>(L.stats + 2 * L.Cout);
        p.bn.timeout_ns = 2000000000ull;    // all CTAs are co-resident: 2 s means a bug, trap instead of hanging the GPU
        p.bn.tw_shift = 0;
        while ((1 << p.bn.tw_shift) < p.TW) ++p.bn.tw_shift;
    }
    GS_REQUIRE(!(p.direct && L.stats), "conv: statistics need the bf16 TMA-store epilogue (Co %% 8 == 0)");
    if (L.residual) {
        GS_REQUIRE(L.res_ld % 8 == 0 && L.Cout % 8 == 0 && (reinterpret_cast<uintptr_t>(L.residual) & 15) == 0,
                   "conv: residual needs Co %% 8 == 0 and 16-byte alignment");
    }

    CUtensorMap tmA, tmB, tmC;
    {
        const uint64_t dims[4] = {(uint64_t)L.a_C, (uint64_t)L.a_W, (uint64_t)L.a_H, (uint64_t)L.N};
        const uint64_t str[3] = {(uint64_t)L.a_ld * 2, (uint64_t)L.a_ld * 2 * L.a_W, (uint64_t)L.a_ld * 2 * L.a_W * L.a_H};
        const uint32_t box[4] = {64, (uint32_t)(p.TW * L.a_estride), (uint32_t)(p.TH * L.a_estride), 1};
        const uint32_t es[4] = {1, (uint32_t)L.a_estride, (uint32_t)L.a_estride, 1};
        if (encode_tmap_4d(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, L.a_ptr, dims, str, box, es,
                           CU_TENSOR_MAP_SWIZZLE_128B)) return -1;
    }
    if (L.b_mn) {
        const uint64_t dims[4] = {(uint64_t)L.Cout, (uint64_t)L.kw, (uint64_t)L.kh, (uint64_t)L.Kc};
        const uint64_t str[3] = {(uint64_t)L.b_cols_max * 2, (uint64_t)L.b_cols_max * 2 * L.kw,
                                 (uint64_t)L.b_cols_max * 2 * L.kw * L.kh};
        const uint32_t box[4] = {64, 1, 1, 64};
        const uint32_t es[4] = {1, 1, 1, 1};
        if (encode_tmap_4d(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, L.b_ptr, dims, str, box, es,
                           CU_TENSOR_MAP_SWIZZLE_128B)) return -1;
    } else {
        const uint64_t dims[4] = {(uint64_t)L.Kc, (uint64_t)L.kw, (uint64_t)L.kh, (uint64_t)L.Cout};
        const uint64_t str[3] = {(uint64_t)L.b_cols_max * 2, (uint64_t)L.b_cols_max * 2 * L.kw,
                                 (uint64_t)L.b_cols_max * 2 * L.kw * L.kh};
        const uint32_t box[4] = {64, 1, 1, (uint32_t)p.b_box_rows};
        const uint32_t es[4] = {1, 1, 1, 1};
        if (encode_tmap_4d(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, L.b_ptr, dims, str, box, es,
                           CU_TENSOR_MAP_SWIZZLE_128B)) return -1;
    }
    if (!p.direct) {
        const uint64_t dims[4] = {(uint64_t)L.Cout, (uint64_t)L.Wo, (uint64_t)L.Ho, (uint64_t)L.N};
        const uint64_t str[3] = {(uint64_t)L.out_ld * 2, (uint64_t)L.out_ld * 2 * L.Wo,
                                 (uint64_t)L.out_ld * 2 * L.Wo * L.Ho};
        const uint32_t box[4] = {64, (uint32_t)p.TW, (uint32_t)p.TH, 1};
        const uint32_t es[4] = {1, 1, 1, 1};
        if (encode_tmap_4d(&tmC, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, L.out, dims, str, box, es,
                           CU_TENSOR_MAP_SWIZZLE_128B)) return -1;
    } else {
        tmC = tmA;  // unused
    }
    CUtensorMap tmR = tmC;
    if (!p.direct && L.residual != nullptr) {
        const uint64_t dims[4] = {(uint64_t)L.Cout, (uint64_t)L.Wo, (uint64_t)L.Ho, (uint64_t)L.N};
        const uint64_t str[3] = {(uint64_t)L.res_ld * 2, (uint64_t)L.res_ld * 2 * L.Wo,
                                 (uint64_t)L.res_ld * 2 * L.Wo * L.Ho};
        const uint32_t box[4] = {64, (uint32_t)p.TW, (uint32_t)p.TH, 1};
        const uint32_t es[4] = {1, 1, 1, 1};
        if (encode_tmap_4d(&tmR, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, L.residual, dims, str, box, es,
                           CU_TENSOR_MAP_SWIZZLE_128B)) return -1;
    }
    static bool attr_set = false;
    if (!attr_set) {
        GS_CUDA_OK(cudaFuncSetAttribute(igemm_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        IgCfg<1>::kSmemBytes));
        GS_CUDA_OK(cudaFuncSetAttribute(igemm_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        IgCfg<2>::kSmemBytes));
        GS_CUDA_OK(cudaFuncSetAttribute(igemm_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        IgCfg<2>::kSmemBytes));
        attr_set = true;
    }
    const int units_max = num_sms() / cg;      // CTAs, or CTA pairs (148 SMs = 74 TPCs)
    GS_REQUIRE(p.n_tiles <= units_max, "conv: %d output-channel tiles exceed the SM count", p.n_tiles);
    const int m_units = (p.m_tiles + cg - 1) / cg;
    int groups = units_max / p.n_tiles;
    if (groups > m_units) groups = m_units;
    const int grid = groups * p.n_tiles * cg;
    if (cg == 1) {
        gs::launch<2>(igemm_kernel<1, false>, dim3(grid), dim3(kIgThreads), IgCfg<1>::kSmemBytes, stream, tmA, tmB, tmC, tmR, p);
    } else {
        if (p.wide)
            gs::launch_pair<2>(igemm_kernel<2, true>, dim3(grid), dim3(kIgThreads), IgCfg<2>::kSmemBytes, stream, tmA, tmB, tmC,
                               tmR, p);
        else
            gs::launch_pair<2>(igemm_kernel<2, false>, dim3(grid), dim3(kIgThreads), IgCfg<2>::kSmemBytes, stream, tmA, tmB, tmC,
                               tmR, p);
    }
    GS_LAUNCHED();
    return 0;
}

// ------------------------------------------------------------------------------------------------
// zero insertion for the data gradient of strided convolutions:
//   up[n, ho*S, wo*S, :] = dy[n, ho, wo, :], zero elsewhere.   (memory-bound, 16-byte vectors)
// ------------------------------------------------------------------------------------------------
__global__ void zero_insert_kernel(const uint4* __restrict__ dy, long long dy_ld8, uint4* __restrict__ up, int N,
                                   int Ho, int Wo, int Hu, int Wu, int S, int C8) {
    pdl_sync();
    const long long total = static_cast<long long>(N) * Hu * Wu * C8;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int c = static_cast<int>(i % C8);
        long long t = i / C8;
        const int wu = static_cast<int>(t % Wu); t /= Wu;
        const int hu = static_cast<int>(t % Hu);
        const int n = static_cast<int>(t / Hu);
        uint4 v = make_uint4(0, 0, 0, 0);
        if ((hu % S) == 0 && (wu % S) == 0) {
            const long long pix = (static_cast<long long>(n) * Ho + hu / S) * Wo + wu / S;
            v = __ldg(dy + pix * dy_ld8 + c);
        }
        up[i] = v;
    }
}

// ------------------------------------------------------------------------------------------------
// wgrad: dW[co][r][s][ci] += sum_pix dY[pix, co] * X[pix shifted by (r,s), ci]
//   M = 128 output channels (A = dY, MN-major), N <= 256 input channels (B = X, MN-major),
//   K = pixels, 64 per pipeline stage.  grid.x = co_tiles * ci_tiles * taps, grid.y = split-K.
// ------------------------------------------------------------------------------------------------
constexpr int kWgBoxBytes = 64 * 64 * 2;           // [64 pixels][64 channels] bf16
constexpr int kWgABytes = 2 * kWgBoxBytes;
constexpr int kWgThreads = 192;
constexpr int kWgMaxUnits = 148;
// CG = 2: a CTA pair (cta_group::2, M = 256 output channels): each CTA stages its own 128 output channels of dY and HALF
// of the X tile -- 32 KB instead of 48 KB of L2 -> SM traffic per 64-pixel chunk, which is what bounds the main loop.
template <int CG>
struct WgCfg {
    static constexpr int kStages = CG == 1 ? 4 : 6;
    static constexpr int kBBytes = (4 / CG) * kWgBoxBytes;
    static constexpr int kStageBytes = kWgABytes + kBBytes;
    static constexpr int kSmemBytes = 1024 + kStages * kStageBytes + 256;
};

// Work decomposition ("stream-K"): an ITEM is one (output-channel tile, input-channel tile, filter tap) accumulator of
// the weight gradient, its K loop runs over `chunks_total` 64-pixel chunks.  The linearised (item-major, chunk-minor)
// space is cut into `units` contiguous ranges of equal WEIGHTED length (weight = bytes a chunk of that item loads), one per
// CTA (or CTA pair): every SM gets the same amount of work whatever items x split-K would have quantised to (round 1:
// grid = items x splitk, e.g. 54 x 3 = 162 CTAs on 148 SMs = two waves).  A unit whose range crosses an item boundary
// flushes its accumulator (fp32 red.add into dW) and starts the next one in the other TMEM buffer.
struct WgradParams {
    int N, Ho, Wo, TH, TW, tiles_h, tiles_w, chunks_total;
    int Co, Ci, co_tiles, ci_tiles;
    int kh, kw, stride, pad, dil;
    float* dw;
    long long row_stride;  // kh*kw*Ci_max
    int Ci_max;
    int units, items;
    int start_item[kWgMaxUnits + 1];     // unit u owns [ (start_item[u], start_chunk[u]), (start_item[u+1], start_chunk[u+1]) )
    int start_chunk[kWgMaxUnits + 1];
};

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
                 : "memory");
}

template <int CG>
struct WgItem {
    int r, s, co0, ci0, ci_valid, co_valid, n_half, cb0, a_boxes, b_boxes;
    uint32_t umma_n;
    __device__ __forceinline__ WgItem(const WgradParams& p, int item, int rank) {
        // (CG 2: the two CTAs of a pair take output-channel tiles 2*cot and 2*cot + 1; a pair with an odd tile count ends on
        // a phantom tile: its dY loads are out of bounds = zero and its epilogue stores nothing)
        int t = item;
        const int tap = t % (p.kh * p.kw); t /= (p.kh * p.kw);
        const int cit = t % p.ci_tiles;
        const int cot = t / p.ci_tiles;
        r = tap / p.kw; s = tap - r * p.kw;
        co0 = (cot * CG + rank) * 128; ci0 = cit * 256;
        ci_valid = p.Ci - ci0; if (ci_valid > 256) ci_valid = 256;
        co_valid = p.Co - co0; if (co_valid > 128) co_valid = 128;
        umma_n = (ci_valid + 15) & ~15;
        n_half = static_cast<int>(umma_n) / CG;          // this CTA's share of the X tile (input channels)
        cb0 = ci0 + rank * (CG == 2 ? n_half : 0);
        a_boxes = CG == 2 ? 2 : (co_valid + 63) >> 6;   // pair: both CTAs always load 2 boxes (equal byte counts)
        b_boxes = ((CG == 2 ? n_half : ci_valid) + 63) >> 6;
    }
};

template <int CG>
__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_kernel(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmX,
             const __grid_constant__ WgradParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    constexpr int kWgStages = WgCfg<CG>::kStages;
    constexpr int kWgStageBytes = WgCfg<CG>::kStageBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kWgStages * kWgStageBytes);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + kWgStages;
    uint64_t* tfull_bar = bars + 2 * kWgStages;   // [2] accumulator buffer complete
    uint64_t* tempty_bar = tfull_bar + 2;          // [2] accumulator buffer drained
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int rank = CG == 2 ? static_cast<int>(cluster_ctarank()) : 0;
    pdl_trigger();
    const int unit = blockIdx.x / CG;
    const int it0 = p.start_item[unit], ch0 = p.start_chunk[unit];
    const int it1 = p.start_item[unit + 1], ch1 = p.start_chunk[unit + 1];
    const int tiles_hw = p.tiles_h * p.tiles_w;

    if (warp == 0 && elect_one()) {
        tma_prefetch_desc(&tmDY);
        tma_prefetch_desc(&tmX);
        for (int i = 0; i < kWgStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 4 * CG); }
        fence_mbar_init();
    }
    if (warp == 1) {
        if (CG == 2) tmem_alloc_pair(tmem_ptr, 512);
        else tmem_alloc(tmem_ptr, 512);
    }
    tc_fence_before_sync();
    if (CG == 2) cluster_sync_all();
    else __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_ptr;
    pdl_wait();     // prologue done; global memory from here on

    if (warp == 0) {
        if (elect_one()) {
            int stage = 0; uint32_t phase = 0;
            for (int item = it0; item <= it1; ++item) {
                const int c_begin = item == it0 ? ch0 : 0;
                const int c_end = item == it1 ? ch1 : p.chunks_total;
                if (c_begin >= c_end) continue;
                const WgItem<CG> w(p, item, rank);
                const uint32_t tx = CG * (w.a_boxes + w.b_boxes) * kWgBoxBytes;   // pair: the leader's barrier counts both CTAs
                for (int c = c_begin; c < c_end; ++c) {
                    const int img = c / tiles_hw;
                    const int rem = c - img * tiles_hw;
                    const int ho0 = (rem / p.tiles_w) * p.TH;
                    const int wo0 = (rem % p.tiles_w) * p.TW;
                    const int ih = ho0 * p.stride - p.pad + w.r * p.dil;
                    const int iw = wo0 * p.stride - p.pad + w.s * p.dil;
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* a_dst = smem + stage * kWgStageBytes;
                    uint8_t* b_dst = a_dst + kWgABytes;
                    if (CG == 1) {
                        mbar_arrive_expect_tx(&full_bar[stage], tx);
                        for (int j = 0; j < w.a_boxes; ++j)
                            tma_load_4d(a_dst + j * kWgBoxBytes, &tmDY, &full_bar[stage], w.co0 + j * 64, wo0, ho0, img);
                        for (int j = 0; j < w.b_boxes; ++j)
                            tma_load_4d(b_dst + j * kWgBoxBytes, &tmX, &full_bar[stage], w.ci0 + j * 64, iw, ih, img);
                    } else {
                        if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], tx);
                        for (int j = 0; j < w.a_boxes; ++j)
                            tma_load_4d_pair(a_dst + j * kWgBoxBytes, &tmDY, &full_bar[stage], w.co0 + j * 64, wo0, ho0, img);
                        for (int j = 0; j < w.b_boxes; ++j)
                            tma_load_4d_pair(b_dst + j * kWgBoxBytes, &tmX, &full_bar[stage], w.cb0 + j * 64, iw, ih, img);
                    }
                    if (++stage == kWgStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (rank == 0 && elect_one()) {
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            for (int item = it0; item <= it1; ++item) {
                const int c_begin = item == it0 ? ch0 : 0;
                const int c_end = item == it1 ? ch1 : p.chunks_total;
                if (c_begin >= c_end) continue;
                const WgItem<CG> w(p, item, rank);
                const uint32_t idesc = make_idesc_bf16(128 * CG, w.umma_n, true, true);
                if (CG == 2) mbar_wait_cluster(&tempty_bar[acc], acc_phase ^ 1);
                else mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
                tc_fence_after_sync();
                const uint32_t d_tmem = tmem_base + acc * 256;
                uint32_t accumulate = 0;
                for (int c = c_begin; c < c_end; ++c) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after_sync();
                    const uint32_t a_addr = smem_u32(smem + stage * kWgStageBytes);
                    const uint32_t b_addr = a_addr + kWgABytes;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t adesc = make_smem_desc_sw128(a_addr + k * 2048, kWgBoxBytes, 1024);
                        const uint64_t bdesc = make_smem_desc_sw128(b_addr + k * 2048, kWgBoxBytes, 1024);
                        if (CG == 2) umma_bf16_ss_pair(d_tmem, adesc, bdesc, idesc, accumulate);
                        else umma_bf16_ss(d_tmem, adesc, bdesc, idesc, accumulate);
                        accumulate = 1;
                    }
                    if (CG == 2) umma_commit_pair(&empty_bar[stage]);
                    else umma_commit(&empty_bar[stage]);
                    if (++stage == kWgStages) { stage = 0; phase ^= 1; }
                }
                if (CG == 2) umma_commit_pair(&tfull_bar[acc]);
                else umma_commit(&tfull_bar[acc]);
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
        }
    } else {
        // epilogue warps 2..5: drain the finished accumulator buffer into dW with vector fp32 reductions while the MMA
        // thread already accumulates the unit's next item in the other buffer
        const int q = warp & 3;
        int acc = 0; uint32_t acc_phase = 0;
        const bool vec_ok = (p.Ci_max % 4 == 0) && (p.Ci % 4 == 0);
        for (int item = it0; item <= it1; ++item) {
            const int c_begin = item == it0 ? ch0 : 0;
            const int c_end = item == it1 ? ch1 : p.chunks_total;
            if (c_begin >= c_end) continue;
            const WgItem<CG> w(p, item, rank);
            const int co = w.co0 + q * 32 + lane;
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after_sync();
            const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * 256;
            float* dst_row = p.dw + static_cast<long long>(co) * p.row_stride +
                             static_cast<long long>(w.r * p.kw + w.s) * p.Ci_max + w.ci0;
            const int nchunks = (w.ci_valid + 31) >> 5;
            for (int c = 0; c < nchunks; ++c) {
                uint32_t raw[32];
                tmem_ld_32x32b_x32(t_addr + c * 32, raw);
                tmem_ld_wait();
                if (c == nchunks - 1) {
                    // the whole share of this warp is in registers (or already reduced): hand the buffer back
                    tc_fence_before_sync();
                    __syncwarp();
                    if (lane == 0) ig_release_acc<CG>(&tempty_bar[acc]);
                }
                if (co < p.Co) {
                    if (vec_ok) {
#pragma unroll
                        for (int g4 = 0; g4 < 8; ++g4) {
                            const int col = c * 32 + g4 * 4;
                            if (col < w.ci_valid)
                                red_add_v4(dst_row + col, __uint_as_float(raw[g4 * 4 + 0]), __uint_as_float(raw[g4 * 4 + 1]),
                                           __uint_as_float(raw[g4 * 4 + 2]), __uint_as_float(raw[g4 * 4 + 3]));
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const int col = c * 32 + i;
                            if (col < w.ci_valid) atomicAdd(dst_row + col, __uint_as_float(raw[i]));
                        }
                    }
                }
            }
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
        }
    }
    tc_fence_before_sync();
    if (CG == 2) cluster_sync_all();
    else __syncthreads();
    if (warp == 1) {
        if (CG == 2) tmem_dealloc_pair(tmem_base, 512);
        else tmem_dealloc(tmem_base, 512);
    }
}

// Equal-weight partition of the (item, chunk) space (host side, a few hundred items at most).
static void wgrad_partition(WgradParams& p, int cg, int max_units) {
    const int taps = p.kh * p.kw;
    const int items = (int)gs_ceil_div(p.co_tiles, cg) * p.ci_tiles * taps;
    p.items = items;
    // GS_WGRAD_WEIGHT: 0 = every chunk costs the same (the main loop is bound by the per-chunk pipeline latency, measured:
    // a 24 KB chunk takes as long as a 48 KB one), 1 = cost proportional to the bytes a chunk loads
    static const int wmode = [] { const char* e = getenv("GS_WGRAD_WEIGHT"); return e ? atoi(e) : 0; }();
    auto weight = [&](int item) {
        if (wmode == 0) return 1;
        int t = item / taps;
        const int cit = t % p.ci_tiles, cot = t / p.ci_tiles;
        int ci_valid = p.Ci - cit * 256; if (ci_valid > 256) ci_valid = 256;
        if (cg == 2) return 2 + (int)gs_ceil_div(gs_round_up(ci_valid, 16) / 2, 64);
        int co_valid = p.Co - cot * 128; if (co_valid > 128) co_valid = 128;
        return (int)gs_ceil_div(co_valid, 64) + (int)gs_ceil_div(ci_valid, 64);
    };
    // at least 8 chunks (512 pixels) per unit on average, so the fp32 reductions stay a small tail
    long long total_chunks = (long long)items * p.chunks_total;
    int units = (int)(total_chunks / 8 < max_units ? total_chunks / 8 : max_units);
    if (units < 1) units = 1;
    if (units > kWgMaxUnits) units = kWgMaxUnits;
    p.units = units;
    double W = 0.0;
    for (int i = 0; i < items; ++i) W += (double)p.chunks_total * weight(i);
    p.start_item[0] = 0; p.start_chunk[0] = 0;
    int item = 0;
    double prefix = 0.0;           // weighted work before `item`
    const int snap = p.chunks_total >= 16 ? 4 : 0;   // do not leave slivers of < 4 chunks next to an item boundary
    for (int u = 1; u < units; ++u) {
        const double pos = W * u / units;
        while (item < items && prefix + (double)p.chunks_total * weight(item) <= pos) {
            prefix += (double)p.chunks_total * weight(item);
            ++item;
        }
        int it = item, ch = 0;
        if (item < items) {
            ch = (int)((pos - prefix) / weight(item) + 0.5);
            if (ch < snap) ch = 0;
            if (ch > p.chunks_total - snap) { it = item + 1; ch = 0; }
        }
        // monotone
        if (it < p.start_item[u - 1] || (it == p.start_item[u - 1] && ch < p.start_chunk[u - 1])) {
            it = p.start_item[u - 1]; ch = p.start_chunk[u - 1];
        }
        p.start_item[u] = it; p.start_chunk[u] = ch;
    }
    p.start_item[units] = items; p.start_chunk[units] = 0;
}

static int launch_wgrad(const gs_conv_geom* g, const void* x, const void* dy, float* dw, cudaStream_t stream) {
    GS_REQUIRE(g->x_ld % 8 == 0 && g->y_ld % 8 == 0, "wgrad: pitches must be multiples of 8 elements");
    GS_REQUIRE(g->stride >= 1 && g->stride <= 2, "wgrad: stride %d not supported", g->stride);
    WgradParams p{};
    p.N = g->N; p.Ho = g->Ho; p.Wo = g->Wo;
    choose_tile(g->Wo, 64, &p.TH, &p.TW);
    p.tiles_h = (int)gs_ceil_div(g->Ho, p.TH);
    p.tiles_w = (int)gs_ceil_div(g->Wo, p.TW);
    p.chunks_total = g->N * p.tiles_h * p.tiles_w;
    p.Co = g->Co; p.Ci = g->Ci;
    p.co_tiles = (int)gs_ceil_div(g->Co, 128);
    p.ci_tiles = (int)gs_ceil_div(g->Ci, 256);
    p.kh = g->kh; p.kw = g->kw; p.stride = g->stride; p.pad = g->pad; p.dil = g->dil;
    p.dw = dw; p.Ci_max = g->Ci_max;
    p.row_stride = static_cast<long long>(g->kh) * g->kw * g->Ci_max;
    static const int cg_env = [] { const char* e = getenv("GS_WGRAD_CTA_GROUP"); return e ? atoi(e) : 2; }();
    static const double pair_min_macs = [] { const char* e = getenv("GS_WGRAD_PAIR_MIN_MACS"); return e ? atof(e) : 2.5e10; }();
    // pairs pay off only for the big layers (measured round 1: >= ~2.5e10 MACs)
    const double macs = (double)g->N * g->Ho * g->Wo * g->Co * g->Ci * g->kh * g->kw;
    const int cg = (cg_env == 2 && p.co_tiles >= 2 && macs >= pair_min_macs) ? 2 : 1;
    wgrad_partition(p, cg, num_sms() / cg);

    CUtensorMap tmDY, tmX;
    {
        const uint64_t dims[4] = {(uint64_t)g->Co, (uint64_t)g->Wo, (uint64_t)g->Ho, (uint64_t)g->N};
        const uint64_t str[3] = {(uint64_t)g->y_ld * 2, (uint64_t)g->y_ld * 2 * g->Wo, (uint64_t)g->y_ld * 2 * g->Wo * g->Ho};
        const uint32_t box[4] = {64, (uint32_t)p.TW, (uint32_t)p.TH, 1};
        const uint32_t es[4] = {1, 1, 1, 1};
        if (encode_tmap_4d(&tmDY, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, dy, dims, str, box, es, CU_TENSOR_MAP_SWIZZLE_128B))
            return -1;
    }
    {
        const uint64_t dims[4] = {(uint64_t)g->Ci, (uint64_t)g->W, (uint64_t)g->H, (uint64_t)g->N};
        const uint64_t str[3] = {(uint64_t)g->x_ld * 2, (uint64_t)g->x_ld * 2 * g->W, (uint64_t)g->x_ld * 2 * g->W * g->H};
        const uint32_t box[4] = {64, (uint32_t)(p.TW * g->stride), (uint32_t)(p.TH * g->stride), 1};
        const uint32_t es[4] = {1, (uint32_t)g->stride, (uint32_t)g->stride, 1};
        if (encode_tmap_4d(&tmX, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, x, dims, str, box, es, CU_TENSOR_MAP_SWIZZLE_128B))
            return -1;
    }
    static bool attr_set = false;
    if (!attr_set) {
        GS_CUDA_OK(cudaFuncSetAttribute(wgrad_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, WgCfg<1>::kSmemBytes));
        GS_CUDA_OK(cudaFuncSetAttribute(wgrad_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, WgCfg<2>::kSmemBytes));
        attr_set = true;
    }
    if (cg == 1) {
        gs::launch<4>(wgrad_kernel<1>, dim3(p.units), dim3(kWgThreads), WgCfg<1>::kSmemBytes, stream, tmDY, tmX, p);
    } else {
        gs::launch_pair<4>(wgrad_kernel<2>, dim3(p.units * 2), dim3(kWgThreads), WgCfg<2>::kSmemBytes, stream, tmDY, tmX, p);
    }
    GS_LAUNCHED();
    return 0;
}

}  // namespace gs

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
using namespace gs;

static int check_geom(const gs_conv_geom* g) {
    GS_REQUIRE(g != nullptr, "conv: null geometry");
    GS_REQUIRE(g->N > 0 && g->H > 0 && g->W > 0 && g->Ho > 0 && g->Wo > 0, "conv: empty tensor");
    GS_REQUIRE(g->Ci > 0 && g->Co > 0 && g->Ci <= g->Ci_max && g->Co <= g->Co_max,
               "conv: active slice [%d,%d] exceeds max-width weight [%d,%d]", g->Co, g->Ci, g->Co_max, g->Ci_max);
    GS_REQUIRE(g->kh >= 1 && g->kw >= 1 && g->dil >= 1 && g->pad >= 0, "conv: bad kernel geometry");
    GS_REQUIRE(g->x_ld >= g->Ci && g->y_ld >= g->Co, "conv: pitch smaller than channel count");
    const int ho = (g->H + 2 * g->pad - g->dil * (g->kh - 1) - 1) / g->stride + 1;
    const int wo = (g->W + 2 * g->pad - g->dil * (g->kw - 1) - 1) / g->stride + 1;
    GS_REQUIRE(ho == g->Ho && wo == g->Wo, "conv: output size (%d,%d) inconsistent with geometry (expect %d,%d)",
               g->Ho, g->Wo, ho, wo);
    return 0;
}

extern "C" int gs_debug_set_trace(void* device_buffer_u64x256) {
    unsigned long long* ptr = reinterpret_cast<unsigned long long*>(device_buffer_u64x256);
    GS_CUDA_OK(cudaMemcpyToSymbol(gs::g_trace, &ptr, sizeof(ptr)));
    return 0;
}

extern "C" int gs_conv2d_fwd(const gs_conv_geom* g, const void* x, const void* w_krsc, void* y, const float* scale,
                             const float* shift, const void* residual, int32_t res_ld, int32_t flags, double* stats,
                             void* stream) {
    return gs_conv2d_fwd_syncbn(g, x, w_krsc, y, scale, shift, residual, res_ld, flags, stats, nullptr, stream);
}

extern "C" int gs_conv2d_fwd_syncbn(const gs_conv_geom* g, const void* x, const void* w_krsc, void* y, const float* scale,
                                    const float* shift, const void* residual, int32_t res_ld, int32_t flags, double* stats,
                                    const gs_sync_desc* sync, void* stream) {
    if (check_geom(g)) return -1;
    IgemmLaunch L{};
    L.a_ptr = x; L.a_C = g->Ci; L.a_ld = g->x_ld; L.a_H = g->H; L.a_W = g->W; L.a_estride = g->stride;
    L.b_ptr = w_krsc; L.b_rows_max = g->Co_max; L.b_cols_max = g->Ci_max;
    L.N = g->N; L.Ho = g->Ho; L.Wo = g->Wo; L.Cout = g->Co; L.Kc = g->Ci; L.kh = g->kh; L.kw = g->kw;
    L.in_mul = g->stride; L.base = -g->pad; L.step = g->dil;
    L.out = y; L.out_ld = g->y_ld; L.out_f32 = (flags & GS_EPI_OUT_F32) ? 1 : 0;
    L.scale = scale; L.shift = shift; L.residual = residual; L.res_ld = res_ld;
    L.relu = (flags & GS_EPI_RELU) ? 1 : 0; L.stats = stats; L.sync = sync;
    return launch_igemm(L, static_cast<cudaStream_t>(stream));
}

extern "C" int gs_conv2d_fwd_bn(const gs_conv_geom* g, const void* x, const void* w_krsc, void* y, const float* shift,
                                double* stats, double count, const float* gamma, const float* beta, float* running_mean,
                                float* running_var, float momentum, float eps, float* aff, const void* residual,
                                int32_t res_ld, int32_t relu, void* z, int32_t z_ld, void* stream) {
    if (check_geom(g)) return -1;
    GS_REQUIRE(stats != nullptr && aff != nullptr && z != nullptr && count > 0, "conv + BN: null pointer / empty count");
    GS_REQUIRE(z_ld % 8 == 0 && (reinterpret_cast<uintptr_t>(z) & 15) == 0, "conv + BN: z must be 16-byte aligned with a pitch multiple of 8");
    GS_REQUIRE(residual == nullptr || (res_ld % 8 == 0 && (reinterpret_cast<uintptr_t>(residual) & 15) == 0),
               "conv + BN: residual must be 16-byte aligned with a pitch multiple of 8");
    IgemmParams::BnTail bn{};
    bn.relu = relu ? 1 : 0;
    bn.inv_count = 1.0 / count;
    bn.unbias = count > 1.0 ? count / (count - 1.0) : 1.0;
    bn.gamma = gamma; bn.beta = beta; bn.rm = running_mean; bn.rv = running_var; bn.momentum = momentum; bn.eps = eps;
    bn.aff = aff;
    bn.res = reinterpret_cast<const __nv_bfloat16*>(residual); bn.res_ld = res_ld;
    bn.z = reinterpret_cast<__nv_bfloat16*>(z); bn.z_ld = z_ld;
    IgemmLaunch L{};
    L.a_ptr = x; L.a_C = g->Ci; L.a_ld = g->x_ld; L.a_H = g->H; L.a_W = g->W; L.a_estride = g->stride;
    L.b_ptr = w_krsc; L.b_rows_max = g->Co_max; L.b_cols_max = g->Ci_max;
    L.N = g->N; L.Ho = g->Ho; L.Wo = g->Wo; L.Cout = g->Co; L.Kc = g->Ci; L.kh = g->kh; L.kw = g->kw;
    L.in_mul = g->stride; L.base = -g->pad; L.step = g->dil;
    L.out = y; L.out_ld = g->y_ld; L.out_f32 = 0;
    L.shift = shift; L.stats = stats; L.bn = &bn;
    return launch_igemm(L, static_cast<cudaStream_t>(stream));
}

extern "C" int64_t gs_conv2d_dgrad_workspace_bytes(const gs_conv_geom* g) {
    if (g == nullptr || g->stride == 1) return 0;
    const int64_t Hu = (int64_t)(g->Ho - 1) * g->stride + 1, Wu = (int64_t)(g->Wo - 1) * g->stride + 1;
    return (int64_t)g->N * Hu * Wu * gs_round_up(g->Co, 8) * 2;
}

extern "C" int gs_conv2d_dgrad(const gs_conv_geom* g, const void* dy, const void* w_krsc, void* dx,
                               const void* residual, int32_t res_ld, void* workspace, const gs_bn_bwd_fuse* fuse,
                               void* stream_) {
    if (check_geom(g)) return -1;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    IgemmLaunch L{};
    if (g->stride == 1) {
        L.a_ptr = dy; L.a_C = g->Co; L.a_ld = g->y_ld; L.a_H = g->Ho; L.a_W = g->Wo;
    } else {
        GS_REQUIRE(workspace != nullptr, "dgrad: stride %d needs a workspace", g->stride);
        GS_REQUIRE(g->Co % 8 == 0 && g->y_ld % 8 == 0, "dgrad: strided path needs Co %% 8 == 0");
        const int Hu = (g->Ho - 1) * g->stride + 1, Wu = (g->Wo - 1) * g->stride + 1;
        const int C8 = g->Co / 8;
        const long long total = (long long)g->N * Hu * Wu * C8;
        const int blocks = (int)(gs_ceil_div(total, 256) < 148 * 16 ? gs_ceil_div(total, 256) : 148 * 16);
        gs::launch(zero_insert_kernel, dim3(blocks), dim3(256), 0, stream, reinterpret_cast<const uint4*>(dy), g->y_ld / 8,
                                                       reinterpret_cast<uint4*>(workspace), g->N, g->Ho, g->Wo, Hu, Wu,
                                                       g->stride, C8);
        GS_LAUNCHED();
        L.a_ptr = workspace; L.a_C = g->Co; L.a_ld = g->Co; L.a_H = Hu; L.a_W = Wu;
    }
    L.a_estride = 1;
    // B operand = the forward weight W[co][r][s][ci] read IN PLACE as an MN-major operand (K = co rows)
    L.b_ptr = w_krsc; L.b_rows_max = g->Co_max; L.b_cols_max = g->Ci_max; L.b_mn = 1;
    L.N = g->N; L.Ho = g->H; L.Wo = g->W; L.Cout = g->Ci; L.Kc = g->Co; L.kh = g->kh; L.kw = g->kw;
    // dx[h] = sum_r dy_up[h + pad - r*dil] * w[:, r]
    L.in_mul = 1; L.base = g->pad; L.step = -g->dil;
    L.out = dx; L.out_ld = g->x_ld; L.out_f32 = 0;
    L.scale = nullptr; L.shift = nullptr; L.residual = residual; L.res_ld = res_ld; L.relu = 0; L.stats = nullptr;
    L.fuse = fuse;
    return launch_igemm(L, stream);
}

extern "C" int gs_conv2d_wgrad(const gs_conv_geom* g, const void* x, const void* dy, float* dw_krsc, void* stream) {
    if (check_geom(g)) return -1;
    return launch_wgrad(g, x, dy, dw_krsc, static_cast<cudaStream_t>(stream));
}
