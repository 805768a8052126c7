// gs_comm.cuh -- device side of the SyncBN statistic exchange over NVLink peer memory, shared by the stand-alone
// kernel (gs_comm.cu: gs_syncbn_allreduce) and the DynBN kernels that fold the exchange into their own launch
// (gs_norm.cu: block 0 exchanges while the other blocks wait on a flag -- no extra launch on the critical chain).
//
// Protocol ("flag in data"): every fp64 value travels as two 8-byte words {32 data bits | 32-bit sequence tag} stored
// straight into slot[seq % kCommSlots][my rank] of every PEER's inbox (P2P st.global over NVLink, no fence, no separate
// flag: one one-way NVLink latency); the receiver polls its own inbox until both words carry the tag of this exchange and
// sums the `world` contributions in rank order (bit-identical result on every rank).  The sequence number lives in device
// memory (CUDA-graph safe).  A rank can be at most one exchange ahead of the slowest rank, so 4 slots never collide.
#pragma once
#include <stdint.h>
#include <stdio.h>

namespace gs {

constexpr int kCommSlots = 4;
constexpr int kCommMaxWorld = 8;
constexpr int kCommSlotDoubles = 2 * 4096;   // 2*C doubles, C <= 4096

struct PeerPtrs {
    ulonglong2* p[kCommMaxWorld];
};

// what a kernel needs to run the exchange itself (world <= 1: no exchange)
struct SyncArgs {
    PeerPtrs peers;
    int rank, world;
    unsigned long long* seq_dev;
    unsigned long long timeout_ns;
    // 0: the kernel runs the whole exchange (push + poll); 1: PUSH only (a producer kernel: its last block sends the final
    // local sums and leaves the sequence counter alone); 2: POLL only (the consumer of a phase-1 producer).  Splitting the
    // exchange puts the NVLink flight time behind the producer's tail and the consumer's launch instead of on the chain.
    int phase;
};

__device__ __forceinline__ void st_u64_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ ulonglong2 ld_v2_sys(const ulonglong2* p) {
    ulonglong2 v;
    asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long gtimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ unsigned long long ld_acquire_gpu(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// PUSH half, by all `nthreads` threads of the LAST block of a producer kernel (every other block's atomics on `stats` are
// fenced before that block took its ticket): the final local sums, read through L2, go to every peer's inbox under the
// tag of the NEXT exchange.  The sequence counter is bumped by the consumer's poll.
__device__ __forceinline__ void syncbn_push_block(const double* __restrict__ stats, int n, const PeerPtrs& peers, int rank,
                                                  int world, const unsigned long long* seq_dev, int tid, int nthreads) {
    const unsigned long long seq = __ldcg(seq_dev) + 1;
    const unsigned long long tag = (seq & 0xFFFFFFFFull) << 32;
    const int slot = static_cast<int>(seq % kCommSlots);
    for (int i = tid; i < n; i += nthreads) {
        const unsigned long long b = static_cast<unsigned long long>(__double_as_longlong(__ldcg(stats + i)));
        const unsigned long long w0 = (b & 0xFFFFFFFFull) | tag, w1 = (b >> 32) | tag;
        for (int r = 0; r < world; ++r) {
            if (r == rank) continue;
            unsigned long long* dst = reinterpret_cast<unsigned long long*>(
                peers.p[r] + (static_cast<size_t>(slot) * world + rank) * kCommSlotDoubles + i);
            st_u64_sys(dst, w0);
            st_u64_sys(dst + 1, w1);
        }
    }
}

// All `nthreads` threads of ONE block call this (tid = 0 .. nthreads-1).  stats[0..n) holds the LOCAL sums on entry and
// the sums over all ranks on return.  dgamma / dbeta (may be NULL; then n = 2C: [sum g | sum g*xhat]) are incremented by
// the LOCAL sums first -- the BN parameter gradients, averaged later by the gradient all-reduce.
// Ends with a __syncthreads(); thread 0 bumps the sequence counter last.
__device__ __forceinline__ void syncbn_exchange_block(double* __restrict__ stats, int n, const PeerPtrs& peers, int rank,
                                                      int world, unsigned long long* seq_dev, float* __restrict__ dgamma,
                                                      float* __restrict__ dbeta, unsigned long long timeout_ns, int tid,
                                                      int nthreads, bool push = true) {
    const unsigned long long seq = *seq_dev + 1;     // every thread reads the counter; thread 0 bumps it at the very end
    const unsigned long long tag = (seq & 0xFFFFFFFFull) << 32;
    const int slot = static_cast<int>(seq % kCommSlots);
    const int C = n >> 1;
    // 1. push the local contribution to every peer (+ parameter gradients from the local sums)
    for (int i = tid; i < n; i += nthreads) {
        const double v = stats[i];
        const unsigned long long b = static_cast<unsigned long long>(__double_as_longlong(v));
        const unsigned long long w0 = (b & 0xFFFFFFFFull) | tag, w1 = (b >> 32) | tag;
        for (int r = 0; push && r < world; ++r) {      // (push == false: the producer kernel has sent them already)
            if (r == rank) continue;
            unsigned long long* dst = reinterpret_cast<unsigned long long*>(
                peers.p[r] + (static_cast<size_t>(slot) * world + rank) * kCommSlotDoubles + i);
            st_u64_sys(dst, w0);
            st_u64_sys(dst + 1, w1);
        }
        if (i < C) { if (dbeta) dbeta[i] += static_cast<float>(v); }
        else if (dgamma) dgamma[i - C] += static_cast<float>(v);
    }
    // 2. poll the own inbox, reduce in rank order (thread i re-reads ITS local values: nobody else wrote them)
    const ulonglong2* inbox = peers.p[rank] + static_cast<size_t>(slot) * world * kCommSlotDoubles;
    const unsigned long long t0 = gtimer();
    for (int i = tid; i < n; i += nthreads) {
        const double mine = stats[i];
        double s = 0.0;
        for (int r = 0; r < world; ++r) {
            if (r == rank) { s += mine; continue; }
            const ulonglong2* src = inbox + static_cast<size_t>(r) * kCommSlotDoubles + i;
            ulonglong2 w = ld_v2_sys(src);
            unsigned int spins = 0;
            while ((w.x & 0xFFFFFFFF00000000ull) != tag || (w.y & 0xFFFFFFFF00000000ull) != tag) {
                if ((++spins & 1023u) == 0 && gtimer() - t0 > timeout_ns) {
                    printf("gaiaseg_b200: SyncBN peer exchange timed out (rank %d waiting for rank %d, seq %llu)\n", rank, r,
                           seq);
                    __trap();
                }
                w = ld_v2_sys(src);
            }
            s += __longlong_as_double(static_cast<long long>((w.x & 0xFFFFFFFFull) | (w.y << 32)));
        }
        stats[i] = s;
    }
    __syncthreads();      // every thread has read *seq_dev long before, but keep the bump strictly last
    if (tid == 0) *seq_dev = seq;
}

}  // namespace gs
