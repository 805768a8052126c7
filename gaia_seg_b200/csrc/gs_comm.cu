// gs_comm.cu -- SyncBN statistic exchange over NVLink peer memory (one kernel, no NCCL call per layer).
//
// The reference's SyncBN issues, per BN layer and direction, an NCCL all_gather / all_reduce of a few KB from the
// host (2 x ~140 layers per step: [EXT] torch.nn.SyncBatchNorm under gaiavision DynSyncBN,
// configs/_dynamic_/models/pspnet_ar50to101v2_gsync.py:20-23): latency- and launch-bound.  Here every rank owns an
// IPC-shared inbox; gs_syncbn_allreduce is ONE small kernel with a flag-in-data ("low latency") protocol:
//   1. every fp64 value is split into two 32-bit halves, each stored together with the 32-bit sequence tag of this
//      exchange as ONE 8-byte word (single-copy atomic) straight into slot[seq % NSLOT][my rank] of every PEER's inbox
//      (P2P st.global over NVLink) -- no memory fence, no separate flag, so the cost is ONE one-way NVLink latency;
//   2. the kernel then polls its own inbox until both words of a value carry the current tag and sums the `world`
//      contributions in rank order (its own from registers) -> bit-identical result on every rank.
// Sequence numbers come from a device-resident counter, so the kernel is CUDA-graph safe.  A rank can be at most
// one exchange ahead of the slowest rank (it needs everybody's data to finish), so NSLOT = 4 slots never collide,
// and a slot's stale words carry an older tag.  Different GPUs run their kernels concurrently by construction (one
// process per GPU); the bounded spin turns a missing peer into a trapped kernel instead of a hung box.
// Optionally the same kernel first accumulates the BN parameter gradients from the LOCAL sums (dbeta += sum g,
// dgamma += sum g*xhat) -- they are averaged later by the gradient all-reduce like every other parameter gradient.
#include "../../include/gaiaseg_b200.h"
#include "gs_host.h"

#include <stdio.h>
#include <string.h>

namespace gs {

constexpr int kCommSlots = 4;
constexpr int kCommMaxWorld = 8;
constexpr int kCommSlotDoubles = 2 * 4096;   // 2*C doubles, C <= 4096
constexpr int kCommThreads = 1024;
constexpr int kCommPerThread = kCommSlotDoubles / kCommThreads;

struct PeerPtrs {
    ulonglong2* p[kCommMaxWorld];
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

__global__ void __launch_bounds__(kCommThreads) syncbn_allreduce_kernel(double* __restrict__ stats, int n, PeerPtrs peers,
                                                                        int rank, int world, unsigned long long* seq_dev,
                                                                        float* __restrict__ dgamma, float* __restrict__ dbeta) {
    const int tid = threadIdx.x;
    const unsigned long long seq = *seq_dev + 1;     // every thread reads the counter; thread 0 bumps it at the very end
    const unsigned long long tag = (seq & 0xFFFFFFFFull) << 32;
    const int slot = static_cast<int>(seq % kCommSlots);
    double mine[kCommPerThread];
    // 1. push the local contribution to every peer
#pragma unroll
    for (int k = 0; k < kCommPerThread; ++k) {
        const int i = tid + k * kCommThreads;
        if (i < n) {
            const double v = stats[i];
            mine[k] = v;
            const unsigned long long b = static_cast<unsigned long long>(__double_as_longlong(v));
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
    // BN parameter gradients from the LOCAL sums (n = 2C: [sum g | sum g*xhat])
    if (dbeta != nullptr || dgamma != nullptr) {
        const int C = n >> 1;
#pragma unroll
        for (int k = 0; k < kCommPerThread; ++k) {
            const int i = tid + k * kCommThreads;
            if (i < n) {
                if (i < C) { if (dbeta) dbeta[i] += static_cast<float>(mine[k]); }
                else if (dgamma) dgamma[i - C] += static_cast<float>(mine[k]);
            }
        }
    }
    // 2. poll the own inbox, reduce in rank order
    const ulonglong2* inbox = peers.p[rank] + static_cast<size_t>(slot) * world * kCommSlotDoubles;
    const unsigned long long t0 = gtimer();
#pragma unroll
    for (int k = 0; k < kCommPerThread; ++k) {
        const int i = tid + k * kCommThreads;
        if (i < n) {
            double s = 0.0;
            for (int r = 0; r < world; ++r) {
                if (r == rank) { s += mine[k]; continue; }
                const ulonglong2* src = inbox + static_cast<size_t>(r) * kCommSlotDoubles + i;
                ulonglong2 w = ld_v2_sys(src);
                unsigned int spins = 0;
                while ((w.x & 0xFFFFFFFF00000000ull) != tag || (w.y & 0xFFFFFFFF00000000ull) != tag) {
                    if ((++spins & 1023u) == 0 && gtimer() - t0 > 10000000000ull) {
                        printf("gaiaseg_b200: SyncBN peer exchange timed out (rank %d waiting for rank %d, seq %llu)\n", rank,
                               r, seq);
                        __trap();
                    }
                    w = ld_v2_sys(src);
                }
                s += __longlong_as_double(static_cast<long long>((w.x & 0xFFFFFFFFull) | (w.y << 32)));
            }
            stats[i] = s;
        }
    }
    __syncthreads();      // every thread has read *seq_dev long before, but keep the bump strictly last
    if (tid == 0) *seq_dev = seq;
}

}  // namespace gs

using namespace gs;

extern "C" int64_t gs_comm_inbox_bytes(int32_t world) {
    return static_cast<int64_t>(kCommSlots) * world * kCommSlotDoubles * static_cast<int64_t>(sizeof(ulonglong2));
}

extern "C" int gs_ipc_alloc(int64_t bytes, void** dev_ptr, void* handle_out_64) {
    GS_REQUIRE(bytes > 0 && dev_ptr && handle_out_64, "ipc_alloc: bad arguments");
    void* p = nullptr;
    GS_CUDA_OK(cudaMalloc(&p, static_cast<size_t>(bytes)));
    GS_CUDA_OK(cudaMemset(p, 0, static_cast<size_t>(bytes)));
    cudaIpcMemHandle_t h;
    GS_CUDA_OK(cudaIpcGetMemHandle(&h, p));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    memcpy(handle_out_64, &h, 64);
    *dev_ptr = p;
    GS_CUDA_OK(cudaDeviceSynchronize());
    return 0;
}

extern "C" int gs_ipc_open(const void* handle_64, void** dev_ptr) {
    GS_REQUIRE(handle_64 && dev_ptr, "ipc_open: bad arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle_64, 64);
    void* p = nullptr;
    GS_CUDA_OK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    *dev_ptr = p;
    return 0;
}

extern "C" int gs_ipc_close(void* dev_ptr) {
    GS_CUDA_OK(cudaIpcCloseMemHandle(dev_ptr));
    return 0;
}

extern "C" int gs_ipc_free(void* dev_ptr) {
    GS_CUDA_OK(cudaFree(dev_ptr));
    return 0;
}

extern "C" int gs_syncbn_allreduce(double* stats, int32_t n, const void* const* peer_inboxes, int32_t rank, int32_t world,
                                   void* seq_dev, float* dgamma, float* dbeta, void* stream) {
    GS_REQUIRE(stats && peer_inboxes && seq_dev, "syncbn_allreduce: null pointer");
    GS_REQUIRE(world >= 1 && world <= kCommMaxWorld && rank >= 0 && rank < world, "syncbn_allreduce: bad rank %d / world %d",
               rank, world);
    GS_REQUIRE(n > 0 && n <= kCommSlotDoubles, "syncbn_allreduce: %d values exceed the slot size %d", n, kCommSlotDoubles);
    GS_REQUIRE((dgamma == nullptr && dbeta == nullptr) || n % 2 == 0, "syncbn_allreduce: parameter gradients need n = 2C");
    PeerPtrs pp{};
    for (int r = 0; r < world; ++r) {
        GS_REQUIRE(peer_inboxes[r] != nullptr, "syncbn_allreduce: inbox of rank %d is not mapped", r);
        pp.p[r] = reinterpret_cast<ulonglong2*>(const_cast<void*>(peer_inboxes[r]));
    }
    int threads = ((n + 31) / 32) * 32;
    if (threads > kCommThreads) threads = kCommThreads;
    syncbn_allreduce_kernel<<<1, kCommThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        stats, n, pp, rank, world, reinterpret_cast<unsigned long long*>(seq_dev), dgamma, dbeta);
    (void)threads;
    GS_LAUNCHED();
    return 0;
}
