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
#include "gs_comm.cuh"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

namespace gs {

constexpr int kCommThreads = 1024;

__global__ void __launch_bounds__(kCommThreads) syncbn_allreduce_kernel(double* __restrict__ stats, int n, PeerPtrs peers,
                                                                        int rank, int world, unsigned long long* seq_dev,
                                                                        float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                                        unsigned long long timeout_ns) {
    pdl_sync();
    syncbn_exchange_block(stats, n, peers, rank, world, seq_dev, dgamma, dbeta, timeout_ns, threadIdx.x, kCommThreads);
}

// ------------------------------------------------------------------------------------------------
// gradient all-reduce over peer memory (replaces the NCCL all-reduce of the flat gradient buffer)
// ------------------------------------------------------------------------------------------------
// Every rank's flat fp32 gradient buffer is IPC-mapped by all peers.  Two-shot all-reduce of a range in three
// stream-ordered launches (all capturable in a CUDA graph, sequence counter in device memory):
//   1. barrier "ready":  my gradients of the range are complete (stream order) -> flag to every peer, wait for theirs;
//   2. reduce + push:    rank r owns the r-th shard of the range: s = sum over ranks (fixed order, P2P 16-byte loads)
//                        and stores s into EVERY rank's buffer (P2P stores) -- each element is touched by one rank only;
//   3. barrier "done":   all shards have landed everywhere (and nobody reads my buffer any more).
// Traffic per rank: (world-1)/world of the range read and written over NVLink, i.e. what a ring all-reduce moves, but
// with every link busy at once (NVSwitch) and no host-side NCCL call.
struct GradPtrs {
    float* g[kCommMaxWorld];
};
struct FlagPtrs {
    unsigned long long* f[kCommMaxWorld];   // per rank: [2 flag sets][kCommMaxWorld]
};

__global__ void peer_barrier_kernel(FlagPtrs flags, int rank, int world, unsigned long long* seq_dev, int set, int bump,
                                    unsigned long long timeout_ns) {
    pdl_sync();
    const int tid = threadIdx.x;
    const unsigned long long seq = *seq_dev + (bump ? 1ull : 0ull);
    __syncthreads();
    if (tid < world && tid != rank) {
        __threadfence_system();
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flags.f[tid] + set * kCommMaxWorld + rank), "l"(seq) : "memory");
        const unsigned long long* mine = flags.f[rank] + set * kCommMaxWorld + tid;
        const unsigned long long t0 = gtimer();
        unsigned int spins = 0;
        unsigned long long v;
        do {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(mine) : "memory");
            if ((++spins & 1023u) == 0 && gtimer() - t0 > timeout_ns) {
                printf("gaiaseg_b200: gradient all-reduce barrier timed out (rank %d waiting for rank %d, seq %llu)\n", rank,
                       tid, seq);
                __trap();
            }
        } while (v < seq);
    }
    __syncthreads();
    if (tid == 0 && bump) *seq_dev = seq;
}

__device__ __forceinline__ float4 ld_f4_sys(const float4* p) {
    float4 v;
    asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_f4_sys(float4* p, const float4& v) {
    asm volatile("st.volatile.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__global__ void __launch_bounds__(256) grad_reduce_push_kernel(GradPtrs ptrs, long long lo4, long long hi4, int world) {
    pdl_sync();
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = lo4 + blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < hi4; i += stride) {
        float4 v[kCommMaxWorld];
#pragma unroll
        for (int r = 0; r < kCommMaxWorld; ++r)
            if (r < world) v[r] = ld_f4_sys(reinterpret_cast<const float4*>(ptrs.g[r]) + i);
        float4 s = v[0];
#pragma unroll
        for (int r = 1; r < kCommMaxWorld; ++r)
            if (r < world) { s.x += v[r].x; s.y += v[r].y; s.z += v[r].z; s.w += v[r].w; }
#pragma unroll
        for (int r = 0; r < kCommMaxWorld; ++r)
            if (r < world) st_f4_sys(reinterpret_cast<float4*>(ptrs.g[r]) + i, s);
    }
    __threadfence_system();
}

// Spin budget of the peer kernels.  A rank may legitimately wait for a peer that is busy with rank-local host work
// (checkpoint write, dataset.evaluate, a stalled DataLoader): the default matches NCCL's watchdog scale (10 minutes),
// GS_COMM_TIMEOUT_S overrides it (tests and benchmarks use a short budget so that a bug traps instead of hanging a box).
unsigned long long comm_timeout_ns() {
    static const unsigned long long v = [] {
        const char* e = getenv("GS_COMM_TIMEOUT_S");
        double s = e ? atof(e) : 600.0;
        if (!(s > 0.0)) s = 600.0;
        return static_cast<unsigned long long>(s * 1e9);
    }();
    return v;
}

}  // namespace gs

using namespace gs;

extern "C" int64_t gs_comm_flags_bytes(void) { return 2 * kCommMaxWorld * 8; }

extern "C" int gs_grad_allreduce(const void* const* peer_grads, int64_t offset, int64_t count, const void* const* peer_flags,
                                 int32_t rank, int32_t world, void* seq_dev, void* stream) {
    GS_REQUIRE(peer_grads && peer_flags && seq_dev, "grad_allreduce: null pointer");
    GS_REQUIRE(world >= 1 && world <= kCommMaxWorld && rank >= 0 && rank < world, "grad_allreduce: bad rank %d / world %d", rank,
               world);
    GS_REQUIRE(offset >= 0 && count >= 0 && offset % 4 == 0 && count % 4 == 0,
               "grad_allreduce: offset / count must be multiples of 4 elements");
    if (count == 0 || world == 1) return 0;
    GradPtrs gp{};
    FlagPtrs fp{};
    for (int r = 0; r < world; ++r) {
        GS_REQUIRE(peer_grads[r] && peer_flags[r], "grad_allreduce: buffers of rank %d are not mapped", r);
        GS_REQUIRE((reinterpret_cast<uintptr_t>(peer_grads[r]) & 15) == 0, "grad_allreduce: gradient buffer not 16-byte aligned");
        gp.g[r] = reinterpret_cast<float*>(const_cast<void*>(peer_grads[r]));
        fp.f[r] = reinterpret_cast<unsigned long long*>(const_cast<void*>(peer_flags[r]));
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    unsigned long long* seq = reinterpret_cast<unsigned long long*>(seq_dev);
    const long long n4 = count / 4, o4 = offset / 4;
    const long long per = (n4 + world - 1) / world;
    long long lo4 = o4 + per * rank, hi4 = lo4 + per;
    if (lo4 > o4 + n4) lo4 = o4 + n4;
    if (hi4 > o4 + n4) hi4 = o4 + n4;
    gs::launch(peer_barrier_kernel, dim3(1), dim3(32), 0, st, fp, rank, world, seq, 0, 1, comm_timeout_ns());
    if (hi4 > lo4) {
        long long blocks = (hi4 - lo4 + 256 * 4 - 1) / (256 * 4);
        const long long cap = static_cast<long long>(num_sms()) * 4;
        if (blocks > cap) blocks = cap;
        gs::launch(grad_reduce_push_kernel, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, st, gp, lo4, hi4, world);
    }
    gs::launch(peer_barrier_kernel, dim3(1), dim3(32), 0, st, fp, rank, world, seq, 1, 0, comm_timeout_ns());
    GS_LAUNCHED();
    return 0;
}

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
    // PDL kind 8: the single exchange block is dispatched while the producer of `stats` still runs (it waits in
    // griddepcontrol.wait), so the exchange starts the moment the producer completes instead of one launch latency later
    gs::launch<8>(syncbn_allreduce_kernel, dim3(1), dim3(kCommThreads), 0, static_cast<cudaStream_t>(stream), 
        stats, n, pp, rank, world, reinterpret_cast<unsigned long long*>(seq_dev), dgamma, dbeta, comm_timeout_ns());
    (void)threads;
    GS_LAUNCHED();
    return 0;
}
