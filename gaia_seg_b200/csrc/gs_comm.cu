// gs_comm.cu -- SyncBN statistic exchange over NVLink peer memory (one kernel, no NCCL call per layer).
//
// The reference's SyncBN issues, per BN layer and direction, an NCCL all_gather / all_reduce of a few KB from the
// host (2 x ~140 layers per step: [EXT] torch.nn.SyncBatchNorm under gaiavision DynSyncBN,
// configs/_dynamic_/models/pspnet_ar50to101v2_gsync.py:20-23): latency- and launch-bound.  Here every rank owns an
// IPC-shared inbox; gs_syncbn_allreduce is ONE small kernel that
//   1. stores the local packed sums into slot[seq % NSLOT][my rank] of EVERY rank's inbox (P2P st.global over NVLink),
//   2. publishes a release flag carrying the sequence number,
//   3. spins (acquire, bounded) until the flags of all ranks for this sequence number have arrived in its own inbox,
//   4. sums the `world` contributions in rank order -> bit-identical result on every rank.
// Sequence numbers come from a device-resident counter, so the kernel is CUDA-graph safe.  A rank can be at most
// one exchange ahead of the slowest rank (it needs everybody's flag to finish), so NSLOT >= 2 slots never collide.
// Different GPUs run their kernels concurrently by construction (one process per GPU); the bounded spin turns a
// missing peer into a trapped kernel instead of a hung box.
#include "../../include/gaiaseg_b200.h"
#include "gs_host.h"

#include <stdio.h>
#include <string.h>

namespace gs {

constexpr int kCommSlots = 4;
constexpr int kCommMaxWorld = 8;
constexpr int kCommSlotDoubles = 2 * 4096;   // 2*C doubles, C <= 4096

struct PeerPtrs {
    double* p[kCommMaxWorld];
};

__host__ __device__ inline size_t comm_flag_offset_doubles(int world) {
    return static_cast<size_t>(kCommSlots) * world * kCommSlotDoubles;
}

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long gtimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__global__ void __launch_bounds__(512) syncbn_allreduce_kernel(double* __restrict__ stats, int n, PeerPtrs peers, int rank,
                                                               int world, unsigned long long* seq_dev) {
    __shared__ unsigned long long s_seq;
    const int tid = threadIdx.x;
    if (tid == 0) s_seq = *seq_dev + 1;
    __syncthreads();
    const unsigned long long seq = s_seq;
    const int slot = static_cast<int>(seq % kCommSlots);
    // 1. push the local contribution to every rank (own inbox included)
    for (int r = 0; r < world; ++r) {
        double* dst = peers.p[r] + (static_cast<size_t>(slot) * world + rank) * kCommSlotDoubles;
        for (int i = tid; i < n; i += blockDim.x) dst[i] = stats[i];
    }
    __threadfence_system();
    __syncthreads();
    // 2. publish, 3. wait
    if (tid < world) {
        unsigned long long* flags = reinterpret_cast<unsigned long long*>(peers.p[tid] + comm_flag_offset_doubles(world));
        st_release_sys(flags + slot * world + rank, seq);
        const unsigned long long* mine =
            reinterpret_cast<const unsigned long long*>(peers.p[rank] + comm_flag_offset_doubles(world)) + slot * world + tid;
        const unsigned long long t0 = gtimer();
        unsigned int spins = 0;
        while (ld_acquire_sys(mine) < seq) {
            if ((++spins & 1023u) == 0 && gtimer() - t0 > 10000000000ull) {
                printf("gaiaseg_b200: SyncBN peer exchange timed out (rank %d waiting for rank %d, seq %llu)\n", rank, tid, seq);
                __trap();
            }
        }
    }
    __syncthreads();
    // 4. reduce in rank order
    const double* inbox = peers.p[rank] + static_cast<size_t>(slot) * world * kCommSlotDoubles;
    for (int i = tid; i < n; i += blockDim.x) {
        double s = 0.0;
        for (int r = 0; r < world; ++r) s += __ldcg(inbox + static_cast<size_t>(r) * kCommSlotDoubles + i);
        stats[i] = s;
    }
    if (tid == 0) *seq_dev = seq;
}

}  // namespace gs

using namespace gs;

extern "C" int64_t gs_comm_inbox_bytes(int32_t world) {
    return static_cast<int64_t>(comm_flag_offset_doubles(world)) * 8 + static_cast<int64_t>(kCommSlots) * world * 8;
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
                                   void* seq_dev, void* stream) {
    GS_REQUIRE(stats && peer_inboxes && seq_dev, "syncbn_allreduce: null pointer");
    GS_REQUIRE(world >= 1 && world <= kCommMaxWorld && rank >= 0 && rank < world, "syncbn_allreduce: bad rank %d / world %d",
               rank, world);
    GS_REQUIRE(n > 0 && n <= kCommSlotDoubles, "syncbn_allreduce: %d values exceed the slot size %d", n, kCommSlotDoubles);
    PeerPtrs pp{};
    for (int r = 0; r < world; ++r) {
        GS_REQUIRE(peer_inboxes[r] != nullptr, "syncbn_allreduce: inbox of rank %d is not mapped", r);
        pp.p[r] = reinterpret_cast<double*>(const_cast<void*>(peer_inboxes[r]));
    }
    syncbn_allreduce_kernel<<<1, 512, 0, static_cast<cudaStream_t>(stream)>>>(
        stats, n, pp, rank, world, reinterpret_cast<unsigned long long*>(seq_dev));
    GS_LAUNCHED();
    return 0;
}
