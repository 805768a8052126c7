// gs_host.h -- host-side helpers shared by the C-ABI translation units.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

namespace gs {

void set_error(const char* fmt, ...);
extern std::atomic<int64_t> g_launches;

// Encodes a rank-4 tiled tensor map (dims/strides listed innermost first; strides for dims 1..3 in
// bytes).  Returns 0 / -1 (error message set).  The driver entry point is resolved at run time with
// cudaGetDriverEntryPoint so the library loads on a machine without libcuda (CPU-only CI).
int encode_tmap_4d(CUtensorMap* m, CUtensorMapDataType dt, const void* base, const uint64_t dims[4],
                   const uint64_t strides_bytes[3], const uint32_t box[4], const uint32_t estr[4],
                   CUtensorMapSwizzle swz);

int num_sms();

// spin budget of the peer-memory kernels in ns (GS_COMM_TIMEOUT_S, default 600 s; gs_comm.cu)
unsigned long long comm_timeout_ns();

// Programmatic dependent launch (PDL).  Every kernel of this library starts with pdl_sync() (or, for the tensor-core
// kernels, pdl_trigger() at the top and pdl_wait() after the barrier / TMEM / tensor-map prologue), and is launched with
// cudaLaunchAttributeProgrammaticStreamSerialization: the launch latency, block dispatch and prologue of kernel i+1 overlap
// the tail of kernel i on the same stream (~1050 conv + ~1100 BN launches of 10-40 us per sandwich cycle), while
// griddepcontrol.wait still orders every global-memory access after the COMPLETION of all earlier kernels -- the
// semantics of a plain stream are unchanged.  Legal inside captured CUDA graphs (programmatic edges).  GS_PDL=0 disables it.
// GS_PDL is a bit mask: 1 = memory-bound kernels (except the three DynBN training kernels, which have their own bits:
// 16 = BN apply forward, 32 = BN-backward reduce, 64 = BN-backward apply), 2 = igemm (conv fwd / dgrad), 4 = wgrad, 8 = the one-block SyncBN
// exchange kernel (several ranks; its block is dispatched while the producer of the sums still runs: 57.3 -> 55.7 ms per
// cycle at N = 2).  Default 14, MEASURED on the
// sandwich cycle (profiles/r02_pdl_sweep.md): early-scheduled blocks of the memory-bound kernels take the scheduling gaps
// the low-priority side-stream wgrad kernels live on (52.3 -> 55.4 ms), the tensor-core kernels gain (52.3 -> 50.7 ms).
bool pdl_enabled(int kind_bit = 1);

template <int KIND = 1, typename... KArgs, typename... Args>
inline void launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl_enabled(KIND) ? 1 : 0;
    (void)cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);   // errors surface through GS_LAUNCHED()
}

// cluster-of-2 variant (CTA pairs of the tensor-core kernels)
template <int KIND = 1, typename... KArgs, typename... Args>
inline void launch_pair(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl_enabled(KIND) ? 2 : 1;
    (void)cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace gs

#ifdef __CUDACC__
namespace gs {
// allow the next kernel of the stream to be scheduled (its blocks still wait for OUR completion in pdl_wait)
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// block until every earlier kernel of the stream has completed and its writes are visible
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_sync() { pdl_trigger(); pdl_wait(); }
}  // namespace gs
#endif

#define GS_REQUIRE(cond, ...)          \
    do {                               \
        if (!(cond)) {                 \
            gs::set_error(__VA_ARGS__); \
            return -1;                 \
        }                              \
    } while (0)

#define GS_CUDA_OK(expr)                                                                       \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            gs::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return -2;                                                                         \
        }                                                                                      \
    } while (0)

// after every kernel launch: count it and surface launch-configuration errors
#define GS_LAUNCHED()                       \
    do {                                    \
        gs::g_launches.fetch_add(1);        \
        GS_CUDA_OK(cudaGetLastError());     \
    } while (0)

static inline int64_t gs_ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int gs_round_up(int a, int b) { return (a + b - 1) / b * b; }
