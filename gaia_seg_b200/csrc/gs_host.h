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

}  // namespace gs

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
