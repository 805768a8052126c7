// gs_host.cu -- error state, launch counter, device check, tensor-map encoding.
#include "gs_host.h"

#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

#include <mutex>

#include "../../include/gaiaseg_b200.h"

namespace gs {

static thread_local char t_err[512] = "";
std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_err, sizeof(t_err), fmt, ap);
    va_end(ap);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess) {
            fn = reinterpret_cast<EncodeTiledFn>(p);
        }
    });
    return fn;
}

int encode_tmap_4d(CUtensorMap* m, CUtensorMapDataType dt, const void* base, const uint64_t dims[4],
                   const uint64_t strides_bytes[3], const uint32_t box[4], const uint32_t estr[4],
                   CUtensorMapSwizzle swz) {
    EncodeTiledFn fn = get_encode();
    GS_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled not available (no CUDA driver?)");
    CUresult r = fn(m, dt, 4, const_cast<void*>(base), dims, strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    GS_REQUIRE(r == CUDA_SUCCESS,
               "cuTensorMapEncodeTiled failed (%d): base=%p dims=[%llu,%llu,%llu,%llu] strides=[%llu,%llu,%llu] "
               "box=[%u,%u,%u,%u] estr=[%u,%u,%u,%u]",
               (int)r, base, (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2],
               (unsigned long long)dims[3], (unsigned long long)strides_bytes[0],
               (unsigned long long)strides_bytes[1], (unsigned long long)strides_bytes[2], box[0], box[1], box[2],
               box[3], estr[0], estr[1], estr[2], estr[3]);
    return 0;
}

bool pdl_enabled(int kind_bit) {
    static const int mask = [] { const char* e = getenv("GS_PDL"); return e == nullptr ? 14 : atoi(e); }();
    return (mask & kind_bit) != 0;
}

int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 148;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    }
    return n;
}

}  // namespace gs

extern "C" {

int gs_version(void) { return GS_ABI_VERSION; }

const char* gs_last_error(void) { return gs::t_err; }

int gs_device_check(void) {
    int dev = 0;
    GS_CUDA_OK(cudaGetDevice(&dev));
    int major = 0, minor = 0;
    GS_CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    GS_CUDA_OK(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
    GS_REQUIRE(major == 10, "libgaiaseg_b200 is built for sm_100a only; device %d is sm_%d%d", dev, major, minor);
    return 0;
}

int64_t gs_launch_count(void) { return gs::g_launches.load(); }
void gs_reset_launch_count(void) { gs::g_launches.store(0); }

}  // extern "C"
