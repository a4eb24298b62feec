// Library-level entry points: version, thread-local error string, device properties.
#include "common.cuh"
#include <mutex>

namespace apgs {

char* err_buf() {
    static thread_local char buf[512] = {0};
    return buf;
}

int sm_count() {
    static thread_local int cached_dev = -1;
    static thread_local int cached = 148;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) cached = n;
        cached_dev = dev;
    }
    return cached;
}

cudaError_t ensure_dynamic_smem(const void* func, size_t bytes) {
    if (bytes <= 48 * 1024) return cudaSuccess;
    struct Entry { const void* f; int dev; size_t bytes; };
    static Entry table[256];
    static int n = 0;
    static std::mutex mu;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lock(mu);
    for (int i = 0; i < n; ++i)
        if (table[i].f == func && table[i].dev == dev) {
            if (table[i].bytes >= bytes) return cudaSuccess;
            e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
            if (e == cudaSuccess) table[i].bytes = bytes;
            return e;
        }
    e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess && n < 256) table[n++] = {func, dev, bytes};
    return e;
}

}  // namespace apgs

extern "C" int apgs_version(void) { return APGS_ABI_VERSION; }
extern "C" const char* apgs_last_error(void) { return apgs::err_buf(); }
