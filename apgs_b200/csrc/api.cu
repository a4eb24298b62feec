// Library-level entry points: version, thread-local error string, device properties.
#include "common.cuh"

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

}  // namespace apgs

extern "C" int apgs_version(void) { return APGS_ABI_VERSION; }
extern "C" const char* apgs_last_error(void) { return apgs::err_buf(); }
