// Shared host/device helpers for the apgs_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <math.h>
#include "../../include/apgs_b200.h"

namespace apgs {

// ---- thread-local error string behind apgs_last_error() ------------------------------------
char* err_buf();
inline int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(err_buf(), 512, fmt, ap);
    va_end(ap);
    return code;
}
inline int cuda_fail(cudaError_t e, const char* what) {
    snprintf(err_buf(), 512, "%s: %s", what, cudaGetErrorString(e));
    return (int)e;
}
#define APGS_CUDA_OK(expr, what)                                  \
    do {                                                          \
        cudaError_t _e = (expr);                                  \
        if (_e != cudaSuccess) return ::apgs::cuda_fail(_e, what); \
    } while (0)
#define APGS_LAUNCH_OK(what) APGS_CUDA_OK(cudaGetLastError(), what)

// B200: 148 SMs; cached per device at first use.
int sm_count();

// Opt a kernel into > 48 KB of dynamic shared memory.  The attribute is set once per (function, device, size high-water
// mark) and remembered, so steady-state calls -- and CUDA-graph capture -- issue nothing but the launch itself.
cudaError_t ensure_dynamic_smem(const void* func, size_t bytes);

__host__ __device__ inline bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) == 0; }

// ---- exact division by an invariant for numerators < 2^31 (Granlund-Montgomery) -------------
struct FastDiv {
    unsigned long long m;   // multiplier
    unsigned int sh;        // total shift
    unsigned int d;
};
inline FastDiv make_fastdiv(unsigned int d) {
    FastDiv f;
    f.d = d;
    unsigned int l = 0;
    while ((1ull << l) < d) ++l;
    f.sh = 31 + l;
    f.m = ((1ull << f.sh) / d) + 1;
    return f;
}
__device__ __forceinline__ unsigned int fastdiv(unsigned int n, const FastDiv& f) {
    return (unsigned int)(((unsigned long long)n * f.m) >> f.sh);
}

// ---- warp helpers --------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

}  // namespace apgs
