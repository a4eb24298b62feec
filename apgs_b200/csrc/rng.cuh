// Counter-based random numbers for draws made inside the kernels: Philox4x32-10 (Salmon et al., SC'11; the generator
// behind torch's CUDA RNG), keyed by a 64-bit seed the HOST draws from torch's generator per launch
// (torch.randint into a device tensor: fresh on every CUDA-graph replay) and counted by (particle, element, attempt),
// so a draw is a pure function of (seed, indices) -- independent of the launch geometry.
#pragma once
#include "common.cuh"

namespace apgs {

__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
    constexpr unsigned int M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const unsigned int hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
        const unsigned int hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += W0; key.y += W1;
    }
    return ctr;
}
// 24-bit uniform in (0, 1): never 0, so logs are finite
__device__ __forceinline__ float u01_open(unsigned int x) { return ((x >> 8) + 0.5f) * 5.9604644775390625e-08f; }

// two independent standard normals from two 32-bit words (Box-Muller)
// (SFU forms: __logf / __sincosf are accurate to ~1e-6 absolute on these ranges -- a perturbation of the variate far below
// the resolution of the 24-bit uniforms it is made from)
__device__ __forceinline__ void box_muller(unsigned int a, unsigned int b, float& n0, float& n1) {
    const float r = sqrtf(-2.f * __logf(u01_open(a)));
    float s, c;
    __sincosf(6.283185307179586f * u01_open(b) - 3.141592653589793f, &s, &c);      // angle in (-pi, pi)
    n0 = r * c; n1 = r * s;
}

// standard Gamma(alpha) by Marsaglia & Tsang (2000), the method ATen's sample_gamma uses; `first` are the random words of
// attempt 0 (x = first normal, u = third word), further attempts (probability < 5 % for alpha >= 1) draw new words.
// alpha < 1: boost  Gamma(alpha) = Gamma(alpha + 1) * U^(1/alpha).
__device__ __forceinline__ float gamma_mt(float alpha, float x0, unsigned int u0w, unsigned int boost_w, uint4 ctr, uint2 key) {
    float scale = 1.f;
    // U^(1/alpha) = 2^(lg2(U) / alpha) with the SFU forms (relative error ~1e-6 |lg2 U| / alpha on a random factor: far below the
    // resolution of the 24-bit uniform it is made from; powf costs ~100 instructions and half a warp takes this branch)
    if (alpha < 1.f) {
        float l, e;
        asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(u01_open(boost_w)));
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(__fdividef(l, alpha)));
        scale = e;
        alpha += 1.f;
    }
    const float d = alpha - 0.3333333333f, c = rsqrtf(9.f * d);
    float x = x0;
    unsigned int uw = u0w;
    for (int attempt = 1; attempt < 64; ++attempt) {
        const float t = 1.f + c * x;
        if (t > 0.f) {
            const float v = t * t * t;
            const float u = u01_open(uw);
            if (__logf(u) < 0.5f * x * x + d - d * v + d * __logf(v)) return scale * d * v;
        }
        ctr.w = (unsigned int)attempt;
        const uint4 w = philox4x32_10(ctr, key);
        float dummy;
        box_muller(w.x, w.y, x, dummy);
        uw = w.z;
    }
    return scale * d;                       // unreachable in practice (64 rejections in a row: p < 1e-80)
}

}  // namespace apgs
