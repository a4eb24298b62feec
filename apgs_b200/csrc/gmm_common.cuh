// Device helpers shared by the GMM kernels: torch.distributions log-densities restated in fp32
// with the library's operation order (no FMA contraction on the reference-order paths).
#pragma once
#include "tile.cuh"

namespace apgs {

#define LOG_SQRT_2PI_F 0.918938533204672741780329736406f
#define FLT_EPS_F 1.1920928955078125e-07f

__device__ __forceinline__ float xlogy_f(float a, float b) { return a == 0.f ? 0.f : __fmul_rn(a, logf(b)); }

// torch/distributions/gamma.py:89-98
__device__ __forceinline__ float gamma_lp(float v, float c, float r) {
    float t = __fadd_rn(xlogy_f(c, r), xlogy_f(__fsub_rn(c, 1.f), v));
    t = __fsub_rn(t, __fmul_rn(r, v));
    return __fsub_rn(t, lgammaf(c));
}
// torch/distributions/normal.py:87-101
__device__ __forceinline__ float normal_lp(float v, float loc, float scale) {
    const float var = __fmul_rn(scale, scale);
    const float d = __fsub_rn(v, loc);
    float t = __fdiv_rn(-__fmul_rn(d, d), __fmul_rn(2.f, var));
    t = __fsub_rn(t, logf(scale));
    return __fsub_rn(t, LOG_SQRT_2PI_F);
}
// 1 / sqrt(v)  as the reference writes it (models.py:34,74,147,177): IEEE sqrt then IEEE divide
__device__ __forceinline__ float inv_sqrt_ref(float v) { return __fdiv_rn(1.f, __fsqrt_rn(v)); }

template <int LANES, typename T>
__device__ __forceinline__ T group_sum(T v) {
#pragma unroll
    for (int o = LANES / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}


// Compensated fp32 accumulation (Kahan): a lane adds <= ceil(N/LANES) terms; the pair (sum, comp) is
// widened to fp64 only when lanes are combined.  No FMA contraction may touch these four operations.
__device__ __forceinline__ void kahan_add(float& s, float& c, float v) {
    const float y = __fsub_rn(v, c);
    const float t = __fadd_rn(s, y);
    c = __fsub_rn(__fsub_rn(t, s), y);
    s = t;
}
struct KahanF {
    float s, c;
    __device__ __forceinline__ KahanF() : s(0.f), c(0.f) {}
    __device__ __forceinline__ void add(float v) { kahan_add(s, c, v); }
    __device__ __forceinline__ double value() const { return (double)s - (double)c; }
};

// Packed fp32x2 arithmetic (sm_100: SASS FFMA2 / FMUL2): two fp32 lanes per issue slot.  The fused z step is bound by
// instruction issue, not by the FMA pipe, so halving the number of FMA-pipe instructions is a direct win.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

// plain fp32 partial sum of a handful of terms (same interface as KahanF)
struct PlainF {
    float s;
    __device__ __forceinline__ PlainF() : s(0.f) {}
    __device__ __forceinline__ void add(float v) { s = __fadd_rn(s, v); }
    __device__ __forceinline__ double value() const { return (double)s; }
};

// tanh with ~1e-7 absolute error from a PRE-SCALED argument a = 2*log2(e)*x:
//   tanh(x) = 1 - 2 / (2^a + 1)      -> MUFU.EX2, FADD, MUFU.RCP, FFMA  (flush-to-zero forms: no denormal
// fix-up code; 2^a underflowing to 0 gives -1 and overflowing to +inf gives +1, the correct saturations)
#define TANH_PRESCALE 2.885390081777926814719849f
__device__ __forceinline__ float ex2_ftz(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_ftz(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// log with the SFU (lg2.approx, abs error ~2^-22 * ln 2 on the result): used for log q of a categorical, whose
// magnitude is <= 16 (eps clamp), so the error is ~1e-7 relative to the sums it enters
__device__ __forceinline__ float log_fast(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y * 0.693147180559945309417f;
}
// lgamma(x), x > 0, for the Gamma normaliser of the eta step: shift up by four (lgamma(x) = lgamma(x+4) - log(x(x+1)(x+2)(x+3)))
// and Stirling's series at y = x + 4 >= 4 (first omitted term 1/(1680 y^7) < 4e-8).  ~25 instructions and two SFU logs
// instead of the ~100 of lgammaf(); absolute error ~3e-7 |lgamma| + 1e-6, the same order as lgammaf's own (a few ulp).
__device__ __forceinline__ float lgamma_pos(float x) {
    const bool big = x > 1.0e8f;                        // the shift is irrelevant there and x^4 would overflow
    const float y = big ? x : x + 4.f;
    const float ry = rcp_ftz(y), ry2 = ry * ry;
    const float series = ry * fmaf(ry2, fmaf(ry2, 7.9365079365e-4f, -2.7777777778e-3f), 8.3333333333e-2f);
    const float st = fmaf(y - 0.5f, log_fast(y), -y) + (0.918938533204672742f + series);
    const float prod = (x * (x + 1.f)) * ((x + 2.f) * (x + 3.f));
    return big ? st : st - log_fast(prod);
}
__device__ __forceinline__ float tanh_prescaled(float a) {
    return fmaf(rcp_ftz(ex2_ftz(a) + 1.f), -2.f, 1.f);
}

}  // namespace apgs
