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
__device__ __forceinline__ float lg2_ftz(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
#define LN2_F 0.693147180559945309417f
#define LOG_EPS_F (-15.9423851528787f)          // log(FLT_EPSILON)
#define LOG_1M_EPS_F (-1.1920929665620834e-07f)   // log(1 - FLT_EPSILON)
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

// ---------------------------------------------------------------------------------------------------------------
// Enc_apg_eta with one-hot z (apgs/gmm/models.py:56-66 + kls_gmm.py:92-119), shared by the instance-major statistics
// kernel (gmm.cu: k_gmm_nss_im) and by the fused z step that emits the NEXT eta step's inputs (gmm_fused.cu).
// Because z_n is one-hot, the two Linear heads on cat(x_n, z_n) take only K values per data point: the factors
// { x'_n (D), gamma_n (K) } are tabulated once per CTA over (point, label), and a particle accumulates
// sum gamma_k, sum gamma_k x'_d, sum gamma_k x'_d^2 (the neural statistics) and, centred at `cen`, the plain one-hot
// statistics { cnt_k, sum z (x - c), sum z (x - c)^2 } that the eta-step log-likelihoods are functions of.
// Both kernels run EXACTLY this code, so they agree bit for bit.
// ---------------------------------------------------------------------------------------------------------------
// lanes that split a particle's points in the statistics accumulation (k_gmm_nss_im: 4 lanes x 64 particles per CTA).
// The fused z step keeps the same partition -- lanes >= kEtaNssLanes of its 8 idle in that pass -- so both agree bit for bit.
#ifndef APGS_NSS_IM_LANES
#define APGS_NSS_IM_LANES 4
#endif
constexpr int kEtaNssLanes = APGS_NSS_IM_LANES;

template <int K, int D>
struct EtaTab {
    // row stride (floats): a multiple of four with an odd number of quads, so consecutive rows fall on distinct banks
    static constexpr int RS = (((K + D + 3) / 4) | 1) * 4;
    static constexpr int TE = K + 2 * K * D;                  // neural statistics per particle
    static constexpr int NS = K * (1 + 2 * D);                // plain statistics per particle
    static constexpr int RW = (TE > NS ? TE : NS) | 1;        // cross-lane reduction row (odd stride)
};

// one table row: the heads evaluated on cat(x_n, e_lab), in the accumulation order of the un-tabulated form
template <int K, int D>
__device__ __forceinline__ void eta_head_factors(const float* __restrict__ Wg, const float* __restrict__ bg,
                                                 const float* __restrict__ Wo, const float* __restrict__ bo,
                                                 const float (&xin)[D], int lab, float* row) {
    constexpr int IN = D + K;
    float lg[K], mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        float acc = __ldg(bg + k);
#pragma unroll
        for (int d = 0; d < D; ++d) acc = fmaf(__ldg(Wg + k * IN + d), xin[d], acc);
        acc += __ldg(Wg + k * IN + D + lab);
        lg[k] = acc;
        mx = fmaxf(mx, acc);
    }
    float se = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k) { lg[k] = ex2_ftz((lg[k] - mx) * 1.4426950408889634f); se += lg[k]; }
    const float rse = rcp_ftz(se);
    float xp[D];
#pragma unroll
    for (int d = 0; d < D; ++d) {
        float acc = __ldg(bo + d);
#pragma unroll
        for (int i = 0; i < D; ++i) acc = fmaf(__ldg(Wo + d * IN + i), xin[i], acc);
        xp[d] = acc + __ldg(Wo + d * IN + D + lab);
    }
#pragma unroll
    for (int d = 0; d < D; ++d) row[d] = xp[d];
#pragma unroll
    for (int k = 0; k < K; ++k) row[D + k] = lg[k] * rse;
}

template <int K, int D, bool STATS>
struct EtaNssAcc {
    static constexpr int KS = STATS ? K : 1;
    float ag[K], a2[K][D], a3[K][D];                 // sum gamma_k, sum gamma_k x'_d, sum gamma_k x'_d^2
    f32x2 a2p[K], a3p[K];                            // the same, packed, when D == 2
    float cen[KS][D];
    f32x2 c2p[KS], c3p[KS], ncen2[KS];
    float c1[KS], c2[KS][D], c3[KS][D];

    __device__ __forceinline__ void init(const float (&center)[K][D]) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            ag[k] = 0.f; a2p[k] = 0ull; a3p[k] = 0ull;
#pragma unroll
            for (int d = 0; d < D; ++d) { a2[k][d] = 0.f; a3[k][d] = 0.f; }
        }
        if (STATS) {
#pragma unroll
            for (int k = 0; k < K; ++k) {
                c1[k] = 0.f;
#pragma unroll
                for (int d = 0; d < D; ++d) { c2[k][d] = 0.f; c3[k][d] = 0.f; cen[k][d] = center[k][d]; }
                c2p[k] = pack2(0.f, 0.f); c3p[k] = c2p[k];
                ncen2[k] = pack2(-cen[k][0], -cen[k][D - 1]);
            }
        }
    }
    // t: the table row of (point, lab); xn: the point itself (shared memory)
    __device__ __forceinline__ void add(const float* t, int lab, const float* xn) {
        if (D == 2 && K >= 2) {
            const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(t);          // (x'_0, x'_1), (gamma_0, gamma_1)
            float g[K];
            unpack2(v.y, g[0], g[1]);
#pragma unroll
            for (int k = 2; k < K; ++k) g[k] = t[D + k];
            const f32x2 xp2 = v.x, sq2 = mul2(xp2, xp2);                           // ob_expand, ob_expand**2
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const f32x2 gk = pack2(g[k], g[k]);
                ag[k] = __fadd_rn(ag[k], g[k]);
                a2p[k] = fma2(gk, xp2, a2p[k]);                                    // z_expand * ob_expand
                a3p[k] = fma2(gk, sq2, a3p[k]);                                    // z_expand * ob_expand**2
            }
        } else {
            float xp[D], sq[D];
#pragma unroll
            for (int d = 0; d < D; ++d) { xp[d] = t[d]; sq[d] = __fmul_rn(xp[d], xp[d]); }
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const float gk = t[D + k];
                ag[k] = __fadd_rn(ag[k], gk);
#pragma unroll
                for (int d = 0; d < D; ++d) { a2[k][d] = fmaf(gk, xp[d], a2[k][d]); a3[k][d] = fmaf(gk, sq[d], a3[k][d]); }
            }
        }
        if (STATS) {
            float ind[K];
#pragma unroll
            for (int k = 0; k < K; ++k) { ind[k] = (lab == k) ? 1.f : 0.f; c1[k] += ind[k]; }
            if (D == 2) {
                f32x2 tt = pack2(xn[0], xn[D - 1]);
#pragma unroll
                for (int k = 0; k < K; ++k) tt = fma2(pack2(ind[k], ind[k]), ncen2[k], tt);      // x - c_lab
                const f32x2 t2 = mul2(tt, tt);
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const f32x2 ik = pack2(ind[k], ind[k]);
                    c2p[k] = fma2(ik, tt, c2p[k]);
                    c3p[k] = fma2(ik, t2, c3p[k]);
                }
            } else {
#pragma unroll
                for (int d = 0; d < D; ++d) {
                    float c = 0.f;
#pragma unroll
                    for (int k = 0; k < K; ++k) c = fmaf(ind[k], cen[k][d], c);
                    const float tt = xn[d] - c, t2 = tt * tt;
#pragma unroll
                    for (int k = 0; k < K; ++k) { c2[k][d] = fmaf(ind[k], tt, c2[k][d]); c3[k][d] = fmaf(ind[k], t2, c3[k][d]); }
                }
            }
        }
    }
    // this lane's partial neural statistics -> mine[0 .. K + 2 K D)
    __device__ __forceinline__ void spill_nss(float* mine) {
        constexpr int KD = K * D;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            if (D == 2 && K >= 2) { unpack2(a2p[k], a2[k][0], a2[k][D - 1]); unpack2(a3p[k], a3[k][0], a3[k][D - 1]); }
            mine[k] = ag[k];
#pragma unroll
            for (int d = 0; d < D; ++d) { mine[K + k * D + d] = a2[k][d]; mine[K + KD + k * D + d] = a3[k][d]; }
        }
    }
    // this lane's partial plain statistics -> mine[0 .. K (1 + 2 D))
    __device__ __forceinline__ void spill_stats(float* mine) {
        constexpr int ST = 1 + 2 * D;
#pragma unroll
        for (int k = 0; k < KS; ++k) {
            if (D == 2) { unpack2(c2p[k], c2[k][0], c2[k][D - 1]); unpack2(c3p[k], c3[k][0], c3[k][D - 1]); }
            mine[k * ST] = c1[k];
#pragma unroll
            for (int d = 0; d < D; ++d) { mine[k * ST + 1 + d] = c2[k][d]; mine[k * ST + 1 + D + d] = c3[k][d]; }
        }
    }
};

// Normal-Gamma posterior element from the three neural statistics (kls_gmm.py:110-118; fp32, reference order, Q1 / Q2)
__device__ __forceinline__ void ng_posterior_elem(float f1, float f2, float f3, float a0, float b0, float m0, float n0,
                                                  float& alpha, float& beta, float& mu, float& nu) {
    const float cnt = (f1 == 0.f) ? 1.f : f1;                                   // Q1
    const float x_bar = __fdiv_rn(f2, cnt);
    alpha = __fadd_rn(a0, __fmul_rn(cnt, 0.5f));                                // x / 2 == x * 0.5 exactly (no IEEE-division slow path)
    nu = __fadd_rn(n0, cnt);
    const float cn = __fadd_rn(cnt, n0);
    mu = __fdiv_rn(__fadd_rn(__fmul_rn(m0, n0), f2), cn);
    const float u1 = __fmul_rn(__fsub_rn(f3, __fdiv_rn(__fmul_rn(f2, f2), cnt)), 0.5f);
    const float dx = __fsub_rn(x_bar, n0);                                       // Q2: prior_nu
    const float u2 = __fmul_rn(__fmul_rn(__fdiv_rn(__fmul_rn(cnt, n0), cn), __fmul_rn(dx, dx)), 0.5f);
    beta = __fadd_rn(__fadd_rn(b0, u1), u2);
}

// A particle's LANES partials meet in a scratch block (LANES rows of RW floats): lane j finishes the posterior elements
// e = j, j + LANES, ... (fp64 across lanes), then every lane sums its share of the plain statistics (fp32, lane order).
// Called by all lanes of the warp (the __syncwarp()s are unconditional); only `valid` particles write.
template <int K, int D, int LANES, bool STATS>
__device__ __forceinline__ void eta_nss_finish(EtaNssAcc<K, D, STATS>& acc, float* blk, int j, bool valid, long long r,
                                               const float* ra, const float* rb, const float* rm, const float* rn,
                                               float* oa, float* ob, float* om, float* on, float* stats) {
    constexpr int KD = K * D, RW = EtaTab<K, D>::RW, NS = EtaTab<K, D>::NS;
    constexpr int EPL = (KD + LANES - 1) / LANES;                  // posterior elements per lane
    float* mine = blk + j * RW;
    acc.spill_nss(mine);
    __syncwarp();
    double n1[EPL], n2[EPL], n3[EPL];
#pragma unroll
    for (int q = 0; q < EPL; ++q) {
        n1[q] = 0.0; n2[q] = 0.0; n3[q] = 0.0;
        const int e = j + q * LANES;
        if (valid && e < KD) {
            const int k = e / D;
#pragma unroll
            for (int i = 0; i < LANES; ++i) {
                const float* o = blk + i * RW;
                n1[q] += (double)o[k]; n2[q] += (double)o[K + e]; n3[q] += (double)o[K + KD + e];
            }
        }
    }
    if (STATS) {
        __syncwarp();
        acc.spill_stats(mine);
        __syncwarp();
        if (valid) {
            float* so = stats + r * NS;
            for (int e = j; e < NS; e += LANES) {
                float v = 0.f;
#pragma unroll
                for (int i = 0; i < LANES; ++i) v += blk[i * RW + e];
                so[e] = v;
            }
        }
    }
#pragma unroll
    for (int q = 0; q < EPL; ++q) {
        const int e = j + q * LANES;
        if (!(valid && e < KD)) continue;
        float alpha, beta, mu, nu;
        ng_posterior_elem((float)n1[q], (float)n2[q], (float)n3[q], __ldg(ra + e), __ldg(rb + e), __ldg(rm + e), __ldg(rn + e), alpha, beta,
                          mu, nu);
        const long long o = r * KD + e;
        oa[o] = alpha; ob[o] = beta; om[o] = mu; on[o] = nu;
    }
}

}  // namespace apgs
