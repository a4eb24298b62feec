// Fused GMM z-step for sm_100a (SURVEY.md section 8f row N1 + rows G3/W2):
//   Enc_apg_z.forward (apgs/gmm/models.py:102-115): per (point n, cluster k) MLP
//       logit_nk = W2 . tanh(W1 [x_n ; mu_k ; tau_k] + b1) + b2,   q_n = softmax_k(logit_n)
//   + categorical draw of z_n by inverse CDF from a supplied uniform (RNG stays in PyTorch)
//   + the z-step importance weight (apgs/gmm/objectives.py:197-227) against z_old
// in ONE pass over the particles: reads x (per instance), tau/mu (48 B), z_old (720 B one-hot, or 64 B as labels) and
// the uniforms (240 B) per particle and writes z_new (720 B / 64 B) + a few scalars, instead of PyTorch's ~100 KB of
// (S,B,N,K,H) activations per particle.  The first Linear is split into its x / mu / tau column blocks:
//   W1 [x;mu_k;tau_k] + b1 = (W1x x_n) + (W1m mu_k + W1t tau_k + b1) =: xh_n + e_k
// and, because 2^(a+b) = 2^a 2^b, the exponential inside tanh(u) = 1 - 2/(2^(2u log2 e) + 1) factorises too:
//   2^(xh'_n + e'_k) = Ex_n * Ee_k,   Ex_n = 2^xh'_n (a function of the data only: one table per CTA, see "tile mapping"),
//   Ee_k = 2^e'_k (K*H values per particle, computed once in the prologue and kept in shared memory).
// The inner loop is then  y = fma(Ex, Ee, c)  per tanh, and the output layer's  sum_h w_h / y_h  is accumulated four
// hidden units at a time as ONE fraction (numerator and denominator built pairwise, one MUFU.RCP per four terms), in
// packed fp32x2: 1.75 instructions (3.5 FMA-pipe cycles) and 0.25 SFU operations per tanh instead of 13 / 2.
// Compute-bound (FMA pipe), not HBM-bound.
#include "gmm_common.cuh"
#include "rng.cuh"

#ifndef APGS_ZF_EXPT
#define APGS_ZF_EXPT 0      // measurement-only builds: 1 = no MLP loop, 2 = no per-point epilogue (lean launch),
                            // 3 = constant uniforms (no Philox), 4 = constant Ee table (no per-particle first layer)
#endif

namespace apgs {

struct ZFusedArgs {
    const float* x;                 // (B, N, D)
    const float *tau, *mu;          // (R, K, D)
    const float *W1, *b1, *W2, *b2; // (H, 3D), (H), (H), (1)
    const float* z_old;             // (R, N, K) one-hot, or null (sample-only mode)
    const float* z_forced;          // (R, N, K) one-hot to use instead of sampling, or null
    const float* u;                 // (R, N) uniforms in [0,1), required unless z_forced or seed
    const unsigned long long* seed; // device word keying the in-kernel Philox stream that replaces `u`, or null
    // label mode (compact latents inside a sweep): assignments as one byte per point instead of K one-hot floats
    const unsigned char* zl_old;    // (R, N) labels of the backward term, or null
    unsigned char* zl_new;          // (R, N) labels out, or null (then z_new one-hot is written)
    long long LN;                   // bytes between the label rows of consecutive particles (>= N)
    const float* rpi;               // (K)
    int S, B, N, H;
    float* z_new;                   // (R, N, K) out
    float* q_probs;                 // (R, N, K) out, optional
    float *log_w, *log_q_f, *dens, *log_w_f, *log_w_b;   // (R), optional
    float *lw_local, *lq_local;                           // (R, N), optional
    // NSS (label mode only): also emit what the NEXT eta step needs from the labels just drawn -- Enc_apg_eta's
    // Normal-Gamma posterior parameters (apgs/gmm/models.py:56-66, kls_gmm.py:92-119) and the one-hot statistics of
    // z_new centred at this particle's mu -- so that no separate pass over the labels is needed after the resampling
    // step (per-particle functions of (z, mu) travel with the particle through the ancestor gather).
    const float *Wg, *bg, *Wo, *bo;                       // Enc_apg_eta heads (K, D+K), (K), (D, D+K), (D)
    const float *ra, *rb, *rm, *rn;                       // prior alpha, beta, mu, nu (K, D)
    float *oa, *ob, *om, *on;                             // posterior out (R, K, D)
    float* stats;                                         // (R, K, 1 + 2 D) out
    // per-instance tables made once per objective by k_gmm_z_tables (everything of the step that depends only on the data
    // and the encoder weights), or null: the CTA then builds its own copy in the prologue
    const float* tables;
    // ancestors (S,B) of a resampling step whose gather was deferred, or null: the particle's INPUT latents (tau, mu, old
    // assignments) are read from row (src_anc[s,b], b); outputs and random numbers belong to row (s,b)
    const int* src_anc;
    const int* src_anc_z;          // the same for the old assignments when they went through other resampling steps (null: src_anc)
};

// ---- per-instance tables ------------------------------------------------------------------------------------------
// Everything of the z step that is a function of (instance data, encoder weights) alone -- the Ex table, its per-unit
// maxima, the output weights, b2' and the clamped log prior, the points themselves -- is the same for the ~S/32 CTAs of
// an instance and for every sweep of an objective.  k_gmm_z_tables writes it once per instance as one contiguous block
// in the kernel's own shared-memory layout, and a CTA of the step fetches its instance's block with ONE bulk copy
// (cp.async.bulk, ~10 KB from L2) that lands while the per-particle prologue runs.  Behind the copied part sits the
// first layer's particle-side weights, pre-scaled (rows of 8 floats per hidden unit: W1m'[D], W1t'[D], b1' - 31), read
// with two L1-resident 16-byte loads per unit.
struct ZTabLayout {
    int xm, aw, misc, x, copy_floats, wt, total;      // offsets in floats (the Ex table sits at 0)
};
__host__ __device__ inline ZTabLayout zf_tab_layout(int N, int H4, int D) {
    ZTabLayout L;
    const int EXS = (N + 31) / 32 * 32;
    L.xm = H4 * EXS;
    L.aw = L.xm + H4;
    L.misc = L.aw + H4;                                // b2', log pi_k (K <= 6), [7]: 1 = the factorised loop is safe
    L.x = L.misc + 8;
    L.copy_floats = L.x + (N * D + 3) / 4 * 4;
    L.wt = L.copy_floats;
    L.total = L.wt + H4 * 8;
    return L;
}

// in-kernel uniforms of the categorical draw: one Philox block per (particle row, lane, pass) -- its four words are the
// uniforms of that lane's PB points of the pass; [0, 1) with 24 bits like torch.rand
__device__ __forceinline__ uint4 philox_block(long long r, int slot, uint2 key) {
    return philox4x32_10(make_uint4((unsigned int)r, (unsigned int)((unsigned long long)r >> 32), (unsigned int)slot, 0x7a737465u), key);
}
__device__ __forceinline__ float u01_co(unsigned int w) { return (float)(w >> 8) * 5.9604644775390625e-08f; }

// ---- tile mapping -----------------------------------------------------------------------------------------------
// A CTA owns ONE instance b and P particles s0..s0+P-1 of it (rows r = s*B + b: P separate N*K*4-byte rows, each
// moved by its own bulk copy).  Everything that depends only on the data is then shared by the whole CTA:
// Ex_n[h] = 2^(W1x x_n)' is tabulated ONCE per CTA in shared memory (N*H ex2 instead of P*N*H), which takes the
// exponentials out of the inner loop altogether -- per quad of hidden units a thread issues 6 + 2K LDS.128 (Ex of its
// four points, the output weights, Ee of the K clusters) and, per cluster and pair of points, 14 packed-fp32
// instructions and two MUFU.RCP.
//
// Range contract of the factorised loop.  The fraction's denominator is the product of four y = Ex*Ee + 1 >= 1 (four
// hidden units of one point); the y are carried scaled by 2^-31 (y' = Ex * (2^-31 Ee) + 2^-31, the 2^-31 folded
// into the Ee table; the accumulated sum comes out scaled by 2^31 and is rescaled once after the loop), so y in
// [1, 2^62] keeps the product of four y' inside [2^-124, 2^124], the numerators (weight x three y') inside
// |w| [2^-93, 2^95], and no clamp is needed.  A particle may
// run this loop iff  (a) every e'_k[h] is in [-80, 80],  (b) XM[h] = max_n xh'_n[h] <= 126 (Ex finite),
// (c) max_h (XM[h] + max_k e'_k[h]) <= 61, and (d) max_h |W2[h]| <= 2^24.  XM is per CTA, the e' test per particle; both are evaluated in the
// prologue.  Any other particle (saturating pre-activations far beyond anything a trained encoder produces, or
// NaN/inf inputs) runs the plain loop  y = 2^(xh' + e') + 1, 1/y by MUFU.RCP per element  -- exact for every
// input, ~3x slower, and in practice never taken.

// sum over four hidden units of A_h / y_h as ONE fraction (pairwise: n/d = (A0 y1 + A1 y0)/(y0 y1), then the two pairs
// combined the same way), for two points at a time in packed fp32: 14 packed instructions and one reciprocal per lane
// for four tanh terms -- 3.5 FMA-pipe cycles per term against 4.25 for inverting each y (batch inversion) first.
__device__ __forceinline__ f32x2 frac4(const f32x2 (&ex)[4], const f32x2 (&ee)[4], const f32x2 (&aw)[4], f32x2 c2,
                                       f32x2 acc) {
    const f32x2 y0 = fma2(ex[0], ee[0], c2), y1 = fma2(ex[1], ee[1], c2);
    const f32x2 y2 = fma2(ex[2], ee[2], c2), y3 = fma2(ex[3], ee[3], c2);
    const f32x2 n01 = fma2(aw[0], y1, mul2(aw[1], y0)), d01 = mul2(y0, y1);
    const f32x2 n23 = fma2(aw[2], y3, mul2(aw[3], y2)), d23 = mul2(y2, y3);
    const f32x2 n = fma2(n01, d23, mul2(n23, d01));
    float dl, dh;
    unpack2(mul2(d01, d23), dl, dh);
    return fma2(n, pack2(rcp_ftz(dl), rcp_ftz(dh)), acc);
}

// ev_ptr: this particle's table, per quad of hidden units h0..h3 and cluster k the four values Ee'_k[h0..h3] (one
// LDS.128) -- index ee_slot(h, k); aw_ptr: the CTA-wide output weights -2 W2[h].  H4 = H rounded up to 4 (padding rows: Ee' = 0, weight 0).  The accumulators
// come out scaled by 2^31 (y' = 2^-31 y and the weights are unscaled), the caller multiplies by 2^-31.
template <int K>
__device__ __forceinline__ int ee_slot(int h, int k) { return (h >> 2) * (4 * K) + k * 4 + (h & 3); }

template <int K>
__device__ __forceinline__ void mlp_quad_fast(const unsigned char* ex_ptr, int ex_stride_b, const unsigned char* ev_ptr,
                                              const unsigned char* aw_ptr, int H4, f32x2 (&L02)[K], f32x2 (&L13)[K]) {
    const f32x2 c2 = pack2(4.656612873077393e-10f, 4.656612873077393e-10f);   // 2^-31
    // unroll 1: unroll 2 measured slower (1.60 vs 1.53 ms: the second quad's table loads raise the register pressure), and
    // unroll 2 / 4 / 8 with a compiler barrier between the quads (immediate offsets, no pointer updates) 1.17 / 1.16 / 1.19
    // vs 1.155 ms
#pragma unroll 1
    for (const unsigned char* ev_end = ev_ptr + (size_t)H4 * K * 4; ev_ptr != ev_end;      // one compare, no separate counter
         ev_ptr += 4 * K * 4, ex_ptr += 4 * (size_t)ex_stride_b, aw_ptr += 16) {
        f32x2 ex02[4], ex13[4], aw[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const ulonglong2 exq = *reinterpret_cast<const ulonglong2*>(ex_ptr + (size_t)i * ex_stride_b);
            ex02[i] = exq.x; ex13[i] = exq.y;                         // (Ex[q0], Ex[q2]), (Ex[q1], Ex[q3])
        }
        {
            const float4 v = *reinterpret_cast<const float4*>(aw_ptr);
            aw[0] = pack2(v.x, v.x); aw[1] = pack2(v.y, v.y); aw[2] = pack2(v.z, v.z); aw[3] = pack2(v.w, v.w);
        }
        const float4* evq = reinterpret_cast<const float4*>(ev_ptr);
#pragma unroll
        for (int k = 0; k < K; ++k) {                                 // cluster k of the four units: one LDS.128
            const float4 v = evq[k];
            const f32x2 ee[4] = {pack2(v.x, v.x), pack2(v.y, v.y), pack2(v.z, v.z), pack2(v.w, v.w)};
            L02[k] = frac4(ex02, ee, aw, c2, L02[k]);
            L13[k] = frac4(ex13, ee, aw, c2, L13[k]);
        }
    }
}

// the plain loop: ev holds the raw e'_k[h] at ee_slot(h, k), aw the weights -2 W2[h];
// xh' is recomputed exactly as the Ex table does
template <int K, int D, int PB>
__device__ __forceinline__ void mlp_quad_plain(const float* W1, const float (&xv)[PB][D], const float* ev, const float* aw,
                                               int H, float (&logit)[PB][K]) {
    for (int h = 0; h < H; ++h) {
        const float wm2 = aw[h];
#pragma unroll
        for (int q = 0; q < PB; ++q) {
            float acc = 0.f;
#pragma unroll
            for (int d = 0; d < D; ++d) acc = fmaf(TANH_PRESCALE * __ldg(W1 + h * (3 * D) + d), xv[q][d], acc);
#pragma unroll
            for (int k = 0; k < K; ++k)      // 2^a = +inf gives 1/y = 0 (tanh = 1), 2^a = 0 gives 1/y = 1 (tanh = -1)
                logit[q][k] = fmaf(wm2, rcp_ftz(ex2_ftz(acc + ev[ee_slot<K>(h, k)]) + 1.f), logit[q][k]);
        }
    }
}

// floats of a particle's private shared-memory block: Ee table (H4 K + 4) + per-unit maxima (H4); with NSS at least the
// LANES x RW scratch rows of the eta-statistics reduction that reuses it
template <int K, int D, int LANES, bool NSS>
__host__ __device__ inline int zf_block_floats(int H4) {
    int b = H4 * K + 4 + H4;
    if (NSS && b < LANES * EtaTab<K, D>::RW) b = LANES * EtaTab<K, D>::RW;
    return (b + 3) / 4 * 4;
}

// magnitudes beyond which the particle-side pre-activations are not guaranteed finite (5 products of two such values)
#define ZF_IN_MAX 1.0e17f

// The per-instance tables, built by all threads of a CTA (the z step's own prologue when no tables are handed in, and
// k_gmm_z_tables): Ex_n[h] = 2^(2 log2(e) W1x x_n) in the MLP's slot order with its per-unit maximum exponent XM[h]
// (+inf: the CTA takes the plain loop), the output weights -2 W2[h], b2' = b2 + sum_h W2[h]  (tanh = 1 - 2/y  =>
// sum_h w2_h tanh_h = sum_h w2_h - sum_h 2 w2_h / y_h; the loops accumulate the second sum), the clamped log prior of
// the assignments, and the instance's points.  The caller synchronises the CTA afterwards.
template <int K, int D, int LANES, int PB>
__device__ __forceinline__ void zf_cta_tables(const float* xb, const float* W1, const float* b1, const float* W2,
                                              const float* b2, const float* rpi, int N, int H, float* s_ex, float* s_xm,
                                              float* s_aw, float* s_misc, float* s_x) {
    const int H4 = (H + 3) / 4 * 4, EXS = (N + LANES * PB - 1) / (LANES * PB) * (LANES * PB);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nthr = blockDim.x, nwarp = nthr >> 5;
    constexpr int XS = 2;                                       // slots per lane kept in registers (N <= 64)
    float xr[XS][D];
#pragma unroll
    for (int si = 0; si < XS; ++si) {
        const int slot = lane + 32 * si, qs = slot & 3, q = ((qs & 1) << 1) | (qs >> 1);   // slot order q0 q2 q1 q3
        const int n = (slot >> 5) * (LANES * PB) + q * LANES + ((slot & 31) >> 2);
#pragma unroll
        for (int d = 0; d < D; ++d) xr[si][d] = (slot < EXS && n < N) ? __ldg(xb + n * D + d) : NAN;
    }
    for (int h = H + warp; h < H4; h += nwarp)                  // padding rows of the last quad of hidden units
        for (int slot = lane; slot < EXS; slot += 32) s_ex[(size_t)h * EXS + slot] = 0.f;
    for (int i = threadIdx.x; i < H4; i += nthr) s_aw[i] = (i < H) ? -2.f * __ldg(W2 + i) : 0.f;
    for (int i = threadIdx.x; i < N * D; i += nthr) s_x[i] = __ldg(xb + i);
#pragma unroll 4
    for (int h = warp; h < H; h += nwarp) {
        float w[D];
#pragma unroll
        for (int d = 0; d < D; ++d) w[d] = TANH_PRESCALE * __ldg(W1 + h * (3 * D) + d);
        float mx = -INFINITY;
        // the unit's particle-side weights: beyond ZF_IN_MAX (or NaN) the pre-activations e' may not be finite -> plain loop
        bool bad = (lane >= D && lane < 3 * D) ? !(fabsf(__ldg(W1 + h * (3 * D) + lane)) <= ZF_IN_MAX)
                                               : (lane == 3 * D ? !(fabsf(__ldg(b1 + h)) <= ZF_IN_MAX) : false);
#pragma unroll
        for (int si = 0; si < XS; ++si) {
            const int slot = lane + 32 * si;
            if (slot < EXS) {
                const int qs = slot & 3, q = ((qs & 1) << 1) | (qs >> 1);
                const bool in = (slot >> 5) * (LANES * PB) + q * LANES + ((slot & 31) >> 2) < N;
                float acc = 0.f;
#pragma unroll
                for (int d = 0; d < D; ++d) acc = fmaf(w[d], xr[si][d], acc);
                if (in) { mx = fmaxf(mx, acc); bad |= !(acc == acc); }
                // an underflow to 0 is exact enough (the sum of exponents is then < -46); overflow is caught via XM
                s_ex[(size_t)h * EXS + slot] = in ? ex2_ftz(acc) : 0.f;
            }
        }
        for (int slot = lane + 32 * XS; slot < EXS; slot += 32) {
            const int qs = slot & 3, q = ((qs & 1) << 1) | (qs >> 1);
            const int n = (slot >> 5) * (LANES * PB) + q * LANES + ((slot & 31) >> 2);
            float acc = 0.f;
            if (n < N) {
#pragma unroll
                for (int d = 0; d < D; ++d) acc = fmaf(w[d], __ldg(xb + n * D + d), acc);
                mx = fmaxf(mx, acc);
                bad |= !(acc == acc);
            }
            s_ex[(size_t)h * EXS + slot] = (n < N) ? ex2_ftz(acc) : 0.f;
        }
        mx = warp_max(bad ? INFINITY : mx);
        if (lane == 0) s_xm[h] = (mx <= 126.f) ? mx : INFINITY;     // +inf: the instance takes the plain loop
    }
    if (warp == 0) {
        float part = 0.f, wmax = 0.f;
        for (int i = lane; i < H; i += 32) {
            const float w = __ldg(W2 + i);
            part += w;
            wmax = fmaxf(wmax, (w == w) ? fabsf(w) : INFINITY);
        }
        part = warp_sum(part);
        wmax = warp_max(wmax);
        if (lane == 0) {
            s_misc[0] = __ldg(b2) + part;
            // (d) output weights: the numerators are |w| [2^-93, 2^95]; beyond 2^24 (or NaN) the instance takes the plain
            // loop (flag in s_misc[7])
            s_misc[7] = (wmax <= 16777216.f) ? 1.f : 0.f;
        }
        if (lane < 6) {
            float v = 0.f;
            if (lane < K) {
                float sp = 0.f;
#pragma unroll
                for (int k = 0; k < K; ++k) sp = __fadd_rn(sp, __ldg(rpi + k));
                v = logf(fminf(fmaxf(__fdiv_rn(__ldg(rpi + lane), sp), FLT_EPS_F), 1.f - FLT_EPS_F));
            }
            s_misc[1 + lane] = v;
        }
    }
}

// row h of the particle-side first-layer weights, pre-scaled for the factorised loop: W1m'[D], W1t'[D], b1' - 31
// (2^-31 folded into the Ee table, see the range contract above); padding rows are zero
template <int D>
__device__ __forceinline__ void zf_wt_row(const float* W1, const float* b1, int h, bool live, float (&wm)[D], float (&wt)[D],
                                          float& bh) {
#pragma unroll
    for (int d = 0; d < D; ++d) {
        wm[d] = live ? TANH_PRESCALE * __ldg(W1 + h * 3 * D + D + d) : 0.f;
        wt[d] = live ? TANH_PRESCALE * __ldg(W1 + h * 3 * D + 2 * D + d) : 0.f;
    }
    bh = live ? __fadd_rn(TANH_PRESCALE * __ldg(b1 + h), -31.f) : 0.f;
}

template <int K, int D>
__global__ void __launch_bounds__(256) k_gmm_z_tables(const float* x, const float* W1, const float* b1, const float* W2,
                                                      const float* b2, const float* rpi, int N, int H, float* tables) {
    static_assert(2 * D + 1 <= 8 && K <= 6, "table rows of 8 floats");
    extern __shared__ __align__(128) unsigned char smem[];
    float* sm = reinterpret_cast<float*>(smem);
    const int H4 = (H + 3) / 4 * 4;
    const ZTabLayout L = zf_tab_layout(N, H4, D);
    const int b = (int)blockIdx.x;
    for (int i = threadIdx.x; i < L.total; i += blockDim.x) sm[i] = 0.f;
    __syncthreads();
    zf_cta_tables<K, D, 8, 4>(x + (size_t)b * N * D, W1, b1, W2, b2, rpi, N, H, sm, sm + L.xm, sm + L.aw, sm + L.misc, sm + L.x);
    for (int h = threadIdx.x; h < H4; h += blockDim.x) {
        float wm[D], wt[D], bh;
        zf_wt_row<D>(W1, b1, h, h < H, wm, wt, bh);
#pragma unroll
        for (int d = 0; d < D; ++d) { sm[L.wt + h * 8 + d] = wm[d]; sm[L.wt + h * 8 + D + d] = wt[d]; }
        sm[L.wt + h * 8 + 2 * D] = bh;
    }
    __syncthreads();
    float4* dst = reinterpret_cast<float4*>(tables + (size_t)b * L.total);      // L.total is a multiple of 4 floats
    for (int i = threadIdx.x; i < L.total / 4; i += blockDim.x) dst[i] = reinterpret_cast<const float4*>(sm)[i];
}

// LEAN: the launch a sweep repeats -- labels in, labels out, uniforms given, no one-hot rows, no per-point outputs.
// The options are compile-time false there, so their tests, pointers and staging code drop out of the hot variant.
// HC / NC: hidden width and points per instance when known at compile time (the reference's 32 and 60), else 0.
// TABM: 1 / 0 = per-instance tables handed in / not, known at compile time (the hot lean instantiations: the other path's
// code -- predicated off at run time otherwise, but still issued -- drops out); 2 = decided at run time.
template <int K, int D, int P, int LANES, int PB, bool LEAN, int HC = 0, int NC = 0, bool NSS = false, int TABM = 2>
__global__ void __launch_bounds__(P* LANES, (LEAN ? 4 : 3) * 32 / P) k_gmm_z_fused(ZFusedArgs a) {
    static_assert(LANES == 8 && PB == 4, "the Ex table layout assumes 8 lanes x 4 points per particle and pass");
    static_assert(!NSS || LEAN, "the eta statistics ride on the lean label-mode launch");
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ unsigned long long bar;
    constexpr int KD = K * D;
    constexpr int RK = K;                                      // table floats per hidden unit
    const int N = NC ? NC : a.N, H = HC ? HC : a.H, H4 = (H + 3) / 4 * 4;
    const int NK = N * K;
    const int nIt = (N + LANES * PB - 1) / (LANES * PB);       // passes of 32 points
    const int EXS = nIt * LANES * PB;                          // floats of Ex per hidden unit
    const int EPs = H4 * RK + 4;                               // floats of a particle's Ee table
    // per-particle private block: [ Ee table (EPs) | max_k e'_k[h] (H4) ]; once the particle's lanes are through the
    // MLP the block is scratch for the cross-lane reduction of the eta statistics (NSS: LANES rows of RW floats)
    const int EPB = zf_block_floats<K, D, LANES, NSS>(H4);
    const bool lab_old = LEAN || a.zl_old != nullptr, lab_new = LEAN || a.zl_new != nullptr;
    const bool has_bwd = LEAN || a.z_old != nullptr || lab_old;
    const bool forced = !LEAN && a.z_forced != nullptr;
    const bool rng = a.seed != nullptr && !forced;
    const bool have_tab = TABM == 2 ? a.tables != nullptr : TABM == 1;
    uint2 key = make_uint2(0u, 0u);
    if (rng) { const unsigned long long sd = __ldg(a.seed); key = make_uint2((unsigned int)sd, (unsigned int)(sd >> 32)); }
    const bool hot_new = !LEAN && a.z_new != nullptr;                      // one-hot z_new wanted (labels may come along)
    const bool tiles = !LEAN && (a.z_old != nullptr || forced || hot_new); // any one-hot row staged in / out?
    const bool want_q = !LEAN && a.q_probs != nullptr, want_lw = !LEAN && a.lw_local != nullptr;
    const bool want_lq = !LEAN && a.lq_local != nullptr;
    const size_t tile_f = tiles ? ((size_t)P * NK + 3) / 4 * 4 : 0;
    const ZTabLayout TL = zf_tab_layout(N, H4, D);
    float* s_zo = reinterpret_cast<float*>(smem);              // z_old rows
    float* s_zn = s_zo + tile_f;                               // z_new rows (staged in from z_forced, out to z_new)
    // the per-instance tables, contiguous in the layout of zf_tab_layout (one bulk copy when they are handed in)
    float* s_ex = s_zn + tile_f;                               // H4 * EXS, slot order inside a quad: q0 q2 q1 q3
    float* s_xm = s_ex + TL.xm;                                // H4
    float* s_aw = s_ex + TL.aw;                                // H4: -2 W2[h], 0 in the padding rows
    float* s_misc = s_ex + TL.misc;                            // 8: b2', log pi_k
    float* s_x = s_ex + TL.x;                                  // N * D: the instance's points (read after the MLP)
    float* s_e = s_ex + TL.copy_floats;                        // P * EPB
    // Normal tables: per particle K rows {mu_d, 1/(2 var_d), log sigma_d + c}_d of TRS floats (16-byte loads)
    constexpr int TRS = (3 * D + 3) / 4 * 4;
    float* s_tab = s_e + (size_t)P * EPB;                      // P * K * TRS
    // lean variant: the particles' uniforms and old labels are prefetched into shared memory with cp.async, so nothing
    // of them lives in registers across the MLP
    const int NU = (N + 3) / 4 * 4, LS = (N + 15) & ~15;
    float* s_u = s_tab + (size_t)P * K * TRS;                  // P * NU (lean)
    unsigned char* s_lb = reinterpret_cast<unsigned char*>(s_u + (size_t)P * NU);   // P * LS (lean): old labels in, new out
    float* s_nt = reinterpret_cast<float*>(s_lb + (size_t)P * LS);                 // N * K rows of EtaTab::RS floats (NSS)
    const int b = (int)blockIdx.x;                             // grid = (B, ceil(S / P)): no division
    const int s0 = (int)blockIdx.y * P;
    const int p = threadIdx.x / LANES, j = threadIdx.x % LANES;
    const bool valid = s0 + p < a.S;
    const long long r = (long long)(s0 + p) * a.B + b;
    const long long rs = (valid && a.src_anc != nullptr) ? (long long)__ldg(a.src_anc + r) * a.B + b : r;   // row of tau, mu
    const long long rz = (valid && a.src_anc_z != nullptr) ? (long long)__ldg(a.src_anc_z + r) * a.B + b : rs;  // row of z_old
    const int rows = min(P, a.S - s0);
    const unsigned int rb = (unsigned int)NK * 4u;
    // a row goes through the bulk-copy engine when its size keeps every row 16-byte aligned
    const bool old_tile = !LEAN && a.z_old != nullptr;
    const bool bulk_in = (rb & 15u) == 0 && (!old_tile || aligned(a.z_old, 16)) && (!forced || aligned(a.z_forced, 16));
    const unsigned int tab_bytes = have_tab ? (unsigned int)TL.copy_floats * 4u : 0u;
    const bool use_bar = !LEAN || have_tab;
    if (use_bar && threadIdx.x == 0) {
        mbar_init(&bar, 1);
        // the waits sit behind the CTA barrier that follows the per-particle prologue, which also publishes the init
        mbar_expect_tx(&bar, tab_bytes + ((!LEAN && bulk_in) ? rb * (unsigned int)rows * ((old_tile ? 1u : 0u) + (forced ? 1u : 0u)) : 0u));
        if (have_tab) bulk_g2s(s_ex, a.tables + (size_t)b * TL.total, tab_bytes, &bar);
    }
    if (!LEAN) {
        __syncthreads();                                           // the particles' own lanes issue their row copies
        if (valid) {
            if (bulk_in) {
                if (j == 0) {
                    if (old_tile) bulk_g2s(s_zo + (size_t)p * NK, a.z_old + rz * NK, rb, &bar);
                    if (forced) bulk_g2s(s_zn + (size_t)p * NK, a.z_forced + r * NK, rb, &bar);
                }
            } else {
                for (int i = j; i < NK; i += LANES) {
                    if (old_tile) s_zo[(size_t)p * NK + i] = __ldg(a.z_old + rz * NK + i);
                    if (forced) s_zn[(size_t)p * NK + i] = __ldg(a.z_forced + r * NK + i);
                }
            }
        }
    }

    if (LEAN && valid) {
        const float* ur = a.u + r * N;
        if (rng) {
            // the uniforms of this lane's own points (read back by the same thread in the epilogue), made here -- while the
            // label rows and the CTA's tables are in flight -- so that the main loop carries no generator state
            static_assert(PB <= 4, "one Philox block carries four uniforms");
            for (int it = 0; it < nIt; ++it) {
#if APGS_ZF_EXPT == 3
                const uint4 w = make_uint4(key.x, key.y, (unsigned int)r, (unsigned int)(it * LANES + j));
#else
                const uint4 w = philox_block(r, it * LANES + j, key);
#endif
                const unsigned int ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                for (int q = 0; q < PB; ++q) {
                    const int n = it * (LANES * PB) + j + q * LANES;
                    if (n < N) s_u[(size_t)p * NU + n] = u01_co(ww[q]);
                }
            }
        } else if ((N & 3) == 0 && aligned(a.u, 16)) {               // rows of N floats: every row start is 16-byte aligned
            for (int c = j; c < N / 4; c += LANES)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(s_u + (size_t)p * NU + c * 4)), "l"(ur + c * 4) : "memory");
        } else {
            for (int i = j; i < N; i += LANES) s_u[(size_t)p * NU + i] = __ldg(ur + i);
        }
        const unsigned char* lr = a.zl_old + rz * a.LN;
        if ((a.LN & 15) == 0 && a.LN >= LS && aligned(a.zl_old, 16)) {
            for (int c = j; c < LS / 16; c += LANES)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(s_lb + (size_t)p * LS + c * 16)), "l"(lr + c * 16) : "memory");
        } else {
            for (int i = j; i < N; i += LANES) s_lb[(size_t)p * LS + i] = __ldg(lr + i);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    }

    // ---- per-CTA tables (no tables handed in): built here, while the rows are in flight
    const float* xb = a.x + (size_t)b * N * D;
    if (!have_tab) zf_cta_tables<K, D, LANES, PB>(xb, a.W1, a.b1, a.W2, a.b2, a.rpi, N, H, s_ex, s_xm, s_aw, s_misc, s_x);
    if constexpr (NSS) {
        // per-CTA table of Enc_apg_eta's head factors over (point, label) (see gmm_common.cuh)
        for (int idx = threadIdx.x; idx < NK; idx += P * LANES) {
            const int n = idx / K, lab = idx - n * K;
            float xin[D];
#pragma unroll
            for (int d = 0; d < D; ++d) xin[d] = __ldg(xb + n * D + d);
            eta_head_factors<K, D>(a.Wg, a.bg, a.Wo, a.bo, xin, lab, s_nt + (size_t)idx * EtaTab<K, D>::RS);
        }
    }
    // ---- per-particle pieces: e''_k[h] = 2 log2(e) (b1 + W1m mu_k + W1t tau_k) - 31, Normal tables.  The Ee table is
    // written for the factorised loop right away (2^e'' = 2^-31 Ee_k[h]); the rare particle that fails the range test
    // below rewrites its entries with the raw e' for the plain loop.
    float elo = INFINITY, ehi = -INFINITY;                     // range of e'' over this lane's entries
    bool in_ok = true;                                         // inputs small enough for every e' to be finite
    float* ep = s_e + (size_t)p * EPB;
    float* em = ep + EPs;
    float tk[KD], mk[KD];
    if (valid) {
#pragma unroll
        for (int e = 0; e < KD; ++e) { tk[e] = __ldg(a.tau + rs * KD + e); mk[e] = __ldg(a.mu + rs * KD + e); }
#pragma unroll
        for (int e = 0; e < KD; ++e) in_ok &= (fabsf(tk[e]) <= ZF_IN_MAX) & (fabsf(mk[e]) <= ZF_IN_MAX);   // false for NaN
        const float* wtp = a.tables + TL.wt;                   // instance 0's copy: the rows are the same for all
#pragma unroll 4
        for (int h = j; h < H4; h += LANES) {
            const bool live = h < H;
            float wm[D], wt[D], bh;
            if (have_tab) {
                const float4 v0 = __ldg(reinterpret_cast<const float4*>(wtp + h * 8));
                const float4 v1 = __ldg(reinterpret_cast<const float4*>(wtp + h * 8 + 4));
                const float v[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
                for (int d = 0; d < D; ++d) { wm[d] = v[d]; wt[d] = v[D + d]; }
                bh = v[2 * D];
            } else {
                zf_wt_row<D>(a.W1, a.b1, h, live, wm, wt, bh);
            }
            float emax = -INFINITY;
#pragma unroll
            for (int k = 0; k < K; ++k) {
#if APGS_ZF_EXPT == 4
                emax = -31.f; elo = -31.f;
                ep[ee_slot<K>(h, k)] = bh * mk[k * D] + wm[0] * tk[k * D];
#else
                float acc = bh;
#pragma unroll
                for (int d = 0; d < D; ++d) {
                    acc = fmaf(wm[d], mk[k * D + d], acc);
                    acc = fmaf(wt[d], tk[k * D + d], acc);
                }
                if (live) { emax = fmaxf(emax, acc); elo = fminf(elo, acc); }
                ep[ee_slot<K>(h, k)] = live ? ex2_ftz(acc) : 0.f;              // 2^-31 Ee_k[h]; padding rows 0
#endif
            }
            em[h] = emax;
            ehi = fmaxf(ehi, emax);
        }
        float* tab = s_tab + (size_t)p * (K * TRS);
        for (int e = j; e < KD; e += LANES) {
            float tf = 0.f, mf = 0.f;
#pragma unroll
            for (int ee = 0; ee < KD; ++ee) if (ee == e) { tf = tk[ee]; mf = mk[ee]; }
            // sigma = 1 / sqrt(tau) (apgs/gmm/models.py:177):  1 / (2 sigma^2) = tau / 2  and  log sigma = -log(tau) / 2, without
            // the IEEE sqrt / two divides / logf of the literal form (~45 instructions per particle-warp): the quotient differs
            // from the reference's rounding chain by <= 2.5 ulp, the SFU log by ~1e-7 absolute -- both far inside the 1e-5 of
            // the summed magnitudes the weights are held to; tau <= 0 or NaN gives NaN / inf exactly where the reference does
            float* row = tab + (e / D) * TRS + (e % D) * 3;
            row[0] = mf;
            row[1] = 0.5f * tf;
            row[2] = fmaf(lg2_ftz(tf), -0.5f * LN2_F, LOG_SQRT_2PI_F);
        }
    }
    if (LEAN) asm volatile("cp.async.wait_group 0;" ::: "memory");   // this thread's uniforms / labels landed
    __syncthreads();                                               // tables built here / mbarrier init / cooperative rows
    if (use_bar) mbar_wait(&bar, 0);                               // tables and rows staged by the copy engine landed
    const float b2 = s_misc[0];
    float logpi[K];
#pragma unroll
    for (int k = 0; k < K; ++k) logpi[k] = s_misc[1 + k];
    float bound = -INFINITY;
    if (valid)
        for (int h = j; h < H; h += LANES) {                                         // each lane reads the em it wrote
            const float xm = s_xm[h];
            bound = fmaxf(bound, (xm <= 126.f) ? xm + em[h] : INFINITY);             // +inf stays +inf whatever em is
        }
#pragma unroll
    for (int o = LANES / 2; o > 0; o >>= 1) {
        bound = fmaxf(bound, __shfl_xor_sync(0xffffffffu, bound, o));
        ehi = fmaxf(ehi, __shfl_xor_sync(0xffffffffu, ehi, o));
        elo = fminf(elo, __shfl_xor_sync(0xffffffffu, elo, o));
    }
    // with e'' = e' - 31:  XM + e' <= 61,  -80 <= e' <= 80;  in_ok is the same for all lanes of a particle
    const bool fast = in_ok && (s_misc[7] != 0.f) && (bound <= 30.f) && (ehi <= 49.f) && (elo >= -111.f);
    if (valid && !fast) {                         // plain loop: raw e' (each lane rewrites the entries it owns)
        for (int h = j; h < H; h += LANES) {
            const float bh = __ldg(a.b1 + h);
#pragma unroll
            for (int k = 0; k < K; ++k) {
                float acc = bh;                   // unscaled weights, one scaling of the sum: no error growth under cancellation
#pragma unroll
                for (int d = 0; d < D; ++d) {
                    acc = fmaf(__ldg(a.W1 + h * 3 * D + D + d), mk[k * D + d], acc);
                    acc = fmaf(__ldg(a.W1 + h * 3 * D + 2 * D + d), tk[k * D + d], acc);
                }
                ep[ee_slot<K>(h, k)] = TANH_PRESCALE * acc;
            }
        }
    }
    __syncwarp();                                                  // a particle's lanes share a warp: publishes s_e

    PlainF A_llf, A_llb, A_lpzf, A_lpzb, A_lqzf, A_lqzb, A_loc;     // <= ceil(N/LANES) terms per lane
    if (valid) {
        const float* zo = s_zo + (size_t)p * NK;
        float* zn = s_zn + (size_t)p * NK;
        const unsigned char* zlo = a.zl_old + rz * a.LN;         // label rows of this particle (label mode)
        unsigned char* zln = a.zl_new + r * a.LN;
        const float* tab = s_tab + (size_t)p * (K * TRS);
        const float b2l = b2 * 1.4426950408889634f;              // logits are carried in log2 units: l' = log2(e) logit
#pragma unroll(NC ? 2 : 1)
        for (int it = 0; it < nIt; ++it) {
            const int n0 = it * (LANES * PB) + j;
            float logit[PB][K], uq[PB];
            int lbq[PB];
#pragma unroll
            for (int q = 0; q < PB; ++q) {
                const int n = n0 + q * LANES;
                if (!LEAN) {
                    uq[q] = (n < N && !forced && !rng) ? __ldg(a.u + r * N + n) : 0.f;  // in flight during the MLP
                    lbq[q] = (n < N && lab_old) ? (int)__ldg(zlo + n) : 0;
                }
            }
            // ---- the MLP: hidden unit h for 4 points x K clusters at a time; points are paired as (0,2) and (1,3)
            if (fast) {
                f32x2 L02[K], L13[K];
#pragma unroll
                for (int k = 0; k < K; ++k) { L02[k] = 0ull; L13[k] = 0ull; }
                const unsigned char* exb = reinterpret_cast<const unsigned char*>(s_ex + (size_t)(it * LANES + j) * PB);
#if APGS_ZF_EXPT != 1
                mlp_quad_fast<K>(exb, EXS * 4, reinterpret_cast<const unsigned char*>(ep),
                                 reinterpret_cast<const unsigned char*>(s_aw), H4, L02, L13);
#endif

#pragma unroll
                for (int k = 0; k < K; ++k) {
                    unpack2(L02[k], logit[0][k], logit[2][k]);
                    unpack2(L13[k], logit[1][k], logit[3][k]);
                }
#pragma unroll
                for (int q = 0; q < PB; ++q)
#pragma unroll
                    for (int k = 0; k < K; ++k) logit[q][k] = fmaf(logit[q][k], 4.656612873077393e-10f * 1.4426950408889634f, b2l);
            } else {
#pragma unroll
                for (int q = 0; q < PB; ++q)
#pragma unroll
                    for (int k = 0; k < K; ++k) logit[q][k] = b2;
                float xq[PB][D];
#pragma unroll
                for (int q = 0; q < PB; ++q)
#pragma unroll
                    for (int d = 0; d < D; ++d) xq[q][d] = (n0 + q * LANES < N) ? s_x[(n0 + q * LANES) * D + d] : 0.f;
                mlp_quad_plain<K, D, PB>(a.W1, xq, ep, s_aw, H, logit);
#pragma unroll
                for (int q = 0; q < PB; ++q)
#pragma unroll
                    for (int k = 0; k < K; ++k) logit[q][k] *= 1.4426950408889634f;
            }
            if (!LEAN && rng) {
                const uint4 w = philox_block(r, it * LANES + j, key);
                const unsigned int ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                for (int q = 0; q < PB; ++q) uq[q] = u01_co(ww[q]);
            }
            // ---- per point: softmax, draw, log-densities
#pragma unroll
            for (int q = 0; q < PB; ++q) {
                const int n = n0 + q * LANES;
                if (n >= N) continue;
#if APGS_ZF_EXPT == 2
                if (LEAN) { A_loc.add(logit[q][0] + logit[q][1] + logit[q][K - 1]); continue; }
#endif
                if (LEAN) { uq[q] = s_u[(size_t)p * NU + n]; lbq[q] = (int)s_lb[(size_t)p * LS + n]; }
                float xn[D];                                          // the point itself: only the log-densities need it
#pragma unroll
                for (int d = 0; d < D; ++d) xn[d] = s_x[n * D + d];
                // softmax in log2 units: e_k = 2^(l'_k - max), q_k = e_k / sum e.  The draw compares u * sum e against the
                // running sums of e (no normalised probabilities needed), and log q_k = ln 2 (l'_k - (max + log2 sum e)):
                // ONE lg2 serves the forward and the backward label.  (The reference normalises twice -- F.softmax, then
                // Categorical's probs / probs.sum(-1), apgs/gmm/models.py:104,115 -- which is this quotient up to rounding.)
                float mx = logit[q][0];
#pragma unroll
                for (int k = 1; k < K; ++k) mx = fmaxf(mx, logit[q][k]);
                float pr[K], se = 0.f;
#pragma unroll
                for (int k = 0; k < K; ++k) { pr[k] = ex2_ftz(logit[q][k] - mx); se = __fadd_rn(se, pr[k]); }
                if (want_q) {
                    const float rse = rcp_ftz(se);
#pragma unroll
                    for (int k = 0; k < K; ++k) a.q_probs[(r * N + n) * K + k] = pr[k] * rse;     // F.softmax(., -1)
                }
                int lf = 0;
                if (forced) {
                    float best = zn[n * K];
#pragma unroll
                    for (int k = 1; k < K; ++k) { const float v = zn[n * K + k]; if (v > best) { best = v; lf = k; } }
                } else {
                    // inverse CDF on the unnormalised probabilities: label = #{k : cdf_k <= u * sum}
                    const float t = __fmul_rn(uq[q], se);
                    float c = 0.f;
#pragma unroll
                    for (int k = 0; k < K - 1; ++k) { c = __fadd_rn(c, pr[k]); lf += (t >= c) ? 1 : 0; }
                    if (hot_new) {
#pragma unroll
                        for (int k = 0; k < K; ++k) zn[n * K + k] = (k == lf) ? 1.f : 0.f;
                    }
                }
                if (LEAN) s_lb[(size_t)p * LS + n] = (unsigned char)lf;      // the old label of this point is already in lbq
                else if (lab_new) zln[n] = (unsigned char)lf;
                const float lnorm = __fadd_rn(mx, lg2_ftz(se));              // log2 of the softmax normaliser
                float slf = logit[q][0], lpf = logpi[0];
#pragma unroll
                for (int k = 1; k < K; ++k) if (lf == k) { slf = logit[q][k]; lpf = logpi[k]; }
                // Categorical.log_prob clamps the probability to [eps, 1 - eps] (torch/distributions/utils.py clamp_probs)
                const float lqf = fminf(fmaxf(__fmul_rn(__fsub_rn(slf, lnorm), LN2_F), LOG_EPS_F), LOG_1M_EPS_F);
                float llf = 0.f;
                {
                    float row[TRS];
#pragma unroll
                    for (int i = 0; i < TRS / 4; ++i)
                        *reinterpret_cast<float4*>(row + 4 * i) = *reinterpret_cast<const float4*>(tab + lf * TRS + 4 * i);
#pragma unroll
                    for (int d = 0; d < D; ++d) {
                        const float df = __fsub_rn(xn[d], row[3 * d]);
                        llf = __fadd_rn(llf, __fsub_rn(-__fmul_rn(__fmul_rn(df, df), row[3 * d + 1]), row[3 * d + 2]));
                    }
                }
                A_llf.add(llf); A_lqzf.add(lqf); A_lpzf.add(lpf);
                if (has_bwd) {
                    int lb = lbq[q];
                    if (!lab_old) {
                        float best = zo[n * K];
#pragma unroll
                        for (int k = 1; k < K; ++k) { const float v = zo[n * K + k]; if (v > best) { best = v; lb = k; } }
                    }
                    float slb = logit[q][0], lpb = logpi[0];
#pragma unroll
                    for (int k = 1; k < K; ++k) if (lb == k) { slb = logit[q][k]; lpb = logpi[k]; }
                    const float lqb = fminf(fmaxf(__fmul_rn(__fsub_rn(slb, lnorm), LN2_F), LOG_EPS_F), LOG_1M_EPS_F);
                    float llb = 0.f;
                    {
                        float row[TRS];
#pragma unroll
                        for (int i = 0; i < TRS / 4; ++i)
                            *reinterpret_cast<float4*>(row + 4 * i) = *reinterpret_cast<const float4*>(tab + lb * TRS + 4 * i);
#pragma unroll
                        for (int d = 0; d < D; ++d) {
                            const float db = __fsub_rn(xn[d], row[3 * d]);
                            llb = __fadd_rn(llb, __fsub_rn(-__fmul_rn(__fmul_rn(db, db), row[3 * d + 1]), row[3 * d + 2]));
                        }
                    }
                    if (!LEAN) { A_llb.add(llb); A_lqzb.add(lqb); A_lpzb.add(lpb); }   // only log_w_b needs these sums
                    const float wf = __fsub_rn(__fadd_rn(llf, lpf), lqf);
                    const float wb = __fsub_rn(__fadd_rn(llb, lpb), lqb);
                    const float loc = __fsub_rn(wf, wb);
                    A_loc.add(loc);
                    if (want_lw) a.lw_local[r * N + n] = loc;
                }
                if (want_lq) a.lq_local[r * N + n] = lqf;
            }
        }
    }
    if (LEAN) {
        // ---- label rows out: the N label bytes of a particle sit in shared memory; its lanes store them as 16-byte
        // vectors (pad bytes of the caller's row are never written)
        __syncwarp();
        if (valid) {
            unsigned char* zln = a.zl_new + r * a.LN;
            const unsigned char* src = s_lb + (size_t)p * LS;
            if ((a.LN & 15) == 0 && aligned(a.zl_new, 16)) {
                for (int off = j * 16; off < N; off += LANES * 16) {
                    int len = N - off;
                    if (len >= 16) {
                        *reinterpret_cast<uint4*>(zln + off) = *reinterpret_cast<const uint4*>(src + off);
                    } else {
                        int o = off;
                        if (len & 8) { *reinterpret_cast<uint2*>(zln + o) = *reinterpret_cast<const uint2*>(src + o); o += 8; }
                        if (len & 4) { *reinterpret_cast<unsigned int*>(zln + o) = *reinterpret_cast<const unsigned int*>(src + o); o += 4; }
                        if (len & 2) { *reinterpret_cast<unsigned short*>(zln + o) = *reinterpret_cast<const unsigned short*>(src + o); o += 2; }
                        if (len & 1) zln[o] = src[o];
                    }
                }
            } else {
                for (int n = j; n < N; n += LANES) zln[n] = src[n];
            }
        }
    }
    if constexpr (NSS) {
        // ---- the next eta step's inputs from the labels just drawn (k_gmm_nss_im's arithmetic in its order: lanes
        // j < kEtaNssLanes own the points n = j, j + kEtaNssLanes, ...; the other lanes contribute zeros)
        EtaNssAcc<K, D, true> acc;
        {
            float cen[K][D];
            const float* tab = s_tab + (size_t)p * (K * TRS);
#pragma unroll
            for (int k = 0; k < K; ++k)
#pragma unroll
                for (int d = 0; d < D; ++d) cen[k][d] = valid ? tab[k * TRS + d * 3] : 0.f;      // this particle's mu
            acc.init(cen);
        }
        __syncwarp();                                  // the labels of the particle's other lanes
        if (valid && j < kEtaNssLanes) {
            for (int n = j; n < N; n += kEtaNssLanes) {
                const int lab = min((int)s_lb[(size_t)p * LS + n], K - 1);
                acc.add(s_nt + (size_t)(n * K + lab) * EtaTab<K, D>::RS, lab, s_x + n * D);
            }
        }
        __syncwarp();                                  // every lane of the particle is through with its Ee / em block
        eta_nss_finish<K, D, LANES, true>(acc, ep, j, valid, r, a.ra, a.rb, a.rm, a.rn, a.oa, a.ob, a.om, a.on, a.stats);
    }
    const double S_llf = group_sum<LANES>(A_llf.value()), S_lpzf = group_sum<LANES>(A_lpzf.value());
    const double S_lqzf = group_sum<LANES>(A_lqzf.value());
    double S_llb = 0.0, S_lpzb = 0.0, S_lqzb = 0.0;                // the lean variant never writes log_w_f / log_w_b
    if (!LEAN) {
        S_llb = group_sum<LANES>(A_llb.value());
        S_lpzb = group_sum<LANES>(A_lpzb.value());
        S_lqzb = group_sum<LANES>(A_lqzb.value());
    }
    const double S_loc = group_sum<LANES>(A_loc.value());
    if (valid && j == 0) {
        const float ll_f = (float)S_llf, ll_b = (float)S_llb;
        const float lpz_f = (float)S_lpzf, lpz_b = (float)S_lpzb, lqz_f = (float)S_lqzf, lqz_b = (float)S_lqzb;
        if (a.log_w != nullptr) a.log_w[r] = (float)S_loc;
        if (!LEAN && a.log_w_f != nullptr) a.log_w_f[r] = __fsub_rn(__fadd_rn(ll_f, lpz_f), lqz_f);
        if (!LEAN && a.log_w_b != nullptr) a.log_w_b[r] = __fsub_rn(__fadd_rn(ll_b, lpz_b), lqz_b);
        if (a.log_q_f != nullptr) a.log_q_f[r] = lqz_f;
        if (a.dens != nullptr) a.dens[r] = __fadd_rn(ll_f, lpz_f);
    }
    // ---- z_new rows out (one bulk store per particle row); nothing to do in label mode
    if (LEAN || a.z_new == nullptr) return;
    const bool bulk_out = (rb & 15u) == 0 && aligned(a.z_new, 16);
    if (bulk_out) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> async proxy
        __syncthreads();
        if (valid && j == 0) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(a.z_new + r * NK),
                         "r"(smem_u32(s_zn + (size_t)p * NK)), "r"(rb)
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // smem may be released at exit
        }
    } else {
        __syncthreads();
        if (valid)
            for (int i = j; i < NK; i += LANES) a.z_new[r * NK + i] = s_zn[(size_t)p * NK + i];
    }
}

// 32 particles per CTA (3 CTAs/SM).  Measured alternatives on the lean launch: 16 particles (6 CTAs/SM, phases less
// correlated, twice the Ex-table work) 1.23 ms vs 1.20; tau / mu loads hoisted above the Ex table 1.22 ms; a CTA that
// walks four chunks of its instance (one Ex table, no CTA barrier after the first chunk) 1.21 ms.
#ifndef APGS_ZF_P
#define APGS_ZF_P 16
#endif
constexpr int kFP = APGS_ZF_P, kFLanes = 8, kFPB = 4;

template <int K, int D, bool NSS>
static int launch_zfused_lean(const ZFusedArgs& a, cudaStream_t st, size_t smem_used) {
    auto kern = a.H == 32 ? (a.N == 60 ? (a.tables ? k_gmm_z_fused<K, D, kFP, kFLanes, kFPB, true, 32, 60, NSS, 1>
                                                   : k_gmm_z_fused<K, D, kFP, kFLanes, kFPB, true, 32, 60, NSS, 0>)
                                       : k_gmm_z_fused<K, D, kFP, kFLanes, kFPB, true, 32, 0, NSS>)
                          : k_gmm_z_fused<K, D, kFP, kFLanes, kFPB, true, 0, 0, NSS>;
    APGS_CUDA_OK(ensure_dynamic_smem(reinterpret_cast<const void*>(kern), smem_used), "smem attr");
    const int chunks = (a.S + kFP - 1) / kFP;
    kern<<<dim3((unsigned int)a.B, (unsigned int)chunks), kFP * kFLanes, smem_used, st>>>(a);
    APGS_LAUNCH_OK("k_gmm_z_fused launch");
    return 0;
}

template <int K, int D>
static int launch_zfused(const ZFusedArgs& a, cudaStream_t st) {
    const int H4 = (a.H + 3) / 4 * 4;
    const bool tiles = a.z_old != nullptr || a.z_forced != nullptr || a.z_new != nullptr;
    const bool nss = a.oa != nullptr;
    const bool lean = a.zl_old && a.zl_new && (a.u || a.seed) && !tiles && !a.q_probs && !a.lw_local && !a.lq_local &&
                      !a.log_w_f && !a.log_w_b;
    if (nss && !lean) return fail(APGS_E_BADARG, "gmm_z_fused: the eta statistics ride on the lean label-mode launch only");
    constexpr bool kNssOk = K * D <= kFLanes;
    if (nss && !kNssOk) return fail(APGS_E_UNSUPPORTED, "gmm_z_fused: eta statistics need K*D <= %d", kFLanes);
    const size_t tile_f = tiles ? ((size_t)kFP * a.N * K + 3) / 4 * 4 : 0;
    const int EPB = (nss && kNssOk) ? zf_block_floats<K, D, kFLanes, kNssOk>(H4) : zf_block_floats<K, D, kFLanes, false>(H4);
    const ZTabLayout TL = zf_tab_layout(a.N, H4, D);
    if (a.tables && !aligned(a.tables, 16)) return fail(APGS_E_BADARG, "gmm_z_fused: tables must be 16-byte aligned");
    const size_t smem = (2 * tile_f + (size_t)TL.copy_floats + (size_t)kFP * EPB + (size_t)kFP * K * ((3 * D + 3) / 4 * 4)) * 4;
    const size_t smem_lean = smem + (size_t)kFP * ((a.N + 3) / 4 * 4) * 4 + (size_t)kFP * ((a.N + 15) / 16 * 16) +
                             (nss ? (size_t)a.N * K * EtaTab<K, D>::RS * 4 : 0);
    const size_t smem_used = lean ? smem_lean : smem;
    // dynamic + static (mbarrier) shared memory against the opt-in limit of one CTA (227 KB on sm_100)
    if (smem_used + 256 > 200 * 1024) {
        if (lean && !nss && smem + 256 <= 200 * 1024) {
            // the lean variant's staging of uniforms / labels does not fit: the generic variant (same numbers) does
            APGS_CUDA_OK(ensure_dynamic_smem(reinterpret_cast<const void*>(k_gmm_z_fused<K, D, kFP, kFLanes, kFPB, false>), smem), "smem attr");
            const int chunks = (a.S + kFP - 1) / kFP;
            if (chunks > 65535) return fail(APGS_E_SHAPE, "gmm_z_fused: S=%d too large for one launch", a.S);
            k_gmm_z_fused<K, D, kFP, kFLanes, kFPB, false><<<dim3((unsigned int)a.B, (unsigned int)chunks), kFP * kFLanes, smem, st>>>(a);
            APGS_LAUNCH_OK("k_gmm_z_fused launch");
            return 0;
        }
        return fail(APGS_E_SHAPE, "gmm_z_fused: N=%d H=%d too large for one tile", a.N, a.H);
    }
    const int chunks = (a.S + kFP - 1) / kFP;
    if (chunks > 65535) return fail(APGS_E_SHAPE, "gmm_z_fused: S=%d too large for one launch", a.S);
    if (lean) {
        if constexpr (kNssOk) {
            if (nss) return launch_zfused_lean<K, D, true>(a, st, smem_used);
        }
        return launch_zfused_lean<K, D, false>(a, st, smem_used);
    }
    auto kern = (a.H == 32 && a.N == 60 ? k_gmm_z_fused<K, D, kFP, kFLanes, kFPB, false, 32, 60>
                                        : k_gmm_z_fused<K, D, kFP, kFLanes, kFPB, false>);
    APGS_CUDA_OK(ensure_dynamic_smem(reinterpret_cast<const void*>(kern), smem_used), "smem attr");
    kern<<<dim3((unsigned int)a.B, (unsigned int)chunks), kFP * kFLanes, smem_used, st>>>(a);
    APGS_LAUNCH_OK("k_gmm_z_fused launch");
    return 0;
}

}  // namespace apgs

using namespace apgs;

static int z_step_fused(const float* x, const float* tau, const float* mu, const float* W1, const float* b1,
                        const float* W2, const float* b2, int64_t H, const apgs_gmm_prior_t* prior, const float* z_old,
                        const unsigned char* z_old_lab, const float* z_forced, const float* uniforms, int64_t S,
                        int64_t B, int64_t N, int64_t K, int64_t D, float* z_new, unsigned char* z_new_lab,
                        int64_t label_stride, float* q_probs, const apgs_gmm_out_t* out, void* stream,
                        const float* const* heads = nullptr, const apgs_gmm_post_out_t* post_out = nullptr,
                        float* stats_out = nullptr, const uint64_t* philox_seed = nullptr, const float* tables = nullptr,
                        const int32_t* src_ancestors = nullptr, const int32_t* src_ancestors_z = nullptr) {
    if (S <= 0 || B <= 0 || N <= 0 || H <= 0 || H > 1024)
        return fail(APGS_E_SHAPE, "gmm_z_step_fused: bad sizes S=%lld B=%lld N=%lld H=%lld", (long long)S, (long long)B,
                    (long long)N, (long long)H);
    if (!x || !tau || !mu || !W1 || !b1 || !W2 || !b2 || !prior || !prior->pi || (!z_new && !z_new_lab))
        return fail(APGS_E_BADARG, "gmm_z_step_fused: NULL input");
    if (!z_forced && !uniforms && !philox_seed) return fail(APGS_E_BADARG, "gmm_z_step_fused: uniforms (or z_forced) required");
    if (z_old && z_old_lab) return fail(APGS_E_BADARG, "gmm_z_step_fused: give z_old either one-hot or as labels, not both");
    if (K > 255) return fail(APGS_E_SHAPE, "gmm_z_step_fused: K=%lld does not fit a label byte", (long long)K);
    ZFusedArgs a = {};
    a.x = x; a.tau = tau; a.mu = mu; a.W1 = W1; a.b1 = b1; a.W2 = W2; a.b2 = b2;
    a.z_old = z_old; a.zl_old = z_old_lab; a.z_forced = z_forced; a.u = uniforms; a.rpi = prior->pi;
    a.seed = reinterpret_cast<const unsigned long long*>(philox_seed);
    a.tables = tables;
    a.src_anc = src_ancestors;
    a.src_anc_z = src_ancestors_z;
    a.S = (int)S; a.B = (int)B; a.N = (int)N; a.H = (int)H;
    a.z_new = z_new; a.zl_new = z_new_lab; a.LN = label_stride; a.q_probs = q_probs;
    if (out) {
        a.log_w = out->log_w; a.log_q_f = out->log_q_f; a.dens = out->density; a.log_w_f = out->log_w_f;
        a.log_w_b = out->log_w_b; a.lw_local = out->log_w_local; a.lq_local = out->log_q_local;
    }
    if (post_out) {
        if (!heads || !heads[0] || !heads[1] || !heads[2] || !heads[3] || !stats_out || !post_out->alpha || !post_out->beta ||
            !post_out->mu || !post_out->nu || !prior->alpha || !prior->beta || !prior->mu || !prior->nu)
            return fail(APGS_E_BADARG, "gmm_z_step_fused: eta statistics need the heads, the Normal-Gamma prior and all outputs");
        a.Wg = heads[0]; a.bg = heads[1]; a.Wo = heads[2]; a.bo = heads[3];
        a.ra = prior->alpha; a.rb = prior->beta; a.rm = prior->mu; a.rn = prior->nu;
        a.oa = post_out->alpha; a.ob = post_out->beta; a.om = post_out->mu; a.on = post_out->nu;
        a.stats = stats_out;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (K == 3 && D == 2) return launch_zfused<3, 2>(a, st);
    if (K == 2 && D == 2) return launch_zfused<2, 2>(a, st);
    if (K == 4 && D == 2) return launch_zfused<4, 2>(a, st);
    if (K == 5 && D == 2) return launch_zfused<5, 2>(a, st);
    if (K == 3 && D == 3) return launch_zfused<3, 3>(a, st);
    return fail(APGS_E_UNSUPPORTED, "gmm_z_step_fused: (K=%lld, D=%lld) not instantiated", (long long)K, (long long)D);
}

// Measurement only (bench.py: roofline.mlp_loop_alone): the main loop of the fused z step -- mlp_quad_fast itself, K = 3,
// H = 32, 32 particles x 8 lanes per CTA as in the step -- on synthetic shared-memory tables, with no prologue and no
// epilogue.  A CTA runs `passes` passes of eight quads of hidden units; S*B/32 CTAs x 2 passes is the loop work of one
// launch of the step at N <= 64.  What it shows: packed FFMA2 / FMUL2 take TWO issue slots each, so the loop (84 packed
// + 27 other instructions per quad) needs ~212 cycles per quad and scheduler against 168 of FMA-pipe time.
__global__ void __launch_bounds__(256, 4) k_gmm_z_loop_probe(float* out, int passes) {
    constexpr int K = 3, H4 = 32, EXS = 64, EPB = 132;
    __shared__ __align__(16) float s_ex[H4 * EXS];
    __shared__ __align__(16) float s_aw[H4];
    __shared__ __align__(16) float s_e[32 * EPB];
    for (int i = threadIdx.x; i < H4 * EXS; i += 256) s_ex[i] = 1.f + 1e-3f * (float)(i & 15);
    for (int i = threadIdx.x; i < H4; i += 256) s_aw[i] = 0.01f * (float)(i + 1);
    for (int i = threadIdx.x; i < 32 * EPB; i += 256) s_e[i] = 4.6e-10f * (1.f + 1e-3f * (float)(i & 7));
    __syncthreads();
    const int p = threadIdx.x / 8, j = threadIdx.x % 8;
    f32x2 L02[K], L13[K];
#pragma unroll
    for (int k = 0; k < K; ++k) { L02[k] = 0ull; L13[k] = 0ull; }
    for (int r = 0; r < passes; ++r)
        mlp_quad_fast<K>(reinterpret_cast<const unsigned char*>(s_ex + (size_t)((r & 1) * 8 + j) * 4), EXS * 4,
                         reinterpret_cast<const unsigned char*>(s_e + p * EPB), reinterpret_cast<const unsigned char*>(s_aw), H4, L02, L13);
    float acc = 0.f, lo, hi;
#pragma unroll
    for (int k = 0; k < K; ++k) { unpack2(L02[k], lo, hi); acc += lo + hi; unpack2(L13[k], lo, hi); acc += lo + hi; }
    out[(size_t)blockIdx.x * 256 + threadIdx.x] = acc;
}

template <int K, int D>
static int launch_ztables(const float* x, const float* W1, const float* b1, const float* W2, const float* b2,
                          const float* rpi, int B, int N, int H, float* tables, cudaStream_t st) {
    const ZTabLayout L = zf_tab_layout(N, (H + 3) / 4 * 4, D);
    const size_t smem = (size_t)L.total * 4;
    if (smem + 256 > 200 * 1024) return fail(APGS_E_SHAPE, "gmm_z_tables: N=%d H=%d too large for one tile", N, H);
    APGS_CUDA_OK(ensure_dynamic_smem(reinterpret_cast<const void*>(k_gmm_z_tables<K, D>), smem), "smem attr");
    k_gmm_z_tables<K, D><<<(unsigned int)B, 256, smem, st>>>(x, W1, b1, W2, b2, rpi, N, H, tables);
    APGS_LAUNCH_OK("k_gmm_z_tables launch");
    return 0;
}

extern "C" int apgs_gmm_z_loop_probe_f32(int64_t ctas, int64_t passes, float* out, void* stream) {
    if (ctas <= 0 || ctas > 0x7fffffff || passes <= 0 || passes > (1 << 20) || !out)
        return fail(APGS_E_BADARG, "gmm_z_loop_probe: bad arguments");
    k_gmm_z_loop_probe<<<(unsigned int)ctas, 256, 0, (cudaStream_t)stream>>>(out, (int)passes);
    APGS_LAUNCH_OK("k_gmm_z_loop_probe launch");
    return 0;
}

extern "C" size_t apgs_gmm_z_tables_bytes(int64_t B, int64_t N, int64_t H, int64_t K, int64_t D) {
    if (B <= 0 || N <= 0 || H <= 0 || H > 1024 || K <= 0 || D <= 0) return 0;
    return (size_t)B * (size_t)zf_tab_layout((int)N, (int)((H + 3) / 4 * 4), (int)D).total * 4;
}

extern "C" int apgs_gmm_z_tables_f32(const float* x, const float* W1, const float* b1, const float* W2, const float* b2,
                                     int64_t H, const apgs_gmm_prior_t* prior, int64_t B, int64_t N, int64_t K, int64_t D,
                                     float* tables, void* stream) {
    if (B <= 0 || N <= 0 || H <= 0 || H > 1024)
        return fail(APGS_E_SHAPE, "gmm_z_tables: bad sizes B=%lld N=%lld H=%lld", (long long)B, (long long)N, (long long)H);
    if (!x || !W1 || !b1 || !W2 || !b2 || !prior || !prior->pi || !tables) return fail(APGS_E_BADARG, "gmm_z_tables: NULL input");
    if (!aligned(tables, 16)) return fail(APGS_E_BADARG, "gmm_z_tables: tables must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    if (K == 3 && D == 2) return launch_ztables<3, 2>(x, W1, b1, W2, b2, prior->pi, (int)B, (int)N, (int)H, tables, st);
    if (K == 2 && D == 2) return launch_ztables<2, 2>(x, W1, b1, W2, b2, prior->pi, (int)B, (int)N, (int)H, tables, st);
    if (K == 4 && D == 2) return launch_ztables<4, 2>(x, W1, b1, W2, b2, prior->pi, (int)B, (int)N, (int)H, tables, st);
    if (K == 5 && D == 2) return launch_ztables<5, 2>(x, W1, b1, W2, b2, prior->pi, (int)B, (int)N, (int)H, tables, st);
    if (K == 3 && D == 3) return launch_ztables<3, 3>(x, W1, b1, W2, b2, prior->pi, (int)B, (int)N, (int)H, tables, st);
    return fail(APGS_E_UNSUPPORTED, "gmm_z_tables: (K=%lld, D=%lld) not instantiated", (long long)K, (long long)D);
}

extern "C" int apgs_gmm_z_step_fused_f32(const float* x, const float* tau, const float* mu, const float* W1,
                                         const float* b1, const float* W2, const float* b2, int64_t H,
                                         const apgs_gmm_prior_t* prior, const float* z_old, const float* z_forced,
                                         const float* uniforms, int64_t S, int64_t B, int64_t N, int64_t K, int64_t D,
                                         float* z_new, float* q_probs, const apgs_gmm_out_t* out, void* stream) {
    return z_step_fused(x, tau, mu, W1, b1, W2, b2, H, prior, z_old, nullptr, z_forced, uniforms, S, B, N, K, D, z_new,
                        nullptr, N, q_probs, out, stream);
}

extern "C" int apgs_gmm_z_step_fused_labels_f32(const float* x, const float* tau, const float* mu, const float* W1,
                                                const float* b1, const float* W2, const float* b2, int64_t H,
                                                const apgs_gmm_prior_t* prior, const uint8_t* z_old_labels,
                                                const float* uniforms, int64_t S, int64_t B, int64_t N, int64_t K,
                                                int64_t D, uint8_t* z_new_labels, int64_t label_stride,
                                                float* z_new_onehot, float* q_probs, const apgs_gmm_out_t* out,
                                                void* stream) {
    if (!z_new_labels) return fail(APGS_E_BADARG, "gmm_z_step_fused_labels: z_new_labels is NULL");
    if (label_stride < N) return fail(APGS_E_BADARG, "gmm_z_step_fused_labels: label_stride %lld < N", (long long)label_stride);
    return z_step_fused(x, tau, mu, W1, b1, W2, b2, H, prior, nullptr, z_old_labels, nullptr, uniforms, S, B, N, K, D,
                        z_new_onehot, z_new_labels, label_stride, q_probs, out, stream);
}

extern "C" int apgs_gmm_z_step_fused_labels_eta_f32(const float* x, const float* tau, const float* mu, const float* W1,
                                                    const float* b1, const float* W2, const float* b2, int64_t H,
                                                    const apgs_gmm_prior_t* prior, const uint8_t* z_old_labels,
                                                    const float* uniforms, int64_t S, int64_t B, int64_t N, int64_t K,
                                                    int64_t D, uint8_t* z_new_labels, int64_t label_stride,
                                                    const float* Wg, const float* bg, const float* Wo, const float* bo,
                                                    const apgs_gmm_post_out_t* post_out, float* stats_out,
                                                    const apgs_gmm_out_t* out, void* stream) {
    if (!z_new_labels || !z_old_labels || !uniforms) return fail(APGS_E_BADARG, "gmm_z_step_fused_labels_eta: labels in / out and uniforms are required");
    if (label_stride < N) return fail(APGS_E_BADARG, "gmm_z_step_fused_labels_eta: label_stride %lld < N", (long long)label_stride);
    if (!post_out || !stats_out) return fail(APGS_E_BADARG, "gmm_z_step_fused_labels_eta: post_out / stats_out is NULL");
    if (out && (out->log_w_f || out->log_w_b || out->log_w_local || out->log_q_local))
        return fail(APGS_E_BADARG, "gmm_z_step_fused_labels_eta: only log_w / log_q_f / density are available on this launch");
    const float* heads[4] = {Wg, bg, Wo, bo};
    return z_step_fused(x, tau, mu, W1, b1, W2, b2, H, prior, nullptr, z_old_labels, nullptr, uniforms, S, B, N, K, D,
                        nullptr, z_new_labels, label_stride, nullptr, out, stream, heads, post_out, stats_out);
}

extern "C" int apgs_gmm_z_step_fused_tab_f32(const float* x, const float* tau, const float* mu, const float* W1,
                                             const float* b1, const float* W2, const float* b2, int64_t H,
                                             const apgs_gmm_prior_t* prior, const float* z_old_onehot,
                                             const uint8_t* z_old_labels, const uint64_t* philox_seed, int64_t S, int64_t B,
                                             int64_t N, int64_t K, int64_t D, float* z_new_onehot, uint8_t* z_new_labels,
                                             int64_t label_stride, float* q_probs, const float* Wg, const float* bg,
                                             const float* Wo, const float* bo, const apgs_gmm_post_out_t* post_out,
                                             float* stats_out, const apgs_gmm_out_t* out, const float* tables,
                                             const int32_t* src_ancestors, const int32_t* src_ancestors_z, void* stream) {
    if (!philox_seed) return fail(APGS_E_BADARG, "gmm_z_step_fused_tab: philox_seed is NULL");
    if (z_new_labels && label_stride < N) return fail(APGS_E_BADARG, "gmm_z_step_fused_tab: label_stride %lld < N", (long long)label_stride);
    if (post_out || stats_out) {
        if (!post_out || !stats_out || !z_old_labels || !z_new_labels || z_new_onehot || q_probs)
            return fail(APGS_E_BADARG, "gmm_z_step_fused_tab: the eta statistics ride on the labels-in / labels-out launch only");
        if (out && (out->log_w_f || out->log_w_b || out->log_w_local || out->log_q_local))
            return fail(APGS_E_BADARG, "gmm_z_step_fused_tab: only log_w / log_q_f / density are available with the eta statistics");
    }
    const float* heads[4] = {Wg, bg, Wo, bo};
    return z_step_fused(x, tau, mu, W1, b1, W2, b2, H, prior, z_old_onehot, z_old_labels, nullptr, nullptr, S, B, N, K, D,
                        z_new_onehot, z_new_labels, z_new_labels ? label_stride : N, q_probs, out, stream,
                        post_out ? heads : nullptr, post_out, stats_out, philox_seed, tables, src_ancestors, src_ancestors_z);
}

extern "C" int apgs_gmm_z_step_fused_rng_f32(const float* x, const float* tau, const float* mu, const float* W1,
                                             const float* b1, const float* W2, const float* b2, int64_t H,
                                             const apgs_gmm_prior_t* prior, const float* z_old_onehot,
                                             const uint8_t* z_old_labels, const uint64_t* philox_seed, int64_t S, int64_t B,
                                             int64_t N, int64_t K, int64_t D, float* z_new_onehot, uint8_t* z_new_labels,
                                             int64_t label_stride, float* q_probs, const float* Wg, const float* bg,
                                             const float* Wo, const float* bo, const apgs_gmm_post_out_t* post_out,
                                             float* stats_out, const apgs_gmm_out_t* out, void* stream) {
    if (!philox_seed) return fail(APGS_E_BADARG, "gmm_z_step_fused_rng: philox_seed is NULL");
    return apgs_gmm_z_step_fused_tab_f32(x, tau, mu, W1, b1, W2, b2, H, prior, z_old_onehot, z_old_labels, philox_seed, S, B, N, K,
                                         D, z_new_onehot, z_new_labels, label_stride, q_probs, Wg, bg, Wo, bo, post_out,
                                         stats_out, out, nullptr, nullptr, nullptr, stream);
}
