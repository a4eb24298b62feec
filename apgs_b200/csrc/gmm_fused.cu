// Fused GMM z-step for sm_100a (SURVEY.md section 8f row N1 + rows G3/W2):
//   Enc_apg_z.forward (apgs/gmm/models.py:102-115): per (point n, cluster k) MLP
//       logit_nk = W2 . tanh(W1 [x_n ; mu_k ; tau_k] + b1) + b2,   q_n = softmax_k(logit_n)
//   + categorical draw of z_n by inverse CDF from a supplied uniform (RNG stays in PyTorch)
//   + the z-step importance weight (apgs/gmm/objectives.py:197-227) against z_old
// in ONE pass over the particles: reads x (per instance), tau/mu (48 B), z_old (720 B) and the uniforms
// (240 B) per particle and writes z_new (720 B) + a few scalars, instead of PyTorch's ~100 KB of
// (S,B,N,K,H) activations per particle.  The first Linear is split into its x / mu / tau column blocks:
//   W1 [x;mu_k;tau_k] + b1 = (W1x x_n) + (W1m mu_k + W1t tau_k + b1) =: xh_n + e_k
// and, because 2^(a+b) = 2^a 2^b, the exponential inside tanh(u) = 1 - 2/(2^(2u log2 e) + 1) factorises too:
//   2^(xh'_n + e'_k) = Ex_n * Ee_k,   Ex_n = 2^xh'_n (one ex2 per point and hidden unit, shared by the K clusters),
//   Ee_k = 2^e'_k (K*H values per particle, computed once in the prologue and kept in shared memory).
// The inner loop is then  y = fma(Ex, Ee, 1) -> clamp -> batch inversion of four y -> fma  per tanh:
// ~7 instructions and 0.6 SFU operations instead of 13 / 2.  Compute-bound (issue + SFU), not HBM-bound.
#include "gmm_common.cuh"

namespace apgs {

struct ZFusedArgs {
    const float* x;                 // (B, N, D)
    const float *tau, *mu;          // (R, K, D)
    const float *W1, *b1, *W2, *b2; // (H, 3D), (H), (H), (1)
    const float* z_old;             // (R, N, K) one-hot, or null (sample-only mode)
    const float* z_forced;          // (R, N, K) one-hot to use instead of sampling, or null
    const float* u;                 // (R, N) uniforms in [0,1), required unless z_forced
    const float* rpi;               // (K)
    long long R;
    int B, N, H;
    float* z_new;                   // (R, N, K) out
    float* q_probs;                 // (R, N, K) out, optional
    float *log_w, *log_q_f, *dens, *log_w_f, *log_w_b;   // (R), optional
    float *lw_local, *lq_local;                           // (R, N), optional
};

template <int K, int D, int P, int LANES, int PB>
__global__ void __launch_bounds__(P* LANES, 3) k_gmm_z_fused(ZFusedArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ unsigned long long bar;
    constexpr int KD = K * D;
    const int N = a.N, H = a.H;
    const int EP = K * H + 1;                                  // padded per-particle stride of e
    const bool has_bwd = a.z_old != nullptr;
    const bool forced = a.z_forced != nullptr;
    const size_t tile_f = (size_t)P * N * K;
    float* s_zo = reinterpret_cast<float*>(smem);              // z_old tile (or z_forced when !has_bwd)
    float* s_zn = s_zo + tile_f;                               // z_new tile (staged in from z_forced, out to z_new)
    float* s_w1x = s_zn + tile_f;                              // H*D
    float* s_w2 = s_w1x + H * D;                               // H
    float* s_e = s_w2 + H;                                     // P * EP
    float* s_tab = s_e + (size_t)P * EP;                       // P * KD * 3: mu, 1/(2 var), log sigma + c
    const long long r0 = (long long)blockIdx.x * P;
    const int rows = (int)min((long long)P, a.R - r0);
    if (threadIdx.x == 0) mbar_init(&bar, 1);
    __syncthreads();
    const unsigned int tb = (unsigned int)(rows * N * K * 4);
    StageReq req[2] = {{s_zo, (has_bwd ? a.z_old : a.x) + (has_bwd ? r0 * N * K : 0), has_bwd ? tb : 0u},
                       {s_zn, (forced ? a.z_forced : a.x) + (forced ? r0 * N * K : 0), forced ? tb : 0u}};
    stage_issue(req, 2, &bar);

    // ---- weights and the per-particle pieces, while the tiles are in flight
    // W1x and e are stored pre-multiplied by 2*log2(e): the tanh argument needs no further scaling
    for (int i = threadIdx.x; i < H * D; i += blockDim.x)
        s_w1x[i] = TANH_PRESCALE * __ldg(a.W1 + (i / D) * (3 * D) + (i % D));
    // tanh = 1 - 2/y  =>  sum_h w2_h tanh_h = (sum_h w2_h) - sum_h (2 w2_h) / y_h : the loop accumulates the second sum
    for (int i = threadIdx.x; i < H; i += blockDim.x) s_w2[i] = -2.f * __ldg(a.W2 + i);
    float b2 = __ldg(a.b2);
    for (int i = 0; i < H; ++i) b2 += __ldg(a.W2 + i);
    const int p = threadIdx.x / LANES, j = threadIdx.x % LANES;
    const bool valid = p < rows;
    const long long r = r0 + p;
    if (valid) {
        float tk[KD], mk[KD];
#pragma unroll
        for (int e = 0; e < KD; ++e) { tk[e] = __ldg(a.tau + r * KD + e); mk[e] = __ldg(a.mu + r * KD + e); }
#pragma unroll
        for (int k = 0; k < K; ++k) {                              // e[k][h] = b1 + W1m mu_k + W1t tau_k
            for (int h = j; h < H; h += LANES) {
                float acc = __ldg(a.b1 + h);
#pragma unroll
                for (int d = 0; d < D; ++d) {
                    acc = fmaf(__ldg(a.W1 + h * 3 * D + D + d), mk[k * D + d], acc);
                    acc = fmaf(__ldg(a.W1 + h * 3 * D + 2 * D + d), tk[k * D + d], acc);
                }
                s_e[p * EP + k * H + h] = ex2_ftz(fminf(fmaxf(TANH_PRESCALE * acc, -80.f), 80.f));   // Ee_k[h]
            }
        }
        float* tab = s_tab + (size_t)p * (KD * 3);
        for (int e = j; e < KD; e += LANES) {
            float tf = 0.f, mf = 0.f;
#pragma unroll
            for (int ee = 0; ee < KD; ++ee) if (ee == e) { tf = tk[ee]; mf = mk[ee]; }
            const float sg = inv_sqrt_ref(tf);
            tab[e * 3 + 0] = mf;
            tab[e * 3 + 1] = __fdiv_rn(1.f, __fmul_rn(2.f, __fmul_rn(sg, sg)));
            tab[e * 3 + 2] = __fadd_rn(logf(sg), LOG_SQRT_2PI_F);
        }
    }
    float logpi[K];
    {
        float sp = 0.f;
#pragma unroll
        for (int k = 0; k < K; ++k) sp = __fadd_rn(sp, __ldg(a.rpi + k));
#pragma unroll
        for (int k = 0; k < K; ++k)
            logpi[k] = logf(fminf(fmaxf(__fdiv_rn(__ldg(a.rpi + k), sp), FLT_EPS_F), 1.f - FLT_EPS_F));
    }
    stage_wait(&bar, 0);

    PlainF A_llf, A_llb, A_lpzf, A_lpzb, A_lqzf, A_lqzb, A_loc;     // <= ceil(N/LANES) terms per lane
    if (valid) {
        const float* zo = s_zo + (size_t)p * N * K;
        float* zn = s_zn + (size_t)p * N * K;
        const float* ev = s_e + p * EP;
        const float* tab = s_tab + (size_t)p * (KD * 3);
        const float* xb = a.x + (size_t)(r % a.B) * N * D;
        for (int n0 = j; n0 < N; n0 += LANES * PB) {
            float xv[PB][D], logit[PB][K];
#pragma unroll
            for (int q = 0; q < PB; ++q) {
                const int n = n0 + q * LANES;
#pragma unroll
                for (int d = 0; d < D; ++d) xv[q][d] = (n < N) ? __ldg(xb + n * D + d) : 0.f;
#pragma unroll
                for (int k = 0; k < K; ++k) logit[q][k] = b2;
            }
            // ---- the MLP: hidden unit h for PB points x K clusters at a time
            if (PB == 4) {
                // points are paired as (0,2) and (1,3) in packed fp32x2 registers: y02 = Ex02 * Ee + 1 etc.
                f32x2 L02[K], L13[K];
#pragma unroll
                for (int k = 0; k < K; ++k) { L02[k] = pack2(b2, b2); L13[k] = L02[k]; }
                f32x2 xd02[D], xd13[D];
#pragma unroll
                for (int d = 0; d < D; ++d) { xd02[d] = pack2(xv[0][d], xv[2][d]); xd13[d] = pack2(xv[1][d], xv[3][d]); }
                const f32x2 one2 = pack2(1.f, 1.f), zero2 = pack2(0.f, 0.f);
                for (int h = 0; h < H; ++h) {
                    f32x2 xh02 = zero2, xh13 = zero2;
#pragma unroll
                    for (int d = 0; d < D; ++d) {
                        const float w = s_w1x[h * D + d];
                        const f32x2 w2p = pack2(w, w);
                        xh02 = fma2(w2p, xd02[d], xh02);
                        xh13 = fma2(w2p, xd13[d], xh13);
                    }
                    float x0, x1, x2, x3;
                    unpack2(xh02, x0, x2);
                    unpack2(xh13, x1, x3);
                    // Ex_n[h] = 2^xh'; 0 / +inf at the extremes are fine: Ee is clamped to [2^-80, 2^80], so no 0 * inf
                    const f32x2 ex02 = pack2(ex2_ftz(x0), ex2_ftz(x2)), ex13 = pack2(ex2_ftz(x1), ex2_ftz(x3));
                    const float wm2 = s_w2[h];                                // -2 * W2[h]
                    const f32x2 wm2p = pack2(wm2, wm2);
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        const float ee = ev[k * H + h];                       // Ee_k[h]
                        const f32x2 eep = pack2(ee, ee);
                        // 1/y for four points with ONE reciprocal (batch inversion).  y is clamped at 2^30 so the
                        // product of four stays below 2^121; 2/(2^30+1) < eps/2, i.e. tanh is already 1.0f there.
                        float y0, y1, y2, y3;
                        unpack2(fma2(ex02, eep, one2), y0, y2);
                        unpack2(fma2(ex13, eep, one2), y1, y3);
                        y0 = fminf(y0, 1073741824.f); y1 = fminf(y1, 1073741824.f);
                        y2 = fminf(y2, 1073741824.f); y3 = fminf(y3, 1073741824.f);
                        const f32x2 yA = pack2(y0, y2), yB = pack2(y1, y3);
                        float p01, p23;
                        unpack2(mul2(yA, yB), p01, p23);
                        const float r = rcp_ftz(p01 * p23);
                        const f32x2 rr = pack2(r * p23, r * p01);             // (1/(y0 y1), 1/(y2 y3))
                        L02[k] = fma2(wm2p, mul2(rr, yB), L02[k]);            // 1/y0 = y1/(y0 y1), 1/y2 = y3/(y2 y3)
                        L13[k] = fma2(wm2p, mul2(rr, yA), L13[k]);
                    }
                }
#pragma unroll
                for (int k = 0; k < K; ++k) { unpack2(L02[k], logit[0][k], logit[2][k]); unpack2(L13[k], logit[1][k], logit[3][k]); }
            } else {
                for (int h = 0; h < H; ++h) {
                    const float wm2 = s_w2[h];
#pragma unroll
                    for (int q = 0; q < PB; ++q) {
                        float acc = 0.f;
#pragma unroll
                        for (int d = 0; d < D; ++d) acc = fmaf(s_w1x[h * D + d], xv[q][d], acc);
                        const float ex = ex2_ftz(acc);
#pragma unroll
                        for (int k = 0; k < K; ++k)
                            logit[q][k] = fmaf(wm2, rcp_ftz(fminf(fmaf(ex, ev[k * H + h], 1.f), 1073741824.f)), logit[q][k]);
                    }
                }
            }
            // ---- per point: softmax, draw, log-densities
#pragma unroll
            for (int q = 0; q < PB; ++q) {
                const int n = n0 + q * LANES;
                if (n >= N) continue;
                float mx = logit[q][0];
#pragma unroll
                for (int k = 1; k < K; ++k) mx = fmaxf(mx, logit[q][k]);
                float pr[K], se = 0.f;
#pragma unroll
                for (int k = 0; k < K; ++k) { pr[k] = ex2_ftz((logit[q][k] - mx) * 1.4426950408889634f); se += pr[k]; }
                const float rse = rcp_ftz(se);
#pragma unroll
                for (int k = 0; k < K; ++k) pr[k] *= rse;                         // F.softmax(., -1)
                float sq = 0.f;
#pragma unroll
                for (int k = 0; k < K; ++k) sq = __fadd_rn(sq, pr[k]);            // Categorical renormalisation
                const float rsq = rcp_ftz(sq);
                if (a.q_probs != nullptr) {
#pragma unroll
                    for (int k = 0; k < K; ++k) a.q_probs[(r * N + n) * K + k] = pr[k];
                }
                int lf = 0;
                if (forced) {
                    float best = zn[n * K];
#pragma unroll
                    for (int k = 1; k < K; ++k) { const float v = zn[n * K + k]; if (v > best) { best = v; lf = k; } }
                } else {
                    // inverse CDF on the (unnormalised) probabilities: label = #{k : cdf_k <= u * sum}
                    const float t = __fmul_rn(__ldg(a.u + r * N + n), sq);
                    float c = 0.f;
#pragma unroll
                    for (int k = 0; k < K - 1; ++k) { c = __fadd_rn(c, pr[k]); lf += (t >= c) ? 1 : 0; }
#pragma unroll
                    for (int k = 0; k < K; ++k) zn[n * K + k] = (k == lf) ? 1.f : 0.f;
                }
                float plf = pr[0], lpf = logpi[0];
#pragma unroll
                for (int k = 1; k < K; ++k) if (lf == k) { plf = pr[k]; lpf = logpi[k]; }
                const float lqf = log_fast(fminf(fmaxf(plf * rsq, FLT_EPS_F), 1.f - FLT_EPS_F));
                float llf = 0.f;
#pragma unroll
                for (int d = 0; d < D; ++d) {
                    const float* tf = tab + (lf * D + d) * 3;
                    const float df = __fsub_rn(xv[q][d], tf[0]);
                    llf = __fadd_rn(llf, __fsub_rn(-__fmul_rn(__fmul_rn(df, df), tf[1]), tf[2]));
                }
                A_llf.add(llf); A_lqzf.add(lqf); A_lpzf.add(lpf);
                if (has_bwd) {
                    int lb = 0;
                    float best = zo[n * K];
#pragma unroll
                    for (int k = 1; k < K; ++k) { const float v = zo[n * K + k]; if (v > best) { best = v; lb = k; } }
                    float plb = pr[0], lpb = logpi[0];
#pragma unroll
                    for (int k = 1; k < K; ++k) if (lb == k) { plb = pr[k]; lpb = logpi[k]; }
                    const float lqb = log_fast(fminf(fmaxf(plb * rsq, FLT_EPS_F), 1.f - FLT_EPS_F));
                    float llb = 0.f;
#pragma unroll
                    for (int d = 0; d < D; ++d) {
                        const float* tbp = tab + (lb * D + d) * 3;
                        const float db = __fsub_rn(xv[q][d], tbp[0]);
                        llb = __fadd_rn(llb, __fsub_rn(-__fmul_rn(__fmul_rn(db, db), tbp[1]), tbp[2]));
                    }
                    A_llb.add(llb); A_lqzb.add(lqb); A_lpzb.add(lpb);
                    const float wf = __fsub_rn(__fadd_rn(llf, lpf), lqf);
                    const float wb = __fsub_rn(__fadd_rn(llb, lpb), lqb);
                    const float loc = __fsub_rn(wf, wb);
                    A_loc.add(loc);
                    if (a.lw_local != nullptr) a.lw_local[r * N + n] = loc;
                }
                if (a.lq_local != nullptr) a.lq_local[r * N + n] = lqf;
            }
        }
    }
    const double S_llf = group_sum<LANES>(A_llf.value()), S_llb = group_sum<LANES>(A_llb.value());
    const double S_lpzf = group_sum<LANES>(A_lpzf.value()), S_lpzb = group_sum<LANES>(A_lpzb.value());
    const double S_lqzf = group_sum<LANES>(A_lqzf.value()), S_lqzb = group_sum<LANES>(A_lqzb.value());
    const double S_loc = group_sum<LANES>(A_loc.value());
    if (valid && j == 0) {
        const float ll_f = (float)S_llf, ll_b = (float)S_llb;
        const float lpz_f = (float)S_lpzf, lpz_b = (float)S_lpzb, lqz_f = (float)S_lqzf, lqz_b = (float)S_lqzb;
        if (a.log_w != nullptr) a.log_w[r] = (float)S_loc;
        if (a.log_w_f != nullptr) a.log_w_f[r] = __fsub_rn(__fadd_rn(ll_f, lpz_f), lqz_f);
        if (a.log_w_b != nullptr) a.log_w_b[r] = __fsub_rn(__fadd_rn(ll_b, lpz_b), lqz_b);
        if (a.log_q_f != nullptr) a.log_q_f[r] = lqz_f;
        if (a.dens != nullptr) a.dens[r] = __fadd_rn(ll_f, lpz_f);
    }
    // ---- z_new tile out (one bulk store per CTA)
    tile_store(a.z_new + r0 * N * K, s_zn, tb);
}

constexpr int kFP = 32, kFLanes = 8, kFPB = 4;

template <int K, int D>
static int launch_zfused(const ZFusedArgs& a, cudaStream_t st) {
    const size_t smem = ((size_t)kFP * a.N * K * 2 + (size_t)a.H * (D + 1) + (size_t)kFP * (K * a.H + 1) +
                         (size_t)kFP * K * D * 3) * 4;
    if (smem > 200 * 1024) return fail(APGS_E_SHAPE, "gmm_z_fused: N=%d H=%d too large for one tile", a.N, a.H);
    auto kern = k_gmm_z_fused<K, D, kFP, kFLanes, kFPB>;
    APGS_CUDA_OK(ensure_dynamic_smem(reinterpret_cast<const void*>(kern), smem), "smem attr");
    const long long grid = (a.R + kFP - 1) / kFP;
    kern<<<(unsigned int)grid, kFP * kFLanes, smem, st>>>(a);
    APGS_LAUNCH_OK("k_gmm_z_fused launch");
    return 0;
}

}  // namespace apgs

using namespace apgs;

extern "C" int apgs_gmm_z_step_fused_f32(const float* x, const float* tau, const float* mu, const float* W1,
                                         const float* b1, const float* W2, const float* b2, int64_t H,
                                         const apgs_gmm_prior_t* prior, const float* z_old, const float* z_forced,
                                         const float* uniforms, int64_t S, int64_t B, int64_t N, int64_t K, int64_t D,
                                         float* z_new, float* q_probs, const apgs_gmm_out_t* out, void* stream) {
    if (S <= 0 || B <= 0 || N <= 0 || H <= 0 || H > 1024)
        return fail(APGS_E_SHAPE, "gmm_z_step_fused: bad sizes S=%lld B=%lld N=%lld H=%lld", (long long)S, (long long)B,
                    (long long)N, (long long)H);
    if (!x || !tau || !mu || !W1 || !b1 || !W2 || !b2 || !prior || !prior->pi || !z_new)
        return fail(APGS_E_BADARG, "gmm_z_step_fused: NULL input");
    if (!z_forced && !uniforms) return fail(APGS_E_BADARG, "gmm_z_step_fused: uniforms (or z_forced) required");
    ZFusedArgs a = {};
    a.x = x; a.tau = tau; a.mu = mu; a.W1 = W1; a.b1 = b1; a.W2 = W2; a.b2 = b2;
    a.z_old = z_old; a.z_forced = z_forced; a.u = uniforms; a.rpi = prior->pi;
    a.R = (long long)S * B; a.B = (int)B; a.N = (int)N; a.H = (int)H;
    a.z_new = z_new; a.q_probs = q_probs;
    if (out) {
        a.log_w = out->log_w; a.log_q_f = out->log_q_f; a.dens = out->density; a.log_w_f = out->log_w_f;
        a.log_w_b = out->log_w_b; a.lw_local = out->log_w_local; a.lq_local = out->log_q_local;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (K == 3 && D == 2) return launch_zfused<3, 2>(a, st);
    if (K == 2 && D == 2) return launch_zfused<2, 2>(a, st);
    if (K == 4 && D == 2) return launch_zfused<4, 2>(a, st);
    if (K == 5 && D == 2) return launch_zfused<5, 2>(a, st);
    if (K == 3 && D == 3) return launch_zfused<3, 3>(a, st);
    return fail(APGS_E_UNSUPPORTED, "gmm_z_step_fused: (K=%lld, D=%lld) not instantiated", (long long)K, (long long)D);
}
