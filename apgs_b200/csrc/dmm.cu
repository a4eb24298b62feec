// DMM (ring mixture) kernels for sm_100a.
//
//   k_dmm_nss_pool    D1  neural-sufficient-statistic pooling  (apgs/dmm/models.py:134-139, :38-43)
//   k_dmm_mu_step     D2  mu-step importance weight            (apgs/dmm/objectives.py:116-151 + Decoder tail
//                                                               apgs/dmm/models.py:190-209)
//   k_dmm_local_step  D3  {z, beta}-step weight per point, its sum, and the one-shot weight
//                                                              (objectives.py:153-188, :80-114; models.py:80-106)
//
// One WARP owns one particle (s,b); its lanes stride over the N points, so every per-point array
// ((S,B,N,*) rows contiguous per particle) is read with fully coalesced warp transactions straight from HBM --
// no (S,B,N,K,*) temporaries, no shared-memory staging needed.  The (K,D) means of the particle sit in a
// warp-private shared table for the label-indexed lookups.  fp32 arithmetic in the reference's operation order
// (torch.distributions formulas restated); per-lane partial sums are Kahan-compensated and widened to fp64 only
// for the cross-lane reduction.
#include "gmm_common.cuh"

namespace apgs {

constexpr int kDK = 16, kDD = 8;          // compile-time bounds on K and D (tables / local arrays)
constexpr int kWarps = 8;                 // particles per CTA

// torch/distributions/beta.py:87-91 -> dirichlet.py:90-97 on [v, 1-v] with concentrations [c1, c0]
__device__ __forceinline__ float beta_lp(float v, float c1, float c0) {
    const float a = xlogy_f(__fsub_rn(c1, 1.f), v);
    const float b = xlogy_f(__fsub_rn(c0, 1.f), __fsub_rn(1.f, v));
    const float s = __fadd_rn(a, b);
    return __fsub_rn(__fadd_rn(s, lgammaf(__fadd_rn(c1, c0))), __fadd_rn(lgammaf(c1), lgammaf(c0)));
}
// the same with the two logs supplied (they are shared by the proposal and the prior term of a point) and the
// normaliser through the Stirling lgamma of gmm_common.cuh (~1e-6 absolute; the sums these enter are O(1..100))
__device__ __forceinline__ float beta_norm(float c1, float c0) {
    return __fsub_rn(lgamma_pos(__fadd_rn(c1, c0)), __fadd_rn(lgamma_pos(c1), lgamma_pos(c0)));
}
__device__ __forceinline__ float beta_lp_logs(float c1, float c0, float lv, float l1v, float norm) {
    const float e1 = __fsub_rn(c1, 1.f), e0 = __fsub_rn(c0, 1.f);
    const float a = e1 == 0.f ? 0.f : __fmul_rn(e1, lv);
    const float b = e0 == 0.f ? 0.f : __fmul_rn(e0, l1v);
    return __fadd_rn(__fadd_rn(a, b), norm);
}
__device__ __forceinline__ float cat_lp(float p_l, float sum_p) {     // quirk Q4: eps-clamped categorical
    return logf(fminf(fmaxf(__fdiv_rn(p_l, sum_p), FLT_EPS_F), 1.f - FLT_EPS_F));
}
__device__ __forceinline__ int argmax_row(const float* __restrict__ row, int K) {
    int l = 0;
    float best = __ldg(row);
    for (int k = 1; k < K; ++k) { const float v = __ldg(row + k); if (v > best) { best = v; l = k; } }
    return l;
}

// ---------------------------------------------------------------------------------------------- D1
struct PoolArgs {
    const float *nss1, *nss2;       // (R, N, J), (R, N, K)
    const float *mu_mu, *mu_sigma;  // (D)
    float* out;                     // (R, K, J + 2D)
    long long R;
    int N, K, J, D;
};

__global__ void __launch_bounds__(kWarps * 32) k_dmm_nss_pool(PoolArgs a) {
    const long long r = (long long)blockIdx.x * kWarps + (threadIdx.x >> 5);
    if (r >= a.R) return;
    const int lane = threadIdx.x & 31;
    const int N = a.N, K = a.K, J = a.J, D = a.D, W = J + 2 * D;
    const float* n1 = a.nss1 + r * N * J;
    const float* n2 = a.nss2 + r * N * K;
    float* o = a.out + r * K * W;
    // lane -> output element (k, j): the warp re-reads the particle's rows from L1 (N*(J+K)*4 B per particle)
    for (int e = lane; e < K * J; e += 32) {
        const int k = e / J, j = e % J;
        double num = 0.0, den = 0.0;
        for (int n = 0; n < N; ++n) {
            const float w = __ldg(n2 + n * K + k);
            num += (double)__fmul_rn(__ldg(n1 + n * J + j), w);
            den += (double)w;
        }
        o[k * W + j] = __fdiv_rn((float)num, __fadd_rn((float)den, 1e-8f));        // models.py:136
    }
    for (int e = lane; e < K * 2 * D; e += 32) {                                   // cat(prior_mu, prior_sigma)
        const int k = e / (2 * D), c = e % (2 * D);
        o[k * W + J + c] = c < D ? __ldg(a.mu_mu + c) : __ldg(a.mu_sigma + c - D);
    }
}

// ------------------------------------------------------------------------------------------ D2 / D3
struct DmmArgs {
    const float* x;                          // (B, N, D)
    const float *mu, *mu_old;                // (R, K, D);  mu_old null unless mu step
    const float *q_mu, *q_sigma;             // (R, K, D) proposal of mu (mu step / one-shot), else null
    const float *z, *z_old;                  // (R, N, K) one-hot
    const float *h2, *h2_old;                // (R, N, D) unit-normalised decoder output for beta / beta_old
    const float* q_probs;                    // (R, N, K)       (local step / one-shot)
    const float *beta, *beta_old;            // (R, N)
    const float *con1, *con0, *con1_old, *con0_old;   // (R, N)
    const float *prior_mu_mu, *prior_mu_sigma;         // (D)
    const float* prior_pi;                   // (K)
    const float *prior_con1, *prior_con0;    // (1)
    const float *radi, *recon_sigma;         // (1)
    long long R;
    int B, N, K, D;
    int mode;                                // 0 mu step, 1 local step, 2 one-shot
    float *log_w, *log_q_f, *dens, *ll_f, *log_w_f, *log_w_b;      // (R), optional
    float *lw_local, *lq_local, *ll_local;                          // (R, N), optional
};

__global__ void __launch_bounds__(kWarps * 32, 4) k_dmm_step(DmmArgs a) {
    __shared__ float s_mu[kWarps][2][kDK * kDD];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long r = (long long)blockIdx.x * kWarps + wid;
    if (r >= a.R) return;                                       // whole warp exits together
    const int N = a.N, K = a.K, D = a.D, KD = K * D, mode = a.mode;
    const bool mu_terms = mode != 1, bwd = mode != 2, local = mode != 0;
    float* tf = s_mu[wid][0];
    float* tb = s_mu[wid][1];
    float lq_mu_f = 0.f, lp_mu_f = 0.f, lq_mu_b = 0.f, lp_mu_b = 0.f;
    for (int e = lane; e < KD; e += 32) {
        const float m = __ldg(a.mu + r * KD + e);
        tf[e] = m;
        const float mo = (mode == 0) ? __ldg(a.mu_old + r * KD + e) : m;
        tb[e] = mo;
        if (mu_terms) {
            const int d = e % D;
            const float qm = __ldg(a.q_mu + r * KD + e), qs = __ldg(a.q_sigma + r * KD + e);
            const float pm = __ldg(a.prior_mu_mu + d), ps = __ldg(a.prior_mu_sigma + d);
            lq_mu_f += normal_lp(m, qm, qs);
            lp_mu_f += normal_lp(m, pm, ps);
            if (mode == 0) { lq_mu_b += normal_lp(mo, qm, qs); lp_mu_b += normal_lp(mo, pm, ps); }
        }
    }
    __syncwarp();
    lq_mu_f = warp_sum(lq_mu_f); lp_mu_f = warp_sum(lp_mu_f); lq_mu_b = warp_sum(lq_mu_b); lp_mu_b = warp_sum(lp_mu_b);

    const float radi = __ldg(a.radi), sig = __ldg(a.recon_sigma);
    const float var2 = __fmul_rn(2.f, __fmul_rn(sig, sig)), lsig = logf(sig);
    float logpi[kDK], spi = 0.f;
    if (local) {
        for (int k = 0; k < K; ++k) spi = __fadd_rn(spi, __ldg(a.prior_pi + k));
        for (int k = 0; k < K; ++k) logpi[k] = cat_lp(__ldg(a.prior_pi + k), spi);
    }
    const float pc1 = local ? __ldg(a.prior_con1) : 1.f, pc0 = local ? __ldg(a.prior_con0) : 1.f;
    const float pnorm = local ? __fsub_rn(lgammaf(__fadd_rn(pc1, pc0)), __fadd_rn(lgammaf(pc1), lgammaf(pc0))) : 0.f;
    const float* xb = a.x + (size_t)(r % a.B) * N * D;
    KahanF A_llf, A_llb, A_lqf, A_lqb, A_lpf, A_lpb, A_loc, A_lqz, A_lqbeta, A_lpz, A_lpbeta;
    for (int n = lane; n < N; n += 32) {
        const long long pn = r * N + n;
        const int lf = argmax_row(a.z + pn * K, K);
        const int lb = (mode == 1) ? argmax_row(a.z_old + pn * K, K) : lf;
        // likelihood: Normal(h2 * radi + mu[z], sigma).log_prob(x), summed over D (models.py:204-209)
        float llf = 0.f, llb = 0.f;
        for (int d = 0; d < D; ++d) {
            const float xv = __ldg(xb + n * D + d);
            const float hf = __ldg(a.h2 + pn * D + d);
            const float df = __fsub_rn(xv, __fadd_rn(__fmul_rn(hf, radi), tf[lf * D + d]));
            llf = __fadd_rn(llf, __fsub_rn(__fsub_rn(__fdiv_rn(-__fmul_rn(df, df), var2), lsig), LOG_SQRT_2PI_F));
            if (bwd) {
                const float hb = (mode == 1) ? __ldg(a.h2_old + pn * D + d) : hf;
                const float db = __fsub_rn(xv, __fadd_rn(__fmul_rn(hb, radi), tb[lb * D + d]));
                llb = __fadd_rn(llb, __fsub_rn(__fsub_rn(__fdiv_rn(-__fmul_rn(db, db), var2), lsig), LOG_SQRT_2PI_F));
            }
        }
        A_llf.add(llf);
        A_llb.add(llb);
        if (local) {
            float sq = 0.f;
            for (int k = 0; k < K; ++k) sq = __fadd_rn(sq, __ldg(a.q_probs + pn * K + k));
            const float bf = __ldg(a.beta + pn);
            const float lqz = cat_lp(__ldg(a.q_probs + pn * K + lf), sq);
            const float lvf = logf(bf), l1vf = logf(__fsub_rn(1.f, bf));
            const float c1f = __ldg(a.con1 + pn), c0f = __ldg(a.con0 + pn);
            const float lqbeta = beta_lp_logs(c1f, c0f, lvf, l1vf, beta_norm(c1f, c0f));
            const float lpz = logpi[lf], lpbeta = beta_lp_logs(pc1, pc0, lvf, l1vf, pnorm);
            const float lqf = __fadd_rn(lqz, lqbeta);                                  // objectives.py:162
            const float lpf = __fadd_rn(__fadd_rn(llf, lpz), lpbeta);                  // :164
            A_lqf.add(lqf); A_lpf.add(lpf);
            A_lqz.add(lqz); A_lqbeta.add(lqbeta); A_lpz.add(lpz); A_lpbeta.add(lpbeta);
            if (a.lq_local) a.lq_local[pn] = lqf;
            if (a.ll_local) a.ll_local[pn] = llf;
            if (mode == 1) {
                const float bb = __ldg(a.beta_old + pn);
                const float lvb = logf(bb), l1vb = logf(__fsub_rn(1.f, bb));
                const float c1b = __ldg(a.con1_old + pn), c0b = __ldg(a.con0_old + pn);
                const float lqb = __fadd_rn(cat_lp(__ldg(a.q_probs + pn * K + lb), sq),
                                            beta_lp_logs(c1b, c0b, lvb, l1vb, beta_norm(c1b, c0b)));
                const float lpb = __fadd_rn(__fadd_rn(llb, logpi[lb]), beta_lp_logs(pc1, pc0, lvb, l1vb, pnorm));
                A_lqb.add(lqb); A_lpb.add(lpb);
                const float loc = __fsub_rn(__fsub_rn(lpf, lqf), __fsub_rn(lpb, lqb));  // :165,171-172
                A_loc.add(loc);
                if (a.lw_local) a.lw_local[pn] = loc;
            }
        }
    }
    const float ll_f = (float)warp_sum(A_llf.value()), ll_b = (float)warp_sum(A_llb.value());
    const float lqf = (float)warp_sum(A_lqf.value()), lqb = (float)warp_sum(A_lqb.value());
    const float lpf = (float)warp_sum(A_lpf.value()), lpb = (float)warp_sum(A_lpb.value());
    const float loc = (float)warp_sum(A_loc.value());
    const float lqz = (float)warp_sum(A_lqz.value()), lqbeta = (float)warp_sum(A_lqbeta.value());
    const float lpz = (float)warp_sum(A_lpz.value()), lpbeta = (float)warp_sum(A_lpbeta.value());
    if (lane != 0) return;
    float lw, lwf = 0.f, lwb = 0.f, lq, dens;
    if (mode == 0) {                    // objectives.py:123-135,150
        lwf = __fsub_rn(__fadd_rn(lp_mu_f, ll_f), lq_mu_f);
        lwb = __fsub_rn(__fadd_rn(lp_mu_b, ll_b), lq_mu_b);
        lw = __fsub_rn(lwf, lwb);
        lq = lq_mu_f;
        dens = lp_mu_f;
    } else if (mode == 1) {             // objectives.py:171-172,187
        lw = loc;
        lwf = __fsub_rn(lpf, lqf);
        lwb = __fsub_rn(lpb, lqb);
        lq = lqf;
        dens = lpf;
    } else {                            // one-shot: objectives.py:88-91
        lq = __fadd_rn(__fadd_rn(lq_mu_f, lqz), lqbeta);
        dens = __fadd_rn(__fadd_rn(__fadd_rn(ll_f, lp_mu_f), lpz), lpbeta);
        lw = __fsub_rn(dens, lq);
        lwf = lw;
    }
    if (a.log_w) a.log_w[r] = lw;
    if (a.log_q_f) a.log_q_f[r] = lq;
    if (a.dens) a.dens[r] = dens;
    if (a.ll_f) a.ll_f[r] = ll_f;
    if (a.log_w_f) a.log_w_f[r] = lwf;
    if (a.log_w_b) a.log_w_b[r] = lwb;
}

}  // namespace apgs

using namespace apgs;

static int check_dmm(int64_t S, int64_t B, int64_t N, int64_t K, int64_t D, const char* who) {
    if (S <= 0 || B <= 0 || N <= 0 || K <= 0 || D <= 0) return fail(APGS_E_SHAPE, "%s: sizes must be positive", who);
    if (K > kDK || D > kDD) return fail(APGS_E_UNSUPPORTED, "%s: K <= %d and D <= %d supported (K=%lld D=%lld)", who, kDK, kDD,
                                        (long long)K, (long long)D);
    if (S * B >= (1ll << 40) || N >= (1 << 24)) return fail(APGS_E_SHAPE, "%s: problem too large", who);
    return 0;
}

extern "C" int apgs_dmm_nss_pool_f32(const float* nss1, const float* nss2, const float* prior_mu_mu,
                                     const float* prior_mu_sigma, int64_t S, int64_t B, int64_t N, int64_t K,
                                     int64_t J, int64_t D, float* out, void* stream) {
    if (int e = check_dmm(S, B, N, K, D, "dmm_nss_pool")) return e;
    if (J <= 0) return fail(APGS_E_SHAPE, "dmm_nss_pool: num_nss must be positive");
    if (!nss1 || !nss2 || !prior_mu_mu || !prior_mu_sigma || !out) return fail(APGS_E_BADARG, "dmm_nss_pool: NULL pointer");
    PoolArgs a = {nss1, nss2, prior_mu_mu, prior_mu_sigma, out, (long long)S * B, (int)N, (int)K, (int)J, (int)D};
    const long long grid = (a.R + kWarps - 1) / kWarps;
    k_dmm_nss_pool<<<(unsigned int)grid, kWarps * 32, 0, (cudaStream_t)stream>>>(a);
    APGS_LAUNCH_OK("k_dmm_nss_pool launch");
    return 0;
}

static int launch_dmm(DmmArgs& a, const apgs_dmm_prior_t* pr, const apgs_dmm_out_t* out, int64_t S, int64_t B, int64_t N,
                      int64_t K, int64_t D, void* stream, const char* who) {
    if (int e = check_dmm(S, B, N, K, D, who)) return e;
    if (!pr || !out) return fail(APGS_E_BADARG, "%s: NULL prior/out struct", who);
    if (!pr->mu_mu || !pr->mu_sigma || !pr->pi || !pr->con1 || !pr->con0 || !pr->radi || !pr->recon_sigma)
        return fail(APGS_E_BADARG, "%s: NULL prior field", who);
    a.prior_mu_mu = pr->mu_mu; a.prior_mu_sigma = pr->mu_sigma; a.prior_pi = pr->pi; a.prior_con1 = pr->con1;
    a.prior_con0 = pr->con0; a.radi = pr->radi; a.recon_sigma = pr->recon_sigma;
    a.R = (long long)S * B; a.B = (int)B; a.N = (int)N; a.K = (int)K; a.D = (int)D;
    a.log_w = out->log_w; a.log_q_f = out->log_q_f; a.dens = out->density; a.ll_f = out->ll_f; a.log_w_f = out->log_w_f;
    a.log_w_b = out->log_w_b; a.lw_local = out->log_w_local; a.lq_local = out->log_q_local; a.ll_local = out->ll_local;
    const long long grid = (a.R + kWarps - 1) / kWarps;
    k_dmm_step<<<(unsigned int)grid, kWarps * 32, 0, (cudaStream_t)stream>>>(a);
    APGS_LAUNCH_OK("k_dmm_step launch");
    return 0;
}

extern "C" int apgs_dmm_logw_mu_f32(const float* x, const float* z, const float* hidden2, const float* q_mu,
                                    const float* q_sigma, const float* mu, const float* mu_old,
                                    const apgs_dmm_prior_t* prior, int64_t S, int64_t B, int64_t N, int64_t K, int64_t D,
                                    const apgs_dmm_out_t* out, void* stream) {
    if (!x || !z || !hidden2 || !q_mu || !q_sigma || !mu || !mu_old) return fail(APGS_E_BADARG, "dmm_logw_mu: NULL pointer");
    DmmArgs a = {};
    a.x = x; a.z = z; a.h2 = hidden2; a.q_mu = q_mu; a.q_sigma = q_sigma; a.mu = mu; a.mu_old = mu_old; a.mode = 0;
    return launch_dmm(a, prior, out, S, B, N, K, D, stream, "dmm_logw_mu");
}

extern "C" int apgs_dmm_logw_local_f32(const float* x, const float* mu, const float* q_probs, const float* z,
                                       const float* beta, const float* con1, const float* con0, const float* hidden2,
                                       const float* z_old, const float* beta_old, const float* con1_old,
                                       const float* con0_old, const float* hidden2_old, const apgs_dmm_prior_t* prior,
                                       int64_t S, int64_t B, int64_t N, int64_t K, int64_t D, const apgs_dmm_out_t* out,
                                       void* stream) {
    if (!x || !mu || !q_probs || !z || !beta || !con1 || !con0 || !hidden2 || !z_old || !beta_old || !con1_old ||
        !con0_old || !hidden2_old)
        return fail(APGS_E_BADARG, "dmm_logw_local: NULL pointer");
    DmmArgs a = {};
    a.x = x; a.mu = mu; a.q_probs = q_probs; a.z = z; a.beta = beta; a.con1 = con1; a.con0 = con0; a.h2 = hidden2;
    a.z_old = z_old; a.beta_old = beta_old; a.con1_old = con1_old; a.con0_old = con0_old; a.h2_old = hidden2_old;
    a.mode = 1;
    return launch_dmm(a, prior, out, S, B, N, K, D, stream, "dmm_logw_local");
}

extern "C" int apgs_dmm_logw_oneshot_f32(const float* x, const float* q_mu, const float* q_sigma, const float* mu,
                                         const float* q_probs, const float* z, const float* beta, const float* con1,
                                         const float* con0, const float* hidden2, const apgs_dmm_prior_t* prior,
                                         int64_t S, int64_t B, int64_t N, int64_t K, int64_t D,
                                         const apgs_dmm_out_t* out, void* stream) {
    if (!x || !q_mu || !q_sigma || !mu || !q_probs || !z || !beta || !con1 || !con0 || !hidden2)
        return fail(APGS_E_BADARG, "dmm_logw_oneshot: NULL pointer");
    DmmArgs a = {};
    a.x = x; a.q_mu = q_mu; a.q_sigma = q_sigma; a.mu = mu; a.q_probs = q_probs; a.z = z; a.beta = beta; a.con1 = con1;
    a.con0 = con0; a.h2 = hidden2; a.mode = 2;
    return launch_dmm(a, prior, out, S, B, N, K, D, stream, "dmm_logw_oneshot");
}
