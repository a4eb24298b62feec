// Fused DMM encoders / decoder for sm_100a (second half: Enc_apg_local; plus the
// Enc_apg_mu statistics and the decoder MLP, which the reference evaluates as ~25 PyTorch ops on (S,B,N,H) tensors).
//
//   k_dmm_pool_onehot      Enc_apg_mu (apgs/dmm/models.py:129-139): z is one-hot, so the two per-point MLPs on
//                          cat(x_n, z_n) take only K distinct values per data point.  The host tabulates them ONCE PER
//                          INSTANCE (tab[b][n][label] -- (B,N,K,.) rows through the same nn.Linear layers, S times less
//                          work) and this kernel does the pooling D1 as a label-indexed gather-accumulate over z.
//   k_dmm_local_fused      Enc_apg_local (apgs/dmm/models.py:76-101): per point the pi MLP (D -> H -> ReLU -> 1 on
//                          x_n - mu_k for each k), softmax, categorical draw by inverse CDF from a supplied uniform, and
//                          the two Beta-concentration MLPs (D -> H -> tanh -> 1, exp) on x_n - mu_{z_n} -- for the new
//                          assignment and, when z_old is given, for the old one too (the backward kernel).
//   k_dmm_decoder_hidden2  Decoder.recon_mu + unit normalisation (apgs/dmm/models.py:202-203).
//
// One warp owns one particle and its lanes stride over the N points; per-particle tables (W mu_k) sit in warp-private
// shared memory, weights in CTA-shared memory.  The first layers split as W (x - mu) + b = (W x + b) - W mu.
// tanh(u) = 1 - 2 / (2^(2 log2(e) u) + 1) with the scale folded into the first-layer weights (gmm_common.cuh).
#include "gmm_common.cuh"
#include "rng.cuh"

namespace apgs {

constexpr int kFW = 8;                    // warps (particles) per CTA

// ------------------------------------------------------------------------------------------ pooling
struct PoolTabArgs {
    const float *tab1, *tab2;       // (B, N, K, J), (B, N, K, K): nss1 / softmax(nss2) of point n GIVEN label
    const float* z;                 // (S, B, N, K) one-hot
    const float *mu_mu, *mu_sigma;  // (D)
    float* out;                     // (S, B, K, J + 2D)
    int S, B, N, K, J, D;
};

__global__ void __launch_bounds__(kFW * 32) k_dmm_pool_onehot(PoolTabArgs a) {
    extern __shared__ __align__(16) unsigned char s_lab_all[];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = (int)(blockIdx.x % (unsigned int)a.B);
    const int s = (int)(blockIdx.x / (unsigned int)a.B) * kFW + wid;
    if (s >= a.S) return;                                         // whole warp
    const int N = a.N, K = a.K, J = a.J, D = a.D, W = J + 2 * D;
    const long long r = (long long)s * a.B + b;
    unsigned char* lab = s_lab_all + (size_t)wid * ((N + 15) / 16 * 16);
    const float* zr = a.z + r * N * K;
    for (int n = lane; n < N; n += 32) {                          // first maximum, as argmax (models.py:86)
        int l = 0;
        float best = __ldg(zr + n * K);
        for (int k = 1; k < K; ++k) { const float v = __ldg(zr + n * K + k); if (v > best) { best = v; l = k; } }
        lab[n] = (unsigned char)l;
    }
    __syncwarp();
    const float* t1 = a.tab1 + (size_t)b * N * K * J;
    const float* t2 = a.tab2 + (size_t)b * N * K * K;
    float* o = a.out + r * K * W;
    for (int e = lane; e < K * J; e += 32) {
        const int k = e / J, j = e % J;
        double num = 0.0, den = 0.0;
#pragma unroll 4
        for (int n = 0; n < N; ++n) {
            const int l = lab[n];
            const float w = __ldg(t2 + ((size_t)n * K + l) * K + k);
            num += (double)__fmul_rn(__ldg(t1 + ((size_t)n * K + l) * J + j), w);
            den += (double)w;
        }
        o[k * W + j] = __fdiv_rn((float)num, __fadd_rn((float)den, 1e-8f));        // models.py:136
    }
    for (int e = lane; e < K * 2 * D; e += 32) {                                   // cat(prior_mu, prior_sigma)
        const int k = e / (2 * D), c = e % (2 * D);
        o[k * W + J + c] = c < D ? __ldg(a.mu_mu + c) : __ldg(a.mu_sigma + c - D);
    }
}

// Register-tiled variant for compile-time (K, J): lanes stride over the points (coalesced z reads, one vector load
// per table row) and keep all K*J numerators + K denominators of their points in registers; the 32 per-lane partials
// (<= ceil(N/32) fp32 terms each) meet in a butterfly.  ~4x fewer instructions than the element-per-lane loop above
// and no serial N-long load chain.
template <int K, int J>
__global__ void __launch_bounds__(kFW * 32) k_dmm_pool_onehot_t(PoolTabArgs a) {
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = (int)(blockIdx.x % (unsigned int)a.B);
    const int s = (int)(blockIdx.x / (unsigned int)a.B) * kFW + wid;
    if (s >= a.S) return;                                         // whole warp
    const int N = a.N, D = a.D, W = J + 2 * D;
    const long long r = (long long)s * a.B + b;
    const float* zr = a.z + r * N * K;
    const float* t1 = a.tab1 + (size_t)b * N * K * J;
    const float* t2 = a.tab2 + (size_t)b * N * K * K;
    float num[K][J], den[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        den[k] = 0.f;
#pragma unroll
        for (int j = 0; j < J; ++j) num[k][j] = 0.f;
    }
    for (int n = lane; n < N; n += 32) {
        int l = 0;
        float best = __ldg(zr + n * K);
#pragma unroll
        for (int k = 1; k < K; ++k) { const float v = __ldg(zr + n * K + k); if (v > best) { best = v; l = k; } }
        const float* r1 = t1 + ((size_t)n * K + l) * J;
        const float* r2 = t2 + ((size_t)n * K + l) * K;
        float v1[J], w[K];
#pragma unroll
        for (int j = 0; j < J; ++j) v1[j] = __ldg(r1 + j);
#pragma unroll
        for (int k = 0; k < K; ++k) w[k] = __ldg(r2 + k);
#pragma unroll
        for (int k = 0; k < K; ++k) {
            den[k] += w[k];
#pragma unroll
            for (int j = 0; j < J; ++j) num[k][j] = fmaf(v1[j], w[k], num[k][j]);
        }
    }
    float* o = a.out + r * K * W;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const float dk = __fadd_rn(warp_sum(den[k]), 1e-8f);                        // models.py:136
#pragma unroll
        for (int j = 0; j < J; ++j) {
            const float v = warp_sum(num[k][j]);
            if (lane == ((k * J + j) & 31)) o[k * W + j] = __fdiv_rn(v, dk);
        }
    }
    for (int e = lane; e < K * 2 * D; e += 32) {                                   // cat(prior_mu, prior_sigma)
        const int k = e / (2 * D), c = e % (2 * D);
        o[k * W + J + c] = c < D ? __ldg(a.mu_mu + c) : __ldg(a.mu_sigma + c - D);
    }
}

// ------------------------------------------------------------------------------------ local encoder
struct LocalEncArgs {
    const float* x;                          // (B, N, D)
    const float* mu;                         // (R, K, D)
    const float *pW0, *pb0, *pW2, *pb2;      // pi_log_prob:    (H, D) (H) (H) (1)
    const float *aW0, *ab0, *aW2, *ab2;      // angle_log_con1: (H, D) (H) (H) (1)
    const float *cW0, *cb0, *cW2, *cb2;      // angle_log_con0
    const float* u;                          // (R, N) uniforms, or null when z_forced
    const float* z_forced;                   // (R, N, K) or null
    const float* z_old;                      // (R, N, K) or null
    float *q_probs, *z;                      // (R, N, K) out
    float *con1, *con0;                      // (R, N) out: concentrations at z
    float *con1_old, *con0_old;              // (R, N) out: concentrations at z_old (when z_old)
    long long R;
    int B, N, H;
    // optional draws made in the kernel (Philox keyed by two host-drawn words, rng.cuh): the categorical draw's uniform when
    // `u` is null, and beta ~ Beta(con1, con0) as torch draws it (two standard Gammas, normalised, clamped) -> beta_out (R, N)
    const unsigned long long* seed;
    float* beta_out;
    // optional: the decoder's unit-normalised output of the drawn beta (Decoder.recon_mu, 1 -> HD -> tanh -> D) -> h2_out (R, N, D)
    const float *dW0, *db0, *dW2, *db2;
    float* h2_out;
    int HD;
};

// sum_h w2[h] tanh(pre[h]) over one point, pre[h] = W'[h] . d + b'[h] already scaled by 2 log2(e):
//   = sum_h w2[h]  -  sum_h 2 w2[h] / (2^pre[h] + 1);   wt holds {W'[h][0..D-1], b'[h], -2 w2[h]} per hidden unit
template <int D>
__device__ __forceinline__ float tanh_net(const float* __restrict__ wt, int H, const float (&d)[D], float w2sum) {
    float acc0 = w2sum, acc1 = 0.f;
    int h = 0;
#pragma unroll 2
    for (; h + 1 < H; h += 2) {
        const float* w0 = wt + h * (D + 2);
        const float* w1 = w0 + (D + 2);
        float p0 = w0[D], p1 = w1[D];
#pragma unroll
        for (int i = 0; i < D; ++i) { p0 = fmaf(w0[i], d[i], p0); p1 = fmaf(w1[i], d[i], p1); }
        // two reciprocals with one MUFU.RCP: 1/y0 = y1/(y0 y1); y clamped so the product stays finite
        const float y0 = fminf(ex2_ftz(p0) + 1.f, 9.223372e18f), y1 = fminf(ex2_ftz(p1) + 1.f, 9.223372e18f);
        const float rr = rcp_ftz(y0 * y1);
        acc0 = fmaf(w0[D + 1], rr * y1, acc0);
        acc1 = fmaf(w1[D + 1], rr * y0, acc1);
    }
    if (h < H) {
        const float* w0 = wt + h * (D + 2);
        float p0 = w0[D];
#pragma unroll
        for (int i = 0; i < D; ++i) p0 = fmaf(w0[i], d[i], p0);
        acc0 = fmaf(w0[D + 1], rcp_ftz(ex2_ftz(p0) + 1.f), acc0);
    }
    return acc0 + acc1;
}

template <int K, int D, int HC = 0>     // HC: hidden width when known at compile time (the reference's 32), else 0
__global__ void __launch_bounds__(kFW * 32) k_dmm_local_fused(LocalEncArgs a) {
    extern __shared__ __align__(16) float s_f[];
    const int H = HC ? HC : a.H, N = a.N;
    constexpr int WS = D + 2;
    float* s_pi = s_f;                          // [H][D+2]: W0[h][:], b0[h], W2[h]
    float* s_c1 = s_pi + H * WS;                // [H][D+2]: 2log2e W0[h][:], 2log2e b0[h], -2 W2[h]
    float* s_c0 = s_c1 + H * WS;
    float* s_mh = s_c0 + H * WS;                // [kFW][H][K]: W0_pi mu_k
    float* s_mu = s_mh + (size_t)kFW * H * K;   // [kFW][K][D]
    float* s_dw = s_mu + (size_t)kFW * K * D;   // decoder (optional): [HD][2 + D]: w0', b0', -2 W2[d][h]
    __shared__ float s_sum[2];
    __shared__ float s_db[D];
    const bool dec = a.h2_out != nullptr;
    if (dec) {
        constexpr float TWO_PI = 6.283185307179586f;
        for (int i = threadIdx.x; i < a.HD * (2 + D); i += blockDim.x) {
            const int h = i / (2 + D), c = i % (2 + D);
            s_dw[i] = c == 0 ? TANH_PRESCALE * TWO_PI * __ldg(a.dW0 + h)
                             : (c == 1 ? TANH_PRESCALE * __ldg(a.db0 + h) : -2.f * __ldg(a.dW2 + (c - 2) * a.HD + h));
        }
        if (threadIdx.x >= 64 && threadIdx.x < 64 + D) {
            const int d = threadIdx.x - 64;
            float sacc = __ldg(a.db2 + d);
            for (int h = 0; h < a.HD; ++h) sacc += __ldg(a.dW2 + d * a.HD + h);
            s_db[d] = sacc;
        }
    }
    for (int i = threadIdx.x; i < H * WS; i += blockDim.x) {
        const int h = i / WS, c = i % WS;
        s_pi[i] = c < D ? __ldg(a.pW0 + h * D + c) : (c == D ? __ldg(a.pb0 + h) : __ldg(a.pW2 + h));
        s_c1[i] = c < D ? TANH_PRESCALE * __ldg(a.aW0 + h * D + c)
                        : (c == D ? TANH_PRESCALE * __ldg(a.ab0 + h) : -2.f * __ldg(a.aW2 + h));
        s_c0[i] = c < D ? TANH_PRESCALE * __ldg(a.cW0 + h * D + c)
                        : (c == D ? TANH_PRESCALE * __ldg(a.cb0 + h) : -2.f * __ldg(a.cW2 + h));
    }
    if (threadIdx.x < 64) {                      // sum_h W2[h] + b2 of the two tanh nets (warp 0: con1, warp 1: con0)
        const int which = threadIdx.x >> 5, ln = threadIdx.x & 31;
        const float* w2 = which ? a.cW2 : a.aW2;
        float part = 0.f;
        for (int h = ln; h < H; h += 32) part += __ldg(w2 + h);
        part = warp_sum(part);
        if (ln == 0) s_sum[which] = part + __ldg(which ? a.cb2 : a.ab2);
    }
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long r = (long long)blockIdx.x * kFW + wid;
    const bool live = r < a.R;
    float* mh = s_mh + (size_t)wid * H * K;
    float* mu = s_mu + wid * K * D;
    if (live) {
        for (int e = lane; e < K * D; e += 32) mu[e] = __ldg(a.mu + r * K * D + e);
    }
    __syncthreads();
    if (!live) return;
    for (int e = lane; e < H * K; e += 32) {     // W0_pi mu_k, layout [h][k]
        const int h = e / K, k = e % K;
        float acc = 0.f;
#pragma unroll
        for (int d = 0; d < D; ++d) acc = fmaf(s_pi[h * WS + d], mu[k * D + d], acc);
        mh[e] = acc;
    }
    __syncwarp();
    const float pb2 = __ldg(a.pb2), sum1 = s_sum[0], sum0 = s_sum[1];
    const float* xb = a.x + (size_t)(r % a.B) * N * D;
    const bool forced = a.z_forced != nullptr, has_old = a.z_old != nullptr;
    const bool draw = a.seed != nullptr;
    uint2 key = make_uint2(0u, 0u);
    unsigned int st_lo = 0u, st_hi = 0u;
    if (draw) {
        const unsigned long long sd = __ldg(a.seed), st = __ldg(a.seed + 1);
        key = make_uint2((unsigned int)sd, (unsigned int)(sd >> 32));
        st_lo = (unsigned int)st; st_hi = (unsigned int)(st >> 32);
    }
    for (int n = lane; n < N; n += 32) {
        const long long pn = r * N + n;
        float xv[D];
#pragma unroll
        for (int d = 0; d < D; ++d) xv[d] = __ldg(xb + n * D + d);
        // Philox words of this point: block 0 = {uniform of the categorical draw, two words -> two normals, uniform of the
        // first Gamma}, block 1 = {uniform of the second Gamma, the two boost uniforms (alpha < 1)}
        uint4 w0 = make_uint4(0u, 0u, 0u, 0u), w1 = w0;
        const uint4 cbase = make_uint4((unsigned int)pn, (unsigned int)((unsigned long long)pn >> 32) ^ st_lo, st_hi, 0u);
        if (draw) {
            w0 = philox4x32_10(cbase, key);
            w1 = philox4x32_10(make_uint4(cbase.x, cbase.y, cbase.z ^ 0x10000000u, 0u), key);
        }
        const float un = forced ? 0.f : (a.u != nullptr ? __ldg(a.u + pn) : (float)(w0.x >> 8) * 5.9604644775390625e-08f);
        // ---- pi logits: W2 . relu((W0 x + b0) - W0 mu_k) + b2   (models.py:78-79)
        float lg[K];
#pragma unroll
        for (int k = 0; k < K; ++k) lg[k] = pb2;
#pragma unroll 4
        for (int h = 0; h < H; ++h) {
            const float* w = s_pi + h * WS;
            float xh = w[D];
#pragma unroll
            for (int d = 0; d < D; ++d) xh = fmaf(w[d], xv[d], xh);
            const float w2 = w[D + 1];
#pragma unroll
            for (int k = 0; k < K; ++k) lg[k] = fmaf(w2, fmaxf(xh - mh[h * K + k], 0.f), lg[k]);
        }
        float mx = lg[0];
#pragma unroll
        for (int k = 1; k < K; ++k) mx = fmaxf(mx, lg[k]);
        float pr[K], se = 0.f;
#pragma unroll
        for (int k = 0; k < K; ++k) { pr[k] = ex2_ftz((lg[k] - mx) * 1.4426950408889634f); se += pr[k]; }
        const float rse = rcp_ftz(se);
        float sq = 0.f;
#pragma unroll
        for (int k = 0; k < K; ++k) { pr[k] *= rse; sq = __fadd_rn(sq, pr[k]); a.q_probs[pn * K + k] = pr[k]; }
        // ---- assignment: the forced one, or the inverse CDF of the (unnormalised) probabilities at u * sum
        int lf = 0;
        if (forced) {
            float best = __ldg(a.z_forced + pn * K);
#pragma unroll
            for (int k = 1; k < K; ++k) { const float v = __ldg(a.z_forced + pn * K + k); if (v > best) { best = v; lf = k; } }
        } else {
            const float t = __fmul_rn(un, sq);
            float c = 0.f;
#pragma unroll
            for (int k = 0; k < K - 1; ++k) { c = __fadd_rn(c, pr[k]); lf += (t >= c) ? 1 : 0; }
        }
#pragma unroll
        for (int k = 0; k < K; ++k) a.z[pn * K + k] = (k == lf) ? 1.f : 0.f;
        // ---- Beta concentrations exp(MLP(x - mu_z))   (models.py:86-88)
        float dv[D];
#pragma unroll
        for (int d = 0; d < D; ++d) dv[d] = xv[d] - mu[lf * D + d];
        const float c1v = expf(tanh_net<D>(s_c1, H, dv, sum1)), c0v = expf(tanh_net<D>(s_c0, H, dv, sum0));
        a.con1[pn] = c1v;
        a.con0[pn] = c0v;
        if (draw && a.beta_out != nullptr) {
            // Beta(con1, con0).sample() as torch draws it (Dirichlet on [con1, con0], apgs/dmm/models.py:90): two standard
            // Gammas (Marsaglia-Tsang, clamped to the smallest normal like _standard_gamma), g1 / (g1 + g0), clamped to
            // [tiny, 1 - eps].  Distribution-equal, not stream-equal, to torch's draw.
            float n1, n0;
            box_muller(w0.y, w0.z, n1, n0);
            const float g1 = fmaxf(gamma_mt(c1v, n1, w0.w, w1.y, make_uint4(cbase.x, cbase.y, cbase.z ^ 0x20000000u, 0u), key), 1.17549435e-38f);
            const float g0 = fmaxf(gamma_mt(c0v, n0, w1.x, w1.z, make_uint4(cbase.x, cbase.y, cbase.z ^ 0x30000000u, 0u), key), 1.17549435e-38f);
            const float bt = fminf(fmaxf(__fdiv_rn(g1, __fadd_rn(g1, g0)), 1.17549435e-38f), 1.f - FLT_EPS_F);
            a.beta_out[pn] = bt;
            if (dec) {
                // Decoder.recon_mu(2 pi beta), unit-normalised (apgs/dmm/models.py:202-203): k_dmm_decoder_hidden2's arithmetic
                float o[D];
#pragma unroll
                for (int d = 0; d < D; ++d) o[d] = s_db[d];
                for (int h = 0; h < a.HD; ++h) {
                    const float* w = s_dw + h * (2 + D);
                    const float rcp = rcp_ftz(ex2_ftz(fmaf(w[0], bt, w[1])) + 1.f);
#pragma unroll
                    for (int d = 0; d < D; ++d) o[d] = fmaf(w[2 + d], rcp, o[d]);
                }
                float ss = 0.f;
#pragma unroll
                for (int d = 0; d < D; ++d) ss = __fadd_rn(ss, __fmul_rn(o[d], o[d]));
                const float nrm = __fsqrt_rn(ss);
#pragma unroll
                for (int d = 0; d < D; ++d) a.h2_out[pn * D + d] = __fdiv_rn(o[d], nrm);
            }
        }
        if (has_old) {                                                      // models.py:99-101 (backward kernel)
            int lb = 0;
            float best = __ldg(a.z_old + pn * K);
#pragma unroll
            for (int k = 1; k < K; ++k) { const float v = __ldg(a.z_old + pn * K + k); if (v > best) { best = v; lb = k; } }
            if (lb == lf) {
                a.con1_old[pn] = c1v;
                a.con0_old[pn] = c0v;
            } else {
#pragma unroll
                for (int d = 0; d < D; ++d) dv[d] = xv[d] - mu[lb * D + d];
                a.con1_old[pn] = expf(tanh_net<D>(s_c1, H, dv, sum1));
                a.con0_old[pn] = expf(tanh_net<D>(s_c0, H, dv, sum0));
            }
        }
    }
}

// ------------------------------------------------------------------------------------------ decoder
struct DecArgs {
    const float* beta;                  // (M)
    const float *W0, *b0, *W2, *b2;     // (H, 1) (H) (D, H) (D)
    float* out;                         // (M, D)
    long long M;
    int H;
};

template <int D>
__global__ void __launch_bounds__(256) k_dmm_decoder_hidden2(DecArgs a) {
    extern __shared__ __align__(16) float s_w[];           // [H][2 + D]: w0', b0', -2 W2[d][h]
    __shared__ float s_b[D];
    const int H = a.H;
    constexpr int WS = 2 + D;
    constexpr float TWO_PI = 6.283185307179586f;
    for (int i = threadIdx.x; i < H * WS; i += blockDim.x) {
        const int h = i / WS, c = i % WS;
        s_w[i] = c == 0 ? TANH_PRESCALE * TWO_PI * __ldg(a.W0 + h)
                        : (c == 1 ? TANH_PRESCALE * __ldg(a.b0 + h) : -2.f * __ldg(a.W2 + (c - 2) * H + h));
    }
    if (threadIdx.x < D) {
        float sacc = __ldg(a.b2 + threadIdx.x);
        for (int h = 0; h < H; ++h) sacc += __ldg(a.W2 + threadIdx.x * H + h);
        s_b[threadIdx.x] = sacc;
    }
    __syncthreads();
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.M; i += stride) {
        const float bt = __ldg(a.beta + i);
        float o[D];
#pragma unroll
        for (int d = 0; d < D; ++d) o[d] = s_b[d];
        for (int h = 0; h < H; ++h) {
            const float* w = s_w + h * WS;
            const float rcp = rcp_ftz(ex2_ftz(fmaf(w[0], bt, w[1])) + 1.f);
#pragma unroll
            for (int d = 0; d < D; ++d) o[d] = fmaf(w[2 + d], rcp, o[d]);
        }
        float ss = 0.f;
#pragma unroll
        for (int d = 0; d < D; ++d) ss = __fadd_rn(ss, __fmul_rn(o[d], o[d]));
        const float nrm = __fsqrt_rn(ss);                                   // models.py:203
#pragma unroll
        for (int d = 0; d < D; ++d) a.out[i * D + d] = __fdiv_rn(o[d], nrm);
    }
}

template <int K, int D>
static int launch_local(const LocalEncArgs& a, cudaStream_t st) {
    const size_t smem = ((size_t)3 * a.H * (D + 2) + (size_t)kFW * a.H * K + (size_t)kFW * K * D +
                         (a.h2_out != nullptr ? (size_t)a.HD * (2 + D) : 0)) * 4;
    if (smem > 200 * 1024) return fail(APGS_E_SHAPE, "dmm_local_encoder_fused: H=%d too large", a.H);
    auto kern = a.H == 32 ? k_dmm_local_fused<K, D, 32> : k_dmm_local_fused<K, D>;
    APGS_CUDA_OK(ensure_dynamic_smem(reinterpret_cast<const void*>(kern), smem), "smem attr");
    const long long grid = (a.R + kFW - 1) / kFW;
    kern<<<(unsigned int)grid, kFW * 32, smem, st>>>(a);
    APGS_LAUNCH_OK("k_dmm_local_fused launch");
    return 0;
}

}  // namespace apgs

using namespace apgs;

extern "C" int apgs_dmm_nss_pool_onehot_f32(const float* tab1, const float* tab2, const float* z,
                                            const float* prior_mu_mu, const float* prior_mu_sigma, int64_t S, int64_t B,
                                            int64_t N, int64_t K, int64_t J, int64_t D, float* out, void* stream) {
    if (S <= 0 || B <= 0 || N <= 0 || K <= 0 || J <= 0 || D <= 0 || K > 255)
        return fail(APGS_E_SHAPE, "dmm_nss_pool_onehot: bad sizes");
    if (S * B >= (1ll << 40) || N >= (1 << 20) || B * ((S + kFW - 1) / kFW) > 0x7fffffffLL)
        return fail(APGS_E_SHAPE, "dmm_nss_pool_onehot: problem too large");
    if (!tab1 || !tab2 || !z || !prior_mu_mu || !prior_mu_sigma || !out)
        return fail(APGS_E_BADARG, "dmm_nss_pool_onehot: NULL pointer");
    PoolTabArgs a = {tab1, tab2, z, prior_mu_mu, prior_mu_sigma, out, (int)S, (int)B, (int)N, (int)K, (int)J, (int)D};
    const size_t smem = (size_t)kFW * ((N + 15) / 16 * 16);
    if (smem > 200 * 1024) return fail(APGS_E_SHAPE, "dmm_nss_pool_onehot: N=%lld too large", (long long)N);
    APGS_CUDA_OK(ensure_dynamic_smem(reinterpret_cast<const void*>(k_dmm_pool_onehot), smem), "smem attr");
    const long long grid = B * ((S + kFW - 1) / kFW);
    if (K == 4 && J == 8) {
        k_dmm_pool_onehot_t<4, 8><<<(unsigned int)grid, kFW * 32, 0, (cudaStream_t)stream>>>(a);
    } else if (K == 3 && J == 8) {
        k_dmm_pool_onehot_t<3, 8><<<(unsigned int)grid, kFW * 32, 0, (cudaStream_t)stream>>>(a);
    } else
    k_dmm_pool_onehot<<<(unsigned int)grid, kFW * 32, smem, (cudaStream_t)stream>>>(a);
    APGS_LAUNCH_OK("k_dmm_pool_onehot launch");
    return 0;
}

static int local_encoder(const float* x, const float* mu, const apgs_dmm_mlp_t* pi,
                         const apgs_dmm_mlp_t* con1_net, const apgs_dmm_mlp_t* con0_net,
                         int64_t H, const float* uniforms, const float* z_forced,
                         const float* z_old, int64_t S, int64_t B, int64_t N, int64_t K,
                         int64_t D, float* q_probs, float* z, float* con1, float* con0,
                         float* con1_old, float* con0_old, void* stream, const uint64_t* seed, float* beta,
                         const apgs_dmm_mlp_t* decoder, int64_t HD, float* hidden2) {
    if (S <= 0 || B <= 0 || N <= 0 || H <= 0 || H > 1024)
        return fail(APGS_E_SHAPE, "dmm_local_encoder_fused: bad sizes S=%lld B=%lld N=%lld H=%lld", (long long)S,
                    (long long)B, (long long)N, (long long)H);
    if (S * B >= (1ll << 33) || N >= (1 << 24)) return fail(APGS_E_SHAPE, "dmm_local_encoder_fused: problem too large");
    if (!x || !mu || !pi || !con1_net || !con0_net || !q_probs || !z || !con1 || !con0)
        return fail(APGS_E_BADARG, "dmm_local_encoder_fused: NULL pointer");
    const apgs_dmm_mlp_t* nets[3] = {pi, con1_net, con0_net};
    for (int i = 0; i < 3; ++i)
        if (!nets[i]->W0 || !nets[i]->b0 || !nets[i]->W2 || !nets[i]->b2)
            return fail(APGS_E_BADARG, "dmm_local_encoder_fused: NULL weight");
    if (!uniforms && !z_forced && !seed) return fail(APGS_E_BADARG, "dmm_local_encoder_fused: uniforms (or z_forced, or a seed) required");
    if ((beta || hidden2) && !seed) return fail(APGS_E_BADARG, "dmm_local_encoder_draw: the Beta draw needs philox_seed");
    if (hidden2 && (!beta || !decoder || !decoder->W0 || !decoder->b0 || !decoder->W2 || !decoder->b2 || HD <= 0 || HD > 4096))
        return fail(APGS_E_BADARG, "dmm_local_encoder_draw: hidden2 needs beta, the decoder weights and its hidden width");
    if (z_old && (!con1_old || !con0_old))
        return fail(APGS_E_BADARG, "dmm_local_encoder_fused: con1_old/con0_old outputs required with z_old");
    LocalEncArgs a = {};
    a.x = x; a.mu = mu;
    a.pW0 = pi->W0; a.pb0 = pi->b0; a.pW2 = pi->W2; a.pb2 = pi->b2;
    a.aW0 = con1_net->W0; a.ab0 = con1_net->b0; a.aW2 = con1_net->W2; a.ab2 = con1_net->b2;
    a.cW0 = con0_net->W0; a.cb0 = con0_net->b0; a.cW2 = con0_net->W2; a.cb2 = con0_net->b2;
    a.u = uniforms; a.z_forced = z_forced; a.z_old = z_old;
    a.q_probs = q_probs; a.z = z; a.con1 = con1; a.con0 = con0; a.con1_old = con1_old; a.con0_old = con0_old;
    a.R = (long long)S * B; a.B = (int)B; a.N = (int)N; a.H = (int)H;
    a.seed = reinterpret_cast<const unsigned long long*>(seed); a.beta_out = beta; a.h2_out = hidden2; a.HD = (int)HD;
    if (hidden2) { a.dW0 = decoder->W0; a.db0 = decoder->b0; a.dW2 = decoder->W2; a.db2 = decoder->b2; }
    cudaStream_t st = (cudaStream_t)stream;
    if (K == 4 && D == 2) return launch_local<4, 2>(a, st);
    if (K == 3 && D == 2) return launch_local<3, 2>(a, st);
    if (K == 2 && D == 2) return launch_local<2, 2>(a, st);
    if (K == 5 && D == 2) return launch_local<5, 2>(a, st);
    if (K == 6 && D == 2) return launch_local<6, 2>(a, st);
    if (K == 4 && D == 3) return launch_local<4, 3>(a, st);
    if (K == 3 && D == 3) return launch_local<3, 3>(a, st);
    return fail(APGS_E_UNSUPPORTED, "dmm_local_encoder_fused: (K=%lld, D=%lld) not instantiated", (long long)K,
                (long long)D);
}

extern "C" int apgs_dmm_local_encoder_fused_f32(const float* x, const float* mu, const apgs_dmm_mlp_t* pi,
                                                const apgs_dmm_mlp_t* con1_net, const apgs_dmm_mlp_t* con0_net,
                                                int64_t H, const float* uniforms, const float* z_forced,
                                                const float* z_old, int64_t S, int64_t B, int64_t N, int64_t K,
                                                int64_t D, float* q_probs, float* z, float* con1, float* con0,
                                                float* con1_old, float* con0_old, void* stream) {
    return local_encoder(x, mu, pi, con1_net, con0_net, H, uniforms, z_forced, z_old, S, B, N, K, D, q_probs, z, con1, con0,
                         con1_old, con0_old, stream, nullptr, nullptr, nullptr, 0, nullptr);
}

extern "C" int apgs_dmm_local_encoder_draw_f32(const float* x, const float* mu, const apgs_dmm_mlp_t* pi,
                                               const apgs_dmm_mlp_t* con1_net, const apgs_dmm_mlp_t* con0_net,
                                               int64_t H, const float* uniforms, const float* z_forced,
                                               const float* z_old, const uint64_t* philox_seed, int64_t S, int64_t B,
                                               int64_t N, int64_t K, int64_t D, float* q_probs, float* z, float* con1,
                                               float* con0, float* con1_old, float* con0_old, float* beta,
                                               const apgs_dmm_mlp_t* decoder, int64_t decoder_hidden, float* hidden2,
                                               void* stream) {
    if (!philox_seed || !beta) return fail(APGS_E_BADARG, "dmm_local_encoder_draw: philox_seed and beta are required");
    return local_encoder(x, mu, pi, con1_net, con0_net, H, uniforms, z_forced, z_old, S, B, N, K, D, q_probs, z, con1, con0,
                         con1_old, con0_old, stream, philox_seed, beta, decoder, decoder_hidden, hidden2);
}

// bit 0: the fused local encoder is compiled for (K, D); bit 1: the decoder kernel for D.  The weight / pooling kernels
// of dmm.cu take K and D at run time.
extern "C" int apgs_dmm_supported(int64_t K, int64_t D) {
    const bool loc = (D == 2 && K >= 2 && K <= 6) || (D == 3 && (K == 3 || K == 4));
    return (loc ? 1 : 0) | ((D == 2 || D == 3) ? 2 : 0);
}

extern "C" int apgs_dmm_decoder_hidden2_f32(const float* beta, const apgs_dmm_mlp_t* net, int64_t H, int64_t D,
                                            int64_t M, float* hidden2, void* stream) {
    if (M <= 0 || H <= 0 || H > 4096 || D <= 0) return fail(APGS_E_SHAPE, "dmm_decoder_hidden2: bad sizes");
    if (!beta || !net || !net->W0 || !net->b0 || !net->W2 || !net->b2 || !hidden2)
        return fail(APGS_E_BADARG, "dmm_decoder_hidden2: NULL pointer");
    DecArgs a = {beta, net->W0, net->b0, net->W2, net->b2, hidden2, (long long)M, (int)H};
    const size_t smem = (size_t)H * (2 + D) * 4;
    const long long want = (M + 255) / 256;
    const unsigned int grid = (unsigned int)(want < (long long)sm_count() * 8 ? want : (long long)sm_count() * 8);
    cudaStream_t st = (cudaStream_t)stream;
    if (D == 2) {
        APGS_CUDA_OK(ensure_dynamic_smem(reinterpret_cast<const void*>(k_dmm_decoder_hidden2<2>), smem), "smem attr");
        k_dmm_decoder_hidden2<2><<<grid, 256, smem, st>>>(a);
    } else if (D == 3) {
        APGS_CUDA_OK(ensure_dynamic_smem(reinterpret_cast<const void*>(k_dmm_decoder_hidden2<3>), smem), "smem attr");
        k_dmm_decoder_hidden2<3><<<grid, 256, smem, st>>>(a);
    } else {
        return fail(APGS_E_UNSUPPORTED, "dmm_decoder_hidden2: D=%lld not instantiated", (long long)D);
    }
    APGS_LAUNCH_OK("k_dmm_decoder_hidden2 launch");
    return 0;
}
