// Training mode of the GMM sweep on sm_100a: backward kernels for the two proposal log-densities
// whose gradients the APG loss needs,  loss = sum_s w_s.detach() * (-log q_s)  (apgs/gmm/objectives.py:183,220).
//
//   k_gmm_z_logq_bwd     d/d{W1,b1,W2,b2} of  sum_{r,n} G[r,n] * log q(z_n | x_n, tau, mu)  for Enc_apg_z
//                        (apgs/gmm/models.py:102-115): per (point n, cluster k) MLP 3D -> H -> 1, softmax over k,
//                        OneHotCategorical log-prob (torch's eps clamp: zero gradient outside [eps, 1 - eps]).
//                        G is the upstream gradient (the detached per-point weights of the loss, or any other).
//                        The forward kernel's q_probs are re-used; the hidden layer is recomputed from the factorised
//                        exponent tables of the forward kernel (gmm_fused.cu): 2^(xh'_n + e'_k) = Ex_n * Ee_k.
//   k_label_table_sum    out[r,c] = sum_n T[b,n,label(r,n),c]       and its adjoint
//   k_label_table_sum_bwd   gT[b,n,l,c] = sum_s [label(s,b,n) == l] * gout[s,b,c]
//                        -- with one-hot z the two Linear heads of Enc_apg_eta (models.py:56-66) take only K values per
//                        data point, so the host tabulates their outputs once per INSTANCE with ordinary (differentiable)
//                        PyTorch ops on a (B,N,K,.) tensor and the per-particle neural statistics are label-indexed sums
//                        of that table: the only S-sized work of the eta step's graph is this pair of kernels.
//
// Deterministic: partial sums travel through shared / global arrays and are combined in a fixed order (no atomics).
#include <stdlib.h>
#include "gmm_common.cuh"

namespace apgs {

struct ZBwdArgs {
    const float* x;                     // (B, N, D)
    const float *tau, *mu;              // (R, K, D)
    const float *W1, *b1, *W2;          // (H, 3D), (H), (H)
    const float* q;                     // (R, N, K) q_probs of the forward kernel
    const float* z;                     // (R, N, K) one-hot assignment, or null
    const unsigned char* zl;            // (R, LN) labels, or null
    long long LN;
    const float* g;                     // upstream gradient d loss / d log q[r, n]
    long long g_sr, g_sn;               // element (r, n) at g[r * g_sr + n * g_sn]  (g_sn = 0: one value per particle)
    int S, B, N, H, chunks;             // chunks of kZP particles handled by one CTA
    int force_point_major;              // tests: run the generic point-major kernel even when H <= 32
    float* partial;                     // [gridDim.y * gridDim.x][GS]
};

constexpr int kZP = 32, kZL = 8, kZPT = 8;       // particles per chunk, lanes per particle, points per lane and block

// values accumulated per hidden unit: [dW2, db1 / (4 W2), dW1x (D), dW1m (D), dW1t (D)] (the factor 4 W2_h is applied last)
template <int D> struct ZBwdV { static constexpr int V = 2 + 3 * D; };

template <int K, int D>
__global__ void __launch_bounds__(kZP* kZL) k_gmm_z_logq_bwd(ZBwdArgs a) {
    extern __shared__ __align__(16) float smf[];
    constexpr int KD = K * D, V = ZBwdV<D>::V, NW = kZP * kZL / 32;
    const int N = a.N, H = a.H;
    const int NP = (N + kZL * kZPT - 1) / (kZL * kZPT) * (kZL * kZPT);       // points padded to blocks of 64
    const int EPS = H * K + 1;                                               // per-particle stride of the Ee table
    float* s_ex = smf;                                   // [H][NP]   Ex (or raw xh' when the CTA is not `xok`)
    float* s_e = s_ex + (size_t)H * NP;                  // [kZP][EPS] Ee (or raw e')
    float* s_x = s_e + (size_t)kZP * EPS;                // [NP][D]
    float* s_acc = s_x + (size_t)NP * D;                 // [NW][H][V] (+1: db2)
    float* s_w2 = s_acc + (size_t)NW * (H * V + 1);      // [H]
    __shared__ int s_xok;
    const int b = blockIdx.x;
    const int p = threadIdx.x / kZL, j = threadIdx.x % kZL, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float* xb = a.x + (size_t)b * N * D;
    if (threadIdx.x == 0) s_xok = 1;
    for (int i = threadIdx.x; i < NW * (H * V + 1); i += blockDim.x) s_acc[i] = 0.f;
    for (int i = threadIdx.x; i < NP * D; i += blockDim.x) s_x[i] = (i < N * D) ? __ldg(xb + i) : 0.f;
    for (int i = threadIdx.x; i < H; i += blockDim.x) s_w2[i] = __ldg(a.W2 + i);
    __syncthreads();
    // ---- per-CTA table of the data part of the pre-activations
    bool okx = true;
    for (int idx = threadIdx.x; idx < H * NP; idx += blockDim.x) {
        const int h = idx / NP, n = idx - h * NP;
        float acc = 0.f;
        if (n < N) {
#pragma unroll
            for (int d = 0; d < D; ++d) acc = fmaf(TANH_PRESCALE * __ldg(a.W1 + h * 3 * D + d), s_x[n * D + d], acc);
            if (!(fabsf(acc) <= 126.f)) okx = false;
        }
        s_ex[idx] = acc;
    }
    if (!okx) s_xok = 0;                                  // benign race: every writer stores 0
    __syncthreads();
    const bool xok = s_xok != 0;
    if (xok)
        for (int idx = threadIdx.x; idx < H * NP; idx += blockDim.x) {
            const int n = idx % NP;
            s_ex[idx] = (n < N) ? ex2_ftz(s_ex[idx]) : 0.f;          // padding points: y = 1, and their g is 0
        }
    __syncthreads();
    float gb2 = 0.f;
    float* my_acc = s_acc + (size_t)warp * (H * V + 1);
    for (int c = 0; c < a.chunks; ++c) {
        const long long s = ((long long)blockIdx.y * a.chunks + c) * kZP + p;
        const bool valid = s < a.S;
        const long long r = s * a.B + b;
        float tk[KD], mk[KD];
#pragma unroll
        for (int e = 0; e < KD; ++e) { tk[e] = valid ? __ldg(a.tau + r * KD + e) : 1.f; mk[e] = valid ? __ldg(a.mu + r * KD + e) : 0.f; }
        // ---- particle part of the pre-activations: e'_k[h]; the factorised form needs every |e'| <= 126
        float* ep = s_e + (size_t)p * EPS;
        bool oke = true;
        __syncwarp();                                    // the previous chunk's readers of this block are done
        for (int h = j; h < H; h += kZL) {
            const float bh = __ldg(a.b1 + h);
#pragma unroll
            for (int k = 0; k < K; ++k) {
                float acc = bh;
#pragma unroll
                for (int d = 0; d < D; ++d) {
                    acc = fmaf(__ldg(a.W1 + h * 3 * D + D + d), mk[k * D + d], acc);
                    acc = fmaf(__ldg(a.W1 + h * 3 * D + 2 * D + d), tk[k * D + d], acc);
                }
                const float ee = TANH_PRESCALE * acc;
                if (!(fabsf(ee) <= 126.f)) oke = false;
                ep[h * K + k] = ee;
            }
        }
#pragma unroll
        for (int o = kZL / 2; o > 0; o >>= 1) oke = (__shfl_xor_sync(0xffffffffu, (int)oke, o) != 0) && oke;
        const bool fast = xok && oke;
        __syncwarp();
        if (fast)
            for (int h = j; h < H; h += kZL)
#pragma unroll
                for (int k = 0; k < K; ++k) ep[h * K + k] = ex2_ftz(ep[h * K + k]);
        __syncwarp();
        for (int n0 = 0; n0 < N; n0 += kZL * kZPT) {
            // ---- this lane's points of the block: upstream gradient times d log q / d logit_k = (k == z) - q_k
            float gk[kZPT][K];
#pragma unroll
            for (int i = 0; i < kZPT; ++i) {
                const int n = n0 + j + kZL * i;
#pragma unroll
                for (int k = 0; k < K; ++k) gk[i][k] = 0.f;
                if (valid && n < N) {
                    float qv[K], sq = 0.f;
#pragma unroll
                    for (int k = 0; k < K; ++k) { qv[k] = __ldg(a.q + (r * N + n) * K + k); sq = __fadd_rn(sq, qv[k]); }
                    int lf = 0;
                    if (a.zl != nullptr) {
                        lf = min((int)__ldg(a.zl + r * a.LN + n), K - 1);
                    } else {
                        float best = __ldg(a.z + (r * N + n) * K);
#pragma unroll
                        for (int k = 1; k < K; ++k) { const float v = __ldg(a.z + (r * N + n) * K + k); if (v > best) { best = v; lf = k; } }
                    }
                    float ql = qv[0];
#pragma unroll
                    for (int k = 1; k < K; ++k) if (lf == k) ql = qv[k];
                    const float pn = __fdiv_rn(ql, sq);                          // Categorical's renormalisation
                    const bool inside = pn >= FLT_EPS_F && pn <= 1.f - FLT_EPS_F;   // clamp_probs: zero gradient outside
                    const float up = inside ? __ldg(a.g + r * a.g_sr + n * a.g_sn) : 0.f;
#pragma unroll
                    for (int k = 0; k < K; ++k) { gk[i][k] = up * (((k == lf) ? 1.f : 0.f) - qv[k]); gb2 += gk[i][k]; }
                }
            }
            // ---- hidden units: t = tanh(a) = 1 - 2 r, 1 - t^2 = 4 (r - r^2), r = 1 / (2^(xh' + e') + 1)
            for (int h = 0; h < H; ++h) {
                float G2 = 0.f, A[K], X[D];
#pragma unroll
                for (int k = 0; k < K; ++k) A[k] = 0.f;
#pragma unroll
                for (int d = 0; d < D; ++d) X[d] = 0.f;
                float ee[K];
#pragma unroll
                for (int k = 0; k < K; ++k) ee[k] = ep[h * K + k];
                const float* exr = s_ex + (size_t)h * NP + n0 + j;
#pragma unroll
                for (int i = 0; i < kZPT; ++i) {
                    const float ex = exr[kZL * i];
                    float sn = 0.f;
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        const float y = fast ? fmaf(ex, ee[k], 1.f) : (ex2_ftz((xok ? 0.f : ex) + ee[k]) + 1.f);
                        const float rr = rcp_ftz(y);
                        const float t = fmaf(-2.f, rr, 1.f);
                        const float v = fmaf(-rr, rr, rr);
                        const float cc = gk[i][k] * v;
                        G2 = fmaf(gk[i][k], t, G2);
                        A[k] += cc;
                        sn += cc;
                    }
                    const int n = n0 + j + kZL * i;
#pragma unroll
                    for (int d = 0; d < D; ++d) X[d] = fmaf(sn, s_x[n * D + d], X[d]);
                }
                // the particle's lanes, then the warp's particles
#pragma unroll
                for (int o = 1; o < kZL; o <<= 1) {
                    G2 += __shfl_xor_sync(0xffffffffu, G2, o);
#pragma unroll
                    for (int k = 0; k < K; ++k) A[k] += __shfl_xor_sync(0xffffffffu, A[k], o);
#pragma unroll
                    for (int d = 0; d < D; ++d) X[d] += __shfl_xor_sync(0xffffffffu, X[d], o);
                }
                float vals[V];
                vals[0] = G2;
                vals[1] = 0.f;
#pragma unroll
                for (int k = 0; k < K; ++k) vals[1] += A[k];
#pragma unroll
                for (int d = 0; d < D; ++d) {
                    float m = 0.f, t = 0.f;
#pragma unroll
                    for (int k = 0; k < K; ++k) { m = fmaf(A[k], mk[k * D + d], m); t = fmaf(A[k], tk[k * D + d], t); }
                    vals[2 + d] = X[d]; vals[2 + D + d] = m; vals[2 + 2 * D + d] = t;
                }
#pragma unroll
                for (int v = 0; v < V; ++v) {
#pragma unroll
                    for (int o = kZL; o < 32; o <<= 1) vals[v] += __shfl_xor_sync(0xffffffffu, vals[v], o);
                }
                if (lane == 0) {
#pragma unroll
                    for (int v = 0; v < V; ++v) my_acc[h * V + v] += vals[v];
                }
            }
        }
    }
    // NOTE on the slow path: when the CTA is not `xok`, s_ex holds the raw xh' (see above) and `fast` is false for every
    // particle, so y = 2^(xh' + e') + 1 exactly; when it is `xok` but a particle is not, s_ex holds Ex = 2^xh' and the
    // exponent cannot be re-assembled -- such particles (|e'| > 126: tanh saturated for every point unless xh' cancels
    // it to within 126) take y = 2^e' + 1, i.e. the saturated value, whose gradient is zero to fp32 precision.
    gb2 = warp_sum(gb2);
    if (lane == 0) my_acc[H * V] = gb2;
    __syncthreads();
    const int GS = H * V + 1;
    float* out = a.partial + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * GS;
    for (int i = threadIdx.x; i < GS; i += blockDim.x) {
        float sum = 0.f;
#pragma unroll
        for (int w = 0; w < NW; ++w) sum += s_acc[(size_t)w * GS + i];
        out[i] = sum;
    }
}

// all points of one particle x the HQ hidden units of one lane.  FAST: y = Ex * Ee + 1 from the factorised tables; else the
// exact form y = 2^(xh' + e') + 1 (s_ex holds the raw xh' when the CTA is not `xok`; see the note in the point-major kernel)
#ifndef APGS_ZBWD_UNROLL
#define APGS_ZBWD_UNROLL 2
#endif
#ifndef APGS_ZBWD_MINB
#define APGS_ZBWD_MINB 2
#endif
template <int K, int D, int HQ, bool FAST>
__device__ __forceinline__ void zbwd_points(int N, const float* __restrict__ gp, const float* __restrict__ s_x,
                                            const float* __restrict__ exl, int H8, const float (&ee)[HQ][K], bool xok,
                                            float (&G2)[HQ], float (&A)[HQ][K], float (&X)[HQ][D]) {
    constexpr int kUnroll = APGS_ZBWD_UNROLL;
#pragma unroll kUnroll
    for (int n = 0; n < N; ++n) {
        float g[K], xn[D], ex[HQ];
#pragma unroll
        for (int k = 0; k < K; ++k) g[k] = gp[n * K + k];
#pragma unroll
        for (int d = 0; d < D; ++d) xn[d] = s_x[n * D + d];
        if (HQ == 4) {
            const float4 v = *reinterpret_cast<const float4*>(exl + (size_t)n * H8);
            ex[0] = v.x; ex[1 % HQ] = v.y; ex[2 % HQ] = v.z; ex[3 % HQ] = v.w;
        } else {
#pragma unroll
            for (int i = 0; i < HQ; ++i) ex[i] = exl[(size_t)n * H8 + i];
        }
#pragma unroll
        for (int i = 0; i < HQ; ++i) {
            float sn = 0.f;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                // t = 1 - 2 r and 1 - t^2 = 4 r (1 - r):  g t = g - 2 (g r)  (the caller adds sum g once per particle) and
                // g r (1 - r) = gr - gr r: six FMA-pipe operations and one MUFU per term
                const float y = FAST ? fmaf(ex[i], ee[i][k], 1.f) : (ex2_ftz((xok ? 0.f : ex[i]) + ee[i][k]) + 1.f);
                const float rr = rcp_ftz(y);
                const float gr = g[k] * rr;
                const float cc = fmaf(-gr, rr, gr);
                G2[i] += gr;                                 // sum g r
                A[i][k] += cc;
                sn += cc;
            }
#pragma unroll
            for (int d = 0; d < D; ++d) X[i][d] = fmaf(sn, xn[d], X[i][d]);
        }
    }
}

// ---- hidden-unit-major variant (H <= 8 HQ): lane j of a particle OWNS the hidden units h = j HQ .. j HQ + HQ - 1 and walks
// all N points of the particle, so every accumulator is private to its thread for the whole chunk: no shuffles in the
// main loop (the point-major kernel above spends ~30 % of its instructions reducing 6 sums per hidden unit across the
// warp).  The upstream-gradient factors g[n][k] of a particle are staged once in shared memory by its 8 lanes; Ex is laid
// out [n][h] so that a lane reads its HQ units with one vector load.  Per-thread totals live in shared memory slots owned
// by the thread (no atomics); the CTA's particles are combined in a fixed order at the end.
template <int K, int D, int HQ>
__global__ void __launch_bounds__(kZP* kZL, APGS_ZBWD_MINB) k_gmm_z_logq_bwd_t(ZBwdArgs a) {
    extern __shared__ __align__(16) float smf[];
    constexpr int KD = K * D, V = ZBwdV<D>::V, H8 = HQ * kZL, NT = kZP * kZL;
    const int N = a.N, H = a.H;
    float* s_ex = smf;                                   // [N][H8]   Ex (or raw xh' when the CTA is not `xok`)
    float* s_x = s_ex + (size_t)N * H8;                  // [N][D] (+pad)
    float* s_g = s_x + (size_t)((N * D + 3) / 4 * 4);    // [kZP][N*K]
    float* s_tot = s_g + (size_t)kZP * ((N * K + 3) / 4 * 4);     // [NT][HQ*V + 1] (odd stride: a warp's slots on 32 banks)
    constexpr int TS = HQ * V + 1;
    __shared__ int s_xok;
    __shared__ float s_gb2[NT / 32];
    const int GP = (N * K + 3) / 4 * 4;
    const int b = blockIdx.x;
    const int p = threadIdx.x / kZL, j = threadIdx.x % kZL, lane = threadIdx.x & 31;
    const float* xb = a.x + (size_t)b * N * D;
    if (threadIdx.x == 0) s_xok = 1;
    for (int i = threadIdx.x; i < N * D; i += NT) s_x[i] = __ldg(xb + i);
    float* tot = s_tot + (size_t)threadIdx.x * TS;
#pragma unroll
    for (int i = 0; i < HQ * V; ++i) tot[i] = 0.f;
    __syncthreads();
    bool okx = true;
    for (int idx = threadIdx.x; idx < N * H8; idx += NT) {
        const int n = idx / H8, h = idx - n * H8;
        float acc = 0.f;
        if (h < H) {
#pragma unroll
            for (int d = 0; d < D; ++d) acc = fmaf(TANH_PRESCALE * __ldg(a.W1 + h * 3 * D + d), s_x[n * D + d], acc);
            if (!(fabsf(acc) <= 126.f)) okx = false;
        }
        s_ex[idx] = acc;
    }
    if (!okx) s_xok = 0;
    __syncthreads();
    const bool xok = s_xok != 0;
    if (xok)
        for (int idx = threadIdx.x; idx < N * H8; idx += NT) s_ex[idx] = ex2_ftz(s_ex[idx]);
    __syncthreads();
    // this lane's hidden units: first-layer weights of the particle part, output weight is applied by the finishing kernel
    float wm[HQ][D], wt[HQ][D], bh[HQ];
#pragma unroll
    for (int i = 0; i < HQ; ++i) {
        const int h = j * HQ + i;
        bh[i] = (h < H) ? __ldg(a.b1 + h) : 0.f;
#pragma unroll
        for (int d = 0; d < D; ++d) {
            wm[i][d] = (h < H) ? __ldg(a.W1 + h * 3 * D + D + d) : 0.f;
            wt[i][d] = (h < H) ? __ldg(a.W1 + h * 3 * D + 2 * D + d) : 0.f;
        }
    }
    float gb2 = 0.f;
    float* gp = s_g + (size_t)p * GP;
    for (int c = 0; c < a.chunks; ++c) {
        const long long s = ((long long)blockIdx.y * a.chunks + c) * kZP + p;
        const bool valid = s < a.S;
        const long long r = s * a.B + b;
        float tk[KD], mk[KD];
#pragma unroll
        for (int e = 0; e < KD; ++e) { tk[e] = valid ? __ldg(a.tau + r * KD + e) : 1.f; mk[e] = valid ? __ldg(a.mu + r * KD + e) : 0.f; }
        float ee[HQ][K];
        bool oke = true;
#pragma unroll
        for (int i = 0; i < HQ; ++i)
#pragma unroll
            for (int k = 0; k < K; ++k) {
                float acc = bh[i];
#pragma unroll
                for (int d = 0; d < D; ++d) { acc = fmaf(wm[i][d], mk[k * D + d], acc); acc = fmaf(wt[i][d], tk[k * D + d], acc); }
                ee[i][k] = TANH_PRESCALE * acc;
                if (!(fabsf(ee[i][k]) <= 126.f)) oke = false;
            }
#pragma unroll
        for (int o = kZL / 2; o > 0; o >>= 1) oke = (__shfl_xor_sync(0xffffffffu, (int)oke, o) != 0) && oke;
        const bool fast = xok && oke;
        if (fast) {
#pragma unroll
            for (int i = 0; i < HQ; ++i)
#pragma unroll
                for (int k = 0; k < K; ++k) ee[i][k] = ex2_ftz(ee[i][k]);
        }
        // ---- the particle's lanes stage g[n][k] = upstream * ((k == z_n) - q_nk) (zero outside the clamp / for dead particles)
        __syncwarp();
        float sgp = 0.f;
        for (int n0 = 0; n0 < N; n0 += kZL * kZPT) {
            // all global loads of this lane's (up to) 8 points first -- 40 independent requests in flight -- then the
            // arithmetic (the IEEE division below keeps the compiler from overlapping the iterations by itself)
            float qv[kZPT][K], up[kZPT];
            int lf[kZPT];
#pragma unroll
            for (int i = 0; i < kZPT; ++i) {
                const int n = n0 + j + kZL * i;
                const bool live = valid && n < N;
#pragma unroll
                for (int k = 0; k < K; ++k) qv[i][k] = live ? __ldg(a.q + (r * N + n) * K + k) : 1.f;
                up[i] = live ? __ldg(a.g + r * a.g_sr + n * a.g_sn) : 0.f;
                lf[i] = 0;
                if (live) {
                    if (a.zl != nullptr) {
                        lf[i] = min((int)__ldg(a.zl + r * a.LN + n), K - 1);
                    } else {
                        float best = __ldg(a.z + (r * N + n) * K);
#pragma unroll
                        for (int k = 1; k < K; ++k) { const float v = __ldg(a.z + (r * N + n) * K + k); if (v > best) { best = v; lf[i] = k; } }
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < kZPT; ++i) {
                const int n = n0 + j + kZL * i;
                float sq = 0.f;
#pragma unroll
                for (int k = 0; k < K; ++k) sq = __fadd_rn(sq, qv[i][k]);
                float ql = qv[i][0];
#pragma unroll
                for (int k = 1; k < K; ++k) if (lf[i] == k) ql = qv[i][k];
                const float pn = __fdiv_rn(ql, sq);
                const bool inside = pn >= FLT_EPS_F && pn <= 1.f - FLT_EPS_F;
                const float u = (valid && inside) ? up[i] : 0.f;
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const float gk = u * (((k == lf[i]) ? 1.f : 0.f) - qv[i][k]);
                    if (n < N) gp[n * K + k] = gk;
                    gb2 += gk;                       // u = 0 for the padding points
                    sgp += gk;
                }
            }
        }
#pragma unroll
        for (int o = kZL / 2; o > 0; o >>= 1) sgp += __shfl_xor_sync(0xffffffffu, sgp, o);     // sum of g over the particle
        __syncwarp();
        // ---- all points x this lane's hidden units
        float G2[HQ], A[HQ][K], X[HQ][D];
#pragma unroll
        for (int i = 0; i < HQ; ++i) {
            G2[i] = 0.f;
#pragma unroll
            for (int k = 0; k < K; ++k) A[i][k] = 0.f;
#pragma unroll
            for (int d = 0; d < D; ++d) X[i][d] = 0.f;
        }
        if (valid) {
            // two copies of the loop: the per-particle `fast` flag must not put a second MUFU (ex2) into every term
            if (fast) zbwd_points<K, D, HQ, true>(N, gp, s_x, s_ex + j * HQ, H8, ee, true, G2, A, X);
            else zbwd_points<K, D, HQ, false>(N, gp, s_x, s_ex + j * HQ, H8, ee, xok, G2, A, X);
        }
        // ---- fold the particle's tau / mu in and add to this thread's totals
#pragma unroll
        for (int i = 0; i < HQ; ++i) {
            float* t = tot + i * V;
            float sb = 0.f;
#pragma unroll
            for (int k = 0; k < K; ++k) sb += A[i][k];
            t[0] += fmaf(-2.f, G2[i], sgp);              // sum g tanh = sum g - 2 sum g r
            t[1] += sb;
#pragma unroll
            for (int d = 0; d < D; ++d) {
                float m = 0.f, tt = 0.f;
#pragma unroll
                for (int k = 0; k < K; ++k) { m = fmaf(A[i][k], mk[k * D + d], m); tt = fmaf(A[i][k], tk[k * D + d], tt); }
                t[2 + d] += X[i][d]; t[2 + D + d] += m; t[2 + 2 * D + d] += tt;
            }
        }
    }
    gb2 = warp_sum(gb2);
    if (lane == 0) s_gb2[threadIdx.x >> 5] = gb2;
    __syncthreads();
    // ---- the CTA's particles in a fixed order; partial row layout [h][V] + db2, as the point-major kernel
    const int GS = H * V + 1;
    float* out = a.partial + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * GS;
    for (int idx = threadIdx.x; idx < H8 * V; idx += NT) {
        const int h = idx / V, v = idx - h * V;
        if (h >= H) continue;
        const int jj = h / HQ, i = h - jj * HQ;
        float sum = 0.f;
        for (int pp = 0; pp < kZP; ++pp) sum += s_tot[(size_t)(pp * kZL + jj) * TS + i * V + v];
        out[h * V + v] = sum;
    }
    if (threadIdx.x == 0) {
        float sum = 0.f;
#pragma unroll
        for (int w = 0; w < NT / 32; ++w) sum += s_gb2[w];
        out[H * V] = sum;
    }
}

// rows [M][GS] -> [NC1][GS] doubles (CTA c sums rows c, c + NC1, ...), then one CTA finishes
__global__ void k_reduce_rows_stage1(const float* __restrict__ part, long long M, int GS, double* __restrict__ ws) {
    for (int col = threadIdx.x; col < GS; col += blockDim.x) {
        double acc = 0.0;
        for (long long row = blockIdx.x; row < M; row += gridDim.x) acc += (double)__ldg(part + row * GS + col);
        ws[(size_t)blockIdx.x * GS + col] = acc;
    }
}

template <int D>
__global__ void k_z_bwd_finish(const double* __restrict__ ws, int NC1, int H, const float* __restrict__ W2,
                               float* gW1, float* gb1, float* gW2, float* gb2) {
    constexpr int V = ZBwdV<D>::V;
    const int GS = H * V + 1;
    for (int col = threadIdx.x; col < GS; col += blockDim.x) {
        double acc = 0.0;
        for (int c = 0; c < NC1; ++c) acc += ws[(size_t)c * GS + col];
        if (col == GS - 1) { if (gb2) gb2[0] = (float)acc; continue; }
        const int h = col / V, v = col - h * V;
        const double f = 4.0 * (double)__ldg(W2 + h);                   // d tanh / d a = 4 (r - r^2), times the output weight
        if (v == 0) { if (gW2) gW2[h] = (float)acc; }
        else if (v == 1) { if (gb1) gb1[h] = (float)(f * acc); }
        else if (gW1) gW1[h * 3 * D + (v - 2)] = (float)(f * acc);      // columns: x (D), mu (D), tau (D) -- W1's own order
    }
}

constexpr int kRedCtas = 128;

template <int K, int D>
static int launch_z_bwd(ZBwdArgs a, float* gW1, float* gb1, float* gW2, float* gb2, void* ws, size_t ws_bytes, cudaStream_t st) {
    constexpr int V = ZBwdV<D>::V;
    const int GS = a.H * V + 1;
    const int NP = (a.N + kZL * kZPT - 1) / (kZL * kZPT) * (kZL * kZPT);
    const int NW = kZP * kZL / 32;
    const size_t smem = ((size_t)a.H * NP + (size_t)kZP * (a.H * K + 1) + (size_t)NP * D + (size_t)NW * GS + a.H) * 4;
    if (smem > 200 * 1024) return fail(APGS_E_SHAPE, "gmm_z_logq_backward: N=%d H=%d too large for one tile", a.N, a.H);
    const int nchunk = (a.S + kZP - 1) / kZP;
    // a few chunks per CTA amortise the data table; keep >= ~4 waves of CTAs on the 148 SMs
    int chunks = 1;
    while (chunks < 8 && (long long)a.B * ((nchunk + 2 * chunks - 1) / (2 * chunks)) >= 4LL * 4 * sm_count()) chunks *= 2;
    a.chunks = chunks;
    const int gy = (nchunk + chunks - 1) / chunks;
    if (gy > 65535) return fail(APGS_E_SHAPE, "gmm_z_logq_backward: S=%d too large for one launch", a.S);
    const long long M = (long long)a.B * gy;
    const size_t need = (size_t)M * GS * sizeof(float) + 256 + (size_t)kRedCtas * GS * sizeof(double);
    if (ws == nullptr || ws_bytes < need)
        return fail(APGS_E_WORKSPACE, "gmm_z_logq_backward: workspace of %zu bytes required, got %zu", need, ws_bytes);
    a.partial = reinterpret_cast<float*>(ws);
    double* ws2 = reinterpret_cast<double*>(reinterpret_cast<char*>(ws) + (((size_t)M * GS * sizeof(float) + 255) & ~(size_t)255));
    // hidden-unit-major variant for H <= 32 (the reference's 32): no cross-lane reductions in the main loop
    const int HQ = a.H <= 8 ? 1 : (a.H <= 16 ? 2 : 4);
    const size_t smem_t = ((size_t)a.N * HQ * kZL + (size_t)((a.N * D + 3) / 4 * 4) + (size_t)kZP * ((a.N * K + 3) / 4 * 4) +
                           (size_t)kZP * kZL * (HQ * V + 1)) * 4;
    if (a.H <= 32 && smem_t <= 200 * 1024 && !a.force_point_major) {
        void (*kern)(ZBwdArgs) = HQ == 1 ? k_gmm_z_logq_bwd_t<K, D, 1> : (HQ == 2 ? k_gmm_z_logq_bwd_t<K, D, 2> : k_gmm_z_logq_bwd_t<K, D, 4>);
        APGS_CUDA_OK(ensure_dynamic_smem(reinterpret_cast<const void*>(kern), smem_t), "smem attr");
        kern<<<dim3((unsigned int)a.B, (unsigned int)gy), kZP * kZL, smem_t, st>>>(a);
        APGS_LAUNCH_OK("k_gmm_z_logq_bwd_t launch");
    } else {
        APGS_CUDA_OK(ensure_dynamic_smem(reinterpret_cast<const void*>(k_gmm_z_logq_bwd<K, D>), smem), "smem attr");
        k_gmm_z_logq_bwd<K, D><<<dim3((unsigned int)a.B, (unsigned int)gy), kZP * kZL, smem, st>>>(a);
        APGS_LAUNCH_OK("k_gmm_z_logq_bwd launch");
    }
    k_reduce_rows_stage1<<<kRedCtas, 256, 0, st>>>(a.partial, M, GS, ws2);
    APGS_LAUNCH_OK("k_reduce_rows_stage1 launch");
    k_z_bwd_finish<D><<<1, 256, 0, st>>>(ws2, kRedCtas, a.H, a.W2, gW1, gb1, gW2, gb2);
    APGS_LAUNCH_OK("k_z_bwd_finish launch");
    return 0;
}

// ================================================================================================================
// label-indexed table sums (eta step graph)
// ================================================================================================================
struct LtsArgs {
    const float* tab;                   // (B, N, K, C)
    const float* z;                     // (R, N, K) one-hot, or null
    const unsigned char* zl;            // (R, LN) labels, or null
    long long LN;
    int S, B, N, K, C;
    float* out;                         // fwd: (R, C)
    const float* gout;                  // bwd: (R, C)
    float* gtab;                        // bwd: (B, N, K, C)
};

__device__ __forceinline__ int lts_label(const LtsArgs& a, long long r, int n) {
    if (a.zl != nullptr) return min((int)__ldg(a.zl + r * a.LN + n), a.K - 1);
    const float* zr = a.z + (r * a.N + n) * a.K;
    int lab = 0;
    float best = __ldg(zr);
    for (int k = 1; k < a.K; ++k) { const float v = __ldg(zr + k); if (v > best) { best = v; lab = k; } }
    return lab;
}

// forward: CTA = instance b x 32 particles, 8 lanes split the points; table of the instance in shared memory
template <int CMAX>
__global__ void __launch_bounds__(256) k_label_table_sum(LtsArgs a) {
    extern __shared__ __align__(16) float s_tab[];       // [N*K][C]
    const int b = blockIdx.x, C = a.C, NK = a.N * a.K;
    for (int i = threadIdx.x; i < NK * C; i += blockDim.x) s_tab[i] = __ldg(a.tab + (size_t)b * NK * C + i);
    __syncthreads();
    const int p = threadIdx.x / 8, j = threadIdx.x % 8;
    const long long s = (long long)blockIdx.y * 32 + p;
    const bool valid = s < a.S;
    const long long r = s * a.B + b;
    float acc[CMAX];
#pragma unroll
    for (int c = 0; c < CMAX; ++c) acc[c] = 0.f;
    if (valid)
        for (int n = j; n < a.N; n += 8) {
            const float* t = s_tab + (size_t)(n * a.K + lts_label(a, r, n)) * C;
#pragma unroll
            for (int c = 0; c < CMAX; ++c) if (c < C) acc[c] += t[c];
        }
#pragma unroll
    for (int c = 0; c < CMAX; ++c) {
        double v = (double)acc[c];
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (valid && j == 0 && c < C) a.out[r * C + c] = (float)v;
    }
}

// adjoint: CTA = instance b; thread t owns the pairs (n, c) = t, t + 256, ... and adds gout[s, b, c] to the row of the
// label particle s gives point n -- every (n, label, c) accumulator has exactly one owner, so no atomics; particles are
// staged through shared memory in tiles
constexpr int kLtsTile = 64;
template <int KT>                       // KT = K when K <= 5 (accumulators in registers), 0: generic
__global__ void __launch_bounds__(256) k_label_table_sum_bwd(LtsArgs a) {
    extern __shared__ __align__(16) float s_g[];          // [N*K][C] accumulators | tile: gout [kLtsTile][C], labels [kLtsTile][N]
    const int b = blockIdx.x, C = a.C, N = a.N, K = a.K, NKC = N * K * C;
    float* s_go = s_g + NKC;
    unsigned char* s_lab = reinterpret_cast<unsigned char*>(s_go + kLtsTile * C);
    for (int i = threadIdx.x; i < NKC; i += blockDim.x) s_g[i] = 0.f;
    // gridDim.y CTAs share the particles of the instance; each writes its own partial table (summed by the caller)
    const int per = (a.S + gridDim.y - 1) / gridDim.y;
    const int s_begin = blockIdx.y * per, s_end = min(a.S, s_begin + per);
    for (int s0 = s_begin; s0 < s_end; s0 += kLtsTile) {
        const int ns = min(kLtsTile, s_end - s0);
        __syncthreads();
        for (int i = threadIdx.x; i < ns * C; i += blockDim.x) {
            const int si = i / C, c = i - si * C;
            s_go[i] = __ldg(a.gout + ((long long)(s0 + si) * a.B + b) * C + c);
        }
        for (int i = threadIdx.x; i < ns * N; i += blockDim.x) {
            const int si = i / N, n = i - si * N;
            s_lab[i] = (unsigned char)lts_label(a, (long long)(s0 + si) * a.B + b, n);
        }
        __syncthreads();
        for (int pc = threadIdx.x; pc < N * C; pc += blockDim.x) {
            const int n = pc / C, c = pc - n * C;
            if (KT > 0) {
                // the K accumulators of this (n, c) in registers, a predicated add per label value: no read-modify-write
                // chain through shared memory (consecutive particles often carry the same label)
                float acc[KT > 0 ? KT : 1];
#pragma unroll
                for (int k = 0; k < KT; ++k) acc[k] = 0.f;
#pragma unroll 8
                for (int si = 0; si < ns; ++si) {
                    const int lab = s_lab[si * N + n];
                    const float g = s_go[si * C + c];
#pragma unroll
                    for (int k = 0; k < KT; ++k)
                        if (lab == k) acc[k] += g;
                }
#pragma unroll
                for (int k = 0; k < KT; ++k) s_g[(size_t)(n * K + k) * C + c] += acc[k];
            } else {
                for (int si = 0; si < ns; ++si) s_g[(size_t)(n * K + s_lab[si * N + n]) * C + c] += s_go[si * C + c];
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < NKC; i += blockDim.x) a.gtab[((size_t)blockIdx.y * a.B + b) * NKC + i] = s_g[i];
}


// ================================================================================================================
// log q(tau, mu | neural statistics) of Enc_apg_eta and its Jacobian w.r.t. the statistics (eta step graph)
//   nss[r] = { s1_k (K), s2_kd (K D), s3_kd (K D) }  ->  Normal-Gamma update (kls_gmm.py:106-119, quirks Q1 / Q2)
//   log q = sum_kd [ Gamma(alpha, beta).log_prob(tau) + Normal(m, 1/sqrt(nu tau)).log_prob(mu) ]  (objectives.py:167)
//   jac[r][c] = d log q[r] / d nss[r][c]   (zero for a count that quirk Q1 replaced)
// One thread per particle; replaces ~150 elementwise PyTorch launches (forward + autograd) per eta step.
// ================================================================================================================
__device__ __forceinline__ float digamma_pos(float x) {
    // psi(x) = psi(x + n) - sum_{i<n} 1/(x + i), then the asymptotic series at y >= 6 (next term 1/(240 y^8) < 3e-9)
    float acc = 0.f;
    while (x < 6.f) { acc += __fdiv_rn(1.f, x); x += 1.f; }
    const float r = __fdiv_rn(1.f, x), r2 = r * r;
    const float series = r2 * (0.0833333333333f - r2 * (0.00833333333333f - r2 * 0.00396825396825f));
    return logf(x) - 0.5f * r - series - acc;
}

template <int K, int D>
__global__ void __launch_bounds__(256) k_gmm_eta_logq_jac(const float* __restrict__ nss, const float* __restrict__ tau,
                                                          const float* __restrict__ mu, const float* __restrict__ ra,
                                                          const float* __restrict__ rb, const float* __restrict__ rm,
                                                          const float* __restrict__ rn, long long R,
                                                          float* __restrict__ lq_out, float* __restrict__ jac) {
    constexpr int KD = K * D, C = K + 2 * KD;
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    const float* s = nss + r * C;
    float lq = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const float s1 = __ldg(s + k);
        const bool replaced = s1 == 0.f;                                     // Q1
        const float c = replaced ? 1.f : s1;
        float gc = 0.f;
#pragma unroll
        for (int d = 0; d < D; ++d) {
            const int e = k * D + d;
            const float s2 = __ldg(s + K + e), s3 = __ldg(s + K + KD + e);
            const float a0 = __ldg(ra + e), b0 = __ldg(rb + e), m0 = __ldg(rm + e), n0 = __ldg(rn + e);
            const float t = __ldg(tau + r * KD + e), u = __ldg(mu + r * KD + e);
            const float cn = c + n0, xb = s2 / c, A = c * n0 / cn, dx = xb - n0;      // Q2: (x_bar - prior_nu)
            const float alpha = a0 + 0.5f * c, nu = n0 + c, m = (m0 * n0 + s2) / cn;
            const float beta = b0 + 0.5f * (s3 - s2 * s2 / c) + 0.5f * A * dx * dx;
            const float lbeta = logf(beta), lt = logf(t), dm_ = u - m;
            lq += (alpha * lbeta + (alpha - 1.f) * lt - beta * t - lgammaf(alpha))
                  + (-0.5f * dm_ * dm_ * nu * t + 0.5f * logf(nu * t) - LOG_SQRT_2PI_F);
            const float g_alpha = lbeta + lt - digamma_pos(alpha);
            const float g_beta = alpha / beta - t;
            const float g_m = dm_ * nu * t;
            const float g_nu = -0.5f * dm_ * dm_ * t + 0.5f / nu;
            jac[r * C + K + KD + e] = 0.5f * g_beta;                                               // d / d s3
            jac[r * C + K + e] = g_beta * (-xb + A * dx / c) + g_m / cn;                           // d / d s2
            gc += 0.5f * g_alpha + g_nu - g_m * m / cn
                  + g_beta * (0.5f * s2 * s2 / (c * c) + 0.5f * (n0 * n0 / (cn * cn)) * dx * dx - A * dx * xb / c);
        }
        jac[r * C + k] = replaced ? 0.f : gc;                                                      // d / d s1
    }
    lq_out[r] = lq;
}
}  // namespace apgs

using namespace apgs;

extern "C" size_t apgs_gmm_z_logq_backward_workspace_bytes(int64_t S, int64_t B, int64_t H, int64_t D) {
    if (S <= 0 || B <= 0 || H <= 0 || D <= 0) return 0;
    const size_t GS = (size_t)H * (2 + 3 * D) + 1;
    const size_t M = (size_t)B * ((S + kZP - 1) / kZP);            // upper bound (one chunk per CTA)
    return ((M * GS * sizeof(float) + 255) & ~(size_t)255) + 256 + (size_t)kRedCtas * GS * sizeof(double);
}

extern "C" int apgs_gmm_z_logq_backward_f32(const float* x, const float* tau, const float* mu, const float* W1,
                                            const float* b1, const float* W2, int64_t H, const float* q_probs,
                                            const float* z_onehot, const uint8_t* z_labels, int64_t label_stride,
                                            const float* grad_logq, int64_t g_stride_particle, int64_t g_stride_point,
                                            int64_t S, int64_t B, int64_t N, int64_t K, int64_t D, float* gW1, float* gb1,
                                            float* gW2, float* gb2, void* workspace, size_t workspace_bytes, void* stream) {
    if (S <= 0 || B <= 0 || N <= 0 || H <= 0 || H > 1024 || S >= (1ll << 31) || B >= 65536ll * 32768)
        return fail(APGS_E_SHAPE, "gmm_z_logq_backward: bad sizes S=%lld B=%lld N=%lld H=%lld", (long long)S, (long long)B,
                    (long long)N, (long long)H);
    if (!x || !tau || !mu || !W1 || !b1 || !W2 || !q_probs || !grad_logq) return fail(APGS_E_BADARG, "gmm_z_logq_backward: NULL input");
    if ((z_onehot == nullptr) == (z_labels == nullptr))
        return fail(APGS_E_BADARG, "gmm_z_logq_backward: give the assignment either one-hot or as labels");
    if (z_labels && label_stride < N) return fail(APGS_E_BADARG, "gmm_z_logq_backward: label_stride %lld < N", (long long)label_stride);
    ZBwdArgs a = {};
    a.x = x; a.tau = tau; a.mu = mu; a.W1 = W1; a.b1 = b1; a.W2 = W2; a.q = q_probs; a.z = z_onehot; a.zl = z_labels;
    a.LN = label_stride; a.g = grad_logq; a.g_sr = g_stride_particle; a.g_sn = g_stride_point;
    a.S = (int)S; a.B = (int)B; a.N = (int)N; a.H = (int)H;
    {   // test hook: APGS_ZBWD_POINT_MAJOR=1 runs the generic point-major kernel even when H <= 32
        const char* e = getenv("APGS_ZBWD_POINT_MAJOR");
        a.force_point_major = (e != nullptr && e[0] == '1') ? 1 : 0;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (K == 3 && D == 2) return launch_z_bwd<3, 2>(a, gW1, gb1, gW2, gb2, workspace, workspace_bytes, st);
    if (K == 2 && D == 2) return launch_z_bwd<2, 2>(a, gW1, gb1, gW2, gb2, workspace, workspace_bytes, st);
    if (K == 4 && D == 2) return launch_z_bwd<4, 2>(a, gW1, gb1, gW2, gb2, workspace, workspace_bytes, st);
    if (K == 5 && D == 2) return launch_z_bwd<5, 2>(a, gW1, gb1, gW2, gb2, workspace, workspace_bytes, st);
    if (K == 3 && D == 3) return launch_z_bwd<3, 3>(a, gW1, gb1, gW2, gb2, workspace, workspace_bytes, st);
    return fail(APGS_E_UNSUPPORTED, "gmm_z_logq_backward: (K=%lld, D=%lld) not instantiated", (long long)K, (long long)D);
}

static int lts_check(const LtsArgs& a, const char* who) {
    if (a.S <= 0 || a.B <= 0 || a.N <= 0 || a.K <= 0 || a.K > 255 || a.C <= 0 || a.C > 64)
        return fail(APGS_E_SHAPE, "%s: bad sizes S=%d B=%d N=%d K=%d C=%d", who, a.S, a.B, a.N, a.K, a.C);
    if (!a.tab && !a.gtab) return fail(APGS_E_BADARG, "%s: NULL table", who);
    if ((a.z == nullptr) == (a.zl == nullptr)) return fail(APGS_E_BADARG, "%s: give the assignment either one-hot or as labels", who);
    if (a.zl && a.LN < a.N) return fail(APGS_E_BADARG, "%s: label_stride < N", who);
    return 0;
}

extern "C" int apgs_gmm_label_table_sum_f32(const float* table, const float* z_onehot, const uint8_t* z_labels,
                                            int64_t label_stride, int64_t S, int64_t B, int64_t N, int64_t K, int64_t C,
                                            float* out, void* stream) {
    LtsArgs a = {};
    a.tab = table; a.z = z_onehot; a.zl = z_labels; a.LN = label_stride;
    a.S = (int)S; a.B = (int)B; a.N = (int)N; a.K = (int)K; a.C = (int)C; a.out = out;
    if (int e = lts_check(a, "gmm_label_table_sum")) return e;
    if (!table || !out) return fail(APGS_E_BADARG, "gmm_label_table_sum: NULL pointer");
    const size_t smem = (size_t)N * K * C * 4;
    if (smem > 200 * 1024) return fail(APGS_E_SHAPE, "gmm_label_table_sum: table of %zu bytes does not fit shared memory", smem);
    const long long gy = (S + 31) / 32;
    if (gy > 65535) return fail(APGS_E_SHAPE, "gmm_label_table_sum: S too large for one launch");
    cudaStream_t st = (cudaStream_t)stream;
    if (C <= 16) {
        APGS_CUDA_OK(ensure_dynamic_smem(reinterpret_cast<const void*>(k_label_table_sum<16>), smem), "smem attr");
        k_label_table_sum<16><<<dim3((unsigned int)B, (unsigned int)gy), 256, smem, st>>>(a);
    } else {
        APGS_CUDA_OK(ensure_dynamic_smem(reinterpret_cast<const void*>(k_label_table_sum<64>), smem), "smem attr");
        k_label_table_sum<64><<<dim3((unsigned int)B, (unsigned int)gy), 256, smem, st>>>(a);
    }
    APGS_LAUNCH_OK("k_label_table_sum launch");
    return 0;
}

extern "C" int apgs_gmm_label_table_sum_backward_splits(int64_t S, int64_t B) {
    // enough CTAs for ~4 waves on the SMs, at least one tile of particles each
    long long sp = 1;
    while (sp < 16 && B * sp < 4LL * sm_count() && S / (sp * 2) >= kLtsTile) sp *= 2;
    return (int)sp;
}

extern "C" int apgs_gmm_label_table_sum_backward_f32(const float* grad_out, const float* z_onehot, const uint8_t* z_labels,
                                                     int64_t label_stride, int64_t S, int64_t B, int64_t N, int64_t K,
                                                     int64_t C, float* grad_table, void* stream) {
    LtsArgs a = {};
    a.gout = grad_out; a.z = z_onehot; a.zl = z_labels; a.LN = label_stride;
    a.S = (int)S; a.B = (int)B; a.N = (int)N; a.K = (int)K; a.C = (int)C; a.gtab = grad_table;
    if (int e = lts_check(a, "gmm_label_table_sum_backward")) return e;
    if (!grad_out || !grad_table) return fail(APGS_E_BADARG, "gmm_label_table_sum_backward: NULL pointer");
    const size_t smem = ((size_t)N * K * C + (size_t)kLtsTile * C) * 4 + (size_t)kLtsTile * N;
    if (smem > 200 * 1024) return fail(APGS_E_SHAPE, "gmm_label_table_sum_backward: N K C too large for shared memory");
    void (*kern)(LtsArgs) = K == 2 ? k_label_table_sum_bwd<2> : K == 3 ? k_label_table_sum_bwd<3> : K == 4 ? k_label_table_sum_bwd<4>
                            : K == 5 ? k_label_table_sum_bwd<5> : k_label_table_sum_bwd<0>;
    APGS_CUDA_OK(ensure_dynamic_smem(reinterpret_cast<const void*>(kern), smem), "smem attr");
    const int splits = apgs_gmm_label_table_sum_backward_splits(S, B);     // grad_table holds `splits` partial tables
    kern<<<dim3((unsigned int)B, (unsigned int)splits), 256, smem, (cudaStream_t)stream>>>(a);
    APGS_LAUNCH_OK("k_label_table_sum_bwd launch");
    return 0;
}

extern "C" int apgs_gmm_eta_logq_jac_f32(const float* nss, const float* tau, const float* mu, const float* prior_alpha,
                                         const float* prior_beta, const float* prior_mu, const float* prior_nu, int64_t S,
                                         int64_t B, int64_t K, int64_t D, float* log_q, float* jac, void* stream) {
    if (S <= 0 || B <= 0 || S * B >= (1ll << 40)) return fail(APGS_E_SHAPE, "gmm_eta_logq_jac: bad sizes S=%lld B=%lld", (long long)S, (long long)B);
    if (!nss || !tau || !mu || !prior_alpha || !prior_beta || !prior_mu || !prior_nu || !log_q || !jac)
        return fail(APGS_E_BADARG, "gmm_eta_logq_jac: NULL pointer");
    const long long R = S * B;
    const unsigned int grid = (unsigned int)((R + 255) / 256);
    cudaStream_t st = (cudaStream_t)stream;
#define APGS_ELQ(K_, D_)                                                                                                 \
    if (K == K_ && D == D_) {                                                                                            \
        k_gmm_eta_logq_jac<K_, D_><<<grid, 256, 0, st>>>(nss, tau, mu, prior_alpha, prior_beta, prior_mu, prior_nu, R, log_q, jac); \
        APGS_LAUNCH_OK("k_gmm_eta_logq_jac launch");                                                                     \
        return 0;                                                                                                        \
    }
    APGS_ELQ(3, 2) APGS_ELQ(2, 2) APGS_ELQ(4, 2) APGS_ELQ(5, 2) APGS_ELQ(3, 3) APGS_ELQ(4, 3) APGS_ELQ(3, 1)
#undef APGS_ELQ
    return fail(APGS_E_UNSUPPORTED, "gmm_eta_logq_jac: (K=%lld, D=%lld) not instantiated", (long long)K, (long long)D);
}
