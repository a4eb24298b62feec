// GMM reweighting kernels for sm_100a.
//
//   k_gmm_nss_posterior  G1: neural sufficient statistics -> Normal-Gamma posterior parameters
//                            (apgs/gmm/kls_gmm.py:92-119, quirks Q1/Q2)
//   k_gmm_weights        G2/G3/G4: log importance weights of the eta step, the z step, the joint
//                            step and the one-shot (apgs/gmm/objectives.py:74-227,
//                            apgs/gmm/models.py:139-185; torch.distributions formulas restated)
//
// Both are HBM-bound streams over particles r = s*B + b whose rows are contiguous in memory.
// A CTA owns a tile of P consecutive particles: the large per-particle rows (z one-hot, q_probs,
// gamma, x') are staged into shared memory by the bulk-copy engine (tile.cuh) while the threads
// evaluate the small (K,D) terms; LANES lanes then split the N points of each particle.
// fp32 arithmetic in the reference's operation order; N-sums are accumulated in fp64.
#include "gmm_common.cuh"
#include "rng.cuh"

namespace apgs {

// ============================================================================================
// G1
// ============================================================================================
struct NssArgs {
    const float* xp;      // (R, N, D)          (un-fused: outputs of the two Linear heads)
    const float* gamma;   // (R, N, K)
    // fused heads (gamma == nullptr): gamma = softmax(Wg [x;z] + bg), xp = Wo [x;z] + bo computed in-kernel
    const float* x;       // (B, N, D)
    const float* z;       // (R, N, K) or nullptr (Enc_rws_eta: heads see x only)
    const float *Wg, *bg, *Wo, *bo;   // (K, D[+K]), (K), (D, D[+K]), (D)
    int B;
    const float *ra, *rb, *rm, *rn;   // prior alpha, beta, mu, nu  (K, D)
    float *oa, *ob, *om, *on;         // posterior (R, K, D)
    long long R;
    int N;
    // STATS: also emit the plain one-hot statistics of z centred at `center` (R, K, D):
    //   stats[r][k] = { sum_n z_nk,  sum_n z_nk (x_nd - c_kd)  (d < D),  sum_n z_nk (x_nd - c_kd)^2  (d < D) }
    // from which the eta-step log-likelihoods follow in O(K D) per particle (k_gmm_eta_weights_stats)
    const float* center;
    float* stats;         // (R, K, 1 + 2 D)
    const unsigned char* zl;   // (R, N) labels instead of the one-hot z (compact latents inside a sweep), or null
    long long LN;              // bytes between the label rows of consecutive particles (>= N)
    // deferred gathers (k_gmm_nss_im, label mode): particle (s,b) reads its label row from row (z_anc[s,b], b) and its centre
    // from row (c_anc[s,b], b) -- ancestor maps (S,B) int32 of resampling steps that were never materialised -- or null
    const int *z_anc, *c_anc;
    // per-instance head tables made once per objective by k_gmm_eta_tables ([N K][RS] factor rows + [N][D] points, padded to
    // 4 floats per instance), or null: the CTA tabulates them itself
    const float* tabs;
};

template <int K, int D, int P, int LANES, bool FUSED, bool HAS_Z, bool STATS = false>
__global__ void __launch_bounds__(P* LANES, STATS ? 4 : 1) k_gmm_nss_posterior(NssArgs a) {
    static_assert(!STATS || (FUSED && HAS_Z), "plain statistics need the raw x and z");
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ unsigned long long bar;
    __shared__ float s_wg[K * (D + K) + K], s_wo[D * (D + K) + D];
    const int N = a.N;
    constexpr bool fused = FUSED, has_z = HAS_Z;
    constexpr int IN = D + (HAS_Z ? K : 0);                      // input width of the heads
    float* s_g = reinterpret_cast<float*>(smem);                 // P*N*K  (gamma, or z when fused)
    float* s_x = s_g + (size_t)P * N * K;                        // P*N*D  (xp; unused when fused)
    const long long r0 = (long long)blockIdx.x * P;
    const int rows = (int)min((long long)P, a.R - r0);
    if (threadIdx.x == 0) mbar_init(&bar, 1);
    __syncthreads();
    StageReq req[2] = {{s_g, (fused ? (has_z ? a.z : a.x) : a.gamma) + ((fused && !has_z) ? 0 : r0 * N * K),
                        (fused && !has_z) ? 0u : (unsigned int)(rows * N * K * 4)},
                       {s_x, (fused ? a.x : a.xp) + (fused ? 0 : r0 * N * D), fused ? 0u : (unsigned int)(rows * N * D * 4)}};
    stage_issue(req, 2, &bar);
    if (fused) {
        for (int i = threadIdx.x; i < K * IN; i += blockDim.x) s_wg[i] = __ldg(a.Wg + i);
        for (int i = threadIdx.x; i < K; i += blockDim.x) s_wg[K * (D + K) + i] = __ldg(a.bg + i);
        for (int i = threadIdx.x; i < D * IN; i += blockDim.x) s_wo[i] = __ldg(a.Wo + i);
        for (int i = threadIdx.x; i < D; i += blockDim.x) s_wo[D * (D + K) + i] = __ldg(a.bo + i);
    }
    const int p = threadIdx.x / LANES, j = threadIdx.x % LANES;
    const bool valid = p < rows;
    stage_wait(&bar, 0);

    // per-lane partial sums (<= ceil(N/LANES) terms) in fp32, combined across lanes in fp64: fewer roundings than
    // the reference's own fp32 reduction over all N terms, and no fp64 converts per point
    float s1[K], s2[K][D], s3[K][D];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        s1[k] = 0.f;
#pragma unroll
        for (int d = 0; d < D; ++d) { s2[k][d] = 0.f; s3[k][d] = 0.f; }
    }
    float c1[STATS ? K : 1], c2[STATS ? K : 1][D], c3[STATS ? K : 1][D], cen[STATS ? K : 1][D];
    f32x2 c2p[STATS ? K : 1], c3p[STATS ? K : 1], ncen2[STATS ? K : 1];       // D == 2: (d0, d1) pairs
    if (STATS) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            c1[k] = 0.f;
#pragma unroll
            for (int d = 0; d < D; ++d) {
                c2[k][d] = 0.f; c3[k][d] = 0.f;
                cen[k][d] = valid ? __ldg(a.center + (r0 + p) * (K * D) + k * D + d) : 0.f;
            }
            c2p[k] = pack2(0.f, 0.f); c3p[k] = c2p[k];
            ncen2[k] = pack2(-cen[k][0], -cen[k][D - 1]);
        }
    }
    if (valid) {
        const float* g = s_g + (size_t)p * N * K;
        const float* x = s_x + (size_t)p * N * D;
        const float* xb = fused ? a.x + (size_t)((r0 + p) % a.B) * N * D : nullptr;
        for (int n = j; n < N; n += LANES) {
            float xv[D], gv[K];
            if (!fused) {
#pragma unroll
                for (int d = 0; d < D; ++d) xv[d] = x[n * D + d];
#pragma unroll
                for (int k = 0; k < K; ++k) gv[k] = g[n * K + k];
            } else {
                // the two Linear heads on cat(x_n, z_n) (models.py:17-22,56-65) and the softmax over k
                float in[IN];
#pragma unroll
                for (int d = 0; d < D; ++d) in[d] = __ldg(xb + n * D + d);
                if (HAS_Z) {
#pragma unroll
                    for (int k = 0; k < K; ++k) in[D + (HAS_Z ? k : 0)] = g[n * K + k];
                }
                if (STATS) {
                    // indicator of the label the reference scores (first maximum, models.py:178), the point centred
                    // at ITS cluster's centre, then K masked accumulations (packed fp32x2 when D == 2)
                    int lab = 0;
                    float best = in[D];
#pragma unroll
                    for (int k = 1; k < K; ++k) { const float v = in[D + (HAS_Z ? k : 0)]; if (v > best) { best = v; lab = k; } }
                    float ind[K];
#pragma unroll
                    for (int k = 0; k < K; ++k) { ind[k] = (lab == k) ? 1.f : 0.f; c1[k] += ind[k]; }
                    if (D == 2) {
                        f32x2 t = pack2(in[0], in[1]);
#pragma unroll
                        for (int k = 0; k < K; ++k) t = fma2(pack2(ind[k], ind[k]), ncen2[k], t);      // x - c_lab
                        const f32x2 t2 = mul2(t, t);
#pragma unroll
                        for (int k = 0; k < K; ++k) {
                            const f32x2 ik = pack2(ind[k], ind[k]);
                            c2p[k] = fma2(ik, t, c2p[k]);
                            c3p[k] = fma2(ik, t2, c3p[k]);
                        }
                    } else {
#pragma unroll
                        for (int d = 0; d < D; ++d) {
                            float c = 0.f;
#pragma unroll
                            for (int k = 0; k < K; ++k) c = fmaf(ind[k], cen[k][d], c);
                            const float t = in[d] - c, t2 = t * t;
#pragma unroll
                            for (int k = 0; k < K; ++k) { c2[k][d] = fmaf(ind[k], t, c2[k][d]); c3[k][d] = fmaf(ind[k], t2, c3[k][d]); }
                        }
                    }
                }
                float lg[K], mx = -INFINITY;
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    float acc = s_wg[K * (D + K) + k];
#pragma unroll
                    for (int i = 0; i < IN; ++i) acc = fmaf(s_wg[k * IN + i], in[i], acc);
                    lg[k] = acc;
                    mx = fmaxf(mx, acc);
                }
                float se = 0.f;
#pragma unroll
                for (int k = 0; k < K; ++k) { lg[k] = ex2_ftz((lg[k] - mx) * 1.4426950408889634f); se += lg[k]; }
                const float rse = rcp_ftz(se);
#pragma unroll
                for (int k = 0; k < K; ++k) gv[k] = lg[k] * rse;
#pragma unroll
                for (int d = 0; d < D; ++d) {
                    float acc = s_wo[D * (D + K) + d];
#pragma unroll
                    for (int i = 0; i < IN; ++i) acc = fmaf(s_wo[d * IN + i], in[i], acc);
                    xv[d] = acc;
                }
            }
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const float gk = gv[k];
                s1[k] = __fadd_rn(s1[k], gk);
#pragma unroll
                for (int d = 0; d < D; ++d) {
                    s2[k][d] = __fadd_rn(s2[k][d], __fmul_rn(gk, xv[d]));                      // z_expand * ob_expand
                    s3[k][d] = __fadd_rn(s3[k][d], __fmul_rn(gk, __fmul_rn(xv[d], xv[d])));    // z_expand * ob_expand**2
                }
            }
        }
    }
    double t1[K], t2[K][D], t3[K][D];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        t1[k] = group_sum<LANES>((double)s1[k]);
#pragma unroll
        for (int d = 0; d < D; ++d) {
            t2[k][d] = group_sum<LANES>((double)s2[k][d]);
            t3[k][d] = group_sum<LANES>((double)s3[k][d]);
        }
    }
    if (STATS) {
        // <= ceil(N/LANES) fp32 terms per lane, centred; the LANES partials of a particle meet in shared memory
        constexpr int ST = 1 + 2 * D, NS = K * ST, NSP = NS | 1;
        __syncthreads();                                      // every warp is done with the z tile: reuse its memory
        float* s_red = reinterpret_cast<float*>(smem);        // [P * LANES][NSP] (the launcher sizes for it)
        float* mine = s_red + (size_t)threadIdx.x * NSP;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            if (D == 2) { unpack2(c2p[k], c2[k][0], c2[k][D - 1]); unpack2(c3p[k], c3[k][0], c3[k][D - 1]); }
            mine[k * ST] = c1[k];
#pragma unroll
            for (int d = 0; d < D; ++d) { mine[k * ST + 1 + d] = c2[k][d]; mine[k * ST + 1 + D + d] = c3[k][d]; }
        }
        __syncwarp();
        if (valid) {
            float* so = a.stats + (r0 + p) * NS;
            for (int e = j; e < NS; e += LANES) {
                float v = 0.f;
#pragma unroll
                for (int i = 0; i < LANES; ++i) v += s_red[(size_t)(p * LANES + i) * NSP + e];
                so[e] = v;
            }
        }
    }
    if (!valid) return;
    // lanes 0..K*D-1 finish one (k,d) element each (kls_gmm.py:110-118, fp32, reference order)
    for (int e = j; e < K * D; e += LANES) {
        const int k = e / D, d = e % D;
        float n1 = 0.f, n2 = 0.f, n3 = 0.f;
#pragma unroll
        for (int kk = 0; kk < K; ++kk)
#pragma unroll
            for (int dd = 0; dd < D; ++dd)
                if (kk == k && dd == d) { n1 = (float)t1[kk]; n2 = (float)t2[kk][dd]; n3 = (float)t3[kk][dd]; }
        const float a0 = __ldg(a.ra + e), b0 = __ldg(a.rb + e), m0 = __ldg(a.rm + e), n0 = __ldg(a.rn + e);
        const float cnt = (n1 == 0.f) ? 1.f : n1;                                   // Q1
        const float x_bar = __fdiv_rn(n2, cnt);
        const float alpha = __fadd_rn(a0, __fmul_rn(cnt, 0.5f));
        const float nu = __fadd_rn(n0, cnt);
        const float cn = __fadd_rn(cnt, n0);
        const float mu = __fdiv_rn(__fadd_rn(__fmul_rn(m0, n0), n2), cn);
        const float u1 = __fmul_rn(__fsub_rn(n3, __fdiv_rn(__fmul_rn(n2, n2), cnt)), 0.5f);
        const float dx = __fsub_rn(x_bar, n0);                                       // Q2: prior_nu
        const float u2 = __fmul_rn(__fmul_rn(__fdiv_rn(__fmul_rn(cnt, n0), cn), __fmul_rn(dx, dx)), 0.5f);
        const float beta = __fadd_rn(__fadd_rn(b0, u1), u2);
        const long long o = (r0 + p) * (K * D) + e;
        a.oa[o] = alpha; a.ob[o] = beta; a.om[o] = mu; a.on[o] = nu;
    }
}

// ============================================================================================
// G1 for the fused heads with one-hot z (Enc_apg_eta), instance-major.
// A CTA owns ONE instance b and P particles of it (rows r = s*B + b, one bulk copy per 720-byte z row), as the fused
// z step does.  With z one-hot, the two Linear heads on cat(x_n, z_n) -- and therefore gamma_n, x'_n and every product
// gamma_nk x'_nd, gamma_nk x'_nd^2 that enters the neural statistics -- take only K distinct values per data point:
// the factors are tabulated ONCE PER CTA in shared memory, T[n][label] = { x' (D), gamma (K) }, and the per-particle
// work shrinks to: label of the point, two loads, K (1 + 2D) accumulating FMAs in packed fp32x2 (+ the centred
// one-hot statistics for the eta-step weights).  ~30 instructions per point instead of ~100.
// ============================================================================================
constexpr int kP = 32, kLanes = 8;

// row stride (floats) of the staged z rows: a multiple of 4 (16-byte bulk-copy destinations) that is 8 or 24 mod 32, so
// that the four particles of a warp reading z[n][k] (n = lane-in-particle + 8 it) hit 32 different banks
__host__ __device__ inline int nss_im_zstride(int NK) {
    int st = (NK + 3) / 4 * 4;
    while ((st & 31) != 8 && (st & 31) != 24) st += 4;
    return st;
}

template <int K, int D, int P, int LANES, bool STATS, int NC = 0>   // NC: points per instance when known at compile time
__global__ void __launch_bounds__(P* LANES, 4) k_gmm_nss_im(NssArgs a, int S) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ unsigned long long bar;
    constexpr int KD = K * D;
    constexpr int TE = EtaTab<K, D>::TE;                                // table planes: gamma, gamma x', gamma x'^2
    constexpr int RW = EtaTab<K, D>::RW;                                // reduction row (odd stride: no bank conflicts)
    const int N = NC ? NC : a.N, NK = N * K;
    const int ZS = nss_im_zstride(NK);
    // z rows, later the cross-lane reduction buffer.  Label mode stages 1-byte labels instead of one-hot rows (P * LS bytes), so
    // the region is the reduction buffer alone: 27 instead of 58 KB per CTA at the reference shape -- 4 CTAs per SM, not 3.
    const size_t zreg = a.zl != nullptr ? max(((size_t)P * ((N + 15) & ~15) + 3) / 4, (size_t)P * LANES * RW)
                                        : max((size_t)P * ZS, (size_t)P * LANES * RW);
    float* s_z = reinterpret_cast<float*>(smem);
    // The table holds the FACTORS, one short row per (point, label): { x'_0..x'_{D-1}, gamma_0..gamma_{K-1} }, and the
    // products gamma_k x'_d, gamma_k x'_d^2 are formed by the FMAs that accumulate them (FMA pipe has headroom, the
    // LSU does not): K + D loaded values per point (one LDS.128 + one LDS.32 at K = 3, D = 2) and K(1 + 2D)
    // accumulating FMAs, instead of the K(1 + 2D) = 15 LDS.32 + 15 FADD of a table of products (0.37 ms, LSU-bound;
    // an array-of-structs row of 16 products cost ~2.3x the wavefronts, pairs of product planes were slower too).
    constexpr int RS = EtaTab<K, D>::RS;
    static_assert(RS <= TE, "the launcher sizes the table region as TE floats per row");
    float* s_tab = s_z + (zreg + 3) / 4 * 4;                            // [NK][RS] (region sized [NK][TE])
    float* s_x = s_tab + (size_t)RS * NK;                               // [N][D], right behind the rows (one block with them)
    const unsigned int tab_bytes = a.tabs != nullptr ? (unsigned int)((RS * NK + N * D + 3) / 4 * 4) * 4u : 0u;
    const int b = (int)(blockIdx.x % (unsigned int)a.B);
    const int s0 = (int)(blockIdx.x / (unsigned int)a.B) * P;
    const int p = threadIdx.x / LANES, j = threadIdx.x % LANES;
    const bool valid = s0 + p < S;
    const long long r = (long long)(s0 + p) * a.B + b;
    const long long rz = (valid && a.z_anc != nullptr) ? (long long)__ldg(a.z_anc + r) * a.B + b : r;     // row of the labels
    const long long rc = (valid && a.c_anc != nullptr) ? (long long)__ldg(a.c_anc + r) * a.B + b : r;     // row of the centre
    const int rows = min(P, S - s0);
    const unsigned int rb = (unsigned int)NK * 4u;
    const bool labels = a.zl != nullptr;
    const bool bulk_in = !labels && (rb & 15u) == 0 && aligned(a.z, 16);
    unsigned char* s_lab = reinterpret_cast<unsigned char*>(s_z);       // [P][LS] in label mode (the z region is free then)
    const int LS = (N + 15) & ~15;                                      // 16-byte rows: one LDG.128 per 16 labels when the
    const bool lab_vec = labels && (a.LN & 15) == 0 && a.LN >= LS && aligned(a.zl, 16);   // caller pads its rows likewise
    if (threadIdx.x == 0) mbar_init(&bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bar, (bulk_in ? rb * (unsigned int)rows : 0u) + tab_bytes);
        if (tab_bytes) bulk_g2s(s_tab, a.tabs + (size_t)b * (tab_bytes / 4), tab_bytes, &bar);    // the instance's tables: one copy
    }
    if (valid) {
        if (labels) {
            if (lab_vec) {
                // cp.async: no register in between, so the table build below overlaps the HBM latency of the rows
                const unsigned char* src = a.zl + rz * a.LN;
                for (int i = j; i < LS / 16; i += LANES)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(s_lab + (size_t)p * LS + i * 16)),
                                 "l"(src + i * 16) : "memory");
                asm volatile("cp.async.commit_group;" ::: "memory");
            } else {
                for (int i = j; i < N; i += LANES) s_lab[(size_t)p * LS + i] = __ldg(a.zl + rz * a.LN + i);
            }
        } else if (bulk_in) {
            if (j == 0) bulk_g2s(s_z + (size_t)p * ZS, a.z + r * NK, rb, &bar);
        } else {
            for (int i = j; i < NK; i += LANES) s_z[(size_t)p * ZS + i] = __ldg(a.z + r * NK + i);
        }
    }
    // ---- per-CTA table over (point, label), while the rows are in flight
    const float* xb = a.x + (size_t)b * N * D;
    for (int idx = threadIdx.x; idx < (a.tabs != nullptr ? 0 : NK); idx += blockDim.x) {
        const int n = idx / K, lab = idx % K;
        float xin[D];
#pragma unroll
        for (int d = 0; d < D; ++d) xin[d] = __ldg(xb + n * D + d);
        if (lab == 0) {
#pragma unroll
            for (int d = 0; d < D; ++d) s_x[n * D + d] = xin[d];
        }
        eta_head_factors<K, D>(a.Wg, a.bg, a.Wo, a.bo, xin, lab, s_tab + (size_t)idx * RS);   // models.py:56-65
    }
    EtaNssAcc<K, D, STATS> acc;
    {
        float cen[K][D];
#pragma unroll
        for (int k = 0; k < K; ++k)
#pragma unroll
            for (int d = 0; d < D; ++d) cen[k][d] = (STATS && valid) ? __ldg(a.center + rc * KD + k * D + d) : 0.f;
        acc.init(cen);
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");        // this thread's label rows (no-op otherwise)
    stage_wait(&bar, 0);                                        // rows landed; table and s_x visible

    if (valid) {
        const float* zr = s_z + (size_t)p * ZS;
        // left to the compiler's own unrolling: `#pragma unroll 8` (N = 60) measured 0.44 ms, `unroll 2` 0.351, default 0.344
        for (int n = j; n < N; n += LANES) {
            int lab = 0;
            if (labels) {
                lab = min((int)s_lab[(size_t)p * LS + n], K - 1);
            } else {
                float best = zr[n * K];
#pragma unroll
                for (int k = 1; k < K; ++k) { const float v = zr[n * K + k]; if (v > best) { best = v; lab = k; } }
            }
            acc.add(s_tab + (size_t)(n * K + lab) * RS, lab, s_x + n * D);
        }
    }
    // ---- the LANES partials of a particle meet in shared memory (the z rows are no longer needed)
    __syncthreads();
    eta_nss_finish<K, D, LANES, STATS>(acc, s_z + (size_t)p * LANES * RW, j, valid, r, a.ra, a.rb, a.rm, a.rn,
                                       a.oa, a.ob, a.om, a.on, a.stats);
}

// ============================================================================================
// G2 / G3 / G4
// ============================================================================================
enum { STEP_ONESHOT = 0, STEP_ETA = 1, STEP_Z = 2, STEP_JOINT = 3 };

struct GmmWArgs {
    const float* x;                   // (B, N, D)   -- read once per instance, not S-replicated (F9)
    const float* z_f;                 // (R, N, K) assignment of the forward term
    const float* z_b;                 // (R, N, K) assignment of the backward term (== z_f for ETA)
    const float* q;                   // (R, N, K) q_probs or null
    const float *pa, *pb, *pm, *pn;   // posterior (R, K, D) or null
    const float *ra, *rb, *rm, *rn;   // prior (K, D)
    const float* rpi;                 // prior pi (K)
    const float *tau_f, *mu_f, *tau_b, *mu_b;   // (R, K, D)
    long long R;
    int B, N, step;
    float *log_w, *log_q_f, *dens, *log_w_f, *log_w_b;   // (R), each optional
    float *lw_local, *lq_local;                           // (R, N), optional (Z step)
};

template <int K, int D, int P, int LANES, int STEP>
__global__ void __launch_bounds__(P* LANES) k_gmm_weights(GmmWArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ unsigned long long bar;
    constexpr int KD = K * D;
    const int N = a.N;
    constexpr int step = STEP;
    constexpr bool has_eta = step != STEP_Z, has_z = step != STEP_ETA, has_bwd = step != STEP_ONESHOT;
    const bool zb_distinct = has_bwd && (a.z_b != a.z_f);
    const size_t tile_f = (size_t)P * N * K;
    float* s_zf = reinterpret_cast<float*>(smem);
    float* s_zb = zb_distinct ? s_zf + tile_f : s_zf;
    float* s_q = s_zf + tile_f * (zb_distinct ? 2 : 1);
    float* s_tab = s_q + (has_z ? tile_f : 0);          // [P][2][KD][3]: mu, 1/(2 var), log sigma + c
    const long long r0 = (long long)blockIdx.x * P;
    const int rows = (int)min((long long)P, a.R - r0);
    if (threadIdx.x == 0) mbar_init(&bar, 1);
    __syncthreads();
    const unsigned int tb = (unsigned int)(rows * N * K * 4);
    StageReq req[3] = {{s_zf, a.z_f + r0 * N * K, tb},
                       {s_zb, a.z_b + r0 * N * K, zb_distinct ? tb : 0u},
                       {s_q, a.q + r0 * N * K, has_z ? tb : 0u}};
    stage_issue(req, 3, &bar);

    // ---- while the tile is in flight: the (K,D) terms of this particle
    const int p = threadIdx.x / LANES, j = threadIdx.x % LANES;
    const bool valid = p < rows;
    const long long r = r0 + p;
    float lq_mu_f = 0.f, lq_tau_f = 0.f, lp_mu_f = 0.f, lp_tau_f = 0.f;
    float lq_mu_b = 0.f, lq_tau_b = 0.f, lp_mu_b = 0.f, lp_tau_b = 0.f;
    if (valid) {
        // One (k,d) element per lane.  The log-densities are written with the logs shared between the terms that
        // need them:  log(1/sqrt(nu*tau)) = -(log nu + log tau)/2,  (1/sqrt(nu*tau))^2 = 1/(nu*tau),  sigma^2 = 1/tau.
        // These identities move individual terms by a few ulp relative to the reference's operation order
        // (1/sqrt then square then log again, quirk Q3) -- two orders of magnitude inside the 1e-5 budget -- and
        // cut the per-particle transcendental count from 20 to 5 per element.
        float* tab = s_tab + (size_t)p * (2 * KD * 3);
        for (int e = j; e < KD; e += LANES) {
            const float tf = __ldg(a.tau_f + r * KD + e), mf = __ldg(a.mu_f + r * KD + e);
            float tbk = tf, mbk = mf;
            if (has_bwd) { tbk = __ldg(a.tau_b + r * KD + e); mbk = __ldg(a.mu_b + r * KD + e); }
            const float Ltf = logf(tf), Ltb = has_bwd ? logf(tbk) : Ltf;
            tab[e * 3 + 0] = mf;
            tab[e * 3 + 1] = 0.5f * tf;                                   // 1 / (2 sigma^2)
            tab[e * 3 + 2] = fmaf(-0.5f, Ltf, LOG_SQRT_2PI_F);            // log sigma + c
            tab[(KD + e) * 3 + 0] = mbk;
            tab[(KD + e) * 3 + 1] = 0.5f * tbk;
            tab[(KD + e) * 3 + 2] = fmaf(-0.5f, Ltb, LOG_SQRT_2PI_F);
            if (has_eta) {
                const float a0 = __ldg(a.ra + e), b0 = __ldg(a.rb + e), m0 = __ldg(a.rm + e), n0 = __ldg(a.rn + e);
                const float qa = __ldg(a.pa + r * KD + e), qb = __ldg(a.pb + r * KD + e);
                const float qm = __ldg(a.pm + r * KD + e), qn = __ldg(a.pn + r * KD + e);
                // Gamma(c, r).log_prob(t) = c log r + (c-1) log t - r t - lgamma(c)      (gamma.py:89-98)
                const float gq = fmaf(qa, logf(qb), -lgammaf(qa)), gp = fmaf(a0, logf(b0), -lgammaf(a0));
                // Normal(m, 1/sqrt(n t)).log_prob(v) = -(v-m)^2 n t / 2 + (log n + log t)/2 - c   (normal.py:87-101)
                const float Lqn = logf(qn), Ln0 = logf(n0);
                {
                    const float dq = mf - qm, dp = mf - m0;
                    lq_tau_f += fmaf(qa - 1.f, Ltf, gq) - qb * tf;
                    lp_tau_f += fmaf(a0 - 1.f, Ltf, gp) - b0 * tf;
                    lq_mu_f += fmaf(-0.5f * dq * dq, qn * tf, 0.5f * (Lqn + Ltf)) - LOG_SQRT_2PI_F;
                    lp_mu_f += fmaf(-0.5f * dp * dp, n0 * tf, 0.5f * (Ln0 + Ltf)) - LOG_SQRT_2PI_F;
                }
                if (has_bwd) {
                    const float dq = mbk - qm, dp = mbk - m0;
                    lq_tau_b += fmaf(qa - 1.f, Ltb, gq) - qb * tbk;
                    lp_tau_b += fmaf(a0 - 1.f, Ltb, gp) - b0 * tbk;
                    lq_mu_b += fmaf(-0.5f * dq * dq, qn * tbk, 0.5f * (Lqn + Ltb)) - LOG_SQRT_2PI_F;
                    lp_mu_b += fmaf(-0.5f * dp * dp, n0 * tbk, 0.5f * (Ln0 + Ltb)) - LOG_SQRT_2PI_F;
                }
            }
        }
    }
    // log q / log p of eta: means.sum + precisions.sum (objectives.py:167-168)
    const float lq_eta_f = __fadd_rn(group_sum<LANES>(lq_mu_f), group_sum<LANES>(lq_tau_f));
    const float lp_eta_f = __fadd_rn(group_sum<LANES>(lp_mu_f), group_sum<LANES>(lp_tau_f));
    const float lq_eta_b = __fadd_rn(group_sum<LANES>(lq_mu_b), group_sum<LANES>(lq_tau_b));
    const float lp_eta_b = __fadd_rn(group_sum<LANES>(lp_mu_b), group_sum<LANES>(lp_tau_b));

    // log pi of the (renormalised, clamped) prior (models.py:167-170, quirk Q4)
    float logpi[K];
    {
        float sp = 0.f;
#pragma unroll
        for (int k = 0; k < K; ++k) sp = __fadd_rn(sp, __ldg(a.rpi + k));
#pragma unroll
        for (int k = 0; k < K; ++k)
            logpi[k] = logf(fminf(fmaxf(__fdiv_rn(__ldg(a.rpi + k), sp), FLT_EPS_F), 1.f - FLT_EPS_F));
    }

    stage_wait(&bar, 0);           // tile + tables visible

    PlainF A_llf, A_llb, A_lpzf, A_lpzb, A_lqzf, A_lqzb, A_loc;      // <= ceil(N/LANES) terms per lane
    if (valid) {
        const float* zf = s_zf + (size_t)p * N * K;
        const float* zb = s_zb + (size_t)p * N * K;
        const float* qq = s_q + (size_t)p * N * K;
        const float* tab = s_tab + (size_t)p * (2 * KD * 3);
        const float* xb = a.x + (size_t)(r % a.B) * N * D;
        for (int n = j; n < N; n += LANES) {
            // labels = argmax over K (first maximum; models.py:178)
            int lf = 0, lb = 0;
            {
                float best = zf[n * K];
#pragma unroll
                for (int k = 1; k < K; ++k) { const float v = zf[n * K + k]; if (v > best) { best = v; lf = k; } }
            }
            if (zb_distinct) {
                float best = zb[n * K];
#pragma unroll
                for (int k = 1; k < K; ++k) { const float v = zb[n * K + k]; if (v > best) { best = v; lb = k; } }
            } else {
                lb = lf;
            }
            float llf = 0.f, llb = 0.f;
#pragma unroll
            for (int d = 0; d < D; ++d) {
                const float xv = __ldg(xb + n * D + d);
                const float* tf = tab + (lf * D + d) * 3;
                const float df = __fsub_rn(xv, tf[0]);
                llf = __fadd_rn(llf, __fsub_rn(-__fmul_rn(__fmul_rn(df, df), tf[1]), tf[2]));
                if (has_bwd) {
                    const float* tbp = tab + (KD + lb * D + d) * 3;
                    const float db = __fsub_rn(xv, tbp[0]);
                    llb = __fadd_rn(llb, __fsub_rn(-__fmul_rn(__fmul_rn(db, db), tbp[1]), tbp[2]));
                }
            }
            A_llf.add(llf);
            A_llb.add(llb);
            if (has_z) {
                float sq = 0.f;
#pragma unroll
                for (int k = 0; k < K; ++k) sq = __fadd_rn(sq, qq[n * K + k]);
                const float rsq = rcp_ftz(sq);
                const float lqf = log_fast(fminf(fmaxf(qq[n * K + lf] * rsq, FLT_EPS_F), 1.f - FLT_EPS_F));
                float lpf = logpi[0];
#pragma unroll
                for (int k = 1; k < K; ++k) if (lf == k) lpf = logpi[k];
                A_lqzf.add(lqf);
                A_lpzf.add(lpf);
                float lqb = 0.f, lpb = 0.f;
                if (has_bwd) {
                    lqb = log_fast(fminf(fmaxf(qq[n * K + lb] * rsq, FLT_EPS_F), 1.f - FLT_EPS_F));
                    lpb = logpi[0];
#pragma unroll
                    for (int k = 1; k < K; ++k) if (lb == k) lpb = logpi[k];
                    A_lqzb.add(lqb);
                    A_lpzb.add(lpb);
                }
                if (step == STEP_Z) {
                    // objectives.py:206,215-216: (ll_f + log_p_f - log_q_f) - (ll_b + log_p_b - log_q_b)
                    const float wf = __fsub_rn(__fadd_rn(llf, lpf), lqf);
                    const float wb = __fsub_rn(__fadd_rn(llb, lpb), lqb);
                    const float loc = __fsub_rn(wf, wb);
                    A_loc.add(loc);
                    if (a.lw_local != nullptr) a.lw_local[r * N + n] = loc;
                    if (a.lq_local != nullptr) a.lq_local[r * N + n] = lqf;
                }
            }
        }
    }
    const double S_llf = group_sum<LANES>(A_llf.value()), S_llb = group_sum<LANES>(A_llb.value());
    const double S_lpzf = group_sum<LANES>(A_lpzf.value()), S_lpzb = group_sum<LANES>(A_lpzb.value());
    const double S_lqzf = group_sum<LANES>(A_lqzf.value()), S_lqzb = group_sum<LANES>(A_lqzb.value());
    const double S_loc = group_sum<LANES>(A_loc.value());
    if (!valid || j != 0) return;

    const float ll_f = (float)S_llf, ll_b = (float)S_llb;
    const float lpz_f = (float)S_lpzf, lpz_b = (float)S_lpzb, lqz_f = (float)S_lqzf, lqz_b = (float)S_lqzb;
    float lw, lwf = 0.f, lwb = 0.f, lqf, dens;
    if (step == STEP_ETA) {                 // objectives.py:172,179-180,194
        lwf = __fsub_rn(__fadd_rn(ll_f, lp_eta_f), lq_eta_f);
        lwb = __fsub_rn(__fadd_rn(ll_b, lp_eta_b), lq_eta_b);
        lw = __fsub_rn(lwf, lwb);
        lqf = lq_eta_f;
        dens = lp_eta_f;
    } else if (step == STEP_Z) {            // objectives.py:216-218,226
        lw = (float)S_loc;
        lwf = __fsub_rn(__fadd_rn(ll_f, lpz_f), lqz_f);
        lwb = __fsub_rn(__fadd_rn(ll_b, lpz_b), lqz_b);
        lqf = lqz_f;
        dens = __fadd_rn(ll_f, lpz_f);
    } else if (step == STEP_ONESHOT) {      // objectives.py:89-92,108
        const float lp = __fadd_rn(__fadd_rn(ll_f, lp_eta_f), lpz_f);
        lqf = __fadd_rn(lq_eta_f, lqz_f);
        lw = __fsub_rn(lp, lqf);
        lwf = lw;
        dens = lp;
    } else {                                // joint: objectives.py:128,139-140,156
        lwf = __fadd_rn(__fsub_rn(__fsub_rn(__fadd_rn(ll_f, lp_eta_f), lq_eta_f), lqz_f), lpz_f);
        lwb = __fsub_rn(__fadd_rn(__fsub_rn(__fadd_rn(ll_b, lp_eta_b), lq_eta_b), lpz_b), lqz_b);
        lw = __fsub_rn(lwf, lwb);
        lqf = __fadd_rn(lq_eta_f, lqz_f);
        dens = __fadd_rn(__fadd_rn(ll_f, lp_eta_f), lpz_f);
    }
    if (a.log_w != nullptr) a.log_w[r] = lw;
    if (a.log_q_f != nullptr) a.log_q_f[r] = lqf;
    if (a.dens != nullptr) a.dens[r] = dens;
    if (a.log_w_f != nullptr) a.log_w_f[r] = lwf;
    if (a.log_w_b != nullptr) a.log_w_b[r] = lwb;
}

// ============================================================================================
// G2 from statistics: the eta step scores the SAME assignment z under the new and the old (tau, mu)
// (objectives.py:163-180), so both log-likelihoods are functions of the one-hot statistics of z that
// k_gmm_nss_posterior<STATS> already accumulated while it read z:
//   sum_n z_nk log N(x_nd | mu, 1/sqrt(tau)) = -(S2c - 2 dl S1c + cnt dl^2) tau / 2 - cnt (c - log(tau)/2),  dl = mu - c_kd
// The statistics are centred at c (the old mu), so there is no cancellation for the backward term and only
// O(|mu_new - mu_old|) of it for the forward term.  O(K D) work and ~300 bytes per particle instead of a second
// pass over the 720-byte z row and all N points.
// ============================================================================================
struct EtaStatsArgs {
    const float *stats, *center;      // (R, K, 1+2D), (R, K, D)
    const float *pa, *pb, *pm, *pn;   // posterior (R, K, D)
    const float *ra, *rb, *rm, *rn;   // prior (K, D)
    const float *tau_f, *mu_f, *tau_b, *mu_b;
    long long R;
    float *log_w, *log_q_f, *dens, *log_w_f, *log_w_b;
    // optional fused draw: standard Gamma(alpha) and standard Normal variates -> tau, mu written out and scored
    const float *g_std, *e_std;       // (R, K, D) or null
    float *tau_out, *mu_out;          // (R, K, D), required with the draws
    // ... or drawn in the kernel: device pointer to two 64-bit words {seed, stream} the host took from torch's generator
    const unsigned long long* seed;
    // deferred gather: the OLD latents (center, tau_b, mu_b) of particle r = s B + b are read from row old_anc[r] B + b
    const int* old_anc;
    int B;
};

template <int K, int D, int LANES>
__global__ void __launch_bounds__(256) k_gmm_eta_weights_stats(EtaStatsArgs a) {
    constexpr int KD = K * D, ST = 1 + 2 * D;
    // per (k,d) constants of the prior, once per CTA: a0, b0, m0, n0, a0 log b0 - lgamma(a0), log n0
    __shared__ float s_pr[KD][6];
    if (threadIdx.x < KD) {
        const int e = threadIdx.x;
        const float a0 = __ldg(a.ra + e), b0 = __ldg(a.rb + e), n0 = __ldg(a.rn + e);
        s_pr[e][0] = a0; s_pr[e][1] = b0; s_pr[e][2] = __ldg(a.rm + e); s_pr[e][3] = n0;
        s_pr[e][4] = fmaf(a0, logf(b0), -lgammaf(a0));
        s_pr[e][5] = logf(n0);
    }
    __syncthreads();
    const long long r = (long long)blockIdx.x * (256 / LANES) + threadIdx.x / LANES;
    const int j = threadIdx.x % LANES;
    const bool valid = r < a.R;
    float lq_f = 0.f, lp_f = 0.f, lq_b = 0.f, lp_b = 0.f, ll_f = 0.f, ll_b = 0.f;
    if (valid) {
        const long long ro = a.old_anc != nullptr ? (long long)__ldg(a.old_anc + r) * a.B + (r % a.B) : r;
        for (int e = j; e < KD; e += LANES) {
            const int k = e / D, d = e % D;
            const float tbk = __ldg(a.tau_b + ro * KD + e), mbk = __ldg(a.mu_b + ro * KD + e);
            const float a0 = s_pr[e][0], b0 = s_pr[e][1], m0 = s_pr[e][2], n0 = s_pr[e][3], gp = s_pr[e][4], Ln0 = s_pr[e][5];
            const float qa = __ldg(a.pa + r * KD + e), qb = __ldg(a.pb + r * KD + e);
            const float qm = __ldg(a.pm + r * KD + e), qn = __ldg(a.pn + r * KD + e);
            float tf, mf;
            if (a.g_std != nullptr) {
                // tau = Gamma(alpha, beta).sample() = max(standard_gamma(alpha) / beta, tiny)   (torch gamma.py:73-77)
                // mu  = Normal(q_mu, 1/sqrt(q_nu tau)).sample() = e * scale + q_mu               (models.py:69-78)
                // the same IEEE operations the PyTorch elementwise kernels perform, so the draws are bit-identical
                tf = fmaxf(__fdiv_rn(__ldg(a.g_std + r * KD + e), qb), 1.17549435e-38f);
                const float sc = __fdiv_rn(1.f, __fsqrt_rn(__fmul_rn(qn, tf)));
                mf = __fadd_rn(__fmul_rn(__ldg(a.e_std + r * KD + e), sc), qm);
                a.tau_out[r * KD + e] = tf;
                a.mu_out[r * KD + e] = mf;
            } else if (a.seed != nullptr) {
                // the same draw with the variates made here (rng.cuh): standard Gamma(alpha) by Marsaglia-Tsang and a
                // standard normal by Box-Muller from Philox words counted by (particle, element) -- no _standard_gamma /
                // randn launches and no variate tensors; distribution-equal, not stream-equal, to torch's draws
                const unsigned long long sd = __ldg(a.seed), st = __ldg(a.seed + 1);
                const uint2 key = make_uint2((unsigned int)sd, (unsigned int)(sd >> 32));
                const uint4 ctr = make_uint4((unsigned int)r, (unsigned int)((unsigned long long)r >> 32) ^ (unsigned int)st,
                                             (unsigned int)e ^ (unsigned int)(st >> 32), 0u);
                const uint4 w = philox4x32_10(ctr, key);
                float x0, en;
                box_muller(w.x, w.y, x0, en);
                const float g = gamma_mt(qa, x0, w.z, w.w, ctr, key);
                // the stream is not torch's anyway: SFU quotient / reciprocal square root (~2 ulp) instead of the IEEE divide,
                // square root and divide of the explicit-variate path (their slow-path branches cost ~10 % of this kernel)
                tf = fmaxf(__fdividef(g, qb), 1.17549435e-38f);
                const float sc = rsqrtf(__fmul_rn(qn, tf));
                mf = fmaf(en, sc, qm);
                a.tau_out[r * KD + e] = tf;
                a.mu_out[r * KD + e] = mf;
            } else {
                tf = __ldg(a.tau_f + r * KD + e);
                mf = __ldg(a.mu_f + r * KD + e);
            }
            const float* sp = a.stats + r * (K * ST) + k * ST;
            const float cnt = __ldg(sp), S1 = __ldg(sp + 1 + d), S2 = __ldg(sp + 1 + D + d);
            const float cen = __ldg(a.center + ro * KD + e);
            // SFU logs (abs error ~1e-7 |log|) and the Stirling lgamma: each enters the sums with a factor <= alpha, i.e.
            // ~1e-7 relative to the magnitudes summed.  Same shared-log forms as k_gmm_weights (see the note there).
            const float Ltf = log_fast(tf), Ltb = log_fast(tbk);
            const float gq = fmaf(qa, log_fast(qb), -lgamma_pos(qa));
            const float Lqn = log_fast(qn);
            {
                const float dq = mf - qm, dp = mf - m0, dl = mf - cen;
                lq_f += (fmaf(qa - 1.f, Ltf, gq) - qb * tf) + (fmaf(-0.5f * dq * dq, qn * tf, 0.5f * (Lqn + Ltf)) - LOG_SQRT_2PI_F);
                lp_f += (fmaf(a0 - 1.f, Ltf, gp) - b0 * tf) + (fmaf(-0.5f * dp * dp, n0 * tf, 0.5f * (Ln0 + Ltf)) - LOG_SQRT_2PI_F);
                const float ss = fmaf(dl, fmaf(cnt, dl, -2.f * S1), S2);          // sum z (x - mu)^2
                ll_f += fmaf(-0.5f * tf, ss, -cnt * fmaf(-0.5f, Ltf, LOG_SQRT_2PI_F));
            }
            {
                const float dq = mbk - qm, dp = mbk - m0, dl = mbk - cen;
                lq_b += (fmaf(qa - 1.f, Ltb, gq) - qb * tbk) + (fmaf(-0.5f * dq * dq, qn * tbk, 0.5f * (Lqn + Ltb)) - LOG_SQRT_2PI_F);
                lp_b += (fmaf(a0 - 1.f, Ltb, gp) - b0 * tbk) + (fmaf(-0.5f * dp * dp, n0 * tbk, 0.5f * (Ln0 + Ltb)) - LOG_SQRT_2PI_F);
                const float ss = fmaf(dl, fmaf(cnt, dl, -2.f * S1), S2);
                ll_b += fmaf(-0.5f * tbk, ss, -cnt * fmaf(-0.5f, Ltb, LOG_SQRT_2PI_F));
            }
        }
    }
    lq_f = group_sum<LANES>(lq_f); lp_f = group_sum<LANES>(lp_f); ll_f = group_sum<LANES>(ll_f);
    lq_b = group_sum<LANES>(lq_b); lp_b = group_sum<LANES>(lp_b); ll_b = group_sum<LANES>(ll_b);
    if (!valid || j != 0) return;
    const float lwf = __fsub_rn(__fadd_rn(ll_f, lp_f), lq_f);                     // objectives.py:172,179-180,194
    const float lwb = __fsub_rn(__fadd_rn(ll_b, lp_b), lq_b);
    if (a.log_w != nullptr) a.log_w[r] = __fsub_rn(lwf, lwb);
    if (a.log_q_f != nullptr) a.log_q_f[r] = lq_f;
    if (a.dens != nullptr) a.dens[r] = lp_f;
    if (a.log_w_f != nullptr) a.log_w_f[r] = lwf;
    if (a.log_w_b != nullptr) a.log_w_b[r] = lwb;
}

#ifndef APGS_ETA_W_LANES
#define APGS_ETA_W_LANES 1
#endif
template <int K, int D>
static int launch_eta_stats(const EtaStatsArgs& a, cudaStream_t st) {
    // lanes per particle, measured at K D = 6 on B200 (ms per launch): 8 lanes 0.172 (a quarter of them idle, three shuffle
    // rounds for six sums), 4: 0.168, 2: 0.120, 1: 0.111 -- one thread walks the particle's K D elements (a warp still reads
    // 32 adjacent 24-byte rows)
    constexpr int EL = APGS_ETA_W_LANES;
    const long long grid = (a.R + 256 / EL - 1) / (256 / EL);
    k_gmm_eta_weights_stats<K, D, EL><<<(unsigned int)grid, 256, 0, st>>>(a);
    APGS_LAUNCH_OK("k_gmm_eta_weights_stats launch");
    return 0;
}

// Enc_apg_eta's head factors over (point, label) and the points, per instance, in k_gmm_nss_im's shared-memory layout: built
// once per objective (the heads and the data do not change inside it) and fetched by the CTAs with one bulk copy.
template <int K, int D>
__global__ void __launch_bounds__(256) k_gmm_eta_tables(const float* __restrict__ x, const float* Wg, const float* bg, const float* Wo,
                                                        const float* bo, int N, float* __restrict__ tabs) {
    constexpr int RS = EtaTab<K, D>::RS;
    const int NK = N * K, b = (int)blockIdx.x;
    const size_t stride = (size_t)(RS * NK + N * D + 3) / 4 * 4;
    float* out = tabs + (size_t)b * stride;
    const float* xb = x + (size_t)b * N * D;
    for (int idx = threadIdx.x; idx < NK; idx += blockDim.x) {
        const int n = idx / K, lab = idx % K;
        float xin[D], row[RS];
#pragma unroll
        for (int i = 0; i < RS; ++i) row[i] = 0.f;
#pragma unroll
        for (int d = 0; d < D; ++d) xin[d] = __ldg(xb + n * D + d);
        if (lab == 0) {
#pragma unroll
            for (int d = 0; d < D; ++d) out[(size_t)RS * NK + n * D + d] = xin[d];
        }
        eta_head_factors<K, D>(Wg, bg, Wo, bo, xin, lab, row);
#pragma unroll
        for (int i = 0; i < RS; ++i) out[(size_t)idx * RS + i] = row[i];
    }
    for (size_t i = (size_t)RS * NK + (size_t)N * D + threadIdx.x; i < stride; i += blockDim.x) out[i] = 0.f;
}

template <int K, int D>
static int launch_eta_tables(const float* x, const float* Wg, const float* bg, const float* Wo, const float* bo, int B, int N,
                             float* tabs, cudaStream_t st) {
    k_gmm_eta_tables<K, D><<<(unsigned int)B, 256, 0, st>>>(x, Wg, bg, Wo, bo, N, tabs);
    APGS_LAUNCH_OK("k_gmm_eta_tables launch");
    return 0;
}

// ----------------------------------------------------------------------------------- launchers

// lanes per particle / particles per CTA of the instance-major statistics kernel.  4 lanes x 64 particles: the per-lane
// costs that do not depend on the number of points (cross-lane reduction, posterior update, row pointers) and the
// per-CTA head table are spread over twice the points per thread compared with 8 x 32.
constexpr int kImLanes = kEtaNssLanes, kImP = 256 / kImLanes;

template <int K, int D>
static int launch_nss_im(const NssArgs& a, cudaStream_t st) {
    constexpr int NS = K * (1 + 2 * D), TE = K + 2 * K * D;
    constexpr int RW = (TE > NS ? TE : NS) | 1;
    const long long S = a.R / a.B;
    const size_t zrows = a.zl != nullptr ? ((size_t)kImP * ((a.N + 15) & ~15) + 3) / 4 : (size_t)kImP * nss_im_zstride(a.N * K);
    const size_t zreg = (max(zrows, (size_t)kImP * kImLanes * RW) + 3) / 4 * 4;
    const size_t smem = (zreg + (size_t)a.N * K * TE + (size_t)a.N * D) * 4;
    if (smem > 200 * 1024) {
        // the one-hot rows of 64 particles do not fit: 32 particles x 8 lanes
        const size_t zrows8 = a.zl != nullptr ? ((size_t)kP * ((a.N + 15) & ~15) + 3) / 4 : (size_t)kP * nss_im_zstride(a.N * K);
        const size_t zreg8 = (max(zrows8, (size_t)kP * kLanes * RW) + 3) / 4 * 4;
        const size_t smem8 = (zreg8 + (size_t)a.N * K * TE + (size_t)a.N * D) * 4;
        if (smem8 > 200 * 1024) return fail(APGS_E_SHAPE, "gmm_eta_posterior: N=%d too large for one tile", a.N);
        void (*kern8)(NssArgs, int) = a.stats != nullptr ? k_gmm_nss_im<K, D, kP, kLanes, true> : k_gmm_nss_im<K, D, kP, kLanes, false>;
        APGS_CUDA_OK(ensure_dynamic_smem(reinterpret_cast<const void*>(kern8), smem8), "smem attr");
        const long long grid8 = (long long)a.B * ((S + kP - 1) / kP);
        if (grid8 > 0x7fffffffLL) return fail(APGS_E_SHAPE, "gmm_eta_posterior: S*B too large for one launch");
        kern8<<<(unsigned int)grid8, kP * kLanes, smem8, st>>>(a, (int)S);
        APGS_LAUNCH_OK("k_gmm_nss_im launch");
        return 0;
    }
    void (*kern)(NssArgs, int) = a.stats != nullptr ? k_gmm_nss_im<K, D, kImP, kImLanes, true> : k_gmm_nss_im<K, D, kImP, kImLanes, false>;
    if (a.stats != nullptr && a.N == 60) kern = k_gmm_nss_im<K, D, kImP, kImLanes, true, 60>;      // the reference's data set
    APGS_CUDA_OK(ensure_dynamic_smem(reinterpret_cast<const void*>(kern), smem), "smem attr");
    const long long grid = (long long)a.B * ((S + kImP - 1) / kImP);
    if (grid > 0x7fffffffLL) return fail(APGS_E_SHAPE, "gmm_eta_posterior: S*B too large for one launch");
    kern<<<(unsigned int)grid, kImP * kImLanes, smem, st>>>(a, (int)S);
    APGS_LAUNCH_OK("k_gmm_nss_im launch");
    return 0;
}

template <int K, int D>
static int launch_nss(const NssArgs& a, cudaStream_t st) {
    if constexpr (K * D <= kLanes) {
        if (a.gamma == nullptr && (a.z != nullptr || a.zl != nullptr)) return launch_nss_im<K, D>(a, st);
    }
    size_t smem = (size_t)kP * a.N * (a.gamma == nullptr ? K : K + D) * 4;   // the fused variants only stage z
    if (a.stats != nullptr) smem = max(smem, (size_t)kP * kLanes * ((K * (1 + 2 * D)) | 1) * 4);
    if (smem > 200 * 1024) return fail(APGS_E_SHAPE, "gmm_nss_posterior: N=%d too large for one tile", a.N);
    void (*kern)(NssArgs) = k_gmm_nss_posterior<K, D, kP, kLanes, false, false>;
    if (a.gamma == nullptr) kern = a.z != nullptr ? (a.stats != nullptr ? k_gmm_nss_posterior<K, D, kP, kLanes, true, true, true>
                                                                        : k_gmm_nss_posterior<K, D, kP, kLanes, true, true>)
                                                  : k_gmm_nss_posterior<K, D, kP, kLanes, true, false>;
    APGS_CUDA_OK(ensure_dynamic_smem(reinterpret_cast<const void*>(kern), smem), "smem attr");
    const long long grid = (a.R + kP - 1) / kP;
    kern<<<(unsigned int)grid, kP * kLanes, smem, st>>>(a);
    APGS_LAUNCH_OK("k_gmm_nss_posterior launch");
    return 0;
}

template <int K, int D>
static int launch_weights(const GmmWArgs& a, cudaStream_t st) {
    const bool has_z = a.step != STEP_ETA, has_bwd = a.step != STEP_ONESHOT;
    const int ntile = 1 + ((has_bwd && a.z_b != a.z_f) ? 1 : 0) + (has_z ? 1 : 0);
    const size_t smem = (size_t)kP * a.N * K * 4 * ntile + (size_t)kP * 2 * K * D * 3 * 4;
    if (smem > 200 * 1024) return fail(APGS_E_SHAPE, "gmm_logw: N=%d too large for one tile", a.N);
    void (*kern)(GmmWArgs) = k_gmm_weights<K, D, kP, kLanes, STEP_ONESHOT>;
    if (a.step == STEP_ETA) kern = k_gmm_weights<K, D, kP, kLanes, STEP_ETA>;
    if (a.step == STEP_Z) kern = k_gmm_weights<K, D, kP, kLanes, STEP_Z>;
    if (a.step == STEP_JOINT) kern = k_gmm_weights<K, D, kP, kLanes, STEP_JOINT>;
    APGS_CUDA_OK(ensure_dynamic_smem(reinterpret_cast<const void*>(kern), smem), "smem attr");
    const long long grid = (a.R + kP - 1) / kP;
    kern<<<(unsigned int)grid, kP * kLanes, smem, st>>>(a);
    APGS_LAUNCH_OK("k_gmm_weights launch");
    return 0;
}

#define APGS_KD_DISPATCH(K_, D_, FN)                                              \
    do {                                                                          \
        if (K_ == 3 && D_ == 2) return FN<3, 2>(a, st);                           \
        if (K_ == 2 && D_ == 2) return FN<2, 2>(a, st);                           \
        if (K_ == 4 && D_ == 2) return FN<4, 2>(a, st);                           \
        if (K_ == 5 && D_ == 2) return FN<5, 2>(a, st);                           \
        if (K_ == 3 && D_ == 3) return FN<3, 3>(a, st);                           \
        if (K_ == 4 && D_ == 3) return FN<4, 3>(a, st);                           \
        if (K_ == 3 && D_ == 1) return FN<3, 1>(a, st);                           \
    } while (0)

}  // namespace apgs

using namespace apgs;

// which kernel families are compiled for (K, D): bit 0 the reweighting / posterior kernels of this file (G1-G4),
// bit 1 the fused z step (gmm_fused.cu), bit 2 the fused eta heads (+ statistics), bit 3 their label mode (K D <= 8 lanes)
extern "C" int apgs_gmm_supported(int64_t K, int64_t D) {
    const bool base = (K == 3 && D == 2) || (K == 2 && D == 2) || (K == 4 && D == 2) || (K == 5 && D == 2) ||
                      (K == 3 && D == 3) || (K == 4 && D == 3) || (K == 3 && D == 1);
    const bool zf = (K == 3 && D == 2) || (K == 2 && D == 2) || (K == 4 && D == 2) || (K == 5 && D == 2) || (K == 3 && D == 3);
    return (base ? 1 : 0) | (zf ? 2 : 0) | (base ? 4 : 0) | ((base && K * D <= apgs::kLanes) ? 8 : 0);
}

static int check_gmm_dims(int64_t S, int64_t B, int64_t N, int64_t K, int64_t D, const char* who) {
    if (S <= 0 || B <= 0 || N <= 0) return fail(APGS_E_SHAPE, "%s: S=%lld B=%lld N=%lld must be positive", who,
                                                (long long)S, (long long)B, (long long)N);
    if (S * B >= (1ll << 40) || N >= (1 << 20)) return fail(APGS_E_SHAPE, "%s: problem too large", who);
    (void)K; (void)D;
    return 0;
}

extern "C" int apgs_gmm_nss_posterior_f32(const float* xp, const float* gamma, int64_t S, int64_t B, int64_t N,
                                          int64_t K, int64_t D, const float* prior_alpha, const float* prior_beta,
                                          const float* prior_mu, const float* prior_nu, float* post_alpha,
                                          float* post_beta, float* post_mu, float* post_nu, void* stream) {
    if (int e = check_gmm_dims(S, B, N, K, D, "gmm_nss_posterior")) return e;
    if (!xp || !gamma || !prior_alpha || !prior_beta || !prior_mu || !prior_nu || !post_alpha || !post_beta ||
        !post_mu || !post_nu)
        return fail(APGS_E_BADARG, "gmm_nss_posterior: NULL pointer");
    NssArgs a = {};
    a.xp = xp; a.gamma = gamma;
    a.ra = prior_alpha; a.rb = prior_beta; a.rm = prior_mu; a.rn = prior_nu;
    a.oa = post_alpha; a.ob = post_beta; a.om = post_mu; a.on = post_nu;
    a.R = (long long)S * B; a.N = (int)N; a.B = (int)B;
    cudaStream_t st = (cudaStream_t)stream;
    APGS_KD_DISPATCH(K, D, launch_nss);
    return fail(APGS_E_UNSUPPORTED, "gmm_nss_posterior: (K=%lld, D=%lld) not instantiated", (long long)K, (long long)D);
}

static int gmm_logw(int step, const float* x, const float* z_f, const float* z_b, const float* q,
                    const apgs_gmm_post_t* postp, const apgs_gmm_prior_t* priorp, const float* tau_f,
                    const float* mu_f, const float* tau_b, const float* mu_b, int64_t S, int64_t B, int64_t N,
                    int64_t K, int64_t D, const apgs_gmm_out_t* out, void* stream) {
    if (!priorp || !out) return fail(APGS_E_BADARG, "gmm_logw: NULL prior/out struct");
    const float* prior[4] = {priorp->alpha, priorp->beta, priorp->mu, priorp->nu};
    const float* prior_pi = priorp->pi;
    const float* postv[4] = {postp ? postp->alpha : nullptr, postp ? postp->beta : nullptr,
                             postp ? postp->mu : nullptr, postp ? postp->nu : nullptr};
    const float* const* post = postp ? postv : nullptr;
    float *log_w = out->log_w, *log_q_f = out->log_q_f, *dens = out->density, *log_w_f = out->log_w_f,
          *log_w_b = out->log_w_b, *lw_local = out->log_w_local, *lq_local = out->log_q_local;
    if (int e = check_gmm_dims(S, B, N, K, D, "gmm_logw")) return e;
    const bool has_eta = step != STEP_Z, has_z = step != STEP_ETA, has_bwd = step != STEP_ONESHOT;
    if (!x || !z_f || !tau_f || !mu_f || !prior || !prior_pi) return fail(APGS_E_BADARG, "gmm_logw: NULL input");
    for (int i = 0; i < 4; ++i)
        if (!prior[i]) return fail(APGS_E_BADARG, "gmm_logw: NULL prior");
    if (has_eta) {
        if (!post) return fail(APGS_E_BADARG, "gmm_logw: posterior parameters required");
        for (int i = 0; i < 4; ++i)
            if (!post[i]) return fail(APGS_E_BADARG, "gmm_logw: NULL posterior parameter");
    }
    if (has_z && !q) return fail(APGS_E_BADARG, "gmm_logw: q_probs required");
    if (has_bwd && (!z_b || !tau_b || !mu_b)) return fail(APGS_E_BADARG, "gmm_logw: backward inputs required");
    GmmWArgs a = {};
    a.x = x; a.z_f = z_f; a.z_b = has_bwd ? z_b : z_f; a.q = has_z ? q : z_f;
    if (has_eta) { a.pa = post[0]; a.pb = post[1]; a.pm = post[2]; a.pn = post[3]; }
    a.ra = prior[0]; a.rb = prior[1]; a.rm = prior[2]; a.rn = prior[3]; a.rpi = prior_pi;
    a.tau_f = tau_f; a.mu_f = mu_f; a.tau_b = has_bwd ? tau_b : tau_f; a.mu_b = has_bwd ? mu_b : mu_f;
    a.R = (long long)S * B; a.B = (int)B; a.N = (int)N; a.step = step;
    a.log_w = log_w; a.log_q_f = log_q_f; a.dens = dens; a.log_w_f = log_w_f; a.log_w_b = log_w_b;
    a.lw_local = lw_local; a.lq_local = lq_local;
    cudaStream_t st = (cudaStream_t)stream;
    APGS_KD_DISPATCH(K, D, launch_weights);
    return fail(APGS_E_UNSUPPORTED, "gmm_logw: (K=%lld, D=%lld) not instantiated", (long long)K, (long long)D);
}

extern "C" int apgs_gmm_logw_oneshot_f32(const float* x, const apgs_gmm_post_t* post, const apgs_gmm_prior_t* prior,
                                         const float* tau, const float* mu, const float* q_probs, const float* z,
                                         int64_t S, int64_t B, int64_t N, int64_t K, int64_t D,
                                         const apgs_gmm_out_t* out, void* stream) {
    return gmm_logw(STEP_ONESHOT, x, z, nullptr, q_probs, post, prior, tau, mu, nullptr, nullptr, S, B, N, K, D, out,
                    stream);
}

extern "C" int apgs_gmm_logw_eta_f32(const float* x, const float* z, const apgs_gmm_post_t* post,
                                     const apgs_gmm_prior_t* prior, const float* tau, const float* mu,
                                     const float* tau_old, const float* mu_old, int64_t S, int64_t B, int64_t N,
                                     int64_t K, int64_t D, const apgs_gmm_out_t* out, void* stream) {
    return gmm_logw(STEP_ETA, x, z, z, nullptr, post, prior, tau, mu, tau_old, mu_old, S, B, N, K, D, out, stream);
}

extern "C" int apgs_gmm_logw_z_f32(const float* x, const float* tau, const float* mu, const float* q_probs,
                                   const apgs_gmm_prior_t* prior, const float* z, const float* z_old, int64_t S,
                                   int64_t B, int64_t N, int64_t K, int64_t D, const apgs_gmm_out_t* out,
                                   void* stream) {
    return gmm_logw(STEP_Z, x, z, z_old, q_probs, nullptr, prior, tau, mu, tau, mu, S, B, N, K, D, out, stream);
}

extern "C" int apgs_gmm_logw_joint_f32(const float* x, const apgs_gmm_post_t* post, const apgs_gmm_prior_t* prior,
                                       const float* tau, const float* mu, const float* tau_old, const float* mu_old,
                                       const float* q_probs, const float* z, const float* z_old, int64_t S,
                                       int64_t B, int64_t N, int64_t K, int64_t D, const apgs_gmm_out_t* out,
                                       void* stream) {
    return gmm_logw(STEP_JOINT, x, z, z_old, q_probs, post, prior, tau, mu, tau_old, mu_old, S, B, N, K, D, out,
                    stream);
}

extern "C" int apgs_gmm_eta_posterior_fused_f32(const float* x, const float* z, const float* Wg, const float* bg,
                                                const float* Wo, const float* bo, int64_t S, int64_t B, int64_t N,
                                                int64_t K, int64_t D, const float* prior_alpha, const float* prior_beta,
                                                const float* prior_mu, const float* prior_nu, float* post_alpha,
                                                float* post_beta, float* post_mu, float* post_nu, void* stream) {
    if (int e = check_gmm_dims(S, B, N, K, D, "gmm_eta_posterior_fused")) return e;
    if (!x || !Wg || !bg || !Wo || !bo || !prior_alpha || !prior_beta || !prior_mu || !prior_nu || !post_alpha ||
        !post_beta || !post_mu || !post_nu)
        return fail(APGS_E_BADARG, "gmm_eta_posterior_fused: NULL pointer");
    NssArgs a = {};
    a.x = x; a.z = z; a.Wg = Wg; a.bg = bg; a.Wo = Wo; a.bo = bo;
    a.ra = prior_alpha; a.rb = prior_beta; a.rm = prior_mu; a.rn = prior_nu;
    a.oa = post_alpha; a.ob = post_beta; a.om = post_mu; a.on = post_nu;
    a.R = (long long)S * B; a.N = (int)N; a.B = (int)B;
    cudaStream_t st = (cudaStream_t)stream;
    APGS_KD_DISPATCH(K, D, launch_nss);
    return fail(APGS_E_UNSUPPORTED, "gmm_eta_posterior_fused: (K=%lld, D=%lld) not instantiated", (long long)K,
                (long long)D);
}

extern "C" int apgs_gmm_eta_posterior_stats_f32(const float* x, const float* z, const float* Wg, const float* bg,
                                                const float* Wo, const float* bo, int64_t S, int64_t B, int64_t N,
                                                int64_t K, int64_t D, const float* prior_alpha, const float* prior_beta,
                                                const float* prior_mu, const float* prior_nu, float* post_alpha,
                                                float* post_beta, float* post_mu, float* post_nu, const float* center,
                                                float* stats, void* stream) {
    if (int e = check_gmm_dims(S, B, N, K, D, "gmm_eta_posterior_stats")) return e;
    if (!x || !z || !Wg || !bg || !Wo || !bo || !prior_alpha || !prior_beta || !prior_mu || !prior_nu || !post_alpha ||
        !post_beta || !post_mu || !post_nu || !center || !stats)
        return fail(APGS_E_BADARG, "gmm_eta_posterior_stats: NULL pointer");
    NssArgs a = {};
    a.x = x; a.z = z; a.Wg = Wg; a.bg = bg; a.Wo = Wo; a.bo = bo;
    a.ra = prior_alpha; a.rb = prior_beta; a.rm = prior_mu; a.rn = prior_nu;
    a.oa = post_alpha; a.ob = post_beta; a.om = post_mu; a.on = post_nu;
    a.center = center; a.stats = stats;
    a.R = (long long)S * B; a.N = (int)N; a.B = (int)B;
    cudaStream_t st = (cudaStream_t)stream;
    APGS_KD_DISPATCH(K, D, launch_nss);
    return fail(APGS_E_UNSUPPORTED, "gmm_eta_posterior_stats: (K=%lld, D=%lld) not instantiated", (long long)K,
                (long long)D);
}

static int logw_eta_stats(const float* stats, const float* center, const apgs_gmm_post_t* post,
                          const apgs_gmm_prior_t* prior, const float* tau, const float* mu, const float* g_std,
                          const float* e_std, float* tau_out, float* mu_out, const float* tau_old, const float* mu_old,
                          int64_t S, int64_t B, int64_t K, int64_t D, const apgs_gmm_out_t* out, void* stream,
                          const uint64_t* seed = nullptr, const int32_t* old_anc = nullptr);

extern "C" int apgs_gmm_logw_eta_stats_f32(const float* stats, const float* center, const apgs_gmm_post_t* post,
                                           const apgs_gmm_prior_t* prior, const float* tau, const float* mu,
                                           const float* tau_old, const float* mu_old, int64_t S, int64_t B, int64_t K,
                                           int64_t D, const apgs_gmm_out_t* out, void* stream) {
    if (!tau || !mu) return fail(APGS_E_BADARG, "gmm_logw_eta_stats: NULL input");
    return logw_eta_stats(stats, center, post, prior, tau, mu, nullptr, nullptr, nullptr, nullptr, tau_old, mu_old, S, B,
                          K, D, out, stream);
}

extern "C" int apgs_gmm_draw_logw_eta_stats_f32(const float* stats, const float* center, const apgs_gmm_post_t* post,
                                                const apgs_gmm_prior_t* prior, const float* gamma_std,
                                                const float* normal_std, float* tau, float* mu, const float* tau_old,
                                                const float* mu_old, int64_t S, int64_t B, int64_t K, int64_t D,
                                                const apgs_gmm_out_t* out, void* stream) {
    if (!gamma_std || !normal_std || !tau || !mu) return fail(APGS_E_BADARG, "gmm_draw_logw_eta_stats: NULL input");
    return logw_eta_stats(stats, center, post, prior, nullptr, nullptr, gamma_std, normal_std, tau, mu, tau_old, mu_old, S,
                          B, K, D, out, stream);
}

extern "C" int apgs_gmm_draw_logw_eta_stats_rng_f32(const float* stats, const float* center, const apgs_gmm_post_t* post,
                                                    const apgs_gmm_prior_t* prior, const uint64_t* philox_seed, float* tau,
                                                    float* mu, const float* tau_old, const float* mu_old, int64_t S,
                                                    int64_t B, int64_t K, int64_t D, const apgs_gmm_out_t* out, void* stream) {
    if (!philox_seed || !tau || !mu) return fail(APGS_E_BADARG, "gmm_draw_logw_eta_stats_rng: NULL input");
    return logw_eta_stats(stats, center, post, prior, nullptr, nullptr, nullptr, nullptr, tau, mu, tau_old, mu_old, S, B, K, D,
                          out, stream, philox_seed);
}

extern "C" int apgs_gmm_draw_logw_eta_stats_rng_via_f32(const float* stats, const float* center, const apgs_gmm_post_t* post,
                                                        const apgs_gmm_prior_t* prior, const uint64_t* philox_seed, float* tau,
                                                        float* mu, const float* tau_old, const float* mu_old,
                                                        const int32_t* old_ancestors, int64_t S, int64_t B, int64_t K,
                                                        int64_t D, const apgs_gmm_out_t* out, void* stream) {
    if (!philox_seed || !tau || !mu) return fail(APGS_E_BADARG, "gmm_draw_logw_eta_stats_rng_via: NULL input");
    if (B >= (1ll << 31)) return fail(APGS_E_SHAPE, "gmm_draw_logw_eta_stats_rng_via: B too large");
    return logw_eta_stats(stats, center, post, prior, nullptr, nullptr, nullptr, nullptr, tau, mu, tau_old, mu_old, S, B, K, D,
                          out, stream, philox_seed, old_ancestors);
}

static int logw_eta_stats(const float* stats, const float* center, const apgs_gmm_post_t* post,
                          const apgs_gmm_prior_t* prior, const float* tau, const float* mu, const float* g_std,
                          const float* e_std, float* tau_out, float* mu_out, const float* tau_old, const float* mu_old,
                          int64_t S, int64_t B, int64_t K, int64_t D, const apgs_gmm_out_t* out, void* stream,
                          const uint64_t* seed, const int32_t* old_anc) {
    if (S <= 0 || B <= 0 || S * B >= (1ll << 36))
        return fail(APGS_E_SHAPE, "gmm_logw_eta_stats: bad sizes S=%lld B=%lld", (long long)S, (long long)B);
    if (!stats || !center || !post || !prior || !out || !tau_old || !mu_old)
        return fail(APGS_E_BADARG, "gmm_logw_eta_stats: NULL input");
    if (!post->alpha || !post->beta || !post->mu || !post->nu || !prior->alpha || !prior->beta || !prior->mu || !prior->nu)
        return fail(APGS_E_BADARG, "gmm_logw_eta_stats: NULL prior / posterior parameter");
    EtaStatsArgs a = {};
    a.stats = stats; a.center = center;
    a.pa = post->alpha; a.pb = post->beta; a.pm = post->mu; a.pn = post->nu;
    a.ra = prior->alpha; a.rb = prior->beta; a.rm = prior->mu; a.rn = prior->nu;
    a.tau_f = tau; a.mu_f = mu; a.tau_b = tau_old; a.mu_b = mu_old;
    a.g_std = g_std; a.e_std = e_std; a.tau_out = tau_out; a.mu_out = mu_out;
    a.seed = reinterpret_cast<const unsigned long long*>(seed);
    a.old_anc = old_anc; a.B = (int)B;
    a.R = (long long)S * B;
    a.log_w = out->log_w; a.log_q_f = out->log_q_f; a.dens = out->density; a.log_w_f = out->log_w_f;
    a.log_w_b = out->log_w_b;
    cudaStream_t st = (cudaStream_t)stream;
    APGS_KD_DISPATCH(K, D, launch_eta_stats);
    return fail(APGS_E_UNSUPPORTED, "gmm_logw_eta_stats: (K=%lld, D=%lld) not instantiated", (long long)K, (long long)D);
}

extern "C" int apgs_gmm_eta_posterior_stats_labels_f32(const float* x, const uint8_t* z_labels, const float* Wg,
                                                       const float* bg, const float* Wo, const float* bo, int64_t S,
                                                       int64_t B, int64_t N, int64_t K, int64_t D,
                                                       const float* prior_alpha, const float* prior_beta,
                                                       const float* prior_mu, const float* prior_nu, float* post_alpha,
                                                       float* post_beta, float* post_mu, float* post_nu,
                                                       const float* center, float* stats, int64_t label_stride,
                                                       void* stream) {
    if (int e = check_gmm_dims(S, B, N, K, D, "gmm_eta_posterior_stats_labels")) return e;
    if (label_stride < N) return fail(APGS_E_BADARG, "gmm_eta_posterior_stats_labels: label_stride %lld < N", (long long)label_stride);
    if (!x || !z_labels || !Wg || !bg || !Wo || !bo || !prior_alpha || !prior_beta || !prior_mu || !prior_nu ||
        !post_alpha || !post_beta || !post_mu || !post_nu)
        return fail(APGS_E_BADARG, "gmm_eta_posterior_stats_labels: NULL pointer");
    if ((center == nullptr) != (stats == nullptr))
        return fail(APGS_E_BADARG, "gmm_eta_posterior_stats_labels: center and stats go together");
    if (K * D > kLanes) return fail(APGS_E_UNSUPPORTED, "gmm_eta_posterior_stats_labels: K*D <= %d supported", kLanes);
    NssArgs a = {};
    a.x = x; a.zl = z_labels; a.LN = label_stride; a.Wg = Wg; a.bg = bg; a.Wo = Wo; a.bo = bo;
    a.ra = prior_alpha; a.rb = prior_beta; a.rm = prior_mu; a.rn = prior_nu;
    a.oa = post_alpha; a.ob = post_beta; a.om = post_mu; a.on = post_nu;
    a.center = center; a.stats = stats;
    a.R = (long long)S * B; a.N = (int)N; a.B = (int)B;
    cudaStream_t st = (cudaStream_t)stream;
    APGS_KD_DISPATCH(K, D, launch_nss);
    return fail(APGS_E_UNSUPPORTED, "gmm_eta_posterior_stats_labels: (K=%lld, D=%lld) not instantiated", (long long)K,
                (long long)D);
}

extern "C" int apgs_gmm_eta_posterior_stats_labels_via_f32(const float* x, const uint8_t* z_labels, const float* Wg,
                                                       const float* bg, const float* Wo, const float* bo, int64_t S,
                                                       int64_t B, int64_t N, int64_t K, int64_t D,
                                                       const float* prior_alpha, const float* prior_beta,
                                                       const float* prior_mu, const float* prior_nu, float* post_alpha,
                                                       float* post_beta, float* post_mu, float* post_nu,
                                                       const float* center, float* stats, int64_t label_stride,
                                                       const int32_t* z_ancestors, const int32_t* center_ancestors,
                                                       const float* tables, void* stream) {
    if (int e = check_gmm_dims(S, B, N, K, D, "gmm_eta_posterior_stats_labels_via")) return e;
    if (tables && !aligned(tables, 16)) return fail(APGS_E_BADARG, "gmm_eta_posterior_stats_labels_via: tables must be 16-byte aligned");
    if (label_stride < N) return fail(APGS_E_BADARG, "gmm_eta_posterior_stats_labels_via: label_stride %lld < N", (long long)label_stride);
    if (!x || !z_labels || !Wg || !bg || !Wo || !bo || !prior_alpha || !prior_beta || !prior_mu || !prior_nu ||
        !post_alpha || !post_beta || !post_mu || !post_nu)
        return fail(APGS_E_BADARG, "gmm_eta_posterior_stats_labels_via: NULL pointer");
    if ((center == nullptr) != (stats == nullptr))
        return fail(APGS_E_BADARG, "gmm_eta_posterior_stats_labels_via: center and stats go together");
    if (K * D > kLanes) return fail(APGS_E_UNSUPPORTED, "gmm_eta_posterior_stats_labels_via: K*D <= %d supported", kLanes);
    NssArgs a = {};
    a.x = x; a.zl = z_labels; a.LN = label_stride; a.Wg = Wg; a.bg = bg; a.Wo = Wo; a.bo = bo;
    a.ra = prior_alpha; a.rb = prior_beta; a.rm = prior_mu; a.rn = prior_nu;
    a.oa = post_alpha; a.ob = post_beta; a.om = post_mu; a.on = post_nu;
    a.center = center; a.stats = stats;
    a.z_anc = z_ancestors; a.c_anc = center_ancestors; a.tabs = tables;
    a.R = (long long)S * B; a.N = (int)N; a.B = (int)B;
    cudaStream_t st = (cudaStream_t)stream;
    APGS_KD_DISPATCH(K, D, launch_nss);
    return fail(APGS_E_UNSUPPORTED, "gmm_eta_posterior_stats_labels_via: (K=%lld, D=%lld) not instantiated", (long long)K,
                (long long)D);
}

template <int K, int D>
static size_t eta_tables_floats(int64_t N) { return (size_t)(EtaTab<K, D>::RS * N * K + N * D + 3) / 4 * 4; }

// (K, D) of the instance-major statistics kernel (K D <= 8 and compiled): floats per instance, 0 if unsupported
static size_t eta_tables_floats_kd(int64_t K, int64_t D, int64_t N) {
    if (K == 3 && D == 2) return eta_tables_floats<3, 2>(N);
    if (K == 2 && D == 2) return eta_tables_floats<2, 2>(N);
    if (K == 4 && D == 2) return eta_tables_floats<4, 2>(N);
    if (K == 3 && D == 1) return eta_tables_floats<3, 1>(N);
    return 0;
}

extern "C" size_t apgs_gmm_eta_tables_bytes(int64_t B, int64_t N, int64_t K, int64_t D) {
    if (B <= 0 || N <= 0) return 0;
    return (size_t)B * eta_tables_floats_kd(K, D, N) * 4;
}

extern "C" int apgs_gmm_eta_tables_f32(const float* x, const float* Wg, const float* bg, const float* Wo, const float* bo,
                                       int64_t B, int64_t N, int64_t K, int64_t D, float* tables, void* stream) {
    if (B <= 0 || N <= 0 || B > 0x7fffffffLL || N > (1 << 20)) return fail(APGS_E_SHAPE, "gmm_eta_tables: bad sizes");
    if (!x || !Wg || !bg || !Wo || !bo || !tables) return fail(APGS_E_BADARG, "gmm_eta_tables: NULL pointer");
    if (!aligned(tables, 16)) return fail(APGS_E_BADARG, "gmm_eta_tables: tables must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    if (K == 3 && D == 2) return launch_eta_tables<3, 2>(x, Wg, bg, Wo, bo, (int)B, (int)N, tables, st);
    if (K == 2 && D == 2) return launch_eta_tables<2, 2>(x, Wg, bg, Wo, bo, (int)B, (int)N, tables, st);
    if (K == 4 && D == 2) return launch_eta_tables<4, 2>(x, Wg, bg, Wo, bo, (int)B, (int)N, tables, st);
    if (K == 3 && D == 1) return launch_eta_tables<3, 1>(x, Wg, bg, Wo, bo, (int)B, (int)N, tables, st);
    return fail(APGS_E_UNSUPPORTED, "gmm_eta_tables: (K=%lld, D=%lld) not instantiated", (long long)K, (long long)D);
}
