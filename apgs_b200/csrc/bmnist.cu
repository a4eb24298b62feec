// Bouncing-MNIST kernels for sm_100a.
//
//   k_stn_crop      B1  Affine_Transformer.frame_to_digit  (apgs/bmnist/affine_transformer.py:54-65)
//   k_stn_paste     B2  Affine_Transformer.digit_to_frame  (:41-52)
//   k_stn_paste_ll  B2+B3  Dec_digit.forward tail (apgs/bmnist/models.py:148-154) + MBern_log_prob (:156-162):
//                       recon = clamp(sum_k paste_k, 0, 1), ll = sum_px log(recon+eps) x + log(1-recon+eps)(1-x)
//   k_logw_where / k_logw_what  B4  Normal log-densities of z_where / z_what under proposal and prior and the
//                       step weights (apgs/bmnist/objectives.py:198-207,241; models.py:97-104)
//
// The reference's transformer is affine_grid + grid_sample(mode='nearest', align_corners=True) with a DIAGONAL
// theta, so the sampling grid is separable: per (s,b,t,k) there is one source column per output column and one
// source row per output row,
//     src[j] = rint(((lin[j] * scale + t) + 1) * (n_in - 1)/2)      (fp32, one rounding per operation)
// lin = affine_grid's base axis (torch.linspace(-1,1,n_out), supplied by the host so it is bit-identical to the
// reference's).  A CTA owns one (s,b,t): it builds the 2*K index vectors in shared memory once and then streams
// pixels with coalesced accesses; the (S*B*T*K, H, W, 2) grid tensor and the T-fold / K-fold repeats of the
// reference never exist, and `frames` may have a stride-0 particle axis (read once per instance).
#include "tile.cuh"

namespace apgs {

__device__ __forceinline__ int stn_index(float lin, float scale, float t, int n_in) {
    const float v = __fadd_rn(__fmul_rn(lin, scale), t);
    const float u = __fmul_rn(__fadd_rn(v, 1.f), (float)(n_in - 1) * 0.5f);
    const float r = rintf(u);
    return (r >= 0.f && r <= (float)(n_in - 1)) ? (int)r : -1;
}

struct StnArgs {
    const float* src;          // crop: frames;            paste: digits (S,B,K,DP,DP)
    long long ss, sb, st;      // crop: element strides of frames over (s, b, t)
    const float* z_where;      // (S,B,T,K,2)
    const float* lin;          // base axis of the OUTPUT resolution
    float* dst;                // crop: (S,B,T,K,DP,DP);  paste: (S,B,T,K,FP,FP) or recon (S,B,T,FP,FP)
    const float* frames;       // paste_ll: observed frames (strides fs, fb, ft)
    long long fs, fb, ft;
    float* ll;                 // paste_ll: (S,B,T)
    int S, B, T, K, FP, DP;
};

// ---------------------------------------------------------------------------------------------- B1
__global__ void __launch_bounds__(256) k_stn_crop(StnArgs a) {
    extern __shared__ int s_idx[];                 // [K][2][DP]: row index, col index
    const int sbt = blockIdx.x;
    const int t = sbt % a.T, b = (sbt / a.T) % a.B, s = sbt / (a.T * a.B);
    const int K = a.K, DP = a.DP, FP = a.FP;
    const float scale = (float)((double)DP / (double)FP);          // FloatTensor([[DP/FP, 0], [0, DP/FP]]) (:37)
    const float tr = (float)((double)(FP - DP) / (double)FP);      // python float -> fp32 scalar multiply (:61)
    for (int i = threadIdx.x; i < K * 2 * DP; i += blockDim.x) {
        const int k = i / (2 * DP), ax = (i / DP) & 1, j = i % DP;
        const float* zw = a.z_where + ((long long)sbt * K + k) * 2;
        // ax 0: rows use -z_y * tr (the y axis is flipped, :62);  ax 1: columns use +z_x * tr
        const float tt = ax == 0 ? -__fmul_rn(__ldg(zw + 1), tr) : __fmul_rn(__ldg(zw), tr);
        s_idx[i] = stn_index(__ldg(a.lin + j), scale, tt, FP);
    }
    __syncthreads();
    const float* fr = a.src + s * a.ss + b * a.sb + t * a.st;
    float* out = a.dst + (long long)sbt * K * DP * DP;
    for (int o = threadIdx.x; o < K * DP * DP; o += blockDim.x) {
        const int k = o / (DP * DP), i = (o / DP) % DP, j = o % DP;
        const int r = s_idx[(k * 2 + 0) * DP + i], c = s_idx[(k * 2 + 1) * DP + j];
        out[o] = (r >= 0 && c >= 0) ? __ldg(fr + r * FP + c) : 0.f;
    }
}

// ---------------------------------------------------------------------------------- B2 and B2+B3
template <bool WITH_LL>
__global__ void __launch_bounds__(256) k_stn_paste(StnArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ unsigned long long bar;
    __shared__ double s_red[8];
    const int sbt = blockIdx.x;
    const int t = sbt % a.T, b = (sbt / a.T) % a.B, s = sbt / (a.T * a.B);
    const int K = a.K, DP = a.DP, FP = a.FP;
    float* s_dig = reinterpret_cast<float*>(smem);                    // K*DP*DP
    int* s_idx = reinterpret_cast<int*>(s_dig + K * DP * DP);         // [K][2][FP]
    if (threadIdx.x == 0) mbar_init(&bar, 1);
    __syncthreads();
    // digits of (s,b): dense (S,B,K,DP,DP), or -- residual mode, K = 1 -- one DP x DP image at explicit (s,b) strides
    const float* dig = (a.ss | a.sb) ? a.src + s * a.ss + b * a.sb : a.src + ((long long)s * a.B + b) * K * DP * DP;
    StageReq req[1] = {{s_dig, dig, (unsigned int)(K * DP * DP * 4)}};
    stage_issue(req, 1, &bar);
    const float scale = (float)((double)FP / (double)DP);             // FloatTensor([[FP/DP, 0], [0, FP/DP]]) (:36)
    const float tr = (float)((double)(FP - DP) / (double)DP);         // (:48)
    for (int i = threadIdx.x; i < K * 2 * FP; i += blockDim.x) {
        const int k = i / (2 * FP), ax = (i / FP) & 1, j = i % FP;
        const float* zw = a.z_where + ((long long)sbt * K + k) * 2;
        // ax 0: rows use +z_y * tr;  ax 1: columns use -z_x * tr (the x axis is flipped, :49)
        const float tt = ax == 0 ? __fmul_rn(__ldg(zw + 1), tr) : -__fmul_rn(__ldg(zw), tr);
        s_idx[i] = stn_index(__ldg(a.lin + j), scale, tt, DP);
    }
    stage_wait(&bar, 0);

    if (!WITH_LL) {
        float* out = a.dst + (long long)sbt * K * FP * FP;
        // residual mode (frames given, K = 1): out = frame - paste, the frame_left update of objectives.py:87-89
        const float* res = a.frames != nullptr ? a.frames + s * a.fs + b * a.fb + t * a.ft : nullptr;
        for (int o = threadIdx.x; o < K * FP * FP; o += blockDim.x) {
            const int k = o / (FP * FP), i = (o / FP) % FP, j = o % FP;
            const int r = s_idx[(k * 2 + 0) * FP + i], c = s_idx[(k * 2 + 1) * FP + j];
            const float v = (r >= 0 && c >= 0) ? s_dig[(k * DP + r) * DP + c] : 0.f;
            __stcs(out + o, res != nullptr ? __fsub_rn(__ldg(res + o), v) : v);
        }
        return;
    }
    const float* fr = a.frames + s * a.fs + b * a.fb + t * a.ft;
    float* rec = a.dst != nullptr ? a.dst + (long long)sbt * FP * FP : nullptr;
    const float EPS = 1e-9f;
    float acc = 0.f, comp = 0.f;                                      // Kahan partial of this thread
    // Background pixels (no digit pasted: recon == 0, ~3/4 of a frame) take their two logarithms from registers: the same
    // logf of the same arguments, evaluated once per thread instead of once per pixel -- bit-identical.
    float zero;                                                       // opaque to the compiler: the run-time logf, not a folded constant
    asm volatile("mov.f32 %0, 0f00000000;" : "=f"(zero));
    const float log_eps = logf(__fadd_rn(zero, EPS)), log_one = logf(__fadd_rn(__fsub_rn(1.f, zero), EPS));
    const int step_i = (int)blockDim.x / FP, step_j = (int)blockDim.x % FP;
    int i = (int)threadIdx.x / FP, j = (int)threadIdx.x % FP;          // (row, column) of pixel o, advanced without divisions
    for (int o = threadIdx.x; o < FP * FP; o += blockDim.x) {
        float v = 0.f;
        for (int k = 0; k < K; ++k) {                                 // .sum(-3): k = 0, 1, 2 in order
            const int r = s_idx[(k * 2 + 0) * FP + i], c = s_idx[(k * 2 + 1) * FP + j];
            v = __fadd_rn(v, (r >= 0 && c >= 0) ? s_dig[(k * DP + r) * DP + c] : 0.f);
        }
        v = fminf(fmaxf(v, 0.f), 1.f);                                // torch.clamp(., 0, 1)
        if (rec != nullptr) __stcs(rec + o, v);
        const float x = __ldg(fr + o);
        // log(x_mean + EPS) * x + log(1 - x_mean + EPS) * (1 - x)      (models.py:161-162)
        float l0 = log_eps, l1 = log_one;
        if (v != 0.f) {
            l0 = logf(__fadd_rn(v, EPS));
            l1 = logf(__fadd_rn(__fsub_rn(1.f, v), EPS));
        }
        const float term = __fadd_rn(__fmul_rn(l0, x), __fmul_rn(l1, __fsub_rn(1.f, x)));
        i += step_i;
        j += step_j;
        if (j >= FP) { j -= FP; ++i; }
        const float y = __fsub_rn(term, comp);
        const float tsum = __fadd_rn(acc, y);
        comp = __fsub_rn(__fsub_rn(tsum, acc), y);
        acc = tsum;
    }
    double d = warp_sum((double)acc - (double)comp);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = d;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += s_red[w];
        a.ll[sbt] = (float)tot;
    }
}

// ------------------------------------------------------------------------------- bilinear mode (north_star item 4)
// The reference's transformer samples with mode='nearest' (that is the parity mode, above); the bilinear
// variant is the same separable transform with grid_sample(mode='bilinear', padding_mode='zeros', align_corners=True)
// semantics: per output row / column a lower source index and a fraction, out = sum of the four in-bounds neighbours
// weighted (1-fr)(1-fc), (1-fr) fc, fr (1-fc), fr fc.  CTA per (s,b,t); index / fraction vectors in shared memory.
__device__ __forceinline__ void stn_axis_bilinear(float lin, float scale, float t, int n_in, int& i0, float& f) {
    const float v = __fadd_rn(__fmul_rn(lin, scale), t);
    const float u = __fmul_rn(__fadd_rn(v, 1.f), (float)(n_in - 1) * 0.5f);
    const float fl = floorf(u);
    // far out of range: park the index where both neighbours are out of bounds
    i0 = (fl >= -2.f && fl <= (float)n_in) ? (int)fl : -2;
    f = u - fl;
}

template <bool PASTE>
__global__ void __launch_bounds__(256) k_stn_bilinear(StnArgs a) {
    extern __shared__ __align__(16) unsigned char smem_b[];
    const int sbt = blockIdx.x;
    const int t = sbt % a.T, b = (sbt / a.T) % a.B, s = sbt / (a.T * a.B);
    const int K = a.K, DP = a.DP, FP = a.FP;
    const int NO = PASTE ? FP : DP, NI = PASTE ? DP : FP;            // output / input resolution
    int* s_i0 = reinterpret_cast<int*>(smem_b);                      // [K][2][NO]
    float* s_f = reinterpret_cast<float*>(s_i0 + K * 2 * NO);        // [K][2][NO]
    const float scale = (float)((double)NO / (double)NI);
    const float tr = PASTE ? (float)((double)(FP - DP) / (double)DP) : (float)((double)(FP - DP) / (double)FP);
    for (int i = threadIdx.x; i < K * 2 * NO; i += blockDim.x) {
        const int k = i / (2 * NO), ax = (i / NO) & 1, j = i % NO;
        const float* zw = a.z_where + ((long long)sbt * K + k) * 2;
        float tt;
        if (PASTE) tt = ax == 0 ? __fmul_rn(__ldg(zw + 1), tr) : -__fmul_rn(__ldg(zw), tr);      // affine_transformer.py:47-49
        else tt = ax == 0 ? -__fmul_rn(__ldg(zw + 1), tr) : __fmul_rn(__ldg(zw), tr);            // :60-62
        stn_axis_bilinear(__ldg(a.lin + j), scale, tt, NI, s_i0[i], s_f[i]);
    }
    __syncthreads();
    float* out = a.dst + (long long)sbt * K * NO * NO;
    for (int o = threadIdx.x; o < K * NO * NO; o += blockDim.x) {
        const int k = o / (NO * NO), i = (o / NO) % NO, j = o % NO;
        const float* img = PASTE ? a.src + (((long long)s * a.B + b) * K + k) * DP * DP
                                 : a.src + s * a.ss + b * a.sb + t * a.st;
        const int r0 = s_i0[(k * 2 + 0) * NO + i], c0 = s_i0[(k * 2 + 1) * NO + j];
        const float fr = s_f[(k * 2 + 0) * NO + i], fc = s_f[(k * 2 + 1) * NO + j];
        float v = 0.f;
#pragma unroll
        for (int dr = 0; dr < 2; ++dr)
#pragma unroll
            for (int dc = 0; dc < 2; ++dc) {
                const int r = r0 + dr, c = c0 + dc;
                if (r >= 0 && r < NI && c >= 0 && c < NI)
                    v = fmaf(__ldg(img + r * NI + c), (dr ? fr : 1.f - fr) * (dc ? fc : 1.f - fc), v);
            }
        out[o] = v;
    }
}

// ---------------------------------------------------------------------------------------------- B4
#define LOG_SQRT_2PI_F 0.918938533204672741780329736406f
__device__ __forceinline__ float normal_lp(float v, float loc, float scale) {      // torch normal.py:87-101
    const float var = __fmul_rn(scale, scale);
    const float d = __fsub_rn(v, loc);
    float t = __fdiv_rn(-__fmul_rn(d, d), __fmul_rn(2.f, var));
    t = __fsub_rn(t, logf(scale));
    return __fsub_rn(t, LOG_SQRT_2PI_F);
}

struct WhereArgs {
    const float *q_mean, *q_std;         // (R, K, D)
    const float *z_new, *z_prev;         // (R, K, D); z_prev null at t = 0
    const float *z_old, *z_old_prev;     // (R, K, D) or null (one-shot: no backward)
    const float *mu0, *sig0, *sigt;      // (D)
    const float *ll_f, *ll_b;            // (R) or null
    long long R;
    int K, D;
    float *log_w, *log_q_f, *log_p_f, *log_q_b, *log_p_b;    // (R), optional
};

__global__ void __launch_bounds__(256) k_logw_where(WhereArgs a) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= a.R) return;
    const int K = a.K, D = a.D;
    float lqf = 0.f, lpf = 0.f, lqb = 0.f, lpb = 0.f;
    for (int k = 0; k < K; ++k) {
        float qf = 0.f, pf = 0.f, qb = 0.f, pb = 0.f;                 // .sum(-1) over D, then summed over k
        for (int d = 0; d < D; ++d) {
            const long long o = (r * K + k) * D + d;
            const float m = __ldg(a.q_mean + o), sd = __ldg(a.q_std + o), zn = __ldg(a.z_new + o);
            qf = __fadd_rn(qf, normal_lp(zn, m, sd));
            pf = __fadd_rn(pf, a.z_prev ? normal_lp(zn, __ldg(a.z_prev + o), __ldg(a.sigt + d))
                                        : normal_lp(zn, __ldg(a.mu0 + d), __ldg(a.sig0 + d)));
            if (a.z_old) {
                const float zo = __ldg(a.z_old + o);
                qb = __fadd_rn(qb, normal_lp(zo, m, sd));
                pb = __fadd_rn(pb, a.z_old_prev ? normal_lp(zo, __ldg(a.z_old_prev + o), __ldg(a.sigt + d))
                                                : normal_lp(zo, __ldg(a.mu0 + d), __ldg(a.sig0 + d)));
            }
        }
        lqf = __fadd_rn(lqf, qf); lpf = __fadd_rn(lpf, pf); lqb = __fadd_rn(lqb, qb); lpb = __fadd_rn(lpb, pb);
    }
    if (a.log_q_f) a.log_q_f[r] = lqf;
    if (a.log_p_f) a.log_p_f[r] = lpf;
    if (a.log_q_b) a.log_q_b[r] = lqb;
    if (a.log_p_b) a.log_p_b[r] = lpb;
    if (a.log_w) {
        // objectives.py:198-199,206: (log_p_f - log_q_f) - (log_p_b - log_q_b) + ll_f - ll_b
        float w = __fsub_rn(__fsub_rn(lpf, lqf), __fsub_rn(lpb, lqb));
        if (a.ll_f) w = __fadd_rn(w, __ldg(a.ll_f + r));
        if (a.ll_b) w = __fsub_rn(w, __ldg(a.ll_b + r));
        a.log_w[r] = w;
    }
}

struct WhatArgs {
    const float *q_mu, *q_std;           // (R, K, Z)
    const float *z_new, *z_old;          // (R, K, Z); z_old null for the one-shot
    const float *ll_f, *ll_b;            // (R, T)
    const float* extra_p;                // (R) added to log p (one-shot: log_p_where), or null
    const float* extra_q;                // (R) added to log q (one-shot: log_q_where), or null
    long long R;
    int K, Z, T;
    float *log_w, *log_q_f, *dens, *ll_sum;   // (R), optional
};

__global__ void __launch_bounds__(256) k_logw_what(WhatArgs a) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= a.R) return;
    float lqf = 0.f, lpf = 0.f, lqb = 0.f, lpb = 0.f;
    for (int k = 0; k < a.K; ++k) {
        float qf = 0.f, pf = 0.f, qb = 0.f, pb = 0.f;
        for (int z = 0; z < a.Z; ++z) {
            const long long o = (r * a.K + k) * a.Z + z;
            const float m = __ldg(a.q_mu + o), sd = __ldg(a.q_std + o), zn = __ldg(a.z_new + o);
            qf = __fadd_rn(qf, normal_lp(zn, m, sd));
            pf = __fadd_rn(pf, normal_lp(zn, 0.f, 1.f));              // Dec_digit prior N(0,1) (models.py:129-130)
            if (a.z_old) {
                const float zo = __ldg(a.z_old + o);
                qb = __fadd_rn(qb, normal_lp(zo, m, sd));
                pb = __fadd_rn(pb, normal_lp(zo, 0.f, 1.f));
            }
        }
        lqf = __fadd_rn(lqf, qf); lpf = __fadd_rn(lpf, pf); lqb = __fadd_rn(lqb, qb); lpb = __fadd_rn(lpb, pb);
    }
    float llf = 0.f, llb = 0.f;
    for (int t = 0; t < a.T; ++t) {
        llf = __fadd_rn(llf, __ldg(a.ll_f + r * a.T + t));
        if (a.ll_b) llb = __fadd_rn(llb, __ldg(a.ll_b + r * a.T + t));
    }
    float lp = __fadd_rn(llf, lpf), lq = lqf;
    if (a.extra_p) lp = __fadd_rn(__fadd_rn(__ldg(a.extra_p + r), lpf), llf);     // log_p_where + log_p_what + ll (:147)
    if (a.extra_q) lq = __fadd_rn(__ldg(a.extra_q + r), lqf);                     // log_q_where + log_q_what (:148)
    if (a.log_q_f) a.log_q_f[r] = lq;
    if (a.dens) a.dens[r] = lp;
    if (a.ll_sum) a.ll_sum[r] = llf;
    if (a.log_w) {
        // one-shot (:149): log_p - log_q.   what step (:241): ll_f + log_p_f - log_q_f - (ll_b + log_p_b - log_q_b)
        float w = __fsub_rn(lp, lq);
        if (a.z_old) w = __fsub_rn(w, __fsub_rn(__fadd_rn(llb, lpb), lqb));
        a.log_w[r] = w;
    }
}

}  // namespace apgs

using namespace apgs;

static int check_stn(int64_t S, int64_t B, int64_t T, int64_t K, int64_t FP, int64_t DP, const char* who) {
    if (S <= 0 || B <= 0 || T <= 0 || K <= 0 || FP <= 1 || DP <= 1)
        return fail(APGS_E_SHAPE, "%s: all sizes must be positive (S=%lld B=%lld T=%lld K=%lld FP=%lld DP=%lld)", who,
                    (long long)S, (long long)B, (long long)T, (long long)K, (long long)FP, (long long)DP);
    if (S * B * T >= (1ll << 31) || FP > 4096 || DP > 4096 || K > 64) return fail(APGS_E_SHAPE, "%s: problem too large", who);
    return 0;
}

extern "C" int apgs_bmnist_crop_f32(const float* frames, int64_t stride_s, int64_t stride_b, int64_t stride_t,
                                    const float* z_where, const float* lin_dp, int64_t S, int64_t B, int64_t T,
                                    int64_t K, int64_t FP, int64_t DP, float* digits_out, void* stream) {
    if (int e = check_stn(S, B, T, K, FP, DP, "bmnist_crop")) return e;
    if (!frames || !z_where || !lin_dp || !digits_out) return fail(APGS_E_BADARG, "bmnist_crop: NULL pointer");
    StnArgs a = {};
    a.src = frames; a.ss = stride_s; a.sb = stride_b; a.st = stride_t; a.z_where = z_where; a.lin = lin_dp;
    a.dst = digits_out; a.S = (int)S; a.B = (int)B; a.T = (int)T; a.K = (int)K; a.FP = (int)FP; a.DP = (int)DP;
    const size_t smem = (size_t)K * 2 * DP * sizeof(int);
    k_stn_crop<<<(unsigned int)(S * B * T), 256, smem, (cudaStream_t)stream>>>(a);
    APGS_LAUNCH_OK("k_stn_crop launch");
    return 0;
}

extern "C" int apgs_bmnist_crop_bilinear_f32(const float* frames, int64_t stride_s, int64_t stride_b, int64_t stride_t,
                                             const float* z_where, const float* lin_dp, int64_t S, int64_t B, int64_t T,
                                             int64_t K, int64_t FP, int64_t DP, float* digits_out, void* stream) {
    if (int e = check_stn(S, B, T, K, FP, DP, "bmnist_crop_bilinear")) return e;
    if (!frames || !z_where || !lin_dp || !digits_out) return fail(APGS_E_BADARG, "bmnist_crop_bilinear: NULL pointer");
    StnArgs a = {};
    a.src = frames; a.ss = stride_s; a.sb = stride_b; a.st = stride_t; a.z_where = z_where; a.lin = lin_dp;
    a.dst = digits_out; a.S = (int)S; a.B = (int)B; a.T = (int)T; a.K = (int)K; a.FP = (int)FP; a.DP = (int)DP;
    const size_t smem = (size_t)K * 2 * DP * 8;
    k_stn_bilinear<false><<<(unsigned int)(S * B * T), 256, smem, (cudaStream_t)stream>>>(a);
    APGS_LAUNCH_OK("k_stn_bilinear<crop> launch");
    return 0;
}

extern "C" int apgs_bmnist_paste_bilinear_f32(const float* digits, const float* z_where, const float* lin_fp, int64_t S,
                                              int64_t B, int64_t T, int64_t K, int64_t FP, int64_t DP, float* frames_out,
                                              void* stream) {
    if (int e = check_stn(S, B, T, K, FP, DP, "bmnist_paste_bilinear")) return e;
    if (!digits || !z_where || !lin_fp || !frames_out) return fail(APGS_E_BADARG, "bmnist_paste_bilinear: NULL pointer");
    StnArgs a = {};
    a.src = digits; a.z_where = z_where; a.lin = lin_fp; a.dst = frames_out;
    a.S = (int)S; a.B = (int)B; a.T = (int)T; a.K = (int)K; a.FP = (int)FP; a.DP = (int)DP;
    const size_t smem = (size_t)K * 2 * FP * 8;
    if (smem > 48 * 1024) return fail(APGS_E_SHAPE, "bmnist_paste_bilinear: K*FP too large");
    k_stn_bilinear<true><<<(unsigned int)(S * B * T), 256, smem, (cudaStream_t)stream>>>(a);
    APGS_LAUNCH_OK("k_stn_bilinear<paste> launch");
    return 0;
}

static int launch_paste(bool with_ll, const float* digits, const float* z_where, const float* lin_fp,
                        const float* frames, int64_t fs, int64_t fb, int64_t ft, int64_t S, int64_t B, int64_t T,
                        int64_t K, int64_t FP, int64_t DP, float* out, float* ll, void* stream) {
    StnArgs a = {};
    a.src = digits; a.z_where = z_where; a.lin = lin_fp; a.dst = out; a.frames = frames; a.fs = fs; a.fb = fb; a.ft = ft;
    a.ll = ll; a.S = (int)S; a.B = (int)B; a.T = (int)T; a.K = (int)K; a.FP = (int)FP; a.DP = (int)DP;
    const size_t smem = (size_t)K * DP * DP * 4 + (size_t)K * 2 * FP * sizeof(int);
    if (smem > 200 * 1024) return fail(APGS_E_SHAPE, "bmnist_paste: K*DP*DP too large for shared memory");
    cudaStream_t st = (cudaStream_t)stream;
    if (with_ll) {
        APGS_CUDA_OK(ensure_dynamic_smem(reinterpret_cast<const void*>(k_stn_paste<true>), smem), "smem attr");
        k_stn_paste<true><<<(unsigned int)(S * B * T), 256, smem, st>>>(a);
    } else {
        APGS_CUDA_OK(ensure_dynamic_smem(reinterpret_cast<const void*>(k_stn_paste<false>), smem), "smem attr");
        k_stn_paste<false><<<(unsigned int)(S * B * T), 256, smem, st>>>(a);
    }
    APGS_LAUNCH_OK("k_stn_paste launch");
    return 0;
}

extern "C" int apgs_bmnist_paste_f32(const float* digits, const float* z_where, const float* lin_fp, int64_t S,
                                     int64_t B, int64_t T, int64_t K, int64_t FP, int64_t DP, float* frames_out,
                                     void* stream) {
    if (int e = check_stn(S, B, T, K, FP, DP, "bmnist_paste")) return e;
    if (!digits || !z_where || !lin_fp || !frames_out) return fail(APGS_E_BADARG, "bmnist_paste: NULL pointer");
    return launch_paste(false, digits, z_where, lin_fp, nullptr, 0, 0, 0, S, B, T, K, FP, DP, frames_out, nullptr, stream);
}

extern "C" int apgs_bmnist_paste_subtract_f32(const float* frame, int64_t frame_stride_s, int64_t frame_stride_b,
                                              const float* digit, int64_t digit_stride_s, int64_t digit_stride_b,
                                              const float* z_where, const float* lin_fp, int64_t S, int64_t B,
                                              int64_t FP, int64_t DP, float* out, void* stream) {
    if (int e = check_stn(S, B, 1, 1, FP, DP, "bmnist_paste_subtract")) return e;
    if (!frame || !digit || !z_where || !lin_fp || !out) return fail(APGS_E_BADARG, "bmnist_paste_subtract: NULL pointer");
    if (frame_stride_s < 0 || frame_stride_b < 0 || digit_stride_s < 0 || digit_stride_b <= 0)
        return fail(APGS_E_BADARG, "bmnist_paste_subtract: bad stride");
    StnArgs a = {};
    a.src = digit; a.ss = digit_stride_s; a.sb = digit_stride_b; a.z_where = z_where; a.lin = lin_fp; a.dst = out;
    a.frames = frame; a.fs = frame_stride_s; a.fb = frame_stride_b; a.ft = 0;
    a.S = (int)S; a.B = (int)B; a.T = 1; a.K = 1; a.FP = (int)FP; a.DP = (int)DP;
    const size_t smem = (size_t)DP * DP * 4 + (size_t)2 * FP * sizeof(int);
    APGS_CUDA_OK(ensure_dynamic_smem(reinterpret_cast<const void*>(k_stn_paste<false>), smem), "smem attr");
    k_stn_paste<false><<<(unsigned int)(S * B), 256, smem, (cudaStream_t)stream>>>(a);
    APGS_LAUNCH_OK("k_stn_paste launch");
    return 0;
}

extern "C" int apgs_bmnist_paste_ll_f32(const float* digits, const float* z_where, const float* lin_fp,
                                        const float* frames, int64_t stride_s, int64_t stride_b, int64_t stride_t,
                                        int64_t S, int64_t B, int64_t T, int64_t K, int64_t FP, int64_t DP,
                                        float* ll_out, float* recon_out, void* stream) {
    if (int e = check_stn(S, B, T, K, FP, DP, "bmnist_paste_ll")) return e;
    if (!digits || !z_where || !lin_fp || !frames || !ll_out) return fail(APGS_E_BADARG, "bmnist_paste_ll: NULL pointer");
    return launch_paste(true, digits, z_where, lin_fp, frames, stride_s, stride_b, stride_t, S, B, T, K, FP, DP,
                        recon_out, ll_out, stream);
}

extern "C" int apgs_bmnist_logw_where_f32(const float* q_mean, const float* q_std, const float* z_new,
                                          const float* z_prev, const float* z_old, const float* z_old_prev,
                                          const float* prior_mu0, const float* prior_sigma0, const float* prior_sigmat,
                                          const float* ll_f, const float* ll_b, int64_t S, int64_t B, int64_t K,
                                          int64_t D, float* log_w, float* log_q_f, float* log_p_f, float* log_q_b, float* log_p_b,
                                          void* stream) {
    if (S <= 0 || B <= 0 || K <= 0 || D <= 0) return fail(APGS_E_SHAPE, "bmnist_logw_where: sizes must be positive");
    if (!q_mean || !q_std || !z_new || !prior_mu0 || !prior_sigma0 || !prior_sigmat)
        return fail(APGS_E_BADARG, "bmnist_logw_where: NULL pointer");
    WhereArgs a = {q_mean, q_std, z_new, z_prev, z_old, z_old_prev, prior_mu0, prior_sigma0, prior_sigmat, ll_f, ll_b,
                   (long long)S * B, (int)K, (int)D, log_w, log_q_f, log_p_f, log_q_b, log_p_b};
    const long long grid = (a.R + 255) / 256;
    k_logw_where<<<(unsigned int)grid, 256, 0, (cudaStream_t)stream>>>(a);
    APGS_LAUNCH_OK("k_logw_where launch");
    return 0;
}

extern "C" int apgs_bmnist_logw_what_f32(const float* q_mu, const float* q_std, const float* z_new, const float* z_old,
                                         const float* ll_f, const float* ll_b, const float* extra_log_p,
                                         const float* extra_log_q, int64_t S, int64_t B, int64_t K, int64_t Z,
                                         int64_t T, float* log_w, float* log_q_f, float* density, float* ll_sum,
                                         void* stream) {
    if (S <= 0 || B <= 0 || K <= 0 || Z <= 0 || T <= 0) return fail(APGS_E_SHAPE, "bmnist_logw_what: sizes must be positive");
    if (!q_mu || !q_std || !z_new || !ll_f) return fail(APGS_E_BADARG, "bmnist_logw_what: NULL pointer");
    if (z_old && !ll_b) return fail(APGS_E_BADARG, "bmnist_logw_what: ll_b required with z_old");
    WhatArgs a = {q_mu, q_std, z_new, z_old, ll_f, ll_b, extra_log_p, extra_log_q, (long long)S * B, (int)K, (int)Z, (int)T,
                  log_w, log_q_f, density, ll_sum};
    const long long grid = (a.R + 255) / 256;
    k_logw_what<<<(unsigned int)grid, 256, 0, (cudaStream_t)stream>>>(a);
    APGS_LAUNCH_OK("k_logw_what launch");
    return 0;
}
