// Bouncing-MNIST template matching for sm_100a:
//   conved = conv2d(frame_left (1, S*B, FP, FP), template_k (S*B, 1, DP, DP), groups = S*B)   apgs/bmnist/objectives.py:72-73
//   conved = softmax(conved.view(S, B, CP*CP), -1),  CP = FP - DP + 1                          :74
// i.e. for every particle-instance pair ONE image is cross-correlated with ITS OWN template -- M = S*B independent
// 2-D correlations, 3.7 MFMA each at 96/28; no operand is shared between problems, so this is FMA-pipe work (a GEMM view
// would have N = 1).  PyTorch's depthwise kernel reaches ~2.5 TFMA/s on it and was 74 % of the bmnist sweep.
//
// One CTA per image.  The image sits in shared memory with a padded row stride (FP + 1: lanes of a warp own different
// rows of the same column segment, so their window loads hit 32 different banks), the template is read as broadcast
// LDS.128.  A thread owns SEG = 23 consecutive output columns of one output row (69 = 3 x 23): for each template row u
// it loads the 50-element image row window and the 28 template values into registers once and issues 23 x 28 = 644
// FMAs on them -- 12 FMAs per shared-memory load, 23 independent accumulation chains.  The softmax over the 4761
// scores of the image is fused behind it (block max / sum), so the raw score map never reaches HBM.
// 56 registers -> 5 CTAs per SM, so the C4 grid (640 images) is one wave.  Measured alternatives (ncu, same shape):
//   this kernel ......................................................... 0.129 ms, issue 76 %, FMA pipe 66 %
//   packed fma.rn.f32x2 on two column strips (half the FMA instructions) . 0.128 ms, issue 52 %, FMA pipe 70 %,
//     stall = math-pipe throttle: FFMA2 occupies the FMA pipe for two cycles, so the packed form moves the limit from
//     issue to the pipe without raising throughput -- the scalar form is kept.
#include "common.cuh"
#include "tf32_pack.cuh"

namespace apgs {

struct MatchArgs {
    const float* frame;        // image (s,b) at frame + s*fs + b*fb, FP*FP dense row-major
    long long fs, fb;
    const float* tmpl;         // template (s,b) at tmpl + s*ts + b*tb, DP*DP dense
    long long ts, tb;
    float* out;                // (S*B, CP*CP)
    int S, B, FP, DP, softmax;
    int packed;                // out is the packed TF32 hi / lo operand of k_linear_tc5pk (tf32_pack.cuh) instead of (S*B, CP*CP)
    // residual mode: the image is frame - paste(prev digit 0) - paste(prev digit 1), the frame_left of the sequential matches
    // (apgs/bmnist/objectives.py:87-89), formed while the image is staged; digit p of (s,b) at prev[p] + s*ps[p] + b*pb[p]
    // (DP*DP dense), its z_where at prev_z[p] + (s*B + b)*2, lin = affine_grid's base axis at the frame resolution
    const float* prev[2];
    long long ps[2], pb[2];
    const float* prev_z[2];
    const float* lin;
    int n_prev;
};

// source index of Affine_Transformer.digit_to_frame along one axis (bmnist.cu stn_index; affine_transformer.py:41-52)
__device__ __forceinline__ int match_stn_index(float lin, float scale, float t, int n_in) {
    const float v = __fadd_rn(__fmul_rn(lin, scale), t);
    const float u = __fmul_rn(__fadd_rn(v, 1.f), (float)(n_in - 1) * 0.5f);
    const float r = rintf(u);
    return (r >= 0.f && r <= (float)(n_in - 1)) ? (int)r : -1;
}

__device__ __forceinline__ float block_reduce(float v, float* s_red, bool is_max) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = is_max ? warp_max(v) : warp_sum(v);
    __syncthreads();                               // s_red may still be read from the previous reduction
    if (lane == 0) s_red[w] = v;
    __syncthreads();
    float r = s_red[0];
    for (int i = 1; i < nw; ++i) r = is_max ? fmaxf(r, s_red[i]) : r + s_red[i];
    return r;
}

// e / sum of the softmax.  1 <= sum <= CP*CP (the maximum contributes exp(0)), so for e >= 2^-113 numerator and quotient are
// normal numbers and div.rn stays on its fast path; below that (probabilities under 1e-34 -- a peaked map over 4761 positions
// has many, and as denormal numerators they sent most warps through the division's slow path: 0.134 -> 0.151 ms per call on
// residual frames) the quotient is e * (1 / sum), full speed on denormals and within one ulp of the division.
__device__ __forceinline__ float softmax_quotient(float e, float sum, float inv_sum) {
    return e >= 0x1p-113f ? __fdiv_rn(e, sum) : __fmul_rn(e, inv_sum);
}

template <int FP, int DP, int SEG>
__global__ void __launch_bounds__(224, 5) k_template_match(MatchArgs a) {
    constexpr int CP = FP - DP + 1, NSEG = CP / SEG, FS = FP + 1, WIN = SEG + DP - 1;
    static_assert(CP % SEG == 0 && DP % 4 == 0, "tile shape");
    static_assert(CP * NSEG <= 224, "one thread per (output row, column segment)");
    extern __shared__ __align__(16) float s_m[];
    float* s_t = s_m;                              // DP*DP (16-byte aligned rows)
    float* s_f = s_m + DP * DP;                    // FP x (FP+1)
    __shared__ float s_red[8];
    const int img = blockIdx.x, s = img / a.B, b = img % a.B;
    const float* f = a.frame + s * a.fs + b * a.fb;
    const float* t = a.tmpl + s * a.ts + b * a.tb;
    // frame_left = (frame - paste_0) - paste_1 (:87-89).  A pasted digit covers only the ~DP x DP pixels whose source index is
    // in range (the transformer magnifies by FP / DP), and v - 0.f == v: only that footprint is touched, digit by digit in
    // order -- bit-identical to k_stn_paste's residual mode.  The index vectors are built before the image is staged (their
    // global loads overlap it) and a thread fetches all of its digit pixels before it subtracts any (independent loads).
    __shared__ short s_pi[2][2][FP];                                      // [digit][rows / columns][frame coordinate] -> digit coordinate
    __shared__ int s_lo[2][2], s_n[2][2];                                 // first in-range frame coordinate and how many follow
    if (a.n_prev > 0) {
        const float scale = (float)((double)FP / (double)DP), tr = (float)((double)(FP - DP) / (double)DP);
        for (int i = threadIdx.x; i < a.n_prev * 2 * FP; i += blockDim.x) {
            const int p = i / (2 * FP), ax = (i / FP) & 1, j = i % FP;
            const float* zw = a.prev_z[p] + ((long long)s * a.B + b) * 2;
            const float tt = ax == 0 ? __fmul_rn(__ldg(zw + 1), tr) : -__fmul_rn(__ldg(zw), tr);
            s_pi[p][ax][j] = (short)match_stn_index(__ldg(a.lin + j), scale, tt, DP);
        }
    }
    for (int i = threadIdx.x; i < FP * FP; i += blockDim.x) s_f[(i / FP) * FS + (i % FP)] = __ldg(f + i);
    if (a.n_prev > 0) {
        __syncthreads();
        const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
        if (w < a.n_prev * 2) {                                           // warp w: one (digit, axis); the in-range run is contiguous
            static_assert(FP % 32 == 0, "lanes stride over the frame coordinates");
            int lo = FP, hi = -1;
#pragma unroll
            for (int j = lane; j < FP; j += 32)
                if (s_pi[w >> 1][w & 1][j] >= 0) { lo = min(lo, j); hi = max(hi, j); }
            lo = __reduce_min_sync(0xffffffffu, lo);
            hi = __reduce_max_sync(0xffffffffu, hi);
            if (lane == 0) { s_lo[w >> 1][w & 1] = lo; s_n[w >> 1][w & 1] = hi >= lo ? hi - lo + 1 : 0; }
        }
        __syncthreads();
        constexpr int PER = 5;                                            // footprint pixels per thread: (DP + 2)^2 <= 224 * 5
        for (int p = 0; p < a.n_prev; ++p) {
            const float* dg = a.prev[p] + s * a.ps[p] + b * a.pb[p];
            const int nr = s_n[p][0], nc = s_n[p][1], r0 = s_lo[p][0], c0 = s_lo[p][1];
            for (int base = 0; base < nr * nc; base += PER * (int)blockDim.x) {
                float d[PER];
                int at[PER];
#pragma unroll
                for (int q = 0; q < PER; ++q) {
                    const int i = base + q * (int)blockDim.x + (int)threadIdx.x;
                    at[q] = -1;
                    d[q] = 0.f;
                    if (i < nr * nc) {
                        const int r = r0 + i / nc, c = c0 + i % nc;
                        const int rr = s_pi[p][0][r], cc = s_pi[p][1][c];
                        if (rr >= 0 && cc >= 0) { at[q] = r * FS + c; d[q] = __ldg(dg + rr * DP + cc); }
                    }
                }
#pragma unroll
                for (int q = 0; q < PER; ++q)
                    if (at[q] >= 0) s_f[at[q]] = __fsub_rn(s_f[at[q]], d[q]);
            }
            __syncthreads();                                              // digit 1 subtracts from what digit 0 left
        }
    }
    for (int i = threadIdx.x; i < DP * DP; i += blockDim.x) s_t[i] = __ldg(t + i);
    __syncthreads();
    const bool live = threadIdx.x < CP * NSEG;
    const int i0 = live ? threadIdx.x % CP : 0, seg = live ? threadIdx.x / CP : 0;
    float acc[SEG];
#pragma unroll
    for (int c = 0; c < SEG; ++c) acc[c] = 0.f;
    if (live) {
#pragma unroll 1
        for (int u = 0; u < DP; ++u) {
            const float* row = s_f + (i0 + u) * FS + seg * SEG;
            float w[WIN], tr[DP];
#pragma unroll
            for (int q = 0; q < WIN; ++q) w[q] = row[q];
#pragma unroll
            for (int q = 0; q < DP / 4; ++q) {
                const float4 v = *reinterpret_cast<const float4*>(s_t + u * DP + q * 4);
                tr[q * 4] = v.x; tr[q * 4 + 1] = v.y; tr[q * 4 + 2] = v.z; tr[q * 4 + 3] = v.w;
            }
#pragma unroll
            for (int v = 0; v < DP; ++v)
#pragma unroll
                for (int c = 0; c < SEG; ++c) acc[c] = fmaf(tr[v], w[v + c], acc[c]);
        }
    }
    float* o = a.out + (size_t)img * CP * CP + i0 * CP + seg * SEG;
    if (a.softmax) {
        float mx = -INFINITY;
        if (live) {
#pragma unroll
            for (int c = 0; c < SEG; ++c) mx = fmaxf(mx, acc[c]);
        }
        mx = block_reduce(mx, s_red, true);
        float se = 0.f;
        if (live) {
#pragma unroll
            for (int c = 0; c < SEG; ++c) { acc[c] = expf(acc[c] - mx); se += acc[c]; }
        }
        se = block_reduce(se, s_red, false);
        // The normalised row goes through shared memory (the image is no longer needed: both reductions have passed a CTA
        // barrier behind the main loop), so that global memory sees whole lines instead of 23 scalar stores per thread at a
        // lane stride of 69 floats (0.137 -> ... ms per call at the C4 shape).
        constexpr int NCH = (CP * CP + 31) / 32;                        // 32-column chunks of the packed operand
        static_assert(NCH * 32 <= FP * FS, "the row fits the image region");
        float* s_row = s_f;
        const float inv_se = __frcp_rn(se);
        if (live) {
            const int k0 = i0 * CP + seg * SEG;                          // lanes: consecutive i0 -> stride 69 = 5 mod 32: no bank conflicts
#pragma unroll
            for (int c = 0; c < SEG; ++c) s_row[k0 + c] = softmax_quotient(acc[c], se, inv_se);
        }
        for (int k = CP * CP + (int)threadIdx.x; k < NCH * 32; k += blockDim.x) s_row[k] = 0.f;
        __syncthreads();
        if (!a.packed) {
            float* orow = a.out + (size_t)img * CP * CP;
            for (int k = threadIdx.x; k < CP * CP; k += blockDim.x) orow[k] = s_row[k];
        } else {
            // the next consumer is Enc_coor's Linear(CP*CP -> 400) on the tensor cores: the row already split into TF32 halves
            // and in the MMA's shared-memory order (tf32_pack.cuh), so that the GEMM's stages are plain bulk copies.  One
            // 16-byte store per (group of four columns, half); columns up to the chunk boundary are written as zeros.
            for (int g = threadIdx.x; g < NCH * 8; g += blockDim.x) {
                const float4 v = *reinterpret_cast<const float4*>(s_row + 4 * g);
                unsigned int h0, h1, h2, h3, l0, l1, l2, l3;
                split_tf32(v.x, h0, l0); split_tf32(v.y, h1, l1); split_tf32(v.z, h2, l2); split_tf32(v.w, h3, l3);
                float* d = a.out + packed_index(kPackRowsX, img, 4 * g, NCH);
                *reinterpret_cast<uint4*>(d) = make_uint4(h0, h1, h2, h3);
                *reinterpret_cast<uint4*>(d + kPackRowsX * 32) = make_uint4(l0, l1, l2, l3);
            }
        }
    } else if (live) {
#pragma unroll
        for (int c = 0; c < SEG; ++c) o[c] = acc[c];
    }
}

// any other (FP, DP): one thread per score, operands in shared memory.  Correct, not tuned.
__global__ void __launch_bounds__(256) k_template_match_generic(MatchArgs a) {
    extern __shared__ __align__(16) float s_m[];
    __shared__ float s_red[8];
    const int FP = a.FP, DP = a.DP, CP = FP - DP + 1;
    float* s_t = s_m;
    float* s_f = s_m + DP * DP;
    const int img = blockIdx.x, s = img / a.B, b = img % a.B;
    const float* f = a.frame + s * a.fs + b * a.fb;
    const float* t = a.tmpl + s * a.ts + b * a.tb;
    for (int i = threadIdx.x; i < FP * FP; i += blockDim.x) s_f[i] = __ldg(f + i);
    for (int i = threadIdx.x; i < DP * DP; i += blockDim.x) s_t[i] = __ldg(t + i);
    __syncthreads();
    float* o = a.out + (size_t)img * CP * CP;
    float mx = -INFINITY;
    for (int e = threadIdx.x; e < CP * CP; e += blockDim.x) {
        const int i = e / CP, j = e % CP;
        float acc = 0.f;
        for (int u = 0; u < DP; ++u)
            for (int v = 0; v < DP; ++v) acc = fmaf(s_t[u * DP + v], s_f[(i + u) * FP + j + v], acc);
        o[e] = acc;
        mx = fmaxf(mx, acc);
    }
    if (!a.softmax) return;
    mx = block_reduce(mx, s_red, true);
    float se = 0.f;
    for (int e = threadIdx.x; e < CP * CP; e += blockDim.x) { const float v = expf(o[e] - mx); o[e] = v; se += v; }
    se = block_reduce(se, s_red, false);
    for (int e = threadIdx.x; e < CP * CP; e += blockDim.x) o[e] = __fdiv_rn(o[e], se);
}

}  // namespace apgs

using namespace apgs;

struct MatchPrev {
    const float* digit[2];
    int64_t ds[2], db[2];
    const float* z[2];
    const float* lin;
    int n;
};

static int template_match_launch(const float* frame, int64_t frame_stride_s, int64_t frame_stride_b, const float* tmpl,
                                 int64_t tmpl_stride_s, int64_t tmpl_stride_b, int64_t S, int64_t B, int64_t FP, int64_t DP,
                                 int softmax, int packed, float* out, void* stream, const MatchPrev* prev = nullptr) {
    if (S <= 0 || B <= 0 || FP <= 0 || DP <= 0 || DP > FP || S * B > 0x7fffffffLL)
        return fail(APGS_E_SHAPE, "bmnist_template_match: bad sizes S=%lld B=%lld FP=%lld DP=%lld", (long long)S, (long long)B,
                    (long long)FP, (long long)DP);
    if (!frame || !tmpl || !out) return fail(APGS_E_BADARG, "bmnist_template_match: NULL pointer");
    if (frame_stride_s < 0 || frame_stride_b < 0 || tmpl_stride_s < 0 || tmpl_stride_b < 0)
        return fail(APGS_E_BADARG, "bmnist_template_match: negative stride");
    if (packed && !(FP == 96 && DP == 28))
        return fail(APGS_E_UNSUPPORTED, "bmnist_template_match_packed: compiled for FP=96, DP=28 only (got %lld, %lld)", (long long)FP,
                    (long long)DP);
    MatchArgs a = {frame, (long long)frame_stride_s, (long long)frame_stride_b, tmpl, (long long)tmpl_stride_s,
                   (long long)tmpl_stride_b, out, (int)S, (int)B, (int)FP, (int)DP, softmax, packed,
                   {nullptr, nullptr}, {0, 0}, {0, 0}, {nullptr, nullptr}, nullptr, 0};
    if (prev != nullptr && prev->n > 0) {
        if (!(FP == 96 && DP == 28))
            return fail(APGS_E_UNSUPPORTED, "bmnist_template_match_residual: compiled for FP=96, DP=28 only (got %lld, %lld)",
                        (long long)FP, (long long)DP);
        if (prev->n > 2 || !prev->lin) return fail(APGS_E_BADARG, "bmnist_template_match_residual: at most two digits, lin required");
        for (int p = 0; p < prev->n; ++p) {
            if (!prev->digit[p] || !prev->z[p]) return fail(APGS_E_BADARG, "bmnist_template_match_residual: NULL digit / z_where");
            if (prev->ds[p] < 0 || prev->db[p] < 0) return fail(APGS_E_BADARG, "bmnist_template_match_residual: negative stride");
            a.prev[p] = prev->digit[p]; a.ps[p] = (long long)prev->ds[p]; a.pb[p] = (long long)prev->db[p]; a.prev_z[p] = prev->z[p];
        }
        a.lin = prev->lin;
        a.n_prev = prev->n;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned int grid = (unsigned int)(S * B);
    if (FP == 96 && DP == 28) {
        const size_t smem = (size_t)(28 * 28 + 96 * 97) * 4;
        APGS_CUDA_OK(ensure_dynamic_smem(reinterpret_cast<const void*>(k_template_match<96, 28, 23>), smem), "smem attr");
        k_template_match<96, 28, 23><<<grid, 224, smem, st>>>(a);
    } else {
        const size_t smem = (size_t)(DP * DP + FP * FP) * 4;
        if (smem > 200 * 1024) return fail(APGS_E_SHAPE, "bmnist_template_match: FP=%lld too large", (long long)FP);
        APGS_CUDA_OK(ensure_dynamic_smem(reinterpret_cast<const void*>(k_template_match_generic), smem), "smem attr");
        k_template_match_generic<<<grid, 256, smem, st>>>(a);
    }
    APGS_LAUNCH_OK("k_template_match launch");
    return 0;
}

extern "C" int apgs_bmnist_template_match_f32(const float* frame, int64_t frame_stride_s, int64_t frame_stride_b,
                                              const float* tmpl, int64_t tmpl_stride_s, int64_t tmpl_stride_b,
                                              int64_t S, int64_t B, int64_t FP, int64_t DP, int softmax, float* out,
                                              void* stream) {
    return template_match_launch(frame, frame_stride_s, frame_stride_b, tmpl, tmpl_stride_s, tmpl_stride_b, S, B, FP, DP, softmax, 0,
                                 out, stream);
}

extern "C" int apgs_bmnist_template_match_packed_f32(const float* frame, int64_t frame_stride_s, int64_t frame_stride_b,
                                                     const float* tmpl, int64_t tmpl_stride_s, int64_t tmpl_stride_b,
                                                     int64_t S, int64_t B, int64_t FP, int64_t DP, float* out_packed,
                                                     void* stream) {
    if (out_packed && !aligned(out_packed, 16)) return fail(APGS_E_BADARG, "bmnist_template_match_packed: out must be 16-byte aligned");
    return template_match_launch(frame, frame_stride_s, frame_stride_b, tmpl, tmpl_stride_s, tmpl_stride_b, S, B, FP, DP, 1, 1,
                                 out_packed, stream);
}

extern "C" int apgs_bmnist_template_match_residual_f32(const float* frame, int64_t frame_stride_s, int64_t frame_stride_b,
                                                       const float* tmpl, int64_t tmpl_stride_s, int64_t tmpl_stride_b,
                                                       const float* digit0, int64_t digit0_stride_s, int64_t digit0_stride_b,
                                                       const float* z_where0, const float* digit1, int64_t digit1_stride_s,
                                                       int64_t digit1_stride_b, const float* z_where1, const float* lin_fp,
                                                       int64_t S, int64_t B, int64_t FP, int64_t DP, int softmax, int packed,
                                                       float* out, void* stream) {
    if (packed && out && !aligned(out, 16)) return fail(APGS_E_BADARG, "bmnist_template_match_residual: packed out must be 16-byte aligned");
    if (packed && !softmax) return fail(APGS_E_BADARG, "bmnist_template_match_residual: the packed output is the softmax map");
    MatchPrev pv = {{digit0, digit1}, {digit0_stride_s, digit1_stride_s}, {digit0_stride_b, digit1_stride_b}, {z_where0, z_where1},
                    lin_fp, digit0 ? (digit1 ? 2 : 1) : 0};
    if (!digit0 && digit1) return fail(APGS_E_BADARG, "bmnist_template_match_residual: digit1 without digit0");
    return template_match_launch(frame, frame_stride_s, frame_stride_b, tmpl, tmpl_stride_s, tmpl_stride_b, S, B, FP, DP, softmax,
                                 packed, out, stream, &pv);
}
