// out = act(x W^T + b) for the one dense contraction on the bouncing-MNIST path: Enc_coor's Linear(4761 -> 400) on the softmax of
// the template-match map (apgs/bmnist/models.py:13-16, called 150 times per 5-sweep objective with M = S*B = 640 rows).
// The shape is skinny (640 x 400 outputs, K = 4761, rows K-contiguous but only 4-byte aligned), which cuBLAS serves with its
// SIMT fp32 kernels plus split-K (~100 us per call).  Here the product runs on the tensor cores with fp32-level accuracy:
//   * 3xTF32 error compensation: a = a_hi + a_lo, w = w_hi + w_lo with the halves exactly representable in TF32
//     (cvt.rna.tf32), and  a w ~= a_hi w_hi + a_lo w_hi + a_hi w_lo  accumulated in fp32 by mma.sync.m16n8k8 -- the dropped
//     a_lo w_lo term is ~2^-22 relative; the halves are made in registers, no second copy of the operands exists;
//   * split-K over the 148 SMs: a CTA owns a 64 x 80 output tile and one slice of K (multiples of 32), writes its partial
//     tile to a workspace, and a second tiny kernel adds the slices in a fixed order with the bias and the ReLU
//     (deterministic: no atomics);
//   * operands staged through shared memory with 4-byte cp.async (rows are not 16-byte aligned), two stages; rows padded to
//     36 floats so that the fragment loads of a warp hit 32 distinct banks.
// Measured on B200 (benchmarks/linear_bench.py, M = 640): 0.102 ms per call against 0.115 ms for torch's fp32 linear, error
// 4e-7 of sum|x w| against 1.7e-6.  The warp-level TF32 MMA issues once per ~32 cycles and scheduler on sm_100 (32 MAC per
// cycle: the rate of the CUDA cores), so the three MMAs per product bound this kernel; a plain fp32 kernel of the same tiling
// (128 x 80 CTA tile, 8 x 8 outputs per thread, packed FFMA2) was built and measured at 0.128-0.15 ms (barrier / load-latency
// bound at 2-3 CTAs per SM) and dropped.  The next step for this layer is tcgen05 (kind::tf32 at full tensor-core rate, the
// operand halves materialised as separate shared-memory tiles).
#include <stdlib.h>
#include "tile.cuh"

namespace apgs {

constexpr int kLBM = 64, kLBN = 80, kLBK = 32, kLPad = 36, kLThreads = 128;
#ifndef APGS_LIN_CTAS_PER_SM
#define APGS_LIN_CTAS_PER_SM 4
#endif

__device__ __forceinline__ unsigned int to_tf32(float v) {
    unsigned int r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ void split_tf32(float v, unsigned int& hi, unsigned int& lo) {
    hi = to_tf32(v);
    lo = to_tf32(v - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const unsigned int (&a)[4], const unsigned int (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void cp_async4(float* dst, const float* src, bool ok) {
    const int n = ok ? 4 : 0;                                   // src-size 0: the destination is zero-filled
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(n) : "memory");
}

// partial[split][M][N] = x[:, k-slice] W[:, k-slice]^T
__global__ void __launch_bounds__(kLThreads) k_linear_tf32x3(const float* __restrict__ x, const float* __restrict__ W, int M, int N, int K,
                                                             int k_per_split, float* __restrict__ partial) {
    __shared__ __align__(16) float As[2][kLBM][kLPad];
    __shared__ __align__(16) float Ws[2][kLBN][kLPad];
    const int n0 = (int)blockIdx.x * kLBN, m0 = (int)blockIdx.y * kLBM, split = (int)blockIdx.z;
    const int kb = split * k_per_split, ke = min(K, kb + k_per_split);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm = (warp >> 1) * 32, wn = (warp & 1) * 40;           // warp tile 32 x 40: 2 m16 tiles x 5 n8 tiles

    auto stage = [&](int buf, int k0) {
        // 64 x 32 of x and 80 x 32 of W: lane = column of the chunk (a warp copies 32 consecutive floats of a row per request),
        // warp w takes rows w, w + 4, ...: one row pointer stepped by 4 K, one column test per chunk
        const int kc = k0 + lane;
        const bool kok = kc < ke;
        const float* xs = x + (size_t)m0 * K + kc;
        const float* wsrc = W + (size_t)n0 * K + kc;
#pragma unroll
        for (int jj = 0; jj < kLBM / 4; ++jj) {
            const int r = warp + 4 * jj;
            const bool ok = kok && (m0 + r < M);
            cp_async4(&As[buf][r][lane], ok ? xs + (size_t)r * K : x, ok);
        }
#pragma unroll
        for (int jj = 0; jj < kLBN / 4; ++jj) {
            const int r = warp + 4 * jj;
            const bool ok = kok && (n0 + r < N);
            cp_async4(&Ws[buf][r][lane], ok ? wsrc + (size_t)r * K : W, ok);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    // the tensor core adds into its fp32 accumulator with truncation; keeping the running sum of a chunk (32 k) small and adding
    // it to the total with an ordinary rounded FADD keeps the result at the accuracy of an fp32 GEMM
    float acc[2][5][4], tot[2][5][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 5; ++j)
#pragma unroll
            for (int c = 0; c < 4; ++c) { acc[i][j][c] = 0.f; tot[i][j][c] = 0.f; }

    const int nchunks = (ke - kb + kLBK - 1) / kLBK;
    if (nchunks > 0) stage(0, kb);
    for (int ch = 0; ch < nchunks; ++ch) {
        const int buf = ch & 1;
        if (ch + 1 < nchunks) {
            stage(buf ^ 1, kb + (ch + 1) * kLBK);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < kLBK; kk += 8) {
            unsigned int ah[2][4], al[2][4];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const float* ap = &As[buf][wm + i * 16 + g][kk + t];
                split_tf32(ap[0], ah[i][0], al[i][0]);                  // (row g,     col t)
                split_tf32(ap[8 * kLPad], ah[i][1], al[i][1]);          // (row g + 8, col t)
                split_tf32(ap[4], ah[i][2], al[i][2]);                  // (row g,     col t + 4)
                split_tf32(ap[8 * kLPad + 4], ah[i][3], al[i][3]);      // (row g + 8, col t + 4)
            }
#pragma unroll
            for (int j = 0; j < 5; ++j) {
                const float* bp = &Ws[buf][wn + j * 8 + g][kk + t];     // B[k][n] = W[n][k]: (k = t, n = g), (k = t + 4, n = g)
                unsigned int bh[2], bl[2];
                split_tf32(bp[0], bh[0], bl[0]);
                split_tf32(bp[4], bh[1], bl[1]);
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    mma_tf32(acc[i][j], al[i], bh);                     // small terms first
                    mma_tf32(acc[i][j], ah[i], bl);
                    mma_tf32(acc[i][j], ah[i], bh);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 5; ++j)
#pragma unroll
                for (int c = 0; c < 4; ++c) { tot[i][j][c] += acc[i][j][c]; acc[i][j][c] = 0.f; }
        __syncthreads();                                                // the buffer is refilled two chunks later
    }
    float* out = partial + (size_t)split * M * N;
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 5; ++j) {
            const int r = m0 + wm + i * 16 + g, c = n0 + wn + j * 8 + 2 * t;
            if (r < M) {
                if (c < N) out[(size_t)r * N + c] = tot[i][j][0];
                if (c + 1 < N) out[(size_t)r * N + c + 1] = tot[i][j][1];
            }
            if (r + 8 < M) {
                if (c < N) out[(size_t)(r + 8) * N + c] = tot[i][j][2];
                if (c + 1 < N) out[(size_t)(r + 8) * N + c + 1] = tot[i][j][3];
            }
        }
}

__global__ void __launch_bounds__(256) k_linear_reduce(const float* __restrict__ partial, const float* __restrict__ bias, long long MN,
                                                       int N, int splits, int relu, float* __restrict__ out) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= MN) return;
    float v = bias != nullptr ? __ldg(bias + (int)(i % N)) : 0.f;
    for (int s = 0; s < splits; ++s) v += __ldg(partial + (size_t)s * MN + i);      // fixed order: deterministic
    out[i] = relu ? fmaxf(v, 0.f) : v;
}

static int linear_splits(long long M, long long N, long long K) {
    const long long tiles = ((M + kLBM - 1) / kLBM) * ((N + kLBN - 1) / kLBN);
    static const int per_sm = [] {                                             // measurement knob: CTAs per SM the split count aims at
        const char* e = getenv("APGS_LIN_CTAS_PER_SM");
        const int v = e ? atoi(e) : 0;
        return (v >= 1 && v <= 16) ? v : APGS_LIN_CTAS_PER_SM;
    }();
    long long s = ((long long)per_sm * sm_count() + tiles - 1) / tiles;       // a CTA is four warps: several per SM hide the staging latency
    const long long smax = (K + 4 * kLBK - 1) / (4 * kLBK);                    // at least four chunks per slice
    if (s > smax) s = smax;
    if (s < 1) s = 1;
    if (s > 64) s = 64;
    return (int)s;
}

}  // namespace apgs

using namespace apgs;

extern "C" size_t apgs_linear_tf32x3_workspace_bytes(int64_t M, int64_t N, int64_t K) {
    if (M <= 0 || N <= 0 || K <= 0) return 0;
    return (size_t)linear_splits(M, N, K) * (size_t)M * (size_t)N * sizeof(float);
}

extern "C" int apgs_linear_tf32x3_f32(const float* x, const float* W, const float* bias, int64_t M, int64_t K, int64_t N,
                                      int relu, float* out, void* workspace, size_t workspace_bytes, void* stream) {
    if (M <= 0 || N <= 0 || K <= 0 || M >= (1ll << 24) || N >= (1ll << 24) || K >= (1ll << 30))
        return fail(APGS_E_SHAPE, "linear_tf32x3: bad sizes M=%lld K=%lld N=%lld", (long long)M, (long long)K, (long long)N);
    if (!x || !W || !out) return fail(APGS_E_BADARG, "linear_tf32x3: NULL pointer");
    const int splits = linear_splits(M, N, K);
    const size_t need = (size_t)splits * (size_t)M * (size_t)N * sizeof(float);
    if (!workspace || workspace_bytes < need)
        return fail(APGS_E_WORKSPACE, "linear_tf32x3: workspace of %zu bytes required, got %zu", need, workspace_bytes);
    int kps = (int)((K + splits - 1) / splits);
    kps = (kps + kLBK - 1) / kLBK * kLBK;                                       // slices start on chunk boundaries
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid((unsigned int)((N + kLBN - 1) / kLBN), (unsigned int)((M + kLBM - 1) / kLBM), (unsigned int)splits);
    k_linear_tf32x3<<<grid, kLThreads, 0, st>>>(x, W, (int)M, (int)N, (int)K, kps, reinterpret_cast<float*>(workspace));
    APGS_LAUNCH_OK("k_linear_tf32x3 launch");
    const long long MN = (long long)M * N;
    k_linear_reduce<<<(unsigned int)((MN + 255) / 256), 256, 0, st>>>(reinterpret_cast<const float*>(workspace), bias, MN, (int)N,
                                                                      splits, relu, out);
    APGS_LAUNCH_OK("k_linear_reduce launch");
    return 0;
}

