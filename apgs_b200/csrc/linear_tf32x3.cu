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
#include "tf32_pack.cuh"

namespace apgs {

constexpr int kLBM = 64, kLBN = 80, kLBK = 32, kLPad = 36, kLThreads = 128;
#ifndef APGS_LIN_CTAS_PER_SM
#define APGS_LIN_CTAS_PER_SM 4
#endif

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
                                                       int N, int splits, int relu, float* __restrict__ out,
                                                       const int* __restrict__ err = nullptr) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= MN) return;
    if (err != nullptr && *err != 0) { out[i] = __int_as_float(0x7fc00000); return; }   // an MMA pipeline timed out: NaN, loudly, without a host sync
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


// ---------------------------------------------------------------------------------------------------------------------------
// The same 3xTF32 product on the 5th-generation tensor cores: tcgen05.mma kind::tf32 issued by one thread, operand halves as
// separate shared-memory tiles (K-major, no swizzle: core matrices of 8 rows x 16 bytes, 128 bytes apart along K, 1 KB apart
// along the rows), accumulator (128 x 80 fp32) in tensor memory, read back with tcgen05.ld for the split-K partial tile.
// CTA tile 128 x 80, K chunks of 32 in two stages: all four warps stage a chunk (scalar loads -- rows are only 4-byte
// aligned -- split into TF32 halves in registers, one 16-byte store per core-matrix row), fence to the async proxy, one
// thread issues 12 MMAs and commits them to the stage's mbarrier, which releases the stage two chunks later.
namespace apgs {

#define APGS_TC5_CTAS 2
constexpr int kTBM = 128, kTBN = 80, kTBK = 32, kTThreads = 128, kTCols = 128;
constexpr int kTKH = kTBK / 16;                                                  // 16-k halves per chunk
constexpr int kTRounds = (kTBM / 8 + kTBN / 8) * kTKH / 4;                       // staging rounds per warp and chunk
constexpr int kTSBO = kTBK / 4 * 128;                                            // bytes between 8-row groups of an operand tile
constexpr int kTABytes = kTBM * kTBK * 4, kTBBytes = kTBN * kTBK * 4;            // one half (hi or lo) of a stage
constexpr int kTStage = 2 * kTABytes + 2 * kTBBytes;                             // A_hi | A_lo | B_hi | B_lo

__device__ __forceinline__ unsigned long long umma_desc(unsigned int smem_addr) {
    // K-major, SWIZZLE_NONE: leading (K) byte offset 128, stride (8-row group) byte offset 1024, descriptor version 1 (sm_100)
    return (unsigned long long)((smem_addr >> 4) & 0x3FFFu) | (8ull << 16) | ((unsigned long long)(kTSBO >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void umma_tf32(unsigned int tmem_d, unsigned long long a, unsigned long long b, unsigned int idesc, unsigned int acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                 ::"r"(tmem_d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
// bounded wait: a malformed descriptor must end in an error code, not in a hung GPU
__device__ __forceinline__ bool mbar_wait_bounded(unsigned long long* bar, unsigned int parity) {
    for (int it = 0; it < (1 << 22); ++it) {
        unsigned int ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (ok) return true;
    }
    return false;
}

__global__ void __launch_bounds__(kTThreads) k_linear_tc5(const float* __restrict__ x, const float* __restrict__ W, int M, int N, int K,
                                                          int k_per_split, float* __restrict__ partial, int* __restrict__ err) {
    extern __shared__ __align__(1024) unsigned char tsm[];
    __shared__ unsigned long long bars[2];
    __shared__ unsigned int tmem_slot;
    const int n0 = (int)blockIdx.x * kTBN, m0 = (int)blockIdx.y * kTBM, split = (int)blockIdx.z;
    const int kb = split * k_per_split, ke = min(K, kb + k_per_split);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(kTCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned int tmem = tmem_slot;
    // instruction descriptor: D = F32, A = B = TF32, both K-major, N = 80, M = 128
    constexpr unsigned int idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned int)(kTBN >> 3) << 17) | ((unsigned int)(kTBM >> 4) << 24);

    const int row8 = lane & 7, kg = lane >> 3;                      // a warp fills 8 rows x 16 k per round: one 16-byte store per lane
    const int nchunks = (ke - kb + kTBK - 1) / kTBK;
    bool ok = true;
    for (int ch = 0; ch < nchunks; ++ch) {
        const int s = ch & 1, k0 = kb + ch * kTBK;
        unsigned char* st = tsm + (size_t)s * kTStage;
        // ---- global -> registers: rounds of (8 rows x 16 k): lane -> (row lane & 7, four consecutive k of group lane >> 3), so a
        // lane owns one 16-byte row of a core matrix and stores it with one 16-byte store (a per-core-matrix mapping with
        // 4-byte stores and four times fewer sectors per load measured slower: 0.098 vs 0.089 ms).  A: 16 row groups x 2 halves of
        // the chunk, B: 10 x 2: 52 rounds, 13 per warp.
        float v[kTRounds][4];
#pragma unroll
        for (int it = 0; it < kTRounds; ++it) {
            const int rd = warp + 4 * it;                            // round index
            const bool isA = rd < kTBM / 8 * kTKH;
            const int q = isA ? rd : rd - kTBM / 8 * kTKH;
            const int rg = q / kTKH, half = q % kTKH;                // row group, which 16 k of the chunk
            const int r = rg * 8 + row8, kk = k0 + half * 16 + kg * 4;
            const bool rok = isA ? (m0 + r < M) : (n0 + r < N);
            const float* src = (isA ? x + (size_t)(m0 + r) * K : W + (size_t)(n0 + r) * K) + kk;
#pragma unroll
            for (int i = 0; i < 4; ++i) v[it][i] = (rok && kk + i < ke) ? __ldg(src + i) : 0.f;
        }
        // ---- the stage must have been released by the MMAs of chunk ch - 2
        if (ch >= 2) ok = ok && mbar_wait_bounded(&bars[s], (unsigned int)(((ch >> 1) - 1) & 1));
#pragma unroll
        for (int it = 0; it < kTRounds; ++it) {
            const int rd = warp + 4 * it;
            const bool isA = rd < kTBM / 8 * kTKH;
            const int q = isA ? rd : rd - kTBM / 8 * kTKH;
            const int rg = q / kTKH, half = q % kTKH;
            unsigned int hi[4], lo[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) split_tf32(v[it][i], hi[i], lo[i]);
            unsigned char* base = st + (isA ? 0 : 2 * kTABytes);
            const int half_bytes = isA ? kTABytes : kTBBytes;
            const int off = rg * kTSBO + (half * 4 + kg) * 128 + row8 * 16;
            *reinterpret_cast<uint4*>(base + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<uint4*>(base + half_bytes + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // generic-proxy stores -> visible to the tensor core
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const unsigned int a_hi = smem_u32(st), a_lo = a_hi + kTABytes, b_hi = a_hi + 2 * kTABytes, b_lo = b_hi + kTBBytes;
#pragma unroll
            for (int j = 0; j < kTBK / 8; ++j) {                            // one MMA covers 8 k = two core-matrix columns = 256 bytes
                const unsigned int o = (unsigned int)j * 256u;
                umma_tf32(tmem, umma_desc(a_lo + o), umma_desc(b_hi + o), idesc, (ch > 0 || j > 0) ? 1u : 0u);
                umma_tf32(tmem, umma_desc(a_hi + o), umma_desc(b_lo + o), idesc, 1u);
                umma_tf32(tmem, umma_desc(a_hi + o), umma_desc(b_hi + o), idesc, 1u);
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.b64 [%0];" ::"r"(smem_u32(&bars[s])) : "memory");
        }
    }
    // ---- all MMAs done: the last commit arrives after every earlier one
    if (nchunks > 0) {
        const int last = nchunks - 1;
        ok = ok && mbar_wait_bounded(&bars[last & 1], (unsigned int)((last >> 1) & 1));
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (!ok && tid == 0 && err != nullptr) atomicExch(err, 1);
    // ---- accumulator rows 32 w .. 32 w + 31 belong to warp w: 80 fp32 columns, 16 at a time
    float* out = partial + (size_t)split * M * N;
    const int r = m0 + warp * 32 + lane;
#pragma unroll
    for (int c0 = 0; c0 < kTBN; c0 += 16) {
        unsigned int d[16];
        if (nchunks > 0 && ok) {
            const unsigned int taddr = tmem + ((unsigned int)(warp * 32) << 16) + (unsigned int)c0;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                         : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]), "=r"(d[4]), "=r"(d[5]), "=r"(d[6]), "=r"(d[7]),
                           "=r"(d[8]), "=r"(d[9]), "=r"(d[10]), "=r"(d[11]), "=r"(d[12]), "=r"(d[13]), "=r"(d[14]), "=r"(d[15])
                         : "r"(taddr) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) d[i] = 0u;
        }
        if (r < M) {
#pragma unroll
            for (int i = 0; i < 16; ++i)
                if (n0 + c0 + i < N) out[(size_t)r * N + n0 + c0 + i] = __uint_as_float(d[i]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(kTCols) : "memory");
}

// ---------------------------------------------------------------------------------------------------------------------------
// Packed operands (r2): both operands arrive ALREADY split into TF32 halves and laid out tile by tile in the canonical
// shared-memory order, so a stage is two bulk copies (cp.async.bulk, completion on the stage's mbarrier) and no register
// ever touches an operand.  The weights are packed once per set of weights (k_pack_rows: they are constant over the 150 calls
// of an objective), the activations are written in this layout by the kernel that produces them (k_template_match's
// epilogue).  Layout of a packed matrix of R rows x K columns with TR rows per tile (TR = 128 for x, 80 for W):
//   tile (rt, c) = rows rt*TR .. rt*TR+TR-1, columns 32c .. 32c+31, stored at float offset (rt * ceil(K/32) + c) * 2*TR*32 as
//   { hi[TR x 32], lo[TR x 32] }, element (r, k) of a half at ((r % TR) / 8) * 256 + ((k % 32) / 4) * 32 + (r % 8) * 4 + k % 4
//   (core matrices of 8 rows x 16 bytes, 128 bytes apart along K, 1 KB apart along the rows); rows >= R and columns >= K zero.
// Warp-specialised: lane 0 of warp 0 issues the copies, lane 0 of warp 1 the MMAs (12 per 32-k chunk, same order as
// k_linear_tc5: the two kernels give bit-identical results), all four warps read the accumulator out of tensor memory.
static_assert(kPackRowsX == kTBM && kPackRowsW == kTBN, "packed tiles are the CTA tiles");
// one thread per (row, group of four columns) of the padded matrix: 16 bytes of hi and 16 bytes of lo
__global__ void __launch_bounds__(256) k_pack_rows(const float* __restrict__ src, long long R, long long K, long long ld, int TR,
                                                   float* __restrict__ dst) {
    const long long nchunks = (K + 31) / 32, kg = nchunks * 8, Rp = (R + TR - 1) / TR * TR;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= Rp * kg) return;
    const long long r = idx / kg, k = (idx % kg) * 4;
    unsigned int hi[4], lo[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) split_tf32((r < R && k + i < K) ? __ldg(src + r * ld + k + i) : 0.f, hi[i], lo[i]);
    float* d = dst + packed_index(TR, r, k, nchunks);
    *reinterpret_cast<uint4*>(d) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(d + (size_t)TR * 32) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

template <int ST>
__global__ void __launch_bounds__(kTThreads) k_linear_tc5pk(const float* __restrict__ Ap, const float* __restrict__ Wp, int M, int N,
                                                            int nchunks_total, int chunks_per_split, float* __restrict__ partial,
                                                            int* __restrict__ err) {
    extern __shared__ __align__(1024) unsigned char tsm[];
    __shared__ unsigned long long full[ST], empty[ST];
    __shared__ unsigned int tmem_slot;
    const int n_tile = (int)blockIdx.x, m_tile = (int)blockIdx.y, split = (int)blockIdx.z;
    const int n0 = n_tile * kTBN, m0 = m_tile * kTBM;
    const int c0 = split * chunks_per_split;
    const int nch = max(0, min(nchunks_total, c0 + chunks_per_split) - c0);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < ST; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(kTCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned int tmem = tmem_slot;
    constexpr unsigned int idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned int)(kTBN >> 3) << 17) | ((unsigned int)(kTBM >> 4) << 24);
    bool ok = true;
    if (tid == 0) {                                                          // ---- producer: two bulk copies per stage
        const float* a_src = Ap + ((size_t)m_tile * nchunks_total + c0) * (size_t)(2 * kTBM * kTBK);
        const float* b_src = Wp + ((size_t)n_tile * nchunks_total + c0) * (size_t)(2 * kTBN * kTBK);
        for (int ch = 0; ch < nch && ok; ++ch) {
            const int s = ch % ST;
            if (ch >= ST) ok = mbar_wait_bounded(&empty[s], (unsigned int)((ch / ST - 1) & 1));   // released by the MMAs of chunk ch - ST
            if (!ok) break;
            unsigned char* st = tsm + (size_t)s * kTStage;
            mbar_expect_tx(&full[s], (unsigned int)kTStage);
            bulk_g2s(st, a_src + (size_t)ch * (2 * kTBM * kTBK), 2 * kTABytes, &full[s]);
            bulk_g2s(st + 2 * kTABytes, b_src + (size_t)ch * (2 * kTBN * kTBK), 2 * kTBBytes, &full[s]);
        }
    } else if (tid == 32) {                                                  // ---- MMA issuer
        for (int ch = 0; ch < nch && ok; ++ch) {
            const int s = ch % ST;
            ok = mbar_wait_bounded(&full[s], (unsigned int)((ch / ST) & 1));
            if (!ok) break;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const unsigned int a_hi = smem_u32(tsm + (size_t)s * kTStage), a_lo = a_hi + kTABytes, b_hi = a_hi + 2 * kTABytes, b_lo = b_hi + kTBBytes;
#pragma unroll
            for (int j = 0; j < kTBK / 8; ++j) {
                const unsigned int o = (unsigned int)j * 256u;
                umma_tf32(tmem, umma_desc(a_lo + o), umma_desc(b_hi + o), idesc, (ch > 0 || j > 0) ? 1u : 0u);
                umma_tf32(tmem, umma_desc(a_hi + o), umma_desc(b_lo + o), idesc, 1u);
                umma_tf32(tmem, umma_desc(a_hi + o), umma_desc(b_hi + o), idesc, 1u);
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.b64 [%0];" ::"r"(smem_u32(&empty[s])) : "memory");
        }
    }
    // The other 126 threads sleep at this barrier until every copy and MMA has been ISSUED.  (A parity wait only distinguishes
    // adjacent phases: a thread that skipped the loop must not poll the last commit's parity before the barrier has reached
    // that phase -- it would pass on an earlier phase of the same parity.)
    __syncwarp();
    __syncthreads();
    // ---- all MMAs done: the last commit arrives after every earlier one
    if (nch > 0) {
        const int last = nch - 1;
        ok = mbar_wait_bounded(&empty[last % ST], (unsigned int)((last / ST) & 1)) && ok;
    }
    ok = __syncthreads_and(ok ? 1 : 0) != 0;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (!ok && lane == 0 && err != nullptr) atomicExch(err, 1);
    float* out = partial + (size_t)split * M * N;
    const int r = m0 + warp * 32 + lane;
#pragma unroll
    for (int cc = 0; cc < kTBN; cc += 16) {
        unsigned int d[16];
        if (nch > 0 && ok) {
            const unsigned int taddr = tmem + ((unsigned int)(warp * 32) << 16) + (unsigned int)cc;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                         : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]), "=r"(d[4]), "=r"(d[5]), "=r"(d[6]), "=r"(d[7]),
                           "=r"(d[8]), "=r"(d[9]), "=r"(d[10]), "=r"(d[11]), "=r"(d[12]), "=r"(d[13]), "=r"(d[14]), "=r"(d[15])
                         : "r"(taddr) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) d[i] = 0u;
        }
        if (r < M) {
#pragma unroll
            for (int i = 0; i < 16; ++i)
                if (n0 + cc + i < N) out[(size_t)r * N + n0 + cc + i] = __uint_as_float(d[i]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(kTCols) : "memory");
}

static int linear_tc5_splits(long long M, long long N, long long K, int ctas_per_sm = APGS_TC5_CTAS) {
    const long long tiles = ((M + kTBM - 1) / kTBM) * ((N + kTBN - 1) / kTBN);
    long long s = ((long long)ctas_per_sm * sm_count()) / tiles;                 // one wave of the CTAs the shared-memory stages allow
    const long long smax = (K + 4 * kTBK - 1) / (4 * kTBK);
    if (s > smax) s = smax;
    if (s < 1) s = 1;
    if (s > 64) s = 64;
    return (int)s;
}

}  // namespace apgs

extern "C" size_t apgs_linear_tc5_workspace_bytes(int64_t M, int64_t N, int64_t K) {
    if (M <= 0 || N <= 0 || K <= 0) return 0;
    return (size_t)apgs::linear_tc5_splits(M, N, K) * (size_t)M * (size_t)N * sizeof(float) + 256;     // + the error word
}

extern "C" int apgs_linear_tc5_f32(const float* x, const float* W, const float* bias, int64_t M, int64_t K, int64_t N, int relu,
                                   float* out, void* workspace, size_t workspace_bytes, void* stream) {
    using namespace apgs;
    if (M <= 0 || N <= 0 || K <= 0 || M >= (1ll << 24) || N >= (1ll << 24) || K >= (1ll << 30))
        return fail(APGS_E_SHAPE, "linear_tc5: bad sizes M=%lld K=%lld N=%lld", (long long)M, (long long)K, (long long)N);
    if (!x || !W || !out) return fail(APGS_E_BADARG, "linear_tc5: NULL pointer");
    const int splits = linear_tc5_splits(M, N, K);
    const size_t need = (size_t)splits * (size_t)M * (size_t)N * sizeof(float) + 256;
    if (!workspace || workspace_bytes < need)
        return fail(APGS_E_WORKSPACE, "linear_tc5: workspace of %zu bytes required, got %zu", need, workspace_bytes);
    int kps = (int)((K + splits - 1) / splits);
    kps = (kps + kTBK - 1) / kTBK * kTBK;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem = (size_t)2 * kTStage + 1024;
    APGS_CUDA_OK(ensure_dynamic_smem(reinterpret_cast<const void*>(k_linear_tc5), smem), "smem attr");
    float* part = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(workspace) + 256);
    int* err = reinterpret_cast<int*>(workspace);
    APGS_CUDA_OK(cudaMemsetAsync(err, 0, sizeof(int), st), "memset");
    dim3 grid((unsigned int)((N + kTBN - 1) / kTBN), (unsigned int)((M + kTBM - 1) / kTBM), (unsigned int)splits);
    k_linear_tc5<<<grid, kTThreads, smem, st>>>(x, W, (int)M, (int)N, (int)K, kps, part, err);
    APGS_LAUNCH_OK("k_linear_tc5 launch");
    const long long MN = (long long)M * N;
    k_linear_reduce<<<(unsigned int)((MN + 255) / 256), 256, 0, st>>>(part, bias, MN, (int)N, splits, relu, out, err);
    APGS_LAUNCH_OK("k_linear_reduce launch");
    return 0;
}


// ---- packed operands (layout above k_pack_rows)
// 2 stages x 2 CTAs per SM (default): the split-K of apgs_linear_tc5_f32, hence bit-identical to it (0.037 ms per call at the C4
// shape).  APGS_TC5PK_STAGES=4: 4 stages x 1 CTA per SM, fewer and longer K slices -- 0.033 ms, but the longer fp32 accumulation
// chains in tensor memory cost accuracy (4.4e-6 of sum|x w| against 2.1e-6), so it stays a measurement knob.
static int tc5pk_stages() {
    const char* e = getenv("APGS_TC5PK_STAGES");
    return (e && atoi(e) == 4) ? 4 : 2;
}

extern "C" size_t apgs_linear_tc5_packed_bytes(int64_t rows, int64_t K, int which) {
    if (rows <= 0 || K <= 0 || (which != 0 && which != 1)) return 0;
    const int64_t TR = which == 0 ? apgs::kTBM : apgs::kTBN;
    return (size_t)((rows + TR - 1) / TR) * (size_t)((K + 31) / 32) * (size_t)(2 * TR * 32) * sizeof(float);
}

extern "C" int apgs_linear_tc5_pack_f32(const float* src, int64_t rows, int64_t K, int64_t ld, int which, float* dst, void* stream) {
    using namespace apgs;
    if (rows <= 0 || K <= 0 || ld < K || rows >= (1ll << 24) || K >= (1ll << 30) || (which != 0 && which != 1))
        return fail(APGS_E_SHAPE, "linear_tc5_pack: bad sizes rows=%lld K=%lld ld=%lld which=%d", (long long)rows, (long long)K,
                    (long long)ld, which);
    if (!src || !dst) return fail(APGS_E_BADARG, "linear_tc5_pack: NULL pointer");
    if (!aligned(dst, 16)) return fail(APGS_E_BADARG, "linear_tc5_pack: dst must be 16-byte aligned");
    const int TR = which == 0 ? kTBM : kTBN;
    const long long n = (rows + TR - 1) / TR * TR * ((K + 31) / 32) * 8;
    k_pack_rows<<<(unsigned int)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, rows, K, ld, TR, dst);
    APGS_LAUNCH_OK("k_pack_rows launch");
    return 0;
}

extern "C" size_t apgs_linear_tc5_packed_workspace_bytes(int64_t M, int64_t N, int64_t K) {
    if (M <= 0 || N <= 0 || K <= 0) return 0;
    return (size_t)apgs::linear_tc5_splits(M, N, K, tc5pk_stages() == 4 ? 1 : 2) * (size_t)M * (size_t)N * sizeof(float) + 256;
}

extern "C" int apgs_linear_tc5_packed_f32(const float* x_packed, const float* w_packed, const float* bias, int64_t M, int64_t K,
                                          int64_t N, int relu, float* out, void* workspace, size_t workspace_bytes, void* stream) {
    using namespace apgs;
    if (M <= 0 || N <= 0 || K <= 0 || M >= (1ll << 24) || N >= (1ll << 24) || K >= (1ll << 30))
        return fail(APGS_E_SHAPE, "linear_tc5_packed: bad sizes M=%lld K=%lld N=%lld", (long long)M, (long long)K, (long long)N);
    if (!x_packed || !w_packed || !out) return fail(APGS_E_BADARG, "linear_tc5_packed: NULL pointer");
    if (!aligned(x_packed, 16) || !aligned(w_packed, 16))
        return fail(APGS_E_BADARG, "linear_tc5_packed: packed operands must be 16-byte aligned");
    const int stages = tc5pk_stages();
    const int splits = linear_tc5_splits(M, N, K, stages == 4 ? 1 : 2);
    const size_t need = (size_t)splits * (size_t)M * (size_t)N * sizeof(float) + 256;
    if (!workspace || workspace_bytes < need)
        return fail(APGS_E_WORKSPACE, "linear_tc5_packed: workspace of %zu bytes required, got %zu", need, workspace_bytes);
    int kps = (int)((K + splits - 1) / splits);
    kps = (kps + kTBK - 1) / kTBK * kTBK;
    const int nchunks = (int)((K + kTBK - 1) / kTBK);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem = (size_t)stages * kTStage + 1024;
    float* part = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(workspace) + 256);
    int* err = reinterpret_cast<int*>(workspace);
    APGS_CUDA_OK(cudaMemsetAsync(err, 0, sizeof(int), st), "memset");
    dim3 grid((unsigned int)((N + kTBN - 1) / kTBN), (unsigned int)((M + kTBM - 1) / kTBM), (unsigned int)splits);
    if (stages == 4) {
        APGS_CUDA_OK(ensure_dynamic_smem(reinterpret_cast<const void*>(k_linear_tc5pk<4>), smem), "smem attr");
        k_linear_tc5pk<4><<<grid, kTThreads, smem, st>>>(x_packed, w_packed, (int)M, (int)N, nchunks, kps / kTBK, part, err);
    } else {
        APGS_CUDA_OK(ensure_dynamic_smem(reinterpret_cast<const void*>(k_linear_tc5pk<2>), smem), "smem attr");
        k_linear_tc5pk<2><<<grid, kTThreads, smem, st>>>(x_packed, w_packed, (int)M, (int)N, nchunks, kps / kTBK, part, err);
    }
    APGS_LAUNCH_OK("k_linear_tc5pk launch");
    const long long MN = (long long)M * N;
    k_linear_reduce<<<(unsigned int)((MN + 255) / 256), 256, 0, st>>>(part, bias, MN, (int)N, splits, relu, out, err);
    APGS_LAUNCH_OK("k_linear_reduce launch");
    return 0;
}
