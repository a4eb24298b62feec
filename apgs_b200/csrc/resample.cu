// Resampling kernels for sm_100a: weight normalisation / ESS / log-sum-exp (W1), CDF + ancestor
// search (R2 systematic, R3 multinomial), ancestor gather (R4) and the fused entry (R5).
//
// Replaces apgs/resampler.py:22-54 and the per-model resample_variables (see include/apgs_b200.h).
// All HBM-bound: no tensor cores.  Layout facts used:
//   * log_w is (S,B) with B innermost, so a warp reading 32 consecutive instances of one particle
//     row is one 128 B transaction;
//   * a latent is (S,B,inner): the output of the gather is a dense stream, the input is B chunks of
//     `inner` bytes per output row.
#include "resample.cuh"
#include "tile.cuh"

namespace apgs {

// =============================================================================================
// K_anc: one CTA owns TB adjacent instances (columns) and ALL S particles of them.
//   phase 1  load the column tile (coalesced), column max
//   phase 2  e = exp(x - max), column sum
//   phase 3  w = e / sum  (optionally stored), sum of squares -> ess, lse
//   phase 4  CDF over the particle axis (EXACT: one sequential chain per column; BLOCKED: strips)
//   phase 5  normalise by the last entry
//   phase 6  ancestor = lower_bound(ncdf, position)  (binary search, coalesced index stores)
// The tile lives in shared memory (S*TB*4 B) or, for very long particle axes, in a global
// workspace with the same (row, column) addressing.
// =============================================================================================
template <int TB, typename T, typename Op>
__device__ __forceinline__ T col_reduce(T v, Op op, T* scratch, int col) {
    // reduce over all threads of the CTA that share `col` (= tid % TB)
    if (TB < 32) {
#pragma unroll
        for (int o = TB; o < 32; o <<= 1) v = op(v, __shfl_xor_sync(0xffffffffu, v, o));
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    if (lane < TB) scratch[warp * TB + lane] = v;
    __syncthreads();
    T r = scratch[col];
    for (int w = 1; w < nwarp; ++w) r = op(r, scratch[w * TB + col]);
    __syncthreads();
    return r;
}

// Tile accessors: shared memory (compile-time stride TB, LDS/STS with immediate offsets) or the
// global workspace (stride B).
template <int TB, bool SMEM>
struct Tile {
    float* p;
    long long gs;
    __device__ __forceinline__ float ld(int i) const { return SMEM ? p[i * TB] : p[(long long)i * gs]; }
    __device__ __forceinline__ void st(int i, float v) const {
        if (SMEM) p[i * TB] = v; else p[(long long)i * gs] = v;
    }
};

// One sequential chain per column (bit-exact with ATen's CPU loops).  Only the adds are dependent:
// values are fetched / converted / stored in register batches of CH so the chain costs one DADD
// (or FADD) latency per particle instead of a shared-memory round trip.
template <int TB, bool SMEM, typename ACC>
__device__ __forceinline__ void cdf_chain(const Tile<TB, SMEM>& t, int S) {
    constexpr int CH = 16;
    ACC acc = (ACC)0;
    const int Sfull = S - (S % CH);
    float v[CH];
    if (Sfull > 0) {
#pragma unroll
        for (int k = 0; k < CH; ++k) v[k] = t.ld(k);
    }
    for (int i0 = 0; i0 < Sfull; i0 += CH) {
        float nx[CH];
        if (i0 + CH < Sfull) {
#pragma unroll
            for (int k = 0; k < CH; ++k) nx[k] = t.ld(i0 + CH + k);
        }
        float c[CH];
#pragma unroll
        for (int k = 0; k < CH; ++k) {
            if (sizeof(ACC) == 8) acc = (ACC)((double)acc + (double)v[k]);
            else acc = (ACC)__fadd_rn((float)acc, v[k]);
            c[k] = (float)acc;
        }
#pragma unroll
        for (int k = 0; k < CH; ++k) t.st(i0 + k, c[k]);
#pragma unroll
        for (int k = 0; k < CH; ++k) v[k] = nx[k];
    }
    for (int i = Sfull; i < S; ++i) {
        if (sizeof(ACC) == 8) acc = (ACC)((double)acc + (double)t.ld(i));
        else acc = (ACC)__fadd_rn((float)acc, t.ld(i));
        t.st(i, (float)acc);
    }
}

// Exact systematic chain on an fp64 shared tile: the serial path is LDS.64 -> DADD -> STS.64 only
// (fp32<->fp64 conversions are quarter-rate and were 2/3 of the chain when done inline).
template <int TB>
__device__ __forceinline__ void cdf_chain_f64(double* t, int S) {
    constexpr int CH = 8;
    double acc = 0.0;
    const int Sfull = S - (S % CH);
    double v[CH];
    if (Sfull > 0) {
#pragma unroll
        for (int k = 0; k < CH; ++k) v[k] = t[k * TB];
    }
    for (int i0 = 0; i0 < Sfull; i0 += CH) {
        double nx[CH];
        if (i0 + CH < Sfull) {
#pragma unroll
            for (int k = 0; k < CH; ++k) nx[k] = t[(i0 + CH + k) * TB];
        }
#pragma unroll
        for (int k = 0; k < CH; ++k) { acc += v[k]; v[k] = acc; }
#pragma unroll
        for (int k = 0; k < CH; ++k) t[(i0 + k) * TB] = v[k];
#pragma unroll
        for (int k = 0; k < CH; ++k) v[k] = nx[k];
    }
    for (int i = Sfull; i < S; ++i) { acc += t[i * TB]; t[i * TB] = acc; }
}

template <int TB, bool SMEM>
__global__ void __launch_bounds__(1024, 1) k_ancestors(AncArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* dscr = reinterpret_cast<double*>(smem_raw);                 // blockDim.x doubles
    float* stile = reinterpret_cast<float*>(smem_raw + (size_t)blockDim.x * sizeof(double));

    const int tid = threadIdx.x, nthr = blockDim.x;
    const int col = tid % TB, rg = tid / TB, R = nthr / TB;
    const long long b0 = (long long)blockIdx.x * TB;
    const long long b = b0 + col;
    const bool colok = b < a.B;
    const int S = (int)a.S;
    const long long B = a.B;

    // fp64 CDF tile (only when a.dbl): S*TB doubles after the fp32 tile, 8-byte aligned
    double* dt = reinterpret_cast<double*>(stile + (((size_t)S * TB + 1) & ~(size_t)1)) + col;
    const bool dbl = SMEM && a.dbl;

    Tile<TB, SMEM> tile;
    if (SMEM) { tile.p = stile + col; tile.gs = TB; }
    else      { tile.p = a.ws + b; tile.gs = B; }
    // in global mode out-of-range columns never touch memory
    const bool touch = SMEM || colok;

    float cmax = 0.f, csum = 1.f;
    if (a.w_norm == nullptr) {
        // ---- phase 1
        float m = -INFINITY;
        for (int i = rg; i < S; i += R) {
            float x = colok ? __ldg(a.log_w + (long long)i * B + b) : 0.f;
            if (touch) tile.st(i, x);
            m = fmaxf(m, x);
        }
        cmax = col_reduce<TB>(m, [](float p, float q) { return fmaxf(p, q); },
                              reinterpret_cast<float*>(dscr), col);
        // ---- phase 2
        float ps = 0.f;
        if (touch)
            for (int i = rg; i < S; i += R) {
                float e = expf(tile.ld(i) - cmax);
                tile.st(i, e);
                ps += e;
            }
        double dsum = col_reduce<TB>((double)ps, [](double p, double q) { return p + q; }, dscr, col);
        csum = (float)dsum;
        // ---- phase 3
        float pq = 0.f;
        if (touch)
            for (int i = rg; i < S; i += R) {
                float w = __fdiv_rn(tile.ld(i), csum);
                if (dbl) dt[i * TB] = (double)w; else tile.st(i, w);
                if (a.w_out != nullptr && colok) a.w_out[(long long)i * B + b] = w;
                pq = fmaf(w, w, pq);
            }
        if (a.ess != nullptr || a.lse != nullptr) {
            double dsq = col_reduce<TB>((double)pq, [](double p, double q) { return p + q; }, dscr, col);
            if (rg == 0 && colok) {
                if (a.ess != nullptr) a.ess[b] = __fdiv_rn(1.f, (float)dsq);
                if (a.lse != nullptr) a.lse[b] = cmax + logf(csum);
            }
        }
        if (a.strategy == APGS_MULTINOMIAL) {
            // Categorical(probs) renormalises once more before torch.multinomial
            // (torch/distributions/categorical.py:70, reached from apgs/resampler.py:41)
            float pw = 0.f;
            if (touch)
                for (int i = rg; i < S; i += R) pw += tile.ld(i);
            const float sw = (float)col_reduce<TB>((double)pw, [](double p, double q) { return p + q; }, dscr, col);
            if (touch)
                for (int i = rg; i < S; i += R) tile.st(i, __fdiv_rn(tile.ld(i), sw));
        }
    } else {
        for (int i = rg; i < S; i += R) {
            float w = colok ? __ldg(a.w_norm + (long long)i * B + b) : 0.f;
            if (dbl) dt[i * TB] = (double)w; else if (touch) tile.st(i, w);
        }
    }
    if (a.strategy < 0) return;
    __syncthreads();

    // ---- phase 4: CDF
    if (!a.blocked) {
        if (rg == 0 && touch) {
            if (dbl)                                cdf_chain_f64<TB>(dt, S);               // ATen CPU cumsum<float>
            else if (a.strategy == APGS_SYSTEMATIC) cdf_chain<TB, SMEM, double>(tile, S);
            else                               cdf_chain<TB, SMEM, float>(tile, S);    // ATen CPU multinomial
        }
    } else {
        const int L = (S + R - 1) / R;
        const int i0 = min(rg * L, S), i1 = min(i0 + L, S);
        double loc = 0.0;
        if (touch)
            for (int i = i0; i < i1; ++i) loc += (double)tile.ld(i);
        dscr[rg * TB + col] = loc;
        __syncthreads();
        if (rg == 0) {
            double run = 0.0;
            for (int r = 0; r < R; ++r) {
                double v = dscr[r * TB + col];
                dscr[r * TB + col] = run;
                run += v;
            }
        }
        __syncthreads();
        double acc = dscr[rg * TB + col];
        if (touch)
            for (int i = i0; i < i1; ++i) {
                acc += (double)tile.ld(i);
                tile.st(i, (float)acc);
            }
    }
    __syncthreads();

    // ---- phase 5: normalise by the last CDF entry (the row max of a monotone CDF; resampler.py:32-33)
    if (dbl) {
        const float norm = (float)dt[(S - 1) * TB];
        for (int i = rg; i < S; i += R) tile.st(i, __fdiv_rn((float)dt[i * TB], norm));
    } else {
        const float norm = touch ? tile.ld(S - 1) : 1.f;
        __syncthreads();
        if (touch)
            for (int i = rg; i < S; i += R) tile.st(i, __fdiv_rn(tile.ld(i), norm));
    }
    __syncthreads();

    // ---- phase 6: ancestors
    if (!colok) return;
    const float fS = (float)S;
    if (a.strategy == APGS_SYSTEMATIC) {
        // Positions (u + s)/S increase with s and so do their ancestors: a thread owns a STRIP of consecutive
        // particles, locates the first one by binary search and then gallops forward from the previous answer
        // (typically 1-2 probes per particle instead of log2 S; identical result: the lower bound of pos in the CDF).
        const float u = __ldg(a.u_sys + b);
        const int per = (S + R - 1) / R;
        const int i0 = min(rg * per, S), i1 = min(i0 + per, S);
        int anc = 0;
        for (int i = i0; i < i1; ++i) {
            const float pos = __fdiv_rn(__fadd_rn(u, (float)i), fS);          // resampler.py:28
            int lo, hi;
            if (i == i0) {
                lo = 0; hi = S;
            } else if (anc < S && tile.ld(anc) < pos) {                       // advance: cdf[anc] < pos
                int step = 1;
                lo = anc;
                while (lo + step < S && tile.ld(lo + step) < pos) { lo += step; step <<= 1; }
                hi = min(lo + step, S);
                lo = lo + 1;
            } else {
                lo = hi = anc;
            }
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (tile.ld(mid) < pos) lo = mid + 1; else hi = mid;
            }
            anc = lo;
            const int out = min(anc, S - 1);                                  // memory safety on NaN / degenerate weights
            if (a.anc32 != nullptr) a.anc32[(long long)i * B + b] = out;
            if (a.anc64 != nullptr) a.anc64[(long long)i * B + b] = out;
        }
        return;
    }
    for (int i = rg; i < S; i += R) {
        int lo = 0, hi = S;
        const double ud = __ldg(a.u_mult + (long long)i * a.us + b * a.ub);
        while (lo < hi) {
            int mid = (lo + hi) >> 1;
            if ((double)tile.ld(mid) < ud) lo = mid + 1; else hi = mid;
        }
        if (lo > S - 1) lo = S - 1;                        // memory safety on NaN / degenerate weights
        if (a.anc32 != nullptr) a.anc32[(long long)i * B + b] = lo;
        if (a.anc64 != nullptr) a.anc64[(long long)i * B + b] = lo;
    }
}

static const size_t kSmemTileBudget = 200 * 1024;

struct AncPlan { int TB; int nthr; int tile_in_smem; int dbl; size_t smem; int grid; };

#ifndef APGS_ANC_MIN_TB
#define APGS_ANC_MIN_TB 4
#endif
static AncPlan plan_ancestors(long long S, long long B, bool want_dbl) {
    AncPlan p;
    int TB = 32;
    const size_t per = want_dbl ? 12 : 4;       // bytes of shared memory per (particle, column)
    while (TB > 1 && ((size_t)S * TB + 2) * per > kSmemTileBudget) TB >>= 1;
    p.tile_in_smem = (((size_t)S * TB + 2) * per <= kSmemTileBudget) ? 1 : 0;
    p.dbl = (want_dbl && p.tile_in_smem) ? 1 : 0;
    if (p.tile_in_smem) {
        // spread over the 148 SMs: shrink the column tile while the grid is under ~2 CTAs/SM
        while (TB > APGS_ANC_MIN_TB && (B + TB - 1) / TB < 2 * sm_count()) TB >>= 1;
    } else {
        TB = 8;
    }
    while (TB > 1 && TB / 2 >= B) TB >>= 1;               // do not waste lanes on tiny B
    long long want = (S * TB + 7) / 8;
    int nthr = 128;
    while (nthr < 1024 && nthr < want) nthr <<= 1;
    if (!p.tile_in_smem) nthr = 1024;
    p.TB = TB;
    p.nthr = nthr;
    p.smem = (size_t)nthr * sizeof(double) + (p.tile_in_smem ? ((size_t)S * TB + 2) * (p.dbl ? 12 : 4) : 0);
    p.grid = (int)((B + TB - 1) / TB);
    return p;
}

template <int TB>
static int launch_anc_tb(const AncArgs& a, const AncPlan& p, cudaStream_t st) {
    if (p.tile_in_smem) {
        APGS_CUDA_OK(ensure_dynamic_smem(reinterpret_cast<const void*>(k_ancestors<TB, true>), p.smem),
                     "cudaFuncSetAttribute(k_ancestors)");
        k_ancestors<TB, true><<<p.grid, p.nthr, p.smem, st>>>(a);
    } else {
        k_ancestors<TB, false><<<p.grid, p.nthr, p.smem, st>>>(a);
    }
    APGS_LAUNCH_OK("k_ancestors launch");
    return 0;
}

static int launch_ancestors(AncArgs a, void* workspace, size_t workspace_bytes, cudaStream_t st) {
    AncPlan p = plan_ancestors(a.S, a.B, a.strategy == APGS_SYSTEMATIC && !a.blocked);
    a.tile_in_smem = p.tile_in_smem;
    a.dbl = p.dbl;
    // particle axes beyond a shared-memory tile: the split-S pipeline (resample_large.cu) unless the caller insists on
    // the exact sequential chain
    if (!p.tile_in_smem && (a.blocked || a.strategy < 0)) return launch_ancestors_large(a, workspace, workspace_bytes, st);
    if (!p.tile_in_smem) {
        size_t need = (size_t)a.S * a.B * sizeof(float);
        if (workspace == nullptr || workspace_bytes < need)
            return fail(APGS_E_WORKSPACE, "ancestors: workspace of %zu bytes required (S=%lld B=%lld), got %zu",
                        need, a.S, a.B, workspace_bytes);
        a.ws = reinterpret_cast<float*>(workspace);
    }
    switch (p.TB) {
        case 32: return launch_anc_tb<32>(a, p, st);
        case 16: return launch_anc_tb<16>(a, p, st);
        case 8:  return launch_anc_tb<8>(a, p, st);
        case 4:  return launch_anc_tb<4>(a, p, st);
        case 2:  return launch_anc_tb<2>(a, p, st);
        default: return launch_anc_tb<1>(a, p, st);
    }
}

// =============================================================================================
// K_gather: dst_l[s,b,:] = src_l[anc[s,b], b, :] for up to 8 latents in one launch.
// Output-stationary: consecutive threads own consecutive 16/8/4-byte vectors of the dense output,
// so stores are perfectly coalesced and each source chunk is read with full-width vectors.  The
// int64 gather index the reference materialises (apgs/resampler.py:48,53) never exists.
// Variants measured on B200 at the GMM shape (ms for all-distinct / sigma=2 / degenerate weights):
//   this flat output-stationary mapping, 8 x 16 B per thread ......... 0.278 / 0.245 / 0.194
//   row-stationary (G lanes per row, 4 rows in flight per thread) .... 0.298 / 0.232 / 0.173
//   warp-span (32 rows per warp, source offsets by shuffle) .......... 0.320 / 0.258 / 0.227
//   cp.async.bulk (TMA) row loads into a shared ring + bulk store ..... 0.417 (copy engine ~120 cycles / sub-KB request)
// The flat mapping wins where it matters (every row read from HBM), so it is the one kept.
// =============================================================================================
struct GLat {
    const char* src;
    char* dst;
    unsigned int vlog;         // log2(vector bytes)
    unsigned int vpr;          // vectors per (s,b) row
    unsigned int Bv;           // vectors per particle row  (= B * vpr)
    unsigned int rows_per_blk; // >1: block covers whole rows;  1: row split into chunks
    unsigned int chunks;       // chunks per row when rows_per_blk == 1
    unsigned int blk_begin;    // first block of this latent
    FastDiv d_vpr, d_Bv, d_chunks;
    const int* first;          // optional (S,B) int32: the latent still sits in the order BEFORE an earlier resampling step
                               // whose ancestors these are -- the row read is src[first[anc[s,b], b], b]
};
struct GatherArgs {
    GLat lat[APGS_MAX_LATENTS];
    int n;
    const int* anc32;
    const long long* anc64;
    long long S, B;
};

constexpr int kGThreads = 256;
constexpr int kGUnroll = 8;
constexpr int kGTile = kGThreads * kGUnroll;   // vectors per block

template <typename V>
__device__ __forceinline__ void gather_block(const GLat& L, const GatherArgs& g, unsigned int lb) {
    const V* __restrict__ src = reinterpret_cast<const V*>(L.src);
    V* __restrict__ dst = reinterpret_cast<V*>(L.dst);
    long long s0;
    unsigned int t0, tend;          // block covers row-local vector range [t0, tend) starting at row s0
    unsigned int nrows;
    if (L.rows_per_blk > 1) {
        s0 = (long long)lb * L.rows_per_blk;
        long long left = g.S - s0;
        nrows = left < (long long)L.rows_per_blk ? (unsigned int)left : L.rows_per_blk;
        t0 = 0;
        tend = nrows * L.Bv;
    } else {
        unsigned int row = fastdiv(lb, L.d_chunks);
        unsigned int ch = lb - row * L.chunks;
        s0 = row;
        t0 = ch * kGTile;
        tend = min(t0 + (unsigned int)kGTile, L.Bv);
        nrows = 1;
    }
    const long long Bv = L.Bv;
    V val[kGUnroll];
    long long so[kGUnroll];
#pragma unroll
    for (int k = 0; k < kGUnroll; ++k) {
        unsigned int t = t0 + k * kGThreads + threadIdx.x;
        so[k] = -1;
        if (t < tend) {
            unsigned int ds = (L.rows_per_blk > 1) ? fastdiv(t, L.d_Bv) : 0u;
            unsigned int v = t - ds * L.Bv;                   // vector within the particle row
            unsigned int bcol = fastdiv(v, L.d_vpr);
            long long s = s0 + ds;
            long long anc = g.anc32 ? (long long)__ldg(g.anc32 + s * g.B + bcol)
                                    : __ldg(g.anc64 + s * g.B + bcol);
            if (L.first != nullptr) anc = (long long)__ldg(L.first + anc * g.B + bcol);
            so[k] = anc * Bv + v;
        }
    }
#pragma unroll
    for (int k = 0; k < kGUnroll; ++k)
        if (so[k] >= 0) val[k] = __ldg(src + so[k]);
#pragma unroll
    for (int k = 0; k < kGUnroll; ++k) {
        unsigned int t = t0 + k * kGThreads + threadIdx.x;
        if (so[k] >= 0) __stcs(dst + (s0 * Bv + t), val[k]);
    }
}

__global__ void __launch_bounds__(kGThreads) k_gather(const __grid_constant__ GatherArgs g) {
    int l = 0;
#pragma unroll 1
    for (int i = 1; i < g.n; ++i)
        if (blockIdx.x >= g.lat[i].blk_begin) l = i;
    const GLat& L = g.lat[l];
    const unsigned int lb = blockIdx.x - L.blk_begin;
    switch (L.vlog) {
        case 4: gather_block<uint4>(L, g, lb); break;
        case 3: gather_block<uint2>(L, g, lb); break;
        case 2: gather_block<unsigned int>(L, g, lb); break;
        case 1: gather_block<unsigned short>(L, g, lb); break;
        default: gather_block<unsigned char>(L, g, lb); break;
    }
}

static int launch_gather(const int* anc32, const long long* anc64, long long S, long long B,
                         const apgs_latent_t* lat, int n, cudaStream_t st, const int32_t* const* first = nullptr) {
    if (n < 0 || n > APGS_MAX_LATENTS) return fail(APGS_E_BADARG, "gather: n_latents=%d out of [0,%d]", n, APGS_MAX_LATENTS);
    if ((anc32 == nullptr) == (anc64 == nullptr)) return fail(APGS_E_BADARG, "gather: exactly one of anc32/anc64 must be given");
    if (n == 0 || S == 0 || B == 0) return 0;
    GatherArgs g;
    g.n = n; g.anc32 = anc32; g.anc64 = anc64; g.S = S; g.B = B;
    unsigned long long blocks = 0;
    for (int i = 0; i < n; ++i) {
        const apgs_latent_t& a = lat[i];
        if (a.src == nullptr || a.dst == nullptr || a.inner_bytes <= 0)
            return fail(APGS_E_BADARG, "gather: latent %d has null pointer or inner_bytes<=0", i);
        unsigned int vlog = 4;
        while (vlog > 0 && ((a.inner_bytes & ((1ll << vlog) - 1)) != 0 ||
                            !aligned(a.src, 1u << vlog) || !aligned(a.dst, 1u << vlog))) --vlog;
        long long vpr = a.inner_bytes >> vlog;
        long long Bv = B * vpr;
        if (Bv >= (1ll << 31) - kGTile) return fail(APGS_E_SHAPE, "gather: latent %d row of %lld vectors too long", i, Bv);
        GLat& L = g.lat[i];
        L.src = reinterpret_cast<const char*>(a.src);
        L.dst = reinterpret_cast<char*>(a.dst);
        L.first = first ? first[i] : nullptr;
        L.vlog = vlog; L.vpr = (unsigned int)vpr; L.Bv = (unsigned int)Bv;
        L.d_vpr = make_fastdiv(L.vpr); L.d_Bv = make_fastdiv(L.Bv);
        unsigned long long nb;
        if (Bv * 2 <= kGTile) {
            L.rows_per_blk = (unsigned int)(kGTile / Bv);
            L.chunks = 1;
            nb = (unsigned long long)((S + L.rows_per_blk - 1) / L.rows_per_blk);
        } else {
            L.rows_per_blk = 1;
            L.chunks = (unsigned int)((Bv + kGTile - 1) / kGTile);
            nb = (unsigned long long)S * L.chunks;
        }
        L.d_chunks = make_fastdiv(L.chunks);
        if (blocks + nb >= (1ull << 31)) return fail(APGS_E_SHAPE, "gather: grid too large");
        L.blk_begin = (unsigned int)blocks;
        blocks += nb;
    }
    k_gather<<<(unsigned int)blocks, kGThreads, 0, st>>>(g);
    APGS_LAUNCH_OK("k_gather launch");
    return 0;
}

// =============================================================================================
// W2: softmax over S for every (b, n) column of log_w_local (S,B,N) + sum over n.
// =============================================================================================
__global__ void __launch_bounds__(256) k_local_sum(const float* __restrict__ lwl, long long rows, int N,
                                                   float* __restrict__ out) {
    // one warp per (s,b) row of N contiguous floats
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    double acc = 0.0;
    for (int n = lane; n < N; n += 32) acc += (double)__ldg(lwl + row * N + n);
    acc = warp_sum(acc);
    if (lane == 0) out[row] = (float)acc;
}

__global__ void __launch_bounds__(256) k_local_softmax(const float* __restrict__ lwl, long long S, long long C,
                                                       float* __restrict__ w) {
    // C = B*N columns, softmax over S; thread per column, lanes along the contiguous axis
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float m = -INFINITY;
    for (long long s = 0; s < S; ++s) m = fmaxf(m, __ldg(lwl + s * C + c));
    double acc = 0.0;                                               // S is large here (beyond a shared-memory tile)
    for (long long s = 0; s < S; ++s) acc += (double)expf(__ldg(lwl + s * C + c) - m);
    const float sum = (float)acc;
    for (long long s = 0; s < S; ++s) w[s * C + c] = __fdiv_rn(expf(__ldg(lwl + s * C + c) - m), sum);
}

// Single-pass tile variant: a CTA owns TC = 16 adjacent columns and ALL S rows of them in shared memory -- one read and
// one write of the (S,B,N) tensor instead of three reads and one write.  256 threads = 64 row groups x 4 column quads:
// a thread moves 16 bytes per row (cp.async straight into the tile, S/64 requests in flight per thread, no registers),
// the column maxima / sums meet through warp shuffles (8 row groups per warp) and an 8 x 16 shared array, exp is evaluated
// once per element (kept in the tile) and normalised with the same IEEE division; 16-byte streaming stores.
constexpr int kLsTC = 16, kLsThreads = 256;

__device__ __forceinline__ float4 ls_reduce4(float4 v, float (*s_red)[kLsTC + 4], bool is_max) {
    // across the 8 row groups of a warp (lanes differ in bits 2..4), then across the 8 warps
#pragma unroll
    for (int o = 4; o < 32; o <<= 1) {
        const float a = __shfl_xor_sync(0xffffffffu, v.x, o), b = __shfl_xor_sync(0xffffffffu, v.y, o);
        const float c = __shfl_xor_sync(0xffffffffu, v.z, o), d = __shfl_xor_sync(0xffffffffu, v.w, o);
        if (is_max) { v.x = fmaxf(v.x, a); v.y = fmaxf(v.y, b); v.z = fmaxf(v.z, c); v.w = fmaxf(v.w, d); }
        else { v.x += a; v.y += b; v.z += c; v.w += d; }
    }
    const int q = threadIdx.x & 3, warp = threadIdx.x >> 5;
    __syncthreads();                                           // previous use of s_red is over
    if ((threadIdx.x & 31) < 4) *reinterpret_cast<float4*>(&s_red[warp][q * 4]) = v;
    __syncthreads();
    float4 r = *reinterpret_cast<const float4*>(&s_red[0][q * 4]);
#pragma unroll
    for (int i = 1; i < kLsThreads / 32; ++i) {
        const float4 t = *reinterpret_cast<const float4*>(&s_red[i][q * 4]);
        if (is_max) { r.x = fmaxf(r.x, t.x); r.y = fmaxf(r.y, t.y); r.z = fmaxf(r.z, t.z); r.w = fmaxf(r.w, t.w); }
        else { r.x += t.x; r.y += t.y; r.z += t.z; r.w += t.w; }
    }
    return r;
}

template <bool VEC>
__global__ void __launch_bounds__(kLsThreads) k_local_softmax_tile(const float* __restrict__ lwl, int S, long long C,
                                                                   float* __restrict__ w) {
    extern __shared__ __align__(16) float s_tile[];                 // [S][TC]
    __shared__ __align__(16) float s_red[kLsThreads / 32][kLsTC + 4];
    const int q = threadIdx.x & 3, rg = threadIdx.x >> 2;           // column quad, row group (0..63)
    constexpr int RG = kLsThreads / 4;
    const long long col = (long long)blockIdx.x * kLsTC + q * 4;
    const bool live = col < C;                                      // VEC: C % 4 == 0, so a quad is all in or all out
    if (VEC) {
        if (live)
            for (int s = rg; s < S; s += RG)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(s_tile + s * kLsTC + q * 4)),
                             "l"(lwl + (long long)s * C + col) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    } else {
        for (int s = rg; s < S; s += RG)
#pragma unroll
            for (int i = 0; i < 4; ++i) s_tile[s * kLsTC + q * 4 + i] = (col + i < C) ? __ldg(lwl + (long long)s * C + col + i) : -INFINITY;
    }
    // each thread only touches the tile elements it loaded itself: no barrier needed before the passes below
    float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    if (live || !VEC)
        for (int s = rg; s < S; s += RG) {
            const float4 v = *reinterpret_cast<const float4*>(s_tile + s * kLsTC + q * 4);
            m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
        }
    m = ls_reduce4(m, s_red, true);
    float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
    if (live || !VEC)
        for (int s = rg; s < S; s += RG) {
            float4* t = reinterpret_cast<float4*>(s_tile + s * kLsTC + q * 4);
            float4 v = *t;
            v.x = expf(v.x - m.x); v.y = expf(v.y - m.y); v.z = expf(v.z - m.z); v.w = expf(v.w - m.w);
            *t = v;
            sum.x += v.x; sum.y += v.y; sum.z += v.z; sum.w += v.w;
        }
    sum = ls_reduce4(sum, s_red, false);
    if (VEC) {
        if (live)
            for (int s = rg; s < S; s += RG) {
                float4 v = *reinterpret_cast<const float4*>(s_tile + s * kLsTC + q * 4);
                v.x = __fdiv_rn(v.x, sum.x); v.y = __fdiv_rn(v.y, sum.y); v.z = __fdiv_rn(v.z, sum.z); v.w = __fdiv_rn(v.w, sum.w);
                __stcs(reinterpret_cast<float4*>(w + (long long)s * C + col), v);
            }
    } else {
        const float sm[4] = {sum.x, sum.y, sum.z, sum.w};
        for (int s = rg; s < S; s += RG)
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (col + i < C) w[(long long)s * C + col + i] = __fdiv_rn(s_tile[s * kLsTC + q * 4 + i], sm[i]);
    }
}

}  // namespace apgs

using namespace apgs;

// =============================================================================================
// C ABI
// =============================================================================================
static int check_sb(long long S, long long B, const char* who) {
    if (S <= 0 || B <= 0) return fail(APGS_E_SHAPE, "%s: S=%lld B=%lld must be positive", who, S, B);
    if (S >= (1ll << 31) || B >= (1ll << 31)) return fail(APGS_E_SHAPE, "%s: S=%lld B=%lld too large", who, S, B);
    return 0;
}

extern "C" size_t apgs_ancestors_workspace_bytes(int64_t S, int64_t B) {
    if (S <= 0 || B <= 0) return 0;
    if (((size_t)S + 2) * 12 <= kSmemTileBudget) return 0;      // CDF tile fits in shared memory
    return ancestors_large_workspace_bytes(S, B);               // >= S*B*4, which the exact global-tile variant needs
}

extern "C" size_t apgs_resample_workspace_bytes(int64_t S, int64_t B) {
    if (S <= 0 || B <= 0) return 0;
    size_t idx = ((size_t)S * (size_t)B * sizeof(int32_t) + 255) & ~(size_t)255;
    return idx + apgs_ancestors_workspace_bytes(S, B);
}

extern "C" int apgs_weight_stats_f32(const float* log_w, int64_t S, int64_t B, float* w_out, float* ess,
                                     float* lse, void* workspace, size_t workspace_bytes, void* stream) {
    if (int e = check_sb(S, B, "weight_stats")) return e;
    if (log_w == nullptr) return fail(APGS_E_BADARG, "weight_stats: log_w is NULL");
    AncArgs a = {};
    a.log_w = log_w; a.S = S; a.B = B; a.strategy = -1;
    a.w_out = w_out; a.ess = ess; a.lse = lse;
    return launch_ancestors(a, workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int apgs_sample_ancestors_f32(const float* log_w, const float* w_norm, int64_t S, int64_t B,
                                         int strategy, const void* uniforms, int64_t u_stride_s,
                                         int64_t u_stride_b, int cdf_mode, int64_t* anc64, int32_t* anc32,
                                         float* w_out, float* ess, float* lse, void* workspace,
                                         size_t workspace_bytes, void* stream) {
    if (int e = check_sb(S, B, "sample_ancestors")) return e;
    if ((log_w == nullptr) == (w_norm == nullptr))
        return fail(APGS_E_BADARG, "sample_ancestors: exactly one of log_w / w_norm must be given");
    if (strategy != APGS_SYSTEMATIC && strategy != APGS_MULTINOMIAL)
        return fail(APGS_E_BADARG, "sample_ancestors: unknown strategy %d", strategy);
    if (uniforms == nullptr) return fail(APGS_E_BADARG, "sample_ancestors: uniforms is NULL");
    if (cdf_mode < APGS_CDF_AUTO || cdf_mode > APGS_CDF_BLOCKED)
        return fail(APGS_E_BADARG, "sample_ancestors: unknown cdf_mode %d", cdf_mode);
    if (anc64 == nullptr && anc32 == nullptr)
        return fail(APGS_E_BADARG, "sample_ancestors: no ancestor output given");
    AncArgs a = {};
    a.log_w = log_w; a.w_norm = w_norm; a.S = S; a.B = B; a.strategy = strategy;
    if (strategy == APGS_SYSTEMATIC) a.u_sys = reinterpret_cast<const float*>(uniforms);
    else { a.u_mult = reinterpret_cast<const double*>(uniforms); a.us = u_stride_s; a.ub = u_stride_b; }
    a.blocked = (cdf_mode == APGS_CDF_BLOCKED) || (cdf_mode == APGS_CDF_AUTO && S > 8192);
    a.anc64 = reinterpret_cast<long long*>(anc64); a.anc32 = anc32;
    a.w_out = w_out; a.ess = ess; a.lse = lse;
    return launch_ancestors(a, workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int apgs_gather_ancestors(const int32_t* anc32, const int64_t* anc64, int64_t S, int64_t B,
                                     const apgs_latent_t* latents, int n_latents, void* stream) {
    if (int e = check_sb(S, B, "gather_ancestors")) return e;
    if (latents == nullptr && n_latents > 0) return fail(APGS_E_BADARG, "gather_ancestors: latents is NULL");
    return launch_gather(anc32, reinterpret_cast<const long long*>(anc64), S, B, latents, n_latents,
                         (cudaStream_t)stream);
}

__global__ void __launch_bounds__(256) k_compose_ancestors(const int* __restrict__ prev, const int* __restrict__ next,
                                                           long long R, int B, int* __restrict__ out) {
    const long long r = (long long)blockIdx.x * 256 + threadIdx.x;
    if (r < R) out[r] = __ldg(prev + (long long)__ldg(next + r) * B + (r % B));
}

extern "C" int apgs_compose_ancestors(const int32_t* prev, const int32_t* next, int64_t S, int64_t B, int32_t* out,
                                      void* stream) {
    if (int e = check_sb(S, B, "compose_ancestors")) return e;
    if (!prev || !next || !out) return fail(APGS_E_BADARG, "compose_ancestors: NULL pointer");
    if (B >= (1ll << 31) || S * B >= (1ll << 38)) return fail(APGS_E_SHAPE, "compose_ancestors: S=%lld B=%lld too large", (long long)S, (long long)B);
    const long long R = S * B;
    k_compose_ancestors<<<(unsigned int)((R + 255) / 256), 256, 0, (cudaStream_t)stream>>>(prev, next, R, (int)B, out);
    APGS_LAUNCH_OK("k_compose_ancestors launch");
    return 0;
}

extern "C" int apgs_gather_ancestors_via(const int32_t* anc32, int64_t S, int64_t B, const apgs_latent_t* latents,
                                         const int32_t* const* first, int n_latents, void* stream) {
    if (int e = check_sb(S, B, "gather_ancestors_via")) return e;
    if (anc32 == nullptr) return fail(APGS_E_BADARG, "gather_ancestors_via: anc32 is NULL");
    if ((latents == nullptr || first == nullptr) && n_latents > 0) return fail(APGS_E_BADARG, "gather_ancestors_via: latents / first is NULL");
    return launch_gather(anc32, nullptr, S, B, latents, n_latents, (cudaStream_t)stream, first);
}

extern "C" int apgs_resample_f32(const float* log_w, const float* w_norm, int64_t S, int64_t B, int strategy,
                                 const void* uniforms, int64_t u_stride_s, int64_t u_stride_b, int cdf_mode,
                                 const apgs_latent_t* latents, int n_latents, int64_t* anc64, float* w_out,
                                 float* ess, float* lse, void* workspace, size_t workspace_bytes,
                                 void* stream) {
    if (int e = check_sb(S, B, "resample")) return e;
    size_t need = apgs_resample_workspace_bytes(S, B);
    if (workspace == nullptr || workspace_bytes < need)
        return fail(APGS_E_WORKSPACE, "resample: workspace of %zu bytes required, got %zu", need, workspace_bytes);
    if (!aligned(workspace, 256)) return fail(APGS_E_ALIGN, "resample: workspace must be 256-byte aligned");
    size_t idx_bytes = ((size_t)S * (size_t)B * sizeof(int32_t) + 255) & ~(size_t)255;
    int32_t* anc32 = reinterpret_cast<int32_t*>(workspace);
    char* rest = reinterpret_cast<char*>(workspace) + idx_bytes;
    int e = apgs_sample_ancestors_f32(log_w, w_norm, S, B, strategy, uniforms, u_stride_s, u_stride_b, cdf_mode,
                                      anc64, anc32, w_out, ess, lse, rest, workspace_bytes - idx_bytes, stream);
    if (e) return e;
    return apgs_gather_ancestors(anc32, nullptr, S, B, latents, n_latents, stream);
}

extern "C" int apgs_local_weight_stats_f32(const float* log_w_local, int64_t S, int64_t B, int64_t N,
                                           float* w_local_out, float* log_w_out, void* stream) {
    if (int e = check_sb(S, B, "local_weight_stats")) return e;
    if (N <= 0 || N >= (1ll << 31)) return fail(APGS_E_SHAPE, "local_weight_stats: N=%lld", (long long)N);
    if (log_w_local == nullptr) return fail(APGS_E_BADARG, "local_weight_stats: input is NULL");
    cudaStream_t st = (cudaStream_t)stream;
    if (log_w_out != nullptr) {
        long long rows = (long long)S * B;
        long long blocks = (rows + 7) / 8;
        if (blocks >= (1ll << 31)) return fail(APGS_E_SHAPE, "local_weight_stats: too many rows");
        k_local_sum<<<(unsigned int)blocks, 256, 0, st>>>(log_w_local, rows, (int)N, log_w_out);
        APGS_LAUNCH_OK("k_local_sum launch");
    }
    if (w_local_out != nullptr) {
        long long C = (long long)B * N;
        const size_t tile = (size_t)S * kLsTC * sizeof(float);
        if (tile <= kSmemTileBudget) {
            long long blocks = (C + kLsTC - 1) / kLsTC;
            if (blocks >= (1ll << 31)) return fail(APGS_E_SHAPE, "local_weight_stats: too many columns");
            const bool vec = (C & 3) == 0 && aligned(log_w_local, 16) && aligned(w_local_out, 16);
            auto kern = vec ? k_local_softmax_tile<true> : k_local_softmax_tile<false>;
            APGS_CUDA_OK(ensure_dynamic_smem(reinterpret_cast<const void*>(kern), tile), "smem attr");
            kern<<<(unsigned int)blocks, kLsThreads, tile, st>>>(log_w_local, (int)S, C, w_local_out);
            APGS_LAUNCH_OK("k_local_softmax_tile launch");
        } else {
            // S beyond a shared-memory tile: three passes over global memory
            long long blocks = (C + 255) / 256;
            if (blocks >= (1ll << 31)) return fail(APGS_E_SHAPE, "local_weight_stats: too many columns");
            k_local_softmax<<<(unsigned int)blocks, 256, 0, st>>>(log_w_local, S, C, w_local_out);
            APGS_LAUNCH_OK("k_local_softmax launch");
        }
    }
    return 0;
}
