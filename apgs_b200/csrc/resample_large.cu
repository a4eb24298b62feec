// Split-S ancestor pipeline for sm_100a: particle axes too long for a shared-memory CDF tile (S > ~51k; BPG runs
// Resampler('systematic', S * 100), apgs/gmm/evaluation.py:27,41-42).  The one-CTA-per-instance-tile kernel of
// resample.cu then falls to 1-2 CTAs with a global-memory tile (3.3 ms at S = 1e5, B = 10).  Here S is split into
// chunks of 1024 rows so that the whole GPU works on every instance, in five or six light launches over L2-resident data:
//
//   k_lg_max      chunk maxima of log_w                                   (S,B) read, coalesced along B
//   k_lg_exp      e = exp(x - max_b) stored TRANSPOSED ws[b][s] (32x32 shared-memory tiles), chunk sums in fp64
//   k_lg_weights  w = e / sum_b in place, chunk sums of w and w^2 in fp64
//   k_lg_renorm   (multinomial) the Categorical re-normalisation p = w / fp32(sum w), chunk sums of p
//   k_lg_scan     CDF: exclusive prefix of the chunk sums + in-chunk fp64 scan (lane strips + warp scan), normalised by
//                 the total and rounded to fp32 in place; chunk 0 also writes ess / lse
//   k_lg_search   ancestor = lower_bound(cdf_b, position): one thread per particle, lanes along s, so the probes of a
//                 warp share their top levels (the transposed layout makes the last levels share sectors too)
//
// Same arithmetic as the BLOCKED CDF mode of resample.cu (strip-parallel fp64 sums rounded to fp32; equal to the
// sequential fp64 chain except where re-association crosses an fp32 rounding boundary, ~2^-29 per element).
#include "resample.cuh"

namespace apgs {

struct LgArgs {
    AncArgs a;
    float* ws;          // [B][S]: e, then w, then the normalised CDF
    float* cmax;        // [nchunk][B]
    double* csum;       // [nchunk][B] sum of e (or of the given weights)
    double* cw;         // [nchunk][B] sum of w
    double* cq;         // [nchunk][B] sum of w^2
    double* cp;         // [nchunk][B] sum of the re-normalised probabilities (multinomial only)
    int nchunk;
};

__global__ void __launch_bounds__(256) k_lg_max(LgArgs g) {
    __shared__ float s_m[8][32];
    const long long S = g.a.S, B = g.a.B;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const long long b = (long long)blockIdx.y * 32 + lane;
    const long long r0 = (long long)blockIdx.x * kLargeChunk, r1 = min(r0 + kLargeChunk, S);
    float m = -INFINITY;
    if (b < B)
        for (long long r = r0 + w; r < r1; r += 8) m = fmaxf(m, __ldg(g.a.log_w + r * B + b));
    s_m[w][lane] = m;
    __syncthreads();
    if (w == 0 && b < B) {
#pragma unroll
        for (int i = 1; i < 8; ++i) m = fmaxf(m, s_m[i][lane]);
        g.cmax[(long long)blockIdx.x * B + b] = m;
    }
}

// e (or the given normalised weights) -> ws[b][s], chunk sums.  have_logw: exponentiate against the column maximum.
__global__ void __launch_bounds__(256) k_lg_exp(LgArgs g, int have_logw) {
    __shared__ float s_t[8][32][33];
    __shared__ double s_s[8][32];
    const long long S = g.a.S, B = g.a.B;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const long long b0 = (long long)blockIdx.y * 32, b = b0 + lane;
    const long long r0 = (long long)blockIdx.x * kLargeChunk, r1 = min(r0 + kLargeChunk, S);
    float gmax = 0.f;
    if (have_logw && b < B) {
        gmax = -INFINITY;
        for (int c = 0; c < g.nchunk; ++c) gmax = fmaxf(gmax, g.cmax[(long long)c * B + b]);
    }
    const float* src = have_logw ? g.a.log_w : g.a.w_norm;
    double acc = 0.0;
    for (long long rg = r0 + (long long)w * 32; rg < r1; rg += 8 * 32) {        // 32-row groups of this warp
        float part = 0.f;
#pragma unroll 4
        for (int i = 0; i < 32; ++i) {
            const long long r = rg + i;
            float v = 0.f;
            if (r < r1 && b < B) {
                v = __ldg(src + r * B + b);
                if (have_logw) v = expf(v - gmax);
            }
            s_t[w][i][lane] = v;
            part += v;
        }
        acc += (double)part;
        __syncwarp();
        const long long r = rg + lane;
#pragma unroll 4
        for (int j = 0; j < 32; ++j)
            if (r < r1 && b0 + j < B) g.ws[(b0 + j) * S + r] = s_t[w][lane][j];
        __syncwarp();
    }
    s_s[w][lane] = acc;
    __syncthreads();
    if (w == 0 && b < B) {
#pragma unroll
        for (int i = 1; i < 8; ++i) acc += s_s[i][lane];
        g.csum[(long long)blockIdx.x * B + b] = acc;
    }
}

__device__ __forceinline__ double block_sum_d(double v, double* s_red) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) s_red[w] = v;
    __syncthreads();
    double r = 0.0;
    for (int i = 0; i < 8; ++i) r += s_red[i];
    return r;
}

// sum over the chunks of one instance column (warp 0, lanes stride over chunks), broadcast to the CTA.  Every CTA of a
// column runs the same code, so all of them get bit-identical totals.
__device__ __forceinline__ double column_total(const double* arr, int nchunk, long long B, long long b, double* s_bc) {
    if (threadIdx.x < 32) {
        double t = 0.0;
        for (int c = threadIdx.x; c < nchunk; c += 32) t += arr[(long long)c * B + b];
        t = warp_sum(t);
        if (threadIdx.x == 0) *s_bc = t;
    }
    __syncthreads();
    const double r = *s_bc;
    __syncthreads();
    return r;
}

// w = e / total in place; chunk sums of w and w^2.  grid (nchunk, B); thread t owns 4 consecutive rows.
__global__ void __launch_bounds__(256) k_lg_weights(LgArgs g, int have_logw) {
    __shared__ double s_red[8];
    const long long S = g.a.S, B = g.a.B, b = blockIdx.y;
    const long long r0 = (long long)blockIdx.x * kLargeChunk + threadIdx.x * 4;
    __shared__ double s_bc;
    const float tot = have_logw ? (float)column_total(g.csum, g.nchunk, B, b, &s_bc) : 1.f;
    float* p = g.ws + b * S;
    double sw = 0.0, sq = 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const long long r = r0 + i;
        if (r < S) {
            float v = p[r];
            if (have_logw) { v = __fdiv_rn(v, tot); p[r] = v; }
            sw += (double)v;
            sq += (double)__fmul_rn(v, v);
        }
    }
    sw = block_sum_d(sw, s_red);
    sq = block_sum_d(sq, s_red);
    if (threadIdx.x == 0) {
        g.cw[(long long)blockIdx.x * B + b] = sw;
        g.cq[(long long)blockIdx.x * B + b] = sq;
    }
}

// multinomial only: Categorical(probs) renormalises once more before torch.multinomial (categorical.py:70, reached from
// apgs/resampler.py:41): p = w / fp32(sum w) in place, chunk sums of p
__global__ void __launch_bounds__(256) k_lg_renorm(LgArgs g) {
    __shared__ double s_red[8];
    const long long S = g.a.S, B = g.a.B, b = blockIdx.y;
    const long long r0 = (long long)blockIdx.x * kLargeChunk + threadIdx.x * 4;
    __shared__ double s_bc;
    const float sw = (float)column_total(g.cw, g.nchunk, B, b, &s_bc);
    float* p = g.ws + b * S;
    double sp = 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const long long r = r0 + i;
        if (r < S) {
            const float v = __fdiv_rn(p[r], sw);
            p[r] = v;
            sp += (double)v;
        }
    }
    sp = block_sum_d(sp, s_red);
    if (threadIdx.x == 0) g.cp[(long long)blockIdx.x * B + b] = sp;
}

// optional (S,B) weights output, straight from log_w with the same arithmetic as k_lg_exp / k_lg_weights
__global__ void __launch_bounds__(256) k_lg_wout(LgArgs g) {
    const long long S = g.a.S, B = g.a.B;
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= S * B) return;
    const long long b = i % B;
    float gmax = -INFINITY;
    double t = 0.0;
    for (int c = 0; c < g.nchunk; ++c) { gmax = fmaxf(gmax, g.cmax[(long long)c * B + b]); t += g.csum[(long long)c * B + b]; }
    g.a.w_out[i] = __fdiv_rn(expf(__ldg(g.a.log_w + i) - gmax), (float)t);
}

// CDF of one (chunk, instance): ws[b][r] <- fp32( (offset + inclusive fp64 prefix) ) / fp32(total)
__global__ void __launch_bounds__(256) k_lg_scan(LgArgs g, int have_logw, int finalize_only) {
    __shared__ double s_w[8];
    const long long S = g.a.S, B = g.a.B, b = blockIdx.y;
    const int c = blockIdx.x;
    __shared__ double s_bc[4];
    __shared__ float s_gm;
    const double* cs = (g.a.strategy == APGS_MULTINOMIAL && have_logw) ? g.cp : g.cw;
    if (threadIdx.x < 32) {                       // warp 0: offset of this chunk, total, and (chunk 0) the ESS / lse sums
        double off = 0.0, tot = 0.0, q = 0.0, se = 0.0;
        float gmax = -INFINITY;
        for (int i = threadIdx.x; i < g.nchunk; i += 32) {
            const double v = cs[(long long)i * B + b];
            if (i < c) off += v;
            tot += v;
            if (c == 0) {
                q += g.cq[(long long)i * B + b];
                if (have_logw) { se += g.csum[(long long)i * B + b]; gmax = fmaxf(gmax, g.cmax[(long long)i * B + b]); }
            }
        }
        off = warp_sum(off); tot = warp_sum(tot);
        if (c == 0) { q = warp_sum(q); se = warp_sum(se); gmax = warp_max(gmax); }
        if (threadIdx.x == 0) { s_bc[0] = off; s_bc[1] = tot; s_bc[2] = q; s_bc[3] = se; s_gm = gmax; }
    }
    __syncthreads();
    const double off = s_bc[0], tot = s_bc[1];
    if (c == 0 && threadIdx.x == 0) {
        if (g.a.ess != nullptr) g.a.ess[b] = __fdiv_rn(1.f, (float)s_bc[2]);
        if (g.a.lse != nullptr && have_logw) g.a.lse[b] = s_gm + logf((float)s_bc[3]);
    }
    if (finalize_only) return;
    const float norm = (float)tot;
    float* p = g.ws + b * S;
    const long long r0 = (long long)c * kLargeChunk + threadIdx.x * 4;
    float v[4];
    double loc = 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i) { v[i] = (r0 + i < S) ? p[r0 + i] : 0.f; loc += (double)v[i]; }
    // exclusive prefix of the 256 strip sums: warp scan, then the 8 warp totals
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    double inc = loc;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_w[w] = inc;
    __syncthreads();
    double base = off;
    for (int i = 0; i < w; ++i) base += s_w[i];
    double run = base + (inc - loc);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        run += (double)v[i];
        if (r0 + i < S) p[r0 + i] = __fdiv_rn((float)run, norm);
    }
}

// CTA = 32 instance columns (one per warp) x 32 consecutive particles (lanes): the probes of a warp walk one CDF
// together; the 32 x 32 results are transposed through shared memory so that the index stores are coalesced along b.
__global__ void __launch_bounds__(1024) k_lg_search(LgArgs g) {
    __shared__ int s_a[32][33];
    const long long S = g.a.S, B = g.a.B;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const long long b = (long long)blockIdx.y * 32 + w;
    const long long s = (long long)blockIdx.x * 32 + lane;
    long long lo = 0, hi = S;
    if (b < B && s < S) {
        const float* cdf = g.ws + b * S;
        if (g.a.strategy == APGS_SYSTEMATIC) {
            const float pos = __fdiv_rn(__fadd_rn(__ldg(g.a.u_sys + b), (float)s), (float)S);      // resampler.py:28
            while (lo < hi) {
                const long long mid = (lo + hi) >> 1;
                if (cdf[mid] < pos) lo = mid + 1; else hi = mid;
            }
        } else {
            const double ud = __ldg(g.a.u_mult + s * g.a.us + b * g.a.ub);
            while (lo < hi) {
                const long long mid = (lo + hi) >> 1;
                if ((double)cdf[mid] < ud) lo = mid + 1; else hi = mid;
            }
        }
        if (lo > S - 1) lo = S - 1;                        // memory safety on NaN / degenerate weights
    }
    s_a[w][lane] = (int)lo;
    __syncthreads();
    // warp w now stores particle row s0 + w: 32 consecutive instances
    const long long so = (long long)blockIdx.x * 32 + w, bo = (long long)blockIdx.y * 32 + lane;
    if (so < S && bo < B) {
        const int v = s_a[lane][w];
        if (g.a.anc32 != nullptr) g.a.anc32[so * B + bo] = v;
        if (g.a.anc64 != nullptr) g.a.anc64[so * B + bo] = v;
    }
}

static size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

size_t ancestors_large_workspace_bytes(long long S, long long B) {
    const size_t nchunk = (size_t)((S + kLargeChunk - 1) / kLargeChunk);
    return align256((size_t)S * B * 4) + align256(nchunk * B * 4) + 4 * align256(nchunk * B * 8);
}

int launch_ancestors_large(const AncArgs& a, void* workspace, size_t workspace_bytes, cudaStream_t st) {
    const size_t need = ancestors_large_workspace_bytes(a.S, a.B);
    if (workspace == nullptr || workspace_bytes < need)
        return fail(APGS_E_WORKSPACE, "ancestors: workspace of %zu bytes required (S=%lld B=%lld), got %zu", need, a.S,
                    a.B, workspace_bytes);
    if (!aligned(workspace, 256)) return fail(APGS_E_ALIGN, "ancestors: workspace must be 256-byte aligned");
    if (a.B > 65535) return fail(APGS_E_SHAPE, "ancestors: B=%lld too large for S=%lld (split-S path takes B <= 65535)", a.B, a.S);
    LgArgs g = {};
    g.a = a;
    g.nchunk = (int)((a.S + kLargeChunk - 1) / kLargeChunk);
    char* p = reinterpret_cast<char*>(workspace);
    g.ws = reinterpret_cast<float*>(p);            p += align256((size_t)a.S * a.B * 4);
    g.cmax = reinterpret_cast<float*>(p);          p += align256((size_t)g.nchunk * a.B * 4);
    g.csum = reinterpret_cast<double*>(p);         p += align256((size_t)g.nchunk * a.B * 8);
    g.cw = reinterpret_cast<double*>(p);           p += align256((size_t)g.nchunk * a.B * 8);
    g.cq = reinterpret_cast<double*>(p);           p += align256((size_t)g.nchunk * a.B * 8);
    g.cp = reinterpret_cast<double*>(p);
    const int have_logw = a.w_norm == nullptr ? 1 : 0;
    const dim3 tile_grid((unsigned int)g.nchunk, (unsigned int)((a.B + 31) / 32));
    const dim3 col_grid((unsigned int)g.nchunk, (unsigned int)a.B);
    if (have_logw) {
        k_lg_max<<<tile_grid, 256, 0, st>>>(g);
        APGS_LAUNCH_OK("k_lg_max launch");
    }
    k_lg_exp<<<tile_grid, 256, 0, st>>>(g, have_logw);
    APGS_LAUNCH_OK("k_lg_exp launch");
    k_lg_weights<<<col_grid, 256, 0, st>>>(g, have_logw);
    APGS_LAUNCH_OK("k_lg_weights launch");
    if (a.w_out != nullptr && have_logw) {
        k_lg_wout<<<(unsigned int)((a.S * a.B + 255) / 256), 256, 0, st>>>(g);
        APGS_LAUNCH_OK("k_lg_wout launch");
    }
    const int stats_only = a.strategy < 0 ? 1 : 0;
    if (a.strategy == APGS_MULTINOMIAL && have_logw) {          // given weights are already what torch.multinomial sees
        k_lg_renorm<<<col_grid, 256, 0, st>>>(g);
        APGS_LAUNCH_OK("k_lg_renorm launch");
    }
    k_lg_scan<<<stats_only ? dim3(1, (unsigned int)a.B) : col_grid, 256, 0, st>>>(g, have_logw, stats_only);
    APGS_LAUNCH_OK("k_lg_scan launch");
    if (stats_only) return 0;
    k_lg_search<<<dim3((unsigned int)((a.S + 31) / 32), (unsigned int)((a.B + 31) / 32)), 1024, 0, st>>>(g);
    APGS_LAUNCH_OK("k_lg_search launch");
    return 0;
}

}  // namespace apgs
