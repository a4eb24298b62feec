// Enc_coor's two heads on the hidden layer, one launch (apgs/bmnist/models.py:16-26,31-32):
//   q_mean = tanh(W2m relu(W1m h + b1m) + b2m),   q_std = exp(W2s relu(W1s h + b1s) + b2s),   h (M, H) = enc_hidden(conved)
// With M = S*B = 640 rows and H = 400 -> 200 -> 2 these are four skinny GEMMs that cuBLAS serves with ~13 launches per call
// (SIMT sgemm + split-K reductions + bias / ReLU / tanh / exp elementwise kernels: ~65 us of launches for 0.1 GFMA), 150 calls
// per 5-sweep objective -- more than the 4761 -> 400 layer in front of them.  Here: a CTA owns R = 8 rows, stages them in shared
// memory, and thread u owns hidden unit u of the concatenated first layers (2 * H2 units): it streams its column of the
// TRANSPOSED, 4-k interleaved first-layer weights ([H/4][2 H2][4] floats, built once per set of weights by the host: one
// coalesced 16-byte load per four k) against the R rows (broadcast 16-byte shared loads): 4 R FMAs per 1 + R loads.  Two threads
// per unit take one half of k each (the chain of dependent L2 loads is what bounds a launch this small) and meet in shared
// memory: every sum is in a fixed order.  The second layers (2 x Z outputs of H2 terms per row) are warp reductions on the ReLU'd tile.  fp32 throughout;
// agrees with the nn.Sequential modules to the rounding of a different summation order (~1e-7 relative).
#include "common.cuh"

namespace apgs {

struct HeadsArgs {
    const float* hidden;   // (M, H)
    const float4* w1t;     // [K4][U] float4: element (k4, u).{x,y,z,w} = W1cat[u][4 k4 + 0..3], W1cat = [W1m; W1s] (U = 2 H2 rows), zero padded
    const float* b1;       // (U) = [b1m; b1s]
    const float *w2m, *b2m, *w2s, *b2s;   // (Z, H2), (Z)
    float *mean, *std;     // (M, Z)
    int M, H, K4, H2, Z;
};

template <int R>
__global__ void __launch_bounds__(1024) k_coor_heads(HeadsArgs a, int UT) {
    extern __shared__ __align__(16) float s_hd[];
    const int U = 2 * a.H2, KP = a.K4 * 4;
    float* s_h = s_hd;                 // [R][KP]
    float* s_a = s_hd + R * KP;        // [R][U]   ReLU'd first layers
    float* s_p = s_a + R * U;          // [R][U]   partial sums of the second half of k
    const int m0 = (int)blockIdx.x * R;
    for (int i = threadIdx.x; i < R * KP; i += blockDim.x) {
        const int r = i / KP, k = i % KP;
        s_h[i] = (m0 + r < a.M && k < a.H) ? __ldg(a.hidden + (size_t)(m0 + r) * a.H + k) : 0.f;
    }
    __syncthreads();
    // thread (part, u): unit u, half `part` of k -- two threads per unit halve the chain of dependent L2 loads
    const int part = (int)threadIdx.x / UT, u0 = (int)threadIdx.x % UT;
    const int kh = (a.K4 + 1) / 2, kb = part * kh, ke = min(a.K4, kb + kh);
    for (int u = u0; u < U; u += UT) {
        float acc[R];
        const float bias = part == 0 ? __ldg(a.b1 + u) : 0.f;
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = bias;
        const float4* wp = a.w1t + u;
#pragma unroll(R == 8 ? 3 : 5)
        for (int k4 = kb; k4 < ke; ++k4) {
            const float4 w = __ldg(wp + (size_t)k4 * U);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const float4 h = *reinterpret_cast<const float4*>(s_h + r * KP + k4 * 4);
                acc[r] = fmaf(h.x, w.x, acc[r]);
                acc[r] = fmaf(h.y, w.y, acc[r]);
                acc[r] = fmaf(h.z, w.z, acc[r]);
                acc[r] = fmaf(h.w, w.w, acc[r]);
            }
        }
        float* dst = part == 0 ? s_a : s_p;
#pragma unroll
        for (int r = 0; r < R; ++r) dst[r * U + u] = acc[r];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < R * U; i += blockDim.x) s_a[i] = fmaxf(s_a[i] + s_p[i], 0.f);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    for (int p = warp; p < R * 2 * a.Z; p += nwarp) {
        const int r = p / (2 * a.Z), o = p % (2 * a.Z), head = o / a.Z, z = o % a.Z;
        const float* w2 = (head ? a.w2s : a.w2m) + (size_t)z * a.H2;
        const float* act = s_a + r * U + head * a.H2;
        float sum = 0.f;
        for (int i = lane; i < a.H2; i += 32) sum = fmaf(act[i], __ldg(w2 + i), sum);
        sum = warp_sum(sum);
        if (lane == 0 && m0 + r < a.M) {
            const float v = sum + __ldg((head ? a.b2s : a.b2m) + z);
            if (head) a.std[(size_t)(m0 + r) * a.Z + z] = expf(v);
            else a.mean[(size_t)(m0 + r) * a.Z + z] = tanhf(v);
        }
    }
}

}  // namespace apgs

using namespace apgs;

extern "C" int apgs_bmnist_coor_heads_f32(const float* hidden, const float* w1t, const float* b1, const float* w2_mean,
                                          const float* b2_mean, const float* w2_std, const float* b2_std, int64_t M, int64_t H,
                                          int64_t H2, int64_t Z, float* q_mean, float* q_std, void* stream) {
    if (M <= 0 || H <= 0 || H2 <= 0 || Z <= 0 || M >= (1ll << 30) || H > 8192 || H2 > 4096 || Z > 16)
        return fail(APGS_E_SHAPE, "bmnist_coor_heads: bad sizes M=%lld H=%lld H2=%lld Z=%lld", (long long)M, (long long)H, (long long)H2,
                    (long long)Z);
    if (!hidden || !w1t || !b1 || !w2_mean || !b2_mean || !w2_std || !b2_std || !q_mean || !q_std)
        return fail(APGS_E_BADARG, "bmnist_coor_heads: NULL pointer");
    if (!aligned(w1t, 16)) return fail(APGS_E_BADARG, "bmnist_coor_heads: w1t must be 16-byte aligned");
    // rows per CTA: 8, or 5 when that still leaves no more CTAs than SMs (M = 640 -> 128 CTAs on 148 SMs instead of 80)
    const int R = (M + 4) / 5 <= sm_count() ? 5 : 8;
    const int K4 = (int)((H + 3) / 4), U = (int)(2 * H2);
    const size_t smem = (size_t)R * (K4 * 4 + 2 * U) * sizeof(float);
    if (smem > 200 * 1024) return fail(APGS_E_SHAPE, "bmnist_coor_heads: H=%lld H2=%lld too large for one tile", (long long)H, (long long)H2);
    HeadsArgs a = {hidden, reinterpret_cast<const float4*>(w1t), b1, w2_mean, b2_mean, w2_std, b2_std, q_mean, q_std,
                   (int)M, (int)H, K4, (int)H2, (int)Z};
    const int UT = min(512, max(32, (U + 31) / 32 * 32));            // threads per half of k
    const unsigned int grid = (unsigned int)((M + R - 1) / R);
    if (R == 5) {
        APGS_CUDA_OK(ensure_dynamic_smem(reinterpret_cast<const void*>(k_coor_heads<5>), smem), "smem attr");
        k_coor_heads<5><<<grid, 2 * UT, smem, (cudaStream_t)stream>>>(a, UT);
    } else {
        APGS_CUDA_OK(ensure_dynamic_smem(reinterpret_cast<const void*>(k_coor_heads<8>), smem), "smem attr");
        k_coor_heads<8><<<grid, 2 * UT, smem, (cudaStream_t)stream>>>(a, UT);
    }
    APGS_LAUNCH_OK("k_coor_heads launch");
    return 0;
}
