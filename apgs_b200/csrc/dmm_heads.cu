// Enc_rws_mu / Enc_apg_mu heads of the DMM task in one launch (apgs/dmm/models.py: mean_mu, mean_log_sigma on the pooled
// statistics; forward :43-44 / :88-89):
//   q_mu (M,Z) = W2m tanh(W1m v + b1m) + b2m,   q_sigma (M,Z) = exp(W2s tanh(W1s v + b1s) + b2s),   v (M,H) = cat(nss, prior_mu, prior_nu)
// M = S*B*K rows (102400 at the C3 shape) of a tiny MLP (H = num_nss + 2 D = 12 -> 32 -> 2): as nn.Sequential modules that is
// two skinny GEMM pairs, their split-K reductions, two tanh, an exp and the bias adds -- ~10 launches per call, 8 % of the C3 step
// in launches of a few microseconds each.  Here a thread owns a row: the row sits in registers, both heads' weights in shared
// memory (broadcast reads), every sum in ascending k, tanhf / expf as torch's own kernels use them.
#include "common.cuh"

namespace apgs {

struct MlpHeadsArgs {
    const float* v;                              // (M, H)
    const float *w1m, *b1m, *w2m, *b2m;          // (H2, H), (H2), (Z, H2), (Z)
    const float *w1s, *b1s, *w2s, *b2s;
    float *mean, *std;                           // (M, Z)
    long long M;
    int H, H2, Z;
};

template <int HP>                                // H padded to a multiple of 4
__global__ void __launch_bounds__(256) k_mlp_heads_rowwise(MlpHeadsArgs a) {
    extern __shared__ __align__(16) float s_w[];
    const int H2 = a.H2, Z = a.Z, U = 2 * H2;
    float* s_w1 = s_w;                           // [U][HP]   (mean units, then sigma units; zero padded)
    float* s_b1 = s_w1 + U * HP;                 // [U]
    float* s_w2 = s_b1 + U;                      // [2][Z][H2]
    float* s_b2 = s_w2 + 2 * Z * H2;             // [2][Z]
    for (int i = threadIdx.x; i < U * HP; i += blockDim.x) {
        const int u = i / HP, k = i % HP;
        s_w1[i] = k < a.H ? __ldg((u < H2 ? a.w1m + (size_t)u * a.H : a.w1s + (size_t)(u - H2) * a.H) + k) : 0.f;
    }
    for (int i = threadIdx.x; i < U; i += blockDim.x) s_b1[i] = __ldg(i < H2 ? a.b1m + i : a.b1s + (i - H2));
    for (int i = threadIdx.x; i < 2 * Z * H2; i += blockDim.x) s_w2[i] = __ldg(i < Z * H2 ? a.w2m + i : a.w2s + (i - Z * H2));
    for (int i = threadIdx.x; i < 2 * Z; i += blockDim.x) s_b2[i] = __ldg(i < Z ? a.b2m + i : a.b2s + (i - Z));
    __syncthreads();
    const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= a.M) return;
    float x[HP];
#pragma unroll
    for (int k = 0; k < HP; ++k) x[k] = k < a.H ? __ldg(a.v + m * a.H + k) : 0.f;
    constexpr int ZMAX = 4;
    float out[2][ZMAX];
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int z = 0; z < ZMAX; ++z) out[h][z] = z < Z ? s_b2[h * Z + z] : 0.f;
    for (int h = 0; h < 2; ++h) {
        for (int u = 0; u < H2; ++u) {
            const float* w = s_w1 + (h * H2 + u) * HP;
            float acc = s_b1[h * H2 + u];
#pragma unroll
            for (int k4 = 0; k4 < HP / 4; ++k4) {
                const float4 w4 = *reinterpret_cast<const float4*>(w + 4 * k4);
                acc = fmaf(x[4 * k4], w4.x, acc);
                acc = fmaf(x[4 * k4 + 1], w4.y, acc);
                acc = fmaf(x[4 * k4 + 2], w4.z, acc);
                acc = fmaf(x[4 * k4 + 3], w4.w, acc);
            }
            const float t = tanhf(acc);
#pragma unroll
            for (int z = 0; z < ZMAX; ++z)
                if (z < Z) out[h][z] = fmaf(t, s_w2[(h * Z + z) * H2 + u], out[h][z]);
        }
    }
#pragma unroll
    for (int z = 0; z < ZMAX; ++z)
        if (z < Z) {
            a.mean[m * Z + z] = out[0][z];
            a.std[m * Z + z] = expf(out[1][z]);
        }
}

template <int HP>
static int launch_heads(const MlpHeadsArgs& a, cudaStream_t st) {
    const size_t smem = (size_t)(2 * a.H2 * HP + 2 * a.H2 + 2 * a.Z * a.H2 + 2 * a.Z) * sizeof(float);
    if (smem > 48 * 1024) return fail(APGS_E_SHAPE, "dmm_mu_heads: H2=%d too large", a.H2);
    k_mlp_heads_rowwise<HP><<<(unsigned int)((a.M + 255) / 256), 256, smem, st>>>(a);
    APGS_LAUNCH_OK("k_mlp_heads_rowwise launch");
    return 0;
}

}  // namespace apgs

using namespace apgs;

extern "C" int apgs_dmm_mu_heads_f32(const float* v, const float* w1_mean, const float* b1_mean, const float* w2_mean,
                                     const float* b2_mean, const float* w1_sigma, const float* b1_sigma, const float* w2_sigma,
                                     const float* b2_sigma, int64_t M, int64_t H, int64_t H2, int64_t Z, float* q_mu,
                                     float* q_sigma, void* stream) {
    if (M <= 0 || H <= 0 || H2 <= 0 || Z <= 0 || M >= (1ll << 40) || H2 > 1024)
        return fail(APGS_E_SHAPE, "dmm_mu_heads: bad sizes M=%lld H=%lld H2=%lld Z=%lld", (long long)M, (long long)H, (long long)H2,
                    (long long)Z);
    if (H > 32 || Z > 4)
        return fail(APGS_E_UNSUPPORTED, "dmm_mu_heads: compiled for H <= 32 inputs and Z <= 4 outputs (got %lld, %lld)", (long long)H,
                    (long long)Z);
    if (!v || !w1_mean || !b1_mean || !w2_mean || !b2_mean || !w1_sigma || !b1_sigma || !w2_sigma || !b2_sigma || !q_mu || !q_sigma)
        return fail(APGS_E_BADARG, "dmm_mu_heads: NULL pointer");
    MlpHeadsArgs a = {v, w1_mean, b1_mean, w2_mean, b2_mean, w1_sigma, b1_sigma, w2_sigma, b2_sigma, q_mu, q_sigma, (long long)M,
                      (int)H, (int)H2, (int)Z};
    cudaStream_t st = (cudaStream_t)stream;
    switch ((H + 3) / 4) {
        case 1: return launch_heads<4>(a, st);
        case 2: return launch_heads<8>(a, st);
        case 3: return launch_heads<12>(a, st);
        case 4: return launch_heads<16>(a, st);
        case 5: return launch_heads<20>(a, st);
        case 6: return launch_heads<24>(a, st);
        case 7: return launch_heads<28>(a, st);
        default: return launch_heads<32>(a, st);
    }
}
