// Throughput of the fused z step's inner loop (frac4 over quads of hidden units, packed fp32x2) as a function of the number
// of resident warps per scheduler: separates "FMA-pipe bound" from "latency bound" for k_gmm_z_fused's main loop.
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o mlp_loop mlp_loop.cu ; run: ./mlp_loop
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float rcp_ftz(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
template <int MODE>
__device__ __forceinline__ f32x2 frac4(const f32x2 (&ex)[4], const f32x2 (&ee)[4], const f32x2 (&aw)[4], f32x2 c2, f32x2 acc) {
    const f32x2 y0 = fma2(ex[0], ee[0], c2), y1 = fma2(ex[1], ee[1], c2);
    const f32x2 y2 = fma2(ex[2], ee[2], c2), y3 = fma2(ex[3], ee[3], c2);
    const f32x2 n01 = fma2(aw[0], y1, mul2(aw[1], y0)), d01 = mul2(y0, y1);
    const f32x2 n23 = fma2(aw[2], y3, mul2(aw[3], y2)), d23 = mul2(y2, y3);
    const f32x2 n = fma2(n01, d23, mul2(n23, d01));
    float dl, dh;
    unpack2(mul2(d01, d23), dl, dh);
    if (MODE == 1) return fma2(n, pack2(dl, dh), acc);
    return fma2(n, pack2(rcp_ftz(dl), rcp_ftz(dh)), acc);
}
template <int K, int MODE>
__device__ __forceinline__ void mlp_quad_fast(const unsigned char* ex_ptr, int ex_stride_b, const unsigned char* ev_ptr,
                                              const unsigned char* aw_ptr, int H4, f32x2 (&L02)[K], f32x2 (&L13)[K]) {
    const f32x2 c2 = pack2(4.656612873077393e-10f, 4.656612873077393e-10f);
#pragma unroll 1
    for (int h = 0; h < H4; h += 4) {
        f32x2 ex02[4], ex13[4], aw[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const ulonglong2 exq = *reinterpret_cast<const ulonglong2*>(ex_ptr + (size_t)(h + i) * ex_stride_b);
            ex02[i] = exq.x; ex13[i] = exq.y;
        }
        {
            const float4 v = *reinterpret_cast<const float4*>(aw_ptr + (size_t)h * 4);
            aw[0] = pack2(v.x, v.x); aw[1] = pack2(v.y, v.y); aw[2] = pack2(v.z, v.z); aw[3] = pack2(v.w, v.w);
        }
        const float4* evq = reinterpret_cast<const float4*>(ev_ptr + (size_t)h * K * 4);
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const float4 v = evq[k];
            const f32x2 ee[4] = {pack2(v.x, v.x), pack2(v.y, v.y), pack2(v.z, v.z), pack2(v.w, v.w)};
            L02[k] = frac4<MODE>(ex02, ee, aw, c2, L02[k]);
            L13[k] = frac4<MODE>(ex13, ee, aw, c2, L13[k]);
        }
    }
}
constexpr int K = 3, H4 = 32, EXS = 64, EPB = 132;
// pure streams of packed FMAs on register operands: 12 independent accumulators; BC: one operand is a scalar broadcast
template <bool BC>
__global__ void __launch_bounds__(256, 4) k_ffma2(float* out, int reps, float s0, float s1) {
    f32x2 acc[12];
    for (int i = 0; i < 12; ++i) acc[i] = pack2(1.f + threadIdx.x * 1e-6f, 1.f + i * 1e-6f);
    const f32x2 x = BC ? pack2(s0, s0) : pack2(s0, s0 * 1.0001f), y = pack2(s1, s1 * 0.9999f);
    for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int u = 0; u < 7; ++u)
#pragma unroll
            for (int i = 0; i < 12; ++i) acc[i] = fma2(acc[i], x, y);
    }
    float a = 0.f, lo, hi;
    for (int i = 0; i < 12; ++i) { unpack2(acc[i], lo, hi); a += lo + hi; }
    out[blockIdx.x * 256 + threadIdx.x] = a;
}
template <int MODE>
__global__ void __launch_bounds__(256, 4) k_loop(float* out, int reps) {
    extern __shared__ __align__(16) unsigned char smem[];
    float* s_ex = reinterpret_cast<float*>(smem);           // H4 * EXS
    float* s_aw = s_ex + H4 * EXS;                          // H4
    float* s_e = s_aw + H4;                                 // 32 * EPB
    for (int i = threadIdx.x; i < H4 * EXS; i += 256) s_ex[i] = 1.f + 1e-3f * (i & 15);
    for (int i = threadIdx.x; i < H4; i += 256) s_aw[i] = 0.01f * (i + 1);
    for (int i = threadIdx.x; i < 32 * EPB; i += 256) s_e[i] = 4.6e-10f * (1.f + 1e-3f * (i & 7));
    __syncthreads();
    const int p = threadIdx.x / 8, j = threadIdx.x % 8;
    f32x2 L02[K], L13[K];
    for (int k = 0; k < K; ++k) { L02[k] = 0ull; L13[k] = 0ull; }
    for (int r = 0; r < reps; ++r) {
        const unsigned char* exb = reinterpret_cast<const unsigned char*>(s_ex + (size_t)((r & 1) * 8 + j) * 4);
        mlp_quad_fast<K, MODE>(exb, EXS * 4, reinterpret_cast<const unsigned char*>(s_e + p * EPB), reinterpret_cast<const unsigned char*>(s_aw), H4, L02, L13);
    }
    float a = 0.f, lo, hi;
    for (int k = 0; k < K; ++k) { unpack2(L02[k], lo, hi); a += lo + hi; unpack2(L13[k], lo, hi); a += lo + hi; }
    out[blockIdx.x * 256 + threadIdx.x] = a;
}
int main() {
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    const int sms = pr.multiProcessorCount; int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    float* out; cudaMalloc(&out, sizeof(float) * 256 * sms * 8);
    const int reps = 512;                                   // passes of 8 quad iterations
    const size_t base = (size_t)(H4 * EXS + H4 + 32 * EPB) * 4;
    cudaFuncSetAttribute(k_loop<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(k_loop<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    printf("SMs %d, clock %.0f MHz (nominal)\n", sms, clk_khz / 1e3);
    for (int ctas = 1; ctas <= 4; ++ctas) {                 // resident CTAs per SM (8 warps each): limited through the smem size
        size_t smem = 220 * 1024 / ctas - 1024; if (smem < base) smem = base; if (smem > 200 * 1024) smem = 200 * 1024;
        int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_loop<0>, 256, smem);
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        k_loop<0><<<sms * occ, 256, smem>>>(out, 8);
        cudaEventRecord(a); k_loop<0><<<sms * occ, 256, smem>>>(out, reps); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        {
            cudaEventRecord(a); k_loop<1><<<sms * occ, 256, smem>>>(out, reps); cudaEventRecord(b); cudaEventSynchronize(b);
            float m1; cudaEventElapsedTime(&m1, a, b);
            printf("   without the reciprocals: %.1f cycles per iteration per scheduler\n", m1 * 1e-3 * 1.965e9 / (reps * 8.0) / (occ * 2));
            for (int bc = 0; bc < 2; ++bc) {
                cudaEventRecord(a);
                if (bc) k_ffma2<true><<<sms * occ, 256>>>(out, 1024, 1.0001f, 1e-7f); else k_ffma2<false><<<sms * occ, 256>>>(out, 1024, 1.0001f, 1e-7f);
                cudaEventRecord(b); cudaEventSynchronize(b);
                cudaEventElapsedTime(&m1, a, b);
                printf("   pure FFMA2 stream (%s): %.2f cycles per FFMA2 per scheduler\n", bc ? "scalar-broadcast operand" : "packed operands",
                       m1 * 1e-3 * 1.965e9 / (1024.0 * 84) / (occ * 2));
            }
        }
        const double iters = (double)reps * 8;              // quad iterations per warp
        const double cyc = ms * 1e-3 * 1.965e9;             // at 1965 MHz
        printf("CTAs/SM %d (warps/scheduler %d): %.3f ms, %.0f cycles per quad iteration per warp, %.1f cycles per iteration per scheduler (84 FFMA2/FMUL2: floor 168)\n",
               occ, occ * 2, ms, cyc / iters, cyc / iters / (occ * 2));
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
