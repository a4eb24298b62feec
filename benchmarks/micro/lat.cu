// dependent-issue latency of DADD / FADD / LDS.64+DADD+STS.64 on one warp (sm_100a)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, int n, double w) {
    __shared__ double sm[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = w * (i + 1);
    __syncthreads();
    double acc = out[0];
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) acc += w;
    long long t1 = clock64();
    float f = (float)acc, fw = (float)w;
    long long t2 = clock64();
    for (int i = 0; i < n; ++i) f += fw;
    long long t3 = clock64();
    double a2 = acc;
    long long t4 = clock64();
#pragma unroll 16
    for (int i = 0; i < 4096; ++i) { a2 += sm[i]; sm[i] = a2; }
    long long t5 = clock64();
    float f2 = f;
    long long t6 = clock64();
    for (int i = 0; i < n; ++i) { f2 = (float)((double)f2 + w); }
    long long t7 = clock64();
    if (threadIdx.x == 0) { out[0] = acc + f + a2 + sm[100] + f2; cyc[0] = t1 - t0; cyc[1] = t3 - t2; cyc[2] = t5 - t4; cyc[3] = t7 - t6; }
}
int main() {
    double* d; long long* c; cudaMalloc(&d, 8); cudaMalloc(&c, 64); cudaMemset(d, 0, 8);
    int n = 4096;
    for (int it = 0; it < 2; ++it) k<<<1, 4>>>(d, c, n, 1e-3);
    cudaDeviceSynchronize();
    long long h[4]; cudaMemcpy(h, c, 32, cudaMemcpyDeviceToHost);
    printf("DADD dep %.1f cyc, FADD dep %.1f cyc, LDS64+DADD+STS64 %.1f cyc/elem, F2F+DADD+F2F %.1f\n", h[0] / (double)n, h[1] / (double)n, h[2] / 4096.0, h[3]/(double)n);
    return 0;
}
