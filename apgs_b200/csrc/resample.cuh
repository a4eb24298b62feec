// Shared declarations of the ancestor kernels (resample.cu: one CTA per instance tile; resample_large.cu: the
// split-S pipeline for particle axes that do not fit a shared-memory tile).
#pragma once
#include "common.cuh"

namespace apgs {

struct AncArgs {
    const float* log_w;
    const float* w_norm;
    const float* u_sys;        // [B]
    const double* u_mult;      // strided
    long long us, ub;
    long long S, B;
    int strategy;              // APGS_SYSTEMATIC / APGS_MULTINOMIAL / -1 (stats only)
    int blocked;               // CDF mode
    int tile_in_smem;
    int dbl;                   // exact systematic CDF kept in a shared fp64 tile (no conversions on the chain)
    float* ws;                 // global tile (S,B) when !tile_in_smem
    long long* anc64;
    int* anc32;
    float* w_out;
    float* ess;
    float* lse;
};

// ---- split-S pipeline (resample_large.cu) ----------------------------------------------------------------------
constexpr int kLargeChunk = 1024;          // particle rows per CTA
size_t ancestors_large_workspace_bytes(long long S, long long B);
int launch_ancestors_large(const AncArgs& a, void* workspace, size_t workspace_bytes, cudaStream_t st);

}  // namespace apgs
