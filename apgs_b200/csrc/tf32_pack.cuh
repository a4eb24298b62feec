// TF32 halves and the packed (MMA-ready) operand layout shared by k_pack_rows / k_linear_tc5pk (linear_tf32x3.cu) and the
// kernel that writes its result straight into that layout (k_template_match, bmnist_match.cu).
#pragma once
#include "common.cuh"

namespace apgs {

__device__ __forceinline__ unsigned int to_tf32(float v) {
    unsigned int r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return r;
}
// v = hi + lo (+ ~2^-22 |v|), both halves exactly representable in TF32
__device__ __forceinline__ void split_tf32(float v, unsigned int& hi, unsigned int& lo) {
    hi = to_tf32(v);
    lo = to_tf32(v - __uint_as_float(hi));
}

// Packed matrix of R rows x K columns with TR rows per tile: tile (rt, c) = rows rt*TR .. rt*TR+TR-1, columns 32c .. 32c+31 at
// float offset (rt * ceil(K/32) + c) * 2*TR*32 as { hi[TR x 32], lo[TR x 32] }; element (r, k) of a half at
// ((r % TR) / 8) * 256 + ((k % 32) / 4) * 32 + (r % 8) * 4 + k % 4 -- the K-major, no-swizzle canonical shared-memory layout of
// tcgen05.mma (core matrices of 8 rows x 16 bytes, 128 bytes apart along K, 1 KB apart along the rows), so that a tile is one
// bulk copy away from the tensor core.  Returns the index of the hi element; the lo element sits TR*32 floats behind it.
__host__ __device__ inline size_t packed_index(int TR, long long r, long long k, long long nchunks) {
    const long long tile = (r / TR) * nchunks + k / 32;
    return (size_t)tile * (size_t)(2 * TR * 32) + (size_t)(((r % TR) / 8) * 256 + ((k % 32) / 4) * 32 + (r % 8) * 4 + (k % 4));
}
constexpr int kPackRowsX = 128, kPackRowsW = 80;      // rows per tile of the activations / of the weights (the CTA tile of k_linear_tc5pk)

}  // namespace apgs
