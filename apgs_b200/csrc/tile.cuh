// Particle-tile staging for sm_100a: per-particle rows that are contiguous in HBM ((S,B,...)
// latents) are brought into shared memory with the bulk-copy engine (cp.async.bulk -> SASS UBLKCP,
// completion on an mbarrier), so one thread moves a whole tile and every other thread keeps
// computing.  Falls back to cooperative vector loads when a slice is not 16-byte sized/aligned.
#pragma once
#include "common.cuh"

namespace apgs {

__device__ __forceinline__ unsigned int smem_u32(const void* p) {
    return (unsigned int)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned int bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned int parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
// global -> shared bulk copy; bytes % 16 == 0, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, unsigned int bytes,
                                         unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// One staging request: `bytes` contiguous bytes from `src` to `dst` (shared).
struct StageReq {
    void* dst;
    const void* src;
    unsigned int bytes;
};

// Stage up to NREQ slices.  Called by ALL threads of the CTA.  Slices that qualify go through the
// bulk-copy engine (issued by thread 0 against `bar`); the rest are copied cooperatively.
// Returns after issuing; call stage_wait() before reading shared memory.
template <int NREQ>
__device__ __forceinline__ void stage_issue(const StageReq (&req)[NREQ], int nreq, unsigned long long* bar) {
    unsigned int tx = 0;
#pragma unroll
    for (int i = 0; i < NREQ; ++i) {
        if (i < nreq && req[i].bytes > 0) {
            const bool bulk = ((req[i].bytes & 15u) == 0) &&
                              ((reinterpret_cast<uintptr_t>(req[i].src) & 15u) == 0);
            if (bulk) tx += req[i].bytes;
        }
    }
    if (threadIdx.x == 0) {
        mbar_expect_tx(bar, tx);                     // tx may be 0: the arrive alone completes the phase
#pragma unroll
        for (int i = 0; i < NREQ; ++i) {
            if (i < nreq && req[i].bytes > 0) {
                const bool bulk = ((req[i].bytes & 15u) == 0) &&
                                  ((reinterpret_cast<uintptr_t>(req[i].src) & 15u) == 0);
                if (bulk) bulk_g2s(req[i].dst, req[i].src, req[i].bytes, bar);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < NREQ; ++i) {
        if (i < nreq && req[i].bytes > 0) {
            const bool bulk = ((req[i].bytes & 15u) == 0) &&
                              ((reinterpret_cast<uintptr_t>(req[i].src) & 15u) == 0);
            if (!bulk) {                              // rare: ragged tail tile or unaligned view
                const unsigned int nw = req[i].bytes >> 2;
                const float* s = reinterpret_cast<const float*>(req[i].src);
                float* d = reinterpret_cast<float*>(req[i].dst);
                for (unsigned int j = threadIdx.x; j < nw; j += blockDim.x) d[j] = __ldg(s + j);
            }
        }
    }
}
__device__ __forceinline__ void stage_wait(unsigned long long* bar, unsigned int parity) {
    mbar_wait(bar, parity);
    __syncthreads();                                  // also publishes the cooperative copies
}

}  // namespace apgs
