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
// shared -> global bulk store of a finished tile.  Called by ALL threads after they wrote `src_smem`
// with ordinary stores; one thread hands the slice to the bulk-copy engine (UBLKCP) when it is
// 16-byte sized/aligned, otherwise the CTA stores it cooperatively.
__device__ __forceinline__ void tile_store(void* dst_gmem, const void* src_smem, unsigned int bytes) {
    const bool bulk = ((bytes & 15u) == 0) && ((reinterpret_cast<uintptr_t>(dst_gmem) & 15u) == 0);
    if (bulk) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> async proxy
        __syncthreads();
        if (threadIdx.x == 0 && bytes > 0) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem),
                         "r"(smem_u32(src_smem)), "r"(bytes)
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // smem may be released at exit
        }
    } else {
        __syncthreads();
        const unsigned int nw = bytes >> 2;
        const float* s = reinterpret_cast<const float*>(src_smem);
        float* d = reinterpret_cast<float*>(dst_gmem);
        for (unsigned int j = threadIdx.x; j < nw; j += blockDim.x) d[j] = s[j];
    }
}

__device__ __forceinline__ void stage_wait(unsigned long long* bar, unsigned int parity) {
    mbar_wait(bar, parity);
    __syncthreads();                                  // also publishes the cooperative copies
}

}  // namespace apgs
