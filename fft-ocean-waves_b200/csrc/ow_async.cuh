// ow_async.cuh — the asynchronous-copy machinery of sm_90+/sm_100a the frame kernels stage their inputs with: mbarriers with
// transaction counts, 1-D bulk copies (cp.async.bulk -> SASS UBLKCP) and 2-D tensor-map copies (cp.async.bulk.tensor -> SASS
// UTMALDG), L2 eviction-priority policies. Thin inline-PTX wrappers; device only.
//
// Why: the row and column kernels replace the reference's 2*log2(N) full-grid butterfly dispatches (src/main.cpp:626-661) by
// one pass each, and ncu showed both passes stalled on the latency of their global loads (long_scoreboard), not on bandwidth.
// A bulk copy is issued by ONE thread, needs no registers for the data in flight and signals an mbarrier when the bytes have
// landed in shared memory, so a persistent CTA can have its NEXT tile's input in flight while all its warps transform the current one.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace ow {

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
// Make freshly initialised barriers visible to the async proxy (the copy engines) before the first copy names them.
__device__ __forceinline__ void mbar_init_fence() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// One arrival + the number of bytes the copies issued next will deliver.
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Blocks until the phase with the given parity has completed. try_wait suspends the thread in hardware for a bounded time per call;
// a copy that never completes (a malformed tensor map, a wrong byte count) must not hang the device: after ~seconds of polling
// the kernel traps, which the host sees as a launch failure (cudaErrorLaunchFailure) instead of a wedged GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    unsigned spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 24)) __trap();
    }
}

__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}

// 1-D bulk copy global -> shared, completion on `bar`. dst, src and bytes are multiples of 16.
__device__ __forceinline__ void bulk_load(void* dst, const void* src, unsigned bytes, uint64_t* bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
                 : "memory");
}

// 2-D tensor-map copy global -> shared: the box of `tmap` whose first element is (x, y) (x = innermost coordinate, in elements).
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tmap, int x, int y, uint64_t* bar, uint64_t policy) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;" ::"r"(
                     smem_u32(dst)),
                 "l"(tmap), "r"(x), "r"(y), "r"(smem_u32(bar)), "l"(policy)
                 : "memory");
}
__device__ __forceinline__ void tma_prefetch_descriptor(const CUtensorMap* tmap) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}

// Barrier among the `nthreads` threads (a multiple of 32) that use barrier resource `id` (1..15; 0 is __syncthreads).
__device__ __forceinline__ void named_barrier(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
#endif

}  // namespace ow
