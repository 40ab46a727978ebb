// Bulk asynchronous copies (the TMA engine's cp.async.bulk: SASS UBLKCP) and the mbarrier handshakes around them,
// shared by the streaming kernels (prep_onepass.cu, qmap_fold.cu).
//
// The streaming kernels are issue-bound when every thread loads and stores its own values (address arithmetic,
// register double buffers and their copies are most of the issue stream).  With bulk copies one elected thread
// moves whole row segments (128 - 256 cells of one day) between global and shared memory; the arithmetic warps
// only see shared memory at compile-time offsets.
//
// Every wait is bounded: a handshake that never completes traps (the launch then reports an error) instead of
// hanging the device.
#pragma once
#include <cstdint>

namespace attrici {
namespace bulk {

__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(void* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
// makes the initialised barriers visible to the async proxy (before the first bulk copy names them)
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

__device__ __forceinline__ void mbar_expect_tx(void* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(void* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(void* bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok) : "r"(smem_addr(bar)), "r"(parity), "r"(0x989680u) : "memory");   // suspend-time hint: the thread sleeps in hardware
    return ok != 0;
}
// waits for the completion of the barrier phase with the given parity
__device__ __forceinline__ void mbar_wait(void* bar, unsigned parity) {
    if (mbar_try_wait(bar, parity)) return;
    unsigned long long t0, t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while (!mbar_try_wait(bar, parity)) {
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        if (t - t0 > 4000000000ull) __trap();   // 4 s: no legitimate wait of these kernels lasts milliseconds
    }
}

// global -> shared; `bytes` a multiple of 16, both addresses 16-byte aligned; completes `bytes` of the barrier's
// transaction count
__device__ __forceinline__ void g2s(void* dst, const void* src, unsigned bytes, void* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_addr(dst)), "l"(src), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}
// shared -> global, tracked by the issuing thread's bulk async-group
__device__ __forceinline__ void s2g(void* dst, const void* src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_addr(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all bulk groups of this thread have finished READING their shared-memory source (the stage may be rewritten)
__device__ __forceinline__ void wait_group_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// shared-memory writes of this thread become visible to the async proxy (before a bulk store reads them)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

}  // namespace bulk
}  // namespace attrici
