// Bulk asynchronous copies (the TMA engine's 1-D mode, cp.async.bulk) and the mbarriers that track them, for sm_100a.
// A regular, contiguous tile -- a channel row's input window, a tile of output samples -- moves global <-> shared
// memory as ONE instruction issued by one thread; the data path bypasses the register file and the LSU pipes, and the
// consumer threads learn that the bytes have landed from the mbarrier's transaction count.
//   global -> shared : SASS UBLKCP.S.G, completion SYNCS.ARRIVE.TRANS64 / SYNCS.PHASECHK.TRANS64.TRYWAIT
//   shared -> global : SASS UBLKCP.G.S, bulk async-groups
// Addresses and sizes must be multiples of 16 bytes.
#pragma once
#include <cstdint>

namespace sdrgpu {
namespace tma {

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int arrivals)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(arrivals) : "memory");
}

// makes the initialised barriers visible to the async proxy (the copy engine) before the first copy names them
__device__ __forceinline__ void fence_barrier_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

// orders this thread's generic-proxy shared-memory accesses with later async-proxy ones (bulk copies reading or
// overwriting shared memory that ordinary loads / stores have touched)
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// one arrival + the number of bytes the copies issued next will deliver
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}

// global -> shared, completion counted on `bar`
__device__ __forceinline__ void bulk_load(void *dst_shared, const void *src_global, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst_shared)),
                 "l"(src_global), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}

// blocks until the barrier's phase with the given parity has completed (hardware sleep, not a spin on memory)
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(smem_addr(bar)),
        "r"(parity)
        : "memory");
}

// shared -> global as a bulk async-group
__device__ __forceinline__ void bulk_store(void *dst_global, const void *src_shared, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_global), "r"(smem_addr(src_shared)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// until the committed stores have finished READING shared memory (the source may be overwritten)
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

}  // namespace tma
}  // namespace sdrgpu
