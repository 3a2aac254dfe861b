// Tensor memory (TMEM, 256 KB per SM on sm_100a) used as a thread-private backing store.
//
// The resident Lenia kernel keeps two per-thread arrays that no other thread ever touches: the world state (64 floats
// per thread) and the kernel-spectrum multipliers (64 floats per thread).  With the 32x32b access shape a warp reads and
// writes its own 32 TMEM lanes (lane quarter = warp id % 4), one 32-bit column per register, so TMEM behaves as a second
// register file of 512 columns per lane quarter: no bank conflicts, no shared-memory bandwidth, and the 128 KB of shared
// memory the two arrays used to take are what lets two worlds (two CTAs) share an SM.  No tensor-core instruction is
// involved.  tools/tmem_probe.cu measures the path (0 mismatches, ~470-600 B/cycle/SM with a wait per chunk).
//
// Every ld is followed by tmem_wait_ld*() that takes the loaded registers as in/out operands: tcgen05.ld is asynchronous
// and the compiler must not schedule a consumer above the wait.
#pragma once
#include <cstdint>

namespace lnx {
namespace tm {

__device__ __forceinline__ void alloc(uint32_t* smem_dst, uint32_t ncols) {  // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void dealloc(uint32_t addr, uint32_t ncols) {  // the warp that allocated
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// address of column `col` of the calling warp's lane quarter
__device__ __forceinline__ uint32_t warp_addr(uint32_t base, int warp, int col) {
    return base + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)col;
}

__device__ __forceinline__ void ld8(uint32_t a, float* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                 : "r"(a)
                 : "memory");
}
__device__ __forceinline__ void st8(uint32_t a, const float* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(a), "f"(v[0]), "f"(v[1]), "f"(v[2]),
                 "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
                 : "memory");
}
__device__ __forceinline__ void st4(uint32_t a, float4 v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// wait for the loads issued so far; the registers of the chunk(s) about to be consumed pass through the statement
__device__ __forceinline__ void wait_ld8(float* v) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]), "+f"(v[4]), "+f"(v[5]), "+f"(v[6]), "+f"(v[7])::"memory");
}
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

}  // namespace tm
}  // namespace lnx
