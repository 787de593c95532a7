// fpc_umma.cuh -- the tcgen05 (5th-generation tensor core) plumbing of the bf16 predictor:
// TMEM allocation, shared-memory matrix descriptors, the instruction descriptor, MMA issue,
// commit -> mbarrier, TMEM -> register loads.  sm_100a only.
//
// Operand layout in shared memory ("K-major, no swizzle", the canonical interleaved form):
// a tile of R rows (M or N index) x K bf16 is stored as 8 x 16-byte core matrices,
//     byte offset(r, k) = (k / 8) * (R * 16) + r * 16 + (k % 8) * 2
// i.e. all rows of one 8-wide k-chunk are contiguous (16 B per row), chunks follow each other.
// In the descriptor: leading-dimension byte offset (K direction, chunk to chunk) = R * 16,
// stride-dimension byte offset (M/N direction, 8-row group to group) = 128.
// One tcgen05.mma.kind::f16 consumes K = 16 (two chunks); advancing the start address by
// 2 * R * 16 bytes steps to the next K = 16 slab.
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

#include "fpc_common.cuh"

namespace fpc {
namespace umma {

// ---- TMEM ------------------------------------------------------------------------------
// one full warp; writes the TMEM base address (lane 0, column c) to *smem_dst
__device__ __forceinline__ void tmem_alloc(uint32_t *smem_dst, uint32_t ncols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy shared-memory writes -> visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- descriptors ---------------------------------------------------------------------------
// shared-memory matrix descriptor, K-major, SWIZZLE_NONE, rows = R of the tile
__device__ __forceinline__ uint64_t smem_desc(uint32_t smem_addr, uint32_t rows)
{
    const uint64_t lbo = (uint64_t)(rows * 16u) >> 4;   // K direction: next 8-wide chunk
    const uint64_t sbo = 128u >> 4;                      // M/N direction: next 8-row group
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (lbo << 16) | (sbo << 32) | (1ull << 46);   // version 1 (sm_100)
}
// instruction descriptor: bf16 x bf16 -> f32, A and B K-major, dense
__host__ __device__ constexpr uint32_t instr_desc_bf16(uint32_t M, uint32_t N)
{
    return (1u << 4)            // D format f32
           | (1u << 7)          // A format bf16
           | (1u << 10)         // B format bf16
           | ((N >> 3) << 17)   // N / 8
           | ((M >> 4) << 24);  // M / 16
}

// D[tmem] (+)= A[smem] * B[smem]^T, one K = 16 slab.  One thread issues.
__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// The same, executed by a whole converged warp: one elected lane issues.  The issuing code then stays warp-uniform, so
// the compiler keeps the descriptors on the uniform datapath; issued from an `if (lane == 0)` region every operand goes
// through an ELECT / R2UR.BROADCAST sequence first and one MMA costs ~150 cycles of issue time (measured: the 330 MMAs
// of a frame took as long as the whole GRU phase).
__device__ __forceinline__ void mma_bf16_elect(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void commit_elect(uint64_t *bar)
{
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
        ::"r"(smem_u32(bar)) : "memory");
}
// all MMAs issued so far by this thread arrive on the mbarrier when they have completed
__device__ __forceinline__ void commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- TMEM -> registers: warp w reads lanes 32*(w%4)..+31, thread t its own lane, 16 columns ----
// The destination registers are undefined until tcgen05.wait::ld; ld16_wait() names them as
// read-write operands of the wait so the compiler cannot move a use above it.
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16_wait(uint32_t (&r)[16])
{
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :
                 : "memory");
}

// byte offset of element (r, k) of an R-row operand tile (see the header comment)
__host__ __device__ constexpr uint32_t tile_off(uint32_t r, uint32_t k, uint32_t R)
{
    return (k >> 3) * (R * 16u) + r * 16u + (k & 7u) * 2u;
}

}  // namespace umma
}  // namespace fpc
