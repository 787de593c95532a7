// fpc_common.cuh -- status codes, launch accounting and the few PTX wrappers the kernels use.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/fpc_b200.h"

namespace fpc {

extern thread_local int g_last_cuda_error;
extern unsigned long long g_launch_count;

inline int cuda_fail(cudaError_t e)
{
    g_last_cuda_error = (int)e;
    return FPC_ERR_CUDA;
}

#define FPC_CUDA_TRY(expr)                                   \
    do {                                                     \
        cudaError_t _e = (expr);                             \
        if (_e != cudaSuccess) return ::fpc::cuda_fail(_e);  \
    } while (0)

// call right after a <<<>>> launch
#define FPC_LAUNCH_CHECK()                                   \
    do {                                                     \
        __atomic_fetch_add(&::fpc::g_launch_count, 1ULL, __ATOMIC_RELAXED); \
        cudaError_t _e = cudaGetLastError();                 \
        if (_e != cudaSuccess) return ::fpc::cuda_fail(_e);  \
    } while (0)

// Per-device one-time state (kernel attributes, SM count, copy pipelines) is indexed by the CUDA device
// ordinal: cudaFuncSetAttribute and the SM count are per device, and one process may drive several GPUs
// through the C ABI.  Ordinals beyond the table share the last slot (attributes are then set every launch).
constexpr int kMaxDevices = 64;
inline int device_slot()
{
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0) d = 0;
    return d < kMaxDevices ? d : kMaxDevices - 1;
}
// opt-in to more than 48 KB of dynamic shared memory, once per (kernel instantiation, device)
template <typename Kernel>
inline int ensure_dynamic_smem(Kernel kernel, int bytes, bool (&done)[kMaxDevices])
{
    const int slot = device_slot();
    if (!done[slot] || slot == kMaxDevices - 1) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
        if (e != cudaSuccess) return cuda_fail(e);
        done[slot] = true;
    }
    return FPC_OK;
}

// ------------------------------------------------------------------------------------------
// packed images (device memory, produced by fpc_pack_*; layouts documented in DESIGN.md 4)
// ------------------------------------------------------------------------------------------
constexpr int kIn = FPC_IN_FEATURES;   // 20
constexpr int kH1 = FPC_GRU1;          // 384
constexpr int kH2 = FPC_GRU2;          // 128
constexpr int kFc = FPC_FC;            // 18
constexpr int kDim = FPC_CODE_DIMS;    // 17
constexpr int kSurv = FPC_SURVIVORS;   // 5

// fp32 weight stream: "groups" of kGk consecutive k for the 384 gate columns of one pass
// (128 hidden units x {r,z,n}); group = [3 gates][kGk/2 k-pairs][64 unit pairs][(e0,k) (e1,k) (e0,k+1) (e1,k+1)]
// floats: a warp reads 128 contiguous bytes per (gate, k-pair), and the two hidden units of a thread sit in one
// 64-bit register pair -- the operand form of the packed fma.rn.f32x2 (SASS FFMA2) the gate GEMM issues.  kGk = 8 halves the per-group hand-off cost (mbarrier wait, LDS latency
// exposed at the group boundary) relative to 4; the x part of GRU 1 is padded from 20 to 24 inputs with zero
// weights (fma(0, 0, acc) leaves the canonical chain's value unchanged).
constexpr int kGk = 8;
constexpr int kGroupFloats = 6 * 64 * kGk;               // 3072
constexpr int kGroupBytes = kGroupFloats * 4;            // 12288
constexpr int kG1x = (kIn + kGk - 1) / kGk;              // 3 groups: GRU1 input part (20 -> 24)
constexpr int kG1h = kH1 / kGk;                          // 48 groups: GRU1 hidden part
constexpr int kG1 = kG1x + kG1h;                         // 51 per pass, 3 passes
constexpr int kG2x = kH1 / kGk;                          // 48: GRU2 input part (h1')
constexpr int kG2h = kH2 / kGk;                          // 16: GRU2 hidden part
constexpr int kG2 = kG2x + kG2h;                         // 64
constexpr int kGroupsPerFrame = 3 * kG1 + kG2;           // 217
constexpr int kStreamFloats = kGroupsPerFrame * kGroupFloats;
constexpr int kBiasFloats = 4 * 4 * 128;                 // [3 GRU1 passes + GRU2][br,bz,bni,bnh][128]
constexpr int kFcFloats = kFc * kH2;                     // 2304
constexpr int kPackedF32Floats = ((kStreamFloats + kBiasFloats + kFcFloats + kFc + 3) / 4) * 4;
// Every CTA re-reads the whole weight image every frame, nearly in lock-step, so each L2 line of a single
// image takes ~150 back-to-back requests.  The image is therefore stored kWeightReplicas times (still
// L2-resident) and CTA b streams replica b % kWeightReplicas, spreading the requests over more L2 lines.
constexpr int kWeightReplicas = 8;
constexpr size_t kPackedF32ReplicaBytes = (((size_t)kPackedF32Floats * 4 + 255) / 256) * 256;

struct PackedVq {        // one VQ codebook file
    int dtype, stages, K, Kp;   // Kp = K rounded up to a multiple of 4
    long long off_t[2];   // byte offset of stage s transposed [17][K]   (file dtype)
    long long off_r[2];   // byte offset of stage s row-major  [K][17]   (file dtype)
    // screening data (fpc_vq_screen.cuh), always fp32:
    long long off_f[2];   // stage s transposed shadow [17][Kp], (float)c, zero padded
    long long off_n[2];   // stage s squared norms [Kp], padded with 3e38
    long long off_g;      // Gram table [K][Kp] = 2 * <c0_k0, c1_k1>   (2-stage books)
    long long off_cmax;   // float[2]: upper bounds of max_k ||c_k|| per stage
    // tensor-core screen (fpc_tc.cuh, fpc_vq_tc.cuh): per stage the fp16-pair B operand image, Kp64 rows in tiles of 64
    // codewords (8 KB each, four 2 KB K-slabs), scaled by the power of two beta_s = ((float *)(base + off_tcbeta))[s]
    // The images are stored kWeightReplicas times (every CTA streams them every frame: one copy would be an L2 hot
    // spot, exactly as for the weights); replica r starts b_rep_stride bytes after replica r - 1, and the stages of
    // one replica are contiguous.
    long long off_b[2];
    long long off_tcbeta;
    long long b_rep_stride;
    int Kp64, pad_;
};
struct PackedScl {
    int dtype, n;
    long long off;
};
struct PackedCodebooks {  // header at offset 0 of the packed codebook image
    PackedVq vq, bl;
    PackedScl scl, blscl;
};
constexpr size_t kCbHeaderBytes = 512;
constexpr size_t kCbVqMaxBytes = (size_t)2 * FPC_MAX_VQ_ENTRIES * kDim * 8 * 2    // both layouts, f64
                                 + (size_t)2 * (kDim + 1) * FPC_MAX_VQ_ENTRIES * 4   // fp32 shadows + norms
                                 + (size_t)8 * 2 * FPC_MAX_VQ_ENTRIES * 128          // tensor-core operand images, 8 replicas
                                 + (size_t)FPC_MAX_VQ_ENTRIES * FPC_MAX_VQ_ENTRIES * 4 + 8192;   // Gram table, cmax, alignment slack
constexpr size_t kCbSclMaxBytes = (size_t)FPC_MAX_SCL_ENTRIES * 8;
constexpr size_t kPackedCbBytes = kCbHeaderBytes + 2 * kCbVqMaxBytes + 2 * kCbSclMaxBytes;

// ------------------------------------------------------------------------------------------
// PTX: mbarrier + 1-D bulk async copy (TMA engine, SASS UBLKCP)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// non-blocking probe of a phase (mbarrier.test_wait never suspends the thread)
__device__ __forceinline__ bool mbar_test_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Spin on the phase.  A hand-off that never completes is a protocol bug: after ~8 s the thread traps,
// so the launch fails with an error instead of hanging the device.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 16000000000LL) __trap();
    }
}
// (Round 2, measured and rejected: the same loop around try_wait with a long suspend-time hint, so that waiting warps
// park instead of polling -- in the role-split fp32 kernel the polling loops are 40 % of all executed warp-instructions.
// It did not make the gate GEMM faster, and on the barriers of the bulk-copy weight ring it produced wrong results under
// full load: weights of a neighbouring group were read now and then (results differed by ~1e-5 between tiles holding the
// same utterances), while the same build with the plain loop on those two barriers was bit-exact.  Not understood; the
// hint is only used where it has always been, on the `go` barrier of the VQ helper roles, fpc_vq_tc.cuh.)
// global -> shared bulk copy, completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(smem_dst)),
        "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
// packed fp32 FMA on a register pair: d.x = fma(a.x, b, c.x), d.y = fma(a.y, b, c.y), each lane IEEE round-to-nearest
// (PTX fma.rn.f32x2, SASS FFMA2 with the scalar-broadcast operand form).  Half the issue slots of two FFMAs and the
// 64-bit operands straddle both register banks.
__device__ __forceinline__ float2 fma2(float2 a, float b, float2 c)
{
    unsigned long long ra, rb, rc, rd;
    asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b), "f"(b));
    asm("mov.b64 %0, {%1, %2};" : "=l"(rc) : "f"(c.x), "f"(c.y));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    float2 d;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
    return d;
}

// d = a * b + c on both lanes with a full pair for b
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c)
{
    unsigned long long ra, rb, rc, rd;
    asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(rc) : "f"(c.x), "f"(c.y));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    float2 d;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
    return d;
}
// d = a - b on both lanes (add.rn.f32x2 with the second operand negated; exact IEEE subtraction per lane)
__device__ __forceinline__ float2 sub2(float2 a, float2 b)
{
    unsigned long long ra, rb, rd;
    asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(-b.x), "f"(-b.y));
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    float2 d;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
    return d;
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

}  // namespace fpc
