// fpc_encode_fp32.cu -- the fused, persistent closed-loop frame-step kernel (fp32 FFMA path).
//
// Replaces the whole body of Wavernn.encoder (/root/reference/src/models/wavernn.py:165-256):
// per frame  forward (:194 -> :63-102), residual (:196), thresholds (:202-208), the injected
// scl_quantize / vq_quantize calls (:217-240 -> quantization/vq_func.py), the feedback (:242 or
// :252), for a tile of MT = 4*TU utterances per CTA with the frame loop INSIDE the kernel.
// Nothing leaves the SM between frames except the per-frame outputs; the GRU states, the
// decoded frame that is fed back and all quantiser scratch live in shared memory.
//
// The loop of one frame is a dependency chain  x(t) -> GRU 1 -> GRU 2 -> FC -> quantiser -> x(t+1),  but two thirds of
// the predictor's arithmetic -- the hidden part W_hh h1(t) of GRU 1 -- only needs h1(t), which is known before the
// quantiser of frame t starts.  The CTA is therefore split by ROLE (512 threads = 4 warpgroups, registers moved between
// them with setmaxnreg):
//   warps 0..7   gate GEMMs.  Per frame: [wait for x(t)]  input part of GRU 1 (3 groups of 8 k per pass) added to the
//                parked hidden part, gate epilogue -> h1(t) in place;  GRU 2 -> h2(t), hand-over to the tail;  then the
//                hidden part of GRU 1 for frame t+1 (144 groups) while the tail quantises frame t; its r / z / n_h
//                partial sums are parked in TENSOR MEMORY (tcgen05.st, 18 TU columns per thread; TMEM is otherwise
//                idle in an fp32 kernel and shared memory has no room for 129 KB of partial sums).
//                Thread (tg = warp>>1, ug = (warp&1)*32 + lane) owns hidden units {2ug, 2ug+1} (+128*pass) for
//                utterances {tg, tg+4, ..., tg+4(TU-1)}; every dot product is one ascending-k FFMA chain seeded with
//                the bias, hidden part first, then input part (the canonical order of the oracle, fpc_oracle.c).
//   warps 8..11  the tail of the frame: FC + 2 tanh, residual, thresholds, scalar quantiser, the m-best VQ with its
//                distance screen on the tensor cores (fpc_vq_tc.cuh; 128 searching threads, one warp per TMEM lane
//                quarter), feedback, outputs.
//   warp 12      one lane streams the 2.68 MB packed weight image (217 groups of 12 KB per frame, in the order the
//                GEMM warps consume them) through a 4-stage ring with 1-D bulk async copies (UBLKCP).
//   warp 13      one lane streams the codebook operand images of the VQ screen;  warps 14, 15 issue its MMAs.
// GEMM warps and tail meet at two mbarriers per frame (x_ready, h2_ready); h2 is double-buffered, h1 is updated in place.
// A launch covers the frame range [f0, f1) of every utterance; with EncodeParams::state the recurrent state of each
// tile is carried from one launch to the next (fpc_encode_host cuts a batch along time that way).
#include <cstdlib>

#include "fpc_common.cuh"
#include "fpc_math.cuh"
#include "fpc_vq.cuh"
#include "fpc_vq_search.cuh"
#define FPC_VQ_TC_OUTLINE __forceinline__      // see fpc_vq_tc.cuh
#include "fpc_vq_tc.cuh"
#include "fpc_encode.cuh"

namespace fpc {

constexpr int kStages = 4;             // weight ring depth
static_assert(kStages <= 8, "h1 is updated in place: a GEMM warp may run at most kStages groups ahead of the slowest");
constexpr int kTailThreads = 128;
constexpr int kThreads = kComputeThreads + kTailThreads + 128;   // 2 GEMM warpgroups + tail warpgroup + helper warpgroup
constexpr int kTailWarp0 = kComputeThreads / 32;                 // 8
constexpr int kHelpWarp0 = kTailWarp0 + kTailThreads / 32;       // 12
// registers per thread after setmaxnreg; the launch allocates 512 x 128.  The tail keeps its 128: it calls out-of-line
// functions (the VQ search), and those are compiled against the launch allocation.
constexpr int kRegGemm = 176, kRegTail = 128, kRegHelp = 32;      // (literal in the setmaxnreg instructions below)
static_assert(kComputeThreads * kRegGemm + kTailThreads * kRegTail + 128 * kRegHelp <= 65536, "register file");
constexpr int kLd1 = kH1 + 4;          // 388: padded row strides (floats) -> conflict-free float4 rows
constexpr int kLd2 = kH2 + 4;          // 132
constexpr int kLdX = 24;               // input frame row (20 used)
constexpr int kLdFc = kH2 + 4;         // 132: float4-aligned rows, conflict-free for 8 consecutive rows
constexpr int kVqNB = 2;               // codebook ring of the VQ screen, 8 KB chunks

// tensor memory: parked hidden-part sums of the two warp sets (warps 0..3 / 4..7 share the lane quarters), then the
// accumulator units of the VQ screen
template <int TU> struct Tm {
    static constexpr int kArr = 2 * TU;                    // columns of one gate array (TU unit pairs)
    static constexpr int kPassCols = 3 * kArr;             // r, z, n_h
    static constexpr int kSetCols = 3 * kPassCols;         // three passes
    static constexpr int kParkCols = ((2 * kSetCols + 63) / 64) * 64;
    static constexpr int kUnits = (512 - kParkCols) / 64;  // 64-column units left for the screen
    static_assert(kUnits >= 2, "tensor memory");
};

template <int TU> struct Smem {
    static constexpr int MT = 4 * TU;
    static constexpr int kMtMax = (5 * MT + 127) / 128;
    static constexpr int offRing = 0;
    static constexpr int offH1 = offRing + kStages * kGroupBytes;
    static constexpr int offH2a = offH1 + MT * kLd1 * 4;
    static constexpr int offH2b = offH2a + MT * kLd2 * 4;
    static constexpr int offXin = offH2b + MT * kLd2 * 4;
    static constexpr int offBias = offXin + MT * kLdX * 4;
    static constexpr int offFc = offBias + kBiasFloats * 4;              // 18 x 129 weights + 18 bias
    static constexpr int offRs = offFc + ((kFc * kLdFc + kFc + 3) / 4) * 16;
    static constexpr int offRq = offRs + MT * kLdR * 4;                  // quantised residual rows (stride 20)
    static constexpr int offMisc = offRq + MT * 20 * 4;
    // misc: m1[MT] m2[MT] (float), idx0/idx1/idx2[MT] (int), listA[MT] listB[MT] tail[MT] (int), counts[4], tmem slot, pad
    static constexpr int offScl = ((offMisc + (8 * MT + 8) * 4 + 15) / 16) * 16;   // both scalar tables, file dtype, 2 x 2 KB
    static constexpr int offBars = offScl + 2 * FPC_MAX_SCL_ENTRIES * 8 + 16;   // + {n, dtype} of the two scalar tables
    static constexpr int kNumBars = 2 * kStages + 2;
    static constexpr int offVqSh = ((offBars + kNumBars * 8 + 127) / 128) * 128;
    // the transient blocks of the screen are contiguous: together they are the scratch of the exact fallback search
    static constexpr int offVqSmall = offVqSh + 512;
    static constexpr int offVqA = offVqSmall + ((vq_tc_small_bytes(MT) + 127) / 128) * 128;
    static constexpr int offPart = offVqA + kMtMax * tc::kTileBytes;
    static constexpr int offBring = offPart + ((vq_tc_part_bytes(kMtMax) + 127) / 128) * 128;
    static constexpr int kScratchBytes = offBring - offVqSmall;
    static constexpr int total = offBring + kVqNB * kVtChunkBytes;
    static constexpr int kStateFloats = MT * (kLd1 + kLd2);              // carried [h1 | h2] (+ MT * kLdX of the input frame)
    static_assert(total <= 227 * 1024, "shared memory");
};

// ring pipeline state of a compute thread
struct Pipe {
    int s;
    uint32_t ph;
    __device__ __forceinline__ void advance()
    {
        if (++s == kStages) { s = 0; ph ^= 1u; }
    }
};

// ---- tensor memory <-> registers, N consecutive columns of this thread's lane (N a sum of 16 / 8 / 4 / 2) ----
template <int N> struct TmIo;
template <> struct TmIo<2> {
    static __device__ __forceinline__ void st(uint32_t ta, const uint32_t *r)
    {
        asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1,%2};" ::"r"(ta), "r"(r[0]), "r"(r[1]) : "memory");
    }
    static __device__ __forceinline__ void ld(uint32_t ta, uint32_t *r)
    {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(ta) : "memory");
    }
};
template <> struct TmIo<4> {
    static __device__ __forceinline__ void st(uint32_t ta, const uint32_t *r)
    {
        asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(ta), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
    }
    static __device__ __forceinline__ void ld(uint32_t ta, uint32_t *r)
    {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(ta) : "memory");
    }
};
template <> struct TmIo<8> {
    static __device__ __forceinline__ void st(uint32_t ta, const uint32_t *r)
    {
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(ta), "r"(r[0]), "r"(r[1]), "r"(r[2]),
                     "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                     : "memory");
    }
    static __device__ __forceinline__ void ld(uint32_t ta, uint32_t *r)
    {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                     : "r"(ta)
                     : "memory");
    }
};
template <> struct TmIo<16> {
    static __device__ __forceinline__ void st(uint32_t ta, const uint32_t *r)
    {
        asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(ta),
                     "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
                     "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
                     : "memory");
    }
    static __device__ __forceinline__ void ld(uint32_t ta, uint32_t *r)
    {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                       "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                     : "r"(ta)
                     : "memory");
    }
};
template <int N> __device__ __forceinline__ void tmem_st_n(uint32_t ta, const uint32_t *r)
{
    static_assert(N >= 0 && N % 2 == 0, "even column counts");
    if constexpr (N >= 16) { TmIo<16>::st(ta, r); tmem_st_n<N - 16>(ta + 16, r + 16); }
    else if constexpr (N >= 8) { TmIo<8>::st(ta, r); tmem_st_n<N - 8>(ta + 8, r + 8); }
    else if constexpr (N >= 4) { TmIo<4>::st(ta, r); tmem_st_n<N - 4>(ta + 4, r + 4); }
    else if constexpr (N >= 2) { TmIo<2>::st(ta, r); }
}
template <int N> __device__ __forceinline__ void tmem_ld_n(uint32_t ta, uint32_t *r)
{
    if constexpr (N >= 16) { TmIo<16>::ld(ta, r); tmem_ld_n<N - 16>(ta + 16, r + 16); }
    else if constexpr (N >= 8) { TmIo<8>::ld(ta, r); tmem_ld_n<N - 8>(ta + 8, r + 8); }
    else if constexpr (N >= 4) { TmIo<4>::ld(ta, r); tmem_ld_n<N - 4>(ta + 4, r + 4); }
    else if constexpr (N >= 2) { TmIo<2>::ld(ta, r); }
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// the registers of a load are undefined until tcgen05.wait::ld; naming them as read-write operands keeps every use below
template <int N> __device__ __forceinline__ void tmem_wait_ld(uint32_t (&a)[N], uint32_t (&b)[N], uint32_t (&c)[N])
{
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < N; ++i) asm volatile("" : "+r"(a[i]), "+r"(b[i]), "+r"(c[i]));
}

// ------------------------------------------------------------------------------------------
// one "part" of a pass: NG groups of 8 k, accumulating into r, z and the third gate (n_i or n_h)
// (Measured and rejected on the B200: weight tile of group g+1 prefetched into a second register buffer, barrier
// probed two groups ahead, whole 4-k chunks prefetched with the rows innermost, two groups per barrier round trip,
// two half-size CTAs per SM, all operands of group g+1 loaded at the end of iteration g (26 LDS, then 168 FFMA2 with
// nothing in between: 275 k cycles instead of 223 k, also with the two warps of a scheduler skewed by half a group)
// -- all equal or slower than this plain loop; ptxas orders loads and FMAs by operand
// readiness whatever the source says.  tools/gemm_bounds.sh: the weight stream alone needs 51 k cycles per frame,
// the arithmetic alone 214 k of the 223 k the stage takes.)
// ------------------------------------------------------------------------------------------
template <int TU>
__device__ __forceinline__ void gemm_part(float2 (&ar)[TU], float2 (&az)[TU], float2 (&an)[TU], int ng,
                                          const float *__restrict__ arow, int lda, const float4 *__restrict__ ring,
                                          uint64_t *full, uint64_t *empty, Pipe &pp, int ug, int lane, uint32_t zmask)
{
    constexpr int NP = kGk / 2;      // k-pairs per group
    for (int g = 0; g < ng; ++g) {
#ifndef FPC_DEBUG_NO_STREAM          // (debug builds only, tools/gemm_bounds.sh: the arithmetic without the weight stream etc.)
        mbar_wait(&full[pp.s], pp.ph);
#endif
        const float4 *sw = ring + pp.s * (kGroupFloats / 4) + ug;
        float4 w[3][NP];             // [gate][k-pair] = (e0,k) (e1,k) (e0,k+1) (e1,k+1)
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int q = 0; q < NP; ++q) w[c][q] = sw[(c * NP + q) * 64];
#ifndef FPC_DEBUG_STREAM_ONLY
#pragma unroll
        for (int i = 0; i < TU; ++i) {
#pragma unroll
            for (int q = 0; q < kGk / 4; ++q) {
                const float4 a = *reinterpret_cast<const float4 *>(arow + (size_t)(4 * i) * lda + kGk * g + 4 * q);
                float2 r = ar[i], z = az[i], n = an[i];
                // ascending k: the canonical chain of every accumulator
                r = fma2(make_float2(w[0][2 * q].x, w[0][2 * q].y), a.x, r);
                z = fma2(make_float2(w[1][2 * q].x, w[1][2 * q].y), a.x, z);
                n = fma2(make_float2(w[2][2 * q].x, w[2][2 * q].y), a.x, n);
                r = fma2(make_float2(w[0][2 * q].z, w[0][2 * q].w), a.y, r);
                z = fma2(make_float2(w[1][2 * q].z, w[1][2 * q].w), a.y, z);
                n = fma2(make_float2(w[2][2 * q].z, w[2][2 * q].w), a.y, n);
                r = fma2(make_float2(w[0][2 * q + 1].x, w[0][2 * q + 1].y), a.z, r);
                z = fma2(make_float2(w[1][2 * q + 1].x, w[1][2 * q + 1].y), a.z, z);
                n = fma2(make_float2(w[2][2 * q + 1].x, w[2][2 * q + 1].y), a.z, n);
                r = fma2(make_float2(w[0][2 * q + 1].z, w[0][2 * q + 1].w), a.w, r);
                z = fma2(make_float2(w[1][2 * q + 1].z, w[1][2 * q + 1].w), a.w, z);
                n = fma2(make_float2(w[2][2 * q + 1].z, w[2][2 * q + 1].w), a.w, n);
                ar[i] = r; az[i] = z; an[i] = n;
            }
        }
#endif
#ifndef FPC_DEBUG_NO_STREAM
        // The stage may be refilled as soon as all eight warps have arrived, so every weight load of this group must have
        // RETURNED before the arrive -- not merely have been issued.  ptxas orders the arrive against the other memory
        // operations only: with TU = 8 it sat directly behind the last LDS.128 of the group, 60 FFMA2 ahead of where the
        // source has it, and under full load a stage was now and then overwritten under a load in flight (weights of
        // group g + 4 instead of g: results off by ~1e-4 in one row group of a tile, a few times per launch).  The
        // barrier address is therefore made to depend on what was loaded, through a mask that is zero at run time
        // only (a constant zero is folded away and the dependence with it).
        // (Through the accumulators of the last row rather than the weight registers themselves: every LDS.128 of the group
        // feeds them, and the arrive then sits behind the FFMA2s instead of stalling the warp on the load latency.)
        uint32_t dep = (__float_as_uint(ar[TU - 1].x) | __float_as_uint(az[TU - 1].x) | __float_as_uint(an[TU - 1].x)) & zmask;
        __syncwarp();
        if (lane == 0) mbar_arrive(reinterpret_cast<uint64_t *>(reinterpret_cast<char *>(&empty[pp.s]) + dep));
#endif
        pp.advance();
    }
}

// hidden part of one GRU 1 pass (128 hidden units): bias + W_hh h1 for r, z, n_h -> parked in tensor memory
template <int TU>
__device__ __forceinline__ void gru1_hidden_pass(const float *__restrict__ hrow, const float *__restrict__ bias, uint32_t tpark,
                                                 const float4 *__restrict__ ring, uint64_t *full, uint64_t *empty, Pipe &pp, int ug, int lane,
                                                 uint32_t zmask)
{
    float2 ar[TU], az[TU], anh[TU];
    const float2 br = *reinterpret_cast<const float2 *>(bias + 0 * 128 + 2 * ug);
    const float2 bz = *reinterpret_cast<const float2 *>(bias + 1 * 128 + 2 * ug);
    const float2 bh = *reinterpret_cast<const float2 *>(bias + 3 * 128 + 2 * ug);
#pragma unroll
    for (int i = 0; i < TU; ++i) { ar[i] = br; az[i] = bz; anh[i] = bh; }
    gemm_part<TU>(ar, az, anh, kG1h, hrow, kLd1, ring, full, empty, pp, ug, lane, zmask);
    uint32_t v[3][2 * TU];
#pragma unroll
    for (int i = 0; i < TU; ++i) {
        v[0][2 * i] = __float_as_uint(ar[i].x); v[0][2 * i + 1] = __float_as_uint(ar[i].y);
        v[1][2 * i] = __float_as_uint(az[i].x); v[1][2 * i + 1] = __float_as_uint(az[i].y);
        v[2][2 * i] = __float_as_uint(anh[i].x); v[2][2 * i + 1] = __float_as_uint(anh[i].y);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) tmem_st_n<2 * TU>(tpark + c * (2 * TU), v[c]);
}

// input part of one GRU 1 pass and the gate epilogue: h1 is updated in place (a thread reads and writes only its own units)
template <int TU>
__device__ __forceinline__ void gru1_input_pass(const float *__restrict__ xrow, const float *__restrict__ bias, uint32_t tpark,
                                                float *__restrict__ hown, const float4 *__restrict__ ring, uint64_t *full, uint64_t *empty,
                                                Pipe &pp, int ug, int lane, uint32_t zmask)
{
    uint32_t vr[2 * TU], vz[2 * TU], vh[2 * TU];
    tmem_ld_n<2 * TU>(tpark, vr);
    tmem_ld_n<2 * TU>(tpark + 2 * TU, vz);
    tmem_ld_n<2 * TU>(tpark + 4 * TU, vh);
    tmem_wait_ld<2 * TU>(vr, vz, vh);
    float2 ar[TU], az[TU], ani[TU];
    const float2 bi = *reinterpret_cast<const float2 *>(bias + 2 * 128 + 2 * ug);
#pragma unroll
    for (int i = 0; i < TU; ++i) {
        ar[i] = make_float2(__uint_as_float(vr[2 * i]), __uint_as_float(vr[2 * i + 1]));
        az[i] = make_float2(__uint_as_float(vz[2 * i]), __uint_as_float(vz[2 * i + 1]));
        ani[i] = bi;
    }
    gemm_part<TU>(ar, az, ani, kG1x, xrow, kLdX, ring, full, empty, pp, ug, lane, zmask);
#pragma unroll
    for (int i = 0; i < TU; ++i) {
        float2 *hp = reinterpret_cast<float2 *>(hown + (size_t)(4 * i) * kLd1);
        const float2 ho = *hp;
        float2 hn;
        hn.x = gru_update(ar[i].x, az[i].x, ani[i].x, __uint_as_float(vh[2 * i]), ho.x);
        hn.y = gru_update(ar[i].y, az[i].y, ani[i].y, __uint_as_float(vh[2 * i + 1]), ho.y);
        *hp = hn;
    }
}

// GRU 2 (one pass of 128 units): hidden part, input part (the new h1), gate epilogue into the other h2 buffer
template <int TU>
__device__ __forceinline__ void gru2_pass(const float *__restrict__ xrow, const float *__restrict__ hrow, const float *__restrict__ hold,
                                          float *__restrict__ hnew, const float *__restrict__ bias, const float4 *__restrict__ ring,
                                          uint64_t *full, uint64_t *empty, Pipe &pp, int ug, int lane, uint32_t zmask)
{
    float2 ar[TU], az[TU], ani[TU], anh[TU];
    const float2 br = *reinterpret_cast<const float2 *>(bias + 0 * 128 + 2 * ug);
    const float2 bz = *reinterpret_cast<const float2 *>(bias + 1 * 128 + 2 * ug);
    const float2 bi = *reinterpret_cast<const float2 *>(bias + 2 * 128 + 2 * ug);
    const float2 bh = *reinterpret_cast<const float2 *>(bias + 3 * 128 + 2 * ug);
#pragma unroll
    for (int i = 0; i < TU; ++i) { ar[i] = br; az[i] = bz; ani[i] = bi; anh[i] = bh; }
    gemm_part<TU>(ar, az, anh, kG2h, hrow, kLd2, ring, full, empty, pp, ug, lane, zmask);
    gemm_part<TU>(ar, az, ani, kG2x, xrow, kLd1, ring, full, empty, pp, ug, lane, zmask);
#pragma unroll
    for (int i = 0; i < TU; ++i) {
        const float2 ho = *reinterpret_cast<const float2 *>(hold + (size_t)(4 * i) * kLd2);
        float2 hn;
        hn.x = gru_update(ar[i].x, az[i].x, ani[i].x, anh[i].x, ho.x);
        hn.y = gru_update(ar[i].y, az[i].y, ani[i].y, anh[i].y, ho.y);
        *reinterpret_cast<float2 *>(hnew + (size_t)(4 * i) * kLd2) = hn;
    }
}

// position s of the per-frame consumption order -> group of the packed image (fpc_pack.cu stores pass by pass, input
// part before hidden part): hidden parts of the three GRU 1 passes, their input parts, GRU 2 hidden, GRU 2 input
__device__ __forceinline__ int stream_group(int s)
{
    if (s < 3 * kG1h) { const int p = s / kG1h; return p * kG1 + kG1x + (s - p * kG1h); }
    s -= 3 * kG1h;
    if (s < 3 * kG1x) { const int p = s / kG1x; return p * kG1 + (s - p * kG1x); }
    s -= 3 * kG1x;
    return s < kG2h ? 3 * kG1 + kG2x + s : 3 * kG1 + (s - kG2h);
}

// ------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------
template <int TU>
__global__ void __launch_bounds__(kThreads, 1) encode_fp32_kernel(EncodeParams P)
{
    using S = Smem<TU>;
    using T = Tm<TU>;
    using VqSh = VqTcShared<kVqNB, T::kUnits>;
    constexpr int MT = S::MT;
    constexpr int NE = (MT * 20 + kTailThreads - 1) / kTailThreads;   // frame elements per tail thread
    extern __shared__ __align__(1024) unsigned char smem[];
    float4 *ring = reinterpret_cast<float4 *>(smem + S::offRing);
    float *h1 = reinterpret_cast<float *>(smem + S::offH1);
    float *h2buf[2] = {reinterpret_cast<float *>(smem + S::offH2a), reinterpret_cast<float *>(smem + S::offH2b)};
    float *xin = reinterpret_cast<float *>(smem + S::offXin);
    float *bias = reinterpret_cast<float *>(smem + S::offBias);
    float *wfc = reinterpret_cast<float *>(smem + S::offFc);
    float *bfc = wfc + kFc * kLdFc;
    float *rs = reinterpret_cast<float *>(smem + S::offRs);
    float *rq = reinterpret_cast<float *>(smem + S::offRq);
    float *m1s = reinterpret_cast<float *>(smem + S::offMisc);
    float *m2s = m1s + MT;
    int *idx0s = reinterpret_cast<int *>(m2s + MT);
    int *idx1s = idx0s + MT;
    int *idx2s = idx1s + MT;
    int *listA = idx2s + MT;
    int *listB = listA + MT;
    int *ftail = listB + MT;
    int *counts = ftail + MT;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(counts + 4);
    unsigned char *sclbuf = smem + S::offScl;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + S::offBars);
    uint64_t *empty = full + kStages;
    uint64_t *x_ready = empty + kStages;       // tail -> GEMM warps: the input frame (and everything else of the frame) is done
    uint64_t *h2_ready = x_ready + 1;          // GEMM warps -> tail: h2 of the frame is in its buffer
    VqSh *vsh = reinterpret_cast<VqSh *>(smem + S::offVqSh);
    static_assert(sizeof(VqSh) <= 512, "control block");

    const int tid = threadIdx.x;
    // The warp index through a shuffle: the compiler then KNOWS it is the same in all lanes.  With tid >> 5 every branch on
    // the warp's role counted as divergent, and inside such a region every tcgen05.mma was issued through ELECT + five
    // R2UR.BROADCAST (~200 cycles per MMA instead of ~10: the descriptors did not stay on the uniform datapath).
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
    const int my_tiles = P.ntiles > (int)blockIdx.x ? (P.ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int nframes = P.f1 - P.f0;

    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kComputeThreads / 32);
        }
        mbar_init(x_ready, 1);
        mbar_init(h2_ready, 1);
        vq_tc_init<kVqNB>(vsh, kTailThreads / 32);
        mbar_fence_init();
    }
    if (warp == kTailWarp0) umma::tmem_alloc(tmem_slot, 512);
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tb = *tmem_slot;
    const uint32_t tb_vq = tb + T::kParkCols;

    // ---------------- helper warpgroup: weight stream, codebook stream, MMA issue of the VQ screen ----------------
    if (warp >= kHelpWarp0) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");      // kRegHelp
        if (warp == kHelpWarp0) {
            if (lane == 0) {
                const long long total = (long long)my_tiles * nframes * kGroupsPerFrame;
                const char *src = reinterpret_cast<const char *>(P.wstream) + (size_t)(blockIdx.x % kWeightReplicas) * kPackedF32ReplicaBytes;
                int s = 0, gf = 0;
                uint32_t wraps = 0;
#ifdef FPC_DEBUG_NO_STREAM
                const long long ntotal = 0 * total;
#else
                const long long ntotal = total;
#endif
                for (long long g = 0; g < ntotal; ++g) {
                    if (wraps > 0) mbar_wait(&empty[s], (wraps - 1) & 1u);
                    mbar_arrive_expect_tx(&full[s], kGroupBytes);
                    bulk_g2s(smem + S::offRing + s * kGroupBytes, src + (size_t)stream_group(gf) * kGroupBytes, kGroupBytes, &full[s]);
                    if (++gf == kGroupsPerFrame) gf = 0;
                    if (++s == kStages) { s = 0; ++wraps; }
                }
            }
        } else if (P.mode == kModeQuantize) {
            // one publication with last = 1 per frame ends the frame for these roles (fpc_vq_tc.cuh)
            VqTcCount vn{0u, 0u, 0u, 0u};
            const long long frames = (long long)my_tiles * nframes;
            if (warp == kHelpWarp0 + 1) {
                if (lane == 0)
                    for (long long f = 0; f < frames; ++f) {
                        bool done = false;
                        while (!done) done = vq_tc_produce_phase<kVqNB>(vsh, P.cb, vn);
                    }
            } else {
                for (long long f = 0; f < frames; ++f) {
                    bool done = false;
                    while (!done) done = vq_tc_issue_phase<kVqNB>(vsh, tb_vq, vn, lane, warp - (kHelpWarp0 + 2));
                }
            }
        }
        return;
    }

    const PackedCodebooks *cbh = reinterpret_cast<const PackedCodebooks *>(P.cb);

    // ---------------- tail warps ----------------
    if (warp >= kTailWarp0) {
        const int ttid = tid - kComputeThreads;
        const int twarp = ttid >> 5;
        {
            const float *tail = P.wstream + kStreamFloats;
            for (int i = ttid; i < kFcFloats; i += kTailThreads) wfc[(i >> 7) * kLdFc + (i & 127)] = tail[kBiasFloats + i];
            if (ttid < kFc) bfc[ttid] = tail[kBiasFloats + kFcFloats + ttid];
            // the scalar tables (<= 256 levels each) next to the state: from global memory every scalar search paid L2 latencies
            if (cbh != nullptr) {
                const long long *src0 = reinterpret_cast<const long long *>(P.cb + cbh->scl.off);
                const long long *src1 = reinterpret_cast<const long long *>(P.cb + cbh->blscl.off);
                const int n0 = cbh->scl.n * (cbh->scl.dtype == FPC_F32 ? 4 : 8), n1 = cbh->blscl.n * (cbh->blscl.dtype == FPC_F32 ? 4 : 8);
                for (int i = ttid; i < (n0 + 7) / 8; i += kTailThreads) reinterpret_cast<long long *>(sclbuf)[i] = src0[i];
                for (int i = ttid; i < (n1 + 7) / 8; i += kTailThreads)
                    reinterpret_cast<long long *>(sclbuf + FPC_MAX_SCL_ENTRIES * 8)[i] = src1[i];
                if (ttid == 0) {
                    int *meta = reinterpret_cast<int *>(sclbuf + 2 * FPC_MAX_SCL_ENTRIES * 8);
                    meta[0] = cbh->scl.n; meta[1] = cbh->scl.dtype; meta[2] = cbh->blscl.n; meta[3] = cbh->blscl.dtype;
                }
            }
            if (ttid == 0) vsh->tmem_base = tb_vq;       // the VQ screen reads it from its control block
        }
        uint32_t phh = 0;
        const bool prof = P.prof != nullptr && ttid == 0;
        long long pt[kPhCount] = {}, pt0 = 0;
    #define FPC_PHASE(ph) do { if (prof) { const long long t_ = clock64(); pt[ph] += t_ - pt0; pt0 = t_; } } while (0)

        for (int tile = blockIdx.x; tile < P.ntiles; tile += gridDim.x) {
            const int b0 = tile * MT;
            float *carry = P.state ? reinterpret_cast<float *>(P.state) + (size_t)tile * (S::kStateFloats + MT * kLdX) : nullptr;
            if (carry && P.f0 > 0) {
                for (int i = ttid; i < MT * kLdX; i += kTailThreads) xin[i] = carry[S::kStateFloats + i];
            } else {
                for (int i = ttid; i < MT * kLdX; i += kTailThreads) xin[i] = 0.0f;      // frame 0 input is all zero
            }
            for (int i = ttid; i < MT * kLdR; i += kTailThreads) rs[i] = 0.0f;
            named_bar_sync(1, kTailThreads);
            if (ttid == 0) mbar_arrive(x_ready);
            int cur = 0;

            for (int fr = P.f0; fr < P.f1; ++fr) {
                // element ownership for this frame: e = ttid + 128 q -> (row u, feature j)
                float featv[NE], fov[NE], rsv[NE];
    #pragma unroll
                for (int q = 0; q < NE; ++q) {
                    const int e = ttid + kTailThreads * q;
                    const int u = e / 20, j = e - u * 20;
                    featv[q] = 0.0f; fov[q] = 0.0f; rsv[q] = 0.0f;
                    if (e < MT * 20 && b0 + u < P.B) {
                        const size_t fo = (size_t)(b0 + u) * P.L + fr;
                        if (P.mode != kModeDecode) featv[q] = __ldg(P.feat + fo * 20 + j);
                        else featv[q] = j < kFc ? __ldg(P.rq_in + fo * kFc + j) : __ldg(P.pitch_in + fo * 2 + (j - kFc));
                    }
                }
                if (prof) pt0 = clock64();
                mbar_wait(h2_ready, phh); phh ^= 1u;
                const float *h2n = h2buf[cur ^ 1];
                cur ^= 1;
                FPC_PHASE(kPhWaitH);
                // ---- relu, dual_fc, 2*tanh (wavernn.py:87-92); residual (:196) ----
                {
                    // the thread's NE outputs advance together (NE independent ascending-k chains, operands as float4)
                    float acc[NE];
                    const float4 *wr4[NE], *hv4[NE];
                    bool on[NE];
    #pragma unroll
                    for (int q = 0; q < NE; ++q) {
                        const int e = ttid + kTailThreads * q;
                        const int u = e / 20, j = e - u * 20;
                        on[q] = e < MT * 20 && j < kFc;
                        acc[q] = on[q] ? bfc[j] : 0.0f;
                        wr4[q] = reinterpret_cast<const float4 *>(wfc + (on[q] ? j : 0) * kLdFc);
                        hv4[q] = reinterpret_cast<const float4 *>(h2n + (on[q] ? u : 0) * kLd2);
                    }
    #pragma unroll 2
                    for (int k4 = 0; k4 < kH2 / 4; ++k4) {
    #pragma unroll
                        for (int q = 0; q < NE; ++q) {
                            const float4 w4 = wr4[q][k4], h4 = hv4[q][k4];
                            float a = acc[q];
                            a = __fmaf_rn(w4.x, fmaxf(h4.x, 0.0f), a);
                            a = __fmaf_rn(w4.y, fmaxf(h4.y, 0.0f), a);
                            a = __fmaf_rn(w4.z, fmaxf(h4.z, 0.0f), a);
                            a = __fmaf_rn(w4.w, fmaxf(h4.w, 0.0f), a);
                            acc[q] = a;
                        }
                    }
    #pragma unroll
                    for (int q = 0; q < NE; ++q) {
                        if (on[q]) {
                            const int e = ttid + kTailThreads * q;
                            const int u = e / 20, j = e - u * 20;
                            const float f = __fmul_rn(2.0f, tanh_c(acc[q]));
                            fov[q] = f;
                            if (P.mode != kModeDecode) {
                                rsv[q] = __fsub_rn(featv[q], f);
                                rs[u * kLdR + 3 + j] = rsv[q];
                            }
                        }
                    }
                }
                if (P.mode != kModeDecode) {
                    for (int i = ttid; i < MT * 20; i += kTailThreads) rq[i] = 0.0f;
                    named_bar_sync(1, kTailThreads);

                    FPC_PHASE(kPhFc);
                    // ---- indicators (:201-212) and the scalar quantiser for c0 (:217-225): four lanes per utterance ----
                    {
                        constexpr int G = 4;
                        const int u = ttid / G, part = ttid % G;          // u < 32 >= MT
                        const bool valid = u < MT && b0 + u < P.B;
                        const int uc = u < MT ? u : 0;
                        float m1 = 0.0f, m2 = 0.0f;
                        if (part == 0 && valid) {
                            if (P.mask == nullptr) {
                                float s = 0.0f;
    #pragma unroll
                                for (int j = 1; j < kFc; ++j) s = __fadd_rn(s, fabsf(rs[uc * kLdR + 3 + j]));
                                m1 = fabsf(rs[uc * kLdR + 3]) > P.l1 ? 1.0f : 0.0f;
                                m2 = s > P.l2 ? 1.0f : 0.0f;
                            } else {
                                const size_t fo = (size_t)(b0 + u) * P.L + fr;
                                m1 = __ldg(P.mask + fo * 2);
                                m2 = __ldg(P.mask + fo * 2 + 1);
                            }
                        }
                        m1 = __shfl_sync(0xffffffffu, m1, lane - part);
                        m2 = __shfl_sync(0xffffffffu, m2, lane - part);
                        int i0 = -1;
                        float qv = 0.0f;
                        bool coded = false;
                        if (P.mode == kModeQuantize) {            // (every lane takes part in the group's shuffles)
                            const int which = (m1 != 0.0f) ? 0 : 1;       // above / below threshold table (:217-225)
                            const unsigned char *sclt = sclbuf + which * (FPC_MAX_SCL_ENTRIES * 8);
                            const int *meta = reinterpret_cast<const int *>(sclbuf + 2 * FPC_MAX_SCL_ENTRIES * 8) + 2 * which;
                            const int sn = valid ? meta[0] : 0, sdt = meta[1];
                            const float x0 = rs[uc * kLdR + 3];
                            float qf = 0.0f;
                            double qd = 0.0;
                            const int if32 = group_scl_nearest<float, G>(reinterpret_cast<const float *>(sclt), sdt == FPC_F32 ? sn : 0, x0, part, qf);
                            const int if64 = group_scl_nearest<double, G>(reinterpret_cast<const double *>(sclt), sdt == FPC_F32 ? 0 : sn, x0, part, qd);
                            if (sn > 0) {
                                coded = true;
                                i0 = sdt == FPC_F32 ? if32 : if64;
                                qv = sdt == FPC_F32 ? qf : (float)qd;
                            }
                        }
                        if (part == 0 && u < MT) {
                            if (coded) rq[u * 20] = qv;
                            m1s[u] = m1; m2s[u] = m2; idx0s[u] = i0; idx1s[u] = -1; idx2s[u] = -1;
                        }
                    }
                    named_bar_sync(1, kTailThreads);

                    FPC_PHASE(kPhScalar);
                    if (P.mode == kModeQuantize) {
                        // ---- VQ for c1..c17 (:228-240): compact the tile rows by branch ----
                        if (twarp == 0) {
                            const bool valid = lane < MT && b0 + lane < P.B;
                            const bool above = valid && m2s[lane < MT ? lane : 0] != 0.0f;
                            const bool below = valid && !above && cbh->bl.stages > 0;
                            const unsigned ba = __ballot_sync(0xffffffffu, above);
                            const unsigned bb = __ballot_sync(0xffffffffu, below);
                            const unsigned lt = (1u << lane) - 1u;
                            if (above) listA[__popc(ba & lt)] = lane;
                            if (below) listB[__popc(bb & lt)] = lane;
                            if (lane == 0) { counts[0] = __popc(ba); counts[1] = __popc(bb); }
                        }
                        named_bar_sync(1, kTailThreads);
                        const int nA = counts[0], nB = counts[1];
                        VqTcMem vmem;
                        vmem.a = smem + S::offVqA;
                        vmem.small = reinterpret_cast<char *>(smem) + S::offVqSmall;
                        vmem.part = smem + S::offPart;
                        vmem.bring = smem + S::offBring;
                        vmem.scratch = reinterpret_cast<char *>(smem) + S::offVqSmall;
                        vmem.scratch_bytes = S::kScratchBytes;
                        vmem.tail = ftail;
                        // one call site for both books (above / below threshold): a single copy of the search.  The helper roles
                        // are owed exactly one publication with last = 1 per frame.
                        const bool tcB = nB > 0 && cbh->bl.K >= 64;
    #pragma unroll 1
                        for (int book = 0; book < 2; ++book) {
                            const int nrows = book ? nB : nA;
                            if (nrows > 0)
                                vq_tc_dispatch<MT, kVqNB, kTailThreads>(book ? cbh->bl : cbh->vq, P.cb, book ? listB : listA, nrows, rs, rq, idx1s, idx2s,
                                                                        vmem, vsh, (book == 1 || !tcB) ? 1 : 0, ttid, prof ? pt + kPhVqDbg : nullptr);
                        }
                        if (!(nA > 0 && cbh->vq.K >= 64) && !tcB) vq_tc_publish_idle<kVqNB, kTailThreads>(vsh, ttid);
                    }
                } else {
                    named_bar_sync(1, kTailThreads);
                }

                FPC_PHASE(kPhVq);
                // ---- feedback (:242 / :252), outputs, next input frame ----
    #pragma unroll
                for (int q = 0; q < NE; ++q) {
                    const int e = ttid + kTailThreads * q;
                    const int u = e / 20, j = e - u * 20;
                    if (e < MT * 20) {
                        const bool valid = b0 + u < P.B;
                        const size_t fo = (size_t)(b0 + u) * P.L + fr;
                        float cin;
                        if (j < kFc) {
                            float ro, rqo, ruo;
                            if (P.mode == kModeQuantize) {
                                rqo = rq[u * 20 + j];
                                ro = rsv[q];
                                ruo = 0.0f;
                                cin = __fadd_rn(fov[q], rqo);
                            } else if (P.mode == kModeResidual) {
                                const float m = j == 0 ? m1s[u] : m2s[u];
                                ruo = __fmul_rn(rsv[q], __fsub_rn(1.0f, m));
                                ro = __fmul_rn(rsv[q], m);
                                rqo = 0.0f;
                                cin = __fadd_rn(fov[q], ro);
                            } else {
                                ro = rqo = ruo = 0.0f;
                                cin = __fadd_rn(fov[q], featv[q]);   // decode: featv holds r_qtz[t]
                            }
                            if (valid && P.mode != kModeDecode) {
                                P.r[fo * kFc + j] = ro;
                                P.r_qtz[fo * kFc + j] = rqo;
                                if (P.r_under) P.r_under[fo * kFc + j] = ruo;
                            }
                        } else {
                            cin = featv[q];   // pitch pass-through (:178)
                        }
                        xin[u * kLdX + j] = cin;
                        if (valid) P.c_in[fo * 20 + j] = cin;
                    }
                }
                if (P.mode != kModeDecode && ttid < MT && b0 + ttid < P.B) {
                    const size_t fo = (size_t)(b0 + ttid) * P.L + fr;
                    const float m1 = m1s[ttid], m2 = m2s[ttid];
                    // the reference fills ind*_mask only in the threshold branch (:204,208)
                    if (P.ind1) P.ind1[fo] = P.mask ? 0.0f : m1;
                    if (P.ind2) P.ind2[fo] = P.mask ? 0.0f : m2;
                    if (P.idx) {
                        int4 v;
                        v.x = idx0s[ttid]; v.y = idx1s[ttid]; v.z = idx2s[ttid];
                        v.w = (m1 != 0.0f ? 1 : 0) | (m2 != 0.0f ? 2 : 0);
                        *reinterpret_cast<int4 *>(P.idx + fo * 4) = v;
                    }
                }
                named_bar_sync(1, kTailThreads);
                if (carry && fr + 1 == P.f1)
                    for (int i = ttid; i < MT * kLdX; i += kTailThreads) carry[S::kStateFloats + i] = xin[i];
                if (ttid == 0 && fr + 1 < P.f1) mbar_arrive(x_ready);       // (the next tile's set-up arrives for the last frame)
                FPC_PHASE(kPhOut);
                if (prof) pt[kPhFrames] += 1;
            }
        }
        if (prof)
            for (int i = 0; i < kPhCount; ++i)
                if (pt[i] != 0) atomicAdd(reinterpret_cast<unsigned long long *>(P.prof) + (size_t)blockIdx.x * kPhCount + i, (unsigned long long)pt[i]);
    #undef FPC_PHASE
        named_bar_sync(1, kTailThreads);
        if (warp == kTailWarp0) umma::tmem_dealloc(tb, 512);
        return;
    }

// ---------------- gate GEMM warps ----------------
    asm volatile("setmaxnreg.inc.sync.aligned.u32 176;");     // kRegGemm
    // Row group per WARP, unit pair per lane: the activation loads of the gate GEMM are then full-warp broadcasts
    // (one shared-memory wavefront instead of two) and a weight LDS.128 reads 512 distinct bytes.  The GEMM needs
    // 48 + 14 wavefronts per k step against 84 FP32-pipe cycles; with row groups inside the warp it was 48 + 28.
    const int tg = warp >> 1;
    const int ug = (warp & 1) * 32 + lane;
    const uint32_t tpark = tb + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)((warp >> 2) * T::kSetCols);
    {
        const float *tail = P.wstream + kStreamFloats;
        for (int i = tid; i < kBiasFloats; i += kComputeThreads) bias[i] = tail[i];
    }
    Pipe pp{0, 0u};
    uint32_t phx = 0;
    const uint32_t zmask = (uint32_t)(P.f0 >> 31);      // zero (f0 >= 0), but not to the compiler: see gemm_part
    const bool prof = P.prof != nullptr && tid == 0;
    long long t_busy = 0, t_wait = 0;
    for (int tile = blockIdx.x; tile < P.ntiles; tile += gridDim.x) {
        float *carry = P.state ? reinterpret_cast<float *>(P.state) + (size_t)tile * (S::kStateFloats + MT * kLdX) : nullptr;
        // the tail has set up the first input frame of this tile -- and is therefore done with the previous tile (it read h2)
        long long t0 = prof ? clock64() : 0;
        mbar_wait(x_ready, phx); phx ^= 1u;
        if (prof) { const long long t = clock64(); t_wait += t - t0; t0 = t; }
        if (carry && P.f0 > 0) {       // continue the recurrence where the launch of the previous frame range stopped
            for (int i = tid; i < MT * kLd1; i += kComputeThreads) h1[i] = carry[i];
            for (int i = tid; i < MT * kLd2; i += kComputeThreads) h2buf[0][i] = carry[MT * kLd1 + i];
        } else {
            for (int i = tid; i < MT * kLd1; i += kComputeThreads) h1[i] = 0.0f;       // h1 = h2 = None -> zeros
            for (int i = tid; i < MT * kLd2; i += kComputeThreads) h2buf[0][i] = 0.0f;
        }
        named_bar_sync(2, kComputeThreads);
        int cur = 0;
        // hidden part of GRU 1 for the first frame of the range
#pragma unroll 1
        for (int pass = 0; pass < 3; ++pass)
            gru1_hidden_pass<TU>(h1 + tg * kLd1, bias + pass * 512, tpark + pass * T::kPassCols, ring, full, empty, pp, ug, lane, zmask);
        tmem_wait_st();
        for (int fr = P.f0; fr < P.f1; ++fr) {
            if (fr > P.f0) {
                if (prof) { const long long t = clock64(); t_busy += t - t0; t0 = t; }
                mbar_wait(x_ready, phx); phx ^= 1u;
                if (prof) { const long long t = clock64(); t_wait += t - t0; t0 = t; }
            }
            // ---- GRU 1 (wavernn.py:71): input part + gates, three passes of 128 hidden units ----
#pragma unroll 1
            for (int pass = 0; pass < 3; ++pass)
                gru1_input_pass<TU>(xin + tg * kLdX, bias + pass * 512, tpark + pass * T::kPassCols, h1 + tg * kLd1 + pass * 128 + 2 * ug,
                                    ring, full, empty, pp, ug, lane, zmask);
            named_bar_sync(2, kComputeThreads);
            // ---- GRU 2 (wavernn.py:76): input is the new h1 ----
            gru2_pass<TU>(h1 + tg * kLd1, h2buf[cur] + tg * kLd2, h2buf[cur] + tg * kLd2 + 2 * ug, h2buf[cur ^ 1] + tg * kLd2 + 2 * ug,
                          bias + 3 * 512, ring, full, empty, pp, ug, lane, zmask);
            named_bar_sync(2, kComputeThreads);
            if (tid == 0) mbar_arrive(h2_ready);
            cur ^= 1;
            // ---- hidden part of GRU 1 for the next frame, while the tail quantises this one ----
            if (fr + 1 < P.f1) {
#pragma unroll 1
                for (int pass = 0; pass < 3; ++pass)
                    gru1_hidden_pass<TU>(h1 + tg * kLd1, bias + pass * 512, tpark + pass * T::kPassCols, ring, full, empty, pp, ug, lane, zmask);
                tmem_wait_st();
            }
        }
        if (prof) { const long long t = clock64(); t_busy += t - t0; t0 = t; }
        if (carry) {
            for (int i = tid; i < MT * kLd1; i += kComputeThreads) carry[i] = h1[i];
            for (int i = tid; i < MT * kLd2; i += kComputeThreads) carry[MT * kLd1 + i] = h2buf[cur][i];
        }
        named_bar_sync(2, kComputeThreads);
        // an odd number of frames leaves the current h2 in buffer 1: the next tile starts from buffer 0 again
    }
    if (prof) {
        unsigned long long *pb = reinterpret_cast<unsigned long long *>(P.prof) + (size_t)blockIdx.x * kPhCount;
        atomicAdd(pb + kPhGru, (unsigned long long)t_busy);
        atomicAdd(pb + kPhWaitX, (unsigned long long)t_wait);
    }
}

template <int TU>
static int launch_encode(const EncodeParams &P, int grid, cudaStream_t st)
{
    using S = Smem<TU>;
    static bool configured[kMaxDevices] = {};
    { const int rc = ensure_dynamic_smem(encode_fp32_kernel<TU>, S::total, configured); if (rc != FPC_OK) return rc; }
    encode_fp32_kernel<TU><<<grid, kThreads, S::total, st>>>(P);
    FPC_LAUNCH_CHECK();
    return FPC_OK;
}

int num_sms()
{
    static int cached[kMaxDevices] = {};
    const int slot = device_slot();
    if (cached[slot] == 0 || slot == kMaxDevices - 1) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
        cached[slot] = n;
    }
    return cached[slot];
}

// ------------------------------------------------------------------------------------------
// Launch plan.  Utterances are independent (wavernn.py:217,228 only loop over k), tiles of one height cost the
// same, and a CTA owns an SM, so a launch finishes in  waves x (height + fixed)  tile-frames.  One tile height
// for the whole batch wastes up to a wave (12 500 utterances on 148 SMs: 391 tiles of 32 = 2.64 waves run as 3).
// The batch is therefore cut into at most kMaxSegments consecutive utterance ranges, each launched with its
// own tile height: whole waves of the tallest tile first, then the heights that fit what is left
// (12 500 -> 2 waves of 32 + 1 wave of 24, cost 106 instead of 114).  Every range is an ordinary launch of the
// same kernel on offset pointers, so results do not depend on the plan.
// ------------------------------------------------------------------------------------------
static double plan_rest(int B, int sms, const int *heights, int nh, double fixed, int depth, EncodeSegment *out, int *nout)
{
    if (B <= 0) { *nout = 0; return 0.0; }
    double best = 1e300;
    EncodeSegment best_seg[kMaxSegments];
    int best_n = 0;
    for (int c = 0; c < nh; ++c) {
        const int h = heights[c];
        const long long per_wave = (long long)sms * h;
        const int wmax = (int)((B + per_wave - 1) / per_wave);
        // this height closes the plan ...
        {
            const double cost = wmax * (h + fixed);
            if (cost < best - 1e-9) { best = cost; best_n = 1; best_seg[0] = EncodeSegment{h, 0, B}; }
        }
        // ... or takes w whole waves and hands the rest to another height
        if (depth + 1 < kMaxSegments) {
            const int wlo = depth == 0 ? (wmax - 3 > 1 ? wmax - 3 : 1) : 1;     // the bulk goes to the first segment
            for (int w = wlo; w < wmax; ++w) {
                const int take = (int)(w * per_wave);
                EncodeSegment sub[kMaxSegments];
                int nsub = 0;
                const double cost = w * (h + fixed) + plan_rest(B - take, sms, heights, nh, fixed, depth + 1, sub, &nsub) + 0.25;
                if (cost < best - 1e-9) {           // + 0.25 per extra launch: prefer fewer segments on near-ties
                    best = cost; best_n = 1 + nsub;
                    best_seg[0] = EncodeSegment{h, 0, take};
                    for (int i = 0; i < nsub; ++i) { best_seg[1 + i] = sub[i]; best_seg[1 + i].first += take; }
                }
            }
        }
    }
    for (int i = 0; i < best_n; ++i) out[i] = best_seg[i];
    *nout = best_n;
    return best;
}

int plan_segments(int B, int sms, const int *heights, int nh, double fixed, EncodeSegment *out)
{
    int n = 0;
    plan_rest(B, sms, heights, nh, fixed, 0, out, &n);
    return n;
}

static const int kHeightsF32[4] = {32, 28, 24, 16};
static const int kHeightsBf16Plan[2] = {64, 32};

int encode_plan(int B, int precision, int sms, int *segments)
{
    EncodeSegment seg[kMaxSegments];
    const int n = precision == FPC_PREC_BF16 ? plan_segments(B, sms, kHeightsBf16Plan, 2, 6.0, seg)
                                             : plan_segments(B, sms, kHeightsF32, 4, 6.0, seg);
    for (int i = 0; i < n; ++i) { segments[3 * i] = seg[i].height; segments[3 * i + 1] = seg[i].first; segments[3 * i + 2] = seg[i].count; }
    return n;
}

static size_t state_bytes_per_tile(int mt) { return (size_t)(mt * (kLd1 + kLd2) + mt * kLdX) * sizeof(float); }

size_t encode_fp32_state_bytes(int B)
{
    const int sms = num_sms();
    if (sms <= 0 || B <= 0) return 0;
    EncodeSegment seg[kMaxSegments];
    const int n = plan_segments(B, sms, kHeightsF32, 4, 6.0, seg);
    size_t total = 0;
    for (int i = 0; i < n; ++i) total += (size_t)((seg[i].count + seg[i].height - 1) / seg[i].height) * state_bytes_per_tile(seg[i].height);
    return total;
}

// the launch parameters of one utterance range of the batch
EncodeParams segment_params(const EncodeParams &P, int first, int count)
{
    EncodeParams Q = P;
    const size_t fo = (size_t)first * P.L;
    if (Q.feat) Q.feat += fo * 20;
    if (Q.mask) Q.mask += fo * 2;
    if (Q.rq_in) Q.rq_in += fo * kFc;
    if (Q.pitch_in) Q.pitch_in += fo * 2;
    if (Q.c_in) Q.c_in += fo * 20;
    if (Q.r) Q.r += fo * kFc;
    if (Q.r_qtz) Q.r_qtz += fo * kFc;
    if (Q.r_under) Q.r_under += fo * kFc;
    if (Q.ind1) Q.ind1 += fo;
    if (Q.ind2) Q.ind2 += fo;
    if (Q.idx) Q.idx += fo * 4;
    Q.B = count;
    return Q;
}

int run_encode_fp32(EncodeParams P, cudaStream_t st, int force_tu)
{
    if (P.f0 < 0 || P.f1 > P.L || P.f0 >= P.f1) return FPC_ERR_ARG;
    const int sms = num_sms();
    if (sms <= 0) return cuda_fail(cudaErrorNoDevice);
    EncodeSegment seg[kMaxSegments];
    int n = 1;
    if (force_tu == 0) {          // debugging / tests: one tile height for the whole batch (16, 24, 28 or 32)
        const char *e = getenv("FPC_FP32_TILE");
        if (e != nullptr && atoi(e) > 0) force_tu = atoi(e) / 4;
    }
    if (force_tu > 0) seg[0] = EncodeSegment{4 * force_tu, 0, P.B};
    else n = plan_segments(P.B, sms, kHeightsF32, 4, 6.0, seg);
    char *state = reinterpret_cast<char *>(P.state);
    for (int i = 0; i < n; ++i) {
        EncodeParams Q = segment_params(P, seg[i].first, seg[i].count);
        const int mt = seg[i].height;
        Q.ntiles = (Q.B + mt - 1) / mt;
        Q.state = state;
        if (state) state += (size_t)Q.ntiles * state_bytes_per_tile(mt);
        const int grid = Q.ntiles < sms ? Q.ntiles : sms;
        int rc;
        switch (mt / 4) {
            case 4: rc = launch_encode<4>(Q, grid, st); break;
            case 6: rc = launch_encode<6>(Q, grid, st); break;
            case 7: rc = launch_encode<7>(Q, grid, st); break;
            case 8: rc = launch_encode<8>(Q, grid, st); break;
            default: rc = FPC_ERR_ARG;
        }
        if (rc != FPC_OK) return rc;
    }
    return FPC_OK;
}

}  // namespace fpc
