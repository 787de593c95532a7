// fpc_encode_fp32.cu -- the fused, persistent closed-loop frame-step kernel (fp32 FFMA path).
//
// Replaces the whole body of Wavernn.encoder (/root/reference/src/models/wavernn.py:165-256):
// per frame  forward (:194 -> :63-102), residual (:196), thresholds (:202-208), the injected
// scl_quantize / vq_quantize calls (:217-240 -> quantization/vq_func.py), the feedback (:242 or
// :252), for a tile of MT = 4*TU utterances per CTA with the frame loop INSIDE the kernel.
// Nothing leaves the SM between frames except the per-frame outputs; the GRU states, the
// decoded frame that is fed back and all quantiser scratch live in shared memory.
//
// Work split inside a CTA (384 threads = 3 warpgroups; setmaxnreg moves the producer group's
// registers to the two compute groups, 232 / 40 per thread):
//   warps 0..7  compute.  Thread (tg = warp>>1, ug = (warp&1)*32 + lane) owns hidden units
//               {2ug, 2ug+1} (+128*pass) for utterances {tg, tg+4, ..., tg+4(TU-1)} of the tile,
//               i.e. a TU x 2 x {r,z,n_i,n_h} register tile; every dot product is one
//               ascending-k FFMA chain (the canonical order the oracle follows).
//   warp 8      producer (warps 9..11 only exist to complete the warpgroup).  One lane streams the 2.68 MB packed
//               weight image (217 groups of 12 KB per frame) through a 4-stage shared-memory ring with 1-D bulk
//               async copies (TMA engine, UBLKCP), mbarrier full/empty handshakes; the image stays L2-resident (it is
//               re-read by every CTA every frame) and the ring runs ahead across frame boundaries.
// After the GRUs: FC + 2*tanh, residual, thresholds, scalar quantiser (one warp per utterance), then the screened
// m-best VQ search (fpc_vq_screen.cuh) with the tile's residual vectors broadcast from shared memory.
// A launch covers the frame range [f0, f1) of every utterance; with EncodeParams::state the recurrent state of each
// tile is carried from one launch to the next (fpc_encode_host cuts a batch along time that way).
#include "fpc_common.cuh"
#include "fpc_math.cuh"
#include "fpc_vq.cuh"
#include "fpc_vq_search.cuh"
#include "fpc_vq_screen.cuh"
#include "fpc_encode.cuh"

namespace fpc {

constexpr int kStages = 4;             // weight ring depth
constexpr int kThreads = kComputeThreads + 128;   // 2 compute warpgroups + 1 producer warpgroup
constexpr int kLd1 = kH1 + 4;          // 388: padded row strides (floats) -> conflict-free float4 rows
constexpr int kLd2 = kH2 + 4;          // 132
constexpr int kLdX = 24;               // input frame row (20 used)
constexpr int kLdFc = kH2 + 4;         // 132: float4-aligned rows, conflict-free for 8 consecutive rows

template <int TU> struct Smem {
    static constexpr int MT = 4 * TU;
    static constexpr int kStateSet = MT * (kLd1 + kLd2);                 // floats: [h1 | h2] of one set
    static constexpr int offRing = 0;
    static constexpr int offSetA = offRing + kStages * kGroupBytes;
    static constexpr int offSetB = offSetA + kStateSet * 4;
    static constexpr int offXin = offSetB + kStateSet * 4;
    static constexpr int offBias = offXin + MT * kLdX * 4;
    static constexpr int offFc = offBias + kBiasFloats * 4;              // 18 x 129 weights + 18 bias
    static constexpr int offRs = offFc + ((kFc * kLdFc + kFc + 3) / 4) * 16;
    static constexpr int offRq = offRs + MT * kLdR * 4;                  // quantised residual rows (stride 20)
    static constexpr int offMisc = offRq + MT * 20 * 4;
    // misc: m1[MT] m2[MT] (float), idx0/idx1/idx2[MT] (int), listA[MT] listB[MT] (int), counts[4]
    static constexpr int offScl = ((offMisc + (7 * MT + 4) * 4 + 15) / 16) * 16;   // both scalar tables, file dtype, 2 x 2 KB
    static constexpr int offBars = offScl + 2 * FPC_MAX_SCL_ENTRIES * 8 + 16;   // + {n, dtype} of the two scalar tables
    static constexpr int total = ((offBars + 2 * kStages * 8 + 127) / 128) * 128;
    // VQ scratch aliases the DEAD state set (the one holding the previous frame's h1/h2)
    static constexpr int kScratchBytes = kStateSet * 4;
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
                                          uint64_t *full, uint64_t *empty, Pipe &pp, int ug, int lane)
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
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[pp.s]);
#endif
        pp.advance();
    }
}

// one pass = 128 hidden units of a GRU: bias init, input part, hidden part, gate epilogue
template <int TU>
__device__ __forceinline__ void gru_pass(int ngx, int ngh, const float *__restrict__ xrow, int ldx,
                                         const float *__restrict__ hrow, int ldh, const float *__restrict__ hold,
                                         float *__restrict__ hnew, int ldo, const float *__restrict__ bias,
                                         const float4 *__restrict__ ring, uint64_t *full, uint64_t *empty, Pipe &pp,
                                         int ug, int lane)
{
    float2 ar[TU], az[TU], ani[TU], anh[TU];
    const float2 br = *reinterpret_cast<const float2 *>(bias + 0 * 128 + 2 * ug);
    const float2 bz = *reinterpret_cast<const float2 *>(bias + 1 * 128 + 2 * ug);
    const float2 bi = *reinterpret_cast<const float2 *>(bias + 2 * 128 + 2 * ug);
    const float2 bh = *reinterpret_cast<const float2 *>(bias + 3 * 128 + 2 * ug);
#pragma unroll
    for (int i = 0; i < TU; ++i) { ar[i] = br; az[i] = bz; ani[i] = bi; anh[i] = bh; }
    gemm_part<TU>(ar, az, ani, ngx, xrow, ldx, ring, full, empty, pp, ug, lane);
    gemm_part<TU>(ar, az, anh, ngh, hrow, ldh, ring, full, empty, pp, ug, lane);
#pragma unroll
    for (int i = 0; i < TU; ++i) {
        const float2 ho = *reinterpret_cast<const float2 *>(hold + (size_t)(4 * i) * ldo);
        float2 hn;
        hn.x = gru_update(ar[i].x, az[i].x, ani[i].x, anh[i].x, ho.x);
        hn.y = gru_update(ar[i].y, az[i].y, ani[i].y, anh[i].y, ho.y);
        *reinterpret_cast<float2 *>(hnew + (size_t)(4 * i) * ldo) = hn;
    }
}

// ------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------
template <int TU>
__global__ void __launch_bounds__(kThreads, 1) encode_fp32_kernel(EncodeParams P)
{
    using S = Smem<TU>;
    constexpr int MT = S::MT;
    constexpr int NE = (MT * 20 + kComputeThreads - 1) / kComputeThreads;   // frame elements per thread
    extern __shared__ __align__(128) unsigned char smem[];
    float4 *ring = reinterpret_cast<float4 *>(smem + S::offRing);
    float *setA = reinterpret_cast<float *>(smem + S::offSetA);
    float *setB = reinterpret_cast<float *>(smem + S::offSetB);
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
    int *counts = listB + MT;
    unsigned char *sclbuf = smem + S::offScl;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + S::offBars);
    uint64_t *empty = full + kStages;

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int my_tiles = P.ntiles > (int)blockIdx.x ? (P.ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kComputeThreads / 32);
        }
        mbar_fence_init();
    }
    __syncthreads();
    // ---------------- producer warpgroup (only one lane works; it hands its registers over) ----------------
    if (warp >= kComputeThreads / 32) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        if (warp == kComputeThreads / 32 && lane == 0) {
            const long long total = (long long)my_tiles * (P.f1 - P.f0) * kGroupsPerFrame;
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
                bulk_g2s(smem + S::offRing + s * kGroupBytes, src + (size_t)gf * kGroupBytes, kGroupBytes, &full[s]);
                if (++gf == kGroupsPerFrame) gf = 0;
                if (++s == kStages) { s = 0; ++wraps; }
            }
        }
        return;
    }

    // ---------------- compute warps ----------------
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    // Row group per WARP, unit pair per lane: the activation loads of the gate GEMM are then full-warp broadcasts
    // (one shared-memory wavefront instead of two) and a weight LDS.128 reads 512 distinct bytes.  The GEMM needs
    // 48 + 14 wavefronts per k step against 84 FP32-pipe cycles; with row groups inside the warp it was 48 + 28.
    const int tg = warp >> 1;
    const int ug = (warp & 1) * 32 + lane;
    const PackedCodebooks *cbh = reinterpret_cast<const PackedCodebooks *>(P.cb);

    // constants resident for the whole kernel
    {
        const float *tail = P.wstream + kStreamFloats;
        for (int i = tid; i < kBiasFloats; i += kComputeThreads) bias[i] = tail[i];
        for (int i = tid; i < kFcFloats; i += kComputeThreads) wfc[(i >> 7) * kLdFc + (i & 127)] = tail[kBiasFloats + i];
        if (tid < kFc) bfc[tid] = tail[kBiasFloats + kFcFloats + tid];
        // the scalar tables (<= 256 levels each) next to the state: the VQ streams 150 KB of codebook through L1 every
        // frame, so from global memory every scalar search paid L2 latencies
        if (cbh != nullptr) {
            const long long *src0 = reinterpret_cast<const long long *>(P.cb + cbh->scl.off);
            const long long *src1 = reinterpret_cast<const long long *>(P.cb + cbh->blscl.off);
            const int n0 = cbh->scl.n * (cbh->scl.dtype == FPC_F32 ? 4 : 8), n1 = cbh->blscl.n * (cbh->blscl.dtype == FPC_F32 ? 4 : 8);
            for (int i = tid; i < (n0 + 7) / 8; i += kComputeThreads) reinterpret_cast<long long *>(sclbuf)[i] = src0[i];
            for (int i = tid; i < (n1 + 7) / 8; i += kComputeThreads)
                reinterpret_cast<long long *>(sclbuf + FPC_MAX_SCL_ENTRIES * 8)[i] = src1[i];
            if (tid == 0) {
                int *meta = reinterpret_cast<int *>(sclbuf + 2 * FPC_MAX_SCL_ENTRIES * 8);
                meta[0] = cbh->scl.n; meta[1] = cbh->scl.dtype; meta[2] = cbh->blscl.n; meta[3] = cbh->blscl.dtype;
            }
        }
    }
    Pipe pp{0, 0u};
    const bool prof = P.prof != nullptr && tid == 0;
    long long pt[kPhCount] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, pt0 = 0;
#define FPC_PHASE(ph) do { if (prof) { const long long t_ = clock64(); pt[ph] += t_ - pt0; pt0 = t_; } } while (0)

    for (int tile = blockIdx.x; tile < P.ntiles; tile += gridDim.x) {
        const int b0 = tile * MT;
        float *cur = setA, *nxt = setB;
        float *carry = P.state ? reinterpret_cast<float *>(P.state) + (size_t)tile * (S::kStateSet + MT * kLdX) : nullptr;
        if (carry && P.f0 > 0) {       // continue the recurrence where the launch of the previous frame range stopped
            for (int i = tid; i < S::kStateSet; i += kComputeThreads) cur[i] = carry[i];
            for (int i = tid; i < MT * kLdX; i += kComputeThreads) xin[i] = carry[S::kStateSet + i];
        } else {
            for (int i = tid; i < S::kStateSet; i += kComputeThreads) cur[i] = 0.0f;   // h1 = h2 = None -> zeros
            for (int i = tid; i < MT * kLdX; i += kComputeThreads) xin[i] = 0.0f;      // frame 0 input is all zero
        }
        for (int i = tid; i < MT * kLdR; i += kComputeThreads) rs[i] = 0.0f;
        named_bar_sync(1, kComputeThreads);

        for (int fr = P.f0; fr < P.f1; ++fr) {
            // element ownership for this frame: e = tid + 256 q -> (row u, feature j)
            float featv[NE], fov[NE], rsv[NE];
#pragma unroll
            for (int q = 0; q < NE; ++q) {
                const int e = tid + kComputeThreads * q;
                const int u = e / 20, j = e - u * 20;
                featv[q] = 0.0f; fov[q] = 0.0f; rsv[q] = 0.0f;
                if (e < MT * 20 && b0 + u < P.B) {
                    const size_t fo = (size_t)(b0 + u) * P.L + fr;
                    if (P.mode != kModeDecode) featv[q] = __ldg(P.feat + fo * 20 + j);
                    else featv[q] = j < kFc ? __ldg(P.rq_in + fo * kFc + j) : __ldg(P.pitch_in + fo * 2 + (j - kFc));
                }
            }
            float *h1c = cur, *h2c = cur + MT * kLd1;
            float *h1n = nxt, *h2n = nxt + MT * kLd1;

            if (prof) pt0 = clock64();
            // ---- GRU 1: three passes of 128 hidden units (wavernn.py:71) ----
#pragma unroll 1
            for (int pass = 0; pass < 3; ++pass) {
                gru_pass<TU>(kG1x, kG1h, xin + tg * kLdX, kLdX, h1c + tg * kLd1, kLd1,
                             h1c + tg * kLd1 + pass * 128 + 2 * ug, h1n + tg * kLd1 + pass * 128 + 2 * ug, kLd1,
                             bias + pass * 512, ring, full, empty, pp, ug, lane);
            }
            named_bar_sync(1, kComputeThreads);
            // ---- GRU 2 (wavernn.py:76): input is the new h1 ----
            gru_pass<TU>(kG2x, kG2h, h1n + tg * kLd1, kLd1, h2c + tg * kLd2, kLd2, h2c + tg * kLd2 + 2 * ug,
                         h2n + tg * kLd2 + 2 * ug, kLd2, bias + 3 * 512, ring, full, empty, pp, ug, lane);
            named_bar_sync(1, kComputeThreads);

            FPC_PHASE(kPhGru);
            // ---- relu, dual_fc, 2*tanh (wavernn.py:87-92); residual (:196) ----
            {
                // the thread's NE outputs advance together (NE independent ascending-k chains, operands as float4)
                float acc[NE];
                const float4 *wr4[NE], *hv4[NE];
                bool on[NE];
#pragma unroll
                for (int q = 0; q < NE; ++q) {
                    const int e = tid + kComputeThreads * q;
                    const int u = e / 20, j = e - u * 20;
                    on[q] = e < MT * 20 && j < kFc;
                    acc[q] = on[q] ? bfc[j] : 0.0f;
                    wr4[q] = reinterpret_cast<const float4 *>(wfc + (on[q] ? j : 0) * kLdFc);
                    hv4[q] = reinterpret_cast<const float4 *>(h2n + (on[q] ? u : 0) * kLd2);
                }
#pragma unroll 4
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
                        const int e = tid + kComputeThreads * q;
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
                for (int i = tid; i < MT * 20; i += kComputeThreads) rq[i] = 0.0f;
                named_bar_sync(1, kComputeThreads);

                FPC_PHASE(kPhFc);
                // ---- indicators (:201-212) and the scalar quantiser for c0 (:217-225) ----
                for (int u = warp; u < MT; u += 8) {
                    const bool valid = b0 + u < P.B;
                    float m1 = 0.0f, m2 = 0.0f;
                    if (lane == 0 && valid) {
                        if (P.mask == nullptr) {
                            float s = 0.0f;
#pragma unroll
                            for (int j = 1; j < kFc; ++j) s = __fadd_rn(s, fabsf(rs[u * kLdR + 3 + j]));
                            m1 = fabsf(rs[u * kLdR + 3]) > P.l1 ? 1.0f : 0.0f;
                            m2 = s > P.l2 ? 1.0f : 0.0f;
                        } else {
                            const size_t fo = (size_t)(b0 + u) * P.L + fr;
                            m1 = __ldg(P.mask + fo * 2);
                            m2 = __ldg(P.mask + fo * 2 + 1);
                        }
                    }
                    m1 = __shfl_sync(0xffffffffu, m1, 0);
                    m2 = __shfl_sync(0xffffffffu, m2, 0);
                    int i0 = -1;
                    if (P.mode == kModeQuantize && valid) {
                        const int which = (m1 != 0.0f) ? 0 : 1;       // above / below threshold table (:217-225)
                        const unsigned char *sclt = sclbuf + which * (FPC_MAX_SCL_ENTRIES * 8);
                        const int *meta = reinterpret_cast<const int *>(sclbuf + 2 * FPC_MAX_SCL_ENTRIES * 8) + 2 * which;
                        struct { int n, dtype; } sb = {meta[0], meta[1]};
                        if (sb.n > 0) {
                            const float x0 = rs[u * kLdR + 3];
                            float qv;
                            if (sb.dtype == FPC_F32) {
                                float q;
                                i0 = warp_scl_nearest<float>(reinterpret_cast<const float *>(sclt), sb.n, x0, lane, q);
                                qv = q;
                            } else {
                                double q;
                                i0 = warp_scl_nearest<double>(reinterpret_cast<const double *>(sclt), sb.n, x0, lane, q);
                                qv = (float)q;
                            }
                            if (lane == 0) rq[u * 20] = qv;
                        }
                    }
                    if (lane == 0) { m1s[u] = m1; m2s[u] = m2; idx0s[u] = i0; idx1s[u] = -1; idx2s[u] = -1; }
                }
                named_bar_sync(1, kComputeThreads);

                FPC_PHASE(kPhScalar);
                if (P.mode == kModeQuantize) {
                    // ---- VQ for c1..c17 (:228-240): compact the tile rows by branch ----
                    if (warp == 0) {
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
                    named_bar_sync(1, kComputeThreads);
                    const int nA = counts[0], nB = counts[1];
                    char *scratch = reinterpret_cast<char *>(cur);   // dead state set (see Smem)
                    // one call site for both books (above / below threshold): a single inlined copy of the search
#pragma unroll 1
                    for (int book = 0; book < 2; ++book) {
                        const int nrows = book ? nB : nA;
                        if (nrows > 0)
                            vq_dispatch_screened(book ? cbh->bl : cbh->vq, P.cb, book ? listB : listA, nrows, MT, rs, rq, idx1s, idx2s,
                                                 scratch, S::kScratchBytes, tid, prof ? pt + kPhVqDbg : nullptr);
                    }
                }
            } else {
                named_bar_sync(1, kComputeThreads);
            }

            FPC_PHASE(kPhVq);
            // ---- feedback (:242 / :252), outputs, next input frame ----
#pragma unroll
            for (int q = 0; q < NE; ++q) {
                const int e = tid + kComputeThreads * q;
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
            if (P.mode != kModeDecode && tid < MT && b0 + tid < P.B) {
                const size_t fo = (size_t)(b0 + tid) * P.L + fr;
                const float m1 = m1s[tid], m2 = m2s[tid];
                // the reference fills ind*_mask only in the threshold branch (:204,208)
                if (P.ind1) P.ind1[fo] = P.mask ? 0.0f : m1;
                if (P.ind2) P.ind2[fo] = P.mask ? 0.0f : m2;
                if (P.idx) {
                    int4 v;
                    v.x = idx0s[tid]; v.y = idx1s[tid]; v.z = idx2s[tid];
                    v.w = (m1 != 0.0f ? 1 : 0) | (m2 != 0.0f ? 2 : 0);
                    *reinterpret_cast<int4 *>(P.idx + fo * 4) = v;
                }
            }
            named_bar_sync(1, kComputeThreads);
            FPC_PHASE(kPhOut);
            if (prof) pt[kPhFrames] += 1;
            float *t = cur; cur = nxt; nxt = t;
        }
        if (carry) {
            for (int i = tid; i < S::kStateSet; i += kComputeThreads) carry[i] = cur[i];
            for (int i = tid; i < MT * kLdX; i += kComputeThreads) carry[S::kStateSet + i] = xin[i];
            named_bar_sync(1, kComputeThreads);
        }
    }
    if (prof)
        for (int i = 0; i < kPhCount; ++i) atomicAdd(reinterpret_cast<unsigned long long *>(P.prof) + (size_t)blockIdx.x * kPhCount + i, (unsigned long long)pt[i]);
#undef FPC_PHASE
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
