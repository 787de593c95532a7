// fpc_encode_bf16.cu -- fused persistent closed-loop frame-step kernel, bf16 predictor on the
// 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM, weights fed by the TMA engine).
//
// Same contract as fpc_encode_fp32.cu (Wavernn.encoder, /root/reference/src/models/wavernn.py:165-256)
// with the GRU gate contractions (:71,76) evaluated as bf16 x bf16 -> fp32 tensor-core products; the
// thresholds, both quantisers (exact fp32/fp64 arithmetic of quantization/vq_func.py) and the feedback
// are the same CUDA-core code as in the fp32 kernel.
//
// Orientation.  Per CTA a tile of NU (32 or 64) utterances.  The products are computed transposed,
//     D^T [128 gate rows x NU utterances] += W_tile [128 x 16] * Act [NU x 16]^T
// so that TMEM lane = hidden unit, TMEM column = utterance: the weight tile is the A operand
// (M = 128), the activations are the B operand (N = NU), and a pass of 128 hidden units needs
// 4 * NU accumulator columns (r, z, n_i, n_h).  The recurrent state is kept as bf16 operand tiles in
// shared memory ([x(32) | h1(384)] ping-pong, h2 in place): what the tensor core reads IS the state.
//
// Roles (384 threads): warps 0-7 compute -- gate epilogue straight from TMEM (tcgen05.ld; lane quarter
// = warp % 4, column half = warp / 4), then FC, residual, thresholds, scalar quantiser, m-best VQ and
// the feedback; warp 8 lane 0 streams the packed bf16 weight image (1.35 MB per frame, L2 resident)
// through a 3-stage x 12 KB ring with bulk async copies; warp 9 lane 0 issues the MMAs.  Hand-offs are
// mbarriers: ring full/empty (TMA tx bytes / tcgen05.commit), accumulator full/empty, activations
// ready (generic-proxy writes fenced to the async proxy).
#include "fpc_common.cuh"
#include "fpc_math.cuh"
#include "fpc_vq.cuh"
#include "fpc_vq_search.cuh"
#include "fpc_vq_screen.cuh"
#define FPC_VQ_TC_OUTLINE __forceinline__      // the search inlined: it then gets the compute warps' 216 registers (out of line: the launch allocation, 168)
#include "fpc_vq_tc.cuh"
#include "fpc_encode.cuh"
#include "fpc_umma.cuh"

namespace fpc {

constexpr int kBStages = 3;                                   // weight ring depth (a fourth stage measured no faster; the VQ screen and the L1 use the 12 KB)
constexpr int kBTileBytes = 128 * 16 * 2;                      // one A tile: 128 gate rows x K = 16, bf16
constexpr int kBStageBytes = 3 * kBTileBytes;                  // r, z and the third gate (n_i or n_h)
constexpr int kBG1Steps = 2 + kH1 / 16;                        // 26: x padded to 32, then h1
constexpr int kBG2Steps = kH1 / 16 + kH2 / 16;                 // 32: h1' then h2
constexpr int kBStepsPerFrame = 3 * kBG1Steps + kBG2Steps;     // 110
constexpr int kBStreamElems = kBStepsPerFrame * 3 * 2048;
constexpr size_t kBStreamBytes = (size_t)kBStreamElems * 2;    // 1 351 680
constexpr int kBTailFloats = ((kBiasFloats + kFcFloats + kFc + 3) / 4) * 4;
constexpr int kXK = 32 + kH1;                                  // K extent of the [x | h1] operand tile
constexpr int kBThreads = kComputeThreads + 128;
constexpr int kLdFcB = kH2 + 1;

size_t packed_bf16_bytes() { return kBStreamBytes + (size_t)kBTailFloats * 4; }

// ------------------------------------------------------------------------------------------
// weight image: [110 steps][3 tiles][2 k-chunks][128 rows][8 k] bf16, then the fp32 tail
// (biases [4 passes][br+bhr, bz+bhz, b_in, b_hn][128], FC weights, FC bias)
// ------------------------------------------------------------------------------------------
__global__ void pack_weights_bf16_kernel(fpc_weights w, unsigned char *__restrict__ out)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < kBStreamElems) {
        const int step = t / 6144, r0 = t - step * 6144;
        const int tile = r0 / 2048, r1 = r0 - tile * 2048;
        const int chunk = r1 / 1024, r2 = r1 - chunk * 1024;
        const int row = r2 >> 3, kk = chunk * 8 + (r2 & 7);
        float v = 0.0f;
        if (step < 3 * kBG1Steps) {
            const int pass = step / kBG1Steps, i = step - pass * kBG1Steps;
            const int grow = tile * kH1 + pass * 128 + row;          // gate row (r, z, n) of rnn1
            if (i < 2) {
                const int k = 16 * i + kk;
                if (k < kIn) v = w.w_ih1[(size_t)grow * kIn + k];
            } else {
                v = w.w_hh1[(size_t)grow * kH1 + 16 * (i - 2) + kk];
            }
        } else {
            const int i = step - 3 * kBG1Steps;
            const int grow = tile * kH2 + row;
            if (i < kH1 / 16) v = w.w_ih2[(size_t)grow * kH1 + 16 * i + kk];
            else v = w.w_hh2[(size_t)grow * kH2 + 16 * (i - kH1 / 16) + kk];
        }
        reinterpret_cast<__nv_bfloat16 *>(out)[t] = __float2bfloat16_rn(v);
        return;
    }
    t -= kBStreamElems;
    float *tail = reinterpret_cast<float *>(out + kBStreamBytes);
    if (t < kBiasFloats) {
        const int pass = t / 512, kind = (t >> 7) & 3, u = t & 127;
        const float *bi = pass < 3 ? w.b_ih1 : w.b_ih2;
        const float *bh = pass < 3 ? w.b_hh1 : w.b_hh2;
        const int H = pass < 3 ? kH1 : kH2;
        const int j = pass < 3 ? pass * 128 + u : u;
        float v;
        if (kind == 0) v = __fadd_rn(bi[j], bh[j]);
        else if (kind == 1) v = __fadd_rn(bi[H + j], bh[H + j]);
        else if (kind == 2) v = bi[2 * H + j];
        else v = bh[2 * H + j];
        tail[t] = v;
        return;
    }
    t -= kBiasFloats;
    if (t < kFcFloats) { tail[kBiasFloats + t] = w.w_fc[t]; return; }
    t -= kFcFloats;
    if (t < kFc) tail[kBiasFloats + kFcFloats + t] = w.b_fc[t];
}

int pack_weights_bf16(const fpc_weights *w, void *d_packed, cudaStream_t st)
{
    const int n = kBStreamElems + kBiasFloats + kFcFloats + kFc;
    pack_weights_bf16_kernel<<<(n + 255) / 256, 256, 0, st>>>(*w, (unsigned char *)d_packed);
    FPC_LAUNCH_CHECK();
    return FPC_OK;
}

// ------------------------------------------------------------------------------------------
template <int NU> struct SmemB {
    static constexpr int kX1Bytes = NU * kXK * 2;
    static constexpr int kH2Bytes = NU * kH2 * 2;
    static constexpr int offRing = 0;
    static constexpr int offX1a = offRing + kBStages * kBStageBytes;
    static constexpr int offX1b = offX1a + kX1Bytes;
    static constexpr int offH2 = offX1b + kX1Bytes;
    static constexpr int offBias = offH2 + kH2Bytes;
    static constexpr int offFc = offBias + kBiasFloats * 4;
    static constexpr int offRs = offFc + ((kFc * kLdFcB + kFc + 3) / 4) * 16;
    static constexpr int offRq = offRs + NU * kLdR * 4;
    static constexpr int offMisc = offRq + NU * 20 * 4;
    // misc: m1[NU] m2[NU] (float), idx0/idx1/idx2[NU], listA[NU], listB[NU] (int), counts[4], tmem base, pad
    static constexpr int offScl = ((offMisc + (7 * NU + 8) * 4 + 15) / 16) * 16;      // both scalar tables, file dtype, 2 x 2 KB + {n, dtype} x 2
    static constexpr int offBars = offScl + 2 * FPC_MAX_SCL_ENTRIES * 8 + 16;
    static constexpr int kNumBars = 2 * kBStages + 6;
    static constexpr int kBase = ((offBars + kNumBars * 8 + 127) / 128) * 128;
    static constexpr int kScratchBytes = kX1Bytes;   // the dead [x | h1] tile doubles as VQ scratch
    // Tensor-core VQ screen (fpc_vq_tc.cuh).  The A tiles live in the scratch when they fit (64-utterance tiles: 3 x 16 KB
    // of 53 248), else behind the control block; then the small tables, the scan partials and a codebook ring of two
    // 8 KB chunks.
    static constexpr int kMtMax = (5 * NU + 127) / 128;
    static constexpr int kABytes = kMtMax * tc::kTileBytes;
    static constexpr bool kAInScratch = kABytes <= kScratchBytes;
    static constexpr bool kSmallInScratch = kAInScratch && kABytes + vq_tc_small_bytes(NU) <= kScratchBytes;
    static constexpr int offVqSh = kBase;
    static constexpr int offVqSmall = offVqSh + 512;                                                // used when !kSmallInScratch
    static constexpr int offVqA = offVqSmall + (kSmallInScratch ? 0 : ((vq_tc_small_bytes(NU) + 127) / 128) * 128);   // used when !kAInScratch
    static constexpr int offPart = offVqA + (kAInScratch ? 0 : kABytes);
    static constexpr int offBring = offPart + vq_tc_part_bytes(kMtMax);
    // Two chunks, not as many as fit: what the CTA leaves of the 228 KB is L1, and L1 is where the register spills of this
    // kernel live (with 227 KB of shared memory every spill reload was an L2 round trip: the MMA issuer then needed
    // 1.4 k cycles per chunk, tools/phase_profile.py trace).
    static constexpr int kNB = 2;
    static constexpr int total = offBring + kNB * kVtChunkBytes;
    // During the VQ phase of a frame the weight ring of the gate GEMMs is idle -- if its producer is held back until
    // the phase is over (it then refills the ring under the feedback phase, before GRU 1 of the next frame needs it).
    // The screen uses that memory as further ring slots: with two slots only, every codebook copy (~650 cycles from L2)
    // was serialised with the MMAs of the chunk before it (tools/phase_profile.py trace, round 2).
    static constexpr int kNBLent = (kBStages * kBStageBytes) / kVtChunkBytes;      // 4
    static constexpr int kNBTotal = kNB + kNBLent;
    static_assert(total <= 227 * 1024, "shared memory");
};

__device__ __forceinline__ float tanh_fast(float x)
{
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float sigmoid_fast(float x)
{
    return __fmaf_rn(0.5f, tanh_fast(__fmul_rn(0.5f, x)), 0.5f);
}
__device__ __forceinline__ void reg_fence16(uint32_t (&r)[16])
{
    asm volatile(""
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]));
}
__device__ __forceinline__ float bf16_bits_to_float(unsigned short b) { return __uint_as_float((uint32_t)b << 16); }
__device__ __forceinline__ unsigned short float_to_bf16_bits(float f) { return __bfloat16_as_ushort(__float2bfloat16_rn(f)); }

// gate epilogue of one 128-unit pass: TMEM (lane = unit, column = utterance) -> new bf16 state
template <int NU>
__device__ __forceinline__ void gate_epilogue(uint32_t tb, int q, int hsel, int lane, const float *__restrict__ bias,
                                              const unsigned char *__restrict__ hold_base,
                                              unsigned char *__restrict__ hnew_base, int k0)
{
    const int jl = 32 * q + lane;
    const float br = bias[jl], bz = bias[128 + jl], bi = bias[256 + jl], bh = bias[384 + jl];
    const uint32_t koff = (uint32_t)((k0 + jl) >> 3) * (NU * 16) + (uint32_t)(jl & 7) * 2;
    constexpr int HALF = NU / 2;
#pragma unroll 1
    for (int c = 0; c < HALF / 16; ++c) {
        const int col0 = hsel * HALF + 16 * c;
        const uint32_t ta = tb + ((uint32_t)(32 * q) << 16) + (uint32_t)col0;
        uint32_t R[16], Z[16], NI[16], NH[16];
        umma::tmem_ld16(ta, R);
        umma::tmem_ld16(ta + NU, Z);
        umma::tmem_ld16(ta + 2 * NU, NI);
        umma::tmem_ld16(ta + 3 * NU, NH);
        umma::tmem_ld16_wait(R);
        reg_fence16(Z); reg_fence16(NI); reg_fence16(NH);
#pragma unroll
        for (int t = 0; t < 16; ++t) {
            const uint32_t off = koff + (uint32_t)(col0 + t) * 16;
            const float hold = bf16_bits_to_float(*reinterpret_cast<const unsigned short *>(hold_base + off));
            const float r = sigmoid_fast(__fadd_rn(__uint_as_float(R[t]), br));
            const float z = sigmoid_fast(__fadd_rn(__uint_as_float(Z[t]), bz));
            const float n = tanh_fast(__fmaf_rn(r, __fadd_rn(__uint_as_float(NH[t]), bh), __fadd_rn(__uint_as_float(NI[t]), bi)));
            const float hn = __fmaf_rn(z, __fsub_rn(hold, n), n);
            *reinterpret_cast<unsigned short *>(hnew_base + off) = float_to_bf16_bits(hn);
        }
    }
}

template <int NU>
__global__ void __launch_bounds__(kBThreads, 1) encode_bf16_kernel(EncodeParams P)
{
    using S = SmemB<NU>;
    constexpr int NE = (NU * 20 + kComputeThreads - 1) / kComputeThreads;
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *x1[2] = {smem + S::offX1a, smem + S::offX1b};
    unsigned char *h2t = smem + S::offH2;
    float *bias = reinterpret_cast<float *>(smem + S::offBias);
    float *wfc = reinterpret_cast<float *>(smem + S::offFc);
    float *bfc = wfc + kFc * kLdFcB;
    float *rs = reinterpret_cast<float *>(smem + S::offRs);
    float *rq = reinterpret_cast<float *>(smem + S::offRq);
    float *m1s = reinterpret_cast<float *>(smem + S::offMisc);
    float *m2s = m1s + NU;
    int *idx0s = reinterpret_cast<int *>(m2s + NU);
    int *idx1s = idx0s + NU;
    int *idx2s = idx1s + NU;
    int *listA = idx2s + NU;
    int *listB = listA + NU;
    int *counts = listB + NU;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(counts + 4);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + S::offBars);
    uint64_t *empty = full + kBStages;
    uint64_t *acc_full = empty + kBStages;     // [2]: the gate accumulators are double-buffered in tensor memory (4 NU columns
    uint64_t *acc_empty = acc_full + 2;        // [2]   each): the MMAs of pass p + 1 run under the gate epilogue of pass p
    uint64_t *act_ready = acc_empty + 2;
    uint64_t *vq_done = act_ready + 1;        // compute warps -> weight producer: the screen no longer uses the weight ring

    const int tid = threadIdx.x;
    // The warp index through a shuffle: the compiler then KNOWS it is the same in all lanes.  With tid >> 5 every branch on
    // the warp's role counted as divergent, and inside such a region every tcgen05.mma was issued through ELECT + five
    // R2UR.BROADCAST (~200 cycles per MMA instead of ~10: the descriptors did not stay on the uniform datapath).
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
    const int my_tiles = P.ntiles > (int)blockIdx.x ? (P.ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    VqTcShared<S::kNBTotal> *vsh = reinterpret_cast<VqTcShared<S::kNBTotal> *>(smem + S::offVqSh);
    static_assert(sizeof(VqTcShared<S::kNBTotal>) <= 512, "control block");
    if (tid == 0) {
        for (int s = 0; s < kBStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 3); }      // three issuing warps
        for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 3); mbar_init(&acc_empty[b], kComputeThreads / 32); }
        mbar_init(act_ready, kComputeThreads / 32);
        mbar_init(vq_done, 1);
        vq_tc_init<S::kNBTotal>(vsh, kComputeThreads / 32);
        vsh->ring2_addr = smem_u32(smem + S::offRing);
        mbar_fence_init();
    }
    // all 512 columns: the gate accumulators use 4 NU of them; the VQ screen uses all of them while the gate
    // accumulators are dead (between the GRU 2 epilogue and the next frame's activations)
    if (warp == 0) umma::tmem_alloc(tmem_slot, 512);
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tb = *tmem_slot;
    if (tid == 0) {
        vsh->tmem_base = tb;       // the VQ screen reads it from its control block
        if (P.prof != nullptr && blockIdx.x == 0) vsh->trace = P.prof + 8192;      // debug trace (first 64 chunks / units)
    }

    // ---------------- dedicated warpgroup: TMA producer and MMA issuer ----------------
    if (warp >= kComputeThreads / 32) {
        // 256 x 216 + 128 x 72 = 64 512 = 384 x 168 (the launch allocation): the increase can always be granted
        asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
        // (warp first, lane inside: a condition that mixes the lane in would make the issuing warps' whole branch count as
        // divergent for the compiler, with the slow tcgen05.mma issue sequence that goes with it)
        if (warp == kComputeThreads / 32) {
          if (lane == 0) {
            const long long total = (long long)my_tiles * (P.f1 - P.f0) * kBStepsPerFrame;
            const char *src = reinterpret_cast<const char *>(P.wstream);
            int s = 0, gf = 0;
            uint32_t wraps = 0, frames = 0;
            const bool lend = P.mode == kModeQuantize;         // the ring is lent to the VQ screen between two frames
            for (long long g = 0; g < total; ++g) {
                if (lend && gf == 0 && frames > 0) mbar_wait(vq_done, (frames - 1) & 1u);
                if (wraps > 0) mbar_wait(&empty[s], (wraps - 1) & 1u);
                mbar_arrive_expect_tx(&full[s], kBStageBytes);
                bulk_g2s(smem + S::offRing + s * kBStageBytes, src + (size_t)gf * kBStageBytes, kBStageBytes, &full[s]);
                if (++gf == kBStepsPerFrame) { gf = 0; ++frames; }
                if (++s == kBStages) { s = 0; ++wraps; }
            }
          } else if (lane == 1 && P.mode == kModeQuantize) {
            // the codebook images of the VQ screen, streamed by another lane of this warp.  (Not by one of the issuing
            // warps: a region under `lane == ...` inside their frame loop makes the compiler treat the rest of the loop as
            // not converged, and every tcgen05.mma is then issued through ELECT + five R2UR.BROADCAST, ~200 cycles each.)
            VqTcCount vn_stream{0u, 0u, 0u, 0u};
            const long long frames = (long long)my_tiles * (P.f1 - P.f0);
            for (long long f = 0; f < frames; ++f) {
                bool done = false;
                while (!done) done = vq_tc_produce_phase<S::kNBTotal>(vsh, P.cb, vn_stream);
            }
          }
        } else {
            // MMA issuers: warps 9, 10, 11, one GATE each (r, z, n -- separate accumulators, so the split does not touch
            // any summation order).  A single thread needs ~100-200 cycles per tcgen05 instruction (descriptor moves to
            // the uniform datapath, issue), which made the 330 MMAs of a frame as long as the whole GRU phase; three
            // warps issue their 110 concurrently.  Each warp runs its loop converged and one elected lane issues
            // (fpc_umma.cuh).  For the VQ screen warps 9 and 11 issue alternate chunks and warp 10 streams the codebook.
            const int gate = warp - (kComputeThreads / 32 + 1);          // 0 = r, 1 = z, 2 = n
            const uint32_t idesc = umma::instr_desc_bf16(128, NU);
            const uint32_t ring_a = smem_u32(smem + S::offRing);
            const uint32_t x1a0 = smem_u32(x1[0]);       // (the two [x | h1] tiles are adjacent: address arithmetic instead of an indexed
                                                          //  local array keeps the descriptors on the uniform datapath)
            const uint32_t h2a = smem_u32(h2t);
            int s = 0;
            uint32_t ph = 0, n_act = 0, n_acc = 0;
            VqTcCount vn{0u, 0u, 0u, 0u};
            for (int tile = 0; tile < my_tiles; ++tile) {
                int cur = 0;
                for (int fr = P.f0; fr < P.f1; ++fr) {
                    // ---- GRU 1: three passes of 128 hidden units ----
                    mbar_wait(act_ready, n_act & 1u); ++n_act;
                    for (int pass = 0; pass < 3; ++pass) {
                        const uint32_t ab = n_acc & 1u, tacc = tb + ab * (4 * NU);      // accumulator buffer of this pass
                        mbar_wait(&acc_empty[ab], ((n_acc >> 1) & 1u) ^ 1u); ++n_acc;
                        umma::fence_after_sync();
                        for (int i = 0; i < kBG1Steps; ++i) {
                            mbar_wait(&full[s], ph);
                            umma::fence_after_sync();
                            const uint32_t a0 = ring_a + s * kBStageBytes + gate * kBTileBytes;
                            const uint64_t bd = umma::smem_desc(x1a0 + (uint32_t)cur * S::kX1Bytes + (uint32_t)(2 * i) * (NU * 16), NU);
                            // gates r, z: one accumulator over x and h; gate n: n_i over the two x steps, n_h over the h steps
                            const bool nh = gate == 2 && i >= 2;
                            umma::mma_bf16_elect(tacc + (uint32_t)((nh ? 3 : gate) * NU), umma::smem_desc(a0, 128), bd, idesc, nh ? i > 2 : i > 0);
                            umma::commit_elect(&empty[s]);
                            if (++s == kBStages) { s = 0; ph ^= 1u; }
                        }
                        umma::commit_elect(&acc_full[ab]);
                    }
                    // ---- GRU 2: input h1' (the other [x | h1] tile), hidden h2 ----
                    mbar_wait(act_ready, n_act & 1u); ++n_act;
                    const uint32_t ab = n_acc & 1u, tacc = tb + ab * (4 * NU);
                    mbar_wait(&acc_empty[ab], ((n_acc >> 1) & 1u) ^ 1u); ++n_acc;
                    umma::fence_after_sync();
                    for (int i = 0; i < kBG2Steps; ++i) {
                        mbar_wait(&full[s], ph);
                        umma::fence_after_sync();
                        const uint32_t a0 = ring_a + s * kBStageBytes + gate * kBTileBytes;
                        const bool xp = i < kH1 / 16;
                        const uint32_t baddr = xp ? x1a0 + (uint32_t)(cur ^ 1) * S::kX1Bytes + (uint32_t)(4 + 2 * i) * (NU * 16)
                                                  : h2a + (uint32_t)(2 * (i - kH1 / 16)) * (NU * 16);
                        const uint64_t bd = umma::smem_desc(baddr, NU);
                        const bool nh = gate == 2 && !xp;
                        umma::mma_bf16_elect(tacc + (uint32_t)((nh ? 3 : gate) * NU), umma::smem_desc(a0, 128), bd, idesc, nh ? i > kH1 / 16 : i > 0);
                        umma::commit_elect(&empty[s]);
                        if (++s == kBStages) { s = 0; ph ^= 1u; }
                    }
                    umma::commit_elect(&acc_full[ab]);
                    cur ^= 1;
                    // the searches of this frame: stages of the codebooks as published by the compute warps
                    if (P.mode == kModeQuantize && gate != 1) {      // warps 9 and 11 issue alternate chunks
                        bool done = false;
                        while (!done) done = vq_tc_issue_phase<S::kNBTotal>(vsh, tb, vn, lane, gate >> 1);
                    }
                }
            }
        }
        return;
    }

    // ---------------- compute warps ----------------
    asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
    const int q = warp & 3, hsel = warp >> 2;
    const PackedCodebooks *cbh = reinterpret_cast<const PackedCodebooks *>(P.cb);
    {
        const float *tail = reinterpret_cast<const float *>(reinterpret_cast<const char *>(P.wstream) + kBStreamBytes);
        for (int i = tid; i < kBiasFloats; i += kComputeThreads) bias[i] = tail[i];
        for (int i = tid; i < kFcFloats; i += kComputeThreads) wfc[(i >> 7) * kLdFcB + (i & 127)] = tail[kBiasFloats + i];
        if (tid < kFc) bfc[tid] = tail[kBiasFloats + kFcFloats + tid];
        // the scalar tables (<= 256 levels each) next to the state: from global memory every scalar search paid L2 latencies
        if (cbh != nullptr) {
            unsigned char *sclbuf = smem + S::offScl;
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
    uint32_t n_full = 0;
    const bool prof = P.prof != nullptr && tid == 0;
    long long pt[kPhCount] = {}, pt0 = 0;
#define FPC_PHASE(ph) do { if (prof) { const long long t_ = clock64(); pt[ph] += t_ - pt0; pt0 = t_; } } while (0)

    for (int tile = blockIdx.x; tile < P.ntiles; tile += gridDim.x) {
        const int b0 = tile * NU;
        int cur = 0;
        // h1 = h2 = None -> zeros, frame 0 input all zero (wavernn.py:177-178,189)
        // (or, when a launch continues an earlier frame range, the state that launch left: the bf16 B-operand tiles)
        int4 *carry = P.state ? reinterpret_cast<int4 *>(P.state) + (size_t)tile * ((S::kX1Bytes + S::kH2Bytes) / 16) : nullptr;
        if (carry && P.f0 > 0) {
            for (int i = tid; i < S::kX1Bytes / 16; i += kComputeThreads) reinterpret_cast<int4 *>(x1[0])[i] = carry[i];
            for (int i = tid; i < S::kH2Bytes / 16; i += kComputeThreads) reinterpret_cast<int4 *>(h2t)[i] = carry[S::kX1Bytes / 16 + i];
        } else {
            for (int i = tid; i < S::kX1Bytes / 16; i += kComputeThreads) reinterpret_cast<int4 *>(x1[0])[i] = make_int4(0, 0, 0, 0);
            for (int i = tid; i < S::kH2Bytes / 16; i += kComputeThreads) reinterpret_cast<int4 *>(h2t)[i] = make_int4(0, 0, 0, 0);
        }
        for (int i = tid; i < NU * kLdR; i += kComputeThreads) rs[i] = 0.0f;
        umma::fence_async_smem();
        named_bar_sync(1, kComputeThreads);
        if (lane == 0) mbar_arrive(act_ready);

        for (int fr = P.f0; fr < P.f1; ++fr) {
            float featv[NE], fov[NE], rsv[NE];
#pragma unroll
            for (int e2 = 0; e2 < NE; ++e2) {
                const int e = tid + kComputeThreads * e2;
                const int u = e / 20, j = e - u * 20;
                featv[e2] = 0.0f; fov[e2] = 0.0f; rsv[e2] = 0.0f;
                if (e < NU * 20 && b0 + u < P.B) {
                    const size_t fo = (size_t)(b0 + u) * P.L + fr;
                    if (P.mode != kModeDecode) featv[e2] = __ldg(P.feat + fo * 20 + j);
                    else featv[e2] = j < kFc ? __ldg(P.rq_in + fo * kFc + j) : __ldg(P.pitch_in + fo * 2 + (j - kFc));
                }
            }
            if (prof) pt0 = clock64();
            // ---- GRU 1 (wavernn.py:71): gate epilogues of the three passes ----
#pragma unroll 1
            for (int pass = 0; pass < 3; ++pass) {
                const uint32_t ab = n_full & 1u;
                mbar_wait(&acc_full[ab], (n_full >> 1) & 1u); ++n_full;
                umma::fence_after_sync();
                gate_epilogue<NU>(tb + ab * (4 * NU), q, hsel, lane, bias + pass * 512, x1[cur], x1[cur ^ 1], 32 + pass * 128);
                umma::fence_before_sync();
                if (pass == 2) umma::fence_async_smem();     // h1' is the B operand of GRU 2
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(&acc_empty[ab]);
                    if (pass == 2) mbar_arrive(act_ready);
                }
            }
            // ---- GRU 2 (wavernn.py:76), state updated in place ----
            const uint32_t ab2 = n_full & 1u;
            mbar_wait(&acc_full[ab2], (n_full >> 1) & 1u); ++n_full;
            umma::fence_after_sync();
            gate_epilogue<NU>(tb + ab2 * (4 * NU), q, hsel, lane, bias + 3 * 512, h2t, h2t, 0);
            umma::fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[ab2]);
            named_bar_sync(1, kComputeThreads);

            FPC_PHASE(kPhGru);
            // ---- relu, dual_fc, 2*tanh (wavernn.py:87-92); residual (:196) ----
#pragma unroll
            for (int e2 = 0; e2 < NE; ++e2) {
                const int e = tid + kComputeThreads * e2;
                const int u = e / 20, j = e - u * 20;
                if (e < NU * 20 && j < kFc) {
                    float a = bfc[j];
                    const float *wr = wfc + j * kLdFcB;
#pragma unroll 4
                    for (int kc = 0; kc < kH2 / 8; ++kc) {
                        const int4 pk = *reinterpret_cast<const int4 *>(h2t + (size_t)kc * (NU * 16) + u * 16);
                        const uint32_t wds[4] = {(uint32_t)pk.x, (uint32_t)pk.y, (uint32_t)pk.z, (uint32_t)pk.w};
#pragma unroll
                        for (int h = 0; h < 4; ++h) {
                            const float lo = __uint_as_float(wds[h] << 16), hi = __uint_as_float(wds[h] & 0xffff0000u);
                            a = __fmaf_rn(wr[kc * 8 + 2 * h], fmaxf(lo, 0.0f), a);
                            a = __fmaf_rn(wr[kc * 8 + 2 * h + 1], fmaxf(hi, 0.0f), a);
                        }
                    }
                    const float f = __fmul_rn(2.0f, tanh_c(a));
                    fov[e2] = f;
                    if (P.mode != kModeDecode) {
                        rsv[e2] = __fsub_rn(featv[e2], f);
                        rs[u * kLdR + 3 + j] = rsv[e2];
                    }
                }
            }
            if (P.mode != kModeDecode) {
                for (int i = tid; i < NU * 20; i += kComputeThreads) rq[i] = 0.0f;
                named_bar_sync(1, kComputeThreads);

                FPC_PHASE(kPhFc);
                // ---- indicators (:201-212) and the scalar quantiser for c0 (:217-225) ----
                // four lanes per utterance (256 threads = 64 utterances at once; a 32-utterance tile uses eight)
                {
                    constexpr int G = kComputeThreads / NU;                // 4 or 8 lanes per utterance
                    const int u = tid / G, part = tid % G;
                    const bool valid = b0 + u < P.B;
                    float m1 = 0.0f, m2 = 0.0f;
                    if (part == 0 && valid) {
                        if (P.mask == nullptr) {
                            float sacc = 0.0f;
#pragma unroll
                            for (int j = 1; j < kFc; ++j) sacc = __fadd_rn(sacc, fabsf(rs[u * kLdR + 3 + j]));
                            m1 = fabsf(rs[u * kLdR + 3]) > P.l1 ? 1.0f : 0.0f;
                            m2 = sacc > P.l2 ? 1.0f : 0.0f;
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
                        const unsigned char *sclt = smem + S::offScl + which * (FPC_MAX_SCL_ENTRIES * 8);
                        const int *meta = reinterpret_cast<const int *>(smem + S::offScl + 2 * FPC_MAX_SCL_ENTRIES * 8) + 2 * which;
                        const int sn = valid ? meta[0] : 0, sdt = meta[1];
                        const float x0 = rs[u * kLdR + 3];
                        // (the tables of one warp's utterances may differ in dtype only between above / below: both branches are
                        //  executed by the lanes that need them, with the group shuffles inside kept convergent per group)
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
                    if (part == 0) {
                        if (coded) rq[u * 20] = qv;
                        m1s[u] = m1; m2s[u] = m2; idx0s[u] = i0; idx1s[u] = -1; idx2s[u] = -1;
                    }
                }
                named_bar_sync(1, kComputeThreads);

                FPC_PHASE(kPhScalar);
                if (P.mode == kModeQuantize) {
                    // ---- VQ for c1..c17 (:228-240): rows compacted by branch ----
                    if (warp == 0) {
                        int na = 0, nb = 0;
#pragma unroll
                        for (int base = 0; base < NU; base += 32) {
                            const int u = base + lane;
                            const bool valid = b0 + u < P.B;
                            const bool above = valid && m2s[u] != 0.0f;
                            const bool below = valid && !above && cbh->bl.stages > 0;
                            const unsigned ba = __ballot_sync(0xffffffffu, above);
                            const unsigned bb = __ballot_sync(0xffffffffu, below);
                            const unsigned lt = (1u << lane) - 1u;
                            if (above) listA[na + __popc(ba & lt)] = u;
                            if (below) listB[nb + __popc(bb & lt)] = u;
                            na += __popc(ba);
                            nb += __popc(bb);
                        }
                        if (lane == 0) { counts[0] = na; counts[1] = nb; }
                    }
                    named_bar_sync(1, kComputeThreads);
                    const int nA = counts[0], nB = counts[1];
                    char *scratch = reinterpret_cast<char *>(x1[cur]);   // old [x | h1]: dead once GRU 1 has finished
                    VqTcMem vmem;
                    unsigned char *abase = S::kAInScratch ? reinterpret_cast<unsigned char *>(scratch) : smem + S::offVqA;
                    vmem.a = abase;
                    vmem.small = S::kSmallInScratch ? scratch + S::kABytes : reinterpret_cast<char *>(smem) + S::offVqSmall;
                    vmem.part = smem + S::offPart;
                    vmem.bring = smem + S::offBring;
                    vmem.scratch = scratch;
                    vmem.scratch_bytes = S::kScratchBytes;
                    vmem.tail = reinterpret_cast<int *>(smem + S::offPart);   // idle by the time the fallback runs
                    // one call site for both books (above / below threshold): a single inlined copy of the search.  The
                    // helper roles are owed exactly one publication with last = 1 per frame.
                    const bool tcB = nB > 0 && cbh->bl.K >= 64;
#pragma unroll 1
                    for (int book = 0; book < 2; ++book) {
                        const int nrows = book ? nB : nA;
                        if (nrows > 0)
                            vq_tc_dispatch<NU, S::kNBTotal>(book ? cbh->bl : cbh->vq, P.cb, book ? listB : listA, nrows, rs, rq, idx1s, idx2s,
                                                       vmem, vsh, (book == 1 || !tcB) ? 1 : 0, tid, prof ? pt + kPhVqDbg : nullptr);
                    }
                    if (!(nA > 0 && cbh->vq.K >= 64) && !tcB) vq_tc_publish_idle<S::kNBTotal>(vsh, tid);
                    // every accumulator unit of the frame has been read, so every MMA that read a codebook chunk is complete:
                    // the weight producer may refill its ring
                    if (tid == 0) mbar_arrive(vq_done);
                }
            } else {
                named_bar_sync(1, kComputeThreads);
            }

            FPC_PHASE(kPhVq);
            // ---- feedback (:242 / :252), outputs, next input frame (bf16 B-operand tile) ----
            unsigned char *xn = x1[cur ^ 1];
#pragma unroll
            for (int e2 = 0; e2 < NE; ++e2) {
                const int e = tid + kComputeThreads * e2;
                const int u = e / 20, j = e - u * 20;
                if (e < NU * 20) {
                    const bool valid = b0 + u < P.B;
                    const size_t fo = (size_t)(b0 + u) * P.L + fr;
                    float cin;
                    if (j < kFc) {
                        float ro, rqo, ruo;
                        if (P.mode == kModeQuantize) {
                            rqo = rq[u * 20 + j];
                            ro = rsv[e2];
                            ruo = 0.0f;
                            cin = __fadd_rn(fov[e2], rqo);
                        } else if (P.mode == kModeResidual) {
                            const float m = j == 0 ? m1s[u] : m2s[u];
                            ruo = __fmul_rn(rsv[e2], __fsub_rn(1.0f, m));
                            ro = __fmul_rn(rsv[e2], m);
                            rqo = 0.0f;
                            cin = __fadd_rn(fov[e2], ro);
                        } else {
                            ro = rqo = ruo = 0.0f;
                            cin = __fadd_rn(fov[e2], featv[e2]);
                        }
                        if (valid && P.mode != kModeDecode) {
                            P.r[fo * kFc + j] = ro;
                            P.r_qtz[fo * kFc + j] = rqo;
                            if (P.r_under) P.r_under[fo * kFc + j] = ruo;
                        }
                    } else {
                        cin = featv[e2];
                    }
                    *reinterpret_cast<unsigned short *>(xn + (size_t)(j >> 3) * (NU * 16) + u * 16 + (j & 7) * 2) = float_to_bf16_bits(cin);
                    if (valid) P.c_in[fo * 20 + j] = cin;
                }
            }
            // K padding 20..31 of the x part: the tile served as VQ scratch, so rewrite the zeros
            for (int i = tid; i < NU * 12; i += kComputeThreads) {
                const int u = i / 12, k = 20 + (i - u * 12);
                *reinterpret_cast<unsigned short *>(xn + (size_t)(k >> 3) * (NU * 16) + u * 16 + (k & 7) * 2) = 0;
            }
            if (P.mode != kModeDecode && tid < NU && b0 + tid < P.B) {
                const size_t fo = (size_t)(b0 + tid) * P.L + fr;
                const float m1 = m1s[tid], m2 = m2s[tid];
                if (P.ind1) P.ind1[fo] = P.mask ? 0.0f : m1;
                if (P.ind2) P.ind2[fo] = P.mask ? 0.0f : m2;
                if (P.idx) {
                    int4 v;
                    v.x = idx0s[tid]; v.y = idx1s[tid]; v.z = idx2s[tid];
                    v.w = (m1 != 0.0f ? 1 : 0) | (m2 != 0.0f ? 2 : 0);
                    *reinterpret_cast<int4 *>(P.idx + fo * 4) = v;
                }
            }
            umma::fence_async_smem();
            named_bar_sync(1, kComputeThreads);
            if (lane == 0 && fr + 1 < P.f1) mbar_arrive(act_ready);
            FPC_PHASE(kPhOut);
            if (prof) pt[kPhFrames] += 1;
            cur ^= 1;
        }
        if (carry) {       // x1[cur] = [next input frame | h1], h2t = h2: everything the next frame range needs
            for (int i = tid; i < S::kX1Bytes / 16; i += kComputeThreads) carry[i] = reinterpret_cast<const int4 *>(x1[cur])[i];
            for (int i = tid; i < S::kH2Bytes / 16; i += kComputeThreads) carry[S::kX1Bytes / 16 + i] = reinterpret_cast<const int4 *>(h2t)[i];
            named_bar_sync(1, kComputeThreads);
        }
    }
    if (prof)
        for (int i = 0; i < kPhCount; ++i) atomicAdd(reinterpret_cast<unsigned long long *>(P.prof) + (size_t)blockIdx.x * kPhCount + i, (unsigned long long)pt[i]);
#undef FPC_PHASE
    named_bar_sync(1, kComputeThreads);
    if (warp == 0) umma::tmem_dealloc(tb, 512);
}

template <int NU>
static int launch_encode_bf16(const EncodeParams &P, int grid, cudaStream_t st)
{
    using S = SmemB<NU>;
    static bool configured[kMaxDevices] = {};
    { const int rc = ensure_dynamic_smem(encode_bf16_kernel<NU>, S::total, configured); if (rc != FPC_OK) return rc; }
    encode_bf16_kernel<NU><<<grid, kBThreads, S::total, st>>>(P);
    FPC_LAUNCH_CHECK();
    return FPC_OK;
}

static const int kHeightsBf16[2] = {64, 32};
// per utterance: the [x(32) | h1(384)] and h2(128) bf16 rows
static size_t state_bytes_per_tile_bf16(int nu) { return (size_t)nu * (size_t)(kXK + kH2) * 2; }

size_t encode_bf16_state_bytes(int B)
{
    const int sms = num_sms();
    if (sms <= 0 || B <= 0) return 0;
    EncodeSegment seg[kMaxSegments];
    const int n = plan_segments(B, sms, kHeightsBf16, 2, 6.0, seg);
    size_t total = 0;
    for (int i = 0; i < n; ++i) total += (size_t)((seg[i].count + seg[i].height - 1) / seg[i].height) * state_bytes_per_tile_bf16(seg[i].height);
    return total;
}

int run_encode_bf16(EncodeParams P, cudaStream_t st, int force_nu)
{
    if (P.f0 < 0 || P.f1 > P.L || P.f0 >= P.f1) return FPC_ERR_ARG;
    const int sms = num_sms();
    if (sms <= 0) return cuda_fail(cudaErrorNoDevice);
    // launch plan as in the fp32 kernel (fpc_encode_fp32.cu): whole waves of 64-utterance tiles, the rest with
    // the height that needs the fewest tile-frames
    EncodeSegment seg[kMaxSegments];
    int n = 1;
    if (force_nu > 0) seg[0] = EncodeSegment{force_nu, 0, P.B};
    else n = plan_segments(P.B, sms, kHeightsBf16, 2, 6.0, seg);
    char *state = reinterpret_cast<char *>(P.state);
    for (int i = 0; i < n; ++i) {
        EncodeParams Q = segment_params(P, seg[i].first, seg[i].count);
        const int nu = seg[i].height;
        Q.ntiles = (Q.B + nu - 1) / nu;
        Q.state = state;
        if (state) state += (size_t)Q.ntiles * state_bytes_per_tile_bf16(nu);
        const int grid = Q.ntiles < sms ? Q.ntiles : sms;
        int rc;
        if (nu == 32) rc = launch_encode_bf16<32>(Q, grid, st);
        else if (nu == 64) rc = launch_encode_bf16<64>(Q, grid, st);
        else rc = FPC_ERR_ARG;
        if (rc != FPC_OK) return rc;
    }
    return FPC_OK;
}

}  // namespace fpc
