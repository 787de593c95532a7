// fpc_vq_tc.cuh -- the m-best VQ search of the fused frame-step kernels with its distance screen on the tensor cores.
//
// Same result, bit for bit, as vq_search_rows_screened (fpc_vq_screen.cuh) and therefore as the reference's
// quantize_mstage (/root/reference/src/quantization/vq_func.py:82-131): this file only replaces the fp32 FMA screen
// (17 FMAs per vector-codeword pair on the CUDA cores) by tcgen05 GEMMs of fp16-pair operands (fpc_tc.cuh) and a scan
// of the accumulators in TMEM.  Decisions are taken only when the screened values prove them.  Margin: M = 2^-16 R =
// 256 u R with u = 2^-24, R = (||x|| + Cmax0 + Cmax1)^2.  A screened value is within ~35 u R of the real one (operand
// format 10 u R, accumulation measured <= 3 u R and asserted <= 64 u R in tests/test_gpu_quant_kmeans.py, the fp32 norm of
// the stage-1 search vector 17 u R, its rounding 2 u R) and numpy's own evaluation within 20 u R, so two candidates are
// compared to within 110 u R.  Rows the screen cannot decide have their candidates re-ranked with the reference's
// exact arithmetic when all of them are known, and go to the exact block-wide search (vq_search_rows) otherwise.
//
// Work split.  The tile's 8 compute warps build the A operand (search vectors, scaled by -2 beta, as fp16 pairs), scan
// the accumulators and decide.  One elected thread of another warp issues the MMAs, one streams the codebook image
// (B operand, 8 KB per 64 codewords) from L2 through a small ring with bulk async copies.  The three roles meet at
// mbarriers: `go` (a phase -- one stage of one book -- is published with its chunk count, tile count and addresses),
// `ack`, ring full/empty, TMEM unit full/empty.  TMEM: 8 units of 64 columns; a unit is (chunk of 64 codewords, tile of
// 128 rows), lane = row, column = codeword.
//   stage 0 of a two-stage book: rows = the vectors, replicated over the lane quarters so that all warps scan (each
//     replica takes every rep-th chunk); scores carry a per-vector offset so they are positive, and every thread keeps
//     the six smallest packed keys (value | codeword) of its columns; a warp per vector then extracts the eight
//     smallest of the row's lists and applies the survivor-set logic of fpc_vq_screen.cuh (margin, exact re-rank of
//     near-ties with dist17<T>).
//   last stage: rows = (vector, survivor) pairs x' = x - c0[k_s]; the scan keeps the smallest score and the runner-up
//     per row without indices (tc::Scan); per vector the five rows are merged with their ||x'||^2.
#pragma once
#include "fpc_tc.cuh"
#include "fpc_vq_screen.cuh"

// The searches are out of line by default (one copy per dtype).  A kernel that moves registers between its warpgroups with
// setmaxnreg.inc on a path that does NOT contain the call cannot use out-of-line device functions (cicc 12.9 crashes on
// that combination, and a callee is compiled against the launch allocation anyway): it defines FPC_VQ_TC_OUTLINE as
// __forceinline__ before including this file.
#ifndef FPC_VQ_TC_OUTLINE
#define FPC_VQ_TC_OUTLINE __noinline__
#endif

namespace fpc {

constexpr int kVtUnits = 8;                          // TMEM units of 64 columns
constexpr int kVtChunkBytes = 64 * tc::kK * 2;       // one 64-codeword tile of a B image
constexpr int kVtSlabBytes = kVtChunkBytes / 4;      // its K = 16 slab
// scan partials: [tile][row][column half] (best, second, column, -) for up to `mtmax` tiles, or 6 keys per thread
// (+ the eight smallest keys and a mark per vector behind them)
__host__ __device__ constexpr int vq_tc_part_bytes(int mtmax) { return mtmax * 128 * 2 * 16 > 256 * 24 + 64 * 36 ? mtmax * 128 * 2 * 16 : 256 * 24 + 64 * 36; }

// nchunks / mtiles / a_addr: the MMA job of this phase.  p_nchunks / b_off: the streaming job that STARTS with this phase
// (stage 0 of a two-stage book publishes the chunks of both stages -- their images are contiguous -- so that the last
// stage's codewords arrive while the survivors are still being selected; the last stage then publishes p_nchunks = 0).
struct VqTcCtl { int nchunks, mtiles, last, p_nchunks; uint32_t a_addr, ring_addr; long long b_off; };   // addresses: shared window

// control block + barriers; lives in shared memory for the whole kernel (phase parities run across frames)
template <int NB, int UNITS = kVtUnits> struct VqTcShared {
    VqTcCtl ctl;
    // The compute warps' running counts and the TMEM base live HERE, not in registers: they would be live across the
    // whole frame loop, and the gate GEMM of the fp32 kernel has no register to spare (six more spilled inside its
    // inner loop and cost 12 % of the frame).
    uint32_t tmem_base, cnt_go, cnt_b, cnt_u;
    // ring slots beyond the first kVtOwnSlots live elsewhere in shared memory (the bf16 kernel lends the idle weight ring
    // of the gate GEMMs to the screen); 0 = none.  Set once by the kernel.
    uint32_t ring2_addr;
    long long *trace;             // debug (tools/phase_profile.py): event times of the first phases of CTA 0, or null
    uint64_t go, ack;
    uint64_t b_full[NB], b_empty[NB];
    uint64_t d_full[UNITS], d_empty[UNITS];
};
constexpr int kVtOwnSlots = 2;
__device__ __forceinline__ uint32_t vq_tc_slot_addr(uint32_t ring, uint32_t ring2, int slot)
{
    return slot < kVtOwnSlots ? ring + (uint32_t)slot * kVtChunkBytes : ring2 + (uint32_t)(slot - kVtOwnSlots) * kVtChunkBytes;
}
struct VqTcCount { uint32_t go, b, u, ring; };       // running use counts of one role (phases, B chunks, TMEM units); ring address in use

template <int NB, int UNITS>
__device__ __forceinline__ void vq_tc_init(VqTcShared<NB, UNITS> *sh, int compute_warps)
{
    sh->cnt_go = sh->cnt_b = sh->cnt_u = 0u;
    sh->ring2_addr = 0u;
    sh->trace = nullptr;
    mbar_init(&sh->go, 1);
    mbar_init(&sh->ack, 3);            // two issuing warps + the streamer
    for (int i = 0; i < NB; ++i) { mbar_init(&sh->b_full[i], 1); mbar_init(&sh->b_empty[i], 1); }
    for (int i = 0; i < UNITS; ++i) { mbar_init(&sh->d_full[i], 1); mbar_init(&sh->d_empty[i], compute_warps); }
}

// The helper threads sit at `go` for the whole predictor phase of every frame.  They wait with a long suspend-time
// hint: the hardware parks the thread until the phase completes (or the hint expires), so it neither polls nor fetches
// instructions next to the gate GEMM of the compute warps on the same scheduler.
__device__ __forceinline__ void mbar_wait_idle(uint64_t *bar, uint32_t parity)
{
    const long long t0 = clock64();
    for (;;) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity), "r"(10000000u)
            : "memory");
        if (ok) return;
        if (clock64() - t0 > 16000000000LL) __trap();
    }
}

// ---- producer thread: one phase.  Returns true after the last phase of a frame ----
template <int NB, int UNITS>
__device__ __forceinline__ bool vq_tc_produce_phase(VqTcShared<NB, UNITS> *sh, const char *__restrict__ cbbase, VqTcCount &n)
{
    mbar_wait_idle(&sh->go, n.go & 1u); ++n.go;
    const int nchunks = sh->ctl.p_nchunks, last = sh->ctl.last;
    const long long off = sh->ctl.b_off;
    const uint32_t ring = sh->ctl.ring_addr, ring2 = sh->ring2_addr;
    mbar_arrive(&sh->ack);
    for (int c = 0; c < nchunks; ++c, ++n.b) {
        const int bs = (int)(n.b % NB);
        const uint32_t use = n.b / NB;
        if (use > 0) mbar_wait(&sh->b_empty[bs], (use - 1) & 1u);
        if (sh->trace && n.b < 64) sh->trace[n.b] = clock64();
        mbar_arrive_expect_tx(&sh->b_full[bs], kVtChunkBytes);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(vq_tc_slot_addr(ring, ring2, bs)),
                     "l"(cbbase + off + (long long)c * kVtChunkBytes), "r"((uint32_t)kVtChunkBytes), "r"(smem_u32(&sh->b_full[bs]))
                     : "memory");
    }
    return last != 0;
}

// Event times of the first chunks for tools/phase_profile.py -- in debug builds only (-DFPC_VQ_TRACE_ON): a branch on the
// lane inside the issue loop makes the compiler treat the whole loop as divergent, and every tcgen05.mma is then issued
// through ELECT + five R2UR.BROADCAST (~200 cycles per MMA) instead of straight from uniform registers.
#ifdef FPC_VQ_TRACE_ON
#define FPC_VQ_TRACE(slot) do { if (lane == 0 && sh->trace && n.b < 64) sh->trace[(slot) + n.b] = clock64(); } while (0)
#else
#define FPC_VQ_TRACE(slot) do { } while (0)
#endif

// ---- MMA issuer WARPS (two of them, `which` = 0 / 1, alternate chunks; all 32 lanes run the loop converged and one
//      elected lane issues): one phase ----
template <int NB, int UNITS>
__device__ __forceinline__ bool vq_tc_issue_phase(VqTcShared<NB, UNITS> *sh, uint32_t tb, VqTcCount &n, int lane, int which)
{
    mbar_wait_idle(&sh->go, n.go & 1u); ++n.go;
    const int nchunks = sh->ctl.nchunks, mtiles = sh->ctl.mtiles, last = sh->ctl.last;
    const uint32_t a_addr = sh->ctl.a_addr;
    // the ring address is the one the streaming job was published with; a phase without a streaming job of its own (the
    // last stage of a two-stage book) continues in the ring of the previous phase
    if (sh->ctl.p_nchunks > 0) n.ring = sh->ctl.ring_addr;
    const uint32_t bring_addr = n.ring;
    __syncwarp();
    if (lane == 0) mbar_arrive(&sh->ack);
    umma::fence_after_sync();
    const uint32_t idesc = tc::instr_desc_f16(128, 64);
    // descriptors differ only in their address field (bytes >> 4, low 14 bits): one base each, then additions
    const uint64_t adesc0 = umma::smem_desc(a_addr, 128);
    const uint32_t ring2 = sh->ring2_addr;
    for (int c = 0; c < nchunks; ++c, ++n.b) {
        if ((c & 1) != which) { n.u += mtiles; continue; }       // the other issuing warp's chunk
        const int bs = (int)(n.b % NB);
        mbar_wait(&sh->b_full[bs], (n.b / NB) & 1u);
        FPC_VQ_TRACE(64);
        umma::fence_after_sync();
        FPC_VQ_TRACE(192);
        const uint64_t bd = umma::smem_desc(vq_tc_slot_addr(bring_addr, ring2, bs), 64);
        // tile by tile, each tile's four K = 16 slabs back to back and its commit right behind them: a dependent chain of
        // MMAs into one accumulator costs nothing (tools/ubench_umma.cu: 48 cycles per MMA either way), and the first
        // tile of a chunk is handed to the scanning warps while the tensor pipe still works on the others
#pragma unroll
        for (int m = 0; m < 3; ++m)
            if (m < mtiles) {
                const uint32_t use = (n.u + m) / UNITS;
                if (use > 0) { mbar_wait(&sh->d_empty[(n.u + m) % UNITS], (use - 1) & 1u); umma::fence_after_sync(); }
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                    umma::mma_bf16_elect(tb + ((n.u + m) % UNITS) * 64u, adesc0 + (uint64_t)((m * tc::kTileBytes + ks * tc::kSlabBytes) >> 4),
                                         bd + (uint64_t)((ks * kVtSlabBytes) >> 4), idesc, ks > 0);
                umma::commit_elect(&sh->d_full[(n.u + m) % UNITS]);
            }
        FPC_VQ_TRACE(320);
        umma::commit_elect(&sh->b_empty[bs]);
        FPC_VQ_TRACE(128);
        n.u += mtiles;
    }
    return last != 0;
}

// ---- compute warps: publish a phase (A tiles are complete) ----
template <int NB, int NT = kComputeThreads, int UNITS>
__device__ __forceinline__ void vq_tc_publish(VqTcShared<NB, UNITS> *sh, VqTcCount &n, int nchunks, int mtiles, uint32_t a_addr, int p_nchunks,
                                              long long b_off, uint32_t ring_addr, int last, int tid)
{
    umma::fence_async_smem();                    // generic-proxy writes of the A tiles -> async proxy (tcgen05.mma reads)
    named_bar_sync(1, NT);
    if (tid == 0) {
        if (n.go > 0) mbar_wait(&sh->ack, (n.go - 1) & 1u);      // both other roles have read the previous control word
        sh->ctl.nchunks = nchunks; sh->ctl.mtiles = mtiles; sh->ctl.last = last; sh->ctl.a_addr = a_addr;
        sh->ctl.p_nchunks = p_nchunks; sh->ctl.b_off = b_off; sh->ctl.ring_addr = ring_addr;
        mbar_arrive(&sh->go);
    }
    ++n.go;
}
// a frame without any search still tells the other two roles that the frame is over
template <int NB, int NT = kComputeThreads, int UNITS>
__device__ FPC_VQ_TC_OUTLINE void vq_tc_publish_idle(VqTcShared<NB, UNITS> *sh, int tid)
{
    VqTcCount n{sh->cnt_go, 0u, 0u, 0u};
    vq_tc_publish<NB, NT, UNITS>(sh, n, 0, 0, 0u, 0, 0, 0u, 1, tid);
    named_bar_sync(1, NT);          // every thread has read cnt_go
    if (tid == 0) sh->cnt_go = n.go;
    named_bar_sync(1, NT);
}

// Six smallest packed keys of a stream: new t_i = min(t_i, max(t_{i-1}, x)), all from the old values (11 instructions).
// The keys are positive fp32 bit patterns (value bits | codeword index in the low 10 mantissa bits) and are compared
// as floats: positive floats order like their bit patterns, and FMNMX issues at twice the rate of the integer VIMNMX.
struct Top6 {
    float t0, t1, t2, t3, t4, t5;
    __device__ __forceinline__ void reset() { t0 = t1 = t2 = t3 = t4 = t5 = __int_as_float(0x7f800000); }
    __device__ __forceinline__ void insert(float x)
    {
        const float n5 = fminf(t5, fmaxf(t4, x)), n4 = fminf(t4, fmaxf(t3, x)), n3 = fminf(t3, fmaxf(t2, x)), n2 = fminf(t2, fmaxf(t1, x)),
                    n1 = fminf(t1, fmaxf(t0, x));
        t0 = fminf(t0, x); t1 = n1; t2 = n2; t3 = n3; t4 = n4; t5 = n5;
    }
};

// Stage-0 selection helper: ranks the six keys of list `l` of vector `v` among the NL lists of the row; a key of rank
// r < 8 is the (r+1)-th smallest of the row.  A list only kept its six smallest keys, so whatever it dropped is larger
// than its sixth key: r6[v] = smallest rank of any list's sixth key tells the decision how many of the row's smallest
// keys are certainly complete (a candidate set that reaches a sixth key may miss a seventh of that list).
// (kHalves = 2: 256 scanning threads, list l = (replica l >> 1, column half l & 1), kept by thread 32 (quarter + 4 half) + lane;
//  kHalves = 1: 128 scanning threads, list l = replica l, kept by thread `row`.)
template <int NL, int kHalves = 2>
__device__ __forceinline__ void vq_tc_rank_list(const unsigned *__restrict__ keys, int P2, int v, int l, unsigned *__restrict__ g8,
                                                int *__restrict__ exh)
{
    unsigned mine[6];
    int rank[6] = {0, 0, 0, 0, 0, 0};
    {
        const int row = (kHalves == 2 ? (l >> 1) : l) * P2 + v;
        const unsigned *kp = keys + (kHalves == 2 ? (32 * (((row >> 5) & 3) + 4 * (l & 1)) + (row & 31)) : row) * 6;
#pragma unroll
        for (int p6 = 0; p6 < 6; ++p6) mine[p6] = kp[p6];
    }
#pragma unroll
    for (int o = 0; o < NL; ++o) {
        const int row = (kHalves == 2 ? (o >> 1) : o) * P2 + v;
        const unsigned *kp = keys + (kHalves == 2 ? (32 * (((row >> 5) & 3) + 4 * (o & 1)) + (row & 31)) : row) * 6;
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            const unsigned k = kp[j];
#pragma unroll
            for (int p6 = 0; p6 < 6; ++p6) rank[p6] += k < mine[p6] ? 1 : 0;
        }
    }
#pragma unroll
    for (int p6 = 0; p6 < 6; ++p6)
        if (rank[p6] < 8) {
            g8[v * 8 + rank[p6]] = mine[p6];
            if (p6 == 5) atomicMin(&exh[v], rank[p6]);
        }
}

// Exact search for the rows the screen could not decide (and for books too small for the screen): the block-wide
// vq_search_rows (fpc_vq_search.cuh), the reference's arithmetic for every candidate.  Out of line: one copy per dtype.
template <typename T, int NT = kComputeThreads>
__device__ FPC_VQ_TC_OUTLINE void vq_tc_fallback(const PackedVq &bk, const char *__restrict__ cbbase, const int *__restrict__ flist, int nflag, int maxn,
                                            const float *__restrict__ rs, float *__restrict__ rq, int *__restrict__ idx1, int *__restrict__ idx2,
                                            char *__restrict__ scratch, int scratch_bytes, int tid)
{
    (void)maxn;
    int sb = 32;
    while (sb > 8 && (int)vq_fixed_bytes<T>(sb) + 1024 * (int)sizeof(T) > scratch_bytes) sb >>= 1;
    int vbe = (scratch_bytes - (int)vq_fixed_bytes<T>(sb)) / (1024 * (int)sizeof(T));
    vbe = vbe > 8 ? 8 : vbe;
    for (int off = 0; off < nflag; off += sb)
        vq_search_rows<T, NT>(bk, cbbase, flist + off, min(sb, nflag - off), sb, rs, rq, idx1, idx2, scratch, vbe, tid, nullptr);
}

struct VqTcMem {
    unsigned char *a;        // A tiles, MTMAX x 16 KB (transient)
    char *small;             // transient: surv, nrm, marg, flag, flist, list copy  (vq_tc_small_bytes(maxn))
    unsigned char *part;     // transient: vq_tc_part_bytes(MTMAX)
    unsigned char *bring;    // B ring, NB x 8 KB (written by the producer thread only)
    char *scratch;           // the caller's scratch for the CUDA-core fallback search (may contain a / small / part)
    int scratch_bytes;
    int *tail;               // MAXN ints OUTSIDE scratch[0, scratch_bytes): the fallback's row list
};
__host__ __device__ constexpr int vq_tc_small_bytes(int maxn) { return maxn * (5 * 4 + 5 * 4 + 2 * 4 + 4 + 4 + 4) + 64; }

// Every one of the NT searching threads (256: two warps per TMEM lane quarter, each scanning one half of a unit's 64
// columns; 128: one warp per quarter scanning both halves) calls this with identical arguments and tid in [0, NT).
// `last` = this is the last search of the frame.
template <typename T, int MAXN, int NB, int NT = kComputeThreads, int UNITS>
__device__ FPC_VQ_TC_OUTLINE void vq_tc_search(const PackedVq &bk_in, const char *__restrict__ cbbase, const int *__restrict__ list, int n,
                                             const float *__restrict__ rs, float *__restrict__ rq, int *__restrict__ idx1, int *__restrict__ idx2,
                                             const VqTcMem &mem, VqTcShared<NB, UNITS> *sh, int last, int tid, long long *dbg)
{
    constexpr int MTMAX = (5 * MAXN + 127) / 128;
    const uint32_t tb = sh->tmem_base;
    VqTcCount cnt{sh->cnt_go, sh->cnt_b, sh->cnt_u, 0u};
    static_assert(MAXN <= 64 && MTMAX <= 3, "rows: at most 64 vectors x 5 survivors = 3 tiles");
    const int warp = tid >> 5, lane = tid & 31;
    static_assert(NT == 256 || NT == 128, "searching threads");
    constexpr int kHalves = NT / 128;           // warps per lane quarter
    constexpr int kLoads = 2 / kHalves;         // 32-column loads a warp makes per unit
    const int q = warp & 3, hh = warp >> 2;
    PackedVq bk;
    {
        const volatile long long *src = reinterpret_cast<const volatile long long *>(&bk_in);
        long long *dst = reinterpret_cast<long long *>(&bk);
#pragma unroll
        for (int i = 0; i < (int)(sizeof(PackedVq) / 8); ++i) dst[i] = src[i];
    }
    const bool two = bk.stages == 2;
    const int ns = two ? kSurv : 1;
    const int nchunks = bk.Kp64 >> 6;
    const float *betas = reinterpret_cast<const float *>(cbbase + bk.off_tcbeta);
    const float *cmax = reinterpret_cast<const float *>(cbbase + bk.off_cmax);
    int *surv = reinterpret_cast<int *>(mem.small);
    float *nrm = reinterpret_cast<float *>(surv + MAXN * 5);
    float *marg = nrm + MAXN * 5;
    int *flag = reinterpret_cast<int *>(marg + MAXN * 2);
    int *flist = flag + MAXN;
    int *cntw = flist + MAXN;
    const float inf = __int_as_float(0x7f800000);
    long long tq0 = dbg ? clock64() : 0;
#define FPC_VQT(i) do { if (dbg) { const long long t_ = clock64(); dbg[i] += t_ - tq0; tq0 = t_; } } while (0)

    // ---- per-vector constants: margin M = 2^-15 R, ||x||^2 ----
    if (tid < n) {
        const float *xr = rs + list[tid] * kLdR + 4;
        float n2 = 0.0f;
#pragma unroll
        for (int d = 0; d < kDim; ++d) n2 = __fmaf_ru(xr[d], xr[d], n2);
        const float csum = __fadd_ru(cmax[0], two ? cmax[1] : 0.0f);
        const float r = __fadd_ru(__fsqrt_ru(n2), csum);
        marg[2 * tid] = __fmul_ru(__fmul_ru(r, r), 1.52587890625e-5f);       // M = 2^-16 R = 256 u R
        marg[2 * tid + 1] = n2;
        flag[tid] = 0;
        reinterpret_cast<int *>(mem.part + 256 * 24 + MAXN * 32)[tid] = 8;      // stage 0: smallest rank of a list's sixth key
    }
    named_bar_sync(1, NT);
    FPC_VQT(2);

    if (two) {
        // ================= stage 0: the survivor set =================
        const float beta = betas[0];
        const int P2 = n <= 32 ? 32 : 64, rep = 128 / P2;
        if (tid < 128) {
            const int v = tid & (P2 - 1), rho = tid / P2;
            if (v < n) {
                const float *xr = rs + list[v] * kLdR + 4;
                float xs[kDim];
                float amax = 0.0f;
#pragma unroll
                for (int d = 0; d < kDim; ++d) { xs[d] = -2.0f * beta * xr[d]; amax = fmaxf(amax, fabsf(xs[d])); }
                const bool ok = amax <= tc::kMaxScaledX;                                  // NaN -> false
                if (!ok) {
#pragma unroll
                    for (int d = 0; d < kDim; ++d) xs[d] = 0.0f;
                    if (rho == 0) flag[v] = 1;
                }
                // offset beta^2 (||x||^2 + M) as (o / 1024) x 1024: every score of the row becomes beta^2 (d + M) > 0
                const float off = ok ? __fmul_rn(__fmul_rn(beta, beta), __fadd_ru(marg[2 * v + 1], marg[2 * v])) * 9.765625e-4f : 0.0f;
                __half o0, o1, o2;
                tc::split3(off, o0, o1, o2);
                const __half one = __float2half_rn(1.0f);
                tc::store_row<true>(mem.a, rho * P2 + v, xs, one, one, one, o0, o1, o2);
            }
        }
        vq_tc_publish<NB, NT, UNITS>(sh, cnt, nchunks, 1, smem_u32(mem.a), 2 * nchunks, bk.off_b[0] + (long long)(blockIdx.x % kWeightReplicas) * bk.b_rep_stride, smem_u32(mem.bring), 0, tid);
        // scan: quarter q belongs to replica (32 q) / P2 and takes the chunks c = replica (mod rep)
        const int my_rep = (32 * q) / P2;
        Top6 tk;
        tk.reset();
        for (int c = 0; c < nchunks; ++c, ++cnt.u) {
            const int ds = (int)(cnt.u % UNITS);
            const long long tw0 = dbg ? clock64() : 0;
            mbar_wait(&sh->d_full[ds], (cnt.u / UNITS) & 1u);
            if (dbg) dbg[c == 0 ? 8 : 9] += clock64() - tw0;

            umma::fence_after_sync();
            if (c % rep == my_rep && (32 * q) % P2 < n) {
#pragma unroll
                for (int ld = 0; ld < kLoads; ++ld) {
                    const int h2 = kLoads == 1 ? hh : ld;
                    uint32_t v[32];
                    tc::tmem_ld32(tb + ((uint32_t)(32 * q) << 16) + (uint32_t)(ds * 64 + 32 * h2), v);
                    tc::tmem_ld_wait(v);
                    if (ld == kLoads - 1) {
                        umma::fence_before_sync();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&sh->d_empty[ds]);
                    }
                    const unsigned cbase = (unsigned)(64 * c + 32 * h2);
#pragma unroll
                    for (int j = 0; j < 32; ++j) tk.insert(__uint_as_float((v[j] & 0xfffffc00u) | (cbase + j)));
                }
            } else {
                umma::fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(&sh->d_empty[ds]);
            }
        }
        {
            float *kp = reinterpret_cast<float *>(mem.part) + tid * 6;
            kp[0] = tk.t0; kp[1] = tk.t1; kp[2] = tk.t2; kp[3] = tk.t3; kp[4] = tk.t4; kp[5] = tk.t5;
        }
        named_bar_sync(1, NT);
        FPC_VQT(3);
        // ---- the eight smallest keys of every vector: thread (vector v, list l) ranks the six keys of list l among all
        //      2 rep sorted lists of the row (the keys are distinct: they carry the codeword index); l is uniform per warp ----
        unsigned *g8 = reinterpret_cast<unsigned *>(mem.part + 256 * 24);        // [n][8]
        int *exh = reinterpret_cast<int *>(g8 + MAXN * 8);                        // [n]
        {
            const int v = tid & (P2 - 1), l = tid / P2;                           // kHalves rep = NT / P2 lists
            if (v < n) {
                const unsigned *keys = reinterpret_cast<const unsigned *>(mem.part);
                if (rep == 2) vq_tc_rank_list<2 * kHalves, kHalves>(keys, P2, v, l, g8, exh);
                else vq_tc_rank_list<4 * kHalves, kHalves>(keys, P2, v, l, g8, exh);
            }
        }
        named_bar_sync(1, NT);
        if (tid < n) {
            const int v = tid;
            unsigned g[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) g[r] = g8[v * 8 + r];
            const int r6 = exh[v];
            const float M = __fmul_ru(marg[2 * v], __fmul_rn(beta, beta));      // the keys are in the scale beta^2
            const float v5 = __uint_as_float(g[4] & 0xfffffc00u);
            const float thr = __fadd_ru(__fmaf_ru(v5, 4.8828125e-4f, v5), M);
            int nc = 5;
#pragma unroll
            for (int r = 5; r < 8; ++r) nc += (__uint_as_float(g[r] & 0xfffffc00u) <= thr) ? 1 : 0;
            // the nc smallest keys are the candidates; they are all there unless one of them is a list's sixth key
            const bool exhausted = r6 < nc;
            bool ok = !exhausted && (__uint_as_float(g[5] & 0xfffffc00u) > thr);
            int ks[kSurv];
#pragma unroll
            for (int r = 0; r < kSurv; ++r) ks[r] = (int)(g[r] & 1023u);
            // near-tie around rank 5 with all candidates known: a warp ranks them exactly below (exh[v] = 2 + nc)
            const bool rerank = !ok && !exhausted && nc < 8;
            bool bad = false;
#pragma unroll
            for (int r = 0; r < kSurv; ++r) bad = bad || ks[r] >= bk.K;
#pragma unroll
            for (int r = 0; r < kSurv; ++r) surv[v * kSurv + r] = (ok && !bad) ? ks[r] : r;     // flagged rows keep harmless indices
            exh[v] = rerank ? 2 + nc : 0;
            if (!rerank && (!ok || bad)) flag[v] = 1;
        }
        named_bar_sync(1, NT);
        // ---- near-ties: lane r < nc evaluates candidate r with the reference's exact arithmetic (dist17<T>), then the five
        //      smallest (distance, index) pairs in order; rare, a warp per vector ----
#pragma unroll 1
        for (int v = warp; v < n; v += NT / 32) {
            const int st = exh[v];
            if (st < 2) continue;
            const int nc = st - 2;
            T d = Rn<T>::inf();
            int ki = 0x7fffffff;
            if (lane < nc) {
                const int kk = (int)(g8[v * 8 + lane] & 1023u);
                if (kk < bk.K) {
                    ki = kk;
                    const T *crow = reinterpret_cast<const T *>(cbbase + bk.off_r[0]) + (size_t)ki * kDim;
                    const float *xr = rs + list[v] * kLdR + 4;
                    T xv[kDim], cv[kDim];
#pragma unroll
                    for (int dd = 0; dd < kDim; ++dd) { xv[dd] = (T)xr[dd]; cv[dd] = crow[dd]; }
                    d = dist17<T>(xv, cv);
                }
            }
            bool good = true;
#pragma unroll
            for (int r = 0; r < kSurv; ++r) {
                T wd = d;
                int wi = ki;
                warp_argmin(wd, wi);
                good = good && wi != 0x7fffffff;
                if (lane == 0) surv[v * kSurv + r] = wi != 0x7fffffff ? wi : r;
                if (ki == wi) { d = Rn<T>::inf(); ki = 0x7fffffff; }
            }
            if (lane == 0 && !good) flag[v] = 1;
        }
        named_bar_sync(1, NT);
        FPC_VQT(4);
    }

    // ================= last stage: nearest entry jointly over the survivors =================
    {
        const int sc1 = two ? 1 : 0;
        const float beta = betas[sc1];
        const int nrows = n * ns;
        const int mtiles = (nrows + 127) >> 7;
        const T *cbr0 = reinterpret_cast<const T *>(cbbase + bk.off_r[0]);
        for (int r = tid; r < nrows; r += NT) {
            const int v = two ? r / kSurv : r, s = two ? r - v * kSurv : 0;
            const float *xr = rs + list[v] * kLdR + 4;
            float xs[kDim];
            float nn = 0.0f, amax = 0.0f;
            if (two) {
                const T *crow = cbr0 + (size_t)surv[v * kSurv + s] * kDim;
#pragma unroll
                for (int d = 0; d < kDim; ++d) xs[d] = __fsub_rn(xr[d], (float)crow[d]);
            } else {
#pragma unroll
                for (int d = 0; d < kDim; ++d) xs[d] = xr[d];
            }
#pragma unroll
            for (int d = 0; d < kDim; ++d) {
                nn = __fmaf_rn(xs[d], xs[d], nn);
                xs[d] = -2.0f * beta * xs[d];
                amax = fmaxf(amax, fabsf(xs[d]));
            }
            if (!(amax <= tc::kMaxScaledX)) {
#pragma unroll
                for (int d = 0; d < kDim; ++d) xs[d] = 0.0f;
                flag[v] = 1;
            }
            nrm[r] = nn;
            const __half one = __float2half_rn(1.0f);
            tc::store_row<true>(mem.a + (size_t)(r >> 7) * tc::kTileBytes, r & 127, xs, one, one, one);
        }
        vq_tc_publish<NB, NT, UNITS>(sh, cnt, nchunks, mtiles, smem_u32(mem.a), two ? 0 : nchunks,
                          bk.off_b[sc1] + (long long)(blockIdx.x % kWeightReplicas) * bk.b_rep_stride, smem_u32(mem.bring), last, tid);
        tc::Scan sc[MTMAX];
#pragma unroll
        for (int m = 0; m < MTMAX; ++m) sc[m].reset();
        for (int c = 0; c < nchunks; ++c) {
#pragma unroll
            for (int m = 0; m < MTMAX; ++m) {
                if (m < mtiles) {
                    const int ds = (int)(cnt.u % UNITS);
                    const long long tw0 = dbg ? clock64() : 0;
                    mbar_wait(&sh->d_full[ds], (cnt.u / UNITS) & 1u);
                    (void)tw0;
                    umma::fence_after_sync();
                    if (128 * m + 32 * q < nrows) {
#pragma unroll
                        for (int ld = 0; ld < kLoads; ++ld) {
                            uint32_t v[32];
                            tc::tmem_ld32(tb + ((uint32_t)(32 * q) << 16) + (uint32_t)(ds * 64 + 32 * (kLoads == 1 ? hh : ld)), v);
                            tc::tmem_ld_wait(v);
                            if (ld == kLoads - 1) {
                                umma::fence_before_sync();
                                __syncwarp();
                                if (lane == 0) mbar_arrive(&sh->d_empty[ds]);
                            }
                            sc[m].feed(v, kLoads == 1 ? 2 * c : 4 * c + 2 * ld);     // group ids: per half / over all 64 columns
                        }
                    } else {
                        umma::fence_before_sync();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&sh->d_empty[ds]);
                    }
                    ++cnt.u;
                }
            }
        }
        // the running counts go back to the control block BEFORE the barriers below: the next search (the other book of
        // the same frame) reads them on entry, with no barrier of its own in between
        if (tid == 0) { sh->cnt_go = cnt.go; sh->cnt_u = cnt.u; }
        {
            float4 *pp = reinterpret_cast<float4 *>(mem.part);
#pragma unroll
            for (int m = 0; m < MTMAX; ++m) {
                if (m < mtiles) {
                    float best, second;
                    int jb;
                    sc[m].finish(best, second, jb);
                    if (kLoads == 1) {
                        const int col = 64 * (sc[m].ga >> 1) + 32 * hh + 16 * (sc[m].ga & 1) + jb;
                        pp[(m * 128 + 32 * q + lane) * 2 + hh] = make_float4(best, second, __int_as_float(col), 0.0f);
                    } else {
                        pp[(m * 128 + 32 * q + lane) * 2] = make_float4(best, second, __int_as_float(16 * sc[m].ga + jb), 0.0f);
                        pp[(m * 128 + 32 * q + lane) * 2 + 1] = make_float4(inf, inf, __int_as_float(0), 0.0f);
                    }
                }
            }
        }
        named_bar_sync(1, NT);
        FPC_VQT(5);
        if (tid < n && !flag[tid]) {
            const int v = tid;
            const float4 *pp = reinterpret_cast<const float4 *>(mem.part);
            const float inv = __fdiv_rn(1.0f, __fmul_rn(beta, beta));            // power of two: exact
            // per survivor row: smallest distance (codeword known) and runner-up (codeword unknown)
            float db[kSurv], dsec[kSurv];
            int dj[kSurv];
            float b = inf;
#pragma unroll
            for (int s = 0; s < kSurv; ++s) {
                db[s] = inf; dsec[s] = inf; dj[s] = 0;
                if (s < ns) {
                    const int r = v * ns + s;
                    const float4 p0 = pp[r * 2], p1 = pp[r * 2 + 1];
                    float rb = p0.x;
                    const float rsec = fminf(fminf(p0.y, p1.y), fmaxf(p0.x, p1.x));
                    int rj = __float_as_int(p0.z);
                    if (p1.x < rb) { rb = p1.x; rj = __float_as_int(p1.z); }
                    const float nn = nrm[r];
                    db[s] = __fmaf_rn(rb, inv, nn);
                    dsec[s] = __fmaf_rn(rsec, inv, nn);
                    dj[s] = rj;
                    b = fminf(b, db[s]);
                }
            }
            // Margin of the last stage.  Its scores involve only x' = x - c0[k_s] and the last codebook, so their errors scale
            // with R' = (max_s ||x'_s|| + Cmax_last)^2, not with R: M' = 2^-16 R'.  For a two-stage book one more term: x' is
            // rounded to fp32 here while the reference keeps x - c0 in the codebook's dtype -- <= u (||x|| + Cmax0) per
            // component, i.e. <= 2 sqrt(17) u sqrt(R' R) on a distance -- covered by 2^-20 sqrt(R' R) = 16 u sqrt(R' R).
            float nmax = 0.0f;
#pragma unroll
            for (int s = 0; s < kSurv; ++s)
                if (s < ns) nmax = fmaxf(nmax, nrm[v * ns + s]);
            const float rp = __fadd_ru(__fsqrt_ru(__fmul_ru(nmax, 1.000002f)), cmax[sc1]);
            const float Rp = __fmul_ru(rp, rp);
            float M1 = __fmul_ru(Rp, 1.52587890625e-5f);
            if (two) M1 = __fadd_ru(M1, __fmul_ru(__fsqrt_ru(__fmul_ru(Rp, __fmul_ru(marg[2 * v], 65536.0f))), 9.5367431640625e-7f));
            const float thr = __fadd_ru(__fadd_ru(b, fabsf(b) * 9.5367431640625e-7f), M1);
            // candidates that may be the exact winner: everything at or below thr.  A runner-up below thr is a candidate
            // whose codeword the scan did not record -> the exact search; otherwise the candidates are the rows' winners
            int ncand = 0, ws = 0;
            bool unknown = false, bad = false;
#pragma unroll
            for (int s = 0; s < kSurv; ++s) {
                if (s < ns) {
                    unknown = unknown || !(dsec[s] > thr);
                    if (db[s] <= thr) { ++ncand; ws = s; bad = bad || dj[s] >= bk.K; }
                }
            }
            bool decided = !unknown && !bad && ncand >= 1;
            if (decided && ncand > 1) {
                // several known candidates inside the margin: the reference's arithmetic decides (vq_func.py:103-106,18):
                // diff = x - (0 + CB0[k_s]);  dist = sum((diff - CB1[j])^2)   in the codebook's dtype
                const float *xr = rs + list[v] * kLdR + 4;
                const T *cbr1 = reinterpret_cast<const T *>(cbbase + bk.off_r[1]);
                T bd = Rn<T>::inf();
                bool tie = false;
#pragma unroll
                for (int s = 0; s < kSurv; ++s) {
                    if (s < ns && db[s] <= thr) {
                        const T *c0 = cbr0 + (size_t)surv[v * kSurv + s] * kDim;
                        const T *c1 = cbr1 + (size_t)dj[s] * kDim;
                        T xd[kDim], cd[kDim];
#pragma unroll
                        for (int d = 0; d < kDim; ++d) {
                            xd[d] = Rn<T>::sub((T)xr[d], Rn<T>::add((T)0, c0[d]));
                            cd[d] = c1[d];
                        }
                        const T dd = dist17<T>(xd, cd);
                        if (dd < bd) { bd = dd; ws = s; tie = false; }
                        else if (dd == bd) tie = true;          // equal distances across survivor ranks: the exact search knows the ranks
                    }
                }
                decided = !tie;
            }
            if (decided) {
                const int row = list[v];
                if (two) { idx1[row] = surv[v * kSurv + ws]; idx2[row] = dj[ws]; }
                else { idx1[row] = dj[ws]; idx2[row] = -1; }
            } else {
                flag[v] = 1;
            }
        }
        named_bar_sync(1, NT);
    }

    // ---- quantised vectors of the decided rows: csum = 0; csum += CB[i][index[i,0]]  (vq_func.py:127-129) ----
    {
        const T *cbr0 = reinterpret_cast<const T *>(cbbase + bk.off_r[0]);
        const T *cbr1 = reinterpret_cast<const T *>(cbbase + bk.off_r[1]);
        for (int e = tid; e < n * kDim; e += NT) {
            const int v = e / kDim, d = e - v * kDim;
            if (!flag[v]) {
                const int row = list[v];
                T csum = Rn<T>::add((T)0, cbr0[(size_t)idx1[row] * kDim + d]);
                if (two) csum = Rn<T>::add(csum, cbr1[(size_t)idx2[row] * kDim + d]);
                rq[row * 20 + 1 + d] = (float)csum;
            }
        }
        if (warp == 0) {
            int nf2 = 0;
            for (int base = 0; base < n; base += 32) {
                const int v = base + lane;
                const bool f = v < n && flag[v] != 0;
                const unsigned bf = __ballot_sync(0xffffffffu, f);
                if (f) flist[nf2 + __popc(bf & ((1u << lane) - 1u))] = list[v];
                nf2 += __popc(bf);
            }
            if (lane == 0) cntw[0] = nf2;
        }
        named_bar_sync(1, NT);
    }
    const int nflag = cntw[0];
    FPC_VQT(6);
    if (dbg) { dbg[0] += n; dbg[1] += nflag; }
    if (nflag > 0) {
        // the exact CUDA-core search for the undecided rows; it reuses the scratch, so the row list moves out of it
        const int mine = tid < nflag ? flist[tid] : 0;
        named_bar_sync(1, NT);
        int *tail = mem.tail;
        if (tid < nflag) tail[tid] = mine;
        named_bar_sync(1, NT);
        vq_tc_fallback<T, NT>(bk, cbbase, tail, nflag, MAXN, rs, rq, idx1, idx2, mem.scratch, mem.scratch_bytes, tid);
        FPC_VQT(7);
    }
#undef FPC_VQT
}

// One VQ search of the fused kernels (one book, the rows of `list`): tensor-core screen when the book has at least
// 64 entries, the CUDA-core search otherwise.  Returns true if a phase was published (the caller owes the other roles
// one publication with last = 1 per frame).
template <int MAXN, int NB, int NT = kComputeThreads, int UNITS>
__device__ __forceinline__ bool vq_tc_dispatch(const PackedVq &bk, const char *cbbase, const int *list, int n, const float *rs, float *rq,
                                               int *idx1, int *idx2, const VqTcMem &mem, VqTcShared<NB, UNITS> *sh, int last, int tid, long long *dbg)
{
    if (bk.K >= 64) {
        if (bk.dtype == FPC_F32) vq_tc_search<float, MAXN, NB, NT, UNITS>(bk, cbbase, list, n, rs, rq, idx1, idx2, mem, sh, last, tid, dbg);
        else vq_tc_search<double, MAXN, NB, NT, UNITS>(bk, cbbase, list, n, rs, rq, idx1, idx2, mem, sh, last, tid, dbg);
        return true;
    }
    if (bk.dtype == FPC_F32) vq_tc_fallback<float, NT>(bk, cbbase, list, n, MAXN, rs, rq, idx1, idx2, mem.scratch, mem.scratch_bytes, tid);
    else vq_tc_fallback<double, NT>(bk, cbbase, list, n, MAXN, rs, rq, idx1, idx2, mem.scratch, mem.scratch_bytes, tid);
    return false;
}

}  // namespace fpc
