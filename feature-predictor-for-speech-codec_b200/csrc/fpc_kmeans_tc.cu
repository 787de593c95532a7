// fpc_kmeans_tc.cu -- the assignment step of the Lloyd iteration with the distance screen on the tensor cores.
//
// Reference: /root/reference/src/quantization/cb_func.py:56-68 (find_nearest: float64 direct-form distances, first
// minimum) and :82-86 (per-centroid sums and counts).  Same contract and same results as kmeans_assign_kernel
// (fpc_kmeans.cu): the index of every vector is the one numpy's float64 evaluation picks.  What changes is where the
// N x K x 17 contraction runs: here it is a tcgen05 GEMM (vectors x centroids, fp16-pair operands, fp32 accumulators in
// TMEM, fpc_tc.cuh), the CUDA cores only scan the accumulators for the smallest score and the runner-up of every
// vector (one min instruction per score) and decide the vectors whose runner-up is more than the slack away; the
// few others are re-done by a warp with the fp32 screen + float64 evaluation the CUDA-core kernel uses.
//
// One CTA per SM, 12 warps:
//   warps 0..7   scan.  Warp w reads TMEM lanes 32 (w % 4) .. +31 (one vector per lane) and the column half w / 4
//                of every 128-centroid chunk; then merge, decision, fallback, float64 atomics for the sums.
//   warps 8..10  load.  Thread r converts vectors r (and r + 96) of the next 128-vector tile (brought on chip by one
//                bulk copy) into fp16-pair A operand rows (scaled by -2 beta) in a ring of 16 KB tiles.
//   warp 11      lane 0 issues the MMAs: per tile and 128-centroid chunk four K = 16 slabs into one of four
//                128-column TMEM slots; tcgen05.commit releases the A tile and publishes the slot.  Lane 1 issues
//                the bulk copies of the raw tiles.
// The B operand (the whole codebook, <= 1024 x 128 B) is resident in shared memory: one bulk copy per launch from
// the image kmeans_tc_pack_kernel writes (also: the fp32 shadow and the scale).
#include "fpc_common.cuh"
#include "fpc_vq.cuh"
#include "fpc_tc.cuh"

namespace fpc {

// Scanner warps: 8 (two per scheduler).  The scan is bound by the min-instruction rate of the ALU pipe, not by latency:
// 16 warps (measured) were 25 % slower -- same scan time, twice the merge / decision overhead and register spills.
constexpr int kTcScanWarps = 8;
constexpr int kTcColParts = kTcScanWarps / 4;                  // column parts of a 128-centroid chunk (one per warp of a lane quarter)
constexpr int kTcLdPerUnit = 128 / (32 * kTcColParts);         // 32-column TMEM loads per thread and chunk
// roles: kTcScanWarps scanning warps, two converting ("loader") warps, one warp whose lane 0 issues the bulk copies, one
// warp that issues the MMAs
constexpr int kTcScan = 32 * kTcScanWarps, kTcLoad = 64, kTcThreads = kTcScan + kTcLoad + 64;   // a multiple of 128 threads
constexpr int kTcASlots = 2, kTcDSlots = 4, kTcRawSlots = 4;
constexpr int kTcRawFloats = 128 * kDim;           // one tile of the data set: 8704 contiguous bytes
constexpr int kTcShadowLd = 20;                    // fp32 shadow row: c0..c16, -, ||c||^2, -
// Slack of a decision, both screens: 2^-17 (||x|| + Cmax)^2 = 128 u R.  A comparison involves two scores: the
// tensor-core scores are within 10 u R (operand format, fpc_tc.cuh) plus the accumulation error, measured at <= 3 u R
// in total over the scales of tests/test_gpu_quant_kmeans.py::test_tc_screen_score_error_budget (asserted <= 64 u R);
// the fp32 screen of the fallback is within 25 u R (21 u R as in fpc_kmeans.cu + the 22-bit shadow it reads here).
constexpr float kTcSlackRel = 7.62939453125e-6f;
constexpr float kF32SlackRel = 7.62939453125e-6f;

struct KmeansTcMeta { float beta, cmax; int K, Kp; unsigned int n_fallback, n_rows; unsigned long long prof[8]; };
__host__ __device__ constexpr size_t tc_image_bytes(int Kp) { return (size_t)Kp * tc::kK * 2; }
__host__ __device__ constexpr size_t tc_pack_bytes(int K)
{
    const int Kp = (K + 127) / 128 * 128;
    return 256 + tc_image_bytes(Kp) + (size_t)K * kTcShadowLd * 4;
}

// ------------------------------------------------------------------------------------------
// codebook -> B operand image [Kp / 128 tiles][16 KB], fp32 shadow [K][20], meta
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024, 1)
kmeans_tc_pack_kernel(const double *__restrict__ cb, int K, unsigned char *__restrict__ out)
{
    __shared__ float s_red[32], s_red2[32];
    const int Kp = (K + 127) / 128 * 128;
    KmeansTcMeta *meta = reinterpret_cast<KmeansTcMeta *>(out);
    unsigned char *image = out + 256;
    float *shadow = reinterpret_cast<float *>(out + 256 + tc_image_bytes(Kp));
    float amax = 0.0f, n2max = 0.0f;
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        float a = 0.0f;
#pragma unroll
        for (int d = 0; d < kDim; ++d) {
            const float v = __double2float_ru(fabs(cb[(size_t)k * kDim + d]));
            amax = fmaxf(amax, v);
            a = __fmaf_ru(v, v, a);
        }
        n2max = fmaxf(n2max, a);
    }
    for (int off = 16; off > 0; off >>= 1) {
        amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, off));
        n2max = fmaxf(n2max, __shfl_xor_sync(0xffffffffu, n2max, off));
    }
    if ((threadIdx.x & 31) == 0) { s_red[threadIdx.x >> 5] = amax; s_red2[threadIdx.x >> 5] = n2max; }
    __syncthreads();
    amax = 0.0f; n2max = 0.0f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { amax = fmaxf(amax, s_red[w]); n2max = fmaxf(n2max, s_red2[w]); }
    const float beta = tc::scale_for(amax);
    if (threadIdx.x == 0) { meta->beta = beta; meta->cmax = __fsqrt_ru(n2max); meta->K = K; meta->Kp = Kp; meta->n_fallback = 0; meta->n_rows = 0; for (int i = 0; i < 8; ++i) meta->prof[i] = 0; }
    for (int k = threadIdx.x; k < Kp; k += blockDim.x) {
        unsigned char *tile = image + (size_t)(k >> 7) * tc::kTileBytes;
        float xs[kDim];
        __half n0, n1, n2;
        if (k < K) {
            double nn = 0.0;
#pragma unroll
            for (int d = 0; d < kDim; ++d) {
                const double c = cb[(size_t)k * kDim + d];
                xs[d] = (float)(c * (double)beta);                // beta is a power of two: only the fp32 rounding of c
                shadow[(size_t)k * kTcShadowLd + d] = (float)c;
                nn += c * c;
            }
            shadow[(size_t)k * kTcShadowLd + 17] = 0.0f;
            shadow[(size_t)k * kTcShadowLd + 18] = (float)nn;
            shadow[(size_t)k * kTcShadowLd + 19] = 0.0f;
            tc::split3((float)(nn * (double)beta * (double)beta), n0, n1, n2);
        } else {
#pragma unroll
            for (int d = 0; d < kDim; ++d) xs[d] = 0.0f;
            n0 = n1 = n2 = __float2half_rn(tc::kPadNorm);
        }
        tc::store_row<false>(tile, k & 127, xs, n0, n1, n2);
    }
}

// fp32 screen value of one (vector, centroid) pair, the chain of fpc_kmeans.cu (screen17): ||c||^2 + sum (-2 x_d) c_d
__device__ __forceinline__ float tc_screen17(const float (&x)[kDim], const float *__restrict__ row)
{
    float s = row[18];
#pragma unroll
    for (int d = 0; d < kDim; ++d) s = __fmaf_rn(-2.0f * x[d], row[d], s);
    return s;
}

// One undecided vector, by a whole warp: fp32 screen of this lane's centroids k = lane + 32 j from the RESIDENT operand
// image (hi + lo of an fp16 pair is exact in fp32, so the image doubles as a 22-bit shadow of beta c; the norm is
// n0 + n1 + n2), then float64 direct-form distances of the candidates inside the fp32 slack, ascending k with strict <
// (numpy's first minimum, cb_func.py:66).  Everything in the scale of the operands (x beta^2), which the bounds do not see.
__device__ __noinline__ int tc_exact_assign(const float *__restrict__ xrow, float rr2, float beta, int K,
                                            const unsigned char *__restrict__ sB, const double *__restrict__ cb, int lane)
{
    const float inf = __int_as_float(0x7f800000);
    float xu[kDim], xs[kDim];
#pragma unroll
    for (int d = 0; d < kDim; ++d) { xu[d] = __ldg(xrow + d); xs[d] = -2.0f * beta * xu[d]; }
    const float slack32 = __fmul_ru(__fadd_ru(__fmul_ru(rr2, kF32SlackRel), 1e-30f), beta * beta);
    float sv[32];
    float m = inf;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        const int k = lane + 32 * j;
        sv[j] = inf;
        if (k < K) {
            const unsigned char *rowp = sB + (size_t)(k >> 7) * tc::kTileBytes + (size_t)(k & 127) * 16;
            uint32_t wds[28];
#pragma unroll
            for (int c8 = 0; c8 < 7; ++c8) {
                const uint4 t4 = *reinterpret_cast<const uint4 *>(rowp + (size_t)c8 * 2048);
                wds[4 * c8] = t4.x; wds[4 * c8 + 1] = t4.y; wds[4 * c8 + 2] = t4.z; wds[4 * c8 + 3] = t4.w;
            }
            auto h = [&](int e) { return __half2float(__ushort_as_half((unsigned short)(wds[e >> 1] >> ((e & 1) * 16)))); };
            float sacc = __fadd_rn(__fadd_rn(h(51), h(52)), h(53));
#pragma unroll
            for (int d = 0; d < kDim; ++d) sacc = __fmaf_rn(xs[d], __fadd_rn(h(3 * d), h(3 * d + 1)), sacc);
            sv[j] = sacc;
            m = fminf(m, sacc);
        }
    }
    for (int off = 16; off > 0; off >>= 1) m = fminf(m, __shfl_xor_sync(0xffffffffu, m, off));
    const float thr = __fadd_ru(m, slack32);
    double bd = Rn<double>::inf();
    int bk = 0x7fffffff;
#pragma unroll 1
    for (int j = 0; j < 32; ++j) {
        float sj = inf;
#pragma unroll
        for (int t = 0; t < 32; ++t) sj = t == j ? sv[t] : sj;
        if (sj <= thr) {
            const int k = lane + 32 * j;
            double xd[kDim], cd[kDim];
#pragma unroll
            for (int d = 0; d < kDim; ++d) { xd[d] = (double)xu[d]; cd[d] = cb[(size_t)k * kDim + d]; }
            const double dd = dist17<double>(xd, cd);
            if (dd < bd || bk == 0x7fffffff) { bd = dd; bk = k; }      // ascending k per lane: first minimum kept
        }
    }
    warp_argmin(bd, bk);
    return bk == 0x7fffffff ? 0 : bk;     // NaN input: no candidate; numpy's argmin of all-NaN distances is 0
}

__global__ void __launch_bounds__(kTcThreads, 1)
kmeans_assign_tc_kernel(const float *__restrict__ data, long N, const double *__restrict__ cb, int K,
                        const unsigned char *__restrict__ packed, double *__restrict__ sums, double *__restrict__ counts,
                        int32_t *__restrict__ idx_out, int R, long ntiles)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    const KmeansTcMeta *meta = reinterpret_cast<const KmeansTcMeta *>(packed);
    const int Kp = (K + 127) / 128 * 128;
    const int nchunks = Kp >> 7;
    unsigned char *sB = smem;
    unsigned char *sA = smem + tc_image_bytes(Kp);
    float *raw = reinterpret_cast<float *>(sA + kTcASlots * tc::kTileBytes);       // [kTcRawSlots][128][17] fp32, as in HBM
    float *pm = raw + kTcRawSlots * kTcRawFloats;                                   // [3][128][4]: best, second, column of column quarters 1..3
    int *dec = reinterpret_cast<int *>(pm + 3 * 128 * 4);                               // [2][128] decided centroid (or -1), two tiles
    uint64_t *bars = reinterpret_cast<uint64_t *>(dec + 2 * 128);
    uint64_t *b_full = bars, *a_full = bars + 1, *a_empty = a_full + kTcASlots, *d_full = a_empty + kTcASlots,
             *d_empty = d_full + kTcDSlots, *raw_full = d_empty + kTcDSlots, *raw_empty = raw_full + kTcRawSlots,
             *dec_full = raw_empty + kTcRawSlots, *dec_empty = dec_full + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(dec_empty + 2);

    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;      // (warp index known to be uniform)
    const long my_tiles = ntiles > (long)blockIdx.x ? (ntiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;

    if (tid == 0) {
        mbar_init(b_full, 1);
        for (int s = 0; s < kTcASlots; ++s) { mbar_init(&a_full[s], kTcLoad / 32); mbar_init(&a_empty[s], 1); }
        for (int s = 0; s < kTcDSlots; ++s) { mbar_init(&d_full[s], 1); mbar_init(&d_empty[s], kTcScan / 32); }
        for (int s = 0; s < kTcRawSlots; ++s) { mbar_init(&raw_full[s], 1); mbar_init(&raw_empty[s], (kTcScan + kTcLoad) / 32); }
        for (int s = 0; s < 2; ++s) { mbar_init(&dec_full[s], kTcScan / 32); mbar_init(&dec_empty[s], kTcLoad / 32); }
        mbar_fence_init();
    }
    if (warp == 0) umma::tmem_alloc(tmem_slot, 512);
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tb = *tmem_slot;
    const float beta = meta->beta, cmax = meta->cmax;
    // a tile of 128 vectors is 8704 contiguous bytes of the data set: one bulk copy brings it on chip, where the loaders
    // (operand conversion) and the scanners (norms, sums) read their rows -- a per-thread row load from global memory
    // is a 68-byte-strided access that costs the L1 seventeen wavefronts per instruction.  The last, partial tile of
    // the data set is read from global memory instead (its size need not be a multiple of 16 bytes).
    auto tile_index = [&](long i) { return (long)blockIdx.x + i * gridDim.x; };
    auto tile_is_bulk = [&](long t) { return (t + 1) * 128 <= N; };

    if (warp == kTcThreads / 32 - 1) {
        // ---------------- MMA issuer: the whole warp runs the loop converged, one elected lane issues ----------------
        // (Issued from an `if (lane == 0)` region every tcgen05.mma goes through ELECT + R2UR sequences, 100-200 cycles
        //  each -- more than the tensor pipe needs for the MMA; in a warp-uniform loop the descriptors stay in uniform
        //  registers.  Nothing lane-dependent may sit in this warp's loop: the copies are issued by the warp before it.)
        mbar_wait(b_full, 0);            // the codebook image
        const uint32_t idesc = tc::instr_desc_f16(128, 128);
        const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
        uint32_t u = 0;
        for (long i = 0; i < my_tiles; ++i) {
            const int sa = (int)(i % kTcASlots);
            const uint32_t na = (uint32_t)(i / kTcASlots);
            mbar_wait(&a_full[sa], na & 1u);
            umma::fence_after_sync();
            for (int c = 0; c < nchunks; ++c, ++u) {
                const int ds = (int)(u % kTcDSlots);
                const uint32_t nd = u / kTcDSlots;
                if (nd > 0) { mbar_wait(&d_empty[ds], (nd - 1) & 1u); umma::fence_after_sync(); }
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                    umma::mma_bf16_elect(tb + (uint32_t)ds * 128u, umma::smem_desc(a0 + sa * tc::kTileBytes + ks * tc::kSlabBytes, 128),
                                         umma::smem_desc(b0 + c * tc::kTileBytes + ks * tc::kSlabBytes, 128), idesc, ks > 0);
                umma::commit_elect(&d_full[ds]);
            }
            umma::commit_elect(&a_empty[sa]);
        }
    } else if (warp == kTcThreads / 32 - 2) {
        if (lane == 0) {
            // ---------------- bulk copies: the codebook image once, then the raw tiles, kTcRawSlots tiles ahead ----------------
            const uint32_t bytes = (uint32_t)tc_image_bytes(Kp);
            mbar_arrive_expect_tx(b_full, bytes);
            for (uint32_t off = 0; off < bytes; off += tc::kTileBytes)
                bulk_g2s(sB + off, packed + 256 + off, tc::kTileBytes, b_full);
            for (long i = 0; i < my_tiles; ++i) {
                const int sr = (int)(i % kTcRawSlots);
                const uint32_t nr = (uint32_t)(i / kTcRawSlots);
                if (nr > 0) mbar_wait(&raw_empty[sr], (nr - 1) & 1u);
                const long t = tile_index(i);
                if (tile_is_bulk(t)) {
                    mbar_arrive_expect_tx(&raw_full[sr], kTcRawFloats * 4);
                    bulk_g2s(raw + sr * kTcRawFloats, data + t * kTcRawFloats, kTcRawFloats * 4, &raw_full[sr]);
                } else {
                    mbar_arrive(&raw_full[sr]);
                }
            }
        }
    } else if (warp >= kTcScan / 32) {
        // ---------------- loaders: vectors -> fp16-pair A tiles ----------------
        const int r0 = tid - kTcScan;
        const __half one = __float2half_rn(1.0f);
        // The loaders also add the vectors of a DECIDED tile to the sums (cb_func.py:82-86), one tile behind their
        // conversions: 2 304 float64 atomics per tile, which the L2 takes at about one per cycle and SM -- done by the
        // scanning warps they were 27 % of a tile; here they run beside the scan of the next tile.
        auto accumulate = [&](long i) {
            const int sr = (int)(i % kTcRawSlots), sd = (int)(i & 1);
            const long t = tile_index(i);
            mbar_wait(&dec_full[sd], (uint32_t)(i >> 1) & 1u);
            for (int r = r0; r < 128; r += kTcLoad) {
                const int b = dec[sd * 128 + r];
                if (b >= 0) {
                    float x[kDim];
                    if (tile_is_bulk(t)) {
#pragma unroll
                        for (int d = 0; d < kDim; ++d) x[d] = raw[sr * kTcRawFloats + r * kDim + d];
                    } else {
#pragma unroll
                        for (int d = 0; d < kDim; ++d) x[d] = __ldg(data + (t * 128 + r) * kDim + d);
                    }
                    double *s2 = sums, *c2 = counts;
                    if (R > 1) {
                        const int rep = (int)(((long)blockIdx.x * 128 + r) % R);
                        s2 += (size_t)rep * K * kDim;
                        c2 += (size_t)rep * K;
                    }
                    s2 += (size_t)b * kDim;
#pragma unroll
                    for (int d = 0; d < kDim; ++d) atomicAdd(s2 + d, (double)x[d]);
                    atomicAdd(c2 + b, 1.0);
                }
            }
            __syncwarp();
            if (lane == 0) { mbar_arrive(&dec_empty[sd]); mbar_arrive(&raw_empty[sr]); }
        };
        for (long i = 0; i < my_tiles; ++i) {
            const int sa = (int)(i % kTcASlots), sr = (int)(i % kTcRawSlots);
            const uint32_t na = (uint32_t)(i / kTcASlots);
            const long t = tile_index(i);
            mbar_wait(&raw_full[sr], (uint32_t)(i / kTcRawSlots) & 1u);
            if (na > 0) mbar_wait(&a_empty[sa], (na - 1) & 1u);
#pragma unroll 1
            for (int r = r0; r < 128; r += kTcLoad) {
                float xs[kDim];
                if (tile_is_bulk(t)) {
#pragma unroll
                    for (int d = 0; d < kDim; ++d) xs[d] = raw[sr * kTcRawFloats + r * kDim + d];
                } else {
                    const long row = t * 128 + r;
#pragma unroll
                    for (int d = 0; d < kDim; ++d) xs[d] = row < N ? __ldg(data + row * kDim + d) : 0.0f;
                }
#pragma unroll
                for (int d = 0; d < kDim; ++d) {
                    float v = -2.0f * beta * xs[d];                       // exact (powers of two)
                    if (!(fabsf(v) <= tc::kMaxScaledX)) v = 0.0f;         // out of range (or NaN): the scan side sends the row to the exact path
                    xs[d] = v;
                }
                tc::store_row<true>(sA + sa * tc::kTileBytes, r, xs, one, one, one);
            }
            umma::fence_async_smem();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&a_full[sa]);
                if (!sums) mbar_arrive(&raw_empty[sr]);          // (with sums the slot is released after its accumulation)
            }
            if (sums && i > 0) accumulate(i - 1);
        }
        if (sums && my_tiles > 0) accumulate(my_tiles - 1);
    } else {
        // ---------------- scanners ----------------
        const int q = warp & 3, cq = warp >> 2;
        const int r = 32 * q + lane;
        uint32_t u = 0;
#ifdef FPC_KMEANS_TC_PROF
        long long pf[8] = {0, 0, 0, 0, 0, 0, 0, 0}, pt = clock64();
#define TCP(i) do { const long long t_ = clock64(); pf[i] += t_ - pt; pt = t_; } while (0)
#else
#define TCP(i) do { } while (0)
#endif
        for (long i = 0; i < my_tiles; ++i) {
            const long t = tile_index(i);
            const long row = t * 128 + r;
            const bool valid = row < N;
            const int sr = (int)(i % kTcRawSlots);
            tc::Scan sc;
            sc.reset();
            for (int c = 0; c < nchunks; ++c, ++u) {
                const int ds = (int)(u % kTcDSlots);
                TCP(0);
                mbar_wait(&d_full[ds], (u / kTcDSlots) & 1u);
                TCP(1);
                umma::fence_after_sync();
                const uint32_t ta = tb + ((uint32_t)(32 * q) << 16) + (uint32_t)(ds * 128 + 32 * kTcLdPerUnit * cq);
                if (kTcLdPerUnit == 1) {
                    uint32_t v[32];
                    tc::tmem_ld32(ta, v);
                    tc::tmem_ld_wait(v);
                    umma::fence_before_sync();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&d_empty[ds]);
                    sc.feed(v, 2 * c);
                } else {
                    uint32_t v0[32], v1[32];
                    tc::tmem_ld32(ta, v0);
                    tc::tmem_ld32(ta + 32, v1);
                    tc::tmem_ld_wait(v0);
                    tc::tmem_ld_wait(v1);
                    umma::fence_before_sync();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&d_empty[ds]);
                    sc.feed(v0, 4 * c);
                    sc.feed(v1, 4 * c + 2);
                }
            }
            TCP(0);
            float best, second;
            int jb;
            sc.finish(best, second, jb);
            // group id -> first column: 2 (or 4) groups of 16 per chunk and column part
            int col = kTcLdPerUnit == 1 ? 128 * (sc.ga >> 1) + 32 * cq + 16 * (sc.ga & 1) + jb
                                        : 128 * (sc.ga >> 2) + 64 * cq + 16 * (sc.ga & 3) + jb;
            if (cq > 0) { float *p = pm + ((cq - 1) * 128 + r) * 4; p[0] = best; p[1] = second; p[2] = __int_as_float(col); }
            // this thread's vector (the raw tile has long arrived: the MMAs above consumed its conversion)
            float x[kDim];
            mbar_wait(&raw_full[sr], (uint32_t)(i / kTcRawSlots) & 1u);
            if (tile_is_bulk(t)) {
#pragma unroll
                for (int d = 0; d < kDim; ++d) x[d] = raw[sr * kTcRawFloats + r * kDim + d];
            } else {
#pragma unroll
                for (int d = 0; d < kDim; ++d) x[d] = valid ? __ldg(data + row * kDim + d) : 0.0f;
            }
            // the slot is refilled once every warp has arrived: the loads above must have returned, not just been issued
            // (see gemm_part in fpc_encode_fp32.cu) -- the barrier address depends on the loaded registers
            uint32_t dep = 0u;
#pragma unroll
            for (int d = 0; d < kDim; ++d) dep |= __float_as_uint(x[d]);
            dep &= (uint32_t)((unsigned long long)N >> 63);       // zero at run time only (a constant zero is folded away)
            __syncwarp();
            if (lane == 0) mbar_arrive(reinterpret_cast<uint64_t *>(reinterpret_cast<char *>(&raw_empty[sr]) + dep));
            TCP(2);
            named_bar_sync(1, kTcScan);
            TCP(3);
            if (cq == 0) {
#pragma unroll
                for (int o = 0; o < kTcColParts - 1; ++o) {
                    const float *p = pm + (o * 128 + r) * 4;
                    const float ob = p[0], os = p[1];
                    const int oc = __float_as_int(p[2]);
                    second = fminf(fminf(second, os), fmaxf(best, ob));
                    if (ob < best) { best = ob; col = oc; }
                }
                // per-vector constants of the decision: R = (||x|| + Cmax)^2 <= 2 (||x||^2 + Cmax^2), rounded up
                float nx = 0.0f;
#pragma unroll
                for (int d = 0; d < kDim; ++d) nx = __fmaf_ru(x[d], x[d], nx);
                const float rr2 = __fmul_ru(2.0f, __fadd_ru(nx, __fmul_ru(cmax, cmax)));
                const float slack_tc = __fmul_ru(__fmul_ru(rr2, kTcSlackRel), beta * beta);   // scores are scaled by beta^2
                // every |(-2 beta x)_d| <= 2 beta ||x||: inside the operand range if 4 beta^2 ||x||^2 <= 1024^2 (NaN -> false)
                const bool in_range = __fmul_ru(4.0f * beta * beta, nx) <= tc::kMaxScaledX * tc::kMaxScaledX;
                bool decided = in_range && col < K && (second > __fadd_ru(best, slack_tc));
                int b = decided ? col : -1;
                // the undecided vectors (near-ties, duplicate centroids, out-of-range rows): one at a time by the whole warp
                unsigned todo = __ballot_sync(0xffffffffu, valid && !decided);
                if (todo && lane == 0) atomicAdd(const_cast<unsigned int *>(&meta->n_fallback), (unsigned)__popc(todo));
                while (todo) {
                    const int src = __ffs(todo) - 1;
                    todo &= todo - 1;
                    const int bk = tc_exact_assign(data + (row - lane + src) * kDim, __shfl_sync(0xffffffffu, rr2, src), beta, K, sB, cb, lane);
                    if (lane == src) b = bk;
                }
                if (sums) {
                    if (i >= 2) mbar_wait(&dec_empty[i & 1], (uint32_t)((i >> 1) - 1) & 1u);     // the loaders are done with tile i - 2
                    dec[(i & 1) * 128 + r] = valid ? b : -1;
                }
                if (valid && idx_out) idx_out[row] = b;
            }
            TCP(4);
            if (sums) {
                __syncwarp();
                if (lane == 0) mbar_arrive(&dec_full[i & 1]);        // all eight scanning warps arrive (four of them wrote)
            }
            named_bar_sync(1, kTcScan);          // the partial minima of this tile have been read: the next scan may overwrite them
            TCP(5);
            TCP(6);
        }
#ifdef FPC_KMEANS_TC_PROF
        if (tid == 0 && blockIdx.x == 0) {
            for (int k2 = 0; k2 < 7; ++k2) const_cast<unsigned long long *>(meta->prof)[k2] = (unsigned long long)pf[k2];
            const_cast<unsigned long long *>(meta->prof)[7] = (unsigned long long)my_tiles;
        }
#endif
#undef TCP
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tb, 512);
}

size_t kmeans_tc_smem_bytes(int K)
{
    const int Kp = (K + 127) / 128 * 128;
    return tc_image_bytes(Kp) + (size_t)kTcASlots * tc::kTileBytes + (size_t)kTcRawSlots * kTcRawFloats * 4 + 3 * 128 * 4 * 4 + 128 * 4 +
           128 * 4 + (1 + 2 * kTcASlots + 2 * kTcDSlots + 2 * kTcRawSlots + 4) * 8 + 16;
}

int num_sms();

// d_pack: tc_pack_bytes(K) bytes of scratch (image + shadow); acc_* are the tables the atomics go to (R replicas)
int run_kmeans_assign_tc(const float *d_data, long N, const double *d_cb, int K, double *acc_sums, double *acc_counts,
                         int32_t *d_idx, int R, void *d_pack, cudaStream_t st)
{
    const int sms = num_sms();
    if (sms <= 0) return cuda_fail(cudaErrorNoDevice);
    kmeans_tc_pack_kernel<<<1, 1024, 0, st>>>(d_cb, K, (unsigned char *)d_pack);
    FPC_LAUNCH_CHECK();
    static bool configured[kMaxDevices] = {};
    { const int rc = ensure_dynamic_smem(kmeans_assign_tc_kernel, (int)kmeans_tc_smem_bytes(FPC_MAX_VQ_ENTRIES), configured);
      if (rc != FPC_OK) return rc; }
    if ((reinterpret_cast<uintptr_t>(d_data) & 15) != 0) return FPC_ERR_ARG;      // bulk copies of whole tiles (torch allocations are 256-byte aligned)
    const long ntiles = (N + 127) / 128;
    const int grid = (int)(ntiles < sms ? ntiles : sms);
    kmeans_assign_tc_kernel<<<grid, kTcThreads, kmeans_tc_smem_bytes(K), st>>>(d_data, N, d_cb, K, (const unsigned char *)d_pack,
                                                                                acc_sums, acc_counts, d_idx, R, ntiles);
    FPC_LAUNCH_CHECK();
    return FPC_OK;
}

size_t kmeans_tc_pack_bytes(int K) { return tc_pack_bytes(K); }

// ------------------------------------------------------------------------------------------
// selftest: the raw screened scores of up to 128 vectors against a codebook, un-scaled, so that a test can measure
// the error of the tensor-core path against float64 (the accumulation term of the error budget in fpc_tc.cuh)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1)
tc_scores_selftest_kernel(const float *__restrict__ x, int n, int K, const unsigned char *__restrict__ packed, float *__restrict__ out)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    const KmeansTcMeta *meta = reinterpret_cast<const KmeansTcMeta *>(packed);
    const int Kp = (K + 127) / 128 * 128, nchunks = Kp >> 7;
    unsigned char *sB = smem, *sA = smem + tc_image_bytes(Kp);
    __shared__ uint64_t bar[2];
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) umma::tmem_alloc(&tmem_base, 128);
    if (tid == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); mbar_fence_init(); }
    const float beta = meta->beta;
    {
        float xs[kDim];
#pragma unroll
        for (int d = 0; d < kDim; ++d) xs[d] = tid < n ? -2.0f * beta * x[(size_t)tid * kDim + d] : 0.0f;
        const __half one = __float2half_rn(1.0f);
        tc::store_row<true>(sA, tid, xs, one, one, one);
    }
    umma::fence_async_smem();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tb = tmem_base;
    if (tid == 0) {
        const uint32_t bytes = (uint32_t)tc_image_bytes(Kp);
        mbar_arrive_expect_tx(&bar[0], bytes);
        for (uint32_t off = 0; off < bytes; off += tc::kTileBytes) bulk_g2s(sB + off, packed + 256 + off, tc::kTileBytes, &bar[0]);
    }
    mbar_wait(&bar[0], 0);
    for (int c = 0; c < nchunks; ++c) {
        if (tid == 0) {
            const uint32_t idesc = tc::instr_desc_f16(128, 128);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
                tc::mma_f16(tb, umma::smem_desc(smem_u32(sA) + ks * tc::kSlabBytes, 128),
                            umma::smem_desc(smem_u32(sB) + c * tc::kTileBytes + ks * tc::kSlabBytes, 128), idesc, ks > 0);
            umma::commit(&bar[1]);
        }
        mbar_wait(&bar[1], c & 1);
        umma::fence_after_sync();
        for (int c0 = 0; c0 < 128; c0 += 32) {
            uint32_t v[32];
            tc::tmem_ld32(tb + ((uint32_t)(warp * 32) << 16) + c0, v);
            tc::tmem_ld_wait(v);
            if (tid < n)
#pragma unroll
                for (int j = 0; j < 32; ++j) out[(size_t)tid * Kp + c * 128 + c0 + j] = __uint_as_float(v[j]) / (beta * beta);
        }
        umma::fence_before_sync();
        __syncthreads();
        umma::fence_after_sync();
    }
    if (warp == 0) umma::tmem_dealloc(tb, 128);
}

}  // namespace fpc

// scores[v][k] = ||c_k||^2 - 2 <x_v, c_k> as the tensor-core screen computes them (n <= 128 vectors, K <= 1024 centroids;
// out: n x Kp floats, Kp = K rounded up to 128; d_pack: fpc_kmeans_workspace_bytes(n, K) bytes of scratch)
extern "C" int fpc_selftest_tc_scores(const float *d_x, int n, const double *d_cb, int K, float *d_out, void *d_pack, void *stream)
{
    using namespace fpc;
    if (!d_x || !d_cb || !d_out || !d_pack) return FPC_ERR_ARG;
    if (n < 1 || n > 128 || K < 1 || K > FPC_MAX_VQ_ENTRIES) return FPC_ERR_SHAPE;
    cudaStream_t st = (cudaStream_t)stream;
    kmeans_tc_pack_kernel<<<1, 1024, 0, st>>>(d_cb, K, (unsigned char *)d_pack);
    FPC_LAUNCH_CHECK();
    const int Kp = (K + 127) / 128 * 128;
    const size_t smem = tc_image_bytes(Kp) + tc::kTileBytes;
    FPC_CUDA_TRY(cudaFuncSetAttribute(tc_scores_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tc_scores_selftest_kernel<<<1, 128, smem, st>>>(d_x, n, K, (const unsigned char *)d_pack, d_out);
    FPC_LAUNCH_CHECK();
    return FPC_OK;
}
