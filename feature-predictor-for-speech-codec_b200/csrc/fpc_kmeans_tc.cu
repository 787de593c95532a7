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
// One CTA per SM, 13 warps:
//   warps 0..7   scan.  Warp w reads TMEM lanes 32 (w % 4) .. +31 (one vector per lane) and the column half w / 4
//                of every 128-centroid chunk; then merge, decision, fallback, float64 atomics for the sums.
//   warps 8..11  load.  Thread r converts vector r of the next 128-vector tile into the fp16-pair A operand row
//                (scaled by -2 beta) in a 4-deep ring of 16 KB tiles.
//   warp 12      lane 0 issues the MMAs: per tile and 128-centroid chunk four K = 16 slabs into one of four
//                128-column TMEM slots; tcgen05.commit releases the A tile and publishes the slot.
// The B operand (the whole codebook, <= 1024 x 128 B) is resident in shared memory: one bulk copy per launch from
// the image kmeans_tc_pack_kernel writes (also: the fp32 shadow and the scale).
#include "fpc_common.cuh"
#include "fpc_vq.cuh"
#include "fpc_tc.cuh"

namespace fpc {

constexpr int kTcScan = 256, kTcLoad = 128, kTcThreads = kTcScan + kTcLoad + 32;
constexpr int kTcASlots = 4, kTcDSlots = 4;
constexpr int kTcShadowLd = 20;                    // fp32 shadow row: c0..c16, -, ||c||^2, -
// slack of a TC decision: 2^-15 (||x|| + Cmax)^2 = 512 u R (fpc_tc.cuh: the operand format costs <= 10 u R per score,
// the accumulation is measured; a comparison involves two scores)
constexpr float kTcSlackRel = 3.0517578125e-5f;
constexpr float kF32SlackRel = 7.62939453125e-6f; // 128 u R: the fp32 screen of the fallback (fpc_kmeans.cu)

struct KmeansTcMeta { float beta, cmax; int K, Kp; };
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
    if (threadIdx.x == 0) { meta->beta = beta; meta->cmax = __fsqrt_ru(n2max); meta->K = K; meta->Kp = Kp; }
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

struct TcSmem {
    static constexpr int offA = 0;                                    // after the B image (dynamic offset)
};

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
    float *pm = reinterpret_cast<float *>(sA + kTcASlots * tc::kTileBytes);        // [128][4]: best, second, column (hh = 1 half)
    int *dec = reinterpret_cast<int *>(pm + 128 * 4);                               // [128] decided centroid (or -1)
    uint64_t *bars = reinterpret_cast<uint64_t *>(dec + 128);
    uint64_t *b_full = bars, *a_full = bars + 1, *a_empty = a_full + kTcASlots, *d_full = a_empty + kTcASlots,
             *d_empty = d_full + kTcDSlots;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(d_empty + kTcDSlots);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long my_tiles = ntiles > (long)blockIdx.x ? (ntiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;

    if (tid == 0) {
        mbar_init(b_full, 1);
        for (int s = 0; s < kTcASlots; ++s) { mbar_init(&a_full[s], kTcLoad / 32); mbar_init(&a_empty[s], 1); }
        for (int s = 0; s < kTcDSlots; ++s) { mbar_init(&d_full[s], 1); mbar_init(&d_empty[s], kTcScan / 32); }
        mbar_fence_init();
    }
    if (warp == 0) umma::tmem_alloc(tmem_slot, 512);
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tb = *tmem_slot;
    const float beta = meta->beta, cmax = meta->cmax;

    if (warp == kTcThreads / 32 - 1) {
        // ---------------- MMA issuer (+ the one bulk copy of the codebook image) ----------------
        if (lane == 0) {
            const uint32_t bytes = (uint32_t)tc_image_bytes(Kp);
            mbar_arrive_expect_tx(b_full, bytes);
            for (uint32_t off = 0; off < bytes; off += tc::kTileBytes)
                bulk_g2s(sB + off, packed + 256 + off, tc::kTileBytes, b_full);
            mbar_wait(b_full, 0);
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
                        tc::mma_f16(tb + (uint32_t)ds * 128u, umma::smem_desc(a0 + sa * tc::kTileBytes + ks * tc::kSlabBytes, 128),
                                    umma::smem_desc(b0 + c * tc::kTileBytes + ks * tc::kSlabBytes, 128), idesc, ks > 0);
                    umma::commit(&d_full[ds]);
                }
                umma::commit(&a_empty[sa]);
            }
        }
    } else if (warp >= kTcScan / 32) {
        // ---------------- loaders: vectors -> fp16-pair A tiles ----------------
        const int r = tid - kTcScan;
        const __half one = __float2half_rn(1.0f);
        for (long i = 0; i < my_tiles; ++i) {
            const int sa = (int)(i % kTcASlots);
            const uint32_t na = (uint32_t)(i / kTcASlots);
            const long row = ((long)blockIdx.x + i * gridDim.x) * 128 + r;
            float xs[kDim];
#pragma unroll
            for (int d = 0; d < kDim; ++d) xs[d] = row < N ? __ldg(data + row * kDim + d) : 0.0f;
#pragma unroll
            for (int d = 0; d < kDim; ++d) {
                float v = -2.0f * beta * xs[d];                       // exact (powers of two)
                if (!(fabsf(v) <= tc::kMaxScaledX)) v = 0.0f;         // out of range (or NaN): the scan side sends the row to the exact path
                xs[d] = v;
            }
            if (na > 0) mbar_wait(&a_empty[sa], (na - 1) & 1u);
            tc::store_row<true>(sA + sa * tc::kTileBytes, r, xs, one, one, one);
            umma::fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&a_full[sa]);
        }
    } else {
        // ---------------- scanners ----------------
        const int q = warp & 3, hh = warp >> 2;
        const int r = 32 * q + lane;
        const float inf = __int_as_float(0x7f800000);
        const float *shadow = reinterpret_cast<const float *>(packed + 256 + tc_image_bytes(Kp));
        uint32_t u = 0;
        for (long i = 0; i < my_tiles; ++i) {
            const long row = ((long)blockIdx.x + i * gridDim.x) * 128 + r;
            const bool valid = row < N;
            float x[kDim];
#pragma unroll
            for (int d = 0; d < kDim; ++d) x[d] = valid ? __ldg(data + row * kDim + d) : 0.0f;
            tc::Scan sc;
            sc.reset();
            for (int c = 0; c < nchunks; ++c, ++u) {
                const int ds = (int)(u % kTcDSlots);
                mbar_wait(&d_full[ds], (u / kTcDSlots) & 1u);
                umma::fence_after_sync();
                uint32_t v0[32], v1[32];
                const uint32_t ta = tb + ((uint32_t)(32 * q) << 16) + (uint32_t)(ds * 128 + 64 * hh);
                tc::tmem_ld32(ta, v0);
                tc::tmem_ld32(ta + 32, v1);
                tc::tmem_ld_wait2(v0, v1);
                umma::fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(&d_empty[ds]);
                sc.feed(v0, v1, 2 * c);
            }
            float best, second;
            int jb;
            sc.finish(best, second, jb);
            int col = 128 * (sc.ga >> 1) + 64 * hh + 32 * (sc.ga & 1) + jb;
            if (hh == 1) { pm[r * 4] = best; pm[r * 4 + 1] = second; pm[r * 4 + 2] = __int_as_float(col); }
            named_bar_sync(1, kTcScan);
            if (hh == 0) {
                const float ob = pm[r * 4], os = pm[r * 4 + 1];
                const int oc = __float_as_int(pm[r * 4 + 2]);
                second = fminf(fminf(second, os), fmaxf(best, ob));
                if (ob < best) { best = ob; col = oc; }
                // per-vector constants of the decision
                float nx = 0.0f, amax = 0.0f;
#pragma unroll
                for (int d = 0; d < kDim; ++d) { nx = __fmaf_ru(x[d], x[d], nx); amax = fmaxf(amax, fabsf(x[d])); }
                const float rr = __fadd_ru(__fsqrt_ru(nx), cmax);
                const float rr2 = __fmul_ru(rr, rr);
                const float slack_tc = __fmul_ru(__fmul_ru(rr2, kTcSlackRel), beta * beta);   // scores are scaled by beta^2
                const bool in_range = 2.0f * beta * amax <= tc::kMaxScaledX;                   // NaN -> false
                bool decided = in_range && col < K && (second > __fadd_ru(best, slack_tc));
                int b = decided ? col : -1;
                // the undecided vectors (near-ties, duplicate centroids, out-of-range rows): one at a time by the whole
                // warp -- fp32 screen of all K centroids (32 per lane), then float64 direct-form distances of the
                // candidates inside the fp32 slack, ascending k with strict < (numpy's first minimum)
                unsigned todo = __ballot_sync(0xffffffffu, valid && !decided);
                while (todo) {
                    const int src = __ffs(todo) - 1;
                    todo &= todo - 1;
                    float xs[kDim];
#pragma unroll
                    for (int d = 0; d < kDim; ++d) xs[d] = __shfl_sync(0xffffffffu, x[d], src);
                    const float slack32 = __fadd_ru(__fmul_ru(__shfl_sync(0xffffffffu, rr2, src), kF32SlackRel), 1e-30f);
                    float m = inf;
                    for (int k = lane; k < K; k += 32) m = fminf(m, tc_screen17(xs, shadow + (size_t)k * kTcShadowLd));
                    for (int off = 16; off > 0; off >>= 1) m = fminf(m, __shfl_xor_sync(0xffffffffu, m, off));
                    const float thr = __fadd_ru(m, slack32);
                    double bd = Rn<double>::inf();
                    int bk = 0x7fffffff;
                    for (int k = lane; k < K; k += 32) {
                        if (tc_screen17(xs, shadow + (size_t)k * kTcShadowLd) <= thr) {
                            double xd[kDim], cd[kDim];
#pragma unroll
                            for (int d = 0; d < kDim; ++d) { xd[d] = (double)xs[d]; cd[d] = cb[(size_t)k * kDim + d]; }
                            const double dd = dist17<double>(xd, cd);
                            if (dd < bd || bk == 0x7fffffff) { bd = dd; bk = k; }     // ascending k per lane: first minimum kept
                        }
                    }
                    warp_argmin(bd, bk);
                    if (bk == 0x7fffffff) bk = 0;         // NaN input: numpy's argmin returns the first NaN; index 0 here
                    if (lane == src) b = bk;
                }
                dec[r] = valid ? b : -1;
                if (valid && idx_out) idx_out[row] = b;
            }
            named_bar_sync(1, kTcScan);
            // sums and counts (cb_func.py:82-86): float64 atomics, the 17 dimensions split between the two warps of a row
            if (sums) {
                const int b = dec[r];
                if (b >= 0) {
                    double *s2 = sums;
                    double *c2 = counts;
                    if (R > 1) {
                        const int rep = (int)(((long)blockIdx.x * kTcScan + tid) % R);
                        s2 += (size_t)rep * K * kDim;
                        c2 += (size_t)rep * K;
                    }
                    s2 += (size_t)b * kDim;
                    if (hh == 0) {
#pragma unroll
                        for (int d = 0; d < 9; ++d) atomicAdd(s2 + d, (double)x[d]);
                    } else {
#pragma unroll
                        for (int d = 9; d < kDim; ++d) atomicAdd(s2 + d, (double)x[d]);
                        atomicAdd(c2 + b, 1.0);
                    }
                }
            }
        }
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tb, 512);
}

size_t kmeans_tc_smem_bytes(int K)
{
    const int Kp = (K + 127) / 128 * 128;
    return tc_image_bytes(Kp) + (size_t)kTcASlots * tc::kTileBytes + 128 * 4 * 4 + 128 * 4 + (1 + 2 * kTcASlots + 2 * kTcDSlots) * 8 + 16;
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
    const long ntiles = (N + 127) / 128;
    const int grid = (int)(ntiles < sms ? ntiles : sms);
    kmeans_assign_tc_kernel<<<grid, kTcThreads, kmeans_tc_smem_bytes(K), st>>>(d_data, N, d_cb, K, (const unsigned char *)d_pack,
                                                                                acc_sums, acc_counts, d_idx, R, ntiles);
    FPC_LAUNCH_CHECK();
    return FPC_OK;
}

size_t kmeans_tc_pack_bytes(int K) { return tc_pack_bytes(K); }

}  // namespace fpc
