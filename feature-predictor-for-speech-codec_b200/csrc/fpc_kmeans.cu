// fpc_kmeans.cu -- Lloyd iteration of the codebook learner.
//
// Reference: /root/reference/src/quantization/cb_func.py
//   find_nearest (:56-68)   dist = np.sum((data - codebook) ** 2, -1) in float64 (float32 data
//                           minus float64 codebook promotes), np.argmin -> first minimum
//   update       (:71-100)  per-centroid float64 sums and counts, codebook = sum / (count + 1e-20)
//   quantize     (:103-112) nearest-centroid gather
//
// Exactness.  The reference decides by float64 direct-form distances (numpy's roundings, dist17<double> in
// fpc_vq.cuh reproduces them), first minimum on ties; the indices here are bit-identical to that.  To avoid 51 FP64
// operations per (vector, centroid) pair every pair is screened in fp32 with the expanded form
//     s_k = ||c_k||^2 - 2 <x, c_k>             (one 17-step FMA chain seeded with ||c_k||^2; d_k = s_k + ||x||^2)
// on the fp32 shadow (float)c_k.  With u = 2^-24 and R = (||x|| + ||c_k||)^2 the 17 FMA roundings, the rounding of
// the shadow and of the stored norm give |s_k - (d_k - ||x||^2)| <= 21 u R (numpy's own float64 evaluation of d_k
// is exact at this scale), so two screened values of one vector compare with error <= 42 u R.
// Issue form: a thread owns two vectors as the two lanes of fma.rn.f32x2 and sweeps eight centroids per step,
//     acc[r] = (-2 x_d of vector 0, -2 x_d of vector 1) * broadcast(c[r][d]) + acc[r],   r = 0..7,
// so 17 packed FMAs cover two (vector, centroid) pairs and the packed pair operand is the one the eight
// consecutive instructions share.  That matters: on sm_100a FFMA2 issues every 2 cycles only when the 64-bit
// pair operand comes from the operand-reuse cache; with three fresh 64-bit operands it takes 3
// (tools/ubench_ffma2.cu).  The shadow is stored in blocks of four rows, transposed ([d][r], one LDS.128 per
// dimension of four rows), with ||c||^2 as ready-made (n, n) pairs for the seed.  One divergence-free
// sweep keeps the two smallest s_k; with the per-vector slack A = 128 u (||x|| + Cmax)^2 (rounded up), a runner-up
// more than A above the smallest proves the smallest is the exact argmin -- no float64 work at all.  Otherwise
// (near-ties, duplicate centroids) the candidates within A of the smallest are re-evaluated in float64, ascending
// k with strict <.  The screen never picks a winner it cannot prove.
//
// Sums.  The reference accumulates in data order on one thread.  Here every CTA first reduces
// its vectors into per-warp-aggregated float64 atomics on the global (K,17) table; the order
// of additions therefore differs and centroids agree with the reference to ~1e-15 relative
// (tests state 1e-12).  With several ranks the caller all-reduces sums and counts (NCCL)
// between fpc_kmeans_assign_accumulate and fpc_kmeans_finalize.
#include <stdlib.h>

#include "fpc_common.cuh"
#include "fpc_vq.cuh"

namespace fpc {

constexpr int kKmThreads = 512;
constexpr int kKmV = 2;        // vectors per thread = the two lanes of the packed FMA (4 per thread at 256 threads measured 24.2 vs 26.5 iters/s)
constexpr int kKmLd64 = 18;   // float64 codeword row stride in shared memory (16-byte aligned rows)
constexpr int kKmLd32 = 20;   // float32 shadow row stride (float4 aligned)
constexpr int kKmRows = 4;    // centroid rows per sweep step (the shadow is padded to a multiple)
constexpr int kKmNB = 2;      // blocks per sweep step: 8 rows share each packed pair of vectors
constexpr int kKmPad = kKmRows * kKmNB;     // rows the shadow is padded to
constexpr int kStepRows = kKmRows * kKmNB;
static_assert(kKmRows == 4, "the shadow block layout is one float4 per dimension");

// Screen value of one (vector, centroid) pair: s = ||c||^2 + sum_d (-2 x_d) c_d as one ascending-d FMA chain seeded
// with ||c||^2.  The sweep evaluates it two vectors at a time (one packed lane each, see below); this scalar form is
// the same chain, so the rescan reproduces the sweep's value bit for bit.
__device__ __forceinline__ float screen17(const float (&x)[kDim], const float *__restrict__ cb32, int k)
{
    const float *c = cb32 + (k >> 2) * (kKmLd32 * kKmRows) + (k & 3);
    float s = c[18 * kKmRows + (k & 3)];      // ||c||^2 of row r = k & 3 sits at block slot 18*4 + 2r (c is offset by r)
#pragma unroll
    for (int d = 0; d < kDim; ++d) s = __fmaf_rn(-2.0f * x[d], c[d * kKmRows], s);
    return s;
}

template <typename TD>
__device__ __forceinline__ double exact17(const TD (&x)[kDim], const double *__restrict__ c)
{
    double xd[kDim], cd[kDim];
#pragma unroll
    for (int d = 0; d < kDim; ++d) { xd[d] = (double)x[d]; cd[d] = c[d]; }
    return dist17<double>(xd, cd);
}

// one vector per thread; codebook (float64 + fp32 shadow) resident in shared memory.
// TD = float: the training vectors as the encoder leaves them.  TD = double: the vectors of a later stage,
// r = quantize(cb, r) - r, which the reference keeps in float64 (train_cb.py:200).  The screen then runs on the
// float32 ROUNDING of the vector -- that moves a screened value by at most 2 u ||x|| ||c|| <= u R / 2, far inside the
// slack of 128 u R -- while the float64 re-evaluation of near-ties and the accumulation use the float64 vector itself.
template <typename TD>
__global__ void __launch_bounds__(kKmThreads, 1)
kmeans_assign_kernel(const TD *__restrict__ data, long N, const double *__restrict__ cb, int K,
                     double *__restrict__ sums, double *__restrict__ counts, int32_t *__restrict__ idx_out, int R)
{
    extern __shared__ __align__(16) unsigned char smem[];
    double *cb64 = reinterpret_cast<double *>(smem);                              // [K][18]
    float *cb32 = reinterpret_cast<float *>(smem + (size_t)K * kKmLd64 * 8);      // [Kp][20]: c0..c16, -, ||c||^2 twice
    const int Kp = (K + kKmPad - 1) / kKmPad * kKmPad;      // padded with rows that can never win (s = +inf)
    // shadow layout: blocks of kKmRows = 4 rows, transposed: blk[d][r] for d < 17, blk[17] unused, then the four
    // ||c||^2 as duplicated pairs (n0,n0,n1,n1 | n2,n2,n3,n3), so one LDS.128 feeds one dimension of four rows
    for (int i = threadIdx.x; i < Kp * kKmLd32; i += kKmThreads)
        cb32[i] = (i % (kKmLd32 * kKmRows)) >= 18 * kKmRows ? __int_as_float(0x7f800000) : 0.0f;
    __syncthreads();
    for (int i = threadIdx.x; i < K * kDim; i += kKmThreads) {
        const int k = i / kDim, d = i - k * kDim;
        const double v = cb[i];
        cb64[k * kKmLd64 + d] = v;
        cb32[(k >> 2) * (kKmLd32 * kKmRows) + d * kKmRows + (k & 3)] = (float)v;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < K; k += kKmThreads) {          // ||c_k||^2 in float64, rounded, next to c16
        double n2 = 0.0;
#pragma unroll
        for (int d = 0; d < kDim; ++d) n2 += cb64[k * kKmLd64 + d] * cb64[k * kKmLd64 + d];
        cb32[(k >> 2) * (kKmLd32 * kKmRows) + 18 * kKmRows + 2 * (k & 3)] = (float)n2;
        cb32[(k >> 2) * (kKmLd32 * kKmRows) + 18 * kKmRows + 2 * (k & 3) + 1] = (float)n2;
    }
    // Cmax^2 = max_k ||c_k||^2 (fp32, rounded up) for the absolute slack of the screen
    __shared__ float s_c2[kKmThreads / 32];
    {
        float c2 = 0.0f;
        for (int k = threadIdx.x; k < K; k += kKmThreads) {
            float a = 0.0f;
#pragma unroll
            for (int d = 0; d < kDim; ++d) {
                const float v = __double2float_ru(fabs(cb64[k * kKmLd64 + d]));
                a = __fmaf_ru(v, v, a);
            }
            c2 = fmaxf(c2, a);
        }
        for (int off = 16; off > 0; off >>= 1) c2 = fmaxf(c2, __shfl_xor_sync(0xffffffffu, c2, off));
        if ((threadIdx.x & 31) == 0) s_c2[threadIdx.x >> 5] = c2;
    }
    __syncthreads();
    float cmax2 = 0.0f;
    for (int w = 0; w < kKmThreads / 32; ++w) cmax2 = fmaxf(cmax2, s_c2[w]);
    const float cmax = __fsqrt_ru(cmax2);                      // upper bound of max ||c_k||

    const int lane = threadIdx.x & 31;
    const long per_round = (long)kKmThreads * kKmV;
    const long stride = (long)gridDim.x * per_round;
    const long nround = (N + stride - 1) / stride;
    for (long it = 0; it < nround; ++it) {
        const long i0 = it * stride + (long)blockIdx.x * per_round + threadIdx.x;     // vectors i0 + j * kKmThreads
        // xp[g][d] = (-2 x_d of vector 2g, -2 x_d of vector 2g+1): one packed lane per vector
        float2 xp[kKmV / 2][kDim];
        float slack[kKmV], m1[kKmV], m2[kKmV];
        int bi[kKmV];
#pragma unroll
        for (int j = 0; j < kKmV; ++j) {
            const long i = i0 + (long)j * kKmThreads;
            float x[kDim];
#pragma unroll
            for (int d = 0; d < kDim; ++d) x[d] = i < N ? (float)__ldg(data + i * kDim + d) : 0.0f;
            // per-vector screen constants: -2 x (exact), ||x||^2 and the slack A = 128 u (||x|| + Cmax)^2, rounded up
#pragma unroll
            for (int d = 0; d < kDim; ++d) {
                if (j & 1) xp[j >> 1][d].y = -2.0f * x[d];
                else xp[j >> 1][d].x = -2.0f * x[d];
            }
            float nx = 0.0f;
#pragma unroll
            for (int d = 0; d < kDim; ++d) nx = __fmaf_ru(x[d], x[d], nx);
            const float rr = __fadd_ru(__fsqrt_ru(nx), cmax);
            slack[j] = __fadd_ru(__fmul_ru(__fmul_ru(rr, rr), 7.62939453125e-6f), 1e-30f);   // 128 * 2^-24
            m1[j] = __int_as_float(0x7f800000); m2[j] = m1[j]; bi[j] = 0;
        }
        // One divergence-free sweep keeps the two smallest screened values of every vector.  A step takes kKmRows
        // centroid rows; for each dimension the packed pair of two vectors is multiplied by the broadcast scalar
        // c[r][d] of every row in turn (fma.rn.f32x2, scalar-broadcast form: the pair operand is reused by the
        // kKmRows consecutive instructions, which is what lets FFMA2 issue every 2 cycles), 17 packed FMAs per two
        // (vector, centroid) pairs and nothing else.  All s_k of a vector share the ||x||^2 term, so two of them
        // compare with error <= 2 * 19 u R: if the runner-up is more than the slack above the smallest, the smallest
        // is the exact argmin (no float64 work at all); otherwise the vector is ambiguous (near-tie, duplicate
        // centroids) and the candidates inside the slack are decided in float64 exactly as numpy does, ascending k
        // with strict <, i.e. first minimum.
        for (int k = 0; k < Kp; k += kStepRows) {
            const float4 *blk = reinterpret_cast<const float4 *>(cb32 + k * kKmLd32);
            float2 acc[kStepRows][kKmV / 2];
#pragma unroll
            for (int nb = 0; nb < kKmNB; ++nb) {
                const float4 na = blk[nb * kKmLd32 + 18], nb4 = blk[nb * kKmLd32 + 19], c0 = blk[nb * kKmLd32];
                const float cr[kKmRows] = {c0.x, c0.y, c0.z, c0.w};
                const float2 cn[kKmRows] = {make_float2(na.x, na.y), make_float2(na.z, na.w), make_float2(nb4.x, nb4.y),
                                            make_float2(nb4.z, nb4.w)};
#pragma unroll
                for (int g = 0; g < kKmV / 2; ++g)
#pragma unroll
                    for (int r = 0; r < kKmRows; ++r) acc[nb * kKmRows + r][g] = fma2(xp[g][0], cr[r], cn[r]);
            }
#pragma unroll
            for (int d = 1; d < kDim; ++d) {
                float cr[kStepRows];
#pragma unroll
                for (int nb = 0; nb < kKmNB; ++nb) {
                    const float4 cd = blk[nb * kKmLd32 + d];
                    cr[nb * kKmRows] = cd.x; cr[nb * kKmRows + 1] = cd.y; cr[nb * kKmRows + 2] = cd.z; cr[nb * kKmRows + 3] = cd.w;
                }
#pragma unroll
                for (int g = 0; g < kKmV / 2; ++g)
#pragma unroll
                    for (int r = 0; r < kStepRows; ++r) acc[r][g] = fma2(xp[g][d], cr[r], acc[r][g]);
            }
            // (a (lo, hi) merge tree per step with FMNMX3 needs fewer ALU instructions, but then the winning row has
            // to be recovered after the sweep with per-thread shared-memory addresses; measured 37 % slower)
#pragma unroll
            for (int r = 0; r < kStepRows; ++r)
#pragma unroll
                for (int j = 0; j < kKmV; ++j) {
                    const float s = (j & 1) ? acc[r][j >> 1].y : acc[r][j >> 1].x;
                    const bool p = s < m1[j];
                    m2[j] = fminf(m2[j], fmaxf(m1[j], s));
                    m1[j] = fminf(m1[j], s);
                    bi[j] = p ? k + r : bi[j];
                }
        }
#pragma unroll
        for (int j = 0; j < kKmV; ++j) {
            const long i = i0 + (long)j * kKmThreads;
            const bool valid = i < N;
            const float thr = __fadd_ru(m1[j], slack[j]);
            float x[kDim];
            TD xe[kDim];                  // the vector as given: what the exact distances and the sums use
            const bool need_x = (sums != nullptr) || !(m2[j] > thr);
            if (need_x) {
#pragma unroll
                for (int d = 0; d < kDim; ++d) {
                    xe[d] = valid ? __ldg(data + i * kDim + d) : (TD)0;   // L1/L2 hit
                    x[d] = (float)xe[d];
                }
            }
            if (!(m2[j] > thr)) {
                double best = 0.0;
                bool have = false;
                for (int k = 0; k < K; ++k) {
                    const float s = screen17(x, cb32, k);
                    if (s <= thr) {
                        const double d = exact17<TD>(xe, cb64 + k * kKmLd64);
                        if (!have || d < best) { best = d; bi[j] = k; have = true; }
                    }
                }
            }
            const int b = bi[j];
            if (valid && idx_out) idx_out[i] = b;

            // accumulate.  Small codebooks (the first steps of the grow-by-one schedule, cb_func.py:34-47) would
            // hammer a handful of addresses, so there the lanes that chose the same centroid are merged first
            // (uniform full-mask shuffles, ascending lane order).
            if (sums) {
                const int key = valid ? b : -1;
                if (R > 1) {
                    // Small codebooks (the grow-by-one schedule of vq_train spends most of its 4102 iterations there,
                    // cb_func.py:34-47) would hammer a handful of addresses: the table is replicated R times in the
                    // workspace (R K >= 512 rows), every thread adds into its own replica -- lanes of a warp never
                    // collide while R >= 32 -- and kmeans_fold_kernel sums the replicas in a fixed order.
                    if (valid) {
                        const int rep = (int)(((long)blockIdx.x * kKmThreads + threadIdx.x + (long)j * 7) % R);
                        double *s2 = sums + ((size_t)rep * K + b) * kDim;
#pragma unroll
                        for (int d = 0; d < kDim; ++d) atomicAdd(s2 + d, (double)xe[d]);
                        atomicAdd(counts + (size_t)rep * K + b, 1.0);
                    }
                } else if (K <= 32) {
                    // no workspace: lanes that chose the same centroid are merged first (uniform full-mask shuffles)
                    const unsigned peers = __match_any_sync(0xffffffffu, key);
                    const int leader = __ffs(peers) - 1;
                    double acc[kDim];
#pragma unroll
                    for (int d = 0; d < kDim; ++d) acc[d] = 0.0;
                    for (int src = 0; src < 32; ++src) {
                        const bool take = (peers >> src) & 1u;
#pragma unroll
                        for (int d = 0; d < kDim; ++d) {
                            const double v = __shfl_sync(0xffffffffu, (double)xe[d], src);
                            if (take) acc[d] += v;
                        }
                    }
                    if (valid && lane == leader) {
#pragma unroll
                        for (int d = 0; d < kDim; ++d) atomicAdd(sums + (size_t)b * kDim + d, acc[d]);
                        atomicAdd(counts + b, (double)__popc(peers));
                    }
                } else if (valid) {
#pragma unroll
                    for (int d = 0; d < kDim; ++d) atomicAdd(sums + (size_t)b * kDim + d, (double)xe[d]);
                    atomicAdd(counts + b, 1.0);
                }
            }
        }
    }
}

// sums[k][d] += sum_r ws_sums[r][k][d], counts[k] += sum_r ws_counts[r][k]  (ascending r: a fixed order)
__global__ void kmeans_fold_kernel(const double *__restrict__ ws_sums, const double *__restrict__ ws_counts, int R, int K,
                                   double *__restrict__ sums, double *__restrict__ counts)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < K * kDim) {
        double a = 0.0;
        for (int r = 0; r < R; ++r) a += ws_sums[(size_t)r * K * kDim + t];
        sums[t] += a;
    } else if (t < K * kDim + K) {
        const int k = t - K * kDim;
        double a = 0.0;
        for (int r = 0; r < R; ++r) a += ws_counts[(size_t)r * K + k];
        counts[k] += a;
    }
}

// codebook = sums / (counts + 1e-20); statistics of cb_func.py:92-97
__global__ void kmeans_finalize_kernel(double *__restrict__ sums, double *__restrict__ counts, int K,
                                       double n_total, double *__restrict__ cb_out, double *__restrict__ stats, int reset)
{
    __shared__ double s_min[32], s_max[32], s_empty[32], s_w2[32];
    __shared__ double s_n;
    if (!(n_total > 0.0)) {
        // nb_vectors = sum of the counts (exact: they are integers), so callers need not bring it from the host
        double c = 0.0;
        for (int k = threadIdx.x; k < K; k += blockDim.x) c += counts[k];
        for (int off = 16; off > 0; off >>= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
        if ((threadIdx.x & 31) == 0) s_min[threadIdx.x >> 5] = c;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += s_min[w];
            s_n = t;
        }
        __syncthreads();
        n_total = s_n;
        __syncthreads();
    }
    double mn = 1e300, mx = -1e300, em = 0.0, w2 = 0.0;
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        const double c = counts[k];
        const double den = c + 1e-20;
#pragma unroll
        for (int d = 0; d < kDim; ++d) cb_out[(size_t)k * kDim + d] = sums[(size_t)k * kDim + d] / den;
        if (reset) {            // leave the accumulator zeroed for the next iteration (no separate memset launches)
#pragma unroll
            for (int d = 0; d < kDim; ++d) sums[(size_t)k * kDim + d] = 0.0;
            counts[k] = 0.0;
        }
        mn = fmin(mn, c); mx = fmax(mx, c);
        if (c == 0.0) em += 1.0;
        const double f = c / n_total;
        w2 += f * f;
    }
    for (int off = 16; off > 0; off >>= 1) {
        mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, off));
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, off));
        em += __shfl_xor_sync(0xffffffffu, em, off);
        w2 += __shfl_xor_sync(0xffffffffu, w2, off);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { s_min[warp] = mn; s_max[warp] = mx; s_empty[warp] = em; s_w2[warp] = w2; }
    __syncthreads();
    if (warp == 0) {
        const int nw = blockDim.x >> 5;
        mn = lane < nw ? s_min[lane] : 1e300;
        mx = lane < nw ? s_max[lane] : -1e300;
        em = lane < nw ? s_empty[lane] : 0.0;
        w2 = lane < nw ? s_w2[lane] : 0.0;
        for (int off = 16; off > 0; off >>= 1) {
            mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, off));
            mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, off));
            em += __shfl_xor_sync(0xffffffffu, em, off);
            w2 += __shfl_xor_sync(0xffffffffu, w2, off);
        }
        if (lane == 0 && stats) { stats[0] = mn; stats[1] = mx; stats[2] = em; stats[3] = w2; stats[4] = n_total; }
    }
}

__global__ void kmeans_gather_kernel(const double *__restrict__ cb, int K, const int32_t *__restrict__ idx, long N,
                                     double *__restrict__ q)
{
    const long total = N * kDim;
    for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
        const long i = t / kDim;
        const int d = (int)(t - i * kDim);
        const int k = idx[i];
        q[t] = (k >= 0 && k < K) ? cb[(size_t)k * kDim + d] : 0.0;
    }
}

int num_sms();
// tensor-core screen (fpc_kmeans_tc.cu)
int run_kmeans_assign_tc(const float *d_data, long N, const double *d_cb, int K, double *acc_sums, double *acc_counts,
                         int32_t *d_idx, int R, void *d_pack, cudaStream_t st);
size_t kmeans_tc_pack_bytes(int K);
constexpr int kTcMinK = 128;      // below this the codebook is a fraction of one MMA tile: the CUDA-core kernel is as fast

static bool tc_enabled()
{
    const char *e = getenv("FPC_KMEANS_TC");        // A/B switch for measurements and tests: FPC_KMEANS_TC=0 -> CUDA-core screen
    return !(e && e[0] == '0');                    // (read on every call: a test flips it between two calls)
}

}  // namespace fpc

using namespace fpc;


// ------------------------------------------------------------------------------------------
// Column sums of the float32 data IN ROW ORDER, in float32: what np.mean(data, 0) of the reference's vq_train
// (cb_func.py:34) adds up -- NumPy reduces a C-contiguous (N, 17) array over axis 0 row by row, out[j] += a[i][j], in the
// array's own dtype.  The sum is inherently serial (17 dependent chains of N additions); one CTA does it: warps 1..7
// bring the next 256 rows into shared memory while 17 lanes of warp 0 add the current ones.  carry[17] is read at the
// start and written at the end, so the ranks of a sharded data set continue one another's sums in rank order.
// ------------------------------------------------------------------------------------------
namespace fpc {
constexpr int kSeqRows = 256;
template <typename TD>
__global__ void __launch_bounds__(256, 1) kmeans_colsum_seq_kernel(const TD *__restrict__ data, long N, TD *__restrict__ carry)
{
    __shared__ TD buf[2][kSeqRows * kDim / (sizeof(TD) / 4)];
    const int tid = threadIdx.x;
    constexpr int kRows = kSeqRows / (sizeof(TD) / 4);        // rows per chunk: the buffers stay below 48 KB
    const long nchunks = (N + kRows - 1) / kRows;
    TD acc = tid < kDim ? carry[tid] : (TD)0;
    auto load = [&](long c, int b, int t0, int nt) {
        const long first = c * kRows * (long)kDim;
        const long cnt = min((long)kRows, N - c * kRows) * kDim;
        for (long i = t0; i < cnt; i += nt) buf[b][i] = __ldg(data + first + i);
    };
    if (nchunks > 0) load(0, 0, tid, 256);
    __syncthreads();
    for (long c = 0; c < nchunks; ++c) {
        const int b = (int)(c & 1);
        if (tid >= 32) {
            if (c + 1 < nchunks) load(c + 1, b ^ 1, tid - 32, 224);
        } else if (tid < kDim) {
            const int rows = (int)min((long)kRows, N - c * kRows);
#pragma unroll 4
            for (int r = 0; r < rows; ++r) acc = Rn<TD>::add(acc, buf[b][r * kDim + tid]);
        }
        __syncthreads();
    }
    if (tid < kDim) carry[tid] = acc;
}
}  // namespace fpc

extern "C" {

// replicas of the (K,17)+(K) accumulation table: R K >= 2048 rows, none from that many entries up (the float64
// atomics of 50 M vectors on fewer rows serialise in the L2)
static int repl_rows()
{
    static int cached = 0;
    if (cached == 0) {
        const char *e = getenv("FPC_KMEANS_REPL_ROWS");     // measurement switch
        cached = e ? atoi(e) : 2048;     // measured on 50 M vectors: K = 512 16.2 -> 11.5 ms, K = 256 14.1 -> 9.7 ms against 512 rows
        if (cached < 1) cached = 2048;
    }
    return cached;
}
static int kmeans_replicas(int K) { const int rows = repl_rows(); return K >= rows ? 1 : (rows + K - 1) / K; }

static size_t replica_bytes(int K)
{
    const int R = kmeans_replicas(K);
    return R > 1 ? (((size_t)R * K * (kDim + 1) * sizeof(double) + 255) & ~(size_t)255) : 0;
}

size_t fpc_kmeans_workspace_bytes(long N, int K)
{
    (void)N;
    if (K < 1) return 0;
    // replicated accumulation tables of small codebooks + the packed operand image of the tensor-core screen
    return replica_bytes(K) + (K >= kTcMinK ? kmeans_tc_pack_bytes(K) : 0);
}

}  // extern "C"

template <typename TD>
static int kmeans_assign_any(const TD *d_data, long N, const double *d_cb, int K, double *d_sums,
                             double *d_counts, int32_t *d_idx, void *d_workspace, size_t workspace_bytes,
                             void *stream)
{
    constexpr bool kF32 = sizeof(TD) == 4;
    if (N < 0 || K < 1) return FPC_ERR_ARG;
    if (K > FPC_MAX_VQ_ENTRIES) return FPC_ERR_CODEBOOK;
    if (N == 0) return FPC_OK;
    if (!d_data || !d_cb) return FPC_ERR_ARG;
    if ((d_sums == nullptr) != (d_counts == nullptr)) return FPC_ERR_ARG;
    if (!d_sums && !d_idx) return FPC_ERR_ARG;
    const int sms = num_sms();
    if (sms <= 0) return cuda_fail(cudaErrorNoDevice);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem = (size_t)K * kKmLd64 * 8 + (size_t)((K + kKmPad - 1) / kKmPad * kKmPad) * kKmLd32 * 4;
    static bool configured[kMaxDevices] = {};
    { const int rc = ensure_dynamic_smem(kmeans_assign_kernel<TD>, (int)((size_t)FPC_MAX_VQ_ENTRIES * (kKmLd64 * 8 + kKmLd32 * 4)), configured);
      if (rc != FPC_OK) return rc; }
    long blocks = (N + (long)kKmThreads * kKmV - 1) / ((long)kKmThreads * kKmV);
    if (blocks > sms) blocks = sms;
    // with a workspace and a small codebook the sums go to replicated tables first (see the kernel)
    const bool have_ws = d_workspace && workspace_bytes >= fpc_kmeans_workspace_bytes(N, K);
    int R = d_sums ? kmeans_replicas(K) : 1;
    if (R > 1 && !have_ws) R = 1;
    double *acc_sums = d_sums, *acc_counts = d_counts;
    if (R > 1) {
        acc_sums = (double *)d_workspace;
        acc_counts = acc_sums + (size_t)R * K * kDim;
        FPC_CUDA_TRY(cudaMemsetAsync(d_workspace, 0, (size_t)R * K * (kDim + 1) * sizeof(double), st));
    }
    bool on_tc = false;
    if constexpr (kF32) {
        if (K >= kTcMinK && have_ws && tc_enabled() && (reinterpret_cast<uintptr_t>(d_data) & 15) == 0) {
            // distance screen on the tensor cores (fpc_kmeans_tc.cu); same decisions, same accumulation
            const int rc = run_kmeans_assign_tc(d_data, N, d_cb, K, acc_sums, acc_counts, d_idx, R,
                                                (char *)d_workspace + replica_bytes(K), st);
            if (rc != FPC_OK) return rc;
            on_tc = true;
        }
    }
    if (!on_tc) {       // (float64 vectors always take the CUDA-core kernel: later training stages only, a13)
        kmeans_assign_kernel<TD><<<(int)blocks, kKmThreads, smem, st>>>(d_data, N, d_cb, K, acc_sums, acc_counts, d_idx, R);
        FPC_LAUNCH_CHECK();
    }
    if (R > 1) {
        kmeans_fold_kernel<<<(K * (kDim + 1) + 255) / 256, 256, 0, st>>>(acc_sums, acc_counts, R, K, d_sums, d_counts);
        FPC_LAUNCH_CHECK();
    }
    return FPC_OK;
}

extern "C" {

int fpc_kmeans_assign_accumulate(const float *d_data, long N, const double *d_cb, int K, double *d_sums,
                                 double *d_counts, int32_t *d_idx, void *d_workspace, size_t workspace_bytes,
                                 void *stream)
{
    return kmeans_assign_any<float>(d_data, N, d_cb, K, d_sums, d_counts, d_idx, d_workspace, workspace_bytes, stream);
}

int fpc_kmeans_assign_accumulate_f64(const double *d_data, long N, const double *d_cb, int K, double *d_sums,
                                     double *d_counts, int32_t *d_idx, void *d_workspace, size_t workspace_bytes,
                                     void *stream)
{
    return kmeans_assign_any<double>(d_data, N, d_cb, K, d_sums, d_counts, d_idx, d_workspace, workspace_bytes, stream);
}

int fpc_kmeans_finalize(const double *d_sums, const double *d_counts, int K, double n_total, double *d_cb_out,
                        double *d_stats, void *stream)
{
    if (!d_sums || !d_counts || !d_cb_out || K < 1) return FPC_ERR_ARG;
    kmeans_finalize_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(const_cast<double *>(d_sums), const_cast<double *>(d_counts), K, n_total,
                                                                 d_cb_out, d_stats, 0);
    FPC_LAUNCH_CHECK();
    return FPC_OK;
}

/* fpc_kmeans_finalize on ONE accumulator  d_acc = [sums (K,17) | counts (K)]  (the layout a single all-reduce message
 * wants), which is left zeroed for the next Lloyd iteration. */
int fpc_kmeans_finalize_acc(double *d_acc, int K, double n_total, double *d_cb_out, double *d_stats, void *stream)
{
    if (!d_acc || !d_cb_out || K < 1) return FPC_ERR_ARG;
    kmeans_finalize_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(d_acc, d_acc + (size_t)K * kDim, K, n_total, d_cb_out, d_stats, 1);
    FPC_LAUNCH_CHECK();
    return FPC_OK;
}

int fpc_kmeans_colsum_f32(const float *d_data, long N, float *d_carry, void *stream)
{
    if (N < 0 || (N > 0 && d_data == nullptr) || d_carry == nullptr) return FPC_ERR_ARG;
    fpc::kmeans_colsum_seq_kernel<float><<<1, 256, 0, (cudaStream_t)stream>>>(d_data, N, d_carry);
    FPC_LAUNCH_CHECK();
    return FPC_OK;
}

int fpc_kmeans_colsum_f64(const double *d_data, long N, double *d_carry, void *stream)
{
    if (N < 0 || (N > 0 && d_data == nullptr) || d_carry == nullptr) return FPC_ERR_ARG;
    fpc::kmeans_colsum_seq_kernel<double><<<1, 256, 0, (cudaStream_t)stream>>>(d_data, N, d_carry);
    FPC_LAUNCH_CHECK();
    return FPC_OK;
}

int fpc_kmeans_gather(const double *d_cb, int K, const int32_t *d_idx, long N, double *d_q, void *stream)
{
    if (N < 0 || K < 1) return FPC_ERR_ARG;
    if (N == 0) return FPC_OK;
    if (!d_cb || !d_idx || !d_q) return FPC_ERR_ARG;
    long blocks = (N * kDim + 255) / 256;
    const int cap = 8 * (num_sms() > 0 ? num_sms() : 148);
    if (blocks > cap) blocks = cap;
    kmeans_gather_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(d_cb, K, d_idx, N, d_q);
    FPC_LAUNCH_CHECK();
    return FPC_OK;
}

}  // extern "C"
