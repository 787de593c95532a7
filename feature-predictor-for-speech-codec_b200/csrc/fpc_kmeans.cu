// fpc_kmeans.cu -- Lloyd iteration of the codebook learner.
//
// Reference: /root/reference/src/quantization/cb_func.py
//   find_nearest (:56-68)   dist = np.sum((data - codebook) ** 2, -1) in float64 (float32 data
//                           minus float64 codebook promotes), np.argmin -> first minimum
//   update       (:71-100)  per-centroid float64 sums and counts, codebook = sum / (count + 1e-20)
//   quantize     (:103-112) nearest-centroid gather
//
// Exactness.  The assignment is decided by EXACT float64 direct-form distances evaluated with
// numpy's roundings (dist17<double>, fpc_vq.cuh), so indices are bit-identical to the
// reference's, ties included.  To avoid paying 51 FP64 operations for every (vector, centroid)
// pair, each pair is first screened in fp32: s = sum_d fma(t, t, .) with t = x_d - (float)c_d.
// With delta_d = |t32_d - t64_d| <= 2^-24 (|c_d| + |t_d|) and 2ab <= g a^2 + b^2/g (g = 2^-12):
//     |s - d64| <= (2^-12 + 2^-19) d64 + 2^-34 max_k ||c_k||^2          
// so with REL = 2^-10 and ABS = 2^-32 Cmax^2 + 1e-30 (both carry a >= 2x safety factor, which
// also absorbs rounding `best` to fp32), s > best (1 + REL) + ABS proves d64 > best: that
// centroid cannot win (a later index also loses ties) and is skipped.  Every survivor is
// re-evaluated in float64.  The screen only removes provable losers; it never picks a winner.
//
// Sums.  The reference accumulates in data order on one thread.  Here every CTA first reduces
// its vectors into per-warp-aggregated float64 atomics on the global (K,17) table; the order
// of additions therefore differs and centroids agree with the reference to ~1e-15 relative
// (tests state 1e-12).  With several ranks the caller all-reduces sums and counts (NCCL)
// between fpc_kmeans_assign_accumulate and fpc_kmeans_finalize.
#include "fpc_common.cuh"
#include "fpc_vq.cuh"

namespace fpc {

constexpr int kKmThreads = 512;
constexpr int kKmLd64 = 18;   // float64 codeword row stride in shared memory (16-byte aligned rows)
constexpr int kKmLd32 = 20;   // float32 shadow row stride (float4 aligned)

#define FPC_KM_SCREEN_REL 9.765625e-04f            /* 2^-10 */
#define FPC_KM_SCREEN_ABS_SCALE 2.3283064365386963e-10f /* 2^-32 */

__device__ __forceinline__ float screen_threshold(double best, float abs_slack)
{
    // round best UP to fp32 so the threshold never undershoots
    const float bf = __double2float_ru(best);
    return __fadd_ru(__fmaf_ru(bf, FPC_KM_SCREEN_REL, bf), abs_slack);
}

__device__ __forceinline__ float screen17(const float (&x)[kDim], const float *__restrict__ c)
{
    // packed lanes: (d, d+1) pairs through add.rn.f32x2 / fma.rn.f32x2 -- the same per-lane IEEE operations as the
    // scalar form (t = x - c, acc = fma(t, t, acc)), half the instructions
    const float4 c0 = *reinterpret_cast<const float4 *>(c);
    const float4 c1 = *reinterpret_cast<const float4 *>(c + 4);
    const float4 c2 = *reinterpret_cast<const float4 *>(c + 8);
    const float4 c3 = *reinterpret_cast<const float4 *>(c + 12);
    const float c16 = c[16];
    float2 acc = make_float2(0.0f, 0.0f), t;
    t = sub2(make_float2(x[0], x[1]), make_float2(c0.x, c0.y)); acc = fma2(t, t, acc);
    t = sub2(make_float2(x[2], x[3]), make_float2(c0.z, c0.w)); acc = fma2(t, t, acc);
    t = sub2(make_float2(x[4], x[5]), make_float2(c1.x, c1.y)); acc = fma2(t, t, acc);
    t = sub2(make_float2(x[6], x[7]), make_float2(c1.z, c1.w)); acc = fma2(t, t, acc);
    t = sub2(make_float2(x[8], x[9]), make_float2(c2.x, c2.y)); acc = fma2(t, t, acc);
    t = sub2(make_float2(x[10], x[11]), make_float2(c2.z, c2.w)); acc = fma2(t, t, acc);
    t = sub2(make_float2(x[12], x[13]), make_float2(c3.x, c3.y)); acc = fma2(t, t, acc);
    t = sub2(make_float2(x[14], x[15]), make_float2(c3.z, c3.w)); acc = fma2(t, t, acc);
    const float tl = x[16] - c16;
    return __fmaf_rn(tl, tl, acc.x) + acc.y;
}

__device__ __forceinline__ double exact17(const float (&x)[kDim], const double *__restrict__ c)
{
    double xd[kDim], cd[kDim];
#pragma unroll
    for (int d = 0; d < kDim; ++d) { xd[d] = (double)x[d]; cd[d] = c[d]; }
    return dist17<double>(xd, cd);
}

// one vector per thread; codebook (float64 + fp32 shadow) resident in shared memory
__global__ void __launch_bounds__(kKmThreads, 1)
kmeans_assign_kernel(const float *__restrict__ data, long N, const double *__restrict__ cb, int K,
                     double *__restrict__ sums, double *__restrict__ counts, int32_t *__restrict__ idx_out)
{
    extern __shared__ __align__(16) unsigned char smem[];
    double *cb64 = reinterpret_cast<double *>(smem);                              // [K][18]
    float *cb32 = reinterpret_cast<float *>(smem + (size_t)K * kKmLd64 * 8);      // [K][20]
    for (int i = threadIdx.x; i < K * kDim; i += kKmThreads) {
        const int k = i / kDim, d = i - k * kDim;
        const double v = cb[i];
        cb64[k * kKmLd64 + d] = v;
        cb32[k * kKmLd32 + d] = (float)v;
    }
    __syncthreads();
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
    float abs_slack = 0.0f;
    for (int w = 0; w < kKmThreads / 32; ++w) abs_slack = fmaxf(abs_slack, s_c2[w]);
    abs_slack = __fadd_ru(__fmul_ru(abs_slack, FPC_KM_SCREEN_ABS_SCALE), 1e-30f);

    const int lane = threadIdx.x & 31;
    const long stride = (long)gridDim.x * kKmThreads;
    const long nround = (N + stride - 1) / stride;
    for (long it = 0; it < nround; ++it) {
        const long i = it * stride + (long)blockIdx.x * kKmThreads + threadIdx.x;
        const bool valid = i < N;
        float x[kDim];
#pragma unroll
        for (int d = 0; d < kDim; ++d) x[d] = valid ? __ldg(data + i * kDim + d) : 0.0f;

        // exact distance to centroid 0 seeds the running minimum (np.argmin: first minimum)
        double best = exact17(x, cb64);
        int bi = 0;
        float thr = screen_threshold(best, abs_slack);
        for (int k = 1; k < K; ++k) {
            const float s = screen17(x, cb32 + k * kKmLd32);
            if (s <= thr) {   // cannot be excluded: decide in float64 exactly as numpy does
                const double d = exact17(x, cb64 + k * kKmLd64);
                if (d < best) {
                    best = d; bi = k;
                    thr = screen_threshold(best, abs_slack);
                }
            }
        }
        if (valid && idx_out) idx_out[i] = bi;

        // accumulate.  Small codebooks (the first steps of the grow-by-one schedule,
        // cb_func.py:34-47) would hammer a handful of addresses, so there the lanes that chose
        // the same centroid are merged first (uniform full-mask shuffles, ascending lane order).
        if (sums) {
            const int key = valid ? bi : -1;
            if (K <= 32) {
                const unsigned peers = __match_any_sync(0xffffffffu, key);
                const int leader = __ffs(peers) - 1;
                double acc[kDim];
#pragma unroll
                for (int d = 0; d < kDim; ++d) acc[d] = 0.0;
                for (int src = 0; src < 32; ++src) {
                    const bool take = (peers >> src) & 1u;
#pragma unroll
                    for (int d = 0; d < kDim; ++d) {
                        const double v = __shfl_sync(0xffffffffu, (double)x[d], src);
                        if (take) acc[d] += v;
                    }
                }
                if (valid && lane == leader) {
#pragma unroll
                    for (int d = 0; d < kDim; ++d) atomicAdd(sums + (size_t)bi * kDim + d, acc[d]);
                    atomicAdd(counts + bi, (double)__popc(peers));
                }
            } else if (valid) {
#pragma unroll
                for (int d = 0; d < kDim; ++d) atomicAdd(sums + (size_t)bi * kDim + d, (double)x[d]);
                atomicAdd(counts + bi, 1.0);
            }
        }
    }
}

// codebook = sums / (counts + 1e-20); statistics of cb_func.py:92-97
__global__ void kmeans_finalize_kernel(const double *__restrict__ sums, const double *__restrict__ counts, int K,
                                       double n_total, double *__restrict__ cb_out, double *__restrict__ stats)
{
    __shared__ double s_min[32], s_max[32], s_empty[32], s_w2[32];
    double mn = 1e300, mx = -1e300, em = 0.0, w2 = 0.0;
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        const double c = counts[k];
        const double den = c + 1e-20;
#pragma unroll
        for (int d = 0; d < kDim; ++d) cb_out[(size_t)k * kDim + d] = sums[(size_t)k * kDim + d] / den;
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
        if (lane == 0 && stats) { stats[0] = mn; stats[1] = mx; stats[2] = em; stats[3] = w2; }
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

}  // namespace fpc

using namespace fpc;

extern "C" {

size_t fpc_kmeans_workspace_bytes(long N, int K)
{
    (void)N; (void)K;
    return 0;
}

int fpc_kmeans_assign_accumulate(const float *d_data, long N, const double *d_cb, int K, double *d_sums,
                                 double *d_counts, int32_t *d_idx, void *d_workspace, size_t workspace_bytes,
                                 void *stream)
{
    (void)d_workspace; (void)workspace_bytes;
    if (N < 0 || K < 1) return FPC_ERR_ARG;
    if (K > FPC_MAX_VQ_ENTRIES) return FPC_ERR_CODEBOOK;
    if (N == 0) return FPC_OK;
    if (!d_data || !d_cb) return FPC_ERR_ARG;
    if ((d_sums == nullptr) != (d_counts == nullptr)) return FPC_ERR_ARG;
    if (!d_sums && !d_idx) return FPC_ERR_ARG;
    const int sms = num_sms();
    if (sms <= 0) return cuda_fail(cudaErrorNoDevice);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem = (size_t)K * (kKmLd64 * 8 + kKmLd32 * 4);
    static size_t configured = 0;
    if (smem > configured) {
        FPC_CUDA_TRY(cudaFuncSetAttribute(kmeans_assign_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)((size_t)FPC_MAX_VQ_ENTRIES * (kKmLd64 * 8 + kKmLd32 * 4))));
        configured = (size_t)FPC_MAX_VQ_ENTRIES * (kKmLd64 * 8 + kKmLd32 * 4);
    }
    long blocks = (N + kKmThreads - 1) / kKmThreads;
    if (blocks > sms) blocks = sms;
    kmeans_assign_kernel<<<(int)blocks, kKmThreads, smem, st>>>(d_data, N, d_cb, K, d_sums, d_counts, d_idx);
    FPC_LAUNCH_CHECK();
    return FPC_OK;
}

int fpc_kmeans_finalize(const double *d_sums, const double *d_counts, int K, double n_total, double *d_cb_out,
                        double *d_stats, void *stream)
{
    if (!d_sums || !d_counts || !d_cb_out || K < 1) return FPC_ERR_ARG;
    kmeans_finalize_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(d_sums, d_counts, K, n_total, d_cb_out, d_stats);
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
