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
//     s_k = ||c_k||^2 - 2 <x, c_k>             (8 packed FMAs + 1 FMA per pair; d_k = s_k + ||x||^2)
// on the fp32 shadow (float)c_k.  With u = 2^-24 and R = (||x|| + ||c_k||)^2 the 18 FMA roundings, the rounding of
// the shadow and of the stored norm give |s_k - (d_k - ||x||^2)| <= 21 u R (numpy's own float64 evaluation of d_k
// is exact at this scale), so two screened values of one vector compare with error <= 42 u R.  One divergence-free
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
#include "fpc_common.cuh"
#include "fpc_vq.cuh"

namespace fpc {

constexpr int kKmThreads = 256;
constexpr int kKmV = 4;        // vectors per thread: the 5 broadcast LDS of a centroid row serve 4 screens (and 4-way ILP)
constexpr int kKmLd64 = 18;   // float64 codeword row stride in shared memory (16-byte aligned rows)
constexpr int kKmLd32 = 20;   // float32 shadow row stride (float4 aligned)

// centroid row in registers: 16 floats, then (c16, ||c||^2)
struct KmRow { float4 c0, c1, c2, c3; float2 tail; };
__device__ __forceinline__ KmRow load_row(const float *__restrict__ c)
{
    KmRow r;
    r.c0 = *reinterpret_cast<const float4 *>(c);
    r.c1 = *reinterpret_cast<const float4 *>(c + 4);
    r.c2 = *reinterpret_cast<const float4 *>(c + 8);
    r.c3 = *reinterpret_cast<const float4 *>(c + 12);
    r.tail = *reinterpret_cast<const float2 *>(c + 16);
    return r;
}
// s = cn + sum_d xm_d * c_d  (xm = -2 x), packed over (d, d+1) pairs
__device__ __forceinline__ float screen17(const float2 (&xm)[8], float xm16, const KmRow &row)
{
    const float4 c0 = row.c0, c1 = row.c1, c2 = row.c2, c3 = row.c3;
    const float2 tail = row.tail;
    float2 acc = make_float2(tail.y, 0.0f);
    acc = fma2(make_float2(c0.x, c0.y), xm[0], acc);
    acc = fma2(make_float2(c0.z, c0.w), xm[1], acc);
    acc = fma2(make_float2(c1.x, c1.y), xm[2], acc);
    acc = fma2(make_float2(c1.z, c1.w), xm[3], acc);
    acc = fma2(make_float2(c2.x, c2.y), xm[4], acc);
    acc = fma2(make_float2(c2.z, c2.w), xm[5], acc);
    acc = fma2(make_float2(c3.x, c3.y), xm[6], acc);
    acc = fma2(make_float2(c3.z, c3.w), xm[7], acc);
    return __fmaf_rn(tail.x, xm16, acc.x) + acc.y;
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
    for (int k = threadIdx.x; k < K; k += kKmThreads) {          // ||c_k||^2 in float64, rounded, next to c16
        double n2 = 0.0;
#pragma unroll
        for (int d = 0; d < kDim; ++d) n2 += cb64[k * kKmLd64 + d] * cb64[k * kKmLd64 + d];
        cb32[k * kKmLd32 + 17] = (float)n2;
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
        float2 xm[kKmV][8];
        float xm16[kKmV], slack[kKmV], m1[kKmV], m2[kKmV];
        int bi[kKmV];
#pragma unroll
        for (int j = 0; j < kKmV; ++j) {
            const long i = i0 + (long)j * kKmThreads;
            float x[kDim];
#pragma unroll
            for (int d = 0; d < kDim; ++d) x[d] = i < N ? __ldg(data + i * kDim + d) : 0.0f;
            // per-vector screen constants: xm = -2 x (exact), ||x||^2 and the slack A = 128 u (||x|| + Cmax)^2, rounded up
#pragma unroll
            for (int d = 0; d < 8; ++d) xm[j][d] = make_float2(-2.0f * x[2 * d], -2.0f * x[2 * d + 1]);
            xm16[j] = -2.0f * x[16];
            float nx = 0.0f;
#pragma unroll
            for (int d = 0; d < kDim; ++d) nx = __fmaf_ru(x[d], x[d], nx);
            const float rr = __fadd_ru(__fsqrt_ru(nx), cmax);
            slack[j] = __fadd_ru(__fmul_ru(__fmul_ru(rr, rr), 7.62939453125e-6f), 1e-30f);   // 128 * 2^-24
            m1[j] = __int_as_float(0x7f800000); m2[j] = m1[j]; bi[j] = 0;
        }
        // One divergence-free sweep keeps the two smallest screened values of every vector.  All s_k of a vector
        // share the ||x||^2 term, so two of them compare with error <= 2 * 21 u R: if the runner-up is more than
        // the slack above the smallest, the smallest is the exact argmin (no float64 work at all); otherwise the
        // vector is ambiguous (near-tie, duplicate centroids) and the candidates inside the slack are decided in
        // float64 exactly as numpy does, ascending k with strict <, i.e. first minimum.
        for (int k = 0; k < K; ++k) {
            const KmRow row = load_row(cb32 + k * kKmLd32);
#pragma unroll
            for (int j = 0; j < kKmV; ++j) {
                const float s = screen17(xm[j], xm16[j], row);
                const bool p = s < m1[j];
                m2[j] = fminf(m2[j], fmaxf(m1[j], s));
                m1[j] = fminf(m1[j], s);
                bi[j] = p ? k : bi[j];
            }
        }
#pragma unroll
        for (int j = 0; j < kKmV; ++j) {
            const long i = i0 + (long)j * kKmThreads;
            const bool valid = i < N;
            const float thr = __fadd_ru(m1[j], slack[j]);
            float x[kDim];
            const bool need_x = (sums != nullptr) || !(m2[j] > thr);
            if (need_x) {
#pragma unroll
                for (int d = 0; d < kDim; ++d) x[d] = valid ? __ldg(data + i * kDim + d) : 0.0f;   // L1/L2 hit
            }
            if (!(m2[j] > thr)) {
                double best = 0.0;
                bool have = false;
                for (int k = 0; k < K; ++k) {
                    const float s = screen17(xm[j], xm16[j], load_row(cb32 + k * kKmLd32));
                    if (s <= thr) {
                        const double d = exact17(x, cb64 + k * kKmLd64);
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
                        for (int d = 0; d < kDim; ++d) atomicAdd(sums + (size_t)b * kDim + d, acc[d]);
                        atomicAdd(counts + b, (double)__popc(peers));
                    }
                } else if (valid) {
#pragma unroll
                    for (int d = 0; d < kDim; ++d) atomicAdd(sums + (size_t)b * kDim + d, (double)x[d]);
                    atomicAdd(counts + b, 1.0);
                }
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
    long blocks = (N + (long)kKmThreads * kKmV - 1) / ((long)kKmThreads * kKmV);
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
