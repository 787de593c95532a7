// fpc_train.cu -- the data formats either side of the closed loop.
//
//   fpc_compact_rows          train_cb.py:177-187  training sets from the encoder's qtz=False outputs:
//                                                  scalar set  [k for k in r[:,:,0].flatten() if k != 0]
//                                                  vector set  [row for row in r[:,:,-17:] if sum(abs(row)) != 0]
//                                                  (order preserving, on the device: k-means then consumes the
//                                                  residuals where the encoder left them)
//   fpc_kmeans_stage_residual train_cb.py:199-200,210-211  r = quantize(codebook, r) - r  (the flipped sign is the
//                                                  reference's) for the next stage
//   fpc_dequantize            models/wavernn.py:217-240 read backwards: index record -> r_qtz, the receiver's half of
//                                                  the quantisers (the reference has no working decoder, :367-379)
//   fpc_pack_frames / fpc_unpack_frames            a fixed 32-bit word per frame as the wire format of the index
//                                                  record (the reference defines none)
// All of it is byte/index work bound by HBM: one pass over the rows, coalesced, grid sized to the SM count.
#include "fpc_common.cuh"
#include "fpc_vq.cuh"

namespace fpc {

int num_sms();

constexpr int kCpThreads = 256;
constexpr int kCpRows = 1024;          // rows per CTA tile (4 per thread)

// keep[row] = (sum_j |x[row][col0 + j]| != 0), written as the reference writes it: NaN rows are kept
__device__ __forceinline__ bool row_kept(const float *__restrict__ row, int ncols)
{
    float s = 0.0f;
    for (int j = 0; j < ncols; ++j) s += fabsf(row[j]);
    return !(s == 0.0f);
}

// pass 1: kept rows per tile
__global__ void compact_count_kernel(const float *__restrict__ src, long n, int stride, int col0, int ncols,
                                     unsigned int *__restrict__ tile_counts)
{
    __shared__ unsigned int s_cnt;
    for (long tile = blockIdx.x; tile * kCpRows < n; tile += gridDim.x) {
        if (threadIdx.x == 0) s_cnt = 0;
        __syncthreads();
        unsigned int c = 0;
        for (int q = 0; q < kCpRows / kCpThreads; ++q) {
            const long i = tile * kCpRows + q * kCpThreads + threadIdx.x;
            if (i < n && row_kept(src + i * stride + col0, ncols)) ++c;
        }
        for (int off = 16; off > 0; off >>= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
        if ((threadIdx.x & 31) == 0 && c) atomicAdd(&s_cnt, c);
        __syncthreads();
        if (threadIdx.x == 0) tile_counts[tile] = s_cnt;
        __syncthreads();
    }
}

// pass 2: exclusive scan of the tile counts (one CTA; a few thousand tiles at most), total -> *count
__global__ void compact_scan_kernel(unsigned int *__restrict__ tile_counts, long ntiles, long long *__restrict__ count)
{
    __shared__ unsigned long long s_warp[32];
    __shared__ unsigned long long s_base;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (long base = 0; base < ntiles; base += blockDim.x) {
        const long t = base + threadIdx.x;
        const unsigned long long v = t < ntiles ? tile_counts[t] : 0;
        unsigned long long x = v;
        for (int off = 1; off < 32; off <<= 1) {
            const unsigned long long y = __shfl_up_sync(0xffffffffu, x, off);
            if (lane >= off) x += y;
        }
        if (lane == 31) s_warp[warp] = x;
        __syncthreads();
        if (warp == 0) {
            unsigned long long w = lane < (int)(blockDim.x >> 5) ? s_warp[lane] : 0;
            for (int off = 1; off < 32; off <<= 1) {
                const unsigned long long y = __shfl_up_sync(0xffffffffu, w, off);
                if (lane >= off) w += y;
            }
            s_warp[lane] = w;            // inclusive scan of the warp totals
        }
        __syncthreads();
        const unsigned long long before = s_base + (warp ? s_warp[warp - 1] : 0) + (x - v);
        // (fpc_compact_rows refuses 2^32 rows or more, so the offsets fit the 32-bit slots they overwrite)
        if (t < ntiles) tile_counts[t] = (unsigned int)before;
        __syncthreads();
        if (threadIdx.x == 0) s_base += s_warp[(blockDim.x >> 5) - 1];
        __syncthreads();
    }
    if (threadIdx.x == 0) *count = (long long)s_base;
}

// pass 3: order-preserving scatter.  A warp handles 32 consecutive rows: ballot gives each kept row its slot.
__global__ void compact_scatter_kernel(const float *__restrict__ src, long n, int stride, int col0, int ncols,
                                       const unsigned int *__restrict__ tile_offsets, float *__restrict__ dst)
{
    __shared__ unsigned int s_warp[kCpThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (long tile = blockIdx.x; tile * kCpRows < n; tile += gridDim.x) {
        unsigned int running = tile_offsets[tile];
        for (int q = 0; q < kCpRows / kCpThreads; ++q) {
            const long i = tile * kCpRows + q * kCpThreads + threadIdx.x;
            const bool keep = i < n && row_kept(src + i * stride + col0, ncols);
            const unsigned int bal = __ballot_sync(0xffffffffu, keep);
            if (lane == 0) s_warp[warp] = __popc(bal);
            __syncthreads();
            unsigned int before = running;
            for (int w = 0; w < warp; ++w) before += s_warp[w];
            unsigned int total = 0;
            for (int w = 0; w < kCpThreads / 32; ++w) total += s_warp[w];
            if (keep) {
                const size_t slot = (size_t)before + __popc(bal & ((1u << lane) - 1u));
                const float *row = src + i * stride + col0;
                for (int j = 0; j < ncols; ++j) dst[slot * ncols + j] = row[j];
            }
            running += total;
            __syncthreads();
        }
    }
}

// next[i] = (float)(cb[idx[i]] - (double)data[i])   (float64 codebook minus float32 data promotes, then the k-means
// input is float32 again)
// (TO = double keeps the difference as the reference does, train_cb.py:200: float64 codebook minus float32 or float64 data)
template <typename TD, typename TO>
__global__ void stage_residual_kernel(const double *__restrict__ cb, int K, const int32_t *__restrict__ idx,
                                      const TD *__restrict__ data, long N, TO *__restrict__ next)
{
    const long total = N * kDim;
    for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
        const long i = t / kDim;
        const int d = (int)(t - i * kDim);
        const int k = idx[i];
        const double q = (k >= 0 && k < K) ? cb[(size_t)k * kDim + d] : 0.0;
        next[t] = (TO)__dsub_rn(q, (double)data[t]);
    }
}

// ---- receiver side of the quantisers ----
template <typename T>
__device__ __forceinline__ float dequant_vq(const char *__restrict__ cbbase, const PackedVq &bk, int i1, int i2, int d)
{
    if (bk.stages < 1 || i1 < 0 || i1 >= bk.K) return 0.0f;
    const T *r0 = reinterpret_cast<const T *>(cbbase + bk.off_r[0]);
    T csum = Rn<T>::add((T)0, r0[(size_t)i1 * kDim + d]);            // csum = 0; csum += CB[i][index]  (vq_func.py:127-129)
    if (bk.stages == 2 && i2 >= 0 && i2 < bk.K) {
        const T *r1 = reinterpret_cast<const T *>(cbbase + bk.off_r[1]);
        csum = Rn<T>::add(csum, r1[(size_t)i2 * kDim + d]);
    }
    return (float)csum;
}

__device__ __forceinline__ float dequant_scl(const char *__restrict__ cbbase, const PackedScl &sb, int i)
{
    if (sb.n <= 0 || i < 0 || i >= sb.n) return 0.0f;
    if (sb.dtype == FPC_F32) return reinterpret_cast<const float *>(cbbase + sb.off)[i];
    return (float)reinterpret_cast<const double *>(cbbase + sb.off)[i];
}

__global__ void dequantize_kernel(const char *__restrict__ cb, const int4 *__restrict__ idx, long n, float *__restrict__ rq)
{
    const PackedCodebooks *h = reinterpret_cast<const PackedCodebooks *>(cb);
    const long total = n * kFc;
    for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
        const long i = t / kFc;
        const int j = (int)(t - i * kFc);
        const int4 v = idx[i];
        float out;
        if (j == 0) {
            out = dequant_scl(cb, (v.w & 1) ? h->scl : h->blscl, v.x);
        } else {
            const PackedVq &bk = (v.w & 2) ? h->vq : h->bl;
            out = bk.dtype == FPC_F32 ? dequant_vq<float>(cb, bk, v.y, v.z, j - 1) : dequant_vq<double>(cb, bk, v.y, v.z, j - 1);
        }
        rq[t] = out;
    }
}

// ---- wire format: one 32-bit word per frame ----
//   bit 0 ind1, bit 1 ind2, bits 2-9 scalar index, bits 10-19 VQ index (stage 1 or the below book), bits 20-29 VQ
//   stage-2 index.  "Nothing coded" (-1 in the record: a below-threshold frame without below-threshold books) needs no
//   bits: which books exist is configuration both sides share, so the receiver restores the -1 from the codebook set.
__global__ void pack_frames_kernel(const int4 *__restrict__ idx, long n, uint32_t *__restrict__ words)
{
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const int4 v = idx[i];
        const uint32_t s = v.x < 0 ? 0u : (uint32_t)v.x & 0xffu;
        const uint32_t a = v.y < 0 ? 0u : (uint32_t)v.y & 0x3ffu;
        const uint32_t b = v.z < 0 ? 0u : (uint32_t)v.z & 0x3ffu;
        words[i] = ((uint32_t)v.w & 3u) | (s << 2) | (a << 10) | (b << 20);
    }
}

__global__ void unpack_frames_kernel(const char *__restrict__ cb, const uint32_t *__restrict__ words, long n,
                                     int4 *__restrict__ idx)
{
    const PackedCodebooks *h = reinterpret_cast<const PackedCodebooks *>(cb);
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const uint32_t w = words[i];
        int4 v;
        v.w = (int)(w & 3u);
        const PackedScl &sb = (v.w & 1) ? h->scl : h->blscl;
        const PackedVq &bk = (v.w & 2) ? h->vq : h->bl;
        v.x = sb.n > 0 ? (int)((w >> 2) & 0xffu) : -1;
        v.y = bk.stages >= 1 ? (int)((w >> 10) & 0x3ffu) : -1;
        v.z = bk.stages == 2 ? (int)((w >> 20) & 0x3ffu) : -1;
        idx[i] = v;
    }
}

static int grid_for(long items, int threads)
{
    const int sms = num_sms();
    long blocks = (items + threads - 1) / threads;
    const long cap = (long)(sms > 0 ? sms : 1) * 8;
    if (blocks > cap) blocks = cap;
    return (int)(blocks < 1 ? 1 : blocks);
}

}  // namespace fpc

using namespace fpc;

extern "C" {

size_t fpc_compact_workspace_bytes(long n_rows)
{
    if (n_rows <= 0) return 0;
    return (size_t)((n_rows + kCpRows - 1) / kCpRows) * sizeof(unsigned int);
}

int fpc_compact_rows(const float *d_src, long n_rows, int src_stride, int col0, int ncols, float *d_dst,
                     long long *d_count, void *d_workspace, size_t workspace_bytes, void *stream)
{
    if (n_rows < 0 || ncols < 1 || col0 < 0 || src_stride < col0 + ncols) return FPC_ERR_ARG;
    if (!d_count) return FPC_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    if (n_rows == 0) {
        FPC_CUDA_TRY(cudaMemsetAsync(d_count, 0, sizeof(long long), st));
        return FPC_OK;
    }
    if (!d_src || !d_dst) return FPC_ERR_ARG;
    if (n_rows >= (1LL << 32)) return FPC_ERR_SHAPE;         // tile offsets are 32-bit
    if (!d_workspace || workspace_bytes < fpc_compact_workspace_bytes(n_rows)) return FPC_ERR_WORKSPACE;
    unsigned int *tiles = (unsigned int *)d_workspace;
    const long ntiles = (n_rows + kCpRows - 1) / kCpRows;
    const int sms = num_sms();
    if (sms <= 0) return cuda_fail(cudaErrorNoDevice);
    const int grid = (int)(ntiles < (long)sms * 8 ? ntiles : (long)sms * 8);
    compact_count_kernel<<<grid, kCpThreads, 0, st>>>(d_src, n_rows, src_stride, col0, ncols, tiles);
    FPC_LAUNCH_CHECK();
    compact_scan_kernel<<<1, 1024, 0, st>>>(tiles, ntiles, d_count);
    FPC_LAUNCH_CHECK();
    compact_scatter_kernel<<<grid, kCpThreads, 0, st>>>(d_src, n_rows, src_stride, col0, ncols, tiles, d_dst);
    FPC_LAUNCH_CHECK();
    return FPC_OK;
}

int fpc_kmeans_stage_residual(const double *d_cb, int K, const int32_t *d_idx, const float *d_data, long N,
                              float *d_next, void *stream)
{
    if (N < 0 || K < 1) return FPC_ERR_ARG;
    if (N == 0) return FPC_OK;
    if (!d_cb || !d_idx || !d_data || !d_next) return FPC_ERR_ARG;
    stage_residual_kernel<float, float><<<grid_for(N * kDim, 256), 256, 0, (cudaStream_t)stream>>>(d_cb, K, d_idx, d_data, N, d_next);
    FPC_LAUNCH_CHECK();
    return FPC_OK;
}

int fpc_kmeans_stage_residual_f64(const double *d_cb, int K, const int32_t *d_idx, const void *d_data, int data_is_f64, long N,
                                  double *d_next, void *stream)
{
    if (N < 0 || K < 1) return FPC_ERR_ARG;
    if (N == 0) return FPC_OK;
    if (!d_cb || !d_idx || !d_data || !d_next) return FPC_ERR_ARG;
    if (data_is_f64)
        stage_residual_kernel<double, double><<<grid_for(N * kDim, 256), 256, 0, (cudaStream_t)stream>>>(d_cb, K, d_idx, (const double *)d_data, N, d_next);
    else
        stage_residual_kernel<float, double><<<grid_for(N * kDim, 256), 256, 0, (cudaStream_t)stream>>>(d_cb, K, d_idx, (const float *)d_data, N, d_next);
    FPC_LAUNCH_CHECK();
    return FPC_OK;
}

int fpc_dequantize(const void *d_packed_codebooks, const int32_t *d_idx, long n_frames, float *d_r_qtz, void *stream)
{
    if (n_frames < 0) return FPC_ERR_ARG;
    if (n_frames == 0) return FPC_OK;
    if (!d_packed_codebooks || !d_idx || !d_r_qtz) return FPC_ERR_ARG;
    dequantize_kernel<<<grid_for(n_frames * kFc, 256), 256, 0, (cudaStream_t)stream>>>(
        (const char *)d_packed_codebooks, (const int4 *)d_idx, n_frames, d_r_qtz);
    FPC_LAUNCH_CHECK();
    return FPC_OK;
}

int fpc_pack_frames(const int32_t *d_idx, long n_frames, uint32_t *d_words, void *stream)
{
    if (n_frames < 0) return FPC_ERR_ARG;
    if (n_frames == 0) return FPC_OK;
    if (!d_idx || !d_words) return FPC_ERR_ARG;
    pack_frames_kernel<<<grid_for(n_frames, 256), 256, 0, (cudaStream_t)stream>>>((const int4 *)d_idx, n_frames, d_words);
    FPC_LAUNCH_CHECK();
    return FPC_OK;
}

int fpc_unpack_frames(const void *d_packed_codebooks, const uint32_t *d_words, long n_frames, int32_t *d_idx, void *stream)
{
    if (n_frames < 0) return FPC_ERR_ARG;
    if (n_frames == 0) return FPC_OK;
    if (!d_packed_codebooks || !d_words || !d_idx) return FPC_ERR_ARG;
    unpack_frames_kernel<<<grid_for(n_frames, 256), 256, 0, (cudaStream_t)stream>>>(
        (const char *)d_packed_codebooks, d_words, n_frames, (int4 *)d_idx);
    FPC_LAUNCH_CHECK();
    return FPC_OK;
}

}  // extern "C"
