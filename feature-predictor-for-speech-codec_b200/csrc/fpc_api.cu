// fpc_api.cu -- C-ABI entry points of the encode / decode / quantiser paths (include/fpc_b200.h).
#include <mutex>
#include <vector>
#include "fpc_common.cuh"
#include "fpc_vq.cuh"
#include "fpc_vq_search.cuh"
#include "fpc_vq_screen.cuh"
#include "fpc_encode.cuh"

namespace fpc {

// ------------------------------------------------------------------------------------------
// stand-alone VQ quantiser (vq_func.py:134-164): tiles of 32 vectors per CTA iteration
// ------------------------------------------------------------------------------------------
constexpr int kQTile = 32;

template <typename T>
__global__ void __launch_bounds__(kComputeThreads, 1)
vq_quantize_kernel(const float *__restrict__ x, long n, const char *__restrict__ cb, int which, T *__restrict__ q,
                   int32_t *__restrict__ idx, int scratch_bytes)
{
    extern __shared__ __align__(128) unsigned char smem[];
    float *rs = reinterpret_cast<float *>(smem);                 // [32][24]
    float *rq = rs + kQTile * kLdR;                              // [32][20]
    int *idx1 = reinterpret_cast<int *>(rq + kQTile * 20);       // [32]
    int *idx2 = idx1 + kQTile;
    int *list = idx2 + kQTile;
    char *scratch = reinterpret_cast<char *>(list + kQTile);
    const PackedCodebooks *h = reinterpret_cast<const PackedCodebooks *>(cb);
    const PackedVq &bk = which == 0 ? h->vq : h->bl;
    const int tid = threadIdx.x;
    const long ntiles = (n + kQTile - 1) / kQTile;
    for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long base = tile * kQTile;
        const int nb = (int)min((long)kQTile, n - base);
        for (int i = tid; i < kQTile * kDim; i += kComputeThreads) {
            const int r = i / kDim, d = i - r * kDim;
            rs[r * kLdR + 4 + d] = r < nb ? x[(base + r) * kDim + d] : 0.0f;
        }
        if (tid < kQTile) list[tid] = tid;
        __syncthreads();
        vq_search_rows_screened<T>(bk, cb, list, nb, kQTile, rs, rq, idx1, idx2, scratch, scratch_bytes, tid, q + base * kDim);
        if (tid < nb) {
            idx[(base + tid) * bk.stages] = idx1[tid];
            if (bk.stages > 1) idx[(base + tid) * bk.stages + 1] = idx2[tid];
        }
        __syncthreads();
    }
}

template <typename T>
__global__ void scl_quantize_kernel(const float *__restrict__ x, long n, const T *__restrict__ codes, int n_code,
                                    T *__restrict__ q, int32_t *__restrict__ idx)
{
    const int lane = threadIdx.x & 31;
    const long warp = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
    for (long i = warp; i < n; i += nwarps) {
        T qv;
        const int bi = warp_scl_nearest<T>(codes, n_code, x[i], lane, qv);
        if (lane == 0) { q[i] = qv; idx[i] = bi; }
    }
}

// cb_tot (wavernn.py:189,221-240) from the index record
__global__ void index_histogram_kernel(const int4 *__restrict__ idx, long n, unsigned long long *__restrict__ hist)
{
    __shared__ unsigned int sh[FPC_HIST_TOTAL];
    for (int i = threadIdx.x; i < FPC_HIST_TOTAL; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const int4 v = idx[i];
        if (v.x >= 0 && v.x < 256) atomicAdd(&sh[((v.w & 1) ? FPC_HIST_SCL : FPC_HIST_BL_SCL) + v.x], 1u);
        if (v.w & 2) {        // above threshold: cb_tot[2] += cb_t[0], cb_tot[3] += cb_t[1]  (:232-234)
            if (v.y >= 0 && v.y < 1024) atomicAdd(&sh[FPC_HIST_VQ1 + v.y], 1u);
            if (v.z >= 0 && v.z < 1024) atomicAdd(&sh[FPC_HIST_VQ2 + v.z], 1u);
        } else {              // below threshold: the LAST stage of the book is counted, cb_tot[4] += cb_t[-1]  (:240)
            const int last = v.z >= 0 ? v.z : v.y;
            if (last >= 0 && last < 1024) atomicAdd(&sh[FPC_HIST_BL_VQ + last], 1u);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < FPC_HIST_TOTAL; i += blockDim.x)
        if (sh[i]) atomicAdd(&hist[i], (unsigned long long)sh[i]);
}

}  // namespace fpc

namespace fpc { long long *g_phase_buffer = nullptr; }

using namespace fpc;

extern "C" {

/* debug: 8 int64 counters (device memory, zeroed by the caller); CTA 0 adds the cycles it spent per phase */
int fpc_debug_set_phase_buffer(void *d_buf) { g_phase_buffer = (long long *)d_buf; return FPC_OK; }

size_t fpc_encode_workspace_bytes(int B, int L, int precision)
{
    (void)B; (void)L; (void)precision;
    return 0;   // all per-frame state lives in shared memory
}

int fpc_encode(const void *d_packed_weights, const void *d_packed_codebooks, const fpc_encode_io *io, int precision,
               void *d_workspace, size_t workspace_bytes, void *stream)
{
    (void)d_workspace; (void)workspace_bytes;
    if (!d_packed_weights || !io) return FPC_ERR_ARG;
    if (io->B < 0 || io->L < 0) return FPC_ERR_ARG;
    if (io->B == 0 || io->L == 0) return FPC_OK;
    if (!io->d_feat || !io->d_c_in || !io->d_r || !io->d_r_qtz) return FPC_ERR_ARG;
    if (io->qtz && !d_packed_codebooks) return FPC_ERR_ARG;
    if (precision != FPC_PREC_FP32 && precision != FPC_PREC_BF16) return FPC_ERR_UNSUPPORTED;
    EncodeParams P;
    P.wstream = (const float *)d_packed_weights;
    P.cb = (const char *)d_packed_codebooks;
    P.feat = io->d_feat;
    P.mask = io->d_mask;
    P.rq_in = nullptr;
    P.pitch_in = nullptr;
    P.c_in = io->d_c_in; P.r = io->d_r; P.r_qtz = io->d_r_qtz; P.r_under = io->d_r_under;
    P.ind1 = io->d_ind1; P.ind2 = io->d_ind2; P.idx = io->d_idx;
    P.B = io->B; P.L = io->L;
    P.mode = io->qtz ? kModeQuantize : kModeResidual;
    P.l1 = io->l1; P.l2 = io->l2;
    P.ntiles = 0;
    P.f0 = 0; P.f1 = io->L; P.state = nullptr;
    P.prof = g_phase_buffer;
    if (precision == FPC_PREC_BF16) return run_encode_bf16(P, (cudaStream_t)stream, 0);
    return run_encode_fp32(P, (cudaStream_t)stream, 0);
}

int fpc_encode_plan(int B, int precision, int sms, int *segments)
{
    if (B < 0 || !segments) return -FPC_ERR_ARG;
    if (precision != FPC_PREC_FP32 && precision != FPC_PREC_BF16) return -FPC_ERR_UNSUPPORTED;
    if (sms <= 0) sms = num_sms();
    if (sms <= 0) return -FPC_ERR_CUDA;
    return encode_plan(B, precision, sms, segments);
}

// ---- host-buffer form: time-chunked upload / compute / download pipeline ----
namespace {
// One copy pipeline (upload stream, download stream, events) per device, created on first use and kept for the
// life of the process.  `busy` is recorded at the end of every call: the next call on the same device makes its
// copy streams wait for it, so two calls that share a workspace are ordered even when they are issued on
// different streams.  Calls on ONE device must not be issued concurrently from several host threads (the
// mutex only protects the table, not the order of the enqueued work).
struct HostPipe {
    cudaStream_t s_in = nullptr, s_out = nullptr;
    cudaEvent_t busy = nullptr;
    bool used = false;
    std::vector<cudaEvent_t> ev;
    int ensure(size_t nev)
    {
        if (!s_in) FPC_CUDA_TRY(cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking));
        if (!busy) FPC_CUDA_TRY(cudaEventCreateWithFlags(&busy, cudaEventDisableTiming));
        if (!s_out) FPC_CUDA_TRY(cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking));
        while (ev.size() < nev) {
            cudaEvent_t e;
            FPC_CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            ev.push_back(e);
        }
        return FPC_OK;
    }
};
HostPipe g_host_pipes[kMaxDevices];
std::mutex g_host_pipe_mutex;
const int kHostWords[8] = {20, 20, 18, 18, 18, 1, 1, 4};   // feat, c_in, r, r_qtz, r_under, ind1, ind2, idx (4-byte words per frame)
size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }
}  // namespace

size_t fpc_encode_host_workspace_bytes(int B, int L, int precision)
{
    if (B <= 0 || L <= 0) return 0;
    size_t total = 0;
    for (int i = 0; i < 8; ++i) total += align256((size_t)B * L * kHostWords[i] * 4);
    total += align256(precision == FPC_PREC_FP32 ? encode_fp32_state_bytes(B) : encode_bf16_state_bytes(B));
    total += align256(sizeof(unsigned long long) * FPC_HIST_TOTAL);
    return total;
}

int fpc_encode_host(const void *d_packed_weights, const void *d_packed_codebooks, const fpc_encode_host_io *io,
                    int precision, int chunks, void *d_workspace, size_t workspace_bytes, void *stream)
{
    if (!d_packed_weights || !io) return FPC_ERR_ARG;
    if (io->B < 0 || io->L < 0) return FPC_ERR_ARG;
    if (io->B == 0 || io->L == 0) return FPC_OK;
    if (!io->h_feat) return FPC_ERR_ARG;
    if (io->qtz && !d_packed_codebooks) return FPC_ERR_ARG;
    if (precision != FPC_PREC_FP32 && precision != FPC_PREC_BF16) return FPC_ERR_UNSUPPORTED;
    const int B = io->B, L = io->L;
    if (!d_workspace || workspace_bytes < fpc_encode_host_workspace_bytes(B, L, precision)) return FPC_ERR_WORKSPACE;
    if (chunks <= 0) chunks = L / 48 < 1 ? 1 : (L / 48 > 16 ? 16 : L / 48);
    if (chunks > L) chunks = L;
    if (chunks > 64) chunks = 64;

    char *ws = (char *)d_workspace;
    void *dev[8];
    for (int i = 0; i < 8; ++i) { dev[i] = ws; ws += align256((size_t)B * L * kHostWords[i] * 4); }
    void *state = ws;
    ws += align256(precision == FPC_PREC_FP32 ? encode_fp32_state_bytes(B) : encode_bf16_state_bytes(B));
    unsigned long long *dev_hist = (unsigned long long *)ws;
    void *host_out[8] = {nullptr, io->h_c_in, io->h_r, io->h_r_qtz, io->h_r_under, io->h_ind1, io->h_ind2, io->h_idx};

    HostPipe &hp = g_host_pipes[device_slot()];
    {
        std::lock_guard<std::mutex> lock(g_host_pipe_mutex);
        const int rc = hp.ensure((size_t)2 * chunks + 3);
        if (rc != FPC_OK) return rc;
    }
    cudaStream_t st = (cudaStream_t)stream;
    cudaEvent_t ev_start = hp.ev[2 * chunks], ev_done = hp.ev[2 * chunks + 1];
    FPC_CUDA_TRY(cudaEventRecord(ev_start, st));            // everything queued before this call on `stream` goes first
    FPC_CUDA_TRY(cudaStreamWaitEvent(hp.s_in, ev_start, 0));
    FPC_CUDA_TRY(cudaStreamWaitEvent(hp.s_out, ev_start, 0));
    if (hp.used) {                                          // ... and so does the previous call on this device, whatever its stream
        FPC_CUDA_TRY(cudaStreamWaitEvent(hp.s_in, hp.busy, 0));
        FPC_CUDA_TRY(cudaStreamWaitEvent(st, hp.busy, 0));
    }

    EncodeParams P;
    P.wstream = (const float *)d_packed_weights;
    P.cb = (const char *)d_packed_codebooks;
    P.feat = (const float *)dev[0];
    P.mask = nullptr; P.rq_in = nullptr; P.pitch_in = nullptr;
    P.c_in = (float *)dev[1]; P.r = (float *)dev[2]; P.r_qtz = (float *)dev[3]; P.r_under = (float *)dev[4];
    P.ind1 = (float *)dev[5]; P.ind2 = (float *)dev[6]; P.idx = (int32_t *)dev[7];
    P.B = B; P.L = L;
    P.mode = io->qtz ? kModeQuantize : kModeResidual;
    P.l1 = io->l1; P.l2 = io->l2;
    P.ntiles = 0;
    P.state = chunks > 1 ? state : nullptr;
    P.prof = nullptr;

    auto range = [&](int c) { return (int)((long long)L * c / chunks); };
    // a frame range of a (B, L, words) array is a 2-D block: B rows of (f1 - f0) * words * 4 bytes, pitch L * words * 4
    for (int c = 0; c < chunks; ++c) {
        const int f0 = range(c), f1 = range(c + 1);
        const size_t pitch = (size_t)L * kHostWords[0] * 4, off = (size_t)f0 * kHostWords[0] * 4;
        FPC_CUDA_TRY(cudaMemcpy2DAsync((char *)dev[0] + off, pitch, (const char *)io->h_feat + off, pitch,
                                       (size_t)(f1 - f0) * kHostWords[0] * 4, (size_t)B, cudaMemcpyHostToDevice, hp.s_in));
        FPC_CUDA_TRY(cudaEventRecord(hp.ev[c], hp.s_in));
    }
    for (int c = 0; c < chunks; ++c) {
        P.f0 = range(c); P.f1 = range(c + 1);
        FPC_CUDA_TRY(cudaStreamWaitEvent(st, hp.ev[c], 0));
        const int rc = precision == FPC_PREC_BF16 ? run_encode_bf16(P, st, 0) : run_encode_fp32(P, st, 0);
        if (rc != FPC_OK) return rc;
        FPC_CUDA_TRY(cudaEventRecord(hp.ev[chunks + c], st));
        FPC_CUDA_TRY(cudaStreamWaitEvent(hp.s_out, hp.ev[chunks + c], 0));
        for (int i = 1; i < 8; ++i) {
            if (!host_out[i]) continue;
            const size_t pitch = (size_t)L * kHostWords[i] * 4, off = (size_t)P.f0 * kHostWords[i] * 4;
            FPC_CUDA_TRY(cudaMemcpy2DAsync((char *)host_out[i] + off, pitch, (const char *)dev[i] + off, pitch,
                                           (size_t)(P.f1 - P.f0) * kHostWords[i] * 4, (size_t)B, cudaMemcpyDeviceToHost,
                                           hp.s_out));
        }
    }
    if (io->h_hist) {
        // cb_tot (wavernn.py:189,221-240) of the whole call: counted on the device from the index record, 28 KB down
        if (!io->qtz) {
            FPC_CUDA_TRY(cudaMemsetAsync(dev_hist, 0, sizeof(unsigned long long) * FPC_HIST_TOTAL, st));
        } else {
            const int rc = fpc_index_histogram((const int32_t *)dev[7], (long)B * L, dev_hist, st);
            if (rc != FPC_OK) return rc;
        }
        cudaEvent_t ev_hist = hp.ev[2 * chunks + 2];
        FPC_CUDA_TRY(cudaEventRecord(ev_hist, st));
        FPC_CUDA_TRY(cudaStreamWaitEvent(hp.s_out, ev_hist, 0));
        FPC_CUDA_TRY(cudaMemcpyAsync(io->h_hist, dev_hist, sizeof(unsigned long long) * FPC_HIST_TOTAL, cudaMemcpyDeviceToHost, hp.s_out));
    }
    FPC_CUDA_TRY(cudaEventRecord(ev_done, hp.s_out));
    FPC_CUDA_TRY(cudaStreamWaitEvent(st, ev_done, 0));     // `stream` completes only after the last download
    FPC_CUDA_TRY(cudaEventRecord(hp.busy, st));
    hp.used = true;
    return FPC_OK;
}

int fpc_decode(const void *d_packed_weights, const float *d_r_qtz, const float *d_pitch, int B, int L, float *d_c_out,
               int precision, void *d_workspace, size_t workspace_bytes, void *stream)
{
    (void)d_workspace; (void)workspace_bytes;
    if (!d_packed_weights) return FPC_ERR_ARG;
    if (B < 0 || L < 0) return FPC_ERR_ARG;
    if (B == 0 || L == 0) return FPC_OK;
    if (!d_r_qtz || !d_pitch || !d_c_out) return FPC_ERR_ARG;
    if (precision != FPC_PREC_FP32 && precision != FPC_PREC_BF16) return FPC_ERR_UNSUPPORTED;
    EncodeParams P;
    P.wstream = (const float *)d_packed_weights;
    P.cb = nullptr; P.feat = nullptr; P.mask = nullptr;
    P.rq_in = d_r_qtz; P.pitch_in = d_pitch;
    P.c_in = d_c_out; P.r = nullptr; P.r_qtz = nullptr; P.r_under = nullptr;
    P.ind1 = nullptr; P.ind2 = nullptr; P.idx = nullptr;
    P.B = B; P.L = L; P.mode = kModeDecode; P.l1 = 0.0f; P.l2 = 0.0f; P.ntiles = 0; P.prof = nullptr;
    P.f0 = 0; P.f1 = L; P.state = nullptr;
    if (precision == FPC_PREC_BF16) return run_encode_bf16(P, (cudaStream_t)stream, 0);
    return run_encode_fp32(P, (cudaStream_t)stream, 0);
}

int fpc_index_histogram(const int32_t *d_idx, long n_frames, unsigned long long *d_hist, void *stream)
{
    if (!d_hist || n_frames < 0) return FPC_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    FPC_CUDA_TRY(cudaMemsetAsync(d_hist, 0, sizeof(unsigned long long) * FPC_HIST_TOTAL, st));
    if (n_frames == 0) return FPC_OK;
    if (!d_idx) return FPC_ERR_ARG;
    long blocks = (n_frames + 1023) / 1024;
    const int cap = 4 * (num_sms() > 0 ? num_sms() : 148);
    if (blocks > cap) blocks = cap;
    index_histogram_kernel<<<(int)blocks, 256, 0, st>>>((const int4 *)d_idx, n_frames, d_hist);
    FPC_LAUNCH_CHECK();
    return FPC_OK;
}

/* stand-alone vq_quantize on a packed codebook image; `which` 0 = cfg['cb_path'] slot,
 * 1 = cfg['bl_cb_path'] slot (the Python mirror caches one packed image per codebook file). */
int fpc_vq_quantize_packed(const float *d_x, long n, const void *d_packed_codebooks, int which, int dtype, int stages,
                           void *d_q, int32_t *d_idx, void *stream)
{
    if (n < 0) return FPC_ERR_ARG;
    if (n == 0) return FPC_OK;
    if (!d_x || !d_packed_codebooks || !d_q || !d_idx) return FPC_ERR_ARG;
    if (stages < 1 || stages > 2) return FPC_ERR_CODEBOOK;
    cudaStream_t st = (cudaStream_t)stream;
    const int sms = num_sms();
    if (sms <= 0) return cuda_fail(cudaErrorNoDevice);
    long tiles = (n + kQTile - 1) / kQTile;
    const int grid = (int)(tiles < sms ? tiles : sms);
    const size_t fixed = (size_t)(kQTile * kLdR + kQTile * 20 + 3 * kQTile) * 4;
    if (dtype == FPC_F32) {
        const size_t smem = fixed + vq_fixed_bytes<float>(kQTile) + 8 * 1024 * sizeof(float);
        static bool cfg[kMaxDevices] = {};
        { const int rc = ensure_dynamic_smem(vq_quantize_kernel<float>, (int)smem, cfg); if (rc != FPC_OK) return rc; }
        vq_quantize_kernel<float><<<grid, kComputeThreads, smem, st>>>(d_x, n, (const char *)d_packed_codebooks, which,
                                                                          (float *)d_q, d_idx, (int)(smem - fixed));
    } else if (dtype == FPC_F64) {
        const size_t smem = fixed + vq_fixed_bytes<double>(kQTile) + 8 * 1024 * sizeof(double);
        static bool cfg[kMaxDevices] = {};
        { const int rc = ensure_dynamic_smem(vq_quantize_kernel<double>, (int)smem, cfg); if (rc != FPC_OK) return rc; }
        vq_quantize_kernel<double><<<grid, kComputeThreads, smem, st>>>(d_x, n, (const char *)d_packed_codebooks, which,
                                                                           (double *)d_q, d_idx, (int)(smem - fixed));
    } else {
        return FPC_ERR_CODEBOOK;
    }
    FPC_LAUNCH_CHECK();
    return FPC_OK;
}

int fpc_scl_quantize(const float *d_x, long n, const void *d_codes, int dtype, int n_code, void *d_q, int32_t *d_idx,
                     void *stream)
{
    if (n < 0) return FPC_ERR_ARG;
    if (n == 0) return FPC_OK;
    if (!d_x || !d_codes || !d_q || !d_idx) return FPC_ERR_ARG;
    if (n_code < 1 || n_code > FPC_MAX_SCL_ENTRIES) return FPC_ERR_CODEBOOK;
    cudaStream_t st = (cudaStream_t)stream;
    long blocks = (n + 7) / 8;
    if (blocks > 1184) blocks = 1184;
    if (dtype == FPC_F32)
        scl_quantize_kernel<float><<<(int)blocks, 256, 0, st>>>(d_x, n, (const float *)d_codes, n_code, (float *)d_q, d_idx);
    else if (dtype == FPC_F64)
        scl_quantize_kernel<double><<<(int)blocks, 256, 0, st>>>(d_x, n, (const double *)d_codes, n_code, (double *)d_q, d_idx);
    else
        return FPC_ERR_CODEBOOK;
    FPC_LAUNCH_CHECK();
    return FPC_OK;
}

}  // extern "C"
