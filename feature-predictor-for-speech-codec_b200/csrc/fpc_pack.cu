// fpc_pack.cu -- re-tiling of the predictor parameters and codebook files into the images the
// frame-step kernel streams.  Reference objects: the state_dict of Wavernn
// (/root/reference/src/models/wavernn.py:37-38,48-52) and the .npy files loaded at
// quantization/vq_func.py:141,171.
#include "fpc_common.cuh"
#include "fpc_encode.cuh"
#include "fpc_tc.cuh"

namespace fpc {

thread_local int g_last_cuda_error = 0;
unsigned long long g_launch_count = 0;

// ------------------------------------------------------------------------------------------
// fp32 weight stream.  One thread per destination float.
//   group g of a pass -> [gate][k-pair p][unit pair ug 0..63][(e0,2p) (e1,2p) (e0,2p+1) (e1,2p+1)]
//   hidden unit j = pass*128 + 2*ug + e ; k = kGk*(group index within part) + 2*p + {0,1}
// ------------------------------------------------------------------------------------------
__global__ void pack_weights_f32_kernel(fpc_weights w, float *__restrict__ out0)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    float *__restrict__ out = reinterpret_cast<float *>(reinterpret_cast<char *>(out0) + (size_t)blockIdx.y * kPackedF32ReplicaBytes);
    if (t < kStreamFloats) {
        int g = t / kGroupFloats;
        int r = t - g * kGroupFloats;
        // [gate][k-pair p][unit pair ug][(e0,2p) (e1,2p) (e0,2p+1) (e1,2p+1)]
        int gate = r / ((kGk / 2) * 256), r2 = r - gate * ((kGk / 2) * 256);
        int pq = r2 / 256, ug = (r2 >> 2) & 63, kk = 2 * pq + ((r2 >> 1) & 1), e = r2 & 1;
        float v = 0.0f;
        if (g < 3 * kG1) {
            int pass = g / kG1, gi = g - pass * kG1;
            int j = pass * 128 + 2 * ug + e;
            int row = gate * kH1 + j;
            if (gi < kG1x) {
                int k = kGk * gi + kk;
                if (k < kIn) v = w.w_ih1[(size_t)row * kIn + k];
            } else {
                v = w.w_hh1[(size_t)row * kH1 + kGk * (gi - kG1x) + kk];
            }
        } else {
            int gi = g - 3 * kG1;
            int j = 2 * ug + e;
            int row = gate * kH2 + j;
            if (gi < kG2x) v = w.w_ih2[(size_t)row * kH1 + kGk * gi + kk];
            else v = w.w_hh2[(size_t)row * kH2 + kGk * (gi - kG2x) + kk];
        }
        out[t] = v;
        return;
    }
    t -= kStreamFloats;
    if (t < kBiasFloats) {
        // [pass 0..3][kind: br, bz, b_in, b_hn][unit 0..127]; pass 3 is GRU2
        int pass = t / 512, kind = (t >> 7) & 3, u = t & 127;
        const float *bi = pass < 3 ? w.b_ih1 : w.b_ih2;
        const float *bh = pass < 3 ? w.b_hh1 : w.b_hh2;
        int H = pass < 3 ? kH1 : kH2;
        int j = pass < 3 ? pass * 128 + u : u;
        float v;
        if (kind == 0) v = __fadd_rn(bi[j], bh[j]);
        else if (kind == 1) v = __fadd_rn(bi[H + j], bh[H + j]);
        else if (kind == 2) v = bi[2 * H + j];
        else v = bh[2 * H + j];
        out[kStreamFloats + t] = v;
        return;
    }
    t -= kBiasFloats;
    if (t < kFcFloats) {
        out[kStreamFloats + kBiasFloats + t] = w.w_fc[t];
        return;
    }
    t -= kFcFloats;
    if (t < kFc) out[kStreamFloats + kBiasFloats + kFcFloats + t] = w.b_fc[t];
}

// ------------------------------------------------------------------------------------------
// codebooks: each VQ stage is stored twice, transposed [17][K] for the search (coalesced
// per-thread codeword loads) and row-major [K][17] for the gathers.
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void pack_vq_kernel(const T *__restrict__ src, int stages, int K, T *__restrict__ dst_t0,
                               T *__restrict__ dst_r0, T *__restrict__ dst_t1, T *__restrict__ dst_r1)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    int per = K * kDim;
    if (t >= stages * per) return;
    int s = t / per, r = t - s * per;
    int k = r / kDim, d = r - k * kDim;
    T v = src[t];
    T *dt = s == 0 ? dst_t0 : dst_t1;
    T *dr = s == 0 ? dst_r0 : dst_r1;
    dt[(size_t)d * K + k] = v;
    dr[r] = v;
}

// screening data of one stage: fp32 shadow [17][Kp], squared norms [Kp] (float64 sum, rounded),
// and an upper bound of max ||c|| (atomicMax on the bit pattern of a non-negative float)
template <typename T>
__global__ void pack_screen_kernel(const T *__restrict__ src /* [K][17] of this stage */, int K, int Kp,
                                   float *__restrict__ shadow, float *__restrict__ norms, float *__restrict__ cmax)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= Kp) return;
    if (k >= K) {
        for (int d = 0; d < kDim; ++d) shadow[(size_t)d * Kp + k] = 0.0f;
        norms[k] = 3.0e38f;
        return;
    }
    double n2 = 0.0;
    for (int d = 0; d < kDim; ++d) {
        const double c = (double)src[(size_t)k * kDim + d];
        shadow[(size_t)d * Kp + k] = (float)c;
        n2 += c * c;
    }
    norms[k] = (float)n2;
    const float nb = __fmul_ru(__fsqrt_ru(__double2float_ru(n2)), 1.0000001f);
    atomicMax(reinterpret_cast<unsigned int *>(cmax), __float_as_uint(nb));
}

// Gram table of a two-stage book: G[k0][k1] = 2 <c0_k0, c1_k1>, float64 dot rounded to fp32
template <typename T>
__global__ void pack_gram_kernel(const T *__restrict__ c0, const T *__restrict__ c1, int K, int Kp, float *__restrict__ G)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)K * Kp) return;
    const int k0 = (int)(t / Kp), k1 = (int)(t - (long long)k0 * Kp);
    double acc = 0.0;
    if (k1 < K)
        for (int d = 0; d < kDim; ++d) acc += (double)c0[(size_t)k0 * kDim + d] * (double)c1[(size_t)k1 * kDim + d];
    G[t] = (float)(2.0 * acc);
}

// Tensor-core screen of the m-best search (fpc_vq_tc.cuh): the B operand image of one stage.  One CTA: largest |c|
// -> power-of-two scale beta, then every codeword as an fp16-pair row (fpc_tc.cuh) in tiles of 64 rows; rows K..Kp64-1
// are padding that can never win (norm entries 3 x 65504).  Entries 54..56 hold 1024: the vector side multiplies them
// by its offset / 1024 (stage 0 of two-stage books).
template <typename T>
__global__ void __launch_bounds__(1024, 1)
pack_tc_image_kernel(const T *__restrict__ src /* [K][17] */, int K, int Kp64, unsigned char *__restrict__ image0, long long rep_stride,
                     float *__restrict__ beta_out)
{
    unsigned char *__restrict__ image = image0 + (size_t)blockIdx.x * rep_stride;       // one CTA per replica
    __shared__ float s_red[32];
    float amax = 0.0f;
    for (int i = threadIdx.x; i < K * kDim; i += blockDim.x) amax = fmaxf(amax, __double2float_ru(fabs((double)src[i])));
    for (int off = 16; off > 0; off >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, off));
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = amax;
    __syncthreads();
    amax = 0.0f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) amax = fmaxf(amax, s_red[w]);
    const float beta = tc::scale_for(amax);
    if (threadIdx.x == 0 && blockIdx.x == 0) *beta_out = beta;
    const __half k1024 = __float2half_rn(1024.0f);
    for (int k = threadIdx.x; k < Kp64; k += blockDim.x) {
        float xs[kDim];
        __half n0, n1, n2;
        if (k < K) {
            double nn = 0.0;
#pragma unroll
            for (int d = 0; d < kDim; ++d) {
                const double c = (double)src[(size_t)k * kDim + d];
                xs[d] = (float)(c * (double)beta);
                nn += c * c;
            }
            tc::split3((float)(nn * (double)beta * (double)beta), n0, n1, n2);
        } else {
#pragma unroll
            for (int d = 0; d < kDim; ++d) xs[d] = 0.0f;
            n0 = n1 = n2 = __float2half_rn(tc::kPadNorm);
        }
        tc::store_row<false, 64>(image + (size_t)(k >> 6) * (64 * tc::kK * 2), k & 63, xs, n0, n1, n2, k1024, k1024, k1024);
    }
}

template <typename T>
__global__ void copy_kernel(const T *__restrict__ src, int n, T *__restrict__ dst)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) dst[t] = src[t];
}

static int check_vq(const void *p, int dtype, int stages, int K, bool required)
{
    if (stages == 0) return required ? FPC_ERR_CODEBOOK : FPC_OK;
    if (!p) return FPC_ERR_ARG;
    if (dtype != FPC_F32 && dtype != FPC_F64) return FPC_ERR_CODEBOOK;
    if (stages < 1 || stages > 2) return FPC_ERR_CODEBOOK;       // >2 stages raise in vq_func.py:111
    if (K < kSurv || K > FPC_MAX_VQ_ENTRIES) return FPC_ERR_CODEBOOK;  // sorted()[:5] needs >= 5 entries
    return FPC_OK;
}

static int check_scl(const void *p, int dtype, int n, bool required)
{
    if (n == 0) return required ? FPC_ERR_CODEBOOK : FPC_OK;
    if (!p) return FPC_ERR_ARG;
    if (dtype != FPC_F32 && dtype != FPC_F64) return FPC_ERR_CODEBOOK;
    if (n < 1 || n > FPC_MAX_SCL_ENTRIES) return FPC_ERR_CODEBOOK;
    return FPC_OK;
}

static int pack_one_vq(const void *src, int dtype, int stages, int K, char *base, size_t &cursor, PackedVq &h,
                       cudaStream_t st)
{
    h.dtype = dtype; h.stages = stages; h.K = K; h.Kp = (K + 3) & ~3;
    h.off_t[0] = h.off_t[1] = h.off_r[0] = h.off_r[1] = 0;
    h.off_f[0] = h.off_f[1] = h.off_n[0] = h.off_n[1] = h.off_g = h.off_cmax = 0;
    h.off_b[0] = h.off_b[1] = 0; h.off_tcbeta = 0; h.b_rep_stride = 0; h.Kp64 = (K + 63) & ~63; h.pad_ = 0;
    if (stages == 0) return FPC_OK;
    size_t es = dtype == FPC_F32 ? 4 : 8;
    size_t one = (((size_t)K * kDim * es) + 255) / 256 * 256;
    for (int s = 0; s < stages; ++s) {
        h.off_t[s] = (long long)cursor; cursor += one;
        h.off_r[s] = (long long)cursor; cursor += one;
    }
    int n = stages * K * kDim;
    int blocks = (n + 255) / 256;
    if (dtype == FPC_F32)
        pack_vq_kernel<float><<<blocks, 256, 0, st>>>((const float *)src, stages, K, (float *)(base + h.off_t[0]),
                                                      (float *)(base + h.off_r[0]), (float *)(base + h.off_t[1]),
                                                      (float *)(base + h.off_r[1]));
    else
        pack_vq_kernel<double><<<blocks, 256, 0, st>>>((const double *)src, stages, K, (double *)(base + h.off_t[0]),
                                                       (double *)(base + h.off_r[0]), (double *)(base + h.off_t[1]),
                                                       (double *)(base + h.off_r[1]));
    FPC_LAUNCH_CHECK();
    // ---- screening data ----
    const int Kp = h.Kp;
    h.off_cmax = (long long)cursor; cursor += 256;
    FPC_CUDA_TRY(cudaMemsetAsync(base + h.off_cmax, 0, 256, st));
    for (int s = 0; s < stages; ++s) {
        h.off_f[s] = (long long)cursor; cursor += (((size_t)kDim * Kp * 4) + 255) / 256 * 256;
        h.off_n[s] = (long long)cursor; cursor += (((size_t)Kp * 4) + 255) / 256 * 256;
        float *cm = (float *)(base + h.off_cmax) + s;
        if (dtype == FPC_F32)
            pack_screen_kernel<float><<<(Kp + 127) / 128, 128, 0, st>>>((const float *)src + (size_t)s * K * kDim, K, Kp,
                                                                         (float *)(base + h.off_f[s]), (float *)(base + h.off_n[s]), cm);
        else
            pack_screen_kernel<double><<<(Kp + 127) / 128, 128, 0, st>>>((const double *)src + (size_t)s * K * kDim, K, Kp,
                                                                          (float *)(base + h.off_f[s]), (float *)(base + h.off_n[s]), cm);
        FPC_LAUNCH_CHECK();
    }
    // ---- tensor-core operand images and their scales ----
    {
        const long long off_beta = (long long)cursor; cursor += 256;
        h.off_tcbeta = off_beta;
        const size_t one_image = (size_t)h.Kp64 * tc::kK * 2;
        h.b_rep_stride = (long long)(stages * one_image);
        for (int s = 0; s < stages; ++s) {
            h.off_b[s] = (long long)(cursor + s * one_image);
            float *bo = (float *)(base + off_beta) + s;
            if (dtype == FPC_F32)
                pack_tc_image_kernel<float><<<kWeightReplicas, 1024, 0, st>>>((const float *)src + (size_t)s * K * kDim, K, h.Kp64,
                                                                               (unsigned char *)(base + h.off_b[s]), h.b_rep_stride, bo);
            else
                pack_tc_image_kernel<double><<<kWeightReplicas, 1024, 0, st>>>((const double *)src + (size_t)s * K * kDim, K, h.Kp64,
                                                                                (unsigned char *)(base + h.off_b[s]), h.b_rep_stride, bo);
            FPC_LAUNCH_CHECK();
        }
        cursor += (size_t)kWeightReplicas * stages * one_image;
    }
    if (stages == 2) {
        h.off_g = (long long)cursor; cursor += (((size_t)K * Kp * 4) + 255) / 256 * 256;
        const long long n2 = (long long)K * Kp;
        const int gb = (int)((n2 + 255) / 256);
        if (dtype == FPC_F32)
            pack_gram_kernel<float><<<gb, 256, 0, st>>>((const float *)src, (const float *)src + (size_t)K * kDim, K, Kp,
                                                         (float *)(base + h.off_g));
        else
            pack_gram_kernel<double><<<gb, 256, 0, st>>>((const double *)src, (const double *)src + (size_t)K * kDim, K, Kp,
                                                          (float *)(base + h.off_g));
        FPC_LAUNCH_CHECK();
    }
    return FPC_OK;
}

static int pack_one_scl(const void *src, int dtype, int n, char *base, size_t &cursor, PackedScl &h, cudaStream_t st)
{
    h.dtype = dtype; h.n = n; h.off = 0;
    if (n == 0) return FPC_OK;
    h.off = (long long)cursor;
    cursor += kCbSclMaxBytes;
    if (dtype == FPC_F32) copy_kernel<float><<<1, 256, 0, st>>>((const float *)src, n, (float *)(base + h.off));
    else copy_kernel<double><<<1, 256, 0, st>>>((const double *)src, n, (double *)(base + h.off));
    FPC_LAUNCH_CHECK();
    return FPC_OK;
}

}  // namespace fpc

using namespace fpc;

extern "C" {

int fpc_version(void) { return FPC_VERSION; }

const char *fpc_status_string(int s)
{
    switch (s) {
        case FPC_OK: return "ok";
        case FPC_ERR_ARG: return "bad argument (null pointer or negative size)";
        case FPC_ERR_SHAPE: return "unsupported dimensions";
        case FPC_ERR_CODEBOOK: return "bad codebook (stages must be 1 or 2, 5..1024 entries, <=256 scalar levels, f32/f64)";
        case FPC_ERR_WORKSPACE: return "workspace or packed buffer too small";
        case FPC_ERR_CUDA: return "CUDA runtime error";
        case FPC_ERR_UNSUPPORTED: return "precision/mode not supported by this build";
        default: return "unknown status";
    }
}

int fpc_last_cuda_error(void) { return g_last_cuda_error; }

unsigned long long fpc_launch_count(void) { return __atomic_load_n(&g_launch_count, __ATOMIC_RELAXED); }

size_t fpc_packed_codebooks_bytes(void) { return kPackedCbBytes; }

int fpc_pack_codebooks(const fpc_codebooks *cb, void *d_packed, size_t packed_bytes, void *stream)
{
    if (!cb || !d_packed) return FPC_ERR_ARG;
    if (packed_bytes < kPackedCbBytes) return FPC_ERR_WORKSPACE;
    int rc;
    if ((rc = check_vq(cb->vq, cb->vq_dtype, cb->vq_stages, cb->vq_entries, false))) return rc;
    if ((rc = check_vq(cb->bl_vq, cb->bl_vq_dtype, cb->bl_vq_stages, cb->bl_vq_entries, false))) return rc;
    if ((rc = check_scl(cb->scl, cb->scl_dtype, cb->scl_entries, false))) return rc;
    if ((rc = check_scl(cb->bl_scl, cb->bl_scl_dtype, cb->bl_scl_entries, false))) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    PackedCodebooks h;
    size_t cursor = kCbHeaderBytes;
    char *base = (char *)d_packed;
    if ((rc = pack_one_vq(cb->vq, cb->vq_dtype, cb->vq_stages, cb->vq_entries, base, cursor, h.vq, st))) return rc;
    cursor = kCbHeaderBytes + kCbVqMaxBytes;
    if ((rc = pack_one_vq(cb->bl_vq, cb->bl_vq_dtype, cb->bl_vq_stages, cb->bl_vq_entries, base, cursor, h.bl, st)))
        return rc;
    cursor = kCbHeaderBytes + 2 * kCbVqMaxBytes;
    if ((rc = pack_one_scl(cb->scl, cb->scl_dtype, cb->scl_entries, base, cursor, h.scl, st))) return rc;
    if ((rc = pack_one_scl(cb->bl_scl, cb->bl_scl_dtype, cb->bl_scl_entries, base, cursor, h.blscl, st))) return rc;
    // The header is tiny and the source is a host stack object: a synchronous copy is the
    // safe choice (cudaMemcpyAsync from pageable memory stages it before returning anyway).
    FPC_CUDA_TRY(cudaMemcpyAsync(d_packed, &h, sizeof(h), cudaMemcpyHostToDevice, st));
    return FPC_OK;
}

size_t fpc_packed_weights_bytes(int precision)
{
    if (precision == FPC_PREC_FP32) return kPackedF32ReplicaBytes * kWeightReplicas;
    if (precision == FPC_PREC_BF16) return packed_bf16_bytes();
    return 0;
}

int fpc_pack_weights(const fpc_weights *w, int precision, void *d_packed, size_t packed_bytes, void *stream)
{
    if (!w || !d_packed) return FPC_ERR_ARG;
    if (!w->w_ih1 || !w->w_hh1 || !w->b_ih1 || !w->b_hh1 || !w->w_ih2 || !w->w_hh2 || !w->b_ih2 || !w->b_hh2 ||
        !w->w_fc || !w->b_fc)
        return FPC_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    if (precision == FPC_PREC_FP32) {
        if (packed_bytes < kPackedF32ReplicaBytes * kWeightReplicas) return FPC_ERR_WORKSPACE;
        int n = kStreamFloats + kBiasFloats + kFcFloats + kFc;
        pack_weights_f32_kernel<<<dim3((n + 255) / 256, kWeightReplicas), 256, 0, st>>>(*w, (float *)d_packed);
        FPC_LAUNCH_CHECK();
        return FPC_OK;
    }
    if (precision == FPC_PREC_BF16) {
        if (packed_bytes < packed_bf16_bytes()) return FPC_ERR_WORKSPACE;
        return pack_weights_bf16(w, d_packed, st);
    }
    return FPC_ERR_UNSUPPORTED;
}

}  // extern "C"
