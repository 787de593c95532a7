// fpc_umma_selftest.cu -- D (128 x N, f32) = A (128 x K, bf16) * B (N x K, bf16)^T on the tcgen05
// tensor cores, through exactly the helpers (fpc_umma.cuh) and the operand layout the bf16
// predictor uses.  Exposed as fpc_selftest_umma so the GPU tests can pin the descriptor
// encodings against a plain fp32 reference of the same product.
#include "fpc_umma.cuh"

namespace fpc {

__global__ void __launch_bounds__(128, 1)
umma_selftest_kernel(const __nv_bfloat16 *__restrict__ A, const __nv_bfloat16 *__restrict__ B, int N, int K,
                     float *__restrict__ D)
{
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char *sA = smem;                         // 128 x K
    unsigned char *sB = smem + (size_t)128 * K * 2;   // N x K
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) umma::tmem_alloc(&tmem_base, 256);
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    for (int i = tid; i < 128 * K; i += 128) {
        const int r = i / K, k = i - r * K;
        *reinterpret_cast<__nv_bfloat16 *>(sA + umma::tile_off(r, k, 128)) = A[i];
    }
    for (int i = tid; i < N * K; i += 128) {
        const int r = i / K, k = i - r * K;
        *reinterpret_cast<__nv_bfloat16 *>(sB + umma::tile_off(r, k, N)) = B[i];
    }
    umma::fence_async_smem();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tb = tmem_base;
    if (tid == 0) {
        const uint32_t idesc = umma::instr_desc_bf16(128, N);
        for (int ks = 0; ks < K / 16; ++ks) {
            const uint64_t ad = umma::smem_desc(smem_u32(sA) + ks * 2 * 128 * 16, 128);
            const uint64_t bd = umma::smem_desc(smem_u32(sB) + ks * 2 * N * 16, N);
            umma::mma_bf16(tb, ad, bd, idesc, ks > 0 ? 1u : 0u);
        }
        umma::commit(&bar);
    }
    mbar_wait(&bar, 0);
    umma::fence_after_sync();
    for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t r[16];
        umma::tmem_ld16(tb + ((uint32_t)(warp * 32) << 16) + c0, r);
        umma::tmem_ld16_wait(r);
#pragma unroll
        for (int j = 0; j < 16; ++j) D[(size_t)tid * N + c0 + j] = __uint_as_float(r[j]);
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tb, 256);
}

}  // namespace fpc

extern "C" int fpc_selftest_umma(const void *d_a_bf16, const void *d_b_bf16, int N, int K, float *d_out, void *stream)
{
    using namespace fpc;
    if (!d_a_bf16 || !d_b_bf16 || !d_out) return FPC_ERR_ARG;
    if (N < 16 || N > 256 || (N & 15) || K < 16 || (K & 15)) return FPC_ERR_SHAPE;
    const size_t smem = (size_t)(128 + N) * K * 2;
    if (smem > 200 * 1024) return FPC_ERR_SHAPE;
    FPC_CUDA_TRY(cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    umma_selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>((const __nv_bfloat16 *)d_a_bf16,
                                                                 (const __nv_bfloat16 *)d_b_bf16, N, K, d_out);
    FPC_LAUNCH_CHECK();
    return FPC_OK;
}
