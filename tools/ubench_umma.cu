// Micro-benchmark: time per tcgen05.mma (kind::f16, M = 128, N = 64 / 128 / 256, K = 16) when consecutive MMAs accumulate
// into the SAME tensor-memory tile (a dependent chain, as the four K = 16 slabs of a K = 64 product do) and when they go
// round-robin over several tiles.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o ubench_umma ubench_umma.cu
#include <cstdio>
#include "../feature-predictor-for-speech-codec_b200/csrc/fpc_umma.cuh"
#include "../feature-predictor-for-speech-codec_b200/csrc/fpc_tc.cuh"
using namespace fpc;

template <int N, int TILES>
__global__ void __launch_bounds__(128, 1) k(long long *out, int iters)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint32_t slot;
    __shared__ uint64_t bar;
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t *>(smem)[i] = 0;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    if (warp == 0) umma::tmem_alloc(&slot, 512);
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tb = slot;
    if (warp == 1) {
        const uint32_t idesc = tc::instr_desc_f16(128, N);
        const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 32768);
        umma::fence_async_smem();
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int t = 0; t < TILES; ++t)
                umma::mma_bf16_elect(tb + (uint32_t)((t * N) % 512), umma::smem_desc(a0 + (it & 3) * 4096, 128),
                                     umma::smem_desc(b0 + (it & 3) * (N * 32), N), idesc, it > 0);
        }
        umma::commit_elect(&bar);
        mbar_wait(&bar, 0);
        const long long t1 = clock64();
        if ((threadIdx.x & 31) == 0) out[0] = (t1 - t0);
    }
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tb, 512);
}

template <int N, int TILES> void run(const char *what)
{
    long long *d, h = 0;
    cudaMalloc(&d, 8);
    const int iters = 256;
    cudaFuncSetAttribute(k<N, TILES>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    for (int r = 0; r < 2; ++r) {
        k<N, TILES><<<1, 128, 96 * 1024>>>(d, iters);
        cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    }
    printf("%-40s N=%3d tiles=%d : %7.1f cycles per MMA (%lld cycles for %d MMAs) %s\n", what, N, TILES, (double)h / (iters * TILES), h, iters * TILES,
           cudaGetErrorString(cudaGetLastError()));
    cudaFree(d);
}

int main()
{
    run<64, 1>("dependent chain");
    run<64, 2>("two tiles round-robin");
    run<64, 3>("three tiles round-robin");
    run<64, 8>("eight tiles round-robin");
    run<128, 1>("dependent chain");
    run<128, 4>("four tiles round-robin");
    run<256, 1>("dependent chain");
    run<256, 2>("two tiles round-robin");
    return 0;
}
