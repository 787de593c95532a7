// Dependent float64 add chain: cycles per DADD (what bounds fpc::ord_sum_kernel, whose sums are one serial chain per
// centroid by specification).   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ubench_dadd tools/ubench_dadd.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void chain(const float *x, double *out, long long *cyc, int n)
{
    float v[32];
    for (int j = 0; j < 32; ++j) v[j] = x[j];
    double acc = out[0];
    const long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int j = 0; j < 32; ++j) acc += (double)v[j];
    }
    const long long t1 = clock64();
    out[threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void chain_d(const double *x, double *out, long long *cyc, int n)
{
    double v[32];
    for (int j = 0; j < 32; ++j) v[j] = x[j];
    double acc = out[0];
    const long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int j = 0; j < 32; ++j) acc += v[j];
    }
    const long long t1 = clock64();
    out[threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main()
{
    float *x; double *xd, *out; long long *cyc;
    cudaMalloc(&x, 256); cudaMalloc(&xd, 512); cudaMalloc(&out, 1024); cudaMalloc(&cyc, 8);
    cudaMemset(x, 0, 256); cudaMemset(xd, 0, 512); cudaMemset(out, 0, 1024);
    for (int warps = 1; warps <= 8; warps *= 2) {
        long long c = 0;
        chain<<<1, 32 * warps>>>(x, out, cyc, 1000);
        cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        printf("float32 -> float64 add chain, %d warp(s): %.2f cycles per add\n", warps, c / 32000.0);
        chain_d<<<1, 32 * warps>>>(xd, out, cyc, 1000);
        cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        printf("float64 add chain,            %d warp(s): %.2f cycles per add\n", warps, c / 32000.0);
    }
    return cudaGetLastError() != cudaSuccess;
}
