// Micro-benchmark: issue cost of FFMA vs FFMA2 (fma.rn.f32x2) on sm_100a, per SM sub-partition.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_ffma2 ubench_ffma2.cu ; run: ./ubench_ffma2
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c)
{
    unsigned long long ra, rb, rc, rd;
    ra = ((unsigned long long)__float_as_uint(a.y) << 32) | __float_as_uint(a.x);
    rb = ((unsigned long long)__float_as_uint(b.y) << 32) | __float_as_uint(b.x);
    rc = ((unsigned long long)__float_as_uint(c.y) << 32) | __float_as_uint(c.x);
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return make_float2(__uint_as_float((unsigned)rd), __uint_as_float((unsigned)(rd >> 32)));
}

__device__ __forceinline__ float2 fma2b(float2 a, float b, float2 c)
{
    unsigned long long ra, rb, rc, rd;
    asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b), "f"(b));
    asm("mov.b64 %0, {%1, %2};" : "=l"(rc) : "f"(c.x), "f"(c.y));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return make_float2(__uint_as_float((unsigned)rd), __uint_as_float((unsigned)(rd >> 32)));
}

constexpr int kIters = 4096;

// MODE 0: scalar FFMA, NACC independent accumulators; MODE 1: FFMA2 with distinct 64-bit a and b;
// MODE 2: FFMA2 where b is a broadcast pair (same register twice is not expressible; uses (s,s) pair);
// MODE 3: FFMA2 interleaved 8:5 with FMNMX-style ALU ops
template <int MODE, int NACC>
__global__ void bench(float *out, long long *cycles, float s0, float s1)
{
    float2 acc[NACC];
    float2 a[4];
    for (int i = 0; i < NACC; ++i) acc[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f);
    for (int i = 0; i < 4; ++i) a[i] = make_float2(s0 + i * 1e-3f, s1 - i * 1e-3f);
    float m1 = 1e30f, m2 = 1e30f;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
            if (MODE == 0) {
                acc[i].x = fmaf(acc[i].x, a[i & 3].x, a[(i + 1) & 3].y);
                acc[i].y = fmaf(acc[i].y, a[(i + 2) & 3].x, a[(i + 3) & 3].y);
            } else if (MODE == 1) {
                acc[i] = fma2(acc[i], a[i & 3], a[(i + 1) & 3]);
            } else if (MODE == 2) {
                acc[i] = fma2(a[i & 3], a[(i + 1) & 3], acc[i]);
            } else if (MODE == 4) {            // both multiplicands shared by every instruction (max operand reuse)
                acc[i] = fma2(a[0], a[1], acc[i]);
            } else if (MODE == 5) {            // one multiplicand shared by runs of 4, the other changes
                acc[i] = fma2(a[(i >> 2) & 3], a[i & 3], acc[i]);
            } else if (MODE == 6) {            // scalar, both multiplicands shared
                acc[i].x = fmaf(a[0].x, a[1].x, acc[i].x);
                acc[i].y = fmaf(a[0].y, a[1].y, acc[i].y);
            } else if (MODE == 7) {            // GEMM form: pair w shared by runs of 4, broadcast scalar activation changes
                acc[i] = fma2b(a[(i >> 2) & 3], (i & 1) ? a[(i >> 1) & 1].x : a[(i >> 1) & 1].y, acc[i]);
            } else if (MODE == 8) {            // GEMM form transposed: broadcast scalar shared by runs of 4, pair changes
                acc[i] = fma2b(a[i & 3], (i & 4) ? a[(i >> 3) & 3].x : a[(i >> 3) & 3].y, acc[i]);
            } else if (MODE == 9) {            // scalar, one multiplicand shared by runs of 4
                acc[i].x = fmaf(a[(i >> 2) & 3].x, a[i & 3].y, acc[i].x);
                acc[i].y = fmaf(a[(i >> 2) & 3].x, a[(i + 1) & 3].y, acc[i].y);
            } else {
                acc[i] = fma2(a[i & 3], a[(i + 1) & 3], acc[i]);
                if ((i & 7) == 7) {
                    const float s = acc[i].x;
                    m2 = fminf(m2, fmaxf(m1, s));
                    m1 = fminf(m1, s);
                }
            }
        }
    }
    const long long t1 = clock64();
    float r = m1 + m2;
    for (int i = 0; i < NACC; ++i) r += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

// The GRU gate GEMM inner block of fpc_encode_fp32.cu without its loads: 4 k x 3 gates x 7 rows of
// acc[gate][row] += w[k][gate] (packed pair) * a[row][k] (broadcast scalar).  ORDER 0: rows innermost (the pair is
// shared by 7 consecutive FFMA2); ORDER 1: gates innermost (the scalar is shared by 3); DEP: chain length test.
template <int ORDER>
__global__ void gemm_block(float *out, long long *cycles, const float4 *src)
{
    float2 acc[3][7];
    float4 a[7], w[2][3];
    for (int i = 0; i < 7; ++i) a[i] = src[i + threadIdx.x % 3];
    for (int p = 0; p < 2; ++p) for (int g = 0; g < 3; ++g) w[p][g] = src[8 + p * 3 + g + threadIdx.x % 5];
    for (int g = 0; g < 3; ++g) for (int i = 0; i < 7; ++i) acc[g][i] = make_float2(g, i);
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < kIters; ++it) {
        if (ORDER == 0) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
#pragma unroll
                for (int g = 0; g < 3; ++g) {
                    const float4 wv = w[kk >> 1][g];
                    const float2 wp = (kk & 1) ? make_float2(wv.z, wv.w) : make_float2(wv.x, wv.y);
#pragma unroll
                    for (int i = 0; i < 7; ++i) {
                        const float av = kk == 0 ? a[i].x : kk == 1 ? a[i].y : kk == 2 ? a[i].z : a[i].w;
                        acc[g][i] = fma2b(wp, av, acc[g][i]);
                    }
                }
        } else {
#pragma unroll
            for (int i = 0; i < 7; ++i)
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    const float av = kk == 0 ? a[i].x : kk == 1 ? a[i].y : kk == 2 ? a[i].z : a[i].w;
#pragma unroll
                    for (int g = 0; g < 3; ++g) {
                        const float4 wv = w[kk >> 1][g];
                        const float2 wp = (kk & 1) ? make_float2(wv.z, wv.w) : make_float2(wv.x, wv.y);
                        acc[g][i] = fma2b(wp, av, acc[g][i]);
                    }
                }
        }
    }
    const long long t1 = clock64();
    float r = 0.f;
    for (int g = 0; g < 3; ++g) for (int i = 0; i < 7; ++i) r += acc[g][i].x + acc[g][i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

// dependent-issue latency: NCH independent chains, each instruction depends on the one NCH before it; fully unrolled
template <int NCH, bool PACKED>
__global__ void latency(float *out, long long *cycles, float s0, float s1)
{
    float2 acc[NCH];
    for (int i = 0; i < NCH; ++i) acc[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f);
    const float2 w = make_float2(s0, s1);
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < 64; ++it) {
#pragma unroll
        for (int u = 0; u < 64; ++u) {
#pragma unroll
            for (int i = 0; i < NCH; ++i) {
                if (PACKED) acc[i] = fma2b(w, s1, acc[i]);
                else acc[i].x = fmaf(s0, s1, acc[i].x);
            }
        }
    }
    const long long t1 = clock64();
    float r = 0.f;
    for (int i = 0; i < NCH; ++i) r += acc[i].x + acc[i].y;
    out[threadIdx.x] = r;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}
template <int NCH, bool PACKED>
void run_latency()
{
    float *out; long long *cyc, h;
    cudaMalloc(&out, 1024 * 4); cudaMalloc(&cyc, 8);
    latency<NCH, PACKED><<<1, 32>>>(out, cyc, 1.0001f, 0.9999f);
    cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%s, %2d independent chains, one warp: %.2f cycles per instruction -> chain step every %.1f cycles\n",
           PACKED ? "FFMA2" : "FFMA ", NCH, (double)h / (64.0 * 64 * NCH), (double)h / (64.0 * 64));
    cudaFree(out); cudaFree(cyc);
}

template <int ORDER>
void run_block(const char *name, int warps_per_sched)
{
    float *out; long long *cyc, h; float4 *src;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8); cudaMalloc(&src, 64 * 16);
    cudaMemset(src, 0, 64 * 16);
    const int threads = 128 * warps_per_sched;
    gemm_block<ORDER><<<148, threads>>>(out, cyc, src);
    gemm_block<ORDER><<<148, threads>>>(out, cyc, src);
    cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-28s warps/sched=%d : %.3f cycles per FFMA2\n", name, warps_per_sched,
           (double)h / ((double)kIters * 84 * warps_per_sched));
    cudaFree(out); cudaFree(cyc); cudaFree(src);
}

template <int MODE, int NACC>
void run(const char *name, int warps_per_sched)
{
    float *out; long long *cyc, h;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    const int threads = 128 * warps_per_sched;
    bench<MODE, NACC><<<148, threads>>>(out, cyc, 1.0001f, 0.9999f);
    bench<MODE, NACC><<<148, threads>>>(out, cyc, 1.0001f, 0.9999f);
    cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double per = (double)h / ((double)kIters * NACC * warps_per_sched);
    printf("%-28s NACC=%2d warps/sched=%d : %.3f cycles per warp-instruction-slot (%s)\n", name, NACC, warps_per_sched, per,
           (MODE == 0 || MODE == 6 || MODE == 9) ? "2 FFMA per slot" : "1 FFMA2 per slot");
    cudaFree(out); cudaFree(cyc);
}

int main()
{
    for (int w = 1; w <= 4; w *= 2) {
        run<0, 16>("scalar FFMA x2", w);
        run<1, 16>("FFMA2 acc as multiplicand", w);
        run<2, 16>("FFMA2 acc as addend", w);
        run<2, 32>("FFMA2 acc as addend", w);
        run<3, 32>("FFMA2 + top-2 every 8", w);
        run<4, 32>("FFMA2 both mult shared", w);
        run<5, 32>("FFMA2 one mult shared x4", w);
        run<6, 16>("scalar both mult shared", w);
        run<9, 16>("scalar one mult shared x4", w);
        run<7, 32>("FFMA2 bcast, pair shared x4", w);
        run<8, 32>("FFMA2 bcast shared x4", w);
    }
    for (int w = 1; w <= 2; ++w) {
        run_block<0>("GEMM block, rows innermost", w);
        run_block<1>("GEMM block, gates innermost", w);
    }
    run_latency<1, true>(); run_latency<2, true>(); run_latency<4, true>(); run_latency<8, true>(); run_latency<16, true>();
    run_latency<1, false>(); run_latency<2, false>(); run_latency<4, false>(); run_latency<8, false>();
    run<2, 1>("FFMA2 dependent chain", 1);
    run<2, 2>("FFMA2 2 chains", 1);
    run<2, 4>("FFMA2 4 chains", 1);
    run<2, 8>("FFMA2 8 chains", 1);
    return 0;
}
