// fpc_vq.cuh -- device-side quantiser arithmetic shared by the fused frame-step kernel and the
// stand-alone quantiser kernels.
//
// Reference: /root/reference/src/quantization/vq_func.py:10-24 (vq_quantize_mbest),
// :82-131 (quantize_mstage), :167-185 (scl_quantize).  numpy evaluates
// np.sum((x - cb) ** 2, -1) as t = x - c, e = t * t (each rounded, no fma) followed by its
// pairwise reduction over the contiguous axis: for 17 terms, eight interleaved partial sums
// r[j] = e[j] + e[8+j], ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), then + e[16].  The functions
// here perform exactly those roundings, in the codebook's dtype, so codeword indices are
// bit-identical to the reference's, ties included (lowest index wins: Python's stable
// sorted() and np.argmin).
#pragma once
#include "fpc_common.cuh"

namespace fpc {

template <typename T> struct Rn;
template <> struct Rn<float> {
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float inf() { return __int_as_float(0x7f800000); }
};
template <> struct Rn<double> {
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double inf() { return __longlong_as_double(0x7ff0000000000000LL); }
};

// squared distance of x to codeword c over the 17 code dimensions, numpy order
template <typename T>
__device__ __forceinline__ T dist17(const T (&x)[kDim], const T (&c)[kDim])
{
    T e[kDim];
#pragma unroll
    for (int d = 0; d < kDim; ++d) {
        T t = Rn<T>::sub(x[d], c[d]);
        e[d] = Rn<T>::mul(t, t);
    }
    T r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = Rn<T>::add(e[j], e[8 + j]);
    T res = Rn<T>::add(Rn<T>::add(Rn<T>::add(r[0], r[1]), Rn<T>::add(r[2], r[3])),
                       Rn<T>::add(Rn<T>::add(r[4], r[5]), Rn<T>::add(r[6], r[7])));
    return Rn<T>::add(res, e[16]);
}

// lexicographic (distance, index) minimum across the warp; every lane gets the result
template <typename T>
__device__ __forceinline__ void warp_argmin(T &d, int &i)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        T od = __shfl_xor_sync(0xffffffffu, d, off);
        int oi = __shfl_xor_sync(0xffffffffu, i, off);
        if (od < d || (od == d && oi < i)) { d = od; i = oi; }
    }
}

// fp32 fast path of the same reduction: distances are non-negative, so their bit patterns
// order like unsigned integers and two REDUX.MIN instructions replace the shuffle tree.
__device__ __forceinline__ void warp_argmin(float &d, int &i, int /*tag*/)
{
    unsigned db = __float_as_uint(d);
    unsigned m = __reduce_min_sync(0xffffffffu, db);
    unsigned cand = db == m ? (unsigned)i : 0xffffffffu;
    unsigned w = __reduce_min_sync(0xffffffffu, cand);
    d = __uint_as_float(m);
    i = (int)w;
}

// scalar quantiser for one value by one warp (vq_func.py:175-176): argmin over n codes of
// (x - code)^2, first minimum.  codes in global or shared memory (n <= 256 = 8 per lane: the loads are issued
// together, so one memory latency is exposed instead of eight).
template <typename T>
__device__ __forceinline__ int warp_scl_nearest(const T *__restrict__ codes, int n, float x, int lane, T &q)
{
    static_assert(FPC_MAX_SCL_ENTRIES == 256, "eight codes per lane");
    T c[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) c[j] = lane + 32 * j < n ? codes[lane + 32 * j] : (T)0;
    T best = Rn<T>::inf();
    int bi = 0x7fffffff;
    const T xv = (T)x;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int k = lane + 32 * j;
        const T t = Rn<T>::sub(xv, c[j]);
        const T d = Rn<T>::mul(t, t);
        if (k < n && (d < best || bi == 0x7fffffff)) { best = d; bi = k; }  // ascending k per lane: first min kept
    }
    warp_argmin(best, bi);
    if (bi == 0x7fffffff) bi = 0;
    q = codes[bi];
    return bi;
}

// The same search by a group of `G` consecutive lanes (G = 4 or 8; every lane of the warp must call): lane `part` of
// the group scans a contiguous quarter / eighth of the table in ascending order with strict < (first minimum), the group
// then keeps the lexicographically smallest (distance, index).  Same arithmetic, same result as warp_scl_nearest; 32 / G
// values are quantised per warp at once instead of one.
template <typename T, int G>
__device__ __forceinline__ int group_scl_nearest(const T *__restrict__ codes, int n, float x, int part, T &q)
{
    const int per = (n + G - 1) / G;
    const int k0 = part * per, k1 = min(n, k0 + per);
    const T xv = (T)x;
    T best = Rn<T>::inf();
    int bi = 0x7fffffff;
    for (int k = k0; k < k1; ++k) {
        const T t = Rn<T>::sub(xv, codes[k]);
        const T d = Rn<T>::mul(t, t);
        if (d < best || bi == 0x7fffffff) { best = d; bi = k; }
    }
#pragma unroll
    for (int off = 1; off < G; off <<= 1) {
        const T od = __shfl_xor_sync(0xffffffffu, best, off);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
        if (oi != 0x7fffffff && (bi == 0x7fffffff || od < best || (od == best && oi < bi))) { best = od; bi = oi; }
    }
    if (bi == 0x7fffffff) bi = 0;
    q = codes[bi];
    return bi;
}

}  // namespace fpc
