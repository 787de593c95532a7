// fpc_tc.cuh -- nearest-codeword SCREENING on the 5th-generation tensor cores.
//
// The searches of this library (quantize_mstage, /root/reference/src/quantization/vq_func.py:82-131, and find_nearest,
// quantization/cb_func.py:56-68) are contractions  score[v][k] = ||c_k||^2 - 2 <x_v, c_k>  (N x 17 . 17 x K) followed by a
// minimum.  The decision itself must be the reference's (float64 / float32 direct-form distances in numpy's rounding
// order), so the tensor cores only SCREEN: they produce every score to within a proven error, a scan of the
// accumulators finds the smallest score and the runner-up of every vector, and a vector whose runner-up is more than
// a margin away is decided; the others go to the exact code paths.
//
// Operand format.  tcgen05.mma.kind::f16 multiplies fp16 pairs exactly (11 + 11 significand bits fit fp32) and
// accumulates in fp32, so a 22-bit value is carried as a (hi, lo) pair of fp16 and a product as three terms:
//     x c  ~  xh ch + xh cl + xl ch          (dropped: xl cl <= 2^-22 |x c|)
// Per dimension d the K extent holds   A side (vectors): xh, xh, xl    B side (codewords): ch, cl, ch
// i.e. 51 entries for 17 dimensions; entries 51..53 carry the norm: A side 1, 1, 1; B side the 3-way fp16 split of
// beta^2 ||c||^2; entries 54..56 may carry a per-vector offset (A side: 3-way split of offset / 1024, B side: 1024) that
// makes every score of a vector positive, which the key-packed top-k scan of the m-best search needs; entries 57..63 are
// zero.  K = 64 = four tcgen05 K-slabs, 128 bytes per row.
// Values are scaled by a power of two beta (exact) so that the largest |c| lies in [4, 8): fp16 then keeps an absolute
// resolution of 2^-25 (subnormal hi/lo parts included), far below the 2^-22 relative target at that scale, and the
// vector side  -2 beta x  may be up to 1024 (an x 128 times the largest codeword entry) before it leaves the format --
// such rows are handed to the exact path.  Scores come out scaled by beta^2.
//
// Error of a screened score against the real value, with u = 2^-24 and R = (||x|| + Cmax)^2:
//     fp16 pair representation of x and c:   2 * 2^-22 ||x|| ||c|| * 2   <=  4 u R       (||x|| ||c|| <= R / 4)
//     dropped lo * lo term:                   2^-22 ||x|| ||c|| * 2       <=  2 u R
//     absolute 2^-25 floor of the format:     <= 2 u R at the chosen scale
//     fp32 rounding of beta c (float64 books) and the 3-way norm split:   <=  2 u R
//     accumulation inside the tensor core:    measured by fpc_selftest_tc_scores against float64
// The margins used by the callers (2^-15 R) leave two orders of magnitude for the last item.
//
// Tile layout in shared memory: "K-major, no swizzle" (fpc_umma.cuh) with R = 128 rows per tile,
//     byte offset(r, k) = (k / 8) * 2048 + r * 16 + (k % 8) * 2,   16 KB per 128 x 64 tile, 4 KB per K-slab.
#pragma once
#include <cuda_fp16.h>

#include "fpc_common.cuh"
#include "fpc_umma.cuh"

namespace fpc {
namespace tc {

constexpr int kK = 64;                         // K extent of both operands
constexpr int kTileRows = 128;
constexpr int kTileBytes = kTileRows * kK * 2; // 16384
constexpr int kSlabBytes = kTileBytes / 4;     // one K = 16 slab of a tile
constexpr float kMaxScaledX = 1024.0f;         // |(-2 beta x)_d| beyond this leaves the proven range
constexpr float kPadNorm = 65504.0f;           // norm entries of padding codewords (three of them: 196 512, above any real score)

// instruction descriptor: fp16 x fp16 -> f32, A and B K-major, dense
__host__ __device__ constexpr uint32_t instr_desc_f16(uint32_t M, uint32_t N)
{
    return (1u << 4) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    umma::mma_bf16(tmem_d, adesc, bdesc, idesc, accumulate);      // same instruction; the operand type lives in idesc
}

// beta = 2^e with  4 <= beta * cmax_abs < 8  (cmax_abs = largest |c| entry of the book; 1 for an all-zero book)
__host__ __device__ inline float scale_for(float cmax_abs)
{
    if (!(cmax_abs > 0.0f) || !(cmax_abs < 1e30f)) return 1.0f;
    int e;
    frexpf(cmax_abs, &e);                 // cmax_abs = m * 2^e, m in [0.5, 1)
    return ldexpf(1.0f, 3 - e);           // beta * cmax_abs = 8 m in [4, 8)
}

__device__ __forceinline__ uint32_t pack_h2(__half a, __half b)
{
    return (uint32_t)__half_as_ushort(a) | ((uint32_t)__half_as_ushort(b) << 16);
}

// (hi, lo) fp16 pair of an fp32 value: hi = rn(v), lo = rn(v - hi) (the subtraction is exact)
__device__ __forceinline__ void split2(float v, __half &hi, __half &lo)
{
    hi = __float2half_rn(v);
    lo = __float2half_rn(__fsub_rn(v, __half2float(hi)));
}

// Writes row r of a kRows-row operand tile from the 17 scaled values xs (vector side: -2 beta x; codeword side: beta c),
// the three norm entries 51..53 (vector side: 1, 1, 1; codeword side: split of beta^2 ||c||^2) and the three offset
// entries 54..56 (o0..o2; callers that add no per-vector offset pass zeros).
template <bool kVectorSide, int kRows = kTileRows>
__device__ __forceinline__ void store_row(unsigned char *tile, int r, const float (&xs)[kDim], __half n0, __half n1, __half n2,
                                          __half o0 = __ushort_as_half(0), __half o1 = __ushort_as_half(0), __half o2 = __ushort_as_half(0))
{
    __half e[kK];
#pragma unroll
    for (int d = 0; d < kDim; ++d) {
        __half hi, lo;
        split2(xs[d], hi, lo);
        e[3 * d] = hi;
        e[3 * d + 1] = kVectorSide ? hi : lo;
        e[3 * d + 2] = kVectorSide ? lo : hi;
    }
    e[51] = n0; e[52] = n1; e[53] = n2;
    e[54] = o0; e[55] = o1; e[56] = o2;
#pragma unroll
    for (int i = 57; i < kK; ++i) e[i] = __float2half_rn(0.0f);
#pragma unroll
    for (int c = 0; c < kK / 8; ++c) {
        uint4 w;
        w.x = pack_h2(e[8 * c], e[8 * c + 1]); w.y = pack_h2(e[8 * c + 2], e[8 * c + 3]);
        w.z = pack_h2(e[8 * c + 4], e[8 * c + 5]); w.w = pack_h2(e[8 * c + 6], e[8 * c + 7]);
        *reinterpret_cast<uint4 *>(tile + (size_t)c * (kRows * 16) + (size_t)r * 16) = w;
    }
}

// 3-way fp16 split of a non-negative fp32 value below 65504 * (1 + 2^-11 + 2^-22)
__device__ __forceinline__ void split3(float v, __half &a, __half &b, __half &c)
{
    a = __float2half_rn(v);
    const float r1 = __fsub_rn(v, __half2float(a));
    b = __float2half_rn(r1);
    c = __float2half_rn(__fsub_rn(r1, __half2float(b)));
}

// ---- TMEM -> registers, 32 columns of this thread's lane ----
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
// the registers of a load are undefined until tcgen05.wait::ld; naming them as read-write operands keeps every use below
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&a)[32])
{
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) asm volatile("" : "+r"(a[i]));
}

__device__ __forceinline__ float fmin3(float a, float b, float c) { return fminf(fminf(a, b), c); }

// minimum of 16 values (8 instructions with 3-input minima)
__device__ __forceinline__ float min16(const uint32_t *v)
{
    float t[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) t[i] = fmin3(__uint_as_float(v[3 * i]), __uint_as_float(v[3 * i + 1]), __uint_as_float(v[3 * i + 2]));
    return fminf(fmin3(t[0], t[1], t[2]), fmin3(t[3], t[4], __uint_as_float(v[15])));
}
__device__ __forceinline__ float min16f(const float (&v)[16])
{
    float t[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) t[i] = fmin3(v[3 * i], v[3 * i + 1], v[3 * i + 2]);
    return fminf(fmin3(t[0], t[1], t[2]), fmin3(t[3], t[4], v[15]));
}

// Running state of one thread's scan of its vector's scores: the smallest value and the runner-up are found
// WITHOUT carrying indices through the 1000+ comparisons.  Every column belongs to one "group" (16 consecutive
// columns) and one "class" (its position inside the group); two columns never share both.  The scan keeps the two
// smallest group minima (with the group of the smallest) and the minimum of every class.  The smallest score is the
// smallest group minimum; its column is (that group, the class holding the same value); the runner-up is
//     min(second smallest group minimum, second smallest class minimum):
// every other column lies outside the winner's group or outside its class, so both terms are minima over columns
// other than the winner's, and the true runner-up is covered by one of them.  About 1.3 min instructions per score
// instead of 4 with packed indices, and no index arithmetic in the loop.
struct Scan {
    float a1, a2;       // two smallest group minima
    int ga;             // group of a1
    float cls[16];      // class minima
    __device__ __forceinline__ void reset()
    {
        a1 = a2 = __int_as_float(0x7f800000);
        ga = 0;
#pragma unroll
        for (int j = 0; j < 16; ++j) cls[j] = __int_as_float(0x7f800000);
    }
    __device__ __forceinline__ void group(float g, int gid)
    {
        a2 = fminf(a2, fmaxf(a1, g));
        ga = g < a1 ? gid : ga;
        a1 = fminf(a1, g);
    }
    // 32 columns = two groups (ids gid, gid + 1)
    __device__ __forceinline__ void feed(const uint32_t (&v)[32], int gid)
    {
        group(min16(v), gid);
        group(min16(v + 16), gid + 1);
#pragma unroll
        for (int j = 0; j < 16; ++j) cls[j] = fmin3(cls[j], __uint_as_float(v[j]), __uint_as_float(v[j + 16]));
    }
    // smallest score, runner-up, class of the smallest (short dependency chains: 16 classes, all in parallel)
    __device__ __forceinline__ void finish(float &best, float &second, int &jbest) const
    {
        const float inf = __int_as_float(0x7f800000);
        const float b1 = min16f(cls);
        int jb = 0;
#pragma unroll
        for (int j = 0; j < 16; ++j) jb = cls[j] == b1 ? j : jb;
        float o[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) o[j] = j == jb ? inf : cls[j];
        best = a1;                       // == b1
        second = fminf(a2, min16f(o));
        jbest = jb;
    }
};

}  // namespace tc
}  // namespace fpc
