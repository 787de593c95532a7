// fpc_vq_search.cuh -- block-wide m-best VQ search (256 threads), used inside the fused
// frame-step kernel and by the stand-alone quantiser kernel.
// Reference: /root/reference/src/quantization/vq_func.py:82-131 (quantize_mstage) and :10-24
// (vq_quantize_mbest).  For two stages the reference's survivor merge reduces to
//   argmin over (survivor rank s, stage-1 entry j) of || (x - CB0[k_s]) - CB1[j] ||^2
// with ties going to the earlier survivor rank, then the lower j (strict < at :115-118 and
// the stable sorted() at :20); the 5 survivors k_s are the stage-0 entries with the smallest
// (distance, index).  tests/test_oracle_golden.py pins this against the reference itself.
#pragma once
#include "fpc_vq.cuh"

namespace fpc {

constexpr int kComputeThreads = 256;
constexpr int kLdR = 24;   // residual row: c0 at [3], c1..c17 at [4..20] (float4 aligned)

template <typename T> struct VqPart { T d; int i; int pad; };

// scratch layout for up to `maxn` rows:  dl [5*maxn][20] T | part [5*maxn][8] | surv [maxn][5] | dbuf [vb][1024] T
template <typename T> __host__ __device__ constexpr size_t vq_fixed_bytes(int maxn)
{
    return (size_t)5 * maxn * 20 * sizeof(T) + (size_t)5 * maxn * 8 * sizeof(VqPart<T>) + (size_t)((maxn * 5 * 4 + 15) / 16) * 16;
}

// thread t of NT owns codewords t + NT * s, s = 0 .. 1024/NT - 1, held four slots (one "chunk") at a time
template <typename T, int NT>
__device__ __forceinline__ void load_codewords(T (&cw)[4][kDim], const T *__restrict__ cbt, int K, int tid, int ch)
{
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int c = tid + NT * (4 * ch + q);
#pragma unroll
        for (int d = 0; d < kDim; ++d) cw[q][d] = c < K ? cbt[(size_t)d * K + c] : (T)0;
    }
}

template <typename T>
__device__ __forceinline__ void load_row_vector(T (&x)[kDim], const float *__restrict__ xr)
{
#pragma unroll
    for (int d4 = 0; d4 < 4; ++d4) {
        const float4 f = *reinterpret_cast<const float4 *>(xr + 4 * d4);
        x[4 * d4 + 0] = (T)f.x; x[4 * d4 + 1] = (T)f.y; x[4 * d4 + 2] = (T)f.z; x[4 * d4 + 3] = (T)f.w;
    }
    x[16] = (T)xr[16];
}

// nearest codeword of `x` among this thread's codewords, merged across the warp; lane 0
// records the warp's (distance, index) pair
template <typename T, int NT>
__device__ __forceinline__ void search_own(const T (&x)[kDim], const T (&cw)[4][kDim], int nq, int K, int tid, int ch,
                                           VqPart<T> *__restrict__ out)
{
    T bd = Rn<T>::inf();
    int bi = 0x7fffffff;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        if (4 * ch + q < nq) {
            const int c = tid + NT * (4 * ch + q);
            const T d = dist17<T>(x, cw[q]);
            if (c < K && (d < bd || bi == 0x7fffffff)) { bd = d; bi = c; }
        }
    }
    warp_argmin(bd, bi);
    if ((tid & 31) == 0) { out->d = bd; out->i = bi; }
}

// merge the 8 partial results (warps x chunks) of one search; all lanes return the winner
template <typename T>
__device__ __forceinline__ void merge_parts(const VqPart<T> *__restrict__ p, int lane, T &bd, int &bi)
{
    bd = lane < 8 ? p[lane].d : Rn<T>::inf();
    bi = lane < 8 ? p[lane].i : 0x7fffffff;
    warp_argmin(bd, bi);
    if (bi == 0x7fffffff) bi = 0;
}

// ------------------------------------------------------------------------------------------
// VQ search over a list of rows, all 256 compute threads of the block.
//   bk:    1 or 2 stages of K codewords; thread t owns codewords t, t+256, t+512, t+768
//   list:  list[0..n) row ids; the search vector of a row is rs[row*kLdR + 4 .. +20]
//   out:   rq[row*20 + 1..17] (float), idx1[row], idx2[row]; optionally qglobal[row*17..] in the
//          codebook dtype (the stand-alone quantiser returns the file's dtype, vq_func.py:161-164)
//   scratch: vq_fixed_bytes<T>(maxn) + vb*1024*sizeof(T) bytes, vb >= 1
// Every thread of the 256 must call this with identical arguments (barrier id 1 is used).
// ------------------------------------------------------------------------------------------
template <typename T, int NT = kComputeThreads>
__device__ __noinline__ void vq_search_rows(const PackedVq &bk, const char *__restrict__ cbbase, const int *__restrict__ list, int n,
                               int maxn, const float *__restrict__ rs, float *__restrict__ rq, int *__restrict__ idx1,
                               int *__restrict__ idx2, char *__restrict__ scratch, int vb, int tid,
                               T *__restrict__ qglobal = nullptr)
{
    constexpr int NW = NT / 32;            // warps
    constexpr int NCH = 256 / NT;          // chunks of four codeword slots per thread (NW * NCH = 8 partials per search)
    static_assert(NT == 256 || NT == 128, "block-wide search is written for 256 or 128 threads");
    const int warp = tid >> 5, lane = tid & 31;
    const int K = bk.K;
    const int nq = (K + NT - 1) / NT;   // codeword slots per thread in use
    const T *cbt0 = reinterpret_cast<const T *>(cbbase + bk.off_t[0]);
    const T *cbr0 = reinterpret_cast<const T *>(cbbase + bk.off_r[0]);
    T *dl = reinterpret_cast<T *>(scratch);
    VqPart<T> *part = reinterpret_cast<VqPart<T> *>(scratch + (size_t)5 * maxn * 20 * sizeof(T));
    int *surv = reinterpret_cast<int *>(reinterpret_cast<char *>(part) + (size_t)5 * maxn * 8 * sizeof(VqPart<T>));
    T *dbuf = reinterpret_cast<T *>(scratch + vq_fixed_bytes<T>(maxn));

    T cw[4][kDim];

    if (bk.stages == 1) {
#pragma unroll 1
        for (int ch = 0; ch < NCH; ++ch) {
            load_codewords<T, NT>(cw, cbt0, K, tid, ch);
            for (int v = 0; v < n; ++v) {
                T x[kDim];
                load_row_vector<T>(x, rs + list[v] * kLdR + 4);
                search_own<T, NT>(x, cw, nq, K, tid, ch, &part[v * 8 + ch * NW + warp]);
            }
        }
        named_bar_sync(1, NT);
        for (int v = warp; v < n; v += NW) {
            T bd; int bi;
            merge_parts<T>(&part[v * 8], lane, bd, bi);
            const int row = list[v];
            if (lane < kDim) {
                const T csum = Rn<T>::add((T)0, cbr0[(size_t)bi * kDim + lane]);   // csum = 0; csum += CB[0][i]
                rq[row * 20 + 1 + lane] = (float)csum;
                if (qglobal) qglobal[(size_t)row * kDim + lane] = csum;
            }
            if (lane == 0) { idx1[row] = bi; idx2[row] = -1; }
        }
        named_bar_sync(1, NT);
        return;
    }

    // ---- stage 0: distances to all K entries, 5 best per vector, stage-1 search vectors ----
    for (int base = 0; base < n; base += vb) {
        const int nb = min(vb, n - base);
#pragma unroll 1
        for (int ch = 0; ch < NCH; ++ch) {
            load_codewords<T, NT>(cw, cbt0, K, tid, ch);
            for (int v = 0; v < nb; ++v) {
                T x[kDim];
                load_row_vector<T>(x, rs + list[base + v] * kLdR + 4);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (4 * ch + q < nq) {
                        const int c = tid + NT * (4 * ch + q);
                        const T d = dist17<T>(x, cw[q]);
                        dbuf[v * 1024 + c] = c < K ? d : Rn<T>::inf();
                    }
                }
            }
        }
        named_bar_sync(1, NT);
        for (int v = warp; v < nb; v += NW) {
            const int nj = (nq * NT) >> 5;
            T loc[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) loc[j] = j < nj ? dbuf[v * 1024 + lane + 32 * j] : Rn<T>::inf();
            const float *xr = rs + list[base + v] * kLdR + 4;
            for (int s = 0; s < kSurv; ++s) {
                T bd = Rn<T>::inf();
                int bj = 0;
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (loc[j] < bd) { bd = loc[j]; bj = j; }
                T wd = bd;
                int wi = (bd < Rn<T>::inf()) ? lane + 32 * bj : 0x7fffffff;
                warp_argmin(wd, wi);
                if (wi == 0x7fffffff) wi = 0;   // fewer than 5 finite distances (NaN/inf input)
                if ((wi & 31) == lane) {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (j == (wi >> 5)) loc[j] = Rn<T>::inf();
                }
                if (lane == 0) surv[(base + v) * kSurv + s] = wi;
                if (lane < kDim) {   // diff = x - (0 + CB[0][wi])   (vq_func.py:103-106)
                    const T csum = Rn<T>::add((T)0, cbr0[(size_t)wi * kDim + lane]);
                    dl[((base + v) * kSurv + s) * 20 + lane] = Rn<T>::sub((T)xr[lane], csum);
                }
            }
        }
        named_bar_sync(1, NT);
    }

    // ---- stage 1: nearest entry for every (vector, survivor) pair ----
    const T *cbt1 = reinterpret_cast<const T *>(cbbase + bk.off_t[1]);
    const T *cbr1 = reinterpret_cast<const T *>(cbbase + bk.off_r[1]);
#pragma unroll 1
    for (int ch = 0; ch < NCH; ++ch) {
        load_codewords<T, NT>(cw, cbt1, K, tid, ch);
        for (int p = 0; p < n * kSurv; ++p) {
            const T *xr = dl + p * 20;
            T x[kDim];
#pragma unroll
            for (int d = 0; d < kDim; ++d) x[d] = xr[d];
            search_own<T, NT>(x, cw, nq, K, tid, ch, &part[p * 8 + ch * NW + warp]);
        }
    }
    named_bar_sync(1, NT);
    // ---- merge: strict < across survivor ranks keeps the earlier rank on ties (:115-125) ----
    for (int v = warp; v < n; v += NW) {
        T fd = Rn<T>::inf();
        int fs = 0, fi = 0;
        for (int s = 0; s < kSurv; ++s) {
            T bd; int bi;
            merge_parts<T>(&part[(v * kSurv + s) * 8], lane, bd, bi);
            if (s == 0 || bd < fd) { fd = bd; fs = s; fi = bi; }
        }
        const int i0 = surv[v * kSurv + fs];
        const int row = list[v];
        if (lane < kDim) {
            T csum = Rn<T>::add((T)0, cbr0[(size_t)i0 * kDim + lane]);
            csum = Rn<T>::add(csum, cbr1[(size_t)fi * kDim + lane]);
            rq[row * 20 + 1 + lane] = (float)csum;
            if (qglobal) qglobal[(size_t)row * kDim + lane] = csum;
        }
        if (lane == 0) { idx1[row] = i0; idx2[row] = fi; }
    }
    named_bar_sync(1, NT);
}

// scratch_bytes is what the caller really has; vb (vectors per stage-0 batch) follows from it
__device__ __forceinline__ void vq_dispatch(const PackedVq &bk, const char *cbbase, const int *list, int n, int maxn,
                                            const float *rs, float *rq, int *idx1, int *idx2, char *scratch,
                                            int scratch_bytes, int tid)
{
    if (bk.dtype == FPC_F32) {
        int vb = (scratch_bytes - (int)vq_fixed_bytes<float>(maxn)) / (1024 * 4);
        vb = vb > 8 ? 8 : vb;
        vq_search_rows<float>(bk, cbbase, list, n, maxn, rs, rq, idx1, idx2, scratch, vb, tid);
    } else {
        int vb = (scratch_bytes - (int)vq_fixed_bytes<double>(maxn)) / (1024 * 8);
        vb = vb > 8 ? 8 : vb;
        vq_search_rows<double>(bk, cbbase, list, n, maxn, rs, rq, idx1, idx2, scratch, vb, tid);
    }
}

}  // namespace fpc
