// fpc_kmeans_ordered.cu -- per-centroid float64 sums in DATA ORDER: the accumulation loop of cb_func.update
// (/root/reference/src/quantization/cb_func.py:82-86)
//
//     for i in range(nb_vectors):  n = nearest[i];  count[n] += 1;  sum[n] += data[i]
//
// adds the vectors of a centroid one after the other, in float64, in the order they stand in `data`.  The assign
// kernels accumulate with float64 atomics in scheduling order -- the same sums to ~1e-16 relative, but not the same
// bits, and not the same bits from run to run.  This file is the opt-in exact variant: a STABLE counting sort of the
// row numbers by centroid (three HBM-bound integer passes) and then one warp per centroid that walks its rows in
// ascending order, lanes = the 17 dimensions.  With the indices of fpc_kmeans_assign_accumulate (bit-identical to
// NumPy's float64 argmin) the updated codebook then equals the reference's bit for bit.
//
//   ord_count_kernel    a warp per tile of kTileRows rows: per-centroid counts of the tile          -> cnt[t][k]
//   ord_group_kernel    exclusive scan of cnt over the tiles of a group, group totals              -> gtot[g][k]
//   ord_offsets_kernel  scan of gtot over the groups, centroid sizes, centroid start offsets       -> coff[k]
//   ord_scatter_kernel  a warp per tile again: row i goes to perm[coff + group + tile offset + rank inside the tile]
//   ord_sum_kernel      a warp per centroid: sums[k][d] += data[perm[i]][d] for i ascending, counts[k] += size
//
// Algorithmic bytes per vector: idx 4 B read twice, perm 4 B written and read, the vector (68 B float32) read once
// = 84 B, plus K * 4 B of tile counts per 2048 rows (2 B per vector at K = 1024), written once and read twice.
#include <cstdlib>

#include "fpc_common.cuh"

namespace fpc {

constexpr int kOrdDim = 17;
constexpr int kOrdMaxK = 2048;
constexpr int kOrdTileRows = 2048;      // rows of one warp's tile
constexpr int kOrdWarps = 4;            // warps per CTA of the tile kernels (kOrdWarps * K counters in shared memory)
constexpr int kOrdGroups = 256;         // tile groups of the two-level scan

// acc += v[0] + ... + v[31] in order, in float64.  (Measured and dropped: widening float32 -> float64 on the integer
// pipe instead of F2F.F64.F32 -- 245 registers and 8.3 ms instead of 5.6.)
template <typename TD>
__device__ __forceinline__ void ord_add32(double &acc, const TD (&v)[32])
{
#pragma unroll
    for (int j = 0; j < 32; ++j) acc += (double)v[j];
}

struct OrdPlan {
    long tiles, tiles_per_group;
    size_t off_cnt, off_gtot, off_coff, off_cur, off_perm, bytes;
};

static OrdPlan ord_plan(long N, int K)
{
    OrdPlan p;
    p.tiles = (N + kOrdTileRows - 1) / kOrdTileRows;
    p.tiles_per_group = (p.tiles + kOrdGroups - 1) / kOrdGroups;
    if (p.tiles_per_group < 1) p.tiles_per_group = 1;
    size_t o = 0;
    p.off_cnt = o;  o += (size_t)p.tiles * K * sizeof(unsigned);          o = (o + 255) & ~(size_t)255;
    p.off_gtot = o; o += (size_t)kOrdGroups * K * sizeof(unsigned);       o = (o + 255) & ~(size_t)255;
    p.off_coff = o; o += (size_t)(K + 1) * sizeof(unsigned);              o = (o + 255) & ~(size_t)255;
    p.off_cur = o;  o += (size_t)K * sizeof(unsigned);                    o = (o + 255) & ~(size_t)255;
    p.off_perm = o; o += (size_t)N * sizeof(unsigned);                    o = (o + 255) & ~(size_t)255;
    p.bytes = o;
    return p;
}

__global__ void __launch_bounds__(kOrdWarps * 32) ord_count_kernel(const int *__restrict__ idx, long N, int K, long tiles,
                                                                   unsigned *__restrict__ cnt)
{
    extern __shared__ unsigned s_cnt[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned *mine = s_cnt + (size_t)warp * K;
    for (long t = (long)blockIdx.x * kOrdWarps + warp; t < tiles; t += (long)gridDim.x * kOrdWarps) {
        for (int k = lane; k < K; k += 32) mine[k] = 0u;
        __syncwarp();
        const long r0 = t * kOrdTileRows;
        for (int j = lane; j < kOrdTileRows; j += 32) {
            const long r = r0 + j;
            if (r < N) {
                const int key = idx[r];
                if (key >= 0 && key < K) atomicAdd(&mine[key], 1u);      // an index outside [0, K) belongs to no centroid
            }
        }
        __syncwarp();
        for (int k = lane; k < K; k += 32) cnt[(size_t)t * K + k] = mine[k];
        __syncwarp();
    }
}

// thread = (group, centroid): exclusive scan over the group's tiles, in place; the group's total to gtot
__global__ void ord_group_kernel(unsigned *__restrict__ cnt, long tiles, long tiles_per_group, int K,
                                 unsigned *__restrict__ gtot)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int g = blockIdx.y;
    if (k >= K) return;
    const long t0 = (long)g * tiles_per_group;
    long t1 = t0 + tiles_per_group;
    if (t1 > tiles) t1 = tiles;
    unsigned run = 0u;
    long t = t0;
    for (; t + 8 <= t1; t += 8) {           // eight loads in flight, then the eight dependent stores
        unsigned v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = cnt[(size_t)(t + j) * K + k];
#pragma unroll
        for (int j = 0; j < 8; ++j) { cnt[(size_t)(t + j) * K + k] = run; run += v[j]; }
    }
    for (; t < t1; ++t) {
        const unsigned v = cnt[(size_t)t * K + k];
        cnt[(size_t)t * K + k] = run;
        run += v;
    }
    gtot[(size_t)g * K + k] = run;
}

// one CTA: gtot -> exclusive over the groups (in place); centroid sizes -> exclusive start offsets coff[0..K]
__global__ void __launch_bounds__(1024) ord_offsets_kernel(unsigned *__restrict__ gtot, int K, unsigned *__restrict__ coff)
{
    __shared__ unsigned s_size[kOrdMaxK];
    __shared__ unsigned s_part[32];
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        unsigned run = 0u;
        for (int g = 0; g < kOrdGroups; ++g) {
            const unsigned v = gtot[(size_t)g * K + k];
            gtot[(size_t)g * K + k] = run;
            run += v;
        }
        s_size[k] = run;
    }
    __syncthreads();
    // exclusive scan of s_size[0..K): every thread owns kOrdMaxK / 1024 = 2 consecutive entries
    const int per = (K + (int)blockDim.x - 1) / (int)blockDim.x;
    const int k0 = threadIdx.x * per;
    unsigned local = 0u;
    for (int j = 0; j < per; ++j) if (k0 + j < K) local += s_size[k0 + j];
    unsigned inc = local;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int off = 1; off < 32; off <<= 1) {
        const unsigned o = __shfl_up_sync(0xffffffffu, inc, off);
        if (lane >= off) inc += o;
    }
    if (lane == 31) s_part[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        unsigned w = lane < (int)(blockDim.x >> 5) ? s_part[lane] : 0u;
        unsigned winc = w;
        for (int off = 1; off < 32; off <<= 1) {
            const unsigned o = __shfl_up_sync(0xffffffffu, winc, off);
            if (lane >= off) winc += o;
        }
        s_part[lane] = winc - w;
    }
    __syncthreads();
    unsigned run = s_part[warp] + inc - local;
    for (int j = 0; j < per; ++j) {
        if (k0 + j < K) { coff[k0 + j] = run; run += s_size[k0 + j]; }
    }
    if (k0 <= K - 1 && K - 1 < k0 + per) coff[K] = run;
}

__global__ void __launch_bounds__(kOrdWarps * 32) ord_scatter_kernel(const int *__restrict__ idx, long N, int K, long tiles,
                                                                     long tiles_per_group, const unsigned *__restrict__ cnt,
                                                                     const unsigned *__restrict__ gtot,
                                                                     const unsigned *__restrict__ coff,
                                                                     unsigned *__restrict__ perm)
{
    extern __shared__ unsigned s_cnt[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned *mine = s_cnt + (size_t)warp * K;
    const unsigned lt = (1u << lane) - 1u;
    for (long t = (long)blockIdx.x * kOrdWarps + warp; t < tiles; t += (long)gridDim.x * kOrdWarps) {
        const long g = t / tiles_per_group;
        for (int k = lane; k < K; k += 32) mine[k] = coff[k] + gtot[(size_t)g * K + k] + cnt[(size_t)t * K + k];
        __syncwarp();
        const long r0 = t * kOrdTileRows;
        int ahead = r0 + lane < N ? idx[r0 + lane] : -1;
        for (int j = 0; j < kOrdTileRows; j += 32) {        // 32 rows at a time, in order: the sort is stable
            const long r = r0 + j + lane;
            int key = ahead;
            if (j + 32 < kOrdTileRows) ahead = r + 32 < N ? idx[r + 32] : -1;       // the next step's keys, under this one
            if (key < 0 || key >= K) key = -1;
            const unsigned same = __match_any_sync(0xffffffffu, key);
            const int leader = __ffs(same) - 1;
            unsigned pos = 0u;
            if (key >= 0 && lane == leader) { pos = mine[key]; mine[key] = pos + __popc(same); }
            pos = __shfl_sync(0xffffffffu, pos, leader);
            if (key >= 0) perm[pos + __popc(same & lt)] = (unsigned)r;
            __syncwarp();
        }
    }
}

// One warp per centroid, lanes = dimensions.  The additions are one dependent float64 chain per centroid -- that IS the
// specification (8 cycles per DADD, tools/ubench_dadd.cu) -- but the loads are not.  Three stages run ahead of the adder:
// kOrdAhead chunks (of 32 rows) ahead every lane pulls one row into L2 (prefetch); NB - 1 chunks ahead the 32 vectors of
// a chunk are loaded into one of NB register buffers (the loop is unrolled over the buffers, so none is ever copied);
// U steps before that, their row numbers.  What is left is the gather itself: 68-byte rows at scattered addresses
// (measured: a deeper cp.async ring in shared memory is slower, 7.5 vs 5.3 ms for 50 M vectors -- the bound is DRAM
// access at sector granularity, not latency).
template <typename TD>
__global__ void __launch_bounds__(128) ord_sum_kernel(const TD *__restrict__ data, const unsigned *__restrict__ perm,
                                                      const unsigned *__restrict__ coff, int K, double *__restrict__ sums,
                                                      double *__restrict__ counts)
{
    const int lane = threadIdx.x & 31;
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int k = blockIdx.x * (blockDim.x >> 5) + warp;
    if (k >= K) return;
    // broadcast from lane 0: ptxas then knows the loop bounds are warp-uniform and emits the shuffles below without a
    // WARPSYNC / collective bracket around each of them (with the brackets: 8.7 ms instead of 5.3)
    const unsigned b = __shfl_sync(0xffffffffu, coff[k], 0), e = __shfl_sync(0xffffffffu, coff[k + 1], 0);
    const int d = lane < kOrdDim ? lane : 0;
    double acc = sums[(size_t)k * kOrdDim + d];
    constexpr unsigned kOrdAhead = 12;      // chunks between the L2 prefetch of a row and its load
    constexpr unsigned kRowsAhead = 24;     // chunks between the copy of a chunk's row numbers and their first use
    constexpr unsigned kRowRing = 32;
    constexpr int NB = sizeof(TD) == 4 ? 4 : 2;
    const unsigned nc = (e - b + 31u) / 32u;
    const unsigned left = (e - b) & 31u;
    // Row numbers: ncu's source page showed half of the stall samples waiting for perm[] (read 4 steps ahead into
    // registers).  They now travel through a shared-memory ring, copied kRowsAhead chunks ahead by cp.async (one 4-byte
    // copy per lane and chunk; each lane reads back only what it copied itself).
    __shared__ unsigned s_rows[4][kRowRing][32];
    auto rows_issue = [&](unsigned c) {
        const unsigned i = b + c * 32u + lane;
        unsigned *dst = &s_rows[warp][c % kRowRing][lane];
        if (c < nc && i < e)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst)), "l"(perm + i) : "memory");
        else
            *dst = 0xffffffffu;
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    auto rows_get = [&](unsigned c) -> unsigned { return s_rows[warp][c % kRowRing][lane]; };
    auto pull = [&](unsigned row) {
        if (row != 0xffffffffu) {
            const char *p0 = reinterpret_cast<const char *>(data + (size_t)row * kOrdDim);
            asm volatile("prefetch.global.L2 [%0];" ::"l"(p0));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(p0 + kOrdDim * sizeof(TD) - 1));
        }
    };
    TD buf[NB][32];
    auto fetch = [&](TD (&v)[32], unsigned mine) {
        if (mine == 0xffffffffu) mine = 0u;                               // past the end: row 0, loaded and not added
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const unsigned row = __shfl_sync(0xffffffffu, mine, j);
            v[j] = data[(size_t)row * kOrdDim + d];
        }
    };
    constexpr int U = 4;                    // steps per loop body (a multiple of NB: no buffer is ever copied)
    if (nc > 0) {
        for (unsigned q = 0; q < kRowsAhead; ++q) rows_issue(q);
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        for (unsigned q = NB - 1; q <= kOrdAhead; ++q) pull(rows_get(q));
#pragma unroll
        for (int u = 0; u < NB - 1; ++u) fetch(buf[u], rows_get(u));
    }
    for (unsigned c = 0; c < nc; c += U) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const unsigned cc = c + u;                                    // the chunk added in this step: buf[u % NB]
            if (cc < nc) {
                rows_issue(cc + kRowsAhead);
                asm volatile("cp.async.wait_group %0;" ::"n"(kRowsAhead - kOrdAhead - 1) : "memory");   // chunk cc + kOrdAhead + 1 is here
                if (cc + NB - 1 < nc) fetch(buf[(u + NB - 1) % NB], rows_get(cc + NB - 1));
                pull(rows_get(cc + kOrdAhead + 1));
                if (cc + 1 < nc || left == 0u) {
                    ord_add32(acc, buf[u % NB]);                          // ascending row order, float64: cb_func.py:86
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) if ((unsigned)j < left) acc += (double)buf[u % NB][j];
                }
            }
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (lane < kOrdDim) sums[(size_t)k * kOrdDim + lane] = acc;
    if (lane == 0) counts[k] += (double)(e - b);
}

// The same sums as a sweep over L2-sized blocks of rows (one launch per block).  A centroid's rows are scattered over
// the whole set, so ord_sum_kernel gathers 68-byte rows straight from DRAM (96-128 bytes fetched per row, no page
// locality).  Here every launch first pulls the NEXT block of rows into L2 with coalesced prefetches (streaming DRAM
// traffic) and then every centroid adds its rows of THIS block, which the previous launch left in L2; the running sum
// and the position in the centroid's row list are carried in global memory (float64, exact) from launch to launch.
constexpr size_t kOrdBlockBytes = (size_t)32 << 20;

// Streams [p, p + bytes) through L2 with real 16-byte loads (a prefetch instruction is a hint that a busy memory
// system drops; these are not).  `first`, `count`: this CTA's position among the CTAs that share the range.
__device__ __forceinline__ void ord_prefetch_range(const char *p, size_t bytes, unsigned first, unsigned count, unsigned *sink)
{
    const uintptr_t a0 = (reinterpret_cast<uintptr_t>(p) + 15) & ~(uintptr_t)15;
    const uintptr_t a1 = (reinterpret_cast<uintptr_t>(p) + bytes) & ~(uintptr_t)15;
    if (bytes < 32 || a1 <= a0) return;
    const uint4 *q = reinterpret_cast<const uint4 *>(a0);
    const size_t n = (a1 - a0) / 16;
    unsigned x = 0u;
    for (size_t l = (size_t)first * blockDim.x + threadIdx.x; l < n; l += (size_t)count * blockDim.x) {
        uint4 v;
        asm volatile("ld.global.cg.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(q + l));
        x ^= v.x ^ v.y ^ v.z ^ v.w;
    }
    if (x == 0x9e3779b9u && bytes == 1) *sink = x;        // never true: keeps the loads
}

__global__ void ord_prefetch_kernel(const char *p, size_t bytes, const unsigned *__restrict__ coff, unsigned *__restrict__ cur, int K)
{
    ord_prefetch_range(p, bytes, blockIdx.x, gridDim.x, cur);
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < K; k += gridDim.x * blockDim.x) cur[k] = coff[k];
}

template <typename TD>
__global__ void __launch_bounds__(128) ord_sum_block_kernel(const TD *__restrict__ data, const unsigned *__restrict__ perm,
                                                            const unsigned *__restrict__ coff, unsigned *__restrict__ cur, int K,
                                                            double *__restrict__ sums, double *__restrict__ counts,
                                                            unsigned row_end, const char *next, size_t next_bytes, int last,
                                                            unsigned sum_ctas)
{
    if (blockIdx.x >= sum_ctas) {           // the CTAs behind the centroids' stream the NEXT block of rows into L2
        ord_prefetch_range(next, next_bytes, blockIdx.x - sum_ctas, gridDim.x - sum_ctas, cur);
        return;
    }
    const int lane = threadIdx.x & 31;
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int k = blockIdx.x * (blockDim.x >> 5) + warp;
    if (k >= K) return;
    const unsigned i0 = __shfl_sync(0xffffffffu, cur[k], 0), e = __shfl_sync(0xffffffffu, coff[k + 1], 0);
    const int d = lane < kOrdDim ? lane : 0;
    double acc = sums[(size_t)k * kOrdDim + d];
    auto rows_at = [&](unsigned c) -> unsigned {          // row numbers of chunk c from the cursor; past the list: no row
        const unsigned i = i0 + c * 32u + lane;
        return i < e ? perm[i] : 0xffffffffu;
    };
    // rows of this block: the list is ascending, so they are a prefix of the chunk
    auto inside = [&](unsigned rows) -> unsigned { return (unsigned)__popc(__ballot_sync(0xffffffffu, rows < row_end)); };
    auto fetch = [&](TD (&v)[32], unsigned mine) {
        if (mine >= row_end) mine = 0u;                   // not of this block: row 0, loaded and not added
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const unsigned row = __shfl_sync(0xffffffffu, mine, j);
            v[j] = data[(size_t)row * kOrdDim + d];
        }
    };
    auto add = [&](const TD (&v)[32], unsigned n) {
        if (n == 32u) {
            ord_add32(acc, v);                                                 // ascending row order, float64: cb_func.py:86
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) if ((unsigned)j < n) acc += (double)v[j];
        }
    };
    // NB register buffers of 32 rows: chunk c + NB - 1 is loaded while chunk c is added; the loop is unrolled over
    // U = 4 steps so that no buffer is ever copied, and the row numbers of a chunk are read U steps before its loads.
    constexpr int NB = sizeof(TD) == 4 ? 4 : 2;
    constexpr int U = 4;
    TD buf[NB][32];
    unsigned cn[NB], mrow[U], total = 0u;
#pragma unroll
    for (int u = 0; u < NB; ++u) cn[u] = 0u;
#pragma unroll
    for (int u = 0; u < NB - 1; ++u) {
        const unsigned r = rows_at(u);
        cn[u] = inside(r);
        if (cn[u]) fetch(buf[u], r);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) mrow[u] = rows_at(NB - 1 + u);
    bool done = cn[0] == 0u;
    for (unsigned c = 0; !done; c += U) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (!done) {
                const unsigned cc = c + u;                                // the chunk added in this step: buf[u % NB]
                const unsigned ahead = cn[u % NB] == 32u ? inside(mrow[u]) : 0u;      // chunk cc + NB - 1
                if (ahead) fetch(buf[(u + NB - 1) % NB], mrow[u]);
                mrow[u] = rows_at(cc + NB - 1 + U);
                add(buf[u % NB], cn[u % NB]);
                total += cn[u % NB];
                done = cn[u % NB] < 32u;
                cn[(u + NB - 1) % NB] = ahead;
                if (!done && cn[(u + 1) % NB] == 0u) done = true;
            }
        }
    }
    const unsigned i1 = i0 + total;
    if (!last && lane < 4 && i1 + lane * 32u < e)          // the row numbers the next launch starts with
        asm volatile("prefetch.global.L2 [%0];" ::"l"(perm + i1 + lane * 32u));
    if (lane < kOrdDim) sums[(size_t)k * kOrdDim + lane] = acc;
    if (lane == 0) {
        cur[k] = i1;
        if (last) counts[k] += (double)(e - coff[k]);
    }
}

template <typename TD>
static int accumulate_ordered(const TD *d_data, long N, const int32_t *d_idx, int K, double *d_sums, double *d_counts,
                              void *d_ws, size_t ws_bytes, cudaStream_t st)
{
    if (N < 0 || K < 1 || K > kOrdMaxK || N > 0x7fffffffL) return FPC_ERR_SHAPE;
    if (!d_sums || !d_counts) return FPC_ERR_ARG;
    if (N == 0) return FPC_OK;
    if (!d_data || !d_idx || !d_ws) return FPC_ERR_ARG;
    const OrdPlan p = ord_plan(N, K);
    if (ws_bytes < p.bytes) return FPC_ERR_WORKSPACE;
    char *ws = static_cast<char *>(d_ws);
    unsigned *cnt = reinterpret_cast<unsigned *>(ws + p.off_cnt);
    unsigned *gtot = reinterpret_cast<unsigned *>(ws + p.off_gtot);
    unsigned *coff = reinterpret_cast<unsigned *>(ws + p.off_coff);
    unsigned *perm = reinterpret_cast<unsigned *>(ws + p.off_perm);
    const size_t smem = (size_t)kOrdWarps * K * sizeof(unsigned);
    long ctas = (p.tiles + kOrdWarps - 1) / kOrdWarps;
    if (ctas > 148 * 6) ctas = 148 * 6;
    ord_count_kernel<<<(unsigned)ctas, kOrdWarps * 32, smem, st>>>(d_idx, N, K, p.tiles, cnt);
    FPC_LAUNCH_CHECK();
    ord_group_kernel<<<dim3((K + 127) / 128, kOrdGroups), 128, 0, st>>>(cnt, p.tiles, p.tiles_per_group, K, gtot);
    FPC_LAUNCH_CHECK();
    ord_offsets_kernel<<<1, 1024, 0, st>>>(gtot, K, coff);
    FPC_LAUNCH_CHECK();
    ord_scatter_kernel<<<(unsigned)ctas, kOrdWarps * 32, smem, st>>>(d_idx, N, K, p.tiles, p.tiles_per_group, cnt, gtot, coff, perm);
    FPC_LAUNCH_CHECK();
    // FPC_KMEANS_ORDERED_VARIANT=g: the one-pass gather (ord_sum_kernel) instead of the blocked sweep, for A/B.
    // (Also measured and dropped: one cp.async.bulk per row into a shared-memory ring -- 10.7 ms, the copy engine's
    // fixed cost per 80-byte copy; 4-byte cp.async per lane -- 7.5 ms.)
    static const char *variant = getenv("FPC_KMEANS_ORDERED_VARIANT");
    if (variant != nullptr && variant[0] == 'g') {
        ord_sum_kernel<TD><<<(K + 3) / 4, 128, 0, st>>>(d_data, perm, coff, K, d_sums, d_counts);
        FPC_LAUNCH_CHECK();
        return FPC_OK;
    }
    unsigned *cur = reinterpret_cast<unsigned *>(ws + p.off_cur);
    const long block_rows = (long)(kOrdBlockBytes / (kOrdDim * sizeof(TD)));
    const long blocks = (N + block_rows - 1) / block_rows;
    const char *base = reinterpret_cast<const char *>(d_data);
    const size_t row_bytes = kOrdDim * sizeof(TD);
    auto span = [&](long blk) -> size_t {                 // bytes of block blk (0 past the end)
        if (blk >= blocks) return 0;
        const long r0 = blk * block_rows, r1 = (blk + 1) * block_rows < N ? (blk + 1) * block_rows : N;
        return (size_t)(r1 - r0) * row_bytes;
    };
    ord_prefetch_kernel<<<148 * 2, 256, 0, st>>>(base, span(0), coff, cur, K);
    FPC_LAUNCH_CHECK();
    for (long blk = 0; blk < blocks; ++blk) {
        const long r1 = (blk + 1) * block_rows < N ? (blk + 1) * block_rows : N;
        const unsigned sum_ctas = (unsigned)(K + 3) / 4;
        const unsigned pf_ctas = span(blk + 1) ? 148u * 2u : 0u;
        ord_sum_block_kernel<TD><<<sum_ctas + pf_ctas, 128, 0, st>>>(d_data, perm, coff, cur, K, d_sums, d_counts, (unsigned)r1,
                                                                     base + (size_t)(blk + 1) * block_rows * row_bytes, span(blk + 1),
                                                                     blk + 1 == blocks, sum_ctas);
        FPC_LAUNCH_CHECK();
    }
    return FPC_OK;
}

}  // namespace fpc

extern "C" {

size_t fpc_kmeans_ordered_workspace_bytes(long N, int K)
{
    if (N < 0 || K < 1 || K > fpc::kOrdMaxK) return 0;
    return fpc::ord_plan(N, K).bytes;
}

int fpc_kmeans_accumulate_ordered(const void *d_data, int data_is_f64, long N, const int32_t *d_idx, int K,
                                          double *d_sums, double *d_counts, void *d_workspace, size_t workspace_bytes,
                                          void *stream)
{
    if (data_is_f64)
        return fpc::accumulate_ordered<double>(static_cast<const double *>(d_data), N, d_idx, K, d_sums, d_counts, d_workspace,
                                               workspace_bytes, (cudaStream_t)stream);
    return fpc::accumulate_ordered<float>(static_cast<const float *>(d_data), N, d_idx, K, d_sums, d_counts, d_workspace,
                                          workspace_bytes, (cudaStream_t)stream);
}

}
