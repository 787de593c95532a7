// fpc_vq_screen.cuh -- m-best VQ search with fp32 screening and an exact fallback.
//
// Same result, bit for bit, as vq_search_rows (fpc_vq_search.cuh), i.e. as the reference's
// quantize_mstage (/root/reference/src/quantization/vq_func.py:82-131) -- but the 6 x 1024 direct-form
// distances per vector (50 dependent fp operations each) are replaced by a screen that is ~4x cheaper,
// and only rows whose decision is not PROVABLY the same are handed to the exact search.
//
// Screen (all fp32, codebook shadow (float)c, norms and Gram table precomputed by fpc_pack_codebooks):
//   stage 0:   a_k     = ||c0_k||^2 - 2 <x, c0_k>                       (17 FMA;  d_k = a_k + ||x||^2)
//   stage 1:   b_{s,j} = a_{k_s} + (||c1_j||^2 - 2 <x, c1_j>) + 2 <c0_{k_s}, c1_j>
//                      = || (x - c0_{k_s}) - c1_j ||^2 - ||x||^2          (one dot product per j, shared by
//                                                                        the 5 survivors; Gram row add)
// Error budget.  With u = 2^-24, R = (||x|| + Cmax0 + Cmax1)^2 (every distance is <= R):
//   |fp32 value - real value| <= 19 u R per dot-product form (shadow rounding u, norm rounding u, 17 FMA),
//   Gram rounding + adds <= 4 u R, rounding of diff = x - c0 in the reference <= 2 u R, and numpy's own
//   evaluation of the direct form is within 20 u R of the real value.  Everything together < 70 u R.
//   A decision between two screened values is safe when they differ by more than twice that (140 u R).
//   The screen uses the margin M = 2^-15 R = 512 u R (plus a relative 2^-11 where a value was truncated
//   to make room for an index; the truncation is <= 2^-13 relative per value).
// Decisions:
//   stage 0 -- the survivor SET is exact if the 6th smallest screened value exceeds the 5th by more than
//              the margin (the order inside the set only breaks ties, which the margin rules out);
//   stage 1 -- the winner (s, j) is exact if the second smallest screened value exceeds the smallest by
//              more than the margin.
// Otherwise the row is flagged and re-done by vq_search_rows (exact arithmetic, reference tie rules).
// On the synthetic codebooks a few percent of the rows are flagged.
#pragma once
#include "fpc_vq_search.cuh"

namespace fpc {

struct ScreenPart { float v1; int i1; float v2; int pad; };

// scratch: keybuf u32[vb][1024] | surv int[maxn][5] | sval float[maxn][5] | marg float[maxn][2] |
//          part ScreenPart[maxn][16] | flag int[maxn] | flist int[maxn] | xm2 float[maxn][20] (-2 x, 16-byte rows) |
//          goff int[maxn][5] (Gram row offsets) | cnt int[4]
__host__ __device__ constexpr size_t screen_fixed_bytes(int maxn) { return (size_t)maxn * 416 + 16; }

// Register blocking of the dot products: a thread holds TWO codewords (34 + 2 registers) and streams
// TWO vectors from shared memory against them -- four independent FMA chains, half the registers of a
// 4-codeword tile; the codebook is covered in two sub-passes (entries 2t, 2t+1 and 512+2t, 512+2t+1).
struct Cw2 { float c0[kDim], c1[kDim], n0, n1; };

__device__ __forceinline__ void load_cw2(Cw2 &w, const float *__restrict__ cf, const float *__restrict__ nf, int Kp, int k)
{
#pragma unroll
    for (int d = 0; d < kDim; ++d) {
        const float2 t = *reinterpret_cast<const float2 *>(cf + (size_t)d * Kp + k);
        w.c0[d] = t.x; w.c1[d] = t.y;
    }
    const float2 t = *reinterpret_cast<const float2 *>(nf + k);
    w.n0 = t.x; w.n1 = t.y;
}

// a[cw][vec] = ||c_cw||^2 + <xm_vec, c_cw>, with xm = -2 x prepared once per vector in shared memory
__device__ __forceinline__ void dot2x2(const Cw2 &w, const float *__restrict__ xa, const float *__restrict__ xb, float &a00,
                                       float &a10, float &a01, float &a11)
{
    a00 = w.n0; a10 = w.n1; a01 = w.n0; a11 = w.n1;
#pragma unroll
    for (int d4 = 0; d4 < 4; ++d4) {
        const float4 pa = *reinterpret_cast<const float4 *>(xa + 4 * d4);
        const float4 pb = *reinterpret_cast<const float4 *>(xb + 4 * d4);
        const float ea[4] = {pa.x, pa.y, pa.z, pa.w};
        const float eb[4] = {pb.x, pb.y, pb.z, pb.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int d = 4 * d4 + e;
            a00 = __fmaf_rn(ea[e], w.c0[d], a00); a10 = __fmaf_rn(ea[e], w.c1[d], a10);
            a01 = __fmaf_rn(eb[e], w.c0[d], a01); a11 = __fmaf_rn(eb[e], w.c1[d], a11);
        }
    }
    const float ta = xa[16], tb = xb[16];
    a00 = __fmaf_rn(ta, w.c0[16], a00); a10 = __fmaf_rn(ta, w.c1[16], a10);
    a01 = __fmaf_rn(tb, w.c0[16], a01); a11 = __fmaf_rn(tb, w.c1[16], a11);
}

// Gram rows of one vector pair (5 survivors each, the thread's two columns)
struct GramBuf { float2 a[kSurv], c[kSurv]; };

__device__ __forceinline__ void gram_load(GramBuf &g, const float *__restrict__ Gk /* G + this thread's column */,
                                          const int *__restrict__ goff, int va, int vc)
{
#pragma unroll
    for (int s2 = 0; s2 < kSurv; ++s2) {
        g.a[s2] = *reinterpret_cast<const float2 *>(Gk + goff[va * kSurv + s2]);
        g.c[s2] = *reinterpret_cast<const float2 *>(Gk + goff[vc * kSurv + s2]);
    }
}

__device__ __forceinline__ unsigned screen_key(float v, unsigned k)
{
    return (__float_as_uint(fmaxf(v, 0.0f)) & 0xfffffc00u) | k;
}

// warp: smallest value + its index, and the smallest of everything else; lane 0 stores
__device__ __forceinline__ void warp_top2_store(float m1, int i1, float m2, int lane, ScreenPart *dst)
{
    m1 = fmaxf(m1, 0.0f); m2 = fmaxf(m2, 0.0f);          // non-negative: bit patterns order like uints
    const unsigned b1 = __float_as_uint(m1);
    const unsigned w1 = __reduce_min_sync(0xffffffffu, b1);
    const unsigned who = __ballot_sync(0xffffffffu, b1 == w1);
    const int winner = __ffs(who) - 1;
    const int wi = __shfl_sync(0xffffffffu, i1, winner);
    const unsigned b2 = __float_as_uint(lane == winner ? m2 : m1);
    const unsigned w2 = __reduce_min_sync(0xffffffffu, b2);
    if (lane == 0) {
        ScreenPart p;
        p.v1 = __uint_as_float(w1); p.i1 = wi; p.v2 = __uint_as_float(w2); p.pad = 0;
        *dst = p;
    }
}

__device__ __forceinline__ void ins3(unsigned &t0, unsigned &t1, unsigned &t2, unsigned x)
{
    unsigned lo = min(t0, x); x = max(t0, x); t0 = lo;
    lo = min(t1, x); x = max(t1, x); t1 = lo;
    t2 = min(t2, x);
}

__device__ __forceinline__ void ins4(unsigned &t0, unsigned &t1, unsigned &t2, unsigned &t3, unsigned x)
{
    unsigned lo = min(t0, x); x = max(t0, x); t0 = lo;
    lo = min(t1, x); x = max(t1, x); t1 = lo;
    lo = min(t2, x); x = max(t2, x); t2 = lo;
    t3 = min(t3, x);
}

// top-2 of a stream of non-negative floats with the index of the smallest
__device__ __forceinline__ void upd2(float &m1, int &i1, float &m2, float v, int idx)
{
    const bool p = v < m1;
    m2 = fminf(m2, fmaxf(m1, v));
    m1 = fminf(m1, v);
    i1 = p ? idx : i1;
}

// Every one of the NT (256 or 128) compute threads must call this with identical arguments (barrier id 1).
// scratch_bytes >= screen_fixed_bytes(maxn) + 8192 and large enough for the exact fallback
// (vq_fixed_bytes<T>(8) + 1024 * sizeof(T)).  A thread covers the codebook in NSUB = 512 / NT sub-passes of two
// codewords; NSUB * (NT / 32) = 16 partial results per vector either way.
template <typename T, int NT = kComputeThreads>
__device__ __forceinline__ void vq_search_rows_screened(const PackedVq &bk_in, const char *__restrict__ cbbase, const int *__restrict__ list,
                                        int n, int maxn, const float *__restrict__ rs, float *__restrict__ rq,
                                        int *__restrict__ idx1, int *__restrict__ idx2, char *__restrict__ scratch,
                                        int scratch_bytes, int tid, T *__restrict__ qglobal = nullptr, long long *dbg = nullptr)
{
    constexpr int NW = NT / 32, NSUB = 512 / NT;
    static_assert(NT == 256 || NT == 128, "block-wide search is written for 256 or 128 threads");
    const int warp = tid >> 5, lane = tid & 31;
    // The header is re-read here on every call (volatile): its fields are loop-invariant for the
    // frame loop of the fused kernels, and left to itself the compiler hoists a dozen 64-bit offsets
    // out of that loop and keeps them live through the gate GEMM, which then spills.
    PackedVq bk;
    {
        const volatile long long *src = reinterpret_cast<const volatile long long *>(&bk_in);
        long long *dst = reinterpret_cast<long long *>(&bk);
#pragma unroll
        for (int i = 0; i < (int)(sizeof(PackedVq) / 8); ++i) dst[i] = src[i];
    }
    const int Kp = bk.Kp;
    const bool two = bk.stages == 2;
    long long tq0 = dbg ? clock64() : 0;
#define FPC_VQT(i) do { if (dbg) { const long long t_ = clock64(); dbg[i] += t_ - tq0; tq0 = t_; } } while (0)
    int vb = (scratch_bytes - (int)screen_fixed_bytes(maxn)) / 4096;
    vb = (vb > 8 ? 8 : vb) & ~1;                       // vectors are processed in pairs
    unsigned *keybuf = reinterpret_cast<unsigned *>(scratch);
    int *surv = reinterpret_cast<int *>(scratch + (size_t)vb * 4096);
    float *sval = reinterpret_cast<float *>(surv + maxn * 5);
    float *marg = sval + maxn * 5;
    ScreenPart *part = reinterpret_cast<ScreenPart *>(marg + maxn * 2);
    int *flag = reinterpret_cast<int *>(part + maxn * 16);
    int *flist = flag + maxn;
    float *xm2 = reinterpret_cast<float *>(flist + maxn);     // 16-byte aligned: every block above is a multiple of 16 per row
    int *goff = reinterpret_cast<int *>(xm2 + maxn * 20);
    int *cnt = goff + maxn * 5;
    const float *cmax = reinterpret_cast<const float *>(cbbase + bk.off_cmax);

    // ---- per-vector constants: margin M and the offset nx that makes screened values positive ----
    if (tid < n) {
        const float *xr = rs + list[tid] * kLdR + 4;
        float n2 = 0.0f;
#pragma unroll
        for (int d = 0; d < kDim; ++d) n2 = __fmaf_ru(xr[d], xr[d], n2);
        const float csum = __fadd_ru(cmax[0], two ? cmax[1] : 0.0f);
        const float r = __fadd_ru(__fsqrt_ru(n2), csum);
        const float M = __fmul_ru(__fmul_ru(r, r), 3.0517578125e-5f);   // 2^-15 R
        marg[2 * tid] = M;
        marg[2 * tid + 1] = __fadd_ru(n2, M);
        flag[tid] = 0;
#pragma unroll
        for (int d = 0; d < kDim; ++d) xm2[tid * 20 + d] = -2.0f * xr[d];      // exact (power of two)
    }
    named_bar_sync(1, NT);
    FPC_VQT(2);

    Cw2 w;
    if (two) {
        // ============ stage 0: survivor set ============
        const float *cf = reinterpret_cast<const float *>(cbbase + bk.off_f[0]);
        const float *nf = reinterpret_cast<const float *>(cbbase + bk.off_n[0]);
        for (int base = 0; base < n; base += vb) {
            const int nb = min(vb, n - base);
#pragma unroll 1
            for (int h = 0; h < NSUB; ++h) {
                const int k = 2 * tid + 2 * NT * h;
                const bool act = k < Kp;
                if (act) load_cw2(w, cf, nf, Kp, k);
                for (int v = 0; v < nb; v += 2) {
                    const int va = base + v, vc = min(base + v + 1, n - 1);     // odd tail: the last vector twice
                    uint2 ka = make_uint2(0xffffffffu, 0xffffffffu), kc = ka;
                    if (act) {
                        float a00, a10, a01, a11;
                        dot2x2(w, xm2 + va * 20, xm2 + vc * 20, a00, a10, a01, a11);
                        const float nxa = marg[2 * va + 1], nxc = marg[2 * vc + 1];
                        ka.x = screen_key(a00 + nxa, (unsigned)k); ka.y = screen_key(a10 + nxa, (unsigned)k + 1);
                        kc.x = screen_key(a01 + nxc, (unsigned)k); kc.y = screen_key(a11 + nxc, (unsigned)k + 1);
                    }
                    *reinterpret_cast<uint2 *>(keybuf + v * 1024 + k) = ka;
                    if (v + 1 < nb) *reinterpret_cast<uint2 *>(keybuf + (v + 1) * 1024 + k) = kc;
                }
            }
            named_bar_sync(1, NT);
            FPC_VQT(3);
#pragma unroll 1
            for (int vw = warp; vw < nb; vw += NW) {
                const int v = base + vw;
                unsigned t0 = 0xffffffffu, t1 = 0xffffffffu, t2 = 0xffffffffu, t3 = 0xffffffffu;
#pragma unroll 2
                for (int i = 0; i < 8; ++i) {
                    const uint4 kk = reinterpret_cast<const uint4 *>(keybuf + vw * 1024)[i * 32 + lane];
                    ins4(t0, t1, t2, t3, kk.x); ins4(t0, t1, t2, t3, kk.y); ins4(t0, t1, t2, t3, kk.z); ins4(t0, t1, t2, t3, kk.w);
                }
                unsigned g[8];
                int popped = 0;
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const unsigned gm = __reduce_min_sync(0xffffffffu, t0);
                    g[r] = gm;
                    if (t0 == gm) { t0 = t1; t1 = t2; t2 = t3; t3 = 0xffffffffu; ++popped; }
                }
                // a lane that gave up all four of its keys may hide a fifth: undecidable here
                const bool exhausted = __any_sync(0xffffffffu, popped >= 4);
                const float M = marg[2 * v];
                const float v5 = __uint_as_float(g[4] & 0xfffffc00u);
                const float thr = __fadd_ru(__fmaf_ru(v5, 4.8828125e-4f, v5), M);
                // keys above thr are provably worse than the five smallest screened keys, so the exact top 5 is
                // among the nc keys at or below it
                int nc = 5;
#pragma unroll
                for (int r = 5; r < 8; ++r) nc += (__uint_as_float(g[r] & 0xfffffc00u) <= thr) ? 1 : 0;   // NaN-safe below
                bool ok = !exhausted && (__uint_as_float(g[5] & 0xfffffc00u) > thr);
                if (!ok && !exhausted && nc < 8) {      // g is sorted: the first key not counted in nc is above thr
                    // near-tie around rank 5: decide among the nc candidates with the reference's exact arithmetic
                    unsigned mine = g[0];
#pragma unroll
                    for (int r = 1; r < 8; ++r) mine = lane == r ? g[r] : mine;
                    T d = Rn<T>::inf();
                    int ki = 0x7fffffff;
                    if (lane < nc) {
                        ki = (int)(mine & 1023u);
                        const T *crow = reinterpret_cast<const T *>(cbbase + bk.off_r[0]) + (size_t)ki * kDim;
                        const float *xr = rs + list[v] * kLdR + 4;
                        T xv[kDim], cv[kDim];
#pragma unroll
                        for (int dd = 0; dd < kDim; ++dd) { xv[dd] = (T)xr[dd]; cv[dd] = crow[dd]; }
                        d = dist17<T>(xv, cv);
                    }
#pragma unroll
                    for (int r = 0; r < kSurv; ++r) {
                        T wd = d;
                        int wi = ki;
                        warp_argmin(wd, wi);
                        g[r] = (unsigned)wi;                  // only the index part is used from here on
                        if (ki == wi) { d = Rn<T>::inf(); ki = 0x7fffffff; }
                    }
                    ok = true;
                }
                if (lane < kSurv) {
                    const unsigned gl = lane == 0 ? g[0] : lane == 1 ? g[1] : lane == 2 ? g[2] : lane == 3 ? g[3] : g[4];
                    const int ks = ok ? (int)(gl & 1023u) : lane;          // flagged rows keep harmless indices
                    surv[v * kSurv + lane] = ks;
                    goff[v * kSurv + lane] = ks * Kp;
                    // full-precision screened value of the survivor (||x - c0||^2 up to the common offset);
                    // (float)c of the row-major copy is the shadow value, and the row is contiguous
                    float a = nf[ks];
                    const float *xr = rs + list[v] * kLdR + 4;
                    const T *crow = reinterpret_cast<const T *>(cbbase + bk.off_r[0]) + (size_t)ks * kDim;
#pragma unroll
                    for (int d = 0; d < kDim; ++d) a = __fmaf_rn(-2.0f * xr[d], (float)crow[d], a);
                    // stored with the margin already added: the last stage uses it as is (one add, done once here)
                    sval[v * kSurv + lane] = fmaxf(a + marg[2 * v + 1], 0.0f) + M;
                }
                if (lane == 0 && !ok) flag[v] = 1;
            }
            named_bar_sync(1, NT);
            FPC_VQT(4);
        }
    }

    // ============ last stage: nearest entry (jointly over the survivors) ============
    {
        const int sc = two ? 1 : 0;
        const float *cf = reinterpret_cast<const float *>(cbbase + bk.off_f[sc]);
        const float *nf = reinterpret_cast<const float *>(cbbase + bk.off_n[sc]);
        const float *G = reinterpret_cast<const float *>(cbbase + bk.off_g);
        const int ns = two ? kSurv : 1;
#pragma unroll 1
        for (int h = 0; h < NSUB; ++h) {
            const int k = 2 * tid + 2 * NT * h;
            const bool act = k < Kp;
            if (act) load_cw2(w, cf, nf, Kp, k);
            // Gram rows (2 vectors x 5 survivors, one float2 each) are fetched a whole vector pair ahead into
            // the other of two register buffers (the loop handles two pairs per trip, so no copies are needed):
            // an L2 round trip is several hundred cycles, the work on one pair about as long.
            GramBuf gx, gy;
#pragma unroll
            for (int s2 = 0; s2 < kSurv; ++s2) {
                gx.a[s2] = make_float2(0.0f, 0.0f); gx.c[s2] = gx.a[s2]; gy.a[s2] = gx.a[s2]; gy.c[s2] = gx.a[s2];
            }
            const bool ld = act && two;
            const float *Gk = G + k;
            if (ld) gram_load(gx, Gk, goff, 0, min(1, n - 1));
            auto pair = [&](const GramBuf &g, int v) {
                const int va = v, vc = min(v + 1, n - 1);
                const float inf = __int_as_float(0x7f800000);
                float ma1 = inf, ma2 = inf, mc1 = inf, mc2 = inf;
                int ia = 0, ic = 0;
                if (act) {
                    float a00, a10, a01, a11;
                    dot2x2(w, xm2 + va * 20, xm2 + vc * 20, a00, a10, a01, a11);
                    const float Ma = marg[2 * va], nxa = marg[2 * va + 1], Mc = marg[2 * vc], nxc = marg[2 * vc + 1];
#pragma unroll
                    for (int s2 = 0; s2 < kSurv; ++s2) {
                        if (s2 < ns) {
                            const float bsa = two ? sval[va * kSurv + s2] : nxa;
                            const float bsc = two ? sval[vc * kSurv + s2] : nxc;
                            const int ib = (s2 << 10) | k;
                            upd2(ma1, ia, ma2, (a00 + bsa) + g.a[s2].x, ib);
                            upd2(ma1, ia, ma2, (a10 + bsa) + g.a[s2].y, ib + 1);
                            upd2(mc1, ic, mc2, (a01 + bsc) + g.c[s2].x, ib);
                            upd2(mc1, ic, mc2, (a11 + bsc) + g.c[s2].y, ib + 1);
                        }
                    }
                }
                warp_top2_store(ma1, ia, ma2, lane, &part[va * 16 + h * NW + warp]);
                if (v + 1 < n) warp_top2_store(mc1, ic, mc2, lane, &part[vc * 16 + h * NW + warp]);
            };
            for (int v = 0; v < n; v += 4) {
                if (ld && v + 2 < n) gram_load(gy, Gk, goff, v + 2, min(v + 3, n - 1));
                pair(gx, v);
                if (v + 2 < n) {
                    if (ld && v + 4 < n) gram_load(gx, Gk, goff, v + 4, min(v + 5, n - 1));
                    pair(gy, v + 2);
                }
            }
        }
        named_bar_sync(1, NT);
        FPC_VQT(5);
        if (tid < n && !flag[tid]) {
            const int v = tid;
            float b = part[v * 16].v1, sec = part[v * 16].v2;
            int bi = part[v * 16].i1;
#pragma unroll
            for (int w2 = 1; w2 < 16; ++w2) {
                const ScreenPart p = part[v * 16 + w2];
                if (p.v1 < b) { sec = fminf(fminf(sec, b), p.v2); b = p.v1; bi = p.i1; }
                else { sec = fminf(sec, p.v1); }
            }
            const float thr = __fadd_ru(__fmaf_ru(b, 9.5367431640625e-7f, b), marg[2 * v]);
            if (sec > thr) {
                const int row = list[v];
                if (two) { idx1[row] = surv[v * kSurv + (bi >> 10)]; idx2[row] = bi & 1023; }
                else { idx1[row] = bi & 1023; idx2[row] = -1; }
            } else {
                flag[v] = 1;
            }
        }
        named_bar_sync(1, NT);
    }

    // ---- quantised vectors of the decided rows: csum = 0; csum += CB[i][index[i,0]]  (vq_func.py:127-129) ----
    {
        const T *cbr0 = reinterpret_cast<const T *>(cbbase + bk.off_r[0]);
        const T *cbr1 = reinterpret_cast<const T *>(cbbase + bk.off_r[1]);
        for (int e = tid; e < n * kDim; e += NT) {
            const int v = e / kDim, d = e - v * kDim;
            if (!flag[v]) {
                const int row = list[v];
                T csum = Rn<T>::add((T)0, cbr0[(size_t)idx1[row] * kDim + d]);
                if (two) csum = Rn<T>::add(csum, cbr1[(size_t)idx2[row] * kDim + d]);
                rq[row * 20 + 1 + d] = (float)csum;
                if (qglobal) qglobal[(size_t)row * kDim + d] = csum;
            }
        }
        if (warp == 0) {
            int nf2 = 0;
            for (int base = 0; base < n; base += 32) {
                const int v = base + lane;
                const bool f = v < n && flag[v] != 0;
                const unsigned bf = __ballot_sync(0xffffffffu, f);
                if (f) flist[nf2 + __popc(bf & ((1u << lane) - 1u))] = list[v];
                nf2 += __popc(bf);
            }
            if (lane == 0) cnt[0] = nf2;
        }
        named_bar_sync(1, NT);
    }

    // ---- exact search for the undecided rows ----
    const int nflag = cnt[0];
    FPC_VQT(6);
    if (dbg) { dbg[0] += n; dbg[1] += nflag; }     // debug counters (rows searched, rows sent to the exact search)
    if (nflag > 0) {
        // the row list must survive the fallback, which reuses the scratch: keep it in registers
        int mine = tid < nflag ? flist[tid] : 0;
        named_bar_sync(1, NT);
        int sb = 32;
        while (sb > 8 && (int)vq_fixed_bytes<T>(sb) + 1024 * (int)sizeof(T) + 4 * maxn > scratch_bytes) sb >>= 1;
        int *keep = reinterpret_cast<int *>(scratch + scratch_bytes - 4 * maxn);   // tail of the scratch: the list
        if (tid < nflag) keep[tid] = mine;
        named_bar_sync(1, NT);
        int vbe = (scratch_bytes - 4 * maxn - (int)vq_fixed_bytes<T>(sb)) / (1024 * (int)sizeof(T));
        vbe = vbe > 8 ? 8 : vbe;
        for (int off = 0; off < nflag; off += sb)
            vq_search_rows<T, NT>(bk, cbbase, keep + off, min(sb, nflag - off), sb, rs, rq, idx1, idx2, scratch, vbe, tid, qglobal);
        FPC_VQT(7);
    }
#undef FPC_VQT
}

// dtype dispatch used by the fused frame-step kernels
template <int NT = kComputeThreads>
__device__ __forceinline__ void vq_dispatch_screened(const PackedVq &bk, const char *cbbase, const int *list, int n, int maxn,
                                                     const float *rs, float *rq, int *idx1, int *idx2, char *scratch,
                                                     int scratch_bytes, int tid, long long *dbg = nullptr)
{
    if (bk.dtype == FPC_F32)
        vq_search_rows_screened<float, NT>(bk, cbbase, list, n, maxn, rs, rq, idx1, idx2, scratch, scratch_bytes, tid, nullptr, dbg);
    else
        vq_search_rows_screened<double, NT>(bk, cbbase, list, n, maxn, rs, rq, idx1, idx2, scratch, scratch_bytes, tid, nullptr, dbg);
}

}  // namespace fpc
