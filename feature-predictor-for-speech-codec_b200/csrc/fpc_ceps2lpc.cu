// fpc_ceps2lpc.cu -- cepstrum -> LPC, the step both callers run right after Wavernn.encoder
// (/root/reference/src/synthesis_qtz.py:158-160, generate_qtz_features.py:61-64).
//
// Reference: /root/reference/src/ceps2lpc/ceps2lpc_vct.py -- ceps2lpc_v (:122-162): c0 += 4, idct (:35-43),
// 10^x * COMPENSATION (:133), interp_band_gain to 161 bins (:45-57), autocorrelation = irfft(320)[:17] (:137-140),
// -40 dB noise floor and lag window (:144-148), then _celt_lpc_s (:60-88) per frame in a Python loop.
// Frames are independent: one thread per frame, a grid-stride loop over frames (the tables below are built once per
// CTA).  The interpolated spectrum is LINEAR in the 18 band energies E[b], so the 17 autocorrelation lags
//     acr[n] = (X0 + 2 sum_{k=1..159} X_k cos(2 pi k n / 320)) / 320,   X_k = (1 - f_k) E[b_k] + f_k E[b_k + 1]
// (bin 160 is never written by the interpolation, :49-56) collapse to an 18 x 17 float64 table T[b][n] built from
// the reference's float32 interpolation weights:  acr[n] = sum_b E[b] T[b][n], 306 float64 FMAs per frame instead of
// 2 703.  That skips the reference's float32 rounding of every X_k (and torch's float32 FFT): against a float64
// evaluation of the whole algorithm (the test suite's ceps2lpc_f64 ground truth) the result is CLOSER than the
// reference's own (tests/test_ceps2lpc_bitrate.py).  Everything else is float32 in the reference's operation order,
// including the two early exits of the Levinson recursion.  The recursion is ill-conditioned (1e-4 noise floor), so
// LPC parity is a stated tolerance, not bit-exact.
#include "fpc_common.cuh"

namespace fpc {

constexpr int kBands = 18, kLpc = 16, kWin = 320;

__global__ void __launch_bounds__(128) ceps2lpc_kernel(const float *__restrict__ ceps, long n, int stride,
                                                      float *__restrict__ lpc_out, float *__restrict__ err_out,
                                                      float *__restrict__ rc_out)
{
    __shared__ float s_dct[kBands * kBands];
    __shared__ float s_comp[kBands];
    __shared__ double s_cos[kWin];
    __shared__ int s_band[kWin / 2];        // band index of bin k
    __shared__ float s_frac[kWin / 2];      // interpolation weight of bin k
    __shared__ __align__(16) double s_T[kBands][kLpc + 2];   // lag table, row = band (padded to 18 for 16-byte loads)
    __shared__ float s_out[2][128][kLpc + 1];                // lpc / rc of the CTA's 128 frames, for coalesced stores
    const float comp[kBands] = {0.8f, 1.0f, 1.0f, 1.0f, 1.0f, 1.0f, 1.0f, 1.0f, 0.666667f, 0.5f, 0.5f, 0.5f,
                                0.333333f, 0.25f, 0.25f, 0.2f, 0.166667f, 0.173913f};
    const int eband[kBands] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 10, 12, 14, 16, 20, 24, 28, 34, 40};
    for (int t = threadIdx.x; t < kBands * kBands; t += blockDim.x) {
        const int i = t / kBands, j = t - i * kBands;
        float v = cospif((float)((i + 0.5) * j / kBands));
        if (j == 0) v = __fmul_rn(v, sqrtf(0.5f));
        s_dct[t] = v;
    }
    for (int t = threadIdx.x; t < kWin; t += blockDim.x) s_cos[t] = cospi(2.0 * t / kWin);
    if (threadIdx.x < kBands) s_comp[threadIdx.x] = comp[threadIdx.x];
    for (int t = threadIdx.x; t < kWin / 2; t += blockDim.x) {
        int b = 0;
        while (b < kBands - 2 && t >= eband[b + 1] * 4) ++b;
        const int bs = (eband[b + 1] - eband[b]) * 4;
        s_band[t] = b;
        s_frac[t] = (float)((double)(t - eband[b] * 4) / bs);
    }
    __syncthreads();

    for (int t = threadIdx.x; t < kBands * (kLpc + 2); t += blockDim.x) {
        const int b = t / (kLpc + 2), m = t - b * (kLpc + 2);
        double sum = 0.0;
        if (m <= kLpc) {
            const int k0 = eband[b > 0 ? b - 1 : 0] * 4, k1 = eband[b < kBands - 1 ? b + 1 : kBands - 1] * 4;
            for (int k = (k0 > 1 ? k0 : 1); k < k1; ++k) {
                const float fr = s_frac[k];
                // the reference's float32 weights of bin k: (1 - frac) on E[b_k], frac on E[b_k + 1]  (:56)
                const double w = s_band[k] == b ? (double)(float)(1.0 - (double)fr) : s_band[k] == b - 1 ? (double)fr : 0.0;
                sum = fma(w, s_cos[(k * m) % kWin], sum);
            }
            sum = ((b == 0 ? 1.0 : 0.0) + 2.0 * sum) / kWin;      // bin 0: frac = 0 -> E[0]
        }
        s_T[b][m] = sum;
    }
    __syncthreads();

    for (long base = (long)blockIdx.x * 128; base < n; base += (long)gridDim.x * 128) {
    const long f = base + threadIdx.x;
    const bool live = f < n;
    float c[kBands];
#pragma unroll
    for (int j = 0; j < kBands; ++j) c[j] = live ? ceps[f * stride + j] : 0.0f;
    c[0] = __fadd_rn(c[0], 4.0f);
    const float k2 = sqrtf(2.0f / kBands);
    double acc[kLpc + 2];
#pragma unroll
    for (int m = 0; m < kLpc + 2; ++m) acc[m] = 0.0;
    for (int i = 0; i < kBands; ++i) {
        float sm = 0.0f;
#pragma unroll
        for (int j = 0; j < kBands; ++j) sm = __fadd_rn(sm, __fmul_rn(c[j], s_dct[i * kBands + j]));
        const double e = (double)__fmul_rn(exp10f(__fmul_rn(sm, k2)), s_comp[i]);
#pragma unroll
        for (int m = 0; m < kLpc + 2; m += 2) {
            const double2 t = *reinterpret_cast<const double2 *>(&s_T[i][m]);
            acc[m] = fma(e, t.x, acc[m]);
            acc[m + 1] = fma(e, t.y, acc[m + 1]);
        }
    }
    float ac[kLpc + 1];
#pragma unroll
    for (int m = 0; m <= kLpc; ++m) ac[m] = (float)acc[m];
    ac[0] = __fadd_rn(ac[0], __fadd_rn(__fmul_rn(ac[0], 0.0001f), (float)(320.0 / 12 / 38.)));
#pragma unroll
    for (int i = 1; i <= kLpc; ++i) ac[i] = __fmul_rn(ac[i], (float)(1 - 0.00006 * i * i));

    // ---- _celt_lpc_s (ceps2lpc_vct.py:60-88): Levinson-Durbin with the reference's early exits ----
    float lp[kLpc], rc[kLpc];
#pragma unroll
    for (int i = 0; i < kLpc; ++i) { lp[i] = 0.0f; rc[i] = 0.0f; }
    float error = ac[0];
    if (ac[0] != 0.0f) {
        bool done = false;
#pragma unroll
        for (int i = 0; i < kLpc; ++i) {
            if (!done) {
                float rr = 0.0f;
#pragma unroll
                for (int j = 0; j < i; ++j) rr = __fadd_rn(rr, __fmul_rn(lp[j], ac[i - j]));
                rr = __fadd_rn(rr, ac[i + 1]);
                const float r = __fdiv_rn(-rr, error);
                rc[i] = r;
                lp[i] = r;
#pragma unroll
                for (int j = 0; j < (i + 1) / 2; ++j) {
                    const float t1 = lp[j], t2 = lp[i - 1 - j];
                    lp[j] = __fadd_rn(t1, __fmul_rn(r, t2));
                    lp[i - 1 - j] = __fadd_rn(t2, __fmul_rn(r, t1));
                }
                error = __fsub_rn(error, __fmul_rn(__fmul_rn(r, r), error));
                if (error < __fdiv_rn(ac[0], 1024.0f)) done = true;
                if (error < __fmul_rn(0.001f, ac[0])) done = true;
            }
        }
    }
    // every thread wrote 16 + 16 consecutive floats 64 B apart: through shared memory the CTA stores its 128 frames as
    // two contiguous 8 KB runs instead
#pragma unroll
    for (int i = 0; i < kLpc; ++i) { s_out[0][threadIdx.x][i] = lp[i]; s_out[1][threadIdx.x][i] = rc[i]; }
    if (live && err_out) err_out[f] = error;
    __syncthreads();
    const long rows = (n - base) < 128 ? (n - base) : 128;
    for (int t = threadIdx.x; t < rows * kLpc; t += 128) {
        lpc_out[base * kLpc + t] = s_out[0][t >> 4][t & 15];
        if (rc_out) rc_out[base * kLpc + t] = s_out[1][t >> 4][t & 15];
    }
    __syncthreads();
    }
}

}  // namespace fpc

extern "C" int fpc_ceps2lpc(const float *d_ceps, long n, int stride, float *d_lpc, float *d_err, float *d_rc, void *stream)
{
    using namespace fpc;
    if (n < 0 || stride < kBands) return FPC_ERR_ARG;
    if (n == 0) return FPC_OK;
    if (!d_ceps || !d_lpc) return FPC_ERR_ARG;
    long blocks = (n + 127) / 128;
    if (blocks > 148 * 8) blocks = 148 * 8;          // grid-stride beyond one wave: the tables are built once per CTA
    ceps2lpc_kernel<<<(unsigned)blocks, 128, 0, (cudaStream_t)stream>>>(d_ceps, n, stride, d_lpc, d_err, d_rc);
    FPC_LAUNCH_CHECK();
    return FPC_OK;
}
