/*
 * fpc_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the reference's closed-loop predictive-coding hot path and
 * of its k-means codebook update.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library; the product path
 * (feature-predictor-for-speech-codec_b200/) never does.
 *
 * What it follows in the reference (paths relative to /root/reference/src):
 *   models/wavernn.py:63-102     Wavernn.forward  (GRU 20->384, GRU 384->128, relu,
 *                                dual_fc summed twice == 2*tanh(W relu(h2)+b))
 *   models/wavernn.py:165-256    Wavernn.encoder  (closed loop, thresholds, feedback)
 *   quantization/vq_func.py:10-24,82-131,134-164   m-best multi-stage VQ search
 *   quantization/vq_func.py:167-185                scalar quantizer
 *   quantization/cb_func.py:56-68,71-100,103-112   find_nearest / update / quantize
 *
 * Arithmetic contract ("canonical fp32"):
 *   - The quantizer arithmetic is the reference's own, bit for bit: numpy evaluates
 *     np.sum((x - cb) ** 2, -1) as t=x-c (rounded), e=t*t (rounded), then numpy's
 *     pairwise reduction (8 interleaved partial sums for 8 <= n <= 128), in the
 *     codebook's dtype (float32 or float64).  Ties resolve to the lowest index
 *     (stable sorted() / argmin).
 *   - The predictor (torch.nn.GRU / Linear on MKL) has no documented summation order,
 *     so the oracle fixes one: every dot product is a single fused-multiply-add chain
 *     in ascending k, seeded with the bias, hidden part first then input part; sigmoid
 *     and tanh are the polynomial forms below (only IEEE add/mul/fma/div, so the CUDA
 *     kernel can reproduce them exactly).  The oracle-vs-reference difference this
 *     introduces (~1e-7) is measured against golden vectors produced by the real
 *     reference (tests/golden, oracle/gen_golden.py).
 *
 * Build: see oracle/build.py  (gcc -O2 -ffp-contract=off -mfma -fopenmp -shared -fPIC).
 * -ffp-contract=off is REQUIRED: every fma in here is explicit.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_SURVIVORS 5 /* vq_func.py:3 */
#define ORC_MAX_STAGES 2
#define ORC_MAX_DIM 128

/* ------------------------------------------------------------------------------------------
 * canonical transcendental functions (shared definition with csrc/fpc_math.cuh)
 * ---------------------------------------------------------------------------------------- */
static inline float orc_exp_core(float x) /* x already clamped to [-87, 88] */
{
    const float magic = 12582912.0f; /* 1.5 * 2^23: round-to-nearest-even integer extraction */
    float t = fmaf(x, 1.44269504088896341f, magic);
    float n = t - magic;
    float r = fmaf(n, -0.693359375f, x);          /* ln2 hi (Cephes split) */
    r = fmaf(n, 2.12194440e-4f, r);               /* ln2 lo */
    float p = 1.9875691500e-4f;
    p = fmaf(p, r, 1.3981999507e-3f);
    p = fmaf(p, r, 8.3334519073e-3f);
    p = fmaf(p, r, 4.1665795894e-2f);
    p = fmaf(p, r, 1.6666665459e-1f);
    p = fmaf(p, r, 5.0000001201e-1f);
    float y = fmaf(p, r * r, r) + 1.0f;
    /* scale by 2^n through the exponent field; y in [0.70, 1.42], n in [-126, 127] */
    int32_t ni = (int32_t)n;
    union { float f; int32_t i; } u;
    u.f = y;
    u.i += ni * (1 << 23);
    return u.f;
}

float orc_expf(float x)
{
    if (x > 88.0f) x = 88.0f;
    if (x < -87.0f) x = -87.0f;
    return orc_exp_core(x);
}

float orc_sigmoidf(float x)
{
    float a = -x;
    if (a > 88.0f) a = 88.0f;
    if (a < -87.0f) a = -87.0f;
    return 1.0f / (1.0f + orc_exp_core(a));
}

float orc_tanhf(float x)
{
    float ax = fabsf(x);
    float y;
    if (ax < 0.625f) {
        float z = x * x;
        float p = -5.70498872745e-3f;
        p = fmaf(p, z, 2.06390887954e-2f);
        p = fmaf(p, z, -5.37397155531e-2f);
        p = fmaf(p, z, 1.33314422036e-1f);
        p = fmaf(p, z, -3.33332819422e-1f);
        return fmaf(p * z, x, x);
    }
    if (ax > 10.0f) ax = 10.0f;
    float e = orc_exp_core(ax + ax);
    y = 1.0f - 2.0f * (1.0f / (e + 1.0f));
    return x < 0.0f ? -y : y;
}

/* ------------------------------------------------------------------------------------------
 * predictor: torch GRU cell semantics (gate row order r, z, n) -- wavernn.py:71,76
 * ---------------------------------------------------------------------------------------- */
static void orc_gru_cell(int in_dim, int H, const float *w_ih, const float *w_hh,
                         const float *b_ih, const float *b_hh, const float *x, const float *h,
                         float *h_out)
{
    for (int j = 0; j < H; ++j) {
        const float *wir = w_ih + (size_t)j * in_dim;
        const float *wiz = w_ih + (size_t)(H + j) * in_dim;
        const float *win = w_ih + (size_t)(2 * H + j) * in_dim;
        const float *whr = w_hh + (size_t)j * H;
        const float *whz = w_hh + (size_t)(H + j) * H;
        const float *whn = w_hh + (size_t)(2 * H + j) * H;
        float ar = b_ih[j] + b_hh[j];
        float az = b_ih[H + j] + b_hh[H + j];
        float ani = b_ih[2 * H + j];
        float anh = b_hh[2 * H + j];
        /* hidden part first: it does not depend on the frame that is fed back, so the CUDA kernel can run it for
         * frame t+1 while frame t is still in the quantiser */
        for (int k = 0; k < H; ++k) {
            ar = fmaf(whr[k], h[k], ar);
            az = fmaf(whz[k], h[k], az);
            anh = fmaf(whn[k], h[k], anh);
        }
        for (int k = 0; k < in_dim; ++k) {
            ar = fmaf(wir[k], x[k], ar);
            az = fmaf(wiz[k], x[k], az);
            ani = fmaf(win[k], x[k], ani);
        }
        float r = orc_sigmoidf(ar);
        float z = orc_sigmoidf(az);
        float n = orc_tanhf(fmaf(r, anh, ani));
        h_out[j] = fmaf(z, h[j] - n, n); /* (1-z)*n + z*h */
    }
}

typedef struct {
    int in_features, h1, h2, fc;
    const float *w_ih1, *w_hh1, *b_ih1, *b_hh1;
    const float *w_ih2, *w_hh2, *b_ih2, *b_hh2;
    const float *w_fc, *b_fc;
} orc_weights;

/* one frame of Wavernn.forward with T=1 (wavernn.py:63-102); h1/h2 updated in place */
static void orc_predictor_step(const orc_weights *w, const float *x, float *h1, float *h2,
                               float *y, float *tmp /* >= max(h1,h2) floats */)
{
    orc_gru_cell(w->in_features, w->h1, w->w_ih1, w->w_hh1, w->b_ih1, w->b_hh1, x, h1, tmp);
    memcpy(h1, tmp, sizeof(float) * (size_t)w->h1);
    orc_gru_cell(w->h1, w->h2, w->w_ih2, w->w_hh2, w->b_ih2, w->b_hh2, h1, h2, tmp);
    memcpy(h2, tmp, sizeof(float) * (size_t)w->h2);
    for (int j = 0; j < w->fc; ++j) {
        float a = w->b_fc[j];
        const float *wr = w->w_fc + (size_t)j * w->h2;
        for (int k = 0; k < w->h2; ++k) {
            float v = h2[k] > 0.0f ? h2[k] : 0.0f; /* relu, wavernn.py:87 */
            a = fmaf(wr[k], v, a);
        }
        y[j] = 2.0f * orc_tanhf(a); /* dual_fc applied twice and summed, wavernn.py:89-92 */
    }
}

/* teacher-forced forward over a whole sequence (B,T,in) -> (B,T,fc); used by tests */
int orc_forward(int in_features, int h1n, int h2n, int fcn, const float *w_ih1,
                const float *w_hh1, const float *b_ih1, const float *b_hh1, const float *w_ih2,
                const float *w_hh2, const float *b_ih2, const float *b_hh2, const float *w_fc,
                const float *b_fc, const float *x, int B, int T, float *y, float *h1_io,
                float *h2_io)
{
    orc_weights w = {in_features, h1n, h2n, fcn, w_ih1, w_hh1, b_ih1, b_hh1,
                     w_ih2, w_hh2, b_ih2, b_hh2, w_fc, b_fc};
    int hm = h1n > h2n ? h1n : h2n;
#pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < B; ++b) {
        float *tmp = (float *)malloc(sizeof(float) * (size_t)hm);
        for (int t = 0; t < T; ++t)
            orc_predictor_step(&w, x + ((size_t)b * T + t) * in_features, h1_io + (size_t)b * h1n,
                               h2_io + (size_t)b * h2n, y + ((size_t)b * T + t) * fcn, tmp);
        free(tmp);
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * quantizers, instantiated for float32 and float64 codebooks (numpy computes in the
 * codebook's dtype: float32 residual - float64 codebook promotes to float64)
 * ---------------------------------------------------------------------------------------- */
#define REAL float
#define SUF(name) name##_f32
#include "fpc_oracle_vq.inc"
#undef REAL
#undef SUF
#define REAL double
#define SUF(name) name##_f64
#include "fpc_oracle_vq.inc"
#undef REAL
#undef SUF

/* dtype: 0 = float32 codebook, 1 = float64 codebook.  x is always float32 (it comes from a
 * torch float32 tensor, wavernn.py:219,230).  q_out is float64 so the caller sees exactly
 * what numpy returned before the cast on assignment (wavernn.py:220,232). */
int orc_vq_quantize(int dtype, const void *cb, int stages, const int *K, int ndim, const float *x,
                    int n, double *q_out, int32_t *idx_out)
{
    if (stages < 1 || stages > ORC_MAX_STAGES || ndim < 1 || ndim > ORC_MAX_DIM) return 1;
    for (int s = 0; s < stages; ++s)
        if (K[s] < ORC_SURVIVORS) return 2;
    int rc = 0;
#pragma omp parallel for schedule(static) if (n > 64)
    for (int i = 0; i < n; ++i) {
        if (dtype == 0)
            orc_quantize_mstage_f32((const float *)cb, stages, K, ndim, x + (size_t)i * ndim,
                                    q_out + (size_t)i * ndim, idx_out + (size_t)i * stages);
        else
            orc_quantize_mstage_f64((const double *)cb, stages, K, ndim, x + (size_t)i * ndim,
                                    q_out + (size_t)i * ndim, idx_out + (size_t)i * stages);
    }
    return rc;
}

int orc_scl_quantize(int dtype, const void *codes, int n_code, const float *x, int n, double *q_out,
                     int32_t *idx_out)
{
    if (n_code < 1) return 1;
    for (int i = 0; i < n; ++i) {
        if (dtype == 0)
            idx_out[i] = orc_scl_nearest_f32((const float *)codes, n_code, x[i], q_out + i);
        else
            idx_out[i] = orc_scl_nearest_f64((const double *)codes, n_code, x[i], q_out + i);
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * closed-loop encoder -- wavernn.py:165-256
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    /* above-threshold VQ (cfg['cb_path']), below-threshold VQ (cfg['bl_cb_path'], stages==0 -> '') */
    int vq_dtype, vq_stages, vq_K[ORC_MAX_STAGES];
    const void *vq;
    int bl_dtype, bl_stages, bl_K[ORC_MAX_STAGES];
    const void *bl;
    /* scalar codebooks (cfg['scl_cb_path'], cfg['bl_scl_cb_path'], n==0 -> '') */
    int scl_dtype, scl_n;
    const void *scl;
    int blscl_dtype, blscl_n;
    const void *blscl;
} orc_codebooks;

static int orc_encode_one(const orc_weights *w, const orc_codebooks *cb, const float *feat, int L,
                          float l1, float l2, const float *mask, int qtz, float *c_in, float *r,
                          float *r_qtz, float *r_under, float *ind1_o, float *ind2_o, int32_t *idx)
{
    const int C = w->in_features, F = w->fc, V = F - 1;
    int hm = w->h1 > w->h2 ? w->h1 : w->h2;
    float *h1 = (float *)calloc((size_t)w->h1, sizeof(float));
    float *h2 = (float *)calloc((size_t)w->h2, sizeof(float));
    float *tmp = (float *)malloc(sizeof(float) * (size_t)hm);
    float *x = (float *)calloc((size_t)C, sizeof(float)); /* frame 0 input: all zero (:177-178) */
    float *fo = (float *)malloc(sizeof(float) * (size_t)F);
    float *rs = (float *)malloc(sizeof(float) * (size_t)F);
    double *q = (double *)malloc(sizeof(double) * (size_t)F);
    for (int i = 0; i < L; ++i) {
        const float *ft = feat + (size_t)i * C;
        orc_predictor_step(w, x, h1, h2, fo, tmp);
        for (int j = 0; j < F; ++j) rs[j] = ft[j] - fo[j]; /* :196 */
        float m1, m2;
        if (mask == NULL) {
            float s = 0.0f;
            for (int j = 1; j < F; ++j) s = s + fabsf(rs[j]);
            m1 = fabsf(rs[0]) > l1 ? 1.0f : 0.0f; /* strict >, fp32 (:202) */
            m2 = s > l2 ? 1.0f : 0.0f;            /* (:206) */
            ind1_o[i] = m1;
            ind2_o[i] = m2;
        } else {
            m1 = mask[(size_t)i * 2 + 0];
            m2 = mask[(size_t)i * 2 + 1];
            ind1_o[i] = 0.0f; /* only the threshold branch fills the masks (:204,208) */
            ind2_o[i] = 0.0f;
        }
        int32_t *ix = idx + (size_t)i * 4;
        ix[0] = ix[1] = ix[2] = -1;
        ix[3] = (m1 != 0.0f ? 1 : 0) | (m2 != 0.0f ? 2 : 0);
        float *ro = r + (size_t)i * F, *rq = r_qtz + (size_t)i * F, *ru = r_under + (size_t)i * F;
        float *co = c_in + (size_t)i * C;
        if (qtz) {
            for (int j = 0; j < F; ++j) { ro[j] = rs[j]; rq[j] = 0.0f; ru[j] = 0.0f; } /* :197 */
            /* scalar quantizer for c0 (:217-225) */
            if (m1 != 0.0f) {
                if (cb->scl_dtype == 0) ix[0] = orc_scl_nearest_f32((const float *)cb->scl, cb->scl_n, rs[0], q);
                else ix[0] = orc_scl_nearest_f64((const double *)cb->scl, cb->scl_n, rs[0], q);
                rq[0] = (float)q[0];
            } else if (cb->blscl_n > 0) {
                if (cb->blscl_dtype == 0) ix[0] = orc_scl_nearest_f32((const float *)cb->blscl, cb->blscl_n, rs[0], q);
                else ix[0] = orc_scl_nearest_f64((const double *)cb->blscl, cb->blscl_n, rs[0], q);
                rq[0] = (float)q[0];
            }
            /* VQ for c1..c17 (:228-240) */
            if (m2 != 0.0f) {
                int32_t id[ORC_MAX_STAGES] = {-1, -1};
                if (cb->vq_dtype == 0) orc_quantize_mstage_f32((const float *)cb->vq, cb->vq_stages, cb->vq_K, V, rs + 1, q, id);
                else orc_quantize_mstage_f64((const double *)cb->vq, cb->vq_stages, cb->vq_K, V, rs + 1, q, id);
                for (int j = 0; j < V; ++j) rq[1 + j] = (float)q[j];
                ix[1] = id[0];
                ix[2] = cb->vq_stages > 1 ? id[1] : -1;
            } else if (cb->bl_stages > 0) {
                int32_t id[ORC_MAX_STAGES] = {-1, -1};
                if (cb->bl_dtype == 0) orc_quantize_mstage_f32((const float *)cb->bl, cb->bl_stages, cb->bl_K, V, rs + 1, q, id);
                else orc_quantize_mstage_f64((const double *)cb->bl, cb->bl_stages, cb->bl_K, V, rs + 1, q, id);
                for (int j = 0; j < V; ++j) rq[1 + j] = (float)q[j];
                ix[1] = id[0];
                ix[2] = cb->bl_stages > 1 ? id[1] : -1;
            }
            for (int j = 0; j < F; ++j) co[j] = fo[j] + rq[j]; /* :242 */
        } else {
            /* residual-generation mode for codebook training (:244-252) */
            ru[0] = rs[0] * (1.0f - m1);
            ro[0] = rs[0] * m1;
            for (int j = 1; j < F; ++j) {
                ru[j] = rs[j] * (1.0f - m2);
                ro[j] = rs[j] * m2;
            }
            for (int j = 0; j < F; ++j) { rq[j] = 0.0f; co[j] = fo[j] + ro[j]; }
        }
        for (int j = F; j < C; ++j) co[j] = ft[j]; /* pitch pass-through (:178) */
        memcpy(x, co, sizeof(float) * (size_t)C);  /* feedback: next input is this decoded frame */
    }
    free(h1); free(h2); free(tmp); free(x); free(fo); free(rs); free(q);
    return 0;
}

int orc_encode(int in_features, int h1n, int h2n, int fcn, const float *w_ih1, const float *w_hh1,
               const float *b_ih1, const float *b_hh1, const float *w_ih2, const float *w_hh2,
               const float *b_ih2, const float *b_hh2, const float *w_fc, const float *b_fc,
               int vq_dtype, int vq_stages, const int *vq_K, const void *vq, int bl_dtype,
               int bl_stages, const int *bl_K, const void *bl, int scl_dtype, int scl_n,
               const void *scl, int blscl_dtype, int blscl_n, const void *blscl, const float *feat,
               int B, int L, float l1, float l2, const float *mask, int qtz, float *c_in, float *r,
               float *r_qtz, float *r_under, float *ind1, float *ind2, int32_t *idx, int nthreads)
{
    if (fcn + 2 != in_features) return 1; /* 18 cepstra + 2 pitch (wavernn.py:178,196) */
    if (vq_stages < 0 || vq_stages > ORC_MAX_STAGES || bl_stages < 0 || bl_stages > ORC_MAX_STAGES) return 2;
    orc_weights w = {in_features, h1n, h2n, fcn, w_ih1, w_hh1, b_ih1, b_hh1,
                     w_ih2, w_hh2, b_ih2, b_hh2, w_fc, b_fc};
    orc_codebooks cb;
    memset(&cb, 0, sizeof(cb));
    cb.vq_dtype = vq_dtype; cb.vq_stages = vq_stages; cb.vq = vq;
    for (int s = 0; s < vq_stages; ++s) { cb.vq_K[s] = vq_K[s]; if (vq_K[s] < ORC_SURVIVORS) return 3; }
    cb.bl_dtype = bl_dtype; cb.bl_stages = bl_stages; cb.bl = bl;
    for (int s = 0; s < bl_stages; ++s) { cb.bl_K[s] = bl_K[s]; if (bl_K[s] < ORC_SURVIVORS) return 3; }
    cb.scl_dtype = scl_dtype; cb.scl_n = scl_n; cb.scl = scl;
    cb.blscl_dtype = blscl_dtype; cb.blscl_n = blscl_n; cb.blscl = blscl;
    if (qtz && (vq_stages < 1 || scl_n < 1)) return 4;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    const size_t C = (size_t)in_features, F = (size_t)fcn;
#pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < B; ++b) {
        orc_encode_one(&w, &cb, feat + (size_t)b * L * C, L, l1, l2,
                       mask ? mask + (size_t)b * L * 2 : NULL, qtz, c_in + (size_t)b * L * C,
                       r + (size_t)b * L * F, r_qtz + (size_t)b * L * F, r_under + (size_t)b * L * F,
                       ind1 + (size_t)b * L, ind2 + (size_t)b * L, idx + (size_t)b * L * 4);
    }
    return 0;
}

/* receiver side: replay the recurrence from transmitted indices (the reference's
 * own decoder, wavernn.py:367-379, is broken -- this follows the encoder's feedback equation
 * c[t] = f(c[t-1]) + dequant(idx[t]) so that decode(encode(x)) == c_in bit for bit) */
int orc_decode(int in_features, int h1n, int h2n, int fcn, const float *w_ih1, const float *w_hh1,
               const float *b_ih1, const float *b_hh1, const float *w_ih2, const float *w_hh2,
               const float *b_ih2, const float *b_hh2, const float *w_fc, const float *b_fc,
               const float *r_qtz, const float *pitch /*B,L,2*/, int B, int L, float *c_out)
{
    orc_weights w = {in_features, h1n, h2n, fcn, w_ih1, w_hh1, b_ih1, b_hh1,
                     w_ih2, w_hh2, b_ih2, b_hh2, w_fc, b_fc};
    const int C = in_features, F = fcn;
    int hm = h1n > h2n ? h1n : h2n;
#pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < B; ++b) {
        float *h1 = (float *)calloc((size_t)h1n, sizeof(float));
        float *h2 = (float *)calloc((size_t)h2n, sizeof(float));
        float *tmp = (float *)malloc(sizeof(float) * (size_t)hm);
        float *x = (float *)calloc((size_t)C, sizeof(float));
        float *fo = (float *)malloc(sizeof(float) * (size_t)F);
        for (int i = 0; i < L; ++i) {
            orc_predictor_step(&w, x, h1, h2, fo, tmp);
            float *co = c_out + ((size_t)b * L + i) * C;
            const float *rq = r_qtz + ((size_t)b * L + i) * F;
            for (int j = 0; j < F; ++j) co[j] = fo[j] + rq[j];
            for (int j = F; j < C; ++j) co[j] = pitch[((size_t)b * L + i) * 2 + (j - F)];
            memcpy(x, co, sizeof(float) * (size_t)C);
        }
        free(h1); free(h2); free(tmp); free(x); free(fo);
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * k-means / LBG -- cb_func.py
 * ---------------------------------------------------------------------------------------- */
/* find_nearest (cb_func.py:56-68): float64 direct-form distance in numpy's pairwise order
 * (float32 data - float64 codebook promotes), first minimum wins. */
int orc_find_nearest(const float *data, long N, const double *cb, int K, int ndim, int32_t *idx)
{
    if (ndim > ORC_MAX_DIM) return 1;
#pragma omp parallel for schedule(static)
    for (long i = 0; i < N; ++i) {
        const float *x = data + (size_t)i * ndim;
        double best = INFINITY;
        int bi = 0;
        for (int k = 0; k < K; ++k) {
            double d = orc_dist_f64(cb + (size_t)k * ndim, x, ndim);
            if (d < best || k == 0) { best = d; bi = k; }
        }
        idx[i] = bi;
    }
    return 0;
}

/* update (cb_func.py:71-100): one Lloyd iteration.  Sums are accumulated in float64 in data
 * order, exactly as the reference's Python loop does; empty clusters collapse to the zero
 * vector through sum/(count+1e-20).  stats = {min count, max count, #empty, w2}. */
int orc_kmeans_update(const float *data, long N, const double *cb, int K, int ndim, double *cb_out,
                      int32_t *idx_out /* may be NULL */, double *counts_out /* K, may be NULL */,
                      double *stats /* 4, may be NULL */)
{
    int32_t *idx = idx_out ? idx_out : (int32_t *)malloc(sizeof(int32_t) * (size_t)N);
    int rc = orc_find_nearest(data, N, cb, K, ndim, idx);
    if (rc) { if (!idx_out) free(idx); return rc; }
    double *count = (double *)calloc((size_t)K, sizeof(double));
    memset(cb_out, 0, sizeof(double) * (size_t)K * ndim);
    for (long i = 0; i < N; ++i) {
        int n = idx[i];
        count[n] += 1.0;
        for (int d = 0; d < ndim; ++d) cb_out[(size_t)n * ndim + d] += (double)data[(size_t)i * ndim + d];
    }
    double mn = INFINITY, mx = -INFINITY, empty = 0.0, w2 = 0.0;
    for (int k = 0; k < K; ++k) {
        for (int d = 0; d < ndim; ++d) cb_out[(size_t)k * ndim + d] /= (count[k] + 1e-20);
        if (count[k] < mn) mn = count[k];
        if (count[k] > mx) mx = count[k];
        if (count[k] == 0.0) empty += 1.0;
        double f = count[k] / (double)N;
        w2 += f * f;
    }
    if (stats) { stats[0] = mn; stats[1] = mx; stats[2] = empty; stats[3] = w2; }
    if (counts_out) memcpy(counts_out, count, sizeof(double) * (size_t)K);
    free(count);
    if (!idx_out) free(idx);
    return 0;
}

/* quantize (cb_func.py:103-112): nearest-centroid gather, float64 out */
int orc_kmeans_quantize(const float *data, long N, const double *cb, int K, int ndim, double *q_out,
                        int32_t *idx_out)
{
    int32_t *idx = idx_out ? idx_out : (int32_t *)malloc(sizeof(int32_t) * (size_t)N);
    int rc = orc_find_nearest(data, N, cb, K, ndim, idx);
    if (!rc)
        for (long i = 0; i < N; ++i)
            memcpy(q_out + (size_t)i * ndim, cb + (size_t)idx[i] * ndim, sizeof(double) * (size_t)ndim);
    if (!idx_out) free(idx);
    return rc;
}

/* The same three functions for FLOAT64 training vectors: the data of a later training stage is
 * r = quantize(cb, r) - r in float64 (train_cb.py:200), and numpy then evaluates (data - codebook) ** 2, the sums of
 * `update` and the seed mean in float64 on it. */
static double orc_dist_dd(const double *c, const double *x, int n)
{
    double e[ORC_MAX_DIM];
    for (int d = 0; d < n; ++d) {
        double t = x[d] - c[d];
        e[d] = t * t;
    }
    if (n < 8) {
        double res = 0.0;
        for (int d = 0; d < n; ++d) res = res + e[d];
        return res;
    }
    double r[8];
    for (int j = 0; j < 8; ++j) r[j] = e[j];
    int i;
    for (i = 8; i < n - (n % 8); i += 8)
        for (int j = 0; j < 8; ++j) r[j] = r[j] + e[i + j];
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) res = res + e[i];
    return res;
}

int orc_find_nearest_d(const double *data, long N, const double *cb, int K, int ndim, int32_t *idx)
{
    if (ndim > ORC_MAX_DIM) return 1;
#pragma omp parallel for schedule(static)
    for (long i = 0; i < N; ++i) {
        const double *x = data + (size_t)i * ndim;
        double best = INFINITY;
        int bi = 0;
        for (int k = 0; k < K; ++k) {
            double d = orc_dist_dd(cb + (size_t)k * ndim, x, ndim);
            if (d < best || k == 0) { best = d; bi = k; }
        }
        idx[i] = bi;
    }
    return 0;
}

int orc_kmeans_update_d(const double *data, long N, const double *cb, int K, int ndim, double *cb_out,
                        int32_t *idx_out /* may be NULL */, double *counts_out /* K, may be NULL */,
                        double *stats /* 4, may be NULL */)
{
    int32_t *idx = idx_out ? idx_out : (int32_t *)malloc(sizeof(int32_t) * (size_t)N);
    int rc = orc_find_nearest_d(data, N, cb, K, ndim, idx);
    if (rc) { if (!idx_out) free(idx); return rc; }
    double *count = (double *)calloc((size_t)K, sizeof(double));
    memset(cb_out, 0, sizeof(double) * (size_t)K * ndim);
    for (long i = 0; i < N; ++i) {
        int n = idx[i];
        count[n] += 1.0;
        for (int d = 0; d < ndim; ++d) cb_out[(size_t)n * ndim + d] += data[(size_t)i * ndim + d];
    }
    double mn = INFINITY, mx = -INFINITY, empty = 0.0, w2 = 0.0;
    for (int k = 0; k < K; ++k) {
        for (int d = 0; d < ndim; ++d) cb_out[(size_t)k * ndim + d] /= (count[k] + 1e-20);
        if (count[k] < mn) mn = count[k];
        if (count[k] > mx) mx = count[k];
        if (count[k] == 0.0) empty += 1.0;
        double f = count[k] / (double)N;
        w2 += f * f;
    }
    if (stats) { stats[0] = mn; stats[1] = mx; stats[2] = empty; stats[3] = w2; }
    if (counts_out) memcpy(counts_out, count, sizeof(double) * (size_t)K);
    free(count);
    if (!idx_out) free(idx);
    return 0;
}

int orc_kmeans_quantize_d(const double *data, long N, const double *cb, int K, int ndim, double *q_out,
                          int32_t *idx_out)
{
    int32_t *idx = idx_out ? idx_out : (int32_t *)malloc(sizeof(int32_t) * (size_t)N);
    int rc = orc_find_nearest_d(data, N, cb, K, ndim, idx);
    if (!rc)
        for (long i = 0; i < N; ++i)
            memcpy(q_out + (size_t)i * ndim, cb + (size_t)idx[i] * ndim, sizeof(double) * (size_t)ndim);
    if (!idx_out) free(idx);
    return rc;
}

int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
