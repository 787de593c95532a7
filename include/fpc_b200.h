/*
 * fpc_b200.h -- C ABI of the B200-native closed-loop predictive-coding path.
 *
 * The reference (haiciyang/Feature-predictor-for-speech-codec) is pure Python and has no
 * FFI of its own; its plug-in seam is the set of Python call signatures listed below
 * (SURVEY.md section 8b).  Each entry point here is what a binding for that call site binds
 * to.  Paths are relative to /root/reference/src.
 *
 *   fpc_encode                 <- models/wavernn.py:165-256   Wavernn.encoder (whole frame loop,
 *                                 including the injected vq_quantize / scl_quantize calls at
 *                                 :219-240 and the feedback at :242,252)
 *   fpc_encode_host            <- synthesis_qtz.py:149-160, generate_qtz_features.py:55-70   the host sequence
 *                                 feat.to('cuda') -> encoder -> results .cpu(), copies overlapped with the kernel
 *   fpc_decode                 <- models/wavernn.py:367-379   Wavernn.decoder (receiver replay)
 *   fpc_pack_weights           <- models/wavernn.py:37-38,48-52  parameters of rnn1/rnn2/dual_fc
 *   fpc_pack_codebooks         <- quantization/vq_func.py:141,171  the np.load of the four files
 *   fpc_vq_quantize_packed     <- quantization/vq_func.py:134-164  vq_quantize / quantize_mstage
 *   fpc_scl_quantize           <- quantization/vq_func.py:167-185  scl_quantize
 *   fpc_index_histogram        <- models/wavernn.py:189,221-240    cb_tot accumulation
 *   fpc_kmeans_assign_accumulate <- quantization/cb_func.py:56-68,82-86  find_nearest + sums
 *   fpc_kmeans_accumulate_ordered <- quantization/cb_func.py:82-86  the same sums, in data order (bit-exact)
 *   fpc_kmeans_finalize        <- quantization/cb_func.py:88-97    divide, cluster statistics
 *   fpc_kmeans_gather          <- quantization/cb_func.py:103-112  quantize
 *   fpc_ceps2lpc               <- ceps2lpc/ceps2lpc_vct.py:122-162 ceps2lpc_v
 *   fpc_compact_rows           <- train_cb.py:177-187              training sets from the qtz=False outputs
 *   fpc_kmeans_stage_residual  <- train_cb.py:199-200,210-211      r = quantize(cb, r) - r for the next stage
 *   fpc_dequantize, fpc_pack_frames, fpc_unpack_frames  (no counterpart: the reference has no working receiver,
 *                                 models/wavernn.py:367-379, and defines no bitstream)
 *
 * Conventions
 *   - Plain C: pointers, sizes, a stream handle.  No torch / C++ types cross this boundary.
 *   - Every pointer named d_* is DEVICE memory owned by the caller; the library never
 *     allocates or frees device memory and never synchronises the stream.  The caller keeps
 *     buffers alive until the stream has drained.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *   - Every function returns an fpc_status; 0 is success.  Nothing throws, nothing prints.
 *   - There is no CPU fallback: without a CUDA device every compute entry point returns
 *     FPC_ERR_CUDA.
 */
#ifndef FPC_B200_H_
#define FPC_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
/* the library is built with -fvisibility=hidden; exactly the declarations below are exported */
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

#define FPC_VERSION 100 /* 0.1.0 */

typedef enum {
    FPC_OK = 0,
    FPC_ERR_ARG = 1,          /* null pointer / negative size */
    FPC_ERR_SHAPE = 2,        /* dimensions this build does not implement */
    FPC_ERR_CODEBOOK = 3,     /* bad codebook (stages, entries, dtype) */
    FPC_ERR_WORKSPACE = 4,    /* workspace / packed buffer too small */
    FPC_ERR_CUDA = 5,         /* a CUDA runtime call failed; see fpc_last_cuda_error() */
    FPC_ERR_UNSUPPORTED = 6   /* precision / mode not built */
} fpc_status;

/* arithmetic of the predictor's dense contractions */
typedef enum {
    FPC_PREC_FP32 = 0, /* FFMA, canonical summation order: bit-exact against oracle/ */
    FPC_PREC_BF16 = 1  /* bf16 operands on tcgen05 tensor cores, fp32 accumulate in TMEM */
} fpc_precision;

/* element type of a codebook file; vq_func.py:18 computes in the file's dtype */
typedef enum { FPC_F32 = 0, FPC_F64 = 1 } fpc_dtype;

/* Fixed model geometry of this build (wavernn.py:24 with the arguments of
 * synthesis_qtz.py:79-85 / BASELINE.json): in_features 20, gru_units1 384, gru_units2 128,
 * fc_units 18, code_dims 17, SURVIVORS 5 (vq_func.py:3). */
#define FPC_IN_FEATURES 20
#define FPC_GRU1 384
#define FPC_GRU2 128
#define FPC_FC 18
#define FPC_CODE_DIMS 17
#define FPC_SURVIVORS 5
#define FPC_MAX_VQ_ENTRIES 1024
#define FPC_MAX_SCL_ENTRIES 256

int fpc_version(void);
const char *fpc_status_string(int status);
/* cudaError_t of the most recent failing CUDA call on this thread (0 if none) */
int fpc_last_cuda_error(void);
/* number of kernel launches issued by this library since load (for bench.py's gpu_launches) */
unsigned long long fpc_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * predictor weights
 * ------------------------------------------------------------------------------------------- */
/* Parameters in torch's own layout, all float32 on the device (state_dict of the module):
 * GRU gate row order r, z, n. */
typedef struct {
    const float *w_ih1; /* (3*384, 20)   rnn1.weight_ih_l0 */
    const float *w_hh1; /* (3*384, 384)  rnn1.weight_hh_l0 */
    const float *b_ih1; /* (3*384)       rnn1.bias_ih_l0   */
    const float *b_hh1; /* (3*384)       rnn1.bias_hh_l0   */
    const float *w_ih2; /* (3*128, 384)  rnn2.weight_ih_l0 */
    const float *w_hh2; /* (3*128, 128)  rnn2.weight_hh_l0 */
    const float *b_ih2; /* (3*128)       rnn2.bias_ih_l0   */
    const float *b_hh2; /* (3*128)       rnn2.bias_hh_l0   */
    const float *w_fc;  /* (18, 128)     dual_fc.0.weight  */
    const float *b_fc;  /* (18)          dual_fc.0.bias    */
} fpc_weights;

/* bytes of the kernel-side weight image for a precision */
size_t fpc_packed_weights_bytes(int precision);
/* re-tile the parameters into the streaming order of the frame-step kernel (one kernel launch) */
int fpc_pack_weights(const fpc_weights *w, int precision, void *d_packed, size_t packed_bytes, void *stream);

/* ---------------------------------------------------------------------------------------------
 * codebooks
 * ------------------------------------------------------------------------------------------- */
/* The four files Wavernn.encoder reads through cfg (wavernn.py:219-237), already on the device
 * in their on-disk layout: VQ (stages, K, 17) row-major, scalar (n) -- stages in {1,2}, all
 * stages of one file share K (np.load gives a rectangular array), 5 <= K <= 1024, n <= 256.
 * stages == 0 / n == 0 means the cfg path was '' (below-threshold residual is dropped). */
typedef struct {
    const void *vq;      int vq_dtype;     int vq_stages;  int vq_entries;      /* cfg['cb_path'] */
    const void *bl_vq;   int bl_vq_dtype;  int bl_vq_stages; int bl_vq_entries; /* cfg['bl_cb_path'] */
    const void *scl;     int scl_dtype;    int scl_entries;                     /* cfg['scl_cb_path'] */
    const void *bl_scl;  int bl_scl_dtype; int bl_scl_entries;                  /* cfg['bl_scl_cb_path'] */
} fpc_codebooks;

size_t fpc_packed_codebooks_bytes(void);
int fpc_pack_codebooks(const fpc_codebooks *cb, void *d_packed, size_t packed_bytes, void *stream);

/* ---------------------------------------------------------------------------------------------
 * closed-loop encoder  (Wavernn.encoder, wavernn.py:165-256)
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    /* inputs */
    const float *d_feat;   /* (B, L, 20) normalised features; read-only */
    const float *d_mask;   /* (B, L, 2) external indicator or NULL -> thresholds (wavernn.py:201-212) */
    int B, L;
    float l1, l2;          /* thresholds on |r0| and sum|r1..17|, strict >, fp32 (:202,206) */
    int qtz;               /* 1: quantise and feed back (:214-242); 0: residual generation (:244-252) */
    /* outputs, all written in full by the call (the callee of the reference allocates zeros) */
    float *d_c_in;         /* (B, L, 20) decoded frames  == c_in[:,1:,:] (:256) */
    float *d_r;            /* (B, L, 18) raw residual (qtz) or masked residual (!qtz) */
    float *d_r_qtz;        /* (B, L, 18) quantised residual (zeros when !qtz) */
    float *d_r_under;      /* (B, L, 18) below-threshold residual (!qtz) / zeros; may be NULL */
    float *d_ind1;         /* (B, L) 0/1; may be NULL */
    float *d_ind2;         /* (B, L) 0/1; may be NULL */
    int32_t *d_idx;        /* (B, L, 4) [scalar idx, vq stage-1 (or below) idx, vq stage-2 idx,
                              flags bit0=ind1 bit1=ind2]; -1 = nothing coded; may be NULL */
} fpc_encode_io;

/* scratch the encoder needs for a batch of B utterances (0 is a valid answer) */
size_t fpc_encode_workspace_bytes(int B, int L, int precision);

int fpc_encode(const void *d_packed_weights, const void *d_packed_codebooks, const fpc_encode_io *io,
               int precision, void *d_workspace, size_t workspace_bytes, void *stream);

/* The launch plan fpc_encode uses for a batch of B utterances on a device with `sms` multiprocessors (sms <= 0: the
 * current device): up to 3 consecutive utterance ranges, each launched with its own tile height (utterances per
 * CTA) so that no partial wave of tall tiles is left (utterances are independent, wavernn.py:217,228, so results do
 * not depend on the plan).  Writes (tile height, first utterance, count) triples to `segments` (room for 9 ints) and
 * returns their number, or a negative status.  Host-only arithmetic; no CUDA call when sms > 0. */
int fpc_encode_plan(int B, int precision, int sms, int *segments);

/* Threading: the library keeps per-device one-time state (kernel attributes, SM count, the copy pipeline of
 * fpc_encode_host) indexed by the CUDA device current at the call, so one process may drive several GPUs.  Calls for
 * DIFFERENT devices may come from different host threads; calls for the same device must be issued from one thread
 * at a time.  fpc_encode_host orders itself after the previous fpc_encode_host call on the same device even when
 * the two calls use different streams (they may share a workspace). */

/* ---- host-buffer form of the closed loop ------------------------------------------------------------------------
 * Replaces the reference's host sequence  feat.to('cuda') -> model_f.encoder(...) -> results .cpu()
 * (synthesis_qtz.py:149-160, generate_qtz_features.py:55-70) for callers whose features and results live in host
 * memory.  The utterances are cut along TIME into `chunks` frame ranges; the upload of range c+1, the kernel of
 * range c (the recurrent state is carried between launches in the workspace) and the download of range c-1 run
 * concurrently on three streams, so the PCIe copies hide behind the arithmetic.  Results are bit-identical to
 * fpc_encode on the same inputs.  Host buffers should be pinned (cudaHostAlloc / torch pin_memory); pageable memory
 * works but serialises the copies.  Thresholds only (no external mask).  The call is asynchronous: the results are
 * complete when `stream` has been synchronised. */
typedef struct fpc_encode_host_io {
    const float *h_feat;   /* (B, L, 20) float32, host */
    int B, L;
    float l1, l2;
    int qtz;
    float *h_c_in;         /* (B, L, 20) host outputs; any of them may be NULL (not downloaded) */
    float *h_r;            /* (B, L, 18) */
    float *h_r_qtz;        /* (B, L, 18) */
    float *h_r_under;      /* (B, L, 18) */
    float *h_ind1;         /* (B, L) */
    float *h_ind2;         /* (B, L) */
    int32_t *h_idx;        /* (B, L, 4) */
    unsigned long long *h_hist;   /* FPC_HIST_TOTAL counters: cb_tot of the call (layout: fpc_index_histogram); may be NULL */
} fpc_encode_host_io;

/* device scratch fpc_encode_host needs: the device copies of the input, of every output, and the carried state */
size_t fpc_encode_host_workspace_bytes(int B, int L, int precision);

/* chunks <= 0: chosen by the library (48+ frames per range, at most 16 ranges). */
int fpc_encode_host(const void *d_packed_weights, const void *d_packed_codebooks, const fpc_encode_host_io *io,
                    int precision, int chunks, void *d_workspace, size_t workspace_bytes, void *stream);

/* receiver side: c[t] = predictor(c[t-1]) + r_qtz[t], pitch passed through.
 * d_r_qtz (B,L,18), d_pitch (B,L,2) -> d_c_out (B,L,20). */
int fpc_decode(const void *d_packed_weights, const float *d_r_qtz, const float *d_pitch, int B, int L,
               float *d_c_out, int precision, void *d_workspace, size_t workspace_bytes, void *stream);

/* cb_tot (wavernn.py:189,221-240): five count tables from the index record.
 * d_hist: 256 + 256 + 1024 + 1024 + 1024 uint64 counters at the offsets below, zeroed by the call. */
#define FPC_HIST_SCL 0
#define FPC_HIST_BL_SCL 256
#define FPC_HIST_VQ1 512
#define FPC_HIST_VQ2 1536
#define FPC_HIST_BL_VQ 2560
#define FPC_HIST_TOTAL 3584
int fpc_index_histogram(const int32_t *d_idx, long n_frames, unsigned long long *d_hist, void *stream);

/* ---------------------------------------------------------------------------------------------
 * stand-alone quantisers  (vq_func.py:134-185)
 * ------------------------------------------------------------------------------------------- */
/* vq_quantize on a codebook file that fpc_pack_codebooks has already placed in a packed image
 * (the host mirror caches one image per file, where the reference re-reads the .npy on every
 * call, vq_func.py:141).  which: 0 = the cfg['cb_path'] slot of the image, 1 = the
 * cfg['bl_cb_path'] slot.  dtype / stages must equal what the slot was packed with.
 * d_x (n,17) float32; d_q (n,17) of `dtype` (numpy returns the codebook dtype, vq_func.py:161-164);
 * d_idx (n,stages) int32. */
int fpc_vq_quantize_packed(const float *d_x, long n, const void *d_packed_codebooks, int which, int dtype,
                           int stages, void *d_q, int32_t *d_idx, void *stream);
/* d_x (n) float32; d_codes (n_code) of `dtype`; d_q (n) of `dtype`; d_idx (n) int32 */
int fpc_scl_quantize(const float *d_x, long n, const void *d_codes, int dtype, int n_code, void *d_q,
                     int32_t *d_idx, void *stream);

/* ---------------------------------------------------------------------------------------------
 * k-means codebook learning  (cb_func.py)
 * ------------------------------------------------------------------------------------------- */
/* One assignment pass of cb_func.update over this rank's shard: nearest centroid in float64
 * direct form (first minimum), then per-centroid float64 sums and counts ADDED into d_sums (K,17)
 * and d_counts (K) (caller zeroes them; with several ranks the caller all-reduces them before
 * fpc_kmeans_finalize).  d_idx (N) int32 may be NULL.  d_workspace: fpc_kmeans_workspace_bytes(N, K) bytes of
 * scratch for codebooks below 512 entries (replicated accumulation tables that keep the float64 atomics of small
 * codebooks off a handful of addresses); NULL is allowed and only slower. */
size_t fpc_kmeans_workspace_bytes(long N, int K);
int fpc_kmeans_assign_accumulate(const float *d_data, long N, const double *d_cb, int K, double *d_sums,
                                 double *d_counts, int32_t *d_idx, void *d_workspace, size_t workspace_bytes,
                                 void *stream);
/* codebook = sums / (counts + 1e-20); stats[5] = {min count, max count, #empty, sum (count/N)^2, N}.
 * n_total <= 0: N is taken as the sum of the counts (what nb_vectors is), which saves a host round trip per
 * iteration when the counts were all-reduced on the device. */
int fpc_kmeans_finalize(const double *d_sums, const double *d_counts, int K, double n_total, double *d_cb_out,
                        double *d_stats, void *stream);

/* The same on ONE accumulator  d_acc = [sums (K,17) | counts (K)]  -- K * 18 float64, the layout a single all-reduce
 * message wants: fpc_kmeans_assign_accumulate takes d_acc and d_acc + 17 K, ranks all-reduce d_acc in place, and this
 * call divides, writes the statistics and leaves d_acc ZEROED for the next Lloyd iteration (cb_func.py:71-100 with no
 * memset / pack / unpack launches between the iterations).  n_total <= 0: the sum of the counts. */
int fpc_kmeans_finalize_acc(double *d_acc, int K, double n_total, double *d_cb_out, double *d_stats, void *stream);
/* The same for float64 vectors: the training data of a later stage, r = quantize(cb, r) - r, which the reference keeps
 * in float64 (train_cb.py:200, cb_func.py:56-100).  The screen runs on the float32 rounding of a vector (inside its
 * slack), near-ties are decided and the sums accumulated with the float64 vector.  CUDA-core kernel for every K. */
int fpc_kmeans_assign_accumulate_f64(const double *d_data, long N, const double *d_cb, int K, double *d_sums,
                                     double *d_counts, int32_t *d_idx, void *d_workspace, size_t workspace_bytes,
                                     void *stream);
/* d_carry[17] += column sums of the float32 rows IN ROW ORDER, in float32: the additions np.mean(data, 0) performs for
 * centroid 0 of vq_train (cb_func.py:34; NumPy reduces a C-contiguous (N,17) float32 array over axis 0 row by row in
 * float32).  Serial by nature: one CTA, ~8 ns per row.  Ranks of a sharded data set call it in rank order on the carry
 * of the previous rank. */
int fpc_kmeans_colsum_f32(const float *d_data, long N, float *d_carry, void *stream);
int fpc_kmeans_colsum_f64(const double *d_data, long N, double *d_carry, void *stream);   /* float64 rows, float64 additions */
/* The accumulation loop of cb_func.update IN DATA ORDER (cb_func.py:82-86: `count[n] += 1; sum[n] += data[i]` for
 * i = 0, 1, ...): given the indices d_idx (N) that fpc_kmeans_assign_accumulate wrote, ADDS to d_sums (K,17) the
 * vectors of each centroid one after the other in ascending row order, in float64, and their number to d_counts (K).
 * Where the assign kernel's own sums (float64 atomics, scheduling order) match the reference to ~1e-16 relative, these
 * match it bit for bit and are the same from run to run.  A stable counting sort of the row numbers by centroid plus
 * one warp per centroid; K <= 2048, N < 2^31.  d_data: float32 rows, or float64 rows when data_is_f64.
 * d_workspace: fpc_kmeans_ordered_workspace_bytes(N, K) bytes (row permutation + tile counts). */
size_t fpc_kmeans_ordered_workspace_bytes(long N, int K);
int fpc_kmeans_accumulate_ordered(const void *d_data, int data_is_f64, long N, const int32_t *d_idx, int K,
                                  double *d_sums, double *d_counts, void *d_workspace, size_t workspace_bytes,
                                  void *stream);
/* q[i] = cb[idx[i]]  (cb_func.quantize after fpc_kmeans_assign_accumulate filled idx) */
int fpc_kmeans_gather(const double *d_cb, int K, const int32_t *d_idx, long N, double *d_q, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Codebook-training data path (train_cb.py:141-211), on the device.
 * fpc_compact_rows keeps, in order, the rows of a (n_rows, src_stride) float32 array whose columns
 * [col0, col0+ncols) do not sum (in absolute value) to zero, and writes those columns densely:
 *   vector set  train_cb.py:187   r = [r[i] for i in range(len(r)) if sum(abs(r[i])) != 0]   (col0 = 1, ncols = 17)
 *   scalar set  train_cb.py:177   [k for k in r[:,:,0].flatten() if k != 0]                   (col0 = 0, ncols = 1)
 * d_dst must hold n_rows * ncols floats; *d_count (device) receives the number of rows kept.
 * ------------------------------------------------------------------------------------------- */
size_t fpc_compact_workspace_bytes(long n_rows);
int fpc_compact_rows(const float *d_src, long n_rows, int src_stride, int col0, int ncols, float *d_dst,
                     long long *d_count, void *d_workspace, size_t workspace_bytes, void *stream);
/* next[i] = (float)(cb[idx[i]] - data[i]) -- the next stage's training vectors (train_cb.py:200,211; the sign is the
 * reference's, flipped relative to the encoder's x - csum) */
int fpc_kmeans_stage_residual(const double *d_cb, int K, const int32_t *d_idx, const float *d_data, long N,
                              float *d_next, void *stream);
/* next[i] = cb[idx[i]] - data[i] KEPT IN FLOAT64, as train_cb.py:200 leaves it (float64 codebook minus float32 data of
 * the first stage, or minus the float64 data of a later one: data_is_f64) */
int fpc_kmeans_stage_residual_f64(const double *d_cb, int K, const int32_t *d_idx, const void *d_data, int data_is_f64,
                                  long N, double *d_next, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Receiver side and wire format of the index record idx (n_frames, 4) that fpc_encode writes.
 * fpc_dequantize rebuilds r_qtz (n_frames, 18) from the record, bit-identical to the encoder's own r_qtz
 * (scalar table entry for c0; csum = 0 + CB0[i1] (+ CB1[i2]) in the codebook dtype, cast to float32);
 * fpc_dequantize + fpc_decode is the decoder.
 * fpc_pack_frames / fpc_unpack_frames: one 32-bit word per frame -- bit 0 ind1, bit 1 ind2, bits 2-9 scalar index,
 * bits 10-19 VQ index (stage 1 / below book), bits 20-29 VQ stage-2 index.  -1 entries ("nothing coded") are
 * restored from the codebook set, which both sides share.
 * ------------------------------------------------------------------------------------------- */
int fpc_dequantize(const void *d_packed_codebooks, const int32_t *d_idx, long n_frames, float *d_r_qtz, void *stream);
int fpc_pack_frames(const int32_t *d_idx, long n_frames, uint32_t *d_words, void *stream);
int fpc_unpack_frames(const void *d_packed_codebooks, const uint32_t *d_words, long n_frames, int32_t *d_idx,
                      void *stream);

/* ---------------------------------------------------------------------------------------------
 * self-test of the tensor-core plumbing (tcgen05.mma / TMEM) the bf16 predictor is built on:
 * d_out (128,N) f32 = A (128,K) bf16 x B (N,K) bf16 ^T.  16 <= N <= 256, N % 16 == 0, K % 16 == 0.
 * ------------------------------------------------------------------------------------------- */
/* ---------------------------------------------------------------------------------------------
 * cepstrum -> LPC  (ceps2lpc/ceps2lpc_vct.py:122-162 ceps2lpc_v, the step after Wavernn.encoder in
 * synthesis_qtz.py:158-160 and generate_qtz_features.py:61-64)
 * d_ceps (n, stride) float32, the first 18 columns of a row are the de-normalised cepstrum; d_lpc (n,16);
 * d_err (n) final prediction error or NULL; d_rc (n,16) reflection coefficients or NULL.
 * ------------------------------------------------------------------------------------------- */
int fpc_ceps2lpc(const float *d_ceps, long n, int stride, float *d_lpc, float *d_err, float *d_rc, void *stream);

/* debug aid: d_buf = 8 int64 counters in device memory (zeroed by the caller) or NULL to switch off;
 * CTA 0 of fpc_encode (fp32) adds the SM cycles it spent in [GRU, FC, thresholds+scalar, VQ, feedback] and the
 * number of frames.  Used by tools/phase_profile.py; not part of the reference-facing surface. */
int fpc_debug_set_phase_buffer(void *d_buf);

int fpc_selftest_umma(const void *d_a_bf16, const void *d_b_bf16, int N, int K, float *d_out, void *stream);

/* Self-test of the tensor-core distance screen (csrc/fpc_tc.cuh): scores[v][k] = ||c_k||^2 - 2 <x_v, c_k> of n <= 128
 * vectors against K <= 1024 float64 centroids exactly as the screen of fpc_kmeans_assign_accumulate computes them
 * (fp16-pair operands, tcgen05, fp32 accumulators), so a test can measure their error against float64.
 * d_out: n x Kp floats (Kp = K rounded up to 128); d_pack: fpc_kmeans_workspace_bytes(n, K) bytes of scratch. */
int fpc_selftest_tc_scores(const float *d_x, int n, const double *d_cb, int K, float *d_out, void *d_pack, void *stream);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* FPC_B200_H_ */
