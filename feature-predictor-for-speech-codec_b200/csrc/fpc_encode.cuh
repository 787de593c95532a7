// fpc_encode.cuh -- launch parameters shared by the frame-step kernels and the C-ABI layer.
#pragma once
#include "fpc_common.cuh"

namespace fpc {

enum { kModeResidual = 0, kModeQuantize = 1, kModeDecode = 2 };

struct EncodeParams {
    const float *wstream;   // packed fp32 weights (fpc_pack_weights)
    const char *cb;         // packed codebooks (fpc_pack_codebooks) or null
    const float *feat;      // (B,L,20)   [decode: unused]
    const float *mask;      // (B,L,2) or null
    const float *rq_in;     // decode: (B,L,18)
    const float *pitch_in;  // decode: (B,L,2)
    float *c_in, *r, *r_qtz, *r_under, *ind1, *ind2;
    int32_t *idx;
    int B, L, mode;
    float l1, l2;
    int ntiles;
    int f0, f1;             // frame range of this launch, [f0, f1) within L (whole utterance: 0, L)
    void *state;            // per-tile recurrent state carried between launches of consecutive ranges, or null
    long long *prof;        // optional debug buffer: per-phase cycle totals of CTA 0 (fpc_debug_set_phase_buffer)
};

// launch plan: consecutive utterance ranges of the batch, each with its own tile height (fpc_encode_fp32.cu)
constexpr int kMaxSegments = 3;
struct EncodeSegment { int height, first, count; };
int plan_segments(int B, int sms, const int *heights, int nh, double fixed, EncodeSegment *out);
int encode_plan(int B, int precision, int sms, int *segments);   // (height, first, count) triples; returns their number
EncodeParams segment_params(const EncodeParams &P, int first, int count);

int run_encode_fp32(EncodeParams P, cudaStream_t st, int force_tu);
size_t encode_fp32_state_bytes(int B);   // size of EncodeParams::state for a batch of B utterances
// bf16 tensor-core predictor (fpc_encode_bf16.cu); wstream then points at the bf16 image
int run_encode_bf16(EncodeParams P, cudaStream_t st, int force_nu);
size_t packed_bf16_bytes();
size_t encode_bf16_state_bytes(int B);
int pack_weights_bf16(const fpc_weights *w, void *d_packed, cudaStream_t st);
int num_sms();
extern long long *g_phase_buffer;   // device pointer or null
enum { kPhGru = 0, kPhFc = 1, kPhScalar = 2, kPhVq = 3, kPhOut = 4, kPhFrames = 5, kPhVqDbg = 6 /* 10 counters of the search */,
       kPhWaitX = 16 /* GEMM warps waiting for the tail */, kPhWaitH = 17 /* tail waiting for h2 */, kPhCount = 32 };

}  // namespace fpc
