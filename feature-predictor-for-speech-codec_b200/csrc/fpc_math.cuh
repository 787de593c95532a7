// fpc_math.cuh -- canonical fp32 transcendental functions of the predictor.
//
// These are the device-side twins of orc_exp_core / orc_sigmoidf / orc_tanhf in
// the CPU restatement (fpc_oracle.c).  They use only IEEE-754 round-to-nearest add / mul / fma / div
// (explicit intrinsics, never contracted), so the fp32 CUDA path reproduces the oracle bit
// for bit.  torch's CPU sigmoid/tanh (reference: torch.nn.GRU inside wavernn.py:71,76 and
// nn.Tanh at :51) differ from these by <= 2 ulp, which is covered by the 1e-4 feature
// tolerance of BASELINE.json and measured in tests/test_oracle_golden.py.
#pragma once
#include <cuda_runtime.h>

namespace fpc {

__device__ __forceinline__ float exp_core(float x)  // x in [-87, 88]
{
    const float magic = 12582912.0f;  // 1.5 * 2^23
    float t = __fmaf_rn(x, 1.44269504088896341f, magic);
    float n = __fsub_rn(t, magic);
    float r = __fmaf_rn(n, -0.693359375f, x);
    r = __fmaf_rn(n, 2.12194440e-4f, r);
    float p = 1.9875691500e-4f;
    p = __fmaf_rn(p, r, 1.3981999507e-3f);
    p = __fmaf_rn(p, r, 8.3334519073e-3f);
    p = __fmaf_rn(p, r, 4.1665795894e-2f);
    p = __fmaf_rn(p, r, 1.6666665459e-1f);
    p = __fmaf_rn(p, r, 5.0000001201e-1f);
    float y = __fadd_rn(__fmaf_rn(p, __fmul_rn(r, r), r), 1.0f);
    int ni = (int)n;
    return __int_as_float(__float_as_int(y) + ni * (1 << 23));
}

__device__ __forceinline__ float sigmoid_c(float x)
{
    float a = -x;
    if (a > 88.0f) a = 88.0f;
    if (a < -87.0f) a = -87.0f;
    return __fdiv_rn(1.0f, __fadd_rn(1.0f, exp_core(a)));
}

__device__ __forceinline__ float tanh_c(float x)
{
    float ax = fabsf(x);
    if (ax < 0.625f) {
        float z = __fmul_rn(x, x);
        float p = -5.70498872745e-3f;
        p = __fmaf_rn(p, z, 2.06390887954e-2f);
        p = __fmaf_rn(p, z, -5.37397155531e-2f);
        p = __fmaf_rn(p, z, 1.33314422036e-1f);
        p = __fmaf_rn(p, z, -3.33332819422e-1f);
        return __fmaf_rn(__fmul_rn(p, z), x, x);
    }
    if (ax > 10.0f) ax = 10.0f;
    float e = exp_core(__fadd_rn(ax, ax));
    float y = __fsub_rn(1.0f, __fmul_rn(2.0f, __fdiv_rn(1.0f, __fadd_rn(e, 1.0f))));
    return x < 0.0f ? -y : y;
}

// GRU cell tail (torch semantics, wavernn.py:71): r, z gates, candidate n, convex update
__device__ __forceinline__ float gru_update(float ar, float az, float ani, float anh, float h)
{
    float r = sigmoid_c(ar);
    float z = sigmoid_c(az);
    float n = tanh_c(__fmaf_rn(r, anh, ani));
    return __fmaf_rn(z, __fsub_rn(h, n), n);
}

}  // namespace fpc
