// Per-element arithmetic of the diagonal GDN / IGDN (layers.py:19-27), shared by the streaming kernels (gdn.cu) and the fused
// first-layer kernel (conv0_gdn.cu).  Forward uses only __f*_rn intrinsics so ptxas cannot contract mul+add into an FMA: a fused
// beta + gamma*x2 differs in the last bit and would flip round() on latents that sit on a half-integer.
#pragma once
#include "common.cuh"

namespace sic {
namespace gdnm {

constexpr float kOffset = 3.814697265625e-06f;  // 2^-18, layers.py:8

__device__ __forceinline__ void eff_params(const float *__restrict__ beta_param, const float *__restrict__ gamma_weight, int c,
                                           float &beta, float &gamma) {
    float b = __ldg(beta_param + c), w = __ldg(gamma_weight + c);
    beta = __fadd_rn(__fmul_rn(b, b), -kOffset);   // layers.py:20
    gamma = __fadd_rn(__fmul_rn(w, w), -kOffset);  // layers.py:21
}

// optional fused conv bias: PyTorch runs conv (no bias) -> add_(bias) -> GDN as three passes; the add is folded in here
// (rn(x + b), the same rounding) and its gradient (sum of dx per channel) comes out of the backward's reduction for free.
// -0.0f is the neutral element that keeps every bit of x, including the sign of zero.
__device__ __forceinline__ float load_bias(const float *__restrict__ bias, int c) { return bias != nullptr ? __ldg(bias + c) : -0.0f; }

template <bool INVERSE>
__device__ __forceinline__ float gdn1(float xin, float bias, float beta, float gamma) {
    float x = __fadd_rn(xin, bias);
    float x2 = __fmul_rn(x, x);
    float p = __fmul_rn(gamma, x2);
    float s = __fadd_rn(beta, p);
    float d = __fsqrt_rn(s);
    return INVERSE ? __fmul_rn(x, d) : __fdiv_rn(x, d);
}

// GDN : y = x/d    dx = g*beta/d^3            h = -1/2 g x / d^3
// IGDN: y = x*d    dx = g*(s + gamma x^2)/d   h = +1/2 g x / d          dbeta_c = sum h ; dgamma_c = sum h x^2
template <bool INVERSE>
__device__ __forceinline__ void gdn_bwd1(float x, float g, float beta, float gamma, float &dx, float &hb, float &hg) {
    float x2 = x * x;
    float gx2 = gamma * x2;
    float s = beta + gx2;
    float r = rsqrtf(s);
    if (INVERSE) {
        dx = g * (s + gx2) * r;
        hb = 0.5f * g * x * r;
    } else {
        float r3 = r * r * r;
        dx = g * beta * r3;
        hb = -0.5f * g * x * r3;
    }
    hg = hb * x2;
}

}  // namespace gdnm
}  // namespace sic
