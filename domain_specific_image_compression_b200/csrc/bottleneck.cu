// K1: fused quantise + likelihood + rate (forward) and its analytic backward.
//
// Replaces, per call, the eager chains of  model.py:27-35 (quantize),  distributions.py:20-31 (Student-t density),
// distributions.py:39-46 (Gaussian z prior) and the .sum() reductions of model.py:77  (paths relative to
// /root/reference/code/modelv2).
//
// Work decomposition: a ROW is one (b,c) plane of HW elements; rows are cut into SEGMENTS of `seg` elements and one
// warp owns one (row, segment) unit.  In the broadcast/channel layouts sigma/nu are row constants, so the lgamma /
// log prefactor is evaluated once per unit (redundantly across the 32 lanes: one issue slot either way) and the
// per-element work is  mul, mul, mul, log1p, fma.  Loads/stores are 128-bit, coalesced, L1-bypassing; four vectors
// are in flight per lane.  Each unit leaves one partial sum; the last CTA to retire (ticket counter) folds the partials
// per patch in a fixed order with a float64 accumulator => bit-reproducible rate, single launch.
//
// HBM-bound by design: 12 B/element (broadcast) against ~25 issue slots per element-lane.
#include "common.cuh"

namespace sic {
namespace {

constexpr int kWarpsPerCta = 8;
constexpr int kThreads = kWarpsPerCta * 32;
constexpr int kMinSeg = 128;
constexpr int kMaxSeg = 2048;
constexpr size_t kWsHeader = 256;  // bytes reserved for the ticket counter

enum { MODE_T_BCAST = 0, MODE_T_SPATIAL = 1, MODE_GAUSS = 2, MODE_T_ROWWISE_MASK = 0 };

struct Shape {
    int B, C, HW, seg, segs;  // segs = segments per row
    long units;               // rows * segs
};

inline Shape make_shape(int B, int C, int HW) {
    Shape s{B, C, HW, kMaxSeg, 0, 0};
    long rows = (long)B * C;
    const long want = (long)sm_count() * 32;  // ≥ 32 warps per SM before segments are allowed to grow
    while (s.seg > kMinSeg && rows * ((HW + s.seg - 1) / s.seg) < want) s.seg >>= 1;
    s.segs = (HW + s.seg - 1) / s.seg;
    s.units = rows * s.segs;
    return s;
}

// ---- Student-t prefactor ---------------------------------------------------------------------------------------
// D(a) = lgamma(a+1/2) - lgamma(a) evaluated WITHOUT cancellation: shift a by 4 (exact recurrence) and use the
// asymptotic series of the ratio at b = a+4 >= 5.  fp32 abs error 4e-7 (the reference's two fp32 lgamma calls: 1.5e-5).
__device__ __forceinline__ float lgamma_half_step(float a) {
    float b = a + 4.0f;
    float num = a * (a + 1.0f) * (a + 2.0f) * (a + 3.0f);
    float den = (a + 0.5f) * (a + 1.5f) * (a + 2.5f) * (a + 3.5f);
    float rb = 1.0f / b, rb2 = rb * rb;
    float series = rb * (-0.125f + rb2 * (5.2083333e-3f - rb2 * 1.5625e-3f));
    return logf(sqrtf(b) * (num / den)) + series;
}

struct TConst {
    float inv_sigma, inv_nu, A, Bc;  // nll = A + Bc * log1p((x*inv_sigma)^2 * inv_nu)
};
__device__ __forceinline__ TConst t_const(float sigma_raw, float nu_raw) {
    float s = clamp_keep_nan(sigma_raw, kSigmaMin, kSigmaMax);  // distributions.py:23
    float n = clamp_keep_nan(nu_raw, kNuMin, kNuMax);           // distributions.py:24
    // logC = lgamma((nu+1)/2) - lgamma(nu/2) - 0.5 log(nu pi) - log sigma     (distributions.py:27)
    float logC = lgamma_half_step(0.5f * n) - 0.5f * logf(n * 3.14159265358979f * s * s);
    TConst c;
    c.inv_sigma = 1.0f / s;
    c.inv_nu = 1.0f / n;
    c.A = -logC * kLog2e;
    c.Bc = 0.5f * (n + 1.0f) * kLog2e;  // distributions.py:29,31
    return c;
}
__device__ __forceinline__ float t_nll(float x, const TConst &c) {
    float q = x * c.inv_sigma;
    return fmaf(c.Bc, log1pf(q * q * c.inv_nu), c.A);
}

struct GConst {
    float A, Bc;  // nll = A + Bc * x^2
};
__device__ __forceinline__ GConst g_const(float log_sigma) {
    float s = clamp_keep_nan(expf(log_sigma), kSigmaMin, kSigmaMax);  // distributions.py:42-43
    float var = s * s;
    GConst c;
    c.A = 0.5f * logf(6.28318530717959f * var) * kLog2e;  // distributions.py:45-46
    c.Bc = 0.5f / var * kLog2e;
    return c;
}

__device__ __forceinline__ float quantize1(float y, int quant_mode, float noise) {
    if (quant_mode == SIC_QUANT_ROUND) return rintf(y);  // round-half-to-even, keeps -0.0 (model.py:33)
    if (quant_mode >= SIC_QUANT_NOISE_TENSOR) return y + noise;  // model.py:30-31
    return y;
}

template <int MODE>
__device__ __forceinline__ float elem_nll(float x, const TConst &tc, const GConst &gc, float sg, float nu) {
    if (MODE == MODE_GAUSS) return fmaf(gc.Bc, x * x, gc.A);
    if (MODE == MODE_T_SPATIAL) return t_nll(x, t_const(sg, nu));
    return t_nll(x, tc);
}

// fold per-unit partials into per-patch sums; executed by the last CTA only
__device__ void finalize_bits(const float *psum, long units_per_patch, int B, float *bits) {
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int b = warp; b < B; b += kWarpsPerCta) {
        const float *p = psum + (long)b * units_per_patch;
        double acc = 0.0;
        for (long i = lane; i < units_per_patch; i += 32) acc += (double)__ldcg(p + i);
        acc = warp_sum(acc);
        if (lane == 0) bits[b] = (float)acc;
    }
}

__device__ __forceinline__ bool retire_and_check_last(unsigned int *ticket) {
    __shared__ bool last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (last) __threadfence();
    return last;
}

// ------------------------------------------------------------------------------------------------------------------
template <int MODE, bool VEC>
__global__ void __launch_bounds__(kThreads) bottleneck_fwd_kernel(
    const float *__restrict__ y, const float *__restrict__ noise, uint64_t *__restrict__ philox,
    const float *__restrict__ mu, const float *__restrict__ sigma, const float *__restrict__ nu, Shape sh, int quant_mode,
    int mu_layout, float *__restrict__ y_tilde, float *__restrict__ nll, float *__restrict__ bits, float *__restrict__ psum,
    unsigned int *__restrict__ ticket) {
    const int lane = threadIdx.x & 31;
    const long unit = (long)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
    if (unit < sh.units) {
        const long row = unit / sh.segs;
        const int seg_i = (int)(unit - row * sh.segs);
        const int e0 = seg_i * sh.seg;
        const int e1 = min(e0 + sh.seg, sh.HW);
        const long base = row * (long)sh.HW;
        TConst tc{};
        GConst gc{};
        float mu_row = 0.0f;
        if (MODE == MODE_T_BCAST) tc = t_const(__ldg(sigma + row), __ldg(nu + row));
        if (MODE == MODE_GAUSS) gc = g_const(__ldg(sigma + row % sh.C));
        if (mu != nullptr && mu_layout != SIC_PARAM_SPATIAL) mu_row = __ldg(mu + (mu_layout == SIC_PARAM_CHANNEL ? row % sh.C : row));
        uint2 key = make_uint2(0, 0);
        uint2 off = make_uint2(0, 0);
        if (quant_mode == SIC_QUANT_NOISE_PHILOX) {
            uint64_t s = philox[0], o = philox[1];
            key = make_uint2((uint32_t)s, (uint32_t)(s >> 32));
            off = make_uint2((uint32_t)o, (uint32_t)(o >> 32));
        }
        float acc = 0.0f;
        if (VEC) {
            constexpr int U = 4;  // vectors in flight per lane
            const int v0 = e0 >> 2, v1 = e1 >> 2;
            const long vbase = base >> 2;
            for (int v = v0 + lane; v < v1; v += 32 * U) {
                float4 yy[U], nn[U], ss[U], uu[U], mm[U];
#pragma unroll
                for (int k = 0; k < U; ++k) {
                    int vv = v + 32 * k;
                    if (vv < v1) {
                        long gi = vbase + vv;
                        yy[k] = ldg_stream(reinterpret_cast<const float4 *>(y) + gi);
                        if (quant_mode == SIC_QUANT_NOISE_TENSOR) nn[k] = ldg_stream(reinterpret_cast<const float4 *>(noise) + gi);
                        if (MODE == MODE_T_SPATIAL) {
                            ss[k] = ldg_stream(reinterpret_cast<const float4 *>(sigma) + gi);
                            uu[k] = ldg_stream(reinterpret_cast<const float4 *>(nu) + gi);
                        }
                        if (mu != nullptr && mu_layout == SIC_PARAM_SPATIAL) mm[k] = ldg_stream(reinterpret_cast<const float4 *>(mu) + gi);
                    }
                }
#pragma unroll
                for (int k = 0; k < U; ++k) {
                    int vv = v + 32 * k;
                    if (vv < v1) {
                        long gi = vbase + vv;
                        if (quant_mode == SIC_QUANT_NOISE_PHILOX) {
                            uint4 r = philox4x32_10(make_uint4((uint32_t)gi, (uint32_t)((uint64_t)gi >> 32), off.x, off.y), key);
                            nn[k] = make_float4(u32_to_noise(r.x), u32_to_noise(r.y), u32_to_noise(r.z), u32_to_noise(r.w));
                        }
                        float4 q, l;
                        float4 m4 = (mu != nullptr && mu_layout == SIC_PARAM_SPATIAL) ? mm[k] : make_float4(mu_row, mu_row, mu_row, mu_row);
                        q.x = quantize1(yy[k].x, quant_mode, nn[k].x);
                        q.y = quantize1(yy[k].y, quant_mode, nn[k].y);
                        q.z = quantize1(yy[k].z, quant_mode, nn[k].z);
                        q.w = quantize1(yy[k].w, quant_mode, nn[k].w);
                        l.x = elem_nll<MODE>(q.x - m4.x, tc, gc, ss[k].x, uu[k].x);
                        l.y = elem_nll<MODE>(q.y - m4.y, tc, gc, ss[k].y, uu[k].y);
                        l.z = elem_nll<MODE>(q.z - m4.z, tc, gc, ss[k].z, uu[k].z);
                        l.w = elem_nll<MODE>(q.w - m4.w, tc, gc, ss[k].w, uu[k].w);
                        if (y_tilde != nullptr) stg_stream(reinterpret_cast<float4 *>(y_tilde) + gi, q);
                        if (nll != nullptr) stg_stream(reinterpret_cast<float4 *>(nll) + gi, l);
                        acc += (l.x + l.y) + (l.z + l.w);
                    }
                }
            }
        } else {
            for (int e = e0 + lane; e < e1; e += 32) {
                long gi = base + e;
                float n1 = 0.0f;
                if (quant_mode == SIC_QUANT_NOISE_TENSOR) n1 = noise[gi];
                if (quant_mode == SIC_QUANT_NOISE_PHILOX) {
                    long v = gi >> 2;
                    uint4 r = philox4x32_10(make_uint4((uint32_t)v, (uint32_t)((uint64_t)v >> 32), off.x, off.y), key);
                    int j = (int)(gi & 3);
                    n1 = u32_to_noise(j == 0 ? r.x : j == 1 ? r.y : j == 2 ? r.z : r.w);
                }
                float q = quantize1(y[gi], quant_mode, n1);
                float m1 = (mu != nullptr && mu_layout == SIC_PARAM_SPATIAL) ? mu[gi] : mu_row;
                float sg = 0.f, nv = 0.f;
                if (MODE == MODE_T_SPATIAL) { sg = sigma[gi]; nv = nu[gi]; }
                float l = elem_nll<MODE>(q - m1, tc, gc, sg, nv);
                if (y_tilde != nullptr) y_tilde[gi] = q;
                if (nll != nullptr) nll[gi] = l;
                acc += l;
            }
        }
        acc = warp_sum(acc);
        if (lane == 0) psum[unit] = acc;
    }
    if (retire_and_check_last(ticket)) {
        finalize_bits(psum, (long)sh.C * sh.segs, sh.B, bits);
        if (threadIdx.x == 0) {
            *ticket = 0;
            if (quant_mode == SIC_QUANT_NOISE_PHILOX) philox[1] += 1;  // every other CTA has retired: safe to advance
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// backward.  Row constants in float64 where the formula cancels (digamma difference minus 1/(2 nu) ~ 1/(4 nu^2)).
struct TBwdConst {
    float sigma2nu, np1, inv_sigma, inv_nu, mask_s, mask_n, Knu;
};
__device__ __forceinline__ TBwdConst t_bwd_const(float sigma_raw, float nu_raw) {
    float s = clamp_keep_nan(sigma_raw, kSigmaMin, kSigmaMax);
    float n = clamp_keep_nan(nu_raw, kNuMin, kNuMax);
    TBwdConst c;
    c.sigma2nu = s * s * n;
    c.np1 = n + 1.0f;
    c.inv_sigma = 1.0f / s;
    c.inv_nu = 1.0f / n;
    c.mask_s = (sigma_raw >= kSigmaMin && sigma_raw <= kSigmaMax) ? 1.0f : 0.0f;  // torch.clamp passes grad on the closed interval
    c.mask_n = (nu_raw >= kNuMin && nu_raw <= kNuMax) ? 1.0f : 0.0f;
    // Knu = 0.5*(psi((nu+1)/2) - psi(nu/2)) - 1/(2 nu)
    double a = 0.5 * (double)n, b = a + 4.0, rb = 1.0 / b, rb2 = rb * rb;
    double E = rb * (0.5 + rb * (0.125 + rb2 * (-0.015625 + rb2 * 0.0078125)));
    E += 1.0 / a - 1.0 / (a + 0.5);
    E += 1.0 / (a + 1.0) - 1.0 / (a + 1.5);
    E += 1.0 / (a + 2.0) - 1.0 / (a + 2.5);
    E += 1.0 / (a + 3.0) - 1.0 / (a + 3.5);
    c.Knu = (float)(0.5 * E - 0.25 / a);
    return c;
}

// per element: returns dnll/dx, and the sigma / nu integrands (without upstream gradient)
__device__ __forceinline__ void t_bwd_elem(float x, const TBwdConst &c, float &dx, float &ds, float &dn) {
    float q = x * x;
    float den = c.sigma2nu + q;
    float r = c.np1 * __fdividef(1.0f, den);  // (nu+1)/(nu sigma^2 + x^2)
    dx = kLog2e * r * x;
    float rq = r * q;
    ds = kLog2e * c.inv_sigma * (1.0f - rq);
    float u = q * __fdividef(1.0f, c.sigma2nu);
    dn = -kLog2e * (c.Knu - 0.5f * log1pf(u) + 0.5f * rq * c.inv_nu);
}

template <int MODE, bool VEC>
__global__ void __launch_bounds__(kThreads) bottleneck_bwd_kernel(
    const float *__restrict__ yt, const float *__restrict__ mu, const float *__restrict__ sigma, const float *__restrict__ nu,
    const float *__restrict__ g_nll, const float *__restrict__ g_bits, const float *__restrict__ g_yt, Shape sh, int quant_mode,
    int mu_layout, float *__restrict__ dy, float *__restrict__ dmu, float *__restrict__ dsigma, float *__restrict__ dnu,
    float *__restrict__ pa, float *__restrict__ pb, float *__restrict__ pc, unsigned int *__restrict__ ticket) {
    const int lane = threadIdx.x & 31;
    const long unit = (long)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
    const bool pass_dy = quant_mode != SIC_QUANT_ROUND;
    if (unit < sh.units) {
        const long row = unit / sh.segs;
        const int seg_i = (int)(unit - row * sh.segs);
        const int e0 = seg_i * sh.seg, e1 = min(e0 + sh.seg, sh.HW);
        const long base = row * (long)sh.HW;
        const int b = (int)(row / sh.C);
        TBwdConst tc{};
        float g_inv_var = 0.f, g_mask = 0.f;
        if (MODE == MODE_T_BCAST) tc = t_bwd_const(__ldg(sigma + row), __ldg(nu + row));
        if (MODE == MODE_GAUSS) {
            float sr = expf(__ldg(sigma + row % sh.C));
            float s = clamp_keep_nan(sr, kSigmaMin, kSigmaMax);
            g_inv_var = 1.0f / (s * s);
            g_mask = (sr >= kSigmaMin && sr <= kSigmaMax) ? 1.0f : 0.0f;
        }
        float mu_row = 0.0f;
        if (mu != nullptr && mu_layout != SIC_PARAM_SPATIAL) mu_row = __ldg(mu + (mu_layout == SIC_PARAM_CHANNEL ? row % sh.C : row));
        const float gb = g_bits != nullptr ? __ldg(g_bits + b) : 0.0f;
        float acc_s = 0.f, acc_n = 0.f, acc_m = 0.f;

        auto one = [&](long gi, float ytv, float gl, float gy, float sg, float nv, float m1, float &o_dy, float &o_ds, float &o_dn, float &o_dm) {
            float gt = gl + gb;
            float x = ytv - m1;
            float dxe, dse = 0.f, dne = 0.f;
            if (MODE == MODE_GAUSS) {
                dxe = kLog2e * x * g_inv_var;
                dse = kLog2e * (1.0f - x * x * g_inv_var) * g_mask;  // d/dlog_sigma
            } else if (MODE == MODE_T_SPATIAL) {
                TBwdConst c = t_bwd_const(sg, nv);
                t_bwd_elem(x, c, dxe, dse, dne);
                dse *= c.mask_s;
                dne *= c.mask_n;
            } else {
                t_bwd_elem(x, tc, dxe, dse, dne);
            }
            float gx = gt * dxe;
            o_dy = pass_dy ? gy + gx : 0.0f;
            o_ds = gt * dse;
            o_dn = gt * dne;
            o_dm = -gx;
        };

        if (VEC) {
            constexpr int U = 2;
            const int v0 = e0 >> 2, v1 = e1 >> 2;
            const long vbase = base >> 2;
            for (int v = v0 + lane; v < v1; v += 32 * U) {
                float4 a[U], gl[U], gy[U], ss[U], uu[U], mm[U];
#pragma unroll
                for (int k = 0; k < U; ++k) {
                    int vv = v + 32 * k;
                    if (vv < v1) {
                        long gi = vbase + vv;
                        a[k] = ldg_stream(reinterpret_cast<const float4 *>(yt) + gi);
                        gl[k] = g_nll ? ldg_stream(reinterpret_cast<const float4 *>(g_nll) + gi) : make_float4(0, 0, 0, 0);
                        gy[k] = (g_yt && pass_dy) ? ldg_stream(reinterpret_cast<const float4 *>(g_yt) + gi) : make_float4(0, 0, 0, 0);
                        if (MODE == MODE_T_SPATIAL) {
                            ss[k] = ldg_stream(reinterpret_cast<const float4 *>(sigma) + gi);
                            uu[k] = ldg_stream(reinterpret_cast<const float4 *>(nu) + gi);
                        }
                        if (mu != nullptr && mu_layout == SIC_PARAM_SPATIAL) mm[k] = ldg_stream(reinterpret_cast<const float4 *>(mu) + gi);
                    }
                }
#pragma unroll
                for (int k = 0; k < U; ++k) {
                    int vv = v + 32 * k;
                    if (vv < v1) {
                        long gi = vbase + vv;
                        float4 m4 = (mu != nullptr && mu_layout == SIC_PARAM_SPATIAL) ? mm[k] : make_float4(mu_row, mu_row, mu_row, mu_row);
                        float4 o, s4, n4, d4;
                        one(gi, a[k].x, gl[k].x, gy[k].x, ss[k].x, uu[k].x, m4.x, o.x, s4.x, n4.x, d4.x);
                        one(gi, a[k].y, gl[k].y, gy[k].y, ss[k].y, uu[k].y, m4.y, o.y, s4.y, n4.y, d4.y);
                        one(gi, a[k].z, gl[k].z, gy[k].z, ss[k].z, uu[k].z, m4.z, o.z, s4.z, n4.z, d4.z);
                        one(gi, a[k].w, gl[k].w, gy[k].w, ss[k].w, uu[k].w, m4.w, o.w, s4.w, n4.w, d4.w);
                        if (dy) stg_stream(reinterpret_cast<float4 *>(dy) + gi, o);
                        if (MODE == MODE_T_SPATIAL) {
                            if (dsigma) stg_stream(reinterpret_cast<float4 *>(dsigma) + gi, s4);
                            if (dnu) stg_stream(reinterpret_cast<float4 *>(dnu) + gi, n4);
                        } else {
                            acc_s += (s4.x + s4.y) + (s4.z + s4.w);
                            acc_n += (n4.x + n4.y) + (n4.z + n4.w);
                        }
                        if (dmu != nullptr && mu_layout == SIC_PARAM_SPATIAL) stg_stream(reinterpret_cast<float4 *>(dmu) + gi, d4);
                        else acc_m += (d4.x + d4.y) + (d4.z + d4.w);
                    }
                }
            }
        } else {
            for (int e = e0 + lane; e < e1; e += 32) {
                long gi = base + e;
                float sg = 0.f, nv = 0.f;
                if (MODE == MODE_T_SPATIAL) { sg = sigma[gi]; nv = nu[gi]; }
                float m1 = (mu != nullptr && mu_layout == SIC_PARAM_SPATIAL) ? mu[gi] : mu_row;
                float o, s1, n1, d1;
                one(gi, yt[gi], g_nll ? g_nll[gi] : 0.f, (g_yt && pass_dy) ? g_yt[gi] : 0.f, sg, nv, m1, o, s1, n1, d1);
                if (dy) dy[gi] = o;
                if (MODE == MODE_T_SPATIAL) {
                    if (dsigma) dsigma[gi] = s1;
                    if (dnu) dnu[gi] = n1;
                } else {
                    acc_s += s1;
                    acc_n += n1;
                }
                if (dmu != nullptr && mu_layout == SIC_PARAM_SPATIAL) dmu[gi] = d1;
                else acc_m += d1;
            }
        }
        acc_s = warp_sum(acc_s);
        acc_n = warp_sum(acc_n);
        acc_m = warp_sum(acc_m);
        if (lane == 0) {
            if (MODE == MODE_T_BCAST) { acc_s *= tc.mask_s; acc_n *= tc.mask_n; }
            pa[unit] = acc_s;
            pb[unit] = acc_n;
            pc[unit] = acc_m;
        }
    }
    if (retire_and_check_last(ticket)) {
        // fold partials: broadcast -> per row over segments; channel -> per channel over (b, segment), fixed order
        const long rows = (long)sh.B * sh.C;
        if (MODE != MODE_T_SPATIAL) {
            if (MODE == MODE_T_BCAST) {
                for (long r = threadIdx.x; r < rows; r += kThreads) {
                    double s = 0, n = 0;
                    for (int k = 0; k < sh.segs; ++k) { s += (double)__ldcg(pa + r * sh.segs + k); n += (double)__ldcg(pb + r * sh.segs + k); }
                    if (dsigma) dsigma[r] = (float)s;
                    if (dnu) dnu[r] = (float)n;
                }
            } else {
                for (int c = threadIdx.x; c < sh.C; c += kThreads) {
                    double s = 0;
                    for (int bb = 0; bb < sh.B; ++bb)
                        for (int k = 0; k < sh.segs; ++k) s += (double)__ldcg(pa + ((long)bb * sh.C + c) * sh.segs + k);
                    if (dsigma) dsigma[c] = (float)s;
                }
            }
        }
        if (dmu != nullptr && mu_layout == SIC_PARAM_BROADCAST) {
            for (long r = threadIdx.x; r < rows; r += kThreads) {
                double m = 0;
                for (int k = 0; k < sh.segs; ++k) m += (double)__ldcg(pc + r * sh.segs + k);
                dmu[r] = (float)m;
            }
        } else if (dmu != nullptr && mu_layout == SIC_PARAM_CHANNEL) {
            for (int c = threadIdx.x; c < sh.C; c += kThreads) {
                double m = 0;
                for (int bb = 0; bb < sh.B; ++bb)
                    for (int k = 0; k < sh.segs; ++k) m += (double)__ldcg(pc + ((long)bb * sh.C + c) * sh.segs + k);
                dmu[c] = (float)m;
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) *ticket = 0;
    }
}

inline bool aligned16(const void *p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

int resolve_mode(int lik_mode, int param_layout, int *mode) {
    if (lik_mode == SIC_LIK_STUDENTT_DENSITY && param_layout == SIC_PARAM_BROADCAST) { *mode = MODE_T_BCAST; return 0; }
    if (lik_mode == SIC_LIK_STUDENTT_DENSITY && param_layout == SIC_PARAM_SPATIAL) { *mode = MODE_T_SPATIAL; return 0; }
    if (lik_mode == SIC_LIK_GAUSSIAN && param_layout == SIC_PARAM_CHANNEL) { *mode = MODE_GAUSS; return 0; }
    return SIC_E_UNSUPPORTED;
}

}  // namespace

// cdf_diff mode lives in bottleneck_cdf.cu
int bottleneck_cdfdiff_fwd(const float *y, const float *noise, uint64_t *philox, const float *mu, const float *sigma,
                           const float *nu, int B, int C, int HW, int quant_mode, int param_layout, float *y_tilde, float *nll,
                           float *bits, void *workspace, size_t workspace_bytes, cudaStream_t st);
int bottleneck_cdfdiff_bwd(const float *y_tilde, const float *mu, const float *sigma, const float *nu, const float *g_nll,
                           const float *g_bits, const float *g_ytilde, int B, int C, int HW, int quant_mode, int param_layout,
                           float *dy, float *dmu, float *dsigma, float *dnu, void *workspace, size_t workspace_bytes,
                           cudaStream_t st);

}  // namespace sic

using namespace sic;

extern "C" size_t sic_bottleneck_workspace_bytes(int B, int C, int HW) {
    if (B <= 0 || C <= 0 || HW <= 0) return kWsHeader;
    long units = (long)B * C * ((HW + kMinSeg - 1) / kMinSeg);
    return kWsHeader + (size_t)units * 3 * sizeof(float);
}

extern "C" int sic_bottleneck_fwd(const float *y, const float *noise, uint64_t *philox, const float *mu,
                                  const float *sigma, const float *nu, int B, int C, int HW, int quant_mode, int lik_mode,
                                  int param_layout, float *y_tilde, float *nll, float *bits, void *workspace,
                                  size_t workspace_bytes, void *stream) {
    SIC_CHECK_ARG(B > 0 && C > 0 && HW > 0, "sic_bottleneck_fwd: empty shape B=%d C=%d HW=%d", B, C, HW);
    SIC_CHECK_ARG(y && bits && sigma && workspace, "sic_bottleneck_fwd: null required pointer");
    SIC_CHECK_ARG(y_tilde || quant_mode == SIC_QUANT_NONE, "sic_bottleneck_fwd: y_tilde may be NULL only with SIC_QUANT_NONE");
    SIC_CHECK_ARG(quant_mode >= SIC_QUANT_NONE && quant_mode <= SIC_QUANT_NOISE_PHILOX, "sic_bottleneck_fwd: bad quant_mode %d", quant_mode);
    SIC_CHECK_ARG(quant_mode != SIC_QUANT_NOISE_TENSOR || noise, "sic_bottleneck_fwd: noise tensor required");
    SIC_CHECK_ARG(quant_mode != SIC_QUANT_NOISE_PHILOX || philox, "sic_bottleneck_fwd: philox state required");
    SIC_CHECK_ARG((long)B * C * HW < (1L << 40), "sic_bottleneck_fwd: tensor too large");
    if (workspace_bytes < sic_bottleneck_workspace_bytes(B, C, HW)) {
        set_error("sic_bottleneck_fwd: workspace %zu < %zu bytes", workspace_bytes, sic_bottleneck_workspace_bytes(B, C, HW));
        return SIC_E_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (lik_mode == SIC_LIK_STUDENTT_CDFDIFF)
        return bottleneck_cdfdiff_fwd(y, noise, philox, mu, sigma, nu, B, C, HW, quant_mode, param_layout, y_tilde, nll, bits,
                                      workspace, workspace_bytes, st);
    int mode;
    if (resolve_mode(lik_mode, param_layout, &mode) != 0) {
        set_error("sic_bottleneck_fwd: unsupported lik_mode %d with param_layout %d", lik_mode, param_layout);
        return SIC_E_UNSUPPORTED;
    }
    SIC_CHECK_ARG(mode == MODE_GAUSS || nu, "sic_bottleneck_fwd: nu required for Student-t");
    int mu_layout = mode == MODE_GAUSS ? SIC_PARAM_CHANNEL : param_layout;
    Shape sh = make_shape(B, C, HW);
    bool vec = (HW % 4 == 0) && aligned16(y) && aligned16(noise) && aligned16(y_tilde) && aligned16(nll) &&
               (mode != MODE_T_SPATIAL || (aligned16(sigma) && aligned16(nu))) && (param_layout != SIC_PARAM_SPATIAL || aligned16(mu));
    unsigned int *ticket = reinterpret_cast<unsigned int *>(workspace);
    float *psum = reinterpret_cast<float *>(static_cast<char *>(workspace) + kWsHeader);
    dim3 grid((unsigned)((sh.units + kWarpsPerCta - 1) / kWarpsPerCta)), block(kThreads);
#define LAUNCH(M, V)                                                                                                       \
    bottleneck_fwd_kernel<M, V><<<grid, block, 0, st>>>(y, noise, philox, mu, sigma, nu, sh, quant_mode, mu_layout, y_tilde, \
                                                        nll, bits, psum, ticket)
    if (mode == MODE_T_BCAST) { if (vec) LAUNCH(MODE_T_BCAST, true); else LAUNCH(MODE_T_BCAST, false); }
    else if (mode == MODE_T_SPATIAL) { if (vec) LAUNCH(MODE_T_SPATIAL, true); else LAUNCH(MODE_T_SPATIAL, false); }
    else { if (vec) LAUNCH(MODE_GAUSS, true); else LAUNCH(MODE_GAUSS, false); }
#undef LAUNCH
    SIC_CHECK_LAUNCH("sic_bottleneck_fwd");
    return 0;
}

extern "C" int sic_bottleneck_bwd(const float *y_tilde, const float *mu, const float *sigma, const float *nu,
                                  const float *g_nll, const float *g_bits, const float *g_ytilde, int B, int C, int HW,
                                  int quant_mode, int lik_mode, int param_layout, float *dy, float *dmu, float *dsigma,
                                  float *dnu, void *workspace, size_t workspace_bytes, void *stream) {
    SIC_CHECK_ARG(B > 0 && C > 0 && HW > 0, "sic_bottleneck_bwd: empty shape B=%d C=%d HW=%d", B, C, HW);
    SIC_CHECK_ARG(y_tilde && sigma && workspace, "sic_bottleneck_bwd: null required pointer");
    SIC_CHECK_ARG(dmu == nullptr || mu != nullptr, "sic_bottleneck_bwd: dmu requested without mu");
    if (workspace_bytes < sic_bottleneck_workspace_bytes(B, C, HW)) {
        set_error("sic_bottleneck_bwd: workspace %zu < %zu bytes", workspace_bytes, sic_bottleneck_workspace_bytes(B, C, HW));
        return SIC_E_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (lik_mode == SIC_LIK_STUDENTT_CDFDIFF)
        return bottleneck_cdfdiff_bwd(y_tilde, mu, sigma, nu, g_nll, g_bits, g_ytilde, B, C, HW, quant_mode, param_layout, dy,
                                      dmu, dsigma, dnu, workspace, workspace_bytes, st);
    int mode;
    if (resolve_mode(lik_mode, param_layout, &mode) != 0) {
        set_error("sic_bottleneck_bwd: unsupported lik_mode %d with param_layout %d", lik_mode, param_layout);
        return SIC_E_UNSUPPORTED;
    }
    SIC_CHECK_ARG(mode == MODE_GAUSS || nu, "sic_bottleneck_bwd: nu required for Student-t");
    int mu_layout = mode == MODE_GAUSS ? SIC_PARAM_CHANNEL : param_layout;
    Shape sh = make_shape(B, C, HW);
    bool vec = (HW % 4 == 0) && aligned16(y_tilde) && aligned16(g_nll) && aligned16(g_ytilde) && aligned16(dy) &&
               (mode != MODE_T_SPATIAL || (aligned16(sigma) && aligned16(nu) && aligned16(dsigma) && aligned16(dnu))) &&
               (param_layout != SIC_PARAM_SPATIAL || (aligned16(mu) && aligned16(dmu)));
    unsigned int *ticket = reinterpret_cast<unsigned int *>(workspace);
    float *pa = reinterpret_cast<float *>(static_cast<char *>(workspace) + kWsHeader);
    float *pb = pa + sh.units, *pc = pb + sh.units;
    dim3 grid((unsigned)((sh.units + kWarpsPerCta - 1) / kWarpsPerCta)), block(kThreads);
#define LAUNCH(M, V)                                                                                                          \
    bottleneck_bwd_kernel<M, V><<<grid, block, 0, st>>>(y_tilde, mu, sigma, nu, g_nll, g_bits, g_ytilde, sh, quant_mode, mu_layout, \
                                                        dy, dmu, dsigma, dnu, pa, pb, pc, ticket)
    if (mode == MODE_T_BCAST) { if (vec) LAUNCH(MODE_T_BCAST, true); else LAUNCH(MODE_T_BCAST, false); }
    else if (mode == MODE_T_SPATIAL) { if (vec) LAUNCH(MODE_T_SPATIAL, true); else LAUNCH(MODE_T_SPATIAL, false); }
    else { if (vec) LAUNCH(MODE_GAUSS, true); else LAUNCH(MODE_GAUSS, false); }
#undef LAUNCH
    SIC_CHECK_LAUNCH("sic_bottleneck_bwd");
    return 0;
}
