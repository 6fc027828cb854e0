// Kernel templates of K1 (shared by bottleneck.cu: density / Gaussian modes, and bottleneck_cdf.cu: cdf_diff modes).
// See bottleneck.cu for the design notes.
#pragma once
#include "common.cuh"

namespace sic {
namespace {

constexpr int kWarpsPerCta = 8;
constexpr int kThreads = kWarpsPerCta * 32;
constexpr int kMinSeg = 128;
constexpr int kMaxSeg = 2048;
constexpr size_t kWsHeader = 256;  // bytes reserved for the ticket counter

enum { MODE_T_BCAST = 0, MODE_T_SPATIAL = 1, MODE_GAUSS = 2, MODE_CDF_BCAST = 3, MODE_CDF_SPATIAL = 4 };
__host__ __device__ constexpr bool is_bcast(int m) { return m == MODE_T_BCAST || m == MODE_CDF_BCAST; }
__host__ __device__ constexpr bool is_spatial(int m) { return m == MODE_T_SPATIAL || m == MODE_CDF_SPATIAL; }
__host__ __device__ constexpr bool is_cdf(int m) { return m == MODE_CDF_BCAST || m == MODE_CDF_SPATIAL; }

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
    float rb = __fdividef(1.0f, b), rb2 = rb * rb;
    float series = rb * (-0.125f + rb2 * (5.2083333e-3f - rb2 * 1.5625e-3f));
    // ln(sqrt(b) num/den) through MUFU.LG2 / MUFU.RCP: abs error ~3e-7, still 50x tighter than the reference's two fp32 lgammas;
    // this runs once per (row, segment) in the broadcast layout but once per ELEMENT in the spatial layout
    return 0.69314718055994531f * (0.5f * __log2f(b) + __log2f(num * __fdividef(1.0f, den))) + series;
}

constexpr float kPanel = 1.5f;   // cdf_diff quadrature: panel width in units of the local density scale
constexpr int kMaxPanels = 48;

struct TConst {
    float inv_sigma, inv_nu, A, B2;  // nll = A + B2 * log2(1 + (x*inv_sigma)^2 * inv_nu),  B2 = (nu+1)/2
    float nu, lc2, kstep;            // cdf_diff: log2 of the unit-scale normaliser, quadrature panel scale
    float rn, inv_rn, inv_ds;        // cdf_diff forward: sqrt(nu), 1/sqrt(nu), panels per unit of asinh(t/sqrt(nu))
};
__device__ __forceinline__ TConst t_const(float sigma_raw, float nu_raw) {
    float s = clamp_keep_nan(sigma_raw, kSigmaMin, kSigmaMax);  // distributions.py:23
    float n = clamp_keep_nan(nu_raw, kNuMin, kNuMax);           // distributions.py:24
    // logC = lgamma((nu+1)/2) - lgamma(nu/2) - 0.5 log(nu pi) - log sigma     (distributions.py:27)
    float logC = lgamma_half_step(0.5f * n) - 0.34657359027997264f * __log2f(n * 3.14159265358979f * s * s);
    TConst c;
    c.inv_sigma = __fdividef(1.0f, s);
    c.inv_nu = __fdividef(1.0f, n);
    c.A = -logC * kLog2e;
    c.B2 = 0.5f * (n + 1.0f);  // distributions.py:29,31: (nu+1)/2 * log1p(.) * LOG2E == (nu+1)/2 * log2(1+.)
    c.nu = n;
    c.lc2 = -c.A + __log2f(s);         // log2 Gamma((nu+1)/2) / (sqrt(nu pi) Gamma(nu/2))
    c.kstep = kPanel * rsqrtf(n + 1.0f);
    c.inv_rn = rsqrtf(n);
    c.rn = n * c.inv_rn;
    c.inv_ds = (n + 1.0f) * rsqrtf(n + 1.0f) * (1.0f / kPanel);
    return c;
}
// One MUFU.LG2 per element.  Error budget in bits: B2 * (2^-24 rounding of 1+u  +  2^-22 MUFU) * log2e-ish <= 1.7e-5 at the
// extreme nu = 100, against a tolerance of 1e-4 + 1e-5|nll| (the reference's own fp32 chain is off by up to 7e-5).
__device__ __forceinline__ float t_nll(float x, const TConst &c) {
    float q = x * c.inv_sigma;
    return fmaf(c.B2, __log2f(fmaf(q * q, c.inv_nu, 1.0f)), c.A);
}


// ---- discretised likelihood (north_star: P = T_nu(y~+1/2) - T_nu(y~-1/2)) ---------------------------------------------
// torch has no Student-t CDF (the reference script calls one that does not exist: SURVEY D5).  A CDF difference cancels
// catastrophically in fp32 both in the tails and for sigma >> 1, so the bin mass is integrated directly:
//     P = int_lo^hi f(t) dt,  f(t) = Cn (1 + t^2/nu)^-(nu+1)/2,  lo/hi = (x -+ 1/2)/sigma
// by 8-point Gauss-Legendre on panels whose width follows the local scale of f, marching outward from the end nearest
// zero:  step = 1.5 sqrt((nu + t^2)/(nu + 1))  (calibrated against scipy in float64 over nu in [2,100], sigma in
// [1e-3,1e3], |x| <= 50 sigma: worst relative error 1.8e-7; one panel whenever sigma >= 0.82).  Values are scaled by
// f(nearest edge) so nothing underflows; the gradients are integrals / edge values of the same scaled density.
// Cost per panel: 8 x (LG2 + EX2) + ~40 FMA  =>  FP32/SFU-bound, not HBM-bound (ncu decides; see profiles/).
__device__ __constant__ float kGLx[4] = {0.1834346424956498f, 0.5255324099163290f, 0.7966664774136267f, 0.9602898564975363f};
__device__ __constant__ float kGLw[4] = {0.3626837833783620f, 0.3137066458778873f, 0.2223810344533745f, 0.1012285362903763f};

// One MUFU instruction each.  __log2f / exp2f wrap the same MUFU.LG2 / MUFU.EX2 in denormal handling (FSETP + FMUL before, FMUL /
// FADD after: the r02c SASS count has 83 M FSETP next to 88 M MUFU); the quadrature's log argument is >= 1 and a density that
// underflows to a denormal relative to the bin's peak contributes nothing, so the flush-to-zero forms are exact enough here.
__device__ __forceinline__ float lg2_ftz(float x) {
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float ex2_ftz(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// scaled density: 2^(-B2 log2(1+t^2/nu) - eref)
__device__ __forceinline__ float t_fs(float t, float inv_nu, float B2, float eref, float &l2, float &w) {
    w = fmaf(t * t, inv_nu, 1.0f);
    l2 = lg2_ftz(w);
    return ex2_ftz(fmaf(-B2, l2, -eref));
}

// integral over [a,b], 0 <= a <= b, of the scaled density (S) and, if WANT_G, of density * dlog f/dnu (Q)
template <bool WANT_G>
__device__ __forceinline__ void t_integrate(float a, float b, const TConst &c, float eref, float gK, float gC, float &S, float &Q) {
    float t = a;
    for (int p = 0; p < kMaxPanels && t < b; ++p) {
        float step = c.kstep * sqrtf(fmaf(t, t, c.nu));
        float t1 = t + step;
        if (b - t1 < 0.25f * step || p == kMaxPanels - 1) t1 = b;
        float m = 0.5f * (t + t1), h = 0.5f * (t1 - t);
        float acc = 0.f, accg = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
#pragma unroll
            for (int sgn = -1; sgn <= 1; sgn += 2) {
                float tt = fmaf(h, sgn * kGLx[i], m);
                float l2, w;
                float fs = t_fs(tt, c.inv_nu, c.B2, eref, l2, w);
                acc = fmaf(kGLw[i], fs, acc);
                if (WANT_G) {
                    float g = gK - 0.34657359027997264f * l2 + gC * (1.0f - __fdividef(1.0f, w));
                    accg = fmaf(kGLw[i] * fs, g, accg);
                }
            }
        }
        S = fmaf(h, acc, S);
        if (WANT_G) Q = fmaf(h, accg, Q);
        t = t1;
    }
}

// bin mass in scaled form: P = 2^(lc2 + eref) * S
template <bool WANT_G>
__device__ __forceinline__ void t_bin(float x, const TConst &c, float gK, float gC, float &S, float &Q, float &eref, float &lo, float &hi) {
    lo = (x - 0.5f) * c.inv_sigma;
    hi = (x + 0.5f) * c.inv_sigma;
    S = 0.f;
    Q = 0.f;
    if (lo >= 0.f || hi <= 0.f) {
        float a = lo >= 0.f ? lo : -hi, b = lo >= 0.f ? hi : -lo;
        eref = -c.B2 * __log2f(fmaf(a * a, c.inv_nu, 1.0f));
        t_integrate<WANT_G>(a, b, c, eref, gK, gC, S, Q);
    } else {
        eref = 0.f;
        t_integrate<WANT_G>(0.f, hi, c, 0.f, gK, gC, S, Q);
        t_integrate<WANT_G>(0.f, -lo, c, 0.f, gK, gC, S, Q);
    }
}

__device__ __forceinline__ float t_cdf_nll(float x, const TConst &c) {
    float S, Q, eref, lo, hi;
    t_bin<false>(x, c, 0.f, 0.f, S, Q, eref, lo, hi);
    return -(c.lc2 + eref + __log2f(S));
}


// ---- warp-uniform forward -------------------------------------------------------------------------------------------------
// History (profiles/): the per-lane marching loop above kept 12 of 32 lanes busy (panel counts differ inside a warp: elements
// near zero with a small sigma need 4+ panels, their neighbours one).  Round 1's answer pooled the panels of a warp's 32 elements
// and dealt them out 32 at a time: all lanes busy, but ncu (profiles/r02b_ncu_cdfdiff_pooled.txt) showed what the dealing costs -
// 1824 M warp instructions for 84 M elements of which only 119 M are MUFU; ISETP/BRA/SHFL/BSYNC/SEL are another 500 M; 121
// registers, 16 resident warps, MUFU pipe at 18 % of its peak rate: bound by instruction issue and latency, not by the SFU.
// This version removes the control flow instead of balancing it:
//   * the bin [lo, hi] is ONE interval in s = asinh(t / sqrt(nu)) (odd, monotonic; dt/ds = sqrt(nu + t^2), so equal steps in s are
//     the marching rule's local-scale panels, and a bin that straddles zero needs no splitting);
//   * every lane cuts ITS OWN interval into n equal s-steps, n = the largest count any lane of the warp (and any of the 4
//     elements a lane holds) asks for: the panel loop has a warp-uniform trip count, no divergence, no shuffles, no shared
//     memory.  Lanes that needed fewer panels get narrower ones (more accurate, never less).  In the broadcast layout all lanes
//     of a warp share sigma and nu, so n is set by the element nearest zero and the excess is small;
//   * 4 elements x 8 nodes per trip are 32 independent LG2/EX2 chains per lane: the latency the 16-warp occupancy could not hide.
// Same calibration as before (float64 against scipy: <= 2.8e-7 relative at an s-step of 1.5/sqrt(nu+1), kMaxPanels cap).
__device__ __forceinline__ float asinh_signed(float u) {
    const float a = fabsf(u);
    return copysignf(__logf(a + sqrtf(fmaf(a, a, 1.0f))), u);
}

template <int E>
__device__ __forceinline__ void uniform_cdf_nll(const float (&x)[E], bool active, const TConst &c, float (&out)[E]) {
    float s0[E], ds[E], t0[E], hi[E], eref[E], S[E];
    int n = 0;
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const float lo = (x[e] - 0.5f) * c.inv_sigma;
        hi[e] = (x[e] + 0.5f) * c.inv_sigma;
        const float near = fminf(fmaxf(0.0f, lo), hi[e]);              // where the density peaks inside the bin: the scaling reference
        eref[e] = -c.B2 * __log2f(fmaf(near * near, c.inv_nu, 1.0f));
        s0[e] = asinh_signed(lo * c.inv_rn);
        ds[e] = asinh_signed(hi[e] * c.inv_rn) - s0[e];
        n = max(n, (int)fminf(fmaxf(ceilf(ds[e] * c.inv_ds), 1.0f), (float)kMaxPanels));
        t0[e] = lo;
        S[e] = 0.0f;
    }
    n = __reduce_max_sync(0xffffffffu, active ? n : 0);
    const float inv_n = __fdividef(1.0f, (float)max(n, 1));
#pragma unroll
    for (int e = 0; e < E; ++e) ds[e] *= inv_n;
    for (int k = 1; k <= n; ++k) {                                     // warp-uniform trip count
#pragma unroll
        for (int e = 0; e < E; ++e) {
            float t1 = hi[e];                                          // exact ends; closed form inside
            if (k < n) {
                const float ex = __expf(fmaf((float)k, ds[e], s0[e]));
                t1 = c.rn * 0.5f * (ex - __fdividef(1.0f, ex));
            }
            const float m = 0.5f * (t0[e] + t1), h = 0.5f * (t1 - t0[e]);
            float acc = 0.0f;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
#pragma unroll
                for (int sgn = -1; sgn <= 1; sgn += 2) {
                    float l2, w;
                    acc = fmaf(kGLw[i], t_fs(fmaf(h, sgn * kGLx[i], m), c.inv_nu, c.B2, eref[e], l2, w), acc);
                }
            }
            S[e] = fmaf(h, acc, S[e]);
            t0[e] = t1;
        }
    }
#pragma unroll
    for (int e = 0; e < E; ++e) out[e] = active ? -(c.lc2 + eref[e] + __log2f(S[e])) : 0.0f;
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

// `active`: whether this lane holds a real element.  The cdf_diff modes take a warp-wide maximum of the panel count, so every
// lane of the warp must make the call (inactive lanes ask for no panels); the other modes ignore the flag.
template <int MODE>
__device__ __forceinline__ float elem_nll(float x, bool active, const TConst &tc, const GConst &gc, float sg, float nu) {
    if (MODE == MODE_GAUSS) return fmaf(gc.Bc, x * x, gc.A);
    if (MODE == MODE_T_SPATIAL) return t_nll(x, t_const(sg, nu));
    if (is_cdf(MODE)) {
        const float xs[1] = {x};
        float r[1];
        if (MODE == MODE_CDF_SPATIAL) uniform_cdf_nll<1>(xs, active, t_const(active ? sg : 1.0f, active ? nu : 4.0f), r);
        else uniform_cdf_nll<1>(xs, active, tc, r);
        return r[0];
    }
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
// QUANT and HAS_MU are compile-time so that the prefetch registers of unused optional inputs (noise tensor, mu map)
// disappear: the common training instance (broadcast, Philox, no mu) must stay <= 64 registers for 4 CTAs per SM.
template <int MODE, bool VEC, int QUANT, bool HAS_MU>
__global__ void __launch_bounds__(kThreads, MODE == MODE_CDF_SPATIAL ? 3 : 4) bottleneck_fwd_kernel(
    const float *__restrict__ y, const float *__restrict__ noise, uint64_t *__restrict__ philox,
    const float *__restrict__ mu_in, const float *__restrict__ sigma, const float *__restrict__ nu, Shape sh,
    int mu_layout, float *__restrict__ y_tilde, float *__restrict__ nll, float *__restrict__ bits, float *__restrict__ psum,
    unsigned int *__restrict__ ticket) {
    constexpr int quant_mode = QUANT;
    const float *__restrict__ mu = HAS_MU ? mu_in : nullptr;
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
        if (is_bcast(MODE)) tc = t_const(__ldg(sigma + row), __ldg(nu + row));
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
            // vectors in flight per lane (spatial mode streams 3 tensors; the cdf_diff modes are SFU/issue-bound: one vector, the
            // registers go to the 32 independent quadrature chains instead)
            constexpr int U = is_cdf(MODE) ? 1 : is_spatial(MODE) ? 2 : 4;
            const int v0 = e0 >> 2, v1 = e1 >> 2;
            const long vbase = base >> 2;
            for (int vb = v0; vb < v1; vb += 32 * U) {   // warp-uniform trip count (the cdf_diff modes are warp-cooperative)
                const int v = vb + lane;
                float4 yy[U], nn[U], ss[U], uu[U], mm[U];
#pragma unroll
                for (int k = 0; k < U; ++k) {
                    int vv = v + 32 * k;
                    if (vv < v1) {
                        long gi = vbase + vv;
                        yy[k] = ldg_stream(reinterpret_cast<const float4 *>(y) + gi);
                        if (quant_mode == SIC_QUANT_NOISE_TENSOR) nn[k] = ldg_stream(reinterpret_cast<const float4 *>(noise) + gi);
                        if (is_spatial(MODE)) {
                            ss[k] = ldg_stream(reinterpret_cast<const float4 *>(sigma) + gi);
                            uu[k] = ldg_stream(reinterpret_cast<const float4 *>(nu) + gi);
                        }
                        if (mu != nullptr && mu_layout == SIC_PARAM_SPATIAL) mm[k] = ldg_stream(reinterpret_cast<const float4 *>(mu) + gi);
                    }
                }
#pragma unroll
                for (int k = 0; k < U; ++k) {
                    const int vv = v + 32 * k;
                    const bool ok = vv < v1;
                    const long gi = vbase + vv;
                    float4 q = make_float4(0.f, 0.f, 0.f, 0.f), l, m4 = make_float4(mu_row, mu_row, mu_row, mu_row);
                    if (ok) {
                        if (quant_mode == SIC_QUANT_NOISE_PHILOX) {
                            uint4 r = philox4x32_10(make_uint4((uint32_t)gi, (uint32_t)((uint64_t)gi >> 32), off.x, off.y), key);
                            nn[k] = make_float4(u32_to_noise(r.x), u32_to_noise(r.y), u32_to_noise(r.z), u32_to_noise(r.w));
                        }
                        if (mu != nullptr && mu_layout == SIC_PARAM_SPATIAL) m4 = mm[k];
                        q.x = quantize1(yy[k].x, quant_mode, nn[k].x);
                        q.y = quantize1(yy[k].y, quant_mode, nn[k].y);
                        q.z = quantize1(yy[k].z, quant_mode, nn[k].z);
                        q.w = quantize1(yy[k].w, quant_mode, nn[k].w);
                    }
                    if (MODE == MODE_CDF_BCAST) {      // the lane's four elements share one warp-uniform panel loop
                        const float xs[4] = {q.x - m4.x, q.y - m4.y, q.z - m4.z, q.w - m4.w};
                        float r[4];
                        uniform_cdf_nll<4>(xs, ok, tc, r);
                        l = make_float4(r[0], r[1], r[2], r[3]);
                    } else {
                        l.x = elem_nll<MODE>(q.x - m4.x, ok, tc, gc, ss[k].x, uu[k].x);
                        l.y = elem_nll<MODE>(q.y - m4.y, ok, tc, gc, ss[k].y, uu[k].y);
                        l.z = elem_nll<MODE>(q.z - m4.z, ok, tc, gc, ss[k].z, uu[k].z);
                        l.w = elem_nll<MODE>(q.w - m4.w, ok, tc, gc, ss[k].w, uu[k].w);
                    }
                    if (ok) {
                        if (y_tilde != nullptr) stg_stream(reinterpret_cast<float4 *>(y_tilde) + gi, q);
                        if (nll != nullptr) stg_stream(reinterpret_cast<float4 *>(nll) + gi, l);
                        acc += (l.x + l.y) + (l.z + l.w);
                    }
                }
            }
        } else {
            for (int eb = e0; eb < e1; eb += 32) {   // warp-uniform trip count
                const int e = eb + lane;
                const bool ok = e < e1;
                const long gi = base + e;
                float q = 0.f, m1 = mu_row, sg = 0.f, nv = 0.f;
                if (ok) {
                    float n1 = 0.0f;
                    if (quant_mode == SIC_QUANT_NOISE_TENSOR) n1 = noise[gi];
                    if (quant_mode == SIC_QUANT_NOISE_PHILOX) {
                        long v = gi >> 2;
                        uint4 r = philox4x32_10(make_uint4((uint32_t)v, (uint32_t)((uint64_t)v >> 32), off.x, off.y), key);
                        int j = (int)(gi & 3);
                        n1 = u32_to_noise(j == 0 ? r.x : j == 1 ? r.y : j == 2 ? r.z : r.w);
                    }
                    q = quantize1(y[gi], quant_mode, n1);
                    if (mu != nullptr && mu_layout == SIC_PARAM_SPATIAL) m1 = mu[gi];
                    if (is_spatial(MODE)) { sg = sigma[gi]; nv = nu[gi]; }
                }
                float l = elem_nll<MODE>(q - m1, ok, tc, gc, sg, nv);
                if (ok) {
                    if (y_tilde != nullptr) y_tilde[gi] = q;
                    if (nll != nullptr) nll[gi] = l;
                    acc += l;
                }
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
    float sigma2nu, inv_sigma2nu, np1, inv_sigma, inv_nu, mask_s, mask_n, Knu;
};
__device__ __forceinline__ TBwdConst t_bwd_const(float sigma_raw, float nu_raw) {
    float s = clamp_keep_nan(sigma_raw, kSigmaMin, kSigmaMax);
    float n = clamp_keep_nan(nu_raw, kNuMin, kNuMax);
    TBwdConst c;
    c.sigma2nu = s * s * n;
    c.inv_sigma2nu = 1.0f / c.sigma2nu;
    c.np1 = n + 1.0f;
    c.inv_sigma = 1.0f / s;
    c.inv_nu = 1.0f / n;
    c.mask_s = (sigma_raw >= kSigmaMin && sigma_raw <= kSigmaMax) ? 1.0f : 0.0f;  // torch.clamp passes grad on the closed interval
    c.mask_n = (nu_raw >= kNuMin && nu_raw <= kNuMax) ? 1.0f : 0.0f;
    // Knu = 0.5*(psi((nu+1)/2) - psi(nu/2)) - 1/(2 nu)  ~ 1/(4 nu^2): the three terms cancel to 1e-4 of their size, which is why
    // this used to be a binary64 chain (12 DDIVs per lane per unit, and per ELEMENT in the spatial layout).  Shifting
    // a = nu/2 by 4 with the psi recurrence and telescoping 1/(4a) the same way leaves a sum of POSITIVE terms that float32
    // carries to 4e-7 relative (checked against scipy digamma over nu in [2,100]):
    //   Knu = sum_{k<4} (1/8) / ((a+k)(a+k+1/2)(a+k+1))  +  rb^2 (1/16 - rb^2/128 + rb^4/256),   rb = 1/(a+4)
    const float a = 0.5f * n;
    float acc = 0.0f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float ak = a + (float)k;
        acc += __fdividef(0.125f, ak * (ak + 0.5f) * (ak + 1.0f));
    }
    const float rb = __fdividef(1.0f, a + 4.0f), rb2 = rb * rb;
    c.Knu = acc + rb2 * (0.0625f + rb2 * (-0.0078125f + rb2 * 0.00390625f));
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
    // log1p(u) = ln2 * log2(1+u); the rounding of 1+u costs <= 6e-8 absolute on a term that is summed against Knu ~ 1/(4 nu^2)
    float l1p = 0.69314718055994531f * __log2f(fmaf(q, c.inv_sigma2nu, 1.0f));
    dn = -kLog2e * (c.Knu - 0.5f * l1p + 0.5f * rq * c.inv_nu);
}

// cdf_diff backward: d(-log2 P)/d{x, sigma, nu} = -log2e * dP/P with
//   dP/dx = (f(hi) - f(lo))/sigma,  dP/dsigma = -(hi f(hi) - lo f(lo))/sigma,  dP/dnu = int f * dlog f/dnu
//   dlog f/dnu = Knu - 1/2 log1p(t^2/nu) + (nu+1)/(2 nu) * (1 - 1/(1+t^2/nu))
__device__ __forceinline__ void cdf_bwd_elem(float x, const TConst &c, const TBwdConst &bc, float &dx, float &ds, float &dn) {
    float S, Q, eref, lo, hi;
    t_bin<true>(x, c, bc.Knu, 0.5f * bc.np1 * c.inv_nu, S, Q, eref, lo, hi);
    // f(hi) - f(lo) cancels when the bin is narrow against the scale (sigma >> 1): form it as f(lo) * expm1(log ratio),
    // with the ratio of (1 + t^2/nu) taken from the exact difference hi^2 - lo^2 = (hi - lo)(hi + lo).
    float l2, w_lo;
    float f_lo = t_fs(lo, c.inv_nu, c.B2, eref, l2, w_lo);
    float delta = (hi - lo) * (hi + lo) * c.inv_nu / w_lo;            // w_hi / w_lo - 1  (> -1)
    float df = f_lo * expm1f(-c.B2 * log1pf(delta));                  // f(hi) - f(lo), scaled
    float f_hi = f_lo + df;
    float k = -kLog2e * __fdividef(1.0f, S);
    dx = k * df * c.inv_sigma;
    ds = -k * ((hi - lo) * f_hi + lo * df) * c.inv_sigma;
    dn = k * Q;
}

template <int MODE, bool VEC>
__global__ void __launch_bounds__(kThreads, MODE == MODE_T_BCAST ? 4 : 3) bottleneck_bwd_kernel(
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
        TConst fc{};
        if (is_bcast(MODE)) { tc = t_bwd_const(__ldg(sigma + row), __ldg(nu + row)); if (is_cdf(MODE)) fc = t_const(__ldg(sigma + row), __ldg(nu + row)); }
        if (MODE == MODE_GAUSS) {
            float sr = expf(__ldg(sigma + row % sh.C));
            float s = clamp_keep_nan(sr, kSigmaMin, kSigmaMax);
            g_inv_var = 1.0f / (s * s);
            g_mask = (sr >= kSigmaMin && sr <= kSigmaMax) ? 1.0f : 0.0f;
        }
        float mu_row = 0.0f;
        if (mu != nullptr && mu_layout != SIC_PARAM_SPATIAL) mu_row = __ldg(mu + (mu_layout == SIC_PARAM_CHANNEL ? row % sh.C : row));
        const float gb = g_bits != nullptr ? __ldg(g_bits + b) : 0.0f;
        float acc_s = 0.f, acc_n = 0.f, acc_m = 0.f, acc_0 = 0.f;

        auto one = [&](long gi, float ytv, float gl, float gy, float sg, float nv, float m1, float &o_dy, float &o_ds, float &o_dn, float &o_dm) {
            float gt = gl + gb;
            float x = ytv - m1;
            float dxe, dse = 0.f, dne = 0.f;
            if (MODE == MODE_GAUSS) {
                dxe = kLog2e * x * g_inv_var;
                dse = kLog2e * (1.0f - x * x * g_inv_var) * g_mask;  // d/dlog_sigma
            } else if (is_spatial(MODE)) {
                TBwdConst c = t_bwd_const(sg, nv);
                if (is_cdf(MODE)) cdf_bwd_elem(x, t_const(sg, nv), c, dxe, dse, dne);
                else t_bwd_elem(x, c, dxe, dse, dne);
                dse *= c.mask_s;
                dne *= c.mask_n;
            } else if (is_cdf(MODE)) {
                cdf_bwd_elem(x, fc, tc, dxe, dse, dne);
            } else {
                // broadcast density (the training instance): only the raw sums are accumulated per element,
                //   acc_s <- sum gt (1 - rq),  acc_n <- sum gt log2(1+u),  acc_0 <- sum gt ;  the row constants multiply once per unit
                float q = x * x;
                float rc = __fdividef(1.0f, tc.sigma2nu + q);
                float rq = tc.np1 * q * rc;
                dxe = kLog2e * tc.np1 * x * rc;
                dse = 1.0f - rq;
                dne = __log2f(fmaf(q, tc.inv_sigma2nu, 1.0f));
                acc_0 += gt;
            }
            float gx = gt * dxe;
            o_dy = pass_dy ? gy + gx : 0.0f;
            o_ds = gt * dse;
            o_dn = gt * dne;
            o_dm = -gx;
        };

        if (MODE == MODE_T_BCAST && VEC && g_nll == nullptr && mu == nullptr && dmu == nullptr) {
            // The training call: the loss reaches the likelihood only through the per-patch bit counts (g_bits), there is no mu and
            // sigma / nu are per row.  Then the upstream gradient gt = g_bits[b] is a row constant: it is factored out of the three
            // sums (applied once per lane below) and folded into the dy coefficient; no dmu work at all.  r02o: the generic loop
            // costs 45 instructions per element (126 M warp instructions, 66 % issue utilisation, 80 % of HBM); this one 13.
            constexpr int U = 2;
            const int v0 = e0 >> 2, v1 = e1 >> 2;
            const long vbase = base >> 2;
            const bool has_gy = g_yt != nullptr && pass_dy;
            const float cdx = pass_dy ? kLog2e * gb : 0.0f;          // dy = gy + gt * log2e * (nu+1) x / (nu sigma^2 + x^2)
            float cnt = 0.f;
            auto elem = [&](float x, float gyv, float &o) {
                const float q = x * x;
                const float t = tc.np1 * __fdividef(1.0f, tc.sigma2nu + q);
                acc_s += 1.0f - t * q;
                acc_n += __log2f(fmaf(q, tc.inv_sigma2nu, 1.0f));
                o = fmaf(cdx * t, x, gyv);
            };
            for (int v = v0 + lane; v < v1; v += 32 * U) {
                float4 a[U], gy[U];
#pragma unroll
                for (int k = 0; k < U; ++k) {
                    const int vv = v + 32 * k;
                    if (vv < v1) {
                        a[k] = ldg_stream(reinterpret_cast<const float4 *>(yt) + vbase + vv);
                        gy[k] = has_gy ? ldg_stream(reinterpret_cast<const float4 *>(g_yt) + vbase + vv) : make_float4(0, 0, 0, 0);
                    }
                }
#pragma unroll
                for (int k = 0; k < U; ++k) {
                    const int vv = v + 32 * k;
                    if (vv < v1) {
                        float4 o;
                        elem(a[k].x, gy[k].x, o.x);
                        elem(a[k].y, gy[k].y, o.y);
                        elem(a[k].z, gy[k].z, o.z);
                        elem(a[k].w, gy[k].w, o.w);
                        cnt += 4.0f;
                        if (dy) stg_stream(reinterpret_cast<float4 *>(dy) + vbase + vv, o);
                    }
                }
            }
            acc_s *= gb;            // sum gt (1 - rq)
            acc_n *= gb;            // sum gt log2(1 + u)
            acc_0 = gb * cnt;       // sum gt
        } else if (VEC) {
            constexpr int U = is_spatial(MODE) ? 1 : 2;
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
                        if (is_spatial(MODE)) {
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
                        if (is_spatial(MODE)) {
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
                if (is_spatial(MODE)) { sg = sigma[gi]; nv = nu[gi]; }
                float m1 = (mu != nullptr && mu_layout == SIC_PARAM_SPATIAL) ? mu[gi] : mu_row;
                float o, s1, n1, d1;
                one(gi, yt[gi], g_nll ? g_nll[gi] : 0.f, (g_yt && pass_dy) ? g_yt[gi] : 0.f, sg, nv, m1, o, s1, n1, d1);
                if (dy) dy[gi] = o;
                if (is_spatial(MODE)) {
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
        if (MODE == MODE_T_BCAST) {
            acc_0 = warp_sum(acc_0);
            // ds = L2E/sigma * sum gt(1-rq);  dn = -L2E (Knu S0 - ln2/2 sum gt log2(1+u) + (S0 - sum gt(1-rq)) / (2 nu))
            float raw_s = acc_s, raw_l = acc_n;
            acc_s = kLog2e * tc.inv_sigma * raw_s;
            acc_n = -kLog2e * (tc.Knu * acc_0 - 0.34657359027997264f * raw_l + 0.5f * tc.inv_nu * (acc_0 - raw_s));
        }
        if (lane == 0) {
            if (is_bcast(MODE)) { acc_s *= tc.mask_s; acc_n *= tc.mask_n; }
            pa[unit] = acc_s;
            pb[unit] = acc_n;
            pc[unit] = acc_m;
        }
    }
    if (retire_and_check_last(ticket)) {
        // fold partials: broadcast -> per row over segments; channel -> per channel over (b, segment), fixed order
        const long rows = (long)sh.B * sh.C;
        if (!is_spatial(MODE)) {
            if (is_bcast(MODE)) {
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

}  // namespace
}  // namespace sic
