// K2: GDN / IGDN with diagonal gamma — forward (bit-exact replay of the eager op sequence) and backward.
//
// Reference: /root/reference/code/modelv2/layers.py:19-27
//     beta  = beta_param**2 - 2^-18                      (:20)   two roundings
//     gamma = gamma_conv.weight**2 - 2^-18               (:21)   two roundings
//     denom = sqrt(beta + conv2d(x**2, gamma, groups=C)) (:23)   x2=rn(x*x); p=rn(gamma*x2); s=rn(beta+p); d=sqrt_rn(s)
//     y     = x*denom (inverse) | x/denom                (:24-27) rn(x*d) | div_rn(x,d)
// The eager path is 5 launches / ~11 tensor passes; this is one pass at 8 B/element (fwd) and 12 B/element (bwd, the
// denominator is recomputed instead of saved).  Both are HBM-bound streaming kernels: 128-bit L1-bypassing accesses,
// 4 vectors in flight per thread, grid sized to a multiple of the SM count.
//
// Forward uses only __f*_rn intrinsics so ptxas cannot contract mul+add into an FMA: a fused beta+gamma*x2 differs in
// the last bit and would flip round() on latents that sit on a half-integer.
#include "gdn_math.cuh"

namespace sic {
namespace {

using namespace gdnm;

constexpr int kThreads = 256;
constexpr int kUnroll = 4;
constexpr int kChunk4 = 2048;  // float4 per backward unit (8192 elements)

// NCHW vector path: one CTA per (plane, chunk of kChunk4 float4) so that beta/gamma are CTA constants and the only
// per-vector integer work is one add (the first version divided every vector index by HW/4 and C: 15 of its 30
// instructions per element were index math, ncu r01).
template <bool INVERSE>
__global__ void __launch_bounds__(kThreads) gdn_fwd_vec_kernel(const float4 *__restrict__ x, const float *__restrict__ bias_p,
                                                               const float *__restrict__ beta_param,
                                                               const float *__restrict__ gamma_weight, int C, int hw4, int chunks,
                                                               float4 *__restrict__ y) {
    const int plane = blockIdx.x / chunks;
    const int chunk = blockIdx.x - plane * chunks;
    float beta, gamma;
    eff_params(beta_param, gamma_weight, plane % C, beta, gamma);
    const float bias = load_bias(bias_p, plane % C);
    const int v_begin = chunk * kChunk4, v_end = min(v_begin + kChunk4, hw4);
    const float4 *xp = x + (long)plane * hw4;
    float4 *yp = y + (long)plane * hw4;
    for (int v0 = v_begin + threadIdx.x; v0 < v_end; v0 += kThreads * kUnroll) {
        float4 a[kUnroll];
#pragma unroll
        for (int k = 0; k < kUnroll; ++k) {
            int v = v0 + k * kThreads;
            if (v < v_end) a[k] = ldg_stream(xp + v);
        }
#pragma unroll
        for (int k = 0; k < kUnroll; ++k) {
            int v = v0 + k * kThreads;
            if (v < v_end) {
                float4 o;
                o.x = gdn1<INVERSE>(a[k].x, bias, beta, gamma);
                o.y = gdn1<INVERSE>(a[k].y, bias, beta, gamma);
                o.z = gdn1<INVERSE>(a[k].z, bias, beta, gamma);
                o.w = gdn1<INVERSE>(a[k].w, bias, beta, gamma);
                stg_stream(yp + v, o);
            }
        }
    }
}

template <bool INVERSE>
__global__ void __launch_bounds__(kThreads) gdn_fwd_scalar_kernel(const float *__restrict__ x, const float *__restrict__ bias_p,
                                                                  const float *__restrict__ beta_param,
                                                                  const float *__restrict__ gamma_weight, long n, int HW, int C,
                                                                  int channels_last, float *__restrict__ y) {
    for (long i = (long)blockIdx.x * kThreads + threadIdx.x; i < n; i += (long)gridDim.x * kThreads) {
        int c = channels_last ? (int)(i % C) : (int)((i / HW) % C);
        float beta, gamma;
        eff_params(beta_param, gamma_weight, c, beta, gamma);
        y[i] = gdn1<INVERSE>(x[i], load_bias(bias_p, c), beta, gamma);
    }
}

// channels-last (NHWC) vector path: 4 consecutive elements are 4 consecutive channels (C % 4 == 0).  blockDim is a multiple
// of C/4 and every stride is a multiple of blockDim, so a thread keeps ONE channel quad for its whole life: the
// re-parameterised beta/gamma are hoisted into registers and (backward) the per-channel sums need no atomics.
struct Quad {
    float b[4], g[4], a[4];  // beta, gamma, conv bias
};
__device__ __forceinline__ Quad load_quad(const float *__restrict__ bias_p, const float *__restrict__ beta_param,
                                          const float *__restrict__ gamma_weight, int c) {
    Quad q;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        eff_params(beta_param, gamma_weight, c + j, q.b[j], q.g[j]);
        q.a[j] = load_bias(bias_p, c + j);
    }
    return q;
}

// One CTA per CONTIGUOUS chunk of iters * kUnroll * blockDim float4 (32 KB at 256 threads, iters = 2), like the NCHW kernel: the
// hardware hands chunks to SMs as they free up, and concurrently running CTAs touch one moving window of DRAM pages.  Round 1 ran
// this as a persistent grid-stride loop (CTAs stepping gridDim * 32 KB apart): 93 % of the HBM peak against the NCHW kernel's 98 %
// on the same bytes (profiles/r02b_kernel_sweep_quick.json).  5 CTAs per SM: ptxas fits the GDN instance in 48 registers.
template <bool INVERSE, int ITERS>
__global__ void __launch_bounds__(kThreads, 5) gdn_fwd_nhwc_kernel(const float4 *__restrict__ x, const float *__restrict__ bias_p,
                                                                const float *__restrict__ beta_param,
                                                                const float *__restrict__ gamma_weight, unsigned n4, int c4,
                                                                float4 *__restrict__ y) {
    constexpr int iters = ITERS;
    const Quad q = load_quad(bias_p, beta_param, gamma_weight, (int)(threadIdx.x % c4) * 4);
    const unsigned step = blockDim.x * kUnroll;
    unsigned v0 = blockIdx.x * (step * (unsigned)iters) + threadIdx.x;      // chunk base is a multiple of blockDim, hence of c4
#pragma unroll
    for (int it = 0; it < iters; ++it, v0 += step) {
        float4 a[kUnroll];
#pragma unroll
        for (int k = 0; k < kUnroll; ++k) {
            unsigned v = v0 + k * blockDim.x;
            if (v < n4) a[k] = ldg_stream(x + v);
        }
#pragma unroll
        for (int k = 0; k < kUnroll; ++k) {
            unsigned v = v0 + k * blockDim.x;
            if (v < n4) {
                float4 o;
                o.x = gdn1<INVERSE>(a[k].x, q.a[0], q.b[0], q.g[0]);
                o.y = gdn1<INVERSE>(a[k].y, q.a[1], q.b[1], q.g[1]);
                o.z = gdn1<INVERSE>(a[k].z, q.a[2], q.b[2], q.g[2]);
                o.w = gdn1<INVERSE>(a[k].w, q.a[3], q.b[3], q.g[3]);
                stg_stream(y + v, o);
            }
        }
    }
}

// ---- backward (per-element formulas: gdn_math.cuh) ---------------------------------------------------------------
__device__ __forceinline__ void block_sum3(float &a, float &b, float &c) {
    __shared__ float sa[kThreads / 32], sb[kThreads / 32], sc[kThreads / 32];
    a = warp_sum(a);
    b = warp_sum(b);
    c = warp_sum(c);
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { sa[warp] = a; sb[warp] = b; sc[warp] = c; }
    __syncthreads();
    if (warp == 0) {
        a = lane < kThreads / 32 ? sa[lane] : 0.f;
        b = lane < kThreads / 32 ? sb[lane] : 0.f;
        c = lane < kThreads / 32 ? sc[lane] : 0.f;
        a = warp_sum(a);
        b = warp_sum(b);
        c = warp_sum(c);
    }
}

// one CTA = one (plane, chunk) unit; partials laid out [3][C][B*chunks] (dbeta, dgamma, dbias) so the finalize kernel
// reads them contiguously
template <bool INVERSE, bool VEC>
__global__ void __launch_bounds__(kThreads) gdn_bwd_kernel(const float *__restrict__ x, const float *__restrict__ bias_p,
                                                           const float *__restrict__ g, const float *__restrict__ beta_param,
                                                           const float *__restrict__ gamma_weight, int C, int HW, int chunks,
                                                           float *__restrict__ dx, float *__restrict__ part) {
    const int plane = blockIdx.x / chunks;
    const int chunk = blockIdx.x - plane * chunks;
    const int c = plane % C, b = plane / C;
    float beta, gamma;
    eff_params(beta_param, gamma_weight, c, beta, gamma);
    const float bias = load_bias(bias_p, c);
    float acc_b = 0.f, acc_g = 0.f, acc_x = 0.f;
    const long base = (long)plane * HW;
    if (VEC) {
        const int hw4 = HW >> 2;
        const int v_begin = chunk * kChunk4, v_end = min(v_begin + kChunk4, hw4);
        const float4 *x4 = reinterpret_cast<const float4 *>(x + base), *g4 = reinterpret_cast<const float4 *>(g + base);
        float4 *dx4 = reinterpret_cast<float4 *>(dx + base);
        constexpr int U = 2;
        for (int v0 = v_begin + threadIdx.x; v0 < v_end; v0 += kThreads * U) {
            float4 xa[U], ga[U];
#pragma unroll
            for (int k = 0; k < U; ++k) {
                int v = v0 + k * kThreads;
                if (v < v_end) { xa[k] = ldg_stream(x4 + v); ga[k] = ldg_stream(g4 + v); }
            }
#pragma unroll
            for (int k = 0; k < U; ++k) {
                int v = v0 + k * kThreads;
                if (v < v_end) {
                    float4 o;
                    float hb, hg;
                    gdn_bwd1<INVERSE>(xa[k].x + bias, ga[k].x, beta, gamma, o.x, hb, hg); acc_b += hb; acc_g += hg;
                    gdn_bwd1<INVERSE>(xa[k].y + bias, ga[k].y, beta, gamma, o.y, hb, hg); acc_b += hb; acc_g += hg;
                    gdn_bwd1<INVERSE>(xa[k].z + bias, ga[k].z, beta, gamma, o.z, hb, hg); acc_b += hb; acc_g += hg;
                    gdn_bwd1<INVERSE>(xa[k].w + bias, ga[k].w, beta, gamma, o.w, hb, hg); acc_b += hb; acc_g += hg;
                    acc_x += (o.x + o.y) + (o.z + o.w);
                    stg_stream(dx4 + v, o);
                }
            }
        }
    } else {
        const int e_begin = chunk * kChunk4 * 4, e_end = min(e_begin + kChunk4 * 4, HW);
        for (int e = e_begin + threadIdx.x; e < e_end; e += kThreads) {
            float o, hb, hg;
            gdn_bwd1<INVERSE>(x[base + e] + bias, g[base + e], beta, gamma, o, hb, hg);
            dx[base + e] = o;
            acc_b += hb;
            acc_g += hg;
            acc_x += o;
        }
    }
    block_sum3(acc_b, acc_g, acc_x);
    if (threadIdx.x == 0) {
        const long P = (long)gridDim.x / C;  // = B * chunks partials per channel
        long slot = (long)c * P + (long)b * chunks + chunk;
        part[slot] = acc_b;
        part[(long)C * P + slot] = acc_g;
        part[2 * (long)C * P + slot] = acc_x;
    }
}

// NHWC backward: one CTA per contiguous chunk of iters * U * blockDim float4 of x (and of g), thread-private sums for the thread's
// channel quad, one shared-memory fold per CTA, partials [3][C][gridDim.x] folded per channel by the finalize kernel.
// Round 1 used 592 persistent CTAs with a statically dealt bulk and a finely dealt tail: 88-89 % of the HBM peak at the 256^2 site
// where the NCHW kernel (one CTA per chunk, dynamic dispatch) reaches 99 % on the same bytes.  The price of the chunked grid is one
// partial per (chunk, channel): 0.8 % extra traffic at 4096-float4 chunks.
constexpr int kBwdNhwcU = 2;   // float4 of x and of g in flight per thread and pass

template <bool INVERSE>
__global__ void __launch_bounds__(kThreads, 4) gdn_bwd_nhwc_kernel(const float4 *__restrict__ x, const float *__restrict__ bias_p,
                                                                const float4 *__restrict__ g, const float *__restrict__ beta_param,
                                                                const float *__restrict__ gamma_weight, unsigned n4, int C, int iters,
                                                                float4 *__restrict__ dx, float *__restrict__ part) {
    extern __shared__ float sm[];  // [blockDim.x][12]
    const int c4 = C >> 2;
    const int cq = (int)(threadIdx.x % c4);
    const Quad q = load_quad(bias_p, beta_param, gamma_weight, cq * 4);
    float ab[4] = {0.f, 0.f, 0.f, 0.f}, ag[4] = {0.f, 0.f, 0.f, 0.f}, ax[4] = {0.f, 0.f, 0.f, 0.f};
    constexpr int U = kBwdNhwcU;
    const unsigned step = blockDim.x * U;
    unsigned v0 = blockIdx.x * (step * (unsigned)iters) + threadIdx.x;   // every stride is a multiple of blockDim, hence of C/4
    for (int it = 0; it < iters; ++it, v0 += step) {
        float4 xa[U], ga[U];
#pragma unroll
        for (int k = 0; k < U; ++k) {
            unsigned v = v0 + k * blockDim.x;
            if (v < n4) { xa[k] = ldg_stream(x + v); ga[k] = ldg_stream(g + v); }
        }
#pragma unroll
        for (int k = 0; k < U; ++k) {
            unsigned v = v0 + k * blockDim.x;
            if (v < n4) {
                float4 o;
                float hb, hg;
                gdn_bwd1<INVERSE>(xa[k].x + q.a[0], ga[k].x, q.b[0], q.g[0], o.x, hb, hg); ab[0] += hb; ag[0] += hg; ax[0] += o.x;
                gdn_bwd1<INVERSE>(xa[k].y + q.a[1], ga[k].y, q.b[1], q.g[1], o.y, hb, hg); ab[1] += hb; ag[1] += hg; ax[1] += o.y;
                gdn_bwd1<INVERSE>(xa[k].z + q.a[2], ga[k].z, q.b[2], q.g[2], o.z, hb, hg); ab[2] += hb; ag[2] += hg; ax[2] += o.z;
                gdn_bwd1<INVERSE>(xa[k].w + q.a[3], ga[k].w, q.b[3], q.g[3], o.w, hb, hg); ab[3] += hb; ag[3] += hg; ax[3] += o.w;
                stg_stream(dx + v, o);
            }
        }
    }
    float *mine = sm + threadIdx.x * 12;
#pragma unroll
    for (int j = 0; j < 4; ++j) { mine[j] = ab[j]; mine[4 + j] = ag[j]; mine[8 + j] = ax[j]; }
    __syncthreads();
    if ((int)threadIdx.x < c4) {  // fold the blockDim/c4 threads that share this quad, fixed order
        float sb[4] = {0.f, 0.f, 0.f, 0.f}, sg[4] = {0.f, 0.f, 0.f, 0.f}, sx[4] = {0.f, 0.f, 0.f, 0.f};
        for (unsigned t = threadIdx.x; t < blockDim.x; t += c4) {
#pragma unroll
            for (int j = 0; j < 4; ++j) { sb[j] += sm[t * 12 + j]; sg[j] += sm[t * 12 + 4 + j]; sx[j] += sm[t * 12 + 8 + j]; }
        }
        const long P = gridDim.x;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            long c = cq * 4 + j;
            part[c * P + blockIdx.x] = sb[j];
            part[(long)C * P + c * P + blockIdx.x] = sg[j];
            part[2 * (long)C * P + c * P + blockIdx.x] = sx[j];
        }
    }
}

// one CTA per channel: fixed-order float64 fold + chain rule through the squared re-parameterisation (layers.py:20-21).
// The chunked NHWC backward leaves thousands of partials per channel (8192 at the 256^2 site); with 128 threads and one dependent
// load per iteration the fold alone took ~35 us (r02c: 246 us of main kernel became 281 us per site).  512 threads, four independent
// loads in flight per thread and array: the fold is back to a few microseconds.  Summation order is fixed by (thread, position).
constexpr int kFinThreads = 512;
// grid (C, 3): blockIdx.y = 0 d(beta), 1 d(gamma), 2 d(bias) - three CTAs per channel instead of one (r02d: 9.1 us for the 12.6 MB of
// partials of the 256^2 site on 128 CTAs; the fold is latency-bound, so it wants every SM)
__global__ void __launch_bounds__(kFinThreads) gdn_bwd_finalize_kernel(const float *__restrict__ part, const float *__restrict__ beta_param,
                                                                     const float *__restrict__ gamma_weight, int C, long P,
                                                                     float *__restrict__ dbeta_param, float *__restrict__ dgamma_weight,
                                                                     float *__restrict__ dbias) {
    const int c = blockIdx.x, which = blockIdx.y;
    float *dst = which == 0 ? dbeta_param : (which == 1 ? dgamma_weight : dbias);
    if (dst == nullptr) return;
    const float *pp = part + (long)which * C * P + (long)c * P;
    double s = 0.0;
    long i = threadIdx.x;
    for (; i + 3 * kFinThreads < P; i += 4 * kFinThreads) {
        float v0 = __ldcs(pp + i), v1 = __ldcs(pp + i + kFinThreads), v2 = __ldcs(pp + i + 2 * kFinThreads), v3 = __ldcs(pp + i + 3 * kFinThreads);
        s += ((double)v0 + (double)v1) + ((double)v2 + (double)v3);
    }
    for (; i < P; i += kFinThreads) s += (double)__ldcs(pp + i);
    __shared__ double sh[kFinThreads / 32];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        s = threadIdx.x < kFinThreads / 32 ? sh[threadIdx.x] : 0.0;
        s = warp_sum(s);
        if (threadIdx.x == 0) {
            if (which == 0) s *= 2.0 * (double)beta_param[c];
            else if (which == 1) s *= 2.0 * (double)gamma_weight[c];
            dst[c] = (float)s;
        }
    }
}

inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline int nhwc_threads(int C) { return (C / 4) * (kThreads / (C / 4)); }  // largest multiple of C/4 that is <= 256
// chunk sizing for the NHWC kernels: as many passes per CTA as keep at least ~8 CTAs per SM in the grid (small sites get one
// pass per CTA so that every SM has work), at most `max_iters`
inline int nhwc_iters(long n4, int per_pass, int max_iters) {
    long it = n4 / ((long)per_pass * sm_count() * 8);
    return (int)(it < 1 ? 1 : (it > max_iters ? max_iters : it));
}
// Backward: 4 resident CTAs per SM (launch bounds).  Among the pass counts between half the cap and the cap, take the one whose grid
// fills its last wave best: at the 128^2 sites the cap (8 passes) gives 2048 CTAs = 3.46 waves of 592, 7 passes give 2341 = 3.95.
inline int nhwc_bwd_iters(long n4, int threads) {
    const int cap = nhwc_iters(n4, threads * kBwdNhwcU, n4 > (16L << 20) ? 16 : 8);
    const long slots = 4L * sm_count();
    int best = cap;
    double best_fill = -1.0;
    for (int it = cap; it >= (cap + 1) / 2 && it >= 1; --it) {
        const long chunk = (long)threads * kBwdNhwcU * it, grid = (n4 + chunk - 1) / chunk;
        const long waves = (grid + slots - 1) / slots;
        const double fill = (double)grid / (double)(waves * slots);
        if (fill > best_fill + 0.02) { best_fill = fill; best = it; }   // prefer more passes per CTA (fewer partials) unless clearly better
    }
    return best;
}
inline long nhwc_bwd_grid(long n4, int threads) {
    const long chunk = (long)threads * kBwdNhwcU * nhwc_bwd_iters(n4, threads);
    return (n4 + chunk - 1) / chunk;
}
inline int bwd_chunks(int HW) {
    int per = kChunk4 * 4;
    return (HW + per - 1) / per;
}

SIC_REGISTER_KERNEL("gdn_fwd_vec_kernel<0>", gdn_fwd_vec_kernel<false>);
SIC_REGISTER_KERNEL("gdn_fwd_vec_kernel<1>", gdn_fwd_vec_kernel<true>);
SIC_REGISTER_KERNEL("gdn_fwd_nhwc_kernel<0,2>", gdn_fwd_nhwc_kernel<false, 2>);
SIC_REGISTER_KERNEL("gdn_fwd_nhwc_kernel<1,2>", gdn_fwd_nhwc_kernel<true, 2>);
SIC_REGISTER_KERNEL("gdn_bwd_kernel<0,1>", gdn_bwd_kernel<false, true>);
SIC_REGISTER_KERNEL("gdn_bwd_kernel<1,1>", gdn_bwd_kernel<true, true>);
SIC_REGISTER_KERNEL("gdn_bwd_nhwc_kernel<0>", gdn_bwd_nhwc_kernel<false>);
SIC_REGISTER_KERNEL("gdn_bwd_nhwc_kernel<1>", gdn_bwd_nhwc_kernel<true>);

}  // namespace
}  // namespace sic

using namespace sic;

extern "C" int sic_gdn_fwd(const float *x, const float *bias, const float *beta_param, const float *gamma_weight, int B, int C,
                           int HW, int inverse, int channels_last, float *y, void *stream) {
    SIC_CHECK_ARG(B > 0 && C > 0 && HW > 0, "sic_gdn_fwd: empty shape B=%d C=%d HW=%d", B, C, HW);
    SIC_CHECK_ARG(x && y && beta_param && gamma_weight, "sic_gdn_fwd: null pointer");
    const long n = (long)B * C * HW;
    cudaStream_t st = (cudaStream_t)stream;
    const int sms = sm_count();
    const bool al = aligned16(x) && aligned16(y) && n / 4 < (1L << 31);
    if (!channels_last && HW % 4 == 0 && al) {
        const int hw4 = HW / 4, chunks = (hw4 + kChunk4 - 1) / kChunk4;
        const long units = (long)B * C * chunks;
        SIC_CHECK_ARG(units < (1L << 31), "sic_gdn_fwd: too many units");
        if (inverse) gdn_fwd_vec_kernel<true><<<(unsigned)units, kThreads, 0, st>>>((const float4 *)x, bias, beta_param, gamma_weight, C, hw4, chunks, (float4 *)y);
        else gdn_fwd_vec_kernel<false><<<(unsigned)units, kThreads, 0, st>>>((const float4 *)x, bias, beta_param, gamma_weight, C, hw4, chunks, (float4 *)y);
    } else if (channels_last && C % 4 == 0 && C <= 4 * kThreads && al) {
        unsigned n4 = (unsigned)(n / 4);
        const int c4 = C / 4, threads = nhwc_threads(C);
        const int iters = nhwc_iters(n4, threads * kUnroll, 2);
        const long chunk = (long)threads * kUnroll * iters;
        const unsigned grid = (unsigned)(((long)n4 + chunk - 1) / chunk);
#define SIC_FWD_NHWC(INV, IT) gdn_fwd_nhwc_kernel<INV, IT><<<grid, threads, 0, st>>>((const float4 *)x, bias, beta_param, gamma_weight, n4, c4, (float4 *)y)
        if (inverse) { if (iters == 2) SIC_FWD_NHWC(true, 2); else SIC_FWD_NHWC(true, 1); }
        else { if (iters == 2) SIC_FWD_NHWC(false, 2); else SIC_FWD_NHWC(false, 1); }
#undef SIC_FWD_NHWC
    } else {
        long want = (n + kThreads - 1) / kThreads;
        unsigned grid = (unsigned)(want < (long)sms * 16 ? want : (long)sms * 16);
        if (inverse) gdn_fwd_scalar_kernel<true><<<grid, kThreads, 0, st>>>(x, bias, beta_param, gamma_weight, n, HW, C, channels_last, y);
        else gdn_fwd_scalar_kernel<false><<<grid, kThreads, 0, st>>>(x, bias, beta_param, gamma_weight, n, HW, C, channels_last, y);
    }
    SIC_CHECK_LAUNCH("sic_gdn_fwd");
    return 0;
}

extern "C" size_t sic_gdn_bwd_workspace_bytes(int B, int C, int HW) {
    if (B <= 0 || C <= 0 || HW <= 0) return 0;
    size_t nchw = (size_t)3 * C * B * bwd_chunks(HW) * sizeof(float);
    size_t nhwc = 0;
    if (C % 4 == 0 && C <= 4 * kThreads) nhwc = (size_t)3 * C * nhwc_bwd_grid((long)B * C * HW / 4, nhwc_threads(C)) * sizeof(float);
    return nchw > nhwc ? nchw : nhwc;
}

// number of partials per (array, channel) the main kernel leaves in the workspace for this shape and layout
static long gdn_bwd_partials_per_channel(int B, int C, int HW, int channels_last) {
    if (channels_last) return nhwc_bwd_grid((long)B * C * HW / 4, nhwc_threads(C));
    return (long)B * bwd_chunks(HW);
}

extern "C" int sic_gdn_bwd_fold(const float *beta_param, const float *gamma_weight, int B, int C, int HW, int channels_last, float *dbias,
                                float *dbeta_param, float *dgamma_weight, const void *workspace, size_t workspace_bytes, void *stream) {
    SIC_CHECK_ARG(B > 0 && C > 0 && HW > 0, "sic_gdn_bwd_fold: empty shape B=%d C=%d HW=%d", B, C, HW);
    SIC_CHECK_ARG(beta_param && gamma_weight && workspace, "sic_gdn_bwd_fold: null pointer");
    if (workspace_bytes < sic_gdn_bwd_workspace_bytes(B, C, HW)) {
        set_error("sic_gdn_bwd_fold: workspace %zu < %zu bytes", workspace_bytes, sic_gdn_bwd_workspace_bytes(B, C, HW));
        return SIC_E_WORKSPACE;
    }
    if (channels_last && (C % 4 != 0 || C > 4 * kThreads)) {
        set_error("sic_gdn_bwd_fold: channels_last needs C %% 4 == 0 and C <= %d", 4 * kThreads);
        return SIC_E_UNSUPPORTED;
    }
    const long P = gdn_bwd_partials_per_channel(B, C, HW, channels_last);
    gdn_bwd_finalize_kernel<<<dim3(C, 3), kFinThreads, 0, (cudaStream_t)stream>>>(static_cast<const float *>(workspace), beta_param, gamma_weight, C, P,
                                                                                  dbeta_param, dgamma_weight, dbias);
    SIC_CHECK_LAUNCH("sic_gdn_bwd_fold");
    return 0;
}

extern "C" int sic_gdn_bwd(const float *x, const float *bias, const float *g, const float *beta_param, const float *gamma_weight,
                           int B, int C, int HW, int inverse, int channels_last, float *dx, float *dbias, float *dbeta_param,
                           float *dgamma_weight, void *workspace, size_t workspace_bytes, void *stream) {
    int rc = sic_gdn_bwd_partials(x, bias, g, beta_param, gamma_weight, B, C, HW, inverse, channels_last, dx, workspace, workspace_bytes, stream);
    if (rc != 0) return rc;
    return sic_gdn_bwd_fold(beta_param, gamma_weight, B, C, HW, channels_last, dbias, dbeta_param, dgamma_weight, workspace, workspace_bytes, stream);
}

extern "C" int sic_gdn_bwd_partials(const float *x, const float *bias, const float *g, const float *beta_param, const float *gamma_weight,
                                    int B, int C, int HW, int inverse, int channels_last, float *dx, void *workspace, size_t workspace_bytes,
                                    void *stream) {
    SIC_CHECK_ARG(B > 0 && C > 0 && HW > 0, "sic_gdn_bwd: empty shape B=%d C=%d HW=%d", B, C, HW);
    SIC_CHECK_ARG(x && g && dx && beta_param && gamma_weight && workspace, "sic_gdn_bwd: null pointer");
    if (workspace_bytes < sic_gdn_bwd_workspace_bytes(B, C, HW)) {
        set_error("sic_gdn_bwd: workspace %zu < %zu bytes", workspace_bytes, sic_gdn_bwd_workspace_bytes(B, C, HW));
        return SIC_E_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    float *part = static_cast<float *>(workspace);
    if (channels_last) {
        const long n = (long)B * C * HW;
        if (C % 4 != 0 || C > 4 * kThreads || !aligned16(x) || !aligned16(g) || !aligned16(dx) || n / 4 >= (1L << 31)) {
            set_error("sic_gdn_bwd: channels_last needs C %% 4 == 0, C <= %d and 16-byte aligned tensors", 4 * kThreads);
            return SIC_E_UNSUPPORTED;
        }
        const int threads = nhwc_threads(C);
        const unsigned n4 = (unsigned)(n / 4);
        const int iters = nhwc_bwd_iters(n4, threads);
        const unsigned grid = (unsigned)nhwc_bwd_grid(n4, threads);
        const size_t smem = (size_t)threads * 12 * sizeof(float);
        if (inverse) gdn_bwd_nhwc_kernel<true><<<grid, threads, smem, st>>>((const float4 *)x, bias, (const float4 *)g, beta_param, gamma_weight, n4, C, iters, (float4 *)dx, part);
        else gdn_bwd_nhwc_kernel<false><<<grid, threads, smem, st>>>((const float4 *)x, bias, (const float4 *)g, beta_param, gamma_weight, n4, C, iters, (float4 *)dx, part);
        SIC_CHECK_LAUNCH("sic_gdn_bwd (nhwc)");
        return 0;
    }
    const int chunks = bwd_chunks(HW);
    const long units = (long)B * C * chunks;
    SIC_CHECK_ARG(units < (1L << 31), "sic_gdn_bwd: too many units");
    const bool vec = HW % 4 == 0 && aligned16(x) && aligned16(g) && aligned16(dx);
    if (inverse) {
        if (vec) gdn_bwd_kernel<true, true><<<(unsigned)units, kThreads, 0, st>>>(x, bias, g, beta_param, gamma_weight, C, HW, chunks, dx, part);
        else gdn_bwd_kernel<true, false><<<(unsigned)units, kThreads, 0, st>>>(x, bias, g, beta_param, gamma_weight, C, HW, chunks, dx, part);
    } else {
        if (vec) gdn_bwd_kernel<false, true><<<(unsigned)units, kThreads, 0, st>>>(x, bias, g, beta_param, gamma_weight, C, HW, chunks, dx, part);
        else gdn_bwd_kernel<false, false><<<(unsigned)units, kThreads, 0, st>>>(x, bias, g, beta_param, gamma_weight, C, HW, chunks, dx, part);
    }
    SIC_CHECK_LAUNCH("sic_gdn_bwd");
    return 0;
}
