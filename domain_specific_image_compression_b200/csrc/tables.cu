// K3: integer CDF tables for the entropy coder, K4: symbols + per-patch support.   Compile this TU with -fmad=false.
//
// Reference: /root/reference/code/modelv2/eval_selfcontained_entropy.py
//   :14-15  gaussian_cdf            :17-23  pmf_to_uint16_cdf
//   :37-47  z support + Gaussian PMF + table          :51-61  y support + Student-t PMF + table
//   :39-40,48 / :52-53,62  per-patch min/max (floor/ceil -+ tail) and int32 symbols
// The reference builds one table per (symbol, channel, h, w) with a Python loop over the batch, two host syncs per patch
// per latent, and calls a Student-t CDF that torch does not implement.  Here: one warp per table row, the L+1 edge CDFs
// evaluated by the lanes in parallel in binary64 following spec SIC-CDF-1 — only IEEE + - * / and integer
// bit moves in a fixed order, so the decoder (any device, any compiler) regenerates identical uint16 tables.
#include "common.cuh"

namespace sic {
namespace {

// ---- SIC-CDF-1 scalar functions (binary64, round-to-nearest, no contraction) --------------------------------------
__device__ __forceinline__ double bits2d(unsigned long long u) { return __longlong_as_double((long long)u); }
__device__ __forceinline__ unsigned long long d2bits(double d) { return (unsigned long long)__double_as_longlong(d); }

__device__ double xlog(double x) {
    const double LN2_HI = 6.93147180369123816490e-01, LN2_LO = 1.90821492927058770002e-10;
    unsigned long long b = d2bits(x);
    int e = (int)(b >> 52) - 1023;
    double m = bits2d((b & 0x000fffffffffffffULL) | 0x3ff0000000000000ULL);
    if (m > 1.41421356237309514547e+00) { m = m * 0.5; e = e + 1; }
    double s = (m - 1.0) / (m + 1.0);
    double z = s * s;
    double p = 1.0 / 25.0;
#pragma unroll 1
    for (int k = 11; k >= 0; --k) {
        double c = 1.0 / (double)(2 * k + 1);
        p = p * z;
        p = p + c;
    }
    double r = 2.0 * s;
    r = r * p;
    double de = (double)e;
    double hi = de * LN2_HI;
    double lo = de * LN2_LO;
    lo = lo + r;
    return hi + lo;
}

__device__ double xexp(double x) {
    const double LN2_HI = 6.93147180369123816490e-01, LN2_LO = 1.90821492927058770002e-10;
    if (x < -700.0) return 0.0;
    if (x > 700.0) x = 700.0;
    double kf = floor(x * 1.44269504088896338700e+00 + 0.5);
    double r = x - kf * LN2_HI;
    r = r - kf * LN2_LO;
    double p = 1.0;
#pragma unroll 1
    for (int n = 16; n >= 1; --n) {
        double t = r * p;
        t = t / (double)n;
        p = 1.0 + t;
    }
    int k = (int)kf;
    return p * bits2d((unsigned long long)(k + 1023) << 52);
}

__device__ double xlgamma(double x) {
    double acc = 1.0;
    while (x < 12.0) { acc = acc * x; x = x + 1.0; }
    double xi = 1.0 / x;
    double xi2 = xi * xi;
    double ser = 1.0 / 156.0;
    ser = ser * xi2; ser = ser + (-691.0 / 360360.0);
    ser = ser * xi2; ser = ser + (1.0 / 1188.0);
    ser = ser * xi2; ser = ser + (-1.0 / 1680.0);
    ser = ser * xi2; ser = ser + (1.0 / 1260.0);
    ser = ser * xi2; ser = ser + (-1.0 / 360.0);
    ser = ser * xi2; ser = ser + (1.0 / 12.0);
    ser = ser * xi;
    double r = (x - 0.5) * xlog(x);
    r = r - x;
    r = r + 9.18938533204672780563e-01;
    r = r + ser;
    return r - xlog(acc);
}

__device__ double xbetacf(double a, double b, double x) {
    const double TINY = 1e-300, EPS = 1e-16;
    double qab = a + b, qap = a + 1.0, qam = a - 1.0;
    double c = 1.0;
    double d = qab * x; d = d / qap; d = 1.0 - d;
    if (fabs(d) < TINY) d = TINY;
    d = 1.0 / d;
    double h = d;
#pragma unroll 1
    for (int m = 1; m <= 300; ++m) {
        double dm = (double)m, m2 = (double)(2 * m);
        double num = dm * (b - dm); num = num * x;
        double den = (qam + m2) * (a + m2);
        double aa = num / den;
        double t = aa * d; d = 1.0 + t; if (fabs(d) < TINY) d = TINY;
        t = aa / c; c = 1.0 + t; if (fabs(c) < TINY) c = TINY;
        d = 1.0 / d;
        t = d * c; h = h * t;
        num = (a + dm) * (qab + dm); num = num * x;
        den = (a + m2) * (qap + m2);
        aa = -(num / den);
        t = aa * d; d = 1.0 + t; if (fabs(d) < TINY) d = TINY;
        t = aa / c; c = 1.0 + t; if (fabs(c) < TINY) c = TINY;
        d = 1.0 / d;
        double del = d * c;
        h = h * del;
        if (fabs(del - 1.0) < EPS) break;
    }
    return h;
}

// Student-t CDF (loc 0, scale 1); lbeta = lgamma(nu/2) + lgamma(1/2) - lgamma(nu/2 + 1/2) is a row constant
__device__ double xtcdf(double t, double nu, double lbeta) {
    if (t == 0.0) return 0.5;
    if (!(fabs(t) < 1e100)) return t > 0.0 ? 1.0 : 0.0;
    double t2 = t * t;
    double den = nu + t2;
    double x = nu / den;
    double xc = t2 / den;
    double a = 0.5 * nu;
    double lx = a * xlog(x);
    double lxc = 0.5 * xlog(xc);
    double front = xexp((lx + lxc) - lbeta);
    double tail;
    if (x < (a + 1.0) / (a + 2.5)) {
        double I = front * xbetacf(a, 0.5, x);
        I = I / a;
        tail = 0.5 * I;
    } else {
        double J = front * xbetacf(0.5, a, xc);
        J = J / 0.5;
        tail = 0.5 * (1.0 - J);
    }
    return t > 0.0 ? 1.0 - tail : tail;
}

__device__ double xerfc(double x) {
    double ax = fabs(x);
    double x2 = ax * ax;
    double r;
    if (ax < 2.0) {
        double term = ax, sum = ax, tx2 = 2.0 * x2;
#pragma unroll 1
        for (int n = 1; n <= 200; ++n) {
            term = term * tx2;
            term = term / (double)(2 * n + 1);
            sum = sum + term;
            if (term < 1e-17 * sum) break;
        }
        double er = 1.12837916709551255856e+00 * xexp(-x2);
        er = er * sum;
        r = 1.0 - er;
    } else {
        const double TINY = 1e-300;
        double f = ax, C = ax, D = 0.0;
#pragma unroll 1
        for (int n = 1; n <= 500; ++n) {
            double an = 0.5 * (double)n;
            double t = an * D; D = ax + t; if (D == 0.0) D = TINY;
            t = an / C; C = ax + t; if (C == 0.0) C = TINY;
            D = 1.0 / D;
            double delta = C * D;
            f = f * delta;
            if (fabs(delta - 1.0) < 1e-16) break;
        }
        r = xexp(-x2) * 5.64189583547756279280e-01;
        r = r / f;
    }
    return x >= 0.0 ? r : 2.0 - r;
}

__device__ double xncdf(double t) {
    if (!(fabs(t) < 1e100)) return t > 0.0 ? 1.0 : 0.0;
    return 0.5 * xerfc(-(t * 7.07106781186547572737e-01));
}

// ---- K3 ---------------------------------------------------------------------------------------------------------------
constexpr int kRowsPerCta = 4;

// dynamic smem: kRowsPerCta * stride floats (edge CDFs, then PMF in place)
__global__ void __launch_bounds__(kRowsPerCta * 32) cdf_tables_kernel(int kind, const float *__restrict__ sigma,
                                                                       const float *__restrict__ nu, int n_rows, int rows_per_patch,
                                                                       int C, const int32_t *__restrict__ mins,
                                                                       const int32_t *__restrict__ maxs, int stride,
                                                                       uint16_t *__restrict__ out) {
    extern __shared__ float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.x * kRowsPerCta + warp;
    if (row >= n_rows) return;
    float *F = smem + (size_t)warp * stride;
    const int patch = row / rows_per_patch;
    const int mn = mins[patch];
    int L = maxs[patch] - mn + 1;
    if (L > stride - 1) L = stride - 1;  // host sizes stride = max L + 1; defensive
    uint16_t *o = out + (size_t)row * stride;
    float sg;
    double dnu = 0.0, lbeta = 0.0;
    if (kind == 0) {
        sg = (float)xexp((double)sigma[row % C]);  // sigma_z = exp(log_sigma), unclamped (eval_selfcontained_entropy.py:32)
    } else {
        sg = sigma[row];
        dnu = (double)nu[row];
        double a = 0.5 * dnu;
        lbeta = xlgamma(a) + 5.72364942924700081938e-01;
        lbeta = lbeta - xlgamma(a + 0.5);
    }
    // edges e_k = (mn + k) - 1/2, k = 0..L ; upper edge of symbol k == lower edge of symbol k+1 exactly in fp32
    for (int k = lane; k <= L; k += 32) {
        float edge = (k == 0) ? (float)mn - 0.5f : (float)(mn + k - 1) + 0.5f;
        float t = edge / sg;
        double Fk = kind == 0 ? xncdf((double)t) : xtcdf((double)t, dnu, lbeta);
        F[k] = (float)Fk;
    }
    __syncwarp();
    // PMF with the 1e-12 floor (:45,:59), in place: F[k] <- pmf_k for k < L.  Lanes own disjoint k; read k+1 before any write
    // of it by staging through registers one 32-wide batch at a time.
    for (int k0 = 0; k0 < L; k0 += 32) {
        int k = k0 + lane;
        float p = 0.f;
        if (k < L) {
            p = F[k + 1] - F[k];
            if (!(p >= 1e-12f)) p = 1e-12f;
        }
        __syncwarp();
        if (k < L) F[k] = p;
        __syncwarp();
    }
    if (lane == 0) {
        double acc = 0.0;
        for (int k = 0; k < L; ++k) acc = acc + (double)F[k];
        float S = (float)acc;  // :46,:60  pmf / pmf.sum
        o[0] = 0;
        acc = 0.0;
        for (int k = 0; k < L; ++k) {
            float q = F[k] / S;
            acc = acc + (double)q;  // :18 cumsum (float64 accumulator, float32 outputs)
            float c = (float)acc;
            if (k == L - 1 && c < 1.0f) c = 1.0f;  // :21
            float sc = c * 65535.0f;               // :22
            if (sc < 0.0f) sc = 0.0f;
            if (sc > 65535.0f) sc = 65535.0f;
            o[k + 1] = (uint16_t)sc;               // truncation, like astype(np.uint16)
        }
    }
    __syncwarp();
    for (int k = L + 1 + lane; k < stride; k += 32) o[k] = 0;
}

// ---- K4 ---------------------------------------------------------------------------------------------------------------
// order-preserving float <-> int map so that atomicMin/atomicMax on ints give float min/max (deterministic: min/max commute)
__device__ __forceinline__ int f2ord(float f) {
    int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__global__ void minmax_init_kernel(int B, int32_t *mins, int32_t *maxs) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B) { mins[b] = 0x7fffffff; maxs[b] = (int)0x80000000; }
}

__global__ void __launch_bounds__(256) minmax_kernel(const float *__restrict__ q, long n_per_patch, int do_round, int blocks_per_patch,
                                                     int32_t *mins, int32_t *maxs) {
    const int b = blockIdx.x / blocks_per_patch, blk = blockIdx.x - b * blocks_per_patch;
    const float *p = q + (long)b * n_per_patch;
    float lo = INFINITY, hi = -INFINITY;
    for (long i = (long)blk * 256 + threadIdx.x; i < n_per_patch; i += (long)blocks_per_patch * 256) {
        float v = p[i];
        if (do_round) v = rintf(v);
        lo = fminf(lo, v);
        hi = fmaxf(hi, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(mins + b, f2ord(lo));
        atomicMax(maxs + b, f2ord(hi));
    }
}

__device__ __forceinline__ int sat_int(float v) {
    v = fminf(fmaxf(v, -2147483000.0f), 2147483000.0f);
    return (int)v;
}

// turn the ordered-int float extrema into  floor(min)-tail / ceil(max)+tail  (:39-40, :52-53)
__global__ void minmax_finish_kernel(int B, int tail, int32_t *mins, int32_t *maxs) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B) {
        mins[b] = sat_int(floorf(ord2f(mins[b]))) - tail;
        maxs[b] = sat_int(ceilf(ord2f(maxs[b]))) + tail;
    }
}

__global__ void __launch_bounds__(256) symbols_kernel(const float *__restrict__ q, long n_per_patch, long n, int do_round,
                                                      const int32_t *__restrict__ mins, int32_t *__restrict__ sym) {
    for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long)gridDim.x * 256) {
        float v = q[i];
        if (do_round) v = rintf(v);
        sym[i] = sat_int(v) - mins[i / n_per_patch];  // int32(q) - min  (:48, :62)
    }
}

SIC_REGISTER_KERNEL("cdf_tables_kernel", cdf_tables_kernel);
SIC_REGISTER_KERNEL("minmax_kernel", minmax_kernel);
SIC_REGISTER_KERNEL("symbols_kernel", symbols_kernel);
}  // namespace
}  // namespace sic

using namespace sic;

extern "C" int sic_quantize_indices(const float *q, int B, long n_per_patch, int do_round, int tail, int32_t *sym,
                                    int32_t *mins, int32_t *maxs, void *stream) {
    SIC_CHECK_ARG(B > 0 && n_per_patch > 0, "sic_quantize_indices: empty input B=%d n=%ld", B, n_per_patch);
    SIC_CHECK_ARG(q && mins && maxs, "sic_quantize_indices: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    long want = (n_per_patch + 256 * 8 - 1) / (256 * 8);
    int bpp = (int)(want < 64 ? (want < 1 ? 1 : want) : 64);
    minmax_init_kernel<<<(B + 127) / 128, 128, 0, st>>>(B, mins, maxs);
    minmax_kernel<<<B * bpp, 256, 0, st>>>(q, n_per_patch, do_round, bpp, mins, maxs);
    minmax_finish_kernel<<<(B + 127) / 128, 128, 0, st>>>(B, tail, mins, maxs);
    if (sym != nullptr) {
        long n = (long)B * n_per_patch;
        long blocks = (n + 255) / 256;
        long cap = (long)sm_count() * 16;
        symbols_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, st>>>(q, n_per_patch, n, do_round, mins, sym);
    }
    SIC_CHECK_LAUNCH("sic_quantize_indices");
    return 0;
}

extern "C" int sic_build_cdf_tables(int kind, const float *sigma, const float *nu, int n_rows, int rows_per_patch, int C,
                                    const int32_t *mins, const int32_t *maxs, int stride, uint16_t *out, void *stream) {
    SIC_CHECK_ARG(kind == 0 || kind == 1, "sic_build_cdf_tables: kind must be 0 (Gaussian) or 1 (Student-t)");
    SIC_CHECK_ARG(n_rows > 0 && rows_per_patch > 0 && C > 0 && stride >= 2, "sic_build_cdf_tables: bad extents");
    SIC_CHECK_ARG(sigma && mins && maxs && out && (kind == 0 || nu), "sic_build_cdf_tables: null pointer");
    SIC_CHECK_ARG(stride <= 4097, "sic_build_cdf_tables: support wider than 4096 symbols (stride %d)", stride);
    cudaStream_t st = (cudaStream_t)stream;
    size_t smem = (size_t)kRowsPerCta * stride * sizeof(float);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(cdf_tables_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            set_error("sic_build_cdf_tables: cannot reserve %zu B of shared memory: %s", smem, cudaGetErrorString(e));
            return (int)e;
        }
    }
    cdf_tables_kernel<<<(n_rows + kRowsPerCta - 1) / kRowsPerCta, kRowsPerCta * 32, smem, st>>>(kind, sigma, nu, n_rows, rows_per_patch,
                                                                                              C, mins, maxs, stride, out);
    SIC_CHECK_LAUNCH("sic_build_cdf_tables");
    return 0;
}
