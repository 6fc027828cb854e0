/* CPU oracle (TEST INFRASTRUCTURE, never linked into the product) for the integer CDF tables.
 *
 * Restates, in plain C, the table recipe of
 *   /root/reference/code/modelv2/eval_selfcontained_entropy.py:14-23  (gaussian_cdf, pmf_to_uint16_cdf)
 *   /root/reference/code/modelv2/eval_selfcontained_entropy.py:37-47  (z support + Gaussian PMF)
 *   /root/reference/code/modelv2/eval_selfcontained_entropy.py:51-61  (y support + Student-t PMF)
 * with the three defects of that script repaired as intent.  torch has no
 * Student-t CDF, so the CDF itself is defined by the spec "SIC-CDF-1" in DESIGN.md: binary64,
 * round-to-nearest, only + - * / and integer bit moves, fixed evaluation order, no FMA contraction
 * (compile with -ffp-contract=off).  The CUDA kernel follows the same written spec independently;
 * the two must agree bit for bit, and both are checked against scipy (stdtr / ndtr) in tests.
 *
 * PARITY UNPINNED by the reference for the table build (its script cannot run); pmf_to_uint16_cdf
 * alone is pinned against the reference's own function (tests/golden/pmf_to_cdf.npz).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

static const double LN2_HI = 6.93147180369123816490e-01;  /* 0x3fe62e42fee00000 */
static const double LN2_LO = 1.90821492927058770002e-10;  /* 0x3dea39ef35793c76 */
static const double INV_LN2 = 1.44269504088896338700e+00;
static const double SQRT2 = 1.41421356237309514547e+00;
static const double INV_SQRT2 = 7.07106781186547572737e-01;
static const double HALF_LOG_2PI = 9.18938533204672780563e-01;
static const double HALF_LOG_PI = 5.72364942924700081938e-01;
static const double TWO_OVER_SQRTPI = 1.12837916709551255856e+00;
static const double INV_SQRTPI = 5.64189583547756279280e-01;

static double u2d(uint64_t u) { double d; memcpy(&d, &u, 8); return d; }
static uint64_t d2u(double d) { uint64_t u; memcpy(&u, &d, 8); return u; }

/* natural log of a positive normal double: 2*atanh((m-1)/(m+1)) on m in [sqrt(.5), sqrt(2)] */
double sic_oracle_log(double x) {
    uint64_t bits = d2u(x);
    int e = (int)(bits >> 52) - 1023;
    double m = u2d((bits & 0x000fffffffffffffULL) | 0x3ff0000000000000ULL);
    if (m > SQRT2) { m = m * 0.5; e = e + 1; }
    double s = (m - 1.0) / (m + 1.0);
    double z = s * s;
    double p = 1.0 / 25.0;
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

double sic_oracle_exp(double x) {
    if (x < -700.0) return 0.0;
    if (x > 700.0) x = 700.0;
    double kf = floor(x * INV_LN2 + 0.5);
    double r = x - kf * LN2_HI;
    r = r - kf * LN2_LO;
    double p = 1.0;
    for (int n = 16; n >= 1; --n) {
        double t = r * p;
        t = t / (double)n;
        p = 1.0 + t;
    }
    int k = (int)kf;
    double scale = u2d((uint64_t)(k + 1023) << 52);
    return p * scale;
}

/* log Gamma for x >= 0.5: shift to x >= 12 then Stirling with 7 Bernoulli terms */
double sic_oracle_lgamma(double x) {
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
    double r = (x - 0.5) * sic_oracle_log(x);
    r = r - x;
    r = r + HALF_LOG_2PI;
    r = r + ser;
    return r - sic_oracle_log(acc);
}

/* continued fraction of the regularised incomplete beta function (modified Lentz) */
static double betacf(double a, double b, double x) {
    const double TINY = 1e-300, EPS = 1e-16;
    double qab = a + b, qap = a + 1.0, qam = a - 1.0;
    double c = 1.0;
    double d = qab * x; d = d / qap; d = 1.0 - d;
    if (fabs(d) < TINY) d = TINY;
    d = 1.0 / d;
    double h = d;
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

/* Student-t CDF, location 0, scale 1, nu > 0 */
double sic_oracle_tcdf(double t, double nu) {
    if (t == 0.0) return 0.5;
    if (!(fabs(t) < 1e100)) return t > 0.0 ? 1.0 : 0.0;
    double t2 = t * t;
    double den = nu + t2;
    double x = nu / den;
    double xc = t2 / den;
    double a = 0.5 * nu;
    double lbeta = sic_oracle_lgamma(a) + HALF_LOG_PI;
    lbeta = lbeta - sic_oracle_lgamma(a + 0.5);
    double lx = a * sic_oracle_log(x);
    double lxc = 0.5 * sic_oracle_log(xc);
    double front = sic_oracle_exp((lx + lxc) - lbeta);
    double tail;
    if (x < (a + 1.0) / (a + 2.5)) {
        double I = front * betacf(a, 0.5, x);
        I = I / a;
        tail = 0.5 * I;
    } else {
        double J = front * betacf(0.5, a, xc);
        J = J / 0.5;
        tail = 0.5 * (1.0 - J);
    }
    return t > 0.0 ? 1.0 - tail : tail;
}

double sic_oracle_erfc(double x) {
    double ax = fabs(x);
    double x2 = ax * ax;
    double r;
    if (ax < 2.0) {
        double term = ax, sum = ax, tx2 = 2.0 * x2;
        for (int n = 1; n <= 200; ++n) {
            term = term * tx2;
            term = term / (double)(2 * n + 1);
            sum = sum + term;
            if (term < 1e-17 * sum) break;
        }
        double erf = TWO_OVER_SQRTPI * sic_oracle_exp(-x2);
        erf = erf * sum;
        r = 1.0 - erf;
    } else {
        const double TINY = 1e-300;
        double f = ax, C = ax, D = 0.0;
        for (int n = 1; n <= 500; ++n) {
            double an = 0.5 * (double)n;
            double t = an * D; D = ax + t; if (D == 0.0) D = TINY;
            t = an / C; C = ax + t; if (C == 0.0) C = TINY;
            D = 1.0 / D;
            double delta = C * D;
            f = f * delta;
            if (fabs(delta - 1.0) < 1e-16) break;
        }
        r = sic_oracle_exp(-x2) * INV_SQRTPI;
        r = r / f;
    }
    return x >= 0.0 ? r : 2.0 - r;
}

/* standard normal CDF; the reference writes it 0.5*(1+erf(x/sqrt 2)) (eval_selfcontained_entropy.py:14-15) */
double sic_oracle_ncdf(double t) {
    if (!(fabs(t) < 1e100)) return t > 0.0 ? 1.0 : 0.0;
    return 0.5 * sic_oracle_erfc(-(t * INV_SQRT2));
}

/* One table row: support mn..mn+L-1, edges k-1/2 .. k+1/2, scale sigma (float32), kind 0 = Gaussian, 1 = Student-t(nu).
 * fp32 data flow mirrors the reference (CDF values, PMF, normalisation and scaling are float32 tensors there);
 * the two reductions are sequential with a float64 accumulator (torch CPU cumsum's accumulation type). */
void sic_oracle_table_row(int kind, float sigma, float nu, int mn, int L, uint16_t *out) {
    float pmf[4096];
    if (L > 4096) L = 4096;
    float lo_edge = (float)mn - 0.5f;
    float prev = (float)(kind ? sic_oracle_tcdf((double)(lo_edge / sigma), (double)nu)
                              : sic_oracle_ncdf((double)(lo_edge / sigma)));
    double acc = 0.0;
    for (int k = 0; k < L; ++k) {
        float up = (float)(mn + k) + 0.5f;
        float tu = up / sigma;
        float Fu = (float)(kind ? sic_oracle_tcdf((double)tu, (double)nu) : sic_oracle_ncdf((double)tu));
        float p = Fu - prev;
        if (!(p >= 1e-12f)) p = 1e-12f;     /* clamp(min=1e-12): NaN-free inputs assumed, NaN maps to the floor */
        pmf[k] = p;
        acc = acc + (double)p;
        prev = Fu;
    }
    float S = (float)acc;
    out[0] = 0;
    acc = 0.0;
    for (int k = 0; k < L; ++k) {
        float q = pmf[k] / S;
        acc = acc + (double)q;
        float c = (float)acc;
        if (k == L - 1 && c < 1.0f) c = 1.0f;
        float sc = c * 65535.0f;
        if (sc < 0.0f) sc = 0.0f;
        if (sc > 65535.0f) sc = 65535.0f;
        out[k + 1] = (uint16_t)sc;
    }
}

/* n rows; row r uses sigma[r], nu[r] (nu ignored for kind 0), support of patch patch_of_row[r]; out stride = stride */
void sic_oracle_tables(int kind, const float *sigma, const float *nu, const int *patch_of_row, const int *mins,
                       const int *maxs, int n_rows, int stride, uint16_t *out) {
    for (int r = 0; r < n_rows; ++r) {
        int b = patch_of_row[r];
        int L = maxs[b] - mins[b] + 1;
        sic_oracle_table_row(kind, sigma[r], kind ? nu[r] : 0.0f, mins[b], L, out + (size_t)r * stride);
    }
}

/* float32 sigma_z = f32(exp64(log_sigma)) : the z prior scale is taken UNCLAMPED (eval_selfcontained_entropy.py:32) */
float sic_oracle_exp_f32(float x) { return (float)sic_oracle_exp((double)x); }
