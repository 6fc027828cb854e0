// N3: SSIM statistics of one MS-SSIM scale, fused forward + backward.
//
// The distortion of the reference's loss is 1 - piq.multi_scale_ssim(x_hat.clamp(0,1), x, data_range=1,
// scale_weights=[.3,.5,.2]) (/root/reference/code/modelv2/model.py:93-102; piq 0.8.0 is third-party, its algorithm is
// restated in losses.py — PARITY UNPINNED).  In eager PyTorch one scale is 5 depthwise 11x11 convolutions + ~15 elementwise
// launches forward and about twice that backward; after GDN and the likelihood were fused this chain was the largest block
// of the training step (ncu launch list r01: 30 % of kernel time).  Here, per scale:
//   forward : one kernel — halo tile of x and y in shared memory, separable Gaussian (11+11 taps instead of 121) of
//             x, y, x^2, y^2, xy, then per pixel  cs = (2 s_xy + c2)/(s_xx + s_yy + c2),  ss = (2 mu_x mu_y + c1)/(mu_x^2+mu_y^2+c1) * cs,
//             per-CTA partial sums of ss and cs (folded in a fixed order by the host -> deterministic), and — when a
//             gradient will be needed — five per-pixel maps from which any upstream (g_ss, g_cs) gradient is a linear mix;
//   backward: one kernel — transposed (full) separable Gaussian of the three mixed maps, then
//             dX = B_mu + 2 X B_xx + Y B_xy.
// Data is tiny (3-channel images), so these kernels are launch/latency sized, not roofline material.
#include "common.cuh"

namespace sic {
namespace {

constexpr int kT = 32;            // output tile edge
constexpr int kK = 11;            // Gaussian taps
constexpr int kHalo = kT + kK - 1;  // 42
constexpr int kThreads = 256;

// normalised 1-D Gaussian, sigma 1.5, 11 taps (float32 values of exp(-d^2/4.5)/sum, as losses._gaussian_window builds them);
// the 2-D window is its outer product
__constant__ float c_g[kK] = {1.028380357e-03f, 7.598758209e-03f, 3.600077331e-02f, 1.093606874e-01f, 2.130055279e-01f, 2.660117149e-01f, 2.130055279e-01f, 1.093606874e-01f, 3.600077331e-02f, 7.598758209e-03f, 1.028380357e-03f};

// ---- forward ---------------------------------------------------------------------------------------------------------
// maps (nullable): [5][planes][Hv][Wv] = M2mu, M2xx, M2xy, l, T   (see ssim_bwd_kernel)
// Layout of an image operand: bands == 0: planar [planes, H, W]; bands == C: channels-last [B, H, W, C] with plane = b * C + c (what the
// fused last synthesis layer writes and the fused first analysis layer reads) - element stride C, no conversion pass either way.
struct PlaneView {
    long base;
    int stride;
};
__device__ __forceinline__ PlaneView plane_view(int plane, int bands, int H, int W) {
    if (bands == 0) return {(long)plane * H * W, 1};
    const int b = plane / bands;
    return {(long)b * H * W * bands + (plane - b * bands), bands};
}

// clamp01: X is clamped to [0, 1] as it is staged (the x_hat.clamp(0, 1) of model.py:98 without its own pass)
__global__ void __launch_bounds__(kThreads) ssim_fwd_kernel(const float *__restrict__ X, const float *__restrict__ Y, int H, int W,
                                                            int x_bands, int y_bands, int clamp01, float c1, float c2,
                                                            float *__restrict__ part_ss, float *__restrict__ part_cs,
                                                            float *__restrict__ maps, long map_stride, float *__restrict__ x_pool,
                                                            float *__restrict__ y_pool) {
    __shared__ float sx[kHalo][kHalo + 1], sy[kHalo][kHalo + 1];
    __shared__ float hb[5][kHalo][kT + 1];
    __shared__ float red[2][kThreads / 32];
    const int plane = blockIdx.z;
    const int Hv = H - (kK - 1), Wv = W - (kK - 1);
    const int oy0 = blockIdx.y * kT, ox0 = blockIdx.x * kT;
    const PlaneView vx = plane_view(plane, x_bands, H, W), vy = plane_view(plane, y_bands, H, W);
    const float *xp = X + vx.base, *yp = Y + vy.base;
    for (int i = threadIdx.x; i < kHalo * kHalo; i += kThreads) {
        int r = i / kHalo, c = i - r * kHalo;
        int gy = oy0 + r, gx = ox0 + c;
        bool in = gy < H && gx < W;
        float xv = in ? __ldg(xp + ((long)gy * W + gx) * vx.stride) : 0.f;
        if (clamp01) xv = fminf(fmaxf(xv, 0.f), 1.f);
        sx[r][c] = xv;
        sy[r][c] = in ? __ldg(yp + ((long)gy * W + gx) * vy.stride) : 0.f;
    }
    __syncthreads();
    if (x_pool != nullptr) {
        // next scale's inputs as a by-product (2x2 average pooling, losses.py / piq): this CTA owns the input rows / columns
        // [32 by, 32 by + 32) of its 42-wide halo tile, the last tile of each axis everything up to the image edge (<= 42: the
        // halo always reaches it).  H and W are even here (the host keeps the torch pooling for odd sizes).
        const int Hp = H >> 1, Wp = W >> 1;
        const int rows = (blockIdx.y == gridDim.y - 1 ? H - oy0 : kT) >> 1, cols = (blockIdx.x == gridDim.x - 1 ? W - ox0 : kT) >> 1;
        for (int i = threadIdx.x; i < rows * cols; i += kThreads) {
            const int pr = i / cols, pc = i - pr * cols;
            const long o = (long)plane * Hp * Wp + (long)((oy0 >> 1) + pr) * Wp + ((ox0 >> 1) + pc);
            x_pool[o] = 0.25f * ((sx[2 * pr][2 * pc] + sx[2 * pr][2 * pc + 1]) + (sx[2 * pr + 1][2 * pc] + sx[2 * pr + 1][2 * pc + 1]));
            y_pool[o] = 0.25f * ((sy[2 * pr][2 * pc] + sy[2 * pr][2 * pc + 1]) + (sy[2 * pr + 1][2 * pc] + sy[2 * pr + 1][2 * pc + 1]));
        }
    }
    for (int i = threadIdx.x; i < kHalo * kT; i += kThreads) {   // horizontal pass
        int r = i / kT, c = i - r * kT;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, a4 = 0.f;
#pragma unroll
        for (int k = 0; k < kK; ++k) {
            float g = c_g[k], xv = sx[r][c + k], yv = sy[r][c + k];
            a0 = fmaf(g, xv, a0);
            a1 = fmaf(g, yv, a1);
            a2 = fmaf(g, xv * xv, a2);
            a3 = fmaf(g, yv * yv, a3);
            a4 = fmaf(g, xv * yv, a4);
        }
        hb[0][r][c] = a0; hb[1][r][c] = a1; hb[2][r][c] = a2; hb[3][r][c] = a3; hb[4][r][c] = a4;
    }
    __syncthreads();
    float acc_ss = 0.f, acc_cs = 0.f;
    for (int i = threadIdx.x; i < kT * kT; i += kThreads) {      // vertical pass + per-pixel SSIM terms
        int r = i / kT, c = i - r * kT;
        int oy = oy0 + r, ox = ox0 + c;
        if (oy < Hv && ox < Wv) {
            float mx = 0.f, my = 0.f, exx = 0.f, eyy = 0.f, exy = 0.f;
#pragma unroll
            for (int k = 0; k < kK; ++k) {
                float g = c_g[k];
                mx = fmaf(g, hb[0][r + k][c], mx);
                my = fmaf(g, hb[1][r + k][c], my);
                exx = fmaf(g, hb[2][r + k][c], exx);
                eyy = fmaf(g, hb[3][r + k][c], eyy);
                exy = fmaf(g, hb[4][r + k][c], exy);
            }
            float sxx = exx - mx * mx, syy = eyy - my * my, sxy = exy - mx * my;
            float A1 = 2.f * mx * my + c1, A2 = mx * mx + my * my + c1;
            float B1 = 2.f * sxy + c2, B2 = sxx + syy + c2;
            float l = A1 / A2, cs = B1 / B2;
            acc_ss += l * cs;
            acc_cs += cs;
            if (maps != nullptr) {
                float iB2 = 1.0f / B2;
                long o = (long)plane * Hv * Wv + (long)oy * Wv + ox;
                maps[o] = 2.f * iB2 * (cs * mx - my);                       // M2mu : d(sum cs)/d mu_x
                maps[map_stride + o] = -cs * iB2;                          // M2xx : d(sum cs)/d E[x^2]
                maps[2 * map_stride + o] = 2.f * iB2;                      // M2xy : d(sum cs)/d E[xy]
                maps[3 * map_stride + o] = l;
                maps[4 * map_stride + o] = cs * 2.f * (my - l * mx) / A2;  // T    : cs * dl/d mu_x
            }
        }
    }
    acc_ss = warp_sum(acc_ss);
    acc_cs = warp_sum(acc_cs);
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = acc_ss; red[1][threadIdx.x >> 5] = acc_cs; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f, c = 0.f;
        for (int w = 0; w < kThreads / 32; ++w) { s += red[0][w]; c += red[1][w]; }
        long tiles = (long)gridDim.x * gridDim.y;
        long t = (long)plane * tiles + (long)blockIdx.y * gridDim.x + blockIdx.x;
        part_ss[t] = s;
        part_cs[t] = c;
    }
}

// ---- backward --------------------------------------------------------------------------------------------------------
// F = g_ss[plane] * mean(ss) + g_cs[plane] * mean(cs).  With k = g_ss*l + g_cs (per pixel):
//   M_mu = g_ss*T + k*M2mu ,  M_xx = k*M2xx ,  M_xy = k*M2xy          (all scaled by 1/n_valid)
//   dF/dX(q) = sum_p G(p..q) M_mu(p) + 2 X(q) sum_p G M_xx(p) + Y(q) sum_p G M_xy(p)      (transposed window: q = p + tap)
// g_scale (nullable): device scalar multiplying g_ss and g_cs (the gradient arriving at the MS-SSIM value, when g_ss / g_cs hold the
// derivatives of that value w.r.t. this scale's means: sic_msssim_combine).  dX has the layout of X; with clamp01 it is zero where
// X lies outside [0, 1] (the backward of the clamp folded into the forward's staging).
__global__ void __launch_bounds__(kThreads) ssim_bwd_kernel(const float *__restrict__ X, const float *__restrict__ Y,
                                                            const float *__restrict__ maps, long map_stride,
                                                            const float *__restrict__ g_ss, const float *__restrict__ g_cs,
                                                            const float *__restrict__ g_scale, int H, int W, int x_bands, int y_bands,
                                                            int clamp01, const float *__restrict__ g_pool, float *__restrict__ dX) {
    __shared__ float sm[3][kHalo][kHalo + 1];
    __shared__ float hb[3][kHalo][kT + 1];
    const int plane = blockIdx.z;
    const int Hv = H - (kK - 1), Wv = W - (kK - 1);
    const int qy0 = blockIdx.y * kT, qx0 = blockIdx.x * kT;
    const float inv_n = (g_scale ? __ldg(g_scale) : 1.f) / ((float)Hv * (float)Wv);
    const float gs = (g_ss ? __ldg(g_ss + plane) : 0.f) * inv_n, gc = (g_cs ? __ldg(g_cs + plane) : 0.f) * inv_n;
    const PlaneView vx = plane_view(plane, x_bands, H, W), vy = plane_view(plane, y_bands, H, W);
    const float *mp = maps + (long)plane * Hv * Wv;
    for (int i = threadIdx.x; i < kHalo * kHalo; i += kThreads) {
        int r = i / kHalo, c = i - r * kHalo;
        int py = qy0 + r - (kK - 1), px = qx0 + c - (kK - 1);     // valid-map coordinate that reaches q with tap (kK-1 - ...)
        float m0 = 0.f, m1 = 0.f, m2 = 0.f;
        if (py >= 0 && py < Hv && px >= 0 && px < Wv) {
            long o = (long)py * Wv + px;
            float k = fmaf(gs, __ldg(mp + 3 * map_stride + o), gc);
            m0 = fmaf(gs, __ldg(mp + 4 * map_stride + o), k * __ldg(mp + o));
            m1 = k * __ldg(mp + map_stride + o);
            m2 = k * __ldg(mp + 2 * map_stride + o);
        }
        sm[0][r][c] = m0; sm[1][r][c] = m1; sm[2][r][c] = m2;
    }
    __syncthreads();
    // q = p + t  (t = tap index 0..10)  =>  p = q - t: halo column (c + 10 - t)
    for (int i = threadIdx.x; i < kHalo * kT; i += kThreads) {
        int r = i / kT, c = i - r * kT;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
        for (int t = 0; t < kK; ++t) {
            float g = c_g[t];
            int cc = c + (kK - 1) - t;
            a0 = fmaf(g, sm[0][r][cc], a0);
            a1 = fmaf(g, sm[1][r][cc], a1);
            a2 = fmaf(g, sm[2][r][cc], a2);
        }
        hb[0][r][c] = a0; hb[1][r][c] = a1; hb[2][r][c] = a2;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kT * kT; i += kThreads) {
        int r = i / kT, c = i - r * kT;
        int qy = qy0 + r, qx = qx0 + c;
        if (qy < H && qx < W) {
            float b0 = 0.f, b1 = 0.f, b2 = 0.f;
#pragma unroll
            for (int t = 0; t < kK; ++t) {
                float g = c_g[t];
                int rr = r + (kK - 1) - t;
                b0 = fmaf(g, hb[0][rr][c], b0);
                b1 = fmaf(g, hb[1][rr][c], b1);
                b2 = fmaf(g, hb[2][rr][c], b2);
            }
            const long q = (long)qy * W + qx, o = vx.base + q * vx.stride;
            const float xraw = __ldg(X + o);
            const float xv = clamp01 ? fminf(fmaxf(xraw, 0.f), 1.f) : xraw;
            float d = b0 + 2.f * xv * b1 + __ldg(Y + vy.base + q * vy.stride) * b2;
            // the gradient that arrives through the pooled copy of X (the coarser scales): avg_pool2d backward, 1/4 to each of the four
            if (g_pool != nullptr) d = fmaf(0.25f, __ldg(g_pool + (long)plane * (H >> 1) * (W >> 1) + (long)(qy >> 1) * (W >> 1) + (qx >> 1)), d);
            if (clamp01 && !(xraw >= 0.f && xraw <= 1.f)) d = 0.f;                  // clamp backward: the gradient passes inside [0, 1] only
            dX[o] = d;
        }
    }
}

// ---- combination of the scales -----------------------------------------------------------------------------------------
// piq.multi_scale_ssim's tail (restated in losses.py): per (image, band) plane  prod_{l < L-1} relu(mean cs_l)^w_l * relu(mean ss_{L-1})^w_{L-1},
// weights normalised by their sum, then the mean over bands and images - and, for the backward, the derivative of that value
// with respect to every per-plane mean (what ssim_bwd_kernel takes as g_ss / g_cs).  One CTA: the data is L x 2 x planes x tiles floats.
// Replaces ~13 tiny torch launches forward and ~20 backward per training step (profiles/r02ap_ncu_launches_bench_step.txt).
constexpr int kMaxLevels = 8;
struct CombineArgs {
    int levels, planes, normalize;   // normalize: weights divided by their sum (piq does that to user-supplied weights, not to its defaults)
    int tiles[kMaxLevels];
    long off[kMaxLevels];        // level l: part_ss at part + off[l], part_cs at part + off[l] + planes * tiles[l]
    float inv_n[kMaxLevels];     // 1 / ((H_l - 10)(W_l - 10))
};

__global__ void __launch_bounds__(kThreads) msssim_combine_kernel(const float *__restrict__ part, const float *__restrict__ weights,
                                                                  CombineArgs a, float *__restrict__ out, float *__restrict__ coef) {
    __shared__ float red[kThreads / 32];
    const int L = a.levels;
    float w[kMaxLevels], wsum = 0.f;
#pragma unroll
    for (int l = 0; l < kMaxLevels; ++l) {
        w[l] = l < L ? __ldg(weights + l) : 0.f;
        wsum += w[l];
    }
    if (!a.normalize) wsum = 1.f;
    const float inv_planes = 1.0f / (float)a.planes;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float acc = 0.f;                                             // lane 0 of each warp: sum of its planes' values
    // one WARP per plane: the lanes fetch the tile partials side by side (a thread per plane walked up to 64 of them one after the
    // other: 11 us of pure load latency), the xor tree gives every lane the same fixed-order sum
    for (int plane = warp; plane < a.planes; plane += kThreads / 32) {
        float t[kMaxLevels], pw[kMaxLevels];
        float prod = 1.f;
#pragma unroll
        for (int l = 0; l < kMaxLevels; ++l) {
            t[l] = 0.f;
            pw[l] = 1.f;
            if (l < L) {
                // the last scale contributes its SSIM mean, the others their contrast-structure mean
                const float *src = part + a.off[l] + (l == L - 1 ? 0 : (long)a.planes * a.tiles[l]) + (long)plane * a.tiles[l];
                float sum = 0.f;
                for (int k = lane; k < a.tiles[l]; k += 32) sum += __ldg(src + k);
                sum = warp_sum(sum);
                t[l] = fmaxf(sum * a.inv_n[l], 0.f);
                pw[l] = powf(t[l], w[l] / wsum);
                prod *= pw[l];
            }
        }
        if (lane == 0) acc += prod;
        if (coef != nullptr && lane == 0) {
#pragma unroll
            for (int l = 0; l < kMaxLevels; ++l) {
                if (l < L) {
                    float others = 1.f;
#pragma unroll
                    for (int k = 0; k < kMaxLevels; ++k)
                        if (k < L && k != l) others *= pw[k];
                    const float wl = w[l] / wsum;
                    const float d = t[l] > 0.f ? inv_planes * others * wl * powf(t[l], wl - 1.f) : 0.f;      // relu: nothing passes at <= 0
                    coef[((long)l * 2 + 0) * a.planes + plane] = l == L - 1 ? d : 0.f;                        // d value / d mean ss_l
                    coef[((long)l * 2 + 1) * a.planes + plane] = l == L - 1 ? 0.f : d;                        // d value / d mean cs_l
                }
            }
        }
    }
    if (lane == 0) red[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
#pragma unroll
        for (int wq = 0; wq < kThreads / 32; ++wq) s += red[wq];
        *out = s * inv_planes;
    }
}

SIC_REGISTER_KERNEL("ssim_fwd_kernel", ssim_fwd_kernel);
SIC_REGISTER_KERNEL("msssim_combine_kernel", msssim_combine_kernel);
SIC_REGISTER_KERNEL("ssim_bwd_kernel", ssim_bwd_kernel);
}  // namespace
}  // namespace sic

using namespace sic;

extern "C" long sic_ssim_tiles(int H, int W) {
    if (H < 11 || W < 11) return 0;
    return (long)((H - 10 + kT - 1) / kT) * ((W - 10 + kT - 1) / kT);
}

extern "C" int sic_ssim_fwd_ex(const float *X, const float *Y, int planes, int H, int W, int x_bands, int y_bands, int clamp01, float c1,
                               float c2, float *part_ss, float *part_cs, float *maps, float *x_pool, float *y_pool, void *stream) {
    SIC_CHECK_ARG(planes > 0 && H >= 11 && W >= 11, "sic_ssim_fwd: needs planes > 0 and images of at least 11x11 (got %d, %dx%d)", planes, H, W);
    SIC_CHECK_ARG(X && Y && part_ss && part_cs, "sic_ssim_fwd: null pointer");
    SIC_CHECK_ARG(planes <= 65535, "sic_ssim_fwd: more than 65535 (batch x channel) planes");
    SIC_CHECK_ARG(x_bands >= 0 && y_bands >= 0 && (x_bands == 0 || planes % x_bands == 0) && (y_bands == 0 || planes % y_bands == 0),
                  "sic_ssim_fwd: bands (%d, %d) must be 0 (planar) or divide planes = %d (channels-last)", x_bands, y_bands, planes);
    SIC_CHECK_ARG((x_pool == nullptr) == (y_pool == nullptr), "sic_ssim_fwd_pool: x_pool and y_pool go together");
    SIC_CHECK_ARG(x_pool == nullptr || ((H | W) & 1) == 0, "sic_ssim_fwd_pool: pooled outputs need even H and W (got %dx%d)", H, W);
    cudaStream_t st = (cudaStream_t)stream;
    const int Hv = H - 10, Wv = W - 10;
    dim3 grid((Wv + kT - 1) / kT, (Hv + kT - 1) / kT, planes);
    ssim_fwd_kernel<<<grid, kThreads, 0, st>>>(X, Y, H, W, x_bands, y_bands, clamp01, c1, c2, part_ss, part_cs, maps,
                                               (long)planes * Hv * Wv, x_pool, y_pool);
    SIC_CHECK_LAUNCH("sic_ssim_fwd");
    return 0;
}

extern "C" int sic_ssim_fwd_pool(const float *X, const float *Y, int planes, int H, int W, float c1, float c2, float *part_ss,
                                 float *part_cs, float *maps, float *x_pool, float *y_pool, void *stream) {
    return sic_ssim_fwd_ex(X, Y, planes, H, W, 0, 0, 0, c1, c2, part_ss, part_cs, maps, x_pool, y_pool, stream);
}

extern "C" int sic_ssim_fwd(const float *X, const float *Y, int planes, int H, int W, float c1, float c2, float *part_ss,
                            float *part_cs, float *maps, void *stream) {
    return sic_ssim_fwd_pool(X, Y, planes, H, W, c1, c2, part_ss, part_cs, maps, nullptr, nullptr, stream);
}

extern "C" int sic_ssim_bwd_ex(const float *X, const float *Y, const float *maps, const float *g_ss, const float *g_cs,
                               const float *g_scale, const float *g_pool, int planes, int H, int W, int x_bands, int y_bands, int clamp01,
                               float *dX, void *stream) {
    SIC_CHECK_ARG(planes > 0 && H >= 11 && W >= 11, "sic_ssim_bwd: bad extents");
    SIC_CHECK_ARG(X && Y && maps && dX, "sic_ssim_bwd: null pointer");
    SIC_CHECK_ARG(planes <= 65535, "sic_ssim_bwd: more than 65535 (batch x channel) planes");
    SIC_CHECK_ARG(x_bands >= 0 && y_bands >= 0 && (x_bands == 0 || planes % x_bands == 0) && (y_bands == 0 || planes % y_bands == 0),
                  "sic_ssim_bwd: bands (%d, %d) must be 0 (planar) or divide planes = %d (channels-last)", x_bands, y_bands, planes);
    SIC_CHECK_ARG(g_pool == nullptr || ((H | W) & 1) == 0, "sic_ssim_bwd_pool: a pooled gradient needs even H and W (got %dx%d)", H, W);
    cudaStream_t st = (cudaStream_t)stream;
    const int Hv = H - 10, Wv = W - 10;
    dim3 grid((W + kT - 1) / kT, (H + kT - 1) / kT, planes);
    ssim_bwd_kernel<<<grid, kThreads, 0, st>>>(X, Y, maps, (long)planes * Hv * Wv, g_ss, g_cs, g_scale, H, W, x_bands, y_bands, clamp01,
                                               g_pool, dX);
    SIC_CHECK_LAUNCH("sic_ssim_bwd");
    return 0;
}

extern "C" int sic_ssim_bwd_pool(const float *X, const float *Y, const float *maps, const float *g_ss, const float *g_cs,
                                 const float *g_pool, int planes, int H, int W, float *dX, void *stream) {
    return sic_ssim_bwd_ex(X, Y, maps, g_ss, g_cs, nullptr, g_pool, planes, H, W, 0, 0, 0, dX, stream);
}

extern "C" int sic_ssim_bwd(const float *X, const float *Y, const float *maps, const float *g_ss, const float *g_cs, int planes,
                            int H, int W, float *dX, void *stream) {
    return sic_ssim_bwd_pool(X, Y, maps, g_ss, g_cs, nullptr, planes, H, W, dX, stream);
}

extern "C" int sic_msssim_combine(const float *part, const long *offsets, const int *tiles, const long *n_valid, int levels, int planes,
                                  const float *weights, int normalize, float *out, float *coef, void *stream) {
    SIC_CHECK_ARG(levels >= 1 && levels <= kMaxLevels, "sic_msssim_combine: %d scales (1..%d supported)", levels, kMaxLevels);
    SIC_CHECK_ARG(planes > 0, "sic_msssim_combine: no planes");
    SIC_CHECK_ARG(part && offsets && tiles && n_valid && weights && out, "sic_msssim_combine: null pointer");
    CombineArgs a;
    a.levels = levels;
    a.planes = planes;
    a.normalize = normalize != 0;
    for (int l = 0; l < kMaxLevels; ++l) {
        a.tiles[l] = l < levels ? tiles[l] : 0;
        a.off[l] = l < levels ? offsets[l] : 0;
        a.inv_n[l] = 1.f;
        if (l < levels) {
            SIC_CHECK_ARG(tiles[l] > 0 && n_valid[l] > 0 && offsets[l] >= 0, "sic_msssim_combine: bad geometry of scale %d", l);
            a.inv_n[l] = 1.0f / (float)n_valid[l];
        }
    }
    msssim_combine_kernel<<<1, kThreads, 0, (cudaStream_t)stream>>>(part, weights, a, out, coef);
    SIC_CHECK_LAUNCH("sic_msssim_combine");
    return 0;
}
