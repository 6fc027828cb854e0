// Last synthesis layer, ConvTranspose2d(N, 3, 5, stride 2, padding 2, output_padding 1) (/root/reference/code/modelv2/layers.py:96-98,
// `deconv(N, 3)`): the scatter/gather halves of its GEMM formulation.
//
// cuDNN serves this layer (3 output bands: not a 16-byte channels-last vector) with its legacy non-tensor-core engines - forward 238 us,
// backward 438 us per cfg2 step (profiles/r02i_kernel_bench_deconv.json) for 5 GFLOP each.  As a GEMM it is tiny and HBM-bound:
//   forward   D[p, m] = sum_ci a[p, ci] Wm[m, ci]            m = (kh * 5 + kw) * 3 + co  (75 of 80 columns), p = (b, iy, ix)
//             = a 1x1 convolution N -> 80 on the channels-last activations: cuDNN's sm_100 tensor-op kernels (and their dgrad / wgrad
//             in the backward, through autograd)
//             x_hat[b, 2 iy - 2 + kh, 2 ix - 2 + kw, co] += D[p, m]                                                    col2im (here)
//   backward  dD[p, m] = g[b, 2 iy - 2 + kh, 2 ix - 2 + kw, co]                                                        im2col (here)
// col2im is written as a GATHER: the thread of input position (b, iy, ix) owns the 2 x 2 output pixels (2 iy + dy, 2 ix + dx) and
// sums the taps that land on them in a fixed order, so there are no atomics, every D element is read exactly once, and the result is
// deterministic.  Output pixel (2 iy + dy, .) receives kernel rows kh == dy (mod 2) from input rows iy + (2 + dy - kh) / 2.
// D is position-major ([P][80], the channels-last output of the 1x1 convolution), so a CTA first stages the D rows of its 8 x 32
// block of positions plus a one-position halo in shared memory with coalesced 128-bit loads (row stride 81 floats: the gather's
// column reads then hit 32 different banks), and gathers from there.
#include "common.cuh"

namespace sic {
namespace {

constexpr int kDcM = 75;        // 5 x 5 taps x 3 bands
constexpr int kDcMP = 80;       // row length of D: 75 padded to a multiple of 4 floats (16-byte rows)
constexpr int kDcTH = 8, kDcTW = 32;                 // positions per CTA: 8 rows x 32 columns, one thread each
constexpr int kDcThreads = kDcTH * kDcTW;
constexpr int kDcSH = kDcTH + 2, kDcSW = kDcTW + 2;  // staged block incl. halo
constexpr int kDcStride = 81;                        // shared-memory row stride in floats (odd: conflict-free column reads)
constexpr size_t kDcSmem = (size_t)kDcSH * kDcSW * kDcStride * sizeof(float);   // 110 KB

__global__ void __launch_bounds__(kDcThreads, 2) deconv_rgb_col2im_kernel(const float *__restrict__ D, const float *__restrict__ bias, int B, int H,
                                                                     int W, int tiles_x, int tiles_y, float *__restrict__ out) {
    extern __shared__ float sD[];
    const int tile = blockIdx.x;
    const int b = tile / (tiles_x * tiles_y), tr = tile - b * tiles_x * tiles_y;
    const int ty = tr / tiles_x, tx = tr - ty * tiles_x;
    const int y0 = ty * kDcTH, x0 = tx * kDcTW;
    // stage rows (y0-1 .. y0+TH, x0-1 .. x0+TW) of D: 20 float4 per position, positions outside the image are never read.
    // Eight loads per thread are issued before the first store (r02am: with one load -> four stores per loop turn the kernel had 8 KB
    // of loads in flight per SM and ran at 1.3 TB/s).
    constexpr int kStageU = 8;
    constexpr int kStageN = kDcSH * kDcSW * (kDcMP / 4);
    for (int i0 = threadIdx.x; i0 < kStageN; i0 += kDcThreads * kStageU) {
        float4 v[kStageU];
        int dst[kStageU];
#pragma unroll
        for (int u = 0; u < kStageU; ++u) {
            const int i = i0 + u * kDcThreads;
            dst[u] = -1;
            if (i < kStageN) {
                const int pos = i / (kDcMP / 4), q = i - pos * (kDcMP / 4);
                const int sy = pos / kDcSW, sx = pos - sy * kDcSW;
                const int gy = y0 - 1 + sy, gx = x0 - 1 + sx;
                if ((unsigned)gy < (unsigned)H && (unsigned)gx < (unsigned)W) {
                    v[u] = ldg_stream(reinterpret_cast<const float4 *>(D + (((size_t)b * H + gy) * W + gx) * kDcMP) + q);
                    dst[u] = pos * kDcStride + 4 * q;
                }
            }
        }
#pragma unroll
        for (int u = 0; u < kStageU; ++u) {
            if (dst[u] >= 0) {
                float *d = sD + dst[u];
                d[0] = v[u].x; d[1] = v[u].y; d[2] = v[u].z; d[3] = v[u].w;
            }
        }
    }
    __syncthreads();
    const int ly = threadIdx.x / kDcTW, lx = threadIdx.x - ly * kDcTW;
    const int iy = y0 + ly, ix = x0 + lx;
    if (iy >= H || ix >= W) return;
    const float b0 = bias ? __ldg(bias) : 0.f, b1 = bias ? __ldg(bias + 1) : 0.f, b2 = bias ? __ldg(bias + 2) : 0.f;
    float acc[2][2][3];
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) { acc[dy][dx][0] = b0; acc[dy][dx][1] = b1; acc[dy][dx][2] = b2; }
#pragma unroll
    for (int kh = 0; kh < 5; ++kh) {
        const int dy = kh & 1, oy = (2 + dy - kh) / 2;               // source input row of this kernel row, relative to iy
        if ((unsigned)(iy + oy) >= (unsigned)H) continue;
#pragma unroll
        for (int kw = 0; kw < 5; ++kw) {
            const int dx = kw & 1, ox = (2 + dx - kw) / 2;
            if ((unsigned)(ix + ox) >= (unsigned)W) continue;
            const float *src = sD + ((ly + 1 + oy) * kDcSW + (lx + 1 + ox)) * kDcStride + (kh * 5 + kw) * 3;
            acc[dy][dx][0] += src[0];
            acc[dy][dx][1] += src[1];
            acc[dy][dx][2] += src[2];
        }
    }
    const int OW = 2 * W;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy) {   // 2 pixels x 3 bands = 6 consecutive floats, 8-byte aligned
        float2 *o = reinterpret_cast<float2 *>(out + (((size_t)b * 2 * H + 2 * iy + dy) * OW + 2 * ix) * 3);
        __stcs(o, make_float2(acc[dy][0][0], acc[dy][0][1]));
        __stcs(o + 1, make_float2(acc[dy][0][2], acc[dy][1][0]));
        __stcs(o + 2, make_float2(acc[dy][1][1], acc[dy][1][2]));
    }
}

// the adjoint gather: one thread per position writes its row of dD (75 values + 5 zeros)
__global__ void __launch_bounds__(256) deconv_rgb_im2col_kernel(const float *__restrict__ g, int B, int H, int W, float *__restrict__ dD) {
    const long P = (long)B * H * W;
    const long p = (long)blockIdx.x * 256 + threadIdx.x;
    if (p >= P) return;
    const int hw = H * W;
    const int b = (int)(p / hw), rem = (int)(p - (long)b * hw);
    const int iy = rem / W, ix = rem - iy * W;
    const int OH = 2 * H, OW = 2 * W;
    float row[kDcMP];
#pragma unroll
    for (int kh = 0; kh < 5; ++kh) {
        const int oy = 2 * iy - 2 + kh;
        const bool rok = (unsigned)oy < (unsigned)OH;
#pragma unroll
        for (int kw = 0; kw < 5; ++kw) {
            const int ox = 2 * ix - 2 + kw;
            float v0 = 0.f, v1 = 0.f, v2 = 0.f;
            if (rok && (unsigned)ox < (unsigned)OW) {
                const float *s = g + (((size_t)b * OH + oy) * OW + ox) * 3;
                v0 = __ldg(s); v1 = __ldg(s + 1); v2 = __ldg(s + 2);
            }
            row[(kh * 5 + kw) * 3] = v0; row[(kh * 5 + kw) * 3 + 1] = v1; row[(kh * 5 + kw) * 3 + 2] = v2;
        }
    }
#pragma unroll
    for (int k = kDcM; k < kDcMP; ++k) row[k] = 0.f;
    float4 *d = reinterpret_cast<float4 *>(dD + (size_t)p * kDcMP);
#pragma unroll
    for (int q = 0; q < kDcMP / 4; ++q) stg_stream(d + q, make_float4(row[4 * q], row[4 * q + 1], row[4 * q + 2], row[4 * q + 3]));
}

SIC_REGISTER_KERNEL("deconv_rgb_col2im_kernel", deconv_rgb_col2im_kernel);
SIC_REGISTER_KERNEL("deconv_rgb_im2col_kernel", deconv_rgb_im2col_kernel);

}  // namespace
}  // namespace sic

using namespace sic;

extern "C" int sic_deconv_rgb_col2im(const float *D, const float *bias, int B, int H, int W, float *out, void *stream) {
    SIC_CHECK_ARG(B > 0 && H > 0 && W > 0, "sic_deconv_rgb_col2im: empty shape B=%d H=%d W=%d", B, H, W);
    SIC_CHECK_ARG(D && out, "sic_deconv_rgb_col2im: null pointer");
    SIC_CHECK_ARG(((uintptr_t)out & 7) == 0 && ((uintptr_t)D & 15) == 0, "sic_deconv_rgb_col2im: D must be 16-byte, out 8-byte aligned");
    const int tiles_x = (W + kDcTW - 1) / kDcTW, tiles_y = (H + kDcTH - 1) / kDcTH;
    const long tiles = (long)B * tiles_x * tiles_y;
    SIC_CHECK_ARG(tiles < (1L << 31), "sic_deconv_rgb_col2im: too many tiles");
    cudaError_t e = cudaFuncSetAttribute(deconv_rgb_col2im_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDcSmem);
    if (e != cudaSuccess) {
        set_error("sic_deconv_rgb_col2im: cannot reserve %zu B of shared memory: %s", kDcSmem, cudaGetErrorString(e));
        return (int)e;
    }
    deconv_rgb_col2im_kernel<<<(unsigned)tiles, kDcThreads, kDcSmem, (cudaStream_t)stream>>>(D, bias, B, H, W, tiles_x, tiles_y, out);
    SIC_CHECK_LAUNCH("sic_deconv_rgb_col2im");
    return 0;
}

extern "C" int sic_deconv_rgb_im2col(const float *grad_out, int B, int H, int W, float *dD, void *stream) {
    SIC_CHECK_ARG(B > 0 && H > 0 && W > 0, "sic_deconv_rgb_im2col: empty shape B=%d H=%d W=%d", B, H, W);
    SIC_CHECK_ARG(grad_out && dD, "sic_deconv_rgb_im2col: null pointer");
    SIC_CHECK_ARG(((uintptr_t)dD & 15) == 0, "sic_deconv_rgb_im2col: dD must be 16-byte aligned");
    const long P = (long)B * H * W;
    SIC_CHECK_ARG((P + 255) / 256 < (1L << 31), "sic_deconv_rgb_im2col: too many positions");
    deconv_rgb_im2col_kernel<<<(unsigned)((P + 255) / 256), 256, 0, (cudaStream_t)stream>>>(grad_out, B, H, W, dD);
    SIC_CHECK_LAUNCH("sic_deconv_rgb_im2col");
    return 0;
}
