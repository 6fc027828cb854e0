// Bias add (+ ReLU) after a cuDNN convolution and its adjoint, for position-major ("channels-last") activations [P, C].
//
// PyTorch runs a biased convolution as cudnn_convolution (no bias) -> add_(bias) [-> relu_], and its backward as threshold_backward ->
// convolution_backward + a separate sum over (batch, height, width) for d(bias) - a reduce_kernel that takes 11 - 18 us on these small
// tensors (profiles/r02ax_ncu_launches_bench_step.txt).  The convolutions followed by a GDN / IGDN already hand their bias to the GDN
// kernel (gdn.cu); this file covers the rest - the hyper analysis / synthesis layers (layers.py:104-139, conv -> ReLU) and the last
// analysis convolution (layers.py:73) - and the d(bias) of the last synthesis layer (deconv_rgb.cu):
//   forward   t[p, c] = act(t[p, c] + bias[c]) in place: the single fp32 add PyTorch does, ReLU as `v < 0 ? 0 : v` (keeps -0 and NaN like
//             clamp_min) - bit-identical values
//   backward  dt = g * (y > 0) (ReLU only), d(bias)[c] = sum_p dt[p, c]: one pass, per-CTA partial sums in a fixed order, then a fold.
#include "common.cuh"

namespace sic {
namespace {

constexpr int kBaRowsPerCta = 64;
constexpr int kBaFoldLanes = 8;

__global__ void __launch_bounds__(256) bias_act_fwd_kernel(float *__restrict__ t, const float *__restrict__ bias, long n, int C, int relu) {
    const long stride = (long)gridDim.x * 256;
    if ((C & 3) == 0 && (((uintptr_t)t) & 15) == 0) {
        float4 *t4 = reinterpret_cast<float4 *>(t);
        const long n4 = n >> 2;
        for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < n4; i += stride) {
            const int c = (int)((i << 2) % C);
            const float4 b = make_float4(__ldg(bias + c), __ldg(bias + c + 1), __ldg(bias + c + 2), __ldg(bias + c + 3));   // bias may sit anywhere in a flat parameter buffer
            float4 v = t4[i];
            v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
            if (relu) {
                v.x = v.x < 0.f ? 0.f : v.x; v.y = v.y < 0.f ? 0.f : v.y;
                v.z = v.z < 0.f ? 0.f : v.z; v.w = v.w < 0.f ? 0.f : v.w;
            }
            t4[i] = v;
        }
    } else {
        for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < n; i += stride) {
            float v = t[i] + __ldg(bias + (int)(i % C));
            if (relu) v = v < 0.f ? 0.f : v;
            t[i] = v;
        }
    }
}

// blockDim.x = T, a multiple of C: a thread keeps one channel for life.  part[cta][c].
__global__ void __launch_bounds__(1024) bias_act_bwd_kernel(const float *__restrict__ g, const float *__restrict__ y, float *__restrict__ dt,
                                                            long P, int C, int rows_per_cta, float *__restrict__ part) {
    extern __shared__ float red[];
    const int T = blockDim.x;
    const long r0 = (long)blockIdx.x * rows_per_cta;
    const long r1 = r0 + rows_per_cta < P ? r0 + rows_per_cta : P;
    const long e0 = r0 * C, e1 = r1 * C;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    long i = e0 + threadIdx.x;
    for (; i + 3L * T < e1; i += 4L * T) {
        float v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = __ldg(g + i + (long)u * T);
        if (y != nullptr) {
            float m[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) m[u] = __ldg(y + i + (long)u * T);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                v[u] = m[u] > 0.f ? v[u] : 0.f;                    // threshold_backward: the gradient passes where the output is > 0
                dt[i + (long)u * T] = v[u];
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) acc[u] += v[u];
    }
    for (; i < e1; i += T) {
        float v = __ldg(g + i);
        if (y != nullptr) {
            v = __ldg(y + i) > 0.f ? v : 0.f;
            dt[i] = v;
        }
        acc[0] += v;
    }
    red[threadIdx.x] = (acc[0] + acc[1]) + (acc[2] + acc[3]);
    __syncthreads();
    if (threadIdx.x < C) {
        float s = 0.f;
        for (int j = threadIdx.x; j < T; j += C) s += red[j];       // the T / C threads of this channel, in thread order
        part[(size_t)blockIdx.x * C + threadIdx.x] = s;
    }
}

// d(bias)[c] = sum over the CTAs' partials: eight lanes per channel, contiguous eighths, fixed xor tree (binary64)
__global__ void __launch_bounds__(256) bias_grad_fold_kernel(const float *__restrict__ part, int n_part, int C, float *__restrict__ dbias) {
    const int t = blockIdx.x * 256 + threadIdx.x;
    const int c = t / kBaFoldLanes, sub = t % kBaFoldLanes;
    const int per = (n_part + kBaFoldLanes - 1) / kBaFoldLanes;
    const int p0 = sub * per, p1 = min(p0 + per, n_part);
    double s = 0.0;
    if (c < C) {
        int p = p0;
        for (; p + 8 <= p1; p += 8) {
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = __ldg(part + (size_t)(p + u) * C + c);
#pragma unroll
            for (int u = 0; u < 8; ++u) s += (double)v[u];
        }
        for (; p < p1; ++p) s += (double)__ldg(part + (size_t)p * C + c);
    }
#pragma unroll
    for (int o = 1; o < kBaFoldLanes; o <<= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (c < C && sub == 0) dbias[c] = (float)s;
}

SIC_REGISTER_KERNEL("bias_act_fwd_kernel", bias_act_fwd_kernel);
SIC_REGISTER_KERNEL("bias_act_bwd_kernel", bias_act_bwd_kernel);
SIC_REGISTER_KERNEL("bias_grad_fold_kernel", bias_grad_fold_kernel);

inline int ba_grid(long P) {
    const long want = (P + kBaRowsPerCta - 1) / kBaRowsPerCta;
    const long cap = 2L * sm_count();
    return (int)(want < cap ? want : cap);
}

}  // namespace
}  // namespace sic

using namespace sic;

extern "C" int sic_bias_act_fwd(float *t, const float *bias, long P, int C, int relu, void *stream) {
    SIC_CHECK_ARG(P > 0 && C > 0, "sic_bias_act_fwd: empty shape P=%ld C=%d", P, C);
    SIC_CHECK_ARG(t && bias, "sic_bias_act_fwd: null pointer");
    const long n = P * C;
    const long want = ((n >> 2) + 255) / 256 + 1;
    const long cap = 8L * sm_count();
    bias_act_fwd_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, (cudaStream_t)stream>>>(t, bias, n, C, relu != 0);
    SIC_CHECK_LAUNCH("sic_bias_act_fwd");
    return 0;
}

extern "C" size_t sic_bias_grad_workspace_bytes(long P, int C) {
    return (P > 0 && C > 0) ? (size_t)ba_grid(P) * C * sizeof(float) : 0;
}

extern "C" int sic_bias_act_bwd(const float *g, const float *y, long P, int C, float *dt, float *dbias, void *workspace,
                                size_t workspace_bytes, void *stream) {
    SIC_CHECK_ARG(P > 0 && C > 0 && C <= 1024, "sic_bias_act_bwd: needs P > 0 and 0 < C <= 1024 (got P=%ld C=%d)", P, C);
    SIC_CHECK_ARG(g && dbias && workspace, "sic_bias_act_bwd: null pointer");
    SIC_CHECK_ARG((y == nullptr) == (dt == nullptr), "sic_bias_act_bwd: y (the ReLU output) and dt go together");
    SIC_CHECK_ARG(workspace_bytes >= sic_bias_grad_workspace_bytes(P, C) && (((uintptr_t)workspace) & 3) == 0,
                  "sic_bias_act_bwd: workspace of %zu B, needs %zu B", workspace_bytes, sic_bias_grad_workspace_bytes(P, C));
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = ba_grid(P);
    const int rows = (int)((P + grid - 1) / grid);
    const int T = C >= 256 ? C : (256 / C) * C;                       // a multiple of C: each thread owns one channel
    bias_act_bwd_kernel<<<grid, T, T * sizeof(float), st>>>(g, y, dt, P, C, rows, (float *)workspace);
    SIC_CHECK_LAUNCH("sic_bias_act_bwd");
    bias_grad_fold_kernel<<<(C * kBaFoldLanes + 255) / 256, 256, 0, st>>>((const float *)workspace, grid, C, dbias);
    SIC_CHECK_LAUNCH("sic_bias_act_bwd (fold)");
    return 0;
}
