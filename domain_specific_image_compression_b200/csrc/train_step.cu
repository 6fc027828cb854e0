// Tail of the data-parallel training step on the FLAT parameter / gradient buffers of trainer.py: global-norm
// clipping + Adam, /root/reference/code/modelv2/train.py:182-183 (optim.Adam) and :200-203 (clip_grad_norm_ at OPTIM.grad_clip,
// then the optimizer step).
//
// Round 2 ran this as torch ops on the flat buffers: vector_norm, four scalar ops for the clip coefficient, an in-place multiply of
// the 26-60 MB gradient and the library's multi-tensor Adam (90 us for one 6.4 M-element tensor: it was built for lists of tensors) -
// 125 us per cfg2 step (profiles/r02ap_ncu_launches_bench_step.txt).  Here: two launches that move the minimum,
//   grad_sumsq_kernel   reads g once, per-CTA partial sums of squares in double (fixed order: deterministic), bumps the step counter
//   adam_clip_kernel    folds the partials (every CTA, same order), derives the clip coefficient, then one pass over p, g, m, v:
//                       16 B read + 12 B written per parameter, nothing else.
// Everything the update needs lives on the device (step counter, norm), so the step stays CUDA-graph capturable.
//
// Update rule (Kingma & Ba 2015, in the operation order of torch.optim.Adam's fused CUDA implementation so the two agree to the last
// bits): g' = (g / world) * coef (+ wd * p);  m = b1 m + (1 - b1) g';  v = b2 v + (1 - b2) g'^2;
//        p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps),   coef = min(1, clip / (|g / world| + 1e-6)).
#include "common.cuh"

namespace sic {
namespace {

constexpr int kOptThreads = 256;
constexpr int kOptU = 4;           // 128-bit loads in flight per thread and tensor

__global__ void __launch_bounds__(kOptThreads) grad_sumsq_kernel(const float *__restrict__ g, long n, double *__restrict__ part,
                                                                 float *__restrict__ step) {
    const long n4 = n >> 2;
    const float4 *g4 = reinterpret_cast<const float4 *>(g);
    const long stride = (long)gridDim.x * kOptThreads;
    double acc = 0.0;
    long i = (long)blockIdx.x * kOptThreads + threadIdx.x;
    for (; i + (kOptU - 1) * stride < n4; i += kOptU * stride) {
        float4 v[kOptU];
#pragma unroll
        for (int u = 0; u < kOptU; ++u) v[u] = __ldg(g4 + i + u * stride);      // read again by adam_clip_kernel: keep it in L2
#pragma unroll
        for (int u = 0; u < kOptU; ++u) {
            const float s = fmaf(v[u].x, v[u].x, fmaf(v[u].y, v[u].y, fmaf(v[u].z, v[u].z, v[u].w * v[u].w)));
            acc += (double)s;
        }
    }
    for (; i < n4; i += stride) {
        const float4 v = __ldg(g4 + i);
        acc += (double)fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, v.w * v.w)));
    }
    if (blockIdx.x == 0 && threadIdx.x < (int)(n & 3)) {                         // the up-to-three elements past the last float4
        const float t = __ldg(g + (n4 << 2) + threadIdx.x);
        acc += (double)(t * t);
    }
    __shared__ double red[kOptThreads / 32];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < kOptThreads / 32; ++w) s += red[w];
        part[blockIdx.x] = s;
        if (blockIdx.x == 0) *step += 1.0f;                                      // t of the update the next launch applies
    }
}

struct AdamArgs {
    float inv_world, clip, lr, beta1, beta2, eps, weight_decay;
};

__device__ __forceinline__ void adam1(float &p, float g, float &m, float &v, const AdamArgs &a, bool scale_world, float coef,
                                      float step_size, float bc2_sqrt) {
    if (scale_world) g *= a.inv_world;
    g *= coef;
    if (a.weight_decay != 0.f) g = fmaf(p, a.weight_decay, g);
    m = fmaf(a.beta1, m, fmaf(-a.beta1, g, g));
    const float g2 = g * g;
    v = fmaf(a.beta2, v, fmaf(-a.beta2, g2, g2));
    const float denom = sqrtf(v) / bc2_sqrt + a.eps;
    p -= step_size * m / denom;
}

__global__ void __launch_bounds__(kOptThreads) adam_clip_kernel(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m,
                                                                float *__restrict__ v, long n, const double *__restrict__ part, int n_part,
                                                                const float *__restrict__ step, AdamArgs a, float *__restrict__ norm_out) {
    __shared__ double red[kOptThreads / 32];
    __shared__ float s_coef, s_step_size, s_bc2_sqrt;
    {   // every CTA folds the same partials in the same order: one global norm, bit-identical in all of them
        double acc = 0.0;
        for (int i = threadIdx.x; i < n_part; i += kOptThreads) acc += part[i];
        acc = warp_sum(acc);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < kOptThreads / 32; ++w) s += red[w];
            const float norm = (float)sqrt(s) * a.inv_world;                     // of the MEAN gradient (what the clip sees)
            float coef = 1.f;
            if (a.clip > 0.f) coef = fminf(a.clip / (norm + 1e-6f), 1.f);       // clip_grad_norm_: max_norm / (total_norm + 1e-6), <= 1
            const float t = *step;
            const float bc1 = 1.f - powf(a.beta1, t), bc2 = 1.f - powf(a.beta2, t);
            s_coef = coef;
            s_step_size = a.lr / bc1;
            s_bc2_sqrt = sqrtf(bc2);
            if (blockIdx.x == 0 && norm_out) *norm_out = norm;
        }
        __syncthreads();
    }
    const float coef = s_coef, step_size = s_step_size, bc2_sqrt = s_bc2_sqrt;
    const bool scale_world = a.inv_world != 1.f;
    const long n4 = n >> 2;
    float4 *p4 = reinterpret_cast<float4 *>(p), *m4 = reinterpret_cast<float4 *>(m), *v4 = reinterpret_cast<float4 *>(v);
    const float4 *g4 = reinterpret_cast<const float4 *>(g);
    const long stride = (long)gridDim.x * kOptThreads;
    for (long i = (long)blockIdx.x * kOptThreads + threadIdx.x; i < n4; i += 2 * stride) {
        const long j = i + stride;
        const bool two = j < n4;
        float4 P0 = p4[i], G0 = ldg_stream(g4 + i), M0 = m4[i], V0 = v4[i];      // eight 128-bit loads in flight before the first use
        float4 P1, G1, M1, V1;
        if (two) { P1 = p4[j]; G1 = ldg_stream(g4 + j); M1 = m4[j]; V1 = v4[j]; }
        adam1(P0.x, G0.x, M0.x, V0.x, a, scale_world, coef, step_size, bc2_sqrt);
        adam1(P0.y, G0.y, M0.y, V0.y, a, scale_world, coef, step_size, bc2_sqrt);
        adam1(P0.z, G0.z, M0.z, V0.z, a, scale_world, coef, step_size, bc2_sqrt);
        adam1(P0.w, G0.w, M0.w, V0.w, a, scale_world, coef, step_size, bc2_sqrt);
        p4[i] = P0; m4[i] = M0; v4[i] = V0;
        if (two) {
            adam1(P1.x, G1.x, M1.x, V1.x, a, scale_world, coef, step_size, bc2_sqrt);
            adam1(P1.y, G1.y, M1.y, V1.y, a, scale_world, coef, step_size, bc2_sqrt);
            adam1(P1.z, G1.z, M1.z, V1.z, a, scale_world, coef, step_size, bc2_sqrt);
            adam1(P1.w, G1.w, M1.w, V1.w, a, scale_world, coef, step_size, bc2_sqrt);
            p4[j] = P1; m4[j] = M1; v4[j] = V1;
        }
    }
    if (blockIdx.x == 0 && threadIdx.x < (int)(n & 3)) {
        const long e = (n4 << 2) + threadIdx.x;
        float P = p[e], M = m[e], V = v[e];
        adam1(P, g[e], M, V, a, scale_world, coef, step_size, bc2_sqrt);
        p[e] = P; m[e] = M; v[e] = V;
    }
}


// ---- rate-distortion loss tail (model.py:75-107 of the reference: R = clamp((sum nll_y + sum nll_z) / (N H W), 0), D = 1 - MS-SSIM or the
// MSE, loss = lambda D + R) on the per-patch bit counts K1 has already reduced.  Eager: ~8 scalar launches forward and as many backward,
// each ~2 us of launch latency on the step's critical path between the synthesis transform and its backward; here one launch each way.
__global__ void __launch_bounds__(32) rd_loss_fwd_kernel(const float *__restrict__ bits_y, int ny, const float *__restrict__ bits_z, int nz,
                                                         const float *__restrict__ dist, int similarity, float pixels, float lambda,
                                                         float *__restrict__ loss, float *__restrict__ R_out, float *__restrict__ D_out,
                                                         float *__restrict__ pass) {
    float sy = 0.f, sz = 0.f;
    for (int i = threadIdx.x; i < ny; i += 32) sy += __ldg(bits_y + i);
    for (int i = threadIdx.x; i < nz; i += 32) sz += __ldg(bits_z + i);
    sy = warp_sum(sy);
    sz = warp_sum(sz);
    if (threadIdx.x == 0) {
        const float r_raw = (sy + sz) / pixels;
        const float R = fmaxf(r_raw, 0.f);
        const float d = __ldg(dist);
        const float D = similarity ? 1.0f - d : d;
        *loss = __fadd_rn(__fmul_rn(lambda, D), R);   // lambda * D + R as the reference's two ops round it (model.py:105)
        *R_out = R;
        *D_out = D;
        *pass = r_raw >= 0.f ? 1.f : 0.f;          // clamp(min=0): the gradient passes where the input is >= 0
    }
}

__global__ void __launch_bounds__(256) rd_loss_bwd_kernel(const float *__restrict__ g_loss, const float *__restrict__ pass, float pixels,
                                                          float lambda, int similarity, int ny, int nz, float *__restrict__ g_bits_y,
                                                          float *__restrict__ g_bits_z, float *__restrict__ g_dist) {
    const float g = __ldg(g_loss);
    const float gb = g * __ldg(pass) / pixels;
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i < ny) g_bits_y[i] = gb;
    if (i < nz) g_bits_z[i] = gb;
    if (i == 0) *g_dist = similarity ? -(g * lambda) : g * lambda;
}

// ---- gradient pack: the per-parameter gradients autograd produced -> their slices of the flat gradient buffer, one launch per
// bucket.  torch.cat of the same ~80 tensors took 50 us for 25 MB (CatArrayBatchedCopy, 1 TB/s: profiles/r02ax_ncu_launches_bench_step.txt).
// The (pointer, offset, size) table travels in the kernel parameters, so a captured CUDA graph keeps it by value.
constexpr int kPackMax = 128;        // tensors per launch: 128 x (8 + 8 + 4 + 4) B = 3 KB of the 4 KB parameter space
constexpr int kPackChunk = 4096;     // floats per CTA
struct PackArgs {
    const float *src[kPackMax];
    long dst_off[kPackMax];
    int numel[kPackMax];
    int blk_end[kPackMax];           // running total of CTAs up to and including tensor t
    int n;
};

__global__ void __launch_bounds__(256) pack_flat_kernel(PackArgs a, float *__restrict__ dst) {
    const int b = blockIdx.x;
    int lo = 0, hi = a.n - 1;
    while (lo < hi) {                // the tensor this CTA's chunk belongs to: first t with blk_end[t] > b
        const int mid = (lo + hi) >> 1;
        if (a.blk_end[mid] > b) hi = mid;
        else lo = mid + 1;
    }
    const int first = lo ? a.blk_end[lo - 1] : 0;
    const long e0 = (long)(b - first) * kPackChunk;
    const int left = a.numel[lo] - (int)e0;
    const int cnt = left < kPackChunk ? left : kPackChunk;
    const float *s = a.src[lo] + e0;
    float *d = dst + a.dst_off[lo] + e0;
    if ((((uintptr_t)s | (uintptr_t)d) & 15) == 0) {
        const int c4 = cnt >> 2;
        const float4 *s4 = reinterpret_cast<const float4 *>(s);
        float4 *d4 = reinterpret_cast<float4 *>(d);
        float4 v[kPackChunk / 4 / 256];
#pragma unroll
        for (int u = 0; u < kPackChunk / 4 / 256; ++u) {
            const int i = threadIdx.x + u * 256;
            if (i < c4) v[u] = ldg_stream(s4 + i);
        }
#pragma unroll
        for (int u = 0; u < kPackChunk / 4 / 256; ++u) {
            const int i = threadIdx.x + u * 256;
            if (i < c4) d4[i] = v[u];        // default caching: the all-reduce / the norm kernel read it next
        }
        for (int i = (c4 << 2) + threadIdx.x; i < cnt; i += 256) d[i] = __ldg(s + i);
    } else {                         // a slice that starts off a 16-byte boundary (an odd-sized tensor came before it)
        float v[kPackChunk / 256];
#pragma unroll
        for (int u = 0; u < kPackChunk / 256; ++u) {
            const int i = threadIdx.x + u * 256;
            if (i < cnt) v[u] = __ldg(s + i);
        }
#pragma unroll
        for (int u = 0; u < kPackChunk / 256; ++u) {
            const int i = threadIdx.x + u * 256;
            if (i < cnt) d[i] = v[u];
        }
    }
}

SIC_REGISTER_KERNEL("pack_flat_kernel", pack_flat_kernel);
SIC_REGISTER_KERNEL("grad_sumsq_kernel", grad_sumsq_kernel);
SIC_REGISTER_KERNEL("adam_clip_kernel", adam_clip_kernel);

inline int opt_grid(long n) {
    const long want = ((n >> 2) + kOptThreads - 1) / kOptThreads;
    const long cap = (long)sm_count() * 4;                       // four resident CTAs per SM, every thread with several vectors
    return (int)(want < 1 ? 1 : (want < cap ? want : cap));
}

}  // namespace
}  // namespace sic

using namespace sic;

extern "C" size_t sic_clip_adam_workspace_bytes(long n) { return n > 0 ? (size_t)opt_grid(n) * sizeof(double) : 0; }

extern "C" int sic_clip_adam_step(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, long n, float *step,
                                  float inv_world, float clip, float lr, float beta1, float beta2, float eps, float weight_decay,
                                  float *norm_out, void *ws, size_t ws_bytes, void *stream) {
    SIC_CHECK_ARG(n > 0, "sic_clip_adam_step: empty parameter buffer n=%ld", n);
    SIC_CHECK_ARG(param && grad && exp_avg && exp_avg_sq && step && ws, "sic_clip_adam_step: null pointer");
    SIC_CHECK_ARG((((uintptr_t)param | (uintptr_t)grad | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15) == 0,
                  "sic_clip_adam_step: buffers must be 16-byte aligned");
    SIC_CHECK_ARG(((uintptr_t)ws & 7) == 0 && ws_bytes >= sic_clip_adam_workspace_bytes(n),
                  "sic_clip_adam_step: workspace of %zu B, needs %zu B (8-byte aligned)", ws_bytes, sic_clip_adam_workspace_bytes(n));
    SIC_CHECK_ARG(lr >= 0.f && beta1 >= 0.f && beta1 < 1.f && beta2 >= 0.f && beta2 < 1.f && eps >= 0.f && inv_world > 0.f,
                  "sic_clip_adam_step: invalid hyper-parameters lr=%g betas=(%g, %g) eps=%g inv_world=%g", lr, beta1, beta2, eps, inv_world);
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = opt_grid(n);
    grad_sumsq_kernel<<<grid, kOptThreads, 0, st>>>(grad, n, (double *)ws, step);
    SIC_CHECK_LAUNCH("sic_clip_adam_step (norm)");
    AdamArgs a{inv_world, clip, lr, beta1, beta2, eps, weight_decay};
    adam_clip_kernel<<<grid, kOptThreads, 0, st>>>(param, grad, exp_avg, exp_avg_sq, n, (const double *)ws, grid, step, a, norm_out);
    SIC_CHECK_LAUNCH("sic_clip_adam_step (update)");
    return 0;
}

extern "C" int sic_rd_loss_fwd(const float *bits_y, int ny, const float *bits_z, int nz, const float *dist, int similarity, long pixels,
                               float lambda, float *loss, float *R, float *D, float *pass, void *stream) {
    SIC_CHECK_ARG(ny > 0 && nz > 0 && pixels > 0, "sic_rd_loss_fwd: empty input (ny=%d nz=%d pixels=%ld)", ny, nz, pixels);
    SIC_CHECK_ARG(bits_y && bits_z && dist && loss && R && D && pass, "sic_rd_loss_fwd: null pointer");
    rd_loss_fwd_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(bits_y, ny, bits_z, nz, dist, similarity != 0, (float)pixels, lambda, loss, R, D, pass);
    SIC_CHECK_LAUNCH("sic_rd_loss_fwd");
    return 0;
}

extern "C" int sic_rd_loss_bwd(const float *g_loss, const float *pass, long pixels, float lambda, int similarity, int ny, int nz,
                               float *g_bits_y, float *g_bits_z, float *g_dist, void *stream) {
    SIC_CHECK_ARG(ny > 0 && nz > 0 && pixels > 0, "sic_rd_loss_bwd: empty input (ny=%d nz=%d pixels=%ld)", ny, nz, pixels);
    SIC_CHECK_ARG(g_loss && pass && g_bits_y && g_bits_z && g_dist, "sic_rd_loss_bwd: null pointer");
    const int n = ny > nz ? ny : nz;
    rd_loss_bwd_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(g_loss, pass, (float)pixels, lambda, similarity != 0, ny, nz, g_bits_y,
                                                                          g_bits_z, g_dist);
    SIC_CHECK_LAUNCH("sic_rd_loss_bwd");
    return 0;
}

extern "C" int sic_pack_flat(const float *const *srcs, const long *numels, const long *dst_offsets, int n, float *dst, void *stream) {
    SIC_CHECK_ARG(n > 0 && srcs && numels && dst_offsets && dst, "sic_pack_flat: nothing to pack or null pointer (n=%d)", n);
    cudaStream_t st = (cudaStream_t)stream;
    for (int t0 = 0; t0 < n; t0 += kPackMax) {
        PackArgs a;
        a.n = n - t0 < kPackMax ? n - t0 : kPackMax;
        long blocks = 0;
        for (int t = 0; t < a.n; ++t) {
            const long ne = numels[t0 + t];
            SIC_CHECK_ARG(srcs[t0 + t] && ne > 0 && ne < (1L << 31) && dst_offsets[t0 + t] >= 0 && (((uintptr_t)srcs[t0 + t]) & 3) == 0,
                          "sic_pack_flat: tensor %d: null / empty / >= 2^31 elements / negative offset", t0 + t);
            a.src[t] = srcs[t0 + t];
            a.dst_off[t] = dst_offsets[t0 + t];
            a.numel[t] = (int)ne;
            blocks += (ne + kPackChunk - 1) / kPackChunk;
            SIC_CHECK_ARG(blocks < (1L << 31), "sic_pack_flat: too many chunks");
            a.blk_end[t] = (int)blocks;
        }
        for (int t = a.n; t < kPackMax; ++t) { a.src[t] = nullptr; a.dst_off[t] = 0; a.numel[t] = 0; a.blk_end[t] = (int)blocks; }
        pack_flat_kernel<<<(unsigned)blocks, 256, 0, st>>>(a, dst);
        SIC_CHECK_LAUNCH("sic_pack_flat");
    }
    return 0;
}
