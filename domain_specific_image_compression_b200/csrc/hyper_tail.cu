// N4: the tail of the hyper-synthesis transform in one launch.
//
// Reference: /root/reference/code/modelv2/layers.py:141-152 (pool -> mlp_sigma / mlp_nu -> expand) and model.py:48-55
// (sigma = exp(log_sigma).mean((2,3)), nu = clamp(exp(log_nu).mean((2,3)), min_nu, max_nu)); the decoder repeats it at
// eval_selfcontained_entropy.py:99-106 between the z decode and the y table build.  In eager PyTorch that is ~15 launches on
// [B,N,1,1] tensors: an adaptive average pool, four 1x1 convolutions (cuDNN/cuBLAS GEMMs with bias), two ReLUs, two exps, two means
// over (h,w) of an already constant map, a clamp.  Here: one CTA per patch, everything in shared memory,
//   p[c]     = (sum_s t[b,c,s]) / (h w)                        lane l sums s = l, l+32, ... then a warp tree: one fixed order for
//                                                              NCHW and channels-last alike
//   hid[j]   = relu(b1[j] + sum_c W1[j,c] p[c])                warp per row, same lane/tree order
//   out[m]   = b2[m] + sum_n W2[m,n] hid[n]
//   sigma    = exp(out_sigma) ;  nu = clamp(exp(out_nu), min_nu, max_nu)
// written straight in kernel K1 / K3's broadcast layout [B*M].  The mean over (h,w) of a map that is constant over (h,w) is the
// value itself (exactly so in the reference too whenever h*w is a power of two), so it is not replayed.
// The summation order is fixed and independent of the batch size, so encoder and decoder derive bit-identical sigma/nu (and hence
// CDF tables) whatever batch each of them runs - cuDNN/cuBLAS give no such guarantee across batch sizes.  Against the eager chain
// the result differs by accumulation order only (~1e-6 relative; bit-exactness against a library GEMM is not attainable).
//
// Backward (training): one CTA per patch for the per-patch chain (d log, d hidden through the ReLU masks, d p, and the broadcast
// write of dt = dp / (h w)), then one small kernel that folds the weight / bias gradients over the batch in a fixed order.
#include "common.cuh"

namespace sic {
namespace {

constexpr int kHtThreads = 1024, kHtWarps = kHtThreads / 32;   // one CTA per patch: the chain is latency-bound, so it gets a full SM's warps

struct Mlp {
    const float *w1, *b1, *w2, *b2;   // [N,N], [N], [M,N], [M]
};

// Two rows at a time (independent loads in flight for both), each in the order `lane l sums c = l, l+32, ... then a warp tree`.
__device__ __forceinline__ void dot_rows2(const float *__restrict__ w0, const float *__restrict__ w1, const float *v, int n, int lane, float &r0,
                                          float &r1) {
    float a0 = 0.f, a1 = 0.f;
#pragma unroll 4
    for (int c = lane; c < n; c += 32) {
        const float x = v[c];
        a0 = fmaf(__ldg(w0 + c), x, a0);
        a1 = fmaf(__ldg(w1 + c), x, a1);
    }
    r0 = warp_sum(a0);
    r1 = warp_sum(a1);
}

// save layout per patch: p[N], hid_sigma[N], hid_nu[N], e_nu[M] (= exp(log_nu) before the clamp)
// r02i: 100 us per launch with 8 warps per patch (16 channels pooled one after the other per warp through lane-strided,
// 512-byte-apart loads; 32 / 48 dependent row dots per warp).  Now 32 warps, channels-last pooling four channels per 128-bit load with
// all loads of a lane independent, two rows per dot: same sums in the same order, a few microseconds.
__global__ void __launch_bounds__(kHtThreads) hyper_tail_fwd_kernel(const float *__restrict__ t, int N, int M, int HW, int channels_last,
                                                                  Mlp ms, Mlp mn, float min_nu, float max_nu, float *__restrict__ sigma,
                                                                  float *__restrict__ nu, float *__restrict__ save) {
    extern __shared__ float sm[];   // p[N], hs[N], hn[N]
    float *p = sm, *hs = sm + N, *hn = sm + 2 * N;
    const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float *tb = t + (size_t)b * N * HW;
    if (channels_last && (N & 3) == 0 && ((uintptr_t)tb & 15) == 0) {
        for (int q = warp; q < N / 4; q += kHtWarps) {           // lane l sums positions l, l+32, ... of four channels at once
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
            for (int sidx = lane; sidx < HW; sidx += 32) {
                const float4 v = __ldg(reinterpret_cast<const float4 *>(tb + (size_t)sidx * N) + q);
                a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
            }
            a.x = warp_sum(a.x); a.y = warp_sum(a.y); a.z = warp_sum(a.z); a.w = warp_sum(a.w);
            if (lane == 0) {
                p[4 * q] = __fdiv_rn(a.x, (float)HW); p[4 * q + 1] = __fdiv_rn(a.y, (float)HW);
                p[4 * q + 2] = __fdiv_rn(a.z, (float)HW); p[4 * q + 3] = __fdiv_rn(a.w, (float)HW);
            }
        }
    } else {
        for (int c = warp; c < N; c += kHtWarps) {
            float a = 0.f;
            if (channels_last) {
#pragma unroll 8
                for (int sidx = lane; sidx < HW; sidx += 32) a += __ldg(tb + (size_t)sidx * N + c);
            } else {
#pragma unroll 8
                for (int sidx = lane; sidx < HW; sidx += 32) a += __ldg(tb + (size_t)c * HW + sidx);
            }
            a = warp_sum(a);
            if (lane == 0) p[c] = __fdiv_rn(a, (float)HW);
        }
    }
    __syncthreads();
    for (int j = warp; j < N; j += kHtWarps) {                   // row j of both first layers
        float a0, a1;
        dot_rows2(ms.w1 + (size_t)j * N, mn.w1 + (size_t)j * N, p, N, lane, a0, a1);
        if (lane == 0) {
            hs[j] = fmaxf(a0 + __ldg(ms.b1 + j), 0.f);
            hn[j] = fmaxf(a1 + __ldg(mn.b1 + j), 0.f);
        }
    }
    __syncthreads();
    float *sv = save ? save + (size_t)b * (3 * N + M) : nullptr;
    for (int j = warp; j < 2 * M; j += 2 * kHtWarps) {           // rows j and j + kHtWarps of the stacked second layers
        const int j1 = j + kHtWarps;
        const bool sec0 = j >= M, sec1 = j1 >= M;
        const int r0 = sec0 ? j - M : j, r1 = (j1 < 2 * M) ? (sec1 ? j1 - M : j1) : r0;
        const Mlp &m0 = sec0 ? mn : ms, &m1 = sec1 ? mn : ms;
        float a0 = 0.f, a1 = 0.f;
        {   // the two rows may read different hidden vectors: two accumulation chains, each in the fixed order
#pragma unroll 4
            for (int c = lane; c < N; c += 32) {
                a0 = fmaf(__ldg(m0.w2 + (size_t)r0 * N + c), (sec0 ? hn : hs)[c], a0);
                a1 = fmaf(__ldg(m1.w2 + (size_t)r1 * N + c), (sec1 ? hn : hs)[c], a1);
            }
            a0 = warp_sum(a0);
            a1 = warp_sum(a1);
        }
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                if (k == 1 && j1 >= 2 * M) break;
                const bool second = k ? sec1 : sec0;
                const int r = k ? r1 : r0;
                const float e = expf((k ? a1 : a0) + __ldg((second ? mn : ms).b2 + r));
                if (second) {
                    nu[(size_t)b * M + r] = fminf(fmaxf(e, min_nu), max_nu);
                    if (sv) sv[3 * N + r] = e;
                } else {
                    sigma[(size_t)b * M + r] = e;
                }
            }
        }
    }
    if (sv) for (int c = threadIdx.x; c < N; c += kHtThreads) { sv[c] = p[c]; sv[N + c] = hs[c]; sv[2 * N + c] = hn[c]; }
}

// scratch layout per patch: dlog_sigma[M], dlog_nu[M], dhid_sigma[N], dhid_nu[N]
// The two matrix-vector products of the chain (W2^T dlog: 2N outputs of length M; W1^T dhid: N outputs of length 2N) are split over
// the CTA's 1024 threads as (output, chunk of the reduction axis): every thread runs a short, 8-way unrolled chain of independent
// loads, the chunk partials meet in shared memory and are added in chunk order (fixed order: deterministic).  r02i: 57 us per
// launch with one thread per output walking the whole axis.
__global__ void __launch_bounds__(kHtThreads) hyper_tail_bwd_kernel(const float *__restrict__ dsigma, const float *__restrict__ dnu,
                                                                  const float *__restrict__ sigma, const float *__restrict__ save, int N, int M,
                                                                  int HW, int channels_last, Mlp ms, Mlp mn, float min_nu, float max_nu,
                                                                  float *__restrict__ dt, float *__restrict__ scratch) {
    extern __shared__ float sm[];   // dls[M], dln[M], dhs[N], dhn[N], dp[N], red[kHtThreads]
    float *dls = sm, *dln = sm + M, *dhs = sm + 2 * M, *dhn = dhs + N, *dp = dhn + N, *red = dp + N;
    const int b = blockIdx.x, tid = threadIdx.x;
    const float *sv = save + (size_t)b * (3 * N + M);
    float *sc = scratch + (size_t)b * (2 * M + 2 * N);
    for (int m = tid; m < M; m += kHtThreads) {
        const float gs = dsigma ? dsigma[(size_t)b * M + m] : 0.f, gn = dnu ? dnu[(size_t)b * M + m] : 0.f;
        const float e = sv[3 * N + m];
        const float a = gs * sigma[(size_t)b * M + m];                         // d/dlog_sigma of exp
        const float c = (e >= min_nu && e <= max_nu) ? gn * e : 0.f;          // clamp passes gradient on the closed interval (torch.clamp)
        dls[m] = a; dln[m] = c;
        sc[m] = a; sc[M + m] = c;
    }
    __syncthreads();
    // dhid[n] = relu'(hid[n]) * sum_m W2[m,n] dlog[m]        (n over both MLPs: 2N outputs)
    for (int base = 0; base < 2 * N; base += kHtThreads) {
        const int n_out = min(2 * N - base, kHtThreads);
        const int chunks = kHtThreads / n_out, len = (M + chunks - 1) / chunks;
        const int chunk = tid / n_out, o = tid - chunk * n_out, n = base + o;
        float a = 0.f;
        if (chunk < chunks) {
            const bool second = n >= N;
            const int r = second ? n - N : n;
            const float *w2 = (second ? mn : ms).w2 + r;
            const float *dl = second ? dln : dls;
            const int m1 = min(M, (chunk + 1) * len);
#pragma unroll 8
            for (int m = chunk * len; m < m1; ++m) a = fmaf(__ldg(w2 + (size_t)m * N), dl[m], a);
        }
        red[tid] = a;
        __syncthreads();
        if (tid < n_out) {
            float t = 0.f;
            for (int k = 0; k < chunks; ++k) t += red[k * n_out + tid];
            const bool second = n >= N;
            const int r = second ? n - N : n;
            t = sv[(second ? 2 * N : N) + r] > 0.f ? t : 0.f;
            (second ? dhn : dhs)[r] = t;
            sc[2 * M + n] = t;
        }
        __syncthreads();
    }
    // dp[c] = (sum_j W1s[j,c] dhs[j] + W1n[j,c] dhn[j]) / (h w)    (reduction axis: the stacked 2N hidden units)
    const float inv = 1.0f / (float)HW;
    for (int base = 0; base < N; base += kHtThreads) {
        const int n_out = min(N - base, kHtThreads);
        const int chunks = kHtThreads / n_out, len = (2 * N + chunks - 1) / chunks;
        const int chunk = tid / n_out, o = tid - chunk * n_out, c = base + o;
        float a = 0.f;
        if (chunk < chunks) {
            const int j1 = min(2 * N, (chunk + 1) * len);
#pragma unroll 8
            for (int j = chunk * len; j < j1; ++j) {
                const bool second = j >= N;
                const int r = second ? j - N : j;
                a = fmaf(__ldg((second ? mn : ms).w1 + (size_t)r * N + c), (second ? dhn : dhs)[r], a);
            }
        }
        red[tid] = a;
        __syncthreads();
        if (tid < n_out) {
            float t = 0.f;
            for (int k = 0; k < chunks; ++k) t += red[k * n_out + tid];
            dp[c] = t * inv;
        }
        __syncthreads();
    }
    if (dt) {
        float *db = dt + (size_t)b * N * HW;
        const long n = (long)N * HW;
        if (channels_last) for (long i = tid; i < n; i += kHtThreads) db[i] = dp[i % N];
        else for (long i = tid; i < n; i += kHtThreads) db[i] = dp[i / HW];
    }
}

// dW2[m,n] = sum_b dlog[b,m] hid[b,n]; db2[m] = sum_b dlog[b,m]; dW1[j,c] = sum_b dhid[b,j] p[b,c]; db1[j] = sum_b dhid[b,j]
__global__ void __launch_bounds__(256) hyper_tail_wgrad_kernel(const float *__restrict__ save, const float *__restrict__ scratch, int B, int N, int M,
                                                             float *dw1s, float *db1s, float *dw2s, float *db2s, float *dw1n, float *db1n,
                                                             float *dw2n, float *db2n) {
    const int per = N * N + N + M * N + M;                   // one MLP: W1, b1, W2, b2
    const int e = blockIdx.x * 256 + threadIdx.x;
    if (e >= 2 * per) return;
    const bool second = e >= per;
    int i = second ? e - per : e;
    const int SV = 3 * N + M, SC = 2 * M + 2 * N;
    const int hid_off = second ? 2 * N : N, dl_off = second ? M : 0, dh_off = 2 * M + (second ? N : 0);
    float a = 0.f;
    float *dst;
    if (i < N * N) {                                         // dW1[j,c]
        const int j = i / N, c = i - j * N;
        for (int b = 0; b < B; ++b) a = fmaf(scratch[(size_t)b * SC + dh_off + j], save[(size_t)b * SV + c], a);
        dst = (second ? dw1n : dw1s) + i;
    } else if ((i -= N * N) < N) {                           // db1[j]
        for (int b = 0; b < B; ++b) a += scratch[(size_t)b * SC + dh_off + i];
        dst = (second ? db1n : db1s) + i;
    } else if ((i -= N) < M * N) {                           // dW2[m,n]
        const int m = i / N, n = i - m * N;
        for (int b = 0; b < B; ++b) a = fmaf(scratch[(size_t)b * SC + dl_off + m], save[(size_t)b * SV + hid_off + n], a);
        dst = (second ? dw2n : dw2s) + i;
    } else {                                                 // db2[m]
        i -= M * N;
        for (int b = 0; b < B; ++b) a += scratch[(size_t)b * SC + dl_off + i];
        dst = (second ? db2n : db2s) + i;
    }
    *dst = a;
}

SIC_REGISTER_KERNEL("hyper_tail_fwd_kernel", hyper_tail_fwd_kernel);
SIC_REGISTER_KERNEL("hyper_tail_bwd_kernel", hyper_tail_bwd_kernel);

}  // namespace
}  // namespace sic

using namespace sic;

extern "C" size_t sic_hyper_tail_save_floats(int B, int N, int M) { return (B > 0 && N > 0 && M > 0) ? (size_t)B * (3 * N + M) : 0; }
extern "C" size_t sic_hyper_tail_scratch_floats(int B, int N, int M) { return (B > 0 && N > 0 && M > 0) ? (size_t)B * (2 * M + 2 * N) : 0; }

extern "C" int sic_hyper_tail_fwd(const float *t, int B, int N, int M, int HW, int channels_last, const float *w1s, const float *b1s,
                                  const float *w2s, const float *b2s, const float *w1n, const float *b1n, const float *w2n,
                                  const float *b2n, float min_nu, float max_nu, float *sigma, float *nu, float *save, void *stream) {
    SIC_CHECK_ARG(B > 0 && N > 0 && M > 0 && HW > 0, "sic_hyper_tail_fwd: empty shape B=%d N=%d M=%d HW=%d", B, N, M, HW);
    SIC_CHECK_ARG(t && w1s && b1s && w2s && b2s && w1n && b1n && w2n && b2n && sigma && nu, "sic_hyper_tail_fwd: null pointer");
    SIC_CHECK_ARG(N <= 4096, "sic_hyper_tail_fwd: N=%d exceeds the shared-memory staging (4096)", N);
    Mlp ms{w1s, b1s, w2s, b2s}, mn{w1n, b1n, w2n, b2n};
    hyper_tail_fwd_kernel<<<B, kHtThreads, 3 * N * sizeof(float), (cudaStream_t)stream>>>(t, N, M, HW, channels_last, ms, mn, min_nu, max_nu,
                                                                                           sigma, nu, save);
    SIC_CHECK_LAUNCH("sic_hyper_tail_fwd");
    return 0;
}

extern "C" int sic_hyper_tail_bwd(const float *dsigma, const float *dnu, const float *sigma, const float *save, int B, int N, int M, int HW,
                                  int channels_last, const float *w1s, const float *w2s, const float *w1n, const float *w2n, float min_nu,
                                  float max_nu, float *dt, float *dw1s, float *db1s, float *dw2s, float *db2s, float *dw1n, float *db1n,
                                  float *dw2n, float *db2n, float *scratch, void *stream) {
    SIC_CHECK_ARG(B > 0 && N > 0 && M > 0 && HW > 0, "sic_hyper_tail_bwd: empty shape B=%d N=%d M=%d HW=%d", B, N, M, HW);
    SIC_CHECK_ARG(sigma && save && w1s && w2s && w1n && w2n && scratch, "sic_hyper_tail_bwd: null pointer");
    SIC_CHECK_ARG(dw1s && db1s && dw2s && db2s && dw1n && db1n && dw2n && db2n, "sic_hyper_tail_bwd: null gradient pointer");
    SIC_CHECK_ARG(N <= 2048 && M <= 4096, "sic_hyper_tail_bwd: N=%d / M=%d exceed the shared-memory staging", N, M);
    cudaStream_t st = (cudaStream_t)stream;
    Mlp ms{w1s, nullptr, w2s, nullptr}, mn{w1n, nullptr, w2n, nullptr};
    hyper_tail_bwd_kernel<<<B, kHtThreads, (2 * M + 3 * N + kHtThreads) * sizeof(float), st>>>(dsigma, dnu, sigma, save, N, M, HW, channels_last, ms, mn,
                                                                                 min_nu, max_nu, dt, scratch);
    SIC_CHECK_LAUNCH("sic_hyper_tail_bwd");
    const int total = 2 * (N * N + N + M * N + M);
    hyper_tail_wgrad_kernel<<<(total + 255) / 256, 256, 0, st>>>(save, scratch, B, N, M, dw1s, db1s, dw2s, db2s, dw1n, db1n, dw2n, db2n);
    SIC_CHECK_LAUNCH("sic_hyper_tail_bwd (weights)");
    return 0;
}
