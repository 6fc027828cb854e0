/* sic.h — C ABI of the B200-native Student-t entropy bottleneck + GDN hot path.
 *
 * The reference (Dimitrinov74/Domain-Specific-Image-Compression, code/modelv2) has no FFI layer: its boundary is
 * the Python API of model.py / layers.py / distributions.py / eval_selfcontained_entropy.py.  Each entry point
 * below replaces the eager-op chain of one of those functions (cited per function, paths relative to
 * /root/reference/code/modelv2) and is what a binding on the reference side would call (INTEGRATION.md shows the
 * ctypes stubs).
 *
 * Conventions (all functions):
 *   - plain pointers + extents; device pointers unless the name says `_host`; tensors are contiguous NCHW float32.
 *   - `stream` is a cudaStream_t passed as void*; work is only ENQUEUED (no allocation, no synchronisation),
 *     so calls are re-entrant across streams/threads.  Outputs and workspaces are caller-allocated.
 *   - return 0 on success; >0 = cudaError_t; <0 = SIC_E_* ; text via sic_last_error() (thread local).
 *   - workspaces must be zero-filled ONCE by the caller before first use; kernels leave them zeroed again.
 */
#ifndef SIC_H_
#define SIC_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SIC_VERSION 100

enum { SIC_E_BADARG = -1, SIC_E_WORKSPACE = -2, SIC_E_UNSUPPORTED = -3, SIC_E_OVERFLOW = -4, SIC_E_TRUNCATED = -5,
       SIC_E_CORRUPT = -6 /* a coded stream decodes to the end but not back to the coder's initial state, or leaves words over */ };

/* quantisation of the latent, model.py:27-35 */
enum {
    SIC_QUANT_NONE = 0,         /* y_tilde = y (likelihood-only call: StudentT.neg_log2_prob on a given tensor) */
    SIC_QUANT_ROUND = 1,        /* torch.round: half-to-even, keeps -0.0 (model.py:32-33)                       */
    SIC_QUANT_NOISE_TENSOR = 2, /* y + noise, noise supplied (parity mode of model.py:29-31)                    */
    SIC_QUANT_NOISE_PHILOX = 3  /* y + U(-1/2,1/2) from in-kernel Philox4x32-10 (seed, offset read on device)   */
};

/* likelihood model */
enum {
    SIC_LIK_STUDENTT_DENSITY = 0, /* -log2 t_nu(y~; mu, sigma): distributions.py:20-31 (what forward()/loss use) */
    SIC_LIK_GAUSSIAN = 1,         /* FactorizedGaussian, `sigma` argument = log_sigma[C]: distributions.py:39-46 */
    SIC_LIK_STUDENTT_CDFDIFF = 2  /* -log2 (T_nu(y~+1/2) - T_nu(y~-1/2)) : north_star / intent of
                                     eval_selfcontained_entropy.py:56-59                                        */
};

/* where sigma / nu / mu live */
enum {
    SIC_PARAM_BROADCAST = 0, /* one value per (b,c): [B*C]  (spatial_params=False, model.py:53-55) */
    SIC_PARAM_SPATIAL = 1,   /* one value per element: [B*C*HW] (spatial_params=True, model.py:49-51) */
    SIC_PARAM_CHANNEL = 2    /* one value per channel: [C] (Gaussian z prior) */
};

int sic_version(void);
const char *sic_last_error(void);

/* Introspection of the loaded binary: the hot kernels by name (template arguments as ncu prints them, without spaces: "gdn_bwd_nhwc_kernel<0>", "bottleneck_bwd_kernel<0,1>", ...)
 * and the registers per thread each was compiled to (cudaFuncGetAttributes).  bench.py uses it to quote the DRAM traffic of a
 * committed ncu capture (profiles/) only when that capture is of the same register allocation as the library that is running. */
int sic_kernel_count(void);
const char *sic_kernel_name(int i);
int sic_kernel_registers(const char *name);

/* ------------------------------------------------------------------------------------------------------------------
 * K1  fused quantise + likelihood + rate.   Replaces model.py:27-35 (quantize), distributions.py:20-31 / :39-46
 * (neg_log2_prob) and the two .sum() of model.py:77 (≈25 eager launches) with one launch.
 *   y        [B,C,HW]      latent
 *   noise    [B,C,HW]      only for SIC_QUANT_NOISE_TENSOR, else NULL
 *   philox   device uint64[2] = {seed, offset}, only for SIC_QUANT_NOISE_PHILOX (vector index v=i/4 is the counter).
 *            IN/OUT: the kernel adds 1 to the offset when it retires, so a CUDA-graph replay draws fresh noise.
 *   mu       NULL (=0, the reference has no location head) or same layout as sigma
 *   sigma,nu per `param_layout` (raw values; the clamps of distributions.py:23-24 are applied inside)
 *   y_tilde  [B,C,HW] out; nll [B,C,HW] out (may be NULL: rate only); bits [B] out = sum of nll per patch
 *   workspace: sic_bottleneck_workspace_bytes(B,C,HW) bytes, zero-initialised once.
 * Reduction order is fixed (warp tree -> per-(row,segment) partials -> per-patch sequential in float64): deterministic. */
size_t sic_bottleneck_workspace_bytes(int B, int C, int HW);
int sic_bottleneck_fwd(const float *y, const float *noise, uint64_t *philox, const float *mu, const float *sigma,
                       const float *nu, int B, int C, int HW, int quant_mode, int lik_mode, int param_layout,
                       float *y_tilde, float *nll, float *bits, void *workspace, size_t workspace_bytes, void *stream);

/* Analytic backward of K1 (autograd of the same op chains in the reference; formulas SURVEY.md 8(a')).
 *   g_nll [B,C,HW] or NULL, g_bits [B] or NULL (upstream of nll / bits); g_ytilde [B,C,HW] or NULL.
 *   dy = g_ytilde + (g_nll + g_bits[b]) * dnll/dy~   (0 for SIC_QUANT_ROUND: torch.round has zero gradient)
 *   dsigma / dnu / dmu in `param_layout` (broadcast: reduced over HW; channel: reduced over B and HW, for the Gaussian
 *   `dsigma` is d/dlog_sigma).  Clamp masks are closed intervals (torch.clamp).  Any of them may be NULL. */
int sic_bottleneck_bwd(const float *y_tilde, const float *mu, const float *sigma, const float *nu, const float *g_nll,
                       const float *g_bits, const float *g_ytilde, int B, int C, int HW, int quant_mode, int lik_mode,
                       int param_layout, float *dy, float *dmu, float *dsigma, float *dnu, void *workspace,
                       size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------------------------------------------------------
 * K2  GDN / IGDN, diagonal gamma — the path the reference executes (layers.py:19-27; depthwise 1x1 conv).
 *   beta_param [C], gamma_weight [C] are the STORED parameters (`beta`, `gamma_conv.weight`); the re-parameterisation
 *   beta = beta_param^2 - 2^-18, gamma = w^2 - 2^-18 (layers.py:20-21) is applied inside.
 *   Forward replays the eager rounding sequence with IEEE ops (mul, mul, add, sqrt.rn, div.rn | mul): bit-exact.
 *   channels_last != 0: x is stored NHWC (channel index = i % C) instead of NCHW.
 *   bias (nullable): the per-channel bias of the producing convolution, folded in as y = GDN(x + bias) with the same
 *   rounding as PyTorch's separate add_ (conv -> add_(bias) -> GDN, layers.py:29-31,49-73); backward then also returns
 *   dbias[c] = sum dx, which saves the add pass and the bias-gradient reduction over every GDN site. */
int sic_gdn_fwd(const float *x, const float *bias, const float *beta_param, const float *gamma_weight, int B, int C, int HW,
                int inverse, int channels_last, float *y, void *stream);
size_t sic_gdn_bwd_workspace_bytes(int B, int C, int HW);
/* sic_gdn_bwd = sic_gdn_bwd_partials (the streaming kernel: dx, and per-CTA partial sums of d(beta), d(gamma), d(bias) left in the
 * workspace) followed by sic_gdn_bwd_fold (fixed-order binary64 fold of those partials + chain rule through the squared
 * re-parameterisation) on the same stream.  The two halves are exported so that each kernel can be timed on its own. */
int sic_gdn_bwd_partials(const float *x, const float *bias, const float *g, const float *beta_param, const float *gamma_weight, int B,
                         int C, int HW, int inverse, int channels_last, float *dx, void *workspace, size_t workspace_bytes,
                         void *stream);
int sic_gdn_bwd_fold(const float *beta_param, const float *gamma_weight, int B, int C, int HW, int channels_last, float *dbias,
                     float *dbeta_param, float *dgamma_weight, const void *workspace, size_t workspace_bytes, void *stream);
int sic_gdn_bwd(const float *x, const float *bias, const float *g, const float *beta_param, const float *gamma_weight, int B,
                int C, int HW, int inverse, int channels_last, float *dx, float *dbias, float *dbeta_param,
                float *dgamma_weight, void *workspace, size_t workspace_bytes, void *stream);

/* G3  GDN / IGDN with a DENSE C x C gamma (north_star's tensor-core contraction; the reference stores the matrix,
 * layers.py:13, but never uses it: SURVEY.md D3).  s[p,i] = beta_i + sum_j gamma_ij x[p,j]^2, y = x/sqrt(s) | x*sqrt(s).
 *   x, y: CHANNELS-LAST activations viewed as [positions, C] (positions = B*H*W); beta_param [C] and gamma_param [C,C]
 *   (row i = output channel) are the stored parameters, re-parameterised inside like layers.py:20-21.
 *   tcgen05.mma kind::tf32 with an exact hi/lo split of x^2: tolerance-only mode (gamma at TF32 precision).
 *   This build: C in {32, 64, 96, 128, 192} (gamma and one x^2 tile resident in shared memory; 192 only in the pipelined
 *   kernel), otherwise SIC_E_UNSUPPORTED. */
int sic_gdn_dense_fwd(const float *x, const float *beta_param, const float *gamma_param, long positions, int C, int inverse,
                      float *y, void *stream);
/* The same operation with the kernel chosen explicitly (both are kept so each can be parity-tested and timed):
 *   SIC_DENSE_SERIAL     one CTA per SM walks load -> square -> MMA -> epilogue tile by tile (csrc/gdn_dense.cu);
 *   SIC_DENSE_PIPELINED  warp-specialised producer / MMA / epilogue roles over mbarrier pipelines, two TMEM accumulator
 *                        stages, gamma as the A operand so the epilogue needs no transpose (csrc/gdn_dense_ws.cu).
 * sic_gdn_dense_fwd uses SIC_DENSE_DEFAULT. */
enum { SIC_DENSE_SERIAL = 0, SIC_DENSE_PIPELINED = 1, SIC_DENSE_DEFAULT = SIC_DENSE_PIPELINED };
int sic_gdn_dense_fwd_variant(const float *x, const float *beta_param, const float *gamma_param, long positions, int C,
                              int inverse, float *y, int variant, void *stream);

/* G3 backward (SURVEY.md 8(a') G3) on the same pipelined tcgen05 kernel, two launches on `stream`:
 *   pass 1: s = beta + gamma x^2 (MMA), h = -1/2 g x / d^3 | +1/2 g x / d, direct = g / d | g d, per-channel partial sums of h;
 *   pass 2: t = gamma^T h (MMA, gamma transposed into the A operand), dx = direct + 2 x t.
 *   x, g, h, direct, dx: channels-last [positions, C]; h and direct are caller-allocated outputs (h is also what the caller
 *   contracts with x^2 for d(gamma_eff)_ij = sum_p h_i x_j^2, a plain library GEMM); dbeta_part [part_rows, C] with
 *   part_rows >= sic_gdn_dense_bwd_part_rows(positions, C): d(beta_eff) = column sums.  Gradients are w.r.t. the EFFECTIVE
 *   beta/gamma; the chain rule through the squared re-parameterisation (x 2 beta_param, x 2 gamma_param) is the caller's.
 *   gamma is consumed at TF32 precision as in the forward.  C in {32, 64, 96, 128, 192}.  Device-validated in round 2
 *   (tests/test_gpu_gdn.py::test_dense_gdn_fused_backward_vs_float64, 10 shapes incl. ragged tails and C = 192) and the default
 *   backward of GDN(dense=True). */
int sic_gdn_dense_bwd_part_rows(long positions, int C);
int sic_gdn_dense_bwd(const float *x, const float *g, const float *beta_param, const float *gamma_param, long positions, int C,
                      int inverse, float *h, float *direct, float *dx, float *dbeta_part, int part_rows, void *stream);

/* G3 backward, third pass: d(gamma_eff)[i][j] = sum_p h[p][i] * x[p][j]^2 on tcgen05 (kind::tf32, both operands MN-major straight
 * from the channels-last layout, exact hi/lo split of h and x^2, one TMEM accumulator for the whole kernel, per-CTA partials
 * folded in a fixed order).  h is pass 1's output of sic_gdn_dense_bwd.  dgamma_eff [C, C] row i = output channel; the chain rule
 * through gamma = gamma_param^2 - 2^-18 (x 2 gamma_param) is the caller's, as for sic_gdn_dense_bwd.  One streaming pass: 8 B/element.
 * workspace: sic_gdn_dense_dgamma_workspace_bytes(positions, C) bytes, any content.  C in {32, 64, 96, 128, 192}. */
size_t sic_gdn_dense_dgamma_workspace_bytes(long positions, int C);
int sic_gdn_dense_dgamma(const float *x, const float *h, long positions, int C, float *dgamma_eff, void *workspace,
                         size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------------------------------------------------------
 * N2  first analysis layer as ONE kernel: conv 3 -> C, 3x3, stride 1, zero padding 1 (layers.py:49-50, `conv(3, N, 3, 1)`) + bias +
 * GDN (layers.py:51, arithmetic :19-27).  Replaces cuDNN conv (legacy non-tensor-core engine: 3 channels) + add_(bias) + K2, and in
 * the backward K2's backward + cuDNN wgrad; the C x 256 x 256 intermediate is neither written nor read (csrc/conv0_gdn.cu).
 *   x [B, H, W, 3]: the image, CHANNELS-LAST.  w [C, 3, 3, 3] in (kh, kw, cin) order = a channels_last [C,3,3,3] weight's memory.
 *   bias (nullable) [C], beta_param [C], gamma_weight [C]: stored parameters as for sic_gdn_fwd.
 *   y [B, H, W, C] channels-last.  v_out (nullable): conv + bias before the GDN, same layout (tests).
 *   tcgen05 kind::tf32 with exact hi/lo splits of weights and taps: the convolution is evaluated to fp32 accuracy (2^-22 relative);
 *   GDN as v * rsqrt(beta + gamma v^2).  Tolerance-only (training) path: the bit-exact latents of eval/compress stay on cuDNN + K2.
 * bwd: grad_y [B, H, W, C] -> dw [C, 27] (same order as w), dbias (nullable), dbeta_param, dgamma_weight (chain rule through the squared
 *   re-parameterisation applied); v is recomputed from the image, the image gets no gradient.  Deterministic (fixed-order binary64
 *   fold of per-CTA partials).  workspace: sic_conv0_gdn_bwd_workspace_bytes bytes, any content.  C in {32, 64, 96, 128, 192}. */
int sic_conv0_gdn_fwd(const float *x, const float *w, const float *bias, const float *beta_param, const float *gamma_weight, int B,
                      int H, int W, int C, float *y, float *v_out, void *stream);
size_t sic_conv0_gdn_bwd_workspace_bytes(int B, int H, int W, int C);
int sic_conv0_gdn_bwd(const float *x, const float *w, const float *bias, const float *beta_param, const float *gamma_weight,
                      const float *grad_y, int B, int H, int W, int C, float *dw, float *dbias, float *dbeta_param,
                      float *dgamma_weight, void *workspace, size_t workspace_bytes, void *stream);

/* N2, synthesis side: the gather halves of the last layer ConvTranspose2d(N, 3, 5, stride 2, padding 2, output_padding 1)
 * (layers.py:96-98) in its GEMM formulation (csrc/deconv_rgb.cu); the GEMM itself (a 1x1 convolution N -> 80) is a library call.
 *   D [P, 80] position-major (the channels-last output of the 1x1 convolution), P = B*H*W over the INPUT grid, column
 *   m = (kh * 5 + kw) * 3 + co for m < 75, columns 75..79 ignored (read) / zero (written).
 *   col2im: out [B, 2H, 2W, 3] channels-last = bias + sum of the taps that land on each output pixel, gathered in a fixed order
 *           (no atomics: deterministic; every tap element of D is read exactly once).  bias nullable.
 *   im2col: dD[p, m] = grad_out[b, 2 iy - 2 + kh, 2 ix - 2 + kw, co] (0 outside the image): the adjoint gather. */
int sic_deconv_rgb_col2im(const float *D, const float *bias, int B, int H, int W, float *out, void *stream);
int sic_deconv_rgb_im2col(const float *grad_out, int B, int H, int W, float *dD, void *stream);

/* ------------------------------------------------------------------------------------------------------------------
 * N4  tail of the hyper-synthesis transform in one launch: layers.py:141-152 (AdaptiveAvgPool2d(1) -> mlp_sigma / mlp_nu, two 1x1
 * convolutions with a ReLU each) + model.py:54-55 (exp, mean over an already constant map, clamp of nu); the decoder runs the same
 * chain at eval_selfcontained_entropy.py:99-106 between the z decode and the y tables.
 *   t [B, N, HW] (NCHW) or [B, HW, N] (channels_last != 0): output of h_s.h_s (the two transposed convolutions + ReLU, cuDNN).
 *   w1* [N, N], b1* [N], w2* [M, N], b2* [M]: the 1x1-conv weights of mlp_sigma (s) / mlp_nu (n), as stored ([out, in, 1, 1]).
 *   sigma, nu [B, M]: K1 / K3's broadcast layout; nu clamped to [min_nu, max_nu], sigma = exp(log_sigma) unclamped (model.py:54).
 *   save (nullable; training): sic_hyper_tail_save_floats(B,N,M) floats kept for the backward (pooled input, hidden layers, exp(log_nu)).
 * Fixed summation order, independent of the batch size: encoder and decoder derive bit-identical sigma / nu from identical z.
 * bwd: dsigma / dnu [B, M] (nullable = 0) -> dt (layout of t, nullable) and the eight parameter gradients (summed over the batch in a
 * fixed order); scratch: sic_hyper_tail_scratch_floats(B,N,M) floats.  Two launches. */
size_t sic_hyper_tail_save_floats(int B, int N, int M);
size_t sic_hyper_tail_scratch_floats(int B, int N, int M);
int sic_hyper_tail_fwd(const float *t, int B, int N, int M, int HW, int channels_last, const float *w1s, const float *b1s,
                       const float *w2s, const float *b2s, const float *w1n, const float *b1n, const float *w2n, const float *b2n,
                       float min_nu, float max_nu, float *sigma, float *nu, float *save, void *stream);
int sic_hyper_tail_bwd(const float *dsigma, const float *dnu, const float *sigma, const float *save, int B, int N, int M, int HW,
                       int channels_last, const float *w1s, const float *w2s, const float *w1n, const float *w2n, float min_nu,
                       float max_nu, float *dt, float *dw1s, float *db1s, float *dw2s, float *db2s, float *dw1n, float *db1n,
                       float *dw2n, float *db2n, float *scratch, void *stream);

/* ------------------------------------------------------------------------------------------------------------------
 * K4  symbols: eval_selfcontained_entropy.py:39-40,48 / :52-53,62 (per patch, no host sync).
 *   q [B,n_per_patch] float latent; do_round != 0 applies torch.round first.
 *   mins[b] = floor(min q_b) - tail ; maxs[b] = ceil(max q_b) + tail ; sym = int32(q) - mins[b]. */
int sic_quantize_indices(const float *q, int B, long n_per_patch, int do_round, int tail, int32_t *sym, int32_t *mins,
                         int32_t *maxs, void *stream);

/* K3  integer CDF tables: eval_selfcontained_entropy.py:17-23 (pmf_to_uint16_cdf), :41-47 (Gaussian), :54-61 (Student-t).
 *   kind 0: Gaussian, `sigma` = log_sigma[C] of the z prior (taken unclamped, :32), rows = B*C, row r -> channel r % C.
 *   kind 1: Student-t, sigma/nu raw per row (n_rows = B*rows_per_patch), location 0.
 *   Row r belongs to patch r / rows_per_patch; support mins[b]..maxs[b] (device int32).  out [n_rows, stride] uint16 with
 *   the symbol axis last; entries past L+1 are zero.  Arithmetic per spec SIC-CDF-1: deterministic, IEEE-only. */
int sic_build_cdf_tables(int kind, const float *sigma, const float *nu, int n_rows, int rows_per_patch, int C,
                         const int32_t *mins, const int32_t *maxs, int stride, uint16_t *out, void *stream);

/* ------------------------------------------------------------------------------------------------------------------
 * N3  SSIM statistics of ONE scale of the MS-SSIM distortion (model.py:93-102 calls piq.multi_scale_ssim; 11x11 Gaussian
 * window sigma 1.5, "valid" support, k1=.01, k2=.03 -> c1 = 1e-4, c2 = 9e-4 at data_range 1).
 *   X (reconstruction, clamped), Y (target): [planes, H, W] contiguous, planes = B*C.
 *   fwd: part_ss / part_cs [planes * sic_ssim_tiles(H,W)] per-tile sums of the ss and cs maps over the valid region (the host
 *        folds them and divides by (H-10)(W-10)); maps (nullable) [5, planes, H-10, W-10]: what bwd needs.
 *   bwd: dX = d(g_ss[plane]*mean(ss) + g_cs[plane]*mean(cs))/dX ; g_ss / g_cs [planes], either may be NULL (= 0). */
/* The same two kernels with the 2x2 average pooling between MS-SSIM scales folded in (even H and W): the forward also writes the next
 * scale's inputs x_pool / y_pool [planes, H/2, W/2] (nullable pair), the backward adds the gradient that arrives through that pooled
 * copy (g_pool [planes, H/2, W/2], nullable; avg_pool2d backward = 1/4 to each of the four pixels).  Replaces two avg_pool2d and one
 * avg_pool2d_backward + add per scale boundary. */
int sic_ssim_fwd_pool(const float *X, const float *Y, int planes, int H, int W, float c1, float c2, float *part_ss, float *part_cs,
                      float *maps, float *x_pool, float *y_pool, void *stream);
int sic_ssim_bwd_pool(const float *X, const float *Y, const float *maps, const float *g_ss, const float *g_cs, const float *g_pool,
                      int planes, int H, int W, float *dX, void *stream);
/* _ex: the same kernels with (a) either image in channels-last layout - x_bands / y_bands = 0 for planar [planes, H, W], = C for a
 * [B, H, W, C] tensor with plane = b * C + c (no conversion pass between the synthesis transform and the loss); dX has X's layout -
 * (b) the x_hat.clamp(0, 1) of model.py:98 folded in (clamp01: X clamped as it is read, dX zero outside [0, 1]), (c) g_scale
 * (nullable device scalar) multiplying g_ss and g_cs. */
int sic_ssim_fwd_ex(const float *X, const float *Y, int planes, int H, int W, int x_bands, int y_bands, int clamp01, float c1, float c2,
                    float *part_ss, float *part_cs, float *maps, float *x_pool, float *y_pool, void *stream);
int sic_ssim_bwd_ex(const float *X, const float *Y, const float *maps, const float *g_ss, const float *g_cs, const float *g_scale,
                    const float *g_pool, int planes, int H, int W, int x_bands, int y_bands, int clamp01, float *dX, void *stream);
/* MS-SSIM value from the per-tile partial sums of its scales, one launch: per plane prod_{l<L-1} relu(mean cs_l)^w_l *
 * relu(mean ss_{L-1})^w_{L-1} with w = weights (device [levels]; divided by their sum when normalize != 0, as piq does to
 * user-supplied weights), then the mean over planes -> out (device scalar);
 * coef (nullable, device [levels][2][planes]): d out / d mean ss_l (index 0) and / d mean cs_l (index 1), the g_ss / g_cs of
 * sic_ssim_bwd_ex.  part: one device buffer; HOST arrays offsets[l] (float offset of scale l's part_ss; its part_cs follows at
 * + planes * tiles[l]), tiles[l] = sic_ssim_tiles(H_l, W_l), n_valid[l] = (H_l - 10)(W_l - 10). */
int sic_msssim_combine(const float *part, const long *offsets, const int *tiles, const long *n_valid, int levels, int planes,
                       const float *weights, int normalize, float *out, float *coef, void *stream);
long sic_ssim_tiles(int H, int W);
int sic_ssim_fwd(const float *X, const float *Y, int planes, int H, int W, float c1, float c2, float *part_ss, float *part_cs,
                 float *maps, void *stream);
int sic_ssim_bwd(const float *X, const float *Y, const float *maps, const float *g_ss, const float *g_cs, int planes, int H,
                 int W, float *dX, void *stream);

/* ------------------------------------------------------------------------------------------------------------------
 * E1  entropy coder, format SIC-RANS-1 (replaces torchac.encode_float_cdf / decode_float_cdf at
 * eval_selfcontained_entropy.py:48,62,96,116).  HOST buffers.  One call = one stream (one patch, one latent).
 *   sym [n] int32 in [0,L); tables [n/sym_per_row rows, stride] uint16; returns bytes written (>=0) or SIC_E_*. */
long sic_rans_encode_host(const int32_t *sym, long n, const uint16_t *tables, int stride, int L, long sym_per_row,
                          uint8_t *out, long cap);
int sic_rans_decode_host(const uint8_t *in, long nbytes, long n, const uint16_t *tables, int stride, int L,
                         long sym_per_row, int32_t *sym);


/* N1  the same coder on the GPU: one warp per stream, all streams of a batch concurrently; bytes identical to the host coder.
 *   sym [n_streams, n] int32; tables [n_streams * rows_per_stream, stride] uint16; Ls [n_streams] int32 (symbols per
 *   stream's support); out [n_streams, cap] with cap >= 128 + 2n, cap % 4 == 0; out_nbytes [n_streams] (-1: symbol out of
 *   range or support L outside [1, min(4096, stride-1)]).  decode: status [n_streams] = 0, SIC_E_TRUNCATED, SIC_E_CORRUPT (final
 *   states / word count do not close) or SIC_E_BADARG (L outside [1, min(4096, stride-1)]).  All pointers are DEVICE pointers. */
/* sic_rans_encode_ws: the same encoder in two phases - a whole-GPU pass that turns every symbol into its (start, width, reciprocal)
 * triple, then the one-warp-per-stream state machine with the division replaced by a multiply (exact) - for the price of a workspace of
 * sic_rans_encode_workspace_bytes(n_streams, n) bytes (8 per symbol, any content).  Identical bytes. */
size_t sic_rans_encode_workspace_bytes(int n_streams, long n);
int sic_rans_encode_ws(const int32_t *sym, const uint16_t *tables, const int32_t *Ls, int n_streams, long n, long sym_per_row,
                       long rows_per_stream, int stride, uint8_t *out, long cap, int32_t *out_nbytes, void *workspace,
                       size_t workspace_bytes, void *stream);
int sic_rans_encode(const int32_t *sym, const uint16_t *tables, const int32_t *Ls, int n_streams, long n, long sym_per_row,
                    long rows_per_stream, int stride, uint8_t *out, long cap, int32_t *out_nbytes, void *stream);
int sic_rans_decode(const uint8_t *in, const int32_t *nbytes, const uint16_t *tables, const int32_t *Ls, int n_streams, long n,
                    long sym_per_row, long rows_per_stream, int stride, long cap, int32_t *sym, int32_t *status, void *stream);

/* ------------------------------------------------------------------------------------------------------------------
 * (e)  tail of the data-parallel training step on the flat parameter / gradient buffers (trainer.py): global-norm clipping
 * (train.py:200-202, torch.nn.utils.clip_grad_norm_) + optim.Adam (train.py:182-183), two launches (csrc/train_step.cu).
 *   grad: the SUM over the ranks of the gradients (what the all-reduce leaves); inv_world = 1 / world size (1 on one GPU).
 *   step: device scalar, the number of updates applied so far; incremented here (CUDA-graph capturable: no host state).
 *   clip <= 0: no clipping.  norm_out (nullable): the pre-clip global norm of the mean gradient.  grad is not modified.
 *   workspace: sic_clip_adam_workspace_bytes(n) bytes, 8-byte aligned. */
/* Tail of rate_distortion_loss (model.py:75-107): R = max((sum bits_y + sum bits_z) / pixels, 0), D = 1 - *dist (similarity != 0:
 * *dist is the MS-SSIM value) or *dist (the MSE), loss = lambda D + R; one launch forward, one backward.  bits_y [ny], bits_z [nz]:
 * per-patch bit counts from sic_bottleneck_fwd; loss, R, D, pass: device scalars (pass = 1 where the clamp lets the gradient through,
 * kept for the backward); bwd: g_bits_y [ny], g_bits_z [nz], g_dist (scalar) from the scalar g_loss. */
int sic_rd_loss_fwd(const float *bits_y, int ny, const float *bits_z, int nz, const float *dist, int similarity, long pixels, float lambda,
                    float *loss, float *R, float *D, float *pass, void *stream);
int sic_rd_loss_bwd(const float *g_loss, const float *pass, long pixels, float lambda, int similarity, int ny, int nz, float *g_bits_y,
                    float *g_bits_z, float *g_dist, void *stream);
/* Bias add (+ ReLU) after a bias-free cuDNN convolution, and its adjoint, on position-major activations [P, C] (channels-last,
 * P = B*H*W); csrc/bias_act.cu.  Replaces PyTorch's add_(bias) / relu_ and, in the backward, threshold_backward + the sum over
 * (B, H, W) for d(bias) (layers.py:104-139 hyper transforms, :73 last analysis convolution; also d(bias) of the last synthesis layer).
 *   fwd: t = act(t + bias) IN PLACE; relu != 0: act = ReLU as `v < 0 ? 0 : v`.  Same fp32 values as PyTorch's two ops.
 *   bwd: y / dt NULL: d(bias)[c] = sum_p g[p, c].  Else y = the forward's output: dt = g where y > 0 else 0, d(bias) = sum_p dt.
 *        Deterministic (fixed-order partial sums).  workspace: sic_bias_grad_workspace_bytes(P, C) bytes. */
int sic_bias_act_fwd(float *t, const float *bias, long P, int C, int relu, void *stream);
size_t sic_bias_grad_workspace_bytes(long P, int C);
int sic_bias_act_bwd(const float *g, const float *y, long P, int C, float *dt, float *dbias, void *workspace, size_t workspace_bytes,
                     void *stream);
/* Gradient pack of trainer.FlatTrainer: n device tensors (HOST arrays: srcs[t] device pointer, numels[t], dst_offsets[t] in floats)
 * copied to dst + dst_offsets[t], one launch per 128 tensors; the table is passed by value (graph-capturable). */
int sic_pack_flat(const float *const *srcs, const long *numels, const long *dst_offsets, int n, float *dst, void *stream);
size_t sic_clip_adam_workspace_bytes(long n);
int sic_clip_adam_step(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, long n, float *step, float inv_world,
                       float clip, float lr, float beta1, float beta2, float eps, float weight_decay, float *norm_out,
                       void *workspace, size_t workspace_bytes, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* SIC_H_ */
