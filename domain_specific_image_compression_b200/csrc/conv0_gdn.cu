// N2 at the largest site: the first analysis layer  conv 3->C 3x3 (stride 1, pad 1) + bias + GDN  as ONE kernel, forward and backward.
//
// Reference: /root/reference/code/modelv2/layers.py:49-51 (`conv(3, N, 3, 1)` followed by `GDN(N)`), GDN arithmetic layers.py:19-27.
// In the cfg2 step this site is 16 x C x 256 x 256.  Unfused it costs, per step (profiles/r02d conv probe + kernel sweep):
//   forward   cuDNN conv 552 us (legacy non-tensor-core engine: 3 channels; writes 537 MB) + GDN 172 us (read 537 MB, write 537 MB)
//   backward  GDN backward 273 us (read x, g; write dx: 1.6 GB) + cuDNN wgrad 820 us (legacy engine)
// The convolution has K = 27 inputs per output: it is not a GEMM worth a library call, it is a prologue.  Fused:
//   forward   read the image (12.6 MB), write y (537 MB):                 one third of the traffic, no intermediate tensor
//   backward  read g (537 MB) and the image; v = conv + bias is RECOMPUTED (27 MACs per element, on the tensor core); dv, d(beta),
//             d(gamma), d(bias) and the weight gradient dW = dv^T patches (a second tensor-core contraction, K = positions) never
//             leave the SM:                                                one seventh of the traffic
//
// One kernel template, BWD = false / true.  Structure = the pipelined dense-GDN kernel (gdn_dense_ws.cu):
//   producers (4 warps)  one position per thread and tile: the 27 image taps -> exact tf32 hi + lo -> im2col rows [position][32 k] in the
//                        K-major SWIZZLE_128B operand layout (B operand of GEMM 1); backward additionally the TRANSPOSED tile
//                        [k][position] in the same K-major layout (B operand of GEMM 2, whose K axis is the position)
//   MMA warp             GEMM 1: v[c, p] = sum_k W[c, k] patch[p, k], weights = A operand (M = output channels, so TMEM lanes are
//                        channels and every global access of the epilogue is coalesced in channels-last memory), 3 exact-split terms
//                        x 4 K-steps of 8.  Backward, GEMM 2: dW[c, k] += sum_p dv[c, p] patch[p, k] with dv read FROM TENSOR MEMORY
//                        (the .ts MMA form: lane = channel, column = position is exactly how the epilogue holds dv) into a per-tile
//                        accumulator that the epilogue warps drain into a shared-memory running sum (see `sdw`)
//   bias                 rides in the GEMMs: im2col column 27 is the constant 1 and weight column 27 is the bias, so GEMM 1 returns
//                        conv + bias and GEMM 2 returns d(bias) = sum dv in column 27 of dW - no per-element add, no per-lane sum
//   epilogue (16 warps)  forward: TMEM -> GDN -> y.  Backward: TMEM v + global g -> dv = g beta / d^3, per-lane sums of
//                        d(beta), d(gamma); dv split exactly into hi (stored over v, in place) and lo (third TMEM region)
//   end of kernel        per-CTA partials of dW and of the three per-channel sums; a second tiny kernel folds the <= 148 partials in a
//                        fixed order (binary64, deterministic, no atomics) and applies the chain rule to the stored parameters.
// C <= 128: one M = 128 block, 128 positions per tile.  C == 192: two M = 128 blocks (rows >= 192 zero), 64 positions per tile.
// Numerics: w = wh + wl, p = ph + pl (exact tf32 splits), v = wh ph + wl ph + wh pl in fp32: the dropped wl pl term is 2^-22 relative,
// i.e. the convolution is evaluated to fp32 accuracy - tighter than cuDNN's TF32 kernels, but not bit-identical to any of them
// (accumulation order), so model.eval() / compress() keep the cuDNN + GDN path whose latents are pinned bit-exactly against the
// reference; the fused layer serves training (noise quantisation: no bit-exactness contract) and is switchable (layers.FUSE_FIRST_LAYER).
// For the same reason its GDN uses y = v * rsqrt(beta + gamma v^2) with the 2-ulp MUFU rsqrt instead of the IEEE sqrt + division replay
// of gdn_math.cuh (~27 instructions per element, which 16 epilogue warps could not hide behind the 64 KB per tile they write).
#include "gdn_dense_ws.cuh"
#include "gdn_math.cuh"

namespace sic {
namespace {

using namespace umma;
using namespace gdnm;

constexpr int kC0Epi = 16, kC0Prod = 4;   // the epilogue does the per-element math: it gets the warps
constexpr int kC0EpiThreads = kC0Epi * 32, kC0ProdThreads = kC0Prod * 32;
constexpr int kC0Threads = kC0EpiThreads + kC0ProdThreads + 32;
constexpr int kC0K = 27;                 // 3 x 3 x 3 taps, K index = (kh * 3 + kw) * 3 + cin (channels-last order of weight and image)
constexpr int kC0KP = 32;                // K padded to one 128-byte swizzle row
constexpr int kC0One = 27;               // the constant-one column of the im2col rows (weight column = bias)

template <int C, bool BWD>
struct C0Cfg {
    static constexpr int MB = C > 128 ? 2 : 1;                 // M = 128 blocks
    static constexpr int TN = 128 / MB;                        // positions per tile (UMMA N of GEMM 1, K of GEMM 2)
    static constexpr int ROWS_W = MB * 128;
    static constexpr uint32_t W_BYTES = ROWS_W * 128;          // one of wh / wl: one 128-byte K row per output channel
    static constexpr uint32_t P_BYTES = TN * 128;              // one of ph / pl (and of the transposed pht / plt)
    static constexpr int NS1 = BWD ? 2 : 4;                    // stages of the im2col tile
    static constexpr int NS2 = BWD ? 3 : 0;                    // stages of the transposed tile (lives until GEMM 2 of its tile)
    static constexpr size_t SMEM = 2 * (size_t)W_BYTES + (size_t)(NS1 + NS2) * 2 * P_BYTES + 1024;
    static constexpr int NP = TN / 4;                          // columns per epilogue warp and block (4 warps share a lane quadrant)
    // TMEM: two v stages of MB * TN = 128 columns; backward: + dv_lo (128) + the dW accumulators (32 per block)
    static constexpr uint32_t COL_DVLO = 256, COL_DW = 384;
    static constexpr uint32_t TMEM_COLS = BWD ? 512 : 256;
    static_assert(C % 32 == 0 && (C <= 128 || C == 192), "first-layer kernel: C in {32,64,96,128,192}");
    static_assert(SMEM <= 227u * 1024u, "operands do not fit in shared memory");
};

struct C0Geom {
    int B, H, W;
    long P;          // B * H * W positions
};

__device__ __forceinline__ float tf32_hi(float v) { return __uint_as_float(__float_as_uint(v) & 0xFFFFE000u); }

__device__ __forceinline__ void split4(const float4 &v, float4 &hi, float4 &lo) {
    hi.x = tf32_hi(v.x); lo.x = v.x - hi.x;
    hi.y = tf32_hi(v.y); lo.y = v.y - hi.y;
    hi.z = tf32_hi(v.z); lo.z = v.z - hi.z;
    hi.w = tf32_hi(v.w); lo.w = v.w - hi.w;
}

__device__ __forceinline__ void sts32(uint32_t saddr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(saddr), "f"(v) : "memory"); }

// 16 registers per lane -> 32 TMEM lanes (the warp's quadrant) x 16 consecutive 32-bit columns (no wait: see tmem_st_wait)
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::"r"(taddr),
        "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
        "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
        "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
        "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }

__device__ __forceinline__ float rsqrt_fast(float s) {   // one MUFU.RSQ; s >= beta_eff > 0 here, so no denormal / sign handling
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(s));
    return r;
}

__device__ __forceinline__ float ldg_stream1(const float *p) {
    float r;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}

// weights -> A operand (hi, lo), K-major SWIZZLE_128B, one 128-byte row per output channel; rows >= C are zero
// K column 27 carries the bias: the im2col rows hold 1.0 there (kC0One), so GEMM 1 adds the bias and GEMM 2 returns d(bias) = sum dv
template <int ROWS_W>
__device__ __forceinline__ void stage_weights(const float *__restrict__ w, const float *__restrict__ bias, int C, uint32_t sWh, uint32_t sWl,
                                              int tid) {
    for (int idx = tid; idx < ROWS_W * 8; idx += kC0Threads) {
        const int i = idx >> 3, c = idx & 7;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (i < C) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (4 * c + j < kC0K) v[j] = __ldg(w + (size_t)i * kC0K + 4 * c + j);
                else if (4 * c + j == kC0One && bias != nullptr) v[j] = __ldg(bias + i);
            }
        }
        float4 hi, lo;
        split4(make_float4(v[0], v[1], v[2], v[3]), hi, lo);
        const uint32_t off = sw128_offset(i, 0, c, ROWS_W);
        sts128(sWh + off, hi);
        sts128(sWl + off, lo);
    }
}

template <int C, bool BWD>
__global__ void __launch_bounds__(kC0Threads, 1)
conv0_gdn_kernel(const float *__restrict__ x, const float *__restrict__ w, const float *__restrict__ bias, const float *__restrict__ beta_param,
                 const float *__restrict__ gamma_weight, C0Geom g, float *__restrict__ y, float *__restrict__ v_out,   // forward outputs
                 const float *__restrict__ gy, float *__restrict__ part_dw, float *__restrict__ part_sums) {       // backward in / out
    using Cfg = C0Cfg<C, BWD>;
    constexpr int TN = Cfg::TN, MB = Cfg::MB, NS1 = Cfg::NS1, NS2 = BWD ? Cfg::NS2 : 1, NP = Cfg::NP;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t sWh = (smem_u32(smem_raw) + 1023u) & ~1023u, sWl = sWh + Cfg::W_BYTES;
    const uint32_t sStage1 = sWl + Cfg::W_BYTES;                          // NS1 x { ph, pl }
    const uint32_t sStage2 = sStage1 + NS1 * 2 * Cfg::P_BYTES;            // NS2 x { pht, plt }
    // barriers: full1[NS1], empty1[NS1], full2[NS2], empty2[NS2], tfull[2], tempty[2] (fwd) / dvfull[2] (bwd), dvlo_empty, done
    __shared__ __align__(8) uint64_t bars[2 * NS1 + 2 * NS2 + 6];
    __shared__ uint32_t tmem_base_slot;
    __shared__ float red[BWD ? 4 * MB * 128 * 2 : 1];
    // dW of this CTA, [block][k][channel row]: GEMM 2 starts a FRESH tensor-memory accumulator every tile and the epilogue warps
    // add it in here with round-to-nearest fp32 adds.  One accumulator for the whole kernel (first version) is biased: the tensor
    // core truncates when it adds into a large running sum, and over ~2600 accumulating MMAs per CTA at the 16 x 256^2 site dW came
    // out 5.8e-5 low against the sum over batch slices (r02f, tests/test_gpu_conv0.py::test_full_site_properties).
    __shared__ float sdw[BWD ? MB * kC0KP * 128 : 1];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t bar_full1 = smem_u32(&bars[0]), bar_empty1 = smem_u32(&bars[NS1]);
    const uint32_t bar_full2 = smem_u32(&bars[2 * NS1]), bar_empty2 = smem_u32(&bars[2 * NS1 + NS2]);
    const uint32_t bar_tfull = smem_u32(&bars[2 * NS1 + 2 * NS2]), bar_epi = bar_tfull + 16;   // tempty (fwd) / dvfull (bwd)
    const uint32_t bar_dvlo = bar_epi + 16, bar_done = bar_dvlo + 8;
    stage_weights<Cfg::ROWS_W>(w, bias, C, sWh, sWl, tid);
    if (BWD) {
        for (int i = tid; i < MB * kC0KP * 128; i += kC0Threads) sdw[i] = 0.f;
    }
    if (BWD) {   // rows 27..31 of the transposed tiles are never written by the producers: zero all stages once
        uint4 *p = reinterpret_cast<uint4 *>(smem_raw + (sStage2 - smem_u32(smem_raw)));
        for (uint32_t i = tid; i < NS2 * 2 * Cfg::P_BYTES / 16; i += kC0Threads) p[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    if (tid == 0) {
        for (int s = 0; s < NS1; ++s) {
            mbar_init(bar_full1 + 8 * s, kC0ProdThreads);
            mbar_init(bar_empty1 + 8 * s, 1);
        }
        for (int s = 0; s < NS2; ++s) {
            mbar_init(bar_full2 + 8 * s, kC0ProdThreads);
            mbar_init(bar_empty2 + 8 * s, 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(bar_tfull + 8 * a, 1);
            mbar_init(bar_epi + 8 * a, kC0EpiThreads);
        }
        mbar_init(bar_dvlo, 1);
        mbar_init(bar_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == kC0Epi + kC0Prod) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_slot)), "r"(Cfg::TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = tmem_base_slot;
    const long n_tiles = (g.P + TN - 1) / TN;

    if (warp < kC0Epi) {
        // ===================================================== epilogue
        const int q = warp & 3, cg = warp >> 2;
        float beta[MB], gamma[MB];
        float sum_b[MB], sum_g[MB];
#pragma unroll
        for (int blk = 0; blk < MB; ++blk) {
            const int ch = blk * 128 + q * 32 + lane;
            beta[blk] = 1.f; gamma[blk] = 0.f;
            sum_b[blk] = sum_g[blk] = 0.f;
            if (ch < C) eff_params(beta_param, gamma_weight, ch, beta[blk], gamma[blk]);
        }
        auto drain_dw = [&]() {   // this warp's 8 of the 32 k columns of the per-tile dW accumulator -> += into shared memory
#pragma unroll
            for (int blk = 0; blk < MB; ++blk) {
                float d[8];
                tmem_ld8(tmem_base + ((uint32_t)(q * 32) << 16) + Cfg::COL_DW + blk * kC0KP + cg * 8, d);
#pragma unroll
                for (int j = 0; j < 8; ++j) sdw[(blk * kC0KP + cg * 8 + j) * 128 + q * 32 + lane] += d[j];
            }
        };
        long it = 0;
        for (long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const uint32_t a = (uint32_t)(it & 1), aph = (uint32_t)((it >> 1) & 1);
            const long p0 = tile * TN + cg * NP;                   // first position of this warp's columns
            const long left = g.P - p0;
            const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16);
            float gq[16];
            if (BWD) {                                              // first batch of g requested before the accumulator is awaited
                const int ch = q * 32 + lane;
#pragma unroll
                for (int j = 0; j < 16; ++j) gq[j] = (q * 32 < C && j < left) ? ldg_stream1(gy + (size_t)(p0 + j) * C + ch) : 0.f;
            }
            mbar_wait(bar_tfull + 8 * a, aph);
            fence_after_sync();
            bool lo_free = !BWD || it == 0;
#pragma unroll
            for (int blk = 0; blk < MB; ++blk) {
                const int ch = blk * 128 + q * 32 + lane;
                const bool ok = blk * 128 + q * 32 < C;            // warp-uniform
                const uint32_t col = (uint32_t)(blk * TN + cg * NP);
#pragma unroll 1
                for (int k0 = 0; k0 < NP; k0 += 16) {
                    float acc[16];
                    tmem_ld16(tlane + a * 128 + col + k0, acc);    // warp-collective: every lane loads
                    const bool whole = left >= k0 + 16;            // warp-uniform: all 16 positions of the batch exist
                    if (!BWD) {
                        if (ok) {
                            float *yp = y + (size_t)(p0 + k0) * C + ch;
                            if (whole && v_out == nullptr) {       // the hot path: 6 instructions per element
#pragma unroll
                                for (int j = 0; j < 16; ++j) {
                                    const float v = acc[j];              // conv + bias (the bias came through the ones column)
                                    __stcs(yp + (size_t)j * C, v * rsqrt_fast(fmaf(gamma[blk], v * v, beta[blk])));   // training path, see header
                                }
                            } else {
#pragma unroll
                                for (int j = 0; j < 16; ++j) {
                                    if (k0 + j < left) {
                                        const float v = acc[j];
                                        if (v_out != nullptr) __stcs(v_out + (size_t)(p0 + k0 + j) * C + ch, v);
                                        __stcs(yp + (size_t)j * C, v * rsqrt_fast(fmaf(gamma[blk], v * v, beta[blk])));
                                    }
                                }
                            }
                        }
                    } else {
                        if (blk != 0 || k0 != 0) {
                            const float *gp = gy + (size_t)(p0 + k0) * C + ch;
                            if (ok && whole) {
#pragma unroll
                                for (int j = 0; j < 16; ++j) gq[j] = ldg_stream1(gp + (size_t)j * C);
                            } else {
#pragma unroll
                                for (int j = 0; j < 16; ++j) gq[j] = (ok && k0 + j < left) ? ldg_stream1(gp + (size_t)j * C) : 0.f;
                            }
                        }
                        float lo[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j) {             // GDN backward (gdn_math.cuh gdn_bwd1<false>) with the one-MUFU rsqrt;
                            const float v = acc[j];                 // conv + bias; the constant factor -1/2 of the two sums is applied once, at the end
                            const float x2 = v * v;
                            const float r = rsqrt_fast(fmaf(gamma[blk], x2, beta[blk]));
                            const float gr3 = gq[j] * (r * r * r);  // g / d^3; gq = 0 past the end: no contribution
                            const float t = gr3 * v;
                            sum_b[blk] += t;                        // d(beta)  = -1/2 sum g v / d^3
                            sum_g[blk] = fmaf(t, x2, sum_g[blk]);   // d(gamma) = -1/2 sum g v^3 / d^3
                            const float dv = gr3 * beta[blk];       // d(bias) = sum dv comes out of GEMM 2 (ones column)
                            acc[j] = tf32_hi(dv);
                            lo[j] = dv - acc[j];
                        }
                        if (!lo_free) {                             // GEMM 2 of the previous tile is complete: its dW is final, dv_lo is free
                            mbar_wait(bar_dvlo, (uint32_t)((it - 1) & 1));
                            fence_after_sync();
                            lo_free = true;
                            drain_dw();
                        }
                        tmem_st16(tlane + a * 128 + col + k0, acc);                  // dv_hi over v, in place
                        tmem_st16(tlane + Cfg::COL_DVLO + col + k0, lo);
                    }
                }
            }
            if (BWD) tmem_st_wait();
            fence_before_sync();
            mbar_arrive(bar_epi + 8 * a);
        }
        if (BWD) {
            mbar_wait(bar_done, 0);                                 // GEMM 2 of this CTA's last tile
            fence_after_sync();
            drain_dw();
#pragma unroll
            for (int blk = 0; blk < MB; ++blk) {
                float *r = red + ((cg * MB + blk) * 128 + q * 32 + lane) * 2;
                r[0] = -0.5f * sum_b[blk]; r[1] = -0.5f * sum_g[blk];
            }
        }
    } else if (warp < kC0Epi + kC0Prod) {
        // ===================================================== producers: one position per thread and tile
        const int pp = tid - kC0EpiThreads;
        const bool active = pp < TN;
        float pn[kC0K];
        float one = 0.f;                                            // 1.0 for a position inside the batch, else 0 (its whole row is zero)
        auto request = [&](long t) {
#pragma unroll
            for (int k = 0; k < kC0K; ++k) pn[k] = 0.f;
            const long p = t * TN + pp;
            one = (active && p < g.P) ? 1.f : 0.f;
            if (active && p < g.P) {
                const int hw = g.H * g.W;
                const int b = (int)(p / hw), rem = (int)(p - (long)b * hw);
                const int oy = rem / g.W, ox = rem - oy * g.W;
                const float *img = x + ((size_t)b * hw + (size_t)oy * g.W + ox) * 3;
#pragma unroll
                for (int kh = 0; kh < 3; ++kh) {
                    const bool rok = (unsigned)(oy + kh - 1) < (unsigned)g.H;
#pragma unroll
                    for (int kw = 0; kw < 3; ++kw) {
                        if (rok && (unsigned)(ox + kw - 1) < (unsigned)g.W) {
                            const float *s = img + ((kh - 1) * g.W + (kw - 1)) * 3;
                            pn[(kh * 3 + kw) * 3 + 0] = __ldg(s);
                            pn[(kh * 3 + kw) * 3 + 1] = __ldg(s + 1);
                            pn[(kh * 3 + kw) * 3 + 2] = __ldg(s + 2);
                        }
                    }
                }
            }
        };
        auto prefetch_g = [&](long t) {   // backward: DRAM latency of the epilogue's grad_y reads is taken tiles ahead by one bulk L2 prefetch
            if (BWD && pp == 0 && t < n_tiles) {
                const long rows = (g.P - t * TN < TN) ? (g.P - t * TN) : TN;
                prefetch_l2_bulk(gy + (size_t)t * TN * C, (uint32_t)(rows * C * 4));
            }
        };
        request(blockIdx.x);
        prefetch_g(blockIdx.x);
        prefetch_g(blockIdx.x + (long)gridDim.x);
        prefetch_g(blockIdx.x + 2 * (long)gridDim.x);
        uint32_t u = 0;
        for (long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++u) {
            const uint32_t s1 = u % NS1, ph1 = (u / NS1) & 1;
            prefetch_g(tile + 3 * (long)gridDim.x);
            const uint32_t sPh = sStage1 + s1 * 2 * Cfg::P_BYTES, sPl = sPh + Cfg::P_BYTES;
            mbar_wait(bar_empty1 + 8 * s1, ph1 ^ 1);
            if (active) {
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    float4 v4;
                    v4.x = 4 * c + 0 < kC0K ? pn[(4 * c + 0) % kC0K] : 0.f;
                    v4.y = 4 * c + 1 < kC0K ? pn[(4 * c + 1) % kC0K] : 0.f;
                    v4.z = 4 * c + 2 < kC0K ? pn[(4 * c + 2) % kC0K] : 0.f;
                    v4.w = 4 * c + 3 < kC0K ? pn[(4 * c + 3) % kC0K] : (4 * c + 3 == kC0One ? one : 0.f);
                    float4 hi, lo;
                    split4(v4, hi, lo);
                    const uint32_t off = sw128_offset(pp, 0, c, TN);
                    sts128(sPh + off, hi);
                    sts128(sPl + off, lo);
                }
            }
            fence_proxy_async();
            mbar_arrive(bar_full1 + 8 * s1);
            if (BWD) {
                const uint32_t s2 = u % NS2, ph2 = (u / NS2) & 1;
                const uint32_t sPht = sStage2 + s2 * 2 * Cfg::P_BYTES, sPlt = sPht + Cfg::P_BYTES;
                mbar_wait(bar_empty2 + 8 * s2, ph2 ^ 1);
                if (active) {
#pragma unroll
                    for (int k = 0; k < kC0K; ++k) {   // row k, column = position: lanes write consecutive floats of one swizzled row
                        const float hi = tf32_hi(pn[k]);
                        const uint32_t off = sw128_offset(k, pp >> 5, (pp & 31) >> 2, kC0KP) + (uint32_t)((pp & 3) << 2);
                        sts32(sPht + off, hi);
                        sts32(sPlt + off, pn[k] - hi);
                    }
                    // the ones row (its lo part stays zero from the start-up clear)
                    sts32(sPht + sw128_offset(kC0One, pp >> 5, (pp & 31) >> 2, kC0KP) + (uint32_t)((pp & 3) << 2), one);
                }
                fence_proxy_async();
                mbar_arrive(bar_full2 + 8 * s2);
            }
            request(tile + gridDim.x);
        }
    } else {
        // ===================================================== MMA warp
        const uint32_t idesc1 = idesc_tf32(128, TN), idesc2 = idesc_tf32(128, kC0KP);
        auto gemm1 = [&](uint32_t tmem_v, uint32_t sPh) {            // v = W patch^T: three exact-split terms x four K = 8 steps
#pragma unroll
            for (int blk = 0; blk < MB; ++blk) {
#pragma unroll
                for (int term = 0; term < 3; ++term) {               // wh ph, wl ph, wh pl
                    const uint64_t dA = smem_desc((term == 1 ? sWl : sWh) + blk * 128 * 128);
                    const uint64_t dB = smem_desc(term == 2 ? sPh + Cfg::P_BYTES : sPh);
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        const uint64_t adv = (uint64_t)((ks * 32) >> 4);
                        mma_tf32(tmem_v + blk * TN, dA + adv, dB + adv, idesc1, (uint32_t)((term | ks) != 0));
                    }
                }
            }
        };
        auto gemm2 = [&](uint32_t tmem_dvhi, uint32_t sPht) {   // dW(tile) = dv patch: dvh pht, dvl pht, dvh plt; K = positions; fresh accumulator
#pragma unroll
            for (int blk = 0; blk < MB; ++blk) {
#pragma unroll
                for (int term = 0; term < 3; ++term) {
                    const uint32_t ta = (term == 1 ? tmem_base + Cfg::COL_DVLO : tmem_dvhi) + blk * TN;
                    const uint32_t sb = term == 2 ? sPht + Cfg::P_BYTES : sPht;
#pragma unroll
                    for (int ks = 0; ks < TN / 8; ++ks) {
                        const uint64_t dB = smem_desc(sb + (ks >> 2) * (kC0KP * 128) + (ks & 3) * 32);
                        mma_tf32_ts(tmem_base + Cfg::COL_DW + blk * kC0KP, ta + ks * 8, dB, idesc2, (uint32_t)((term | ks) != 0));
                    }
                }
            }
        };
        long n_local = 0;
        for (long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) ++n_local;
        for (long it = 0; it < n_local + (BWD ? 1 : 0); ++it) {
            const uint32_t a = (uint32_t)(it & 1), aph = (uint32_t)((it >> 1) & 1);
            if (it < n_local) {
                const uint32_t s1 = (uint32_t)(it % NS1), ph1 = (uint32_t)((it / NS1) & 1);
                if (!BWD) mbar_wait(bar_epi + 8 * a, aph ^ 1);       // forward: the epilogue has drained this accumulator stage
                // backward: v[a] was last read (as dv_hi) by GEMM 2 of tile it-2, issued by this thread one iteration ago: in order
                mbar_wait(bar_full1 + 8 * s1, ph1);
                fence_after_sync();
                if (elect_one_sync()) {
                    gemm1(tmem_base + a * 128, sStage1 + s1 * 2 * Cfg::P_BYTES);
                    mma_commit(bar_empty1 + 8 * s1);
                    mma_commit(bar_tfull + 8 * a);
                }
                __syncwarp();
            }
            if (BWD && it >= 1) {
                const long jt = it - 1;
                const uint32_t b = (uint32_t)(jt & 1), bph = (uint32_t)((jt >> 1) & 1);
                const uint32_t s2 = (uint32_t)(jt % NS2), ph2 = (uint32_t)((jt / NS2) & 1);
                mbar_wait(bar_epi + 8 * b, bph);                     // dv_hi / dv_lo of tile jt are in tensor memory
                mbar_wait(bar_full2 + 8 * s2, ph2);
                fence_after_sync();
                if (elect_one_sync()) {
                    gemm2(tmem_base + b * 128, sStage2 + s2 * 2 * Cfg::P_BYTES);
                    mma_commit(bar_empty2 + 8 * s2);
                    mma_commit(bar_dvlo);
                    if (jt == n_local - 1) mma_commit(bar_done);
                }
                __syncwarp();
            }
        }
    }
    fence_before_sync();
    __syncthreads();
    if (BWD) {
        // ===================================================== read-out: per-CTA partials
        for (int i = tid; i < MB * 128 * 2; i += kC0Threads) {      // the two per-channel sums: fixed order over the four column groups
            const int row = i / 2, which = i - row * 2;               // row = blk * 128 + channel-in-block = channel
            if (row < C) {
                float sacc = 0.f;
#pragma unroll
                for (int c4 = 0; c4 < 4; ++c4) sacc += red[((c4 * MB + row / 128) * 128 + (row & 127)) * 2 + which];
                part_sums[(size_t)blockIdx.x * 2 * C + (size_t)which * C + row] = sacc;
            }
        }
        for (int i = tid; i < MB * 128 * kC0KP; i += kC0Threads) {
            const int ch = i / kC0KP, k = i - ch * kC0KP;             // ch = blk * 128 + row
            if (ch < C) part_dw[((size_t)blockIdx.x * C + ch) * kC0KP + k] = sdw[((ch / 128) * kC0KP + k) * 128 + (ch & 127)];
        }
    }
    if (warp == kC0Epi + kC0Prod) {
        fence_after_sync();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(Cfg::TMEM_COLS));
    }
}

// fixed-order fold of the per-CTA partials (binary64) + chain rule to the stored parameters (layers.py:20-21):
//   dW[c][k] = sum parts (k < 27);  dbias[c] = sum parts of the ones column;  dbeta_param = 2 beta_param sum h;  dgamma_weight = 2 w sum h v^2
constexpr int kC0FinLanes = 8;      // lanes sharing one output: each folds a contiguous eighth of the per-CTA partials, then a fixed xor tree
__global__ void __launch_bounds__(256) conv0_gdn_bwd_finalize_kernel(const float *__restrict__ part_dw, const float *__restrict__ part_sums,
                                                                   int n_part, int C, const float *__restrict__ beta_param,
                                                                   const float *__restrict__ gamma_weight, float *__restrict__ dw,
                                                                   float *__restrict__ dbias, float *__restrict__ dbeta,
                                                                   float *__restrict__ dgamma) {
    // One thread per output walked the 148 partials as 19 dependent batches of loads (14 us for 4.6 K outputs on 15 CTAs: pure
    // latency).  Eight lanes per output: 3 batches each, 120 CTAs.
    const int t = blockIdx.x * 256 + threadIdx.x;
    const int e = t / kC0FinLanes, sub = t % kC0FinLanes;
    const int n_dw = C * (kC0K + 1);
    const int per = (n_part + kC0FinLanes - 1) / kC0FinLanes;
    const int p0 = sub * per, p1 = min(p0 + per, n_part);
    const bool is_dw = e < n_dw, is_sum = !is_dw && e < n_dw + 2 * C;
    const float *src = part_sums;
    size_t step = 0;
    if (is_dw) {
        const int c = e / (kC0K + 1), k = e - c * (kC0K + 1);       // k == 27: the ones column = d(bias)
        src = part_dw + (size_t)c * kC0KP + k;
        step = (size_t)C * kC0KP;
    } else if (is_sum) {
        src = part_sums + (e - n_dw);
        step = (size_t)2 * C;
    }
    double s = 0.0;
    if (is_dw || is_sum) {
        int p = p0;
        for (; p + 8 <= p1; p += 8) {                               // eight independent loads in flight
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = __ldcs(src + (size_t)(p + u) * step);
#pragma unroll
            for (int u = 0; u < 8; ++u) s += (double)v[u];
        }
        for (; p < p1; ++p) s += (double)__ldcs(src + (size_t)p * step);
    }
#pragma unroll
    for (int o = 1; o < kC0FinLanes; o <<= 1) s += __shfl_xor_sync(0xffffffffu, s, o);      // same tree on every lane: deterministic
    if (sub != 0) return;
    if (is_dw) {
        const int c = e / (kC0K + 1), k = e - c * (kC0K + 1);
        if (k == kC0One) {
            if (dbias != nullptr) dbias[c] = (float)s;
        } else {
            dw[c * kC0K + k] = (float)s;
        }
    } else if (is_sum) {
        const int r = e - n_dw, which = r / C, c = r - which * C;
        if (which == 0) dbeta[c] = (float)(2.0 * (double)beta_param[c] * s);
        else dgamma[c] = (float)(2.0 * (double)gamma_weight[c] * s);
    }
}

template <int C, bool BWD>
inline int c0_grid(long P) {
    const long n_tiles = (P + C0Cfg<C, BWD>::TN - 1) / C0Cfg<C, BWD>::TN;
    return (int)(n_tiles < sm_count() ? n_tiles : sm_count());
}

template <int C, bool BWD>
int c0_reserve(const char *what) {
    cudaError_t e = cudaFuncSetAttribute(conv0_gdn_kernel<C, BWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C0Cfg<C, BWD>::SMEM);
    if (e != cudaSuccess) {
        set_error("%s: cannot reserve %zu B of shared memory: %s", what, C0Cfg<C, BWD>::SMEM, cudaGetErrorString(e));
        return (int)e;
    }
    return 0;
}

template <int C>
int launch_c0_fwd(const float *x, const float *w, const float *bias, const float *beta_param, const float *gamma_weight, C0Geom g, float *y,
                  float *v_out, cudaStream_t st) {
    if (int rc = c0_reserve<C, false>("sic_conv0_gdn_fwd")) return rc;
    conv0_gdn_kernel<C, false><<<c0_grid<C, false>(g.P), kC0Threads, C0Cfg<C, false>::SMEM, st>>>(x, w, bias, beta_param, gamma_weight, g, y, v_out,
                                                                                                nullptr, nullptr, nullptr);
    SIC_CHECK_LAUNCH("sic_conv0_gdn_fwd");
    return 0;
}

template <int C>
int launch_c0_bwd(const float *x, const float *w, const float *bias, const float *beta_param, const float *gamma_weight, const float *gy,
                  C0Geom g, float *dw, float *dbias, float *dbeta, float *dgamma, float *ws, cudaStream_t st) {
    if (int rc = c0_reserve<C, true>("sic_conv0_gdn_bwd")) return rc;
    const int grid = c0_grid<C, true>(g.P);
    float *part_dw = ws, *part_sums = ws + (size_t)grid * C * kC0KP;
    conv0_gdn_kernel<C, true><<<grid, kC0Threads, C0Cfg<C, true>::SMEM, st>>>(x, w, bias, beta_param, gamma_weight, g, nullptr, nullptr, gy,
                                                                             part_dw, part_sums);
    SIC_CHECK_LAUNCH("sic_conv0_gdn_bwd");
    const int n_out = C * (kC0K + 1) + 2 * C;
    conv0_gdn_bwd_finalize_kernel<<<(n_out * kC0FinLanes + 255) / 256, 256, 0, st>>>(part_dw, part_sums, grid, C, beta_param, gamma_weight, dw, dbias, dbeta,
                                                                     dgamma);
    SIC_CHECK_LAUNCH("sic_conv0_gdn_bwd (finalize)");
    return 0;
}

SIC_REGISTER_KERNEL("conv0_gdn_kernel<128,fwd>", conv0_gdn_kernel<128, false>);
SIC_REGISTER_KERNEL("conv0_gdn_kernel<128,bwd>", conv0_gdn_kernel<128, true>);
SIC_REGISTER_KERNEL("conv0_gdn_kernel<192,fwd>", conv0_gdn_kernel<192, false>);
SIC_REGISTER_KERNEL("conv0_gdn_kernel<192,bwd>", conv0_gdn_kernel<192, true>);

inline int c0_tn(int C) { return C > 128 ? 64 : 128; }

}  // namespace
}  // namespace sic

using namespace sic;

#define SIC_C0_DISPATCH(C, CALL, what)                                                      \
    switch (C) {                                                                            \
        case 32: return CALL(32);                                                           \
        case 64: return CALL(64);                                                           \
        case 96: return CALL(96);                                                           \
        case 128: return CALL(128);                                                         \
        case 192: return CALL(192);                                                         \
        default:                                                                            \
            set_error(what ": C=%d unsupported (C in {32,64,96,128,192})", C);              \
            return SIC_E_UNSUPPORTED;                                                       \
    }

extern "C" int sic_conv0_gdn_fwd(const float *x, const float *w, const float *bias, const float *beta_param, const float *gamma_weight,
                                 int B, int H, int W, int C, float *y, float *v_out, void *stream) {
    SIC_CHECK_ARG(B > 0 && H > 0 && W > 0 && C > 0, "sic_conv0_gdn_fwd: empty shape B=%d H=%d W=%d C=%d", B, H, W, C);
    SIC_CHECK_ARG(x && w && beta_param && gamma_weight && y, "sic_conv0_gdn_fwd: null pointer");
    SIC_CHECK_ARG((long)H * W < (1L << 30), "sic_conv0_gdn_fwd: image too large");
    C0Geom g{B, H, W, (long)B * H * W};
    cudaStream_t st = (cudaStream_t)stream;
#define SIC_C0_FWD(CC) launch_c0_fwd<CC>(x, w, bias, beta_param, gamma_weight, g, y, v_out, st)
    SIC_C0_DISPATCH(C, SIC_C0_FWD, "sic_conv0_gdn_fwd")
#undef SIC_C0_FWD
}

extern "C" size_t sic_conv0_gdn_bwd_workspace_bytes(int B, int H, int W, int C) {
    if (B <= 0 || H <= 0 || W <= 0 || C <= 0) return 0;
    const long P = (long)B * H * W, tn = c0_tn(C), n_tiles = (P + tn - 1) / tn;
    const long grid = n_tiles < sm_count() ? n_tiles : sm_count();
    return (size_t)grid * C * (kC0KP + 3) * sizeof(float);
}

extern "C" int sic_conv0_gdn_bwd(const float *x, const float *w, const float *bias, const float *beta_param, const float *gamma_weight,
                                 const float *grad_y, int B, int H, int W, int C, float *dw, float *dbias, float *dbeta_param,
                                 float *dgamma_weight, void *workspace, size_t workspace_bytes, void *stream) {
    SIC_CHECK_ARG(B > 0 && H > 0 && W > 0 && C > 0, "sic_conv0_gdn_bwd: empty shape B=%d H=%d W=%d C=%d", B, H, W, C);
    SIC_CHECK_ARG(x && w && beta_param && gamma_weight && grad_y && dw && dbeta_param && dgamma_weight && workspace,
                  "sic_conv0_gdn_bwd: null pointer");
    SIC_CHECK_ARG((long)H * W < (1L << 30), "sic_conv0_gdn_bwd: image too large");
    SIC_CHECK_ARG(((uintptr_t)workspace & 15) == 0, "sic_conv0_gdn_bwd: workspace must be 16-byte aligned");
    if (workspace_bytes < sic_conv0_gdn_bwd_workspace_bytes(B, H, W, C)) {
        set_error("sic_conv0_gdn_bwd: workspace %zu < %zu bytes", workspace_bytes, sic_conv0_gdn_bwd_workspace_bytes(B, H, W, C));
        return SIC_E_WORKSPACE;
    }
    C0Geom g{B, H, W, (long)B * H * W};
    cudaStream_t st = (cudaStream_t)stream;
    float *ws = static_cast<float *>(workspace);
#define SIC_C0_BWD(CC) launch_c0_bwd<CC>(x, w, bias, beta_param, gamma_weight, grad_y, g, dw, dbias, dbeta_param, dgamma_weight, ws, st)
    SIC_C0_DISPATCH(C, SIC_C0_BWD, "sic_conv0_gdn_bwd")
#undef SIC_C0_BWD
}
