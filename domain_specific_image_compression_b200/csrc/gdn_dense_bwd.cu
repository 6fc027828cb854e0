// G3 backward: dense-gamma GDN / IGDN gradients on the pipelined tcgen05 kernel of gdn_dense_ws.cu (SURVEY.md 8(a') G3).
//
//   s_i = beta_i + sum_j gamma_ij x_j^2,  d = sqrt(s)
//   GDN :  h_i = -1/2 g_i x_i / d_i^3,  dx_j = g_j / d_j + 2 x_j sum_i gamma_ij h_i
//   IGDN:  h_i = +1/2 g_i x_i / d_i,    dx_j = g_j d_j  + 2 x_j sum_i gamma_ij h_i
//   d(beta_eff)_i = sum_p h_i ,  d(gamma_eff)_ij = sum_p h_i x_j^2   (the latter: gdn_dense_dgamma.cu, on the emitted h)
//
// Two launches of the same warp-specialised producer / MMA / epilogue pipeline as the forward (identical barriers, descriptors,
// stage ring and TMEM layout; only the producer's pre-op, the orientation of gamma and the epilogue differ):
//   pass 1  MMA as in the forward (s); epilogue emits h and direct = g/d | g d and accumulates sum_p h per lane (= per channel)
//   pass 2  gamma TRANSPOSED into the A operand, h split hi/lo as the B operand (t = gamma^T h); epilogue dx = direct + 2 x t
// Algorithmic traffic: pass 1 reads x, g and writes h, direct; pass 2 reads h, x, direct and writes dx: 32 B/element, against
// ~100 B/element for the elementwise + cuBLAS chain through torch that it replaces.
//
// STATUS: the default backward of GDN(dense=True) since round 2 (tests/test_gpu_gdn.py: 10 shapes vs float64, the whole model with all
// 13 sites dense); d(gamma) is the third launch, csrc/gdn_dense_dgamma.cu.  Measured: profiles/r02z_kernel_bench_dense.json,
// profiles/r02d_ncu_kernels.txt (both passes run at 5.8 - 6.5 TB/s of DRAM traffic: the cost is the 40 B/element the three passes move).
#include "gdn_dense_ws.cuh"

namespace sic {
namespace {

using namespace umma;
using namespace dense_ws;

// What the pipeline computes.  The producer / MMA / barrier machinery is the same for all five; they differ in the producer's
// pre-op, in the orientation of gamma and in the epilogue:
//   OP_FWD / OP_FWD_INV    s = beta + gamma x^2 ;  y = x / sqrt(s)  |  x sqrt(s)
//   OP_BWD1 / OP_BWD1_INV  the same s, then (SURVEY 8(a') G3)  GDN:  direct = g / d,  h = -1/2 g x / d^3
//                                                             IGDN: direct = g d,    h = +1/2 g x / d      ; also sum_p h per channel
//   OP_BWD2                t = gamma^T h  (gamma transposed into the A operand, h split hi/lo like x^2) ;  dx = direct + 2 x t
enum { OP_FWD = 0, OP_FWD_INV = 1, OP_BWD1 = 2, OP_BWD1_INV = 3, OP_BWD2 = 4 };   // only the OP_BWD* ones are instantiated here
template <int OP>
struct OpTraits {
    static constexpr bool kInverse = OP == OP_FWD_INV || OP == OP_BWD1_INV;
    static constexpr bool kFwd = OP == OP_FWD || OP == OP_FWD_INV;
    static constexpr bool kBwd1 = OP == OP_BWD1 || OP == OP_BWD1_INV;
    static constexpr bool kBwd2 = OP == OP_BWD2;
    static_assert(OP >= OP_FWD && OP <= OP_BWD2 && (kInverse || !kInverse), "unknown op");
};

// epilogue tensors:  forward a = x, o0 = y;   pass 1: a = x, b = g, o0 = h, o1 = direct;   pass 2: a = x, b = direct, o0 = dx
struct EpiPtrs {
    const float *a, *b;
    float *o0, *o1;
};

// Epilogue of one warp for one tile: NP positions of this lane's output channel (block A), then of its block-B channel when
// C = 192.  The inputs are fetched in batches of 16 positions through a 2-deep register ring: the first two batches are requested
// before the accumulator barrier is awaited (their L2 latency hides behind the MMA), every later batch while the two before it
// are being processed and stored.  In the forward, 32 + 16 live values keep the role inside the 96-register budget of a 17-warp
// CTA without spills (spill traffic shares the L1 data pipe with the tensor core's operand reads, this kernel's critical resource).
template <int C, int OP, bool FULL>
__device__ __forceinline__ void epilogue_tile(const EpiPtrs &p, long p0, int left, int cA, bool okA, float betaA, int cB, bool okB,
                                              float betaB, uint32_t taddr, uint32_t bar, uint32_t parity, float &sumA, float &sumB) {
    using Cfg = WsCfg<C>;
    using T = OpTraits<OP>;
    constexpr int NP = Cfg::NP;
    static_assert(NP % 16 == 0, "epilogue batches are 16 columns");
    constexpr int NBB = NP / 16;                                 // batches per block
    constexpr int NB = Cfg::kTwoBlocks ? 2 * NBB : NBB;
    const long offA = p0 * C + cA, offB = p0 * C + cB;
    float ring[2][16];
    float ring2[T::kFwd ? 1 : 2][T::kFwd ? 1 : 16];              // second input (g or direct) of the backward passes
    auto load = [&](int slot, int b) {
        const bool blkB = b >= NBB;
        const int k0 = (blkB ? b - NBB : b) * 16;
        const long off = blkB ? offB : offA;
        const bool ok = blkB ? okB : okA;
#pragma unroll
        for (int j = 0; j < 16; ++j) ring[slot][j] = (ok && (FULL || k0 + j < left)) ? __ldcg(p.a + off + (long)(k0 + j) * C) : 0.f;
        if constexpr (!T::kFwd) {
#pragma unroll
            for (int j = 0; j < 16; ++j) ring2[slot][j] = (ok && (FULL || k0 + j < left)) ? __ldcg(p.b + off + (long)(k0 + j) * C) : 0.f;
        }
    };
    load(0, 0);
    if (NB > 1) load(1, 1);
    mbar_wait(bar, parity);
    fence_after_sync();
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        const bool blkB = b >= NBB;
        const int k0 = (blkB ? b - NBB : b) * 16;
        const long off = blkB ? offB : offA;
        const bool ok = blkB ? okB : okA;
        const float beta = blkB ? betaB : betaA;
        float acc[16];
        tmem_ld16(taddr + (blkB ? kColsB : 0) + k0, acc);
        float hsum = 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            if (ok && (FULL || k0 + j < left)) {
                const long e = off + (long)(k0 + j) * C;
                const float xv = ring[b & 1][j];
                if constexpr (T::kFwd) {
                    __stcs(p.o0 + e, xv * norm_factor<T::kInverse>(beta + acc[j]));
                } else if constexpr (T::kBwd1) {
                    const float gv = ring2[b & 1][j];
                    const float s = beta + acc[j];
                    const float r = norm_factor<false>(s);        // 1/sqrt(s)
                    float direct, h;
                    if (T::kInverse) {
                        direct = gv * (s * r);                   // g d
                        h = 0.5f * gv * xv * r;                  // 1/2 g x / d
                    } else {
                        direct = gv * r;                         // g / d
                        h = -0.5f * direct * xv * (r * r);       // -1/2 g x / d^3
                    }
                    __stcs(p.o0 + e, h);
                    __stcs(p.o1 + e, direct);
                    hsum += h;
                } else {
                    __stcs(p.o0 + e, fmaf(2.0f * xv, acc[j], ring2[b & 1][j]));   // dx = direct + 2 x t
                }
            }
        }
        if constexpr (T::kBwd1) {
            if (blkB) sumB += hsum;
            else sumA += hsum;
        }
        if (b + 2 < NB) load(b & 1, b + 2);
    }
}

// in0: the tensor the producers feed to the MMA (x; h in pass 2).  out0 / in1 / in2 / out1 / part: see EpiPtrs and OP_*.
template <int C, int OP>
__global__ void __launch_bounds__(kThreadsWS, 1) gdn_dense_ws_kernel(const float *__restrict__ in0, const float *__restrict__ beta_param,
                                                                     const float *__restrict__ gamma_param, long P, int use_prefetch,
                                                                     int gamma_in_tmem, float *__restrict__ out0,
                                                                     const float *__restrict__ in1, const float *__restrict__ in2,
                                                                     float *__restrict__ out1, float *__restrict__ part) {
    using Cfg = WsCfg<C>;
    using T = OpTraits<OP>;
    constexpr int V = C / 4;                         // float4 per position
    constexpr int TN = Cfg::TN, NS = Cfg::NS, ROWS_G = Cfg::ROWS_G, NP = Cfg::NP, KS = Cfg::KS, KBS = Cfg::KBS;
    constexpr int VH = Cfg::VH, PER = Cfg::PER, ACC_COLS = Cfg::ACC_COLS;
    constexpr uint32_t B_BYTES = Cfg::B_BYTES;
    extern __shared__ uint8_t smem_raw[];
    // 32-bit shared-window addresses throughout (st.shared, descriptors); SWIZZLE_128B operands need 1024-byte alignment
    const uint32_t sG = (smem_u32(smem_raw) + 1023u) & ~1023u;   // [ROWS_G x C]  gamma, re-parameterised
    const uint32_t sStage = sG + Cfg::G_BYTES;                   // NS x { hi [TN x 32 KBS], lo [TN x 32 KBS] }
    __shared__ __align__(8) uint64_t bars[2 * NS + 4];           // full[NS], empty[NS], tmem_full[2], tmem_empty[2]
    __shared__ uint32_t tmem_base_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t bar_full = smem_u32(&bars[0]), bar_empty = smem_u32(&bars[NS]);
    const uint32_t bar_tfull = smem_u32(&bars[2 * NS]), bar_tempty = smem_u32(&bars[2 * NS + 2]);

    // ---- one-time setup: gamma -> A operand (layers.py:21 applied to the C x C matrix), barriers, TMEM
    for (int idx = tid; idx < ROWS_G * V; idx += kThreadsWS) {
        const int i = idx / V, c4 = idx - i * V;
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < C) {
            if constexpr (T::kBwd2) {                // operand row i, K index k holds gamma[k][i]  (t_i = sum_k gamma_ki h_k)
                const float *col = gamma_param + (size_t)(4 * c4) * C + i;
                g = make_float4(__ldg(col), __ldg(col + C), __ldg(col + 2 * C), __ldg(col + 3 * C));
            } else {
                g = __ldg(reinterpret_cast<const float4 *>(gamma_param) + (size_t)i * V + c4);
            }
            g.x = g.x * g.x - kReparamOffset; g.y = g.y * g.y - kReparamOffset;
            g.z = g.z * g.z - kReparamOffset; g.w = g.w * g.w - kReparamOffset;
        }
        sts128(sG + sw128_offset(i, c4 >> 3, c4 & 7, ROWS_G), g);
    }
    if (tid == 0) {
        for (int s = 0; s < NS; ++s) {
            mbar_init(bar_full + 8 * s, kProdThreads);
            mbar_init(bar_empty + 8 * s, 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(bar_tfull + 8 * a, 1);
            mbar_init(bar_tempty + 8 * a, kEpiThreads);
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    // TMEM: two accumulator stages, plus (one-block shapes, gamma_in_tmem) a copy of gamma as the A operand of a
    // tensor-memory-sourced MMA: the tensor core then reads only the x^2 tile from shared memory, which halves its share of the
    // L1/shared data pipe (ncu: that pipe, shared with the producers' stores and the epilogue's accesses, is the busiest unit)
    constexpr uint32_t TMEM_COLS = Cfg::kTwoBlocks ? 2 * ACC_COLS : 512;
    const bool a_tmem = !Cfg::kTwoBlocks && !T::kBwd2 && gamma_in_tmem != 0;
    if (warp == kEpiWarps + kProdWarps) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_slot)), "r"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    fence_proxy_async();                             // gamma was written by the generic proxy, the MMA reads through the async proxy
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = tmem_base_slot;
    const uint32_t tmem_gamma = tmem_base + 2 * ACC_COLS;   // lane = output channel, column = input channel
    if (a_tmem) {
        if (warp < 4) {                              // one warp per TMEM lane quadrant: row = 32 * warp + lane
            const int row = warp * 32 + lane;
#pragma unroll 1
            for (int c0 = 0; c0 < C; c0 += 32) {
                float g[32];
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4) {
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (row < C) {
                        v = __ldg(reinterpret_cast<const float4 *>(gamma_param + (size_t)row * C + c0) + j4);
                        v.x = v.x * v.x - kReparamOffset; v.y = v.y * v.y - kReparamOffset;
                        v.z = v.z * v.z - kReparamOffset; v.w = v.w * v.w - kReparamOffset;
                    }
                    g[4 * j4] = v.x; g[4 * j4 + 1] = v.y; g[4 * j4 + 2] = v.z; g[4 * j4 + 3] = v.w;
                }
                tmem_st32(tmem_gamma + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, g);
            }
        }
        fence_before_sync();
        __syncthreads();
        fence_after_sync();
    }
    const long n_tiles = (P + TN - 1) / TN;

    if (warp < kEpiWarps) {
        // ===================================================== epilogue: TMEM -> y
        const int q = warp & 3;                      // TMEM lane quadrant this warp may read (hardware rule: warp id % 4)
        const int col0 = (warp >> 2) * NP;           // its half of the TN position columns
        const int cA = q * 32 + lane;                // block A (M = 128): accumulator row = lane of the quadrant
        const bool okA = q * 32 < C;                 // warp-uniform (C is a multiple of 32); idle warps only keep the barriers in step
        const int cB = 128 + q * 16 + (lane & 15);   // block B (M = 64): row j lives in lane 32*(j/16) + j%16
        const bool okB = Cfg::kTwoBlocks && lane < 16;
        float betaA = 1.f, betaB = 1.f;
        if (okA) {
            const float b = __ldg(beta_param + cA);
            betaA = b * b - kReparamOffset;          // layers.py:20
        }
        if (okB) {
            const float b = __ldg(beta_param + cB);
            betaB = b * b - kReparamOffset;
        }
        EpiPtrs ep;
        ep.a = T::kBwd2 ? in1 : in0;                 // x
        ep.b = T::kBwd2 ? in2 : in1;                 // pass 1: g, pass 2: direct (unused in the forward)
        ep.o0 = out0;
        ep.o1 = out1;
        float sumA = 0.f, sumB = 0.f;                // pass 1: sum over this warp's positions of h for the lane's channel(s)
        long it = 0;
        for (long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const uint32_t a = (uint32_t)(it & 1), aph = (uint32_t)((it >> 1) & 1);
            const long p0 = tile * TN + col0;
            const long left = P - p0;                // positions of this half that exist (may be <= 0 on the last tile)
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + a * ACC_COLS + (uint32_t)col0;
            if (!okA || left <= 0) {
                mbar_wait(bar_tfull + 8 * a, aph);
            } else if (left >= NP) {
                epilogue_tile<C, OP, true>(ep, p0, NP, cA, okA, betaA, cB, okB, betaB, taddr, bar_tfull + 8 * a, aph, sumA, sumB);
            } else {
                epilogue_tile<C, OP, false>(ep, p0, (int)left, cA, okA, betaA, cB, okB, betaB, taddr, bar_tfull + 8 * a, aph, sumA, sumB);
            }
            fence_before_sync();
            mbar_arrive(bar_tempty + 8 * a);         // accumulator stage a may be overwritten
        }
        if constexpr (T::kBwd1) {                    // d(beta_eff) partials: one row per (CTA, position half), folded by the caller
            float *row = part + ((size_t)blockIdx.x * 2 + (size_t)(warp >> 2)) * C;
            if (okA) row[cA] = sumA;
            if (okB) row[cB] = sumB;
        }
    } else if (warp < kEpiWarps + kProdWarps) {
        // ===================================================== producer: x -> x^2 (hi, lo) -> smem, one K sub-step at a time
        const int ptid = tid - kEpiThreads;
        float4 xn[PER];
        auto request = [&](long t, int h) {          // channels [h*32*KBS, (h+1)*32*KBS) of the TN positions of tile t -> registers
            const long q0 = t * TN;
            const long vld = P - q0;                 // <= 0 past the end
            const float4 *src = reinterpret_cast<const float4 *>(in0 + q0 * C) + h * VH;
#pragma unroll
            for (int k = 0; k < PER; ++k) {
                const int idx = ptid + k * kProdThreads, r = idx / VH, c4 = idx - r * VH;
                xn[k] = (r < vld) ? ldg_keep(src + (long)r * V + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        // DRAM latency of the producers' tensor is taken by an L2 prefetch two tiles ahead (one thread, one bulk instruction, no
        // registers); the register loads one sub-step ahead then hit L2.  The tensors only the EPILOGUE reads (g in pass 1; x and
        // direct in pass 2) are prefetched for the tile the producers are just starting: the epilogue of that tile runs one tile
        // time (~3 us) later.  Round 2's first capture prefetched them two tiles ahead as well: 57-85 MB of lines in flight across
        // 148 CTAs, L2 read hit rate 29 % / 19 % and twice the algorithmic DRAM reads (profiles/r02c_ncu_dense_bwd.txt) - the lines
        // were evicted before use.  SIC_DENSE_BWD_PF=far restores that behaviour for A/B timing.
        auto prefetch = [&](long t) {
            const long q0 = t * TN;
            if (use_prefetch && ptid == 0 && q0 < P) {
                const long rows = (P - q0 < TN) ? (P - q0) : TN;
                prefetch_l2_bulk(in0 + q0 * C, (uint32_t)(rows * C * 4));
                if (use_prefetch == 2) {
                    if constexpr (!T::kFwd) prefetch_l2_bulk(in1 + q0 * C, (uint32_t)(rows * C * 4));
                    if constexpr (T::kBwd2) prefetch_l2_bulk(in2 + q0 * C, (uint32_t)(rows * C * 4));
                }
            }
        };
        auto prefetch_epilogue = [&](long t) {
            const long q0 = t * TN;
            if (use_prefetch == 1 && ptid == 0 && q0 < P) {
                const long rows = (P - q0 < TN) ? (P - q0) : TN;
                if constexpr (!T::kFwd) prefetch_l2_bulk(in1 + q0 * C, (uint32_t)(rows * C * 4));    // g (pass 1) / x (pass 2)
                if constexpr (T::kBwd2) prefetch_l2_bulk(in2 + q0 * C, (uint32_t)(rows * C * 4));   // direct
            }
        };
        request(blockIdx.x, 0);
        prefetch(blockIdx.x + (long)gridDim.x);
        uint32_t u = 0;                              // sub-step counter: stage = u % NS, phase = (u / NS) & 1
        for (long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            prefetch_epilogue(tile);
#pragma unroll 1
            for (int h = 0; h < KS; ++h, ++u) {
                const uint32_t s = u % NS, ph = (u / NS) & 1;
                const uint32_t sHi = sStage + s * 2 * B_BYTES, sLo = sHi + B_BYTES;
                mbar_wait(bar_empty + 8 * s, ph ^ 1);    // the MMAs that read this stage have completed
#pragma unroll
                for (int k = 0; k < PER; ++k) {
                    const int idx = ptid + k * kProdThreads, r = idx / VH, c4 = idx - r * VH;
                    const float4 v = xn[k];
                    const float4 sq = T::kBwd2 ? v : make_float4(v.x * v.x, v.y * v.y, v.z * v.z, v.w * v.w);   // pass 2 contracts h itself
                    float4 hi, lo;
                    hi.x = __uint_as_float(__float_as_uint(sq.x) & 0xFFFFE000u); lo.x = sq.x - hi.x;
                    hi.y = __uint_as_float(__float_as_uint(sq.y) & 0xFFFFE000u); lo.y = sq.y - hi.y;
                    hi.z = __uint_as_float(__float_as_uint(sq.z) & 0xFFFFE000u); lo.z = sq.z - hi.z;
                    hi.w = __uint_as_float(__float_as_uint(sq.w) & 0xFFFFE000u); lo.w = sq.w - hi.w;
                    const uint32_t off = sw128_offset(r, c4 >> 3, c4 & 7, TN);
                    sts128(sHi + off, hi);
                    sts128(sLo + off, lo);
                }
                fence_proxy_async();                 // generic-proxy writes -> visible to the tensor core's async proxy
                mbar_arrive(bar_full + 8 * s);
                if (h + 1 < KS) {
                    request(tile, h + 1);            // next K sub-step of this tile
                } else {
                    request(tile + gridDim.x, 0);    // next tile (prefetched into L2 one tile ago)
                    prefetch(tile + 2 * (long)gridDim.x);
                }
            }
        }
    } else {
        // ===================================================== MMA warp: every lane follows the barriers, one elected lane issues
        const uint32_t idescA = idesc_tf32(128, TN), idescB = idesc_tf32(64, TN);
        const uint64_t descGA = smem_desc(sG), descGB = smem_desc(sG + (128 / 8) * 1024);   // block B starts at gamma row 128
        long it = 0;
        uint32_t u = 0;
        for (long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const uint32_t a = (uint32_t)(it & 1), aph = (uint32_t)((it >> 1) & 1);
            const uint32_t tmem_d = tmem_base + a * ACC_COLS;
            mbar_wait(bar_tempty + 8 * a, aph ^ 1);  // the epilogue has drained this accumulator stage
#pragma unroll
            for (int h = 0; h < KS; ++h, ++u) {
                const uint32_t s = u % NS, ph = (u / NS) & 1;
                const uint32_t sHi = sStage + s * 2 * B_BYTES;
                const uint64_t descHi = smem_desc(sHi), descLo = smem_desc(sHi + B_BYTES);
                mbar_wait(bar_full + 8 * s, ph);     // x^2 of this sub-step is in shared memory
                fence_after_sync();
                if (elect_one_sync()) {
#pragma unroll
                    for (int pass = 0; pass < 2; ++pass) {
                        const uint64_t dB = pass == 0 ? descHi : descLo;
#pragma unroll
                        for (int kb = 0; kb < KBS; ++kb) {
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks) {     // 4 x (8 tf32 = 32 B) inside one 128-byte swizzle row
                                const uint64_t advG = (uint64_t)((((h * KBS + kb) * (ROWS_G * 128)) + ks * 32) >> 4);
                                const uint64_t advB = (uint64_t)((kb * (TN * 128) + ks * 32) >> 4);
                                const uint32_t accumulate = (h | pass | kb | ks) != 0;
                                if (a_tmem) mma_tf32_ts(tmem_d, tmem_gamma + (uint32_t)((h * KBS + kb) * 32 + ks * 8), dB + advB, idescA, accumulate);
                                else mma_tf32(tmem_d, descGA + advG, dB + advB, idescA, accumulate);
                                if (Cfg::kTwoBlocks) mma_tf32(tmem_d + kColsB, descGB + advG, dB + advB, idescB, accumulate);
                            }
                        }
                    }
                    mma_commit(bar_empty + 8 * s);   // shared-memory stage free once these MMAs have read it
                    if (h == KS - 1) mma_commit(bar_tfull + 8 * a);   // accumulator complete
                }
                __syncwarp();
            }
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == kEpiWarps + kProdWarps) {
        fence_after_sync();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(TMEM_COLS));
    }
}

// 0: no L2 prefetch, 1 (default): epilogue-only tensors prefetched for the current tile, 2 (SIC_DENSE_BWD_PF=far): everything two ahead
inline int dense_bwd_prefetch_mode() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("SIC_DENSE_BWD_PF");
        v = !dense_prefetch_enabled() ? 0 : (e && e[0] == 'f') ? 2 : 1;
    }
    return v;
}

template <int C, int OP>
int launch_dense_op(const char *what, const float *in0, const float *beta_param, const float *gamma_param, long P, float *out0,
                    const float *in1, const float *in2, float *out1, float *part, cudaStream_t st) {
    const size_t smem = WsCfg<C>::SMEM;
    auto kern = gdn_dense_ws_kernel<C, OP>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
        set_error("%s: cannot reserve %zu B of shared memory: %s", what, smem, cudaGetErrorString(e));
        return (int)e;
    }
    kern<<<dense_grid(P, WsCfg<C>::TN), kThreadsWS, smem, st>>>(in0, beta_param, gamma_param, P, dense_bwd_prefetch_mode(),
                                                             dense_gamma_in_tmem(), out0, in1, in2, out1, part);
    cudaError_t le = cudaGetLastError();
    if (le != cudaSuccess) {
        set_error("%s: launch failed: %s", what, cudaGetErrorString(le));
        return (int)le;
    }
    return 0;
}

// backward = pass 1 (s, h, direct, dbeta partials) then pass 2 (dx) on the same stream
template <int C>
int launch_dense_bwd(const float *x, const float *g, const float *beta_param, const float *gamma_param, long P, int inverse, float *h,
                     float *direct, float *dx, float *dbeta_part, cudaStream_t st) {
    const char *what = "sic_gdn_dense_bwd";
    int rc = inverse ? launch_dense_op<C, OP_BWD1_INV>(what, x, beta_param, gamma_param, P, h, g, nullptr, direct, dbeta_part, st)
                     : launch_dense_op<C, OP_BWD1>(what, x, beta_param, gamma_param, P, h, g, nullptr, direct, dbeta_part, st);
    if (rc != 0) return rc;
    return launch_dense_op<C, OP_BWD2>(what, h, beta_param, gamma_param, P, dx, x, direct, nullptr, nullptr, st);
}

}  // namespace

static int dense_tile_positions(int C) { return C > 128 ? dense_ws::WsCfg<192>::TN : dense_ws::WsCfg<128>::TN; }

}  // namespace sic

using namespace sic;

extern "C" int sic_gdn_dense_bwd_part_rows(long positions, int C) {
    if (positions <= 0 || C <= 0) return 0;
    return 2 * dense_ws::dense_grid(positions, dense_tile_positions(C));   // one row per (persistent CTA, position half)
}

extern "C" int sic_gdn_dense_bwd(const float *x, const float *g, const float *beta_param, const float *gamma_param, long positions,
                                 int C, int inverse, float *h, float *direct, float *dx, float *dbeta_part, int part_rows,
                                 void *stream) {
    SIC_CHECK_ARG(positions > 0 && C > 0, "sic_gdn_dense_bwd: empty shape positions=%ld C=%d", positions, C);
    SIC_CHECK_ARG(x && g && beta_param && gamma_param && h && direct && dx && dbeta_part, "sic_gdn_dense_bwd: null pointer");
    SIC_CHECK_ARG(((uintptr_t)x & 15) == 0 && ((uintptr_t)g & 15) == 0 && ((uintptr_t)h & 15) == 0 && ((uintptr_t)direct & 15) == 0 &&
                      ((uintptr_t)dx & 15) == 0 && ((uintptr_t)gamma_param & 15) == 0,
                  "sic_gdn_dense_bwd: tensors must be 16-byte aligned");
    SIC_CHECK_ARG(h != dx && direct != dx && h != direct, "sic_gdn_dense_bwd: h, direct and dx must be distinct buffers");
    SIC_CHECK_ARG(part_rows >= sic_gdn_dense_bwd_part_rows(positions, C), "sic_gdn_dense_bwd: dbeta_part has %d rows, needs %d",
                  part_rows, sic_gdn_dense_bwd_part_rows(positions, C));
    cudaStream_t st = (cudaStream_t)stream;
    switch (C) {
        case 32: return launch_dense_bwd<32>(x, g, beta_param, gamma_param, positions, inverse, h, direct, dx, dbeta_part, st);
        case 64: return launch_dense_bwd<64>(x, g, beta_param, gamma_param, positions, inverse, h, direct, dx, dbeta_part, st);
        case 96: return launch_dense_bwd<96>(x, g, beta_param, gamma_param, positions, inverse, h, direct, dx, dbeta_part, st);
        case 128: return launch_dense_bwd<128>(x, g, beta_param, gamma_param, positions, inverse, h, direct, dx, dbeta_part, st);
        case 192: return launch_dense_bwd<192>(x, g, beta_param, gamma_param, positions, inverse, h, direct, dx, dbeta_part, st);
        default:
            set_error("sic_gdn_dense_bwd: C=%d unsupported (C in {32,64,96,128,192})", C);
            return SIC_E_UNSUPPORTED;
    }
}
