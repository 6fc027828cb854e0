// G3, pipelined version: dense-gamma GDN / IGDN on tcgen05 with warp-specialised producer / MMA / epilogue roles.
//
//   s[p,i] = beta_i + sum_j gamma_ij * x[p,j]^2 ,   y = x / sqrt(s)   (IGDN: x * sqrt(s))        (layers.py:13,19-27)
//
// gdn_dense.cu runs load -> square -> MMA -> epilogue strictly one after the other on every tile (48 % of the HBM peak).  Here
// the three phases of consecutive tiles overlap:
//
//   producer warps (8)   x tile (TN positions x C, one contiguous block of channels-last memory) -> registers (requested one
//                        tile ahead) -> x^2 split exactly into tf32 hi + lo -> shared memory in the K-major SWIZZLE_128B UMMA
//                        layout -> mbarrier full[s]
//   MMA warp (1 thread)  D[c_out, pos] (TMEM, 2 accumulator stages) = G[c_out, c_in] (smem, resident) * X2[pos, c_in]^T,
//                        2 * C/8 tcgen05.mma kind::tf32 per M block; tcgen05.commit -> empty[s] and tmem_full[a]
//   epilogue warps (8)   tcgen05.ld: lane = output channel, columns = positions.  With channels-last activations the lanes of a
//                        warp therefore address CONSECUTIVE floats for any fixed position: x is re-read (an L2 hit: the producer
//                        touched the tile microseconds earlier with a normal-priority load) and y = x * rsqrt(beta + acc) is
//                        written with coalesced warp accesses and no shared-memory transpose -> mbarrier tmem_empty[a]
//
// gamma is the A operand (M = output channels), the x^2 tile is the B operand (N = positions).  That orientation is what removes
// the transpose of the first version (its TMEM lanes were positions, so a thread held 32 channels of ONE position and had to go
// through padded shared memory to store coalesced).
//   C <= 128: one M = 128 block (rows >= C zero-padded), TN = 128 positions per tile.
//   C == 192: gamma (144 KB) still fits beside one x^2 stage when TN = 48: an M = 128 block (channels 0..127) plus an M = 64
//             block (channels 128..191; its accumulator rows live in lanes 32*(j/16) + j%16, i.e. the low half of each warp's
//             TMEM quadrant).  This covers the N = 192 model of BASELINE.json configs[3].
// HBM traffic stays at the algorithmic 8 B/element; the epilogue's second read of x is L2 traffic (ncu: profiles/).
#include "gdn_dense_ws.cuh"

namespace sic {
namespace {

using namespace umma;
using namespace dense_ws;

// Epilogue of one warp for one tile: NP positions of this lane's output channel (block A), then of its block-B channel when
// C = 192.  x is fetched in batches of 16 positions through a 2-deep register ring: the first two batches are requested before
// the accumulator barrier is awaited (their L2 latency hides behind the MMA), every later batch while the two before it are being
// normalised and stored.  32 + 16 live values keep the role inside the 96-register budget of a 17-warp CTA without spills
// (spill traffic shares the L1 data pipe with the tensor core's operand reads, which is this kernel's critical resource).
template <int C, bool INVERSE, bool FULL>
__device__ __forceinline__ void epilogue_tile(const float *__restrict__ x, float *__restrict__ y, long p0, int left, int cA, bool okA,
                                              float betaA, int cB, bool okB, float betaB, uint32_t taddr, uint32_t bar,
                                              uint32_t parity) {
    using Cfg = WsCfg<C>;
    constexpr int NP = Cfg::NP;
    static_assert(NP % 16 == 0, "epilogue batches are 16 columns");
    constexpr int NBB = NP / 16;                                 // batches per block
    constexpr int NB = Cfg::kTwoBlocks ? 2 * NBB : NBB;
    const float *xA = x + p0 * C + cA, *xB = x + p0 * C + cB;
    float *yA = y + p0 * C + cA, *yB = y + p0 * C + cB;
    float ring[2][16];
    auto load = [&](float (&dst)[16], int b) {
        const bool blkB = b >= NBB;
        const int k0 = (blkB ? b - NBB : b) * 16;
        const float *src = blkB ? xB : xA;
        const bool ok = blkB ? okB : okA;
#pragma unroll
        for (int j = 0; j < 16; ++j) dst[j] = (ok && (FULL || k0 + j < left)) ? __ldcg(src + (long)(k0 + j) * C) : 0.f;
    };
    load(ring[0], 0);
    if (NB > 1) load(ring[1], 1);
    mbar_wait(bar, parity);
    fence_after_sync();
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        const bool blkB = b >= NBB;
        const int k0 = (blkB ? b - NBB : b) * 16;
        float *dst = blkB ? yB : yA;
        const bool ok = blkB ? okB : okA;
        const float beta = blkB ? betaB : betaA;
        float acc[16];
        tmem_ld16(taddr + (blkB ? kColsB : 0) + k0, acc);
#pragma unroll
        for (int j = 0; j < 16; ++j)
            if (ok && (FULL || k0 + j < left)) __stcs(dst + (long)(k0 + j) * C, ring[b & 1][j] * norm_factor<INVERSE>(beta + acc[j]));
        if (b + 2 < NB) load(ring[b & 1], b + 2);
    }
}

template <int C, bool INVERSE>
__global__ void __launch_bounds__(kThreadsWS, 1) gdn_dense_ws_kernel(const float *__restrict__ x, const float *__restrict__ beta_param,
                                                                     const float *__restrict__ gamma_param, long P, int use_prefetch,
                                                                     int gamma_in_tmem, float *__restrict__ y) {
    using Cfg = WsCfg<C>;
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
            g = __ldg(reinterpret_cast<const float4 *>(gamma_param) + (size_t)i * V + c4);
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
    const bool a_tmem = !Cfg::kTwoBlocks && gamma_in_tmem != 0;
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
        long it = 0;
        for (long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const uint32_t a = (uint32_t)(it & 1), aph = (uint32_t)((it >> 1) & 1);
            const long p0 = tile * TN + col0;
            const long left = P - p0;                // positions of this half that exist (may be <= 0 on the last tile)
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + a * ACC_COLS + (uint32_t)col0;
            if (!okA || left <= 0) {
                mbar_wait(bar_tfull + 8 * a, aph);
            } else if (left >= NP) {
                epilogue_tile<C, INVERSE, true>(x, y, p0, NP, cA, okA, betaA, cB, okB, betaB, taddr, bar_tfull + 8 * a, aph);
            } else {
                epilogue_tile<C, INVERSE, false>(x, y, p0, (int)left, cA, okA, betaA, cB, okB, betaB, taddr, bar_tfull + 8 * a, aph);
            }
            fence_before_sync();
            mbar_arrive(bar_tempty + 8 * a);         // accumulator stage a may be overwritten
        }
    } else if (warp < kEpiWarps + kProdWarps) {
        // ===================================================== producer: x -> x^2 (hi, lo) -> smem, one K sub-step at a time
        const int ptid = tid - kEpiThreads;
        float4 xn[PER];
        auto request = [&](long t, int h) {          // channels [h*32*KBS, (h+1)*32*KBS) of the TN positions of tile t -> registers
            const long q0 = t * TN;
            const long vld = P - q0;                 // <= 0 past the end
            const float4 *src = reinterpret_cast<const float4 *>(x + q0 * C) + h * VH;
#pragma unroll
            for (int k = 0; k < PER; ++k) {
                const int idx = ptid + k * kProdThreads, r = idx / VH, c4 = idx - r * VH;
                xn[k] = (r < vld) ? ldg_keep(src + (long)r * V + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        // DRAM latency is taken by an L2 prefetch two tiles ahead (one thread, one bulk instruction, no registers); the register
        // loads one sub-step ahead then hit L2, so the in-flight window is not bounded by the producers' register file
        auto prefetch = [&](long t) {
            const long q0 = t * TN;
            if (use_prefetch && ptid == 0 && q0 < P) {
                const long rows = (P - q0 < TN) ? (P - q0) : TN;
                prefetch_l2_bulk(x + q0 * C, (uint32_t)(rows * C * 4));
            }
        };
        request(blockIdx.x, 0);
        prefetch(blockIdx.x + (long)gridDim.x);
        uint32_t u = 0;                              // sub-step counter: stage = u % NS, phase = (u / NS) & 1
        for (long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
#pragma unroll 1
            for (int h = 0; h < KS; ++h, ++u) {
                const uint32_t s = u % NS, ph = (u / NS) & 1;
                const uint32_t sHi = sStage + s * 2 * B_BYTES, sLo = sHi + B_BYTES;
                mbar_wait(bar_empty + 8 * s, ph ^ 1);    // the MMAs that read this stage have completed
#pragma unroll
                for (int k = 0; k < PER; ++k) {
                    const int idx = ptid + k * kProdThreads, r = idx / VH, c4 = idx - r * VH;
                    const float4 v = xn[k];
                    const float4 sq = make_float4(v.x * v.x, v.y * v.y, v.z * v.z, v.w * v.w);
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

template <int C>
int launch_dense_ws(const float *x, const float *beta_param, const float *gamma_param, long P, int inverse, float *y, cudaStream_t st) {
    const size_t smem = WsCfg<C>::SMEM;
    auto kern = inverse ? gdn_dense_ws_kernel<C, true> : gdn_dense_ws_kernel<C, false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
        set_error("sic_gdn_dense_fwd (pipelined): cannot reserve %zu B of shared memory: %s", smem, cudaGetErrorString(e));
        return (int)e;
    }
    const long n_tiles = (P + WsCfg<C>::TN - 1) / WsCfg<C>::TN;
    const int grid = (int)(n_tiles < sm_count() ? n_tiles : sm_count());   // persistent: one CTA per SM
    kern<<<grid, kThreadsWS, smem, st>>>(x, beta_param, gamma_param, P, dense_prefetch_enabled(), dense_gamma_in_tmem(), y);
    SIC_CHECK_LAUNCH("sic_gdn_dense_fwd (pipelined)");
    return 0;
}

SIC_REGISTER_KERNEL("gdn_dense_ws_kernel<128,0>", gdn_dense_ws_kernel<128, false>);
SIC_REGISTER_KERNEL("gdn_dense_ws_kernel<128,1>", gdn_dense_ws_kernel<128, true>);
SIC_REGISTER_KERNEL("gdn_dense_ws_kernel<192,0>", gdn_dense_ws_kernel<192, false>);
SIC_REGISTER_KERNEL("gdn_dense_ws_kernel<192,1>", gdn_dense_ws_kernel<192, true>);

}  // namespace

int gdn_dense_ws_dispatch(const float *x, const float *beta_param, const float *gamma_param, long positions, int C, int inverse,
                          float *y, cudaStream_t st) {
    switch (C) {
        case 32: return launch_dense_ws<32>(x, beta_param, gamma_param, positions, inverse, y, st);
        case 64: return launch_dense_ws<64>(x, beta_param, gamma_param, positions, inverse, y, st);
        case 96: return launch_dense_ws<96>(x, beta_param, gamma_param, positions, inverse, y, st);
        case 128: return launch_dense_ws<128>(x, beta_param, gamma_param, positions, inverse, y, st);
        case 192: return launch_dense_ws<192>(x, beta_param, gamma_param, positions, inverse, y, st);
        default:
            set_error("sic_gdn_dense_fwd: C=%d unsupported (gamma and one x^2 hi/lo tile must be resident in shared memory: "
                      "C in {32,64,96,128,192})", C);
            return SIC_E_UNSUPPORTED;
    }
}

}  // namespace sic
