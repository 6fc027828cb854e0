// G3 backward, third pass: d(gamma_eff)_ij = sum_p h[p,i] * x[p,j]^2 on tcgen05 (SURVEY.md 8(a') G3: "dgamma_ij = sum_{b,hw} h_i x_j^2
// (GEMM, K = B*H'*W')"; the C x C parameter is /root/reference/code/modelv2/layers.py:13).
//
// Round 1 left this contraction to torch ((X*X) as a full tensor pass, then a cuBLAS GEMM with K = one million positions): 1.4 of
// the 2.3 ms of the dense backward at 16x128x256^2.  Here it is one streaming pass over h and x (8 B/element, the algorithmic
// minimum for a separate pass):
//
//   producers (8 warps)  h tile and x tile, TN positions x C channels each, straight from channels-last memory -> h split EXACTLY
//                        into tf32 hi + lo, x^2 likewise -> four shared-memory tiles in the NATURAL layout [position row][32-channel
//                        128-byte chunk].  No transpose anywhere: with K = positions that natural layout IS the canonical MN-major
//                        UMMA operand layout, for A = h^T and for B = (x^2)^T alike.  For 32-bit (tf32) MN-major operands the only
//                        layout the tensor core accepts is SWIZZLE_128B_BASE32B (cute: Layout_MN_SW128_32B_Atom, "for mn-major
//                        tf32 operands, SW128_32B is the only available smem layout"): atoms of 4 K-rows x 128 B, the 32-byte chunk
//                        index XORed with (row & 3), atoms LBO apart along M/N and SBO = 512 B apart along K.  (The first version
//                        used the 16-byte SWIZZLE_128B pattern of the K-major kernels: the MMAs ran and returned zeros.)
//   MMA warp             D[i, j] (TMEM, fp32) += hh*qh + hl*qh + hh*ql over K = 8 positions per instruction (a_major = b_major = MN in
//                        the instruction descriptor); the dropped hl*ql term is 2^-22 relative.  A plain tf32 product (truncating
//                        both operands) is biased by ~ -7e-4 relative, which fails the 1e-4 gradient tolerance; the split costs
//                        tensor time only, and the pass is HBM-bound.
//   flush warps (4)      The tensor core TRUNCATES when it adds into a large running sum: with ONE accumulator for the whole kernel
//                        (first version) d(gamma) came out 1.35e-4 low at 10^6 positions (5e-8 per accumulating MMA, linear in the
//                        position count; scripts/dgamma_bias_probe.py).  So an accumulator only lives for kDgFlushTiles = 16 tiles
//                        (192 MMAs): then these warps read it (tcgen05.ld) and add it with round-to-nearest fp32 adds into the
//                        per-CTA partial [C x C] in global memory (L2-resident; each element is owned by one thread: no atomics,
//                        fixed order).  C <= 128: two accumulators alternate, so the MMAs never wait; C = 192 (no room for two):
//                        the MMA warp waits for the read-out (~1 us per 16 tiles).
//   afterwards           a second tiny kernel folds the <= 148 partials in a fixed order (deterministic).
// C <= 128: one M = 128 block (rows >= C unused).  C == 192: M = 128 block (i = 0..127) + M = 64 block (i = 128..191, TMEM rows
// 32*(r/16) + r%16 as in gdn_dense_ws.cu), N = 192 for both.
#include "gdn_dense_ws.cuh"

namespace sic {
namespace {

using namespace umma;

constexpr int kDgProdWarps = 8, kDgProdThreads = kDgProdWarps * 32;
constexpr int kDgFlushWarps = 4, kDgFlushThreads = kDgFlushWarps * 32;
constexpr int kDgThreads = kDgProdThreads + 32 + kDgFlushThreads;   // producers, the MMA warp, the flush warps
constexpr int kDgTN = 32;                             // positions (= K) per shared-memory stage
constexpr int kDgFlushTiles = 16;                     // tiles per accumulator life (16 x 12 = 192 accumulating MMAs: bias <= 1e-5)

template <int C>
struct DgCfg {
    static constexpr uint32_t TILE = kDgTN * C * 4;                   // one of hh / hl / qh / ql
    static constexpr uint32_t STAGE = 4 * TILE;
    // C < 128: the M = 128 instruction still walks four 32-channel atoms of the A operand; the ones past C produce accumulator
    // rows nobody reads, but their addresses must stay inside the CTA's allocation -> pad behind the last stage
    static constexpr uint32_t PAD = C < 128 ? 4u * kDgTN * 128u : 0u;
    static constexpr int NS_FIT = (int)((227u * 1024u - 2048u - PAD) / STAGE);
    static constexpr int NS = NS_FIT > 4 ? 4 : NS_FIT;
    static constexpr size_t SMEM = (size_t)NS * STAGE + PAD + 1024;
    static constexpr bool kTwoBlocks = C > 128;
    static constexpr int PER = kDgTN * (C / 4) / kDgProdThreads;      // float4 of h (and of x) per producer thread and stage
    static constexpr uint32_t COLS_B = 256;                           // TMEM column of the M = 64 block
    static constexpr int NACC = kTwoBlocks ? 1 : 2;                   // accumulators that alternate between flush periods
    static constexpr uint32_t ACC_STRIDE = 128;                       // TMEM columns between the two accumulators (C <= 128)
    static constexpr uint32_t TMEM_COLS = kTwoBlocks ? 512 : 256;
    static_assert(C % 32 == 0 && (C <= 128 || C == 192), "C in {32,64,96,128,192}");
    static_assert((kDgTN * (C / 4)) % kDgProdThreads == 0 && NS >= 2, "stage geometry");
};

// MN-major SWIZZLE_128B_BASE32B operand (layout type 1): 4 K-rows of 128 bytes per atom; atoms `lbo` bytes apart along M/N;
// 4-row groups 512 B apart along K
__device__ __forceinline__ uint64_t smem_desc_mn(uint32_t saddr, uint32_t lbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)(lbo >> 4) << 16;
    d |= (uint64_t)(512 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)1 << 61;
    return d;
}
// byte offset of the 16-byte chunk c (0..7) of row r in 32-channel column block kb of a [rows x C] tile: Swizzle<2,5,2> on the byte
// address, i.e. the 32-byte chunk index (c >> 1) XOR (r & 3)
__device__ __forceinline__ uint32_t sw128b32_offset(int r, int kb, int c, int rows) {
    return (uint32_t)kb * (uint32_t)rows * 128u + (uint32_t)r * 128u + (uint32_t)((((c >> 1) ^ (r & 3)) << 5) | ((c & 1) << 4));
}
__device__ __forceinline__ uint32_t idesc_tf32_mn(int M, int N) { return idesc_tf32(M, N) | (1u << 15) | (1u << 16); }

__device__ __forceinline__ void split_tf32(const float4 &v, float4 &hi, float4 &lo) {
    hi.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u); lo.x = v.x - hi.x;
    hi.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u); lo.y = v.y - hi.y;
    hi.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u); lo.z = v.z - hi.z;
    hi.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u); lo.w = v.w - hi.w;
}

template <int C>
__global__ void __launch_bounds__(kDgThreads, 1) gdn_dense_dgamma_kernel(const float *__restrict__ x, const float *__restrict__ h, long P,
                                                                       float *__restrict__ partial) {
    using Cfg = DgCfg<C>;
    constexpr int V = C / 4, TN = kDgTN, NS = Cfg::NS, PER = Cfg::PER;
    constexpr uint32_t TILE = Cfg::TILE, LBO = TN * 128;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t sStage = (smem_u32(smem_raw) + 1023u) & ~1023u;       // NS x { hh, hl, qh, ql }
    __shared__ __align__(8) uint64_t bars[2 * NS + 4];                   // full[NS], empty[NS], acc_full[2], acc_free[2]
    __shared__ uint32_t tmem_base_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t bar_full = smem_u32(&bars[0]), bar_empty = smem_u32(&bars[NS]);
    const uint32_t bar_acc_full = smem_u32(&bars[2 * NS]), bar_acc_free = smem_u32(&bars[2 * NS + 2]);
    constexpr int NACC = Cfg::NACC;
    if (tid == 0) {
        for (int s = 0; s < NS; ++s) {
            mbar_init(bar_full + 8 * s, kDgProdThreads);
            mbar_init(bar_empty + 8 * s, 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(bar_acc_full + 8 * a, 1);
            mbar_init(bar_acc_free + 8 * a, kDgFlushThreads);
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == kDgProdWarps) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_slot)), "r"(Cfg::TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = tmem_base_slot;
    const long n_tiles = (P + TN - 1) / TN;

    long n_local = 0;                                                    // tiles of this CTA, and its flush periods
    for (long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) ++n_local;
    const long n_periods = (n_local + kDgFlushTiles - 1) / kDgFlushTiles;
    if (warp < kDgProdWarps) {
        // ===================================================== producers: h, x -> (hh, hl, qh, ql) tiles
        float4 hn[PER], xn[PER];
        auto request = [&](long t) {
            const long q0 = t * TN, vld = P - q0;                        // vld <= 0 past the end: zeros contribute nothing
            const float4 *hs = reinterpret_cast<const float4 *>(h + q0 * C), *xs = reinterpret_cast<const float4 *>(x + q0 * C);
#pragma unroll
            for (int k = 0; k < PER; ++k) {
                const int idx = tid + k * kDgProdThreads, r = idx / V;
                const bool ok = r < vld;
                hn[k] = ok ? ldg_stream(hs + idx) : make_float4(0.f, 0.f, 0.f, 0.f);     // the tile is one contiguous block
                xn[k] = ok ? ldg_stream(xs + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        auto prefetch = [&](long t) {                                    // DRAM latency is taken two tiles ahead by one bulk L2 prefetch
            const long q0 = t * TN;
            if (tid == 0 && q0 < P) {
                const long rows = (P - q0 < TN) ? (P - q0) : TN;
                prefetch_l2_bulk(h + q0 * C, (uint32_t)(rows * C * 4));
                prefetch_l2_bulk(x + q0 * C, (uint32_t)(rows * C * 4));
            }
        };
        request(blockIdx.x);
        prefetch(blockIdx.x + (long)gridDim.x);
        prefetch(blockIdx.x + 2 * (long)gridDim.x);
        uint32_t u = 0;
        for (long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++u) {
            const uint32_t s = u % NS, ph = (u / NS) & 1;
            const uint32_t sHH = sStage + s * Cfg::STAGE, sHL = sHH + TILE, sQH = sHL + TILE, sQL = sQH + TILE;
            mbar_wait(bar_empty + 8 * s, ph ^ 1);
#pragma unroll
            for (int k = 0; k < PER; ++k) {
                const int idx = tid + k * kDgProdThreads, r = idx / V, c4 = idx - r * V;
                const uint32_t off = sw128b32_offset(r, c4 >> 3, c4 & 7, TN);
                float4 hi, lo;
                split_tf32(hn[k], hi, lo);
                sts128(sHH + off, hi);
                sts128(sHL + off, lo);
                const float4 v = xn[k];
                split_tf32(make_float4(v.x * v.x, v.y * v.y, v.z * v.z, v.w * v.w), hi, lo);
                sts128(sQH + off, hi);
                sts128(sQL + off, lo);
            }
            fence_proxy_async();
            mbar_arrive(bar_full + 8 * s);
            request(tile + gridDim.x);
            prefetch(tile + 3 * (long)gridDim.x);
        }
    } else if (warp == kDgProdWarps) {
        // ===================================================== MMA warp
        const uint32_t idescA = idesc_tf32_mn(128, C > 128 ? 192 : (C < 32 ? 32 : C)), idescB = idesc_tf32_mn(64, 192);
        uint32_t u = 0;
        for (long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++u) {
            const uint32_t s = u % NS, ph = (u / NS) & 1;
            const uint32_t sHH = sStage + s * Cfg::STAGE, sHL = sHH + TILE, sQH = sHL + TILE, sQL = sQH + TILE;
            const uint32_t period = u / kDgFlushTiles, in_period = u % kDgFlushTiles, acc = period % NACC;
            const uint32_t tmem_acc = tmem_base + acc * Cfg::ACC_STRIDE;
            if (in_period == 0 && period >= (uint32_t)NACC)             // the flush warps have read this accumulator's previous life
                mbar_wait(bar_acc_free + 8 * acc, ((period / NACC) - 1) & 1);
            mbar_wait(bar_full + 8 * s, ph);
            fence_after_sync();
            if (elect_one_sync()) {
#pragma unroll
                for (int term = 0; term < 3; ++term) {                   // hh*qh, hl*qh, hh*ql
                    const uint32_t sa = term == 1 ? sHL : sHH, sb = term == 2 ? sQL : sQH;
#pragma unroll
                    for (int ks = 0; ks < TN / 8; ++ks) {                // 8 positions (two 4-row atoms, 1024 bytes) per instruction
                        const uint64_t dA = smem_desc_mn(sa + ks * 1024, LBO), dB = smem_desc_mn(sb + ks * 1024, LBO);
                        const uint32_t accumulate = (in_period | (uint32_t)term | (uint32_t)ks) != 0;   // fresh accumulator every period
                        mma_tf32(tmem_acc, dA, dB, idescA, accumulate);
                        if (Cfg::kTwoBlocks)                             // rows i = 128..191: A starts four 32-channel atoms further on
                            mma_tf32(tmem_acc + Cfg::COLS_B, smem_desc_mn(sa + ks * 1024 + 4 * LBO, LBO), dB, idescB, accumulate);
                    }
                }
                mma_commit(bar_empty + 8 * s);
                if (in_period == kDgFlushTiles - 1 || tile + gridDim.x >= n_tiles) mma_commit(bar_acc_full + 8 * acc);   // period complete
            }
            __syncwarp();
        }
    } else {
        // ===================================================== flush warps: TMEM accumulator of one period -> += partial[blockIdx.x][i][j]
        const int lq = warp & 3;                                         // TMEM lane quadrant = warp % 4 (warps 9..12 -> 1, 2, 3, 0)
        float *out = partial + (size_t)blockIdx.x * C * C;
        const int iA = lq * 32 + lane;
        const int iB = 128 + lq * 16 + (lane & 15);
        for (long p = 0; p < n_periods; ++p) {
            const uint32_t acc = (uint32_t)(p % NACC);
            const uint32_t tmem_acc = tmem_base + acc * Cfg::ACC_STRIDE;
            mbar_wait(bar_acc_full + 8 * acc, (uint32_t)((p / NACC) & 1));
            fence_after_sync();
            if (lq * 32 < C && lq * 32 < 128) {
#pragma unroll 1
                for (int j0 = 0; j0 < C; j0 += 16) {
                    float a16[16];
                    tmem_ld16(tmem_acc + ((uint32_t)(lq * 32) << 16) + (uint32_t)j0, a16);
#pragma unroll
                    for (int k4 = 0; k4 < 4; ++k4) {
                        float4 *dst = reinterpret_cast<float4 *>(out + (size_t)iA * C + j0 + 4 * k4);
                        float4 v = make_float4(a16[4 * k4], a16[4 * k4 + 1], a16[4 * k4 + 2], a16[4 * k4 + 3]);
                        if (p > 0) { const float4 o = *dst; v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
                        *dst = v;
                    }
                }
            }
            if (Cfg::kTwoBlocks) {
#pragma unroll 1
                for (int j0 = 0; j0 < C; j0 += 16) {
                    float a16[16];
                    tmem_ld16(tmem_acc + ((uint32_t)(lq * 32) << 16) + Cfg::COLS_B + (uint32_t)j0, a16);   // warp-collective: all lanes load
                    if (lane < 16) {
#pragma unroll
                        for (int k4 = 0; k4 < 4; ++k4) {
                            float4 *dst = reinterpret_cast<float4 *>(out + (size_t)iB * C + j0 + 4 * k4);
                            float4 v = make_float4(a16[4 * k4], a16[4 * k4 + 1], a16[4 * k4 + 2], a16[4 * k4 + 3]);
                            if (p > 0) { const float4 o = *dst; v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
                            *dst = v;
                        }
                    }
                }
            }
            fence_before_sync();
            mbar_arrive(bar_acc_free + 8 * acc);                         // this accumulator may start its next life
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == kDgProdWarps) {
        fence_after_sync();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(Cfg::TMEM_COLS));
    }
}

// fixed-order fold of the per-CTA partials: out[e] = sum_c partial[c][e]
__global__ void __launch_bounds__(256) dgamma_fold_kernel(const float *__restrict__ partial, int n_part, int CC, float *__restrict__ out) {
    const int e = blockIdx.x * 256 + threadIdx.x;
    if (e >= CC) return;
    float s = 0.f;
    int c = 0;
    for (; c + 8 <= n_part; c += 8) {                    // eight independent loads in flight, summed in index order
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = __ldcs(partial + (size_t)(c + u) * CC + e);
#pragma unroll
        for (int u = 0; u < 8; ++u) s += v[u];
    }
    for (; c < n_part; ++c) s += __ldcs(partial + (size_t)c * CC + e);
    out[e] = s;
}

inline int dgamma_grid(long P) {
    const long n_tiles = (P + kDgTN - 1) / kDgTN;
    return (int)(n_tiles < sm_count() ? n_tiles : sm_count());
}

template <int C>
int launch_dgamma(const float *x, const float *h, long P, float *dgamma_eff, float *partial, cudaStream_t st) {
    auto kern = gdn_dense_dgamma_kernel<C>;
    const size_t smem = DgCfg<C>::SMEM;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
        set_error("sic_gdn_dense_dgamma: cannot reserve %zu B of shared memory: %s", smem, cudaGetErrorString(e));
        return (int)e;
    }
    const int grid = dgamma_grid(P);
    kern<<<grid, kDgThreads, smem, st>>>(x, h, P, partial);
    SIC_CHECK_LAUNCH("sic_gdn_dense_dgamma");
    dgamma_fold_kernel<<<(C * C + 255) / 256, 256, 0, st>>>(partial, grid, C * C, dgamma_eff);
    SIC_CHECK_LAUNCH("sic_gdn_dense_dgamma (fold)");
    return 0;
}

SIC_REGISTER_KERNEL("gdn_dense_dgamma_kernel<128>", gdn_dense_dgamma_kernel<128>);
SIC_REGISTER_KERNEL("gdn_dense_dgamma_kernel<192>", gdn_dense_dgamma_kernel<192>);

}  // namespace
}  // namespace sic

using namespace sic;

extern "C" size_t sic_gdn_dense_dgamma_workspace_bytes(long positions, int C) {
    if (positions <= 0 || C <= 0) return 0;
    return (size_t)dgamma_grid(positions) * C * C * sizeof(float);
}

extern "C" int sic_gdn_dense_dgamma(const float *x, const float *h, long positions, int C, float *dgamma_eff, void *workspace,
                                    size_t workspace_bytes, void *stream) {
    SIC_CHECK_ARG(positions > 0 && C > 0, "sic_gdn_dense_dgamma: empty shape positions=%ld C=%d", positions, C);
    SIC_CHECK_ARG(x && h && dgamma_eff && workspace, "sic_gdn_dense_dgamma: null pointer");
    SIC_CHECK_ARG(((uintptr_t)x & 15) == 0 && ((uintptr_t)h & 15) == 0 && ((uintptr_t)workspace & 15) == 0,
                  "sic_gdn_dense_dgamma: tensors must be 16-byte aligned");
    if (!(C == 32 || C == 64 || C == 96 || C == 128 || C == 192)) {      // before the workspace check: an unsupported C is the cause
        set_error("sic_gdn_dense_dgamma: C=%d unsupported (C in {32,64,96,128,192})", C);
        return SIC_E_UNSUPPORTED;
    }
    if (workspace_bytes < sic_gdn_dense_dgamma_workspace_bytes(positions, C)) {
        set_error("sic_gdn_dense_dgamma: workspace %zu < %zu bytes", workspace_bytes, sic_gdn_dense_dgamma_workspace_bytes(positions, C));
        return SIC_E_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    float *partial = static_cast<float *>(workspace);
    switch (C) {
        case 32: return launch_dgamma<32>(x, h, positions, dgamma_eff, partial, st);
        case 64: return launch_dgamma<64>(x, h, positions, dgamma_eff, partial, st);
        case 96: return launch_dgamma<96>(x, h, positions, dgamma_eff, partial, st);
        case 128: return launch_dgamma<128>(x, h, positions, dgamma_eff, partial, st);
        case 192: return launch_dgamma<192>(x, h, positions, dgamma_eff, partial, st);
        default:
            set_error("sic_gdn_dense_dgamma: C=%d unsupported (C in {32,64,96,128,192})", C);
            return SIC_E_UNSUPPORTED;
    }
}
