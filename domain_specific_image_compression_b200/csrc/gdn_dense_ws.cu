// G3, pipelined version: dense-gamma GDN / IGDN on tcgen05 with warp-specialised producer / MMA / epilogue roles.
//
//   s[p,i] = beta_i + sum_j gamma_ij * x[p,j]^2 ,   y = x / sqrt(s)   (IGDN: x * sqrt(s))        (layers.py:13,19-27)
//
// gdn_dense.cu runs load -> square -> MMA -> epilogue strictly one after the other on every tile (48 % of the HBM peak).  Here
// the three phases of consecutive tiles overlap:
//
//   producer warps (8)   x tile (128 positions x C, one contiguous block of channels-last memory) -> registers (requested one
//                        tile ahead) -> x^2 split exactly into tf32 hi + lo -> shared memory in the K-major SWIZZLE_128B UMMA
//                        layout -> mbarrier full[s]
//   MMA warp (1 thread)  D[c_out, pos] (TMEM, 2 accumulator stages) = G[c_out, c_in] (smem, resident) * X2[pos, c_in]^T,
//                        2 * C/8 tcgen05.mma kind::tf32 of shape 128 x 128 x 8; tcgen05.commit -> empty[s] and tmem_full[a]
//   epilogue warps (8)   tcgen05.ld: lane = output channel, columns = positions.  With channels-last activations the 32 lanes of a
//                        warp therefore address 32 CONSECUTIVE floats for any fixed position: x is re-read (an L2 hit, the
//                        producer touched the tile microseconds earlier) and y = x * rsqrt(beta + acc) is written with fully
//                        coalesced 128-byte warp accesses and no shared-memory transpose -> mbarrier tmem_empty[a]
//
// gamma is the A operand (M = 128 output channels, rows >= C zero-padded), the x^2 tile is the B operand (N = 128 positions).
// The orientation is what removes the transpose of the first version (its TMEM lanes were positions, so a thread held 32
// channels of ONE position and had to go through padded shared memory to store coalesced).
// HBM traffic stays at the algorithmic 8 B/element; the epilogue's second read of x is L2 traffic.
#include "umma.cuh"

namespace sic {
namespace {

using namespace umma;

constexpr int kTileN = 128;                 // positions per tile == UMMA N
constexpr int kRowsA = 128;                 // UMMA M: output channels, zero-padded
constexpr int kEpiWarps = 8, kProdWarps = 8;
constexpr int kEpiThreads = kEpiWarps * 32, kProdThreads = kProdWarps * 32;
constexpr int kThreadsWS = kEpiThreads + kProdThreads + 32;   // + the MMA warp
constexpr int kAccCols = 128;               // TMEM columns per accumulator stage (= kTileN)
constexpr float kReparamOffset = 3.814697265625e-06f;  // 2^-18, layers.py:8

__host__ __device__ constexpr int ws_stages(int C) { return ((size_t)kRowsA * C * 4 * 5 + 1024 <= 227u * 1024u) ? 2 : 1; }
__host__ __device__ constexpr size_t ws_smem_bytes(int C) { return (size_t)kRowsA * C * 4 * (1 + 2 * ws_stages(C)) + 1024; }

// One epilogue warp, one half tile: 64 positions of this lane's output channel.  x is requested before the accumulator barrier
// is awaited so the L2 latency hides behind the MMA; FULL = all 64 positions exist (no predicates on the fast path).
template <int C, bool INVERSE, bool FULL>
__device__ __forceinline__ void epilogue_half(const float *__restrict__ xp, float *__restrict__ yp, int left, float beta,
                                              uint32_t taddr, uint32_t bar, uint32_t parity) {
    float xv[64];
#pragma unroll
    for (int j = 0; j < 64; ++j) xv[j] = (FULL || j < left) ? __ldcg(xp + (long)j * C) : 0.f;
    mbar_wait(bar, parity);
    fence_after_sync();
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
        float acc[16];
        tmem_ld16(taddr + ch * 16, acc);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const float s = beta + acc[j];
            float d;
            if (INVERSE) asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(d) : "f"(s));
            else asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(d) : "f"(s));
            const int k = ch * 16 + j;
            if (FULL || k < left) __stcs(yp + (long)k * C, xv[k] * d);
        }
    }
}

template <int C, bool INVERSE>
__global__ void __launch_bounds__(kThreadsWS, 1) gdn_dense_ws_kernel(const float *__restrict__ x, const float *__restrict__ beta_param,
                                                                     const float *__restrict__ gamma_param, long P,
                                                                     float *__restrict__ y) {
    static_assert(C % 32 == 0 && C >= 32 && C <= 128, "dense GDN kernel: C in {32,64,96,128}");
    constexpr int KB = C / 32;                       // K-blocks of 32 fp32 (one 128-byte swizzle row)
    constexpr int V = C / 4;                         // float4 per position
    constexpr int NS = ws_stages(C);                 // shared-memory stages of (hi, lo)
    constexpr uint32_t OPER_BYTES = kRowsA * C * 4;  // one K-major operand of 128 rows
    constexpr int PER = kTileN * V / kProdThreads;   // float4 per producer thread per tile
    extern __shared__ uint8_t smem_raw[];
    // 32-bit shared-window addresses throughout (st.shared, descriptors); SWIZZLE_128B operands need 1024-byte alignment
    const uint32_t sG = (smem_u32(smem_raw) + 1023u) & ~1023u;   // [128 x C]  gamma, re-parameterised
    const uint32_t sStage = sG + OPER_BYTES;                     // NS x { hi [128 x C], lo [128 x C] }
    __shared__ __align__(8) uint64_t bars[2 * NS + 4];   // full[NS], empty[NS], tmem_full[2], tmem_empty[2]
    __shared__ uint32_t tmem_base_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t bar_full = smem_u32(&bars[0]), bar_empty = smem_u32(&bars[NS]);
    const uint32_t bar_tfull = smem_u32(&bars[2 * NS]), bar_tempty = smem_u32(&bars[2 * NS + 2]);

    // ---- one-time setup: gamma -> A operand (layers.py:21 applied to the C x C matrix), barriers, TMEM
    for (int idx = tid; idx < kRowsA * V; idx += kThreadsWS) {
        const int i = idx / V, c4 = idx - i * V;
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < C) {
            g = __ldg(reinterpret_cast<const float4 *>(gamma_param) + (size_t)i * V + c4);
            g.x = g.x * g.x - kReparamOffset; g.y = g.y * g.y - kReparamOffset;
            g.z = g.z * g.z - kReparamOffset; g.w = g.w * g.w - kReparamOffset;
        }
        sts128(sG + sw128_offset(i, c4 >> 3, c4 & 7, kRowsA), g);
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
    if (warp == kEpiWarps + kProdWarps) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_slot)), "r"(2 * kAccCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    fence_proxy_async();                             // gamma was written by the generic proxy, the MMA reads through the async proxy
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = tmem_base_slot;
    const long n_tiles = (P + kTileN - 1) / kTileN;

    if (warp < kEpiWarps) {
        // ===================================================== epilogue: TMEM -> y
        const int q = warp & 3;                      // TMEM lane quadrant this warp may read (hardware rule: warp id % 4)
        const int col0 = (warp >> 2) * 64;           // its half of the 128 position columns
        const int c = q * 32 + lane;                 // output channel of this lane
        const bool ch_ok = q * 32 < C;               // warp-uniform (C is a multiple of 32); idle warps only keep the barriers in step
        float beta = 1.f;
        if (ch_ok) {
            const float b = __ldg(beta_param + c);
            beta = b * b - kReparamOffset;           // layers.py:20
        }
        long it = 0;
        for (long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const uint32_t a = (uint32_t)(it & 1), aph = (uint32_t)((it >> 1) & 1);
            const long p0 = tile * kTileN + col0;
            const long left = P - p0;                // positions of this half that exist (may be <= 0 on the last tile)
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + a * kAccCols + (uint32_t)col0;
            if (!ch_ok || left <= 0) {
                mbar_wait(bar_tfull + 8 * a, aph);
            } else if (left >= 64) {
                epilogue_half<C, INVERSE, true>(x + p0 * C + c, y + p0 * C + c, 64, beta, taddr, bar_tfull + 8 * a, aph);
            } else {
                epilogue_half<C, INVERSE, false>(x + p0 * C + c, y + p0 * C + c, (int)left, beta, taddr, bar_tfull + 8 * a, aph);
            }
            fence_before_sync();
            mbar_arrive(bar_tempty + 8 * a);         // accumulator stage a may be overwritten
        }
    } else if (warp < kEpiWarps + kProdWarps) {
        // ===================================================== producer: x -> x^2 (hi, lo) -> smem
        const int ptid = tid - kEpiThreads;
        float4 xn[PER];
        auto request = [&](long t) {
            const long q0 = t * kTileN;
            const long vld = P - q0;                 // <= 0 past the end
            const float4 *src = reinterpret_cast<const float4 *>(x + q0 * C);
#pragma unroll
            for (int k = 0; k < PER; ++k) {
                const int idx = ptid + k * kProdThreads, r = idx / V;
                xn[k] = (r < vld) ? ldg_stream(src + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        request(blockIdx.x);
        long it = 0;
        for (long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const uint32_t s = (uint32_t)(it % NS), ph = (uint32_t)((it / NS) & 1);
            const uint32_t sHi = sStage + s * 2 * OPER_BYTES, sLo = sHi + OPER_BYTES;
            mbar_wait(bar_empty + 8 * s, ph ^ 1);    // the MMAs that read this stage have completed
#pragma unroll
            for (int k = 0; k < PER; ++k) {
                const int idx = ptid + k * kProdThreads, r = idx / V, c4 = idx - r * V;
                const float4 v = xn[k];
                const float4 sq = make_float4(v.x * v.x, v.y * v.y, v.z * v.z, v.w * v.w);
                float4 hi, lo;
                hi.x = __uint_as_float(__float_as_uint(sq.x) & 0xFFFFE000u); lo.x = sq.x - hi.x;
                hi.y = __uint_as_float(__float_as_uint(sq.y) & 0xFFFFE000u); lo.y = sq.y - hi.y;
                hi.z = __uint_as_float(__float_as_uint(sq.z) & 0xFFFFE000u); lo.z = sq.z - hi.z;
                hi.w = __uint_as_float(__float_as_uint(sq.w) & 0xFFFFE000u); lo.w = sq.w - hi.w;
                const uint32_t off = sw128_offset(r, c4 >> 3, c4 & 7, kTileN);
                sts128(sHi + off, hi);
                sts128(sLo + off, lo);
            }
            fence_proxy_async();                     // generic-proxy writes -> visible to the tensor core's async proxy
            mbar_arrive(bar_full + 8 * s);
            request(tile + gridDim.x);               // next tile's HBM reads fly while the MMA and the epilogue run
        }
    } else {
        // ===================================================== MMA warp: every lane follows the barriers, lane 0 issues
        const uint32_t idesc = idesc_tf32(kRowsA, kTileN);
        const uint64_t descG = smem_desc(sG);
        long it = 0;
        for (long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const uint32_t s = (uint32_t)(it % NS), ph = (uint32_t)((it / NS) & 1);
            const uint32_t a = (uint32_t)(it & 1), aph = (uint32_t)((it >> 1) & 1);
            const uint32_t sHi = sStage + s * 2 * OPER_BYTES;
            const uint64_t descHi = smem_desc(sHi), descLo = smem_desc(sHi + OPER_BYTES);
            mbar_wait(bar_tempty + 8 * a, aph ^ 1);  // the epilogue has drained this accumulator stage
            mbar_wait(bar_full + 8 * s, ph);         // x^2 of this tile is in shared memory
            fence_after_sync();
            if (lane == 0) {
                const uint32_t tmem_d = tmem_base + a * kAccCols;
                uint32_t accumulate = 0;
#pragma unroll
                for (int pass = 0; pass < 2; ++pass) {
                    const uint64_t dB = pass == 0 ? descHi : descLo;
#pragma unroll
                    for (int kb = 0; kb < KB; ++kb) {
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {     // 4 x (8 tf32 = 32 B) inside one 128-byte swizzle row
                            const uint64_t adv = (uint64_t)((kb * (kRowsA * 128) + ks * 32) >> 4);
                            mma_tf32(tmem_d, descG + adv, dB + adv, idesc, accumulate);
                            accumulate = 1;
                        }
                    }
                }
                mma_commit(bar_empty + 8 * s);       // shared-memory stage free once these MMAs have read it
                mma_commit(bar_tfull + 8 * a);       // accumulator complete
            }
            __syncwarp();
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == kEpiWarps + kProdWarps) {
        fence_after_sync();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(2 * kAccCols));
    }
}

template <int C>
int launch_dense_ws(const float *x, const float *beta_param, const float *gamma_param, long P, int inverse, float *y, cudaStream_t st) {
    const size_t smem = ws_smem_bytes(C);
    auto kern = inverse ? gdn_dense_ws_kernel<C, true> : gdn_dense_ws_kernel<C, false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
        set_error("sic_gdn_dense_fwd (pipelined): cannot reserve %zu B of shared memory: %s", smem, cudaGetErrorString(e));
        return (int)e;
    }
    const long n_tiles = (P + kTileN - 1) / kTileN;
    const int grid = (int)(n_tiles < sm_count() ? n_tiles : sm_count());   // persistent: one CTA per SM
    kern<<<grid, kThreadsWS, smem, st>>>(x, beta_param, gamma_param, P, y);
    SIC_CHECK_LAUNCH("sic_gdn_dense_fwd (pipelined)");
    return 0;
}

}  // namespace

int gdn_dense_ws_dispatch(const float *x, const float *beta_param, const float *gamma_param, long positions, int C, int inverse,
                          float *y, cudaStream_t st) {
    switch (C) {
        case 32: return launch_dense_ws<32>(x, beta_param, gamma_param, positions, inverse, y, st);
        case 64: return launch_dense_ws<64>(x, beta_param, gamma_param, positions, inverse, y, st);
        case 96: return launch_dense_ws<96>(x, beta_param, gamma_param, positions, inverse, y, st);
        case 128: return launch_dense_ws<128>(x, beta_param, gamma_param, positions, inverse, y, st);
        default:
            set_error("sic_gdn_dense_fwd: C=%d unsupported (gamma and the x^2 hi/lo tile are resident in shared memory: C in "
                      "{32,64,96,128}; wider layers need K-streaming)", C);
            return SIC_E_UNSUPPORTED;
    }
}

}  // namespace sic
