// G3: GDN / IGDN with a DENSE gamma (C x C) — the channel contraction on tcgen05 tensor cores.
//
//   s[p,i] = beta_i + sum_j gamma_ij * x[p,j]^2 ,   y = x / sqrt(s)   (IGDN: x * sqrt(s))
//
// The reference allocates the C x C parameter (`self.gamma`, layers.py:13) but its forward only ever uses the diagonal
// through a depthwise 1x1 conv (layers.py:21-23).  The dense contraction is what north_star names; its
// oracle is F.conv2d(x**2, gamma.view(C,C,1,1), beta) and it equals the reference when gamma is diagonal.
//
// Mapping (channels-last activations: a tile of 128 positions x C channels is ONE contiguous block of memory):
//   D[128 pos x C_out] (TMEM, fp32)  +=  A[128 pos x C_in] (smem, x^2)  *  B[C_out x C_in]^T (smem, gamma)      kind::tf32
//   - A and B are K-major in the canonical 128-byte-swizzled UMMA layout (8 rows x 128 B atoms, 32 fp32 per K-block);
//   - x^2 is produced by the SIMT threads between the global load and the MMA, so they write A directly in that layout
//     (a TMA tensor load would land x, not x^2; there is nothing left for TMA to stage);
//   - TF32 keeps 10 mantissa bits: x^2 is split EXACTLY into hi = tf32(x^2) and lo = x^2 - hi and both are multiplied
//     (2 MMAs per K-step), so the activations carry ~2^-20; gamma is used at TF32 precision (a parameter perturbation of
//     2^-11, which the oracle applies as well)  =>  tolerance-only mode, NOT the bit-exact path (that is the diagonal one);
//   - one elected thread issues tcgen05.mma; completion comes back through tcgen05.commit -> mbarrier;
//   - epilogue: tcgen05.ld (32 lanes x 32 columns per warp) -> rsqrt(beta + acc) -> smem (conflict-free rotated float4
//     stores) -> every thread multiplies the x it kept in registers and streams y out with coalesced 128-bit stores.
// Even dense, the op is HBM-bound (8 B/element against 2C flop/B... C/4 flop/B = 32 at C=128): the MMA time per tile
// (2 x 128^3 MAC at the tf32 rate ~ 1.1 us) hides under the 128 KB of HBM traffic per tile (~2.9 us per SM).
#include "common.cuh"

namespace sic {
namespace {

constexpr int kTileM = 128;      // positions per tile == UMMA M == TMEM lanes
constexpr int kThreads = 512;     // 16 warps: 8 float4 of x per thread for the current tile + 8 prefetched for the next
constexpr float kOffset = 3.814697265625e-06f;  // 2^-18, layers.py:8

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// byte offset of the 16-byte chunk (row r, K-block kb, chunk c in [0,8)) in a K-major SWIZZLE_128B operand with `rows` rows
__device__ __forceinline__ uint32_t sw128_offset(int r, int kb, int c, int rows) {
    return (uint32_t)kb * (uint32_t)rows * 128u + (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u + (uint32_t)((c ^ (r & 7)) << 4);
}

// UMMA shared-memory descriptor, K-major, SWIZZLE_128B, 8-row groups 1024 B apart (cute::UMMA::SmemDescriptor bit layout)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);  // start address      bits [0,14)
    d |= (uint64_t)1 << 16;                    // leading byte off.  bits [16,30)  (ignored for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;          // stride byte offset bits [32,46)
    d |= (uint64_t)1 << 46;                    // version = 1 (Blackwell)
    d |= (uint64_t)2 << 61;                    // layout type SWIZZLE_128B
    return d;
}

// kind::tf32, fp32 accumulate, A and B K-major (cute::UMMA::InstrDescriptor bit layout)
__device__ __forceinline__ uint32_t umma_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// C is a template parameter so that the x tile stays in registers (C/8 float4 per thread) and the K loops unroll.
template <int C, bool INVERSE>
__global__ void __launch_bounds__(kThreads, 1) gdn_dense_fwd_kernel(const float *__restrict__ x, const float *__restrict__ beta_param,
                                                                    const float *__restrict__ gamma_param, long P,
                                                                    float *__restrict__ y) {
    static_assert(C % 32 == 0 && C >= 32 && C <= 128, "dense GDN kernel: C in {32,64,96,128}");
    constexpr int KB = C / 32;               // K-blocks of 32 fp32 (128 B)
    constexpr int V = C / 4;                 // float4 per position
    constexpr int PER_THREAD = kTileM * V / kThreads;
    constexpr uint32_t A_BYTES = kTileM * C * 4, G_BYTES = C * C * 4;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *sG = smem, *sAhi = smem + G_BYTES, *sAlo = sAhi + A_BYTES;
    float *sBeta = reinterpret_cast<float *>(sAlo + A_BYTES);
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // ---- one-time setup: gamma (re-parameterised, layers.py:21 applied to the C x C matrix) into the swizzled B operand
    for (int idx = tid; idx < C * V; idx += kThreads) {
        int i = idx / V, c4 = idx - i * V;
        float4 g = __ldg(reinterpret_cast<const float4 *>(gamma_param) + idx);
        g.x = g.x * g.x - kOffset; g.y = g.y * g.y - kOffset; g.z = g.z * g.z - kOffset; g.w = g.w * g.w - kOffset;
        *reinterpret_cast<float4 *>(sG + sw128_offset(i, c4 >> 3, c4 & 7, C)) = g;
    }
    for (int c = tid; c < C; c += kThreads) {
        float b = __ldg(beta_param + c);
        sBeta[c] = b * b - kOffset;          // layers.py:20
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_slot)), "r"(128));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(&bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem_d = tmem_base_slot;
    const uint32_t idesc = umma_idesc_tf32(kTileM, C);
    const uint64_t descG = umma_desc(smem_u32(sG)), descHi = umma_desc(smem_u32(sAhi)), descLo = umma_desc(smem_u32(sAlo));
    uint32_t phase = 0;

    const long n_tiles = (P + kTileM - 1) / kTileM;
    // software pipeline: the x tile of the NEXT iteration is requested from HBM before this iteration's MMA/epilogue, so the
    // load latency (the dominant stall of the first version: ncu long_scoreboard 7.6/issue) hides behind tensor + epilogue work
    float4 xn[PER_THREAD];
    auto prefetch = [&](long t) {
        const long q0 = t * kTileM;
        const int vld = (int)min((long)kTileM, P - q0);
#pragma unroll
        for (int k = 0; k < PER_THREAD; ++k) {
            int idx = tid + k * kThreads, r = idx / V;
            xn[k] = (t < n_tiles && r < vld) ? ldg_stream(reinterpret_cast<const float4 *>(x + q0 * C) + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    prefetch(blockIdx.x);
    for (long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long p0 = tile * kTileM;
        const int valid = (int)min((long)kTileM, P - p0);
        // ---- take the prefetched x (coalesced: the tile is one contiguous block), square, split, write A_hi / A_lo swizzled
        float4 xr[PER_THREAD];
#pragma unroll
        for (int k = 0; k < PER_THREAD; ++k) xr[k] = xn[k];
        prefetch(tile + gridDim.x);
#pragma unroll
        for (int k = 0; k < PER_THREAD; ++k) {
            int idx = tid + k * kThreads, r = idx / V, c4 = idx - r * V;
            float4 q = make_float4(xr[k].x * xr[k].x, xr[k].y * xr[k].y, xr[k].z * xr[k].z, xr[k].w * xr[k].w);
            float4 hi, lo;
            hi.x = __uint_as_float(__float_as_uint(q.x) & 0xFFFFE000u); lo.x = q.x - hi.x;
            hi.y = __uint_as_float(__float_as_uint(q.y) & 0xFFFFE000u); lo.y = q.y - hi.y;
            hi.z = __uint_as_float(__float_as_uint(q.z) & 0xFFFFE000u); lo.z = q.z - hi.z;
            hi.w = __uint_as_float(__float_as_uint(q.w) & 0xFFFFE000u); lo.w = q.w - hi.w;
            uint32_t off = sw128_offset(r, c4 >> 3, c4 & 7, kTileM);
            *reinterpret_cast<float4 *>(sAhi + off) = hi;
            *reinterpret_cast<float4 *>(sAlo + off) = lo;
        }
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // generic-proxy smem writes -> visible to the MMA (async proxy)
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
        __syncthreads();
        // ---- one thread issues 2 * C/8 MMAs (hi then lo) of shape 128 x C x 8 and commits to the mbarrier
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            uint32_t acc = 0;
#pragma unroll
            for (int pass = 0; pass < 2; ++pass) {
                const uint64_t dA = pass == 0 ? descHi : descLo;
#pragma unroll
                for (int kb = 0; kb < KB; ++kb) {
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {   // 4 x (8 tf32 = 32 B) inside one 128 B swizzle row
                        uint64_t a = dA + (uint64_t)((kb * (kTileM * 128) + ks * 32) >> 4);
                        uint64_t b = descG + (uint64_t)((kb * (C * 128) + ks * 32) >> 4);
                        umma_tf32(tmem_d, a, b, idesc, acc);
                        acc = 1;
                    }
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&bar)) : "memory");
        }
        mbar_wait(smem_u32(&bar), phase);
        phase ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        // ---- epilogue 1: TMEM -> registers -> 1/sqrt(beta + acc) (or sqrt) -> smem, row-major [128][C+4] over the (now free)
        //      A_hi/A_lo region; the 4-float row padding puts the 8 lanes of a quarter-warp on 8 different bank groups
        constexpr int SD = C + 4;
        float *sD = reinterpret_cast<float *>(sAhi);
        {
            const int lg = warp & 3;                       // TMEM lane group this warp may touch
            const int row = lg * 32 + lane;
            for (int cb = (warp >> 2) * 32; cb < C; cb += (kThreads / 128) * 32) {
                float v[32];
                tmem_ld32(tmem_d + ((uint32_t)(lg * 32) << 16) + (uint32_t)cb, v);
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    float s = sBeta[cb + j] + v[j];
                    v[j] = INVERSE ? __fsqrt_rn(s) : rsqrtf(s);
                }
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4)
                    *reinterpret_cast<float4 *>(sD + row * SD + cb + j4 * 4) = make_float4(v[j4 * 4], v[j4 * 4 + 1], v[j4 * 4 + 2], v[j4 * 4 + 3]);
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
        __syncthreads();
        // ---- epilogue 2: y = x * d with the x kept in registers, coalesced 128-bit streaming stores
#pragma unroll
        for (int k = 0; k < PER_THREAD; ++k) {
            int idx = tid + k * kThreads, r = idx / V;
            if (r < valid) {
                float4 d = *reinterpret_cast<const float4 *>(sD + r * SD + (idx - r * V) * 4);
                stg_stream(reinterpret_cast<float4 *>(y + p0 * C) + idx, make_float4(xr[k].x * d.x, xr[k].y * d.y, xr[k].z * d.z, xr[k].w * d.w));
            }
        }
        __syncthreads();                                   // sD aliases A_hi/A_lo: the next tile must not overwrite it early
    }
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_d), "r"(128));
}

template <int C>
int launch_dense(const float *x, const float *beta_param, const float *gamma_param, long P, int inverse, float *y, cudaStream_t st) {
    const size_t smem = (size_t)C * C * 4 + 2 * (size_t)kTileM * C * 4 + (size_t)C * 4 + 1024;
    auto kern = inverse ? gdn_dense_fwd_kernel<C, true> : gdn_dense_fwd_kernel<C, false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
        set_error("sic_gdn_dense_fwd: cannot reserve %zu B of shared memory: %s", smem, cudaGetErrorString(e));
        return (int)e;
    }
    long n_tiles = (P + kTileM - 1) / kTileM;
    int grid = (int)(n_tiles < sm_count() ? n_tiles : sm_count());   // persistent: one CTA per SM
    kern<<<grid, kThreads, smem, st>>>(x, beta_param, gamma_param, P, y);
    SIC_CHECK_LAUNCH("sic_gdn_dense_fwd");
    return 0;
}

}  // namespace
}  // namespace sic

namespace sic {
// gdn_dense_ws.cu: the warp-specialised, pipelined kernel (SIC_DENSE_PIPELINED)
int gdn_dense_ws_dispatch(const float *x, const float *beta_param, const float *gamma_param, long positions, int C, int inverse,
                          float *y, cudaStream_t st);
}  // namespace sic

using namespace sic;

static int dense_serial_dispatch(const float *x, const float *beta_param, const float *gamma_param, long positions, int C,
                                 int inverse, float *y, cudaStream_t st) {
    switch (C) {
        case 32: return launch_dense<32>(x, beta_param, gamma_param, positions, inverse, y, st);
        case 64: return launch_dense<64>(x, beta_param, gamma_param, positions, inverse, y, st);
        case 96: return launch_dense<96>(x, beta_param, gamma_param, positions, inverse, y, st);
        case 128: return launch_dense<128>(x, beta_param, gamma_param, positions, inverse, y, st);
        default:
            set_error("sic_gdn_dense_fwd: C=%d unsupported (this build keeps gamma and the x^2 hi/lo tiles resident in shared "
                      "memory: C in {32,64,96,128}; wider layers need K-streaming)", C);
            return SIC_E_UNSUPPORTED;
    }
}

extern "C" int sic_gdn_dense_fwd_variant(const float *x, const float *beta_param, const float *gamma_param, long positions, int C,
                                         int inverse, float *y, int variant, void *stream) {
    SIC_CHECK_ARG(positions > 0 && C > 0, "sic_gdn_dense_fwd: empty shape positions=%ld C=%d", positions, C);
    SIC_CHECK_ARG(x && y && beta_param && gamma_param, "sic_gdn_dense_fwd: null pointer");
    SIC_CHECK_ARG(((uintptr_t)x & 15) == 0 && ((uintptr_t)y & 15) == 0 && ((uintptr_t)gamma_param & 15) == 0,
                  "sic_gdn_dense_fwd: tensors must be 16-byte aligned");
    SIC_CHECK_ARG(variant == SIC_DENSE_SERIAL || variant == SIC_DENSE_PIPELINED, "sic_gdn_dense_fwd: unknown variant %d", variant);
    cudaStream_t st = (cudaStream_t)stream;
    if (variant == SIC_DENSE_PIPELINED) return gdn_dense_ws_dispatch(x, beta_param, gamma_param, positions, C, inverse, y, st);
    return dense_serial_dispatch(x, beta_param, gamma_param, positions, C, inverse, y, st);
}

extern "C" int sic_gdn_dense_fwd(const float *x, const float *beta_param, const float *gamma_param, long positions, int C,
                                 int inverse, float *y, void *stream) {
    return sic_gdn_dense_fwd_variant(x, beta_param, gamma_param, positions, C, inverse, y, SIC_DENSE_DEFAULT, stream);
}
