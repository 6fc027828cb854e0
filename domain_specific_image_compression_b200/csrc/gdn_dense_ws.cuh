// Geometry, constants and small helpers shared by the pipelined dense-gamma GDN kernels (gdn_dense_ws.cu: forward,
// gdn_dense_bwd.cu: the two backward passes).  See gdn_dense_ws.cu for the design.
#pragma once
#include <stdlib.h>

#include "umma.cuh"

namespace sic {
namespace dense_ws {

using namespace umma;

constexpr int kEpiWarps = 8, kProdWarps = 8;
constexpr int kEpiThreads = kEpiWarps * 32, kProdThreads = kProdWarps * 32;
constexpr int kThreadsWS = kEpiThreads + kProdThreads + 32;   // + the MMA warp
constexpr int kColsB = 128;                 // TMEM column offset of the M = 64 block inside an accumulator stage
constexpr float kReparamOffset = 3.814697265625e-06f;  // 2^-18, layers.py:8

// Tile geometry.  The K (input-channel) axis of a tile is fed to the tensor core in KS sub-steps of KBS 128-byte K-blocks, each
// sub-step through one shared-memory stage, so the producers fill sub-step u+1 while the MMAs of sub-step u run.
template <int C>
struct WsCfg {
    static constexpr bool kTwoBlocks = C > 128;                     // C == 192: M = 128 block + M = 64 block
    static constexpr int TN = kTwoBlocks ? 96 : 128;                // positions per tile == UMMA N
    static constexpr int ROWS_G = kTwoBlocks ? C : 128;             // rows of the gamma operand per K-block
    static constexpr int NP = TN / 2;                               // positions per epilogue warp
    static constexpr int KB = C / 32;                               // K-blocks of 32 fp32 (one 128-byte swizzle row)
    static constexpr int KS = C == 192 ? 6 : C == 128 ? 2 : C == 96 ? 3 : 1;   // sub-steps per tile
    static constexpr int KBS = KB / KS;                             // K-blocks per sub-step
    static constexpr uint32_t G_BYTES = ROWS_G * C * 4;
    static constexpr uint32_t B_BYTES = TN * KBS * 128;             // one of hi / lo of one sub-step
    static constexpr int NS_FIT = (int)((227u * 1024u - 1024u - G_BYTES) / (2 * B_BYTES));
    static constexpr int NS = NS_FIT > 4 ? 4 : NS_FIT;              // shared-memory stages
    static constexpr size_t SMEM = (size_t)G_BYTES + (size_t)NS * 2 * B_BYTES + 1024;
    static constexpr int ACC_COLS = kTwoBlocks ? 256 : 128;         // TMEM columns per accumulator stage
    static constexpr int VH = KBS * 8;                              // float4 per position per sub-step
    static constexpr int PER = TN * VH / kProdThreads;              // float4 per producer thread per sub-step
    static_assert(C % 32 == 0 && (C <= 128 || C == 192), "dense GDN kernel: C in {32,64,96,128,192}");
    static_assert(KB % KS == 0 && (TN * VH) % kProdThreads == 0, "producer threads must tile the sub-step");
    static_assert(NS >= 1 && SMEM <= 227u * 1024u, "operands do not fit in shared memory");
};

__device__ __forceinline__ float4 ldg_keep(const float4 *p) {   // read-only path, NORMAL L2 priority (the epilogue re-reads the tile)
    float4 r;
    asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}

template <bool INVERSE>
__device__ __forceinline__ float norm_factor(float s) {
    float d;
    if (INVERSE) asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(d) : "f"(s));
    else asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(d) : "f"(s));
    return d;
}

// SIC_DENSE_PREFETCH=0 switches the bulk L2 prefetch off (A/B timing only; results are identical either way)
inline int dense_prefetch_enabled() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("SIC_DENSE_PREFETCH");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v;
}

// SIC_DENSE_TS=1: gamma as a tensor-memory A operand (experimental until parity-tested on the device); default: shared memory
inline int dense_gamma_in_tmem() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("SIC_DENSE_TS");
        v = (e && e[0] == '1') ? 1 : 0;
    }
    return v;
}

inline int dense_grid(long P, int TN) {
    const long n_tiles = (P + TN - 1) / TN;
    return (int)(n_tiles < sm_count() ? n_tiles : sm_count());   // persistent: one CTA per SM
}

}  // namespace dense_ws
}  // namespace sic
