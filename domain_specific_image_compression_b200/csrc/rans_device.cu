// N1 / E1 on the GPU: interleaved rANS encoder / decoder, format SIC-RANS-1 — byte-identical to rans_host.cpp.
//
// Replaces the CPU coder + per-patch device->host copies of /root/reference/code/modelv2/eval_selfcontained_entropy.py
// :48,62 (encode) and :96,116 (decode), where torchac runs serially on the host for every patch and stream.
// One WARP per stream (one patch, one latent): lane l owns the symbols i with i mod 32 == l — exactly the 32 interleaved
// lanes of the format — so an iteration codes 32 symbols at once, and the only serial dependence is the word cursor,
// which advances by a warp ballot + popcount.  All patches of a batch run concurrently (patches are independent).
//   encoder: walks the iterations backwards and fills the word area from its END, lane ranks descending, which is the
//            reversed emission order the format prescribes; then slides the words down behind the 32 state words.
//   decoder: walks forwards; a lane that drops below 2^16 pulls the next word, rank = popc(ballot & lanes below).
// Table rows (uint16, possibly with zero-width symbols) are widened on the fly exactly like the host coder:
//   c'_k = floor(c_k (65536 - L) / 65535) + k.   In broadcast mode (one row per channel) the row in use is staged in
// shared memory, so the decoder's binary search runs there.
#include "common.cuh"

namespace sic {
namespace {

constexpr int kWarps = 4;  // streams per CTA
constexpr uint32_t kLow = 1u << 16;
constexpr int kMaxL = 4096;

__device__ __forceinline__ uint32_t widen(uint32_t c, uint32_t k, uint32_t L) { return c * (65536u - L) / 65535u + k; }

__global__ void __launch_bounds__(kWarps * 32) rans_encode_kernel(const int32_t *__restrict__ sym, const uint16_t *__restrict__ tables,
                                                                  const int32_t *__restrict__ Ls, int n_streams, long n,
                                                                  long sym_per_row, long rows_per_stream, int stride, uint8_t *out,
                                                                  long cap, int32_t *__restrict__ out_nbytes) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int s = blockIdx.x * kWarps + warp;
    if (s >= n_streams) return;
    const uint32_t L = (uint32_t)Ls[s];
    const int32_t *sy = sym + (long)s * n;
    const uint16_t *tab = tables + (long)s * rows_per_stream * stride;
    uint8_t *o = out + (long)s * cap;
    uint16_t *words = reinterpret_cast<uint16_t *>(o + 128);
    const long wcap = (cap - 128) / 2;
    long pos = wcap;  // words are written at [pos, wcap), growing downwards
    uint32_t x = kLow;
    bool bad = false;
    if (Ls[s] < 1 || Ls[s] > kMaxL || Ls[s] > stride - 1) {   // no support / wider than a table row: nothing can be coded
        if (lane == 0) out_nbytes[s] = -1;
        return;
    }
    const long iters = (n + 31) / 32;
    // The only serial dependence is the state x (and the word cursor).  Symbols and their table entries do not depend on it, so they
    // are fetched kEncU iterations at a time (independent loads in flight) before the kEncU state updates run: the first version
    // paid two dependent global loads (symbol -> table row) per iteration, ~1300 cycles each, 4.2 ms for a 196 608-symbol stream.
    constexpr int kEncU = 8;
    const bool small = n < (1L << 31) && sym_per_row < (1L << 31);
    for (long jb = iters; jb > 0; jb -= kEncU) {
        uint32_t c0s[kEncU], fs[kEncU];
#pragma unroll
        for (int u = 0; u < kEncU; ++u) {
            const long j = jb - 1 - u;
            c0s[u] = 0u;
            fs[u] = 0u;                                               // 0 = no symbol here (every real symbol has width >= 1)
            const long i = j * 32 + lane;
            if (j >= 0 && i < n) {
                int32_t v = sy[i];
                if (v < 0 || (uint32_t)v >= L) { bad = true; v = 0; }
                // row of this symbol: 32-bit division whenever the stream fits 32-bit indices (always, in this codec) - a 64-bit division
                // is ~120 instructions, and ONE warp walks the whole stream: its instruction count is the coder's speed
                const long r = small ? (long)((uint32_t)i / (uint32_t)sym_per_row) : i / sym_per_row;
                const uint16_t *row = tab + r * stride;
                const uint32_t a = widen(row[v], (uint32_t)v, L), b = widen(row[v + 1], (uint32_t)v + 1, L);
                c0s[u] = a;
                fs[u] = b - a;
            }
        }
#pragma unroll
        for (int u = 0; u < kEncU; ++u) {
            if (jb - 1 - u < 0) break;                                // warp-uniform
            bool emit = false;
            uint32_t word = 0;
            const uint32_t f = fs[u];
            if (f != 0u) {
                if (x >= (f << 16)) { emit = true; word = x & 0xffffu; x >>= 16; }
                x = ((x / f) << 16) + (x % f) + c0s[u];
            }
            unsigned mask = __ballot_sync(0xffffffffu, emit);
            int cnt = __popc(mask);
            if (emit) {
                int rank = __popc(mask & ((1u << lane) - 1u));   // lanes below me come earlier in decode order
                long at = pos - cnt + rank;
                if (at >= 0) words[at] = (uint16_t)word;
            }
            pos -= cnt;
        }
    }
    bad = __any_sync(0xffffffffu, bad) || pos < 0;
    // 32 final states first (lane 0 first), little endian
    reinterpret_cast<uint32_t *>(o)[lane] = x;
    // slide the words down so that they follow the states: dst index w <- src index pos + w  (dst < src, ascending order is safe)
    const long nw = wcap - (pos < 0 ? 0 : pos);
    for (long w0 = 0; w0 < nw; w0 += 32) {
        long w = w0 + lane;
        uint16_t v = 0;
        if (w < nw) v = words[pos + w];
        __syncwarp();
        if (w < nw) words[w] = v;
        __syncwarp();
    }
    if (lane == 0) out_nbytes[s] = bad ? -1 : (int32_t)(128 + 2 * nw);
}

// ---- two-phase encoder ----------------------------------------------------------------------------------------------------------
// ONE warp walks a whole stream, so its instruction count IS the coder's speed (r02o ncu: 510 instructions and 1350 cycles per
// 32-symbol iteration, IPC 0.38, 4.2 ms for a 196 608-symbol stream while 144 SMs idle).  Everything that does not depend on the
// coder state is therefore moved out of that warp:
//   phase A (rans_prepare_kernel, one thread per symbol, the whole GPU): symbol -> table row -> widened (c0, f) and r = floor(2^32 / f),
//           packed in 8 bytes: lo = c0 | (f - 1) << 16 (c0 <= 65535 and 1 <= f <= 65536 always), hi = r;  0xFFFFFFFF = bad symbol
//   phase B (rans_encode_prepared_kernel, one warp per stream): per iteration one 8-byte load per lane (prefetched 8 deep) and the
//           state update with the division replaced by q = umulhi(x, r), one or two corrections: exact for every 32-bit x
//           (q_est >= floor(x / f) - 1 because r >= 2^32 / f - 1).
// Same bytes as the single-kernel encoder above (kept: sic_rans_encode without a workspace) and as the host coder.
__global__ void __launch_bounds__(256) rans_prepare_kernel(const int32_t *__restrict__ sym, const uint16_t *__restrict__ tables,
                                                           const int32_t *__restrict__ Ls, int n_streams, long n, long sym_per_row,
                                                           long rows_per_stream, int stride, uint2 *__restrict__ prep) {
    const long t = (long)blockIdx.x * 256 + threadIdx.x;
    if (t >= (long)n_streams * n) return;
    const int s = (int)(t / n);
    const long i = t - (long)s * n;
    const int Lraw = Ls[s];
    uint2 o = make_uint2(0xFFFFFFFFu, 0u);
    if (Lraw >= 1 && Lraw <= kMaxL && Lraw <= stride - 1) {
        const uint32_t L = (uint32_t)Lraw;
        const int32_t v = sym[t];
        if (v >= 0 && (uint32_t)v < L) {
            const uint16_t *row = tables + ((long)s * rows_per_stream + i / sym_per_row) * stride;
            const uint32_t c0 = widen(row[v], (uint32_t)v, L), c1 = widen(row[v + 1], (uint32_t)v + 1, L);
            const uint32_t f = c1 - c0;
            o.x = c0 | ((f - 1u) << 16);
            o.y = f == 1u ? 0xFFFFFFFFu : (uint32_t)((1ull << 32) / f);
        }
    }
    prep[t] = o;
}

__global__ void __launch_bounds__(kWarps * 32) rans_encode_prepared_kernel(const uint2 *__restrict__ prep, const int32_t *__restrict__ Ls,
                                                                           int n_streams, long n, int stride, uint8_t *out, long cap,
                                                                           int32_t *__restrict__ out_nbytes) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int s = blockIdx.x * kWarps + warp;
    if (s >= n_streams) return;
    const uint2 *pp = prep + (long)s * n;
    uint8_t *o = out + (long)s * cap;
    uint16_t *words = reinterpret_cast<uint16_t *>(o + 128);
    const long wcap = (cap - 128) / 2;
    long pos = wcap;  // words are written at [pos, wcap), growing downwards
    uint32_t x = kLow;
    bool bad = false;
    if (Ls[s] < 1 || Ls[s] > kMaxL || Ls[s] > stride - 1) {
        if (lane == 0) out_nbytes[s] = -1;
        return;
    }
    const long iters = (n + 31) / 32;
    constexpr int U = 8;
    uint2 nxt[U];
    auto fetch = [&](long jb) {                                       // the U iterations jb-1 ... jb-U (descending)
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long i = (jb - 1 - u) * 32 + lane;
            nxt[u] = make_uint2(0u, 0u);                              // hi == 0: no symbol here (r >= 65536 for every real symbol)
            if (jb - 1 - u >= 0 && i < n) nxt[u] = __ldcs(pp + i);
        }
    };
    fetch(iters);
    for (long jb = iters; jb > 0; jb -= U) {
        uint2 e[U];
#pragma unroll
        for (int u = 0; u < U; ++u) e[u] = nxt[u];
        fetch(jb - U);                                                // the next batch's loads fly while this batch's states update
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (jb - 1 - u < 0) break;                                // warp-uniform
            bool emit = false;
            uint32_t word = 0;
            if (e[u].x == 0xFFFFFFFFu) {
                bad = true;
            } else if (e[u].y != 0u) {
                const uint32_t c0 = e[u].x & 0xffffu, f = (e[u].x >> 16) + 1u, r = e[u].y;
                if (x >= (f << 16)) { emit = true; word = x & 0xffffu; x >>= 16; }   // 32-bit shift exactly as the host coder (rans_host.cpp:63)
                uint32_t q = __umulhi(x, r), rem = x - q * f;
                if (rem >= f) { ++q; rem -= f; }
                if (rem >= f) { ++q; rem -= f; }
                x = (q << 16) + rem + c0;
            }
            unsigned mask = __ballot_sync(0xffffffffu, emit);
            int cnt = __popc(mask);
            if (emit) {
                int rank = __popc(mask & ((1u << lane) - 1u));
                long at = pos - cnt + rank;
                if (at >= 0) words[at] = (uint16_t)word;
            }
            pos -= cnt;
        }
    }
    bad = __any_sync(0xffffffffu, bad) || pos < 0;
    reinterpret_cast<uint32_t *>(o)[lane] = x;
    const long nw = wcap - (pos < 0 ? 0 : pos);
    for (long w0 = 0; w0 < nw; w0 += 32) {
        long w = w0 + lane;
        uint16_t v = 0;
        if (w < nw) v = words[pos + w];
        __syncwarp();
        if (w < nw) words[w] = v;
        __syncwarp();
    }
    if (lane == 0) out_nbytes[s] = bad ? -1 : (int32_t)(128 + 2 * nw);
}

__global__ void __launch_bounds__(kWarps * 32) rans_decode_kernel(const uint8_t *__restrict__ in, const int32_t *__restrict__ nbytes,
                                                                  const uint16_t *__restrict__ tables, const int32_t *__restrict__ Ls,
                                                                  int n_streams, long n, long sym_per_row, long rows_per_stream,
                                                                  int stride, long cap, int32_t *__restrict__ sym,
                                                                  int32_t *__restrict__ status) {
    extern __shared__ uint32_t srow_all[];  // [kWarps][stride] widened row of the channel in flight (broadcast mode)
    // the next kWin words of the stream, per warp: a lane that renormalises takes its word from here instead of paying a dependent
    // global load in (almost) every iteration; refilled with one coalesced read when fewer than 32 words are left in it
    constexpr int kWin = 256;
    __shared__ uint16_t swin_all[kWarps][kWin];
    // slot -> symbol: a 64-bucket table per staged row (lut[b] = last symbol whose start is <= b << 10; lut[64] = L - 1) brackets the
    // answer in [lut[b], lut[b + 1]], then a binary search inside the bracket: 0-1 steps for the buckets of the distribution's body,
    // <= 5 in the tails (a linear scan from lut[b] walked through up to 30 one-count tail symbols whenever ONE lane of the warp landed
    // there: 149 instructions per iteration, r02ak); the plain binary search over the whole row was ~7 dependent shared-memory reads
    __shared__ uint16_t slut_all[kWarps][66];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int s = blockIdx.x * kWarps + warp;
    if (s >= n_streams) return;
    uint32_t *srow = srow_all + (size_t)warp * stride;
    const uint32_t L = (uint32_t)Ls[s];
    const long nb = nbytes[s];
    const uint8_t *ip = in + (long)s * cap;
    const uint16_t *words = reinterpret_cast<const uint16_t *>(ip + 128);
    const long nw = nb >= 128 ? (nb - 128) / 2 : 0;
    const uint16_t *tab = tables + (long)s * rows_per_stream * stride;
    int32_t *so = sym + (long)s * n;
    // supports come from a container or a dict the caller did not produce: a patch with min > max gives L <= 0, which as
    // uint32 would run the staging loop and the binary search far outside the row
    if (Ls[s] < 1 || Ls[s] > kMaxL || Ls[s] > stride - 1) {
        if (lane == 0) status[s] = SIC_E_BADARG;
        return;
    }
    if (nb < 128 || nb > cap) {
        if (lane == 0) status[s] = SIC_E_TRUNCATED;
        return;
    }
    uint32_t x = reinterpret_cast<const uint32_t *>(ip)[lane];
    long pos = 0;
    bool trunc = false;
    uint16_t *swin = swin_all[warp];
    long wbase = 0;
    auto refill = [&](long base) {
        __syncwarp();
        for (int k = lane; k < kWin; k += 32) swin[k] = base + k < nw ? words[base + k] : (uint16_t)0;
        __syncwarp();
        wbase = base;
    };
    refill(0);
    const bool staged = sym_per_row % 32 == 0;   // then the 32 symbols of an iteration share one row
    const bool small = n < (1L << 31) && sym_per_row < (1L << 31);   // 32-bit index arithmetic (a 64-bit division is ~120 instructions)
    long cur_row = -1, next_row = 0;
    const long its_per_row = staged ? sym_per_row / 32 : 0;
    long it_in_row = 0;
    const long iters = (n + 31) / 32;
    for (long j = 0; j < iters; ++j) {
        const long i = j * 32 + lane;
        if (staged) {
            if (it_in_row == its_per_row) { it_in_row = 0; ++next_row; }       // staged: every row spans sym_per_row / 32 whole iterations
            ++it_in_row;
            const long r = next_row;
            if (r != cur_row) {
                __syncwarp();
                const uint16_t *row = tab + r * stride;
                for (uint32_t k = lane; k <= L; k += 32) srow[k] = widen(row[k], k, L);
                cur_row = r;
                __syncwarp();
                for (uint32_t b = lane; b < 64u; b += 32) {          // two buckets per lane, binary search each
                    const uint32_t t0 = b << 10;
                    uint32_t lo = 0, hi = L;
                    while (hi - lo > 1) {
                        uint32_t mid = (lo + hi) >> 1;
                        if (srow[mid] <= t0) lo = mid; else hi = mid;
                    }
                    slut_all[warp][b] = (uint16_t)lo;
                }
                if (lane == 0) slut_all[warp][64] = (uint16_t)(L - 1);
                __syncwarp();
            }
        }
        bool need = false;
        if (i < n) {
            const uint32_t slot = x & 0xffffu;
            uint32_t lo = 0, hi = L, c0, c1;
            if (staged) {
                lo = slut_all[warp][slot >> 10];                      // srow[lo] <= (slot >> 10) << 10 <= slot
                hi = (uint32_t)slut_all[warp][(slot >> 10) + 1] + 1u;  // srow[hi] > ((slot >> 10) + 1) << 10 > slot  (srow[L] = 65536)
                while (hi - lo > 1) {
                    uint32_t mid = (lo + hi) >> 1;
                    if (srow[mid] <= slot) lo = mid; else hi = mid;
                }
                c0 = srow[lo];
                c1 = srow[lo + 1];
            } else {
                const uint16_t *row = tab + (small ? (long)((uint32_t)i / (uint32_t)sym_per_row) : i / sym_per_row) * stride;
                while (hi - lo > 1) {
                    uint32_t mid = (lo + hi) >> 1;
                    if (widen(row[mid], mid, L) <= slot) lo = mid; else hi = mid;
                }
                c0 = widen(row[lo], lo, L);
                c1 = widen(row[lo + 1], lo + 1, L);
            }
            x = (c1 - c0) * (x >> 16) + slot - c0;
            need = x < kLow;
            so[i] = (int32_t)lo;
        }
        unsigned mask = __ballot_sync(0xffffffffu, need);
        if (pos + 32 > wbase + kWin) refill(pos);                        // warp-uniform: an iteration takes at most 32 words
        if (need) {
            long at = pos + __popc(mask & ((1u << lane) - 1u));
            if (at < nw) x = (x << 16) | (uint32_t)swin[at - wbase];
            else trunc = true;
        }
        pos += __popc(mask);
    }
    trunc = __any_sync(0xffffffffu, trunc);
    // a stream that decodes to the end must leave every lane at the encoder's initial state and consume every word:
    // anything else is damage (flipped bits, stray words) that would otherwise decode silently to garbage
    const bool corrupt = __any_sync(0xffffffffu, x != kLow) || pos != nw;
    if (lane == 0) status[s] = trunc ? SIC_E_TRUNCATED : (corrupt ? SIC_E_CORRUPT : 0);
}

SIC_REGISTER_KERNEL("rans_encode_kernel", rans_encode_kernel);
SIC_REGISTER_KERNEL("rans_prepare_kernel", rans_prepare_kernel);
SIC_REGISTER_KERNEL("rans_encode_prepared_kernel", rans_encode_prepared_kernel);
SIC_REGISTER_KERNEL("rans_decode_kernel", rans_decode_kernel);
}  // namespace
}  // namespace sic

using namespace sic;

extern "C" int sic_rans_encode(const int32_t *sym, const uint16_t *tables, const int32_t *Ls, int n_streams, long n,
                               long sym_per_row, long rows_per_stream, int stride, uint8_t *out, long cap, int32_t *out_nbytes,
                               void *stream) {
    SIC_CHECK_ARG(n_streams > 0 && n >= 0 && sym_per_row > 0 && rows_per_stream > 0 && stride >= 2, "sic_rans_encode: bad extents");
    SIC_CHECK_ARG(sym && tables && Ls && out && out_nbytes, "sic_rans_encode: null pointer");
    SIC_CHECK_ARG(cap >= 128 + 2 * n && cap % 4 == 0, "sic_rans_encode: cap must be a multiple of 4 and >= 128 + 2n (= %ld)", 128 + 2 * n);
    SIC_CHECK_ARG(((uintptr_t)out & 3) == 0, "sic_rans_encode: out must be 4-byte aligned");
    rans_encode_kernel<<<(n_streams + kWarps - 1) / kWarps, kWarps * 32, 0, (cudaStream_t)stream>>>(
        sym, tables, Ls, n_streams, n, sym_per_row, rows_per_stream, stride, out, cap, out_nbytes);
    SIC_CHECK_LAUNCH("sic_rans_encode");
    return 0;
}

extern "C" size_t sic_rans_encode_workspace_bytes(int n_streams, long n) {
    return (n_streams > 0 && n > 0) ? (size_t)n_streams * (size_t)n * sizeof(uint2) : 0;
}

extern "C" int sic_rans_encode_ws(const int32_t *sym, const uint16_t *tables, const int32_t *Ls, int n_streams, long n,
                                  long sym_per_row, long rows_per_stream, int stride, uint8_t *out, long cap, int32_t *out_nbytes,
                                  void *workspace, size_t workspace_bytes, void *stream) {
    SIC_CHECK_ARG(n_streams > 0 && n >= 0 && sym_per_row > 0 && rows_per_stream > 0 && stride >= 2, "sic_rans_encode_ws: bad extents");
    SIC_CHECK_ARG(sym && tables && Ls && out && out_nbytes, "sic_rans_encode_ws: null pointer");
    SIC_CHECK_ARG(cap >= 128 + 2 * n && cap % 4 == 0, "sic_rans_encode_ws: cap must be a multiple of 4 and >= 128 + 2n (= %ld)", 128 + 2 * n);
    SIC_CHECK_ARG(((uintptr_t)out & 3) == 0, "sic_rans_encode_ws: out must be 4-byte aligned");
    if (n == 0) return sic_rans_encode(sym, tables, Ls, n_streams, n, sym_per_row, rows_per_stream, stride, out, cap, out_nbytes, stream);
    SIC_CHECK_ARG(workspace != nullptr && ((uintptr_t)workspace & 7) == 0, "sic_rans_encode_ws: workspace must be 8-byte aligned");
    if (workspace_bytes < sic_rans_encode_workspace_bytes(n_streams, n)) {
        set_error("sic_rans_encode_ws: workspace %zu < %zu bytes", workspace_bytes, sic_rans_encode_workspace_bytes(n_streams, n));
        return SIC_E_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    uint2 *prep = static_cast<uint2 *>(workspace);
    const long total = (long)n_streams * n;
    SIC_CHECK_ARG((total + 255) / 256 < (1L << 31), "sic_rans_encode_ws: too many symbols");
    rans_prepare_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(sym, tables, Ls, n_streams, n, sym_per_row, rows_per_stream, stride, prep);
    SIC_CHECK_LAUNCH("sic_rans_encode_ws (prepare)");
    rans_encode_prepared_kernel<<<(n_streams + kWarps - 1) / kWarps, kWarps * 32, 0, st>>>(prep, Ls, n_streams, n, stride, out, cap, out_nbytes);
    SIC_CHECK_LAUNCH("sic_rans_encode_ws");
    return 0;
}

extern "C" int sic_rans_decode(const uint8_t *in, const int32_t *nbytes, const uint16_t *tables, const int32_t *Ls, int n_streams,
                               long n, long sym_per_row, long rows_per_stream, int stride, long cap, int32_t *sym, int32_t *status,
                               void *stream) {
    SIC_CHECK_ARG(n_streams > 0 && n >= 0 && sym_per_row > 0 && rows_per_stream > 0 && stride >= 2 && stride <= kMaxL + 1,
                  "sic_rans_decode: bad extents");
    SIC_CHECK_ARG(in && nbytes && tables && Ls && sym && status, "sic_rans_decode: null pointer");
    SIC_CHECK_ARG(cap % 4 == 0 && ((uintptr_t)in & 3) == 0, "sic_rans_decode: streams must be 4-byte aligned (cap %% 4 == 0)");
    size_t smem = (size_t)kWarps * stride * sizeof(uint32_t);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(rans_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            set_error("sic_rans_decode: cannot reserve %zu B of shared memory: %s", smem, cudaGetErrorString(e));
            return (int)e;
        }
    }
    rans_decode_kernel<<<(n_streams + kWarps - 1) / kWarps, kWarps * 32, smem, (cudaStream_t)stream>>>(
        in, nbytes, tables, Ls, n_streams, n, sym_per_row, rows_per_stream, stride, cap, sym, status);
    SIC_CHECK_LAUNCH("sic_rans_decode");
    return 0;
}
