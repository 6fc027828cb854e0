// E1 (host side): interleaved rANS coder, format SIC-RANS-1.
//
// Stands in for torchac.encode_float_cdf / decode_float_cdf at
// /root/reference/code/modelv2/eval_selfcontained_entropy.py:48,62,96,116 (torchac 0.9.3 is a third-party module that is
// not vendored with the reference; the bitstream format is therefore ours).  32 interleaved lanes so that the same
// stream can be produced/consumed by one warp on the GPU; 32-bit states, 16-bit renormalisation words, 16 probability bits.
// Tables come straight from sic_build_cdf_tables (uint16, may contain zero-width symbols); each row is widened once to
// 16-bit totals with a guaranteed count >= 1 per symbol:  c'_k = floor(c_k * (65536 - L) / 65535) + k.
#include <stdint.h>
#include <string.h>

#include <vector>

#include "sic.h"

namespace sic {
void set_error(const char *fmt, ...);
}

namespace {
constexpr int kLanes = 32;
constexpr uint32_t kLow = 1u << 16;

struct WideRows {
    std::vector<uint32_t> c;  // [rows][L+1]
    int L;
    WideRows(const uint16_t *tables, long rows, int stride, int L_) : c((size_t)rows * (L_ + 1)), L(L_) {
        const uint32_t scale = 65536u - (uint32_t)L;
        for (long r = 0; r < rows; ++r) {
            const uint16_t *src = tables + (size_t)r * stride;
            uint32_t *dst = c.data() + (size_t)r * (L + 1);
            for (int k = 0; k <= L; ++k) dst[k] = (uint32_t)src[k] * scale / 65535u + (uint32_t)k;
        }
    }
    const uint32_t *row(long r) const { return c.data() + (size_t)r * (L + 1); }
};
}  // namespace

extern "C" long sic_rans_encode_host(const int32_t *sym, long n, const uint16_t *tables, int stride, int L, long sym_per_row,
                                     uint8_t *out, long cap) {
    if (!sym || !tables || !out || n < 0 || L < 1 || L > 4096 || stride < L + 1 || sym_per_row < 1) {
        sic::set_error("sic_rans_encode_host: bad argument (n=%ld L=%d stride=%d sym_per_row=%ld)", n, L, stride, sym_per_row);
        return SIC_E_BADARG;
    }
    const long rows = (n + sym_per_row - 1) / sym_per_row;
    WideRows wide(tables, rows, stride, L);
    uint32_t state[kLanes];
    for (int l = 0; l < kLanes; ++l) state[l] = kLow;
    std::vector<uint16_t> emitted;
    emitted.reserve((size_t)n / 2 + 16);
    // the decoder walks symbols forwards, so encode backwards and reverse the emitted words
    long r = rows - 1, next_row_start = r * sym_per_row;
    const uint32_t *cw = rows ? wide.row(r) : nullptr;
    for (long i = n - 1; i >= 0; --i) {
        if (i < next_row_start) { --r; next_row_start -= sym_per_row; cw = wide.row(r); }
        int32_t s = sym[i];
        if (s < 0 || s >= L) {
            sic::set_error("sic_rans_encode_host: symbol %d at %ld outside [0,%d)", s, i, L);
            return SIC_E_BADARG;
        }
        const uint32_t start = cw[s], freq = cw[s + 1] - start;
        uint32_t &x = state[i & (kLanes - 1)];
        if (x >= (freq << 16)) { emitted.push_back((uint16_t)x); x >>= 16; }
        x = ((x / freq) << 16) + (x % freq) + start;
    }
    const long total = 4L * kLanes + 2L * (long)emitted.size();
    if (total > cap) {
        sic::set_error("sic_rans_encode_host: output needs %ld bytes, capacity %ld", total, cap);
        return SIC_E_OVERFLOW;
    }
    uint8_t *p = out;
    for (int l = 0; l < kLanes; ++l, p += 4) {
        uint32_t x = state[l];
        p[0] = (uint8_t)x; p[1] = (uint8_t)(x >> 8); p[2] = (uint8_t)(x >> 16); p[3] = (uint8_t)(x >> 24);
    }
    for (size_t w = emitted.size(); w-- > 0; p += 2) {
        p[0] = (uint8_t)emitted[w];
        p[1] = (uint8_t)(emitted[w] >> 8);
    }
    return total;
}

extern "C" int sic_rans_decode_host(const uint8_t *in, long nbytes, long n, const uint16_t *tables, int stride, int L,
                                    long sym_per_row, int32_t *sym) {
    if (!in || !tables || !sym || n < 0 || L < 1 || L > 4096 || stride < L + 1 || sym_per_row < 1) {
        sic::set_error("sic_rans_decode_host: bad argument (n=%ld L=%d stride=%d sym_per_row=%ld)", n, L, stride, sym_per_row);
        return SIC_E_BADARG;
    }
    if (nbytes < 4L * kLanes) {
        sic::set_error("sic_rans_decode_host: stream shorter than the %d-byte state header", 4 * kLanes);
        return SIC_E_TRUNCATED;
    }
    const long rows = (n + sym_per_row - 1) / sym_per_row;
    WideRows wide(tables, rows, stride, L);
    uint32_t state[kLanes];
    for (int l = 0; l < kLanes; ++l) {
        const uint8_t *p = in + 4 * l;
        state[l] = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
    }
    const uint8_t *p = in + 4 * kLanes, *end = in + nbytes;
    long r = 0, row_end = sym_per_row;
    const uint32_t *cw = rows ? wide.row(0) : nullptr;
    for (long i = 0; i < n; ++i) {
        if (i >= row_end) { ++r; row_end += sym_per_row; cw = wide.row(r); }
        uint32_t &x = state[i & (kLanes - 1)];
        const uint32_t slot = x & 0xffffu;
        int lo = 0, hi = L;  // invariant cw[lo] <= slot < cw[hi]
        while (hi - lo > 1) {
            int mid = (lo + hi) >> 1;
            if (cw[mid] <= slot) lo = mid; else hi = mid;
        }
        x = (cw[lo + 1] - cw[lo]) * (x >> 16) + slot - cw[lo];
        if (x < kLow) {
            if (p + 2 > end) {
                sic::set_error("sic_rans_decode_host: stream truncated at symbol %ld", i);
                return SIC_E_TRUNCATED;
            }
            x = (x << 16) | (uint32_t)p[0] | ((uint32_t)p[1] << 8);
            p += 2;
        }
        sym[i] = lo;
    }
    // every lane must be back at the encoder's initial state and every word consumed (see rans_device.cu)
    for (int l = 0; l < kLanes; ++l) {
        if (state[l] != kLow) {
            sic::set_error("sic_rans_decode_host: lane %d ends in state 0x%x, not 0x%x: damaged stream", l, state[l], kLow);
            return SIC_E_CORRUPT;
        }
    }
    if (p != end && p + 1 != end) {   // an odd trailing byte cannot hold a word
        sic::set_error("sic_rans_decode_host: %ld stray bytes after the last word", (long)(end - p));
        return SIC_E_CORRUPT;
    }
    return 0;
}
