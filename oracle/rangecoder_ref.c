/* CPU oracle (TEST INFRASTRUCTURE) for the entropy coder, format "SIC-RANS-1" (DESIGN.md).
 *
 * The reference codes with torchac 0.9.3 (eval_selfcontained_entropy.py:48,62,96,116), a third-party module
 * that is neither vendored under /root/reference nor installed, and whose API the call sites misuse
 *  => PARITY UNPINNED.  This file is a plain sequential statement of our own
 * format; the product encoder/decoder (host C++ and CUDA) are written separately and must produce/consume
 * identical bytes.
 *
 * Format, one stream = one (patch, latent) pair:
 *   symbols s_0..s_{n-1} in (C,h,w) row-major order; symbol i uses table row i / sym_per_row.
 *   table row = uint16 cdf[L+1] from pmf_to_uint16_cdf (cdf[0]=0, cdf[L]=65535), possibly with zero-width
 *   symbols; the coder widens it to 16-bit totals with every symbol >= 1 count:
 *        c'_k = floor(c_k * (65536 - L) / 65535) + k ,  k = 0..L      (c'_0 = 0, c'_L = 65536)
 *   32 interleaved rANS lanes (32-bit state in [2^16, 2^32), 16-bit words, 16 probability bits);
 *   lane l owns symbols i with i mod 32 == l.  Decoder order: read 32 initial states (uint32 LE, lane 0 first),
 *   then for j = 0.., for l = 0..31: decode symbol 32j+l, and if the state fell below 2^16 pull the next
 *   uint16 LE word.  The encoder runs that schedule backwards and reverses its output.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline uint32_t widen(uint32_t c, uint32_t k, uint32_t L) { return (c * (65536u - L)) / 65535u + k; }

/* returns number of bytes written, or -1 if cap is too small / a symbol is out of range */
long sic_oracle_rans_encode(const int32_t *sym, long n, const uint16_t *tables, int stride, int L, long sym_per_row,
                            uint8_t *out, long cap) {
    uint32_t state[32];
    for (int l = 0; l < 32; ++l) state[l] = 1u << 16;
    uint16_t *words = (uint16_t *)malloc(sizeof(uint16_t) * (size_t)(n + 1));
    long nw = 0;
    for (long i = n - 1; i >= 0; --i) {
        int l = (int)(i & 31);
        int32_t s = sym[i];
        if (s < 0 || s >= L) { free(words); return -1; }
        const uint16_t *row = tables + (size_t)(i / sym_per_row) * stride;
        uint32_t c0 = widen(row[s], (uint32_t)s, (uint32_t)L);
        uint32_t c1 = widen(row[s + 1], (uint32_t)s + 1, (uint32_t)L);
        uint32_t f = c1 - c0;
        uint32_t x = state[l];
        if (x >= (f << 16)) { words[nw++] = (uint16_t)(x & 0xffffu); x >>= 16; }
        state[l] = ((x / f) << 16) + (x % f) + c0;
    }
    long total = 128 + 2 * nw;
    if (total > cap) { free(words); return -1; }
    for (int l = 0; l < 32; ++l) {
        uint32_t x = state[l];
        out[4 * l + 0] = (uint8_t)(x); out[4 * l + 1] = (uint8_t)(x >> 8);
        out[4 * l + 2] = (uint8_t)(x >> 16); out[4 * l + 3] = (uint8_t)(x >> 24);
    }
    for (long w = 0; w < nw; ++w) {
        uint16_t v = words[nw - 1 - w];
        out[128 + 2 * w] = (uint8_t)v; out[128 + 2 * w + 1] = (uint8_t)(v >> 8);
    }
    free(words);
    return total;
}

/* returns 0 on success, -1 on a truncated stream */
int sic_oracle_rans_decode(const uint8_t *in, long nbytes, long n, const uint16_t *tables, int stride, int L,
                           long sym_per_row, int32_t *sym) {
    if (nbytes < 128) return -1;
    uint32_t state[32];
    for (int l = 0; l < 32; ++l)
        state[l] = (uint32_t)in[4 * l] | ((uint32_t)in[4 * l + 1] << 8) | ((uint32_t)in[4 * l + 2] << 16) |
                   ((uint32_t)in[4 * l + 3] << 24);
    long pos = 128;
    for (long i = 0; i < n; ++i) {
        int l = (int)(i & 31);
        const uint16_t *row = tables + (size_t)(i / sym_per_row) * stride;
        uint32_t x = state[l];
        uint32_t slot = x & 0xffffu;
        int lo = 0, hi = L;                       /* largest k with c'_k <= slot */
        while (hi - lo > 1) {
            int mid = (lo + hi) >> 1;
            if (widen(row[mid], (uint32_t)mid, (uint32_t)L) <= slot) lo = mid; else hi = mid;
        }
        uint32_t c0 = widen(row[lo], (uint32_t)lo, (uint32_t)L);
        uint32_t c1 = widen(row[lo + 1], (uint32_t)lo + 1, (uint32_t)L);
        x = (c1 - c0) * (x >> 16) + slot - c0;
        if (x < (1u << 16)) {
            if (pos + 2 > nbytes) return -1;
            x = (x << 16) | (uint32_t)in[pos] | ((uint32_t)in[pos + 1] << 8);
            pos += 2;
        }
        state[l] = x;
        sym[i] = lo;
    }
    return 0;
}
