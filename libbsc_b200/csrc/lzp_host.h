// libbsc_b200/csrc/lzp_host.h -- the INVERSE of libbsc's LZP preprocessing stage, on the host (lzp.cpp:564-674 decoder, 813-887
// chunk container).  BASELINE.json's north star keeps libbsc/lzp on the host; the forward stage stays in the reference (its five
// encoder variants are not restated), but undoing it is cheap and lets bsc_decompress accept blocks made with the reference's
// DEFAULT options (lzpHashSize 15, lzpMinLen 128).  Sequential byte work after the GPU stages; unlike the reference it never
// writes past the caller's capacity (corrupt streams return LIBBSC_DATA_CORRUPT).
#pragma once
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

namespace lzp_host {

enum { kMatchFlag = 0xf2 };

// one chunk: literals, escaped flags (flag, 255) and matches (flag, length bytes: 254 = "254 more and continue") against the
// position last seen under the hash of the previous four bytes
inline int decode_chunk(const unsigned char *in, const unsigned char *in_end, unsigned char *out, long cap, int hash_bits, int min_len)
{
    if (in_end - in < 4) return -5;                                         // LIBBSC_UNEXPECTED_EOB
    if (cap < 4) return -6;
    const uint32_t mask = (uint32_t(1) << hash_bits) - 1u;
    int *seen = (int *)calloc((size_t)mask + 1, sizeof(int));
    if (!seen) return -2;
    long o = 0;
    while (o < 4) out[o++] = *in++;
    auto last4 = [&]() { return uint32_t(out[o - 1]) | (uint32_t(out[o - 2]) << 8) | (uint32_t(out[o - 3]) << 16) | (uint32_t(out[o - 4]) << 24); };
    uint32_t ctx = last4();
    int rc = 0;
    while (in < in_end) {
        const uint32_t slot = ((ctx >> 15) ^ ctx ^ (ctx >> 3)) & mask;
        const int from = seen[slot]; seen[slot] = (int)o;
        const unsigned char b = *in++;
        if (b != kMatchFlag || from <= 0) {                                 // plain literal
            if (o >= cap) { rc = -6; break; }
            out[o++] = b; ctx = (ctx << 8) | b;
            continue;
        }
        if (in >= in_end) { rc = -6; break; }
        if (*in == 255) {                                                   // escaped flag byte
            ++in;
            if (o >= cap) { rc = -6; break; }
            out[o++] = kMatchFlag; ctx = (ctx << 8) | kMatchFlag;
            continue;
        }
        long len = min_len;
        for (;;) { if (in >= in_end) { rc = -6; break; } const unsigned char l = *in++; len += l; if (l != 254) break; }
        if (rc || o + len > cap) { rc = -6; break; }
        for (const unsigned char *src = out + from; len > 0; --len) out[o++] = *src++;   // forward byte copy: source and target may overlap
        ctx = last4();
    }
    free(seen);
    return rc ? rc : (int)o;
}

// bsc_lzp_decompress: byte 0 = number of chunks; if > 1: {int32 rawSize, int32 packedSize} per chunk, then the chunks (raw when equal)
inline int decompress(const unsigned char *in, int n, unsigned char *out, int cap, int hash_bits, int min_len)
{
    if (n < 1) return -5;
    const int chunks = in[0];
    if (chunks == 1) return decode_chunk(in + 1, in + n, out, cap, hash_bits, min_len);
    if (chunks == 0 || n < 1 + 8 * chunks) return -6;
    long ip = 1 + 8L * chunks, op = 0;
    for (int c = 0; c < chunks; ++c) {
        int32_t raw, packed; memcpy(&raw, in + 1 + 8 * c, 4); memcpy(&packed, in + 5 + 8 * c, 4);
        if (raw < 0 || packed < 0 || ip + packed > n || op + raw > cap) return -6;
        int r = packed;
        if (packed != raw) r = decode_chunk(in + ip, in + ip + packed, out + op, raw, hash_bits, min_len);
        else memcpy(out + op, in + ip, (size_t)packed);
        if (r < 0) return r;
        if (r != raw) return -6;
        ip += packed; op += raw;
    }
    return (int)op;
}

}  // namespace lzp_host
