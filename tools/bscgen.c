/* tools/bscgen.c -- deterministic synthetic input generators.
 * Bench/test infrastructure; not part of the libbsc API.  Plain C, no dependencies.
 *   bscgen_rand(seed, out, n)  uniform bytes                    (config C1)
 *   bscgen_skew(seed, out, n)  bytes = floor(256*u^2)           (config C5, ~7.585 bits/byte)
 *   bscgen_text(seed, out, n)  Zipf-ish word text, enwik-shaped (configs C2-C4)
 *   bscgen_adler32(p, n)       checksum helper for the KATs
 */
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>

static uint64_t sm64(uint64_t *S)
{
    uint64_t z = (*S += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

void bscgen_rand(uint64_t seed, unsigned char *o, size_t n)
{
    uint64_t S = seed;
    for (size_t i = 0; i < n; i += 8) {
        uint64_t r = sm64(&S);
        for (int j = 0; j < 8 && i + j < n; ++j) o[i + j] = (unsigned char)(r >> (8 * j));
    }
}

void bscgen_skew(uint64_t seed, unsigned char *o, size_t n)
{
    uint64_t S = seed;
    for (size_t i = 0; i < n; i += 2) {
        uint64_t r = sm64(&S);
        uint64_t a = (uint32_t)r, b = r >> 32;
        o[i] = (unsigned char)((a * a) >> 56);
        if (i + 1 < n) o[i + 1] = (unsigned char)((b * b) >> 56);
    }
}

#ifndef TEXT_VARIANT
#define TEXT_VARIANT 0
#endif

void bscgen_text(uint64_t seed, unsigned char *o, size_t n)
{
    static const char L[] = "etaoinshrdlucmfwypvbgkqjxz";
    enum { W = 8191 };
    unsigned char (*words)[10] = malloc((size_t)W * 10);
    unsigned char *wlen = malloc(W);
    uint64_t S = seed;
    for (int w = 0; w < W; ++w) {
        uint64_t r = sm64(&S);
        int len = 2 + (int)(r & 7); r >>= 3;
        for (int j = 0; j < len; ++j) {
            unsigned u = (unsigned)((r >> (6 * j)) & 63);
            words[w][j] = (unsigned char)L[(u * u * 26) >> 12];
        }
        wlen[w] = (unsigned char)len;
    }
    size_t p = 0; int sLen = 8, inSentence = 0, sentences = 0, start = 1;
    while (p < n) {
        uint64_t r = sm64(&S);
        int k = (int)(r % 13); r >>= 8;
        int rank = (1 << k) + (int)(r & ((1u << k) - 1)) - 1;
        for (int j = 0; j < wlen[rank] && p < n; ++j) {
            unsigned char c = words[rank][j];
            if (j == 0 && start) c = (unsigned char)(c - 32);
            o[p++] = c;
        }
        start = 0;
        if (++inSentence == sLen) {
            if (p < n) o[p++] = '.';
            sLen = 8 + (int)((r >> 20) & 15); inSentence = 0; start = 1;
            ++sentences;
            if (p < n) o[p++] = (sentences % 20 == 0) ? '\n' : ' ';
        } else {
            if (p < n) o[p++] = ' ';
        }
    }
    free(words); free(wlen);
}

uint32_t bscgen_adler32(const unsigned char *p, size_t n)
{
    uint32_t a = 1, b = 0;
    while (n) {
        size_t k = n < 5552 ? n : 5552; n -= k;
        while (k--) { a += *p++; b += a; }
        a %= 65521; b %= 65521;
    }
    return (b << 16) | a;
}
