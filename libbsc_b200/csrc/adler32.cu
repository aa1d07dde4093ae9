// libbsc_b200/csrc/adler32.cu -- Adler-32 of a device buffer.
//
// bsc_compress / bsc_decompress checksum the raw data, the payload and the header
// (libbsc/libbsc/libbsc.cpp:331-333, 347, 545, 616; libbsc/adler32/adler32.cpp:83-203).  The
// reference keeps this on the host (SURVEY.md 8f next #4 suggests moving it); doing it on the
// device lets a whole block stay resident in HBM from H2D to D2H.
//
// Adler-32 is a pair of modular sums, so it is an ordinary parallel reduction:
//   A = 1 + sum d_i            (mod 65521)
//   B = n + sum (n - i) d_i    (mod 65521)
// Each thread handles 64 consecutive bytes with 128-bit loads (weights relative to the chunk end),
// the CTA reduces and adds its (A, B) contribution with two 64-bit atomics.
#include "common.cuh"
#include "stages.cuh"

#define ADLER_MOD 65521u
#define AD_THREADS 256
#define AD_CHUNK   64

namespace {

__global__ void __launch_bounds__(AD_THREADS) adler_partial(const u8 *__restrict__ p, u32 n, unsigned long long *__restrict__ acc)
{
    __shared__ unsigned long long s_a[AD_THREADS / 32], s_b[AD_THREADS / 32];
    const u64 first = ((u64)blockIdx.x * AD_THREADS + threadIdx.x) * AD_CHUNK;
    u32 a = 0; u64 b = 0;
    if (first < n) {
        const u32 len = (u32)min((u64)AD_CHUNK, (u64)n - first);
        u32 wsum = 0;                                     // sum (len - k) * d_k, k = 0..len-1  (<= 64*64*255)
        if (len == AD_CHUNK && ((((size_t)(p + first)) & 15) == 0)) {
#pragma unroll
            for (int v = 0; v < AD_CHUNK / 16; ++v) {
                uint4 q = ld_stream_v4(p + first + 16 * v);
                u32 w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int k = 0; k < 4; ++k) { u32 d = (w[j] >> (8 * k)) & 255u; a += d; wsum += (u32)(AD_CHUNK - (16 * v + 4 * j + k)) * d; }
            }
        } else {
            for (u32 k = 0; k < len; ++k) { u32 d = p[first + k]; a += d; wsum += (len - k) * d; }
        }
        // weight of byte i in B is (n - i) = (len - k) + (n - first - len)
        u64 tail = ((u64)n - first - len) % ADLER_MOD;
        b = (u64)wsum + (u64)a * tail;                    // < 2^20 + 2^14 * 2^16
    }
    unsigned long long ra = a, rb = b % ADLER_MOD;
    for (int o = 16; o; o >>= 1) { ra += __shfl_xor_sync(0xffffffffu, ra, o); rb += __shfl_xor_sync(0xffffffffu, rb, o); }
    if ((threadIdx.x & 31) == 0) { s_a[threadIdx.x >> 5] = ra; s_b[threadIdx.x >> 5] = rb; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < AD_THREADS / 32; ++i) { ra += s_a[i]; rb += s_b[i]; }
        atomicAdd(acc, ra % ADLER_MOD); atomicAdd(acc + 1, rb % ADLER_MOD);
    }
}

__global__ void adler_final(unsigned long long *acc, u32 n, u32 *out)
{
    u32 A = (u32)((1 + acc[0]) % ADLER_MOD);
    u32 B = (u32)((n % ADLER_MOD + acc[1]) % ADLER_MOD);
    *out = (B << 16) | A;
    acc[0] = 0; acc[1] = 0;                               // ready for the next use on this stream
}

}  // namespace

// d_mail words 60..63 hold the two 64-bit accumulators (kept zero between calls).
void stage_adler32_async(Ctx *ctx, const u8 *d_p, int n, int slot)
{
    unsigned long long *acc = (unsigned long long *)(ctx->d_mail + 60);
    if (n > 0) {
        u32 blocks = ceil_div((u64)n, (u64)AD_THREADS * AD_CHUNK);
        LAUNCH(ctx, adler_partial, blocks, AD_THREADS, 0, d_p, (u32)n, acc);
    }
    LAUNCH(ctx, adler_final, 1, 1, 0, acc, (u32)(n > 0 ? n : 0), ctx->d_mail + slot);
}

u32 stage_adler32(Ctx *ctx, const u8 *d_p, int n)
{
    stage_adler32_async(ctx, d_p, n, 0);
    ctx->fetch_mail(1);
    return ctx->h_mail[0];
}
