// libbsc_b200/csrc/bwt_decode.cu -- inverse Burrows-Wheeler transform on the device.
//
// Replaces bsc_bwt_decode (libbsc/bwt/bwt.cpp:283-332: libsais_unbwt on the CPU,
// libcubwt_unbwt on the GPU, libcubwt.cu:2953-3132).
//
// Conventions: the conceptual matrix of T$ has n+1 rows; row 0 starts with '$',
// and L is its last column with the '$' entry (row `index`) removed.  LF(row) = 1 + C[c] +
// occ_c(row); walking LF from row 0 yields T backwards.
//
//   unbwt_hist   256-bin histogram of L                                   (reads n)
//   unbwt_lf     LF mapping = destination of a STABLE counting sort by symbol: per 4 KB tile a
//                warp match-any multisplit ranks the bytes, per-symbol decoupled look-back gives
//                the tile's global offsets in the same pass              (reads n, writes 4n)
//   unbwt_walk   K ~ n/64 start rows (one per 64-row window, jittered); one thread per segment chases LF until it
//                reaches another start row.  Pass 1 records (length, successor); a device-side
//                pointer-jumping list ranking (no host round trip, cf. libcubwt.cu:3077-3086)
//                turns that into each segment's text offset; pass 2 walks again and writes the
//                bytes straight to their final place.  This stage is sector/latency bound (one
//                dependent 4-byte gather per output byte), not stream-bandwidth bound.
#include "common.cuh"
#include "stages.cuh"
#include "lf_map.cuh"

namespace {

// Segment starts ("marks"): one row per window of 64 rows, at a pseudo-random offset inside the window.  Evenly spaced marks would
// be correct too, but on structured inputs LF advances rows in arithmetic progressions (a text of period 8 and length 2^26 keeps the
// row's residue mod 64 for millions of steps), and then one thread walks a chain of millions of rows alone.  The reference jitters its
// marks for the same reason (libcubwt.cu:2727-2742).  Window 0 starts at row 0, where the walk begins.
__device__ __forceinline__ u32 mark_row(u32 w, u32 n)
{
    u32 h = w * 0x9E3779B1u; h ^= h >> 15; h *= 0x85EBCA77u; h ^= h >> 13;
    const u32 r = (w << 6) + (w == 0 ? 0u : (h & 63u));
    return r <= n ? r : (w << 6);                        // rows are 0 .. n; the last window may be short
}
__device__ __forceinline__ bool is_mark(u32 row, u32 n) { return mark_row(row >> 6, n) == row; }

// One thread per segment.  EMIT = false: record length and successor.  EMIT = true: write the bytes.
template <bool EMIT>
__global__ void __launch_bounds__(256) unbwt_walk(const u32 *__restrict__ LF, const u8 *__restrict__ L, u32 n, u32 index, u32 K,
                                                  u32 *__restrict__ seg_len, u32 *__restrict__ seg_next, const u32 *__restrict__ seg_dist, u8 *__restrict__ out)
{
    u32 k = blockIdx.x * 256 + threadIdx.x;
    if (k >= K) return;
    u32 row = mark_row(k, n), len = 0, next = K;         // K = sentinel "end of text"
    long long o = 0;
    if (EMIT) { o = (long long)seg_dist[k] - 1; if (o >= (long long)n) o = (long long)n - 1; }
    while (row != index) {
        if (EMIT) { if (o >= 0) out[o] = L[row < index ? row : row - 1]; --o; }
        ++len;
        row = __ldg(LF + row);
        if (is_mark(row, n)) { next = row >> 6; break; }
        if (len > n) break;                              // cannot happen for a permutation; corrupt-input guard
    }
    if (!EMIT) { seg_len[k] = len; seg_next[k] = next; }
}

// Wyllie pointer jumping: dist[k] = number of bytes from the start of segment k to the end.
__global__ void __launch_bounds__(256) unbwt_jump(const u32 *__restrict__ dist_in, const u32 *__restrict__ next_in, u32 *__restrict__ dist_out, u32 *__restrict__ next_out, u32 K)
{
    u32 k = blockIdx.x * 256 + threadIdx.x;
    if (k > K) return;                                   // node K is the sentinel (dist 0, next K)
    u32 nx = next_in[k];
    dist_out[k] = dist_in[k] + dist_in[nx];
    next_out[k] = next_in[nx];
}

__global__ void unbwt_init_sentinel(u32 *dist, u32 *next, u32 K) { dist[K] = 0; next[K] = K; }

}  // namespace

int stage_bwt_decode(Ctx *ctx, u8 *d_T, int n_, int index_)
{
    if (d_T == nullptr || n_ < 0 || index_ <= 0 || index_ > n_) return LIBBSC_BAD_PARAMETER;
    if (n_ <= 1) return LIBBSC_NO_ERROR;
    const u32 n = (u32)n_, index = (u32)index_;
    Arena &A = ctx->arena;
    const size_t mark = A.mark();

    const u32 tiles = ceil_div(n, LF_TILE);
    u8  *Lp   = A.get<u8>((size_t)n + 64);
    u32 *LF   = A.get<u32>((size_t)n + 2);
    u32 *hist = A.get<u32>(256 + 64);
    u64 *lb   = A.get<u64>((size_t)tiles * 256);
    const u32 K = ceil_div((u64)n + 1, 64);                  // one segment per window of 64 rows (mark_row)
    u32 *dist[2] = { A.get<u32>((size_t)K + 1), A.get<u32>((size_t)K + 1) };
    u32 *next[2] = { A.get<u32>((size_t)K + 1), A.get<u32>((size_t)K + 1) };

    CUDA_TRY(cudaMemcpyAsync(Lp, d_T, n, cudaMemcpyDeviceToDevice, ctx->stream));
    CUDA_TRY(cudaMemsetAsync(Lp + n, 0, 64, ctx->stream));
    CUDA_TRY(cudaMemsetAsync(hist, 0, sizeof(u32) * (256 + 64), ctx->stream));
    CUDA_TRY(cudaMemsetAsync(lb, 0, sizeof(u64) * (size_t)tiles * 256, ctx->stream));

    LAUNCH(ctx, unbwt_hist, min(ceil_div(n, 256 * 64), (u32)(B200_SMS * 8)), 256, 0, Lp, n, hist);
    LAUNCH(ctx, unbwt_scan256, 1, 32, 0, hist);
    PROF_BYTES(ctx, 5.0 * n);
    LAUNCH(ctx, unbwt_lf<true>, tiles, LF_THREADS, 0, Lp, n, index, hist, hist + 256, lb, LF);

    PROF_BYTES(ctx, 4.0 * n);
    LAUNCH(ctx, unbwt_walk<false>, ceil_div(K, 256), 256, 0, LF, Lp, n, index, K, dist[0], next[0], (const u32 *)nullptr, (u8 *)nullptr);
    LAUNCH(ctx, unbwt_init_sentinel, 1, 1, 0, dist[0], next[0], K);
    int cur = 0;
    for (u32 span = 1; span < K + 1; span <<= 1) {
        LAUNCH(ctx, unbwt_jump, ceil_div(K + 1, 256), 256, 0, dist[cur], next[cur], dist[cur ^ 1], next[cur ^ 1], K);
        cur ^= 1;
    }
    PROF_BYTES(ctx, 6.0 * n);
    LAUNCH(ctx, unbwt_walk<true>, ceil_div(K, 256), 256, 0, LF, Lp, n, index, K, (u32 *)nullptr, (u32 *)nullptr, dist[cur], d_T);
    A.release(mark);
    return LIBBSC_NO_ERROR;
}
