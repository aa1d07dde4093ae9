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
//   unbwt_walk   K ~ n/64 evenly spaced start rows; one thread per segment chases LF until it
//                reaches another start row.  Pass 1 records (length, successor); a device-side
//                pointer-jumping list ranking (no host round trip, cf. libcubwt.cu:3077-3086)
//                turns that into each segment's text offset; pass 2 walks again and writes the
//                bytes straight to their final place.  This stage is sector/latency bound (one
//                dependent 4-byte gather per output byte), not stream-bandwidth bound.
#include "common.cuh"
#include "stages.cuh"

#define LF_THREADS 256
#define LF_ITEMS   16
#define LF_TILE    (LF_THREADS * LF_ITEMS)
#define LF_WARPS   (LF_THREADS / 32)

#define LB_FLAG_AGG    (1ull << 62)
#define LB_FLAG_PREFIX (2ull << 62)
#define LB_FLAG_MASK   (3ull << 62)

namespace {

__global__ void __launch_bounds__(256) unbwt_hist(const u8 *__restrict__ L, u32 n, u32 *__restrict__ ghist)
{
    __shared__ u32 sh[8][256];
    for (int i = threadIdx.x; i < 8 * 256; i += 256) (&sh[0][0])[i] = 0;
    __syncthreads();
    u32 *mine = sh[threadIdx.x >> 5];
    const u32 nvec = n / 16;
    const uint4 *V = (const uint4 *)L;
    for (u32 i = blockIdx.x * 256 + threadIdx.x; i < nvec; i += gridDim.x * 256) {
        uint4 v = ld_stream_v4(V + i);
        u32 w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            atomicAdd(&mine[w[j] & 255], 1u); atomicAdd(&mine[(w[j] >> 8) & 255], 1u);
            atomicAdd(&mine[(w[j] >> 16) & 255], 1u); atomicAdd(&mine[w[j] >> 24], 1u);
        }
    }
    if (blockIdx.x == 0) for (u32 i = nvec * 16 + threadIdx.x; i < n; i += 256) atomicAdd(&mine[L[i]], 1u);
    __syncthreads();
    u32 s = 0;
    for (int w = 0; w < 8; ++w) s += sh[w][threadIdx.x];
    if (s) atomicAdd(&ghist[threadIdx.x], s);
}

// exclusive scan of 256 counters, one warp
__global__ void unbwt_scan256(u32 *h)
{
    u32 lane = threadIdx.x, v[8], sum = 0;
    for (int j = 0; j < 8; ++j) { v[j] = h[lane * 8 + j]; sum += v[j]; }
    u32 incl = sum;
    for (int o = 1; o < 32; o <<= 1) { u32 t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (u32)o) incl += t; }
    u32 run = incl - sum;
    for (int j = 0; j < 8; ++j) { h[lane * 8 + j] = run; run += v[j]; }
}

__global__ void __launch_bounds__(LF_THREADS, 4)
unbwt_lf(const u8 *__restrict__ L, u32 n, u32 index, const u32 *__restrict__ cbase, u32 *tile_counter, u64 *lookback, u32 *__restrict__ LF)
{
    __shared__ u32 whist[LF_WARPS][256];
    __shared__ u32 goff[256];
    __shared__ __align__(16) u8 bytes[LF_TILE];
    __shared__ u32 s_tile;
    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (tid == 0) s_tile = atomicAdd(tile_counter, 1u);
    for (int i = tid; i < LF_WARPS * 256; i += LF_THREADS) (&whist[0][0])[i] = 0;
    __syncthreads();
    const u32 tile = s_tile, base = tile * LF_TILE;
    const u32 valid = min((u32)LF_TILE, n - base);
    {   // cooperative 16-byte loads of the tile (L is 16-byte aligned and padded)
        uint4 v = make_uint4(0, 0, 0, 0);
        if (tid * 16 < valid) v = ld_stream_v4(L + base + tid * 16);
        ((uint4 *)bytes)[tid] = v;
    }
    __syncthreads();

    const u32 wbase = warp * (32 * LF_ITEMS) + lane;
    u32 sym[LF_ITEMS], pos[LF_ITEMS];
#pragma unroll
    for (int i = 0; i < LF_ITEMS; ++i) {
        u32 off = wbase + 32 * i;
        bool ok = off < valid;
        u32 d = ok ? bytes[off] : 256u;                  // 256 = "no item" (never matches a symbol)
        sym[i] = d;
        u32 m = __match_any_sync(0xffffffffu, d);
        u32 leader = __ffs(m) - 1, pre = 0;
        if (ok && lane == leader) { pre = whist[warp][d]; whist[warp][d] = pre + __popc(m); }
        pre = __shfl_sync(0xffffffffu, pre, leader);
        pos[i] = pre + __popc(m & lanemask_lt());
        __syncwarp();
    }
    __syncthreads();
    {
        const u32 d = tid;
        u32 run = 0;
#pragma unroll
        for (int w = 0; w < LF_WARPS; ++w) { u32 t = whist[w][d]; whist[w][d] = run; run += t; }
        const u32 count = run;
        u64 *mine = lookback + (size_t)tile * 256 + d;
        u32 gexcl = 0;
        if (tile == 0) st_relaxed(mine, LB_FLAG_PREFIX | (u64)count);
        else {
            st_relaxed(mine, LB_FLAG_AGG | (u64)count);
            for (u32 t = tile; t-- > 0; ) {
                const u64 *theirs = lookback + (size_t)t * 256 + d;
                u64 v;
                do { v = ld_relaxed(theirs); } while ((v & LB_FLAG_MASK) == 0);
                gexcl += (u32)v;
                if ((v & LB_FLAG_MASK) == LB_FLAG_PREFIX) break;
            }
            st_relaxed(mine, LB_FLAG_PREFIX | (u64)(gexcl + count));
        }
        goff[d] = 1u + cbase[d] + gexcl;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < LF_ITEMS; ++i) {
        u32 off = wbase + 32 * i;
        if (off < valid) {
            u32 li = base + off;                         // index into L
            u32 row = li < index ? li : li + 1;
            LF[row] = goff[sym[i]] + whist[warp][sym[i]] + pos[i];
        }
    }
    if (tile == 0 && tid == 0) LF[index] = 0;            // the '$' row; never followed
}

// One thread per segment.  EMIT = false: record length and successor.  EMIT = true: write the bytes.
template <bool EMIT>
__global__ void __launch_bounds__(256) unbwt_walk(const u32 *__restrict__ LF, const u8 *__restrict__ L, u32 n, u32 index, u32 stride, u32 K,
                                                  u32 *__restrict__ seg_len, u32 *__restrict__ seg_next, const u32 *__restrict__ seg_dist, u8 *__restrict__ out)
{
    u32 k = blockIdx.x * 256 + threadIdx.x;
    if (k >= K) return;
    u32 row = k * stride, len = 0, next = K;             // K = sentinel "end of text"
    long long o = 0;
    if (EMIT) { o = (long long)seg_dist[k] - 1; if (o >= (long long)n) o = (long long)n - 1; }
    while (row != index) {
        if (EMIT) { if (o >= 0) out[o] = L[row < index ? row : row - 1]; --o; }
        ++len;
        row = __ldg(LF + row);
        if (row % stride == 0) { next = row / stride; break; }
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
    u32 stride = 64;
    const u32 K = ceil_div((u64)n + 1, stride);
    u32 *dist[2] = { A.get<u32>((size_t)K + 1), A.get<u32>((size_t)K + 1) };
    u32 *next[2] = { A.get<u32>((size_t)K + 1), A.get<u32>((size_t)K + 1) };

    CUDA_TRY(cudaMemcpyAsync(Lp, d_T, n, cudaMemcpyDeviceToDevice, ctx->stream));
    CUDA_TRY(cudaMemsetAsync(Lp + n, 0, 64, ctx->stream));
    CUDA_TRY(cudaMemsetAsync(hist, 0, sizeof(u32) * (256 + 64), ctx->stream));
    CUDA_TRY(cudaMemsetAsync(lb, 0, sizeof(u64) * (size_t)tiles * 256, ctx->stream));

    LAUNCH(ctx, unbwt_hist, min(ceil_div(n, 256 * 64), (u32)(B200_SMS * 8)), 256, 0, Lp, n, hist);
    LAUNCH(ctx, unbwt_scan256, 1, 32, 0, hist);
    PROF_BYTES(ctx, 5.0 * n);
    LAUNCH(ctx, unbwt_lf, tiles, LF_THREADS, 0, Lp, n, index, hist, hist + 256, lb, LF);

    PROF_BYTES(ctx, 4.0 * n);
    LAUNCH(ctx, unbwt_walk<false>, ceil_div(K, 256), 256, 0, LF, Lp, n, index, stride, K, dist[0], next[0], (const u32 *)nullptr, (u8 *)nullptr);
    LAUNCH(ctx, unbwt_init_sentinel, 1, 1, 0, dist[0], next[0], K);
    int cur = 0;
    for (u32 span = 1; span < K + 1; span <<= 1) {
        LAUNCH(ctx, unbwt_jump, ceil_div(K + 1, 256), 256, 0, dist[cur], next[cur], dist[cur ^ 1], next[cur ^ 1], K);
        cur ^= 1;
    }
    PROF_BYTES(ctx, 6.0 * n);
    LAUNCH(ctx, unbwt_walk<true>, ceil_div(K, 256), 256, 0, LF, Lp, n, index, stride, K, (u32 *)nullptr, (u32 *)nullptr, dist[cur], d_T);
    A.release(mark);
    return LIBBSC_NO_ERROR;
}
