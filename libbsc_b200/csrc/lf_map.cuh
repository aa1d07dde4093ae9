// libbsc_b200/csrc/lf_map.cuh -- symbol histogram + LF mapping (destination of a stable counting sort by symbol), shared by the
// inverse BWT (bwt_decode.cu) and the inverse sort transform (st_decode.cu).
//
//   unbwt_hist   256-bin histogram of L                                   (reads n)
//   unbwt_lf     per 4 KB tile a warp match-any multisplit ranks the bytes, per-symbol decoupled look-back gives
//                the tile's global offsets in the same pass              (reads n, writes 4n)
#pragma once
#include "common.cuh"

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

// SENTINEL = true: BWT conventions (row `index` is the removed '$' entry, LF counts the '$' row: + 1).
// SENTINEL = false: plain stable counting-sort destination over n rows (inverse ST: cyclic rotations, no sentinel).
template <bool SENTINEL>
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
        goff[d] = (SENTINEL ? 1u : 0u) + cbase[d] + gexcl;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < LF_ITEMS; ++i) {
        u32 off = wbase + 32 * i;
        if (off < valid) {
            u32 li = base + off;                         // index into L
            u32 row = (!SENTINEL || li < index) ? li : li + 1;
            LF[row] = goff[sym[i]] + whist[warp][sym[i]] + pos[i];
        }
    }
    if (SENTINEL && tile == 0 && tid == 0) LF[index] = 0;            // the '$' row; never followed
}

}  // namespace
