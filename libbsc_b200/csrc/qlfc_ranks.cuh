// libbsc_b200/csrc/qlfc_ranks.cuh -- QLFC stage 1 (backward move-to-front rank of every run), parallel.
// Included by qlfc.cu inside its anonymous namespace (needs SubBlock).
//
// Reference: bsc_qlfc_transform, qlfc.cpp:200-255 (SIMD) / 398-455 (scalar): scanning the runs of
// a sub-block from the end, a run's rank is the position of its symbol in the recency list, a
// symbol met for the first time gets the number of distinct symbols seen so far, and the very
// last run is forced to 1.  Both cases are the same quantity:
//
//     rank(i) = #{ s != c_i : next_s(i) < next_c(i) },   next_s(i) = first run after i with symbol s
//                                                         (infinity if there is none)
//
// which only needs, at every run, the table of next occurrences.  That table is carried across
// tiles of RK_TILE runs with a two-level scheme, so all tiles are processed concurrently:
//   q_rank_first  per tile: first occurrence of each symbol inside the tile
//   q_rank_scan   per sub-block and symbol (256 threads): backward min-scan over the tiles ->
//                 next-occurrence table at every tile end; also emits the MTF-order header table
//                 (symbols by first appearance, duplicate-terminated, qlfc.cpp:252-253) and nsym
//   q_rank_tile   per tile (one warp): walk the tile backwards; per run 2 x LDS.128 per lane +
//                 a warp add-reduction give the rank, then next_c := i.
#pragma once

#define RK_TILE 4096
#define RK_INF  0xffffffffu

__device__ __forceinline__ bool rk_locate(const SubBlock *__restrict__ sbs, u32 nBlocks, u32 tile, u32 &s, u32 &b, u32 &e)
{
    for (s = 0; s < nBlocks; ++s) {
        const u32 tb = sbs[s].tile_base, tn = sbs[s].tiles;
        if (tile >= tb && tile < tb + tn) {
            b = sbs[s].run_begin + (tile - tb) * RK_TILE;
            e = min(b + (u32)RK_TILE, sbs[s].run_end);
            return true;
        }
    }
    return false;
}

__global__ void __launch_bounds__(128) q_rank_first(const u8 *__restrict__ run_sym, const SubBlock *__restrict__ sbs, u32 nBlocks, u32 total_tiles,
                                                    u32 *__restrict__ first_tab)
{
    __shared__ u32 s_first[4][256];
    const u32 warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const u32 tile = blockIdx.x * 4 + warp;
    if (tile >= total_tiles) return;
    u32 s, b, e;
    if (!rk_locate(sbs, nBlocks, tile, s, b, e)) return;
    for (int k = lane; k < 256; k += 32) s_first[warp][k] = RK_INF;
    __syncwarp();
    for (u32 i = b + lane; i < e; i += 32) atomicMin(&s_first[warp][run_sym[i]], i);
    __syncwarp();
    for (int k = lane; k < 256; k += 32) first_tab[(size_t)tile * 256 + k] = s_first[warp][k];
}

// one CTA per sub-block, thread = symbol
__global__ void __launch_bounds__(256) q_rank_scan(SubBlock *__restrict__ sbs, const u32 *__restrict__ first_tab, u32 *__restrict__ next_tab, u8 *__restrict__ mtf_out)
{
    __shared__ u32 s_first[256];
    SubBlock &sb = sbs[blockIdx.x];
    const u32 sym = threadIdx.x;
    u32 nxt = RK_INF;
    for (u32 t = sb.tiles; t-- > 0; ) {
        const size_t o = (size_t)(sb.tile_base + t) * 256 + sym;
        next_tab[o] = nxt;
        nxt = min(first_tab[o], nxt);
    }
    s_first[sym] = nxt;                                   // first occurrence in the whole sub-block
    __syncthreads();
    u32 before = 0, present = 0;
    for (int k = 0; k < 256; ++k) { const u32 f = s_first[k]; before += (f < nxt); present += (f != RK_INF); }
    u8 *mtf = mtf_out + blockIdx.x * 256;
    if (nxt != RK_INF) {                                  // slots past the terminator are never read
        mtf[before] = (u8)sym;
        if (before + 1 == present && present < 256) mtf[present] = (u8)sym;             // duplicate terminator
    }
    if (sym == 0) sb.nsym = present;
}

__global__ void __launch_bounds__(128) q_rank_tile(const u8 *__restrict__ run_sym, u8 *__restrict__ run_rank, const SubBlock *__restrict__ sbs, u32 nBlocks,
                                                   u32 total_tiles, const u32 *__restrict__ next_tab)
{
    __shared__ __align__(16) u32 s_next[4][256];
    const u32 warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const u32 tile = blockIdx.x * 4 + warp;
    if (tile >= total_tiles) return;
    u32 s, b, e;
    if (!rk_locate(sbs, nBlocks, tile, s, b, e)) return;
    const bool last_tile = (e == sbs[s].run_end);
    {
        const uint4 *src = (const uint4 *)(next_tab + (size_t)tile * 256);
        uint4 *dst = (uint4 *)s_next[warp];
        dst[lane * 2] = src[lane * 2]; dst[lane * 2 + 1] = src[lane * 2 + 1];
    }
    __syncwarp();
    const uint4 *mine = (const uint4 *)s_next[warp] + lane * 2;
    for (u32 hi = e; hi > b; ) {
        const u32 cnt = min(32u, hi - b);
        const u32 mysym = lane < cnt ? run_sym[hi - 1 - lane] : 0;     // lane j holds run hi-1-j
        u32 myrank = 0;
        for (u32 j = 0; j < cnt; ++j) {
            const u32 c = __shfl_sync(0xffffffffu, mysym, j);
            const u32 nc = s_next[warp][c];
            const uint4 a = mine[0], q = mine[1];
            u32 k = (a.x < nc) + (a.y < nc) + (a.z < nc) + (a.w < nc) + (q.x < nc) + (q.y < nc) + (q.z < nc) + (q.w < nc);
            k = __reduce_add_sync(0xffffffffu, k);
            __syncwarp();
            if (lane == 0) s_next[warp][c] = hi - 1 - j;
            __syncwarp();
            if (lane == j) myrank = k;
        }
        if (last_tile && hi == e && lane == 0) myrank = 1;                 // qlfc.cpp:249: the final run is coded as rank 1
        if (lane < cnt) run_rank[hi - 1 - lane] = (u8)myrank;
        hi -= cnt;
    }
}
