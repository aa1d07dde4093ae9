// libbsc_b200/csrc/list_rank.cuh -- in-place Wyllie pointer jumping on 64-bit (successor | distance << 32) pairs, shared by the inverse
// BWT (bwt_decode.cu: lists of segment nodes) and the inverse sort transform (st_decode.cu: lists of rows).
//
// pair[i] = successor | weight << 32.  A list END points to itself and carries LR_DONE in its distance word (its own weight counts).
// After convergence every element holds (its list's end, sum of the weights from itself to the end inclusive of the end's) | LR_DONE.
// Any 64-bit snapshot of pair[j] is a valid (successor, distance) statement, so reading a pair that another thread has already
// advanced in the same round only speeds convergence up: no ping-pong buffers.  Rounds after convergence return at once (device-side
// flag per round, no host round trip); launch ceil(log2(longest list)) + 1 of them.
#pragma once
#include "common.cuh"

#define LR_DONE 0x80000000u

namespace {

__global__ void __launch_bounds__(256) lr_jump(u64 *pair, u32 n, u32 *flags, int round)
{
    if (round > 0 && flags[round - 1] == 0) return;     // converged in an earlier round (uniform for the whole grid)
    const u32 i = blockIdx.x * 256 + threadIdx.x;
    bool more = false;
    if (i < n) {
        const u64 p = ld_relaxed(pair + i);
        const u32 d = (u32)(p >> 32);
        if (!(d & LR_DONE)) {
            const u32 nx = (u32)p;
            const u64 q = ld_relaxed(pair + nx);
            const u32 qd = (u32)(q >> 32);
            const u32 nd = (d + (qd & ~LR_DONE)) | (qd & LR_DONE);
            st_relaxed(pair + i, (u64)(u32)q | ((u64)nd << 32));
            more = !(qd & LR_DONE);
        }
    }
    // ONE store per CTA: two million threads storing to the same word took 0.12 ms per round (profiles/r2g_call_g.log)
    if (__syncthreads_or(more) && threadIdx.x == 0) flags[round] = 1;
}

}  // namespace

// enqueue enough rounds for lists of up to `longest` elements; flags: >= 40 zeroed words
static inline void lr_rank(Ctx *ctx, u64 *pair, u32 n, u32 longest, u32 *flags)
{
    int rounds = 1; while ((1ull << rounds) < (u64)longest) ++rounds;
    for (int r = 0; r <= rounds; ++r) LAUNCH(ctx, lr_jump, ceil_div(n, 256), 256, 0, pair, n, flags, r);
}
