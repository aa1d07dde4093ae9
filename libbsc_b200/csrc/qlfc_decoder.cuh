// libbsc_b200/csrc/qlfc_decoder.cuh -- QLFC stage 2 DECODER (qlfc.cpp:1672-1927): one warp per sub-block.
// Included by qlfc.cu after qlfc_coder.cuh (uses CoderSmem's layout, the counter-file indices, q_mix/q_up/q_down).
//
// The decoder is one serial recurrence per stream: every decision needs the counters picked by the previous
// decisions, so the warp runs in lock-step (all lanes decode the same thing from shared-memory broadcasts;
// lanes only differ when a run is expanded or the MTF list is rotated) and per-decision LATENCY is everything.
// What the ncu source view of this kernel showed (profiles/r1f_ncu_decoder.txt), and what is done about it:
//   * ~10 % of all stall samples sat on `S2UR SR_CgaCtaId -> ULEA`: ptxas re-derives the base of the CTA's
//     shared-memory window (it depends on the CTA's rank in its cluster) seven times per run instead of
//     keeping it.  All shared-memory traffic of this kernel therefore goes through ld.shared/st.shared on an
//     explicit 32-bit base address that is computed once and made opaque to the compiler (struct SM);
//   * a shared-memory load costs ~34 cycles if its consumer follows immediately, so everything the NEXT run's
//     first decision needs is fetched while the current run is still being decoded:
//       - the front of the MTF list lives in registers (c, m1, m2, m3): the next symbol is always m1;
//       - the two histories of the next symbol are loaded right after the MTF move;
//       - the rank state of the next run depends on one bit not known yet (run < 3): both candidates and
//         their first-decision counters are loaded, the bit selects at the top of the loop;
//       - the run-length state for the most common rank (1) and its counters are loaded before the rank
//         is decoded.
#pragma once

#include <cstddef>

// shared-memory accessors on an explicit base (byte offsets inside CoderSmem)
struct SM {
    u32 b;
    __device__ __forceinline__ u32 ld8(u32 off) const { u32 v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(b + off)); return v; }
    __device__ __forceinline__ u32 ld16(u32 off) const { u32 v; asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(b + off)); return v; }
    __device__ __forceinline__ void st8(u32 off, u32 v) const { asm volatile("st.shared.u8 [%0], %1;" :: "r"(b + off), "r"(v)); }
    __device__ __forceinline__ void st16(u32 off, u32 v) const { asm volatile("st.shared.u16 [%0], %1;" :: "r"(b + off), "r"(v)); }
    // counters (indices in u16 units, as in qlfc_coder.cuh)
    __device__ __forceinline__ int cnt(u32 idx) const { return (int)ld16((u32)offsetof(CoderSmem, s16) + 2u * idx); }
    __device__ __forceinline__ void set(u32 idx, int v) const { st16((u32)offsetof(CoderSmem, s16) + 2u * idx, (u32)v); }
};
constexpr u32 O_RANK_STATE = (u32)offsetof(CoderSmem, rank_state), O_RUN_STATE = (u32)offsetof(CoderSmem, run_state);
constexpr u32 O_TAG_STATE = (u32)offsetof(CoderSmem, tag_state), O_TAG_CHAR = (u32)offsetof(CoderSmem, tag_char);
constexpr u32 O_RANK_HIST = (u32)offsetof(CoderSmem, rankHist), O_RUN_HIST = (u32)offsetof(CoderSmem, runHist);
constexpr u32 O_MTF = (u32)offsetof(CoderSmem, mtf), O_INWIN = (u32)offsetof(CoderSmem, inwin);

// Index (into the counter file) of rare counter `idx` of one kind, through the direct-mapped write-back cache.
// Warp-uniform: every lane performs the same accesses.
__device__ __forceinline__ u32 dcache_get(const SM &sm, u32 val_base, u32 tags_off, short *__restrict__ cold, u32 idx, u32 &misses)
{
    const u32 slot = cache_slot(idx), want = cache_tag(idx);
    const u32 t = sm.ld16(tags_off + 2u * slot);
    if (t != want) {
        if (t) cold[cache_unslot(slot, t)] = (short)sm.cnt(val_base + slot);
        sm.set(val_base + slot, (u16)cold[idx]);
        sm.st16(tags_off + 2u * slot, want);
        ++misses;
    }
    return val_base + slot;
}

// cold path of the decoder's input window (by-value arguments: keeps the coder state in registers)
__device__ __noinline__ void rc_refill_window(const u8 *__restrict__ in, u32 win_addr, u32 base, u32 limit)
{
    const u32 lane = threadIdx.x & 31;
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const u32 o = base + lane * 8 + k;
        const u32 v = o < limit ? in[o] : 0u;
        asm volatile("st.shared.u8 [%0], %1;" :: "r"(win_addr + lane * 8 + k), "r"(v) : "memory");
    }
    __syncwarp();
}

struct Rc2Dec {
    const u8 *in; u32 pos, limit, code, range;
    u32 win; u32 wbase;              // 256-byte shared-memory window [wbase, wbase+256) of the stream (pos is always even)
    __device__ __forceinline__ void refill() { wbase = pos; rc_refill_window(in, win, pos, limit); }
    __device__ __forceinline__ u32 get16() {
        if (pos - wbase >= 256u) refill();
        u32 v; asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(win + (pos - wbase)));
        pos += 2; return v;
    }
};

// one binary decision against three counters whose values (s, c, g) the caller has already loaded
template <int K> __device__ __forceinline__ u32 dec3v(const SM &sm, Rc2Dec &rc, u32 is, u32 ic, u32 ig, int s, int c, int g)
{
    const int p = q_mix<K>(s, c, g);
    if (rc.range < 0x10000u) { rc.range <<= 16; rc.code = (rc.code << 16) | rc.get16(); }
    const u32 r = (rc.range >> 12) * (u32)p;
    if (rc.code >= r) {                                      // warp-uniform branch: one side does everything for its outcome
        rc.code -= r; rc.range -= r;
        sm.set(is, q_down<K, 0>(s)); sm.set(ic, q_down<K, 1>(c)); sm.set(ig, q_down<K, 2>(g));
        return 1u;
    }
    rc.range = r;
    sm.set(is, q_up<K, 0>(s)); sm.set(ic, q_up<K, 1>(c)); sm.set(ig, q_up<K, 2>(g));
    return 0u;
}
template <int K> __device__ __forceinline__ u32 dec3(const SM &sm, Rc2Dec &rc, u32 is, u32 ic, u32 ig)
{
    const int s = sm.cnt(is), c = sm.cnt(ic), g = sm.cnt(ig);
    return dec3v<K>(sm, rc, is, ic, ig, s, c, g);
}
// the plain 1/2 decisions of the stream header
__device__ __forceinline__ u32 dec_half(Rc2Dec &rc)
{
    if (rc.range < 0x10000u) { rc.range <<= 16; rc.code = (rc.code << 16) | rc.get16(); }
    const u32 r = (rc.range >> 12) * 2048u;
    const u32 bit = rc.code >= r;
    rc.code -= bit ? r : 0u; rc.range = bit ? rc.range - r : r;
    return bit;
}

__global__ void __launch_bounds__(32, 1) q_decode2(const u8 *__restrict__ in_all, SubBlock *__restrict__ sbs, short *__restrict__ cold_all,
                                                   const QTables *__restrict__ tables, u8 *__restrict__ out_all, const u32 *__restrict__ sb_list)
{
    extern __shared__ __align__(16) u8 q_smem_raw[];
    coder_smem_init(*reinterpret_cast<CoderSmem *>(q_smem_raw), tables);
    SM sm; sm.b = (u32)__cvta_generic_to_shared(q_smem_raw);
    asm volatile("" : "+r"(sm.b) :: "memory");               // from here on: explicit addresses only (see the header comment)

    const u32 sid = sb_list[blockIdx.x];
    SubBlock &sb = sbs[sid];
    short *cold_s = cold_all + (size_t)blockIdx.x * 2 * COLD_PAD, *cold_c = cold_s + COLD_PAD;
    const u32 lane = threadIdx.x;
    u32 st_cached = 0, st_miss = 0;

    Rc2Dec rc; rc.in = in_all + sb.out_off; rc.pos = 0; rc.limit = sb.out_cap; rc.code = 0; rc.range = 0xffffffffu;
    rc.win = sm.b + O_INWIN; rc.wbase = 0; rc.refill();
    for (int i = 0; i < 3; ++i) rc.code = (rc.code << 16) | rc.get16();
    u32 n = 0; for (int b = 0; b < 32; ++b) n = (n << 1) | dec_half(rc);
    if (n > sb.in_size) { if (lane == 0) sb.result = LIBBSC_DATA_CORRUPT; return; }   // would overrun the output slice

    int ctxRank0 = 0, ctxRank4 = 0, ctxRun = 0, maxRank = 7, avgRank = 0;
    {
        u32 used8 = 0; int prev = -1;
        for (int d = 0; d < 256; ++d) {
            int c = 0;
            for (int bit = 7; bit >= 0; --bit) {
                bool can0, can1; header_options(used8, prev, c, bit, can0, can1);
                if (can0 && can1) c = 2 * c + (int)dec_half(rc);
                else if (can1) c = 2 * c + 1;
                else if (can0) c = 2 * c;
            }
            c &= 255;
            sm.st8(O_MTF + d, (u32)c);
            if (c == prev) { maxRank = ilog2_dev((u32)(d - 1)); break; }
            prev = c; if ((u32)(c >> 3) == lane) used8 |= 1u << (c & 7);
        }
    }
    __syncwarp();

    u8 *out = out_all + sb.in_start;
    u32 c = sm.ld8(O_MTF), m1 = sm.ld8(O_MTF + 1), m2 = sm.ld8(O_MTF + 2), m3 = sm.ld8(O_MTF + 3);
    u32 rhU = sm.ld8(O_RUN_HIST + c);
    u32 st = sm.ld8(O_RANK_STATE + ((ctxRun << 11) | (ctxRank4 << 3) | sm.ld8(O_RANK_HIST + c)));
    int tS = sm.cnt(R_RT_STATE + st), tC = sm.cnt(R_RT_CHAR + c), tG = sm.cnt(R_RT_SHARED);
    u32 st2z = sm.ld8(O_RUN_STATE + ((ctxRank0 << 10) | (ctxRun << 6) | (rhU < 7 ? rhU : 7)));       // run state if rank == 1
    for (u32 i = 0; i < n; ) {
        int rank = 1; u32 b;
        const u32 rhq = rhU < 7 ? rhU : 7;
        const bool plain = avgRank < 32;
        // first-decision counters of the run length, should the rank turn out to be 1 (the rank decisions never touch them)
        const int uS0 = sm.cnt(R_UT_STATE + st2z), uC0 = sm.cnt(R_UT_CHAR + c), uG0 = sm.cnt(R_UT_SHARED);
        if (plain) {
            b = dec3v<K_RANK_T>(sm, rc, R_RT_STATE + st, R_RT_CHAR + c, R_RT_SHARED, tS, tC, tG);
            if (!b) sm.st8(O_RANK_HIST + c, 0);
            else {
                u32 e = 1;
                while ((int)e != maxRank) {
                    b = dec3<K_RANK_E>(sm, rc, R_RE_STATE + st * 8 + e - 1, R_RE_CHAR + c * 8 + e - 1, R_RE_SHARED + e - 1);
                    if (!b) break;
                    if (++e >= 7) break;                                      // e <= maxRank <= 7 in valid streams
                }
                sm.st8(O_RANK_HIST + c, e);
                if (e <= M_MAXE) {
                    const u32 bs = R_RM_STATE + st * M_ROW + (1u << e) - 2u, bc = R_RM_CHAR + c * M_ROW + (1u << e) - 2u, bg = R_WIDE_SHARED + e * 256;
                    for (int bit = (int)e - 1; bit >= 0; --bit) {
                        b = dec3<K_RANK_M>(sm, rc, bs + rank, bc + rank, bg + rank);
                        rank = 2 * rank + (int)b;
                    }
                } else {
                    for (int bit = (int)e - 1; bit >= 0; --bit) {
                        const u32 is = dcache_get(sm, C_STATE_VAL, O_TAG_STATE, cold_s, wide_idx(e, st, rank), st_miss);
                        const u32 ic = dcache_get(sm, C_CHAR_VAL, O_TAG_CHAR, cold_c, wide_idx(e, c, rank), st_miss);
                        st_cached += 2;
                        b = dec3<K_RANK_M>(sm, rc, is, ic, R_WIDE_SHARED + e * 256 + rank);
                        rank = 2 * rank + (int)b;
                    }
                }
            }
        } else {
            rank = 0;
            for (int node = 1, bit = maxRank; bit >= 0; --bit) {
                const u32 is = dcache_get(sm, C_STATE_VAL, O_TAG_STATE, cold_s, wide_idx(8, st, node), st_miss);
                const u32 ic = dcache_get(sm, C_CHAR_VAL, O_TAG_CHAR, cold_c, wide_idx(8, c, node), st_miss);
                st_cached += 2;
                b = dec3<K_RANK_P>(sm, rc, is, ic, R_WIDE_SHARED + 8 * 256 + node);
                node = 2 * node + (int)b; rank = 2 * rank + (int)b;
            }
            sm.st8(O_RANK_HIST + c, (u32)ilog2_dev((u32)rank));
        }
        rank &= 255;
        // push c `rank` places back: mtf[0..rank-1] = mtf[1..rank]; mtf[rank] = c  (qlfc.cpp:1830-1860)
        const u32 cur = c;
        if (rank == 1) { sm.st8(O_MTF, m1); sm.st8(O_MTF + 1, cur); c = m1; m1 = cur; }
        else if (rank == 2) { sm.st8(O_MTF, m1); sm.st8(O_MTF + 1, m2); sm.st8(O_MTF + 2, cur); c = m1; m1 = m2; m2 = cur; }
        else if (rank == 3) { sm.st8(O_MTF, m1); sm.st8(O_MTF + 1, m2); sm.st8(O_MTF + 2, m3); sm.st8(O_MTF + 3, cur); c = m1; m1 = m2; m2 = m3; m3 = cur; }
        else if (rank != 0) {
            __syncwarp();
            for (int basep = 0; basep < rank; basep += 32) {
                const int p = basep + (int)lane;
                const u32 v = sm.ld8(O_MTF + p + 1);
                __syncwarp();
                if (p < rank) sm.st8(O_MTF + p, v);
                __syncwarp();
            }
            if (lane == 0) sm.st8(O_MTF + rank, cur);
            __syncwarp();
            c = sm.ld8(O_MTF); m1 = sm.ld8(O_MTF + 1); m2 = sm.ld8(O_MTF + 2); m3 = sm.ld8(O_MTF + 3);
        }
        // (c, m1, m2, m3) now describe the NEXT run; `cur` is this run's symbol
        const u32 rhRn = sm.ld8(O_RANK_HIST + c), rhUn = sm.ld8(O_RUN_HIST + c);

        avgRank = (avgRank * 124 + rank * 4) >> 7;
        const int rank0 = rank - 1;
        u32 st2 = st2z, run = 1;
        if (rank0 == 0) b = dec3v<K_RUN_T>(sm, rc, R_UT_STATE + st2z, R_UT_CHAR + cur, R_UT_SHARED, uS0, uC0, uG0);
        else {
            st2 = sm.ld8(O_RUN_STATE + ((ctxRank0 << 10) | (ctxRun << 6) | (((u32)rank0 < 7u ? rank0 : 7) << 3) | rhq));
            b = dec3<K_RUN_T>(sm, rc, R_UT_STATE + st2, R_UT_CHAR + cur, R_UT_SHARED);
        }
        // both candidates for the next run's rank state (its ctxRun gets one more bit: run < 3)
        const u32 ctxRank4n = ((ctxRank4 << 2) | ((u32)rank0 < 3u ? rank0 : 3)) & 0xff;
        const u32 ctxRunN = (ctxRun << 1) & 0xf;
        const u32 stA = sm.ld8(O_RANK_STATE + (((ctxRunN | 1u) << 11) | (ctxRank4n << 3) | rhRn)), stB = sm.ld8(O_RANK_STATE + ((ctxRunN << 11) | (ctxRank4n << 3) | rhRn));
        if (!b) sm.st8(O_RUN_HIST + cur, (rhU + 2) >> 2);
        else {
            u32 e = 1;
            for (;;) {
                const u32 k = e - 1;
                if (k < UE_RES) b = dec3<K_RUN_E>(sm, rc, R_UE_STATE + st2 * UE_RES + k, R_UE_CHAR + cur * UE_RES + k, R_UE_SHARED + k);
                else {
                    const u32 is = dcache_get(sm, C_STATE_VAL, O_TAG_STATE, cold_s, ue_idx(st2, k), st_miss);
                    const u32 ic = dcache_get(sm, C_CHAR_VAL, O_TAG_CHAR, cold_c, ue_idx(cur, k), st_miss);
                    st_cached += 2;
                    b = dec3<K_RUN_E>(sm, rc, is, ic, R_UE_SHARED + k);
                }
                if (!b) break;
                if (++e >= 31) break;                                         // corrupt-input guard
            }
            sm.st8(O_RUN_HIST + cur, ((rhU + 3 * e + 3) >> 2) & 255u);
            if (e <= M_MAXE) {
                const u32 bs = R_UM_STATE + st2 * M_ROW + (1u << e) - 2u, bc = R_UM_CHAR + cur * M_ROW + (1u << e) - 2u, bg = R_NARROW_SHARED + e * 32;
                for (int node = 1, bit = (int)e - 1; bit >= 0; --bit) {
                    b = dec3<K_RUN_M>(sm, rc, bs + node, bc + node, bg + node);
                    run = 2 * run + b; node = 2 * node + (int)b;
                }
            } else {
                for (int node = 1, bit = (int)e - 1; bit >= 0; --bit) {
                    const u32 is = dcache_get(sm, C_STATE_VAL, O_TAG_STATE, cold_s, narrow_idx(e, st2, node), st_miss);
                    const u32 ic = dcache_get(sm, C_CHAR_VAL, O_TAG_CHAR, cold_c, narrow_idx(e, cur, node), st_miss);
                    st_cached += 2;
                    b = dec3<K_RUN_M>(sm, rc, is, ic, R_NARROW_SHARED + e * 32 + node);
                    run = 2 * run + b; node = node + 1;                       // qlfc.cpp:1119: linear contexts above 5 bits
                }
            }
        }
        const bool shortRun = run < 3;
        ctxRank0 = ((ctxRank0 << 1) | (rank0 == 0)) & 0x7;
        ctxRank4 = (int)ctxRank4n;
        ctxRun   = (int)(ctxRunN | (shortRun ? 1u : 0u));
        st = shortRun ? stA : stB;
        rhU = rank != 0 ? rhUn : sm.ld8(O_RUN_HIST + c);                       // rank 0 (corrupt input only): same symbol again
        // first-decision counters of the next run (nothing writes the rank counters until then) and its run state for rank 1
        tS = sm.cnt(R_RT_STATE + st); tC = sm.cnt(R_RT_CHAR + c); tG = sm.cnt(R_RT_SHARED);
        st2z = sm.ld8(O_RUN_STATE + ((ctxRank0 << 10) | (ctxRun << 6) | (rhU < 7 ? rhU : 7)));

        if (run > n - i) run = n - i;                                        // never write past n
        if (run <= 32) { if (lane < run) out[i + lane] = (u8)cur; }
        else for (u32 k = lane; k < run; k += 32) out[i + k] = (u8)cur;
        i += run;
    }
    if (lane == 0) { sb.result = (int)n; sb.stat_cached = st_cached; sb.stat_miss = st_miss; }
}
