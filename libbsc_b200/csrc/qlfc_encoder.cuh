// libbsc_b200/csrc/qlfc_encoder.cuh -- QLFC stage 2 ENCODER as a three-warp pipeline.
// Included by qlfc.cu after qlfc_coder.cuh (uses CoderSmem, the counter-file layout, Rc2Enc).
//
// Facts this design rests on (qlfc.cpp:829-1129):
//   * every model context is a function of the INPUT (symbols, ranks, run lengths) only -- never of
//     the coder state -- so probabilities can be produced ahead of the range coder;
//   * within one run every binary decision touches a different counter, so all decisions of a run
//     can be evaluated simultaneously, one lane per decision;
//   * the rank decisions and the run-length decisions use disjoint counter arrays, so they can be
//     evaluated by two different warps without any ordering between them;
//   * the only truly serial recurrence is the range coder (range/low, rangecoder.h:83-177).
//
//   warp 0  "rank model":  per run, lane d evaluates rank decision d  (first bit | unary exponent | mantissa / escape)
//   warp 1  "run model" :  per run, lane d evaluates run-length decision d (first bit | unary exponent | mantissa)
//   warp 2  "coder"     :  consumes the (bit, p) records in stream order from a shared-memory ring
//
// Both model warps derive the ring position of every run from the same arithmetic (the number of
// decisions of a run follows from rank, run length, maxRank and the escape flag), so they never
// talk to each other; the coder consumes up to min(progress of warp 0, progress of warp 1).
// The first profiled version of the encoder (one warp doing everything, profiles/r1b) executed
// ~530 instructions per run on the critical warp; here the critical warp executes ~100.
#pragma once

#define QE4_RING 2048
#define QE4_BIT  0x2000u                                    // record: bits 0..12 p, bit 13 the coded bit, bit 14 run start
#define QE4_RUN  0x4000u

struct Enc4Pipe {
    u16 ring[QE4_RING];
    volatile u32 progA, progB, head;                        // records completed by warp 0 / warp 1, records consumed
    volatile u32 hdr_len, hdr_ready, max_rank;              // published by warp 0 after the stream header
    volatile u32 doneA, doneB, fail;
    int prm[7][2][12];                                      // [class][bit] = w0 w1 w2 | Ms Ks | Mc Kc | Mg Kg
};

// parameters of one decision class for one outcome, in the (p*M + K) >> 12 form of the counter moves
__device__ __forceinline__ void enc4_fill_params(Enc4Pipe &P, u32 lane)
{
    for (int i = lane; i < 14; i += 32) {
        const int k = i >> 1, b = i & 1;
        int *q = P.prm[k][b];
        q[0] = bscb_param(k, 0); q[1] = bscb_param(k, 1); q[2] = bscb_param(k, 2);
        for (int who = 0; who < 3; ++who) {
            const int th0 = bscb_param(k, 3 + 4 * who), ar0 = bscb_param(k, 4 + 4 * who), th1 = bscb_param(k, 5 + 4 * who), ar1 = bscb_param(k, 6 + 4 * who);
            q[3 + 2 * who] = b ? 4096 - ar1 : 4096 - ar0;
            q[4 + 2 * who] = b ? th1 * ar1 + 4095 : (4096 - th0) * ar0;
        }
    }
}

// wait until records [pos, pos+cnt) of the ring may be overwritten; false = give up (failure / watchdog)
__device__ __forceinline__ bool enc4_wait_room(Enc4Pipe &P, u32 pos, u32 cnt)
{
    for (u32 spins = 0; (int)(pos + cnt - P.head) > QE4_RING; ) {
        if (P.fail || ++spins > (1u << 26)) { if (!P.fail) P.fail = 2; return false; }
    }
    return true;
}

// Evaluate one decision: counters at s16 indices (is, ic, ig); returns the record.  Rare counters
// (cached == true) go through the direct-mapped caches; the caller guarantees that no two lanes of
// the same call use the same cache slot (else it serialises the lanes).
__device__ __forceinline__ u32 enc4_decide(CoderSmem &S, Enc4Pipe &P, int K, u32 bit, u32 is, u32 ic, u32 ig, bool cached, u32 cs, u32 cc,
                                           short *__restrict__ cold_s, short *__restrict__ cold_c, u32 &misses)
{
    if (cached) {
        const u32 slot_s = cache_slot(cs), slot_c = cache_slot(cc);
        u32 t = S.tag_state[slot_s];
        if (t != cache_tag(cs)) { if (t) cold_s[cache_unslot(slot_s, t)] = (short)S.s16[C_STATE_VAL + slot_s];
                                  S.s16[C_STATE_VAL + slot_s] = (u16)cold_s[cs]; S.tag_state[slot_s] = (u16)cache_tag(cs); ++misses; }
        t = S.tag_char[slot_c];
        if (t != cache_tag(cc)) { if (t) cold_c[cache_unslot(slot_c, t)] = (short)S.s16[C_CHAR_VAL + slot_c];
                                  S.s16[C_CHAR_VAL + slot_c] = (u16)cold_c[cc]; S.tag_char[slot_c] = (u16)cache_tag(cc); ++misses; }
        is = C_STATE_VAL + slot_s; ic = C_CHAR_VAL + slot_c;
    }
    const int *q = P.prm[K][bit];
    const int s = S.s16[is], c = S.s16[ic], g = S.s16[ig];
    const int p = (c * q[0] + s * q[1] + g * q[2]) >> 5;
    S.s16[is] = (u16)((s * q[3] + q[4]) >> 12);
    S.s16[ic] = (u16)((c * q[5] + q[6]) >> 12);
    S.s16[ig] = (u16)((g * q[7] + q[8]) >> 12);
    return (u32)p | (bit ? QE4_BIT : 0u);
}

// run the per-lane decisions of one chunk; lanes with cached counters that collide on a cache slot are serialised
__device__ __forceinline__ u32 enc4_chunk(CoderSmem &S, Enc4Pipe &P, bool act, int K, u32 bit, u32 is, u32 ic, u32 ig, bool cached, u32 cs, u32 cc,
                                          short *__restrict__ cold_s, short *__restrict__ cold_c, u32 lane, u32 &n_cached, u32 &misses)
{
    const u32 cmask = __ballot_sync(0xffffffffu, act && cached);
    bool any_clash = false;
    if (cmask) {                                            // warp-uniform, rare
        bool clash = false;
        if (act && cached) {
            const u32 slot_s = cache_slot(cs), slot_c = cache_slot(cc);
            const u32 ms = __match_any_sync(cmask, slot_s), mc = __match_any_sync(cmask, slot_c);   // both executed by all lanes of cmask
            clash = (__popc(ms) > 1) | (__popc(mc) > 1);
        }
        any_clash = __any_sync(0xffffffffu, clash);
        n_cached += 2 * __popc(cmask);
    }
    u32 rec = 0;
    if (!any_clash) { if (act) rec = enc4_decide(S, P, K, bit, is, ic, ig, cached, cs, cc, cold_s, cold_c, misses); }
    else for (u32 turn = 0; turn < 32; ++turn) {
        if (act && lane == turn) rec = enc4_decide(S, P, K, bit, is, ic, ig, cached, cs, cc, cold_s, cold_c, misses);
        __syncwarp();
    }
    __syncwarp();
    return rec;
}

__global__ void __launch_bounds__(96, 1) q_encode4(const u32 *__restrict__ run_pos, const u8 *__restrict__ run_sym, const u8 *__restrict__ run_rank,
                                                   SubBlock *__restrict__ sbs, const u8 *__restrict__ mtf_all, short *__restrict__ cold_all,
                                                   const QTables *__restrict__ tables, u8 *__restrict__ out_all, const u32 *__restrict__ sb_list)
{
    extern __shared__ __align__(16) u8 q_smem_raw[];
    CoderSmem &S = *reinterpret_cast<CoderSmem *>(q_smem_raw);
    Enc4Pipe &P = *reinterpret_cast<Enc4Pipe *>(q_smem_raw + ((sizeof(CoderSmem) + 15) & ~(size_t)15));
    const u32 warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const u32 sid = sb_list ? sb_list[blockIdx.x] : blockIdx.x;
    SubBlock &sb = sbs[sid];

    if (warp == 0) coder_smem_init(S, tables);
    if (warp == 1) {
        enc4_fill_params(P, lane);
        if (lane == 0) { P.progA = 0; P.progB = 0; P.head = 0; P.hdr_len = 0; P.hdr_ready = 0; P.max_rank = 7; P.doneA = 0; P.doneB = 0; P.fail = 0; }
    }
    __syncthreads();

    const u32 rb = sb.run_begin, re = sb.run_end;

    if (warp == 2) {
        // ---------------------------------------- coder ----------------------------------------
        Rc2Enc rc; rc.low32 = 0; rc.carry = 0; rc.range = 0xffffffffu; rc.cache = 0; rc.pending = 0; rc.pos = 0; rc.out = out_all + sb.out_off;
        const long long eob = (long long)sb.out_cap - 16;
        u32 h = 0; int result = 0; bool eob_hit = eob <= 0;
        for (;;) {
            u32 limit, spins = 0;
            for (;;) {
                const u32 a = P.progA, b = P.progB;
                limit = min(a, b);
                if (limit != h) break;
                if (P.doneA && P.doneB && min(P.progA, P.progB) == h) { limit = h; break; }
                if (P.fail == 2 || ++spins > (1u << 27)) { result = LIBBSC_GPU_ERROR; break; }
            }
            if (result || limit == h) break;
            __threadfence_block();
            u32 nxt = P.ring[h & (QE4_RING - 1)];
            while (h != limit) {
                const u32 rec = nxt; ++h;
                nxt = P.ring[h & (QE4_RING - 1)];               // prefetch (harmless past `limit`)
                if (eob_hit && (rec & QE4_RUN)) { result = LIBBSC_NOT_COMPRESSIBLE; break; }   // qlfc.cpp:898-901
                if (rc.range < 0x10000u) { rc.shift(); rc.range <<= 16; eob_hit = (long long)rc.pos >= eob; }
                const u32 r = (rc.range >> 12) * (rec & 0x1fffu);
                if (rec & QE4_BIT) { const u32 s = rc.low32 + r; rc.carry += (s < rc.low32); rc.low32 = s; rc.range -= r; }
                else rc.range = r;
            }
            if (lane == 0) P.head = h;
            if (result) break;
        }
        if (result == 0) result = (int)rc.finish();
        if (lane == 0) { if (result < 0 && P.fail == 0) P.fail = 1; sb.result = result; }
        return;
    }

    short *cold_s = cold_all + (size_t)sid * 2 * COLD_PAD, *cold_c = cold_s + COLD_PAD;
    u32 n_cached = 0, misses = 0;
    int ctxRank0 = 0, ctxRank4 = 0, ctxRun = 0, avgRank = 0, maxRank = 7;
    u32 off = 0;                                               // ring position (record index) of the current run

    if (warp == 0) {
        // ----------------------- stream header: n as 32 raw bits, then the MTF order (qlfc.cpp:851-891) -----------------------
        const u32 n = sb.in_size;
        if (lane < 32) P.ring[lane] = (u16)(2048u | (((n >> (31 - lane)) & 1u) ? QE4_BIT : 0u));
        off = 32;
        {
            const u8 *mtf = mtf_all + sid * 256;
            u32 used8 = 0; int prev = -1;
            for (int d = 0; d < 256; ++d) {
                const int c = mtf[d];
                for (int bit = 7; bit >= 0; --bit) {
                    bool can0, can1; header_options(used8, prev, c >> (bit + 1), bit, can0, can1);
                    if (can0 && can1) { if (lane == 0) P.ring[off & (QE4_RING - 1)] = (u16)(2048u | (((c >> bit) & 1) ? QE4_BIT : 0u)); ++off; }   // <= 2048 records: fits the empty ring
                }
                if (c == prev) { maxRank = ilog2_dev((u32)(d - 1)); break; }
                prev = c; if ((u32)(c >> 3) == lane) used8 |= 1u << (c & 7);
            }
        }
        __syncwarp();
        __threadfence_block();
        if (lane == 0) { P.max_rank = (u32)maxRank; P.hdr_len = off; P.progA = off; __threadfence_block(); P.hdr_ready = 1; }
    } else {
        for (u32 spins = 0; !P.hdr_ready; ) if (++spins > (1u << 27)) { P.fail = 2; break; }
        __threadfence_block();
        maxRank = (int)P.max_rank; off = P.hdr_len;
        if (lane == 0) P.progB = off;
    }

    bool stop = false;
    for (u32 t0 = rb; t0 < re && !stop; t0 += 32) {
        const u32 cnt = min(32u, re - t0);
        u32 my_sym = 0, my_rank = 0, my_len = 0;               // lane j prefetches run t0 + j
        if (lane < cnt) { my_sym = run_sym[t0 + lane]; my_rank = run_rank[t0 + lane]; my_len = run_pos[t0 + lane + 1] - run_pos[t0 + lane]; }
        for (u32 j = 0; j < cnt; ++j) {
            const u32 c = __shfl_sync(0xffffffffu, my_sym, j);
            const u32 rank = __shfl_sync(0xffffffffu, my_rank, j);
            const u32 run = __shfl_sync(0xffffffffu, my_len, j);
            const bool esc = avgRank >= 32;
            const u32 er = (u32)ilog2_dev(rank), eu = (u32)ilog2_dev(run);
            const int rank0 = (int)rank - 1;
            // decisions of this run: [rank part nA][run part nB]
            const u32 nE = (!esc && rank != 1) ? (er - 1) + ((int)er < maxRank ? 1u : 0u) : 0u;
            const u32 nM = esc ? (u32)maxRank + 1u : (rank != 1 ? er : 0u);
            const u32 nA = (esc ? 0u : 1u) + nE + nM;
            const u32 nB = 1u + (run != 1 ? 2u * eu : 0u);

            if (warp == 0) {
                const u32 st1 = S.rank_state[(ctxRun << 11) | (ctxRank4 << 3) | S.rankHist[c]];
                S.rankHist[c] = (u8)((esc || rank != 1) ? er : 0);          // all lanes store the same value
                if (!enc4_wait_room(P, off, nA)) { stop = true; break; }
                // lane d: first bit | exponent index k | mantissa level l
                const u32 d = lane;
                const bool act = d < nA;
                const u32 dT = esc ? 0xffffffffu : 0u;                    // escape mode has no first-bit decision
                const bool isT = d == dT, isE = !esc && d >= 1 && d <= nE;
                const u32 k = d - 1, l = esc ? d : d - 1 - nE;
                const u32 e_m = esc ? (u32)maxRank + 1u : er, v = esc ? (rank | (1u << e_m)) : rank;
                const u32 bp = e_m - 1 - (l < e_m ? l : 0), node = v >> (bp + 1);
                const int K = isT ? K_RANK_T : isE ? K_RANK_E : (esc ? K_RANK_P : K_RANK_M);
                const u32 bit = isT ? (rank != 1) : isE ? (k + 1 < er) : ((v >> bp) & 1u);
                const u32 bank = esc ? 8u : er;
                const bool cached = !isT && !isE && (esc || er > M_MAXE);
                const u32 pos = isT ? 0u : isE ? k : node;
                const u32 is = (isT ? R_RT_STATE + st1 : isE ? R_RE_STATE + st1 * 8 : R_RM_STATE + st1 * M_ROW + (1u << er) - 2u) + pos;
                const u32 ic = (isT ? R_RT_CHAR + c : isE ? R_RE_CHAR + c * 8 : R_RM_CHAR + c * M_ROW + (1u << er) - 2u) + pos;
                const u32 ig = (isT ? R_RT_SHARED : isE ? R_RE_SHARED : R_WIDE_SHARED + bank * 256) + pos;
                u32 rec = enc4_chunk(S, P, act, K, bit, is, ic, ig, cached, wide_idx(bank, st1, node), wide_idx(bank, c, node), cold_s, cold_c, lane, n_cached, misses);
                if (d == 0) rec |= QE4_RUN;
                if (act) P.ring[(off + d) & (QE4_RING - 1)] = (u16)rec;
            } else {
                const int rh = S.runHist[c];
                const u32 st2 = S.run_state[(ctxRank0 << 10) | (ctxRun << 6) | ((rank0 < 7 ? rank0 : 7) << 3) | (rh < 7 ? rh : 7)];
                S.runHist[c] = (u8)(run == 1 ? (rh + 2) >> 2 : (rh + 3 * (int)eu + 3) >> 2);
                if (!enc4_wait_room(P, off, nA + nB)) { stop = true; break; }
                for (u32 base = 0; base < nB; base += 32) {              // nB <= 61; one chunk unless the run is >= 64 Ki long
                    const u32 d = base + lane;
                    const bool act = d < nB;
                    const bool isT = d == 0, isE = d >= 1 && d <= eu;
                    const u32 k = d - 1, l = d - 1 - eu;
                    const u32 bp = eu - 1 - (l < eu ? l : 0);
                    const bool tree = eu <= M_MAXE;
                    const u32 node = tree ? (run >> (bp + 1)) : 1u + l;
                    const int K = isT ? K_RUN_T : isE ? K_RUN_E : K_RUN_M;
                    const u32 bit = isT ? (run != 1) : isE ? (k + 1 < eu) : ((run >> bp) & 1u);
                    const bool cached = isE ? (k >= UE_RES) : (!isT && !tree);
                    const u32 pos = isT ? 0u : isE ? k : node;
                    const u32 is = (isT ? R_UT_STATE + st2 : isE ? R_UE_STATE + st2 * UE_RES : R_UM_STATE + st2 * M_ROW + (1u << eu) - 2u) + pos;
                    const u32 ic = (isT ? R_UT_CHAR + c : isE ? R_UE_CHAR + c * UE_RES : R_UM_CHAR + c * M_ROW + (1u << eu) - 2u) + pos;
                    const u32 ig = (isT ? R_UT_SHARED : isE ? R_UE_SHARED : R_NARROW_SHARED + eu * 32) + pos;
                    const u32 cs = isE ? ue_idx(st2, k) : narrow_idx(eu, st2, node), cc = isE ? ue_idx(c, k) : narrow_idx(eu, c, node);
                    const u32 rec = enc4_chunk(S, P, act, K, bit, is, ic, ig, cached, cs, cc, cold_s, cold_c, lane, n_cached, misses);
                    if (act) P.ring[(off + nA + d) & (QE4_RING - 1)] = (u16)rec;
                }
            }
            off += nA + nB;
            avgRank = (avgRank * 124 + (int)rank * 4) >> 7;
            ctxRank0 = ((ctxRank0 << 1) | (rank0 == 0)) & 0x7;
            ctxRank4 = ((ctxRank4 << 2) | (rank0 < 3 ? rank0 : 3)) & 0xff;
            ctxRun   = ((ctxRun << 1) | (run < 3)) & 0xf;
            if ((j & 3) == 3 || j + 1 == cnt) {                        // publish progress every 4 runs
                __syncwarp();
                __threadfence_block();
                if (lane == 0) { if (warp == 0) P.progA = off; else P.progB = off; }
                if (P.fail) { stop = true; break; }
            }
        }
    }
    __syncwarp();
    __threadfence_block();
    if (lane == 0) { if (warp == 0) { P.progA = off; P.doneA = 1; } else { P.progB = off; P.doneB = 1; } }
    misses = __reduce_add_sync(0xffffffffu, misses);
    if (lane == 0) { if (warp == 0) sb.stat_cached = n_cached; else sb.stat_miss = misses; }
}
