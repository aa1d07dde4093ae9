// libbsc_b200/csrc/qlfc_encoder.cuh -- QLFC stage 2 ENCODER as a five-warp pipeline.
// Included by qlfc.cu after qlfc_coder.cuh (uses CoderSmem, the counter-file layout, Rc2Enc).
//
// Facts this design rests on (qlfc.cpp:829-1129):
//   * every model context is a function of the INPUT (symbols, ranks, run lengths) only -- never of
//     the coder state -- so probabilities can be produced ahead of the range coder;
//   * within one run every binary decision touches a different counter, so all decisions of a run
//     can be evaluated simultaneously, one lane per decision;
//   * the four decision groups of a run -- rank first-bit + exponent, rank mantissa (or escape),
//     run-length first-bit + exponent, run-length mantissa -- use disjoint counter arrays, so four
//     different warps can evaluate them without any ordering between them;
//   * the only truly serial recurrence is the range coder (range/low, rangecoder.h:83-177).
//
//   warp 0..3  "model" : per run, lane d evaluates decision d of this warp's group
//   warp 4     "coder" : consumes the (bit, p) records in stream order from a shared-memory ring
//
// All model warps derive the ring position of every run from the same arithmetic (the number of
// decisions of a run follows from rank, run length, maxRank and the escape flag), so they never
// talk to each other; the coder consumes up to the minimum of the four progress counters.
//
// The per-run CONTEXTS are computed 32 runs at a time, one lane per run ("vector prologue"):
// sliding-window contexts come from warp ballots, the per-symbol histories from match_any chains,
// avgRank from a 32-step serial integer loop, ring offsets from a warp scan, and the state-table
// look-ups are one vector shared-memory load.  History of the critical warp's instruction count per
// run (ncu, profiles/): one warp doing everything 530 -> 2 model warps 205 -> vector prologue 170
// -> 4 model warps (this file).
#pragma once

#define QE_RING 2048
#define QE_BIT  0x2000u                                     // record: bits 0..12 p, bit 13 the coded bit, bit 14 run start
#define QE_RUN  0x4000u
#define QE_MODELS 4

struct EncPipe {
    u16 ring[QE_RING];
    volatile u32 prog[QE_MODELS], done[QE_MODELS];          // records completed by each model warp / its end flag
    volatile u32 head;                                      // records consumed by the coder
    volatile u32 hdr_len, hdr_ready, max_rank;              // published by warp 0 after the stream header
    volatile u32 fail;
    int prm[7][2][12];                                      // [class][bit] = w0 w1 w2 | Ms Ks | Mc Kc | Mg Kg
    u8  hist2[2][256];                                      // private history copies of warps 1 and 3 (warps 0, 2 use CoderSmem's)
};

// parameters of one decision class for one outcome, in the (p*M + K) >> 12 form of the counter moves
__device__ __forceinline__ void enc_fill_params(EncPipe &P, u32 lane)
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

// Evaluate one decision: counters at s16 indices (is, ic, ig); returns the record.  Rare counters
// (cached == true) go through the direct-mapped caches; the caller guarantees that no two lanes of
// the same call use the same cache slot (else it serialises the lanes).
__device__ __forceinline__ u32 enc_decide(CoderSmem &S, EncPipe &P, int K, u32 bit, u32 is, u32 ic, u32 ig, bool cached, u32 cs, u32 cc,
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
    return (u32)p | (bit ? QE_BIT : 0u);
}

// run the per-lane decisions of one chunk; lanes with cached counters that collide on a cache slot are serialised
__device__ __forceinline__ u32 enc_chunk(CoderSmem &S, EncPipe &P, bool act, int K, u32 bit, u32 is, u32 ic, u32 ig, bool cached, u32 cs, u32 cc,
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
    if (!any_clash) { if (act) rec = enc_decide(S, P, K, bit, is, ic, ig, cached, cs, cc, cold_s, cold_c, misses); }
    else for (u32 turn = 0; turn < 32; ++turn) {
        if (act && lane == turn) rec = enc_decide(S, P, K, bit, is, ic, ig, cached, cs, cc, cold_s, cold_c, misses);
        __syncwarp();
    }
    __syncwarp();
    return rec;
}

// bits [32-lane, ...) of (prev:cur): the flags of the runs before this lane's run, most recent in bit 0
__device__ __forceinline__ u32 enc_window(u32 prev_rev, u32 cur_rev, u32 lane) { return (u32)((((u64)prev_rev << 32) | cur_rev) >> (32u - lane)); }

__global__ void __launch_bounds__(160, 1) q_encode5(const u32 *__restrict__ run_pos, const u8 *__restrict__ run_sym, const u8 *__restrict__ run_rank,
                                                    SubBlock *__restrict__ sbs, const u8 *__restrict__ mtf_all, short *__restrict__ cold_all,
                                                    const QTables *__restrict__ tables, u8 *__restrict__ out_all, const u32 *__restrict__ sb_list)
{
    extern __shared__ __align__(16) u8 q_smem_raw[];
    CoderSmem &S = *reinterpret_cast<CoderSmem *>(q_smem_raw);
    EncPipe &P = *reinterpret_cast<EncPipe *>(q_smem_raw + ((sizeof(CoderSmem) + 15) & ~(size_t)15));
    const u32 warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const u32 sid = sb_list ? sb_list[blockIdx.x] : blockIdx.x;
    SubBlock &sb = sbs[sid];

    if (warp == 0) coder_smem_init(S, tables);
    if (warp == 1) {
        enc_fill_params(P, lane);
        for (int i = lane; i < 512; i += 32) (&P.hist2[0][0])[i] = 0;
        if (lane < QE_MODELS) { P.prog[lane] = 0; P.done[lane] = 0; }
        if (lane == 0) { P.head = 0; P.hdr_len = 0; P.hdr_ready = 0; P.max_rank = 7; P.fail = 0; }
    }
    __syncthreads();

    const u32 rb = sb.run_begin, re = sb.run_end;

    if (warp == QE_MODELS) {
        // ---------------------------------------- coder ----------------------------------------
        Rc2Enc rc; rc.init(out_all + sb.out_off);
        const long long eob = (long long)sb.out_cap - 16;
        u32 h = 0; int result = 0; bool eob_hit = eob <= 0;
        for (;;) {
            u32 limit, spins = 0;
            for (;;) {
                limit = min(min(P.prog[0], P.prog[1]), min(P.prog[2], P.prog[3]));
                if (limit != h) break;
                if (P.done[0] && P.done[1] && P.done[2] && P.done[3] && min(min(P.prog[0], P.prog[1]), min(P.prog[2], P.prog[3])) == h) break;
                if (P.fail == 2 || ++spins > (1u << 27)) { result = LIBBSC_GPU_ERROR; break; }
            }
            if (result || limit == h) break;
            __threadfence_block();
            while (h != limit) {
                if (!eob_hit && (h & 3u) == 0 && limit - h >= 4u) {
                    // fast path: four records per 64-bit shared-memory load, no end-of-buffer test.  The
                    // test of qlfc.cpp:898-901 can only start to matter after a renormalisation moved the
                    // write position past the limit; from then on the checked path below takes over.
                    const uint2 q = *reinterpret_cast<const uint2 *>(&P.ring[h & (QE_RING - 1)]);
                    const u32 recs[4] = {q.x & 0xffffu, q.x >> 16, q.y & 0xffffu, q.y >> 16};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const u32 rec = recs[k]; ++h;
                        if (rc.range < 0x10000u) { rc.shift(); rc.range <<= 16; eob_hit = (long long)rc.pos >= eob; }
                        rc.step((rec >> 13) & 1u, rec & 0x1fffu);
                        if (eob_hit) break;
                    }
                    continue;
                }
                const u32 rec = P.ring[h & (QE_RING - 1)]; ++h;
                if (eob_hit && (rec & QE_RUN)) { result = LIBBSC_NOT_COMPRESSIBLE; break; }   // qlfc.cpp:898-901
                if (rc.range < 0x10000u) { rc.shift(); rc.range <<= 16; eob_hit = (long long)rc.pos >= eob; }
                rc.step((rec >> 13) & 1u, rec & 0x1fffu);
            }
            if (lane == 0) P.head = h;
            if (result) break;
        }
        if (result == 0) result = (int)rc.finish();
        if (lane == 0) { if (result < 0 && P.fail == 0) P.fail = 1; sb.result = result; }
        return;
    }

    // ------------------------------------------ model warps ------------------------------------------
    const bool rank_side = warp < 2;                          // warps 0,1: rank decisions; warps 2,3: run-length decisions
    u8 *my_hist = warp == 0 ? S.rankHist : warp == 2 ? S.runHist : warp == 1 ? P.hist2[0] : P.hist2[1];
    short *cold_s = cold_all + (size_t)sid * 2 * COLD_PAD, *cold_c = cold_s + COLD_PAD;
    u32 n_cached = 0, misses = 0;
    u32 maxRank = 7;
    u32 base_off = 0;                                          // ring position (record index) of the current batch of runs

    if (warp == 0) {
        // ----------------------- stream header: n as 32 raw bits, then the MTF order (qlfc.cpp:851-891) -----------------------
        const u32 n = sb.in_size;
        P.ring[lane] = (u16)(2048u | (((n >> (31 - lane)) & 1u) ? QE_BIT : 0u));
        u32 off = 32;
        {
            const u8 *mtf = mtf_all + sid * 256;
            u32 used8 = 0; int prev = -1;
            for (int d = 0; d < 256; ++d) {
                const int c = mtf[d];
                for (int bit = 7; bit >= 0; --bit) {
                    bool can0, can1; header_options(used8, prev, c >> (bit + 1), bit, can0, can1);
                    if (can0 && can1) { if (lane == 0) P.ring[off & (QE_RING - 1)] = (u16)(2048u | (((c >> bit) & 1) ? QE_BIT : 0u)); ++off; }   // <= 2048 records: fits the empty ring
                }
                if (c == prev) { maxRank = (u32)ilog2_dev((u32)(d - 1)); break; }
                prev = c; if ((u32)(c >> 3) == lane) used8 |= 1u << (c & 7);
            }
        }
        base_off = off;
        __syncwarp();
        __threadfence_block();
        if (lane == 0) { P.max_rank = maxRank; P.hdr_len = off; P.prog[0] = off; __threadfence_block(); P.hdr_ready = 1; }
    } else {
        for (u32 spins = 0; !P.hdr_ready; ) if (++spins > (1u << 27)) { P.fail = 2; break; }
        __threadfence_block();
        maxRank = P.max_rank; base_off = P.hdr_len;
        if (lane == 0) P.prog[warp] = base_off;
    }

    // ---- state carried from batch to batch ----
    u32 avg = 0;                                               // avgRank after the last run of the previous batch
    u32 pf0 = 0, plo = 0, phi = 0, prn = 0;                    // bit-reversed flag ballots of the previous batch
    u32 head_seen = 0;
    bool stop = false;

    for (u32 t0 = rb; t0 < re && !stop; t0 += 32) {
        // ================= vector prologue: lane j <-> run t0 + j =================
        const u32 cnt = min(32u, re - t0);
        const bool live = lane < cnt;
        u32 sym = 256u + lane, rank = 1, len = 1;              // dead lanes: unique pseudo-symbols, no decisions
        if (live) { sym = run_sym[t0 + lane]; rank = run_rank[t0 + lane]; len = run_pos[t0 + lane + 1] - run_pos[t0 + lane]; }
        const u32 er = (u32)ilog2_dev(rank), eu = (u32)ilog2_dev(len), rank0 = rank - 1;
        u32 my_avg = 0;                                        // avgRank seen by this lane's run (qlfc.cpp:1048: serial integer recurrence)
        {
            u32 a = avg;
#pragma unroll 8
            for (u32 j = 0; j < 32; ++j) {
                const u32 r = __shfl_sync(0xffffffffu, rank, j);
                if (lane == j) my_avg = a;
                if (j < cnt) a = (a * 124u + r * 4u) >> 7;
            }
            avg = a;
        }
        const bool esc = my_avg >= 32u;
        const u32 nE = (!esc && rank != 1) ? (er - 1) + (er < maxRank ? 1u : 0u) : 0u;
        const u32 nM = esc ? maxRank + 1u : (rank != 1 ? er : 0u);
        const u32 nA = live ? (esc ? 0u : 1u) + nE + nM : 0u;
        const u32 nB = live ? 1u + (len != 1 ? 2u * eu : 0u) : 0u;
        u32 incl = nA + nB;                                    // ring offsets: exclusive warp scan of the decision counts
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const u32 t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (u32)o) incl += t; }
        const u32 my_off = incl - (nA + nB), batch_total = __shfl_sync(0xffffffffu, incl, 31);
        // sliding-window contexts (qlfc.cpp:1123-1125) from ballots; most recent run in the low bits
        const u32 q3 = rank0 < 3 ? rank0 : 3;
        const u32 bf0 = __brev(__ballot_sync(0xffffffffu, live && rank0 == 0));
        const u32 blo = __brev(__ballot_sync(0xffffffffu, live && (q3 & 1u))), bhi = __brev(__ballot_sync(0xffffffffu, live && (q3 & 2u)));
        const u32 brn = __brev(__ballot_sync(0xffffffffu, live && len < 3));
        const u32 ctxRank0 = enc_window(pf0, bf0, lane) & 7u, ctxRun = enc_window(prn, brn, lane) & 15u;
        const u32 wl = enc_window(plo, blo, lane) & 15u, wh = enc_window(phi, bhi, lane) & 15u;
        const u32 ctxRank4 = (wl & 1u) | ((wh & 1u) << 1) | ((wl & 2u) << 1) | ((wh & 2u) << 2) | ((wl & 4u) << 2) | ((wh & 4u) << 3) | ((wl & 8u) << 3) | ((wh & 8u) << 4);
        pf0 = bf0; plo = blo; phi = bhi; prn = brn;
        // per-symbol histories: value left by the previous run of the same symbol
        const u32 same = __match_any_sync(0xffffffffu, sym);
        const u32 below = same & lanemask_lt();
        const u32 prevl = below ? 31u - (u32)__clz(below) : lane;
        const bool last_of_sym = live && (same >> lane) == 1u;   // no later run of this symbol in the batch
        u32 my_st;
        if (rank_side) {
            const u32 er_prev = __shfl_sync(0xffffffffu, er, prevl);
            const u32 rh = below ? er_prev : (u32)my_hist[sym & 255u];             // rankHistory = exponent of the previous rank (0 for rank 1)
            my_st = S.rank_state[(ctxRun << 11) | (ctxRank4 << 3) | rh];
            __syncwarp();
            if (last_of_sym) my_hist[sym] = (u8)er;
        } else {
            const u32 occ = __popc(below), maxocc = __reduce_max_sync(0xffffffffu, live ? occ : 0u);
            u32 vin = my_hist[sym & 255u], vout = 0;
            for (u32 round = 0; round <= maxocc; ++round) {       // resolve same-symbol chains in occurrence order
                const u32 t = __shfl_sync(0xffffffffu, vout, prevl);
                if (occ == round) { if (round) vin = t; vout = (len == 1) ? (vin + 2u) >> 2 : (vin + 3u * eu + 3u) >> 2; }
            }
            my_st = S.run_state[(ctxRank0 << 10) | (ctxRun << 6) | ((rank0 < 7 ? rank0 : 7u) << 3) | (vin < 7 ? vin : 7u)];
            __syncwarp();
            if (last_of_sym) my_hist[sym] = (u8)vout;
        }
        const u32 pack = er | (eu << 3) | ((esc ? 1u : 0u) << 8) | (nE << 9) | (nM << 12) | (nA << 16) | (nB << 20);
        __syncwarp();

        // ================= per run: one lane per decision of this warp's group =================
        for (u32 j = 0; j < cnt; ++j) {
            const u32 c = __shfl_sync(0xffffffffu, sym, j);
            const u32 val = __shfl_sync(0xffffffffu, rank_side ? rank : len, j);      // the coded value: rank or run length
            const u32 pk = __shfl_sync(0xffffffffu, pack, j), st = __shfl_sync(0xffffffffu, my_st, j);
            const u32 off = base_off + __shfl_sync(0xffffffffu, my_off, j);
            const u32 r_er = pk & 7u, r_eu = (pk >> 3) & 31u, r_nE = (pk >> 9) & 7u, r_nM = (pk >> 12) & 15u, r_nA = (pk >> 16) & 15u, r_nB = (pk >> 20) & 63u;
            const bool r_esc = (pk >> 8) & 1u;
            {   // room in the ring for this run's records (re-read the consumer's head only when needed)
                const u32 end = off + r_nA + r_nB;
                for (u32 spins = 0; (int)(end - head_seen) > QE_RING; ) {
                    head_seen = P.head;
                    if (P.fail || ++spins > (1u << 26)) { if (!P.fail) P.fail = 2; stop = true; break; }
                }
                if (stop) break;
            }
            if (warp == 0) {            // rank: first bit (lane 0) and unary exponent (lanes 1..nE); nothing in escape mode
                const u32 d = lane, k = d - 1;
                const bool act = !r_esc && d <= r_nE, isT = d == 0;
                const u32 bit = isT ? (val != 1) : (k + 1 < r_er);
                const u32 is = isT ? R_RT_STATE + st : R_RE_STATE + st * 8 + k;
                const u32 ic = isT ? R_RT_CHAR + c : R_RE_CHAR + c * 8 + k;
                const u32 ig = isT ? R_RT_SHARED : R_RE_SHARED + k;
                u32 rec = enc_chunk(S, P, act, isT ? K_RANK_T : K_RANK_E, bit, is, ic, ig, false, 0, 0, cold_s, cold_c, lane, n_cached, misses);
                if (isT) rec |= QE_RUN;
                if (act) P.ring[(off + d) & (QE_RING - 1)] = (u16)rec;
            } else if (warp == 1) {     // rank: mantissa tree of depth e (or the escape tree of depth maxRank+1)
                const u32 l = lane;
                const bool act = l < r_nM;
                const u32 e_m = r_esc ? maxRank + 1u : r_er, v = r_esc ? (val | (1u << e_m)) : val;
                const u32 bp = e_m - 1 - (l < e_m ? l : 0), node = v >> (bp + 1), bit = (v >> bp) & 1u;
                const u32 bank = r_esc ? 8u : r_er;
                const bool cached = r_esc || r_er > M_MAXE;
                const u32 rowoff = (1u << r_er) - 2u + node;
                u32 rec = enc_chunk(S, P, act, r_esc ? K_RANK_P : K_RANK_M, bit, R_RM_STATE + st * M_ROW + rowoff, R_RM_CHAR + c * M_ROW + rowoff,
                                    R_WIDE_SHARED + bank * 256 + node, cached, wide_idx(bank, st, node), wide_idx(bank, c, node), cold_s, cold_c, lane, n_cached, misses);
                if (r_esc && l == 0) rec |= QE_RUN;
                if (act) P.ring[(off + (r_esc ? 0u : 1u + r_nE) + l) & (QE_RING - 1)] = (u16)rec;
            } else if (warp == 2) {     // run length: first bit (lane 0) and unary exponent (lanes 1..eu)
                const u32 d = lane, k = d - 1;
                const u32 nTE = val != 1 ? 1u + r_eu : 1u;
                const bool act = d < nTE, isT = d == 0;
                const u32 bit = isT ? (val != 1) : (k + 1 < r_eu);
                const bool cached = !isT && k >= UE_RES;
                const u32 is = isT ? R_UT_STATE + st : R_UE_STATE + st * UE_RES + k;
                const u32 ic = isT ? R_UT_CHAR + c : R_UE_CHAR + c * UE_RES + k;
                const u32 ig = isT ? R_UT_SHARED : R_UE_SHARED + k;
                const u32 rec = enc_chunk(S, P, act, isT ? K_RUN_T : K_RUN_E, bit, is, ic, ig, cached, ue_idx(st, k), ue_idx(c, k), cold_s, cold_c, lane, n_cached, misses);
                if (act) P.ring[(off + r_nA + d) & (QE_RING - 1)] = (u16)rec;
            } else {                    // run length: mantissa (tree for exponents <= 5, linear contexts above; qlfc.cpp:1119)
                const u32 l = lane;
                const bool act = val != 1 && l < r_eu;
                const u32 bp = r_eu - 1 - (l < r_eu ? l : 0), bit = (val >> bp) & 1u;
                const bool tree = r_eu <= M_MAXE;
                const u32 node = tree ? (val >> (bp + 1)) : 1u + l;
                const u32 rowoff = (1u << r_eu) - 2u + node;
                const u32 rec = enc_chunk(S, P, act, K_RUN_M, bit, R_UM_STATE + st * M_ROW + rowoff, R_UM_CHAR + c * M_ROW + rowoff,
                                          R_NARROW_SHARED + r_eu * 32 + node, !tree, narrow_idx(r_eu, st, node), narrow_idx(r_eu, c, node), cold_s, cold_c, lane, n_cached, misses);
                if (act) P.ring[(off + r_nA + 1u + r_eu + l) & (QE_RING - 1)] = (u16)rec;
            }
            if ((j & 3) == 3 || j + 1 == cnt) {                        // publish progress every 4 runs
                __syncwarp();
                __threadfence_block();
                if (lane == 0) P.prog[warp] = off + r_nA + r_nB;
                if (P.fail) { stop = true; break; }
            }
        }
        base_off += batch_total;
    }
    __syncwarp();
    __threadfence_block();
    if (lane == 0) { P.prog[warp] = base_off; P.done[warp] = 1; }
    misses = __reduce_add_sync(0xffffffffu, misses);
    if (lane == 0) { atomicAdd(&sb.stat_cached, n_cached); atomicAdd(&sb.stat_miss, misses); }
}
