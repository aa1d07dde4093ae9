// libbsc_b200/csrc/qlfc_decoder3.cuh -- QLFC stage 2 DECODER, third generation (qlfc.cpp:1672-1927).
// Included by qlfc.cu after qlfc_coder.cuh; ALSO compiled for the host by tools/qdec3_host.cpp (QD3_HOST), which
// runs the very same source with the 32 lanes emulated one after the other, so that the lane logic below is
// checked bit-for-bit against the oracle on the CPU (tests/test_qdec3_host.py) before it ever reaches a GPU.
//
// One warp per stream, as before: the decoder is a serial recurrence (the context of every decision depends
// on the previous decisions), so per-decision latency is everything.  q_decode2 walks the decisions one by one
// and pays, for every decision, address arithmetic -> three shared-memory counter loads -> three-term mix ->
// range-coder step -> three counter moves: ~230 cycles per decision measured (profiles/r1g).
//
// What this version does differently -- SPECULATIVE EVALUATION across the lanes:
//   * the rank of a run is coded as first-bit / unary exponent / mantissa-tree decisions whose contexts depend
//     only on (state, symbol) -- both known when the run starts -- and on the path taken.  All 70 candidate
//     contexts (1 first-bit, 7 exponent, 62 tree nodes of the compact rows for exponents 1..5) are evaluated
//     at once, two or three per lane: counter loads, mix, and the probability goes to a small shared array;
//   * eight more lanes evaluate the run-length first-bit for the eight possible rank classes (its state
//     depends on min(rank-1, 7));
//   * the serial part that remains is the range-coder walk over ready probabilities: first-bit and exponent
//     probabilities arrive in ONE 128-bit load, the mantissa walk loads the (left, right) children pair of the
//     node it is about to decide while that decision is being resolved, so no load sits between two decisions;
//   * the counter moves are done afterwards by the lanes that own the counters on the decoded path, off the
//     critical path of the next decision;
//   * the range-coder step is branch-free: the next 16-bit unit is always preloaded, renormalisation is a
//     select; the input window is checked once per run (a run consumes at most 152 bytes), not per decision.
// Everything rare (rank exponents 6-7, the escape mode, run lengths > 1) keeps the serial decision-by-decision
// code on the shared counter file / the write-back caches; those paths touch counters no lane owns.
#pragma once

#include <cstddef>

#ifdef QD3_HOST
#define QD3_FN static inline
#define QD3_FN_MEMBER
#define QD3_LANES for (u32 lane = 0; lane < 32; ++lane)
#define QD3_L(x) x[lane]
#define QD3_SYNC() do { } while (0)
#define QD3_PARAM(k, i) ((int)bscb_static_params[k][i])
#else
#define QD3_PARAM(k, i) ((int)c_params[k][i])     // lane-dependent class: a run-time look-up (once per stream)
#define QD3_FN __device__ __forceinline__
#define QD3_FN_MEMBER __device__ __forceinline__
#define QD3_LANES
#define QD3_L(x) x
#define QD3_SYNC() __syncwarp()
#endif

struct Dec3Smem {
    CoderSmem cs;
    alignas(16) u16 px[16];      // [0] rank first-bit, [1..7] rank exponent k = 0..6, [8..15] run first-bit for rank class q = 0..7
    alignas(16) u8  st2[16];     // [q] run state for rank class q
    alignas(16) u16 pm[128];     // rank mantissa probabilities at the compact row offsets 1..61 (+ slack for the pair loads)
    alignas(16) u8  win[272];    // staged window of the input stream
};
constexpr u32 O3_PX = (u32)offsetof(Dec3Smem, px), O3_ST2 = (u32)offsetof(Dec3Smem, st2), O3_PM = (u32)offsetof(Dec3Smem, pm), O3_WIN = (u32)offsetof(Dec3Smem, win);
constexpr u32 O3_RANK_STATE = (u32)offsetof(CoderSmem, rank_state), O3_RUN_STATE = (u32)offsetof(CoderSmem, run_state);
constexpr u32 O3_TAG_STATE = (u32)offsetof(CoderSmem, tag_state), O3_TAG_CHAR = (u32)offsetof(CoderSmem, tag_char);
constexpr u32 O3_RANK_HIST = (u32)offsetof(CoderSmem, rankHist), O3_RUN_HIST = (u32)offsetof(CoderSmem, runHist);
constexpr u32 O3_MTF = (u32)offsetof(CoderSmem, mtf), O3_S16 = (u32)offsetof(CoderSmem, s16);
constexpr u32 QD3_WIN_BYTES = 272, QD3_RUN_ROOM = 100;      // refill at a run start when more than this is consumed: 100 + 152 + 2 < 272

struct U4 { u32 x, y, z, w; };
struct U2 { u32 x, y; };

// shared-memory accessors on an explicit base (byte offsets inside Dec3Smem); see qlfc_decoder.cuh for why
struct SM3 {
#ifdef QD3_HOST
    u8 *b;
    u32 ld8(u32 off) const { return b[off]; }
    u32 ld16(u32 off) const { u16 v; memcpy(&v, b + off, 2); return v; }
    u32 ld32(u32 off) const { u32 v; memcpy(&v, b + off, 4); return v; }
    U2  ld64(u32 off) const { U2 v; memcpy(&v, b + off, 8); return v; }
    U4  ld128(u32 off) const { U4 v; memcpy(&v, b + off, 16); return v; }
    void st8(u32 off, u32 v) const { b[off] = (u8)v; }
    void st16(u32 off, u32 v) const { u16 t = (u16)v; memcpy(b + off, &t, 2); }
    u32 at(u32 off) const { return off; }                              // "absolute address" of an offset (host: the offset itself)
    u32 ld16a(u32 a) const { return ld16(a); }
    void st16a(u32 a, u32 v) const { st16(a, v); }
#else
    u32 b;
    __device__ __forceinline__ u32 ld8(u32 off) const { u32 v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(b + off)); return v; }
    __device__ __forceinline__ u32 ld16(u32 off) const { u32 v; asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(b + off)); return v; }
    __device__ __forceinline__ u32 ld32(u32 off) const { u32 v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(b + off)); return v; }
    __device__ __forceinline__ U2  ld64(u32 off) const { U2 v; asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(b + off)); return v; }
    __device__ __forceinline__ U4  ld128(u32 off) const { U4 v; asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(b + off)); return v; }
    __device__ __forceinline__ void st8(u32 off, u32 v) const { asm volatile("st.shared.u8 [%0], %1;" :: "r"(b + off), "r"(v) : "memory"); }
    __device__ __forceinline__ void st16(u32 off, u32 v) const { asm volatile("st.shared.u16 [%0], %1;" :: "r"(b + off), "r"(v) : "memory"); }
    __device__ __forceinline__ u32 at(u32 off) const { return b + off; }     // absolute shared-memory address of an offset
    __device__ __forceinline__ u32 ld16a(u32 a) const { u32 v; asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
    __device__ __forceinline__ void st16a(u32 a, u32 v) const { asm volatile("st.shared.u16 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
#endif
    // counters (indices in u16 units, as in qlfc_coder.cuh)
    QD3_FN_MEMBER int cnt(u32 idx) const { return (int)ld16(O3_S16 + 2u * idx); }
    QD3_FN_MEMBER void set(u32 idx, int v) const { st16(O3_S16 + 2u * idx, (u32)v); }
};

QD3_FN int qd3_ilog2(u32 v) {
#ifdef QD3_HOST
    return 31 - __builtin_clz(v | 1u);
#else
    return 31 - __clz(v | 1u);
#endif
}

// Index (into the counter file) of rare counter `idx` through the direct-mapped write-back cache (uniform).
QD3_FN u32 qd3_cache_get(const SM3 &sm, u32 val_base, u32 tags_off, short *__restrict__ cold, u32 idx, u32 &misses)
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

// ---- range decoder (rangecoder.h:203-240), branch-free step ---------------------------------------------------
struct Rc3 {
    const u8 *in; u32 limit;
    u32 code, range;
    u32 nx;                          // the 16-bit unit at `pos`, already loaded
    u32 pos, wbase;                  // next unread unit (byte offset in the stream, always even); window = [wbase, wbase + 272)
};

// one decision with P(bit = 0) = p / 4096
QD3_FN u32 qd3_step(const SM3 &sm, Rc3 &rc, u32 p)
{
    const bool need = rc.range < 0x10000u;
    rc.code = need ? (rc.code << 16) | rc.nx : rc.code;
    rc.range = need ? rc.range << 16 : rc.range;
    rc.pos += need ? 2u : 0u;
    rc.nx = sm.ld16(O3_WIN + (rc.pos - rc.wbase));
    const u32 r = (rc.range >> 12) * p;
    const bool bit = rc.code >= r;
    rc.code -= bit ? r : 0u;
    rc.range = bit ? rc.range - r : r;
    return bit ? 1u : 0u;
}

// one serial decision against three counters of the shared counter file (the rare paths)
template <int K> QD3_FN u32 qd3_dec3(const SM3 &sm, Rc3 &rc, u32 is, u32 ic, u32 ig)
{
    const int s = sm.cnt(is), c = sm.cnt(ic), g = sm.cnt(ig);
    const u32 b = qd3_step(sm, rc, (u32)q_mix<K>(s, c, g));
    sm.set(is, b ? q_down<K, 0>(s) : q_up<K, 0>(s));
    sm.set(ic, b ? q_down<K, 1>(c) : q_up<K, 1>(c));
    sm.set(ig, b ? q_down<K, 2>(g) : q_up<K, 2>(g));
    return b;
}

// ---- per-lane registers ------------------------------------------------------------------------------------------
struct Qd3Lane {
    // two mantissa-tree slots: compact row offsets lane and lane + 32
    u32 okA, eA, jA, dA, gxA;        // is a real node; its exponent (level), node number, depth of the node, shared-counter index
    u32 okB, eB, jB, dB, gxB;
    // slot X: lane 0 rank first-bit, 1..7 rank exponent k = lane - 1, 8..15 run first-bit for rank class q = lane - 8 (16..31 mirror 0..15, never store)
    u32 xS, xC, xG, xMul;            // state index = xS + (state) * xMul, symbol index = xC + c * xMul, shared index = xG
    int xw0, xw1, xw2;               // mix weights (symbol, state, shared)
    int xMs0, xKs0, xMs1, xKs1, xMc0, xKc0, xMc1, xKc1, xMg0, xKg0, xMg1, xKg1;   // counter moves (state/symbol/shared, bit 0/1): v' = (v * M + K) >> 12
    // values loaded by the evaluation phase, used again by the update phase
    int sA, cA, gA, sB, cB, gB, sX, cX, gX;
    u32 iSX, iCX;
    u32 used8, tmp;
    u32 mtfv;                        // qlfc_decoder6.cuh: lane l holds position l of the MTF list
};

QD3_FN void qd3_lane_init(Qd3Lane &r, u32 lane)
{
#define QD3_SLOT(o_, OK, E, J, D, GX) { const u32 o = (o_); const u32 e = (u32)qd3_ilog2(o + 2u), j = o + 2u - (1u << e); \
        const u32 ok = (j >= 1u && e >= 1u && e <= M_MAXE && o < M_ROW) ? 1u : 0u; \
        r.OK = ok; r.E = e; r.J = j; r.D = (u32)qd3_ilog2(j); r.GX = R_WIDE_SHARED + (ok ? e * 256u + j : 0u); }
    QD3_SLOT(lane, okA, eA, jA, dA, gxA) QD3_SLOT(lane + 32u, okB, eB, jB, dB, gxB)
#undef QD3_SLOT
    const u32 l = lane & 15u;
    int k;
    if (l == 0)     { k = K_RANK_T; r.xS = R_RT_STATE; r.xC = R_RT_CHAR; r.xG = R_RT_SHARED; r.xMul = 1; }
    else if (l < 8) { k = K_RANK_E; r.xS = R_RE_STATE + (l - 1); r.xC = R_RE_CHAR + (l - 1); r.xG = R_RE_SHARED + (l - 1); r.xMul = 8; }
    else            { k = K_RUN_T;  r.xS = R_UT_STATE; r.xC = R_UT_CHAR; r.xG = R_UT_SHARED; r.xMul = 1; }
    r.xw0 = QD3_PARAM(k, 0); r.xw1 = QD3_PARAM(k, 1); r.xw2 = QD3_PARAM(k, 2);
#define QD3_MOVE(who, M0, K0, M1, K1) { const int th0 = QD3_PARAM(k, 3 + 4 * who), ar0 = QD3_PARAM(k, 4 + 4 * who), th1 = QD3_PARAM(k, 5 + 4 * who), ar1 = QD3_PARAM(k, 6 + 4 * who); \
        r.M0 = 4096 - ar0; r.K0 = (4096 - th0) * ar0;  /* q_up */  r.M1 = 4096 - ar1; r.K1 = th1 * ar1 + 4095;  /* q_down */ }
    QD3_MOVE(0, xMs0, xKs0, xMs1, xKs1) QD3_MOVE(1, xMc0, xKc0, xMc1, xKc1) QD3_MOVE(2, xMg0, xKg0, xMg1, xKg1)
#undef QD3_MOVE
    r.sA = r.cA = r.gA = r.sB = r.cB = r.gB = r.sX = r.cX = r.gX = 0; r.iSX = r.iCX = 0; r.used8 = 0; r.tmp = 0;
}

#ifdef QD3_HOST
#define QD3_LREGS Qd3Lane lr[32]
#define QD3_LREGS_PARAM Qd3Lane (&lr)[32]
#else
#define QD3_LREGS Qd3Lane lr
#define QD3_LREGS_PARAM Qd3Lane &lr
#endif

// (re)load the input window at rc.pos: 8 bytes per lane + 16 more by lanes 0..15
#define QD3_REFILL() do { rc.wbase = rc.pos; QD3_SYNC(); \
        QD3_LANES { for (u32 k_ = 0; k_ < 8; ++k_) { const u32 w_ = lane * 8u + k_, o_ = rc.wbase + w_; sm.st8(O3_WIN + w_, o_ < rc.limit ? rc.in[o_] : 0u); } \
                    if (lane < 16u) { const u32 w_ = 256u + lane, o_ = rc.wbase + w_; sm.st8(O3_WIN + w_, o_ < rc.limit ? rc.in[o_] : 0u); } } \
        QD3_SYNC(); } while (0)

// which symbols can still appear in the MTF-order header (qlfc.cpp:857-891): lane l owns symbols 8l..8l+7
#ifdef QD3_HOST
#define QD3_HEADER_OPTIONS(prev, prefix, bit, can0, can1) do { can0 = can1 = false; \
        for (u32 lane = 0; lane < 32; ++lane) for (int k_ = 0; k_ < 8; ++k_) { const int c_ = 8 * (int)lane + k_; \
            if ((c_ == (prev) || !((lr[lane].used8 >> k_) & 1u)) && ((c_ >> ((bit) + 1)) == (prefix))) { if (c_ & (1 << (bit))) can1 = true; else can0 = true; } } } while (0)
#else
#define QD3_HEADER_OPTIONS(prev, prefix, bit, can0, can1) header_options(lr.used8, (prev), (prefix), (bit), can0, can1)
#endif

// Decodes one stream into out[0 .. n).  `sm` must point at an initialised Dec3Smem (tables copied, counters 2048, the
// rest zero).  Returns the decoded length or a (negative) libbsc error code.
// PROF (device only, BSCB200_QDEC_PROF=1): cycle counts per phase, printed for the first stream of the launch.
#ifdef QD3_HOST
#define QD3_T(k) do { } while (0)
#else
#define QD3_T(k) do { if (PROF) { const long long t_ = clock64(); prof_t[k] += t_ - prof_last; prof_last = t_; } } while (0)
#endif
// Stream prologue shared by both decoders: coder start-up, the 32-bit length, the MTF-order table.  Returns 0 or an error.
QD3_FN int qd3_prologue(const SM3 &sm, Rc3 &rc, QD3_LREGS_PARAM, const u8 *__restrict__ in, u32 in_limit, u32 out_cap, u32 &n, int &maxRank)
{
#ifndef QD3_HOST
    const u32 lane = threadIdx.x & 31u;
#endif
    QD3_LANES { qd3_lane_init(QD3_L(lr), lane); }

    rc.in = in; rc.limit = in_limit; rc.code = 0; rc.range = 0xffffffffu; rc.pos = 0; rc.wbase = 0; rc.nx = 0;
    QD3_REFILL();
    rc.code = (sm.ld16(O3_WIN + 2) << 16) | sm.ld16(O3_WIN + 4);            // rangecoder.h:203-211: three units, the first falls out of 32 bits
    rc.pos = 6; rc.nx = sm.ld16(O3_WIN + 6);
    n = 0;
    for (int b = 0; b < 32; ++b) n = (n << 1) | qd3_step(sm, rc, 2048u);
    if (n > out_cap) return LIBBSC_DATA_CORRUPT;                            // would overrun the output slice

    maxRank = 7;
    int prev = -1;
    for (int d = 0; d < 256; ++d) {
        int c = 0;
        for (int bit = 7; bit >= 0; --bit) {
            bool can0, can1; QD3_HEADER_OPTIONS(prev, c, bit, can0, can1);
            if (can0 && can1) {
                if (rc.pos - rc.wbase > 256u) QD3_REFILL();
                c = 2 * c + (int)qd3_step(sm, rc, 2048u);
            }
            else if (can1) c = 2 * c + 1;
            else if (can0) c = 2 * c;
        }
        c &= 255;
        sm.st8(O3_MTF + d, (u32)c);
        if (c == prev) { maxRank = qd3_ilog2((u32)(d - 1)); break; }
        prev = c;
        QD3_LANES { if ((u32)(c >> 3) == lane) QD3_L(lr).used8 |= 1u << (c & 7); }
    }
    QD3_SYNC();
    return 0;
}

template <bool PROF> QD3_FN int qd3_decode_stream(const SM3 &sm, const u8 *__restrict__ in, u32 in_limit, u8 *__restrict__ out, u32 out_cap,
                             short *__restrict__ cold_s, short *__restrict__ cold_c, u32 &st_cached, u32 &st_miss)
{
    QD3_LREGS;
#ifndef QD3_HOST
    const u32 lane = threadIdx.x & 31u;
#endif
    Rc3 rc; u32 n; int maxRank;
    { const int err = qd3_prologue(sm, rc, lr, in, in_limit, out_cap, n, maxRank); if (err) return err; }

    u32 ctxRank0 = 0, ctxRank4 = 0, ctxRun = 0; int avgRank = 0;
    u32 c = sm.ld8(O3_MTF), m1 = sm.ld8(O3_MTF + 1), m2 = sm.ld8(O3_MTF + 2), m3 = sm.ld8(O3_MTF + 3);
    u32 rhU = sm.ld8(O3_RUN_HIST + c);
    u32 st = sm.ld8(O3_RANK_STATE + ((ctxRun << 11) | (ctxRank4 << 3) | sm.ld8(O3_RANK_HIST + c)));

    long long prof_t[8] = {0, 0, 0, 0, 0, 0, 0, 0}, prof_last = 0; u32 prof_runs = 0;
    (void)prof_t; (void)prof_last; (void)prof_runs;
#ifndef QD3_HOST
    if (PROF) prof_last = clock64();
#endif
    for (u32 i = 0; i < n; ) {
        const bool plain = avgRank < 32;
        const u32 rhq = rhU < 7 ? rhU : 7;
        if (rc.pos - rc.wbase > QD3_RUN_ROOM) QD3_REFILL(); else QD3_SYNC();   // the sync orders the previous run's counter moves before this run's loads

        // ---- evaluation phase: every candidate decision of this run's rank code + the run-length first bit ----
        QD3_LANES {
            Qd3Lane &r = QD3_L(lr);
            const u32 q = lane & 7u;
            const u32 st2 = sm.ld8(O3_RUN_STATE + ((ctxRank0 << 10) | (ctxRun << 6) | (q << 3) | rhq));
            int pA = 0, pB = 0;
            if (plain) {
                r.sA = sm.cnt(R_RM_STATE + st * M_ROW + lane);       r.cA = sm.cnt(R_RM_CHAR + c * M_ROW + lane);       r.gA = sm.cnt(r.gxA);
                r.sB = sm.cnt(R_RM_STATE + st * M_ROW + lane + 32u); r.cB = sm.cnt(R_RM_CHAR + c * M_ROW + lane + 32u); r.gB = sm.cnt(r.gxB);
            }
            r.iSX = r.xS + ((lane & 8u) ? st2 : st) * r.xMul; r.iCX = r.xC + c * r.xMul;
            r.sX = sm.cnt(r.iSX); r.cX = sm.cnt(r.iCX); r.gX = sm.cnt(r.xG);
            if (plain) { pA = q_mix<K_RANK_M>(r.sA, r.cA, r.gA); pB = q_mix<K_RANK_M>(r.sB, r.cB, r.gB); }
            const int pX = (r.cX * r.xw0 + r.sX * r.xw1 + r.gX * r.xw2) >> 5;
            if (plain) { if (r.okA) sm.st16(O3_PM + 2u * lane, (u32)pA); if (r.okB) sm.st16(O3_PM + 2u * (lane + 32u), (u32)pB); }
            if (lane < 16u) sm.st16(O3_PX + 2u * lane, (u32)pX);
            if (lane >= 8u && lane < 16u) sm.st8(O3_ST2 + q, st2);
        }
        QD3_SYNC();
        QD3_T(0);

        // ---- the serial walk over ready probabilities (uniform) ----
        const U4 px = sm.ld128(O3_PX), pu = sm.ld128(O3_PX + 16u);
        const U2 s2 = sm.ld64(O3_ST2);
        u32 rank = 1, e = 0, nE = 0, lastE = 0, bT = 0, b;
        if (plain) {
            bT = qd3_step(sm, rc, px.x & 0xffffu);
            if (bT) {
                e = 1;
#define QD3_E(P) { if ((int)e == maxRank) break; b = qd3_step(sm, rc, (P)); ++nE; lastE = b; if (!b) break; ++e; }
                do { QD3_E(px.x >> 16) QD3_E(px.y & 0xffffu) QD3_E(px.y >> 16) QD3_E(px.z & 0xffffu) QD3_E(px.z >> 16) QD3_E(px.w & 0xffffu) } while (0);
#undef QD3_E
                if (e <= M_MAXE) {
                    const u32 mb = O3_PM + 2u * ((1u << e) - 2u);                 // pair (gap, root) of level e
                    u32 pair = sm.ld32(mb), p = pair >> 16, node = 1;
                    for (int bit = (int)e - 1; bit >= 0; --bit) {
                        pair = sm.ld32(mb + 4u * node);                           // (left, right) children of `node`
                        b = qd3_step(sm, rc, p);
                        p = b ? pair >> 16 : pair & 0xffffu;
                        node = 2u * node + b;
                    }
                    rank = node;
                } else {
                    for (int bit = (int)e - 1; bit >= 0; --bit) {
                        const u32 is = qd3_cache_get(sm, C_STATE_VAL, O3_TAG_STATE, cold_s, wide_idx(e, st, rank), st_miss);
                        const u32 ic = qd3_cache_get(sm, C_CHAR_VAL, O3_TAG_CHAR, cold_c, wide_idx(e, c, rank), st_miss);
                        st_cached += 2;
                        b = qd3_dec3<K_RANK_M>(sm, rc, is, ic, R_WIDE_SHARED + e * 256u + rank);
                        rank = 2u * rank + b;
                    }
                }
            }
            sm.st8(O3_RANK_HIST + c, e);
        } else {
            rank = 0;
            for (int node = 1, bit = maxRank; bit >= 0; --bit) {
                const u32 is = qd3_cache_get(sm, C_STATE_VAL, O3_TAG_STATE, cold_s, wide_idx(8, st, (u32)node), st_miss);
                const u32 ic = qd3_cache_get(sm, C_CHAR_VAL, O3_TAG_CHAR, cold_c, wide_idx(8, c, (u32)node), st_miss);
                st_cached += 2;
                b = qd3_dec3<K_RANK_P>(sm, rc, is, ic, R_WIDE_SHARED + 8u * 256u + (u32)node);
                node = 2 * node + (int)b; rank = 2u * rank + b;
            }
            sm.st8(O3_RANK_HIST + c, (u32)qd3_ilog2(rank));
        }
        rank &= 255u;
        QD3_T(1);

        // ---- run-length first bit: the probability for this rank class is already there ----
        const u32 rank0 = rank - 1u;
        const u32 q = rank0 < 7u ? rank0 : 7u;
        const u32 pw = q < 2 ? pu.x : q < 4 ? pu.y : q < 6 ? pu.z : pu.w;
        const u32 st2 = ((q < 4 ? s2.x : s2.y) >> (8u * (q & 3u))) & 255u;
        const u32 bU = qd3_step(sm, rc, (q & 1u) ? pw >> 16 : pw & 0xffffu);

        QD3_T(2);
        // ---- update phase: the lanes that own a counter on the decoded path move it ----
        QD3_LANES {
            Qd3Lane &r = QD3_L(lr);
            bool onX = false; u32 bitX = 0;
            if (lane == 0) { onX = plain; bitX = bT; }
            else if (lane < 8u) { const u32 k = lane - 1u; onX = plain && bT && k < nE; bitX = (k + 1u < nE) ? 1u : lastE; }
            else if (lane < 16u) { onX = (lane - 8u) == q; bitX = bU; }
            if (onX) {
                sm.set(r.iSX, (r.sX * (bitX ? r.xMs1 : r.xMs0) + (bitX ? r.xKs1 : r.xKs0)) >> 12);
                sm.set(r.iCX, (r.cX * (bitX ? r.xMc1 : r.xMc0) + (bitX ? r.xKc1 : r.xKc0)) >> 12);
                sm.set(r.xG,  (r.gX * (bitX ? r.xMg1 : r.xMg0) + (bitX ? r.xKg1 : r.xKg0)) >> 12);
            }
            if (plain && bT && e <= M_MAXE) {
                const u32 shA = (e - r.dA) & 31u, shB = (e - r.dB) & 31u;
                const bool onA = r.okA && r.eA == e && (rank >> shA) == r.jA, onB = r.okB && r.eB == e && (rank >> shB) == r.jB;
                if (onA) {
                    const u32 bit = (rank >> ((shA - 1u) & 31u)) & 1u;
                    sm.set(R_RM_STATE + st * M_ROW + lane, bit ? q_down<K_RANK_M, 0>(r.sA) : q_up<K_RANK_M, 0>(r.sA));
                    sm.set(R_RM_CHAR + c * M_ROW + lane,   bit ? q_down<K_RANK_M, 1>(r.cA) : q_up<K_RANK_M, 1>(r.cA));
                    sm.set(r.gxA,                          bit ? q_down<K_RANK_M, 2>(r.gA) : q_up<K_RANK_M, 2>(r.gA));
                }
                if (onB) {
                    const u32 bit = (rank >> ((shB - 1u) & 31u)) & 1u;
                    sm.set(R_RM_STATE + st * M_ROW + lane + 32u, bit ? q_down<K_RANK_M, 0>(r.sB) : q_up<K_RANK_M, 0>(r.sB));
                    sm.set(R_RM_CHAR + c * M_ROW + lane + 32u,   bit ? q_down<K_RANK_M, 1>(r.cB) : q_up<K_RANK_M, 1>(r.cB));
                    sm.set(r.gxB,                                bit ? q_down<K_RANK_M, 2>(r.gB) : q_up<K_RANK_M, 2>(r.gB));
                }
            }
        }

        QD3_T(3);
        // ---- push c `rank` places back: mtf[0..rank-1] = mtf[1..rank]; mtf[rank] = c  (qlfc.cpp:1830-1860) ----
        const u32 cur = c;
        if (rank == 1) { sm.st8(O3_MTF, m1); sm.st8(O3_MTF + 1, cur); c = m1; m1 = cur; }
        else if (rank == 2) { sm.st8(O3_MTF, m1); sm.st8(O3_MTF + 1, m2); sm.st8(O3_MTF + 2, cur); c = m1; m1 = m2; m2 = cur; }
        else if (rank == 3) { sm.st8(O3_MTF, m1); sm.st8(O3_MTF + 1, m2); sm.st8(O3_MTF + 2, m3); sm.st8(O3_MTF + 3, cur); c = m1; m1 = m2; m2 = m3; m3 = cur; }
        else if (rank != 0) {
            QD3_SYNC();
            for (u32 basep = 0; basep < rank; basep += 32) {
                QD3_LANES { QD3_L(lr).tmp = sm.ld8(O3_MTF + basep + lane + 1u); }
                QD3_SYNC();
                QD3_LANES { if (basep + lane < rank) sm.st8(O3_MTF + basep + lane, QD3_L(lr).tmp); }
                QD3_SYNC();
            }
            sm.st8(O3_MTF + rank, cur);
            QD3_SYNC();
            c = sm.ld8(O3_MTF); m1 = sm.ld8(O3_MTF + 1); m2 = sm.ld8(O3_MTF + 2); m3 = sm.ld8(O3_MTF + 3);
        }
        // (c, m1, m2, m3) now describe the NEXT run; `cur` is this run's symbol
        const u32 rhRn = sm.ld8(O3_RANK_HIST + c), rhUn = sm.ld8(O3_RUN_HIST + c);
        avgRank = (avgRank * 124 + (int)rank * 4) >> 7;
        // both candidates for the next run's rank state (its ctxRun gets one more bit: run < 3)
        const u32 ctxRank4n = ((ctxRank4 << 2) | (rank0 < 3u ? rank0 : 3u)) & 0xffu;
        const u32 ctxRunN = (ctxRun << 1) & 0xfu;
        const u32 stA = sm.ld8(O3_RANK_STATE + (((ctxRunN | 1u) << 11) | (ctxRank4n << 3) | rhRn)), stB = sm.ld8(O3_RANK_STATE + ((ctxRunN << 11) | (ctxRank4n << 3) | rhRn));

        QD3_T(4);
        u32 run = 1;
        if (!bU) sm.st8(O3_RUN_HIST + cur, (rhU + 2u) >> 2);
        else {
            u32 eu = 1;
            for (;;) {
                const u32 k = eu - 1u;
                if (k < UE_RES) b = qd3_dec3<K_RUN_E>(sm, rc, R_UE_STATE + st2 * UE_RES + k, R_UE_CHAR + cur * UE_RES + k, R_UE_SHARED + k);
                else {
                    const u32 is = qd3_cache_get(sm, C_STATE_VAL, O3_TAG_STATE, cold_s, ue_idx(st2, k), st_miss);
                    const u32 ic = qd3_cache_get(sm, C_CHAR_VAL, O3_TAG_CHAR, cold_c, ue_idx(cur, k), st_miss);
                    st_cached += 2;
                    b = qd3_dec3<K_RUN_E>(sm, rc, is, ic, R_UE_SHARED + k);
                }
                if (!b) break;
                if (++eu >= 31u) break;                                          // corrupt-input guard
            }
            sm.st8(O3_RUN_HIST + cur, ((rhU + 3u * eu + 3u) >> 2) & 255u);
            if (eu <= M_MAXE) {
                const u32 bs = R_UM_STATE + st2 * M_ROW + (1u << eu) - 2u, bc = R_UM_CHAR + cur * M_ROW + (1u << eu) - 2u, bg = R_NARROW_SHARED + eu * 32u;
                for (u32 node = 1, bit = eu; bit > 0; --bit) {
                    b = qd3_dec3<K_RUN_M>(sm, rc, bs + node, bc + node, bg + node);
                    run = 2u * run + b; node = 2u * node + b;
                }
            } else {
                for (u32 node = 1, bit = eu; bit > 0; --bit) {
                    const u32 is = qd3_cache_get(sm, C_STATE_VAL, O3_TAG_STATE, cold_s, narrow_idx(eu, st2, node), st_miss);
                    const u32 ic = qd3_cache_get(sm, C_CHAR_VAL, O3_TAG_CHAR, cold_c, narrow_idx(eu, cur, node), st_miss);
                    st_cached += 2;
                    b = qd3_dec3<K_RUN_M>(sm, rc, is, ic, R_NARROW_SHARED + eu * 32u + node);
                    run = 2u * run + b; node = node + 1u;                        // qlfc.cpp:1119: linear contexts above 5 bits
                }
            }
        }
        QD3_T(5);
        const bool shortRun = run < 3u;
        ctxRank0 = ((ctxRank0 << 1) | (rank0 == 0u ? 1u : 0u)) & 0x7u;
        ctxRank4 = ctxRank4n;
        ctxRun   = ctxRunN | (shortRun ? 1u : 0u);
        st = shortRun ? stA : stB;
        rhU = rank != 0 ? rhUn : sm.ld8(O3_RUN_HIST + c);                         // rank 0 (corrupt input only): same symbol again

        if (run > n - i) run = n - i;                                            // never write past n
        QD3_LANES { for (u32 k = lane; k < run; k += 32) out[i + k] = (u8)cur; }
        i += run;
        QD3_T(6);
        if (PROF) ++prof_runs;
    }
#ifndef QD3_HOST
    if (PROF && blockIdx.x == 0 && threadIdx.x == 0)
        printf("[qdec3 prof] runs %u; cycles/run: eval %.1f walk %.1f runbit %.1f update %.1f mtf+next %.1f run>1 %.1f tail %.1f\n", prof_runs,
               (double)prof_t[0] / prof_runs, (double)prof_t[1] / prof_runs, (double)prof_t[2] / prof_runs, (double)prof_t[3] / prof_runs,
               (double)prof_t[4] / prof_runs, (double)prof_t[5] / prof_runs, (double)prof_t[6] / prof_runs);
#endif
    return (int)n;
}

// ---- the serial decoder (decision-by-decision walk, the structure of q_decode2) on this file's plumbing -------------
// Same prefetching as q_decode2 (the next run's first-decision counters and states are loaded while the current run
// is still being decoded), but: the branch-free range-coder step with a preloaded next unit, counter moves as selects
// (no two-sided branch per decision), the input window checked once per run, the front of the MTF list kept in
// registers and written back only when a rank > 3 needs the list in shared memory, and run expansion as ONE
// unconditional 32-byte store per run (later runs overwrite the excess; byte address A is always written by lane
// A mod 32, so the order of two stores to one address is program order of one thread).
template <int K> QD3_FN u32 qd3_dec3v(const SM3 &sm, Rc3 &rc, u32 is, u32 ic, u32 ig, int s, int c, int g)
{
    const u32 b = qd3_step(sm, rc, (u32)q_mix<K>(s, c, g));
    sm.set(is, b ? q_down<K, 0>(s) : q_up<K, 0>(s));
    sm.set(ic, b ? q_down<K, 1>(c) : q_up<K, 1>(c));
    sm.set(ig, b ? q_down<K, 2>(g) : q_up<K, 2>(g));
    return b;
}

template <bool PROF> QD3_FN int qd3_decode_stream_serial(const SM3 &sm, const u8 *__restrict__ in, u32 in_limit, u8 *__restrict__ out, u32 out_cap,
                                    short *__restrict__ cold_s, short *__restrict__ cold_c, u32 &st_cached, u32 &st_miss)
{
    QD3_LREGS;
#ifndef QD3_HOST
    const u32 lane = threadIdx.x & 31u;
#endif
    Rc3 rc; u32 n; int maxRank;
    { const int err = qd3_prologue(sm, rc, lr, in, in_limit, out_cap, n, maxRank); if (err) return err; }

    u32 ctxRank0 = 0, ctxRank4 = 0, ctxRun = 0; int avgRank = 0;
    u32 c, m1, m2, m3;
    { const u32 f = sm.ld32(O3_MTF); c = f & 255u; m1 = (f >> 8) & 255u; m2 = (f >> 16) & 255u; m3 = f >> 24; }
    u32 rhU = sm.ld8(O3_RUN_HIST + c);
    u32 st = sm.ld8(O3_RANK_STATE + ((ctxRun << 11) | (ctxRank4 << 3) | sm.ld8(O3_RANK_HIST + c)));
    int tS = sm.cnt(R_RT_STATE + st), tC = sm.cnt(R_RT_CHAR + c), tG = sm.cnt(R_RT_SHARED);
    u32 st2z = sm.ld8(O3_RUN_STATE + ((ctxRank0 << 10) | (ctxRun << 6) | (rhU < 7 ? rhU : 7)));       // run state if rank == 1

    long long prof_t[8] = {0, 0, 0, 0, 0, 0, 0, 0}, prof_last = 0; u32 prof_runs = 0;
    (void)prof_t; (void)prof_last; (void)prof_runs;
#ifndef QD3_HOST
    if (PROF) prof_last = clock64();
#endif
    for (u32 i = 0; i < n; ) {
        u32 rank = 1, b;
        const u32 rhq = rhU < 7 ? rhU : 7;
        const bool plain = avgRank < 32;
        if (rc.pos - rc.wbase > QD3_RUN_ROOM) QD3_REFILL();
        // first-decision counters of the run length, should the rank turn out to be 1 (the rank decisions never touch them)
        const int uS0 = sm.cnt(R_UT_STATE + st2z), uC0 = sm.cnt(R_UT_CHAR + c), uG0 = sm.cnt(R_UT_SHARED);
        QD3_T(0);
        if (plain) {
            b = qd3_dec3v<K_RANK_T>(sm, rc, R_RT_STATE + st, R_RT_CHAR + c, R_RT_SHARED, tS, tC, tG);
            if (!b) sm.st8(O3_RANK_HIST + c, 0);
            else {
                u32 e = 1;
                while ((int)e != maxRank) {
                    b = qd3_dec3<K_RANK_E>(sm, rc, R_RE_STATE + st * 8 + e - 1, R_RE_CHAR + c * 8 + e - 1, R_RE_SHARED + e - 1);
                    if (!b) break;
                    if (++e >= 7) break;                                      // e <= maxRank <= 7 in valid streams
                }
                sm.st8(O3_RANK_HIST + c, e);
                if (e <= M_MAXE) {
                    const u32 bs = R_RM_STATE + st * M_ROW + (1u << e) - 2u, bc = R_RM_CHAR + c * M_ROW + (1u << e) - 2u, bg = R_WIDE_SHARED + e * 256;
                    for (int bit = (int)e - 1; bit >= 0; --bit) {
                        b = qd3_dec3<K_RANK_M>(sm, rc, bs + rank, bc + rank, bg + rank);
                        rank = 2u * rank + b;
                    }
                } else {
                    for (int bit = (int)e - 1; bit >= 0; --bit) {
                        const u32 is = qd3_cache_get(sm, C_STATE_VAL, O3_TAG_STATE, cold_s, wide_idx(e, st, rank), st_miss);
                        const u32 ic = qd3_cache_get(sm, C_CHAR_VAL, O3_TAG_CHAR, cold_c, wide_idx(e, c, rank), st_miss);
                        st_cached += 2;
                        b = qd3_dec3<K_RANK_M>(sm, rc, is, ic, R_WIDE_SHARED + e * 256u + rank);
                        rank = 2u * rank + b;
                    }
                }
            }
        } else {
            rank = 0;
            for (int node = 1, bit = maxRank; bit >= 0; --bit) {
                const u32 is = qd3_cache_get(sm, C_STATE_VAL, O3_TAG_STATE, cold_s, wide_idx(8, st, (u32)node), st_miss);
                const u32 ic = qd3_cache_get(sm, C_CHAR_VAL, O3_TAG_CHAR, cold_c, wide_idx(8, c, (u32)node), st_miss);
                st_cached += 2;
                b = qd3_dec3<K_RANK_P>(sm, rc, is, ic, R_WIDE_SHARED + 8u * 256u + (u32)node);
                node = 2 * node + (int)b; rank = 2u * rank + b;
            }
            sm.st8(O3_RANK_HIST + c, (u32)qd3_ilog2(rank));
        }
        rank &= 255u;
        QD3_T(1);

        // push c `rank` places back (qlfc.cpp:1830-1860); positions 0..3 of the list live in (c, m1, m2, m3)
        const u32 cur = c;
        if (rank == 1) { c = m1; m1 = cur; }
        else if (rank == 2) { c = m1; m1 = m2; m2 = cur; }
        else if (rank == 3) { c = m1; m1 = m2; m2 = m3; m3 = cur; }
        else if (rank != 0) {
            sm.st8(O3_MTF, c); sm.st8(O3_MTF + 1, m1); sm.st8(O3_MTF + 2, m2); sm.st8(O3_MTF + 3, m3);
            QD3_SYNC();
            for (u32 basep = 0; basep < rank; basep += 32) {
                QD3_LANES { QD3_L(lr).tmp = sm.ld8(O3_MTF + basep + lane + 1u); }
                QD3_SYNC();
                QD3_LANES { if (basep + lane < rank) sm.st8(O3_MTF + basep + lane, QD3_L(lr).tmp); }
                QD3_SYNC();
            }
            sm.st8(O3_MTF + rank, cur);
            QD3_SYNC();
            const u32 f = sm.ld32(O3_MTF); c = f & 255u; m1 = (f >> 8) & 255u; m2 = (f >> 16) & 255u; m3 = f >> 24;
        }
        // (c, m1, m2, m3) now describe the NEXT run; `cur` is this run's symbol
        const u32 rhRn = sm.ld8(O3_RANK_HIST + c), rhUn = sm.ld8(O3_RUN_HIST + c);
        avgRank = (avgRank * 124 + (int)rank * 4) >> 7;
        const u32 rank0 = rank - 1u;
        u32 st2 = st2z, run = 1;
        QD3_T(4);
        if (rank0 == 0) b = qd3_dec3v<K_RUN_T>(sm, rc, R_UT_STATE + st2z, R_UT_CHAR + cur, R_UT_SHARED, uS0, uC0, uG0);
        else {
            st2 = sm.ld8(O3_RUN_STATE + ((ctxRank0 << 10) | (ctxRun << 6) | ((rank0 < 7u ? rank0 : 7u) << 3) | rhq));
            b = qd3_dec3<K_RUN_T>(sm, rc, R_UT_STATE + st2, R_UT_CHAR + cur, R_UT_SHARED);
        }
        // both candidates for the next run's rank state (its ctxRun gets one more bit: run < 3)
        const u32 ctxRank4n = ((ctxRank4 << 2) | (rank0 < 3u ? rank0 : 3u)) & 0xffu;
        const u32 ctxRunN = (ctxRun << 1) & 0xfu;
        const u32 stA = sm.ld8(O3_RANK_STATE + (((ctxRunN | 1u) << 11) | (ctxRank4n << 3) | rhRn)), stB = sm.ld8(O3_RANK_STATE + ((ctxRunN << 11) | (ctxRank4n << 3) | rhRn));
        QD3_T(2);
        if (!b) sm.st8(O3_RUN_HIST + cur, (rhU + 2u) >> 2);
        else {
            u32 eu = 1;
            for (;;) {
                const u32 k = eu - 1u;
                if (k < UE_RES) b = qd3_dec3<K_RUN_E>(sm, rc, R_UE_STATE + st2 * UE_RES + k, R_UE_CHAR + cur * UE_RES + k, R_UE_SHARED + k);
                else {
                    const u32 is = qd3_cache_get(sm, C_STATE_VAL, O3_TAG_STATE, cold_s, ue_idx(st2, k), st_miss);
                    const u32 ic = qd3_cache_get(sm, C_CHAR_VAL, O3_TAG_CHAR, cold_c, ue_idx(cur, k), st_miss);
                    st_cached += 2;
                    b = qd3_dec3<K_RUN_E>(sm, rc, is, ic, R_UE_SHARED + k);
                }
                if (!b) break;
                if (++eu >= 31u) break;                                          // corrupt-input guard
            }
            sm.st8(O3_RUN_HIST + cur, ((rhU + 3u * eu + 3u) >> 2) & 255u);
            if (eu <= M_MAXE) {
                const u32 bs = R_UM_STATE + st2 * M_ROW + (1u << eu) - 2u, bc = R_UM_CHAR + cur * M_ROW + (1u << eu) - 2u, bg = R_NARROW_SHARED + eu * 32u;
                for (u32 node = 1, bit = eu; bit > 0; --bit) {
                    b = qd3_dec3<K_RUN_M>(sm, rc, bs + node, bc + node, bg + node);
                    run = 2u * run + b; node = 2u * node + b;
                }
            } else {
                for (u32 node = 1, bit = eu; bit > 0; --bit) {
                    const u32 is = qd3_cache_get(sm, C_STATE_VAL, O3_TAG_STATE, cold_s, narrow_idx(eu, st2, node), st_miss);
                    const u32 ic = qd3_cache_get(sm, C_CHAR_VAL, O3_TAG_CHAR, cold_c, narrow_idx(eu, cur, node), st_miss);
                    st_cached += 2;
                    b = qd3_dec3<K_RUN_M>(sm, rc, is, ic, R_NARROW_SHARED + eu * 32u + node);
                    run = 2u * run + b; node = node + 1u;                        // qlfc.cpp:1119: linear contexts above 5 bits
                }
            }
        }
        QD3_T(5);
        const bool shortRun = run < 3u;
        ctxRank0 = ((ctxRank0 << 1) | (rank0 == 0u ? 1u : 0u)) & 0x7u;
        ctxRank4 = ctxRank4n;
        ctxRun   = ctxRunN | (shortRun ? 1u : 0u);
        st = shortRun ? stA : stB;
        rhU = rank != 0 ? rhUn : sm.ld8(O3_RUN_HIST + c);                         // rank 0 (corrupt input only): same symbol again
        // first-decision counters of the next run (nothing writes the rank counters until then) and its run state for rank 1
        tS = sm.cnt(R_RT_STATE + st); tC = sm.cnt(R_RT_CHAR + c); tG = sm.cnt(R_RT_SHARED);
        st2z = sm.ld8(O3_RUN_STATE + ((ctxRank0 << 10) | (ctxRun << 6) | (rhU < 7 ? rhU : 7)));

        // run expansion: byte address A is always written by lane A mod 32
        if (run <= 32u && i + 32u <= n) { QD3_LANES { out[i + ((lane - i) & 31u)] = (u8)cur; } }
        else {
            if (run > n - i) run = n - i;                                        // never write past n
            QD3_LANES { for (u32 k = (lane - i) & 31u; k < run; k += 32) out[i + k] = (u8)cur; }
        }
        i += run;
        QD3_T(6);
        if (PROF) ++prof_runs;
    }
#ifndef QD3_HOST
    if (PROF && blockIdx.x == 0 && threadIdx.x == 0)
        printf("[qdec3 serial prof] runs %u; cycles/run: top %.1f rank %.1f runbit %.1f mtf+hist %.1f run>1 %.1f tail %.1f\n", prof_runs,
               (double)prof_t[0] / prof_runs, (double)prof_t[1] / prof_runs, (double)prof_t[2] / prof_runs, (double)prof_t[4] / prof_runs,
               (double)prof_t[5] / prof_runs, (double)prof_t[6] / prof_runs);
#endif
    return (int)n;
}

// ---- the pipelined serial decoder: two-way speculation inside the uniform instruction stream ---------------------------
// Measured on q_decode3<1> (profiles/r1h): a rank decision costs ~138 cycles, most of it the chain address -> three
// shared-memory loads -> three-term mix in front of the range-coder step.  Here the counters of BOTH possible next
// decisions (exponent: next exponent decision | root of the mantissa tree; mantissa node: left | right child; last
// mantissa decision: the two candidate run-length first-bit contexts) are loaded and mixed while the current decision
// is being resolved; the outcome only selects.  More instructions, shorter chain.
struct C3 { u32 is, ic, ig; int s, c, g; };
QD3_FN C3 qd3_load3(const SM3 &sm, u32 is, u32 ic, u32 ig) { C3 x; x.is = is; x.ic = ic; x.ig = ig; x.s = sm.cnt(is); x.c = sm.cnt(ic); x.g = sm.cnt(ig); return x; }
template <int K> QD3_FN u32 qd3_decide(const SM3 &sm, Rc3 &rc, const C3 &x, u32 p)
{
    const u32 b = qd3_step(sm, rc, p);
    sm.set(x.is, b ? q_down<K, 0>(x.s) : q_up<K, 0>(x.s));
    sm.set(x.ic, b ? q_down<K, 1>(x.c) : q_up<K, 1>(x.c));
    sm.set(x.ig, b ? q_down<K, 2>(x.g) : q_up<K, 2>(x.g));
    return b;
}

template <bool PROF> QD3_FN int qd3_decode_stream_pipe(const SM3 &sm, const u8 *__restrict__ in, u32 in_limit, u8 *__restrict__ out, u32 out_cap,
                                    short *__restrict__ cold_s, short *__restrict__ cold_c, u32 &st_cached, u32 &st_miss)
{
    QD3_LREGS;
#ifndef QD3_HOST
    const u32 lane = threadIdx.x & 31u;
#endif
    Rc3 rc; u32 n; int maxRank;
    { const int err = qd3_prologue(sm, rc, lr, in, in_limit, out_cap, n, maxRank); if (err) return err; }

    u32 ctxRank0 = 0, ctxRank4 = 0, ctxRun = 0; int avgRank = 0;
    u32 c, m1, m2, m3;
    { const u32 f = sm.ld32(O3_MTF); c = f & 255u; m1 = (f >> 8) & 255u; m2 = (f >> 16) & 255u; m3 = f >> 24; }
    u32 rhU = sm.ld8(O3_RUN_HIST + c);
    u32 st = sm.ld8(O3_RANK_STATE + ((ctxRun << 11) | (ctxRank4 << 3) | sm.ld8(O3_RANK_HIST + c)));
    int tS = sm.cnt(R_RT_STATE + st), tC = sm.cnt(R_RT_CHAR + c), tG = sm.cnt(R_RT_SHARED);
    u32 st2z = sm.ld8(O3_RUN_STATE + ((ctxRank0 << 10) | (ctxRun << 6) | (rhU < 7 ? rhU : 7)));       // run state if rank == 1

    long long prof_t[8] = {0, 0, 0, 0, 0, 0, 0, 0}, prof_last = 0; u32 prof_runs = 0;
    (void)prof_t; (void)prof_last; (void)prof_runs;
#ifndef QD3_HOST
    if (PROF) prof_last = clock64();
#endif
    for (u32 i = 0; i < n; ) {
        u32 rank = 1, b;
        const u32 rhq = rhU < 7 ? rhU : 7;
        const bool plain = avgRank < 32;
        if (rc.pos - rc.wbase > QD3_RUN_ROOM) QD3_REFILL();
        // first-decision counters of the run length, should the rank turn out to be 1 (the rank decisions never touch them)
        const int uS0 = sm.cnt(R_UT_STATE + st2z), uC0 = sm.cnt(R_UT_CHAR + c), uG0 = sm.cnt(R_UT_SHARED);
        QD3_T(0);
        bool haveU = false; u32 st2 = st2z, pU = 0; int uS = uS0;
        if (plain) {
            // Two-way speculation: while a decision is being resolved, the counters of BOTH possible next decisions are
            // already loaded and mixed, so a load never sits between two decisions.
            C3 T; T.is = R_RT_STATE + st; T.ic = R_RT_CHAR + c; T.ig = R_RT_SHARED; T.s = tS; T.c = tC; T.g = tG;
            C3 E = qd3_load3(sm, R_RE_STATE + st * 8, R_RE_CHAR + c * 8, R_RE_SHARED);             // exponent decision 0, needed if the first bit is 1
            u32 pE = (u32)q_mix<K_RANK_E>(E.s, E.c, E.g);
            b = qd3_decide<K_RANK_T>(sm, rc, T, (u32)q_mix<K_RANK_T>(T.s, T.c, T.g));
            if (!b) sm.st8(O3_RANK_HIST + c, 0);
            else {
                u32 e = 1;
                C3 M; M.is = M.ic = M.ig = 0; M.s = M.c = M.g = 0; u32 pM = 0; bool haveM = false;
                while ((int)e != maxRank) {
                    // after exponent decision e-1: 1 -> exponent decision e, 0 -> the root of the mantissa tree of level e
                    const C3 En = qd3_load3(sm, R_RE_STATE + st * 8 + e, R_RE_CHAR + c * 8 + e, R_RE_SHARED + e);
                    const u32 pEn = (u32)q_mix<K_RANK_E>(En.s, En.c, En.g);
                    C3 Mr = M; u32 pMr = 0;
                    if (e <= M_MAXE) {
                        Mr = qd3_load3(sm, R_RM_STATE + st * M_ROW + (1u << e) - 1u, R_RM_CHAR + c * M_ROW + (1u << e) - 1u, R_WIDE_SHARED + e * 256u + 1u);
                        pMr = (u32)q_mix<K_RANK_M>(Mr.s, Mr.c, Mr.g);
                    }
                    b = qd3_decide<K_RANK_E>(sm, rc, E, pE);
                    if (!b) { M = Mr; pM = pMr; haveM = e <= M_MAXE; break; }
                    if (++e >= 7) break;                                      // e <= maxRank <= 7 in valid streams
                    E = En; pE = pEn;
                }
                sm.st8(O3_RANK_HIST + c, e);
                if (e <= M_MAXE) {
                    const u32 bs = R_RM_STATE + st * M_ROW + (1u << e) - 2u, bc = R_RM_CHAR + c * M_ROW + (1u << e) - 2u, bg = R_WIDE_SHARED + e * 256u;
                    if (!haveM) { M = qd3_load3(sm, bs + 1u, bc + 1u, bg + 1u); pM = (u32)q_mix<K_RANK_M>(M.s, M.c, M.g); }
                    u32 node = 1;
                    for (int bit = (int)e - 1; bit > 0; --bit) {
                        const C3 c0 = qd3_load3(sm, bs + 2u * node, bc + 2u * node, bg + 2u * node), c1 = qd3_load3(sm, bs + 2u * node + 1u, bc + 2u * node + 1u, bg + 2u * node + 1u);
                        const u32 p0 = (u32)q_mix<K_RANK_M>(c0.s, c0.c, c0.g), p1 = (u32)q_mix<K_RANK_M>(c1.s, c1.c, c1.g);
                        b = qd3_decide<K_RANK_M>(sm, rc, M, pM);
                        node = 2u * node + b;
                        M.is = b ? c1.is : c0.is; M.ic = b ? c1.ic : c0.ic; M.ig = b ? c1.ig : c0.ig;
                        M.s = b ? c1.s : c0.s; M.c = b ? c1.c : c0.c; M.g = b ? c1.g : c0.g; pM = b ? p1 : p0;
                    }
                    {   // last mantissa decision: rank = 2 node + b, so both candidates of the run-length first bit are known
                        const u32 qa = 2u * node - 1u < 7u ? 2u * node - 1u : 7u, qb = 2u * node < 7u ? 2u * node : 7u;
                        const u32 sb2 = O3_RUN_STATE + ((ctxRank0 << 10) | (ctxRun << 6) | rhq);
                        const u32 st2a = sm.ld8(sb2 + (qa << 3)), st2b = sm.ld8(sb2 + (qb << 3));
                        const int ua = sm.cnt(R_UT_STATE + st2a), ub = sm.cnt(R_UT_STATE + st2b);
                        const u32 pa = (u32)q_mix<K_RUN_T>(ua, uC0, uG0), pb = (u32)q_mix<K_RUN_T>(ub, uC0, uG0);
                        b = qd3_decide<K_RANK_M>(sm, rc, M, pM);
                        node = 2u * node + b;
                        st2 = b ? st2b : st2a; uS = b ? ub : ua; pU = b ? pb : pa; haveU = true;
                    }
                    rank = node;
                } else {
                    for (int bit = (int)e - 1; bit >= 0; --bit) {
                        const u32 is = qd3_cache_get(sm, C_STATE_VAL, O3_TAG_STATE, cold_s, wide_idx(e, st, rank), st_miss);
                        const u32 ic = qd3_cache_get(sm, C_CHAR_VAL, O3_TAG_CHAR, cold_c, wide_idx(e, c, rank), st_miss);
                        st_cached += 2;
                        b = qd3_dec3<K_RANK_M>(sm, rc, is, ic, R_WIDE_SHARED + e * 256u + rank);
                        rank = 2u * rank + b;
                    }
                }
            }
        } else {
            rank = 0;
            for (int node = 1, bit = maxRank; bit >= 0; --bit) {
                const u32 is = qd3_cache_get(sm, C_STATE_VAL, O3_TAG_STATE, cold_s, wide_idx(8, st, (u32)node), st_miss);
                const u32 ic = qd3_cache_get(sm, C_CHAR_VAL, O3_TAG_CHAR, cold_c, wide_idx(8, c, (u32)node), st_miss);
                st_cached += 2;
                b = qd3_dec3<K_RANK_P>(sm, rc, is, ic, R_WIDE_SHARED + 8u * 256u + (u32)node);
                node = 2 * node + (int)b; rank = 2u * rank + b;
            }
            sm.st8(O3_RANK_HIST + c, (u32)qd3_ilog2(rank));
        }
        rank &= 255u;
        QD3_T(1);

        // push c `rank` places back (qlfc.cpp:1830-1860); positions 0..3 of the list live in (c, m1, m2, m3)
        const u32 cur = c;
        if (rank == 1) { c = m1; m1 = cur; }
        else if (rank == 2) { c = m1; m1 = m2; m2 = cur; }
        else if (rank == 3) { c = m1; m1 = m2; m2 = m3; m3 = cur; }
        else if (rank != 0) {
            sm.st8(O3_MTF, c); sm.st8(O3_MTF + 1, m1); sm.st8(O3_MTF + 2, m2); sm.st8(O3_MTF + 3, m3);
            QD3_SYNC();
            for (u32 basep = 0; basep < rank; basep += 32) {
                QD3_LANES { QD3_L(lr).tmp = sm.ld8(O3_MTF + basep + lane + 1u); }
                QD3_SYNC();
                QD3_LANES { if (basep + lane < rank) sm.st8(O3_MTF + basep + lane, QD3_L(lr).tmp); }
                QD3_SYNC();
            }
            sm.st8(O3_MTF + rank, cur);
            QD3_SYNC();
            const u32 f = sm.ld32(O3_MTF); c = f & 255u; m1 = (f >> 8) & 255u; m2 = (f >> 16) & 255u; m3 = f >> 24;
        }
        // (c, m1, m2, m3) now describe the NEXT run; `cur` is this run's symbol
        const u32 rhRn = sm.ld8(O3_RANK_HIST + c), rhUn = sm.ld8(O3_RUN_HIST + c);
        avgRank = (avgRank * 124 + (int)rank * 4) >> 7;
        const u32 rank0 = rank - 1u;
        u32 run = 1;
        QD3_T(4);
        if (rank0 == 0) b = qd3_dec3v<K_RUN_T>(sm, rc, R_UT_STATE + st2z, R_UT_CHAR + cur, R_UT_SHARED, uS0, uC0, uG0);
        else if (haveU) {
            C3 U; U.is = R_UT_STATE + st2; U.ic = R_UT_CHAR + cur; U.ig = R_UT_SHARED; U.s = uS; U.c = uC0; U.g = uG0;
            b = qd3_decide<K_RUN_T>(sm, rc, U, pU);
        } else {
            st2 = sm.ld8(O3_RUN_STATE + ((ctxRank0 << 10) | (ctxRun << 6) | ((rank0 < 7u ? rank0 : 7u) << 3) | rhq));
            b = qd3_dec3<K_RUN_T>(sm, rc, R_UT_STATE + st2, R_UT_CHAR + cur, R_UT_SHARED);
        }
        // both candidates for the next run's rank state (its ctxRun gets one more bit: run < 3)
        const u32 ctxRank4n = ((ctxRank4 << 2) | (rank0 < 3u ? rank0 : 3u)) & 0xffu;
        const u32 ctxRunN = (ctxRun << 1) & 0xfu;
        const u32 stA = sm.ld8(O3_RANK_STATE + (((ctxRunN | 1u) << 11) | (ctxRank4n << 3) | rhRn)), stB = sm.ld8(O3_RANK_STATE + ((ctxRunN << 11) | (ctxRank4n << 3) | rhRn));
        QD3_T(2);
        if (!b) sm.st8(O3_RUN_HIST + cur, (rhU + 2u) >> 2);
        else {
            u32 eu = 1;
            for (;;) {
                const u32 k = eu - 1u;
                if (k < UE_RES) b = qd3_dec3<K_RUN_E>(sm, rc, R_UE_STATE + st2 * UE_RES + k, R_UE_CHAR + cur * UE_RES + k, R_UE_SHARED + k);
                else {
                    const u32 is = qd3_cache_get(sm, C_STATE_VAL, O3_TAG_STATE, cold_s, ue_idx(st2, k), st_miss);
                    const u32 ic = qd3_cache_get(sm, C_CHAR_VAL, O3_TAG_CHAR, cold_c, ue_idx(cur, k), st_miss);
                    st_cached += 2;
                    b = qd3_dec3<K_RUN_E>(sm, rc, is, ic, R_UE_SHARED + k);
                }
                if (!b) break;
                if (++eu >= 31u) break;                                          // corrupt-input guard
            }
            sm.st8(O3_RUN_HIST + cur, ((rhU + 3u * eu + 3u) >> 2) & 255u);
            if (eu <= M_MAXE) {
                const u32 bs = R_UM_STATE + st2 * M_ROW + (1u << eu) - 2u, bc = R_UM_CHAR + cur * M_ROW + (1u << eu) - 2u, bg = R_NARROW_SHARED + eu * 32u;
                for (u32 node = 1, bit = eu; bit > 0; --bit) {
                    b = qd3_dec3<K_RUN_M>(sm, rc, bs + node, bc + node, bg + node);
                    run = 2u * run + b; node = 2u * node + b;
                }
            } else {
                for (u32 node = 1, bit = eu; bit > 0; --bit) {
                    const u32 is = qd3_cache_get(sm, C_STATE_VAL, O3_TAG_STATE, cold_s, narrow_idx(eu, st2, node), st_miss);
                    const u32 ic = qd3_cache_get(sm, C_CHAR_VAL, O3_TAG_CHAR, cold_c, narrow_idx(eu, cur, node), st_miss);
                    st_cached += 2;
                    b = qd3_dec3<K_RUN_M>(sm, rc, is, ic, R_NARROW_SHARED + eu * 32u + node);
                    run = 2u * run + b; node = node + 1u;                        // qlfc.cpp:1119: linear contexts above 5 bits
                }
            }
        }
        QD3_T(5);
        const bool shortRun = run < 3u;
        ctxRank0 = ((ctxRank0 << 1) | (rank0 == 0u ? 1u : 0u)) & 0x7u;
        ctxRank4 = ctxRank4n;
        ctxRun   = ctxRunN | (shortRun ? 1u : 0u);
        st = shortRun ? stA : stB;
        rhU = rank != 0 ? rhUn : sm.ld8(O3_RUN_HIST + c);                         // rank 0 (corrupt input only): same symbol again
        // first-decision counters of the next run (nothing writes the rank counters until then) and its run state for rank 1
        tS = sm.cnt(R_RT_STATE + st); tC = sm.cnt(R_RT_CHAR + c); tG = sm.cnt(R_RT_SHARED);
        st2z = sm.ld8(O3_RUN_STATE + ((ctxRank0 << 10) | (ctxRun << 6) | (rhU < 7 ? rhU : 7)));

        // run expansion: byte address A is always written by lane A mod 32
        if (run <= 32u && i + 32u <= n) { QD3_LANES { out[i + ((lane - i) & 31u)] = (u8)cur; } }
        else {
            if (run > n - i) run = n - i;                                        // never write past n
            QD3_LANES { for (u32 k = (lane - i) & 31u; k < run; k += 32) out[i + k] = (u8)cur; }
        }
        i += run;
        QD3_T(6);
        if (PROF) ++prof_runs;
    }
#ifndef QD3_HOST
    if (PROF && blockIdx.x == 0 && threadIdx.x == 0)
        printf("[qdec3 pipe prof] runs %u; cycles/run: top %.1f rank %.1f runbit %.1f mtf+hist %.1f run>1 %.1f tail %.1f\n", prof_runs,
               (double)prof_t[0] / prof_runs, (double)prof_t[1] / prof_runs, (double)prof_t[2] / prof_runs, (double)prof_t[4] / prof_runs,
               (double)prof_t[5] / prof_runs, (double)prof_t[6] / prof_runs);
#endif
    return (int)n;
}

#ifndef QD3_HOST
template <int MODE, bool PROF> __global__ void __launch_bounds__(32, 1) q_decode3(const u8 *__restrict__ in_all, SubBlock *__restrict__ sbs, short *__restrict__ cold_all,
                                                   const QTables *__restrict__ tables, u8 *__restrict__ out_all, const u32 *__restrict__ sb_list)
{
    extern __shared__ __align__(16) u8 q_smem_raw[];
    Dec3Smem &D = *reinterpret_cast<Dec3Smem *>(q_smem_raw);
    coder_smem_init(D.cs, tables);
    {
        u32 *w = (u32 *)D.px;
        for (u32 i = threadIdx.x; i < (sizeof(Dec3Smem) - offsetof(Dec3Smem, px)) / 4; i += 32) w[i] = 0;
        __syncwarp();
    }
    SM3 sm; sm.b = (u32)__cvta_generic_to_shared(q_smem_raw);
    asm volatile("" : "+r"(sm.b) :: "memory");               // from here on: explicit addresses only

    const u32 sid = sb_list[blockIdx.x];
    SubBlock &sb = sbs[sid];
    short *cold_s = cold_all + (size_t)blockIdx.x * 2 * COLD_PAD, *cold_c = cold_s + COLD_PAD;
    u32 st_cached = 0, st_miss = 0;
    const int r = MODE == 0 ? qd3_decode_stream<PROF>(sm, in_all + sb.out_off, sb.out_cap, out_all + sb.in_start, sb.in_size, cold_s, cold_c, st_cached, st_miss)
                : MODE == 1 ? qd3_decode_stream_serial<PROF>(sm, in_all + sb.out_off, sb.out_cap, out_all + sb.in_start, sb.in_size, cold_s, cold_c, st_cached, st_miss)
                            : qd3_decode_stream_pipe<PROF>(sm, in_all + sb.out_off, sb.out_cap, out_all + sb.in_start, sb.in_size, cold_s, cold_c, st_cached, st_miss);
    if (threadIdx.x == 0) { sb.result = r; sb.stat_cached = st_cached; sb.stat_miss = st_miss; }
}
#endif
