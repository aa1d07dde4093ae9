// libbsc_b200/csrc/qlfc_coder.cuh -- QLFC stage 2 (context model + binary range coder), device engine.
// Included by qlfc.cu inside its anonymous namespace (needs SubBlock, QTables, K_*, bscb_param).
//
// Why this looks the way it does.  One coder stream is a strictly serial recurrence (every binary
// decision reads three adaptive counters whose addresses depend on the previous decisions, then
// updates them and the range coder).  The format allows only <= 8 streams per block
// (coder.cpp:52-59), so per-stream LATENCY is everything.  Measured history (profiles/):
//   r1a  counters in global memory                 ~370 cycles / decision (each load an L2 round trip)
//   r1b  16 K-entry shared-memory caches           ~200 cycles / decision (working set did not fit)
// so the counter file is now laid out for shared memory: one SM per stream holds
//
//   * state tables                                  40 KB  (tables.h data, qlfc_tables.inc)
//   * first-bit, unary-exponent (index < 8) and model-wide shared counters      25 KB
//   * mantissa counters for exponents 1..5, stored COMPACTLY (2^e - 1 tree nodes per row instead of
//     the reference's 256/32-wide rows: 62 entries per state / symbol)          124 KB
//   * two small direct-mapped write-back caches (4 K entries each) for everything rare: rank
//     exponents 6-7, the escape bank, run exponents >= 6 and run-exponent indices >= 8; the
//     1.7 M counters behind them live in HBM                                    24 KB
//
// The decoder warp runs in lock-step: all 32 lanes execute the same decision sequence on the same
// data (shared-memory broadcasts), lanes only diverge to expand runs / rotate the MTF list.
// The encoder is a two-warp pipeline, see q_encode3 below.
#pragma once

// ---- shared-memory counter file (indices in u16 units) ---------------------------------------------
constexpr u32 R_RT_SHARED = 0, R_RT_STATE = 2, R_RT_CHAR = R_RT_STATE + 256;
constexpr u32 R_RE_SHARED = R_RT_CHAR + 256, R_RE_STATE = R_RE_SHARED + 8, R_RE_CHAR = R_RE_STATE + 2048;
constexpr u32 R_UT_SHARED = R_RE_CHAR + 2048, R_UT_STATE = R_UT_SHARED + 2, R_UT_CHAR = R_UT_STATE + 256;
constexpr u32 UE_RES = 8;                                   // run-exponent indices kept resident
constexpr u32 R_UE_SHARED = R_UT_CHAR + 256, R_UE_STATE = R_UE_SHARED + 32, R_UE_CHAR = R_UE_STATE + 256 * UE_RES;
constexpr u32 R_WIDE_SHARED = R_UE_CHAR + 256 * UE_RES;    // 9 banks x 256 (rank mantissa e=0..7, escape)
constexpr u32 R_NARROW_SHARED = R_WIDE_SHARED + 9 * 256;   // 32 banks x 32 (run mantissa)
constexpr u32 M_ROW = 62;                                   // compact row: exponents 1..5 -> offsets 2^e-2 .. 2^(e+1)-3
constexpr u32 M_MAXE = 5;
constexpr u32 R_RM_STATE = R_NARROW_SHARED + 32 * 32, R_RM_CHAR = R_RM_STATE + 256 * M_ROW;
constexpr u32 R_UM_STATE = R_RM_CHAR + 256 * M_ROW, R_UM_CHAR = R_UM_STATE + 256 * M_ROW;
constexpr u32 R_END = R_UM_CHAR + 256 * M_ROW;

constexpr int  QC_LOG = 12;                                 // 4 K entries per cache
constexpr u32  QC_SLOTS = 1u << QC_LOG, QC_MASK = QC_SLOTS - 1;
constexpr u32  C_STATE_VAL = R_END;                         // cache of the by-state rare banks
constexpr u32  C_CHAR_VAL = C_STATE_VAL + QC_SLOTS;         // cache of the by-symbol rare banks
constexpr u32  S16_COUNT = (C_CHAR_VAL + QC_SLOTS + 1) & ~1u;
// index space behind each cache: 9 wide banks x [256][256], 32 narrow banks x [256][32], run exponent [256][32]
constexpr u32  COLD_WIDE = 65536, COLD_NARROW = 8192;
constexpr u32  COLD_UE = 9 * COLD_WIDE + 32 * COLD_NARROW;
constexpr u32  COLD_COUNT = COLD_UE + 8192;
constexpr u32  COLD_PAD = (COLD_COUNT + 255) & ~255u;
// Slot of a cold index: 11 hashed bits + 1 class bit (rank banks / run banks never share a slot, so the
// encoder's rank-model warp and run-model warp can use the caches concurrently).
// 10 hashed bits + 2 class bits: classes 0,1 = rank banks (rank mantissa e>5, escape), 2 = run mantissa
// e>5, 3 = run exponent index >= 8.  The encoder's model warps each touch exactly one class group, so
// they can use the caches concurrently without ever sharing a slot.
__device__ __forceinline__ u32 cache_slot(u32 idx)
{
    const u32 h = idx >> 10;
    const u32 cls = idx < 9u * 65536u ? (h & 1u) : (idx < 9u * 65536u + 32u * 8192u ? 2u : 3u);
    return (cls << 10) | ((idx ^ (h * 1237u)) & 1023u);
}
__device__ __forceinline__ u32 cache_tag(u32 idx) { return (idx >> 10) + 1u; }
__device__ __forceinline__ u32 cache_unslot(u32 slot, u32 tag) { const u32 h = tag - 1u; return (h << 10) | (((slot & 1023u) ^ (h * 1237u)) & 1023u); }

struct CoderSmem {
    u8    rank_state[32768];
    u8    run_state[8192];
    u16   s16[S16_COUNT];
    u16   tag_state[QC_SLOTS];   // 0 = empty, else 1 + (cold index >> 11)
    u16   tag_char[QC_SLOTS];
    u8    rankHist[256], runHist[256];
    u8    mtf[256 + 32];
    __align__(16) u8 inwin[256];     // decoder: staged window of the input stream
};

__device__ __forceinline__ void coder_smem_init(CoderSmem &S, const QTables *__restrict__ g)
{
    const u32 lane = threadIdx.x & 31;
    const uint4 *src = (const uint4 *)g; uint4 *dst = (uint4 *)S.rank_state;      // rank_state and run_state are contiguous
    for (u32 i = lane; i < sizeof(QTables) / 16; i += 32) dst[i] = src[i];
    u32 *w = (u32 *)S.s16;
    for (u32 i = lane; i < S16_COUNT / 2; i += 32) w[i] = 0x08000800u;            // every counter starts at 2048
    u32 *t = (u32 *)S.tag_state;
    for (u32 i = lane; i < (2 * QC_SLOTS * 2 + 512 + 288) / 4; i += 32) t[i] = 0; // tags, histories, mtf
    __syncwarp();
}

__device__ __forceinline__ u32 wide_idx(u32 bank, u32 x, u32 node) { return bank * COLD_WIDE + x * 256 + node; }
__device__ __forceinline__ u32 narrow_idx(u32 e, u32 x, u32 node) { return 9 * COLD_WIDE + e * COLD_NARROW + x * 32 + node; }
__device__ __forceinline__ u32 ue_idx(u32 x, u32 k) { return COLD_UE + x * 32 + k; }
__device__ __forceinline__ u32 m_off(u32 e, u32 node) { return (1u << e) - 2u + node; }   // position inside a compact row

// Index (into S.s16) of rare counter `idx` of one kind, through the direct-mapped write-back cache.
// Warp-uniform version (decoder): every lane performs the same accesses.
__device__ __forceinline__ u32 cache_get(CoderSmem &S, u32 val_base, u16 *tags, short *__restrict__ cold, u32 idx, u32 &misses)
{
    const u32 slot = cache_slot(idx), want = cache_tag(idx);
    const u32 t = tags[slot];
    if (t != want) {
        if (t) cold[cache_unslot(slot, t)] = (short)S.s16[val_base + slot];
        S.s16[val_base + slot] = (u16)cold[idx];
        tags[slot] = (u16)want;
        ++misses;
    }
    return val_base + slot;
}

// ---- counters ---------------------------------------------------------------------------------------
// All parameters are compile-time immediates (bscb_param is constexpr).  Counters provably stay in
// [1, 4095] from their start value 2048 for every parameter set, so they are kept as unsigned 16-bit.
template <int K> __device__ __forceinline__ int q_mix(int s, int c, int g)
{
    return (c * bscb_param(K, 0) + s * bscb_param(K, 1) + g * bscb_param(K, 2)) >> 5;
}
// Counter moves as ONE multiply-add and ONE shift.  With integer p:
//   bit 0 (predictor.h:50-53):  p + floor(((4096-TH0-p)*AR0)/4096) = floor((p*(4096-AR0) + (4096-TH0)*AR0) / 4096)
//   bit 1 (predictor.h:55-58):  p - floor(((p-TH1)*AR1)/4096)      = floor((p*(4096-AR1) + TH1*AR1 + 4095) / 4096)
// (checked exhaustively for p in [0,4096] and all 21 parameter sets; tests/test_oracle.py)
template <int K, int WHO> __device__ __forceinline__ int q_up(int p)
{
    return (p * (4096 - bscb_param(K, 4 + 4 * WHO)) + (4096 - bscb_param(K, 3 + 4 * WHO)) * bscb_param(K, 4 + 4 * WHO)) >> 12;
}
template <int K, int WHO> __device__ __forceinline__ int q_down(int p)
{
    return (p * (4096 - bscb_param(K, 6 + 4 * WHO)) + bscb_param(K, 5 + 4 * WHO) * bscb_param(K, 6 + 4 * WHO) + 4095) >> 12;
}

// ---- range coder (rangecoder.h:38-271), 16-bit units -------------------------------------------------
struct Rc2Enc {
    u64 low;                          // bits 0..31 = low32, bit 32 = pending carry (rangecoder.h:43-52 `ari.low`)
    u32 range, cache, pending, pos;
    u8 *out;
    __device__ __forceinline__ void init(u8 *o) { low = 0; range = 0xffffffffu; cache = 0; pending = 0; pos = 0; out = o; }
    __device__ __forceinline__ void put16(u32 v) { *(u16 *)(out + pos) = (u16)v; pos += 2; }   // all lanes store the same value
    __device__ __forceinline__ void shift() {
        const u32 low32 = (u32)low, carry = (u32)(low >> 32);
        if (low32 < 0xffff0000u || carry) {
            put16(cache + carry);
#pragma unroll 1
            for (; pending; --pending) put16(carry - 1);        // one time in 65536: keep it small
            cache = low32 >> 16;
        } else pending++;
        low = (u64)(u32)(low32 << 16);
    }
    // branch-free step: low += bit ? r : 0 ; range = bit ? range - r : r
    __device__ __forceinline__ void step(u32 bit, u32 p) {
        const u32 r = (range >> 12) * p;
        low += bit ? (u64)r : 0ull;
        range = bit ? range - r : r;
    }
    __device__ __forceinline__ void encode(u32 bit, int p) {
        if (range < 0x10000u) { shift(); range <<= 16; }
        step(bit, (u32)p);
    }
    __device__ u32 finish() { if (range < 0x10000u) shift(); shift(); shift(); shift(); return pos; }
};

// cold path of the decoder's input window (by-value arguments: keeps the coder state in registers)
__device__ __noinline__ void rc_refill_window(const u8 *__restrict__ in, u8 *win, u32 base, u32 limit)
{
    const u32 lane = threadIdx.x & 31;
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 8; ++k) { const u32 o = base + lane * 8 + k; win[lane * 8 + k] = o < limit ? in[o] : (u8)0; }
    __syncwarp();
}

struct Rc2Dec {
    const u8 *in; u32 pos, limit, code, range;
    u8 *win; u32 wbase;              // 256-byte shared-memory window [wbase, wbase+256) of the stream (pos is always even)
    __device__ __forceinline__ void refill() { wbase = pos; rc_refill_window(in, win, pos, limit); }
    __device__ __forceinline__ u32 get16() {
        if (pos - wbase >= 256u) refill();
        const u32 v = *(const u16 *)(win + (pos - wbase));
        pos += 2; return v;
    }
    __device__ __forceinline__ u32 decode(int p) {
        if (range < 0x10000u) { range <<= 16; code = (code << 16) | get16(); }
        const u32 r = (range >> 12) * (u32)p;
        const u32 bit = code >= r;
        code -= bit ? r : 0u; range = bit ? range - r : r;
        return bit;
    }
};

// one binary decision against three shared-memory counters (indices into S.s16), decoder side
template <int K> __device__ __forceinline__ u32 dec3(CoderSmem &S, Rc2Dec &rc, u32 is, u32 ic, u32 ig)
{
    const int s = S.s16[is], c = S.s16[ic], g = S.s16[ig];
    const int p = q_mix<K>(s, c, g);
    if (rc.range < 0x10000u) { rc.range <<= 16; rc.code = (rc.code << 16) | rc.get16(); }
    const u32 r = (rc.range >> 12) * (u32)p;
    if (rc.code >= r) {                                      // warp-uniform branch: one side does everything for its outcome
        rc.code -= r; rc.range -= r;
        S.s16[is] = (u16)q_down<K, 0>(s); S.s16[ic] = (u16)q_down<K, 1>(c); S.s16[ig] = (u16)q_down<K, 2>(g);
        return 1u;
    }
    rc.range = r;
    S.s16[is] = (u16)q_up<K, 0>(s); S.s16[ic] = (u16)q_up<K, 1>(c); S.s16[ig] = (u16)q_up<K, 2>(g);
    return 0u;
}

// ---------------------------------------------------------------------------------------------------
// decoder (qlfc.cpp:1672-1927).  One warp per sub-block, lock-step; runs are expanded warp-wide.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32, 1) q_decode2(const u8 *__restrict__ in_all, SubBlock *__restrict__ sbs, short *__restrict__ cold_all,
                                                   const QTables *__restrict__ tables, u8 *__restrict__ out_all, const u32 *__restrict__ sb_list)
{
    extern __shared__ __align__(16) u8 q_smem_raw[];
    CoderSmem &S = *reinterpret_cast<CoderSmem *>(q_smem_raw);
    coder_smem_init(S, tables);

    const u32 sid = sb_list[blockIdx.x];
    SubBlock &sb = sbs[sid];
    short *cold_s = cold_all + (size_t)blockIdx.x * 2 * COLD_PAD, *cold_c = cold_s + COLD_PAD;
    const u32 lane = threadIdx.x;
    u32 st_cached = 0, st_miss = 0;

    Rc2Dec rc; rc.in = in_all + sb.out_off; rc.pos = 0; rc.limit = sb.out_cap; rc.code = 0; rc.range = 0xffffffffu;
    rc.win = S.inwin; rc.wbase = 0; rc.refill();
    for (int i = 0; i < 3; ++i) rc.code = (rc.code << 16) | rc.get16();
    u32 n = 0; for (int b = 0; b < 32; ++b) n = (n << 1) | rc.decode(2048);
    if (n > sb.in_size) { if (lane == 0) sb.result = LIBBSC_DATA_CORRUPT; return; }   // would overrun the output slice

    int ctxRank0 = 0, ctxRank4 = 0, ctxRun = 0, maxRank = 7, avgRank = 0;
    {
        u32 used8 = 0; int prev = -1;
        for (int d = 0; d < 256; ++d) {
            int c = 0;
            for (int bit = 7; bit >= 0; --bit) {
                bool can0, can1; header_options(used8, prev, c, bit, can0, can1);
                if (can0 && can1) c = 2 * c + (int)rc.decode(2048);
                else if (can1) c = 2 * c + 1;
                else if (can0) c = 2 * c;
            }
            c &= 255;
            S.mtf[d] = (u8)c;
            if (c == prev) { maxRank = ilog2_dev((u32)(d - 1)); break; }
            prev = c; if ((u32)(c >> 3) == lane) used8 |= 1u << (c & 7);
        }
    }
    __syncwarp();

    u8 *out = out_all + sb.in_start;
    for (u32 i = 0; i < n; ) {
        const u32 c = S.mtf[0];
        int rank = 1; u32 b;
        u32 st = S.rank_state[(ctxRun << 11) | (ctxRank4 << 3) | S.rankHist[c]];
        if (avgRank < 32) {
            b = dec3<K_RANK_T>(S, rc, R_RT_STATE + st, R_RT_CHAR + c, R_RT_SHARED);
            if (!b) S.rankHist[c] = 0;
            else {
                u32 e = 1;
                while ((int)e != maxRank) {
                    b = dec3<K_RANK_E>(S, rc, R_RE_STATE + st * 8 + e - 1, R_RE_CHAR + c * 8 + e - 1, R_RE_SHARED + e - 1);
                    if (!b) break;
                    if (++e >= 7) break;                                      // e <= maxRank <= 7 in valid streams
                }
                S.rankHist[c] = (u8)e;
                if (e <= M_MAXE) {
                    const u32 bs = R_RM_STATE + st * M_ROW + (1u << e) - 2u, bc = R_RM_CHAR + c * M_ROW + (1u << e) - 2u, bg = R_WIDE_SHARED + e * 256;
                    for (int bit = (int)e - 1; bit >= 0; --bit) {
                        b = dec3<K_RANK_M>(S, rc, bs + rank, bc + rank, bg + rank);
                        rank = 2 * rank + (int)b;
                    }
                } else {
                    for (int bit = (int)e - 1; bit >= 0; --bit) {
                        const u32 is = cache_get(S, C_STATE_VAL, S.tag_state, cold_s, wide_idx(e, st, rank), st_miss);
                        const u32 ic = cache_get(S, C_CHAR_VAL, S.tag_char, cold_c, wide_idx(e, c, rank), st_miss);
                        st_cached += 2;
                        b = dec3<K_RANK_M>(S, rc, is, ic, R_WIDE_SHARED + e * 256 + rank);
                        rank = 2 * rank + (int)b;
                    }
                }
            }
        } else {
            rank = 0;
            for (int node = 1, bit = maxRank; bit >= 0; --bit) {
                const u32 is = cache_get(S, C_STATE_VAL, S.tag_state, cold_s, wide_idx(8, st, node), st_miss);
                const u32 ic = cache_get(S, C_CHAR_VAL, S.tag_char, cold_c, wide_idx(8, c, node), st_miss);
                st_cached += 2;
                b = dec3<K_RANK_P>(S, rc, is, ic, R_WIDE_SHARED + 8 * 256 + node);
                node = 2 * node + (int)b; rank = 2 * rank + (int)b;
            }
            S.rankHist[c] = (u8)ilog2_dev((u32)rank);
        }
        rank &= 255;
        // push c `rank` places back: mtf[0..rank-1] = mtf[1..rank]; mtf[rank] = c  (qlfc.cpp:1830-1860)
        if (rank >= 1 && rank <= 3) {                        // common case: every lane does the same few moves, no barrier needed
            const u8 m1 = S.mtf[1], m2 = S.mtf[2], m3 = S.mtf[3];
            S.mtf[0] = m1;
            if (rank == 1) S.mtf[1] = (u8)c;
            else { S.mtf[1] = m2; if (rank == 2) S.mtf[2] = (u8)c; else { S.mtf[2] = m3; S.mtf[3] = (u8)c; } }
        } else {
            __syncwarp();
            for (int basep = 0; basep < rank; basep += 32) {
                const int p = basep + (int)lane;
                const u8 v = S.mtf[p + 1];
                __syncwarp();
                if (p < rank) S.mtf[p] = v;
                __syncwarp();
            }
            if (lane == 0) S.mtf[rank] = (u8)c;
            __syncwarp();
        }

        avgRank = (avgRank * 124 + rank * 4) >> 7;
        const int rank0 = rank - 1;
        const int rh = S.runHist[c];
        st = S.run_state[(ctxRank0 << 10) | (ctxRun << 6) | (((u32)rank0 < 7u ? rank0 : 7) << 3) | (rh < 7 ? rh : 7)];
        u32 run = 1;
        b = dec3<K_RUN_T>(S, rc, R_UT_STATE + st, R_UT_CHAR + c, R_UT_SHARED);
        if (!b) S.runHist[c] = (u8)((rh + 2) >> 2);
        else {
            u32 e = 1;
            for (;;) {
                const u32 k = e - 1;
                if (k < UE_RES) b = dec3<K_RUN_E>(S, rc, R_UE_STATE + st * UE_RES + k, R_UE_CHAR + c * UE_RES + k, R_UE_SHARED + k);
                else {
                    const u32 is = cache_get(S, C_STATE_VAL, S.tag_state, cold_s, ue_idx(st, k), st_miss);
                    const u32 ic = cache_get(S, C_CHAR_VAL, S.tag_char, cold_c, ue_idx(c, k), st_miss);
                    st_cached += 2;
                    b = dec3<K_RUN_E>(S, rc, is, ic, R_UE_SHARED + k);
                }
                if (!b) break;
                if (++e >= 31) break;                                         // corrupt-input guard
            }
            S.runHist[c] = (u8)((rh + 3 * e + 3) >> 2);
            if (e <= M_MAXE) {
                const u32 bs = R_UM_STATE + st * M_ROW + (1u << e) - 2u, bc = R_UM_CHAR + c * M_ROW + (1u << e) - 2u, bg = R_NARROW_SHARED + e * 32;
                for (int node = 1, bit = (int)e - 1; bit >= 0; --bit) {
                    b = dec3<K_RUN_M>(S, rc, bs + node, bc + node, bg + node);
                    run = 2 * run + b; node = 2 * node + (int)b;
                }
            } else {
                for (int node = 1, bit = (int)e - 1; bit >= 0; --bit) {
                    const u32 is = cache_get(S, C_STATE_VAL, S.tag_state, cold_s, narrow_idx(e, st, node), st_miss);
                    const u32 ic = cache_get(S, C_CHAR_VAL, S.tag_char, cold_c, narrow_idx(e, c, node), st_miss);
                    st_cached += 2;
                    b = dec3<K_RUN_M>(S, rc, is, ic, R_NARROW_SHARED + e * 32 + node);
                    run = 2 * run + b; node = node + 1;                       // qlfc.cpp:1119: linear contexts above 5 bits
                }
            }
        }
        ctxRank0 = ((ctxRank0 << 1) | (rank0 == 0)) & 0x7;
        ctxRank4 = ((ctxRank4 << 2) | ((u32)rank0 < 3u ? rank0 : 3)) & 0xff;
        ctxRun   = ((ctxRun << 1) | (run < 3)) & 0xf;

        if (run > n - i) run = n - i;                                         // never write past n
        for (u32 k = lane; k < run; k += 32) out[i + k] = (u8)c;
        i += run;
    }
    if (lane == 0) { sb.result = (int)n; sb.stat_cached = st_cached; sb.stat_miss = st_miss; }
}

