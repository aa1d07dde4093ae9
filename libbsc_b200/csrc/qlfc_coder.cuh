// libbsc_b200/csrc/qlfc_coder.cuh -- QLFC stage 2 (context model + binary range coder), device engine.
// Included by qlfc.cu inside its anonymous namespace (needs SubBlock, QTables, c_params, K_*).
//
// Why this looks the way it does.  One coder stream is a strictly serial recurrence (every binary
// decision reads three adaptive counters whose addresses depend on the previous decisions, then
// updates them and the range coder).  The format allows only <= 8 streams per block
// (coder.cpp:52-59), so per-stream LATENCY is everything.  First measurements (profiles/r1a_*)
// showed ~370 cycles per decision with the counters in global memory (every dependent load is an
// L2 round trip because the preceding store invalidates the L1 line).  Here the whole working set
// lives in shared memory of one SM per stream:
//
//   * state tables                       40 KB   (tables.h data, qlfc_tables.inc)
//   * "resident" counters                49 KB   every counter of the first-bit / unary-exponent
//                                                decisions + all model-wide shared counters
//   * two direct-mapped write-back caches 96 KB  for the mantissa/escape banks indexed by state and
//                                                by symbol (1.7 M counters in HBM behind them;
//                                                separate caches so the two lookups of one
//                                                decision can never evict each other)
//
// The warp runs in lock-step: all 32 lanes execute the same decision sequence on the same data
// (shared-memory broadcasts), so no intra-warp synchronisation is needed on the model; lanes only
// diverge to prefetch run records (encoder) and to expand runs / rotate the MTF list (decoder).
#pragma once

// ---- shared-memory counter file (indices in int16 units) -------------------------------------------
constexpr u32 R_RT_SHARED = 0, R_RT_STATE = 2, R_RT_CHAR = R_RT_STATE + 256;
constexpr u32 R_RE_SHARED = R_RT_CHAR + 256, R_RE_STATE = R_RE_SHARED + 8, R_RE_CHAR = R_RE_STATE + 2048;
constexpr u32 R_UT_SHARED = R_RE_CHAR + 2048, R_UT_STATE = R_UT_SHARED + 2, R_UT_CHAR = R_UT_STATE + 256;
constexpr u32 R_UE_SHARED = R_UT_CHAR + 256, R_UE_STATE = R_UE_SHARED + 32, R_UE_CHAR = R_UE_STATE + 8192;
constexpr u32 R_WIDE_SHARED = R_UE_CHAR + 8192;            // 9 banks x 256 (rank mantissa e=0..7, escape)
constexpr u32 R_NARROW_SHARED = R_WIDE_SHARED + 9 * 256;   // 32 banks x 32 (run mantissa)
constexpr u32 R_END = R_NARROW_SHARED + 32 * 32;

constexpr int  QC_LOG = 14;                                 // 16 K entries per cache
constexpr u32  QC_SLOTS = 1u << QC_LOG, QC_MASK = QC_SLOTS - 1;
constexpr u32  C_STATE_VAL = R_END;                         // cache of the by-state banks
constexpr u32  C_CHAR_VAL = C_STATE_VAL + QC_SLOTS;         // cache of the by-symbol banks
constexpr u32  S16_COUNT = C_CHAR_VAL + QC_SLOTS;
// index space behind each cache: 9 wide banks x [256][256], then 32 narrow banks x [256][32]
constexpr u32  COLD_WIDE = 65536, COLD_NARROW = 8192;
constexpr u32  COLD_COUNT = 9 * COLD_WIDE + 32 * COLD_NARROW;          // 851968 counters per kind
constexpr u32  COLD_PAD = (COLD_COUNT + 255) & ~255u;

struct CoderSmem {
    u8    rank_state[32768];
    u8    run_state[8192];
    short s16[S16_COUNT];
    u8    tag_state[QC_SLOTS];
    u8    tag_char[QC_SLOTS];
    u8    rankHist[256], runHist[256];
    u8    mtf[256 + 32];
};

__device__ __forceinline__ void coder_smem_init(CoderSmem &S, const QTables *__restrict__ g)
{
    const u32 lane = threadIdx.x;
    const uint4 *src = (const uint4 *)g; uint4 *dst = (uint4 *)S.rank_state;      // rank_state and run_state are contiguous
    for (u32 i = lane; i < sizeof(QTables) / 16; i += 32) dst[i] = src[i];
    u32 *w = (u32 *)S.s16;
    for (u32 i = lane; i < S16_COUNT / 2; i += 32) w[i] = 0x08000800u;            // every counter starts at 2048
    u32 *t = (u32 *)S.tag_state;
    for (u32 i = lane; i < (2 * QC_SLOTS + 512 + 288) / 4; i += 32) t[i] = 0;     // tags, histories, mtf
    __syncwarp();
}

// Index (into S.s16) of counter `idx` of one kind, loading it through the direct-mapped cache.
__device__ __forceinline__ u32 cache_get(CoderSmem &S, u32 val_base, u8 *tags, short *__restrict__ cold, u32 idx)
{
    const u32 h = idx >> QC_LOG, slot = (idx ^ (h * 1237u)) & QC_MASK, want = h + 1;
    const u32 t = tags[slot];
    if (t != want) {                                         // warp-uniform branch
        if (t) cold[((t - 1) << QC_LOG) | ((slot ^ ((t - 1) * 1237u)) & QC_MASK)] = S.s16[val_base + slot];
        S.s16[val_base + slot] = cold[idx];
        tags[slot] = (u8)want;
    }
    return val_base + slot;
}

// ---- counters ---------------------------------------------------------------------------------------
template <int K> __device__ __forceinline__ int q_mix(int s, int c, int g)
{
    return (c * c_params[K][0] + s * c_params[K][1] + g * c_params[K][2]) >> 5;
}
template <int K, int WHO> __device__ __forceinline__ int q_learn(int p, u32 bit)
{
    const int th0 = c_params[K][3 + 4 * WHO], ar0 = c_params[K][4 + 4 * WHO], th1 = c_params[K][5 + 4 * WHO], ar1 = c_params[K][6 + 4 * WHO];
    const int up = p + (((4096 - th0 - p) * ar0) >> 12), down = p - (((p - th1) * ar1) >> 12);
    return bit ? down : up;
}

// ---- range coder (rangecoder.h:38-271), 16-bit units -------------------------------------------------
struct Rc2Enc {
    u32 low32, carry, range, cache, pending, pos;
    u8 *out;
    __device__ __forceinline__ void put16(u32 v) { *(u16 *)(out + pos) = (u16)v; pos += 2; }   // all lanes store the same value
    __device__ __noinline__ void shift() {
        if (low32 < 0xffff0000u || carry) {
            put16(cache + carry);
            for (; pending; --pending) put16(carry - 1);
            cache = low32 >> 16; carry = 0;
        } else pending++;
        low32 <<= 16;
    }
    __device__ __forceinline__ void encode(u32 bit, int p) {
        if (range < 0x10000u) { shift(); range <<= 16; }
        const u32 r = (range >> 12) * (u32)p;
        if (bit) { const u32 s = low32 + r; carry += (s < low32); low32 = s; range -= r; }
        else range = r;
    }
    __device__ u32 finish() { if (range < 0x10000u) shift(); shift(); shift(); shift(); return pos; }
};

struct Rc2Dec {
    const u8 *in; u32 pos, limit, code, range;
    __device__ __forceinline__ u32 get16() { u32 v = 0; if (pos + 1 < limit) v = (u32)in[pos] | ((u32)in[pos + 1] << 8); pos += 2; return v; }
    __device__ __forceinline__ u32 decode(int p) {
        if (range < 0x10000u) { range <<= 16; code = (code << 16) | get16(); }
        const u32 r = (range >> 12) * (u32)p;
        const u32 bit = code >= r;
        code -= bit ? r : 0u; range = bit ? range - r : r;
        return bit;
    }
};

// one binary decision against three shared-memory counters (indices into S.s16)
template <int K> __device__ __forceinline__ void enc3(CoderSmem &S, Rc2Enc &rc, u32 is, u32 ic, u32 ig, u32 bit)
{
    const int s = S.s16[is], c = S.s16[ic], g = S.s16[ig];
    const int p = q_mix<K>(s, c, g);
    S.s16[is] = (short)q_learn<K, 0>(s, bit); S.s16[ic] = (short)q_learn<K, 1>(c, bit); S.s16[ig] = (short)q_learn<K, 2>(g, bit);
    rc.encode(bit, p);
}
template <int K> __device__ __forceinline__ u32 dec3(CoderSmem &S, Rc2Dec &rc, u32 is, u32 ic, u32 ig)
{
    const int s = S.s16[is], c = S.s16[ic], g = S.s16[ig];
    const u32 bit = rc.decode(q_mix<K>(s, c, g));
    S.s16[is] = (short)q_learn<K, 0>(s, bit); S.s16[ic] = (short)q_learn<K, 1>(c, bit); S.s16[ig] = (short)q_learn<K, 2>(g, bit);
    return bit;
}

// index helpers for the cached banks.  wide bank b = 0..7 rank mantissa by exponent, 8 = escape.
__device__ __forceinline__ u32 wide_idx(u32 bank, u32 x, u32 node) { return bank * COLD_WIDE + x * 256 + node; }
__device__ __forceinline__ u32 narrow_idx(u32 e, u32 x, u32 node) { return 9 * COLD_WIDE + e * COLD_NARROW + x * 32 + node; }

// ---------------------------------------------------------------------------------------------------
// encoder (qlfc.cpp:829-1129)
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32, 1) q_encode2(const u32 *__restrict__ run_pos, const u8 *__restrict__ run_sym, const u8 *__restrict__ run_rank,
                                                   SubBlock *__restrict__ sbs, const u8 *__restrict__ mtf_all, short *__restrict__ cold_all,
                                                   const QTables *__restrict__ tables, u8 *__restrict__ out_all, const u32 *__restrict__ sb_list)
{
    extern __shared__ __align__(16) u8 q_smem_raw[];
    CoderSmem &S = *reinterpret_cast<CoderSmem *>(q_smem_raw);
    coder_smem_init(S, tables);

    const u32 sid = sb_list ? sb_list[blockIdx.x] : blockIdx.x;
    SubBlock &sb = sbs[sid];
    short *cold_s = cold_all + (size_t)sid * 2 * COLD_PAD, *cold_c = cold_s + COLD_PAD;
    const u32 lane = threadIdx.x;

    Rc2Enc rc; rc.low32 = 0; rc.carry = 0; rc.range = 0xffffffffu; rc.cache = 0; rc.pending = 0; rc.pos = 0; rc.out = out_all + sb.out_off;
    const long long eob = (long long)sb.out_cap - 16;
    int ctxRank0 = 0, ctxRank4 = 0, ctxRun = 0, maxRank = 7, avgRank = 0;

    const u32 n = sb.in_size;
    for (int b = 31; b >= 0; --b) rc.encode((n >> b) & 1u, 2048);
    {   // MTF-order header (qlfc.cpp:857-891)
        const u8 *mtf = mtf_all + sid * 256;
        u32 used8 = 0; int prev = -1;
        for (int d = 0; d < 256; ++d) {
            int c = mtf[d];
            for (int bit = 7; bit >= 0; --bit) {
                bool can0, can1; header_options(used8, prev, c >> (bit + 1), bit, can0, can1);
                if (can0 && can1) rc.encode((c >> bit) & 1u, 2048);
            }
            if (c == prev) { maxRank = ilog2_dev((u32)(d - 1)); break; }
            prev = c; if ((u32)(c >> 3) == lane) used8 |= 1u << (c & 7);
        }
    }

    int result = 0;
    const u32 rb = sb.run_begin, re = sb.run_end;
    for (u32 t0 = rb; t0 < re && result == 0; t0 += 32) {
        const u32 cnt = min(32u, re - t0);
        u32 my_sym = 0, my_rank = 0, my_len = 0;             // lane j prefetches run t0 + j
        if (lane < cnt) { my_sym = run_sym[t0 + lane]; my_rank = run_rank[t0 + lane]; my_len = run_pos[t0 + lane + 1] - run_pos[t0 + lane]; }
        for (u32 j = 0; j < cnt; ++j) {
            if ((long long)rc.pos >= eob) { result = LIBBSC_NOT_COMPRESSIBLE; break; }   // qlfc.cpp:898-901
            const u32 c = __shfl_sync(0xffffffffu, my_sym, j);
            const int rank = (int)__shfl_sync(0xffffffffu, my_rank, j);
            const int run = (int)__shfl_sync(0xffffffffu, my_len, j);

            u32 st = S.rank_state[(ctxRun << 11) | (ctxRank4 << 3) | S.rankHist[c]];
            if (avgRank < 32) {
                enc3<K_RANK_T>(S, rc, R_RT_STATE + st, R_RT_CHAR + c, R_RT_SHARED, rank != 1);
                if (rank == 1) S.rankHist[c] = 0;
                else {
                    const int e = ilog2_dev((u32)rank);
                    S.rankHist[c] = (u8)e;
                    for (int b = 1; b < e; ++b) enc3<K_RANK_E>(S, rc, R_RE_STATE + st * 8 + b - 1, R_RE_CHAR + c * 8 + b - 1, R_RE_SHARED + b - 1, 1);
                    if (e < maxRank)          enc3<K_RANK_E>(S, rc, R_RE_STATE + st * 8 + e - 1, R_RE_CHAR + c * 8 + e - 1, R_RE_SHARED + e - 1, 0);
                    for (int node = 1, bit = e - 1; bit >= 0; --bit) {
                        const u32 bb = ((u32)rank >> bit) & 1u;
                        const u32 is = cache_get(S, C_STATE_VAL, S.tag_state, cold_s, wide_idx(e, st, node));
                        const u32 ic = cache_get(S, C_CHAR_VAL, S.tag_char, cold_c, wide_idx(e, c, node));
                        enc3<K_RANK_M>(S, rc, is, ic, R_WIDE_SHARED + e * 256 + node, bb);
                        node = 2 * node + (int)bb;
                    }
                }
            } else {
                S.rankHist[c] = (u8)ilog2_dev((u32)rank);
                for (int node = 1, bit = maxRank; bit >= 0; --bit) {
                    const u32 bb = ((u32)rank >> bit) & 1u;
                    const u32 is = cache_get(S, C_STATE_VAL, S.tag_state, cold_s, wide_idx(8, st, node));
                    const u32 ic = cache_get(S, C_CHAR_VAL, S.tag_char, cold_c, wide_idx(8, c, node));
                    enc3<K_RANK_P>(S, rc, is, ic, R_WIDE_SHARED + 8 * 256 + node, bb);
                    node = 2 * node + (int)bb;
                }
            }
            avgRank = (avgRank * 124 + rank * 4) >> 7;
            const int rank0 = rank - 1;
            const int rh = S.runHist[c];
            st = S.run_state[(ctxRank0 << 10) | (ctxRun << 6) | ((rank0 < 7 ? rank0 : 7) << 3) | (rh < 7 ? rh : 7)];

            enc3<K_RUN_T>(S, rc, R_UT_STATE + st, R_UT_CHAR + c, R_UT_SHARED, run != 1);
            if (run == 1) S.runHist[c] = (u8)((rh + 2) >> 2);
            else {
                const int e = ilog2_dev((u32)run);
                S.runHist[c] = (u8)((rh + 3 * e + 3) >> 2);
                for (int b = 1; b < e; ++b) enc3<K_RUN_E>(S, rc, R_UE_STATE + st * 32 + b - 1, R_UE_CHAR + c * 32 + b - 1, R_UE_SHARED + b - 1, 1);
                enc3<K_RUN_E>(S, rc, R_UE_STATE + st * 32 + e - 1, R_UE_CHAR + c * 32 + e - 1, R_UE_SHARED + e - 1, 0);
                for (int node = 1, bit = e - 1; bit >= 0; --bit) {
                    const u32 bb = ((u32)run >> bit) & 1u;
                    const u32 is = cache_get(S, C_STATE_VAL, S.tag_state, cold_s, narrow_idx(e, st, node));
                    const u32 ic = cache_get(S, C_CHAR_VAL, S.tag_char, cold_c, narrow_idx(e, c, node));
                    enc3<K_RUN_M>(S, rc, is, ic, R_NARROW_SHARED + e * 32 + node, bb);
                    node = (e <= 5) ? 2 * node + (int)bb : node + 1;          // qlfc.cpp:1119
                }
            }
            ctxRank0 = ((ctxRank0 << 1) | (rank0 == 0)) & 0x7;
            ctxRank4 = ((ctxRank4 << 2) | (rank0 < 3 ? rank0 : 3)) & 0xff;
            ctxRun   = ((ctxRun << 1) | (run < 3)) & 0xf;
        }
    }
    if (result == 0) result = (int)rc.finish();
    if (lane == 0) sb.result = result;
}

// ---------------------------------------------------------------------------------------------------
// decoder (qlfc.cpp:1672-1927)
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

    Rc2Dec rc; rc.in = in_all + sb.out_off; rc.pos = 0; rc.limit = sb.out_cap; rc.code = 0; rc.range = 0xffffffffu;
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
                int e = 1;
                while (e != maxRank) {
                    b = dec3<K_RANK_E>(S, rc, R_RE_STATE + st * 8 + e - 1, R_RE_CHAR + c * 8 + e - 1, R_RE_SHARED + e - 1);
                    if (!b) break;
                    if (++e >= 7) break;                                      // e <= maxRank <= 7 in valid streams
                }
                S.rankHist[c] = (u8)e;
                for (int bit = e - 1; bit >= 0; --bit) {
                    const u32 is = cache_get(S, C_STATE_VAL, S.tag_state, cold_s, wide_idx(e, st, rank));
                    const u32 ic = cache_get(S, C_CHAR_VAL, S.tag_char, cold_c, wide_idx(e, c, rank));
                    b = dec3<K_RANK_M>(S, rc, is, ic, R_WIDE_SHARED + e * 256 + rank);
                    rank = 2 * rank + (int)b;
                }
            }
        } else {
            rank = 0;
            for (int node = 1, bit = maxRank; bit >= 0; --bit) {
                const u32 is = cache_get(S, C_STATE_VAL, S.tag_state, cold_s, wide_idx(8, st, node));
                const u32 ic = cache_get(S, C_CHAR_VAL, S.tag_char, cold_c, wide_idx(8, c, node));
                b = dec3<K_RANK_P>(S, rc, is, ic, R_WIDE_SHARED + 8 * 256 + node);
                node = 2 * node + (int)b; rank = 2 * rank + (int)b;
            }
            S.rankHist[c] = (u8)ilog2_dev((u32)rank);
        }
        rank &= 255;
        __syncwarp();
        // push c `rank` places back: mtf[0..rank-1] = mtf[1..rank]; mtf[rank] = c  (qlfc.cpp:1830-1860)
        for (int basep = 0; basep < rank; basep += 32) {
            const int p = basep + (int)lane;
            const u8 v = S.mtf[p + 1];
            __syncwarp();
            if (p < rank) S.mtf[p] = v;
            __syncwarp();
        }
        if (lane == 0) S.mtf[rank] = (u8)c;
        __syncwarp();

        avgRank = (avgRank * 124 + rank * 4) >> 7;
        const int rank0 = rank - 1;
        const int rh = S.runHist[c];
        st = S.run_state[(ctxRank0 << 10) | (ctxRun << 6) | (((u32)rank0 < 7u ? rank0 : 7) << 3) | (rh < 7 ? rh : 7)];
        u32 run = 1;
        b = dec3<K_RUN_T>(S, rc, R_UT_STATE + st, R_UT_CHAR + c, R_UT_SHARED);
        if (!b) S.runHist[c] = (u8)((rh + 2) >> 2);
        else {
            int e = 1;
            for (;;) {
                b = dec3<K_RUN_E>(S, rc, R_UE_STATE + st * 32 + e - 1, R_UE_CHAR + c * 32 + e - 1, R_UE_SHARED + e - 1);
                if (!b) break;
                if (++e >= 31) break;                                         // corrupt-input guard
            }
            S.runHist[c] = (u8)((rh + 3 * e + 3) >> 2);
            for (int node = 1, bit = e - 1; bit >= 0; --bit) {
                const u32 is = cache_get(S, C_STATE_VAL, S.tag_state, cold_s, narrow_idx(e, st, node));
                const u32 ic = cache_get(S, C_CHAR_VAL, S.tag_char, cold_c, narrow_idx(e, c, node));
                b = dec3<K_RUN_M>(S, rc, is, ic, R_NARROW_SHARED + e * 32 + node);
                run = 2 * run + b;
                node = (e <= 5) ? 2 * node + (int)b : node + 1;
            }
        }
        ctxRank0 = ((ctxRank0 << 1) | (rank0 == 0)) & 0x7;
        ctxRank4 = ((ctxRank4 << 2) | ((u32)rank0 < 3u ? rank0 : 3)) & 0xff;
        ctxRun   = ((ctxRun << 1) | (run < 3)) & 0xf;

        if (run > n - i) run = n - i;                                         // never write past n
        for (u32 k = lane; k < run; k += 32) out[i + k] = (u8)c;
        i += run;
    }
    if (lane == 0) sb.result = (int)n;
}
