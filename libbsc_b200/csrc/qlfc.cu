// libbsc_b200/csrc/qlfc.cu -- QLFC entropy stage (static model, coder id 1) on the device.
//
// Replaces bsc_coder_compress / bsc_coder_decompress (libbsc/coder/coder.cpp:244-347) and the
// static QLFC coder under them (libbsc/coder/qlfc/qlfc.cpp:200-255 transform, 829-1129 encoder,
// 1672-1927 decoder; coder/common/rangecoder.h; coder/common/predictor.h:45-61;
// qlfc_model.h:115-241).  The bit stream is reproduced exactly; the data that define it (two
// state tables + the F_* constants) come from qlfc_tables.inc.
//
// Device decomposition of the encoder:
//   q_split      sub-block boundaries of the container (coder.cpp:70-109), one CTA.
//   q_run_*      parallel run detection over the whole block: run start positions + symbols.
//   q_ranks      stage 1 of QLFC (backward move-to-front rank per run), one warp per sub-block,
//                the 256-entry recency list packed 8 symbols per lane and updated with
//                byte-SIMD compares; also emits the MTF-order table.
//   q_encode     stage 2: context model + binary range coder.  The format fixes <= 8 independent
//                streams per block and each stream is a serial recurrence (adaptive counters +
//                range/low), so this kernel runs one warp per sub-block in lock-step (all lanes
//                execute the same decisions; lanes are used for prefetch/broadcast of run
//                records).  Throughput comes from running many blocks' streams concurrently.
//   q_decode     inverse of both stages, one warp per sub-block, warp-wide run expansion.
#include "common.cuh"
#include "stages.cuh"
#include "qlfc_tables.inc"

#define Q_MAX_SUB 8

namespace {

// ---------------------------------------------------------------------------------------------
// counter layout (int16, all start at 2048) -- our own flat layout
// ---------------------------------------------------------------------------------------------
constexpr u32 WIDE  = 256 + 65536 + 65536;               // shared[256] | by_state[256][256] | by_char[256][256]
constexpr u32 NARROW = 32 + 8192 + 8192;                 // shared[32]  | by_state[256][32]  | by_char[256][32]
constexpr u32 O_RT_SHARED = 0, O_RT_STATE = 2, O_RT_CHAR = O_RT_STATE + 256;
constexpr u32 O_RE_SHARED = O_RT_CHAR + 256, O_RE_STATE = O_RE_SHARED + 8, O_RE_CHAR = O_RE_STATE + 2048;
constexpr u32 O_RM = O_RE_CHAR + 2048;                   // 8 wide banks (mantissa by exponent)
constexpr u32 O_RP = O_RM + 8 * WIDE;                    // escape bank
constexpr u32 O_UT_SHARED = O_RP + WIDE, O_UT_STATE = O_UT_SHARED + 2, O_UT_CHAR = O_UT_STATE + 256;
constexpr u32 O_UE = O_UT_CHAR + 256;                    // narrow bank (run exponent)
constexpr u32 O_UM = O_UE + NARROW;                      // 32 narrow banks (run mantissa by exponent)
constexpr u32 MODEL_SHORTS = O_UM + 32 * NARROW;
constexpr u32 MODEL_SHORTS_PAD = (MODEL_SHORTS + 127) & ~127u;

enum { K_RANK_T, K_RANK_E, K_RANK_M, K_RANK_P, K_RUN_T, K_RUN_E, K_RUN_M };

__constant__ short c_params[7][15];

struct SubBlock {
    u32 in_start, in_size;       // slice of the block
    u32 run_begin, run_end;      // slice of the run arrays
    u32 out_off, out_cap;        // slice of the temp output / (decode) input stream offset, size
    int result;                  // bytes produced or error
    u32 nsym;
};

// ---------------------------------------------------------------------------------------------
// sub-block split (coder.cpp:70-109): sample i = 1, 33, 65, ...; cut after every (total/nBlocks)
// sampled changes.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) q_split(const u8 *__restrict__ in, u32 n, u32 nBlocks, u32 *__restrict__ start, u32 *__restrict__ size)
{
    __shared__ u32 s_warp[32];
    __shared__ u32 s_total;
    __shared__ u32 s_cut[Q_MAX_SUB];
    const u32 samples = n > 1 ? (n - 2) / 32 + 1 : 0;    // i = 1 + 32*j < n
    const u32 per_thread = (samples + 1023) / 1024;
    const u32 b = threadIdx.x * per_thread, e = min(b + per_thread, samples);
    u32 cnt = 0;
    for (u32 j = b; j < e; ++j) { u32 i = 1 + 32 * j; cnt += (in[i] != in[i - 1]); }
    u32 lane = threadIdx.x & 31, w = threadIdx.x >> 5, incl = cnt;
    for (int o = 1; o < 32; o <<= 1) { u32 t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (u32)o) incl += t; }
    if (lane == 31) s_warp[w] = incl;
    if (threadIdx.x < Q_MAX_SUB) s_cut[threadIdx.x] = 0;
    __syncthreads();
    u32 pre = 0; for (u32 i = 0; i < w; ++i) pre += s_warp[i];
    u32 excl = pre + incl - cnt;
    if (threadIdx.x == 1023) s_total = pre + incl;
    __syncthreads();
    const u32 total = s_total;
    if (total > nBlocks) {
        const u32 per = total / nBlocks;
        // cut id (1..nBlocks-1) happens at the sample where the running change count reaches id*per
        u32 run = excl;
        for (u32 j = b; j < e; ++j) {
            u32 i = 1 + 32 * j;
            if (in[i] != in[i - 1]) { ++run; if (run % per == 0) { u32 id = run / per; if (id < nBlocks) s_cut[id] = i; } }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            u32 prev = 0;
            for (u32 id = 0; id < nBlocks; ++id) {
                u32 nxt = (id + 1 < nBlocks) ? s_cut[id + 1] : n;
                start[id] = prev; size[id] = nxt - prev; prev = nxt;
            }
        }
    } else if (threadIdx.x == 0) {
        for (u32 p = 0; p < nBlocks; ++p) { start[p] = (n / nBlocks) * p; size[p] = (p != nBlocks - 1) ? n / nBlocks : n - (n / nBlocks) * (nBlocks - 1); }
    }
}

// ---------------------------------------------------------------------------------------------
// run detection: a run starts where the byte changes or a sub-block starts.
// ---------------------------------------------------------------------------------------------
#define RUN_THREADS 256
#define RUN_ITEMS   16
#define RUN_TILE    (RUN_THREADS * RUN_ITEMS)

__device__ __forceinline__ bool is_run_head(const u8 *__restrict__ in, u32 i, const u32 *s_start, u32 nBlocks)
{
    if (i == 0 || in[i] != in[i - 1]) return true;
    for (u32 s = 1; s < nBlocks; ++s) if (s_start[s] == i) return true;
    return false;
}

__global__ void __launch_bounds__(RUN_THREADS) q_run_count(const u8 *__restrict__ in, u32 n, const u32 *__restrict__ sb_start, u32 nBlocks, u32 *__restrict__ tile_count)
{
    __shared__ u32 s_w[RUN_THREADS / 32];
    __shared__ u32 s_start[Q_MAX_SUB];
    if (threadIdx.x < nBlocks) s_start[threadIdx.x] = sb_start[threadIdx.x];
    __syncthreads();
    u32 base = blockIdx.x * RUN_TILE + threadIdx.x * RUN_ITEMS, c = 0;
    for (int k = 0; k < RUN_ITEMS; ++k) { u32 i = base + k; if (i < n) c += is_run_head(in, i, s_start, nBlocks); }
    for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) { u32 t = 0; for (int i = 0; i < RUN_THREADS / 32; ++i) t += s_w[i]; tile_count[blockIdx.x] = t; }
}

// exclusive scan of tile counts in place (single CTA); total -> *total_out
__global__ void __launch_bounds__(1024) q_scan_tiles(u32 *__restrict__ tile_count, u32 tiles, u32 *__restrict__ total_out)
{
    __shared__ u32 s_w[32];
    u32 per = (tiles + 1023) / 1024, b = threadIdx.x * per, e = min(b + per, tiles), c = 0;
    for (u32 i = b; i < e; ++i) c += tile_count[i];
    u32 lane = threadIdx.x & 31, w = threadIdx.x >> 5, incl = c;
    for (int o = 1; o < 32; o <<= 1) { u32 t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (u32)o) incl += t; }
    if (lane == 31) s_w[w] = incl;
    __syncthreads();
    u32 pre = 0; for (u32 i = 0; i < w; ++i) pre += s_w[i];
    u32 run = pre + incl - c;
    for (u32 i = b; i < e; ++i) { u32 t = tile_count[i]; tile_count[i] = run; run += t; }
    if (threadIdx.x == 1023) *total_out = pre + incl;
}

__global__ void __launch_bounds__(RUN_THREADS) q_run_write(const u8 *__restrict__ in, u32 n, const u32 *__restrict__ sb_start, u32 nBlocks,
                                                           const u32 *__restrict__ tile_excl, u32 *__restrict__ run_pos, u8 *__restrict__ run_sym)
{
    __shared__ u32 s_w[RUN_THREADS / 32];
    __shared__ u32 s_start[Q_MAX_SUB];
    if (threadIdx.x < nBlocks) s_start[threadIdx.x] = sb_start[threadIdx.x];
    __syncthreads();
    u32 base = blockIdx.x * RUN_TILE + threadIdx.x * RUN_ITEMS, c = 0;
    bool hd[RUN_ITEMS];
    for (int k = 0; k < RUN_ITEMS; ++k) { u32 i = base + k; hd[k] = (i < n) && is_run_head(in, i, s_start, nBlocks); c += hd[k]; }
    u32 lane = threadIdx.x & 31, w = threadIdx.x >> 5, incl = c;
    for (int o = 1; o < 32; o <<= 1) { u32 t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (u32)o) incl += t; }
    if (lane == 31) s_w[w] = incl;
    __syncthreads();
    u32 pre = tile_excl[blockIdx.x]; for (u32 i = 0; i < w; ++i) pre += s_w[i];
    u32 o = pre + incl - c;
    for (int k = 0; k < RUN_ITEMS; ++k) if (hd[k]) { run_pos[o] = base + k; run_sym[o] = in[base + k]; ++o; }
}

// run index of each sub-block's first run (binary search over run_pos) + sentinel
__global__ void q_run_bounds(u32 *run_pos, u32 R, u32 n, const u32 *__restrict__ sb_start, const u32 *__restrict__ sb_size, u32 nBlocks,
                             SubBlock *__restrict__ sb)
{
    u32 s = threadIdx.x;
    if (s == 0) run_pos[R] = n;                          // sentinel: run t spans [run_pos[t], run_pos[t+1])
    if (s >= nBlocks) return;
    u32 lo = 0, hi = R;                                  // first run with pos >= sb_start[s]
    while (lo < hi) { u32 mid = (lo + hi) >> 1; if (run_pos[mid] < sb_start[s]) lo = mid + 1; else hi = mid; }
    sb[s].in_start = sb_start[s]; sb[s].in_size = sb_size[s]; sb[s].run_begin = lo;
    u32 endpos = sb_start[s] + sb_size[s];
    u32 lo2 = lo, hi2 = R;
    while (lo2 < hi2) { u32 mid = (lo2 + hi2) >> 1; if (run_pos[mid] < endpos) lo2 = mid + 1; else hi2 = mid; }
    sb[s].run_end = lo2;
}

// ---------------------------------------------------------------------------------------------
// QLFC stage 1 (qlfc.cpp:200-255 / 398-455): backward MTF rank per run.  One warp per sub-block.
// Lane l keeps the list positions of symbols 8l..8l+7 as bytes of (lo, hi).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) q_ranks(const u8 *__restrict__ run_sym, u8 *__restrict__ run_rank, SubBlock *__restrict__ sbs, u8 *__restrict__ mtf_out)
{
    SubBlock &sb = sbs[blockIdx.x];
    const u32 lane = threadIdx.x;
    const u32 rb = sb.run_begin, re = sb.run_end;
    // initial order: identity, except symbols 0 and 1 swapped when the last byte is 0 (qlfc.cpp:405-408)
    u32 lo = (8 * lane) | ((8 * lane + 1) << 8) | ((8 * lane + 2) << 16) | ((8 * lane + 3) << 24);
    u32 hi = (8 * lane + 4) | ((8 * lane + 5) << 8) | ((8 * lane + 6) << 16) | ((8 * lane + 7) << 24);
    if (re > rb && run_sym[re - 1] == 0 && lane == 0) lo = (lo & 0xffff0000u) | 0x0001u;   // pos[0]=1, pos[1]=0
    u32 seen = 0, nsym = 0;

    for (u32 hiR = re; hiR > rb; ) {
        u32 cnt = min(32u, hiR - rb);
        // lane j handles run (hiR-1-j)
        u32 mysym = lane < cnt ? run_sym[hiR - 1 - lane] : 0, myrank = 0;
        for (u32 j = 0; j < cnt; ++j) {
            u32 c = __shfl_sync(0xffffffffu, mysym, j);
            u32 word = (c & 4) ? hi : lo;
            u32 wsel = __shfl_sync(0xffffffffu, word, c >> 3);
            u32 sm = __shfl_sync(0xffffffffu, seen, c >> 3);
            u32 r = (wsel >> ((c & 3) * 8)) & 255u;
            u32 r4 = r * 0x01010101u;
            lo = __vsub4(lo, __vcmpltu4(lo, r4));        // positions < r move one place back
            hi = __vsub4(hi, __vcmpltu4(hi, r4));
            if (lane == (c >> 3)) {
                u32 clr = ~(255u << ((c & 3) * 8));
                if (c & 4) hi &= clr; else lo &= clr;    // c goes to the front
                seen |= 1u << (c & 7);
            }
            u32 out = r;
            if (!((sm >> (c & 7)) & 1u)) out = nsym++;   // last occurrence: ordinal from the end
            if (lane == j) myrank = out;
        }
        if (lane < cnt) run_rank[hiR - 1 - lane] = (u8)myrank;
        hiR -= cnt;
    }
    __syncwarp();
    if (lane == 0 && re > rb) run_rank[re - 1] = 1;       // qlfc.cpp:249

    // MTF-order table: mtf[pos[c]] = c, then duplicate-terminate after the used symbols
    __shared__ u8 s_mtf[256];
    __shared__ u8 s_seen[256];
    for (int k = 0; k < 8; ++k) {
        u32 c = 8 * lane + k, p = ((k & 4 ? hi : lo) >> ((k & 3) * 8)) & 255u;
        s_mtf[p] = (u8)c; s_seen[c] = (seen >> k) & 1u;
    }
    __syncwarp();
    if (lane == 0) {
        for (int d = 1; d < 256; ++d) if (!s_seen[s_mtf[d]]) { s_mtf[d] = s_mtf[d - 1]; break; }
        sb.nsym = nsym;
    }
    __syncwarp();
    for (int k = 0; k < 8; ++k) mtf_out[blockIdx.x * 256 + 8 * lane + k] = s_mtf[8 * lane + k];
}

// ---------------------------------------------------------------------------------------------
// model helpers
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) q_model_init(short *__restrict__ models, size_t total_shorts)
{
    size_t i = ((size_t)blockIdx.x * 256 + threadIdx.x) * 8;
    if (i + 8 <= total_shorts) *(uint4 *)(models + i) = make_uint4(0x08000800u, 0x08000800u, 0x08000800u, 0x08000800u);
    else for (; i < total_shorts; ++i) models[i] = 2048;
}

template <int K> __device__ __forceinline__ int mix3(int s, int c, int g)
{
    return (c * c_params[K][0] + s * c_params[K][1] + g * c_params[K][2]) >> 5;
}
template <int K, int WHO> __device__ __forceinline__ int learn(int p, u32 bit)
{
    const short *q = &c_params[K][3 + 4 * WHO];
    return bit ? p - (((p - q[2]) * q[3]) >> 12) : p + (((4096 - q[0] - p) * q[1]) >> 12);
}

struct Counters3 { short *s, *c, *g; };

// ---------------------------------------------------------------------------------------------
// range coder (rangecoder.h:38-271), 16-bit output units
// ---------------------------------------------------------------------------------------------
struct RcEnc {
    u32 low32, carry, range, cache, pending, pos;
    u8 *out;
    __device__ __forceinline__ void put16(u32 v) { if (lane_id() == 0) *(u16 *)(out + pos) = (u16)v; pos += 2; }
    __device__ void shift() {
        if (low32 < 0xffff0000u || carry) {
            put16(cache + carry);
            for (; pending; --pending) put16(carry - 1);
            cache = low32 >> 16; carry = 0;
        } else pending++;
        low32 <<= 16;
    }
    __device__ __forceinline__ void encode(u32 bit, int p) {
        if (range < 0x10000u) { shift(); range <<= 16; }
        u32 r = (range >> 12) * (u32)p;
        if (bit) { u32 s = low32 + r; carry += (s < low32); low32 = s; range -= r; }
        else range = r;
    }
    __device__ u32 finish() {
        if (range < 0x10000u) shift();
        shift(); shift(); shift();
        return pos;
    }
};

struct RcDec {
    const u8 *in; u32 pos, limit, code, range;
    __device__ __forceinline__ u32 get16() { u32 v = 0; if (pos + 1 < limit) v = (u32)in[pos] | ((u32)in[pos + 1] << 8); pos += 2; return v; }
    __device__ __forceinline__ u32 decode(int p) {
        if (range < 0x10000u) { range <<= 16; code = (code << 16) | get16(); }
        u32 r = (range >> 12) * (u32)p;
        u32 bit = code >= r;
        if (bit) { code -= r; range -= r; } else range = r;
        return bit;
    }
};

// encode/decode one binary decision against three counters
template <int K> __device__ __forceinline__ void enc_decision(RcEnc &rc, short *ps, short *pc, short *pg, u32 bit)
{
    int s = *ps, c = *pc, g = *pg;
    int p = mix3<K>(s, c, g);
    s = learn<K, 0>(s, bit); c = learn<K, 1>(c, bit); g = learn<K, 2>(g, bit);
    if (lane_id() == 0) { *ps = (short)s; *pc = (short)c; *pg = (short)g; }
    __syncwarp();
    rc.encode(bit, p);
}
template <int K> __device__ __forceinline__ u32 dec_decision(RcDec &rc, short *ps, short *pc, short *pg)
{
    int s = *ps, c = *pc, g = *pg;
    u32 bit = rc.decode(mix3<K>(s, c, g));
    s = learn<K, 0>(s, bit); c = learn<K, 1>(c, bit); g = learn<K, 2>(g, bit);
    if (lane_id() == 0) { *ps = (short)s; *pc = (short)c; *pg = (short)g; }
    __syncwarp();
    return bit;
}

struct RunCtx {
    int ctxRank0, ctxRank4, ctxRun, maxRank, avgRank;
};

__device__ __forceinline__ int ilog2_dev(u32 v) { return 31 - __clz(v | 1u); }

// which symbols can still appear in the MTF-order header (qlfc.cpp:857-891): lane l owns symbols 8l..8l+7
__device__ __forceinline__ void header_options(u32 used8, int prev, int prefix, int bit, bool &can0, bool &can1)
{
    u32 lane = lane_id(); bool c0 = false, c1 = false;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        int c = 8 * (int)lane + k;
        bool cand = (c == prev || !((used8 >> k) & 1u)) && ((c >> (bit + 1)) == prefix);
        if (cand) { if (c & (1 << bit)) c1 = true; else c0 = true; }
    }
    can0 = __any_sync(0xffffffffu, c0); can1 = __any_sync(0xffffffffu, c1);
}

struct QTables { u8 rank_state[32768]; u8 run_state[8192]; };

__device__ __forceinline__ void load_tables(u8 *s_tab, const QTables *__restrict__ g)
{
    const uint4 *src = (const uint4 *)g; uint4 *dst = (uint4 *)s_tab;
    for (u32 i = threadIdx.x; i < sizeof(QTables) / 16; i += blockDim.x) dst[i] = src[i];
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// QLFC stage 2 encoder (qlfc.cpp:829-1129).  One warp per sub-block, lock-step.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) q_encode(const u32 *__restrict__ run_pos, const u8 *__restrict__ run_sym, const u8 *__restrict__ run_rank,
                                               SubBlock *__restrict__ sbs, const u8 *__restrict__ mtf_all, short *__restrict__ models,
                                               const QTables *__restrict__ tables, u8 *__restrict__ out_all, const u32 *__restrict__ sb_list)
{
    extern __shared__ __align__(16) u8 s_tab[];
    load_tables(s_tab, tables);
    const u8 *t_rank = s_tab, *t_run = s_tab + 32768;
    __shared__ u8 s_rankHist[256], s_runHist[256];

    const u32 sid = sb_list ? sb_list[blockIdx.x] : blockIdx.x;
    SubBlock &sb = sbs[sid];
    short *M = models + (size_t)sid * MODEL_SHORTS_PAD;
    const u32 lane = threadIdx.x;
    for (int i = lane; i < 256; i += 32) { s_rankHist[i] = 0; s_runHist[i] = 0; }
    __syncwarp();

    RcEnc rc; rc.low32 = 0; rc.carry = 0; rc.range = 0xffffffffu; rc.cache = 0; rc.pending = 0; rc.pos = 0; rc.out = out_all + sb.out_off;
    const long long eob = (long long)sb.out_cap - 16;
    RunCtx x; x.ctxRank0 = x.ctxRank4 = x.ctxRun = 0; x.maxRank = 7; x.avgRank = 0;

    const u32 n = sb.in_size;
    for (int b = 31; b >= 0; --b) rc.encode((n >> b) & 1u, 2048);

    {   // MTF-order header
        const u8 *mtf = mtf_all + sid * 256;
        u32 used8 = 0; int prev = -1;
        for (int d = 0; d < 256; ++d) {
            int c = mtf[d];
            for (int bit = 7; bit >= 0; --bit) {
                bool can0, can1; header_options(used8, prev, c >> (bit + 1), bit, can0, can1);
                if (can0 && can1) rc.encode((c >> bit) & 1u, 2048);
            }
            if (c == prev) { x.maxRank = ilog2_dev((u32)(d - 1)); break; }
            prev = c; if ((u32)(c >> 3) == lane) used8 |= 1u << (c & 7);
        }
    }

    int result = 0;
    const u32 rb = sb.run_begin, re = sb.run_end;
    for (u32 t0 = rb; t0 < re && result == 0; t0 += 32) {
        const u32 cnt = min(32u, re - t0);
        // lane j prefetches run t0+j
        u32 my_sym = 0, my_rank = 0, my_len = 0;
        if (lane < cnt) { my_sym = run_sym[t0 + lane]; my_rank = run_rank[t0 + lane]; my_len = run_pos[t0 + lane + 1] - run_pos[t0 + lane]; }
        for (u32 j = 0; j < cnt; ++j) {
            if ((long long)rc.pos >= eob) { result = LIBBSC_NOT_COMPRESSIBLE; break; }   // qlfc.cpp:898-901
            const int c = (int)__shfl_sync(0xffffffffu, my_sym, j);
            const int rank = (int)__shfl_sync(0xffffffffu, my_rank, j);
            const int run = (int)__shfl_sync(0xffffffffu, my_len, j);

            int st = t_rank[(x.ctxRun << 11) | (x.ctxRank4 << 3) | s_rankHist[c]];
            if (x.avgRank < 32) {
                enc_decision<K_RANK_T>(rc, M + O_RT_STATE + st, M + O_RT_CHAR + c, M + O_RT_SHARED, rank != 1);
                if (rank == 1) { if (lane == 0) s_rankHist[c] = 0; }
                else {
                    const int e = ilog2_dev((u32)rank);
                    if (lane == 0) s_rankHist[c] = (u8)e;
                    for (int b = 1; b < e; ++b) enc_decision<K_RANK_E>(rc, M + O_RE_STATE + st * 8 + b - 1, M + O_RE_CHAR + c * 8 + b - 1, M + O_RE_SHARED + b - 1, 1);
                    if (e < x.maxRank)          enc_decision<K_RANK_E>(rc, M + O_RE_STATE + st * 8 + e - 1, M + O_RE_CHAR + c * 8 + e - 1, M + O_RE_SHARED + e - 1, 0);
                    short *bank = M + O_RM + (u32)e * WIDE;
                    for (int node = 1, bit = e - 1; bit >= 0; --bit) {
                        u32 bb = ((u32)rank >> bit) & 1u;
                        enc_decision<K_RANK_M>(rc, bank + 256 + st * 256 + node, bank + 256 + 65536 + c * 256 + node, bank + node, bb);
                        node = 2 * node + (int)bb;
                    }
                }
            } else {
                if (lane == 0) s_rankHist[c] = (u8)ilog2_dev((u32)rank);
                short *bank = M + O_RP;
                for (int node = 1, bit = x.maxRank; bit >= 0; --bit) {
                    u32 bb = ((u32)rank >> bit) & 1u;
                    enc_decision<K_RANK_P>(rc, bank + 256 + st * 256 + node, bank + 256 + 65536 + c * 256 + node, bank + node, bb);
                    node = 2 * node + (int)bb;
                }
            }
            x.avgRank = (x.avgRank * 124 + rank * 4) >> 7;
            const int rank0 = rank - 1;
            const int rh = s_runHist[c];
            st = t_run[(x.ctxRank0 << 10) | (x.ctxRun << 6) | ((rank0 < 7 ? rank0 : 7) << 3) | (rh < 7 ? rh : 7)];

            enc_decision<K_RUN_T>(rc, M + O_UT_STATE + st, M + O_UT_CHAR + c, M + O_UT_SHARED, run != 1);
            if (run == 1) { if (lane == 0) s_runHist[c] = (u8)((rh + 2) >> 2); }
            else {
                const int e = ilog2_dev((u32)run);
                if (lane == 0) s_runHist[c] = (u8)((rh + 3 * e + 3) >> 2);
                short *eb = M + O_UE;
                for (int b = 1; b < e; ++b) enc_decision<K_RUN_E>(rc, eb + 32 + st * 32 + b - 1, eb + 32 + 8192 + c * 32 + b - 1, eb + b - 1, 1);
                enc_decision<K_RUN_E>(rc, eb + 32 + st * 32 + e - 1, eb + 32 + 8192 + c * 32 + e - 1, eb + e - 1, 0);
                short *bank = M + O_UM + (u32)e * NARROW;
                for (int node = 1, bit = e - 1; bit >= 0; --bit) {
                    u32 bb = ((u32)run >> bit) & 1u;
                    enc_decision<K_RUN_M>(rc, bank + 32 + st * 32 + node, bank + 32 + 8192 + c * 32 + node, bank + node, bb);
                    node = (e <= 5) ? 2 * node + (int)bb : node + 1;          // qlfc.cpp:1119
                }
            }
            __syncwarp();
            x.ctxRank0 = ((x.ctxRank0 << 1) | (rank0 == 0)) & 0x7;
            x.ctxRank4 = ((x.ctxRank4 << 2) | (rank0 < 3 ? rank0 : 3)) & 0xff;
            x.ctxRun   = ((x.ctxRun << 1) | (run < 3)) & 0xf;
        }
    }
    if (result == 0) result = (int)rc.finish();
    if (lane == 0) sb.result = result;
}

// ---------------------------------------------------------------------------------------------
// decoder (qlfc.cpp:1672-1927).  One warp per sub-block, lock-step; runs are expanded warp-wide.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) q_decode(const u8 *__restrict__ in_all, SubBlock *__restrict__ sbs, short *__restrict__ models,
                                               const QTables *__restrict__ tables, u8 *__restrict__ out_all, const u32 *__restrict__ sb_list)
{
    extern __shared__ __align__(16) u8 s_tab[];
    load_tables(s_tab, tables);
    const u8 *t_rank = s_tab, *t_run = s_tab + 32768;
    __shared__ u8 s_rankHist[256], s_runHist[256], s_mtf[256 + 32];

    const u32 sid = sb_list[blockIdx.x];
    SubBlock &sb = sbs[sid];
    short *M = models + (size_t)blockIdx.x * MODEL_SHORTS_PAD;
    const u32 lane = threadIdx.x;
    for (int i = lane; i < 256; i += 32) { s_rankHist[i] = 0; s_runHist[i] = 0; s_mtf[i] = 0; }
    __syncwarp();

    RcDec rc; rc.in = in_all + sb.out_off; rc.pos = 0; rc.limit = sb.out_cap; rc.code = 0; rc.range = 0xffffffffu;
    for (int i = 0; i < 3; ++i) rc.code = (rc.code << 16) | rc.get16();
    u32 n = 0; for (int b = 0; b < 32; ++b) n = (n << 1) | rc.decode(2048);
    if (n > sb.in_size) { if (lane == 0) sb.result = LIBBSC_DATA_CORRUPT; return; }   // would overrun the output slice

    RunCtx x; x.ctxRank0 = x.ctxRank4 = x.ctxRun = 0; x.maxRank = 7; x.avgRank = 0;
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
            if (lane == 0) s_mtf[d] = (u8)c;
            if (c == prev) { x.maxRank = ilog2_dev((u32)(d - 1)); break; }
            prev = c; if ((u32)(c >> 3) == lane) used8 |= 1u << (c & 7);
        }
    }
    __syncwarp();

    u8 *out = out_all + sb.in_start;
    for (u32 i = 0; i < n; ) {
        const int c = s_mtf[0];
        int rank = 1; u32 b;
        int st = t_rank[(x.ctxRun << 11) | (x.ctxRank4 << 3) | s_rankHist[c]];
        if (x.avgRank < 32) {
            b = dec_decision<K_RANK_T>(rc, M + O_RT_STATE + st, M + O_RT_CHAR + c, M + O_RT_SHARED);
            if (!b) { if (lane == 0) s_rankHist[c] = 0; }
            else {
                int e = 1;
                while (e != x.maxRank) {
                    b = dec_decision<K_RANK_E>(rc, M + O_RE_STATE + st * 8 + e - 1, M + O_RE_CHAR + c * 8 + e - 1, M + O_RE_SHARED + e - 1);
                    if (!b) break;
                    if (++e >= 7) break;                                     // e <= maxRank <= 7 in valid streams
                }
                if (lane == 0) s_rankHist[c] = (u8)e;
                short *bank = M + O_RM + (u32)e * WIDE;
                for (int bit = e - 1; bit >= 0; --bit) {
                    b = dec_decision<K_RANK_M>(rc, bank + 256 + st * 256 + rank, bank + 256 + 65536 + c * 256 + rank, bank + rank);
                    rank = 2 * rank + (int)b;
                }
            }
        } else {
            rank = 0;
            short *bank = M + O_RP;
            for (int node = 1, bit = x.maxRank; bit >= 0; --bit) {
                b = dec_decision<K_RANK_P>(rc, bank + 256 + st * 256 + node, bank + 256 + 65536 + c * 256 + node, bank + node);
                node = 2 * node + (int)b; rank = 2 * rank + (int)b;
            }
            if (lane == 0) s_rankHist[c] = (u8)ilog2_dev((u32)rank);
        }
        rank &= 255;
        __syncwarp();
        // push c `rank` places back: mtf[0..rank-1] = mtf[1..rank]; mtf[rank] = c  (qlfc.cpp:1830-1860)
        for (int basep = 0; basep < rank; basep += 32) {
            int p = basep + (int)lane;
            u8 v = s_mtf[p + 1];
            __syncwarp();
            if (p < rank) s_mtf[p] = v;
            __syncwarp();
        }
        if (lane == 0) s_mtf[rank] = (u8)c;
        __syncwarp();

        x.avgRank = (x.avgRank * 124 + rank * 4) >> 7;
        const int rank0 = rank - 1;
        const int rh = s_runHist[c];
        st = t_run[(x.ctxRank0 << 10) | (x.ctxRun << 6) | (((u32)rank0 < 7u ? rank0 : 7) << 3) | (rh < 7 ? rh : 7)];
        u32 run = 1;
        b = dec_decision<K_RUN_T>(rc, M + O_UT_STATE + st, M + O_UT_CHAR + c, M + O_UT_SHARED);
        if (!b) { if (lane == 0) s_runHist[c] = (u8)((rh + 2) >> 2); }
        else {
            int e = 1;
            short *eb = M + O_UE;
            for (;;) {
                b = dec_decision<K_RUN_E>(rc, eb + 32 + st * 32 + e - 1, eb + 32 + 8192 + c * 32 + e - 1, eb + e - 1);
                if (!b) break;
                if (++e >= 31) break;                                        // corrupt-input guard
            }
            if (lane == 0) s_runHist[c] = (u8)((rh + 3 * e + 3) >> 2);
            short *bank = M + O_UM + (u32)e * NARROW;
            for (int node = 1, bit = e - 1; bit >= 0; --bit) {
                b = dec_decision<K_RUN_M>(rc, bank + 32 + st * 32 + node, bank + 32 + 8192 + c * 32 + node, bank + node);
                run = 2 * run + b;
                node = (e <= 5) ? 2 * node + (int)b : node + 1;
            }
        }
        __syncwarp();
        x.ctxRank0 = ((x.ctxRank0 << 1) | (rank0 == 0)) & 0x7;
        x.ctxRank4 = ((x.ctxRank4 << 2) | ((u32)rank0 < 3u ? rank0 : 3)) & 0xff;
        x.ctxRun   = ((x.ctxRun << 1) | (run < 3)) & 0xf;

        if (run > n - i) run = n - i;                                        // never write past n
        for (u32 k = lane; k < run; k += 32) out[i + k] = (u8)c;
        i += run;
    }
    if (lane == 0) sb.result = (int)n;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static int coder_num_blocks(int n)                        // coder.cpp:52-59
{
    if (n < 256 * 1024) return 1;
    if (n < 4 * 1024 * 1024) return 2;
    if (n < 16 * 1024 * 1024) return 4;
    return 8;
}

static const QTables *get_tables(Ctx *ctx)
{
    if (!ctx->qlfc_tables) {
        QTables *d = nullptr;
        CUDA_TRY(cudaMalloc((void **)&d, sizeof(QTables)));
        CUDA_TRY(cudaMemcpyAsync(d->rank_state, bscb_rank_state_tab, 32768, cudaMemcpyHostToDevice, ctx->stream));
        CUDA_TRY(cudaMemcpyAsync(d->run_state, bscb_run_state_tab, 8192, cudaMemcpyHostToDevice, ctx->stream));
        CUDA_TRY(cudaMemcpyToSymbolAsync(c_params, bscb_static_params, sizeof(c_params), 0, cudaMemcpyHostToDevice, ctx->stream));
        ctx->sync();
        ctx->qlfc_tables = d;
    }
    return (const QTables *)ctx->qlfc_tables;
}

static void init_models(Ctx *ctx, short *models, int count)
{
    size_t total = (size_t)count * MODEL_SHORTS_PAD;
    LAUNCH(ctx, q_model_init, ceil_div(total, 256 * 8), 256, 0, models, total);
}

int stage_coder_compress(Ctx *ctx, const u8 *d_in, u8 *d_out, int n_, int coder, int features)
{
    if (coder != 1) return (coder == 2 || coder == 3) ? LIBBSC_NOT_SUPPORTED : LIBBSC_BAD_PARAMETER;
    if (n_ <= 0) return LIBBSC_BAD_PARAMETER;
    const u32 n = (u32)n_;
    const int nBlocks = coder_num_blocks(n_);
    const QTables *tables = get_tables(ctx);
    Arena &A = ctx->arena;
    const size_t mark = A.mark();

    u32 *d_start = A.get<u32>(2 * Q_MAX_SUB), *d_size = d_start + Q_MAX_SUB;
    SubBlock *d_sb = A.get<SubBlock>(Q_MAX_SUB);
    const u32 run_tiles = ceil_div(n, RUN_TILE);
    u32 *tile_cnt = A.get<u32>(run_tiles + 1);
    u8 *mtf = A.get<u8>(256 * Q_MAX_SUB);
    short *models = A.get<short>((size_t)nBlocks * MODEL_SHORTS_PAD);
    u8 *tmp = A.get<u8>((size_t)n + (4096 + 256) * Q_MAX_SUB);

    // 1. split
    if (nBlocks > 1) LAUNCH(ctx, q_split, 1, 1024, 0, d_in, n, (u32)nBlocks, d_start, d_size);
    else { u32 h[2] = {0, n}; CUDA_TRY(cudaMemcpyAsync(d_start, &h[0], 4, cudaMemcpyHostToDevice, ctx->stream)); CUDA_TRY(cudaMemcpyAsync(d_size, &h[1], 4, cudaMemcpyHostToDevice, ctx->stream)); ctx->sync(); }
    // 2. runs
    LAUNCH(ctx, q_run_count, run_tiles, RUN_THREADS, 0, d_in, n, d_start, (u32)nBlocks, tile_cnt);
    LAUNCH(ctx, q_scan_tiles, 1, 1024, 0, tile_cnt, run_tiles, ctx->d_mail);
    CUDA_TRY(cudaMemcpyAsync(ctx->d_mail + 8, d_start, sizeof(u32) * 2 * Q_MAX_SUB, cudaMemcpyDeviceToDevice, ctx->stream));
    ctx->fetch_mail(8 + 2 * Q_MAX_SUB);
    const u32 R = ctx->h_mail[0];
    u32 h_start[Q_MAX_SUB], h_size[Q_MAX_SUB];
    for (int b = 0; b < nBlocks; ++b) { h_start[b] = ctx->h_mail[8 + b]; h_size[b] = ctx->h_mail[8 + Q_MAX_SUB + b]; }

    u32 *run_pos = A.get<u32>((size_t)R + 2);
    u8 *run_sym = A.get<u8>((size_t)R + 32);
    u8 *run_rank = A.get<u8>((size_t)R + 32);
    LAUNCH(ctx, q_run_write, run_tiles, RUN_THREADS, 0, d_in, n, d_start, (u32)nBlocks, tile_cnt, run_pos, run_sym);
    LAUNCH(ctx, q_run_bounds, 1, 32, 0, run_pos, R, n, d_start, d_size, (u32)nBlocks, d_sb);
    // output slices in tmp (256-byte aligned, 4 KB slack each); capacity = the sub-block's input size (coder.cpp:193)
    SubBlock h_sb[Q_MAX_SUB];
    CUDA_TRY(cudaMemcpyAsync(h_sb, d_sb, sizeof(SubBlock) * nBlocks, cudaMemcpyDeviceToHost, ctx->stream));
    ctx->sync();
    for (u32 b = 0, off = 0; b < (u32)nBlocks; ++b) {
        h_sb[b].out_off = off; off += (u32)align_up((size_t)h_size[b] + 4096, 256);   // 16-bit stores need even offsets
        h_sb[b].out_cap = (nBlocks == 1) ? n - 1 : h_size[b];
        h_sb[b].result = 0; h_sb[b].nsym = 0;
    }
    CUDA_TRY(cudaMemcpyAsync(d_sb, h_sb, sizeof(SubBlock) * nBlocks, cudaMemcpyHostToDevice, ctx->stream));
    // 3. ranks, 4. encode
    PROF_BYTES(ctx, 2.0 * R);
    LAUNCH(ctx, q_ranks, nBlocks, 32, 0, run_sym, run_rank, d_sb, mtf);
    init_models(ctx, models, nBlocks);
    CUDA_TRY(cudaFuncSetAttribute(q_encode, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(QTables)));
    PROF_BYTES(ctx, (double)n);                          // + c written; the launch is latency-, not bandwidth-bound
    LAUNCH(ctx, q_encode, nBlocks, 32, sizeof(QTables), run_pos, run_sym, run_rank, d_sb, mtf, models, tables, tmp, (const u32 *)nullptr);
    CUDA_TRY(cudaMemcpyAsync(h_sb, d_sb, sizeof(SubBlock) * nBlocks, cudaMemcpyDeviceToHost, ctx->stream));
    ctx->sync();

    int result;
    if (nBlocks == 1) {                                   // coder.cpp:113-119
        result = h_sb[0].result;
        if (result >= 0) {
            u8 one = 1;
            CUDA_TRY(cudaMemcpyAsync(d_out, &one, 1, cudaMemcpyHostToDevice, ctx->stream));
            CUDA_TRY(cudaMemcpyAsync(d_out + 1, tmp + h_sb[0].out_off, (size_t)result, cudaMemcpyDeviceToDevice, ctx->stream));
            ctx->sync();
            result += 1;
        }
        A.release(mark);
        return result;
    }

    int res[Q_MAX_SUB];
    const bool parallel_rules = (features & LIBBSC_FEATURE_MULTITHREADING) != 0;
    int ptr = 1 + 8 * nBlocks;
    result = 0;
    if (parallel_rules) {                                 // coder.cpp:159-240
        int total = ptr;
        for (int b = 0; b < nBlocks; ++b) { res[b] = h_sb[b].result < 0 ? (int)h_size[b] : h_sb[b].result; total += res[b]; }
        if (total >= n_) result = LIBBSC_NOT_COMPRESSIBLE;
    } else {                                              // coder.cpp:111-155
        for (int b = 0; b < nBlocks && result == 0; ++b) {
            int room = (int)h_size[b]; if (room > n_ - ptr) room = n_ - ptr;
            int r = h_sb[b].result;
            if (room != (int)h_size[b]) {                 // clipped output size: redo this sub-block with the clipped room
                h_sb[b].out_cap = (u32)(room > 0 ? room : 0); h_sb[b].result = 0;
                u32 *d_list = A.get<u32>(1); u32 one = (u32)b;
                CUDA_TRY(cudaMemcpyAsync(d_sb + b, &h_sb[b], sizeof(SubBlock), cudaMemcpyHostToDevice, ctx->stream));
                CUDA_TRY(cudaMemcpyAsync(d_list, &one, 4, cudaMemcpyHostToDevice, ctx->stream));
                ctx->sync();
                init_models(ctx, models + (size_t)b * MODEL_SHORTS_PAD, 1);
                LAUNCH(ctx, q_encode, 1, 32, sizeof(QTables), run_pos, run_sym, run_rank, d_sb, mtf, models, tables, tmp, d_list);
                CUDA_TRY(cudaMemcpyAsync(&h_sb[b], d_sb + b, sizeof(SubBlock), cudaMemcpyDeviceToHost, ctx->stream));
                ctx->sync();
                r = h_sb[b].result;
            }
            if (r < 0) { if (ptr + (int)h_size[b] >= n_) { result = LIBBSC_NOT_COMPRESSIBLE; break; } r = (int)h_size[b]; h_sb[b].result = -1; }
            res[b] = r; ptr += r;
        }
        ptr = 1 + 8 * nBlocks;
    }
    if (result == 0) {
        u8 *hdr = (u8 *)(ctx->h_mail + 32);               // pinned scratch (<= 65 bytes)
        hdr[0] = (u8)nBlocks;
        for (int b = 0; b < nBlocks; ++b) {
            u32 a = h_size[b], c = (u32)res[b];
            memcpy(hdr + 1 + 8 * b, &a, 4); memcpy(hdr + 5 + 8 * b, &c, 4);
        }
        CUDA_TRY(cudaMemcpyAsync(d_out, hdr, (size_t)(1 + 8 * nBlocks), cudaMemcpyHostToDevice, ctx->stream));
        for (int b = 0; b < nBlocks; ++b) {
            // coder.cpp:137-141, 224-231: a sub-block whose stored size equals its input size is raw
            const u8 *src = (res[b] == (int)h_size[b]) ? d_in + h_start[b] : tmp + h_sb[b].out_off;
            CUDA_TRY(cudaMemcpyAsync(d_out + ptr, src, (size_t)res[b], cudaMemcpyDeviceToDevice, ctx->stream));
            ptr += res[b];
        }
        ctx->sync();
        result = ptr;
    }
    A.release(mark);
    return result;
}

int stage_coder_decompress(Ctx *ctx, const u8 *d_in, int in_size, u8 *d_out, int out_cap, int coder, int features)
{
    (void)features;
    if (coder != 1) return (coder == 2 || coder == 3) ? LIBBSC_NOT_SUPPORTED : LIBBSC_BAD_PARAMETER;
    if (in_size < 1) return LIBBSC_UNEXPECTED_EOB;
    const QTables *tables = get_tables(ctx);
    Arena &A = ctx->arena;
    const size_t mark = A.mark();

    u8 hdr[1 + 8 * 255];
    {
        int want = in_size < 65 ? in_size : 65;
        CUDA_TRY(cudaMemcpyAsync(ctx->h_mail + 32, d_in, (size_t)want, cudaMemcpyDeviceToHost, ctx->stream));
        ctx->sync();
        memcpy(hdr, ctx->h_mail + 32, (size_t)want);
    }
    const int nBlocks = hdr[0];
    SubBlock h_sb[Q_MAX_SUB]; u32 list[Q_MAX_SUB]; int nlist = 0;
    memset(h_sb, 0, sizeof h_sb);
    if (nBlocks == 1) {
        h_sb[0].in_start = 0; h_sb[0].in_size = (u32)out_cap; h_sb[0].out_off = 1; h_sb[0].out_cap = (u32)(in_size - 1);
        list[nlist++] = 0;
    } else {
        if (nBlocks == 0 || nBlocks > Q_MAX_SUB || in_size < 1 + 8 * nBlocks) { A.release(mark); return LIBBSC_DATA_CORRUPT; }
        long long inPtr = 1 + 8 * nBlocks, outPtr = 0;
        for (int b = 0; b < nBlocks; ++b) {
            int rawSize, packed; memcpy(&rawSize, hdr + 1 + 8 * b, 4); memcpy(&packed, hdr + 5 + 8 * b, 4);
            if (rawSize < 0 || packed < 0 || inPtr + packed > in_size || outPtr + rawSize > out_cap) { A.release(mark); return LIBBSC_DATA_CORRUPT; }
            h_sb[b].in_start = (u32)outPtr; h_sb[b].in_size = (u32)rawSize; h_sb[b].out_off = (u32)inPtr; h_sb[b].out_cap = (u32)packed;
            if (packed != rawSize) list[nlist++] = (u32)b;
            else { h_sb[b].result = rawSize; if (rawSize) CUDA_TRY(cudaMemcpyAsync(d_out + outPtr, d_in + inPtr, (size_t)rawSize, cudaMemcpyDeviceToDevice, ctx->stream)); }
            inPtr += packed; outPtr += rawSize;
        }
    }
    if (nlist > 0) {
        SubBlock *d_sb = A.get<SubBlock>(Q_MAX_SUB);
        u32 *d_list = A.get<u32>(Q_MAX_SUB);
        short *models = A.get<short>((size_t)nlist * MODEL_SHORTS_PAD);
        CUDA_TRY(cudaMemcpyAsync(d_sb, h_sb, sizeof h_sb, cudaMemcpyHostToDevice, ctx->stream));
        CUDA_TRY(cudaMemcpyAsync(d_list, list, sizeof list, cudaMemcpyHostToDevice, ctx->stream));
        ctx->sync();
        init_models(ctx, models, nlist);
        CUDA_TRY(cudaFuncSetAttribute(q_decode, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(QTables)));
        PROF_BYTES(ctx, (double)in_size + (double)out_cap);
        LAUNCH(ctx, q_decode, nlist, 32, sizeof(QTables), d_in, d_sb, models, tables, d_out, d_list);
        CUDA_TRY(cudaMemcpyAsync(h_sb, d_sb, sizeof h_sb, cudaMemcpyDeviceToHost, ctx->stream));
    }
    ctx->sync();
    int total = 0, err = 0;
    for (int b = 0; b < (nBlocks == 1 ? 1 : nBlocks); ++b) { if (h_sb[b].result < 0) err = h_sb[b].result; total += h_sb[b].result; }
    A.release(mark);
    return err ? err : total;
}
