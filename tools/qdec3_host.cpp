// tools/qdec3_host.cpp -- compiles the single-warp QLFC coders of libbsc_b200/csrc FOR THE HOST (QD3_HOST): the decoder's
// lane-parallel phases run as loops over 32 emulated lanes, the shared-memory counter file is a plain byte
// array.  Test infrastructure only (tests/test_qdec3_host.py compares it with the oracle on the CPU); the
// product never links or loads this file -- it checks the LOGIC of the CUDA decoder where there is no GPU.
//
//   g++ -O2 -shared -fPIC -o tools/bin/libqdec3_host.so tools/qdec3_host.cpp
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <cstddef>

typedef uint8_t  u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;

#define LIBBSC_DATA_CORRUPT -6
#define LIBBSC_NOT_ENOUGH_MEMORY -2
#define LIBBSC_NOT_COMPRESSIBLE -3

// just enough of the CUDA dialect for qlfc_tables.inc / qlfc_coder.cuh
#define __CUDACC__ 1
#define __host__
#define __device__
#define __forceinline__ inline
struct uint4 { u32 x, y, z, w; };
static struct { u32 x; } threadIdx;
static inline void __syncwarp() {}
#define __align__(n) alignas(n)

#include "../libbsc_b200/csrc/qlfc_tables.inc"
#include "../libbsc_b200/csrc/qlfc_tables2.inc"

#ifdef QD6_TABSIM
// LRU model of the L1 behind the state-table look-ups of a TablesInGlobal layout: fully associative, 128-byte lines, three capacities.
// (diagnostic build only: g++ -DQD6_TABSIM; results in DESIGN.md 4.5)
static const int TS_CAP[3] = {64, 128, 224};
static struct { unsigned line[224]; int n; unsigned long long hit, miss; } ts_lru[3];
static unsigned long long ts_touch = 0; static unsigned char ts_seen[41 * 1024 / 128 + 1];
static void tabsim_touch(unsigned idx)
{
    const unsigned ln = idx >> 7; ++ts_touch; ts_seen[ln] = 1;
    for (int c = 0; c < 3; ++c) {
        auto &L = ts_lru[c]; int k = 0;
        while (k < L.n && L.line[k] != ln) ++k;
        if (k < L.n) ++L.hit; else { ++L.miss; if (L.n < TS_CAP[c]) k = L.n++; else k = L.n - 1; }
        for (; k > 0; --k) L.line[k] = L.line[k - 1];
        L.line[0] = ln;
    }
}
#define QD6_TABSIM_TOUCH(i) tabsim_touch(i)
extern "C" void qdec6_tabsim_report(unsigned long long *out)       // touches, distinct lines, then (hits, misses) x 3 capacities; resets
{
    unsigned d = 0; for (unsigned i = 0; i < sizeof(ts_seen); ++i) d += ts_seen[i];
    out[0] = ts_touch; out[1] = d;
    for (int c = 0; c < 3; ++c) { out[2 + 2 * c] = ts_lru[c].hit; out[3 + 2 * c] = ts_lru[c].miss; }
    memset(ts_lru, 0, sizeof(ts_lru)); memset(ts_seen, 0, sizeof(ts_seen)); ts_touch = 0;
}
#endif

namespace {
enum { K_RANK_T, K_RANK_E, K_RANK_M, K_RANK_P, K_RUN_T, K_RUN_E, K_RUN_M };
struct SubBlock { u32 in_start, in_size, run_begin, run_end, out_off, out_cap; int result; u32 nsym, tile_base, tiles, stat_cached, stat_miss; };
struct QTables { u8 rank_state[32768]; u8 run_state[8192]; };
#include "../libbsc_b200/csrc/qlfc_coder.cuh"
#define QD3_HOST 1
#include "../libbsc_b200/csrc/qlfc_lanes.cuh"
#include "../libbsc_b200/csrc/qlfc_fast.cuh"
#include "../libbsc_b200/csrc/qlfc_decoder6.cuh"
#include "../libbsc_b200/csrc/qlfc_adaptive.cuh"
}

// ---- fast coder (coder id 3), libbsc_b200/csrc/qlfc_fast.cuh ---------------------------------------------------------
static u8 *fast_smem_new(short **cold_out)
{
    FastSmem *F = (FastSmem *)calloc(1, sizeof(FastSmem));
    short *cold = (short *)malloc(sizeof(short) * (size_t)QF_COLD);
    if (!F || !cold) { free(F); free(cold); return nullptr; }
    for (u32 i = 0; i < 256 * 8; ++i) F->re[i] = 4096;
    for (u32 i = 0; i < 256 * QF_ROW; ++i) { F->rm[i] = 4096; F->um[i] = 1024; }
    for (u32 i = 0; i < 256 * 32; ++i) F->ue[i] = 1024;
    for (u32 i = 0; i < QF_COLD; ++i) cold[i] = i < QF_COLD_RANK ? 4096 : 1024;
    *cold_out = cold;
    return (u8 *)F;
}

extern "C" int qfast_host_decode(const unsigned char *in, unsigned in_size, unsigned char *out, unsigned out_cap, unsigned *stats)
{
    short *cold = nullptr; u8 *smem = fast_smem_new(&cold);
    if (!smem) return LIBBSC_NOT_ENOUGH_MEMORY;
    SM3 sm; sm.b = smem;
    u32 st_cached = 0, st_miss = 0;
    const int r = qf_decode_stream(sm, in, in_size, out, out_cap, cold, st_cached, st_miss);
    if (stats) { stats[0] = st_cached; stats[1] = st_miss; }
    free(smem); free(cold);
    return r;
}

// run_pos has nruns + 1 entries (the last one = in_size); mtf is the 256-byte MTF-order table of the transform
extern "C" int qfast_host_encode(const unsigned *run_pos, const unsigned char *run_sym, const unsigned char *run_rank, unsigned nruns, unsigned in_size,
                                 const unsigned char *mtf, unsigned char *out, unsigned out_cap, unsigned *stats)
{
    short *cold = nullptr; u8 *smem = fast_smem_new(&cold);
    if (!smem) return LIBBSC_NOT_ENOUGH_MEMORY;
    SM3 sm; sm.b = smem;
    u32 st_cached = 0, st_miss = 0;
    const int r = qf_encode_stream(sm, run_pos, run_sym, run_rank, 0, nruns, in_size, mtf, out, out_cap, cold, st_cached, st_miss);
    if (stats) { stats[0] = st_cached; stats[1] = st_miss; }
    free(smem); free(cold);
    return r;
}

// ---- layout-templated serial decoder (libbsc_b200/csrc/qlfc_decoder6.cuh): layout 0 = full (205 KB), 1 = diet (101 KB) ------------
static const QTables *host_tables()
{
    static QTables t; static bool ready = false;
    if (!ready) { memcpy(t.rank_state, bscb_rank_state_tab, 32768); memcpy(t.run_state, bscb_run_state_tab, 8192); ready = true; }
    return &t;
}
template <class LY> static int qdec6_run(const unsigned char *in, unsigned in_size, unsigned char *out, unsigned out_cap, unsigned *stats)
{
    u8 *smem = (u8 *)calloc(1, LY::BYTES + LY::SHIFT);       // the host keeps the whole image; a TG layout just never reads its first SHIFT bytes
    short *cold = (short *)malloc(sizeof(short) * 2 * (size_t)COLD_PAD);
    if (!smem || !cold) { free(smem); free(cold); return LIBBSC_NOT_ENOUGH_MEMORY; }
    memcpy(smem + LY::O_RANK_STATE, bscb_rank_state_tab, 32768);
    memcpy(smem + LY::O_RUN_STATE, bscb_run_state_tab, 8192);
    for (u32 i = 0; i < LY::S16_COUNT; ++i) { const u16 v = 2048; memcpy(smem + LY::O_S16 + 2 * i, &v, 2); }
    for (size_t i = 0; i < 2 * (size_t)COLD_PAD; ++i) cold[i] = 2048;
    SM3 sm; sm.b = smem;
    u32 st_cached = 0, st_miss = 0;
    int moves[QD6_MOVES]; qd6_fill_moves(moves);
    const int r = qd6_decode_stream<LY, false>(sm, in, in_size, out, out_cap, cold, cold + COLD_PAD, moves, (const u8 *)host_tables(), st_cached, st_miss);
    if (stats) { stats[0] = st_cached; stats[1] = st_miss; }
    free(smem); free(cold);
    return r;
}
extern "C" int qdec6_host_decode(const unsigned char *in, unsigned in_size, unsigned char *out, unsigned out_cap, unsigned *stats, int layout)
{
    return layout == 0 ? qdec6_run<LayoutFull>(in, in_size, out, out_cap, stats)         // 0: the full layout (what the adaptive coder derives from)
         : layout == 1 ? qdec6_run<LayoutDiet>(in, in_size, out, out_cap, stats)         // 1: two streams per SM
         : layout == 2 ? qdec6_run<LayoutDietTG>(in, in_size, out, out_cap, stats)       // 2: state tables in global memory, three per SM
         : layout == 3 ? qdec6_run<LayoutDiet4>(in, in_size, out, out_cap, stats)        // 3: four per SM
                       : qdec6_run<LayoutDiet5>(in, in_size, out, out_cap, stats);       // 4: five per SM
}
extern "C" unsigned qdec6_smem_bytes(int layout) { return layout == 0 ? LayoutFull::BYTES : layout == 1 ? LayoutDiet::BYTES : layout == 2 ? LayoutDietTG::BYTES : layout == 3 ? LayoutDiet4::BYTES : LayoutDiet5::BYTES; }

// ---- adaptive coder (coder id 2), libbsc_b200/csrc/qlfc_adaptive.cuh ---------------------------------------------------
static u8 *adaptive_smem_new(short **cold_out)
{
    u8 *smem = (u8 *)calloc(1, QA_BYTES);
    short *cold = (short *)malloc(sizeof(short) * 2 * (size_t)COLD_PAD);
    if (!smem || !cold) { free(smem); free(cold); return nullptr; }
    memcpy(smem + ALY::O_RANK_STATE, bscb_rank_state_tab, 32768);
    memcpy(smem + ALY::O_RUN_STATE, bscb_run_state_tab, 8192);
    for (u32 i = 0; i < ALY::S16_COUNT; ++i) { const u16 v = 2048; memcpy(smem + ALY::O_S16 + 2 * i, &v, 2); }
    memcpy(smem + OA_STRETCH, bscb_stretch_le16, 2 * 4097);
    memcpy(smem + OA_SQUASH, bscb_squash_le16, 2 * 4097);
    SM3 sm; sm.b = smem;
    for (u32 m = 0; m < QA_MIXERS; ++m) qa_init_mixer(sm, m);
    for (size_t i = 0; i < 2 * (size_t)COLD_PAD; ++i) cold[i] = 2048;
    *cold_out = cold;
    return smem;
}
extern "C" int qadapt_host_decode(const unsigned char *in, unsigned in_size, unsigned char *out, unsigned out_cap, unsigned *stats)
{
    short *cold = nullptr; u8 *smem = adaptive_smem_new(&cold);
    if (!smem) return LIBBSC_NOT_ENOUGH_MEMORY;
    SM3 sm; sm.b = smem;
    u32 st_cached = 0, st_miss = 0;
    const int r = qa_decode_stream(sm, in, in_size, out, out_cap, cold, cold + COLD_PAD, st_cached, st_miss);
    if (stats) { stats[0] = st_cached; stats[1] = st_miss; }
    free(smem); free(cold);
    return r;
}
extern "C" int qadapt_host_encode(const unsigned *run_pos, const unsigned char *run_sym, const unsigned char *run_rank, unsigned nruns, unsigned in_size,
                                  const unsigned char *mtf, unsigned char *out, unsigned out_cap, unsigned *stats)
{
    short *cold = nullptr; u8 *smem = adaptive_smem_new(&cold);
    if (!smem) return LIBBSC_NOT_ENOUGH_MEMORY;
    SM3 sm; sm.b = smem;
    u32 st_cached = 0, st_miss = 0;
    const int r = qa_encode_stream(sm, run_pos, run_sym, run_rank, 0, nruns, in_size, mtf, out, out_cap, cold, cold + COLD_PAD, st_cached, st_miss);
    if (stats) { stats[0] = st_cached; stats[1] = st_miss; }
    free(smem); free(cold);
    return r;
}
extern "C" unsigned qadapt_smem_bytes(void) { return QA_BYTES; }
