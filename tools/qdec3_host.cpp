// tools/qdec3_host.cpp -- compiles libbsc_b200/csrc/qlfc_decoder3.cuh FOR THE HOST (QD3_HOST): the decoder's
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

namespace {
enum { K_RANK_T, K_RANK_E, K_RANK_M, K_RANK_P, K_RUN_T, K_RUN_E, K_RUN_M };
struct SubBlock { u32 in_start, in_size, run_begin, run_end, out_off, out_cap; int result; u32 nsym, tile_base, tiles, stat_cached, stat_miss; };
struct QTables { u8 rank_state[32768]; u8 run_state[8192]; };
#include "../libbsc_b200/csrc/qlfc_coder.cuh"
#define QD3_HOST 1
#include "../libbsc_b200/csrc/qlfc_decoder3.cuh"
}

// Decodes one QLFC static stream (what bsc_qlfc_static_decode_block reads) of `in_size` bytes into out[0..out_cap).
// mode 0: the speculative decoder, mode 1: the serial decoder, mode 2: the pipelined serial decoder
extern "C" int qdec3_host_decode(const unsigned char *in, unsigned in_size, unsigned char *out, unsigned out_cap, unsigned *stats, int mode)
{
    u8 *smem = (u8 *)calloc(1, sizeof(Dec3Smem));
    short *cold = (short *)malloc(sizeof(short) * 2 * (size_t)COLD_PAD);
    if (!smem || !cold) { free(smem); free(cold); return LIBBSC_NOT_ENOUGH_MEMORY; }
    Dec3Smem *D = (Dec3Smem *)smem;
    memcpy(D->cs.rank_state, bscb_rank_state_tab, 32768);
    memcpy(D->cs.run_state, bscb_run_state_tab, 8192);
    for (u32 i = 0; i < S16_COUNT; ++i) D->cs.s16[i] = 2048;
    for (size_t i = 0; i < 2 * (size_t)COLD_PAD; ++i) cold[i] = 2048;
    SM3 sm; sm.b = smem;
    u32 st_cached = 0, st_miss = 0;
    const int r = mode == 0 ? qd3_decode_stream<false>(sm, in, in_size, out, out_cap, cold, cold + COLD_PAD, st_cached, st_miss)
                : mode == 1 ? qd3_decode_stream_serial<false>(sm, in, in_size, out, out_cap, cold, cold + COLD_PAD, st_cached, st_miss)
                            : qd3_decode_stream_pipe<false>(sm, in, in_size, out, out_cap, cold, cold + COLD_PAD, st_cached, st_miss);
    if (stats) { stats[0] = st_cached; stats[1] = st_miss; }
    free(smem); free(cold);
    return r;
}
