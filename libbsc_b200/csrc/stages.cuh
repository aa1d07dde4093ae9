// libbsc_b200/csrc/stages.cuh -- device-resident stage entry points (host functions that enqueue
// kernels on ctx->stream; all data pointers are DEVICE pointers unless noted).
#pragma once
#include "common.cuh"

// Forward BWT in place (bsc_bwt_encode, bwt.cpp:178).  num_indexes/indexes are HOST pointers (may be NULL).
int stage_bwt_encode(Ctx *ctx, u8 *d_T, int n, unsigned char *num_indexes, int *indexes);
// Inverse BWT in place (bsc_bwt_decode, bwt.cpp:283).
int stage_bwt_decode(Ctx *ctx, u8 *d_T, int n, int index);
// Sort transform of order k in place (bsc_st_encode, st.cpp:990; k = 3..8).
int stage_st_encode(Ctx *ctx, u8 *d_T, int n, int k);
// Inverse sort transform in place (bsc_st_decode, st.cpp:1491; k = 3..8, index = row of rotation 0).
int stage_st_decode(Ctx *ctx, u8 *d_T, int n, int k, int index);
// Coder container (bsc_coder_compress, coder.cpp:244).  d_out must hold n + 4096 bytes.
// bare_out_size >= 0: one QLFC stream without the container (bsc_qlfc_*_encode_block, qlfc.h:55-77), output capacity bare_out_size.
int stage_coder_compress(Ctx *ctx, const u8 *d_in, u8 *d_out, int n, int coder, int features, int bare_out_size = -1);
// bsc_coder_decompress (coder.cpp:273).  `in_size` bounds the readable input (device padded by >= 64 bytes);
// out_cap bounds the writable output.
int stage_coder_decompress(Ctx *ctx, const u8 *d_in, int in_size, u8 *d_out, int out_cap, int coder, int features, bool bare = false);   // bare: one stream, no container byte
// Adler-32 of a device buffer (bsc_adler32, adler32.cpp:83).  Result lands in ctx->d_mail[slot]
// (asynchronously); adler32_fetch() waits and returns it.
void stage_adler32_async(Ctx *ctx, const u8 *d_p, int n, int slot);
u32  stage_adler32(Ctx *ctx, const u8 *d_p, int n);
