/* include/libbsc_b200.h -- C ABI of the B200-native block-sorting hot path.
 *
 * Drop-in boundary: the first group of functions has EXACTLY the names, argument meaning, buffer
 * ownership and error behaviour of the libbsc 3.3.5 entry points it replaces (reference file:line
 * cited per function), so that libbsc/libbsc/libbsc.cpp -- or any program linking libbsc.a --
 * can bind to this library instead of bwt.o / st.o / coder.o / qlfc.o.
 * All pointers in that group are HOST pointers; the library moves the block to the current CUDA
 * device, runs the stage there and moves the result back.  Plain pointers and ints only.
 *
 * There is no CPU implementation behind these symbols: without a usable CUDA device they return
 * LIBBSC_GPU_NOT_SUPPORTED.
 *
 * The second group (bscb200_*) are extensions for callers that keep blocks resident in HBM and
 * manage their own streams (bench.py, the multi-GPU block scheduler).
 */
#ifndef LIBBSC_B200_H
#define LIBBSC_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* error codes, features, ids: libbsc/libbsc.h:41-76 */
#ifndef LIBBSC_NO_ERROR
#define LIBBSC_NO_ERROR                0
#define LIBBSC_BAD_PARAMETER          -1
#define LIBBSC_NOT_ENOUGH_MEMORY      -2
#define LIBBSC_NOT_COMPRESSIBLE       -3
#define LIBBSC_NOT_SUPPORTED          -4
#define LIBBSC_UNEXPECTED_EOB         -5
#define LIBBSC_DATA_CORRUPT           -6
#define LIBBSC_GPU_ERROR              -7
#define LIBBSC_GPU_NOT_SUPPORTED      -8
#define LIBBSC_GPU_NOT_ENOUGH_MEMORY  -9
#define LIBBSC_FEATURE_FASTMODE        1
#define LIBBSC_FEATURE_MULTITHREADING  2
#define LIBBSC_FEATURE_LARGEPAGES      4
#define LIBBSC_FEATURE_CUDA            8
#define LIBBSC_HEADER_SIZE             28
#endif
#define LIBBSC_BLOCKSORTER_BWT         1   /* 3..8 = ST3..ST8 */
#define LIBBSC_CODER_QLFC_STATIC       1
#define LIBBSC_CODER_QLFC_ADAPTIVE     2   /* experimental: LIBBSC_NOT_SUPPORTED unless BSCB200_ENABLE_ADAPTIVE=1 (csrc/qlfc_adaptive.cuh) */
#define LIBBSC_CODER_QLFC_FAST         3   /* experimental: LIBBSC_NOT_SUPPORTED unless BSCB200_ENABLE_FAST=1 (csrc/qlfc_fast.cuh) */

/* ---- group 1: libbsc-compatible entry points -------------------------------------------- */

/* libbsc/libbsc.h:95,104 ; libbsc/libbsc/libbsc.cpp:46-66.  Call once before anything else. */
int bsc_init(int features);
int bsc_init_full(int features, void *(*malloc_fn)(size_t), void *(*zero_malloc_fn)(size_t), void (*free_fn)(void *));

/* libbsc/libbsc.h:118 ; libbsc.cpp:213-338.  output holds n + 28 bytes; input may equal output.
 * bsc_compress: LZP (lzpHashSize/lzpMinLen != 0) is outside the replaced path: LIBBSC_NOT_SUPPORTED.
 * bsc_decompress: blocks with an LZP stage are accepted (the inverse stage runs on the host after the GPU stages). */
int bsc_compress(const unsigned char *input, unsigned char *output, int n, int lzpHashSize, int lzpMinLen, int blockSorter, int coder, int features);
/* libbsc/libbsc.h:128 ; libbsc.cpp:68-81 */
int bsc_store(const unsigned char *input, unsigned char *output, int n, int features);
/* libbsc/libbsc.h:139 ; libbsc.cpp:340-418 */
int bsc_block_info(const unsigned char *blockHeader, int headerSize, int *pBlockSize, int *pDataSize, int features);
/* libbsc/libbsc.h:150 ; libbsc.cpp:522-617 */
int bsc_decompress(const unsigned char *input, int inputSize, unsigned char *output, int outputSize, int features);

/* libbsc/bwt/bwt.h:45,56,68 ; libbsc/bwt/bwt.cpp:54-69, 178-231, 283-332.  In place on T. */
int bsc_bwt_init(int features);
int bsc_bwt_encode(unsigned char *T, int n, unsigned char *num_indexes, int *indexes, int features);
int bsc_bwt_decode(unsigned char *T, int n, int index, unsigned char num_indexes, int *indexes, int features);

/* libbsc/st/st.h:47,57,68 ; libbsc/st/st.cpp:990-1012 (k = 3..8; st.cu:334 for 7, 8).  In place on T. */
int bsc_st_init(int features);
int bsc_st_encode(unsigned char *T, int n, int k, int features);
int bsc_st_decode(unsigned char *T, int n, int k, int index, int features);   /* st.h:68 ; st.cpp:1491-1527 (csrc/st_decode.cu) */

/* libbsc/coder/coder.h:45,56,66 ; libbsc/coder/coder.cpp:244-347.  output of compress holds n + 4096 bytes. */
int bsc_coder_init(int features);
int bsc_coder_compress(const unsigned char *input, unsigned char *output, int n, int coder, int features);
int bsc_coder_decompress(const unsigned char *input, unsigned char *output, int coder, int features);
int bsc_qlfc_init(int features);                                              /* libbsc/coder/qlfc/qlfc.h:45 */
/* libbsc/coder/qlfc/qlfc.h:55-99 ; qlfc.cpp:2138-2226: ONE QLFC stream (what the container holds per sub-block).  encode: fails
 * with LIBBSC_NOT_COMPRESSIBLE when the stream would come within 16 bytes of outputSize (rangecoder.h:118-127).  decode: size-less
 * like the reference's; the stream is taken to end at most n + 16 bytes after `input` (n = the size coded in the stream). */
int bsc_qlfc_static_encode_block(const unsigned char *input, unsigned char *output, int inputSize, int outputSize);
int bsc_qlfc_adaptive_encode_block(const unsigned char *input, unsigned char *output, int inputSize, int outputSize);
int bsc_qlfc_fast_encode_block(const unsigned char *input, unsigned char *output, int inputSize, int outputSize);
int bsc_qlfc_static_decode_block(const unsigned char *input, unsigned char *output);
int bsc_qlfc_adaptive_decode_block(const unsigned char *input, unsigned char *output);
int bsc_qlfc_fast_decode_block(const unsigned char *input, unsigned char *output);

/* libbsc/adler32/adler32.h ; adler32.cpp:83 */
unsigned int bsc_adler32(const unsigned char *T, int n, int features);

/* libbsc/platform/platform.h:190-216 ; platform.cpp:192-260 (allocator hooks honoured for host temporaries) */
int   bsc_platform_init(int features, void *(*malloc_fn)(size_t), void *(*zero_malloc_fn)(size_t), void (*free_fn)(void *));
void *bsc_malloc(size_t size);
void *bsc_zero_malloc(size_t size);
void  bsc_free(void *address);

/* ---- group 2: extensions ---------------------------------------------------------------- */

/* bsc_coder_decompress with explicit bounds (what bsc_decompress uses internally) */
int bscb200_coder_decompress(const unsigned char *input, int inputSize, unsigned char *output, int outputCapacity, int coder, int features);

/* A context = one CUDA stream + one HBM workspace on `device`.  cuda_stream may be a
 * cudaStream_t owned by the caller (e.g. a torch stream) or NULL for a private stream.
 * One block at a time per context; use several contexts for concurrency. */
void              *bscb200_ctx_create(int device, void *cuda_stream);
void               bscb200_ctx_destroy(void *ctx);
int                bscb200_ctx_reserve(void *ctx, long long bytes);
int                bscb200_lzp_decompress_host(const unsigned char *input, int n, unsigned char *output, int outputCapacity, int lzpHashSize, int lzpMinLen);   /* inverse of the reference LZP stage (libbsc/lzp/lzp.h), host only */
int                bscb200_lzp_compress_host(const unsigned char *input, unsigned char *output, int n, int lzpHashSize, int lzpMinLen, int features);   /* the reference LZP stage forward (libbsc/lzp/lzp.h), host only */
int                bscb200_device_count(void);                    /* CUDA devices visible to the process */
long long          bscb200_device_free_bytes(void);               /* free HBM on the current device, -1 on error */
void               bscb200_release_pools(void);                   /* free the pooled contexts and sort slabs of every device (no call in flight) */
int                bscb200_set_device(int device);                /* bind the calling thread: all entry points use the current device */
long long          bscb200_workspace_bytes(int n, int blockSorter);   /* per-context workspace (~13 n + 64 MB): staged block + coder stage */
long long          bscb200_workspace_bytes_decode(int n);         /* the same (kept for callers of the round-1 ABI) */
long long          bscb200_scratch_bytes(int n, int blockSorter);  /* one device-wide sort slab (~58 n for BWT); BSCB200_SORT_SLABS of them per GPU (default 3) */
unsigned long long bscb200_ctx_kernel_launches(void *ctx);
unsigned long long bscb200_total_kernel_launches(void);
/* per-kernel CUDA-event timing of everything launched through ctx (bench.py's roofline leg) */
void               bscb200_ctx_set_profile(void *ctx, int on);
int                bscb200_ctx_profile_report(void *ctx, char *buf, int cap);
const char        *bscb200_version(void);

/* Device-resident variants: all d_* pointers are device pointers on the context's device.
 * Same return values as their host-pointer counterparts. */
int bscb200_compress_device(void *ctx, const unsigned char *d_input, unsigned char *d_output /* n+28 */, int n, int blockSorter, int coder, int features);
int bscb200_decompress_device(void *ctx, const unsigned char *d_input, int inputSize, unsigned char *d_output, int outputSize, int features);
int bscb200_bwt_encode_device(void *ctx, unsigned char *d_T, int n, unsigned char *num_indexes /* host */, int *indexes /* host */);
int bscb200_bwt_decode_device(void *ctx, unsigned char *d_T, int n, int index);
int bscb200_st_encode_device(void *ctx, unsigned char *d_T, int n, int k);
int bscb200_st_decode_device(void *ctx, unsigned char *d_T, int n, int k, int index);
int bscb200_coder_compress_device(void *ctx, const unsigned char *d_in, unsigned char *d_out /* n+4096 */, int n, int coder, int features);
int bscb200_coder_decompress_device(void *ctx, const unsigned char *d_in, int inputSize, unsigned char *d_out, int outputCapacity, int coder, int features);
unsigned int bscb200_adler32_device(void *ctx, const unsigned char *d_p, int n);

#ifdef __cplusplus
}
#endif
#endif
