/* oracle/bsc_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C CPU restatement of the libbsc 3.3.5 hot path (forward/inverse BWT, ST-k forward,
 * QLFC static coder, coder container, block framing).  It exists so that tests can check the
 * CUDA product path bit-for-bit on a box that has no /root/reference.  Nothing in the product
 * library (libbsc_b200/) may include, link or call this; only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline leg do.
 *
 * Pinning: every function here is checked against the UNMODIFIED reference compiled from
 * /root/reference (oracle/_ref/libbsc_ref.so, recipe in oracle/Makefile) and against the
 * known-answer values of SURVEY.md Appendix C (tests/test_oracle.py).  The reference ships no
 * tests or golden vectors of its own.
 *
 * All orc_* functions mirror the argument meaning and return conventions of the libbsc
 * function named in their comment.
 */
#ifndef BSC_ORACLE_H
#define BSC_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_NO_ERROR            0
#define ORC_BAD_PARAMETER      -1
#define ORC_NOT_ENOUGH_MEMORY  -2
#define ORC_NOT_COMPRESSIBLE   -3
#define ORC_NOT_SUPPORTED      -4
#define ORC_UNEXPECTED_EOB     -5
#define ORC_DATA_CORRUPT       -6

#define ORC_FEATURE_MULTITHREADING 2
#define ORC_HEADER_SIZE 28

/* libbsc/adler32/adler32.cpp:83 bsc_adler32 */
unsigned int orc_adler32(const unsigned char *p, int n);

/* libbsc/bwt/bwt.cpp:178 bsc_bwt_encode (CPU branch) ; index conventions libsais.c:6867-6893 */
int orc_bwt_encode(unsigned char *T, int n, unsigned char *num_indexes, int *indexes);
/* libbsc/bwt/bwt.cpp:283 bsc_bwt_decode */
int orc_bwt_decode(unsigned char *T, int n, int index);

/* libbsc/st/st.cpp:990 bsc_st_encode ; k = 3..8 (7,8 defined by st.cu:99-163) */
int orc_st_encode(unsigned char *T, int n, int k);
/* libbsc/st/st.cpp:1491 bsc_st_decode */
int orc_st_decode(unsigned char *T, int n, int k, int index);

/* libbsc/coder/qlfc/qlfc.cpp:398-455 bsc_qlfc_transform (scalar variant).
 * ranks[0..R) in forward run order, returns R. */
int orc_qlfc_transform(const unsigned char *in, int n, unsigned char *ranks, unsigned char mtf[256]);

/* libbsc/coder/qlfc/qlfc.cpp:2138 / 829 bsc_qlfc_static_encode_block */
int orc_qlfc_static_encode_block(const unsigned char *in, unsigned char *out, int inSize, int outSize);
/* libbsc/coder/qlfc/qlfc.cpp:2186 / 1672 bsc_qlfc_static_decode_block */
int orc_qlfc_static_decode_block(const unsigned char *in, unsigned char *out);

/* libbsc/coder/qlfc/qlfc.cpp:2153 / 463 bsc_qlfc_adaptive_encode_block, 2200 / 1366 bsc_qlfc_adaptive_decode_block (coder id 2) */
int orc_qlfc_adaptive_encode_block(const unsigned char *in, unsigned char *out, int inSize, int outSize);
int orc_qlfc_adaptive_decode_block(const unsigned char *in, unsigned char *out);

/* libbsc/coder/qlfc/qlfc.cpp:2168 / 1135 bsc_qlfc_fast_encode_block, 2213 / 1933 bsc_qlfc_fast_decode_block (coder id 3) */
int orc_qlfc_fast_encode_block(const unsigned char *in, unsigned char *out, int inSize, int outSize);
int orc_qlfc_fast_decode_block(const unsigned char *in, unsigned char *out);

/* libbsc/coder/coder.cpp:70 bsc_coder_split_blocks */
int orc_coder_num_blocks(int n);
void orc_coder_split_blocks(const unsigned char *in, int n, int nBlocks, int *start, int *size);
/* libbsc/coder/coder.cpp:244 bsc_coder_compress (coder 1 static, 2 adaptive, 3 fast) */
int orc_coder_compress(const unsigned char *in, unsigned char *out, int n, int coder, int features);
/* libbsc/coder/coder.cpp:273 bsc_coder_decompress */
int orc_coder_decompress(const unsigned char *in, unsigned char *out, int coder);

/* libbsc/lzp/lzp.cpp:813 bsc_lzp_decompress (decoder only; writes at most outCap bytes) */
int orc_lzp_decompress(const unsigned char *in, unsigned char *out, int n, int outCap, int hashSize, int minLen);

/* libbsc/libbsc/libbsc.cpp:68,213,340,522 (orc_compress: lzpHashSize = lzpMinLen = 0 only; orc_decompress undoes LZP) */
int orc_store(const unsigned char *in, unsigned char *out, int n);
int orc_compress(const unsigned char *in, unsigned char *out, int n, int blockSorter, int coder, int features);
int orc_block_info(const unsigned char *hdr, int hdrSize, int *pBlockSize, int *pDataSize);
int orc_decompress(const unsigned char *in, int inSize, unsigned char *out, int outSize);

#ifdef __cplusplus
}
#endif
#endif
