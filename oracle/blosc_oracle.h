/*
 * blosc_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C, single-threaded restatement of the c-blosc hot path (filters, BloscLZ,
 * LZ4 block codec, chunk framing).  It exists to check the CUDA product path; it is
 * never linked into, imported by or called from the product (c-blosc_b200/).  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
 * may load it.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_ref.py checks every function below
 * against the unmodified reference compiled from /root/reference (oracle/_ref) and
 * tests/test_oracle_golden.py against the compat .cdata golden chunks.
 *
 * Not restated here: the decode-only zlib / zstd paths.  Their
 * algorithms live in third-party libraries vendored by the reference (zlib 1.3.1, zstd 1.5.6,
 * internal-complibs/); the checker for those is the reference itself -- oracle/_ref is built
 * with both -- plus the reference's golden chunks (tests/test_zlib_decode.py,
 * tests/test_zstd_decode.py).  orc_decompress_ctx() returns -5 for such chunks, like a
 * reference built without those codecs.
 *
 * Every function cites the reference file:line whose behaviour it restates.
 */
#ifndef BLOSC_ORACLE_H
#define BLOSC_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- filters (blosc/shuffle.c:367-443, shuffle-generic.h:32-81, bitshuffle-generic.c) ---- */
void orc_shuffle(size_t typesize, size_t blocksize, const uint8_t* src, uint8_t* dst);
void orc_unshuffle(size_t typesize, size_t blocksize, const uint8_t* src, uint8_t* dst);
int  orc_bitshuffle(size_t typesize, size_t blocksize, const uint8_t* src, uint8_t* dst);
int  orc_bitunshuffle(size_t typesize, size_t blocksize, const uint8_t* src, uint8_t* dst);

/* ---- codecs ---- */
/* blosc/blosclz.c:421-613 */
int orc_blosclz_compress(int clevel, const void* input, int length, void* output,
                         int maxout, int split_block);
/* blosc/blosclz.c:679-789 */
int orc_blosclz_decompress(const void* input, int length, void* output, int maxout);
/* internal-complibs/lz4-1.10.0/lz4.c:1453-1469 -> 1382-1403 -> 930-1338 */
int orc_lz4_compress_fast(const char* src, char* dst, int srcSize, int dstCapacity,
                          int acceleration);
/* internal-complibs/lz4-1.10.0/lz4.c:2451-2456 -> 2022-2445 */
/* the library's opt-in segment-parallel parse, restated (not the reference's byte stream; any LZ4 decoder reads it) */
int orc_lz4_compress_segmented(const char* src, char* dst, int srcSize, int dstCapacity, int acceleration, int seg_bytes, int warm);
void orc_set_lz4_segmented(int seg_bytes, int warm);   /* orc_compress_ctx uses it for LZ4 when seg_bytes != 0 */
int orc_lz4_decompress_safe(const char* src, char* dst, int compressedSize, int dstCapacity);

/* ---- chunk framing (blosc/blosc.c) ---- */
/* blosc.c:962-1060 (+ split_block 929-959 for the FORWARD_COMPAT split mode) */
int32_t orc_compute_blocksize(int compcode, int clevel, int32_t typesize, int32_t nbytes,
                              int32_t forced_blocksize);
/* blosc.c:1282-1308 (serial path: do_job -> serial_blosc 803-867) */
int orc_compress_ctx(int clevel, int doshuffle, size_t typesize, size_t nbytes,
                     const void* src, void* dest, size_t destsize, const char* compressor,
                     size_t blocksize, int numinternalthreads);
/* blosc.c:1520-1535 */
int orc_decompress_ctx(const void* src, void* dest, size_t destsize, int numinternalthreads);
/* blosc.c:1574-1703 */
int orc_getitem(const void* src, int start, int nitems, void* dest);

#ifdef __cplusplus
}
#endif
#endif
