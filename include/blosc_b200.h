/*
 * blosc_b200.h -- C ABI of libblosc_b200.so, a B200 (sm_100a) drop-in for the hot path
 * of c-blosc 1.21: the blocked shuffle -> LZ compress / LZ decompress -> unshuffle
 * pipeline behind blosc_compress_ctx() / blosc_decompress_ctx() / blosc_getitem().
 *
 * Every declaration below replaces the reference declaration cited next to it
 * (paths are relative to the reference checkout, blosc/blosc.h) and keeps its name,
 * argument meaning, return codes and on-wire chunk format (README_CHUNK_FORMAT.rst).
 * Differences a caller can observe:
 *   - `src` / `dest` may be host pointers (as in the reference; the library stages them
 *     over PCIe) OR CUDA device pointers (detected with cudaPointerGetAttributes), so a
 *     GPU-resident caller never leaves HBM;
 *   - `numinternalthreads` is validated like the reference but otherwise ignored: the
 *     CUDA grid is the thread pool (reference blosc.c:1706-1949);
 *   - "blosclz" and "lz4" chunks are byte-identical to the reference's; "lz4hc" is accepted and
 *     written in its (= LZ4's, blosc.h:96) format by a hash-chain parser run with LZ4HC's search
 *     effort -- every reference build decodes the chunks, the header is the reference's, the
 *     bytes are not LZ4_compress_HC's; other compressors report -5 exactly like a reference built
 *     with -DDEACTIVATE_ZLIB/ZSTD/SNAPPY (blosc.c:573,1197-1208).  Decoding is wider: zlib and
 *     zstd chunks decode too (serial GPU decoders, one lane per stream); snappy chunks report -5;
 *   - there is no CPU codec: without a CUDA device every compress/decompress call
 *     prints a message on stderr and returns -1.
 */
#ifndef BLOSC_B200_H
#define BLOSC_B200_H

#include <limits.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- constants, blosc.h:20-117 ---- */
#define BLOSC_VERSION_MAJOR 1
#define BLOSC_VERSION_MINOR 21
#define BLOSC_VERSION_RELEASE 7
#define BLOSC_VERSION_STRING "1.21.7.dev-b200"
#define BLOSC_VERSION_DATE "$Date:: 2024-06-24 #$"
#define BLOSC_VERSION_FORMAT 2
#define BLOSC_MIN_HEADER_LENGTH 16
#define BLOSC_MAX_OVERHEAD BLOSC_MIN_HEADER_LENGTH
#define BLOSC_MAX_BUFFERSIZE (INT_MAX - BLOSC_MAX_OVERHEAD)
#define BLOSC_MAX_TYPESIZE 255
#define BLOSC_MAX_BLOCKSIZE ((INT_MAX - BLOSC_MAX_TYPESIZE * sizeof(int32_t)) / 3)
#define BLOSC_MAX_THREADS 256
#define BLOSC_NOSHUFFLE 0
#define BLOSC_SHUFFLE 1
#define BLOSC_BITSHUFFLE 2
#define BLOSC_DOSHUFFLE 0x1
#define BLOSC_MEMCPYED 0x2
#define BLOSC_DOBITSHUFFLE 0x4
#define BLOSC_BLOSCLZ 0
#define BLOSC_LZ4 1
#define BLOSC_LZ4HC 2
#define BLOSC_SNAPPY 3
#define BLOSC_ZLIB 4
#define BLOSC_ZSTD 5
#define BLOSC_BLOSCLZ_COMPNAME "blosclz"
#define BLOSC_LZ4_COMPNAME "lz4"
#define BLOSC_LZ4HC_COMPNAME "lz4hc"
#define BLOSC_SNAPPY_COMPNAME "snappy"
#define BLOSC_ZLIB_COMPNAME "zlib"
#define BLOSC_ZSTD_COMPNAME "zstd"
#define BLOSC_BLOSCLZ_LIB 0
#define BLOSC_LZ4_LIB 1
#define BLOSC_SNAPPY_LIB 2
#define BLOSC_ZLIB_LIB 3
#define BLOSC_ZSTD_LIB 4
#define BLOSC_BLOSCLZ_FORMAT BLOSC_BLOSCLZ_LIB
#define BLOSC_LZ4_FORMAT BLOSC_LZ4_LIB
#define BLOSC_LZ4HC_FORMAT BLOSC_LZ4_LIB
#define BLOSC_ZLIB_FORMAT BLOSC_ZLIB_LIB
#define BLOSC_ZSTD_FORMAT BLOSC_ZSTD_LIB
#define BLOSC_BLOSCLZ_VERSION_FORMAT 1
#define BLOSC_LZ4_VERSION_FORMAT 1
#define BLOSC_ZLIB_VERSION_FORMAT 1
#define BLOSC_ZSTD_VERSION_FORMAT 1
#define BLOSC_ALWAYS_SPLIT 1
#define BLOSC_NEVER_SPLIT 2
#define BLOSC_AUTO_SPLIT 3
#define BLOSC_FORWARD_COMPAT_SPLIT 4

/* ---- the hot path (the drop-in boundary) ---- */

/* replaces blosc.h:245-248 / blosc.c:1282-1308.  >0 compressed bytes; 0 does not fit in
 * destsize / input too large / destsize < 16; -10 bad clevel|doshuffle|typesize; -5 codec not
 * available; -1 internal / no device. */
int blosc_compress_ctx(int clevel, int doshuffle, size_t typesize, size_t nbytes, const void* src,
                       void* dest, size_t destsize, const char* compressor, size_t blocksize,
                       int numinternalthreads);

/* replaces blosc.h:301-302 / blosc.c:1520-1535.  >=0 decompressed bytes; -1 malformed chunk or
 * destsize too small; -5 / -9 unknown codec / codec format version. */
int blosc_decompress_ctx(const void* src, void* dest, size_t destsize, int numinternalthreads);

/* replaces blosc.h:312 / blosc.c:1574-1703.  `start`, `nitems` in elements of the chunk's typesize. */
int blosc_getitem(const void* src, int start, int nitems, void* dest);

/* ---- global-state front-end over the ctx path (blosc.h:127-222,256-299,321-352) ---- */
void blosc_init(void);                                                   /* blosc.h:142 */
void blosc_destroy(void);                                                /* blosc.h:151 */
int  blosc_compress(int clevel, int doshuffle, size_t typesize, size_t nbytes, const void* src,
                    void* dest, size_t destsize);                        /* blosc.h:221-222 */
int  blosc_decompress(const void* src, void* dest, size_t destsize);     /* blosc.h:279 */
int  blosc_get_nthreads(void);                                           /* blosc.h:321 */
int  blosc_set_nthreads(int nthreads);                                   /* blosc.h:332 */
const char* blosc_get_compressor(void);                                  /* blosc.h:338 */
int  blosc_set_compressor(const char* compname);                         /* blosc.h:352 */
int  blosc_get_blocksize(void);                                          /* blosc.h:494 */
void blosc_set_blocksize(size_t blocksize);                              /* blosc.h:506 */
void blosc_set_splitmode(int splitmode);                                 /* blosc.h:527 */
int  blosc_free_resources(void);                                         /* blosc.h:411 */

/* ---- names / introspection (host-only header readers) ---- */
int  blosc_compcode_to_compname(int compcode, const char** compname);    /* blosc.h:364 */
int  blosc_compname_to_compcode(const char* compname);                   /* blosc.h:374 */
const char* blosc_list_compressors(void);                                /* blosc.h:388 */
const char* blosc_get_version_string(void);                              /* blosc.h:396 */
int  blosc_get_complib_info(const char* compname, char** complib, char** version);   /* blosc.h:417 */
void blosc_cbuffer_sizes(const void* cbuffer, size_t* nbytes, size_t* cbytes, size_t* blocksize);  /* blosc.h:431 */
int  blosc_cbuffer_validate(const void* cbuffer, size_t cbytes, size_t* nbytes);                   /* blosc.h:441 */
void blosc_cbuffer_metainfo(const void* cbuffer, size_t* typesize, int* flags);                    /* blosc.h:459 */
void blosc_cbuffer_versions(const void* cbuffer, int* version, int* compversion);                  /* blosc.h:468 */
const char* blosc_cbuffer_complib(const void* cbuffer);                                            /* blosc.h:477 */

/* ---- B200 extensions (not in the reference) ---- */

/* One filter over one block, the GPU counterpart of the reference's internal
 * blosc_internal_{shuffle,unshuffle,bitshuffle,bitunshuffle} (blosc/shuffle.c:367-443) that
 * its unit tests call.  mode: 0 shuffle, 1 unshuffle, 2 bitshuffle, 3 bitunshuffle.
 * src/dest host or device.  Returns 0, or -1 on device failure. */
int blosc_b200_filter(int mode, size_t typesize, size_t blocksize, const void* src, void* dest);

/* Frames: buffers larger than one chunk (a Blosc-1 chunk holds at most BLOSC_MAX_BUFFERSIZE
 * bytes, blosc.h:40).  The buffer is cut into `chunksize`-byte pieces (0 = 256 MiB; rounded down
 * to a multiple of typesize), each compressed exactly as blosc_compress_ctx() would with
 * destsize = piece + 16, several in flight at once so that PCIe transfers overlap the kernels.
 * The result is a 32-byte header + u64 offset table + ordinary Blosc-1 chunks (layout in
 * blosc_b200.c); blosc_b200_frame_chunk() locates chunk i so that any Blosc-1 library can decode
 * it.  src/dest/frame may be host or device memory.  Returns: compress -> frame bytes, 0 if it
 * does not fit in destsize (always fits in blosc_b200_frame_bound()), <0 like blosc_compress_ctx;
 * decompress -> nbytes or -1; getitem -> bytes copied or <0 (items may span chunks). */
size_t    blosc_b200_frame_bound(size_t nbytes, size_t typesize, size_t chunksize);
long long blosc_b200_frame_compress(int clevel, int doshuffle, size_t typesize, size_t nbytes, const void* src,
                                    void* dest, size_t destsize, const char* compressor, size_t blocksize,
                                    size_t chunksize, int numinternalthreads);
long long blosc_b200_frame_decompress(const void* frame, size_t framesize, void* dest, size_t destsize,
                                      int numinternalthreads);
long long blosc_b200_frame_getitem(const void* frame, size_t framesize, size_t start, size_t nitems, void* dest);
int       blosc_b200_frame_info(const void* frame, size_t framesize, size_t* nbytes, size_t* cbytes,
                                size_t* chunksize, size_t* nchunks);
long long blosc_b200_frame_chunk(const void* frame, size_t framesize, size_t i, size_t* chunk_cbytes);

/* Select the CUDA device used by the calling thread's subsequent calls with HOST pointers
 * (device pointers carry their device).  Multi-GPU callers run one process (or thread) per GPU. */
int blosc_b200_set_device(int dev);

/* Per-kernel CUDA-event timing of the calls made since the last reset (bench.py's roofline
 * leg).  kind: 0 filter, 1 encode, 2 scan, 3 compact, 4 decode, 5 unfilter. */
void blosc_b200_set_profiling(int on);
void blosc_b200_prof_reset(void);
int  blosc_b200_prof_get(int kind, double* ms_total, long long* launches);
long long blosc_b200_launch_count(void);      /* kernels launched by this library so far */

#ifdef __cplusplus
}
#endif
#endif
