/*
 * b2_args.h -- plain-C argument blocks shared by the host framing code
 * (blosc_b200.c), the CUDA backend (backend_cuda.cu) and the device kernels.
 */
#ifndef B2_ARGS_H
#define B2_ARGS_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { FILT_SHUFFLE = 0, FILT_UNSHUFFLE = 1, FILT_BITSHUFFLE = 2, FILT_BITUNSHUFFLE = 3 };
enum { B2_CODEC_BLOSCLZ = 0, B2_CODEC_LZ4 = 1, B2_CODEC_ZLIB = 2, B2_CODEC_ZSTD = 3 /* the last two: decode only */ };

typedef struct FilterArgs {
  const uint8_t* src;
  uint8_t* dst;
  long long nbytes;       /* total bytes */
  int blocksize;          /* Blosc block size (last block may be shorter) */
  int typesize;
  int mode;
} FilterArgs;

typedef struct StreamMap {
  long long nbytes;      /* uncompressed size of the whole buffer */
  int blocksize;
  int nsplits;           /* streams per full block */
  int first_block;       /* first selected block (getitem decodes a sub-range) */
  int nfull;             /* number of selected full-size blocks */
  int leftover;          /* bytes of the (selected) short last block, 0 if none */
  int nstreams;          /* nfull*nsplits + (leftover ? 1 : 0) */
} StreamMap;

typedef struct EncodeArgs {
  StreamMap map;
  const uint8_t* in;     /* filtered (or original) bytes, block-major */
  uint8_t* slots;        /* per-stream output slots at the same offsets as `in` */
  int* csizes;           /* [nstreams] compressed size; == stream length means "stored raw" */
  int* needs;            /* [nstreams] smallest `maxout` with which the codec would still have succeeded */
  int codec, clevel, accel, split_flag;
  int table_bytes;       /* shared-memory bytes per warp */
  int* queue;            /* zero-initialised work counter: warps pull stream numbers from it */
} EncodeArgs;

typedef struct ScanArgs {
  const int* csizes;
  const int* needs;
  int* bstarts;          /* [nblocks] out */
  int* result;           /* [0] = total cbytes (clamped to INT_MAX), [1] = fits */
  int nsplits, nfull, has_leftover;
  int blocksize, leftover;
  int serial;            /* 1: reproduce serial_blosc's per-split maxout clamp (blosc.c:646-651); 0: t_blosc's total-fit rule */
  long long destsize;
} ScanArgs;

typedef struct CompactArgs {
  StreamMap map;
  const uint8_t* in;     /* raw splits are copied from here */
  const uint8_t* slots;
  const int* csizes;
  const int* bstarts;
  const int* result;
  uint8_t* dest;
  uint32_t hdr0;         /* version | versionlz<<8 | flags<<16 | typesize<<24 */
  int nbytes32, nblocks;
} CompactArgs;

typedef struct DecodeArgs {
  StreamMap map;
  const uint8_t* chunk;  /* whole compressed chunk */
  int cbytes;            /* header cbytes (bounds for every read) */
  uint8_t* out;          /* uncompressed (still filtered) bytes */
  long long out_shift;   /* subtracted from the buffer offset (getitem decodes into a small scratch) */
  int codec;
  int* status;           /* 0 ok, else min of the negative error codes */
  int* queue;            /* zero-initialised work counter */
} DecodeArgs;

#ifdef __cplusplus
}
#endif
#endif
