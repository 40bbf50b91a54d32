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

/* the per-workspace int32 words behind `result` / `queue` / `done` / `status` (zero between calls: the
 * last warp of every encode / decode launch puts the counters back) */
enum { B2_R_CBYTES = 0, B2_R_FITS = 1, B2_R_STATUS = 2, B2_R_QUEUE = 3, B2_R_DONE = 4, B2_R_STATUS_OUT = 5, B2_R_WORDS = 16 };
#define B2_FOLD_SCAN_MAX_BLOCKS 65536   /* above this the 1024-thread scan_kernel is launched instead of the in-kernel scan */

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

typedef struct EncodeArgs {
  StreamMap map;
  const uint8_t* in;     /* filtered (or original) bytes, block-major */
  uint8_t* slots;        /* per-stream output slots at the same offsets as `in` */
  int* csizes;           /* [nstreams] compressed size; == stream length means "stored raw" */
  int* needs;            /* [nstreams] smallest `maxout` with which the codec would still have succeeded */
  int codec, clevel, accel, split_flag;
  int table_bytes;       /* shared-memory bytes per warp */
  int num_sms;           /* filled in by the backend (team kernel: spreads the walker warps over the SM sub-partitions) */
  int many;              /* other chunks are in flight on this device (frames): streams per SM matter more than latency */
  int* queue;            /* work counter: warps pull stream numbers from it.  It only ever counts up: a launch
                          * adds exactly nstreams + (warps launched) tickets, the backend keeps the running base */
  unsigned queue_base;   /* first ticket of this launch (filled in by the backend) */
  unsigned* queue_base_host;   /* HOST word behind it, owned by the workspace */
  int* done;             /* zero-initialised count of finished streams: the warp that completes it runs the scan */
  int fold_scan;         /* 1: that warp computes bstarts / cbytes / the fit verdict (scan) in this launch */
  ScanArgs scan;
} EncodeArgs;

/* ---- segment-parallel LZ4 parse (dev_lz4fast.cuh) ---- */
#ifndef B2_FAST_SEG
#define B2_FAST_SEG 256     /* bytes per segment (dev_lz4fast.cuh FAST_SEG) */
#endif
#define B2_FAST_WIN_MAX (64 * 1024)   /* bytes of a stream that one parse CTA keeps in shared memory */
typedef struct FastSeg {   /* one record per segment of FAST_SEG bytes */
  uint16_t nbytes;         /* bytes in the segment's slot (0: no match, the segment is all literals) */
  uint16_t l1;             /* literals in front of the segment's first match */
  uint16_t tail;           /* literals after its last match (the whole segment when nbytes == 0) */
  uint16_t lt;             /* slot position of the last sequence's token */
  uint16_t lm, lo;         /* length and offset of the last match (its length bytes are not in the slot) */
  uint16_t pad0, pad1;
  uint32_t dst;            /* by the stream scan: offset of the merged output inside the stream's LZ4 block,
                            * 0xffffffff when the segment only continues the previous segment's last match */
  uint32_t pin;            /* by the stream scan: literals pending in front of this segment */
  uint32_t run;            /* by the stream scan: final length of the last match (with the segments it swallowed) */
  uint32_t pad2;
} FastSeg;

typedef struct FastArgs {
  StreamMap map;
  const uint8_t* in;       /* filtered (or original) bytes, block-major */
  uint8_t* slots;
  uint16_t* prev;          /* [nbytes] hash-chain index written by index_kernel */
  FastSeg* segs;           /* stream-major: stream idx starts at idx * segs_full (the leftover stream comes last) */
  int* seg_done;           /* [nstreams] zero-initialised count of parsed groups; the warp that completes a stream scans it */
  int* ptail;              /* [nstreams] literals after the stream's last match */
  int* csizes;
  int* needs;
  int segs_full, segs_left;        /* segments per full stream / of the leftover stream */
  int win_bytes;                   /* bytes per parse window (one CTA): multiple of 32 segments */
  int threads;                     /* threads of a parse CTA (<= segments per window; they draw segments from a counter) */
  int groups_full, groups_left;    /* windows per full stream / of the leftover stream */
  int depth, accel;
  int hash_mask;                   /* 0xffff: chains over 6-byte hashes (lz4); 0: over 4-byte hashes (lz4hc) */
  int lazy;                        /* matches shorter than this are weighed against the next position's match (lz4hc) */
  int* queue;
  unsigned queue_base;
  unsigned* queue_base_host;
  int* done;
  int fold_scan;
  ScanArgs scan;
} FastArgs;

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
  const FastSeg* segs;   /* non-NULL: the compressed streams are fast-parsed segments to be stitched (dev_lz4fast.cuh) */
  const int* ptail;
  int segs_full, segs_left;
} CompactArgs;

typedef struct DecodeArgs {
  StreamMap map;
  const uint8_t* chunk;  /* whole compressed chunk */
  int cbytes;            /* header cbytes (bounds for every read) */
  uint8_t* out;          /* uncompressed (still filtered) bytes */
  long long out_shift;   /* subtracted from the buffer offset (getitem decodes into a small scratch) */
  int codec;
  int* status;           /* zero-initialised accumulator: 0 ok, else min of the negative error codes */
  int* queue;            /* work counter, see EncodeArgs */
  unsigned queue_base;
  unsigned* queue_base_host;
  int* done;             /* zero-initialised count of finished streams */
  int* status_out;       /* the last warp publishes the verdict here and zeroes the three words above */
  int many;              /* other calls are running on this device: streams per SM matter more than latency */
} DecodeArgs;

#ifdef __cplusplus
}
#endif
#endif
