/*
 * dev_filters.cuh -- byte-shuffle / bit-shuffle filters and their inverses, sm_100a.
 *
 * Semantics (reference): blosc/shuffle-generic.h:32-81 (byte transpose
 * dst[j*N+i] = src[i*ts+j], tail copied), blosc/shuffle.c:393-443 +
 * blosc/bitshuffle-generic.c:125-139,208-220 (bit-plane transpose, only when the
 * element count is a multiple of 8, otherwise the block is copied).
 *
 * The filters operate per Blosc block (the last block may be shorter).  Work item =
 * (block, tile of 1024 elements); one warp per item:
 *   fast path  (typesize 1/2/4/8/16, 16-byte aligned block bases, N % 32 == 0):
 *     each lane moves 4 elements with 128-bit coalesced accesses and transposes
 *     them in registers with PRMT (__byte_perm); every plane / bit-row store of the
 *     warp is one contiguous 128-byte line.  The bit variants stage the byte planes
 *     of the tile in warp-private shared memory and run 8x8 bit transposes on them.
 *   generic path (any typesize / alignment / remainder): byte-granular loops.
 * HBM traffic is the algorithmic minimum: every byte is read once and written once.
 */
#pragma once
#include "b2_args.h"
#include "dev_common.cuh"

#define FILT_TILE 1024            /* elements per warp work item */

/* 4x4 byte transpose of the words (a,b,c,d): a'=(a0,b0,c0,d0) b'=(a1,b1,c1,d1) ... */
DEV void bt4x4(u32& a, u32& b, u32& c, u32& d) {
  const u32 t0 = __byte_perm(a, b, 0x5140), t1 = __byte_perm(c, d, 0x5140);
  const u32 t2 = __byte_perm(a, b, 0x7362), t3 = __byte_perm(c, d, 0x7362);
  a = __byte_perm(t0, t1, 0x5410);
  b = __byte_perm(t0, t1, 0x7632);
  c = __byte_perm(t2, t3, 0x5410);
  d = __byte_perm(t2, t3, 0x7632);
}

/* 8x8 bit transpose of the 64-bit value (lo | hi<<32): output byte k collects bit k
 * of every input byte, input byte m landing in bit m (bitshuffle-generic.h:42-49). */
DEV void bit8x8(u32& lo, u32& hi) {
  u64 x = ((u64)hi << 32) | lo, t;
  t = (x ^ (x >> 7)) & 0x00AA00AA00AA00AAull;  x = x ^ t ^ (t << 7);
  t = (x ^ (x >> 14)) & 0x0000CCCC0000CCCCull; x = x ^ t ^ (t << 14);
  t = (x ^ (x >> 28)) & 0x00000000F0F0F0F0ull; x = x ^ t ^ (t << 28);
  lo = (u32)x; hi = (u32)(x >> 32);
}

/* Load 4 consecutive elements (4*TS bytes at p, aligned to min(16,4*TS)) and return
 * the TS plane words: pl[j] = byte j of the four elements. */
template <int TS>
DEV void load4_to_planes(const u8* __restrict__ p, u32 (&pl)[TS]) {
  if (TS == 1) {
    pl[0] = *(const u32*)p;
  } else if (TS == 2) {
    const uint2 v = *(const uint2*)p;
    pl[0] = __byte_perm(v.x, v.y, 0x6420);
    pl[1] = __byte_perm(v.x, v.y, 0x7531);
  } else {
    constexpr int WPE = TS / 4;                 /* words per element */
    u32 w[4 * (TS >= 4 ? TS / 4 : 1)];
#pragma unroll
    for (int q = 0; q < TS / 4; q++) {          /* TS/4 x 128-bit loads */
      const uint4 v = ((const uint4*)p)[q];
      w[4 * q] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w;
    }
#pragma unroll
    for (int k = 0; k < WPE; k++) {
      u32 a = w[k], b = w[WPE + k], c = w[2 * WPE + k], d = w[3 * WPE + k];
      bt4x4(a, b, c, d);
      pl[(4 * k) % TS] = a; pl[(4 * k + 1) % TS] = b; pl[(4 * k + 2) % TS] = c; pl[(4 * k + 3) % TS] = d;
    }
  }
}

/* Inverse: from TS plane words rebuild 4 consecutive elements and store them. */
template <int TS>
DEV void planes_to_store4(u8* __restrict__ p, const u32 (&pl)[TS]) {
  if (TS == 1) {
    *(u32*)p = pl[0];
  } else if (TS == 2) {
    *(uint2*)p = make_uint2(__byte_perm(pl[0], pl[1], 0x5140), __byte_perm(pl[0], pl[1], 0x7362));
  } else {
    constexpr int WPE = TS / 4;
    u32 w[4 * (TS >= 4 ? TS / 4 : 1)];
#pragma unroll
    for (int k = 0; k < WPE; k++) {
      u32 a = pl[(4 * k) % TS], b = pl[(4 * k + 1) % TS], c = pl[(4 * k + 2) % TS], d = pl[(4 * k + 3) % TS];
      bt4x4(a, b, c, d);
      w[k] = a; w[WPE + k] = b; w[2 * WPE + k] = c; w[3 * WPE + k] = d;
    }
#pragma unroll
    for (int q = 0; q < TS / 4; q++) ((uint4*)p)[q] = make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
  }
}

/* ---- fast tile paths: a full tile of FILT_TILE elements starting at element e0 ----
 * All the loads of a batch of iterations are issued before the first transpose / store, so that a lane has
 * 64-128 bytes in flight whatever the typesize (the copy is latency-bound otherwise: 8 bytes per lane and
 * iteration at typesize 2). */
template <int TS>
DEV void tile_shuffle(const u8* __restrict__ s, u8* __restrict__ d, int N, int e0) {
  const int lane = lane_id();
  constexpr int B = TS <= 4 ? 8 : (TS == 8 ? 4 : 2);          /* iterations per batch */
#pragma unroll 1
  for (int q0 = 0; q0 < FILT_TILE / 128; q0 += B) {
    u32 pl[B][TS];
#pragma unroll
    for (int q = 0; q < B; q++) load4_to_planes<TS>(s + (long long)(e0 + (q0 + q) * 128 + 4 * lane) * TS, pl[q]);
#pragma unroll
    for (int q = 0; q < B; q++) {
      const int e = e0 + (q0 + q) * 128 + 4 * lane;
#pragma unroll
      for (int j = 0; j < TS; j++) *(u32*)(d + (long long)j * N + e) = pl[q][j];
    }
  }
}

template <int TS>
DEV void tile_unshuffle(const u8* __restrict__ s, u8* __restrict__ d, int N, int e0) {
  const int lane = lane_id();
  constexpr int B = TS <= 4 ? 8 : 2;
#pragma unroll 1
  for (int q0 = 0; q0 < FILT_TILE / 128; q0 += B) {
    u32 pl[B][TS];
#pragma unroll
    for (int q = 0; q < B; q++) {
      const int e = e0 + (q0 + q) * 128 + 4 * lane;
#pragma unroll
      for (int j = 0; j < TS; j++) pl[q][j] = *(const u32*)(s + (long long)j * N + e);
    }
#pragma unroll
    for (int q = 0; q < B; q++) planes_to_store4<TS>(d + (long long)(e0 + (q0 + q) * 128 + 4 * lane) * TS, pl[q]);
  }
}

/* sm: warp-private TS*FILT_TILE bytes, plane j at sm + j*FILT_TILE */
template <int TS>
DEV void tile_bitshuffle(const u8* __restrict__ s, u8* __restrict__ d, int N, int e0, u8* sm) {
  const int lane = lane_id();
  const int rowlen = N >> 3;
#pragma unroll 4
  for (int q = 0; q < FILT_TILE / 128; q++) {
    const int el = q * 128 + 4 * lane;
    u32 pl[TS];
    load4_to_planes<TS>(s + (long long)(e0 + el) * TS, pl);
#pragma unroll
    for (int j = 0; j < TS; j++) *(u32*)(sm + j * FILT_TILE + el) = pl[j];
  }
  __syncwarp();
  for (int j = 0; j < TS; j++) {
    const uint4 v0 = *(const uint4*)(sm + j * FILT_TILE + 32 * lane);
    const uint4 v1 = *(const uint4*)(sm + j * FILT_TILE + 32 * lane + 16);
    u32 lo0 = v0.x, hi0 = v0.y, lo1 = v0.z, hi1 = v0.w, lo2 = v1.x, hi2 = v1.y, lo3 = v1.z, hi3 = v1.w;
    bit8x8(lo0, hi0); bit8x8(lo1, hi1); bit8x8(lo2, hi2); bit8x8(lo3, hi3);
    bt4x4(lo0, lo1, lo2, lo3);      /* lo_k = byte k of the four groups -> bit-rows 0..3 */
    bt4x4(hi0, hi1, hi2, hi3);      /* bit-rows 4..7 */
    u8* row = d + (long long)(8 * j) * rowlen + (e0 >> 3) + 4 * lane;
    *(u32*)(row) = lo0;
    *(u32*)(row + (long long)rowlen) = lo1;
    *(u32*)(row + 2ll * rowlen) = lo2;
    *(u32*)(row + 3ll * rowlen) = lo3;
    *(u32*)(row + 4ll * rowlen) = hi0;
    *(u32*)(row + 5ll * rowlen) = hi1;
    *(u32*)(row + 6ll * rowlen) = hi2;
    *(u32*)(row + 7ll * rowlen) = hi3;
  }
  __syncwarp();
}

template <int TS>
DEV void tile_bitunshuffle(const u8* __restrict__ s, u8* __restrict__ d, int N, int e0, u8* sm) {
  const int lane = lane_id();
  const int rowlen = N >> 3;
#pragma unroll 2
  for (int j = 0; j < TS; j++) {
    const u8* row = s + (long long)(8 * j) * rowlen + (e0 >> 3) + 4 * lane;
    u32 lo0 = *(const u32*)(row);
    u32 lo1 = *(const u32*)(row + (long long)rowlen);
    u32 lo2 = *(const u32*)(row + 2ll * rowlen);
    u32 lo3 = *(const u32*)(row + 3ll * rowlen);
    u32 hi0 = *(const u32*)(row + 4ll * rowlen);
    u32 hi1 = *(const u32*)(row + 5ll * rowlen);
    u32 hi2 = *(const u32*)(row + 6ll * rowlen);
    u32 hi3 = *(const u32*)(row + 7ll * rowlen);
    bt4x4(lo0, lo1, lo2, lo3);      /* back to per-group words */
    bt4x4(hi0, hi1, hi2, hi3);
    bit8x8(lo0, hi0); bit8x8(lo1, hi1); bit8x8(lo2, hi2); bit8x8(lo3, hi3);
    *(uint4*)(sm + j * FILT_TILE + 32 * lane) = make_uint4(lo0, hi0, lo1, hi1);
    *(uint4*)(sm + j * FILT_TILE + 32 * lane + 16) = make_uint4(lo2, hi2, lo3, hi3);
  }
  __syncwarp();
  for (int q = 0; q < FILT_TILE / 128; q++) {
    const int el = q * 128 + 4 * lane;
    u32 pl[TS];
#pragma unroll
    for (int j = 0; j < TS; j++) pl[j] = *(const u32*)(sm + j * FILT_TILE + el);
    planes_to_store4<TS>(d + (long long)(e0 + el) * TS, pl);
  }
  __syncwarp();
}

/* ---- generic (any typesize / alignment) element-range paths ---- */
DEV void range_generic(int mode, const u8* __restrict__ s, u8* __restrict__ d, int ts, int N, int e0, int e1) {
  const int lane = lane_id();
  const int ne = e1 - e0;
  if (mode == FILT_SHUFFLE) {
    for (int j = 0; j < ts; j++)
      for (int i = lane; i < ne; i += 32) d[(long long)j * N + e0 + i] = s[(long long)(e0 + i) * ts + j];
  } else if (mode == FILT_UNSHUFFLE) {
    const long long nb = (long long)ne * ts;
    for (long long o = lane; o < nb; o += 32) {
      const int i = (int)(o / ts), j = (int)(o - (long long)i * ts);
      d[(long long)e0 * ts + o] = s[(long long)j * N + e0 + i];
    }
  } else if (mode == FILT_BITSHUFFLE) {          /* e0, e1 multiples of 8 */
    const int rowlen = N >> 3, nb8 = ne >> 3;
    for (int r = 0; r < 8 * ts; r++) {
      const int b = r >> 3, k = r & 7;
      for (int i = lane; i < nb8; i += 32) {
        const u8* p = s + (long long)(e0 + 8 * i) * ts + b;
        u32 v = 0;
#pragma unroll
        for (int m = 0; m < 8; m++) v |= ((p[(long long)m * ts] >> k) & 1u) << m;
        d[(long long)r * rowlen + (e0 >> 3) + i] = (u8)v;
      }
    }
  } else {                                       /* FILT_BITUNSHUFFLE */
    const int rowlen = N >> 3;
    const long long nb = (long long)ne * ts;
    for (long long o = lane; o < nb; o += 32) {
      const int i = (int)(o / ts), b = (int)(o - (long long)i * ts);
      const int e = e0 + i;
      const u8* p = s + (long long)(8 * b) * rowlen + (e >> 3);
      u32 v = 0;
#pragma unroll
      for (int k = 0; k < 8; k++) v |= ((p[(long long)k * rowlen] >> (e & 7)) & 1u) << k;
      d[(long long)e * ts + b] = (u8)v;
    }
  }
}

template <int TS>
DEV void tile_fast(int mode, const u8* __restrict__ s, u8* __restrict__ d, int N, int e0, u8* sm) {
  if (mode == FILT_SHUFFLE) tile_shuffle<TS>(s, d, N, e0);
  else if (mode == FILT_UNSHUFFLE) tile_unshuffle<TS>(s, d, N, e0);
  else if (mode == FILT_BITSHUFFLE) tile_bitshuffle<TS>(s, d, N, e0, sm);
  else tile_bitunshuffle<TS>(s, d, N, e0, sm);
}

/* One warp work item: tile `t` of block whose (already offset) bases are s/d.
 * t == ntiles_full_block is the "tail" item (bytes past N*ts).  TSK > 0: the kernel was instantiated for this
 * typesize (the register count of the typesize-16 transposes must not cap the occupancy of the typesize-4 ones). */
template <int TSK>
DEV void filter_item(int mode, const u8* __restrict__ s, u8* __restrict__ d, int ts, int bsize, int t,
                     int tail_item, u8* sm) {
  const int N = bsize / ts;
  const bool bitmode = mode >= FILT_BITSHUFFLE;
  if (t == tail_item) {
    const long long off = (long long)N * ts;
    warp_copy_bytes(d + off, s + off, (int)(bsize - off));
    return;
  }
  const int e0 = t * FILT_TILE;
  if (e0 >= N) return;
  const int e1 = e0 + FILT_TILE < N ? e0 + FILT_TILE : N;
  if (bitmode && (N & 7)) {                       /* shuffle.c:412-415: block is copied */
    const long long o0 = (long long)e0 * ts, o1 = (long long)e1 * ts;
    for (long long o = o0 + lane_id(); o < o1; o += 32) d[o] = s[o];
    return;
  }
  const bool aligned = ((((uintptr_t)s) | ((uintptr_t)d)) & 15u) == 0 && (N & 31) == 0;
  if (aligned && e1 - e0 == FILT_TILE) {
    if (TSK > 0) { tile_fast<TSK == 0 ? 1 : TSK>(mode, s, d, N, e0, sm); return; }
    switch (ts) {
      case 1: if (bitmode) { tile_fast<1>(mode, s, d, N, e0, sm); return; } break;
      case 2: tile_fast<2>(mode, s, d, N, e0, sm); return;
      case 4: tile_fast<4>(mode, s, d, N, e0, sm); return;
      case 8: tile_fast<8>(mode, s, d, N, e0, sm); return;
      case 16: tile_fast<16>(mode, s, d, N, e0, sm); return;
      default: break;
    }
  }
  range_generic(mode, s, d, ts, N, e0, e1);
}

#define FILT_WARPS 4
/* dynamic shared memory: FILT_WARPS * 16 * FILT_TILE bytes for the bit modes (0 otherwise) */
template <int TSK>
__global__ void __launch_bounds__(FILT_WARPS * 32, TSK == 0 ? 4 : 8) filter_kernel(FilterArgs a) {
#ifdef SIMT_EMU
  u8* smem = simt::g_dynsmem;
#else
  extern __shared__ __align__(16) u8 smem[];
#endif
  const int warp = (int)(threadIdx.x >> 5);
  u8* sm = smem + (size_t)warp * (16 * FILT_TILE);
  const long long nblocks = (a.nbytes + a.blocksize - 1) / a.blocksize;
  const int tiles_per_block = (a.blocksize / a.typesize + FILT_TILE - 1) / FILT_TILE;
  const int ipb = tiles_per_block + 1;            /* + tail item */
  const long long nitems = nblocks * ipb;
  for (long long it = (long long)blockIdx.x * FILT_WARPS + warp; it < nitems; it += (long long)gridDim.x * FILT_WARPS) {
    long long b;
    int t;
    if (nitems < 0x7fffffffll) { const unsigned ui = (unsigned)it; b = ui / (unsigned)ipb; t = (int)(ui - (unsigned)b * (unsigned)ipb); }
    else { b = it / ipb; t = (int)(it - b * ipb); }
    const long long b0 = b * a.blocksize;
    const long long rem = a.nbytes - b0;
    const int bsize = rem < a.blocksize ? (int)rem : a.blocksize;
    filter_item<TSK>(a.mode, a.src + b0, a.dst + b0, a.typesize, bsize, t, tiles_per_block, sm);
  }
}
