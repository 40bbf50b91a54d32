/*
 * dev_lz4.cuh -- LZ4 block codec, one warp per stream, sm_100a.
 *
 * Encoder: bit-exact replay of LZ4_compress_fast's greedy parse
 * (reference: internal-complibs/lz4-1.10.0/lz4.c:930-1338, called from
 * blosc/blosc.c:413-420 with acceleration = 10 - clevel and maxout = neblock, i.e.
 * limitedOutput).  The sequential "probe, skip, probe" search loop (lz4.c:1043-1101)
 * is evaluated 32 probes at a time: lane l computes the position the l-th probe
 * WOULD visit (closed form of the skip schedule), hashes it, and resolves the hash
 * table state it would observe -- the table as left by earlier rounds, overridden by
 * the nearest lower lane with the same hash (__match_any_sync).  The first hitting
 * lane wins, and only probes up to and including it are committed to the table, so
 * the table evolves exactly as in the serial code and the output is byte-identical.
 *
 * Decoder: LZ4_decompress_safe semantics (lz4.c:2022-2445) with warp-wide literal
 * and (period-replicating) match copies.
 */
#pragma once
#include "dev_common.cuh"

#define LZ4_MFLIMIT 12
#define LZ4_LASTLITERALS 5
#define LZ4_TABLE_BYTES 16384          /* lz4.h:163,696  LZ4_MEMORY_USAGE 14 */

template <bool U16>
DEV u32 lz4_hash_at(const u8* __restrict__ s, int pos) {       /* lz4.c:777-806 */
  if (U16) return (ld_u32(s + pos) * 2654435761u) >> (32 - 13);
  const u64 seq = (u64)ld_u32(s + pos) | ((u64)s[pos + 4] << 32);   /* low 5 bytes are all hash5 uses */
  return (u32)(((seq << 24) * 889523592379ull) >> (64 - 12));
}

/* Offset from the search start of the it-th probe of the skip schedule
 * (lz4.c:1043-1053): steps are 1, then accel + (k >> 6) for k = 0,1,2,... */
DEV long long lz4_probe_offset(int it, int accel) {
  if (it == 0) return 0;
  const long long m = it - 1;
  const long long c = m >> 6;
  return 1 + m * accel + 32 * c * (c - 1) + c * (m - 64 * c);
}

/* Returns the compressed size, or 0 when the stream does not fit in `cap`
 * (LZ4_compress_fast's limitedOutput failure).  Uniform across the warp.
 * `tabmem` is LZ4_TABLE_BYTES of shared memory private to this warp. */
template <bool U16>
DEV int lz4_encode_warp(const u8* __restrict__ s, const int n, u8* __restrict__ d, const int cap,
                        const int accel, void* tabmem) {
  const int lane = lane_id();
  u16* tab16 = (u16*)tabmem;
  u32* tab32 = (u32*)tabmem;
#define LZ4_TGET(h) (U16 ? (int)tab16[h] : (int)tab32[h])
#define LZ4_TPUT(h, v) do { if (U16) tab16[h] = (u16)(v); else tab32[h] = (u32)(v); } while (0)

  for (int i = lane; i < LZ4_TABLE_BYTES / 4; i += 32) tab32[i] = 0;   /* LZ4_initStream, lz4.c:1384 */
  __syncwarp();

  const bool limited = (long long)cap < (long long)n + n / 255 + 16;   /* lz4.c:1388,1395 */
  const int olimit = cap;
  const int mfl1 = n - LZ4_MFLIMIT + 1;       /* mflimitPlusOne */
  const int matchlimit = n - LZ4_LASTLITERALS;
  int ip = 1, anchor = 0, op = 0;
  bool more = n >= LZ4_MFLIMIT + 1;           /* lz4.c:1002 */
  /* first byte (lz4.c:1005-1010): table[hash(0)] = 0, which the zeroed table already says */

  while (more) {
    /* ---- find a match: 32 probes per round ---- */
    int match = 0;
    bool found_any = false;
    for (int base_it = 0;; base_it += 32) {
      const int it = base_it + lane;
      const bool valid = ip + lz4_probe_offset(it + 1, accel) <= mfl1;   /* else: `goto _last_literals` (lz4.c:1055) */
      const int pos = valid ? ip + (int)lz4_probe_offset(it, accel) : 0;
      u32 h = 0x80000000u | (u32)lane, seq = 0;
      if (valid) { seq = ld_u32(s + pos); h = lz4_hash_at<U16>(s, pos); }
      const unsigned vmask = __ballot_sync(FULLMASK, valid);
      const unsigned peers = __match_any_sync(FULLMASK, h);
      const unsigned lower = peers & ((1u << lane) - 1u);
      int cand = 0;
      bool hit = false;
      if (valid) {
        cand = lower ? ip + (int)lz4_probe_offset(base_it + (31 - __clz((int)lower)), accel) : LZ4_TGET(h);
        if (U16 || cand + 65535 >= pos) hit = ld_u32(s + cand) == seq;       /* lz4.c:1090-1101 */
      }
      const unsigned found = __ballot_sync(FULLMASK, hit);
      const int nvalid = __popc(vmask);                       /* valid lanes form a prefix */
      const int f = found ? __ffs((int)found) - 1 : 32;
      const int last = f < nvalid - 1 ? f : nvalid - 1;       /* last probe committed to the table */
      if (valid && lane <= last) {
        const unsigned le = last >= 31 ? FULLMASK : ((1u << (last + 1)) - 1u);
        if ((((peers & le) >> lane) >> 1) == 0) LZ4_TPUT(h, pos);  /* highest committed lane per hash wins */
      }
      __syncwarp();
      if (found) {
        ip = __shfl_sync(FULLMASK, pos, f);
        match = __shfl_sync(FULLMASK, cand, f);
        found_any = true;
        break;
      }
      if (nvalid < 32) break;                                  /* ran into the end: last literals */
    }
    if (!found_any) break;

    /* ---- catch up (lz4.c:1107-1109) ---- */
    for (;;) {
      const int a = ip - 1 - lane, b = match - 1 - lane;
      const bool ok = a >= anchor && b >= 0 && s[a] == s[b];
      const unsigned m = __ballot_sync(FULLMASK, ok);
      const int back = m == FULLMASK ? 32 : __ffs((int)~m) - 1;
      ip -= back; match -= back;
      if (back < 32) break;
    }

    /* ---- literals (lz4.c:1112-1136) ---- */
    const int lit = ip - anchor;
    int token = op++;
    if (limited && op + lit + (2 + 1 + LZ4_LASTLITERALS) + lit / 255 > olimit) return 0;
    u32 tokval;
    if (lit >= 15) {
      const int len = lit - 15, nff = len / 255;
      tokval = 15u << 4;
      warp_fill_bytes(d + op, nff, 255);
      if (lane == 0) d[op + nff] = (u8)(len - nff * 255);
      op += nff + 1;
    } else tokval = (u32)lit << 4;
    warp_copy_bytes(d + op, s + anchor, lit);
    op += lit;

    for (;;) {   /* _next_match (lz4.c:1138-1226) */
      const int off = ip - match;
      if (lane == 0) { d[op] = (u8)off; d[op + 1] = (u8)(off >> 8); }
      op += 2;
      int mc = warp_count_match(s, ip + 4, match + 4, matchlimit);
      ip += mc + 4;
      if (limited && op + (1 + LZ4_LASTLITERALS) + (mc + 240) / 255 > olimit) return 0;
      if (mc >= 15) {
        tokval += 15;
        mc -= 15;
        const int nff = mc / 255;
        warp_fill_bytes(d + op, nff, 255);
        if (lane == 0) d[op + nff] = (u8)(mc - nff * 255);
        op += nff + 1;
      } else tokval += (u32)mc;
      if (lane == 0) d[token] = (u8)tokval;

      anchor = ip;
      if (ip >= mfl1) { more = false; break; }                 /* lz4.c:1230-1233 */

      /* fill table at ip-2, then test the next position (lz4.c:1236-1294) */
      const u32 h2 = lz4_hash_at<U16>(s, ip - 2);
      const u32 h = lz4_hash_at<U16>(s, ip);
      int cand = 0;
      if (lane == 0) {
        LZ4_TPUT(h2, ip - 2);
        cand = LZ4_TGET(h);
        LZ4_TPUT(h, ip);
      }
      cand = __shfl_sync(FULLMASK, cand, 0);
      __syncwarp();
      if ((U16 || cand + 65535 >= ip) && ld_u32(s + cand) == ld_u32(s + ip)) {
        token = op++;
        tokval = 0;
        match = cand;
        continue;
      }
      ip++;                                                    /* lz4.c:1298 */
      break;
    }
  }

  /* ---- last literals (lz4.c:1302-1329) ---- */
  const int lastRun = n - anchor;
  if (limited && op + lastRun + 1 + (lastRun + 255 - 15) / 255 > olimit) return 0;
  if (lastRun >= 15) {
    const int acc = lastRun - 15, nff = acc / 255;
    if (lane == 0) d[op] = (u8)(15u << 4);
    op++;
    warp_fill_bytes(d + op, nff, 255);
    if (lane == 0) d[op + nff] = (u8)(acc - nff * 255);
    op += nff + 1;
  } else {
    if (lane == 0) d[op] = (u8)(lastRun << 4);
    op++;
  }
  warp_copy_bytes(d + op, s + anchor, lastRun);
  op += lastRun;
  return op;
#undef LZ4_TGET
#undef LZ4_TPUT
}

/* LZ4_decompress_safe for one stream (lz4.c:2451-2456; safe-loop rules :2234-2436).
 * Returns the number of bytes written or -1.  offset==0 is rejected. */
DEV int lz4_decode_warp(const u8* __restrict__ in, const int csize, u8* out, const int cap) {
  const int iend = csize, oend = cap;
  int ip = 0, op = 0;
  if (cap == 0) return (csize == 1 && in[0] == 0) ? 0 : -1;   /* lz4.c:2062-2066 */
  if (csize == 0) return -1;
  for (;;) {
    const u32 token = in[ip++];
    int len = (int)(token >> 4);
    if (len == 15) {                                          /* read_variable_length(ip, iend-15, 1) */
      u32 sb;
      if (ip >= iend - 15) return -1;
      do {
        sb = in[ip++];
        len += (int)sb;
        if (ip > iend - 15) return -1;
        if (len > oend) return -1;                            /* same verdict as the cpy>oend test below, no int overflow */
      } while (sb == 255);
    }
    int cpy = op + len;
    if (cpy > oend - LZ4_MFLIMIT || ip + len > iend - (2 + 1 + LZ4_LASTLITERALS)) {   /* lz4.c:2289-2331 */
      if (ip + len != iend || cpy > oend) return -1;
      warp_copy_bytes(out + op, in + ip, len);
      op += len;
      break;
    }
    warp_copy_bytes(out + op, in + ip, len);
    ip += len; op = cpy;
    const int off = (int)in[ip] | ((int)in[ip + 1] << 8);
    ip += 2;
    const int match = op - off;
    len = (int)(token & 15u);
    if (len == 15) {                                          /* read_variable_length(ip, iend-4, 0) */
      u32 sb;
      do {
        sb = in[ip++];
        len += (int)sb;
        if (ip > iend - LZ4_LASTLITERALS + 1) return -1;
        if (len > oend) return -1;                            /* keeps `len` from overflowing on hostile input */
      } while (sb == 255);
    }
    len += 4;
    if (match < 0 || off == 0) return -1;                     /* lz4.c:2356 */
    cpy = op + len;
    if (cpy > oend - LZ4_LASTLITERALS) return -1;             /* lz4.c:2423 */
    __syncwarp();                                             /* earlier output must be visible to all lanes */
    warp_copy_match(out, op, match, len);
    __syncwarp();
    op = cpy;
  }
  __syncwarp();
  return op;
}
