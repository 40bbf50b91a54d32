/*
 * dev_common.cuh -- device-side helpers shared by the sm_100a kernels.
 *
 * Coding model used by every codec routine in this directory: one warp owns one LZ
 * stream and behaves like a scalar processor whose registers are replicated in all
 * 32 lanes (every lane computes the same ip/op/anchor), with explicitly parallel
 * sections (candidate search, match extension, literal / match copies) where the
 * lanes fan out and re-converge through a ballot.  Memory writes in the scalar
 * sections are done by lane 0 only.
 *
 * The same sources compile under g++ against tests/emu/simt_emu.h (a lock-step SIMT
 * emulator used by the CPU test-suite); nothing here depends on that.
 */
#pragma once
#ifdef __CUDACC__
#include <cuda_runtime.h>
#include <stdint.h>
#endif

typedef unsigned char u8;
typedef unsigned short u16;
typedef unsigned int u32;
typedef unsigned long long u64;

#define DEV __device__ __forceinline__
#define FULLMASK 0xffffffffu

/* status codes written by the decode kernels (mirrors blosc_d's returns, blosc.c:761-782) */
#define B2_ERR_BOUNDS (-1)
#define B2_ERR_CODEC (-2)

DEV int lane_id() { return (int)(threadIdx.x & 31u); }

/* Named CTA barriers for warp-specialised producer / consumer hand-offs: `nthreads` threads in total
 * take part (arrivers + waiters); bar_arrive does not block.  Barrier 0 is __syncthreads. */
DEV void bar_sync(int id, int nthreads) {
#ifdef SIMT_EMU
  simt::named_barrier(id, nthreads, true);
#else
  asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(nthreads) : "memory");
#endif
}
DEV void bar_arrive(int id, int nthreads) {
#ifdef SIMT_EMU
  simt::named_barrier(id, nthreads, false);
#else
  asm volatile("bar.arrive %0, %1;" :: "r"(id), "r"(nthreads) : "memory");
#endif
}

/* Unaligned little-endian 32-bit load. On the GPU: two aligned word loads + funnel
 * shift (global/shared loads must be naturally aligned). */
DEV u32 ld_u32(const u8* p) {
#ifdef SIMT_EMU
  u32 v; memcpy(&v, p, 4); return v;
#else
  const uintptr_t a = (uintptr_t)p;
  const u32* w = (const u32*)(a & ~(uintptr_t)3);
  const u32 sh = (u32)(a & 3u) * 8u;
  const u32 lo = w[0];
  if (sh == 0) return lo;
  return __funnelshift_r(lo, w[1], sh);
#endif
}

/* Shared-memory byte access through a 32-bit shared-window address (no generic->shared
 * conversion in the inner loops).  In the emulator a "shared address" is a plain pointer. */
#ifdef SIMT_EMU
typedef u8* smem_addr_t;
DEV smem_addr_t smem_addr(void* p) { return (u8*)p; }
DEV u32 smem_ld_u8(smem_addr_t base, u32 off) { return base[off]; }
DEV void smem_st_u8(smem_addr_t base, u32 off, u32 v) { base[off] = (u8)v; }
DEV u32 smem_ld_u32(smem_addr_t base, u32 off) { u32 v; memcpy(&v, base + off, 4); return v; }   /* off % 4 == 0 */
#else
typedef u32 smem_addr_t;
DEV smem_addr_t smem_addr(void* p) { return (u32)__cvta_generic_to_shared(p); }
DEV u32 smem_ld_u8(smem_addr_t base, u32 off) {
  u32 v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(base + off));
  return v;
}
DEV void smem_st_u8(smem_addr_t base, u32 off, u32 v) {
  asm volatile("st.shared.u8 [%0], %1;" :: "r"(base + off), "r"(v) : "memory");
}
DEV u32 smem_ld_u32(smem_addr_t base, u32 off) {                 /* off % 4 == 0 */
  u32 v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(base + off));
  return v;
}
#endif

/* 8 bytes of shared memory at a 32-bit shared-window address (off % 8 == 0) */
DEV void smem_ld_u32x2(smem_addr_t base, u32 off, u32& x, u32& y) {
#ifdef SIMT_EMU
  memcpy(&x, base + off, 4); memcpy(&y, base + off + 4, 4);
#else
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(x), "=r"(y) : "r"(base + off));
#endif
}

DEV void st_u32_bytes(u8* p, u32 v) {   /* unaligned 32-bit store, byte by byte */
  p[0] = (u8)v; p[1] = (u8)(v >> 8); p[2] = (u8)(v >> 16); p[3] = (u8)(v >> 24);
}

/* number of equal leading bytes (0..4) given x = a ^ b of two little-endian words */
DEV int eq_bytes32(u32 x) { return x ? ((__ffs((int)x) - 1) >> 3) : 4; }

/* Warp-parallel common-prefix length: number of i >= 0 with s[p+i] == s[q+i] and
 * p+i < limit (q < p).  Each lane compares 16 bytes per round; rounds are
 * independent loads, one ballot each.  Uniform result in all lanes. */
DEV int warp_count_match(const u8* __restrict__ s, int p, int q, int limit) {
  const int lane = lane_id();
  int total = 0;
  for (;;) {
    const int i = total + lane * 16;
    const int avail = limit - (p + i);
    int eq = 0;
    if (avail >= 16) {
      const u32 x0 = ld_u32(s + p + i) ^ ld_u32(s + q + i);
      const u32 x1 = ld_u32(s + p + i + 4) ^ ld_u32(s + q + i + 4);
      const u32 x2 = ld_u32(s + p + i + 8) ^ ld_u32(s + q + i + 8);
      const u32 x3 = ld_u32(s + p + i + 12) ^ ld_u32(s + q + i + 12);
      if (x0) eq = eq_bytes32(x0);
      else if (x1) eq = 4 + eq_bytes32(x1);
      else if (x2) eq = 8 + eq_bytes32(x2);
      else eq = 12 + eq_bytes32(x3);
    } else {
      while (eq < avail && s[p + i + eq] == s[q + i + eq]) eq++;
    }
    const unsigned full = __ballot_sync(FULLMASK, eq == 16);
    if (full == FULLMASK) { total += 512; continue; }
    const int fl = __ffs((int)~full) - 1;
    const int e = __shfl_sync(FULLMASK, eq, fl);
    return total + fl * 16 + e;
  }
}

/* Warp-cooperative byte copy between non-overlapping buffers (any alignment). */
DEV void warp_copy_bytes(u8* __restrict__ dst, const u8* __restrict__ src, int n) {
  for (int k = lane_id(); k < n; k += 32) dst[k] = src[k];
}

DEV void warp_fill_bytes(u8* dst, int n, u8 v) {
  for (int k = lane_id(); k < n; k += 32) dst[k] = v;
}

/* Long periodic fill: out[op+k] = out[match + k % off] for k < len when the period `off`
 * is a power of two <= 512 (the byte/bit-shuffled planes Blosc feeds the codecs are full of
 * these: zero planes are off==1 runs, the bench.c planes repeat with period 256).  Each
 * lane's 16-byte phase is then loop invariant: gather it once, store aligned uint4s. */
DEV void warp_fill_period_pow2(u8* out, int op, int match, int len) {
  const int off = op - match;
  const int lane = lane_id();
  int head = (int)((16u - (u32)((uintptr_t)(out + op) & 15u)) & 15u);
  if (head > len) head = len;
  for (int k = lane; k < head; k += 32) out[op + k] = out[match + (k & (off - 1))];
  const int body = (len - head) & ~15;
  if (body > 0) {
    u32 w[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
      u32 v = 0;
#pragma unroll
      for (int t = 0; t < 4; t++) v |= (u32)out[match + ((head + 16 * lane + 4 * q + t) & (off - 1))] << (8 * t);
      w[q] = v;
    }
    const uint4 pat = make_uint4(w[0], w[1], w[2], w[3]);
    uint4* dst = (uint4*)(out + op + head);
    for (int c = lane; c < (body >> 4); c += 32) dst[c] = pat;     /* 512 % off == 0: same phase every round */
  }
  for (int k = head + body + lane; k < len; k += 32) out[op + k] = out[match + (k & (off - 1))];
}

/* LZ77 match copy out[op..op+len) = out[match..], forward semantics with overlap
 * (period = op - match).  Reads only bytes < op, which were written (and made
 * visible with __syncwarp) before this call, so all lanes are independent. */
DEV void warp_copy_match(u8* out, int op, int match, int len) {
  const int off = op - match;
  const int lane = lane_id();
  if (len >= 96 && off <= 512 && (off & (off - 1)) == 0) { warp_fill_period_pow2(out, op, match, len); return; }
  if (off >= len) {
    for (int k = lane; k < len; k += 32) out[op + k] = out[match + k];
  } else if (off == 1) {
    const u8 v = out[match];
    for (int k = lane; k < len; k += 32) out[op + k] = v;
  } else {
    for (int k = lane; k < len; k += 32) out[op + k] = out[match + (k % off)];
  }
}
