/*
 * simt_emu.h -- TEST INFRASTRUCTURE ONLY.
 *
 * A tiny lock-step SIMT emulator so the warp-cooperative device code under
 * c-blosc_b200/csrc/ (kernels_*.cuh) can be compiled with g++ and exercised on a
 * machine without a GPU (this dev container has none).  Every CUDA thread of a CTA
 * is a fiber; warp collectives (__shfl_sync, __ballot_sync, __match_any_sync,
 * __syncwarp) and __syncthreads are rendezvous points resolved by a scheduler.
 * Between rendezvous points the lanes of a warp run one after another (in
 * alternating lane order, to shake out missing __syncwarp()s), so this checks
 * functional correctness and memory safety (build with -fsanitize=address), not
 * performance.  The product never includes this file.
 */
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>

namespace simt {

struct Dim3 {
  unsigned x, y, z;
  Dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};

enum State { RUNNABLE, WAIT_WARP, WAIT_CTA, WAIT_BAR, DONE };
enum Op { OP_NONE, OP_SYNCWARP, OP_SHFL, OP_SHFL_UP, OP_SHFL_DOWN, OP_SHFL_XOR, OP_BALLOT, OP_MATCH_ANY };

struct Thread {
  void* sp;
  char* stack;
  Dim3 tidx;
  unsigned lin;
  State st;
  Op op;
  unsigned mask;
  uint64_t val;
  int arg;
  uint64_t res;
};

extern Thread* cur;
extern Dim3 g_blockIdx, g_blockDim, g_gridDim;
extern unsigned char* g_dynsmem;
extern unsigned long long g_collectives;

uint64_t warp_collective(Op op, unsigned mask, uint64_t val, int arg);
void cta_barrier();
void named_barrier(int id, int count, bool wait);   /* bar.sync / bar.arrive id, count (count = THREADS expected) */
void launch(Dim3 grid, Dim3 block, size_t dynsmem, const std::function<void()>& body);

}  // namespace simt

/* ---- CUDA surface ------------------------------------------------------- */
#define __device__
#define __global__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __restrict__
#define __shared__ static
#define SIMT_EMU 1

#define threadIdx (simt::cur->tidx)
#define blockIdx (simt::g_blockIdx)
#define blockDim (simt::g_blockDim)
#define gridDim (simt::g_gridDim)

struct uint2 { unsigned x, y; };
struct uint4 { unsigned x, y, z, w; };
struct int2 { int x, y; };
static inline uint2 make_uint2(unsigned a, unsigned b) { return uint2{a, b}; }
static inline uint4 make_uint4(unsigned a, unsigned b, unsigned c, unsigned d) { return uint4{a, b, c, d}; }

template <typename T> static inline uint64_t simt_pack(T v) { uint64_t u = 0; memcpy(&u, &v, sizeof(T)); return u; }
template <typename T> static inline T simt_unpack(uint64_t u) { T v; memcpy(&v, &u, sizeof(T)); return v; }

template <typename T> static inline T __shfl_sync(unsigned m, T v, int src, int width = 32) {
  (void)width; return simt_unpack<T>(simt::warp_collective(simt::OP_SHFL, m, simt_pack(v), src));
}
template <typename T> static inline T __shfl_up_sync(unsigned m, T v, unsigned d, int width = 32) {
  (void)width; return simt_unpack<T>(simt::warp_collective(simt::OP_SHFL_UP, m, simt_pack(v), (int)d));
}
template <typename T> static inline T __shfl_down_sync(unsigned m, T v, unsigned d, int width = 32) {
  (void)width; return simt_unpack<T>(simt::warp_collective(simt::OP_SHFL_DOWN, m, simt_pack(v), (int)d));
}
template <typename T> static inline T __shfl_xor_sync(unsigned m, T v, int x, int width = 32) {
  (void)width; return simt_unpack<T>(simt::warp_collective(simt::OP_SHFL_XOR, m, simt_pack(v), x));
}
static inline unsigned __ballot_sync(unsigned m, int pred) {
  return (unsigned)simt::warp_collective(simt::OP_BALLOT, m, pred ? 1 : 0, 0);
}
static inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
static inline int __all_sync(unsigned m, int pred) { return __ballot_sync(m, !pred) == 0; }
template <typename T> static inline unsigned __match_any_sync(unsigned m, T v) {
  return (unsigned)simt::warp_collective(simt::OP_MATCH_ANY, m, simt_pack(v), 0);
}
static inline void __syncwarp(unsigned m = 0xffffffffu) { simt::warp_collective(simt::OP_SYNCWARP, m, 0, 0); }
static inline void __syncthreads() { simt::cta_barrier(); }
static inline void __nanosleep(unsigned) {}
static inline void __threadfence() {}
static inline void __threadfence_block() {}

static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __ffsll(long long v) { return __builtin_ffsll(v); }
static inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz((unsigned)v); }
static inline int __clzll(long long v) { return v == 0 ? 64 : __builtin_clzll((unsigned long long)v); }
static inline unsigned __brev(unsigned v) {
  unsigned r = 0;
  for (int i = 0; i < 32; i++) r |= ((v >> i) & 1u) << (31 - i);
  return r;
}
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned s) {
  uint64_t t = ((uint64_t)b << 32) | a;
  unsigned r = 0;
  for (int i = 0; i < 4; i++) {
    unsigned sel = (s >> (4 * i)) & 0xf;
    unsigned byte = (unsigned)(t >> (8 * (sel & 7))) & 0xff;
    if (sel & 8) byte = (byte & 0x80) ? 0xff : 0x00;
    r |= byte << (8 * i);
  }
  return r;
}
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned sh) {
  uint64_t t = ((uint64_t)hi << 32) | lo;
  return (unsigned)(t >> (sh & 31));
}
static inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned sh) {
  uint64_t t = ((uint64_t)hi << 32) | lo;
  return (unsigned)((t << (sh & 31)) >> 32);
}
template <typename T> static inline T __ldg(const T* p) { return *p; }
template <typename T> static inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
template <typename T> static inline T atomicOr(T* p, T v) { T o = *p; *p = o | v; return o; }
template <typename T> static inline T atomicAnd(T* p, T v) { T o = *p; *p = o & v; return o; }
template <typename T> static inline T atomicMin(T* p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <typename T> static inline T atomicMax(T* p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <typename T> static inline T atomicCAS(T* p, T c, T v) { T o = *p; if (o == c) *p = v; return o; }
template <typename T> static inline T atomicExch(T* p, T v) { T o = *p; *p = v; return o; }
static inline unsigned umin(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned umax(unsigned a, unsigned b) { return a > b ? a : b; }
#ifndef SIMT_NO_MINMAX
template <typename T> static inline T min(T a, T b) { return a < b ? a : b; }
template <typename T> static inline T max(T a, T b) { return a > b ? a : b; }
#endif
