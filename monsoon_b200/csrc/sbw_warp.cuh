// sbw_warp.cuh -- the SPMD vocabulary of the warp-per-game engine (sbw_*.cuh).
//
// One game = one warp.  The game's working set lives in SHARED memory (struct WG, sbw_core.cuh) and all 32 lanes walk
// the rules together:
//   * "uniform" code: every lane executes it with identical values (reads are shared-memory broadcasts, writes store the
//     same value from every lane).  It may never look at the lane index.
//   * "lane" code: written between FOR_LANES(l) ... END_LANES, or as a lambda handed to one of the collectives below
//     (w_ballot / w_or / w_sum / ...).  Lane l works on element l (entity slot, deck slot, tile).  END_LANES is a
//     __syncwarp(): lane-written shared memory is visible to everybody afterwards and the warp is converged again.
// The same source also compiles for the HOST (no __CUDA_ARCH__): uniform code runs once, FOR_LANES is a loop over
// l = 0..31, per-lane values (LV<T>) become arrays.  That build exists for the tests only (tests/wsim): it lets the
// whole engine be checked against the oracle without a GPU, and -- because the lane index does not exist outside
// FOR_LANES there -- the host compiler rejects any use of it in uniform code.
// Inside one FOR_LANES section a lane must not read what another lane writes in the same section (the host build runs
// the lanes one after the other, the GPU runs them at once): move data with two sections and an LV temporary.
#pragma once
#include <stdint.h>

#if defined(__CUDA_ARCH__)
#define SBW __device__
#define SBW_NI __device__ __noinline__
#define SBW_FI __device__ __forceinline__
#define W_FULL 0xFFFFFFFFu
#define FOR_LANES(l) { const int l = (int)(threadIdx.x & 31u);
#define END_LANES } __syncwarp();
template <class T> struct LV { T v; };
#define LVAL(x, l) ((x).v)
#define W_SHARED(p) __builtin_assume(__isShared(p))
#else
#include <math.h>
#include <string.h>
#define SBW static
#define SBW_NI static __attribute__((noinline))
#define SBW_FI static inline
#define FOR_LANES(l) for (int l = 0; l < 32; l++) {
#define END_LANES }
template <class T> struct LV { T v[32]; };
#define LVAL(x, l) ((x).v[l])
#define W_SHARED(p) ((void)0)
#ifndef __align__
#define __align__(n) __attribute__((aligned(n)))
#endif
#endif

typedef signed char i8;
typedef unsigned char u8;
typedef short i16;
typedef unsigned short u16;
typedef unsigned int u32;
typedef unsigned long long u64;

// ---------------------------------------------------------------- scalar intrinsics with host twins
#if defined(__CUDA_ARCH__)
SBW_FI u32 w_mulhi(u32 a, u32 b) { return __umulhi(a, b); }
SBW_FI int w_popc(u32 x) { return __popc(x); }
SBW_FI int w_ffs(u32 x) { return __ffs(x); }
SBW_FI int w_clz(u32 x) { return __clz(x); }
SBW_FI u32 w_brev(u32 x) { return __brev(x); }
SBW_FI double d_add(double a, double b) { return __dadd_rn(a, b); }
SBW_FI double d_sub(double a, double b) { return __dsub_rn(a, b); }
SBW_FI double d_mul(double a, double b) { return __dmul_rn(a, b); }
SBW_FI double d_fma(double a, double b, double c) { return __fma_rn(a, b, c); }
SBW_FI double d_hilo(u32 hi, u32 lo) { return __hiloint2double((int)hi, (int)lo); }
#else
SBW_FI u32 w_mulhi(u32 a, u32 b) { return (u32)(((u64)a * (u64)b) >> 32); }
SBW_FI int w_popc(u32 x) { return __builtin_popcount(x); }
SBW_FI int w_ffs(u32 x) { return __builtin_ffs((int)x); }
SBW_FI int w_clz(u32 x) { return x ? __builtin_clz(x) : 32; }
SBW_FI u32 w_brev(u32 x) { u32 r = 0; for (int i = 0; i < 32; i++) r |= ((x >> i) & 1u) << (31 - i); return r; }
// the host build is compiled with -ffp-contract=off: every operation rounds once, like the _rn intrinsics
SBW_FI double d_add(double a, double b) { return a + b; }
SBW_FI double d_sub(double a, double b) { return a - b; }
SBW_FI double d_mul(double a, double b) { return a * b; }
SBW_FI double d_fma(double a, double b, double c) { return fma(a, b, c); }
SBW_FI double d_hilo(u32 hi, u32 lo) { u64 b = ((u64)hi << 32) | lo; double d; memcpy(&d, &b, 8); return d; }
#endif
// one out-of-line IEEE FP64 division (each inline expansion is ~30 instructions)
#if defined(__CUDA_ARCH__)
SBW_NI double d_div(double a, double b) { return __ddiv_rn(a, b); }
#else
SBW_NI double d_div(double a, double b) { return a / b; }
#endif

// ---------------------------------------------------------------- collectives over a per-lane expression f(lane)
template <class F> SBW_FI u32 w_ballot(F f) {
#if defined(__CUDA_ARCH__)
  return __ballot_sync(W_FULL, f((int)(threadIdx.x & 31u)));
#else
  u32 m = 0;
  for (int l = 0; l < 32; l++) m |= (f(l) ? 1u : 0u) << l;
  return m;
#endif
}
template <class F> SBW_FI u32 w_or(F f) {  // OR of f(lane) over the warp (one REDUX instruction)
#if defined(__CUDA_ARCH__)
  return __reduce_or_sync(W_FULL, (u32)f((int)(threadIdx.x & 31u)));
#else
  u32 m = 0;
  for (int l = 0; l < 32; l++) m |= (u32)f(l);
  return m;
#endif
}
template <class F> SBW_FI int w_sum(F f) {
#if defined(__CUDA_ARCH__)
  return __reduce_add_sync(W_FULL, (int)f((int)(threadIdx.x & 31u)));
#else
  int s = 0;
  for (int l = 0; l < 32; l++) s += (int)f(l);
  return s;
#endif
}
template <class F> SBW_FI int w_min(F f) {
#if defined(__CUDA_ARCH__)
  return __reduce_min_sync(W_FULL, (int)f((int)(threadIdx.x & 31u)));
#else
  int s = 0x7FFFFFFF;
  for (int l = 0; l < 32; l++) { int v = (int)f(l); s = v < s ? v : s; }
  return s;
#endif
}
template <class F> SBW_FI int w_max(F f) {
#if defined(__CUDA_ARCH__)
  return __reduce_max_sync(W_FULL, (int)f((int)(threadIdx.x & 31u)));
#else
  int s = -0x7FFFFFFF - 1;
  for (int l = 0; l < 32; l++) { int v = (int)f(l); s = v > s ? v : s; }
  return s;
#endif
}
// value held by lane `src` (uniform src), as a uniform value
template <class T> SBW_FI T w_bcast(const LV<T>& x, int src) {
#if defined(__CUDA_ARCH__)
  return __shfl_sync(W_FULL, x.v, src);
#else
  return x.v[src];
#endif
}
// inclusive prefix sum over lanes (Hillis-Steele; the host twin adds in the same order, so both builds round alike)
SBW_FI void w_scan_add(LV<double>& x) {
#if defined(__CUDA_ARCH__)
  const int lane = (int)(threadIdx.x & 31u);
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const double o = __shfl_up_sync(W_FULL, x.v, off);
    if (lane >= off) x.v = __dadd_rn(x.v, o);
  }
#else
  for (int off = 1; off < 32; off <<= 1) {
    double o[32];
    for (int l = 0; l < 32; l++) o[l] = l >= off ? x.v[l - off] : 0.0;
    for (int l = off; l < 32; l++) x.v[l] = x.v[l] + o[l];
  }
#endif
}
