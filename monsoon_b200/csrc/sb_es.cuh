// (mu + lambda) evolution-strategy operators on the device -- SURVEY 8(f) row f1.
//   WeightVector.mutate                 evo/weights.py:20-40
//   Population.generate_offspring       evo/population.py:75-90
//   Population.select_from_combined     evo/population.py:92-176 (top-mu, sigma reset, diversity injection)
// The population (weights, sigmas: f64 [rows][nf]) stays resident in HBM between generations; one thread owns one
// row and draws from that row's counter stream philox(counter=(draw, row, tag, generation), key=seed), so the
// kernels need no communication.  exp / log are written with correctly rounded basic operations only
// (__dmul_rn / __dadd_rn / __ddiv_rn / __dsqrt_rn: no FMA contraction, no libdevice), so the results are
// bit-identical to the CPU oracle (oracle/sb_oracle_es.c).
#pragma once
#include "sb_engine.cuh"

#define ES_TAG_OFFSPRING 0xE5u
#define ES_TAG_RESET 0xE6u
#define ES_TAG_INJECT 0xE7u
#define ES_MAX_FEATURES 64
#define ES_MAX_MU 4096

SBD_FI double es_ldexp(double x, int e) { return __dmul_rn(x, __longlong_as_double((long long)(1023 + e) << 52)); }
SBD double es_log(double x) {  // x > 0, normal
  unsigned long long u = (unsigned long long)__double_as_longlong(x);
  int e = (int)((u >> 52) & 0x7FF) - 1022;  // x = m * 2^e, m in [0.5, 1)
  double m = __longlong_as_double((long long)((u & 0x000FFFFFFFFFFFFFull) | 0x3FE0000000000000ull));
  if (m < 0.70710678118654752440) { m = __dmul_rn(m, 2.0); e -= 1; }
  const double z = __ddiv_rn(__dadd_rn(m, -1.0), __dadd_rn(m, 1.0)), z2 = __dmul_rn(z, z);  // log m = 2 atanh z
  double p = 1.0 / 27.0;
  #pragma unroll 1
  for (int k = 25; k >= 1; k -= 2) p = __dadd_rn(__dmul_rn(p, z2), __ddiv_rn(1.0, (double)k));
  return __dadd_rn(__dmul_rn((double)e, 0.693147180559945309417232), __dmul_rn(__dmul_rn(2.0, z), p));
}
SBD double es_exp(double y) {  // |y| < 700
  const double inv_fact[13] = {1.0, 1.0, 0.5, 1.0 / 6, 1.0 / 24, 1.0 / 120, 1.0 / 720, 1.0 / 5040, 1.0 / 40320,
                               1.0 / 362880, 1.0 / 3628800, 1.0 / 39916800, 1.0 / 479001600};
  const double kf = __dmul_rn(y, 1.44269504088896340736);
  const int k = (int)__dadd_rn(kf, kf >= 0 ? 0.5 : -0.5);
  const double r = __dadd_rn(__dadd_rn(y, -__dmul_rn((double)k, 0.693147180369123816490)), -__dmul_rn((double)k, 1.90821492927058770002e-10));
  double p = 1.0 / 6227020800.0;
  #pragma unroll
  for (int i = 12; i >= 0; i--) p = __dadd_rn(__dmul_rn(p, r), inv_fact[i]);
  return es_ldexp(p, k);
}

struct EsRng { u32 lo, hi, row, tag, gen, draw; };
SBD_FI void es_block(EsRng& r, u32 w[4]) {  // full Philox4x32-10 block (the engine's helper returns two words)
  u32 c0 = r.draw++, c1 = r.row, c2 = r.tag, c3 = r.gen, k0 = r.lo, k1 = r.hi;
#pragma unroll
  for (int i = 0; i < 10; i++) {
    const u32 h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
    const u32 h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
    const u32 n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
    c0 = n0; c1 = l1; c2 = n2; c3 = l0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  w[0] = c0; w[1] = c1; w[2] = c2; w[3] = c3;
}
SBD_FI int es_below(EsRng& r, int n) { u32 w[4]; es_block(r, w); return (int)__umulhi(w[0], (u32)n); }
SBD_FI double es_u53(u32 a, u32 b) {
  return __ddiv_rn(__dadd_rn(__dmul_rn((double)(a >> 5), 67108864.0), (double)(b >> 6)), 9007199254740992.0);
}
SBD_FI double es_uniform(EsRng& r, double lo, double hi) {
  u32 w[4];
  es_block(r, w);
  return __dadd_rn(lo, __dmul_rn(__dadd_rn(hi, -lo), es_u53(w[0], w[1])));
}
SBD double es_normal(EsRng& r) {  // Marsaglia polar: one block per attempt, first variate only
  for (;;) {
    u32 w[4];
    es_block(r, w);
    const double u = __dadd_rn(__dmul_rn(2.0, es_u53(w[0], w[1])), -1.0), v = __dadd_rn(__dmul_rn(2.0, es_u53(w[2], w[3])), -1.0);
    const double s = __dadd_rn(__dmul_rn(u, u), __dmul_rn(v, v));
    if (s >= 1.0 || s == 0.0) continue;
    return __dmul_rn(u, __dsqrt_rn(__ddiv_rn(__dmul_rn(-2.0, es_log(s)), s)));
  }
}
// WeightVector.mutate on one row held in registers / local memory
SBD void es_mutate(EsRng& r, int nf, double tau, double tau_prime, double min_sigma, double* w, double* s) {
  double ind[ES_MAX_FEATURES];
  const double g = es_normal(r);
  #pragma unroll 1
  for (int i = 0; i < nf; i++) ind[i] = es_normal(r);
  #pragma unroll 1
  for (int i = 0; i < nf; i++) {
    const double v = __dmul_rn(s[i], es_exp(__dadd_rn(__dmul_rn(tau_prime, g), __dmul_rn(tau, ind[i]))));
    s[i] = v > min_sigma ? v : min_sigma;
  }
  #pragma unroll 1
  for (int i = 0; i < nf; i++) {
    const double v = __dadd_rn(w[i], __dadd_rn(0.0, __dmul_rn(s[i], es_normal(r))));
    w[i] = v < 0.0 ? 0.0 : v > 1.0 ? 1.0 : v;
  }
}

// rows [0, mu) are the parents; child c is written to row mu + c
__global__ void __launch_bounds__(128) k_es_offspring(unsigned long long seed, u32 generation, int mu, int lambda, int nf, double tau,
                                                     double tau_prime, double min_sigma, double* w, double* s, int* parent_out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= lambda) return;
  EsRng r = {(u32)seed, (u32)(seed >> 32), (u32)c, ES_TAG_OFFSPRING, generation, 0u};
  const int p = es_below(r, mu);
  if (parent_out) parent_out[c] = p;
  double cw[ES_MAX_FEATURES], cs[ES_MAX_FEATURES];
  #pragma unroll 1
  for (int i = 0; i < nf; i++) { cw[i] = w[(size_t)p * nf + i]; cs[i] = s[(size_t)p * nf + i]; }
  es_mutate(r, nf, tau, tau_prime, min_sigma, cw, cs);
  #pragma unroll 1
  for (int i = 0; i < nf; i++) { w[(size_t)(mu + c) * nf + i] = cw[i]; s[(size_t)(mu + c) * nf + i] = cs[i]; }
}

// top-mu by fitness, descending, ties in original order (Python's sorted(..., reverse=True) is stable): every row
// counts the rows that precede it; the fitness vector (a few thousand doubles) is read through L1/L2.
__global__ void __launch_bounds__(128) k_es_select(int total, int mu, int nf, const double* fitness, const double* w, const double* s,
                                                  double* w_out, double* s_out, double* fit_out, int* order_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const double f = fitness[i];
  int rank = 0;
  #pragma unroll 4
  for (int j = 0; j < total; j++) { const double fj = __ldg(fitness + j); rank += (fj > f) || (fj == f && j < i); }
  if (rank >= mu) return;
  #pragma unroll 1
  for (int k = 0; k < nf; k++) { w_out[(size_t)rank * nf + k] = w[(size_t)i * nf + k]; s_out[(size_t)rank * nf + k] = s[(size_t)i * nf + k]; }
  fit_out[rank] = f;
  if (order_out) order_out[rank] = i;
}

__global__ void __launch_bounds__(128) k_es_reset_sigmas(unsigned long long seed, u32 generation, int mu, int nf, double initial_sigma, double* s) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= mu) return;
  EsRng r = {(u32)seed, (u32)(seed >> 32), (u32)i, ES_TAG_RESET, generation, 0u};
  const double lo = __dmul_rn(initial_sigma, 0.5), hi = __dmul_rn(initial_sigma, 1.5);
  #pragma unroll 1
  for (int k = 0; k < nf; k++) {
    const double v = es_uniform(r, lo, hi);
    s[(size_t)i * nf + k] = v > 1e-10 ? v : 1e-10;
  }
}

// one CTA: thread 0 draws the permutation (choice(mu, mu/2, replace=False)), then the threads share the chosen rows
__global__ void __launch_bounds__(256) k_es_inject_diversity(unsigned long long seed, u32 generation, int mu, int nf, double tau, double tau_prime,
                                                            double min_sigma, double initial_sigma, double* w, double* s, int* chosen_out) {
  __shared__ short perm[ES_MAX_MU];
  const int k = mu / 2 > 1 ? mu / 2 : 1;
  if (threadIdx.x == 0) {
    EsRng pr = {(u32)seed, (u32)(seed >> 32), 0xFFFFFFFFu, ES_TAG_INJECT, generation, 0u};
    for (int i = 0; i < mu; i++) perm[i] = (short)i;
    for (int i = mu - 1; i > 0; i--) { const int j = es_below(pr, i + 1); const short t = perm[i]; perm[i] = perm[j]; perm[j] = t; }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < k; c += blockDim.x) {
    const int row = perm[c];
    if (chosen_out) chosen_out[c] = row;
    EsRng r = {(u32)seed, (u32)(seed >> 32), (u32)row, ES_TAG_INJECT, generation, 0u};
    double cw[ES_MAX_FEATURES], cs[ES_MAX_FEATURES], orig[ES_MAX_FEATURES];
    #pragma unroll 1
    for (int i = 0; i < nf; i++) {
      cw[i] = w[(size_t)row * nf + i];
      orig[i] = s[(size_t)row * nf + i];
      const double b = __dmul_rn(orig[i], 5.0);
      cs[i] = b > 1e-10 ? b : 1e-10;
    }
    #pragma unroll 1
    for (int rep = 0; rep < 3; rep++) es_mutate(r, nf, __dmul_rn(tau, 2.0), __dmul_rn(tau_prime, 2.0), min_sigma, cw, cs);
    const double floor_sigma = __dmul_rn(initial_sigma, 0.5);
    #pragma unroll 1
    for (int i = 0; i < nf; i++) {
      const double v = orig[i] > floor_sigma ? orig[i] : floor_sigma;
      w[(size_t)row * nf + i] = cw[i];
      s[(size_t)row * nf + i] = v > 1e-10 ? v : 1e-10;
    }
  }
}
