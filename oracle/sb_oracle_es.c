/* TEST INFRASTRUCTURE -- CPU restatement of the (mu + lambda) evolution-strategy operators:
 *   WeightVector.mutate                 evo/weights.py:20-40
 *   Population.generate_offspring       evo/population.py:75-90
 *   Population.select_from_combined     evo/population.py:92-176 (top-mu, sigma reset, diversity injection)
 * with the injected counter-based stream instead of numpy's process-global MT19937:
 *   block(row, draw) = philox(counter=(draw, row, tag, generation), key=seed)
 * tag 0xE5 offspring (row = child index), 0xE6 sigma reset (row = survivor index), 0xE7 diversity injection
 * (row = survivor index; row 0xFFFFFFFF draws the permutation).  Every row owns its draw counter, so rows are
 * independent (one thread per row on the GPU) while the call shapes inside a row are the reference's:
 * randint -> normal() -> normal(size=n) -> normal(0, sigmas).
 *
 * Floating point: exp / log are restated with IEEE +,-,*,/ and sqrt only (no libm, no FMA contraction), so this
 * file and the CUDA kernels produce IDENTICAL bits; against numpy's exp the results agree to ~1e-15 relative
 * (tests use 1e-12).  Parity status: pinned live against the reference's Population / WeightVector driven with
 * the same stream (oracle/validate_vs_reference.py check_es) and by tests/golden/es_operators.npz.
 */
#include <stdint.h>
#include <string.h>
#include <math.h>

static void philox_es(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4]) {
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* ---- deterministic exp / log (IEEE basic operations only) */
static double det_ldexp(double x, int e) { /* x * 2^e for results in the normal range */
  union { double d; uint64_t u; } s;
  s.u = (uint64_t)(1023 + e) << 52;
  return x * s.d;
}
double sbo_det_log(double x) { /* x > 0, normal */
  union { double d; uint64_t u; } s;
  s.d = x;
  int e = (int)((s.u >> 52) & 0x7FF) - 1022;           /* x = m * 2^e, m in [0.5, 1) */
  s.u = (s.u & 0x000FFFFFFFFFFFFFull) | 0x3FE0000000000000ull;
  double m = s.d;
  if (m < 0.70710678118654752440) { m = m * 2.0; e -= 1; } /* m in [sqrt(1/2), sqrt(2)) */
  double z = (m - 1.0) / (m + 1.0), z2 = z * z;          /* log m = 2 atanh z, |z| <= 0.1716 */
  double p = 1.0 / 27.0;
  for (int k = 25; k >= 1; k -= 2) p = p * z2 + 1.0 / (double)k;
  return (double)e * 0.693147180559945309417232 + 2.0 * z * p;
}
double sbo_det_exp(double y) { /* |y| < 700 */
  double kf = y * 1.44269504088896340736;
  int k = (int)(kf + (kf >= 0 ? 0.5 : -0.5));
  double r = (y - (double)k * 0.693147180369123816490) - (double)k * 1.90821492927058770002e-10; /* ln2 hi/lo */
  double p = 1.0 / 6227020800.0;                                                             /* 1/13! */
  static const double inv_fact[13] = {1.0, 1.0, 0.5, 1.0 / 6, 1.0 / 24, 1.0 / 120, 1.0 / 720, 1.0 / 5040, 1.0 / 40320,
                                      1.0 / 362880, 1.0 / 3628800, 1.0 / 39916800, 1.0 / 479001600};
  for (int i = 12; i >= 0; i--) p = p * r + inv_fact[i];
  return det_ldexp(p, k);
}

/* ---- per-row stream */
typedef struct { uint32_t lo, hi, row, tag, gen, draw; } EsRng;
static void es_block(EsRng *r, uint32_t w[4]) { philox_es(r->draw++, r->row, r->tag, r->gen, r->lo, r->hi, w); }
static int es_below(EsRng *r, int n) { uint32_t w[4]; es_block(r, w); return (int)(((uint64_t)w[0] * (uint64_t)n) >> 32); }
static double u53(uint32_t a, uint32_t b) { return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) / 9007199254740992.0; }
static double es_uniform(EsRng *r, double lo, double hi) { uint32_t w[4]; es_block(r, w); return lo + (hi - lo) * u53(w[0], w[1]); }
static double es_normal(EsRng *r) { /* Marsaglia polar, one block per attempt, first variate only */
  for (;;) {
    uint32_t w[4];
    es_block(r, w);
    double u = 2.0 * u53(w[0], w[1]) - 1.0, v = 2.0 * u53(w[2], w[3]) - 1.0;
    double s = u * u + v * v;
    if (s >= 1.0 || s == 0.0) continue;
    return u * sqrt(-2.0 * sbo_det_log(s) / s);
  }
}

/* WeightVector.mutate (evo/weights.py:20-40) on one row */
static void es_mutate(EsRng *r, int nf, double tau, double tau_prime, double min_sigma, double *w, double *s) {
  double ind[64];
  double g = es_normal(r);
  for (int i = 0; i < nf; i++) ind[i] = es_normal(r);
  for (int i = 0; i < nf; i++) {
    double v = s[i] * sbo_det_exp(tau_prime * g + tau * ind[i]);
    s[i] = v > min_sigma ? v : min_sigma;
  }
  for (int i = 0; i < nf; i++) {
    double v = w[i] + (0.0 + s[i] * es_normal(r));
    w[i] = v < 0.0 ? 0.0 : v > 1.0 ? 1.0 : v;
  }
}

/* Population.generate_offspring: rows [0, mu) are the parents, rows [mu, mu+lambda) are written */
void sbo_es_offspring(uint64_t seed, uint32_t generation, int mu, int lambda, int nf, double tau, double tau_prime,
                      double min_sigma, double *w, double *s, int32_t *parent_out) {
  for (int c = 0; c < lambda; c++) {
    EsRng r = {(uint32_t)seed, (uint32_t)(seed >> 32), (uint32_t)c, 0xE5u, generation, 0};
    int p = es_below(&r, mu);
    if (parent_out) parent_out[c] = p;
    memcpy(w + (size_t)(mu + c) * nf, w + (size_t)p * nf, sizeof(double) * nf);
    memcpy(s + (size_t)(mu + c) * nf, s + (size_t)p * nf, sizeof(double) * nf);
    es_mutate(&r, nf, tau, tau_prime, min_sigma, w + (size_t)(mu + c) * nf, s + (size_t)(mu + c) * nf);
  }
}

/* top-mu of `total` rows by fitness, descending, ties in original order (sorted(..., reverse=True) is stable);
 * order_out[k] = source row of survivor k */
void sbo_es_select(int total, int mu, int nf, const double *fitness, const double *w, const double *s, double *w_out,
                   double *s_out, double *fit_out, int32_t *order_out) {
  for (int i = 0; i < total; i++) {
    int rank = 0;
    for (int j = 0; j < total; j++) rank += fitness[j] > fitness[i] || (fitness[j] == fitness[i] && j < i);
    if (rank < mu) {
      memcpy(w_out + (size_t)rank * nf, w + (size_t)i * nf, sizeof(double) * nf);
      memcpy(s_out + (size_t)rank * nf, s + (size_t)i * nf, sizeof(double) * nf);
      fit_out[rank] = fitness[i];
      if (order_out) order_out[rank] = i;
    }
  }
}

/* sigma reset (evo/population.py:128-139): uniform(0.5, 1.5) x initial_sigma per feature, floor 1e-10 (set_sigmas) */
void sbo_es_reset_sigmas(uint64_t seed, uint32_t generation, int mu, int nf, double initial_sigma, double *s) {
  for (int i = 0; i < mu; i++) {
    EsRng r = {(uint32_t)seed, (uint32_t)(seed >> 32), (uint32_t)i, 0xE6u, generation, 0};
    for (int k = 0; k < nf; k++) {
      double v = es_uniform(&r, initial_sigma * 0.5, initial_sigma * 1.5);
      s[(size_t)i * nf + k] = v > 1e-10 ? v : 1e-10;
    }
  }
}

/* diversity injection (evo/population.py:146-170): choice(mu, mu/2, replace=False) = first half of a
 * Fisher-Yates permutation; each chosen row: sigmas x5, three mutate(2 tau, 2 tau', min_sigma), then
 * sigmas = max(original, initial_sigma / 2) (floor 1e-10).  chosen_out: mu/2 row indices in draw order. */
void sbo_es_inject_diversity(uint64_t seed, uint32_t generation, int mu, int nf, double tau, double tau_prime, double min_sigma,
                             double initial_sigma, double *w, double *s, int32_t *chosen_out) {
  int perm[4096];
  int k = mu / 2 > 1 ? mu / 2 : 1;
  EsRng pr = {(uint32_t)seed, (uint32_t)(seed >> 32), 0xFFFFFFFFu, 0xE7u, generation, 0};
  for (int i = 0; i < mu; i++) perm[i] = i;
  for (int i = mu - 1; i > 0; i--) { int j = es_below(&pr, i + 1); int t = perm[i]; perm[i] = perm[j]; perm[j] = t; }
  for (int c = 0; c < k; c++) {
    int row = perm[c];
    if (chosen_out) chosen_out[c] = row;
    EsRng r = {(uint32_t)seed, (uint32_t)(seed >> 32), (uint32_t)row, 0xE7u, generation, 0};
    double orig[64];
    double *sr = s + (size_t)row * nf, *wr = w + (size_t)row * nf;
    for (int i = 0; i < nf; i++) { orig[i] = sr[i]; double b = sr[i] * 5.0; sr[i] = b > 1e-10 ? b : 1e-10; }
    for (int rep = 0; rep < 3; rep++) es_mutate(&r, nf, tau * 2, tau_prime * 2, min_sigma, wr, sr);
    for (int i = 0; i < nf; i++) {
      double v = orig[i] > initial_sigma * 0.5 ? orig[i] : initial_sigma * 0.5;
      sr[i] = v > 1e-10 ? v : 1e-10;
    }
  }
}
