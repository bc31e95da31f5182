/*
 * oracle/s4b_rng.h -- TEST INFRASTRUCTURE (CPU oracle), not product code.
 *
 * Random number source shared *by specification* (not by code) between the CPU
 * oracle and the CUDA product ("s4b-rng v1").  The reference draws BART
 * randomness from R's global generator (Mersenne-Twister + inversion; see
 * /root/reference/src/init.cpp:259,298,750,919 GetRNGstate/PutRNGstate and
 * SURVEY.md App. B) and Stan randomness from boost::ecuyer1988
 * (/root/reference/src/interruptable_sampler.hpp:104,
 * src/include/stan/services/util/create_rng.hpp:25-31).  Neither R nor Boost
 * exists in this image, and a sequential generator cannot feed N parallel
 * latent draws, so both sides of the parity test use this counter-based
 * generator, or -- in replay mode -- a pre-recorded "tape" of draws.
 *
 *   block   : Philox4x32-10 (Salmon et al. 2011), key = (seed_lo, seed_hi),
 *             counter = (c0, c1, c2, stream)
 *   uniform : u = (((o0 >> 6) << 26 | (o1 >> 6)) + 0.5) * 2^-52   in (0,1)
 *   normal  : Wichura AS241 PPND16 inversion of one uniform (R's default
 *             normal.kind = "Inversion" also inverts, R nmath/qnorm.c)
 *   exp(1)  : -log(u)
 *   index   : floor(u * n), clamped to n-1
 *
 * Counter layout of a stream: c0 = draw index inside a substream, c1 = low 32 bits of
 * the substream's step number, c2 = (sub << 30) | high 30 bits of the step, c3 = stream.
 * The BART stream is keyed by tree step so that independent pieces of work do not
 * serialise on one counter: sub 0 = proposal draws of tree step `step`, sub 1 = its
 * decision draws (accept uniform, then one normal per bottom node, left to right),
 * sub 2 = sampleTreesFromPrior (step = call * numTrees + tree).  Inside a substream draws
 * are consumed in order; the Stan stream uses (step 0, sub 0) and carries idx into step.
 * The probit-latent stream is keyed per observation: counter = (sub-draw, epoch, obs
 * index, stream 2).  A replay tape is consumed strictly in program order.
 */
#ifndef S4B_ORACLE_RNG_H
#define S4B_ORACLE_RNG_H

#include <math.h>
#include <stdint.h>
#include <stddef.h>

#define S4B_STREAM_BART   0u
#define S4B_STREAM_STAN   1u
#define S4B_STREAM_LATENT 2u

static inline void s4b_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
  uint32_t k0 = key[0], k1 = key[1];
  for (int round = 0; round < 10; ++round) {
    uint64_t p0 = (uint64_t) 0xD2511F53u * c0;
    uint64_t p1 = (uint64_t) 0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t) (p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t) p1;
    uint32_t n2 = (uint32_t) (p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t) p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static inline double s4b_bits_to_uniform(uint32_t o0, uint32_t o1)
{
  uint64_t k = ((uint64_t) (o0 >> 6) << 26) | (uint64_t) (o1 >> 6);
  return ((double) k + 0.5) * 0x1.0p-52;
}

/* Wichura (1988) AS241 PPND16, as in R nmath/qnorm.c */
static inline double s4b_qnorm(double p)
{
  double q = p - 0.5, r, val;
  if (fabs(q) <= 0.425) {
    r = 0.180625 - q * q;
    val = q * (((((((r * 2509.0809287301226727 +
                     33430.575583588128105) * r + 67265.770927008700853) * r
                   + 45921.953931549871457) * r + 13731.693765509461125) * r
                 + 1971.5909503065514427) * r + 133.14166789178437745) * r
               + 3.387132872796366608)
          / (((((((r * 5226.495278852545925 +
                   28729.085735721942674) * r + 39307.89580009271061) * r
                 + 21213.794301586595867) * r + 5394.1960214247511077) * r
               + 687.1870074920579083) * r + 42.313330701600911252) * r + 1.0);
    return val;
  }
  r = q < 0.0 ? p : 1.0 - p;
  r = sqrt(-log(r));
  if (r <= 5.0) {
    r -= 1.6;
    val = (((((((r * 7.7454501427834140764e-4 +
                 0.0227238449892691845833) * r + 0.24178072517745061177) *
               r + 1.27045825245236838258) * r +
              3.64784832476320460504) * r + 5.7694972214606914055) *
            r + 4.6303378461565452959) * r + 1.42343711074968357734)
          / (((((((r * 1.05075007164441684324e-9 + 5.475938084995344946e-4) *
                  r + 0.0151986665636164571966) * r +
                 0.14810397642748007459) * r + 0.68976733498510000455) *
               r + 1.6763848301838038494) * r +
              2.05319162663775882187) * r + 1.0);
  } else {
    r -= 5.0;
    val = (((((((r * 2.01033439929228813265e-7 +
                 2.71155556874348757815e-5) * r +
                0.0012426609473880784386) * r + 0.026532189526576123093) *
              r + 0.29656057182850489123) * r +
             1.7848265399172913358) * r + 5.4637849111641143699) *
           r + 6.6579046435011037772)
          / (((((((r * 2.04426310338993978564e-15 + 1.4215117583164458887e-7) *
                  r + 1.8463183175100546818e-5) * r +
                 7.868691311456132591e-4) * r + 0.0148753612908506148525)
               * r + 0.13692988092273580531) * r +
              0.59983220655588793769) * r + 1.0);
  }
  return q < 0.0 ? -val : val;
}

/* sequential stream with optional tape (replay) and optional recording */
typedef struct s4b_rng {
  uint32_t key[2];
  uint32_t stream;
  uint32_t sub, idx;
  uint64_t step;
  uint64_t counter;      /* total number of draws consumed (for tests) */
  /* replay: if tape != NULL draws are read from it (uniforms and normals
     interleaved in consumption order); record: if rec != NULL every draw is
     appended (until rec_cap) */
  const double* tape; size_t tape_len; size_t tape_pos;
  double* rec; size_t rec_cap; size_t rec_len;
  int tape_underrun;
} s4b_rng;

static inline void s4b_rng_init(s4b_rng* g, uint64_t seed, uint32_t stream)
{
  g->key[0] = (uint32_t) seed; g->key[1] = (uint32_t) (seed >> 32);
  g->stream = stream; g->counter = 0; g->step = 0; g->sub = 0; g->idx = 0;
  g->tape = NULL; g->tape_len = 0; g->tape_pos = 0;
  g->rec = NULL; g->rec_cap = 0; g->rec_len = 0; g->tape_underrun = 0;
}

/* switch to substream (step, sub); a no-op when already there */
static inline void s4b_rng_enter(s4b_rng* g, uint64_t step, uint32_t sub)
{
  if (g->step != step || g->sub != sub) { g->step = step; g->sub = sub; g->idx = 0; }
}

static inline double s4b_rng_raw_uniform(s4b_rng* g)
{
  uint32_t ctr[4] = { g->idx, (uint32_t) g->step, (g->sub << 30) | (uint32_t) ((g->step >> 32) & 0x3FFFFFFFu), g->stream };
  uint32_t out[4];
  s4b_philox4x32_10(ctr, g->key, out);
  g->idx++;
  if (g->idx == 0) g->step++;
  return s4b_bits_to_uniform(out[0], out[1]);
}

static inline double s4b_rng_record(s4b_rng* g, double v)
{
  g->counter++;
  if (g->rec != NULL && g->rec_len < g->rec_cap) g->rec[g->rec_len] = v;
  if (g->rec != NULL) g->rec_len++;
  return v;
}

static inline double s4b_rng_uniform(s4b_rng* g)
{
  if (g->tape != NULL) {
    if (g->tape_pos >= g->tape_len) { g->tape_underrun = 1; return 0.5; }
    return s4b_rng_record(g, g->tape[g->tape_pos++]);
  }
  return s4b_rng_record(g, s4b_rng_raw_uniform(g));
}

static inline double s4b_rng_normal(s4b_rng* g)
{
  if (g->tape != NULL) {
    if (g->tape_pos >= g->tape_len) { g->tape_underrun = 1; return 0.0; }
    return s4b_rng_record(g, g->tape[g->tape_pos++]);
  }
  return s4b_rng_record(g, s4b_qnorm(s4b_rng_raw_uniform(g)));
}

/* Gamma(shape, 1) by Marsaglia & Tsang (2000) on the stream's normal and uniform draws (shape < 1 through the usual
 * U^(1/shape) boost); used for the k hyperprior of the leaf prior */
static inline double s4b_rng_gamma(s4b_rng* g, double shape)
{
  double boost = 1.0;
  if (shape < 1.0) { boost = pow(s4b_rng_uniform(g), 1.0 / shape); shape += 1.0; }
  const double d = shape - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
  for (;;) {
    double x = s4b_rng_normal(g);
    double v = 1.0 + c * x;
    if (v <= 0.0) continue;
    v = v * v * v;
    double u = s4b_rng_uniform(g);
    if (log(u) < 0.5 * x * x + d - d * v + d * log(v)) return boost * d * v;
    if (g->tape != NULL && g->tape_underrun) return boost * d;
  }
}

static inline size_t s4b_rng_index(s4b_rng* g, size_t n)
{
  size_t k = (size_t) (s4b_rng_uniform(g) * (double) n);
  return k >= n ? n - 1 : k;
}

/* keyed (per-observation) draws for the probit latents */
static inline double s4b_keyed_uniform(uint64_t seed, uint32_t obs, uint32_t epoch, uint32_t sub)
{
  uint32_t key[2] = { (uint32_t) seed, (uint32_t) (seed >> 32) };
  uint32_t ctr[4] = { sub, epoch, obs, S4B_STREAM_LATENT };
  uint32_t out[4];
  s4b_philox4x32_10(ctr, key, out);
  return s4b_bits_to_uniform(out[0], out[1]);
}

/* z ~ N(mean, 1) truncated to z > 0 (positive != 0) or z < 0 (positive == 0), by inversion from ONE keyed uniform
 * (sub-draw 0): with m = +-mean, x = m - Phi^-1(u Phi(m)) is N(m, 1) conditioned on x > 0 for every m (Phi(m) from erfc, so
 * the far tail m << 0 keeps full relative accuracy).  dbarts draws the same distribution with Robert's (1995) rejection
 * sampler (ext_rng_simulate{Lower,Upper}TruncatedNormalScale1, un-vendored); its stream cannot be reproduced anyway, and
 * a fixed one-draw recipe has no data-dependent loop on the GPU. */
static inline double s4b_keyed_truncnorm(uint64_t seed, uint32_t obs, uint32_t epoch, double mean, int positive)
{
  double m = positive ? mean : -mean;   /* reduce to lower truncation at 0 of N(m,1) */
  double u = s4b_keyed_uniform(seed, obs, epoch, 0u);
  double pm = 0.5 * erfc(-m * 0.70710678118654752440);
  double arg = u * pm;
  if (arg < 1e-300) arg = 1e-300;
  double x = m - s4b_qnorm(arg);
  if (!(x > 0.0)) x = 0.0;               /* rounding at the boundary */
  return positive ? x : -x;
}

#endif
