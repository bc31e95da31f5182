/*
 * oracle/oracle_glmm.c -- TEST INFRASTRUCTURE (CPU oracle), not product code.
 *
 * CPU restatement of the stan4bart GLMM density and its gradient on the default
 * path (prior_dist in {0,1}, no intercept, optional observation weights; ranef blocks of any size, the ones with more than two
 * coefficients through the scaled onion rows of make_theta_L):
 *   maths          /root/reference/src/stan_files/continuous.stan:1-429
 *   op order       /root/reference/src/stan_files/continuous.hpp:2168-2638 (log_prob_impl,
 *                  lpdfs hard-coded <false> = all constants kept)
 *   make_theta_L   continuous.stan:2-59      make_b  continuous.stan:61-94
 *   decov_lp       continuous.stan:96-122 (continuous.hpp:823-916)
 *   transforms     src/include/stan/math/prim/fun/lb_constrain.hpp:64 (exp, J = +x),
 *                  lub_constrain.hpp:109-110 (inv_logit, J = -|x| - 2 log1p(exp(-|x|)))
 *   write_array    continuous.hpp:2640-2938 (order: params, then aux, beta, b, theta_L)
 *   set_offset / set_response / get_aux / get_parametric_mean  continuous.hpp:3626-3768
 * The reference differentiates by reverse-mode AD (model/gradient.hpp:21-35); here the
 * gradient is hand-derived; for blocks with more than two coefficients the Jacobian of the block factor T with
 * respect to its parameters is taken by the complex-step method (exact to rounding: no subtraction), everything
 * around it stays analytic.  Pinned by tests/golden/glmm_*.json (torch fp64 autograd).
 */
#include "s4b_oracle.h"

#include <complex.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

struct or_glmm {
  s4b_glmm_data d;
  double *X, *y, *offset, *weights;
  double *prior_scale, *prior_mean, *shape, *scale, *concentration, *regularization, *w, *delta;
  int32_t *p, *l, *v, *u;
  int len_rho, len_z_T, num_params, has_aux;
  /* coefficient priors beyond normal (continuous.stan:262-270): z_beta may be longer than K (product_normal) and the
   * shrinkage priors add positive parameters global[hs], local[hs][K], caux[hs > 0], mix[K], one_over_lambda */
  double* prior_df; int32_t* num_normals;
  int hs, len_zbeta, len_extra;
};
#define OR_MAX_NC 16

#define HALF_LOG_2PI 0.91893853320467274178

static void* dup_mem(const void* src, size_t bytes) { void* r = malloc(bytes ? bytes : 1); if (bytes) memcpy(r, src, bytes); return r; }

or_glmm* or_glmm_create(const s4b_glmm_data* d)
{
  for (int i = 0; i < d->t; ++i) if (d->p[i] < 1 || d->p[i] > OR_MAX_NC) return NULL;
  if (d->prior_dist < 0 || d->prior_dist > 7) return NULL;
  if ((d->prior_dist == 3 || d->prior_dist == 4) && d->is_binary) return NULL;      /* hs_prior reads aux[1] */
  if ((d->prior_dist == 2 || d->prior_dist == 3 || d->prior_dist == 4 || d->prior_dist == 6) && d->K > 0 && !d->prior_df) return NULL;
  if (d->prior_dist == 7 && d->K > 0 && !d->num_normals) return NULL;
  or_glmm* m = (or_glmm*) calloc(1, sizeof(or_glmm));
  m->d = *d;
  size_t N = (size_t) d->N, K = (size_t) d->K, t = (size_t) d->t;
  m->X = (double*) dup_mem(d->X, sizeof(double) * N * K);
  m->y = (double*) dup_mem(d->y, sizeof(double) * N);
  m->offset = (double*) calloc(N ? N : 1, sizeof(double));
  m->weights = d->weights ? (double*) dup_mem(d->weights, sizeof(double) * N) : NULL;
  m->prior_scale = (double*) dup_mem(d->prior_scale, sizeof(double) * K);
  m->prior_mean = (double*) dup_mem(d->prior_mean, sizeof(double) * K);
  m->p = (int32_t*) dup_mem(d->p, sizeof(int32_t) * t);
  m->l = (int32_t*) dup_mem(d->l, sizeof(int32_t) * t);
  m->shape = (double*) dup_mem(d->shape, sizeof(double) * t);
  m->scale = (double*) dup_mem(d->scale, sizeof(double) * t);
  m->concentration = (double*) dup_mem(d->concentration, sizeof(double) * (size_t) d->len_concentration);
  m->regularization = (double*) dup_mem(d->regularization, sizeof(double) * (size_t) d->len_regularization);
  m->w = (double*) dup_mem(d->w, sizeof(double) * (size_t) d->num_non_zero);
  m->v = (int32_t*) dup_mem(d->v, sizeof(int32_t) * (size_t) d->num_non_zero);
  m->u = (int32_t*) dup_mem(d->u, sizeof(int32_t) * (N + 1));
  /* transformed data, src/stan_sampler.cpp:142-182: delta restarts at concentration[0] for every term */
  m->delta = (double*) calloc((size_t) d->len_concentration + 1, sizeof(double));
  int pos = 0, sum_p = 0;
  for (int i = 0; i < d->t; ++i) {
    if (d->p[i] > 1) for (int j = 0; j < d->p[i]; ++j) m->delta[pos++] = d->concentration[j];
    sum_p += d->p[i];
  }
  m->len_rho = sum_p - d->t;
  m->len_z_T = 0;
  for (int i = 0; i < d->t; ++i) if (d->p[i] > 2) m->len_z_T += (d->p[i] - 2) * (d->p[i] - 1);      /* continuous.stan:258 */
  m->has_aux = d->is_binary ? 0 : 1;
  m->prior_df = d->prior_df ? (double*) dup_mem(d->prior_df, sizeof(double) * K) : NULL;
  m->num_normals = d->num_normals ? (int32_t*) dup_mem(d->num_normals, sizeof(int32_t) * K) : NULL;
  m->hs = d->prior_dist == 3 ? 2 : (d->prior_dist == 4 ? 4 : 0);
  m->len_zbeta = d->K;
  if (d->prior_dist == 7) { m->len_zbeta = 0; for (int k = 0; k < d->K; ++k) { if (d->num_normals[k] < 2) { or_glmm_free(m); return NULL; } m->len_zbeta += d->num_normals[k]; } }
  m->len_extra = m->hs + m->hs * d->K + (m->hs > 0) + ((d->prior_dist == 5 || d->prior_dist == 6) ? d->K : 0) + (d->prior_dist == 6);
  m->num_params = m->len_zbeta + m->len_extra + d->q + m->len_z_T + m->len_rho + d->len_concentration + d->t + m->has_aux;
  return m;
}

void or_glmm_free(or_glmm* m)
{
  if (!m) return;
  free(m->X); free(m->y); free(m->offset); free(m->weights); free(m->prior_df); free(m->num_normals); free(m->prior_scale); free(m->prior_mean); free(m->p); free(m->l);
  free(m->shape); free(m->scale); free(m->concentration); free(m->regularization); free(m->w); free(m->v); free(m->u); free(m->delta);
  free(m);
}

int or_glmm_num_params(const or_glmm* m) { return m->num_params; }
int or_glmm_num_constrained(const or_glmm* m) { return m->num_params + m->has_aux + m->d.K + m->d.q + m->d.len_theta_L; }
void or_glmm_set_offset(or_glmm* m, const double* offset) { memcpy(m->offset, offset, sizeof(double) * (size_t) m->d.N); }
void or_glmm_set_response(or_glmm* m, const double* y) { memcpy(m->y, y, sizeof(double) * (size_t) m->d.N); }

void or_glmm_data_terms(const or_glmm* m, const double* beta, const double* b, double* S, double* gbeta, double* gb)
{
  int64_t N = m->d.N; int K = m->d.K, q = m->d.q;
  double s = 0.0;
  for (int k = 0; k < K; ++k) gbeta[k] = 0.0;
  for (int k = 0; k < q; ++k) gb[k] = 0.0;
  for (int64_t i = 0; i < N; ++i) {
    double eta = m->offset[i];
    for (int k = 0; k < K; ++k) eta += m->X[(size_t) k * (size_t) N + (size_t) i] * beta[k];
    for (int32_t z = m->u[i]; z < m->u[i + 1]; ++z) eta += m->w[z] * b[m->v[z]];
    double e = m->y[i] - eta;
    if (m->weights) {            /* continuous.stan:358-366: -0.5 sum w (y - eta)^2 / sigma^2 */
      double we = m->weights[i] * e;
      s += we * e;
      e = we;
    } else s += e * e;
    for (int k = 0; k < K; ++k) gbeta[k] += m->X[(size_t) k * (size_t) N + (size_t) i] * e;
    for (int32_t z = m->u[i]; z < m->u[i + 1]; ++z) gb[m->v[z]] += m->w[z] * e;
  }
  *S = s;
}

/* constrained and transformed quantities for a given unconstrained q */
typedef struct {
  const double *z_beta, *extra_u, *z_b, *z_T, *rho_u, *zeta_u, *tau_u;
  double aux_u;
  double *rho, *zeta, *tau, *beta, *b, *theta_L, *extra;
  double aux_unscaled, aux, disp;
} Params;

static void params_alloc(const or_glmm* m, Params* P)
{
  P->rho = (double*) calloc((size_t) m->len_rho + 1, sizeof(double));
  P->zeta = (double*) calloc((size_t) m->d.len_concentration + 1, sizeof(double));
  P->tau = (double*) calloc((size_t) m->d.t + 1, sizeof(double));
  P->beta = (double*) calloc((size_t) m->d.K + 1, sizeof(double));
  P->b = (double*) calloc((size_t) m->d.q + 1, sizeof(double));
  P->theta_L = (double*) calloc((size_t) m->d.len_theta_L + 1, sizeof(double));
  P->extra = (double*) calloc((size_t) m->len_extra + 1, sizeof(double));
}
static void params_free(Params* P) { free(P->rho); free(P->zeta); free(P->tau); free(P->beta); free(P->b); free(P->theta_L); free(P->extra); }

/* lower-triangular factor T (row major, nc x nc) of one ranef block with nc >= 2 coefficients: continuous.stan:20-50.
 * The off-diagonal entries of row r + 1 are scaled with the standard deviation of row r, exactly as the Stan program
 * does.  Complex arguments: the imaginary part carries the complex-step derivative. */
static void block_T(int nc, double complex c, const double complex* zeta, const double complex* rho, const double complex* zT, double complex* T)
{
  double complex trace = c * c * (double) nc, zs = 0.0;
  for (int k = 0; k < nc; ++k) zs += zeta[k];
  for (int k = 0; k < nc * nc; ++k) T[k] = 0.0;
  double complex sd = csqrt(zeta[0] / zs * trace);
  T[0] = sd;
  sd = csqrt(zeta[1] / zs * trace);
  double complex r21 = 2.0 * rho[0] - 1.0;
  T[nc + 1] = sd * csqrt(1.0 - r21 * r21);
  T[nc] = sd * r21;
  int zmark = 0;
  for (int r = 2; r < nc; ++r) {
    double complex dot = 0.0;
    for (int k = 0; k < r; ++k) dot += zT[zmark + k] * zT[zmark + k];
    double complex sf = csqrt(rho[r - 1] / dot) * sd;
    sd = csqrt(zeta[r] / zs * trace);
    for (int k = 0; k < r; ++k) T[r * nc + k] = zT[zmark + k] * sf;
    T[r * nc + r] = csqrt(1.0 - rho[r - 1]) * sd;
    zmark += r;
  }
}

/* beta as a function of z_beta, the constrained extra parameters and aux: continuous.stan:293-322 with hs_prior /
 * hsplus_prior (:123-143) and CFt (:146-158).  Complex arguments carry the complex-step derivative. */
static void coef_beta(const or_glmm* m, const double complex* z, const double complex* ex, double complex aux, double complex* beta)
{
  const s4b_glmm_data* d = &m->d;
  const int K = d->K, hs = m->hs;
  switch (d->prior_dist) {
  case 0: for (int k = 0; k < K; ++k) beta[k] = z[k]; break;
  case 1: for (int k = 0; k < K; ++k) beta[k] = z[k] * m->prior_scale[k] + m->prior_mean[k]; break;
  case 2:
    for (int k = 0; k < K; ++k) {
      double complex z1 = z[k], z2 = z1 * z1, z3 = z2 * z1, z5 = z2 * z3, z7 = z2 * z5, z9 = z2 * z7;
      double df = m->prior_df[k], df2 = df * df, df3 = df2 * df, df4 = df2 * df2;
      double complex cft = z1 + (z3 + z1) / (4 * df) + (5 * z5 + 16 * z3 + 3 * z1) / (96 * df2)
                           + (3 * z7 + 19 * z5 + 17 * z3 - 15 * z1) / (384 * df3)
                           + (79 * z9 + 776 * z7 + 1482 * z5 - 1920 * z3 - 945 * z1) / (92160 * df4);
      beta[k] = cft * m->prior_scale[k] + m->prior_mean[k];
    }
    break;
  case 3: case 4: {
    const double complex* global = ex;
    const double complex* local = ex + hs;              /* local[j][k] = local[j * K + k] */
    double complex caux = ex[hs + hs * K];
    double complex c2 = d->slab_scale * d->slab_scale * caux;
    double complex tau = global[0] * csqrt(global[1]) * d->global_prior_scale * aux;
    for (int k = 0; k < K; ++k) {
      double complex lambda = local[k] * csqrt(local[K + k]);
      if (hs == 4) lambda = lambda * (local[2 * K + k] * csqrt(local[3 * K + k]));
      double complex l2 = lambda * lambda;
      double complex lt = csqrt(c2 * l2 / (c2 + tau * tau * l2));
      beta[k] = z[k] * lt * tau;
    }
    break;
  }
  case 5: for (int k = 0; k < K; ++k) beta[k] = m->prior_mean[k] + m->prior_scale[k] * csqrt(2.0 * ex[k]) * z[k]; break;
  case 6: for (int k = 0; k < K; ++k) beta[k] = m->prior_mean[k] + ex[K] * m->prior_scale[k] * csqrt(2.0 * ex[k]) * z[k]; break;
  default: {
    int zp = 0;
    for (int k = 0; k < K; ++k) {
      double complex v = z[zp++];
      for (int n = 2; n <= m->num_normals[k]; ++n) v = v * z[zp++];
      beta[k] = v * pow(m->prior_scale[k], (double) m->num_normals[k]) + m->prior_mean[k];
    }
  }
  }
}

static double inv_logit(double x) { return x >= 0.0 ? 1.0 / (1.0 + exp(-x)) : exp(x) / (1.0 + exp(x)); }

static void transform(const or_glmm* m, const double* q, Params* P)
{
  const s4b_glmm_data* d = &m->d;
  int pos = 0;
  P->z_beta = q + pos; pos += m->len_zbeta;
  P->extra_u = q + pos; pos += m->len_extra;
  for (int i = 0; i < m->len_extra; ++i) P->extra[i] = exp(P->extra_u[i]);
  P->z_b = q + pos; pos += d->q;
  P->z_T = q + pos; pos += m->len_z_T;
  P->rho_u = q + pos; pos += m->len_rho;
  P->zeta_u = q + pos; pos += d->len_concentration;
  P->tau_u = q + pos; pos += d->t;
  P->aux_u = m->has_aux ? q[pos] : 0.0;
  for (int i = 0; i < m->len_rho; ++i) P->rho[i] = inv_logit(P->rho_u[i]);
  for (int i = 0; i < d->len_concentration; ++i) P->zeta[i] = exp(P->zeta_u[i]);
  for (int i = 0; i < d->t; ++i) P->tau[i] = exp(P->tau_u[i]);
  if (m->has_aux) {
    P->aux_unscaled = exp(P->aux_u);
    if (d->prior_dist_for_aux == 0) P->aux = P->aux_unscaled;
    else {
      P->aux = d->prior_scale_for_aux * P->aux_unscaled;
      if (d->prior_dist_for_aux <= 2) P->aux += d->prior_mean_for_aux;
    }
    P->disp = P->aux;
  } else { P->aux_unscaled = 0.0; P->aux = 1.0; P->disp = 1.0; }
  {
    int nin = m->len_zbeta + m->len_extra;
    double complex* in = (double complex*) malloc(sizeof(double complex) * (size_t) (nin + d->K + 1));
    for (int i = 0; i < m->len_zbeta; ++i) in[i] = P->z_beta[i];
    for (int i = 0; i < m->len_extra; ++i) in[m->len_zbeta + i] = P->extra[i];
    coef_beta(m, in, in + m->len_zbeta, P->aux, in + nin);
    for (int k = 0; k < d->K; ++k) P->beta[k] = creal(in[nin + k]);
    free(in);
  }
  /* make_theta_L (continuous.stan:2-59) and make_b (:61-94) */
  int zeta_mark = 0, rho_mark = 0, th = 0, b_mark = 0, zT_mark = 0;
  for (int i = 0; i < d->t; ++i) {
    if (m->p[i] == 1) {
      double theta = P->tau[i] * m->scale[i] * P->disp;
      P->theta_L[th++] = theta;
      for (int s = 0; s < m->l[i]; ++s) P->b[b_mark + s] = theta * P->z_b[b_mark + s];
      b_mark += m->l[i];
    } else if (m->p[i] > 2) {
      const int nc = m->p[i];
      double complex T[OR_MAX_NC * OR_MAX_NC], zc[OR_MAX_NC], rc[OR_MAX_NC], tc[OR_MAX_NC * OR_MAX_NC];
      const int nzT = (nc - 2) * (nc + 1) / 2;          /* 2 + 3 + ... + (nc - 1) */
      for (int k = 0; k < nc; ++k) zc[k] = P->zeta[zeta_mark + k];
      for (int k = 0; k < nc - 1; ++k) rc[k] = P->rho[rho_mark + k];
      for (int k = 0; k < nzT; ++k) tc[k] = P->z_T[zT_mark + k];
      block_T(nc, P->tau[i] * m->scale[i] * P->disp, zc, rc, tc, T);
      for (int c2 = 0; c2 < nc; ++c2) for (int r = c2; r < nc; ++r) P->theta_L[th++] = creal(T[r * nc + c2]);     /* vech */
      for (int j = 0; j < m->l[i]; ++j) {
        for (int r = 0; r < nc; ++r) { double acc = 0.0; for (int k = 0; k <= r; ++k) acc += creal(T[r * nc + k]) * P->z_b[b_mark + k]; P->b[b_mark + r] = acc; }
        b_mark += nc;
      }
      zeta_mark += nc; rho_mark += nc - 1; zT_mark += nzT;
    } else {
      double c = P->tau[i] * m->scale[i] * P->disp;
      double trace = c * c * 2.0;
      double zs = P->zeta[zeta_mark] + P->zeta[zeta_mark + 1];
      double pi1 = P->zeta[zeta_mark] / zs, pi2 = P->zeta[zeta_mark + 1] / zs;
      zeta_mark += 2;
      double sd1 = sqrt(pi1 * trace), sd2 = sqrt(pi2 * trace);
      double T21c = 2.0 * P->rho[rho_mark++] - 1.0;
      double T11 = sd1, T22 = sd2 * sqrt(1.0 - T21c * T21c), T21 = sd2 * T21c;
      P->theta_L[th++] = T11; P->theta_L[th++] = T21; P->theta_L[th++] = T22;
      for (int j = 0; j < m->l[i]; ++j) {
        double z0 = P->z_b[b_mark], z1 = P->z_b[b_mark + 1];
        P->b[b_mark] = T11 * z0; P->b[b_mark + 1] = T21 * z0 + T22 * z1;
        b_mark += 2;
      }
    }
  }
}

int or_glmm_log_prob_grad(const or_glmm* m, const double* q, double* lp_out, double* grad)
{
  const s4b_glmm_data* d = &m->d;
  Params P; params_alloc(m, &P); transform(m, q, &P);
  int K = d->K, nq = d->q, t = d->t;
  double N = (double) d->N;
  double lp = 0.0;
  /* Jacobians */
  for (int i = 0; i < m->len_rho; ++i) { double x = P.rho_u[i]; lp += -fabs(x) - 2.0 * log1p(exp(-fabs(x))); }
  for (int i = 0; i < d->len_concentration; ++i) lp += P.zeta_u[i];
  for (int i = 0; i < t; ++i) lp += P.tau_u[i];
  if (m->has_aux) lp += P.aux_u;
  /* likelihood */
  double* gbeta = (double*) calloc((size_t) K + 1, sizeof(double));
  double* gb = (double*) calloc((size_t) nq + 1, sizeof(double));
  double S; or_glmm_data_terms(m, P.beta, P.b, &S, gbeta, gb);
  double sigma = m->has_aux ? P.aux : 1.0;
  lp += -0.5 * S / (sigma * sigma) - N * log(sigma) - N * HALF_LOG_2PI;
  /* priors */
  double d_au_prior = 0.0;
  if (m->has_aux && d->prior_dist_for_aux > 0 && d->prior_scale_for_aux > 0.0) {
    double au = P.aux_unscaled;
    if (d->prior_dist_for_aux == 1) { lp += -0.5 * au * au - HALF_LOG_2PI + 0.693147180559945286; d_au_prior = -au; }
    else if (d->prior_dist_for_aux == 2) {
      double nu = d->prior_df_for_aux;
      lp += lgamma(0.5 * (nu + 1.0)) - lgamma(0.5 * nu) - 0.5 * log(nu * 3.14159265358979323846) - 0.5 * (nu + 1.0) * log1p(au * au / nu) + 0.693147180559945286;
      d_au_prior = -(nu + 1.0) * au / (nu + au * au);
    } else { lp += -au; d_au_prior = -1.0; }
  }
  if (d->prior_dist >= 1) { for (int k = 0; k < m->len_zbeta; ++k) lp += -0.5 * P.z_beta[k] * P.z_beta[k]; lp -= m->len_zbeta * HALF_LOG_2PI; }
  /* the extra parameters of the shrinkage priors: lb_constrain Jacobian + their priors (continuous.stan:381-408);
   * d_extra[i] = d prior / d (constrained value) */
  double* d_extra = (double*) calloc((size_t) m->len_extra + 1, sizeof(double));
  for (int i = 0; i < m->len_extra; ++i) lp += P.extra_u[i];
  if (m->hs > 0) {
    const int hs = m->hs;
    const double* local = P.extra + hs;
    double* d_local = d_extra + hs;
    for (int j = 0; j < hs; ++j) for (int k = 0; k < K; ++k) {
      double x = local[j * K + k];
      if ((j & 1) == 0) { lp += -0.5 * x * x - HALF_LOG_2PI; d_local[j * K + k] = -x; }   /* normal_lpdf(local[j] | 0, 1), - log_half once below */
      else {
        double a = 0.5 * (j == 1 ? m->prior_df[k] : m->prior_scale[k]);        /* hs_plus: prior_scale as a second df */
        lp += a * log(a) - lgamma(a) - (a + 1.0) * log(x) - a / x;
        d_local[j * K + k] = -(a + 1.0) / x + a / (x * x);
      }
    }
    lp += 0.693147180559945286 * (hs / 2);       /* "- log_half" is subtracted once per vector statement (continuous.stan:384, :393, :395) */
    { double x = P.extra[0]; lp += -0.5 * x * x - HALF_LOG_2PI + 0.693147180559945286; d_extra[0] = -x; }
    { double x = P.extra[1], a = 0.5 * d->global_prior_df; lp += a * log(a) - lgamma(a) - (a + 1.0) * log(x) - a / x; d_extra[1] = -(a + 1.0) / x + a / (x * x); }
    { double x = P.extra[hs + hs * K], a = 0.5 * d->slab_df; lp += a * log(a) - lgamma(a) - (a + 1.0) * log(x) - a / x; d_extra[hs + hs * K] = -(a + 1.0) / x + a / (x * x); }
  } else if (d->prior_dist == 5 || d->prior_dist == 6) {
    for (int k = 0; k < K; ++k) { lp += -P.extra[k]; d_extra[k] = -1.0; }
    if (d->prior_dist == 6) {
      double nu = m->prior_df[0], x = P.extra[K];
      lp += -(0.5 * nu) * 0.693147180559945286 - lgamma(0.5 * nu) + (0.5 * nu - 1.0) * log(x) - 0.5 * x;
      d_extra[K] = (0.5 * nu - 1.0) / x - 0.5;
    }
  }
  for (int k = 0; k < nq; ++k) lp += -0.5 * P.z_b[k] * P.z_b[k];
  lp -= nq * HALF_LOG_2PI;
  for (int k = 0; k < m->len_z_T; ++k) lp += -0.5 * P.z_T[k] * P.z_T[k];
  lp -= m->len_z_T * HALF_LOG_2PI;
  {
    int pos_reg = 0, pos_rho = 0;
    for (int i = 0; i < t; ++i) if (m->p[i] > 1) {
      double nu = m->regularization[pos_reg++] + 0.5 * (m->p[i] - 2);
      double r = P.rho[pos_rho++];
      lp += (nu - 1.0) * log(r) + (nu - 1.0) * log1p(-r) + lgamma(2.0 * nu) - 2.0 * lgamma(nu);
      for (int j = 2; j < m->p[i]; ++j) {            /* continuous.stan:111-115: shape1 = j / 2, shape2 = nu - (j - 1) / 2 */
        nu -= 0.5;
        double s1 = 0.5 * j, s2 = nu;
        r = P.rho[pos_rho++];
        lp += (s1 - 1.0) * log(r) + (s2 - 1.0) * log1p(-r) + lgamma(s1 + s2) - lgamma(s1) - lgamma(s2);
      }
    }
  }
  for (int i = 0; i < d->len_concentration; ++i) lp += (m->delta[i] - 1.0) * log(P.zeta[i]) - P.zeta[i] - lgamma(m->delta[i]);
  for (int i = 0; i < t; ++i) lp += (m->shape[i] - 1.0) * log(P.tau[i]) - P.tau[i] - lgamma(m->shape[i]);

  /* ---- gradient ---- */
  int pos = 0;
  double* g_zbeta = grad + pos; pos += m->len_zbeta;
  double* g_extra = grad + pos; pos += m->len_extra;
  double* g_zb = grad + pos; pos += nq;
  double* g_zT = grad + pos; pos += m->len_z_T;
  double* g_rho = grad + pos; pos += m->len_rho;
  double* g_zeta = grad + pos; pos += d->len_concentration;
  double* g_tau = grad + pos; pos += t;
  double inv_s2 = 1.0 / (sigma * sigma);
  double d_disp = 0.0;
  if (d->prior_dist <= 1) {
    for (int k = 0; k < K; ++k) {
      double dbeta = gbeta[k] * inv_s2;
      g_zbeta[k] = d->prior_dist == 0 ? dbeta : dbeta * m->prior_scale[k] - P.z_beta[k];
    }
  } else {
    /* d beta / d (z_beta, extras, aux) by the complex-step method: the map is a handful of elementary functions of K numbers */
    const double h = 1e-20;
    int nin = m->len_zbeta + m->len_extra;
    double complex* in = (double complex*) malloc(sizeof(double complex) * (size_t) (nin + K + 1));
    for (int ip = 0; ip < nin + (m->hs > 0 ? 1 : 0); ++ip) {
      for (int i = 0; i < m->len_zbeta; ++i) in[i] = P.z_beta[i];
      for (int i = 0; i < m->len_extra; ++i) in[m->len_zbeta + i] = P.extra[i];
      double complex aux = P.aux;
      if (ip < nin) in[ip] += h * I; else aux += h * I;
      coef_beta(m, in, in + m->len_zbeta, aux, in + nin);
      double acc = 0.0;
      for (int k = 0; k < K; ++k) acc += gbeta[k] * inv_s2 * (cimag(in[nin + k]) / h);
      if (ip < m->len_zbeta) g_zbeta[ip] = acc - P.z_beta[ip];
      else if (ip < nin) { int e = ip - m->len_zbeta; g_extra[e] = (acc + d_extra[e]) * P.extra[e] + 1.0; }
      else d_disp += acc;                               /* hs_prior's error_scale is aux */
    }
    free(in);
  }
  /* the Stan program declares (p - 2)(p - 1) elements of z_T per block but its onion rows consume 2 + ... + (p - 1) of them
   * through one running mark (continuous.stan:9, :41-44, :258): the surplus elements only see their normal prior */
  for (int k = 0; k < m->len_z_T; ++k) g_zT[k] = -P.z_T[k];
  int zeta_mark = 0, rho_mark = 0, th = 0, b_mark = 0, pos_reg = 0, zT_mark = 0;
  for (int i = 0; i < t; ++i) {
    if (m->p[i] > 2) {
      const int nc = m->p[i];
      const int nzT = (nc - 2) * (nc + 1) / 2;
      double Tr[OR_MAX_NC * OR_MAX_NC], A[OR_MAX_NC * OR_MAX_NC];
      for (int k = 0; k < nc * nc; ++k) { Tr[k] = 0.0; A[k] = 0.0; }
      { int thk = th; for (int c2 = 0; c2 < nc; ++c2) for (int r = c2; r < nc; ++r) Tr[r * nc + c2] = P.theta_L[thk++]; th = thk; }
      for (int j = 0; j < m->l[i]; ++j) {
        for (int k = 0; k < nc; ++k) {              /* g_zb = T' db - z ; A[r][k] += db[r] z[k] */
          double acc = 0.0;
          for (int r = k; r < nc; ++r) acc += Tr[r * nc + k] * (gb[b_mark + r] * inv_s2);
          g_zb[b_mark + k] = acc - P.z_b[b_mark + k];
        }
        for (int r = 0; r < nc; ++r) for (int k = 0; k <= r; ++k) A[r * nc + k] += (gb[b_mark + r] * inv_s2) * P.z_b[b_mark + k];
        b_mark += nc;
      }
      /* d L / d (c, zeta, rho, z_T) = sum A .* dT/dparam, dT/dparam by complex step */
      const double h = 1e-20;
      double complex Tc[OR_MAX_NC * OR_MAX_NC], zc[OR_MAX_NC], rc[OR_MAX_NC], tc[OR_MAX_NC * OR_MAX_NC];
      const double cval = P.tau[i] * m->scale[i] * P.disp;
      const int npar = 1 + nc + (nc - 1) + nzT;
      double dpar[1 + OR_MAX_NC + OR_MAX_NC + OR_MAX_NC * OR_MAX_NC];
      for (int ip = 0; ip < npar; ++ip) {
        double complex cc = cval;
        for (int k = 0; k < nc; ++k) zc[k] = P.zeta[zeta_mark + k];
        for (int k = 0; k < nc - 1; ++k) rc[k] = P.rho[rho_mark + k];
        for (int k = 0; k < nzT; ++k) tc[k] = P.z_T[zT_mark + k];
        if (ip == 0) cc += I * h;
        else if (ip < 1 + nc) zc[ip - 1] += I * h;
        else if (ip < 1 + nc + nc - 1) rc[ip - 1 - nc] += I * h;
        else tc[ip - 1 - nc - (nc - 1)] += I * h;
        block_T(nc, cc, zc, rc, tc, Tc);
        double acc = 0.0;
        for (int r = 0; r < nc; ++r) for (int k = 0; k <= r; ++k) acc += A[r * nc + k] * (cimag(Tc[r * nc + k]) / h);
        dpar[ip] = acc;
      }
      const double d_c = dpar[0];
      for (int k = 0; k < nc; ++k) {
        const double zv = P.zeta[zeta_mark + k];
        const double d_z = dpar[1 + k] + (m->delta[zeta_mark + k] - 1.0) / zv - 1.0;
        g_zeta[zeta_mark + k] = d_z * zv + 1.0;
      }
      {
        double nu = m->regularization[pos_reg++] + 0.5 * (nc - 2);
        for (int k = 0; k < nc - 1; ++k) {
          double s1, s2;
          if (k == 0) { s1 = nu; s2 = nu; } else { nu -= 0.5; s1 = 0.5 * (k + 1); s2 = nu; }
          const double rho = P.rho[rho_mark + k];
          const double d_rho = dpar[1 + nc + k] + (s1 - 1.0) / rho - (s2 - 1.0) / (1.0 - rho);
          g_rho[rho_mark + k] = d_rho * rho * (1.0 - rho) + (1.0 - 2.0 * rho);
        }
      }
      for (int k = 0; k < nzT; ++k) g_zT[zT_mark + k] += dpar[1 + nc + (nc - 1) + k];
      double d_tau = d_c * m->scale[i] * P.disp;
      d_disp += d_c * P.tau[i] * m->scale[i];
      d_tau += (m->shape[i] - 1.0) / P.tau[i] - 1.0;
      g_tau[i] = d_tau * P.tau[i] + 1.0;
      zeta_mark += nc; rho_mark += nc - 1; zT_mark += nzT;
    } else if (m->p[i] == 1) {
      double theta = P.theta_L[th++];
      double d_theta = 0.0;
      for (int s = 0; s < m->l[i]; ++s) {
        double db = gb[b_mark + s] * inv_s2;
        g_zb[b_mark + s] = theta * db - P.z_b[b_mark + s];
        d_theta += db * P.z_b[b_mark + s];
      }
      b_mark += m->l[i];
      double d_tau = d_theta * m->scale[i] * P.disp;
      d_disp += d_theta * P.tau[i] * m->scale[i];
      d_tau += (m->shape[i] - 1.0) / P.tau[i] - 1.0;
      g_tau[i] = d_tau * P.tau[i] + 1.0;
    } else {
      double T11 = P.theta_L[th], T21 = P.theta_L[th + 1], T22 = P.theta_L[th + 2]; th += 3;
      double dT11 = 0.0, dT21 = 0.0, dT22 = 0.0;
      for (int j = 0; j < m->l[i]; ++j) {
        double db0 = gb[b_mark] * inv_s2, db1 = gb[b_mark + 1] * inv_s2;
        double z0 = P.z_b[b_mark], z1 = P.z_b[b_mark + 1];
        g_zb[b_mark] = T11 * db0 + T21 * db1 - z0;
        g_zb[b_mark + 1] = T22 * db1 - z1;
        dT11 += db0 * z0; dT21 += db1 * z0; dT22 += db1 * z1;
        b_mark += 2;
      }
      double c = P.tau[i] * m->scale[i] * P.disp;
      double trace = 2.0 * c * c;
      double z1v = P.zeta[zeta_mark], z2v = P.zeta[zeta_mark + 1], zs = z1v + z2v;
      double pi1 = z1v / zs, pi2 = z2v / zs;
      double sd1 = sqrt(pi1 * trace), sd2 = sqrt(pi2 * trace);
      double rho = P.rho[rho_mark];
      double r = 2.0 * rho - 1.0, sq = sqrt(1.0 - r * r);
      double d_c = (dT11 * T11 + dT21 * T21 + dT22 * T22) / c;
      double d_sd1 = dT11, d_sd2 = dT21 * r + dT22 * sq;
      double d_pi1 = d_sd1 * sd1 / (2.0 * pi1), d_pi2 = d_sd2 * sd2 / (2.0 * pi2);
      double dot = d_pi1 * pi1 + d_pi2 * pi2;
      double d_z1 = (d_pi1 - dot) / zs, d_z2 = (d_pi2 - dot) / zs;
      double d_r = dT21 * sd2 - dT22 * sd2 * r / sq;
      double d_rho = 2.0 * d_r;
      double nu = m->regularization[pos_reg++] + 0.5 * (m->p[i] - 2);
      d_rho += (nu - 1.0) / rho - (nu - 1.0) / (1.0 - rho);
      g_rho[rho_mark] = d_rho * rho * (1.0 - rho) + (1.0 - 2.0 * rho);
      d_z1 += (m->delta[zeta_mark] - 1.0) / z1v - 1.0;
      d_z2 += (m->delta[zeta_mark + 1] - 1.0) / z2v - 1.0;
      g_zeta[zeta_mark] = d_z1 * z1v + 1.0;
      g_zeta[zeta_mark + 1] = d_z2 * z2v + 1.0;
      double d_tau = d_c * m->scale[i] * P.disp;
      d_disp += d_c * P.tau[i] * m->scale[i];
      d_tau += (m->shape[i] - 1.0) / P.tau[i] - 1.0;
      g_tau[i] = d_tau * P.tau[i] + 1.0;
      zeta_mark += 2; rho_mark += 1;
    }
  }
  if (m->has_aux) {
    double d_aux = d_disp + S / (sigma * sigma * sigma) - N / sigma;
    double d_au = d_aux * (d->prior_dist_for_aux == 0 ? 1.0 : d->prior_scale_for_aux) + d_au_prior;
    grad[pos] = d_au * P.aux_unscaled + 1.0;
  }
  *lp_out = lp;
  int bad = !isfinite(lp);
  for (int i = 0; i < m->num_params; ++i) if (!isfinite(grad[i])) bad = 1;
  free(gbeta); free(gb); free(d_extra); params_free(&P);
  return bad;
}

void or_glmm_write_array(const or_glmm* m, const double* q, double* out)
{
  const s4b_glmm_data* d = &m->d;
  Params P; params_alloc(m, &P); transform(m, q, &P);
  int pos = 0;
  for (int k = 0; k < m->len_zbeta; ++k) out[pos++] = P.z_beta[k];
  for (int k = 0; k < m->len_extra; ++k) out[pos++] = P.extra[k];
  for (int k = 0; k < d->q; ++k) out[pos++] = P.z_b[k];
  for (int k = 0; k < m->len_z_T; ++k) out[pos++] = P.z_T[k];
  for (int k = 0; k < m->len_rho; ++k) out[pos++] = P.rho[k];
  for (int k = 0; k < d->len_concentration; ++k) out[pos++] = P.zeta[k];
  for (int k = 0; k < d->t; ++k) out[pos++] = P.tau[k];
  if (m->has_aux) { out[pos++] = P.aux_unscaled; out[pos++] = P.aux; }
  for (int k = 0; k < d->K; ++k) out[pos++] = P.beta[k];
  for (int k = 0; k < d->q; ++k) out[pos++] = P.b[k];
  for (int k = 0; k < d->len_theta_L; ++k) out[pos++] = P.theta_L[k];
  params_free(&P);
}

double or_glmm_get_aux(const or_glmm* m, const double* constrained) { return constrained[m->num_params]; }

void or_glmm_parametric_mean(const or_glmm* m, const double* constrained, double* out, int include_fixed, int include_random)
{
  const s4b_glmm_data* d = &m->d;
  const double* beta = constrained + m->num_params + m->has_aux;
  const double* b = beta + d->K;
  int64_t N = d->N;
  for (int64_t i = 0; i < N; ++i) {
    double eta = 0.0;
    if (include_fixed) for (int k = 0; k < d->K; ++k) eta += m->X[(size_t) k * (size_t) N + (size_t) i] * beta[k];
    if (include_random && d->t > 0) for (int32_t z = m->u[i]; z < m->u[i + 1]; ++z) eta += m->w[z] * b[m->v[z]];
    out[i] = eta;
  }
}
