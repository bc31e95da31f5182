/*
 * oracle/oracle_nuts.c -- TEST INFRASTRUCTURE (CPU oracle), not product code.
 *
 * Plain-C restatement of the NUTS machinery stan4bart drives
 * (/root/reference/src/interruptable_sampler.hpp:118-210), following the vendored
 * Stan 2.28 headers under /root/reference/src/include/stan:
 *   mcmc/hmc/nuts/base_nuts.hpp:78-204 (transition), :247-352 (build_tree), :222-226 (criterion)
 *   mcmc/hmc/nuts/adapt_diag_e_nuts.hpp:25-49
 *   mcmc/hmc/base_hmc.hpp:81-143 (init_stepsize), :175-180 (sample_stepsize)
 *   mcmc/hmc/hamiltonians/diag_e_metric.hpp:20-50, base_hamiltonian.hpp:61-70
 *   mcmc/hmc/integrators/base_leapfrog.hpp:17-22, expl_leapfrog.hpp:16-31
 *   mcmc/stepsize_adaptation.hpp:49-73, var_adaptation.hpp:17-46, windowed_adaptation.hpp:23-110
 *   math/prim/fun/welford_var_estimator.hpp:10-45
 *   services/util/initialize.hpp:60-216, generate_transitions.hpp:43-77, mcmc_writer.hpp:97-129
 * Randomness is s4b-rng v1 (stream 1), NOT boost::ecuyer1988: the reference's own
 * NUTS draws cannot be replayed here ("parity unpinned" vs the reference binary).
 * What does pin this file: exact posterior moments by quadrature of an independent
 * transcription of the density (tests/test_exact_stan_posterior.py).
 */
#include "s4b_oracle.h"
#include "s4b_rng.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct { double *q, *p, *g; double V; } PsPoint;

struct or_nuts {
  or_glmm* model;
  s4b_stan_control ctl;
  int d, num_constrained;
  s4b_rng rng;
  PsPoint z;
  double* inv_metric;
  double nom_epsilon, epsilon, epsilon_jitter;
  int max_depth; double max_deltaH;
  int depth, n_leapfrog, divergent; double energy;
  /* stepsize adaptation */
  double sa_counter, sa_s_bar, sa_x_bar, sa_mu, sa_delta, sa_gamma, sa_kappa, sa_t0;
  /* windowed variance adaptation (unsigned arithmetic as in the reference) */
  uint32_t num_warmup, init_buffer, term_buffer, base_window, window_counter, next_window, window_size;
  double wf_n; double *wf_m, *wf_m2;
  int adapt_flag;
  double* cont_params; double lp, accept_stat;
  int64_t num_grad;
};

static void ps_alloc(PsPoint* z, int d) { z->q = (double*) calloc((size_t) d + 1, sizeof(double)); z->p = (double*) calloc((size_t) d + 1, sizeof(double)); z->g = (double*) calloc((size_t) d + 1, sizeof(double)); z->V = 0.0; }
static void ps_free(PsPoint* z) { free(z->q); free(z->p); free(z->g); }
static void ps_copy(PsPoint* dst, const PsPoint* src, int d) { memcpy(dst->q, src->q, sizeof(double) * (size_t) d); memcpy(dst->p, src->p, sizeof(double) * (size_t) d); memcpy(dst->g, src->g, sizeof(double) * (size_t) d); dst->V = src->V; }
static double* vec_new(int d) { return (double*) calloc((size_t) d + 1, sizeof(double)); }

static void update_potential_gradient(or_nuts* s, PsPoint* z)
{
  double lp; double* grad = vec_new(s->d);
  int bad = or_glmm_log_prob_grad(s->model, z->q, &lp, grad);
  s->num_grad++;
  if (!bad) { z->V = -lp; for (int i = 0; i < s->d; ++i) z->g[i] = -grad[i]; }
  else { z->V = INFINITY; for (int i = 0; i < s->d; ++i) z->g[i] = -z->g[i]; }   /* stale g negated, base_hamiltonian.hpp:61-70 */
  free(grad);
}
static double kinetic(const or_nuts* s, const PsPoint* z) { double t = 0.0; for (int i = 0; i < s->d; ++i) t += z->p[i] * s->inv_metric[i] * z->p[i]; return 0.5 * t; }
static double hamiltonian(const or_nuts* s, const PsPoint* z) { return kinetic(s, z) + z->V; }
static void sample_p(or_nuts* s, PsPoint* z) { for (int i = 0; i < s->d; ++i) z->p[i] = s4b_rng_normal(&s->rng) / sqrt(s->inv_metric[i]); }
static void dtau_dp(const or_nuts* s, const PsPoint* z, double* out) { for (int i = 0; i < s->d; ++i) out[i] = s->inv_metric[i] * z->p[i]; }

static void evolve(or_nuts* s, PsPoint* z, double eps)
{
  int d = s->d;
  for (int i = 0; i < d; ++i) z->p[i] -= 0.5 * eps * z->g[i];
  for (int i = 0; i < d; ++i) z->q[i] += eps * s->inv_metric[i] * z->p[i];
  update_potential_gradient(s, z);
  for (int i = 0; i < d; ++i) z->p[i] -= 0.5 * eps * z->g[i];
}

static void init_stepsize(or_nuts* s)
{
  int d = s->d;
  PsPoint z_init; ps_alloc(&z_init, d); ps_copy(&z_init, &s->z, d);
  if (s->nom_epsilon == 0 || s->nom_epsilon > 1e7 || isnan(s->nom_epsilon)) { ps_free(&z_init); return; }
  sample_p(s, &s->z); update_potential_gradient(s, &s->z);
  double H0 = hamiltonian(s, &s->z);
  evolve(s, &s->z, s->nom_epsilon);
  double h = hamiltonian(s, &s->z); if (isnan(h)) h = INFINITY;
  double delta_H = H0 - h;
  int direction = delta_H > log(0.8) ? 1 : -1;
  for (;;) {
    ps_copy(&s->z, &z_init, d);
    sample_p(s, &s->z); update_potential_gradient(s, &s->z);
    H0 = hamiltonian(s, &s->z);
    evolve(s, &s->z, s->nom_epsilon);
    h = hamiltonian(s, &s->z); if (isnan(h)) h = INFINITY;
    delta_H = H0 - h;
    if (direction == 1 && !(delta_H > log(0.8))) break;
    else if (direction == -1 && !(delta_H < log(0.8))) break;
    else s->nom_epsilon = direction == 1 ? 2.0 * s->nom_epsilon : 0.5 * s->nom_epsilon;
    if (s->nom_epsilon > 1e7 || s->nom_epsilon == 0) break;   /* reference throws */
  }
  ps_copy(&s->z, &z_init, d);
  ps_free(&z_init);
}

static int criterion(const double* ps_minus, const double* ps_plus, const double* rho, int d)
{
  double a = 0.0, b = 0.0;
  for (int i = 0; i < d; ++i) { a += ps_plus[i] * rho[i]; b += ps_minus[i] * rho[i]; }
  return a > 0 && b > 0;
}
static double log_sum_exp2(double a, double b)
{
  if (a == -INFINITY) return b;
  if (a == INFINITY && b == INFINITY) return INFINITY;
  if (a > b) return a + log1p(exp(b - a));
  return b + log1p(exp(a - b));
}

static int build_tree(or_nuts* s, int depth, PsPoint* z_propose, double* p_sharp_beg, double* p_sharp_end, double* rho,
                      double* p_beg, double* p_end, double H0, double sign, int* n_leapfrog, double* log_sum_weight, double* sum_metro_prob)
{
  int d = s->d;
  if (depth == 0) {
    evolve(s, &s->z, sign * s->epsilon);
    ++(*n_leapfrog);
    double h = hamiltonian(s, &s->z); if (isnan(h)) h = INFINITY;
    if ((h - H0) > s->max_deltaH) s->divergent = 1;
    *log_sum_weight = log_sum_exp2(*log_sum_weight, H0 - h);
    if (H0 - h > 0) *sum_metro_prob += 1; else *sum_metro_prob += exp(H0 - h);
    ps_copy(z_propose, &s->z, d);
    dtau_dp(s, &s->z, p_sharp_beg); memcpy(p_sharp_end, p_sharp_beg, sizeof(double) * (size_t) d);
    for (int i = 0; i < d; ++i) rho[i] += s->z.p[i];
    memcpy(p_beg, s->z.p, sizeof(double) * (size_t) d); memcpy(p_end, p_beg, sizeof(double) * (size_t) d);
    return !s->divergent;
  }
  double log_sum_weight_init = -INFINITY;
  double *p_init_end = vec_new(d), *p_sharp_init_end = vec_new(d), *rho_init = vec_new(d);
  int valid_init = build_tree(s, depth - 1, z_propose, p_sharp_beg, p_sharp_init_end, rho_init, p_beg, p_init_end, H0, sign, n_leapfrog, &log_sum_weight_init, sum_metro_prob);
  if (!valid_init) { free(p_init_end); free(p_sharp_init_end); free(rho_init); return 0; }
  PsPoint z_propose_final; ps_alloc(&z_propose_final, d); ps_copy(&z_propose_final, &s->z, d);
  double log_sum_weight_final = -INFINITY;
  double *p_final_beg = vec_new(d), *p_sharp_final_beg = vec_new(d), *rho_final = vec_new(d);
  int valid_final = build_tree(s, depth - 1, &z_propose_final, p_sharp_final_beg, p_sharp_end, rho_final, p_final_beg, p_end, H0, sign, n_leapfrog, &log_sum_weight_final, sum_metro_prob);
  int result = 0;
  if (valid_final) {
    double log_sum_weight_subtree = log_sum_exp2(log_sum_weight_init, log_sum_weight_final);
    *log_sum_weight = log_sum_exp2(*log_sum_weight, log_sum_weight_subtree);
    if (log_sum_weight_final > log_sum_weight_subtree) ps_copy(z_propose, &z_propose_final, d);
    else {
      double accept_prob = exp(log_sum_weight_final - log_sum_weight_subtree);
      if (s4b_rng_uniform(&s->rng) < accept_prob) ps_copy(z_propose, &z_propose_final, d);
    }
    double* rho_subtree = vec_new(d);
    for (int i = 0; i < d; ++i) { rho_subtree[i] = rho_init[i] + rho_final[i]; rho[i] += rho_subtree[i]; }
    int persist = criterion(p_sharp_beg, p_sharp_end, rho_subtree, d);
    for (int i = 0; i < d; ++i) rho_subtree[i] = rho_init[i] + p_final_beg[i];
    persist &= criterion(p_sharp_beg, p_sharp_final_beg, rho_subtree, d);
    for (int i = 0; i < d; ++i) rho_subtree[i] = rho_final[i] + p_init_end[i];
    persist &= criterion(p_sharp_init_end, p_sharp_end, rho_subtree, d);
    free(rho_subtree);
    result = persist;
  }
  ps_free(&z_propose_final);
  free(p_init_end); free(p_sharp_init_end); free(rho_init); free(p_final_beg); free(p_sharp_final_beg); free(rho_final);
  return result;
}

static void base_transition(or_nuts* s)
{
  int d = s->d;
  /* sample_stepsize */
  s->epsilon = s->nom_epsilon;
  if (s->epsilon_jitter) s->epsilon *= 1.0 + s->epsilon_jitter * (2.0 * s4b_rng_uniform(&s->rng) - 1.0);
  memcpy(s->z.q, s->cont_params, sizeof(double) * (size_t) d);
  sample_p(s, &s->z);
  update_potential_gradient(s, &s->z);
  PsPoint z_fwd, z_bck, z_sample, z_propose;
  ps_alloc(&z_fwd, d); ps_alloc(&z_bck, d); ps_alloc(&z_sample, d); ps_alloc(&z_propose, d);
  ps_copy(&z_fwd, &s->z, d); ps_copy(&z_bck, &s->z, d); ps_copy(&z_sample, &s->z, d); ps_copy(&z_propose, &s->z, d);
  double *p_fwd_fwd = vec_new(d), *p_sharp_fwd_fwd = vec_new(d), *p_fwd_bck = vec_new(d), *p_sharp_fwd_bck = vec_new(d);
  double *p_bck_fwd = vec_new(d), *p_sharp_bck_fwd = vec_new(d), *p_bck_bck = vec_new(d), *p_sharp_bck_bck = vec_new(d);
  double *rho = vec_new(d), *rho_fwd = vec_new(d), *rho_bck = vec_new(d), *rho_ext = vec_new(d);
  size_t nb = sizeof(double) * (size_t) d;
  memcpy(p_fwd_fwd, s->z.p, nb); dtau_dp(s, &s->z, p_sharp_fwd_fwd);
  memcpy(p_fwd_bck, s->z.p, nb); memcpy(p_sharp_fwd_bck, p_sharp_fwd_fwd, nb);
  memcpy(p_bck_fwd, s->z.p, nb); memcpy(p_sharp_bck_fwd, p_sharp_fwd_fwd, nb);
  memcpy(p_bck_bck, s->z.p, nb); memcpy(p_sharp_bck_bck, p_sharp_fwd_fwd, nb);
  memcpy(rho, s->z.p, nb);
  double log_sum_weight = 0.0;
  double H0 = hamiltonian(s, &s->z);
  int n_leapfrog = 0; double sum_metro_prob = 0.0;
  s->depth = 0; s->divergent = 0;
  while (s->depth < s->max_depth) {
    memset(rho_fwd, 0, nb); memset(rho_bck, 0, nb);
    int valid_subtree = 0;
    double log_sum_weight_subtree = -INFINITY;
    if (s4b_rng_uniform(&s->rng) > 0.5) {
      ps_copy(&s->z, &z_fwd, d);
      memcpy(rho_bck, rho, nb); memcpy(p_bck_fwd, p_fwd_fwd, nb); memcpy(p_sharp_bck_fwd, p_sharp_fwd_fwd, nb);
      valid_subtree = build_tree(s, s->depth, &z_propose, p_sharp_fwd_bck, p_sharp_fwd_fwd, rho_fwd, p_fwd_bck, p_fwd_fwd, H0, 1.0, &n_leapfrog, &log_sum_weight_subtree, &sum_metro_prob);
      ps_copy(&z_fwd, &s->z, d);
    } else {
      ps_copy(&s->z, &z_bck, d);
      memcpy(rho_fwd, rho, nb); memcpy(p_fwd_bck, p_bck_bck, nb); memcpy(p_sharp_fwd_bck, p_sharp_bck_bck, nb);
      valid_subtree = build_tree(s, s->depth, &z_propose, p_sharp_bck_fwd, p_sharp_bck_bck, rho_bck, p_bck_fwd, p_bck_bck, H0, -1.0, &n_leapfrog, &log_sum_weight_subtree, &sum_metro_prob);
      ps_copy(&z_bck, &s->z, d);
    }
    if (!valid_subtree) break;
    ++s->depth;
    if (log_sum_weight_subtree > log_sum_weight) ps_copy(&z_sample, &z_propose, d);
    else {
      double accept_prob = exp(log_sum_weight_subtree - log_sum_weight);
      if (s4b_rng_uniform(&s->rng) < accept_prob) ps_copy(&z_sample, &z_propose, d);
    }
    log_sum_weight = log_sum_exp2(log_sum_weight, log_sum_weight_subtree);
    for (int i = 0; i < d; ++i) rho[i] = rho_bck[i] + rho_fwd[i];
    int persist = criterion(p_sharp_bck_bck, p_sharp_fwd_fwd, rho, d);
    for (int i = 0; i < d; ++i) rho_ext[i] = rho_bck[i] + p_fwd_bck[i];
    persist &= criterion(p_sharp_bck_bck, p_sharp_fwd_bck, rho_ext, d);
    for (int i = 0; i < d; ++i) rho_ext[i] = rho_fwd[i] + p_bck_fwd[i];
    persist &= criterion(p_sharp_bck_fwd, p_sharp_fwd_fwd, rho_ext, d);
    if (!persist) break;
  }
  s->n_leapfrog = n_leapfrog;
  double accept_prob = sum_metro_prob / (double) n_leapfrog;
  ps_copy(&s->z, &z_sample, d);
  s->energy = hamiltonian(s, &s->z);
  memcpy(s->cont_params, s->z.q, nb);
  s->lp = -s->z.V; s->accept_stat = accept_prob;
  ps_free(&z_fwd); ps_free(&z_bck); ps_free(&z_sample); ps_free(&z_propose);
  free(p_fwd_fwd); free(p_sharp_fwd_fwd); free(p_fwd_bck); free(p_sharp_fwd_bck);
  free(p_bck_fwd); free(p_sharp_bck_fwd); free(p_bck_bck); free(p_sharp_bck_bck);
  free(rho); free(rho_fwd); free(rho_bck); free(rho_ext);
}

/* ---- adaptation ---- */
static void window_restart(or_nuts* s) { s->window_counter = 0; s->window_size = s->base_window; s->next_window = s->init_buffer + s->window_size - 1u; }
static void set_window_params(or_nuts* s, uint32_t num_warmup, uint32_t init_buffer, uint32_t term_buffer, uint32_t base_window)
{
  if (num_warmup < 20u) return;
  if (init_buffer + base_window + term_buffer > num_warmup) {
    s->num_warmup = num_warmup;
    s->init_buffer = (uint32_t) (0.15 * num_warmup);
    s->term_buffer = (uint32_t) (0.10 * num_warmup);
    s->base_window = num_warmup - (s->init_buffer + s->term_buffer);
    return;    /* no restart() here in the reference (windowed_adaptation.hpp:49-75) */
  }
  s->num_warmup = num_warmup; s->init_buffer = init_buffer; s->term_buffer = term_buffer; s->base_window = base_window;
  window_restart(s);
}
static int adaptation_window(const or_nuts* s) { return s->window_counter >= s->init_buffer && s->window_counter < s->num_warmup - s->term_buffer && s->window_counter != s->num_warmup; }
static int end_adaptation_window(const or_nuts* s) { return s->window_counter == s->next_window && s->window_counter != s->num_warmup; }
static void compute_next_window(or_nuts* s)
{
  if (s->next_window == s->num_warmup - s->term_buffer - 1u) return;
  s->window_size *= 2u;
  s->next_window = s->window_counter + s->window_size;
  if (s->next_window == s->num_warmup - s->term_buffer - 1u) return;
  uint32_t next_window_boundary = s->next_window + 2u * s->window_size;
  if (next_window_boundary >= s->num_warmup - s->term_buffer) s->next_window = s->num_warmup - s->term_buffer - 1u;
}
static int learn_variance(or_nuts* s)
{
  int d = s->d;
  if (adaptation_window(s)) {
    s->wf_n += 1.0;
    for (int i = 0; i < d; ++i) { double delta = s->z.q[i] - s->wf_m[i]; s->wf_m[i] += delta / s->wf_n; s->wf_m2[i] += delta * (s->z.q[i] - s->wf_m[i]); }
  }
  if (end_adaptation_window(s)) {
    compute_next_window(s);
    double n = s->wf_n;
    for (int i = 0; i < d; ++i) {
      double var = s->inv_metric[i];
      if (n > 1.0) var = s->wf_m2[i] / (n - 1.0);
      s->inv_metric[i] = (n / (n + 5.0)) * var + 1e-3 * (5.0 / (n + 5.0));
    }
    s->wf_n = 0.0; memset(s->wf_m, 0, sizeof(double) * (size_t) d); memset(s->wf_m2, 0, sizeof(double) * (size_t) d);
    ++s->window_counter;
    return 1;
  }
  ++s->window_counter;
  return 0;
}
static void learn_stepsize(or_nuts* s, double adapt_stat)
{
  s->sa_counter += 1.0;
  adapt_stat = adapt_stat > 1 ? 1 : adapt_stat;
  double eta = 1.0 / (s->sa_counter + s->sa_t0);
  s->sa_s_bar = (1.0 - eta) * s->sa_s_bar + eta * (s->sa_delta - adapt_stat);
  double x = s->sa_mu - s->sa_s_bar * sqrt(s->sa_counter) / s->sa_gamma;
  double x_eta = pow(s->sa_counter, -s->sa_kappa);
  s->sa_x_bar = (1.0 - x_eta) * s->sa_x_bar + x_eta * x;
  s->nom_epsilon = exp(x);
}

static void adapt_transition(or_nuts* s)
{
  base_transition(s);
  if (s->adapt_flag) {
    learn_stepsize(s, s->accept_stat);
    if (learn_variance(s)) {
      init_stepsize(s);
      s->sa_mu = log(10 * s->nom_epsilon);
      s->sa_counter = 0; s->sa_s_bar = 0; s->sa_x_bar = 0;
    }
  }
}

or_nuts* or_nuts_create(or_glmm* model, const s4b_stan_control* ctl, int chain_id, int num_warmup)
{
  or_nuts* s = (or_nuts*) calloc(1, sizeof(or_nuts));
  s->model = model; s->ctl = *ctl; s->d = or_glmm_num_params(model); s->num_constrained = or_glmm_num_constrained(model);
  int d = s->d;
  s4b_rng_init(&s->rng, (uint64_t) ctl->seed | ((uint64_t) (uint32_t) chain_id << 32), S4B_STREAM_STAN);
  ps_alloc(&s->z, d);
  s->inv_metric = vec_new(d); s->wf_m = vec_new(d); s->wf_m2 = vec_new(d); s->cont_params = vec_new(d);
  /* services/util/initialize.hpp: uniform(-R, R) on the unconstrained scale, <= 100 tries */
  double* grad = vec_new(d);
  for (int attempt = 0; attempt < 100; ++attempt) {
    for (int i = 0; i < d; ++i) s->cont_params[i] = ctl->init_radius == 0.0 ? 0.0 : -ctl->init_radius + 2.0 * ctl->init_radius * s4b_rng_uniform(&s->rng);
    double lp; int bad = or_glmm_log_prob_grad(model, s->cont_params, &lp, grad); s->num_grad++;
    if (!bad) break;
  }
  free(grad);
  for (int i = 0; i < d; ++i) s->inv_metric[i] = 1.0;
  s->nom_epsilon = 0.1; s->epsilon = 0.1; s->epsilon_jitter = 0.0; s->max_depth = 5; s->max_deltaH = 1000;
  if (ctl->stepsize > 0) s->nom_epsilon = ctl->stepsize;
  if (ctl->stepsize_jitter > 0 && ctl->stepsize_jitter < 1) s->epsilon_jitter = ctl->stepsize_jitter;
  if (ctl->max_treedepth > 0) s->max_depth = ctl->max_treedepth;
  s->sa_mu = 0.5; s->sa_delta = 0.5; s->sa_gamma = 0.05; s->sa_kappa = 0.75; s->sa_t0 = 10;
  s->sa_mu = log(10 * ctl->stepsize);
  if (ctl->adapt_delta > 0 && ctl->adapt_delta < 1) s->sa_delta = ctl->adapt_delta;
  if (ctl->adapt_gamma > 0) s->sa_gamma = ctl->adapt_gamma;
  if (ctl->adapt_kappa > 0) s->sa_kappa = ctl->adapt_kappa;
  if (ctl->adapt_t0 > 0) s->sa_t0 = ctl->adapt_t0;
  /* windowed_adaptation ctor: all zero then restart() */
  s->num_warmup = s->init_buffer = s->term_buffer = s->base_window = 0; window_restart(s);
  set_window_params(s, (uint32_t) (num_warmup * ctl->skip), ctl->adapt_init_buffer, ctl->adapt_term_buffer, ctl->adapt_window);
  s->adapt_flag = 1;
  memcpy(s->z.q, s->cont_params, sizeof(double) * (size_t) d);
  init_stepsize(s);
  return s;
}

void or_nuts_free(or_nuts* s) { if (!s) return; ps_free(&s->z); free(s->inv_metric); free(s->wf_m); free(s->wf_m2); free(s->cont_params); free(s); }
int or_nuts_num_pars(const or_nuts* s) { return 7 + s->num_constrained; }

void or_nuts_run(or_nuts* s, int warmup, double* out)
{
  (void) warmup;
  for (int m = 0; m < s->ctl.skip; ++m) adapt_transition(s);
  if (out) {
    out[0] = s->lp; out[1] = s->accept_stat; out[2] = s->epsilon; out[3] = s->depth; out[4] = s->n_leapfrog; out[5] = s->divergent; out[6] = s->energy;
    or_glmm_write_array(s->model, s->cont_params, out + 7);
  }
}

void or_nuts_disengage_adaptation(or_nuts* s) { s->adapt_flag = 0; s->nom_epsilon = exp(s->sa_x_bar); }
double or_nuts_stepsize(const or_nuts* s) { return s->nom_epsilon; }
void or_nuts_get_metric(const or_nuts* s, double* m) { memcpy(m, s->inv_metric, sizeof(double) * (size_t) s->d); }
void or_nuts_get_q(const or_nuts* s, double* q) { memcpy(q, s->cont_params, sizeof(double) * (size_t) s->d); }
int64_t or_nuts_num_grad_evals(const or_nuts* s) { return s->num_grad; }
