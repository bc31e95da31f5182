/*
 * oracle/oracle_gibbs.c -- TEST INFRASTRUCTURE (CPU oracle), not product code.
 *
 * Restates the alternating BART <-> Stan loop of stan4bart's host sampler:
 *   createSampler   /root/reference/src/init.cpp:190-310
 *   run             /root/reference/src/init.cpp:678-965 (loop body :752-917)
 *   disengage       /root/reference/src/init.cpp:995-1004
 * including the user offset and its offset_type variants (init.cpp:84-88, :236-252, :762-795, :831-839).
 */
#include "s4b_oracle.h"

#include <stdlib.h>
#include <string.h>

struct or_sampler {
  s4b_common_control cc;
  or_glmm* model; or_nuts* nuts; or_bart* bart;
  int64_t n, n_test; int p;
  double *bartOffset, *stanOffset, *bartLatents;
  double* stan_curr;   /* last saved Stan draw */
  int num_pars;
  double* userOffset;  /* copy of cc.user_offset or NULL */
  double* k_samples; int num_k;   /* k of every iteration of the last run */
};
enum { OFFSET_DEFAULT = 0, OFFSET_FIXEF, OFFSET_RANEF, OFFSET_BART, OFFSET_PARAMETRIC };

/* stanOffset from the tree-only training fit (init.cpp:277-285, :831-839) */
static void set_stan_offset(or_sampler* s, const double* tree_fit)
{
  size_t n = (size_t) s->n;
  if (s->userOffset && s->cc.offset_type == OFFSET_BART) memcpy(s->stanOffset, s->userOffset, sizeof(double) * n);
  else {
    memcpy(s->stanOffset, tree_fit, sizeof(double) * n);
    if (s->userOffset && s->cc.offset_type == OFFSET_DEFAULT) for (size_t j = 0; j < n; ++j) s->stanOffset[j] += s->userOffset[j];
  }
  or_glmm_set_offset(s->model, s->stanOffset);
}

or_sampler* or_sampler_create(const s4b_bart_config* bcfg, const double* y_bart, const double* x_bart, const double* x_test,
                              const s4b_glmm_data* gdata, const s4b_stan_control* sctl, const s4b_common_control* cctl,
                              const double* bart_offset_init)
{
  or_sampler* s = (or_sampler*) calloc(1, sizeof(or_sampler));
  s->cc = *cctl; s->n = bcfg->n; s->n_test = bcfg->n_test; s->p = (int) bcfg->p;
  size_t n = (size_t) s->n;
  s->model = or_glmm_create(gdata);
  if (!s->model) { free(s); return NULL; }
  s->nuts = or_nuts_create(s->model, sctl, 1, cctl->warmup);                        /* init.cpp:211-212 */
  s->num_pars = or_nuts_num_pars(s->nuts);
  s->stan_curr = (double*) calloc((size_t) s->num_pars, sizeof(double));
  s->bart = or_bart_create(bcfg, y_bart, x_bart, x_test);                           /* :215-228 */
  s->bartOffset = (double*) calloc(n ? n : 1, sizeof(double));
  s->stanOffset = (double*) calloc(n ? n : 1, sizeof(double));
  if (cctl->is_binary) s->bartLatents = (double*) calloc(n ? n : 1, sizeof(double));
  if (cctl->user_offset) { s->userOffset = (double*) malloc(sizeof(double) * (n ? n : 1)); memcpy(s->userOffset, cctl->user_offset, sizeof(double) * n); }
  s->cc.user_offset = NULL;
  if (s->userOffset && s->cc.offset_type != OFFSET_BART) {                          /* :236-247 */
    memcpy(s->bartOffset, s->userOffset, sizeof(double) * n);
    if (bart_offset_init && s->cc.offset_type == OFFSET_DEFAULT) for (size_t j = 0; j < n; ++j) s->bartOffset[j] += bart_offset_init[j];
  } else if (bart_offset_init) memcpy(s->bartOffset, bart_offset_init, sizeof(double) * n);  /* :248-252 */
  or_bart_set_offset(s->bart, s->bartOffset, 1);                                    /* :255 */
  if (!cctl->is_binary) or_bart_set_sigma(s->bart, cctl->sigma_init);               /* :256-257 */
  or_bart_sample_trees_from_prior(s->bart);                                         /* :261 */
  double* first = (double*) calloc(n ? n : 1, sizeof(double));
  or_bart_run(s->bart, first, NULL, NULL, NULL);                                    /* :273 */
  for (size_t j = 0; j < n; ++j) first[j] -= s->bartOffset[j];                      /* :275 */
  set_stan_offset(s, first);                                                        /* :277-287 */
  free(first);
  if (cctl->is_binary) { or_bart_store_latents(s->bart, s->bartLatents); or_glmm_set_response(s->model, s->bartLatents); }  /* :288-291 */
  return s;
}

void or_sampler_free(or_sampler* s)
{
  if (!s) return;
  or_bart_free(s->bart); or_nuts_free(s->nuts); or_glmm_free(s->model);
  free(s->bartOffset); free(s->stanOffset); free(s->bartLatents); free(s->stan_curr); free(s->userOffset); free(s->k_samples); free(s);
}

int or_sampler_num_stan_pars(const or_sampler* s) { return s->num_pars; }
or_bart* or_sampler_bart(or_sampler* s) { return s->bart; }
or_nuts* or_sampler_nuts(or_sampler* s) { return s->nuts; }
void or_sampler_get_range(const or_sampler* s, double* out2) { double r[3]; or_bart_get_range(s->bart, r); out2[0] = r[0]; out2[1] = r[1]; }
void or_sampler_disengage_adaptation(or_sampler* s) { or_nuts_disengage_adaptation(s->nuts); }

void or_sampler_run(or_sampler* s, int num_iter, int is_warmup, double* stan, double* train, double* test, uint32_t* varcount, double* sigma)
{
  size_t n = (size_t) s->n, nt = (size_t) s->n_test;
  double* tmp_train = (double*) calloc(n ? n : 1, sizeof(double));
  double* tmp_test = (double*) calloc(nt ? nt : 1, sizeof(double));
  free(s->k_samples);
  s->k_samples = (double*) calloc((size_t) (num_iter > 0 ? num_iter : 1), sizeof(double)); s->num_k = num_iter;
  for (int iter = 0; iter < num_iter; ++iter) {
    size_t slot = s->cc.keep_fits ? (size_t) iter : 0;
    /* A. Stan block, init.cpp:758-819 */
    or_nuts_run(s->nuts, is_warmup, s->stan_curr);
    if (stan) memcpy(stan + slot * (size_t) s->num_pars, s->stan_curr, sizeof(double) * (size_t) s->num_pars);
    if (!s->userOffset || s->cc.offset_type == OFFSET_DEFAULT || s->cc.offset_type == OFFSET_BART)   /* :762-776 */
      or_glmm_parametric_mean(s->model, s->stan_curr + 7, s->bartOffset, 1, 1);
    else if (s->cc.offset_type == OFFSET_RANEF) or_glmm_parametric_mean(s->model, s->stan_curr + 7, s->bartOffset, 1, 0);   /* :778-782 */
    else if (s->cc.offset_type == OFFSET_FIXEF) or_glmm_parametric_mean(s->model, s->stan_curr + 7, s->bartOffset, 0, 1);   /* :784-788 */
    if (s->userOffset) {
      if (s->cc.offset_type == OFFSET_PARAMETRIC) memcpy(s->bartOffset, s->userOffset, sizeof(double) * n);               /* :790-793 */
      else if (s->cc.offset_type != OFFSET_BART) for (size_t j = 0; j < n; ++j) s->bartOffset[j] += s->userOffset[j];
    }
    if (!s->cc.is_binary) or_bart_set_sigma(s->bart, or_glmm_get_aux(s->model, s->stan_curr + 7));  /* :796-800 */
    int update_scale_mod = 1 << (8 * iter / num_iter);                              /* :816 */
    or_bart_set_offset(s->bart, s->bartOffset, is_warmup && iter % update_scale_mod == 0);
    /* B. BART block, init.cpp:821-916 */
    double sig;
    or_bart_run(s->bart, tmp_train, nt ? tmp_test : NULL, varcount ? varcount + slot * (size_t) s->p : NULL, &sig);
    for (size_t j = 0; j < n; ++j) tmp_train[j] -= s->bartOffset[j];                /* :828-829 */
    set_stan_offset(s, tmp_train);                                                  /* :831-842 */
    if (s->cc.is_binary) { or_bart_store_latents(s->bart, s->bartLatents); or_glmm_set_response(s->model, s->bartLatents); }
    if (train) memcpy(train + slot * n, tmp_train, sizeof(double) * n);
    if (test && nt) memcpy(test + slot * nt, tmp_test, sizeof(double) * nt);
    if (sigma) sigma[slot] = sig;
    s->k_samples[iter] = or_bart_get_k(s->bart);
  }
  free(tmp_train); free(tmp_test);
}

int or_sampler_last_k(const or_sampler* s, double* out, int capacity)
{
  int m = s->num_k < capacity ? s->num_k : capacity;
  for (int i = 0; i < m; ++i) out[i] = s->k_samples[i];
  return m;
}
