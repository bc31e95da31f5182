/*
 * include/stan4bart_b200.h -- C ABI of the B200-native stan4bart Gibbs hot path.
 *
 * Plain pointers and sizes only.  Every entry point names the reference interface it
 * replaces (paths relative to /root/reference).  All functions return 0 on success and a
 * non-zero code on failure; s4b_last_error() returns the message (thread local).  There is
 * no CPU fallback: without a CUDA device every compute entry point fails.
 *
 * Three layers, mirroring the reference's own boundaries:
 *   gpubart_*  : the dbarts C-callable table stan4bart binds at load time
 *                (BARTFunctionTable, src/init.cpp:54-81, lookup :1113-1147)
 *   glmm_*     : the Stan model / gradient hook (stan::model::gradient,
 *                src/include/stan/model/gradient.hpp:21-35; continuous_model members
 *                src/stan_files/continuous.hpp:3626-3768)
 *   s4b_sampler_* : the .Call routines of the host sampler (src/init.cpp:1215-1229)
 */
#ifndef STAN4BART_B200_H
#define STAN4BART_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define S4B_TRACE_RECORD_LEN 32

/* dbarts Control + Model as configured by R/stan4bart_fit.R:437-479 */
typedef struct s4b_bart_config {
  int64_t n, p, n_test;
  int32_t num_trees;       /* n.trees */
  int32_t n_cuts;          /* n.cuts, uniform cut points, <= 255 */
  int32_t thin;            /* n.thin = skip.bart */
  int32_t min_obs;         /* minNumObservationsInNode (5) */
  int32_t is_binary;
  int32_t max_ctas;        /* 0 = the whole GPU; > 0 confines this chain's sweep kernel to that many SMs (several chains per GPU) */
  double birth_death_prob, swap_prob, change_prob, birth_prob;   /* .5 .1 .4 .5 */
  double base, power;      /* cgm tree prior */
  double k;                /* normal(k) leaf prior */
  double node_scale;       /* .5 continuous, 3 binary (R/stan4bart_fit.R:479) */
  uint64_t seed;
  /* bart_args split.probs (R/stan4bart_fit.R:466, tests/testthat/test-09-bartArgs.R:20-39): relative probability of every
   * predictor when a splitting variable is drawn and in the rule prior; NULL = uniform.  Normalised internally to integer
   * weights summing to ~2^30, so that the CPU oracle and the device select and weigh variables identically. */
  const double* split_probs;
  /* observation weights (`weights` of stan4bart(), handed to dbarts as data weights: R/stan4bart_fit.R:449): y_i ~ N(f(x_i),
   * sigma^2 / w_i), leaf statistics become sum w, sum w r; n rows, finite and >= 0; NULL = unweighted.  Weighted fits run the
   * persistent sweep kernel with two-value bins (sum w r, sum w). */
  const double* weights;
  /* bart_args k = chi(degreesOfFreedom, scale) (the `!kPrior->isFixed` of src/init.cpp:731): k_df > 0 => k is sampled after every
   * sweep, starting from `k`; k_scale <= 0 or infinite => the improper chi(df, Inf).  k_df = 0 => fixed k. */
  double k_df, k_scale;
  /* bart_args n.cuts given per predictor (R/stan4bart_fit.R:446-451 recycles it over the columns): p entries in [1, n_cuts],
   * n_cuts being the largest; NULL = n_cuts for every predictor */
  const int32_t* n_cuts_var;
  /* change rule: 0 (default) = Metropolis-Hastings ratio including the proposal term |I_new| P(old var) / (|I_old| P(new var))
   * (the cut of the new rule is drawn from the interval that ancestors and descendants leave for the NEW variable, the reverse
   * move draws from the interval of the OLD one): the chain then has the model's exact posterior as its stationary law, which
   * tests/test_exact_posterior.py checks by brute-force enumeration.  1 = prior x likelihood ratio only, the form the change
   * step is remembered to have in dbarts / BayesTree (not verifiable here: dbarts is not vendored); exact only for p = 1. */
  int32_t change_symmetric;
  /* bart_args use.quantiles (R/stan4bart_fit.R:437-451 passes it on to dbartsControl): cut points between the distinct sorted values of
   * every predictor (all gaps when there are at most n.cuts + 1 distinct values, else n.cuts of them evenly spaced in rank) instead of
   * uniform over the range; observation-sharded chains exchange their distinct values at create, so every rank gets the cuts of the whole column */
  int32_t use_quantiles;
} s4b_bart_config;

/* the `data.stan` list, R/stan4bart_fit.R:259-365 / src/stan_sampler.cpp:112-380 (default path) */
typedef struct s4b_glmm_data {
  int64_t N;
  int32_t K, is_binary, prior_dist, prior_dist_for_aux, t, q, len_theta_L, len_concentration, len_regularization, reserved;
  int64_t num_non_zero;
  const double* X; const double* y; const double* prior_scale; const double* prior_mean;
  double prior_scale_for_aux, prior_mean_for_aux, prior_df_for_aux;
  const int32_t* p; const int32_t* l; const double* shape; const double* scale;
  const double* concentration; const double* regularization;
  const double* w; const int32_t* v; const int32_t* u;
  /* data.stan `has_weights` / `weights` (R/stan4bart_fit.R:255-262, continuous.stan:358-366): NULL = unweighted */
  const double* weights;
  /* non-default coefficient priors (continuous.stan:184-186, :207-215): prior_dist 2 student_t (prior_df), 3 hs / 4 hs_plus
   * (prior_df, global_prior_df, global_prior_scale, slab_df, slab_scale; hs_plus reads prior_scale as a second df), 5 laplace,
   * 6 lasso (prior_df[0]), 7 product_normal (num_normals, K entries >= 2).  prior_df / num_normals may be NULL when unused. */
  const double* prior_df; const int32_t* num_normals;
  double global_prior_df, global_prior_scale, slab_df, slab_scale;
} s4b_glmm_data;

/* StanControl, src/stan_sampler.hpp:28-42 */
typedef struct s4b_stan_control {
  uint32_t seed; int32_t skip;
  double init_radius, adapt_gamma, adapt_delta, adapt_kappa, adapt_t0;
  uint32_t adapt_init_buffer, adapt_term_buffer, adapt_window; int32_t max_treedepth;
  double stepsize, stepsize_jitter;
} s4b_stan_control;

/* control.common, src/init.cpp:1015-1051 */
/* commonControl (init.cpp:1010-1040).  user_offset / offset_type: the `offset` argument of stan4bart() and its debugging
 * variants (init.cpp:84-88, :236-252, :762-795, :831-839): 0 default (added to both halves' offsets), 1 fixef / 2 ranef (the
 * user vector stands in for that part of the parametric mean), 3 bart (it replaces the BART fit seen by Stan), 4 parametric
 * (it replaces the whole Stan part seen by BART).  user_offset NULL = none. */
typedef struct s4b_common_control {
  int32_t warmup, iter, is_binary, keep_fits;
  double sigma_init;
  int32_t offset_type, reserved;
  const double* user_offset;
} s4b_common_control;

const char* s4b_last_error(void);
int s4b_device_count(void);
/* cudaSetDevice for this library's runtime instance (one process per GPU: call once with LOCAL_RANK) */
int s4b_set_device(int device);
/* all later objects created on this thread use `cuda_stream` (a cudaStream_t; NULL = a private stream) */
int s4b_set_stream(void* cuda_stream);

/* ------------------------------------------------------------------ gpubart_* */
typedef struct gpubart_fit gpubart_fit;
/* initializeFit + initializeControl/Data/Model (init.cpp:215-228); x, x_test column major */
int gpubart_create(const s4b_bart_config* cfg, const double* y, const double* x, const double* x_test, gpubart_fit** out);
/* invalidateFit (init.cpp:161-164) */
int gpubart_free(gpubart_fit* fit);
/* setOffset(fit, offset, updateScale) (init.cpp:255, :817) */
int gpubart_set_offset(gpubart_fit* fit, const double* offset, int update_scale);
/* setResponse(fit, response) (BARTFunctionTable, init.cpp:68): a new response vector, n entries */
int gpubart_set_response(gpubart_fit* fit, const double* y);
/* setSigma (init.cpp:257, :799) */
int gpubart_set_sigma(gpubart_fit* fit, double sigma);
/* current k of the leaf prior (a draw when k is modelled) */
int gpubart_get_k(gpubart_fit* fit, double* k);
/* sampleTreesFromPrior (init.cpp:261) */
int gpubart_sample_trees_from_prior(gpubart_fit* fit);
/* runSamplerWithResults(fit, 0, results[numSamples = 1]) (init.cpp:273, :824); layout src/bart_util.hpp:14-35 */
int gpubart_run_sampler_with_results(gpubart_fit* fit, double* train, double* test, uint32_t* varcount, double* sigma);
/* Several chains, one launch (grid.y = chain): one runSamplerWithResults step of `count` fits whose sweep kernels
 * run as ONE cooperative launch, chain c on the CTAs of row c.  The fits must have been created on the same thread / stream with the
 * same shape class (rows per thread, predictors, `thin`) and max_ctas = SMs / count (or less); traced, replayed, weighted and sharded
 * fits are refused.  The chains evolve exactly as if run one by one.  Results of each fit: gpubart_collect_results (any pointer NULL). */
int gpubart_run_batched(gpubart_fit* const* fits, int count);
int gpubart_collect_results(gpubart_fit* fit, double* train, double* test, uint32_t* varcount, double* sigma);
/* storeLatents / getLatentVariables (init.cpp:289, :845) */
int gpubart_store_latents(gpubart_fit* fit, double* out);
/* fit->sharedScratch.dataScale.{min,max,range} (init.cpp:324-325) */
int gpubart_get_data_range(gpubart_fit* fit, double* min_max_range);
/* predict(fit, x_test, n, testOffset, result) (init.cpp:398) */
int gpubart_predict(gpubart_fit* fit, const double* x_test, int64_t n, const double* test_offset, double* out);
/* getTrees -> FlattenedTrees (init.cpp:577-666): pre-order, var < 0 => leaf, value = cut point | leaf mu */
/* keepTrees (dbarts control$keepTrees; stan4bart_exportBARTState / stan4bart_createStoredBARTSampler / stan4bart_predictBART on a
 * stored sampler, init.cpp:354-446; stan4bart_getTrees on stored samples, :514-671).  With a store of `capacity` draws every
 * runSamplerWithResults call (and every s4b_sampler_run iteration) appends the trees of its kept draw; capacity 0 frees it.
 * predict_stored: out [n x count], draws first .. first + count - 1; get_stored_trees: the flattened table of one draw. */
int gpubart_set_keep_trees(gpubart_fit* fit, int64_t capacity);
/* setControl(fit, control) with control->keepTrees toggled (init.cpp:737-744: off for warm-up runs, on for sampling): whether
 * runSamplerWithResults appends to the store */
int gpubart_set_keep_trees_active(gpubart_fit* fit, int on);
/* (min, range) of the response scale each stored draw's leaf values are expressed in; out2 has 2 * count entries */
int gpubart_get_stored_scales(gpubart_fit* fit, int64_t first, int64_t count, double* out2);
int gpubart_num_stored(gpubart_fit* fit, int64_t* out);
int gpubart_predict_stored(gpubart_fit* fit, const double* x_test, int64_t n, const double* test_offset, int64_t first, int64_t count, double* out);
int gpubart_num_stored_nodes(gpubart_fit* fit, int64_t sample, int64_t* out);
int gpubart_get_stored_trees(gpubart_fit* fit, int64_t sample, int32_t* tree_no, int64_t* n_obs, int32_t* var, double* value);
/* stan4bart_exportBARTState -> a host blob (cut points + stored draws); stan4bart_createStoredBARTSampler -> a prediction-only
 * object rebuilt from such a blob in any later process (R: the state saved inside the fitted object, R/generics.R:183-190) */
typedef struct gpubart_stored gpubart_stored;
int gpubart_stored_export_size(gpubart_fit* fit, int64_t* bytes);
int gpubart_stored_export(gpubart_fit* fit, void* out, int64_t bytes);
int gpubart_stored_import(const void* blob, int64_t bytes, gpubart_stored** out);
int gpubart_stored_free(gpubart_stored* st);
int gpubart_stored_count(gpubart_stored* st, int64_t* out);
int gpubart_stored_get_scales(gpubart_stored* st, int64_t first, int64_t count, double* out2);
int gpubart_stored_predict(gpubart_stored* st, const double* x_test, int64_t n, const double* test_offset, int64_t first, int64_t count, double* out);
/* printInitialSummary (dbarts table entry called by stan4bart_printInitialSummary, init.cpp:981): the text of the summary,
 * including the "power and base for tree prior:" and "tree split probabilities:" lines that tests/testthat/test-09-bartArgs.R
 * parses.  *needed = bytes including the terminator; out may be NULL to query. */
int gpubart_summary(gpubart_fit* fit, char* out, size_t cap, size_t* needed);
int gpubart_num_nodes(gpubart_fit* fit, int64_t* out);
int gpubart_get_trees(gpubart_fit* fit, int32_t* tree_no, int64_t* n_obs, int32_t* var, double* value);
/* parity instrumentation (no reference counterpart) */
int gpubart_node_assignment(gpubart_fit* fit, int tree, int64_t* heap_index);
int gpubart_leaf_stats(gpubart_fit* fit, int tree, int max_leaves, int64_t* heap_index, int64_t* count, double* sum, double* sumsq, int* num_leaves);
int gpubart_get_residual(gpubart_fit* fit, double* out);
int gpubart_set_trace(gpubart_fit* fit, size_t cap_records);
int gpubart_get_trace(gpubart_fit* fit, double* out, size_t cap_records, size_t* num_records);
int gpubart_set_tape(gpubart_fit* fit, const double* tape, size_t len);
int gpubart_set_record(gpubart_fit* fit, size_t cap);
int gpubart_get_record(gpubart_fit* fit, double* out, size_t cap, size_t* len);
int gpubart_rng_counter(gpubart_fit* fit, uint64_t* out);
int gpubart_set_use_graph(gpubart_fit* fit, int use_graph);
/* 0: one kernel launch per tree step; 1: those launches captured in a CUDA graph; 2: persistent on-chip sweep kernel
 * (default when the chain fits the register file + shared memory) */
int gpubart_set_sweep_mode(gpubart_fit* fit, int mode);
int gpubart_get_sweep_mode(gpubart_fit* fit, int* mode);
/* the software-pipelined sweep kernel (csrc/sweep_pipe.cuh): on by default for unweighted, unsharded, register-resident chains
 * without parity trace / replay.  Decided per tree step on the device: runs of steps whose trees are small enough go through it,
 * the steps in between through the synchronous kernel.  get: whether it is enabled, sweeps offered to it, tree steps it ran */
int gpubart_set_pipeline(gpubart_fit* fit, int on);
int gpubart_get_pipeline(gpubart_fit* fit, int* enabled, int64_t* sweeps_launched, int64_t* steps_pipelined);
/* tree steps that did not fit the pipelined kernel so far: [0] all, [1] tree too large, [2] more than 8 statistic slots, [3] too many cells */
int gpubart_pipeline_misfits(gpubart_fit* fit, uint32_t* out4);
/* micro-benchmark hook: `reps` launches of the leaf-statistics pass for `tree`, returns mean ms per launch */
int gpubart_time_leaf_stats(gpubart_fit* fit, int tree, int reps, double* ms_per_launch);
int gpubart_num_tree_steps(gpubart_fit* fit, int64_t* out);
/* device milliseconds (CUDA events on the launching stream) spent in the per-tree kernels since the last reset */
int gpubart_tree_step_ms(gpubart_fit* fit, int reset, double* ms);
/* SM-cycle counters of the controller phases of k_tree_step: [0] pass of the last block, [1] partial reduction, [2] tree load,
 * [3] MH decision + leaf draws, [4] write-back + next tree load, [5] next proposal / update, [6] descriptor publish, [7] steps counted,
 * [8..15] finer controller counters of the persistent kernel (see sweep_kernel.cuh); [16..19] worker sub-phases; out24 has 24 entries */
int gpubart_get_profile(gpubart_fit* fit, uint64_t* out24, int reset);
/* the cycle counters cost a few percent of the controller's time: off by default, switch on before the sweeps to profile */
int gpubart_set_profile(gpubart_fit* fit, int on);

/* ------------------------------------------------------------------ glmm_* */
typedef struct glmm_model glmm_model;
/* createStanModelFromExpression (src/stan_sampler.cpp:112-380) */
int glmm_create(const s4b_glmm_data* data, glmm_model** out);
int glmm_free(glmm_model* m);
int glmm_num_params(glmm_model* m, int* d, int* num_constrained);
/* set_offset / set_response (continuous.hpp:3626-3635), host pointers */
int glmm_set_offset(glmm_model* m, const double* offset);
int glmm_set_response(glmm_model* m, const double* y);
/* stan::model::gradient (model/gradient.hpp:21-35): status != 0 <=> exception path (V = +inf) */
int glmm_log_prob_grad(glmm_model* m, const double* q, double* lp, double* grad, int* status);
/* write_array (continuous.hpp:2640-2938) */
int glmm_write_array(glmm_model* m, const double* q, double* out);
/* names of the stored Stan rows ('\n' separated; lp__ ... energy__, then constrained_param_names, continuous.hpp:3115-3204):
 * the dimnames of the reference's `stan` result (src/stan_sampler.cpp:478-489, :577-596).  *needed = bytes incl. terminator */
int glmm_stan_row_names(glmm_model* m, char* out, size_t cap, size_t* needed);
/* get_parametric_mean (continuous.hpp:3662-3768) */
int glmm_parametric_mean(glmm_model* m, const double* constrained, double* out, int include_fixed, int include_random);
/* device data terms only: S = sum e^2, X'e, Z'e */
int glmm_data_terms(glmm_model* m, const double* beta, const double* b, double* S, double* gbeta, double* gb);
int glmm_num_grad_evals(glmm_model* m, int64_t* out);
/* 0: one device pass per evaluation (the literal north-star path); 1 (default when K + q <= 512): one device pass per Gibbs
 * sweep anchors an exact quadratic expansion of the data terms in (beta, b), every further evaluation of the sweep is O((K+q)^2)
 * on the host */
int glmm_set_mode(glmm_model* m, int mode);
int glmm_get_mode(glmm_model* m, int* mode);
int glmm_num_device_passes(glmm_model* m, int64_t* out);
/* measurement: device milliseconds of the data pass alone (CUDA events on the model's stream, mean of `reps` launches); flush_l2 != 0
 * evicts the L2 before every timed launch; *bulk (optional) = 1 when the pass is the bulk-copy (TMA) kernel */
int glmm_time_data_pass(glmm_model* m, int reps, int flush_l2, double* ms, int* bulk);

/* The Stan half alone: the reference's StanSampler over a model (src/stan_sampler.cpp:382-476 constructor: diag_e NUTS with the
 * windowed adaptation of `ctl`; run(isWarmup) :478-489: `ctl.skip` transitions, the last one written to out[num_pars] -- lp__,
 * accept_stat__, stepsize__, treedepth__, n_leapfrog__, divergent__, energy__, then the constrained parameters).  The model must
 * outlive the sampler.  Used by the Gibbs sampler below; exposed for tests of the NUTS control against exact posteriors. */
typedef struct glmm_nuts glmm_nuts;
int glmm_nuts_create(glmm_model* m, const s4b_stan_control* ctl, int chain_id, int num_warmup, glmm_nuts** out);
int glmm_nuts_free(glmm_nuts* s);
int glmm_nuts_num_pars(glmm_nuts* s, int* out);
int glmm_nuts_run(glmm_nuts* s, int is_warmup, double* out);
int glmm_nuts_disengage_adaptation(glmm_nuts* s);
int glmm_nuts_stepsize(glmm_nuts* s, double* out);

/* ------------------------------------------------------------------ s4b_sampler_* */
typedef struct s4b_sampler s4b_sampler;
/* stan4bart_create (init.cpp:190-310) */
int s4b_sampler_create(const s4b_bart_config* bcfg, const double* y_bart, const double* x_bart, const double* x_test,
                       const s4b_glmm_data* gdata, const s4b_stan_control* sctl, const s4b_common_control* cctl,
                       const double* bart_offset_init, s4b_sampler** out);
/* stan4bart_finalize */
int s4b_sampler_free(s4b_sampler* s);
int s4b_sampler_num_stan_pars(s4b_sampler* s, int* out);
/* stan4bart_run(sampler, numIter, isWarmup, "both") (init.cpp:678-965).  Output buffers may be NULL.
 * stan [num_pars x S], train [n x S], test [n_test x S], varcount [p x S], sigma [S]; S = keep_fits ? num_iter : 1 */
/* Several chains of one GPU with their BART sweeps batched into ONE launch per Gibbs iteration (grid.y = chain).
 * Create a group for `count` samplers (all on the same device, each created with max_ctas = SMs / count and the same BART shape class),
 * attach every sampler, then call s4b_sampler_run for each of them from its OWN host thread with the same num_iter: at its BART block a
 * chain waits for the others, the last one enqueues k_sweep_batch for all.  The chains advance in lock-step; each evolves exactly as on
 * its own with the synchronous sweep kernel.  A chain that stops early times the others out (error) instead of hanging them. */
typedef struct s4b_batch_group s4b_batch_group;
int s4b_batch_group_create(int count, s4b_batch_group** out);
int s4b_batch_group_free(s4b_batch_group* g);
int s4b_batch_group_launches(s4b_batch_group* g, int64_t* out);
int s4b_sampler_set_batch_group(s4b_sampler* s, s4b_batch_group* g);   /* NULL detaches */
int s4b_sampler_run(s4b_sampler* s, int num_iter, int is_warmup, double* stan, double* train, double* test, uint32_t* varcount, double* sigma);
/* k of every iteration of the last s4b_sampler_run (the `k` row of the reference's BART results when k is modelled,
 * src/bart_util.hpp:25); *count = how many were written (<= capacity) */
int s4b_sampler_last_k(s4b_sampler* s, double* out, int capacity, int* count);
/* the per-iteration callback of stan4bart_run (init.cpp:849-911: an R closure evaluated with yhat.train, yhat.test and the Stan
 * draw of the iteration in scope).  `fn` is called on the host after every iteration with host copies of the iteration's
 * training fit [n], test fit [n_test] (NULL when there is no test sample) and Stan row [num_pars]; a non-zero return value
 * stops the run with an error.  fn == NULL removes the callback. */
typedef int (*s4b_iteration_callback)(void* user, int iteration, const double* stan_row, const double* yhat_train, const double* yhat_test);
int s4b_sampler_set_callback(s4b_sampler* s, s4b_iteration_callback fn, void* user);
/* stan4bart_disengageAdaptation (init.cpp:995-1004) */
int s4b_sampler_disengage_adaptation(s4b_sampler* s);
/* stan4bart_getBARTDataRange (init.cpp:316-330) */
int s4b_sampler_get_bart_data_range(s4b_sampler* s, double* out2);
/* stan4bart_getParametricMean (init.cpp:332-347) */
int s4b_sampler_get_parametric_mean(s4b_sampler* s, double* out);
/* stan4bart_predictBART (init.cpp:354-403), on the live sampler's current trees */
int s4b_sampler_predict_bart(s4b_sampler* s, const double* x_test, int64_t n, const double* test_offset, double* out);
gpubart_fit* s4b_sampler_bart(s4b_sampler* s);
glmm_model* s4b_sampler_glmm(s4b_sampler* s);
/* running posterior means kept on device (keep_fits = FALSE path): mean train / test BART fit
 * and mean parametric part over the sampling-phase sweeps run so far */
int s4b_sampler_get_means(s4b_sampler* s, double* mean_bart_train, double* mean_bart_test, double* mean_parametric, int64_t* num_draws);
/* route every iteration's N-length vectors (parametric mean, BART fit, latents) through pinned host memory, as the
 * reference's host-side vectors bartOffset / stanOffset / bartLatents do (init.cpp:143-145); reports bytes per iteration */
int s4b_sampler_set_host_plumbing(s4b_sampler* s, int on, int64_t* h2d_bytes_per_iter, int64_t* d2h_bytes_per_iter);
/* timing split of the last run: milliseconds in the Stan block and the BART block (CUDA events), leapfrog count */
int s4b_sampler_last_run_stats(s4b_sampler* s, double* ms_stan, double* ms_bart, int64_t* n_grad_evals, int64_t* n_tree_steps);

/* ------------------------------------------------------------------ s4b_shard_*  (observation-sharded chains)
 * No counterpart in the reference: stan4bart runs one chain on one host thread (R/stan4bart.R:288-300 fans whole chains
 * out with parallel::clusterMap).  BASELINE config E / SURVEY.md 8(e) shard the ROWS of one chain over the GPUs of an
 * NVSwitch box: one process per GPU, trees / RNG / NUTS state replicated, per-node statistics and GLMM reductions
 * exchanged through peer-mapped mailboxes inside the kernels (csrc/shard.hpp).  Set-up: every rank creates a context,
 * publishes its 64-byte IPC handle (any transport: torch.distributed all_gather, MPI, a file), attaches the world's
 * handles, states its row range, then creates the sampler with the *_sharded constructors.  All ranks must make the
 * same sequence of calls with the same seeds. */
typedef struct s4b_shard s4b_shard;
int s4b_shard_create(int rank, int world, s4b_shard** out);
int s4b_shard_free(s4b_shard* sh);
int s4b_shard_ipc_handle(s4b_shard* sh, unsigned char* out64);
/* handles: world x 64 bytes, ordered by rank (the own entry is ignored) */
int s4b_shard_attach(s4b_shard* sh, const unsigned char* handles);
/* this rank holds rows [first_obs, first_obs + n_local) of total_obs */
int s4b_shard_set_obs_range(s4b_shard* sh, int64_t first_obs, int64_t total_obs);
/* in-place all-reduce of a host vector over the ranks (op 0 = sum in rank order, 1 = max); collective */
int s4b_shard_allreduce(s4b_shard* sh, double* vec, int64_t n, int op);
/* Reference collective ("NCCL ncclAllReduce as reference implementation"): the small all-reduces of a sharded chain
 * (cut-point ranges, response range of the rescale, the (1 + K + q) GLMM reductions, the Gram matrix, s4b_shard_allreduce) through
 * ncclAllReduce over NVLink instead of the peer-mapped mailboxes.  Rank 0 obtains the 128-byte unique id, every rank passes it to
 * nccl_init (collective), use_nccl switches the path.  NCCL is loaded at run time.  The per-tree-step exchange stays inside the
 * sweep kernel (NCCL's launch latency per call is several times the whole tree step); tests check both paths against each other. */
int s4b_shard_nccl_unique_id(unsigned char* out128);
int s4b_shard_nccl_init(s4b_shard* sh, const unsigned char* id128);
int s4b_shard_use_nccl(s4b_shard* sh, int on);
/* initializeFit / glmm / stan4bart_create on this rank's rows; n, N are the local row counts */
int gpubart_create_sharded(const s4b_bart_config* cfg, const double* y, const double* x, const double* x_test, s4b_shard* sh, gpubart_fit** out);
int glmm_create_sharded(const s4b_glmm_data* d, s4b_shard* sh, glmm_model** out);
int s4b_sampler_create_sharded(const s4b_bart_config* bcfg, const double* y_bart, const double* x_bart, const double* x_test,
                               const s4b_glmm_data* gdata, const s4b_stan_control* sctl, const s4b_common_control* cctl,
                               const double* bart_offset_init, s4b_shard* sh, s4b_sampler** out);

#ifdef __cplusplus
}
#endif
#endif
