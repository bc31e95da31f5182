/*
 * glue/gpubart_shim.h -- declarations of the dbarts C-callable table as re-implemented over the GPU sampler
 * (glue/gpubart_shim.cpp).  Each gpubart_shim_<name> has exactly the signature of BARTFunctionTable::<name>
 * (/root/reference/src/init.cpp:54-81; `storeLatents` is the registered name of the table's getLatentVariables, :1140).
 *
 * SEXP arguments: with R (-DGPUBART_SHIM_WITH_R) they are the S4 objects dbarts defines (dbartsControl, dbartsData, dbartsModel,
 * the state list); without R they point at the plain structs below, which carry the same fields.
 */
#ifndef GPUBART_SHIM_H
#define GPUBART_SHIM_H

#include <stddef.h>
#include <stdint.h>

#include "../include/stan4bart_b200.h"

#ifdef GPUBART_SHIM_WITH_R
#  include <Rinternals.h>
#else
#  ifndef R_INTERNALS_H_
typedef struct SEXPREC* SEXP;      /* opaque: points at one of the *_expr structs below */
#  endif
#endif

#ifdef __cplusplus
#  include <dbarts/bartFit.hpp>
#  include <dbarts/control.hpp>
#  include <dbarts/data.hpp>
#  include <dbarts/model.hpp>
#  include <dbarts/results.hpp>

extern "C" {
#endif

/* dbartsControl as stan4bart builds it: dbartsControl(n.chains = 1, n.samples = 1, n.burn = 0, n.thin = skip.bart, n.threads = 1,
 * updateState = FALSE, keepTrees = ..., n.trees = ..., binary filled in C++), R/stan4bart_fit.R:437-445 */
typedef struct gpubart_control_expr {
  int32_t binary, verbose, keep_training_fits, use_quantiles, keep_trees;
  int32_t n_samples, n_burn, n_trees, n_chains, n_threads, n_thin, min_obs;
  uint64_t rng_seed;
} gpubart_control_expr;

/* dbartsData: y, x, x.test, weights, offset, offset.test, n.cuts, sigma (R/lme4_functions.R:176, R/stan4bart_fit.R:446-451) */
typedef struct gpubart_data_expr {
  const double* y; const double* x; const double* x_test; const double* weights; const double* offset; const double* offset_test;
  int64_t n, p, n_test;
  const int32_t* n_cuts; int32_t n_cuts_len;     /* recycled over the predictors */
  double sigma;
} gpubart_data_expr;

/* dbartsModel from parsePriors(control, data, cgm(power, base, split.probs), normal(k) | chi(df, scale), fixed(1)) with
 * node.scale 0.5 | 3 (R/stan4bart_fit.R:456-479) */
typedef struct gpubart_model_expr {
  double p_birth_death, p_swap, p_change, p_birth;
  double node_scale;                  /* <= 0: the default for the response type */
  double base, power; const double* split_probs;
  double k, k_df, k_scale;            /* k_df > 0: k ~ chi(k_df, k_scale) */
} gpubart_model_expr;

/* the state of createStateExpression / initializeState: the stored draws of one chain; release with free() */
typedef struct gpubart_state_expr { int64_t bytes; unsigned char data[]; } gpubart_state_expr;

#ifdef __cplusplus
void gpubart_shim_initializeFit(dbarts::BARTFit* fit, dbarts::Control* control, dbarts::Model* model, dbarts::Data* data);
void gpubart_shim_invalidateFit(dbarts::BARTFit* fit);
void gpubart_shim_initializeControl(dbarts::Control* control, SEXP controlExpr);
void gpubart_shim_initializeData(dbarts::Data* data, SEXP dataExpr);
void gpubart_shim_invalidateData(dbarts::Data* data);
void gpubart_shim_initializeModel(dbarts::Model* model, SEXP modelExpr, const dbarts::Control* control, const dbarts::Data* data);
void gpubart_shim_invalidateModel(dbarts::Model* model);
SEXP gpubart_shim_createStateExpression(const dbarts::BARTFit* fit);
void gpubart_shim_initializeState(dbarts::BARTFit* fit, SEXP stateExpr);
void gpubart_shim_setControl(dbarts::BARTFit* fit, const dbarts::Control* control);
void gpubart_shim_runSamplerWithResults(dbarts::BARTFit* fit, std::size_t numBurnIn, dbarts::Results* results);
void gpubart_shim_predict(const dbarts::BARTFit* fit, const double* x_test, std::size_t numTestObservations, const double* testOffset, double* result);
void gpubart_shim_setResponse(dbarts::BARTFit* fit, const double* response);
void gpubart_shim_setOffset(dbarts::BARTFit* fit, const double* offset, bool updateScale);
void gpubart_shim_setSigma(dbarts::BARTFit* fit, const double* sigma);
void gpubart_shim_sampleTreesFromPrior(dbarts::BARTFit* fit);
void gpubart_shim_printInitialSummary(const dbarts::BARTFit* fit);
void gpubart_shim_storeLatents(const dbarts::BARTFit* fit, double* target);
void gpubart_shim_printTrees(const dbarts::BARTFit* fit, const std::size_t* chainIndices, std::size_t numChainIndices,
                             const std::size_t* sampleIndices, std::size_t numSampleIndices,
                             const std::size_t* treeIndices, std::size_t numTreeIndices);
dbarts::FlattenedTrees* gpubart_shim_getTrees(const dbarts::BARTFit* fit, const std::size_t* chainIndices, std::size_t numChainIndices,
                                              const std::size_t* sampleIndices, std::size_t numSampleIndices,
                                              const std::size_t* treeIndices, std::size_t numTreeIndices, bool useLiveTrees);
#endif

/* the stand-in for R_GetCCallable("dbarts", name): NULL for an unknown name */
void (*gpubart_shim_lookup(const char* name))(void);
int gpubart_shim_num_entries(void);
const char* gpubart_shim_entry_name(int i);
#ifdef GPUBART_SHIM_WITH_R
void gpubart_shim_register(void);
#endif

#ifdef __cplusplus
}
#endif
#endif
