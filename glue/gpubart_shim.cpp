// glue/gpubart_shim.cpp -- the dbarts C-callable table, re-implemented over the GPU sampler.
//
// stan4bart binds 20 functions of the dbarts package at load time (BARTFunctionTable, /root/reference/src/init.cpp:54-81;
// lookupBARTFunctions, :1113-1147: R_GetCCallable("dbarts", <name>)).  This file exports those 20 functions with the very
// signatures of the table, implemented over the C ABI of include/stan4bart_b200.h (gpubart_*), against the shim headers
// include/dbarts_shim/dbarts/*.hpp.  Built as stan4bart_b200/libgpubart_shim.so by stan4bart_b200/build.py.
//
// Two ways to bind it:
//   * with R:  compile with -DGPUBART_SHIM_WITH_R, call gpubart_shim_register() from an R_init_<pkg>: the functions are
//     registered with R_RegisterCCallable("dbarts", name, fn), so an UNMODIFIED stan4bart finds them where it looks for dbarts';
//   * without R (this image; tests/test_shim_gpu.py): gpubart_shim_lookup(name) is the stand-in for R_GetCCallable, and the
//     SEXP arguments of initializeControl / initializeData / initializeModel / initializeState point at the plain structs of
//     glue/gpubart_shim.h (the fields of dbartsControl / dbartsData / dbartsModel that stan4bart sets, R/stan4bart_fit.R:437-479).
//
// Errors: dbarts reports through Rf_error / ext_throwError (a longjmp out of the call); without R the shim throws
// std::runtime_error carrying s4b_last_error().
#include "gpubart_shim.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <stdexcept>
#include <string>
#include <vector>

#ifdef GPUBART_SHIM_WITH_R
#  include <R_ext/Print.h>
#  include <R_ext/Rdynload.h>
#  define SHIM_PRINTF Rprintf
#else
#  define SHIM_PRINTF std::printf
#endif

using namespace dbarts;

namespace {

[[noreturn]] void fail(const std::string& what)
{
#ifdef GPUBART_SHIM_WITH_R
  Rf_error("%s", what.c_str());
#else
  throw std::runtime_error(what);
#endif
}
void check(int rc, const char* where)
{
  if (rc != 0) fail(std::string(where) + ": " + s4b_last_error());
}
gpubart_fit* live(const BARTFit* fit)
{
  if (fit == NULL || fit->impl == NULL) fail("the fit has no live device sampler (it was created from stored state, or invalidated)");
  return static_cast<gpubart_fit*>(fit->impl);
}
void refresh_scale(BARTFit* fit)
{
  double r[3] = { -0.5, 0.5, 1.0 };
  check(gpubart_get_data_range(live(fit), r), "dataScale");
  fit->sharedScratch.dataScale.min = r[0]; fit->sharedScratch.dataScale.max = r[1]; fit->sharedScratch.dataScale.range = r[2];
}

}  // namespace

extern "C" {

// ---- initializeControl / Data / Model: the S4 objects of dbarts, or their plain stand-ins ----
void gpubart_shim_initializeControl(Control* control, SEXP controlExpr)
{
  new (control) Control();
#ifdef GPUBART_SHIM_WITH_R
  // slots of dbartsControl as stan4bart fills them (R/stan4bart_fit.R:437-445; slot names as in dbarts' R/control.R, recalled)
  control->responseIsBinary = Rf_asLogical(R_do_slot(controlExpr, Rf_install("binary"))) == TRUE;
  control->verbose = Rf_asLogical(R_do_slot(controlExpr, Rf_install("verbose"))) == TRUE;
  control->keepTrainingFits = Rf_asLogical(R_do_slot(controlExpr, Rf_install("keepTrainingFits"))) == TRUE;
  control->useQuantiles = Rf_asLogical(R_do_slot(controlExpr, Rf_install("useQuantiles"))) == TRUE;
  control->keepTrees = Rf_asLogical(R_do_slot(controlExpr, Rf_install("keepTrees"))) == TRUE;
  control->defaultNumSamples = (std::size_t) Rf_asInteger(R_do_slot(controlExpr, Rf_install("n.samples")));
  control->defaultNumBurnIn = (std::size_t) Rf_asInteger(R_do_slot(controlExpr, Rf_install("n.burn")));
  control->numTrees = (std::size_t) Rf_asInteger(R_do_slot(controlExpr, Rf_install("n.trees")));
  control->numChains = (std::size_t) Rf_asInteger(R_do_slot(controlExpr, Rf_install("n.chains")));
  control->numThreads = (std::size_t) Rf_asInteger(R_do_slot(controlExpr, Rf_install("n.threads")));
  control->treeThinningRate = (std::uint32_t) Rf_asInteger(R_do_slot(controlExpr, Rf_install("n.thin")));
  {
    const int seed = Rf_asInteger(R_do_slot(controlExpr, Rf_install("rngSeed")));
    // NA: dbarts draws from R's generator (man/stan4bart.Rd:209-228); the device stream is then seeded from it once
    control->rngSeed = seed == NA_INTEGER ? (std::uint64_t) (unif_rand() * 4294967296.0) : (std::uint64_t) (unsigned int) seed;
  }
#else
  const gpubart_control_expr* e = reinterpret_cast<const gpubart_control_expr*>(controlExpr);
  if (e == NULL) fail("initializeControl: NULL control");
  control->responseIsBinary = e->binary != 0; control->verbose = e->verbose != 0; control->keepTrainingFits = e->keep_training_fits != 0;
  control->useQuantiles = e->use_quantiles != 0; control->keepTrees = e->keep_trees != 0;
  control->defaultNumSamples = (std::size_t) e->n_samples; control->defaultNumBurnIn = (std::size_t) e->n_burn;
  control->numTrees = (std::size_t) e->n_trees; control->numChains = (std::size_t) e->n_chains; control->numThreads = (std::size_t) e->n_threads;
  control->treeThinningRate = (std::uint32_t) e->n_thin; control->rngSeed = e->rng_seed;
  if (e->min_obs > 0) control->minNumObservationsInNode = (std::uint32_t) e->min_obs;
#endif
}

void gpubart_shim_initializeData(Data* data, SEXP dataExpr)
{
  new (data) Data();
#ifdef GPUBART_SHIM_WITH_R
  SEXP x = R_do_slot(dataExpr, Rf_install("x"));
  SEXP dims = Rf_getAttrib(x, R_DimSymbol);
  data->y = REAL(R_do_slot(dataExpr, Rf_install("y")));
  data->x = REAL(x);
  data->numObservations = (std::size_t) INTEGER(dims)[0]; data->numPredictors = (std::size_t) INTEGER(dims)[1];
  SEXP xt = R_do_slot(dataExpr, Rf_install("x.test"));
  if (!Rf_isNull(xt)) { data->x_test = REAL(xt); data->numTestObservations = (std::size_t) INTEGER(Rf_getAttrib(xt, R_DimSymbol))[0]; }
  SEXP w = R_do_slot(dataExpr, Rf_install("weights")); if (!Rf_isNull(w)) data->weights = REAL(w);
  SEXP off = R_do_slot(dataExpr, Rf_install("offset")); if (!Rf_isNull(off)) data->offset = REAL(off);
  SEXP offt = R_do_slot(dataExpr, Rf_install("offset.test")); if (!Rf_isNull(offt)) data->testOffset = REAL(offt);
  data->sigmaEstimate = Rf_asReal(R_do_slot(dataExpr, Rf_install("sigma")));
  SEXP cuts = R_do_slot(dataExpr, Rf_install("n.cuts"));
  std::uint32_t* mc = new std::uint32_t[data->numPredictors];
  for (std::size_t j = 0; j < data->numPredictors; ++j) mc[j] = (std::uint32_t) INTEGER(cuts)[j % (std::size_t) XLENGTH(cuts)];
  data->maxNumCuts = mc;
#else
  const gpubart_data_expr* e = reinterpret_cast<const gpubart_data_expr*>(dataExpr);
  if (e == NULL || e->y == NULL || e->x == NULL) fail("initializeData: NULL data");
  data->y = e->y; data->x = e->x; data->x_test = e->x_test; data->weights = e->weights; data->offset = e->offset; data->testOffset = e->offset_test;
  data->numObservations = (std::size_t) e->n; data->numPredictors = (std::size_t) e->p; data->numTestObservations = e->x_test ? (std::size_t) e->n_test : 0;
  data->sigmaEstimate = e->sigma;
  std::uint32_t* mc = new std::uint32_t[data->numPredictors > 0 ? data->numPredictors : 1];
  for (std::size_t j = 0; j < data->numPredictors; ++j) mc[j] = (std::uint32_t) (e->n_cuts_len > 0 ? e->n_cuts[j % (std::size_t) e->n_cuts_len] : 100);
  data->maxNumCuts = mc;
#endif
}

void gpubart_shim_invalidateData(Data* data)
{
  if (data == NULL) return;
  delete [] data->maxNumCuts; data->maxNumCuts = NULL;
}

void gpubart_shim_initializeModel(Model* model, SEXP modelExpr, const Control* control, const Data* data)
{
  new (model) Model(false);
  model->treePrior = new TreePrior(); model->muPrior = new EndNodePrior(); model->sigmaSqPrior = new ResidualVariancePrior(); model->kPrior = new EndNodeHyperprior();
  model->nodeScale = control->responseIsBinary ? 3.0 : 0.5;
#ifdef GPUBART_SHIM_WITH_R
  // slots of dbartsModel (dbarts' R/model.R, recalled): p.birth_death, p.swap, p.change, p.birth, node.scale, tree.prior (power,
  // base, splitProbabilities), node.prior / node.hyperprior (k or chi(degreesOfFreedom, scale)), resid.prior
  model->birthOrDeathProbability = Rf_asReal(R_do_slot(modelExpr, Rf_install("p.birth_death")));
  model->swapProbability = Rf_asReal(R_do_slot(modelExpr, Rf_install("p.swap")));
  model->changeProbability = Rf_asReal(R_do_slot(modelExpr, Rf_install("p.change")));
  model->birthProbability = Rf_asReal(R_do_slot(modelExpr, Rf_install("p.birth")));
  model->nodeScale = Rf_asReal(R_do_slot(modelExpr, Rf_install("node.scale")));
  SEXP tp = R_do_slot(modelExpr, Rf_install("tree.prior"));
  model->treePrior->power = Rf_asReal(R_do_slot(tp, Rf_install("power")));
  model->treePrior->base = Rf_asReal(R_do_slot(tp, Rf_install("base")));
  SEXP sp = R_do_slot(tp, Rf_install("splitProbabilities"));
  if (!Rf_isNull(sp) && (std::size_t) XLENGTH(sp) == data->numPredictors) {
    model->treePrior->splitProbabilities = new double[data->numPredictors];
    std::memcpy(model->treePrior->splitProbabilities, REAL(sp), sizeof(double) * data->numPredictors);
  }
  SEXP hp = R_do_slot(modelExpr, Rf_install("node.hyperprior"));
  if (Rf_inherits(hp, "dbartsChiHyperprior")) {
    model->kPrior->isFixed = false;
    model->kPrior->degreesOfFreedom = Rf_asReal(R_do_slot(hp, Rf_install("degreesOfFreedom")));
    model->kPrior->scale = Rf_asReal(R_do_slot(hp, Rf_install("scale")));
  } else model->kPrior->k = model->muPrior->k = Rf_asReal(R_do_slot(hp, Rf_install("k")));
#else
  const gpubart_model_expr* e = reinterpret_cast<const gpubart_model_expr*>(modelExpr);
  if (e == NULL) fail("initializeModel: NULL model");
  model->birthOrDeathProbability = e->p_birth_death; model->swapProbability = e->p_swap; model->changeProbability = e->p_change; model->birthProbability = e->p_birth;
  if (e->node_scale > 0.0) model->nodeScale = e->node_scale;
  model->treePrior->base = e->base; model->treePrior->power = e->power;
  if (e->split_probs != NULL) {
    model->treePrior->splitProbabilities = new double[data->numPredictors];
    std::memcpy(model->treePrior->splitProbabilities, e->split_probs, sizeof(double) * data->numPredictors);
  }
  model->muPrior->k = e->k; model->kPrior->k = e->k;
  if (e->k_df > 0.0) { model->kPrior->isFixed = false; model->kPrior->degreesOfFreedom = e->k_df; model->kPrior->scale = e->k_scale; }
#endif
}

void gpubart_shim_invalidateModel(Model* model)
{
  if (model == NULL) return;
  if (model->treePrior != NULL) delete [] model->treePrior->splitProbabilities;
  delete model->treePrior; delete model->muPrior; delete model->sigmaSqPrior; delete model->kPrior;
  model->treePrior = NULL; model->muPrior = NULL; model->sigmaSqPrior = NULL; model->kPrior = NULL;
}

// ---- initializeFit / invalidateFit: the caller owns raw storage of sizeof(BARTFit) (init.cpp:227-228, :161-164) ----
void gpubart_shim_initializeFit(BARTFit* fit, Control* control, Model* model, Data* data)
{
  std::memset(static_cast<void*>(fit), 0, sizeof(BARTFit));
  fit->control = *control; fit->model = *model; fit->data = *data;
  fit->sharedScratch.dataScale.min = -0.5; fit->sharedScratch.dataScale.max = 0.5; fit->sharedScratch.dataScale.range = 1.0;
  fit->currentNumSamples = 0; fit->impl = NULL; fit->stored = NULL; fit->storeEnabled = false; fit->storeBase = 0;
  if (control->numChains > 1 && !control->keepTrees) fail("the device sampler runs one chain per fit (stan4bart sets n.chains = 1, R/stan4bart_fit.R:438)");
  if (data->y == NULL || data->x == NULL) return;           // a prediction-only fit: initializeState supplies the trees

  s4b_bart_config cfg; std::memset(&cfg, 0, sizeof cfg);
  cfg.n = (int64_t) data->numObservations; cfg.p = (int64_t) data->numPredictors; cfg.n_test = (int64_t) data->numTestObservations;
  cfg.num_trees = (int32_t) control->numTrees; cfg.thin = (int32_t) (control->treeThinningRate > 0 ? control->treeThinningRate : 1);
  cfg.min_obs = (int32_t) control->minNumObservationsInNode; cfg.is_binary = control->responseIsBinary ? 1 : 0;
  cfg.birth_death_prob = model->birthOrDeathProbability; cfg.swap_prob = model->swapProbability; cfg.change_prob = model->changeProbability;
  cfg.birth_prob = model->birthProbability; cfg.base = model->treePrior->base; cfg.power = model->treePrior->power;
  cfg.k = model->kPrior->isFixed ? model->muPrior->k : model->kPrior->k; cfg.node_scale = model->nodeScale; cfg.seed = control->rngSeed;
  cfg.split_probs = model->treePrior->splitProbabilities; cfg.weights = data->weights;
  if (!model->kPrior->isFixed) { cfg.k_df = model->kPrior->degreesOfFreedom; cfg.k_scale = model->kPrior->scale; }
  std::vector<int32_t> ncuts(data->numPredictors);
  int32_t mx = 1;
  for (std::size_t j = 0; j < data->numPredictors; ++j) { ncuts[j] = (int32_t) data->maxNumCuts[j]; if (ncuts[j] > mx) mx = ncuts[j]; }
  cfg.n_cuts = mx; cfg.n_cuts_var = ncuts.data(); cfg.use_quantiles = control->useQuantiles ? 1 : 0;
  gpubart_fit* g = NULL;
  check(gpubart_create(&cfg, data->y, data->x, data->x_test, &g), "initializeFit");
  fit->impl = g;
  if (data->offset != NULL) check(gpubart_set_offset(g, data->offset, 1), "initializeFit (offset)");
  if (control->keepTrees) {
    check(gpubart_set_keep_trees(g, (int64_t) (control->defaultNumSamples > 0 ? control->defaultNumSamples : 1)), "initializeFit (keepTrees)");
    fit->storeEnabled = true;
  }
  refresh_scale(fit);
}

void gpubart_shim_invalidateFit(BARTFit* fit)
{
  if (fit == NULL) return;
  if (fit->impl != NULL) gpubart_free(static_cast<gpubart_fit*>(fit->impl));
  if (fit->stored != NULL) gpubart_stored_free(static_cast<gpubart_stored*>(fit->stored));
  fit->impl = NULL; fit->stored = NULL;
}

// ---- setControl: stan4bart uses it for `verbose` (init.cpp:264-267, :295-296) and to switch keepTrees per run (:737-743) ----
void gpubart_shim_setControl(BARTFit* fit, const Control* control)
{
  const bool was = fit->control.keepTrees;
  if (control->numTrees != fit->control.numTrees || control->responseIsBinary != fit->control.responseIsBinary)
    fail("setControl: the number of trees and the response type are fixed when the fit is created");
  fit->control = *control;
  if (fit->impl == NULL) return;
  gpubart_fit* g = live(fit);
  if (control->keepTrees && !fit->storeEnabled) {
    // first run that keeps trees: room for the sampling-phase draws (init.cpp:218-221 sets defaultNumSamples = iter - warmup)
    check(gpubart_set_keep_trees(g, (int64_t) (control->defaultNumSamples > 0 ? control->defaultNumSamples : 1)), "setControl (keepTrees)");
    fit->storeEnabled = true; fit->currentNumSamples = 0;
  }
  if (fit->storeEnabled && control->keepTrees != was) check(gpubart_set_keep_trees_active(g, control->keepTrees ? 1 : 0), "setControl (keepTrees)");
  if (fit->storeEnabled && !control->keepTrees && !was) check(gpubart_set_keep_trees_active(g, 0), "setControl (keepTrees)");
}

// ---- the four per-iteration entries ----
void gpubart_shim_runSamplerWithResults(BARTFit* fit, std::size_t numBurnIn, Results* results)
{
  gpubart_fit* g = live(fit);
  if (fit->storeEnabled) check(gpubart_set_keep_trees_active(g, 0), "runSampler (burn-in)");
  for (std::size_t b = 0; b < numBurnIn; ++b) check(gpubart_run_sampler_with_results(g, NULL, NULL, NULL, NULL), "runSampler (burn-in)");
  if (fit->storeEnabled) check(gpubart_set_keep_trees_active(g, fit->control.keepTrees ? 1 : 0), "runSampler");
  const std::size_t n = results->numObservations, nt = results->numTestObservations, p = results->numPredictors;
  for (std::size_t s = 0; s < results->numSamples; ++s) {
    check(gpubart_run_sampler_with_results(g, results->trainingSamples + s * n, results->testSamples != NULL ? results->testSamples + s * nt : NULL,
                                           results->variableCountSamples + s * p, results->sigmaSamples + s), "runSamplerWithResults");
    if (results->kSamples != NULL) check(gpubart_get_k(g, results->kSamples + s), "runSamplerWithResults (k)");
  }
  if (fit->storeEnabled) { int64_t k = 0; check(gpubart_num_stored(g, &k), "runSamplerWithResults"); fit->currentNumSamples = (std::size_t) k; }
  refresh_scale(fit);
}

void gpubart_shim_setOffset(BARTFit* fit, const double* offset, bool updateScale)
{
  check(gpubart_set_offset(live(fit), offset, updateScale ? 1 : 0), "setOffset");
  fit->data.offset = offset;
  if (updateScale) refresh_scale(fit);
}

void gpubart_shim_setSigma(BARTFit* fit, const double* sigma)
{
  if (sigma == NULL) fail("setSigma: NULL");
  check(gpubart_set_sigma(live(fit), *sigma), "setSigma");
}

void gpubart_shim_storeLatents(const BARTFit* fit, double* target) { check(gpubart_store_latents(live(fit), target), "storeLatents"); }

// ---- the rest of the table ----
void gpubart_shim_setResponse(BARTFit* fit, const double* response)
{
  check(gpubart_set_response(live(fit), response), "setResponse");
  fit->data.y = response;
}

void gpubart_shim_sampleTreesFromPrior(BARTFit* fit) { check(gpubart_sample_trees_from_prior(live(fit)), "sampleTreesFromPrior"); }

void gpubart_shim_printInitialSummary(const BARTFit* fit)
{
  size_t need = 0;
  check(gpubart_summary(live(fit), NULL, 0, &need), "printInitialSummary");
  std::vector<char> buf(need + 1);
  check(gpubart_summary(live(fit), buf.data(), buf.size(), &need), "printInitialSummary");
  SHIM_PRINTF("%s", buf.data());
}

// predict: with stored draws (keepTrees: the StoredBARTSampler of init.cpp:354-403) one column per stored draw, expressed in the
// scale the caller left in sharedScratch.dataScale (init.cpp:433-435 sets the identity scale and R un-scales with range.bart,
// R/generics.R:671-674); otherwise the live sampler's current trees.
void gpubart_shim_predict(const BARTFit* fit, const double* x_test, std::size_t numTestObservations, const double* testOffset, double* result)
{
  const bool have_store = fit->stored != NULL || (fit->impl != NULL && fit->storeEnabled && fit->control.keepTrees);
  if (!have_store) { check(gpubart_predict(live(fit), x_test, (int64_t) numTestObservations, testOffset, result), "predict"); return; }
  const int64_t n = (int64_t) numTestObservations, S = (int64_t) fit->currentNumSamples;
  if (S == 0 || n == 0) return;
  std::vector<double> scales((size_t) (2 * S));
  if (fit->stored != NULL) {
    check(gpubart_stored_predict(static_cast<gpubart_stored*>(fit->stored), x_test, n, NULL, 0, S, result), "predict (stored)");
    check(gpubart_stored_get_scales(static_cast<gpubart_stored*>(fit->stored), 0, S, scales.data()), "predict (stored)");
  } else {
    check(gpubart_predict_stored(live(fit), x_test, n, NULL, 0, S, result), "predict (stored)");
    check(gpubart_get_stored_scales(live(fit), 0, S, scales.data()), "predict (stored)");
  }
  const DataScale& ds = fit->sharedScratch.dataScale;
  for (int64_t s = 0; s < S; ++s) {
    double* col = result + (size_t) s * (size_t) n;
    if (!fit->control.responseIsBinary) {
      const double mn = scales[(size_t) (2 * s)], rg = scales[(size_t) (2 * s + 1)];
      for (int64_t i = 0; i < n; ++i) col[i] = ds.min + (((col[i] - mn) / rg - 0.5) + 0.5) * ds.range;
    }
    if (testOffset != NULL) for (int64_t i = 0; i < n; ++i) col[i] += testOffset[i];
  }
}

// createStateExpression / initializeState: the stored draws as one opaque blob per chain (R: a list of raw vectors)
SEXP gpubart_shim_createStateExpression(const BARTFit* fit)
{
  int64_t bytes = 0;
  check(gpubart_stored_export_size(live(fit), &bytes), "createStateExpression");
#ifdef GPUBART_SHIM_WITH_R
  SEXP out = PROTECT(Rf_allocVector(VECSXP, 1));
  SEXP raw = PROTECT(Rf_allocVector(RAWSXP, (R_xlen_t) bytes));
  check(gpubart_stored_export(live(fit), RAW(raw), bytes), "createStateExpression");
  SET_VECTOR_ELT(out, 0, raw);
  UNPROTECT(2);
  return out;
#else
  gpubart_state_expr* st = static_cast<gpubart_state_expr*>(std::malloc(sizeof(gpubart_state_expr) + (size_t) bytes));
  if (st == NULL) fail("createStateExpression: out of memory");
  st->bytes = bytes;
  check(gpubart_stored_export(live(fit), st->data, bytes), "createStateExpression");
  return reinterpret_cast<SEXP>(st);
#endif
}

void gpubart_shim_initializeState(BARTFit* fit, SEXP stateExpr)
{
  const void* blob = NULL; int64_t bytes = 0;
#ifdef GPUBART_SHIM_WITH_R
  if (XLENGTH(stateExpr) != 1) Rf_error("initializeState: one chain per stored sampler on the device (got %d)", (int) XLENGTH(stateExpr));
  SEXP raw = VECTOR_ELT(stateExpr, 0);
  blob = RAW(raw); bytes = (int64_t) XLENGTH(raw);
#else
  const gpubart_state_expr* st = reinterpret_cast<const gpubart_state_expr*>(stateExpr);
  if (st == NULL) fail("initializeState: NULL state");
  blob = st->data; bytes = st->bytes;
#endif
  if (fit->stored != NULL) { gpubart_stored_free(static_cast<gpubart_stored*>(fit->stored)); fit->stored = NULL; }
  gpubart_stored* s = NULL;
  check(gpubart_stored_import(blob, bytes, &s), "initializeState");
  fit->stored = s;
  int64_t k = 0;
  check(gpubart_stored_count(s, &k), "initializeState");
  fit->currentNumSamples = (std::size_t) k;
}

namespace {
// flattened trees of the requested samples / trees.  Stored draws come from the live fit's device store; a prediction-only fit
// (imported state) cannot list trees: its blob holds them, but stan4bart extracts trees from the sampler that produced them.
struct Flat { std::vector<std::size_t> sample, tree, nobs; std::vector<std::int32_t> var; std::vector<double> value; };
static void collect(const BARTFit* fit, const std::size_t* sampleIndices, std::size_t numSampleIndices, const std::size_t* treeIndices,
             std::size_t numTreeIndices, bool live_trees, Flat& out)
{
  gpubart_fit* g = live(fit);
  const std::size_t T = fit->control.numTrees;
  std::vector<char> want(T, numTreeIndices == 0 ? 1 : 0);
  for (std::size_t i = 0; i < numTreeIndices; ++i) { if (treeIndices[i] >= T) fail("getTrees: tree index out of range"); want[treeIndices[i]] = 1; }
  const std::size_t rounds = live_trees ? 1 : numSampleIndices;
  for (std::size_t r = 0; r < rounds; ++r) {
    int64_t k = 0;
    const std::size_t smp = live_trees ? 0 : sampleIndices[r];
    if (live_trees) check(gpubart_num_nodes(g, &k), "getTrees"); else check(gpubart_num_stored_nodes(g, (int64_t) smp, &k), "getTrees");
    std::vector<int32_t> tno((size_t) k), var((size_t) k); std::vector<int64_t> nobs((size_t) k); std::vector<double> val((size_t) k);
    if (live_trees) check(gpubart_get_trees(g, tno.data(), nobs.data(), var.data(), val.data()), "getTrees");
    else check(gpubart_get_stored_trees(g, (int64_t) smp, tno.data(), nobs.data(), var.data(), val.data()), "getTrees");
    for (int64_t i = 0; i < k; ++i) if (want[(size_t) tno[(size_t) i]]) {
      out.sample.push_back(smp); out.tree.push_back((std::size_t) tno[(size_t) i]); out.nobs.push_back((std::size_t) nobs[(size_t) i]);
      out.var.push_back(var[(size_t) i]); out.value.push_back(val[(size_t) i]);
    }
  }
}
}  // namespace

FlattenedTrees* gpubart_shim_getTrees(const BARTFit* fit, const std::size_t* chainIndices, std::size_t numChainIndices,
                                      const std::size_t* sampleIndices, std::size_t numSampleIndices,
                                      const std::size_t* treeIndices, std::size_t numTreeIndices, bool useLiveTrees)
{
  for (std::size_t i = 0; i < numChainIndices; ++i) if (chainIndices[i] != 0) fail("getTrees: one chain per fit on the device");
  const bool live_trees = useLiveTrees || !(fit->storeEnabled && fit->control.keepTrees);
  Flat f;
  collect(fit, sampleIndices, numSampleIndices, treeIndices, numTreeIndices, live_trees, f);
  FlattenedTrees* out = static_cast<FlattenedTrees*>(::operator new(sizeof(FlattenedTrees)));
  const std::size_t m = f.tree.size();
  out->totalNumNodes = m;
  out->chainNumber = new std::size_t[m > 0 ? m : 1]; out->sampleNumber = new std::size_t[m > 0 ? m : 1]; out->treeNumber = new std::size_t[m > 0 ? m : 1];
  out->numObservations = new std::size_t[m > 0 ? m : 1]; out->variable = new std::int32_t[m > 0 ? m : 1]; out->value = new double[m > 0 ? m : 1];
  for (std::size_t i = 0; i < m; ++i) {
    out->chainNumber[i] = 0; out->sampleNumber[i] = f.sample[i]; out->treeNumber[i] = f.tree[i]; out->numObservations[i] = f.nobs[i];
    out->variable[i] = f.var[i]; out->value[i] = f.value[i];
  }
  return out;
}

// printTrees: one line per node, children indented under their parent (pre-order), e.g.
//   tree 3, sample 1
//     x5 <= 0.4123 (n = 80)
//       mu = -0.0312 (n = 31)
void gpubart_shim_printTrees(const BARTFit* fit, const std::size_t* chainIndices, std::size_t numChainIndices,
                             const std::size_t* sampleIndices, std::size_t numSampleIndices,
                             const std::size_t* treeIndices, std::size_t numTreeIndices)
{
  for (std::size_t i = 0; i < numChainIndices; ++i) if (chainIndices[i] != 0) fail("printTrees: one chain per fit on the device");
  const bool live_trees = !(fit->storeEnabled && fit->control.keepTrees);
  Flat f;
  collect(fit, sampleIndices, numSampleIndices, treeIndices, numTreeIndices, live_trees, f);
  std::vector<int> open;          // children still to come per open internal node
  for (std::size_t i = 0; i < f.tree.size(); ++i) {
    if (open.empty()) {
      if (live_trees) SHIM_PRINTF("tree %zu\n", f.tree[i] + 1); else SHIM_PRINTF("tree %zu, sample %zu\n", f.tree[i] + 1, f.sample[i] + 1);
    }
    for (std::size_t d = 0; d <= open.size(); ++d) SHIM_PRINTF("  ");
    if (f.var[i] >= 0) { SHIM_PRINTF("x%d <= %.6g (n = %zu)\n", (int) f.var[i] + 1, f.value[i], f.nobs[i]); open.push_back(2); }
    else {
      SHIM_PRINTF("mu = %.6g (n = %zu)\n", f.value[i], f.nobs[i]);
      while (!open.empty() && --open.back() == 0) open.pop_back();
    }
  }
}

// ---- binding ----
struct ShimEntry { const char* name; void (*fn)(void); };
#define ENTRY(n, f) { n, reinterpret_cast<void (*)(void)>(f) }
static const ShimEntry kTable[] = {
  ENTRY("initializeFit", gpubart_shim_initializeFit), ENTRY("invalidateFit", gpubart_shim_invalidateFit),
  ENTRY("initializeControl", gpubart_shim_initializeControl), ENTRY("initializeData", gpubart_shim_initializeData),
  ENTRY("invalidateData", gpubart_shim_invalidateData), ENTRY("initializeModel", gpubart_shim_initializeModel),
  ENTRY("invalidateModel", gpubart_shim_invalidateModel), ENTRY("createStateExpression", gpubart_shim_createStateExpression),
  ENTRY("initializeState", gpubart_shim_initializeState), ENTRY("setControl", gpubart_shim_setControl),
  ENTRY("runSamplerWithResults", gpubart_shim_runSamplerWithResults), ENTRY("predict", gpubart_shim_predict),
  ENTRY("setResponse", gpubart_shim_setResponse), ENTRY("setOffset", gpubart_shim_setOffset), ENTRY("setSigma", gpubart_shim_setSigma),
  ENTRY("sampleTreesFromPrior", gpubart_shim_sampleTreesFromPrior), ENTRY("printInitialSummary", gpubart_shim_printInitialSummary),
  ENTRY("storeLatents", gpubart_shim_storeLatents), ENTRY("printTrees", gpubart_shim_printTrees), ENTRY("getTrees", gpubart_shim_getTrees),
};

// stand-in for R_GetCCallable("dbarts", name) (init.cpp:1115-1146)
void (*gpubart_shim_lookup(const char* name))(void)
{
  for (std::size_t i = 0; i < sizeof kTable / sizeof kTable[0]; ++i) if (std::strcmp(kTable[i].name, name) == 0) return kTable[i].fn;
  return NULL;
}
int gpubart_shim_num_entries(void) { return (int) (sizeof kTable / sizeof kTable[0]); }
const char* gpubart_shim_entry_name(int i) { return (i >= 0 && i < gpubart_shim_num_entries()) ? kTable[i].name : NULL; }

#ifdef GPUBART_SHIM_WITH_R
// call from R_init_<package>: an unmodified stan4bart then binds these where it looks for dbarts' C callables
void gpubart_shim_register(void)
{
  for (std::size_t i = 0; i < sizeof kTable / sizeof kTable[0]; ++i) R_RegisterCCallable("dbarts", kTable[i].name, reinterpret_cast<DL_FUNC>(kTable[i].fn));
}
#endif

}  // extern "C"
