// glue/init_gpu.cpp -- the R `.Call` layer of stan4bart over the GPU sampler.
//
// Registers the same 12 routines as /root/reference/src/init.cpp:1215-1229, with the same names, arities, argument meaning
// and result shapes, implemented over the C ABI of include/stan4bart_b200.h (s4b_sampler_* / gpubart_*).  With this file
// compiled into the package's shared object in place of src/init.cpp, the unmodified R code (R/stan4bart_fit.R:42-56, :579;
// R/generics.R:190, :667) drives the CUDA path.  The alternative, finer-grained binding -- keeping src/init.cpp and Stan on
// the host and replacing only dbarts -- is glue/gpubart_shim.cpp.
//
// R is not available in the build image: this file is compiled with -fsyntax-only against the declarations of
// tests/r_stub/ (tests/test_glue_cpu.py); it has never been run under R.  Needs -DGPUBART_SHIM_WITH_R for the S4 parsers of
// glue/gpubart_shim.cpp, which it reuses for dbartsControl / dbartsData / dbartsModel.
#ifndef GPUBART_SHIM_WITH_R
#  define GPUBART_SHIM_WITH_R 1
#endif
#include "gpubart_shim.h"

#include <R_ext/Print.h>
#include <R_ext/Rdynload.h>
#include <R_ext/Random.h>
#include <R_ext/Utils.h>

#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace {

struct GpuSampler {
  s4b_sampler* h;
  int defaultWarmup, defaultIter, verbose, refresh, numPars;
  bool responseIsBinary, keepFits, keepTrees;
  std::size_t n, nTest, p;
  bool kIsModeled;
  SEXP callback, callbackEnv;              // kept alive by the external pointer's protected slot
  std::vector<std::string> rowNames;
  const double* userOffset; int offsetType;
  dbarts::Control bartControl; dbarts::Data bartData; dbarts::Model bartModel;
  GpuSampler() : h(NULL), bartModel(false) { }
};
struct StoredSampler { gpubart_stored* st; bool binary; std::size_t numPredictors; };

SEXP listElement(SEXP list, const char* name)
{
  SEXP names = Rf_getAttrib(list, R_NamesSymbol);
  if (Rf_isNull(names)) return R_NilValue;
  for (R_xlen_t i = 0; i < XLENGTH(list); ++i) if (std::strcmp(CHAR(STRING_ELT(names, i)), name) == 0) return VECTOR_ELT(list, i);
  return R_NilValue;
}
int getInt(SEXP list, const char* name, int dflt) { SEXP e = listElement(list, name); if (Rf_isNull(e) || XLENGTH(e) == 0) return dflt; int v = Rf_asInteger(e); return v == NA_INTEGER ? dflt : v; }
double getReal(SEXP list, const char* name, double dflt) { SEXP e = listElement(list, name); if (Rf_isNull(e) || XLENGTH(e) == 0) return dflt; double v = Rf_asReal(e); return ISNAN(v) ? dflt : v; }
bool getBool(SEXP list, const char* name, bool dflt) { SEXP e = listElement(list, name); if (Rf_isNull(e) || XLENGTH(e) == 0) return dflt; int v = Rf_asLogical(e); return v == NA_LOGICAL ? dflt : v == TRUE; }
const double* getRealVec(SEXP list, const char* name) { SEXP e = listElement(list, name); return (Rf_isNull(e) || XLENGTH(e) == 0 || !Rf_isReal(e)) ? NULL : REAL(e); }
const int* getIntVec(SEXP list, const char* name) { SEXP e = listElement(list, name); return (Rf_isNull(e) || XLENGTH(e) == 0 || !Rf_isInteger(e)) ? NULL : INTEGER(e); }
void check(int rc) { if (rc != 0) Rf_error("%s", s4b_last_error()); }

GpuSampler* samplerOf(SEXP ptr, const char* who)
{
  GpuSampler* s = static_cast<GpuSampler*>(R_ExternalPtrAddr(ptr));
  if (s == NULL) Rf_error("%s called on NULL external pointer", who);
  return s;
}

void samplerFinalizer(SEXP ptr)
{
  GpuSampler* s = static_cast<GpuSampler*>(R_ExternalPtrAddr(ptr));
  if (s == NULL) return;
  if (s->h != NULL) s4b_sampler_free(s->h);
  gpubart_shim_invalidateModel(&s->bartModel); gpubart_shim_invalidateData(&s->bartData);
  delete s;
  R_ClearExternalPtr(ptr);
}
void storedFinalizer(SEXP ptr)
{
  StoredSampler* s = static_cast<StoredSampler*>(R_ExternalPtrAddr(ptr));
  if (s == NULL) return;
  if (s->st != NULL) gpubart_stored_free(s->st);
  delete s;
  R_ClearExternalPtr(ptr);
}

// the per-iteration callback of stan4bart_run (init.cpp:849-911): an R closure evaluated with (yhat.train, yhat.test, stan_pars)
struct CallbackState { SEXP closure, yhatTrain, yhatTest, stanPars, results; std::size_t n, nTest; int numPars; int numIter; R_xlen_t resultLength; };
int iterationCallback(void* user, int iteration, const double* stanRow, const double* yhatTrain, const double* yhatTest)
{
  CallbackState& cb = *static_cast<CallbackState*>(user);
  R_CheckUserInterrupt();                                   // the reference checks once per transition (src/stan_sampler.hpp:44-48)
  if (cb.closure == R_NilValue) return 0;
  std::memcpy(REAL(cb.yhatTrain), yhatTrain, cb.n * sizeof(double));
  if (cb.nTest > 0 && yhatTest != NULL) std::memcpy(REAL(cb.yhatTest), yhatTest, cb.nTest * sizeof(double));
  std::memcpy(REAL(cb.stanPars), stanRow, (std::size_t) cb.numPars * sizeof(double));
  SEXP value = PROTECT(Rf_eval(cb.closure, R_GlobalEnv));
  if (cb.results == R_NilValue) {                           // first call fixes the length of a result
    cb.resultLength = XLENGTH(value);
    cb.results = Rf_allocMatrix(REALSXP, (int) cb.resultLength, cb.numIter);
    R_PreserveObject(cb.results);
  }
  if (XLENGTH(value) == cb.resultLength) {
    SEXP dbl = PROTECT(Rf_coerceVector(value, REALSXP));
    std::memcpy(REAL(cb.results) + (std::size_t) iteration * (std::size_t) cb.resultLength, REAL(dbl), (std::size_t) cb.resultLength * sizeof(double));
    UNPROTECT(1);
  }
  UNPROTECT(1);
  return 0;
}

SEXP namedList(int n, const char* const* names)
{
  SEXP out = PROTECT(Rf_allocVector(VECSXP, n));
  SEXP nm = PROTECT(Rf_allocVector(STRSXP, n));
  for (int i = 0; i < n; ++i) SET_STRING_ELT(nm, i, Rf_mkChar(names[i]));
  Rf_setAttrib(out, R_NamesSymbol, nm);
  UNPROTECT(2);
  return out;
}

}  // namespace

extern "C" {

// stan4bart_create(bartControl, bartData, bartModel, stanData, stanControl, commonControl) -> external pointer   (init.cpp:190-310)
static SEXP createSampler(SEXP bartControlExpr, SEXP bartDataExpr, SEXP bartModelExpr, SEXP stanDataExpr, SEXP stanControlExpr, SEXP commonControlExpr)
{
  GpuSampler* s = new GpuSampler();
  // ---- control.common (init.cpp:1015-1051) ----
  s->defaultWarmup = getInt(commonControlExpr, "warmup", 1000); s->defaultIter = getInt(commonControlExpr, "iter", 2000);
  s->verbose = getInt(commonControlExpr, "verbose", 0); s->refresh = getInt(commonControlExpr, "refresh", 200);
  s->responseIsBinary = getBool(commonControlExpr, "is_binary", false); s->keepFits = getBool(commonControlExpr, "keep_fits", true);
  s->callback = listElement(commonControlExpr, "callback"); s->callbackEnv = listElement(commonControlExpr, "callbackEnv");
  if (s->callback != R_NilValue && !Rf_isFunction(s->callback)) { delete s; Rf_error("callback must be a function or NULL"); }
  s->userOffset = getRealVec(commonControlExpr, "offset"); s->offsetType = getInt(commonControlExpr, "offset_type", 0);
  const double* bartOffsetInit = getRealVec(commonControlExpr, "bart_offset_init");
  const double sigmaInit = getReal(commonControlExpr, "sigma_init", 1.0);
  if (!(sigmaInit > 0.0)) { delete s; Rf_error("sigma_init must be greater than 0"); }

  // ---- data.stan (src/stan_sampler.cpp:67-80, :112-380) ----
  s4b_glmm_data gd; std::memset(&gd, 0, sizeof gd);
  gd.N = getInt(stanDataExpr, "N", 0); gd.K = getInt(stanDataExpr, "K", 0); gd.is_binary = getInt(stanDataExpr, "is_binary", 0);
  gd.prior_dist = getInt(stanDataExpr, "prior_dist", 1); gd.prior_dist_for_aux = getInt(stanDataExpr, "prior_dist_for_aux", 0);
  gd.t = getInt(stanDataExpr, "t", 0); gd.q = getInt(stanDataExpr, "q", 0); gd.len_theta_L = getInt(stanDataExpr, "len_theta_L", 0);
  gd.len_concentration = getInt(stanDataExpr, "len_concentration", 0); gd.len_regularization = getInt(stanDataExpr, "len_regularization", 0);
  gd.num_non_zero = getInt(stanDataExpr, "num_non_zero", 0);
  gd.X = getRealVec(stanDataExpr, "X"); gd.y = getRealVec(stanDataExpr, "y");
  gd.prior_scale = getRealVec(stanDataExpr, "prior_scale"); gd.prior_mean = getRealVec(stanDataExpr, "prior_mean");
  gd.prior_scale_for_aux = getReal(stanDataExpr, "prior_scale_for_aux", 0.0); gd.prior_mean_for_aux = getReal(stanDataExpr, "prior_mean_for_aux", 0.0);
  gd.prior_df_for_aux = getReal(stanDataExpr, "prior_df_for_aux", 1.0);
  gd.p = getIntVec(stanDataExpr, "p"); gd.l = getIntVec(stanDataExpr, "l"); gd.shape = getRealVec(stanDataExpr, "shape"); gd.scale = getRealVec(stanDataExpr, "scale");
  gd.concentration = getRealVec(stanDataExpr, "concentration"); gd.regularization = getRealVec(stanDataExpr, "regularization");
  gd.w = getRealVec(stanDataExpr, "w"); gd.v = getIntVec(stanDataExpr, "v"); gd.u = getIntVec(stanDataExpr, "u");
  if (getInt(stanDataExpr, "has_weights", 0) != 0) gd.weights = getRealVec(stanDataExpr, "weights");
  gd.prior_df = getRealVec(stanDataExpr, "prior_df"); gd.num_normals = getIntVec(stanDataExpr, "num_normals");
  gd.global_prior_df = getReal(stanDataExpr, "global_prior_df", 0.0); gd.global_prior_scale = getReal(stanDataExpr, "global_prior_scale", 0.0);
  gd.slab_df = getReal(stanDataExpr, "slab_df", 0.0); gd.slab_scale = getReal(stanDataExpr, "slab_scale", 0.0);
  if (getInt(stanDataExpr, "has_intercept", 0) != 0) { delete s; Rf_error("has_intercept: stan4bart's front end never sets it (R/rstanarm_functions.R:420-447)"); }

  // ---- control.stan (src/stan_sampler.cpp:82-96, :395-458; skip default init.cpp:206-209) ----
  s4b_stan_control sc; std::memset(&sc, 0, sizeof sc);
  sc.seed = (uint32_t) getInt(stanControlExpr, "seed", 0);
  sc.skip = getInt(stanControlExpr, "skip", NA_INTEGER);
  if (sc.skip == NA_INTEGER) { sc.skip = (2000 - s->defaultWarmup) / 1000; if (sc.skip < 1) sc.skip = 1; }
  sc.init_radius = getReal(stanControlExpr, "init_r", 2.0); sc.adapt_gamma = getReal(stanControlExpr, "adapt_gamma", 0.05);
  sc.adapt_delta = getReal(stanControlExpr, "adapt_delta", 0.8); sc.adapt_kappa = getReal(stanControlExpr, "adapt_kappa", 0.75);
  sc.adapt_t0 = getReal(stanControlExpr, "adapt_t0", 10.0);
  sc.adapt_init_buffer = (uint32_t) getInt(stanControlExpr, "adapt_init_buffer", 75); sc.adapt_term_buffer = (uint32_t) getInt(stanControlExpr, "adapt_term_buffer", 50);
  sc.adapt_window = (uint32_t) getInt(stanControlExpr, "adapt_window", 25); sc.max_treedepth = getInt(stanControlExpr, "max_treedepth", 10);
  sc.stepsize = getReal(stanControlExpr, "stepsize", 1.0); sc.stepsize_jitter = getReal(stanControlExpr, "stepsize_jitter", 0.0);

  // ---- dbartsControl / dbartsData / dbartsModel (init.cpp:215-225) ----
  GetRNGstate();                                        // an NA rngSeed is drawn from R's generator (init.cpp:259)
  gpubart_shim_initializeControl(&s->bartControl, bartControlExpr);
  PutRNGstate();
  s->keepTrees = s->bartControl.keepTrees;
  s->bartControl.responseIsBinary = s->responseIsBinary;
  gpubart_shim_initializeData(&s->bartData, bartDataExpr);
  gpubart_shim_initializeModel(&s->bartModel, bartModelExpr, &s->bartControl, &s->bartData);
  s->n = s->bartData.numObservations; s->nTest = s->bartData.numTestObservations; s->p = s->bartData.numPredictors;
  s->kIsModeled = !s->bartModel.kPrior->isFixed;
  s4b_bart_config bc; std::memset(&bc, 0, sizeof bc);
  bc.n = (int64_t) s->n; bc.p = (int64_t) s->p; bc.n_test = (int64_t) s->nTest; bc.num_trees = (int32_t) s->bartControl.numTrees;
  bc.thin = (int32_t) s->bartControl.treeThinningRate; bc.min_obs = (int32_t) s->bartControl.minNumObservationsInNode; bc.is_binary = s->responseIsBinary ? 1 : 0;
  bc.birth_death_prob = s->bartModel.birthOrDeathProbability; bc.swap_prob = s->bartModel.swapProbability; bc.change_prob = s->bartModel.changeProbability;
  bc.birth_prob = s->bartModel.birthProbability; bc.base = s->bartModel.treePrior->base; bc.power = s->bartModel.treePrior->power;
  bc.k = s->kIsModeled ? s->bartModel.kPrior->k : s->bartModel.muPrior->k; bc.node_scale = s->bartModel.nodeScale; bc.seed = s->bartControl.rngSeed;
  bc.split_probs = s->bartModel.treePrior->splitProbabilities; bc.weights = s->bartData.weights;
  if (s->kIsModeled) { bc.k_df = s->bartModel.kPrior->degreesOfFreedom; bc.k_scale = s->bartModel.kPrior->scale; }
  std::vector<int32_t> ncuts(s->p);
  bc.n_cuts = 1;
  for (std::size_t j = 0; j < s->p; ++j) { ncuts[j] = (int32_t) s->bartData.maxNumCuts[j]; if (ncuts[j] > bc.n_cuts) bc.n_cuts = ncuts[j]; }
  bc.n_cuts_var = ncuts.data();

  s4b_common_control cc; std::memset(&cc, 0, sizeof cc);
  cc.warmup = s->defaultWarmup; cc.iter = s->defaultIter; cc.is_binary = s->responseIsBinary ? 1 : 0; cc.keep_fits = s->keepFits ? 1 : 0;
  cc.sigma_init = sigmaInit; cc.offset_type = s->offsetType; cc.user_offset = s->userOffset;
  if (s4b_sampler_create(&bc, s->bartData.y, s->bartData.x, s->bartData.x_test, &gd, &sc, &cc, bartOffsetInit, &s->h) != 0) {
    std::string msg = s4b_last_error();
    gpubart_shim_invalidateModel(&s->bartModel); gpubart_shim_invalidateData(&s->bartData);
    delete s;
    Rf_error("%s", msg.c_str());
  }
  if (s->keepTrees) check(gpubart_set_keep_trees(s4b_sampler_bart(s->h), (int64_t) (s->defaultIter - s->defaultWarmup > 0 ? s->defaultIter - s->defaultWarmup : 1)));
  check(s4b_sampler_num_stan_pars(s->h, &s->numPars));
  {
    size_t need = 0;
    check(glmm_stan_row_names(s4b_sampler_glmm(s->h), NULL, 0, &need));
    std::vector<char> buf(need + 1);
    check(glmm_stan_row_names(s4b_sampler_glmm(s->h), buf.data(), buf.size(), &need));
    for (char* tok = std::strtok(buf.data(), "\n"); tok != NULL; tok = std::strtok(NULL, "\n")) s->rowNames.push_back(tok);
  }
  // the data live in R objects that must outlive the sampler: keep them reachable from the external pointer
  SEXP keep = PROTECT(Rf_allocVector(VECSXP, 6));
  SET_VECTOR_ELT(keep, 0, bartDataExpr); SET_VECTOR_ELT(keep, 1, bartModelExpr); SET_VECTOR_ELT(keep, 2, stanDataExpr);
  SET_VECTOR_ELT(keep, 3, commonControlExpr); SET_VECTOR_ELT(keep, 4, s->callback); SET_VECTOR_ELT(keep, 5, s->callbackEnv);
  SEXP result = PROTECT(R_MakeExternalPtr(s, R_NilValue, keep));
  R_RegisterCFinalizerEx(result, samplerFinalizer, FALSE);
  UNPROTECT(2);
  return result;
}

// stan4bart_run(sampler, numIter, isWarmup, resultsType) -> list(stan = [pars x S], bart = list(sigma, train, test, varcount[, k])[, callback])
// (init.cpp:678-965; result layout src/bart_util.cpp:13-81, src/stan_sampler.cpp:577-596)
static SEXP run(SEXP samplerExpr, SEXP numIterExpr, SEXP isWarmupExpr, SEXP resultsTypeExpr)
{
  GpuSampler* s = samplerOf(samplerExpr, "run");
  const int numIter = Rf_asInteger(numIterExpr);
  if (numIter == NA_INTEGER || numIter < 1) Rf_error("num_iter must be greater than or equal to 1");
  const bool isWarmup = Rf_asLogical(isWarmupExpr) == TRUE;
  int resultsType = 0;                                   // "both" | "bart" | "stan" (R passes the name; integers 0 / 1 / 2 also accepted)
  if (Rf_isString(resultsTypeExpr) && XLENGTH(resultsTypeExpr) > 0) {
    const char* rt = CHAR(STRING_ELT(resultsTypeExpr, 0));
    resultsType = std::strcmp(rt, "bart") == 0 ? 1 : (std::strcmp(rt, "stan") == 0 ? 2 : 0);
  } else if (!Rf_isNull(resultsTypeExpr)) { const int v = Rf_asInteger(resultsTypeExpr); if (v != NA_INTEGER) resultsType = v; }
  const bool wantBart = resultsType != 2, wantStan = resultsType != 1;
  const std::size_t S = s->keepFits ? (std::size_t) numIter : 1;
  int protectCount = 0;

  SEXP stanExpr = R_NilValue, sigmaExpr = R_NilValue, trainExpr = R_NilValue, testExpr = R_NilValue, varcountExpr = R_NilValue, kExpr = R_NilValue;
  if (wantStan) {
    stanExpr = PROTECT(Rf_allocMatrix(REALSXP, s->numPars, (int) S)); ++protectCount;
    SEXP dn = PROTECT(Rf_allocVector(VECSXP, 2)); ++protectCount;
    SEXP rn = PROTECT(Rf_allocVector(STRSXP, s->numPars)); ++protectCount;
    for (int i = 0; i < s->numPars; ++i) SET_STRING_ELT(rn, i, Rf_mkChar(s->rowNames[(std::size_t) i].c_str()));
    SET_VECTOR_ELT(dn, 0, rn);
    Rf_setAttrib(stanExpr, R_DimNamesSymbol, dn);
  }
  if (wantBart) {
    sigmaExpr = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t) S)); ++protectCount;
    trainExpr = PROTECT(Rf_allocMatrix(REALSXP, (int) s->n, (int) S)); ++protectCount;
    if (s->nTest > 0) { testExpr = PROTECT(Rf_allocMatrix(REALSXP, (int) s->nTest, (int) S)); ++protectCount; }
    varcountExpr = PROTECT(Rf_allocMatrix(INTSXP, (int) s->p, (int) S)); ++protectCount;
    if (s->kIsModeled) { kExpr = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t) S)); ++protectCount; }
  }
  CallbackState cb;
  cb.closure = R_NilValue; cb.results = R_NilValue; cb.n = s->n; cb.nTest = s->nTest; cb.numPars = s->numPars; cb.numIter = numIter; cb.resultLength = 0;
  cb.yhatTrain = cb.yhatTest = cb.stanPars = R_NilValue;
  if (s->callback != R_NilValue) {
    cb.yhatTrain = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t) s->n)); ++protectCount;
    cb.yhatTest = s->nTest > 0 ? Rf_allocVector(REALSXP, (R_xlen_t) s->nTest) : R_NilValue; PROTECT(cb.yhatTest); ++protectCount;
    cb.stanPars = PROTECT(Rf_allocVector(REALSXP, s->numPars)); ++protectCount;
    SEXP nm = PROTECT(Rf_allocVector(STRSXP, s->numPars)); ++protectCount;
    for (int i = 0; i < s->numPars; ++i) SET_STRING_ELT(nm, i, Rf_mkChar(s->rowNames[(std::size_t) i].c_str()));
    Rf_setAttrib(cb.stanPars, R_NamesSymbol, nm);
    cb.closure = PROTECT(Rf_lang4(s->callback, cb.yhatTrain, cb.yhatTest, cb.stanPars)); ++protectCount;
  }
  check(s4b_sampler_set_callback(s->h, iterationCallback, &cb));
  // keepTrees only for the sampling runs (init.cpp:737-744): s4b_sampler_run switches the store by `is_warmup`
  if (s->verbose > 0) Rprintf("starting %s, %d draws, %s\n", isWarmup ? "warmup" : "sampling", numIter, resultsType == 0 ? "both BART and Stan" : (resultsType == 1 ? "BART only" : "Stan only"));
  GetRNGstate();
  std::vector<uint32_t> vc(wantBart ? s->p * S : 1);
  const int rc = s4b_sampler_run(s->h, numIter, isWarmup ? 1 : 0, wantStan ? REAL(stanExpr) : NULL, wantBart ? REAL(trainExpr) : NULL,
                                 (wantBart && s->nTest > 0) ? REAL(testExpr) : NULL, wantBart ? vc.data() : NULL, wantBart ? REAL(sigmaExpr) : NULL);
  PutRNGstate();
  s4b_sampler_set_callback(s->h, NULL, NULL);
  if (rc != 0) { if (cb.results != R_NilValue) R_ReleaseObject(cb.results); UNPROTECT(protectCount); Rf_error("%s", s4b_last_error()); }
  if (wantBart) {
    for (std::size_t i = 0; i < s->p * S; ++i) INTEGER(varcountExpr)[i] = (int) vc[i];
    if (s->kIsModeled) {
      std::vector<double> k((std::size_t) numIter); int cnt = 0;
      check(s4b_sampler_last_k(s->h, k.data(), numIter, &cnt));
      for (std::size_t i = 0; i < S; ++i) REAL(kExpr)[i] = s->keepFits ? k[i] : k[(std::size_t) (cnt > 0 ? cnt - 1 : 0)];
    }
  }
  // ---- list(stan =, bart =[, callback =]) with keep_fits; list(callback =) without (init.cpp:920-960) ----
  SEXP result;
  if (s->keepFits) {
    const char* names[3]; int len = 0;
    if (wantStan) names[len++] = "stan";
    if (wantBart) names[len++] = "bart";
    if (s->callback != R_NilValue) names[len++] = "callback";
    result = PROTECT(namedList(len, names)); ++protectCount;
    int pos = 0;
    if (wantStan) SET_VECTOR_ELT(result, pos++, stanExpr);
    if (wantBart) {
      const char* bn[5] = { "sigma", "train", "test", "varcount", "k" };
      SEXP bart = PROTECT(namedList(s->kIsModeled ? 5 : 4, bn)); ++protectCount;
      SET_VECTOR_ELT(bart, 0, sigmaExpr); SET_VECTOR_ELT(bart, 1, trainExpr); SET_VECTOR_ELT(bart, 2, testExpr); SET_VECTOR_ELT(bart, 3, varcountExpr);
      if (s->kIsModeled) SET_VECTOR_ELT(bart, 4, kExpr);
      SET_VECTOR_ELT(result, pos++, bart);
    }
    if (s->callback != R_NilValue) SET_VECTOR_ELT(result, pos, cb.results);
  } else {
    const char* names[1] = { "callback" };
    result = PROTECT(namedList(1, names)); ++protectCount;
    SET_VECTOR_ELT(result, 0, cb.results);
  }
  if (cb.results != R_NilValue) R_ReleaseObject(cb.results);
  UNPROTECT(protectCount);
  return result;
}

static SEXP printInitialSummary(SEXP samplerExpr)
{
  GpuSampler* s = samplerOf(samplerExpr, "printInitialSummary");
  size_t need = 0;
  check(gpubart_summary(s4b_sampler_bart(s->h), NULL, 0, &need));
  std::vector<char> buf(need + 1);
  check(gpubart_summary(s4b_sampler_bart(s->h), buf.data(), buf.size(), &need));
  Rprintf("bart init:\n%s", buf.data());
  if (s->userOffset != NULL) {
    Rprintf("\nuser offset: %f", s->userOffset[0]);
    for (std::size_t i = 1; i < (s->n < 5 ? s->n : 5); ++i) Rprintf(", %f", s->userOffset[i]);
    if (s->n > 5) Rprintf("...");
    Rprintf("\n");
  }
  return R_NilValue;
}

static SEXP disengageAdaptation(SEXP samplerExpr) { check(s4b_sampler_disengage_adaptation(samplerOf(samplerExpr, "disengageAdaptation")->h)); return R_NilValue; }

static SEXP finalize(void) { return R_NilValue; }     // every sampler is released by its own finalizer

// stan4bart_exportBARTState(sampler) -> list(<raw vector>): the stored draws of this chain (init.cpp:409-416)
static SEXP exportBARTState(SEXP samplerExpr)
{
  GpuSampler* s = samplerOf(samplerExpr, "exportBARTState");
  int64_t bytes = 0;
  check(gpubart_stored_export_size(s4b_sampler_bart(s->h), &bytes));
  SEXP out = PROTECT(Rf_allocVector(VECSXP, 1));
  SEXP raw = PROTECT(Rf_allocVector(RAWSXP, (R_xlen_t) bytes));
  check(gpubart_stored_export(s4b_sampler_bart(s->h), RAW(raw), bytes));
  SET_VECTOR_ELT(out, 0, raw);
  UNPROTECT(2);
  return out;
}

// stan4bart_createStoredBARTSampler(control, data, model, state) (init.cpp:418-446); `state` = the list of one chain's export
static SEXP createStoredBARTSampler(SEXP controlExpr, SEXP dataExpr, SEXP /* modelExpr */, SEXP stateExpr)
{
  if (XLENGTH(stateExpr) != 1) Rf_error("one chain per stored BART sampler (R/stan4bart_fit.R:572-580 creates one per chain)");
  SEXP raw = VECTOR_ELT(stateExpr, 0);
  StoredSampler* s = new StoredSampler();
  s->binary = Rf_asLogical(R_do_slot(controlExpr, Rf_install("binary"))) == TRUE;
  s->numPredictors = (std::size_t) INTEGER(Rf_getAttrib(R_do_slot(dataExpr, Rf_install("x")), R_DimSymbol))[1];
  if (gpubart_stored_import(RAW(raw), (int64_t) XLENGTH(raw), &s->st) != 0) { delete s; Rf_error("%s", s4b_last_error()); }
  SEXP result = PROTECT(R_MakeExternalPtr(s, R_NilValue, R_NilValue));
  R_RegisterCFinalizerEx(result, storedFinalizer, FALSE);
  UNPROTECT(1);
  return result;
}

// stan4bart_predictBART(storedSampler, x_test, offset_test) -> [n x numSamples] in the identity scale (init.cpp:354-403, :433-435;
// R un-scales with range.bart, R/generics.R:671-674)
static SEXP predictBART(SEXP storedExpr, SEXP xTestExpr, SEXP offsetTestExpr)
{
  StoredSampler* s = static_cast<StoredSampler*>(R_ExternalPtrAddr(storedExpr));
  if (s == NULL) Rf_error("predictBART called on NULL external pointer");
  if (Rf_isNull(xTestExpr)) return R_NilValue;
  if (!Rf_isReal(xTestExpr)) Rf_error("x.test must be of type real");
  int* dims = INTEGER(Rf_getAttrib(xTestExpr, R_DimSymbol));
  if ((std::size_t) dims[1] != s->numPredictors) Rf_error("dimensions of x_test must match the number of predictors");
  const int64_t n = dims[0];
  const double* off = NULL;
  if (!Rf_isNull(offsetTestExpr) && Rf_isReal(offsetTestExpr) && !(XLENGTH(offsetTestExpr) == 1 && ISNA(REAL(offsetTestExpr)[0]))) {
    if (XLENGTH(offsetTestExpr) != n) Rf_error("length of offset.test must equal number of rows in x.test");
    off = REAL(offsetTestExpr);
  }
  int64_t S = 0;
  check(gpubart_stored_count(s->st, &S));
  SEXP result = PROTECT(Rf_allocMatrix(REALSXP, (int) n, (int) S));
  check(gpubart_stored_predict(s->st, REAL(xTestExpr), n, NULL, 0, S, REAL(result)));
  std::vector<double> scales((std::size_t) (2 * (S > 0 ? S : 1)));
  check(gpubart_stored_get_scales(s->st, 0, S, scales.data()));
  for (int64_t k = 0; k < S; ++k) {
    double* col = REAL(result) + (std::size_t) k * (std::size_t) n;
    if (!s->binary) for (int64_t i = 0; i < n; ++i) col[i] = (col[i] - scales[(std::size_t) (2 * k)]) / scales[(std::size_t) (2 * k + 1)] - 0.5;
    if (off != NULL) for (int64_t i = 0; i < n; ++i) col[i] += off[i];
  }
  UNPROTECT(1);
  return result;
}

static SEXP getParametricMean(SEXP samplerExpr)
{
  GpuSampler* s = samplerOf(samplerExpr, "getParametricMean");
  SEXP result = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t) s->n));
  check(s4b_sampler_get_parametric_mean(s->h, REAL(result)));
  UNPROTECT(1);
  return result;
}

static SEXP getBARTDataRange(SEXP samplerExpr)
{
  GpuSampler* s = samplerOf(samplerExpr, "getBARTDataRange");
  SEXP result = PROTECT(Rf_allocVector(REALSXP, 2));
  check(s4b_sampler_get_bart_data_range(s->h, REAL(result)));
  UNPROTECT(1);
  return result;
}

// stan4bart_getTrees(storedSampler | sampler, chainIndices, sampleIndices, treeIndices, current) -> data.frame(sample, tree, n, var, value)
// (init.cpp:514-671).  On the device the trees of the draws live with the sampler that produced them, so this takes the sampler.
static SEXP getTrees(SEXP samplerExpr, SEXP /* chainIndicesExpr */, SEXP sampleIndicesExpr, SEXP treeIndicesExpr, SEXP currentExpr)
{
  GpuSampler* s = samplerOf(samplerExpr, "getTrees");
  gpubart_fit* g = s4b_sampler_bart(s->h);
  const bool live = Rf_asLogical(currentExpr) == TRUE || !s->keepTrees;
  int64_t stored = 0;
  check(gpubart_num_stored(g, &stored));
  std::vector<int64_t> samples;
  if (!live) {
    if (Rf_isNull(sampleIndicesExpr)) for (int64_t i = 0; i < stored; ++i) samples.push_back(i);
    else for (R_xlen_t i = 0; i < XLENGTH(sampleIndicesExpr); ++i) samples.push_back((int64_t) INTEGER(sampleIndicesExpr)[i] - 1);
  } else samples.push_back(-1);
  std::vector<char> want(s->bartControl.numTrees, Rf_isNull(treeIndicesExpr) ? 1 : 0);
  if (!Rf_isNull(treeIndicesExpr)) for (R_xlen_t i = 0; i < XLENGTH(treeIndicesExpr); ++i) {
    const int t = INTEGER(treeIndicesExpr)[i] - 1;
    if (t < 0 || (std::size_t) t >= want.size()) Rf_error("tree index out of range");
    want[(std::size_t) t] = 1;
  }
  std::vector<int> sampleNo, treeNo, nObs, var; std::vector<double> value;
  for (std::size_t r = 0; r < samples.size(); ++r) {
    int64_t k = 0;
    if (live) check(gpubart_num_nodes(g, &k)); else check(gpubart_num_stored_nodes(g, samples[r], &k));
    std::vector<int32_t> tno((std::size_t) k), v((std::size_t) k); std::vector<int64_t> nn((std::size_t) k); std::vector<double> val((std::size_t) k);
    if (live) check(gpubart_get_trees(g, tno.data(), nn.data(), v.data(), val.data()));
    else check(gpubart_get_stored_trees(g, samples[r], tno.data(), nn.data(), v.data(), val.data()));
    for (int64_t i = 0; i < k; ++i) if (want[(std::size_t) tno[(std::size_t) i]]) {
      sampleNo.push_back((int) samples[r] + 1); treeNo.push_back(tno[(std::size_t) i] + 1); nObs.push_back((int) nn[(std::size_t) i]);
      var.push_back(v[(std::size_t) i] >= 0 ? v[(std::size_t) i] + 1 : v[(std::size_t) i]); value.push_back(val[(std::size_t) i]);
    }
  }
  const R_xlen_t m = (R_xlen_t) treeNo.size();
  const char* colsSaved[5] = { "sample", "tree", "n", "var", "value" };
  const char* colsLive[4] = { "tree", "n", "var", "value" };
  SEXP df = PROTECT(live ? namedList(4, colsLive) : namedList(5, colsSaved));
  int col = 0;
  if (!live) { SEXP c = PROTECT(Rf_allocVector(INTSXP, m)); for (R_xlen_t i = 0; i < m; ++i) INTEGER(c)[i] = sampleNo[(std::size_t) i]; SET_VECTOR_ELT(df, col++, c); UNPROTECT(1); }
  { SEXP c = PROTECT(Rf_allocVector(INTSXP, m)); for (R_xlen_t i = 0; i < m; ++i) INTEGER(c)[i] = treeNo[(std::size_t) i]; SET_VECTOR_ELT(df, col++, c); UNPROTECT(1); }
  { SEXP c = PROTECT(Rf_allocVector(INTSXP, m)); for (R_xlen_t i = 0; i < m; ++i) INTEGER(c)[i] = nObs[(std::size_t) i]; SET_VECTOR_ELT(df, col++, c); UNPROTECT(1); }
  { SEXP c = PROTECT(Rf_allocVector(INTSXP, m)); for (R_xlen_t i = 0; i < m; ++i) INTEGER(c)[i] = var[(std::size_t) i]; SET_VECTOR_ELT(df, col++, c); UNPROTECT(1); }
  { SEXP c = PROTECT(Rf_allocVector(REALSXP, m)); for (R_xlen_t i = 0; i < m; ++i) REAL(c)[i] = value[(std::size_t) i]; SET_VECTOR_ELT(df, col++, c); UNPROTECT(1); }
  SEXP cls = PROTECT(Rf_mkString("data.frame"));
  Rf_setAttrib(df, R_ClassSymbol, cls);
  SEXP rn = PROTECT(Rf_allocVector(INTSXP, 2));          // compact row names c(NA, -m)
  INTEGER(rn)[0] = NA_INTEGER; INTEGER(rn)[1] = -(int) m;
  Rf_setAttrib(df, R_RowNamesSymbol, rn);
  UNPROTECT(3);
  return df;
}

static SEXP printTrees(SEXP samplerExpr, SEXP chainIndicesExpr, SEXP sampleIndicesExpr, SEXP treeIndicesExpr)
{
  SEXP cur = PROTECT(Rf_ScalarLogical(FALSE));
  SEXP df = PROTECT(getTrees(samplerExpr, chainIndicesExpr, sampleIndicesExpr, treeIndicesExpr, cur));
  const int ncol = (int) XLENGTH(df);
  const int* tree = INTEGER(VECTOR_ELT(df, ncol - 4)); const int* n = INTEGER(VECTOR_ELT(df, ncol - 3)); const int* var = INTEGER(VECTOR_ELT(df, ncol - 2));
  const double* value = REAL(VECTOR_ELT(df, ncol - 1));
  std::vector<int> open;
  for (R_xlen_t i = 0; i < XLENGTH(VECTOR_ELT(df, 0)); ++i) {
    if (open.empty()) Rprintf("tree %d\n", tree[i]);
    for (std::size_t d = 0; d <= open.size(); ++d) Rprintf("  ");
    if (var[i] >= 0) { Rprintf("x%d <= %.6g (n = %d)\n", var[i], value[i], n[i]); open.push_back(2); }
    else { Rprintf("mu = %.6g (n = %d)\n", value[i], n[i]); while (!open.empty() && --open.back() == 0) open.pop_back(); }
  }
  UNPROTECT(2);
  return R_NilValue;
}

#define DEF_FUNC(_N_, _F_, _A_) { _N_, (DL_FUNC) (void (*)(void)) &_F_, _A_ }
static R_CallMethodDef R_callMethods[] = {
  DEF_FUNC("stan4bart_create", createSampler, 6),
  DEF_FUNC("stan4bart_run", run, 4),
  DEF_FUNC("stan4bart_printInitialSummary", printInitialSummary, 1),
  DEF_FUNC("stan4bart_disengageAdaptation", disengageAdaptation, 1),
  DEF_FUNC("stan4bart_finalize", finalize, 0),
  DEF_FUNC("stan4bart_exportBARTState", exportBARTState, 1),
  DEF_FUNC("stan4bart_createStoredBARTSampler", createStoredBARTSampler, 4),
  DEF_FUNC("stan4bart_predictBART", predictBART, 3),
  DEF_FUNC("stan4bart_getParametricMean", getParametricMean, 1),
  DEF_FUNC("stan4bart_getBARTDataRange", getBARTDataRange, 1),
  DEF_FUNC("stan4bart_printTrees", printTrees, 4),
  DEF_FUNC("stan4bart_getTrees", getTrees, 5),
  { NULL, NULL, 0 }
};
#undef DEF_FUNC

void R_init_stan4bart(DllInfo* info)
{
  R_registerRoutines(info, NULL, R_callMethods, NULL, NULL);
  R_useDynamicSymbols(info, FALSE);
}

}  // extern "C"
