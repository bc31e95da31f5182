// tests/shim_driver.cpp -- TEST INFRASTRUCTURE.  Drives the GPU sampler the way stan4bart's host code drives dbarts: only through a
// BARTFunctionTable filled by name (the analogue of lookupBARTFunctions, /root/reference/src/init.cpp:1113-1147), with the dbarts
// types of include/dbarts_shim, in the call order of createSampler (init.cpp:215-298) and of the run loop (:737-744, :796-847,
// :914-915), then predict / getTrees / printTrees on the stored draws as predictBART / getTrees do (:354-403, :514-671).
// Reads a problem from a binary file, writes everything it saw to another; tests/test_shim_gpu.py compares that with the
// direct C ABI run on the same seeds.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <vector>

#include "gpubart_shim.h"

namespace {

struct BARTFunctionTable {          // same members, same signatures as init.cpp:54-81
  void (*initializeFit)(dbarts::BARTFit* fit, dbarts::Control* control, dbarts::Model* model, dbarts::Data* data);
  void (*invalidateFit)(dbarts::BARTFit* fit);
  void (*initializeControl)(dbarts::Control* control, SEXP controlExpr);
  void (*initializeData)(dbarts::Data* data, SEXP dataExpr);
  void (*invalidateData)(dbarts::Data* data);
  void (*initializeModel)(dbarts::Model* model, SEXP modelExpr, const dbarts::Control* control, const dbarts::Data* data);
  void (*invalidateModel)(dbarts::Model* model);
  SEXP (*createStateExpression)(const dbarts::BARTFit* fit);
  void (*initializeState)(dbarts::BARTFit* fit, SEXP stateExpr);
  void (*setControl)(dbarts::BARTFit* fit, const dbarts::Control* control);
  void (*runSamplerWithResults)(dbarts::BARTFit* fit, std::size_t numBurnIn, dbarts::Results* results);
  void (*predict)(const dbarts::BARTFit* fit, const double* x_test, std::size_t numTestObservations, const double* testOffset, double* result);
  void (*setResponse)(dbarts::BARTFit* fit, const double* response);
  void (*setOffset)(dbarts::BARTFit* fit, const double* offset, bool updateState);
  void (*setSigma)(dbarts::BARTFit* fit, const double* sigma);
  void (*sampleTreesFromPrior)(dbarts::BARTFit* fit);
  void (*printInitialSummary)(const dbarts::BARTFit* fit);
  void (*getLatentVariables)(const dbarts::BARTFit*, double*);
  void (*printTrees)(const dbarts::BARTFit*, const std::size_t*, std::size_t, const std::size_t*, std::size_t, const std::size_t*, std::size_t);
  dbarts::FlattenedTrees* (*getTrees)(const dbarts::BARTFit*, const std::size_t*, std::size_t, const std::size_t*, std::size_t, const std::size_t*,
                                      std::size_t, bool);
};
BARTFunctionTable bartFunctions;

template <class F> void bind(F& slot, const char* name)
{
  void (*p)(void) = gpubart_shim_lookup(name);
  if (p == NULL) throw std::runtime_error(std::string("missing table entry ") + name);
  slot = reinterpret_cast<F>(p);
}
void lookupBARTFunctions()
{
  bind(bartFunctions.initializeFit, "initializeFit"); bind(bartFunctions.invalidateFit, "invalidateFit");
  bind(bartFunctions.initializeControl, "initializeControl"); bind(bartFunctions.initializeData, "initializeData");
  bind(bartFunctions.invalidateData, "invalidateData"); bind(bartFunctions.initializeModel, "initializeModel");
  bind(bartFunctions.invalidateModel, "invalidateModel"); bind(bartFunctions.setControl, "setControl");
  bind(bartFunctions.predict, "predict"); bind(bartFunctions.setResponse, "setResponse"); bind(bartFunctions.setOffset, "setOffset");
  bind(bartFunctions.setSigma, "setSigma"); bind(bartFunctions.createStateExpression, "createStateExpression");
  bind(bartFunctions.initializeState, "initializeState"); bind(bartFunctions.runSamplerWithResults, "runSamplerWithResults");
  bind(bartFunctions.sampleTreesFromPrior, "sampleTreesFromPrior"); bind(bartFunctions.printInitialSummary, "printInitialSummary");
  bind(bartFunctions.getLatentVariables, "storeLatents"); bind(bartFunctions.printTrees, "printTrees"); bind(bartFunctions.getTrees, "getTrees");
}

// a results window that slides over a numSamples-deep buffer, one draw per sampler call (what the reference's IterableBartResults does)
struct SlidingResults : dbarts::Results {
  double *sigma0, *train0, *test0, *k0; std::uint32_t* vc0; std::size_t depth;
  SlidingResults(std::size_t n, std::size_t p, std::size_t nt, std::size_t S, bool kModeled) : dbarts::Results(n, p, nt, S, 1, kModeled),
    sigma0(sigmaSamples), train0(trainingSamples), test0(testSamples), k0(kSamples), vc0(variableCountSamples), depth(S) { numSamples = 1; }
  void advance() { sigmaSamples += 1; trainingSamples += numObservations; if (testSamples) testSamples += numTestObservations; variableCountSamples += numPredictors; if (kSamples) kSamples += 1; }
  ~SlidingResults() { sigmaSamples = sigma0; trainingSamples = train0; testSamples = test0; kSamples = k0; variableCountSamples = vc0; numSamples = depth; }
};

std::vector<double> read_vec(FILE* f, std::size_t n) { std::vector<double> v(n); if (n && fread(v.data(), sizeof(double), n, f) != n) throw std::runtime_error("short read"); return v; }
void write_vec(FILE* f, const char* tag, const double* v, std::size_t n)
{
  char name[16]; std::memset(name, 0, sizeof name); std::strncpy(name, tag, 15);
  std::uint64_t len = n;
  fwrite(name, 1, 16, f); fwrite(&len, sizeof len, 1, f); if (n) fwrite(v, sizeof(double), n, f);
}

}  // namespace

int main(int argc, char** argv)
{
  if (argc != 3) { std::fprintf(stderr, "usage: shim_driver problem.bin out.bin\n"); return 2; }
  try {
    FILE* in = std::fopen(argv[1], "rb");
    if (!in) throw std::runtime_error("cannot open the problem file");
    std::int64_t hdr[10];
    if (fread(hdr, sizeof(std::int64_t), 10, in) != 10) throw std::runtime_error("short header");
    const std::size_t n = (std::size_t) hdr[0], p = (std::size_t) hdr[1], nt = (std::size_t) hdr[2];
    const int trees = (int) hdr[3], binary = (int) hdr[4], warm = (int) hdr[5], iters = (int) hdr[6], kmod = (int) hdr[8];
    const std::uint64_t seed = (std::uint64_t) hdr[7];
    std::vector<double> y = read_vec(in, n), x = read_vec(in, n * p), xt = read_vec(in, nt * p), off0 = read_vec(in, n), off1 = read_vec(in, n), y2 = read_vec(in, n);
    std::fclose(in);
    FILE* out = std::fopen(argv[2], "wb");
    if (!out) throw std::runtime_error("cannot open the output file");

    lookupBARTFunctions();
    // ---- createSampler (init.cpp:215-298) ----
    gpubart_control_expr ce; std::memset(&ce, 0, sizeof ce);
    ce.binary = binary; ce.keep_training_fits = 1; ce.keep_trees = 1; ce.n_samples = 1; ce.n_trees = trees; ce.n_chains = 1; ce.n_threads = 1; ce.n_thin = 1; ce.rng_seed = seed;
    std::int32_t ncuts = 100;
    gpubart_data_expr de; std::memset(&de, 0, sizeof de);
    de.y = y.data(); de.x = x.data(); de.x_test = nt ? xt.data() : NULL; de.n = (std::int64_t) n; de.p = (std::int64_t) p; de.n_test = (std::int64_t) nt; de.n_cuts = &ncuts; de.n_cuts_len = 1; de.sigma = 1.0;
    gpubart_model_expr me; std::memset(&me, 0, sizeof me);
    me.p_birth_death = 0.5; me.p_swap = 0.1; me.p_change = 0.4; me.p_birth = 0.5; me.base = 0.95; me.power = 2.0; me.k = 2.0;
    if (kmod) { me.k_df = 1.25; me.k_scale = 0.0; }

    dbarts::Control bartControl; dbarts::Data bartData; dbarts::Model bartModel(false);
    bartFunctions.initializeControl(&bartControl, reinterpret_cast<SEXP>(&ce));
    const bool keepTrees = bartControl.keepTrees;
    bartControl.keepTrees = false;
    if (keepTrees) { bartControl.defaultNumSamples = (std::size_t) iters; bartControl.defaultNumBurnIn = (std::size_t) warm; }
    bartControl.responseIsBinary = binary != 0;
    bartFunctions.initializeData(&bartData, reinterpret_cast<SEXP>(&de));
    bartFunctions.initializeModel(&bartModel, reinterpret_cast<SEXP>(&me), &bartControl, &bartData);
    dbarts::BARTFit* fit = static_cast<dbarts::BARTFit*>(::operator new(sizeof(dbarts::BARTFit)));
    bartFunctions.initializeFit(fit, &bartControl, &bartModel, &bartData);
    bartFunctions.setOffset(fit, off0.data(), true);
    const double sigma_init = 1.3;
    if (!binary) bartFunctions.setSigma(fit, &sigma_init);
    bartFunctions.sampleTreesFromPrior(fit);
    dbarts::Control quiet = fit->control; quiet.verbose = false;
    bartFunctions.setControl(fit, &quiet);
    {
      dbarts::Results firstDraw(n, fit->data.numPredictors, fit->data.numTestObservations, 1, quiet.numChains, !fit->model.kPrior->isFixed);
      bartFunctions.runSamplerWithResults(fit, 0, &firstDraw);
      write_vec(out, "first_train", firstDraw.trainingSamples, n);
      if (nt) write_vec(out, "first_test", firstDraw.testSamples, nt);
    }
    std::vector<double> latents(n, 0.0);
    if (binary) { bartFunctions.getLatentVariables(fit, latents.data()); write_vec(out, "first_latents", latents.data(), n); }
    // ---- run(warmup) then run(sampling) (init.cpp:728-744, :796-847, :914-915) ----
    for (int phase = 0; phase < 2; ++phase) {
      const bool isWarmup = phase == 0;
      const int numIter = isWarmup ? warm : iters;
      SlidingResults bartSamples(n, p, nt, (std::size_t) numIter, !bartModel.kPrior->isFixed);
      if (keepTrees) { bartControl.keepTrees = !isWarmup; bartFunctions.setControl(fit, &bartControl); }
      for (int iter = 0; iter < numIter; ++iter) {
        const double* off = (iter % 2 == 0) ? off1.data() : off0.data();
        const double sg = 1.0 + 0.05 * iter;
        if (!binary) bartFunctions.setSigma(fit, &sg);
        bartFunctions.setOffset(fit, off, isWarmup && iter % (1 << (8 * iter / numIter)) == 0);
        bartFunctions.runSamplerWithResults(fit, 0, &bartSamples);
        if (binary) bartFunctions.getLatentVariables(fit, latents.data());
        bartSamples.advance();
      }
      write_vec(out, isWarmup ? "warm_train" : "samp_train", bartSamples.train0, n * (std::size_t) numIter);
      if (nt) write_vec(out, isWarmup ? "warm_test" : "samp_test", bartSamples.test0, nt * (std::size_t) numIter);
      write_vec(out, isWarmup ? "warm_sigma" : "samp_sigma", bartSamples.sigma0, (std::size_t) numIter);
      std::vector<double> vc(p * (std::size_t) numIter);
      for (std::size_t i = 0; i < vc.size(); ++i) vc[i] = (double) bartSamples.vc0[i];
      write_vec(out, isWarmup ? "warm_varcount" : "samp_varcount", vc.data(), vc.size());
      if (bartSamples.k0) write_vec(out, isWarmup ? "warm_k" : "samp_k", bartSamples.k0, (std::size_t) numIter);
      if (binary) write_vec(out, isWarmup ? "warm_latents" : "samp_latents", latents.data(), n);
    }
    const double range[3] = { fit->sharedScratch.dataScale.min, fit->sharedScratch.dataScale.max, fit->sharedScratch.dataScale.range };
    write_vec(out, "data_range", range, 3);
    const double stored = (double) fit->currentNumSamples;
    write_vec(out, "num_stored", &stored, 1);
    // ---- predict on the stored draws of the live fit (original scale), then through an exported / re-imported stored sampler with
    // the identity scale that createStoredBARTSampler sets (init.cpp:418-446) ----
    std::vector<double> pred(n * fit->currentNumSamples);
    bartFunctions.predict(fit, x.data(), n, NULL, pred.data());
    write_vec(out, "pred_live", pred.data(), pred.size());
    SEXP state = bartFunctions.createStateExpression(fit);
    {
      dbarts::Control sc; dbarts::Data sdata; dbarts::Model sm(false);
      bartFunctions.initializeControl(&sc, reinterpret_cast<SEXP>(&ce));
      sc.numChains = 1; sc.keepTrees = true; sc.responseIsBinary = binary != 0;
      gpubart_data_expr de0 = de; de0.y = NULL; de0.x = NULL; de0.x_test = NULL;       // a stored sampler has no training data
      sdata.numObservations = n; sdata.numPredictors = p;
      (void) de0;
      bartFunctions.initializeModel(&sm, reinterpret_cast<SEXP>(&me), &sc, &sdata);
      dbarts::BARTFit* sfit = static_cast<dbarts::BARTFit*>(::operator new(sizeof(dbarts::BARTFit)));
      bartFunctions.initializeFit(sfit, &sc, &sm, &sdata);
      bartFunctions.initializeState(sfit, state);
      sfit->sharedScratch.dataScale.min = -0.5; sfit->sharedScratch.dataScale.max = 0.5; sfit->sharedScratch.dataScale.range = 1.0;
      std::vector<double> sp(n * sfit->currentNumSamples);
      bartFunctions.predict(sfit, x.data(), n, NULL, sp.data());
      write_vec(out, "pred_stored", sp.data(), sp.size());
      bartFunctions.invalidateFit(sfit); ::operator delete(sfit);
      bartFunctions.invalidateModel(&sm); bartFunctions.invalidateData(&sdata);
    }
    std::free(state);
    // ---- getTrees: stored sample 0 all trees; live trees 0 and 2 ----
    {
      std::size_t chain0 = 0, sample0 = 0, pick[2] = { 0, 2 };
      dbarts::FlattenedTrees* ft = bartFunctions.getTrees(fit, &chain0, 1, &sample0, 1, NULL, 0, false);
      std::vector<double> flat(4 * ft->totalNumNodes);
      for (std::size_t i = 0; i < ft->totalNumNodes; ++i) { flat[4 * i] = (double) ft->treeNumber[i]; flat[4 * i + 1] = (double) ft->numObservations[i]; flat[4 * i + 2] = (double) ft->variable[i]; flat[4 * i + 3] = ft->value[i]; }
      write_vec(out, "trees_stored0", flat.data(), flat.size());
      delete [] ft->value; delete [] ft->variable; delete [] ft->numObservations; delete [] ft->treeNumber; delete [] ft->sampleNumber; delete [] ft->chainNumber;
      ::operator delete(ft);
      ft = bartFunctions.getTrees(fit, &chain0, 1, NULL, 0, pick, 2, true);
      flat.assign(4 * ft->totalNumNodes, 0.0);
      for (std::size_t i = 0; i < ft->totalNumNodes; ++i) { flat[4 * i] = (double) ft->treeNumber[i]; flat[4 * i + 1] = (double) ft->numObservations[i]; flat[4 * i + 2] = (double) ft->variable[i]; flat[4 * i + 3] = ft->value[i]; }
      write_vec(out, "trees_live02", flat.data(), flat.size());
      delete [] ft->value; delete [] ft->variable; delete [] ft->numObservations; delete [] ft->treeNumber; delete [] ft->sampleNumber; delete [] ft->chainNumber;
      ::operator delete(ft);
      std::size_t one_tree = 1;
      bartFunctions.printTrees(fit, &chain0, 1, &sample0, 1, &one_tree, 1);
      bartFunctions.printInitialSummary(fit);
    }
    // ---- setResponse: a new response, then one more draw ----
    {
      bartControl.keepTrees = false; bartFunctions.setControl(fit, &bartControl);
      bartFunctions.setResponse(fit, y2.data());
      dbarts::Results r(n, p, nt, 1, 1, !bartModel.kPrior->isFixed);
      bartFunctions.runSamplerWithResults(fit, 0, &r);
      write_vec(out, "after_setresp", r.trainingSamples, n);
    }
    bartFunctions.invalidateFit(fit); ::operator delete(fit);
    bartFunctions.invalidateModel(&bartModel); bartFunctions.invalidateData(&bartData);
    std::fclose(out);
    return 0;
  } catch (const std::exception& e) {
    std::fprintf(stderr, "shim_driver: %s\n", e.what());
    return 1;
  }
}
