// include/dbarts_shim/dbarts/results.hpp -- shim of <dbarts/results.hpp> (see control.hpp).  Layout as the reference uses it
// (/root/reference/src/bart_util.hpp:14-64, src/bart_util.cpp:18-64, src/init.cpp:268-281): one buffer per quantity,
// samples contiguous, `numSamples` deep; the caller owns the object, the buffers belong to it.
#ifndef DBARTS_RESULTS_HPP
#define DBARTS_RESULTS_HPP

#include <cstddef>
#include <cstdint>

namespace dbarts {
  struct Results {
    double* sigmaSamples;                     // numSamples x numChains
    double* trainingSamples;                  // numObservations x numSamples x numChains
    double* testSamples;                      // numTestObservations x numSamples x numChains, NULL without test rows
    std::uint32_t* variableCountSamples;      // numPredictors x numSamples x numChains
    double* kSamples;                         // numSamples x numChains, NULL unless k is modelled

    std::size_t numObservations, numPredictors, numTestObservations, numSamples, numChains;

    Results(std::size_t numObservations, std::size_t numPredictors, std::size_t numTestObservations, std::size_t numSamples,
            std::size_t numChains, bool kIsModeled) :
      sigmaSamples(NULL), trainingSamples(NULL), testSamples(NULL), variableCountSamples(NULL), kSamples(NULL),
      numObservations(numObservations), numPredictors(numPredictors), numTestObservations(numTestObservations),
      numSamples(numSamples), numChains(numChains)
    {
      sigmaSamples = new double[getNumSigmaSamples()];
      trainingSamples = new double[getNumTrainingSamples() > 0 ? getNumTrainingSamples() : 1];
      if (numTestObservations > 0) testSamples = new double[getNumTestSamples()];
      variableCountSamples = new std::uint32_t[getNumVariableCountSamples() > 0 ? getNumVariableCountSamples() : 1];
      if (kIsModeled) kSamples = new double[getNumSigmaSamples()];
    }
    ~Results() { delete [] kSamples; delete [] variableCountSamples; delete [] testSamples; delete [] trainingSamples; delete [] sigmaSamples; }

    std::size_t getNumSigmaSamples() const { return numSamples * numChains; }
    std::size_t getNumTrainingSamples() const { return numObservations * numSamples * numChains; }
    std::size_t getNumTestSamples() const { return numTestObservations * numSamples * numChains; }
    std::size_t getNumVariableCountSamples() const { return numPredictors * numSamples * numChains; }

   private:
    Results(const Results&);
    Results& operator=(const Results&);
  };
}

#endif
