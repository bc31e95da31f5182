// include/dbarts_shim/dbarts/control.hpp -- shim of dbarts' <dbarts/control.hpp> for building stan4bart's src/init.cpp against
// the GPU sampler instead of the dbarts package (SURVEY.md 8b "C++ <-> BART").  dbarts itself is not vendored in the reference
// tree; this header exposes the members the reference reads and writes (/root/reference/src/init.cpp:216-222, :264-271, :375,
// :389-394, :463-467, :535-539, :737-743; src/bart_util.cpp:30-63) plus what the shim needs to configure the device sampler.
#ifndef DBARTS_CONTROL_HPP
#define DBARTS_CONTROL_HPP

#include <cstddef>
#include <cstdint>

namespace dbarts {
  struct Control {
    bool responseIsBinary;      // init.cpp:222
    bool verbose;               // init.cpp:265-266, :295
    bool keepTrainingFits;
    bool useQuantiles;          // cut points between the distinct sorted values (s4b_bart_config::use_quantiles)
    bool keepTrees;             // init.cpp:216-217, :739-741, :375
    std::size_t defaultNumSamples;   // init.cpp:219
    std::size_t defaultNumBurnIn;    // init.cpp:220
    std::size_t numTrees;            // init.cpp:465, :539
    std::size_t numChains;           // init.cpp:271, :389, :426
    std::size_t numThreads;
    std::uint32_t treeThinningRate;  // n.thin = skip.bart (R/stan4bart_fit.R:439)
    std::uint32_t printEvery;
    std::uint32_t printCutoffs;
    std::uint64_t rngSeed;           // dbartsControl rngSeed; stan4bart leaves it NA and seeds R's generator (man/stan4bart.Rd:209-228)
    std::uint32_t minNumObservationsInNode;   // dbarts' compile-time default 5, settable here

    Control() :
      responseIsBinary(false), verbose(false), keepTrainingFits(true), useQuantiles(false), keepTrees(false),
      defaultNumSamples(1), defaultNumBurnIn(0), numTrees(75), numChains(1), numThreads(1), treeThinningRate(1),
      printEvery(100), printCutoffs(0), rngSeed(0), minNumObservationsInNode(5) { }
  };
}

#endif
