// include/dbarts_shim/dbarts/bartFit.hpp -- shim of <dbarts/bartFit.hpp> (see control.hpp).  The reference allocates raw storage
// of sizeof(BARTFit) and has initializeFit construct into it (/root/reference/src/init.cpp:227-228, :429-430), releases it with
// invalidateFit + ::operator delete (:161-164), and reads these members directly: control (:264), data.numPredictors /
// numTestObservations (:269-270, :371), model.kPrior->isFixed (:272), sharedScratch.dataScale.{min,max,range} (:324-325,
// :433-435), currentNumSamples (:375, :466, :538).  Everything else of dbarts' BARTFit is replaced by the device sampler
// behind `impl` (gpubart_fit / gpubart_stored of include/stan4bart_b200.h).
#ifndef DBARTS_BART_FIT_HPP
#define DBARTS_BART_FIT_HPP

#include <cstddef>
#include <cstdint>

#include "control.hpp"
#include "data.hpp"
#include "model.hpp"

namespace dbarts {
  struct Results;

  struct DataScale { double min, max, range; };
  struct SharedScratch { DataScale dataScale; };

  // getTrees: parallel arrays, one entry per node in depth-first (pre-order) order; variable < 0 marks a bottom node, value is
  // the cut value or the bottom node's mu (init.cpp:583-666; the caller frees the arrays with delete [] and the struct with
  // ::operator delete)
  struct FlattenedTrees {
    std::size_t totalNumNodes;
    std::size_t* chainNumber;
    std::size_t* sampleNumber;
    std::size_t* treeNumber;
    std::size_t* numObservations;
    std::int32_t* variable;
    double* value;
  };

  struct BARTFit {
    Control control;
    Model model;
    Data data;
    SharedScratch sharedScratch;
    std::size_t currentNumSamples;      // stored draws so far (keepTrees)

    // ---- device side (opaque to the reference) ----
    void* impl;          // gpubart_fit*: the live sampler
    void* stored;        // gpubart_stored*: draws imported with initializeState (prediction only)
    bool storeEnabled;   // a tree store of control.defaultNumSamples draws exists on the device
    std::size_t storeBase;   // draws in the device store that were discarded by a later setControl(keepTrees) cycle
  };
}

#endif
