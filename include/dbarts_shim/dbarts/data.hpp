// include/dbarts_shim/dbarts/data.hpp -- shim of <dbarts/data.hpp> (see control.hpp).  Members read by the reference:
// numObservations, numPredictors, numTestObservations (/root/reference/src/init.cpp:230, :269-270, :339, :371, :699-702, :729-730).
// The arrays stay owned by the caller (R objects in the reference), column major like dbartsData's x / x.test.
#ifndef DBARTS_DATA_HPP
#define DBARTS_DATA_HPP

#include <cstddef>
#include <cstdint>

namespace dbarts {
  enum VariableType { ORDINAL, CATEGORICAL };

  struct Data {
    const double* y;
    const double* x;                // numObservations x numPredictors, column major
    const double* x_test;           // numTestObservations x numPredictors, column major, or NULL
    const double* weights;          // or NULL
    const double* offset;           // or NULL
    const double* testOffset;       // or NULL
    std::size_t numObservations;
    std::size_t numPredictors;
    std::size_t numTestObservations;
    double sigmaEstimate;
    const VariableType* variableTypes;   // ordinal only on the device
    const std::uint32_t* maxNumCuts;     // n.cuts per predictor (owned: freed by invalidateData)

    Data() : y(NULL), x(NULL), x_test(NULL), weights(NULL), offset(NULL), testOffset(NULL), numObservations(0), numPredictors(0),
             numTestObservations(0), sigmaEstimate(1.0), variableTypes(NULL), maxNumCuts(NULL) { }
  };
}

#endif
