// include/dbarts_shim/dbarts/model.hpp -- shim of <dbarts/model.hpp> (see control.hpp).  The reference constructs `Model(false)`
// (/root/reference/src/init.cpp:109, :153) and reads `model.kPrior->isFixed` (:272, :731).  The priors are the ones stan4bart
// configures: cgm tree prior, normal end-node prior, fixed(1) residual prior, fixed or chi(df, scale) hyperprior on k
// (R/stan4bart_fit.R:456-479).
#ifndef DBARTS_MODEL_HPP
#define DBARTS_MODEL_HPP

#include <cstddef>

namespace dbarts {
  struct TreePrior {              // cgm
    double base, power;
    double* splitProbabilities;   // per predictor, or NULL (uniform); owned
    TreePrior() : base(0.95), power(2.0), splitProbabilities(NULL) { }
  };
  struct EndNodePrior {           // normal(k): mu ~ N(0, (node.scale / (k sqrt(numTrees)))^2)
    double k;
    EndNodePrior() : k(2.0) { }
  };
  struct ResidualVariancePrior {  // stan4bart always passes fixed(1): sigma comes from Stan through setSigma
    bool isFixed;
    double value;
    ResidualVariancePrior() : isFixed(true), value(1.0) { }
  };
  struct EndNodeHyperprior {      // fixed(k) or chi(degreesOfFreedom, scale)
    bool isFixed;
    double k;
    double degreesOfFreedom, scale;
    EndNodeHyperprior() : isFixed(true), k(2.0), degreesOfFreedom(1.25), scale(0.0) { }
  };

  struct Model {
    double birthOrDeathProbability, swapProbability, changeProbability, birthProbability;
    double nodeScale;             // 0.5 continuous, 3.0 binary (R/stan4bart_fit.R:479)
    TreePrior* treePrior;
    EndNodePrior* muPrior;
    ResidualVariancePrior* sigmaSqPrior;
    EndNodeHyperprior* kPrior;

    explicit Model(bool /* allocate: the reference always passes false and lets initializeModel fill the priors */ = false) :
      birthOrDeathProbability(0.5), swapProbability(0.1), changeProbability(0.4), birthProbability(0.5), nodeScale(0.5),
      treePrior(NULL), muPrior(NULL), sigmaSqPrior(NULL), kPrior(NULL) { }
  };
}

#endif
