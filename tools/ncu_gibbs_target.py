"""ncu target: a few whole Gibbs sweeps of config C (binary probit Friedman, n = 1 M, 200 trees, n_test = n) on one chain, plus
the stand-alone leaf-statistics pass -- for captures of the kernels around the sweep kernel (k_glmm_data_terms, k_finish_sweep,
k_apply_offset_binary, k_glmm_linear_predictor, k_leaf_stats, ...).
usage: python tools/ncu_gibbs_target.py [n] [sweeps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stan4bart_b200.frontend import friedman_problem
from stan4bart_b200.sampler import Sampler
from stan4bart_b200.structs import bart_config, stan_control

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
sweeps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
pr = friedman_problem(n, binary=True, seed=99)
s = Sampler(bart_config(n, 9, n_test=n, num_trees=200, is_binary=True, seed=4711), pr["y"], pr["x_bart"], pr["x_test"], pr["stan_data"],
            stan_control(seed=4712), warmup=sweeps, iter_=2 * sweeps, keep_fits=False)
s.run(sweeps, True)
b = s.bart()
for t in (0, 100, 199):
    b.time_leaf_stats(t, 2)
print("done")
