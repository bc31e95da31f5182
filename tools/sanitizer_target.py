"""compute-sanitizer target: a few sweeps of the pipelined / synchronous sweep kernels and of the GLMM passes on a small problem.
usage: compute-sanitizer --tool memcheck|racecheck|synccheck python tools/sanitizer_target.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from stan4bart_b200.frontend import friedman_problem
from stan4bart_b200.sampler import GpuBart, GlmmModel, Sampler
from stan4bart_b200.structs import bart_config, stan_control

n = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
pr = friedman_problem(n, binary=True, seed=3)
g = GpuBart(bart_config(n, 9, num_trees=30, seed=1, is_binary=True), pr["y"], pr["x_bart"])
for _ in range(4):
    g.run()
print("bart", g.pipeline())
for bulk in ("0", "1"):
    os.environ["S4B_GLMM_BULK"] = bulk
    m = GlmmModel(pr["stan_data"])
    m.set_mode(0)
    q = np.random.default_rng(1).uniform(-0.5, 0.5, m.d)
    print("glmm bulk=" + bulk, m.log_prob_grad(q)[0])
s = Sampler(bart_config(n, 9, n_test=n, num_trees=30, is_binary=True, seed=5), pr["y"], pr["x_bart"], pr["x_test"], pr["stan_data"], stan_control(seed=6),
            warmup=3, iter_=6, keep_fits=False)
s.run(3, True)
print("gibbs ok")
