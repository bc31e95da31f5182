"""A multilevel model with many grouping levels (random intercept + slope per level): what the column path of the GLMM data
pass and the sweep-level expansion with the sparse Gram matrix deliver.
usage: python tools/many_levels_bench.py [n] [levels] [sweeps] [trees]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stan4bart_b200.frontend import build_stan_data, init_fit
from stan4bart_b200.sampler import Sampler
from stan4bart_b200.structs import bart_config, stan_control

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
levels = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
sweeps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
trees = int(sys.argv[4]) if len(sys.argv) > 4 else 200
adapt = 60
rng = np.random.default_rng(1)
xb = np.asfortranarray(rng.random((n, 9)))
g = rng.integers(0, levels, n); g[:levels] = np.arange(levels)
Xf = rng.standard_normal((n, 2))
b0, b1 = rng.standard_normal(levels) * 0.7, rng.standard_normal(levels) * 0.3
y = 10 * np.sin(np.pi * xb[:, 0] * xb[:, 1]) + 5 * xb[:, 2] + Xf @ np.array([1.0, -0.5]) + b0[g] + b1[g] * xb[:, 3] + rng.standard_normal(n)
sd = build_stan_data(Xf, y, [(g, np.column_stack([np.ones(n), xb[:, 3]]))])
offset_init, sigma_init = init_fit(sd, False)
out = {"workload": "continuous response, n=%d, %d trees, one grouping factor with %d levels, random intercept + slope (K=2, q=%d)" % (n, trees, levels, sd.q)}
for mode in (1, 0):
    s = Sampler(bart_config(n, 9, n_test=n, num_trees=trees, seed=5), y, xb, xb.copy(order="F"), sd, stan_control(seed=6),
                warmup=adapt, iter_=adapt + 2 * sweeps, keep_fits=False, sigma_init=sigma_init, bart_offset_init=offset_init)
    s.glmm().set_mode(mode)
    k = sweeps if mode == 1 else max(2, sweeps // 5)
    s.run(adapt if mode == 1 else 10, True, results=False)
    s.run(2, True, results=False)
    t0 = time.time()
    s.run(k, True, results=False)
    dt = time.time() - t0
    st = s.last_run_stats()
    out["mode%d" % mode] = {"sweeps_per_s": k / dt, "ms_stan_block": st["ms_stan"] / k, "ms_bart_block": st["ms_bart"] / k,
                            "grad_evals_per_sweep": st["grad_evals"] / k,
                            "note": "sweep-level expansion, Gram matrix in pieces" if mode == 1 else "one device pass (column path) per gradient evaluation"}
    del s
print(json.dumps(out))
