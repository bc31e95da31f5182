"""ncu target: k_leaf_stats on a synthetic 32 M-row fit (the HBM-bound size of bench.py's roofline.leaf_stat.hbm_bound_size).
usage: ncu --set full --clock-control none --import-source on -k regex:k_leaf_stats -s 4 -c 1 -o gpurun_out/leaf python tools/ncu_leaf_target.py [n]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from stan4bart_b200.sampler import GpuBart
from stan4bart_b200.structs import bart_config

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32_000_000
rg = np.random.default_rng(1)
x = np.asfortranarray(rg.random((n, 3)))
y = 10 * np.sin(np.pi * x[:, 0] * x[:, 1]) + 5 * x[:, 2] + rg.standard_normal(n)
g = GpuBart(bart_config(n, 3, num_trees=8, seed=3), y, x)
g.set_sigma(1.0)
g.sample_trees_from_prior()
for _ in range(int(os.environ.get("LEAF_SWEEPS", "1"))):
    g.run()
tr = g.trees()
for t in range(8):
    rules = int(np.sum(tr["var"][tr["tree"] == t] >= 0))
    if 1 <= rules <= 7:
        ms = g.time_leaf_stats(t, 8)
        print("tree", t, "rules", rules, "us", ms * 1e3, "frac 11B", 11.0 * n / ms / 1e6 / 6546.6, flush=True)
