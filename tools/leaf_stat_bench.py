"""The stand-alone leaf-statistics pass (csrc/leaf_stats.cuh) at several sizes: achieved GB/s in SURVEY.md's 11 B / row model and in
bytes actually read (8 + rules of the tree), against the measured HBM peak.  Beyond ~10 M rows the working set (8 B residual +
1 B per rule and row) no longer fits the 126 MB L2, so back-to-back launches stream from HBM.
usage: python tools/leaf_stat_bench.py [n ...]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from stan4bart_b200.sampler import GpuBart
from stan4bart_b200.structs import bart_config

sizes = [int(a) for a in sys.argv[1:]] or [1_000_000, 4_000_000, 16_000_000, 32_000_000]
peak = 6546.6
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
out = []
for n in sizes:
    rng = np.random.default_rng(1)
    x = np.asfortranarray(rng.random((n, 9)))
    y = 10 * np.sin(np.pi * x[:, 0] * x[:, 1]) + 20 * (x[:, 2] - 0.5) ** 2 + 5 * x[:, 4] + rng.standard_normal(n)
    T = 20
    g = GpuBart(bart_config(n, 9, num_trees=T, seed=3), y, x)
    g.set_sigma(1.0)
    g.sample_trees_from_prior()
    for _ in range(3):
        g.run()
    tr = g.trees()
    for t in range(0, T, 5):
        sel = tr["tree"] == t
        rules = int(np.sum(tr["var"][sel] >= 0))
        ms = g.time_leaf_stats(t, 30)
        out.append({"n": n, "tree": t, "rules": rules, "bottom_nodes": int(np.sum(tr["var"][sel] < 0)), "launch_us": ms * 1e3,
                    "gbs_11B_model": 11.0 * n / ms / 1e6, "gbs_bytes_read": (8.0 + rules) * n / ms / 1e6,
                    "frac_of_peak_11B_model": 11.0 * n / ms / 1e6 / peak, "frac_of_peak_bytes_read": (8.0 + rules) * n / ms / 1e6 / peak,
                    "working_set_mb": (8.0 + rules) * n / 1e6})
    del g
print(json.dumps({"peak_gbs": peak, "note": "CUDA events around 30 back-to-back launches of k_leaf_stats", "runs": out}, indent=1))
