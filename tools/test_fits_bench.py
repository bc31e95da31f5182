"""k_test_fits at scale: the treatment z INSIDE bart() (a 10th BART column), so the counterfactual test design (z flipped) differs
from the training design and the test fits are computed by tree traversal every sweep instead of being aliased to the training fits.
usage: python tools/test_fits_bench.py [n] [trees] [sweeps]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from stan4bart_b200.frontend import friedman_data
from stan4bart_b200.sampler import GpuBart
from stan4bart_b200.structs import bart_config

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
T = int(sys.argv[2]) if len(sys.argv) > 2 else 200
sweeps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
d = friedman_data(n, ranef=False, causal=True, binary=True, seed=99)
x = np.asfortranarray(np.column_stack([d["x"][:, [0, 1, 2, 4, 5, 6, 7, 8, 9]], d["z"]]))
xt = x.copy(order="F"); xt[:, 9] = 1.0 - xt[:, 9]
out = {}
for label, x_test in (("with_counterfactual_test_fits", xt), ("no_test_sample", None)):
    g = GpuBart(bart_config(n, 10, n_test=n if x_test is not None else 0, num_trees=T, seed=1, is_binary=True), d["y"], x, x_test)
    for _ in range(30):
        g.run()
    import ctypes as C
    from stan4bart_b200 import _lib
    t0 = time.time()
    for _ in range(sweeps):
        _lib.check(g.L.gpubart_run_sampler_with_results(g.h, None, None, None, None))
    g.data_range()           # synchronises
    out[label] = {"ms_per_sweep_wall": (time.time() - t0) / sweeps * 1e3}
    del g
dt = out["with_counterfactual_test_fits"]["ms_per_sweep_wall"] - out["no_test_sample"]["ms_per_sweep_wall"]
out["k_test_fits_ms_per_sweep"] = dt
out["k_test_fits_gbs"] = (10.0 + 8.0) * n / (dt * 1e-3) / 1e9 if dt > 0 else None
out["note"] = "per test row: 10 binned predictor bytes read + 8 bytes written; all %d trees staged in shared memory, one thread per row" % T
out["workload"] = "binary probit Friedman, n = n_test = %d, 10 BART columns (treatment inside bart()), %d trees" % (n, T)
print(json.dumps(out))
