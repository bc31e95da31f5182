"""Per-phase SM-cycle counters of the BART sweep (see gpubart_get_profile) at the benchmark shape.
usage: python tools/phase_profile.py [n] [bart|gibbs] [binary|continuous]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stan4bart_b200.frontend import friedman_problem
from stan4bart_b200.sampler import GpuBart, Sampler
from stan4bart_b200.structs import bart_config, stan_control

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
what = sys.argv[2] if len(sys.argv) > 2 else "bart"
binary = (sys.argv[3] if len(sys.argv) > 3 else "continuous") == "binary"
pr = friedman_problem(n, binary=binary)
cfg = bart_config(n, 9, n_test=n if what == "gibbs" else 0, num_trees=200, seed=1, is_binary=binary)
if what == "gibbs":
    s = Sampler(cfg, pr["y"], pr["x_bart"], pr["x_test"], pr["stan_data"], stan_control(seed=2), warmup=40, iter_=100, keep_fits=False,
                sigma_init=pr["sigma_init"], bart_offset_init=pr["bart_offset_init"])
    s.run(40, True, results=False)
    s.disengage_adaptation()
    g = s.bart()
    g.set_profile(True); g.profile(); g.tree_step_ms()
    t0 = time.time()
    s.run(20, False, results=False)
    dt = time.time() - t0
    print("gibbs ms/sweep wall", dt / 20 * 1e3, "bart device ms/sweep", g.tree_step_ms() / 20, s.last_run_stats())
else:
    g = GpuBart(cfg, pr["y"], pr["x_bart"])
    if not binary:
        g.set_sigma(1.0)
    for _ in range(30):
        g.run()
    g.set_profile(True); g.profile(); g.tree_step_ms()
    t0 = time.time()
    for _ in range(20):
        g.run()
    dt = time.time() - t0
    print("ms/sweep wall", dt / 20 * 1e3, "device ms/sweep", g.tree_step_ms() / 20)
prof = g.profile()
names = ["pass(cta0)", "pass+barrier+helper", "reduce", "decide", "seq_propose", "update", "-"]
print(json.dumps({"steps": prof["steps"], "cycles_per_step": dict(zip(names, prof["cycles_per_step"].values())),
                  "controller": prof["controller_cycles_per_step"], "worker": prof["worker_cycles_per_step"], "decision_fine": prof["decision_fine_cycles_per_step"]}, indent=1))
tr = g.trees()
print("nodes per tree", len(tr["var"]) / 200, "sweep mode", g.sweep_mode())
