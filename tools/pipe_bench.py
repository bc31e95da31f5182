"""BART-only timing of the sweep kernels at the benchmark shape: pipelined (sweep_pipe.cuh) against synchronous.
usage: python tools/pipe_bench.py [n] [trees] [sweeps]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from stan4bart_b200.frontend import friedman_problem
from stan4bart_b200.sampler import GpuBart
from stan4bart_b200.structs import bart_config

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
T = int(sys.argv[2]) if len(sys.argv) > 2 else 200
sweeps = int(sys.argv[3]) if len(sys.argv) > 3 else 30
pr = friedman_problem(n, binary=True)
out = {}
for mode in ("sync", "pipe"):
    cfg = bart_config(n, 9, num_trees=T, seed=1, is_binary=True)
    g = GpuBart(cfg, pr["y"], pr["x_bart"])
    g.set_pipeline(mode == "pipe")
    for _ in range(40):
        g.run()
    g.tree_step_ms()
    t0 = time.time()
    for _ in range(sweeps):
        g.run()
    dt = time.time() - t0
    dev = g.tree_step_ms()
    tr = g.trees()
    out[mode] = {"wall_ms_per_sweep": dt / sweeps * 1e3, "device_ms_per_sweep": dev / sweeps, "us_per_tree_step": dev / sweeps / T * 1e3,
                 "nodes_per_tree": len(tr["var"]) / T, "pipeline": g.pipeline()}
    del g
print(json.dumps({"n": n, "trees": T, "sweeps": sweeps, **out}, indent=1))
if os.environ.get("S4B_PIPE_PROF"):
    cfg = bart_config(n, 9, num_trees=T, seed=1, is_binary=True)
    g = GpuBart(cfg, pr["y"], pr["x_bart"])
    for _ in range(40):
        g.run()
    g.profile()          # reset the counters
    for _ in range(10):
        g.run()
    import ctypes as C
    from stan4bart_b200 import _lib
    out = (C.c_uint64 * 24)()
    _lib.check(g.L.gpubart_get_profile(g.h, out, 1))
    v = [int(x) for x in out]
    steps = max(1, v[5])
    print("pipe profile (cycles per pipelined step):", json.dumps({"steps": v[5], "worker": dict(zip(["wait_decision", "update", "walk", "accumulate", "reduce_arrive"], [x / steps for x in v[0:5]])),
          "controller": dict(zip(["plan_fetch", "wait_rows", "row_reduce", "decide", "deltas_arrive"], [x / steps for x in v[8:13]]))}))
