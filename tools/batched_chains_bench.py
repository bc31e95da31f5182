"""Several chains batched in one launch (k_sweep_batch, grid.y = chain) against the same chains launched one after the other and on
concurrent host threads / streams: BART sweeps per second, config D shape (IHDP-like, 25 covariates, 200 trees, 8 chains per GPU).
usage: python tools/batched_chains_bench.py [n] [chains] [sweeps]"""
import json, os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from stan4bart_b200.frontend import ihdp_problem
from stan4bart_b200.sampler import GpuBart
from stan4bart_b200.structs import bart_config

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
chains = int(sys.argv[2]) if len(sys.argv) > 2 else 8
sweeps = int(sys.argv[3]) if len(sys.argv) > 3 else 40
pr = ihdp_problem(n)
p = pr["x_bart"].shape[1]
sms = 148
out = {"workload": "config D shape: n=%d, p=%d, 200 trees, %d chains on one GPU, %d SMs per chain, BART half only" % (n, p, chains, sms // chains)}
mk = lambda c, pipe: bart_config(n, p, num_trees=200, seed=100 + c, max_ctas=sms // chains)
def make(pipe):
    fits = [GpuBart(mk(c, pipe), pr["y"], pr["x_bart"]) for c in range(chains)]
    for f in fits:
        f.set_sigma(1.0); f.sample_trees_from_prior(); f.set_pipeline(pipe)
    return fits
def timed(fn, fits):
    for _ in range(10):
        fn(fits)
    fits[0].latents()
    t0 = time.time()
    for _ in range(sweeps):
        fn(fits)
    for f in fits:
        f.rng_counter()
    return chains * sweeps / (time.time() - t0)
fits = make(False)
out["batched_one_launch"] = timed(lambda fs: GpuBart.run_batched(fs, results=False), fits)
out["one_after_the_other_synchronous_kernel"] = timed(lambda fs: [GpuBart.run_batched([f], results=False) for f in fs], fits)
del fits
fits = make(True)
out["one_after_the_other_pipelined_kernel"] = timed(lambda fs: [GpuBart.run_batched([f], results=False) if False else f.run() for f in fs], fits)
print(json.dumps(out))
