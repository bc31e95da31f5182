"""Several chains per GPU (BASELINE config D: IHDP-shaped, n = 500 000, 25 covariates, 8 chains per GPU): every chain has its
own host thread and stream, so one chain's host-side NUTS overlaps another chain's sweep kernel on the device.
With max_ctas > 0 every chain's sweep kernel is confined to that many SMs, so that the kernels of different chains are
resident side by side (8 chains x 18 SMs on a 148-SM B200) and hide each other's barrier and decision latency.
usage: python tools/multi_chain_bench.py [n] [chains] [sweeps] [trees] [max_ctas] [ihdp|configC]
(configC: the binary probit Friedman problem of the headline bench, to see what a second chain on the same GPU adds)"""
import json
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stan4bart_b200.frontend import friedman_problem, ihdp_problem
from stan4bart_b200.sampler import Sampler
from stan4bart_b200.structs import bart_config, stan_control

n = int(sys.argv[1]) if len(sys.argv) > 1 else 500000
chains = int(sys.argv[2]) if len(sys.argv) > 2 else 8
sweeps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
trees = int(sys.argv[4]) if len(sys.argv) > 4 else 200
max_ctas = int(sys.argv[5]) if len(sys.argv) > 5 else 0
problem = sys.argv[6] if len(sys.argv) > 6 else "ihdp"
adapt = 160
binary = problem == "configC"
pr = friedman_problem(n, binary=True, seed=99) if binary else ihdp_problem(n)
p_bart = pr["x_bart"].shape[1]
kw = dict(warmup=adapt, iter_=adapt + 3 * sweeps, keep_fits=False)
if not binary:
    kw.update(sigma_init=pr["sigma_init"], bart_offset_init=pr["bart_offset_init"])
samplers = [None] * chains
barrier = threading.Barrier(chains + 1)
group = None
if max_ctas > 0 and os.environ.get("S4B_NO_BATCH_LEG") is None:
    from stan4bart_b200.sampler import BatchGroup
    group = BatchGroup(chains)
times = {}


def work(c):
    cfg = bart_config(n, p_bart, n_test=n, num_trees=trees, is_binary=binary, seed=100 + c, max_ctas=max_ctas)
    s = Sampler(cfg, pr["y"], pr["x_bart"], pr["x_test"], pr["stan_data"], stan_control(seed=200 + c), **kw)
    s.run(adapt, True, results=False)
    s.disengage_adaptation()
    s.run(5, False, results=False)
    samplers[c] = s
    barrier.wait()          # everybody adapted
    barrier.wait()          # sequential leg done by the main thread
    s.run(sweeps, False, results=False)
    barrier.wait()          # threaded leg done
    if group is not None:   # the same chains with their BART sweeps batched into one launch per iteration (grid.y = chain)
        s.set_batch_group(group)
        s.run(3, False, results=False)
        barrier.wait()
        s.run(sweeps, False, results=False)
        barrier.wait()
        s.set_batch_group(None)


th = [threading.Thread(target=work, args=(c,)) for c in range(chains)]
for t in th:
    t.start()
barrier.wait()
t0 = time.time()
for s in samplers:          # one chain after the other on the main thread
    s.run(sweeps, False, results=False)
t_seq = time.time() - t0
stats = samplers[0].last_run_stats()
barrier.wait()
t0 = time.time()
barrier.wait()
t_par = time.time() - t0
t_batch = None
if group is not None:
    barrier.wait()
    t0 = time.time()
    barrier.wait()
    t_batch = time.time() - t0
for t in th:
    t.join()
what = "config C: binary probit Friedman" if binary else "config D shape: IHDP-like continuous"
print(json.dumps({"workload": "%s, n=%d, p=%d, %d trees, %d chains on one GPU, %s SMs per chain" % (what, n, p_bart, trees, chains, max_ctas or "all"),
                  "sweeps_per_s_sequential": chains * sweeps / t_seq, "sweeps_per_s_threaded": chains * sweeps / t_par,
                  "sweeps_per_s_batched_one_launch_per_iteration": (chains * sweeps / t_batch) if t_batch else None,
                  "ms_stan_block": stats["ms_stan"] / sweeps, "ms_bart_block": stats["ms_bart"] / sweeps,
                  "bart_sweep_mode": samplers[0].bart().sweep_mode()}))
