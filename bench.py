#!/usr/bin/env python
"""bench.py -- Gibbs sweeps/sec of the stan4bart hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (config C of BASELINE.json): binary-probit Friedman causal data, n = 1 000 000, 200 trees,
9 BART predictors, fixed effects X4 + z, (1 + X4 | g.1) + (1 | g.2), counterfactual test design
(n_test = n), one chain per GPU (weak scaling, no collective on the data path).
A "step" is one full Gibbs sweep: Stan block (one NUTS transition, every gradient on device) + BART block
(200 tree updates) + plumbing.  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

METRIC = "gibbs_sweeps_per_sec"
UNIT = "sweeps/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", "--rows", dest="n", type=int, default=1_000_000, help="observations (use --rows under torchrun: its parser trips over --n)")
    ap.add_argument("--trees", type=int, default=200)
    ap.add_argument("--adapt", type=int, default=200, help="adaptation sweeps before adaptation is disengaged (untimed; >= 150 so that the metric windows of Stan run)")
    ap.add_argument("--continuous", action="store_true", help="continuous response (configs B / E) instead of the probit model of config C")
    ap.add_argument("--weighted", action="store_true", help="observation weights (`weights` of stan4bart()): a non-default branch, not the headline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--shard-rows", action="store_true",
                    help="N > 1 only: ONE chain whose --n rows are sharded over the N GPUs (BASELINE config E; strong scaling) instead of "
                         "one independent chain per GPU")
    ap.add_argument("--cpu-budget-s", type=float, default=25.0)
    ap.add_argument("--ref-budget-s", type=float, default=150.0)
    return ap.parse_args()


def workload_config(args):
    what = "continuous Friedman causal" if args.continuous else "config C: binary probit Friedman causal"
    if args.weighted:
        what += " with observation weights"
    return {"workload": "%s, n=%d, %d trees, p_bart=9, K=2, q=18, n_test=%d, "
                        "1 chain per GPU" % (what, args.n, args.trees, args.n),
            "n": args.n, "trees": args.trees, "chains_per_gpu": 1, "parallelism": "chain-per-GPU, no data-path collective",
            "l2": "inputs larger than L2, no explicit flush: one sweep streams ~%d MB of distinct N-length arrays (BART: binned X, "
                  "residual, response, offset, fits, latents; GLMM: X, Z index / value streams, response, offset, residual; running "
                  "means) against 126 MB of L2, so every step's kernels start from HBM; inside k_sweep the chain's residuals and "
                  "predictors are then held on chip for the 200 tree steps by design" % int(round((9 + 8 * 19 + 12 + 8) * args.n / 1e6)),
            "adapt_sweeps": args.adapt}


def make_problem(args):
    from stan4bart_b200.frontend import friedman_problem
    pr = friedman_problem(args.n, binary=not args.continuous, seed=99)
    return add_weights(pr, args.n) if args.weighted else pr


def add_weights(pr, n):
    pr["weights"] = np.random.default_rng(4).gamma(3.0, 1.0 / 3.0, n)
    pr["stan_data"].weights = pr["weights"]
    return pr


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """`nvidia-smi -lms 100` running from before the warm-up steps until after the timed region (the timed region itself
    is only tens of milliseconds, shorter than one nvidia-smi start-up)."""

    QUERY = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.QUERY, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return

        def pump():
            for line in self.proc.stdout:
                self.lines.append(line.strip())
        self.thread = threading.Thread(target=pump, daemon=True)
        self.thread.start()

    def stop(self):
        if self.proc is not None:
            time.sleep(0.25)            # let at least one more sample land
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def summary(self):
        samples = [[x.strip() for x in ln.split(",")] for ln in self.lines if ln]
        sm, mx, reasons = [], [], set()
        for smp in samples:
            try:
                sm.append(float(smp[0])); mx.append(float(smp[1]))
            except Exception:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), smp[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm),
                "window": "sampled every 100 ms from the warm-up steps through the timed and e2e regions"}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def avg_levels(trees, n):
    """Mean number of tree levels an observation walks, from the flattened trees (pre-order, n per node)."""
    tot = 0.0
    tree_ids = trees["tree"]
    for t in np.unique(tree_ids):
        sel = np.nonzero(tree_ids == t)[0]
        var, cnt = trees["var"][sel], trees["n"][sel]
        stack = []      # remaining children counts
        depth = 0
        for k in range(len(sel)):
            d = len(stack)
            if var[k] >= 0:
                stack.append(2)
            else:
                tot += d * float(cnt[k])
                while stack:
                    stack[-1] -= 1
                    if stack[-1] == 0:
                        stack.pop()
                    else:
                        break
    ntrees = len(np.unique(tree_ids))
    return tot / (float(n) * ntrees)


def cpu_baseline(args, pr, budget_s):
    """The CPU oracle (port of the reference algorithm, one thread per chain as the reference configures dbarts,
    R/stan4bart_fit.R:437-439) timed on a bounded sample of the same workload."""
    import oracle_lib as O
    from stan4bart_b200.structs import bart_config, stan_control
    sd = pr["stan_data"]
    cfg = bart_config(args.n, 9, n_test=args.n, num_trees=args.trees, is_binary=not args.continuous, seed=12345, weights=pr.get("weights"))
    extra = {"sigma_init": pr["sigma_init"], "bart_offset_init": pr["bart_offset_init"]} if args.continuous else {}
    t0 = time.time()
    s = O.OracleSampler(cfg, pr["y"], pr["x_bart"], pr["x_test"], sd, stan_control(seed=1), warmup=10, iter_=20, keep_fits=False, **extra)
    t_create = time.time() - t0
    t0 = time.time()
    s.run(1, True)
    t_first = time.time() - t0
    k = int(max(1, min(5, (budget_s - t_create - t_first) // max(t_first, 1e-3))))
    t0 = time.time()
    s.run(k, True)
    dt = time.time() - t0
    return {"value": k / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": "CPU restatement of the reference algorithm (oracle/, not the reference binary): same n=%d, %d trees, "
                      "1 chain on 1 thread, %d full sweeps after 1 warm-up sweep (setup %.1f s excluded); these are early, "
                      "un-adapted sweeps with few leapfrogs each, i.e. the CPU's best case" % (args.n, args.trees, k, t_create)}


# ------------------------------------------------------------------------------------------------
def _ref_worker(args_dict, seed, conn):
    try:
        import oracle_lib as O
        from stan4bart_b200.frontend import friedman_problem
        from stan4bart_b200.structs import bart_config, stan_control
        n, trees = args_dict["n"], args_dict["trees"]
        binary = not args_dict.get("continuous", False)
        pr = friedman_problem(n, binary=binary, seed=99)
        if args_dict.get("weighted"):
            pr = add_weights(pr, n)
        cfg = bart_config(n, 9, n_test=n, num_trees=trees, is_binary=binary, seed=seed, weights=pr.get("weights"))
        extra = {} if binary else {"sigma_init": pr["sigma_init"], "bart_offset_init": pr["bart_offset_init"]}
        s = O.OracleSampler(cfg, pr["y"], pr["x_bart"], pr["x_test"], pr["stan_data"], stan_control(seed=seed), warmup=10, iter_=20,
                            keep_fits=False, **extra)
        conn.send(("ready", 0.0))
        while True:
            cmd = conn.recv()
            if cmd[0] == "run":
                t0 = time.time()
                s.run(cmd[1], True)
                conn.send(("done", time.time() - t0))
            else:
                break
    except Exception as e:  # pragma: no cover
        conn.send(("error", repr(e)))


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  The real reference cannot be built here
    (needs R + dbarts + Eigen/Boost/TBB), so this times the oracle port, parallelised the way the reference is:
    one single-threaded process per chain (R/stan4bart_fit.R:495-542)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    chains = max(1, args.gpus)
    cores = os.cpu_count() or 1
    procs = min(chains, cores)
    ctx = mp.get_context("fork")
    workers = []
    for c in range(procs):
        a, b = ctx.Pipe()
        p = ctx.Process(target=_ref_worker, args=(dict(n=args.n, trees=args.trees, continuous=args.continuous, weighted=args.weighted), 12345 + c, b), daemon=True)
        p.start()
        workers.append((p, a))
    for _, a in workers:
        tag, val = a.recv()
        if tag != "ready":
            raise RuntimeError("reference worker failed: %s" % (val,))

    def run_all(k):
        t0 = time.time()
        for _, a in workers:
            a.send(("run", k))
        for _, a in workers:
            tag, val = a.recv()
            if tag != "done":
                raise RuntimeError("reference worker failed: %s" % (val,))
        return time.time() - t0

    t_first = run_all(1)                                   # warm-up sweep, also sizes the bounded sample
    warm_done = 1
    k = int(max(1, min(args.steps, (args.ref_budget_s - t_first) // max(t_first, 1e-3))))
    dt = run_all(k)
    for _, a in workers:
        a.send(("stop",))
    value = procs * k / dt
    sample = ("oracle port (CPU restatement of the reference algorithm, not the reference binary), %d chain(s) in %d "
              "single-threaded process(es), n=%d, %d trees, %d full sweeps timed after %d warm-up sweep "
              "(requested steps=%d warmup=%d bounded by --ref-budget-s)" % (procs, procs, args.n, args.trees, k, warm_done, args.steps, args.warmup))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": k, "warmup": warm_done,
            "ms_per_step": 1000.0 * dt / k, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config(args),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    from stan4bart_b200 import _lib
    from stan4bart_b200.sampler import Sampler
    from stan4bart_b200.structs import bart_config, stan_control

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly one line, the JSON: anything libraries print on fd 1 meanwhile (NCCL's version banner) goes to stderr
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl")
    torch.cuda.set_device(local_rank)
    L = _lib.load()
    _lib.require_device()
    _lib.check(L.s4b_set_device(local_rank))
    stream = torch.cuda.Stream()
    _lib.check(L.s4b_set_stream(stream.cuda_stream))

    from stan4bart_b200.dist import barrier, chain_seed, max_over_ranks

    pr = make_problem(args)
    n, T = args.n, args.trees
    sharded = bool(args.shard_rows) and world > 1
    shard_ctx = None
    chains = world
    if sharded:
        # one chain, rows dealt out in contiguous blocks; every rank runs the replicated controller with the same seeds
        from stan4bart_b200.frontend import shard_problem
        from stan4bart_b200.shard import ShardContext, row_range
        shard_ctx = ShardContext.from_torch_distributed(total_obs=n)
        lo, hi = row_range(n, rank, world)
        pr = shard_problem(pr, lo, hi)
        n = hi - lo
        chains = 1
    seed_rank = 0 if sharded else rank
    sd = pr["stan_data"]
    cfg = bart_config(n, 9, n_test=n, num_trees=T, is_binary=not args.continuous, seed=chain_seed(12345, seed_rank), weights=pr.get("weights"))
    s = Sampler(cfg, pr["y"], pr["x_bart"], pr["x_test"], sd, stan_control(seed=chain_seed(1000, seed_rank)), warmup=args.adapt,
                iter_=args.adapt + args.steps, keep_fits=False, shard=shard_ctx,
                **({"sigma_init": pr["sigma_init"], "bart_offset_init": pr["bart_offset_init"]} if args.continuous else {}))
    bart = s.bart()
    glmm = s.glmm()
    s.run(args.adapt, True, results=False)
    s.disengage_adaptation()
    W, K = max(3, args.warmup), args.steps

    # ---- device-resident leg: `value` ----
    clocks = ClockSampler(local_rank)
    clocks.start()
    s.run(W, False, results=False)
    bart.tree_step_ms(reset=True)
    passes0 = glmm.num_device_passes()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        ev0.record(stream)
        t_wall0 = time.time()
        out = s.run(K, False, results=False)
        ev1.record(stream)
    barrier()
    t_wall = time.time() - t_wall0
    ms = max_over_ranks(float(ev0.elapsed_time(ev1)))
    stats = s.last_run_stats()
    glmm_passes = glmm.num_device_passes() - passes0
    sweep_ms = bart.tree_step_ms(reset=True)
    names = sd.param_names()
    n_leapfrog = float(out["stan"][names.index("n_leapfrog__")][-1])
    value = chains * K / (ms / 1000.0)

    # ---- roofline of the dominant kernel ----
    trees = bart.trees()
    lv = avg_levels(trees, n)
    # 29 algorithmic bytes per (tree x observation) -- statistics pass 11 B (residual 8 + node id 2 + split column 1)
    # + update pass 18 B (residual read 8 + write 8 + node id 2).  The leaner count for THIS design (residual read + write and
    # one u8 per tree level of the two walks) is reported beside it.
    bytes_per_obs = 29.0
    bytes_per_obs_lean = 16.0 + 2.0 * lv
    peak, peak_src = measured_peak()
    persistent = bart.sweep_mode() == 2
    launch_ms = sweep_ms / K if persistent else sweep_ms / (K * T)    # CUDA events around the sweep launches on the launching stream
    units_per_launch = (T if persistent else 1) * n
    achieved = bytes_per_obs * units_per_launch / (launch_ms * 1e-3) / 1e9
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f).get("k_sweep_dram_bytes_per_launch" if persistent else "k_tree_step_dram_bytes_per_launch")
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": "k_sweep (one launch = one %d-tree sweep)" % T if persistent else "k_tree_step (one launch = one tree)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_tree_obs": bytes_per_obs,
                "units_per_launch": units_per_launch, "avg_tree_levels": lv,
                "launch_us": launch_ms * 1e3, "tree_step_us": sweep_ms / (K * T) * 1e3,
                "achieved_lean_model_gbs": bytes_per_obs_lean * units_per_launch / (launch_ms * 1e-3) / 1e9,
                "lean_bytes_per_tree_obs": bytes_per_obs_lean,
                "note": "algorithmic bytes are what per-tree streaming passes must move (SURVEY.md 8d); the persistent kernel keeps residuals "
                        "in registers and predictors in shared memory, so DRAM traffic per launch is ~17 MB and the binding limit is the "
                        "latency of 200 sequential reduce + Metropolis decisions, not HBM (DESIGN.md section 4)"}

    # ---- end-to-end leg through the C ABI with host buffers ----
    h2d, d2h = s.set_host_plumbing(True)
    pin_train = torch.empty(n, dtype=torch.float64).pin_memory()
    pin_test = torch.empty(n, dtype=torch.float64).pin_memory()
    pin_stan = torch.empty(s.num_pars, dtype=torch.float64).pin_memory()
    s.run_into(W, False, stan=pin_stan.data_ptr(), train=pin_train.data_ptr(), test=pin_test.data_ptr())
    barrier()
    t0 = time.time()
    s.run_into(K, False, stan=pin_stan.data_ptr(), train=pin_train.data_ptr(), test=pin_test.data_ptr())
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.time() - t0)
    s.set_host_plumbing(False)
    clocks.stop()
    e2e = {"value": chains * K / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h + 2 * 8 * n + 8 * s.num_pars),
           "note": "s4b_sampler_run with host result buffers (train + test fits, Stan row) and every N-vector of the sweep "
                   "(parametric mean, BART fit, latents) round-tripped through pinned host memory like the reference's host vectors"}

    # kernels launched inside the timed region, per sweep: k_prepare_sweep + k_sweep + epilogue + epoch bump (BART block),
    # the offset kernel + its epoch bump (1 kernel for a continuous response), parametric mean, the fused GLMM input refresh,
    # the fused running-mean accumulation, plus the GLMM data passes
    per_sweep_bart = 4 if bart.sweep_mode() == 2 else T + 3
    launches = K * (per_sweep_bart + (1 if args.continuous else 2) + 1 + 1 + 1) + glmm_passes

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K,
            "higher_is_better": True, "scaling": "strong" if sharded else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": dict(workload_config(args), **({"sharding": f"one chain, rows sharded over {world} GPUs ({n} rows on rank 0)"} if sharded else {})),
            "clocks": clocks.summary(), "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roofline,
            "breakdown": {"ms_stan_block": stats["ms_stan"] / K, "ms_bart_block": stats["ms_bart"] / K, "grad_evals_per_sweep": stats["grad_evals"] / K,
                          "glmm_device_passes_per_sweep": glmm_passes / K, "glmm_mode": glmm.mode(), "bart_sweep_mode": bart.sweep_mode(),
                          "n_leapfrog_last": n_leapfrog, "wall_s": t_wall, "tree_step_us": sweep_ms / (K * T) * 1e3}}
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args, pr, args.cpu_budget_s)
        else:
            line["cpu_baseline"] = None
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
