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
    ap.add_argument("--chains-per-gpu", type=int, default=3,
                    help="chains resident on every GPU, each on its own host thread and stream (value = aggregate over all chains; the "
                         "one-chain-per-GPU figure is reported beside it)")
    ap.add_argument("--scaled-levels", action="store_true",
                    help="grouping factors with max(5 | 8, n / 2000) levels each (SURVEY.md 8d: 'scale to max(5, N/2000) levels for big N and "
                         "report it'): q = 3 n / 2000 random-effect coefficients instead of 18")
    ap.add_argument("--no-leaf-large", action="store_true", help="skip the 32 M-row leg of the leaf-statistics kernel (HBM-bound size)")
    ap.add_argument("--no-mode0", action="store_true", help="skip the glmm mode 0 leg (one device pass per gradient evaluation)")
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
    q = 18 if not getattr(args, "scaled_levels", False) else 2 * max(5, args.n // 2000) + max(8, args.n // 2000)
    return {"workload": "%s, n=%d, %d trees, p_bart=9, K=2, q=%d, n_test=%d, "
                        "1 chain per GPU" % (what, args.n, args.trees, q, args.n),
            "n": args.n, "trees": args.trees, "chains_per_gpu": 1, "parallelism": "chains over GPUs, no data-path collective",
            "l2": "inputs larger than L2, no explicit flush: one sweep streams ~%d MB of distinct N-length arrays (BART: binned X, "
                  "residual, response, offset, fits, latents; GLMM: X, Z index / value streams, response, offset, residual; running "
                  "means) against 126 MB of L2, so every step's kernels start from HBM; inside k_sweep the chain's residuals and "
                  "predictors are then held on chip for the 200 tree steps by design" % int(round((9 + 8 * 19 + 12 + 8) * args.n / 1e6)),
            "adapt_sweeps": args.adapt}


def ref_config(args, procs):
    """The reference arm runs the GPU arm's configuration: the same chains (chains-per-gpu x gpus), one single-threaded process each."""
    cfg = workload_config(args)
    cpg = max(1, args.chains_per_gpu)
    cfg["chains_per_gpu"] = cpg
    cfg["workload"] = cfg["workload"].replace("1 chain per GPU", "%d chain(s) per GPU" % cpg)
    cfg["cpu_processes"] = procs
    return cfg


def make_problem(args):
    from stan4bart_b200.frontend import friedman_problem
    lv = {"n_g1": max(5, args.n // 2000), "n_g2": max(8, args.n // 2000)} if getattr(args, "scaled_levels", False) else {}
    pr = friedman_problem(args.n, binary=not args.continuous, seed=99, **lv)
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
    O.use_fast(True)                # the -O3 -march=native build of the same sources (oracle/Makefile); parity checks use the strict one
    O.lib()
    build = "-O3 -march=native" if O.fast_build_is_native() else "-O2 (strict build: the native one was compiled for another CPU)"
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
            "sample": "CPU restatement of the reference algorithm (oracle/, gcc %s, not the reference binary): same n=%d, %d trees, "
                      "1 chain on 1 thread, %d full sweeps after 1 warm-up sweep (setup %.1f s excluded); these are early, "
                      "un-adapted sweeps with few leapfrogs each, i.e. the CPU's best case" % (build, args.n, args.trees, k, t_create)}


# ------------------------------------------------------------------------------------------------
def _ref_worker(args_dict, seed, conn):
    try:
        import oracle_lib as O
        O.use_fast(True)
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

    import oracle_lib as O
    O.use_fast(True)
    O.lib()                 # loaded in this process too (the workers are forked from it): the CPU arm's native code is visible here
    build = "-O3 -march=native" if O.fast_build_is_native() else "-O2 (strict build)"
    chains = max(1, args.gpus) * max(1, args.chains_per_gpu)      # the same chains as the GPU arm's configuration, one process each
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
    sample = ("oracle port (CPU restatement of the reference algorithm, gcc %s, not the reference binary), %d chain(s) in %d "
              "single-threaded process(es), n=%d, %d trees, %d full sweeps timed after %d warm-up sweep "
              "(requested steps=%d warmup=%d bounded by --ref-budget-s)" % (build, procs, procs, args.n, args.trees, k, warm_done, args.steps, args.warmup))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": k, "warmup": warm_done,
            "ms_per_step": 1000.0 * dt / k, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": ref_config(args, procs),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
RANK_CORES = []          # this rank's logical CPUs, one per physical core first (chain threads are pinned to them in that order)


def pin_rank_to_cores(local_rank, local_world):
    """Every rank gets its own slice of the host's cores (its chain threads and their NUTS run there): eight ranks sharing one
    NUMA node's scheduler was what bent the 8-GPU curve.  Inside the slice the logical CPUs are ordered one per physical core
    first, so that two chain threads (one running NUTS, one spinning on a stream) do not share a core's two hyper-threads."""
    global RANK_CORES
    try:
        cores = sorted(os.sched_getaffinity(0))
        per = max(1, len(cores) // max(1, local_world))
        mine = cores[local_rank * per:(local_rank + 1) * per] or cores
        os.sched_setaffinity(0, mine)
        first, rest, seen = [], [], set()
        for c in mine:
            try:
                with open("/sys/devices/system/cpu/cpu%d/topology/core_id" % c) as f:
                    core = int(f.read())
                with open("/sys/devices/system/cpu/cpu%d/topology/physical_package_id" % c) as f:
                    core = (int(f.read()), core)
            except (OSError, ValueError):
                core = ("cpu", c)
            (rest if core in seen else first).append(c)
            seen.add(core)
        RANK_CORES = first + rest
        return len(mine)
    except (AttributeError, OSError):
        return None


class ChainThread(threading.Thread):
    """One chain = one host thread + one CUDA stream (the reference runs one worker process per chain, R/stan4bart_fit.R:495-542).
    Chains of one GPU overlap: while one chain's NUTS runs on the host, the other chain's sweep kernel has the device."""

    def __init__(self, index, device, make_sampler):
        super().__init__(daemon=True)
        self.index, self.device, self.make_sampler = index, device, make_sampler
        self.cmd, self.done = None, threading.Event()
        self.go = threading.Event()
        self.sampler = None
        self.stream = None
        self.error = None
        self.result = None

    def run(self):
        import torch

        from stan4bart_b200 import _lib
        try:
            if RANK_CORES:      # chain c of this rank -> its own physical core (the last logical CPU of the slice stays with the main thread)
                try:
                    os.sched_setaffinity(0, {RANK_CORES[self.index % max(1, len(RANK_CORES) - 1)]})
                except OSError:
                    pass
            torch.cuda.set_device(self.device)
            L = _lib.load()
            _lib.check(L.s4b_set_device(self.device))
            self.stream = torch.cuda.Stream()
            _lib.check(L.s4b_set_stream(self.stream.cuda_stream))
            self.sampler = self.make_sampler(self.index)
        except Exception as e:      # pragma: no cover
            self.error = e
        self.done.set()
        while True:
            self.go.wait()
            self.go.clear()
            if self.cmd is None:
                return
            try:
                self.result = self.cmd(self)
            except Exception as e:  # pragma: no cover
                self.error = e
            self.done.set()

    def submit(self, fn):
        self.done.clear()
        self.cmd = fn
        self.go.set()

    def wait(self):
        self.done.wait()
        if self.error is not None:
            raise self.error
        return self.result


def run_all(chains, fn):
    for c in chains:
        c.submit(fn)
    return [c.wait() for c in chains]


def timed_concurrent(torch, chains, fn):
    """Device time (CUDA events) from a common start to the completion of the last chain's work on its stream."""
    tstream = torch.cuda.Stream()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(tstream)
    for c in chains:
        c.stream.wait_event(ev0)

    def work(c):
        out = fn(c)
        e = torch.cuda.Event()
        e.record(c.stream)
        return out, e
    res = run_all(chains, work)
    for _, e in res:
        tstream.wait_event(e)
    ev1.record(tstream)
    ev1.synchronize()
    return float(ev0.elapsed_time(ev1)), [r for r, _ in res]


def parity_at_config(args, pr, sweeps=2):
    """The first Gibbs sweeps AT THE BENCHMARKED CONFIGURATION (same n, trees, kernel instantiation, GLMM mode), CUDA path against
    the strict build of the CPU oracle with the same seeds: Stan rows, fits, variable counts and trees after every sweep."""
    import oracle_lib as O
    from stan4bart_b200.sampler import Sampler
    from stan4bart_b200.structs import bart_config, stan_control
    O.use_fast(False)
    t0 = time.time()
    mk = lambda: bart_config(args.n, 9, n_test=args.n, num_trees=args.trees, is_binary=not args.continuous, seed=4711, weights=pr.get("weights"))
    extra = {"sigma_init": pr["sigma_init"], "bart_offset_init": pr["bart_offset_init"]} if args.continuous else {}
    kw = dict(warmup=10, iter_=20, keep_fits=False, **extra)
    o = O.OracleSampler(mk(), pr["y"], pr["x_bart"], pr["x_test"], pr["stan_data"], stan_control(seed=4712), **kw)
    g = Sampler(mk(), pr["y"], pr["x_bart"], pr["x_test"], pr["stan_data"], stan_control(seed=4712), **kw)
    rel = lambda a, b: float(np.max(np.abs(np.asarray(a) - np.asarray(b)) / (np.abs(np.asarray(a)) + 1.0))) if np.size(a) else 0.0
    worst = {"stan": 0.0, "train": 0.0, "test": 0.0, "tree_values": 0.0}
    status = "ok"
    for k in range(sweeps):
        ro, rg = o.run(1, True), g.run(1, True)
        worst["stan"] = max(worst["stan"], rel(ro["stan"], rg["stan"]))
        worst["train"] = max(worst["train"], rel(ro["bart"]["train"], rg["bart"]["train"]))
        worst["test"] = max(worst["test"], rel(ro["bart"]["test"], rg["bart"]["test"]))
        to, tg = o.bart().trees(), g.bart().trees()
        same_structure = np.array_equal(to["var"], tg["var"]) and np.array_equal(to["n"], tg["n"]) and np.array_equal(to["tree"], tg["tree"])
        if same_structure:
            worst["tree_values"] = max(worst["tree_values"], rel(to["value"], tg["value"]))
        if not same_structure or not np.array_equal(ro["bart"]["varcount"], rg["bart"]["varcount"]):
            status = "FAILED: tree structures / node counts / variable counts differ after sweep %d" % (k + 1)
            break
    tol = 1e-8
    if status == "ok" and max(worst.values()) > tol:
        status = "FAILED: %s differs by %.3e (tolerance %.0e)" % (max(worst, key=worst.get), max(worst.values()), tol)
    glmm_mode, sweep_mode = g.glmm().mode(), g.bart().sweep_mode()
    del o, g
    return {"status": status, "sweeps": sweeps, "tolerance": tol, "max_rel_err": worst, "bit_exact": "tree structures, per-node row counts, variable counts",
            "glmm_mode": glmm_mode, "bart_sweep_mode": sweep_mode, "seconds": time.time() - t0,
            "what": "CUDA path vs CPU oracle (strict build), same seeds, n=%d, %d trees: every sweep's Stan row, train / test fits, trees" % (args.n, args.trees)}


def run_ours(args):
    import torch
    import torch.distributed as dist

    from stan4bart_b200 import _lib
    from stan4bart_b200.sampler import Sampler
    from stan4bart_b200.structs import bart_config, stan_control

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    cores_per_rank = pin_rank_to_cores(local_rank, local_world)
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

    from stan4bart_b200.dist import barrier, chain_seed, max_over_ranks, min_over_ranks

    pr = make_problem(args)
    n, T = args.n, args.trees
    sharded = bool(args.shard_rows) and world > 1
    shard_ctx = None
    # every chain needs a host core for its NUTS thread (one is left to the main thread)
    cpg = 1 if sharded else max(1, min(args.chains_per_gpu, max(1, (cores_per_rank or 4) - 1)))
    chains_total = 1 if sharded else world * cpg
    if sharded:
        # one chain, rows dealt out in contiguous blocks; every rank runs the replicated controller with the same seeds
        from stan4bart_b200.frontend import shard_problem
        from stan4bart_b200.shard import ShardContext, row_range
        shard_ctx = ShardContext.from_torch_distributed(total_obs=n)
        lo, hi = row_range(n, rank, world)
        pr = shard_problem(pr, lo, hi)
        n = hi - lo
    sd = pr["stan_data"]
    extra = {"sigma_init": pr["sigma_init"], "bart_offset_init": pr["bart_offset_init"]} if args.continuous else {}

    def make_sampler(c):
        chain = 0 if sharded else rank * cpg + c
        cfg = bart_config(n, 9, n_test=n, num_trees=T, is_binary=not args.continuous, seed=chain_seed(12345, chain), weights=pr.get("weights"))
        s = Sampler(cfg, pr["y"], pr["x_bart"], pr["x_test"], sd, stan_control(seed=chain_seed(1000, chain)), warmup=args.adapt,
                    iter_=args.adapt + args.steps, keep_fits=False, shard=shard_ctx, **extra)
        s.run(args.adapt, True, results=False)
        s.disengage_adaptation()
        return s

    W, K = max(3, args.warmup), args.steps
    if sharded:
        # the sharded constructors are collective over the ranks: keep them on this thread and stream
        class Inline:
            pass
        ch = Inline()
        ch.sampler, ch.stream = make_sampler(0), stream
        chains = [ch]

        def run_all_local(fn):
            return [fn(ch)]
    else:
        chains = [ChainThread(c, local_rank, make_sampler) for c in range(cpg)]
        for c in chains:
            c.start()
        for c in chains:
            c.wait()
        run_all_local = lambda fn: run_all(chains, fn)
    s0 = chains[0].sampler
    bart, glmm = s0.bart(), s0.glmm()
    names = sd.param_names()

    clocks = ClockSampler(local_rank) if rank == 0 else None          # one nvidia-smi poller per job, not one per rank
    if clocks:
        clocks.start()

    # ---- leg A: ONE chain alone on the GPU.  The sweep kernel's launch time (roofline) is measured here, where no other
    # chain's kernel can sit between its events ----
    s0.run(W, False, results=False)
    bart.tree_step_ms(reset=True)
    passes_before_a = glmm.num_device_passes()
    barrier()
    evA0, evA1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    evA0.record(chains[0].stream)
    if sharded:
        outA = s0.run(K, False, results=False)
    else:
        chains[0].submit(lambda c: c.sampler.run(K, False, results=False))
        outA = chains[0].wait()
    evA1.record(chains[0].stream)
    evA1.synchronize()
    barrier()
    ms_one = max_over_ranks(float(evA0.elapsed_time(evA1)))
    sweep_ms = bart.tree_step_ms(reset=True)
    statsA = s0.last_run_stats()
    value_one = (1 if sharded else world) * K / (ms_one / 1000.0)

    # ---- leg B (the headline when chains_per_gpu > 1): every chain of this GPU at once ----
    passes0 = sum(c.sampler.glmm().num_device_passes() for c in chains)
    if cpg > 1:
        run_all_local(lambda c: c.sampler.run(W, False, results=False))
        barrier()
        t_wall0 = time.time()
        ms_loc, outs = timed_concurrent(torch, chains, lambda c: c.sampler.run(K, False, results=False))
        barrier()
        t_wall = time.time() - t_wall0
        ms = max_over_ranks(ms_loc)
        stats = [c.sampler.last_run_stats() for c in chains]
        out = outs[0]
    else:
        ms, out, stats, t_wall = ms_one, outA, [statsA], ms_one / 1000.0
        passes0 = passes_before_a       # leg A is the timed region
    glmm_passes = sum(c.sampler.glmm().num_device_passes() for c in chains) - passes0
    n_leapfrog = float(out["stan"][names.index("n_leapfrog__")][-1])
    value = chains_total * K / (ms / 1000.0)
    ms_stan = float(np.mean([st["ms_stan"] for st in stats])) / K
    ms_bart = float(np.mean([st["ms_bart"] for st in stats])) / K
    grad_evals = float(np.mean([st["grad_evals"] for st in stats])) / K

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
    traffic, traffic_file = None, {}
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic_file = json.load(f)
            traffic = traffic_file.get("k_sweep_dram_bytes_per_launch" if persistent else "k_tree_step_dram_bytes_per_launch")
    except Exception:
        pass
    # the metric's second half: "leaf-stat HBM GB/s vs peak" -- the stand-alone leaf-statistics pass (one launch streams the
    # residuals and the binned predictor columns of the tree's rules: 11 B per observation in the SURVEY model)
    leaf_stat = None
    if not sharded:
        try:
            reps = 50
            leaf_ms = float(np.median([bart.time_leaf_stats(t, reps) for t in (0, T // 2, T - 1)]))
            leaf_gbs = 11.0 * n / (leaf_ms * 1e-3) / 1e9
            leaf_stat = {"kernel": "k_leaf_stats / k_leaf_stats_small (csrc/leaf_stats.cuh): one launch = one tree's per-leaf (n, sum r, sum r^2) over all rows",
                         "algorithmic_bytes_per_obs": 11.0, "launch_us": leaf_ms * 1e3, "achieved": leaf_gbs, "peak": peak, "unit": "GB/s",
                         "frac": leaf_gbs / peak, "traffic": traffic_file.get("k_leaf_stats_dram_bytes_per_launch"),
                         "timing": "CUDA events around %d back-to-back launches on the launching stream, median over 3 trees; at n = 1 M the "
                                   "pass reads ~10 MB (1.5 us at peak) and is bound by ~10 us of fixed cost (launch, tree set-up, two-level "
                                   "reduction), and the working set stays in L2 between launches; bandwidth-bound sizes (16 M - 32 M rows: "
                                   "see hbm_bound_size) are in profiles/leaf_stat_sizes_r2b.json" % reps}
        except Exception as e:      # pragma: no cover
            leaf_stat = {"error": repr(e)}
        # the same kernel where it IS bound by HBM: 32 M rows (working set 8 B residual + 1 B per rule and row = 290-320 MB, beyond the
        # 126 MB L2, so back-to-back launches stream from HBM); rank 0 of a one-GPU run only
        if leaf_stat is not None and "error" not in leaf_stat and world == 1 and not args.no_leaf_large:
            try:
                from stan4bart_b200.sampler import GpuBart
                nl = 32_000_000
                rg = np.random.default_rng(1)
                xl = np.asfortranarray(rg.random((nl, 3)))
                yl = 10 * np.sin(np.pi * xl[:, 0] * xl[:, 1]) + 5 * xl[:, 2] + rg.standard_normal(nl)
                gl = GpuBart(bart_config(nl, 3, num_trees=8, seed=3), yl, xl)
                gl.set_sigma(1.0)
                gl.sample_trees_from_prior()
                gl.run()
                tr = gl.trees()
                rows = []
                for t in range(8):
                    rules = int(np.sum(tr["var"][tr["tree"] == t] >= 0))
                    if 1 <= rules <= 7:
                        ms_l = gl.time_leaf_stats(t, 20)
                        rows.append((rules, ms_l))
                del gl, xl, yl
                if rows:
                    rules, ms_l = rows[len(rows) // 2]
                    by_rules = {}
                    for rr, mm in rows:
                        by_rules.setdefault(rr, []).append(mm)
                    leaf_stat["hbm_bound_size"] = {"n": nl, "rules_of_the_tree": rules, "launch_us": ms_l * 1e3,
                                                   "by_rules": {str(rr): {"trees": len(v), "launch_us": float(np.median(v)) * 1e3,
                                                                          "frac": 11.0 * nl / float(np.median(v)) / 1e6 / peak,
                                                                          "frac_bytes_actually_read": (8.0 + rr) * nl / float(np.median(v)) / 1e6 / peak,
                                                                          "kernel": "k_leaf_stats_small<2,4,1>" if rr <= 1 else "k_leaf_stats_small<4,3,2>" if rr <= 3 else "k_leaf_stats_small<8,2,2>"}
                                                                for rr, v in sorted(by_rules.items())},
                                                   "achieved": 11.0 * nl / ms_l / 1e6, "frac": 11.0 * nl / ms_l / 1e6 / peak,
                                                   "achieved_bytes_actually_read": (8.0 + rules) * nl / ms_l / 1e6,
                                                   "frac_bytes_actually_read": (8.0 + rules) * nl / ms_l / 1e6 / peak,
                                                   "note": "the leaf-statistics pass on a synthetic 32 M-row fit (3 predictors), the tree in the middle of the 8 prior trees "
                                                           "(by_rules: every tree shape of the fit): 11 algorithmic B per row as above; the kernels actually read "
                                                           "8 + (rules of the tree) B per row; trees with <= 2 / <= 4 bottom nodes take the software-pipelined "
                                                           "kernels with register bins / register counts (csrc/leaf_stats.cuh)"}
            except Exception as e:      # pragma: no cover
                leaf_stat["hbm_bound_size"] = {"error": repr(e)}
    # the GLMM data pass (S, X'e, Z'e over all rows: 44 algorithmic B per row for this model), L2-warm and from HBM
    glmm_pass = None
    if not sharded:
        try:
            gm = glmm
            bpr = 8.0 + 8.0 * sd.K + 4.0 * 3 + 8.0 * 1
            warm_ms, is_bulk = gm.time_data_pass(30, False)
            cold_ms, _ = gm.time_data_pass(20, True)
            glmm_pass = {"kernel": "k_glmm_data_terms_bulk (cp.async.bulk + mbarrier ring)" if is_bulk else "k_glmm_data_terms", "algorithmic_bytes_per_obs": bpr,
                         "launch_us_l2_warm": warm_ms * 1e3, "launch_us_l2_flushed": cold_ms * 1e3,
                         "achieved_l2_warm": bpr * n / warm_ms / 1e6, "achieved": bpr * n / cold_ms / 1e6, "peak": peak, "unit": "GB/s",
                         "frac": bpr * n / cold_ms / 1e6 / peak, "frac_l2_warm": bpr * n / warm_ms / 1e6 / peak,
                         "traffic": traffic_file.get("k_glmm_data_terms_dram_bytes_per_launch"),
                         "timing": "CUDA events around single launches (L2 flushed before each by a 256 MB memset followed by a 256 MB read pass, so that the write-back of the memset's dirty lines is not charged to the launch) and around 30 back-to-back launches; "
                                   "44 MB is 7 us at peak, ~8 us of the launch are fixed cost (profiles/glmm_pass_r2b.json: 9 us for 70 000 rows; 0.68 / 0.73 of peak at 4 M rows)"}
        except Exception as e:      # pragma: no cover
            glmm_pass = {"error": repr(e)}
    roofline = {"bound": "hbm", "kernel": ("k_sweep_pipe + k_sweep (one sweep = one %d-tree pass over all rows: the pipelined kernel takes the steps that fit it, "
                                           "the synchronous kernel the rest)" % T) if persistent else "k_tree_step (one launch = one tree)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_tree_obs": bytes_per_obs,
                "units_per_launch": units_per_launch, "avg_tree_levels": lv,
                "launch_us": launch_ms * 1e3, "tree_step_us": sweep_ms / (K * T) * 1e3,
                "launch_timing": "CUDA events on the launching stream around every sweep launch of the one-chain leg (%d sweeps)" % K,
                "achieved_lean_model_gbs": bytes_per_obs_lean * units_per_launch / (launch_ms * 1e-3) / 1e9,
                "lean_bytes_per_tree_obs": bytes_per_obs_lean,
                "whole_sweep_frac": bytes_per_obs * T * n * (value / max(1, world)) / 1e9 / peak,
                "leaf_stat": leaf_stat, "glmm_data_pass": glmm_pass,
                # where a sweep's time goes, per rank (max / min over ranks): kept here so that the scaling record can attribute a loss
                "step_breakdown": {"ms_stan_block_max": max_over_ranks(ms_stan), "ms_stan_block_min": min_over_ranks(ms_stan),
                                   "ms_bart_block_max": max_over_ranks(ms_bart), "ms_bart_block_min": min_over_ranks(ms_bart),
                                   "ms_per_step_max": max_over_ranks(ms / K), "ms_per_step_min": min_over_ranks(ms / K),
                                   "grad_evals_per_sweep": grad_evals, "host_cores_per_rank": cores_per_rank},
                "note": "algorithmic bytes are what per-tree streaming passes must move (SURVEY.md 8d); the persistent kernel keeps residuals "
                        "in registers and predictors in shared memory, so DRAM traffic per launch is ~17 MB and the binding limit is the "
                        "latency of 200 sequential reduce + Metropolis decisions, not HBM (DESIGN.md section 4)"}

    # ---- the literal north-star Stan path beside it: one device pass per gradient evaluation (glmm mode 0) ----
    value_mode0 = None
    if not sharded and not args.no_mode0:
        k0 = max(2, min(4, K))
        for c in chains:
            c.sampler.glmm().set_mode(0)
        run_all_local(lambda c: c.sampler.run(1, False, results=False))
        ms0, _ = timed_concurrent(torch, chains, lambda c: c.sampler.run(k0, False, results=False))
        for c in chains:
            c.sampler.glmm().set_mode(1)
        value_mode0 = {"value": chains_total * k0 / (max_over_ranks(ms0) / 1000.0), "unit": UNIT, "steps": k0,
                       "note": "glmm_mode 0: k_glmm_data_terms launched for EVERY gradient evaluation (~%d per sweep) instead of once per "
                               "sweep; same chain, same draws to rounding" % int(round(grad_evals))}

    # ---- end-to-end leg through the C ABI with host buffers ----
    pins = []
    for c in chains:
        h2d, d2h = c.sampler.set_host_plumbing(True)
        pins.append((torch.empty(n, dtype=torch.float64).pin_memory(), torch.empty(n, dtype=torch.float64).pin_memory(),
                     torch.empty(c.sampler.num_pars, dtype=torch.float64).pin_memory()))
    for c, pin in zip(chains, pins):
        c.pin = pin
    e2e_run = lambda k: (lambda c: c.sampler.run_into(k, False, stan=c.pin[2].data_ptr(), train=c.pin[0].data_ptr(), test=c.pin[1].data_ptr()))
    run_all_local(e2e_run(W))
    barrier()
    t0 = time.time()
    run_all_local(e2e_run(K))
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.time() - t0)
    for c in chains:
        c.sampler.set_host_plumbing(False)
    if clocks:
        clocks.stop()
    e2e = {"value": chains_total * K / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d) * cpg, "d2h_bytes_per_step": int(d2h + 8 * n + 8 * s0.num_pars) * cpg,
           "note": "s4b_sampler_run with host result buffers (train + test fits, Stan row) and every N-vector of the sweep "
                   "(parametric mean, BART fit, latents) round-tripped through pinned host memory like the reference's host vectors "
                   "(the training fit makes its round trip through the result buffer, as in init.cpp:828-835: counted once); "
                   "bytes are per step of the GPU (all %d chain(s) on it advance one sweep)" % cpg}

    # kernels launched inside the timed region, per sweep and chain.  BART block with the pipelined sweep: k_prepare_sweep +
    # 2 x (k_sweep_pipe + k_sweep) segment launches (those with nothing to do return at once) + epilogue + epoch bump;
    # synchronous sweep: k_prepare_sweep + k_sweep + epilogue + epoch bump; per-tree kernels: T + 3.  Then the offset kernel + its
    # epoch bump (1 kernel for a continuous response), parametric mean, the fused GLMM input refresh, the fused running-mean
    # accumulation, plus the GLMM data passes.
    pl = bart.pipeline()
    if bart.sweep_mode() == 2:
        per_sweep_bart = (1 + 4 + 2) if pl["enabled"] else 4
    else:
        per_sweep_bart = T + 3
    launches = cpg * K * (per_sweep_bart + (1 if args.continuous else 2) + 1 + 1 + 1) + glmm_passes
    sweeps_done = max(1, pl["sweeps_offered"])
    roofline["pipelined_sweep"] = {"enabled": pl["enabled"], "tree_steps_pipelined_frac": pl["steps_pipelined"] / (sweeps_done * T) if pl["enabled"] else 0.0,
                                   "misfits": pl["misfits"],
                                   "note": "csrc/sweep_pipe.cuh: workers one tree step ahead of the controller; steps whose trees are too large for it "
                                           "run in the synchronous kernel (csrc/sweep_kernel.cuh)"}

    cfg_out = workload_config(args)
    cfg_out["chains_per_gpu"] = cpg
    cfg_out["workload"] = cfg_out["workload"].replace("1 chain per GPU", "%d chain(s) per GPU" % cpg)
    if sharded:
        cfg_out["sharding"] = f"one chain, rows sharded over {world} GPUs ({n} rows on rank 0)"
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K,
            "higher_is_better": True, "scaling": "strong" if sharded else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": cfg_out,
            "clocks": clocks.summary() if clocks else None, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roofline,
            "per_chain": {"chains_total": chains_total, "sweeps_per_s_per_chain": value / chains_total,
                          "value_one_chain_per_gpu": value_one, "ms_per_sweep_one_chain_per_gpu": ms_one / K,
                          "note": "a step of the job = every chain advances one Gibbs sweep; with %d chains per GPU one chain's host-side NUTS "
                                  "overlaps the other chain's sweep kernel (the reference also runs its chains side by side, one process each)" % cpg},
            "value_mode0": value_mode0,
            "breakdown": {"ms_stan_block": ms_stan, "ms_bart_block": ms_bart, "grad_evals_per_sweep": grad_evals,
                          "glmm_device_passes_per_sweep": glmm_passes / (K * cpg), "glmm_mode": glmm.mode(), "bart_sweep_mode": bart.sweep_mode(),
                          "n_leapfrog_last": n_leapfrog, "wall_s": t_wall, "tree_step_us": sweep_ms / (K * T) * 1e3}}
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            line["parity_at_config"] = parity_at_config(args, pr)
            line["cpu_baseline"] = cpu_baseline(args, pr, args.cpu_budget_s)
        else:
            line["cpu_baseline"] = None
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
    if not sharded:
        for c in chains:
            c.cmd = None
            c.go.set()
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
