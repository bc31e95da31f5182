"""One rank of the observation-sharded parity run (launched by test_shard_gpu.py / torch.distributed.run).

Runs every sharded case on this rank's rows and saves what it saw to <out>.rank<r>.npz; the test process compares the
pieces of all ranks with the CPU oracle run on the whole data set.  Uses gloo only to hand the IPC handles round."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch.distributed as dist  # noqa: E402

from stan4bart_b200 import _lib  # noqa: E402
from stan4bart_b200.frontend import friedman_problem, shard_problem  # noqa: E402
from stan4bart_b200.sampler import GlmmModel, GpuBart, Sampler  # noqa: E402
from stan4bart_b200.shard import ShardContext, row_range  # noqa: E402
from stan4bart_b200.structs import bart_config, stan_control  # noqa: E402

import shard_cases as SC  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", required=True)
    args = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    # on a box with fewer GPUs than ranks the ranks share devices: the processes are time-sliced on the GPU, the mailboxes are
    # mapped through CUDA IPC exactly as between two GPUs, so the exchange protocol is exercised unchanged (only slower)
    ndev = _lib.load().s4b_device_count()
    _lib.check(_lib.load().s4b_set_device(local % max(ndev, 1)))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    res = {}

    # ---- the small all-reduce ----
    ctx = ShardContext.from_torch_distributed()
    v = SC.allreduce_input(rank)
    res["allreduce_sum"] = ctx.allreduce(v, "sum")
    res["allreduce_max"] = ctx.allreduce(v, "max")
    long_v = SC.allreduce_long_input(rank)
    res["allreduce_long"] = ctx.allreduce(long_v, "sum")

    # ---- the reference collective: the same small all-reduces through ncclAllReduce (needs one GPU per rank) ----
    res["nccl"] = np.array([0.0])
    if ndev >= world:
        ctx.init_nccl()
        ctx.use_nccl(True)
        res["nccl"] = np.array([1.0])
        res["nccl_allreduce_sum"] = ctx.allreduce(v, "sum")
        res["nccl_allreduce_max"] = ctx.allreduce(v, "max")
        res["nccl_allreduce_long"] = ctx.allreduce(long_v, "sum")
        pr = friedman_problem(SC.GLMM_N)
        lo, hi = row_range(SC.GLMM_N, rank, world)
        ctx.set_obs_range(lo, SC.GLMM_N)
        sp = shard_problem(pr, lo, hi)
        m = GlmmModel(sp["stan_data"], shard=ctx)        # Gram matrix and every data pass all-reduced by NCCL
        m.set_mode(0)
        m.set_offset(SC.glmm_offset()[lo:hi])
        out = []
        for q in SC.glmm_points(m.d):
            lp, grad, st = m.log_prob_grad(q)
            out.append(np.concatenate([[lp, st], grad]))
        res["glmm_nccl"] = np.array(out)
        del m
        ctx.use_nccl(False)

    # ---- stand-alone BART on sharded rows ----
    for binary in (False, True):
        tag = "bin" if binary else "cont"
        x, y, off = SC.bart_data(binary)
        n = len(y)
        lo, hi = row_range(n, rank, world)
        ctx.set_obs_range(lo, n)
        cfg = bart_config(hi - lo, x.shape[1], n_test=0, num_trees=SC.BART_TREES, is_binary=binary, seed=SC.BART_SEED)
        g = GpuBart(cfg, y[lo:hi], x[lo:hi], shard=ctx)
        g.set_offset(off[lo:hi], True)
        if not binary:
            g.set_sigma(1.3)
        g.sample_trees_from_prior()
        g.set_trace(SC.BART_TREES * SC.BART_SWEEPS)
        for _ in range(SC.BART_SWEEPS):
            r = g.run()
        res[f"bart_{tag}_trace"] = g.trace()
        res[f"bart_{tag}_train"] = r["train"]
        res[f"bart_{tag}_range"] = np.array(g.data_range())
        res[f"bart_{tag}_leafstats"] = np.column_stack([np.asarray(a, dtype=np.float64) for a in g.leaf_stats(0)])
        res[f"bart_{tag}_residual"] = g.residual()
        tr = g.trees()
        res[f"bart_{tag}_trees_var"] = tr["var"]
        res[f"bart_{tag}_trees_n"] = tr["n"]
        res[f"bart_{tag}_trees_value"] = tr["value"]
        del g
        # the streamed sweep variant (shards beyond the register file), sharded
        os.environ["S4B_FORCE_STREAM"] = "1"
        g = GpuBart(cfg, y[lo:hi], x[lo:hi], shard=ctx)
        del os.environ["S4B_FORCE_STREAM"]
        g.set_offset(off[lo:hi], True)
        if not binary:
            g.set_sigma(1.3)
        g.sample_trees_from_prior()
        g.set_trace(SC.BART_TREES * SC.BART_SWEEPS)
        for _ in range(SC.BART_SWEEPS):
            r = g.run()
        res[f"bart_{tag}_trace_streamed"] = g.trace()
        res[f"bart_{tag}_train_streamed"] = r["train"]
        del g
        # the streaming per-tree kernels (shards too large for the on-chip sweep) exchange through the same mailboxes
        for mode in (0, 1):
            g = GpuBart(cfg, y[lo:hi], x[lo:hi], shard=ctx)
            g.set_sweep_mode(mode)
            g.set_offset(off[lo:hi], True)
            if not binary:
                g.set_sigma(1.3)
            g.sample_trees_from_prior()
            g.set_trace(SC.BART_TREES * SC.BART_SWEEPS)
            for _ in range(SC.BART_SWEEPS):
                r = g.run()
            res[f"bart_{tag}_trace_mode{mode}"] = g.trace()
            res[f"bart_{tag}_train_mode{mode}"] = r["train"]
            del g

    # ---- bart_args use.quantiles on sharded rows: the cut points come from the distinct values of the whole column ----
    x, y, off = SC.quantile_bart_data()
    n = len(y)
    lo, hi = row_range(n, rank, world)
    ctx.set_obs_range(lo, n)
    cfg = bart_config(hi - lo, x.shape[1], n_test=0, num_trees=SC.BART_TREES, seed=SC.BART_SEED, n_cuts=40, use_quantiles=True)
    g = GpuBart(cfg, y[lo:hi], x[lo:hi], shard=ctx)
    g.set_offset(off[lo:hi], True)
    g.set_sigma(1.3)
    g.sample_trees_from_prior()
    g.set_trace(SC.BART_TREES * SC.BART_SWEEPS)
    for _ in range(SC.BART_SWEEPS):
        r = g.run()
    res["bart_quant_trace"] = g.trace()
    res["bart_quant_train"] = r["train"]
    tr = g.trees()
    res["bart_quant_trees_var"] = tr["var"]
    res["bart_quant_trees_n"] = tr["n"]
    res["bart_quant_trees_value"] = tr["value"]
    res["bart_quant_varcount"] = r["varcount"]
    del g

    # ---- GLMM density on sharded rows ----
    pr = friedman_problem(SC.GLMM_N)
    lo, hi = row_range(SC.GLMM_N, rank, world)
    ctx.set_obs_range(lo, SC.GLMM_N)
    sp = shard_problem(pr, lo, hi)
    for mode in (0, 1):
        m = GlmmModel(sp["stan_data"], shard=ctx)
        m.set_mode(mode)
        m.set_offset(SC.glmm_offset()[lo:hi])
        out = []
        for q in SC.glmm_points(m.d):
            lp, grad, st = m.log_prob_grad(q)
            out.append(np.concatenate([[lp, st], grad]))
        res[f"glmm_mode{mode}"] = np.array(out)
        del m

    # ---- the whole Gibbs sweep on sharded rows ----
    for binary in (False, True):
        tag = "bin" if binary else "cont"
        pr = friedman_problem(SC.GIBBS_N, binary=binary)
        lo, hi = row_range(SC.GIBBS_N, rank, world)
        ctx.set_obs_range(lo, SC.GIBBS_N)
        sp = shard_problem(pr, lo, hi)
        cfg = bart_config(hi - lo, 9, n_test=hi - lo, num_trees=SC.GIBBS_TREES, is_binary=binary, seed=SC.GIBBS_SEED)
        ctl = stan_control(seed=SC.GIBBS_SEED + 1)
        s = Sampler(cfg, sp["y"], sp["x_bart"], sp["x_test"], sp["stan_data"], ctl, warmup=SC.GIBBS_WARMUP, iter_=SC.GIBBS_ITER,
                    keep_fits=True, sigma_init=pr["sigma_init"], bart_offset_init=sp["bart_offset_init"], shard=ctx)
        b = s.bart()
        b.set_trace(SC.GIBBS_TREES * (SC.GIBBS_WARMUP + SC.GIBBS_SAMPLES))
        w = s.run(SC.GIBBS_WARMUP, True)
        s.disengage_adaptation()
        r = s.run(SC.GIBBS_SAMPLES, False)
        res[f"gibbs_{tag}_trace"] = b.trace()
        for nm, part in (("w", w), ("r", r)):
            res[f"gibbs_{tag}_{nm}_stan"] = part["stan"]
            res[f"gibbs_{tag}_{nm}_train"] = part["bart"]["train"]
            res[f"gibbs_{tag}_{nm}_test"] = part["bart"]["test"]
            res[f"gibbs_{tag}_{nm}_varcount"] = part["bart"]["varcount"]
            res[f"gibbs_{tag}_{nm}_sigma"] = part["bart"]["sigma"]
        res[f"gibbs_{tag}_range"] = np.array(s.data_range())
        res[f"gibbs_{tag}_parmean"] = s.parametric_mean()
        m = s.means()
        res[f"gibbs_{tag}_mean_train"] = m["bart_train"]
        del b, s

    # ---- the same with observation weights (weighted leaf statistics cross the ranks as sum w r / sum w) ----
    pr = friedman_problem(SC.GIBBS_N)
    pr["stan_data"].weights = SC.gibbs_weights()
    lo, hi = row_range(SC.GIBBS_N, rank, world)
    ctx.set_obs_range(lo, SC.GIBBS_N)
    sp = shard_problem(pr, lo, hi)
    cfg = bart_config(hi - lo, 9, n_test=hi - lo, num_trees=SC.GIBBS_TREES, seed=SC.GIBBS_SEED, weights=SC.gibbs_weights()[lo:hi])
    ctl = stan_control(seed=SC.GIBBS_SEED + 1)
    s = Sampler(cfg, sp["y"], sp["x_bart"], sp["x_test"], sp["stan_data"], ctl, warmup=SC.GIBBS_WARMUP, iter_=SC.GIBBS_ITER,
                keep_fits=True, sigma_init=pr["sigma_init"], bart_offset_init=sp["bart_offset_init"], shard=ctx)
    b = s.bart()
    b.set_trace(SC.GIBBS_TREES * SC.GIBBS_WARMUP)
    w = s.run(SC.GIBBS_WARMUP, True)
    res["gibbs_wt_trace"] = b.trace()
    res["gibbs_wt_stan"] = w["stan"]
    res["gibbs_wt_train"] = w["bart"]["train"]
    del b, s

    np.savez(f"{args.out}.rank{rank}.npz", **res)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
