"""Debug: one rank of a 2-rank sharded continuous Gibbs chain, compared stage by stage with the whole-data oracle."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch.distributed as dist
import oracle_lib as O
import shard_cases as SC
from stan4bart_b200 import _lib
from stan4bart_b200.frontend import friedman_problem, shard_problem
from stan4bart_b200.sampler import Sampler
from stan4bart_b200.shard import ShardContext, row_range
from stan4bart_b200.structs import bart_config, stan_control

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
ndev = _lib.load().s4b_device_count()
_lib.check(_lib.load().s4b_set_device(rank % ndev))
dist.init_process_group("gloo", rank=rank, world_size=world)
ctx = ShardContext.from_torch_distributed()
from stan4bart_b200.sampler import GlmmModel, GpuBart
pre = os.environ.get("DBG_PRE", "")
if "a" in pre:
    v = SC.allreduce_input(rank)
    ctx.allreduce(v, "sum"); ctx.allreduce(v, "max"); ctx.allreduce(SC.allreduce_long_input(rank), "sum")
if "b" in pre:
    for binary in (False, True):
        x, y, off = SC.bart_data(binary)
        n0 = len(y)
        lo, hi = row_range(n0, rank, world)
        ctx.set_obs_range(lo, n0)
        cfg = bart_config(hi - lo, x.shape[1], n_test=0, num_trees=SC.BART_TREES, is_binary=binary, seed=SC.BART_SEED)
        variants = os.environ.get("DBG_BVAR", "psmM")
        for var in variants:
            if var == "s": os.environ["S4B_FORCE_STREAM"] = "1"
            g = GpuBart(cfg, y[lo:hi], x[lo:hi], shard=ctx)
            if var == "s": del os.environ["S4B_FORCE_STREAM"]
            if var == "m": g.set_sweep_mode(0)
            if var == "M": g.set_sweep_mode(1)
            g.set_offset(off[lo:hi], True)
            if not binary: g.set_sigma(1.3)
            g.sample_trees_from_prior()
            g.set_trace(SC.BART_TREES * SC.BART_SWEEPS)
            for _ in range(SC.BART_SWEEPS): g.run()
            del g
if "g" in pre:
    pr0 = friedman_problem(SC.GLMM_N)
    lo, hi = row_range(SC.GLMM_N, rank, world)
    ctx.set_obs_range(lo, SC.GLMM_N)
    sp0 = shard_problem(pr0, lo, hi)
    for mode0 in (0, 1):
        m = GlmmModel(sp0["stan_data"], shard=ctx)
        m.set_mode(mode0)
        m.set_offset(SC.glmm_offset()[lo:hi])
        for q in SC.glmm_points(m.d): m.log_prob_grad(q)
        del m
n = SC.GIBBS_N
pr = friedman_problem(n)
lo, hi = row_range(n, rank, world)
ctx.set_obs_range(lo, n)
sp = shard_problem(pr, lo, hi)
mode = int(os.environ.get("DBG_GLMM_MODE", "1"))
cfg = bart_config(hi - lo, 9, n_test=hi - lo, num_trees=SC.GIBBS_TREES, seed=SC.GIBBS_SEED)
ctl = stan_control(seed=SC.GIBBS_SEED + 1)
s = Sampler(cfg, sp["y"], sp["x_bart"], sp["x_test"], sp["stan_data"], ctl, warmup=SC.GIBBS_WARMUP, iter_=SC.GIBBS_ITER,
            keep_fits=True, sigma_init=pr["sigma_init"], bart_offset_init=sp["bart_offset_init"], shard=ctx)
s.glmm().set_mode(mode)
cfgo = bart_config(n, 9, n_test=n, num_trees=SC.GIBBS_TREES, seed=SC.GIBBS_SEED)
o = O.OracleSampler(cfgo, pr["y"], pr["x_bart"], pr["x_test"], pr["stan_data"], ctl, warmup=SC.GIBBS_WARMUP, iter_=SC.GIBBS_ITER,
                    keep_fits=True, sigma_init=pr["sigma_init"], bart_offset_init=pr["bart_offset_init"])
b, ob = s.bart(), o.bart()
def rel(a, c): a = np.asarray(a, float); c = np.asarray(c, float); return float(np.max(np.abs(a - c) / (np.abs(a) + 1.0))) if a.size else 0.0
print(rank, "after create: range", ob.data_range(), b.data_range(), flush=True)
to, tg = ob.trees(), b.trees()
print(rank, "trees var equal", np.array_equal(to["var"], tg["var"]), "value rel", rel(to["value"], tg["value"]) if len(to["value"]) == len(tg["value"]) else "len differs", flush=True)
print(rank, "residual rel", rel(ob.residual()[lo:hi], b.residual()), flush=True)
# the GLMM density right after creation
g, og = s.glmm(), o.glmm() if hasattr(o, "glmm") else None
q = np.random.default_rng(1).uniform(-1, 1, g.d)
print(rank, "glmm lp/grad gpu", g.log_prob_grad(q)[0], flush=True)
iters = int(os.environ.get("DBG_ITERS", "1"))
if os.environ.get("DBG_TRACE"):
    b.set_trace(SC.GIBBS_TREES * 20); ob.set_trace(SC.GIBBS_TREES * 20)
w = o.run(iters, True)
r = s.run(iters, True)
names = sp["stan_data"].param_names()
print(rank, "sweep 1 stan rel", rel(w["stan"], r["stan"]), "lp", w["stan"][0, 0], r["stan"][0, 0], "aux", w["stan"][names.index("aux.1"), 0], r["stan"][names.index("aux.1"), 0], flush=True)
print(rank, "sweep 1 train rel", rel(w["bart"]["train"][lo:hi], r["bart"]["train"]), flush=True)
dist.barrier()
dist.destroy_process_group()
