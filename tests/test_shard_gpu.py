"""Observation-sharded chains on 2 GPUs: one process per GPU, rows split in contiguous blocks,
per-tree statistics and GLMM reductions exchanged through peer-mapped mailboxes inside the kernels.  Every rank must
reproduce the CPU oracle run on the WHOLE data set (integer fields exact, floating fields to the usual tolerance;
the only difference to the single-GPU path is the order of the final sums).  On a box with one GPU both ranks run on that
device as two time-sliced processes (the peer mailboxes are still mapped through CUDA IPC and every exchange still crosses
the process boundary), so the suite never skips; `gpurun --gpus 2 -- python -m pytest tests/test_shard_gpu.py -m gpu` runs it
over NVLink."""
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

import oracle_lib as O
import shard_cases as SC
from common import compare_traces, rel_err
from stan4bart_b200.frontend import friedman_problem
from stan4bart_b200.shard import row_range
from stan4bart_b200.structs import bart_config, stan_control

pytestmark = pytest.mark.gpu
WORLD = 2
HERE = os.path.dirname(os.path.abspath(__file__))


def _device_count():
    from stan4bart_b200 import _lib
    return _lib.load().s4b_device_count()


@pytest.fixture(scope="module")
def ranks():
    if _device_count() < 1:
        pytest.skip("needs a CUDA device")
    with tempfile.TemporaryDirectory() as tmp:
        out = os.path.join(tmp, "shard")
        port = 29500 + os.getpid() % 2000
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={WORLD}", "--master-addr", "127.0.0.1",
               "--master-port", str(port), os.path.join(HERE, "shard_worker.py"), "--out", out]
        # the kernels give up on a silent peer after 30 s; the outer limit only guards the launch itself
        p = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
        assert p.returncode == 0, p.stdout[-4000:] + "\n" + p.stderr[-4000:]
        yield [dict(np.load(f"{out}.rank{r}.npz")) for r in range(WORLD)]


def _cat(ranks, key, axis=0):
    return np.concatenate([r[key] for r in ranks], axis=axis)


def test_small_allreduce_is_rank_ordered_and_identical(ranks):
    v = [SC.allreduce_input(r) for r in range(WORLD)]
    want = v[0].copy()
    for r in range(1, WORLD):
        want = want + v[r]
    for r in ranks:
        assert np.array_equal(r["allreduce_sum"], want)                    # same order of additions: bit-exact
        assert np.array_equal(r["allreduce_max"], np.maximum.reduce(v))
    long_want = SC.allreduce_long_input(0)
    for r in range(1, WORLD):
        long_want = long_want + SC.allreduce_long_input(r)
    for r in ranks:
        assert np.array_equal(r["allreduce_long"], long_want)              # chunked path (> 1024 entries)


@pytest.mark.parametrize("binary", [False, True])
def test_sharded_bart_matches_whole_data_oracle(ranks, binary):
    tag = "bin" if binary else "cont"
    x, y, off = SC.bart_data(binary)
    cfg = bart_config(len(y), x.shape[1], n_test=0, num_trees=SC.BART_TREES, is_binary=binary, seed=SC.BART_SEED)
    o = O.OracleBart(cfg, y, x)
    o.set_offset(off, True)
    if not binary:
        o.set_sigma(1.3)
    o.sample_trees_from_prior()
    o.set_trace(SC.BART_TREES * SC.BART_SWEEPS)
    for _ in range(SC.BART_SWEEPS):
        ro = o.run()
    tr_o = o.trace()
    for r in ranks:
        compare_traces(tr_o, r[f"bart_{tag}_trace"], tol=1e-8)
        assert rel_err(o.data_range(), r[f"bart_{tag}_range"]) <= 1e-12
    assert np.array_equal(ranks[0][f"bart_{tag}_trace"], ranks[1][f"bart_{tag}_trace"])     # ranks agree bit for bit
    for r in ranks:          # streamed sweep variant, sharded: same arithmetic as the register variant
        assert np.array_equal(r[f"bart_{tag}_trace"], r[f"bart_{tag}_trace_streamed"])
    assert np.array_equal(_cat(ranks, f"bart_{tag}_train"), _cat(ranks, f"bart_{tag}_train_streamed"))
    for mode in (0, 1):      # streaming per-tree kernels, sharded
        for r in ranks:
            compare_traces(tr_o, r[f"bart_{tag}_trace_mode{mode}"], tol=1e-8)
        got = _cat(ranks, f"bart_{tag}_train_mode{mode}")
        assert rel_err(ro["train"], got, scale=np.abs(ro["train"]) + 1.0) <= 1e-8
    train = _cat(ranks, f"bart_{tag}_train")
    assert rel_err(ro["train"], train, scale=np.abs(ro["train"]) + 1.0) <= 1e-8
    res = _cat(ranks, f"bart_{tag}_residual")
    assert rel_err(o.residual(), res, scale=np.abs(o.residual()) + 1.0) <= 1e-8
    # leaf statistics of tree 0 are those of the whole data set, on every rank
    heap, cnt, s, ss = o.leaf_stats(0)
    for r in ranks:
        ls = r[f"bart_{tag}_leafstats"]
        assert np.array_equal(ls[:, 0].astype(np.int64), heap) and np.array_equal(ls[:, 1].astype(np.int64), cnt)
        assert rel_err(s, ls[:, 2], scale=np.abs(s) + 1e-3 * cnt) <= 1e-10 and rel_err(ss, ls[:, 3]) <= 1e-10
    to = o.trees()
    for r in ranks:
        assert np.array_equal(to["var"], r[f"bart_{tag}_trees_var"])
        assert np.array_equal(to["n"], r[f"bart_{tag}_trees_n"])            # counts of the whole data set
        assert rel_err(to["value"], r[f"bart_{tag}_trees_value"], scale=np.abs(to["value"]) + 1e-3) <= 1e-8


def test_sharded_quantile_cut_points_match_whole_data_oracle(ranks):
    """bart_args use.quantiles on sharded rows: the cut points are a function of the distinct values of the WHOLE column, so the ranks
    exchange their sorted distinct values at setup.  Cut values bit for bit those of the whole-data oracle - also for values that several
    shards hold, for a predictor that is constant inside one shard only, and for one that is constant everywhere (no cut, never chosen)."""
    x, y, off = SC.quantile_bart_data()
    cfg = bart_config(len(y), x.shape[1], n_test=0, num_trees=SC.BART_TREES, seed=SC.BART_SEED, n_cuts=40, use_quantiles=True)
    o = O.OracleBart(cfg, y, x)
    o.set_offset(off, True)
    o.set_sigma(1.3)
    o.sample_trees_from_prior()
    o.set_trace(SC.BART_TREES * SC.BART_SWEEPS)
    for _ in range(SC.BART_SWEEPS):
        ro = o.run()
    to = o.trees()
    rules = to["var"] >= 0
    assert rules.any() and ro["varcount"][4] == 0
    for r in ranks:
        compare_traces(o.trace(), r["bart_quant_trace"], tol=1e-8)
        assert np.array_equal(to["var"], r["bart_quant_trees_var"]) and np.array_equal(to["n"], r["bart_quant_trees_n"])
        assert np.array_equal(to["value"][rules], r["bart_quant_trees_value"][rules])                   # cut values bit for bit
        assert np.array_equal(ro["varcount"], r["bart_quant_varcount"])
    assert np.array_equal(ranks[0]["bart_quant_trace"], ranks[1]["bart_quant_trace"])
    train = _cat(ranks, "bart_quant_train")
    assert rel_err(ro["train"], train, scale=np.abs(ro["train"]) + 1.0) <= 1e-8


@pytest.mark.parametrize("mode", [0, 1])
def test_sharded_glmm_density_matches_whole_data_oracle(ranks, mode):
    pr = friedman_problem(SC.GLMM_N)
    g = O.OracleGlmm(pr["stan_data"])
    g.set_offset(SC.glmm_offset())
    for k, q in enumerate(SC.glmm_points(g.d)):
        lp, grad, st = g.log_prob_grad(q)
        for r in ranks:
            row = r[f"glmm_mode{mode}"][k]
            assert int(row[1]) == st
            assert abs(row[0] - lp) <= 1e-10 * max(1.0, abs(lp))
            assert rel_err(grad, row[2:], scale=np.abs(grad) + 1.0) <= 1e-10
    assert np.array_equal(ranks[0][f"glmm_mode{mode}"], ranks[1][f"glmm_mode{mode}"])


@pytest.mark.parametrize("binary", [False, True])
def test_sharded_gibbs_matches_whole_data_oracle(ranks, binary):
    tag = "bin" if binary else "cont"
    n = SC.GIBBS_N
    pr = friedman_problem(n, binary=binary)
    cfg = bart_config(n, 9, n_test=n, num_trees=SC.GIBBS_TREES, is_binary=binary, seed=SC.GIBBS_SEED)
    ctl = stan_control(seed=SC.GIBBS_SEED + 1)
    o = O.OracleSampler(cfg, pr["y"], pr["x_bart"], pr["x_test"], pr["stan_data"], ctl, warmup=SC.GIBBS_WARMUP, iter_=SC.GIBBS_ITER,
                        keep_fits=True, sigma_init=pr["sigma_init"], bart_offset_init=pr["bart_offset_init"])
    ob = o.bart()
    ob.set_trace(SC.GIBBS_TREES * (SC.GIBBS_WARMUP + SC.GIBBS_SAMPLES))
    w = o.run(SC.GIBBS_WARMUP, True)
    o.disengage_adaptation()
    s = o.run(SC.GIBBS_SAMPLES, False)
    tr_o = ob.trace()
    for r in ranks:
        compare_traces(tr_o, r[f"gibbs_{tag}_trace"], tol=1e-7)
        for nm, part in (("w", w), ("r", s)):
            assert rel_err(part["stan"], r[f"gibbs_{tag}_{nm}_stan"], scale=np.abs(part["stan"]) + 1.0) <= 1e-7
            assert np.array_equal(part["bart"]["varcount"], r[f"gibbs_{tag}_{nm}_varcount"])
            assert rel_err(part["bart"]["sigma"], r[f"gibbs_{tag}_{nm}_sigma"]) <= 1e-7
        assert rel_err(o.data_range(), r[f"gibbs_{tag}_range"]) <= 1e-10
    # replicated state is bitwise identical on the two ranks
    for key in (f"gibbs_{tag}_trace", f"gibbs_{tag}_w_stan", f"gibbs_{tag}_r_stan"):
        assert np.array_equal(ranks[0][key], ranks[1][key]), key
    for nm, part in (("w", w), ("r", s)):
        for which in ("train", "test"):
            got = _cat(ranks, f"gibbs_{tag}_{nm}_{which}", axis=0)
            want = part["bart"][which]
            assert rel_err(want, got, scale=np.abs(want) + 1.0) <= 1e-7
    lo, hi = row_range(n, 1, WORLD)
    assert ranks[1][f"gibbs_{tag}_parmean"].shape == (hi - lo,)
    got = _cat(ranks, f"gibbs_{tag}_mean_train")
    assert rel_err(s["bart"]["train"].mean(axis=1), got, scale=1.0) <= 1e-7


def test_sharded_weighted_gibbs_matches_whole_data_oracle(ranks):
    n = SC.GIBBS_N
    pr = friedman_problem(n)
    wt = SC.gibbs_weights()
    pr["stan_data"].weights = wt
    cfg = bart_config(n, 9, n_test=n, num_trees=SC.GIBBS_TREES, seed=SC.GIBBS_SEED, weights=wt)
    ctl = stan_control(seed=SC.GIBBS_SEED + 1)
    o = O.OracleSampler(cfg, pr["y"], pr["x_bart"], pr["x_test"], pr["stan_data"], ctl, warmup=SC.GIBBS_WARMUP, iter_=SC.GIBBS_ITER,
                        keep_fits=True, sigma_init=pr["sigma_init"], bart_offset_init=pr["bart_offset_init"])
    ob = o.bart()
    ob.set_trace(SC.GIBBS_TREES * SC.GIBBS_WARMUP)
    w = o.run(SC.GIBBS_WARMUP, True)
    for r in ranks:
        compare_traces(ob.trace(), r["gibbs_wt_trace"], tol=1e-7, ll_difference_only=True)
        assert rel_err(w["stan"], r["gibbs_wt_stan"], scale=np.abs(w["stan"]) + 1.0) <= 1e-7
    assert np.array_equal(ranks[0]["gibbs_wt_trace"], ranks[1]["gibbs_wt_trace"])
    got = _cat(ranks, "gibbs_wt_train", axis=0)
    assert rel_err(w["bart"]["train"], got, scale=np.abs(w["bart"]["train"]) + 1.0) <= 1e-7


def test_nccl_reference_collective_agrees_with_the_mailbox_exchange(ranks):
    """`ncclAllReduce` as the reference implementation of the exchange.  With one GPU per rank the worker repeats the
    small all-reduces and the sharded GLMM density through NCCL (`s4b_shard_use_nccl`); both transports must give the same numbers
    (two ranks: a + b either way round, so bit for bit) and the oracle's density.  On a one-GPU box NCCL cannot place two ranks on
    the same device; the mailbox path (the product path) is what the other tests of this file exercise there."""
    if not all(r["nccl"][0] == 1.0 for r in ranks):
        # one GPU: NCCL refuses two ranks on one device, so the cross-check runs on a communicator of ONE rank in this process -- the
        # library is loaded, the communicator built and ncclAllReduce launched on the fit's stream exactly as with more ranks (the
        # two-rank comparison is in profiles/shard_and_nuts_tests_2gpu_r2.log)
        from stan4bart_b200.sampler import GlmmModel
        from stan4bart_b200.shard import ShardContext
        ctx = ShardContext(0, 1)
        ctx.init_nccl()
        v, long_v = SC.allreduce_input(0), SC.allreduce_long_input(0)
        plain = [ctx.allreduce(v, "sum"), ctx.allreduce(v, "max"), ctx.allreduce(long_v, "sum")]
        ctx.use_nccl(True)
        through_nccl = [ctx.allreduce(v, "sum"), ctx.allreduce(v, "max"), ctx.allreduce(long_v, "sum")]
        for a, b, want in zip(plain, through_nccl, (v, v, long_v)):
            assert np.array_equal(a, want) and np.array_equal(b, want)
        pr = friedman_problem(SC.GLMM_N)
        ctx.set_obs_range(0, SC.GLMM_N)
        m, g = GlmmModel(pr["stan_data"], shard=ctx), O.OracleGlmm(pr["stan_data"])     # every data pass all-reduced by NCCL
        m.set_mode(0); m.set_offset(SC.glmm_offset()); g.set_offset(SC.glmm_offset())
        for q in SC.glmm_points(g.d):
            lp, grad, st = g.log_prob_grad(q)
            lp2, grad2, st2 = m.log_prob_grad(q)
            assert st == st2 and abs(lp - lp2) <= 1e-10 * max(1.0, abs(lp)) and rel_err(grad, grad2, scale=np.abs(grad) + 1.0) <= 1e-10
        del m
        ctx.use_nccl(False)
        return
    for r in ranks:
        assert np.array_equal(r["nccl_allreduce_sum"], r["allreduce_sum"])
        assert np.array_equal(r["nccl_allreduce_max"], r["allreduce_max"])
        assert np.array_equal(r["nccl_allreduce_long"], r["allreduce_long"])
        assert np.allclose(r["glmm_nccl"], r["glmm_mode0"], rtol=1e-13, atol=0.0)
    pr = friedman_problem(SC.GLMM_N)
    g = O.OracleGlmm(pr["stan_data"])
    g.set_offset(SC.glmm_offset())
    for k, q in enumerate(SC.glmm_points(g.d)):
        lp, grad, st = g.log_prob_grad(q)
        for r in ranks:
            row = r["glmm_nccl"][k]
            assert int(row[1]) == st and abs(row[0] - lp) <= 1e-10 * max(1.0, abs(lp))
            assert rel_err(grad, row[2:], scale=np.abs(grad) + 1.0) <= 1e-10
