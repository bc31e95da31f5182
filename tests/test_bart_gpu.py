"""GPU parity tests for the BART half (CUDA path through the C ABI vs the CPU oracle)."""
import numpy as np
import pytest

import oracle_lib as O
from common import REL_TOL, bart_problem, compare_traces, rel_err
from stan4bart_b200.sampler import GpuBart
from stan4bart_b200.structs import bart_config

pytestmark = pytest.mark.gpu


def make_pair(n=500, p=5, n_test=0, binary=False, num_trees=10, seed=11, thin=1, offset=True, **kw):
    x, y, xt = bart_problem(n, p, n_test, binary)
    cfg = bart_config(n, p, n_test=n_test, num_trees=num_trees, is_binary=binary, seed=seed, thin=thin, **kw)
    o = O.OracleBart(cfg, y, x, xt)
    g = GpuBart(cfg, y, x, xt)
    if offset:
        off = 0.3 * x[:, 3] - 0.1
        o.set_offset(off, True); g.set_offset(off, True)
    if not binary:
        o.set_sigma(1.3); g.set_sigma(1.3)
    return o, g, (x, y, xt)


def assert_same_partition(o, g, num_trees):
    for t in range(num_trees):
        ao, ag = o.node_assignment(t), g.node_assignment(t)
        assert np.array_equal(ao, ag), f"tree {t}: node assignment differs for {np.count_nonzero(ao != ag)} observations"


def test_prior_trees_partition_and_residual():
    o, g, _ = make_pair(n=777, num_trees=12)
    o.sample_trees_from_prior(); g.sample_trees_from_prior()
    assert o.rng_counter() == g.rng_counter()
    assert_same_partition(o, g, 12)
    to, tg = o.trees(), g.trees()
    assert np.array_equal(to["var"], tg["var"]) and np.array_equal(to["tree"], tg["tree"])
    assert rel_err(to["value"], tg["value"]) <= REL_TOL
    assert rel_err(o.residual(), g.residual(), scale=1.0) <= 1e-12


def test_leaf_stats_kernel():
    o, g, _ = make_pair(n=2003, num_trees=8)
    o.sample_trees_from_prior(); g.sample_trees_from_prior()
    for t in range(8):
        ho, co, so, sso = o.leaf_stats(t)
        hg, cg, sg, ssg = g.leaf_stats(t)
        assert np.array_equal(ho, hg) and np.array_equal(co, cg)           # counts bit exact
        assert rel_err(sso, ssg) <= REL_TOL
        # sums can cancel: measure against sum |x| <= sqrt(n * sumsq)
        assert rel_err(so, sg, scale=np.sqrt(np.maximum(co, 1) * sso) + 1e-300) <= REL_TOL


def test_leaf_stats_small_tree_kernels(monkeypatch):
    """csrc/leaf_stats.cuh: trees with <= 2 / <= 4 bottom nodes take k_leaf_stats_small<2, 4> / <4, 2> (bins in registers, pattern table
    in a register, software-pipelined loads), larger ones k_leaf_stats (shared-memory bins).  Counts bit exact, sums 1e-10 against the
    oracle, and the small-tree kernels against the shared-memory one (S4B_LEAF_REG_BINS=0) to rounding.  Trees after a few sweeps
    (1-6 bottom nodes), a ragged last quad, several CTAs."""
    n, T = 150_003, 30
    o, g, _ = make_pair(n=n, num_trees=T)
    o.sample_trees_from_prior(); g.sample_trees_from_prior()
    for _ in range(8):
        o.run(); g.run()
    sizes = set()
    for t in range(T):
        monkeypatch.delenv("S4B_LEAF_REG_BINS", raising=False)
        hr, cr, sr, ssr = g.leaf_stats(t)
        monkeypatch.setenv("S4B_LEAF_REG_BINS", "0")
        hs, cs, ss, sss = g.leaf_stats(t)
        monkeypatch.delenv("S4B_LEAF_REG_BINS", raising=False)
        ho, co, so, sso = o.leaf_stats(t)
        scale = np.sqrt(np.maximum(co, 1) * sso) + 1e-300
        assert np.array_equal(hr, hs) and np.array_equal(cr, cs)
        assert rel_err(sss, ssr) <= 1e-13 and rel_err(ss, sr, scale=scale) <= 1e-13
        assert np.array_equal(ho, hr) and np.array_equal(co, cr)
        assert rel_err(sso, ssr) <= REL_TOL
        assert rel_err(so, sr, scale=scale) <= REL_TOL
        sizes.add(len(hr))
    assert min(sizes) <= 2 and any(3 <= k <= 4 for k in sizes) and any(5 <= k <= 8 for k in sizes), sizes


@pytest.mark.parametrize("binary", [False, True])
def test_stepwise_replay_native_rng(binary):
    """Same counter-based RNG on both sides: every tree step of the first sweeps matches."""
    T, sweeps = 10, 6
    o, g, _ = make_pair(n=600, num_trees=T, binary=binary, n_test=50)
    o.sample_trees_from_prior(); g.sample_trees_from_prior()
    o.set_trace(T * sweeps); g.set_trace(T * sweeps)
    for s in range(sweeps):
        ro, rg = o.run(), g.run()
        assert rel_err(ro["train"], rg["train"], scale=np.abs(ro["train"]) + 1.0) <= REL_TOL, f"sweep {s}"
        assert rel_err(ro["test"], rg["test"], scale=np.abs(ro["test"]) + 1.0) <= REL_TOL
        assert np.array_equal(ro["varcount"], rg["varcount"])
        assert ro["sigma"] == pytest.approx(rg["sigma"], rel=1e-14)
    compare_traces(o.trace(), g.trace())
    assert o.rng_counter() == g.rng_counter()
    assert_same_partition(o, g, T)
    if binary:
        assert rel_err(o.latents(), g.latents(), scale=1.0) <= 1e-9


def test_replay_from_tape():
    """Replay mode: the GPU consumes the draws recorded by the oracle, not its own generator."""
    T, sweeps = 7, 5
    o, g, _ = make_pair(n=400, num_trees=T, seed=5)
    o.set_record(200000)
    o.sample_trees_from_prior()
    o.set_trace(T * sweeps)
    outs = [o.run() for _ in range(sweeps)]
    tape = o.record()
    g2 = g
    g2.set_tape(tape)
    g2.sample_trees_from_prior()
    g2.set_trace(T * sweeps)
    for s in range(sweeps):
        rg = g2.run()
        assert rel_err(outs[s]["train"], rg["train"], scale=np.abs(outs[s]["train"]) + 1.0) <= REL_TOL
    compare_traces(o.trace(), g2.trace())


def test_tape_underrun_is_an_error():
    from stan4bart_b200._lib import S4BError
    o, g, _ = make_pair(n=300, num_trees=4)
    g.set_tape(np.full(3, 0.5))
    with pytest.raises(S4BError):
        g.sample_trees_from_prior()
        g.run()


def test_all_move_types_and_all_sweep_modes():
    """Long enough that birth, death, change and swap are all proposed and accepted.  The three execution
    modes (persistent on-chip sweep, per-tree kernels in a CUDA graph, per-tree kernels on the stream) all
    match the oracle; graph and stream launches of the same kernels agree bit for bit."""
    T, sweeps = 20, 30
    o, g, (x, y, xt) = make_pair(n=1500, p=6, num_trees=T, seed=21)
    assert g.sweep_mode() == 2
    cfg = g.cfg
    off = 0.3 * x[:, 3] - 0.1
    others = []
    for mode in (1, 0):
        h = GpuBart(cfg, y, x, xt)
        h.set_offset(off, True); h.set_sigma(1.3); h.set_sweep_mode(mode)
        others.append(h)
    for b in [o, g] + others:
        b.sample_trees_from_prior()
        b.set_trace(T * sweeps)
    for s in range(sweeps):
        ro, rg = o.run(), g.run()
        r1, r0 = others[0].run(), others[1].run()
        assert np.array_equal(r1["train"], r0["train"])
        assert rel_err(ro["train"], rg["train"], scale=np.abs(ro["train"]) + 1.0) <= REL_TOL
        assert rel_err(ro["train"], r1["train"], scale=np.abs(ro["train"]) + 1.0) <= REL_TOL
    tr = o.trace()
    compare_traces(tr, g.trace())
    compare_traces(tr, others[0].trace())
    kinds = tr[:, 0]
    for k in (0, 1, 2, 3):
        assert np.any((kinds == k) & (tr[:, 4] == 1)), f"move type {k} never accepted in this run"
    assert_same_partition(o, g, T)
    assert_same_partition(o, others[0], T)


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_deep_trees_many_slots(mode):
    """Weak leaf prior + tiny min_obs grows trees past 8 statistic slots (several bin passes)."""
    T, sweeps = 4, 60
    x, y, _ = bart_problem(3000, 4, 0)
    cfg = bart_config(3000, 4, num_trees=T, seed=8, min_obs=1, base=0.99, power=0.5, k=0.5)
    o, g = O.OracleBart(cfg, y, x), GpuBart(cfg, y, x)
    g.set_sweep_mode(mode)
    o.set_sigma(0.3); g.set_sigma(0.3)
    o.set_trace(T * sweeps); g.set_trace(T * sweeps)
    for _ in range(sweeps):
        ro, rg = o.run(), g.run()
    tr = o.trace()
    assert tr[:, 8].max() > 9, "trees did not grow deep enough for this test to bite"
    compare_traces(tr, g.trace())
    assert rel_err(ro["train"], rg["train"], scale=np.abs(ro["train"]) + 1.0) <= REL_TOL
    assert_same_partition(o, g, T)


def test_thinning_and_rescale_schedule():
    T = 6
    o, g, (x, y, _) = make_pair(n=500, num_trees=T, thin=3)
    o.sample_trees_from_prior(); g.sample_trees_from_prior()
    rng = np.random.default_rng(0)
    for it in range(6):
        off = 0.3 * x[:, 3] + 0.05 * rng.standard_normal(len(y))
        upd = it % 2 == 0
        o.set_offset(off, upd); g.set_offset(off, upd)
        o.set_sigma(1.0 + 0.1 * it); g.set_sigma(1.0 + 0.1 * it)
        ro, rg = o.run(), g.run()
        assert rel_err(ro["train"], rg["train"], scale=np.abs(ro["train"]) + 1.0) <= REL_TOL
        assert rel_err(o.data_range(), g.data_range()) <= 1e-14
    assert o.rng_counter() == g.rng_counter()


def test_predict_matches_training_fits_and_oracle():
    """test-01-continuous.R:212-254: predict(newdata = train) reproduces the stored training fits."""
    o, g, (x, y, _) = make_pair(n=640, num_trees=15)
    o.sample_trees_from_prior(); g.sample_trees_from_prior()
    for _ in range(5):
        ro, rg = o.run(), g.run()
    off = 0.3 * x[:, 3] - 0.1
    pg = g.predict(x, off)
    assert rel_err(pg, rg["train"], scale=np.abs(pg) + 1.0) <= 1e-10
    xnew = np.asfortranarray(np.random.default_rng(9).random((333, x.shape[1])))
    assert rel_err(o.predict(xnew), g.predict(xnew), scale=1.0) <= 1e-10
    to, tg = o.trees(), g.trees()
    assert np.array_equal(to["var"], tg["var"]) and np.array_equal(to["n"], tg["n"])
    assert rel_err(to["value"], tg["value"]) <= REL_TOL


@pytest.mark.parametrize("n", [1, 3, 4, 5, 17, 255, 256, 257, 1025])
def test_ragged_sizes(n):
    T = 3
    x, y, xt = bart_problem(max(n, 1), 2, 0)
    cfg = bart_config(n, 2, num_trees=T, seed=2, min_obs=1)
    o, g = O.OracleBart(cfg, y, x), GpuBart(cfg, y, x)
    o.set_sigma(0.7); g.set_sigma(0.7)
    o.set_trace(T * 4); g.set_trace(T * 4)
    for _ in range(4):
        ro, rg = o.run(), g.run()
        assert rel_err(ro["train"], rg["train"], scale=np.abs(ro["train"]) + 1.0) <= REL_TOL
    compare_traces(o.trace(), g.trace())


def test_bad_arguments_fail_loudly():
    from stan4bart_b200._lib import S4BError
    x, y, _ = bart_problem(50, 2)
    with pytest.raises(S4BError):
        GpuBart(bart_config(50, 2, n_cuts=300), y, x)
    g = GpuBart(bart_config(50, 2, num_trees=2), y, x)
    with pytest.raises(S4BError):
        g.node_assignment(5)


@pytest.mark.parametrize("binary", [False, True])
def test_streamed_sweep_variant(monkeypatch, binary):
    """Shards beyond the register file run the same persistent sweep with residuals and cached node indices streamed
    from global memory (S4B_FORCE_STREAM selects it at any size): all move types, n not a multiple of anything, and
    bit-for-bit agreement with the register-resident variant (same arithmetic in the same order)."""
    T, sweeps = 12, 25
    x, y, xt = bart_problem(2777, 6, 0, binary, seed=21)
    cfg = bart_config(2777, 6, num_trees=T, is_binary=binary, seed=33)
    off = 0.3 * x[:, 3] - 0.1
    o = O.OracleBart(cfg, y, x, xt)
    g_reg = GpuBart(cfg, y, x, xt)
    monkeypatch.setenv("S4B_FORCE_STREAM", "1")
    g_str = GpuBart(cfg, y, x, xt)
    monkeypatch.delenv("S4B_FORCE_STREAM")
    assert g_str.sweep_mode() == 2
    for b in (o, g_reg, g_str):
        b.set_offset(off, True)
        if not binary:
            b.set_sigma(1.3)
        b.sample_trees_from_prior()
        b.set_trace(T * sweeps)
    for s in range(sweeps):
        ro, r1, r2 = o.run(), g_reg.run(), g_str.run()
        assert np.array_equal(r1["train"], r2["train"]), f"sweep {s}"
        assert rel_err(ro["train"], r2["train"], scale=np.abs(ro["train"]) + 1.0) <= REL_TOL
    compare_traces(o.trace(), g_str.trace())
    assert np.array_equal(g_reg.trace(), g_str.trace())
    assert_same_partition(o, g_str, T)


def test_streamed_sweep_variant_replays_a_tape(monkeypatch):
    T, sweeps = 7, 5
    o, _, (x, y, xt) = make_pair(n=900, num_trees=T, seed=5)
    monkeypatch.setenv("S4B_FORCE_STREAM", "1")
    g = GpuBart(o.cfg, y, x, xt)
    monkeypatch.delenv("S4B_FORCE_STREAM")
    off = 0.3 * x[:, 3] - 0.1
    g.set_offset(off, True); g.set_sigma(1.3)
    o.set_record(200000)
    o.sample_trees_from_prior()
    o.set_trace(T * sweeps)
    outs = [o.run() for _ in range(sweeps)]
    g.set_tape(o.record())
    g.sample_trees_from_prior()
    g.set_trace(T * sweeps)
    for s in range(sweeps):
        rg = g.run()
        assert rel_err(outs[s]["train"], rg["train"], scale=np.abs(outs[s]["train"]) + 1.0) <= REL_TOL
    compare_traces(o.trace(), g.trace())


@pytest.mark.parametrize("nq", [2, 4, 6])
def test_register_variants_of_the_sweep_kernel(monkeypatch, nq):
    """The sweep kernel keeps 4 * NQ observations per thread in registers (NQ grows with the shard); S4B_FORCE_NQ selects an
    instantiation at any size.  Every instantiation reproduces the oracle and the NQ = 1 run bit for bit."""
    T, sweeps = 10, 12
    x, y, xt = bart_problem(5003, 6, 0, False, seed=9)
    cfg = bart_config(5003, 6, num_trees=T, seed=77)
    off = 0.3 * x[:, 3] - 0.1
    o = O.OracleBart(cfg, y, x, xt)
    g1 = GpuBart(cfg, y, x, xt)
    monkeypatch.setenv("S4B_FORCE_NQ", str(nq))
    g2 = GpuBart(cfg, y, x, xt)
    monkeypatch.delenv("S4B_FORCE_NQ")
    for b in (o, g1, g2):
        b.set_offset(off, True); b.set_sigma(1.3)
        b.sample_trees_from_prior()
        b.set_trace(T * sweeps)
    for s in range(sweeps):
        ro, r1, r2 = o.run(), g1.run(), g2.run()
        assert np.array_equal(r1["train"], r2["train"]), f"sweep {s}"
        assert rel_err(ro["train"], r2["train"], scale=np.abs(ro["train"]) + 1.0) <= REL_TOL
    compare_traces(o.trace(), g2.trace())


@pytest.mark.parametrize("binary", [False, True])
def test_full_grid_at_scale(binary):
    """n = 600 000 rows: all 148 CTAs, 16 observations per thread (the instantiation of the n = 1 M benchmark), the
    grid-wide barrier and reduction at full width.  First sweeps against the oracle; partitions bit-exact."""
    T, sweeps = 8, 3
    n = 600000
    x, y, xt = bart_problem(n, 5, 0, binary, seed=4)
    cfg = bart_config(n, 5, num_trees=T, is_binary=binary, seed=5)
    o = O.OracleBart(cfg, y, x, xt)
    g = GpuBart(cfg, y, x, xt)
    assert g.sweep_mode() == 2
    for b in (o, g):
        if not binary:
            b.set_sigma(1.1)
        b.sample_trees_from_prior()
        b.set_trace(T * sweeps)
    for s in range(sweeps):
        ro, rg = o.run(), g.run()
        assert rel_err(ro["train"], rg["train"], scale=np.abs(ro["train"]) + 1.0) <= 1e-9, f"sweep {s}"
    compare_traces(o.trace(), g.trace(), tol=1e-9)
    assert_same_partition(o, g, T)


@pytest.mark.parametrize("binary", [False, True])
def test_production_path_without_sums_of_squares(binary):
    """Without the parity trace the sweep kernel does not accumulate sum(r^2): every Metropolis ratio compares two partitions
    of the same observations, so those terms cancel.  The untraced (production) run must take the same decisions as the
    traced run and as the oracle: same trees, fits equal to rounding."""
    T, sweeps = 15, 30
    x, y, xt = bart_problem(3001, 6, 0, binary, seed=13)
    cfg = bart_config(3001, 6, num_trees=T, is_binary=binary, seed=99)
    off = 0.3 * x[:, 3] - 0.1
    o = O.OracleBart(cfg, y, x, xt)
    g_tr, g_pl = GpuBart(cfg, y, x, xt), GpuBart(cfg, y, x, xt)
    for b in (o, g_tr, g_pl):
        b.set_offset(off, True)
        if not binary:
            b.set_sigma(1.3)
        b.sample_trees_from_prior()
    g_tr.set_trace(T * sweeps)
    for s in range(sweeps):
        ro, r1, r2 = o.run(), g_tr.run(), g_pl.run()
        assert rel_err(r1["train"], r2["train"], scale=np.abs(r1["train"]) + 1.0) <= 1e-10, f"sweep {s}"
        assert rel_err(ro["train"], r2["train"], scale=np.abs(ro["train"]) + 1.0) <= 1e-9, f"sweep {s}"
        assert np.array_equal(ro["varcount"], r2["varcount"])
    to, t2 = o.trees(), g_pl.trees()
    assert np.array_equal(to["var"], t2["var"]) and np.array_equal(to["n"], t2["n"])
    assert_same_partition(o, g_pl, T)
    assert g_tr.rng_counter() == g_pl.rng_counter() == o.rng_counter()


def test_split_probs_weighted_variable_selection():
    """bart_args split.probs (tests/testthat/test-09-bartArgs.R:20-39): predictors are drawn, and enter the rule prior, with the
    given relative probabilities.  Integer weights make the choice exact on both sides: every step matches the oracle in the
    persistent kernel and in the per-tree kernels; a predictor of probability zero is never split on."""
    T, sweeps = 12, 25
    x, y, xt = bart_problem(1500, 6, 0, False, seed=31)
    sp = [1.0, 1.0, 2.0, 0.0, 1.0, 3.0]
    cfg = bart_config(1500, 6, num_trees=T, seed=8, split_probs=sp)
    o = O.OracleBart(cfg, y, x, xt)
    g2, g1 = GpuBart(cfg, y, x, xt), GpuBart(cfg, y, x, xt)
    g1.set_sweep_mode(1)
    for b in (o, g2, g1):
        b.set_sigma(1.2)
        b.sample_trees_from_prior()
        b.set_trace(T * sweeps)
    used = np.zeros(6)
    for s in range(sweeps):
        ro, r2, r1 = o.run(), g2.run(), g1.run()
        assert np.array_equal(ro["varcount"], r2["varcount"]) and np.array_equal(ro["varcount"], r1["varcount"])
        assert rel_err(ro["train"], r2["train"], scale=np.abs(ro["train"]) + 1.0) <= REL_TOL
        used += ro["varcount"]
    compare_traces(o.trace(), g2.trace())
    compare_traces(o.trace(), g1.trace())
    assert used[3] == 0 and used.sum() > 0
    kinds = o.trace()[:, 0]
    assert np.any(kinds == 2) and np.any(kinds == 0)          # change and birth steps drew weighted variables


def test_bart_args_plumb_through_like_test_09():
    """tests/testthat/test-09-bartArgs.R:20-39: bart_args n.trees = 2, power = 2.5, base = 0.9, split.probs = c(X3 = 2, .default = 1)
    must reach the tree sampler; the reference checks it by parsing the initial summary."""
    import re
    x, y, xt = bart_problem(200, 5, 0, False)
    cfg = bart_config(200, 5, num_trees=2, power=2.5, base=0.9, split_probs=[1, 1, 2, 1, 1], seed=1)
    g = GpuBart(cfg, y, x, xt)
    out = g.summary().splitlines()
    pb = [ln for ln in out if "\tpower and base for tree prior:" in ln]
    assert len(pb) == 1
    power, base = (float(v) for v in re.sub(r"^[^0-9]+([0-9.]+ [0-9.]+)$", r"\1", pb[0]).split(" "))
    assert (power, base) == (2.5, 0.9)
    sp = [ln for ln in out if "\ttree split probabilities:" in ln]
    assert len(sp) == 1
    probs = [float(v) for v in re.sub(r"^[^0-9]+((?:[0-9.]+, )*[0-9.]+)$", r"\1", sp[0]).split(", ")]
    assert np.allclose(probs, [1 / 6, 1 / 6, 2 / 6, 1 / 6, 1 / 6], atol=1e-5)
    assert "number of trees: 2" in "\n".join(out)


def test_chain_confined_to_a_share_of_the_sms(monkeypatch):
    """Several chains per GPU (config D): max_ctas confines a chain's sweep kernel to a few SMs (register or streamed variant,
    whichever holds the rows); the draws do not depend on the grid size beyond the order of the final sums."""
    T, sweeps = 10, 8
    n = 60000
    x, y, xt = bart_problem(n, 5, 0, False, seed=2)
    o = O.OracleBart(bart_config(n, 5, num_trees=T, seed=6), y, x, xt)
    fits = [o]
    for cap in (0, 7, 2):
        fits.append(GpuBart(bart_config(n, 5, num_trees=T, seed=6, max_ctas=cap), y, x, xt))
    for b in fits:
        b.set_sigma(1.1)
        b.sample_trees_from_prior()
        b.set_trace(T * sweeps)
    for s in range(sweeps):
        rs = [b.run() for b in fits]
        for r in rs[1:]:
            assert rel_err(rs[0]["train"], r["train"], scale=np.abs(rs[0]["train"]) + 1.0) <= 1e-9, f"sweep {s}"
    for b in fits[1:]:
        compare_traces(o.trace(), b.trace(), tol=1e-9)
        assert b.sweep_mode() == 2


@pytest.mark.parametrize("binary", [False, True])
def test_observation_weights(binary):
    """`weights` of stan4bart() as dbarts data weights: leaf statistics sum w / sum w r; same decisions, partitions and
    leaf draws as the oracle, traced and untraced.  Some zero weights: such rows follow the trees but carry no information."""
    T, sweeps, n = 12, 8, 2500
    x, y, xt = bart_problem(n, 6, 40, binary, seed=21)
    rng = np.random.default_rng(5)
    wt = rng.gamma(2.0, 0.5, n)
    wt[rng.random(n) < 0.05] = 0.0
    cfg = bart_config(n, 6, n_test=40, num_trees=T, is_binary=binary, seed=77, weights=wt)
    off = 0.3 * x[:, 3] - 0.1
    o = O.OracleBart(cfg, y, x, xt)
    g_tr, g_pl = GpuBart(cfg, y, x, xt), GpuBart(cfg, y, x, xt)
    for b in (o, g_tr, g_pl):
        b.set_offset(off, True)
        if not binary:
            b.set_sigma(1.3)
        b.sample_trees_from_prior()
    o.set_trace(T * sweeps); g_tr.set_trace(T * sweeps)
    for s in range(sweeps):
        ro, r1, r2 = o.run(), g_tr.run(), g_pl.run()
        assert rel_err(ro["train"], r1["train"], scale=np.abs(ro["train"]) + 1.0) <= 1e-9, f"sweep {s}"
        assert rel_err(r1["train"], r2["train"], scale=np.abs(r1["train"]) + 1.0) <= 1e-10, f"sweep {s}"
        assert rel_err(ro["test"], r2["test"], scale=np.abs(ro["test"]) + 1.0) <= 1e-9
        assert np.array_equal(ro["varcount"], r2["varcount"])
    compare_traces(o.trace(), g_tr.trace(), ll_difference_only=True)
    assert_same_partition(o, g_pl, T)
    assert g_tr.rng_counter() == g_pl.rng_counter() == o.rng_counter()
    # the weights matter: an unweighted chain from the same seed takes other decisions
    cfg_u = bart_config(n, 6, n_test=40, num_trees=T, is_binary=binary, seed=77)
    u = GpuBart(cfg_u, y, x, xt)
    u.set_offset(off, True)
    if not binary:
        u.set_sigma(1.3)
    u.sample_trees_from_prior()
    for s in range(sweeps):
        ru = u.run()
    assert not np.array_equal(ru["train"], r2["train"])


def test_constant_weights_rescale_sigma():
    """y_i ~ N(mu, sigma^2 / c) for every i is the unweighted model with sigma / sqrt(c)."""
    T, n, c = 10, 1500, 4.0
    x, y, xt = bart_problem(n, 5, 0, False, seed=8)
    gw = GpuBart(bart_config(n, 5, num_trees=T, seed=5, weights=np.full(n, c)), y, x, xt)
    gu = GpuBart(bart_config(n, 5, num_trees=T, seed=5), y, x, xt)
    gw.set_sigma(1.3); gu.set_sigma(1.3 / np.sqrt(c))
    gw.sample_trees_from_prior(); gu.sample_trees_from_prior()
    for s in range(6):
        rw, ru = gw.run(), gu.run()
        assert rel_err(rw["train"], ru["train"], scale=np.abs(ru["train"]) + 1.0) <= 1e-9
    assert np.array_equal(gw.trees()["var"], gu.trees()["var"])


def test_weighted_fit_refuses_the_per_tree_modes():
    from stan4bart_b200._lib import S4BError
    x, y, xt = bart_problem(300, 5, 0, False)
    g = GpuBart(bart_config(300, 5, num_trees=5, weights=np.ones(300)), y, x, xt)
    with pytest.raises(S4BError):
        g.set_sweep_mode(0)
    with pytest.raises(ValueError):
        bart_config(300, 5, weights=-np.ones(300))


@pytest.mark.parametrize("binary", [False, True])
def test_modelled_k_matches_oracle(binary):
    """bart_args k = chi(1.25, Inf): the k draw after every sweep (its own RNG substream) and the leaf prior it feeds."""
    T, sweeps = 12, 10
    o, g, _ = make_pair(n=900, num_trees=T, binary=binary, k_df=1.25)
    o.sample_trees_from_prior(); g.sample_trees_from_prior()
    o.set_trace(T * sweeps); g.set_trace(T * sweeps)
    assert o.k() == g.k() == 2.0
    ks = []
    for s in range(sweeps):
        ro, rg = o.run(), g.run()
        assert rel_err(ro["train"], rg["train"], scale=np.abs(ro["train"]) + 1.0) <= 1e-9, f"sweep {s}"
        assert abs(o.k() - g.k()) <= 1e-10 * o.k(), f"sweep {s}"
        ks.append(g.k())
    compare_traces(o.trace(), g.trace())
    assert o.rng_counter() == g.rng_counter()
    assert len(set(ks)) == sweeps


def test_keep_trees_leaves_the_split_and_observation_weights_alone():
    """Regression: asking for a tree store must not release the split weights / observation weights of the fit."""
    n, T = 800, 8
    x, y, xt = bart_problem(n, 5, 0, False, seed=4)
    wt = np.random.default_rng(1).gamma(2.0, 0.5, n)
    cfg = bart_config(n, 5, num_trees=T, seed=21, weights=wt, split_probs=[0.4, 0.1, 0.2, 0.2, 0.1])
    o, g = O.OracleBart(cfg, y, x, xt), GpuBart(cfg, y, x, xt)
    for b in (o, g):
        b.set_sigma(1.2); b.sample_trees_from_prior()
    g.set_keep_trees(4)
    for s in range(4):
        ro, rg = o.run(), g.run()
        assert rel_err(ro["train"], rg["train"], scale=np.abs(ro["train"]) + 1.0) <= 1e-9
        assert np.array_equal(ro["varcount"], rg["varcount"])
    assert_same_partition(o, g, T)


@pytest.mark.parametrize("case", range(16))
def test_randomised_configurations_against_the_oracle(case, monkeypatch):
    """Random draws over the configuration space (size, predictors, trees, response type, thinning, n.cuts, min leaf size,
    tree prior, move probabilities, split.probs, weights, modelled k, sweep kernel variant): decisions, partitions and fits."""
    rng = np.random.default_rng(1000 + case)
    n = int(rng.choice([37, 150, 600, 2100, 5003]))
    p = int(rng.integers(1, 12))
    T = int(rng.integers(1, 14))
    binary = bool(rng.integers(0, 2))
    n_test = int(rng.choice([0, 1, 33]))
    kw = dict(thin=int(rng.choice([1, 1, 2, 3])), n_cuts=int(rng.choice([1, 3, 20, 100, 255])), min_obs=int(rng.choice([1, 5, 20])),
              base=float(rng.uniform(0.5, 0.99)), power=float(rng.uniform(0.5, 3.0)), k=float(rng.uniform(1.0, 4.0)))
    if rng.random() < 0.4:
        sp = rng.random(p) + 0.05
        sp[rng.random(p) < 0.2] = 0.0
        if not np.any(sp > 0):
            sp[0] = 1.0
        kw["split_probs"] = sp
    if rng.random() < 0.4:
        kw["weights"] = rng.gamma(2.0, 0.5, n)
    if rng.random() < 0.3:
        kw["k_df"] = float(rng.uniform(0.5, 3.0))
    if case >= 10:               # bart_args n.cuts as a vector: one count per predictor
        kw["n_cuts"] = rng.integers(1, 60, p)
    if case in (3, 7, 11, 13):   # bart_args use.quantiles
        kw["use_quantiles"] = True
    variant = rng.choice(["auto", "stream", "nq2", "nq6"])
    if variant == "stream":
        monkeypatch.setenv("S4B_FORCE_STREAM", "1")
    elif variant.startswith("nq") and "weights" not in kw:
        monkeypatch.setenv("S4B_FORCE_NQ", variant[2:])
    x, y, xt = bart_problem(n, p, n_test, binary, seed=case)
    cfg = bart_config(n, p, n_test=n_test, num_trees=T, is_binary=binary, seed=500 + case, **kw)
    o, g = O.OracleBart(cfg, y, x, xt), GpuBart(cfg, y, x, xt)
    off = 0.2 * np.sin(np.arange(n))
    sweeps = 5
    for b in (o, g):
        b.set_offset(off, True)
        if not binary:
            b.set_sigma(float(0.5 + case % 3))
        b.sample_trees_from_prior()
    traced = case % 2 == 0
    if traced:
        o.set_trace(T * sweeps * kw["thin"]); g.set_trace(T * sweeps * kw["thin"])
    for s in range(sweeps):
        ro, rg = o.run(), g.run()
        assert rel_err(ro["train"], rg["train"], scale=np.abs(ro["train"]) + 1.0) <= 1e-8, f"sweep {s}"
        if n_test:
            assert rel_err(ro["test"], rg["test"], scale=np.abs(ro["test"]) + 1.0) <= 1e-8
        assert np.array_equal(ro["varcount"], rg["varcount"])
        assert abs(o.k() - g.k()) <= 1e-9 * o.k()
    if traced:
        compare_traces(o.trace(), g.trace(), tol=1e-8, ll_difference_only="weights" in kw)
    assert_same_partition(o, g, T)
    assert o.rng_counter() == g.rng_counter()


def test_cut_counts_per_predictor():
    """bart_args n.cuts as a vector (R/stan4bart_fit.R:446-451): every predictor has its own number of uniform cut points;
    splits on a predictor never use a cut index beyond its own count."""
    n, T = 1200, 12
    x, y, xt = bart_problem(n, 4, 25, False, seed=5)
    counts = np.array([1, 2, 100, 7])
    cfg = bart_config(n, 4, n_test=25, num_trees=T, seed=3, n_cuts=counts)
    o, g = O.OracleBart(cfg, y, x, xt), GpuBart(cfg, y, x, xt)
    for b in (o, g):
        b.set_sigma(1.0); b.sample_trees_from_prior()
    o.set_trace(T * 8); g.set_trace(T * 8)
    for s in range(8):
        ro, rg = o.run(), g.run()
        assert rel_err(ro["train"], rg["train"], scale=np.abs(ro["train"]) + 1.0) <= 1e-9
        assert rel_err(ro["test"], rg["test"], scale=np.abs(ro["test"]) + 1.0) <= 1e-9
    compare_traces(o.trace(), g.trace())
    assert_same_partition(o, g, T)
    tr = g.trace()
    births = tr[(tr[:, 0] == 0) & (tr[:, 2] >= 0)]
    assert births.shape[0] > 0 and np.all(births[:, 3] < counts[births[:, 2].astype(int)])
    tg, to = g.trees(), o.trees()
    assert np.array_equal(tg["var"], to["var"]) and rel_err(tg["value"], to["value"], scale=np.abs(to["value"]) + 1e-3) <= 1e-9


@pytest.mark.parametrize("binary,streamed", [(False, False), (True, False), (False, True)])
def test_chains_batched_in_one_launch(binary, streamed, monkeypatch):
    """several chains, ONE cooperative launch (k_sweep_batch, grid.y = chain), every chain on its share of the SMs.
    Four chains with different data and seeds: each must evolve exactly like the same chain run on its own (same kernel code, same
    draws: bit for bit) and like its oracle."""
    if streamed:                                   # residuals and node indices streamed from global memory (large shards): k_sweep_batch<1, true>
        monkeypatch.setenv("S4B_FORCE_STREAM", "1")
    sms = 148
    chains, n, T = 4, 5000, 10
    data = [bart_problem(n=n, p=6, binary=binary, seed=60 + c) for c in range(chains)]
    mk = lambda c: bart_config(n, 6, num_trees=T, is_binary=binary, seed=900 + c, max_ctas=sms // chains)
    batch = [GpuBart(mk(c), data[c][1], data[c][0]) for c in range(chains)]
    solo = [GpuBart(mk(c), data[c][1], data[c][0]) for c in range(chains)]
    orc = [O.OracleBart(mk(c), data[c][1], data[c][0]) for c in range(chains)]
    for group in (batch, solo, orc):
        for b in group:
            if not binary:
                b.set_sigma(1.0)
            b.sample_trees_from_prior()
    for g in solo:
        g.set_pipeline(False)                      # the batched launch runs the synchronous kernel's code
    for s in range(6):
        rb = GpuBart.run_batched(batch)
        for c in range(chains):
            rs, ro = solo[c].run(), orc[c].run()
            assert np.array_equal(rb[c]["train"], rs["train"]) and np.array_equal(rb[c]["varcount"], rs["varcount"]), (s, c)
            assert rel_err(ro["train"], rb[c]["train"], scale=np.abs(ro["train"]) + 1.0) <= 1e-9
    for c in range(chains):
        tb, ts = batch[c].trees(), solo[c].trees()
        assert np.array_equal(tb["var"], ts["var"]) and np.array_equal(tb["value"], ts["value"])
        assert batch[c].rng_counter() == solo[c].rng_counter() == orc[c].rng_counter()
    if streamed:
        return                                     # (the streamed layout has no predictor tile: any number of predictors may share a batch)
    with pytest.raises(Exception):                 # a fit of another shape class (other predictor tile) cannot join the batch
        xo, yo, _ = bart_problem(n=n, p=3, binary=binary, seed=1)
        GpuBart.run_batched(batch + [GpuBart(bart_config(n, 3, num_trees=T, is_binary=binary, seed=1, max_ctas=sms // chains), yo, xo)])


def test_two_fits_of_different_shapes_in_one_process():
    """The opt-in limit of dynamic shared memory is an attribute of the kernel FUNCTION: a second fit with fewer predictors (a smaller
    predictor tile) must not lower the limit under the first fit's launches (two stan4bart fits in one session).  Regression test:
    fit A runs, fit B (other shape) is created and runs, fit A runs again -- and still equals its oracle."""
    xa, ya, _ = bart_problem(n=9000, p=9, binary=False, seed=41)
    xb, yb, _ = bart_problem(n=4000, p=2, binary=True, seed=42)
    cfg_a = bart_config(9000, 9, num_trees=12, seed=5)
    oa, ga = O.OracleBart(cfg_a, ya, xa), GpuBart(cfg_a, ya, xa)
    for b in (oa, ga):
        b.set_sigma(1.0); b.sample_trees_from_prior()
    for _ in range(2):
        ro, rg = oa.run(), ga.run()
    gb = GpuBart(bart_config(4000, 2, num_trees=7, seed=6, is_binary=True), yb, xb)
    gb.sample_trees_from_prior()
    for _ in range(2):
        gb.run()
    for _ in range(3):
        ro, rg = oa.run(), ga.run()
        assert rel_err(ro["train"], rg["train"], scale=np.abs(ro["train"]) + 1.0) <= 1e-9
    gb.run()
    assert_same_partition(oa, ga, 12)


def test_quantile_cut_points():
    """bart_args use.quantiles: cut points between the distinct sorted values (discrete predictors get a cut in every gap, a
    constant predictor none); same cuts, bins, decisions and fits as the oracle; test rows are binned against the same cuts."""
    n, T = 3000, 15
    rng = np.random.default_rng(12)
    x = np.empty((n, 5)); xt = np.empty((40, 5))
    for a, m in ((x, n), (xt, 40)):
        a[:, 0] = rng.integers(0, 2, m); a[:, 1] = rng.integers(0, 6, m) * 1.5; a[:, 2] = rng.standard_normal(m) ** 3
        a[:, 3] = 7.0; a[:, 4] = rng.random(m)
    x, xt = np.asfortranarray(x), np.asfortranarray(xt)
    y = 2.0 * x[:, 0] + 0.4 * x[:, 1] + np.tanh(x[:, 2]) + x[:, 4] + 0.2 * rng.standard_normal(n)
    cfg = bart_config(n, 5, n_test=40, num_trees=T, seed=9, n_cuts=np.array([100, 100, 30, 100, 255]), use_quantiles=True)
    o, g = O.OracleBart(cfg, y, x, xt), GpuBart(cfg, y, x, xt)
    for b in (o, g):
        b.set_sigma(0.7); b.sample_trees_from_prior()
    o.set_trace(T * 10); g.set_trace(T * 10)
    for s in range(10):
        ro, rg = o.run(), g.run()
        assert rel_err(ro["train"], rg["train"], scale=np.abs(ro["train"]) + 1.0) <= 1e-9
        assert rel_err(ro["test"], rg["test"], scale=np.abs(ro["test"]) + 1.0) <= 1e-9
        assert np.array_equal(ro["varcount"], rg["varcount"]) and ro["varcount"][3] == 0
    compare_traces(o.trace(), g.trace())
    assert_same_partition(o, g, T)
    tg, to = g.trees(), o.trees()
    rules = tg["var"] >= 0
    assert np.array_equal(tg["var"], to["var"]) and np.array_equal(tg["value"][rules], to["value"][rules])          # cut values bit for bit
    assert rel_err(tg["value"], to["value"], scale=np.abs(to["value"]) + 1e-3) <= 1e-9
    assert set(np.unique(tg["value"][rules & (tg["var"] == 0)])) <= {0.5}
    assert set(np.unique(tg["value"][rules & (tg["var"] == 1)])) <= {0.75, 2.25, 3.75, 5.25, 6.75}
    assert "use quantiles for rule cut points: true" in g.summary()


@pytest.mark.parametrize("binary", [False, True])
def test_pipelined_sweep_kernel_equals_the_synchronous_one(binary):
    """csrc/sweep_pipe.cuh: workers one step ahead of the controller, slot sums corrected with the cross table (slot of t) x (cell
    of t - 1).  Same decisions, draws and trees as the synchronous kernel; fits / residuals differ only by the rounding of the
    slot sums.  Also against the oracle, and the kernel really ran (device-side counter)."""
    n, p, T, sweeps = 6000, 6, 40, 12
    x, y, _ = bart_problem(n=n, p=p, binary=binary, seed=31)
    off = 0.2 * np.sin(np.arange(n) * 0.05)
    res = {}
    for mode in ("pipe", "sync", "oracle"):
        cfg = bart_config(n, p, num_trees=T, is_binary=binary, seed=777)
        s = O.OracleBart(cfg, y, x) if mode == "oracle" else GpuBart(cfg, y, x)
        if mode == "sync":
            s.set_pipeline(False)
        s.set_offset(off, True)
        if not binary:
            s.set_sigma(0.9)
        s.sample_trees_from_prior()
        for _ in range(sweeps):
            r = s.run()
        res[mode] = dict(train=r["train"], trees=s.trees(), resid=s.residual(), vc=r["varcount"])
        if mode == "pipe":
            pl = s.pipeline()
            assert pl["enabled"] and pl["sweeps_offered"] == sweeps and pl["steps_pipelined"] >= 0.5 * sweeps * T, pl
        if mode == "sync":
            assert not s.pipeline()["enabled"] and s.pipeline()["steps_pipelined"] == 0
    for other in ("sync", "oracle"):
        a, b = res["pipe"], res[other]
        assert np.array_equal(a["trees"]["var"], b["trees"]["var"]) and np.array_equal(a["trees"]["n"], b["trees"]["n"]), other
        assert np.array_equal(a["vc"], b["vc"])
        assert rel_err(a["trees"]["value"], b["trees"]["value"], scale=np.abs(b["trees"]["value"]) + 1e-3) <= 1e-9
        assert rel_err(a["train"], b["train"], scale=np.abs(b["train"]) + 1.0) <= 1e-10
        assert rel_err(a["resid"], b["resid"], scale=np.abs(b["resid"]) + 1.0) <= 1e-10
