"""GPU parity of the whole Gibbs sweep (BART block + Stan block + plumbing) against the CPU oracle."""
import numpy as np
import pytest

import oracle_lib as O
from common import compare_traces, rel_err
from stan4bart_b200.frontend import friedman_problem
from stan4bart_b200.sampler import Sampler
from stan4bart_b200.structs import bart_config, stan_control

pytestmark = pytest.mark.gpu


def make_pair(n=100, binary=False, num_trees=11, warmup=7, iter_=13, keep_fits=True, seed=12345, thin=1):
    pr = friedman_problem(n, binary=binary)
    sd = pr["stan_data"]
    cfg = bart_config(n, 9, n_test=n, num_trees=num_trees, is_binary=binary, seed=seed, thin=thin)
    ctl = stan_control(seed=seed + 1)
    kw = dict(warmup=warmup, iter_=iter_, keep_fits=keep_fits, sigma_init=pr["sigma_init"], bart_offset_init=pr["bart_offset_init"])
    o = O.OracleSampler(cfg, pr["y"], pr["x_bart"], pr["x_test"], sd, ctl, **kw)
    g = Sampler(cfg, pr["y"], pr["x_bart"], pr["x_test"], sd, ctl, **kw)
    return o, g, pr


@pytest.mark.parametrize("binary", [False, True])
def test_first_sweeps_match_step_by_step(binary):
    """README config A (n = 100), tests' tiny settings: first K Gibbs sweeps agree step by step."""
    o, g, pr = make_pair(binary=binary)
    K = 7
    ob, gb = o.bart(), g.bart()
    ob.set_trace(11 * K); gb.set_trace(11 * K)
    ro, rg = o.run(K, True), g.run(K, True)
    compare_traces(ob.trace(), gb.trace(), tol=1e-8)
    assert rel_err(ro["stan"], rg["stan"], scale=np.abs(ro["stan"]) + 1.0) <= 1e-8
    assert rel_err(ro["bart"]["train"], rg["bart"]["train"], scale=np.abs(ro["bart"]["train"]) + 1.0) <= 1e-8
    assert rel_err(ro["bart"]["test"], rg["bart"]["test"], scale=np.abs(ro["bart"]["test"]) + 1.0) <= 1e-8
    assert np.array_equal(ro["bart"]["varcount"], rg["bart"]["varcount"])
    assert rel_err(ro["bart"]["sigma"], rg["bart"]["sigma"]) <= 1e-8
    assert rel_err(o.data_range(), g.data_range()) <= 1e-10
    # the saved 7 diagnostics are integers / booleans where they should be
    names = pr["stan_data"].param_names()
    for nm in ("treedepth__", "n_leapfrog__", "divergent__"):
        assert np.array_equal(ro["stan"][names.index(nm)], rg["stan"][names.index(nm)])


def test_result_layout_and_consistency():
    """Shapes of the reference's result list and the invariants of test-01-continuous.R / test-11-callback.R."""
    o, g, pr = make_pair(warmup=7, iter_=13)
    sd = pr["stan_data"]
    w = g.run(7, True)
    g.disengage_adaptation()
    r = g.run(6, False)
    assert w["stan"].shape == (58, 7) and r["stan"].shape == (58, 6)           # 7 + 26 + 25 rows
    assert r["bart"]["train"].shape == (100, 6) and r["bart"]["test"].shape == (100, 6)
    assert r["bart"]["varcount"].shape == (9, 6) and r["bart"]["sigma"].shape == (6,)
    names = sd.param_names()
    assert len(names) == 58 and not any(nm.startswith("gamma") for nm in names)    # test-01:29-32: no intercept
    # stored beta / b rows reproduce the parametric mean the BART block saw (test-11:57-66)
    glmm = O.OracleGlmm(sd)
    last = r["stan"][7:, -1]
    assert rel_err(g.parametric_mean(), glmm.parametric_mean(last), scale=1.0) <= 1e-12
    # sigma handed to BART is the stored aux row
    assert np.allclose(r["bart"]["sigma"], r["stan"][names.index("aux.1")])
    # predict on the training design reproduces the stored training fit of the last sweep (test-01:212-254)
    rng_min, rng_max = g.data_range()
    pred = g.predict_bart(pr["x_bart"])
    assert rel_err(pred, r["bart"]["train"][:, -1], scale=np.abs(pred) + 1.0) <= 1e-10
    # x_test == x_train for this design, so test fits equal training fits
    assert rel_err(r["bart"]["test"][:, -1], r["bart"]["train"][:, -1], scale=1.0) <= 1e-10
    # running means kept on device equal the mean of the stored draws (fitted == rowMeans(extract), test-01:34-57)
    m = g.means()
    assert m["num_draws"] == 6
    assert rel_err(m["bart_train"], r["bart"]["train"].mean(axis=1), scale=1.0) <= 1e-12
    assert rel_err(m["bart_test"], r["bart"]["test"].mean(axis=1), scale=1.0) <= 1e-12


def test_keep_fits_false_and_seed_determinism():
    _, g1, _ = make_pair(keep_fits=False)
    _, g2, _ = make_pair(keep_fits=False)
    _, g3, _ = make_pair(keep_fits=False, seed=777)
    outs = []
    for g in (g1, g2, g3):
        g.run(5, True)
        g.disengage_adaptation()
        outs.append(g.run(5, False))
    assert outs[0]["stan"].shape == (58, 1)
    assert np.array_equal(outs[0]["stan"], outs[1]["stan"])                 # test-05-rng.R: same seed, same draws
    assert np.array_equal(outs[0]["bart"]["train"], outs[1]["bart"]["train"])
    assert not np.array_equal(outs[0]["bart"]["train"], outs[2]["bart"]["train"])


def test_thin_and_larger_n():
    o, g, _ = make_pair(n=3000, num_trees=20, thin=2)
    ob, gb = o.bart(), g.bart()
    ob.set_trace(20 * 2 * 3); gb.set_trace(20 * 2 * 3)
    ro, rg = o.run(3, True), g.run(3, True)
    compare_traces(ob.trace(), gb.trace(), tol=1e-7)
    assert rel_err(ro["bart"]["train"], rg["bart"]["train"], scale=np.abs(ro["bart"]["train"]) + 1.0) <= 1e-7


@pytest.mark.parametrize("binary", [False, True])
def test_posterior_agreement_within_monte_carlo_error(binary):
    """north_star: posterior fitted means and CATE agree within 4 Monte-Carlo standard errors.
    Independent seeds on the two sides (a statistical, not a replay, check).  BART chains mix slowly, so
    the Monte-Carlo error of a posterior mean is estimated from the spread of independent chains.
    (Probit: fitted means and the treatment coefficient on the latent scale.)"""
    n, warm, it, chains = 200, 150, 300, 4
    pr = friedman_problem(n, binary=binary)
    sd = pr["stan_data"]
    kw = dict(warmup=warm, iter_=warm + it, keep_fits=True, sigma_init=pr["sigma_init"], bart_offset_init=pr["bart_offset_init"])
    d = pr["data"]
    names = sd.param_names()
    glmm = O.OracleGlmm(sd)
    means = {"oracle": [], "gpu": []}
    cates = {"oracle": [], "gpu": []}
    for label, cls, seed0 in (("oracle", O.OracleSampler, 100), ("gpu", Sampler, 200)):
        for c in range(chains):
            seed = seed0 + c
            cfg = bart_config(n, 9, n_test=n, num_trees=30, seed=seed, is_binary=binary)
            s = cls(cfg, pr["y"], pr["x_bart"], pr["x_test"], sd, stan_control(seed=seed), **kw)
            s.run(warm, True)
            s.disengage_adaptation()
            r = s.run(it, False)
            par = np.stack([glmm.parametric_mean(r["stan"][7:, k]) for k in range(it)], axis=1)
            means[label].append((r["bart"]["train"] + par).mean(axis=1))
            # readme.md:57-64: icate = (mu.train - mu.test) * (2 z - 1); the counterfactual differs only in z => beta_z
            cates[label].append(r["stan"][names.index("beta.2")].mean())
    mo, mg = np.stack(means["oracle"]), np.stack(means["gpu"])
    se = np.sqrt(mo.var(axis=0, ddof=1) / chains + mg.var(axis=0, ddof=1) / chains)
    diff = np.abs(mo.mean(axis=0) - mg.mean(axis=0))
    assert np.mean(diff <= 4 * se) >= 0.97        # fitted means, observation by observation
    co, cg = np.array(cates["oracle"]), np.array(cates["gpu"])
    assert abs(co.mean() - cg.mean()) <= 4 * np.sqrt(co.var(ddof=1) / chains + cg.var(ddof=1) / chains)
    mu_true = d["mu1"] * d["z"] + d["mu0"] * (1 - d["z"])
    assert np.corrcoef(mg.mean(axis=0), mu_true)[0, 1] >= (0.6 if binary else 0.95)


def _ihdp_pair(n=400, num_trees=9, seed=4321, warmup=6, iter_=11):
    from stan4bart_b200.frontend import ihdp_problem
    pr = ihdp_problem(n)
    cfg = bart_config(n, 25, n_test=n, num_trees=num_trees, is_binary=False, seed=seed)
    ctl = stan_control(seed=seed + 1)
    kw = dict(warmup=warmup, iter_=iter_, keep_fits=True, sigma_init=pr["sigma_init"], bart_offset_init=pr["bart_offset_init"])
    return pr, cfg, ctl, kw


def test_config_d_ihdp_shaped_first_sweeps():
    """BASELINE config D shape (25 covariates: 6 continuous + 19 binary, site-level random intercept, treatment as the
    only fixed effect): first sweeps agree with the oracle step by step; binary covariates bin to two levels."""
    pr, cfg, ctl, kw = _ihdp_pair()
    o = O.OracleSampler(cfg, pr["y"], pr["x_bart"], pr["x_test"], pr["stan_data"], ctl, **kw)
    g = Sampler(cfg, pr["y"], pr["x_bart"], pr["x_test"], pr["stan_data"], ctl, **kw)
    K = 6
    ob, gb = o.bart(), g.bart()
    ob.set_trace(9 * K); gb.set_trace(9 * K)
    ro, rg = o.run(K, True), g.run(K, True)
    compare_traces(ob.trace(), gb.trace(), tol=1e-8)
    assert rel_err(ro["stan"], rg["stan"], scale=np.abs(ro["stan"]) + 1.0) <= 1e-8
    assert rel_err(ro["bart"]["train"], rg["bart"]["train"], scale=np.abs(ro["bart"]["train"]) + 1.0) <= 1e-8
    assert np.array_equal(ro["bart"]["varcount"], rg["bart"]["varcount"])


def test_chains_on_concurrent_host_threads_match_sequential_runs():
    """Config D runs several chains per GPU: every chain is driven by its own host thread on its own stream (the C ABI
    keeps stream and error state per thread), so one chain's host NUTS overlaps another chain's sweep kernel.  The draws
    must not depend on the interleaving."""
    import threading
    pr, cfg0, ctl0, kw = _ihdp_pair(n=600)

    def make(c):
        cfg = bart_config(600, 25, n_test=600, num_trees=9, is_binary=False, seed=100 + c)
        return Sampler(cfg, pr["y"], pr["x_bart"], pr["x_test"], pr["stan_data"], stan_control(seed=200 + c), **kw)

    seq = []
    for c in range(3):
        s = make(c)
        seq.append(s.run(6, True))
        del s
    out = [None] * 3

    def work(c):
        s = make(c)
        out[c] = s.run(6, True)

    th = [threading.Thread(target=work, args=(c,)) for c in range(3)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for c in range(3):
        assert out[c] is not None
        assert np.array_equal(seq[c]["stan"], out[c]["stan"])
        assert np.array_equal(seq[c]["bart"]["train"], out[c]["bart"]["train"])


def test_chains_batched_in_one_launch_through_the_gibbs_loop():
    """Config D as SURVEY.md 8e states it: the chains of one GPU with their BART sweeps batched into ONE launch per Gibbs iteration
    (k_sweep_batch, grid.y = chain), every chain on its own host thread for its NUTS.  Four chains joined in a BatchGroup must
    produce exactly the draws of the same chains run one by one with the synchronous sweep kernel, and the group must have
    launched once per iteration."""
    import threading
    from stan4bart_b200.sampler import BatchGroup
    pr, cfg0, ctl0, kw = _ihdp_pair(n=600)
    chains, iters = 4, 6

    def make(c):
        cfg = bart_config(600, 25, n_test=600, num_trees=9, is_binary=False, seed=100 + c, max_ctas=148 // chains)
        return Sampler(cfg, pr["y"], pr["x_bart"], pr["x_test"], pr["stan_data"], stan_control(seed=200 + c), **kw)

    seq = []
    for c in range(chains):
        s = make(c)
        s.bart().set_pipeline(False)
        seq.append(s.run(iters, True))
        del s
    group = BatchGroup(chains)
    out, err = [None] * chains, [None] * chains
    samplers = [None] * chains

    def work(c):
        try:
            samplers[c] = make(c)
            samplers[c].set_batch_group(group)
            out[c] = samplers[c].run(iters, True)
        except Exception as e:      # pragma: no cover
            err[c] = e

    th = [threading.Thread(target=work, args=(c,)) for c in range(chains)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert all(e is None for e in err), err
    assert group.launches() == iters
    for c in range(chains):
        assert np.array_equal(seq[c]["stan"], out[c]["stan"])
        assert np.array_equal(seq[c]["bart"]["train"], out[c]["bart"]["train"])
        assert np.array_equal(seq[c]["bart"]["test"], out[c]["bart"]["test"])
        assert np.array_equal(seq[c]["bart"]["varcount"], out[c]["bart"]["varcount"])
    for s in samplers:
        s.set_batch_group(None)
    del samplers


@pytest.mark.parametrize("binary", [False, True])
def test_host_boundary_path_gives_the_same_draws(binary):
    """bench.py's `e2e` leg: results written into caller-provided host buffers (on a second stream, overlapping the next
    Stan block) and every N-vector of the sweep round-tripped through pinned host memory like the reference's host
    vectors.  The draws must be exactly those of the device-resident path."""
    _, g1, pr = make_pair(n=1201, binary=binary, num_trees=9, keep_fits=False)
    _, g2, _ = make_pair(n=1201, binary=binary, num_trees=9, keep_fits=False)
    h2d, d2h = g2.set_host_plumbing(True)
    assert h2d == d2h == 8 * 1201 * (3 if binary else 2)
    stan = np.zeros(g2.num_pars)
    train = np.zeros(1201)
    test = np.zeros(1201)
    for k in range(4):
        r1 = g1.run(3, k < 2)
        g2.run_into(3, k < 2, stan=stan.ctypes.data, train=train.ctypes.data, test=test.ctypes.data)
        assert np.array_equal(r1["stan"][:, -1], stan)
        assert np.array_equal(r1["bart"]["train"][:, -1], train)
        assert np.array_equal(r1["bart"]["test"][:, -1], test)
    g2.set_host_plumbing(False)
    r1, r2 = g1.run(2, False), g2.run(2, False)
    assert np.array_equal(r1["stan"], r2["stan"]) and np.array_equal(r1["bart"]["train"], r2["bart"]["train"])


@pytest.mark.parametrize("binary", [False, True])
def test_keep_trees_stored_draws_predict_their_own_training_fits(binary):
    """keepTrees (the reference's strongest self-consistency check, tests/testthat/test-01-continuous.R:212-254):
    predicting from the stored trees of draw s on the training design reproduces the stored training fit of draw s, the
    flattened tree table of the last stored draw is the live sampler's, and the store refuses to overflow."""
    from stan4bart_b200._lib import S4BError
    _, g, pr = make_pair(n=700, binary=binary, num_trees=9, warmup=8, iter_=14)
    g.run(8, True)                                  # warm-up draws are not stored
    g.disengage_adaptation()
    b = g.bart()
    b.set_keep_trees(6)
    r = g.run(6, False)
    assert b.num_stored() == 6
    pred = b.predict_stored(pr["x_bart"])
    assert pred.shape == (700, 6)
    assert rel_err(pred, r["bart"]["train"], scale=np.abs(r["bart"]["train"]) + 1.0) <= 1e-10
    # a different design goes through the same trees: stored draw 5 == the live sampler's current state
    xnew = np.asfortranarray(np.random.default_rng(0).random((50, 9)))
    assert rel_err(b.predict_stored(xnew, first=5, count=1)[:, 0], g.predict_bart(xnew), scale=1.0) <= 1e-12
    live, last = b.trees(), b.stored_trees(5)
    for key in ("tree", "n", "var", "value"):
        assert np.array_equal(live[key], last[key]), key
    assert not np.array_equal(b.stored_trees(0)["value"], last["value"])
    with pytest.raises(S4BError):
        g.run(1, False)                             # a seventh draw does not fit the store
    # exportBARTState / createStoredBARTSampler: the blob predicts on its own, also after the live sampler is gone
    from stan4bart_b200.sampler import StoredBart
    blob = b.export_stored()
    st = StoredBart(blob)
    assert st.count() == 6
    assert np.array_equal(st.predict(xnew), b.predict_stored(xnew))
    assert np.array_equal(st.predict(pr["x_bart"], first=2, count=3), pred[:, 2:5])
    with pytest.raises(S4BError):
        StoredBart(blob[:100])
    b.set_keep_trees(0)
    g.run(1, False)
    assert b.num_stored() == 0


@pytest.mark.parametrize("offset_type", [0, 1, 2, 3, 4])
@pytest.mark.parametrize("binary", [False, True])
def test_user_offset_and_offset_types(offset_type, binary):
    """stan4bart(offset = ...) and its offset_type variants (init.cpp:84-88, :236-252, :762-795, :831-839): default (added to
    both halves), fixef / ranef (stands in for that part of the parametric mean), bart (replaces the BART fit Stan sees),
    parametric (replaces the Stan part BART sees).  First sweeps against the oracle, step by step."""
    n, T, K = 400, 9, 6
    pr = friedman_problem(n, binary=binary)
    sd = pr["stan_data"]
    user = 0.4 * np.sin(np.arange(n) * 0.05) + 0.1
    cfg = bart_config(n, 9, n_test=n, num_trees=T, is_binary=binary, seed=321)
    ctl = stan_control(seed=654)
    kw = dict(warmup=K, iter_=2 * K, keep_fits=True, sigma_init=pr["sigma_init"], bart_offset_init=pr["bart_offset_init"],
              offset=user, offset_type=offset_type)
    o = O.OracleSampler(cfg, pr["y"], pr["x_bart"], pr["x_test"], sd, ctl, **kw)
    g = Sampler(cfg, pr["y"], pr["x_bart"], pr["x_test"], sd, ctl, **kw)
    ob, gb = o.bart(), g.bart()
    ob.set_trace(T * K); gb.set_trace(T * K)
    ro, rg = o.run(K, True), g.run(K, True)
    compare_traces(ob.trace(), gb.trace(), tol=1e-8)
    assert rel_err(ro["stan"], rg["stan"], scale=np.abs(ro["stan"]) + 1.0) <= 1e-8
    assert rel_err(ro["bart"]["train"], rg["bart"]["train"], scale=np.abs(ro["bart"]["train"]) + 1.0) <= 1e-8
    # the offset changes the chain (it is not silently ignored)
    base = Sampler(cfg, pr["y"], pr["x_bart"], pr["x_test"], sd, ctl, warmup=K, iter_=2 * K, keep_fits=True, sigma_init=pr["sigma_init"],
                   bart_offset_init=pr["bart_offset_init"]).run(K, True)
    assert not np.allclose(base["bart"]["train"], rg["bart"]["train"])


def test_iteration_callback_sees_each_draw():
    """The per-iteration callback of stan4bart_run (init.cpp:849-911, tests/testthat/test-11-callback.R): called after every
    iteration with that iteration's Stan row and fits; a truthy return stops the run."""
    from stan4bart_b200._lib import S4BError
    _, g, pr = make_pair(n=300, num_trees=7, warmup=4, iter_=9)
    seen = []
    g.set_callback(lambda it, stan, train, test: seen.append((it, stan.copy(), train.copy(), test.copy())) and False)
    r = g.run(4, True)
    assert [s[0] for s in seen] == [0, 1, 2, 3]
    for k, (_, stan, train, test) in enumerate(seen):
        assert np.array_equal(stan, r["stan"][:, k]) and np.array_equal(train, r["bart"]["train"][:, k])
        assert np.array_equal(test, r["bart"]["test"][:, k])
    g.set_callback(lambda it, stan, train, test: it == 1)
    with pytest.raises(S4BError):
        g.run(4, True)
    g.set_callback(None)
    g.run(2, True)


def test_three_coefficient_block_first_sweeps():
    """`(1 + X4 + X5 | g.1)`: a grouping term with three coefficients (onion rows through z_T) through the whole Gibbs
    sweep, step by step against the oracle."""
    from stan4bart_b200.frontend import build_stan_data, friedman_data, init_fit
    n, nt, seed = 300, 9, 777
    d = friedman_data(n, ranef=True, causal=True, binary=False, seed=3)
    x = d["x"]
    x_bart = np.asfortranarray(x[:, [0, 1, 2, 5, 6, 7, 8, 9]])
    ones = np.ones(n)
    terms = [(d["g1"], np.column_stack([ones, x[:, 3], x[:, 4]])), (d["g2"], ones.reshape(n, 1))]
    sd = build_stan_data(np.column_stack([x[:, 3], d["z"]]), d["y"], terms, is_binary=False)
    offset_init, sigma_init = init_fit(sd, False)
    cfg = bart_config(n, 8, n_test=n, num_trees=nt, is_binary=False, seed=seed)
    ctl = stan_control(seed=seed + 1)
    kw = dict(warmup=6, iter_=8, keep_fits=True, sigma_init=sigma_init, bart_offset_init=offset_init)
    y = np.ascontiguousarray(d["y"])
    o = O.OracleSampler(cfg, y, x_bart, x_bart.copy(order="F"), sd, ctl, **kw)
    g = Sampler(cfg, y, x_bart, x_bart.copy(order="F"), sd, ctl, **kw)
    K = 6
    ob, gb = o.bart(), g.bart()
    ob.set_trace(nt * K); gb.set_trace(nt * K)
    ro, rg = o.run(K, True), g.run(K, True)
    compare_traces(ob.trace(), gb.trace(), tol=1e-8)
    assert ro["stan"].shape == rg["stan"].shape
    assert rel_err(ro["stan"], rg["stan"], scale=np.abs(ro["stan"]) + 1.0) <= 1e-8
    assert rel_err(ro["bart"]["train"], rg["bart"]["train"], scale=np.abs(ro["bart"]["train"]) + 1.0) <= 1e-8
    names = sd.param_names()
    assert sum(nm.startswith("z_T") for nm in names) == 2


@pytest.mark.parametrize("binary", [False, True])
def test_observation_weights_first_sweeps(binary):
    """`stan4bart(weights = ...)` (R/stan4bart_fit.R:255-262, :449): the weights reach both halves -- dbarts data weights in
    the tree statistics and `weights` of continuous.stan:358-366 -- and the sweeps agree with the oracle step by step."""
    n, nt, seed, K = 500, 9, 31, 6
    pr = friedman_problem(n, binary=binary)
    rng = np.random.default_rng(17)
    wt = rng.gamma(3.0, 1.0 / 3.0, n)
    sd = pr["stan_data"]
    sd.weights = wt
    cfg = bart_config(n, 9, n_test=n, num_trees=nt, is_binary=binary, seed=seed, weights=wt)
    ctl = stan_control(seed=seed + 1)
    kw = dict(warmup=K, iter_=8, keep_fits=True, sigma_init=pr["sigma_init"], bart_offset_init=pr["bart_offset_init"])
    o = O.OracleSampler(cfg, pr["y"], pr["x_bart"], pr["x_test"], sd, ctl, **kw)
    g = Sampler(cfg, pr["y"], pr["x_bart"], pr["x_test"], sd, ctl, **kw)
    ob, gb = o.bart(), g.bart()
    ob.set_trace(nt * K); gb.set_trace(nt * K)
    ro, rg = o.run(K, True), g.run(K, True)
    compare_traces(ob.trace(), gb.trace(), tol=1e-8, ll_difference_only=True)
    assert rel_err(ro["stan"], rg["stan"], scale=np.abs(ro["stan"]) + 1.0) <= 1e-8
    assert rel_err(ro["bart"]["train"], rg["bart"]["train"], scale=np.abs(ro["bart"]["train"]) + 1.0) <= 1e-8
    assert rel_err(ro["bart"]["test"], rg["bart"]["test"], scale=np.abs(ro["bart"]["test"]) + 1.0) <= 1e-8
    assert np.array_equal(ro["bart"]["varcount"], rg["bart"]["varcount"])
    # and they matter: the unweighted chain from the same seeds differs
    sd.weights = None
    cfg_u = bart_config(n, 9, n_test=n, num_trees=nt, is_binary=binary, seed=seed)
    u = Sampler(cfg_u, pr["y"], pr["x_bart"], pr["x_test"], sd, ctl, **kw)
    ru = u.run(K, True)
    assert not np.allclose(ru["stan"], rg["stan"])


@pytest.mark.parametrize("prior_dist", [3, 6, 7])
def test_shrinkage_coefficient_priors_first_sweeps(prior_dist):
    """`stan_args = list(prior = hs() / lasso() / product_normal())` (R/stan4bart_fit.R:129-142): the extra parameters of the
    prior travel through NUTS and the stored Stan rows; sweeps agree with the oracle step by step."""
    n, nt, seed, K = 300, 7, 55, 5
    pr = friedman_problem(n, binary=False)
    sd = pr["stan_data"]
    sd.prior_dist = prior_dist
    sd.prior_df = np.array([1.0, 3.0])
    sd.num_normals = np.array([2, 3], dtype=np.int32)
    cfg = bart_config(n, 9, n_test=n, num_trees=nt, seed=seed)
    ctl = stan_control(seed=seed + 1)
    kw = dict(warmup=K, iter_=8, keep_fits=True, sigma_init=pr["sigma_init"], bart_offset_init=pr["bart_offset_init"])
    o = O.OracleSampler(cfg, pr["y"], pr["x_bart"], pr["x_test"], sd, ctl, **kw)
    g = Sampler(cfg, pr["y"], pr["x_bart"], pr["x_test"], sd, ctl, **kw)
    ro, rg = o.run(K, True), g.run(K, True)
    assert ro["stan"].shape == rg["stan"].shape == (len(sd.param_names()), K)
    assert rel_err(ro["stan"], rg["stan"], scale=np.abs(ro["stan"]) + 1.0) <= 1e-7
    assert rel_err(ro["bart"]["train"], rg["bart"]["train"], scale=np.abs(ro["bart"]["train"]) + 1.0) <= 1e-7
    assert np.array_equal(ro["bart"]["varcount"], rg["bart"]["varcount"])


def test_modelled_k_through_the_gibbs_loop():
    """`bart_args = list(k = chi(1.25, Inf))`: the per-iteration k row of the BART results (src/bart_util.hpp:25)."""
    n, nt, seed, K = 300, 8, 91, 6
    pr = friedman_problem(n, binary=False)
    cfg = bart_config(n, 9, n_test=n, num_trees=nt, seed=seed, k_df=1.25)
    ctl = stan_control(seed=seed + 1)
    kw = dict(warmup=K, iter_=8, keep_fits=True, sigma_init=pr["sigma_init"], bart_offset_init=pr["bart_offset_init"])
    o = O.OracleSampler(cfg, pr["y"], pr["x_bart"], pr["x_test"], pr["stan_data"], ctl, **kw)
    g = Sampler(cfg, pr["y"], pr["x_bart"], pr["x_test"], pr["stan_data"], ctl, **kw)
    ro, rg = o.run(K, True), g.run(K, True)
    assert ro["bart"]["k"].shape == rg["bart"]["k"].shape == (K,)
    assert rel_err(ro["bart"]["k"], rg["bart"]["k"]) <= 1e-8
    assert len(set(rg["bart"]["k"])) == K
    assert rel_err(ro["stan"], rg["stan"], scale=np.abs(ro["stan"]) + 1.0) <= 1e-8
    assert rel_err(ro["bart"]["train"], rg["bart"]["train"], scale=np.abs(ro["bart"]["train"]) + 1.0) <= 1e-8
    # fixed k: the row is the configured constant
    cfg_f = bart_config(n, 9, n_test=n, num_trees=nt, seed=seed, k=3.0)
    f = Sampler(cfg_f, pr["y"], pr["x_bart"], pr["x_test"], pr["stan_data"], ctl, **kw)
    assert np.array_equal(f.run(2, True)["bart"]["k"], [3.0, 3.0])


@pytest.mark.parametrize("case", range(10))
def test_randomised_models_against_the_oracle(case):
    """Random model structures through the whole sweep: grouping terms with 1-4 coefficients, coefficient prior family,
    aux prior, response type, weights, user offset, thinning of both halves; a few sweeps step by step."""
    from stan4bart_b200.frontend import build_stan_data, init_fit
    rng = np.random.default_rng(7000 + case)
    n = int(rng.choice([120, 400, 1500]))
    binary = bool(rng.integers(0, 2))
    p_bart, Kf = int(rng.integers(2, 8)), int(rng.integers(0, 4))
    xb = np.asfortranarray(rng.random((n, p_bart)))
    Xf = rng.standard_normal((n, Kf))
    terms = []
    for _ in range(int(rng.integers(0, 3))):
        nc, nlev = int(rng.integers(1, 5)), int(rng.integers(2, 7))
        g = rng.integers(0, nlev, n); g[:nlev] = np.arange(nlev)
        terms.append((g, np.column_stack([np.ones(n)] + [rng.standard_normal(n) for _ in range(nc - 1)])))
    f = 3 * np.sin(3 * xb[:, 0]) + xb[:, 1] + (Xf @ rng.standard_normal(Kf) if Kf else 0.0)
    y = (rng.random(n) < 1 / (1 + np.exp(-f))).astype(float) if binary else f + rng.standard_normal(n)
    wt = rng.gamma(3.0, 1.0 / 3.0, n) if rng.random() < 0.4 else None
    sd = build_stan_data(Xf, y, terms, is_binary=binary, weights=wt)
    sd.prior_dist = int(rng.choice([0, 1, 1, 2, 5, 6, 7] + ([] if binary else [3, 4]))) if Kf else 1
    sd.prior_df = rng.uniform(1.0, 6.0, Kf)
    sd.num_normals = rng.integers(2, 4, Kf).astype(np.int32)
    if not binary:
        sd.prior_dist_for_aux = int(rng.integers(0, 4))
        sd.prior_mean_for_aux, sd.prior_df_for_aux = 0.3, 4.0
    offset_init, sigma_init = init_fit(sd, binary)
    nt = int(rng.integers(2, 9))
    cfg = bart_config(n, p_bart, n_test=n, num_trees=nt, is_binary=binary, seed=case, thin=int(rng.choice([1, 2])), weights=wt)
    ctl = stan_control(seed=case + 50, skip=int(rng.choice([1, 2])))
    kw = dict(warmup=5, iter_=7, keep_fits=True, sigma_init=sigma_init, bart_offset_init=offset_init)
    if rng.random() < 0.3:
        kw.update(offset=rng.standard_normal(n) * 0.2, offset_type=int(rng.integers(0, 5)))
    if Kf == 0 and not terms:
        # no parametric component at all: the reference's front end refuses such a formula (R/lme4_functions.R:205) and Stan's
        # step-size search throws on the parameter-free density (base_hmc.hpp:131-134); the sampler must fail the same way
        from stan4bart_b200._lib import S4BError
        with pytest.raises(S4BError, match="Posterior is improper"):
            Sampler(cfg, y, xb, xb.copy(order="F"), sd, ctl, **kw)
        return
    o = O.OracleSampler(cfg, y, xb, xb.copy(order="F"), sd, ctl, **kw)
    g = Sampler(cfg, y, xb, xb.copy(order="F"), sd, ctl, **kw)
    ro, rg = o.run(5, True), g.run(5, True)
    assert ro["stan"].shape == rg["stan"].shape == (len(sd.param_names()), 5)
    assert rel_err(ro["stan"], rg["stan"], scale=np.abs(ro["stan"]) + 1.0) <= 1e-7
    assert rel_err(ro["bart"]["train"], rg["bart"]["train"], scale=np.abs(ro["bart"]["train"]) + 1.0) <= 1e-7
    assert rel_err(ro["bart"]["test"], rg["bart"]["test"], scale=np.abs(ro["bart"]["test"]) + 1.0) <= 1e-7
    assert np.array_equal(ro["bart"]["varcount"], rg["bart"]["varcount"])


def test_many_grouping_levels_through_the_gibbs_loop():
    """A grouping factor with 400 levels and a random slope (q = 800): the column path of the GLMM data pass and the
    sweep-level expansion with the Gram matrix in pieces (X'WX, X'WZ dense, Z'WZ sparse)."""
    from stan4bart_b200.frontend import build_stan_data, init_fit
    rng = np.random.default_rng(3)
    n, levels, nt = 6000, 400, 6
    xb = np.asfortranarray(rng.random((n, 4)))
    g = rng.integers(0, levels, n); g[:levels] = np.arange(levels)
    Xf = rng.standard_normal((n, 1))
    b = rng.standard_normal(levels) * 0.5
    y = 3 * np.sin(3 * xb[:, 0]) + Xf[:, 0] + b[g] + rng.standard_normal(n)
    sd = build_stan_data(Xf, y, [(g, np.column_stack([np.ones(n), xb[:, 1]]))])
    offset_init, sigma_init = init_fit(sd, False)
    cfg = bart_config(n, 4, n_test=n, num_trees=nt, seed=8)
    ctl = stan_control(seed=9, max_treedepth=5)
    kw = dict(warmup=4, iter_=6, keep_fits=True, sigma_init=sigma_init, bart_offset_init=offset_init)
    o = O.OracleSampler(cfg, y, xb, xb.copy(order="F"), sd, ctl, **kw)
    s = Sampler(cfg, y, xb, xb.copy(order="F"), sd, ctl, **kw)
    assert s.glmm().mode() == 1
    ro, rg = o.run(4, True), s.run(4, True)
    assert rel_err(ro["stan"], rg["stan"], scale=np.abs(ro["stan"]) + 1.0) <= 1e-7
    assert rel_err(ro["bart"]["train"], rg["bart"]["train"], scale=np.abs(ro["bart"]["train"]) + 1.0) <= 1e-7


def test_headline_configuration_first_sweeps_match_the_oracle():
    """The configuration BASELINE.json's metric is quoted on and bench.py times -- binary probit Friedman, n = 1 000 000, 200 trees,
    n_test = n, the NQ = 4 register-resident sweep kernel without the parity trace, GLMM sweep-level expansion (mode 1) -- against the
    CPU oracle with the same seeds: every sweep's Stan row, train / test fits, trees (structures and node counts bit-exact).
    bench.py runs the same check before it times the CPU arm and reports it as `parity_at_config`."""
    import argparse

    import bench
    args = argparse.Namespace(n=1_000_000, trees=200, continuous=False, weighted=False)
    res = bench.parity_at_config(args, bench.make_problem(args), sweeps=2)
    assert res["status"] == "ok", res
    assert res["bart_sweep_mode"] == 2 and res["glmm_mode"] == 1
    assert max(res["max_rel_err"].values()) <= 1e-8
