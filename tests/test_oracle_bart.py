"""CPU: invariants of the BART oracle taken from the reference's own test suite
(tests/testthat/test-05-rng.R determinism, test-01-continuous.R:94-101 varcount, :212-254
predict == stored fits, :160-167 recovery bounds) plus RNG known answers."""
import numpy as np
import pytest

import oracle_lib as O
from common import bart_problem, rel_err
from stan4bart_b200.structs import bart_config


def test_philox_known_answers_and_uniform_range():
    # Random123 kat_vectors: philox4x32-10, counter 0 key 0 -> 6627e8d5 e169c58d ...
    u = np.zeros(1)
    O.lib().or_rng_uniforms(0, 0, 0, 1, O.dptr(u))
    o0, o1 = 0x6627e8d5, 0xe169c58d
    want = (((o0 >> 6) << 26 | (o1 >> 6)) + 0.5) * 2.0 ** -52
    assert u[0] == want
    v = np.zeros(200000)
    O.lib().or_rng_uniforms(123, 0, 0, len(v), O.dptr(v))
    assert 0.0 < v.min() and v.max() < 1.0
    assert abs(v.mean() - 0.5) < 4 / np.sqrt(12 * len(v))


def test_qnorm_as241():
    from scipy.stats import norm
    for p in [1e-300, 1e-20, 1e-12, 1e-8, 0.001, 0.075, 0.3, 0.5, 0.9, 1 - 1e-9]:
        got, want = O.lib().or_rng_qnorm(p), norm.ppf(p)
        assert abs(got - want) <= 1e-14 * max(1.0, abs(want))


def test_truncated_normal_moments():
    from scipy.stats import truncnorm
    for mean, positive in [(0.3, 1), (-2.5, 1), (1.0, 0), (-0.2, 0), (4.0, 1)]:
        z = np.array([O.lib().or_rng_truncnorm(9, i, 3, mean, positive) for i in range(20000)])
        assert np.all(z > 0) if positive else np.all(z < 0)
        a, b = (-mean, np.inf) if positive else (-np.inf, -mean)
        want_m, want_s = truncnorm.mean(a, b, loc=mean), truncnorm.std(a, b, loc=mean)
        assert abs(z.mean() - want_m) < 5 * want_s / np.sqrt(len(z))


def test_truncated_normal_law_including_far_tails():
    """The one-draw inversion x = m - qnorm(u Phi(m)) has exactly the law of N(m, 1) | x > 0: Kolmogorov-Smirnov against
    scipy for means from deep inside the truncated region to far outside it."""
    from scipy.stats import kstest, truncnorm
    for mean, positive in [(0.0, 1), (2.0, 1), (-3.0, 1), (-8.0, 1), (8.0, 1), (6.0, 0), (-6.0, 0)]:
        z = np.array([O.lib().or_rng_truncnorm(123, i, 1, mean, positive) for i in range(20000)])
        assert np.all(np.isfinite(z))
        a, b = (-mean, np.inf) if positive else (-np.inf, -mean)
        stat, pval = kstest(z, truncnorm(a, b, loc=mean).cdf)
        assert pval > 1e-3, (mean, positive, stat, pval)


def make(n=300, p=4, n_test=40, num_trees=15, seed=4, binary=False, **kw):
    x, y, xt = bart_problem(n, p, n_test, binary)
    cfg = bart_config(n, p, n_test=n_test, num_trees=num_trees, seed=seed, is_binary=binary, **kw)
    return O.OracleBart(cfg, y, x, xt), x, y, xt


def test_same_seed_same_draws_different_seed_differs():
    a, *_ = make(seed=12345)
    b, *_ = make(seed=12345)
    c, *_ = make(seed=12346)
    for f in (a, b, c):
        f.sample_trees_from_prior()
    ra = [a.run()["train"] for _ in range(5)]
    rb = [b.run()["train"] for _ in range(5)]
    rc = [c.run()["train"] for _ in range(5)]
    assert all(np.array_equal(u, v) for u, v in zip(ra, rb))
    assert not np.array_equal(ra[-1], rc[-1])


def test_record_then_replay_is_identical():
    a, *_ = make(seed=7)
    a.set_record(100000)
    a.sample_trees_from_prior()
    ra = [a.run()["train"] for _ in range(4)]
    b, *_ = make(seed=999)           # different seed: everything must come from the tape
    b.set_tape(a.record())
    b.sample_trees_from_prior()
    rb = [b.run()["train"] for _ in range(4)]
    assert all(np.array_equal(u, v) for u, v in zip(ra, rb))


def test_predict_equals_training_fit_and_varcount_and_partition():
    f, x, y, xt = make()
    f.sample_trees_from_prior()
    for _ in range(10):
        r = f.run()
    assert rel_err(f.predict(x), r["train"], scale=np.abs(r["train"]) + 1.0) <= 1e-12
    assert rel_err(f.predict(xt), r["test"], scale=np.abs(r["test"]) + 1.0) <= 1e-12
    tr = f.trees()
    internal = tr["var"][tr["var"] >= 0]
    assert np.array_equal(np.bincount(internal, minlength=f.p), r["varcount"])
    # every observation sits in exactly one leaf of every tree and leaf counts add up
    for t in range(15):
        heap, cnt, s, ss = f.leaf_stats(t)
        assert cnt.sum() == f.n
        assign = f.node_assignment(t)
        assert np.array_equal(np.sort(np.unique(assign)), np.sort(heap[cnt > 0]))
    # residual bookkeeping: yresc - totalFits
    rng_min, rng_max, rng_range = f.data_range()
    assert rel_err((r["train"] - rng_min) / rng_range - 0.5 + f.residual(), (y - rng_min) / rng_range - 0.5, scale=1.0) <= 1e-12


def test_offset_and_rescale_keep_fits_in_original_units():
    f, x, y, _ = make(n_test=0)
    f.sample_trees_from_prior()
    for _ in range(3):
        r0 = f.run()
    off = 0.5 * x[:, 0]
    fit_before = r0["train"].copy()
    f.set_offset(off, True)
    # fits of the trees themselves (minus the location shift of the new scale) are unchanged in original units
    mn, mx, rg = f.data_range()
    assert rg == pytest.approx((y - off).max() - (y - off).min())
    assert f.predict(x).std() == pytest.approx(fit_before.std(), rel=1e-9)


def test_binary_latents_have_the_sign_of_y():
    f, x, y, _ = make(binary=True, n_test=0)
    f.set_offset(np.zeros(len(y)), False)
    f.sample_trees_from_prior()
    for _ in range(3):
        f.run()
    z = f.latents()
    assert np.all(z[y > 0] > 0) and np.all(z[y == 0] < 0)


def test_recovers_signal():
    rng = np.random.default_rng(1)
    n = 400
    x = np.asfortranarray(rng.random((n, 5)))
    f_true = 10 * np.sin(np.pi * x[:, 0] * x[:, 1]) + 20 * (x[:, 2] - 0.5) ** 2 + 5 * x[:, 4]
    y = f_true + rng.standard_normal(n)
    fit = O.OracleBart(bart_config(n, 5, num_trees=50, seed=1), y, x)
    fit.set_sigma(1.0)
    acc = np.zeros(n)
    for it in range(150):
        r = fit.run()
        if it >= 50:
            acc += r["train"]
    assert np.corrcoef(acc / 100, f_true)[0, 1] >= 0.95


def test_trace_covers_all_move_types():
    f, *_ = make(n=800, num_trees=20, n_test=0)
    f.set_sigma(1.0)
    f.set_trace(20 * 40)
    for _ in range(40):
        f.run()
    tr = f.trace()
    assert len(tr) == 800
    for k in (0, 1, 2, 3):
        assert np.any((tr[:, 0] == k) & (tr[:, 4] == 1))
    # min_obs is respected by every accepted birth
    births = tr[(tr[:, 0] == 0) & (tr[:, 4] == 1)]
    assert np.all(births[:, 9] >= 5) and np.all(births[:, 10] >= 5)


def test_split_probs_prior_frequencies():
    """Trees drawn from the prior split on predictor j with probability proportional to split.probs (root rules: no
    predictor is exhausted there), and never on a predictor of probability zero."""
    x, y, _ = bart_problem(300, 5, 0, False)
    sp = np.array([1.0, 1.0, 2.0, 0.0, 4.0])
    cfg = bart_config(300, 5, num_trees=200, seed=17, split_probs=sp)
    o = O.OracleBart(cfg, y, x, None)
    counts = np.zeros(5)
    for _ in range(30):
        o.sample_trees_from_prior()
        tr = o.trees()
        roots = np.concatenate([[0], np.nonzero(np.diff(tr["tree"]))[0] + 1])
        rv = tr["var"][roots]
        counts += np.bincount(rv[rv >= 0], minlength=5)
    assert counts[3] == 0
    freq = counts / counts.sum()
    want = sp / sp.sum()
    assert np.all(np.abs(freq - want) < 4 * np.sqrt(want * (1 - want) / counts.sum()) + 1e-9)


def test_observation_weights_semantics():
    """Unit weights are the unweighted model (bit for bit); a constant weight c is sigma / sqrt(c); the leaf statistics of a
    weighted fit are the weighted mean with effective count sum w."""
    n, T = 400, 8
    x, y, xt = bart_problem(n, 4, 0, False, seed=2)

    def run(weights, sigma, sweeps=5):
        o = O.OracleBart(bart_config(n, 4, num_trees=T, seed=9, weights=weights), y, x, xt)
        o.set_sigma(sigma)
        o.sample_trees_from_prior()
        o.set_trace(T * sweeps)
        out = [o.run()["train"].copy() for _ in range(sweeps)]
        return out, o.trace()

    base, tr0 = run(None, 1.1)
    ones, tr1 = run(np.ones(n), 1.1)
    assert all(np.array_equal(a, b) for a, b in zip(base, ones)) and np.array_equal(tr0, tr1)
    c = 2.5
    scaled, _ = run(np.full(n, c), 1.1 * np.sqrt(c))
    assert max(rel_err(a, b, scale=np.abs(a) + 1.0) for a, b in zip(base, scaled)) <= 1e-9
    rng = np.random.default_rng(0)
    other, _ = run(rng.gamma(2.0, 0.5, n), 1.1)
    assert not np.array_equal(other[-1], base[-1])


def test_modelled_k_hyperprior():
    """bart_args k = chi(df, scale): k is redrawn after every sweep from its conjugate conditional; a prior that dominates
    pins it at sqrt(df) * scale, the improper chi(1.25, Inf) lets the leaf values speak."""
    n, T = 300, 10
    x, y, xt = bart_problem(n, 4, 0, False, seed=6)
    fixed = O.OracleBart(bart_config(n, 4, num_trees=T, seed=3), y, x, xt)
    assert fixed.k() == 2.0
    tight = O.OracleBart(bart_config(n, 4, num_trees=T, seed=3, k_df=1e6, k_scale=0.003), y, x, xt)
    free = O.OracleBart(bart_config(n, 4, num_trees=T, seed=3, k_df=1.25), y, x, xt)
    for o in (fixed, tight, free):
        o.set_sigma(1.0); o.sample_trees_from_prior()
    ks = []
    for s in range(12):
        fixed.run(); tight.run(); free.run()
        assert fixed.k() == 2.0
        assert abs(tight.k() - 3.0) < 0.02
        ks.append(free.k())
    assert all(np.isfinite(k) and k > 0 for k in ks) and len(set(ks)) == len(ks)


def test_cut_counts_per_predictor_in_the_oracle():
    n, T = 500, 10
    x, y, xt = bart_problem(n, 3, 0, False, seed=5)
    counts = np.array([1, 3, 50])
    o = O.OracleBart(bart_config(n, 3, num_trees=T, seed=3, n_cuts=counts), y, x, xt)
    o.set_sigma(1.0); o.sample_trees_from_prior()
    o.set_trace(T * 10)
    for s in range(10):
        o.run()
    tr = o.trace()
    rules = tr[(tr[:, 0] == 0) & (tr[:, 2] >= 0)]
    assert rules.shape[0] > 0 and np.all(rules[:, 3] < counts[rules[:, 2].astype(int)])
    with pytest.raises(Exception):
        O.OracleBart(bart_config(n, 3, num_trees=T, n_cuts=np.array([0, 3, 50])), y, x, xt)


def test_quantile_cut_points_in_the_oracle():
    """bart_args use.quantiles (R/stan4bart_fit.R:437-451 -> dbartsControl): cuts between the distinct sorted values.  Known answers of
    the rule as restated in oracle_bart.c:quantile_cuts: few distinct values => a cut in every gap; many => n.cuts of them, every
    (distinct / n.cuts) values starting half a step in; a constant predictor has no cut and is never split on."""
    n, T = 400, 8
    rng = np.random.default_rng(8)
    x = np.empty((n, 4))
    x[:, 0] = rng.integers(0, 2, n)                 # binary: one cut at 0.5
    x[:, 1] = rng.integers(0, 5, n) * 2.0           # 0, 2, .. 8: cuts 1, 3, 5, 7
    x[:, 2] = rng.permutation(n) + 0.0              # 400 distinct values, 7 cuts: step 57, offset 28 => between ranks 28|29, 85|86, ...
    x[:, 3] = 3.25                                  # constant: no cut
    x = np.asfortranarray(x)
    y = x[:, 0] + 0.3 * x[:, 1] + np.sin(x[:, 2] / 50.0) + 0.1 * rng.standard_normal(n)
    o = O.OracleBart(bart_config(n, 4, num_trees=T, seed=3, n_cuts=7, use_quantiles=True), y, x)
    o.set_sigma(0.5); o.sample_trees_from_prior()
    for _ in range(30):
        r = o.run()
    tr = o.trees()
    rules = tr["var"] >= 0
    assert rules.sum() > 0 and not np.any(tr["var"][rules] == 3)
    allowed = {0: {0.5}, 1: {1.0, 3.0, 5.0, 7.0}, 2: {28.5 + 57.0 * k for k in range(7)}}
    for v, c in zip(tr["var"][rules], tr["value"][rules]):
        assert float(c) in allowed[int(v)], (v, c)
    # every cut of the evenly-ranked predictor gets used over a long enough run of prior draws
    seen = set()
    for _ in range(60):
        o.sample_trees_from_prior()
        t2 = o.trees()
        seen |= {float(c) for v, c in zip(t2["var"], t2["value"]) if v == 2}
    assert seen == allowed[2]

