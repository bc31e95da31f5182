"""The reference's own statistical acceptance bounds, on the CUDA path with the reference's defaults
(/root/reference/tests/testthat/test-01-continuous.R:127-168, "nonlinearities are estimated well"):

    80 / 20 train / test split of the n = 100 Friedman data, `y ~ bart(. - g.1 - g.2 - X4 - z) + X4 + z + (1 + X4 | g.1) + (1 | g.2)`,
    defaults (4 chains, 1000 warm-up + 1000 sampling iterations, 75 trees):
      test RMSE (in units of sd(y_train)) <= that of the linear multilevel fit,
      cor(indiv.bart, truth) >= 0.95, cor(indiv.ranef, truth) >= 0.68, cor(indiv.fixef, truth) >= 0.99.

The data come from the same generator recipe with numpy's generator (R's `set.seed(99)` stream cannot be reproduced
without R); lme4 is not available, so the linear competitor is least squares with the grouping structure as dummies
(and, more leniently for the competitor, plain least squares) -- both are linear in X1..X10, which is what the bound is about."""
import numpy as np
import pytest

from stan4bart_b200.frontend import build_stan_data, friedman_data, init_fit
from stan4bart_b200.sampler import Sampler
from stan4bart_b200.structs import bart_config, stan_control

def _acceptance(sampler_class):
    n, n_train = 100, 80
    d = friedman_data(n, ranef=True, causal=True, binary=False, seed=99)
    x, z, y = d["x"], d["z"], d["y"]
    tr, te = np.arange(n_train), np.arange(n_train, n)
    bart_cols = [0, 1, 2, 4, 5, 6, 7, 8, 9]
    ones = np.ones(n)
    terms = lambda rows: [(d["g1"][rows], np.column_stack([ones[rows], x[rows, 3]])), (d["g2"][rows], ones[rows].reshape(-1, 1))]
    X_fixed = np.column_stack([x[:, 3], z])
    sd = build_stan_data(X_fixed[tr], y[tr], terms(tr))
    offset_init, sigma_init = init_fit(sd, False)
    names = sd.param_names()
    ib = [names.index(f"beta.{k + 1}") for k in range(2)]
    ibb = [names.index(f"b.{k + 1}") for k in range(sd.q)]

    # Z of any rows in the training layout (terms ordered by decreasing number of levels: g.2 (8 levels) then g.1 (5 levels x 2))
    def zmat(rows):
        Z = np.zeros((len(rows), sd.q))
        Z[np.arange(len(rows)), d["g2"][rows]] = 1.0
        Z[np.arange(len(rows)), 8 + 2 * d["g1"][rows]] = 1.0
        Z[np.arange(len(rows)), 8 + 2 * d["g1"][rows] + 1] = x[rows, 3]
        return Z
    # the layout assumption above is the one build_stan_data produced
    assert list(sd.p) == [1, 2] and list(sd.l) == [8, 5]
    Xc = X_fixed - sd.xbar

    chains, warmup, iters, trees = 4, 1000, 1000, 75
    bart_tr, bart_te, fix, ran = [], [], [], []
    for c in range(chains):
        cfg = bart_config(n_train, 9, n_test=n - n_train, num_trees=trees, seed=1 + c)
        s = sampler_class(cfg, y[tr], np.asfortranarray(x[np.ix_(tr, bart_cols)]), np.asfortranarray(x[np.ix_(te, bart_cols)]), sd,
                    stan_control(seed=101 + c), warmup=warmup, iter_=warmup + iters, keep_fits=True, sigma_init=sigma_init,
                    bart_offset_init=offset_init)
        s.run(warmup, True)
        s.disengage_adaptation()
        r = s.run(iters, False)
        bart_tr.append(r["bart"]["train"]); bart_te.append(r["bart"]["test"])
        fix.append(r["stan"][ib]); ran.append(r["stan"][ibb])
    bart_tr, bart_te = np.concatenate(bart_tr, axis=1), np.concatenate(bart_te, axis=1)
    beta, b = np.concatenate(fix, axis=1), np.concatenate(ran, axis=1)

    indiv_bart = bart_tr.mean(axis=1)
    indiv_fixef = (Xc[tr] @ beta).mean(axis=1)
    indiv_ranef = (zmat(tr) @ b).mean(axis=1)
    fitted_test = (bart_te + Xc[te] @ beta + zmat(te) @ b).mean(axis=1)
    bart_rmse = np.sqrt(np.mean((y[te] - fitted_test) ** 2)) / y[tr].std(ddof=1)

    def linear_rmse(with_groups):
        cols = [ones, x, z[:, None]]
        if with_groups:
            g1 = np.eye(5)[d["g1"]]
            cols += [g1[:, 1:], g1 * x[:, [3]], np.eye(8)[d["g2"]][:, 1:]]
        A = np.column_stack(cols)
        coef, *_ = np.linalg.lstsq(A[tr], y[tr], rcond=None)
        return np.sqrt(np.mean((y[te] - A[te] @ coef) ** 2)) / y[tr].std(ddof=1)

    fixef_true = d["mu_fixef"] + d["tau"] * z
    assert bart_rmse <= linear_rmse(True) and bart_rmse <= linear_rmse(False), (bart_rmse, linear_rmse(True), linear_rmse(False))
    assert np.corrcoef(indiv_bart, d["mu_bart"][tr])[0, 1] >= 0.95
    assert np.corrcoef(indiv_ranef, d["mu_ranef"][tr])[0, 1] >= 0.68
    assert np.corrcoef(indiv_fixef, fixef_true[tr])[0, 1] >= 0.99


@pytest.mark.gpu
def test_reference_acceptance_bounds_with_defaults():
    _acceptance(Sampler)


def test_reference_acceptance_bounds_hold_for_the_oracle_too():
    import oracle_lib as O
    _acceptance(O.OracleSampler)
