"""Shared helpers for the parity tests."""
import numpy as np

from stan4bart_b200.frontend import friedman_problem
from stan4bart_b200.structs import bart_config, stan_control

REL_TOL = 1e-10   # north_star: leaf statistics, log-density and gradient agree to 1e-10 relative


def rel_err(a, b, scale=None):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if scale is None:
        scale = np.maximum(np.abs(a), np.abs(b))
    scale = np.maximum(scale, 1e-300)
    return float(np.max(np.abs(a - b) / scale)) if a.size else 0.0


def bart_problem(n=500, p=5, n_test=0, binary=False, seed=3):
    rng = np.random.default_rng(seed)
    x = np.asfortranarray(rng.random((n, p)))
    f = 10 * np.sin(np.pi * x[:, 0] * x[:, min(1, p - 1)]) + 5 * x[:, min(2, p - 1)]
    if binary:
        y = (rng.random(n) < 1.0 / (1.0 + np.exp(-(f - f.mean()) / 3.0))).astype(np.float64)
    else:
        y = f + rng.standard_normal(n)
    xt = np.asfortranarray(rng.random((n_test, p))) if n_test else None
    return x, y, xt


def compare_traces(tr_o, tr_g, tol=REL_TOL, ll_difference_only=False):
    """Integer fields exact, floating fields to `tol` (relative, with a small absolute floor for
    log-likelihood sums that can be near zero).  ll_difference_only (weighted fits): the device never forms the
    sum w r^2 term that is common to both sides of a Metropolis ratio, so only new - old log-likelihood is comparable."""
    assert tr_o.shape == tr_g.shape, (tr_o.shape, tr_g.shape)
    int_cols = [0, 1, 2, 3, 4, 8, 9, 10]
    for c in int_cols:
        bad = np.nonzero(tr_o[:, c] != tr_g[:, c])[0]
        assert bad.size == 0, f"trace column {c} differs first at step {bad[0]}: oracle {tr_o[bad[0]]} gpu {tr_g[bad[0]]}"
    # the MH ratio is exp(delta log-lik): compare it on the log scale, relative to the size of the log-likelihoods
    ra, rb = tr_o[:, 5], tr_g[:, 5]
    pos = (ra > 0) & (rb > 0) & np.isfinite(ra) & np.isfinite(rb)
    assert np.array_equal(ra[~pos], rb[~pos]), "zero / infinite / absent ratios differ"
    ll_scale = np.maximum(1.0, np.maximum(np.abs(tr_o[:, 6]), np.abs(tr_o[:, 7])))
    lerr = np.abs(np.log(ra[pos]) - np.log(rb[pos])) / ll_scale[pos]
    assert lerr.size == 0 or lerr.max() <= tol, f"log ratio differs by {lerr.max():.3e} (relative to |log-lik|)"
    if ll_difference_only:
        do, dg = tr_o[:, 7] - tr_o[:, 6], tr_g[:, 7] - tr_g[:, 6]
        fin = np.isfinite(do) & np.isfinite(dg)
        assert np.array_equal(fin, np.isfinite(do) | np.isfinite(dg))
        derr = np.abs(do[fin] - dg[fin]) / ll_scale[fin]
        assert derr.size == 0 or derr.max() <= tol, f"log-likelihood difference differs by {derr.max():.3e}"
    for c in ([] if ll_difference_only else [6, 7]) + list(range(11, tr_o.shape[1])):
        a, b = tr_o[:, c], tr_g[:, c]
        # floors: log-likelihood sums and leaf draws are differences that can cancel to ~0
        floor = 1.0 if c in (6, 7) else 1e-3
        scale = np.maximum(np.maximum(np.abs(a), np.abs(b)), floor)
        with np.errstate(invalid="ignore"):
            err = np.where(a == b, 0.0, np.abs(a - b) / scale)      # equal infinities (overflowing ratios) are equal
        k = int(np.argmax(err))
        assert err[k] <= tol, f"trace column {c} step {k}: oracle {a[k]!r} gpu {b[k]!r} rel {err[k]:.3e}"


def load_glmm_case(path):
    """Rebuild the StanData of a tests/golden/glmm_*.json case (see make_glmm_golden.py)."""
    import json

    from stan4bart_b200.frontend import build_stan_data
    with open(path) as f:
        c = json.load(f)
    N = c["N"]
    terms = [(np.asarray(g["g"]), np.asarray(g["M"]).reshape(N, -1)) for g in c["groups"]]
    sd = build_stan_data(np.asarray(c["X_fixed"]), np.asarray(c["y"]), terms, is_binary=c["binary"])
    if not c["binary"]:
        sd.prior_dist_for_aux = c["aux_prior"]
        sd.prior_mean_for_aux = c["prior_mean_for_aux"]
        sd.prior_df_for_aux = c["prior_df_for_aux"]
        sd.prior_scale_for_aux = c["prior_scale_for_aux"]
    sd.prior_dist = c["prior_dist"]
    sd.prior_scale = np.asarray(c["prior_scale"], dtype=np.float64)
    for key, val in c.get("coef", {}).items():
        setattr(sd, key, np.asarray(val) if isinstance(val, list) else val)
    if "weights" in c:
        sd.weights = np.asarray(c["weights"], dtype=np.float64)
    return sd, c


def golden_cases():
    import glob
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    return sorted(glob.glob(os.path.join(here, "golden", "glmm_*.json")))


def golden_case(name):
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    return os.path.join(here, "golden", f"glmm_{name}.json")
