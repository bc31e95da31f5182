"""Host-side input construction: the pieces of the reference's R front end that fix the
hot path's *inputs* (they run once per fit and are otherwise out of scope).

  friedman_data      inst/common/friedmanData.R:1-126 (same distributional recipe; numpy
                     default_rng instead of R's set.seed(99) stream, which cannot be
                     reproduced without R)
  build_stan_data    R/stan4bart_fit.R:95-365 (default priors: normal(0, 2.5) autoscaled,
                     exponential(1) autoscaled aux, decov(1,1,1,1)); center_x
                     (R/rstanarm_functions.R:420-447); Z layout = lme4 mkReTrms order
                     (terms by decreasing #levels, level-major / coefficient-minor,
                     R/lme4_functions.R:463-471, :1026-1028) in CSR (extract_sparse_parts)
  init_fit           stands in for the lmer/glmer initial fit of R/stan4bart.R:125-186
                     (least squares on the centred fixed effects)
"""
import math

import numpy as np
from scipy.special import ndtr

from .structs import StanData


def friedman_data(n, ranef=True, causal=True, binary=False, seed=99, n_g1=5, n_g2=8):
    rng = np.random.default_rng(seed)
    x = rng.random((n, 10))
    sigma = 1.0
    mu_bart = 10.0 * np.round(np.sin(np.pi * x[:, 0] * x[:, 1]), 14) + 20.0 * (x[:, 2] - 0.5) ** 2 + 5.0 * x[:, 4]
    mu_fixef = 10.0 * x[:, 3]
    out = dict(x=x, sigma=sigma, mu_bart=mu_bart, mu_fixef=mu_fixef)
    mu = mu_bart + mu_fixef
    if ranef:
        g1 = rng.integers(0, n_g1, n)
        Sigma_b1 = np.array([[1.5 ** 2, 0.2], [0.2, 1.0]])
        b1 = rng.standard_normal((n_g1, 2)) @ np.linalg.cholesky(Sigma_b1).T
        g2 = rng.integers(0, n_g2, n)
        b2 = rng.standard_normal(n_g2) * math.sqrt(1.2)
        mu_ranef = b1[g1, 0] + x[:, 3] * b1[g1, 1] + b2[g2]
        mu = mu + mu_ranef
        out.update(g1=g1, g2=g2, b1=b1, b2=b2, mu_ranef=mu_ranef)
    if causal:
        tau = 5.0
        z = (rng.random(n) < 0.2).astype(np.float64)
        mu0, mu1 = mu, mu + tau
        if binary:
            both = np.concatenate([mu0, mu1])
            loc = both.mean()
            scale = both.std(ddof=1) / -1.0364333894937898   # qnorm(0.15)
            mu0 = (mu0 - loc) / scale
            mu1 = (mu1 - loc) / scale
            y0 = (rng.random(n) < ndtr(mu0)).astype(np.float64)
            y1 = (rng.random(n) < ndtr(mu1)).astype(np.float64)
        else:
            y0 = mu0 + rng.standard_normal(n) * sigma
            y1 = mu1 + rng.standard_normal(n) * sigma
        y = y1 * z + y0 * (1.0 - z)
        out.update(z=z, tau=tau, mu0=mu0, mu1=mu1, y0=y0, y1=y1, y=y)
    else:
        if binary:
            loc = mu.mean()
            scale = mu.std(ddof=1) / -1.0364333894937898
            mu = (mu - loc) / scale
            y = (rng.random(n) < ndtr(mu)).astype(np.float64)
        else:
            y = mu + rng.standard_normal(n) * sigma
        out.update(mu=mu, y=y)
    return out


def _ranef_terms_to_csr(n, terms):
    """terms: list of (level_index[n] int, covariates [n, p] with first column = 1 for an
    intercept).  Returns p, l, q, (w, v, u) with lme4's term ordering."""
    order = sorted(range(len(terms)), key=lambda i: -int(terms[i][0].max() + 1 if len(terms[i][0]) else 0))
    # stable for ties (sorted is stable)
    p, l, cols, vals = [], [], [], []
    col_base = 0
    for i in order:
        g, M = terms[i]
        g = np.asarray(g, dtype=np.int64)
        M = np.asarray(M, dtype=np.float64).reshape(n, -1)
        nlev = int(g.max()) + 1
        pc = M.shape[1]
        p.append(pc)
        l.append(nlev)
        for c in range(pc):
            cols.append(col_base + g * pc + c)
            vals.append(M[:, c])
        col_base += nlev * pc
    q = col_base
    if not cols:
        return [], [], 0, (np.zeros(0), np.zeros(0, np.int32), np.zeros(n + 1, np.int32)), order
    cols = np.stack(cols, axis=1)          # [n, nnz_per_row], ascending within a row by construction
    vals = np.stack(vals, axis=1)
    nnz_row = cols.shape[1]
    w = vals.reshape(-1)
    v = cols.reshape(-1).astype(np.int32)
    u = (np.arange(n + 1, dtype=np.int64) * nnz_row).astype(np.int32)
    return p, l, q, (w, v, u), order


def build_stan_data(X_fixed, y, ranef_terms, is_binary=False, prior_scale=2.5, autoscale=True,
                    decov=(1.0, 1.0, 1.0, 1.0), weights=None):
    """Mirror of the data.stan list built at R/stan4bart_fit.R:259-365 on the default path.  weights: observation weights
    (`has_weights` / `weights`); all-one weights are dropped as at R/stan4bart_fit.R:255."""
    X_fixed = np.asarray(X_fixed, dtype=np.float64)
    n = len(y)
    if X_fixed.ndim == 1:
        X_fixed = X_fixed.reshape(n, -1)
    xbar = X_fixed.mean(axis=0) if X_fixed.shape[1] else np.zeros(0)
    xtemp = X_fixed - xbar
    K = xtemp.shape[1]
    y = np.asarray(y, dtype=np.float64)
    p_scale = np.full(K, float(prior_scale))
    scale_aux = 1.0                                        # exponential(rate = 1) => scale 1
    if not is_binary:
        ss = y.std(ddof=1)
        if autoscale:
            p_scale = p_scale * ss
            scale_aux = scale_aux * ss
    if autoscale and K:
        sdx = np.array([1.0 if len(np.unique(xtemp[:, k])) == 1 else xtemp[:, k].std(ddof=1) for k in range(K)])
        p_scale = np.maximum(1e-12, p_scale / sdx)
    p, l, q, (w, v, u), order = _ranef_terms_to_csr(n, ranef_terms)
    reg, conc, shape, scale = decov
    t = len(p)
    len_conc = int(sum(pi for pi in p if pi > 1))
    len_reg = int(sum(1 for pi in p if pi > 1))
    sd = StanData(
        X=xtemp, y=y, is_binary=is_binary, prior_dist=1, prior_scale=p_scale, prior_mean=np.zeros(K),
        prior_dist_for_aux=0 if is_binary else 3,
        prior_scale_for_aux=0.0 if is_binary else scale_aux, prior_mean_for_aux=0.0, prior_df_for_aux=0.0 if is_binary else 1.0,
        p=p, l=l, shape=np.full(t, shape), scale=np.full(t, scale), concentration=np.full(len_conc, conc),
        regularization=np.full(len_reg, reg), w=w, v=v, u=u, q=q)
    sd.xbar = xbar
    if weights is not None and len(weights) > 0 and not np.all(np.asarray(weights) == 1):
        sd.weights = np.ascontiguousarray(weights, dtype=np.float64)
    sd.term_order = order
    return sd


def init_fit(stan_data, is_binary):
    """bart_offset_init / sigma_init stand-in (R/stan4bart.R:125-186 uses lmer/glmer)."""
    n = stan_data.N
    if is_binary:
        return np.zeros(n), 1.0
    A = np.column_stack([np.ones(n), stan_data.X])
    coef, *_ = np.linalg.lstsq(A, stan_data.y, rcond=None)
    fitted = A @ coef
    resid = stan_data.y - fitted
    dof = max(1, n - A.shape[1])
    return fitted, float(math.sqrt(float(resid @ resid) / dof))


def friedman_problem(n, binary=False, seed=99, n_g1=5, n_g2=8, with_test=True):
    """The README model `y ~ bart(. - g.1 - g.2 - X4 - z) + X4 + z + (1 + X4 | g.1) + (1 | g.2)`
    (readme.md:46-55) on Friedman data: returns the BART design (9 columns), the
    counterfactual test design, and the Stan data."""
    d = friedman_data(n, ranef=True, causal=True, binary=binary, seed=seed, n_g1=n_g1, n_g2=n_g2)
    x = d["x"]
    bart_cols = [0, 1, 2, 4, 5, 6, 7, 8, 9]
    x_bart = np.asfortranarray(x[:, bart_cols])
    x_test = np.asfortranarray(x_bart.copy()) if with_test else None   # treatment z is not a BART column
    X_fixed = np.column_stack([x[:, 3], d["z"]])
    ones = np.ones(n)
    terms = [(d["g1"], np.column_stack([ones, x[:, 3]])), (d["g2"], ones.reshape(n, 1))]
    sd = build_stan_data(X_fixed, d["y"], terms, is_binary=binary)
    offset_init, sigma_init = init_fit(sd, binary)
    return dict(data=d, x_bart=x_bart, x_test=x_test, stan_data=sd, y=np.ascontiguousarray(d["y"]),
                bart_offset_init=offset_init, sigma_init=sigma_init)


def shard_problem(prob, lo, hi):
    """Rows [lo, hi) of a problem built by `friedman_problem` (or laid out like it): what one rank of an
    observation-sharded chain holds.  The initial offset / sigma come from the whole-data fit."""
    out = dict(prob)
    out["x_bart"] = np.asfortranarray(prob["x_bart"][lo:hi])
    out["x_test"] = np.asfortranarray(prob["x_test"][lo:hi]) if prob.get("x_test") is not None else None
    out["stan_data"] = prob["stan_data"].rows(lo, hi)
    out["y"] = np.ascontiguousarray(prob["y"][lo:hi])
    out["bart_offset_init"] = np.ascontiguousarray(prob["bart_offset_init"][lo:hi])
    if prob.get("weights") is not None:
        out["weights"] = np.ascontiguousarray(prob["weights"][lo:hi])
    return out


def ihdp_problem(n, seed=7, n_sites=8, with_test=True):
    """BASELINE config D: an IHDP-shaped causal problem (/root/reference/ihdp/data.R:7-10: 6 continuous + 19 binary
    covariates; treatment z; site-level random intercept), scaled to n rows with synthetic covariates.  Model:
    `y ~ bart(. - site - z) + z + (1 | site)`; the counterfactual test design flips nothing BART sees, so it aliases
    the training design exactly as in the Friedman causal example."""
    rng = np.random.default_rng(seed)
    xc = rng.standard_normal((n, 6))
    xb = (rng.random((n, 19)) < rng.uniform(0.1, 0.9, 19)).astype(np.float64)
    x = np.column_stack([xc, xb])
    site = rng.integers(0, n_sites, n)
    b_site = rng.standard_normal(n_sites) * 0.8
    z = (rng.random(n) < ndtr(0.4 * xc[:, 0] - 0.3 * xb[:, 0] - 0.6)).astype(np.float64)
    mu0 = np.exp((xc[:, :3] + 0.5) @ np.array([0.3, 0.2, 0.1])) + xb[:, :5] @ np.array([0.5, -0.4, 0.3, 0.2, -0.1]) + b_site[site]
    y = mu0 + 4.0 * z + rng.standard_normal(n)
    x_bart = np.asfortranarray(x)
    x_test = np.asfortranarray(x_bart.copy()) if with_test else None
    terms = [(site, np.ones((n, 1)))]
    sd = build_stan_data(z.reshape(n, 1), y, terms, is_binary=False)
    offset_init, sigma_init = init_fit(sd, False)
    return dict(data=dict(x=x, z=z, site=site, y=y), x_bart=x_bart, x_test=x_test, stan_data=sd, y=np.ascontiguousarray(y),
                bart_offset_init=offset_init, sigma_init=sigma_init)
