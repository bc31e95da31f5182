"""External pin of the Stan half (NUTS control + density + transforms): the exact posterior of a two-parameter model by quadrature
against long NUTS chains of (a) the CPU oracle and (b) the CUDA path.

The product's NUTS (csrc/nuts.cu) and the oracle's (oracle/oracle_nuts.c) both restate the vendored Stan headers, so their
agreement says nothing independent about the control logic.  Here the target's moments come from a third source: the torch
transcription of continuous.stan (tests/golden/make_glmm_golden.py, which shares no code with either) integrated on a grid over the
two unconstrained parameters.  A wrong Jacobian, prior, acceptance / multinomial weighting, U-turn rule or adaptation that biases the
chain shows up as a shifted mean or variance; error bars come from independent chains."""
import os
import sys

import numpy as np
import pytest
import torch

import oracle_lib as O
from stan4bart_b200.frontend import build_stan_data
from stan4bart_b200.structs import stan_control

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import make_glmm_golden as G  # noqa: E402


def model(binary):
    rng = np.random.default_rng(17)
    N = 24
    if binary:
        X = np.column_stack([rng.standard_normal(N), rng.random(N)])
        sd = build_stan_data(X, (rng.random(N) < 0.5).astype(float), [], is_binary=True)
        sd.y = 0.6 * X[:, 0] - 0.4 * X[:, 1] + rng.standard_normal(N)          # the probit latents Stan conditions on
    else:
        X = rng.standard_normal((N, 1))
        y = 0.8 * X[:, 0] + 0.9 * rng.standard_normal(N)
        sd = build_stan_data(X, y - y.mean(), [])
    offset = 0.3 * rng.standard_normal(N)
    assert sd.num_params == 2
    return sd, offset


def exact_moments(sd, offset, half_width=7.0, points=141):
    """E and E[.^2] of the two constrained parameters (beta.1 and aux.1, or beta.1 and beta.2) by quadrature over the unconstrained
    plane: a coarse grid finds the mode and the scale, the fine grid covers +- half_width posterior standard deviations."""
    def grid_eval(a0, a1, b0, b1, m):
        qa, qb = np.linspace(a0, a1, m), np.linspace(b0, b1, m)
        lp = np.empty((m, m)); f0 = np.empty((m, m)); f1 = np.empty((m, m))
        for i, u in enumerate(qa):
            for j, v in enumerate(qb):
                val, tp = G.log_prob(sd, torch.tensor([u, v]), offset, sd.y)
                lp[i, j] = float(val)
                f0[i, j] = float(tp["beta"][0])
                f1[i, j] = float(tp["aux"]) if not sd.is_binary else float(tp["beta"][1])
        w = np.exp(lp - lp.max()); w /= w.sum()
        return qa, qb, w, f0, f1
    qa, qb, w, _, _ = grid_eval(-12.0, 12.0, -12.0, 12.0, 49)
    ma, mb = (w.sum(1) * qa).sum(), (w.sum(0) * qb).sum()
    sa = max(np.sqrt((w.sum(1) * (qa - ma) ** 2).sum()), 0.25); sb = max(np.sqrt((w.sum(0) * (qb - mb) ** 2).sum()), 0.25)
    qa, qb, w, f0, f1 = grid_eval(ma - half_width * sa, ma + half_width * sa, mb - half_width * sb, mb + half_width * sb, points)
    edge = max(w[0].sum(), w[-1].sum(), w[:, 0].sum(), w[:, -1].sum())
    assert edge < 1e-9, "the grid does not cover the posterior"
    return np.array([(w * f0).sum(), (w * f1).sum(), (w * f0 ** 2).sum(), (w * f1 ** 2).sum()])


def chain_moments(make_sampler, names, binary, chains=6, warmup=300, draws=3000):
    cols = [names.index("beta.1"), names.index("beta.2" if binary else "aux.1")]
    out, acc, div = [], [], 0
    for c in range(chains):
        s = make_sampler(100 + c)
        for _ in range(warmup):
            s.run(True)
        s.disengage_adaptation()
        d = np.stack([s.run(False) for _ in range(draws)])
        x0, x1 = d[:, cols[0]], d[:, cols[1]]
        out.append([x0.mean(), x1.mean(), (x0 ** 2).mean(), (x1 ** 2).mean()])
        acc.append(d[:, names.index("accept_stat__")].mean()); div += int(d[:, names.index("divergent__")].sum())
    out = np.array(out)
    return out.mean(0), out.std(0, ddof=1) / np.sqrt(chains), float(np.mean(acc)), div


def check(exact, mean, se, acc, div, what):
    z = (mean - exact) / np.maximum(se, 1e-4 * (np.abs(exact) + 1e-3))
    assert np.all(np.abs(z) < 4.5), f"{what}: chain moments {mean} vs exact {exact} (z = {z})"
    assert np.all(se < 0.05 * (np.abs(exact) + 0.05)), f"{what}: the chains are too noisy to say anything ({se})"
    assert 0.6 < acc <= 1.0 and div == 0, (what, acc, div)


_EXACT = {}


def exact_for(binary):
    if binary not in _EXACT:
        sd, offset = model(binary)
        _EXACT[binary] = (sd, offset, exact_moments(sd, offset))
    return _EXACT[binary]


@pytest.mark.parametrize("binary", [False, True])
def test_oracle_nuts_has_the_exact_posterior(binary):
    sd, offset, exact = exact_for(binary)

    def make(seed):
        m = O.OracleGlmm(sd)
        m.set_offset(offset)
        s = O.OracleNuts(m, stan_control(seed=seed), num_warmup=300)
        s._keep = m
        return s
    mean, se, acc, div = chain_moments(make, sd.param_names(), binary)
    check(exact, mean, se, acc, div, "oracle")


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("binary", [False, True])
def test_cuda_nuts_has_the_exact_posterior(binary, mode):
    from stan4bart_b200.sampler import GlmmModel, StanSampler
    sd, offset, exact = exact_for(binary)

    def make(seed):
        m = GlmmModel(sd)
        m.set_mode(mode)
        m.set_offset(offset)
        return StanSampler(m, stan_control(seed=seed), num_warmup=300)
    chains, draws = (6, 3000) if mode == 1 else (4, 600)          # mode 0 launches a device pass per gradient evaluation
    mean, se, acc, div = chain_moments(make, sd.param_names(), binary, chains=chains, draws=draws)
    check(exact, mean, se, acc, div, "cuda path, glmm mode %d" % mode)
