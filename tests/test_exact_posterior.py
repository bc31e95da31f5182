"""External pin of the BART tree sampler: the exact posterior of a one-tree model on a tiny data set, enumerated by brute
force from the model (tests/exact_posterior.py: independent of dbarts, of oracle/ and of the CUDA path), against the
visit frequencies of long chains of (a) the CPU oracle and (b) the CUDA path.  A wrong prior, transition or likelihood
ratio in any of the four moves breaks detailed balance with respect to this distribution and shows up here; nobody's
recollection of birthDeathRule.cpp / changeRule.cpp / swapRule.cpp is involved.

Error bars come from independent chains (different seeds), so autocorrelation is accounted for."""
import numpy as np
import pytest

import exact_posterior as EP
import oracle_lib as O
from stan4bart_b200.structs import bart_config

NCUTS, BASE, POWER, K, NODE_SCALE = 3, 0.95, 2.0, 2.0, 0.5


def problem(p, amp, n=36, seed=5):
    """A weak step signal under heavy noise: the posterior then spreads over many tree structures."""
    rng = np.random.default_rng(seed)
    x = np.asfortranarray(rng.random((n, p)))
    f = amp * np.where(x[:, 0] > 0.45, 1.0, -0.6) + (amp * 0.9 * (x[:, p - 1] > 0.7) if p > 1 else 0.0)
    y = f + 0.9 * rng.standard_normal(n)
    return x, y


def exact(x, y, sigma, min_obs, weights=None, split_probs=None):
    lo, rng_ = float(y.min()), float(y.max() - y.min())
    ys = (y - lo) / rng_ - 0.5                                    # the sampler's rescaled response
    leaf_prec = (K * np.sqrt(1.0) / NODE_SCALE) ** 2              # one tree
    return EP.ExactPosterior(x, ys, sigma / rng_, leaf_prec, NCUTS, BASE, POWER, min_obs, weights, split_probs), lo, rng_


def run_chains(make, x, steps, chains, burn=200):
    """`make(seed)` -> a sampler with set_trace / run / trace; one tree, `thin` = steps per run() call."""
    bins = EP.bin_matrix(x, NCUTS)
    freqs, fits = [], []
    for c in range(chains):
        s = make(1000 + c)
        s.set_trace(steps + burn)
        s.run()
        tr = s.trace()
        assert len(tr) == steps + burn
        # the chain starts from the root-only tree; the burn-in steps are replayed for the state, not counted
        visits, fit_sum, m = EP.replay_trace(tr, bins, skip=burn)
        freqs.append({k: v / m for k, v in visits.items()})
        fits.append(fit_sum / m)
    return freqs, np.array(fits)


def check_against_exact(ex, freqs, fits, what):
    C = len(freqs)
    f = np.zeros((C, len(ex.keys)))
    for c, fr in enumerate(freqs):
        for k, v in fr.items():
            assert k in ex.index, f"{what}: the chain visited a tree outside the model's support: {k}"
            f[c, ex.index[k]] = v
    mean, se = f.mean(axis=0), f.std(axis=0, ddof=1) / np.sqrt(C)
    big = ex.prob >= 0.004
    z = (mean[big] - ex.prob[big]) / np.maximum(se[big], 1e-4)
    tv = 0.5 * np.abs(mean - ex.prob).sum()
    worst = int(np.argmax(np.abs(z)))
    assert np.abs(z).max() <= 4.5, (f"{what}: visit frequency of {np.array(ex.keys, dtype=object)[big][worst]} is {mean[big][worst]:.4f}, "
                                    f"exact {ex.prob[big][worst]:.4f} (z = {z[worst]:.1f}); total variation {tv:.4f}")
    assert tv <= 0.03, f"{what}: total variation distance {tv:.4f}"
    # the number of bottom nodes (coarse summary with small error bars)
    for L in np.unique(ex.num_leaves):
        sel = ex.num_leaves == L
        pe, fe = ex.prob[sel].sum(), f[:, sel].sum(axis=1)
        assert abs(fe.mean() - pe) <= 4.5 * max(fe.std(ddof=1) / np.sqrt(C), 1e-4) + 1e-3, f"{what}: P({L} bottom nodes) {fe.mean():.4f} vs exact {pe:.4f}"
    # posterior mean of the fit at every row: checks the leaf draws as well
    fm, fse = fits.mean(axis=0), fits.std(axis=0, ddof=1) / np.sqrt(C)
    zf = (fm - ex.expected_fit) / np.maximum(fse, 1e-5)
    assert np.abs(zf).max() <= 4.5, f"{what}: posterior mean fit off by z = {zf[np.argmax(np.abs(zf))]:.1f}"
    return tv


CASES = {
    # name: (p, min_obs, weighted, split_probs, move probabilities (birth/death, swap, change), signal amplitude)
    "one_variable": (1, 5, False, None, (0.5, 0.1, 0.4), 0.4),
    "two_variables": (2, 3, False, None, (0.5, 0.1, 0.4), 0.7),
    "two_variables_change_swap_heavy": (2, 3, False, None, (0.2, 0.3, 0.5), 0.7),
    "two_variables_weighted_split_probs": (2, 3, True, (0.7, 0.3), (0.5, 0.1, 0.4), 0.7),
}


def case_setup(name):
    p, min_obs, weighted, sp, moves, amp = CASES[name]
    x, y = problem(p, amp)
    sigma = 1.3
    wt = np.random.default_rng(11).gamma(4.0, 0.25, len(y)) if weighted else None
    ex, lo, rng_ = exact(x, y, sigma, min_obs, wt, sp)

    def cfg_for(seed, steps, **kw):
        return bart_config(len(y), p, num_trees=1, n_cuts=NCUTS, thin=steps, min_obs=min_obs, base=BASE, power=POWER, k=K, node_scale=NODE_SCALE,
                           seed=seed, birth_death_prob=moves[0], swap_prob=moves[1], change_prob=moves[2], split_probs=sp, weights=wt, **kw)
    return x, y, sigma, ex, cfg_for


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_chain_has_the_exact_posterior(name):
    x, y, sigma, ex, cfg_for = case_setup(name)
    assert len(ex.keys) > (5 if name == "one_variable" else 50) and ex.prob.max() < 0.6      # a spread-out posterior, not a point mass
    steps, chains, burn = 50000, 12, 200

    def make(seed):
        o = O.OracleBart(cfg_for(seed, steps + burn), y, x)
        o.set_sigma(sigma)
        return o
    freqs, fits = run_chains(make, x, steps, chains, burn)
    check_against_exact(ex, freqs, fits, "oracle " + name)


def test_the_check_detects_a_change_step_without_its_proposal_ratio():
    """Power of the check: with `change_symmetric` (prior x likelihood ratio only, the form remembered from dbarts / BayesTree)
    the chain's stationary law is NOT the model's posterior once the change step can switch variables, and this test sees it."""
    name = "two_variables_change_swap_heavy"
    x, y, sigma, ex, cfg_for = case_setup(name)
    steps, chains, burn = 50000, 12, 200

    def make(seed):
        o = O.OracleBart(cfg_for(seed, steps + burn, change_symmetric=True), y, x)
        o.set_sigma(sigma)
        return o
    freqs, fits = run_chains(make, x, steps, chains, burn)
    with pytest.raises(AssertionError):
        check_against_exact(ex, freqs, fits, "oracle, uncorrected change step")


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CASES))
def test_cuda_chain_has_the_exact_posterior(name):
    from stan4bart_b200.sampler import GpuBart
    x, y, sigma, ex, cfg_for = case_setup(name)
    steps, chains, burn = 40000, 8, 200

    def make(seed):
        g = GpuBart(cfg_for(seed, steps + burn), y, x)
        g.set_sigma(sigma)
        return g
    freqs, fits = run_chains(make, x, steps, chains, burn)
    check_against_exact(ex, freqs, fits, "cuda " + name)
