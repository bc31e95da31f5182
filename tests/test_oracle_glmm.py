"""CPU: the GLMM oracle against the committed golden vectors (torch fp64 autograd transcription of
continuous.stan, tests/golden/make_glmm_golden.py) and against finite differences."""
import os

import numpy as np
import pytest

import oracle_lib as O
from common import golden_case, golden_cases, load_glmm_case, rel_err


@pytest.mark.parametrize("path", golden_cases(), ids=lambda p: os.path.basename(p)[5:-5])
def test_oracle_matches_golden(path):
    sd, c = load_glmm_case(path)
    m = O.OracleGlmm(sd)
    m.set_offset(np.asarray(c["offset"]))
    assert m.d == len(c["q"][0])
    for q, lp, grad, wa in zip(c["q"], c["lp"], c["grad"], c["write_array"]):
        lp_o, g_o, status = m.log_prob_grad(np.asarray(q))
        assert status == 0
        assert abs(lp_o - lp) <= 1e-10 * abs(lp)
        assert rel_err(g_o, grad, scale=np.abs(grad) + 1e-8 * np.max(np.abs(grad))) <= 1e-10
        assert rel_err(m.write_array(np.asarray(q)), wa) <= 1e-12


def test_gradient_matches_finite_differences():
    sd, c = load_glmm_case(golden_case("friedman_like"))
    m = O.OracleGlmm(sd)
    m.set_offset(np.asarray(c["offset"]))
    q = np.asarray(c["q"][0])
    lp0, g, _ = m.log_prob_grad(q)
    for i in range(len(q)):
        h = 1e-6
        qp, qm = q.copy(), q.copy()
        qp[i] += h; qm[i] -= h
        fd = (m.log_prob_grad(qp)[0] - m.log_prob_grad(qm)[0]) / (2 * h)
        assert abs(fd - g[i]) <= 1e-5 * max(1.0, abs(g[i]))


def test_parametric_mean_and_aux():
    sd, c = load_glmm_case(golden_case("friedman_like"))
    m = O.OracleGlmm(sd)
    q = np.asarray(c["q"][0])
    wa = m.write_array(q)
    names = sd.param_names()[7:]
    beta = wa[[names.index(f"beta.{k + 1}") for k in range(sd.K)]]
    b = wa[[names.index(f"b.{k + 1}") for k in range(sd.q)]]
    # dense Z from the CSR parts
    Z = np.zeros((sd.N, sd.q))
    for i in range(sd.N):
        for k in range(sd.u[i], sd.u[i + 1]):
            Z[i, sd.v[k]] += sd.w[k]
    want = sd.X @ beta + Z @ b
    assert rel_err(m.parametric_mean(wa), want, scale=np.abs(want) + 1.0) <= 1e-13
    assert rel_err(m.parametric_mean(wa, fixed=True, random=False), sd.X @ beta, scale=1.0) <= 1e-13
    assert rel_err(m.parametric_mean(wa, fixed=False, random=True), Z @ b, scale=1.0) <= 1e-13
    assert O.lib().or_glmm_get_aux(m.h, O.dptr(wa)) == wa[names.index("aux.1")]


def test_non_finite_maps_to_error_status():
    sd, c = load_glmm_case(golden_case("friedman_like"))
    m = O.OracleGlmm(sd)
    q = np.asarray(c["q"][0]).copy()
    q[-1] = 800.0          # aux overflows
    assert m.log_prob_grad(q)[2] != 0


def test_unsupported_branches_are_rejected():
    from stan4bart_b200.frontend import build_stan_data
    rng = np.random.default_rng(0)
    N = 30
    g = rng.integers(0, 3, N)
    M = np.column_stack([np.ones(N), rng.random(N)])
    sd = build_stan_data(rng.random((N, 1)), rng.standard_normal(N), [(g, M)])
    import ctypes as C
    sd.prior_dist = 8                                   # no such prior family
    s = sd.struct()
    assert not O.lib().or_glmm_create(C.byref(s))
    # the horseshoe priors read the error scale aux[1] (continuous.stan:300, :304): undefined for a binary response
    sb = build_stan_data(rng.random((N, 1)), (rng.random(N) < 0.5).astype(float), [(g, M)], is_binary=True)
    sb.prior_dist = 3
    s = sb.struct()
    assert not O.lib().or_glmm_create(C.byref(s))


@pytest.mark.parametrize("prior_dist", [2, 3, 4, 5, 6, 7])
def test_coefficient_priors_gradient_by_finite_differences(prior_dist):
    """student_t / hs / hs_plus / laplace / lasso / product_normal coefficient priors: gradient against central differences
    of the log density (the goldens pin the values; this covers other sizes, K = 3)."""
    from stan4bart_b200.frontend import build_stan_data
    rng = np.random.default_rng(prior_dist)
    N = 80
    g = rng.integers(0, 4, N)
    sd = build_stan_data(rng.standard_normal((N, 3)), rng.standard_normal(N), [(g, np.ones((N, 1)))])
    sd.prior_dist = prior_dist
    sd.prior_df = np.array([3.0, 1.0, 5.0])
    sd.num_normals = np.array([2, 3, 2], dtype=np.int32)
    m = O.OracleGlmm(sd)
    assert m.d == sd.num_params
    m.set_offset(rng.standard_normal(N) * 0.3)
    q = rng.uniform(-0.7, 0.7, m.d)
    lp, grad, st = m.log_prob_grad(q)
    assert st == 0
    h = 1e-6
    for i in range(m.d):
        qp, qm = q.copy(), q.copy()
        qp[i] += h; qm[i] -= h
        fd = (m.log_prob_grad(qp)[0] - m.log_prob_grad(qm)[0]) / (2 * h)
        assert abs(fd - grad[i]) <= 2e-5 * max(1.0, abs(grad[i])), (i, fd, grad[i])
    assert len(sd.param_names()) == 7 + m.nc == 7 + sd.num_constrained


def test_onion_blocks_gradient_by_finite_differences():
    """Ranef blocks with more than two coefficients (scaled onion rows, z_T): the hand-derived / complex-step gradient against
    central differences of the log density, including the surplus z_T element that the Stan program declares but never uses."""
    sd, c = load_glmm_case(golden_case("four_and_three_binary"))
    assert sd.len_z_T == 8 and list(sd.p) == [4, 3, 2] or sd.len_z_T == 8
    m = O.OracleGlmm(sd)
    m.set_offset(np.asarray(c["offset"]))
    q = np.asarray(c["q"][1])
    lp, g, st = m.log_prob_grad(q)
    assert st == 0
    for k in range(len(q)):
        h = 1e-6
        qp, qm = q.copy(), q.copy()
        qp[k] += h; qm[k] -= h
        fd = (m.log_prob_grad(qp)[0] - m.log_prob_grad(qm)[0]) / (2 * h)
        assert abs(fd - g[k]) <= 1e-5 * max(1.0, abs(g[k])), (k, fd, g[k])
