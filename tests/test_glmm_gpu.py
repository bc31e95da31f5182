"""GPU parity: the device GLMM pass + host chain rule against the golden vectors and the oracle."""
import os

import numpy as np
import pytest

import oracle_lib as O
from common import REL_TOL, golden_case, golden_cases, load_glmm_case, rel_err
from stan4bart_b200.frontend import friedman_problem
from stan4bart_b200.sampler import GlmmModel

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("path", golden_cases(), ids=lambda p: os.path.basename(p)[5:-5])
def test_gpu_matches_golden(path, mode):
    sd, c = load_glmm_case(path)
    m = GlmmModel(sd)
    if sd.K + sd.q > 0:
        m.set_mode(mode)
    m.set_offset(np.asarray(c["offset"]))
    assert m.d == len(c["q"][0])
    for q, lp, grad, wa in zip(c["q"], c["lp"], c["grad"], c["write_array"]):
        lp_g, g_g, status = m.log_prob_grad(np.asarray(q))
        assert status == 0
        assert abs(lp_g - lp) <= REL_TOL * abs(lp)
        assert rel_err(g_g, grad, scale=np.abs(grad) + 1e-8 * np.max(np.abs(grad))) <= REL_TOL
        assert rel_err(m.write_array(np.asarray(q)), wa) <= 1e-12


@pytest.mark.parametrize("n,binary", [(2, False), (31, False), (257, True), (20000, False), (100003, True)])
def test_gpu_matches_oracle_on_friedman(n, binary):
    pr = friedman_problem(max(n, 1), binary=binary, seed=5)
    sd = pr["stan_data"]
    rng = np.random.default_rng(n)
    off = rng.standard_normal(sd.N)
    mo, mg = O.OracleGlmm(sd), GlmmModel(sd)
    mo.set_offset(off); mg.set_offset(off)
    if binary:
        z = rng.standard_normal(sd.N)
        mo.set_response(z); mg.set_response(z)
    for _ in range(3):
        q = rng.uniform(-1, 1, mo.d)
        lo, go, so = mo.log_prob_grad(q)
        lg, gg, sg = mg.log_prob_grad(q)
        assert so == sg == 0
        assert abs(lo - lg) <= REL_TOL * abs(lo)
        assert rel_err(go, gg, scale=np.abs(go) + 1e-8 * np.max(np.abs(go))) <= REL_TOL
        wa = mo.write_array(q)
        assert rel_err(mo.parametric_mean(wa), mg.parametric_mean(wa), scale=1.0) <= 1e-12
        assert rel_err(mo.parametric_mean(wa, True, False), mg.parametric_mean(wa, True, False), scale=1.0) <= 1e-12
        assert rel_err(mo.parametric_mean(wa, False, True), mg.parametric_mean(wa, False, True), scale=1.0) <= 1e-12


def test_sweep_level_expansion_equals_per_evaluation_pass():
    """Mode 1 (one device pass per sweep + exact quadratic expansion) against mode 0 (a pass per evaluation),
    near the anchor point (NUTS-sized moves) and far from it."""
    pr = friedman_problem(50000, seed=11)
    sd = pr["stan_data"]
    rng = np.random.default_rng(1)
    m0, m1 = GlmmModel(sd), GlmmModel(sd)
    m0.set_mode(0); m1.set_mode(1)
    off = rng.standard_normal(sd.N)
    m0.set_offset(off); m1.set_offset(off)
    q0 = rng.uniform(-0.5, 0.5, m0.d)
    passes0 = m1.num_device_passes()
    for scale in (0.0, 1e-4, 1e-2, 1.0):
        for _ in range(3):
            q = q0 + scale * rng.standard_normal(m0.d)
            la, ga, sa = m0.log_prob_grad(q)
            lb, gb, sb = m1.log_prob_grad(q)
            assert sa == sb == 0
            assert abs(la - lb) <= 1e-12 * abs(la)
            assert rel_err(ga, gb, scale=np.abs(ga) + 1e-9 * np.max(np.abs(ga))) <= REL_TOL
    assert m1.num_device_passes() == passes0 + 1          # one pass for the whole "sweep"
    m1.set_offset(off + 0.1)                              # new residual => new anchor
    m1.log_prob_grad(q0); m1.log_prob_grad(q0 + 0.01)
    assert m1.num_device_passes() == passes0 + 2


def test_data_terms_linearity_at_scale():
    """Size-independent property at a large N: S, X'e, Z'e are quadratic / linear in (beta, b)."""
    pr = friedman_problem(300000, seed=8)
    sd = pr["stan_data"]
    m = GlmmModel(sd)
    rng = np.random.default_rng(0)
    m.set_offset(rng.standard_normal(sd.N))
    z0, zb0 = np.zeros(sd.K), np.zeros(sd.q)
    S0, gx0, gz0 = m.data_terms(z0, zb0)
    beta, b = rng.standard_normal(sd.K), rng.standard_normal(sd.q)
    S1, gx1, gz1 = m.data_terms(beta, b)
    S2, gx2, gz2 = m.data_terms(2 * beta, 2 * b)
    # e(theta) = e0 - A theta  =>  g(theta) affine, S quadratic
    assert rel_err(gx2 - gx0, 2 * (gx1 - gx0), scale=np.abs(gx0) + np.abs(gx2)) <= 1e-11
    assert rel_err(gz2 - gz0, 2 * (gz1 - gz0), scale=np.abs(gz0) + np.abs(gz2)) <= 1e-11
    quad = S1 - S0 + 2 * (beta @ gx0 + b @ gz0)       # theta' A'A theta
    quad2 = S2 - S0 + 4 * (beta @ gx0 + b @ gz0)
    assert abs(quad2 - 4 * quad) <= 1e-10 * abs(quad2)
    # deterministic: repeated evaluation is bit identical
    assert m.data_terms(beta, b)[0] == S1


def test_non_finite_maps_to_status():
    sd, c = load_glmm_case(golden_case("friedman_like"))
    m = GlmmModel(sd)
    q = np.asarray(c["q"][0]).copy()
    q[-1] = 800.0
    assert m.log_prob_grad(q)[2] != 0


def test_unsupported_branches_fail_loudly():
    """A grouping term with more coefficients than the host's fixed-size onion scratch (16) is refused, not truncated."""
    from stan4bart_b200._lib import S4BError
    from stan4bart_b200.frontend import build_stan_data
    rng = np.random.default_rng(0)
    N = 60
    g = rng.integers(0, 3, N)
    M = np.column_stack([np.ones(N)] + [rng.random(N) for _ in range(16)])
    sd = build_stan_data(rng.random((N, 1)), rng.standard_normal(N), [(g, M)])
    with pytest.raises(S4BError):
        GlmmModel(sd)


@pytest.mark.parametrize("ncoef", [3, 5])
def test_blocks_with_more_than_two_coefficients_match_oracle(ncoef):
    """continuous.stan:44-70 (scaled onion rows through z_T) at sizes beyond the golden vectors."""
    from stan4bart_b200.frontend import build_stan_data
    rng = np.random.default_rng(ncoef)
    N = 4000
    g1, g2 = rng.integers(0, 7, N), rng.integers(0, 4, N)
    M1 = np.column_stack([np.ones(N)] + [rng.standard_normal(N) for _ in range(ncoef - 1)])
    M2 = np.column_stack([np.ones(N), rng.standard_normal(N)])
    sd = build_stan_data(rng.standard_normal((N, 2)), rng.standard_normal(N), [(g1, M1), (g2, M2)])
    off = rng.standard_normal(N)
    mo, mg = O.OracleGlmm(sd), GlmmModel(sd)
    mo.set_offset(off); mg.set_offset(off)
    assert mo.d == mg.d
    for mode in (0, 1):
        mg.set_mode(mode)
        for _ in range(3):
            q = rng.uniform(-1, 1, mo.d)
            lo, go, so = mo.log_prob_grad(q)
            lg, gg, sg = mg.log_prob_grad(q)
            assert so == sg == 0
            assert abs(lo - lg) <= REL_TOL * abs(lo)
            assert rel_err(go, gg, scale=np.abs(go) + 1e-8 * np.max(np.abs(go))) <= REL_TOL
            assert rel_err(mo.write_array(q), mg.write_array(q)) <= 1e-12


def test_weighted_data_pass_general_path_and_scale():
    """Observation weights outside the fast path (K > 4) and at a size with many blocks, both evaluation modes."""
    from stan4bart_b200.frontend import build_stan_data
    rng = np.random.default_rng(3)
    N = 50001
    g1 = rng.integers(0, 9, N)
    M1 = np.column_stack([np.ones(N), rng.standard_normal(N)])
    wt = rng.gamma(2.0, 0.5, N)
    wt[::97] = 0.0
    sd = build_stan_data(rng.standard_normal((N, 6)), rng.standard_normal(N), [(g1, M1)], weights=wt)
    off = rng.standard_normal(N)
    mo, mg = O.OracleGlmm(sd), GlmmModel(sd)
    mo.set_offset(off); mg.set_offset(off)
    for mode in (0, 1):
        mg.set_mode(mode)
        for _ in range(3):
            q = rng.uniform(-1, 1, mo.d)
            lo, go, so = mo.log_prob_grad(q)
            lg, gg, sg = mg.log_prob_grad(q)
            assert so == sg == 0
            assert abs(lo - lg) <= REL_TOL * abs(lo)
            assert rel_err(go, gg, scale=np.abs(go) + 1e-8 * np.max(np.abs(go))) <= REL_TOL


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("name", ["friedman_like", "four_and_three_binary", "weighted", "no_ranef_flat_prior"])
def test_column_path_matches_golden(monkeypatch, name, mode):
    """The data pass for models with more columns than the shared-memory bins hold (w e written once, dense columns as chunked
    dot products, one warp per column of Z), forced here on the golden models."""
    monkeypatch.setenv("S4B_GLMM_COLUMNS", "1")
    sd, c = load_glmm_case(golden_case(name))
    m = GlmmModel(sd)
    if sd.K + sd.q > 0:
        m.set_mode(mode)
    m.set_offset(np.asarray(c["offset"]))
    for q, lp, grad in zip(c["q"], c["lp"], c["grad"]):
        lp_g, g_g, status = m.log_prob_grad(np.asarray(q))
        assert status == 0
        assert abs(lp_g - lp) <= REL_TOL * abs(lp)
        assert rel_err(g_g, grad, scale=np.abs(grad) + 1e-8 * np.max(np.abs(grad))) <= REL_TOL


@pytest.mark.parametrize("bulk", ["0", "1"])
@pytest.mark.parametrize("path", golden_cases(), ids=lambda p: os.path.basename(p)[5:-5])
def test_bulk_copy_data_pass_matches_golden(monkeypatch, path, bulk):
    """k_glmm_data_terms_bulk (operand streams staged through shared memory by cp.async.bulk + mbarrier, producer warp + consumer
    warps) forced on / off on the golden models (tiny n: one ragged tile per CTA); models outside its scope (K or non-zeros per row
    above 4, no coefficients) take the register version either way."""
    monkeypatch.setenv("S4B_GLMM_BULK", bulk)
    sd, c = load_glmm_case(path)
    m = GlmmModel(sd)
    if sd.K + sd.q > 0:
        m.set_mode(0)
    m.set_offset(np.asarray(c["offset"]))
    for q, lp, grad in zip(c["q"], c["lp"], c["grad"]):
        lp_g, g_g, status = m.log_prob_grad(np.asarray(q))
        assert status == 0
        assert abs(lp_g - lp) <= REL_TOL * abs(lp)
        assert rel_err(g_g, grad, scale=np.abs(grad) + 1e-8 * np.max(np.abs(grad))) <= REL_TOL


@pytest.mark.parametrize("n,binary,weighted", [(1023, False, False), (1025, True, False), (70001, False, True), (300017, True, False)])
def test_bulk_copy_data_pass_matches_oracle_and_register_version(monkeypatch, n, binary, weighted):
    """Sizes around the tile boundaries (1024 rows per tile, three tiles in flight, several tiles per CTA at the largest size),
    weighted and unweighted: against the oracle at 1e-10 and against the register version of the same pass."""
    pr = friedman_problem(n, binary=binary, seed=21)
    sd = pr["stan_data"]
    rng = np.random.default_rng(n)
    if weighted:
        sd.weights = rng.gamma(2.0, 0.5, n)
    off = rng.standard_normal(sd.N)
    monkeypatch.setenv("S4B_GLMM_BULK", "1")
    mb = GlmmModel(sd)
    monkeypatch.setenv("S4B_GLMM_BULK", "0")
    mr = GlmmModel(sd)
    mo = O.OracleGlmm(sd)
    for m in (mb, mr, mo):
        m.set_offset(off)
    mb.set_mode(0); mr.set_mode(0)
    for _ in range(4):
        q = rng.uniform(-1, 1, mo.d)
        lo, go, so = mo.log_prob_grad(q)
        lb, gb, sb = mb.log_prob_grad(q)
        lr, gr, sr = mr.log_prob_grad(q)
        assert so == sb == sr == 0
        assert abs(lo - lb) <= REL_TOL * abs(lo) and abs(lr - lb) <= 1e-12 * abs(lr)
        assert rel_err(go, gb, scale=np.abs(go) + 1e-8 * np.max(np.abs(go))) <= REL_TOL
        assert rel_err(gr, gb, scale=np.abs(gr) + 1e-8 * np.max(np.abs(gr))) <= 1e-11


@pytest.mark.parametrize("levels,weighted", [(150, False), (1500, True)])
def test_many_grouping_levels(levels, weighted):
    """Hundreds to thousands of grouping levels (K + q far beyond the ~110 columns of the binned pass): random intercept and
    slope per level plus a crossed intercept; q = 2 * levels + 40."""
    from stan4bart_b200.frontend import build_stan_data
    rng = np.random.default_rng(levels)
    N = 60000
    g1 = rng.integers(0, levels, N); g1[:levels] = np.arange(levels)
    g2 = rng.integers(0, 40, N)
    M1 = np.column_stack([np.ones(N), rng.standard_normal(N)])
    wt = rng.gamma(2.0, 0.5, N) if weighted else None
    sd = build_stan_data(rng.standard_normal((N, 3)), rng.standard_normal(N), [(g1, M1), (g2, np.ones((N, 1)))], weights=wt)
    assert sd.q == 2 * levels + 40
    off = rng.standard_normal(N)
    mo, mg = O.OracleGlmm(sd), GlmmModel(sd)
    mo.set_offset(off); mg.set_offset(off)
    for mode in (0, 1):          # K + q > 512: the sweep-level expansion uses the Gram matrix in pieces (Z'WZ sparse)
        mg.set_mode(mode)
        for _ in range(3):
            q = rng.uniform(-1, 1, mo.d)
            lo, go, so = mo.log_prob_grad(q)
            lg, gg, sg = mg.log_prob_grad(q)
            assert so == sg == 0
            assert abs(lo - lg) <= REL_TOL * abs(lo)
            assert rel_err(go, gg, scale=np.abs(go) + 1e-8 * np.max(np.abs(go))) <= REL_TOL
            wa = mo.write_array(q)
            assert rel_err(mo.parametric_mean(wa), mg.parametric_mean(wa), scale=1.0) <= 1e-12
    # deterministic: repeated evaluation is bit identical
    a, b = mg.log_prob_grad(q), mg.log_prob_grad(q)
    assert a[0] == b[0] and np.array_equal(a[1], b[1])


@pytest.mark.parametrize("weighted", [False, True])
def test_crossed_factors_with_many_levels_use_the_gram_matrix_on_the_device(weighted, monkeypatch):
    """Two crossed grouping factors with hundreds of levels each fill Z'WZ in (300 x 300 cross block): the host-side sparse expansion
    would cost more than a device pass, so the Gram matrix lives on the device and G d is one kernel (k_gram_matvec).  Same
    density and gradient as the per-evaluation path, the host-side expansion and the oracle; one data pass per residual."""
    from stan4bart_b200.frontend import build_stan_data
    rng = np.random.default_rng(11)
    N, L1, L2 = 120000, 300, 300
    g1 = rng.integers(0, L1, N); g1[:L1] = np.arange(L1)
    g2 = rng.integers(0, L2, N); g2[:L2] = np.arange(L2)
    M1 = np.column_stack([np.ones(N), rng.standard_normal(N)])
    wt = rng.gamma(2.0, 0.5, N) if weighted else None
    sd = build_stan_data(rng.standard_normal((N, 2)), rng.standard_normal(N), [(g1, M1), (g2, np.ones((N, 1)))], weights=wt)
    assert sd.q == 2 * L1 + L2
    off = rng.standard_normal(N)
    mo, mg = O.OracleGlmm(sd), GlmmModel(sd)
    assert mg.mode() == 1                                   # the expansion stays the default thanks to the device-side product
    monkeypatch.setenv("S4B_GLMM_DEVICE_GRAM", "0")
    mh = GlmmModel(sd)                                       # host-side sparse expansion (default mode there: 0)
    mh.set_mode(1)
    for m in (mo, mg, mh):
        m.set_offset(off)
    passes0 = mg.num_device_passes()
    q0 = rng.uniform(-0.5, 0.5, mo.d)
    for scale in (0.0, 1e-3, 0.3):
        q = q0 + scale * rng.standard_normal(mo.d)
        lo, go, so = mo.log_prob_grad(q)
        lg, gg, sg = mg.log_prob_grad(q)
        lh, gh, sh = mh.log_prob_grad(q)
        assert so == sg == sh == 0
        assert abs(lo - lg) <= REL_TOL * abs(lo) and abs(lh - lg) <= 1e-12 * abs(lh)
        assert rel_err(go, gg, scale=np.abs(go) + 1e-8 * np.max(np.abs(go))) <= REL_TOL
        assert rel_err(gh, gg, scale=np.abs(gh) + 1e-8 * np.max(np.abs(gh))) <= 1e-11
    assert mg.num_device_passes() == passes0 + 1            # one data pass anchors all three evaluations
    mg.set_mode(0)
    l0, g0, _ = mg.log_prob_grad(q)
    assert abs(l0 - lg) <= 1e-12 * abs(l0) and rel_err(g0, gg, scale=np.abs(g0) + 1e-8 * np.max(np.abs(g0))) <= 1e-10
    a, b = mg.log_prob_grad(q), mg.log_prob_grad(q)
    assert a[0] == b[0] and np.array_equal(a[1], b[1])


@pytest.mark.parametrize("prior_dist", [1, 3, 4, 5, 6, 7])
def test_stan_row_names_match_the_reference_naming(prior_dist):
    """The names the library reports for the stored Stan rows (dimnames of the reference's `stan` result, src/stan_sampler.cpp:478-489,
    continuous.hpp:3115-3204) equal the front end's list for every coefficient prior family."""
    from stan4bart_b200.sampler import GlmmModel
    pr = friedman_problem(300)
    sd = pr["stan_data"]
    sd.prior_dist = prior_dist
    sd.prior_df = np.full(sd.K, 3.0)
    sd.num_normals = np.full(sd.K, 2, dtype=np.int32)
    sd.global_prior_df, sd.global_prior_scale, sd.slab_df, sd.slab_scale = 1.0, 0.1, 4.0, 2.5
    m = GlmmModel(sd)
    assert m.stan_row_names() == sd.param_names()
