"""Inputs shared by the sharded-chain worker (one process per GPU) and the test that checks its output."""
import numpy as np

from common import bart_problem

BART_TREES, BART_SWEEPS, BART_SEED, BART_N = 12, 6, 4242, 3001
GLMM_N = 2003
GIBBS_N, GIBBS_TREES, GIBBS_SEED, GIBBS_WARMUP, GIBBS_ITER, GIBBS_SAMPLES = 1501, 9, 9876, 6, 11, 5


def allreduce_input(rank):
    rng = np.random.default_rng(100 + rank)
    return rng.standard_normal(37) * 10.0 ** rng.integers(-3, 4, 37)


def allreduce_long_input(rank):
    rng = np.random.default_rng(200 + rank)
    return rng.standard_normal(2500)


def bart_data(binary):
    x, y, _ = bart_problem(n=BART_N, p=6, binary=binary, seed=21)
    off = 0.3 * np.cos(np.arange(BART_N) * 0.01)
    return x, y, off


def quantile_bart_data():
    """Predictors for use.quantiles on sharded rows: continuous, rounded (values shared between the shards), a few integers, one that is
    constant inside the first shard only (it has cut points because the OTHER shard varies), one constant everywhere (no cut at all)."""
    x, y, off = bart_data(False)
    x = x.copy()
    n = len(y)
    x[:, 1] = np.round(x[:, 1], 1)
    x[:, 2] = np.floor(4.0 * x[:, 2])
    x[: (n + 1) // 2 + 7, 3] = 0.25
    x[:, 4] = 1.5
    return x, y, off


def glmm_offset():
    return 0.5 * np.sin(np.arange(GLMM_N) * 0.37)


def glmm_points(d):
    rng = np.random.default_rng(5)
    return [rng.uniform(-1.0, 1.0, d) for _ in range(3)]


def gibbs_weights():
    rng = np.random.default_rng(77)
    return rng.gamma(3.0, 1.0 / 3.0, GIBBS_N)
