"""CPU (gloo, world_size 2): the multi-rank plumbing of bench.py -- chain partition, barrier, max-over-ranks timing."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from stan4bart_b200.dist import aggregate_throughput, barrier, chain_seed, chains_for_rank, max_over_ranks


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        barrier()
        chains = chains_for_rank(5, rank, world)
        # rank r pretends to have needed (10 + 5 r) ms for 20 sweeps of each of its chains
        value, ms = aggregate_throughput(20 * len(chains), 10.0 + 5.0 * rank)
        out[rank] = (chains, value, ms, max_over_ranks(rank))
    finally:
        dist.destroy_process_group()


def test_two_ranks_gloo():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert out[0][0] == [0, 2, 4] and out[1][0] == [1, 3]
    for r in range(world):
        chains, value, ms, mx = out[r]
        assert ms == 15.0                       # max over ranks
        assert value == pytest.approx(100 / 0.015)   # all 5 chains x 20 sweeps over the slowest rank's time
        assert mx == 1.0


def test_partition_properties():
    for world in (1, 2, 4, 8):
        for chains in (1, 8, 64):
            parts = [chains_for_rank(chains, r, world) for r in range(world)]
            flat = sorted(c for p in parts for c in p)
            assert flat == list(range(chains))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    assert len({chain_seed(12345, c) for c in range(64)}) == 64
    with pytest.raises(ValueError):
        chains_for_rank(4, 3, 2)


def test_single_process_fallbacks():
    assert max_over_ranks(3.5) == 3.5
    v, ms = aggregate_throughput(40, 20.0)
    assert v == 2000.0 and ms == 20.0


# ---- observation-sharded chains: host-side row partition (stan4bart_b200/shard.py, frontend.shard_problem) ----
def test_row_range_covers_everything_once():
    from stan4bart_b200.shard import row_range
    for n in (1, 2, 7, 100, 1501, 1000000):
        for world in (1, 2, 3, 8):
            edges = [row_range(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            for a, b in zip(edges[:-1], edges[1:]):
                assert a[1] == b[0]
            assert all(lo % 4 == 0 for lo, _ in edges if lo < n)
    with pytest.raises(ValueError):
        row_range(10, 2, 2)


@pytest.mark.parametrize("weighted", [False, True])
def test_sharded_stan_data_terms_add_up_to_the_whole(weighted):
    """The GLMM reductions a sharded chain all-reduces (S, X'e, Z'e) are sums over rows: the row shards produced by
    StanData.rows() must add up to the whole-data terms on the CPU oracle (with observation weights: each shard carries
    its own rows' weights)."""
    import oracle_lib as O
    from stan4bart_b200.frontend import friedman_problem, shard_problem
    from stan4bart_b200.shard import row_range
    n, world = 403, 3
    pr = friedman_problem(n)
    sd = pr["stan_data"]
    rng = np.random.default_rng(1)
    if weighted:
        pr["weights"] = rng.gamma(2.0, 0.5, n)
        sd.weights = pr["weights"]
    beta, b = rng.standard_normal(sd.K), rng.standard_normal(sd.q)
    off = rng.standard_normal(n)
    whole = O.OracleGlmm(sd)
    whole.set_offset(off)
    S, gbeta, gb = whole.data_terms(beta, b)
    S2, gbeta2, gb2 = 0.0, np.zeros(sd.K), np.zeros(sd.q)
    rows = 0
    for r in range(world):
        lo, hi = row_range(n, r, world)
        sp = shard_problem(pr, lo, hi)
        assert sp["stan_data"].q == sd.q and sp["stan_data"].N == hi - lo and len(sp["y"]) == hi - lo
        if weighted:
            assert np.array_equal(sp["stan_data"].weights, pr["weights"][lo:hi]) and np.array_equal(sp["weights"], pr["weights"][lo:hi])
        part = O.OracleGlmm(sp["stan_data"])
        part.set_offset(off[lo:hi])
        s_, a_, b_ = part.data_terms(beta, b)
        S2 += s_; gbeta2 += a_; gb2 += b_
        rows += hi - lo
    assert rows == n
    assert abs(S - S2) <= 1e-10 * abs(S)
    assert np.allclose(gbeta, gbeta2, rtol=1e-10, atol=1e-10) and np.allclose(gb, gb2, rtol=1e-10, atol=1e-10)
