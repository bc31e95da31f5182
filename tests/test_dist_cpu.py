"""CPU (gloo, world_size 2): the multi-rank plumbing of bench.py -- chain partition, barrier, max-over-ranks timing."""
import os
import socket

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
