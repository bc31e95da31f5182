"""Observation-sharded chains (BASELINE config E): Python side of `s4b_shard_*`.

One process per GPU.  `ShardContext.from_torch_distributed()` creates the mailbox, exchanges the CUDA IPC handles
through the process group (any backend: the handles are 64 plain bytes) and attaches the peers; `row_range` deals the
rows out in contiguous blocks.  The data path itself never touches torch.distributed: the per-tree statistics and the
GLMM reductions travel through the peer-mapped mailboxes inside the kernels (csrc/shard.hpp)."""
import ctypes as C

import numpy as np

from . import _lib
from .structs import dptr, f64

HANDLE_BYTES = 64


def row_range(total_obs, rank, world):
    """Contiguous block of rows for `rank`: sizes differ by at most one, multiples of 4 where possible so that the
    quads of the sweep kernel never straddle ranks."""
    if total_obs < 0 or world < 1 or not (0 <= rank < world):
        raise ValueError("bad row partition request")
    quads = (total_obs + 3) // 4
    lo_q = (quads * rank) // world
    hi_q = (quads * (rank + 1)) // world
    return min(4 * lo_q, total_obs), min(4 * hi_q, total_obs)


class ShardContext:
    def __init__(self, rank, world):
        self.L = _lib.load()
        _lib.require_device()
        self.rank, self.world = int(rank), int(world)
        h = C.c_void_p()
        _lib.check(self.L.s4b_shard_create(self.rank, self.world, C.byref(h)))
        self.h = h
        self.first_obs, self.total_obs = 0, 0

    def __del__(self):
        if getattr(self, "h", None):
            self.L.s4b_shard_free(self.h)
            self.h = None

    def ipc_handle(self):
        buf = (C.c_ubyte * HANDLE_BYTES)()
        _lib.check(self.L.s4b_shard_ipc_handle(self.h, buf))
        return bytes(buf)

    def attach(self, handles):
        """handles: list of `world` 64-byte strings ordered by rank."""
        if len(handles) != self.world or any(len(x) != HANDLE_BYTES for x in handles):
            raise ValueError("need one 64-byte handle per rank")
        flat = (C.c_ubyte * (HANDLE_BYTES * self.world)).from_buffer_copy(b"".join(handles))
        _lib.check(self.L.s4b_shard_attach(self.h, flat))

    def set_obs_range(self, first_obs, total_obs):
        _lib.check(self.L.s4b_shard_set_obs_range(self.h, int(first_obs), int(total_obs)))
        self.first_obs, self.total_obs = int(first_obs), int(total_obs)

    def allreduce(self, vec, op="sum"):
        v = f64(np.array(vec, dtype=np.float64, copy=True).ravel())
        _lib.check(self.L.s4b_shard_allreduce(self.h, dptr(v), v.size, {"sum": 0, "max": 1}[op]))
        return v

    def init_nccl(self):
        """Reference collective: create the NCCL communicator of this context's ranks (collective; the unique id travels through
        the torch.distributed process group) -- `use_nccl(True)` then routes the small all-reduces through ncclAllReduce."""
        buf = (C.c_ubyte * 128)()
        if self.rank == 0:
            _lib.check(self.L.s4b_shard_nccl_unique_id(buf))
        if self.world == 1:                # a communicator of one rank: nothing to broadcast
            _lib.check(self.L.s4b_shard_nccl_init(self.h, buf))
            return
        import torch
        import torch.distributed as dist
        dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
        t = torch.tensor(list(buf), dtype=torch.uint8, device=dev)
        dist.broadcast(t, src=0)
        ident = (C.c_ubyte * 128).from_buffer_copy(bytes(t.cpu().tolist()))
        _lib.check(self.L.s4b_shard_nccl_init(self.h, ident))

    def use_nccl(self, on):
        _lib.check(self.L.s4b_shard_use_nccl(self.h, int(bool(on))))

    @classmethod
    def from_torch_distributed(cls, total_obs=None):
        """Build and attach a context for the current torch.distributed process group (one rank per GPU)."""
        import torch
        import torch.distributed as dist
        rank, world = dist.get_rank(), dist.get_world_size()
        ctx = cls(rank, world)
        if world > 1:
            mine = torch.tensor(list(ctx.ipc_handle()), dtype=torch.uint8)
            dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
            got = [torch.zeros(HANDLE_BYTES, dtype=torch.uint8, device=dev) for _ in range(world)]
            dist.all_gather(got, mine.to(dev))
            ctx.attach([bytes(g.cpu().tolist()) for g in got])
        if total_obs is not None:
            lo, _ = row_range(total_obs, rank, world)
            ctx.set_obs_range(lo, total_obs)
        return ctx
