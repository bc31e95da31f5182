"""Multi-GPU plumbing: chains are the unit of parallelism (the reference runs one R worker process per chain,
R/stan4bart_fit.R:495-542); one process per GPU, chains dealt round-robin, no collective on the data path.
torch.distributed is used only for the start barrier and for taking the max of the per-rank device times."""
import torch
import torch.distributed as dist


def chains_for_rank(num_chains, rank, world_size):
    """Chain c runs on rank c mod world_size."""
    if num_chains < 1 or world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad chain partition request")
    return [c for c in range(num_chains) if c % world_size == rank]


def chain_seed(base_seed, chain):
    """Distinct, reproducible seed per chain (the reference draws one seed per chain, R/stan4bart_fit.R:510-516)."""
    return int(base_seed) + int(chain)


def _device_for_backend():
    return "cuda" if dist.is_initialized() and dist.get_backend() == "nccl" else "cpu"


def barrier():
    if dist.is_initialized() and dist.get_world_size() > 1:
        t = torch.zeros(1, device=_device_for_backend())
        dist.all_reduce(t)
    if torch.cuda.is_available():
        torch.cuda.synchronize()


def max_over_ranks(x):
    """Max of a per-rank scalar (e.g. the device time of the timed region)."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        t = torch.tensor([float(x)], dtype=torch.float64, device=_device_for_backend())
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    return float(x)


def min_over_ranks(x):
    if dist.is_initialized() and dist.get_world_size() > 1:
        t = torch.tensor([float(x)], dtype=torch.float64, device=_device_for_backend())
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return float(t.item())
    return float(x)


def sum_over_ranks(x):
    if dist.is_initialized() and dist.get_world_size() > 1:
        t = torch.tensor([float(x)], dtype=torch.float64, device=_device_for_backend())
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())
    return float(x)


def aggregate_throughput(units_this_rank, ms_this_rank):
    """Whole-job throughput = units processed by all ranks / max over ranks of the elapsed time."""
    total_units = sum_over_ranks(units_this_rank)
    ms = max_over_ranks(ms_this_rank)
    return total_units / (ms / 1000.0), ms
