"""Phase cycle counters of the sweep kernel for an observation-sharded chain (run under torch.distributed.run, one rank
per GPU).  usage: torchrun ... tools/shard_profile.py [total_rows] [trees]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch.distributed as dist  # noqa: E402

from stan4bart_b200 import _lib  # noqa: E402
from stan4bart_b200.frontend import friedman_problem  # noqa: E402
from stan4bart_b200.sampler import GpuBart  # noqa: E402
from stan4bart_b200.shard import ShardContext, row_range  # noqa: E402
from stan4bart_b200.structs import bart_config  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
T = int(sys.argv[2]) if len(sys.argv) > 2 else 200
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
_lib.check(_lib.load().s4b_set_device(int(os.environ.get("LOCAL_RANK", 0))))
dist.init_process_group("gloo", rank=rank, world_size=world)
ctx = ShardContext.from_torch_distributed(total_obs=n)
pr = friedman_problem(n)
lo, hi = row_range(n, rank, world)
cfg = bart_config(hi - lo, 9, num_trees=T, seed=1)
g = GpuBart(cfg, pr["y"][lo:hi], pr["x_bart"][lo:hi], shard=ctx if world > 1 else None)
g.set_sigma(1.0)
for _ in range(30):
    g.run()
g.set_profile(True); g.profile(); g.tree_step_ms()
dist.barrier()
for _ in range(20):
    g.run()
ms = g.tree_step_ms() / 20
prof = g.profile()
if rank == 0:
    print(json.dumps({"world": world, "rows_total": n, "rows_rank0": hi - lo, "device_ms_per_sweep": ms, "us_per_tree_step": ms / T * 1e3,
                      "cycles_per_step": prof["cycles_per_step"], "worker": prof["worker_cycles_per_step"]}, indent=1))
dist.barrier()
dist.destroy_process_group()
