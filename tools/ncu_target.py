"""Small driver for ncu captures: a few BART sweeps at the benchmark shape (n = 1M, 200 trees)."""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stan4bart_b200.frontend import friedman_problem
from stan4bart_b200.sampler import GpuBart
from stan4bart_b200.structs import bart_config

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
sweeps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
binary = len(sys.argv) > 3 and sys.argv[3] == "binary"
pr = friedman_problem(n, binary=binary)
g = GpuBart(bart_config(n, 9, num_trees=200, seed=1, is_binary=binary), pr["y"], pr["x_bart"])
if not binary:
    g.set_sigma(1.0)
for _ in range(sweeps):
    g.run()
print("done", g.num_tree_steps())
