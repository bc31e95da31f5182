"""ncu target: a few BART sweeps at the benchmark shape (binary Friedman, n = 1 M, 200 trees) for kernel captures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stan4bart_b200.frontend import friedman_problem
from stan4bart_b200.sampler import GpuBart
from stan4bart_b200.structs import bart_config
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
sweeps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
pr = friedman_problem(n, binary=True)
g = GpuBart(bart_config(n, 9, num_trees=200, seed=1, is_binary=True), pr["y"], pr["x_bart"])
for _ in range(sweeps):
    g.run()
print(g.pipeline())
