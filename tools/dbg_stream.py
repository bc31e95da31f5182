import os, sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
from common import bart_problem
from stan4bart_b200.sampler import GpuBart
from stan4bart_b200.structs import bart_config
T=4
x, y, xt = bart_problem(2777, 6, 0, False, seed=21)
cfg = bart_config(2777, 6, num_trees=T, is_binary=False, seed=33)
g1 = GpuBart(cfg, y, x, xt)
os.environ["S4B_FORCE_STREAM"]="1"
g2 = GpuBart(cfg, y, x, xt)
for b in (g1,g2):
    b.set_sigma(1.3); b.sample_trees_from_prior(); b.set_trace(T*2)
print("resid equal before", np.array_equal(g1.residual(), g2.residual()))
r1=g1.run(); r2=g2.run()
t1,t2=g1.trace(),g2.trace()
np.set_printoptions(linewidth=200, precision=6)
for i in range(T):
    print(i, "reg", t1[i][:12]); print(i, "str", t2[i][:12])
print("resid diff", np.abs(g1.residual()-g2.residual()).max())
print(g1.leaf_stats(0)); print(g2.leaf_stats(0))
