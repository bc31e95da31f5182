"""Debug helper: run the sharded worker (2 ranks, any number of GPUs) and print where the continuous Gibbs case departs from the oracle."""
import os, subprocess, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import oracle_lib as O
import shard_cases as SC
from stan4bart_b200.frontend import friedman_problem
from stan4bart_b200.structs import bart_config, stan_control

tmp = tempfile.mkdtemp()
out = os.path.join(tmp, "shard")
cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port", "29617",
       os.path.join(ROOT, "tests", "shard_worker.py"), "--out", out]
p = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
print("rc", p.returncode, p.stderr[-2000:])
ranks = [dict(np.load(f"{out}.rank{r}.npz")) for r in range(2)]
import shutil
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
shutil.copy(f"{out}.rank0.npz", os.path.join(ROOT, "gpurun_out", "shard_debug_rank0.npz"))
for binary in (False, True):
    tag = "bin" if binary else "cont"
    n = SC.GIBBS_N
    pr = friedman_problem(n, binary=binary)
    cfg = bart_config(n, 9, n_test=n, num_trees=SC.GIBBS_TREES, is_binary=binary, seed=SC.GIBBS_SEED)
    ctl = stan_control(seed=SC.GIBBS_SEED + 1)
    o = O.OracleSampler(cfg, pr["y"], pr["x_bart"], pr["x_test"], pr["stan_data"], ctl, warmup=SC.GIBBS_WARMUP, iter_=SC.GIBBS_ITER,
                        keep_fits=True, sigma_init=pr["sigma_init"], bart_offset_init=pr["bart_offset_init"])
    ob = o.bart()
    ob.set_trace(SC.GIBBS_TREES * (SC.GIBBS_WARMUP + SC.GIBBS_SAMPLES))
    w = o.run(SC.GIBBS_WARMUP, True)
    tr_o = ob.trace()
    for ri, r in enumerate(ranks):
        tr_g = r[f"gibbs_{tag}_trace"]
        m = min(len(tr_o), len(tr_g))
        d = np.abs(tr_o[:m] - tr_g[:m]) / (np.abs(tr_o[:m]) + 1.0)
        bad = np.nonzero(d.max(axis=1) > 1e-7)[0]
        print(tag, "rank", ri, "first bad trace step", bad[:5], "of", m)
        if bad.size:
            k = bad[0]
            print(" oracle", tr_o[k][:14]); print(" gpu   ", tr_g[k][:14])
        ws = r[f"gibbs_{tag}_w_stan"]
        ds = np.abs(w["stan"] - ws) / (np.abs(w["stan"]) + 1.0)
        print(" stan rel err per warmup sweep", ds.max(axis=0))
        print(" ranks identical trace:", np.array_equal(ranks[0][f"gibbs_{tag}_trace"], ranks[1][f"gibbs_{tag}_trace"]))
