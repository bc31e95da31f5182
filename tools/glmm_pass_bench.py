"""The GLMM data pass alone (S = sum e^2, X'e, Z'e over all rows) at config C's shape: register version against the bulk-copy
(TMA) version, L2-warm (back-to-back launches) and from HBM (L2 flushed before every launch); achieved GB/s in the algorithmic
byte count of DESIGN.md section 4 (44 B per row for the Friedman model) against the measured HBM peak.
usage: python tools/glmm_pass_bench.py [n ...]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stan4bart_b200.frontend import friedman_problem
from stan4bart_b200.sampler import GlmmModel

sizes = [int(a) for a in sys.argv[1:]] or [1_000_000, 4_000_000]
peak = 6546.6
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
out = []
for n in sizes:
    pr = friedman_problem(n, binary=True, seed=99)
    sd = pr["stan_data"]
    bytes_per_row = 8 + 8 * sd.K + 4 * 3 + 8 * 1          # r, X, three index streams, one value stream (two slots are indicators)
    for bulk in ("0", "1"):
        os.environ["S4B_GLMM_BULK"] = bulk
        m = GlmmModel(sd)
        for flush in (False, True):
            ms, is_bulk = m.time_data_pass(30, flush)
            out.append({"n": n, "kernel": "k_glmm_data_terms_bulk" if is_bulk else "k_glmm_data_terms", "l2": "flushed before every launch" if flush else "warm (back-to-back launches)",
                        "launch_us": ms * 1e3, "gbs": bytes_per_row * n / ms / 1e6, "frac_of_peak": bytes_per_row * n / ms / 1e6 / peak})
        del m
print(json.dumps({"peak_gbs": peak, "bytes_per_row": 44, "runs": out}, indent=1))
