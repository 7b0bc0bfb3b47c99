"""Oracle-vs-oracle drift: how far does the residual history of the BASELINE workload move when NOTHING but the floating-point
summation order changes?  ("run the oracle against itself with a different summation order to prove the floor".)
Runs the CPU oracle with reversed dot-product / matrix-row sums (xo_set_sum_order(1)) and compares its history with the committed
fixture (reference order, tests/golden/oracle_<mx>cubed_history.json).  CPU only; 64^3 needs ~45 GB and ~10-25 min on 8 cores.
    python scripts/oracle_drift.py [mx] [levels] [threads]
Writes profiles/r02_oracle_drift_<mx>cubed.json."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from oracle import oracle as O
mx = int(sys.argv[1]) if len(sys.argv) > 1 else 64
levels = int(sys.argv[2]) if len(sys.argv) > 2 else 6
threads = int(sys.argv[3]) if len(sys.argv) > 3 else 0
fx = json.load(open(os.path.join(ROOT, "tests", "golden", "oracle_%dcubed_history.json" % mx)))
assert fx["levels"] == levels
L = O.lib()
if threads:
    L.xo_set_num_threads(threads)
L.xo_set_sum_order(1)
t0 = time.time()
p = O.Problem(fx["options"], nsd=3)
x, r = p.solve()
h1 = np.array(r.hist[:r.nhist]); h0 = np.array(fx["hist"])
m = min(len(h0), len(h1))
rel = np.abs(h1[:m] - h0[:m]) / h0[:m]
def upto(thr):
    k = [i for i in range(m) if h0[i] >= thr * h0[0]]
    return float(rel[k].max()) if k else None
out = {"what": "oracle (reversed dot / row sums, %d threads) vs oracle fixture (reference order, %d threads)" % (L.xo_num_threads(), fx.get("threads", 0)),
       "options": fx["options"], "its": [fx["its"], int(r.its)], "inner_its_equal": [int(v) for v in r.inner_its[:r.n_inner]] == fx["inner_its"],
       "inner_its": [int(v) for v in r.inner_its[:r.n_inner]],
       "max_rel_hist_diff_above_1e-2": upto(1e-2), "max_rel_hist_diff_above_1e-4": upto(1e-4), "max_rel_hist_diff_all": float(rel.max()),
       "rel_hist_diff": [float(v) for v in rel], "seconds": time.time() - t0}
json.dump(out, open(os.path.join(ROOT, "profiles", "r02_oracle_drift_%dcubed.json" % mx), "w"), indent=1)
print(json.dumps({k: out[k] for k in ("its", "inner_its_equal", "max_rel_hist_diff_above_1e-2", "max_rel_hist_diff_above_1e-4", "max_rel_hist_diff_all", "seconds")}))
