"""Generates tests/golden/oracle_64cubed_bjacobi_blocks.json: the oracle at the BASELINE 64^3 workload with the pressure
bjacobi/ILU(0) split into 2 and 8 z-slab blocks (= the product on 2 and 8 GPUs).  ~45 GB RAM, ~8 min per run on 8 cores."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O
abf = " ".join(l for l in O.ABF_OPTS.split("\n") if l.strip())
out = {}
for nb in (2, 8):
    t0 = time.time()
    p = O.Problem("%s -saddle_fieldsplit_u_pc_mg_levels 6 -model 6 -mx 64 -eta0 1 -eta1 1e6 -saddle_ksp_rtol 1e-8 -xo_p_blocks %d" % (abf, nb), nsd=3)
    x, r = p.solve()
    out[nb] = {"its": r.its, "reason": r.reason, "inner": [int(v) for v in r.inner_its[:r.n_inner]], "hist_tail": [float(v) for v in r.hist[r.nhist - 3:r.nhist]], "seconds": time.time() - t0}
    print(nb, out[nb], flush=True)
    del p
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "oracle_64cubed_bjacobi_blocks.json"), "w"), indent=1)
