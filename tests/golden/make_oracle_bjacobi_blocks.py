"""Generates tests/golden/oracle_64cubed_bjacobi_blocks.json: the oracle at the BASELINE 64^3 workload with the pressure
bjacobi/ILU(0) split into 2, 4 and 8 z-slab blocks (= the product on 2, 4 and 8 GPUs; bench.py asserts its multi-GPU iteration counts and
residual history against it).  ~45 GB RAM, ~10 min per run on 8 cores.
    python tests/golden/make_oracle_bjacobi_blocks.py [blocks ...]   (default 2 4 8; existing entries of other block counts are kept)"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O
abf = " ".join(l for l in O.ABF_OPTS.split("\n") if l.strip())
path = os.path.join(ROOT, "tests", "golden", "oracle_64cubed_bjacobi_blocks.json")
out = json.load(open(path)) if os.path.exists(path) else {}
for nb in ([int(v) for v in sys.argv[1:]] or [2, 4, 8]):
    t0 = time.time()
    p = O.Problem("%s -saddle_fieldsplit_u_pc_mg_levels 6 -model 6 -mx 64 -eta0 1 -eta1 1e6 -saddle_ksp_rtol 1e-8 -xo_p_blocks %d" % (abf, nb), nsd=3)
    x, r = p.solve()
    out[str(nb)] = {"its": r.its, "reason": r.reason, "inner": [int(v) for v in r.inner_its[:r.n_inner]], "hist": [float(v) for v in r.hist[:r.nhist]],
                    "hist_tail": [float(v) for v in r.hist[r.nhist - 3:r.nhist]], "seconds": time.time() - t0}
    print(nb, out[str(nb)]["its"], out[str(nb)]["seconds"], flush=True)
    del p
    json.dump(out, open(path, "w"), indent=1)
