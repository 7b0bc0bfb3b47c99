"""Generates tests/golden/oracle_64cubed_history.json: the CPU oracle's residual history of the BASELINE 64^3 workload
(bench.py default: model 6, eta1/eta0 = 1e6, abf.opts tree with 6 MG levels, rtol 1e-8).  ~45 GB of RAM, ~25 min on 8 cores.
    python tests/golden/make_oracle_64cubed.py [mx] [levels]
The GPU parity test compares its own history at the same size against this fixture (tests/test_gpu_parity.py)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
from oracle import oracle as O
mx = int(sys.argv[1]) if len(sys.argv) > 1 else 64
levels = int(sys.argv[2]) if len(sys.argv) > 2 else 6
abf = " ".join(l for l in O.ABF_OPTS.split("\n") if l.strip())
opts = "%s -saddle_fieldsplit_u_pc_mg_levels %d -model 6 -mx %d -eta0 1 -eta1 1e6 -saddle_ksp_rtol 1e-8" % (abf, levels, mx)
t0 = time.time()
p = O.Problem(opts, nsd=3)
x, r = p.solve()
F = p.F()
out = {"options": opts, "mx": mx, "levels": levels, "its": r.its, "reason": r.reason, "hist": [float(v) for v in r.hist[:r.nhist]],
       "inner_its": [int(v) for v in r.inner_its[:r.n_inner]], "cheb_emax_est": [float(v) for v in r.cheb_emax_est[:levels]],
       "x_norm2": float(np.linalg.norm(x)), "x_u_absmax": float(np.max(np.abs(x[:p.nu]))), "p_absmax": float(np.max(np.abs(x[p.nu:]))),
       "true_rel_res": float(np.linalg.norm(F - p.mult(x)) / np.linalg.norm(F)), "seconds": time.time() - t0, "threads": O.lib().xo_num_threads()}
# levels == 3 is abf.opts verbatim (coarsest level 17^3 / 33^3 nodes, solved exactly: banded Cholesky in the oracle, UMFPACK LU in the reference)
name = "oracle_%dcubed_abf3_history.json" % mx if levels == 3 else "oracle_%dcubed_history.json" % mx
json.dump(out, open(os.path.join(ROOT, "tests", "golden", name), "w"), indent=1)
print(json.dumps({k: out[k] for k in ("its", "reason", "seconds", "true_rel_res")}))
