"""History-difference probe (development aid): GPU vs oracle residual histories on the MG test cases."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import exsaddle_b200 as X
from oracle import oracle as O
from tests.test_gpu_parity import MG_CASES, ABF
for cid, nsd, lame, opts, levels in MG_CASES:
    full = "%s %s -saddle_fieldsplit_u_pc_mg_levels %d -saddle_ksp_rtol 1e-8" % (ABF, opts, levels)
    g = X.ExSaddle(full, nsd=nsd, lame=lame).assemble().ksp_setup()
    o = O.Problem(full, nsd=nsd, lame=lame)
    x = g.solve(); xo, r = o.solve()
    h = g.history(); ho = np.array(r.hist[:r.nhist])
    n = min(len(h), len(ho))
    d = np.abs(h[:n] - ho[:n])
    big = ho[:n] >= 1e-4 * ho[0]
    print(cid, g.iterations(), (r.its, r.reason), g.inner_iterations() == list(r.inner_its[:r.n_inner]),
          "max|d|/h0 %.1e" % (d.max() / ho[0]), "maxrel(big) %.1e" % (d[big] / ho[:n][big]).max(), "maxrel %.1e" % (d / ho[:n]).max(),
          "xerr %.1e" % (np.linalg.norm(x - xo) / np.linalg.norm(xo)))
    g.close()
