import sys, json, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import exsaddle_b200 as X
opts = "-model 2 -sinker_n 1 -fs -mx 4 -diagnostics -saddle_ksp_monitor_short"
z = {}
for tag, extra in (("r1", ""), ("r2", " -xsb_ranks 2"), ("r2_first", None)):
    g = X.ExSaddle(("-xsb_ranks 2 " + opts) if extra is None else opts + extra, nsd=3).assemble().ksp_setup()
    r = np.cos(0.3 * np.arange(g.n)) + 0.1
    z[tag] = g.pc_apply(r)
    g.solve()
    print(tag, "its", g.iterations(), ["%g" % v for v in g.history()])
    g.close()
print("pc diff r1-r2", np.linalg.norm(z["r1"] - z["r2"]) / np.linalg.norm(z["r1"]), "r2-r2_first", np.linalg.norm(z["r2"] - z["r2_first"]))
