"""Phase timing probe (development aid): python scripts/phase_probe.py MX ETA1 LEVELS [MAXIT]"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import exsaddle_b200 as X
import bench
mx, eta1, levels = int(sys.argv[1]), float(sys.argv[2]), int(sys.argv[3])
maxit = int(sys.argv[4]) if len(sys.argv) > 4 else 200
class A: pass
a = A(); a.mx, a.eta1, a.levels = mx, eta1, levels
opts = bench.workload_options(a) + " -saddle_ksp_max_it %d -xsb_time_kernels %s" % (maxit, " ".join(sys.argv[5:]))
g = X.ExSaddle(opts, nsd=3)
t = time.time(); g.assemble(); print("assemble %.3f s" % (time.time() - t), g.n, g.nnz, flush=True)
t = time.time(); g.ksp_setup(); print("ksp_setup %.3f s" % (time.time() - t), flush=True)
for l in range(levels):
    print(" level", l, g.mat_info(X.MAT_MG_LEVEL0 + l), g.chebyshev(l), flush=True)
for rep in range(2):
    t = time.time(); x = g.solve(); dt = time.time() - t
    its, reason = g.iterations(); c = g.counters(); inner = g.inner_iterations()
    print("solve %.3f s its %d reason %d inner %s" % (dt, its, reason, inner), c, g.timing(), flush=True)
    h = g.history(); print(" hist", " ".join("%.3e" % v for v in h), flush=True)
F = g.rhs(); print("true rel residual %.3e" % (np.linalg.norm(F - g.mat_mult(X.MAT_A, x)) / np.linalg.norm(F)))
