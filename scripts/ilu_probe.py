"""One pressure-block ILU(0) solve on the bench workload's lattice (for ncu captures of k_ilup_solve / timing by shape).
usage: ilu_probe.py [mx my mz]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import exsaddle_b200 as X
from oracle import oracle as O
mx, my, mz = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (64, 64, 64)
abf = " ".join(l for l in O.ABF_OPTS.split("\n") if l.strip())
lv = 1
while (mx % (1 << lv) == 0 and my % (1 << lv) == 0 and mz % (1 << lv) == 0) and 3 * (mx // (1 << lv) * 2 + 1) * (my // (1 << lv) * 2 + 1) * (mz // (1 << lv) * 2 + 1) > 6000:
    lv += 1
g = X.ExSaddle(abf + " -saddle_fieldsplit_u_pc_mg_levels %d -xsb_matrix_free full -model 6 -mx %d -my %d -mz %d -eta1 1e6 %s" % (lv + 1, mx, my, mz, " ".join(sys.argv[4:])), nsd=3).assemble().ksp_setup()
b = np.cos(0.3 * np.arange(g.np_))
x = g.pc_schur_apply(b)
print("ilu_probe", mx, my, mz, "levels", lv + 1, "device ms per apply: %.4f" % g.time_pc_schur(20), "|x|", float(np.linalg.norm(x)))
g.close()
