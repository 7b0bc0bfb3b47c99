"""Run one exSaddle command line on the GPU and print a JSON summary (development aid / profiles):
    python scripts/run_case.py [--lame] [--nsd 3] [--solves 2] -- <exSaddle options>
Second-solve protocol (exSaddle.c:569-599): the last solve is the one timed."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import exsaddle_b200 as X
av = sys.argv[1:]
lame = "--lame" in av
nsd = int(av[av.index("--nsd") + 1]) if "--nsd" in av else 3
solves = int(av[av.index("--solves") + 1]) if "--solves" in av else 2
opts = " ".join(av[av.index("--") + 1:])
t0 = time.time(); g = X.ExSaddle(opts, nsd=nsd, lame=lame).assemble(); t_asm = time.time() - t0
t0 = time.time(); g.ksp_setup(); t_setup = time.time() - t0
for _ in range(solves):
    x = g.solve()
its, reason = g.iterations()
h = g.history()
b = g.rhs()
r = b - g.mat_mult(X.MAT_A, x)
tm = g.timing()
print(json.dumps({"options": opts, "lame": lame, "n": g.n, "nnz": g.nnz, "its": its, "reason": reason, "inner": g.inner_iterations(),
                  "rnorm0": float(h[0]), "rnorm": float(h[-1]), "true_rel_res": float(np.linalg.norm(r) / np.linalg.norm(b)),
                  "assemble_s": t_asm, "ksp_setup_s": t_setup, "solve_s": tm[1] / 1e3 if tm else None, "counters": g.counters()}))
g.close()
