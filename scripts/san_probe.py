import sys
sys.path.insert(0,'/root/repo')
import exsaddle_b200 as X
from oracle import oracle as O
abf = " ".join(l for l in O.ABF_OPTS.split("\n") if l.strip())
g = X.ExSaddle(abf + " -model 6 -mx 4 -eta1 100 -saddle_fieldsplit_u_pc_mg_levels 2", nsd=3).assemble()
g.ksp_setup()
print("setup ok")
