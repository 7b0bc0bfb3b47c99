"""SpMV micro-benchmark (development aid): python scripts/spmv_bench.py MX [REPS]
Times y = A00 x (BAIJ) and y = A x (AIJ) on device-resident vectors with CUDA events on the library stream."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import exsaddle_b200 as X
mx = int(sys.argv[1]); reps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
more = " " + sys.argv[3] if len(sys.argv) > 3 else ""      # e.g. "-xsb_baij_closed_form"
g = X.ExSaddle("-mx %d -model 6 -eta1 1e6%s" % (mx, more), nsd=3).assemble()
st = torch.cuda.ExternalStream(g.stream())
peak = 6548.2
out = {}
for name, which in (("A00_baij", X.MAT_A00), ("A_aij", X.MAT_A)):
    rows, cols, nnz, bs = g.mat_info(which)
    x = torch.sin(0.37 * torch.arange(cols, dtype=torch.float64, device="cuda")) + 0.1
    y = torch.empty(rows, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    with torch.cuda.stream(st):
        for _ in range(5):
            g.mat_mult_dev(which, x.data_ptr(), y.data_ptr())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(reps):
            g.mat_mult_dev(which, x.data_ptr(), y.data_ptr())
        e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    if bs > 1:
        nblk = nnz // (bs * bs); byts = (8 * bs * bs + 4) * nblk + 4 * (rows // bs + 1) + 16 * rows
    else:
        byts = 12 * nnz + 4 * (rows + 1) + 8 * cols + 8 * rows
    out[name] = {"ms": ms, "GB": byts / 1e9, "GBps": byts / ms / 1e6, "frac": byts / ms / 1e6 / peak}
print(json.dumps(out))
