"""One matrix-free kernel variant, a few products (ncu target): python scripts/mf_one.py MX KERNEL REPS [TILE]
TILE = -xsb_mf_tile variant of the one-pass kernel."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import exsaddle_b200 as X
mx = int(sys.argv[1]); k = int(sys.argv[2]); reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
tile = int(sys.argv[4]) if len(sys.argv) > 4 else 0
g = X.ExSaddle("-mx %d -model 6 -eta1 1e6 -xsb_mf_kernel %d -xsb_mf_tile %d" % (mx, k, tile), nsd=3).assemble()
st = torch.cuda.ExternalStream(g.stream())
rows = g.mat_info(X.MAT_A00)[0]
x = torch.sin(0.37 * torch.arange(rows, dtype=torch.float64, device="cuda")) + 0.1
y = torch.empty(rows, dtype=torch.float64, device="cuda")
torch.cuda.synchronize()
with torch.cuda.stream(st):
    for _ in range(3):
        g.mat_mult_dev(X.MAT_A00_MF, x.data_ptr(), y.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(reps):
        g.mat_mult_dev(X.MAT_A00_MF, x.data_ptr(), y.data_ptr())
    e1.record(st)
torch.cuda.synchronize()
print(json.dumps({"mx": mx, "kernel": k, "tile": tile, "ms": e0.elapsed_time(e1) / reps, "ynorm": float(y.norm())}))
g.close()
