"""Matrix-free vs assembled A00 apply timing (development aid): python scripts/mf_bench.py MX [reps] ["extra options"]
Times the assembled BAIJ product and both matrix-free kernels (-xsb_mf_kernel 1|2) on the same input."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import exsaddle_b200 as X
mx = int(sys.argv[1]); reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
more = " " + sys.argv[3] if len(sys.argv) > 3 else ""
out = {}
ref = None
for name, which, extra in (("A00_baij", X.MAT_A00, ""), ("A00_mf_v3", X.MAT_A00_MF, " -xsb_mf_kernel 3"), ("A00_mf_onepass", X.MAT_A00_MF, " -xsb_mf_kernel 4")):
    g = X.ExSaddle("-mx %d -model 6 -eta1 1e6%s%s" % (mx, extra, more), nsd=3).assemble()
    st = torch.cuda.ExternalStream(g.stream())
    rows = g.mat_info(X.MAT_A00)[0]
    x = torch.sin(0.37 * torch.arange(rows, dtype=torch.float64, device="cuda")) + 0.1
    y = torch.empty(rows, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    with torch.cuda.stream(st):
        for _ in range(3):
            g.mat_mult_dev(which, x.data_ptr(), y.data_ptr())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(reps):
            g.mat_mult_dev(which, x.data_ptr(), y.data_ptr())
        e1.record(st)
    torch.cuda.synchronize()
    if ref is None:
        ref = y.clone()
    out[name] = {"ms": e0.elapsed_time(e1) / reps, "rel_diff_vs_baij": float((y - ref).norm() / ref.norm())}
    if "mf" in name:
        out[name]["GFLOPs"] = 2 * 5900 * mx ** 3 / out[name]["ms"] / 1e6
    g.close(); del x, y
print(json.dumps(out))
