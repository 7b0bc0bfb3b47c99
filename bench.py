#!/usr/bin/env python
"""bench.py -- headline benchmark of the exSaddle solve path on B200 (contract: see the task brief).

Metric (BASELINE.json): Stokes KSP solve time to rtol 1e-8 (s) -- FGMRES + fieldsplit Schur-upper ABF with GMG
(Chebyshev/Jacobi, Galerkin) on the velocity block -- and the Stokes MatMult HBM GB/s against the measured peak.
A "step" is one KSPSolve (zero initial guess, second-solve protocol of exSaddle.c:569-599: set-up is outside the timed
region).  The working set (x, y, viscosity, Galerkin levels; A00 BAIJ 10.4 GB on the assembled path) exceeds the 126 MB L2.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--mx 64] [--eta1 1e6] [--levels 6] [--path operator-free|assembled] [--impl reference]

`value` (default --path operator-free): neither A nor A00 is stored; every fine-level A00 product is the one-pass, TMA-staged,
sum-factorised Q2 element kernel (csrc/xsb_mf1p.cu) with the smoother update fused in.  `roofline` is that kernel against the
measured FP64 FMA peak (it is bound by FP64 issue, not by HBM or tensor throughput: SURVEY 8d), with its HBM view beside it.
`assembled`: the same solve on the reference's data layout (A in MATAIJ, A00 / Galerkin levels BAIJ(3)) with the HBM roofline of
the fine-level BAIJ SpMV and the full-A AIJ MatMult micro-measure.  `strong_128`: the north-star strong-scaling workload
(128^3, 7 levels, operator-free) at the same N.  `parity`: true residual + iteration counts / history against the committed
CPU-oracle fixture of the timed configuration; the process exits 3 if it does not hold (a wrong answer prints no valid time).
`e2e`: the same solve through the host-pointer C-ABI call (pinned host RHS -> device, solution -> host inside the timed region).

N > 1 (torchrun): the SAME problem is cut into z-slabs, one per GPU (strong scaling): NCCL halo exchange in front of
every operator apply, NCCL all-reduce behind every Krylov reduction; value = solve time, max over ranks.
--impl reference: the CPU oracle port (PETSc is not installable here), all host cores, ONE complete timed solve.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ABF = ("-saddle_ksp_type fgmres -fs -saddle_fieldsplit_u_pc_type mg -saddle_fieldsplit_u_ksp_type gcr "
       "-saddle_fieldsplit_u_ksp_rtol 1e-2 -saddle_fieldsplit_u_mg_levels_pc_type jacobi "
       "-saddle_fieldsplit_u_mg_levels_ksp_type chebyshev -saddle_fieldsplit_u_mg_levels_ksp_chebyshev_esteig 0,0.2,0,1.1 "
       "-saddle_fieldsplit_u_mg_levels_ksp_max_it 8 -saddle_fieldsplit_u_mg_levels_ksp_norm_type none "
       "-saddle_fieldsplit_u_pc_mg_galerkin -saddle_fieldsplit_p_ksp_type preonly -saddle_fieldsplit_p_pc_type bjacobi")
ITERS_FILE = os.path.join(ROOT, "profiles", "bench_iterations.json")


def workload_options(a):
    extra = os.environ.get("XSB_BENCH_EXTRA_OPTS", "")   # experiments only (library tuning switches); printed with the options in `config`
    return "%s -saddle_fieldsplit_u_pc_mg_levels %d -model 6 -mx %d -eta0 1 -eta1 %g -saddle_ksp_rtol 1e-8%s" % (ABF, a.levels, a.mx, a.eta1, (" " + extra) if extra else "")


# measured once per kernel change with `ncu --set full` (dram__bytes_read.sum + dram__bytes_write.sum per launch), see profiles/
NCU_TRAFFIC = {("spmv_baij", 64, 1): 10.81e9, ("mf_onepass", 64, 1): 354.9e6}   # mf_onepass: Chebyshev-step launch, profiles/r02_mf_onepass_64cubed_chebstep_raw.csv (305.8 MB read + 49.1 MB written)
FP64_PEAK_TFLOPS = 36.72      # scripts/fp64_peak.cu on this pool's B200 (profiles/r01_fp64_peak.json): 63.1 DFMA/clk/SM at 1965 MHz
MF_FLOP_PER_ELEMENT = 5808.0  # FP64 flops the one-pass element kernel executes per element: 3 lanes x (723 DFMA x 2 + 376 DMUL + 114 DADD), cuobjdump -sass (zero / unit table entries skipped)


def a00_csr_bytes(mx, world=1):
    """AIJ (CSR) bytes of one A00 product at mx^3, SURVEY 8d: 12 B/nnz + 4 B/row pointer + 16 B/row (x, y); the common yardstick
    ("AIJ-equivalent GB/s") for the BAIJ and the matrix-free products.  nnz(A00) = 9 * (sum over a node line of its coupling width)^3."""
    n1 = 2 * mx + 1
    per_dir = sum(3 if (i % 2 or i in (0, n1 - 1)) else 5 for i in range(n1))
    nnz = 9 * per_dir ** 3; rows = 3 * n1 ** 3
    return (12.0 * nnz + 4.0 * (rows + 1) + 16.0 * rows) / world


def timed_solves(g, torch, xdev, steps, barrier):
    """K solves, device-resident RHS, CUDA events on the handle's own stream; returns (ms, counters summed over the steps)."""
    stream = torch.cuda.ExternalStream(g.stream(), device=xdev.device)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0; a00_ns = []; modes = [0, 0, 0, 0]; n_a00 = 0
    with torch.cuda.stream(stream):
        ev0.record(stream)
        for _ in range(steps):
            g.solve_dev(0, xdev.data_ptr())
            c = g.counters(); launches += c["launches"]; a00_ns.append(c["a00_avg_ns"]); n_a00 += c["a00_spmv"]
            modes = [u + v for u, v in zip(modes, c["a00_by_mode"])]
        ev1.record(stream)
    barrier()
    return ev0.elapsed_time(ev1), launches, sum(a00_ns) / len(a00_ns), n_a00, modes


def timed_e2e(g, X, torch, device, F_host, x_host, steps, barrier):
    """K solves through the host-pointer C-ABI call: pinned host RHS -> device, solution -> host inside the timed region."""
    stream = torch.cuda.ExternalStream(g.stream(), device=device)
    barrier()
    with torch.cuda.stream(stream):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            g._chk(g.L.xsb_ksp_solve(g.h, X.api._dp(F_host.numpy()), X.api._dp(x_host.numpy())))
        e1.record(stream)
    barrier()
    return e0.elapsed_time(e1)


def parity_check(g, X, torch, dist, a, world, xdev, its, reason, inner, hist, lame=False):
    """Parity of the TIMED configuration, inside the bench (so a wrong answer cannot print a time):
    (1) true residual ||F - A x|| / ||F|| of the last solve, computed on the slabs (owned rows, all-reduced);
    (2) outer / inner iteration counts and the residual history against the committed CPU-oracle fixture of the same
        configuration and the same number of bjacobi blocks (tests/golden/oracle_<mx>cubed_history.json for one GPU,
        oracle_64cubed_bjacobi_blocks.json[N] for N GPUs: ILU(0) per rank = per z-slab block in the oracle).
    History tolerance: 1e-5 relative while the residual is above 1e-2 ||r0|| -- the oracle-vs-oracle floor at eta1/eta0 = 1e6 is
    2e-6 (profiles/r02_oracle_drift_64cubed.json); counts: outer +-1, inner counts equal up to one iteration in total."""
    part = g.partition()
    n = g.n
    F = torch.from_numpy(g.rhs()).cuda()
    y = torch.empty(n, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    g.mat_mult_dev(X.MAT_A, xdev.data_ptr(), y.data_ptr())
    torch.cuda.synchronize()
    r = F - y
    own = torch.cat([r[part["u_off"]:part["u_off"] + part["u_len"]], r[part["p_off"]:part["p_off"] + part["p_len"]]])
    ownF = torch.cat([F[part["u_off"]:part["u_off"] + part["u_len"]], F[part["p_off"]:part["p_off"] + part["p_len"]]])
    sums = torch.stack([(own * own).sum(), (ownF * ownF).sum()])
    if dist is not None:
        dist.all_reduce(sums)
    true_res = float((sums[0] / sums[1]).sqrt())
    out = {"true_rel_residual": true_res, "outer_its": int(its), "reason": int(reason), "fixture": None}
    ok = reason == 2 and true_res <= 3e-8      # rtol 1e-8 on the FGMRES residual estimate; the true residual tracks it within a small factor
    fx = None
    gold = os.path.join(ROOT, "tests", "golden")
    if not lame and a.eta1 == 1e6:
        if world == 1 and os.path.exists(os.path.join(gold, "oracle_%dcubed_history.json" % a.mx)):
            d = json.load(open(os.path.join(gold, "oracle_%dcubed_history.json" % a.mx)))
            if d.get("levels") == a.levels:
                fx = {"its": d["its"], "inner": d["inner_its"], "hist": d["hist"]}; out["fixture"] = "oracle_%dcubed_history.json" % a.mx
        elif world > 1 and a.mx == 64 and a.levels == 6:
            d = json.load(open(os.path.join(gold, "oracle_64cubed_bjacobi_blocks.json"))).get(str(world))
            if d:
                fx = {"its": d["its"], "inner": d["inner"], "hist": d.get("hist")}; out["fixture"] = "oracle_64cubed_bjacobi_blocks.json[%d]" % world
    if fx:
        m = min(len(inner), len(fx["inner"]))
        inner_diff = sum(abs(u - v) for u, v in zip(inner[:m], fx["inner"][:m]))
        out.update({"oracle_outer_its": fx["its"], "inner_count_diff_total": int(inner_diff)})
        ok = ok and abs(its - fx["its"]) <= 1 and inner_diff <= 1
        if fx["hist"]:
            k = min(len(hist), len(fx["hist"]))
            rel = [abs(hist[i] - fx["hist"][i]) / fx["hist"][i] for i in range(k) if fx["hist"][i] >= 1e-2 * fx["hist"][0]]
            out["max_rel_hist_diff_above_1e-2"] = max(rel) if rel else None
            ok = ok and (not rel or max(rel) <= 1e-5)
    out["ok"] = bool(ok)
    return out


def workload_name(a):
    return "exSaddle3d Stokes sinker (model 6) %d^3 Q2-Q1, eta1/eta0=%g, FGMRES+fieldsplit Schur-upper ABF, GMG %d levels Chebyshev(8)/Jacobi Galerkin, rtol 1e-8" % (a.mx, a.eta1, a.levels)


def config_key(a):
    return "mx%d_eta%g_lv%d" % (a.mx, a.eta1, a.levels)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True); self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([t.strip() for t in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


def a00_bytes(info, by_mode):
    """Algorithmic bytes of one fine-level A00 BAIJ(3) launch, averaged over the launches of the solve:
    (72+4) B per block + 4 B per block-row pointer + 8 B per vector entry streamed (x,y [+b,+idiag,+p_{k-1}])."""
    rows, _, nnz, bs = info
    nblk = nnz // (bs * bs)
    mat = (8 * bs * bs + 4) * nblk + 4 * (rows // bs + 1)
    vecs = {0: 2, 1: 3, 2: 4, 3: 5}
    tot = sum(by_mode)
    if tot == 0:
        return mat + 16 * rows
    return mat + 8 * rows * sum(vecs[m] * by_mode[m] for m in range(4)) / tot


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_reference(a, full_solve):
    """The reference's CPU implementation of the path.  PETSc is not installable here (no PETSc/MPI in the image), so this is
    the oracle port (oracle/xo_*.c, OpenMP), on ALL host cores: the team size is set explicitly because torchrun exports
    OMP_NUM_THREADS=1 (round 1's reference arm ran single-threaded under torchrun and timed out).

    full_solve=True  (--impl reference): ONE complete solve of the workload, timed (exSaddle.c:569-599 protocol without the
                     warm-up solve: a CPU solve has no launch/JIT warm-up to hide and one solve is minutes); value = its seconds.
    full_solve=False (cpu_baseline of the GPU arm): a bounded sample -- the first `sample_outer` outer FGMRES iterations --
                     scaled to the full solve by operator products (cost ~ n_A00 + 1.66 n_A; 1.66 = AIJ bytes of A / CSR bytes
                     of A00), the full solve's counts taken from the committed oracle fixture.  Labelled as an estimate."""
    from oracle import oracle as O
    O.lib().xo_set_num_threads(host_cores())
    cores = O.lib().xo_num_threads()
    opts = workload_options(a)
    t0 = time.time()
    p = O.Problem(opts, nsd=3)
    s = p.solver()
    p.pc_setup(s)
    t_setup = time.time() - t0
    if full_solve:
        x, res = p.solve(s)
        value = res.solve_seconds if res.reason > 0 else None
        base = {"value": value, "unit": "s", "cores": cores, "kind": "port",
                "sample": "ONE complete solve of the same %d^3 system on the oracle port (OpenMP, %d threads): %d outer FGMRES iterations, %d inner GCR iterations, "
                          "%d A00 + %d full-A products, converged reason %d; CPU set-up %.1f s not included" % (a.mx, cores, res.its, sum(res.inner_its[:res.n_inner]), res.n_a00_mult, res.n_a_mult, res.reason, t_setup),
                "outer_its": int(res.its), "reason": int(res.reason), "steps_run": 1, "setup_s": t_setup}
        return base
    full = None   # (outer its, A00 products, full-A products) of the complete solve
    fx = os.path.join(ROOT, "tests", "golden", "oracle_%dcubed_history.json" % a.mx)
    if os.path.exists(fx):
        d = json.load(open(fx))
        if d.get("levels") == a.levels and "-eta1 %g " % a.eta1 in d.get("options", "").replace("1e6", "1e+06") + " ":
            full = (d["its"], 17 * sum(d["inner_its"]), d["its"] + 1 + d["its"] // 30, "oracle fixture")
    if full is None and os.path.exists(ITERS_FILE):
        e = json.load(open(ITERS_FILE)).get(config_key(a), {})
        if e.get("outer_its"):
            full = (e["outer_its"], 17 * e.get("inner_its_total", 0), e["outer_its"] + 1 + e["outer_its"] // 30, "GPU arm's counters")
    s.max_outer_sample = max(1, a.sample_outer)
    x, res = p.solve(s)
    t_sample = res.solve_seconds
    work = lambda n00, nA: n00 + 1.66 * nA
    w_sample = work(res.n_a00_mult, res.n_a_mult)
    if res.reason > 0:        # the sample converged: it IS the full solve
        value, how = t_sample, "complete solve (%d outer iterations)" % res.its
    elif full is not None and w_sample > 0:
        value = t_sample * work(full[1], full[2]) / w_sample
        how = ("ESTIMATE: scaled by operator products to the full solve's %d outer / %d inner iterations = %d A00 + %d full-A products (%s); "
               "the measured complete solve is what `bench.py --impl reference` reports" % (full[0], full[1] // 17, full[1], full[2], full[3]))
    else:
        value, how = None, "no iteration counts of the full solve available"
    return {"value": value, "unit": "s", "cores": cores, "kind": "port", "estimated": res.reason <= 0,
            "sample": "first %d outer FGMRES iterations (%d A00 + %d full-A products, %.2f s) of the same %d^3 system on the oracle port (OpenMP, %d threads), %s; CPU set-up %.1f s not included"
                      % (res.its, res.n_a00_mult, res.n_a_mult, t_sample, a.mx, cores, how, t_setup)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--mx", type=int, default=64)
    ap.add_argument("--eta1", type=float, default=1e6)
    ap.add_argument("--levels", type=int, default=6)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--sample-outer", dest="sample_outer", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-assembled", dest="no_assembled", action="store_true", help="skip the assembled (AIJ / BAIJ) path measured beside the headline")
    ap.add_argument("--no-strong128", dest="no_strong128", action="store_true", help="skip the 128^3 operator-free block (north-star strong-scaling workload)")
    ap.add_argument("--path", default="operator-free", choices=["operator-free", "assembled"], help="which path is the headline `value`")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("NCCL_DEBUG", "WARN")   # keep NCCL's version banner off stdout: rank 0 prints exactly one JSON line
    warmup = max(a.warmup, 3) if a.impl == "b200" else a.warmup
    fits_assembled = 5420.0 * a.mx ** 3 / max(1, world) < 2.0e9   # the AIJ operator needs nnz(A) per rank < 2^31 (32-bit PetscInt)
    if a.path == "assembled" and not fits_assembled:
        raise SystemExit("--path assembled: nnz(A) per rank exceeds 32-bit indices at %d^3 on %d GPU(s)" % (a.mx, world))
    headline_opfree = a.path == "operator-free"

    path_text = {True: "operator-free: neither A nor A00 stored; fine-level A00 products (smoother, GCR, outer MatMult) by the one-pass TMA-staged sum-factorised Q2 element kernel, "
                       "first Galerkin level assembled element by element, coarse levels BAIJ(3)",
                 False: "assembled: A in the reference's MATAIJ layout, A00 and Galerkin levels BAIJ(3)"}
    cfg = {"workload": workload_name(a), "path": path_text[headline_opfree], "unknowns": 3 * (2 * a.mx + 1) ** 3 + (a.mx + 1) ** 3, "mg_levels": a.levels,
           "parallelism": "1 GPU" if world == 1 else "one problem, z-slab partition over %d GPUs (one process per GPU): ghost planes exchanged by one kernel over NVLink peer memory (cudaIpc windows, flag handshake; -xsb_p2p 0: ncclSend/ncclRecv) before every fine-level operator apply, NCCL all-reduce for Krylov dot products; fine MG level on the slab lattice, large coarse levels distributed by node planes (one plane of every product's result traded with each neighbour), small ones replicated, ILU(0) per rank (bjacobi)" % world,
           "l2": "x, y, viscosity and the Galerkin levels (>= 0.7 GB at 64^3; assembled path: A00 BAIJ 10.4 GB) exceed the 126 MB L2; no flush needed",
           "options": workload_options(a)}

    if a.impl == "reference":
        if rank != 0:
            return 0
        base = cpu_reference(a, True)
        line = {"impl": "reference", "metric": "stokes_ksp_solve_time_rtol1e-8", "value": base["value"], "unit": "s", "n_gpus": a.gpus,
                "steps": 1, "warmup": 0, "steps_requested": a.steps, "warmup_requested": a.warmup,
                "ms_per_step": None if base["value"] is None else 1e3 * base["value"],
                "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": cfg, "cpu_baseline": base,
                "e2e": {"value": base["value"], "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        print(json.dumps(line)); return 0

    import numpy as np
    import torch
    import exsaddle_b200 as X
    if not torch.cuda.is_available() or not X.device_available():
        raise SystemExit("bench.py needs a CUDA device: exsaddle_b200 has no CPU path")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def make(b, extra):
        h = X.ExSaddle(workload_options(b) + extra, nsd=3, device=local)
        if world > 1:   # one problem, z-slabs over the ranks: NCCL communicator inside the library, id shipped by torch.distributed
            uid = [X.comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(uid, src=0)
            h.comm_init(uid[0], rank, world)
        return h

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(vals):
        t = torch.tensor(vals, dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t]

    peak, peak_src = peaks()

    def measure(b, opfree, steps, wu, sampler=None, e2e=True, instrument=True, micro=False):
        """One path of one workload: K timed solves (device-resident, CUDA events, V-cycles replayed from CUDA graphs), parity,
        end-to-end solves through the host-pointer call, and one instrumented solve (-xsb_time_kernels: an event pair around every
        fine-level A00 product, graphs off) for the live kernel time of the roofline."""
        g = make(b, " -xsb_matrix_free full" if opfree else "")
        t0 = time.time(); g.assemble(); t_asm = time.time() - t0
        part = g.partition(); own_frac = part["u_len"] / float(g.nu)
        t0 = time.time(); g.ksp_setup(); t_setup = time.time() - t0
        n = g.n
        xdev = torch.empty(n, dtype=torch.float64, device="cuda")
        for _ in range(wu):
            g.solve_dev(0, xdev.data_ptr())
        barrier()
        if sampler:
            sampler.start()
        ms, launches, _, n_a00, modes = timed_solves(g, torch, xdev, steps, barrier)
        clocks = sampler.stop() if sampler else None
        its, reason = g.iterations(); inner = g.inner_iterations(); hist = g.history()
        parity = parity_check(g, X, torch, dist, b, world, xdev, its, reason, inner, [float(v) for v in hist])
        ci = g.comm_info()
        out = {"comm": {"peer_memory_halo": ci["p2p"], "plane_distributed_mg_levels": ci["pdist_levels"]}, "outer_its": int(its), "reason": int(reason), "inner_gcr_its": int(sum(inner)), "rnorm0": float(hist[0]), "rnorm": float(hist[-1]),
               "a00_products_per_solve": n_a00 // steps, "assemble_s": t_asm, "ksp_setup_s": t_setup, "gpu_launches": launches, "parity": parity, "clocks": clocks}
        ms_e2e = None
        if e2e:
            F_host = torch.from_numpy(g.rhs()).pin_memory(); x_host = torch.empty(n, dtype=torch.float64).pin_memory()
            ms_e2e = timed_e2e(g, X, torch, xdev.device, F_host, x_host, steps, barrier)
            del F_host, x_host
        if micro:   # full-A MatMult, 10 warm-up + 30 timed applies (collective on slabs)
            a_info = g.mat_info(X.MAT_A)
            xin = torch.sin(0.37 * torch.arange(n, dtype=torch.float64, device="cuda")) + 0.1
            yout = torch.empty_like(xin)
            stream = torch.cuda.ExternalStream(g.stream(), device=xdev.device)
            stream.wait_stream(torch.cuda.current_stream())   # xin is produced on torch's stream, consumed on the library's
            with torch.cuda.stream(stream):
                for _ in range(10):
                    g.mat_mult_dev(X.MAT_A, xin.data_ptr(), yout.data_ptr())
                m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                m0.record(stream)
                for _ in range(30):
                    g.mat_mult_dev(X.MAT_A, xin.data_ptr(), yout.data_ptr())
                m1.record(stream)
            torch.cuda.synchronize()
            aij_ms = m0.elapsed_time(m1) / 30
            aij_bytes = (12 * a_info[2] + 4 * (a_info[0] + 1) + 16 * a_info[0]) * own_frac
            out["aij_matmult"] = {"kernel": "spmv_csr_kernel (full saddle A, AIJ layout)" if not opfree else "element kernel (A00) + spmv_csr_kernel on A01 / A10 / A11",
                                  "ms": aij_ms, "bytes": aij_bytes, "achieved": aij_bytes / (aij_ms * 1e6), "frac": aij_bytes / (aij_ms * 1e6) / peak,
                                  "frac_of_nominal_8TBps": aij_bytes / (aij_ms * 1e6) / 8000.0}
            del xin, yout
        ms, ms_e2e = allmax([ms, ms_e2e if ms_e2e is not None else 0.0])
        out["value"] = ms / 1e3 / steps; out["e2e"] = ms_e2e / 1e3 / steps if e2e else None
        if instrument:
            a00_info = g.mat_info(X.MAT_A00); nel_local = g.nel
            g.set_option("-xsb_time_kernels"); g.ksp_setup()
            g.solve_dev(0, xdev.data_ptr()); barrier()
            ms_i, _, avg_ns, n_i, modes_i = timed_solves(g, torch, xdev, 1, barrier)
            share = (n_i * avg_ns * 1e-9) / (ms_i / 1e3)
            # device time of that instrumented solve by category (graphs off, rank 0's view): where the step goes
            out["profile"] = {"solve_ms": ms_i, "by_category_ms": {k: round(v[0], 3) for k, v in g.profile().items()}, "stretches": {k: int(v[1]) for k, v in g.profile().items()},
                              "note": "one extra solve with -xsb_time_kernels (an event at every category change, CUDA graphs off), this rank's device time"}
            if opfree:
                nel_apply = nel_local * own_frac if world > 1 else nel_local   # slabs: the kernel applies the layers touching owned planes
                flops_alg = 9000.0 * nel_apply            # ~9.0 kflop per element, sum-factorised
                flops_exec = MF_FLOP_PER_ELEMENT * nel_apply
                hbm_bytes = (8.0 * 27 * a.mx ** 3 if b.mx == a.mx else 8.0 * 27 * b.mx ** 3) / world
                vecs = {0: 2, 1: 3, 2: 4, 3: 5}; tot = max(1, sum(modes_i))
                hbm_bytes += 8.0 * 3 * (2 * b.mx + 1) ** 3 / world * sum(vecs[m] * modes_i[m] for m in range(4)) / tot
                ach = flops_alg / max(avg_ns, 1) / 1e3
                out["roofline"] = {"bound": "fp64", "kernel": "mf_onepass_kernel + mf_shared_kernel (one-pass TMA-staged sum-factorised Q2 element apply of A00 with fused smoother epilogue)",
                                   "achieved": ach, "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": ach / FP64_PEAK_TFLOPS,
                                   "traffic": NCU_TRAFFIC.get(("mf_onepass", b.mx, world)),
                                   "peak_source": "measured FP64 FMA micro-kernel (scripts/fp64_peak.cu, profiles/r01_fp64_peak.json); not in MEASURED_PEAKS.json",
                                   "flops_per_launch": flops_alg, "flops_per_launch_source": "SURVEY 8(d): 9.0 kflop per element (sum-factorised count); executed per SASS count: %.0f flop per element = %.3g per launch" % (MF_FLOP_PER_ELEMENT, flops_exec),
                                   "achieved_executed_flops": flops_exec / max(avg_ns, 1) / 1e3,
                                   "avg_launch_us": avg_ns / 1e3, "launches_timed": n_i, "share_of_step": share,
                                   "hbm_view": {"algorithmic_bytes": hbm_bytes, "achieved_GBps": hbm_bytes / max(avg_ns, 1), "frac_of_measured_peak": hbm_bytes / max(avg_ns, 1) / peak},
                                   "aij_equivalent_GBps": a00_csr_bytes(b.mx, world) / max(avg_ns, 1)}
            else:
                bytes_launch = a00_bytes(a00_info, modes_i) * own_frac   # owned rows only are streamed (own_frac = 1 on one GPU)
                achieved = bytes_launch / avg_ns if avg_ns else 0.0     # B/ns = GB/s
                out["roofline"] = {"bound": "hbm", "kernel": "spmv_baij_kernel<3> (fine-level A00 with fused residual/Chebyshev epilogue)",
                                   "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": NCU_TRAFFIC.get(("spmv_baij", b.mx, world)),
                                   "peak_source": peak_src, "bytes_per_launch": bytes_launch, "avg_launch_us": avg_ns / 1e3,
                                   "launches_timed": n_i, "share_of_step": share, "aij_equivalent_GBps": a00_csr_bytes(b.mx, world) / max(avg_ns, 1)}
        g.close(); del xdev
        torch.cuda.empty_cache()
        return out

    sampler = ClockSampler(local)
    head = measure(a, headline_opfree, a.steps, warmup, sampler=sampler, micro=not headline_opfree)
    other = None
    if headline_opfree and fits_assembled and not a.no_assembled:
        other = measure(a, False, a.steps, 2, micro=True)
    elif not headline_opfree:
        other = measure(a, True, a.steps, 2)
    strong = None
    if a.mx == 64 and not a.no_strong128:
        import copy
        b = copy.copy(a); b.mx, b.levels = 128, 7
        s128 = measure(b, True, 3, 1, e2e=False, instrument=True)
        strong = {"workload": workload_name(b), "path": "operator-free", "value": s128["value"], "unit": "s", "steps": 3, "warmup": 1, "outer_its": s128["outer_its"],
                  "inner_gcr_its": s128["inner_gcr_its"], "true_rel_residual": s128["parity"]["true_rel_residual"], "parity_ok": s128["parity"]["ok"],
                  "profile": s128.get("profile"), "comm": s128.get("comm"), "element_kernel_avg_us": (s128.get("roofline") or {}).get("avg_launch_us"),
                  "note": "north-star strong-scaling workload (128^3, 7 MG levels): same value at every --gpus N, divide N = 1 by N x this for the efficiency"}

    if rank == 0:
        os.makedirs(os.path.dirname(ITERS_FILE), exist_ok=True)
        try:   # outer iteration count of the ONE-rank solve: what the CPU sample is scaled by (per-rank ILU changes it on slabs)
            if world > 1:
                raise RuntimeError("keep the one-rank count")
            d = json.load(open(ITERS_FILE)) if os.path.exists(ITERS_FILE) else {}
            d[config_key(a)] = {"outer_its": head["outer_its"], "inner_its_total": head["inner_gcr_its"], "reason": head["reason"]}
            json.dump(d, open(ITERS_FILE, "w"), indent=1, sort_keys=True)
        except Exception:
            pass
        base = None
        if world == 1 and not a.no_cpu_baseline:
            try:
                import psutil
                need_gb = 45.0 * (a.mx / 64.0) ** 3
                if psutil.virtual_memory().available / 2 ** 30 > need_gb + 8:
                    base = cpu_reference(a, False)
                else:
                    base = {"value": None, "unit": "s", "cores": os.cpu_count(), "kind": "port", "sample": "skipped: host RAM below %.0f GB" % need_gb}
            except Exception as e:   # the checker must never take the bench line down
                base = {"value": None, "unit": "s", "cores": os.cpu_count(), "kind": "port", "sample": "failed: %r" % (e,)}
        n = cfg["unknowns"]
        line = {"metric": "stokes_ksp_solve_time_rtol1e-8", "value": head["value"], "unit": "s", "n_gpus": world, "steps": a.steps,
                "warmup": warmup, "ms_per_step": 1e3 * head["value"], "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": cfg,
                "solve": {k: head[k] for k in ("outer_its", "reason", "inner_gcr_its", "rnorm0", "rnorm", "a00_products_per_solve", "assemble_s", "ksp_setup_s")},
                "parity": head["parity"], "roofline": head.get("roofline"), "profile": head.get("profile"), "comm": head.get("comm"),
                ("assembled" if headline_opfree else "operator_free"): None if other is None else {k: other.get(k) for k in ("value", "e2e", "outer_its", "inner_gcr_its", "gpu_launches", "parity", "roofline", "aij_matmult")},
                "strong_128": strong, "cpu_baseline": base,
                "e2e": {"value": head["e2e"], "unit": "s", "h2d_bytes_per_step": 8 * n, "d2h_bytes_per_step": 8 * n},
                "gpu_launches": head["gpu_launches"], "clocks": head["clocks"]}
        if "aij_matmult" in head:
            line["roofline"]["aij_matmult"] = head["aij_matmult"]
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    bad = [k for k, v in (("headline", head["parity"]), ("other", other["parity"] if other else None)) if v is not None and not v["ok"]]
    if strong is not None and not strong["parity_ok"]:
        bad.append("strong_128")
    if bad:
        sys.stderr.write("bench.py: PARITY FAILED in %s: %s\n" % (bad, json.dumps(head["parity"])))
        return 3
    return 0


if __name__ == "__main__":
    sys.exit(main())
