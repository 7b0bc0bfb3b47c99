"""Slab-partition parity check, run under torchrun with N >= 2 GPUs of one box:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 scripts/slab_check.py
Every rank assembles its z-slab; rank 0 also holds the single-GPU problem.  Checks, in the natural (one-rank) ordering:
MatMult <= 1e-12, RHS equal, Jacobi-GMRES and ABF residual histories / iteration counts / solutions against one GPU."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import exsaddle_b200 as X

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
MX = int(sys.argv[1]) if len(sys.argv) > 1 else 8
ABF = ("-saddle_ksp_type fgmres -fs -saddle_fieldsplit_u_pc_type mg -saddle_fieldsplit_u_ksp_type gcr -saddle_fieldsplit_u_ksp_rtol 1e-2 "
       "-saddle_fieldsplit_u_pc_mg_levels 3 -saddle_fieldsplit_u_mg_levels_pc_type jacobi -saddle_fieldsplit_u_mg_levels_ksp_type chebyshev "
       "-saddle_fieldsplit_u_mg_levels_ksp_chebyshev_esteig 0,0.2,0,1.1 -saddle_fieldsplit_u_mg_levels_ksp_max_it 8 "
       "-saddle_fieldsplit_u_mg_levels_ksp_norm_type none -saddle_fieldsplit_u_pc_mg_galerkin -saddle_fieldsplit_p_ksp_type preonly ")


def make(opts, distributed):
    g = X.ExSaddle(opts, nsd=3, device=local)
    if distributed:
        uid = [X.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        g.comm_init(uid[0], rank, world)
    return g.assemble()


def to_local(g, part, xglob, nu_g):
    """natural-order global vector -> this rank's local [u|p] vector (owned + ghost entries filled from the global one)."""
    L = g.n; x = np.zeros(L)
    nu_loc = part["p_off"] - (part["p_off"] - g.nu) if False else g.nu
    # local u lattice starts at global velocity plane 2*e0, local p lattice at global pressure plane e0
    pu = part["u_len"] // max(1, (2 * (part["k1"] - part["k0"]) + (1 if part["rank"] == part["nranks"] - 1 else 0)))
    pp = part["p_len"] // max(1, ((part["k1"] - part["k0"]) + (1 if part["rank"] == part["nranks"] - 1 else 0)))
    u0 = 2 * part["e0"] * pu; p0 = part["e0"] * pp
    x[:g.nu] = xglob[u0:u0 + g.nu]
    x[g.nu:] = xglob[nu_g + p0: nu_g + p0 + g.np_]
    return x


def gather_owned(g, part, xloc, n_glob, nu_g):
    """owned entries of every rank -> natural-order global vector on all ranks."""
    pieces = [None] * world
    mine = (part["u_glob0"], xloc[part["u_off"]:part["u_off"] + part["u_len"]].copy(), nu_g + part["p_glob0"], xloc[part["p_off"]:part["p_off"] + part["p_len"]].copy())
    dist.all_gather_object(pieces, mine)
    out = np.full(n_glob, np.nan)
    for u0, u, p0, p in pieces:
        out[u0:u0 + len(u)] = u; out[p0:p0 + len(p)] = p
    assert not np.isnan(out).any(), "owned ranges do not tile the global vector"
    return out


results = {}
ok = True
for name, opts, ptype in (("jacobi_gmres", "-model 1 -mx %d -eta1 10 -saddle_pc_type jacobi -saddle_ksp_max_it 25" % MX, "jacobi"),
                          ("abf_pjacobi", ABF + "-saddle_fieldsplit_p_pc_type jacobi -model 6 -mx %d -eta1 100 -saddle_ksp_rtol 1e-8" % MX, "abf"),
                          ("abf_bjacobi_ilu", ABF + "-saddle_fieldsplit_p_pc_type bjacobi -model 6 -mx %d -eta1 100 -saddle_ksp_rtol 1e-8" % MX, "abf_ilu"),
                          ("abf_pjacobi_mf", ABF + "-saddle_fieldsplit_p_pc_type jacobi -xsb_matrix_free -model 6 -mx %d -eta1 100 -saddle_ksp_rtol 1e-8" % MX, "abf"),
                          ("abf_pjacobi_mffull", ABF + "-saddle_fieldsplit_p_pc_type jacobi -xsb_matrix_free full -model 6 -mx %d -eta1 100 -saddle_ksp_rtol 1e-8" % MX, "abf"),
                          # coarse levels distributed by node planes: level 1 of 3 (gathers into the replicated coarsest level), levels 2 and 1 of 4
                          ("abf_pjacobi_pdist3", ABF + "-saddle_fieldsplit_p_pc_type jacobi -xsb_pdist_min_nodes 1 -model 6 -mx %d -eta1 100 -saddle_ksp_rtol 1e-8" % MX, "abf"),
                          ("abf_pjacobi_pdist4_mffull", ABF + "-saddle_fieldsplit_u_pc_mg_levels 4 -saddle_fieldsplit_p_pc_type jacobi -xsb_pdist_min_nodes 1 -xsb_matrix_free full -model 6 -mx %d -eta1 100 -saddle_ksp_rtol 1e-8" % MX, "abf"),
                          # the same with ncclSend / ncclRecv instead of the peer-memory exchange kernel
                          ("abf_pjacobi_pdist4_nccl", ABF + "-saddle_fieldsplit_u_pc_mg_levels 4 -saddle_fieldsplit_p_pc_type jacobi -xsb_pdist_min_nodes 1 -xsb_p2p 0 -model 6 -mx %d -eta1 100 -saddle_ksp_rtol 1e-8" % MX, "abf")):
    gd = make(opts, True)
    part = gd.partition()
    g1 = make(opts, False)          # every rank keeps a one-GPU copy as the reference (small problem)
    n_glob, nu_g = g1.n, g1.nu
    # RHS and MatMult in natural ordering
    F = gather_owned(gd, part, gd.rhs(), n_glob, nu_g)
    eF = np.linalg.norm(F - g1.rhs()) / np.linalg.norm(g1.rhs())
    xg = np.sin(0.37 * np.arange(n_glob)) + 0.1
    y = gather_owned(gd, part, gd.mat_mult(X.MAT_A, to_local(gd, part, xg, nu_g)), n_glob, nu_g)
    y1 = g1.mat_mult(X.MAT_A, xg)
    eA = np.linalg.norm(y - y1) / np.linalg.norm(y1)
    gd.ksp_setup(); g1.ksp_setup()
    xs = gather_owned(gd, part, gd.solve(), n_glob, nu_g)
    x1 = g1.solve()
    h, h1 = gd.history(), g1.history()
    n = min(len(h), len(h1))
    res = {"rhs_err": eF, "matmult_err": eA, "its": gd.iterations(), "its_1gpu": g1.iterations(), "inner": gd.inner_iterations()[:8], "inner_1gpu": g1.inner_iterations()[:8],
           "hist_err_ksp_rel": float(np.max(np.abs(h[:n] - h1[:n])) / h1[0]), "hist_err_first8": float(np.max(np.abs(h[:8] - h1[:8]) / h1[:8])), "sol_err": float(np.linalg.norm(xs - x1) / np.linalg.norm(x1)),
           "true_res": float(np.linalg.norm(g1.rhs() - g1.mat_mult(X.MAT_A, xs)) / np.linalg.norm(g1.rhs()))}
    good = eF <= 1e-13 and eA <= 1e-12
    if ptype != "abf_ilu":   # bjacobi/ILU is per rank: iteration counts legitimately change with the rank count
        # same algorithm, different reduction order: the first iterations agree to rounding; later ones drift by
        # (condition number) x eps, which can move the last iteration across the tolerance (seen at 1 vs 2 ranks in
        # the reference too: testref/exSaddle2d_lame_1 vs _2)
        good = good and abs(gd.iterations()[0] - g1.iterations()[0]) <= 2 and gd.iterations()[1] == g1.iterations()[1] \
            and res["hist_err_first8"] <= 1e-9 and res["hist_err_ksp_rel"] <= 1e-6 and res["sol_err"] <= 1e-5
    else:
        good = good and gd.iterations()[1] == 2 and res["true_res"] <= 2e-8
    res["ok"] = bool(good); ok = ok and good
    results[name] = res
    gd.close(); g1.close()
if rank == 0:
    for k, v in results.items():
        print(k, json.dumps(v))
    os.makedirs("gpurun_out", exist_ok=True); json.dump(results, open("gpurun_out/slab_check_%d.json" % world, "w"), indent=1)
    print("SLAB_CHECK", "PASS" if ok else "FAIL")
dist.barrier(); dist.destroy_process_group()
sys.exit(0 if ok else 1)
