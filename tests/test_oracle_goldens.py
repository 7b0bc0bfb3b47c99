"""Pin the CPU oracle to the reference's own golden outputs.

tests/golden/testref_kat.json holds the numbers extracted from /root/reference/testref/*.ref and the
option strings of /root/reference/Makefile:254-513 (tests/golden/make_golden.py).  No GPU needed.
"""
import json
import os

import numpy as np
import scipy.sparse as sp
import pytest

from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _sig(x, ref_text):
    """relative tolerance implied by the digits printed in the golden (%g -> 6 significant)."""
    return 6e-6 * abs(float(ref_text))


def _nums(lines):
    import re
    return np.array([float(t) for l in lines for t in re.findall(r"[+-]\d\.\d+e[+-]\d+", l)])


def _problem(kat, name, extra=""):
    c = kat[name]
    opts = c["options"].replace("-options_file abf.opts", " ".join(kat["_abf_opts"]))
    return O.Problem(opts + " " + extra, nsd=c["nsd"], lame=c["lame"]), c


# ---- C.1: ||F||_2 = iteration-0 residual of every unpreconditioned-norm (right PC) golden ----------------
RIGHT_PC_CASES = ["exSaddle3d_ar_1", "exSaddle3d_pseudoice_1", "exSaddle2d_ar_1", "exSaddle3d_mg_1", "exSaddle2d_mg_1",
                  "exSaddle3d_lame_mg_1", "exSaddle2d_lame_mg_1", "exSaddle3d_lame_3", "exSaddle3d_lame_4",
                  "exSaddle3d_lame_5", "exSaddle3d_ildl_1"]


@pytest.mark.parametrize("name", RIGHT_PC_CASES)
def test_rhs_norm_matches_golden(kat, name):
    c = kat[name]
    # only the discretisation options matter for ||F||; solver options are ignored here
    o = O.parse_options(c["options"].replace("-options_file abf.opts", ""))
    p = O.Problem({k: v for k, v in o.items() if not k.startswith("saddle_") and k not in ("mg", "fs", "nlevels")},
                  nsd=c["nsd"], lame=c["lame"])
    f = np.linalg.norm(p.F())
    assert abs(f - c["residuals"][0]) <= _sig(f, c["residuals_text"][0]), (f, c["residuals_text"][0])


# ---- C.2: pattern / size pins from -ksp_view goldens ----------------------------------------------------
@pytest.mark.parametrize("m,rows,nnz,alloc", [(2, 402, 52546, 100794), (4, 2312, 381196, 542628), (6, 6934, 1244446, 1585590)])
def test_pattern_counts_3d(m, rows, nnz, alloc):
    p = O.Problem("-mx %d -model 0" % m, nsd=3)
    assert (p.n, p.nnz, p.prealloc) == (rows, nnz, alloc)
    A = p.A()
    # closed forms of SURVEY App. A.5
    assert p.submatrix(0, 0).ia[-1] == 9 * (8 * m + 1) ** 3
    assert p.submatrix(0, 1).ia[-1] == 3 * (5 * m + 1) ** 3
    assert p.submatrix(1, 1).ia[-1] == (3 * m + 1) ** 3
    assert p.mnnz == (3 * m + 1) ** 3
    # columns strictly ascending in every row
    for i in range(0, p.n, 7):
        r = A.ja[A.ia[i]:A.ia[i + 1]]
        assert np.all(np.diff(r) > 0)


def test_pattern_counts_2d():
    m = 5
    p = O.Problem("-mx %d -model 0" % m, nsd=2)
    assert p.submatrix(0, 0).ia[-1] == 4 * (8 * m + 1) ** 2
    assert p.submatrix(0, 1).ia[-1] == 2 * (5 * m + 1) ** 2
    assert p.submatrix(1, 1).ia[-1] == (3 * m + 1) ** 2


def test_noncubic_pattern_is_product_of_directions():
    p = O.Problem("-mx 4 -my 7 -mz 5 -model 1", nsd=3)
    assert p.submatrix(0, 0).ia[-1] == 9 * (8 * 4 + 1) * (8 * 7 + 1) * (8 * 5 + 1)


def test_galerkin_level_sizes_match_ksp_view(kat):
    p, c = _problem(kat, "exSaddle3d_pseudoice_1")
    r = p.pc_setup()
    assert list(r.level_rows[:3]) == [192, 1029, 6591]
    assert list(r.level_nnz[:3]) == [9000, 61731, 1058841]
    assert [6591, 6591, 1] in c["mat_rows"] and [1058841, 1058841] in c["mat_nnz"]
    assert [89373, 89373] in c["mat_nnz"] and p.submatrix(0, 1).ia[-1] == 89373
    assert [1244446, 1585590] in c["mat_nnz"]


# ---- C.3: diagnostics blocks of the Jacobi-GMRES goldens (fixed iteration count => digits must match) ----
@pytest.mark.parametrize("name", ["exSaddle2d_1", "exSaddle3d_1", "exSaddle2d_2", "exSaddle3d_2"])   # _2: the reference on 2 ranks, same output
def test_jacobi_gmres_diagnostics_exact(kat, name):
    p, c = _problem(kat, name)
    x, r = p.solve()
    assert r.its == c["iterations"] and r.reason == -3 and c["reason"] == "DIVERGED_ITS"
    got = p.diagnostics_text(x)
    assert [g.rstrip() for g in got] == [s.rstrip() for s in c["diagnostics"]]
    assert p.banner.rstrip("\n").split("\n") == c["banner"]


@pytest.mark.parametrize("name", ["exSaddle3d_lame_1", "exSaddle2d_lame_1"])
def test_lame_jacobi_gmres_converges_like_golden(kat, name):
    p, c = _problem(kat, name)
    x, r = p.solve()
    # summation order moves long Jacobi-GMRES runs by one iteration (testref/exSaddle2d_lame_1 vs _2: 145 vs 146)
    assert abs(r.its - c["iterations"]) <= 1 and r.reason == 2
    got = _nums(p.diagnostics_text(x)); ref = _nums(c["diagnostics"])
    assert np.allclose(got, ref, rtol=2e-5, atol=1e-7)


@pytest.mark.parametrize("name", ["exSaddle3d_lame_3", "exSaddle3d_lame_4", "exSaddle3d_lame_5"])
def test_right_jacobi_gmres_history(kat, name):
    p, c = _problem(kat, name)
    x, r = p.solve()
    assert r.nhist == len(c["residuals"])
    for i, t in enumerate(c["residuals_text"]):
        assert abs(r.hist[i] - float(t)) <= _sig(r.hist[i], t), (i, r.hist[i], t)
    assert [g.rstrip() for g in p.diagnostics_text(x)] == [s.rstrip() for s in c["diagnostics"]]


# ---- C.3/C.4: the ABF solver (north-star tree) --------------------------------------------------------
def test_abf_pseudoice_history_with_golden_chebyshev_bounds(kat):
    """exSaddle3d_pseudoice_1: all 21 residuals to the printed digits once the golden's own Chebyshev
    bounds (testref/exSaddle3d_pseudoice_1.ref:103,132) replace the irreproducible noisy estimate (App. B.4)."""
    c = kat["exSaddle3d_pseudoice_1"]
    (e1lo, e1hi), (e2lo, e2hi) = c["cheb_bounds"][0], c["cheb_bounds"][1]
    p, c = _problem(kat, "exSaddle3d_pseudoice_1",
                    "-saddle_fieldsplit_u_mg_levels_1_ksp_chebyshev_eigenvalues %r,%r "
                    "-saddle_fieldsplit_u_mg_levels_2_ksp_chebyshev_eigenvalues %r,%r" % (e1lo, e1hi, e2lo, e2hi))
    x, r = p.solve()
    assert r.its == 20 and r.reason == 2 and r.nhist == 21
    for i, t in enumerate(c["residuals_text"]):
        # bounds are given to 6 digits, so allow 2 units in the 6th digit
        assert abs(r.hist[i] - float(t)) <= 2 * _sig(r.hist[i], t), (i, r.hist[i], t)


def test_abf_ar_3d_iteration_counts(kat):
    p, c = _problem(kat, "exSaddle3d_ar_1")
    x, r = p.solve()
    assert r.its == len(c["residuals"]) - 1 == 6
    assert list(r.inner_its[:r.n_inner]) == c["inner_its"] == [7, 5, 5, 7, 6, 6]
    # own noise vector: histories agree to 1e-3..3e-2 relative (App. C.4)
    for i, t in enumerate(c["residuals"]):
        assert abs(r.hist[i] - t) <= 3e-2 * t


def test_abf_ar_2d_iteration_count_within_one(kat):
    p, c = _problem(kat, "exSaddle2d_ar_1")
    x, r = p.solve()
    assert abs(r.its - (len(c["residuals"]) - 1)) <= 1
    for i in range(5):
        assert abs(r.hist[i] - c["residuals"][i]) <= 1e-2 * c["residuals"][i]


def test_chebyshev_ritz_estimates_close_to_golden(kat):
    """The Ritz extremes depend on PETSc's unknown noise vector; ours land within 1% for emax."""
    p, c = _problem(kat, "exSaddle3d_pseudoice_1")
    r = p.pc_setup()
    assert abs(r.cheb_emax_est[1] - c["cheb_ritz"][0][1]) / c["cheb_ritz"][0][1] < 1e-2
    assert abs(r.cheb_emax_est[2] - c["cheb_ritz"][1][1]) / c["cheb_ritz"][1][1] < 1e-2
    assert abs(r.cheb_emax[1] - 1.1 * r.cheb_emax_est[1]) < 1e-14 and abs(r.cheb_emin[1] - 0.2 * r.cheb_emax_est[1]) < 1e-14


# ---- direct-solve KATs: diagnostics of converged goldens at solver tolerance -----------------------------
def test_direct_solve_matches_fs_golden(kat):
    import scipy.sparse.linalg as spla
    c = kat["exSaddle3d_fs_1"]
    p = O.Problem("-model 2 -sinker_n 1 -mx 4", nsd=3)
    x = spla.splu(p.A().scipy().tocsc()).solve(p.F())
    got = _nums(p.diagnostics_text(x)); ref = _nums(c["diagnostics"])
    assert np.allclose(got[-5:], ref[-5:], rtol=1e-5, atol=1e-6)  # pressure norms (golden stops at rtol 1e-5)
    assert np.allclose(got[:-5], ref[:-5], rtol=5e-2, atol=1e-6)  # velocity ~1e-5 of the pressure scale


def test_mms_error_norms(kat):
    """exSaddle2d_mms_1: direct LU, model 101, constant-pressure null space (testref/exSaddle2d_mms_1.ref)."""
    import scipy.sparse.linalg as spla
    p = O.Problem("-model 101 -mx 16", nsd=2)
    # pin one pressure dof through a bordered system to fix the null space, then project as MatNullSpaceRemove does
    A = p.A().scipy().tolil(); F = p.F().copy()
    k = p.nu
    A[k, :] = 0; A[k, k] = 1.0; F[k] = 0.0
    x = spla.splu(A.tocsc()).solve(F)
    x[p.nu:] -= x[p.nu:].mean()      # MatNullSpaceRemove with the constant-pressure vector (exSaddle.c:288-301)
    # left-preconditioned GMRES with an exact LU: the iteration-0 "residual" is ||A^-1 F|| = ||x||
    assert abs(np.linalg.norm(x) - 255.134) < 2e-3
    nx = 2 * 16 + 1
    xs = np.arange(nx) / (nx - 1.0)
    X, Y = np.meshgrid(xs, xs, indexing="xy")
    uref = np.stack([20 * X * Y ** 3, 5 * (X ** 4 - Y ** 4)], axis=-1).reshape(-1)
    eu = np.linalg.norm(uref - x[:p.nu])
    assert abs(eu - 0.000198842) / 0.000198842 < 2e-2
    assert abs(eu / np.linalg.norm(uref) - 1.20852e-06) / 1.20852e-06 < 2e-2


# ---- unit checks of oracle helpers against numpy ---------------------------------------------------
def test_hessenberg_eigenvalues_vs_numpy():
    rng = np.random.default_rng(0)
    for n in (1, 2, 3, 6, 10):
        H = np.triu(rng.standard_normal((n, n)), -1)
        e = np.sort_complex(O.hess_eig(H)); e0 = np.sort_complex(np.linalg.eigvals(H))
        assert np.allclose(e, e0, rtol=1e-10, atol=1e-12)


def test_rander48_stream_is_drand48():
    v = O.rander48(4)
    # first drand48() values after srand48(0x12345678): X1 = (a*X0+c) mod 2^48
    X = ((0x12345678 << 16) | 0x330E)
    ref = []
    for _ in range(4):
        X = (0x5DEECE66D * X + 0xB) & ((1 << 48) - 1); ref.append(X / float(1 << 48))
    assert np.array_equal(v, np.array(ref))


def test_ilu0_exact_on_tridiagonal_and_matches_dense_pattern_restricted():
    import scipy.sparse as sp
    p = O.Problem("-mx 3 -model 0", nsd=3)
    M = p.Mp(); n = M.shape[0]
    lu = np.empty_like(M.a)
    assert O.lib().xo_ilu0(n, O._ip(M.ia), O._ip(M.ja), O._dp(M.a), O._dp(lu)) == 0
    # reference ILU(0): dense IKJ restricted to the pattern
    A = M.scipy().toarray(); pat = A != 0
    for i in range(n):
        for k in range(i):
            if pat[i, k]:
                A[i, k] = A[i, k] * (1.0 / A[k, k])
                for j in range(k + 1, n):
                    if pat[i, j] and pat[k, j]:
                        A[i, j] -= A[i, k] * A[k, j]
    LU = sp.csr_matrix((lu, M.ja, M.ia), shape=M.shape).toarray()
    d = np.diag(LU).copy(); np.fill_diagonal(LU, 1.0 / d)
    assert np.allclose(LU, A * pat, rtol=1e-12, atol=1e-14)
    b = np.arange(n) * 0.1 - 1.0; x = np.empty(n)
    O.lib().xo_ilu0_solve(n, O._ip(M.ia), O._ip(M.ja), O._dp(lu), O._dp(b), O._dp(x))
    L = np.tril(A * pat, -1) + np.eye(n); U = np.triu(A * pat)
    assert np.allclose(x, np.linalg.solve(U, np.linalg.solve(L, b)), rtol=1e-11)


def test_galerkin_and_transfers_vs_scipy(kat):
    import scipy.sparse as sp
    p = O.Problem("-mx 4 -my 2 -mz 2 -model 6 -eta1 100", nsd=3)
    s = O.Solver(); O.lib().xo_solver_abf(s); s.mg_levels = 2
    p.pc_setup(s)
    A1 = p.mg_level(1).scipy(); A0 = p.mg_level(0).scipy()
    nf = (9, 5, 5); nc = (5, 3, 3)

    def P1(nf_, nc_):
        P = np.zeros((nf_, nc_))
        for f in range(nf_):
            if f % 2 == 0:
                P[f, f // 2] = 1.0
            else:
                P[f, f // 2] = 0.5; P[f, f // 2 + 1] = 0.5
        return sp.csr_matrix(P)
    Pn = sp.kron(P1(nf[2], nc[2]), sp.kron(P1(nf[1], nc[1]), P1(nf[0], nc[0])))
    P = sp.kron(Pn, sp.identity(3)).tocsr()
    G = (P.T @ A1 @ P).toarray()
    assert np.allclose(A0.toarray(), G, rtol=1e-12, atol=1e-13)
    assert A0.nnz == 9 * (3 * 5 - 2) * (3 * 3 - 2) * (3 * 3 - 2)
    rng = np.random.default_rng(1)
    xc = rng.standard_normal(P.shape[1]); xf = rng.standard_normal(P.shape[0])
    assert np.allclose(p.prolong_add(0, xc, xf.copy()), xf + P @ xc, rtol=1e-13)
    assert np.allclose(p.restrict(0, xf, P.shape[1]), P.T @ xf, rtol=1e-13)


# ---- monolithic -mg path: rediscretised levels, GMRES/Jacobi smoothers, LU coarse solve ----
def _short(v):
    return "%g" % v if v > 1e-9 else ("%5.3e" % v if v > 1e-11 else "< 1.e-11")   # KSPMonitorDefaultShort


@pytest.mark.parametrize("name", ["exSaddle3d_mg_1", "exSaddle2d_mg_1", "exSaddle2d_lame_mg_1", "exSaddle3d_lame_mg_1",
                                  "exSaddle2d_lame_mg_2", "exSaddle3d_lame_mg_2"])   # _2: 2-rank runs of the reference, same output
def test_monolithic_mg_history_and_diagnostics_match_golden(kat, name):
    from oracle.oracle_mg import MonolithicMG
    c = kat[name]
    M = MonolithicMG(c["options"], nsd=c["nsd"], lame=c["lame"])
    x, its, reason, hist = M.solve()
    assert reason == 2 and its == len(c["residuals"]) - 1
    assert [_short(v) for v in hist] == c["residuals_text"]          # every printed digit of every iteration
    got = M.fine.diagnostics_text(x)
    assert [g.rstrip() for g in got] == [s.rstrip() for s in c["diagnostics"]]
    assert M.fine.banner.rstrip("\n").split("\n") == c["banner"]


def test_monolithic_mg_two_rank_goldens_equal_the_one_rank_ones(kat):
    # the reference's -n 2 runs print the same histories: nothing in the path depends on the partition
    for a, b in (("exSaddle3d_mg_1", "exSaddle3d_mg_2"), ("exSaddle2d_mg_1", "exSaddle2d_mg_2")):
        assert kat[a]["residuals_text"] == kat[b]["residuals_text"]


def test_monolithic_mg_option_errors():
    from oracle.oracle_mg import MonolithicMG
    with pytest.raises(ValueError):
        MonolithicMG("-mx 6 -mg -nlevels 3 -saddle_ksp_type fgmres -saddle_mg_levels_ksp_type gmres -saddle_mg_levels_pc_type jacobi", nsd=3)   # 6 % 4 != 0 (exSaddle.c:220)
    with pytest.raises(ValueError):
        MonolithicMG("-mx 8 -mg -nlevels 1 -saddle_ksp_type fgmres -saddle_mg_levels_ksp_type gmres -saddle_mg_levels_pc_type jacobi", nsd=3)   # exSaddle.c:209


# ---- bjacobi blocks = ranks of the slab partition (SURVEY 8e caveat 2): the oracle predicts the multi-GPU iteration counts ----
def test_bjacobi_block_oracle_matches_recorded_multi_gpu_runs():
    """profiles/r01_slab_check_n{2,4}.json are the product's 2- and 4-GPU runs of the 8^3 ABF case with bjacobi/ILU(0) per rank
    (scripts/slab_check.py on B200s).  The oracle with the same block structure gives the same outer and inner counts."""
    import json
    abf = " ".join(l for l in O.ABF_OPTS.split("\n") if l.strip())
    base = abf + " -saddle_fieldsplit_p_pc_type bjacobi -model 6 -mx 8 -eta1 100 -saddle_ksp_rtol 1e-8"
    x1, r1 = O.Problem(base, nsd=3).solve()
    xb, rb = O.Problem(base + " -xo_p_blocks 1", nsd=3).solve()
    h1, hb = np.array(r1.hist[:r1.nhist]), np.array(rb.hist[:rb.nhist])
    assert r1.its == rb.its and np.max(np.abs(h1 - hb)) <= 1e-10 * h1[0]   # one block = plain ILU(0) (OpenMP reductions: not bitwise)
    for n in (2, 4):
        rec = json.load(open(os.path.join(ROOT, "profiles", "r01_slab_check_n%d.json" % n)))["abf_bjacobi_ilu"]
        x, r = O.Problem(base + " -xo_p_blocks %d" % n, nsd=3).solve()
        assert r.reason == 2 and r.its == rec["its"][0]
        assert [int(v) for v in r.inner_its[:len(rec["inner"])]] == rec["inner"]
        assert r.its != r1.its                                                              # the block structure matters


def test_recorded_multi_gpu_iteration_counts_at_64cubed_match_block_oracle():
    """BASELINE 64^3 workload on 2 and 8 B200s (profiles/r01_bench_64cubed_n2_v2.json, r01_bench_64cubed_n8.json) against the
    oracle with 2 / 8 bjacobi blocks (tests/golden/oracle_64cubed_bjacobi_blocks.json, made by make_oracle_bjacobi_blocks.py):
    45 and 51 outer iterations on both sides; one rank / one block: 42 (oracle_64cubed_history.json)."""
    import json
    fx = json.load(open(os.path.join(ROOT, "tests", "golden", "oracle_64cubed_bjacobi_blocks.json")))
    for n, prof in ((2, "r01_bench_64cubed_n2_v2.json"), (8, "r01_bench_64cubed_n8.json")):
        rec = json.load(open(os.path.join(ROOT, "profiles", prof)))
        assert rec["n_gpus"] == n and rec["solve"]["reason"] == 2 == fx[str(n)]["reason"]
        assert rec["solve"]["outer_its"] == fx[str(n)]["its"]
        assert abs(rec["solve"]["inner_gcr_its"] - sum(fx[str(n)]["inner"])) <= 1


def test_reference_itself_moves_by_one_iteration_between_rank_counts(kat):
    """The +-1 in north_star's "same iteration count +-1" is the reference's own behaviour: its 1- and 2-rank goldens of the
    same command differ by one iteration (reduction order), with diagnostics equal to ~1e-5."""
    a, b = kat["exSaddle2d_lame_1"], kat["exSaddle2d_lame_2"]
    assert a["options"] == b["options"] and (a["iterations"], b["iterations"]) == (145, 146)
    da, db = _nums(a["diagnostics"]), _nums(b["diagnostics"])
    assert np.allclose(da, db, rtol=1e-3, atol=1e-7) and not np.array_equal(da, db)


def test_wavefront_window_of_the_pressure_ilu_sweeps():
    """Claim behind the product's ILU(0) kernels (xsb_ilu.cu): on the 27-point pressure lattice row (i,j,k) only depends on rows of
    the 7 previous wavefronts w = i + 2j + 4k, so (a) rows of one wavefront are independent and (b) an 8-slot ring of x suffices.
    Emulated here on the oracle's factors: wavefront-ordered sweeps through an 8-slot ring reproduce the sequential solve exactly."""
    o = O.Problem("-model 6 -mx 6 -my 5 -mz 4 -eta1 100", nsd=3)
    M = o.Mp(); n = o.np_
    lu = np.empty_like(M.a)
    assert O.lib().xo_ilu0(n, O._ip(M.ia), O._ip(M.ja), O._dp(M.a), O._dp(lu)) == 0
    px, py = 7, 6
    node = np.arange(n); i, j, k = node % px, (node // px) % py, node // (px * py)
    w = i + 2 * j + 4 * k
    nw = int(w.max()) + 1
    order = np.argsort(w, kind="stable"); off = np.concatenate([[0], np.cumsum(np.bincount(w, minlength=nw))])
    pos = np.empty(n, int); pos[order] = np.arange(n) - off[w[order]]
    maxw = int(np.max(np.diff(off)))
    for r in range(n):      # dependency window
        cols = M.ja[M.ia[r]:M.ia[r + 1]]
        lower, upper = cols[cols < r], cols[cols > r]
        assert np.all((w[r] - w[lower] >= 1) & (w[r] - w[lower] <= 7)) and np.all((w[upper] - w[r] >= 1) & (w[upper] - w[r] <= 7))
    rng = np.random.default_rng(0); b = rng.standard_normal(n)
    xs = np.empty(n); O.lib().xo_ilu0_solve(n, O._ip(M.ia), O._ip(M.ja), O._dp(lu), O._dp(b), O._dp(xs))
    ring = np.full(8 * maxw, np.nan); x = np.empty(n)
    for lvl in range(nw):                                    # forward, unit lower
        for r in order[off[lvl]:off[lvl + 1]]:
            s = b[r]
            for q in range(M.ia[r], M.ia[r + 1]):
                c = M.ja[q]
                if c < r:
                    s -= lu[q] * ring[(w[c] & 7) * maxw + pos[c]]
            ring[(lvl & 7) * maxw + pos[r]] = s; x[r] = s
    for lvl in range(nw - 1, -1, -1):                        # backward, columns descending like MatSolve
        for r in order[off[lvl]:off[lvl + 1]]:
            s = x[r]; d = None
            for q in range(M.ia[r + 1] - 1, M.ia[r] - 1, -1):
                c = M.ja[q]
                if c > r:
                    s -= lu[q] * ring[(w[c] & 7) * maxw + pos[c]]
                elif c == r:
                    d = lu[q]
            s = s * d                                        # the factor stores 1 / pivot (xo_ilu0)
            ring[(lvl & 7) * maxw + pos[r]] = s; x[r] = s
    assert not np.isnan(x).any() and np.linalg.norm(x - xs) <= 1e-13 * np.linalg.norm(xs)


@pytest.mark.parametrize("mesh", [(6, 5, 4), (3, 7, 2), (5, 2, 6)])
def test_line_pipelined_schedule_of_the_pressure_ilu_sweeps(mesh):
    """Claim behind the default ILU(0) solve kernel (k_ilup_solve, xsb_ilu.cu), emulated on the oracle's factors with the kernel's data
    flow: one "CTA" per node plane, one "thread" per node line; at in-plane step t thread j handles node i = t - 2j using ONLY
      * its own previous value (a register),
      * the values line j-1 produced at steps t-3, t-2, t-1, read from a 4-slot ring indexed by step (positions i-1, i, i+1),
      * values of the plane below that were produced at in-plane steps <= t + 3 -- so a plane may run as soon as the plane below is four
        steps ahead (the kernel additionally fetches them four steps early), with the sentinel marking what has not been written;
    the factors come from one 13-entry record per row in sweep order, zero where the lattice ends; the backward sweep is the forward
    sweep on the mirrored lattice.  The emulation advances all planes in lock-step with the minimal lag and must meet no sentinel; the
    result equals the sequential MatSolve."""
    mx, my, mz = mesh
    o = O.Problem("-model 6 -mx %d -my %d -mz %d -eta1 100" % mesh, nsd=3)
    M = o.Mp(); n = o.np_
    lu = np.empty_like(M.a)
    assert O.lib().xo_ilu0(n, O._ip(M.ia), O._ip(M.ja), O._dp(M.a), O._dp(lu)) == 0
    px, py, pz = mx + 1, my + 1, mz + 1
    A = sp.csr_matrix((lu, M.ja, M.ia), shape=(n, n)).tocsr()
    rng = np.random.default_rng(1); b = rng.standard_normal(n)
    xs = np.empty(n); O.lib().xo_ilu0_solve(n, O._ip(M.ia), O._ip(M.ja), O._dp(lu), O._dp(b), O._dp(xs))
    offs = [(u % 3 - 1, (u // 3) % 3 - 1, -1) for u in range(9)] + [(-1, -1, 0), (0, -1, 0), (1, -1, 0), (-1, 0, 0)]   # sweep order of the 13 earlier neighbours

    def sweep(rhs, bwd):
        real = (lambda i, j, k: (px - 1 - i, py - 1 - j, pz - 1 - k)) if bwd else (lambda i, j, k: (i, j, k))
        idx = lambda i, j, k: i + px * (j + py * k)
        rec = {}                                   # packed records: mirrored row -> 13 factors (+ 1 / pivot)
        for k in range(pz):
            for j in range(py):
                for i in range(px):
                    r = idx(*real(i, j, k)); L = np.zeros(14)
                    for u, (di, dj, dk) in enumerate(offs):
                        ni, nj, nk = i + di, j + dj, k + dk
                        if 0 <= ni < px and 0 <= nj < py and 0 <= nk < pz:
                            L[u] = A[r, idx(*real(ni, nj, nk))]
                    L[13] = A[r, r]
                    rec[(i, j, k)] = L
        SENT = object()
        out = {}                                   # (i, j, k) mirrored -> value; absent = sentinel
        S = px + 2 * (py - 1); lag = 4
        ring = np.zeros((pz, py, 4)); xprev = np.zeros((pz, py))
        for T in range(S + lag * (pz - 1)):        # global clock; plane k runs its in-plane step t = T - lag k
            for k in range(pz):
                t = T - lag * k
                if not 0 <= t < S:
                    continue
                new = {}
                for j in range(py):
                    i = t - 2 * j
                    if not 0 <= i < px:
                        continue
                    L = rec[(i, j, k)]
                    xv = []
                    for (di, dj, dk) in offs[:9]:  # plane below: must already be there (produced at in-plane step <= t + 3 of plane k-1, i.e. clock <= T - 1)
                        key = (i + di, j + dj, k - 1)
                        if k == 0 or not (0 <= key[0] < px and 0 <= key[1] < py):
                            xv.append(0.0)
                        else:
                            assert key in out, ("sentinel met: the plane below is not far enough ahead", T, k, j, i)
                            assert key[0] + 2 * key[1] <= t + 3
                            xv.append(out[key])
                    for di in (-1, 0, 1):          # line j-1 through the step-indexed ring: slots (t-3, t-2, t-1) & 3
                        ok = j > 0 and 0 <= i + di < px
                        xv.append(ring[k, j - 1, (t - 2 + di) & 3] if ok else 0.0)
                    xv.append(xprev[k, j] if i > 0 else 0.0)
                    acc = rhs[idx(*real(i, j, k))]
                    for u in range(13):
                        acc -= L[u] * xv[u]
                    if bwd:
                        acc *= L[13]
                    new[j] = (i, acc)
                for j, (i, acc) in new.items():    # the block barrier: writes of step t become visible to step t + 1
                    ring[k, j, t & 3] = acc; xprev[k, j] = acc; out[(i, j, k)] = acc
        res = np.empty(n)
        for (i, j, k), v in out.items():
            res[idx(*real(i, j, k))] = v
        assert len(out) == n
        return res

    y = sweep(b, False)
    x = sweep(y, True)
    assert np.linalg.norm(x - xs) <= 1e-13 * np.linalg.norm(xs)


def test_monolithic_mg_fs_coarse_golden(kat):
    """-mg -fs_coarse (exSaddle.c:366-398): outer GMRES (left PCMG, preconditioned norm), GMRES/Jacobi smoothers (2 steps), coarse
    FGMRES preconditioned by PCFIELDSPLIT Schur / UPPER / user Mpscaled_coarse with GMRES + Jacobi on both splits and the nested
    A00 solve inside every Schur-complement product (App. B.2; solver tree as printed by -saddle_ksp_view in the golden).
    13 iterations and CONVERGED_RTOL as in testref/exSaddle3d_mg_fs_coarse_1.ref; residuals to 3e-5 (the nested inexact solves
    at rtol 1e-5 amplify summation-order differences; the first six agree to all printed digits).  The product runs the same tree on
    the device: tests/test_gpu_parity.py::test_golden_monolithic_mg_fs_coarse."""
    from oracle.oracle_mg import MonolithicMG
    c = kat["exSaddle3d_mg_fs_coarse_1"]
    M = MonolithicMG(c["options"], nsd=3)
    x, its, reason, hist = M.solve()
    assert (its, reason) == (c["iterations"], 2) and c["reason"] == "CONVERGED_RTOL" and len(hist) == len(c["residuals"])
    assert [_short(v) for v in hist[:6]] == c["residuals_text"][:6]
    assert np.allclose(hist, c["residuals"], rtol=3e-5, atol=0)
    assert all(2 <= k <= 3 for k in M.coarse_its)       # the coarse FGMRES needs 2-3 iterations per V-cycle
    assert M.fine.banner.rstrip("\n").split("\n") == c["banner"]


# ---- ASM on the reference's element patches (oracle/oracle_asm.py) ----
@pytest.mark.parametrize("name", ["exSaddle2d_asm_1", "exSaddle3d_asm_1"])
def test_asm_element_patch_goldens(kat, name):
    """`-saddle_pc_type asm -saddle_pc_asm_dm_subdomains -set_ksp_dm` on 9 / 8 ranks (Makefile:298, 411): one closed element patch per
    rank (+ -dmdafe_overlap), exact sub-solves, sub-solutions kept on the rank's owned dofs (PC_ASM_RESTRICT): every printed digit
    of every residual of testref/exSaddle{2d,3d}_asm_1.ref (one last-digit rounding difference allowed)."""
    from oracle import oracle_asm as OA
    c = kat[name]
    x, its, reason, hist, p = OA.solve(c["options"], c["nsd"], c["nranks"])
    assert reason == 2 and its == len(c["residuals"]) - 1
    # every printed digit, up to one unit of the sixth digit (SuperLU here, UMFPACK there: 4.97997e-05 against 4.97998e-05 at iteration 23 in 3-D)
    assert np.allclose(hist, c["residuals"], rtol=6e-6, atol=0)   # six printed digits
    assert sum(a != b for a, b in zip([_short(v) for v in hist], c["residuals_text"])) <= 1
    assert p.banner.rstrip("\n").split("\n") == c["banner"]


def test_asm_smoother_inside_monolithic_mg_golden(kat):
    """`-mg -nlevels 2 -saddle_mg_levels_pc_type asm -saddle_mg_levels_pc_asm_dm_subdomains -dmdafe_overlap 1` on 4 ranks, 6 x 4 x 4
    elements (Makefile:418): the same patches as the PC of the GMRES smoother; testref/exSaddle3d_mg_asm_1.ref digit for digit."""
    from oracle.oracle_mg import MonolithicMG
    c = kat["exSaddle3d_mg_asm_1"]
    M = MonolithicMG(c["options"], nsd=3, nranks=c["nranks"])
    x, its, reason, hist = M.solve()
    assert reason == 2 and [_short(v) for v in hist] == c["residuals_text"]


def test_asm_subdomains_host_logic_matches_oracle_and_tiles_the_mesh():
    """xsb_dmda_grid / xsb_asm_subdomain (host integer logic of the product, no GPU) against the oracle's restatement of PETSc's
    DMDA partition + the reference's element fitting: process grids, patches, owned ranges; owned ranges tile both lattices."""
    import exsaddle_b200 as X
    from oracle import oracle_asm as OA
    for nsd, mesh, size, ov in [(2, (12, 12, 1), 9, 1), (3, (6, 6, 6), 8, 0), (3, (6, 4, 4), 4, 1), (2, (8, 5, 1), 6, 0), (3, (8, 8, 8), 12, 2),
                                (3, (4, 6, 10), 5, 1), (2, (16, 4, 1), 4, 1), (3, (64, 64, 64), 64, 1), (2, (7, 9, 1), 1, 3)]:
        grid, sds = OA.subdomains(nsd, mesh, size, ov)
        N = [2 * m + 1 for m in mesh]
        assert X.dmda_grid(nsd, N[0], N[1], N[2] if nsd == 3 else 1, size)[:nsd] == tuple(grid[:nsd])
        cover_u = np.zeros(N[:nsd][::-1], int); cover_p = np.zeros([m + 1 for m in mesh[:nsd]][::-1], int)
        for r, sd in enumerate(sds):
            got = X.asm_subdomain(nsd, mesh[0], mesh[1], mesh[2], size, ov, r)
            assert got == sd, (nsd, mesh, size, r)
            su = tuple(slice(a, b) for a, b in got["own_u"][::-1]); spp = tuple(slice(a, b) for a, b in got["own_p"][::-1])
            cover_u[su] += 1; cover_p[spp] += 1
            for d in range(nsd):   # owned nodes lie inside the rank's patch
                assert 2 * got["lo"][d] <= got["own_u"][d][0] and got["own_u"][d][1] <= 2 * got["hi"][d] + 1
        assert np.all(cover_u == 1) and np.all(cover_p == 1)
    with pytest.raises(X.XsbError):     # 3 ranks on 5 nodes per direction: no rank can hold a whole element column (femixedspace.c:1097)
        X.asm_subdomain(2, 2, 2, 1, 9, 0, 0)


# ---- the reference's plain -fs tree with PETSc's default sub-solvers (oracle only; oracle/oracle_fs.py) ----
@pytest.mark.parametrize("name", ["exSaddle3d_fs_1", "exSaddle2d_fs_1", "exSaddle2d_lame_fs_1", "exSaddle3d_lame_fs_1",
                                  "exSaddle3d_fs_2", "exSaddle2d_fs_2", "exSaddle2d_lame_fs_2", "exSaddle3d_lame_fs_2"])
def test_plain_fs_tree_history_and_diagnostics_match_golden(kat, name):
    """GMRES + PCFIELDSPLIT Schur / UPPER / user Mpscaled with GMRES + ILU(0) on both splits and the nested A00 solve inside every
    Schur-complement product (App. B.2): residual history and diagnostics to every printed digit of testref/*_fs_1.ref; and of
    testref/*_fs_2.ref, the reference on 2 ranks, where the inner PCs are bjacobi with one ILU(0) block per rank on the dofs the
    rank's DMDAs own (the same ownership restatement the ASM goldens pin)."""
    from oracle.oracle_fs import FieldSplitDefault
    c = kat[name]
    F = FieldSplitDefault(c["options"], nsd=c["nsd"], lame=c["lame"], nranks=c["nranks"])
    x, its, reason, hist = F.solve()
    assert reason == 2 and its == len(c["residuals"]) - 1
    assert [_short(v) for v in hist] == c["residuals_text"]
    assert [g.rstrip() for g in F.p.diagnostics_text(x)] == [s.rstrip() for s in c["diagnostics"]]
    assert F.p.banner.rstrip("\n").split("\n") == c["banner"]


def test_history_floor_at_baseline_sizes():
    """The committed oracle-vs-oracle drift records (scripts/oracle_drift.py: the oracle with reversed dot-product / row sums
    against the oracle fixture): same outer and inner iteration counts, histories apart by ~1e-6 -- two orders of magnitude
    above north_star's 1e-8, so that bar is not attainable at eta1/eta0 = 1e6 by any implementation that does not reproduce
    the reference's summation order bit for bit.  The GPU-vs-oracle bound of tests/test_gpu_parity.py sits at this floor."""
    for mx in (32, 64):
        d = json.load(open(os.path.join(ROOT, "profiles", "r02_oracle_drift_%dcubed.json" % mx)))
        assert d["its"][0] == d["its"][1] and d["inner_its_equal"]
        assert 1e-8 < d["max_rel_hist_diff_above_1e-2"] < 1e-5
