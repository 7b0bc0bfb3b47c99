"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle on the same inputs,
against the reference's golden outputs (tests/golden/testref_kat.json), and -- at sizes where the oracle is
slow -- through size-independent properties.  Bars (BASELINE.json north_star): sparsity pattern and index maps
bit-exact; operator apply <= 1e-12 relative; residual histories <= 1e-8 relative with equal iteration counts."""
import numpy as np
import pytest

import exsaddle_b200 as X
from oracle import oracle as O

pytestmark = pytest.mark.gpu

ABF = " ".join(l for l in O.ABF_OPTS.split("\n") if l.strip())

CASES = [  # (id, nsd, lame, options)
    ("stokes2d_solcx", 2, False, "-model 0 -mx 4"),
    ("stokes2d_noncubic", 2, False, "-model 6 -mx 5 -my 3 -eta1 100"),
    ("stokes3d_sinkers_noncubic", 3, False, "-model 1 -mx 4 -my 7 -mz 5 -eta1 10"),
    ("stokes3d_sinker6", 3, False, "-model 6 -mx 6 -eta0 1 -eta1 1e6"),
    ("stokes3d_pseudoice", 3, False, "-model 11 -size_x 0.1 -mx 6"),
    ("stokes3d_solcx_aspect", 3, False, "-model 0 -mx 6 -size_z 0.1 -eta1 1e3"),
    ("lame3d_inclusion", 3, True, "-model 6 -mx 4"),
    ("lame3d_compression", 3, True, "-model 9 -mx 4"),
    ("lame3d_compression2", 3, True, "-model 12 -mx 4 -mu1 10 -lambda1 100"),
    ("lame2d_xsinker", 2, True, "-model 2 -mx 8 -mu1 100 -lambda1 10"),
    ("stokes2d_mms", 2, False, "-model 101 -mx 6"),
    ("stokes3d_tiny", 3, False, "-model 2 -mx 1"),
]


def _relerr(a, b):
    s = max(np.max(np.abs(b)), 1e-300)
    return np.max(np.abs(a - b)) / s


@pytest.fixture(scope="module", params=CASES, ids=[c[0] for c in CASES])
def pair(request):
    _, nsd, lame, opts = request.param
    g = X.ExSaddle(opts, nsd=nsd, lame=lame).assemble()
    o = O.Problem(opts, nsd=nsd, lame=lame)
    yield g, o
    g.close()


# ------------------------------------------------------------------ set-up parity
def test_sizes_and_pattern_bit_exact(pair):
    g, o = pair
    assert (g.n, g.nu, g.np_, g.nnz, g.prealloc, g.nel, g.nbc, g.mnnz) == (o.n, o.nu, o.np_, o.nnz, o.prealloc, o.nel, o.nbc, o.mnnz)
    ia, ja, a, _ = g.mat_csr(X.MAT_A)
    A = o.A()
    assert ia.dtype == np.int32 and np.array_equal(ia, A.ia) and np.array_equal(ja, A.ja)
    mia, mja, ma, _ = g.mat_csr(X.MAT_MP)
    M = o.Mp()
    assert np.array_equal(mia, M.ia) and np.array_equal(mja, M.ja)
    bi, bv = g.bc(); oi, ov = o.bc()
    assert np.array_equal(bi, oi) and np.allclose(bv, ov, rtol=1e-15, atol=0)


def test_assembled_values_rhs_and_coefficients(pair):
    g, o = pair
    _, _, a, _ = g.mat_csr(X.MAT_A)
    assert _relerr(a, o.A().a) <= 1e-13
    _, _, ma, _ = g.mat_csr(X.MAT_MP)
    assert _relerr(ma, o.Mp().a) <= 1e-13
    assert _relerr(g.rhs(), o.F()) <= 1e-13
    cq = o.coeff_qp()
    for gs, os_ in ((0, 0), (1, 1), (2, 2), (3, 3), (4, 4), (5, 5)):   # slot order is shared: eta,Fu0,Fu1,Fu2,Fp,lambda
        assert _relerr(g.coeff_qp(gs), cq[:, :, os_].reshape(-1)) <= 1e-14


def test_subblocks_are_index_set_extractions(pair):
    g, o = pair
    for which, rb, cb in ((X.MAT_A00, 0, 0), (X.MAT_A01, 0, 1), (X.MAT_A10, 1, 0), (X.MAT_A11, 1, 1)):
        ia, ja, a, shape = g.mat_csr(which)
        S = o.submatrix(rb, cb)
        assert shape == S.shape and np.array_equal(ia, S.ia) and np.array_equal(ja, S.ja)
        assert _relerr(a, S.a) <= 1e-13 if len(a) and np.max(np.abs(S.a)) > 0 else np.all(a == 0)


def test_matmult_within_1e12(pair):
    g, o = pair
    rng = np.random.default_rng(7)
    for x in (np.sin(0.37 * np.arange(o.n)) + 0.1, rng.standard_normal(o.n)):
        y = g.mat_mult(X.MAT_A, x); yo = o.mult(x)
        assert np.linalg.norm(y - yo) <= 1e-12 * np.linalg.norm(yo)
        # blocks reproduce the full operator: A x = [A00 xu + A01 xp ; A10 xu + A11 xp]
        xu, xp = x[:o.nu], x[o.nu:]
        yb = np.concatenate([g.mat_mult(X.MAT_A00, xu) + g.mat_mult(X.MAT_A01, xp), g.mat_mult(X.MAT_A10, xu) + g.mat_mult(X.MAT_A11, xp)])
        assert np.linalg.norm(yb - yo) <= 1e-12 * np.linalg.norm(yo)
    assert np.array_equal(g.mat_diagonal(X.MAT_A), o.A().scipy().diagonal()) or _relerr(g.mat_diagonal(X.MAT_A), o.A().scipy().diagonal()) <= 1e-13


def test_operator_is_symmetric_and_linear(pair):
    g, o = pair
    rng = np.random.default_rng(3)
    x, y = rng.standard_normal(o.n), rng.standard_normal(o.n)
    Ax, Ay = g.mat_mult(X.MAT_A, x), g.mat_mult(X.MAT_A, y)
    assert abs(y @ Ax - x @ Ay) <= 1e-11 * (np.linalg.norm(x) * np.linalg.norm(Ay))
    Az = g.mat_mult(X.MAT_A, 2.0 * x - 3.0 * y)
    assert np.linalg.norm(Az - (2.0 * Ax - 3.0 * Ay)) <= 1e-12 * np.linalg.norm(Az)


# ------------------------------------------------------------------ preconditioner pieces
MG_CASES = [("3d_6_l3", 3, False, "-model 11 -size_x 0.1 -mx 6", 3), ("3d_8_l3_contrast", 3, False, "-model 6 -mx 8 -eta1 1e4", 3),
            ("3d_noncubic_l2", 3, False, "-model 1 -mx 4 -my 6 -mz 2", 2), ("2d_16_l4", 2, False, "-model 0 -mx 16 -my 16 -size_y 0.1", 4),
            ("lame3d_4_l2", 3, True, "-model 6 -mx 4 -lambda1 20", 2)]


@pytest.fixture(scope="module", params=MG_CASES, ids=[c[0] for c in MG_CASES])
def abf_pair(request):
    _, nsd, lame, opts, levels = request.param
    full = "%s %s -saddle_fieldsplit_u_pc_mg_levels %d -saddle_ksp_rtol 1e-8" % (ABF, opts, levels)
    g = X.ExSaddle(full, nsd=nsd, lame=lame).assemble().ksp_setup()
    o = O.Problem(full, nsd=nsd, lame=lame)
    res = o.pc_setup()
    yield g, o, res, levels
    g.close()


def test_galerkin_levels(abf_pair):
    g, o, res, levels = abf_pair
    for l in range(levels):
        ia, ja, a, shape = g.mat_csr(X.MAT_MG_LEVEL0 + l)
        L = o.mg_level(l)
        assert shape == L.shape and np.array_equal(ia, L.ia) and np.array_equal(ja, L.ja)
        assert _relerr(a, L.a) <= 1e-12
        assert shape[0] == res.level_rows[l] and len(a) == res.level_nnz[l]


def test_chebyshev_estimates_and_transfers(abf_pair):
    g, o, res, levels = abf_pair
    rng = np.random.default_rng(5)
    for l in range(1, levels):
        emin_est, emax_est, emin, emax = g.chebyshev(l)
        assert abs(emax_est - res.cheb_emax_est[l]) <= 1e-9 * emax_est      # same rander48 noise vector by construction
        assert abs(emin - res.cheb_emin[l]) <= 1e-9 * emax and abs(emax - res.cheb_emax[l]) <= 1e-9 * emax
    for lc in range(levels - 1):
        nf = g.mat_info(X.MAT_MG_LEVEL0 + lc + 1)[0]; nc = g.mat_info(X.MAT_MG_LEVEL0 + lc)[0]
        rf = rng.standard_normal(nf); xc = rng.standard_normal(nc); xf = rng.standard_normal(nf)
        assert _relerr(g.mg_restrict(lc, rf), o.restrict(lc, rf, nc)) <= 1e-14
        assert _relerr(g.mg_interpolate_add(lc, xc, xf), o.prolong_add(lc, xc, xf.copy())) <= 1e-14


def test_vcycle_ilu_and_fieldsplit_apply(abf_pair):
    g, o, res, levels = abf_pair
    rng = np.random.default_rng(11)
    b = rng.standard_normal(o.nu)
    bi, _ = o.bc(); b[bi] = 0.0
    xg = g.pc_mg_apply(b); xo = o.vcycle(b)
    assert np.linalg.norm(xg - xo) <= 1e-9 * np.linalg.norm(xo)
    bp = rng.standard_normal(o.np_)
    M = o.Mp(); lu = np.empty_like(M.a)
    assert O.lib().xo_ilu0(o.np_, O._ip(M.ia), O._ip(M.ja), O._dp(M.a), O._dp(lu)) == 0
    xp = np.empty(o.np_); O.lib().xo_ilu0_solve(o.np_, O._ip(M.ia), O._ip(M.ja), O._dp(lu), O._dp(bp), O._dp(xp))
    assert np.linalg.norm(g.pc_schur_apply(bp) - xp) <= 1e-12 * np.linalg.norm(xp)
    r = o.F() + 1e-3 * rng.standard_normal(o.n) * np.linalg.norm(o.F()) / np.sqrt(o.n)
    zg = g.pc_apply(r); zo, its = o.pc_apply(r)
    assert np.linalg.norm(zg - zo) <= 1e-8 * np.linalg.norm(zo)


@pytest.mark.parametrize("nsd,levels,opts", [(3, 1, "-model 6 -mx 5 -my 3 -mz 4 -eta1 1e4"), (3, 1, "-model 1 -mx 2 -my 6 -mz 9 -eta1 10"), (3, 3, "-model 6 -mx 16 -eta1 1e6"),
                                             (3, 1, "-model 2 -mx 1"), (2, 1, "-model 6 -mx 9 -my 4 -eta1 100"), (3, 3, "-model 6 -mx 40 -my 32 -mz 8 -eta1 1e3")])
def test_pressure_ilu0_solve_pipelined_vs_wavefront_vs_oracle(nsd, levels, opts):
    """bjacobi/ILU(0) on Mpscaled (abf.opts:15): the line-pipelined triangular solve (default) against the oracle's sequential
    MatSolve (<= 1e-12) and, bit for bit, against the wavefront kernel it replaces (-xsb_ilu_kernel 0); twice, because the
    step counters the CTAs hand to each other must be back at zero after a launch."""
    abf = ABF + " -saddle_fieldsplit_u_pc_mg_levels %d" % levels
    o = O.Problem(opts, nsd=nsd)
    M = o.Mp(); lu = np.empty_like(M.a)
    assert O.lib().xo_ilu0(o.np_, O._ip(M.ia), O._ip(M.ja), O._dp(M.a), O._dp(lu)) == 0
    rng = np.random.default_rng(3)
    out = {}
    for kern in (1, 0):
        g = X.ExSaddle(abf + " " + opts + " -xsb_ilu_kernel %d" % kern, nsd=nsd).assemble().ksp_setup()
        res = []
        for rep in range(2):
            bp = np.cos(0.3 * np.arange(o.np_) + rep)
            xp = np.empty(o.np_); O.lib().xo_ilu0_solve(o.np_, O._ip(M.ia), O._ip(M.ja), O._dp(lu), O._dp(bp), O._dp(xp))
            xg = g.pc_schur_apply(bp)
            assert np.linalg.norm(xg - xp) <= 1e-12 * np.linalg.norm(xp), (kern, rep)
            res.append(xg)
        out[kern] = res
        g.close()
    for a, b in zip(out[0], out[1]):
        assert np.array_equal(a, b)


def test_abf_solve_history_matches_oracle(abf_pair):
    g, o, res, levels = abf_pair
    x = g.solve()
    xo, r = o.solve()
    its, reason = g.iterations()
    assert (its, reason) == (r.its, r.reason)
    assert g.inner_iterations() == list(r.inner_its[:r.n_inner])
    h = g.history(); ho = np.array(r.hist[:r.nhist])
    assert len(h) == len(ho)
    # "within 1e-8 relative": KSP-relative, i.e. against ||r_0|| like every KSP tolerance -- held 100x tighter here.
    assert np.max(np.abs(h - ho)) <= 1e-10 * ho[0]
    # entry-by-entry 1e-8 relative while the residual is above 1e-2 ||r_0||; further down the two runs differ by
    # (condition number) x (FMA / reduction-order rounding) -- 3e-8 at 1e-4 ||r_0|| for the eta-contrast 1e4
    # pseudo-ice case -- and only the KSP-relative bound above is meaningful
    big = ho >= 1e-2 * ho[0]
    assert np.max(np.abs(h - ho)[big] / ho[big]) <= 1e-8
    assert np.max(np.abs(h - ho) / ho) <= 1e-3
    assert np.linalg.norm(x - xo) <= 1e-7 * np.linalg.norm(xo)
    # the answer really solves the system: true residual at the requested tolerance
    F = o.F()
    assert np.linalg.norm(F - o.mult(x)) <= 1.05e-8 * np.linalg.norm(F) * 1.5


# ------------------------------------------------------------------ the reference's golden outputs, end to end
def _run(kat, name, extra="", nranks=1):
    c = kat[name]
    opts = c["options"].replace("-options_file abf.opts", ABF) + " " + extra
    text, s, x = X.run_exsaddle(c["exe"], opts, nranks=nranks)
    return c, text, s, x


@pytest.mark.parametrize("name", ["exSaddle2d_1", "exSaddle3d_1"])
def test_golden_jacobi_gmres_output_is_identical(kat, name):
    """BASELINE config 0 (exSaddle2d_1) and the non-cubic 3-D case: the program output diffs clean against testref."""
    c, text, s, x = _run(kat, name)
    ref = open_ref_lines(c)
    assert [l.rstrip() for l in text.rstrip("\n").split("\n")] == ref


def open_ref_lines(c):
    lines = list(c["banner"])
    if "reason" in c:
        verb = "converged" if c["reason"].startswith("CONVERGED") else "did not converge"
        lines.append("Linear saddle_ solve %s due to %s iterations %d" % (verb, c["reason"], c["iterations"]))
    lines += [l.rstrip() for l in c["diagnostics"]]
    return lines


@pytest.mark.parametrize("name", ["exSaddle3d_lame_3", "exSaddle3d_lame_4", "exSaddle3d_lame_5"])
def test_golden_right_jacobi_histories(kat, name):
    c, text, s, x = _run(kat, name)
    h = s.history()
    assert len(h) == len(c["residuals"])
    for v, t in zip(h, c["residuals_text"]):
        assert X.monitor_short(v) == t or abs(v - float(t)) <= 6e-6 * v
    got = [l.rstrip() for l in text.rstrip("\n").split("\n") if l.startswith("|")]
    assert got == [l.rstrip() for l in c["diagnostics"]]


def test_golden_abf_pseudoice_history(kat):
    c = kat["exSaddle3d_pseudoice_1"]
    (a1, b1), (a2, b2) = c["cheb_bounds"][0], c["cheb_bounds"][1]
    c, text, s, x = _run(kat, "exSaddle3d_pseudoice_1",
                         "-saddle_fieldsplit_u_mg_levels_1_ksp_chebyshev_eigenvalues %r,%r -saddle_fieldsplit_u_mg_levels_2_ksp_chebyshev_eigenvalues %r,%r" % (a1, b1, a2, b2))
    h = s.history()
    assert s.iterations() == (20, 2) and len(h) == 21
    for v, t in zip(h, c["residuals_text"]):
        assert abs(v - float(t)) <= 1.2e-5 * v, (v, t)
    # -ksp_view pins: level sizes and nnz
    for l, (rows, nnz) in enumerate([(192, 9000), (1029, 61731), (6591, 1058841)]):
        r, _, z, bs = s.mat_info(X.MAT_MG_LEVEL0 + l)
        assert (r, z, bs) == (rows, nnz, 3)


def test_golden_abf_ar_iteration_counts(kat):
    c, text, s, x = _run(kat, "exSaddle3d_ar_1")
    assert s.iterations()[0] == 6 and s.inner_iterations() == c["inner_its"]
    for v, t in zip(s.history(), c["residuals"]):
        assert abs(v - t) <= 3e-2 * t
    assert s.options_left() == []     # "There are no unused options." (testref/exSaddle3d_ar_1.ref)


# ------------------------------------------------------------------ larger sizes: properties instead of the oracle
def test_32cubed_properties():
    g = X.ExSaddle(ABF + " -mx 32 -model 6 -eta0 1 -eta1 100 -saddle_fieldsplit_u_pc_mg_levels 5 -saddle_ksp_rtol 1e-8", nsd=3).assemble()
    m = 32
    assert g.n == 3 * (2 * m + 1) ** 3 + (m + 1) ** 3 == 859812
    assert g.nnz == 9 * (8 * m + 1) ** 3 + 2 * 3 * (5 * m + 1) ** 3 + (3 * m + 1) ** 3 == 178723696
    rng = np.random.default_rng(0)
    x, y = rng.standard_normal(g.n), rng.standard_normal(g.n)
    Ax, Ay = g.mat_mult(X.MAT_A, x), g.mat_mult(X.MAT_A, y)
    assert abs(y @ Ax - x @ Ay) <= 1e-11 * np.linalg.norm(x) * np.linalg.norm(Ay)
    xu, xp = x[:g.nu], x[g.nu:]
    yb = np.concatenate([g.mat_mult(X.MAT_A00, xu) + g.mat_mult(X.MAT_A01, xp), g.mat_mult(X.MAT_A10, xu) + g.mat_mult(X.MAT_A11, xp)])
    assert np.linalg.norm(yb - Ax) <= 1e-12 * np.linalg.norm(Ax)
    # constant pressure is in the kernel of the gradient block away from Dirichlet rows: sum over rows of A01 p=1
    g.ksp_setup()
    sol = g.solve()
    its, reason = g.iterations()
    assert reason == 2 and its < 60
    F = g.rhs()
    assert np.linalg.norm(F - g.mat_mult(X.MAT_A, sol)) <= 1.5e-8 * np.linalg.norm(F)
    h = g.history()
    assert abs(h[0] - np.linalg.norm(F)) <= 1e-12 * h[0] and np.all(h[1:] <= h[:-1] * (1 + 1e-12))   # FGMRES residual is monotone
    g.close()


# ------------------------------------------------------------------ K4: matrix-free Q2 apply
@pytest.mark.parametrize("nsd,lame,opts", [(3, False, "-model 6 -mx 4 -my 3 -mz 2 -eta1 100"), (2, False, "-model 0 -mx 5 -my 3 -size_x 2.0"), (3, True, "-model 6 -mx 3 -mu1 10 -lambda1 5"),
                                           (3, False, "-model 11 -size_x 0.1 -mx 4"), (3, True, "-model 9 -mx 4"), (3, False, "-model 2 -mx 1"), (2, True, "-model 2 -mx 8 -mu1 100 -lambda1 10")])
def test_gradient_and_divergence_blocks_matrix_free(nsd, lame, opts):
    """A01 x_p and A10 x_u by the closed-form separable stencils (csrc/xsb_grad.cu) against the assembled blocks of the library and of
    the oracle: <= 1e-13 of the result's max norm, constrained rows / columns included, random and smooth inputs."""
    g = X.ExSaddle(opts, nsd=nsd, lame=lame).assemble()
    o = O.Problem(opts, nsd=nsd, lame=lame)
    A = o.A().scipy().tocsr(); nu = o.nu
    rng = np.random.default_rng(2)
    for rep in range(2):
        xp = rng.standard_normal(o.np_) if rep else np.cos(0.3 * np.arange(o.np_)) + 0.2
        xu = rng.standard_normal(nu) if rep else np.sin(0.11 * np.arange(nu)) + 0.1
        y01 = g.mat_mult(X.MAT_A01_MF, xp); y10 = g.mat_mult(X.MAT_A10_MF, xu)
        for got, ref, ref2 in ((y01, g.mat_mult(X.MAT_A01, xp), A[:nu, nu:] @ xp), (y10, g.mat_mult(X.MAT_A10, xu), A[nu:, :nu] @ xu)):
            s = np.max(np.abs(ref2))
            assert np.max(np.abs(got - ref)) <= 1e-13 * s and np.max(np.abs(got - ref2)) <= 1e-13 * s
    bi, _ = o.bc()
    assert np.all(g.mat_mult(X.MAT_A01_MF, np.ones(o.np_))[bi[bi < nu]] == 0.0)      # rows of constrained velocity dofs are exactly zero
    g.close()


@pytest.mark.parametrize("opts,lame", [("-model 6 -mx 4 -eta1 1e4", False), ("-model 1 -mx 3 -my 5 -mz 2 -eta1 10", False),
                                       ("-model 11 -size_x 0.1 -mx 6", False), ("-model 0 -mx 5 -size_z 0.3 -freesliphack", False),
                                       ("-model 12 -mx 4 -mu1 10", True), ("-model 2 -mx 1", False),
                                       ("-model 6 -mx 18 -my 7 -mz 3 -eta1 1e3", False), ("-model 1 -mx 35 -my 11 -mz 2", False)])   # several tiles of the one-pass kernel
@pytest.mark.parametrize("kernel", [3, 4])
def test_matrix_free_apply_matches_assembled_block(opts, lame, kernel):
    g = X.ExSaddle(opts + " -xsb_mf_kernel %d" % kernel, nsd=3, lame=lame).assemble()
    o = O.Problem(opts, nsd=3, lame=lame)
    A00 = o.submatrix(0, 0).scipy()
    rng = np.random.default_rng(2)
    for x in (np.sin(0.37 * np.arange(o.nu)) + 0.1, rng.standard_normal(o.nu)):
        y = g.mat_mult(X.MAT_A00_MF, x)
        yo = A00 @ x
        assert np.linalg.norm(y - yo) <= 1e-12 * np.linalg.norm(yo)
        assert np.linalg.norm(y - g.mat_mult(X.MAT_A00, x)) <= 1e-12 * np.linalg.norm(yo)
        assert np.array_equal(y, g.mat_mult(X.MAT_A00_MF, x))      # fixed summation order: bit-reproducible
    g.close()


def test_matrix_free_solve_matches_assembled_solve():
    base = ABF + " -model 6 -mx 8 -eta1 1e4 -saddle_ksp_rtol 1e-8"
    ga = X.ExSaddle(base, nsd=3).assemble().ksp_setup()
    gm = X.ExSaddle(base + " -xsb_matrix_free", nsd=3).assemble().ksp_setup()
    xa, xm = ga.solve(), gm.solve()
    assert ga.iterations() == gm.iterations() and ga.inner_iterations() == gm.inner_iterations()
    ha, hm = ga.history(), gm.history()
    assert np.max(np.abs(ha - hm)) <= 1e-10 * ha[0]
    assert np.linalg.norm(xa - xm) <= 1e-7 * np.linalg.norm(xa)
    ga.close(); gm.close()


# ------------------------------------------------------------------ operator-free mode (-xsb_matrix_free full): A / A00 never stored
FULL_CASES = [("stokes_8_l3_contrast", False, "-model 6 -mx 8 -eta1 1e4", 3), ("stokes_noncubic_l2", False, "-model 1 -mx 4 -my 6 -mz 2", 2),
              ("stokes_pseudoice_l3", False, "-model 11 -size_x 0.1 -mx 6", 3), ("lame_compression2_l2", True, "-model 12 -mx 4 -mu1 10 -lambda1 100", 2)]


@pytest.mark.parametrize("name,lame,opts,levels", FULL_CASES, ids=[c[0] for c in FULL_CASES])
def test_operator_free_mode_matches_oracle_and_assembled_path(name, lame, opts, levels):
    full = "%s %s -saddle_fieldsplit_u_pc_mg_levels %d -saddle_ksp_rtol 1e-8" % (ABF, opts, levels)
    g = X.ExSaddle(full + " -xsb_matrix_free full", nsd=3, lame=lame).assemble().ksp_setup()
    ga = X.ExSaddle(full, nsd=3, lame=lame).assemble().ksp_setup()
    o = O.Problem(full, nsd=3, lame=lame)
    res = o.pc_setup()
    # sizes, RHS (incl. the Dirichlet lifting through the unmasked element kernel), operator apply
    assert (g.n, g.nu, g.np_, g.nnz) == (o.n, o.nu, o.np_, o.nnz)
    assert _relerr(g.rhs(), o.F()) <= 1e-13
    rng = np.random.default_rng(3)
    for x in (np.sin(0.37 * np.arange(o.n)) + 0.1, rng.standard_normal(o.n)):
        yo = o.mult(x)
        assert np.linalg.norm(g.mat_mult(X.MAT_A, x) - yo) <= 1e-12 * np.linalg.norm(yo)
    for which in (X.MAT_A01, X.MAT_A10, X.MAT_A11, X.MAT_MP):
        (ia, ja, a, shape), (ia2, ja2, a2, shape2) = g.mat_csr(which), ga.mat_csr(which)
        assert shape == shape2 and np.array_equal(ia, ia2) and np.array_equal(ja, ja2) and np.array_equal(a, a2)
    with pytest.raises(X.XsbError):
        g.mat_csr(X.MAT_A)
    # element-wise Galerkin product and everything below it; Jacobi diagonal through the Chebyshev bounds
    for l in range(levels - 1):
        ia, ja, a, shape = g.mat_csr(X.MAT_MG_LEVEL0 + l)
        L = o.mg_level(l)
        assert shape == L.shape and np.array_equal(ia, L.ia) and np.array_equal(ja, L.ja)
        assert _relerr(a, L.a) <= 1e-12
    for l in range(1, levels):
        emin_est, emax_est, emin, emax = g.chebyshev(l)
        assert abs(emax_est - res.cheb_emax_est[l]) <= 1e-9 * emax_est
    # the solve: same iteration counts as the oracle and the assembled GPU path, same history
    x = g.solve(); xa = ga.solve(); xo, r = o.solve()
    assert g.iterations() == ga.iterations() == (r.its, r.reason)
    assert g.inner_iterations() == ga.inner_iterations() == list(r.inner_its[:r.n_inner])
    h = g.history(); ho = np.array(r.hist[:r.nhist])
    assert np.max(np.abs(h - ho)) <= 1e-10 * ho[0]
    assert np.linalg.norm(x - xo) <= 1e-7 * np.linalg.norm(xo)
    F = o.F()
    assert np.linalg.norm(F - o.mult(x)) <= 1.6e-8 * np.linalg.norm(F)
    g.close(); ga.close()


# ------------------------------------------------------------------ monolithic -mg path
@pytest.mark.parametrize("name", ["exSaddle3d_mg_1", "exSaddle2d_mg_1", "exSaddle2d_lame_mg_1", "exSaddle3d_lame_mg_1"])
def test_golden_monolithic_mg_output_is_identical(kat, name):
    """`-mg -nlevels L`: rediscretised levels, GMRES/Jacobi smoothers, LU coarse solve.  The program output (banner,
    residual history as printed by -saddle_ksp_monitor_short, diagnostics) diffs clean against testref/*.ref."""
    c, text, s, x = _run(kat, name)
    ref = list(c["banner"]) + ["  Residual norms for saddle_ solve."] + \
        ["%3d KSP Residual norm %s" % (i, t) for i, t in enumerate(c["residuals_text"])] + [l.rstrip() for l in c["diagnostics"]]
    assert [l.rstrip() for l in text.rstrip("\n").split("\n")] == ref
    assert s.iterations() == (len(c["residuals"]) - 1, 2)


# ------------------------------------------------------------------ plain -fs tree: PETSc's default sub-solvers (xsb_fs.cu)
@pytest.mark.parametrize("name", ["exSaddle3d_fs_1", "exSaddle2d_fs_1", "exSaddle2d_lame_fs_1", "exSaddle3d_lame_fs_1",
                                  "exSaddle3d_fs_2", "exSaddle2d_fs_2", "exSaddle2d_lame_fs_2", "exSaddle3d_lame_fs_2"])
def test_golden_plain_fs_output_is_identical(kat, name):
    """`-fs` without abf.opts (exSaddle.c:303-322, Makefile:282, 347, 396, 480): GMRES + fieldsplit Schur-upper with GMRES + ILU(0)
    of A00 and of Mpscaled, a nested velocity solve inside every Schur-complement product.  The program output (banner, residual
    history as printed by -saddle_ksp_monitor_short, converged-reason line where asked, diagnostics) diffs clean against testref.
    *_fs_2: the reference on 2 ranks (-xsb_ranks 2): both ILU(0)s become bjacobi with one block per rank."""
    c, text, s, x = _run(kat, name, nranks=kat[name]["nranks"])
    ref = list(c["banner"]) + ["  Residual norms for saddle_ solve."] + ["%3d KSP Residual norm %s" % (i, t) for i, t in enumerate(c["residuals_text"])]
    got = [l.rstrip() for l in text.rstrip("\n").split("\n")]
    if "-saddle_ksp_converged_reason" in c["options"]:
        assert got[len(ref)].startswith("Linear saddle_ solve")
        got = got[:len(ref)] + got[len(ref) + 1:]
    assert got == ref + [l.rstrip() for l in c["diagnostics"]]


@pytest.mark.parametrize("opts,nsd,lame", [("-model 6 -mx 3 -my 4 -mz 2 -eta1 10 -fs", 3, False), ("-model 0 -mx 5 -my 4 -fs -saddle_fieldsplit_p_ksp_type preonly", 2, False),
                                           ("-model 8 -mx 3 -fs -saddle_fieldsplit_u_ksp_max_it 7", 3, True)])
def test_plain_fs_tree_matches_oracle(opts, nsd, lame):
    from oracle.oracle_fs import FieldSplitDefault
    g = X.ExSaddle(opts, nsd=nsd, lame=lame).assemble().ksp_setup()
    x = g.solve()
    F = FieldSplitDefault(opts, nsd=nsd, lame=lame)
    xo, its, reason, hist = F.solve()
    assert g.iterations() == (its, reason)
    h = g.history(); ho = np.array(hist)
    assert np.max(np.abs(h - ho)) <= 1e-8 * ho[0]
    assert np.linalg.norm(x - xo) <= 1e-6 * np.linalg.norm(xo)
    g.close()


@pytest.mark.parametrize("opts,nsd,lame", [("-model 1 -mx 4 -my 8 -mz 4 -mg -nlevels 2", 3, False), ("-model 0 -mx 16 -mg -nlevels 4", 2, False),
                                           ("-model 12 -mx 4 -mu1 10 -mg -nlevels 2", 3, True)])
def test_monolithic_mg_matches_oracle(opts, nsd, lame):
    from oracle.oracle_mg import MonolithicMG
    full = opts + " -saddle_ksp_type fgmres -saddle_mg_levels_ksp_type gmres -saddle_mg_levels_pc_type jacobi -saddle_mg_levels_ksp_max_it 10"
    g = X.ExSaddle(full, nsd=nsd, lame=lame).assemble().ksp_setup()
    M = MonolithicMG(full, nsd=nsd, lame=lame)   # configurations chosen to converge within one FGMRES cycle (<= 30 its, rtol 1e-5)
    # one V-cycle on a random residual, then the whole solve
    rng = np.random.default_rng(4)
    r = rng.standard_normal(g.n)
    zg, zo = g.pc_apply(r), M.pc_apply(r)
    assert np.linalg.norm(zg - zo) <= 1e-9 * np.linalg.norm(zo)
    x = g.solve(); xo, its, reason, ho = M.solve()
    assert g.iterations() == (its, reason)
    h = g.history()
    assert np.max(np.abs(h - ho)) <= 1e-9 * ho[0]
    big = ho >= 1e-3 * ho[0]
    assert np.max(np.abs(h - ho)[big] / ho[big]) <= 1e-8
    assert np.linalg.norm(x - xo) <= 1e-6 * np.linalg.norm(xo)
    g.close()


def test_monolithic_mg_option_errors_like_reference():
    for opts, frag in (("-mx 6 -mg -nlevels 3", "incompatible with problem size"), ("-mx 8 -mg -nlevels 1", "-nlevels < 2 specified with -mg"),
                       ("-mx 2 -mg -nlevels 3", "Too much refinement"), ("-mx 8 -fs -mg -nlevels 2", "both -fs and -mg")):
        g = X.ExSaddle(opts + " -saddle_ksp_type fgmres -saddle_mg_levels_ksp_type gmres -saddle_mg_levels_pc_type jacobi", nsd=3).assemble()
        with pytest.raises(X.XsbError) as e:
            g.ksp_setup()
        assert frag in str(e.value)
        g.close()


def test_asm_and_fs_coarse_option_errors_like_reference():
    """exSaddle.c:210, 212 and the limits of the ASM / -fs_coarse support, each with its message."""
    for opts, frag in (("-mx 4 -fs_coarse -saddle_pc_type jacobi", "-fs_coarse supplied without -mg"),
                       ("-mx 4 -fs -set_ksp_dm", "-set_ksp_dm not intended for use with -mg or -fs"),
                       ("-mx 4 -saddle_pc_type asm", "-saddle_pc_asm_dm_subdomains -set_ksp_dm"),
                       ("-mx 4 -saddle_pc_type asm -saddle_pc_asm_dm_subdomains -set_ksp_dm", "ASM sub-solves"),
                       ("-mx 2 -saddle_pc_type asm -saddle_pc_asm_dm_subdomains -set_ksp_dm -saddle_sub_pc_type lu -xsb_ranks 9", "Cannot generate consistent macro element")):
        g = X.ExSaddle(opts, nsd=3).assemble()
        with pytest.raises(X.XsbError) as e:
            g.ksp_setup()
        assert frag in str(e.value), (opts, str(e.value))
        g.close()


def test_golden_monolithic_mg_fs_coarse(kat):
    """-mg -fs_coarse (exSaddle.c:362-400, Makefile:390): the coarse saddle level solved by FGMRES preconditioned with fieldsplit
    Schur / UPPER / user Mpscaled_coarse (GMRES + Jacobi splits, nested velocity solve inside the Schur complement), on the device.
    testref/exSaddle3d_mg_fs_coarse_1.ref: 13 iterations, CONVERGED_RTOL, the first six residuals to every printed digit, all to
    3e-5 (the nested inexact solves at rtol 1e-5 amplify summation-order differences: the oracle meets the same bar); and the
    oracle's history to 1e-4 entry-wise."""
    from oracle.oracle_mg import MonolithicMG
    c = kat["exSaddle3d_mg_fs_coarse_1"]
    opts = c["options"].replace("-saddle_ksp_view", "")
    g = X.ExSaddle(opts, nsd=3).assemble().ksp_setup()
    g.solve()
    its, reason = g.iterations(); h = g.history()
    assert (its, reason) == (c["iterations"], 2) and len(h) == len(c["residuals"])
    assert [X.monitor_short(v) for v in h[:6]] == c["residuals_text"][:6]
    assert np.allclose(h, c["residuals"], rtol=3e-5, atol=0)
    M = MonolithicMG(c["options"], nsd=3)
    xo, ito, ro, ho = M.solve()
    assert ito == its and np.allclose(h, ho, rtol=1e-4, atol=0)
    g.close()


# ------------------------------------------------------------------ BASELINE sizes against committed oracle fixtures
@pytest.mark.parametrize("name", ["exSaddle2d_asm_1", "exSaddle3d_asm_1", "exSaddle3d_mg_asm_1"])
def test_golden_asm_element_patches_output_is_identical(kat, name):
    """ASM on the reference's element patches, one per rank of the golden's communicator (-xsb_ranks 9 / 8 / 4),
    dense pivoted inverses + one-launch apply on the device -- as the top-level PC (Makefile:298, 411) and as the PC of the -mg
    GMRES smoother (:418).  The program output matches testref/*.ref (one last-digit difference allowed)."""
    c, text, s, x = _run(kat, name, nranks=kat[name]["nranks"])
    ref = list(c["banner"]) + ["  Residual norms for saddle_ solve."] + ["%3d KSP Residual norm %s" % (i, t) for i, t in enumerate(c["residuals_text"])]
    got = [l.rstrip() for l in text.rstrip("\n").split("\n")]
    # the sub-solves are exact but not UMFPACK's: at most one residual may differ, by one unit of its sixth digit
    assert len(got) == len(ref) and sum(a != b for a, b in zip(got, ref)) <= 1
    assert np.allclose(s.history(), c["residuals"], rtol=6e-6, atol=0)   # six printed digits


@pytest.mark.parametrize("nsd,size,opts", [(2, 6, "-model 6 -mx 8 -my 5 -eta1 100 -dmdafe_overlap 1"), (3, 5, "-model 1 -mx 4 -my 3 -mz 10 -eta1 10 -dmdafe_overlap 1"),
                                           (3, 1, "-model 6 -mx 3 -eta1 1e3")])
def test_asm_pc_apply_and_solve_match_oracle(nsd, size, opts):
    """PC apply <= 1e-9 (exact dense sub-solves of indefinite patches against SuperLU), iteration count and history against the oracle;
    one rank = one patch = a direct solve (1 iteration)."""
    from oracle import oracle_asm as OA
    full = "-saddle_pc_type asm -saddle_pc_asm_dm_subdomains -set_ksp_dm -saddle_sub_ksp_type preonly -saddle_sub_pc_type lu -saddle_ksp_rtol 1e-8 " + opts
    xo, its, reason, hist, p = OA.solve(full, nsd, size)
    g = X.ExSaddle(full + " -xsb_ranks %d" % size, nsd=nsd).assemble().ksp_setup()
    A = p.A().scipy().tocsr()
    o = O.parse_options(full); mx = int(o.get("mx", 4)); my = int(o.get("my", mx)); mz = int(o.get("mz", mx)) if nsd == 3 else 1
    pc = OA.AsmPC(A, nsd, (mx, my, mz), size, int(o.get("dmdafe_overlap", 0)))
    r = np.cos(0.7 * np.arange(p.n)) + 0.3
    zg = g.pc_apply(r); zo = pc(r)
    assert np.linalg.norm(zg - zo) <= 1e-9 * np.linalg.norm(zo)
    x = g.solve()
    gi, gr = g.iterations()
    assert gr == reason and abs(gi - its) <= 1     # 100+ iterations with restarts: rounding may move the last one across the tolerance
    h = g.history(); k = min(len(h), len(hist), 31)
    assert np.max(np.abs(h[:k] - np.array(hist[:k]))) <= 1e-8 * hist[0]    # first restart cycle
    if size == 1:
        assert its == 1
    assert np.linalg.norm(x - xo) <= 1e-6 * np.linalg.norm(xo)
    g.close()


@pytest.mark.parametrize("mx,levels", [(32, 5), (64, 6)])
def test_baseline_size_history_matches_oracle_fixture(mx, levels):
    """bench.py's workloads (configs[1] 32^3, configs[2] 64^3): the GPU residual history against the CPU oracle's, generated
    once by tests/golden/make_oracle_64cubed.py (the 64^3 oracle solve takes ~8 min on 8 cores, so it is a fixture).
    north_star asks for 1e-8 relative histories and iteration counts +-1.  The counts hold (32^3: 51 = 51; 64^3: 42 or 43
    against 42).  The 1e-8 history bound holds at the sizes of the tests above (<= 8^3) and for the well-conditioned Lame
    config at 64^3 (next test).  At eta1/eta0 = 1e6 it is below the floor ANY two correct evaluations can agree to: the
    ORACLE AGAINST ITSELF, with nothing changed but the order of its dot-product / matrix-row sums (scripts/oracle_drift.py),
    moves its own history by 8.7e-7 (32^3) and 2.1e-6 (64^3) while the residual is above 1e-2 ||r0||, with identical
    outer and inner iteration counts (profiles/r02_oracle_drift_{32,64}cubed.json).  The GPU differs from the oracle by
    2.6e-6 at 64^3 (profiles/r01_fixture_probe.json): the same floor.  The bounds below are those measurements with a
    factor ~4 of head room; the floor itself is asserted by test_oracle_goldens.py::test_history_floor_at_baseline_sizes."""
    import json, os
    path = os.path.join(os.path.dirname(__file__), "golden", "oracle_%dcubed_history.json" % mx)
    fx = json.load(open(path))
    ho = np.array(fx["hist"])
    for extra in ("", " -xsb_matrix_free full"):
        g = X.ExSaddle(fx["options"] + extra, nsd=3).assemble().ksp_setup()
        x = g.solve()
        its, reason = g.iterations()
        h = g.history()
        assert reason == fx["reason"] and abs(its - fx["its"]) <= 1
        n = min(len(h), len(ho)); d = np.abs(h[:n] - ho[:n])
        assert np.max(d) <= 1e-6 * ho[0]
        big = ho[:n] >= 1e-2 * ho[0]; mid = ho[:n] >= 1e-4 * ho[0]
        assert np.max(d[big] / ho[:n][big]) <= 1e-5 and np.max(d[mid] / ho[:n][mid]) <= 6e-5
        inner, inner_o = g.inner_iterations(), fx["inner_its"]
        m = min(len(inner), len(inner_o))
        assert sum(abs(a - b) for a, b in zip(inner[:m], inner_o[:m])) <= 1      # GCR counts of every outer iteration
        for l in range(1, levels):
            assert abs(g.chebyshev(l)[1] - fx["cheb_emax_est"][l]) <= 1e-12 * fx["cheb_emax_est"][l]
        assert abs(np.linalg.norm(x) - fx["x_norm2"]) <= 1e-7 * fx["x_norm2"]
        g.close()


def test_lame_64cubed_history_matches_oracle_fixture():
    """BASELINE config 4 (exSaddle3d_lame -options_file abf.opts -mx 64 -model 6; 6 MG levels, rtol 1e-8): both GPU paths
    against the oracle fixture tests/golden/oracle_64cubed_lame_history.json (make_oracle_lame64.py).  The system is well
    conditioned (mu 1/1, lambda 1/2), so north_star's bars hold as stated: equal iteration counts (outer and every inner),
    residual history within 1e-8 relative -- entry-wise down to 1e-6 ||r0|| (measured: <= 4e-10), and within 1e-11 ||r0|| in
    absolute terms all the way to convergence at 1.5e-9 ||r0||, where an entry-wise ratio only measures rounding of the
    Givens recurrence (1.4e-6 of a 1.4e-12 residual)."""
    import json, os
    fx = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "oracle_64cubed_lame_history.json")))
    ho = np.array(fx["hist"])
    for extra in ("", " -xsb_matrix_free full"):
        g = X.ExSaddle(fx["options"] + extra, nsd=3, lame=True).assemble().ksp_setup()
        x = g.solve()
        its, reason = g.iterations()
        h = g.history()
        assert (its, reason) == (fx["its"], fx["reason"])
        assert g.inner_iterations() == fx["inner_its"]
        big = ho >= 1e-6 * ho[0]
        assert np.max(np.abs(h - ho)[big] / ho[big]) <= 1e-8 and np.max(np.abs(h - ho)) <= 1e-11 * ho[0]
        for l in range(1, fx["levels"]):
            assert abs(g.chebyshev(l)[1] - fx["cheb_emax_est"][l]) <= 1e-12 * fx["cheb_emax_est"][l]
        assert abs(np.linalg.norm(x) - fx["x_norm2"]) <= 1e-8 * fx["x_norm2"]
        F = g.rhs()
        assert np.linalg.norm(F - g.mat_mult(X.MAT_A, x)) <= 1.5e-8 * np.linalg.norm(F)
        g.close()


# ------------------------------------------------------------------ abf.opts verbatim: 3 levels, large coarsest level
@pytest.mark.parametrize("opts,lame", [("-model 6 -mx 8 -eta1 1e4", False), ("-model 1 -mx 4 -my 8 -mz 4 -eta1 10", False), ("-model 6 -mx 8 -mu1 10", True)])
def test_internal_coarse_hierarchy_reproduces_the_dense_coarse_solve(opts, lame):
    """A coarsest level above the dense-inverse limit is solved by CG preconditioned with an internal V-cycle, to LU accuracy
    (abf.opts:7,16 leaves 14 739 / 107 811 coarse rows at 32^3 / 64^3).  Forced here on a small mesh by lowering the limit:
    same iteration counts as the dense coarse solve and as the oracle (dense LU), histories within the usual bars."""
    full = "%s %s -saddle_fieldsplit_u_pc_mg_levels 2 -saddle_ksp_rtol 1e-8" % (ABF, opts)
    gd = X.ExSaddle(full, nsd=3, lame=lame).assemble().ksp_setup()
    gc = X.ExSaddle(full + " -xsb_coarse_dense_max 100", nsd=3, lame=lame).assemble().ksp_setup()
    assert "cg to 1e-13" in gc.view() and "cg to 1e-13" not in gd.view()
    o = O.Problem(full, nsd=3, lame=lame)
    xo, r = o.solve()
    rng = np.random.default_rng(5)
    b = rng.standard_normal(gd.nu)
    zd, zc = gd.pc_mg_apply(b), gc.pc_mg_apply(b)                  # one V-cycle with either coarse solver
    assert np.linalg.norm(zd - zc) <= 1e-11 * np.linalg.norm(zd)
    xd, xc = gd.solve(), gc.solve()
    assert gc.iterations() == gd.iterations() == (r.its, r.reason)
    assert gc.inner_iterations() == gd.inner_iterations() == [int(v) for v in r.inner_its[:r.n_inner]]
    ho = np.array(r.hist[:r.nhist]); hc = gc.history()
    assert np.max(np.abs(hc - ho)) <= 1e-9 * ho[0]
    assert np.linalg.norm(xc - xd) <= 1e-8 * np.linalg.norm(xd)
    gd.close(); gc.close()


@pytest.mark.parametrize("mx", [32, 64])
def test_abf_opts_verbatim_matches_oracle_fixture(mx):
    """`-options_file abf.opts` UNCHANGED (3 MG levels, LU on the coarsest) at the BASELINE sizes: the GPU solves the 17^3 / 33^3
    node coarsest level by CG + internal V-cycle, the oracle by a banded Cholesky factorisation (tests/golden/
    oracle_<mx>cubed_abf3_history.json, make_oracle_64cubed.py <mx> 3).  Counts +-1 and the history floor of the 1e6 contrast."""
    import json, os
    path = os.path.join(os.path.dirname(__file__), "golden", "oracle_%dcubed_abf3_history.json" % mx)
    if not os.path.exists(path):
        pytest.skip("fixture %s not generated" % os.path.basename(path))
    fx = json.load(open(path))
    assert fx["levels"] == 3
    ho = np.array(fx["hist"])
    for extra in ("", " -xsb_matrix_free full"):
        g = X.ExSaddle(fx["options"] + extra, nsd=3).assemble().ksp_setup()
        x = g.solve()
        its, reason = g.iterations()
        h = g.history()
        assert reason == fx["reason"] and abs(its - fx["its"]) <= 1
        n = min(len(h), len(ho)); d = np.abs(h[:n] - ho[:n])
        big = ho[:n] >= 1e-2 * ho[0]
        assert np.max(d[big] / ho[:n][big]) <= 1e-5 and np.max(d) <= 1e-6 * ho[0]
        inner, inner_o = g.inner_iterations(), fx["inner_its"]
        m = min(len(inner), len(inner_o))
        assert sum(abs(a - b) for a, b in zip(inner[:m], inner_o[:m])) <= 1
        F = g.rhs()
        # the true residual tracks the oracle's: at 32^3 the solve ends at iteration 30, just before the first restart would
        # re-compute it, and one-pass classical Gram-Schmidt has let the estimate run ahead of it (oracle: 1.55e-7)
        assert np.linalg.norm(F - g.mat_mult(X.MAT_A, x)) <= (1.5 * fx["true_rel_res"] + 1e-8) * np.linalg.norm(F)
        g.close()


# ------------------------------------------------------------------ MatMultTranspose / KSPView entry points
def test_mult_transpose_and_view():
    opts = ABF + " -model 6 -mx 4 -eta1 100 -saddle_fieldsplit_u_pc_mg_levels 2"
    g = X.ExSaddle(opts, nsd=3).assemble().ksp_setup()
    o = O.Problem(opts, nsd=3)
    rng = np.random.default_rng(9)
    for which, rb, cb in ((X.MAT_A01, 0, 1), (X.MAT_A10, 1, 0), (X.MAT_A11, 1, 1), (X.MAT_A00, 0, 0)):
        M = o.submatrix(rb, cb).scipy()
        x = rng.standard_normal(M.shape[0])
        yo = M.T @ x
        y = g.mat_mult_transpose(which, x)
        assert y.shape == yo.shape and np.linalg.norm(y - yo) <= 1e-12 * max(np.linalg.norm(yo), 1e-300)
    x = rng.standard_normal(o.n)
    yo = o.A().scipy().T @ x
    assert np.linalg.norm(g.mat_mult_transpose(X.MAT_A, x) - yo) <= 1e-12 * np.linalg.norm(yo)
    text = g.view()
    for frag in ("type: fgmres", "FieldSplit with Schur preconditioner, factorization UPPER", "level 1: chebyshev + jacobi", "ilu(0)", "type aij"):
        assert frag in text, text
    g.close()


def test_petsc_binary_dumps_of_operator_and_solution(tmp_path):
    """-dump_operator / -dump_solution / -dump_scaled_mass_matrix (exSaddle.c:488-501, 535-537): the files hold exactly the
    operator and solution the handle reports, in PETSc's binary layout, under the reference's file names."""
    opts = ABF + " -model 6 -mx 4 -eta1 100 -saddle_fieldsplit_u_pc_mg_levels 2 -diagnostics -dump_solution -dump_operator -dump_scaled_mass_matrix"
    text, s, x = X.run_exsaddle("exSaddle3d", opts, outdir=str(tmp_path))
    assert "Dumping solution vector to solution.petscbin." in text and "Finished dumping operator to operator_0.petscbin." in text
    kind, xs = X.read_petsc_binary(str(tmp_path / "solution.petscbin"))
    assert kind == "Vec" and np.array_equal(xs, x)
    ia, ja, a, shape = s.mat_csr(X.MAT_A)
    kind, (ia2, ja2, a2, shape2) = X.read_petsc_binary(str(tmp_path / "operator_0.petscbin"))
    assert kind == "Mat" and tuple(shape2) == tuple(shape) and np.array_equal(ia2, ia) and np.array_equal(ja2, ja) and np.array_equal(a2, a)
    o = O.Problem(opts, nsd=3)
    kind, (mia, mja, ma, mshape) = X.read_petsc_binary(str(tmp_path / "mpscaled.petscbin"))
    M = o.Mp()
    assert np.array_equal(mia, M.ia) and np.array_equal(mja, M.ja) and _relerr(ma, M.a) <= 1e-13
    s.close()


