"""CPU oracle of exSaddle's monolithic `-mg` path.  TEST INFRASTRUCTURE, NOT PRODUCT.

What the reference does with `-mg -nlevels L` (exSaddle.c:215-270, 331-402):
  * one Q2-Q1 mesh per level, m_k = m / 2^(L-1-k) elements per side (exSaddle.c:217-241);
  * coefficients: the fine level's nodal Q1 fields (femixedspace.c:1976-2083) are restricted level by level on the
    pressure lattice, c_k = (P_p^T c_{k+1}) .* 1 / (P_p^T 1)  (MatRestrict + DMCreateInterpolationScale, :2139-2150),
    and interpolated to the coarse quadrature points (:2168-2215);
  * every level operator is RE-ASSEMBLED (PC_MG_GALERKIN_NONE, exSaddle.c:265-270, 339) with its own Dirichlet rows;
  * interpolation between levels = DMComposite block-diagonal (P_u (x) I_nsd, P_p), both (tri)linear on their node
    lattices (DMCreateInterpolation, exSaddle.c:348);
  * PCMG: multiplicative V-cycle, smoothers = exactly `max_it` iterations of left-Jacobi GMRES from the current iterate
    (convergence test skipped), coarse = LU of the coarse saddle matrix; outer FGMRES, right PC.
Assembly comes from the C oracle (oracle/xo_fe.c); the Krylov / multigrid logic here is plain numpy + scipy.sparse.
Pinned by tests/test_oracle_goldens.py against testref/exSaddle{2d,3d}[_lame]_mg_1.ref (residual histories and
diagnostics to the printed digits)."""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from . import oracle as O


def interp_1d(nc):
    """DMDA Q1 interpolation on a line: nc coarse nodes -> 2 nc - 1 fine nodes."""
    nf = 2 * nc - 1
    rows, cols, vals = [], [], []
    for i in range(nf):
        if i % 2 == 0:
            rows.append(i); cols.append(i // 2); vals.append(1.0)
        else:
            rows += [i, i]; cols += [(i - 1) // 2, (i + 1) // 2]; vals += [0.5, 0.5]
    return sp.csr_matrix((vals, (rows, cols)), shape=(nf, nc))


def interp_lattice(dims_c):
    """(tri)linear interpolation on a node lattice, x fastest; dims_c = coarse nodes per direction (1 = flat)."""
    P = None
    for n in dims_c:   # x first (fastest index): kron(Pz, kron(Py, Px))
        Pd = interp_1d(n) if n > 1 else sp.identity(1, format="csr")
        P = Pd if P is None else sp.kron(Pd, P, format="csr")
    return P.tocsr()


def petsc_gmres(apply_A, apply_M, b, rtol=1e-5, atol=1e-50, dtol=1e4, max_it=10000, restart=30, side="left", flexible=False):
    """KSPSolve_GMRES / KSPSolve_FGMRES from a zero initial guess (App. B.5): classical Gram-Schmidt in one pass, Givens QR
    of the Hessenberg, residual estimate |g_{j+1}|, default convergence test rnorm <= max(rtol * rnorm0, atol).
    side="left": iterates on M^-1 A and measures the preconditioned residual; "right" (and FGMRES): unpreconditioned.
    Returns (x, its, reason, hist)."""
    n = len(b); x = np.zeros(n); hist = []; its = 0; reason = 0; rnorm0 = None; ttol = 0.0
    left = side == "left" and not flexible
    while not reason:
        r = b - apply_A(x) if its else b.copy()
        if left:
            r = apply_M(r)
        res = np.linalg.norm(r)
        if rnorm0 is None:
            rnorm0 = res; ttol = max(rtol * rnorm0, atol)
        if len(hist) == its:
            hist.append(res)
        if res == 0.0 or res <= ttol:
            reason = 2 if res > 0.0 or rnorm0 > 0.0 else 3; break
        if its >= max_it:
            reason = -3; break
        m = restart
        V = np.zeros((m + 1, n)); Z = np.zeros((m, n)) if (flexible or not left) else None; H = np.zeros((m + 1, m))
        V[0] = r / res
        g = np.zeros(m + 1); g[0] = res; cs = np.zeros(m); sn = np.zeros(m); R = np.zeros((m + 1, m))
        k = 0
        for j in range(m):
            if left:
                w = apply_M(apply_A(V[j]))
            else:
                Z[j] = apply_M(V[j]); w = apply_A(Z[j])
            h = V[:j + 1] @ w
            w = w - V[:j + 1].T @ h
            hn = np.linalg.norm(w)
            H[:j + 1, j] = h; H[j + 1, j] = hn
            if hn != 0.0:
                V[j + 1] = w / hn
            col = H[:j + 2, j].copy()
            for i in range(j):
                t = col[i]; col[i] = cs[i] * t + sn[i] * col[i + 1]; col[i + 1] = -sn[i] * t + cs[i] * col[i + 1]
            tt = np.hypot(col[j], col[j + 1])
            if tt == 0.0:
                reason = -5; break
            cs[j] = col[j] / tt; sn[j] = col[j + 1] / tt
            g[j + 1] = -sn[j] * g[j]; g[j] = cs[j] * g[j]
            col[j] = cs[j] * col[j] + sn[j] * col[j + 1]; col[j + 1] = 0.0
            R[:j + 2, j] = col
            res = abs(g[j + 1]); k = j + 1; its += 1
            hist.append(res)
            if res <= ttol:
                reason = 2
            elif res >= dtol * rnorm0:
                reason = -4
            elif its >= max_it:
                reason = -3
            elif hn == 0.0:
                reason = 2      # happy breakdown
            if reason:
                break
        if k:
            y = np.zeros(k)
            for i in range(k - 1, -1, -1):
                y[i] = (g[i] - R[i, i + 1:k] @ y[i + 1:k]) / R[i, i]
            x = x + (V[:k].T @ y if left else Z[:k].T @ y)
    return x, its, reason, np.array(hist)


class Level:
    pass


class MonolithicMG:
    def __init__(self, opts, nsd=3, lame=False, nranks=1):
        self.o = O.parse_options(opts) if not isinstance(opts, dict) else dict(opts)
        o = self.o
        self.nsd, self.lame = nsd, lame
        if "mg" not in o:
            raise ValueError("MonolithicMG needs -mg")
        L = int(o.get("nlevels", 1))
        if L < 2:
            raise ValueError("-nlevels < 2 specified with -mg")   # exSaddle.c:209
        mx = int(o.get("mx", 4)); my = int(o.get("my", mx)); mz = int(o.get("mz", mx)) if nsd == 3 else 1
        ratio = 2 ** (L - 1)
        for m in (mx, my) + ((mz,) if nsd == 3 else ()):
            if ratio > m or m % ratio:
                raise ValueError("Coarsening ratio of 2 ^ %d = %d is incompatible with problem size" % (L - 1, ratio))   # exSaddle.c:219-220
        self.levels = [None] * L
        fine = O.Problem(o, nsd=nsd, lame=lame)
        nodal = fine.coeff_nodal()
        probs = [None] * L; probs[L - 1] = fine
        mesh = [None] * L; mesh[L - 1] = (mx, my, mz)
        for k in range(L - 2, -1, -1):
            f = 2 ** (L - 1 - k)
            mk = (mx // f, my // f, (mz // f) if nsd == 3 else 1)
            pd = (mk[0] + 1, mk[1] + 1, (mk[2] + 1) if nsd == 3 else 1)
            Pp = interp_lattice(pd)
            scale = 1.0 / (Pp.T @ np.ones(Pp.shape[0]))           # DMCreateInterpolationScale
            nodal = (Pp.T @ nodal) * scale[:, None]                # MatRestrict, then VecPointwiseMult (:2148-2149)
            ok = dict(o); ok["mx"] = str(mk[0]); ok["my"] = str(mk[1])
            if nsd == 3:
                ok["mz"] = str(mk[2])
            probs[k] = O.Problem(ok, nsd=nsd, lame=lame, nodal=nodal); mesh[k] = mk
        for k in range(L):
            lv = Level(); p = probs[k]
            lv.p = p; lv.A = p.A().scipy().tocsr(); lv.n = p.n
            d = lv.A.diagonal(); lv.idiag = np.where(d == 0.0, 1.0, 1.0 / np.where(d == 0.0, 1.0, d))   # PCSetUp_Jacobi
            if k > 0:
                mkc = mesh[k - 1]
                ud = (2 * mkc[0] + 1, 2 * mkc[1] + 1, (2 * mkc[2] + 1) if nsd == 3 else 1)
                pd = (mkc[0] + 1, mkc[1] + 1, (mkc[2] + 1) if nsd == 3 else 1)
                Pu = sp.kron(interp_lattice(ud), sp.identity(nsd), format="csr")   # MAIJ over the components
                lv.P = sp.block_diag([Pu, interp_lattice(pd)], format="csr")
            self.levels[k] = lv
        self.lu = spla.splu(self.levels[0].A.tocsc())
        self.fs_coarse = "fs_coarse" in o
        if self.fs_coarse:   # exSaddle.c:366-398: FGMRES + fieldsplit Schur-upper (user Mpscaled_coarse) on the coarse level
            for key, want in (("saddle_mg_coarse_ksp_type", "fgmres"), ("saddle_mg_coarse_fieldsplit_u_pc_type", "jacobi"),
                              ("saddle_mg_coarse_fieldsplit_p_pc_type", "jacobi"), ("saddle_mg_coarse_ksp_convergence_test", "default")):
                if o.get(key) != want:
                    raise NotImplementedError("oracle -fs_coarse: the golden's tree only (-%s %s)" % (key, want))
            c = probs[0]; lv0 = self.levels[0]; nu = c.nu
            A = lv0.A
            lv0.A00 = A[:nu, :nu].tocsr(); lv0.A01 = A[:nu, nu:].tocsr(); lv0.A10 = A[nu:, :nu].tocsr(); lv0.A11 = A[nu:, nu:].tocsr()
            jac = lambda d: np.where(d == 0.0, 1.0, 1.0 / np.where(d == 0.0, 1.0, d))
            lv0.id00 = jac(lv0.A00.diagonal()); lv0.idmp = jac(c.Mp().scipy().diagonal()); lv0.nu = nu
            self.coarse_its = []
        self.fine = fine
        self.smooth_its = int(o.get("saddle_mg_levels_ksp_max_it", 2))   # PCMG default: 2 smoothing steps
        self.restart = int(o.get("saddle_mg_levels_ksp_gmres_restart", 30))
        spc = o.get("saddle_mg_levels_pc_type", "sor")
        if o.get("saddle_mg_levels_ksp_type", "chebyshev") != "gmres" or spc not in ("jacobi", "asm"):
            raise NotImplementedError("oracle -mg smoothers: gmres + jacobi | asm (the reference's tests)")
        for k in range(L):
            lv = self.levels[k]
            lv.pc = (lambda d: (lambda v: d * v))(lv.idiag)
        if spc == "asm":   # Makefile:418 (exSaddle3d_mg_asm_1): element-patch ASM with exact sub-solves as the smoother's PC, one patch per rank
            from .oracle_asm import AsmPC
            if "saddle_mg_levels_pc_asm_dm_subdomains" not in o or o.get("saddle_mg_levels_sub_pc_type", "ilu") != "lu":
                raise NotImplementedError("oracle -mg ASM smoother: -saddle_mg_levels_pc_asm_dm_subdomains -saddle_mg_levels_sub_pc_type lu")
            for k in range(1, L):
                self.levels[k].pc = AsmPC(self.levels[k].A, nsd, mesh[k], nranks, int(o.get("dmdafe_overlap", 0)))
        self.n_smooth_mult = 0

    # -- KSPSolve_GMRES, left Jacobi, exactly `its` iterations from the current iterate (KSPConvergedSkip)
    def smooth(self, lv, b, x, its):
        done = 0
        while done < its:
            m = min(self.restart, its - done)
            r = lv.pc(b - lv.A @ x); self.n_smooth_mult += 1
            beta = np.linalg.norm(r)
            if beta == 0.0:
                return x
            V = np.zeros((m + 1, lv.n)); H = np.zeros((m + 1, m))
            V[0] = r / beta
            k = 0
            for j in range(m):
                w = lv.pc(lv.A @ V[j]); self.n_smooth_mult += 1
                h = V[:j + 1] @ w                       # classical Gram-Schmidt, one pass
                w = w - V[:j + 1].T @ h
                hn = np.linalg.norm(w)
                H[:j + 1, j] = h; H[j + 1, j] = hn
                k = j + 1
                if hn == 0.0:
                    break
                V[j + 1] = w / hn
            e1 = np.zeros(k + 1); e1[0] = beta
            y = np.linalg.lstsq(H[:k + 1, :k], e1, rcond=None)[0]
            x = x + V[:k].T @ y
            done += k
            if k < m:
                break
        return x

    def vcycle(self, l, b):
        lv = self.levels[l]
        if l == 0:
            return self.coarse_fieldsplit(b) if self.fs_coarse else self.lu.solve(b)
        x = self.smooth(lv, b, np.zeros(lv.n), self.smooth_its)
        r = b - lv.A @ x
        xc = self.vcycle(l - 1, lv.P.T @ r)
        x = x + lv.P @ xc
        return self.smooth(lv, b, x, self.smooth_its)

    # -- coarse solver of -fs_coarse: FGMRES(rtol 1e-5) with PCFIELDSPLIT Schur / UPPER / user Mpscaled_coarse (App. B.2):
    #      y_p = GMRES[S, Jacobi(Mp)] x_p with S v = A11 v - A10 GMRES[A00, Jacobi](A01 v);  y_u = GMRES[A00, Jacobi](x_u - A01 y_p)
    def coarse_fieldsplit(self, b):
        L = self.levels[0]; nu = L.nu
        ksp_u = lambda rhs: petsc_gmres(lambda v: L.A00 @ v, lambda v: L.id00 * v, rhs)[0]
        S = lambda v: L.A11 @ v - L.A10 @ ksp_u(L.A01 @ v)

        def pc(r):
            yp = petsc_gmres(S, lambda v: L.idmp * v, r[nu:])[0]
            yu = ksp_u(r[:nu] - L.A01 @ yp)
            return np.concatenate([yu, yp])
        x, its, reason, _ = petsc_gmres(lambda v: L.A @ v, pc, b, flexible=True)
        self.coarse_its.append(its)
        return x

    def pc_apply(self, r):
        return self.vcycle(len(self.levels) - 1, r)

    # -- KSPSolve_FGMRES, right PC, unpreconditioned norm (App. B.5)
    def solve(self, b=None):
        o = self.o
        if o.get("saddle_ksp_type", "gmres") != "fgmres":   # the default: GMRES, left PC, preconditioned norm (mg_fs_coarse_1.ref)
            A = self.levels[-1].A
            return petsc_gmres(lambda v: A @ v, self.pc_apply, self.fine.F() if b is None else b,
                               rtol=float(o.get("saddle_ksp_rtol", 1e-5)), max_it=int(o.get("saddle_ksp_max_it", 10000)),
                               restart=int(o.get("saddle_ksp_gmres_restart", 30)), side=o.get("saddle_ksp_pc_side", "left"))
        rtol = float(o.get("saddle_ksp_rtol", 1e-5)); atol = float(o.get("saddle_ksp_atol", 1e-50)); dtol = 1e4
        max_it = int(o.get("saddle_ksp_max_it", 10000)); m = int(o.get("saddle_ksp_gmres_restart", 30))
        A = self.levels[-1].A; n = A.shape[0]
        b = self.fine.F() if b is None else b
        x = np.zeros(n); hist = []; its = 0; reason = 0; rnorm0 = None
        while not reason:
            r = b - A @ x if its else b.copy()
            res = np.linalg.norm(r)
            if rnorm0 is None:
                rnorm0 = res; ttol = max(rtol * rnorm0, atol)
            if len(hist) == its:
                hist.append(res)
            if res <= ttol:
                reason = 2; break
            if its >= max_it:
                reason = -3; break
            V = np.zeros((m + 1, n)); Z = np.zeros((m, n)); H = np.zeros((m + 1, m))
            V[0] = r / res
            g = np.zeros(m + 1); g[0] = res; cs = np.zeros(m); sn = np.zeros(m)
            k = 0
            for j in range(m):
                Z[j] = self.pc_apply(V[j])
                w = A @ Z[j]
                h = V[:j + 1] @ w
                w = w - V[:j + 1].T @ h
                hn = np.linalg.norm(w)
                H[:j + 1, j] = h; H[j + 1, j] = hn
                if hn != 0.0:
                    V[j + 1] = w / hn
                col = H[:j + 2, j].copy()
                for i in range(j):
                    t = col[i]; col[i] = cs[i] * t + sn[i] * col[i + 1]; col[i + 1] = -sn[i] * t + cs[i] * col[i + 1]
                tt = np.hypot(col[j], col[j + 1])
                cs[j] = col[j] / tt; sn[j] = col[j + 1] / tt
                g[j + 1] = -sn[j] * g[j]; g[j] = cs[j] * g[j]
                res = abs(g[j + 1])
                k = j + 1; its += 1
                hist.append(res)
                if res <= ttol:
                    reason = 2
                elif res >= dtol * rnorm0:
                    reason = -4
                elif its >= max_it:
                    reason = -3
                if reason:
                    break
            e1 = np.zeros(k + 1); e1[0] = np.linalg.norm(r)
            y = np.linalg.lstsq(H[:k + 1, :k], e1, rcond=None)[0]
            x = x + Z[:k].T @ y
        return x, its, reason, np.array(hist)
