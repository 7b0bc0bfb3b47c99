"""ctypes front end of the CPU oracle (oracle/xo_*.c).

TEST INFRASTRUCTURE, NOT PRODUCT: only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs import this module.  The product package
(exsaddle_b200) never does.

`Problem(opts)` takes the reference's own command-line strings (e.g. the option strings of
/root/reference/Makefile:254-513 and abf.opts) and reproduces exSaddle.c:124-425 on one rank.
"""
import ctypes as C
import math
import os
import shlex
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "libxo.so")
XO_MAX_LEVELS = 10


def build(force=False):
    """Compile the C restatement with gcc (make -C oracle)."""
    srcs = [os.path.join(_HERE, f) for f in ("xo_fe.c", "xo_solve.c", "xo.h", "xo_internal.h", "Makefile")]
    if not force and os.path.exists(_LIB) and all(os.path.getmtime(_LIB) >= os.path.getmtime(s) for s in srcs):
        return _LIB
    subprocess.run(["make", "-C", _HERE], check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    return _LIB


class Params(C.Structure):
    _fields_ = [("nsd", C.c_int), ("lame", C.c_int), ("mx", C.c_int), ("my", C.c_int), ("mz", C.c_int),
                ("size", C.c_double * 3), ("model", C.c_int), ("c0", C.c_double), ("c1", C.c_double),
                ("lam0", C.c_double), ("lam1", C.c_double), ("sinker_r", C.c_double), ("sinker_c", C.c_double * 3),
                ("sinker_n", C.c_int), ("solcx_xc", C.c_double), ("solcx_nz", C.c_int), ("freeslip", C.c_int)]


class Solver(C.Structure):
    _fields_ = [("ksp_type", C.c_int), ("pc_type", C.c_int), ("pc_side", C.c_int),
                ("rtol", C.c_double), ("atol", C.c_double), ("dtol", C.c_double),
                ("max_it", C.c_int), ("restart", C.c_int),
                ("u_rtol", C.c_double), ("u_max_it", C.c_int), ("u_restart", C.c_int),
                ("mg_levels", C.c_int), ("cheb_its", C.c_int), ("esteig", C.c_double * 4),
                ("esteig_steps", C.c_int), ("noise", C.c_int), ("n_cheb_fixed", C.c_int),
                ("cheb_emin", C.c_double * XO_MAX_LEVELS), ("cheb_emax", C.c_double * XO_MAX_LEVELS),
                ("p_pc", C.c_int), ("max_outer_sample", C.c_int), ("p_blocks", C.c_int)]


class Result(C.Structure):
    _fields_ = [("its", C.c_int), ("reason", C.c_int), ("nhist", C.c_int), ("hist", C.c_double * 2048),
                ("inner_its", C.c_int * 2048), ("n_inner", C.c_int),
                ("cheb_emin_est", C.c_double * XO_MAX_LEVELS), ("cheb_emax_est", C.c_double * XO_MAX_LEVELS),
                ("cheb_emin", C.c_double * XO_MAX_LEVELS), ("cheb_emax", C.c_double * XO_MAX_LEVELS),
                ("level_rows", C.c_int * XO_MAX_LEVELS), ("level_nnz", C.c_int64 * XO_MAX_LEVELS),
                ("setup_seconds", C.c_double), ("solve_seconds", C.c_double),
                ("n_a00_mult", C.c_int64), ("n_a_mult", C.c_int64)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB)
        vp, ip, dp = C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_double)
        L.xo_params_init.argtypes = [C.POINTER(Params), C.c_int, C.c_int]
        L.xo_solver_init.argtypes = [C.POINTER(Solver)]
        L.xo_solver_abf.argtypes = [C.POINTER(Solver)]
        L.xo_create.argtypes = [C.POINTER(Params), C.POINTER(vp)]
        L.xo_create_nodal.argtypes = [C.POINTER(Params), dp, C.POINTER(vp)]
        L.xo_coeff_nodal.argtypes = [vp]; L.xo_coeff_nodal.restype = dp
        L.xo_destroy.argtypes = [vp]
        L.xo_banner.argtypes = [vp]; L.xo_banner.restype = C.c_char_p
        L.xo_error.argtypes = [vp]; L.xo_error.restype = C.c_char_p
        L.xo_sizes.argtypes = [vp, C.POINTER(C.c_int64)]
        for name, rt in (("xo_A_ia", ip), ("xo_A_ja", ip), ("xo_A_a", dp), ("xo_A_raw", dp), ("xo_Mp_ia", ip),
                         ("xo_Mp_ja", ip), ("xo_Mp_a", dp), ("xo_F", dp), ("xo_bc_idx", ip), ("xo_bc_val", dp),
                         ("xo_u_map", ip), ("xo_p_map", ip), ("xo_coeff_qp", dp)):
            getattr(L, name).argtypes = [vp]; getattr(L, name).restype = rt
        L.xo_A_mult.argtypes = [vp, dp, dp]
        L.xo_csr_mult.argtypes = [C.c_int, ip, ip, dp, dp, dp]
        L.xo_submatrix.argtypes = [vp, C.c_int, C.c_int, ip, ip, dp]; L.xo_submatrix.restype = C.c_int64
        L.xo_mg_setup.argtypes = [vp, C.c_int]
        L.xo_mg_level_csr.argtypes = [vp, C.c_int, ip, C.POINTER(ip), C.POINTER(ip), C.POINTER(dp)]
        L.xo_mg_prolong_add.argtypes = [vp, C.c_int, dp, dp]
        L.xo_mg_restrict.argtypes = [vp, C.c_int, dp, dp]
        L.xo_ilu0.argtypes = [C.c_int, ip, ip, dp, dp]
        L.xo_ilu0_solve.argtypes = [C.c_int, ip, ip, dp, dp, dp]
        L.xo_solve.argtypes = [vp, C.POINTER(Solver), dp, dp, C.POINTER(Result)]
        L.xo_pc_setup.argtypes = [vp, C.POINTER(Solver), C.POINTER(Result)]
        L.xo_pc_apply.argtypes = [vp, dp, dp, ip]
        L.xo_vcycle.argtypes = [vp, dp, dp]
        L.xo_diagnostics.argtypes = [vp, dp, dp]
        L.xo_rander48.argtypes = [C.c_int, C.c_int, dp]
        L.xo_hess_eig.argtypes = [C.c_int, dp, C.c_int, dp, dp]
        L.xo_num_threads.restype = C.c_int
        L.xo_set_num_threads.argtypes = [C.c_int]
        L.xo_set_sum_order.argtypes = [C.c_int]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def parse_options(argv, options_file_dir=None):
    """PETSc-style option list -> dict (key without '-' -> value string or True).
    Handles `-options_file f` (exSaddle2d/3d -options_file abf.opts) and '#' comments."""
    if isinstance(argv, str):
        argv = shlex.split(argv, comments=False)
    out = {}
    i = 0
    while i < len(argv):
        k = argv[i]
        if not k.startswith("-"):
            i += 1
            continue
        key = k[1:]
        val = True
        if i + 1 < len(argv) and not (argv[i + 1].startswith("-") and not _is_number(argv[i + 1])):
            val = argv[i + 1]
            i += 1
        i += 1
        if key == "options_file":
            path = val if os.path.isabs(val) else os.path.join(options_file_dir or os.getcwd(), val)
            with open(path) as f:
                for line in f:
                    line = line.split("#", 1)[0].strip()
                    if line:
                        for kk, vv in parse_options(line).items():
                            out.setdefault(kk, vv)
            continue
        out[key] = val
    return out


def _is_number(s):
    try:
        float(s.split(",")[0])
        return True
    except ValueError:
        return False


ABF_OPTS = """
-saddle_ksp_type fgmres
-fs
-saddle_fieldsplit_u_pc_type mg
-saddle_fieldsplit_u_ksp_type gcr
-saddle_fieldsplit_u_ksp_rtol 1e-2
-saddle_fieldsplit_u_pc_mg_levels 3
-saddle_fieldsplit_u_mg_levels_pc_type jacobi
-saddle_fieldsplit_u_mg_levels_ksp_type chebyshev
-saddle_fieldsplit_u_mg_levels_ksp_chebyshev_esteig 0,0.2,0,1.1
-saddle_fieldsplit_u_mg_levels_ksp_max_it 8
-saddle_fieldsplit_u_mg_levels_ksp_norm_type none
-saddle_fieldsplit_u_pc_mg_galerkin
-saddle_fieldsplit_p_ksp_type preonly
-saddle_fieldsplit_p_pc_type bjacobi
-saddle_fieldsplit_u_mg_coarse_pc_factor_mat_solver_type umfpack
"""  # content of /root/reference/abf.opts:2-16 (option names are the interface being mirrored)


def make_params(o, nsd=3, lame=False):
    p = Params()
    lib().xo_params_init(C.byref(p), nsd, int(lame))
    g = lambda k, d, t=float: t(o[k]) if k in o else d
    p.mx = g("mx", 4, int); p.my = g("my", -1, int); p.mz = g("mz", -1, int)
    p.size[0] = g("size_x", 1.0); p.size[1] = g("size_y", 1.0); p.size[2] = g("size_z", 1.0)
    p.model = g("model", -1, int)
    nan = float("nan")
    if lame:
        p.c0 = g("mu0", nan); p.c1 = g("mu1", nan); p.lam0 = g("lambda0", nan); p.lam1 = g("lambda1", nan)
    else:
        p.c0 = g("eta0", nan); p.c1 = g("eta1", nan)
    p.sinker_r = g("sinker_r", nan)
    p.sinker_c[0] = g("sinker_x", nan); p.sinker_c[1] = g("sinker_y", nan); p.sinker_c[2] = g("sinker_z", nan)
    p.sinker_n = g("sinker_n", -1, int); p.solcx_xc = g("solcx_xc", nan); p.solcx_nz = g("solcx_nz", -1, int)
    p.freeslip = 1 if o.get("freesliphack") in (True, "1", "true") else 0
    return p


def make_solver(o):
    s = Solver()
    L = lib()
    L.xo_solver_init(C.byref(s))
    fs = "fs" in o
    ksp = o.get("saddle_ksp_type", "gmres")
    s.ksp_type = {"gmres": 0, "fgmres": 1}[ksp]
    if fs:
        s.pc_type = 2
        if o.get("saddle_fieldsplit_u_pc_type") != "mg" or o.get("saddle_fieldsplit_u_ksp_type") != "gcr" \
                or o.get("saddle_fieldsplit_p_ksp_type") != "preonly":
            raise NotImplementedError("oracle supports -fs only with the abf.opts tree (gcr+mg / preonly)")
    else:
        s.pc_type = {"jacobi": 1, "none": 0}[o.get("saddle_pc_type", "none")]
    side = o.get("saddle_ksp_pc_side")
    s.pc_side = 1 if (side == "right" or ksp == "fgmres") else 0
    s.rtol = float(o.get("saddle_ksp_rtol", 1e-5)); s.atol = float(o.get("saddle_ksp_atol", 1e-50))
    s.max_it = int(o.get("saddle_ksp_max_it", 10000)); s.restart = int(o.get("saddle_ksp_gmres_restart", 30))
    s.u_rtol = float(o.get("saddle_fieldsplit_u_ksp_rtol", 1e-5)); s.u_max_it = int(o.get("saddle_fieldsplit_u_ksp_max_it", 10000))
    s.u_restart = int(o.get("saddle_fieldsplit_u_ksp_gcr_restart", 30))
    s.mg_levels = int(o.get("saddle_fieldsplit_u_pc_mg_levels", 1))
    s.cheb_its = int(o.get("saddle_fieldsplit_u_mg_levels_ksp_max_it", 2))
    if "saddle_fieldsplit_u_mg_levels_ksp_chebyshev_esteig" in o:
        v = [float(t) for t in o["saddle_fieldsplit_u_mg_levels_ksp_chebyshev_esteig"].split(",")]
        for i in range(4):
            s.esteig[i] = v[i]
    s.esteig_steps = int(o.get("saddle_fieldsplit_u_mg_levels_ksp_chebyshev_esteig_steps", 10))
    s.noise = int(o.get("xsb_chebyshev_noise", 0))
    # explicit per-level bounds: -saddle_fieldsplit_u_mg_levels_<l>_ksp_chebyshev_eigenvalues emin,emax (l = 1..levels-1)
    n = 0
    for l in range(1, s.mg_levels):
        key = "saddle_fieldsplit_u_mg_levels_%d_ksp_chebyshev_eigenvalues" % l
        if key in o:
            a, b = (float(t) for t in o[key].split(","))
            s.cheb_emin[l - 1] = a; s.cheb_emax[l - 1] = b; n += 1
    if n and n != s.mg_levels - 1:
        raise ValueError("explicit Chebyshev eigenvalues must be given for every level")
    s.n_cheb_fixed = n
    ppc = o.get("saddle_fieldsplit_p_pc_type", "bjacobi")
    s.p_pc = {"bjacobi": 0, "ilu": 0, "jacobi": 1}[ppc]
    s.p_blocks = int(o.get("xo_p_blocks", 1))   # bjacobi blocks = ranks of the slab partition being mirrored
    return s


class CSR:
    def __init__(self, ia, ja, a, shape):
        self.ia, self.ja, self.a, self.shape = ia, ja, a, shape

    def scipy(self):
        import scipy.sparse as sp
        return sp.csr_matrix((self.a, self.ja, self.ia), shape=self.shape)


class Problem:
    """One exSaddle run: Problem('-model 0 -mx 4 ...', nsd=2)."""

    def __init__(self, opts, nsd=3, lame=False, options_file_dir=None, nodal=None):
        self.L = lib()
        self.o = parse_options(opts, options_file_dir) if not isinstance(opts, dict) else dict(opts)
        self.nsd, self.lame = nsd, lame
        self.params = make_params(self.o, nsd, lame)
        self.h = C.c_void_p()
        if nodal is None:
            rc = self.L.xo_create(C.byref(self.params), C.byref(self.h))
        else:   # coarse level of the -mg hierarchy: coefficients interpolated from nodal Q1 fields (npn x 6)
            nodal = np.ascontiguousarray(nodal, dtype=np.float64)
            rc = self.L.xo_create_nodal(C.byref(self.params), _dp(nodal), C.byref(self.h))
        if rc:
            msg = self.L.xo_error(self.h).decode()
            self.L.xo_destroy(self.h); self.h = None
            raise RuntimeError("oracle: " + msg)
        sz = (C.c_int64 * 8)()
        self.L.xo_sizes(self.h, sz)
        (self.n, self.nu, self.np_, self.nnz, self.prealloc, self.nel, self.nbc, self.mnnz) = [int(v) for v in sz]
        self.banner = self.L.xo_banner(self.h).decode()

    def coeff_nodal(self):
        """nodal Q1 coefficient fields, (p nodes, 6 slots: eta|mu, Fu0, Fu1, Fu2, Fp, lambda)"""
        return self._arr(self.L.xo_coeff_nodal(self.h), self.np_ * 6, np.float64).reshape(self.np_, 6).copy()

    def __del__(self):
        if getattr(self, "h", None):
            self.L.xo_destroy(self.h); self.h = None

    def _arr(self, ptr, n, dtype):
        return np.ctypeslib.as_array(ptr, shape=(n,)).view(dtype)

    def A(self):
        return CSR(self._arr(self.L.xo_A_ia(self.h), self.n + 1, np.int32), self._arr(self.L.xo_A_ja(self.h), self.nnz, np.int32),
                   self._arr(self.L.xo_A_a(self.h), self.nnz, np.float64), (self.n, self.n))

    def A_raw_values(self):
        p = self.L.xo_A_raw(self.h)
        return self._arr(p, self.nnz, np.float64) if p else None

    def Mp(self):
        return CSR(self._arr(self.L.xo_Mp_ia(self.h), self.np_ + 1, np.int32), self._arr(self.L.xo_Mp_ja(self.h), self.mnnz, np.int32),
                   self._arr(self.L.xo_Mp_a(self.h), self.mnnz, np.float64), (self.np_, self.np_))

    def F(self):
        return self._arr(self.L.xo_F(self.h), self.n, np.float64)

    def bc(self):
        return self._arr(self.L.xo_bc_idx(self.h), self.nbc, np.int32), self._arr(self.L.xo_bc_val(self.h), self.nbc, np.float64)

    def coeff_qp(self):
        nqp = 27 if self.nsd == 3 else 9
        return self._arr(self.L.xo_coeff_qp(self.h), self.nel * nqp * 6, np.float64).reshape(self.nel, nqp, 6)

    def mult(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64); y = np.empty(self.n)
        self.L.xo_A_mult(self.h, _dp(x), _dp(y))
        return y

    def submatrix(self, rb, cb):
        nr = self.np_ if rb else self.nu
        nc = self.np_ if cb else self.nu
        ia = np.empty(nr + 1, np.int32)
        nnz = self.L.xo_submatrix(self.h, rb, cb, _ip(ia), None, None)
        ja = np.empty(max(nnz, 1), np.int32); a = np.empty(max(nnz, 1))
        self.L.xo_submatrix(self.h, rb, cb, _ip(ia), _ip(ja), _dp(a))
        return CSR(ia, ja[:nnz], a[:nnz], (nr, nc))

    def solver(self):
        return make_solver(self.o)

    def pc_setup(self, s=None):
        s = s or self.solver()
        self._s = s
        r = Result()
        if self.L.xo_pc_setup(self.h, C.byref(s), C.byref(r)):
            raise RuntimeError("oracle: " + self.L.xo_error(self.h).decode())
        return r

    def mg_level(self, l):
        n = C.c_int(); ia = C.POINTER(C.c_int)(); ja = C.POINTER(C.c_int)(); a = C.POINTER(C.c_double)()
        if self.L.xo_mg_level_csr(self.h, l, C.byref(n), C.byref(ia), C.byref(ja), C.byref(a)):
            raise IndexError(l)
        n = n.value
        iaa = self._arr(ia, n + 1, np.int32)
        nnz = int(iaa[n])
        return CSR(iaa, self._arr(ja, nnz, np.int32), self._arr(a, nnz, np.float64), (n, n))

    def vcycle(self, b):
        b = np.ascontiguousarray(b, dtype=np.float64); x = np.empty(self.nu)
        self.L.xo_vcycle(self.h, _dp(b), _dp(x))
        return x

    def pc_apply(self, r):
        r = np.ascontiguousarray(r, dtype=np.float64); z = np.zeros(self.n); its = C.c_int()
        self.L.xo_pc_apply(self.h, _dp(r), _dp(z), C.byref(its))
        return z, its.value

    def prolong_add(self, lc, xc, xf):
        self.L.xo_mg_prolong_add(self.h, lc, _dp(xc), _dp(xf)); return xf

    def restrict(self, lc, rf, nc):
        bc = np.zeros(nc); self.L.xo_mg_restrict(self.h, lc, _dp(rf), _dp(bc)); return bc

    def solve(self, s=None, b=None):
        s = s or self.solver()
        r = Result(); x = np.zeros(self.n)
        bp = _dp(np.ascontiguousarray(b, dtype=np.float64)) if b is not None else None
        if self.L.xo_solve(self.h, C.byref(s), bp, _dp(x), C.byref(r)):
            raise RuntimeError("oracle: " + self.L.xo_error(self.h).decode())
        return x, r

    def diagnostics(self, x):
        out = np.zeros(5 * self.nsd + 5)
        self.L.xo_diagnostics(self.h, _dp(np.ascontiguousarray(x)), _dp(out))
        return out

    def diagnostics_text(self, x):
        """The 10 lines SaddleReportSolutionDiagnostics prints (exSaddle_io.c:17-56)."""
        d = self.diagnostics(x); n = self.nsd
        tag = "|u,v|" if n == 2 else "|u,v,w|"
        lines = []
        for k, nm in enumerate(("_1  ", "_2  ", "_inf", "_min", "_max")):
            vals = " , ".join("%+1.6e" % d[k * n + c] for c in range(n))
            lines.append("%s%s %s%s" % (tag, nm, vals, " " if n == 2 else ""))
        for k, nm in enumerate(("_1  ", "_2  ", "_inf", "_min", "_max")):
            lines.append("|p|%s        %+1.6e" % (nm, d[5 * n + k]))
        return lines


def monitor_short(v):
    """-ksp_monitor_short number formatting (KSPMonitorDefaultShort): %g above 1e-9, %5.3e below."""
    if v > 1.e-9:
        return "%g" % v
    if v > 1.e-11:
        return "%5.3e" % v
    return "< 1.e-11"


def rander48(n, interval=0):
    v = np.empty(n); lib().xo_rander48(n, interval, _dp(v)); return v


def hess_eig(H):
    H = np.ascontiguousarray(H, dtype=np.float64); n = H.shape[0]
    re = np.empty(n); im = np.empty(n)
    if lib().xo_hess_eig(n, _dp(H), n, _dp(re), _dp(im)):
        raise RuntimeError("hqr failed")
    return re + 1j * im
