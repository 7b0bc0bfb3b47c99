"""ctypes binding of include/exsaddle_b200.h (one Python method per C entry point)."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBPATH = os.path.join(_HERE, "libexsaddle_b200.so")

MAT_A, MAT_A00, MAT_A01, MAT_A10, MAT_A11, MAT_MP, MAT_A00_MF, MAT_A01_MF, MAT_A10_MF, MAT_MG_LEVEL0 = 0, 1, 2, 3, 4, 5, 6, 7, 8, 16
ERR_NO_DEVICE = -6


class XsbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("exsaddle_b200 error %d: %s" % (code, msg))
        self.code = code


def library_path():
    return _LIBPATH


_lib = None


def lib():
    """Load the CUDA library. It must have been built (python -m exsaddle_b200.build); there is no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIBPATH):
            raise ImportError("%s is missing: build it with `python exsaddle_b200/build.py` (nvcc, sm_100a). "
                              "exsaddle_b200 has no CPU or PyTorch fallback." % _LIBPATH)
        L = C.CDLL(_LIBPATH)
        vp, i32p, i64p, dp = C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_double)
        sig = {
            "xsb_create": [C.POINTER(vp), C.c_int, C.c_int, C.c_int], "xsb_reset": [vp], "xsb_destroy": [C.POINTER(vp)],
            "xsb_device_available": [], "xsb_set_option": [vp, C.c_char_p, C.c_char_p], "xsb_set_options": [vp, C.c_char_p],
            "xsb_set_options_file": [vp, C.c_char_p], "xsb_options_left": [vp, C.c_char_p, C.c_int],
            "xsb_assemble": [vp], "xsb_banner": [vp, C.c_char_p, C.c_int], "xsb_get_sizes": [vp, i64p],
            "xsb_mat_get_info": [vp, C.c_int, i64p, i64p, i64p, C.POINTER(C.c_int)],
            "xsb_mat_get_csr": [vp, C.c_int, i32p, i32p, dp], "xsb_mat_mult": [vp, C.c_int, dp, dp],
            "xsb_mat_mult_dev": [vp, C.c_int, vp, vp], "xsb_mat_get_diagonal": [vp, C.c_int, dp],
            "xsb_mat_mult_transpose": [vp, C.c_int, dp, dp], "xsb_ksp_view": [vp, C.c_char_p, C.c_int],
            "xsb_dump_operator": [vp, C.c_int, C.c_char_p], "xsb_dump_vector": [vp, dp, C.c_int64, C.c_char_p],
            "xsb_write_petsc_mat": [C.c_char_p, C.c_int64, C.c_int64, i32p, i32p, dp], "xsb_write_petsc_vec": [C.c_char_p, C.c_int64, dp],
            "xsb_view_fields": [vp, dp, C.c_char_p, C.c_char_p],
            "xsb_write_vts": [C.c_char_p, C.c_int, C.c_int, C.c_int, dp, C.c_int, C.POINTER(C.c_char_p), dp, C.c_int64, C.c_int64],
            "xsb_vec_get_rhs": [vp, dp], "xsb_get_bc": [vp, i32p, dp], "xsb_get_coeff_qp": [vp, C.c_int, dp],
            "xsb_ksp_setup": [vp], "xsb_ksp_reset": [vp], "xsb_get_state": [vp, C.POINTER(C.c_int), C.POINTER(C.c_int)], "xsb_ksp_solve": [vp, dp, dp], "xsb_ksp_solve_dev": [vp, vp, vp],
            "xsb_pc_apply": [vp, dp, dp], "xsb_pc_apply_dev": [vp, vp, vp], "xsb_pc_mg_apply": [vp, dp, dp],
            "xsb_pc_schur_apply": [vp, dp, dp], "xsb_time_pc_schur": [vp, C.c_int, dp], "xsb_time_halo": [vp, C.c_int, dp], "xsb_mg_restrict": [vp, C.c_int, dp, dp],
            "xsb_mg_interpolate_add": [vp, C.c_int, dp, dp],
            "xsb_ksp_get_iterations": [vp, C.POINTER(C.c_int), C.POINTER(C.c_int)],
            "xsb_ksp_get_history": [vp, dp, C.c_int, C.POINTER(C.c_int)],
            "xsb_ksp_get_inner_iterations": [vp, C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_int)],
            "xsb_ksp_get_inner_reasons": [vp, C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_int)],
            "xsb_ksp_get_chebyshev": [vp, C.c_int, dp, dp, dp, dp], "xsb_ksp_get_timing": [vp, dp, dp],
            "xsb_ksp_get_counters": [vp, i64p], "xsb_diagnostics": [vp, dp, dp], "xsb_get_stream": [vp, C.POINTER(vp)],
            "xsb_ksp_get_profile": [vp, dp, i64p, C.c_int, C.POINTER(C.c_int)], "xsb_comm_info": [vp, i64p],
            "xsb_pattern_row": [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64, i32p, C.c_int],
            "xsb_prealloc_total": [C.c_int, C.c_int, C.c_int, C.c_int],
            "xsb_bc_list": [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, i32p, dp, C.c_int],
            "xsb_mg_level_dims": [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)],
            "xsb_grad_line_tables": [C.c_int, C.c_double, i32p, dp, dp, dp, dp],
            "xsb_dmda_grid": [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)],
            "xsb_asm_subdomain": [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)],
            "xsb_slab_range": [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)],
            "xsb_pdist_range": [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)],
            "xsb_slab_layout": [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, i64p],
            "xsb_comm_unique_id": [C.c_char_p], "xsb_comm_init": [vp, C.c_char_p, C.c_int, C.c_int], "xsb_get_partition": [vp, i64p],
        }
        for name, args in sig.items():
            f = getattr(L, name); f.argtypes = args; f.restype = C.c_int
        L.xsb_last_error.argtypes = [vp]; L.xsb_last_error.restype = C.c_char_p
        L.xsb_prealloc_total.restype = C.c_int64
        L._symbols = list(sig) + ["xsb_last_error"]
        _lib = L
    return _lib


def device_available():
    return bool(lib().xsb_device_available())


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


# ---- host-side index maps (no GPU) ------------------------------------------------------------------
def pattern_row(nsd, mx, my, mz, row):
    n = lib().xsb_pattern_row(nsd, mx, my, mz, row, None, 0)
    if n < 0:
        raise XsbError(n, "bad arguments")
    cols = np.empty(n, np.int32)
    lib().xsb_pattern_row(nsd, mx, my, mz, row, _ip(cols), n)
    return cols


def prealloc_total(nsd, mx, my, mz):
    return int(lib().xsb_prealloc_total(nsd, mx, my, mz))


def bc_list(nsd, lame, model, freeslip, mx, my, mz):
    n = lib().xsb_bc_list(nsd, int(lame), model, int(freeslip), mx, my, mz, None, None, 0)
    idx = np.empty(max(n, 1), np.int32); val = np.empty(max(n, 1))
    lib().xsb_bc_list(nsd, int(lame), model, int(freeslip), mx, my, mz, _ip(idx), _dp(val), n)
    return idx[:n], val[:n]


def mg_level_dims(nsd, mx, my, mz, levels, level):
    d = (C.c_int * 3)()
    rc = lib().xsb_mg_level_dims(nsd, mx, my, mz, levels, level, d)
    if rc:
        raise XsbError(rc, "mesh cannot be coarsened to %d levels" % levels)
    return tuple(d)


def slab_range(mz, nranks, rank):
    a, b = C.c_int(), C.c_int()
    rc = lib().xsb_slab_range(mz, nranks, rank, C.byref(a), C.byref(b))
    if rc:
        raise XsbError(rc, "bad slab request")
    return a.value, b.value


def grad_line_tables(m, h):
    """per-direction coefficient tables of the matrix-free gradient / divergence blocks (xsb_grad_line_tables)"""
    import numpy as np
    uP = np.empty(3 * (2 * m + 1), np.int32); uM = np.empty(3 * (2 * m + 1)); uG = np.empty_like(uM); pM = np.empty(5 * (m + 1)); pG = np.empty_like(pM)
    rc = lib().xsb_grad_line_tables(m, h, uP.ctypes.data_as(C.POINTER(C.c_int32)), _dp(uM), _dp(uG), _dp(pM), _dp(pG))
    if rc:
        raise XsbError(rc, "bad table request")
    return uP.reshape(-1, 3), uM.reshape(-1, 3), uG.reshape(-1, 3), pM.reshape(-1, 5), pG.reshape(-1, 5)


def dmda_grid(nsd, M, N, P, size):
    """process grid of DMDACreate{2,3}d with PETSC_DECIDE (xsb_dmda_grid)"""
    out = (C.c_int * 3)()
    rc = lib().xsb_dmda_grid(nsd, M, N, P, size, out)
    if rc:
        raise XsbError(rc, "no process grid")
    return tuple(out)


def asm_subdomain(nsd, mx, my, mz, size, overlap, rank):
    """element patch and owned node ranges of one rank's ASM subdomain (xsb_asm_subdomain)"""
    out = (C.c_int * 18)()
    rc = lib().xsb_asm_subdomain(nsd, mx, my, mz, size, overlap, rank, out)
    if rc:
        raise XsbError(rc, "Cannot generate consistent macro element")
    v = list(out)
    return {"lo": v[0:nsd], "hi": v[3:3 + nsd], "own_u": [(v[6 + d], v[9 + d]) for d in range(nsd)], "own_p": [(v[12 + d], v[15 + d]) for d in range(nsd)]}


def pdist_range(mz, nranks, rank, depth):
    """Node planes of coarse MG level `depth` (0 = first coarse level) that `rank` computes (xsb_pdist_range)."""
    a, b = C.c_int(), C.c_int()
    rc = lib().xsb_pdist_range(mz, nranks, rank, depth, C.byref(a), C.byref(b))
    if rc:
        raise XsbError(rc, "bad plane-range request")
    return a.value, b.value


_PART_KEYS = ("rank", "nranks", "k0", "k1", "e0", "e1", "u_off", "u_len", "p_off", "p_len", "u_glob0", "p_glob0")


def slab_layout(nsd, mx, my, mz, nranks, rank):
    out = (C.c_int64 * 12)()
    rc = lib().xsb_slab_layout(nsd, mx, my, mz, nranks, rank, out)
    if rc:
        raise XsbError(rc, "bad slab layout request")
    return dict(zip(_PART_KEYS, [int(v) for v in out]))


def comm_unique_id():
    buf = C.create_string_buffer(128)
    rc = lib().xsb_comm_unique_id(buf)
    if rc:
        raise XsbError(rc, "ncclGetUniqueId failed (is libnccl.so.2 loadable?)")
    return buf.raw


class ExSaddle:
    """One exSaddle{2d,3d}{,_lame} run on the GPU. `opts` is the reference's own option string."""

    def __init__(self, opts="", nsd=3, lame=False, device=-1):
        self.L = lib()
        self.nsd, self.lame = nsd, bool(lame)
        self.h = C.c_void_p()
        rc = self.L.xsb_create(C.byref(self.h), nsd, int(lame), device)
        if rc:
            raise XsbError(rc, "xsb_create failed")
        if opts:
            self.set_options(opts)

    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self.L.xsb_destroy(C.byref(self.h)); self.h = None

    __del__ = close

    def _chk(self, rc):
        if rc:
            raise XsbError(rc, self.L.xsb_last_error(self.h).decode())

    # options ------------------------------------------------------------------------------------------
    def set_options(self, s):
        self._chk(self.L.xsb_set_options(self.h, s.encode()))

    def set_option(self, key, value=None):
        self._chk(self.L.xsb_set_option(self.h, key.encode(), None if value is None else str(value).encode()))

    def set_options_file(self, path):
        self._chk(self.L.xsb_set_options_file(self.h, path.encode()))

    def options_left(self):
        buf = C.create_string_buffer(1 << 16)
        self._chk(self.L.xsb_options_left(self.h, buf, len(buf)))
        return [l for l in buf.value.decode().split("\n") if l]

    # multi-GPU ----------------------------------------------------------------------------------------
    def comm_init(self, unique_id, rank, nranks):
        self._chk(self.L.xsb_comm_init(self.h, unique_id, rank, nranks)); return self

    def partition(self):
        out = (C.c_int64 * 12)()
        self._chk(self.L.xsb_get_partition(self.h, out))
        return dict(zip(_PART_KEYS, [int(v) for v in out]))

    # FE set-up ----------------------------------------------------------------------------------------
    def assemble(self):
        self._chk(self.L.xsb_assemble(self.h))
        sz = (C.c_int64 * 8)()
        self._chk(self.L.xsb_get_sizes(self.h, sz))
        (self.n, self.nu, self.np_, self.nnz, self.prealloc, self.nel, self.nbc, self.mnnz) = [int(v) for v in sz]
        return self

    def banner(self):
        buf = C.create_string_buffer(4096)
        self._chk(self.L.xsb_banner(self.h, buf, len(buf)))
        return buf.value.decode()

    def mat_info(self, which):
        r, c_, z, bs = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int()
        self._chk(self.L.xsb_mat_get_info(self.h, which, C.byref(r), C.byref(c_), C.byref(z), C.byref(bs)))
        return r.value, c_.value, z.value, bs.value

    def mat_csr(self, which, values=True, pattern=True):
        r, c_, z, _ = self.mat_info(which)
        ia = np.empty(r + 1, np.int32); ja = np.empty(z, np.int32) if pattern else None
        a = np.empty(z) if values else None
        self._chk(self.L.xsb_mat_get_csr(self.h, which, _ip(ia), _ip(ja) if pattern else None, _dp(a) if values else None))
        return ia, ja, a, (r, c_)

    def mat_mult(self, which, x):
        r, c_, _, _ = self.mat_info(which)
        x = np.ascontiguousarray(x, dtype=np.float64)
        assert x.shape == (c_,)
        y = np.empty(r)
        self._chk(self.L.xsb_mat_mult(self.h, which, _dp(x), _dp(y)))
        return y

    def mat_mult_dev(self, which, x_ptr, y_ptr):
        self._chk(self.L.xsb_mat_mult_dev(self.h, which, C.c_void_p(x_ptr), C.c_void_p(y_ptr)))

    def mat_diagonal(self, which):
        r = self.mat_info(which)[0]; d = np.empty(r)
        self._chk(self.L.xsb_mat_get_diagonal(self.h, which, _dp(d))); return d

    def rhs(self):
        F = np.empty(self.n); self._chk(self.L.xsb_vec_get_rhs(self.h, _dp(F))); return F

    def bc(self):
        idx = np.empty(max(self.nbc, 1), np.int32); val = np.empty(max(self.nbc, 1))
        self._chk(self.L.xsb_get_bc(self.h, _ip(idx), _dp(val)))
        return idx[:self.nbc], val[:self.nbc]

    def coeff_qp(self, slot):
        out = np.empty(self.nel * (27 if self.nsd == 3 else 9))
        self._chk(self.L.xsb_get_coeff_qp(self.h, slot, _dp(out))); return out

    # solver -------------------------------------------------------------------------------------------
    def ksp_setup(self):
        self._chk(self.L.xsb_ksp_setup(self.h)); return self

    def solve(self, b=None):
        x = np.empty(self.n)
        bp = None
        if b is not None:
            b = np.ascontiguousarray(b, dtype=np.float64); bp = _dp(b)
        self._chk(self.L.xsb_ksp_solve(self.h, bp, _dp(x)))
        return x

    def solve_dev(self, b_ptr, x_ptr):
        self._chk(self.L.xsb_ksp_solve_dev(self.h, C.c_void_p(b_ptr) if b_ptr else None, C.c_void_p(x_ptr)))

    def pc_apply(self, r):
        r = np.ascontiguousarray(r, dtype=np.float64); z = np.empty(self.n)
        self._chk(self.L.xsb_pc_apply(self.h, _dp(r), _dp(z))); return z

    def pc_mg_apply(self, b):
        b = np.ascontiguousarray(b, dtype=np.float64); x = np.empty(self.nu)
        self._chk(self.L.xsb_pc_mg_apply(self.h, _dp(b), _dp(x))); return x

    def time_pc_schur(self, reps=20):
        ms = C.c_double()
        self._chk(self.L.xsb_time_pc_schur(self.h, reps, C.byref(ms))); return ms.value

    def time_halo(self, reps=50):
        out = (C.c_double * 3)()
        self._chk(self.L.xsb_time_halo(self.h, reps, out)); return {"halo_u_us": out[0], "halo_plane_us": out[1], "a00_kernel_us": out[2]}

    def pc_schur_apply(self, b):
        b = np.ascontiguousarray(b, dtype=np.float64); x = np.empty(self.np_)
        self._chk(self.L.xsb_pc_schur_apply(self.h, _dp(b), _dp(x))); return x

    def mg_restrict(self, lc, rf):
        nc = self.mat_info(MAT_MG_LEVEL0 + lc)[0]
        rf = np.ascontiguousarray(rf, dtype=np.float64); bc = np.empty(nc)
        self._chk(self.L.xsb_mg_restrict(self.h, lc, _dp(rf), _dp(bc))); return bc

    def mg_interpolate_add(self, lc, xc, xf):
        xc = np.ascontiguousarray(xc, dtype=np.float64); xf = np.array(xf, dtype=np.float64)
        self._chk(self.L.xsb_mg_interpolate_add(self.h, lc, _dp(xc), _dp(xf))); return xf

    def iterations(self):
        its, reason = C.c_int(), C.c_int()
        self._chk(self.L.xsb_ksp_get_iterations(self.h, C.byref(its), C.byref(reason)))
        return its.value, reason.value

    def history(self):
        n = C.c_int(); h = np.empty(20000)
        self._chk(self.L.xsb_ksp_get_history(self.h, _dp(h), len(h), C.byref(n)))
        return h[:n.value].copy()

    def inner_iterations(self):
        n = C.c_int(); a = (C.c_int * 20000)()
        self._chk(self.L.xsb_ksp_get_inner_iterations(self.h, a, 20000, C.byref(n)))
        return [a[i] for i in range(n.value)]

    def inner_reasons(self):
        n = C.c_int(); a = (C.c_int * 20000)()
        self._chk(self.L.xsb_ksp_get_inner_reasons(self.h, a, 20000, C.byref(n)))
        return [a[i] for i in range(n.value)]

    def chebyshev(self, level):
        v = [C.c_double() for _ in range(4)]
        self._chk(self.L.xsb_ksp_get_chebyshev(self.h, level, *[C.byref(t) for t in v]))
        return tuple(t.value for t in v)

    def view(self):
        buf = C.create_string_buffer(16384)
        self._chk(self.L.xsb_ksp_view(self.h, buf, len(buf))); return buf.value.decode()

    def mat_mult_transpose(self, which, x):
        rows, cols, _, _ = self.mat_info(which)
        x = np.ascontiguousarray(x, dtype=np.float64); y = np.empty(cols)
        self._chk(self.L.xsb_mat_mult_transpose(self.h, which, _dp(x), _dp(y))); return y

    def view_fields(self, x, outdir=".", tag=""):
        x = np.ascontiguousarray(x, dtype=np.float64)
        self._chk(self.L.xsb_view_fields(self.h, _dp(x), outdir.encode(), tag.encode()))

    def dump_operator(self, which, path):
        self._chk(self.L.xsb_dump_operator(self.h, which, path.encode()))

    def dump_vector(self, x, path):
        x = np.ascontiguousarray(x, dtype=np.float64)
        self._chk(self.L.xsb_dump_vector(self.h, _dp(x), len(x), path.encode()))

    def timing(self):
        a, b = C.c_double(), C.c_double()
        self._chk(self.L.xsb_ksp_get_timing(self.h, C.byref(a), C.byref(b))); return a.value, b.value

    def counters(self):
        out = (C.c_int64 * 8)()
        self._chk(self.L.xsb_ksp_get_counters(self.h, out))
        return {"a00_spmv": out[0], "a_spmv": out[1], "launches": out[2], "a00_avg_ns": out[3],
                "a00_by_mode": [out[4], out[5], out[6], out[7]]}

    def profile(self):
        """-xsb_time_kernels: {category: (ms, stretches)} of the last solve (xsb_ksp_get_profile)."""
        ms = (C.c_double * 32)(); cnt = (C.c_int64 * 32)(); n = C.c_int()
        self._chk(self.L.xsb_ksp_get_profile(self.h, ms, cnt, 32, C.byref(n)))
        names = ["krylov_vectors_and_gaps", "fine_halo", "fine_a00"] + ["mg_level_%d" % l for l in range(10)] + ["coarse_plane_exchange", "transfers", "pressure_ilu", "full_operator", "fieldsplit_a01"]
        return {names[i]: (ms[i], cnt[i]) for i in range(min(n.value, len(names))) if cnt[i]}

    def comm_info(self):
        out = (C.c_int64 * 4)()
        self._chk(self.L.xsb_comm_info(self.h, out))
        return {"rank": out[0], "nranks": out[1], "p2p": bool(out[2]), "pdist_levels": [l for l in range(16) if out[3] >> l & 1]}

    def stream(self):
        p = C.c_void_p()
        self._chk(self.L.xsb_get_stream(self.h, C.byref(p)))
        return p.value or 0

    def diagnostics(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64); out = np.empty(5 * self.nsd + 5)
        self._chk(self.L.xsb_diagnostics(self.h, _dp(x), _dp(out))); return out


def write_petsc_mat(path, ia, ja, a, shape):
    """PETSc binary Mat file from host CSR arrays (no GPU needed)."""
    ia = np.ascontiguousarray(ia, np.int32); ja = np.ascontiguousarray(ja, np.int32); a = np.ascontiguousarray(a, np.float64)
    rc = lib().xsb_write_petsc_mat(path.encode(), shape[0], shape[1], _ip(ia), _ip(ja), _dp(a))
    if rc:
        raise XsbError(rc, "cannot write %s" % path)


def write_petsc_vec(path, x):
    x = np.ascontiguousarray(x, np.float64)
    rc = lib().xsb_write_petsc_vec(path.encode(), len(x), _dp(x))
    if rc:
        raise XsbError(rc, "cannot write %s" % path)


def read_petsc_binary(path):
    """Reader for the files above (the layout PETSc's PetscBinaryIO.py / PetscBinaryRead.m read): returns
    ("Mat", (ia, ja, a, shape)) or ("Vec", x)."""
    raw = open(path, "rb").read()
    cid = int(np.frombuffer(raw, ">i4", 1, 0)[0])
    if cid == 1211214:
        n = int(np.frombuffer(raw, ">i4", 1, 4)[0])
        return "Vec", np.frombuffer(raw, ">f8", n, 8).astype(np.float64)
    if cid == 1211216:
        rows, cols, nnz = (int(v) for v in np.frombuffer(raw, ">i4", 3, 4))
        lens = np.frombuffer(raw, ">i4", rows, 16).astype(np.int64)
        ja = np.frombuffer(raw, ">i4", nnz, 16 + 4 * rows).astype(np.int32)
        a = np.frombuffer(raw, ">f8", nnz, 16 + 4 * rows + 4 * nnz).astype(np.float64)
        ia = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
        assert len(raw) == 16 + 4 * rows + 12 * nnz
        return "Mat", (ia, ja, a, (rows, cols))
    raise ValueError("not a PETSc Mat/Vec binary file (class id %d)" % cid)


def write_vts(path, dims, h, names, data, field_stride, node_stride):
    """VTK XML StructuredGrid with raw appended data from a host array (no GPU needed); field f of node n = data[f*field_stride + n*node_stride]."""
    data = np.ascontiguousarray(data, np.float64); hh = np.ascontiguousarray(h, np.float64)
    arr = (C.c_char_p * len(names))(*[n.encode() for n in names])
    rc = lib().xsb_write_vts(path.encode(), dims[0], dims[1], dims[2], _dp(hh), len(names), arr, _dp(data), field_stride, node_stride)
    if rc:
        raise XsbError(rc, "cannot write %s" % path)


def read_vts(path):
    """Minimal reader of the files above: returns (dims, points[n,3], {name: values[n]})."""
    import re
    raw = open(path, "rb").read()
    head, _, tail = raw.partition(b"<AppendedData encoding=\"raw\">\n_")
    text = head.decode()
    ext = [int(v) for v in re.search(r'WholeExtent="([^"]+)"', text).group(1).split()]
    dims = (ext[1] + 1, ext[3] + 1, ext[5] + 1); n = dims[0] * dims[1] * dims[2]
    out = {}; pts = None
    for m in re.finditer(r'<DataArray type="Float64" Name="([^"]*)" NumberOfComponents="(\d)" format="appended" offset="(\d+)"', text):
        name, nc, off = m.group(1), int(m.group(2)), int(m.group(3))
        nbytes = int(np.frombuffer(tail, "<u8", 1, off)[0])
        assert nbytes == 8 * nc * n
        v = np.frombuffer(tail, "<f8", nc * n, off + 8).copy()
        if name == "Position":
            pts = v.reshape(n, 3)
        else:
            out[name] = v
    return dims, pts, out
