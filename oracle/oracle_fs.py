"""CPU oracle of the reference's plain `-fs` tree (exSaddle.c:303-322 with PETSc's default sub-solvers; goldens *_fs_1).
TEST INFRASTRUCTURE, NOT PRODUCT.  Oracle only: the GPU library implements -fs with the abf.opts tree (GCR + GMG / preonly), not
this one (ILU(0) of the assembled A00 in natural ordering is a ~5000-wavefront sequential sweep: not a GPU algorithm).

  outer   GMRES(30), left PC, preconditioned norm, rtol 1e-5
  PC      PCFIELDSPLIT Schur / UPPER / user Mpscaled (App. B.2):
            y_p = KSP_p[S, PC(Mpscaled)] x_p,   S v = A11 v - A10 KSP_u[A00](A01 v)
            y_u = KSP_u[A00] (x_u - A01 y_p)
  KSP_u   GMRES(30) + ILU(0) (PETSc's default PC for seqaij), rtol 1e-5   (-saddle_fieldsplit_u_ksp_max_it N honoured)
  KSP_p   GMRES(30) + ILU(0) of Mpscaled, rtol 1e-5                       (-saddle_fieldsplit_p_ksp_type preonly honoured)
"""
import numpy as np

from . import oracle as O
from .oracle_mg import petsc_gmres


class Ilu0:
    def __init__(self, A):
        A = A.tocsr(); A.sort_indices()
        self.n = A.shape[0]; self.ia = A.indptr.astype(np.int32); self.ja = A.indices.astype(np.int32)
        self.lu = np.empty(len(A.data))
        a = np.ascontiguousarray(A.data, np.float64)
        if O.lib().xo_ilu0(self.n, O._ip(self.ia), O._ip(self.ja), O._dp(a), O._dp(self.lu)):
            raise RuntimeError("zero pivot in ILU(0)")

    def __call__(self, b):
        b = np.ascontiguousarray(b, np.float64); x = np.empty(self.n)
        O.lib().xo_ilu0_solve(self.n, O._ip(self.ia), O._ip(self.ja), O._dp(self.lu), O._dp(b), O._dp(x))
        return x


class FieldSplitDefault:
    def __init__(self, opts, nsd=3, lame=False):
        self.o = O.parse_options(opts) if not isinstance(opts, dict) else dict(opts)
        o = self.o
        if "fs" not in o or "mg" in o:
            raise ValueError("FieldSplitDefault needs -fs (and no -mg)")
        self.p = O.Problem(o, nsd=nsd, lame=lame)
        p = self.p; nu = p.nu; self.nu = nu
        A = p.A().scipy().tocsr(); self.A = A
        self.A00 = A[:nu, :nu].tocsr(); self.A01 = A[:nu, nu:].tocsr(); self.A10 = A[nu:, :nu].tocsr(); self.A11 = A[nu:, nu:].tocsr()
        self.ilu_u = Ilu0(self.A00); self.ilu_p = Ilu0(p.Mp().scipy())
        self.u_max_it = int(o.get("saddle_fieldsplit_u_ksp_max_it", 10000))
        self.p_preonly = o.get("saddle_fieldsplit_p_ksp_type", "gmres") == "preonly"
        self.u_its = []

    def ksp_u(self, rhs):
        x, its, _, _ = petsc_gmres(lambda v: self.A00 @ v, self.ilu_u, rhs, max_it=self.u_max_it)
        self.u_its.append(its)
        return x

    def pc_apply(self, r):
        nu = self.nu
        if self.p_preonly:
            yp = self.ilu_p(r[nu:])
        else:
            S = lambda v: self.A11 @ v - self.A10 @ self.ksp_u(self.A01 @ v)
            yp = petsc_gmres(S, self.ilu_p, r[nu:])[0]
        yu = self.ksp_u(r[:nu] - self.A01 @ yp)
        return np.concatenate([yu, yp])

    def solve(self):
        o = self.o
        return petsc_gmres(lambda v: self.A @ v, self.pc_apply, self.p.F(), rtol=float(o.get("saddle_ksp_rtol", 1e-5)),
                           max_it=int(o.get("saddle_ksp_max_it", 10000)), restart=int(o.get("saddle_ksp_gmres_restart", 30)))
