"""CPU oracle of the reference's plain `-fs` tree (exSaddle.c:303-322 with PETSc's default sub-solvers; goldens *_fs_1, *_fs_2).
TEST INFRASTRUCTURE, NOT PRODUCT (the product's version of this tree is exsaddle_b200/csrc/xsb_fs.cu).

On more than one rank (goldens *_fs_2, `mpiexec -n 2`) PETSc's default PC of the two inner solves becomes bjacobi with one
ILU(0) block per rank: the rank's rows / columns of A00 and of Mpscaled, i.e. the dofs its DMDAs own (velocity: default node
split of the PETSC_DECIDE grid; pressure: the reference's rule, femixedspace.c:1216-1236 -- both restated in oracle_asm.py), in
the rank-local ordering, which for the slab-shaped partitions of 2 ranks is the natural ordering restricted to the block.

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


class BlockIlu0:
    """PCBJACOBI with ILU(0) blocks: z[block r] = ILU0(A[block r, block r])^-1 r[block r]"""

    def __init__(self, A, blocks):
        A = A.tocsr()
        self.blocks = [(np.asarray(b), Ilu0(A[b][:, b])) for b in blocks]
        cover = np.zeros(A.shape[0], int)
        for b, _ in self.blocks:
            cover[b] += 1
        assert np.all(cover == 1), "bjacobi blocks must tile the rows"

    def __call__(self, r):
        z = np.empty_like(r)
        for b, ilu in self.blocks:
            z[b] = ilu(r[b])
        return z


def rank_blocks(nsd, mesh, nranks):
    """(velocity dofs, pressure nodes) owned by each rank, ascending natural index (split-local numbering)"""
    from .oracle_asm import subdomains
    N = [2 * m + 1 for m in mesh[:nsd]]; P = [m + 1 for m in mesh[:nsd]]
    _, sds = subdomains(nsd, mesh, nranks, 0)
    ub, pb = [], []
    for sd in sds:
        u = [np.arange(*sd["own_u"][d]) for d in range(nsd)]; p = [np.arange(*sd["own_p"][d]) for d in range(nsd)]
        if nsd == 2:
            un = (u[1][:, None] * N[0] + u[0][None, :]).ravel(); pn = (p[1][:, None] * P[0] + p[0][None, :]).ravel()
        else:
            un = ((u[2][:, None, None] * N[1] + u[1][None, :, None]) * N[0] + u[0][None, None, :]).ravel()
            pn = ((p[2][:, None, None] * P[1] + p[1][None, :, None]) * P[0] + p[0][None, None, :]).ravel()
        ub.append(np.sort((un[:, None] * nsd + np.arange(nsd)[None, :]).ravel())); pb.append(np.sort(pn))
    return ub, pb


class FieldSplitDefault:
    def __init__(self, opts, nsd=3, lame=False, nranks=1):
        self.o = O.parse_options(opts) if not isinstance(opts, dict) else dict(opts)
        o = self.o
        if "fs" not in o or "mg" in o:
            raise ValueError("FieldSplitDefault needs -fs (and no -mg)")
        self.p = O.Problem(o, nsd=nsd, lame=lame)
        p = self.p; nu = p.nu; self.nu = nu
        A = p.A().scipy().tocsr(); self.A = A
        self.A00 = A[:nu, :nu].tocsr(); self.A01 = A[:nu, nu:].tocsr(); self.A10 = A[nu:, :nu].tocsr(); self.A11 = A[nu:, nu:].tocsr()
        if nranks == 1:
            self.ilu_u = Ilu0(self.A00); self.ilu_p = Ilu0(p.Mp().scipy())
        else:   # bjacobi + ILU(0) per rank on both splits
            mx = int(o.get("mx", 4)); mesh = (mx, int(o.get("my", mx)), int(o.get("mz", mx)) if nsd == 3 else 1)
            ub, pb = rank_blocks(nsd, mesh, nranks)
            self.ilu_u = BlockIlu0(self.A00, ub); self.ilu_p = BlockIlu0(p.Mp().scipy(), pb)
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
