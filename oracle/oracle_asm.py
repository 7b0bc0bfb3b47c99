"""CPU ORACLE (test infrastructure, never imported by the product): the reference's additive-Schwarz preconditioner on
element-patch subdomains, `-saddle_pc_type asm -saddle_pc_asm_dm_subdomains [-dmdafe_overlap k]` (SURVEY 8f rank 3;
Makefile:297, 410, 417).

What the reference + PETSc build, restated:
  * the velocity DMDA ((2m+1)^nsd nodes, stencil width 2) is created with PETSC_DECIDE processor counts
    (femixedspace.c:1153-1158), so the process grid is PETSc's "squarish" choice (DMSetUp_DA_2D / _3D) and node ownership the
    default split M/m + (M%m > i) -> `dmda_grid`, `dmda_split`;
  * every rank fits whole Q2 elements to its node range by parity (_DMCreate_SaddleQ2_BuildElementLayout,
    femixedspace.c:1074-1124) -> `q2_elem_ranges`; pressure nodes: as many as elements, the last rank one more (:1216-1236);
  * DMCreateDomainDecomposition_DMDAFEQ2Q1 (femixedspace.c:823-837) returns ONE subdomain per rank: the closed box of its
    elements grown by `-dmdafe_overlap` element layers, every velocity and pressure node on it (:745-816, index arithmetic
    :292-596) -- as the INNER index set; PCASM then takes the same set as the solve domain (DM-defined subdomains are never
    grown by -pc_asm_overlap) and, its type being the default PC_ASM_RESTRICT, adds each sub-solution back only on the
    dofs the rank OWNS (reverse local scatter).  That last point is what the goldens pin: summing the patch solutions on
    the whole patches gives 4.36 instead of exSaddle2d_asm_1's first residual 2.47014;
  * sub-solves are exact (`-saddle_sub_pc_type lu`, UMFPACK in the goldens): SuperLU here.
Pinned against testref/exSaddle2d_asm_1.ref (9 ranks, overlap 1), exSaddle3d_asm_1.ref (8 ranks, overlap 0) and
exSaddle3d_mg_asm_1.ref (4 ranks, ASM as the GMRES smoother's PC inside -mg): tests/test_oracle_goldens.py."""
import itertools

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from . import oracle as O
from .oracle_mg import petsc_gmres


def dmda_grid(nsd, M, N, P, size):
    """Process grid (m, n, p) of DMDACreate{2,3}d with PETSC_DECIDE (PETSc da2.c / da3.c, 'try for squarish distribution')."""
    if nsd == 2:
        m = int(0.5 + np.sqrt(float(M) * size / float(N))) or 1
        while m > 0:
            n = size // m
            if m * n == size:
                break
            m -= 1
        if M > N and m < n:
            m, n = n, m
        return m, n, 1
    n = int(0.5 + (float(N) * N * size / (float(P) * M)) ** (1.0 / 3.0)) or 1
    while n > 0:
        pm = size // n
        if n * pm == size:
            break
        n -= 1
    n = n or 1
    m = int(0.5 + np.sqrt(float(M) * size / (float(P) * n))) or 1
    while m > 0:
        p = size // (m * n)
        if m * n * p == size:
            break
        m -= 1
    if M > P and m < p:
        m, p = p, m
    return m, n, p


def dmda_split(M, m):
    """nodes per rank along one direction: M/m, the remainder to the low ranks"""
    return [M // m + (1 if (M % m) > i else 0) for i in range(m)]


def q2_elem_ranges(M, m):
    """per rank along one direction: (first element, number of elements), femixedspace.c:1074-1124"""
    out = []; s = 0
    for w in dmda_split(M, m):
        s_el = s if s % 2 == 0 else s - 1
        e = s + w
        e_el = e if e % 2 == 0 else e - 1
        if (e_el - s_el) % 2 or e_el <= s_el:
            raise ValueError("Cannot generate consistent macro element")   # femixedspace.c:1097-1113
        out.append((s_el // 2, (e_el - s_el) // 2)); s += w
    return out


def subdomains(nsd, mesh, size, overlap):
    """One record per rank: element box [lo, hi) grown by `overlap` (clipped), owned velocity-node and pressure-node ranges."""
    N = [2 * m + 1 for m in mesh[:nsd]]
    grid = dmda_grid(nsd, N[0], N[1], N[2] if nsd == 3 else 1, size)
    er = [q2_elem_ranges(N[d], grid[d]) for d in range(nsd)]
    uo = []; po = []
    for d in range(nsd):
        s = np.cumsum([0] + dmda_split(N[d], grid[d])); uo.append([(int(s[i]), int(s[i + 1])) for i in range(grid[d])])
        el = [ne for _, ne in er[d]]; el[-1] += 1
        s = np.cumsum([0] + el); po.append([(int(s[i]), int(s[i + 1])) for i in range(grid[d])])
    out = []
    for idx in itertools.product(*[range(g) for g in reversed(grid[:nsd])]):
        idx = idx[::-1]
        lo = [max(0, er[d][idx[d]][0] - overlap) for d in range(nsd)]
        hi = [min(mesh[d], er[d][idx[d]][0] + er[d][idx[d]][1] + overlap) for d in range(nsd)]
        out.append({"lo": lo, "hi": hi, "own_u": [uo[d][idx[d]] for d in range(nsd)], "own_p": [po[d][idx[d]] for d in range(nsd)]})
    return grid, out


def _box_dofs(nsd, mesh, ur, pr):
    """natural-ordering dofs of the velocity nodes ur[d] = (a, b) and pressure nodes pr[d] = (a, b) (half-open node ranges)"""
    N = [2 * m + 1 for m in mesh[:nsd]]; P = [m + 1 for m in mesh[:nsd]]
    nun = int(np.prod(N))
    u = [np.arange(*ur[d]) for d in range(nsd)]; p = [np.arange(*pr[d]) for d in range(nsd)]
    if nsd == 2:
        un = (u[1][:, None] * N[0] + u[0][None, :]).ravel(); pn = (p[1][:, None] * P[0] + p[0][None, :]).ravel()
    else:
        un = ((u[2][:, None, None] * N[1] + u[1][None, :, None]) * N[0] + u[0][None, None, :]).ravel()
        pn = ((p[2][:, None, None] * P[1] + p[1][None, :, None]) * P[0] + p[0][None, None, :]).ravel()
    return np.concatenate([(un[:, None] * nsd + np.arange(nsd)[None, :]).ravel(), nsd * nun + pn])


def subdomain_dofs(nsd, mesh, sd):
    """(patch dofs ascending, mask of the dofs the rank owns)"""
    patch = _box_dofs(nsd, mesh, [(2 * sd["lo"][d], 2 * sd["hi"][d] + 1) for d in range(nsd)], [(sd["lo"][d], sd["hi"][d] + 1) for d in range(nsd)])
    own = _box_dofs(nsd, mesh, sd["own_u"], sd["own_p"])
    patch = np.sort(patch)
    return patch, np.isin(patch, own)


class AsmPC:
    """PCApply_ASM, PC_ASM_RESTRICT with DM subdomains: z[owned dofs of rank r] = (A_r^-1 r|patch_r)[owned dofs], all ranks."""

    def __init__(self, A, nsd, mesh, size, overlap):
        self.grid, sds = subdomains(nsd, mesh, size, overlap)
        self.subs = []
        for sd in sds:
            patch, keep = subdomain_dofs(nsd, mesh, sd)
            self.subs.append((patch, keep, spla.splu(sp.csc_matrix(A[patch][:, patch]))))
        cover = np.zeros(A.shape[0], int)
        for patch, keep, _ in self.subs:
            cover[patch[keep]] += 1
        assert np.all(cover == 1), "owned dofs must tile the vector"

    def __call__(self, r):
        z = np.zeros_like(r)
        for patch, keep, lu in self.subs:
            y = lu.solve(r[patch])
            z[patch[keep]] = y[keep]
        return z


def solve(opts, nsd, size, lame=False):
    """`mpiexec -n size ./exSaddle{2,3}d <opts>` with -saddle_pc_type asm: returns (x, its, reason, hist, problem)."""
    o = O.parse_options(opts) if not isinstance(opts, dict) else dict(opts)
    if o.get("saddle_pc_type") != "asm" or "saddle_pc_asm_dm_subdomains" not in o or "set_ksp_dm" not in o:
        raise ValueError("oracle ASM: -saddle_pc_type asm -saddle_pc_asm_dm_subdomains -set_ksp_dm (Makefile:298, 411)")
    if o.get("saddle_sub_pc_type", "ilu") != "lu" or o.get("saddle_sub_ksp_type", "preonly") != "preonly":
        raise NotImplementedError("oracle ASM: exact sub-solves only (-saddle_sub_ksp_type preonly -saddle_sub_pc_type lu)")
    p = O.Problem(o, nsd=nsd, lame=lame)
    mx = int(o.get("mx", 4)); my = int(o.get("my", mx)); mz = int(o.get("mz", mx)) if nsd == 3 else 1
    A = p.A().scipy().tocsr()
    pc = AsmPC(A, nsd, (mx, my, mz), size, int(o.get("dmdafe_overlap", 0)))
    x, its, reason, hist = petsc_gmres(lambda v: A @ v, pc, np.array(p.F()), rtol=float(o.get("saddle_ksp_rtol", 1e-5)),
                                       max_it=int(o.get("saddle_ksp_max_it", 10000)), restart=int(o.get("saddle_ksp_gmres_restart", 30)),
                                       side=o.get("saddle_ksp_pc_side", "left"))
    return x, its, reason, hist, p
