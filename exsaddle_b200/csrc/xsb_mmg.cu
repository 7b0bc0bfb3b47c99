// xsb_mmg.cu -- the reference's monolithic multigrid path, `-mg -nlevels L`.
//
// What exSaddle.c:215-270, 331-402 builds with PETSc, on the device:
//   * one Q2-Q1 mesh per level, m_k = m / 2^(L-1-k) (exSaddle.c:217-241); every level operator is RE-ASSEMBLED
//     (PC_MG_GALERKIN_NONE, :339) by the same element kernels, from coefficient fields restricted level by level on the
//     pressure lattice: c_k = (P_p^T c_{k+1}) .* 1 / (P_p^T 1)  (MatRestrict + DMCreateInterpolationScale,
//     femixedspace.c:2139-2150), then Q1-interpolated to the coarse Gauss points (:2168-2215);
//   * interpolation = DMComposite block-diagonal (P_u (x) I_nsd, P_p), (tri)linear on the two node lattices: the
//     stencil kernels of xsb_mg.cu, no stored P;
//   * smoother = exactly max_it iterations of left-Jacobi GMRES from the current iterate (PCMG skips the convergence
//     test), classical Gram-Schmidt, Hessenberg least squares on the host (10 x 10);
//   * coarse = LU of the coarse saddle matrix (indefinite): dense Gauss-Jordan inverse WITH partial pivoting on the
//     device, applied as a GEMV, so the V-cycle has no host round trip except the smoother's scalar fetches.
// Each coarse level is a child handle assembled by fe_assemble with `nodal_in` set.
#include "xsb.h"

static inline unsigned nblk(int64_t n, int bs = 256) { return (unsigned)((n + bs - 1) / bs); }

struct MmgLevel {
  xsb_ctx ctx = nullptr;      // level problem (finest = the user's handle)
  double *idiag = nullptr;    // Jacobi on the full saddle operator (zero diagonal -> 1)
  double *x = nullptr, *b = nullptr, *r = nullptr, *t = nullptr;
  std::vector<double *> V;    // GMRES basis of the smoother
  double *inv = nullptr;      // coarsest: dense inverse
  void *fsc = nullptr;        // coarsest with -fs_coarse: fieldsplit-preconditioned FGMRES instead of the dense inverse (xsb_fs.cu)
  void *asmpc = nullptr;      // -saddle_mg_levels_pc_type asm: element-patch ASM instead of Jacobi (xsb_asm.cu)
};
struct Mmg { int nlev = 0, smooth_its = 2, restart = 30; std::vector<MmgLevel> lev; };

// ------------------------------------------------------------------ coefficient restriction on the pressure lattice
__global__ void k_pp_restrict(int fnx, int fny, int fnz, int cnx, int cny, int cnz, int nslot, int64_t fnp, int64_t cnp,
                              const double *__restrict__ cf, double *__restrict__ cc)
{
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; if (t >= cnp) return;
  const int I = (int)(t % cnx), J = (int)((t / cnx) % cny), K = (int)(t / ((int64_t)cnx * cny));
  double ones = 0.0, acc[XSB_NSLOT];
  for (int s = 0; s < XSB_NSLOT; ++s) acc[s] = 0.0;
  const int c0 = fnz > 1 ? -1 : 0, c1 = fnz > 1 ? 1 : 0;
  for (int c = c0; c <= c1; ++c) for (int b = -1; b <= 1; ++b) for (int a = -1; a <= 1; ++a) {   // ascending fine index (MatMultTranspose order)
    const int i = 2 * I + a, j = 2 * J + b, k = fnz > 1 ? 2 * K + c : 0;
    if (i < 0 || i >= fnx || j < 0 || j >= fny || k < 0 || k >= fnz) continue;
    const double w = (a ? 0.5 : 1.0) * (b ? 0.5 : 1.0) * (c ? 0.5 : 1.0);
    const int64_t f = i + (int64_t)j * fnx + (int64_t)k * fnx * fny;
    ones += w * 1.0;
    for (int s = 0; s < nslot; ++s) acc[s] += w * cf[(int64_t)s * fnp + f];
  }
  const double scale = 1.0 / ones;   // DMCreateInterpolationScale: VecReciprocal(P^T 1)
  for (int s = 0; s < nslot; ++s) cc[(int64_t)s * cnp + t] = acc[s] * scale;   // VecPointwiseMult (:2149)
}

// ------------------------------------------------------------------ dense inverse with partial pivoting (Gauss-Jordan on [M | I])
__global__ void k_csr_to_dense(int n, const int *__restrict__ ia, const int *__restrict__ ja, const double *__restrict__ a, double *__restrict__ M)
{
  const int row = blockIdx.x * blockDim.x + threadIdx.x; if (row >= n) return;
  for (int k = ia[row]; k < ia[row + 1]; ++k) M[(int64_t)row * n + ja[k]] = a[k];
}
__global__ void k_identity(int n, double *M) { const int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) M[(int64_t)i * n + i] = 1.0; }
// pivot row of column k among rows >= k (first maximum, like LAPACK's idamax), one block
__global__ void __launch_bounds__(1024) k_gjp_pivot(int n, int k, const double *__restrict__ M, int *piv, double *pval, int *flag)
{
  __shared__ double sv[32]; __shared__ int si[32];
  double best = -1.0; int bi = k;
  for (int i = k + threadIdx.x; i < n; i += blockDim.x) { const double v = fabs(M[(int64_t)i * n + k]); if (v > best) { best = v; bi = i; } }
  for (int o = 16; o > 0; o >>= 1) {
    const double ov = __shfl_xor_sync(0xffffffffu, best, o); const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
  if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = best; si[threadIdx.x >> 5] = bi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) if (sv[w] > best || (sv[w] == best && si[w] < bi)) { best = sv[w]; bi = si[w]; }
    *piv = bi; *pval = M[(int64_t)bi * n + k]; if (!(best > 0.0)) *flag = 1;
  }
}
// swap rows k and piv in [M | Inv], then scale row k by 1 / pivot; saves column k of M (the multipliers) first
__global__ void k_gjp_swap_scale(int n, int k, double *M, double *Inv, const int *piv, const double *pval)
{
  const int j = blockIdx.x * blockDim.x + threadIdx.x; if (j >= n) return;
  const int p = *piv;
  const double pivot = *pval;   // read from the pivot kernel's copy: thread j = k overwrites M[p][k] / M[k][k] below
  double a = M[(int64_t)p * n + j], b = Inv[(int64_t)p * n + j];
  if (p != k) { M[(int64_t)p * n + j] = M[(int64_t)k * n + j]; Inv[(int64_t)p * n + j] = Inv[(int64_t)k * n + j]; }
  M[(int64_t)k * n + j] = a / pivot; Inv[(int64_t)k * n + j] = b / pivot;
}
__global__ void k_gjp_col(int n, int k, const double *__restrict__ M, double *__restrict__ colk)
{ const int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) colk[i] = i == k ? 0.0 : M[(int64_t)i * n + k]; }
__global__ void k_gjp_eliminate(int n, int k, double *M, double *Inv, const double *__restrict__ colk)
{
  const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
  if (j >= n) return;
  const double f = colk[i]; if (f == 0.0) return;
  if (j >= k) M[(int64_t)i * n + j] -= f * M[(int64_t)k * n + j];
  Inv[(int64_t)i * n + j] -= f * Inv[(int64_t)k * n + j];
}
__global__ void k_gemv_rows(int n, const double *__restrict__ M, const double *__restrict__ b, double *__restrict__ x)
{
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= n) return;
  double acc = 0.0;
  for (int j = lane; j < n; j += 32) acc += M[(int64_t)row * n + j] * b[j];
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) x[row] = acc;
}
// Inv = M^-1 for a dense n x n matrix M (row-major, destroyed): Gauss-Jordan on [M | I] with partial pivoting
int dense_invert_pivoted(xsb_ctx c, int n, double *M, double *Inv)
{
  cudaStream_t st = c->stream;
  double *colk = nullptr, *pval = nullptr; int *piv = nullptr, *flag = nullptr;
  CUDA_OK(cudaMalloc(&colk, sizeof(double) * ((size_t)n + 2))); pval = colk + n;
  CUDA_OK(cudaMalloc(&piv, sizeof(int) * 2)); flag = piv + 1;
  CUDA_OK(cudaMemsetAsync(piv, 0, sizeof(int) * 2, st));
  CUDA_OK(cudaMemsetAsync(Inv, 0, sizeof(double) * (size_t)n * n, st));
  k_identity<<<nblk(n), 256, 0, st>>>(n, Inv); KERNEL_OK();
  dim3 g2((n + 255) / 256, n);
  for (int k = 0; k < n; ++k) {
    k_gjp_pivot<<<1, 1024, 0, st>>>(n, k, M, piv, pval, flag); KERNEL_OK();
    k_gjp_swap_scale<<<nblk(n), 256, 0, st>>>(n, k, M, Inv, piv, pval); KERNEL_OK();
    k_gjp_col<<<nblk(n), 256, 0, st>>>(n, k, M, colk); KERNEL_OK();
    k_gjp_eliminate<<<g2, 256, 0, st>>>(n, k, M, Inv, colk); KERNEL_OK();
  }
  int hflag = 0; CUDA_OK(cudaMemcpyAsync(&hflag, flag, sizeof(int), cudaMemcpyDeviceToHost, st)); CUDA_OK(cudaStreamSynchronize(st));
  cudaFree(colk); cudaFree(piv);
  if (hflag) return xsb_fail(c, XSB_ERR_BREAKDOWN, "singular matrix in a dense direct solve (-mg coarse level / ASM subdomain)");
  return 0;
}
static int dense_inverse_pivoted(xsb_ctx c, const Csr &A, double **inv_out)
{
  const int n = A.n; cudaStream_t st = c->stream;
  if (n > 6600) return xsb_fail(c, XSB_ERR_SUP, "coarsest -mg level has %d unknowns; the dense coarse solve supports <= 6600 (use more -nlevels)", n);
  double *M = nullptr, *Inv = nullptr;
  XSB_CHK(dev_alloc(c, &Inv, (size_t)n * n));
  CUDA_OK(cudaMalloc(&M, sizeof(double) * (size_t)n * n)); CUDA_OK(cudaMemsetAsync(M, 0, sizeof(double) * (size_t)n * n, st));
  k_csr_to_dense<<<nblk(n), 256, 0, st>>>(n, A.ia, A.ja, A.a, M); KERNEL_OK();
  int rc = dense_invert_pivoted(c, n, M, Inv);
  cudaFree(M);
  if (rc) return rc;
  *inv_out = Inv;
  return 0;
}

// ------------------------------------------------------------------ transfers on [u | p]
static int mmg_restrict(xsb_ctx c, const MmgLevel &F, const MmgLevel &C, const double *rf, double *bc)
{
  const Lattice &lf = F.ctx->lat, &lc = C.ctx->lat;
  Level vf, vc; vf.nx = lf.NX; vf.ny = lf.NY; vf.nz = lf.NZ; vc.nx = lc.NX; vc.ny = lc.NY; vc.nz = lc.NZ;
  XSB_CHK(mg_restrict(c, vf, vc, rf, bc));
  return mg_restrict_scalar(c, lf.PX, lf.PY, lf.PZ, lc.PX, lc.PY, lc.PZ, rf + lf.nu, bc + lc.nu);
}
static int mmg_prolong_add(xsb_ctx c, const MmgLevel &F, const MmgLevel &C, const double *xc, double *xf)
{
  const Lattice &lf = F.ctx->lat, &lc = C.ctx->lat;
  Level vf, vc; vf.nx = lf.NX; vf.ny = lf.NY; vf.nz = lf.NZ; vc.nx = lc.NX; vc.ny = lc.NY; vc.nz = lc.NZ;
  XSB_CHK(mg_prolong_add(c, vf, vc, xc, xf));
  return mg_prolong_add_scalar(c, lf.PX, lf.PY, lf.PZ, lc.PX, lc.PY, xc + lc.nu, xf + lf.nu);
}

// ------------------------------------------------------------------ smoother: KSPSolve_GMRES, left Jacobi, `its` iterations, no test
// smoother PC: Jacobi, or ASM on the reference's element patches (Makefile:417)
static int mmg_pc(xsb_ctx c, MmgLevel &L, int64_t n, const double *in, double *out)
{ return L.asmpc ? asm_apply(c, L.asmpc, in, out) : vec_pmult(c, n, L.idiag, in, out); }

static int mmg_smooth(xsb_ctx c, Mmg &G, MmgLevel &L, const double *b, double *x, int its, bool x_is_zero)
{
  const int64_t n = L.ctx->lat.n; const Ranges rg = whole(n);
  int done = 0;
  std::vector<double> H, hcol, y;
  while (done < its) {
    const int m = its - done < G.restart ? its - done : G.restart;
    while ((int)L.V.size() < m + 1) { double *v = nullptr; XSB_CHK(dev_alloc(c, &v, (size_t)n)); L.V.push_back(v); }
    // r = B (b - A x)
    if (x_is_zero && done == 0) XSB_CHK(mmg_pc(c, L, n, b, L.V[0]));
    else { XSB_CHK(spmv_csr(c, L.ctx->A, x, L.t)); XSB_CHK(vec_aypx(c, n, -1.0, b, L.t)); XSB_CHK(mmg_pc(c, L, n, L.t, L.V[0])); }
    XSB_CHK(vec_mdot(c, rg, L.V[0], nullptr, 0, true, c->scal));
    double beta2; XSB_CHK(vec_fetch(c, c->scal, 1, &beta2));
    const double beta = sqrt(beta2);
    if (beta == 0.0) return 0;
    XSB_CHK(vec_scale(c, n, 1.0 / beta, L.V[0]));
    H.assign((size_t)(m + 1) * m, 0.0); hcol.assign(m + 2, 0.0);
    int k = 0;
    for (int j = 0; j < m; ++j) {
      double *w = L.V[j + 1];
      XSB_CHK(spmv_csr(c, L.ctx->A, L.V[j], L.t));
      XSB_CHK(mmg_pc(c, L, n, L.t, w));                                            // w = B A v_j
      XSB_CHK(vec_mdot(c, rg, w, L.V.data(), j + 1, false, c->scal));              // classical Gram-Schmidt, one pass
      XSB_CHK(vec_maxpy_dev(c, n, w, L.V.data(), j + 1, c->scal, -1.0));
      XSB_CHK(vec_mdot(c, rg, w, nullptr, 0, true, c->scal + j + 1));
      XSB_CHK(vec_scale_by_inv_sqrt(c, n, w, c->scal + j + 1));
      XSB_CHK(vec_fetch(c, c->scal, j + 2, hcol.data()));
      for (int i = 0; i <= j; ++i) H[(size_t)i * m + j] = hcol[i];
      H[(size_t)(j + 1) * m + j] = sqrt(hcol[j + 1]);
      k = j + 1;
      if (hcol[j + 1] == 0.0) break;
    }
    // least squares min || beta e1 - H y || by Givens rotations
    std::vector<double> R(H), g(k + 1, 0.0); g[0] = beta;
    for (int j = 0; j < k; ++j) {
      const double a = R[(size_t)j * m + j], bb = R[(size_t)(j + 1) * m + j], tt = sqrt(a * a + bb * bb);
      const double cs = tt == 0.0 ? 1.0 : a / tt, sn = tt == 0.0 ? 0.0 : bb / tt;
      for (int q = j; q < k; ++q) { const double u = R[(size_t)j * m + q], v = R[(size_t)(j + 1) * m + q]; R[(size_t)j * m + q] = cs * u + sn * v; R[(size_t)(j + 1) * m + q] = -sn * u + cs * v; }
      const double gu = g[j]; g[j] = cs * gu; g[j + 1] = -sn * gu;
    }
    y.assign(k, 0.0);
    for (int i = k - 1; i >= 0; --i) { double s = g[i]; for (int q = i + 1; q < k; ++q) s -= R[(size_t)i * m + q] * y[q]; y[i] = s / R[(size_t)i * m + i]; }
    XSB_CHK(vec_maxpy_host(c, n, x, L.V.data(), k, y.data()));
    done += k;
    if (k < m) break;
    x_is_zero = false;
  }
  return 0;
}

static int mmg_cycle(xsb_ctx c, Mmg &G, int l)
{
  MmgLevel &L = G.lev[l];
  const int64_t n = L.ctx->lat.n;
  if (l == 0 && L.fsc) return fsc_solve(c, L.fsc, L.b, L.x);
  if (l == 0) { k_gemv_rows<<<nblk((int64_t)n * 32), 256, 0, c->stream>>>((int)n, L.inv, L.b, L.x); KERNEL_OK(); return 0; }
  MmgLevel &C = G.lev[l - 1];
  XSB_CHK(vec_set(c, n, 0.0, L.x));
  XSB_CHK(mmg_smooth(c, G, L, L.b, L.x, G.smooth_its, true));
  XSB_CHK(spmv_csr(c, L.ctx->A, L.x, L.r)); XSB_CHK(vec_aypx(c, n, -1.0, L.b, L.r));   // r = b - A x
  XSB_CHK(mmg_restrict(c, L, C, L.r, C.b));
  XSB_CHK(mmg_cycle(c, G, l - 1));
  XSB_CHK(mmg_prolong_add(c, L, C, C.x, L.x));
  return mmg_smooth(c, G, L, L.b, L.x, G.smooth_its, false);
}

int mmg_apply(xsb_ctx c, const double *r, double *z)
{
  Mmg &G = *(Mmg *)c->mmg; MmgLevel &L = G.lev[G.nlev - 1];
  const int64_t n = c->lat.n;
  double *save = L.b; L.b = const_cast<double *>(r);
  int rc = mmg_cycle(c, G, G.nlev - 1);
  L.b = save;
  if (rc) return rc;
  return vec_copy(c, n, L.x, z);
}

void mmg_free(xsb_ctx c)
{
  if (!c->mmg) return;
  Mmg *G = (Mmg *)c->mmg;
  for (int l = 0; l < G->nlev; ++l) { if (G->lev[l].asmpc) asm_free(G->lev[l].asmpc); if (G->lev[l].fsc) fsc_free(G->lev[l].fsc); }
  for (int l = 0; l + 1 < G->nlev; ++l) { xsb_ctx ch = G->lev[l].ctx; if (ch) { dev_free_all(ch); delete ch; } }
  delete G; c->mmg = nullptr;
}

int mmg_setup(xsb_ctx c)
{
  Options &o = c->opt; const int nsd = c->nsd; cudaStream_t st = c->stream;
  if (c->slab.nranks > 1) return xsb_fail(c, XSB_ERR_SUP, "-mg is implemented for one GPU");
  if (c->no_A) return xsb_fail(c, XSB_ERR_SUP, "-mg needs the assembled operator (not -xsb_matrix_free full)");
  mmg_free(c);
  const int L = o.integer("nlevels", 1);
  if (L < 2) return xsb_fail(c, XSB_ERR_SUP, "-nlevels < 2 specified with -mg");                       // exSaddle.c:209
  if (L > XSB_MAX_LEVELS) return xsb_fail(c, XSB_ERR_SUP, "MG levels must be less than %d", XSB_MAX_LEVELS);   // :211
  const std::string spc = o.str("saddle_mg_levels_pc_type", "sor");
  if (o.str("saddle_mg_levels_ksp_type", "chebyshev") != "gmres" || (spc != "jacobi" && spc != "asm"))
    return xsb_fail(c, XSB_ERR_SUP, "-mg smoothers: -saddle_mg_levels_ksp_type gmres -saddle_mg_levels_pc_type jacobi|asm (the reference's tests)");
  if (spc == "asm" && (!o.flag("saddle_mg_levels_pc_asm_dm_subdomains") || o.str("saddle_mg_levels_sub_pc_type", "ilu") != "lu" || o.str("saddle_mg_levels_sub_ksp_type", "preonly") != "preonly"))
    return xsb_fail(c, XSB_ERR_SUP, "-saddle_mg_levels_pc_type asm is supported with the reference's element patches and exact sub-solves: -saddle_mg_levels_pc_asm_dm_subdomains -saddle_mg_levels_sub_pc_type lu");
  o.has("saddle_mg_levels_sub_pc_factor_mat_solver_type");
  const int ratio = 1 << (L - 1);
  const int m[3] = {c->lat.mx, c->lat.my, nsd == 3 ? c->lat.mz : ratio};
  for (int d = 0; d < 3; ++d) {
    if (ratio > m[d]) return xsb_fail(c, XSB_ERR_ARG, "Too much refinement 2 ^ %d = %d requested for the given problem size (%d x %d x %d elements)", L - 1, ratio, c->lat.mx, c->lat.my, c->lat.mz);   // :219
    if (m[d] % ratio) return xsb_fail(c, XSB_ERR_ARG, "Coarsening ratio of 2 ^ %d = %d is incompatible with problem size (%d x %d x %d elements)", L - 1, ratio, c->lat.mx, c->lat.my, c->lat.mz);   // :220
  }
  Mmg *G = new Mmg(); c->mmg = G;
  G->nlev = L; G->lev.resize(L);
  G->smooth_its = o.integer("saddle_mg_levels_ksp_max_it", 2);          // PCMG default: 2 smoothing steps (mg_fs_coarse_1.ref:143)
  G->restart = o.integer("saddle_mg_levels_ksp_gmres_restart", 30);
  o.has("saddle_mg_coarse_pc_factor_mat_solver_type"); o.has("saddle_mg_coarse_redundant_pc_factor_mat_solver_type");
  G->lev[L - 1].ctx = c;
  for (int k = L - 2; k >= 0; --k) {
    xsb_ctx f = G->lev[k + 1].ctx;
    xsb_ctx ch = new xsb_ctx_s(); G->lev[k].ctx = ch;
    ch->nsd = nsd; ch->lame = c->lame; ch->device = c->device; ch->have_device = true; ch->stream = st; ch->opt = c->opt;
    const int fac = 1 << (L - 1 - k);
    char buf[32];
    snprintf(buf, sizeof(buf), "%d", c->lat.mx / fac); ch->opt.kv["mx"] = buf;
    snprintf(buf, sizeof(buf), "%d", c->lat.my / fac); ch->opt.kv["my"] = buf;
    if (nsd == 3) { snprintf(buf, sizeof(buf), "%d", c->lat.mz / fac); ch->opt.kv["mz"] = buf; }
    ch->opt.kv.erase("xsb_matrix_free");
    // restricted nodal coefficient fields of the finer level
    const Lattice &lf = f->lat;
    const int cpx = lf.mx / 2 + 1, cpy = lf.my / 2 + 1, cpz = nsd == 3 ? lf.mz / 2 + 1 : 1;
    const int64_t cnp = (int64_t)cpx * cpy * cpz;
    double *nodal = nullptr; XSB_CHK(dev_alloc(ch, &nodal, (size_t)XSB_NSLOT * cnp));
    k_pp_restrict<<<nblk(cnp, 128), 128, 0, st>>>(lf.PX, lf.PY, lf.PZ, cpx, cpy, cpz, XSB_NSLOT, lf.npn, cnp, f->coeff_nodal, nodal); KERNEL_OK();
    ch->nodal_in = nodal;
    int rc = fe_assemble(ch);
    if (rc) { c->err = "-mg level assembly: " + ch->err; return rc; }
  }
  for (int k = 0; k < L; ++k) {
    MmgLevel &lv = G->lev[k]; const int64_t n = lv.ctx->lat.n;
    XSB_CHK(dev_alloc(c, &lv.idiag, (size_t)n)); XSB_CHK(csr_diag_inv(c, lv.ctx->A, lv.idiag));
    XSB_CHK(dev_alloc(c, &lv.x, (size_t)n)); XSB_CHK(dev_alloc(c, &lv.b, (size_t)n)); XSB_CHK(dev_alloc(c, &lv.r, (size_t)n)); XSB_CHK(dev_alloc(c, &lv.t, (size_t)n));
  }
  if (spc == "asm") {   // one patch per rank of the communicator the reference would run on; every level's DMDA picks its own process grid
    const int size = o.integer("xsb_ranks", 1), ov = o.integer("dmdafe_overlap", 0);
    for (int k = 1; k < L; ++k) XSB_CHK(asm_setup(c, G->lev[k].ctx, size, ov, &G->lev[k].asmpc));
  }
  if (o.flag("fs_coarse")) return fsc_setup(c, G->lev[0].ctx, &G->lev[0].fsc);   // exSaddle.c:362-400
  return dense_inverse_pivoted(c, G->lev[0].ctx->A, &G->lev[0].inv);
}
