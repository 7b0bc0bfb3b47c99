// xsb_ksp.cu -- host-side Krylov drivers over the device kernels: what KSPSolve (exSaddle.c:425) runs for
// the solver trees of the reference's tests (Makefile:254-513, abf.opts):
//   outer  KSPGMRES / KSPFGMRES, restart 30, classical Gram-Schmidt without refinement, Givens QR (App. B.5)
//   PC     PCJACOBI, or PCFIELDSPLIT Schur/UPPER with the user matrix Mpscaled (exSaddle.c:312-321, App. B.2):
//            y_p = ILU0(Mp)^-1 x_p ;  y_u = GCR[A00, PCMG]( x_u - A01 y_p )
//   inner  KSPGCR (App. B.6) preconditioned by one MG V-cycle (xsb_mg.cu)
// The host only sequences kernels and does the O(restart^2) Hessenberg algebra; every vector stays in HBM.
// One stream-synchronising scalar fetch per Krylov iteration (the residual norm the convergence test needs).
#include "xsb.h"
#include <complex>

// ------------------------------------------------------------------ eigenvalues of a small Hessenberg matrix
// Shifted QR iteration in complex arithmetic with Givens rotations and deflation (Wilkinson shift).
int hess_eig(int n, const double *H, int ldh, double *wr, double *wi)
{
  typedef std::complex<double> cd;
  std::vector<cd> A((size_t)n * n);
  for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) A[(size_t)i * n + j] = (j >= i - 1) ? cd(H[(size_t)i * ldh + j], 0.0) : cd(0.0, 0.0);
  auto a = [&](int i, int j) -> cd & { return A[(size_t)i * n + j]; };
  int hi = n - 1, iter = 0;
  std::vector<cd> cs(n), sn(n);
  while (hi >= 0) {
    if (hi == 0) { wr[0] = a(0, 0).real(); wi[0] = a(0, 0).imag(); break; }
    // deflation check from the bottom
    int lo = hi;
    while (lo > 0) {
      double s = std::abs(a(lo - 1, lo - 1)) + std::abs(a(lo, lo));
      if (s == 0.0) s = 1.0;
      if (std::abs(a(lo, lo - 1)) <= 2.3e-16 * s) { a(lo, lo - 1) = 0.0; break; }
      --lo;
    }
    if (lo == hi) { wr[hi] = a(hi, hi).real(); wi[hi] = a(hi, hi).imag(); --hi; iter = 0; continue; }
    if (++iter > 500) return 1;
    // Wilkinson shift: eigenvalue of the trailing 2x2 closest to a(hi,hi)
    cd p = a(hi - 1, hi - 1), q = a(hi - 1, hi), r = a(hi, hi - 1), s = a(hi, hi);
    cd tr = p + s, det = p * s - q * r, disc = std::sqrt(tr * tr - 4.0 * det);
    cd e1 = 0.5 * (tr + disc), e2 = 0.5 * (tr - disc);
    cd mu = std::abs(e1 - s) < std::abs(e2 - s) ? e1 : e2;
    if (iter % 11 == 10) mu += cd(std::abs(a(hi, hi - 1)), 0.0);   // exceptional shift
    // QR step on the active block [lo, hi]
    for (int i = lo; i <= hi; ++i) a(i, i) -= mu;
    for (int k = lo; k < hi; ++k) {
      cd x = a(k, k), y = a(k + 1, k); double nr = std::sqrt(std::norm(x) + std::norm(y));
      cd c_ = nr == 0.0 ? cd(1.0) : x / nr, s_ = nr == 0.0 ? cd(0.0) : y / nr;
      cs[k] = c_; sn[k] = s_;
      for (int j = k; j < n; ++j) { cd u = a(k, j), v = a(k + 1, j); a(k, j) = std::conj(c_) * u + std::conj(s_) * v; a(k + 1, j) = -s_ * u + c_ * v; }
    }
    for (int k = lo; k < hi; ++k) {
      cd c_ = cs[k], s_ = sn[k];
      for (int i = 0; i <= (k + 2 < n - 1 ? k + 2 : n - 1); ++i) { cd u = a(i, k), v = a(i, k + 1); a(i, k) = u * c_ + v * s_; a(i, k + 1) = -u * std::conj(s_) + v * std::conj(c_); }
    }
    for (int i = lo; i <= hi; ++i) a(i, i) += mu;
  }
  return 0;
}

// ------------------------------------------------------------------ operators
// y = A x on the full saddle operator; on slabs: ghost update of x, then the owned velocity and pressure rows
int op_full_mult(xsb_ctx c, const double *x, double *y)
{
  if (c->no_A) {   // operator-free: y_u = A00 x_u (element kernel) + A01 x_p ; y_p = A11 x_p + A10 x_u
    const Lattice &L = c->lat; Epilogue ep;
    XSB_CHK(mf_setup(c));   // returns at once when the element-kernel state is current (the product may precede xsb_ksp_setup)
    XSB_CHK(comm_halo_full(c, const_cast<double *>(x)));
    XSB_CHK(mf_a00_apply(c, x, y, ep));
    if (c->so.mf_grad || !c->ksp_ready) {   // gradient / divergence by their closed-form stencils: no matrix read (before the first set-up: the default of this mode)
      XSB_CHK(grad_apply(c, x + L.nu, y, c->own_u.off0, c->own_u.len0, y));
      XSB_CHK(spmv_csr(c, c->A11, x + L.nu, y + L.nu, c->own_p.off0, c->own_p.len0));
      return div_apply(c, x, y + L.nu, c->own_p.off0, c->own_p.len0, y + L.nu);
    }
    XSB_CHK(spmv_csr(c, c->A01, x + L.nu, y, c->own_u.off0, c->own_u.len0, y));
    XSB_CHK(spmv_csr(c, c->A11, x + L.nu, y + L.nu, c->own_p.off0, c->own_p.len0));
    return spmv_csr(c, c->A10, x, y + L.nu, c->own_p.off0, c->own_p.len0, y + L.nu);
  }
  if (c->slab.nranks == 1) return spmv_csr(c, c->A, x, y);
  XSB_CHK(comm_halo_full(c, const_cast<double *>(x)));
  XSB_CHK(spmv_csr(c, c->A, x, y, c->own_full.off0, c->own_full.len0));
  return spmv_csr(c, c->A, x, y, c->own_full.off1, c->own_full.len1);
}
static int full_mult(xsb_ctx c, const double *x, double *y) { c->n_a++; XSB_CHK(prof_mark(c, PROF_FULL)); XSB_CHK(op_full_mult(c, x, y)); return prof_mark(c, PROF_OTHER); }
static int a00_mult(xsb_ctx c, const double *x, double *y) { Epilogue ep; return spmv_a00_fine(c, c->A00, x, y, ep); }

// ------------------------------------------------------------------ KSPSolve_GCR on A00, right PC = PCMG
static int gcr_solve(xsb_ctx c, const double *b, double *x, int *its_out, int *reason_out)
{
  const SolverOpts &s = c->so; const int64_t n = c->lat.nu; const int m = s.u_restart;
  double *r = c->gcr_r, *dots = c->scal;     // dots[0..m) mdot results, dots[64] r.v, dots[65] v.v, dots[66] ||r||^2
  double h[4];
  int its = 0; bool done = false;
  XSB_CHK(vec_set(c, n, 0.0, x));
  XSB_CHK(vec_copy(c, n, b, r));            // r = b - A*0
  const Ranges &rg = c->own_u;
  XSB_CHK(vec_mdot(c, rg, r, nullptr, 0, true, dots + 66));
  XSB_CHK(vec_fetch(c, dots + 66, 1, h));
  const double rnorm0 = sqrt(h[0]), ttol = fmax(s.u_rtol * rnorm0, 1e-50);
  int reason = 0;
  if (rnorm0 <= ttol) { *its_out = 0; if (reason_out) *reason_out = 3; return 0; }
  while (!done && its < s.u_max_it) {
    for (int k = 0; k < m; ++k) {
      if ((int)c->GV.size() <= k) { double *v = nullptr, *sv = nullptr; c->phase = 1; int rc = dev_alloc(c, &v, (size_t)n); if (!rc) rc = dev_alloc(c, &sv, (size_t)n); c->phase = 0; if (rc) return rc; c->GV.push_back(v); c->GS.push_back(sv); }
      double *v = c->GV[k], *sv = c->GS[k];
      XSB_CHK(mg_vcycle(c, r, sv));                                   // s = B^-1 r
      XSB_CHK(a00_mult(c, sv, v));                                    // v = A s
      if (k > 0) {
        XSB_CHK(vec_mdot(c, rg, v, c->GV.data(), k, false, dots));    // VecMDot
        XSB_CHK(vec_maxpy_dev(c, n, v, c->GV.data(), k, dots, -1.0)); // v -= sum (v.v_i) v_i
        XSB_CHK(vec_maxpy_dev(c, n, sv, c->GS.data(), k, dots, -1.0));
      }
      { double *two[2] = {r, v}; XSB_CHK(vec_mdot(c, rg, v, two, 2, false, dots + 64)); }   // VecDotNorm2: [r.v, v.v]
      XSB_CHK(vec_gcr_update(c, rg, dots + 64, v, sv, x, r, dots + 66));
      XSB_CHK(vec_fetch(c, dots + 66, 1, h));
      const double norm_r = sqrt(h[0]);
      its++;
      if (norm_r <= ttol) reason = 2; else if (norm_r >= 1e4 * rnorm0) reason = -4; else if (its >= s.u_max_it) reason = -3;   // KSPConvergedDefault order
      if (reason) { done = true; break; }
    }
  }
  *its_out = its; if (reason_out) *reason_out = reason ? reason : -3;
  return 0;
}

// ------------------------------------------------------------------ PCApply
int pc_apply(xsb_ctx c, const double *r, double *z, int *inner, int *inner_reason)
{
  const Lattice &L = c->lat;
  if (inner) *inner = 0;
  if (c->so.pc_type == 1) return vec_pmult(c, L.n, c->idiagA, r, z);
  if (c->so.pc_type == 0) return vec_copy(c, L.n, r, z);
  if (c->so.pc_type == 3) return mmg_apply(c, r, z);
  if (c->so.pc_type == 4) return fsd_apply(c, r, z);
  if (c->so.pc_type == 5) return asm_apply(c, c->asmpc, r, z);
  // PCApply_FieldSplit_Schur, PC_FIELDSPLIT_SCHUR_FACT_UPPER
  double *yp = z + L.nu;
  XSB_CHK(prof_mark(c, PROF_ILU));
  if (c->so.p_pc == 0) XSB_CHK(ilu_apply(c, r + L.nu + c->own_p.off0, yp + c->own_p.off0)); else XSB_CHK(vec_pmult(c, L.np, c->mp_idiag, r + L.nu, yp));
  XSB_CHK(prof_mark(c, PROF_CSR));
  XSB_CHK(comm_halo_p(c, yp));
  if (c->so.mf_grad) XSB_CHK(grad_apply(c, yp, c->fs_tu, c->own_u.off0, c->own_u.len0));
  else XSB_CHK(spmv_csr(c, c->A01, yp, c->fs_tu, c->own_u.off0, c->own_u.len0));
  XSB_CHK(prof_mark(c, PROF_OTHER));
  XSB_CHK(vec_aypx(c, L.nu, -1.0, r, c->fs_tu));      // t_u = x_u - A01 y_p
  int its = 0, why = 0;
  XSB_CHK(gcr_solve(c, c->fs_tu, z, &its, &why));
  if (inner) *inner = its;
  if (inner_reason) *inner_reason = why;
  return 0;
}

// ------------------------------------------------------------------ options -> solver tree (KSPSetFromOptions)
static int read_solver_options(xsb_ctx c)
{
  Options &o = c->opt; SolverOpts &s = c->so; s = SolverOpts(); c->mf_opts_read = false;   // the element-kernel options live in `so` too
  const std::string ksp = o.str("saddle_ksp_type", "gmres");
  if (ksp == "gmres") s.ksp_type = 0; else if (ksp == "fgmres") s.ksp_type = 1;
  else return xsb_fail(c, XSB_ERR_SUP, "-saddle_ksp_type %s not supported (gmres|fgmres)", ksp.c_str());
  const bool fs = o.flag("fs"), mg = o.flag("mg");
  if (fs && mg) return xsb_fail(c, XSB_ERR_SUP, "both -fs and -mg supplied");              // exSaddle.c:205
  if (o.flag("fs_coarse") && !mg) return xsb_fail(c, XSB_ERR_SUP, "-fs_coarse supplied without -mg");   // exSaddle.c:210
  if (o.integer("nlevels", 1) > 1 && fs) return xsb_fail(c, XSB_ERR_SUP, "-nlevels > 1 specified with -fs");      // exSaddle.c:207
  if (o.integer("nlevels", 1) > 1 && !mg) return xsb_fail(c, XSB_ERR_SUP, "-nlevels > 1 specified without -mg"); // exSaddle.c:208
  if (o.flag("set_ksp_dm") && (mg || fs)) return xsb_fail(c, XSB_ERR_SUP, "-set_ksp_dm not intended for use with -mg or -fs");   // exSaddle.c:212
  if (mg) {
    s.pc_type = 3;   // monolithic PCMG on the saddle operator (xsb_mmg.cu)
    o.has("fs_coarse");   // fieldsplit coarse solver: validated in mmg_setup / fsc_setup
  } else if (fs && o.str("saddle_fieldsplit_u_pc_type", "") != "mg") {
    s.pc_type = 4;   // plain -fs: PETSc's default sub-solvers (GMRES + ILU(0) on A00, nested Schur solves); validated in fsd_setup (xsb_fs.cu)
    o.has("saddle_fieldsplit_u_ksp_type"); o.has("saddle_fieldsplit_p_ksp_type"); o.has("saddle_fieldsplit_u_ksp_max_it"); o.has("saddle_fieldsplit_p_pc_type"); o.has("xsb_ranks");
  } else if (fs) {
    s.pc_type = 2;
    if (o.str("saddle_fieldsplit_u_ksp_type", "") != "gcr" || o.str("saddle_fieldsplit_p_ksp_type", "") != "preonly")
      return xsb_fail(c, XSB_ERR_SUP, "-fs with -saddle_fieldsplit_u_pc_type mg is supported as the abf.opts tree: fieldsplit_u gcr+mg, fieldsplit_p preonly");
    if (o.str("saddle_fieldsplit_u_mg_levels_ksp_type", "chebyshev") != "chebyshev" || o.str("saddle_fieldsplit_u_mg_levels_pc_type", "jacobi") != "jacobi")
      return xsb_fail(c, XSB_ERR_SUP, "MG smoother must be chebyshev/jacobi");
    o.has("saddle_fieldsplit_u_pc_mg_galerkin"); o.has("saddle_fieldsplit_u_mg_levels_ksp_norm_type"); o.has("saddle_fieldsplit_u_mg_coarse_pc_factor_mat_solver_type");
  } else {
    // the reference sets no PC type here (exSaddle.c:303-402 only does for -fs / -mg), so PETSc's default applies: ilu on one
    // rank, bjacobi + ilu on several.  That default is not implemented: say so instead of silently running unpreconditioned.
    if (!o.has("saddle_pc_type")) return xsb_fail(c, XSB_ERR_SUP, "no -saddle_pc_type given: PETSc's default (ilu / bjacobi+ilu on the saddle matrix) is not implemented; pass -saddle_pc_type jacobi|none, -fs or -mg");
    const std::string pc = o.str("saddle_pc_type", "none");
    if (pc == "jacobi") s.pc_type = 1; else if (pc == "none") s.pc_type = 0;
    else if (pc == "asm") {
      // Makefile:298, 411: one element patch per rank (DMCreateDomainDecomposition of the KSP's DM) with exact sub-solves.  Without
      // -saddle_pc_asm_dm_subdomains / -set_ksp_dm PETSc would cut subdomains from the matrix graph instead: not implemented.
      s.pc_type = 5;
      if (!o.flag("saddle_pc_asm_dm_subdomains") || !o.flag("set_ksp_dm")) return xsb_fail(c, XSB_ERR_SUP, "-saddle_pc_type asm is supported on the reference's element patches: add -saddle_pc_asm_dm_subdomains -set_ksp_dm");
      if (o.str("saddle_sub_pc_type", "ilu") != "lu" || o.str("saddle_sub_ksp_type", "preonly") != "preonly") return xsb_fail(c, XSB_ERR_SUP, "ASM sub-solves: -saddle_sub_ksp_type preonly -saddle_sub_pc_type lu");
      o.has("saddle_sub_pc_factor_mat_solver_type"); o.has("dmdafe_overlap"); o.has("xsb_ranks");
    }
    else return xsb_fail(c, XSB_ERR_SUP, "-saddle_pc_type %s not supported without -fs (jacobi|none|asm)", pc.c_str());
  }
  s.right = (o.str("saddle_ksp_pc_side", ksp == "fgmres" ? "right" : "left") == "right") || s.ksp_type == 1;
  s.rtol = o.real("saddle_ksp_rtol", 1e-5); s.atol = o.real("saddle_ksp_atol", 1e-50); s.dtol = o.real("saddle_ksp_divtol", 1e4);
  s.max_it = o.integer("saddle_ksp_max_it", 10000); s.restart = o.integer("saddle_ksp_gmres_restart", 30);
  s.u_rtol = o.real("saddle_fieldsplit_u_ksp_rtol", 1e-5); s.u_max_it = o.integer("saddle_fieldsplit_u_ksp_max_it", 10000);
  s.u_restart = o.integer("saddle_fieldsplit_u_ksp_gcr_restart", 30);
  s.mg_levels = o.integer("saddle_fieldsplit_u_pc_mg_levels", 1);
  s.cheb_its = o.integer("saddle_fieldsplit_u_mg_levels_ksp_max_it", 2);
  if (o.has("saddle_fieldsplit_u_mg_levels_ksp_chebyshev_esteig")) {
    double v[4] = {0, 0.1, 0, 1.1};
    if (sscanf(o.kv["saddle_fieldsplit_u_mg_levels_ksp_chebyshev_esteig"].c_str(), "%lf,%lf,%lf,%lf", &v[0], &v[1], &v[2], &v[3]) != 4)
      return xsb_fail(c, XSB_ERR_ARG, "-..._ksp_chebyshev_esteig needs a,b,c,d");
    for (int i = 0; i < 4; ++i) s.esteig[i] = v[i];
  }
  s.esteig_steps = o.integer("saddle_fieldsplit_u_mg_levels_ksp_chebyshev_esteig_steps", 10);
  if (s.esteig_steps < 1 || s.esteig_steps > 60) return xsb_fail(c, XSB_ERR_ARG, "-..._ksp_chebyshev_esteig_steps must be in [1,60]");   // the dot-product scratch holds 128 scalars
  if (s.pc_type == 2 && s.cheb_its < 1) return xsb_fail(c, XSB_ERR_ARG, "-saddle_fieldsplit_u_mg_levels_ksp_max_it must be >= 1");
  s.noise = o.integer("xsb_chebyshev_noise", 0);
  s.n_cheb_fixed = 0;
  for (int l = 1; l < s.mg_levels; ++l) {
    char key[128]; snprintf(key, sizeof(key), "saddle_fieldsplit_u_mg_levels_%d_ksp_chebyshev_eigenvalues", l);
    if (o.has(key)) {
      if (sscanf(o.kv[key].c_str(), "%lf,%lf", &s.cheb_emin[l - 1], &s.cheb_emax[l - 1]) != 2) return xsb_fail(c, XSB_ERR_ARG, "-%s needs emin,emax", key);
      s.n_cheb_fixed++;
    }
  }
  if (s.n_cheb_fixed && s.n_cheb_fixed != s.mg_levels - 1) return xsb_fail(c, XSB_ERR_ARG, "explicit Chebyshev eigenvalues must be given for every MG level");
  const std::string ppc = o.str("saddle_fieldsplit_p_pc_type", "bjacobi");
  if (ppc == "bjacobi" || ppc == "ilu") s.p_pc = 0; else if (ppc == "jacobi") s.p_pc = 1;
  else return xsb_fail(c, XSB_ERR_SUP, "-saddle_fieldsplit_p_pc_type %s not supported (bjacobi|ilu|jacobi)", ppc.c_str());
  s.time_kernels = o.flag("xsb_time_kernels");
  s.mf_grad = o.integer("xsb_mf_grad", c->no_A ? 1 : 0);
  c->use_graph = o.integer("xsb_graph", 1);
  { const std::string mfv = o.str("xsb_matrix_free", "0");   // flag: fine-level products by the element kernel; "full": A / A00 never stored
    s.matrix_free = (mfv.empty() || mfv == "1" || mfv == "true" || mfv == "yes" || mfv == "full" || mfv == "2") ? 1 : 0;
    if (c->no_A && !s.matrix_free) return xsb_fail(c, XSB_ERR_ORDER, "the operator was assembled with -xsb_matrix_free full; the option cannot be dropped before xsb_ksp_setup");
    if (!c->no_A && (mfv == "full" || mfv == "2")) return xsb_fail(c, XSB_ERR_ORDER, "-xsb_matrix_free full must be set before xsb_assemble"); }
  if (c->no_A && s.pc_type != 2) return xsb_fail(c, XSB_ERR_SUP, "-xsb_matrix_free full supports the -fs (ABF) solver tree only");
  if (s.restart < 1 || s.restart > 60 || s.u_restart < 1 || s.u_restart > 60) return xsb_fail(c, XSB_ERR_ARG, "restart must be in [1,60]");
  // monitor / view flags of the reference's command lines are accepted and handled by the caller
  o.has("saddle_ksp_monitor_short"); o.has("saddle_ksp_converged_reason"); o.has("saddle_ksp_view"); o.has("diagnostics");
  o.has("options_left"); o.has("saddle_fieldsplit_u_ksp_converged_reason"); o.has("twosolves");
  o.has("dump_solution"); o.has("dump_operator"); o.has("dump_scaled_mass_matrix"); o.has("view_fields");   // written by the caller through xsb_dump_operator / xsb_dump_vector
  return 0;
}

// Releases everything xsb_ksp_setup and the solves allocated (phase-1 allocations): MG hierarchy, Galerkin levels, ILU
// factors, Krylov bases, work vectors.  The assembled operator (phase 0) and the element-kernel state (phase 2) stay.
int ksp_release(xsb_ctx c)
{
  CUDA_OK(cudaStreamSynchronize(c->stream));
  mg_graphs_release(c);
  mmg_free(c); fsd_free(c);
  if (c->asmpc) { asm_free(c->asmpc); c->asmpc = nullptr; }
  dev_free_phase(c, 1);
  c->red = c->scal = nullptr; c->w_t1 = c->w_t2 = c->xdev = c->bdev = c->idiagA = c->gcr_r = c->fs_tu = nullptr;
  c->mp_lu = c->mp_idiag = nullptr; c->mp_block_a = nullptr; c->ilu_rows = c->ilu_lvl_off = c->ilu_diag = c->ilu_fcol = c->ilu_bcol = nullptr;
  c->ilu_fval = c->ilu_bval = c->ilu_binv = nullptr; c->ilu_fn = c->ilu_bn = nullptr; c->MpOwn = Csr();
  c->ilup_on = false; c->ilup_packf = c->ilup_packb = nullptr; c->ilup_prog = nullptr; c->ilup_y = nullptr;
  c->V.clear(); c->Z.clear(); c->GV.clear(); c->GS.clear();
  for (int l = 0; l < XSB_MAX_LEVELS; ++l) { c->lev[l] = Level(); c->sub[l] = Level(); }
  c->nlev = 0; c->nsub = 0; c->cg_p = c->cg_q = nullptr; c->ksp_ready = false;
  return 0;
}

int ksp_setup(xsb_ctx c)
{
  if (!c->assembled) return xsb_fail(c, XSB_ERR_ORDER, "xsb_ksp_setup called before xsb_assemble");
  XSB_CHK(read_solver_options(c));
  const Lattice &L = c->lat;
  // Re-entry (an option changed since the last set-up): everything the previous set-up allocated is released first
  XSB_CHK(ksp_release(c));
  c->phase = 1;
  struct PhaseGuard { xsb_ctx c; ~PhaseGuard() { c->phase = 0; } } guard{c};
  CUDA_OK(cudaEventRecord(c->ev0, c->stream));
  XSB_CHK(dev_alloc(c, &c->red, (size_t)592 * 8)); XSB_CHK(dev_alloc(c, &c->scal, 256));
  if (!c->red_h) CUDA_OK(cudaMallocHost(&c->red_h, sizeof(double) * 256));
  XSB_CHK(dev_alloc(c, &c->w_t1, (size_t)L.n)); XSB_CHK(dev_alloc(c, &c->w_t2, (size_t)L.n)); XSB_CHK(dev_alloc(c, &c->xdev, (size_t)L.n)); XSB_CHK(dev_alloc(c, &c->bdev, (size_t)L.n));
  if (c->so.pc_type == 1) { XSB_CHK(dev_alloc(c, &c->idiagA, (size_t)L.n)); XSB_CHK(csr_diag_inv(c, c->A, c->idiagA)); }
  if (c->so.pc_type == 3) XSB_CHK(mmg_setup(c));
  if (c->so.pc_type == 4) XSB_CHK(fsd_setup(c));
  if (c->so.pc_type == 5) XSB_CHK(asm_setup(c, c, c->opt.integer("xsb_ranks", 1), c->opt.integer("dmdafe_overlap", 0), &c->asmpc));
  if (c->so.pc_type == 2) {
    if (c->so.matrix_free) XSB_CHK(mf_setup(c));
    if (c->so.mf_grad) XSB_CHK(grad_prepare(c));
    XSB_CHK(mg_setup(c));
    if (c->so.p_pc == 0) XSB_CHK(ilu_setup(c));
    else { XSB_CHK(dev_alloc(c, &c->mp_idiag, (size_t)L.np)); XSB_CHK(csr_diag_inv(c, c->Mp, c->mp_idiag)); }
    XSB_CHK(dev_alloc(c, &c->gcr_r, (size_t)L.nu)); XSB_CHK(dev_alloc(c, &c->fs_tu, (size_t)L.nu));
  }
  CUDA_OK(cudaEventRecord(c->ev1, c->stream)); CUDA_OK(cudaEventSynchronize(c->ev1));
  CUDA_OK(cudaEventElapsedTime(&c->setup_ms, c->ev0, c->ev1));
  c->ksp_ready = true;
  return 0;
}

// ------------------------------------------------------------------ KSPSolve_GMRES / KSPSolve_FGMRES
int ksp_solve(xsb_ctx c, const double *b, double *x)
{
  if (!c->ksp_ready) return xsb_fail(c, XSB_ERR_ORDER, "xsb_ksp_solve called before xsb_ksp_setup");
  const SolverOpts &s = c->so; const int64_t n = c->lat.n; const int m = s.restart;
  const bool flex = s.ksp_type == 1, right = s.right, haspc = s.pc_type != 0;
  if (flex && !haspc) return xsb_fail(c, XSB_ERR_SUP, "fgmres without a preconditioner");
  std::vector<double> hh((size_t)(m + 1) * m), cs(m + 1), sn(m + 1), rs(m + 1), y(m + 1), hcol(m + 2);
  double *t1 = c->w_t1, *t2 = c->w_t2;
  c->its = 0; c->reason = 0; c->hist.clear(); c->inner_its.clear(); c->inner_reason.clear();
  c->n_a00 = c->n_a = 0; c->a00_ns_sum = 0; c->a00_timed = 0; c->ev_used = 0; for (int i = 0; i < 4; ++i) c->a00_mode[i] = 0;
  const int64_t launch0 = c->n_launch;
  auto need = [&](std::vector<double *> &W, int k) -> int { while ((int)W.size() <= k) { double *p = nullptr; c->phase = 1; int rc = dev_alloc(c, &p, (size_t)n); c->phase = 0; if (rc) return rc; W.push_back(p); } return 0; };
  CUDA_OK(cudaEventRecord(c->ev0, c->stream));
  XSB_CHK(prof_mark(c, PROF_OTHER));
  if (!b) b = c->F;
  XSB_CHK(vec_set(c, n, 0.0, x));          // initial guess is zero
  XSB_CHK(need(c->V, 0));
  double rnorm0 = 0.0, ttol = 0.0; bool first = true;
  while (!c->reason) {
    double res;
    // KSPInitialResidual: r = b - A x (left PC: M^-1 r)
    if (first) XSB_CHK(vec_copy(c, n, b, t1));   // x = 0: r = b (bitwise)
    else { XSB_CHK(full_mult(c, x, t1)); XSB_CHK(vec_aypx(c, n, -1.0, b, t1)); }
    first = false;
    if (!flex && !right && haspc) XSB_CHK(pc_apply(c, t1, c->V[0], nullptr)); else XSB_CHK(vec_copy(c, n, t1, c->V[0]));
    XSB_CHK(vec_mdot(c, c->own_full, c->V[0], nullptr, 0, true, c->scal));
    XSB_CHK(vec_fetch(c, c->scal, 1, hcol.data()));
    res = sqrt(hcol[0]);
    if (c->its == 0) { rnorm0 = res; ttol = fmax(s.rtol * rnorm0, s.atol); }
    if ((int)c->hist.size() == c->its) c->hist.push_back(res); else c->hist[c->its] = res;
    if (res == 0.0) { c->reason = 3; break; }
    if (res <= ttol) { c->reason = (s.atol >= s.rtol * rnorm0) ? 3 : 2; break; }
    if (c->its >= s.max_it) { c->reason = -3; break; }
    XSB_CHK(vec_scale(c, n, 1.0 / res, c->V[0]));
    rs[0] = res;
    int it = 0;
    while (!c->reason && it < m && c->its < s.max_it) {
      XSB_CHK(need(c->V, it + 1));
      double *w = c->V[it + 1];
      if (flex) {   // z_j = M^-1 v_j ; w = A z_j
        XSB_CHK(need(c->Z, it));
        int inner = 0, why = 0; XSB_CHK(pc_apply(c, c->V[it], c->Z[it], &inner, &why));
        if (s.pc_type == 2) { c->inner_its.push_back(inner); c->inner_reason.push_back(why); }
        XSB_CHK(full_mult(c, c->Z[it], w));
      } else if (right) {
        if (haspc) { XSB_CHK(pc_apply(c, c->V[it], t2, nullptr)); XSB_CHK(full_mult(c, t2, w)); } else XSB_CHK(full_mult(c, c->V[it], w));
      } else {
        if (haspc) { XSB_CHK(full_mult(c, c->V[it], t2)); XSB_CHK(pc_apply(c, t2, w, nullptr)); } else XSB_CHK(full_mult(c, c->V[it], w));
      }
      // classical Gram-Schmidt: h = V^T w ; w -= V h ; ||w||   (coefficients never leave the device)
      XSB_CHK(vec_mdot(c, c->own_full, w, c->V.data(), it + 1, false, c->scal));
      XSB_CHK(vec_maxpy_dev(c, n, w, c->V.data(), it + 1, c->scal, -1.0));
      XSB_CHK(vec_mdot(c, c->own_full, w, nullptr, 0, true, c->scal + it + 1));
      XSB_CHK(vec_scale_by_inv_sqrt(c, n, w, c->scal + it + 1));
      XSB_CHK(vec_fetch(c, c->scal, it + 2, hcol.data()));
      hcol[it + 1] = sqrt(hcol[it + 1]);
      // KSPGMRESUpdateHessenberg
      for (int j = 0; j < it; ++j) { double t = hcol[j]; hcol[j] = cs[j] * t + sn[j] * hcol[j + 1]; hcol[j + 1] = -sn[j] * t + cs[j] * hcol[j + 1]; }
      const double tt = sqrt(hcol[it] * hcol[it] + hcol[it + 1] * hcol[it + 1]);
      if (tt == 0.0) { c->reason = -5; break; }
      cs[it] = hcol[it] / tt; sn[it] = hcol[it + 1] / tt;
      rs[it + 1] = -sn[it] * rs[it]; rs[it] = cs[it] * rs[it];
      hcol[it] = cs[it] * hcol[it] + sn[it] * hcol[it + 1]; hcol[it + 1] = 0.0;
      res = fabs(rs[it + 1]);
      for (int j = 0; j <= it; ++j) hh[(size_t)it * (m + 1) + j] = hcol[j];
      it++; c->its++;
      if ((int)c->hist.size() == c->its) c->hist.push_back(res); else c->hist[c->its] = res;
      if (res <= ttol) c->reason = (s.atol >= s.rtol * rnorm0) ? 3 : 2;
      else if (res >= s.dtol * rnorm0) c->reason = -4;
    }
    // KSPGMRESBuildSoln
    if (it > 0) {
      for (int k = it - 1; k >= 0; --k) { double t = rs[k]; for (int j = k + 1; j < it; ++j) t -= hh[(size_t)j * (m + 1) + k] * y[j]; y[k] = t / hh[(size_t)k * (m + 1) + k]; }
      if (flex) XSB_CHK(vec_maxpy_host(c, n, x, c->Z.data(), it, y.data()));
      else if (right && haspc) { XSB_CHK(vec_set(c, n, 0.0, t1)); XSB_CHK(vec_maxpy_host(c, n, t1, c->V.data(), it, y.data())); XSB_CHK(pc_apply(c, t1, t2, nullptr)); XSB_CHK(vec_axpy(c, n, 1.0, t2, x)); }
      else XSB_CHK(vec_maxpy_host(c, n, x, c->V.data(), it, y.data()));
    }
    if (!c->reason && c->its >= s.max_it) c->reason = -3;
  }
  XSB_CHK(prof_mark(c, PROF_OTHER));
  CUDA_OK(cudaEventRecord(c->ev1, c->stream)); CUDA_OK(cudaEventSynchronize(c->ev1));
  CUDA_OK(cudaEventElapsedTime(&c->solve_ms, c->ev0, c->ev1));
  c->solve_launches = c->n_launch - launch0;
  XSB_CHK(spmv_collect_timing(c));
  if (c->ilup_on) { unsigned e = 0; CUDA_OK(cudaMemcpyAsync(&e, c->ilup_prog, sizeof(e), cudaMemcpyDeviceToHost, c->stream)); CUDA_OK(cudaStreamSynchronize(c->stream));
    if (e) return xsb_fail(c, XSB_ERR_BREAKDOWN, "pipelined ILU(0) solve: a plane waited more than 2 s for the plane below (NaN right-hand side?)"); }
  return comm_p2p_check(c);
}
