// xsb_fs.cu -- the reference's plain `-fs` solver tree, i.e. PCFIELDSPLIT with PETSc's DEFAULT sub-solvers
// (exSaddle.c:303-322 when no abf.opts is given; goldens exSaddle{2d,3d}[_lame]_fs_{1,2}, Makefile:282, 347, 396, 480):
//
//   outer   GMRES(30), left PC, preconditioned norm                                         (xsb_ksp.cu)
//   PC      fieldsplit Schur / UPPER / user Mpscaled:
//             y_p = KSP_p[S, ILU0(Mpscaled)] x_p ,   S v = A11 v - A10 KSP_u[A00](A01 v)     (nested solve per S apply)
//             y_u = KSP_u[A00] (x_u - A01 y_p)
//   KSP_u   GMRES(30) + ILU(0) of A00 in natural ordering (PETSc's default PC for seqaij), rtol 1e-5
//   KSP_p   GMRES(30) + ILU(0) of Mpscaled, rtol 1e-5   (-saddle_fieldsplit_p_ksp_type preonly: the ILU alone)
//
// ILU(0) of A00 is a general sparse factorisation: rows are scheduled by dependency level (longest path in the pattern's DAG,
// integer work done once on the host from the pattern), one CTA sweeps the levels with a block barrier in between; every row
// performs exactly the operations of MatLUFactorNumeric_SeqAIJ / MatSolve_SeqAIJ in their order (IKJ, ascending columns;
// backward sweep descending), so the factors equal the sequential ones.  This tree is for the reference's small regression
// cases (2-D 6^2, 3-D 4^3): the production path is the ABF tree (GCR + GMG), which needs no factorisation of A00.
#include "xsb.h"
#include <functional>

namespace {
struct GIlu {
  int n = 0; int *ia = nullptr, *ja = nullptr, *diag = nullptr; double *lu = nullptr;
  int nlf = 0, nlb = 0; int *foff = nullptr, *frows = nullptr, *boff = nullptr, *brows = nullptr;   // forward / backward level schedules
};
struct Fsd {
  Csr A00s;                 // scalar CSR copy of A00 (the MATSEQAIJ sub-matrix PETSc factors)
  GIlu ilu_u;
  std::vector<double *> Vu, Vp; double *u_t1 = nullptr, *u_t2 = nullptr, *u_rhs = nullptr, *u_sol = nullptr, *p_t1 = nullptr, *p_t2 = nullptr, *p_tmp = nullptr;
  int u_max_it = 10000, p_preonly = 0; double u_rtol = 1e-5, p_rtol = 1e-5; int p_max_it = 10000;
  int64_t n_ksp_u = 0;
};

// scalar CSR rows of a BAIJ matrix (row x of node: its blocks' row x, block after block)
template <int BS>
__global__ void k_baij_scalar(int nb, const int *__restrict__ bia, const int *__restrict__ bja, const double *__restrict__ ba, int *ia, int *ja, double *a)
{
  const int node = blockIdx.x * blockDim.x + threadIdx.x; if (node > nb) return;
  if (node == nb) { ia[BS * nb] = BS * BS * bia[nb]; return; }
  const int b0 = bia[node], nbk = bia[node + 1] - b0;
  for (int x = 0; x < BS; ++x) {
    const int r0 = BS * BS * b0 + x * BS * nbk;
    ia[BS * node + x] = r0;
    for (int s = 0; s < nbk; ++s) for (int y = 0; y < BS; ++y) { ja[r0 + s * BS + y] = BS * bja[b0 + s] + y; a[r0 + s * BS + y] = ba[(int64_t)(b0 + s) * BS * BS + x * BS + y]; }
  }
}

// MatLUFactorNumeric_SeqAIJ restricted to the pattern, level by level (one CTA; rows of a level are independent)
__global__ void __launch_bounds__(1024) k_gilu_factor(int nlev, const int *__restrict__ off, const int *__restrict__ rows, const int *__restrict__ ia, const int *__restrict__ ja,
                                                      const int *__restrict__ diag, double *lu, int *flag)
{
  for (int lv = 0; lv < nlev; ++lv) {
    for (int q = off[lv] + threadIdx.x; q < off[lv + 1]; q += blockDim.x) {
      const int i = rows[q], r0 = ia[i], r1 = ia[i + 1], di = diag[i];
      for (int k = r0; k < di; ++k) {                   // lower entries in ascending column order
        const int r = ja[k]; double mult = lu[k];
        if (mult != 0.0) {
          mult = mult * lu[diag[r]]; lu[k] = mult;      // lu[diag] = 1 / pivot
          int p = k + 1;                                // both rows are sorted: merge the pivot row's upper part into row i
          for (int l = diag[r] + 1; l < ia[r + 1]; ++l) {
            const int col = ja[l];
            while (p < r1 && ja[p] < col) ++p;
            if (p < r1 && ja[p] == col) lu[p] -= mult * lu[l];   // fill outside the pattern is dropped (ILU(0))
          }
        }
      }
      if (lu[di] == 0.0) *flag = 1;
      lu[di] = 1.0 / lu[di];
    }
    __syncthreads();
  }
}
// MatSolve_SeqAIJ: forward with unit L (ascending columns), backward with U (descending columns) and the inverted diagonal
__global__ void __launch_bounds__(1024) k_gilu_solve(int nlf, const int *__restrict__ foff, const int *__restrict__ frows, int nlb, const int *__restrict__ boff, const int *__restrict__ brows,
                                                     const int *__restrict__ ia, const int *__restrict__ ja, const int *__restrict__ diag, const double *__restrict__ lu,
                                                     const double *__restrict__ b, double *x)
{
  for (int lv = 0; lv < nlf; ++lv) {
    for (int q = foff[lv] + threadIdx.x; q < foff[lv + 1]; q += blockDim.x) {
      const int i = frows[q]; double s = b[i];
      for (int k = ia[i]; k < diag[i]; ++k) s -= lu[k] * x[ja[k]];
      x[i] = s;
    }
    __syncthreads();
  }
  for (int lv = 0; lv < nlb; ++lv) {
    for (int q = boff[lv] + threadIdx.x; q < boff[lv + 1]; q += blockDim.x) {
      const int i = brows[q]; double s = x[i];
      for (int k = ia[i + 1] - 1; k > diag[i]; --k) s -= lu[k] * x[ja[k]];
      x[i] = s * lu[diag[i]];
    }
    __syncthreads();
  }
}

int gilu_setup(xsb_ctx c, const Csr &M, GIlu &I)
{
  cudaStream_t st = c->stream; const int n = M.n;
  if (M.nnz > 200000000) return xsb_fail(c, XSB_ERR_SUP, "ILU(0) of A00 (the default -fs tree) is meant for the reference's small cases; use the abf.opts tree at this size");
  I.n = n; I.ia = M.ia; I.ja = M.ja;
  std::vector<int> ia(n + 1), ja(M.nnz);
  CUDA_OK(cudaStreamSynchronize(st));
  CUDA_OK(cudaMemcpy(ia.data(), M.ia, sizeof(int) * (n + 1), cudaMemcpyDeviceToHost));
  CUDA_OK(cudaMemcpy(ja.data(), M.ja, sizeof(int) * M.nnz, cudaMemcpyDeviceToHost));
  // integer analysis of the pattern (host): diagonal positions, forward / backward dependency levels
  std::vector<int> diag(n), lf(n), lb(n);
  int nlf = 0, nlb = 0;
  for (int i = 0; i < n; ++i) {
    int d = -1, l = 0;
    for (int k = ia[i]; k < ia[i + 1]; ++k) { if (ja[k] == i) d = k; else if (ja[k] < i && lf[ja[k]] + 1 > l) l = lf[ja[k]] + 1; }
    if (d < 0) return xsb_fail(c, XSB_ERR_BREAKDOWN, "ILU(0): row %d has no diagonal entry", i);
    diag[i] = d; lf[i] = l; if (l + 1 > nlf) nlf = l + 1;
  }
  for (int i = n - 1; i >= 0; --i) {
    int l = 0;
    for (int k = diag[i] + 1; k < ia[i + 1]; ++k) if (lb[ja[k]] + 1 > l) l = lb[ja[k]] + 1;
    lb[i] = l; if (l + 1 > nlb) nlb = l + 1;
  }
  auto schedule = [&](const std::vector<int> &lev, int nl, std::vector<int> &off, std::vector<int> &rows) {
    off.assign(nl + 1, 0); rows.resize(n);
    for (int i = 0; i < n; ++i) off[lev[i] + 1]++;
    for (int l = 0; l < nl; ++l) off[l + 1] += off[l];
    std::vector<int> cur(off.begin(), off.end() - 1);
    for (int i = 0; i < n; ++i) rows[cur[lev[i]]++] = i;
  };
  std::vector<int> foff, frows, boff, brows; schedule(lf, nlf, foff, frows); schedule(lb, nlb, boff, brows);
  I.nlf = nlf; I.nlb = nlb;
  XSB_CHK(dev_alloc(c, &I.diag, (size_t)n)); XSB_CHK(dev_alloc(c, &I.lu, (size_t)M.nnz));
  XSB_CHK(dev_alloc(c, &I.foff, (size_t)nlf + 1)); XSB_CHK(dev_alloc(c, &I.frows, (size_t)n)); XSB_CHK(dev_alloc(c, &I.boff, (size_t)nlb + 1)); XSB_CHK(dev_alloc(c, &I.brows, (size_t)n));
  // uploads on the handle's stream (ordered with the kernels that read them; the host vectors outlive the synchronisation below)
  CUDA_OK(cudaMemcpyAsync(I.diag, diag.data(), sizeof(int) * n, cudaMemcpyHostToDevice, st));
  CUDA_OK(cudaMemcpyAsync(I.foff, foff.data(), sizeof(int) * (nlf + 1), cudaMemcpyHostToDevice, st)); CUDA_OK(cudaMemcpyAsync(I.frows, frows.data(), sizeof(int) * n, cudaMemcpyHostToDevice, st));
  CUDA_OK(cudaMemcpyAsync(I.boff, boff.data(), sizeof(int) * (nlb + 1), cudaMemcpyHostToDevice, st)); CUDA_OK(cudaMemcpyAsync(I.brows, brows.data(), sizeof(int) * n, cudaMemcpyHostToDevice, st));
  CUDA_OK(cudaMemcpyAsync(I.lu, M.a, sizeof(double) * M.nnz, cudaMemcpyDeviceToDevice, st));
  int *flag = nullptr; XSB_CHK(dev_alloc(c, &flag, 1));
  k_gilu_factor<<<1, 1024, 0, st>>>(nlf, I.foff, I.frows, I.ia, I.ja, I.diag, I.lu, flag); KERNEL_OK();
  int h = 0; CUDA_OK(cudaMemcpyAsync(&h, flag, sizeof(int), cudaMemcpyDeviceToHost, st)); CUDA_OK(cudaStreamSynchronize(st));
  if (h) return xsb_fail(c, XSB_ERR_BREAKDOWN, "zero pivot in ILU(0) of A00");
  return 0;
}
int gilu_apply(xsb_ctx c, const GIlu &I, const double *b, double *x)
{
  k_gilu_solve<<<1, 1024, 0, c->stream>>>(I.nlf, I.foff, I.frows, I.nlb, I.boff, I.brows, I.ia, I.ja, I.diag, I.lu, b, x); KERNEL_OK();
  return 0;
}

// KSPSolve_GMRES from a zero initial guess: left PC, preconditioned norm, classical Gram-Schmidt without refinement, Givens QR,
// restart 30 -- the inner solver of the default tree; same arithmetic as the outer loop of xsb_ksp.cu.
typedef std::function<int(const double *, double *)> Op;
int gmres_left(xsb_ctx c, int64_t n, const Op &A, const Op &M, const double *b, double *x, double rtol, int max_it, int m,
               std::vector<double *> &V, double *t1, double *t2, double *scal, int *its_out)
{
  const Ranges rg = whole(n);
  std::vector<double> hh((size_t)(m + 1) * m), cs(m + 1), sn(m + 1), rs(m + 1), y(m + 1), hcol(m + 2);
  auto need = [&](int k) -> int { while ((int)V.size() <= k) { double *p = nullptr; c->phase = 1; int rc = dev_alloc(c, &p, (size_t)n); c->phase = 0; if (rc) return rc; V.push_back(p); } return 0; };
  int its = 0, reason = 0; double rnorm0 = 0.0, ttol = 0.0; bool first = true;
  XSB_CHK(vec_set(c, n, 0.0, x)); XSB_CHK(need(0));
  while (!reason) {
    if (first) XSB_CHK(M(b, V[0]));                                   // x = 0: M^-1 b
    else { XSB_CHK(A(x, t1)); XSB_CHK(vec_aypx(c, n, -1.0, b, t1)); XSB_CHK(M(t1, V[0])); }
    first = false;
    XSB_CHK(vec_mdot(c, rg, V[0], nullptr, 0, true, scal, true));
    XSB_CHK(vec_fetch(c, scal, 1, hcol.data()));
    double res = sqrt(hcol[0]);
    if (its == 0) { rnorm0 = res; ttol = fmax(rtol * rnorm0, 1e-50); }
    if (res == 0.0 || res <= ttol) { reason = 2; break; }
    if (its >= max_it) { reason = -3; break; }
    XSB_CHK(vec_scale(c, n, 1.0 / res, V[0]));
    rs[0] = res;
    int it = 0;
    while (!reason && it < m && its < max_it) {
      XSB_CHK(need(it + 1));
      double *w = V[it + 1];
      XSB_CHK(A(V[it], t2)); XSB_CHK(M(t2, w));
      XSB_CHK(vec_mdot(c, rg, w, V.data(), it + 1, false, scal, true));
      XSB_CHK(vec_maxpy_dev(c, n, w, V.data(), it + 1, scal, -1.0));
      XSB_CHK(vec_mdot(c, rg, w, nullptr, 0, true, scal + it + 1, true));
      XSB_CHK(vec_scale_by_inv_sqrt(c, n, w, scal + it + 1));
      XSB_CHK(vec_fetch(c, scal, it + 2, hcol.data()));
      hcol[it + 1] = sqrt(hcol[it + 1]);
      for (int j = 0; j < it; ++j) { double t = hcol[j]; hcol[j] = cs[j] * t + sn[j] * hcol[j + 1]; hcol[j + 1] = -sn[j] * t + cs[j] * hcol[j + 1]; }
      const double tt = sqrt(hcol[it] * hcol[it] + hcol[it + 1] * hcol[it + 1]);
      if (tt == 0.0) { reason = -5; break; }
      cs[it] = hcol[it] / tt; sn[it] = hcol[it + 1] / tt;
      rs[it + 1] = -sn[it] * rs[it]; rs[it] = cs[it] * rs[it];
      hcol[it] = cs[it] * hcol[it] + sn[it] * hcol[it + 1]; hcol[it + 1] = 0.0;
      res = fabs(rs[it + 1]);
      for (int j = 0; j <= it; ++j) hh[(size_t)it * (m + 1) + j] = hcol[j];
      it++; its++;
      if (res <= ttol) reason = 2; else if (res >= 1e4 * rnorm0) reason = -4; else if (its >= max_it) reason = -3;
    }
    if (it > 0) {
      for (int k = it - 1; k >= 0; --k) { double t = rs[k]; for (int j = k + 1; j < it; ++j) t -= hh[(size_t)j * (m + 1) + k] * y[j]; y[k] = t / hh[(size_t)k * (m + 1) + k]; }
      CUDA_OK(cudaMemcpyAsync(scal + 64, y.data(), sizeof(double) * it, cudaMemcpyHostToDevice, c->stream));
      CUDA_OK(cudaStreamSynchronize(c->stream));   // y is a local of this frame
      XSB_CHK(vec_maxpy_dev(c, n, x, V.data(), it, scal + 64, 1.0));
    }
  }
  if (its_out) *its_out = its;
  return 0;
}

// KSPSolve_FGMRES from a zero initial guess: right (flexible) PC, unpreconditioned norm, classical Gram-Schmidt, restart m (App. B.5)
int fgmres_right(xsb_ctx c, int64_t n, const Op &A, const Op &M, const double *b, double *x, double rtol, int max_it, int m,
                 std::vector<double *> &V, std::vector<double *> &Z, double *t1, double *scal, int *its_out)
{
  const Ranges rg = whole(n);
  std::vector<double> hh((size_t)(m + 1) * m), cs(m + 1), sn(m + 1), rs(m + 1), y(m + 1), hcol(m + 2);
  auto need = [&](std::vector<double *> &W, int k) -> int { while ((int)W.size() <= k) { double *p = nullptr; c->phase = 1; int rc = dev_alloc(c, &p, (size_t)n); c->phase = 0; if (rc) return rc; W.push_back(p); } return 0; };
  int its = 0, reason = 0; double rnorm0 = 0.0, ttol = 0.0; bool first = true;
  XSB_CHK(vec_set(c, n, 0.0, x)); XSB_CHK(need(V, 0));
  while (!reason) {
    if (first) XSB_CHK(vec_copy(c, n, b, V[0]));
    else { XSB_CHK(A(x, t1)); XSB_CHK(vec_aypx(c, n, -1.0, b, t1)); XSB_CHK(vec_copy(c, n, t1, V[0])); }
    first = false;
    XSB_CHK(vec_mdot(c, rg, V[0], nullptr, 0, true, scal, true));
    XSB_CHK(vec_fetch(c, scal, 1, hcol.data()));
    double res = sqrt(hcol[0]);
    if (its == 0) { rnorm0 = res; ttol = fmax(rtol * rnorm0, 1e-50); }
    if (res == 0.0 || res <= ttol) { reason = 2; break; }
    if (its >= max_it) { reason = -3; break; }
    XSB_CHK(vec_scale(c, n, 1.0 / res, V[0]));
    rs[0] = res;
    int it = 0;
    while (!reason && it < m && its < max_it) {
      XSB_CHK(need(V, it + 1)); XSB_CHK(need(Z, it));
      double *w = V[it + 1];
      XSB_CHK(M(V[it], Z[it])); XSB_CHK(A(Z[it], w));
      XSB_CHK(vec_mdot(c, rg, w, V.data(), it + 1, false, scal, true));
      XSB_CHK(vec_maxpy_dev(c, n, w, V.data(), it + 1, scal, -1.0));
      XSB_CHK(vec_mdot(c, rg, w, nullptr, 0, true, scal + it + 1, true));
      XSB_CHK(vec_scale_by_inv_sqrt(c, n, w, scal + it + 1));
      XSB_CHK(vec_fetch(c, scal, it + 2, hcol.data()));
      hcol[it + 1] = sqrt(hcol[it + 1]);
      for (int j = 0; j < it; ++j) { double t = hcol[j]; hcol[j] = cs[j] * t + sn[j] * hcol[j + 1]; hcol[j + 1] = -sn[j] * t + cs[j] * hcol[j + 1]; }
      const double tt = sqrt(hcol[it] * hcol[it] + hcol[it + 1] * hcol[it + 1]);
      if (tt == 0.0) { reason = -5; break; }
      cs[it] = hcol[it] / tt; sn[it] = hcol[it + 1] / tt;
      rs[it + 1] = -sn[it] * rs[it]; rs[it] = cs[it] * rs[it];
      hcol[it] = cs[it] * hcol[it] + sn[it] * hcol[it + 1]; hcol[it + 1] = 0.0;
      res = fabs(rs[it + 1]);
      for (int j = 0; j <= it; ++j) hh[(size_t)it * (m + 1) + j] = hcol[j];
      it++; its++;
      if (res <= ttol) reason = 2; else if (res >= 1e4 * rnorm0) reason = -4; else if (its >= max_it) reason = -3;
    }
    if (it > 0) {
      for (int k = it - 1; k >= 0; --k) { double t = rs[k]; for (int j = k + 1; j < it; ++j) t -= hh[(size_t)j * (m + 1) + k] * y[j]; y[k] = t / hh[(size_t)k * (m + 1) + k]; }
      CUDA_OK(cudaMemcpyAsync(scal + 32, y.data(), sizeof(double) * it, cudaMemcpyHostToDevice, c->stream));
      CUDA_OK(cudaStreamSynchronize(c->stream));
      XSB_CHK(vec_maxpy_dev(c, n, x, Z.data(), it, scal + 32, 1.0));
    }
  }
  if (its_out) *its_out = its;
  return 0;
}

// -fs_coarse (exSaddle.c:362-400): the coarse level of the monolithic -mg hierarchy solved by FGMRES preconditioned with
// PCFIELDSPLIT Schur / UPPER / user Mpscaled_coarse, both splits GMRES + Jacobi, the Schur complement applied with a nested
// velocity solve (golden exSaddle3d_mg_fs_coarse_1, Makefile:390; solver tree as its -saddle_ksp_view prints it).
struct FsCoarse {
  xsb_ctx P = nullptr; double *id00 = nullptr, *idmp = nullptr;
  std::vector<double *> Vu, Vp, Vo, Zo; double *u_t1 = nullptr, *u_t2 = nullptr, *u_rhs = nullptr, *u_sol = nullptr, *p_t1 = nullptr, *p_t2 = nullptr, *p_tmp = nullptr, *o_t1 = nullptr, *pc_z = nullptr;
  double rtol = 1e-5, u_rtol = 1e-5, p_rtol = 1e-5; int max_it = 10000; std::vector<int> its;
};
}   // namespace

void fsc_free(void *h) { delete (FsCoarse *)h; }
int fsc_setup(xsb_ctx c, xsb_ctx P, void **out)
{
  Options &o = c->opt; const Lattice &L = P->lat;
  const char *need[][2] = {{"saddle_mg_coarse_ksp_type", "fgmres"}, {"saddle_mg_coarse_fieldsplit_u_pc_type", "jacobi"}, {"saddle_mg_coarse_fieldsplit_p_pc_type", "jacobi"}, {"saddle_mg_coarse_ksp_convergence_test", "default"}};
  for (auto &kv : need) if (o.str(kv[0], "") != kv[1]) return xsb_fail(c, XSB_ERR_SUP, "-fs_coarse is supported with the reference's coarse tree: -%s %s (Makefile:390)", kv[0], kv[1]);
  if (o.str("saddle_mg_coarse_fieldsplit_u_ksp_type", "gmres") != "gmres" || o.str("saddle_mg_coarse_fieldsplit_p_ksp_type", "gmres") != "gmres") return xsb_fail(c, XSB_ERR_SUP, "-fs_coarse: the coarse splits use gmres");
  FsCoarse *F = new FsCoarse(); *out = F; F->P = P;
  F->rtol = o.real("saddle_mg_coarse_ksp_rtol", 1e-5); F->max_it = o.integer("saddle_mg_coarse_ksp_max_it", 10000);
  F->u_rtol = o.real("saddle_mg_coarse_fieldsplit_u_ksp_rtol", 1e-5); F->p_rtol = o.real("saddle_mg_coarse_fieldsplit_p_ksp_rtol", 1e-5);
  XSB_CHK(dev_alloc(c, &F->id00, (size_t)L.nu)); XSB_CHK(baij_diag_inv(c, P->A00, F->id00));
  XSB_CHK(dev_alloc(c, &F->idmp, (size_t)L.np)); XSB_CHK(csr_diag_inv(c, P->Mp, F->idmp));
  for (double **v : {&F->u_t1, &F->u_t2, &F->u_rhs, &F->u_sol}) XSB_CHK(dev_alloc(c, v, (size_t)L.nu));
  for (double **v : {&F->p_t1, &F->p_t2, &F->p_tmp}) XSB_CHK(dev_alloc(c, v, (size_t)L.np));
  XSB_CHK(dev_alloc(c, &F->o_t1, (size_t)L.n)); XSB_CHK(dev_alloc(c, &F->pc_z, (size_t)L.n));
  return 0;
}
int fsc_solve(xsb_ctx c, void *h, const double *b, double *x)
{
  FsCoarse *F = (FsCoarse *)h; xsb_ctx P = F->P; const Lattice &L = P->lat; const int64_t nu = L.nu, np = L.np;
  Epilogue plain;
  Op A00 = [&](const double *v, double *y) { return spmv_baij(c, P->A00, v, y, plain); };
  Op Mu = [&](const double *v, double *y) { return vec_pmult(c, nu, F->id00, v, y); };
  Op Mp = [&](const double *v, double *y) { return vec_pmult(c, np, F->idmp, v, y); };
  auto ksp_u = [&](const double *rhs, double *sol) { return gmres_left(c, nu, A00, Mu, rhs, sol, F->u_rtol, 10000, 30, F->Vu, F->u_t1, F->u_t2, c->scal, nullptr); };
  Op S = [&](const double *v, double *y) {
    XSB_CHK(spmv_csr(c, P->A01, v, F->u_rhs));
    XSB_CHK(ksp_u(F->u_rhs, F->u_sol));
    XSB_CHK(spmv_csr(c, P->A10, F->u_sol, F->p_tmp));
    XSB_CHK(spmv_csr(c, P->A11, v, y));
    return vec_axpy(c, np, -1.0, F->p_tmp, y); };
  Op PC = [&](const double *r, double *z) {
    double *yu = z, *yp = z + nu;
    XSB_CHK(gmres_left(c, np, S, Mp, r + nu, yp, F->p_rtol, 10000, 30, F->Vp, F->p_t1, F->p_t2, c->scal + 96, nullptr));
    XSB_CHK(spmv_csr(c, P->A01, yp, F->u_rhs));
    XSB_CHK(vec_aypx(c, nu, -1.0, r, F->u_rhs));
    return ksp_u(F->u_rhs, yu); };
  Op A = [&](const double *v, double *y) { return spmv_csr(c, P->A, v, y); };
  int its = 0;
  XSB_CHK(fgmres_right(c, L.n, A, PC, b, x, F->rtol, F->max_it, 30, F->Vo, F->Zo, F->o_t1, c->scal + 192, &its));
  F->its.push_back(its);
  return 0;
}
int fsc_last_its(void *h) { FsCoarse *F = (FsCoarse *)h; return F->its.empty() ? 0 : F->its.back(); }

// bjacobi on R ranks (goldens *_fs_2): entries that couple dofs of different ranks are zeroed in the copies the ILU(0)s factor.
// ILU(0) of such a matrix in natural ordering IS the block ILU(0) (cross-block entries stay exactly zero through the elimination,
// in-block operations and their order are those of the rank-local factorisation: ownership is a box, so the rank-local DMDA
// ordering is the natural ordering restricted to the box).
__global__ void k_zero_cross_block(int n, const int *__restrict__ ia, const int *__restrict__ ja, double *__restrict__ a, const int *__restrict__ owner)
{
  const int row = blockIdx.x * blockDim.x + threadIdx.x; if (row >= n) return;
  const int o = owner[row];
  for (int k = ia[row]; k < ia[row + 1]; ++k) if (owner[ja[k]] != o) a[k] = 0.0;
}
// owner rank of every velocity dof / pressure node for the communicator size R (xsb_asm_subdomain: PETSc's DMDA ownership)
static int rank_owners(xsb_ctx c, int R, int **own_u, int **own_p)
{
  const Lattice &L = c->lat; const int nsd = L.nsd;
  std::vector<int> hu(L.nu, -1), hp(L.np, -1);
  for (int r = 0; r < R; ++r) {
    int b[18];
    if (xsb_asm_subdomain(nsd, L.mx, L.my, L.mz, R, 0, r, b)) return xsb_fail(c, XSB_ERR_ARG, "-xsb_ranks %d: PETSc's DMDA cannot partition this mesh into whole Q2 elements", R);
    for (int k = nsd == 3 ? b[8] : 0; k < (nsd == 3 ? b[11] : 1); ++k) for (int j = b[7]; j < b[10]; ++j) for (int i = b[6]; i < b[9]; ++i) {
      const int64_t nd = i + (int64_t)j * L.NX + (int64_t)k * L.NX * L.NY;
      for (int d = 0; d < nsd; ++d) hu[nd * nsd + d] = r;
    }
    for (int k = nsd == 3 ? b[14] : 0; k < (nsd == 3 ? b[17] : 1); ++k) for (int j = b[13]; j < b[16]; ++j) for (int i = b[12]; i < b[15]; ++i)
      hp[i + (int64_t)j * L.PX + (int64_t)k * L.PX * L.PY] = r;
  }
  for (int v : hu) if (v < 0) return xsb_fail(c, XSB_ERR_ARG, "rank ownership does not cover the velocity lattice");
  for (int v : hp) if (v < 0) return xsb_fail(c, XSB_ERR_ARG, "rank ownership does not cover the pressure lattice");
  XSB_CHK(dev_alloc(c, own_u, (size_t)L.nu)); XSB_CHK(dev_alloc(c, own_p, (size_t)L.np));
  // on the handle's (non-blocking) stream, and waited for: a blocking cudaMemcpy from pageable memory may return before its DMA
  // has landed, and nothing orders the legacy stream against the kernels that read these arrays
  CUDA_OK(cudaMemcpyAsync(*own_u, hu.data(), sizeof(int) * L.nu, cudaMemcpyHostToDevice, c->stream));
  CUDA_OK(cudaMemcpyAsync(*own_p, hp.data(), sizeof(int) * L.np, cudaMemcpyHostToDevice, c->stream));
  CUDA_OK(cudaStreamSynchronize(c->stream));
  return 0;
}

void fsd_free(xsb_ctx c) { if (c->fsd) { delete (Fsd *)c->fsd; c->fsd = nullptr; } }

int fsd_setup(xsb_ctx c)
{
  if (c->slab.nranks > 1) return xsb_fail(c, XSB_ERR_SUP, "the default -fs tree (GMRES + ILU(0) on A00) is implemented for one GPU; use the abf.opts tree on slabs");
  if (c->no_A) return xsb_fail(c, XSB_ERR_SUP, "the default -fs tree factors the assembled A00 (not -xsb_matrix_free full)");
  fsd_free(c);
  Fsd *F = new Fsd(); c->fsd = F;
  Options &o = c->opt; const Lattice &L = c->lat; cudaStream_t st = c->stream;
  F->u_max_it = o.integer("saddle_fieldsplit_u_ksp_max_it", 10000); F->u_rtol = o.real("saddle_fieldsplit_u_ksp_rtol", 1e-5);
  F->p_max_it = o.integer("saddle_fieldsplit_p_ksp_max_it", 10000); F->p_rtol = o.real("saddle_fieldsplit_p_ksp_rtol", 1e-5);
  const std::string pk = o.str("saddle_fieldsplit_p_ksp_type", "gmres");
  if (pk == "preonly") F->p_preonly = 1; else if (pk != "gmres") return xsb_fail(c, XSB_ERR_SUP, "-saddle_fieldsplit_p_ksp_type %s (gmres|preonly)", pk.c_str());
  const std::string uk = o.str("saddle_fieldsplit_u_ksp_type", "gmres"), up = o.str("saddle_fieldsplit_u_pc_type", "ilu"), pp = o.str("saddle_fieldsplit_p_pc_type", "ilu");
  if (uk != "gmres" || up != "ilu" || (pp != "ilu" && pp != "bjacobi")) return xsb_fail(c, XSB_ERR_SUP, "default -fs tree: fieldsplit_u gmres+ilu, fieldsplit_p gmres|preonly + ilu");
  // scalar CSR of A00
  const Baij &B = c->A00; const int bs = B.bs; Csr &S = F->A00s;
  S.n = S.m = B.nb * bs; S.nnz = B.nblk * bs * bs;
  XSB_CHK(dev_alloc(c, &S.ia, (size_t)S.n + 1)); XSB_CHK(dev_alloc(c, &S.ja, (size_t)S.nnz)); XSB_CHK(dev_alloc(c, &S.a, (size_t)S.nnz));
  if (bs == 3) k_baij_scalar<3><<<(B.nb + 256) / 256, 256, 0, st>>>(B.nb, B.ia, B.ja, B.a, S.ia, S.ja, S.a);
  else k_baij_scalar<2><<<(B.nb + 256) / 256, 256, 0, st>>>(B.nb, B.ia, B.ja, B.a, S.ia, S.ja, S.a);
  KERNEL_OK();
  c->mp_block_a = nullptr;
  const int R = o.integer("xsb_ranks", 1);
  if (R > 1) {   // the reference on R ranks: PETSc's default inner PC is bjacobi, one ILU(0) block per rank
    int *own_u = nullptr, *own_p = nullptr;
    XSB_CHK(rank_owners(c, R, &own_u, &own_p));
    k_zero_cross_block<<<(S.n + 255) / 256, 256, 0, st>>>(S.n, S.ia, S.ja, S.a, own_u); KERNEL_OK();
    XSB_CHK(dev_alloc(c, &c->mp_block_a, (size_t)c->Mp.nnz));
    CUDA_OK(cudaMemcpyAsync(c->mp_block_a, c->Mp.a, sizeof(double) * c->Mp.nnz, cudaMemcpyDeviceToDevice, st));
    k_zero_cross_block<<<(c->Mp.n + 255) / 256, 256, 0, st>>>(c->Mp.n, c->Mp.ia, c->Mp.ja, c->mp_block_a, own_p); KERNEL_OK();
  }
  XSB_CHK(gilu_setup(c, S, F->ilu_u));
  XSB_CHK(ilu_setup(c));   // ILU(0) of Mpscaled (xsb_ilu.cu); of its rank-block-diagonal part when -xsb_ranks > 1
  XSB_CHK(dev_alloc(c, &F->u_t1, (size_t)L.nu)); XSB_CHK(dev_alloc(c, &F->u_t2, (size_t)L.nu)); XSB_CHK(dev_alloc(c, &F->u_rhs, (size_t)L.nu)); XSB_CHK(dev_alloc(c, &F->u_sol, (size_t)L.nu));
  XSB_CHK(dev_alloc(c, &F->p_t1, (size_t)L.np)); XSB_CHK(dev_alloc(c, &F->p_t2, (size_t)L.np)); XSB_CHK(dev_alloc(c, &F->p_tmp, (size_t)L.np));
  return 0;
}

// PCApply_FieldSplit_Schur, UPPER, with the default sub-solvers
int fsd_apply(xsb_ctx c, const double *r, double *z)
{
  Fsd *F = (Fsd *)c->fsd; const Lattice &L = c->lat; const int64_t nu = L.nu, np = L.np;
  Epilogue plain;
  Op A00 = [&](const double *x, double *y) { return spmv_baij(c, c->A00, x, y, plain); };
  Op Mu = [&](const double *x, double *y) { return gilu_apply(c, F->ilu_u, x, y); };
  Op Mp = [&](const double *x, double *y) { return ilu_apply(c, x, y); };
  auto ksp_u = [&](const double *rhs, double *sol) { F->n_ksp_u++; return gmres_left(c, nu, A00, Mu, rhs, sol, F->u_rtol, F->u_max_it, 30, F->Vu, F->u_t1, F->u_t2, c->scal, nullptr); };
  double *yu = z, *yp = z + nu;
  if (F->p_preonly) XSB_CHK(Mp(r + nu, yp));
  else {
    // S v = A11 v - A10 KSP_u(A01 v): MatSchurComplement with a nested velocity solve per application
    Op S = [&](const double *v, double *y) {
      XSB_CHK(spmv_csr(c, c->A01, v, F->u_rhs));
      XSB_CHK(ksp_u(F->u_rhs, F->u_sol));
      XSB_CHK(spmv_csr(c, c->A10, F->u_sol, F->p_tmp));
      XSB_CHK(spmv_csr(c, c->A11, v, y));
      return vec_axpy(c, np, -1.0, F->p_tmp, y);
    };
    XSB_CHK(gmres_left(c, np, S, Mp, r + nu, yp, F->p_rtol, F->p_max_it, 30, F->Vp, F->p_t1, F->p_t2, c->scal + 96, nullptr));
  }
  XSB_CHK(spmv_csr(c, c->A01, yp, F->u_rhs));
  XSB_CHK(vec_aypx(c, nu, -1.0, r, F->u_rhs));      // x_u - A01 y_p
  return ksp_u(F->u_rhs, yu);
}
