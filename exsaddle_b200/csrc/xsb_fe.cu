// xsb_fe.cu -- device-side FE set-up of the Q2-Q1 saddle operator (kernels K5/K6 of SURVEY 2.1).
//
// Replaces, for one rank, the reference's L3 layer (femixedspace.c) and the model functions of models.c:
//   coefficient evaluation      FEMixedSpaceDefineQPwiseProperties   femixedspace.c:1857-1933, models.c:855,1482
//   Q1 projection               ..._Q1Projection (fine level)         femixedspace.c:1976-2083
//   sparsity pattern            MatAssemble_Saddle_NULL               femixedspace.c:2306-2370 (closed form, App. A.5)
//   element assembly            MatAssemble_Saddle / _Schur, VecAssemble_F1/F2   :2373-2786, :2837-2948
//   Dirichlet handling          rhs_diri, MatZeroRowsColumns(1.0)     femixedspace.c:2634-2645, exSaddle.c:276-281
//   sub-blocks                  MatCreateSubMatrix on the u/p ISs     exSaddle.c:319-321
// The mesh is the reference's uniform DMDA box, so J = diag(h) and the global basis gradients are tabulated
// once; assembly is atomic-free: elements are processed in 8 (4 in 2-D) parity colours, one CTA per element.
#include "xsb.h"
#include <cub/cub.cuh>
#include <cstdarg>

// ------------------------------------------------------------------ tables
static void host_tables(const Lattice &L, FeTables &T)
{
  static const double xi1d[3] = {-0.774596669241483, 0.0, 0.774596669241483};
  static const double wt1d[3] = {0.555555555555556, 0.888888888888889, 0.555555555555556};
  const int nsd = L.nsd, n3 = nsd == 3 ? 3 : 1;
  memset(&T, 0, sizeof(T));
  int q = 0;
  for (int kq = 0; kq < n3; ++kq) for (int jq = 0; jq < 3; ++jq) for (int iq = 0; iq < 3; ++iq, ++q) {
    double xi[3] = {xi1d[iq], xi1d[jq], nsd == 3 ? xi1d[kq] : 0.0};
    T.wq[q] = nsd == 3 ? wt1d[iq] * wt1d[jq] * wt1d[kq] : wt1d[iq] * wt1d[jq];
    double b[3][3], g[3][3];
    for (int d = 0; d < 3; ++d) {   // femixedspace.c:1540-1542, 1837-1839
      double x = xi[d];
      b[d][0] = 0.5 * x * (x - 1.0); b[d][1] = (1.0 + x) * (1.0 - x); b[d][2] = 0.5 * (1.0 + x) * x;
      g[d][0] = 0.5 * (2.0 * x - 1.0); g[d][1] = -2.0 * x; g[d][2] = 0.5 * (2.0 * x + 1.0);
    }
    int c = 0;
    for (int k = 0; k < n3; ++k) for (int j = 0; j < 3; ++j) for (int i = 0; i < 3; ++i, ++c) {
      double bk = nsd == 3 ? b[2][k] : 1.0;
      T.Nu[q][c] = b[0][i] * b[1][j] * bk;
      T.Gu[q][c][0] = g[0][i] * b[1][j] * bk / L.hu[0];
      T.Gu[q][c][1] = b[0][i] * g[1][j] * bk / L.hu[1];
      T.Gu[q][c][2] = nsd == 3 ? b[0][i] * b[1][j] * g[2][k] / L.hu[2] : 0.0;
    }
    c = 0;
    for (int k = 0; k < (nsd == 3 ? 2 : 1); ++k) for (int j = 0; j < 2; ++j) for (int i = 0; i < 2; ++i, ++c) {   // :1497-1509
      double sx = i ? 1.0 + xi[0] : 1.0 - xi[0], sy = j ? 1.0 + xi[1] : 1.0 - xi[1], sz = k ? 1.0 + xi[2] : 1.0 - xi[2];
      T.Np[q][c] = nsd == 3 ? 0.125 * sx * sy * sz : 0.25 * sx * sy;
    }
  }
  T.detJ = nsd == 3 ? L.hu[0] * L.hu[1] * L.hu[2] : L.hu[0] * L.hu[1];
}

// ------------------------------------------------------------------ model resolution (host)
static void bprintf(std::string &s, const char *fmt, ...)
{
  char buf[512]; va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap); s += buf;
}

// Option blocks of models.c (defaults, banner text) + BC dispatch models.c:610-648.
int fe_resolve_model(xsb_ctx c)
{
  Options &o = c->opt; Model &m = c->mdl; Lattice &L = c->lat;
  const int nsd = c->nsd;
  L.nsd = nsd;
  L.mx = o.integer("mx", 4); L.my = o.integer("my", L.mx); L.mz = o.integer("mz", L.mx);   // exSaddle.c:178-182
  if (nsd == 2) L.mz = 1;
  if (L.mx < 1 || L.my < 1 || L.mz < 1) return xsb_fail(c, XSB_ERR_ARG, "-mx/-my/-mz must be >= 1");
  double size[3] = {o.real("size_x", 1.0), o.real("size_y", 1.0), o.real("size_z", 1.0)};   // exSaddle.c:183-185
  L.NX = 2 * L.mx + 1; L.NY = 2 * L.my + 1; L.NZ = nsd == 3 ? 2 * L.mz + 1 : 1;              // femixedspace.c:1157-1158
  L.PX = L.mx + 1; L.PY = L.my + 1; L.PZ = nsd == 3 ? L.mz + 1 : 1;                          // femixedspace.c:1248-1249
  L.nun = (int64_t)L.NX * L.NY * L.NZ; L.npn = (int64_t)L.PX * L.PY * L.PZ;
  L.nu = nsd * L.nun; L.np = L.npn; L.n = L.nu + L.np; L.nel = (int64_t)L.mx * L.my * L.mz;
  L.hu[0] = size[0] / (L.NX - 1); L.hu[1] = size[1] / (L.NY - 1); L.hu[2] = nsd == 3 ? size[2] / (L.NZ - 1) : 1.0;
  {   // z-slab partition: shrink the lattice to this rank's local element layers (coordinates stay global)
    Slab &S = c->slab; const int mzg = L.mz;
    S.mz_glob = mzg; S.k0 = 0; S.k1 = mzg; S.e0 = 0; S.e1 = mzg;
    if (S.nranks > 1) {
      if (xsb_slab_range(mzg, S.nranks, S.rank, &S.k0, &S.k1)) return xsb_fail(c, XSB_ERR_ARG, "-mz %d cannot be cut into %d slabs", mzg, S.nranks);
      S.e0 = S.k0 - 2 < 0 ? 0 : S.k0 - 2; S.e1 = S.k1 + 1 > mzg ? mzg : S.k1 + 1;
    }
    const bool last = S.rank == S.nranks - 1;
    L.mz = S.e1 - S.e0; L.zoff = S.e0;
    L.NZ = nsd == 3 ? 2 * L.mz + 1 : 1; L.PZ = nsd == 3 ? L.mz + 1 : 1;
    L.nun = (int64_t)L.NX * L.NY * L.NZ; L.npn = (int64_t)L.PX * L.PY * L.PZ;
    L.nu = nsd * L.nun; L.np = L.npn; L.n = L.nu + L.np; L.nel = (int64_t)L.mx * L.my * L.mz;
    S.ou0 = nsd == 3 ? 2 * (S.k0 - S.e0) : 0; S.ou1 = nsd == 3 ? 2 * (S.k1 - S.e0) + (last ? 1 : 0) : 1;
    S.op0 = nsd == 3 ? S.k0 - S.e0 : 0;       S.op1 = nsd == 3 ? S.k1 - S.e0 + (last ? 1 : 0) : 1;
    const int64_t pu = (int64_t)nsd * L.NX * L.NY, pp = (int64_t)L.PX * L.PY;
    c->own_u = Ranges(); c->own_u.off0 = S.ou0 * pu; c->own_u.len0 = (S.ou1 - S.ou0) * pu;
    c->own_p = Ranges(); c->own_p.off0 = S.op0 * pp; c->own_p.len0 = (S.op1 - S.op0) * pp;
    c->own_full = c->own_u; c->own_full.off1 = L.nu + c->own_p.off0; c->own_full.len1 = c->own_p.len0;
  }
  if (L.n >= INT32_MAX) return xsb_fail(c, XSB_ERR_SUP, "more than 2^31 unknowns: the assembled AIJ path uses 32-bit PetscInt indices");
  m.model = o.integer("model", c->lame ? 6 : 2);   // models.h:9-13
  m.freeslip = o.flag("freesliphack");
  m.size_x = size[0];
  m.bc_type = BC_SOLCX;
  if (c->lame && m.model == 8) m.bc_type = BC_FIXEDBASE;
  if (c->lame && (m.model == 9 || m.model == 10)) m.bc_type = BC_COMPRESSION;
  if (nsd == 3 && m.model == 11) m.bc_type = BC_FIXEDBASE;
  if (c->lame && nsd == 3 && m.model == 12) m.bc_type = BC_COMPRESSION2;
  if (!c->lame && nsd == 2 && m.model == 101) m.bc_type = BC_MMS1;
  static const char *bcn[] = {"SolCx", "FixedBase", "Compression", "Compression2", "StokesMMS1"};
  std::string &b = c->banner; b.clear();
  bprintf(b, "Boundary Conditions: %s\n", bcn[m.bc_type]);
  if (c->lame) {
    m.c0 = o.real("mu0", 1.0); m.lam0 = o.real("lambda0", 1.0);
    switch (m.model) {
    case 2:
      m.c1 = o.real("mu1", 1.0); m.lam1 = o.real("lambda1", 1.0); m.rad = o.real("sinker_r", 0.05); m.nsink = o.integer("sinker_n", 3);
      bprintf(b, "ModelType: LameXSinker\n  params: mu0 %1.4e\n  params: mu1 %1.4e\n  params: lambda0 %1.4e\n  params: lambda1 %1.4e\n  params: num sinkers %d\n  params: sinker radius %1.4e\n", m.c0, m.c1, m.lam0, m.lam1, m.nsink, m.rad);
      if (m.nsink > 8) return xsb_fail(c, XSB_ERR_SUP, "Too many sinkers");
      if (m.rad > 0.05) return xsb_fail(c, XSB_ERR_SUP, "Sinker Radius too big");
      break;
    case 6: case 8: case 10: case 12:
      m.c1 = o.real("mu1", 1.0); m.lam1 = o.real("lambda1", 2.0); m.rad = o.real("sinker_r", 0.25);
      bprintf(b, "ModelType: LameOneSinker\n  params: mu0 %1.4e\n  params: mu1 %1.4e\n  params: lambda0 %1.4e\n  params: lambda1 %1.4e\n  params: rad %1.4e\n", m.c0, m.c1, m.lam0, m.lam1, m.rad);
      break;
    case 9:
      bprintf(b, "ModelType: LameHomogeneous\n  params: mu0 %1.4e\n  params: lambda0 %1.4e\n", m.c0, m.lam0);
      break;
    default: return xsb_fail(c, XSB_ERR_SUP, "Elasticity Model %d not implemented", m.model);
    }
  } else {
    m.c0 = o.real("eta0", 1.0);
    switch (m.model) {
    case 0: case 5:
      if (m.model == 5 && nsd != 3) return xsb_fail(c, XSB_ERR_SUP, "Stokes Model 5 not implemented in 2d");
      m.c1 = o.real("eta1", 1.0); m.xc = o.real("solcx_xc", 0.5); m.nz = o.integer("solcx_nz", 1);
      bprintf(b, "ModelType: %s\n  params: eta0 %1.4e\n  params: eta1 %1.4e\n  params: xc   %1.4e\n  params: nz   %d\n", m.model == 0 ? "StokesSolCx" : "StokesSolCx3d", m.c0, m.c1, m.xc, m.nz);
      if (m.model == 5) bprintf(b, "  params: nz2  %d\n", 1);
      break;
    case 1:
      m.c1 = o.real("eta1", 1.0); m.rad = o.real("sinker_r", 0.1);
      bprintf(b, "ModelType: StokesThreeSinker\n  params: eta0 %1.4e\n  params: eta1 %1.4e\n  params: rad  %1.4e\n", m.c0, m.c1, m.rad);
      break;
    case 2:
      m.c1 = o.real("eta1", 1.0); m.rad = o.real("sinker_r", 0.05); m.nsink = o.integer("sinker_n", 3);
      bprintf(b, "ModelType: StokesXSinker\n  params: eta0 %1.4e\n  params: eta1 %1.4e\n  params: num sinkers %d\n  params: sinker radius %1.4e\n", m.c0, m.c1, m.nsink, m.rad);
      if (m.nsink > 8) return xsb_fail(c, XSB_ERR_SUP, "Too many sinkers");
      if (m.rad > 0.05) return xsb_fail(c, XSB_ERR_SUP, "Sinker Radius too big");
      break;
    case 6:
      m.c1 = o.real("eta1", 1.0); m.rad = o.real("sinker_r", 0.25);
      m.cx = o.real("sinker_x", 0.5); m.cy = o.real("sinker_y", 0.5); m.cz = o.real("sinker_z", 0.5);
      bprintf(b, "ModelType: StokesOneSinker\n  params: eta0 %1.4e\n  params: eta1 %1.4e\n  params: x %1.4e\n  params: y %1.4e\n", m.c0, m.c1, m.cx, m.cy);
      if (nsd == 3) bprintf(b, "  params: z %1.4e\n", m.cz);
      bprintf(b, "  params: rad %1.4e\n", m.rad);
      break;
    case 11:
      if (nsd != 3) return xsb_fail(c, XSB_ERR_SUP, "Stokes Model 11 not implemented in 2d");
      m.c1 = o.real("eta1", 10000.0);
      bprintf(b, "ModelType: PseudoIce\n  params: eta0 %1.4e\n  params: eta1 %1.4e\n", m.c0, m.c1);
      break;
    case 101:
      if (nsd != 2) return xsb_fail(c, XSB_ERR_SUP, "Stokes Model 101 not implemented in 3d");
      bprintf(b, "ModelType: StokesMMS1\n");
      break;
    default: return xsb_fail(c, XSB_ERR_SUP, "Stokes Model %d not implemented", m.model);
    }
  }
  return XSB_OK;
}

// ------------------------------------------------------------------ coefficient kernels
__constant__ double c_posx[8] = {0.27, 0.6, 0.7, 0.2, 0.85, 0.4, 0.16, 0.55};   // models.c:1012-1015
__constant__ double c_posy[8] = {0.63, 0.83, 0.33, 0.2, 0.65, 0.3, 0.84, 0.54};
__constant__ double c_posz[8] = {0.50, 0.40, 0.30, 0.70, 0.65, 0.4, 0.8, 0.50};

__device__ void eval_model(const Model &m, int lame, int nsd, const double *x, double *out)
{
  const bool d3 = nsd == 3;
  double cc = m.c0, lam = m.lam0, rho = 1.0; bool inside = false;
  for (int s = 0; s < XSB_NSLOT; ++s) out[s] = 0.0;
  if (lame) {
    if (m.model == 2) {
      for (int i = 0; i < m.nsink; ++i) {
        double d2 = (x[0] - c_posx[i]) * (x[0] - c_posx[i]) + (x[1] - c_posy[i]) * (x[1] - c_posy[i]);
        if (d3) d2 += (x[2] - c_posz[i]) * (x[2] - c_posz[i]);
        if (d2 < m.rad * m.rad) { inside = true; break; }
      }
      if (inside) { cc = m.c1; lam = m.lam1; rho = 1.1; }
    } else if (m.model != 9) {
      double s2 = (x[0] - 0.5) * (x[0] - 0.5) + (x[1] - 0.5) * (x[1] - 0.5);
      if (d3) s2 += (x[2] - 0.5) * (x[2] - 0.5);
      if (s2 < m.rad * m.rad) { rho = 2.0; cc = m.c1; lam = m.lam1; }
    }
    out[C_ETA] = cc; out[C_LAM] = lam; out[C_FU1] = -rho;
    return;
  }
  const double pi = 3.14159265358979323846264338327950288419716939937510582;   // PETSC_PI
  switch (m.model) {
  case 0:
    if (x[0] > m.xc) cc = m.c1;
    out[C_ETA] = cc; out[C_FU1] = sin(m.nz * pi * x[1]) * cos(1.0 * pi * x[0]);
    return;
  case 5:
    if (x[0] > m.xc) cc = m.c1;
    out[C_ETA] = cc; out[C_FU1] = sin(m.nz * pi * x[1]) * cos(1.0 * pi * x[0]) * sin(1 * pi * x[2]);
    return;
  case 1: {
    const double sx[3] = {0.27, 0.6, 0.7}, sy[3] = {0.63, 0.83, 0.33};
    for (int i = 0; i < 3; ++i) {
      double s2 = (x[0] - sx[i]) * (x[0] - sx[i]) + (x[1] - sy[i]) * (x[1] - sy[i]);
      if (d3) s2 += (x[2] - 0.5) * (x[2] - 0.5);
      if (s2 < m.rad * m.rad) inside = true;
    }
    break; }
  case 2:
    for (int i = 0; i < m.nsink; ++i) {
      double d2 = (x[0] - c_posx[i]) * (x[0] - c_posx[i]) + (x[1] - c_posy[i]) * (x[1] - c_posy[i]);
      if (d3) d2 += (x[2] - c_posz[i]) * (x[2] - c_posz[i]);
      if (d2 < m.rad * m.rad) { inside = true; break; }
    }
    break;
  case 6: {
    double s2 = (x[0] - m.cx) * (x[0] - m.cx) + (x[1] - m.cy) * (x[1] - m.cy);
    if (d3) s2 += (x[2] - m.cz) * (x[2] - m.cz);
    if (s2 < m.rad * m.rad) inside = true;
    break; }
  case 11: {
    double xrel = x[0] / m.size_x;
    out[C_ETA] = xrel * m.c0 + (1 - xrel) * m.c1; out[C_FU2] = 1.0;
    return; }
  case 101:
    out[C_ETA] = 1.0;
    return;
  }
  if (inside) { cc = m.c1; rho = 1.1; }
  out[C_ETA] = cc; out[C_FU1] = -rho;
}

// one thread per (element, quadrature point): x_q = sum_i N_i(xi_q) x_i, then the model (femixedspace.c:1902-1927)
__global__ void coeff_eval_kernel(Lattice L, Model m, int lame, const FeTables *T, double *coeff)
{
  const int nqp = L.nsd == 3 ? 27 : 9, nbu = nqp;
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, tot = L.nel * nqp;
  if (t >= tot) return;
  int64_t e = t / nqp; int q = (int)(t - e * nqp);
  int ei = (int)(e % L.mx), ej = (int)((e / L.mx) % L.my), ek = (int)(e / ((int64_t)L.mx * L.my));
  double xq[3] = {0, 0, 0};
  for (int i = 0; i < nbu; ++i) {
    int ii = i % 3, jj = (i / 3) % 3, kk = i / 9;
    double N = T->Nu[q][i];
    xq[0] += N * (L.hu[0] * (2 * ei + ii)); xq[1] += N * (L.hu[1] * (2 * ej + jj));
    if (L.nsd == 3) xq[2] += N * (L.hu[2] * (2 * (ek + L.zoff) + kk));
  }
  double out[XSB_NSLOT];
  eval_model(m, lame, L.nsd, xq, out);
  for (int s = 0; s < XSB_NSLOT; ++s) coeff[(int64_t)s * tot + t] = out[s];
}

// one thread per pressure node: c_n = sum_e sum_q N_n c_eq / sum_e sum_q N_n, elements in ascending order (:1976-2018)
__global__ void q1_project_kernel(Lattice L, const FeTables *T, const double *coeff, double *nodal)
{
  const int nqp = L.nsd == 3 ? 27 : 9;
  int64_t nd = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (nd >= L.npn) return;
  int pi = (int)(nd % L.PX), pj = (int)((nd / L.PX) % L.PY), pk = (int)(nd / ((int64_t)L.PX * L.PY));
  double acc[XSB_NSLOT], scale = 0.0;
  for (int s = 0; s < XSB_NSLOT; ++s) acc[s] = 0.0;
  const int64_t tot = L.nel * nqp;
  for (int ek = pk - 1; ek <= pk; ++ek) {
    if (L.nsd == 3 ? (ek < 0 || ek >= L.mz) : (ek != pk)) continue;
    for (int ej = pj - 1; ej <= pj; ++ej) {
      if (ej < 0 || ej >= L.my) continue;
      for (int ei = pi - 1; ei <= pi; ++ei) {
        if (ei < 0 || ei >= L.mx) continue;
        int kk3 = L.nsd == 3 ? pk - ek : 0;
        int il = (pi - ei) + 2 * (pj - ej) + 4 * kk3;
        int64_t e = ei + (int64_t)ej * L.mx + (int64_t)(L.nsd == 3 ? ek : 0) * L.mx * L.my;
        double els = 0.0;
        for (int q = 0; q < nqp; ++q) els += T->Np[q][il];
        scale += els;
        for (int s = 0; s < XSB_NSLOT; ++s) {
          double elc = 0.0;
          for (int q = 0; q < nqp; ++q) elc += T->Np[q][il] * coeff[(int64_t)s * tot + e * nqp + q];
          acc[s] += elc;
        }
      }
    }
  }
  for (int s = 0; s < XSB_NSLOT; ++s) nodal[(int64_t)s * L.npn + nd] = acc[s] / scale;
}

// one thread per (element, qp): c_eq = sum_n N_n(xi_q) c_n  (:2036-2083)
__global__ void q1_interp_kernel(Lattice L, const FeTables *T, const double *nodal, double *coeff)
{
  const int nqp = L.nsd == 3 ? 27 : 9, nbp = L.nsd == 3 ? 8 : 4;
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, tot = L.nel * nqp;
  if (t >= tot) return;
  int64_t e = t / nqp; int q = (int)(t - e * nqp);
  int ei = (int)(e % L.mx), ej = (int)((e / L.mx) % L.my), ek = (int)(e / ((int64_t)L.mx * L.my));
  for (int s = 0; s < XSB_NSLOT; ++s) {
    double v = 0.0;
    for (int i = 0; i < nbp; ++i) {
      int64_t nd = (ei + (i & 1)) + (int64_t)(ej + ((i >> 1) & 1)) * L.PX + (int64_t)(ek + (i >> 2)) * L.PX * L.PY;
      v += T->Np[q][i] * nodal[(int64_t)s * L.npn + nd];
    }
    coeff[(int64_t)s * tot + t] = v;
  }
}

// ------------------------------------------------------------------ pattern
__global__ void row_len_kernel(Lattice L, int64_t *len)
{
  int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= L.n) return;
  RowBox b; int comp; row_to_box(L, row, b, &comp);
  len[row] = (int64_t)L.nsd * b.ncu + b.ncp;
}
__global__ void narrow_ia_kernel(int64_t n, const int64_t *ia64, int *ia) { int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; if (i <= n) ia[i] = (int)ia64[i]; }

// one warp per AIJ row: columns in ascending order
__global__ void fill_ja_kernel(Lattice L, const int *ia, int *ja)
{
  int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; int lane = threadIdx.x & 31;
  if (row >= L.n) return;
  RowBox b; int comp; row_to_box(L, row, b, &comp);
  const int nsd = L.nsd, nxu = b.uhi[0] - b.ulo[0] + 1, nyu = b.uhi[1] - b.ulo[1] + 1, nxp = b.phi[0] - b.plo[0] + 1, nyp = b.phi[1] - b.plo[1] + 1;
  const int nuc = nsd * b.ncu, tot = nuc + b.ncp; int *out = ja + ia[row];
  for (int t = lane; t < tot; t += 32) {
    if (t < nuc) {
      int s = t / nsd, d = t - s * nsd; int ii = s % nxu, jj = (s / nxu) % nyu, kk = s / (nxu * nyu);
      out[t] = nsd * ((b.ulo[0] + ii) + (b.ulo[1] + jj) * L.NX + (b.ulo[2] + kk) * L.NX * L.NY) + d;
    } else {
      int s = t - nuc; int ii = s % nxp, jj = (s / nxp) % nyp, kk = s / (nxp * nyp);
      out[t] = (int)L.nu + (b.plo[0] + ii) + (b.plo[1] + jj) * L.PX + (b.plo[2] + kk) * L.PX * L.PY;
    }
  }
}

// Mpscaled / 27-point scalar pattern on the pressure lattice (DMCreateMatrix(dmp), exSaddle.c:315)
__global__ void mp_len_kernel(Lattice L, int64_t *len)
{
  int64_t nd = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; if (nd >= L.npn) return;
  BoxPattern p{L.PX, L.PY, L.PZ, 0};
  len[nd] = box_size(p, (int)(nd % L.PX), (int)((nd / L.PX) % L.PY), (int)(nd / ((int64_t)L.PX * L.PY)));
}
__global__ void mp_ja_kernel(Lattice L, const int *ia, int *ja)
{
  int64_t nd = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; if (nd >= L.npn) return;
  int i = (int)(nd % L.PX), j = (int)((nd / L.PX) % L.PY), k = (int)(nd / ((int64_t)L.PX * L.PY));
  int l0, h0, l1, h1, l2, h2; range_pp(i, L.PX, l0, h0); range_pp(j, L.PY, l1, h1); range_pp(k, L.PZ, l2, h2);
  int c = ia[nd];
  for (int kk = l2; kk <= h2; ++kk) for (int jj = l1; jj <= h1; ++jj) for (int ii = l0; ii <= h0; ++ii) ja[c++] = ii + jj * L.PX + kk * L.PX * L.PY;
}

// ------------------------------------------------------------------ element assembly (K5)
// One CTA per element of the current parity colour. Elements of one colour share no node, so the
// read-modify-write of the global values needs no atomics; colours run in a fixed order, hence the
// result is bit-reproducible.  Per matrix entry the quadrature sum runs q = 0..nqp-1 with the
// surviving B^T D B terms in the reference's k order (femixedspace.c:2531-2559).
// SPLIT (operator-free mode, -xsb_matrix_free full): the velocity block is never formed; the gradient / divergence /
// pressure blocks go straight into their own CSR arrays (same values, same colour order).
struct SplitDst { const int *ia01, *ia10, *ia11; double *a01, *a10, *a11; };
template <int NSD, bool SPLIT>
__global__ void __launch_bounds__(256) assemble_kernel(Lattice L, int lame, int colour, const FeTables *T, const double *coeff,
                                                       const int *ia, double *a, const int *mia, double *ma, double *F, SplitDst sp)
{
  constexpr int NBU = NSD == 3 ? 27 : 9, NBP = NSD == 3 ? 8 : 4, NQP = NBU;
  __shared__ double sG[NQP][NBU][NSD];
  __shared__ double sNp[NQP][NBP];
  __shared__ double sNu[NQP][NBU];
  __shared__ double sfac[NQP], sw[NQP], sil[NQP], sinv[NQP], sF[NQP][4];
  const int ci = colour & 1, cj = (colour >> 1) & 1, ck = (colour >> 2) & 1;
  const int nei = (L.mx - ci + 1) / 2, nej = (L.my - cj + 1) / 2;
  const int64_t t = blockIdx.x;
  const int ei = 2 * (int)(t % nei) + ci, ej = 2 * (int)((t / nei) % nej) + cj, ek = NSD == 3 ? 2 * (int)(t / ((int64_t)nei * nej)) + ck : 0;
  const int64_t e = ei + (int64_t)ej * L.mx + (int64_t)ek * L.mx * L.my;
  const int64_t tot = L.nel * NQP;
  const int tid = threadIdx.x;
  for (int x = tid; x < NQP * NBU * NSD; x += blockDim.x) { int q = x / (NBU * NSD), r = x - q * NBU * NSD, i = r / NSD, d = r - i * NSD; sG[q][i][d] = T->Gu[q][i][d]; }
  for (int x = tid; x < NQP * NBP; x += blockDim.x) sNp[x / NBP][x % NBP] = T->Np[x / NBP][x % NBP];
  for (int x = tid; x < NQP * NBU; x += blockDim.x) sNu[x / NBU][x % NBU] = T->Nu[x / NBU][x % NBU];
  if (tid < NQP) {
    const double wd = T->wq[tid] * T->detJ;
    const double eta = coeff[(int64_t)C_ETA * tot + e * NQP + tid];
    sfac[tid] = eta * T->wq[tid] * T->detJ;   // fac = eta_c * w_qp * detJ (:2523-2526)
    sw[tid] = wd;                             // fac = w_qp * detJ        (:2573)
    if (lame) { const double lam = coeff[(int64_t)C_LAM * tot + e * NQP + tid]; sil[tid] = wd / lam; sinv[tid] = 1.0 / lam + 1.0 / eta; }
    else { sil[tid] = 0.0; sinv[tid] = 1.0 / eta; }
    sF[tid][0] = coeff[(int64_t)C_FU0 * tot + e * NQP + tid]; sF[tid][1] = coeff[(int64_t)C_FU1 * tot + e * NQP + tid];
    sF[tid][2] = coeff[(int64_t)C_FU2 * tot + e * NQP + tid]; sF[tid][3] = coeff[(int64_t)C_FP * tot + e * NQP + tid];
  }
  __syncthreads();
  const int nK = NSD == 3 ? 3 : 1;
  // ---- A11: one thread per (node i, node j) pair -> NSD x NSD block
  for (int pr = tid; !SPLIT && pr < NBU * NBU; pr += blockDim.x) {
    const int i = pr / NBU, j = pr - i * NBU;
    double r[NSD][NSD];
#pragma unroll
    for (int x = 0; x < NSD; ++x)
#pragma unroll
      for (int y = 0; y < NSD; ++y) r[x][y] = 0.0;
    for (int q = 0; q < NQP; ++q) {
      const double D1 = 1.0 * sfac[q], D2 = 2.0 * sfac[q];
      const double xi = sG[q][i][0], yi = sG[q][i][1], xj = sG[q][j][0], yj = sG[q][j][1];
      if (NSD == 2) {
        r[0][0] += xi * D2 * xj; r[0][0] += yi * D1 * yj;
        r[0][1] += yi * D1 * xj;
        r[1][0] += xi * D1 * yj;
        r[1][1] += yi * D2 * yj; r[1][1] += xi * D1 * xj;
      } else {
        const double zi = sG[q][i][NSD - 1], zj = sG[q][j][NSD - 1];
        r[0][0] += xi * D2 * xj; r[0][0] += yi * D1 * yj; r[0][0] += zi * D1 * zj;
        r[0][1] += yi * D1 * xj;
        r[0][NSD - 1] += zi * D1 * xj;
        r[1][0] += xi * D1 * yj;
        r[1][1] += yi * D2 * yj; r[1][1] += xi * D1 * xj; r[1][1] += zi * D1 * zj;
        r[1][NSD - 1] += zi * D1 * yj;
        r[NSD - 1][0] += xi * D1 * zj;
        r[NSD - 1][1] += yi * D1 * zj;
        r[NSD - 1][NSD - 1] += zi * D2 * zj; r[NSD - 1][NSD - 1] += xi * D1 * xj; r[NSD - 1][NSD - 1] += yi * D1 * yj;
      }
    }
    const int gi = 2 * ei + i % 3, gj = 2 * ej + (i / 3) % 3, gk = NSD == 3 ? 2 * ek + i / 9 : 0;
    const int hi = 2 * ei + j % 3, hj = 2 * ej + (j / 3) % 3, hk = NSD == 3 ? 2 * ek + j / 9 : 0;
    RowBox b; row_box_u(L, gi, gj, gk, b);
    const int64_t node = gi + (int64_t)gj * L.NX + (int64_t)gk * L.NX * L.NY;
    const int up = NSD * box_upos(b, hi, hj, hk);
#pragma unroll
    for (int x = 0; x < NSD; ++x) {
      double *row = a + ia[NSD * node + x] + up;
#pragma unroll
      for (int y = 0; y < NSD; ++y) row[y] += r[x][y];
    }
  }
  // ---- A12 / A21: one thread per (velocity node i, pressure node j)
  for (int pr = tid; pr < NBU * NBP; pr += blockDim.x) {
    const int i = pr / NBP, j = pr - i * NBP;
    double g[NSD];
#pragma unroll
    for (int d = 0; d < NSD; ++d) g[d] = 0.0;
    for (int q = 0; q < NQP; ++q)
#pragma unroll
      for (int d = 0; d < NSD; ++d) g[d] -= sG[q][i][d] * sNp[q][j] * sw[q];   // :2576-2579
    const int gi = 2 * ei + i % 3, gj = 2 * ej + (i / 3) % 3, gk = NSD == 3 ? 2 * ek + i / 9 : 0;
    const int pi = ei + (j & 1), pj = ej + ((j >> 1) & 1), pk = NSD == 3 ? ek + (j >> 2) : 0;
    RowBox bu; row_box_u(L, gi, gj, gk, bu);
    RowBox bp; row_box_p(L, pi, pj, pk, bp);
    const int64_t node = gi + (int64_t)gj * L.NX + (int64_t)gk * L.NX * L.NY;
    const int64_t pnode = pi + (int64_t)pj * L.PX + (int64_t)pk * L.PX * L.PY;
    const int pp = NSD * bu.ncu + box_ppos(bu, pi, pj, pk);
    const int upos = NSD * box_upos(bp, gi, gj, gk);
    if (SPLIT) {
      double *prow = sp.a10 + sp.ia10[pnode] + upos; const int ppos = box_ppos(bu, pi, pj, pk);
#pragma unroll
      for (int d = 0; d < NSD; ++d) { sp.a01[sp.ia01[NSD * node + d] + ppos] += g[d]; prow[d] += g[d]; }
    } else {
      double *prow = a + ia[L.nu + pnode] + upos;
#pragma unroll
      for (int d = 0; d < NSD; ++d) { a[ia[NSD * node + d] + pp] += g[d]; prow[d] += g[d]; }   // A21 = A12^T (:2584-2590)
    }
  }
  // ---- A22 (LAME) and Mpscaled: one thread per (pressure node i, pressure node j)
  for (int pr = tid; pr < NBP * NBP; pr += blockDim.x) {
    const int i = pr / NBP, j = pr - i * NBP;
    double a22 = 0.0, s = 0.0;
    for (int q = 0; q < NQP; ++q) {
      a22 -= sNp[q][i] * sNp[q][j] * sil[q];               // :2602-2606
      s -= sinv[q] * sNp[q][i] * sNp[q][j] * sw[q];        // :2924-2928
    }
    const int pi = ei + (i & 1), pj = ej + ((i >> 1) & 1), pk = NSD == 3 ? ek + (i >> 2) : 0;
    const int qi = ei + (j & 1), qj = ej + ((j >> 1) & 1), qk = NSD == 3 ? ek + (j >> 2) : 0;
    RowBox bp; row_box_p(L, pi, pj, pk, bp);
    const int64_t pnode = pi + (int64_t)pj * L.PX + (int64_t)pk * L.PX * L.PY;
    const int pos = box_ppos(bp, qi, qj, qk);
    if (lame) { if (SPLIT) sp.a11[sp.ia11[pnode] + pos] += a22; else a[ia[L.nu + pnode] + NSD * bp.ncu + pos] += a22; }
    ma[mia[pnode] + pos] += s;
  }
  // ---- F1 / F2 (:2695-2710, :2763-2778)
  for (int i = tid; i < NBU + NBP; i += blockDim.x) {
    if (i < NBU) {
      double f[NSD];
#pragma unroll
      for (int d = 0; d < NSD; ++d) f[d] = 0.0;
      for (int q = 0; q < NQP; ++q)
#pragma unroll
        for (int d = 0; d < NSD; ++d) f[d] += sNu[q][i] * sF[q][d] * sw[q];
      const int gi = 2 * ei + i % 3, gj = 2 * ej + (i / 3) % 3, gk = NSD == 3 ? 2 * ek + i / 9 : 0;
      const int64_t node = gi + (int64_t)gj * L.NX + (int64_t)gk * L.NX * L.NY;
#pragma unroll
      for (int d = 0; d < NSD; ++d) F[NSD * node + d] += f[d];
    } else {
      const int j = i - NBU; double f = 0.0;
      for (int q = 0; q < NQP; ++q) f += sNp[q][j] * sF[q][3] * sw[q];
      const int pi = ei + (j & 1), pj = ej + ((j >> 1) & 1), pk = NSD == 3 ? ek + (j >> 2) : 0;
      F[L.nu + pi + (int64_t)pj * L.PX + (int64_t)pk * L.PX * L.PY] += f;
    }
  }
  (void)nK;
}

// ------------------------------------------------------------------ Dirichlet handling
__global__ void bc_mark_kernel(int nbc, const int *idx, const double *val, unsigned char *isbc, double *g)
{
  int t = blockIdx.x * blockDim.x + threadIdx.x; if (t >= nbc) return;
  isbc[idx[t]] = 1; g[idx[t]] = val[t];
}
// MatZeroRowsColumns(A, bc, 1.0) keeping the pattern (femixedspace.c:2645, 2367); one warp per row
__global__ void zero_rows_cols_kernel(int64_t n, const int *ia, const int *ja, double *a, const unsigned char *isbc)
{
  int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; int lane = threadIdx.x & 31;
  if (row >= n) return;
  const bool rb = isbc[row];
  for (int k = ia[row] + lane; k < ia[row + 1]; k += 32) {
    int col = ja[k];
    if (rb) a[k] = (col == row) ? 1.0 : 0.0;
    else if (isbc[col]) a[k] = 0.0;
  }
}
// F[bc] = g[bc]; F += rhs_diri where rhs_diri = -(A_raw g), zero at bc (exSaddle.c:278-281, femixedspace.c:2638-2642)
__global__ void rhs_bc_kernel(int64_t n, const unsigned char *isbc, const double *g, const double *Ag, double *F)
{
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; if (i >= n) return;
  if (isbc[i]) F[i] = g[i] + 0.0; else F[i] = F[i] + 1.0 * (-1.0 * Ag[i]);
}

// ------------------------------------------------------------------ sub-block extraction
__global__ void a00_len_kernel(Lattice L, int64_t *len)
{
  int64_t nd = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; if (nd >= L.nun) return;
  BoxPattern p{L.NX, L.NY, L.NZ, 1};
  len[nd] = box_size(p, (int)(nd % L.NX), (int)((nd / L.NX) % L.NY), (int)(nd / ((int64_t)L.NX * L.NY)));
}
// one warp per velocity node: block columns + NSD x NSD blocks copied from the AIJ rows
__global__ void a00_fill_kernel(Lattice L, const int *ia, const int *ja, const double *a, const int *bia, int *bja, double *ba)
{
  int64_t nd = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; int lane = threadIdx.x & 31;
  if (nd >= L.nun) return;
  const int nsd = L.nsd, bs2 = nsd * nsd, nb = bia[nd + 1] - bia[nd];
  const int r0 = ia[nsd * nd];
  for (int s = lane; s < nb; s += 32) bja[bia[nd] + s] = ja[r0 + nsd * s] / nsd;
  for (int t = lane; t < nb * bs2; t += 32) {
    int s = t / bs2, r = t - s * bs2, x = r / nsd, y = r - x * nsd;
    ba[(int64_t)bia[nd] * bs2 + t] = a[ia[nsd * nd + x] + nsd * s + y];
  }
}
// generic contiguous column-range extraction: rows [r0, r0+nr), entries [off_lo(row), off_hi(row)) of each row
__global__ void sub_len_kernel(Lattice L, int64_t r0, int64_t nr, int pcols, int64_t *len)
{
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; if (t >= nr) return;
  RowBox b; int comp; row_to_box(L, r0 + t, b, &comp);
  len[t] = pcols ? b.ncp : (int64_t)L.nsd * b.ncu;
}
__global__ void sub_fill_kernel(Lattice L, int64_t r0, int64_t nr, int pcols, const int *ia, const int *ja, const double *a,
                                const int *sia, int *sja, double *sa)
{
  int64_t t = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; int lane = threadIdx.x & 31;
  if (t >= nr) return;
  RowBox b; int comp; row_to_box(L, r0 + t, b, &comp);
  const int src = ia[r0 + t] + (pcols ? L.nsd * b.ncu : 0), cnt = sia[t + 1] - sia[t], shift = pcols ? (int)L.nu : 0;
  for (int k = lane; k < cnt; k += 32) { sja[sia[t] + k] = ja[src + k] - shift; sa[sia[t] + k] = a[src + k]; }
}

// closed-form columns of a sub-block row (operator-free mode: there is no AIJ row to copy from)
__global__ void sub_ja_kernel(Lattice L, int64_t r0, int64_t nr, int pcols, const int *sia, int *sja)
{
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; if (t >= nr) return;
  RowBox b; int comp; row_to_box(L, r0 + t, b, &comp);
  int c = sia[t];
  if (pcols) { for (int kk = b.plo[2]; kk <= b.phi[2]; ++kk) for (int jj = b.plo[1]; jj <= b.phi[1]; ++jj) for (int ii = b.plo[0]; ii <= b.phi[0]; ++ii) sja[c++] = ii + jj * L.PX + kk * L.PX * L.PY; }
  else { for (int kk = b.ulo[2]; kk <= b.uhi[2]; ++kk) for (int jj = b.ulo[1]; jj <= b.uhi[1]; ++jj) for (int ii = b.ulo[0]; ii <= b.uhi[0]; ++ii) for (int d = 0; d < L.nsd; ++d) sja[c++] = L.nsd * (ii + jj * L.NX + kk * L.NX * L.NY) + d; }
}
// MatZeroRowsColumns on the off-diagonal blocks: A01 loses the rows, A10 the columns of the constrained velocity dofs
__global__ void zero_split_kernel(int64_t nr, int rows_are_u, const int *ia, const int *ja, double *a, const unsigned char *isbc)
{
  int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; int lane = threadIdx.x & 31;
  if (row >= nr) return;
  const bool rb = rows_are_u && isbc[row];
  for (int k = ia[row] + lane; k < ia[row + 1]; k += 32) if (rb || (!rows_are_u && isbc[ja[k]])) a[k] = 0.0;
}

// ------------------------------------------------------------------ helpers
static int scan_to_ia(xsb_ctx c, int64_t n, int64_t *len64 /* n+1, device, len in [0,n) */, int **ia_out, int64_t *total)
{
  // exclusive scan in place (int64), check 32-bit range, narrow
  void *tmp = nullptr; size_t tb = 0;
  CUDA_OK(cudaMemsetAsync(len64 + n, 0, sizeof(int64_t), c->stream));
  CUDA_OK(cub::DeviceScan::ExclusiveSum(nullptr, tb, len64, len64, n + 1, c->stream));
  CUDA_OK(cudaMalloc(&tmp, tb));
  CUDA_OK(cub::DeviceScan::ExclusiveSum(tmp, tb, len64, len64, n + 1, c->stream));
  int64_t tot = 0;
  CUDA_OK(cudaMemcpyAsync(&tot, len64 + n, sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
  CUDA_OK(cudaStreamSynchronize(c->stream));
  CUDA_OK(cudaFree(tmp));
  if (tot >= (int64_t)INT32_MAX) return xsb_fail(c, XSB_ERR_SUP, "matrix has %lld nonzeros >= 2^31: not representable with 32-bit PetscInt (use the matrix-free path)", (long long)tot);
  XSB_CHK(dev_alloc(c, ia_out, (size_t)n + 1));
  narrow_ia_kernel<<<(unsigned)((n + 1 + 255) / 256), 256, 0, c->stream>>>(n, len64, *ia_out); KERNEL_OK();
  *total = tot;
  return XSB_OK;
}

static inline unsigned nblk(int64_t n, int bs = 256) { return (unsigned)((n + bs - 1) / bs); }

int fe_assemble(xsb_ctx c)
{
  XSB_CHK(fe_resolve_model(c));
  Lattice &L = c->lat; const int nsd = c->nsd, nqp = nsd == 3 ? 27 : 9;
  cudaStream_t st = c->stream;
  // tables
  FeTables T; host_tables(L, T);
  FeTables *dT = nullptr; XSB_CHK(dev_alloc(c, (char **)&dT, sizeof(FeTables)));
  CUDA_OK(cudaMemcpyAsync(dT, &T, sizeof(T), cudaMemcpyHostToDevice, st));
  CUDA_OK(cudaStreamSynchronize(st));
  c->fe_tables = dT;
  // coefficients + Q1 projection
  const int64_t nq = L.nel * nqp;
  XSB_CHK(dev_alloc(c, &c->coeff, (size_t)XSB_NSLOT * nq));
  XSB_CHK(dev_alloc(c, &c->coeff_nodal, (size_t)XSB_NSLOT * L.npn));
  if (c->nodal_in) {   // coarse level of the monolithic -mg hierarchy: restricted nodal fields replace the model (femixedspace.c:2139-2215)
    CUDA_OK(cudaMemcpyAsync(c->coeff_nodal, c->nodal_in, sizeof(double) * XSB_NSLOT * L.npn, cudaMemcpyDeviceToDevice, st));
  } else {
    coeff_eval_kernel<<<nblk(nq), 256, 0, st>>>(L, c->mdl, c->lame, dT, c->coeff); KERNEL_OK();
    q1_project_kernel<<<nblk(L.npn, 128), 128, 0, st>>>(L, dT, c->coeff, c->coeff_nodal); KERNEL_OK();
  }
  q1_interp_kernel<<<nblk(nq), 256, 0, st>>>(L, dT, c->coeff_nodal, c->coeff); KERNEL_OK();
  // operator-free mode (-xsb_matrix_free full): A and A00 are never stored (128^3: nnz(A) = 1.13e10 > 2^31 and 137 GB)
  { const std::string mfv = c->opt.str("xsb_matrix_free", "0"); c->no_A = (mfv == "full" || mfv == "2"); }
  const bool split = c->no_A;
  if (split && nsd != 3) return xsb_fail(c, XSB_ERR_SUP, "-xsb_matrix_free full is implemented for the 3-D executables");
  // AIJ pattern
  int64_t *len64 = nullptr; CUDA_OK(cudaMalloc(&len64, sizeof(int64_t) * (L.n + 1)));
  row_len_kernel<<<nblk(L.n), 256, 0, st>>>(L, len64); KERNEL_OK();
  c->A.n = c->A.m = (int)L.n;
  c->A.inode_bs = nsd; c->A.inode_rows = L.nu;   // the NSD rows of a velocity node share one column pattern (row_box_u)
  SplitDst sp{nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  struct SubDef { Csr *S; int64_t r0, nr; int pcols; };
  SubDef subs[3] = {{&c->A01, 0, L.nu, 1}, {&c->A10, L.nu, L.np, 0}, {&c->A11, L.nu, L.np, 1}};
  if (!split) {
    { int rc = scan_to_ia(c, L.n, len64, &c->A.ia, &c->A.nnz); if (rc) { cudaFree(len64); return rc; } }
    XSB_CHK(dev_alloc(c, &c->A.ja, (size_t)c->A.nnz)); XSB_CHK(dev_alloc(c, &c->A.a, (size_t)c->A.nnz));
    fill_ja_kernel<<<nblk(L.n * 32), 256, 0, st>>>(L, c->A.ia, c->A.ja); KERNEL_OK();
    CUDA_OK(cudaMemsetAsync(c->A.a, 0, sizeof(double) * c->A.nnz, st));
  } else {
    {   // nnz(A) is still reported (xsb_get_sizes): total of the closed-form row lengths
      void *tmp = nullptr; size_t tb = 0; int64_t *tot = nullptr; CUDA_OK(cudaMalloc(&tot, sizeof(int64_t)));
      CUDA_OK(cub::DeviceReduce::Sum(nullptr, tb, len64, tot, L.n, st)); CUDA_OK(cudaMalloc(&tmp, tb));
      CUDA_OK(cub::DeviceReduce::Sum(tmp, tb, len64, tot, L.n, st));
      CUDA_OK(cudaMemcpyAsync(&c->A.nnz, tot, sizeof(int64_t), cudaMemcpyDeviceToHost, st)); CUDA_OK(cudaStreamSynchronize(st));
      CUDA_OK(cudaFree(tmp)); CUDA_OK(cudaFree(tot));
    }
    for (auto &s : subs) {
      sub_len_kernel<<<nblk(s.nr), 256, 0, st>>>(L, s.r0, s.nr, s.pcols, len64); KERNEL_OK();
      s.S->n = (int)s.nr; s.S->m = s.pcols ? (int)L.np : (int)L.nu;
      { int rc = scan_to_ia(c, s.nr, len64, &s.S->ia, &s.S->nnz); if (rc) { cudaFree(len64); return rc; } }
      XSB_CHK(dev_alloc(c, &s.S->ja, (size_t)s.S->nnz)); XSB_CHK(dev_alloc(c, &s.S->a, (size_t)s.S->nnz));
      sub_ja_kernel<<<nblk(s.nr), 256, 0, st>>>(L, s.r0, s.nr, s.pcols, s.S->ia, s.S->ja); KERNEL_OK();
      CUDA_OK(cudaMemsetAsync(s.S->a, 0, sizeof(double) * s.S->nnz, st));
    }
    sp = SplitDst{c->A01.ia, c->A10.ia, c->A11.ia, c->A01.a, c->A10.a, c->A11.a};
  }
  // Mp pattern
  mp_len_kernel<<<nblk(L.npn), 256, 0, st>>>(L, len64); KERNEL_OK();
  c->Mp.n = c->Mp.m = (int)L.npn;
  { int rc = scan_to_ia(c, L.npn, len64, &c->Mp.ia, &c->Mp.nnz); if (rc) { cudaFree(len64); return rc; } }
  XSB_CHK(dev_alloc(c, &c->Mp.ja, (size_t)c->Mp.nnz)); XSB_CHK(dev_alloc(c, &c->Mp.a, (size_t)c->Mp.nnz));
  mp_ja_kernel<<<nblk(L.npn), 256, 0, st>>>(L, c->Mp.ia, c->Mp.ja); KERNEL_OK();
  CUDA_OK(cudaMemsetAsync(c->Mp.a, 0, sizeof(double) * c->Mp.nnz, st));
  XSB_CHK(dev_alloc(c, &c->F, (size_t)L.n));
  CUDA_OK(cudaMemsetAsync(c->F, 0, sizeof(double) * L.n, st));
  // coloured element assembly
  for (int col = 0; col < (nsd == 3 ? 8 : 4); ++col) {
    const int ci = col & 1, cj = (col >> 1) & 1, ck = (col >> 2) & 1;
    const int64_t nei = (L.mx - ci + 1) / 2, nej = (L.my - cj + 1) / 2, nek = nsd == 3 ? (L.mz - ck + 1) / 2 : 1;
    const int64_t ne = nei * nej * nek;
    if (ne <= 0) continue;
    if (ne > 0x7fffffff) return xsb_fail(c, XSB_ERR_SUP, "too many elements per colour");
    if (split) assemble_kernel<3, true><<<(unsigned)ne, 256, 0, st>>>(L, c->lame, col, dT, c->coeff, nullptr, nullptr, c->Mp.ia, c->Mp.a, c->F, sp);
    else if (nsd == 3) assemble_kernel<3, false><<<(unsigned)ne, 256, 0, st>>>(L, c->lame, col, dT, c->coeff, c->A.ia, c->A.a, c->Mp.ia, c->Mp.a, c->F, sp);
    else assemble_kernel<2, false><<<(unsigned)ne, 128, 0, st>>>(L, c->lame, col, dT, c->coeff, c->A.ia, c->A.a, c->Mp.ia, c->Mp.a, c->F, sp);
    KERNEL_OK();
  }
  // Dirichlet data: index list built on the host (integer logic of ISCreate_BCList), applied on the device
  {
    const int zlo = c->slab.e0 == 0, zhi = c->slab.e1 == c->slab.mz_glob;   // which z faces of the local lattice are physical boundaries
    int cap = bc_list_faces(nsd, c->lame, c->mdl.model, c->mdl.freeslip, L.mx, L.my, L.mz, zlo, zhi, nullptr, nullptr, 0);
    std::vector<int32_t> idx(cap > 0 ? cap : 1); std::vector<double> val(cap > 0 ? cap : 1);
    c->nbc = bc_list_faces(nsd, c->lame, c->mdl.model, c->mdl.freeslip, L.mx, L.my, L.mz, zlo, zhi, idx.data(), val.data(), cap);
    if (c->mdl.bc_type == BC_MMS1) {   // values from the coordinates (models.c:505-593)
      for (int t = 0; t < c->nbc; ++t) {
        int64_t nd = idx[t] / 2; int d = idx[t] % 2; double x = L.hu[0] * (nd % L.NX), y = L.hu[1] * (nd / L.NX);
        val[t] = d == 0 ? 20 * x * y * y * y : 5 * (x * x * x * x - y * y * y * y);
      }
    }
    XSB_CHK(dev_alloc(c, &c->bc_idx, (size_t)(c->nbc > 0 ? c->nbc : 1))); XSB_CHK(dev_alloc(c, &c->bc_val, (size_t)(c->nbc > 0 ? c->nbc : 1)));
    XSB_CHK(dev_alloc(c, &c->isbc, (size_t)L.n));
    CUDA_OK(cudaMemsetAsync(c->isbc, 0, L.n, st));
    double *g = nullptr, *Ag = nullptr; XSB_CHK(dev_alloc(c, &g, (size_t)L.n)); XSB_CHK(dev_alloc(c, &Ag, (size_t)L.n));
    CUDA_OK(cudaMemsetAsync(g, 0, sizeof(double) * L.n, st));
    if (c->nbc > 0) {
      CUDA_OK(cudaMemcpyAsync(c->bc_idx, idx.data(), sizeof(int) * c->nbc, cudaMemcpyHostToDevice, st));
      CUDA_OK(cudaMemcpyAsync(c->bc_val, val.data(), sizeof(double) * c->nbc, cudaMemcpyHostToDevice, st));
      bc_mark_kernel<<<nblk(c->nbc), 256, 0, st>>>(c->nbc, c->bc_idx, c->bc_val, c->isbc, g); KERNEL_OK();
    }
    if (!split) {
      XSB_CHK(spmv_csr(c, c->A, g, Ag));   // rhs_diri uses A before rows/columns are zeroed (:2639)
    } else {
      // A_raw g with g_p = 0: velocity rows by the unmasked element kernel, pressure rows by A10 before its columns go
      bool any = false; for (int t = 0; t < c->nbc; ++t) any = any || val[t] != 0.0;
      CUDA_OK(cudaMemsetAsync(Ag, 0, sizeof(double) * L.n, st));
      if (any) { XSB_CHK(mf_a00_apply_raw(c, g, Ag)); XSB_CHK(spmv_csr(c, c->A10, g, Ag + L.nu)); }
    }
    rhs_bc_kernel<<<nblk(L.n), 256, 0, st>>>(L.n, c->isbc, g, Ag, c->F); KERNEL_OK();
    if (!split) { zero_rows_cols_kernel<<<nblk(L.n * 32), 256, 0, st>>>(L.n, c->A.ia, c->A.ja, c->A.a, c->isbc); KERNEL_OK(); }
    else {
      zero_split_kernel<<<nblk(L.nu * 32), 256, 0, st>>>(L.nu, 1, c->A01.ia, c->A01.ja, c->A01.a, c->isbc); KERNEL_OK();
      zero_split_kernel<<<nblk(L.np * 32), 256, 0, st>>>(L.np, 0, c->A10.ia, c->A10.ja, c->A10.a, c->isbc); KERNEL_OK();
    }
    CUDA_OK(cudaStreamSynchronize(st));
  }
  // sub-blocks: A00 as BAIJ(nsd), A01/A10/A11 as CSR
  if (split) {   // shell of the velocity block: sizes and pattern descriptor only (products go through xsb_mf.cu)
    Baij &B = c->A00; B.nb = (int)L.nun; B.bs = nsd; B.pat = BoxPattern{L.NX, L.NY, L.NZ, 1}; B.nblk = 0;
  } else {
    a00_len_kernel<<<nblk(L.nun), 256, 0, st>>>(L, len64); KERNEL_OK();
    Baij &B = c->A00; B.nb = (int)L.nun; B.bs = nsd; B.pat = BoxPattern{L.NX, L.NY, L.NZ, 1};
    { int rc = scan_to_ia(c, L.nun, len64, &B.ia, &B.nblk); if (rc) { cudaFree(len64); return rc; } }
    XSB_CHK(dev_alloc(c, &B.ja, (size_t)B.nblk)); XSB_CHK(dev_alloc(c, &B.a, (size_t)B.nblk * nsd * nsd + 2));   // +2: the tile kernel's last 16-byte load may straddle the end
    a00_fill_kernel<<<nblk(L.nun * 32), 256, 0, st>>>(L, c->A.ia, c->A.ja, c->A.a, B.ia, B.ja, B.a); KERNEL_OK();
    for (auto &s : subs) {
      sub_len_kernel<<<nblk(s.nr), 256, 0, st>>>(L, s.r0, s.nr, s.pcols, len64); KERNEL_OK();
      s.S->n = (int)s.nr; s.S->m = s.pcols ? (int)L.np : (int)L.nu;
      { int rc = scan_to_ia(c, s.nr, len64, &s.S->ia, &s.S->nnz); if (rc) { cudaFree(len64); return rc; } }
      XSB_CHK(dev_alloc(c, &s.S->ja, (size_t)s.S->nnz)); XSB_CHK(dev_alloc(c, &s.S->a, (size_t)s.S->nnz));
      sub_fill_kernel<<<nblk(s.nr * 32), 256, 0, st>>>(L, s.r0, s.nr, s.pcols, c->A.ia, c->A.ja, c->A.a, s.S->ia, s.S->ja, s.S->a); KERNEL_OK();
    }
  }
  CUDA_OK(cudaStreamSynchronize(st));
  CUDA_OK(cudaFree(len64));
  c->assembled = true;
  return XSB_OK;
}
