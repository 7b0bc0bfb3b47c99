/* xo_fe.c -- CPU oracle, part 1: the Q2-Q1 discretisation of exSaddle.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT (see xo.h).  Restates, function by function,
 * /root/reference/femixedspace.c and models.c for ONE rank (natural DMDA/DMComposite
 * numbering: all velocity dofs node-major/component-fastest, then all pressure dofs;
 * SURVEY.md App. A.1).  Every routine cites the reference lines it follows.
 */
#include "xo_internal.h"

/* ------------------------------------------------------------------ params */
void xo_params_init(xo_params *p, int nsd, int lame)
{
  memset(p, 0, sizeof(*p));
  p->nsd = nsd; p->lame = lame;
  p->mx = 4; p->my = -1; p->mz = -1;                /* exSaddle.c:140,179-180 */
  p->size[0] = p->size[1] = p->size[2] = 1.0;       /* exSaddle.c:143-145 */
  p->model = -1;
  p->c0 = p->c1 = p->lam0 = p->lam1 = NAN;
  p->sinker_r = NAN;
  p->sinker_c[0] = p->sinker_c[1] = p->sinker_c[2] = NAN;
  p->sinker_n = -1; p->solcx_xc = NAN; p->solcx_nz = -1; p->freeslip = 0;
}

static double dflt(double v, double d) { return isnan(v) ? d : v; }

/* Resolve model defaults exactly as the static option blocks in models.c do. */
static int resolve_params(xo_problem *P)
{
  xo_params *p = &P->prm;
  char *b = P->banner; size_t cap = sizeof(P->banner); int n = 0;
  if (p->nsd != 2 && p->nsd != 3) return xo_fail(P, "NSD must be 2 or 3 (exSaddle.h:7)");
  if (p->my < 0) p->my = p->mx;                      /* exSaddle.c:179 */
  if (p->mz < 0) p->mz = p->mx;
  if (p->nsd == 2) p->mz = 0;
  if (p->model < 0) p->model = p->lame ? 6 : 2;      /* models.h:9-13 */
  /* BC type (models.c:610-648) */
  P->bc_type = XO_BC_SOLCX;
  if (p->lame && p->model == 8) P->bc_type = XO_BC_FIXEDBASE;
  if (p->lame && (p->model == 9 || p->model == 10)) P->bc_type = XO_BC_COMPRESSION;
  if (p->nsd == 3 && p->model == 11) P->bc_type = XO_BC_FIXEDBASE;
  if (p->lame && p->nsd == 3 && p->model == 12) P->bc_type = XO_BC_COMPRESSION2;
  if (!p->lame && p->nsd == 2 && p->model == 101) P->bc_type = XO_BC_MMS1;
  {
    static const char *nm[] = {"SolCx", "FixedBase", "Compression", "Compression2", "StokesMMS1"};
    n += snprintf(b + n, cap - n, "Boundary Conditions: %s\n", nm[P->bc_type]);
  }
  if (p->lame) {
    switch (p->model) {
    case 2: /* models.c:741-759 */
      p->c0 = dflt(p->c0, 1.0); p->c1 = dflt(p->c1, 1.0); p->lam0 = dflt(p->lam0, 1.0); p->lam1 = dflt(p->lam1, 1.0);
      p->sinker_r = dflt(p->sinker_r, 0.05); if (p->sinker_n < 0) p->sinker_n = 3;
      n += snprintf(b + n, cap - n, "ModelType: LameXSinker\n  params: mu0 %1.4e\n  params: mu1 %1.4e\n  params: lambda0 %1.4e\n  params: lambda1 %1.4e\n  params: num sinkers %d\n  params: sinker radius %1.4e\n",
                    p->c0, p->c1, p->lam0, p->lam1, p->sinker_n, p->sinker_r);
      if (p->sinker_n > 8) return xo_fail(P, "Too many sinkers (models.c:763)");
      if (p->sinker_r > 0.05) return xo_fail(P, "Sinker Radius too big (models.c:766)");
      break;
    case 6: case 8: case 10: case 12: /* models.c:666-680 */
      p->c0 = dflt(p->c0, 1.0); p->c1 = dflt(p->c1, 1.0); p->lam0 = dflt(p->lam0, 1.0); p->lam1 = dflt(p->lam1, 2.0);
      p->sinker_r = dflt(p->sinker_r, 0.25);
      n += snprintf(b + n, cap - n, "ModelType: LameOneSinker\n  params: mu0 %1.4e\n  params: mu1 %1.4e\n  params: lambda0 %1.4e\n  params: lambda1 %1.4e\n  params: rad %1.4e\n",
                    p->c0, p->c1, p->lam0, p->lam1, p->sinker_r);
      break;
    case 9: /* models.c:822-827 */
      p->c0 = dflt(p->c0, 1.0); p->lam0 = dflt(p->lam0, 1.0);
      n += snprintf(b + n, cap - n, "ModelType: LameHomogeneous\n  params: mu0 %1.4e\n  params: lambda0 %1.4e\n", p->c0, p->lam0);
      break;
    default: return xo_fail(P, "Elasticity Model not implemented (models.c:876)");
    }
  } else {
    switch (p->model) {
    case 0: /* models.c:900-912 */
      p->c0 = dflt(p->c0, 1.0); p->c1 = dflt(p->c1, 1.0); p->solcx_xc = dflt(p->solcx_xc, 0.5); if (p->solcx_nz < 0) p->solcx_nz = 1;
      n += snprintf(b + n, cap - n, "ModelType: StokesSolCx\n  params: eta0 %1.4e\n  params: eta1 %1.4e\n  params: xc   %1.4e\n  params: nz   %d\n",
                    p->c0, p->c1, p->solcx_xc, p->solcx_nz);
      break;
    case 1: /* models.c:947-956 */
      p->c0 = dflt(p->c0, 1.0); p->c1 = dflt(p->c1, 1.0); p->sinker_r = dflt(p->sinker_r, 0.1);
      n += snprintf(b + n, cap - n, "ModelType: StokesThreeSinker\n  params: eta0 %1.4e\n  params: eta1 %1.4e\n  params: rad  %1.4e\n", p->c0, p->c1, p->sinker_r);
      break;
    case 2: /* models.c:1025-1046 */
      p->c0 = dflt(p->c0, 1.0); p->c1 = dflt(p->c1, 1.0); p->sinker_r = dflt(p->sinker_r, 0.05); if (p->sinker_n < 0) p->sinker_n = 3;
      n += snprintf(b + n, cap - n, "ModelType: StokesXSinker\n  params: eta0 %1.4e\n  params: eta1 %1.4e\n  params: num sinkers %d\n  params: sinker radius %1.4e\n",
                    p->c0, p->c1, p->sinker_n, p->sinker_r);
      if (p->sinker_n > 8) return xo_fail(P, "Too many sinkers (models.c:1041)");
      if (p->sinker_r > 0.05) return xo_fail(P, "Sinker Radius too big (models.c:1044)");
      break;
    case 5: /* models.c:1351-1365, 3-D only */
      if (p->nsd != 3) return xo_fail(P, "Stokes Model 5 is 3-D only (models.c:1499)");
      p->c0 = dflt(p->c0, 1.0); p->c1 = dflt(p->c1, 1.0); p->solcx_xc = dflt(p->solcx_xc, 0.5); if (p->solcx_nz < 0) p->solcx_nz = 1;
      n += snprintf(b + n, cap - n, "ModelType: StokesSolCx3d\n  params: eta0 %1.4e\n  params: eta1 %1.4e\n  params: xc   %1.4e\n  params: nz   %d\n  params: nz2  %d\n",
                    p->c0, p->c1, p->solcx_xc, p->solcx_nz, 1);
      break;
    case 6: /* models.c:1099-1123 */
      p->c0 = dflt(p->c0, 1.0); p->c1 = dflt(p->c1, 1.0); p->sinker_r = dflt(p->sinker_r, 0.25);
      p->sinker_c[0] = dflt(p->sinker_c[0], 0.5); p->sinker_c[1] = dflt(p->sinker_c[1], 0.5); p->sinker_c[2] = dflt(p->sinker_c[2], 0.5);
      n += snprintf(b + n, cap - n, "ModelType: StokesOneSinker\n  params: eta0 %1.4e\n  params: eta1 %1.4e\n  params: x %1.4e\n  params: y %1.4e\n", p->c0, p->c1, p->sinker_c[0], p->sinker_c[1]);
      if (p->nsd == 3) n += snprintf(b + n, cap - n, "  params: z %1.4e\n", p->sinker_c[2]);
      n += snprintf(b + n, cap - n, "  params: rad %1.4e\n", p->sinker_r);
      break;
    case 11: /* models.c:1453-1459, 3-D only */
      if (p->nsd != 3) return xo_fail(P, "Stokes Model 11 is 3-D only (models.c:1507)");
      p->c0 = dflt(p->c0, 1.0); p->c1 = dflt(p->c1, 10000.0);
      n += snprintf(b + n, cap - n, "ModelType: PseudoIce\n  params: eta0 %1.4e\n  params: eta1 %1.4e\n", p->c0, p->c1);
      break;
    case 101: /* models.c:1393-1395, 2-D only */
      if (p->nsd != 2) return xo_fail(P, "Stokes Model 101 is 2-D only (models.c:1515)");
      n += snprintf(b + n, cap - n, "ModelType: StokesMMS1\n");
      break;
    default: return xo_fail(P, "Stokes Model not implemented (models.c:1521; model 7 needs srand-based sinkers, out of scope)");
    }
  }
  return 0;
}

/* ---------------------------------------------------------- coefficients */
/* models.c:855-881 (Lame) and :1482-1525 (Stokes): out[] in XO_C_* slots */
static void eval_coeff(const xo_params *p, const double x[3], double out[XO_NSLOT])
{
  static const double posx[8] = {0.27, 0.6, 0.7, 0.2, 0.85, 0.4, 0.16, 0.55};   /* models.c:728-731, 1012-1015 */
  static const double posy[8] = {0.63, 0.83, 0.33, 0.2, 0.65, 0.3, 0.84, 0.54};
  static const double posz[8] = {0.50, 0.40, 0.30, 0.70, 0.65, 0.4, 0.8, 0.50};
  const int d3 = (p->nsd == 3);
  double c = p->c0, lam = p->lam0, rho = 1.0;
  int inside = 0, i;
  for (i = 0; i < XO_NSLOT; ++i) out[i] = 0.0;
  if (p->lame) {
    if (p->model == 2) { /* models.c:770-790 */
      for (i = 0; i < p->sinker_n; ++i) {
        double d2 = (x[0] - posx[i]) * (x[0] - posx[i]) + (x[1] - posy[i]) * (x[1] - posy[i]);
        if (d3) d2 += (x[2] - posz[i]) * (x[2] - posz[i]);
        if (d2 < p->sinker_r * p->sinker_r) { inside = 1; break; }
      }
      if (inside) { c = p->c1; lam = p->lam1; rho = 1.1; }
    } else if (p->model == 9) { /* models.c:831-833 */
    } else { /* one sinker: models.c:684-701 */
      double s2 = (x[0] - 0.5) * (x[0] - 0.5) + (x[1] - 0.5) * (x[1] - 0.5);
      if (d3) s2 += (x[2] - 0.5) * (x[2] - 0.5);
      if (s2 < p->sinker_r * p->sinker_r) { rho = 2.0; c = p->c1; lam = p->lam1; }
    }
    out[XO_C_ETA] = c; out[XO_C_LAM] = lam; out[XO_C_FU1] = -rho;
    return;
  }
  switch (p->model) {
  case 0: /* models.c:917-927 */
    if (x[0] > p->solcx_xc) c = p->c1;
    out[XO_C_ETA] = c;
    out[XO_C_FU1] = sin(p->solcx_nz * M_PI * x[1]) * cos(1.0 * M_PI * x[0]);
    return;
  case 5: /* models.c:1370-1378 */
    if (x[0] > p->solcx_xc) c = p->c1;
    out[XO_C_ETA] = c;
    out[XO_C_FU1] = sin(p->solcx_nz * M_PI * x[1]) * cos(1.0 * M_PI * x[0]) * sin(1 * M_PI * x[2]);
    return;
  case 1: { /* models.c:960-989 */
    static const double cx[3] = {0.27, 0.6, 0.7}, cy[3] = {0.63, 0.83, 0.33};
    for (i = 0; i < 3; ++i) {
      double s2 = (x[0] - cx[i]) * (x[0] - cx[i]) + (x[1] - cy[i]) * (x[1] - cy[i]);
      if (d3) s2 += (x[2] - 0.5) * (x[2] - 0.5);
      if (s2 < p->sinker_r * p->sinker_r) inside = 1;
    }
    break; }
  case 2: /* models.c:1048-1066 */
    for (i = 0; i < p->sinker_n; ++i) {
      double d2 = (x[0] - posx[i]) * (x[0] - posx[i]) + (x[1] - posy[i]) * (x[1] - posy[i]);
      if (d3) d2 += (x[2] - posz[i]) * (x[2] - posz[i]);
      if (d2 < p->sinker_r * p->sinker_r) { inside = 1; break; }
    }
    break;
  case 6: { /* models.c:1127-1143 */
    double s2 = (x[0] - p->sinker_c[0]) * (x[0] - p->sinker_c[0]) + (x[1] - p->sinker_c[1]) * (x[1] - p->sinker_c[1]);
    if (d3) s2 += (x[2] - p->sinker_c[2]) * (x[2] - p->sinker_c[2]);
    if (s2 < p->sinker_r * p->sinker_r) inside = 1;
    break; }
  case 11: { /* models.c:1464-1474; size_x re-read from the options DB */
    double xrel = x[0] / p->size[0];
    out[XO_C_ETA] = xrel * p->c0 + (1 - xrel) * p->c1;
    out[XO_C_FU2] = 1.0;
    return; }
  case 101: /* models.c:1424-1433 */
    out[XO_C_ETA] = 1.0;
    return;
  }
  if (inside) { c = p->c1; rho = 1.1; }
  out[XO_C_ETA] = c; out[XO_C_FU1] = -rho;
}

/* ------------------------------------------------------------ basis (A.2) */
/* femixedspace.c:1489-1512 */
static void basis_q1(int nsd, const double *xi_, double *N)
{
  double xi = xi_[0], eta = xi_[1];
  if (nsd == 2) {
    N[0] = 0.25 * (1.0 - xi) * (1.0 - eta); N[1] = 0.25 * (1.0 + xi) * (1.0 - eta);
    N[2] = 0.25 * (1.0 - xi) * (1.0 + eta); N[3] = 0.25 * (1.0 + xi) * (1.0 + eta);
  } else {
    double zeta = xi_[2];
    N[0] = 0.125 * (1.0 - xi) * (1.0 - eta) * (1.0 - zeta); N[1] = 0.125 * (1.0 + xi) * (1.0 - eta) * (1.0 - zeta);
    N[2] = 0.125 * (1.0 - xi) * (1.0 + eta) * (1.0 - zeta); N[3] = 0.125 * (1.0 + xi) * (1.0 + eta) * (1.0 - zeta);
    N[4] = 0.125 * (1.0 - xi) * (1.0 - eta) * (1.0 + zeta); N[5] = 0.125 * (1.0 + xi) * (1.0 - eta) * (1.0 + zeta);
    N[6] = 0.125 * (1.0 - xi) * (1.0 + eta) * (1.0 + zeta); N[7] = 0.125 * (1.0 + xi) * (1.0 + eta) * (1.0 + zeta);
  }
}
/* femixedspace.c:1726-1783 */
static void dbasis_q1(int nsd, const double *xi_, double *Gx, double *Ge, double *Gz)
{
  double xi = xi_[0], eta = xi_[1];
  if (nsd == 2) {
    Gx[0] = -0.25 * (1.0 - eta); Gx[1] = 0.25 * (1.0 - eta); Gx[2] = -0.25 * (1.0 + eta); Gx[3] = 0.25 * (1.0 + eta);
    Ge[0] = -0.25 * (1.0 - xi);  Ge[1] = -0.25 * (1.0 + xi); Ge[2] = 0.25 * (1.0 - xi);   Ge[3] = 0.25 * (1.0 + xi);
  } else {
    double zeta = xi_[2];
    Gx[0] = -0.125 * (1.0 - eta) * (1.0 - zeta); Gx[1] = 0.125 * (1.0 - eta) * (1.0 - zeta);
    Gx[2] = -0.125 * (1.0 + eta) * (1.0 - zeta); Gx[3] = 0.125 * (1.0 + eta) * (1.0 - zeta);
    Gx[4] = -0.125 * (1.0 - eta) * (1.0 + zeta); Gx[5] = 0.125 * (1.0 - eta) * (1.0 + zeta);
    Gx[6] = -0.125 * (1.0 + eta) * (1.0 + zeta); Gx[7] = 0.125 * (1.0 + eta) * (1.0 + zeta);
    Ge[0] = -0.125 * (1.0 - xi) * (1.0 - zeta); Ge[1] = -0.125 * (1.0 + xi) * (1.0 - zeta);
    Ge[2] = 0.125 * (1.0 - xi) * (1.0 - zeta);  Ge[3] = 0.125 * (1.0 + xi) * (1.0 - zeta);
    Ge[4] = -0.125 * (1.0 - xi) * (1.0 + zeta); Ge[5] = -0.125 * (1.0 + xi) * (1.0 + zeta);
    Ge[6] = 0.125 * (1.0 - xi) * (1.0 + zeta);  Ge[7] = 0.125 * (1.0 + xi) * (1.0 + zeta);
    Gz[0] = -0.125 * (1.0 - xi) * (1.0 - eta); Gz[1] = -0.125 * (1.0 + xi) * (1.0 - eta);
    Gz[2] = -0.125 * (1.0 - xi) * (1.0 + eta); Gz[3] = -0.125 * (1.0 + xi) * (1.0 + eta);
    Gz[4] = 0.125 * (1.0 - xi) * (1.0 - eta);  Gz[5] = 0.125 * (1.0 + xi) * (1.0 - eta);
    Gz[6] = 0.125 * (1.0 - xi) * (1.0 + eta);  Gz[7] = 0.125 * (1.0 + xi) * (1.0 + eta);
  }
}
/* femixedspace.c:1515-1556 */
static void basis_q2(int nsd, const double *xi_, double *N)
{
  if (nsd == 2) {
    double xi = xi_[0], eta = xi_[1];
    N[0] = 0.5 * eta * (eta - 1.0) * 0.5 * xi * (xi - 1.0);
    N[1] = 0.5 * eta * (eta - 1.0) * (1.0 + xi) * (1.0 - xi);
    N[2] = 0.5 * eta * (eta - 1.0) * 0.5 * (1.0 + xi) * xi;
    N[3] = (1.0 + eta) * (1.0 - eta) * 0.5 * xi * (xi - 1.0);
    N[4] = (1.0 + eta) * (1.0 - eta) * (1.0 + xi) * (1.0 - xi);
    N[5] = (1.0 + eta) * (1.0 - eta) * 0.5 * (1.0 + xi) * xi;
    N[6] = 0.5 * (1.0 + eta) * eta * 0.5 * xi * (xi - 1.0);
    N[7] = 0.5 * (1.0 + eta) * eta * (1.0 + xi) * (1.0 - xi);
    N[8] = 0.5 * (1.0 + eta) * eta * 0.5 * (1.0 + xi) * xi;
  } else {
    double b[3][3]; int d, i, j, k, cnt = 0;
    for (d = 0; d < 3; ++d) {
      double xi = xi_[d];
      b[d][0] = 0.5 * xi * (xi - 1.0); b[d][1] = (1.0 + xi) * (1.0 - xi); b[d][2] = 0.5 * (1.0 + xi) * xi;
    }
    for (k = 0; k < 3; ++k) for (j = 0; j < 3; ++j) for (i = 0; i < 3; ++i) N[cnt++] = b[0][i] * b[1][j] * b[2][k];
  }
}
/* femixedspace.c:1786-1855 */
static void dbasis_q2(int nsd, const double *xi_, double *Gx, double *Ge, double *Gz)
{
  if (nsd == 2) {
    double xi = xi_[0], eta = xi_[1];
    Gx[0] = 0.5 * eta * (eta - 1.0) * (xi - 0.5);
    Gx[1] = 0.5 * eta * (eta - 1.0) * (-2.0 * xi);
    Gx[2] = 0.5 * eta * (eta - 1.0) * 0.5 * (1.0 + 2.0 * xi);
    Gx[3] = (1.0 + eta) * (1.0 - eta) * (xi - 0.5);
    Gx[4] = (1.0 + eta) * (1.0 - eta) * (-2.0 * xi);
    Gx[5] = (1.0 + eta) * (1.0 - eta) * 0.5 * (1.0 + 2.0 * xi);
    Gx[6] = 0.5 * (1.0 + eta) * eta * (xi - 0.5);
    Gx[7] = 0.5 * (1.0 + eta) * eta * (-2.0 * xi);
    Gx[8] = 0.5 * (1.0 + eta) * eta * 0.5 * (1.0 + 2.0 * xi);
    Ge[0] = (eta - 0.5) * 0.5 * xi * (xi - 1.0);
    Ge[1] = (eta - 0.5) * (1.0 + xi) * (1.0 - xi);
    Ge[2] = (eta - 0.5) * 0.5 * (1.0 + xi) * xi;
    Ge[3] = (-2.0 * eta) * 0.5 * xi * (xi - 1.0);
    Ge[4] = (-2.0 * eta) * (1.0 + xi) * (1.0 - xi);
    Ge[5] = (-2.0 * eta) * 0.5 * (1.0 + xi) * xi;
    Ge[6] = 0.5 * (1.0 + 2.0 * eta) * 0.5 * xi * (xi - 1.0);
    Ge[7] = 0.5 * (1.0 + 2.0 * eta) * (1.0 + xi) * (1.0 - xi);
    Ge[8] = 0.5 * (1.0 + 2.0 * eta) * 0.5 * (1.0 + xi) * xi;
  } else {
    double b[3][3], g[3][3]; int d, i, j, k, cnt = 0;
    for (d = 0; d < 3; ++d) {
      double xi = xi_[d];
      b[d][0] = 0.5 * xi * (xi - 1.0); b[d][1] = (1.0 + xi) * (1.0 - xi); b[d][2] = 0.5 * (1.0 + xi) * xi;
      g[d][0] = 0.5 * (2.0 * xi - 1.0); g[d][1] = -2.0 * xi; g[d][2] = 0.5 * (2.0 * xi + 1.0);
    }
    for (k = 0; k < 3; ++k) for (j = 0; j < 3; ++j) for (i = 0; i < 3; ++i) {
      Gx[cnt] = g[0][i] * b[1][j] * b[2][k];
      Ge[cnt] = b[0][i] * g[1][j] * b[2][k];
      Gz[cnt] = b[0][i] * b[1][j] * g[2][k];
      ++cnt;
    }
  }
}

/* femixedspace.c:1559-1612 (detJ only; keeps the reference's "+ J12*J20" form at :1608) */
static double basis_transformation(int nsd, int nb, const double *Gx, const double *Ge, const double *Gz, const double *co)
{
  int k;
  if (nsd == 2) {
    double J[2][2] = {{0, 0}, {0, 0}};
    for (k = 0; k < nb; ++k) {
      double xc = co[2 * k], yc = co[2 * k + 1];
      J[0][0] += Gx[k] * xc; J[0][1] += Gx[k] * yc; J[1][0] += Ge[k] * xc; J[1][1] += Ge[k] * yc;
    }
    return J[0][0] * J[1][1] - J[0][1] * J[1][0];
  } else {
    double J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    for (k = 0; k < nb; ++k) {
      double xc = co[3 * k], yc = co[3 * k + 1], zc = co[3 * k + 2];
      J[0][0] += Gx[k] * xc; J[0][1] += Gx[k] * yc; J[0][2] += Gx[k] * zc;
      J[1][0] += Ge[k] * xc; J[1][1] += Ge[k] * yc; J[1][2] += Ge[k] * zc;
      J[2][0] += Gz[k] * xc; J[2][1] += Gz[k] * yc; J[2][2] += Gz[k] * zc;
    }
    return J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) - J[0][1] * (J[1][0] * J[2][2] + J[1][2] * J[2][0]) +
           J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
  }
}

/* femixedspace.c:1615-1723 */
static double deriv_global(int nsd, int nb, const double *Gxi, const double *Geta, const double *Gzeta,
                           double *GNx, double *GNy, double *GNz, const double *co)
{
  int k;
  if (nsd == 2) {
    double J[2][2] = {{0, 0}, {0, 0}}, iJ[2][2], detJ;
    for (k = 0; k < nb; ++k) {
      double xc = co[2 * k], yc = co[2 * k + 1];
      J[0][0] += Gxi[k] * xc; J[0][1] += Gxi[k] * yc; J[1][0] += Geta[k] * xc; J[1][1] += Geta[k] * yc;
    }
    detJ = J[0][0] * J[1][1] - J[0][1] * J[1][0];
    iJ[0][0] = J[1][1] / detJ; iJ[0][1] = -J[0][1] / detJ; iJ[1][0] = -J[1][0] / detJ; iJ[1][1] = J[0][0] / detJ;
    for (k = 0; k < nb; ++k) {
      GNx[k] = Gxi[k] * iJ[0][0] + Geta[k] * iJ[0][1];
      GNy[k] = Gxi[k] * iJ[1][0] + Geta[k] * iJ[1][1];
    }
    return detJ;
  } else {
    double J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, iJ[3][3], t4, t6, t8, t10, t12, t14, t17;
    for (k = 0; k < nb; ++k) {
      double xc = co[3 * k], yc = co[3 * k + 1], zc = co[3 * k + 2];
      J[0][0] += Gxi[k] * xc; J[0][1] += Gxi[k] * yc; J[0][2] += Gxi[k] * zc;
      J[1][0] += Geta[k] * xc; J[1][1] += Geta[k] * yc; J[1][2] += Geta[k] * zc;
      J[2][0] += Gzeta[k] * xc; J[2][1] += Gzeta[k] * yc; J[2][2] += Gzeta[k] * zc;
    }
    t4 = J[2][0] * J[0][1]; t6 = J[2][0] * J[0][2]; t8 = J[1][0] * J[0][1];
    t10 = J[1][0] * J[0][2]; t12 = J[0][0] * J[1][1]; t14 = J[0][0] * J[1][2];
    t17 = 0.1e1 / (t4 * J[1][2] - t6 * J[1][1] - t8 * J[2][2] + t10 * J[2][1] + t12 * J[2][2] - t14 * J[2][1]);
    iJ[0][0] = (J[1][1] * J[2][2] - J[1][2] * J[2][1]) * t17;
    iJ[0][1] = -(J[0][1] * J[2][2] - J[0][2] * J[2][1]) * t17;
    iJ[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * t17;
    iJ[1][0] = -(-J[2][0] * J[1][2] + J[1][0] * J[2][2]) * t17;
    iJ[1][1] = (-t6 + J[0][0] * J[2][2]) * t17;
    iJ[1][2] = -(-t10 + t14) * t17;
    iJ[2][0] = (-J[2][0] * J[1][1] + J[1][0] * J[2][1]) * t17;
    iJ[2][1] = -(-t4 + J[0][0] * J[2][1]) * t17;
    iJ[2][2] = (-t8 + t12) * t17;
    for (k = 0; k < nb; ++k) {
      GNx[k] = iJ[0][0] * Gxi[k] + iJ[0][1] * Geta[k] + iJ[0][2] * Gzeta[k];
      GNy[k] = iJ[1][0] * Gxi[k] + iJ[1][1] * Geta[k] + iJ[1][2] * Gzeta[k];
      GNz[k] = iJ[2][0] * Gxi[k] + iJ[2][1] * Geta[k] + iJ[2][2] * Gzeta[k];
    }
    return J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) - J[0][1] * (J[1][0] * J[2][2] + J[1][2] * J[2][0]) +
           J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
  }
}

/* ------------------------------------------------------------------ mesh */
/* DMCreate_SaddleQ2Q1 (femixedspace.c:1136-1350) on one rank + node maps (:852-1039) */
static int build_mesh(xo_problem *P)
{
  const xo_params *p = &P->prm;
  const int nsd = p->nsd;
  int e, ei, ej, ek, ii, jj, kk, mzz = nsd == 3 ? p->mz : 1;
  if (p->mx < 1 || p->my < 1 || (nsd == 3 && p->mz < 1)) return xo_fail(P, "mx,my,mz must be >= 1");
  P->NX = 2 * p->mx + 1; P->NY = 2 * p->my + 1; P->NZ = nsd == 3 ? 2 * p->mz + 1 : 1;   /* :1157-1158 */
  P->PX = p->mx + 1;     P->PY = p->my + 1;     P->PZ = nsd == 3 ? p->mz + 1 : 1;       /* :1248-1249 */
  P->nun = (int64_t)P->NX * P->NY * P->NZ; P->npn = (int64_t)P->PX * P->PY * P->PZ;
  P->nu = nsd * P->nun; P->np = P->npn; P->n = P->nu + P->np;
  P->nel = (int64_t)p->mx * p->my * mzz;
  P->nbu = nsd == 3 ? 27 : 9; P->nbp = nsd == 3 ? 8 : 4; P->nqp = P->nbu;
  if (P->n >= INT32_MAX) return xo_fail(P, "problem too large for 32-bit PetscInt");
  P->u_map = (int *)malloc(sizeof(int) * P->nbu * P->nel);
  P->p_map = (int *)malloc(sizeof(int) * P->nbp * P->nel);
  e = 0;
  for (ek = 0; ek < mzz; ++ek) for (ej = 0; ej < p->my; ++ej) for (ei = 0; ei < p->mx; ++ei, ++e) {   /* :984-986 */
    int c = 0;
    for (kk = 0; kk < (nsd == 3 ? 3 : 1); ++kk) for (jj = 0; jj < 3; ++jj) for (ii = 0; ii < 3; ++ii)
      P->u_map[P->nbu * e + c++] = (2 * ei + ii) + (2 * ej + jj) * P->NX + (2 * ek + kk) * P->NX * P->NY;   /* :995-1029 */
    c = 0;
    for (kk = 0; kk < (nsd == 3 ? 2 : 1); ++kk) for (jj = 0; jj < 2; ++jj) for (ii = 0; ii < 2; ++ii)
      P->p_map[P->nbp * e + c++] = (ei + ii) + (ej + jj) * P->PX + (ek + kk) * P->PX * P->PY;               /* :914-921 */
  }
  /* DMDASetUniformCoordinates (:1353-1362): x_i = xmin + i*(xmax-xmin)/(M-1) */
  P->hu[0] = p->size[0] / (P->NX - 1); P->hu[1] = p->size[1] / (P->NY - 1); P->hu[2] = nsd == 3 ? p->size[2] / (P->NZ - 1) : 0;
  P->hp[0] = p->size[0] / (P->PX - 1); P->hp[1] = p->size[1] / (P->PY - 1); P->hp[2] = nsd == 3 ? p->size[2] / (P->PZ - 1) : 0;
  return 0;
}

static void u_coords(const xo_problem *P, int node, double *x)
{
  int i = node % P->NX, j = (node / P->NX) % P->NY, k = node / (P->NX * P->NY);
  x[0] = 0.0 + P->hu[0] * i; x[1] = 0.0 + P->hu[1] * j; if (P->prm.nsd == 3) x[2] = 0.0 + P->hu[2] * k;
}
static void p_coords(const xo_problem *P, int node, double *x)
{
  int i = node % P->PX, j = (node / P->PX) % P->PY, k = node / (P->PX * P->PY);
  x[0] = 0.0 + P->hp[0] * i; x[1] = 0.0 + P->hp[1] * j; if (P->prm.nsd == 3) x[2] = 0.0 + P->hp[2] * k;
}

/* FEMixedSpaceQuadratureCreate (femixedspace.c:1366-1408) + basis tabulation */
static void build_tables(xo_problem *P)
{
  static const double xi1d[3] = {-0.774596669241483, 0.0, 0.774596669241483};   /* :1379 (15-digit literals) */
  static const double wt1d[3] = {0.555555555555556, 0.888888888888889, 0.555555555555556};
  const int nsd = P->prm.nsd;
  int i, j, k, q = 0;
  for (k = 0; k < (nsd == 3 ? 3 : 1); ++k) for (j = 0; j < 3; ++j) for (i = 0; i < 3; ++i, ++q) {
    P->xi[q][0] = xi1d[i]; P->xi[q][1] = xi1d[j]; P->xi[q][2] = nsd == 3 ? xi1d[k] : 0.0;
    P->wq[q] = nsd == 3 ? wt1d[i] * wt1d[j] * wt1d[k] : wt1d[i] * wt1d[j];
  }
  for (q = 0; q < P->nqp; ++q) {
    basis_q2(nsd, P->xi[q], P->Nu[q]);
    dbasis_q2(nsd, P->xi[q], P->GNuxi[q], P->GNueta[q], P->GNuzeta[q]);
    basis_q1(nsd, P->xi[q], P->Np[q]);
    dbasis_q1(nsd, P->xi[q], P->GNpxi[q], P->GNpeta[q], P->GNpzeta[q]);
  }
}

/* FEMixedSpaceDefineQPwiseProperties (femixedspace.c:1857-1933) */
static void define_qp_properties(xo_problem *P)
{
  const int nsd = P->prm.nsd, nbu = P->nbu, nqp = P->nqp;
  int64_t e;
  P->coeff = (double *)calloc((size_t)P->nel * nqp * XO_NSLOT, sizeof(double));
#pragma omp parallel for schedule(static)
  for (e = 0; e < P->nel; ++e) {
    double co[27 * 3]; int i, d, q;
    for (i = 0; i < nbu; ++i) u_coords(P, P->u_map[nbu * e + i], &co[nsd * i]);
    for (q = 0; q < nqp; ++q) {
      double xq[3] = {0, 0, 0};
      for (i = 0; i < nbu; ++i) for (d = 0; d < nsd; ++d) xq[d] += P->Nu[q][i] * co[nsd * i + d];   /* :1905-1906 */
      eval_coeff(&P->prm, xq, &P->coeff[((size_t)e * nqp + q) * XO_NSLOT]);
    }
  }
}

/* nodal Q1 coefficient fields -> quadrature points (femixedspace.c:2036-2083 fine level, :2168-2215 coarse levels) */
static void interp_nodal_to_qp(xo_problem *P)
{
  const int nbp = P->nbp, nqp = P->nqp;
  const double *nodal = P->coeff_nodal;
  int64_t e;
#pragma omp parallel for schedule(static)
  for (e = 0; e < P->nel; ++e) {
    int i, q, s;
    for (q = 0; q < nqp; ++q) {
      double *c = &P->coeff[((size_t)e * nqp + q) * XO_NSLOT];
      for (s = 0; s < XO_NSLOT; ++s) {
        double v = 0.0;
        for (i = 0; i < nbp; ++i) v += P->Np[q][i] * nodal[(size_t)P->p_map[nbp * e + i] * XO_NSLOT + s];
        c[s] = v;
      }
    }
  }
}

/* FEMixedSpaceDefineQPwiseProperties_Q1Projection, fine level (femixedspace.c:1976-2083) */
static void q1_projection(xo_problem *P)
{
  const int nbp = P->nbp, nqp = P->nqp;
  double *nodal = (double *)calloc((size_t)P->npn * XO_NSLOT, sizeof(double));
  double *scale = (double *)calloc((size_t)P->npn, sizeof(double));
  int64_t e;
  for (e = 0; e < P->nel; ++e) {   /* sequential: ADD_VALUES in element order (:2005-2007) */
    double elc[XO_NSLOT][8], els[8]; int i, q, s;
    memset(elc, 0, sizeof(elc)); memset(els, 0, sizeof(els));
    for (q = 0; q < nqp; ++q) for (i = 0; i < nbp; ++i) {
      const double *c = &P->coeff[((size_t)e * nqp + q) * XO_NSLOT];
      for (s = 0; s < XO_NSLOT; ++s) elc[s][i] += P->Np[q][i] * c[s];
      els[i] += P->Np[q][i];
    }
    for (i = 0; i < nbp; ++i) {
      int nd = P->p_map[nbp * e + i];
      for (s = 0; s < XO_NSLOT; ++s) nodal[(size_t)nd * XO_NSLOT + s] += elc[s][i];
      scale[nd] += els[i];
    }
  }
  {
    int64_t nd; int s;
    for (nd = 0; nd < P->npn; ++nd) for (s = 0; s < XO_NSLOT; ++s) nodal[(size_t)nd * XO_NSLOT + s] /= scale[nd];   /* :2017 */
  }
  P->coeff_nodal = nodal;
  free(scale);
  interp_nodal_to_qp(P);
}

/* ------------------------------------------------------------ BC lists */
static void bc_push(xo_problem *P, int idx, double val)
{
  if (P->nbc == P->bc_cap) {
    P->bc_cap = P->bc_cap ? 2 * P->bc_cap : 1024;
    P->bc_idx = (int *)realloc(P->bc_idx, sizeof(int) * P->bc_cap);
    P->bc_val = (double *)realloc(P->bc_val, sizeof(double) * P->bc_cap);
  }
  P->bc_idx[P->nbc] = idx; P->bc_val[P->nbc] = val; P->nbc++;
}
static double mms1_solx(double x, double y) { return 20 * x * y * y * y; }             /* models.c:462 */
static double mms1_soly(double x, double y) { return 5 * (x * x * x * x - y * y * y * y); } /* models.c:463 */

/* ISCreate_BCList (models.c:610-648) with global=PETSC_TRUE on one rank: si=sj=sk=0, ni=M ... */
static void build_bc(xo_problem *P)
{
  const int nsd = P->prm.nsd, ni = P->NX, nj = P->NY, nk = P->NZ, M = P->NX, N = P->NY, Pz = P->NZ;
  int i, j, k, d;
#define IDX(i, j, k, d) (nsd * ((i) + (j) * ni + (k) * ni * nj) + (d))
  switch (P->bc_type) {
  case XO_BC_SOLCX:   /* models.c:8-158 */
    if (nsd == 2) {
      for (j = 0; j < nj; ++j) bc_push(P, IDX(0, j, 0, 0), 0.0);
      for (i = 0; i < ni; ++i) bc_push(P, IDX(i, 0, 0, 1), 0.0);
      for (j = 0; j < nj; ++j) bc_push(P, IDX(ni - 1, j, 0, 0), 0.0);
      if (P->prm.freeslip) for (i = 0; i < ni; ++i) bc_push(P, IDX(i, nj - 1, 0, 1), 0.0);
    } else {
      for (j = 0; j < nj; ++j) for (k = 0; k < nk; ++k) bc_push(P, IDX(0, j, k, 0), 0.0);
      for (i = 0; i < ni; ++i) for (k = 0; k < nk; ++k) bc_push(P, IDX(i, 0, k, 1), 0.0);
      for (i = 0; i < ni; ++i) for (j = 0; j < nj; ++j) bc_push(P, IDX(i, j, 0, 2), 0.0);
      for (j = 0; j < nj; ++j) for (k = 0; k < nk; ++k) bc_push(P, IDX(ni - 1, j, k, 0), 0.0);
      if (P->prm.freeslip) for (i = 0; i < ni; ++i) for (k = 0; k < nk; ++k) bc_push(P, IDX(i, nj - 1, k, 1), 0.0);
      for (i = 0; i < ni; ++i) for (j = 0; j < nj; ++j) bc_push(P, IDX(i, j, nk - 1, 2), 0.0);
    }
    break;
  case XO_BC_FIXEDBASE:   /* models.c:162-234 */
    for (d = 0; d < nsd; ++d) for (i = 0; i < ni; ++i) for (k = 0; k < nk; ++k) bc_push(P, IDX(i, 0, k, d), 0.0);
    break;
  case XO_BC_COMPRESSION:   /* models.c:239-337; NOTE the reference tests si+ni==N (y count), :270,315 */
    for (d = 0; d < nsd; ++d) for (j = 0; j < nj; ++j) for (k = 0; k < nk; ++k) bc_push(P, IDX(0, j, k, d), d == 0 ? 0.1 : 0.0);
    if (ni == N) for (d = 0; d < nsd; ++d) for (j = 0; j < nj; ++j) for (k = 0; k < nk; ++k) bc_push(P, IDX(ni - 1, j, k, d), d == 0 ? -0.1 : 0.0);
    break;
  case XO_BC_COMPRESSION2:   /* models.c:342-455 */
    for (j = 0; j < nj; ++j) for (k = 0; k < nk; ++k) bc_push(P, IDX(0, j, k, 0), 0.1);
    if (ni == N) for (j = 0; j < nj; ++j) for (k = 0; k < nk; ++k) bc_push(P, IDX(ni - 1, j, k, 0), -0.1);
    for (i = 0; i < ni; ++i) for (k = 0; k < nk; ++k) bc_push(P, IDX(i, 0, k, 1), 0.0);
    for (i = 0; i < ni; ++i) for (j = 0; j < nj; ++j) bc_push(P, IDX(i, j, 0, 2), 0.0);
    (void)Pz;
    for (i = 0; i < ni; ++i) for (j = 0; j < nj; ++j) bc_push(P, IDX(i, j, nk - 1, 2), 0.0);
    break;
  case XO_BC_MMS1:   /* models.c:466-603 (2-D): all four faces, both components; N/M swapped as in :496,:498 */
    for (j = 0; j < nj; ++j) { double c[3]; u_coords(P, 0 + j * ni, c);
      for (d = 0; d < 2; ++d) bc_push(P, IDX(0, j, 0, d), d == 0 ? mms1_solx(c[0], c[1]) : mms1_soly(c[0], c[1])); }
    if (ni == N) for (j = 0; j < nj; ++j) { double c[3]; u_coords(P, ni - 1 + j * ni, c);
      for (d = 0; d < 2; ++d) bc_push(P, IDX(ni - 1, j, 0, d), d == 0 ? mms1_solx(c[0], c[1]) : mms1_soly(c[0], c[1])); }
    for (i = 0; i < ni; ++i) { double c[3]; u_coords(P, i, c);
      for (d = 0; d < 2; ++d) bc_push(P, IDX(i, 0, 0, d), d == 0 ? mms1_solx(c[0], c[1]) : mms1_soly(c[0], c[1])); }
    if (nj == M) for (i = 0; i < ni; ++i) { double c[3]; u_coords(P, i + (nj - 1) * ni, c);
      for (d = 0; d < 2; ++d) bc_push(P, IDX(i, nj - 1, 0, d), d == 0 ? mms1_solx(c[0], c[1]) : mms1_soly(c[0], c[1])); }
    break;
  }
#undef IDX
}

/* ------------------------------------------------- pattern / preallocation */
/* SaddlePreallocation_SEQ (femixedspace.c:181-286): total of the per-row bounds */
static int64_t prealloc_total(const xo_problem *P)
{
  const int nsd = P->prm.nsd; int i, j, k; int64_t tot = 0, m = P->n;
  for (k = 0; k < P->NZ; ++k) for (j = 0; j < P->NY; ++j) for (i = 0; i < P->NX; ++i) {
    int64_t r;
    if (nsd == 2) { int vi = i % 2 == 0, vj = j % 2 == 0; r = (vi && vj) ? 2 * 25 + 9 : (vi || vj) ? 2 * 15 + 6 : 2 * 9 + 4; }
    else { int nmod = i % 2 + j % 2 + k % 2; r = nmod == 0 ? 3 * 125 + 27 : nmod == 1 ? 3 * 75 + 18 : nmod == 2 ? 3 * 45 + 12 : 3 * 27 + 8; }
    if (r > m) r = m;
    tot += nsd * r;
  }
  { int64_t r = nsd == 2 ? 2 * 25 + 9 : 3 * 125 + 27; if (r > m) r = m; tot += P->npn * r; }
  return tot;
}

static int cmp_int(const void *a, const void *b) { int x = *(const int *)a, y = *(const int *)b; return x < y ? -1 : x > y; }
static int uniq_sorted(int *v, int n) { int i, m = 0; for (i = 0; i < n; ++i) if (!m || v[i] != v[m - 1]) v[m++] = v[i]; return m; }

/* Pattern of MatAssemble_Saddle_NULL (femixedspace.c:2306-2370): union over elements of the four
   element blocks, columns sorted ascending.  Built from the element->node maps (node->element adjacency). */
static int build_pattern(xo_problem *P)
{
  const int nsd = P->prm.nsd, nbu = P->nbu, nbp = P->nbp;
  int64_t e, r; int i;
  int *uoff = (int *)calloc(P->nun + 1, sizeof(int)), *poff = (int *)calloc(P->npn + 1, sizeof(int));
  int *uel, *pel;
  for (e = 0; e < P->nel; ++e) { for (i = 0; i < nbu; ++i) uoff[P->u_map[nbu * e + i] + 1]++; for (i = 0; i < nbp; ++i) poff[P->p_map[nbp * e + i] + 1]++; }
  for (r = 0; r < P->nun; ++r) uoff[r + 1] += uoff[r];
  for (r = 0; r < P->npn; ++r) poff[r + 1] += poff[r];
  uel = (int *)malloc(sizeof(int) * uoff[P->nun]); pel = (int *)malloc(sizeof(int) * poff[P->npn]);
  { int *uc = (int *)calloc(P->nun, sizeof(int)), *pc = (int *)calloc(P->npn, sizeof(int));
    for (e = 0; e < P->nel; ++e) {
      for (i = 0; i < nbu; ++i) { int nd = P->u_map[nbu * e + i]; uel[uoff[nd] + uc[nd]++] = (int)e; }
      for (i = 0; i < nbp; ++i) { int nd = P->p_map[nbp * e + i]; pel[poff[nd] + pc[nd]++] = (int)e; }
    }
    free(uc); free(pc); }
  P->ia = (int *)malloc(sizeof(int) * (P->n + 1));
  /* pass 1: row lengths */
  {
    int64_t *len = (int64_t *)malloc(sizeof(int64_t) * (P->nun + P->npn));
#pragma omp parallel for schedule(static) private(i)
    for (r = 0; r < P->nun + P->npn; ++r) {
      int cu[8 * 27], cp[8 * 8], nu_ = 0, np_ = 0, t;
      const int isu = r < P->nun; const int nd = isu ? (int)r : (int)(r - P->nun);
      const int *els = isu ? &uel[uoff[nd]] : &pel[poff[nd]];
      const int nels = isu ? uoff[nd + 1] - uoff[nd] : poff[nd + 1] - poff[nd];
      for (t = 0; t < nels; ++t) { for (i = 0; i < nbu; ++i) cu[nu_++] = P->u_map[nbu * els[t] + i]; for (i = 0; i < nbp; ++i) cp[np_++] = P->p_map[nbp * els[t] + i]; }
      qsort(cu, nu_, sizeof(int), cmp_int); qsort(cp, np_, sizeof(int), cmp_int);
      len[r] = (int64_t)nsd * uniq_sorted(cu, nu_) + uniq_sorted(cp, np_);
    }
    { int64_t tot = 0, row = 0; int d;
      for (r = 0; r < P->nun; ++r) for (d = 0; d < nsd; ++d) { P->ia[row++] = (int)tot; tot += len[r]; if (tot >= INT32_MAX) { free(len); return xo_fail(P, "nnz exceeds 32-bit PetscInt"); } }
      for (r = P->nun; r < P->nun + P->npn; ++r) { P->ia[row++] = (int)tot; tot += len[r]; if (tot >= INT32_MAX) { free(len); return xo_fail(P, "nnz exceeds 32-bit PetscInt"); } }
      P->ia[row] = (int)tot; P->nnz = tot; }
    free(len);
  }
  P->ja = (int *)malloc(sizeof(int) * P->nnz);
#pragma omp parallel for schedule(static) private(i)
  for (r = 0; r < P->nun + P->npn; ++r) {
    int cu[8 * 27], cp[8 * 8], nu_ = 0, np_ = 0, t, d, b;
    const int isu = r < P->nun; const int nd = isu ? (int)r : (int)(r - P->nun);
    const int *els = isu ? &uel[uoff[nd]] : &pel[poff[nd]];
    const int nels = isu ? uoff[nd + 1] - uoff[nd] : poff[nd + 1] - poff[nd];
    for (t = 0; t < nels; ++t) { for (i = 0; i < nbu; ++i) cu[nu_++] = P->u_map[nbu * els[t] + i]; for (i = 0; i < nbp; ++i) cp[np_++] = P->p_map[nbp * els[t] + i]; }
    qsort(cu, nu_, sizeof(int), cmp_int); qsort(cp, np_, sizeof(int), cmp_int);
    nu_ = uniq_sorted(cu, nu_); np_ = uniq_sorted(cp, np_);
    for (d = 0; d < (isu ? nsd : 1); ++d) {
      int row = isu ? nsd * nd + d : (int)P->nu + nd; int *ja = &P->ja[P->ia[row]]; int c = 0;
      for (t = 0; t < nu_; ++t) for (b = 0; b < nsd; ++b) ja[c++] = nsd * cu[t] + b;
      for (t = 0; t < np_; ++t) ja[c++] = (int)P->nu + cp[t];
    }
  }
  /* Mpscaled pattern: DMCreateMatrix(dmp) (exSaddle.c:315) = 27/9-point box stencil = element coupling of Q1 */
  P->mia = (int *)malloc(sizeof(int) * (P->npn + 1));
  { int64_t tot = 0;
    for (r = 0; r < P->npn; ++r) {
      int i0 = (int)(r % P->PX), j0 = (int)((r / P->PX) % P->PY), k0 = (int)(r / (P->PX * P->PY));
      int cx = 1 + (i0 > 0) + (i0 < P->PX - 1), cy = 1 + (j0 > 0) + (j0 < P->PY - 1), cz = 1 + (k0 > 0) + (k0 < P->PZ - 1);
      P->mia[r] = (int)tot; tot += cx * cy * cz;
    }
    P->mia[P->npn] = (int)tot; P->mnnz = tot; }
  P->mja = (int *)malloc(sizeof(int) * P->mnnz);
#pragma omp parallel for schedule(static)
  for (r = 0; r < P->npn; ++r) {
    int i0 = (int)(r % P->PX), j0 = (int)((r / P->PX) % P->PY), k0 = (int)(r / (P->PX * P->PY)), a, b, c, cnt = 0;
    for (c = -1; c <= 1; ++c) for (b = -1; b <= 1; ++b) for (a = -1; a <= 1; ++a) {
      int ii = i0 + a, jj = j0 + b, kk = k0 + c;
      if (ii < 0 || ii >= P->PX || jj < 0 || jj >= P->PY || kk < 0 || kk >= P->PZ) continue;
      P->mja[P->mia[r] + cnt++] = ii + jj * P->PX + kk * P->PX * P->PY;
    }
  }
  free(uoff); free(poff); free(uel); free(pel);
  return 0;
}

static inline int find_col(const int *ja, int lo, int hi, int col)
{
  while (lo < hi) { int mid = lo + ((hi - lo) >> 1); if (ja[mid] < col) lo = mid + 1; else hi = mid; }   /* lo+hi would overflow int at 64^3 */
  return lo;
}

/* --------------------------------------------------------------- assembly */
/* Element matrices of MatAssemble_Saddle (femixedspace.c:2480-2610).  Zero B entries are skipped; the
   surviving terms are accumulated in the reference's order (q outer, k ascending, (B*D)*B). */
static void element_matrices(const xo_problem *P, int64_t e, double *A11, double *A12, double *A22)
{
  const int nsd = P->prm.nsd, nbu = P->nbu, nbp = P->nbp, nqp = P->nqp, ndu = nsd * nbu;
  double co[27 * 3], cop[8 * 3], GNx[27], GNy[27], GNz[27];
  int i, j, q;
  for (i = 0; i < nbu; ++i) u_coords(P, P->u_map[nbu * e + i], &co[nsd * i]);
  for (i = 0; i < nbp; ++i) p_coords(P, P->p_map[nbp * e + i], &cop[nsd * i]);
  memset(A11, 0, sizeof(double) * ndu * ndu); memset(A12, 0, sizeof(double) * ndu * nbp);
  if (A22) memset(A22, 0, sizeof(double) * nbp * nbp);
  for (q = 0; q < nqp; ++q) {   /* A11: :2491-2561 */
    double detJ = deriv_global(nsd, nbu, P->GNuxi[q], P->GNueta[q], P->GNuzeta[q], GNx, GNy, GNz, co);
    double fac = P->coeff[((size_t)e * nqp + q) * XO_NSLOT + XO_C_ETA] * P->wq[q] * detJ;   /* eta or mu (:2522-2526) */
    double D2 = 2.0 * fac, D1 = 1.0 * fac;
    if (nsd == 2) {
      for (i = 0; i < nbu; ++i) {
        double xi_ = GNx[i], yi = GNy[i];
        for (j = 0; j < nbu; ++j) {
          double xj = GNx[j], yj = GNy[j];
          double *r0 = &A11[(2 * i) * ndu + 2 * j], *r1 = &A11[(2 * i + 1) * ndu + 2 * j];
          r0[0] += xi_ * D2 * xj; r0[0] += yi * D1 * yj;      /* k=0, k=2 */
          r0[1] += yi * D1 * xj;                               /* k=2 */
          r1[0] += xi_ * D1 * yj;                              /* k=2 */
          r1[1] += yi * D2 * yj; r1[1] += xi_ * D1 * xj;      /* k=1, k=2 */
        }
      }
    } else {
      for (i = 0; i < nbu; ++i) {
        double xi_ = GNx[i], yi = GNy[i], zi = GNz[i];
        for (j = 0; j < nbu; ++j) {
          double xj = GNx[j], yj = GNy[j], zj = GNz[j];
          double *r0 = &A11[(3 * i) * ndu + 3 * j], *r1 = r0 + ndu, *r2 = r1 + ndu;
          r0[0] += xi_ * D2 * xj; r0[0] += yi * D1 * yj; r0[0] += zi * D1 * zj;   /* k=0,3,4 */
          r0[1] += yi * D1 * xj;                                                  /* k=3 */
          r0[2] += zi * D1 * xj;                                                  /* k=4 */
          r1[0] += xi_ * D1 * yj;                                                 /* k=3 */
          r1[1] += yi * D2 * yj; r1[1] += xi_ * D1 * xj; r1[1] += zi * D1 * zj;   /* k=1,3,5 */
          r1[2] += zi * D1 * yj;                                                  /* k=5 */
          r2[0] += xi_ * D1 * zj;                                                 /* k=4 */
          r2[1] += yi * D1 * zj;                                                  /* k=5 */
          r2[2] += zi * D2 * zj; r2[2] += xi_ * D1 * xj; r2[2] += yi * D1 * yj;   /* k=2,4,5 */
        }
      }
    }
  }
  for (q = 0; q < nqp; ++q) {   /* A12: :2564-2582 */
    double detJ = deriv_global(nsd, nbu, P->GNuxi[q], P->GNueta[q], P->GNuzeta[q], GNx, GNy, GNz, co);
    double fac = P->wq[q] * detJ;
    for (i = 0; i < nbu; ++i) for (j = 0; j < nbp; ++j) {
      A12[(nsd * i + 0) * nbp + j] -= GNx[i] * P->Np[q][j] * fac;
      A12[(nsd * i + 1) * nbp + j] -= GNy[i] * P->Np[q][j] * fac;
      if (nsd == 3) A12[(nsd * i + 2) * nbp + j] -= GNz[i] * P->Np[q][j] * fac;
    }
  }
  if (A22) for (q = 0; q < nqp; ++q) {   /* A22 (LAME): :2594-2609 */
    double detJ = basis_transformation(nsd, nbp, P->GNpxi[q], P->GNpeta[q], P->GNpzeta[q], cop);
    double fac = P->wq[q] * detJ / P->coeff[((size_t)e * nqp + q) * XO_NSLOT + XO_C_LAM];
    for (i = 0; i < nbp; ++i) for (j = 0; j < nbp; ++j) A22[i * nbp + j] -= P->Np[q][i] * P->Np[q][j] * fac;
  }
}

/* MatAssemble_Schur element matrix (femixedspace.c:2895-2931) */
static void element_schur(const xo_problem *P, int64_t e, double *S)
{
  const int nsd = P->prm.nsd, nbp = P->nbp, nqp = P->nqp; double cop[8 * 3]; int i, j, q;
  for (i = 0; i < nbp; ++i) p_coords(P, P->p_map[nbp * e + i], &cop[nsd * i]);
  memset(S, 0, sizeof(double) * nbp * nbp);
  for (q = 0; q < nqp; ++q) {
    const double *c = &P->coeff[((size_t)e * nqp + q) * XO_NSLOT];
    double cinv = P->prm.lame ? 1.0 / c[XO_C_LAM] + 1.0 / c[XO_C_ETA] : 1.0 / c[XO_C_ETA];   /* :2915-2917 */
    double detJ = basis_transformation(nsd, nbp, P->GNpxi[q], P->GNpeta[q], P->GNpzeta[q], cop);
    double fac = P->wq[q] * detJ;
    for (i = 0; i < nbp; ++i) for (j = 0; j < nbp; ++j) S[i * nbp + j] -= cinv * P->Np[q][i] * P->Np[q][j] * fac;
  }
}

/* MatAssemble_Saddle + MatAssemble_Schur + VecAssemble_F1/F2 (femixedspace.c:2373-2786, 2837-2948).
   ADD_VALUES is done colour by colour (elements of one parity class share no node), which is deterministic
   for any thread count; it reorders the <= 8 element contributions to an entry w.r.t. the reference's
   lexicographic element loop (rounding-level difference only). */
static void assemble(xo_problem *P)
{
  const xo_params *p = &P->prm;
  const int nsd = p->nsd, nbu = P->nbu, nbp = P->nbp, nqp = P->nqp, ndu = nsd * nbu;
  const int mzz = nsd == 3 ? p->mz : 1, ncol = nsd == 3 ? 8 : 4;
  int col;
  P->a = (double *)calloc(P->nnz, sizeof(double));
  P->ma = (double *)calloc(P->mnnz, sizeof(double));
  P->F = (double *)calloc(P->n, sizeof(double));
  for (col = 0; col < ncol; ++col) {
    const int ci = col & 1, cj = (col >> 1) & 1, ck = (col >> 2) & 1;
    const int nei = (p->mx - ci + 1) / 2, nej = (p->my - cj + 1) / 2, nek = (mzz - ck + 1) / 2;
    int64_t t, nt = (int64_t)nei * nej * nek;
#pragma omp parallel
    {
      double *A11 = (double *)malloc(sizeof(double) * ndu * ndu), *A12 = (double *)malloc(sizeof(double) * ndu * nbp);
      double A22[64], S[64];
#pragma omp for schedule(static)
      for (t = 0; t < nt; ++t) {
        int ei = 2 * (int)(t % nei) + ci, ej = 2 * (int)((t / nei) % nej) + cj, ek = 2 * (int)(t / ((int64_t)nei * nej)) + ck;
        int64_t e = ei + (int64_t)ej * p->mx + (int64_t)ek * p->mx * p->my;
        int ug[81], pg[8], i, j, q, d;
        element_matrices(P, e, A11, A12, p->lame ? A22 : NULL);
        element_schur(P, e, S);
        for (i = 0; i < nbu; ++i) for (d = 0; d < nsd; ++d) ug[nsd * i + d] = nsd * P->u_map[nbu * e + i] + d;
        for (i = 0; i < nbp; ++i) pg[i] = (int)P->nu + P->p_map[nbp * e + i];
        for (i = 0; i < ndu; ++i) {   /* uu and up rows (:2615-2616) */
          int lo = P->ia[ug[i]], hi = P->ia[ug[i] + 1];
          for (j = 0; j < ndu; ++j) P->a[find_col(P->ja, lo, hi, ug[j])] += A11[i * ndu + j];
          for (j = 0; j < nbp; ++j) P->a[find_col(P->ja, lo, hi, pg[j])] += A12[i * nbp + j];
        }
        for (i = 0; i < nbp; ++i) {   /* pu (A21 = A12^T, :2584-2590) and pp rows (:2617-2619) */
          int lo = P->ia[pg[i]], hi = P->ia[pg[i] + 1];
          for (j = 0; j < ndu; ++j) P->a[find_col(P->ja, lo, hi, ug[j])] += A12[j * nbp + i];
          if (p->lame) for (j = 0; j < nbp; ++j) P->a[find_col(P->ja, lo, hi, pg[j])] += A22[i * nbp + j];
        }
        for (i = 0; i < nbp; ++i) {   /* Mpscaled (:2937) */
          int r = P->p_map[nbp * e + i], lo = P->mia[r], hi = P->mia[r + 1];
          for (j = 0; j < nbp; ++j) P->ma[find_col(P->mja, lo, hi, P->p_map[nbp * e + j])] += S[i * nbp + j];
        }
        {   /* F1, F2 (:2694-2710, :2762-2778) */
          double co[27 * 3], ef[81], efp[8];
          for (i = 0; i < nbu; ++i) u_coords(P, P->u_map[nbu * e + i], &co[nsd * i]);
          memset(ef, 0, sizeof(ef)); memset(efp, 0, sizeof(efp));
          for (q = 0; q < nqp; ++q) {
            const double *c = &P->coeff[((size_t)e * nqp + q) * XO_NSLOT];
            double Fu[3] = {c[XO_C_FU0], c[XO_C_FU1], c[XO_C_FU2]};
            double detJ = basis_transformation(nsd, nbu, P->GNuxi[q], P->GNueta[q], P->GNuzeta[q], co);
            double fac = P->wq[q] * detJ;
            for (i = 0; i < nbu; ++i) for (d = 0; d < nsd; ++d) ef[nsd * i + d] += P->Nu[q][i] * Fu[d] * fac;
            for (i = 0; i < nbp; ++i) efp[i] += P->Np[q][i] * c[XO_C_FP] * fac;
          }
          for (i = 0; i < ndu; ++i) P->F[ug[i]] += ef[i];
          for (i = 0; i < nbp; ++i) P->F[pg[i]] += efp[i];
        }
      }
      free(A11); free(A12);
    }
  }
}

void xo_csr_mult(int n, const int *ia, const int *ja, const double *a, const double *x, double *y)
{
  int i;
  if (xo_sum_order == 1) {   /* drift experiment only: row sums from the last stored entry to the first */
#pragma omp parallel for schedule(static)
    for (i = 0; i < n; ++i) {
      double s = 0.0; int k;
      for (k = ia[i + 1] - 1; k >= ia[i]; --k) s += a[k] * x[ja[k]];
      y[i] = s;
    }
    return;
  }
#pragma omp parallel for schedule(static)
  for (i = 0; i < n; ++i) {
    double s = 0.0; int k;
    for (k = ia[i]; k < ia[i + 1]; ++k) s += a[k] * x[ja[k]];
    y[i] = s;
  }
}

/* Dirichlet handling: femixedspace.c:2634-2645 and exSaddle.c:278-281 */
static void impose_bc(xo_problem *P, int keep_raw)
{
  int64_t i; int t;
  char *isbc = (char *)calloc(P->n, 1);
  double *g = (double *)calloc(P->n, sizeof(double)), *rd = (double *)calloc(P->n, sizeof(double));
  for (t = 0; t < P->nbc; ++t) { isbc[P->bc_idx[t]] = 1; g[P->bc_idx[t]] = P->bc_val[t]; }   /* ImposeDirichletValuesIS(temp) :2638 */
  xo_csr_mult((int)P->n, P->ia, P->ja, P->a, g, rd);                                          /* :2639, A not yet zeroed */
  for (i = 0; i < P->n; ++i) rd[i] = -1.0 * rd[i];                                            /* :2641 */
  for (t = 0; t < P->nbc; ++t) rd[P->bc_idx[t]] = 0.0;                                        /* :2642 */
  if (keep_raw) { P->a_raw = (double *)malloc(sizeof(double) * P->nnz); memcpy(P->a_raw, P->a, sizeof(double) * P->nnz); }
  /* MatZeroRowsColumns(A, ..., 1.0) keeping the pattern (:2645, :2367) */
#pragma omp parallel for schedule(static)
  for (i = 0; i < P->n; ++i) {
    int k;
    for (k = P->ia[i]; k < P->ia[i + 1]; ++k) {
      int c = P->ja[k];
      if (isbc[i]) P->a[k] = (c == i) ? 1.0 : 0.0;
      else if (isbc[c]) P->a[k] = 0.0;
    }
  }
  for (t = 0; t < P->nbc; ++t) P->F[P->bc_idx[t]] = P->bc_val[t];   /* ImposeDirichletValuesIS(Fu) exSaddle.c:278 */
  for (i = 0; i < P->n; ++i) P->F[i] += 1.0 * rd[i];                /* VecAXPY(F,1.0,rhs_diri) exSaddle.c:281 */
  P->isbc = isbc;
  free(g); free(rd);
}

/* ------------------------------------------------------------------ public */
int xo_create(const xo_params *prm, xo_problem **out)
{
  xo_problem *P = (xo_problem *)calloc(1, sizeof(xo_problem));
  double t0 = xo_wtime();
  *out = P;
  P->prm = *prm;
  if (resolve_params(P)) return 1;
  if (build_mesh(P)) return 1;
  build_tables(P);
  build_bc(P);
  define_qp_properties(P);   /* exSaddle.c:238 */
  q1_projection(P);          /* exSaddle.c:251 */
  if (build_pattern(P)) return 1;   /* DMCreateMatrix + MatAssemble_Saddle_NULL exSaddle.c:267-268 */
  assemble(P);               /* exSaddle.c:269, 276-277, 317 */
  impose_bc(P, P->n <= 200000);
  P->create_seconds = xo_wtime() - t0;
  return 0;
}

/* A coarse level of the monolithic -mg hierarchy (exSaddle.c:215-270): same model / BCs on a coarser mesh, with the
   quadrature-point coefficients interpolated from the given nodal Q1 fields (restricted from the finer level,
   femixedspace.c:2139-2215) instead of evaluated from the model.  nodal: npn x XO_NSLOT, node-major. */
int xo_create_nodal(const xo_params *prm, const double *nodal, xo_problem **out)
{
  xo_problem *P = (xo_problem *)calloc(1, sizeof(xo_problem));
  double t0 = xo_wtime();
  *out = P;
  P->prm = *prm;
  if (resolve_params(P)) return 1;
  if (build_mesh(P)) return 1;
  build_tables(P);
  build_bc(P);
  P->coeff = (double *)calloc((size_t)P->nel * P->nqp * XO_NSLOT, sizeof(double));
  P->coeff_nodal = (double *)malloc((size_t)P->npn * XO_NSLOT * sizeof(double));
  memcpy(P->coeff_nodal, nodal, (size_t)P->npn * XO_NSLOT * sizeof(double));
  interp_nodal_to_qp(P);
  if (build_pattern(P)) return 1;
  assemble(P);
  impose_bc(P, P->n <= 200000);
  P->create_seconds = xo_wtime() - t0;
  return 0;
}
const double *xo_coeff_nodal(const xo_problem *P) { return P->coeff_nodal; }

void xo_destroy(xo_problem *P)
{
  if (!P) return;
  xo_solver_free(P);
  free(P->u_map); free(P->p_map); free(P->coeff); free(P->coeff_nodal); free(P->bc_idx); free(P->bc_val);
  free(P->ia); free(P->ja); free(P->a); free(P->a_raw); free(P->mia); free(P->mja); free(P->ma); free(P->F); free(P->isbc);
  free(P);
}

const char *xo_banner(const xo_problem *P) { return P->banner; }
const char *xo_error(const xo_problem *P) { return P->err; }

void xo_sizes(const xo_problem *P, int64_t out[8])
{
  out[0] = P->n; out[1] = P->nu; out[2] = P->np; out[3] = P->nnz; out[4] = prealloc_total(P);
  out[5] = P->nel; out[6] = P->nbc; out[7] = P->mnnz;
}
const int *xo_A_ia(const xo_problem *P) { return P->ia; }
const int *xo_A_ja(const xo_problem *P) { return P->ja; }
const double *xo_A_a(const xo_problem *P) { return P->a; }
const double *xo_A_raw(const xo_problem *P) { return P->a_raw; }
const int *xo_Mp_ia(const xo_problem *P) { return P->mia; }
const int *xo_Mp_ja(const xo_problem *P) { return P->mja; }
const double *xo_Mp_a(const xo_problem *P) { return P->ma; }
const double *xo_F(const xo_problem *P) { return P->F; }
const int *xo_bc_idx(const xo_problem *P) { return P->bc_idx; }
const double *xo_bc_val(const xo_problem *P) { return P->bc_val; }
const int *xo_u_map(const xo_problem *P) { return P->u_map; }
const int *xo_p_map(const xo_problem *P) { return P->p_map; }
const double *xo_coeff_qp(const xo_problem *P) { return P->coeff; }
int xo_ncoeff(const xo_problem *P) { (void)P; return XO_NSLOT; }
void xo_A_mult(const xo_problem *P, const double *x, double *y) { xo_csr_mult((int)P->n, P->ia, P->ja, P->a, x, y); }

/* MatCreateSubMatrix on the DMComposite ISs (exSaddle.c:319-321): stored zeros are kept */
int64_t xo_submatrix(const xo_problem *P, int rb, int cb, int *ia, int *ja, double *a)
{
  const int r0 = rb ? (int)P->nu : 0, r1 = rb ? (int)P->n : (int)P->nu, c0 = cb ? (int)P->nu : 0, c1 = cb ? (int)P->n : (int)P->nu;
  int64_t cnt = 0; int i, k;
  for (i = r0; i < r1; ++i) {
    if (ia) ia[i - r0] = (int)cnt;
    for (k = P->ia[i]; k < P->ia[i + 1]; ++k) {
      int c = P->ja[k];
      if (c >= c0 && c < c1) { if (ja) { ja[cnt] = c - c0; a[cnt] = P->a[k]; } cnt++; }
    }
  }
  if (ia) ia[r1 - r0] = (int)cnt;
  return cnt;
}

/* SaddleReportSolutionDiagnostics (exSaddle_io.c:7-58) */
void xo_diagnostics(const xo_problem *P, const double *x, double *out)
{
  const int nsd = P->prm.nsd; int d; int64_t i;
  for (d = 0; d < nsd; ++d) {
    double n1 = 0, n2 = 0, ni = 0, mn = DBL_MAX, mx = -DBL_MAX;
    for (i = 0; i < P->nun; ++i) { double v = x[nsd * i + d]; n1 += fabs(v); n2 += v * v; if (fabs(v) > ni) ni = fabs(v); if (v < mn) mn = v; if (v > mx) mx = v; }
    out[0 * nsd + d] = n1; out[1 * nsd + d] = sqrt(n2); out[2 * nsd + d] = ni; out[3 * nsd + d] = mn; out[4 * nsd + d] = mx;
  }
  { double n1 = 0, n2 = 0, ni = 0, mn = DBL_MAX, mx = -DBL_MAX;
    for (i = 0; i < P->np; ++i) { double v = x[P->nu + i]; n1 += fabs(v); n2 += v * v; if (fabs(v) > ni) ni = fabs(v); if (v < mn) mn = v; if (v > mx) mx = v; }
    out[5 * nsd + 0] = n1; out[5 * nsd + 1] = sqrt(n2); out[5 * nsd + 2] = ni; out[5 * nsd + 3] = mn; out[5 * nsd + 4] = mx; }
}
