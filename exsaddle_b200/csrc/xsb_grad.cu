// xsb_grad.cu -- the gradient / divergence blocks applied matrix-free.
//
// Replaces MatMult on A01 = A[u,p] and A10 = A[p,u] (assembled by MatAssemble_Saddle, femixedspace.c:2576-2579: element entry
// -sum_q dN^u_i/dx_c N^p_j w_q detJ, then MatZeroRowsColumns on the constrained velocity dofs) where the solve applies them on
// their own: inside every full-operator product of the operator-free mode and in the fieldsplit's x_u - A01 y_p.
// On the uniform box mesh the entry does not depend on the coefficient and separates over the directions:
//     A01[(n_u, c), n_p] = - prod_d T_d ,   T_d = G_d[i_d][P_d] (d == c)  or  M_d[i_d][P_d] (d != c)
//     M_d[l][m] = sum_q w_q b_l(xi_q) psi_m(xi_q) h_d ,   G_d[l][m] = sum_q w_q b'_l(xi_q) psi_m(xi_q)     (1-D Q2 x Q1 element tables)
// summed over the (one or two) elements that contain both nodes -- equal to the assembled blocks to 1e-15 (tests/test_gpu_parity.py)
// for Stokes and Lame, 2-D / 3-D, non-unit box sizes.  So the products stream x and y once and no matrix: 0.1 GB instead of
// 2.4 GB of CSR traffic per outer iteration at 64^3, and A01 / A10 need not be read at all during the solve.
//   k_grad: one thread per velocity node, <= 27 pressure values, all NSD components at once;
//   k_div : one warp per pressure node; lanes take the (j, k) lines of the 5 x 5 (x 5) velocity box, 5 nodes x NSD contiguous values each.
// The 1-D coefficients per node (which pressure nodes, which sums of element-table entries: boundary nodes have one element, corner
// nodes two) are tabulated per direction on the host and uploaded once (grad_prepare): the kernels only load and multiply.
#include "xsb.h"

struct GradTab { double M[3][3][2], G[3][3][2]; };   // [direction][local velocity node][local pressure node]

void grad_tables(const Lattice &L, GradTab &T)
{
  static const double xi[3] = {-0.774596669241483, 0.0, 0.774596669241483};   // femixedspace.c:1379-1380
  static const double w[3] = {0.555555555555556, 0.888888888888889, 0.555555555555556};
  memset(&T, 0, sizeof(T));
  for (int d = 0; d < 3; ++d) for (int q = 0; q < 3; ++q) {
    const double x = xi[q];
    const double b[3] = {0.5 * x * (x - 1.0), (1.0 + x) * (1.0 - x), 0.5 * (1.0 + x) * x}, g[3] = {0.5 * (2.0 * x - 1.0), -2.0 * x, 0.5 * (2.0 * x + 1.0)};
    const double psi[2] = {0.5 * (1.0 - x), 0.5 * (1.0 + x)};
    for (int l = 0; l < 3; ++l) for (int m = 0; m < 2; ++m) { T.M[d][l][m] += w[q] * b[l] * psi[m] * L.hu[d]; T.G[d][l][m] += w[q] * g[l] * psi[m]; }
  }
}

// Per-direction coefficient tables on the device (built once per lattice by grad_prepare):
//   velocity node i  -> up to 3 pressure nodes:  uP[d][i][3] (-1 = none), uM[d][i][3], uG[d][i][3]
//   pressure node P  -> its 5 velocity nodes 2P-2 .. 2P+2:  pM[d][P][5], pG[d][P][5] (0 where the node does not exist / does not couple)
struct GradDev { const int *uP[3]; const double *uM[3], *uG[3], *pM[3], *pG[3]; };

template <int NSD>
__global__ void __launch_bounds__(256) k_grad(int NX, int NY, int PX, int PY, GradDev T, const unsigned char *__restrict__ isbc, const double *__restrict__ xp, double *__restrict__ y,
                                              const double *__restrict__ yadd, int plane0)
{
  // grid: x over the nodes of a plane, y over the planes [plane0, plane0 + gridDim.y)
  const int t = blockIdx.x * blockDim.x + threadIdx.x; if (t >= NX * NY) return;
  const int k = NSD == 3 ? plane0 + (int)blockIdx.y : 0, j = t / NX, i = t - j * NX;
  const int64_t nd = (int64_t)k * NX * NY + t;
  int Pi[3], Pj[3], Pk[3]; double mi[3], gi[3], mj[3], gj[3], mk[3], gk[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    Pi[a] = T.uP[0][3 * i + a]; mi[a] = T.uM[0][3 * i + a]; gi[a] = T.uG[0][3 * i + a];
    Pj[a] = T.uP[1][3 * j + a]; mj[a] = T.uM[1][3 * j + a]; gj[a] = T.uG[1][3 * j + a];
    if (NSD == 3) { Pk[a] = T.uP[2][3 * k + a]; mk[a] = T.uM[2][3 * k + a]; gk[a] = T.uG[2][3 * k + a]; }
    else { Pk[a] = a == 0 ? 0 : -1; mk[a] = a == 0 ? 1.0 : 0.0; gk[a] = 0.0; }
  }
  double a0 = 0.0, a1 = 0.0, a2 = 0.0;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    if (Pk[c] < 0) continue;
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      if (Pj[b] < 0) continue;
      const double *row = xp + ((int64_t)Pk[c] * PY + Pj[b]) * PX;
      double s0 = 0.0, s1 = 0.0;   // sum_i g_i x , sum_i m_i x
#pragma unroll
      for (int a = 0; a < 3; ++a) if (Pi[a] >= 0) { const double v = __ldg(row + Pi[a]); s0 += gi[a] * v; s1 += mi[a] * v; }
      a0 += s0 * mj[b] * mk[c]; a1 += s1 * gj[b] * mk[c];
      if (NSD == 3) a2 += s1 * mj[b] * gk[c];
    }
  }
  const double acc[3] = {a0, a1, a2};
#pragma unroll
  for (int c = 0; c < NSD; ++c) {
    const int64_t dof = nd * NSD + c;
    const double base = yadd ? yadd[dof] : 0.0;
    y[dof] = isbc[dof] ? base : base - acc[c];   // constrained rows of A01 are zero (MatZeroRowsColumns)
  }
}

// One warp per pressure node; lanes take the (j, k) lines of its 5 x 5 (x 5) velocity box, 5 nodes x NSD contiguous values each.
// (A variant that spread each line over 5 NSD adjacent lanes -- fewer cache lines per load instruction -- measured the same: the
// kernel is bound by the latency of its dependent table / vector loads, ~0.3 ms per launch at 64^3.)
template <int NSD>
__global__ void __launch_bounds__(256) k_div(int NX, int NY, int NZ, int PX, int PY, GradDev T, const unsigned char *__restrict__ isbc, const double *__restrict__ xu, double *__restrict__ y,
                                             const double *__restrict__ yadd, int64_t p0, int64_t np)
{
  const int64_t wq = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; const int lane = threadIdx.x & 31;
  if (wq >= np) return;
  const int64_t pn = p0 + wq;
  const int P0 = (int)(pn % PX), P1 = (int)((pn / PX) % PY), P2 = NSD == 3 ? (int)(pn / ((int64_t)PX * PY)) : 0;
  double acc = 0.0;
  const int nl = NSD == 3 ? 25 : 5;
  if (lane < nl) {
    const int dj = lane % 5, dk = lane / 5;
    const int j = 2 * P1 - 2 + dj, k = NSD == 3 ? 2 * P2 - 2 + dk : 0;
    const double mj = T.pM[1][5 * P1 + dj], gj = T.pG[1][5 * P1 + dj];
    const double mk = NSD == 3 ? T.pM[2][5 * P2 + dk] : 1.0, gk = NSD == 3 ? T.pG[2][5 * P2 + dk] : 0.0;
    if (j >= 0 && j < NY && k >= 0 && k < NZ) {
      const int64_t line = ((int64_t)k * NY + j) * NX;
#pragma unroll
      for (int di = 0; di < 5; ++di) {
        const int i = 2 * P0 - 2 + di; if (i < 0 || i >= NX) continue;
        const double mi = T.pM[0][5 * P0 + di], gi = T.pG[0][5 * P0 + di];
        const int64_t dof = (line + i) * NSD;
        if (!isbc[dof]) acc += gi * mj * mk * __ldg(xu + dof);                 // constrained columns of A10 are zero
        if (!isbc[dof + 1]) acc += mi * gj * mk * __ldg(xu + dof + 1);
        if (NSD == 3) { if (!isbc[dof + 2]) acc += mi * mj * gk * __ldg(xu + dof + 2); }
      }
    }
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) y[pn] = (yadd ? yadd[pn] : 0.0) - acc;
}

// host: the coefficient tables of one lattice, uploaded once (phase-0 allocation: they live as long as the assembled problem)
static void line_tables(int m, const double (*M)[2], const double (*G)[2], std::vector<int> &uP, std::vector<double> &uM, std::vector<double> &uG, std::vector<double> &pM, std::vector<double> &pG)
{
  const int N = 2 * m + 1, P = m + 1;
  uP.assign(3 * N, -1); uM.assign(3 * N, 0.0); uG.assign(3 * N, 0.0); pM.assign(5 * P, 0.0); pG.assign(5 * P, 0.0);
  for (int i = 0; i < N; ++i) {
    int n = 0;
    auto add = [&](int Pn, double cm, double cg) { for (int a = 0; a < n; ++a) if (uP[3 * i + a] == Pn) { uM[3 * i + a] += cm; uG[3 * i + a] += cg; return; } uP[3 * i + n] = Pn; uM[3 * i + n] = cm; uG[3 * i + n] = cg; ++n; };
    for (int e = (i - 2) / 2 < 0 ? 0 : (i - 2) / 2; e <= i / 2 && e < m; ++e) {   // elements containing node i, ascending
      const int l = i - 2 * e; if (l < 0 || l > 2) continue;
      for (int lp = 0; lp < 2; ++lp) add(e + lp, M[l][lp], G[l][lp]);
    }
  }
  for (int Pn = 0; Pn < P; ++Pn) for (int di = 0; di < 5; ++di) {
    const int i = 2 * Pn - 2 + di; if (i < 0 || i >= N) continue;
    for (int a = 0; a < 3; ++a) if (uP[3 * i + a] == Pn) { pM[5 * Pn + di] = uM[3 * i + a]; pG[5 * Pn + di] = uG[3 * i + a]; }
  }
}
// host-only view of the tables of one direction (m elements, node spacing h), for the CPU tests: uP/uM/uG hold 3 entries per
// velocity node, pM/pG 5 entries per pressure node
extern "C" int xsb_grad_line_tables(int m, double h, int32_t *uP, double *uM, double *uG, double *pM, double *pG)
{
  if (m < 1 || !(h > 0.0) || !uP || !uM || !uG || !pM || !pG) return XSB_ERR_ARG;
  Lattice L{}; L.hu[0] = L.hu[1] = L.hu[2] = h;
  GradTab T; grad_tables(L, T);
  std::vector<int> a; std::vector<double> b, c, d, e;
  line_tables(m, T.M[0], T.G[0], a, b, c, d, e);
  for (size_t i = 0; i < a.size(); ++i) { uP[i] = a[i]; uM[i] = b[i]; uG[i] = c[i]; }
  for (size_t i = 0; i < d.size(); ++i) { pM[i] = d[i]; pG[i] = e[i]; }
  return XSB_OK;
}

int grad_prepare(xsb_ctx c)
{
  const Lattice &L = c->lat;
  if (c->grad_tab && c->grad_key[0] == L.NX && c->grad_key[1] == L.NY && c->grad_key[2] == L.NZ) return 0;
  GradTab T; grad_tables(L, T);
  const int m[3] = {L.mx, L.my, L.nsd == 3 ? L.mz : 0};
  std::vector<double> blob; std::vector<int> iblob; size_t off_i[3], off_d[3][4];
  for (int d = 0; d < 3; ++d) {
    std::vector<int> uP; std::vector<double> uM, uG, pM, pG;
    if (m[d] > 0) line_tables(m[d], T.M[d], T.G[d], uP, uM, uG, pM, pG);
    off_i[d] = iblob.size(); iblob.insert(iblob.end(), uP.begin(), uP.end());
    const std::vector<double> *v[4] = {&uM, &uG, &pM, &pG};
    for (int q = 0; q < 4; ++q) { off_d[d][q] = blob.size(); blob.insert(blob.end(), v[q]->begin(), v[q]->end()); }
  }
  const int save = c->phase; c->phase = 0;
  double *dd = nullptr; int *di = nullptr;
  int rc = dev_alloc(c, &dd, blob.size() + 1); if (!rc) rc = dev_alloc(c, &di, iblob.size() + 1);
  c->phase = save; if (rc) return rc;
  CUDA_OK(cudaMemcpyAsync(dd, blob.data(), sizeof(double) * blob.size(), cudaMemcpyHostToDevice, c->stream));
  CUDA_OK(cudaMemcpyAsync(di, iblob.data(), sizeof(int) * iblob.size(), cudaMemcpyHostToDevice, c->stream));
  CUDA_OK(cudaStreamSynchronize(c->stream));
  GradDev *G = new GradDev();
  for (int d = 0; d < 3; ++d) { G->uP[d] = di + off_i[d]; G->uM[d] = dd + off_d[d][0]; G->uG[d] = dd + off_d[d][1]; G->pM[d] = dd + off_d[d][2]; G->pG[d] = dd + off_d[d][3]; }
  delete (GradDev *)c->grad_tab; c->grad_tab = G;
  c->grad_key[0] = L.NX; c->grad_key[1] = L.NY; c->grad_key[2] = L.NZ;
  return 0;
}
void grad_free(xsb_ctx c) { delete (GradDev *)c->grad_tab; c->grad_tab = nullptr; }

// y[rows of the velocity nodes of the planes containing dof0 .. dof0+ndofs) = A01 xp (+ yadd);  y, yadd indexed like the velocity vector
int grad_apply(xsb_ctx c, const double *xp, double *y, int64_t dof0, int64_t ndofs, const double *yadd)
{
  const Lattice &L = c->lat;
  if (ndofs <= 0) return 0;
  XSB_CHK(grad_prepare(c));
  const GradDev &T = *(const GradDev *)c->grad_tab;
  const int64_t pn = (int64_t)L.NX * L.NY * L.nsd;   // dofs per node plane: the requested range is a whole number of planes (owned planes of a slab, or everything)
  if (dof0 % pn || ndofs % pn) return xsb_fail(c, XSB_ERR_ARG, "gradient product: row range is not a whole number of node planes");
  const int plane0 = (int)(dof0 / pn), nplanes = (int)(ndofs / pn);
  dim3 grid((unsigned)((L.NX * L.NY + 255) / 256), (unsigned)nplanes);
  if (L.nsd == 3) k_grad<3><<<grid, 256, 0, c->stream>>>(L.NX, L.NY, L.PX, L.PY, T, c->isbc, xp, y, yadd, plane0);
  else k_grad<2><<<grid, 256, 0, c->stream>>>(L.NX, L.NY, L.PX, L.PY, T, c->isbc, xp, y, yadd, 0);
  KERNEL_OK(); return 0;
}
// y[pressure rows p0 .. p0+np) = A10 xu (+ yadd);  y, yadd indexed like the pressure vector
int div_apply(xsb_ctx c, const double *xu, double *y, int64_t p0, int64_t np, const double *yadd)
{
  const Lattice &L = c->lat;
  if (np <= 0) return 0;
  XSB_CHK(grad_prepare(c));
  const GradDev &T = *(const GradDev *)c->grad_tab;
  const unsigned nb = (unsigned)((np * 32 + 255) / 256);
  if (L.nsd == 3) k_div<3><<<nb, 256, 0, c->stream>>>(L.NX, L.NY, L.NZ, L.PX, L.PY, T, c->isbc, xu, y, yadd, p0, np);
  else k_div<2><<<nb, 256, 0, c->stream>>>(L.NX, L.NY, 1, L.PX, L.PY, T, c->isbc, xu, y, yadd, p0, np);
  KERNEL_OK(); return 0;
}
