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

// velocity node i on a line of m elements: the pressure nodes it couples to and the 1-D coefficients (summed over shared elements)
__device__ __forceinline__ int line_u(int i, int m, const double (*M)[2], const double (*G)[2], int *P, double *cm, double *cg)
{
  if (i & 1) { const int e = (i - 1) >> 1; P[0] = e; P[1] = e + 1; cm[0] = M[1][0]; cm[1] = M[1][1]; cg[0] = G[1][0]; cg[1] = G[1][1]; return 2; }
  const int e1 = i >> 1, e0 = e1 - 1;   // node 2 of element e0, node 0 of element e1
  int n = 0;
  if (e0 >= 0) { P[n] = e0; cm[n] = M[2][0]; cg[n] = G[2][0]; ++n; }
  P[n] = e1; cm[n] = (e0 >= 0 ? M[2][1] : 0.0) + (e1 < m ? M[0][0] : 0.0); cg[n] = (e0 >= 0 ? G[2][1] : 0.0) + (e1 < m ? G[0][0] : 0.0); ++n;
  if (e1 < m) { P[n] = e1 + 1; cm[n] = M[0][1]; cg[n] = G[0][1]; ++n; }
  return n;
}
// coefficient between velocity node i and pressure node P along one line
__device__ __forceinline__ void pair_up(int i, int P, int m, const double (*M)[2], const double (*G)[2], double &cm, double &cg)
{
  cm = 0.0; cg = 0.0;
  for (int e = P - 1; e <= P; ++e) { if (e < 0 || e >= m) continue; const int l = i - 2 * e; if (l < 0 || l > 2) continue; cm += M[l][P - e]; cg += G[l][P - e]; }
}

template <int NSD>
__global__ void __launch_bounds__(256) k_grad(Lattice L, GradTab T, const unsigned char *__restrict__ isbc, const double *__restrict__ xp, double *__restrict__ y, const double *__restrict__ yadd, int64_t node0, int64_t nnodes)
{
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; if (t >= nnodes) return;
  const int64_t nd = node0 + t;
  const int i = (int)(nd % L.NX), j = (int)((nd / L.NX) % L.NY), k = NSD == 3 ? (int)(nd / ((int64_t)L.NX * L.NY)) : 0;
  int Pi[3], Pj[3], Pk[3]; double mi[3], gi[3], mj[3], gj[3], mk[3], gk[3];
  const int ni = line_u(i, L.mx, T.M[0], T.G[0], Pi, mi, gi), nj = line_u(j, L.my, T.M[1], T.G[1], Pj, mj, gj);
  int nk = 1; Pk[0] = 0; mk[0] = 1.0; gk[0] = 0.0;
  if (NSD == 3) nk = line_u(k, L.mz, T.M[2], T.G[2], Pk, mk, gk);
  double a0 = 0.0, a1 = 0.0, a2 = 0.0;
  for (int c = 0; c < nk; ++c) for (int b = 0; b < nj; ++b) {
    const double *row = xp + (int64_t)Pk[c] * L.PX * L.PY + (int64_t)Pj[b] * L.PX;
    double s0 = 0.0, s1 = 0.0;   // sum_i g_i x , sum_i m_i x
    for (int a = 0; a < ni; ++a) { const double v = __ldg(row + Pi[a]); s0 += gi[a] * v; s1 += mi[a] * v; }
    a0 += s0 * mj[b] * mk[c]; a1 += s1 * gj[b] * mk[c];
    if (NSD == 3) a2 += s1 * mj[b] * gk[c];
  }
  const double acc[3] = {a0, a1, a2};
#pragma unroll
  for (int c = 0; c < NSD; ++c) {
    const int64_t dof = nd * NSD + c;
    const double base = yadd ? yadd[dof] : 0.0;
    y[dof] = isbc[dof] ? base : base - acc[c];   // constrained rows of A01 are zero (MatZeroRowsColumns)
  }
}

template <int NSD>
__global__ void __launch_bounds__(256) k_div(Lattice L, GradTab T, const unsigned char *__restrict__ isbc, const double *__restrict__ xu, double *__restrict__ y, const double *__restrict__ yadd, int64_t p0, int64_t np)
{
  // the 1-D tables are indexed by lane-dependent node positions: from shared memory (a kernel parameter lives in the constant bank, where
  // divergent indices serialise)
  __shared__ double sM[3][3][2], sG[3][3][2];
  if (threadIdx.x < 18) { (&sM[0][0][0])[threadIdx.x] = (&T.M[0][0][0])[threadIdx.x]; (&sG[0][0][0])[threadIdx.x] = (&T.G[0][0][0])[threadIdx.x]; }
  __syncthreads();
  const int64_t wq = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; const int lane = threadIdx.x & 31;
  if (wq >= np) return;
  const int64_t pn = p0 + wq;
  const int P0 = (int)(pn % L.PX), P1 = (int)((pn / L.PX) % L.PY), P2 = NSD == 3 ? (int)(pn / ((int64_t)L.PX * L.PY)) : 0;
  double acc = 0.0;
  const int nl = NSD == 3 ? 25 : 5;
  if (lane < nl) {
    const int dj = lane % 5, dk = lane / 5;
    const int j = 2 * P1 - 2 + dj, k = NSD == 3 ? 2 * P2 - 2 + dk : 0;
    if (j >= 0 && j < L.NY && k >= 0 && k < L.NZ) {
      double mj, gj, mk = 1.0, gk = 0.0;
      pair_up(j, P1, L.my, sM[1], sG[1], mj, gj);
      if (NSD == 3) pair_up(k, P2, L.mz, sM[2], sG[2], mk, gk);
      const int64_t line = ((int64_t)k * L.NY + j) * L.NX;
      for (int di = 0; di < 5; ++di) {
        const int i = 2 * P0 - 2 + di; if (i < 0 || i >= L.NX) continue;
        double mi, gi; pair_up(i, P0, L.mx, sM[0], sG[0], mi, gi);
        const int64_t dof = (line + i) * NSD;
        const double c0 = gi * mj * mk, c1 = mi * gj * mk, c2 = mi * mj * gk;
        if (!isbc[dof]) acc += c0 * __ldg(xu + dof);                 // constrained columns of A10 are zero
        if (!isbc[dof + 1]) acc += c1 * __ldg(xu + dof + 1);
        if (NSD == 3) { if (!isbc[dof + 2]) acc += c2 * __ldg(xu + dof + 2); }
      }
    }
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) y[pn] = (yadd ? yadd[pn] : 0.0) - acc;
}

// y[rows of the velocity nodes node0 .. node0+nnodes) = A01 xp (+ yadd);  y, yadd indexed like the velocity vector
int grad_apply(xsb_ctx c, const double *xp, double *y, int64_t dof0, int64_t ndofs, const double *yadd)
{
  const Lattice &L = c->lat; GradTab T; grad_tables(L, T);
  const int64_t node0 = dof0 / L.nsd, nnodes = ndofs / L.nsd;
  if (nnodes <= 0) return 0;
  const unsigned nb = (unsigned)((nnodes + 255) / 256);
  if (L.nsd == 3) k_grad<3><<<nb, 256, 0, c->stream>>>(L, T, c->isbc, xp, y, yadd, node0, nnodes);
  else k_grad<2><<<nb, 256, 0, c->stream>>>(L, T, c->isbc, xp, y, yadd, node0, nnodes);
  KERNEL_OK(); return 0;
}
// y[pressure rows p0 .. p0+np) = A10 xu (+ yadd);  y, yadd indexed like the pressure vector
int div_apply(xsb_ctx c, const double *xu, double *y, int64_t p0, int64_t np, const double *yadd)
{
  const Lattice &L = c->lat; GradTab T; grad_tables(L, T);
  if (np <= 0) return 0;
  const unsigned nb = (unsigned)((np * 32 + 255) / 256);
  if (L.nsd == 3) k_div<3><<<nb, 256, 0, c->stream>>>(L, T, c->isbc, xu, y, yadd, p0, np);
  else k_div<2><<<nb, 256, 0, c->stream>>>(L, T, c->isbc, xu, y, yadd, p0, np);
  KERNEL_OK(); return 0;
}
