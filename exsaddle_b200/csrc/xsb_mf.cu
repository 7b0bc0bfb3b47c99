// xsb_mf.cu -- matrix-free, sum-factorised Q2 apply of the velocity block A00 (K4 of SURVEY 2.1).
//
// y = A00 x without reading the 10.4 GB assembled block: per element, gather the 27 x 3 nodal values, evaluate
// grad u at the 27 Gauss points by sum factorisation (three 1-D contractions instead of a 27 x 27 one),
// sigma = eta w |J| (grad u + grad u^T)  (= D B u with D = diag(2,2,2,1,1,1), femixedspace.c:2510-2559),
// apply the transposed contractions and add into y.  HBM traffic is 16 B per dof + 8 B per Gauss point
// (0.16 GB at 64^3 against 10.4 GB), so the kernel is bound by FP64 issue (~5.9 kFMA per element), not by HBM.
//
// Mapping: 9 lanes per element (one lane per (i,j) node column / (a,b) Gauss column), 3 elements per warp.
// Contractions along k stay in registers; contractions along i and j exchange values between the 3 lanes of a
// row / column of the element's 3 x 3 lane tile with warp shuffles.  No shared memory, no atomics: elements are
// processed in 8 parity colours (same-colour elements share no node), colours run in a fixed order, so the
// result is bit-reproducible.  Dirichlet handling follows MatZeroRowsColumns(diag = 1): constrained inputs
// are masked on gather, constrained outputs are skipped on scatter and the epilogue writes y_bc = x_bc.
#include "xsb.h"

struct MfTab { double N[3][3], D[3][3], w[3]; };   // N[q][n], D[q][n]: 1-D Q2 basis / derivative at Gauss point q
__constant__ MfTab c_tab;

static void host_mf_tab(MfTab &T)
{
  static const double xi1d[3] = {-0.774596669241483, 0.0, 0.774596669241483};   // femixedspace.c:1379-1380
  static const double wt1d[3] = {0.555555555555556, 0.888888888888889, 0.555555555555556};
  for (int q = 0; q < 3; ++q) {
    const double x = xi1d[q];
    T.N[q][0] = 0.5 * x * (x - 1.0); T.N[q][1] = (1.0 + x) * (1.0 - x); T.N[q][2] = 0.5 * (1.0 + x) * x;   // :1540-1542
    T.D[q][0] = 0.5 * (2.0 * x - 1.0); T.D[q][1] = -2.0 * x; T.D[q][2] = 0.5 * (2.0 * x + 1.0);             // :1837-1839
    T.w[q] = wt1d[q];
  }
}

#define FULL 0xffffffffu
// values of `v` held by the 3 lanes of my row (same b, a = 0,1,2) / my column (same a, b = 0,1,2)
#define ROW3(v, o) { o[0] = __shfl_sync(FULL, v, rowb); o[1] = __shfl_sync(FULL, v, rowb + 1); o[2] = __shfl_sync(FULL, v, rowb + 2); }
#define COL3(v, o) { o[0] = __shfl_sync(FULL, v, colb); o[1] = __shfl_sync(FULL, v, colb + 3); o[2] = __shfl_sync(FULL, v, colb + 6); }
#define DOT3(c, v) ((c)[0] * (v)[0] + (c)[1] * (v)[1] + (c)[2] * (v)[2])

__global__ void __launch_bounds__(128) mf_a00_kernel(Lattice L, int colour, double ihx, double ihy, double ihz, double detJ,
                                                     const double *__restrict__ eta, const unsigned char *__restrict__ isbc,
                                                     const double *__restrict__ x, double *__restrict__ y)
{
  const int lane = threadIdx.x & 31;
  const int g = lane / 9, r = lane - 9 * g, a = r % 3, b = r / 3;
  const bool active = lane < 27;
  const int rowb = 9 * g + 3 * b, colb = 9 * g + a;
  const int ci = colour & 1, cj = (colour >> 1) & 1, ck = (colour >> 2) & 1;
  const int nei = (L.mx - ci + 1) / 2, nej = (L.my - cj + 1) / 2, nek = (L.mz - ck + 1) / 2;
  const int64_t nelc = (int64_t)nei * nej * nek;
  const int64_t wg = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t t = wg * 3 + g;
  const bool valid = active && t < nelc;
  // lane-specific 1-D coefficients (forward: rows of N/D at my Gauss index; transpose: columns at my node index)
  double Na[3], Da[3], Nb[3], Db[3], NaT[3], DaT[3], NbT[3], DbT[3];
#pragma unroll
  for (int n = 0; n < 3; ++n) {
    Na[n] = c_tab.N[a][n]; Da[n] = c_tab.D[a][n]; Nb[n] = c_tab.N[b][n]; Db[n] = c_tab.D[b][n];
    NaT[n] = c_tab.N[n][a]; DaT[n] = c_tab.D[n][a]; NbT[n] = c_tab.N[n][b]; DbT[n] = c_tab.D[n][b];
  }
  int ei = 0, ej = 0, ek = 0;
  if (valid) { ei = 2 * (int)(t % nei) + ci; ej = 2 * (int)((t / nei) % nej) + cj; ek = 2 * (int)(t / ((int64_t)nei * nej)) + ck; }
  const int64_t e = ei + (int64_t)ej * L.mx + (int64_t)ek * L.mx * L.my;
  const int64_t node0 = (2 * ei + a) + (int64_t)(2 * ej + b) * L.NX + (int64_t)(2 * ek) * L.NX * L.NY, kstride = (int64_t)L.NX * L.NY;
  // gather (Dirichlet columns masked: MatZeroRowsColumns)
  double U[3][3];   // [comp][k]
  double fac[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int64_t i0 = 3 * (node0 + k * kstride);
#pragma unroll
    for (int c = 0; c < 3; ++c) U[c][k] = (valid && !isbc[i0 + c]) ? __ldg(x + i0 + c) : 0.0;
    fac[k] = valid ? __ldg(eta + e * 27 + a + 3 * b + 9 * k) * (c_tab.w[a] * c_tab.w[b] * c_tab.w[k]) * detJ : 0.0;   // eta w |J| at (a,b,k)
  }
  // forward: G[c][d][q] = d u_c / d x_d at Gauss points (a,b,q)
  double G[3][3][3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    double tN[3], tD[3], tNN[3], tND[3], tDN[3], v[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) { ROW3(U[c][k], v); tN[k] = DOT3(Na, v); tD[k] = DOT3(Da, v); }
#pragma unroll
    for (int k = 0; k < 3; ++k) { COL3(tN[k], v); tNN[k] = DOT3(Nb, v); tND[k] = DOT3(Db, v); COL3(tD[k], v); tDN[k] = DOT3(Nb, v); }
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      G[c][0][q] = ihx * (c_tab.N[q][0] * tDN[0] + c_tab.N[q][1] * tDN[1] + c_tab.N[q][2] * tDN[2]);
      G[c][1][q] = ihy * (c_tab.N[q][0] * tND[0] + c_tab.N[q][1] * tND[1] + c_tab.N[q][2] * tND[2]);
      G[c][2][q] = ihz * (c_tab.D[q][0] * tNN[0] + c_tab.D[q][1] * tNN[1] + c_tab.D[q][2] * tNN[2]);
    }
  }
  // Gauss-point work: sigma_cd = eta w |J| (G_cd + G_dc), pre-scaled by 1/h_d for the transposed derivative
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    const double f = fac[q];
    const double sxx = f * (G[0][0][q] + G[0][0][q]), syy = f * (G[1][1][q] + G[1][1][q]), szz = f * (G[2][2][q] + G[2][2][q]);
    const double sxy = f * (G[0][1][q] + G[1][0][q]), sxz = f * (G[0][2][q] + G[2][0][q]), syz = f * (G[1][2][q] + G[2][1][q]);
    G[0][0][q] = ihx * sxx; G[0][1][q] = ihy * sxy; G[0][2][q] = ihz * sxz;
    G[1][0][q] = ihx * sxy; G[1][1][q] = ihy * syy; G[1][2][q] = ihz * syz;
    G[2][0][q] = ihx * sxz; G[2][1][q] = ihy * syz; G[2][2][q] = ihz * szz;
  }
  // transpose: Y[c][k] at node (a,b,k)
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    double rDN[3], rND[3], rNN[3], qA[3], qB[3], v[3], Y[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      rDN[k] = c_tab.N[0][k] * G[c][0][0] + c_tab.N[1][k] * G[c][0][1] + c_tab.N[2][k] * G[c][0][2];
      rND[k] = c_tab.N[0][k] * G[c][1][0] + c_tab.N[1][k] * G[c][1][1] + c_tab.N[2][k] * G[c][1][2];
      rNN[k] = c_tab.D[0][k] * G[c][2][0] + c_tab.D[1][k] * G[c][2][1] + c_tab.D[2][k] * G[c][2][2];
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      COL3(rDN[k], v); qB[k] = DOT3(NbT, v);
      COL3(rND[k], v); qA[k] = DOT3(DbT, v);
      COL3(rNN[k], v); qA[k] += DOT3(NbT, v);
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) { ROW3(qA[k], v); Y[k] = DOT3(NaT, v); ROW3(qB[k], v); Y[k] += DOT3(DaT, v); }
    // scatter: same-colour elements share no node, so a plain read-modify-write is race free
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int64_t i0 = 3 * (node0 + k * kstride) + c;
      if (valid && !isbc[i0]) y[i0] += Y[k];
    }
  }
}

__device__ __forceinline__ double mf_epi(const Epilogue &ep, int64_t i, double ax)
{
  switch (ep.mode) {
  case EPI_RESIDUAL:   return ep.b[i] - ax;
  case EPI_CHEB_FIRST: return ep.pk[i] + ep.s0 * (ep.idiag[i] * (ep.b[i] - ax));
  case EPI_CHEB:       return ep.s0 * ep.pkm1[i] + ep.s1 * ep.pk[i] + ep.s2 * (ep.idiag[i] * (ep.b[i] - ax));
  default:             return ax;
  }
}
// out = epilogue( isbc ? x : (K x) ): identity rows of the constrained dofs + the fused smoother update
__global__ void mf_epilogue_kernel(int64_t n, const unsigned char *__restrict__ isbc, const double *__restrict__ x, const double *__restrict__ kx,
                                   double *__restrict__ out, Epilogue ep)
{
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = mf_epi(ep, i, isbc[i] ? x[i] : kx[i]);
}

int mf_setup(xsb_ctx c)
{
  if (c->nsd != 3) return xsb_fail(c, XSB_ERR_SUP, "-xsb_matrix_free is implemented for the 3-D Q2 velocity block");
  MfTab T; host_mf_tab(T);
  CUDA_OK(cudaMemcpyToSymbolAsync(c_tab, &T, sizeof(T), 0, cudaMemcpyHostToDevice, c->stream));
  if (!c->mf_tmp) XSB_CHK(dev_alloc(c, &c->mf_tmp, (size_t)c->lat.nu));
  return 0;
}

// y = epilogue(A00 x), matrix-free.  x and y must not alias.
int mf_a00_apply(xsb_ctx c, const double *x, double *y, const Epilogue &ep)
{
  const Lattice &L = c->lat; cudaStream_t st = c->stream;
  const double detJ = L.hu[0] * L.hu[1] * L.hu[2];
  const double *eta = c->coeff;   // slot C_ETA (eta, or mu for LAME)
  CUDA_OK(cudaMemsetAsync(c->mf_tmp, 0, sizeof(double) * L.nu, st));
  for (int col = 0; col < 8; ++col) {
    const int ci = col & 1, cj = (col >> 1) & 1, ck = (col >> 2) & 1;
    const int64_t ne = (int64_t)((L.mx - ci + 1) / 2) * ((L.my - cj + 1) / 2) * ((L.mz - ck + 1) / 2);
    if (ne <= 0) continue;
    const int64_t warps = (ne + 2) / 3, blocks = (warps * 32 + 127) / 128;
    mf_a00_kernel<<<(unsigned)blocks, 128, 0, st>>>(L, col, 1.0 / L.hu[0], 1.0 / L.hu[1], 1.0 / L.hu[2], detJ, eta, c->isbc, x, c->mf_tmp); KERNEL_OK();
  }
  int64_t nb = (L.nu + 255) / 256; if (nb > 148 * 16) nb = 148 * 16;
  mf_epilogue_kernel<<<(unsigned)nb, 256, 0, st>>>(L.nu, c->isbc, x, c->mf_tmp, y, ep); KERNEL_OK();
  return 0;
}
