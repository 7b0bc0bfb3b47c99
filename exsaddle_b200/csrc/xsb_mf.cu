// xsb_mf.cu -- matrix-free, sum-factorised Q2 apply of the velocity block A00 (K4 of SURVEY 2.1): set-up, dispatch and the
// 8-colour kernel of round 1.  The production kernel is the one-pass TMA-staged kernel in xsb_mf1p.cu (-xsb_mf_kernel 4,
// default); the colour kernel below (-xsb_mf_kernel 3) is kept as the reference point of the A/B timings and as the path for
// vectors that are not 16-byte aligned (the bulk copies of the one-pass kernel need that).
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

void mf_tab_scaled(const Lattice &L, MfTabS &TS)
{
  MfTab T0; host_mf_tab(T0);
  for (int q = 0; q < 3; ++q) { TS.w[q] = T0.w[q]; for (int n = 0; n < 3; ++n) { TS.N[q][n] = T0.N[q][n]; TS.Dx[q][n] = T0.D[q][n] / L.hu[0]; TS.Dy[q][n] = T0.D[q][n] / L.hu[1]; TS.Dz[q][n] = T0.D[q][n] / L.hu[2]; } }
}

#define FULL 0xffffffffu

// ---------------------------------------------------------------------------------------------------------
// Version 2: 3 lanes per element (10 elements per warp).  Lane `a` owns the x-index a of the element: the
// (j,k) plane of nodes with i = a on gather / scatter and the 9 Gauss points (a, b, q).  Only the contraction
// along x crosses lanes (2 rotating shuffles per value: an all-gather going up, a reduce-scatter coming
// down); the y and z contractions are register-local.  The symmetric gradient E = G + G^T is accumulated
// directly (6 x 9 accumulators per lane instead of 9 x 9), k-slab by k-slab.  Per element: ~650 shuffles
// (v1: ~3900) for ~4.4 kFMA-class instructions.
// ncu on v1 / the first v2 showed both latency bound (long-scoreboard 11 of 15 cycles per issue, FP64 pipe 13 %):
// loads were issued one (c,k,j) at a time behind the constraint test.  Here every load of the element is issued
// up front and unconditionally (27 x values, 9 per-node constraint masks, 9 viscosities), constraints are applied
// with selects, and the scatter either preloads its 27 y values before the Gauss-point work (SCATTER 1) or uses
// fire-and-forget reductions (SCATTER 0; one add per address per launch, so still deterministic).
__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(FULL, v, src); }
// (c,d) -> slot of the symmetric 3x3: xx yy zz xy xz yz
__device__ __forceinline__ constexpr int sym_idx(int c, int d) { return c == d ? c : (c + d == 1 ? 3 : (c + d == 2 ? 4 : 5)); }

template <int SCATTER>
__global__ void __launch_bounds__(128, 2) mf_a00_kernel_v2(Lattice L, int colour, int kz0, int kz1, int reverse, MfTabS T, double detJ,
                                                           const double *__restrict__ eta, const unsigned char *__restrict__ bcnode,
                                                           const double *__restrict__ x, double *__restrict__ y)
{
  const int lane = threadIdx.x & 31;
  const int g = lane / 3, a = lane - 3 * g, base = 3 * g;
  const bool active = lane < 30;
  const int a1 = a == 2 ? 0 : a + 1, a2 = a == 0 ? 2 : a - 1;     // (a+1)%3, (a+2)%3
  const int src1 = base + a1, src2 = base + a2;
  const int ci = colour & 1, cj = (colour >> 1) & 1, ck = (colour >> 2) & 1;
  // elements of this colour inside the chunk of element layers [kz0, kz1): ek = kf, kf + 2, ...
  const int kf = kz0 + (((kz0 & 1) != ck) ? 1 : 0);
  const int nei = (L.mx - ci + 1) / 2, nej = (L.my - cj + 1) / 2, nek = kz1 > kf ? (kz1 - kf + 1) / 2 : 0;
  const int64_t nelc = (int64_t)nei * nej * nek;
  const int64_t ngroups = (nelc + 9) / 10;
  const int64_t jstride = L.NX, kstride = (int64_t)L.NX * L.NY;
  // lane-specific coefficients, rotated so that index r pairs with lane (a+r)%3:
  //   gather  (Gauss index a, node n = (a+r)%3) and scatter (destination node i = (a+r)%3, my Gauss index a) use the same numbers
  double Nr[3], Dr[3];
  Nr[0] = T.N[a][a]; Nr[1] = T.N[a][a1]; Nr[2] = T.N[a][a2];
  Dr[0] = T.Dx[a][a]; Dr[1] = T.Dx[a][a1]; Dr[2] = T.Dx[a][a2];
  const double wa = T.w[a] * detJ;
  // one group of 10 elements per warp (a grid-stride loop measured 10 % slower: the hardware CTA scheduler balances
  // better); odd launches sweep top-down so the planes the previous launch touched last are still in L2
  const int64_t wg = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (wg >= ngroups) return;
  const int64_t t = (reverse ? ngroups - 1 - wg : wg) * 10 + g;
  const bool valid = active && t < nelc;
  int ei = ci, ej = cj, ek = kf;
  if (valid) { ei = 2 * (int)(t % nei) + ci; ej = 2 * (int)((t / nei) % nej) + cj; ek = 2 * (int)(t / ((int64_t)nei * nej)) + kf; }
  const int64_t e = ei + (int64_t)ej * L.mx + (int64_t)ek * L.mx * L.my;
  const int64_t node0 = (2 * ei + a) + (int64_t)(2 * ej) * jstride + (int64_t)(2 * ek) * kstride;   // idle lanes read a valid element
  // ---- every load of the element, issued before any use
  double U[3][9], fac[9]; unsigned bc[9];
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int64_t nd = node0 + j * jstride + k * kstride;
      bc[3 * k + j] = bcnode[nd];
#pragma unroll
      for (int c = 0; c < 3; ++c) U[c][3 * k + j] = __ldg(x + 3 * nd + c);
    }
#pragma unroll
  for (int b = 0; b < 3; ++b)
#pragma unroll
    for (int q = 0; q < 3; ++q) fac[3 * b + q] = __ldcs(eta + e * 27 + a + 3 * b + 9 * q);   // read once per product: evict-first, keep x / y in L2
  unsigned bcmask = 0;   // bit (9c + 3k + j): dof constrained (or lane idle)
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int n = 0; n < 9; ++n) bcmask |= (valid ? ((bc[n] >> c) & 1u) : 1u) << (9 * c + n);
  double E[6][3][3];     // [sym slot][b][q]
#pragma unroll
  for (int s = 0; s < 6; ++s)
#pragma unroll
    for (int b = 0; b < 3; ++b)
#pragma unroll
      for (int q = 0; q < 3; ++q) E[s][b][q] = 0.0;
  // ---- forward: E_cd = d u_c / d x_d + d u_d / d x_c at my 9 Gauss points
#pragma unroll
  for (int c = 0; c < 3; ++c) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      double tN[3], tD[3];
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const double u = ((bcmask >> (9 * c + 3 * k + j)) & 1u) ? 0.0 : U[c][3 * k + j];
        const double u1 = shfl_d(u, src1), u2 = shfl_d(u, src2);
        tN[j] = Nr[0] * u + Nr[1] * u1 + Nr[2] * u2;
        tD[j] = Dr[0] * u + Dr[1] * u1 + Dr[2] * u2;
      }
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        const double gx = T.N[b][0] * tD[0] + T.N[b][1] * tD[1] + T.N[b][2] * tD[2];      // D in x, N in y
        const double gy = T.Dy[b][0] * tN[0] + T.Dy[b][1] * tN[1] + T.Dy[b][2] * tN[2];   // N in x, D in y
        const double gz = T.N[b][0] * tN[0] + T.N[b][1] * tN[1] + T.N[b][2] * tN[2];      // N in x, N in y (D in z below)
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          E[sym_idx(c, 0)][b][q] += T.N[q][k] * gx;
          E[sym_idx(c, 1)][b][q] += T.N[q][k] * gy;
          E[sym_idx(c, 2)][b][q] += T.Dz[q][k] * gz;
        }
      }
    }
  }
  // ---- scatter targets: issue the 27 loads now, consume them after the Gauss-point work and the transposed sums
  double Yold[SCATTER ? 27 : 1];
  if (SCATTER) {
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int n = 0; n < 9; ++n) Yold[9 * c + n] = y[3 * (node0 + (n % 3) * jstride + (n / 3) * kstride) + c];
  }
  // ---- Gauss points: sigma = eta w |J| (G + G^T); diagonal slots hold G_cc once, so double them
#pragma unroll
  for (int b = 0; b < 3; ++b)
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const double f = fac[3 * b + q] * (wa * (T.w[b] * T.w[q]));
      E[0][b][q] = f * (E[0][b][q] + E[0][b][q]); E[1][b][q] = f * (E[1][b][q] + E[1][b][q]); E[2][b][q] = f * (E[2][b][q] + E[2][b][q]);
      E[3][b][q] *= f; E[4][b][q] *= f; E[5][b][q] *= f;
    }
  // ---- transpose: y_c(i,j,k) += sum over Gauss points of sigma_cd d N_(i,j,k) / d x_d
#pragma unroll
  for (int c = 0; c < 3; ++c) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      double rx[3], ry[3], rz[3];   // index b
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        rx[b] = T.N[0][k] * E[sym_idx(c, 0)][b][0] + T.N[1][k] * E[sym_idx(c, 0)][b][1] + T.N[2][k] * E[sym_idx(c, 0)][b][2];
        ry[b] = T.N[0][k] * E[sym_idx(c, 1)][b][0] + T.N[1][k] * E[sym_idx(c, 1)][b][1] + T.N[2][k] * E[sym_idx(c, 1)][b][2];
        rz[b] = T.Dz[0][k] * E[sym_idx(c, 2)][b][0] + T.Dz[1][k] * E[sym_idx(c, 2)][b][1] + T.Dz[2][k] * E[sym_idx(c, 2)][b][2];
      }
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const double qD = T.N[0][j] * rx[0] + T.N[1][j] * rx[1] + T.N[2][j] * rx[2];                     // pairs with D in x
        const double qN = T.Dy[0][j] * ry[0] + T.Dy[1][j] * ry[1] + T.Dy[2][j] * ry[2]
                        + T.N[0][j] * rz[0] + T.N[1][j] * rz[1] + T.N[2][j] * rz[2];                     // pairs with N in x
        // reduce-scatter over the 3 lanes: my contribution to node i = (a+r)%3 is s_r;
        // lane a receives s1 of lane (a+2)%3 and s2 of lane (a+1)%3
        const double s0 = Nr[0] * qN + Dr[0] * qD, s1 = Nr[1] * qN + Dr[1] * qD, s2 = Nr[2] * qN + Dr[2] * qD;
        const double Y = s0 + shfl_d(s1, src2) + shfl_d(s2, src1);
        if (!((bcmask >> (9 * c + 3 * k + j)) & 1u)) {
          const int64_t i0 = 3 * (node0 + j * jstride + k * kstride) + c;
          if (SCATTER) y[i0] = Yold[9 * c + 3 * k + j] + Y; else atomicAdd(y + i0, Y);
        }
      }
    }
  }
}

// per-node constraint mask (bit c = component c is a Dirichlet dof): one byte load per node in the element kernel
__global__ void mf_bcnode_kernel(int64_t nun, const unsigned char *__restrict__ isbc, unsigned char *__restrict__ bcnode)
{
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; if (n >= nun) return;
  bcnode[n] = (unsigned char)((isbc[3 * n] ? 1 : 0) | (isbc[3 * n + 1] ? 2 : 0) | (isbc[3 * n + 2] ? 4 : 0));
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
__global__ void mf_epilogue_kernel(int64_t n, const unsigned char *__restrict__ isbc, const double *__restrict__ x, double *kx,
                                   double *__restrict__ out, Epilogue ep)
{
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double v = kx[i];
    if (isbc && isbc[i]) v = x[i];
    out[i] = mf_epi(ep, i, v);
  }
}

// element layers a product applies: all of the local lattice, or (slabs, constrained operator) only those touching an owned plane
static void mf_layer_range(xsb_ctx c, bool owned_only, int *zlo, int *zhi)
{
  const Lattice &L = c->lat;
  *zlo = 0; *zhi = L.mz;
  if (c->slab.nranks > 1 && owned_only) {   // [k0-1, k1) of the local [k0-2, k1+1): the two extra ghost layers exist for the assembled Galerkin rows
    *zlo = c->slab.k0 - c->slab.e0 - 1; if (*zlo < 0) *zlo = 0;
    *zhi = c->slab.k1 - c->slab.e0; if (*zhi > L.mz) *zhi = L.mz;
  }
}

int mf_setup(xsb_ctx c)
{
  if (c->nsd != 3) return xsb_fail(c, XSB_ERR_SUP, "-xsb_matrix_free is implemented for the 3-D Q2 velocity block");
  if (c->mf_opts_read && c->mf_tmp && c->mf_bcnode) return 0;   // once per option change, not per product
  const int keep_phase = c->phase; c->phase = 2;
  struct PhaseGuard { xsb_ctx c; int p; ~PhaseGuard() { c->phase = p; } } guard{c, keep_phase};
  MfTab T; host_mf_tab(T);
  CUDA_OK(cudaMemcpyToSymbolAsync(c_tab, &T, sizeof(T), 0, cudaMemcpyHostToDevice, c->stream));
  if (!c->mf_tmp) XSB_CHK(dev_alloc(c, &c->mf_tmp, (size_t)c->lat.nu));
  c->so.mf_kernel = c->opt.integer("xsb_mf_kernel", 4);   // 4: one-pass TMA-staged kernel (xsb_mf1p.cu); 3: 8 colour passes, 3 lanes per element, reduction scatter
  c->so.mf_tile = c->opt.integer("xsb_mf_tile", 0);
  c->so.mf_chunk = c->opt.integer("xsb_mf_chunk", 0);     // element layers per z-chunk (0 = no chunking)
  c->so.mf_reverse = c->opt.integer("xsb_mf_reverse", 1); // alternate the sweep direction of successive colour launches
  if (c->so.mf_kernel != 3 && c->so.mf_kernel != 4) return xsb_fail(c, XSB_ERR_ARG, "-xsb_mf_kernel must be 4 (one-pass kernel) or 3 (8-colour kernel)");
  if (c->so.mf_tile != 0 && c->so.mf_tile != 1) return xsb_fail(c, XSB_ERR_ARG, "-xsb_mf_tile must be 0 (16 x 5 elements, one CTA per SM) or 1 (8 x 5, two CTAs per SM)");
  if (!c->mf_bcnode) {
    XSB_CHK(dev_alloc(c, &c->mf_bcnode, (size_t)c->lat.nun));
    mf_bcnode_kernel<<<(unsigned)((c->lat.nun + 255) / 256), 256, 0, c->stream>>>(c->lat.nun, c->isbc, c->mf_bcnode); KERNEL_OK();
  }
  c->mf_opts_read = true;
  if (c->so.mf_kernel == 4) { int zlo, zhi; mf_layer_range(c, true, &zlo, &zhi); XSB_CHK(mf1p_prepare(c, zlo, zhi)); }
  return 0;
}

static int mf_apply_core(xsb_ctx c, const double *x, double *y, const Epilogue &ep, const unsigned char *isbc, const unsigned char *bcnode);
// y = epilogue(A00 x), matrix-free.  x and y must not alias.
int mf_a00_apply(xsb_ctx c, const double *x, double *y, const Epilogue &ep) { return mf_apply_core(c, x, y, ep, c->isbc, c->mf_bcnode); }
// y = K x with no Dirichlet rows / columns (the operator MatAssemble_Saddle holds before MatZeroRowsColumns, used for rhs_diri)
int mf_a00_apply_raw(xsb_ctx c, const double *x, double *y)
{
  XSB_CHK(mf_setup(c));
  if (!c->mf_bczero) { const int kp = c->phase; c->phase = 2; int rc = dev_alloc(c, &c->mf_bczero, (size_t)c->lat.nun); c->phase = kp; if (rc) return rc; }   // dev_alloc zero-fills
  Epilogue ep;
  return mf_apply_core(c, x, y, ep, nullptr, c->mf_bczero);
}
static int mf_apply_core(xsb_ctx c, const double *x, double *y, const Epilogue &ep, const unsigned char *isbc, const unsigned char *bcnode)
{
  const Lattice &L = c->lat; cudaStream_t st = c->stream;
  const double detJ = L.hu[0] * L.hu[1] * L.hu[2];
  const double *eta = c->coeff;   // slot C_ETA (eta, or mu for LAME)
  // Slabs: only the element layers that touch an owned node plane are needed, [k0-1, k1) of the local [k0-2, k1+1); the two
  // extra ghost layers exist for the assembled Galerkin rows.
  int zlo, zhi; mf_layer_range(c, isbc != nullptr, &zlo, &zhi);
  if (c->so.mf_kernel == 4) {
    const uintptr_t al = (uintptr_t)x | (uintptr_t)ep.b | (uintptr_t)ep.idiag | (uintptr_t)ep.pkm1;
    if (!(al & 15)) return mf1p_apply(c, x, y, ep, bcnode, zlo, zhi);
    // vectors off the 16-byte grid (a caller's sub-vector): the colour kernel below has no alignment requirement
  }
  CUDA_OK(cudaMemsetAsync(c->mf_tmp, 0, sizeof(double) * L.nu, st));
  MfTabS TS; mf_tab_scaled(L, TS);
  {
    // optional z-chunks of element layers (-xsb_mf_chunk): the 8 colours run chunk by chunk so a chunk's x / y planes stay in L2
    int chunk = c->so.mf_chunk;
    if (chunk <= 0) chunk = L.mz;
    if (chunk >= L.mz) chunk = L.mz; else chunk &= ~1;   // even chunk starts keep the colour parity pattern regular
    int launch = 0;
    for (int kz0 = zlo; kz0 < zhi; kz0 += chunk) {
      const int kz1 = kz0 + chunk < zhi ? kz0 + chunk : zhi;
      for (int col = 0; col < 8; ++col) {
        const int ci = col & 1, cj = (col >> 1) & 1, ck = (col >> 2) & 1;
        const int kf = kz0 + (((kz0 & 1) != ck) ? 1 : 0);
        const int64_t ne = (int64_t)((L.mx - ci + 1) / 2) * ((L.my - cj + 1) / 2) * (kz1 > kf ? (kz1 - kf + 1) / 2 : 0);
        if (ne <= 0) continue;
        const int64_t warps = (ne + 9) / 10; int64_t blocks = (warps + 3) / 4;
        const int rev = c->so.mf_reverse ? (launch++) & 1 : 0;
        mf_a00_kernel_v2<0><<<(unsigned)blocks, 128, 0, st>>>(L, col, kz0, kz1, rev, TS, detJ, eta, bcnode, x, c->mf_tmp);
        KERNEL_OK();
      }
    }
  }
  int64_t nb = (L.nu + 255) / 256; if (nb > 148 * 16) nb = 148 * 16;
  mf_epilogue_kernel<<<(unsigned)nb, 256, 0, st>>>(L.nu, isbc, x, c->mf_tmp, y, ep);
  KERNEL_OK();
  return 0;
}
