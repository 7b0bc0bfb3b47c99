// xsb_mfull.cu -- operator-free fine level (-xsb_matrix_free full): what the solver still needs from the velocity
// block when neither A nor A00 is stored (128^3: nnz(A) = 1.13e10 does not fit 32-bit PetscInt).
//   * 1 / diag(A00) for the fine-level Jacobi smoother, accumulated from the element matrices with the expressions and
//     the colour order of the assembly kernel (femixedspace.c:2544-2559), so it equals the assembled diagonal;
//   * the first Galerkin operator P^T A00 P (MatPtAP of PCMG, abf.opts:12), element by element: a Q2 element is one
//     cell of the coarse node lattice, so P restricted to it maps its 8 corners to its 27 nodes and
//        A_c = sum_e P_e^T M_e K_e M_e P_e + P^T (I - M) P          (M masks the Dirichlet dofs)
//     with P_e^T M K_e M P_e = sum_q eta w |J| Bt_q^T D Bt_q, Bt_q the gradients of the MASKED interpolated corner
//     functions at the fine Gauss points -- 24 x 24 per element instead of the 81 x 81 element matrix.
// Both are 8-colour, atomic-free, fixed-order sums (bit-reproducible).
#include "xsb.h"
#include <cub/cub.cuh>

static inline unsigned nblk(int64_t n, int bs = 256) { return (unsigned)((n + bs - 1) / bs); }

// ------------------------------------------------------------------ diag(A00)
// one CTA per element of the colour, one thread per (node, component)
__global__ void __launch_bounds__(96) mf_diag_kernel(Lattice L, int colour, const FeTables *T, const double *__restrict__ eta, double *diag)
{
  __shared__ double sfac[27];
  const int ci = colour & 1, cj = (colour >> 1) & 1, ck = (colour >> 2) & 1;
  const int nei = (L.mx - ci + 1) / 2, nej = (L.my - cj + 1) / 2;
  const int64_t t = blockIdx.x;
  const int ei = 2 * (int)(t % nei) + ci, ej = 2 * (int)((t / nei) % nej) + cj, ek = 2 * (int)(t / ((int64_t)nei * nej)) + ck;
  const int64_t e = ei + (int64_t)ej * L.mx + (int64_t)ek * L.mx * L.my;
  if (threadIdx.x < 27) sfac[threadIdx.x] = eta[e * 27 + threadIdx.x] * T->wq[threadIdx.x] * T->detJ;
  __syncthreads();
  if (threadIdx.x >= 81) return;
  const int i = threadIdx.x / 3, comp = threadIdx.x - 3 * i;
  double r = 0.0;
  for (int q = 0; q < 27; ++q) {
    const double D1 = 1.0 * sfac[q], D2 = 2.0 * sfac[q];
    const double xi = T->Gu[q][i][0], yi = T->Gu[q][i][1], zi = T->Gu[q][i][2];
    if (comp == 0) { r += xi * D2 * xi; r += yi * D1 * yi; r += zi * D1 * zi; }
    else if (comp == 1) { r += yi * D2 * yi; r += xi * D1 * xi; r += zi * D1 * zi; }
    else { r += zi * D2 * zi; r += xi * D1 * xi; r += yi * D1 * yi; }
  }
  const int64_t node = (2 * ei + i % 3) + (int64_t)(2 * ej + (i / 3) % 3) * L.NX + (int64_t)(2 * ek + i / 9) * L.NX * L.NY;
  diag[3 * node + comp] += r;
}
__global__ void mf_diag_finish_kernel(int64_t n, const unsigned char *__restrict__ isbc, double *d)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; if (i >= n) return;
  const double v = isbc[i] ? 1.0 : d[i];   // MatZeroRowsColumns(diag = 1.0)
  d[i] = v == 0.0 ? 1.0 : 1.0 / v;           // PCSetUp_Jacobi
}
int mf_diag_inv(xsb_ctx c, double *idiag)
{
  const Lattice &L = c->lat; cudaStream_t st = c->stream;
  CUDA_OK(cudaMemsetAsync(idiag, 0, sizeof(double) * L.nu, st));
  for (int col = 0; col < 8; ++col) {
    const int ci = col & 1, cj = (col >> 1) & 1, ck = (col >> 2) & 1;
    const int64_t ne = (int64_t)((L.mx - ci + 1) / 2) * ((L.my - cj + 1) / 2) * ((L.mz - ck + 1) / 2);
    if (ne <= 0) continue;
    mf_diag_kernel<<<(unsigned)ne, 96, 0, st>>>(L, col, (const FeTables *)c->fe_tables, c->coeff, idiag); KERNEL_OK();
  }
  mf_diag_finish_kernel<<<nblk(L.nu), 256, 0, st>>>(L.nu, c->isbc, idiag); KERNEL_OK();
  return 0;
}

// ------------------------------------------------------------------ element-wise Galerkin product
__global__ void k_box_len27(BoxPattern p, int64_t *len)
{
  int64_t nd = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; if (nd >= (int64_t)p.nx * p.ny * p.nz) return;
  len[nd] = box_size(p, (int)(nd % p.nx), (int)((nd / p.nx) % p.ny), (int)(nd / ((int64_t)p.nx * p.ny)));
}
__global__ void k_narrow32(int64_t n, const int64_t *a, int *b) { int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; if (i <= n) b[i] = (int)a[i]; }
__global__ void k_box_ja27(BoxPattern p, const int *__restrict__ ia, int *__restrict__ ja)
{
  int64_t nd = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; if (nd >= (int64_t)p.nx * p.ny * p.nz) return;
  const int i = (int)(nd % p.nx), j = (int)((nd / p.nx) % p.ny), k = (int)(nd / ((int64_t)p.nx * p.ny));
  int l0, h0, l1, h1, l2, h2; box_range(p, i, p.nx, l0, h0); box_range(p, j, p.ny, l1, h1); box_range(p, k, p.nz, l2, h2);
  int cpos = ia[nd];
  for (int kk = l2; kk <= h2; ++kk) for (int jj = l1; jj <= h1; ++jj) for (int ii = l0; ii <= h0; ++ii) ja[cpos++] = ii + jj * p.nx + kk * p.nx * p.ny;
}

// 1-D interpolation weight of fine node a (0..2 inside the element) from corner A (0..1)
__device__ __forceinline__ double w1(int a, int A) { return a == 1 ? 0.5 : ((a >> 1) == A ? 1.0 : 0.0); }

// P^T (I - M) P: one thread per constrained dof; every contribution is a product of powers of two, the sums are exact,
// so the atomic order cannot change the result.  Runs on the zeroed matrix, before the element contributions.
__global__ void k_galerkin_bc(Lattice L, BoxPattern cp, int nbc, const int *__restrict__ bc_idx, const int *__restrict__ cia, double *ca)
{
  const int t = blockIdx.x * blockDim.x + threadIdx.x; if (t >= nbc) return;
  const int dof = bc_idx[t]; if (dof >= L.nu) return;
  const int64_t nd = dof / 3; const int comp = dof - 3 * (int)nd;
  const int f[3] = {(int)(nd % L.NX), (int)((nd / L.NX) % L.NY), (int)(nd / ((int64_t)L.NX * L.NY))};
  int n[3], c0[3]; double w[3];
  for (int d = 0; d < 3; ++d) { if (f[d] & 1) { n[d] = 2; c0[d] = (f[d] - 1) >> 1; w[d] = 0.5; } else { n[d] = 1; c0[d] = f[d] >> 1; w[d] = 1.0; } }
  const double wI = w[0] * w[1] * w[2];
  for (int Ik = 0; Ik < n[2]; ++Ik) for (int Ij = 0; Ij < n[1]; ++Ij) for (int Ii = 0; Ii < n[0]; ++Ii)
    for (int Jk = 0; Jk < n[2]; ++Jk) for (int Jj = 0; Jj < n[1]; ++Jj) for (int Ji = 0; Ji < n[0]; ++Ji) {
      const int I0 = c0[0] + Ii, I1 = c0[1] + Ij, I2 = c0[2] + Ik, J0 = c0[0] + Ji, J1 = c0[1] + Jj, J2 = c0[2] + Jk;
      const int64_t In = I0 + (int64_t)I1 * cp.nx + (int64_t)I2 * cp.nx * cp.ny;
      const int slot = box_slot(cp, I0, I1, I2, J0, J1, J2);
      atomicAdd(ca + ((int64_t)cia[In] + slot) * 9 + 4 * comp, wI * wI);
    }
}

// one CTA (64 threads) per element of the colour; thread = (corner I, corner J) pair
__global__ void __launch_bounds__(64) k_galerkin_elem(Lattice L, int colour, const FeTables *T, const double *__restrict__ eta,
                                                      const unsigned char *__restrict__ isbc, BoxPattern cp, const int *__restrict__ cia, double *ca)
{
  __shared__ double sG[27][8][3][3];   // [q][corner][component of the masked interpolant][derivative]
  __shared__ double sfac[27];
  __shared__ unsigned char sm[27][3];
  const int ci = colour & 1, cj = (colour >> 1) & 1, ck = (colour >> 2) & 1;
  const int nei = (L.mx - ci + 1) / 2, nej = (L.my - cj + 1) / 2;
  const int64_t t = blockIdx.x;
  const int ei = 2 * (int)(t % nei) + ci, ej = 2 * (int)((t / nei) % nej) + cj, ek = 2 * (int)(t / ((int64_t)nei * nej)) + ck;
  const int64_t e = ei + (int64_t)ej * L.mx + (int64_t)ek * L.mx * L.my;
  const int tid = threadIdx.x;
  if (tid < 27) {
    sfac[tid] = eta[e * 27 + tid] * T->wq[tid] * T->detJ;
    const int64_t node = (2 * ei + tid % 3) + (int64_t)(2 * ej + (tid / 3) % 3) * L.NX + (int64_t)(2 * ek + tid / 9) * L.NX * L.NY;
    for (int c = 0; c < 3; ++c) sm[tid][c] = isbc[3 * node + c];
  }
  __syncthreads();
  for (int x = tid; x < 27 * 8 * 9; x += 64) {
    const int q = x / 72, r = x - 72 * q, I = r / 9, cd = r - 9 * I, c = cd / 3, d = cd - 3 * c;
    const int A = I & 1, B = (I >> 1) & 1, Cc = I >> 2;
    double g = 0.0;
    for (int i = 0; i < 27; ++i) {
      const double w = w1(i % 3, A) * w1((i / 3) % 3, B) * w1(i / 9, Cc);
      if (w != 0.0 && !sm[i][c]) g += w * T->Gu[q][i][d];
    }
    sG[q][I][c][d] = g;
  }
  __syncthreads();
  const int I = tid >> 3, J = tid & 7;
  double blk[3][3];
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int cc = 0; cc < 3; ++cc) blk[c][cc] = 0.0;
  for (int q = 0; q < 27; ++q) {
    const double f = sfac[q];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      // row (I, c) tests with e_c * phi_I^(c); column (J, cc) is the trial e_cc * phi_J^(cc):
      //   eta (grad u + grad u^T) : grad v = delta_{c cc} grad phi_I . grad phi_J + d_cc phi_I d_c phi_J
      double dot = 0.0;
#pragma unroll
      for (int d = 0; d < 3; ++d) dot += sG[q][I][c][d] * sG[q][J][c][d];
      blk[c][c] += f * dot;
#pragma unroll
      for (int cc = 0; cc < 3; ++cc) blk[c][cc] += f * (sG[q][I][c][cc] * sG[q][J][cc][c]);
    }
  }
  const int I0 = ei + (I & 1), I1 = ej + ((I >> 1) & 1), I2 = ek + (I >> 2), J0 = ei + (J & 1), J1 = ej + ((J >> 1) & 1), J2 = ek + (J >> 2);
  const int64_t In = I0 + (int64_t)I1 * cp.nx + (int64_t)I2 * cp.nx * cp.ny;
  double *dst = ca + ((int64_t)cia[In] + box_slot(cp, I0, I1, I2, J0, J1, J2)) * 9;
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int cc = 0; cc < 3; ++cc) dst[3 * c + cc] += blk[c][cc];
}

// C.A = P^T A00 P on the local lattice (C.nx/ny/nz already set to the coarse lattice of the local fine lattice)
int galerkin_elements(xsb_ctx c, Level &C)
{
  const Lattice &L = c->lat; cudaStream_t st = c->stream;
  if (c->nsd != 3) return xsb_fail(c, XSB_ERR_SUP, "element-wise Galerkin product is 3-D only");
  if (C.nx != L.mx + 1 || C.ny != L.my + 1 || C.nz != L.mz + 1) return xsb_fail(c, XSB_ERR_ARG, "element-wise Galerkin: coarse lattice %dx%dx%d does not match the mesh", C.nx, C.ny, C.nz);
  BoxPattern cp{C.nx, C.ny, C.nz, 0};
  const int64_t ncn = (int64_t)C.nx * C.ny * C.nz;
  int64_t *len = nullptr; CUDA_OK(cudaMalloc(&len, sizeof(int64_t) * (ncn + 1)));
  k_box_len27<<<nblk(ncn), 256, 0, st>>>(cp, len); KERNEL_OK();
  CUDA_OK(cudaMemsetAsync(len + ncn, 0, sizeof(int64_t), st));
  void *tmp = nullptr; size_t tb = 0;
  CUDA_OK(cub::DeviceScan::ExclusiveSum(nullptr, tb, len, len, ncn + 1, st));
  CUDA_OK(cudaMalloc(&tmp, tb));
  CUDA_OK(cub::DeviceScan::ExclusiveSum(tmp, tb, len, len, ncn + 1, st));
  int64_t tot = 0; CUDA_OK(cudaMemcpyAsync(&tot, len + ncn, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  CUDA_OK(cudaStreamSynchronize(st)); CUDA_OK(cudaFree(tmp));
  Baij &A = C.A; A.nb = (int)ncn; A.bs = 3; A.nblk = tot; A.pat = cp;
  XSB_CHK(dev_alloc(c, &A.ia, (size_t)ncn + 1)); XSB_CHK(dev_alloc(c, &A.ja, (size_t)tot)); XSB_CHK(dev_alloc(c, &A.a, (size_t)tot * 9 + 2));   // dev_alloc zero-fills
  k_narrow32<<<nblk(ncn + 1), 256, 0, st>>>(ncn, len, A.ia); KERNEL_OK();
  k_box_ja27<<<nblk(ncn), 256, 0, st>>>(cp, A.ia, A.ja); KERNEL_OK();
  if (c->nbc > 0) { k_galerkin_bc<<<nblk(c->nbc), 256, 0, st>>>(L, cp, c->nbc, c->bc_idx, A.ia, A.a); KERNEL_OK(); }
  for (int col = 0; col < 8; ++col) {
    const int ci = col & 1, cj = (col >> 1) & 1, ck = (col >> 2) & 1;
    const int64_t ne = (int64_t)((L.mx - ci + 1) / 2) * ((L.my - cj + 1) / 2) * ((L.mz - ck + 1) / 2);
    if (ne <= 0) continue;
    k_galerkin_elem<<<(unsigned)ne, 64, 0, st>>>(L, col, (const FeTables *)c->fe_tables, c->coeff, c->isbc, cp, A.ia, A.a); KERNEL_OK();
  }
  CUDA_OK(cudaStreamSynchronize(st)); CUDA_OK(cudaFree(len));
  C.owns_A = true;
  return 0;
}
