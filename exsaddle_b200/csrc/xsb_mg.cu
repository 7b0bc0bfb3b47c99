// xsb_mg.cu -- geometric multigrid on the velocity block (K8, K9, K11 of SURVEY 2.1).
//
// Replaces what PETSc's PCMG does for -saddle_fieldsplit_u_pc_type mg with a DMDA (abf.opts:4-13,
// exSaddle.c:408-422): DMCoarsen ((n-1)/2+1 nodes per direction), DMCreateInterpolation (Q1 on the node
// lattice, MAIJ over the NSD components), Galerkin P^T A P (MatPtAP), Chebyshev/Jacobi smoothers with the
// GMRES eigenvalue estimate, LU on the coarsest level.
//   * transfers are stencil kernels on the lattice (no stored P);
//   * every level operator is a BAIJ(NSD) "box pattern" matrix, so block positions are closed-form;
//   * the Galerkin product is a gather: one thread per coarse block, no atomics, fixed summation order;
//   * the coarsest operator is inverted densely on the device (Gauss-Jordan, SPD => no pivoting) and applied
//     as a GEMV: the V-cycle has no host round trip at all.
#include "xsb.h"
#include <cub/cub.cuh>

static inline unsigned nblk(int64_t n, int bs = 256) { return (unsigned)((n + bs - 1) / bs); }

// ------------------------------------------------------------------ K8: transfers
// The fine vector lives on a (possibly slab-local) lattice whose plane 0 is global plane fz0; the coarse vector is
// always indexed globally.  bc = P^T rf for the coarse planes [K0,K1): one thread per coarse node, fine
// contributions gathered in ascending fine index (MatRestrict).
template <int BS>
__global__ void k_restrict(int fnx, int fny, int fnz, int fz0, int cnx, int cny, int K0, int K1, const double *__restrict__ rf, double *__restrict__ bc)
{
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; if (t >= (int64_t)cnx * cny * (K1 - K0)) return;
  const int I = (int)(t % cnx), J = (int)((t / cnx) % cny), K = K0 + (int)(t / ((int64_t)cnx * cny));
  double acc[BS];
#pragma unroll
  for (int d = 0; d < BS; ++d) acc[d] = 0.0;
  for (int c = -1; c <= 1; ++c) for (int b = -1; b <= 1; ++b) for (int a = -1; a <= 1; ++a) {
    const int i = 2 * I + a, j = 2 * J + b, k = 2 * K + c - fz0;
    if (i < 0 || i >= fnx || j < 0 || j >= fny || k < 0 || k >= fnz) continue;
    const double w = (a ? 0.5 : 1.0) * (b ? 0.5 : 1.0) * (c ? 0.5 : 1.0);
    const int64_t f = i + (int64_t)j * fnx + (int64_t)k * fnx * fny;
#pragma unroll
    for (int d = 0; d < BS; ++d) acc[d] += w * rf[BS * f + d];
  }
  const int64_t cn = I + (int64_t)J * cnx + (int64_t)K * cnx * cny;
#pragma unroll
  for (int d = 0; d < BS; ++d) bc[BS * cn + d] = acc[d];
}
// xf += P xc for the fine planes [z0,z1) (local indices): one thread per fine node (MatInterpolateAdd)
template <int BS>
__global__ void k_prolong_add(int fnx, int fny, int z0, int z1, int fz0, int cnx, int cny, const double *__restrict__ xc, double *__restrict__ xf)
{
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; if (t >= (int64_t)fnx * fny * (z1 - z0)) return;
  const int i = (int)(t % fnx), j = (int)((t / fnx) % fny), kl = z0 + (int)(t / ((int64_t)fnx * fny)), k = kl + fz0;
  const int i0 = i >> 1, j0 = j >> 1, k0 = k >> 1, ni = 1 + (i & 1), nj = 1 + (j & 1), nk = 1 + (k & 1);
  const double w = (ni == 2 ? 0.5 : 1.0) * (nj == 2 ? 0.5 : 1.0) * (nk == 2 ? 0.5 : 1.0);
  double acc[BS];
#pragma unroll
  for (int d = 0; d < BS; ++d) acc[d] = 0.0;
  for (int c = 0; c < nk; ++c) for (int b = 0; b < nj; ++b) for (int a = 0; a < ni; ++a) {
    const int64_t cn = (i0 + a) + (int64_t)(j0 + b) * cnx + (int64_t)(k0 + c) * cnx * cny;
#pragma unroll
    for (int d = 0; d < BS; ++d) acc[d] += w * xc[BS * cn + d];
  }
  const int64_t f = i + (int64_t)j * fnx + (int64_t)kl * fnx * fny;
#pragma unroll
  for (int d = 0; d < BS; ++d) xf[BS * f + d] += acc[d];
}
// Slabs.  The fine level lives on the rank's local lattice (ghost planes refreshed in front of every product); a plane-
// distributed coarse level (Level::pdist) keeps global indexing and is current on the planes [rp0-1, rp1+1): its products
// exchange their OUTPUT planes, so a residual or an iterate handed to the transfers already carries valid neighbour planes.
int mg_restrict(xsb_ctx c, const Level &F, const Level &C, const double *rf, double *bc)
{
  const Slab &S = c->slab;
  int fz0 = 0, K0 = 0, K1 = C.nz;
  if (F.dist) {   // owned coarse planes only; the ghost plane below is refreshed first
    XSB_CHK(comm_halo_u(c, const_cast<double *>(rf)));
    fz0 = 2 * S.e0; K0 = S.k0; K1 = S.rank == S.nranks - 1 ? C.nz : S.k1;
  } else if (F.pdist) { K0 = (F.rp0 + 1) / 2; K1 = (F.rp1 + 1) / 2; }   // coarse plane K sits on fine plane 2K: owned with it
  const int64_t nc = (int64_t)C.nx * C.ny * (K1 - K0);
  if (nc > 0) {
    if (c->nsd == 3) k_restrict<3><<<nblk(nc, 128), 128, 0, c->stream>>>(F.nx, F.ny, F.nz, fz0, C.nx, C.ny, K0, K1, rf, bc);
    else k_restrict<2><<<nblk(nc, 128), 128, 0, c->stream>>>(F.nx, F.ny, F.nz, fz0, C.nx, C.ny, K0, K1, rf, bc);
    KERNEL_OK();
  }
  if (C.pdist) return 0;   // the right-hand side is only read on owned rows
  if (F.dist) XSB_CHK(comm_bcast_planes(c, bc, (int64_t)c->nsd * C.nx * C.ny, C.nz));   // replicated coarse level
  else if (F.pdist) XSB_CHK(comm_bcast_plane_ranges(c, bc, (int64_t)c->nsd * C.nx * C.ny, F.cr0.data(), F.cr1.data()));
  return 0;
}
int mg_prolong_add(xsb_ctx c, const Level &F, const Level &C, const double *xc, double *xf)
{
  const Slab &S = c->slab;
  int z0 = 0, z1 = F.nz, fz0 = 0;
  if (F.dist) { z0 = S.ou0; z1 = S.ou1; fz0 = 2 * S.e0; }
  else if (F.pdist) { z0 = F.rp0 - 1 < 0 ? 0 : F.rp0 - 1; z1 = F.rp1 + 1 > F.nz ? F.nz : F.rp1 + 1; }   // neighbour planes too: their coarse parents are current on this rank
  const int64_t nf = (int64_t)F.nx * F.ny * (z1 - z0);
  if (c->nsd == 3) k_prolong_add<3><<<nblk(nf), 256, 0, c->stream>>>(F.nx, F.ny, z0, z1, fz0, C.nx, C.ny, xc, xf);
  else k_prolong_add<2><<<nblk(nf), 256, 0, c->stream>>>(F.nx, F.ny, z0, z1, fz0, C.nx, C.ny, xc, xf);
  KERNEL_OK(); return 0;
}

// scalar (pressure-lattice) transfers of the monolithic -mg path: the same stencils with one dof per node
int mg_restrict_scalar(xsb_ctx c, int fnx, int fny, int fnz, int cnx, int cny, int cnz, const double *rf, double *bc)
{
  const int64_t nc = (int64_t)cnx * cny * cnz;
  k_restrict<1><<<nblk(nc, 128), 128, 0, c->stream>>>(fnx, fny, fnz, 0, cnx, cny, 0, cnz, rf, bc); KERNEL_OK(); return 0;
}
int mg_prolong_add_scalar(xsb_ctx c, int fnx, int fny, int fnz, int cnx, int cny, const double *xc, double *xf)
{
  const int64_t nf = (int64_t)fnx * fny * fnz;
  k_prolong_add<1><<<nblk(nf), 256, 0, c->stream>>>(fnx, fny, 0, fnz, 0, cnx, cny, xc, xf); KERNEL_OK(); return 0;
}

// ------------------------------------------------------------------ K9: Galerkin coarse operator
__global__ void k_box_len(BoxPattern p, int64_t *len)
{
  int64_t nd = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; if (nd >= (int64_t)p.nx * p.ny * p.nz) return;
  len[nd] = box_size(p, (int)(nd % p.nx), (int)((nd / p.nx) % p.ny), (int)(nd / ((int64_t)p.nx * p.ny)));
}
__global__ void k_narrow(int64_t n, const int64_t *a, int *b) { int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; if (i <= n) b[i] = (int)a[i]; }

// A_c[I,J] = sum_{f in N(I)} sum_{g in N(J), g in box(f)} w(f,I) w(g,J) A[f,g];  one thread per (I, J) coarse block
template <int BS>
__global__ void __launch_bounds__(128) k_galerkin(BoxPattern fp, const int *__restrict__ fia, const double *__restrict__ fa,
                                                  BoxPattern cp, const int *__restrict__ cia, int *__restrict__ cja, double *__restrict__ ca)
{
  constexpr int BS2 = BS * BS;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, ncn = (int64_t)cp.nx * cp.ny * cp.nz;
  const int64_t I = t / 27; const int s27 = (int)(t - I * 27);
  if (I >= ncn) return;
  const int Ii = (int)(I % cp.nx), Ij = (int)((I / cp.nx) % cp.ny), Ik = (int)(I / ((int64_t)cp.nx * cp.ny));
  const int Ji = Ii + s27 % 3 - 1, Jj = Ij + (s27 / 3) % 3 - 1, Jk = Ik + s27 / 9 - 1;
  if (Ji < 0 || Ji >= cp.nx || Jj < 0 || Jj >= cp.ny || Jk < 0 || Jk >= cp.nz) return;
  const int slot = box_slot(cp, Ii, Ij, Ik, Ji, Jj, Jk);
  double acc[BS2];
#pragma unroll
  for (int x = 0; x < BS2; ++x) acc[x] = 0.0;
  for (int fk = 2 * Ik - 1; fk <= 2 * Ik + 1; ++fk) {
    if (fk < 0 || fk >= fp.nz) continue;
    for (int fj = 2 * Ij - 1; fj <= 2 * Ij + 1; ++fj) {
      if (fj < 0 || fj >= fp.ny) continue;
      for (int fi = 2 * Ii - 1; fi <= 2 * Ii + 1; ++fi) {
        if (fi < 0 || fi >= fp.nx) continue;
        const double wf = (fi == 2 * Ii ? 1.0 : 0.5) * (fj == 2 * Ij ? 1.0 : 0.5) * (fk == 2 * Ik ? 1.0 : 0.5);
        int l0, h0, l1, h1, l2, h2;
        box_range(fp, fi, fp.nx, l0, h0); box_range(fp, fj, fp.ny, l1, h1); box_range(fp, fk, fp.nz, l2, h2);
        const int nx = h0 - l0 + 1, ny = h1 - l1 + 1;
        const int64_t f = fi + (int64_t)fj * fp.nx + (int64_t)fk * fp.nx * fp.ny;
        const double *rowv = fa + (int64_t)fia[f] * BS2;
        const int g0 = max(2 * Ji - 1, l0), g1 = min(2 * Ji + 1, h0), gj0 = max(2 * Jj - 1, l1), gj1 = min(2 * Jj + 1, h1), gk0 = max(2 * Jk - 1, l2), gk1 = min(2 * Jk + 1, h2);
        for (int gk = gk0; gk <= gk1; ++gk) for (int gj = gj0; gj <= gj1; ++gj) for (int gi = g0; gi <= g1; ++gi) {
          const double w = wf * ((gi == 2 * Ji ? 1.0 : 0.5) * (gj == 2 * Jj ? 1.0 : 0.5) * (gk == 2 * Jk ? 1.0 : 0.5));
          const double *blk = rowv + (int64_t)(((gk - l2) * ny + (gj - l1)) * nx + (gi - l0)) * BS2;
#pragma unroll
          for (int x = 0; x < BS2; ++x) acc[x] += w * blk[x];
        }
      }
    }
  }
  const int64_t pos = cia[I] + slot;
  cja[pos] = (int)(Ji + (int64_t)Jj * cp.nx + (int64_t)Jk * cp.nx * cp.ny);
#pragma unroll
  for (int x = 0; x < BS2; ++x) ca[pos * BS2 + x] = acc[x];
}

static int galerkin(xsb_ctx c, const Level &F, Level &C)
{
  const int bs = c->nsd; cudaStream_t st = c->stream;
  BoxPattern cp{C.nx, C.ny, C.nz, 0};
  const int64_t ncn = (int64_t)C.nx * C.ny * C.nz;
  int64_t *len = nullptr; CUDA_OK(cudaMalloc(&len, sizeof(int64_t) * (ncn + 1)));
  k_box_len<<<nblk(ncn), 256, 0, st>>>(cp, len); KERNEL_OK();
  CUDA_OK(cudaMemsetAsync(len + ncn, 0, sizeof(int64_t), st));
  void *tmp = nullptr; size_t tb = 0;
  CUDA_OK(cub::DeviceScan::ExclusiveSum(nullptr, tb, len, len, ncn + 1, st));
  CUDA_OK(cudaMalloc(&tmp, tb));
  CUDA_OK(cub::DeviceScan::ExclusiveSum(tmp, tb, len, len, ncn + 1, st));
  int64_t tot = 0; CUDA_OK(cudaMemcpyAsync(&tot, len + ncn, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  CUDA_OK(cudaStreamSynchronize(st)); CUDA_OK(cudaFree(tmp));
  Baij &A = C.A; A.nb = (int)ncn; A.bs = bs; A.nblk = tot; A.pat = cp;
  XSB_CHK(dev_alloc(c, &A.ia, (size_t)ncn + 1)); XSB_CHK(dev_alloc(c, &A.ja, (size_t)tot)); XSB_CHK(dev_alloc(c, &A.a, (size_t)tot * bs * bs + 2));
  k_narrow<<<nblk(ncn + 1), 256, 0, st>>>(ncn, len, A.ia); KERNEL_OK();
  if (bs == 3) k_galerkin<3><<<nblk(ncn * 27, 128), 128, 0, st>>>(F.A.pat, F.A.ia, F.A.a, cp, A.ia, A.ja, A.a);
  else k_galerkin<2><<<nblk(ncn * 27, 128), 128, 0, st>>>(F.A.pat, F.A.ia, F.A.a, cp, A.ia, A.ja, A.a);
  KERNEL_OK();
  CUDA_OK(cudaStreamSynchronize(st)); CUDA_OK(cudaFree(len));
  C.owns_A = true;
  return 0;
}

// box-pattern block columns of a lattice matrix
__global__ void k_box_ja(BoxPattern p, const int *__restrict__ ia, int *__restrict__ ja)
{
  int64_t nd = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; if (nd >= (int64_t)p.nx * p.ny * p.nz) return;
  const int i = (int)(nd % p.nx), j = (int)((nd / p.nx) % p.ny), k = (int)(nd / ((int64_t)p.nx * p.ny));
  int l0, h0, l1, h1, l2, h2; box_range(p, i, p.nx, l0, h0); box_range(p, j, p.ny, l1, h1); box_range(p, k, p.nz, l2, h2);
  int cpos = ia[nd];
  for (int kk = l2; kk <= h2; ++kk) for (int jj = l1; jj <= h1; ++jj) for (int ii = l0; ii <= h0; ++ii) ja[cpos++] = ii + jj * p.nx + kk * p.nx * p.ny;
}

// Slab-partitioned fine level -> replicated first coarse level.  Every rank forms P^T A P on its local lattice (the
// rows of its owned coarse planes are complete: SURVEY 8e / Slab comment), copies them into the global box-pattern
// matrix and all ranks exchange their row ranges (ncclBroadcast per rank; set-up only).
static int galerkin_replicate(xsb_ctx c, const Level &F, Level &C)
{
  const int bs = c->nsd, bs2 = bs * bs; cudaStream_t st = c->stream; const Slab &S = c->slab;
  Level T; T.nx = (F.nx - 1) / 2 + 1; T.ny = (F.ny - 1) / 2 + 1; T.nz = (F.nz - 1) / 2 + 1;
  if (c->no_A) XSB_CHK(galerkin_elements(c, T)); else XSB_CHK(galerkin(c, F, T));
  BoxPattern gp{C.nx, C.ny, C.nz, 0};
  const int64_t ncn = (int64_t)C.nx * C.ny * C.nz, pn = (int64_t)C.nx * C.ny;
  int64_t *len = nullptr; CUDA_OK(cudaMalloc(&len, sizeof(int64_t) * (ncn + 1)));
  k_box_len<<<nblk(ncn), 256, 0, st>>>(gp, len); KERNEL_OK();
  CUDA_OK(cudaMemsetAsync(len + ncn, 0, sizeof(int64_t), st));
  void *tmp = nullptr; size_t tb = 0;
  CUDA_OK(cub::DeviceScan::ExclusiveSum(nullptr, tb, len, len, ncn + 1, st));
  CUDA_OK(cudaMalloc(&tmp, tb));
  CUDA_OK(cub::DeviceScan::ExclusiveSum(tmp, tb, len, len, ncn + 1, st));
  std::vector<int64_t> gia(ncn + 1);
  CUDA_OK(cudaMemcpyAsync(gia.data(), len, sizeof(int64_t) * (ncn + 1), cudaMemcpyDeviceToHost, st));
  CUDA_OK(cudaStreamSynchronize(st)); CUDA_OK(cudaFree(tmp));
  Baij &A = C.A; A.nb = (int)ncn; A.bs = bs; A.nblk = gia[ncn]; A.pat = gp;
  XSB_CHK(dev_alloc(c, &A.ia, (size_t)ncn + 1)); XSB_CHK(dev_alloc(c, &A.ja, (size_t)A.nblk)); XSB_CHK(dev_alloc(c, &A.a, (size_t)A.nblk * bs2 + 2));
  k_narrow<<<nblk(ncn + 1), 256, 0, st>>>(ncn, len, A.ia); KERNEL_OK();
  k_box_ja<<<nblk(ncn), 256, 0, st>>>(gp, A.ia, A.ja); KERNEL_OK();
  // my rows: global coarse planes [K0,K1) = local planes [K0-e0, K1-e0) of T (identical boxes, see above)
  const int K0 = S.k0, K1 = S.rank == S.nranks - 1 ? C.nz : S.k1;
  std::vector<int> tia(T.A.nb + 1);
  CUDA_OK(cudaMemcpy(tia.data(), T.A.ia, sizeof(int) * (T.A.nb + 1), cudaMemcpyDeviceToHost));
  const int64_t g0 = gia[K0 * pn], g1 = gia[K1 * pn], l0 = tia[(K0 - S.e0) * pn], l1 = tia[(K1 - S.e0) * pn];
  if (g1 - g0 != l1 - l0) return xsb_fail(c, XSB_ERR_ARG, "slab Galerkin: local and global row ranges differ (%lld vs %lld blocks)", (long long)(l1 - l0), (long long)(g1 - g0));
  CUDA_OK(cudaMemcpyAsync(A.a + g0 * bs2, T.A.a + l0 * bs2, sizeof(double) * (g1 - g0) * bs2, cudaMemcpyDeviceToDevice, st));
  std::vector<int64_t> offs(S.nranks + 1);
  for (int r = 0; r < S.nranks; ++r) { int a0, a1; xsb_slab_range(S.mz_glob, S.nranks, r, &a0, &a1); offs[r] = gia[a0 * pn] * bs2; }
  offs[S.nranks] = gia[ncn] * bs2;
  XSB_CHK(comm_bcast_segments(c, A.a, offs.data()));
  CUDA_OK(cudaStreamSynchronize(st)); CUDA_OK(cudaFree(len));
  dev_free(c, T.A.ia); dev_free(c, T.A.ja); dev_free(c, T.A.a);   // the local product was only the source of the owned rows
  C.owns_A = true;
  return 0;
}

// ------------------------------------------------------------------ Jacobi (PCSetUp_Jacobi: 1/diag, zero -> 1)
template <int BS>
__global__ void k_baij_idiag(BoxPattern p, const int *__restrict__ ia, const double *__restrict__ a, double *__restrict__ idiag)
{
  int64_t nd = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; if (nd >= (int64_t)p.nx * p.ny * p.nz) return;
  const int i = (int)(nd % p.nx), j = (int)((nd / p.nx) % p.ny), k = (int)(nd / ((int64_t)p.nx * p.ny));
  const int slot = box_slot(p, i, j, k, i, j, k);
  const double *blk = a + (int64_t)(ia[nd] + slot) * BS * BS;
  for (int d = 0; d < BS; ++d) { double v = blk[d * BS + d]; idiag[BS * nd + d] = v == 0.0 ? 1.0 : 1.0 / v; }
}
int baij_diag_inv(xsb_ctx c, const Baij &A, double *idiag)
{
  if (A.bs == 3) k_baij_idiag<3><<<nblk(A.nb), 256, 0, c->stream>>>(A.pat, A.ia, A.a, idiag);
  else k_baij_idiag<2><<<nblk(A.nb), 256, 0, c->stream>>>(A.pat, A.ia, A.a, idiag);
  KERNEL_OK(); return 0;
}
__global__ void k_csr_idiag(int n, const int *__restrict__ ia, const int *__restrict__ ja, const double *__restrict__ a, double *__restrict__ idiag)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x; if (i >= n) return;
  double v = 0.0; for (int k = ia[i]; k < ia[i + 1]; ++k) if (ja[k] == i) v = a[k];
  idiag[i] = v == 0.0 ? 1.0 : 1.0 / v;
}
template <int BS>
__global__ void k_baij_diag(BoxPattern p, const int *__restrict__ ia, const double *__restrict__ a, double *__restrict__ diag)
{
  int64_t nd = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; if (nd >= (int64_t)p.nx * p.ny * p.nz) return;
  const int i = (int)(nd % p.nx), j = (int)((nd / p.nx) % p.ny), k = (int)(nd / ((int64_t)p.nx * p.ny));
  const double *blk = a + (int64_t)(ia[nd] + box_slot(p, i, j, k, i, j, k)) * BS * BS;
  for (int d = 0; d < BS; ++d) diag[BS * nd + d] = blk[d * BS + d];
}
__global__ void k_csr_diag(int n, const int *__restrict__ ia, const int *__restrict__ ja, const double *__restrict__ a, double *__restrict__ diag)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x; if (i >= n) return;
  double v = 0.0; for (int k = ia[i]; k < ia[i + 1]; ++k) if (ja[k] == i) v = a[k];
  diag[i] = v;
}
// MatGetDiagonal on the device
int baij_diag(xsb_ctx c, const Baij &A, double *diag)
{
  if (A.bs == 3) k_baij_diag<3><<<nblk(A.nb), 256, 0, c->stream>>>(A.pat, A.ia, A.a, diag);
  else k_baij_diag<2><<<nblk(A.nb), 256, 0, c->stream>>>(A.pat, A.ia, A.a, diag);
  KERNEL_OK(); return 0;
}
int csr_diag(xsb_ctx c, const Csr &A, double *diag) { k_csr_diag<<<nblk(A.n), 256, 0, c->stream>>>(A.n, A.ia, A.ja, A.a, diag); KERNEL_OK(); return 0; }
int csr_diag_inv(xsb_ctx c, const Csr &A, double *idiag) { k_csr_idiag<<<nblk(A.n), 256, 0, c->stream>>>(A.n, A.ia, A.ja, A.a, idiag); KERNEL_OK(); return 0; }

// BAIJ -> scalar CSR on the host (MatGetRowIJ view of a level operator; used by tests / xsb_mat_get_csr)
int baij_to_csr_host(xsb_ctx c, const Baij &A, int32_t *ia, int32_t *ja, double *a)
{
  const int bs = A.bs, bs2 = bs * bs;
  std::vector<int> bia(A.nb + 1), bja(A.nblk); std::vector<double> ba;
  CUDA_OK(cudaMemcpy(bia.data(), A.ia, sizeof(int) * (A.nb + 1), cudaMemcpyDeviceToHost));
  CUDA_OK(cudaMemcpy(bja.data(), A.ja, sizeof(int) * A.nblk, cudaMemcpyDeviceToHost));
  if (a) { ba.resize((size_t)A.nblk * bs2); CUDA_OK(cudaMemcpy(ba.data(), A.a, sizeof(double) * A.nblk * bs2, cudaMemcpyDeviceToHost)); }
  int64_t pos = 0;
  for (int nd = 0; nd < A.nb; ++nd) {
    const int nb = bia[nd + 1] - bia[nd];
    for (int x = 0; x < bs; ++x) {
      if (ia) ia[bs * nd + x] = (int32_t)pos;
      for (int s = 0; s < nb; ++s) for (int y = 0; y < bs; ++y) {
        if (ja) ja[pos] = bs * bja[bia[nd] + s] + y;
        if (a) a[pos] = ba[(size_t)(bia[nd] + s) * bs2 + x * bs + y];
        pos++;
      }
    }
  }
  if (ia) ia[bs * A.nb] = (int32_t)pos;
  return 0;
}

// ------------------------------------------------------------------ K11: dense inverse of the coarsest operator
template <int BS>
__global__ void k_baij_to_dense(int nb, const int *__restrict__ ia, const int *__restrict__ ja, const double *__restrict__ a, int n, double *__restrict__ M)
{
  int64_t blk = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  // one thread per block row (coarsest level is tiny)
  if (blk >= nb) return;
  for (int s = ia[blk]; s < ia[blk + 1]; ++s) for (int x = 0; x < BS; ++x) for (int y = 0; y < BS; ++y)
    M[(int64_t)(BS * blk + x) * n + BS * ja[s] + y] = a[(int64_t)s * BS * BS + x * BS + y];
}
__global__ void k_gj_prepare(int n, int k, const double *__restrict__ M, double *__restrict__ colk, double *__restrict__ rowk, int *__restrict__ flag)
{
  int t = blockIdx.x * blockDim.x + threadIdx.x; if (t >= n) return;
  const double piv = M[(int64_t)k * n + k];
  if (t == 0 && piv == 0.0) *flag = 1;
  const double p = 1.0 / piv;
  colk[t] = M[(int64_t)t * n + k];
  rowk[t] = t == k ? p : M[(int64_t)k * n + t] * p;
}
__global__ void k_gj_update(int n, int k, double *__restrict__ M, const double *__restrict__ colk, const double *__restrict__ rowk)
{
  const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
  if (j >= n) return;
  double *m = M + (int64_t)i * n + j;
  if (i == k) *m = rowk[j];
  else if (j == k) *m = -colk[i] * rowk[k];
  else *m = *m - colk[i] * rowk[j];
}
__global__ void k_gemv(int n, const double *__restrict__ M, const double *__restrict__ b, double *__restrict__ x)
{
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= n) return;
  double acc = 0.0;
  for (int j = lane; j < n; j += 32) acc += M[(int64_t)row * n + j] * b[j];
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) x[row] = acc;
}
static int coarse_invert(xsb_ctx c, Level &L)
{
  const int n = L.A.nb * L.A.bs; cudaStream_t st = c->stream;
  if (n > 6600) return xsb_fail(c, XSB_ERR_SUP, "coarsest MG level has %d dofs; the dense coarse solve supports <= 6600 (use more -saddle_fieldsplit_u_pc_mg_levels)", n);
  XSB_CHK(dev_alloc(c, &L.inv, (size_t)n * n));
  CUDA_OK(cudaMemsetAsync(L.inv, 0, sizeof(double) * n * n, st));
  if (L.A.bs == 3) k_baij_to_dense<3><<<nblk(L.A.nb, 64), 64, 0, st>>>(L.A.nb, L.A.ia, L.A.ja, L.A.a, n, L.inv);
  else k_baij_to_dense<2><<<nblk(L.A.nb, 64), 64, 0, st>>>(L.A.nb, L.A.ia, L.A.ja, L.A.a, n, L.inv);
  KERNEL_OK();
  double *colk = nullptr, *rowk = nullptr; int *flag = nullptr;
  XSB_CHK(dev_alloc(c, &colk, (size_t)n)); XSB_CHK(dev_alloc(c, &rowk, (size_t)n)); XSB_CHK(dev_alloc(c, &flag, 1));
  CUDA_OK(cudaMemsetAsync(flag, 0, sizeof(int), st));
  dim3 g2((n + 255) / 256, n);
  for (int k = 0; k < n; ++k) {
    k_gj_prepare<<<nblk(n), 256, 0, st>>>(n, k, L.inv, colk, rowk, flag); KERNEL_OK();
    k_gj_update<<<g2, 256, 0, st>>>(n, k, L.inv, colk, rowk); KERNEL_OK();
  }
  int hflag = 0; CUDA_OK(cudaMemcpyAsync(&hflag, flag, sizeof(int), cudaMemcpyDeviceToHost, st)); CUDA_OK(cudaStreamSynchronize(st));
  if (hflag) return xsb_fail(c, XSB_ERR_BREAKDOWN, "zero pivot in the coarse-level factorisation");
  return 0;
}

// ------------------------------------------------------------------ Chebyshev / Jacobi smoother (App. B.4)
__global__ void k_cheb_first_zero(int64_t n, double scale, const double *__restrict__ idiag, const double *__restrict__ b, double *__restrict__ p)
{ for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = 0.0 + scale * (idiag[i] * b[i]); }

static int a00_spmv(xsb_ctx c, const Level &L, bool fine, const double *x, double *y, const Epilogue &ep)
{
  if (fine) return spmv_a00_fine(c, L.A, x, y, ep);
  const int lcat = PROF_LVL + (int)(&L - (&L >= c->sub && &L < c->sub + XSB_MAX_LEVELS ? c->sub : c->lev));
  XSB_CHK(prof_mark(c, lcat));
  if (!L.pdist) { XSB_CHK(spmv_baij(c, L.A, x, y, ep)); return prof_mark(c, PROF_OTHER); }
  // plane-distributed level: this rank computes the rows of its node planes (rows and fused epilogue are per row, so the
  // values are bitwise those of the one-GPU product) and trades one plane of the result with each neighbour
  const int pn = L.nx * L.ny;
  if (c->side && L.rp1 - L.rp0 >= 4 && c->opt.integer("xsb_overlap", 1)) {
    // the two planes the neighbours need first; their exchange (a synchronisation point with both neighbours) then runs on a
    // second stream beside the interior rows -- fork / join by events, so the pattern is captured into the V-cycle graphs as is
    XSB_CHK(spmv_baij_pair(c, L.A, x, y, ep, L.rp0 * pn, (L.rp1 - 1) * pn, pn));
    CUDA_OK(cudaEventRecord(c->ev_fork, c->stream)); CUDA_OK(cudaStreamWaitEvent(c->side, c->ev_fork, 0));
    XSB_CHK(comm_halo_planes(c, y, (int64_t)L.A.bs * pn, L.rp0, L.rp1, 1, 1, c->side));
    CUDA_OK(cudaEventRecord(c->ev_join, c->side));
    XSB_CHK(spmv_baij(c, L.A, x, y, ep, (L.rp0 + 1) * pn, (L.rp1 - L.rp0 - 2) * pn));
    XSB_CHK(prof_mark(c, PROF_CHALO));
    CUDA_OK(cudaStreamWaitEvent(c->stream, c->ev_join, 0));
    return prof_mark(c, PROF_OTHER);
  }
  XSB_CHK(spmv_baij(c, L.A, x, y, ep, L.rp0 * pn, (L.rp1 - L.rp0) * pn));
  XSB_CHK(prof_mark(c, PROF_CHALO));
  XSB_CHK(comm_halo_planes(c, y, (int64_t)L.A.bs * pn, L.rp0, L.rp1, 1, 1));
  return prof_mark(c, PROF_OTHER);
}
// set-up products (eigenvalue estimate): not counted; slab levels refresh ghosts and compute owned rows only
static int level_spmv(xsb_ctx c, const Level &L, const double *x, double *y)
{
  Epilogue ep;
  if (c->no_A && &L == &c->lev[c->nlev - 1]) {   // operator-free fine level
    XSB_CHK(comm_halo_u(c, const_cast<double *>(x)));
    return mf_a00_apply(c, x, y, ep);
  }
  if (!L.dist) return spmv_baij(c, L.A, x, y, ep);
  XSB_CHK(comm_halo_u(c, const_cast<double *>(x)));
  const int pn = L.nx * L.ny;
  return spmv_baij(c, L.A, x, y, ep, c->slab.ou0 * pn, (c->slab.ou1 - c->slab.ou0) * pn);
}

// KSPSolve_Chebyshev, Jacobi PC, nonzero initial guess, norm none: first correction + (its-1) recurrence steps.
// Result ends in L.x (buffers are rotated by pointer swap, no copies).
static int cheb_smooth(xsb_ctx c, Level &L, bool fine, int its, bool x_is_zero)
{
  if (its < 1) return 0;
  const int64_t n = (int64_t)L.A.nb * L.A.bs;
  const double scale = 2.0 / (L.emax + L.emin), alpha = 1.0 - scale * L.emin, mu = 1.0 / alpha, omegaprod = 2.0 / alpha;
  double ckm1 = 1.0, ck = mu;
  double *pkm1 = L.x, *pk = L.w0, *pkp1 = L.w1;
  if (x_is_zero && L.pdist) {   // owned planes, then the neighbours' (the right-hand side is only current on owned rows)
    const int64_t pd = (int64_t)L.A.bs * L.nx * L.ny, o = L.rp0 * pd, m = (L.rp1 - L.rp0) * pd;
    k_cheb_first_zero<<<nblk(m) > 2368 ? 2368 : nblk(m), 256, 0, c->stream>>>(m, scale, L.idiag + o, L.b + o, pk + o); KERNEL_OK();
    XSB_CHK(comm_halo_planes(c, pk, pd, L.rp0, L.rp1, 1, 1));
  } else if (x_is_zero) {   // r = b - A*0 = b exactly: skip the product (bitwise identical)
    k_cheb_first_zero<<<nblk(n) > 2368 ? 2368 : nblk(n), 256, 0, c->stream>>>(n, scale, L.idiag, L.b, pk); KERNEL_OK();
  } else {
    Epilogue ep; ep.mode = EPI_CHEB_FIRST; ep.b = L.b; ep.idiag = L.idiag; ep.pk = pkm1; ep.s0 = scale;
    XSB_CHK(a00_spmv(c, L, fine, pkm1, pk, ep));
  }
  for (int it = 1; it < its; ++it) {
    const double ckp1 = 2.0 * mu * ck - ckm1, omega = omegaprod * ck / ckp1;
    Epilogue ep; ep.mode = EPI_CHEB; ep.b = L.b; ep.idiag = L.idiag; ep.pk = pk; ep.pkm1 = pkm1;
    ep.s0 = 1.0 - omega; ep.s1 = omega; ep.s2 = omega * scale;
    XSB_CHK(a00_spmv(c, L, fine, pk, pkp1, ep));
    ckm1 = ck; ck = ckp1;
    double *t = pkm1; pkm1 = pk; pk = pkp1; pkp1 = t;
  }
  L.x = pk; L.w0 = pkm1; L.w1 = pkp1;
  return 0;
}

static int coarse_pcg(xsb_ctx c);

// One V-cycle level of hierarchy H (c->lev: the PCMG hierarchy of the options; c->sub: the internal one under a large coarsest level)
static int mg_cycle(xsb_ctx c, Level *H, int l, int top, bool main)
{
  Level &L = H[l];
  if (l == 0) {   // coarse grid: preonly + LU  ->  x = A^-1 b
    if (!L.inv) return coarse_pcg(c);   // too large for the dense inverse: solved to LU accuracy by V-cycle-preconditioned CG
    const int n = L.A.nb * L.A.bs;
    XSB_CHK(prof_mark(c, PROF_LVL));
    k_gemv<<<nblk((int64_t)n * 32), 256, 0, c->stream>>>(n, L.inv, L.b, L.x); KERNEL_OK();
    return prof_mark(c, PROF_OTHER);
  }
  Level &C = H[l - 1];
  const bool fine = main && l == top;
  const int64_t n = (int64_t)L.A.nb * L.A.bs;
  XSB_CHK(vec_set(c, n, 0.0, L.x));
  XSB_CHK(cheb_smooth(c, L, fine, c->so.cheb_its, true));                 // pre-smooth
  { Epilogue ep; ep.mode = EPI_RESIDUAL; ep.b = L.b; XSB_CHK(a00_spmv(c, L, fine, L.x, L.r, ep)); }   // r = b - A x
  XSB_CHK(prof_mark(c, PROF_XFER)); XSB_CHK(mg_restrict(c, L, C, L.r, C.b)); XSB_CHK(prof_mark(c, PROF_OTHER));
  XSB_CHK(mg_cycle(c, H, l - 1, top, main));
  XSB_CHK(prof_mark(c, PROF_XFER)); XSB_CHK(mg_prolong_add(c, L, C, C.x, L.x)); XSB_CHK(prof_mark(c, PROF_OTHER));
  XSB_CHK(cheb_smooth(c, L, fine, c->so.cheb_its, false));                // post-smooth
  return 0;
}

// Coarsest level of abf.opts' 3-level hierarchy at the BASELINE sizes (14 739 rows at 32^3, 107 811 at 64^3; the reference
// factors it with UMFPACK, abf.opts:16): conjugate gradients on the (symmetric positive definite) Galerkin operator,
// preconditioned by one V-cycle of the internal hierarchy c->sub, to a relative residual of 1e-13 -- the accuracy of a
// direct factorisation, so the outer iteration counts are those of the reference's LU.
static __global__ void k_cg_update(int64_t n, const double *__restrict__ sc /* [rz, pq] */, const double *__restrict__ p, const double *__restrict__ q, double *__restrict__ x, double *__restrict__ r)
{
  const double alpha = sc[0] / sc[1];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) { x[i] += alpha * p[i]; r[i] -= alpha * q[i]; }
}
static int coarse_pcg(xsb_ctx c)
{
  Level &L = c->lev[0]; const int top = c->nsub - 1; Level &S = c->sub[top];
  const int64_t n = (int64_t)L.A.nb * L.A.bs; const Ranges rg = whole(n);
  double *r = L.r, *p = c->cg_p, *q = c->cg_q, *sc = c->scal + 192;   // sc[0] r.z, sc[1] p.q, sc[2] r.r
  double h[4], rz_old = 0.0, bb = 0.0; Epilogue plain;
  XSB_CHK(vec_set(c, n, 0.0, L.x));
  XSB_CHK(vec_copy(c, n, L.b, r));
  c->coarse_solves++;
  for (int it = 0; it < 500; ++it) {
    double *save = S.b; S.b = r;                       // z = M^-1 r: one V-cycle of the internal hierarchy on r
    int rc = mg_cycle(c, c->sub, top, top, false);
    S.b = save; if (rc) return rc;
    double *z = S.x;
    { double *two[1] = {z}; XSB_CHK(vec_mdot(c, rg, r, two, 1, true, sc, true)); }   // [r.z, r.r] (replicated level: local sums)
    XSB_CHK(vec_fetch(c, sc, 2, h));
    const double rz = h[0], rr = h[1];
    if (it == 0) bb = rr;
    if (bb == 0.0 || rr <= 1e-26 * bb) return 0;       // ||r|| <= 1e-13 ||b||
    if (!(rz > 0.0)) return xsb_fail(c, XSB_ERR_BREAKDOWN, "coarse-level CG: the V-cycle preconditioner is not positive definite (r.z = %g)", rz);
    if (it == 0) XSB_CHK(vec_copy(c, n, z, p)); else XSB_CHK(vec_aypx(c, n, rz / rz_old, z, p));   // p = z + beta p
    rz_old = rz;
    XSB_CHK(spmv_baij(c, L.A, p, q, plain));
    { double *two[1] = {q}; XSB_CHK(vec_mdot(c, rg, p, two, 1, false, sc + 1, true)); }             // p.q
    k_cg_update<<<nblk(n) > 1184 ? 1184 : nblk(n), 256, 0, c->stream>>>(n, sc, p, q, L.x, r); KERNEL_OK();
    c->coarse_its++;
  }
  return xsb_fail(c, XSB_ERR_BREAKDOWN, "coarse-level CG did not reach 1e-13 in 500 iterations");
}

// PCApply_MG: one multiplicative V-cycle from a zero initial guess.
// The ~150 launches of a cycle (most of them tiny coarse-level kernels whose launch cost exceeds their run time) are captured
// once into a CUDA graph and replayed (-xsb_graph, default on).  The smoothers rotate three buffers per level, so the pointer
// assignment a cycle starts from repeats with period 3: one graph per starting state, found by (rhs, x, w0, w1) of the fine level.
static void vgraph_free(xsb_ctx c)
{
  for (auto &g : c->vgraphs) { if (g.exec) cudaGraphExecDestroy(g.exec); if (g.graph) cudaGraphDestroy(g.graph); }
  c->vgraphs.clear();
}
void mg_graphs_release(xsb_ctx c) { vgraph_free(c); }
int mg_vcycle(xsb_ctx c, const double *b, double *x)
{
  const int top = c->nlev - 1;
  Level &L = c->lev[top];
  const int64_t n = (int64_t)L.A.nb * L.A.bs;
  const bool graphable = c->use_graph && c->nsub == 0 && !c->so.time_kernels && c->nlev > 1;   // coarse CG and event timing need the host
  if (graphable) {
    for (auto &g : c->vgraphs) {
      if (g.b != b || g.x != L.x || g.w0 != L.w0 || g.w1 != L.w1) continue;
      CUDA_OK(cudaGraphLaunch(g.exec, c->stream));
      c->n_a00 += g.d_a00; c->n_launch += g.d_launch; for (int i = 0; i < 4; ++i) c->a00_mode[i] += g.d_mode[i];
      for (int l = 0; l <= top; ++l) { c->lev[l].x = g.px[l]; c->lev[l].w0 = g.pw0[l]; c->lev[l].w1 = g.pw1[l]; }   // the rotation the cycle leaves behind
      c->graph_replays++;
      return vec_copy(c, n, c->lev[top].x, x);
    }
  }
  xsb_ctx_s::VGraph G; G.b = b; G.x = L.x; G.w0 = L.w0; G.w1 = L.w1;
  const int64_t a0 = c->n_a00, l0 = c->n_launch; int64_t m0[4]; for (int i = 0; i < 4; ++i) m0[i] = c->a00_mode[i];
  const bool capture = graphable && c->vgraphs.size() < 8;
  if (capture) CUDA_OK(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
  double *save = L.b; L.b = const_cast<double *>(b);   // the level reads b in place
  int rc = mg_cycle(c, c->lev, top, top, true);
  c->lev[top].b = save;
  if (capture) {
    cudaError_t e = cudaStreamEndCapture(c->stream, &G.graph);
    if (rc) { if (G.graph) cudaGraphDestroy(G.graph); return rc; }
    if (e != cudaSuccess) return xsb_fail(c, XSB_ERR_CUDA, "V-cycle graph capture failed: %s (disable with -xsb_graph 0)", cudaGetErrorString(e));
    CUDA_OK(cudaGraphInstantiate(&G.exec, G.graph, 0));
    G.d_a00 = c->n_a00 - a0; G.d_launch = c->n_launch - l0; for (int i = 0; i < 4; ++i) G.d_mode[i] = c->a00_mode[i] - m0[i];
    for (int l = 0; l <= top; ++l) { G.px[l] = c->lev[l].x; G.pw0[l] = c->lev[l].w0; G.pw1[l] = c->lev[l].w1; }
    c->vgraphs.push_back(G);
    CUDA_OK(cudaGraphLaunch(G.exec, c->stream));   // capturing records, it does not run
  }
  if (rc) return rc;
  return vec_copy(c, n, c->lev[top].x, x);
}

// ------------------------------------------------------------------ eigenvalue estimate (KSPChebyshev esteig)
// eststeps of left-Jacobi GMRES (classical Gram-Schmidt) on the noisy vector; Ritz values of the Hessenberg.
static int cheb_estimate(xsb_ctx c, Level &L)
{
  const SolverOpts &s = c->so; const int m = s.esteig_steps; const int64_t n = (int64_t)L.A.nb * L.A.bs;
  std::vector<double *> V(m + 1, nullptr);
  for (auto &v : V) XSB_CHK(dev_alloc(c, &v, (size_t)n));
  double *t = L.r;
  std::vector<double> H((size_t)(m + 1) * m, 0.0), h(m + 2);
  const Ranges rg = L.dist ? c->own_u : whole(n);
  const int64_t soff = L.dist ? (int64_t)c->nsd * (2 * c->slab.e0) * L.nx * L.ny : 0;   // noise is a function of the GLOBAL dof index
  XSB_CHK(vec_rander48(c, n, s.noise, t, soff));
  XSB_CHK(vec_pmult(c, n, L.idiag, t, V[0]));                       // v0 = M^-1 b (x0 = 0)
  XSB_CHK(vec_mdot(c, rg, V[0], V.data(), 0, true, c->scal));
  XSB_CHK(vec_fetch(c, c->scal, 1, h.data()));
  const double res0 = sqrt(h[0]); double res = res0;
  if (res0 == 0.0) return xsb_fail(c, XSB_ERR_BREAKDOWN, "zero noise vector in the Chebyshev eigenvalue estimate");
  XSB_CHK(vec_scale(c, n, 1.0 / res0, V[0]));
  std::vector<double> cs(m + 1), sn(m + 1), rs(m + 1); rs[0] = res0;
  int it = 0;
  while (it < m) {
    XSB_CHK(level_spmv(c, L, V[it], t));
    XSB_CHK(vec_pmult(c, n, L.idiag, t, V[it + 1]));               // w = M^-1 A v
    XSB_CHK(vec_mdot(c, rg, V[it + 1], V.data(), it + 1, false, c->scal));
    XSB_CHK(vec_maxpy_dev(c, n, V[it + 1], V.data(), it + 1, c->scal, -1.0));
    XSB_CHK(vec_mdot(c, rg, V[it + 1], V.data(), 0, true, c->scal + it + 1));
    XSB_CHK(vec_scale_by_inv_sqrt(c, n, V[it + 1], c->scal + it + 1));
    XSB_CHK(vec_fetch(c, c->scal, it + 2, h.data()));
    const double tt = sqrt(h[it + 1]);
    for (int j = 0; j <= it; ++j) H[(size_t)j * m + it] = h[j];
    H[(size_t)(it + 1) * m + it] = tt;
    // Givens update only to track the residual for the rtol 1e-12 stop (KSPSetTolerances(kspest,1e-12,...))
    std::vector<double> col(h.begin(), h.begin() + it + 1); col.push_back(tt);
    for (int j = 0; j < it; ++j) { double a = col[j]; col[j] = cs[j] * a + sn[j] * col[j + 1]; col[j + 1] = -sn[j] * a + cs[j] * col[j + 1]; }
    const double d = sqrt(col[it] * col[it] + col[it + 1] * col[it + 1]);
    it++;
    if (d == 0.0) break;
    cs[it - 1] = col[it - 1] / d; sn[it - 1] = col[it] / d;
    rs[it] = -sn[it - 1] * rs[it - 1]; rs[it - 1] = cs[it - 1] * rs[it - 1];
    res = fabs(rs[it]);
    if (res <= 1e-12 * res0 || tt == 0.0) break;
  }
  std::vector<double> Hk((size_t)it * it), wr(it), wi(it);
  for (int i = 0; i < it; ++i) for (int j = 0; j < it; ++j) Hk[(size_t)i * it + j] = H[(size_t)i * m + j];
  if (hess_eig(it, Hk.data(), it, wr.data(), wi.data())) return xsb_fail(c, XSB_ERR_BREAKDOWN, "Hessenberg eigenvalue iteration failed");
  L.emin_est = wr[0]; L.emax_est = wr[0];
  for (int i = 1; i < it; ++i) { if (wr[i] < L.emin_est) L.emin_est = wr[i]; if (wr[i] > L.emax_est) L.emax_est = wr[i]; }
  L.emin = s.esteig[0] * L.emin_est + s.esteig[1] * L.emax_est;
  L.emax = s.esteig[2] * L.emin_est + s.esteig[3] * L.emax_est;
  CUDA_OK(cudaStreamSynchronize(c->stream));
  for (auto &v : V) dev_free(c, v);   // the Arnoldi basis is set-up scratch (11 fine-level vectors: 4.5 GB at 128^3)
  return 0;
}

static int cheb_estimate(xsb_ctx c, Level &L);
// Coarsest level: dense inverse up to 6600 rows; above that (abf.opts' 3 levels at 32^3 and larger) the hierarchy is continued
// INTERNALLY by further Galerkin coarsening until the dense inverse applies, and the level is solved by coarse_pcg.
static int coarse_setup(xsb_ctx c)
{
  Level &L0 = c->lev[0]; const int bs = L0.A.bs;
  c->nsub = 0; c->coarse_its = c->coarse_solves = 0;
  const int64_t dense_max = c->opt.integer("xsb_coarse_dense_max", 6600) < 6600 ? c->opt.integer("xsb_coarse_dense_max", 6600) : 6600;   // tests lower it to exercise the internal hierarchy on small meshes
  if ((int64_t)L0.A.nb * bs <= dense_max) return coarse_invert(c, L0);
  int dims[XSB_MAX_LEVELS][3]; int ns = 1;
  dims[0][0] = L0.nx; dims[0][1] = L0.ny; dims[0][2] = L0.nz;
  while ((int64_t)dims[ns - 1][0] * dims[ns - 1][1] * dims[ns - 1][2] * bs > dense_max) {
    if (ns == XSB_MAX_LEVELS) return xsb_fail(c, XSB_ERR_SUP, "coarsest MG level: internal hierarchy deeper than %d levels", XSB_MAX_LEVELS);
    for (int d = 0; d < 3; ++d) {
      const int nn = dims[ns - 1][d];
      if (d == 2 && c->nsd == 2) { dims[ns][d] = 1; continue; }
      if ((nn - 1) % 2 != 0 || (nn - 1) / 2 + 1 < 2) { if ((int64_t)dims[ns - 1][0] * dims[ns - 1][1] * dims[ns - 1][2] * bs <= 6600) goto done; return xsb_fail(c, XSB_ERR_SUP, "coarsest MG level has %lld dofs and its lattice %dx%dx%d cannot be coarsened further for the internal coarse solver (dense inverse: <= 6600)",
                                                                     (long long)L0.A.nb * bs, L0.nx, L0.ny, L0.nz); }
      dims[ns][d] = (nn - 1) / 2 + 1;
    }
    ++ns;
  }
done:
  if (ns == 1) return coarse_invert(c, L0);
  c->nsub = ns;
  { Level &T = c->sub[ns - 1]; T = Level(); T.nx = L0.nx; T.ny = L0.ny; T.nz = L0.nz; T.A = L0.A; T.owns_A = false; }
  for (int l = ns - 2; l >= 0; --l) {
    Level &F = c->sub[l + 1], &C = c->sub[l]; C = Level();
    C.nx = dims[ns - 1 - l][0]; C.ny = dims[ns - 1 - l][1]; C.nz = dims[ns - 1 - l][2];
    XSB_CHK(galerkin(c, F, C));
  }
  for (int l = 0; l < ns; ++l) {
    Level &L = c->sub[l]; const int64_t n = (int64_t)L.A.nb * bs;
    XSB_CHK(dev_alloc(c, &L.x, (size_t)n)); XSB_CHK(dev_alloc(c, &L.b, (size_t)n)); XSB_CHK(dev_alloc(c, &L.r, (size_t)n));
    XSB_CHK(dev_alloc(c, &L.w0, (size_t)n)); XSB_CHK(dev_alloc(c, &L.w1, (size_t)n)); XSB_CHK(dev_alloc(c, &L.idiag, (size_t)n));
    XSB_CHK(baij_diag_inv(c, L.A, L.idiag));
  }
  XSB_CHK(coarse_invert(c, c->sub[0]));
  for (int l = 1; l < ns; ++l) XSB_CHK(cheb_estimate(c, c->sub[l]));
  const int64_t n0 = (int64_t)L0.A.nb * bs;
  XSB_CHK(dev_alloc(c, &c->cg_p, (size_t)n0)); XSB_CHK(dev_alloc(c, &c->cg_q, (size_t)n0));
  return 0;
}

// ------------------------------------------------------------------ PCSetUp_MG
int mg_setup(xsb_ctx c)
{
  const SolverOpts &s = c->so; const int levels = s.mg_levels; const Lattice &Lt = c->lat;
  if (levels < 1 || levels > XSB_MAX_LEVELS) return xsb_fail(c, XSB_ERR_ARG, "-saddle_fieldsplit_u_pc_mg_levels %d out of range", levels);
  c->nlev = levels;
  const Slab &S = c->slab; const bool dist = S.nranks > 1;
  if (c->no_A && levels < 2) return xsb_fail(c, XSB_ERR_SUP, "-xsb_matrix_free full needs at least 2 MG levels (a one-level PCMG is LU of the assembled A00)");
  if (dist && levels < 2) return xsb_fail(c, XSB_ERR_SUP, "the slab partition needs at least 2 MG levels (the coarse levels are replicated)");
  { Level &L = c->lev[levels - 1]; L = Level(); L.nx = Lt.NX; L.ny = Lt.NY; L.nz = Lt.NZ; L.A = c->A00; L.owns_A = false; L.dist = dist; }
  for (int l = levels - 2; l >= 0; --l) {
    Level &F = c->lev[l + 1], &C = c->lev[l]; C = Level();
    int dims[3];
    if (xsb_mg_level_dims(c->nsd, Lt.mx, Lt.my, S.mz_glob, levels, l, dims)) return xsb_fail(c, XSB_ERR_ARG, "mesh %dx%dx%d cannot be coarsened to %d MG levels (DMCoarsen needs (n-1) divisible by 2)", Lt.mx, Lt.my, S.mz_glob, levels);
    C.nx = dims[0]; C.ny = dims[1]; C.nz = dims[2];
    if (F.dist) XSB_CHK(galerkin_replicate(c, F, C));
    else if (c->no_A && l == levels - 2) XSB_CHK(galerkin_elements(c, C));
    else XSB_CHK(galerkin(c, F, C));
  }
  // Slabs: the coarse levels below the fine one stay distributed by node planes while they are large (-xsb_pdist_min_nodes) and
  // every rank keeps at least one plane; their operators are replicated (set-up), their vectors are current only around the
  // owned planes (Level::pdist).  Smaller levels are replicated and computed redundantly.
  if (dist && c->opt.integer("xsb_pdist", 1)) {
    const int64_t min_nodes = c->opt.integer("xsb_pdist_min_nodes", 100000);
    for (int l = levels - 2; l >= 1; --l) {
      Level &L = c->lev[l];
      if ((int64_t)L.A.nb < min_nodes) break;
      bool ok = true; int mine0 = 0, mine1 = 0;
      for (int r = 0; r < S.nranks; ++r) { int p0, p1; xsb_pdist_range(S.mz_glob, S.nranks, r, levels - 2 - l, &p0, &p1); if (p1 - p0 < 1) ok = false; if (r == S.rank) { mine0 = p0; mine1 = p1; } }
      if (!ok) break;
      L.pdist = true; L.rp0 = mine0; L.rp1 = mine1;
    }
    for (int l = levels - 2; l >= 1; --l) {   // a distributed level above a replicated one gathers the restricted planes
      Level &L = c->lev[l];
      if (!L.pdist || c->lev[l - 1].pdist) continue;
      L.cr0.resize(S.nranks); L.cr1.resize(S.nranks);
      for (int r = 0; r < S.nranks; ++r) xsb_pdist_range(S.mz_glob, S.nranks, r, levels - 2 - l + 1, &L.cr0[r], &L.cr1[r]);
    }
  }
  for (int l = 0; l < levels; ++l) {
    Level &L = c->lev[l]; const int64_t n = (int64_t)L.A.nb * L.A.bs;
    XSB_CHK(dev_alloc(c, &L.x, (size_t)n)); XSB_CHK(dev_alloc(c, &L.b, (size_t)n)); XSB_CHK(dev_alloc(c, &L.r, (size_t)n));
    XSB_CHK(dev_alloc(c, &L.w0, (size_t)n)); XSB_CHK(dev_alloc(c, &L.w1, (size_t)n)); XSB_CHK(dev_alloc(c, &L.idiag, (size_t)n));
    if (L.pdist) { CUDA_OK(cudaMemsetAsync(L.b, 0, sizeof(double) * n, c->stream)); CUDA_OK(cudaMemsetAsync(L.r, 0, sizeof(double) * n, c->stream)); CUDA_OK(cudaMemsetAsync(L.w0, 0, sizeof(double) * n, c->stream)); CUDA_OK(cudaMemsetAsync(L.w1, 0, sizeof(double) * n, c->stream)); }
    if (c->no_A && l == levels - 1) XSB_CHK(mf_diag_inv(c, L.idiag)); else XSB_CHK(baij_diag_inv(c, L.A, L.idiag));
  }
  XSB_CHK(coarse_setup(c));
  for (int l = 1; l < levels; ++l) {
    Level &L = c->lev[l];
    if (s.n_cheb_fixed > 0) {
      if (l - 1 >= s.n_cheb_fixed) return xsb_fail(c, XSB_ERR_ARG, "explicit Chebyshev eigenvalues missing for MG level %d", l);
      L.emin = s.cheb_emin[l - 1]; L.emax = s.cheb_emax[l - 1];
    } else XSB_CHK(cheb_estimate(c, L));
  }
  return 0;
}
