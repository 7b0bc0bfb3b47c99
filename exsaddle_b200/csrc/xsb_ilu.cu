// xsb_ilu.cu -- ILU(0) of the scaled pressure mass matrix and its triangular solves (K10 of SURVEY 2.1).
//
// Replaces -saddle_fieldsplit_p_pc_type bjacobi (one block per rank, sub-PC ILU(0) in natural ordering,
// abf.opts:15; PETSc MatLUFactorNumeric_SeqAIJ / MatSolve_SeqAIJ) applied to Mpscaled
// (MatAssemble_Schur, femixedspace.c:2837).  Mp is a 27-point (9-point in 2-D) stencil on the pressure
// lattice, so row (i,j,k) depends only on rows with a smaller wavefront number w = i + 2j + 4k: all rows
// of one wavefront are independent and the natural-ordering factorisation / substitutions are reproduced
// exactly (same operations per row, same order within a row) by sweeping wavefronts.  The matrix is tiny
// next to A00 (np = (m+1)^3), latency bound, and runs in one 1024-thread CTA with a block barrier per
// wavefront: no grid-wide synchronisation and no host round trips.
#include "xsb.h"

struct PLat { int px, py, pz; };

__device__ __forceinline__ bool wf_row(const PLat &P, int w, int t, int &i, int &j, int &k)
{
  // candidate t -> (j,k); i follows from the wavefront number
  j = t % P.py; k = t / P.py;
  if (k >= P.pz) return false;
  i = w - 2 * j - 4 * k;
  return i >= 0 && i < P.px;
}

// Factorisation: row-wise IKJ restricted to the pattern; multiplier = a_ik * (1/a_kk); diagonal stored inverted.
__global__ void __launch_bounds__(1024) k_ilu0_factor(PLat P, const int *__restrict__ ia, const double *__restrict__ a, double *lu, int *flag)
{
  const BoxPattern pat{P.px, P.py, P.pz, 0};
  const int nw = (P.px - 1) + 2 * (P.py - 1) + 4 * (P.pz - 1) + 1, ncand = P.py * P.pz;
  for (int w = 0; w < nw; ++w) {
    for (int t = threadIdx.x; t < ncand; t += blockDim.x) {
      int i, j, k; if (!wf_row(P, w, t, i, j, k)) continue;
      const int row = i + j * P.px + k * P.px * P.py;
      int l0, h0, l1, h1, l2, h2; range_pp(i, P.px, l0, h0); range_pp(j, P.py, l1, h1); range_pp(k, P.pz, l2, h2);
      const int nx = h0 - l0 + 1, ny = h1 - l1 + 1, nrow = nx * ny * (h2 - l2 + 1), r0 = ia[row];
      double wv[27];
      for (int s = 0; s < nrow; ++s) wv[s] = a[r0 + s];
      const int dslot = ((k - l2) * ny + (j - l1)) * nx + (i - l0);
      for (int s = 0; s < dslot; ++s) {   // lower entries in ascending column order
        const int ci = l0 + s % nx, cj = l1 + (s / nx) % ny, ck = l2 + s / (nx * ny);
        const int crow = ci + cj * P.px + ck * P.px * P.py;
        int m0, g0, m1, g1, m2, g2; range_pp(ci, P.px, m0, g0); range_pp(cj, P.py, m1, g1); range_pp(ck, P.pz, m2, g2);
        const int cnx = g0 - m0 + 1, cny = g1 - m1 + 1, cn = cnx * cny * (g2 - m2 + 1), c0 = ia[crow];
        const int cd = ((ck - m2) * cny + (cj - m1)) * cnx + (ci - m0);
        double mult = wv[s];
        if (mult != 0.0) {
          mult = mult * lu[c0 + cd];   // lu[diag] = 1/pivot
          wv[s] = mult;
          for (int u = cd + 1; u < cn; ++u) {   // U part of the pivot row
            const int gi = m0 + u % cnx, gj = m1 + (u / cnx) % cny, gk = m2 + u / (cnx * cny);
            const int sl = box_slot(pat, i, j, k, gi, gj, gk);
            if (sl >= 0) wv[sl] -= mult * lu[c0 + u];   // fill outside the pattern is dropped (ILU(0))
          }
        }
      }
      if (wv[dslot] == 0.0) *flag = 1;
      for (int s = 0; s < nrow; ++s) lu[r0 + s] = s == dslot ? 1.0 / wv[dslot] : wv[s];
    }
    __syncthreads();
  }
}

// x = U^-1 L^-1 b (MatSolve_SeqAIJ): forward with unit L, backward with U and the inverted diagonal
__global__ void __launch_bounds__(1024) k_ilu0_solve(PLat P, const int *__restrict__ ia, const double *__restrict__ lu, const double *__restrict__ b, double *x)
{
  const int nw = (P.px - 1) + 2 * (P.py - 1) + 4 * (P.pz - 1) + 1, ncand = P.py * P.pz;
  for (int w = 0; w < nw; ++w) {
    for (int t = threadIdx.x; t < ncand; t += blockDim.x) {
      int i, j, k; if (!wf_row(P, w, t, i, j, k)) continue;
      const int row = i + j * P.px + k * P.px * P.py;
      int l0, h0, l1, h1, l2, h2; range_pp(i, P.px, l0, h0); range_pp(j, P.py, l1, h1); range_pp(k, P.pz, l2, h2);
      const int nx = h0 - l0 + 1, ny = h1 - l1 + 1, r0 = ia[row];
      const int dslot = ((k - l2) * ny + (j - l1)) * nx + (i - l0);
      double s = b[row];
      for (int u = 0; u < dslot; ++u) {
        const int col = (l0 + u % nx) + (l1 + (u / nx) % ny) * P.px + (l2 + u / (nx * ny)) * P.px * P.py;
        s -= lu[r0 + u] * x[col];
      }
      x[row] = s;
    }
    __syncthreads();
  }
  for (int w = nw - 1; w >= 0; --w) {
    for (int t = threadIdx.x; t < ncand; t += blockDim.x) {
      int i, j, k; if (!wf_row(P, w, t, i, j, k)) continue;
      const int row = i + j * P.px + k * P.px * P.py;
      int l0, h0, l1, h1, l2, h2; range_pp(i, P.px, l0, h0); range_pp(j, P.py, l1, h1); range_pp(k, P.pz, l2, h2);
      const int nx = h0 - l0 + 1, ny = h1 - l1 + 1, nrow = nx * ny * (h2 - l2 + 1), r0 = ia[row];
      const int dslot = ((k - l2) * ny + (j - l1)) * nx + (i - l0);
      double s = x[row];
      for (int u = nrow - 1; u > dslot; --u) {
        const int col = (l0 + u % nx) + (l1 + (u / nx) % ny) * P.px + (l2 + u / (nx * ny)) * P.px * P.py;
        s -= lu[r0 + u] * x[col];
      }
      x[row] = s * lu[r0 + dslot];
    }
    __syncthreads();
  }
}

int ilu_setup(xsb_ctx c)
{
  const Lattice &L = c->lat; PLat P{L.PX, L.PY, L.PZ};
  XSB_CHK(dev_alloc(c, &c->mp_lu, (size_t)c->Mp.nnz));
  int *flag = nullptr; XSB_CHK(dev_alloc(c, &flag, 1));
  CUDA_OK(cudaMemsetAsync(flag, 0, sizeof(int), c->stream));
  k_ilu0_factor<<<1, 1024, 0, c->stream>>>(P, c->Mp.ia, c->Mp.a, c->mp_lu, flag); KERNEL_OK();
  int h = 0; CUDA_OK(cudaMemcpyAsync(&h, flag, sizeof(int), cudaMemcpyDeviceToHost, c->stream)); CUDA_OK(cudaStreamSynchronize(c->stream));
  if (h) return xsb_fail(c, XSB_ERR_BREAKDOWN, "zero pivot in ILU(0) of Mpscaled");
  return 0;
}

int ilu_apply(xsb_ctx c, const double *b, double *x)
{
  const Lattice &L = c->lat; PLat P{L.PX, L.PY, L.PZ};
  k_ilu0_solve<<<1, 1024, 0, c->stream>>>(P, c->Mp.ia, c->mp_lu, b, x); KERNEL_OK();
  return 0;
}
