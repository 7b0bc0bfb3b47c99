// xsb_ilu.cu -- ILU(0) of the scaled pressure mass matrix and its triangular solves (K10 of SURVEY 2.1).
//
// Replaces -saddle_fieldsplit_p_pc_type bjacobi (one block per rank, sub-PC ILU(0) in natural ordering,
// abf.opts:15; PETSc MatLUFactorNumeric_SeqAIJ / MatSolve_SeqAIJ) applied to Mpscaled
// (MatAssemble_Schur, femixedspace.c:2837).  Mp is a 27-point (9-point in 2-D) stencil on the pressure
// lattice, so row (i,j,k) depends only on rows with a smaller wavefront number w = i + 2j + 4k: all rows
// of one wavefront are independent and the natural-ordering factorisation / substitutions are reproduced
// exactly (same operations per row, same order within a row) by sweeping wavefronts.
//
// The solve is latency bound (449 dependent wavefronts per sweep at 64^3, ~600 rows each, 13 FMAs per row).
// It runs as ONE thread-block cluster of 8 CTAs x 1024 threads: rows of a wavefront are spread over the
// cluster, wavefronts are separated by the hardware cluster barrier (barrier.cluster, ~0.3 us) instead of
// a kernel launch or a grid-wide barrier, and both sweeps are one launch.  The factors are repacked at set-up
// into wavefront order (SoA per wavefront) so every access is coalesced, independent of x and prefetchable;
// x is exchanged through L2 (ld.global.cg / st.global.cg).
#include "xsb.h"
#include <cooperative_groups.h>
#include <cub/cub.cuh>
namespace cg = cooperative_groups;

struct PLat { int px, py, pz; };
#define ILU_CLUSTER 8
#define ILU_TPB 128

__device__ __forceinline__ bool wf_row(const PLat &P, int w, int t, int &i, int &j, int &k)
{
  j = t % P.py; k = t / P.py;
  if (k >= P.pz) return false;
  i = w - 2 * j - 4 * k;
  return i >= 0 && i < P.px;
}

// Factorisation (set-up, once): row-wise IKJ restricted to the pattern; multiplier = a_ik * (1/a_kk); the
// diagonal is stored inverted.  One CTA, block barrier per wavefront.
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

// ---- level schedule (set-up): rows sorted by wavefront number
__global__ void k_lvl_count(PLat P, int *cnt)
{
  int row = blockIdx.x * blockDim.x + threadIdx.x; if (row >= P.px * P.py * P.pz) return;
  const int i = row % P.px, j = (row / P.px) % P.py, k = row / (P.px * P.py);
  atomicAdd(&cnt[i + 2 * j + 4 * k], 1);
}
__global__ void k_lvl_fill(PLat P, const int *off, int *cursor, int *rows, int *diag, const int *ia)
{
  int row = blockIdx.x * blockDim.x + threadIdx.x; if (row >= P.px * P.py * P.pz) return;
  const int i = row % P.px, j = (row / P.px) % P.py, k = row / (P.px * P.py), w = i + 2 * j + 4 * k;
  rows[off[w] + atomicAdd(&cursor[w], 1)] = row;   // order inside a wavefront is irrelevant: its rows are independent
  int l0, h0, l1, h1, l2, h2; range_pp(i, P.px, l0, h0); range_pp(j, P.py, l1, h1); range_pp(k, P.pz, l2, h2);
  diag[row] = ia[row] + ((k - l2) * (h1 - l1 + 1) + (j - l1)) * (h0 - l0 + 1) + (i - l0);
}

// Level-packed factors (set-up): for wavefront w with cnt rows starting at off[w], entry u of local row l lives at
// 13*off[w] + u*cnt + l (SoA per wavefront => coalesced, and affine in the wavefront order => prefetchable).
// Lower entries ascending, upper entries DEscending (the order MatSolve consumes them); unused slots have n < u.
__global__ void k_ilu_pack(int nw, const int *__restrict__ off, const int *__restrict__ rows, const int *__restrict__ ia, const int *__restrict__ ja,
                           const int *__restrict__ diag, const double *__restrict__ lu, const int *__restrict__ lvl_of_q,
                           int *fcol, double *fval, unsigned char *fn, int *bcol, double *bval, unsigned char *bn, double *binv)
{
  const int q = blockIdx.x * blockDim.x + threadIdx.x; if (q >= off[nw]) return;
  const int w = lvl_of_q[q], cnt = off[w + 1] - off[w], l = q - off[w];
  const int64_t base = (int64_t)13 * off[w] + l;
  const int row = rows[q], r0 = ia[row], r1 = ia[row + 1], d = diag[row];
  const int nl = d - r0, nu = r1 - 1 - d;
  fn[q] = (unsigned char)nl; bn[q] = (unsigned char)nu; binv[q] = lu[d];
  for (int u = 0; u < 13; ++u) {
    fcol[base + (int64_t)u * cnt] = u < nl ? ja[r0 + u] : row; fval[base + (int64_t)u * cnt] = u < nl ? lu[r0 + u] : 0.0;
    bcol[base + (int64_t)u * cnt] = u < nu ? ja[r1 - 1 - u] : row; bval[base + (int64_t)u * cnt] = u < nu ? lu[r1 - 1 - u] : 0.0;
  }
}
__global__ void k_lvl_of_q(int nw, const int *__restrict__ off, int *lvl_of_q)
{
  const int w = blockIdx.x; if (w >= nw) return;
  for (int q = off[w] + threadIdx.x; q < off[w + 1]; q += blockDim.x) lvl_of_q[q] = w;
}

__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// x = U^-1 L^-1 b (MatSolve_SeqAIJ): forward with unit L, backward with U and the inverted diagonal.
// One cluster of 8 CTAs with few warps each (the cluster barrier's acquire invalidates L1 once per warp, so its
// cost grows with the warp count: 6.3 ms/apply with 32 warps per CTA, 3.2 ms with 4).  Software pipelining: the
// factors of the NEXT wavefront (independent of x) are loaded into registers before the barrier of the current
// one, so after a barrier only the x values written by earlier wavefronts (L2 hits) are on the critical path.
struct IluRow { double lv[13]; int cv[13]; int row, n; double aux; };

template <bool FWD>
__device__ __forceinline__ void ilu_load(IluRow &R, int o, int cnt, int l, const int *__restrict__ rows, const int *__restrict__ col,
                                         const double *__restrict__ val, const unsigned char *__restrict__ nn, const double *__restrict__ aux)
{
  const int q = o + l; const int64_t base = (int64_t)13 * o + l;
  R.row = rows[q]; R.n = nn[q];
#pragma unroll
  for (int u = 0; u < 13; ++u) { R.lv[u] = val[base + (int64_t)u * cnt]; R.cv[u] = col[base + (int64_t)u * cnt]; }
  R.aux = FWD ? aux[R.row] : aux[q];   // forward: b[row]; backward: 1/pivot
}
template <bool FWD>
__device__ __forceinline__ void ilu_row(const IluRow &R, double *x)
{
  double xv[13];
#pragma unroll
  for (int u = 0; u < 13; ++u) xv[u] = u < R.n ? __ldcg(x + R.cv[u]) : 0.0;
  double s = FWD ? R.aux : __ldcg(x + R.row);
#pragma unroll
  for (int u = 0; u < 13; ++u) if (u < R.n) s -= R.lv[u] * xv[u];   // column order of the sequential sweep (packed that way)
  __stcg(x + R.row, FWD ? s : s * R.aux);
}

__global__ void __cluster_dims__(ILU_CLUSTER, 1, 1) __launch_bounds__(ILU_TPB)
k_ilu0_solve(int nw, int np, const int *__restrict__ off, const int *__restrict__ rows,
             const int *__restrict__ fcol, const double *__restrict__ fval, const unsigned char *__restrict__ fn,
             const int *__restrict__ bcol, const double *__restrict__ bval, const unsigned char *__restrict__ bn, const double *__restrict__ binv,
             const double *__restrict__ b, double *x)
{
  extern __shared__ int soff[];   // nw + 1 wavefront offsets
  cg::cluster_group cl = cg::this_cluster();
  const int gt = (int)cl.block_rank() * ILU_TPB + threadIdx.x, gn = ILU_CLUSTER * ILU_TPB;
  for (int i = threadIdx.x; i <= nw; i += ILU_TPB) soff[i] = off[i];
  // L2 prefetch of everything the two sweeps stream (128-byte lines), issued in sweep order
  {
    const int64_t nv = (int64_t)13 * np;
    for (int64_t i = (int64_t)gt * 16; i < nv; i += (int64_t)gn * 16) prefetch_l2(fval + i);
    for (int64_t i = (int64_t)gt * 32; i < nv; i += (int64_t)gn * 32) prefetch_l2(fcol + i);
    for (int64_t i = (int64_t)gt * 16; i < np; i += (int64_t)gn * 16) prefetch_l2(b + i);
    for (int64_t i = (int64_t)gt * 32; i < np; i += (int64_t)gn * 32) prefetch_l2(rows + i);
    for (int64_t i = nv - 16 - (int64_t)gt * 16; i >= 0; i -= (int64_t)gn * 16) prefetch_l2(bval + i);
    for (int64_t i = nv - 32 - (int64_t)gt * 32; i >= 0; i -= (int64_t)gn * 32) prefetch_l2(bcol + i);
  }
  __syncthreads();
  IluRow cur, nxt;
  bool have = gt < soff[1] - soff[0];
  if (have) ilu_load<true>(cur, soff[0], soff[1] - soff[0], gt, rows, fcol, fval, fn, b);
  for (int w = 0; w < nw; ++w) {
    const int o = soff[w], cnt = soff[w + 1] - o;
    bool have_next = false;
    if (w + 1 < nw) { const int o1 = soff[w + 1], c1 = soff[w + 2] - o1; have_next = gt < c1; if (have_next) ilu_load<true>(nxt, o1, c1, gt, rows, fcol, fval, fn, b); }
    if (have) ilu_row<true>(cur, x);
    for (int l = gt + gn; l < cnt; l += gn) { IluRow t; ilu_load<true>(t, o, cnt, l, rows, fcol, fval, fn, b); ilu_row<true>(t, x); }
    cl.sync();
    cur = nxt; have = have_next;
  }
  have = gt < soff[nw] - soff[nw - 1];
  if (have) ilu_load<false>(cur, soff[nw - 1], soff[nw] - soff[nw - 1], gt, rows, bcol, bval, bn, binv);
  for (int w = nw - 1; w >= 0; --w) {
    const int o = soff[w], cnt = soff[w + 1] - o;
    bool have_next = false;
    if (w > 0) { const int o1 = soff[w - 1], c1 = soff[w] - o1; have_next = gt < c1; if (have_next) ilu_load<false>(nxt, o1, c1, gt, rows, bcol, bval, bn, binv); }
    if (have) ilu_row<false>(cur, x);
    for (int l = gt + gn; l < cnt; l += gn) { IluRow t; ilu_load<false>(t, o, cnt, l, rows, bcol, bval, bn, binv); ilu_row<false>(t, x); }
    cl.sync();
    cur = nxt; have = have_next;
  }
}

// bjacobi block of this rank: rows and columns of the owned pressure planes [op0,op1) of the local lattice, as a
// 27-point matrix on the owned sub-lattice (PETSc PCBJACOBI: the rank's diagonal block of Mpscaled).
__global__ void k_own_len(PLat P, int *len)
{
  int row = blockIdx.x * blockDim.x + threadIdx.x; if (row >= P.px * P.py * P.pz) return;
  const BoxPattern pat{P.px, P.py, P.pz, 0};
  len[row] = box_size(pat, row % P.px, (row / P.px) % P.py, row / (P.px * P.py));
}
__global__ void k_own_fill(PLat P, int op0, int pz_loc, const int *__restrict__ sia, const double *__restrict__ sa, const int *__restrict__ dia, int *dja, double *da)
{
  int row = blockIdx.x * blockDim.x + threadIdx.x; if (row >= P.px * P.py * P.pz) return;
  const int i = row % P.px, j = (row / P.px) % P.py, k = row / (P.px * P.py), ks = k + op0;
  int l0, h0, l1, h1, l2, h2, s2, t2;
  range_pp(i, P.px, l0, h0); range_pp(j, P.py, l1, h1); range_pp(k, P.pz, l2, h2); range_pp(ks, pz_loc, s2, t2);
  const int nx = h0 - l0 + 1, ny = h1 - l1 + 1, srow = i + j * P.px + ks * P.px * P.py;
  int d = dia[row];
  for (int kk = l2; kk <= h2; ++kk) for (int jj = l1; jj <= h1; ++jj) for (int ii = l0; ii <= h0; ++ii) {
    dja[d] = ii + jj * P.px + kk * P.px * P.py;
    da[d] = sa[sia[srow] + ((kk + op0 - s2) * ny + (jj - l1)) * nx + (ii - l0)];
    ++d;
  }
}
static int build_owned_block(xsb_ctx c, Csr &B)
{
  const Lattice &L = c->lat; const Slab &S = c->slab; cudaStream_t st = c->stream;
  PLat P{L.PX, L.PY, S.op1 - S.op0}; const int np = P.px * P.py * P.pz;
  int *len = nullptr; XSB_CHK(dev_alloc(c, &len, (size_t)np + 1)); XSB_CHK(dev_alloc(c, &B.ia, (size_t)np + 1));
  k_own_len<<<(np + 255) / 256, 256, 0, st>>>(P, len); KERNEL_OK();
  void *tmp = nullptr; size_t tb = 0;
  CUDA_OK(cub::DeviceScan::ExclusiveSum(nullptr, tb, len, B.ia, np + 1, st));
  CUDA_OK(cudaMalloc(&tmp, tb));
  CUDA_OK(cub::DeviceScan::ExclusiveSum(tmp, tb, len, B.ia, np + 1, st));
  int tot = 0; CUDA_OK(cudaMemcpyAsync(&tot, B.ia + np, sizeof(int), cudaMemcpyDeviceToHost, st));
  CUDA_OK(cudaStreamSynchronize(st)); CUDA_OK(cudaFree(tmp));
  B.n = B.m = np; B.nnz = tot;
  XSB_CHK(dev_alloc(c, &B.ja, (size_t)tot)); XSB_CHK(dev_alloc(c, &B.a, (size_t)tot));
  k_own_fill<<<(np + 255) / 256, 256, 0, st>>>(P, S.op0, L.PZ, c->Mp.ia, c->Mp.a, B.ia, B.ja, B.a); KERNEL_OK();
  return 0;
}

int ilu_setup(xsb_ctx c)
{
  const Lattice &L = c->lat; cudaStream_t st = c->stream;
  if (c->slab.nranks > 1) XSB_CHK(build_owned_block(c, c->MpOwn)); else c->MpOwn = c->Mp;
  const Csr &M = c->MpOwn;
  PLat P{L.PX, L.PY, c->slab.op1 - c->slab.op0};
  const int np = P.px * P.py * P.pz, nw = (P.px - 1) + 2 * (P.py - 1) + 4 * (P.pz - 1) + 1;
  XSB_CHK(dev_alloc(c, &c->mp_lu, (size_t)M.nnz));
  int *flag = nullptr; XSB_CHK(dev_alloc(c, &flag, 1));
  CUDA_OK(cudaMemsetAsync(flag, 0, sizeof(int), st));
  k_ilu0_factor<<<1, 1024, 0, st>>>(P, M.ia, M.a, c->mp_lu, flag); KERNEL_OK();
  // level schedule
  int *cnt = nullptr, *cursor = nullptr, *diag = nullptr;
  XSB_CHK(dev_alloc(c, &cnt, (size_t)nw + 1)); XSB_CHK(dev_alloc(c, &cursor, (size_t)nw + 1));
  XSB_CHK(dev_alloc(c, &c->ilu_lvl_off, (size_t)nw + 1)); XSB_CHK(dev_alloc(c, &c->ilu_rows, (size_t)np)); XSB_CHK(dev_alloc(c, &diag, (size_t)np));
  CUDA_OK(cudaMemsetAsync(cnt, 0, sizeof(int) * (nw + 1), st)); CUDA_OK(cudaMemsetAsync(cursor, 0, sizeof(int) * (nw + 1), st));
  k_lvl_count<<<(np + 255) / 256, 256, 0, st>>>(P, cnt); KERNEL_OK();
  { void *tmp = nullptr; size_t tb = 0;
    CUDA_OK(cub::DeviceScan::ExclusiveSum(nullptr, tb, cnt, c->ilu_lvl_off, nw + 1, st));
    CUDA_OK(cudaMalloc(&tmp, tb));
    CUDA_OK(cub::DeviceScan::ExclusiveSum(tmp, tb, cnt, c->ilu_lvl_off, nw + 1, st));
    CUDA_OK(cudaStreamSynchronize(st)); CUDA_OK(cudaFree(tmp)); }
  k_lvl_fill<<<(np + 255) / 256, 256, 0, st>>>(P, c->ilu_lvl_off, cursor, c->ilu_rows, diag, M.ia); KERNEL_OK();
  c->ilu_nlvl = nw; c->ilu_diag = diag;
  {
    int *lvl_of_q = nullptr; XSB_CHK(dev_alloc(c, &lvl_of_q, (size_t)np));
    XSB_CHK(dev_alloc(c, &c->ilu_fcol, (size_t)13 * np + 64)); XSB_CHK(dev_alloc(c, &c->ilu_fval, (size_t)13 * np + 64)); XSB_CHK(dev_alloc(c, &c->ilu_fn, (size_t)np));
    XSB_CHK(dev_alloc(c, &c->ilu_bcol, (size_t)13 * np + 64)); XSB_CHK(dev_alloc(c, &c->ilu_bval, (size_t)13 * np + 64)); XSB_CHK(dev_alloc(c, &c->ilu_bn, (size_t)np));
    XSB_CHK(dev_alloc(c, &c->ilu_binv, (size_t)np));
    k_lvl_of_q<<<nw, 256, 0, st>>>(nw, c->ilu_lvl_off, lvl_of_q); KERNEL_OK();
    k_ilu_pack<<<(np + 255) / 256, 256, 0, st>>>(nw, c->ilu_lvl_off, c->ilu_rows, M.ia, M.ja, diag, c->mp_lu, lvl_of_q,
                                                   c->ilu_fcol, c->ilu_fval, c->ilu_fn, c->ilu_bcol, c->ilu_bval, c->ilu_bn, c->ilu_binv); KERNEL_OK();
  }
  int h = 0; CUDA_OK(cudaMemcpyAsync(&h, flag, sizeof(int), cudaMemcpyDeviceToHost, st)); CUDA_OK(cudaStreamSynchronize(st));
  if (h) return xsb_fail(c, XSB_ERR_BREAKDOWN, "zero pivot in ILU(0) of Mpscaled");
  return 0;
}

int ilu_apply(xsb_ctx c, const double *b, double *x)
{
  k_ilu0_solve<<<ILU_CLUSTER, ILU_TPB, sizeof(int) * (c->ilu_nlvl + 1), c->stream>>>(c->ilu_nlvl, c->MpOwn.n, c->ilu_lvl_off, c->ilu_rows, c->ilu_fcol, c->ilu_fval, c->ilu_fn,
                                                       c->ilu_bcol, c->ilu_bval, c->ilu_bn, c->ilu_binv, b, x); KERNEL_OK();
  return 0;
}
