// xsb_spmv.cu -- operator application kernels (K1/K2/K3 of SURVEY 2.1).
//
// Replace PETSc MatMult on the AIJ matrices the reference fills (KSPSolve, exSaddle.c:425; rhs_diri,
// femixedspace.c:2639).  Both kernels are HBM-bandwidth bound (2 flop per 12 B / 8.4 B of matrix):
//   spmv_csr   AIJ layout (the reference's MATAIJ): one warp per row, lanes stride the row so every
//              load instruction of values / column indices is a contiguous 256 B / 128 B segment;
//              matrix stream uses ld.global.cs (evict-first) so the 126 MB L2 stays available for x.
//   spmv_baij  BAIJ(bs) velocity block A00 and every Galerkin level: one warp per block row with fixed lane
//              roles (see the kernel).  The epilogue fuses the vector work that always follows the product in
//              the smoother (residual / Chebyshev update), so the smoother makes exactly one pass over A00
//              per iteration and no separate vector passes.
#include "xsb.h"

__device__ __forceinline__ double ld_stream(const double *p) { return __ldcs(p); }
__device__ __forceinline__ int ld_stream(const int *p) { return __ldcs(p); }

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------ K1: CSR
template <int UNROLL>
__global__ void __launch_bounds__(256) spmv_csr_kernel(int64_t row0, int64_t n, const int *__restrict__ ia, const int *__restrict__ ja,
                                                       const double *__restrict__ a, const double *__restrict__ x, double *y, const double *yadd /* may alias y */)
{
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t row = row0 + warp; row < row0 + n; row += nwarps) {
    const int k0 = ia[row], k1 = ia[row + 1];
    double acc = 0.0;
    int k = k0 + lane;
    for (; k + 32 * (UNROLL - 1) < k1; k += 32 * UNROLL) {
      double v[UNROLL]; int cidx[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) { v[u] = ld_stream(a + k + 32 * u); cidx[u] = ld_stream(ja + k + 32 * u); }
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) acc += v[u] * __ldg(x + cidx[u]);
    }
    for (; k < k1; k += 32) acc += ld_stream(a + k) * __ldg(x + ld_stream(ja + k));
    acc = warp_sum(acc);
    if (lane == 0) y[row] = yadd ? acc + yadd[row] : acc;
  }
}

// Rows of the full saddle operator that belong to one velocity node (its NSD components) have IDENTICAL column patterns
// (the pattern is the node's coupling box) -- PETSc exploits the same fact with its "inode" routines
// (MatMult_SeqAIJ_Inode).  One warp takes the BS rows of a node: every column index and every x value is loaded once and
// used BS times, so the product streams 8 + 4/BS bytes per nonzero of the AIJ arrays instead of 12 and gathers x a third
// as often.  Per row the lane-strided partial sums and the warp tree are those of spmv_csr_kernel: results are bitwise equal.
template <int BS, int UNROLL>
__global__ void __launch_bounds__(256) spmv_csr_inode_kernel(int64_t node0, int64_t nnodes, const int *__restrict__ ia, const int *__restrict__ ja,
                                                             const double *__restrict__ a, const double *__restrict__ x, double *y, const double *yadd /* may alias y */)
{
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t node = node0 + warp; node < node0 + nnodes; node += nwarps) {
    const int64_t r0 = BS * node;
    const int k0 = ia[r0], len = ia[r0 + 1] - k0;
    double acc[BS];
#pragma unroll
    for (int r = 0; r < BS; ++r) acc[r] = 0.0;
    int k = lane;
    for (; k + 32 * (UNROLL - 1) < len; k += 32 * UNROLL) {
      double v[BS][UNROLL]; int cidx[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        cidx[u] = ld_stream(ja + k0 + k + 32 * u);
#pragma unroll
        for (int r = 0; r < BS; ++r) v[r][u] = ld_stream(a + k0 + (int64_t)r * len + k + 32 * u);
      }
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const double xv = __ldg(x + cidx[u]);
#pragma unroll
        for (int r = 0; r < BS; ++r) acc[r] += v[r][u] * xv;
      }
    }
    for (; k < len; k += 32) {
      const double xv = __ldg(x + ld_stream(ja + k0 + k));
#pragma unroll
      for (int r = 0; r < BS; ++r) acc[r] += ld_stream(a + k0 + (int64_t)r * len + k) * xv;
    }
#pragma unroll
    for (int r = 0; r < BS; ++r) acc[r] = warp_sum(acc[r]);
    if (lane == 0) {
#pragma unroll
      for (int r = 0; r < BS; ++r) y[r0 + r] = yadd ? acc[r] + yadd[r0 + r] : acc[r];
    }
  }
}

// Short rows (A01: 8..27 entries per velocity row): LPR lanes per row, 32 / LPR rows per warp, so a load instruction
// still covers 32 consecutive entries of consecutive rows instead of one partly filled row.
template <int LPR>
__global__ void __launch_bounds__(256) spmv_csr_short_kernel(int64_t row0, int64_t n, const int *__restrict__ ia, const int *__restrict__ ja,
                                                             const double *__restrict__ a, const double *__restrict__ x, double *y, const double *yadd /* may alias y */)
{
  constexpr int GPW = 32 / LPR;   // rows per warp
  const int sub = threadIdx.x & (LPR - 1), gw = (threadIdx.x & 31) / LPR;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t rb = warp * GPW; rb < n; rb += nwarps * GPW) {   // trip count is uniform per warp: the shuffles below are full-mask
    const int64_t r = rb + gw, row = row0 + r; const bool ok = r < n;
    const int k0 = ok ? ia[row] : 0, k1 = ok ? ia[row + 1] : 0;
    double acc = 0.0;
    for (int k = k0 + sub; k < k1; k += LPR) acc += ld_stream(a + k) * __ldg(x + ld_stream(ja + k));
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o, LPR);
    if (ok && sub == 0) y[row] = yadd ? acc + yadd[row] : acc;
  }
}

int spmv_csr(xsb_ctx c, const Csr &A, const double *x, double *y, int64_t row0, int64_t nrows, const double *yadd)
{
  if (nrows < 0) nrows = A.n - row0;
  if (nrows <= 0) return XSB_OK;
  if (A.nnz < 40 * (int64_t)A.n) {   // short rows
    const int tpb = 256; int64_t blocks = (nrows * 8 + tpb - 1) / tpb;
    const int64_t cap = 148LL * 8 * 32; if (blocks > cap) blocks = cap;
    spmv_csr_short_kernel<8><<<(unsigned)blocks, tpb, 0, c->stream>>>(row0, nrows, A.ia, A.ja, A.a, x, y, yadd); KERNEL_OK();
    return XSB_OK;
  }
  const int tpb = 256; const int64_t cap = 148LL * 8 * 16;   // persistent-style grid: 148 SMs x 8 resident CTAs x 16 waves
  // node-grouped rows first (velocity rows of the full operator), single rows after
  const int bs = A.inode_bs;
  if (bs > 1 && row0 < A.inode_rows && row0 % bs == 0) {
    int64_t gend = row0 + nrows < A.inode_rows ? row0 + nrows : A.inode_rows; gend -= (gend - row0) % bs;
    const int64_t nnodes = (gend - row0) / bs;
    if (nnodes > 0) {
      int64_t blocks = (nnodes * 32 + tpb - 1) / tpb; if (blocks > cap) blocks = cap;
      if (bs == 3) spmv_csr_inode_kernel<3, 4><<<(unsigned)blocks, tpb, 0, c->stream>>>(row0 / bs, nnodes, A.ia, A.ja, A.a, x, y, yadd);
      else spmv_csr_inode_kernel<2, 4><<<(unsigned)blocks, tpb, 0, c->stream>>>(row0 / bs, nnodes, A.ia, A.ja, A.a, x, y, yadd);
      KERNEL_OK();
      nrows -= gend - row0; row0 = gend;
      if (nrows <= 0) return XSB_OK;
    }
  }
  int64_t blocks = (nrows * 32 + tpb - 1) / tpb;
  if (blocks > cap) blocks = cap;
  spmv_csr_kernel<8><<<(unsigned)blocks, tpb, 0, c->stream>>>(row0, nrows, A.ia, A.ja, A.a, x, y, yadd); KERNEL_OK();
  return XSB_OK;
}

// ------------------------------------------------------------------ K2: BAIJ with fixed lane roles, fused epilogue
__device__ __forceinline__ double epilogue_value(const Epilogue &ep, int64_t i, double ax)
{
  switch (ep.mode) {
  case EPI_RESIDUAL:   return ep.b[i] - ax;                                                             // r = b - A x
  case EPI_CHEB_FIRST: return ep.pk[i] + ep.s0 * (ep.idiag[i] * (ep.b[i] - ax));                        // p1 = x + scale*B(b - A x)
  case EPI_CHEB:       return ep.s0 * ep.pkm1[i] + ep.s1 * ep.pk[i] + ep.s2 * (ep.idiag[i] * (ep.b[i] - ax)); // VecAXPBYPCZ
  default:             return ax;
  }
}

// One warp per block row.  Each lane keeps a FIXED role (block g of the group, row comp ra, col comp ca) for the
// whole kernel: a warp iteration consumes NBW = 32/BS^2 whole blocks (27 of 32 lanes for BS = 3, all 32 for BS = 2)
// with one coalesced value load, one block-column load, one x gather and one FMA per lane -- no per-element
// div/mod, a single accumulator.  UN iterations are issued back to back so UN loads per lane are in flight.
template <int BS, int UN>
__global__ void __launch_bounds__(256) spmv_baij_kernel(int node0, int nb, const int *__restrict__ ia, const int *__restrict__ ja,
                                                        const double *__restrict__ a, const double *__restrict__ x, double *__restrict__ y, Epilogue ep)
{
  constexpr int BS2 = BS * BS, NBW = 32 / BS2, ACTIVE = NBW * BS2;
  const int lane = threadIdx.x & 31;
  const int g = lane / BS2, r = lane - g * BS2, ra = r / BS, ca = r - ra * BS;
  const bool active = lane < ACTIVE;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t node = node0 + warp; node < node0 + nb; node += nwarps) {
    const int b0 = ia[node], nblk = ia[node + 1] - b0;
    const double *__restrict__ av = a + (int64_t)b0 * BS2 + lane;
    const int *__restrict__ cj = ja + b0 + g;
    const int mine = active ? (nblk - g + NBW - 1) / NBW : 0;   // iterations in which this lane's block exists
    double acc = 0.0;
    for (int it = 0; it < mine; it += UN) {
      double v[UN]; int col[UN];
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        const bool ok = it + u < mine;
        v[u] = ok ? ld_stream(av + (int64_t)(it + u) * ACTIVE) : 0.0;
        col[u] = ok ? __ldg(cj + (it + u) * NBW) : 0;
      }
#pragma unroll
      for (int u = 0; u < UN; ++u) acc += v[u] * __ldg(x + (int64_t)BS * col[u] + ca);
    }
    // sum over the column component (lanes r, r+1, ..), then over the NBW block slots
    double t = acc;
#pragma unroll
    for (int s = 1; s < BS; ++s) t += __shfl_down_sync(0xffffffffu, acc, s);
    double tot = t;
#pragma unroll
    for (int s = 1; s < NBW; ++s) tot += __shfl_down_sync(0xffffffffu, t, s * BS2);
    if (g == 0 && ca == 0 && active) { const int64_t i = (int64_t)BS * node + ra; y[i] = epilogue_value(ep, i, tot); }
  }
}

// The same product for TWO row ranges of equal length in one launch (blockIdx.y picks the range): the two boundary planes of a
// plane-distributed coarse level, whose results the neighbours wait for (xsb_mg.cu).  Kept as a separate kernel on purpose.
template <int BS, int UN>
__global__ void __launch_bounds__(256) spmv_baij_pair_kernel(int nodeA, int nodeB, int nb, const int *__restrict__ ia, const int *__restrict__ ja,
                                                             const double *__restrict__ a, const double *__restrict__ x, double *__restrict__ y, Epilogue ep)
{
  constexpr int BS2 = BS * BS, NBW = 32 / BS2, ACTIVE = NBW * BS2;
  const int lane = threadIdx.x & 31;
  const int g = lane / BS2, r = lane - g * BS2, ra = r / BS, ca = r - ra * BS;
  const bool active = lane < ACTIVE;
  const int node0 = blockIdx.y ? nodeB : nodeA;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t node = node0 + warp; node < node0 + nb; node += nwarps) {
    const int b0 = ia[node], nblk = ia[node + 1] - b0;
    const double *__restrict__ av = a + (int64_t)b0 * BS2 + lane;
    const int *__restrict__ cj = ja + b0 + g;
    const int mine = active ? (nblk - g + NBW - 1) / NBW : 0;
    double acc = 0.0;
    for (int it = 0; it < mine; it += UN) {
      double v[UN]; int col[UN];
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        const bool ok = it + u < mine;
        v[u] = ok ? ld_stream(av + (int64_t)(it + u) * ACTIVE) : 0.0;
        col[u] = ok ? __ldg(cj + (it + u) * NBW) : 0;
      }
#pragma unroll
      for (int u = 0; u < UN; ++u) acc += v[u] * __ldg(x + (int64_t)BS * col[u] + ca);
    }
    double t = acc;
#pragma unroll
    for (int s = 1; s < BS; ++s) t += __shfl_down_sync(0xffffffffu, acc, s);
    double tot = t;
#pragma unroll
    for (int s = 1; s < NBW; ++s) tot += __shfl_down_sync(0xffffffffu, t, s * BS2);
    if (g == 0 && ca == 0 && active) { const int64_t i = (int64_t)BS * node + ra; y[i] = epilogue_value(ep, i, tot); }
  }
}
int spmv_baij_pair(xsb_ctx c, const Baij &A, const double *x, double *y, const Epilogue &ep, int nodeA, int nodeB, int nnodes)
{
  if (nnodes <= 0) return XSB_OK;
  if (A.bs != 3) { XSB_CHK(spmv_baij(c, A, x, y, ep, nodeA, nnodes)); return spmv_baij(c, A, x, y, ep, nodeB, nnodes); }
  const int tpb = 256; dim3 grid((unsigned)(((int64_t)nnodes * 32 + tpb - 1) / tpb), 2);
  spmv_baij_pair_kernel<3, 8><<<grid, tpb, 0, c->stream>>>(nodeA, nodeB, nnodes, A.ia, A.ja, A.a, x, y, ep);
  KERNEL_OK();
  return XSB_OK;
}

int spmv_baij(xsb_ctx c, const Baij &A, const double *x, double *y, const Epilogue &ep, int node0, int nnodes)
{
  if (nnodes < 0) nnodes = A.nb - node0;
  if (nnodes <= 0) return XSB_OK;
  const int tpb = 256; int64_t blocks = ((int64_t)nnodes * 32 + tpb - 1) / tpb;
  const int64_t cap = 148LL * 8 * 16;
  if (blocks > cap) blocks = cap;
  if (A.bs == 3) spmv_baij_kernel<3, 8><<<(unsigned)blocks, tpb, 0, c->stream>>>(node0, nnodes, A.ia, A.ja, A.a, x, y, ep);
  else if (A.bs == 2) spmv_baij_kernel<2, 4><<<(unsigned)blocks, tpb, 0, c->stream>>>(node0, nnodes, A.ia, A.ja, A.a, x, y, ep);
  else return xsb_fail(c, XSB_ERR_SUP, "BAIJ block size %d", A.bs);
  KERNEL_OK();
  return XSB_OK;
}

// Fine-level A00 launch with bookkeeping: per-mode launch counters (for the algorithmic byte count) and, with
// -xsb_time_kernels, CUDA events recorded on the launching stream in front of the ghost exchange, in front of the kernel and
// behind it (prof_mark).  Events are only read back after the solve (spmv_collect_timing), so timing adds no synchronisation.
int prof_mark(xsb_ctx c, int cat)
{
  if (!c->so.time_kernels) return XSB_OK;
  if (c->ev_used + 1 > c->evpool.size()) { for (int i = 0; i < 1024; ++i) { cudaEvent_t e; CUDA_OK(cudaEventCreate(&e)); c->evpool.push_back(e); } }
  CUDA_OK(cudaEventRecord(c->evpool[c->ev_used], c->stream));
  if (c->ev_cat.size() <= c->ev_used) c->ev_cat.resize(c->ev_used + 1024);
  c->ev_cat[c->ev_used++] = cat;
  return XSB_OK;
}
int spmv_a00_fine(xsb_ctx c, const Baij &A, const double *x, double *y, const Epilogue &ep)
{
  c->n_a00++; c->a00_mode[ep.mode & 3]++;
  XSB_CHK(prof_mark(c, PROF_FINE_HALO));
  XSB_CHK(comm_halo_u(c, const_cast<double *>(x)));   // ghost planes of the input (no-op on one GPU)
  XSB_CHK(prof_mark(c, PROF_FINE));
  if (c->so.matrix_free) XSB_CHK(mf_a00_apply(c, x, y, ep));
  else { const int pn = c->lat.NX * c->lat.NY; XSB_CHK(spmv_baij(c, A, x, y, ep, c->slab.ou0 * pn, (c->slab.ou1 - c->slab.ou0) * pn)); }
  return prof_mark(c, PROF_OTHER);
}
int spmv_collect_timing(xsb_ctx c)
{
  for (int i = 0; i < PROF_N; ++i) { c->prof_ms[i] = 0; c->prof_cnt[i] = 0; }
  for (size_t i = 0; i + 1 < c->ev_used; ++i) {
    float ms = 0; CUDA_OK(cudaEventElapsedTime(&ms, c->evpool[i], c->evpool[i + 1]));
    const int cat = c->ev_cat[i]; c->prof_ms[cat] += ms; c->prof_cnt[cat]++;
    if (cat == PROF_FINE) { c->a00_ns_sum += 1e6 * (double)ms; c->a00_timed++; }
  }
  c->ev_used = 0;
  return XSB_OK;
}
