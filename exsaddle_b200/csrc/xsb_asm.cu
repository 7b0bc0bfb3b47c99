// xsb_asm.cu -- additive Schwarz on the reference's element-patch subdomains.
//
// Replaces `-saddle_pc_type asm -saddle_pc_asm_dm_subdomains -set_ksp_dm [-dmdafe_overlap k]` with exact sub-solves
// (`-saddle_sub_pc_type lu`; Makefile:297, 410) and the same PC as the smoother of the monolithic -mg path (Makefile:417):
//   * subdomains: DMCreateDomainDecomposition_DMDAFEQ2Q1 (femixedspace.c:823-837) gives every MPI rank ONE subdomain, the
//     closed box of the Q2 elements it owns (fitted to its DMDA node range by parity, :1074-1124) grown by -dmdafe_overlap
//     element layers (:745-816).  There is no MPI here: `-xsb_ranks N` names the communicator size the reference would run
//     on, and the process grid / ownership PETSc's DMDA would choose for it are computed in closed form (xsb_dmda_grid,
//     xsb_asm_subdomain: host integer logic, exported for the CPU tests);
//   * PCASM semantics (PC_ASM_RESTRICT with DM-defined subdomains, never grown by -pc_asm_overlap): the residual is restricted
//     to the patch, solved exactly, and the solution is written back only on the dofs the rank owns.  Owned dofs tile the
//     vector, so z needs neither zeroing nor atomics;
//   * sub-solves: each patch matrix (symmetric indefinite: the p-p block is zero) is inverted densely on the device by
//     Gauss-Jordan with partial pivoting (the -mg coarse solver's kernels), and the PC apply is ONE launch: a warp per owned
//     row forms  z[g(row)] = sum_j Inv_s[row][j] r[g(j)]  -- only the owned rows of every inverse are ever read.
#include "xsb.h"

static inline unsigned nblk(int64_t n, int bs = 256) { return (unsigned)((n + bs - 1) / bs); }

extern "C" {
// DMDACreate{2,3}d with PETSC_DECIDE: the "squarish" process grid of PETSc's DMSetUp_DA_2D / _3D for M x N (x P) nodes on `size` ranks
int xsb_dmda_grid(int nsd, int M, int N, int P, int size, int out[3])
{
  if (!out || size < 1 || M < 1 || N < 1 || (nsd == 3 && P < 1)) return XSB_ERR_ARG;
  int m, n, p = 1;
  if (nsd == 2) {
    m = (int)(0.5 + sqrt((double)M * (double)size / (double)N)); if (!m) m = 1;
    n = 1;
    while (m > 0) { n = size / m; if (m * n == size) break; m--; }
    if (M > N && m < n) { int t = m; m = n; n = t; }
  } else {
    n = (int)(0.5 + pow((double)N * N * (double)size / ((double)P * M), 1.0 / 3.0)); if (!n) n = 1;
    while (n > 0) { int pm = size / n; if (n * pm == size) break; n--; }
    if (!n) n = 1;
    m = (int)(0.5 + sqrt((double)M * (double)size / ((double)P * n))); if (!m) m = 1;
    while (m > 0) { p = size / (m * n); if (m * n * p == size) break; m--; }
    if (M > P && m < p) { int t = m; m = p; p = t; }
  }
  if (m < 1 || n < 1 || p < 1 || m * n * p != size) return XSB_ERR_ARG;
  if (M < m || N < n || (nsd == 3 && P < p)) return XSB_ERR_ARG;   // PETSc: "Partition in x direction is too fine!"
  out[0] = m; out[1] = n; out[2] = p;
  return XSB_OK;
}

// Subdomain of `rank` (x fastest in the process grid): out[0..2] first element, [3..5] one past the last element of the patch
// (after -dmdafe_overlap), [6..8] / [9..11] owned velocity-node range [lo,hi), [12..14] / [15..17] owned pressure-node range.
int xsb_asm_subdomain(int nsd, int mx, int my, int mz, int size, int overlap, int rank, int out[18])
{
  if (!out || overlap < 0 || rank < 0 || rank >= size) return XSB_ERR_ARG;
  const int mesh[3] = {mx, my, nsd == 3 ? mz : 1};
  int grid[3]; if (xsb_dmda_grid(nsd, 2 * mx + 1, 2 * my + 1, nsd == 3 ? 2 * mz + 1 : 1, size, grid)) return XSB_ERR_ARG;
  const int pidx[3] = {rank % grid[0], (rank / grid[0]) % grid[1], rank / (grid[0] * grid[1])};
  for (int d = 0; d < 3; ++d) { out[d] = 0; out[3 + d] = 1; out[6 + d] = 0; out[9 + d] = 1; out[12 + d] = 0; out[15 + d] = 1; }
  for (int d = 0; d < nsd; ++d) {
    const int M = 2 * mesh[d] + 1, g = grid[d];
    int s = 0, pstart = 0, e0 = 0, ne = 0;
    for (int r = 0; r < g; ++r) {   // node ownership M/g (+1 on the low ranks); whole elements fitted by parity (femixedspace.c:1074-1124)
      const int w = M / g + ((M % g) > r ? 1 : 0), e = s + w;
      const int s_el = s % 2 == 0 ? s : s - 1, e_el = e % 2 == 0 ? e : e - 1;
      if ((e_el - s_el) % 2 || e_el <= s_el) return XSB_ERR_ARG;   // "Cannot generate consistent macro element"
      const int nel = (e_el - s_el) / 2, npts = nel + (r == g - 1 ? 1 : 0);   // pressure nodes per rank (:1216-1236)
      if (r == pidx[d]) { e0 = s_el / 2; ne = nel; out[6 + d] = s; out[9 + d] = e; out[12 + d] = pstart; out[15 + d] = pstart + npts; }
      s = e; pstart += npts;
    }
    out[d] = e0 - overlap < 0 ? 0 : e0 - overlap;
    out[3 + d] = e0 + ne + overlap > mesh[d] ? mesh[d] : e0 + ne + overlap;
  }
  return XSB_OK;
}
}   // extern "C"

struct AsmPC {
  int nsub = 0; int64_t n = 0;
  int *idx = nullptr;          // patch dofs of all subdomains, concatenated (ascending inside a patch)
  double *inv = nullptr;       // dense inverses, concatenated
  int *row_sub = nullptr, *row_loc = nullptr;   // for each of the n owned rows: its subdomain and its row inside the patch
  int *sub_off = nullptr, *sub_n = nullptr; int64_t *inv_off = nullptr;   // per subdomain: offset into idx, size, offset into inv
};

__global__ void k_asm_g2l(int ns, const int *__restrict__ idx, int *__restrict__ g2l, int val)
{ const int t = blockIdx.x * blockDim.x + threadIdx.x; if (t < ns) g2l[idx[t]] = val < 0 ? -1 : t; }
__global__ void k_asm_extract(int ns, const int *__restrict__ idx, const int *__restrict__ g2l, const int *__restrict__ ia, const int *__restrict__ ja, const double *__restrict__ a, double *__restrict__ M)
{
  const int r = blockIdx.x * blockDim.x + threadIdx.x; if (r >= ns) return;
  const int g = idx[r];
  for (int k = ia[g]; k < ia[g + 1]; ++k) { const int lc = g2l[ja[k]]; if (lc >= 0) M[(int64_t)r * ns + lc] = a[k]; }
}
__global__ void __launch_bounds__(256) k_asm_apply(int64_t n, const int *__restrict__ row_sub, const int *__restrict__ row_loc, const int *__restrict__ sub_off, const int *__restrict__ sub_n,
                                                   const int64_t *__restrict__ inv_off, const int *__restrict__ idx, const double *__restrict__ inv, const double *__restrict__ r, double *__restrict__ z)
{
  const int64_t q = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; const int lane = threadIdx.x & 31;
  if (q >= n) return;
  const int s = row_sub[q], lr = row_loc[q], ns = sub_n[s]; const int *id = idx + sub_off[s];
  const double *row = inv + inv_off[s] + (int64_t)lr * ns;
  double acc = 0.0;
  for (int j = lane; j < ns; j += 32) acc += row[j] * r[id[j]];
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) z[id[lr]] = acc;
}

int dense_invert_pivoted(xsb_ctx c, int n, double *M, double *Inv);   // xsb_mmg.cu: Gauss-Jordan with partial pivoting, M is destroyed

void asm_free(void *p) { delete (AsmPC *)p; }

// alloc: the handle whose solver phase owns the allocations; P: the problem (lattice + assembled A) the PC is built for
int asm_setup(xsb_ctx alloc, xsb_ctx P, int size, int overlap, void **out)
{
  xsb_ctx c = alloc; const Lattice &L = P->lat; const int nsd = L.nsd; cudaStream_t st = c->stream;
  if (P->slab.nranks > 1 || alloc->slab.nranks > 1) return xsb_fail(c, XSB_ERR_SUP, "ASM element patches are implemented for one GPU");
  if (!P->A.a) return xsb_fail(c, XSB_ERR_SUP, "ASM needs the assembled operator");
  AsmPC *S = new AsmPC(); *out = S; S->nsub = size; S->n = L.n;
  std::vector<int> hidx, hoff(size + 1, 0), hn(size), hsub(L.n, -1), hloc(L.n, -1); std::vector<int64_t> hinv(size + 1, 0);
  const int N[3] = {L.NX, L.NY, L.NZ}, Pn[3] = {L.PX, L.PY, L.PZ};
  for (int r = 0; r < size; ++r) {
    int b[18];
    if (xsb_asm_subdomain(nsd, L.mx, L.my, L.mz, size, overlap, r, b)) { delete S; *out = nullptr; return xsb_fail(c, XSB_ERR_ARG, "-xsb_ranks %d: PETSc's DMDA cannot partition the %d x %d x %d velocity lattice into whole Q2 elements (Cannot generate consistent macro element)", size, L.NX, L.NY, L.NZ); }
    const int base = (int)hidx.size();
    auto owned_u = [&](int i, int j, int k) { return i >= b[6] && i < b[9] && j >= b[7] && j < b[10] && (nsd == 2 || (k >= b[8] && k < b[11])); };
    auto owned_p = [&](int i, int j, int k) { return i >= b[12] && i < b[15] && j >= b[13] && j < b[16] && (nsd == 2 || (k >= b[14] && k < b[17])); };
    const int k0 = nsd == 3 ? 2 * b[2] : 0, k1 = nsd == 3 ? 2 * b[5] : 0;
    for (int k = k0; k <= k1; ++k) for (int j = 2 * b[1]; j <= 2 * b[4]; ++j) for (int i = 2 * b[0]; i <= 2 * b[3]; ++i) {
      const int64_t nd = i + (int64_t)j * N[0] + (int64_t)k * N[0] * N[1];
      for (int d = 0; d < nsd; ++d) { const int g = (int)(nd * nsd + d); if (owned_u(i, j, k)) { hsub[g] = r; hloc[g] = (int)hidx.size() - base; } hidx.push_back(g); }
    }
    const int pk0 = nsd == 3 ? b[2] : 0, pk1 = nsd == 3 ? b[5] : 0;
    for (int k = pk0; k <= pk1; ++k) for (int j = b[1]; j <= b[4]; ++j) for (int i = b[0]; i <= b[3]; ++i) {
      const int g = (int)(L.nu + i + (int64_t)j * Pn[0] + (int64_t)k * Pn[0] * Pn[1]);
      if (owned_p(i, j, k)) { hsub[g] = r; hloc[g] = (int)hidx.size() - base; }
      hidx.push_back(g);
    }
    hn[r] = (int)hidx.size() - base; hoff[r + 1] = (int)hidx.size(); hinv[r + 1] = hinv[r] + (int64_t)hn[r] * hn[r];
    if (hn[r] > 6600) { delete S; *out = nullptr; return xsb_fail(c, XSB_ERR_SUP, "ASM subdomain of rank %d has %d unknowns; the dense sub-solve supports <= 6600 (use more -xsb_ranks)", r, hn[r]); }
  }
  for (int64_t g = 0; g < L.n; ++g) if (hsub[g] < 0) { delete S; *out = nullptr; return xsb_fail(c, XSB_ERR_ARG, "ASM: dof %lld is owned by no subdomain", (long long)g); }
  XSB_CHK(dev_alloc(c, &S->idx, hidx.size())); XSB_CHK(dev_alloc(c, &S->inv, (size_t)hinv[size]));
  XSB_CHK(dev_alloc(c, &S->row_sub, (size_t)L.n)); XSB_CHK(dev_alloc(c, &S->row_loc, (size_t)L.n));
  XSB_CHK(dev_alloc(c, &S->sub_off, (size_t)size + 1)); XSB_CHK(dev_alloc(c, &S->sub_n, (size_t)size)); XSB_CHK(dev_alloc(c, (char **)&S->inv_off, sizeof(int64_t) * ((size_t)size + 1)));
  CUDA_OK(cudaMemcpyAsync(S->idx, hidx.data(), sizeof(int) * hidx.size(), cudaMemcpyHostToDevice, st));
  CUDA_OK(cudaMemcpyAsync(S->row_sub, hsub.data(), sizeof(int) * L.n, cudaMemcpyHostToDevice, st));
  CUDA_OK(cudaMemcpyAsync(S->row_loc, hloc.data(), sizeof(int) * L.n, cudaMemcpyHostToDevice, st));
  CUDA_OK(cudaMemcpyAsync(S->sub_off, hoff.data(), sizeof(int) * (size + 1), cudaMemcpyHostToDevice, st));
  CUDA_OK(cudaMemcpyAsync(S->sub_n, hn.data(), sizeof(int) * size, cudaMemcpyHostToDevice, st));
  CUDA_OK(cudaMemcpyAsync(S->inv_off, hinv.data(), sizeof(int64_t) * (size + 1), cudaMemcpyHostToDevice, st));
  CUDA_OK(cudaStreamSynchronize(st));   // the host vectors go out of scope
  int nmax = 0; for (int r = 0; r < size; ++r) if (hn[r] > nmax) nmax = hn[r];
  double *M = nullptr; int *g2l = nullptr;
  CUDA_OK(cudaMalloc(&M, sizeof(double) * (size_t)nmax * nmax)); CUDA_OK(cudaMalloc(&g2l, sizeof(int) * L.n));
  CUDA_OK(cudaMemsetAsync(g2l, 0xff, sizeof(int) * L.n, st));
  int rc = 0;
  for (int r = 0; r < size && !rc; ++r) {
    const int ns = hn[r]; const int *id = S->idx + hoff[r];
    CUDA_OK(cudaMemsetAsync(M, 0, sizeof(double) * (size_t)ns * ns, st));
    k_asm_g2l<<<nblk(ns), 256, 0, st>>>(ns, id, g2l, 1); KERNEL_OK();
    k_asm_extract<<<nblk(ns, 128), 128, 0, st>>>(ns, id, g2l, P->A.ia, P->A.ja, P->A.a, M); KERNEL_OK();
    k_asm_g2l<<<nblk(ns), 256, 0, st>>>(ns, id, g2l, -1); KERNEL_OK();
    rc = dense_invert_pivoted(c, ns, M, S->inv + hinv[r]);
  }
  cudaStreamSynchronize(st); cudaFree(M); cudaFree(g2l);
  return rc;
}

int asm_apply(xsb_ctx c, void *p, const double *r, double *z)
{
  const AsmPC *S = (const AsmPC *)p;
  k_asm_apply<<<nblk(S->n * 32), 256, 0, c->stream>>>(S->n, S->row_sub, S->row_loc, S->sub_off, S->sub_n, S->inv_off, S->idx, S->inv, r, z); KERNEL_OK();
  return 0;
}
