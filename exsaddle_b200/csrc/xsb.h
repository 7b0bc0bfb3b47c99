// xsb.h -- internal structures of the B200 exSaddle library (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <map>
#include <set>
#include <string>
#include <vector>
#include "../../include/exsaddle_b200.h"

#define XSB_MAX_LEVELS 10   // MG_DEPTH, exSaddle.h:24
#define XSB_NSLOT 6         // coefficient slots at a quadrature point
enum { C_ETA = 0, C_FU0 = 1, C_FU1 = 2, C_FU2 = 3, C_FP = 4, C_LAM = 5 };
enum { BC_SOLCX = 0, BC_FIXEDBASE, BC_COMPRESSION, BC_COMPRESSION2, BC_MMS1 };

#ifdef __CUDACC__
#define HD __host__ __device__ __forceinline__
#else
#define HD inline
#endif

// ---------------------------------------------------------------------------------------------------------
// Lattice description shared by host and device code.
struct Lattice {
  int nsd;
  int mx, my, mz;          // elements of the LOCAL lattice (mz = 1 in 2-D)
  int NX, NY, NZ;          // velocity nodes
  int PX, PY, PZ;          // pressure nodes
  int64_t nun, npn, nu, np, n, nel;
  double hu[3];            // velocity node spacing (from the GLOBAL mesh)
  int zoff;                // first element layer of the local lattice in the global mesh (0 on one GPU)
};

// z-slab partition: this rank owns element layers [k0,k1) of mz_glob and works on the local lattice
// of layers [e0,e1) = [k0-2, k1+1) clipped to the mesh, so that every matrix row of an owned node, and every
// row one node plane below (needed by the Galerkin product), is complete without communication.
struct Slab {
  int rank = 0, nranks = 1, mz_glob = 0;
  int k0 = 0, k1 = 0, e0 = 0, e1 = 0;
  int ou0 = 0, ou1 = 0;    // owned velocity-node planes, local indices [ou0,ou1)
  int op0 = 0, op1 = 0;    // owned pressure-node planes
};
struct Ranges { int64_t off0 = 0, len0 = 0, off1 = 0, len1 = 0; };   // owned index ranges of a ghosted vector

// Coupling ranges along one direction. i = node coordinate, N = nodes in that direction.
// velocity node -> velocity nodes: +-2 from an element-corner (even) node, +-1 from a mid (odd) node.
HD void range_uu(int i, int N, int &lo, int &hi) { if (i & 1) { lo = i - 1; hi = i + 1; } else { lo = i - 2 < 0 ? 0 : i - 2; hi = i + 2 > N - 1 ? N - 1 : i + 2; } }
// velocity node -> pressure nodes
HD void range_up(int i, int P, int &lo, int &hi) { if (i & 1) { lo = (i - 1) >> 1; hi = (i + 1) >> 1; } else { int p = i >> 1; lo = p - 1 < 0 ? 0 : p - 1; hi = p + 1 > P - 1 ? P - 1 : p + 1; } }
// pressure node -> velocity nodes
HD void range_pu(int p, int N, int &lo, int &hi) { lo = 2 * p - 2 < 0 ? 0 : 2 * p - 2; hi = 2 * p + 2 > N - 1 ? N - 1 : 2 * p + 2; }
// pressure node -> pressure nodes (also every coarse MG level: 27-point box)
HD void range_pp(int p, int P, int &lo, int &hi) { lo = p - 1 < 0 ? 0 : p - 1; hi = p + 1 > P - 1 ? P - 1 : p + 1; }

// A "box pattern" block matrix on a node lattice: block row of node (i,j,k) couples to the nodes of the box
// [lo,hi] per direction; q2 = 1 uses range_uu (finest velocity level), q2 = 0 uses range_pp (27-point).
struct BoxPattern { int nx, ny, nz, q2; };
HD void box_range(const BoxPattern &p, int c, int n, int &lo, int &hi) { if (p.q2) range_uu(c, n, lo, hi); else range_pp(c, n, lo, hi); }

// Column boxes of one AIJ row (velocity-node row or pressure-node row): velocity-node box and pressure-node box.
struct RowBox { int ulo[3], uhi[3], plo[3], phi[3]; int ncu, ncp; };
HD void row_box_u(const Lattice &L, int i, int j, int k, RowBox &b)
{
  range_uu(i, L.NX, b.ulo[0], b.uhi[0]); range_uu(j, L.NY, b.ulo[1], b.uhi[1]); range_uu(k, L.NZ, b.ulo[2], b.uhi[2]);
  range_up(i, L.PX, b.plo[0], b.phi[0]); range_up(j, L.PY, b.plo[1], b.phi[1]); range_up(k, L.PZ, b.plo[2], b.phi[2]);
  b.ncu = (b.uhi[0] - b.ulo[0] + 1) * (b.uhi[1] - b.ulo[1] + 1) * (b.uhi[2] - b.ulo[2] + 1);
  b.ncp = (b.phi[0] - b.plo[0] + 1) * (b.phi[1] - b.plo[1] + 1) * (b.phi[2] - b.plo[2] + 1);
}
HD void row_box_p(const Lattice &L, int i, int j, int k, RowBox &b)
{
  range_pu(i, L.NX, b.ulo[0], b.uhi[0]); range_pu(j, L.NY, b.ulo[1], b.uhi[1]); range_pu(k, L.NZ, b.ulo[2], b.uhi[2]);
  range_pp(i, L.PX, b.plo[0], b.phi[0]); range_pp(j, L.PY, b.plo[1], b.phi[1]); range_pp(k, L.PZ, b.plo[2], b.phi[2]);
  b.ncu = (b.uhi[0] - b.ulo[0] + 1) * (b.uhi[1] - b.ulo[1] + 1) * (b.uhi[2] - b.ulo[2] + 1);
  b.ncp = (b.phi[0] - b.plo[0] + 1) * (b.phi[1] - b.plo[1] + 1) * (b.phi[2] - b.plo[2] + 1);
}
// slot of velocity node (ii,jj,kk) / pressure node inside a row's box (ascending node index order)
HD int box_upos(const RowBox &b, int ii, int jj, int kk) { return ((kk - b.ulo[2]) * (b.uhi[1] - b.ulo[1] + 1) + (jj - b.ulo[1])) * (b.uhi[0] - b.ulo[0] + 1) + (ii - b.ulo[0]); }
HD int box_ppos(const RowBox &b, int ii, int jj, int kk) { return ((kk - b.plo[2]) * (b.phi[1] - b.plo[1] + 1) + (jj - b.plo[1])) * (b.phi[0] - b.plo[0] + 1) + (ii - b.plo[0]); }
// AIJ row -> box; returns 1 for a velocity row (comp in *a) and 0 for a pressure row
HD int row_to_box(const Lattice &L, int64_t row, RowBox &b, int *comp)
{
  if (row < L.nu) { int64_t nd = row / L.nsd; *comp = (int)(row - nd * L.nsd); int i = (int)(nd % L.NX), j = (int)((nd / L.NX) % L.NY), k = (int)(nd / ((int64_t)L.NX * L.NY)); row_box_u(L, i, j, k, b); return 1; }
  int64_t nd = row - L.nu; *comp = 0; int i = (int)(nd % L.PX), j = (int)((nd / L.PX) % L.PY), k = (int)(nd / ((int64_t)L.PX * L.PY)); row_box_p(L, i, j, k, b); return 0;
}
// block-row box of a BoxPattern lattice matrix
HD int box_size(const BoxPattern &p, int i, int j, int k)
{ int l0, h0, l1, h1, l2, h2; box_range(p, i, p.nx, l0, h0); box_range(p, j, p.ny, l1, h1); box_range(p, k, p.nz, l2, h2); return (h0 - l0 + 1) * (h1 - l1 + 1) * (h2 - l2 + 1); }
// slot of node (gi,gj,gk) in the block row of node (i,j,k), or -1 when outside the box
HD int box_slot(const BoxPattern &p, int i, int j, int k, int gi, int gj, int gk)
{
  int l0, h0, l1, h1, l2, h2; box_range(p, i, p.nx, l0, h0); box_range(p, j, p.ny, l1, h1); box_range(p, k, p.nz, l2, h2);
  if (gi < l0 || gi > h0 || gj < l1 || gj > h1 || gk < l2 || gk > h2) return -1;
  return ((gk - l2) * (h1 - l1 + 1) + (gj - l1)) * (h0 - l0 + 1) + (gi - l0);
}

// FE tables of the uniform box mesh (host_tables in xsb_fe.cu), kept on the device for the element kernels
struct FeTables {
  double wq[27];            // product Gauss weights (literals of femixedspace.c:1379-1380)
  double Nu[27][27];        // Q2 basis at quadrature points [q][i]
  double Gu[27][27][3];     // Q2 global derivatives d/dx_d  [q][i][d] = dN/dxi_d / h_d
  double Np[27][8];         // Q1 basis [q][i]
  double detJ;              // h_x h_y (h_z)
};

// 1-D Q2 basis / derivative tables of the element kernels: N[q][n], Dd[q][n] = D[q][n] / h_d (uniform mesh: J = diag(h))
struct MfTabS { double N[3][3], Dx[3][3], Dy[3][3], Dz[3][3], w[3]; };

struct Csr  { int n = 0, m = 0; int64_t nnz = 0; int *ia = nullptr, *ja = nullptr; double *a = nullptr;
              int inode_bs = 0; int64_t inode_rows = 0; };   // rows [0, inode_rows) come in groups of inode_bs consecutive rows with identical column patterns (the components of a velocity node)
// BAIJ: block rows = lattice nodes, blocks bs x bs stored row-major, block columns ascending.
struct Baij { int nb = 0, bs = 0; int64_t nblk = 0; int *ia = nullptr, *ja = nullptr; double *a = nullptr; BoxPattern pat{0, 0, 0, 0}; };

struct Level {
  int nx = 0, ny = 0, nz = 0;
  Baij A; bool owns_A = false;
  double *idiag = nullptr;
  double emin = 0, emax = 0, emin_est = NAN, emax_est = NAN;
  double *x = nullptr, *b = nullptr, *r = nullptr, *w0 = nullptr, *w1 = nullptr, *w2 = nullptr;
  double *inv = nullptr;      // dense inverse of the coarsest operator
  bool pdist = false; int rp0 = 0, rp1 = 0;     // plane-distributed coarse level: vectors keep their GLOBAL indexing (and full length), but only the node planes [rp0-1, rp1+1) are kept current; this rank computes the rows of [rp0,rp1) and exchanges one plane with each neighbour
  std::vector<int> cr0, cr1;                    // pdist level above a replicated one: the coarse planes each rank restricts
  bool dist = false;          // z-slab distributed level (vectors on the local lattice, ghost planes, owned-row products)
};

struct Options {
  std::map<std::string, std::string> kv;
  std::set<std::string> used;
  bool has(const std::string &k) { if (kv.count(k)) { used.insert(k); return true; } return false; }
  std::string str(const std::string &k, const std::string &d) { if (has(k)) return kv[k]; return d; }
  double real(const std::string &k, double d) { if (has(k)) return atof(kv[k].c_str()); return d; }
  int integer(const std::string &k, int d) { if (has(k)) return atoi(kv[k].c_str()); return d; }
  bool flag(const std::string &k) { if (!has(k)) return false; const std::string &v = kv[k]; return v.empty() || v == "1" || v == "true" || v == "yes"; }
};

struct Model {   // resolved model parameters (models.c static option blocks)
  int model = -1, bc_type = BC_SOLCX;
  double c0 = 1, c1 = 1, lam0 = 1, lam1 = 1, rad = 0, cx = 0.5, cy = 0.5, cz = 0.5, xc = 0.5, size_x = 1;
  int nsink = 3, nz = 1, freeslip = 0;
};

struct SolverOpts {
  int ksp_type = 0;      // 0 gmres 1 fgmres
  int pc_type = 0;       // 0 none 1 jacobi 2 fieldsplit (abf) 3 monolithic PCMG (-mg) 4 fieldsplit with PETSc's default sub-solvers (plain -fs) 5 ASM on element patches
  int right = 0;
  double rtol = 1e-5, atol = 1e-50, dtol = 1e4; int max_it = 10000, restart = 30;
  double u_rtol = 1e-5; int u_max_it = 10000, u_restart = 30;
  int mg_levels = 1, cheb_its = 2; double esteig[4] = {0, 0.1, 0, 1.1}; int esteig_steps = 10, noise = 0;
  int n_cheb_fixed = 0; double cheb_emin[XSB_MAX_LEVELS], cheb_emax[XSB_MAX_LEVELS];
  int p_pc = 0;          // 0 ilu0 (bjacobi) 1 jacobi
  int time_kernels = 0;
  int mf_tile = 0;       // -xsb_mf_tile: tile variant of the one-pass element kernel (0: 16 x 5 elements, one CTA per SM; 1: 8 x 5, two CTAs per SM)
  int mf_kernel = 4;     // -xsb_mf_kernel: 4 = one-pass TMA-staged kernel (xsb_mf1p.cu), 3 = 8-colour kernel (xsb_mf.cu)
  int mf_reverse = 1;    // -xsb_mf_reverse: successive colour launches sweep the mesh in alternating directions (L2 reuse)
  int mf_chunk = 0;      // -xsb_mf_chunk: element layers per z-chunk of the matrix-free apply (0 = sized for L2)
  int mf_grad = 0;       // -xsb_mf_grad: A01 / A10 products by the closed-form gradient / divergence stencils (xsb_grad.cu); default on in operator-free mode
  int matrix_free = 0;   // -xsb_matrix_free: fine-level A00 products by the sum-factorised element kernel (xsb_mf.cu)
};

// -xsb_time_kernels: categories the solve's time line is cut into (prof_mark, xsb_spmv.cu)
enum { PROF_OTHER = 0, PROF_FINE_HALO, PROF_FINE, PROF_LVL, PROF_CHALO = PROF_LVL + XSB_MAX_LEVELS, PROF_XFER, PROF_ILU, PROF_FULL, PROF_CSR, PROF_N };

struct xsb_ctx_s {
  int nsd = 3, lame = 0, device = 0;
  bool have_device = false;
  cudaStream_t stream = nullptr;
  std::string err, banner;
  Options opt;
  Lattice lat{};
  Model mdl;
  SolverOpts so;
  bool assembled = false, ksp_ready = false;
  bool no_A = false;   // -xsb_matrix_free full: A and A00 are never stored (operator-free fine level)
  // FE data (device)
  double *coeff = nullptr;       // [slot][nel*nqp]
  double *coeff_nodal = nullptr; // [slot][npn]
  int nbc = 0; int *bc_idx = nullptr; double *bc_val = nullptr; unsigned char *isbc = nullptr;
  Csr A, A01, A10, A11, Mp, MpOwn;   // MpOwn: this rank's diagonal block of Mp (bjacobi); = Mp on one GPU
  Baij A00;
  double *F = nullptr;
  double *idiagA = nullptr;
  // solver state
  int nlev = 0; Level lev[XSB_MAX_LEVELS];
  int nsub = 0; Level sub[XSB_MAX_LEVELS];   // internal hierarchy below a coarsest level too large for the dense inverse (sub[nsub-1] aliases lev[0])
  double *cg_p = nullptr, *cg_q = nullptr; int coarse_its = 0, coarse_solves = 0;
  // V-cycle as a CUDA graph (-xsb_graph, default on): one captured graph per buffer-rotation state of the smoothers
  struct VGraph { const double *b = nullptr; double *x = nullptr, *w0 = nullptr, *w1 = nullptr; cudaGraphExec_t exec = nullptr; cudaGraph_t graph = nullptr;
                  int64_t d_a00 = 0, d_launch = 0, d_mode[4] = {0, 0, 0, 0}; double *px[XSB_MAX_LEVELS], *pw0[XSB_MAX_LEVELS], *pw1[XSB_MAX_LEVELS]; };
  std::vector<VGraph> vgraphs; int use_graph = 1; int64_t graph_replays = 0;
  double *mp_block_a = nullptr;   // values of Mpscaled with the entries between different ranks' dofs zeroed (plain -fs tree with -xsb_ranks > 1)
  double *mp_lu = nullptr, *mp_idiag = nullptr; int *ilu_rows = nullptr, *ilu_lvl_off = nullptr, *ilu_diag = nullptr; int ilu_nlvl = 0;
  int *ilu_fcol = nullptr, *ilu_bcol = nullptr; double *ilu_fval = nullptr, *ilu_bval = nullptr, *ilu_binv = nullptr; unsigned char *ilu_fn = nullptr, *ilu_bn = nullptr;
  std::vector<int> ilu_lvl_off_h;
  bool ilup_on = false; double *ilup_packf = nullptr, *ilup_packb = nullptr; unsigned *ilup_prog = nullptr; double *ilup_y = nullptr; int ilup_dims[6] = {0, 0, 0, 0, 0, 0}; size_t ilup_smem = 0;   // line-pipelined ILU solve (xsb_ilu.cu)
  std::vector<double *> V, Z, GV, GS;   // outer Krylov basis, GCR bases
  double *w_t1 = nullptr, *w_t2 = nullptr, *gcr_r = nullptr, *fs_tu = nullptr, *xdev = nullptr, *bdev = nullptr, *mf_tmp = nullptr; unsigned char *mf_bcnode = nullptr, *mf_bczero = nullptr; bool mf_opts_read = false;
  double *mf_part = nullptr; int *mf_zitems = nullptr; int mf_nz = 0, mf_zkey[3] = {-1, -1, -1}, mf_sms = 0, mf_ready = 0;   // one-pass element kernel: partial sums of shared nodes, z-boundary list
  double *red = nullptr;      // device reduction scratch
  double *red_h = nullptr;    // pinned host mirror
  double *scal = nullptr;     // device scalars (dot results consumed by kernels)
  // results
  int its = 0, reason = 0; std::vector<double> hist; std::vector<int> inner_its, inner_reason;
  float setup_ms = 0, solve_ms = 0;
  int64_t n_a00 = 0, n_a = 0, n_launch = 0, solve_launches = 0; double a00_ns_sum = 0; int64_t a00_timed = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, evk0 = nullptr, evk1 = nullptr;
  std::vector<cudaEvent_t> evpool; size_t ev_used = 0;   // -xsb_time_kernels: one event per mark, the stretch up to the next mark is booked on the mark's category
  std::vector<int> ev_cat; double prof_ms[PROF_N] = {0}; int64_t prof_cnt[PROF_N] = {0};
  int64_t a00_mode[4] = {0, 0, 0, 0};                    // fine-level A00 launches per epilogue mode
  Slab slab; void *nccl = nullptr;          // ncclComm_t when nranks > 1
  cudaStream_t side = nullptr; cudaEvent_t ev_fork = nullptr, ev_join = nullptr;   // slabs: the plane exchange of a distributed coarse level runs beside the interior rows of the product
  void *p2p = nullptr;                      // peer-memory halo windows (xsb_comm.cu), survives xsb_reset like the communicator
  Ranges own_full, own_u, own_p;            // owned entries of [u|p], u and p vectors on the local lattice
  std::vector<void *> allocs;   // every device allocation, for xsb_reset
  std::vector<char> alloc_phase; int phase = 0;   // 0: xsb_assemble, 1: xsb_ksp_setup (freed when the solver is set up again), 2: lazily created element-kernel state
  void *fe_tables = nullptr;    // FeTables on the device
  void *mmg = nullptr;          // monolithic -mg hierarchy (xsb_mmg.cu)
  void *asmpc = nullptr;        // -saddle_pc_type asm (xsb_asm.cu)
  void *grad_tab = nullptr; int grad_key[3] = {0, 0, 0};   // per-direction coefficient tables of the gradient / divergence stencils (xsb_grad.cu)
  void *fsd = nullptr;          // default -fs tree: GMRES + ILU(0) sub-solvers (xsb_fs.cu)
  const double *nodal_in = nullptr;   // coarse -mg level: nodal Q1 coefficient fields [slot][p-node] to use instead of the model
};

int xsb_fail(xsb_ctx c, int code, const char *fmt, ...);
#define CUDA_OK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return xsb_fail(c, XSB_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); } while (0)
#define XSB_CHK(call) do { int rc_ = (call); if (rc_) return rc_; } while (0)
#define KERNEL_OK() do { cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) return xsb_fail(c, XSB_ERR_CUDA, "%s:%d kernel launch: %s", __FILE__, __LINE__, cudaGetErrorString(e_)); c->n_launch++; } while (0)

template <class T> int dev_alloc(xsb_ctx c, T **p, size_t n);
int dev_free_all(xsb_ctx c);
int dev_free(xsb_ctx c, void *p);          // one allocation made by dev_alloc
int dev_free_phase(xsb_ctx c, int phase);  // every allocation of one phase

// ---- xsb_fe.cu
int fe_resolve_model(xsb_ctx c);
int fe_assemble(xsb_ctx c);
int bc_list_faces(int nsd, int lame, int model, int freeslip, int mx, int my, int mz, int zlo, int zhi, int32_t *idx, double *val, int cap);
// ---- xsb_spmv.cu
enum { EPI_PLAIN = 0, EPI_RESIDUAL = 1, EPI_CHEB_FIRST = 2, EPI_CHEB = 3 };
struct Epilogue { int mode = EPI_PLAIN; const double *b = nullptr, *idiag = nullptr, *pk = nullptr, *pkm1 = nullptr; double s0 = 0, s1 = 0, s2 = 0; };
int spmv_csr(xsb_ctx c, const Csr &A, const double *x, double *y, int64_t row0 = 0, int64_t nrows = -1, const double *yadd = nullptr);   // y = A x (+ yadd)
int spmv_baij(xsb_ctx c, const Baij &A, const double *x, double *y, const Epilogue &ep, int node0 = 0, int nnodes = -1);
int spmv_baij_pair(xsb_ctx c, const Baij &A, const double *x, double *y, const Epilogue &ep, int nodeA, int nodeB, int nnodes);   // two row ranges, one launch
int spmv_a00_fine(xsb_ctx c, const Baij &A, const double *x, double *y, const Epilogue &ep);   // counted (+ timed) fine-level launch
int spmv_collect_timing(xsb_ctx c);
int prof_mark(xsb_ctx c, int cat);   // -xsb_time_kernels: the time from here to the next mark belongs to category `cat`
// ---- xsb_mf.cu
int mf_setup(xsb_ctx c);
int mf_a00_apply(xsb_ctx c, const double *x, double *y, const Epilogue &ep);
int mf_a00_apply_raw(xsb_ctx c, const double *x, double *y);   // without the Dirichlet rows / columns (A before MatZeroRowsColumns)
void mf_tab_scaled(const Lattice &L, MfTabS &T);
// ---- xsb_mf1p.cu (one-pass, TMA-staged element kernel)
int mf1p_apply(xsb_ctx c, const double *x, double *y, const Epilogue &ep, const unsigned char *bcnode, int zlo, int zhi);
int mf1p_prepare(xsb_ctx c, int zlo, int zhi);   // allocations / tables of the kernel, outside any graph capture
int mf1p_partition(int P, int64_t ncols, int nl, int p, int64_t *lo, int64_t *hi);   // host mirror of the kernel's work split (tests)
// ---- xsb_mfull.cu (operator-free mode)
int mf_diag_inv(xsb_ctx c, double *idiag);                    // 1 / diag(A00) from the element matrices
int galerkin_elements(xsb_ctx c, Level &C);                   // P^T A00 P assembled element by element on the local lattice
// ---- xsb_vec.cu
int vec_set(xsb_ctx c, int64_t n, double a, double *x);
int vec_copy(xsb_ctx c, int64_t n, const double *x, double *y);
int vec_axpy(xsb_ctx c, int64_t n, double a, const double *x, double *y);            // y += a x
int vec_aypx(xsb_ctx c, int64_t n, double a, const double *x, double *y);            // y = x + a y
int vec_scale(xsb_ctx c, int64_t n, double a, double *x);
int vec_pmult(xsb_ctx c, int64_t n, const double *d, const double *x, double *y);    // y = d .* x
int vec_waxpy(xsb_ctx c, int64_t n, double a, const double *x, const double *y, double *w); // w = y + a x
int vec_mdot(xsb_ctx c, const Ranges &rg, const double *w, double *const *V, int k, bool with_norm, double *out_dev, bool local = false); // out[j] = w.V[j], out[k] = w.w (owned entries, all ranks; local: this rank's sum only, for replicated vectors)
int vec_maxpy_dev(xsb_ctx c, int64_t n, double *w, double *const *V, int k, const double *coef_dev, double sign); // w += sign * sum coef[j] V[j]
int vec_maxpy_host(xsb_ctx c, int64_t n, double *w, double *const *V, int k, const double *coef_host);
int vec_scale_by_inv_sqrt(xsb_ctx c, int64_t n, double *w, const double *nrm2_dev);  // w /= sqrt(*nrm2)
int vec_gcr_update(xsb_ctx c, const Ranges &rg, const double *dots_dev /* [r.v, v.v] */, double *v, double *s, double *x, double *r, double *rnorm2_dev);
int vec_fetch(xsb_ctx c, const double *dev, int n, double *host);   // sync copy of n scalars
int vec_diagnostics(xsb_ctx c, const double *x, double *out);
int vec_rander48(xsb_ctx c, int64_t n, int interval, double *x, int64_t stream_offset = 0);
// ---- xsb_comm.cu
int comm_init(xsb_ctx c, const void *unique_id, int rank, int nranks);
int comm_unique_id(void *out128);
int comm_destroy(xsb_ctx c);
int comm_allreduce_sum(xsb_ctx c, double *dev, int n);
int comm_halo_u(xsb_ctx c, double *u);            // fill the velocity ghost planes owned rows read (2 below, 1 above)
int comm_halo_p(xsb_ctx c, double *p);            // pressure ghost planes (1 below, 1 above)
int comm_halo_full(xsb_ctx c, double *x);         // [u|p]
int comm_bcast_segments(xsb_ctx c, double *glob, const int64_t *offs /* nranks+1 */);
int comm_bcast_planes(xsb_ctx c, double *glob, int64_t plane_doubles, int nplanes_glob);   // every rank contributes its owned coarse planes
int comm_bcast_plane_ranges(xsb_ctx c, double *glob, int64_t plane_doubles, const int *p0, const int *p1);   // rank r contributes planes [p0[r], p1[r])
int comm_halo_planes(xsb_ctx c, double *v, int64_t plane_doubles, int o0, int o1, int gb, int ga, cudaStream_t on = nullptr);   // ghost planes of any lattice vector (owned planes [o0,o1))
int comm_p2p_setup(xsb_ctx c);      // collective: peer-memory windows for the halo exchange (NVLink), called by xsb_assemble
int comm_p2p_active(xsb_ctx c);
int comm_p2p_check(xsb_ctx c);      // sticky time-out flag of the peer-memory exchange
void comm_p2p_destroy(xsb_ctx c);
inline Ranges whole(int64_t n) { Ranges r; r.len0 = n; return r; }
// ---- xsb_mg.cu
int mg_setup(xsb_ctx c);
int mg_vcycle(xsb_ctx c, const double *b, double *x);
void mg_graphs_release(xsb_ctx c);
int mg_restrict(xsb_ctx c, const Level &F, const Level &C, const double *rf, double *bc);
int mg_prolong_add(xsb_ctx c, const Level &F, const Level &C, const double *xc, double *xf);
int baij_to_csr_host(xsb_ctx c, const Baij &A, int32_t *ia, int32_t *ja, double *a);
int baij_diag_inv(xsb_ctx c, const Baij &A, double *idiag);
int mg_restrict_scalar(xsb_ctx c, int fnx, int fny, int fnz, int cnx, int cny, int cnz, const double *rf, double *bc);   // pressure lattice
int mg_prolong_add_scalar(xsb_ctx c, int fnx, int fny, int fnz, int cnx, int cny, const double *xc, double *xf);
// ---- xsb_mmg.cu (monolithic -mg)
int mmg_setup(xsb_ctx c);
int mmg_apply(xsb_ctx c, const double *r, double *z);
void mmg_free(xsb_ctx c);
// ---- xsb_grad.cu (gradient / divergence blocks matrix-free)
int grad_apply(xsb_ctx c, const double *xp, double *y, int64_t dof0, int64_t ndofs, const double *yadd = nullptr);   // y_u rows = A01 xp (+ yadd)
int div_apply(xsb_ctx c, const double *xu, double *y, int64_t p0, int64_t np, const double *yadd = nullptr);         // y_p rows = A10 xu (+ yadd)
void grad_free(xsb_ctx c);
int grad_prepare(xsb_ctx c);   // coefficient tables of the current lattice (idempotent)
// ---- xsb_asm.cu (additive Schwarz on the reference's element patches)
int asm_setup(xsb_ctx alloc, xsb_ctx problem, int size, int overlap, void **out);
int asm_apply(xsb_ctx c, void *asmpc, const double *r, double *z);
void asm_free(void *asmpc);
int dense_invert_pivoted(xsb_ctx c, int n, double *M, double *Inv);
// ---- xsb_fs.cu (plain -fs tree with PETSc's default sub-solvers)
int fsd_setup(xsb_ctx c);
int fsd_apply(xsb_ctx c, const double *r, double *z);
void fsd_free(xsb_ctx c);
int fsc_setup(xsb_ctx alloc, xsb_ctx coarse_problem, void **out);   // -fs_coarse: fieldsplit-preconditioned FGMRES on the coarse -mg level
int fsc_solve(xsb_ctx c, void *fsc, const double *b, double *x);
int fsc_last_its(void *fsc);
void fsc_free(void *fsc);
// ---- xsb_ilu.cu
int ilu_setup(xsb_ctx c);
int ilu_apply(xsb_ctx c, const double *b, double *x);
// ---- xsb_ksp.cu
int ksp_setup(xsb_ctx c);
int ksp_release(xsb_ctx c);
int op_full_mult(xsb_ctx c, const double *x, double *y);
int ksp_solve(xsb_ctx c, const double *b_dev, double *x_dev);
int pc_apply(xsb_ctx c, const double *r, double *z, int *inner, int *inner_reason = nullptr);
int hess_eig(int n, const double *H, int ldh, double *wr, double *wi);
int csr_diag_inv(xsb_ctx c, const Csr &A, double *idiag);
int csr_diag(xsb_ctx c, const Csr &A, double *diag);
int baij_diag(xsb_ctx c, const Baij &A, double *diag);
