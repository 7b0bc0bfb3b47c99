/* exsaddle_b200.h -- C ABI of the B200-native exSaddle solve path.
 *
 * Plain C, opaque handles, int error codes (0 = success), caller-owned buffers, no torch / C++ types.
 * Every entry point names the reference interface it replaces (paths relative to psanan/exsaddle).
 * The reference is a PETSc application: its "plugin API" for this path is PETSc's Mat / PC / KSP
 * callback tables (the pattern the reference itself demonstrates with PCRegister + pc->ops in
 * pcildl.c:460-485 and exSaddle.c:110-115) and the options database (-saddle_* keys, abf.opts).
 * INTEGRATION.md shows the PETSc-side shim (MatCreate_ExSaddleB200 / PCCreate_ExSaddleB200) that
 * binds these functions into MatRegister / PCRegister.
 *
 * Vectors use the reference's one-rank DMComposite ordering: all velocity dofs (node-major, component
 * fastest, node = i + j*NX + k*NX*NY on the (2mx+1)(2my+1)(2mz+1) lattice), then all pressure dofs
 * (femixedspace.c:1150-1158, 1243-1249, 1312-1315).  Entry points ending in _dev take device pointers
 * (same ordering) and run asynchronously on the handle's stream; the others take host pointers and are
 * synchronous.  There is no CPU fallback: every compute call returns XSB_ERR_NO_DEVICE without a GPU.
 */
#ifndef EXSADDLE_B200_H_
#define EXSADDLE_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XSB_VERSION 100

/* error classes (PetscErrorCode convention: 0 = OK; callers wrap with CHKERRQ, pcildl.c:65-66,318) */
enum {
  XSB_OK = 0,
  XSB_ERR_ARG = -1,        /* PETSC_ERR_ARG_* : bad argument / option value */
  XSB_ERR_SUP = -2,        /* PETSC_ERR_SUP  : unsupported combination (exSaddle.c:205-213) */
  XSB_ERR_ORDER = -3,      /* PETSC_ERR_ORDER: called before setup */
  XSB_ERR_MEM = -4,        /* PETSC_ERR_MEM */
  XSB_ERR_CUDA = -5,       /* CUDA runtime error, text in xsb_last_error() */
  XSB_ERR_NO_DEVICE = -6,  /* no CUDA device: the library has no CPU path */
  XSB_ERR_BREAKDOWN = -7,  /* zero pivot / Krylov breakdown */
  XSB_ERR_NCCL = -8
};

/* which operator (MatCreateSubMatrix on the DMComposite ISs, exSaddle.c:319-321; Mpscaled :315-318) */
enum {
  XSB_MAT_A = 0,    /* full saddle operator, AIJ (DMCreateMatrix_SaddleAIJ femixedspace.c:599) */
  XSB_MAT_A00 = 1,  /* velocity block */
  XSB_MAT_A01 = 2,  /* gradient block */
  XSB_MAT_A10 = 3,  /* divergence block */
  XSB_MAT_A11 = 4,  /* pressure block (stored zeros for Stokes, -1/lambda mass for LAME) */
  XSB_MAT_MP = 5,   /* scaled pressure mass matrix (MatAssemble_Schur femixedspace.c:2837) */
  XSB_MAT_A00_MF = 6, /* the velocity block applied matrix-free (sum-factorised element kernel; 3-D): xsb_mat_mult only */
  XSB_MAT_A01_MF = 7, /* the gradient block applied by its closed-form stencil (no matrix read): xsb_mat_mult only */
  XSB_MAT_A10_MF = 8, /* the divergence block, likewise */
  XSB_MAT_MG_LEVEL0 = 16 /* + l : Galerkin operator of PCMG level l (0 = coarsest), after xsb_ksp_setup */
};

typedef struct xsb_ctx_s *xsb_ctx;

/* -------- lifecycle: PCCreate_X / PCReset_X / PCDestroy_X (pcildl.c:376-407, 460-485) ------------------ */
int xsb_create(xsb_ctx *ctx, int nsd /* -DNSD: 2|3 */, int lame /* -DLAME */, int device /* CUDA ordinal, -1 = current */);
int xsb_reset(xsb_ctx ctx);        /* frees device state, keeps options; idempotent */
int xsb_destroy(xsb_ctx *ctx);     /* frees the handle; *ctx = NULL */
const char *xsb_last_error(xsb_ctx ctx);
int xsb_device_available(void);    /* 1 if a CUDA device is usable */

/* -------- options database: PetscOptionsGet* (exSaddle.c:169-203, models.c, abf.opts) -------------------- */
int xsb_set_option(xsb_ctx ctx, const char *key /* e.g. "-saddle_ksp_rtol" */, const char *value /* NULL for flags */);
int xsb_set_options(xsb_ctx ctx, const char *cmdline);       /* whole option string, PETSc syntax, '#' comments */
int xsb_set_options_file(xsb_ctx ctx, const char *path);     /* -options_file abf.opts */
int xsb_options_left(xsb_ctx ctx, char *buf, int buflen);    /* -options_left: unused keys, newline separated */

/* -------- FE set-up: exSaddle.c:215-283 -----------------------------------------------------------------
   DMCreate_SaddleQ2Q1 + FEMixedSpaceQuadratureCreate + FEMixedSpaceBCISCreate +
   FEMixedSpaceDefineQPwiseProperties(+_Q1Projection) + DMCreateMatrix + MatAssemble_Saddle_NULL +
   MatAssemble_Saddle + VecAssemble_F1_qp/F2_qp + ImposeDirichletValuesIS + MatAssemble_Schur, on the device. */
int xsb_assemble(xsb_ctx ctx);
int xsb_banner(xsb_ctx ctx, char *buf, int buflen);          /* "Boundary Conditions: ..." / "ModelType: ..." lines */

/* sizes: out[0]=rows out[1]=u dofs out[2]=p dofs out[3]=nnz(A) out[4]=preallocated nnz out[5]=elements
          out[6]=Dirichlet dofs out[7]=nnz(Mp) */
int xsb_get_sizes(xsb_ctx ctx, int64_t out[8]);

/* -------- Mat: MatGetRowIJ + MatSeqAIJGetArray (pcildl.c:305-316), MatMult, MatGetDiagonal ------------ */
int xsb_mat_get_info(xsb_ctx ctx, int which, int64_t *rows, int64_t *cols, int64_t *nnz, int *bs);
int xsb_mat_get_csr(xsb_ctx ctx, int which, int32_t *ia, int32_t *ja, double *a); /* host out; NULL to skip */
int xsb_mat_mult(xsb_ctx ctx, int which, const double *x, double *y);            /* host pointers */
int xsb_mat_mult_dev(xsb_ctx ctx, int which, const double *x, double *y);        /* device pointers */
int xsb_mat_mult_transpose(xsb_ctx ctx, int which, const double *x, double *y);  /* MatMultTranspose, host pointers: A, A00, A11, Mp are
                                                                                    symmetric; A01^T = A10 (femixedspace.c:2584-2590) */
int xsb_mat_get_diagonal(xsb_ctx ctx, int which, double *d);
int xsb_vec_get_rhs(xsb_ctx ctx, double *F);                                      /* F of exSaddle.c:263-281 */
int xsb_get_bc(xsb_ctx ctx, int32_t *idx, double *val);                           /* u_is_global / u_bc_global */
int xsb_get_coeff_qp(xsb_ctx ctx, int slot /* 0 eta|mu 1 Fu0 2 Fu1 3 Fu2 4 Fp 5 lambda */, double *out /* nel*nqp */);

/* -------- KSP / PC: KSPSetFromOptions + KSPSetUp + KSPSolve (exSaddle.c:304-322, 405-425) ---------------
   Solver tree from the -saddle_* options: gmres|fgmres, pc jacobi | fieldsplit(Schur,UPPER,user Mpscaled)
   with fieldsplit_u = gcr + mg(Galerkin, chebyshev/jacobi, LU coarse), fieldsplit_p = preonly + bjacobi/ilu(0). */
int xsb_ksp_setup(xsb_ctx ctx);
int xsb_ksp_reset(xsb_ctx ctx);    /* PCReset_X (pcildl.c:376-394) of a PC that shares the handle with the Mat: frees the solver state
                                      (MG hierarchy, ILU factors, Krylov bases), keeps the assembled operator and the options; idempotent */
int xsb_get_state(xsb_ctx ctx, int *assembled, int *ksp_ready);
int xsb_ksp_solve(xsb_ctx ctx, const double *b /* host, NULL = assembled F */, double *x /* host out */);
int xsb_ksp_solve_dev(xsb_ctx ctx, const double *b /* device, NULL = F */, double *x /* device out */);
int xsb_pc_apply(xsb_ctx ctx, const double *r, double *z);        /* PCApply of the outer PC, host pointers */
int xsb_pc_apply_dev(xsb_ctx ctx, const double *r, double *z);
int xsb_pc_mg_apply(xsb_ctx ctx, const double *b, double *x);     /* PCApply_MG on A00 (one V-cycle), host pointers */
int xsb_pc_schur_apply(xsb_ctx ctx, const double *b, double *x);  /* fieldsplit_p PC on Mpscaled, host pointers */
/* measurement aid: average device time (ms) of `reps` pressure-block solves on resident vectors */
int xsb_time_pc_schur(xsb_ctx ctx, int reps, double *ms_per_apply);
/* measurement aid (collective on slabs): average device time in us of [0] a velocity ghost exchange, [1] a one-plane exchange
   (distributed coarse level), [2] a fine-level A00 product without its exchange */
int xsb_time_halo(xsb_ctx ctx, int reps, double out[3]);
int xsb_mg_restrict(xsb_ctx ctx, int coarse_level, const double *rf, double *bc);     /* MatRestrict */
int xsb_mg_interpolate_add(xsb_ctx ctx, int coarse_level, const double *xc, double *xf); /* MatInterpolateAdd */

/* results of the last solve: KSPGetIterationNumber / KSPGetConvergedReason / KSPGetResidualHistory */
int xsb_ksp_get_iterations(xsb_ctx ctx, int *its, int *reason);
int xsb_ksp_get_history(xsb_ctx ctx, double *hist, int cap, int *n);
int xsb_ksp_get_inner_iterations(xsb_ctx ctx, int *its, int cap, int *n);     /* fieldsplit_u GCR counts */
int xsb_ksp_get_inner_reasons(xsb_ctx ctx, int *reasons, int cap, int *n);    /* KSPConvergedReason of each fieldsplit_u solve (-saddle_fieldsplit_u_ksp_converged_reason) */
int xsb_ksp_get_chebyshev(xsb_ctx ctx, int level, double *emin_est, double *emax_est, double *emin, double *emax);
int xsb_ksp_get_timing(xsb_ctx ctx, double *setup_ms, double *solve_ms);      /* CUDA-event times */
/* KSPView / PCView (pc->ops->view, pcildl.c:429-456; -saddle_ksp_view): the solver tree in use, level sizes, nonzeros and
   Chebyshev bounds, as text (PETSc-like layout, not byte-identical to PETSc's viewer) */
int xsb_ksp_view(xsb_ctx ctx, char *buf, int buflen);
/* counters of the last solve: [0] fine-level A00 block SpMV launches [1] full-A SpMV launches [2] all kernel
   launches [3] average fine-level A00 SpMV device time in ns (CUDA-event pairs on the launching stream around
   every such launch when -xsb_time_kernels is set; read back after the solve, no extra synchronisation)
   [4..7] fine-level A00 launches by fused epilogue: plain y=Ax, residual b-Ax, first Chebyshev step, Chebyshev step */
int xsb_ksp_get_counters(xsb_ctx ctx, int64_t out[8]);
/* -xsb_time_kernels: device time of the last solve by category (the -log_view stages a PETSc user would read; events on the
   launching stream cut the solve into stretches, each booked on one category).  ms[i], count[i] for i < min(cap, *ncat):
   0 Krylov vector kernels + host gaps, 1 ghost exchange in front of fine-level A00 products, 2 fine-level A00 kernels,
   3+l products on MG level l (l < 10; level 0: the coarse solve), 13 plane exchange behind products of distributed coarse
   levels, 14 restriction / interpolation (with their exchanges), 15 pressure-block ILU(0) solves, 16 full-operator products
   of the outer Krylov method, 17 ghost exchange + A01 product of the fieldsplit */
int xsb_ksp_get_profile(xsb_ctx ctx, double *ms, int64_t *count, int cap, int *ncat);
/* the CUDA stream (cudaStream_t) every kernel of this handle is launched on, for event timing by the caller */
int xsb_get_stream(xsb_ctx ctx, void **stream);

/* SaddleReportSolutionDiagnostics (exSaddle_io.c:7-58): out[5*nsd+5] = {1,2,inf,min,max} x comps, then p */
int xsb_diagnostics(xsb_ctx ctx, const double *x /* host */, double *out);

/* -------- PETSc binary dumps: DumpOperator / DumpSolution (exSaddle_io.c:61-88; -dump_operator -> operator_<k>.petscbin,
   -dump_solution -> solution.petscbin, -dump_scaled_mass_matrix -> mpscaled.petscbin, exSaddle.c:488-501, 535-537).
   The files are what PetscViewerBinaryOpen + MatView / VecView write (big-endian; Mat: 1211216, rows, cols, nnz, row
   lengths, columns, values; Vec: 1211214, n, values) and load with PetscBinaryRead (octave_demo.m:10-12) / MatLoad. */
int xsb_dump_operator(xsb_ctx ctx, int which, const char *path);
int xsb_dump_vector(xsb_ctx ctx, const double *x /* host */, int64_t n, const char *path);
/* the writers themselves: host arrays in, no GPU needed */
int xsb_write_petsc_mat(const char *path, int64_t rows, int64_t cols, const int32_t *ia, const int32_t *ja, const double *a);
int xsb_write_petsc_vec(const char *path, int64_t n, const double *x);

/* ViewFields (exSaddle_io.c:128-177; -view_fields): <dir>/<tag>uv[w].vts (velocity components as scalar point fields on
   the velocity lattice) and <dir>/<tag>p.vts, VTK XML StructuredGrid with raw appended data (loads in ParaView / VisIt). */
int xsb_view_fields(xsb_ctx ctx, const double *x /* host */, const char *dir, const char *tag /* "" or "ref_" */);
int xsb_write_vts(const char *path, int nx, int ny, int nz, const double h[3], int nfields, const char *const *names,
                  const double *data, int64_t field_stride, int64_t node_stride);   /* the writer itself: host arrays, no GPU */

/* -------- host-side index maps (no GPU needed; integer logic only) --------------------------------------- */
/* columns of AIJ row `row` in ascending order (pattern of MatAssemble_Saddle_NULL, femixedspace.c:2306-2370) */
int xsb_pattern_row(int nsd, int mx, int my, int mz, int64_t row, int32_t *cols, int cap);
/* preallocated nnz total of SaddlePreallocation_SEQ (femixedspace.c:181-286) */
int64_t xsb_prealloc_total(int nsd, int mx, int my, int mz);
/* Dirichlet dof list of ISCreate_BCList (models.c:610-648); returns count, fills up to cap */
int xsb_bc_list(int nsd, int lame, int model, int freeslip, int mx, int my, int mz, int32_t *idx, double *val, int cap);
/* PCMG level lattice: DMCoarsen n -> (n-1)/2+1; returns 0 or XSB_ERR_ARG when not coarsenable */
int xsb_mg_level_dims(int nsd, int mx, int my, int mz, int levels, int level, int dims[3]);

/* -------- gradient / divergence blocks without a matrix (csrc/xsb_grad.cu) ----------------------------------- */
/* host-only: the per-direction coefficient tables behind XSB_MAT_A01_MF / XSB_MAT_A10_MF for a line of m elements with node
   spacing h.  Velocity node i couples to the pressure nodes uP[3i..3i+2] (-1 = none) with the 1-D mass-type factors uM and
   derivative-type factors uG (sums of the Q2 x Q1 element-table entries over the elements containing both nodes); pM / pG hold the
   same numbers seen from pressure node P: its velocity nodes 2P-2 .. 2P+2.  An entry of A01 (MatAssemble_Saddle,
   femixedspace.c:2576-2579) is minus the product over the directions, G-type in the component's own direction. */
int xsb_grad_line_tables(int m, double h, int32_t *uP, double *uM, double *uG, double *pM, double *pG);

/* -------- ASM on the reference's element patches --------------------------------------- */
/* Process grid PETSc's DMDACreate{2,3}d(PETSC_DECIDE) picks for M x N (x P) nodes on `size` ranks (the velocity DMDA of
   femixedspace.c:1153-1158); XSB_ERR_ARG when `size` cannot be factored onto the lattice. */
int xsb_dmda_grid(int nsd, int M, int N, int P, int size, int out[3]);
/* Subdomain DMCreateDomainDecomposition_DMDAFEQ2Q1 (femixedspace.c:745-837) gives rank `rank` of `size` (x fastest in the process
   grid) with -dmdafe_overlap `overlap`: out[0..2] first element and out[3..5] one past the last element of the closed patch,
   out[6..8] / [9..11] the velocity-node range [lo,hi) the rank owns, out[12..14] / [15..17] the pressure-node range it owns.
   XSB_ERR_ARG where the reference stops with "Cannot generate consistent macro element" (femixedspace.c:1097-1113).
   The solver uses it with `-saddle_pc_type asm -saddle_pc_asm_dm_subdomains -set_ksp_dm -xsb_ranks <size>`. */
int xsb_asm_subdomain(int nsd, int mx, int my, int mz, int size, int overlap, int rank, int out[18]);

/* -------- multi-GPU: z-slab partition, one process per GPU ------------------------------------ */
/* Element z-range [k0,k1) owned by `rank` of `nranks` (elements split like the reference's pressure rule,
   femixedspace.c:1231-1240: mz/nranks each, remainder to the low ranks). */
int xsb_slab_range(int mz, int nranks, int rank, int *k0, int *k1);
/* Node planes [p0,p1) of coarse velocity-MG level `depth` below the fine level (0 = first coarse level, mz+1 planes) whose
   rows `rank` computes when the level is distributed: the planes of its element layers, halved per coarsening (PETSc
   analogue: the DMDA ownership of the coarsened DM, exSaddle.c:408-422 / DMCoarsen).  All ranks' ranges tile the level. */
int xsb_pdist_range(int mz, int nranks, int rank, int depth, int *p0, int *p1);
/* host-only: the partition xsb_get_partition reports, computed from the mesh alone (same out[12] layout) */
int xsb_slab_layout(int nsd, int mx, int my, int mz, int nranks, int rank, int64_t out[12]);
/* NCCL bootstrap (one process per GPU; replaces MPI_Init/PETSC_COMM_WORLD of the reference, femixedspace.c:645,684):
   rank 0 calls xsb_comm_unique_id and ships the 128 bytes to the other ranks (the launcher's store, MPI or a file);
   every rank then calls xsb_comm_init BEFORE xsb_assemble.  nranks = 1 is a no-op. */
int xsb_comm_unique_id(void *out128);
/* out[0]=rank [1]=nranks [2]=1 when ghost planes travel through the peer-memory exchange kernel (0: ncclSend/ncclRecv)
   [3]=bit l set when MG level l is distributed by node planes (after xsb_ksp_setup) */
int xsb_comm_info(xsb_ctx ctx, int64_t out[4]);
int xsb_comm_init(xsb_ctx ctx, const void *unique_id, int rank, int nranks);
/* partition of this rank after xsb_assemble: out[0]=rank [1]=nranks [2..3]=owned element layers [k0,k1)
   [4..5]=local lattice layers [e0,e1) [6]=offset,[7]=length of the owned velocity entries in a local vector
   [8],[9]= same for the owned pressure entries [10],[11]=global (one-rank DMComposite) index of the first owned
   velocity / pressure dof.  Local vectors are [u_local | p_local] on the local lattice (ghost planes included). */
int xsb_get_partition(xsb_ctx ctx, int64_t out[12]);

#ifdef __cplusplus
}
#endif
#endif
