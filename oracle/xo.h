/* xo.h -- CPU oracle for the exSaddle Q2-Q1 solve path.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library; the
 * product (exsaddle_b200/) never includes, links or calls anything in oracle/.
 *
 * It is a plain-C restatement of the reference's algorithm for the path named by
 * BASELINE.json:north_star.  The discretisation follows /root/reference/femixedspace.c
 * and models.c function by function (citations at each function); the PETSc-side
 * algorithms (KSP/PC) that live outside /root/reference are restated from PETSc
 * 3.12-3.14 behaviour.  PETSc itself is absent from the image, so the
 * reference cannot be compiled here; parity is pinned instead on the reference's own
 * golden outputs (testref/ *.ref), see tests/test_oracle_goldens.py.
 */
#ifndef XO_H_
#define XO_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XO_MAX_LEVELS 10 /* MG_DEPTH, exSaddle.h:24 */

/* Options that shape the discrete problem (exSaddle.c:178-185, models.c option lookups).
   Doubles set to NAN / ints set to -1 mean "not given": the model default applies. */
typedef struct {
  int    nsd;         /* -DNSD: 2 or 3 (exSaddle.h:7) */
  int    lame;        /* -DLAME (Makefile:97-99) */
  int    mx, my, mz;  /* -mx -my -mz (exSaddle.c:178-182) */
  double size[3];     /* -size_x/y/z (exSaddle.c:183-185) */
  int    model;       /* -model; -1 = DEFAULT_MODEL (models.h:9-13) */
  double c0, c1;      /* -eta0/-eta1 (Stokes) or -mu0/-mu1 (Lame) */
  double lam0, lam1;  /* -lambda0/-lambda1 */
  double sinker_r;    /* -sinker_r */
  double sinker_c[3]; /* -sinker_x/y/z (model 6 Stokes only) */
  int    sinker_n;    /* -sinker_n */
  double solcx_xc;    /* -solcx_xc */
  int    solcx_nz;    /* -solcx_nz */
  int    freeslip;    /* -freesliphack (models.c:29) */
} xo_params;

/* Solver tree options (subset of the PETSc options the reference's tests use). */
enum { XO_KSP_GMRES = 0, XO_KSP_FGMRES = 1 };
enum { XO_PC_NONE = 0, XO_PC_JACOBI = 1, XO_PC_ABF = 2 };
enum { XO_SIDE_LEFT = 0, XO_SIDE_RIGHT = 1 };
enum { XO_PPC_ILU0 = 0, XO_PPC_JACOBI = 1 };

typedef struct {
  int    ksp_type;   /* -saddle_ksp_type */
  int    pc_type;    /* -saddle_pc_type jacobi | (-fs) fieldsplit */
  int    pc_side;    /* -saddle_ksp_pc_side (gmres default left, fgmres right) */
  double rtol, atol, dtol;
  int    max_it, restart;
  /* fieldsplit_u: GCR + PCMG (abf.opts:4-13) */
  double u_rtol;
  int    u_max_it, u_restart;
  int    mg_levels;       /* -saddle_fieldsplit_u_pc_mg_levels */
  int    cheb_its;        /* -..._mg_levels_ksp_max_it */
  double esteig[4];       /* -..._ksp_chebyshev_esteig a,b,c,d */
  int    esteig_steps;    /* 10 */
  int    noise;           /* 0: rander48 on [0,1), 1: rander48 on [-1,1) */
  int    n_cheb_fixed;    /* >0: explicit (emin,emax) per level 1..; PETSc -ksp_chebyshev_eigenvalues */
  double cheb_emin[XO_MAX_LEVELS], cheb_emax[XO_MAX_LEVELS];
  int    p_pc;            /* XO_PPC_ILU0 (bjacobi/ilu) or XO_PPC_JACOBI */
  int    max_outer_sample; /* >0: stop after this many outer its (CPU-baseline sample) */
  int    p_blocks;         /* bjacobi blocks of the pressure PC: z-slabs of pressure planes like the product's N-rank slab partition (0/1 = one block) */
} xo_solver;

typedef struct {
  int    its;
  int    reason;     /* PETSc KSPConvergedReason: 2 RTOL, 3 ATOL, -3 DIVERGED_ITS, -4 DTOL */
  int    nhist;
  double hist[2048];             /* monitor values, hist[i] = residual at iteration i */
  int    inner_its[2048];            /* GCR iterations of each fieldsplit_u solve */
  int    n_inner;
  double cheb_emin_est[XO_MAX_LEVELS], cheb_emax_est[XO_MAX_LEVELS]; /* Ritz extremes per level */
  double cheb_emin[XO_MAX_LEVELS], cheb_emax[XO_MAX_LEVELS];         /* bounds in use */
  int    level_rows[XO_MAX_LEVELS];
  int64_t level_nnz[XO_MAX_LEVELS];
  double setup_seconds, solve_seconds;
  int64_t n_a00_mult, n_a_mult;
} xo_result;

typedef struct xo_problem_s xo_problem;

void xo_params_init(xo_params *p, int nsd, int lame);
void xo_solver_init(xo_solver *s);
void xo_solver_abf(xo_solver *s); /* abf.opts */

/* Build everything exSaddle.c:215-283 builds: mesh, coefficients, A (BCs imposed), F, Mpscaled. */
int  xo_create(const xo_params *prm, xo_problem **out);
/* coarse level of the monolithic -mg hierarchy: coefficients from nodal Q1 fields (npn x 6, node-major) */
int  xo_create_nodal(const xo_params *prm, const double *nodal, xo_problem **out);
const double *xo_coeff_nodal(const xo_problem *p); /* npn x 6 nodal Q1 coefficient fields (after the projection) */
void xo_destroy(xo_problem *p);
const char *xo_banner(const xo_problem *p); /* BC / model banner lines (models.c PetscPrintf) */
const char *xo_error(const xo_problem *p);

/* sizes: [0] rows, [1] u dofs, [2] p dofs, [3] nnz(A), [4] preallocated nnz, [5] nel, [6] nbc, [7] nnz(Mp) */
void xo_sizes(const xo_problem *p, int64_t out[8]);
const int    *xo_A_ia(const xo_problem *p);
const int    *xo_A_ja(const xo_problem *p);
const double *xo_A_a(const xo_problem *p);
const double *xo_A_raw(const xo_problem *p); /* values before MatZeroRowsColumns (femixedspace.c:2645) */
const int    *xo_Mp_ia(const xo_problem *p);
const int    *xo_Mp_ja(const xo_problem *p);
const double *xo_Mp_a(const xo_problem *p);
const double *xo_F(const xo_problem *p);
const int    *xo_bc_idx(const xo_problem *p);
const double *xo_bc_val(const xo_problem *p);
const int    *xo_u_map(const xo_problem *p);
const int    *xo_p_map(const xo_problem *p);
const double *xo_coeff_qp(const xo_problem *p); /* nel*nqp*ncoeff after Q1 projection */
int           xo_ncoeff(const xo_problem *p);

/* y = A x on the assembled AIJ operator */
void xo_A_mult(const xo_problem *p, const double *x, double *y);
void xo_csr_mult(int n, const int *ia, const int *ja, const double *a, const double *x, double *y);

/* index-set extraction of a sub-block (MatCreateSubMatrix on the u/p ISs, exSaddle.c:319-321):
   rb, cb in {0 (u), 1 (p)}.  Two-pass: call with ja=a=NULL to get nnz and ia. */
int64_t xo_submatrix(const xo_problem *p, int rb, int cb, int *ia, int *ja, double *a);

/* Galerkin hierarchy of A00 as PCMG builds it (App. B.3). Level 0 = coarsest. Returns rows/nnz and CSR. */
int  xo_mg_setup(xo_problem *p, int levels);
int  xo_mg_level_csr(const xo_problem *p, int level, int *n, const int **ia, const int **ja, const double **a);
void xo_mg_prolong_add(const xo_problem *p, int level_coarse, const double *xc, double *xf); /* xf += P xc */
void xo_mg_restrict(const xo_problem *p, int level_coarse, const double *rf, double *bc);    /* bc  = P^T rf */

/* ILU(0) of Mp in natural ordering; returns factors in Mp's pattern (L unit-lower, U with inverted diagonal NOT applied). */
int  xo_ilu0(int n, const int *ia, const int *ja, const double *a, double *lu);
void xo_ilu0_solve(int n, const int *ia, const int *ja, const double *lu, const double *b, double *x);

/* One application of the ABF preconditioner (PCApply_FieldSplit_Schur, UPPER) after xo_solve set it up. */
int  xo_solve(xo_problem *p, const xo_solver *s, const double *b /* NULL: F */, double *x, xo_result *res);
int  xo_pc_setup(xo_problem *p, const xo_solver *s, xo_result *res);
int  xo_pc_apply(xo_problem *p, const double *r, double *z, int *inner_its);
int  xo_vcycle(xo_problem *p, const double *b, double *x); /* PCApply_MG on A00 */

/* SaddleReportSolutionDiagnostics (exSaddle_io.c:7-58): out[0..] = for each of 5 stats (1,2,inf,min,max) nsd comps; then p: 1,2,inf,min,max */
void xo_diagnostics(const xo_problem *p, const double *x, double *out /* 5*nsd + 5 */);

/* helpers exposed for unit tests */
void xo_rander48(int n, int interval, double *v);          /* PETSc rander48 stream, seed 0x12345678 */
int  xo_hess_eig(int n, const double *H, int ldh, double *re, double *im); /* eigenvalues of upper Hessenberg */
int  xo_num_threads(void);
void xo_set_num_threads(int n);   /* OpenMP team size (bench.py sets all host cores: torchrun exports OMP_NUM_THREADS=1) */
void xo_set_sum_order(int o);     /* 0 reference order (default); 1 reversed dot-product / row sums: oracle-vs-oracle drift experiment */

#ifdef __cplusplus
}
#endif
#endif
