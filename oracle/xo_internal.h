/* xo_internal.h -- private structures of the CPU oracle (TEST INFRASTRUCTURE, see xo.h). */
#ifndef XO_INTERNAL_H_
#define XO_INTERNAL_H_

#define _USE_MATH_DEFINES
#define _GNU_SOURCE
#include <math.h>
#include <float.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "xo.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846264338327950288419716939937510582 /* PETSC_PI */
#endif

/* coefficient slots at each quadrature point (StokesCoefficient / LameCoefficient, femixedspace.h:9-22) */
enum { XO_C_ETA = 0, XO_C_FU0 = 1, XO_C_FU1 = 2, XO_C_FU2 = 3, XO_C_FP = 4, XO_C_LAM = 5, XO_NSLOT = 6 };
enum { XO_BC_SOLCX = 0, XO_BC_FIXEDBASE, XO_BC_COMPRESSION, XO_BC_COMPRESSION2, XO_BC_MMS1 };

typedef struct { int n, m; int64_t nnz; int *ia, *ja; double *a; } xo_csr;

typedef struct {
  int     nx, ny, nz;     /* node lattice of this level */
  int     bs;             /* dofs per node */
  xo_csr  A;              /* level operator (finest: alias of A00) */
  int     owns_A;
  double *idiag;          /* Jacobi: 1/diag, zero diag -> 1 (PCSetUp_Jacobi) */
  double  emin, emax;     /* Chebyshev bounds in use */
  double  emin_est, emax_est;
  double *x, *b, *r, *w0, *w1, *w2;   /* level work vectors */
  double *lu; int *piv;   /* dense LU of the coarsest operator */
  double *band; int bw;   /* or, above 8000 rows: banded Cholesky factor (rows x (bw+1), row i holds columns i-bw..i) */
} xo_level;

struct xo_problem_s {
  xo_params prm;
  char   banner[1024];
  char   err[256];
  int    bc_type;
  int    NX, NY, NZ, PX, PY, PZ;
  int64_t nun, npn, nu, np, n, nel, nnz, mnnz;
  int    nbu, nbp, nqp;
  double hu[3], hp[3];
  int   *u_map, *p_map;
  double xi[27][3], wq[27];
  double Nu[27][27], GNuxi[27][27], GNueta[27][27], GNuzeta[27][27];
  double Np[27][8], GNpxi[27][8], GNpeta[27][8], GNpzeta[27][8];
  double *coeff, *coeff_nodal;
  int    nbc, bc_cap; int *bc_idx; double *bc_val; char *isbc;
  int   *ia, *ja; double *a, *a_raw;
  int   *mia, *mja; double *ma;
  double *F;
  double create_seconds;
  /* solver state (xo_solve.c) */
  int    pc_ready;
  xo_solver sopt;
  xo_csr A00, A01, A10, A11;
  int    nlev; xo_level lev[XO_MAX_LEVELS];
  double *mp_lu, *mp_idiag;
  double *idiagA;       /* Jacobi on the full saddle matrix */
  int64_t n_a00_mult, n_a_mult;
  /* GCR workspace */
  double **gcr_V, **gcr_S, *gcr_r, *gcr_val;
  double *fs_tu, *fs_yp;
};

static inline int xo_fail(xo_problem *P, const char *msg) { snprintf(P->err, sizeof(P->err), "%s", msg); return 1; }
static inline double xo_wtime(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }
extern int xo_sum_order;   /* 0 = reference order; 1 = reversed sums (drift experiment, xo_solve.c) */
void xo_solver_free(xo_problem *P);
const char *xo_error(const xo_problem *P);

#endif
