/* xo_solve.c -- CPU oracle, part 2: the PETSc-side algorithms on the solve path.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT (see xo.h).  PETSc is a third-party dependency of the
 * reference (README.md:9, ">= 3.12", unpinned) and is absent from /root/reference and from this
 * image, so the algorithms KSPSolve reaches from exSaddle.c:425 are restated here from PETSc
 * 3.12-3.14 behaviour as catalogued in SURVEY.md App. B: GMRES/FGMRES (B.5), GCR (B.6),
 * PCFIELDSPLIT Schur/UPPER with user Schur-pre (B.2, exSaddle.c:312-321), PCMG V-cycle with DMDA
 * Q1 interpolation and Galerkin coarse operators (B.3), Chebyshev/Jacobi with the GMRES eigenvalue
 * estimate (B.4), bjacobi+ILU(0) (B.7).  Pinned by tests/test_oracle_goldens.py on testref/ *.ref.
 */
#include "xo_internal.h"

int xo_num_threads(void)
{
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void xo_set_num_threads(int n)
{
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* ---------------------------------------------------------------- vectors */
/* Summation-order switch (xo_set_sum_order; 0 = default).  1: dot products and matrix-row sums run from the last term to the
   first.  Used only to measure how far a residual history moves when nothing but the floating-point summation order
   changes (oracle-vs-oracle drift, scripts/oracle_drift.py); the reference's order is the default. */
int xo_sum_order = 0;
void xo_set_sum_order(int o) { xo_sum_order = o; }
static double vdot(int64_t n, const double *x, const double *y)
{
  double s = 0.0; int64_t i;
  if (xo_sum_order == 1) {
#pragma omp parallel for reduction(+ : s) schedule(static)
    for (i = n - 1; i >= 0; --i) s += x[i] * y[i];
    return s;
  }
#pragma omp parallel for reduction(+ : s) schedule(static)
  for (i = 0; i < n; ++i) s += x[i] * y[i];
  return s;
}
static double vnorm(int64_t n, const double *x) { return sqrt(vdot(n, x, x)); }
static void vaxpy(int64_t n, double a, const double *x, double *y)
{
  int64_t i;
#pragma omp parallel for schedule(static)
  for (i = 0; i < n; ++i) y[i] += a * x[i];
}
static void vscale(int64_t n, double a, double *x)
{
  int64_t i;
#pragma omp parallel for schedule(static)
  for (i = 0; i < n; ++i) x[i] *= a;
}
static void vcopy(int64_t n, const double *x, double *y) { memcpy(y, x, sizeof(double) * n); }
static void vzero(int64_t n, double *x) { memset(x, 0, sizeof(double) * n); }
static void vpmult(int64_t n, const double *d, const double *x, double *y)
{
  int64_t i;
#pragma omp parallel for schedule(static)
  for (i = 0; i < n; ++i) y[i] = d[i] * x[i];
}
static double *dvec(int64_t n) { return (double *)calloc((size_t)n, sizeof(double)); }

static void csr_mult(const xo_csr *A, const double *x, double *y) { xo_csr_mult(A->n, A->ia, A->ja, A->a, x, y); }
static void csr_free(xo_csr *A) { free(A->ia); free(A->ja); free(A->a); memset(A, 0, sizeof(*A)); }

/* ------------------------------------------------------- PETSc rander48 */
/* PetscRandom type "rander48", default seed 0x12345678: the 48-bit LCG of drand48
   (X <- 0x5DEECE66D*X + 0xB mod 2^48, X0 = seed<<16 | 0x330E), value X/2^48. */
void xo_rander48(int n, int interval, double *v)
{
  uint64_t X = ((uint64_t)0x12345678u << 16) | 0x330Eu;
  const uint64_t a = 0x5DEECE66DULL, c = 0xBULL, mask = (1ULL << 48) - 1;
  int i;
  for (i = 0; i < n; ++i) {
    double u;
    X = (a * X + c) & mask;
    u = ldexp((double)X, -48);
    v[i] = interval ? 2.0 * u - 1.0 : u;
  }
}

/* ---------------------------------------- eigenvalues of a Hessenberg matrix */
#define SIGN_(a, b) ((b) >= 0.0 ? fabs(a) : -fabs(a))
/* Francis double-shift QR on an upper Hessenberg matrix (EISPACK hqr); used for the Ritz values of
   the GMRES eigen-estimate (KSPComputeEigenvalues_GMRES calls LAPACK hseqr on the same matrix). */
int xo_hess_eig(int n, const double *H, int ldh, double *wr, double *wi)
{
  double *a = (double *)malloc(sizeof(double) * n * n);
  int nn, m, l, k, j, its, i, mmin;
  double z = 0, y, x, w, v, u, t, s, r = 0, q = 0, p = 0, anorm = 0.0;
#define A_(i, j) a[(i) * n + (j)]
  for (i = 0; i < n; ++i) for (j = 0; j < n; ++j) A_(i, j) = (j >= i - 1) ? H[i * ldh + j] : 0.0;
  for (i = 0; i < n; ++i) for (j = (i > 0 ? i - 1 : 0); j < n; ++j) anorm += fabs(A_(i, j));
  nn = n - 1; t = 0.0;
  while (nn >= 0) {
    its = 0;
    do {
      for (l = nn; l >= 1; l--) {
        s = fabs(A_(l - 1, l - 1)) + fabs(A_(l, l));
        if (s == 0.0) s = anorm;
        if ((fabs(A_(l, l - 1)) + s) == s) { A_(l, l - 1) = 0.0; break; }
      }
      x = A_(nn, nn);
      if (l == nn) { wr[nn] = x + t; wi[nn--] = 0.0; }
      else {
        y = A_(nn - 1, nn - 1); w = A_(nn, nn - 1) * A_(nn - 1, nn);
        if (l == nn - 1) {
          p = 0.5 * (y - x); q = p * p + w; z = sqrt(fabs(q)); x += t;
          if (q >= 0.0) {
            z = p + SIGN_(z, p); wr[nn - 1] = wr[nn] = x + z; if (z != 0.0) wr[nn] = x - w / z;
            wi[nn - 1] = wi[nn] = 0.0;
          } else { wr[nn - 1] = wr[nn] = x + p; wi[nn] = z; wi[nn - 1] = -z; }
          nn -= 2;
        } else {
          if (its == 60) { free(a); return 1; }
          if (its == 10 || its == 20) {
            t += x; for (i = 0; i <= nn; ++i) A_(i, i) -= x;
            s = fabs(A_(nn, nn - 1)) + fabs(A_(nn - 1, nn - 2)); y = x = 0.75 * s; w = -0.4375 * s * s;
          }
          ++its;
          for (m = nn - 2; m >= l; m--) {
            z = A_(m, m); r = x - z; s = y - z;
            p = (r * s - w) / A_(m + 1, m) + A_(m, m + 1); q = A_(m + 1, m + 1) - z - r - s; r = A_(m + 2, m + 1);
            s = fabs(p) + fabs(q) + fabs(r); p /= s; q /= s; r /= s;
            if (m == l) break;
            u = fabs(A_(m, m - 1)) * (fabs(q) + fabs(r));
            v = fabs(p) * (fabs(A_(m - 1, m - 1)) + fabs(z) + fabs(A_(m + 1, m + 1)));
            if ((u + v) == v) break;
          }
          for (i = m + 2; i <= nn; ++i) { A_(i, i - 2) = 0.0; if (i != m + 2) A_(i, i - 3) = 0.0; }
          for (k = m; k <= nn - 1; ++k) {
            if (k != m) {
              p = A_(k, k - 1); q = A_(k + 1, k - 1); r = 0.0; if (k != nn - 1) r = A_(k + 2, k - 1);
              if ((x = fabs(p) + fabs(q) + fabs(r)) != 0.0) { p /= x; q /= x; r /= x; }
            }
            if ((s = SIGN_(sqrt(p * p + q * q + r * r), p)) != 0.0) {
              if (k == m) { if (l != m) A_(k, k - 1) = -A_(k, k - 1); }
              else A_(k, k - 1) = -s * x;
              p += s; x = p / s; y = q / s; z = r / s; q /= p; r /= p;
              for (j = k; j <= nn; ++j) {
                p = A_(k, j) + q * A_(k + 1, j);
                if (k != nn - 1) { p += r * A_(k + 2, j); A_(k + 2, j) -= p * z; }
                A_(k + 1, j) -= p * y; A_(k, j) -= p * x;
              }
              mmin = nn < k + 3 ? nn : k + 3;
              for (i = l; i <= mmin; ++i) {
                p = x * A_(i, k) + y * A_(i, k + 1);
                if (k != nn - 1) { p += z * A_(i, k + 2); A_(i, k + 2) -= p * r; }
                A_(i, k + 1) -= p * q; A_(i, k) -= p;
              }
            }
          }
        }
      }
    } while (l < nn - 1);
  }
#undef A_
  free(a);
  return 0;
}

/* ------------------------------------------------------------- sub-blocks */
static void extract_block(const xo_problem *P, int rb, int cb, xo_csr *B)
{
  B->n = rb ? (int)P->np : (int)P->nu; B->m = cb ? (int)P->np : (int)P->nu;
  B->ia = (int *)malloc(sizeof(int) * (B->n + 1));
  B->nnz = xo_submatrix(P, rb, cb, B->ia, NULL, NULL);
  B->ja = (int *)malloc(sizeof(int) * (B->nnz ? B->nnz : 1)); B->a = (double *)malloc(sizeof(double) * (B->nnz ? B->nnz : 1));
  xo_submatrix(P, rb, cb, B->ia, B->ja, B->a);
}

/* --------------------------------------------------- grid transfer (B.3) */
/* DMCreateInterpolation on a DMDA, Q1, refinement ratio 2, non-periodic: fine node f takes coarse node
   f/2 with weight 1 (f even) or coarse nodes (f-1)/2,(f+1)/2 with weight 1/2 each, tensorised; the
   velocity DMDA has dof = NSD so the matrix is MAIJ (same scalar weights for every component). */
static inline int cdim(int nf) { return (nf - 1) / 2 + 1; }   /* DMCoarsen: (n-1)/2+1 */

/* xf += P xc   (MatInterpolateAdd) */
static void prolong_add(const xo_level *F, const xo_level *C, const double *xc, double *xf)
{
  const int bs = F->bs; int64_t nf = (int64_t)F->nx * F->ny * F->nz, f;
#pragma omp parallel for schedule(static)
  for (f = 0; f < nf; ++f) {
    int i = (int)(f % F->nx), j = (int)((f / F->nx) % F->ny), k = (int)(f / ((int64_t)F->nx * F->ny));
    int i0 = i / 2, j0 = j / 2, k0 = k / 2, ni = 1 + (i & 1), nj = 1 + (j & 1), nk = 1 + (k & 1), a, b, c, d;
    double acc[3] = {0, 0, 0};
    for (c = 0; c < nk; ++c) for (b = 0; b < nj; ++b) for (a = 0; a < ni; ++a) {
      double w = (ni == 2 ? 0.5 : 1.0) * (nj == 2 ? 0.5 : 1.0) * (nk == 2 ? 0.5 : 1.0);
      int64_t cn = (i0 + a) + (int64_t)(j0 + b) * C->nx + (int64_t)(k0 + c) * C->nx * C->ny;
      for (d = 0; d < bs; ++d) acc[d] += w * xc[bs * cn + d];
    }
    for (d = 0; d < bs; ++d) xf[bs * f + d] += acc[d];
  }
}
/* bc = P^T rf  (MatRestrict = MatMultTranspose; contributions gathered in ascending fine index) */
static void restrict_to(const xo_level *F, const xo_level *C, const double *rf, double *bc)
{
  const int bs = F->bs; int64_t nc = (int64_t)C->nx * C->ny * C->nz, cidx;
#pragma omp parallel for schedule(static)
  for (cidx = 0; cidx < nc; ++cidx) {
    int I = (int)(cidx % C->nx), J = (int)((cidx / C->nx) % C->ny), K = (int)(cidx / ((int64_t)C->nx * C->ny)), a, b, c, d;
    double acc[3] = {0, 0, 0};
    for (c = -1; c <= 1; ++c) for (b = -1; b <= 1; ++b) for (a = -1; a <= 1; ++a) {
      int i = 2 * I + a, j = 2 * J + b, k = 2 * K + c; double w; int64_t f;
      if (i < 0 || i >= F->nx || j < 0 || j >= F->ny || k < 0 || k >= F->nz) continue;
      w = (a ? 0.5 : 1.0) * (b ? 0.5 : 1.0) * (c ? 0.5 : 1.0);
      f = i + (int64_t)j * F->nx + (int64_t)k * F->nx * F->ny;
      for (d = 0; d < bs; ++d) acc[d] += w * rf[bs * f + d];
    }
    for (d = 0; d < bs; ++d) bc[bs * cidx + d] = acc[d];
  }
}

/* explicit P (dofs x dofs) as CSR, for the Galerkin product */
static void build_P(const xo_level *F, const xo_level *C, xo_csr *Pm)
{
  const int bs = F->bs; int64_t nf = (int64_t)F->nx * F->ny * F->nz, f; int64_t cnt = 0;
  Pm->n = (int)(bs * nf); Pm->m = (int)(bs * (int64_t)C->nx * C->ny * C->nz);
  Pm->ia = (int *)malloc(sizeof(int) * (Pm->n + 1));
  for (f = 0; f < nf; ++f) {
    int i = (int)(f % F->nx), j = (int)((f / F->nx) % F->ny), k = (int)(f / ((int64_t)F->nx * F->ny)), d;
    int e = (1 + (i & 1)) * (1 + (j & 1)) * (1 + (k & 1));
    for (d = 0; d < bs; ++d) { Pm->ia[bs * f + d] = (int)cnt; cnt += e; }
  }
  Pm->ia[Pm->n] = (int)cnt; Pm->nnz = cnt;
  Pm->ja = (int *)malloc(sizeof(int) * cnt); Pm->a = (double *)malloc(sizeof(double) * cnt);
  for (f = 0; f < nf; ++f) {
    int i = (int)(f % F->nx), j = (int)((f / F->nx) % F->ny), k = (int)(f / ((int64_t)F->nx * F->ny));
    int i0 = i / 2, j0 = j / 2, k0 = k / 2, ni = 1 + (i & 1), nj = 1 + (j & 1), nk = 1 + (k & 1), a, b, c, d;
    for (d = 0; d < bs; ++d) {
      int pos = Pm->ia[bs * f + d];
      for (c = 0; c < nk; ++c) for (b = 0; b < nj; ++b) for (a = 0; a < ni; ++a) {
        int64_t cn = (i0 + a) + (int64_t)(j0 + b) * C->nx + (int64_t)(k0 + c) * C->nx * C->ny;
        Pm->ja[pos] = (int)(bs * cn + d);
        Pm->a[pos++] = (ni == 2 ? 0.5 : 1.0) * (nj == 2 ? 0.5 : 1.0) * (nk == 2 ? 0.5 : 1.0);
      }
    }
  }
}

static void csr_transpose(const xo_csr *A, xo_csr *T)
{
  int i, k; int *cnt;
  T->n = A->m; T->m = A->n; T->nnz = A->nnz;
  T->ia = (int *)calloc(T->n + 1, sizeof(int)); T->ja = (int *)malloc(sizeof(int) * A->nnz); T->a = (double *)malloc(sizeof(double) * A->nnz);
  for (k = 0; k < A->nnz; ++k) T->ia[A->ja[k] + 1]++;
  for (i = 0; i < T->n; ++i) T->ia[i + 1] += T->ia[i];
  cnt = (int *)calloc(T->n, sizeof(int));
  for (i = 0; i < A->n; ++i) for (k = A->ia[i]; k < A->ia[i + 1]; ++k) { int c = A->ja[k]; int pos = T->ia[c] + cnt[c]++; T->ja[pos] = i; T->a[pos] = A->a[k]; }
  free(cnt);
}

static int cmp_int2(const void *a, const void *b) { int x = *(const int *)a, y = *(const int *)b; return x < y ? -1 : x > y; }

/* C = A*B by Gustavson's row-wise algorithm; structural (stored zeros propagate, like MatPtAP's symbolic phase);
   columns sorted ascending. */
static int csr_matmat(const xo_csr *A, const xo_csr *B, xo_csr *C)
{
  int i; int64_t tot = 0; int bad = 0;
  C->n = A->n; C->m = B->m;
  C->ia = (int *)malloc(sizeof(int) * (A->n + 1));
  {
    int *len = (int *)malloc(sizeof(int) * A->n);
#pragma omp parallel
    {
      int *mark = (int *)malloc(sizeof(int) * B->m); int j;
      for (j = 0; j < B->m; ++j) mark[j] = -1;
#pragma omp for schedule(dynamic, 256)
      for (i = 0; i < A->n; ++i) {
        int k, l, c = 0;
        for (k = A->ia[i]; k < A->ia[i + 1]; ++k) { int r = A->ja[k]; for (l = B->ia[r]; l < B->ia[r + 1]; ++l) if (mark[B->ja[l]] != i) { mark[B->ja[l]] = i; c++; } }
        len[i] = c;
      }
      free(mark);
    }
    for (i = 0; i < A->n; ++i) { C->ia[i] = (int)tot; tot += len[i]; if (tot >= INT32_MAX) bad = 1; }
    C->ia[A->n] = (int)tot; C->nnz = tot; free(len);
    if (bad) return 1;
  }
  C->ja = (int *)malloc(sizeof(int) * (tot ? tot : 1)); C->a = (double *)malloc(sizeof(double) * (tot ? tot : 1));
#pragma omp parallel
  {
    int *mark = (int *)malloc(sizeof(int) * B->m); double *acc = (double *)calloc(B->m, sizeof(double)); int j;
    for (j = 0; j < B->m; ++j) mark[j] = -1;
#pragma omp for schedule(dynamic, 256)
    for (i = 0; i < A->n; ++i) {
      int k, l, c = 0, *cj = &C->ja[C->ia[i]]; double *ca = &C->a[C->ia[i]];
      for (k = A->ia[i]; k < A->ia[i + 1]; ++k) {
        int r = A->ja[k]; double av = A->a[k];
        for (l = B->ia[r]; l < B->ia[r + 1]; ++l) {
          int col = B->ja[l];
          if (mark[col] != i) { mark[col] = i; cj[c++] = col; acc[col] = av * B->a[l]; }
          else acc[col] += av * B->a[l];
        }
      }
      qsort(cj, c, sizeof(int), cmp_int2);
      for (k = 0; k < c; ++k) ca[k] = acc[cj[k]];
    }
    free(mark); free(acc);
  }
  return 0;
}

/* ------------------------------------------------------------ dense LU */
static int dense_lu(int n, double *a, int *piv)
{
  int i, j, k;
  for (k = 0; k < n; ++k) {
    int p = k; double mx = fabs(a[(size_t)k * n + k]);
    for (i = k + 1; i < n; ++i) if (fabs(a[(size_t)i * n + k]) > mx) { mx = fabs(a[(size_t)i * n + k]); p = i; }
    if (mx == 0.0) return 1;
    piv[k] = p;
    if (p != k) for (j = 0; j < n; ++j) { double t = a[(size_t)k * n + j]; a[(size_t)k * n + j] = a[(size_t)p * n + j]; a[(size_t)p * n + j] = t; }
#pragma omp parallel for schedule(static) private(j) if (n - k > 256)
    for (i = k + 1; i < n; ++i) {
      double l = a[(size_t)i * n + k] / a[(size_t)k * n + k];
      a[(size_t)i * n + k] = l;
      for (j = k + 1; j < n; ++j) a[(size_t)i * n + j] -= l * a[(size_t)k * n + j];
    }
  }
  return 0;
}
static void dense_lu_solve(int n, const double *a, const int *piv, const double *b, double *x)
{
  int i, j;
  for (i = 0; i < n; ++i) x[i] = b[i];
  for (i = 0; i < n; ++i) { int p = piv[i]; if (p != i) { double t = x[i]; x[i] = x[p]; x[p] = t; } }
  for (i = 0; i < n; ++i) { double s = x[i]; for (j = 0; j < i; ++j) s -= a[(size_t)i * n + j] * x[j]; x[i] = s; }
  for (i = n - 1; i >= 0; --i) { double s = x[i]; for (j = i + 1; j < n; ++j) s -= a[(size_t)i * n + j] * x[j]; x[i] = s / a[(size_t)i * n + i]; }
}

/* ------------------------------------------------------ banded Cholesky */
/* Exact sparse direct solve of a large coarsest operator (PETSc: UMFPACK LU, abf.opts:16, on the 14 739 / 107 811 rows abf.opts'
   3 levels leave at 32^3 / 64^3).  The Galerkin operator is symmetric positive definite with a lattice band, so a banded
   L L^T factorisation is an exact LU up to rounding.  B[i*(bw+1) + (j-i+bw)] = entry (i,j), i-bw <= j <= i. */
static int band_cholesky(int n, int bw, double *B)
{
  const int w = bw + 1; int k;
  double *col = (double *)malloc(sizeof(double) * (size_t)w);
  for (k = 0; k < n; ++k) {
    const int m = (k + bw < n - 1 ? k + bw : n - 1) - k;   /* rows k+1 .. k+m hold column k */
    double d = B[(size_t)k * w + bw]; int i;
    if (!(d > 0.0)) { free(col); return 1; }
    d = sqrt(d); B[(size_t)k * w + bw] = d;
    for (i = 1; i <= m; ++i) { double *p = &B[(size_t)(k + i) * w + (bw - i)]; *p /= d; col[i] = *p; }
#pragma omp parallel for schedule(static) if (m > 64)
    for (i = 1; i <= m; ++i) {          /* trailing update: row k+i, columns k+1 .. k+i */
      const double lik = col[i]; double *row = &B[(size_t)(k + i) * w + (bw - i)]; int j;
      for (j = 1; j <= i; ++j) row[j] -= lik * col[j];
    }
  }
  free(col);
  return 0;
}
static void band_cholesky_solve(int n, int bw, const double *B, const double *b, double *x)
{
  const int w = bw + 1; int i, j;
  for (i = 0; i < n; ++i) {             /* L y = b */
    const int j0 = i - bw > 0 ? i - bw : 0; const double *row = &B[(size_t)i * w + (bw - i)]; double s = b[i];
    for (j = j0; j < i; ++j) s -= row[j] * x[j];
    x[i] = s / row[i];
  }
  for (i = n - 1; i >= 0; --i) {        /* L^T x = y, column sweep */
    const int j0 = i - bw > 0 ? i - bw : 0; const double *row = &B[(size_t)i * w + (bw - i)]; const double xi = x[i] / row[i];
    x[i] = xi;
    for (j = j0; j < i; ++j) x[j] -= row[j] * xi;
  }
}

/* -------------------------------------------------------------- ILU(0) */
/* MatLUFactorNumeric_SeqAIJ with ILU(0) pattern, natural ordering, no shift: row-wise IKJ, the pivot is
   stored inverted and the multiplier is a_ik * (1/a_kk).  lu[] shares Mp's pattern; lu[diag] = 1/pivot. */
int xo_ilu0(int n, const int *ia, const int *ja, const double *a, double *lu)
{
  int i, k, *diag = (int *)malloc(sizeof(int) * n); double *w = (double *)calloc(n, sizeof(double));
  memcpy(lu, a, sizeof(double) * ia[n]);
  for (i = 0; i < n; ++i) { diag[i] = -1; for (k = ia[i]; k < ia[i + 1]; ++k) if (ja[k] == i) diag[i] = k; if (diag[i] < 0) { free(diag); free(w); return 1; } }
  for (i = 0; i < n; ++i) {
    for (k = ia[i]; k < ia[i + 1]; ++k) w[ja[k]] = lu[k];
    for (k = ia[i]; k < diag[i]; ++k) {
      int r = ja[k], l; double mult = w[r];
      if (mult != 0.0) {
        mult = mult * lu[diag[r]]; w[r] = mult;
        for (l = diag[r] + 1; l < ia[r + 1]; ++l) w[ja[l]] -= mult * lu[l];   /* fill outside row i's pattern is dropped below */
      }
    }
    if (w[i] == 0.0) { free(diag); free(w); return 2; }   /* zero pivot */
    for (k = ia[i]; k < ia[i + 1]; ++k) { lu[k] = w[ja[k]]; }
    lu[diag[i]] = 1.0 / w[i];
    /* clear the work row, including dropped fill */
    for (k = ia[i]; k < diag[i]; ++k) { int r = ja[k], l; for (l = diag[r] + 1; l < ia[r + 1]; ++l) w[ja[l]] = 0.0; }
    for (k = ia[i]; k < ia[i + 1]; ++k) w[ja[k]] = 0.0;
  }
  free(diag); free(w);
  return 0;
}
void xo_ilu0_solve(int n, const int *ia, const int *ja, const double *lu, const double *b, double *x)
{
  int i, k;
  for (i = 0; i < n; ++i) { double s = b[i]; for (k = ia[i]; k < ia[i + 1] && ja[k] < i; ++k) s -= lu[k] * x[ja[k]]; x[i] = s; }
  for (i = n - 1; i >= 0; --i) {
    double s = x[i], d = 1.0; for (k = ia[i + 1] - 1; k >= ia[i] && ja[k] > i; --k) s -= lu[k] * x[ja[k]];
    d = lu[k];   /* k now at the diagonal: stored 1/pivot */
    x[i] = s * d;
  }
}

/* ------------------------------------------------------------- Jacobi */
/* PCSetUp_Jacobi: inverse diagonal, zero entries replaced by 1.0 (App. B.1) */
static double *jacobi_idiag(const xo_csr *A)
{
  double *d = dvec(A->n); int i, k;
  for (i = 0; i < A->n; ++i) { double v = 0.0; for (k = A->ia[i]; k < A->ia[i + 1]; ++k) if (A->ja[k] == i) v = A->a[k]; d[i] = v == 0.0 ? 1.0 : 1.0 / v; }
  return d;
}

/* ------------------------------------------ GMRES(m) kernel shared by callers */
typedef struct {
  int64_t n; int restart, max_it; double rtol, atol, dtol;
  int flexible, right;
  void (*amult)(void *, const double *, double *); void *actx;
  void (*pc)(void *, const double *, double *); void *pctx;
  int sample_stop;
  /* outputs */
  int its, reason, nhist; double *hist; int hist_cap;
  double *hes; int hes_n;   /* Hessenberg of the last cycle, (restart+1) x restart, column-major by iteration */
} gmres_t;

/* KSPSolve_GMRES / KSPSolve_FGMRES: classical Gram-Schmidt without refinement, Givens QR (App. B.5).
   left:  iterate on M^-1 A, preconditioned residual norm.   right/flexible: z_j = M^-1 v_j, w = A z_j. */
static void gmres_solve(gmres_t *g, const double *b, double *x)
{
  const int64_t n = g->n; const int m = g->restart;
  double **V = (double **)calloc(m + 1, sizeof(double *)), **Z = (double **)calloc(m + 1, sizeof(double *));
  double *hh = (double *)calloc((size_t)(m + 1) * m, sizeof(double));   /* rotated H */
  double *hes = (double *)calloc((size_t)(m + 1) * m, sizeof(double));  /* original H */
  double *cs = dvec(m + 1), *sn = dvec(m + 1), *rs = dvec(m + 1), *y = dvec(m + 1), *hcol = dvec(m + 1);
  double *t1 = dvec(n), *t2 = dvec(n);
  double rnorm0 = 0.0, ttol = 0.0; int it, j, k;
  g->its = 0; g->reason = 0; g->nhist = 0;
  V[0] = dvec(n);
  while (!g->reason) {
    double res;
    /* initial residual of the cycle (KSPInitialResidual) */
    g->amult(g->actx, x, t1);
    for (int64_t i = 0; i < n; ++i) t1[i] = b[i] - t1[i];
    if (!g->flexible && !g->right && g->pc) g->pc(g->pctx, t1, V[0]); else vcopy(n, t1, V[0]);
    res = vnorm(n, V[0]);
    if (g->its == 0) { rnorm0 = res; ttol = fmax(g->rtol * rnorm0, g->atol); }
    if (g->nhist < g->hist_cap) { if (g->nhist == g->its) g->nhist++; g->hist[g->its] = res; }
    if (res == 0.0) { g->reason = 3; break; }
    if (res <= ttol) { g->reason = (res <= g->atol && g->atol >= g->rtol * rnorm0) ? 3 : 2; break; }
    if (g->its >= g->max_it) { g->reason = -3; break; }
    vscale(n, 1.0 / res, V[0]);
    rs[0] = res;
    it = 0;
    memset(hes, 0, sizeof(double) * (size_t)(m + 1) * m);
    while (!g->reason && it < m && g->its < g->max_it) {
      double *w, tt;
      if (!V[it + 1]) V[it + 1] = dvec(n);
      w = V[it + 1];
      if (g->flexible) {                 /* FGMRES: z_j = M^-1 v_j ; w = A z_j */
        if (!Z[it]) Z[it] = dvec(n);
        g->pc(g->pctx, V[it], Z[it]); g->amult(g->actx, Z[it], w);
      } else if (g->right) {             /* right: w = A M^-1 v_j */
        if (g->pc) { g->pc(g->pctx, V[it], t2); g->amult(g->actx, t2, w); } else g->amult(g->actx, V[it], w);
      } else {                           /* left: w = M^-1 A v_j */
        if (g->pc) { g->amult(g->actx, V[it], t2); g->pc(g->pctx, t2, w); } else g->amult(g->actx, V[it], w);
      }
      /* classical Gram-Schmidt, one pass: h = V^T w (VecMDot); w -= V h (VecMAXPY) */
      for (j = 0; j <= it; ++j) hcol[j] = vdot(n, w, V[j]);
      for (j = 0; j <= it; ++j) vaxpy(n, -hcol[j], V[j], w);
      tt = vnorm(n, w);
      if (tt != 0.0) vscale(n, 1.0 / tt, w);
      hcol[it + 1] = tt;
      for (j = 0; j <= it + 1; ++j) hes[(size_t)it * (m + 1) + j] = hcol[j];
      /* KSPGMRESUpdateHessenberg: apply previous rotations, form the new one */
      for (j = 0; j < it; ++j) { double t = hcol[j]; hcol[j] = cs[j] * t + sn[j] * hcol[j + 1]; hcol[j + 1] = -sn[j] * t + cs[j] * hcol[j + 1]; }
      { double t = sqrt(hcol[it] * hcol[it] + hcol[it + 1] * hcol[it + 1]);
        if (t == 0.0) { g->reason = -5; break; }
        cs[it] = hcol[it] / t; sn[it] = hcol[it + 1] / t;
        rs[it + 1] = -sn[it] * rs[it]; rs[it] = cs[it] * rs[it];
        hcol[it] = cs[it] * hcol[it] + sn[it] * hcol[it + 1]; hcol[it + 1] = 0.0;
        res = fabs(rs[it + 1]); }
      for (j = 0; j <= it; ++j) hh[(size_t)it * (m + 1) + j] = hcol[j];
      it++; g->its++;
      if (g->its < g->hist_cap) { g->hist[g->its] = res; g->nhist = g->its + 1; }
      if (res <= ttol) g->reason = (res <= g->atol && g->atol >= g->rtol * rnorm0) ? 3 : 2;
      else if (res >= g->dtol * rnorm0) g->reason = -4;
      else if (g->sample_stop > 0 && g->its >= g->sample_stop) g->reason = -3;
    }
    g->hes_n = it;
    if (g->hes) memcpy(g->hes, hes, sizeof(double) * (size_t)(m + 1) * m);
    /* KSPGMRESBuildSoln: back substitution, x += V y (left/right: through M^-1 for right) or Z y (flexible) */
    if (it > 0) {
      for (k = it - 1; k >= 0; --k) {
        double t = rs[k];
        for (j = k + 1; j < it; ++j) t -= hh[(size_t)j * (m + 1) + k] * y[j];
        y[k] = t / hh[(size_t)k * (m + 1) + k];
      }
      if (g->flexible) { for (j = 0; j < it; ++j) vaxpy(n, y[j], Z[j], x); }
      else if (g->right && g->pc) { vzero(n, t1); for (j = 0; j < it; ++j) vaxpy(n, y[j], V[j], t1); g->pc(g->pctx, t1, t2); vaxpy(n, 1.0, t2, x); }
      else { for (j = 0; j < it; ++j) vaxpy(n, y[j], V[j], x); }
    }
    if (!g->reason && g->its >= g->max_it) g->reason = -3;
  }
  for (j = 0; j <= m; ++j) { free(V[j]); free(Z[j]); }
  free(V); free(Z); free(hh); free(hes); free(cs); free(sn); free(rs); free(y); free(hcol); free(t1); free(t2);
}

/* ------------------------------------------------------- MG on A00 (B.3/B.4) */
static void lev_amult(void *ctx, const double *x, double *y) { csr_mult(&((xo_level *)ctx)->A, x, y); }
static void lev_jacobi(void *ctx, const double *x, double *y) { xo_level *L = (xo_level *)ctx; vpmult(L->A.n, L->idiag, x, y); }

/* KSPChebyshev eigen-estimate: esteig_steps of left-Jacobi GMRES on a noisy rhs, Ritz values of the Hessenberg */
static int cheb_estimate(xo_level *L, const xo_solver *s)
{
  const int n = L->A.n, m = s->esteig_steps; gmres_t g; double *b = dvec(n), *x = dvec(n), hist[64];
  double *hes = (double *)calloc((size_t)(m + 1) * m, sizeof(double)), *H, wr[64], wi[64]; int i, j, k;
  xo_rander48(n, s->noise, b);
  memset(&g, 0, sizeof(g));
  g.n = n; g.restart = m; g.max_it = m; g.rtol = 1e-12; g.atol = 1e-50; g.dtol = 1e30; g.flexible = 0; g.right = 0;
  g.amult = lev_amult; g.actx = L; g.pc = lev_jacobi; g.pctx = L; g.hist = hist; g.hist_cap = 64; g.hes = hes;
  gmres_solve(&g, b, x);
  k = g.hes_n;
  if (k < 1) { free(b); free(x); free(hes); return 1; }
  H = (double *)calloc((size_t)k * k, sizeof(double));
  for (j = 0; j < k; ++j) for (i = 0; i < k; ++i) H[i * k + j] = hes[(size_t)j * (m + 1) + i];
  if (xo_hess_eig(k, H, k, wr, wi)) { free(H); free(b); free(x); free(hes); return 1; }
  L->emin_est = wr[0]; L->emax_est = wr[0];
  for (i = 1; i < k; ++i) { if (wr[i] < L->emin_est) L->emin_est = wr[i]; if (wr[i] > L->emax_est) L->emax_est = wr[i]; }
  L->emin = s->esteig[0] * L->emin_est + s->esteig[1] * L->emax_est;
  L->emax = s->esteig[2] * L->emin_est + s->esteig[3] * L->emax_est;
  free(H); free(b); free(x); free(hes);
  return 0;
}

/* KSPSolve_Chebyshev with Jacobi, nonzero initial guess, no norm (App. B.4): first correction + (max_it-1) passes.
   x_is_zero skips the A*0 product of the down-smoother (bitwise identical: b - A*0 = b). */
static void cheb_smooth(xo_problem *P, xo_level *L, const double *b, double *x, int its, int x_is_zero, int fine)
{
  const int64_t n = L->A.n; int64_t i; int it;
  const double scale = 2.0 / (L->emax + L->emin), alpha = 1.0 - scale * L->emin, mu = 1.0 / alpha, omegaprod = 2.0 / alpha;
  double ckm1 = 1.0, ck = mu, *r = L->r, *pkm1 = L->w0, *pk = L->w1, *pkp1 = L->w2, *tmp;
  if (its < 1) return;
  if (x_is_zero) vcopy(n, b, r);
  else { csr_mult(&L->A, x, r); if (fine) P->n_a00_mult++;
#pragma omp parallel for schedule(static)
    for (i = 0; i < n; ++i) r[i] = b[i] - r[i]; }
  vcopy(n, x, pkm1);
#pragma omp parallel for schedule(static)
  for (i = 0; i < n; ++i) pk[i] = pkm1[i] + scale * (L->idiag[i] * r[i]);   /* p[k] = x + scale*B r */
  for (it = 1; it < its; ++it) {
    const double ckp1 = 2.0 * mu * ck - ckm1, omega = omegaprod * ck / ckp1;
    csr_mult(&L->A, pk, r); if (fine) P->n_a00_mult++;
#pragma omp parallel for schedule(static)
    for (i = 0; i < n; ++i) {
      double ri = b[i] - r[i];
      pkp1[i] = (1.0 - omega) * pkm1[i] + omega * pk[i] + omega * scale * (L->idiag[i] * ri);   /* VecAXPBYPCZ */
    }
    ckm1 = ck; ck = ckp1;
    tmp = pkm1; pkm1 = pk; pk = pkp1; pkp1 = tmp;
  }
  vcopy(n, pk, x);
}

static void mg_cycle(xo_problem *P, int l)
{
  xo_level *L = &P->lev[l];
  if (l == 0) {   /* coarse: preonly + LU */
    if (L->band) band_cholesky_solve(L->A.n, L->bw, L->band, L->b, L->x);
    else dense_lu_solve(L->A.n, L->lu, L->piv, L->b, L->x);
    return;
  }
  {
    xo_level *C = &P->lev[l - 1]; const int64_t n = L->A.n; int64_t i; const int fine = (l == P->nlev - 1);
    cheb_smooth(P, L, L->b, L->x, P->sopt.cheb_its, 1, fine);          /* pre-smooth, x = 0 on entry */
    csr_mult(&L->A, L->x, L->r); if (fine) P->n_a00_mult++;
#pragma omp parallel for schedule(static)
    for (i = 0; i < n; ++i) L->r[i] = L->b[i] - L->r[i];               /* residual */
    restrict_to(L, C, L->r, C->b);                                     /* MatRestrict */
    vzero(C->A.n, C->x);
    mg_cycle(P, l - 1);
    prolong_add(L, C, C->x, L->x);                                     /* MatInterpolateAdd */
    cheb_smooth(P, L, L->b, L->x, P->sopt.cheb_its, 0, fine);          /* post-smooth */
  }
}

int xo_vcycle(xo_problem *P, const double *b, double *x)
{
  xo_level *L = &P->lev[P->nlev - 1];
  vcopy(L->A.n, b, L->b); vzero(L->A.n, L->x);
  mg_cycle(P, P->nlev - 1);
  vcopy(L->A.n, L->x, x);
  return 0;
}

int xo_mg_setup(xo_problem *P, int levels)
{
  int l; const int bs = P->prm.nsd;
  if (levels < 1 || levels > XO_MAX_LEVELS) return xo_fail(P, "bad MG level count");
  if (!P->A00.ia) extract_block(P, 0, 0, &P->A00);
  P->nlev = levels;
  { xo_level *L = &P->lev[levels - 1]; L->nx = P->NX; L->ny = P->NY; L->nz = P->NZ; L->bs = bs; L->A = P->A00; L->owns_A = 0; }
  for (l = levels - 2; l >= 0; --l) {
    xo_level *F = &P->lev[l + 1], *C = &P->lev[l]; xo_csr Pm, R, AP;
    if ((F->nx - 1) % 2 || (F->ny - 1) % 2 || (F->nz > 1 && (F->nz - 1) % 2)) return xo_fail(P, "DMCoarsen: (n-1) not divisible by 2 for the requested MG levels");
    C->nx = cdim(F->nx); C->ny = cdim(F->ny); C->nz = F->nz > 1 ? cdim(F->nz) : 1; C->bs = bs;
    if (C->nx < 2 || C->ny < 2 || (F->nz > 1 && C->nz < 2)) return xo_fail(P, "too many MG levels for this mesh");
    build_P(F, C, &Pm); csr_transpose(&Pm, &R);
    if (csr_matmat(&F->A, &Pm, &AP)) return xo_fail(P, "Galerkin product too large");
    if (csr_matmat(&R, &AP, &C->A)) return xo_fail(P, "Galerkin product too large");   /* A_c = P^T (A P) */
    C->owns_A = 1;
    csr_free(&Pm); csr_free(&R); csr_free(&AP);
  }
  for (l = 0; l < levels; ++l) {
    xo_level *L = &P->lev[l]; const int n = L->A.n;
    L->x = dvec(n); L->b = dvec(n); L->r = dvec(n); L->w0 = dvec(n); L->w1 = dvec(n); L->w2 = dvec(n);
    L->idiag = jacobi_idiag(&L->A);
  }
  { /* coarsest: dense LU (PETSc: UMFPACK sparse LU, abf.opts:16) */
    xo_level *L = &P->lev[0]; const int n = L->A.n; int i, k;
    if (n > 8000) {   /* banded Cholesky (the lattice band of the 27-point block stencil) */
      int bw = 0;
      for (i = 0; i < n; ++i) for (k = L->A.ia[i]; k < L->A.ia[i + 1]; ++k) if (i - L->A.ja[k] > bw) bw = i - L->A.ja[k];
      if ((double)n * (bw + 1) * 8.0 > 16e9) return xo_fail(P, "coarsest MG level too large for the oracle's banded Cholesky (use more levels)");
      L->bw = bw; L->band = (double *)calloc((size_t)n * (bw + 1), sizeof(double));
      for (i = 0; i < n; ++i) for (k = L->A.ia[i]; k < L->A.ia[i + 1]; ++k) if (L->A.ja[k] <= i) L->band[(size_t)i * (bw + 1) + (L->A.ja[k] - i + bw)] = L->A.a[k];
      if (band_cholesky(n, bw, L->band)) return xo_fail(P, "coarse operator is not positive definite");
      return 0;
    }
    L->lu = (double *)calloc((size_t)n * n, sizeof(double)); L->piv = (int *)malloc(sizeof(int) * n);
    for (i = 0; i < n; ++i) for (k = L->A.ia[i]; k < L->A.ia[i + 1]; ++k) L->lu[(size_t)i * n + L->A.ja[k]] = L->A.a[k];
    if (dense_lu(n, L->lu, L->piv)) return xo_fail(P, "singular coarse operator");
  }
  return 0;
}

int xo_mg_level_csr(const xo_problem *P, int level, int *n, const int **ia, const int **ja, const double **a)
{
  if (level < 0 || level >= P->nlev) return 1;
  *n = P->lev[level].A.n; *ia = P->lev[level].A.ia; *ja = P->lev[level].A.ja; *a = P->lev[level].A.a;
  return 0;
}
void xo_mg_prolong_add(const xo_problem *P, int lc, const double *xc, double *xf) { prolong_add(&P->lev[lc + 1], &P->lev[lc], xc, xf); }
void xo_mg_restrict(const xo_problem *P, int lc, const double *rf, double *bc) { restrict_to(&P->lev[lc + 1], &P->lev[lc], rf, bc); }

/* ------------------------------------------------------------ GCR (B.6) */
/* KSPSolve_GCR on A00 with PCMG as right preconditioner; returns iterations */
static int gcr_solve(xo_problem *P, const double *b, double *x)
{
  const xo_solver *s = &P->sopt; const int64_t n = P->nu; const int m = s->u_restart;
  double *r = P->gcr_r, norm_r, rnorm0, ttol; int its = 0, k, i, done = 0;
  double **V = P->gcr_V, **S = P->gcr_S, *val = P->gcr_val;
  vzero(n, x);
  vcopy(n, b, r);   /* r = b - A*0 */
  norm_r = vnorm(n, r); rnorm0 = norm_r; ttol = fmax(s->u_rtol * rnorm0, 1e-50);
  if (norm_r <= ttol) return 0;
  while (!done && its < s->u_max_it) {
    for (k = 0; k < m; ++k) {
      double *v, *sv, r_dot_v, nrm;
      if (!V[k]) { V[k] = dvec(n); S[k] = dvec(n); }
      v = V[k]; sv = S[k];
      xo_vcycle(P, r, sv);                               /* s = B^-1 r */
      csr_mult(&P->A00, sv, v); P->n_a00_mult++;         /* v = A s */
      for (i = 0; i < k; ++i) val[i] = -vdot(n, v, V[i]);   /* VecMDot, negated */
      for (i = 0; i < k; ++i) vaxpy(n, val[i], V[i], v);
      for (i = 0; i < k; ++i) vaxpy(n, val[i], S[i], sv);
      r_dot_v = vdot(n, r, v); nrm = sqrt(vdot(n, v, v));   /* VecDotNorm2 */
      r_dot_v = r_dot_v / nrm;
      vscale(n, 1.0 / nrm, v); vscale(n, 1.0 / nrm, sv);
      vaxpy(n, r_dot_v, sv, x); vaxpy(n, -r_dot_v, v, r);
      norm_r = vnorm(n, r);
      its++;
      if (norm_r <= ttol) { done = 1; break; }
      if (norm_r >= 1e4 * rnorm0) { done = 1; break; }
      if (its >= s->u_max_it) { done = 1; break; }
    }
  }
  return its;
}

/* --------------------------------------- PCApply_FieldSplit_Schur, UPPER (B.2) */
int xo_pc_apply(xo_problem *P, const double *rr, double *z, int *inner_its)
{
  const int64_t nu = P->nu, np = P->np; int64_t i; int its;
  double *yp = z + nu, *tu = P->fs_tu;
  /* y_p = KSP(S, Mp) x_p with ksp preonly: PC(Mp)^-1 x_p */
  if (P->sopt.p_pc == XO_PPC_ILU0) xo_ilu0_solve((int)np, P->mia, P->mja, P->mp_lu, rr + nu, yp);
  else vpmult(np, P->mp_idiag, rr + nu, yp);
  /* t_u = x_u - A01 y_p */
  csr_mult(&P->A01, yp, tu);
#pragma omp parallel for schedule(static)
  for (i = 0; i < nu; ++i) tu[i] = rr[i] + (-1.0) * tu[i];
  its = gcr_solve(P, tu, z);
  if (inner_its) *inner_its = its;
  return 0;
}

typedef struct { xo_problem *P; xo_result *res; } outer_ctx;
static void outer_amult(void *ctx, const double *x, double *y) { xo_problem *P = ((outer_ctx *)ctx)->P; xo_A_mult(P, x, y); P->n_a_mult++; }
static void outer_pc_abf(void *ctx, const double *x, double *y)
{
  outer_ctx *c = (outer_ctx *)ctx; int its;
  xo_pc_apply(c->P, x, y, &its);
  if (c->res && c->res->n_inner < 2048) c->res->inner_its[c->res->n_inner++] = its;
}
static void outer_pc_jacobi(void *ctx, const double *x, double *y) { xo_problem *P = ((outer_ctx *)ctx)->P; vpmult(P->n, P->idiagA, x, y); }

void xo_solver_init(xo_solver *s)
{
  memset(s, 0, sizeof(*s));
  s->ksp_type = XO_KSP_GMRES; s->pc_type = XO_PC_NONE; s->pc_side = -1;
  s->rtol = 1e-5; s->atol = 1e-50; s->dtol = 1e4; s->max_it = 10000; s->restart = 30;   /* App. B.1 */
  s->u_rtol = 1e-5; s->u_max_it = 10000; s->u_restart = 30;
  s->mg_levels = 1; s->cheb_its = 2; s->esteig[0] = 0; s->esteig[1] = 0.1; s->esteig[2] = 0; s->esteig[3] = 1.1;
  s->esteig_steps = 10; s->noise = 0; s->p_pc = XO_PPC_ILU0;
}
/* abf.opts:2-16 */
void xo_solver_abf(xo_solver *s)
{
  xo_solver_init(s);
  s->ksp_type = XO_KSP_FGMRES; s->pc_type = XO_PC_ABF; s->pc_side = XO_SIDE_RIGHT;
  s->u_rtol = 1e-2; s->mg_levels = 3; s->cheb_its = 8;
  s->esteig[0] = 0; s->esteig[1] = 0.2; s->esteig[2] = 0; s->esteig[3] = 1.1;
  s->p_pc = XO_PPC_ILU0;
}

void xo_solver_free(xo_problem *P)
{
  int l, k;
  for (l = 0; l < P->nlev; ++l) {
    xo_level *L = &P->lev[l];
    if (L->owns_A) csr_free(&L->A);
    free(L->idiag); free(L->x); free(L->b); free(L->r); free(L->w0); free(L->w1); free(L->w2); free(L->lu); free(L->piv); free(L->band);
    memset(L, 0, sizeof(*L));
  }
  P->nlev = 0;
  csr_free(&P->A00); csr_free(&P->A01); csr_free(&P->A10); csr_free(&P->A11);
  free(P->mp_lu); free(P->mp_idiag); free(P->idiagA); P->mp_lu = P->mp_idiag = P->idiagA = NULL;
  if (P->gcr_V) { for (k = 0; k < P->sopt.u_restart; ++k) { free(P->gcr_V[k]); free(P->gcr_S[k]); } }
  free(P->gcr_V); free(P->gcr_S); free(P->gcr_r); free(P->gcr_val); P->gcr_V = P->gcr_S = NULL; P->gcr_r = P->gcr_val = NULL;
  free(P->fs_tu); free(P->fs_yp); P->fs_tu = P->fs_yp = NULL;
  P->pc_ready = 0;
}

/* KSPSetUp for the chosen tree (exSaddle.c:304-322, 405-422) */
int xo_pc_setup(xo_problem *P, const xo_solver *s, xo_result *res)
{
  int l;
  xo_solver_free(P);
  P->sopt = *s; P->sopt.max_outer_sample = 0;
  if (s->pc_type == XO_PC_JACOBI) {
    xo_csr A; A.n = (int)P->n; A.m = (int)P->n; A.ia = P->ia; A.ja = P->ja; A.a = P->a; A.nnz = P->nnz;
    P->idiagA = jacobi_idiag(&A);
  } else if (s->pc_type == XO_PC_ABF) {
    extract_block(P, 0, 0, &P->A00); extract_block(P, 0, 1, &P->A01);
    if (xo_mg_setup(P, s->mg_levels)) return 1;
    for (l = 1; l < P->nlev; ++l) {
      xo_level *L = &P->lev[l];
      if (s->n_cheb_fixed > 0) {   /* -ksp_chebyshev_eigenvalues emin,emax per level (level 1 = first above coarse) */
        if (l - 1 >= s->n_cheb_fixed) return xo_fail(P, "explicit Chebyshev bounds missing for a level");
        L->emin = s->cheb_emin[l - 1]; L->emax = s->cheb_emax[l - 1]; L->emin_est = L->emax_est = NAN;
      } else if (cheb_estimate(L, s)) return xo_fail(P, "Chebyshev eigenvalue estimate failed");
    }
    if (s->p_pc == XO_PPC_ILU0) {
      const double *ma = P->ma; double *blk = NULL;
      P->mp_lu = (double *)malloc(sizeof(double) * P->mnnz);
      if (s->p_blocks > 1) {
        /* PCBJACOBI with one block per rank (SURVEY 8e caveat 2): rank r owns the pressure planes of its element layers
           [k0,k1) (mz/N each, remainder to the low ranks; the last rank also owns the top plane), i.e. the product's slab
           rule.  ILU(0) of the block-diagonal part = ILU(0) of Mp with the cross-block entries set to zero (no update can
           reach them: every l_ik u_kj with (i,j) across blocks has one cross-block factor). */
        const int mz = P->PZ - 1, nb = s->p_blocks, q = mz / nb, rr = mz % nb; const int64_t pn = (int64_t)P->PX * P->PY;
        int64_t i; int k;
        if (P->prm.nsd != 3 || mz < nb) return xo_fail(P, "p_blocks needs a 3-D mesh with at least one element layer per block");
        blk = (double *)malloc(sizeof(double) * P->mnnz);
        #define XO_BLOCK_OF(plane) ((plane) >= mz ? nb - 1 : ((plane) < rr * (q + 1) ? (plane) / (q + 1) : rr + ((plane) - rr * (q + 1)) / q))
        for (i = 0; i < P->np; ++i) {
          const int bi = XO_BLOCK_OF((int)(i / pn));
          for (k = P->mia[i]; k < P->mia[i + 1]; ++k) blk[k] = XO_BLOCK_OF((int)(P->mja[k] / pn)) == bi ? P->ma[k] : 0.0;
        }
        #undef XO_BLOCK_OF
        ma = blk;
      }
      if (xo_ilu0((int)P->np, P->mia, P->mja, ma, P->mp_lu)) { free(blk); return xo_fail(P, "ILU(0) of Mpscaled failed"); }
      free(blk);
    } else {
      xo_csr M; M.n = (int)P->np; M.m = M.n; M.ia = P->mia; M.ja = P->mja; M.a = P->ma; M.nnz = P->mnnz;
      P->mp_idiag = jacobi_idiag(&M);
    }
    P->gcr_V = (double **)calloc(s->u_restart, sizeof(double *)); P->gcr_S = (double **)calloc(s->u_restart, sizeof(double *));
    P->gcr_r = dvec(P->nu); P->gcr_val = dvec(s->u_restart + 1);
    P->fs_tu = dvec(P->nu);
  }
  if (res) {
    for (l = 0; l < P->nlev; ++l) {
      res->level_rows[l] = P->lev[l].A.n; res->level_nnz[l] = P->lev[l].A.nnz;
      res->cheb_emin[l] = P->lev[l].emin; res->cheb_emax[l] = P->lev[l].emax;
      res->cheb_emin_est[l] = P->lev[l].emin_est; res->cheb_emax_est[l] = P->lev[l].emax_est;
    }
  }
  P->pc_ready = 1;
  return 0;
}

int xo_solve(xo_problem *P, const xo_solver *s, const double *b, double *x, xo_result *res)
{
  gmres_t g; outer_ctx ctx; double t0;
  memset(res, 0, sizeof(*res));
  t0 = xo_wtime();
  { xo_solver key = *s; key.max_outer_sample = 0;
    if (!P->pc_ready || memcmp(&P->sopt, &key, sizeof(key))) { if (xo_pc_setup(P, s, res)) return 1; } }
  res->setup_seconds = xo_wtime() - t0;
  { int l; for (l = 0; l < P->nlev; ++l) {
      res->level_rows[l] = P->lev[l].A.n; res->level_nnz[l] = P->lev[l].A.nnz;
      res->cheb_emin[l] = P->lev[l].emin; res->cheb_emax[l] = P->lev[l].emax;
      res->cheb_emin_est[l] = P->lev[l].emin_est; res->cheb_emax_est[l] = P->lev[l].emax_est; } }
  ctx.P = P; ctx.res = res;
  memset(&g, 0, sizeof(g));
  g.n = P->n; g.restart = s->restart; g.max_it = s->max_it; g.rtol = s->rtol; g.atol = s->atol; g.dtol = s->dtol;
  g.flexible = (s->ksp_type == XO_KSP_FGMRES);
  g.right = g.flexible ? 1 : (s->pc_side == XO_SIDE_RIGHT);
  g.amult = outer_amult; g.actx = &ctx;
  g.pc = s->pc_type == XO_PC_ABF ? outer_pc_abf : s->pc_type == XO_PC_JACOBI ? outer_pc_jacobi : NULL; g.pctx = &ctx;
  if (g.flexible && !g.pc) return xo_fail(P, "fgmres needs a preconditioner in this oracle");
  g.hist = res->hist; g.hist_cap = 2048; g.sample_stop = s->max_outer_sample;
  P->n_a00_mult = P->n_a_mult = 0;
  memset(x, 0, sizeof(double) * P->n);   /* initial guess is zero */
  t0 = xo_wtime();
  gmres_solve(&g, b ? b : P->F, x);
  res->solve_seconds = xo_wtime() - t0;
  res->its = g.its; res->reason = g.reason; res->nhist = g.nhist;
  res->n_a00_mult = P->n_a00_mult; res->n_a_mult = P->n_a_mult;
  return 0;
}
