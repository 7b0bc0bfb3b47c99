// xsb_io.cu -- PETSc binary writers: what the reference's -dump_operator / -dump_solution /
// -dump_scaled_mass_matrix produce through PetscViewerBinaryOpen + MatView / VecView (exSaddle_io.c:61-88,
// exSaddle.c:488-501, 535-537), so a PETSc / Octave user can load this library's operator and solution
// (octave_demo.m:10-12, PetscBinaryRead).  Format (PETSc's, all big-endian):
//   Mat: int32 1211216 (MAT_FILE_CLASSID), rows, cols, nnz; int32 row lengths[rows]; int32 columns[nnz]; float64 values[nnz]
//   Vec: int32 1211214 (VEC_FILE_CLASSID), n; float64 values[n]
// Pure host code on caller-supplied or device-fetched arrays; no PETSc needed.
#include "xsb.h"

static inline uint32_t be32(uint32_t v) { return ((v & 0xffu) << 24) | ((v & 0xff00u) << 8) | ((v >> 8) & 0xff00u) | (v >> 24); }
static inline uint64_t be64(uint64_t v) { return ((uint64_t)be32((uint32_t)v) << 32) | be32((uint32_t)(v >> 32)); }

static int put_i32(FILE *f, const int32_t *v, size_t n)
{
  std::vector<uint32_t> buf(1 << 16);
  for (size_t i = 0; i < n;) { size_t m = n - i < buf.size() ? n - i : buf.size(); for (size_t k = 0; k < m; ++k) buf[k] = be32((uint32_t)v[i + k]); if (fwrite(buf.data(), 4, m, f) != m) return 1; i += m; }
  return 0;
}
static int put_f64(FILE *f, const double *v, size_t n)
{
  std::vector<uint64_t> buf(1 << 15);
  for (size_t i = 0; i < n;) { size_t m = n - i < buf.size() ? n - i : buf.size(); for (size_t k = 0; k < m; ++k) { uint64_t u; memcpy(&u, &v[i + k], 8); buf[k] = be64(u); } if (fwrite(buf.data(), 8, m, f) != m) return 1; i += m; }
  return 0;
}

extern "C" {

int xsb_write_petsc_mat(const char *path, int64_t rows, int64_t cols, const int32_t *ia, const int32_t *ja, const double *a)
{
  if (!path || !ia || !ja || !a || rows < 0 || cols < 0) return XSB_ERR_ARG;
  const int64_t nnz = ia[rows];
  if (rows > INT32_MAX || cols > INT32_MAX || nnz > INT32_MAX) return XSB_ERR_SUP;   // 32-bit PetscInt header
  FILE *f = fopen(path, "wb"); if (!f) return XSB_ERR_ARG;
  const int32_t hdr[4] = {1211216, (int32_t)rows, (int32_t)cols, (int32_t)nnz};
  std::vector<int32_t> len((size_t)rows);
  for (int64_t i = 0; i < rows; ++i) len[i] = ia[i + 1] - ia[i];
  int rc = put_i32(f, hdr, 4) || put_i32(f, len.data(), (size_t)rows) || put_i32(f, ja, (size_t)nnz) || put_f64(f, a, (size_t)nnz);
  rc = fclose(f) || rc;
  return rc ? XSB_ERR_ARG : XSB_OK;
}

int xsb_write_petsc_vec(const char *path, int64_t n, const double *x)
{
  if (!path || !x || n < 0 || n > INT32_MAX) return XSB_ERR_ARG;
  FILE *f = fopen(path, "wb"); if (!f) return XSB_ERR_ARG;
  const int32_t hdr[2] = {1211214, (int32_t)n};
  int rc = put_i32(f, hdr, 2) || put_f64(f, x, (size_t)n);
  rc = fclose(f) || rc;
  return rc ? XSB_ERR_ARG : XSB_OK;
}

int xsb_dump_operator(xsb_ctx c, int which, const char *path)
{
  if (!c || !path) return XSB_ERR_ARG;
  int64_t rows, cols, nnz; int rc = xsb_mat_get_info(c, which, &rows, &cols, &nnz, nullptr); if (rc) return rc;
  if (nnz >= INT32_MAX) return xsb_fail(c, XSB_ERR_SUP, "operator has %lld nonzeros: the PETSc binary header holds a 32-bit count", (long long)nnz);
  std::vector<int32_t> ia((size_t)rows + 1), ja((size_t)nnz); std::vector<double> a((size_t)nnz);
  rc = xsb_mat_get_csr(c, which, ia.data(), ja.data(), a.data()); if (rc) return rc;
  if (xsb_write_petsc_mat(path, rows, cols, ia.data(), ja.data(), a.data())) return xsb_fail(c, XSB_ERR_ARG, "cannot write %s", path);
  return XSB_OK;
}

int xsb_dump_vector(xsb_ctx c, const double *x, int64_t n, const char *path)
{
  if (!c || !path || !x) return XSB_ERR_ARG;
  if (xsb_write_petsc_vec(path, n, x)) return xsb_fail(c, XSB_ERR_ARG, "cannot write %s", path);
  return XSB_OK;
}

}   // extern "C"
