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

// VTK XML StructuredGrid (.vts) with raw appended data, the container PetscViewerVTKOpen + VecView(DMDA vector) produce for
// ViewFields (exSaddle_io.c:128-177: <tag>uv[w].vts with the velocity components as scalar point fields, <tag>p.vts with p).
// nx, ny, nz: nodes of the lattice; h: node spacing from the origin (DMDASetUniformCoordinates_Saddle); data: nfields
// arrays, field f = data[f * stride0 + node * stride1] (strided so the interleaved velocity vector needs no copy).
int xsb_write_vts(const char *path, int nx, int ny, int nz, const double h[3], int nfields, const char *const *names,
                  const double *data, int64_t stride0, int64_t stride1)
{
  if (!path || !h || !names || !data || nx < 1 || ny < 1 || nz < 1 || nfields < 1) return XSB_ERR_ARG;
  FILE *f = fopen(path, "wb"); if (!f) return XSB_ERR_ARG;
  const int64_t nn = (int64_t)nx * ny * nz;
  fprintf(f, "<?xml version=\"1.0\"?>\n<VTKFile type=\"StructuredGrid\" version=\"1.0\" byte_order=\"LittleEndian\" header_type=\"UInt64\">\n");
  fprintf(f, "  <StructuredGrid WholeExtent=\"0 %d 0 %d 0 %d\">\n    <Piece Extent=\"0 %d 0 %d 0 %d\">\n", nx - 1, ny - 1, nz - 1, nx - 1, ny - 1, nz - 1);
  uint64_t offset = 0;
  fprintf(f, "      <Points>\n        <DataArray type=\"Float64\" Name=\"Position\" NumberOfComponents=\"3\" format=\"appended\" offset=\"%llu\" />\n      </Points>\n", (unsigned long long)offset);
  offset += 8 + 24 * (uint64_t)nn;
  fprintf(f, "      <PointData Scalars=\"ScalarPointData\">\n");
  for (int k = 0; k < nfields; ++k) {
    fprintf(f, "        <DataArray type=\"Float64\" Name=\"%s\" NumberOfComponents=\"1\" format=\"appended\" offset=\"%llu\" />\n", names[k], (unsigned long long)offset);
    offset += 8 + 8 * (uint64_t)nn;
  }
  fprintf(f, "      </PointData>\n    </Piece>\n  </StructuredGrid>\n  <AppendedData encoding=\"raw\">\n_");
  int rc = 0;
  { const uint64_t bytes = 24 * (uint64_t)nn; rc |= fwrite(&bytes, 8, 1, f) != 1;
    std::vector<double> row((size_t)3 * nx);
    for (int k = 0; k < nz && !rc; ++k) for (int j = 0; j < ny && !rc; ++j) {
      for (int i = 0; i < nx; ++i) { row[3 * i] = h[0] * i; row[3 * i + 1] = h[1] * j; row[3 * i + 2] = h[2] * k; }
      rc |= fwrite(row.data(), 8, (size_t)3 * nx, f) != (size_t)3 * nx;
    } }
  std::vector<double> buf((size_t)(nn < 65536 ? nn : 65536));
  for (int k = 0; k < nfields && !rc; ++k) {
    const uint64_t bytes = 8 * (uint64_t)nn; rc |= fwrite(&bytes, 8, 1, f) != 1;
    for (int64_t i0 = 0; i0 < nn && !rc; i0 += (int64_t)buf.size()) {
      const int64_t m = nn - i0 < (int64_t)buf.size() ? nn - i0 : (int64_t)buf.size();
      for (int64_t i = 0; i < m; ++i) buf[i] = data[k * stride0 + (i0 + i) * stride1];
      rc |= fwrite(buf.data(), 8, (size_t)m, f) != (size_t)m;
    }
  }
  fprintf(f, "\n  </AppendedData>\n</VTKFile>\n");
  rc = fclose(f) || rc;
  return rc ? XSB_ERR_ARG : XSB_OK;
}

// ViewFields(dm_saddle, X, tag) (exSaddle_io.c:128-177): <dir>/<tag>uv[w].vts and <dir>/<tag>p.vts from a host solution vector
int xsb_view_fields(xsb_ctx c, const double *x, const char *dir, const char *tag)
{
  if (!c || !x || !dir || !tag) return XSB_ERR_ARG;
  if (!c->assembled) return xsb_fail(c, XSB_ERR_ORDER, "xsb_view_fields before xsb_assemble");
  if (c->slab.nranks > 1) return xsb_fail(c, XSB_ERR_SUP, "xsb_view_fields writes the one-rank lattice; gather the owned ranges first");
  const Lattice &L = c->lat; const int nsd = c->nsd;
  const char *un[3] = {"u", "v", "w"}, *pn[1] = {"p"};
  const double hp[3] = {2.0 * L.hu[0], 2.0 * L.hu[1], nsd == 3 ? 2.0 * L.hu[2] : 1.0};
  const double hu[3] = {L.hu[0], L.hu[1], nsd == 3 ? L.hu[2] : 1.0};
  const std::string pu = std::string(dir) + "/" + tag + (nsd == 3 ? "uvw.vts" : "uv.vts"), pp = std::string(dir) + "/" + tag + "p.vts";
  if (xsb_write_vts(pu.c_str(), L.NX, L.NY, L.NZ, hu, nsd, un, x, 1, nsd)) return xsb_fail(c, XSB_ERR_ARG, "cannot write %s", pu.c_str());
  if (xsb_write_vts(pp.c_str(), L.PX, L.PY, L.PZ, hp, 1, pn, x + L.nu, 0, 1)) return xsb_fail(c, XSB_ERR_ARG, "cannot write %s", pp.c_str());
  return XSB_OK;
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
