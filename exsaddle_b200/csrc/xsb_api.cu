// xsb_api.cu -- the C ABI (include/exsaddle_b200.h): handle lifecycle, options database, host<->device
// staging for the host-pointer entry points, and the host-side integer index maps.
#include "xsb.h"
#include <cstdarg>
#include <fstream>
#include <sstream>

int xsb_fail(xsb_ctx c, int code, const char *fmt, ...)
{
  char buf[1024]; va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
  if (c) c->err = buf;
  return code;
}

template <class T> int dev_alloc(xsb_ctx c, T **p, size_t n)
{
  void *q = nullptr; if (n == 0) n = 1;
  cudaError_t e = cudaMalloc(&q, n * sizeof(T) + 32);   // + slack: 16-byte granular bulk copies may touch the granule holding the last value
  if (e != cudaSuccess) return xsb_fail(c, XSB_ERR_MEM, "cudaMalloc of %zu bytes failed: %s", n * sizeof(T), cudaGetErrorString(e));
  cudaMemsetAsync(q, 0, n * sizeof(T), c->stream);   // ghost entries of slab vectors must be finite
  c->allocs.push_back(q); c->alloc_phase.push_back((char)c->phase); *p = (T *)q;
  return 0;
}
template int dev_alloc<double>(xsb_ctx, double **, size_t);
template int dev_alloc<int>(xsb_ctx, int **, size_t);
template int dev_alloc<char>(xsb_ctx, char **, size_t);
template int dev_alloc<unsigned char>(xsb_ctx, unsigned char **, size_t);
template int dev_alloc<unsigned short>(xsb_ctx, unsigned short **, size_t);
template int dev_alloc<unsigned>(xsb_ctx, unsigned **, size_t);

int dev_free_all(xsb_ctx c)
{
  for (void *p : c->allocs) cudaFree(p);
  c->allocs.clear(); c->alloc_phase.clear();
  return 0;
}
int dev_free(xsb_ctx c, void *p)
{
  if (!p) return 0;
  for (size_t i = 0; i < c->allocs.size(); ++i) if (c->allocs[i] == p) { c->allocs.erase(c->allocs.begin() + i); c->alloc_phase.erase(c->alloc_phase.begin() + i); break; }
  cudaFree(p);
  return 0;
}
int dev_free_phase(xsb_ctx c, int phase)
{
  size_t k = 0;
  for (size_t i = 0; i < c->allocs.size(); ++i) {
    if (c->alloc_phase[i] == phase) cudaFree(c->allocs[i]);
    else { c->allocs[k] = c->allocs[i]; c->alloc_phase[k] = c->alloc_phase[i]; ++k; }
  }
  c->allocs.resize(k); c->alloc_phase.resize(k);
  return 0;
}

extern "C" {

int xsb_device_available(void)
{
  int n = 0; cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) { cudaGetLastError(); return 0; }
  return n > 0;
}

int xsb_create(xsb_ctx *out, int nsd, int lame, int device)
{
  if (!out) return XSB_ERR_ARG;
  *out = nullptr;
  if (nsd != 2 && nsd != 3) return XSB_ERR_ARG;   // exSaddle.h:7-9
  xsb_ctx c = new xsb_ctx_s();
  c->nsd = nsd; c->lame = lame ? 1 : 0;
  *out = c;
  if (!xsb_device_available()) { c->have_device = false; c->err = "no CUDA device: exsaddle_b200 has no CPU path"; return XSB_OK; }
  if (device >= 0) { CUDA_OK(cudaSetDevice(device)); c->device = device; } else CUDA_OK(cudaGetDevice(&c->device));
  CUDA_OK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  CUDA_OK(cudaEventCreate(&c->ev0)); CUDA_OK(cudaEventCreate(&c->ev1)); CUDA_OK(cudaEventCreate(&c->evk0)); CUDA_OK(cudaEventCreate(&c->evk1));
  c->have_device = true;
  return XSB_OK;
}

#define NEED_DEVICE(c) do { if (!(c)) return XSB_ERR_ARG; if (!(c)->have_device) return xsb_fail(c, XSB_ERR_NO_DEVICE, "no CUDA device: exsaddle_b200 has no CPU path"); cudaSetDevice((c)->device); } while (0)

int xsb_reset(xsb_ctx c)
{
  if (!c) return XSB_ERR_ARG;
  if (c->have_device) { cudaSetDevice(c->device); cudaStreamSynchronize(c->stream); mg_graphs_release(c); mmg_free(c); fsd_free(c); grad_free(c); if (c->asmpc) { asm_free(c->asmpc); c->asmpc = nullptr; } dev_free_all(c); if (c->red_h) cudaFreeHost(c->red_h); for (cudaEvent_t e : c->evpool) cudaEventDestroy(e); }
  Options opt = c->opt; int nsd = c->nsd, lame = c->lame, device = c->device; bool hd = c->have_device;
  const int rank = c->slab.rank, nranks = c->slab.nranks; void *nccl = c->nccl, *p2p = c->p2p; cudaStream_t side = c->side; cudaEvent_t evf = c->ev_fork, evj = c->ev_join;
  cudaStream_t st = c->stream; cudaEvent_t e0 = c->ev0, e1 = c->ev1, k0 = c->evk0, k1 = c->evk1;
  *c = xsb_ctx_s();
  c->opt = opt; c->opt.used.clear(); c->nsd = nsd; c->lame = lame; c->device = device; c->have_device = hd;
  c->stream = st; c->ev0 = e0; c->ev1 = e1; c->evk0 = k0; c->evk1 = k1;
  c->slab.rank = rank; c->slab.nranks = nranks; c->nccl = nccl; c->p2p = p2p; c->side = side; c->ev_fork = evf; c->ev_join = evj;
  return XSB_OK;
}

int xsb_destroy(xsb_ctx *pc)
{
  if (!pc || !*pc) return XSB_OK;
  xsb_ctx c = *pc;
  xsb_reset(c);
  comm_destroy(c);
  if (c->have_device && c->side) { cudaStreamDestroy(c->side); cudaEventDestroy(c->ev_fork); cudaEventDestroy(c->ev_join); }
  if (c->have_device) { cudaEventDestroy(c->ev0); cudaEventDestroy(c->ev1); cudaEventDestroy(c->evk0); cudaEventDestroy(c->evk1); cudaStreamDestroy(c->stream); }
  delete c; *pc = nullptr;
  return XSB_OK;
}

const char *xsb_last_error(xsb_ctx c) { return c ? c->err.c_str() : "null handle"; }

// ---------------------------------------------------------------- options
static bool looks_like_value(const std::string &t)
{
  if (t.empty() || t[0] != '-') return true;
  if (t.size() > 1 && (isdigit((unsigned char)t[1]) || t[1] == '.')) return true;   // negative number
  return false;
}

int xsb_set_option(xsb_ctx c, const char *key, const char *value)
{
  if (!c || !key) return XSB_ERR_ARG;
  std::string k = key; if (!k.empty() && k[0] == '-') k = k.substr(1);
  if (k.empty()) return xsb_fail(c, XSB_ERR_ARG, "empty option name");
  if (k == "options_file") return xsb_set_options_file(c, value ? value : "");
  c->opt.kv[k] = value ? value : "";
  c->ksp_ready = false; c->mf_opts_read = false;
  return XSB_OK;
}

int xsb_set_options(xsb_ctx c, const char *cmdline)
{
  if (!c || !cmdline) return XSB_ERR_ARG;
  std::istringstream in(cmdline); std::string line;
  while (std::getline(in, line)) {
    size_t h = line.find('#'); if (h != std::string::npos) line = line.substr(0, h);
    std::istringstream ls(line); std::vector<std::string> tok; std::string t;
    while (ls >> t) tok.push_back(t);
    for (size_t i = 0; i < tok.size();) {
      if (tok[i][0] != '-' || looks_like_value(tok[i])) { ++i; continue; }
      if (i + 1 < tok.size() && looks_like_value(tok[i + 1])) { XSB_CHK(xsb_set_option(c, tok[i].c_str(), tok[i + 1].c_str())); i += 2; }
      else { XSB_CHK(xsb_set_option(c, tok[i].c_str(), nullptr)); i += 1; }
    }
  }
  return XSB_OK;
}

int xsb_set_options_file(xsb_ctx c, const char *path)
{
  if (!c || !path) return XSB_ERR_ARG;
  std::ifstream f(path);
  if (!f) return xsb_fail(c, XSB_ERR_ARG, "cannot open options file %s", path);
  std::stringstream ss; ss << f.rdbuf();
  // command-line options win over file options (PETSc inserts the file first)
  std::map<std::string, std::string> keep = c->opt.kv;
  XSB_CHK(xsb_set_options(c, ss.str().c_str()));
  for (auto &kv : keep) c->opt.kv[kv.first] = kv.second;
  return XSB_OK;
}

int xsb_options_left(xsb_ctx c, char *buf, int buflen)
{
  if (!c || !buf || buflen < 1) return XSB_ERR_ARG;
  std::string s;
  for (auto &kv : c->opt.kv) if (!c->opt.used.count(kv.first)) { s += "-" + kv.first; if (!kv.second.empty()) s += " " + kv.second; s += "\n"; }
  snprintf(buf, buflen, "%s", s.c_str());
  return XSB_OK;
}

// ---------------------------------------------------------------- set-up
int xsb_assemble(xsb_ctx c)
{
  NEED_DEVICE(c);
  if (c->assembled) { Options keep = c->opt; XSB_CHK(xsb_reset(c)); c->opt = keep; }
  CUDA_OK(cudaEventRecord(c->ev0, c->stream));
  int rc = fe_assemble(c);
  if (rc) return rc;
  XSB_CHK(comm_p2p_setup(c));   // slabs: peer-memory halo windows sized for this lattice (collective)
  CUDA_OK(cudaEventRecord(c->ev1, c->stream)); CUDA_OK(cudaEventSynchronize(c->ev1));
  return XSB_OK;
}

int xsb_banner(xsb_ctx c, char *buf, int buflen)
{
  if (!c || !buf || buflen < 1) return XSB_ERR_ARG;
  if (c->banner.empty()) { int rc = fe_resolve_model(c); if (rc) return rc; }
  snprintf(buf, buflen, "%s", c->banner.c_str());
  return XSB_OK;
}

int xsb_get_sizes(xsb_ctx c, int64_t out[8])
{
  if (!c || !out) return XSB_ERR_ARG;
  if (!c->assembled) return xsb_fail(c, XSB_ERR_ORDER, "xsb_get_sizes before xsb_assemble");
  const Lattice &L = c->lat;
  out[0] = L.n; out[1] = L.nu; out[2] = L.np; out[3] = c->A.nnz; out[4] = xsb_prealloc_total(c->nsd, L.mx, L.my, L.mz);
  out[5] = L.nel; out[6] = c->nbc; out[7] = c->Mp.nnz;
  return XSB_OK;
}

static int pick_csr(xsb_ctx c, int which, const Csr **S, const Baij **B)
{
  *S = nullptr; *B = nullptr;
  if (!c->assembled) return xsb_fail(c, XSB_ERR_ORDER, "matrix requested before xsb_assemble");
  if (c->no_A && (which == XSB_MAT_A || which == XSB_MAT_A00)) return xsb_fail(c, XSB_ERR_SUP, "-xsb_matrix_free full: A and A00 are not stored (use xsb_mat_mult, or XSB_MAT_A00_MF)");
  switch (which) {
  case XSB_MAT_A: *S = &c->A; return 0;
  case XSB_MAT_A00: case XSB_MAT_A00_MF: *B = &c->A00; return 0;
  case XSB_MAT_A01: case XSB_MAT_A01_MF: *S = &c->A01; return 0;
  case XSB_MAT_A10: case XSB_MAT_A10_MF: *S = &c->A10; return 0;
  case XSB_MAT_A11: *S = &c->A11; return 0;
  case XSB_MAT_MP: *S = &c->Mp; return 0;
  default:
    if (which >= XSB_MAT_MG_LEVEL0 && which < XSB_MAT_MG_LEVEL0 + c->nlev) { *B = &c->lev[which - XSB_MAT_MG_LEVEL0].A; return 0; }
    return xsb_fail(c, XSB_ERR_ARG, "unknown matrix id %d", which);
  }
}

int xsb_mat_get_info(xsb_ctx c, int which, int64_t *rows, int64_t *cols, int64_t *nnz, int *bs)
{
  NEED_DEVICE(c);
  if (c->no_A && c->assembled && (which == XSB_MAT_A || which == XSB_MAT_A00)) {   // sizes of the operators that are applied but not stored
    const int64_t n = which == XSB_MAT_A ? c->lat.n : c->lat.nu;
    if (rows) *rows = n; if (cols) *cols = n; if (nnz) *nnz = which == XSB_MAT_A ? c->A.nnz : 0; if (bs) *bs = which == XSB_MAT_A ? 1 : c->nsd;
    return XSB_OK;
  }
  const Csr *S; const Baij *B; XSB_CHK(pick_csr(c, which, &S, &B));
  if (S) { if (rows) *rows = S->n; if (cols) *cols = S->m; if (nnz) *nnz = S->nnz; if (bs) *bs = 1; }
  else { if (rows) *rows = (int64_t)B->nb * B->bs; if (cols) *cols = (int64_t)B->nb * B->bs; if (nnz) *nnz = B->nblk * B->bs * B->bs; if (bs) *bs = B->bs; }
  return XSB_OK;
}

int xsb_mat_get_csr(xsb_ctx c, int which, int32_t *ia, int32_t *ja, double *a)
{
  NEED_DEVICE(c);
  const Csr *S; const Baij *B; XSB_CHK(pick_csr(c, which, &S, &B));
  CUDA_OK(cudaStreamSynchronize(c->stream));
  if (S) {
    if (ia) CUDA_OK(cudaMemcpy(ia, S->ia, sizeof(int) * ((size_t)S->n + 1), cudaMemcpyDeviceToHost));
    if (ja) CUDA_OK(cudaMemcpy(ja, S->ja, sizeof(int) * (size_t)S->nnz, cudaMemcpyDeviceToHost));
    if (a) CUDA_OK(cudaMemcpy(a, S->a, sizeof(double) * (size_t)S->nnz, cudaMemcpyDeviceToHost));
    return XSB_OK;
  }
  return baij_to_csr_host(c, *B, ia, ja, a);
}

int xsb_mat_mult_dev(xsb_ctx c, int which, const double *x, double *y)
{
  NEED_DEVICE(c);
  if (c->no_A && c->assembled) {
    if (which == XSB_MAT_A) return op_full_mult(c, x, y);
    if (which == XSB_MAT_A00) which = XSB_MAT_A00_MF;
  }
  const Csr *S; const Baij *B; XSB_CHK(pick_csr(c, which, &S, &B));
  if (which == XSB_MAT_A01_MF) return grad_apply(c, x, y, 0, c->lat.nu);
  if (which == XSB_MAT_A10_MF) return div_apply(c, x, y, 0, c->lat.np);
  if (S) return which == XSB_MAT_A ? op_full_mult(c, x, y) : spmv_csr(c, *S, x, y);
  Epilogue ep;
  if (which == XSB_MAT_A00_MF) { XSB_CHK(mf_setup(c)); return mf_a00_apply(c, x, y, ep); }
  return spmv_baij(c, *B, x, y, ep);
}

int xsb_mat_mult(xsb_ctx c, int which, const double *x, double *y)
{
  NEED_DEVICE(c);
  int64_t rows, cols; XSB_CHK(xsb_mat_get_info(c, which, &rows, &cols, nullptr, nullptr));
  double *dx = nullptr, *dy = nullptr;
  CUDA_OK(cudaMalloc(&dx, sizeof(double) * cols)); CUDA_OK(cudaMalloc(&dy, sizeof(double) * rows));
  CUDA_OK(cudaMemcpyAsync(dx, x, sizeof(double) * cols, cudaMemcpyHostToDevice, c->stream));
  int rc = xsb_mat_mult_dev(c, which, dx, dy);
  if (!rc) { cudaError_t e = cudaMemcpyAsync(y, dy, sizeof(double) * rows, cudaMemcpyDeviceToHost, c->stream); if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream); if (e != cudaSuccess) rc = xsb_fail(c, XSB_ERR_CUDA, "%s", cudaGetErrorString(e)); }
  cudaFree(dx); cudaFree(dy);
  return rc;
}

int xsb_mat_mult_transpose(xsb_ctx c, int which, const double *x, double *y)
{
  NEED_DEVICE(c);
  // the saddle operator and its diagonal blocks are symmetric (MatZeroRowsColumns keeps the symmetry); the gradient and
  // divergence blocks are each other's transposes by construction (the same g[d] goes to both, femixedspace.c:2584-2590)
  if (which == XSB_MAT_A01) which = XSB_MAT_A10; else if (which == XSB_MAT_A10) which = XSB_MAT_A01;
  return xsb_mat_mult(c, which, x, y);
}

int xsb_mat_get_diagonal(xsb_ctx c, int which, double *d)
{
  NEED_DEVICE(c);
  const Csr *S; const Baij *B; XSB_CHK(pick_csr(c, which, &S, &B));
  const int64_t rows = S ? S->n : (int64_t)B->nb * B->bs;
  double *dd = nullptr; CUDA_OK(cudaMalloc(&dd, sizeof(double) * rows));
  int rc = S ? csr_diag(c, *S, dd) : baij_diag(c, *B, dd);   // read on the device: one pass over the matrix, only the diagonal crosses PCIe
  if (!rc) { cudaError_t e = cudaMemcpyAsync(d, dd, sizeof(double) * rows, cudaMemcpyDeviceToHost, c->stream); if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream); if (e != cudaSuccess) rc = xsb_fail(c, XSB_ERR_CUDA, "%s", cudaGetErrorString(e)); }
  cudaFree(dd);
  return rc;
}

int xsb_vec_get_rhs(xsb_ctx c, double *F)
{
  NEED_DEVICE(c);
  if (!c->assembled) return xsb_fail(c, XSB_ERR_ORDER, "xsb_vec_get_rhs before xsb_assemble");
  CUDA_OK(cudaStreamSynchronize(c->stream));
  CUDA_OK(cudaMemcpy(F, c->F, sizeof(double) * c->lat.n, cudaMemcpyDeviceToHost));
  return XSB_OK;
}

int xsb_get_bc(xsb_ctx c, int32_t *idx, double *val)
{
  NEED_DEVICE(c);
  if (!c->assembled) return xsb_fail(c, XSB_ERR_ORDER, "xsb_get_bc before xsb_assemble");
  if (c->nbc == 0) return XSB_OK;
  if (idx) CUDA_OK(cudaMemcpy(idx, c->bc_idx, sizeof(int) * c->nbc, cudaMemcpyDeviceToHost));
  if (val) CUDA_OK(cudaMemcpy(val, c->bc_val, sizeof(double) * c->nbc, cudaMemcpyDeviceToHost));
  return XSB_OK;
}

int xsb_get_coeff_qp(xsb_ctx c, int slot, double *out)
{
  NEED_DEVICE(c);
  if (!c->assembled) return xsb_fail(c, XSB_ERR_ORDER, "xsb_get_coeff_qp before xsb_assemble");
  if (slot < 0 || slot >= XSB_NSLOT) return xsb_fail(c, XSB_ERR_ARG, "coefficient slot %d", slot);
  const int64_t nq = c->lat.nel * (c->nsd == 3 ? 27 : 9);
  CUDA_OK(cudaMemcpy(out, c->coeff + (int64_t)slot * nq, sizeof(double) * nq, cudaMemcpyDeviceToHost));
  return XSB_OK;
}

// ---------------------------------------------------------------- solver
int xsb_ksp_setup(xsb_ctx c) { NEED_DEVICE(c); return ksp_setup(c); }
int xsb_ksp_reset(xsb_ctx c)
{
  if (!c) return XSB_ERR_ARG;
  c->ksp_ready = false;
  if (!c->have_device) return XSB_OK;
  cudaSetDevice(c->device);
  return ksp_release(c);
}
int xsb_get_state(xsb_ctx c, int *assembled, int *ksp_ready)
{ if (!c) return XSB_ERR_ARG; if (assembled) *assembled = c->assembled ? 1 : 0; if (ksp_ready) *ksp_ready = c->ksp_ready ? 1 : 0; return XSB_OK; }

int xsb_ksp_solve_dev(xsb_ctx c, const double *b, double *x) { NEED_DEVICE(c); return ksp_solve(c, b, x); }

int xsb_ksp_solve(xsb_ctx c, const double *b, double *x)
{
  NEED_DEVICE(c);
  if (!c->ksp_ready) return xsb_fail(c, XSB_ERR_ORDER, "xsb_ksp_solve called before xsb_ksp_setup");
  const int64_t n = c->lat.n;
  if (b) CUDA_OK(cudaMemcpyAsync(c->bdev, b, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
  XSB_CHK(ksp_solve(c, b ? c->bdev : nullptr, c->xdev));
  CUDA_OK(cudaMemcpyAsync(x, c->xdev, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
  CUDA_OK(cudaStreamSynchronize(c->stream));
  return XSB_OK;
}

int xsb_pc_apply_dev(xsb_ctx c, const double *r, double *z) { NEED_DEVICE(c); if (!c->ksp_ready) return xsb_fail(c, XSB_ERR_ORDER, "PC not set up"); return pc_apply(c, r, z, nullptr); }

static int staged(xsb_ctx c, int64_t nin, int64_t nout, const double *hin, double *hout, int (*fn)(xsb_ctx, const double *, double *))
{
  double *di = nullptr, *dout = nullptr;
  CUDA_OK(cudaMalloc(&di, sizeof(double) * nin)); CUDA_OK(cudaMalloc(&dout, sizeof(double) * nout));
  CUDA_OK(cudaMemcpyAsync(di, hin, sizeof(double) * nin, cudaMemcpyHostToDevice, c->stream));
  CUDA_OK(cudaMemsetAsync(dout, 0, sizeof(double) * nout, c->stream));
  int rc = fn(c, di, dout);
  if (!rc) { cudaError_t e = cudaMemcpyAsync(hout, dout, sizeof(double) * nout, cudaMemcpyDeviceToHost, c->stream); if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream); if (e != cudaSuccess) rc = xsb_fail(c, XSB_ERR_CUDA, "%s", cudaGetErrorString(e)); }
  cudaFree(di); cudaFree(dout);
  return rc;
}

int xsb_pc_apply(xsb_ctx c, const double *r, double *z)
{
  NEED_DEVICE(c);
  if (!c->ksp_ready) return xsb_fail(c, XSB_ERR_ORDER, "PC not set up");
  return staged(c, c->lat.n, c->lat.n, r, z, [](xsb_ctx cc, const double *a, double *b) { return pc_apply(cc, a, b, nullptr); });
}
int xsb_pc_mg_apply(xsb_ctx c, const double *b, double *x)
{
  NEED_DEVICE(c);
  if (!c->ksp_ready || c->so.pc_type != 2) return xsb_fail(c, XSB_ERR_ORDER, "PCMG not set up");
  return staged(c, c->lat.nu, c->lat.nu, b, x, [](xsb_ctx cc, const double *a, double *bb) { return mg_vcycle(cc, a, bb); });
}
int xsb_pc_schur_apply(xsb_ctx c, const double *b, double *x)
{
  NEED_DEVICE(c);
  if (!c->ksp_ready || c->so.pc_type != 2) return xsb_fail(c, XSB_ERR_ORDER, "fieldsplit PC not set up");
  return staged(c, c->lat.np, c->lat.np, b, x, [](xsb_ctx cc, const double *a, double *bb) {
    if (cc->so.p_pc == 0) return ilu_apply(cc, a + cc->own_p.off0, bb + cc->own_p.off0);   // bjacobi: this rank's block acts on its owned rows
    return vec_pmult(cc, cc->lat.np, cc->mp_idiag, a, bb); });
}

// device time (us) of `reps` back-to-back ghost exchanges of a velocity vector (collective), and of the fine-level A00 kernel
// without its exchange: out[0] = exchange, out[1] = exchange of a single node plane per side (a distributed coarse level),
// out[2] = fine-level product without exchange
int xsb_time_halo(xsb_ctx c, int reps, double out[3])
{
  NEED_DEVICE(c);
  if (!c->ksp_ready || reps < 1 || !out) return xsb_fail(c, XSB_ERR_ORDER, "solver not set up");
  const Lattice &L = c->lat; double *a = c->w_t1, *bb = c->w_t2; float ms = 0;
  XSB_CHK(vec_set(c, L.n, 1.0, a));
  for (int which = 0; which < 3; ++which) {
    auto one = [&]() -> int {
      if (which == 0) return comm_halo_u(c, a);
      if (which == 1) return comm_halo_planes(c, a, (int64_t)L.nsd * (L.mx + 1) * (L.my + 1), c->slab.ou0, c->slab.ou1, 1, 1);
      Epilogue ep; if (c->so.matrix_free) return mf_a00_apply(c, a, bb, ep);
      const int pn = L.NX * L.NY; return spmv_baij(c, c->A00, a, bb, ep, c->slab.ou0 * pn, (c->slab.ou1 - c->slab.ou0) * pn); };
    XSB_CHK(one()); XSB_CHK(one());
    CUDA_OK(cudaEventRecord(c->evk0, c->stream));
    for (int i = 0; i < reps; ++i) XSB_CHK(one());
    CUDA_OK(cudaEventRecord(c->evk1, c->stream)); CUDA_OK(cudaEventSynchronize(c->evk1));
    CUDA_OK(cudaEventElapsedTime(&ms, c->evk0, c->evk1));
    out[which] = 1e3 * (double)ms / reps;
  }
  return XSB_OK;
}

// device time of `reps` pressure-block solves (ILU(0) or Jacobi) on resident vectors, CUDA events on the handle's stream
int xsb_time_pc_schur(xsb_ctx c, int reps, double *ms_per_apply)
{
  NEED_DEVICE(c);
  if (!c->ksp_ready || c->so.pc_type != 2 || reps < 1 || !ms_per_apply) return xsb_fail(c, XSB_ERR_ORDER, "fieldsplit PC not set up");
  double *a = c->w_t1, *bb = c->w_t2;
  XSB_CHK(vec_set(c, c->lat.np, 1.0, a));
  auto one = [&]() { return c->so.p_pc == 0 ? ilu_apply(c, a + c->own_p.off0, bb + c->own_p.off0) : vec_pmult(c, c->lat.np, c->mp_idiag, a, bb); };
  XSB_CHK(one());
  CUDA_OK(cudaEventRecord(c->evk0, c->stream));
  for (int i = 0; i < reps; ++i) XSB_CHK(one());
  CUDA_OK(cudaEventRecord(c->evk1, c->stream)); CUDA_OK(cudaEventSynchronize(c->evk1));
  float ms = 0; CUDA_OK(cudaEventElapsedTime(&ms, c->evk0, c->evk1));
  *ms_per_apply = (double)ms / reps;
  return XSB_OK;
}

int xsb_mg_restrict(xsb_ctx c, int lc, const double *rf, double *bc)
{
  NEED_DEVICE(c);
  if (!c->ksp_ready || lc < 0 || lc + 1 >= c->nlev) return xsb_fail(c, XSB_ERR_ARG, "bad MG level %d", lc);
  const Level &F = c->lev[lc + 1], &C = c->lev[lc];
  const int64_t nf = (int64_t)F.A.nb * F.A.bs, nc = (int64_t)C.A.nb * C.A.bs;
  double *dc = nullptr, *df = nullptr;
  CUDA_OK(cudaMalloc(&dc, sizeof(double) * nc)); CUDA_OK(cudaMalloc(&df, sizeof(double) * nf));
  CUDA_OK(cudaMemcpyAsync(df, rf, sizeof(double) * nf, cudaMemcpyHostToDevice, c->stream));
  int rc = mg_restrict(c, F, C, df, dc);
  if (!rc) { cudaError_t e = cudaMemcpyAsync(bc, dc, sizeof(double) * nc, cudaMemcpyDeviceToHost, c->stream); if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream); if (e != cudaSuccess) rc = xsb_fail(c, XSB_ERR_CUDA, "%s", cudaGetErrorString(e)); }
  cudaFree(dc); cudaFree(df);
  return rc;
}
int xsb_mg_interpolate_add(xsb_ctx c, int lc, const double *xc, double *xf)
{
  NEED_DEVICE(c);
  if (!c->ksp_ready || lc < 0 || lc + 1 >= c->nlev) return xsb_fail(c, XSB_ERR_ARG, "bad MG level %d", lc);
  const Level &F = c->lev[lc + 1], &C = c->lev[lc];
  const int64_t nf = (int64_t)F.A.nb * F.A.bs, nc = (int64_t)C.A.nb * C.A.bs;
  double *dc = nullptr, *df = nullptr;
  CUDA_OK(cudaMalloc(&dc, sizeof(double) * nc)); CUDA_OK(cudaMalloc(&df, sizeof(double) * nf));
  CUDA_OK(cudaMemcpyAsync(dc, xc, sizeof(double) * nc, cudaMemcpyHostToDevice, c->stream));
  CUDA_OK(cudaMemcpyAsync(df, xf, sizeof(double) * nf, cudaMemcpyHostToDevice, c->stream));
  int rc = mg_prolong_add(c, F, C, dc, df);
  if (!rc) { cudaError_t e = cudaMemcpyAsync(xf, df, sizeof(double) * nf, cudaMemcpyDeviceToHost, c->stream); if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream); if (e != cudaSuccess) rc = xsb_fail(c, XSB_ERR_CUDA, "%s", cudaGetErrorString(e)); }
  cudaFree(dc); cudaFree(df);
  return rc;
}

int xsb_ksp_get_iterations(xsb_ctx c, int *its, int *reason) { if (!c) return XSB_ERR_ARG; if (its) *its = c->its; if (reason) *reason = c->reason; return XSB_OK; }
int xsb_ksp_get_history(xsb_ctx c, double *hist, int cap, int *n)
{
  if (!c) return XSB_ERR_ARG;
  int m = (int)c->hist.size(); if (n) *n = m;
  for (int i = 0; i < m && i < cap; ++i) hist[i] = c->hist[i];
  return XSB_OK;
}
int xsb_ksp_get_inner_iterations(xsb_ctx c, int *its, int cap, int *n)
{
  if (!c) return XSB_ERR_ARG;
  int m = (int)c->inner_its.size(); if (n) *n = m;
  for (int i = 0; i < m && i < cap; ++i) its[i] = c->inner_its[i];
  return XSB_OK;
}
int xsb_ksp_get_inner_reasons(xsb_ctx c, int *reasons, int cap, int *n)
{
  if (!c) return XSB_ERR_ARG;
  int m = (int)c->inner_reason.size(); if (n) *n = m;
  for (int i = 0; i < m && i < cap; ++i) reasons[i] = c->inner_reason[i];
  return XSB_OK;
}
int xsb_ksp_get_chebyshev(xsb_ctx c, int level, double *emin_est, double *emax_est, double *emin, double *emax)
{
  if (!c || level < 0 || level >= c->nlev) return XSB_ERR_ARG;
  const Level &L = c->lev[level];
  if (emin_est) *emin_est = L.emin_est; if (emax_est) *emax_est = L.emax_est; if (emin) *emin = L.emin; if (emax) *emax = L.emax;
  return XSB_OK;
}
int xsb_ksp_get_timing(xsb_ctx c, double *setup_ms, double *solve_ms) { if (!c) return XSB_ERR_ARG; if (setup_ms) *setup_ms = c->setup_ms; if (solve_ms) *solve_ms = c->solve_ms; return XSB_OK; }
int xsb_ksp_view(xsb_ctx c, char *buf, int buflen)
{
  if (!c || !buf || buflen < 1) return XSB_ERR_ARG;
  if (!c->ksp_ready) return xsb_fail(c, XSB_ERR_ORDER, "xsb_ksp_view before xsb_ksp_setup");
  const SolverOpts &s = c->so; std::string o; char t[512];
  auto add = [&](const char *fmt, auto... a) { snprintf(t, sizeof(t), fmt, a...); o += t; };
  add("KSP Object: (saddle_) %d GPU(s)\n  type: %s\n    restart=%d, using Classical (unmodified) Gram-Schmidt Orthogonalization with no iterative refinement\n",
      c->slab.nranks, s.ksp_type == 1 ? "fgmres" : "gmres", s.restart);
  add("  maximum iterations=%d, initial guess is zero\n  tolerances:  relative=%g, absolute=%g, divergence=%g.\n", s.max_it, s.rtol, s.atol, s.dtol);
  add("  %s preconditioning\n  using %s norm type for convergence test\n", s.right ? "right" : "left", s.right ? "UNPRECONDITIONED" : "PRECONDITIONED");
  add("PC Object: (saddle_)\n");
  if (s.pc_type == 0) add("  type: none\n");
  else if (s.pc_type == 1) add("  type: jacobi\n");
  else if (s.pc_type == 3) add("  type: mg\n    type is MULTIPLICATIVE, levels=%d cycles=v\n      Not using Galerkin computed coarse grid matrices\n    smoothers: gmres + jacobi, exactly %d iterations; coarse: dense LU (pivoted inverse on the device)\n",
                               c->opt.integer("nlevels", 1), c->opt.integer("saddle_mg_levels_ksp_max_it", 2));
  else {
    add("  type: fieldsplit\n    FieldSplit with Schur preconditioner, factorization UPPER\n    Preconditioner for the Schur complement formed from user provided matrix\n");
    add("    KSP solver for A00 block: (saddle_fieldsplit_u_) gcr, restart=%d, relative=%g; PC mg, MULTIPLICATIVE, levels=%d cycles=v, Galerkin coarse operators\n", s.u_restart, s.u_rtol, c->nlev);
    add("      fine-level products: %s\n", c->no_A ? "operator-free (sum-factorised Q2 element kernel; A and A00 not stored)" : s.matrix_free ? "matrix-free element kernel (A00 also assembled)" : "assembled BAIJ");
    for (int l = 0; l < c->nlev; ++l) {
      const Level &L = c->lev[l]; const long long rows = (long long)L.A.nb * L.A.bs, nz = (long long)L.A.nblk * L.A.bs * L.A.bs;
      if (l == 0 && c->nsub > 0) add("      level 0 (coarse): preonly + lu replaced by cg to 1e-13 preconditioned by an internal %d-level V-cycle (dense inverse at its bottom), rows=%lld, total: nonzeros=%lld, bs=%d; %d coarse solves, %d cg iterations so far\n", c->nsub, rows, nz, L.A.bs, c->coarse_solves, c->coarse_its);
      else if (l == 0) add("      level 0 (coarse): preonly + lu (dense inverse), rows=%lld, total: nonzeros=%lld, bs=%d\n", rows, nz, L.A.bs);
      else add("      level %d: chebyshev + jacobi, maximum iterations=%d, eigenvalue estimates used:  min = %g, max = %g%s, rows=%lld, total: nonzeros=%lld, bs=%d%s\n",
               l, s.cheb_its, L.emin, L.emax, s.n_cheb_fixed ? " (set explicitly)" : "", rows, nz, L.A.bs, L.dist ? ", z-slab distributed" : L.pdist ? ", distributed by node planes (operator replicated)" : "");
    }
    add("    KSP solver for S = A11 - A10 inv(A00) A01: (saddle_fieldsplit_p_) preonly; PC %s on Mpscaled (rows=%d, nonzeros=%lld)\n",
        s.p_pc == 0 ? "bjacobi, one block per GPU, ilu(0) in natural ordering" : "jacobi", c->Mp.n, (long long)c->Mp.nnz);
  }
  add("  linear system matrix: rows=%lld, cols=%lld, total: nonzeros=%lld%s\n", (long long)c->lat.n, (long long)c->lat.n, (long long)c->A.nnz, c->no_A ? " (not stored)" : ", type aij");
  snprintf(buf, buflen, "%s", o.c_str());
  return XSB_OK;
}

int xsb_ksp_get_counters(xsb_ctx c, int64_t out[8])
{
  if (!c || !out) return XSB_ERR_ARG;
  out[0] = c->n_a00; out[1] = c->n_a; out[2] = c->solve_launches; out[3] = c->a00_timed ? (int64_t)(c->a00_ns_sum / c->a00_timed) : 0;
  for (int i = 0; i < 4; ++i) out[4 + i] = c->a00_mode[i];
  return XSB_OK;
}
int xsb_ksp_get_profile(xsb_ctx c, double *ms, int64_t *count, int cap, int *ncat)
{
  if (!c || !ms || !count || !ncat) return XSB_ERR_ARG;
  *ncat = PROF_N;
  for (int i = 0; i < PROF_N && i < cap; ++i) { ms[i] = c->prof_ms[i]; count[i] = c->prof_cnt[i]; }
  return XSB_OK;
}
int xsb_comm_info(xsb_ctx c, int64_t out[4])
{
  if (!c || !out) return XSB_ERR_ARG;
  out[0] = c->slab.rank; out[1] = c->slab.nranks; out[2] = comm_p2p_active(c); out[3] = 0;
  for (int l = 0; l < c->nlev; ++l) if (c->lev[l].pdist) out[3] |= 1LL << l;
  return XSB_OK;
}
int xsb_comm_unique_id(void *out128) { return comm_unique_id(out128); }
int xsb_comm_init(xsb_ctx c, const void *unique_id, int rank, int nranks)
{
  NEED_DEVICE(c);
  if (c->assembled) return xsb_fail(c, XSB_ERR_ORDER, "xsb_comm_init must precede xsb_assemble");
  return comm_init(c, unique_id, rank, nranks);
}
int xsb_get_partition(xsb_ctx c, int64_t out[12])
{
  if (!c || !out) return XSB_ERR_ARG;
  if (!c->assembled) return xsb_fail(c, XSB_ERR_ORDER, "xsb_get_partition before xsb_assemble");
  const Slab &S = c->slab; const Lattice &L = c->lat;
  out[0] = S.rank; out[1] = S.nranks; out[2] = S.k0; out[3] = S.k1; out[4] = S.e0; out[5] = S.e1;
  out[6] = c->own_u.off0; out[7] = c->own_u.len0; out[8] = L.nu + c->own_p.off0; out[9] = c->own_p.len0;
  out[10] = (int64_t)(2 * S.k0) * L.NX * L.NY * L.nsd;   /* global index of the first owned velocity dof */
  out[11] = (int64_t)S.k0 * L.PX * L.PY;                 /* global index of the first owned pressure dof */
  return XSB_OK;
}
int xsb_get_stream(xsb_ctx c, void **stream) { if (!c || !stream) return XSB_ERR_ARG; *stream = (void *)c->stream; return XSB_OK; }

int xsb_diagnostics(xsb_ctx c, const double *x, double *out)
{
  NEED_DEVICE(c);
  if (!c->assembled) return xsb_fail(c, XSB_ERR_ORDER, "xsb_diagnostics before xsb_assemble");
  const int64_t n = c->lat.n; const int no = 5 * c->nsd + 5;
  double *dx = nullptr, *dout = nullptr;
  CUDA_OK(cudaMalloc(&dx, sizeof(double) * n)); CUDA_OK(cudaMalloc(&dout, sizeof(double) * no));
  CUDA_OK(cudaMemcpyAsync(dx, x, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
  int rc = vec_diagnostics(c, dx, dout);
  if (!rc) { cudaError_t e = cudaMemcpyAsync(out, dout, sizeof(double) * no, cudaMemcpyDeviceToHost, c->stream); if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream); if (e != cudaSuccess) rc = xsb_fail(c, XSB_ERR_CUDA, "%s", cudaGetErrorString(e)); }
  cudaFree(dx); cudaFree(dout);
  return rc;
}

// ---------------------------------------------------------------- host-side index maps (integer logic only)
static void make_lattice(int nsd, int mx, int my, int mz, Lattice &L)
{
  memset(&L, 0, sizeof(L));
  L.nsd = nsd; L.mx = mx; L.my = my; L.mz = nsd == 3 ? mz : 1;
  L.NX = 2 * mx + 1; L.NY = 2 * my + 1; L.NZ = nsd == 3 ? 2 * mz + 1 : 1;
  L.PX = mx + 1; L.PY = my + 1; L.PZ = nsd == 3 ? mz + 1 : 1;
  L.nun = (int64_t)L.NX * L.NY * L.NZ; L.npn = (int64_t)L.PX * L.PY * L.PZ;
  L.nu = nsd * L.nun; L.np = L.npn; L.n = L.nu + L.np; L.nel = (int64_t)L.mx * L.my * L.mz;
}

int xsb_pattern_row(int nsd, int mx, int my, int mz, int64_t row, int32_t *cols, int cap)
{
  if ((nsd != 2 && nsd != 3) || mx < 1 || my < 1 || (nsd == 3 && mz < 1)) return XSB_ERR_ARG;
  Lattice L; make_lattice(nsd, mx, my, mz, L);
  if (row < 0 || row >= L.n) return XSB_ERR_ARG;
  RowBox b; int comp; row_to_box(L, row, b, &comp);
  const int len = nsd * b.ncu + b.ncp;
  if (!cols) return len;
  int c = 0;
  for (int kk = b.ulo[2]; kk <= b.uhi[2]; ++kk) for (int jj = b.ulo[1]; jj <= b.uhi[1]; ++jj) for (int ii = b.ulo[0]; ii <= b.uhi[0]; ++ii)
    for (int d = 0; d < nsd; ++d) { if (c < cap) cols[c] = (int32_t)(nsd * (ii + (int64_t)jj * L.NX + (int64_t)kk * L.NX * L.NY) + d); c++; }
  for (int kk = b.plo[2]; kk <= b.phi[2]; ++kk) for (int jj = b.plo[1]; jj <= b.phi[1]; ++jj) for (int ii = b.plo[0]; ii <= b.phi[0]; ++ii)
  { if (c < cap) cols[c] = (int32_t)(L.nu + ii + (int64_t)jj * L.PX + (int64_t)kk * L.PX * L.PY); c++; }
  return len;
}

int64_t xsb_prealloc_total(int nsd, int mx, int my, int mz)
{
  Lattice L; make_lattice(nsd, mx, my, mz, L);
  int64_t tot = 0; const int64_t m = L.n;
  for (int k = 0; k < L.NZ; ++k) for (int j = 0; j < L.NY; ++j) for (int i = 0; i < L.NX; ++i) {
    int64_t r;
    if (nsd == 2) { bool vi = i % 2 == 0, vj = j % 2 == 0; r = (vi && vj) ? 2 * 25 + 9 : (vi || vj) ? 2 * 15 + 6 : 2 * 9 + 4; }   // femixedspace.c:205-211
    else { int nmod = i % 2 + j % 2 + k % 2; r = nmod == 0 ? 3 * 125 + 27 : nmod == 1 ? 3 * 75 + 18 : nmod == 2 ? 3 * 45 + 12 : 3 * 27 + 8; }   // :231-244
    if (r > m) r = m;
    tot += nsd * r;
  }
  int64_t r = nsd == 2 ? 2 * 25 + 9 : 3 * 125 + 27; if (r > m) r = m;   // :263, :278
  return tot + L.npn * r;
}

int xsb_bc_list(int nsd, int lame, int model, int freeslip, int mx, int my, int mz, int32_t *idx, double *val, int cap)
{
  return bc_list_faces(nsd, lame, model, freeslip, mx, my, mz, 1, 1, idx, val, cap);
}

}   // extern "C"

// Dirichlet list of a (local) lattice; zlo/zhi say whether its z = 0 / z = max planes are physical boundary faces
int bc_list_faces(int nsd, int lame, int model, int freeslip, int mx, int my, int mz, int zlo, int zhi, int32_t *idx, double *val, int cap)
{
  Lattice L; make_lattice(nsd, mx, my, mz, L);
  const int ni = L.NX, nj = L.NY, nk = L.NZ, N = L.NY, M = L.NX;
  int type = BC_SOLCX;
  if (model < 0) model = lame ? 6 : 2;
  if (lame && model == 8) type = BC_FIXEDBASE;
  if (lame && (model == 9 || model == 10)) type = BC_COMPRESSION;
  if (nsd == 3 && model == 11) type = BC_FIXEDBASE;
  if (lame && nsd == 3 && model == 12) type = BC_COMPRESSION2;
  if (!lame && nsd == 2 && model == 101) type = BC_MMS1;
  int cnt = 0;
  auto push = [&](int i, int j, int k, int d, double v) { if (idx && cnt < cap) { idx[cnt] = nsd * (i + j * ni + k * ni * nj) + d; if (val) val[cnt] = v; } cnt++; };
  switch (type) {
  case BC_SOLCX:   // models.c:52-79 (2-D), :98-149 (3-D)
    if (nsd == 2) {
      for (int j = 0; j < nj; ++j) push(0, j, 0, 0, 0.0);
      for (int i = 0; i < ni; ++i) push(i, 0, 0, 1, 0.0);
      for (int j = 0; j < nj; ++j) push(ni - 1, j, 0, 0, 0.0);
      if (freeslip) for (int i = 0; i < ni; ++i) push(i, nj - 1, 0, 1, 0.0);
    } else {
      for (int j = 0; j < nj; ++j) for (int k = 0; k < nk; ++k) push(0, j, k, 0, 0.0);
      for (int i = 0; i < ni; ++i) for (int k = 0; k < nk; ++k) push(i, 0, k, 1, 0.0);
      if (zlo) for (int i = 0; i < ni; ++i) for (int j = 0; j < nj; ++j) push(i, j, 0, 2, 0.0);
      for (int j = 0; j < nj; ++j) for (int k = 0; k < nk; ++k) push(ni - 1, j, k, 0, 0.0);
      if (freeslip) for (int i = 0; i < ni; ++i) for (int k = 0; k < nk; ++k) push(i, nj - 1, k, 1, 0.0);
      if (zhi) for (int i = 0; i < ni; ++i) for (int j = 0; j < nj; ++j) push(i, j, nk - 1, 2, 0.0);
    }
    break;
  case BC_FIXEDBASE:   // models.c:197-225
    for (int d = 0; d < nsd; ++d) for (int i = 0; i < ni; ++i) for (int k = 0; k < nk; ++k) push(i, 0, k, d, 0.0);
    break;
  case BC_COMPRESSION:   // models.c:276-328 (the x-max test compares with the y count, :270,:288,:315)
    for (int d = 0; d < nsd; ++d) for (int j = 0; j < nj; ++j) for (int k = 0; k < nk; ++k) push(0, j, k, d, d == 0 ? 0.1 : 0.0);
    if (ni == N) for (int d = 0; d < nsd; ++d) for (int j = 0; j < nj; ++j) for (int k = 0; k < nk; ++k) push(ni - 1, j, k, d, d == 0 ? -0.1 : 0.0);
    break;
  case BC_COMPRESSION2:   // models.c:380-446
    for (int j = 0; j < nj; ++j) for (int k = 0; k < nk; ++k) push(0, j, k, 0, 0.1);
    if (ni == N) for (int j = 0; j < nj; ++j) for (int k = 0; k < nk; ++k) push(ni - 1, j, k, 0, -0.1);
    for (int i = 0; i < ni; ++i) for (int k = 0; k < nk; ++k) push(i, 0, k, 1, 0.0);
    if (zlo) for (int i = 0; i < ni; ++i) for (int j = 0; j < nj; ++j) push(i, j, 0, 2, 0.0);
    if (zhi) for (int i = 0; i < ni; ++i) for (int j = 0; j < nj; ++j) push(i, j, nk - 1, 2, 0.0);
    break;
  case BC_MMS1:   // models.c:505-593; values are filled from the coordinates by the caller
    for (int j = 0; j < nj; ++j) for (int d = 0; d < 2; ++d) push(0, j, 0, d, 0.0);
    if (ni == N) for (int j = 0; j < nj; ++j) for (int d = 0; d < 2; ++d) push(ni - 1, j, 0, d, 0.0);
    for (int i = 0; i < ni; ++i) for (int d = 0; d < 2; ++d) push(i, 0, 0, d, 0.0);
    if (nj == M) for (int i = 0; i < ni; ++i) for (int d = 0; d < 2; ++d) push(i, nj - 1, 0, d, 0.0);
    break;
  }
  return cnt;
}

extern "C" {

int xsb_mg_level_dims(int nsd, int mx, int my, int mz, int levels, int level, int dims[3])
{
  if (levels < 1 || level < 0 || level >= levels) return XSB_ERR_ARG;
  int n[3] = {2 * mx + 1, 2 * my + 1, nsd == 3 ? 2 * mz + 1 : 1};
  for (int l = levels - 1; l > level; --l)
    for (int d = 0; d < 3; ++d) {
      if (n[d] == 1 && d == 2 && nsd == 2) continue;
      if ((n[d] - 1) % 2 != 0) return XSB_ERR_ARG;
      n[d] = (n[d] - 1) / 2 + 1;
      if (n[d] < 2) return XSB_ERR_ARG;
    }
  dims[0] = n[0]; dims[1] = n[1]; dims[2] = n[2];
  return XSB_OK;
}

int xsb_slab_layout(int nsd, int mx, int my, int mz, int nranks, int rank, int64_t out[12])
{
  if (nsd != 3 || !out) return XSB_ERR_ARG;
  int k0, k1; if (xsb_slab_range(mz, nranks, rank, &k0, &k1)) return XSB_ERR_ARG;
  const int e0 = nranks > 1 ? (k0 - 2 < 0 ? 0 : k0 - 2) : 0, e1 = nranks > 1 ? (k1 + 1 > mz ? mz : k1 + 1) : mz;
  const bool last = rank == nranks - 1;
  const int64_t NX = 2 * mx + 1, NY = 2 * my + 1, PX = mx + 1, PY = my + 1, nzl = 2 * (e1 - e0) + 1;
  const int64_t pu = 3 * NX * NY, pp = PX * PY, nu_loc = pu * nzl;
  const int ou0 = 2 * (k0 - e0), ou1 = 2 * (k1 - e0) + (last ? 1 : 0), op0 = k0 - e0, op1 = k1 - e0 + (last ? 1 : 0);
  out[0] = rank; out[1] = nranks; out[2] = k0; out[3] = k1; out[4] = e0; out[5] = e1;
  out[6] = ou0 * pu; out[7] = (ou1 - ou0) * pu; out[8] = nu_loc + op0 * pp; out[9] = (op1 - op0) * pp;
  out[10] = (int64_t)(2 * k0) * pu; out[11] = (int64_t)k0 * pp;
  return XSB_OK;
}

int xsb_slab_range(int mz, int nranks, int rank, int *k0, int *k1)
{
  if (nranks < 1 || rank < 0 || rank >= nranks || mz < nranks) return XSB_ERR_ARG;
  const int q = mz / nranks, r = mz % nranks;
  const int s = rank * q + (rank < r ? rank : r);
  if (k0) *k0 = s; if (k1) *k1 = s + q + (rank < r ? 1 : 0);
  return XSB_OK;
}

// Node planes of coarse MG level `depth` below the fine one (depth 0 = first coarse level, mz+1 planes) that rank `rank` of a
// z-slab partition computes: the planes of its element layers (the last rank also the top plane), halved with every further
// coarsening (coarse plane K sits on fine plane 2K).  The ranges of all ranks tile [0, planes of the level).
int xsb_pdist_range(int mz, int nranks, int rank, int depth, int *p0, int *p1)
{
  int k0, k1; if (depth < 0 || xsb_slab_range(mz, nranks, rank, &k0, &k1)) return XSB_ERR_ARG;
  if (rank == nranks - 1) k1 = mz + 1;
  for (int d = 0; d < depth; ++d) { k0 = (k0 + 1) / 2; k1 = (k1 + 1) / 2; }
  if (p0) *p0 = k0; if (p1) *p1 = k1;
  return XSB_OK;
}

}   // extern "C"
