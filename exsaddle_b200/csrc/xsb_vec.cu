// xsb_vec.cu -- fused Krylov vector kernels (K7 of SURVEY 2.1): replace PETSc VecMDot / VecMAXPY /
// VecDotNorm2 / VecAXPBYPCZ / VecNorm / VecStrideNormAll on the solve path.  All are one-pass HBM-bound
// streams; reductions are two-stage with a fixed grid (148 SMs x 4 CTAs) so results are run-to-run
// deterministic.  Scalars produced by a reduction stay on the device and are consumed by the next kernel
// (no host round trip between VecMDot and VecMAXPY, or between VecDotNorm2 and the GCR update).
#include "xsb.h"

#define RB 592          // reduction grid: 148 SMs x 4
#define RT 256
#define MD 8            // vectors per multi-dot / multi-axpy pass

static inline unsigned gridfor(int64_t n) { int64_t b = (n + RT - 1) / RT; if (b > 148 * 16) b = 148 * 16; return (unsigned)(b < 1 ? 1 : b); }

__global__ void k_set(int64_t n, double a, double *x) { for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) x[i] = a; }
__global__ void k_axpy(int64_t n, double a, const double *__restrict__ x, double *__restrict__ y) { for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) y[i] += a * x[i]; }
__global__ void k_aypx(int64_t n, double a, const double *__restrict__ x, double *__restrict__ y) { for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) y[i] = x[i] + a * y[i]; }
__global__ void k_scale(int64_t n, double a, double *x) { for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) x[i] *= a; }
__global__ void k_pmult(int64_t n, const double *__restrict__ d, const double *__restrict__ x, double *__restrict__ y) { for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) y[i] = d[i] * x[i]; }
__global__ void k_waxpy(int64_t n, double a, const double *__restrict__ x, const double *__restrict__ y, double *__restrict__ w) { for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) w[i] = y[i] + a * x[i]; }

int vec_set(xsb_ctx c, int64_t n, double a, double *x) { if (n <= 0) return 0; if (a == 0.0) { CUDA_OK(cudaMemsetAsync(x, 0, sizeof(double) * n, c->stream)); return 0; } k_set<<<gridfor(n), RT, 0, c->stream>>>(n, a, x); KERNEL_OK(); return 0; }
int vec_copy(xsb_ctx c, int64_t n, const double *x, double *y) { if (n <= 0 || x == y) return 0; CUDA_OK(cudaMemcpyAsync(y, x, sizeof(double) * n, cudaMemcpyDeviceToDevice, c->stream)); return 0; }
int vec_axpy(xsb_ctx c, int64_t n, double a, const double *x, double *y) { if (n <= 0) return 0; k_axpy<<<gridfor(n), RT, 0, c->stream>>>(n, a, x, y); KERNEL_OK(); return 0; }
int vec_aypx(xsb_ctx c, int64_t n, double a, const double *x, double *y) { if (n <= 0) return 0; k_aypx<<<gridfor(n), RT, 0, c->stream>>>(n, a, x, y); KERNEL_OK(); return 0; }
int vec_scale(xsb_ctx c, int64_t n, double a, double *x) { if (n <= 0) return 0; k_scale<<<gridfor(n), RT, 0, c->stream>>>(n, a, x); KERNEL_OK(); return 0; }
int vec_pmult(xsb_ctx c, int64_t n, const double *d, const double *x, double *y) { if (n <= 0) return 0; k_pmult<<<gridfor(n), RT, 0, c->stream>>>(n, d, x, y); KERNEL_OK(); return 0; }
int vec_waxpy(xsb_ctx c, int64_t n, double a, const double *x, const double *y, double *w) { if (n <= 0) return 0; k_waxpy<<<gridfor(n), RT, 0, c->stream>>>(n, a, x, y, w); KERNEL_OK(); return 0; }

// ------------------------------------------------------------------ reductions
struct PtrPack { const double *p[MD]; };

__device__ __forceinline__ double block_sum(double v, double *sh)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  double s = 0.0;
  if (threadIdx.x == 0) for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += sh[i];
  return s;   // valid on thread 0
}

// partial[j*RB + block] = sum over the block's slice of w[i]*V_j[i]   (j < m)
template <int M>
__global__ void __launch_bounds__(RT) k_mdot_partial(Ranges rg, const double *__restrict__ w, PtrPack V, double *__restrict__ partial)
{
  __shared__ double sh[RT / 32];
  double acc[M];
#pragma unroll
  for (int j = 0; j < M; ++j) acc[j] = 0.0;
  const int64_t n = rg.len0 + rg.len1;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = t < rg.len0 ? rg.off0 + t : rg.off1 + (t - rg.len0);   // owned entries only
    const double wi = w[i];
#pragma unroll
    for (int j = 0; j < M; ++j) acc[j] += wi * V.p[j][i];
  }
#pragma unroll
  for (int j = 0; j < M; ++j) { double s = block_sum(acc[j], sh); if (threadIdx.x == 0) partial[(int64_t)j * RB + blockIdx.x] = s; }
}
// out[j] = sum_b partial[j*RB + b]; one CTA per j
__global__ void __launch_bounds__(RT) k_reduce_final(int nb, const double *__restrict__ partial, double *__restrict__ out)
{
  __shared__ double sh[RT / 32];
  double v = 0.0;
  for (int b = threadIdx.x; b < nb; b += blockDim.x) v += partial[(int64_t)blockIdx.x * RB + b];
  double s = block_sum(v, sh);
  if (threadIdx.x == 0) out[blockIdx.x] = s;
}

template <int M> static int mdot_launch(xsb_ctx c, const Ranges &rg, const double *w, const PtrPack &P, double *out)
{
  k_mdot_partial<M><<<RB, RT, 0, c->stream>>>(rg, w, P, c->red); KERNEL_OK();
  k_reduce_final<<<M, RT, 0, c->stream>>>(RB, c->red, out); KERNEL_OK();
  return 0;
}

// out[j] = w . V[j] for j < k ; out[k] = w . w when with_norm (VecMDot + VecNorm^2 in one or few passes)
int vec_mdot(xsb_ctx c, const Ranges &n, const double *w, double *const *V, int k, bool with_norm, double *out, bool local)
{
  const int tot = k + (with_norm ? 1 : 0);
  for (int j0 = 0; j0 < tot; j0 += MD) {
    const int m = tot - j0 < MD ? tot - j0 : MD;
    PtrPack P;
    for (int j = 0; j < MD; ++j) { int g = j0 + j; P.p[j] = g < k ? V[g] : w; }
    switch (m) {
    case 1: XSB_CHK(mdot_launch<1>(c, n, w, P, out + j0)); break;
    case 2: XSB_CHK(mdot_launch<2>(c, n, w, P, out + j0)); break;
    case 3: XSB_CHK(mdot_launch<3>(c, n, w, P, out + j0)); break;
    case 4: XSB_CHK(mdot_launch<4>(c, n, w, P, out + j0)); break;
    case 5: XSB_CHK(mdot_launch<5>(c, n, w, P, out + j0)); break;
    case 6: XSB_CHK(mdot_launch<6>(c, n, w, P, out + j0)); break;
    case 7: XSB_CHK(mdot_launch<7>(c, n, w, P, out + j0)); break;
    default: XSB_CHK(mdot_launch<8>(c, n, w, P, out + j0)); break;
    }
  }
  if (local) return 0;
  return comm_allreduce_sum(c, out, tot);   // VecMDot's MPI_Allreduce: one NCCL all-reduce of the whole block of partial sums
}

// w += sign * sum_j coef[j] V_j, applied in ascending j like a sequence of VecAXPY (VecMAXPY)
template <int M>
__global__ void __launch_bounds__(RT) k_maxpy(int64_t n, double *__restrict__ w, PtrPack V, const double *__restrict__ coef, double sign)
{
  double cf[M];
#pragma unroll
  for (int j = 0; j < M; ++j) cf[j] = sign * coef[j];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double wi = w[i];
#pragma unroll
    for (int j = 0; j < M; ++j) wi += cf[j] * V.p[j][i];
    w[i] = wi;
  }
}
int vec_maxpy_dev(xsb_ctx c, int64_t n, double *w, double *const *V, int k, const double *coef, double sign)
{
  for (int j0 = 0; j0 < k; j0 += MD) {
    const int m = k - j0 < MD ? k - j0 : MD;
    PtrPack P; for (int j = 0; j < MD; ++j) P.p[j] = V[j0 + (j < m ? j : 0)];
    switch (m) {
    case 1: k_maxpy<1><<<gridfor(n), RT, 0, c->stream>>>(n, w, P, coef + j0, sign); break;
    case 2: k_maxpy<2><<<gridfor(n), RT, 0, c->stream>>>(n, w, P, coef + j0, sign); break;
    case 3: k_maxpy<3><<<gridfor(n), RT, 0, c->stream>>>(n, w, P, coef + j0, sign); break;
    case 4: k_maxpy<4><<<gridfor(n), RT, 0, c->stream>>>(n, w, P, coef + j0, sign); break;
    case 5: k_maxpy<5><<<gridfor(n), RT, 0, c->stream>>>(n, w, P, coef + j0, sign); break;
    case 6: k_maxpy<6><<<gridfor(n), RT, 0, c->stream>>>(n, w, P, coef + j0, sign); break;
    case 7: k_maxpy<7><<<gridfor(n), RT, 0, c->stream>>>(n, w, P, coef + j0, sign); break;
    default: k_maxpy<8><<<gridfor(n), RT, 0, c->stream>>>(n, w, P, coef + j0, sign); break;
    }
    KERNEL_OK();
  }
  return 0;
}
int vec_maxpy_host(xsb_ctx c, int64_t n, double *w, double *const *V, int k, const double *coef_host)
{
  if (k <= 0) return 0;
  // stage coefficients through the scalar scratch (second half, first half holds live dot products)
  double *dst = c->scal + 128;
  CUDA_OK(cudaMemcpyAsync(dst, coef_host, sizeof(double) * k, cudaMemcpyHostToDevice, c->stream));
  return vec_maxpy_dev(c, n, w, V, k, dst, 1.0);
}

__global__ void k_scale_inv_sqrt(int64_t n, double *w, const double *nrm2)
{
  const double s = *nrm2; const double f = s > 0.0 ? 1.0 / sqrt(s) : 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) w[i] *= f;
}
int vec_scale_by_inv_sqrt(xsb_ctx c, int64_t n, double *w, const double *nrm2) { k_scale_inv_sqrt<<<gridfor(n), RT, 0, c->stream>>>(n, w, nrm2); KERNEL_OK(); return 0; }

// GCR update (KSPSolve_GCR_cycle): nrm = sqrt(v.v); a = (r.v)/nrm; v /= nrm; s /= nrm; x += a s; r -= a v; also ||r||^2
__global__ void __launch_bounds__(RT) k_gcr_update(Ranges rg, const double *__restrict__ dots, double *__restrict__ v, double *__restrict__ s,
                                                   double *__restrict__ x, double *__restrict__ r, double *__restrict__ partial)
{
  __shared__ double sh[RT / 32];
  const double nrm = sqrt(dots[1]), a = dots[0] / nrm, inv = 1.0 / nrm;
  double acc = 0.0;
  const int64_t n = rg.len0 + rg.len1;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = t < rg.len0 ? rg.off0 + t : rg.off1 + (t - rg.len0);
    const double vi = v[i] * inv, si = s[i] * inv;
    v[i] = vi; s[i] = si;
    x[i] += a * si;
    const double ri = r[i] + (-a) * vi;
    r[i] = ri; acc += ri * ri;
  }
  double t = block_sum(acc, sh);
  if (threadIdx.x == 0) partial[blockIdx.x] = t;
}
int vec_gcr_update(xsb_ctx c, const Ranges &rg, const double *dots, double *v, double *s, double *x, double *r, double *rnorm2)
{
  k_gcr_update<<<RB, RT, 0, c->stream>>>(rg, dots, v, s, x, r, c->red); KERNEL_OK();
  k_reduce_final<<<1, RT, 0, c->stream>>>(RB, c->red, rnorm2); KERNEL_OK();
  return comm_allreduce_sum(c, rnorm2, 1);
}

int vec_fetch(xsb_ctx c, const double *dev, int n, double *host)
{
  CUDA_OK(cudaMemcpyAsync(c->red_h, dev, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
  CUDA_OK(cudaStreamSynchronize(c->stream));
  memcpy(host, c->red_h, sizeof(double) * n);
  return 0;
}

// ------------------------------------------------------------------ diagnostics (exSaddle_io.c:7-58)
// out[5*nsd+5]; one CTA per field (nsd velocity components + pressure)
__global__ void __launch_bounds__(1024) k_diag(int nsd, int64_t nun, int64_t np, const double *__restrict__ x, double *__restrict__ out)
{
  __shared__ double s1[32], s2[32], si[32], smn[32], smx[32];
  const int f = blockIdx.x; const bool isp = f == nsd;
  const int64_t cnt = isp ? np : nun; const int64_t base = isp ? nsd * nun : f; const int stride = isp ? 1 : nsd;
  double n1 = 0, n2 = 0, ni = 0, mn = 1.7976931348623157e308, mx = -1.7976931348623157e308;
  for (int64_t i = threadIdx.x; i < cnt; i += blockDim.x) { double v = x[base + i * stride]; n1 += fabs(v); n2 += v * v; ni = fmax(ni, fabs(v)); mn = fmin(mn, v); mx = fmax(mx, v); }
  for (int o = 16; o > 0; o >>= 1) {
    n1 += __shfl_xor_sync(0xffffffffu, n1, o); n2 += __shfl_xor_sync(0xffffffffu, n2, o);
    ni = fmax(ni, __shfl_xor_sync(0xffffffffu, ni, o)); mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o)); mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { s1[w] = n1; s2[w] = n2; si[w] = ni; smn[w] = mn; smx[w] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 32; ++i) { n1 += s1[i]; n2 += s2[i]; ni = fmax(ni, si[i]); mn = fmin(mn, smn[i]); mx = fmax(mx, smx[i]); }
    if (isp) { double *o = out + 5 * nsd; o[0] = n1; o[1] = sqrt(n2); o[2] = ni; o[3] = mn; o[4] = mx; }
    else { out[0 * nsd + f] = n1; out[1 * nsd + f] = sqrt(n2); out[2 * nsd + f] = ni; out[3 * nsd + f] = mn; out[4 * nsd + f] = mx; }
  }
}
int vec_diagnostics(xsb_ctx c, const double *x, double *out_dev)
{
  k_diag<<<c->nsd + 1, 1024, 0, c->stream>>>(c->nsd, c->lat.nun, c->lat.np, x, out_dev); KERNEL_OK();
  return 0;
}

// ------------------------------------------------------------------ PetscRandom "rander48" stream on the device
// value i = X_{i+1} / 2^48 with X_{k+1} = (a X_k + c) mod 2^48, X_0 = 0x12345678<<16 | 0x330E (drand48 seeding);
// each thread jumps ahead with the composed affine map and then walks 64 consecutive values.
__global__ void k_rander48(int64_t n, int interval, double *x, int64_t soff)
{
  const uint64_t a = 0x5DEECE66DULL, cc = 0xBULL, mask = (1ULL << 48) - 1;
  const int64_t chunk = 64, i0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * chunk;
  if (i0 >= n) return;
  uint64_t A = 1, C = 0, pa = a, pc = cc; uint64_t k = (uint64_t)(i0 + soff);   // f^k = A x + C; soff = global index of local entry 0
  while (k) { if (k & 1) { A = (A * pa) & mask; C = (C * pa + pc) & mask; } pc = (pc * pa + pc) & mask; pa = (pa * pa) & mask; k >>= 1; }
  uint64_t X = (A * ((0x12345678ULL << 16) | 0x330EULL) + C) & mask;
  for (int64_t i = i0; i < n && i < i0 + chunk; ++i) {
    X = (a * X + cc) & mask;
    const double u = (double)X * (1.0 / 281474976710656.0);
    x[i] = interval ? 2.0 * u - 1.0 : u;
  }
}
int vec_rander48(xsb_ctx c, int64_t n, int interval, double *x, int64_t soff)
{
  const int64_t threads = (n + 63) / 64;
  k_rander48<<<(unsigned)((threads + 127) / 128), 128, 0, c->stream>>>(n, interval, x, soff); KERNEL_OK();
  return 0;
}
