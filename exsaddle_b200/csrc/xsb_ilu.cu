// xsb_ilu.cu -- ILU(0) of the scaled pressure mass matrix and its triangular solves (K10 of SURVEY 2.1).
//
// Replaces -saddle_fieldsplit_p_pc_type bjacobi (one block per rank, sub-PC ILU(0) in natural ordering,
// abf.opts:15; PETSc MatLUFactorNumeric_SeqAIJ / MatSolve_SeqAIJ) applied to Mpscaled
// (MatAssemble_Schur, femixedspace.c:2837).  Mp is a 27-point (9-point in 2-D) stencil on the pressure
// lattice, so row (i,j,k) depends only on rows with a smaller wavefront number w = i + 2j + 4k: all rows
// of one wavefront are independent and the natural-ordering factorisation / substitutions are reproduced
// exactly (same operations per row, same order within a row) by sweeping wavefronts.
//
// The solve is latency bound (449 dependent wavefronts per sweep at 64^3, ~600 rows each, 13 FMAs per row).
// It runs as ONE thread-block cluster of 8 CTAs x 1024 threads: rows of a wavefront are spread over the
// cluster, wavefronts are separated by the hardware cluster barrier (barrier.cluster, ~0.3 us) instead of
// a kernel launch or a grid-wide barrier, and both sweeps are one launch.  The factors are repacked at set-up
// into wavefront order (SoA per wavefront) so every access is coalesced, independent of x and prefetchable;
// x is exchanged through L2 (ld.global.cg / st.global.cg).
#include "xsb.h"
#include <cooperative_groups.h>
#include <cub/cub.cuh>
namespace cg = cooperative_groups;

struct PLat { int px, py, pz; };
#define ILU_CLUSTER 8
#define ILU_TPB 128

__device__ __forceinline__ bool wf_row(const PLat &P, int w, int t, int &i, int &j, int &k)
{
  j = t % P.py; k = t / P.py;
  if (k >= P.pz) return false;
  i = w - 2 * j - 4 * k;
  return i >= 0 && i < P.px;
}

// Factorisation (set-up, once): row-wise IKJ restricted to the pattern; multiplier = a_ik * (1/a_kk); the
// diagonal is stored inverted.  One CTA, block barrier per wavefront.
__global__ void __launch_bounds__(1024) k_ilu0_factor(PLat P, const int *__restrict__ ia, const double *__restrict__ a, double *lu, int *flag)
{
  const BoxPattern pat{P.px, P.py, P.pz, 0};
  const int nw = (P.px - 1) + 2 * (P.py - 1) + 4 * (P.pz - 1) + 1, ncand = P.py * P.pz;
  for (int w = 0; w < nw; ++w) {
    for (int t = threadIdx.x; t < ncand; t += blockDim.x) {
      int i, j, k; if (!wf_row(P, w, t, i, j, k)) continue;
      const int row = i + j * P.px + k * P.px * P.py;
      int l0, h0, l1, h1, l2, h2; range_pp(i, P.px, l0, h0); range_pp(j, P.py, l1, h1); range_pp(k, P.pz, l2, h2);
      const int nx = h0 - l0 + 1, ny = h1 - l1 + 1, nrow = nx * ny * (h2 - l2 + 1), r0 = ia[row];
      double wv[27];
      for (int s = 0; s < nrow; ++s) wv[s] = a[r0 + s];
      const int dslot = ((k - l2) * ny + (j - l1)) * nx + (i - l0);
      for (int s = 0; s < dslot; ++s) {   // lower entries in ascending column order
        const int ci = l0 + s % nx, cj = l1 + (s / nx) % ny, ck = l2 + s / (nx * ny);
        const int crow = ci + cj * P.px + ck * P.px * P.py;
        int m0, g0, m1, g1, m2, g2; range_pp(ci, P.px, m0, g0); range_pp(cj, P.py, m1, g1); range_pp(ck, P.pz, m2, g2);
        const int cnx = g0 - m0 + 1, cny = g1 - m1 + 1, cn = cnx * cny * (g2 - m2 + 1), c0 = ia[crow];
        const int cd = ((ck - m2) * cny + (cj - m1)) * cnx + (ci - m0);
        double mult = wv[s];
        if (mult != 0.0) {
          mult = mult * lu[c0 + cd];   // lu[diag] = 1/pivot
          wv[s] = mult;
          for (int u = cd + 1; u < cn; ++u) {   // U part of the pivot row
            const int gi = m0 + u % cnx, gj = m1 + (u / cnx) % cny, gk = m2 + u / (cnx * cny);
            const int sl = box_slot(pat, i, j, k, gi, gj, gk);
            if (sl >= 0) wv[sl] -= mult * lu[c0 + u];   // fill outside the pattern is dropped (ILU(0))
          }
        }
      }
      if (wv[dslot] == 0.0) *flag = 1;
      for (int s = 0; s < nrow; ++s) lu[r0 + s] = s == dslot ? 1.0 / wv[dslot] : wv[s];
    }
    __syncthreads();
  }
}

// ---- level schedule (set-up): rows sorted by wavefront number
__global__ void k_lvl_count(PLat P, int *cnt)
{
  int row = blockIdx.x * blockDim.x + threadIdx.x; if (row >= P.px * P.py * P.pz) return;
  const int i = row % P.px, j = (row / P.px) % P.py, k = row / (P.px * P.py);
  atomicAdd(&cnt[i + 2 * j + 4 * k], 1);
}
__global__ void k_lvl_fill(PLat P, const int *off, int *cursor, int *rows, int *diag, const int *ia)
{
  int row = blockIdx.x * blockDim.x + threadIdx.x; if (row >= P.px * P.py * P.pz) return;
  const int i = row % P.px, j = (row / P.px) % P.py, k = row / (P.px * P.py), w = i + 2 * j + 4 * k;
  rows[off[w] + atomicAdd(&cursor[w], 1)] = row;   // order inside a wavefront is irrelevant: its rows are independent
  int l0, h0, l1, h1, l2, h2; range_pp(i, P.px, l0, h0); range_pp(j, P.py, l1, h1); range_pp(k, P.pz, l2, h2);
  diag[row] = ia[row] + ((k - l2) * (h1 - l1 + 1) + (j - l1)) * (h0 - l0 + 1) + (i - l0);
}

// Level-packed factors (set-up): for wavefront w with cnt rows starting at off[w], entry u of local row l lives at
// 13*off[w] + u*cnt + l (SoA per wavefront => coalesced, and affine in the wavefront order => prefetchable).
// Lower entries ascending, upper entries DEscending (the order MatSolve consumes them); unused slots have n < u.
__global__ void k_ilu_pack(int nw, const int *__restrict__ off, const int *__restrict__ rows, const int *__restrict__ ia, const int *__restrict__ ja,
                           const int *__restrict__ diag, const double *__restrict__ lu, const int *__restrict__ lvl_of_q,
                           int *fcol, double *fval, unsigned char *fn, int *bcol, double *bval, unsigned char *bn, double *binv)
{
  const int q = blockIdx.x * blockDim.x + threadIdx.x; if (q >= off[nw]) return;
  const int w = lvl_of_q[q], cnt = off[w + 1] - off[w], l = q - off[w];
  const int64_t base = (int64_t)13 * off[w] + l;
  const int row = rows[q], r0 = ia[row], r1 = ia[row + 1], d = diag[row];
  const int nl = d - r0, nu = r1 - 1 - d;
  fn[q] = (unsigned char)nl; bn[q] = (unsigned char)nu; binv[q] = lu[d];
  for (int u = 0; u < 13; ++u) {
    fcol[base + (int64_t)u * cnt] = u < nl ? ja[r0 + u] : row; fval[base + (int64_t)u * cnt] = u < nl ? lu[r0 + u] : 0.0;
    bcol[base + (int64_t)u * cnt] = u < nu ? ja[r1 - 1 - u] : row; bval[base + (int64_t)u * cnt] = u < nu ? lu[r1 - 1 - u] : 0.0;
  }
}
__global__ void k_lvl_of_q(int nw, const int *__restrict__ off, int *lvl_of_q)
{
  const int w = blockIdx.x; if (w >= nw) return;
  for (int q = off[w] + threadIdx.x; q < off[w + 1]; q += blockDim.x) lvl_of_q[q] = w;
}

__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// x = U^-1 L^-1 b (MatSolve_SeqAIJ): forward with unit L, backward with U and the inverted diagonal.
// One cluster of 8 CTAs with few warps each (the cluster barrier's acquire invalidates L1 once per warp, so its
// cost grows with the warp count: 6.3 ms/apply with 32 warps per CTA, 3.2 ms with 4).  Software pipelining: the
// factors of the NEXT wavefront (independent of x) are loaded into registers before the barrier of the current
// one, so after a barrier only the x values written by earlier wavefronts (L2 hits) are on the critical path.
struct IluRow { double lv[13]; int cv[13]; int row, n; double aux; };

template <bool FWD>
__device__ __forceinline__ void ilu_load(IluRow &R, int o, int cnt, int l, const int *__restrict__ rows, const int *__restrict__ col,
                                         const double *__restrict__ val, const unsigned char *__restrict__ nn, const double *__restrict__ aux)
{
  const int q = o + l; const int64_t base = (int64_t)13 * o + l;
  R.row = rows[q]; R.n = nn[q];
#pragma unroll
  for (int u = 0; u < 13; ++u) { R.lv[u] = val[base + (int64_t)u * cnt]; R.cv[u] = col[base + (int64_t)u * cnt]; }
  R.aux = FWD ? aux[R.row] : aux[q];   // forward: b[row]; backward: 1/pivot
}
template <bool FWD>
__device__ __forceinline__ void ilu_row(const IluRow &R, double *x)
{
  double xv[13];
#pragma unroll
  for (int u = 0; u < 13; ++u) xv[u] = u < R.n ? __ldcg(x + R.cv[u]) : 0.0;
  double s = FWD ? R.aux : __ldcg(x + R.row);
#pragma unroll
  for (int u = 0; u < 13; ++u) if (u < R.n) s -= R.lv[u] * xv[u];   // column order of the sequential sweep (packed that way)
  __stcg(x + R.row, FWD ? s : s * R.aux);
}

__global__ void __cluster_dims__(ILU_CLUSTER, 1, 1) __launch_bounds__(ILU_TPB)
k_ilu0_solve(int nw, int np, const int *__restrict__ off, const int *__restrict__ rows,
             const int *__restrict__ fcol, const double *__restrict__ fval, const unsigned char *__restrict__ fn,
             const int *__restrict__ bcol, const double *__restrict__ bval, const unsigned char *__restrict__ bn, const double *__restrict__ binv,
             const double *__restrict__ b, double *x)
{
  extern __shared__ int soff[];   // nw + 1 wavefront offsets
  cg::cluster_group cl = cg::this_cluster();
  const int gt = (int)cl.block_rank() * ILU_TPB + threadIdx.x, gn = ILU_CLUSTER * ILU_TPB;
  for (int i = threadIdx.x; i <= nw; i += ILU_TPB) soff[i] = off[i];
  // L2 prefetch of everything the two sweeps stream (128-byte lines), issued in sweep order
  {
    const int64_t nv = (int64_t)13 * np;
    for (int64_t i = (int64_t)gt * 16; i < nv; i += (int64_t)gn * 16) prefetch_l2(fval + i);
    for (int64_t i = (int64_t)gt * 32; i < nv; i += (int64_t)gn * 32) prefetch_l2(fcol + i);
    for (int64_t i = (int64_t)gt * 16; i < np; i += (int64_t)gn * 16) prefetch_l2(b + i);
    for (int64_t i = (int64_t)gt * 32; i < np; i += (int64_t)gn * 32) prefetch_l2(rows + i);
    for (int64_t i = nv - 16 - (int64_t)gt * 16; i >= 0; i -= (int64_t)gn * 16) prefetch_l2(bval + i);
    for (int64_t i = nv - 32 - (int64_t)gt * 32; i >= 0; i -= (int64_t)gn * 32) prefetch_l2(bcol + i);
  }
  __syncthreads();
  IluRow cur, nxt;
  bool have = gt < soff[1] - soff[0];
  if (have) ilu_load<true>(cur, soff[0], soff[1] - soff[0], gt, rows, fcol, fval, fn, b);
  for (int w = 0; w < nw; ++w) {
    const int o = soff[w], cnt = soff[w + 1] - o;
    bool have_next = false;
    if (w + 1 < nw) { const int o1 = soff[w + 1], c1 = soff[w + 2] - o1; have_next = gt < c1; if (have_next) ilu_load<true>(nxt, o1, c1, gt, rows, fcol, fval, fn, b); }
    if (have) ilu_row<true>(cur, x);
    for (int l = gt + gn; l < cnt; l += gn) { IluRow t; ilu_load<true>(t, o, cnt, l, rows, fcol, fval, fn, b); ilu_row<true>(t, x); }
    cl.sync();
    cur = nxt; have = have_next;
  }
  have = gt < soff[nw] - soff[nw - 1];
  if (have) ilu_load<false>(cur, soff[nw - 1], soff[nw] - soff[nw - 1], gt, rows, bcol, bval, bn, binv);
  for (int w = nw - 1; w >= 0; --w) {
    const int o = soff[w], cnt = soff[w + 1] - o;
    bool have_next = false;
    if (w > 0) { const int o1 = soff[w - 1], c1 = soff[w] - o1; have_next = gt < c1; if (have_next) ilu_load<false>(nxt, o1, c1, gt, rows, bcol, bval, bn, binv); }
    if (have) ilu_row<false>(cur, x);
    for (int l = gt + gn; l < cnt; l += gn) { IluRow t; ilu_load<false>(t, o, cnt, l, rows, bcol, bval, bn, binv); ilu_row<false>(t, x); }
    cl.sync();
    cur = nxt; have = have_next;
  }
}

// ------------------------------------------------------------------ line-pipelined triangular solves (default)
// The wavefront kernel above needs a cluster barrier and an L2 round trip per wavefront: 2.45 ms per apply at 64^3 and
// 7.2 ms on a 128^3 two-slab block (13 % of that solve), 100 x off the time the factors take to stream.  The same
// sequential sweep is reorganised so that the dependent chain runs through registers and shared memory:
//   * one CTA per pressure-node plane k, one thread per node line (j, k); a thread marches along i keeping x(i-1) in a
//     register; the three values of line j-1 it needs were produced by the neighbouring thread one to three steps earlier
//     and travel through a 4-deep ring in shared memory -- one block barrier per step (step t handles i = t - 2 j);
//   * the nine values of plane k-1 come from the CTA below, which runs a few steps ahead; they are fetched from L2 four steps
//     before use into a register ring, so the L2 latency is off the chain (hand-over protocol: see "the value is the flag");
//   * the factors are repacked at set-up into one 128-byte record per row, ordered (plane, step, line): a thread streams
//     its records with cp.async three steps ahead (prefetched into L2 32 steps ahead).
// The backward sweep is the forward sweep on the mirrored lattice (i, j, k -> px-1-i, ...).  Per row the operations and
// their order are those of MatSolve_SeqAIJ (ascending columns forward, descending backward): results equal the wavefront
// kernel's bit for bit.  Chain length: (px + 2 py) steps of ~0.15 us for the first plane + ~12 steps of lag per plane.
#define ILUP_PF 4      // register-ring prefetch distance of the plane-below values (steps; a step is ~0.4 us, an L2 round trip ~0.7 us)
#define ILUP_D 3       // cp.async distance of the factor records (steps); ring of ILUP_D + 1 slots
struct IluPipe { int px, py, pz, S, W; };
__device__ __forceinline__ unsigned ld_acquire_gpu_u32(const unsigned *p) { unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_release_gpu_u32(unsigned *p, unsigned v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void cp_async16(void *smem, const void *g) { const unsigned sa = (unsigned)__cvta_generic_to_shared(smem); asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(sa), "l"(g) : "memory"); }
__host__ __device__ __forceinline__ int ilup_jlo(int s, int px) { const int v = s - (px - 1); return v <= 0 ? 0 : (v + 1) >> 1; }

// one 128-byte record per row and direction: [0..12] factors of the 13 earlier neighbours in sweep order, [13] 1/pivot (backward)
__global__ void k_ilup_pack(IluPipe P, const int *__restrict__ ia, const int *__restrict__ diag, const double *__restrict__ lu, double *__restrict__ packf, double *__restrict__ packb)
{
  const int row = blockIdx.x * blockDim.x + threadIdx.x; if (row >= P.px * P.py * P.pz) return;
  const BoxPattern pat{P.px, P.py, P.pz, 0};
  const int i = row % P.px, j = (row / P.px) % P.py, k = row / (P.px * P.py);
  for (int dir = 0; dir < 2; ++dir) {
    const int ib = dir ? P.px - 1 - i : i, jb = dir ? P.py - 1 - j : j, kb = dir ? P.pz - 1 - k : k;
    const int st = ib + 2 * jb;
    double *rec = (dir ? packb : packf) + ((((int64_t)kb * P.S + st) * P.W) + (jb - ilup_jlo(st, P.px))) * 16;
    for (int u = 0; u < 13; ++u) {
      const int di = u < 12 ? u % 3 - 1 : -1, dj = u < 9 ? (u / 3) % 3 - 1 : (u < 12 ? -1 : 0), dk = u < 9 ? -1 : 0;
      const int nib = ib + di, njb = jb + dj, nkb = kb + dk;
      double v = 0.0;
      if (nib >= 0 && nib < P.px && njb >= 0 && njb < P.py && nkb >= 0 && nkb < P.pz) {
        const int ni = dir ? P.px - 1 - nib : nib, nj = dir ? P.py - 1 - njb : njb, nk = dir ? P.pz - 1 - nkb : nkb;
        v = lu[ia[row] + box_slot(pat, i, j, k, ni, nj, nk)];
      }
      rec[u] = v;
    }
    rec[13] = dir ? lu[diag[row]] : 0.0; rec[14] = 0.0; rec[15] = 0.0;
  }
}

// Hand-over between planes without flags or fences: THE VALUE IS THE FLAG.  Both sweeps write into arrays that a small kernel
// has filled with a sentinel (a quiet NaN with a payload no computation produces) beforehand; a consumer that fetches a value
// of the plane below and finds the sentinel fetches it again (ld.relaxed.gpu, served by L2) until the real value has
// arrived.  An 8-byte store is single-copy atomic and nothing else depends on it, so no release / acquire pair is needed (a
// first version published step counters with st.release.gpu: the gpu-scope fence behind it cost ~2.8 us per publication and
// bounded the whole pipeline).  The forward sweep writes y into a scratch vector, the backward sweep reads y and writes x,
// so "not yet written" is distinguishable in both.  A fetch that waits longer than ~2 s raises a sticky error word.
#define ILUP_SENTINEL 0x7FF8DEADBEEF1234ULL
__device__ __forceinline__ double ld_relaxed_gpu_f64(const double *p) { double v; asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_relaxed_gpu_f64(double *p, double v) { asm volatile("st.relaxed.gpu.global.f64 [%0], %1;" :: "l"(p), "d"(v) : "memory"); }
__device__ __forceinline__ bool ilup_missing(double v) { return (unsigned long long)__double_as_longlong(v) == ILUP_SENTINEL; }
__device__ __noinline__ double ilup_refetch(const double *p, unsigned *err)
{
  const long long t0 = clock64(); double v;
  do { v = ld_relaxed_gpu_f64(p); if (clock64() - t0 > 4000000000LL) { atomicExch(err, 1u); return 0.0; } } while (ilup_missing(v));
  return v;
}
__global__ void k_ilup_fill(int64_t n, double *a, double *b)
{
  const double s = __longlong_as_double((long long)ILUP_SENTINEL);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) { a[i] = s; b[i] = s; }
}

// rhs: right-hand side of the sweep (b forward, y backward); out: the vector this sweep writes and reads the plane below from
template <bool BWD>
__device__ __forceinline__ void ilup_sweep(const IluPipe P, const int kb, const double *__restrict__ pack, const double *__restrict__ rhs, double *out,
                                           unsigned *err, double2 *rows, double *xs)
{
  const int T = (int)blockDim.x, t = (int)threadIdx.x, jb = t;
  for (int a = 0; a < 4; ++a) xs[t * 4 + a] = 0.0;
  __syncthreads();
  const bool line = jb < P.py;
  const int64_t plane = (int64_t)P.px * P.py;
  auto ridx = [&](int ib, int jj, int kk) -> int64_t {
    const int i = BWD ? P.px - 1 - ib : ib, j = BWD ? P.py - 1 - jj : jj, k = BWD ? P.pz - 1 - kk : kk;
    return i + (int64_t)j * P.px + (int64_t)k * plane; };
  double w[3][3], ring[ILUP_PF][4], xprev = 0.0;
#pragma unroll
  for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) w[a][b] = 0.0;
#pragma unroll
  for (int a = 0; a < ILUP_PF; ++a) for (int b = 0; b < 4; ++b) ring[a][b] = 0.0;
  const double *pk = pack + (int64_t)kb * P.S * P.W * 16;
  // addresses advance by one node per step: element `ib` of a line sits at base + dx * ib
  const int dx = BWD ? -1 : 1;
  const double *xb0 = out + ridx(0, jb > 0 ? jb - 1 : 0, kb > 0 ? kb - 1 : 0), *xb1 = out + ridx(0, line ? jb : 0, kb > 0 ? kb - 1 : 0), *xb2 = out + ridx(0, jb + 1 < P.py ? jb + 1 : 0, kb > 0 ? kb - 1 : 0);
  double *xown = out + ridx(0, line ? jb : 0, kb);
  const double *rown = rhs + ridx(0, line ? jb : 0, kb);
  const bool v0 = line && kb > 0 && jb > 0, v1 = line && kb > 0, v2 = line && kb > 0 && jb + 1 < P.py;
  double2 *rows_t = rows + (size_t)t * 7; const int slot_stride = 7 * T;   // factor ring [slot][thread][7 x 16 B]: a thread's record is contiguous (pitch 112 B: conflict-free for 128-bit accesses)
  double *xs_me = xs + t * 4; const double *xs_lo = xs + (t > 0 ? t - 1 : 0) * 4;
  static_assert(ILUP_PF == ILUP_D + 1 && ILUP_PF == 4, "ring slots are derived from the unroll index");
  for (int s0 = -2 * ILUP_PF; s0 < P.S; s0 += ILUP_PF) {
#pragma unroll
    for (int u = 0; u < ILUP_PF; ++u) {
      const int s = s0 + u, ib = s - 2 * jb;
      // values fetched ILUP_PF steps ago for this step: plane kb-1 at position ib+1 of the lines jb-1, jb, jb+1 (fetched again if they
      // had not been written yet), and the right-hand side
      {
        double n0 = ring[u][0], n1 = ring[u][1], n2 = ring[u][2];
        if (ilup_missing(n0)) n0 = ilup_refetch(xb0 + dx * (ib + 1), err);
        if (ilup_missing(n1)) n1 = ilup_refetch(xb1 + dx * (ib + 1), err);
        if (ilup_missing(n2)) n2 = ilup_refetch(xb2 + dx * (ib + 1), err);
        w[0][0] = w[0][1]; w[0][1] = w[0][2]; w[0][2] = n0;
        w[1][0] = w[1][1]; w[1][1] = w[1][2]; w[1][2] = n1;
        w[2][0] = w[2][1]; w[2][1] = w[2][2]; w[2][2] = n2;
      }
      const double rhs_now = ring[u][3];
      {   // fetch for step s + ILUP_PF
        const int ib8 = ib + ILUP_PF, p = ib8 + 1;
        const bool pv = p >= 0 && p < P.px;
        ring[u][0] = (pv && v0) ? ld_relaxed_gpu_f64(xb0 + dx * p) : 0.0;
        ring[u][1] = (pv && v1) ? ld_relaxed_gpu_f64(xb1 + dx * p) : 0.0;
        ring[u][2] = (pv && v2) ? ld_relaxed_gpu_f64(xb2 + dx * p) : 0.0;
        const bool rv = line && ib8 >= 0 && ib8 < P.px;
        ring[u][3] = rv ? (BWD ? __ldcg(rown + dx * ib8) : __ldg(rown + dx * ib8)) : 0.0;
      }
      {   // factor record of step s + ILUP_D -> shared memory (1.8 us ahead: covers the DRAM latency, no separate L2 prefetch)
        const int sr = s + ILUP_D, ir = sr - 2 * jb;
        if (line && ir >= 0 && ir < P.px) {
          const double *rec = pk + (sr * P.W + (jb - ilup_jlo(sr, P.px))) * 16;
          double2 *dst = rows_t + ((u + ILUP_D) & ILUP_D) * slot_stride;   // slot of step s + ILUP_D: s0 is a multiple of the ring size, so it is known at compile time
#pragma unroll
          for (int cc = 0; cc < 7; ++cc) cp_async16(dst + cc, rec + 2 * cc);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
      }
      if (s >= 0 && s < P.S) {
        asm volatile("cp.async.wait_group %0;" :: "n"(ILUP_D) : "memory");
        if (line && ib >= 0 && ib < P.px) {
          const double2 *src = rows_t + (u & ILUP_D) * slot_stride;
          double L[14];
#pragma unroll
          for (int cc = 0; cc < 7; ++cc) { const double2 v = src[cc]; L[2 * cc] = v.x; L[2 * cc + 1] = v.y; }
          const bool jm = jb > 0;
          // line jb-1 is two nodes ahead: its positions ib-1, ib, ib+1 were its steps s-3, s-2, s-1 (ring slot = step mod 4, static)
          const double q0 = (jm && ib > 0) ? xs_lo[(u + 1) & 3] : 0.0;
          const double q1 = jm ? xs_lo[(u + 2) & 3] : 0.0;
          const double q2 = (jm && ib + 1 < P.px) ? xs_lo[(u + 3) & 3] : 0.0;
          double acc = rhs_now;
          acc -= L[0] * w[0][0]; acc -= L[1] * w[0][1]; acc -= L[2] * w[0][2];
          acc -= L[3] * w[1][0]; acc -= L[4] * w[1][1]; acc -= L[5] * w[1][2];
          acc -= L[6] * w[2][0]; acc -= L[7] * w[2][1]; acc -= L[8] * w[2][2];
          acc -= L[9] * q0; acc -= L[10] * q1; acc -= L[11] * q2;
          acc -= L[12] * xprev;
          if (BWD) acc *= L[13];
          st_relaxed_gpu_f64(xown + dx * ib, acc);
          xs_me[u & 3] = acc; xprev = acc;
        }
      }
      __syncthreads();
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
}

__global__ void __launch_bounds__(256) k_ilup_solve(IluPipe P, const double *__restrict__ packf, const double *__restrict__ packb, const double *__restrict__ b, double *y, double *x, unsigned *err)
{
  extern __shared__ double2 ilup_smem[];
  const int T = (int)blockDim.x;
  double2 *rows = ilup_smem; double *xs = (double *)(ilup_smem + (size_t)(ILUP_D + 1) * 7 * T);
  ilup_sweep<false>(P, (int)blockIdx.x, packf, b, y, err, rows, xs);
  ilup_sweep<true>(P, P.pz - 1 - (int)blockIdx.x, packb, y, x, err, rows, xs);
}

// bjacobi block of this rank: rows and columns of the owned pressure planes [op0,op1) of the local lattice, as a
// 27-point matrix on the owned sub-lattice (PETSc PCBJACOBI: the rank's diagonal block of Mpscaled).
__global__ void k_own_len(PLat P, int *len)
{
  int row = blockIdx.x * blockDim.x + threadIdx.x; if (row >= P.px * P.py * P.pz) return;
  const BoxPattern pat{P.px, P.py, P.pz, 0};
  len[row] = box_size(pat, row % P.px, (row / P.px) % P.py, row / (P.px * P.py));
}
__global__ void k_own_fill(PLat P, int op0, int pz_loc, const int *__restrict__ sia, const double *__restrict__ sa, const int *__restrict__ dia, int *dja, double *da)
{
  int row = blockIdx.x * blockDim.x + threadIdx.x; if (row >= P.px * P.py * P.pz) return;
  const int i = row % P.px, j = (row / P.px) % P.py, k = row / (P.px * P.py), ks = k + op0;
  int l0, h0, l1, h1, l2, h2, s2, t2;
  range_pp(i, P.px, l0, h0); range_pp(j, P.py, l1, h1); range_pp(k, P.pz, l2, h2); range_pp(ks, pz_loc, s2, t2);
  const int nx = h0 - l0 + 1, ny = h1 - l1 + 1, srow = i + j * P.px + ks * P.px * P.py;
  int d = dia[row];
  for (int kk = l2; kk <= h2; ++kk) for (int jj = l1; jj <= h1; ++jj) for (int ii = l0; ii <= h0; ++ii) {
    dja[d] = ii + jj * P.px + kk * P.px * P.py;
    da[d] = sa[sia[srow] + ((kk + op0 - s2) * ny + (jj - l1)) * nx + (ii - l0)];
    ++d;
  }
}
static int build_owned_block(xsb_ctx c, Csr &B)
{
  const Lattice &L = c->lat; const Slab &S = c->slab; cudaStream_t st = c->stream;
  PLat P{L.PX, L.PY, S.op1 - S.op0}; const int np = P.px * P.py * P.pz;
  int *len = nullptr; XSB_CHK(dev_alloc(c, &len, (size_t)np + 1)); XSB_CHK(dev_alloc(c, &B.ia, (size_t)np + 1));
  k_own_len<<<(np + 255) / 256, 256, 0, st>>>(P, len); KERNEL_OK();
  void *tmp = nullptr; size_t tb = 0;
  CUDA_OK(cub::DeviceScan::ExclusiveSum(nullptr, tb, len, B.ia, np + 1, st));
  CUDA_OK(cudaMalloc(&tmp, tb));
  CUDA_OK(cub::DeviceScan::ExclusiveSum(tmp, tb, len, B.ia, np + 1, st));
  int tot = 0; CUDA_OK(cudaMemcpyAsync(&tot, B.ia + np, sizeof(int), cudaMemcpyDeviceToHost, st));
  CUDA_OK(cudaStreamSynchronize(st)); CUDA_OK(cudaFree(tmp));
  B.n = B.m = np; B.nnz = tot;
  XSB_CHK(dev_alloc(c, &B.ja, (size_t)tot)); XSB_CHK(dev_alloc(c, &B.a, (size_t)tot));
  k_own_fill<<<(np + 255) / 256, 256, 0, st>>>(P, S.op0, L.PZ, c->Mp.ia, c->Mp.a, B.ia, B.ja, B.a); KERNEL_OK();
  return 0;
}

int ilu_setup(xsb_ctx c)
{
  const Lattice &L = c->lat; cudaStream_t st = c->stream;
  if (c->slab.nranks > 1) XSB_CHK(build_owned_block(c, c->MpOwn)); else c->MpOwn = c->Mp;
  if (c->slab.nranks == 1 && c->mp_block_a) c->MpOwn.a = c->mp_block_a;   // plain -fs tree emulating R ranks: cross-rank entries zeroed (xsb_fs.cu)
  const Csr &M = c->MpOwn;
  PLat P{L.PX, L.PY, c->slab.op1 - c->slab.op0};
  const int np = P.px * P.py * P.pz, nw = (P.px - 1) + 2 * (P.py - 1) + 4 * (P.pz - 1) + 1;
  XSB_CHK(dev_alloc(c, &c->mp_lu, (size_t)M.nnz));
  int *flag = nullptr; XSB_CHK(dev_alloc(c, &flag, 1));
  CUDA_OK(cudaMemsetAsync(flag, 0, sizeof(int), st));
  k_ilu0_factor<<<1, 1024, 0, st>>>(P, M.ia, M.a, c->mp_lu, flag); KERNEL_OK();
  // level schedule
  int *cnt = nullptr, *cursor = nullptr, *diag = nullptr;
  XSB_CHK(dev_alloc(c, &cnt, (size_t)nw + 1)); XSB_CHK(dev_alloc(c, &cursor, (size_t)nw + 1));
  XSB_CHK(dev_alloc(c, &c->ilu_lvl_off, (size_t)nw + 1)); XSB_CHK(dev_alloc(c, &c->ilu_rows, (size_t)np)); XSB_CHK(dev_alloc(c, &diag, (size_t)np));
  CUDA_OK(cudaMemsetAsync(cnt, 0, sizeof(int) * (nw + 1), st)); CUDA_OK(cudaMemsetAsync(cursor, 0, sizeof(int) * (nw + 1), st));
  k_lvl_count<<<(np + 255) / 256, 256, 0, st>>>(P, cnt); KERNEL_OK();
  { void *tmp = nullptr; size_t tb = 0;
    CUDA_OK(cub::DeviceScan::ExclusiveSum(nullptr, tb, cnt, c->ilu_lvl_off, nw + 1, st));
    CUDA_OK(cudaMalloc(&tmp, tb));
    CUDA_OK(cub::DeviceScan::ExclusiveSum(tmp, tb, cnt, c->ilu_lvl_off, nw + 1, st));
    CUDA_OK(cudaStreamSynchronize(st)); CUDA_OK(cudaFree(tmp)); }
  k_lvl_fill<<<(np + 255) / 256, 256, 0, st>>>(P, c->ilu_lvl_off, cursor, c->ilu_rows, diag, M.ia); KERNEL_OK();
  c->ilu_nlvl = nw; c->ilu_diag = diag;
  {
    int *lvl_of_q = nullptr; XSB_CHK(dev_alloc(c, &lvl_of_q, (size_t)np));
    XSB_CHK(dev_alloc(c, &c->ilu_fcol, (size_t)13 * np + 64)); XSB_CHK(dev_alloc(c, &c->ilu_fval, (size_t)13 * np + 64)); XSB_CHK(dev_alloc(c, &c->ilu_fn, (size_t)np));
    XSB_CHK(dev_alloc(c, &c->ilu_bcol, (size_t)13 * np + 64)); XSB_CHK(dev_alloc(c, &c->ilu_bval, (size_t)13 * np + 64)); XSB_CHK(dev_alloc(c, &c->ilu_bn, (size_t)np));
    XSB_CHK(dev_alloc(c, &c->ilu_binv, (size_t)np));
    k_lvl_of_q<<<nw, 256, 0, st>>>(nw, c->ilu_lvl_off, lvl_of_q); KERNEL_OK();
    k_ilu_pack<<<(np + 255) / 256, 256, 0, st>>>(nw, c->ilu_lvl_off, c->ilu_rows, M.ia, M.ja, diag, c->mp_lu, lvl_of_q,
                                                   c->ilu_fcol, c->ilu_fval, c->ilu_fn, c->ilu_bcol, c->ilu_bval, c->ilu_bn, c->ilu_binv); KERNEL_OK();
  }
  int h = 0; CUDA_OK(cudaMemcpyAsync(&h, flag, sizeof(int), cudaMemcpyDeviceToHost, st)); CUDA_OK(cudaStreamSynchronize(st));
  if (h) return xsb_fail(c, XSB_ERR_BREAKDOWN, "zero pivot in ILU(0) of Mpscaled");
  // line-pipelined solve (-xsb_ilu_kernel 1, default): needs one resident CTA per plane (they wait for each other)
  c->ilup_on = false;
  if (c->opt.integer("xsb_ilu_kernel", 1) == 1 && P.py <= 256) {   // one thread per node line, at most 256 (register budget of the prefetch rings)
    const int T = ((P.py + 31) / 32) * 32; const size_t smem = (size_t)T * ((ILUP_D + 1) * 7 * 16 + 4 * 8);   // one thread per node line
    int per_sm = 0, sms = 0;
    CUDA_OK(cudaFuncSetAttribute(k_ilup_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_ilup_solve, T, smem));
    CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
    if (smem <= 200 * 1024 && (int64_t)per_sm * sms >= P.pz) {
      const int S = P.px + 2 * (P.py - 1); int W = (P.px - 1) / 2 + 2; if (W > P.py) W = P.py;
      const size_t recs = (size_t)P.pz * S * W * 16;
      XSB_CHK(dev_alloc(c, &c->ilup_packf, recs)); XSB_CHK(dev_alloc(c, &c->ilup_packb, recs)); XSB_CHK(dev_alloc(c, &c->ilup_prog, (size_t)4)); XSB_CHK(dev_alloc(c, &c->ilup_y, (size_t)np));
      CUDA_OK(cudaMemsetAsync(c->ilup_prog, 0, sizeof(unsigned) * 4, st));
      c->ilup_dims[0] = P.px; c->ilup_dims[1] = P.py; c->ilup_dims[2] = P.pz; c->ilup_dims[3] = S; c->ilup_dims[4] = W; c->ilup_dims[5] = T; c->ilup_smem = smem;
      IluPipe Q{P.px, P.py, P.pz, S, W};
      k_ilup_pack<<<(np + 127) / 128, 128, 0, st>>>(Q, M.ia, diag, c->mp_lu, c->ilup_packf, c->ilup_packb); KERNEL_OK();
      c->ilup_on = true;
    }
  }
  return 0;
}

int ilu_apply(xsb_ctx c, const double *b, double *x)
{
  if (c->ilup_on) {
    IluPipe Q{c->ilup_dims[0], c->ilup_dims[1], c->ilup_dims[2], c->ilup_dims[3], c->ilup_dims[4]};
    if (b == x) return xsb_fail(c, XSB_ERR_ARG, "ILU(0) solve: right-hand side and solution must be different vectors");
    const int64_t np = (int64_t)Q.px * Q.py * Q.pz;
    k_ilup_fill<<<(unsigned)((np + 1023) / 1024 > 592 ? 592 : (np + 1023) / 1024), 256, 0, c->stream>>>(np, c->ilup_y, x); KERNEL_OK();
    k_ilup_solve<<<Q.pz, c->ilup_dims[5], c->ilup_smem, c->stream>>>(Q, c->ilup_packf, c->ilup_packb, b, c->ilup_y, x, c->ilup_prog); KERNEL_OK();
    return 0;
  }
  k_ilu0_solve<<<ILU_CLUSTER, ILU_TPB, sizeof(int) * (c->ilu_nlvl + 1), c->stream>>>(c->ilu_nlvl, c->MpOwn.n, c->ilu_lvl_off, c->ilu_rows, c->ilu_fcol, c->ilu_fval, c->ilu_fn,
                                                       c->ilu_bcol, c->ilu_bval, c->ilu_bn, c->ilu_binv, b, x); KERNEL_OK();
  return 0;
}
