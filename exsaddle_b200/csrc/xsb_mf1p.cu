// xsb_mf1p.cu -- one-pass, shared-memory / TMA-staged matrix-free Q2 apply of the velocity block (K4 of SURVEY 2.1).
//
// y = epilogue(A00 x) with ONE pass over x, y, the viscosity field and the epilogue operands (femixedspace.c:2491-2561
// applied element by element, never assembled).  Replaces round 1's memset + 8 colour launches + epilogue (3.6 x the
// algorithmic DRAM traffic: every colour pass re-read x and read-modify-wrote the accumulator; now 1.2 x).
//
//   * The element lattice is cut into columns of TI x TJ elements; the (column, element layer) pairs are linearised
//     (layers fastest) and split evenly over one persistent CTA per SM.  A CTA marches through its layers bottom-up.
//   * Node planes of x, the viscosity of the layer and the epilogue operands (b, 1/diag, p_{k-1}) are brought into shared
//     memory by bulk asynchronous copies (cp.async.bulk global -> shared, SASS UBLKCP, completion on an mbarrier with
//     expect_tx), one copy per node row of the tile, a layer ahead of their use; the three x planes of a layer live in a
//     ring of five.  Copies are issued by single lanes with uniform operands, spread over the warps.
//   * Arithmetic: 3 lanes per element (lane a owns the x-index a), sum-factorised forward / transposed contractions in
//     registers, the x-contraction through warp shuffles; zero / unit entries of the 1-D tables are skipped (exact).
//   * Each element writes its 81 outputs to an element-local slot of shared memory (no atomics, no colouring).  After a
//     block barrier the NODE PHASE sums, for every node of the two planes that are complete in z, the contributions of
//     the <= 4 (+ the carried top plane of the layer below) elements around it in a fixed order, applies the Dirichlet
//     identity rows and the fused smoother epilogue (spmv epilogue modes) and stores y.
//   * Nodes shared with another CTA (tile faces in x / y, segment ends in z) are written as partial sums to
//     part[slot][dof] (slot = parity of the tile in x, y and the z side); a small second kernel adds the <= 8 partial
//     sums of such a node in a fixed order and applies the same epilogue.  ~14 % of the nodes at 64^3.
// Summation order is fixed by the geometry, not by timing: the result is bit-reproducible run to run.
//
// Measured and dropped in round 2 (profiles/r02_summary.md): two / three / four smaller CTAs per SM (instruction-cache misses
// and spills cost more than the occupancy gains), operands read from global memory in the node phase (long-scoreboard stalls),
// a deferred node phase with one barrier per layer (all warps still hit it together), Dirichlet masking applied to the staged
// planes instead of in registers.
#include "xsb.h"

namespace {
// Tile configuration: TI x TJ elements per tile layer (a multiple of 10: one warp per 10 elements), CPS CTAs resident per SM.
template <int TI_, int TJ_, int CPS_, bool STAGE_> struct Cfg {
  static constexpr int TI = TI_, TJ = TJ_, CPS = CPS_;
  static constexpr bool STAGE = STAGE_;                // epilogue operands (b, 1/diag, p_{k-1}) staged in shared memory by bulk copies, or read from global in the node phase
  static constexpr int NEL = TI * TJ;                  // elements per tile layer
  static constexpr int NTHR = (NEL / 10) * 32;
  static constexpr int BX = 2 * TI + 1, BY = 2 * TJ + 1;  // nodes of a tile plane
  static constexpr int ROWP = 3 * BX + 1;              // doubles per staged node row: 3 BX values + 1 (16-byte alignment of the copy)
  static constexpr int ROW_BYTES = ROWP * 8;
  static constexpr int SLOT = BY * ROWP;               // doubles per staged node plane
  static constexpr int NXS = 5;                        // ring of x planes: 3 in use + 2 in flight
  static constexpr int YS = 83;                        // element stride of the element-local output buffer (81 used; 83 = 3 mod 16: conflict-free)
  static constexpr int MS_BYTES = ((NXS * BY * BX + 15) / 16) * 16;
  static constexpr int MPT = (2 * BY * BX + NTHR - 1) / NTHR;   // Dirichlet bytes of the two prefetched planes per thread
  static constexpr int NES = STAGE ? 6 : 0;            // staged operand planes: 3 vectors x the 2 planes a layer completes
  static constexpr int EROWP = TI * 27 + 2;            // doubles per staged viscosity row (TI elements x 27 Gauss points + alignment shift)
  static constexpr int ESLOT = TJ * EROWP;             // one element layer of the tile
  static constexpr size_t SMEM_BYTES = sizeof(double) * ((size_t)NXS * SLOT + NES * SLOT + 2 * ESLOT + (size_t)NEL * YS + BY * BX * 3) + MS_BYTES + 64;
  static_assert(NEL % 10 == 0 && ROW_BYTES % 16 == 0 && (SLOT * 8) % 16 == 0 && EROWP % 2 == 0 && TI % 2 == 0, "tile: multiple of 10 elements, even TI, 16-byte granular bulk copies");
};
typedef Cfg<16, 5, 1, true> CfgWide;    // 80 elements, 256 threads, one CTA per SM (157 KB of shared memory)
typedef Cfg<8, 5, 2, true> CfgTwo;      // 40 elements, 128 threads, two CTAs per SM (80 KB each): one CTA's node phase overlaps the other's arithmetic

struct Args {
  Lattice L;
  int zlo, zhi;              // element layers applied: [zlo, zhi)
  int ntx, nty, P;           // tile columns, persistent CTAs
  long long T;               // column-layers = ntx * nty * (zhi - zlo)
  MfTabS tab; double detJ;
  const double *eta; const unsigned char *bcnode;
  const double *x; double *y; double *part;
  Epilogue ep;
};

// ---- work split: CTA p owns the linearised (column, layer) indices [p T / P, (p+1) T / P)
__host__ __device__ inline long long part_lo(long long p, long long T, long long P) { return p * T / P; }
// does a segment start at (column-layer index idx)?  (then the node plane below it is shared between two CTAs)
__host__ __device__ inline bool seg_start(long long idx, long long T, long long P)
{
  const long long p = (idx * P + T - 1) / T;
  return p > 0 && p < P && part_lo(p, T, P) == idx;
}

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count)
{ asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{ asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
  asm volatile("{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}"
               ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// one node row of a tile: `bytes` from the 16-byte aligned address at or just below src
__device__ __forceinline__ void bulk_row(double *dst, const double *src, unsigned bytes, unsigned long long *bar)
{
  const unsigned long long g = (unsigned long long)(uintptr_t)src & ~15ull;
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(g), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
__device__ __forceinline__ constexpr int sym_idx(int c, int d) { return c == d ? c : (c + d == 1 ? 3 : (c + d == 2 ? 4 : 5)); }

template <int MODE>
__device__ __forceinline__ double epi_value(const Epilogue &ep, double ax, double b, double idiag, double pk, double pkm1)
{
  switch (MODE < 0 ? ep.mode : MODE) {
  case EPI_RESIDUAL:   return b - ax;
  case EPI_CHEB_FIRST: return pk + ep.s0 * (idiag * (b - ax));
  case EPI_CHEB:       return ep.s0 * pkm1 + ep.s1 * pk + ep.s2 * (idiag * (b - ax));
  default:             return ax;
  }
}

// elements of a tile direction that touch local node l (n elements): first element / its local node index, count
__device__ __forceinline__ void touching(int l, int n, int &e0, int &loc0, int &cnt)
{
  if (l & 1) { e0 = l >> 1; loc0 = 1; cnt = 1; return; }
  const int hi = l >> 1, lo = hi - 1;
  if (lo < 0) { e0 = hi; loc0 = 0; cnt = 1; }
  else if (hi >= n) { e0 = lo; loc0 = 2; cnt = 1; }
  else { e0 = lo; loc0 = 2; cnt = 2; }   // second element: e0 + 1 with local node 0
}

template <class C, int MODE>
__global__ void __launch_bounds__(C::NTHR, C::CPS) mf_onepass_kernel(const __grid_constant__ Args A)
{
  constexpr int TI = C::TI, TJ = C::TJ, NEL = C::NEL, NTHR = C::NTHR, NW = C::NTHR / 32, BX = C::BX, BY = C::BY, ROWP = C::ROWP, ROW_BYTES = C::ROW_BYTES,
                SLOT = C::SLOT, NXS = C::NXS, YS = C::YS, MS_BYTES = C::MS_BYTES, MPT = C::MPT, NES = C::NES, EROWP = C::EROWP, ESLOT = C::ESLOT;
  constexpr bool STAGE = C::STAGE;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double *xs = (double *)smem_raw;                 // [NXS][BY][ROWP]   x node planes
  double *es = xs + NXS * SLOT;                    // [3][2][BY][ROWP]  b, 1/diag, p_{k-1} planes of the current layer
  double *etas = es + NES * SLOT;                  // [2][TJ][EROWP]    viscosity at the Gauss points of two element layers
  double *yl = etas + 2 * ESLOT;                   // [NEL][YS]         element-local outputs
  double *carry = yl + NEL * YS;                   // [BY][BX][3]       top-plane sums of the layer below
  unsigned char *ms = (unsigned char *)(carry + BY * BX * 3);   // [NXS][BY][BX] per-node Dirichlet bits
  unsigned long long *bars = (unsigned long long *)(ms + MS_BYTES);   // x planes + viscosity (2, alternating layers), epilogue operands (1)

  const Lattice &L = A.L; const MfTabS &T = A.tab;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane / 3, a = lane - 3 * g, base = 3 * g;
  const int a1 = a == 2 ? 0 : a + 1, a2 = a == 0 ? 2 : a - 1;
  const int src1 = base + a1, src2 = base + a2;
  double Nr[3], Dr[3];
  Nr[0] = T.N[a][a]; Nr[1] = T.N[a][a1]; Nr[2] = T.N[a][a2];
  Dr[0] = T.Dx[a][a]; Dr[1] = T.Dx[a][a1]; Dr[2] = T.Dx[a][a2];
  const double wa = T.w[a] * A.detJ;
  const int64_t NXY = (int64_t)L.NX * L.NY, nu = L.nu;
  constexpr bool need_b = MODE != EPI_PLAIN, need_d = MODE == EPI_CHEB_FIRST || MODE == EPI_CHEB, need_m = MODE == EPI_CHEB;
  constexpr int nepi = (need_b ? 1 : 0) + (need_d ? 1 : 0) + (need_m ? 1 : 0);
  // parity of the 8-byte index of each vector's base address: a row copy starts at the 16-byte boundary at or below its
  // first value, so the values sit `shift` doubles into the staged row
  const int pbx = (int)(((uintptr_t)A.x >> 3) & 1), pbb = (int)(((uintptr_t)A.ep.b >> 3) & 1),
            pbd = (int)(((uintptr_t)A.ep.idiag >> 3) & 1), pbm = (int)(((uintptr_t)A.ep.pkm1 >> 3) & 1), pbe = (int)(((uintptr_t)A.eta >> 3) & 1);

  if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_init(&bars[2], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();

  const int nl = A.zhi - A.zlo;
  unsigned lc = 0;   // layers this CTA has processed: x barrier = lc & 1 with parity (lc >> 1) & 1; epilogue barrier parity lc & 1
  const long long hi_idx = part_lo(blockIdx.x + 1, A.T, A.P);
#pragma unroll 1
  for (long long lo_idx = part_lo(blockIdx.x, A.T, A.P); lo_idx < hi_idx;) {
    const int col = (int)(lo_idx / nl), l0 = (int)(lo_idx - (long long)col * nl);
    int l1 = l0 + (int)(hi_idx - lo_idx); if (l1 > nl) l1 = nl;
    lo_idx += l1 - l0;
    const int s0 = A.zlo + l0, s1 = A.zlo + l1;
    const int tx = col % A.ntx, ty = col / A.ntx;
    const int ei0 = tx * TI, ej0 = ty * TJ;
    const int nti = min(TI, L.mx - ei0), ntj = min(TJ, L.my - ej0), nelt = nti * ntj;
    const int bx = 2 * nti + 1, by = 2 * ntj + 1, i0 = 2 * ei0, j0 = 2 * ej0, nd = by * bx * 3;
    const int slot_lo = (tx & 1) | ((ty & 1) << 1);
    const unsigned eta_row_bytes = (unsigned)(((nti * 27 + 2) & ~1) * 8);   // one row of nti elements + the alignment shift, a multiple of 16 bytes

    // Bulk copies, issued by ONE lane with uniform operands (a per-lane address would make the compiler serialise the
    // warp around every copy).  A node plane of vector v: by rows of ROW_BYTES.
    auto issue_plane = [&](const double *v, int Pl, double *dst, unsigned long long *bar) {
      const double *src = v + 3 * (i0 + (int64_t)L.NX * j0 + NXY * Pl);
      for (int r = 0; r < by; ++r) bulk_row(dst + r * ROWP, src + 3 * (int64_t)L.NX * r, (unsigned)ROW_BYTES, bar);
    };
    auto issue_eta = [&](int s, unsigned long long *bar) {   // viscosity of element layer s: ntj rows of nti elements x 27 Gauss points
      double *dst = etas + (s & 1) * ESLOT;
      for (int r = 0; r < ntj; ++r) bulk_row(dst + r * EROWP, A.eta + 27 * (ei0 + (int64_t)L.mx * ((ej0 + r) + (int64_t)L.my * s)), eta_row_bytes, bar);
    };
    auto load_masks = [&](int Pl) {
      unsigned char *m = ms + (Pl % NXS) * (BY * BX);
      for (int t = tid; t < by * bx; t += NTHR) { const int lj = t / bx, li = t - lj * bx; m[lj * BX + li] = A.bcnode[(i0 + li) + (int64_t)L.NX * (j0 + lj) + NXY * Pl]; }
    };
    const unsigned xbytes = (unsigned)(by * ROW_BYTES), ebytes = (unsigned)(ntj) * eta_row_bytes;

    // ---- segment prologue: the three planes and the viscosity of the first layer
    __syncthreads();   // the previous segment no longer reads shared memory
    if (tid == 0) {
      mbar_expect_tx(&bars[lc & 1], 3 * xbytes + ebytes);
      for (int k = 0; k < 3; ++k) issue_plane(A.x, 2 * s0 + k, xs + ((2 * s0 + k) % NXS) * SLOT, &bars[lc & 1]);
      issue_eta(s0, &bars[lc & 1]);
    }
    for (int k = 0; k < 3; ++k) load_masks(2 * s0 + k);

#pragma unroll 1
    for (int s = s0; s < s1; ++s) {
      const bool more = s + 1 < s1;
      // ---- arm the barriers of everything issued below (one thread, before the block barrier that precedes the copies)
      if (tid == 0) {
        if (more) mbar_expect_tx(&bars[(lc + 1) & 1], 2 * xbytes + ebytes);
        if (STAGE && nepi) mbar_expect_tx(&bars[2], (unsigned)(nepi * 2) * xbytes);
      }
      // Dirichlet bits of the next layer's two new planes: loaded into registers now, stored after the arithmetic
      unsigned char mreg[MPT];
      if (more) {
#pragma unroll
        for (int u = 0; u < MPT; ++u) {
          const int t = tid + u * NTHR, k = t >= by * bx ? 1 : 0, r = t - k * by * bx;
          if (t < 2 * by * bx) { const int lj = r / bx, li = r - lj * bx; mreg[u] = A.bcnode[(i0 + li) + (int64_t)L.NX * (j0 + lj) + NXY * (2 * s + 3 + k)]; }
        }
      }
      mbar_wait(&bars[lc & 1], (lc >> 1) & 1);
      __syncthreads();   // barriers armed; masks of this layer's planes (stored by other threads) visible
      // ---- copies, spread over the warps (job w goes to warp w mod NW): the next layer's two new x planes and viscosity
      // (their ring slots were last read in layer s-1), this layer's epilogue operand planes (consumed by the node phase)
      if (lane == 0) {
        int job = 0;
        if (more) {
          if (job++ % NW == warp) issue_plane(A.x, 2 * s + 3, xs + ((2 * s + 3) % NXS) * SLOT, &bars[(lc + 1) & 1]);
          if (job++ % NW == warp) issue_plane(A.x, 2 * s + 4, xs + ((2 * s + 4) % NXS) * SLOT, &bars[(lc + 1) & 1]);
          if (job++ % NW == warp) issue_eta(s + 1, &bars[(lc + 1) & 1]);
        }
        if (STAGE) {
#pragma unroll
          for (int lk = 0; lk < 2; ++lk) {
            if (need_b) { if (job++ % NW == warp) issue_plane(A.ep.b, 2 * s + lk, es + (0 + lk) * SLOT, &bars[2]); }
            if (need_d) { if (job++ % NW == warp) issue_plane(A.ep.idiag, 2 * s + lk, es + (2 + lk) * SLOT, &bars[2]); }
            if (need_m) { if (job++ % NW == warp) issue_plane(A.ep.pkm1, 2 * s + lk, es + (4 + lk) * SLOT, &bars[2]); }
          }
        }
      }
      __syncwarp();

      // ================= element phase: 3 lanes per element, 10 elements per warp
      if (10 * warp < nelt) {
        const int t = 10 * warp + g;
        const bool valid = lane < 30 && t < nelt;
        const int tv = valid ? t : 0;
        const int tj = tv / nti, ti = tv - tj * nti;
        // Dirichlet bits of my 9 nodes x 3 components (bit 9c + 3k + j); x is read from shared memory where it is used
        unsigned bcmask = 0;
        const double *xp[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const int Pl = 2 * s + k, sl = Pl % NXS;
          xp[k] = xs + sl * SLOT + (2 * tj) * ROWP + 3 * (2 * ti + a);
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            const unsigned m = valid ? ms[sl * (BY * BX) + (2 * tj + j) * BX + 2 * ti + a] : 7u;
#pragma unroll
            for (int c = 0; c < 3; ++c) bcmask |= ((m >> c) & 1u) << (9 * c + 3 * k + j);
          }
        }
        double E[6][3][3];     // [sym slot][b][q]
#pragma unroll
        for (int sI = 0; sI < 6; ++sI)
#pragma unroll
          for (int b = 0; b < 3; ++b)
#pragma unroll
            for (int q = 0; q < 3; ++q) E[sI][b][q] = 0.0;
        // ---- forward: E_cd = d u_c / d x_d + d u_d / d x_c at my 9 Gauss points
#pragma unroll
        for (int c = 0; c < 3; ++c) {
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            double tN[3], tD[3];
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              // alignment shift of row 2 tj + j of plane 2 s + k: (pbx + i0 + j0 + row + plane) & 1 = pbx ^ ((j + k) & 1), i0 / j0 / 2 tj / 2 s being even
              double u = xp[k][j * ROWP + c + (((j + k) & 1) ? (pbx ^ 1) : pbx)];
              if ((bcmask >> (9 * c + 3 * k + j)) & 1u) u = 0.0;   // Dirichlet columns masked (MatZeroRowsColumns)
              const double u1 = shfl_d(u, src1), u2 = shfl_d(u, src2);
              tN[j] = Nr[0] * u + Nr[1] * u1 + Nr[2] * u2;
              tD[j] = Dr[0] * u + Dr[1] * u1 + Dr[2] * u2;
            }
#pragma unroll
            for (int b = 0; b < 3; ++b) {
              // 1-D tables at the middle Gauss point (xi = 0): N[1] = (0, 1, 0), D[1] = (-d, 0, d) -- the zero terms are skipped
              // (exact: they add 0), the unit coefficient needs no multiply
              const double gx = b == 1 ? tD[1] : T.N[b][0] * tD[0] + T.N[b][1] * tD[1] + T.N[b][2] * tD[2];                          // D in x, N in y
              const double gy = b == 1 ? T.Dy[1][0] * tN[0] + T.Dy[1][2] * tN[2] : T.Dy[b][0] * tN[0] + T.Dy[b][1] * tN[1] + T.Dy[b][2] * tN[2];   // N in x, D in y
              const double gz = b == 1 ? tN[1] : T.N[b][0] * tN[0] + T.N[b][1] * tN[1] + T.N[b][2] * tN[2];                          // N in x, N in y (D in z below)
#pragma unroll
              for (int q = 0; q < 3; ++q) {
                if (q == 1) { if (k == 1) { E[sym_idx(c, 0)][b][q] += gx; E[sym_idx(c, 1)][b][q] += gy; } }
                else { E[sym_idx(c, 0)][b][q] += T.N[q][k] * gx; E[sym_idx(c, 1)][b][q] += T.N[q][k] * gy; }
                if (!(q == 1 && k == 1)) E[sym_idx(c, 2)][b][q] += T.Dz[q][k] * gz;
              }
            }
          }
        }
        // ---- Gauss points: sigma = eta w |J| (G + G^T); diagonal slots hold G_cc once, so double them.  Viscosity of my 9 points
        // from the staged layer (row of element row tj; the copy started at the 16-byte boundary below its first value)
        const double *etap = etas + (s & 1) * ESLOT + tj * EROWP + ti * 27 + a + ((pbe + L.mx * ((ej0 + tj) + L.my * s)) & 1);
#pragma unroll
        for (int b = 0; b < 3; ++b)
#pragma unroll
          for (int q = 0; q < 3; ++q) {
            const double f = etap[3 * b + 9 * q] * (wa * (T.w[b] * T.w[q]));
            E[0][b][q] = f * (E[0][b][q] + E[0][b][q]); E[1][b][q] = f * (E[1][b][q] + E[1][b][q]); E[2][b][q] = f * (E[2][b][q] + E[2][b][q]);
            E[3][b][q] *= f; E[4][b][q] *= f; E[5][b][q] *= f;
          }
        // ---- transpose: the element's contribution to y_c at its node (a, j, k), stored element-locally
        double *yo = yl + tv * YS + a;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            double rx[3], ry[3], rz[3];   // index b
#pragma unroll
            for (int b = 0; b < 3; ++b) {
              const double *e0 = E[sym_idx(c, 0)][b], *e1 = E[sym_idx(c, 1)][b], *e2 = E[sym_idx(c, 2)][b];
              if (k == 1) {   // N[1][1] = 1, Dz[1][1] = 0
                rx[b] = T.N[0][1] * e0[0] + e0[1] + T.N[2][1] * e0[2];
                ry[b] = T.N[0][1] * e1[0] + e1[1] + T.N[2][1] * e1[2];
                rz[b] = T.Dz[0][1] * e2[0] + T.Dz[2][1] * e2[2];
              } else {        // N[1][k] = 0
                rx[b] = T.N[0][k] * e0[0] + T.N[2][k] * e0[2];
                ry[b] = T.N[0][k] * e1[0] + T.N[2][k] * e1[2];
                rz[b] = T.Dz[0][k] * e2[0] + T.Dz[1][k] * e2[1] + T.Dz[2][k] * e2[2];
              }
            }
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              double qD, qN;
              if (j == 1) {
                qD = T.N[0][1] * rx[0] + rx[1] + T.N[2][1] * rx[2];                                              // pairs with D in x
                qN = T.Dy[0][1] * ry[0] + T.Dy[2][1] * ry[2] + T.N[0][1] * rz[0] + rz[1] + T.N[2][1] * rz[2];    // pairs with N in x
              } else {
                qD = T.N[0][j] * rx[0] + T.N[2][j] * rx[2];
                qN = T.Dy[0][j] * ry[0] + T.Dy[1][j] * ry[1] + T.Dy[2][j] * ry[2] + T.N[0][j] * rz[0] + T.N[2][j] * rz[2];
              }
              // reduce-scatter over the 3 lanes: my contribution to node i = (a+r)%3 is s_r
              const double sA = Nr[0] * qN + Dr[0] * qD, sB = Nr[1] * qN + Dr[1] * qD, sC = Nr[2] * qN + Dr[2] * qD;
              const double Y = sA + shfl_d(sB, src2) + shfl_d(sC, src1);
              if (valid) yo[c * 27 + 3 * j + 9 * k] = Y;
            }
          }
        }
      }
      if (s + 1 < s1) {
#pragma unroll
        for (int u = 0; u < MPT; ++u) {
          const int t = tid + u * NTHR, k = t >= by * bx ? 1 : 0, r = t - k * by * bx;
          if (t < 2 * by * bx) { const int lj = r / bx, li = r - lj * bx; ms[((2 * s + 3 + k) % NXS) * (BY * BX) + lj * BX + li] = mreg[u]; }
        }
      }
      __syncthreads();

      // ================= node phase: planes 2s and 2s+1 are complete in z; the sums of plane 2s+2 are carried up
      if (STAGE && nepi) mbar_wait(&bars[2], lc & 1);
      {
        const bool first = s == s0, zshared = first && s0 > A.zlo;
        const int sl0 = (2 * s) % NXS, sl1 = (2 * s + 1) % NXS;
        // one warp per node row lj, lanes along the 3 bx dofs of the row (coalesced stores, conflict-free operand reads); the
        // number of touching element rows is uniform per warp, the second element in x is a predicated load: no divergence
        // one thread per node of the tile plane (3 components inside): the touching elements are found once per node
        for (int nn = tid; nn < by * bx; nn += NTHR) {
          const int lj = nn / bx, li = nn - lj * bx;
          int ex, ax, cx, ey, ay, cy; touching(li, nti, ex, ax, cx); touching(lj, ntj, ey, ay, cy);
          const bool shared_xy = (li == 0 && tx > 0) || (li == bx - 1 && tx < A.ntx - 1) || (lj == 0 && ty > 0) || (lj == by - 1 && ty < A.nty - 1);
          const int64_t dofn = 3 * ((i0 + li) + (int64_t)L.NX * (j0 + lj) + NXY * (2 * s));
          const int col0 = lj * ROWP + 3 * li, par = i0 + j0 + lj + 2 * s;
          const unsigned m0 = ms[sl0 * (BY * BX) + lj * BX + li], m1 = ms[sl1 * (BY * BX) + lj * BX + li];
          const double *y00 = yl + (ey * nti + ex) * YS + ax + 3 * ay;   // first touching element, my node inside it
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            double v0 = y00[c * 27], v1 = y00[c * 27 + 9], v2 = y00[c * 27 + 18];
            if (cx == 2) { const double *p = y00 + YS - ax + c * 27; v0 += p[0]; v1 += p[9]; v2 += p[18]; }                        // ascending element index
            if (cy == 2) {
              const double *q = y00 + nti * YS - 3 * ay + c * 27; v0 += q[0]; v1 += q[9]; v2 += q[18];
              if (cx == 2) { const double *p = q + YS - ax; v0 += p[0]; v1 += p[9]; v2 += p[18]; }
            }
            const int d = 3 * nn + c;
            if (!first) v0 = carry[d] + v0;   // the layer below first
            carry[d] = v2;
#pragma unroll
            for (int lk = 0; lk < 2; ++lk) {
              const double v = lk ? v1 : v0;
              const int64_t dof = dofn + c + lk * 3 * NXY;
              if (lk == 0 && zshared) { A.part[(int64_t)(slot_lo | 4) * nu + dof] = v; continue; }
              if (shared_xy) { A.part[(int64_t)slot_lo * nu + dof] = v; continue; }
              const int sl = lk ? sl1 : sl0, pp = par + lk;
              const bool bc = ((lk ? m1 : m0) >> c) & 1u;
              const double xv = xs[sl * SLOT + col0 + c + ((pbx + pp) & 1)];
              double bv = 0.0, dv = 0.0, mv = 0.0;
              if (STAGE) {
                if (need_b) bv = es[(0 + lk) * SLOT + col0 + c + ((pbb + pp) & 1)];
                if (need_d) dv = es[(2 + lk) * SLOT + col0 + c + ((pbd + pp) & 1)];
                if (need_m) mv = es[(4 + lk) * SLOT + col0 + c + ((pbm + pp) & 1)];
              } else {
                if (need_b) bv = __ldg(A.ep.b + dof);
                if (need_d) dv = __ldg(A.ep.idiag + dof);
                if (need_m) mv = __ldg(A.ep.pkm1 + dof);
              }
              A.y[dof] = epi_value<MODE>(A.ep, bc ? xv : v, bv, dv, xv, mv);   // identity rows of the constrained dofs, fused smoother update
            }
          }
        }
      }
      __syncthreads();
      ++lc;
    }
    // ---- segment end: the carried top plane always goes out as a partial sum (finished by mf_shared_kernel)
    for (int d = tid; d < nd; d += NTHR) {
      const int c = d % 3, nn = d / 3, lj = nn / bx, li = nn - lj * bx;
      A.part[(int64_t)slot_lo * nu + 3 * ((i0 + li) + (int64_t)L.NX * (j0 + lj) + NXY * (2 * s1)) + c] = carry[d];
    }
  }
}

// Nodes shared by several CTAs of mf_onepass_kernel: add their partial sums in a fixed order (z-lower first, then tile
// rows, then tile columns), then the same Dirichlet rows + epilogue.  Three families of nodes:
//   [0, n1)        node columns on an interior tile face in x (i = 2 TI t), every j, every plane
//   [n1, n1 + n2)  node rows on an interior tile face in y, i not on an x face
//   z items        (column, plane) pairs whose plane is a segment end: the column's nodes that are on no x / y face
template <class C>
__global__ void __launch_bounds__(256) mf_shared_kernel(Args A, long long n1, long long n2, int nb12, const int *__restrict__ zitems)
{
  constexpr int TI = C::TI, TJ = C::TJ;
  static_assert(C::BX * C::BY <= 512, "a z item is handled by two CTAs of 256 threads");
  const Lattice &L = A.L;
  const int nl = A.zhi - A.zlo, NP = 2 * nl + 1;
  const int64_t NXY = (int64_t)L.NX * L.NY, nu = L.nu;
  int i, j, Pl;
  if ((int)blockIdx.x < nb12) {
    const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
    if (t < n1) { const int f = (int)(t % (A.ntx - 1)); const long long r = t / (A.ntx - 1); i = 2 * TI * (f + 1); j = (int)(r % L.NY); Pl = 2 * A.zlo + (int)(r / L.NY); }
    else if (t < n1 + n2) {
      const long long u = t - n1; const int f = (int)(u % (A.nty - 1)); const long long r = u / (A.nty - 1);
      j = 2 * TJ * (f + 1); i = (int)(r % L.NX); Pl = 2 * A.zlo + (int)(r / L.NX);
      if (i % (2 * TI) == 0 && i > 0 && i < L.NX - 1) return;   // x face: first family
    } else return;
    if (Pl - 2 * A.zlo >= NP) return;
  } else {
    const int item = ((int)blockIdx.x - nb12) / 2, t = (((int)blockIdx.x - nb12) & 1) * 256 + threadIdx.x;
    const int col = zitems[2 * item]; Pl = zitems[2 * item + 1];
    const int tx = col % A.ntx, ty = col / A.ntx;
    const int nti = min(TI, L.mx - tx * TI), ntj = min(TJ, L.my - ty * TJ), bx = 2 * nti + 1, by = 2 * ntj + 1;
    if (t >= bx * by) return;
    const int lj = t / bx, li = t - lj * bx;
    if ((li == 0 && tx > 0) || (li == bx - 1 && tx < A.ntx - 1) || (lj == 0 && ty > 0) || (lj == by - 1 && ty < A.nty - 1)) return;   // faces: first two families
    i = 2 * TI * tx + li; j = 2 * TJ * ty + lj;
  }
  // tile columns containing the node
  const bool xf = i % (2 * TI) == 0 && i > 0 && i < L.NX - 1, yf = j % (2 * TJ) == 0 && j > 0 && j < L.NY - 1;
  const int tx1 = min(i / (2 * TI), A.ntx - 1), ty1 = min(j / (2 * TJ), A.nty - 1);
  const int tx0 = xf ? tx1 - 1 : tx1, ty0 = yf ? ty1 - 1 : ty1;
  const int64_t node = i + (int64_t)L.NX * j + NXY * Pl;
  const unsigned bcm = A.bcnode[node];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const int64_t dof = 3 * node + c;
    double v = 0.0; bool started = false;
    for (int ty = ty0; ty <= ty1; ++ty) for (int tx = tx0; tx <= tx1; ++tx) {
      const int col = ty * A.ntx + tx, slot_lo = (tx & 1) | ((ty & 1) << 1);
      // z contributors of this column at plane Pl: an even plane strictly inside [zlo, zhi] that starts a segment has two
      int nz = 1;
      if (!(Pl & 1)) { const int s = Pl >> 1; if (s > A.zlo && s < A.zhi && seg_start((long long)col * nl + (s - A.zlo), A.T, A.P)) nz = 2; }
      for (int zb = 0; zb < nz; ++zb) {
        const double pv = A.part[(int64_t)(slot_lo | (zb << 2)) * nu + dof];
        if (!started) { v = pv; started = true; } else v += pv;
      }
    }
    const double xv = A.x[dof];
    const double bv = A.ep.mode != EPI_PLAIN ? A.ep.b[dof] : 0.0;
    const double dv = (A.ep.mode == EPI_CHEB_FIRST || A.ep.mode == EPI_CHEB) ? A.ep.idiag[dof] : 0.0;
    const double mv = A.ep.mode == EPI_CHEB ? A.ep.pkm1[dof] : 0.0;
    A.y[dof] = epi_value<-1>(A.ep, ((bcm >> c) & 1u) ? xv : v, bv, dv, xv, mv);
  }
}
}   // namespace

int mf1p_partition(int P, int64_t ncols, int nl, int p, int64_t *lo, int64_t *hi)
{
  const long long T = (long long)ncols * nl;
  if (P < 1 || p < 0 || p >= P || T < 1) return XSB_ERR_ARG;
  *lo = part_lo(p, T, P); *hi = part_lo(p + 1, T, P);
  return 0;
}

// Lazy state of the kernel (shared-memory opt-in, partial-sum buffer, z-item table): everything that allocates or synchronises,
// so that a product inside a CUDA-graph capture finds it done (mf_setup calls mf1p_prepare).
template <class C>
static int mf1p_prepare_t(xsb_ctx c, Args &A, int variant)
{
  const Lattice &L = c->lat; cudaStream_t st = c->stream;
  if (c->mf_ready != variant + 1) {
    cudaDeviceProp prop; CUDA_OK(cudaGetDeviceProperties(&prop, c->device));
    if ((size_t)prop.sharedMemPerBlockOptin < C::SMEM_BYTES) return xsb_fail(c, XSB_ERR_SUP, "one-pass element kernel needs %zu bytes of shared memory per CTA (device offers %zu)", C::SMEM_BYTES, (size_t)prop.sharedMemPerBlockOptin);
    CUDA_OK(cudaFuncSetAttribute(mf_onepass_kernel<C, EPI_PLAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES));
    CUDA_OK(cudaFuncSetAttribute(mf_onepass_kernel<C, EPI_RESIDUAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES));
    CUDA_OK(cudaFuncSetAttribute(mf_onepass_kernel<C, EPI_CHEB_FIRST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES));
    CUDA_OK(cudaFuncSetAttribute(mf_onepass_kernel<C, EPI_CHEB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES));
    c->mf_sms = prop.multiProcessorCount;
    if (!c->mf_part) { const int kp = c->phase; c->phase = 2; int rc = dev_alloc(c, &c->mf_part, (size_t)8 * L.nu); c->phase = kp; if (rc) return rc; }
    c->mf_ready = variant + 1; c->mf_zkey[0] = -1;
  }
  A.part = c->mf_part;
  A.ntx = (L.mx + C::TI - 1) / C::TI; A.nty = (L.my + C::TJ - 1) / C::TJ;
  const int nl = A.zhi - A.zlo, ncols = A.ntx * A.nty, slots = c->mf_sms * C::CPS;
  A.T = (long long)ncols * nl;
  A.P = (int)(A.T < slots ? A.T : slots);
  // (column, plane) pairs whose node plane is a segment end: every segment start above zlo, and the top plane of every column
  if (c->mf_zkey[0] != A.zlo || c->mf_zkey[1] != A.zhi || c->mf_zkey[2] != A.P) {
    std::vector<int> items;
    for (int p = 1; p < A.P; ++p) { const long long idx = part_lo(p, A.T, A.P); if (idx % nl) { items.push_back((int)(idx / nl)); items.push_back(2 * (A.zlo + (int)(idx % nl))); } }
    for (int col = 0; col < ncols; ++col) { items.push_back(col); items.push_back(2 * A.zhi); }
    const size_t cap = (size_t)2 * ((size_t)L.mx * L.my + 4 * (size_t)c->mf_sms + 2);   // any tile variant: columns <= elements per layer
    if (!c->mf_zitems) { const int kp = c->phase; c->phase = 2; int rc = dev_alloc(c, &c->mf_zitems, cap); c->phase = kp; if (rc) return rc; }
    if (items.size() > cap) return xsb_fail(c, XSB_ERR_MEM, "one-pass element kernel: z-item table overflow");
    CUDA_OK(cudaMemcpyAsync(c->mf_zitems, items.data(), sizeof(int) * items.size(), cudaMemcpyHostToDevice, st));
    CUDA_OK(cudaStreamSynchronize(st));   // `items` is pageable, function-local memory
    c->mf_nz = (int)items.size() / 2; c->mf_zkey[0] = A.zlo; c->mf_zkey[1] = A.zhi; c->mf_zkey[2] = A.P;
  }
  return 0;
}

template <class C>
static int mf1p_launch(xsb_ctx c, Args &A, int variant)
{
  const Lattice &L = c->lat; cudaStream_t st = c->stream;
  XSB_CHK(mf1p_prepare_t<C>(c, A, variant));
  const int nl = A.zhi - A.zlo;
  switch (A.ep.mode) {
  case EPI_RESIDUAL:   mf_onepass_kernel<C, EPI_RESIDUAL><<<A.P, C::NTHR, C::SMEM_BYTES, st>>>(A); break;
  case EPI_CHEB_FIRST: mf_onepass_kernel<C, EPI_CHEB_FIRST><<<A.P, C::NTHR, C::SMEM_BYTES, st>>>(A); break;
  case EPI_CHEB:       mf_onepass_kernel<C, EPI_CHEB><<<A.P, C::NTHR, C::SMEM_BYTES, st>>>(A); break;
  default:             mf_onepass_kernel<C, EPI_PLAIN><<<A.P, C::NTHR, C::SMEM_BYTES, st>>>(A);
  }
  KERNEL_OK();
  const int NP = 2 * nl + 1;
  const long long n1 = (long long)(A.ntx - 1) * L.NY * NP, n2 = (long long)(A.nty - 1) * L.NX * NP;
  const int nb12 = (int)((n1 + n2 + 255) / 256);
  mf_shared_kernel<C><<<nb12 + 2 * c->mf_nz, 256, 0, st>>>(A, n1, n2, nb12, c->mf_zitems); KERNEL_OK();
  return 0;
}

// y = epilogue(K x) with the per-node Dirichlet bits `bcnode` (all zero: the operator before MatZeroRowsColumns), element
// layers [zlo, zhi) of the local lattice.  x, y must not alias; with a Chebyshev epilogue ep.pk must be x.
int mf1p_prepare(xsb_ctx c, int zlo, int zhi)
{
  if (zhi <= zlo) return 0;
  Args A; A.L = c->lat; A.zlo = zlo; A.zhi = zhi;
  switch (c->so.mf_tile) {
  case 1: return mf1p_prepare_t<CfgTwo>(c, A, 1);
  default: return mf1p_prepare_t<CfgWide>(c, A, 0);
  }
}

int mf1p_apply(xsb_ctx c, const double *x, double *y, const Epilogue &ep, const unsigned char *bcnode, int zlo, int zhi)
{
  const Lattice &L = c->lat;
  if (zhi <= zlo) return 0;
  if ((ep.mode == EPI_CHEB || ep.mode == EPI_CHEB_FIRST) && ep.pk != x) return xsb_fail(c, XSB_ERR_ARG, "one-pass element kernel: the Chebyshev epilogue's p_k must be the product's input vector");
  Args A; A.L = L; A.zlo = zlo; A.zhi = zhi;
  mf_tab_scaled(L, A.tab); A.detJ = L.hu[0] * L.hu[1] * L.hu[2];
  A.eta = c->coeff; A.bcnode = bcnode; A.x = x; A.y = y; A.part = nullptr; A.ep = ep;
  const int variant = c->so.mf_tile;
  int rc;
  switch (variant) {
  case 1: rc = mf1p_launch<CfgTwo>(c, A, 1); break;
  default: rc = mf1p_launch<CfgWide>(c, A, 0);
  }
  return rc;
}
