// xsb_comm.cu -- NCCL plumbing for the z-slab partition (one process per GPU).
//
// Replaces what PETSc/MPI do implicitly on the solve path: the VecScatter ghost update in front of
// every MatMult on MPIAIJ (femixedspace.c:1150 stencil width 2 for u, :1243 width 1 for p) and the MPI_Allreduce
// behind VecMDot / VecNorm.  Slabs are cut along z, so a ghost plane is one contiguous range of the vector:
// halo exchange is ncclSend / ncclRecv straight from / into the vectors (no pack kernels), both neighbours in one
// group, on the handle's stream (ordered with the kernels that produce and consume the planes).
// NCCL is loaded with dlopen at xsb_comm_init, so the single-GPU library has no link-time NCCL dependency.
#include "xsb.h"
#include <dlfcn.h>

typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess_ = 0 };
enum { ncclFloat64_ = 8, ncclSum_ = 0 };

struct NcclApi {
  void *h = nullptr;
  int (*GetUniqueId)(ncclUniqueId *) = nullptr;
  int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Broadcast)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;

static int nccl_load(xsb_ctx c)
{
  if (g_nccl.h) return 0;
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char *n : names) { g_nccl.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (g_nccl.h) break; }
  if (!g_nccl.h) return xsb_fail(c, XSB_ERR_NCCL, "cannot load libnccl.so.2: %s", dlerror());
#define SYM(field, name) *(void **)(&g_nccl.field) = dlsym(g_nccl.h, name); if (!g_nccl.field) return xsb_fail(c, XSB_ERR_NCCL, "libnccl lacks %s", name)
  SYM(GetUniqueId, "ncclGetUniqueId"); SYM(CommInitRank, "ncclCommInitRank"); SYM(CommDestroy, "ncclCommDestroy");
  SYM(AllReduce, "ncclAllReduce"); SYM(Broadcast, "ncclBroadcast"); SYM(AllGather, "ncclAllGather"); SYM(Send, "ncclSend"); SYM(Recv, "ncclRecv");
  SYM(GroupStart, "ncclGroupStart"); SYM(GroupEnd, "ncclGroupEnd"); SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
  return 0;
}
#define NCCL_OK(call) do { int r_ = (call); if (r_ != ncclSuccess_) return xsb_fail(c, XSB_ERR_NCCL, "%s:%d %s: %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString(r_)); } while (0)

int comm_unique_id(void *out128)
{
  xsb_ctx c = nullptr;
  if (nccl_load(c)) return XSB_ERR_NCCL;
  ncclUniqueId id; if (g_nccl.GetUniqueId(&id) != ncclSuccess_) return XSB_ERR_NCCL;
  memcpy(out128, &id, sizeof(id));
  return 0;
}

int comm_init(xsb_ctx c, const void *unique_id, int rank, int nranks)
{
  if (nranks < 1 || rank < 0 || rank >= nranks) return xsb_fail(c, XSB_ERR_ARG, "bad rank %d of %d", rank, nranks);
  c->slab.rank = rank; c->slab.nranks = nranks;
  if (nranks == 1) return 0;
  if (c->nsd != 3) return xsb_fail(c, XSB_ERR_SUP, "the z-slab partition is implemented for the 3-D executables");
  XSB_CHK(nccl_load(c));
  ncclUniqueId id; memcpy(&id, unique_id, sizeof(id));
  ncclComm_t comm = nullptr;
  NCCL_OK(g_nccl.CommInitRank(&comm, nranks, id, rank));
  c->nccl = comm;
  if (!c->side) { CUDA_OK(cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking)); CUDA_OK(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming)); CUDA_OK(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming)); }
  return 0;
}

int comm_destroy(xsb_ctx c)
{
  comm_p2p_destroy(c);
  if (c->nccl && g_nccl.CommDestroy) g_nccl.CommDestroy((ncclComm_t)c->nccl);
  c->nccl = nullptr;
  return 0;
}

int comm_allreduce_sum(xsb_ctx c, double *dev, int n)
{
  if (c->slab.nranks == 1) return 0;
  NCCL_OK(g_nccl.AllReduce(dev, dev, (size_t)n, ncclFloat64_, ncclSum_, (ncclComm_t)c->nccl, c->stream));
  return 0;
}

// ------------------------------------------------------------------ peer-memory halo exchange (NVLink / NVSwitch)
// The solve exchanges ghost planes in front of every fine-level product and behind every product of a distributed coarse
// level (~3300 exchanges per 128^3 solve), 0.4 - 3 MB each: latency, not bandwidth.  ncclSend/ncclRecv cost 15 - 22 us per
// exchange here; so did a first peer-memory kernel that published a sequence number behind system-scope fences.  The
// exchange is now ONE kernel without fences or flags -- THE VALUE IS THE FLAG (the idea of NCCL's LL protocol, for doubles):
//   * every rank owns a window (cudaMalloc, opened by both neighbours through cudaIpc handles exchanged once with
//     ncclAllGather), two slots per direction, kept filled with a sentinel (a quiet NaN whose payload no computation produces);
//   A. the kernel stores this rank's boundary planes straight into the neighbours' windows (plain 8-byte stores over NVLink);
//   B. each thread then reads its share of the rank's OWN window, re-reading an element until it is no longer the sentinel
//      (an 8-byte store is single-copy atomic, and nothing but the value itself is awaited), copies it into the ghost
//      plane and puts the sentinel back.
// Slots alternate with the parity of an exchange counter in device memory (advanced by the last block, so the kernel replays
// inside CUDA graphs).  Slot s%2 is free again when a rank starts exchange s+2: it has finished part B of s+1, so the
// neighbour finished part A of s+1, which is stream-ordered behind the neighbour's whole kernel s (its reads and its
// sentinel stores).  A wait beyond ~2 s (a neighbour died) raises a sticky error word instead of hanging the GPU.
struct P2P {
  bool on = false;
  char *win = nullptr, *peer_lo = nullptr, *peer_hi = nullptr;   // my window, the windows of rank-1 / rank+1
  unsigned long long *ctl = nullptr;                               // local control words: [0] exchange counter, [1] blocks done, [2] error
  size_t slot_bytes = 0;
};
enum { P2P_HDR = 256 };   // window: [pad][from-below slot 0,1][from-above slot 0,1]
#define P2P_SENTINEL 0x7FF8C0FFEE5EED01ULL
struct P2PDev { unsigned long long *ctl; char *win, *peer_lo, *peer_hi; size_t slot_bytes; };

__device__ __forceinline__ unsigned long long ld_relaxed_sys_u64(const unsigned long long *p)
{ unsigned long long v; asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v; }
__global__ void k_p2p_fill(unsigned long long *w, size_t n) { for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) w[i] = P2P_SENTINEL; }

__global__ void __launch_bounds__(256) k_halo_p2p(double *__restrict__ v, int64_t pd, int o0, int o1, int gb, int ga, P2PDev w)
{
  __shared__ unsigned long long s_seq;
  const bool lo = w.peer_lo != nullptr, hi = w.peer_hi != nullptr;
  if (threadIdx.x == 0) s_seq = *(volatile unsigned long long *)&w.ctl[0];
  __syncthreads();
  const int slot = (int)(s_seq & 1);
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
  // A: my bottom `ga` owned planes are the neighbour below's ghost planes above; my top `gb` owned planes the ghosts below of the neighbour above
  if (lo) { double *dst = (double *)(w.peer_lo + P2P_HDR + (size_t)(2 + slot) * w.slot_bytes); const double *src = v + (int64_t)o0 * pd; for (int64_t i = tid; i < ga * pd; i += nth) dst[i] = src[i]; }
  if (hi) { double *dst = (double *)(w.peer_hi + P2P_HDR + (size_t)(0 + slot) * w.slot_bytes); const double *src = v + (int64_t)(o1 - gb) * pd; for (int64_t i = tid; i < gb * pd; i += nth) dst[i] = src[i]; }
  // B: my ghost planes from my own window
  bool dead = *(volatile unsigned long long *)&w.ctl[2] != 0;
  for (int side = 0; side < 2; ++side) {
    if (side == 0 ? !lo : !hi) continue;
    unsigned long long *src = (unsigned long long *)(w.win + P2P_HDR + (size_t)((side ? 2 : 0) + slot) * w.slot_bytes);
    double *dst = side ? v + (int64_t)o1 * pd : v + (int64_t)(o0 - gb) * pd;
    const int64_t n = (side ? ga : gb) * pd;
    for (int64_t i = tid; i < n; i += nth) {
      unsigned long long x = ld_relaxed_sys_u64(src + i);
      if (x == P2P_SENTINEL && !dead) {
        const long long t0 = clock64();
        do { x = ld_relaxed_sys_u64(src + i); if (clock64() - t0 > 4000000000LL) { atomicExch(&w.ctl[2], 1ULL); dead = true; break; } } while (x == P2P_SENTINEL);
      }
      dst[i] = __longlong_as_double((long long)x);
      src[i] = P2P_SENTINEL;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0 && atomicAdd(&w.ctl[1], 1ULL) == gridDim.x - 1) { w.ctl[1] = 0; *(volatile unsigned long long *)&w.ctl[0] = s_seq + 1; }   // the last block closes the exchange
}

static void p2p_release(xsb_ctx c, bool collective)
{
  P2P *p = (P2P *)c->p2p; if (!p) return;
  cudaStreamSynchronize(c->stream);
  if (p->peer_lo) cudaIpcCloseMemHandle(p->peer_lo);
  if (p->peer_hi) cudaIpcCloseMemHandle(p->peer_hi);
  if (collective && c->nccl && p->ctl) { g_nccl.AllReduce(p->ctl + 4, p->ctl + 4, 1, ncclFloat64_, ncclSum_, (ncclComm_t)c->nccl, c->stream); cudaStreamSynchronize(c->stream); }   // nobody frees a window a neighbour still maps
  if (p->win) cudaFree(p->win);
  if (p->ctl) cudaFree(p->ctl);
  delete p; c->p2p = nullptr;
}

// Collective (called by xsb_assemble on every rank once the local lattice is known): windows sized for the largest message,
// two velocity planes of the fine lattice.  -xsb_p2p 0 keeps ncclSend / ncclRecv.
int comm_p2p_setup(xsb_ctx c)
{
  const Slab &S = c->slab; const Lattice &L = c->lat;
  if (S.nranks == 1) return 0;
  if (!c->opt.integer("xsb_p2p", 1)) { p2p_release(c, true); return 0; }
  const size_t need = (((size_t)2 * L.nsd * L.NX * L.NY * sizeof(double)) + 255) & ~(size_t)255;
  P2P *p = (P2P *)c->p2p;
  if (p && p->on && p->slot_bytes == need) return 0;   // NX, NY are global: every rank takes the same branch
  p2p_release(c, true);
  p = new P2P(); c->p2p = p; p->slot_bytes = need;
  CUDA_OK(cudaMalloc(&p->win, P2P_HDR + 4 * need)); CUDA_OK(cudaMemsetAsync(p->win, 0, P2P_HDR, c->stream));
  k_p2p_fill<<<256, 256, 0, c->stream>>>((unsigned long long *)(p->win + P2P_HDR), 4 * need / 8); KERNEL_OK();
  CUDA_OK(cudaMalloc(&p->ctl, 64)); CUDA_OK(cudaMemsetAsync(p->ctl, 0, 64, c->stream));
  cudaIpcMemHandle_t mine; CUDA_OK(cudaIpcGetMemHandle(&mine, p->win));
  char *hd = nullptr; CUDA_OK(cudaMalloc(&hd, sizeof(mine) * S.nranks));
  CUDA_OK(cudaMemcpyAsync(hd + sizeof(mine) * S.rank, &mine, sizeof(mine), cudaMemcpyHostToDevice, c->stream));
  NCCL_OK(g_nccl.AllGather(hd + sizeof(mine) * S.rank, hd, sizeof(mine), 0 /* ncclInt8 */, (ncclComm_t)c->nccl, c->stream));   // also orders every rank's sentinel fill before anyone's first push
  std::vector<cudaIpcMemHandle_t> all(S.nranks);
  CUDA_OK(cudaMemcpyAsync(all.data(), hd, sizeof(mine) * S.nranks, cudaMemcpyDeviceToHost, c->stream));
  CUDA_OK(cudaStreamSynchronize(c->stream)); CUDA_OK(cudaFree(hd));
  cudaError_t e = cudaSuccess;
  if (S.rank > 0) e = cudaIpcOpenMemHandle((void **)&p->peer_lo, all[S.rank - 1], cudaIpcMemLazyEnablePeerAccess);
  if (e == cudaSuccess && S.rank < S.nranks - 1) e = cudaIpcOpenMemHandle((void **)&p->peer_hi, all[S.rank + 1], cudaIpcMemLazyEnablePeerAccess);
  // all ranks must agree on the transport: one that cannot map its neighbours sends everybody back to NCCL
  double ok = e == cudaSuccess ? 0.0 : 1.0; double *okd = (double *)(p->ctl + 5);
  if (e != cudaSuccess) cudaGetLastError();
  CUDA_OK(cudaMemcpyAsync(okd, &ok, sizeof(double), cudaMemcpyHostToDevice, c->stream));
  NCCL_OK(g_nccl.AllReduce(okd, okd, 1, ncclFloat64_, ncclSum_, (ncclComm_t)c->nccl, c->stream));
  CUDA_OK(cudaMemcpyAsync(&ok, okd, sizeof(double), cudaMemcpyDeviceToHost, c->stream)); CUDA_OK(cudaStreamSynchronize(c->stream));
  CUDA_OK(cudaMemsetAsync(okd, 0, sizeof(double), c->stream));
  p->on = ok == 0.0;
  if (!p->on && c->opt.integer("xsb_p2p", 1) > 1) return xsb_fail(c, XSB_ERR_NCCL, "-xsb_p2p 2: peer-memory windows could not be mapped (%s)", cudaGetErrorString(e));
  return 0;
}
int comm_p2p_active(xsb_ctx c) { const P2P *p = (const P2P *)c->p2p; return p && p->on; }
int comm_p2p_check(xsb_ctx c)
{
  const P2P *p = (const P2P *)c->p2p; if (!p || !p->on) return 0;
  unsigned long long err = 0; CUDA_OK(cudaMemcpyAsync(&err, p->ctl + 2, sizeof(err), cudaMemcpyDeviceToHost, c->stream)); CUDA_OK(cudaStreamSynchronize(c->stream));
  if (err) return xsb_fail(c, XSB_ERR_NCCL, "peer-memory halo exchange timed out waiting for a neighbour rank");
  return 0;
}
void comm_p2p_destroy(xsb_ctx c) { p2p_release(c, false); }   // tear-down is not collective: the windows are idle, no barrier

// Ghost update of a lattice vector whose planes hold `pd` doubles: this rank's owned planes are [o0,o1) in the vector's own
// plane numbering; it needs gb planes below o0 (the top gb owned planes of rank-1) and ga planes at o1 (the bottom ga of rank+1).
int comm_halo_planes(xsb_ctx c, double *v, int64_t pd, int o0, int o1, int gb, int ga, cudaStream_t on)
{
  const Slab &S = c->slab;
  if (S.nranks == 1) return 0;
  cudaStream_t st = on ? on : c->stream;
  const P2P *p = (const P2P *)c->p2p;
  if (p && p->on && (size_t)((gb > ga ? gb : ga) * pd) * sizeof(double) <= p->slot_bytes) {
    P2PDev w{p->ctl, p->win, p->peer_lo, p->peer_hi, p->slot_bytes};
    int64_t blocks = ((gb > ga ? gb : ga) * pd + 1023) / 1024; if (blocks > 148) blocks = 148; if (blocks < 1) blocks = 1;
    k_halo_p2p<<<(unsigned)blocks, 256, 0, st>>>(v, pd, o0, o1, gb, ga, w); KERNEL_OK();
    return 0;
  }
  ncclComm_t comm = (ncclComm_t)c->nccl;
  NCCL_OK(g_nccl.GroupStart());
  if (S.rank > 0) {
    NCCL_OK(g_nccl.Recv(v + (int64_t)(o0 - gb) * pd, (size_t)(gb * pd), ncclFloat64_, S.rank - 1, comm, st));
    NCCL_OK(g_nccl.Send(v + (int64_t)o0 * pd, (size_t)(ga * pd), ncclFloat64_, S.rank - 1, comm, st));
  }
  if (S.rank < S.nranks - 1) {
    NCCL_OK(g_nccl.Recv(v + (int64_t)o1 * pd, (size_t)(ga * pd), ncclFloat64_, S.rank + 1, comm, st));
    NCCL_OK(g_nccl.Send(v + (int64_t)(o1 - gb) * pd, (size_t)(gb * pd), ncclFloat64_, S.rank + 1, comm, st));
  }
  NCCL_OK(g_nccl.GroupEnd());
  return 0;
}
static int halo_planes(xsb_ctx c, double *v, int64_t pd, int o0, int o1, int gb, int ga) { return comm_halo_planes(c, v, pd, o0, o1, gb, ga); }
int comm_halo_u(xsb_ctx c, double *u) { const Lattice &L = c->lat; return halo_planes(c, u, (int64_t)L.nsd * L.NX * L.NY, c->slab.ou0, c->slab.ou1, 2, 1); }
int comm_halo_p(xsb_ctx c, double *p) { const Lattice &L = c->lat; return halo_planes(c, p, (int64_t)L.PX * L.PY, c->slab.op0, c->slab.op1, 1, 1); }
int comm_halo_full(xsb_ctx c, double *x)
{
  if (c->slab.nranks == 1) return 0;
  XSB_CHK(comm_halo_u(c, x));
  return comm_halo_p(c, x + c->lat.nu);
}

// every rank r broadcasts glob[offs[r] .. offs[r+1]) (setup-time replication of operator rows)
int comm_bcast_segments(xsb_ctx c, double *glob, const int64_t *offs)
{
  const Slab &S = c->slab;
  if (S.nranks == 1) return 0;
  ncclComm_t comm = (ncclComm_t)c->nccl;
  NCCL_OK(g_nccl.GroupStart());
  for (int r = 0; r < S.nranks; ++r) {
    double *p = glob + offs[r];
    if (offs[r + 1] > offs[r]) NCCL_OK(g_nccl.Broadcast(p, p, (size_t)(offs[r + 1] - offs[r]), ncclFloat64_, r, comm, c->stream));
  }
  NCCL_OK(g_nccl.GroupEnd());
  return 0;
}

// Replicated coarse vector / operator rows: global plane Z of the first coarse level is owned by the rank whose
// element layers contain layer Z (the last rank also owns the top plane).  Every rank broadcasts its planes.
int comm_bcast_planes(xsb_ctx c, double *glob, int64_t pd, int nplanes_glob)
{
  const Slab &S = c->slab;
  if (S.nranks == 1) return 0;
  ncclComm_t comm = (ncclComm_t)c->nccl;
  NCCL_OK(g_nccl.GroupStart());
  for (int r = 0; r < S.nranks; ++r) {
    int k0, k1; xsb_slab_range(S.mz_glob, S.nranks, r, &k0, &k1);
    if (r == S.nranks - 1) k1 = nplanes_glob;
    double *p = glob + (int64_t)k0 * pd;
    NCCL_OK(g_nccl.Broadcast(p, p, (size_t)((int64_t)(k1 - k0) * pd), ncclFloat64_, r, comm, c->stream));
  }
  NCCL_OK(g_nccl.GroupEnd());
  return 0;
}

// A coarse vector whose planes [p0[r], p1[r]) were computed by rank r (restriction from a plane-distributed level) becomes
// replicated: every rank broadcasts its planes (one group).
int comm_bcast_plane_ranges(xsb_ctx c, double *glob, int64_t pd, const int *p0, const int *p1)
{
  const Slab &S = c->slab;
  if (S.nranks == 1) return 0;
  ncclComm_t comm = (ncclComm_t)c->nccl;
  NCCL_OK(g_nccl.GroupStart());
  for (int r = 0; r < S.nranks; ++r) {
    if (p1[r] <= p0[r]) continue;
    double *p = glob + (int64_t)p0[r] * pd;
    NCCL_OK(g_nccl.Broadcast(p, p, (size_t)((int64_t)(p1[r] - p0[r]) * pd), ncclFloat64_, r, comm, c->stream));
  }
  NCCL_OK(g_nccl.GroupEnd());
  return 0;
}
