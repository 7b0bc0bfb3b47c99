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
  return 0;
}

int comm_destroy(xsb_ctx c)
{
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

// Ghost update of a lattice vector whose planes hold `pd` doubles: this rank's owned planes are local [o0,o1);
// it needs gb planes below o0 (the top gb owned planes of rank-1) and ga planes at o1 (the bottom ga of rank+1).
static int halo_planes(xsb_ctx c, double *v, int64_t pd, int o0, int o1, int gb, int ga)
{
  const Slab &S = c->slab;
  if (S.nranks == 1) return 0;
  ncclComm_t comm = (ncclComm_t)c->nccl;
  NCCL_OK(g_nccl.GroupStart());
  if (S.rank > 0) {
    NCCL_OK(g_nccl.Recv(v + (int64_t)(o0 - gb) * pd, (size_t)(gb * pd), ncclFloat64_, S.rank - 1, comm, c->stream));
    NCCL_OK(g_nccl.Send(v + (int64_t)o0 * pd, (size_t)(ga * pd), ncclFloat64_, S.rank - 1, comm, c->stream));
  }
  if (S.rank < S.nranks - 1) {
    NCCL_OK(g_nccl.Recv(v + (int64_t)o1 * pd, (size_t)(ga * pd), ncclFloat64_, S.rank + 1, comm, c->stream));
    NCCL_OK(g_nccl.Send(v + (int64_t)(o1 - gb) * pd, (size_t)(gb * pd), ncclFloat64_, S.rank + 1, comm, c->stream));
  }
  NCCL_OK(g_nccl.GroupEnd());
  return 0;
}
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

// Row-partitioned product on a replicated level: rank r holds fresh values for planes [r cp, (r+1) cp) of `glob`, cp = ceil(n / N)
// (equal chunks: the vector is allocated with N cp planes, the tail beyond the lattice is padding), so the exchange is ONE
// in-place ncclAllGather instead of N grouped broadcasts.
int comm_allgather_planes(xsb_ctx c, double *glob, int64_t pd, int nplanes)
{
  const Slab &S = c->slab;
  if (S.nranks == 1) return 0;
  const int64_t cp = (nplanes + S.nranks - 1) / S.nranks;
  NCCL_OK(g_nccl.AllGather(glob + (int64_t)S.rank * cp * pd, glob, (size_t)(cp * pd), ncclFloat64_, (ncclComm_t)c->nccl, c->stream));
  return 0;
}
