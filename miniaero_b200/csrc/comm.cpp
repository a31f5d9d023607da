// Halo communicator: NCCL send/recv over NVLink, one process per GPU.  Replaces the reference's
// host-staged MPI point-to-point exchange (CopyGhost.C:41-79: D2H copy, MPI_Isend/Irecv/Waitall, H2D
// copy) with device-to-device transfers enqueued on a CUDA stream.
//
// libnccl is dlopen'ed (the torch-bundled copy when the process already holds it, else the system
// one), so the C-ABI library itself has no link-time NCCL dependency and single-GPU users never load it.
#include "comm.h"

#include <dlfcn.h>

#include <functional>

#include <cstring>
#include <mutex>
#include <string>

#include "host_common.h"
#include "miniaero_b200.h"

namespace {

struct NcclId {
  char internal[128];
};
typedef void *NcclComm;
typedef int (*fn_get_unique_id)(NcclId *);
typedef int (*fn_comm_init_rank)(NcclComm *, int, NcclId, int);
typedef int (*fn_comm_destroy)(NcclComm);
typedef int (*fn_group)(void);
typedef int (*fn_sendrecv)(const void *, size_t, int, int, NcclComm, cudaStream_t);
typedef int (*fn_recv)(void *, size_t, int, int, NcclComm, cudaStream_t);
typedef const char *(*fn_errstr)(int);

struct NcclApi {
  void *handle = nullptr;
  fn_get_unique_id GetUniqueId = nullptr;
  fn_comm_init_rank CommInitRank = nullptr;
  fn_comm_destroy CommDestroy = nullptr;
  fn_group GroupStart = nullptr, GroupEnd = nullptr;
  fn_sendrecv Send = nullptr;
  fn_recv Recv = nullptr;
  fn_errstr GetErrorString = nullptr;
  std::string error;
};

void load_api(NcclApi &a) {
  const char *env = getenv("MINIAERO_NCCL_LIB");
  const char *names[] = {env, "libnccl.so.2", "libnccl.so"};
  for (const char *n : names) {
    if (!n || !*n) continue;
    a.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (a.handle) break;
  }
  if (!a.handle) {
    const char *e = dlerror();  // one call: dlerror() clears the message it returns
    a.error = std::string("cannot dlopen libnccl.so.2: ") + (e ? e : "?");
    return;
  }
#define MA_SYM(field, name)                                      \
  a.field = (decltype(a.field))dlsym(a.handle, name);            \
  if (!a.field) a.error = std::string("libnccl lacks ") + name;
  MA_SYM(GetUniqueId, "ncclGetUniqueId")
  MA_SYM(CommInitRank, "ncclCommInitRank")
  MA_SYM(CommDestroy, "ncclCommDestroy")
  MA_SYM(GroupStart, "ncclGroupStart")
  MA_SYM(GroupEnd, "ncclGroupEnd")
  MA_SYM(Send, "ncclSend")
  MA_SYM(Recv, "ncclRecv")
  MA_SYM(GetErrorString, "ncclGetErrorString")
#undef MA_SYM
}

NcclApi &api() {
  static NcclApi a;
  static std::once_flag once;
  std::call_once(once, load_api, std::ref(a));
  return a;
}

const int kNcclDouble = 8;  // ncclFloat64 (nccl.h ncclDataType_t)

int nccl_fail(const char *what, int rc) {
  NcclApi &a = api();
  return ma_set_error(MA_ERR_NCCL, std::string(what) + ": " + (a.GetErrorString ? a.GetErrorString(rc) : "?"));
}

}  // namespace

struct ma_comm {
  NcclComm comm = nullptr;
  int rank = 0, nranks = 1, device = 0;
};

namespace ma {

int comm_rank(const ma_comm *c) { return c ? c->rank : 0; }
int comm_size(const ma_comm *c) { return c ? c->nranks : 1; }

int comm_exchange(ma_comm *c, const double *sendbuf, double *recvbuf, int row, int npeers, const int *peer_rank,
                  const int *send_count, const int *recv_count, cudaStream_t st) {
  if (!c) return ma_set_error(MA_ERR_NCCL, "halo exchange requested without a communicator");
  NcclApi &a = api();
  if (!a.error.empty()) return ma_set_error(MA_ERR_NCCL, a.error);
  int rc = a.GroupStart();
  if (rc) return nccl_fail("ncclGroupStart", rc);
  size_t so = 0, ro = 0;
  for (int p = 0; p < npeers; ++p) {  // ascending peer rank, as CopyGhost.C:56-72
    const size_t ns = (size_t)send_count[p] * row, nr = (size_t)recv_count[p] * row;
    if (ns) {
      rc = a.Send(sendbuf + so, ns, kNcclDouble, peer_rank[p], c->comm, st);
      if (rc) {
        a.GroupEnd();  // never leave the thread in group mode
        return nccl_fail("ncclSend", rc);
      }
    }
    if (nr) {
      rc = a.Recv(recvbuf + ro, nr, kNcclDouble, peer_rank[p], c->comm, st);
      if (rc) {
        a.GroupEnd();
        return nccl_fail("ncclRecv", rc);
      }
    }
    so += ns;
    ro += nr;
  }
  rc = a.GroupEnd();
  if (rc) return nccl_fail("ncclGroupEnd", rc);
  return MA_OK;
}

}  // namespace ma

extern "C" {

int ma_comm_get_unique_id(unsigned char id[MA_COMM_ID_BYTES]) {
  NcclApi &a = api();
  if (!a.error.empty()) return ma_set_error(MA_ERR_NCCL, a.error);
  NcclId nid;
  int rc = a.GetUniqueId(&nid);
  if (rc) return nccl_fail("ncclGetUniqueId", rc);
  static_assert(sizeof(NcclId) == MA_COMM_ID_BYTES, "id size");
  std::memcpy(id, &nid, sizeof(nid));
  return MA_OK;
}

int ma_comm_create(const unsigned char id[MA_COMM_ID_BYTES], int num_ranks, int rank, int device, ma_comm **out) {
  if (!id || !out || num_ranks < 1 || rank < 0 || rank >= num_ranks)
    return ma_set_error(MA_ERR_INVALID, "ma_comm_create: bad arguments");
  *out = nullptr;
  NcclApi &a = api();
  if (!a.error.empty()) return ma_set_error(MA_ERR_NCCL, a.error);
  cudaError_t ce = cudaSetDevice(device);
  if (ce != cudaSuccess) return ma_set_error(MA_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(ce));
  NcclId nid;
  std::memcpy(&nid, id, sizeof(nid));
  ma_comm *c = new ma_comm();
  c->rank = rank, c->nranks = num_ranks, c->device = device;
  int rc = a.CommInitRank(&c->comm, num_ranks, nid, rank);
  if (rc) {
    delete c;
    return nccl_fail("ncclCommInitRank", rc);
  }
  *out = c;
  return MA_OK;
}

void ma_comm_destroy(ma_comm *c) {
  if (!c) return;
  if (c->comm) api().CommDestroy(c->comm);
  delete c;
}

}  // extern "C"
