// K12 — data-parallel collectives of the BC path (pvr_comm_init / allreduce / reduce_scatter /
// allgather): one NCCL communicator per process (= per GPU), NVLink 5 / NVSwitch inside one node.
//
// The reference has no distributed code; what is exchanged is decided by the host mirror
// (pvr_habitat_b200/parallel.py): BatchNorm1d sums (2 D doubles) in the forward, the flat gradient in buckets during the
// backward, scalars. Calling NCCL directly (instead of through torch.distributed's process group) keeps the
// collectives plain stream-ordered launches: they are captured into the whole-step CUDA graph like any kernel and can
// be forked onto a second stream so that a bucket's all-reduce overlaps the weight-gradient GEMMs that follow it.
//
// NCCL is bound at run time (dlopen of the libnccl.so.2 that ships with torch — already loaded in a torch process), so
// the library has no link-time dependency on it and single-GPU users never touch it.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdint.h>
#include <string.h>

#include <mutex>

#include "pvr_b200.h"

extern void pvr_set_error(const char* fmt, ...);

namespace {

// the few NCCL declarations used (ABI-stable since NCCL 2.x; nccl.h is not needed to build)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclSuccess = 0 };
enum { ncclFloat32 = 7, ncclFloat64 = 8, ncclBfloat16 = 9, ncclInt64 = 4 };  // ncclDataType_t
enum { ncclSum = 0 };

struct Api {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*ReduceScatter)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
};
Api g_api;
std::mutex g_mu;

bool load(const char* path) {
  std::lock_guard<std::mutex> lock(g_mu);
  if (g_api.handle) return true;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h && path && path[0]) h = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
  if (!h) {
    pvr_set_error("pvr_comm: cannot load libnccl.so.2 (%s)", dlerror());
    return false;
  }
#define SYM(field, name)                                                     \
  *reinterpret_cast<void**>(&g_api.field) = dlsym(h, name);                  \
  if (!g_api.field) {                                                        \
    pvr_set_error("pvr_comm: libnccl.so.2 has no symbol %s", name);          \
    dlclose(h);                                                              \
    return false;                                                            \
  }
  SYM(GetUniqueId, "ncclGetUniqueId")
  SYM(CommInitRank, "ncclCommInitRank")
  SYM(CommDestroy, "ncclCommDestroy")
  SYM(AllReduce, "ncclAllReduce")
  SYM(ReduceScatter, "ncclReduceScatter")
  SYM(AllGather, "ncclAllGather")
  SYM(Broadcast, "ncclBroadcast")
  SYM(GetErrorString, "ncclGetErrorString")
  SYM(GetVersion, "ncclGetVersion")
#undef SYM
  g_api.handle = h;
  return true;
}

int dtype_of(int dtype) {
  switch (dtype) {
    case PVR_COMM_F32: return ncclFloat32;
    case PVR_COMM_F64: return ncclFloat64;
    case PVR_COMM_BF16: return ncclBfloat16;
    case PVR_COMM_I64: return ncclInt64;
    default: return -1;
  }
}

int fail(const char* what, ncclResult_t r) {
  pvr_set_error("%s: NCCL error %d (%s)", what, (int)r, g_api.GetErrorString ? g_api.GetErrorString(r) : "?");
  return PVR_ERR_CUDA;
}

}  // namespace

extern "C" int pvr_comm_load(const char* libnccl_path) { return load(libnccl_path) ? PVR_OK : PVR_ERR_CUDA; }

extern "C" int pvr_comm_version(void) {
  int v = 0;
  if (!g_api.handle || g_api.GetVersion(&v) != ncclSuccess) return 0;
  return v;
}

extern "C" int pvr_comm_unique_id(void* id128) {
  if (!id128 || !load(nullptr)) {
    if (!id128) pvr_set_error("pvr_comm_unique_id: invalid argument");
    return PVR_ERR_ARG;
  }
  ncclUniqueId id;
  ncclResult_t r = g_api.GetUniqueId(&id);
  if (r != ncclSuccess) return fail("pvr_comm_unique_id", r);
  memcpy(id128, &id, sizeof(id));
  return PVR_OK;
}

extern "C" int pvr_comm_init(int rank, int world, const void* id128, void** comm_out) {
  if (!id128 || !comm_out || world < 1 || rank < 0 || rank >= world || !load(nullptr)) {
    if (id128 && comm_out) pvr_set_error("pvr_comm_init: invalid argument or NCCL unavailable");
    return PVR_ERR_ARG;
  }
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  ncclComm_t c = nullptr;
  ncclResult_t r = g_api.CommInitRank(&c, world, id, rank);  // binds to the current CUDA device
  if (r != ncclSuccess) return fail("pvr_comm_init", r);
  *comm_out = c;
  return PVR_OK;
}

extern "C" int pvr_comm_destroy(void* comm) {
  if (!comm || !g_api.handle) return PVR_OK;
  ncclResult_t r = g_api.CommDestroy(static_cast<ncclComm_t>(comm));
  return r == ncclSuccess ? PVR_OK : fail("pvr_comm_destroy", r);
}

extern "C" int pvr_comm_allreduce(void* comm, void* buf, int64_t count, int dtype, void* stream) {
  const int dt = dtype_of(dtype);
  if (!comm || !buf || count <= 0 || dt < 0) {
    pvr_set_error("pvr_comm_allreduce: invalid argument");
    return PVR_ERR_ARG;
  }
  ncclResult_t r = g_api.AllReduce(buf, buf, (size_t)count, dt, ncclSum, static_cast<ncclComm_t>(comm),
                                   static_cast<cudaStream_t>(stream));
  return r == ncclSuccess ? PVR_OK : fail("pvr_comm_allreduce", r);
}

extern "C" int pvr_comm_reduce_scatter(void* comm, const void* send, void* recv, int64_t recv_count, int dtype,
                                       void* stream) {
  const int dt = dtype_of(dtype);
  if (!comm || !send || !recv || recv_count <= 0 || dt < 0) {
    pvr_set_error("pvr_comm_reduce_scatter: invalid argument");
    return PVR_ERR_ARG;
  }
  ncclResult_t r = g_api.ReduceScatter(send, recv, (size_t)recv_count, dt, ncclSum, static_cast<ncclComm_t>(comm),
                                       static_cast<cudaStream_t>(stream));
  return r == ncclSuccess ? PVR_OK : fail("pvr_comm_reduce_scatter", r);
}

extern "C" int pvr_comm_allgather(void* comm, const void* send, void* recv, int64_t send_count, int dtype,
                                  void* stream) {
  const int dt = dtype_of(dtype);
  if (!comm || !send || !recv || send_count <= 0 || dt < 0) {
    pvr_set_error("pvr_comm_allgather: invalid argument");
    return PVR_ERR_ARG;
  }
  ncclResult_t r = g_api.AllGather(send, recv, (size_t)send_count, dt, static_cast<ncclComm_t>(comm),
                                   static_cast<cudaStream_t>(stream));
  return r == ncclSuccess ? PVR_OK : fail("pvr_comm_allgather", r);
}

extern "C" int pvr_comm_broadcast(void* comm, void* buf, int64_t count, int dtype, int root, void* stream) {
  const int dt = dtype_of(dtype);
  if (!comm || !buf || count <= 0 || dt < 0) {
    pvr_set_error("pvr_comm_broadcast: invalid argument");
    return PVR_ERR_ARG;
  }
  ncclResult_t r = g_api.Broadcast(buf, buf, (size_t)count, dt, root, static_cast<ncclComm_t>(comm),
                                   static_cast<cudaStream_t>(stream));
  return r == ncclSuccess ? PVR_OK : fail("pvr_comm_broadcast", r);
}
