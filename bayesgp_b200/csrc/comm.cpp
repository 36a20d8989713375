// comm.cpp — NCCL plumbing for observation-sharded models (one process per GPU).
//
// The reference has no distributed path; this is new.  When rows of A are
// sharded across ranks, each Newton iteration all-reduces the packed likelihood partials
// [g_lik | ll | sumsq | nonfinite] and the p x p likelihood Hessian over NVLink; every rank then
// runs the identical prior / Cholesky / step kernels, so replicas stay bit-identical.
// NCCL is loaded with dlopen at communicator creation: single-GPU use never needs it, and when
// the host process already has torch's libnccl.so.2 mapped the same copy is reused.
#include <dlfcn.h>

#include "bgp_internal.h"

namespace bgp {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclSuccess = 0 };
enum { ncclFloat64 = 8 };   // ncclDataType_t: int8 0, uint8 1, int32 2, uint32 3, int64 4, uint64 5, f16 6, f32 7, f64 8
enum { ncclSum = 0 };

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi* nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api.handle ? &api : nullptr;
  tried = true;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* nm : names) {
    api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (api.handle) break;
  }
  if (!api.handle) return nullptr;
  api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.handle, "ncclGetUniqueId");
  api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.handle, "ncclCommInitRank");
  api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.handle, "ncclCommDestroy");
  api.AllReduce = (decltype(api.AllReduce))dlsym(api.handle, "ncclAllReduce");
  api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.handle, "ncclGetErrorString");
  if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllReduce) {
    dlclose(api.handle);
    api.handle = nullptr;
    return nullptr;
  }
  return &api;
}

struct Comm {
  ncclComm_t comm = nullptr;
};

int comm_unique_id(void* id128) {
  NcclApi* api = nccl_api();
  if (!api) {
    set_error("libnccl.so.2 could not be loaded");
    return BGP_ERR_NCCL;
  }
  ncclUniqueId id;
  ncclResult_t r = api->GetUniqueId(&id);
  if (r != ncclSuccess) {
    set_error("ncclGetUniqueId failed: %s", api->GetErrorString ? api->GetErrorString(r) : "?");
    return BGP_ERR_NCCL;
  }
  memcpy(id128, &id, sizeof(id));
  return BGP_OK;
}

int comm_open(Comm** out, int rank, int world, const void* id128) {
  NcclApi* api = nccl_api();
  if (!api) {
    set_error("libnccl.so.2 could not be loaded");
    return BGP_ERR_NCCL;
  }
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  Comm* c = new Comm();
  ncclResult_t r = api->CommInitRank(&c->comm, world, id, rank);
  if (r != ncclSuccess) {
    set_error("ncclCommInitRank failed: %s", api->GetErrorString ? api->GetErrorString(r) : "?");
    delete c;
    return BGP_ERR_NCCL;
  }
  *out = c;
  return BGP_OK;
}

void comm_close(Comm** c) {
  if (!c || !*c) return;
  NcclApi* api = nccl_api();
  if (api && (*c)->comm) api->CommDestroy((*c)->comm);
  delete *c;
  *c = nullptr;
}

int comm_sum(Comm* c, double* buf, size_t count, cudaStream_t st) {
  NcclApi* api = nccl_api();
  if (!api || !c) {
    set_error("collective on a model without a communicator");
    return BGP_ERR_NCCL;
  }
  ncclResult_t r = api->AllReduce(buf, buf, count, ncclFloat64, ncclSum, c->comm, st);
  if (r != ncclSuccess) {
    set_error("ncclAllReduce failed: %s", api->GetErrorString ? api->GetErrorString(r) : "?");
    return BGP_ERR_NCCL;
  }
  return BGP_OK;
}

int comm_create(bgp_model* m, const void* id128) { return comm_open(&m->comm, m->rank, m->world, id128); }

void comm_destroy(bgp_model* m) {
  comm_close(&m->comm);
  comm_close(&m->node_comm);
}

int comm_allreduce_sum(bgp_model* m, double* buf, size_t count) {
  if (m->world <= 1) return BGP_OK;
  return comm_sum(m->comm, buf, count, m->stream);
}

// Node group: ranks holding replicas of the same rows split quadrature nodes, sample blocks and prediction rows.
// Every piece has exactly one owner and the other ranks contribute zeros, so a SUM all-reduce assembles the whole
// bit-exactly (x + 0 = x) whatever the reduction order.
int node_allreduce_sum(bgp_model* m, double* buf, size_t count) {
  if (m->node_world <= 1) return BGP_OK;
  return comm_sum(m->node_comm, buf, count, m->stream);
}

}  // namespace bgp
