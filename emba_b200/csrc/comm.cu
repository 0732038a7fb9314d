// Combination of the per-GPU partial systems over NVLink: NCCL all-reduce of the
// num_ev_map histogram, the cost scalars, A11/b1, A22/b2 and the A12 strips on the handle's stream. NCCL is
// loaded lazily (dlopen) so that the library has no hard dependency on it for single-GPU use.
#include <dlfcn.h>

#include <cstring>

#include "emba_internal.cuh"

namespace emba {

typedef struct { char internal[128]; } nccl_uid_t;
typedef int (*fn_get_uid)(nccl_uid_t*);
typedef int (*fn_init_rank)(void**, int, nccl_uid_t, int);
typedef int (*fn_allreduce)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*fn_destroy)(void*);
typedef const char* (*fn_errstr)(int);

struct NcclApi {
  void* lib = nullptr;
  fn_get_uid get_uid = nullptr;
  fn_init_rank init_rank = nullptr;
  fn_allreduce allreduce = nullptr;
  fn_destroy destroy = nullptr;
  fn_errstr errstr = nullptr;
};

static NcclApi* nccl_api() {
  static NcclApi api;
  if (api.lib) return &api;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* nm : names) {
    api.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (api.lib) break;
  }
  if (!api.lib) return nullptr;
  api.get_uid = (fn_get_uid)dlsym(api.lib, "ncclGetUniqueId");
  api.init_rank = (fn_init_rank)dlsym(api.lib, "ncclCommInitRank");
  api.allreduce = (fn_allreduce)dlsym(api.lib, "ncclAllReduce");
  api.destroy = (fn_destroy)dlsym(api.lib, "ncclCommDestroy");
  api.errstr = (fn_errstr)dlsym(api.lib, "ncclGetErrorString");
  if (!api.get_uid || !api.init_rank || !api.allreduce || !api.destroy) { api.lib = nullptr; return nullptr; }
  return &api;
}

// dtype: 0 = int32 sum, 1 = fp64 sum, 2 = int32 min, 3 = int32 max
int comm_allreduce(Handle* h, void* buf, int64_t count, int dtype) {
  if (h->world <= 1 || count <= 0) return EMBA_OK;
  NcclApi* api = nccl_api();
  if (!api || !h->nccl_comm) { h->err = "multi-GPU shard without a communicator: call emba_comm_init"; return EMBA_E_NCCL; }
  // ncclDataType_t: ncclInt32 = 2, ncclFloat64 = 8; ncclRedOp_t: sum 0, prod 1, max 2, min 3
  const int ty = (dtype == 1) ? 8 : 2;
  const int op = (dtype == 2) ? 3 : (dtype == 3) ? 2 : 0;
  const int r = api->allreduce(buf, buf, (size_t)count, ty, op, h->nccl_comm, h->stream);
  if (r != 0) { h->err = std::string("ncclAllReduce: ") + (api->errstr ? api->errstr(r) : "error"); return EMBA_E_NCCL; }
  h->launches++;
  return EMBA_OK;
}

void comm_destroy(Handle* h) {
  if (h->nccl_comm) {
    NcclApi* api = nccl_api();
    if (api) api->destroy(h->nccl_comm);
    h->nccl_comm = nullptr;
  }
}

}  // namespace emba

using namespace emba;

extern "C" {

int emba_comm_unique_id(void* out128) {
  if (!out128) return EMBA_E_ARG;
  NcclApi* api = nccl_api();
  if (!api) return EMBA_E_NCCL;
  nccl_uid_t id;
  if (api->get_uid(&id) != 0) return EMBA_E_NCCL;
  std::memcpy(out128, &id, 128);
  return EMBA_OK;
}

int emba_comm_init(emba_handle_t hh, const void* id128, int32_t rank, int32_t world) {
  Handle* h = (Handle*)hh;
  if (!h || !id128) return EMBA_E_ARG;
  if (world < 1 || rank < 0 || rank >= world) { h->err = "emba_comm_init: bad rank/world"; return EMBA_E_ARG; }
  NcclApi* api = nccl_api();
  if (!api) { h->err = "emba_comm_init: libnccl.so.2 not found"; return EMBA_E_NCCL; }
  EMBA_CUDA(cudaSetDevice(h->device));
  comm_destroy(h);
  nccl_uid_t id;
  std::memcpy(&id, id128, 128);
  const int r = api->init_rank(&h->nccl_comm, world, id, rank);
  if (r != 0) { h->err = std::string("ncclCommInitRank: ") + (api->errstr ? api->errstr(r) : "error"); h->nccl_comm = nullptr; return EMBA_E_NCCL; }
  return emba_set_shard(hh, rank, world);
}

}  // extern "C"
