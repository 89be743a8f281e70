// NCCL communicator owned by the library, so that the sharded bundle adjustment issues its ONE all-reduce of the reduced
// camera system per LM trial from C, on the handle's own stream — no host callback, no Python in the
// data path, and the call is capturable into the LM loop's CUDA graph.  libnccl.so.2 is resolved at run time (the copy
// torch already loaded when the host is a torch.distributed job, else the system one): the library keeps no link-time
// dependency and single-GPU users never load NCCL.  Rendez-vous is the host's business: rank 0 creates the 128-byte
// unique id, the launcher (torch.distributed broadcast, MPI, a file...) hands it to the other ranks.
#include <dlfcn.h>
#include <nccl.h>

#include <mutex>

#include "common.cuh"

struct vieo_comm {
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1, device = 0;
};

namespace {
struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_nccl;
std::once_flag g_once;

bool nccl_load() {
  std::call_once(g_once, [] {
    void* l = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!l) l = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!l) return;
    g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))dlsym(l, "ncclGetUniqueId");
    g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))dlsym(l, "ncclCommInitRank");
    g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))dlsym(l, "ncclCommDestroy");
    g_nccl.AllReduce = (decltype(g_nccl.AllReduce))dlsym(l, "ncclAllReduce");
    g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))dlsym(l, "ncclGetErrorString");
    if (g_nccl.GetUniqueId && g_nccl.CommInitRank && g_nccl.CommDestroy && g_nccl.AllReduce && g_nccl.GetErrorString) g_nccl.lib = l;
  });
  if (!g_nccl.lib) vieo::set_error("libnccl.so.2 not found or incomplete (dlopen): %s", dlerror() ? dlerror() : "missing symbols");
  return g_nccl.lib != nullptr;
}
#define NCCL_CK(call)                                                                              \
  do {                                                                                             \
    ncclResult_t r_ = (call);                                                                      \
    if (r_ != ncclSuccess) {                                                                       \
      vieo::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString(r_));      \
      return VIEO_E_CUDA;                                                                          \
    }                                                                                              \
  } while (0)
}  // namespace

namespace vieo {
// sum `count` doubles in place over the communicator, ordered on `stream` (used by ba.cu)
int comm_allreduce_f64(vieo_comm* c, double* buf, size_t count, cudaStream_t stream) {
  NCCL_CK(g_nccl.AllReduce(buf, buf, count, ncclDouble, ncclSum, c->comm, stream));
  return VIEO_OK;
}
void comm_info(const vieo_comm* c, int* rank, int* world) {
  *rank = c->rank;
  *world = c->world;
}
}  // namespace vieo

extern "C" {

int vieo_comm_unique_id(uint8_t id[VIEO_COMM_ID_BYTES]) {
  VIEO_ARG(id, "null argument");
  static_assert(sizeof(ncclUniqueId) == VIEO_COMM_ID_BYTES, "ncclUniqueId size");
  if (!nccl_load()) return VIEO_E_CUDA;
  ncclUniqueId u;
  NCCL_CK(g_nccl.GetUniqueId(&u));
  memcpy(id, &u, sizeof(u));
  return VIEO_OK;
}

int vieo_comm_create(const uint8_t id[VIEO_COMM_ID_BYTES], int rank, int world, int device, vieo_comm_t** out) {
  VIEO_ARG(id && out && world >= 1 && rank >= 0 && rank < world, "bad argument");
  int rc = vieo::use_device(device);
  if (rc) return rc;
  if (!nccl_load()) return VIEO_E_CUDA;
  ncclUniqueId u;
  memcpy(&u, id, sizeof(u));
  vieo_comm* c = new vieo_comm();
  c->rank = rank; c->world = world; c->device = device;
  ncclResult_t r = g_nccl.CommInitRank(&c->comm, world, u, rank);
  if (r != ncclSuccess) {
    vieo::set_error("ncclCommInitRank -> %s", g_nccl.GetErrorString(r));
    delete c;
    return VIEO_E_CUDA;
  }
  *out = c;
  return VIEO_OK;
}

void vieo_comm_destroy(vieo_comm_t* c) {
  if (!c) return;
  if (c->comm && g_nccl.lib) {
    cudaSetDevice(c->device);
    g_nccl.CommDestroy(c->comm);
  }
  delete c;
}

/* sum `count` fp64 in place over the ranks, ordered on `stream` (exposed for tests of the communicator itself) */
int vieo_comm_allreduce_f64(vieo_comm_t* c, double* buf_dev, size_t count, void* stream) {
  VIEO_ARG(c && buf_dev, "null argument");
  return vieo::comm_allreduce_f64(c, buf_dev, count, (cudaStream_t)stream);
}

}  // extern "C"
