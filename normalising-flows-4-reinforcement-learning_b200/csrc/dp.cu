// Data-parallel plumbing of the C ABI: one NCCL communicator per device, owned by the library, so that the gradient
// all-reduce can be issued from INSIDE the backward launch sequence (and is captured into its CUDA graph) instead
// of as a separate host-side call after it.  NCCL is bound at run time with dlopen("libnccl.so.2") -- the process
// that imports torch has already loaded it, so the library shares torch's NCCL build -- and the few entry points
// used are declared here (nccl.h: ncclGetUniqueId, ncclCommInitRank, ncclAllReduce, ncclBroadcast, ncclCommDestroy).
#include "dp.cuh"

#include <dlfcn.h>
#include <stdlib.h>

#include <mutex>

namespace nf {
namespace {

struct NcclId { char internal[128]; };               // ncclUniqueId (NCCL_UNIQUE_ID_BYTES = 128)
typedef void* Comm;                                   // ncclComm_t
typedef int (*GetUniqueIdFn)(NcclId*);
typedef int (*CommInitRankFn)(Comm*, int, NcclId, int);
// ncclConfig_t as of NCCL 2.27 (nccl.h ncclConfig_v22700); newer runtimes accept older versions of the structure
struct NcclConfig {
  size_t size; unsigned int magic; unsigned int version;
  int blocking, cgaClusterSize, minCTAs, maxCTAs; const char* netName; int splitShare, trafficClass;
  const char* commName; int collnetEnable, CTAPolicy, shrinkShare, nvlsCTAs;
};
constexpr int kUndefInt = -2147483647 - 1;   // NCCL_CONFIG_UNDEF_INT
typedef int (*CommInitRankConfigFn)(Comm*, int, NcclId, int, NcclConfig*);
typedef int (*AllReduceFn)(const void*, void*, size_t, int, int, Comm, cudaStream_t);
typedef int (*BroadcastFn)(const void*, void*, size_t, int, int, Comm, cudaStream_t);
typedef int (*CommDestroyFn)(Comm);
typedef const char* (*GetErrorStringFn)(int);
constexpr int kNcclFloat32 = 7, kNcclSum = 0, kNcclAvg = 4;

struct Api {
  void* handle = nullptr;
  GetUniqueIdFn get_unique_id = nullptr;
  CommInitRankFn comm_init_rank = nullptr;
  CommInitRankConfigFn comm_init_rank_config = nullptr;
  AllReduceFn all_reduce = nullptr;
  BroadcastFn broadcast = nullptr;
  CommDestroyFn comm_destroy = nullptr;
  GetErrorStringFn error_string = nullptr;
  bool tried = false;
};
Api g_api;
std::mutex g_mu;
struct DevComm { Comm comm = nullptr; int rank = 0, world = 0; };
DevComm g_comm[64];

int load_api() {
  if (g_api.tried) return g_api.all_reduce ? 0 : set_err(NF_E_UNSUPPORTED, "libnccl.so.2 is not available");
  g_api.tried = true;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    g_api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (g_api.handle) break;
  }
  NF_CHECK_ARG(g_api.handle != nullptr, NF_E_UNSUPPORTED, "dlopen(libnccl.so.2) failed: %s", dlerror());
  g_api.get_unique_id = reinterpret_cast<GetUniqueIdFn>(dlsym(g_api.handle, "ncclGetUniqueId"));
  g_api.comm_init_rank = reinterpret_cast<CommInitRankFn>(dlsym(g_api.handle, "ncclCommInitRank"));
  g_api.comm_init_rank_config = reinterpret_cast<CommInitRankConfigFn>(dlsym(g_api.handle, "ncclCommInitRankConfig"));
  g_api.broadcast = reinterpret_cast<BroadcastFn>(dlsym(g_api.handle, "ncclBroadcast"));
  g_api.comm_destroy = reinterpret_cast<CommDestroyFn>(dlsym(g_api.handle, "ncclCommDestroy"));
  g_api.error_string = reinterpret_cast<GetErrorStringFn>(dlsym(g_api.handle, "ncclGetErrorString"));
  g_api.all_reduce = reinterpret_cast<AllReduceFn>(dlsym(g_api.handle, "ncclAllReduce"));
  if (!g_api.get_unique_id || !g_api.comm_init_rank || !g_api.broadcast || !g_api.comm_destroy || !g_api.all_reduce) {
    g_api.all_reduce = nullptr;
    return set_err(NF_E_UNSUPPORTED, "libnccl.so.2 lacks an expected entry point");
  }
  return 0;
}

int nccl_rc(int rc, const char* what) {
  if (rc == 0) return 0;
  return set_err(1000 + rc, "%s failed: %s", what, g_api.error_string ? g_api.error_string(rc) : "nccl error");
}

DevComm* current() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  return &g_comm[dev];
}

}  // namespace

int dp_world() {
  DevComm* c = current();
  return (c && c->comm) ? c->world : 0;
}

int dp_allreduce(float* buf, int64_t n, bool average, cudaStream_t st) {
  DevComm* c = current();
  NF_CHECK_ARG(c && c->comm, NF_E_BADARG, "nf_dp_init has not been called on this device");
  if (n <= 0) return 0;
  return nccl_rc(g_api.all_reduce(buf, buf, (size_t)n, kNcclFloat32, average ? kNcclAvg : kNcclSum, c->comm, st),
                 "ncclAllReduce");
}

}  // namespace nf

using namespace nf;

extern "C" {

int nf_dp_unique_id(void* id128) {
  NF_CHECK_ARG(id128 != nullptr, NF_E_BADARG, "id buffer is NULL");
  std::lock_guard<std::mutex> lk(g_mu);
  NF_TRY(load_api());
  NcclId id;
  NF_TRY(nccl_rc(g_api.get_unique_id(&id), "ncclGetUniqueId"));
  memcpy(id128, &id, sizeof(id));
  return 0;
}

int nf_dp_init(const void* id128, int32_t rank, int32_t world) {
  NF_CHECK_ARG(id128 != nullptr && world >= 1 && rank >= 0 && rank < world, NF_E_BADARG, "bad rank %d / world %d", rank, world);
  std::lock_guard<std::mutex> lk(g_mu);
  NF_TRY(load_api());
  DevComm* c = current();
  NF_CHECK_ARG(c != nullptr, NF_E_BADARG, "no current CUDA device");
  if (c->comm) { g_api.comm_destroy(c->comm); c->comm = nullptr; }
  NcclId id;
  memcpy(&id, id128, sizeof(id));
  Comm comm = nullptr;
  // NF_B200_DP_MAX_CTAS caps the communicator's CTAs (experiment knob; default 0 = NCCL's own choice).  Measured on
  // 2 B200s (C2a, 4096 rows per GPU): capping to 8 / 4 / 2 CTAs makes the step SLOWER (0.85 / 0.89 / 0.98 ms against
  // 0.78 ms) -- the payload is latency-bound and the last bucket's all-reduce is exposed whatever runs beside it.
  const char* e = getenv("NF_B200_DP_MAX_CTAS");
  const int max_ctas = e ? atoi(e) : 0;
  if (max_ctas > 0 && g_api.comm_init_rank_config) {
    NcclConfig cfg;
    cfg.size = sizeof(NcclConfig); cfg.magic = 0xcafebeef; cfg.version = 22703;
    cfg.blocking = cfg.cgaClusterSize = cfg.minCTAs = kUndefInt;
    cfg.maxCTAs = max_ctas;
    cfg.netName = nullptr; cfg.splitShare = cfg.trafficClass = kUndefInt; cfg.commName = nullptr;
    cfg.collnetEnable = cfg.CTAPolicy = cfg.shrinkShare = cfg.nvlsCTAs = kUndefInt;
    NF_TRY(nccl_rc(g_api.comm_init_rank_config(&comm, world, id, rank, &cfg), "ncclCommInitRankConfig"));
  } else {
    NF_TRY(nccl_rc(g_api.comm_init_rank(&comm, world, id, rank), "ncclCommInitRank"));
  }
  c->comm = comm; c->rank = rank; c->world = world;
  return 0;
}

int nf_dp_world(void) { return dp_world(); }

int nf_dp_allreduce(float* buf, int64_t n, int32_t average, void* stream) {
  NF_TRY(check_ptr(buf, "buf", n > 0));
  return dp_allreduce(buf, n, average != 0, reinterpret_cast<cudaStream_t>(stream));
}

int nf_dp_broadcast(float* buf, int64_t n, int32_t root, void* stream) {
  NF_TRY(check_ptr(buf, "buf", n > 0));
  DevComm* c = current();
  NF_CHECK_ARG(c && c->comm, NF_E_BADARG, "nf_dp_init has not been called on this device");
  if (n <= 0) return 0;
  return nccl_rc(g_api.broadcast(buf, buf, (size_t)n, kNcclFloat32, root, c->comm, reinterpret_cast<cudaStream_t>(stream)),
                 "ncclBroadcast");
}

int nf_dp_shutdown(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  DevComm* c = current();
  if (c && c->comm) {
    g_api.comm_destroy(c->comm);
    c->comm = nullptr; c->world = 0;
  }
  return 0;
}

}  // extern "C"
