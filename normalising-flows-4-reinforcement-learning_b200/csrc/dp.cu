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

// ---- gradient all-reduce over NVLink peer memory -------------------------------------------------------------
// A 4 MB gradient arena is latency-bound for a library collective (~0.06 ms of a 0.63 ms step at 8 ranks, and the
// whole arena becomes final only with the last weight-gradient kernel, so there is nothing to overlap it with).  The
// ranks of one box can write each other's memory through NVSwitch, so small arenas are averaged by ONE kernel per rank:
//   scatter  rank r stores slice p of its arena into slot r of peer p's receive region   (peer stores, 128-bit)
//   reduce   rank r sums the world slots of its own slice in rank order (the same order on every rank: results are
//            bit-identical across replicas), scales, and stores the result into slice r of every peer's result region
//   gather   rank r copies the other slices of its result region into its arena
// Synchronisation is per CTA index: CTA b of every rank owns the same element range of every slice, so it waits only
// for the flags CTA b of each peer raised (fence + release store into the peer's flag array, acquire poll on its own),
// never for another CTA of its own grid -- no co-residency requirement, no grid barrier.  Flags carry a call counter
// kept in device memory, so the launch is replayable from a captured CUDA graph.  Buffers are exchanged once as CUDA
// IPC handles (nf_dp_p2p_alloc / nf_dp_p2p_connect); arenas above the threshold, or any failure to connect, use NCCL.
constexpr int kP2pCtas = 64, kP2pThreads = 512, kP2pMaxWorld = 8;
constexpr int64_t kP2pFlagBytesAligned = ((2 * kP2pMaxWorld * kP2pCtas + kP2pCtas + 16) * 4 + 255) / 256 * 256;

struct P2pPeers {
  float* recv[kP2pMaxWorld];      // receive region of every rank (world slots of one slice)
  float* result[kP2pMaxWorld];    // result region of every rank (world slices)
  uint32_t* flags[kP2pMaxWorld];  // [phase 2][source rank kP2pMaxWorld][CTA kP2pCtas]
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// The CTA barrier orders every thread's peer stores before thread p's release store (system scope, cumulative) of this
// CTA's flag of `phase` on peer p -- one fence per flag, not one per thread (a membar.sys in all 32 K threads cost
// ~12 us per phase); thread s then waits for peer s's flag of the same CTA index.  A peer that never arrives (a rank
// that skipped the call) ends the wait after ~5 min with the error word set instead of hanging the device.
__device__ __forceinline__ void p2p_exchange_flags(const P2pPeers& pp, int rank, int world, int phase, uint32_t epoch,
                                                   int* err) {
  __syncthreads();
  // lanes 0..world-1 of warp 0 serve the peers (one flag thread per warp was measured 3.5 us slower at 8 ranks)
  const int t = threadIdx.x;
  if (t < world && t != rank) {
    st_release_sys(pp.flags[t] + (phase * kP2pMaxWorld + rank) * kP2pCtas + blockIdx.x, epoch);
    const uint32_t* f = pp.flags[rank] + (phase * kP2pMaxWorld + t) * kP2pCtas + blockIdx.x;
    const long long t0 = clock64();
    while ((int32_t)(ld_acquire_sys(f) - epoch) < 0) {
      if (clock64() - t0 > 600000000000LL) { *err = 1; break; }
    }
  }
  __syncthreads();
}

__device__ __forceinline__ void p2p_stamp(int* err, int slot) {   // CTA 0's phase boundaries in ns (nf_dp_p2p_timeline)
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    reinterpret_cast<unsigned long long*>(err + 2)[slot] = t;
  }
}

__global__ void __launch_bounds__(kP2pThreads) k_p2p_allreduce(float* __restrict__ g, long long n, long long slice,
                                                               P2pPeers pp, int rank, int world, float scale,
                                                               uint32_t* epochs, int* err) {
  NF_PDL_ENTRY();
  p2p_stamp(err, 0);
  const uint32_t epoch = epochs[blockIdx.x] + 1;
  const long long s4 = slice >> 2;                                  // float4 per slice
  const long long per = (s4 + gridDim.x - 1) / gridDim.x;
  const long long lo = blockIdx.x * per, hi = min(lo + per, s4);
  const bool al = (reinterpret_cast<uintptr_t>(g) & 15) == 0;       // slices are multiples of 4 floats
  auto load_g = [&](long long e) -> float4 {                        // arena element e .. e+3, zero past the end
    if (al && e + 3 < n) return *reinterpret_cast<const float4*>(g + e);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (e < n) v.x = g[e];
    if (e + 1 < n) v.y = g[e + 1];
    if (e + 2 < n) v.z = g[e + 2];
    if (e + 3 < n) v.w = g[e + 3];
    return v;
  };
  auto store_g = [&](long long e, float4 v) {
    if (al && e + 3 < n) { *reinterpret_cast<float4*>(g + e) = v; return; }
    if (e < n) g[e] = v.x;
    if (e + 1 < n) g[e + 1] = v.y;
    if (e + 2 < n) g[e + 2] = v.z;
    if (e + 3 < n) g[e + 3] = v.w;
  };
  // scatter: slice p of this rank's arena -> slot `rank` of peer p
  for (int k = 1; k < world; ++k) {
    const int p = (rank + k) % world;
    float4* dst = reinterpret_cast<float4*>(pp.recv[p] + rank * slice);
    for (long long i = lo + threadIdx.x; i < hi; i += blockDim.x) dst[i] = load_g(p * slice + (i << 2));
  }
  p2p_stamp(err, 1);
  p2p_exchange_flags(pp, rank, world, 0, epoch, err);
  p2p_stamp(err, 2);
  // reduce this rank's slice in rank order, publish it to every peer
  for (long long i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = 0; s < world; ++s) {
      const float4 v = s == rank ? load_g(rank * slice + (i << 2))
                                 : __ldcg(reinterpret_cast<const float4*>(pp.recv[rank] + s * slice) + i);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    acc.x *= scale; acc.y *= scale; acc.z *= scale; acc.w *= scale;
    store_g(rank * slice + (i << 2), acc);
    for (int k = 1; k < world; ++k) {
      const int p = (rank + k) % world;
      reinterpret_cast<float4*>(pp.result[p] + rank * slice)[i] = acc;
    }
  }
  p2p_stamp(err, 3);
  p2p_exchange_flags(pp, rank, world, 1, epoch, err);
  p2p_stamp(err, 4);
  // gather: the peers' reduced slices -> arena
  for (int k = 1; k < world; ++k) {
    const int s = (rank + k) % world;
    const float4* src = reinterpret_cast<const float4*>(pp.result[rank] + s * slice);
    for (long long i = lo + threadIdx.x; i < hi; i += blockDim.x) store_g(s * slice + (i << 2), __ldcg(src + i));
  }
  p2p_stamp(err, 5);
  if (threadIdx.x == 0) epochs[blockIdx.x] = epoch;
}

struct P2pState {
  void* base = nullptr;            // this rank's buffer: [flags + call counters + error word | receive | result]
  int64_t cap = 0;                 // floats per region
  void* peer_base[kP2pMaxWorld] = {};
  P2pPeers peers{};
  int rank = 0, world = 0;
  bool connected = false;
  int64_t max_floats = 0;          // arenas up to this size take the peer-memory path
  int64_t min_floats = 0;          // ... and at least this size when fewer than 8 ranks take part (see nf_dp_p2p_alloc)
};
P2pState g_p2p[64];

P2pState* p2p_current() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  return &g_p2p[dev];
}

void p2p_release(P2pState* s) {
  if (!s) return;
  for (int p = 0; p < kP2pMaxWorld; ++p)
    if (s->peer_base[p] && s->peer_base[p] != s->base) cudaIpcCloseMemHandle(s->peer_base[p]);
  if (s->base) cudaFree(s->base);
  *s = P2pState();
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
  P2pState* s = p2p_current();
  if (s && s->connected && s->world == c->world && n <= s->max_floats && (s->world >= 8 || n >= s->min_floats)) {
    const int64_t slice = (cdiv(n, (int64_t)s->world) + 3) & ~int64_t(3);
    uint32_t* epochs = reinterpret_cast<uint32_t*>(s->base) + 2 * kP2pMaxWorld * kP2pCtas;
    int* err = reinterpret_cast<int*>(epochs + kP2pCtas);
    NF_LAUNCH((k_p2p_allreduce), kP2pCtas, kP2pThreads, 0, st, buf, (long long)n, (long long)slice, s->peers, s->rank,
              s->world, average ? 1.0f / (float)s->world : 1.0f, epochs, err);
    NF_LAUNCHED();
    return 0;
  }
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

int nf_dp_p2p_alloc(int64_t max_floats, void* handle64) {
  NF_CHECK_ARG(handle64 != nullptr && max_floats > 0, NF_E_BADARG, "bad arguments");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
  std::lock_guard<std::mutex> lk(g_mu);
  P2pState* s = p2p_current();
  NF_CHECK_ARG(s != nullptr, NF_E_BADARG, "no current CUDA device");
  p2p_release(s);
  // a slice is the arena divided by the world size rounded up to 4 floats: world * slice <= max_floats + 4 * world
  s->cap = ((max_floats + 4 * kP2pMaxWorld + 63) / 64) * 64;
  const size_t bytes = (size_t)kP2pFlagBytesAligned + 2 * (size_t)s->cap * sizeof(float);
  NF_CUDA(cudaMalloc(&s->base, bytes));
  NF_CUDA(cudaMemset(s->base, 0, bytes));
  NF_CUDA(cudaDeviceSynchronize());
  cudaIpcMemHandle_t h;
  NF_CUDA(cudaIpcGetMemHandle(&h, s->base));
  memcpy(handle64, &h, sizeof(h));
  s->max_floats = max_floats;
  // Measured against NCCL 2.28 on the same buffers (profiles/r2_p2p_allreduce.txt): at 8 ranks the kernel wins at every size
  // (0.25 MB: 27 vs 30 us, 4 MB: 34 vs 44 us); at 2 and 4 ranks NCCL's low-latency protocol is ahead up to 1 MB (2 ranks,
  // 1 MB: 16 vs 20 us) and behind from 4 MB on (24 vs 37 us).  NF_B200_DP_P2P_MIN_BYTES moves the switch-over.
  const char* e = getenv("NF_B200_DP_P2P_MIN_BYTES");
  s->min_floats = (e ? atoll(e) : (2ll << 20)) / 4;
  return 0;
}

int nf_dp_p2p_connect(const void* handles, int32_t rank, int32_t world) {
  NF_CHECK_ARG(handles != nullptr && world >= 2 && world <= kP2pMaxWorld && rank >= 0 && rank < world, NF_E_BADARG,
               "bad rank %d / world %d (peer-memory all-reduce: 2..%d ranks of one box)", rank, world, kP2pMaxWorld);
  std::lock_guard<std::mutex> lk(g_mu);
  P2pState* s = p2p_current();
  NF_CHECK_ARG(s != nullptr && s->base != nullptr, NF_E_BADARG, "nf_dp_p2p_alloc has not been called on this device");
  for (int p = 0; p < world; ++p) {
    if (p == rank) { s->peer_base[p] = s->base; continue; }
    cudaIpcMemHandle_t h;
    memcpy(&h, static_cast<const char*>(handles) + (size_t)p * sizeof(h), sizeof(h));
    void* ptr = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      (void)cudaGetLastError();
      return set_err(NF_E_UNSUPPORTED, "cudaIpcOpenMemHandle(rank %d): %s", p, cudaGetErrorString(e));
    }
    s->peer_base[p] = ptr;
  }
  for (int p = 0; p < world; ++p) {
    char* b = static_cast<char*>(s->peer_base[p]);
    s->peers.flags[p] = reinterpret_cast<uint32_t*>(b);
    s->peers.recv[p] = reinterpret_cast<float*>(b + kP2pFlagBytesAligned);
    s->peers.result[p] = s->peers.recv[p] + s->cap;
  }
  s->rank = rank; s->world = world;
  return 0;
}

int nf_dp_p2p_enable(int32_t on) {
  P2pState* s = p2p_current();
  NF_CHECK_ARG(s != nullptr, NF_E_BADARG, "no current CUDA device");
  NF_CHECK_ARG(!on || (s->base && s->world >= 2), NF_E_BADARG, "nf_dp_p2p_connect has not succeeded on this device");
  s->connected = on != 0;
  return 0;
}

int nf_dp_p2p_status(void) {
  P2pState* s = p2p_current();
  if (!s || !s->connected) return 0;
  int err = 0;
  const char* e = static_cast<const char*>(s->base) + (2 * kP2pMaxWorld * kP2pCtas + kP2pCtas) * sizeof(uint32_t);
  if (cudaMemcpy(&err, e, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  return err ? -1 : 1;
}

int nf_dp_p2p_timeline(uint64_t* ns6) {
  NF_CHECK_ARG(ns6 != nullptr, NF_E_BADARG, "ns6 is NULL");
  P2pState* s = p2p_current();
  NF_CHECK_ARG(s != nullptr && s->base != nullptr, NF_E_BADARG, "nf_dp_p2p_alloc has not been called on this device");
  const char* e = static_cast<const char*>(s->base) + (2 * kP2pMaxWorld * kP2pCtas + kP2pCtas + 2) * sizeof(uint32_t);
  NF_CUDA(cudaMemcpy(ns6, e, 6 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
  return 0;
}

int nf_dp_shutdown(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  DevComm* c = current();
  if (c && c->comm) {
    g_api.comm_destroy(c->comm);
    c->comm = nullptr; c->world = 0;
  }
  p2p_release(p2p_current());
  return 0;
}

}  // extern "C"
