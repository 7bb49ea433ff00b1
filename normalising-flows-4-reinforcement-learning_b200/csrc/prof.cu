#include "prof.cuh"

#include <mutex>
#include <vector>

#include "common.cuh"   // pulls in include/nf_b200.h (NF_API declarations)

namespace nf {
namespace {
struct Rec { cudaEvent_t e0, e1; int kind; double work; };
std::mutex g_mu;
bool g_on = false;
std::vector<Rec> g_recs;
std::vector<cudaEvent_t> g_pool;

cudaEvent_t get_event() {
  if (!g_pool.empty()) { cudaEvent_t e = g_pool.back(); g_pool.pop_back(); return e; }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}
}  // namespace

bool prof_enabled() { return g_on; }

void prof_begin(cudaStream_t st, int kind, double work) {
  std::lock_guard<std::mutex> lk(g_mu);
  Rec r{get_event(), get_event(), kind, work};
  cudaEventRecord(r.e0, st);
  g_recs.push_back(r);
}

void prof_end(cudaStream_t st) {
  std::lock_guard<std::mutex> lk(g_mu);
  // scopes do not nest across kinds on one thread; the last open record of this stream is ours
  if (!g_recs.empty()) cudaEventRecord(g_recs.back().e1, st);
}

}  // namespace nf

using namespace nf;

extern "C" {

int nf_prof_enable(int on) {
  std::lock_guard<std::mutex> lk(g_mu);
  g_on = on != 0;
  return 0;
}

// Sums (and clears) the records: per kind total milliseconds, total work, launch count.
int nf_prof_collect(double* ms, double* work, int64_t* launches, int nkinds) {
  std::lock_guard<std::mutex> lk(g_mu);
  for (int k = 0; k < nkinds; ++k) { ms[k] = 0; work[k] = 0; launches[k] = 0; }
  for (Rec& r : g_recs) {
    cudaEventSynchronize(r.e1);
    float t = 0.f;
    if (cudaEventElapsedTime(&t, r.e0, r.e1) == cudaSuccess && r.kind < nkinds) {
      ms[r.kind] += t; work[r.kind] += r.work; launches[r.kind] += 1;
    }
    g_pool.push_back(r.e0); g_pool.push_back(r.e1);
  }
  g_recs.clear();
  return 0;
}

}  // extern "C"
