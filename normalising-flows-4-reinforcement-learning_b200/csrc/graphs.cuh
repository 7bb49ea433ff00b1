// Launch-sequence replay: every C-ABI entry point of the flow (prepare / forward / inverse / backward)
// is a fixed sequence of 10-500 small kernels that depends only on the descriptor, the batch size
// and the caller's pointers.  At the reference's batch sizes (256-4096) the host cannot issue those
// launches as fast as the B200 retires them, so the sequence is captured once into a CUDA graph
// (stream capture on a private stream), cached under (entry point, descriptor, B, pointers) and
// replayed with one cudaGraphLaunch on the caller's stream whenever the same arguments come back
// -- which is what a training or sampling loop over a caching allocator does.  A key is captured on
// its second sighting (one-off calls stay eager); the cache is a small LRU.
#pragma once
#include <functional>

#include "common.cuh"

namespace nf {

struct GraphKey {
  int op = 0;                 // entry point
  NfFlowDesc desc{};
  long long B = 0;
  long long ws_bytes = 0;
  int flags = 0;
  const void* ptr[12] = {};
};

// Runs body(stream) either eagerly on `user` or as a cached graph launched on `user`.
int run_graphed(const GraphKey& key, cudaStream_t user, const std::function<int(cudaStream_t)>& body);

}  // namespace nf
