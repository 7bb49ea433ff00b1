// Data-parallel communicator owned by the library (see dp.cu).
#pragma once
#include "common.cuh"

namespace nf {

int dp_world();                                                          // 0: nf_dp_init not called on this device
int dp_allreduce(float* buf, int64_t n, bool average, cudaStream_t st);  // in place; capturable into a CUDA graph

}  // namespace nf
