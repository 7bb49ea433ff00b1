// Optional per-kernel-class timing with CUDA events on the launching stream (off by default).
// bench.py turns it on for a separate pass to measure the dominant kernel's average launch duration
// inside the real step (the roofline numerator); the timed throughput pass runs with it off.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace nf {

enum ProfKind {
  PROF_TC_GEMM = 0,     // tcgen05 K-major GEMM (Linear forward / data gradients)
  PROF_TC_WGRAD = 1,    // tcgen05 MN-major GEMM (weight gradients)
  PROF_SIMT_GEMM = 2,   // FFMA GEMM (small / ragged contractions)
  PROF_LN_FWD = 3,      // LeakyReLU + LayerNorm forward
  PROF_LN_BWD = 4,      // LayerNorm + LeakyReLU backward
  PROF_AFFINE = 5,      // affine coupling + log-det (fwd / inv / bwd)
  PROF_PARAM = 6,       // parameter-sized kernels (PLU build, folding, transposes, unfolding)
  PROF_ROW_TAIL = 7,    // row-fused forward / inverse tail
  PROF_ROW_HEAD = 8,    // row-fused backward head
  PROF_ROW_BTAIL = 9,   // row-fused backward tail
  PROF_COL_REDUCE = 10, // multi-job column reduction over the batch rows (skinny weight/bias gradients)
  PROF_TF_ROW = 11,     // transformer flow: row kernels (embedding, residual + LayerNorm, GELU, tail, column sums); work = bytes
  PROF_TF_ATTN = 12,    // transformer flow: causal attention forward / backward; work = FLOPs
  PROF_NKINDS = 13
};

bool prof_enabled();
void prof_begin(cudaStream_t st, int kind, double work);   // work = algorithmic FLOPs or bytes
void prof_end(cudaStream_t st);

struct ProfScope {
  cudaStream_t st;
  bool on;
  ProfScope(cudaStream_t s, int kind, double work) : st(s), on(prof_enabled()) {
    if (on) prof_begin(st, kind, work);
  }
  ~ProfScope() {
    if (on) prof_end(st);
  }
};

}  // namespace nf
