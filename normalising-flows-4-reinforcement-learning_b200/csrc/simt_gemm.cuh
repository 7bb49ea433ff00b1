// Exact-fp32 FFMA GEMM used for the small / ragged contractions of the flow (PLU matmuls,
// last conditioner layer, PLU gradients) and as the on-device cross-check of the tcgen05 path.
#pragma once
#include "common.cuh"

namespace nf {

// C[m,n] (+)= alpha * sum_k opA(m,k) * opB(k,n)   (+ bias[n])
// The reduction dimension may be the concatenation of two segments (K0 then K1), each with its
// own A and B pointer -- this is how [x_cond | y] @ W1^T is done without materialising the
// concatenation (torch_fcnf.py:99-100).
struct GemmP {
  const float* A0 = nullptr; const float* A1 = nullptr;
  const float* B0 = nullptr; const float* B1 = nullptr;
  float* C = nullptr;
  const float* bias = nullptr;
  int M = 0, N = 0, K0 = 0, K1 = 0;
  int lda0 = 0, lda1 = 0, ldb0 = 0, ldb1 = 0, ldc = 0;
  long long sA0 = 0, sA1 = 0, sB0 = 0, sB1 = 0, sC = 0, sBias = 0;  // batch strides (floats)
  int batch = 1;
  int transA = 0;    // 0: A(m,k) = A[m*lda + k]   1: A(m,k) = A[k*lda + m]
  int transB = 0;    // 0: B(k,n) = B[k*ldb + n]   1: B(k,n) = B[n*ldb + k]
  int accumulate = 0;  // C += result instead of C = result
  int splitk = 1;      // > 1: partial sums combined with atomicAdd (C must hold the base value)
  float alpha = 1.0f;
};

int simt_gemm(const GemmP& p, cudaStream_t st);

// picks a split-K factor that fills the GPU for a batch-reduction GEMM
int pick_splitk(int M, int N, int64_t K, int batch);

}  // namespace nf
