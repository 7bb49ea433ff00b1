// Precision dispatch of one Linear / one weight gradient (defined in flow.cu): tcgen05 3xTF32 when the shape and
// the operand pitches allow it, exact-fp32 FFMA otherwise.
#pragma once
#include "tc_gemm.cuh"

namespace nf {

// C[z] (M,N) (+)= [A0 | A1][z] (M,K0+K1) @ [B0 | B1][z] (N,K0+K1)^T (+ bias)
int linear_nt(int prec, const TcGemmP& t, cudaStream_t st);
// C[z] (M,N) += A[z] (K,M)^T @ B[z] (K,N); C must hold the base value
int wgrad_tn(int prec, const TcGemmTnP& t, cudaStream_t st);

// true when wgrad_tn(prec, t) runs a kernel whose splitter threads can take t.colsum / t.skinny along (tc_gemm.cuh)
bool rides_ok(int prec, const TcGemmTnP& t);

}  // namespace nf
