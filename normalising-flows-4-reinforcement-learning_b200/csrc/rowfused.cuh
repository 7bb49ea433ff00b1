// Row-fused kernels for small flow dimensions (D <= 32): everything in a block that is NOT one of
// the wide conditioner GEMMs -- the last Linear (N = D/2 outputs), the affine coupling, the
// log-det row reduction, the invertible PLU matmul of the neighbouring block and, in backward,
// the matching gradient pieces -- runs as one warp-per-row kernel instead of five to eight
// launches of skinny GEMMs and element-wise kernels.
#pragma once
#include "common.cuh"

namespace nf {

constexpr int kRowFusedMaxD = 32;

// forward / inverse tail of block i
struct TailP {
  const float* xh2; long long xh2_net;   // (2, B, C) normalised activations of layer 2 (net 0 = t, 1 = s)
  const float* W3f; long long w_net;     // (2, dt, C) folded last Linear
  const float* b3f; long long b_net;     // (2, dt)
  const float* in; int ld_in;            // forward: xp = x W (B, D) ; inverse: current z (B, D)
  const float* cdev;                     // sum log|s_plu| of this block (device scalar)
  const float* Wmat;                     // forward: W of block i+1 (or null) ; inverse: W^-1 of this block
  float* out; int ld_out;                // forward: block output z (B, D) ; inverse: unused (null)
  float* out2; int ld_out2;              // forward: xp of block i+1 (if Wmat) ; inverse: x = xp W^-1
  float* logdet;                         // (B,) accumulated, may be null
  float* s_out;                          // (B, dt) log-scales for backward, may be null
  long long B; int D, C, inverse;
};
int row_tail(const TailP& p, cudaStream_t st);

// backward head of block i: affine backward, last-Linear gradients, data gradient through the
// last Linear and through LayerNorm + LeakyReLU of layer 2
struct BwdHeadP {
  const float* dz; int ld_dz;            // incoming gradient (B, D), may be null (zero)
  const float* z;                        // block output (B, D) dense
  const float* s;                        // (B, dt) stashed log-scales
  const float* dld;                      // (B,) may be null
  const float* xh2; long long xh2_net;   // (2, B, C)
  const float* rstd; long long rstd_net; // (2, B)
  const uint32_t* mask; long long mask_net; int mw;
  const float* W3f; long long w_net;     // (2, dt, C)
  float* dxp;                            // (B, D) out: [dz_c, dz_t e^-s]
  float* du2; long long du2_net;         // (2, B, C) out
  float* G3; long long g_net;            // (2, dt, C) += dout^T x_hat2      (null: no parameter grads)
  float* g3;                             // (2, dt) += colsum(dout), stride g_net
  float* gb2;                            // (2, C) += colsum(du2), stride g_net
  long long B; int D, C;
};
int row_bwd_head(const BwdHeadP& p, cudaStream_t st);
size_t row_bwd_head_smem(int D, int C);

// backward tail of block i: x_cond part of the first Linear's data gradient, PLU backward
struct BwdTailP {
  const float* du1; long long du1_net;   // (2, B, H1)
  const float* W1T; long long w_net;     // (2, K1, H1): rows 0..dc-1 are the x_cond rows of W1^T
  const float* z;                        // block output (B, D): x_cond = z[:, :dc]
  const float* xin;                      // block input (B, D)
  const float* W;                        // (D, D) of this block
  float* dxp;                            // (B, D) in/out: first dc columns get += dh0
  float* dxin; int ld_dxin;              // (B, D) out: dxp W^T   (may be null)
  float* G1; long long g_net; int ldg;   // (2, H1, K1) first dc columns += du1^T x_cond   (null: skip)
  float* dW;                             // (D, D) += x_in^T dxp   (null: skip)
  long long B; int D, H1;
};
int row_bwd_tail(const BwdTailP& p, cudaStream_t st);
size_t row_bwd_tail_smem(int D, int H1);

constexpr size_t kRowFusedSmemLimit = 96 * 1024;   // keeps >= 2 CTAs per SM; wider flows use the generic path

}  // namespace nf
