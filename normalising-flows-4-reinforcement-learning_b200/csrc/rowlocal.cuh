// Row-local backward kernels for small flow dimensions (D <= 32, C and H1 multiples of 4 up to 1024).
//
// The first generation (rowfused.cu: k_row_bwd_head / k_row_bwd_tail) accumulated the skinny
// weight/bias gradients across rows inside the row kernel (register accumulators, shared-memory
// reduction, atomics), which made a warp walk several rows twice and left the kernel latency-bound
// (ncu: profiles/r1_early_row_bwd_head.txt, 26 % SM throughput, 23 % warps active).  Here every
// row is touched once by one warp with 128-bit accesses and NO cross-row state; all reductions over
// the batch rows are done by one multi-job column-reduction kernel per block (col_reduce).
#pragma once
#include "common.cuh"

namespace nf {

// affine backward + data gradient through the last Linear and LayerNorm-2/LeakyReLU-2, per row:
//   dxp = [dz_c, dz_t e^-s]; dt = -dz_t e^-s; ds = -dz_t z_t - dld         (torch_fcnf.py:101-104 transposed)
//   g = dout W3f;  du2 = LeakyReLU'(mask) rstd (g - mean(g) - x_hat2 mean(g x_hat2))   (:75-78 transposed)
struct BwdHead2P {
  const float* dz; int ld_dz;            // (B, D) incoming gradient, may be null (zero)
  const float* z;                        // (B, D) block output
  const float* s;                        // (B, dt) stashed log-scales
  const float* dld;                      // (B,) may be null
  const float* xh2; long long xh2_net;   // (2, B, C)
  const float* rstd; long long rstd_net; // (2, B)
  const uint32_t* mask; long long mask_net; int mw;
  const float* W3f; long long w_net;     // (2, dt, C)
  float* dxp;                            // (B, D) out
  float* dout; long long dout_net;       // (2, B, dt) out: net 0 = dt, net 1 = ds
  float* du2; long long du2_net;         // (2, B, C) out
  long long B; int D, C;
};
int row_bwd_head2(const BwdHead2P& p, cudaStream_t st);

// x_cond part of the first Linear's data gradient and the PLU matmul backward, per row:
//   dxp[:, :dc] += du1_t W1_t[:, :dc] + du1_s W1_s[:, :dc];  dx_in = dxp W^T     (torch_fcnf.py:98-100, :40 transposed)
struct BwdTail2P {
  const float* du1; long long du1_net;   // (2, B, H1)
  const float* W1T; long long w_net;     // (2, K1, H1): rows 0..dc-1 are the x_cond rows of W1^T
  const float* W;                        // (D, D) of this block
  float* dxp;                            // (B, D) in/out
  float* dxin; int ld_dxin;              // (B, D) out, may be null
  long long B; int D, H1;
};
int row_bwd_tail2(const BwdTail2P& p, cudaStream_t st);

bool row_v2_supported(int D, int C, int H1);

struct TailP;   // rowfused.cuh
// forward / inverse tail with the same contract as row_tail (rowfused.cuh), C % 4 == 0, C <= 1024
int row_tail2(const TailP& p, cudaStream_t st);

// Column reductions over the batch rows, several jobs in one launch:
//   out[z][i][c]  += sum_rows coef[z][row][i] * A[z][row][c]     (i < m <= 32; element strides so_i / so_c)
//   out2[z][c]    += sum_rows A2[z][row][c]                       (optional)
//   out3[z][i]    += sum_rows coef[z][row][i]                     (optional)
// z runs over `batch` (the two nets) with the given strides (0 = shared).
struct ColJob {
  const float* A = nullptr; long long lda = 0, sA = 0;
  const float* coef = nullptr; long long ldcf = 0, sCf = 0; int m = 0;
  float* out = nullptr; long long so_i = 0, so_c = 1, sOut = 0;
  const float* A2 = nullptr; long long lda2 = 0, sA2 = 0;
  float* out2 = nullptr; long long sOut2 = 0;
  float* out3 = nullptr; long long sOut3 = 0;
  int N = 0, batch = 1;
};
constexpr int kColMaxJobs = 8;
int col_reduce(const ColJob* jobs, int njobs, long long B, cudaStream_t st);
// dW (D, D) += X (B, D)^T G (B, D), dense rows, D <= 32
// (batch > 1: one launch for several flow blocks, operands `s?` floats apart)
int xtg_small(const float* X, const float* G, float* dW, long long B, int D, cudaStream_t st, int batch = 1, long long sX = 0,
              long long sG = 0, long long sW = 0);

}  // namespace nf
