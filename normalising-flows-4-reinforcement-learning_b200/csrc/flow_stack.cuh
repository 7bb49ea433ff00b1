// Persistent whole-stack kernel of the coupling flow (forward and inverse), sm_100a.
//
// Rows are independent through the whole stack (torch_fcnf.py:134-141), so a cluster of CTAs that owns a
// 128-row tile can run ALL n blocks without leaving the kernel: per block
//     u1 = [x_cond | y] W1^T -> LeakyReLU -> LayerNorm -> x_hat1      (phase 0, tcgen05 + cluster LN exchange)
//     u2 = x_hat1 W2f^T      -> LeakyReLU -> LayerNorm -> x_hat2      (phase 1)
//     s, t = x_hat2 W3f^T ; affine coupling ; log-det ; PLU matmul of the neighbouring block   (tail, leader CTA)
// (torch_fcnf.py:95-106 forward, :108-117 inverse).  The cluster is (1, H/128, 2): CTA (j, net) owns columns
// [128 j, 128 j + 128) of net `net` (0 = t, 1 = s).  Between phases the activations travel through global memory
// (L2-resident: the x_hat1 tile a CTA wrote is read back by the ncl CTAs of its net with TMA) behind a cluster
// barrier; nothing is launched between blocks.  One launch replaces 3 n launches of the per-layer path.
#pragma once
#include "common.cuh"

namespace nf {

struct StackP {
  // problem
  long long B = 0;
  int D = 0, H = 0, R = 0, n = 0;      // H = H1 = C (the kernel needs equal hidden widths)
  int inverse = 0;
  float eps = 1e-5f;
  int fast_var = 0;     // flax LayerNorm variance
  int raw_hi = 1;
  // operands (all fp32, 16-byte aligned)
  const float* cur = nullptr; int ld_cur = 0;       // (B, ld_cur) running buffer: forward x W_0 on entry; inverse z on entry
  float* cur_w = nullptr;                           //   (same buffer, written in place by the tail)
  const float* y = nullptr;                         // (B, R)
  // first Linear: W1p (H, K1p) per net / block, y columns at column dcp; hi / lo copies at +hi_off / +lo_off
  const float* W1p = nullptr; int K1p = 0, dcp = 0;
  const float* W2f = nullptr;                       // (H, H) per net / block
  long long w_net = 0, w_blk = 0, hi_off = 0, lo_off = 0;   // packed-table strides (floats): net, block, hi / lo copies
  const float* b1 = nullptr; long long b1_net = 0, b1_blk = 0;     // params: first bias
  const float* b2f = nullptr;                       // packed: folded second bias (strides w_net / w_blk)
  const float* W3f = nullptr; const float* b3f = nullptr;   // packed: folded last Linear (dt, H) and bias (strides w_net / w_blk)
  const float* Wmat = nullptr;                      // packed: W (forward) or W^-1 (inverse) of block 0 (stride w_blk)
  const float* cdev = nullptr;                      // (n) sum log|s_plu| per block
  // activations between the phases: (2, B, H) per block slot
  float* xh1 = nullptr; float* xh2 = nullptr; long long xh_net = 0, xh_blk = 0;   // xh_blk = 0: one slot reused by every block
  int store_xh2 = 0;
  float* rstd1 = nullptr; float* rstd2 = nullptr;   // (B) per net / block, strides xh_net / xh_blk; null: not stored
  uint32_t* mask1 = nullptr; uint32_t* mask2 = nullptr; int mw = 0;
  // outputs
  float* out = nullptr; long long out_blk = 0;      // forward: block outputs (B, D): out + blk * out_blk; out_blk = 0: last block only
  float* s_out = nullptr; long long s_blk = 0;      // forward: log-scales (B, dt) per block, may be null
  float* logdet = nullptr;                          // (B) accumulated (read-modify-write), may be null
  float* final_out = nullptr; int ld_final = 0;     // inverse: x of the last step (B, ld_final)
};

bool flow_stack_supported(const StackP& p);
int flow_stack(const StackP& p, cudaStream_t st);

}  // namespace nf
