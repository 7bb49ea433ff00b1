// Row-wise and parameter-sized kernels of the flow (everything that is not a GEMM).
#pragma once
#include "common.cuh"

namespace nf {

// x_hat = LayerNorm_noaffine(LeakyReLU(u + bias)); optional rstd (B), sign mask (B, mw words).
// Batched over `nets` with the given strides (floats / words).
struct LnFwdP {
  const float* u; long long u_net;
  const float* bias; long long bias_net;
  float* xh; long long xh_net;
  float* xh_lo; long long xh_lo_net;   // optional low plane (tf32 split); may be null
  float* rstd; long long rstd_net;     // may be null
  uint32_t* mask; long long mask_net;  // may be null
  int mw, N, nets;
  long long B;
  float eps;
  int fast_var;   // flax nn.LayerNorm: var = max(0, E[x^2] - E[x]^2) instead of the centred two-pass variance
};
int lrelu_ln_fwd(const LnFwdP& p, cudaStream_t st);

// g (in: d x_hat, out: d u) ; colsum[n] += sum_rows du
struct LnBwdP {
  float* g; long long g_net;
  const float* xh; long long xh_net;
  const float* rstd; long long rstd_net;
  const uint32_t* mask; long long mask_net;
  float* colsum; long long colsum_net;   // N floats per net, atomically accumulated
  int mw, N, nets;
  long long B;
};
int ln_lrelu_bwd(const LnBwdP& p, cudaStream_t st);

// ld_in / ld_out: row pitches (floats) of the (B,D) input and output
int affine_fwd(const float* xp, int ld_in, const float* s, const float* t, const float* bias_s, const float* bias_t,
               const float* logdet_const_dev, float logdet_const_host, long long B, int D,
               float* z, int ld_out, float* logdet, float* s_out, cudaStream_t st);
int affine_inv(const float* z, int ld_in, const float* s, const float* t, const float* bias_s, const float* bias_t,
               const float* logdet_const_dev, float logdet_const_host, long long B, int D,
               float* xp, int ld_out, float* logdet, cudaStream_t st);
// ds/dt are (B,dt) each; g3s/g3t (dt) receive column sums of ds/dt (atomic), may be null.
int affine_bwd(const float* dz, const float* z, const float* s_full, const float* dlogdet, long long B, int D,
               float* dxp, float* ds, float* dt, float* g3s, float* g3t, cudaStream_t st, int inverse = 0);

// ---- parameter-sized kernels ---------------------------------------------------------------
int plu_mask(const float* params, const NfParamLayout& pl, const Dims& m, float* packed,
             const PackedLayout& kl, cudaStream_t st);
// D <= 32: mask + PL + W + W^T + log|det| of every block in one kernel; the PLU parameter gradients (the three small GEMMs
// and plu_grad_finish) in one kernel
int plu_build_small(const float* params, const NfParamLayout& pl, const Dims& m, float* packed, const PackedLayout& kl,
                    cudaStream_t st);
// D <= 32: every derived table of the forward direction (PLU tables, folded LayerNorm affine and its transpose, first-layer
// packings, total log|det|) with their hi / lo copies, in one kernel with no dependency between CTAs
int prepare_fwd_fused(const float* params, const NfParamLayout& pl, const Dims& m, float* packed, const PackedLayout& kl,
                      cudaStream_t st);
int plu_grad_small(const float* params, const NfParamLayout& pl, const Dims& m, const float* packed, const PackedLayout& kl,
                   const float* dW, const float* sum_dld, float* dparams, cudaStream_t st, float ld_sign = 1.0f);
int plu_tri_inverse(const float* packed, const PackedLayout& kl, const Dims& m, float* uinv, float* linv,
                    cudaStream_t st);
int fold_ln_affine(const float* params, const NfParamLayout& pl, const Dims& m, float* packed,
                   const PackedLayout& kl, cudaStream_t st);
int pack_w1(const float* params, const NfParamLayout& pl, const Dims& m, float* packed, const PackedLayout& kl,
            cudaStream_t st);
int unfold_grads(const float* params, const NfParamLayout& pl, const Dims& m, float* dparams, cudaStream_t st);
int plu_grad_finish(const float* params, const NfParamLayout& pl, const Dims& m, const float* dLh,
                    const float* dUh, const float* sum_dld, float* dparams, cudaStream_t st, float ld_sign = 1.0f);
int sum_vector(const float* v, long long n, float* out, cudaStream_t st);  // out[0] = sum(v)
int fill_zero(float* p, long long n, cudaStream_t st);
// hi[i] = rna_tf32(t[i]), lo[i] = t[i] - hi[i] for the first n floats of the packed tables (hi = t + hi_off, ...)
int split_tables(float* tables, long long hi_off, long long lo_off, long long n, cudaStream_t st);
// ---- the steps either side of the flow in the training loop -----------------------
// loss[0] = -mean_b( -0.5 |z_b|^2 - D/2 log(2 pi) + logdet_b )      (torch_nfbc_ant.py:269 with a standard-normal prior)
int nll_forward(const float* z, const float* logdet, long long B, int D, float* loss, cudaStream_t st);
// cotangents of that loss scaled by the upstream gradient g[0] (g may be null = 1): dz = g z / B, dld = -g / B
int nll_cotangents(const float* z, const float* g, long long B, int D, float* dz, float* dld, cudaStream_t st);
// one AdamW step (torch.optim.AdamW semantics, torch_nfbc_ant.py:150-155) over the flat parameter arena; slot_mask
// (n_blocks * NF_P_SLOTS bytes on the device) selects the trainable slots
// lp[b] = log N(z_b; 0, I) and its backward dz = -g[b] z   (the prior term of the reference's loss line)
int stdnormal_logprob(const float* z, long long B, int D, float* lp, cudaStream_t st);
int stdnormal_logprob_bwd(const float* z, const float* g, long long B, int D, float* dz, cudaStream_t st);
// per group g of N candidate rows: argmin_n of -0.5 |z|^2 - D/2 log 2 pi + logdet (ties: smaller index); out_x (G, D) =
// the winning row of x, out_idx (G) its index, out_lp (G) its log-probability (each optional)
int group_argmin_logp(const float* z, const float* logdet, const float* x, long long G, long long N, int D, float* out_x,
                      long long* out_idx, float* out_lp, cudaStream_t st);
int noise_augment(const float* x, long long n, float std_, unsigned long long seed, unsigned long long offset, float* out,
                  cudaStream_t st);   // out = x + std * N(0, 1), Philox4x32-10 keyed by (seed, offset + i / 4)
struct AdamWP {
  float* params; const float* grads; float* exp_avg; float* exp_avg_sq;
  const unsigned char* slot_mask;
  float lr_wd_factor;      // 1 - lr * weight_decay
  float one_minus_beta1, beta2, one_minus_beta2, step_size, bc2_sqrt, eps;
};
int adamw_step(const AdamWP& p, const NfParamLayout& pl, int n_blocks, cudaStream_t st);

// up to 4 contiguous device-to-device copies in one launch (the I/O staging around the cached launch graphs)
struct CopySegs { float* dst[4]; const float* src[4]; long long n[4]; int count = 0;
                  void add(float* d, const float* s, long long k) { if (d && s && k > 0 && d != s) { dst[count] = d; src[count] = s; n[count] = k; ++count; } } };
int copy_segments(const CopySegs& c, cudaStream_t st);
int axpy_rows(float* dst, const float* src, long long n, cudaStream_t st);  // dst += src
int copy2d(float* dst, long long ldd, const float* src, long long lds, long long rows, int cols, cudaStream_t st);
// dst[b][c][r] = src[b][r][c]  (rows x cols -> cols x rows), batched with strides in floats
int transpose2d(float* dst, long long ldd, long long sdst0, long long sdst1, const float* src, long long lds,
                long long ssrc0, long long ssrc1, int rows, int cols, int batch0, int batch1, cudaStream_t st);

}  // namespace nf
