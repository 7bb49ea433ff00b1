/*
 * nf_b200.h -- C ABI of the B200-native coupling-flow hot path.
 *
 * Drop-in boundary for the reference's flow (there is no native interface in the reference;
 * every entry point names the Python symbol it replaces, paths relative to /root/reference):
 *
 *   nf_flow_forward   <- RealNVP.forward        il-and-cil/utils/torch_fcnf.py:134-141
 *                        (MetaBlock.forward :95-106, InvertiblePLU.forward :33-42)
 *   nf_flow_inverse   <- RealNVP.reverse        il-and-cil/utils/torch_fcnf.py:143-152
 *                        (MetaBlock.reverse :108-117, InvertiblePLU.reverse :44-58;
 *                         optional log-det as unsupservised-gcrl/nf.py:131-148)
 *   nf_flow_backward  <- torch.autograd through the above (loss.backward(),
 *                        il-and-cil/torch_nfbc_ant.py:269-272; input gradient for the denoise
 *                        step :212-217)
 *   nf_flow_prepare   <- the per-call parameter work the reference redoes inside every
 *                        forward/reverse: W = P L U (:34-38), W^-1 by two triangular solves
 *                        (:45-54), sum log|s| (:41)
 *   nf_affine_logdet_{fwd,inv,bwd} <- the affine coupling + log-det lines :101-104 / :112-113
 *                        as a standalone HBM-bound kernel
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no torch / ATen types.
 *   - Every pointer is a DEVICE pointer to caller-owned memory (fp32, row-major, contiguous,
 *     16-byte aligned) borrowed for the duration of the call; the library allocates nothing on
 *     the data path.  Scratch and stash buffers are sized by the *_bytes queries.
 *   - All work is enqueued asynchronously on `stream` (a cudaStream_t passed as void*).
 *   - Return value: 0 = ok, < 0 = bad argument (NF_E_*), > 0 = cudaError_t.  Nothing throws
 *     across the ABI; nf_last_error() returns a thread-local message for the last failure.
 *   - Re-entrant; one process per GPU.
 */
#ifndef NF_B200_H
#define NF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NF_ABI_VERSION 1

#if defined(__GNUC__)
#define NF_API __attribute__((visibility("default")))
#else
#define NF_API
#endif

#define NF_E_BADARG   (-1)
#define NF_E_ALIGN    (-2)
#define NF_E_WORKSPACE (-3)
#define NF_E_UNSUPPORTED (-4)

/* Arithmetic of the conditioner GEMMs. */
#define NF_PREC_FP32_SIMT 0 /* FFMA, exact fp32 (checker / odd shapes)                         */
#define NF_PREC_TF32X3    1 /* tcgen05 kind::tf32, 3-pass split operands, fp32 accumulate:     */
                            /* fp32-parity mode (<= 1e-5 vs the fp64 arbiter)                   */
#define NF_PREC_BF16      2 /* reserved: not built, rejected with NF_E_UNSUPPORTED                */
#define NF_PREC_TF32      3 /* optional reduced-precision mode: ONE kind::tf32 pass on the raw fp32 operands (10-bit
                             * mantissa, fp32 accumulate), no operand splitting; contract 2e-2 relative on z / log_prob
                             * for D <= 64, n_blocks <= 12 (BASELINE north_star's optional fast mode; errors grow with
                             * depth: measured values in DESIGN.md) */

typedef struct NfFlowDesc {
  int32_t D;        /* in_channels: flow dimension                                     */
  int32_t C;        /* channels: conditioner width (second hidden layer and LN size)    */
  int32_t H1;       /* first hidden width: C (torch_fcnf.py:75) or C+R (nf.py:94)       */
  int32_t R;        /* cond_channels                                                    */
  int32_t n_blocks; /* n_layers                                                         */
  float   ln_eps;   /* 1e-5 (torch nn.LayerNorm) or 1e-6 (flax nn.LayerNorm)            */
  int32_t precision;/* NF_PREC_*                                                        */
  int32_t reserved; /* flags: NF_FLAG_*                                                 */
} NfFlowDesc;
/* flax nn.LayerNorm (nf.py:94-104) computes var = max(0, E[x^2] - E[x]^2); torch nn.LayerNorm the centred variance */
#define NF_FLAG_LN_FAST_VARIANCE 1

/* Per-block slots of the flat parameter arena, in order.  Offsets are in floats, every slot
 * starts on a multiple of 4 floats.  Net 0 is `t`, net 1 is `s` (state_dict order). */
enum {
  NF_P_S = 0, NF_P_U, NF_P_L, NF_P_P, NF_P_PINV,                 /* blocks.i.l.{s,U,L,P,P_inv} */
  NF_P_W1, NF_P_B1, NF_P_G1, NF_P_BE1, NF_P_W2, NF_P_B2, NF_P_G2, NF_P_BE2, NF_P_W3, NF_P_B3,
  NF_P_NET_SLOTS = 10,                                            /* {0,2,3,5,6}.{weight,bias}  */
  NF_P_SLOTS = 5 + 2 * 10
};

typedef struct NfParamLayout {
  int64_t slot_off[NF_P_SLOTS];  /* offset of each slot inside one block (net t then net s) */
  int64_t slot_numel[NF_P_SLOTS];
  int64_t block_stride;          /* floats per block                                         */
  int64_t net_stride;            /* offset(net s slot) - offset(net t slot)                  */
  int64_t total;                 /* floats in the whole arena = n_blocks * block_stride      */
} NfParamLayout;

NF_API int nf_abi_version(void);
NF_API const char* nf_last_error(void);

/* Layout of the flat parameter / gradient arena for `desc` (pure host computation). */
NF_API int nf_param_layout(const NfFlowDesc* desc, NfParamLayout* out);

/* Buffer sizes (bytes). */
NF_API int64_t nf_flow_packed_bytes(const NfFlowDesc* desc);
NF_API int64_t nf_flow_stash_bytes(const NfFlowDesc* desc, int64_t B);
NF_API int64_t nf_flow_workspace_bytes(const NfFlowDesc* desc, int64_t B);

/* what: bit 0 = forward/backward tables, bit 1 = inverse tables (W^-1). */
#define NF_PREP_FWD 1
#define NF_PREP_INV 2
NF_API int nf_flow_prepare(const NfFlowDesc* desc, const float* params, void* packed, int what,
                    void* workspace, int64_t ws_bytes, void* stream);

/* y == NULL in nf_flow_forward / nf_flow_inverse / nf_flow_backward means "the conditioning input is identically zero"
 * (train_auto_NF_rl.py:328-339 evaluates the flow 1 M times with cond = 0): the y segment of the first Linear's
 * reduction, the dy GEMM and the y columns of dW1 are skipped (dy, if given, is zero-filled).
 *
 * z (B,D), logdet (B,), outputs (n_blocks,B,D) or NULL.  `stash` (nf_flow_stash_bytes) receives
 * what nf_flow_backward needs; pass NULL for an inference-only forward. */
NF_API int nf_flow_forward(const NfFlowDesc* desc, const float* params, const void* packed,
                    const float* x, const float* y, int64_t B,
                    float* z, float* logdet, float* outputs, void* stash,
                    void* workspace, int64_t ws_bytes, void* stream);

/* x (B,D) from z (B,D); logdet (B,) or NULL (flax-style reverse log-det);
 * seq (n_blocks+1,B,D) or NULL (return_sequence=True, seq[0] = z). */
NF_API int nf_flow_inverse(const NfFlowDesc* desc, const float* params, const void* packed,
                    const float* z, const float* y, int64_t B,
                    float* x, float* logdet, float* seq,
                    void* workspace, int64_t ws_bytes, void* stream);

/* Given dz (B,D) and dlogdet (B,) (either may be NULL = zero):  dx (B,D) or NULL, dy (B,R) or
 * NULL, dparams (flat arena, layout of nf_param_layout, fully overwritten) or NULL.  `y` must be the
 * tensor given to nf_flow_forward; the kernels read the copy forward left in the stash.
 * All entry points copy the caller's inputs into / outputs out of the workspace and stash with stream-ordered
 * copies around the kernel sequence: the kernels themselves only touch params, packed, stash and workspace,
 * so keeping THOSE buffers persistent is what makes the cached launch graphs replay (see below). */
NF_API int nf_flow_backward(const NfFlowDesc* desc, const float* params, const void* packed,
                     const void* stash, const float* y, int64_t B,
                     const float* dz, const float* dlogdet,
                     float* dx, float* dy, float* dparams,
                     void* workspace, int64_t ws_bytes, void* stream);

/* Backward of the REVERSE direction x = nf_flow_inverse(z, y) (with its log-det), for callers that differentiate
 * through sampling (flax nf.py:131-148 as used by offline-rl/agents/vinf.py:50,68).  `stash` is the one
 * nf_flow_forward(x = the reverse output, y) leaves (the intermediate values of reverse(z) are those of forward(x)).
 * dx (B,D) / dlogdet (B,) are the cotangents of the reverse output and of the reverse log-det (either may be NULL);
 * outputs: dz (B,D) or NULL, dy (B,R) or NULL, dparams (flat arena, fully overwritten) or NULL. */
NF_API int nf_flow_inverse_backward(const NfFlowDesc* desc, const float* params, const void* packed,
                             const void* stash, const float* y, int64_t B,
                             const float* dx, const float* dlogdet,
                             float* dz, float* dy, float* dparams,
                             void* workspace, int64_t ws_bytes, void* stream);

/* The same backward, block range [block_begin, block_end) only (the full pass is the ranges from the top block
 * down, in order, with the same workspace and no other call in between): data-parallel training all-reduces the
 * gradient slice of one range while the next range computes.  dz is read by the range that ends at n_blocks, dx / dy
 * are written by the range that begins at 0, dparams is written for the range's blocks only. */
NF_API int nf_flow_backward_range(const NfFlowDesc* desc, const float* params, const void* packed,
                           const void* stash, const float* y, int64_t B,
                           const float* dz, const float* dlogdet,
                           float* dx, float* dy, float* dparams,
                           int32_t block_begin, int32_t block_end,
                           void* workspace, int64_t ws_bytes, void* stream);

/* ---- data parallel (SURVEY section 8 row e) --------------------------------------------------------------
 * One process per GPU, batch-sharded, full weight replica; the one exchange step is the average of the gradient
 * arena over the ranks.  The library owns an NCCL communicator per device (NCCL is bound at run time with
 * dlopen("libnccl.so.2"), i.e. the build the host process already uses) so that the all-reduce is issued from inside
 * the backward launch sequence and captured into its CUDA graph:
 *   nf_dp_unique_id   rank 0 fills a 128-byte ncclUniqueId; the host distributes it (torch.distributed, MPI, a file)
 *   nf_dp_init        every rank, with its CUDA device current (collective: ncclCommInitRank)
 *   nf_dp_allreduce   in-place sum or average of a device buffer (standalone use: the encoder's arena)
 *   nf_dp_broadcast   in-place broadcast from `root` (initial parameter replica)
 *   nf_dp_p2p_alloc / nf_dp_p2p_connect / nf_dp_p2p_enable / nf_dp_p2p_status
 *                     all-reduce over NVLink peer memory for the ranks of one box (2..8): every rank allocates its
 *                     exchange buffer (`max_floats` = largest arena to take this path) and gets a 64-byte CUDA IPC
 *                     handle; the host gathers the handles of all ranks in rank order and every rank connects; once
 *                     ALL ranks have connected (the host agrees on that), nf_dp_p2p_enable(1) routes nf_dp_allreduce
 *                     and nf_flow_backward_dp of arenas up to max_floats through one peer-store kernel per rank (sum in
 *                     rank order: bit-identical on every replica) instead of NCCL; with fewer than 8 ranks only arenas of
 *                     at least NF_B200_DP_P2P_MIN_BYTES (default 2 MB: below that NCCL's low-latency protocol measured
 *                     faster at 2 and 4 ranks).  status: 1 active, 0 off, -1 a peer did not arrive within the kernel's
 *                     time-out (~5 min).
 *   nf_flow_backward_dp  nf_flow_backward + the averaged gradient arena: blocks are processed from the top down in
 *                     buckets of `bucket_blocks` blocks; each finished bucket (one contiguous slice of the arena) is
 *                     all-reduced on a side stream while the next bucket computes.  dparams must not be NULL. */
NF_API int nf_dp_unique_id(void* id128);
NF_API int nf_dp_init(const void* id128, int32_t rank, int32_t world);
NF_API int nf_dp_world(void);
NF_API int nf_dp_allreduce(float* buf, int64_t n, int32_t average, void* stream);
NF_API int nf_dp_broadcast(float* buf, int64_t n, int32_t root, void* stream);
NF_API int nf_dp_p2p_alloc(int64_t max_floats, void* handle64);
NF_API int nf_dp_p2p_connect(const void* handles, int32_t rank, int32_t world);
NF_API int nf_dp_p2p_enable(int32_t on);
NF_API int nf_dp_p2p_status(void);
NF_API int nf_dp_p2p_timeline(uint64_t* ns6);   /* debug: CTA 0's phase boundaries of the last call, %globaltimer ns */
NF_API int nf_dp_shutdown(void);
NF_API int nf_flow_backward_dp(const NfFlowDesc* desc, const float* params, const void* packed,
                        const void* stash, const float* y, int64_t B,
                        const float* dz, const float* dlogdet,
                        float* dx, float* dy, float* dparams, int32_t bucket_blocks,
                        void* workspace, int64_t ws_bytes, void* stream);

/* Standalone affine coupling + log-det kernels (HBM-bound; torch_fcnf.py:101-104, :112-113).
 *   fwd: z[:, :dc] = xp[:, :dc];  z[:, dc:] = (xp[:, dc:] - t) * exp(-s);  logdet += c - sum_j s
 *   inv: xp[:, dc:] = z[:, dc:] * exp(s) + t;                              logdet += sum_j s - c
 *   bwd: ds = -dz_t * z_t - dlogdet; dt = -dz_t * exp(-s); dxp = [dz_c, dz_t * exp(-s)]
 * s, t are (B, D/2) WITHOUT the last-layer bias; bias_s/bias_t (D/2) are added inside (may be NULL).
 * logdet may be NULL. */
NF_API int nf_affine_logdet_fwd(const float* xp, const float* s, const float* t,
                         const float* bias_s, const float* bias_t, float logdet_const,
                         int64_t B, int32_t D, float* z, float* logdet, float* s_out, void* stream);
NF_API int nf_affine_logdet_inv(const float* z, const float* s, const float* t,
                         const float* bias_s, const float* bias_t, float logdet_const,
                         int64_t B, int32_t D, float* xp, float* logdet, void* stream);
NF_API int nf_affine_logdet_bwd(const float* dz, const float* z, const float* s_full, const float* dlogdet,
                         int64_t B, int32_t D, float* dxp, float* ds, float* dt, void* stream);

/* Standalone PLU build (torch_fcnf.py:34-38, :45-54): W (D,D), optional W_inv (D,D), logdet (1). */
NF_API int nf_plu_build(const float* s, const float* U, const float* L, const float* P, const float* P_inv,
                 int32_t D, float* W, float* W_inv, float* logdet, void* workspace, int64_t ws_bytes,
                 void* stream);

/* One conditioner Linear  C (M,N) = A (M,K) @ W (N,K)^T + bias  (torch_fcnf.py:75-77; nn.Linear weight
 * layout).  Row pitches lda/ldw/ldc in floats.  engine: NF_PREC_FP32_SIMT (FFMA), NF_PREC_TF32X3 (tcgen05,
 * fp32 parity), 3 = tcgen05 single TF32 pass, 4 = 3xTF32 feeding raw fp32 as the hi operand; 5 / 6 = as 1 / 4
 * with 32-element (128 B) pipeline stages instead of 16-element ones. */
NF_API int nf_linear(const float* A, int64_t lda, const float* W, int64_t ldw, const float* bias, float* C,
                     int64_t ldc, int64_t M, int32_t N, int32_t K, int32_t engine, void* stream);

/* Weight gradient of one Linear: dW (n_out, n_in) += dY (B, n_out)^T @ X (B, n_in); the reduction runs over
 * the batch rows (what autograd computes for nn.Linear.weight, torch_fcnf.py:75-77).  engine as nf_linear
 * (0 = FFMA, 1 = tcgen05 3xTF32, 3 = single TF32 pass). */
NF_API int nf_linear_wgrad(const float* dY, int64_t lddy, const float* X, int64_t ldx, float* dW, int64_t lddw,
                           int64_t B, int32_t n_out, int32_t n_in, int32_t engine, void* stream);

/* The steps either side of the flow in the reference's training loop:
 *   nf_nll_forward    loss[0] = -mean_b(-0.5 |z_b|^2 - D/2 log 2 pi + logdet_b)   (torch_nfbc_ant.py:269, N(0, I) prior)
 *   nf_nll_cotangents dz = g z / B, dlogdet = -g / B  with the upstream gradient g[0] (g may be NULL = 1)
 *   nf_adamw_step     one torch.optim.AdamW step (torch_nfbc_ant.py:150-155) over the flat parameter arena in one
 *                     launch; exp_avg / exp_avg_sq are arenas of the same layout, slot_mask (n_blocks * NF_P_SLOTS
 *                     bytes on the device, 1 = update) leaves P / P_inv and frozen parameters alone; step >= 1. */
/*   nf_noise_augment  out[i] = x[i] + std * N(0, 1): the training-time noise augmentation of torch_nfbc_ant.py:264-265 as
 *                     one kernel (counter-based Philox4x32-10 + Box-Muller; the normal at index i is a pure function of
 *                     (seed, offset + i / 4, i % 4), so a caller advances `offset` by ceil(n / 4) per call); out may
 *                     alias x. */
/*   nf_stdnormal_logprob  lp[b] = log N(z_b; 0, I) = -0.5 |z_b|^2 - D/2 log 2 pi: `model.prior.log_prob(z)` of the reference's
 *                     loss line for its MultivariateNormal(0, I) prior (torch_nfbc_ant.py:146,269) as one kernel;
 *   nf_stdnormal_logprob_bwd  dz[b, j] = -g[b] z[b, j]. */
NF_API int nf_stdnormal_logprob(const float* z, int64_t B, int32_t D, float* lp, void* stream);
NF_API int nf_stdnormal_logprob_bwd(const float* z, const float* g, int64_t B, int32_t D, float* dz, void* stream);
/*   nf_logprob_group_argmin  per group g of N candidates: n* = argmin_n [log N(z_gn; 0, I) + logdet_gn] (ties: smaller n),
 *                     out_x[g] = x[g, n*], out_idx[g] = n*, out_lp[g] = the minimum (each output optional): the exploration-goal
 *                     selection of unsupservised-gcrl/train_auto_NF_rl.py:328-333 without the (G, N) round trip. */
NF_API int nf_logprob_group_argmin(const float* z, const float* logdet, const float* x, int64_t G, int64_t N, int32_t D,
                                   float* out_x, int64_t* out_idx, float* out_lp, void* stream);
NF_API int nf_noise_augment(const float* x, int64_t n, float std_, uint64_t seed, uint64_t offset, float* out, void* stream);
NF_API int nf_nll_forward(const float* z, const float* logdet, int64_t B, int32_t D, float* loss, void* stream);
NF_API int nf_nll_cotangents(const float* z, const float* g, int64_t B, int32_t D, float* dz, float* dld, void* stream);
NF_API int nf_adamw_step(const NfFlowDesc* desc, float* params, const float* grads, float* exp_avg, float* exp_avg_sq,
                         const unsigned char* slot_mask, float lr, float beta1, float beta2, float eps,
                         float weight_decay, int64_t step, void* stream);

/* ---- conditioning encoder ---------------------------------------------------------
 * The MLP that produces the flow's conditioning input y in every reference caller (GEncoder,
 * torch_fcnf.py:154-190): n_hidden x [Linear -> LayerNorm(affine) -> SiLU] followed by one Linear.  The reference
 * fixes hidden = 512 and n_hidden = 4; both are free here (hidden % 4 == 0, hidden <= 1024, n_hidden <= 8).
 * Parameter arena (fp32, offsets from nf_encoder_param_layout, every slot 16-byte aligned): for hidden layer i
 *   offs[4 i + 0] Linear weight (hidden, in_i) row-major (nn.Linear layout), offs[4 i + 1] Linear bias (hidden),
 *   offs[4 i + 2] LayerNorm weight (hidden), offs[4 i + 3] LayerNorm bias (hidden);
 * then offs[4 n_hidden] output weight (out_features, hidden) and offs[4 n_hidden + 1] output bias.
 * Forward writes `out` (B, out_features) and, when `stash` is given, what backward needs (a copy of g included).
 * Backward takes dout (B, out_features), writes dparams (same layout as params, overwritten) and optionally dg
 * (B, in_features).  Like the flow calls, both replay a cached launch graph keyed on (desc, B, params, stash,
 * dparams, workspace); g / out / dout / dg are staged through the workspace and may change from call to call.
 * Same conventions as the flow calls: caller-owned buffers, workspace passed in, asynchronous on `stream`. */
typedef struct NfEncoderDesc {
  int32_t in_features, hidden, out_features, n_hidden;
  float ln_eps;      /* 1e-5 (torch) / 1e-6 (flax, nf.py:188-212) */
  int32_t precision; /* NF_PREC_* */
} NfEncoderDesc;
#define NF_ENC_MAX_HIDDEN_LAYERS 8
NF_API int64_t nf_encoder_param_layout(const NfEncoderDesc* desc, int64_t* offs);  /* returns total floats, < 0 on error */
NF_API int64_t nf_encoder_stash_bytes(const NfEncoderDesc* desc, int64_t B);
NF_API int64_t nf_encoder_workspace_bytes(const NfEncoderDesc* desc, int64_t B);
NF_API int nf_encoder_forward(const NfEncoderDesc* desc, const float* params, const float* g, int64_t B, float* out,
                              void* stash, void* workspace, int64_t ws_bytes, void* stream);
NF_API int nf_encoder_backward(const NfEncoderDesc* desc, const float* params, const void* stash, int64_t B,
                               const float* dout, float* dg, float* dparams, void* workspace, int64_t ws_bytes,
                               void* stream);

/* Causal-transformer flow: the model `il-and-cil/torch_nfgcbc.py:27,119` imports from the
 * absent `utils/torch_networks.py` -- RealNVP(input_dims, channels, rep_dim, head_dim, blocks, layers_per_block, prior),
 * model(actions (B, A, 1), cond (B, 1, rep)) -> (z, outputs, logdets) (:136-142), model.reverse(z, cond)
 * (utils/torch_evaluations.py:79-82).  PARITY UNPINNED upstream; the arithmetic is the published TarFlow meta block as
 * restated in oracle/torch_tarflow.py (the checker): T = A tokens (the conditioning token, then action dimensions
 * 0 .. A-2), n_layers x [pre-LN causal attention + GELU MLP], token t yields (xa_t, xb_t), z_t = (x_t - xb_t) exp(-xa_t),
 * logdet = -sum_t xa_t; odd blocks run on the reversed dimension order.  Parameters live in one flat arena
 * (nf_tarflow_param_layout); forward keeps what backward needs in `stash` (NULL = inference); inverse is the
 * autoregressive loop over the A dimensions with the per-layer K / V rows cached in the workspace.  x, z, dz, dx are
 * (B, A); y, dy are (B, R); logdet, dlogdet are (B).  outputs (n_blocks, B, A) is optional.  Conventions as the flow. */
typedef struct NfTarflowDesc {
  int32_t A;          /* action dimensions (tokens), 1 .. 32                                  */
  int32_t C;          /* channels, multiple of 4, <= 1024                                     */
  int32_t R;          /* rep_dim of the conditioning vector                                   */
  int32_t head_dim;   /* 32, 64 or 128; divides C                                             */
  int32_t n_blocks;
  int32_t n_layers;   /* layers per block                                                     */
  int32_t expansion;  /* MLP width = expansion * C                                            */
  float   ln_eps;
  float   logscale_clamp; /* c > 0: xa = c tanh(raw / c); 0: raw                              */
  int32_t precision;  /* NF_PREC_FP32_SIMT or NF_PREC_TF32X3                                  */
} NfTarflowDesc;
/* offs[22] (floats): block_stride, then offsets inside a block: pos_embed (A, C), proj_in.weight (C), proj_in.bias (C),
 * cond_proj.weight (C, R), cond_proj.bias (C), layers, layer_stride, then inside a layer: attn.norm.{weight,bias},
 * attn.qkv.{weight (3C, C),bias}, attn.proj.{weight,bias}, mlp.norm.{weight,bias}, mlp.fc1.{weight (EC, C),bias},
 * mlp.fc2.{weight (C, EC),bias}; then inside the block again: proj_out.weight (2, C), proj_out.bias (2).
 * Returns the arena size in floats (< 0 on error). */
#define NF_TARFLOW_LAYOUT_SLOTS 22
NF_API int64_t nf_tarflow_param_layout(const NfTarflowDesc* desc, int64_t* offs);
NF_API int64_t nf_tarflow_stash_bytes(const NfTarflowDesc* desc, int64_t B);
NF_API int64_t nf_tarflow_workspace_bytes(const NfTarflowDesc* desc, int64_t B);
NF_API int nf_tarflow_forward(const NfTarflowDesc* desc, const float* params, const float* x, const float* y, int64_t B,
                              float* z, float* logdet, float* outputs, void* stash, void* workspace, int64_t ws_bytes,
                              void* stream);
NF_API int nf_tarflow_inverse(const NfTarflowDesc* desc, const float* params, const float* z, const float* y, int64_t B,
                              float* x, float* logdet, void* workspace, int64_t ws_bytes, void* stream);
NF_API int nf_tarflow_backward(const NfTarflowDesc* desc, const float* params, const void* stash, int64_t B,
                               const float* dz, const float* dlogdet, float* dx, float* dy, float* dparams,
                               void* workspace, int64_t ws_bytes, void* stream);

/* Optional per-kernel-class timing (CUDA events on the launching stream; off by default).  Kinds:
 * 0 tcgen05 GEMM, 1 tcgen05 weight-gradient GEMM, 2 FFMA GEMM, 3 LeakyReLU+LayerNorm fwd, 4 its backward,
 * 5 affine coupling kernels, 6 parameter-sized kernels, 7 row-fused forward/inverse tail, 8 row-fused
 * backward head, 9 row-fused backward tail, 10 column reductions over the batch rows, 11 transformer-flow row kernels, 12 transformer-flow
 * attention.  nf_prof_collect sums and clears the records:
 * ms[k] total milliseconds, work[k] algorithmic FLOPs (GEMMs) or bytes (row kernels), launches[k]. */
#define NF_PROF_KINDS 13
NF_API int nf_prof_enable(int on);
NF_API int nf_prof_collect(double* ms, double* work, int64_t* launches, int nkinds);

/* Launch-sequence replay (on by default; NF_B200_GRAPHS=0 disables): prepare / forward / inverse /
 * backward capture their kernel sequence into a CUDA graph the second time the same (descriptor, B,
 * pointers) come back and replay it with one cudaGraphLaunch on the caller's stream afterwards.  Buffers
 * are still caller-owned; a cached graph only refers to the pointers it was keyed on. */
NF_API int nf_graphs_enable(int on);
NF_API int nf_graphs_clear(void);
NF_API int nf_graphs_stats(int64_t* replays, int64_t* captures, int64_t* eager);

/* Run-time options, by name (also NF_B200_<NAME> in the environment, read once): "pdl" (programmatic dependent
 * launch, 1), "streams" (backward on three streams, 1), "tma_store" (bulk-tensor-store epilogues, 1), "nacc" (main
 * accumulators in rotation, 0 = 3), "fused_ln" (LayerNorm in the GEMM epilogue: -1 auto / 0 / 1), "fused_tail"
 * (flow tail in the layer-2 GEMM, 0), "bn256" (256-wide tiles: 0 auto / 1 / 2 never), "stack" (persistent whole-stack
 * forward / inverse kernel: -1 auto / 0 / 1), "ride" (bias / x_cond weight gradients
 * taken along by the weight-gradient GEMMs instead of a column-reduction pass, 1), "defer_wgrad" (conditioner weight
 * gradients of all blocks in two batched launches after the block loop: -1 auto while the flow is small / 0 / 1), "mcast"
 * (weight tiles of the large conditioner GEMMs multicast over CTA pairs: -1 auto / 0 / 1), "whatif" (experiments only: skips
 * backward work, wrong gradients).  Every setting computes the
 * same function (tests/test_flow_gpu.py::test_options_are_equivalent); setting one clears the launch graphs. */
NF_API int nf_set_option(const char* name, int value);
NF_API int nf_get_option(const char* name);

/* Number of kernels this library launched on the calling thread since the last reset
 * (bench.py's gpu_launches). */
NF_API int64_t nf_launch_count(int reset);

#ifdef __cplusplus
}
#endif
#endif /* NF_B200_H */
