// Orchestration of the coupling-flow hot path and the C-ABI entry points (include/nf_b200.h).
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "dp.cuh"
#include "elementwise.cuh"
#include "flow_stack.cuh"
#include <vector>
#include "graphs.cuh"
#include "linear.cuh"
#include "rowfused.cuh"
#include "rowlocal.cuh"
#include "simt_gemm.cuh"
#include "tc_gemm.cuh"

namespace nf {

// ---------------------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static int64_t g_launches = 0;  // process-wide (autograd runs backward on its own thread)
char* err_buf() { return g_err; }
int64_t& launch_counter() { return g_launches; }
int set_err(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

namespace {
struct OptDef { const char* name; const char* env; int def; };
const OptDef kOpts[NF_OPT_COUNT] = {
    {"pdl", "NF_B200_PDL", 1},           {"streams", "NF_B200_STREAMS", 1},       {"tma_store", "NF_B200_TMA_STORE", 1},
    {"nacc", "NF_B200_NACC", 0},         {"fused_ln", "NF_B200_FUSED_LN", -1},    {"fused_tail", "NF_B200_FUSED_TAIL", 0},
    {"bn256", "NF_B200_BN256", 0},       {"ts", "NF_B200_TS", 1},                 {"presplit", "NF_B200_PRESPLIT", 1},
    {"stack", "NF_B200_STACK", -1},      {"stack_nsg", "NF_B200_STACK_NSG", 2},   {"whatif", "NF_B200_WHATIF", 0},
    {"ride", "NF_B200_RIDE", 1},         {"defer_wgrad", "NF_B200_DEFER_WGRAD", -1},
    {"mcast", "NF_B200_MCAST", -1},      {"bn64", "NF_B200_BN64", 1},
    {"tf_ar", "NF_B200_TF_AR", -1},      {"prep_fused", "NF_B200_PREP_FUSED", 1}, {"defer_dy", "NF_B200_DEFER_DY", 1},
    {"defer_cols", "NF_B200_DEFER_COLS", 1}, {"defer_skinny", "NF_B200_DEFER_SKINNY", 0},
};
int g_opt[NF_OPT_COUNT];
bool g_opt_init = false;
void opt_init() {
  if (g_opt_init) return;
  for (int i = 0; i < NF_OPT_COUNT; ++i) {
    const char* e = getenv(kOpts[i].env);
    g_opt[i] = e ? atoi(e) : kOpts[i].def;
  }
  g_opt_init = true;
}
}  // namespace

int option(int id) {
  opt_init();
  return (id >= 0 && id < NF_OPT_COUNT) ? g_opt[id] : 0;
}

int check_desc(const NfFlowDesc* d) {
  NF_CHECK_ARG(d != nullptr, NF_E_BADARG, "desc is NULL");
  NF_CHECK_ARG(d->D >= 2 && d->D <= 4096, NF_E_BADARG, "D=%d out of range [2,4096]", d->D);
  NF_CHECK_ARG(d->C >= 1 && d->C <= 8192, NF_E_BADARG, "C=%d out of range", d->C);
  NF_CHECK_ARG(d->H1 >= 1 && d->H1 <= 16384, NF_E_BADARG, "H1=%d out of range", d->H1);
  NF_CHECK_ARG(d->R >= 0 && d->R <= 65536, NF_E_BADARG, "R=%d out of range", d->R);
  NF_CHECK_ARG(d->n_blocks >= 1 && d->n_blocks <= 4096, NF_E_BADARG, "n_blocks=%d out of range", d->n_blocks);
  NF_CHECK_ARG(d->ln_eps > 0.f, NF_E_BADARG, "ln_eps must be > 0");
  NF_CHECK_ARG(d->precision >= 0 && d->precision <= 3, NF_E_BADARG, "unknown precision %d", d->precision);
  NF_CHECK_ARG(d->precision != NF_PREC_BF16, NF_E_UNSUPPORTED, "the bf16 mode is not built (use NF_PREC_TF32X3)");
  return 0;
}

void param_layout(const Dims& m, NfParamLayout* out) {
  const int64_t DD = (int64_t)m.D * m.D;
  int64_t numel[NF_P_SLOTS];
  numel[NF_P_S] = m.D; numel[NF_P_U] = DD; numel[NF_P_L] = DD; numel[NF_P_P] = DD; numel[NF_P_PINV] = DD;
  for (int net = 0; net < 2; ++net) {
    const int o = net * NF_P_NET_SLOTS;
    numel[NF_P_W1 + o] = (int64_t)m.H1 * m.K1; numel[NF_P_B1 + o] = m.H1;
    numel[NF_P_G1 + o] = m.H1; numel[NF_P_BE1 + o] = m.H1;
    numel[NF_P_W2 + o] = (int64_t)m.C * m.H1; numel[NF_P_B2 + o] = m.C;
    numel[NF_P_G2 + o] = m.C; numel[NF_P_BE2 + o] = m.C;
    numel[NF_P_W3 + o] = (int64_t)m.dt * m.C; numel[NF_P_B3 + o] = m.dt;
  }
  int64_t off = 0;
  for (int s = 0; s < NF_P_SLOTS; ++s) {
    out->slot_off[s] = off;
    out->slot_numel[s] = numel[s];
    off += pad4(numel[s]);
  }
  out->block_stride = off;
  out->net_stride = out->slot_off[NF_P_W1 + NF_P_NET_SLOTS] - out->slot_off[NF_P_W1];
  out->total = off * m.n;
}

PackedLayout packed_layout(const Dims& m) {
  PackedLayout k;
  const int64_t DD = pad4((int64_t)m.D * m.D);
  int64_t off = 0;
  k.W = off; off += DD;
  k.Winv = off; off += DD;
  k.Lh = off; off += DD;
  k.Uh = off; off += DD;
  k.PL = off; off += DD;
  k.WT = off; off += DD;
  k.WinvT = off; off += DD;
  k.net0 = off;
  int64_t n = 0;
  k.W2f = n; n += pad4((int64_t)m.C * m.H1);
  k.b2f = n; n += pad4(m.C);
  k.W3f = n; n += pad4((int64_t)m.dt * m.C);
  k.b3f = n; n += pad4(m.dt);
  k.W2fT = n; n += pad4((int64_t)m.H1 * m.C);
  k.W1T = n; n += pad4((int64_t)m.K1 * m.H1);
  k.dcp = (int)pad4(m.dc);
  k.K1p = (int)pad4(k.dcp + m.R);
  k.W1p = n; n += pad4((int64_t)m.H1 * k.K1p);
  k.net_stride = n;
  off += 2 * n;
  k.block_stride = off;
  k.logdet_c = off * m.n;
  k.logdet_total = k.logdet_c + pad4(m.n);
  k.total = k.logdet_total + 4;
  k.hi_off = k.lo_off = 0;
  if (m.prec == NF_PREC_TF32X3) {
    const int64_t base = pad32(k.total);
    k.hi_off = base; k.lo_off = 2 * base;
    k.total = 3 * base;
  }
  return k;
}

StashLayout stash_layout(const Dims& m, int64_t B) {
  StashLayout s;
  s.mw1 = (m.H1 + 31) / 32;
  s.mw2 = (m.C + 31) / 32;
  int64_t off = 0;
  s.xs = off; off += pad4((int64_t)(m.n + 1) * B * m.D);
  s.ss = off; off += pad4((int64_t)m.n * B * m.dt);
  s.y = off; off += pad4(B * m.R);
  s.blocks = off;
  int64_t n = 0;
  s.xh1 = n; n += pad4(B * m.H1);
  s.xh2 = n; n += pad4(B * m.C);
  s.r1 = n; n += pad4(B);
  s.r2 = n; n += pad4(B);
  s.m1 = n; n += pad4(B * s.mw1);
  s.m2 = n; n += pad4(B * s.mw2);
  s.net_stride = n;
  s.block_stride = 2 * n;
  s.total = off + s.block_stride * m.n;
  return s;
}

// ---------------------------------------------------------------------------------------------
// workspace carving
// ---------------------------------------------------------------------------------------------
struct Carver {
  float* base;
  int64_t off = 0;
  explicit Carver(void* p) : base(reinterpret_cast<float*>(p)) {}
  float* take(int64_t n) {
    float* r = base ? base + off : nullptr;
    off += pad32(n);
    return r;
  }
};

// io_*: staging copies of the caller's inputs / outputs.  The kernels only ever touch caller-owned PERSISTENT
// buffers (workspace, stash, packed, params); the caller's x / y / z / logdet / ... are copied in and out with
// plain stream-ordered copies around the kernel sequence, so the cached launch graph does not depend on
// the I/O pointers (fresh tensors every step would otherwise force a new capture each time).
struct FwdWs { float *xp, *u, *o, *xh1, *xh2, *pp0, *pp1, *io_x, *io_y, *io_z, *io_ld, *io_outs; int64_t total; int Dp; };
static FwdWs carve_fwd(const Dims& m, int64_t B, void* ws) {
  Carver c(ws);
  FwdWs w;
  const int64_t Hm = std::max(m.H1, m.C);
  w.Dp = (int)pad4(m.D);   // internal (B,D) buffers use a 16-byte row pitch so TMA can read x_cond
  w.xp = c.take(B * w.Dp);
  w.u = c.take(2 * B * Hm);
  w.o = c.take(2 * B * m.dt);
  w.xh1 = c.take(2 * B * m.H1);
  w.xh2 = c.take(2 * B * m.C);
  w.pp0 = c.take(B * w.Dp);
  w.pp1 = c.take(B * w.Dp);
  w.io_x = c.take(B * w.Dp);
  w.io_y = c.take(B * m.R);
  w.io_z = c.take(B * w.Dp);
  w.io_ld = c.take(B);
  w.io_outs = c.take((int64_t)m.n * B * m.D);
  w.total = c.off;
  return w;
}

struct BwdWs { float *d0, *d1, *dxp, *dout, *gA, *gB, *dWs, *T1, *dLh, *dUh, *sumdld;
               float *dxp2, *dout2, *gA2, *gB2;   // second set: blocks alternate so side streams can lag one block
               float *io_dz, *io_dld, *io_dy, *io_dx;   // staging of the caller's cotangents and of dx / dy (see FwdWs)
               float *w3t[2];                           // generic path: W3f^T (2, C, dtp) of the current block (alternating)
               float *doutAll, *dxpAll;                 // dout / dxp of every block (deferred column reductions)
               float *gAall, *gBall;                    // du2 / du1 of EVERY block (n, 2, B, Hm) when the weight gradients are deferred
               int64_t total; };
// The conditioner weight gradients of a block only consume du2 / du1.  While the flow is small enough to keep those for
// every block (2 x n x 2 x B x Hm floats), backward skips the per-block weight-gradient GEMMs and runs TWO batched
// launches after the block loop (batch = blocks x nets): 12 launches of 15-16 us with 36-74-way split-K become 2 with a
// 4-6-way split (C2a).  Larger flows keep the per-block launches, which overlap the data-gradient chain on side streams.
constexpr int64_t kDeferWgradMaxBytes = int64_t(640) << 20;
static bool defer_fits(const Dims& m, int64_t B) {
  const int64_t Hm = std::max(m.H1, m.C);
  return (int64_t)m.n * 2 * 2 * B * Hm * 4 <= kDeferWgradMaxBytes;
}
static BwdWs carve_bwd(const Dims& m, int64_t B, void* ws) {
  Carver c(ws);
  BwdWs w;
  const int64_t Hm = std::max(m.H1, m.C);
  const int64_t DD = (int64_t)m.D * m.D;
  w.d0 = c.take(B * m.D);
  w.d1 = c.take(B * m.D);
  w.dxp = c.take(B * m.D);
  w.dout = c.take(2 * B * m.dt);
  w.gA = c.take(2 * B * Hm);
  w.gB = c.take(2 * B * Hm);
  w.dWs = c.take(m.n * DD);
  w.T1 = c.take(m.n * DD);
  w.dLh = c.take(m.n * DD);
  w.dUh = c.take(m.n * DD);
  w.sumdld = c.take(4);
  w.dxp2 = c.take(B * m.D);
  w.dout2 = c.take(2 * B * m.dt);
  w.gA2 = c.take(2 * B * Hm);
  w.gB2 = c.take(2 * B * Hm);
  w.io_dz = c.take(B * m.D);
  w.io_dld = c.take(B);
  w.io_dy = c.take(B * m.R);
  w.io_dx = c.take(B * m.D);
  w.w3t[0] = c.take(2 * (int64_t)m.C * pad4(m.dt));
  w.w3t[1] = c.take(2 * (int64_t)m.C * pad4(m.dt));
  w.gAall = w.gBall = w.doutAll = w.dxpAll = nullptr;
  if (defer_fits(m, B)) {     // (sized by shape alone, so the workspace size does not depend on run-time options)
    w.gAall = c.take((int64_t)m.n * 2 * B * Hm);
    w.gBall = c.take((int64_t)m.n * 2 * B * Hm);
    w.doutAll = c.take((int64_t)m.n * 2 * B * m.dt);
    w.dxpAll = c.take((int64_t)m.n * B * m.D);
  }
  w.total = c.off;
  return w;
}

// ---------------------------------------------------------------------------------------------
// side streams of the backward pass
// ---------------------------------------------------------------------------------------------
// Per block only  head -> data gradient through layer 2/1 -> tail  is a dependent chain; the two weight-
// gradient GEMMs, the dy GEMM and the column reductions only consume its intermediates.  They run on two
// library-owned side streams (fork/join with events, which also become parallel branches of the captured
// graph), so the critical path per block is 3 kernels instead of 8.  Blocks alternate between two sets of
// intermediate buffers and the main stream waits for the side work of block i+2 before it overwrites them.
struct SideStreams {
  cudaStream_t s1 = nullptr, s2 = nullptr, s3 = nullptr;   // s3: gradient all-reduce of finished buckets (data parallel)
  cudaEvent_t evA = nullptr, evB = nullptr, evC = nullptr, done1[2] = {nullptr, nullptr}, done2[2] = {nullptr, nullptr};
  cudaEvent_t evD = nullptr, done3 = nullptr;
  bool ok = false;
};
static SideStreams* side_streams(bool force = false) {
  static SideStreams per_dev[64];
  int dev = 0;
  if ((!force && !option(NF_OPT_STREAMS)) || cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  SideStreams& s = per_dev[dev];
  if (!s.ok) {
    bool good = cudaStreamCreateWithFlags(&s.s1, cudaStreamNonBlocking) == cudaSuccess &&
                cudaStreamCreateWithFlags(&s.s2, cudaStreamNonBlocking) == cudaSuccess &&
                cudaStreamCreateWithFlags(&s.s3, cudaStreamNonBlocking) == cudaSuccess;
    cudaEvent_t* evs[] = {&s.evA, &s.evB, &s.evC, &s.done1[0], &s.done1[1], &s.done2[0], &s.done2[1], &s.evD, &s.done3};
    for (cudaEvent_t* e : evs) good = good && cudaEventCreateWithFlags(e, cudaEventDisableTiming) == cudaSuccess;
    if (!good) { cudaGetLastError(); return nullptr; }
    s.ok = true;
  }
  return &s;
}

struct PrepWs { float *uinv, *linv, *T; int64_t total; };
static PrepWs carve_prep(const Dims& m, void* ws) {
  Carver c(ws);
  PrepWs w;
  const int64_t DD = (int64_t)m.D * m.D;
  w.uinv = c.take(m.n * DD);
  w.linv = c.take(m.n * DD);
  w.T = c.take(m.n * DD);
  w.total = c.off;
  return w;
}


// C[z] (M,N) (+)= [A0 | A1][z] (M,K0+K1) @ [B0 | B1][z] (N,K0+K1)^T : the Linear layers and their data
// gradients.  precision 1 (3xTF32) runs on tcgen05 when the tile shapes make sense and TMA can
// address the operands; everything else (tiny N or K, odd pitches) is exact-fp32 FFMA.
int linear_nt(int prec, const TcGemmP& t0, cudaStream_t st) {
  TcGemmP t = t0;
  // Truncating the hi operand (raw_hi) leaves a one-signed lo, so the dropped lo*lo products are a
  // systematic 2^-22-relative bias that grows with the reduction length; long reductions use the
  // round-to-nearest split (zero-mean lo) -- measured on C5 (K = 1024, 24 blocks): dy error 1.08e-5 -> inside 1e-5.
  if (t.K0 + t.K1 >= 512) t.raw_hi = 0;
  if (prec == NF_PREC_TF32) { t.passes = 1; t.raw_hi = 1; t.B0hi = t.B0lo = t.B1hi = t.B1lo = nullptr; }
  if (prec_tc(prec) && t.N >= 16 && (t.K0 + t.K1) >= 32 && tc_gemm_supported(t)) return tc_gemm(t, st);
  GemmP g;
  g.A0 = t.A0; g.A1 = t.A1; g.B0 = t.B0; g.B1 = t.B1; g.C = t.C; g.bias = t.bias;
  g.M = t.M; g.N = t.N; g.K0 = t.K0; g.K1 = t.K1;
  g.lda0 = (int)t.lda0; g.lda1 = (int)t.lda1; g.ldb0 = (int)t.ldb0; g.ldb1 = (int)t.ldb1; g.ldc = (int)t.ldc;
  g.sA0 = t.sA0; g.sA1 = t.sA1; g.sB0 = t.sB0; g.sB1 = t.sB1; g.sC = t.sC; g.sBias = t.sBias;
  g.batch = t.batch; g.transB = 1; g.accumulate = t.accumulate;
  return simt_gemm(g, st);
}

// true when wgrad_tn(prec, t) will run a kernel that can take colsum / skinny reductions along (option "ride")
bool rides_ok(int prec, const TcGemmTnP& t) {
  return option(NF_OPT_RIDE) != 0 && prec == NF_PREC_TF32X3 && t.M >= 16 && t.N >= 32 && tc_gemm_tn_rides_supported(t);
}

// C[z] (M,N) += A[z] (K,M)^T @ B[z] (K,N): weight gradients (reduction over the K = batch rows).
// C must already hold the base value (the gradient arena is zero-filled first).
int wgrad_tn(int prec, const TcGemmTnP& t0, cudaStream_t st) {
  TcGemmTnP t = t0;
  if (prec == NF_PREC_TF32) t.passes = 1;
  if (prec_tc(prec) && t.M >= 16 && t.N >= 32 && tc_gemm_tn_supported(t)) return tc_gemm_tn(t, st);
  GemmP g;
  g.transA = 1; g.A0 = t.A; g.lda0 = (int)t.lda; g.sA0 = t.sA; g.K0 = (int)t.K;
  g.B0 = t.B; g.ldb0 = (int)t.ldb; g.sB0 = t.sB;
  g.C = t.C; g.ldc = (int)t.ldc; g.sC = t.sC;
  g.M = t.M; g.N = t.N; g.batch = t.batch;
  g.splitk = pick_splitk(g.M, g.N, t.K, t.batch);
  if (g.splitk == 1) g.accumulate = 1;
  return simt_gemm(g, st);
}

// C (B, D) = A (B, D) @ Bk^T with Bk a (D, D) table of the packed block (so its hi / lo copies exist): the
// invertible-linear matmuls x @ W (Bk = W^T), xp @ W^-1 (Bk = W^-T), dxp @ W^T (Bk = W), g @ W^-T (Bk = W^-1).
// Wide flows run them on the tensor cores (measured: the FFMA GEMM took 46 us per launch on the D = 400 ant flow,
// 47 % of its training step); narrow or unaligned ones stay FFMA.
static int plu_matmul(const Dims& m, const PackedLayout& kl, const float* A, int lda, const float* Bk, float* C, int ldc,
                      int64_t B, cudaStream_t st) {
  if (prec_tc(m.prec) && m.D >= 32 && (m.D & 3) == 0 && (lda & 3) == 0 && (ldc & 3) == 0) {
    TcGemmP t;
    t.A0 = A; t.lda0 = lda; t.K0 = m.D;
    t.B0 = Bk; t.ldb0 = m.D;
    if (kl.hi_off) { t.B0hi = Bk + kl.hi_off; t.B0lo = Bk + kl.lo_off; }
    t.C = C; t.ldc = ldc; t.M = (int)B; t.N = m.D;
    t.raw_hi = 0;   // round-to-nearest split: the truncating one biases |z| low by ~1e-5 of the D = 400 log-prob sum
    // The invertible linear layer is the flow's own transform, not a conditioner: in the reduced-precision mode it stays
    // at 3xTF32 (a single TF32 pass here put its 1e-3 rounding straight into z and the inverse: C5 z error 1.4e-3 -> 1.7e-2
    // when this matmul first moved to the tensor cores) -- it is 1-12 % of a block's FLOPs.
    return linear_nt(m.prec == NF_PREC_TF32 ? NF_PREC_TF32X3 : m.prec, t, st);
  }
  GemmP g;
  g.A0 = A; g.lda0 = lda; g.K0 = m.D;
  g.B0 = Bk; g.ldb0 = m.D; g.transB = 1;
  g.C = C; g.ldc = ldc; g.M = (int)B; g.N = m.D;
  return simt_gemm(g, st);
}

// Small flow dimensions use the row-fused kernels of rowfused.cuh for everything around the wide GEMMs.
static bool use_row_fused(const Dims& m) {
  if (m.D > kRowFusedMaxD) return false;
  // row-local kernels: worth it while the skinny weight matrices fit in shared memory (measured: C5 with
  // dt = 16, C = 1024 is 45 % slower on this path than on the GEMM + row-kernel path)
  if (row_v2_supported(m.D, m.C, m.H1))
    return (size_t)2 * m.dt * m.C * 4 <= 40 * 1024 && (size_t)2 * m.dc * m.H1 * 4 <= 40 * 1024;
  return row_bwd_head_smem(m.D, m.C) <= kRowFusedSmemLimit && row_bwd_tail_smem(m.D, m.H1) <= kRowFusedSmemLimit;
}

// The two conditioner MLPs (net 0 = t, net 1 = s) up to the last Linear's raw output o (2,B,dt):
//   u1 = [x_cond | y] W1^T -> LeakyReLU -> LN -> x_hat1 -> W2f -> LeakyReLU -> LN -> x_hat2 -> W3f
// (torch_fcnf.py:74-78, :85-89, :99-100).  `xc` holds x_cond in its first dc columns (row pitch ldxc).
static int conditioners(const Dims& m, const NfParamLayout& pl, const PackedLayout& kl, const float* pb,
                        const float* kb, const float* xc, int ldxc, const float* y, int64_t B, const FwdWs& w,
                        float* xh1, int64_t xh1_net, float* xh2, int64_t xh2_net, float* r1, float* r2,
                        int64_t r_net, uint32_t* m1, uint32_t* m2, int64_t m_net, int mw1, int mw2,
                        bool with_last, cudaStream_t st, const TailP* tail = nullptr, bool* tail_done = nullptr,
                        bool store_xh2 = true) {
  const int64_t Hm = std::max(m.H1, m.C);
  // Linear + LeakyReLU + LayerNorm as ONE tcgen05 kernel (row statistics exchanged across the CTA cluster
  // that spans the row) when the width allows it; otherwise GEMM -> u, then the row-wise kernel.
  if (tail_done) *tail_done = false;
  auto linear_lrelu_ln = [&](TcGemmP g, int N, const float* bias, int64_t bias_net, float* xh, int64_t xh_net,
                             float* r, uint32_t* mk, int mw, const TailP* tl) -> int {
    g.M = (int)B; g.N = N; g.batch = 2;
    if (g.K0 + g.K1 >= 512) g.raw_hi = 0;
    if (m.prec == NF_PREC_TF32) { g.passes = 1; g.raw_hi = 1; g.B0hi = g.B0lo = g.B1hi = g.B1lo = nullptr; }
    TcGemmP f = g;
    f.C = xh; f.ldc = N; f.sC = xh_net; f.bias = bias; f.sBias = bias_net;
    f.epi.kind = TC_EPI_LN_FWD; f.epi.eps = m.eps; f.epi.fast_var = m.fastvar;
    f.epi.rstd = r; f.epi.sRstd = r_net; f.epi.mask = mk; f.epi.sMask = m_net; f.epi.mw = mk ? mw : 0;
    if (prec_tc(m.prec) && (g.K0 + g.K1) >= 32 && tc_gemm_epi_supported(f)) {
      if (tl && tc_gemm_tail_supported(f, tl->D)) {
        // the whole rest of the block (last Linear, affine coupling, log-det, neighbouring PLU matmul) rides on
        // this launch: the cluster holds both nets, its leader CTA finishes the rows
        TcEpi::Tail& t = f.epi.tail;
        t.on = 1; t.store_c = store_xh2 ? 1 : 0; t.D = tl->D; t.inverse = tl->inverse;
        t.ld_in = tl->ld_in; t.ld_out = tl->ld_out; t.ld_out2 = tl->ld_out2;
        t.W3f = tl->W3f; t.w_net = tl->w_net; t.b3f = tl->b3f; t.b_net = tl->b_net;
        t.in = tl->in; t.cdev = tl->cdev; t.Wmat = tl->Wmat;
        t.out = tl->out; t.out2 = tl->out2; t.logdet = tl->logdet; t.s_out = tl->s_out;
        if (tail_done) *tail_done = true;
      }
      return tc_gemm(f, st);
    }
    g.C = w.u; g.ldc = N; g.sC = B * Hm;
    NF_TRY(linear_nt(m.prec, g, st));
    LnFwdP p{};
    p.u = w.u; p.u_net = B * Hm;
    p.bias = bias; p.bias_net = bias_net;
    p.xh = xh; p.xh_net = xh_net;
    p.rstd = r; p.rstd_net = r_net;
    p.mask = mk; p.mask_net = m_net; p.mw = mw;
    p.N = N; p.nets = 2; p.B = B; p.eps = m.eps; p.fast_var = m.fastvar;
    return lrelu_ln_fwd(p, st);
  };
  {  // layer 1: [x_cond | y] W1^T + b1
    TcGemmP g;
    g.A0 = xc; g.lda0 = ldxc; g.K0 = m.dc;
    // y == NULL: the conditioning input is identically zero (train_auto_NF_rl.py:328-339 evaluates the flow 1 M times
    // with cond = 0): the first Linear reduces to x_cond W1[:, :dc]^T + b1, the y segment of the reduction is skipped
    const bool has_y = m.R > 0 && y != nullptr;
    if (has_y) { g.A1 = y; g.lda1 = m.R; g.K1 = m.R; }
    // W1p: W1 with the y columns moved to the 16-byte aligned column dcp (TMA operand)
    g.B0 = kb + kl.net0 + kl.W1p; g.ldb0 = kl.K1p; g.sB0 = kl.net_stride;
    if (has_y) { g.B1 = g.B0 + kl.dcp; g.ldb1 = kl.K1p; g.sB1 = kl.net_stride; }
    if (kl.hi_off) { g.B0hi = g.B0 + kl.hi_off; g.B0lo = g.B0 + kl.lo_off; if (has_y) { g.B1hi = g.B1 + kl.hi_off; g.B1lo = g.B1 + kl.lo_off; } }
    NF_TRY(linear_lrelu_ln(g, m.H1, pb + pl.slot_off[NF_P_B1], pl.net_stride, xh1, xh1_net, r1, m1, mw1, nullptr));
  }
  {  // layer 2: x_hat1 W2f^T + b2f
    TcGemmP g;
    g.A0 = xh1; g.lda0 = m.H1; g.K0 = m.H1; g.sA0 = xh1_net;
    g.B0 = kb + kl.net0 + kl.W2f; g.ldb0 = m.H1; g.sB0 = kl.net_stride;
    if (kl.hi_off) { g.B0hi = g.B0 + kl.hi_off; g.B0lo = g.B0 + kl.lo_off; }
    NF_TRY(linear_lrelu_ln(g, m.C, kb + kl.net0 + kl.b2f, kl.net_stride, xh2, xh2_net, r2, m2, mw2, tail));
  }
  if (with_last) {  // o = x_hat2 W3f^T  (N = D/2: FFMA unless the flow is wide)
    TcGemmP g;
    g.A0 = xh2; g.lda0 = m.C; g.K0 = m.C; g.sA0 = xh2_net;
    g.B0 = kb + kl.net0 + kl.W3f; g.ldb0 = m.C; g.sB0 = kl.net_stride;
    g.C = w.o; g.ldc = m.dt; g.sC = B * m.dt;
    g.M = (int)B; g.N = m.dt; g.batch = 2;
    NF_TRY(linear_nt(m.prec, g, st));
  }
  return 0;
}

// The persistent whole-stack kernel (flow_stack.cuh) replaces the per-layer launches of a forward / inverse pass when
// the flow is narrow (D <= 9), both hidden widths are equal multiples of 128 and the grid is small enough to be
// latency-bound (at most two waves of CTAs; larger batches keep the per-layer kernels, whose epilogues overlap with
// other tiles' MMAs).
static bool stack_params(const Dims& m, const NfParamLayout& pl, const PackedLayout& kl, const float* params,
                         const float* packed, int64_t B, StackP* out) {
  const int mode = option(NF_OPT_STACK);
  if (mode == 0 || m.prec != NF_PREC_TF32X3 || m.H1 != m.C || !kl.hi_off) return false;
  StackP p;
  p.B = B; p.D = m.D; p.H = m.C; p.R = m.R; p.n = m.n; p.eps = m.eps; p.fast_var = m.fastvar;
  p.raw_hi = (m.dc + m.R >= 512 || m.C >= 512) ? 0 : 1;
  p.W1p = packed + kl.net0 + kl.W1p; p.K1p = kl.K1p; p.dcp = kl.dcp;
  p.W2f = packed + kl.net0 + kl.W2f;
  p.w_net = kl.net_stride; p.w_blk = kl.block_stride; p.hi_off = kl.hi_off; p.lo_off = kl.lo_off;
  p.b1 = params + pl.slot_off[NF_P_B1]; p.b1_net = pl.net_stride; p.b1_blk = pl.block_stride;
  p.b2f = packed + kl.net0 + kl.b2f; p.W3f = packed + kl.net0 + kl.W3f; p.b3f = packed + kl.net0 + kl.b3f;
  p.cdev = packed + kl.logdet_c;
  if (mode != 1 && cdiv(B, 128) * 2 * (m.C / 128) > 2 * 148) return false;
  *out = p;
  return true;
}

}  // namespace nf

using namespace nf;

// =============================================================================================
// C ABI
// =============================================================================================
extern "C" {

int nf_abi_version(void) { return NF_ABI_VERSION; }

int nf_set_option(const char* name, int value) {
  NF_CHECK_ARG(name != nullptr, NF_E_BADARG, "option name is NULL");
  opt_init();
  for (int i = 0; i < NF_OPT_COUNT; ++i) {
    if (strcmp(name, kOpts[i].name) == 0) {
      g_opt[i] = value;
      return nf_graphs_clear();   // cached launch graphs embed the old choice
    }
  }
  return set_err(NF_E_BADARG, "unknown option '%s'", name);
}

int nf_get_option(const char* name) {
  opt_init();
  for (int i = 0; name && i < NF_OPT_COUNT; ++i)
    if (strcmp(name, kOpts[i].name) == 0) return g_opt[i];
  return -9999;
}
const char* nf_last_error(void) { return err_buf(); }
int64_t nf_launch_count(int reset) {
  int64_t v = launch_counter();
  if (reset) launch_counter() = 0;
  return v;
}

int nf_param_layout(const NfFlowDesc* desc, NfParamLayout* out) {
  NF_TRY(check_desc(desc));
  NF_CHECK_ARG(out != nullptr, NF_E_BADARG, "out is NULL");
  param_layout(dims_of(desc), out);
  return 0;
}

int64_t nf_flow_packed_bytes(const NfFlowDesc* desc) {
  if (check_desc(desc)) return -1;
  return packed_layout(dims_of(desc)).total * (int64_t)sizeof(float);
}

int64_t nf_flow_stash_bytes(const NfFlowDesc* desc, int64_t B) {
  if (check_desc(desc) || B < 0) return -1;
  return stash_layout(dims_of(desc), B).total * (int64_t)sizeof(float);
}

int64_t nf_flow_workspace_bytes(const NfFlowDesc* desc, int64_t B) {
  if (check_desc(desc) || B < 0) return -1;
  const Dims m = dims_of(desc);
  int64_t t = carve_fwd(m, B, nullptr).total;
  t = std::max(t, carve_bwd(m, B, nullptr).total);
  t = std::max(t, carve_prep(m, nullptr).total);
  return (t + 64) * (int64_t)sizeof(float);
}

int nf_flow_prepare(const NfFlowDesc* desc, const float* params, void* packed_v, int what, void* workspace,
                    int64_t ws_bytes, void* stream) {
  NF_TRY(check_desc(desc));
  NF_TRY(check_ptr(params, "params", true));
  NF_TRY(check_ptr(packed_v, "packed", true));
  const Dims m = dims_of(desc);
  NfParamLayout pl;
  param_layout(m, &pl);
  const PackedLayout kl = packed_layout(m);
  cudaStream_t user_st = reinterpret_cast<cudaStream_t>(stream);
  float* packed = reinterpret_cast<float*>(packed_v);
  const int64_t DD = (int64_t)m.D * m.D;
  if (what & NF_PREP_INV) {
    NF_TRY(check_ptr(workspace, "workspace", true));
    const PrepWs w = carve_prep(m, workspace);
    NF_CHECK_ARG(ws_bytes >= w.total * (int64_t)sizeof(float), NF_E_WORKSPACE, "workspace too small: %lld < %lld",
                 (long long)ws_bytes, (long long)(w.total * sizeof(float)));
  }
  GraphKey key;
  key.op = 1; key.desc = *desc; key.flags = what; key.ws_bytes = ws_bytes;
  key.ptr[0] = params; key.ptr[1] = packed_v; key.ptr[2] = workspace;
  auto body = [&](cudaStream_t st) -> int {
  // forward-direction tables of a small flow dimension: one kernel, no dependent launches (elementwise.cu)
  if (!(what & NF_PREP_INV) && m.D <= 32 && option(NF_OPT_PREP_FUSED)) return prepare_fwd_fused(params, pl, m, packed, kl, st);
  // three independent chains (PLU tables | folded LayerNorm + its transpose | first-layer packings) run as parallel
  // branches; the hi/lo split of all tables joins them
  SideStreams* ss = side_streams();
  cudaStream_t s1 = ss ? ss->s1 : st, s2 = ss ? ss->s2 : st;
  if (ss) {
    NF_CUDA(cudaEventRecord(ss->evA, st));
    NF_CUDA(cudaStreamWaitEvent(s1, ss->evA, 0));
    NF_CUDA(cudaStreamWaitEvent(s2, ss->evA, 0));
  }
  if (m.D <= 32) {
    NF_TRY(plu_build_small(params, pl, m, packed, kl, st));   // mask, PL, W, W^T, log|det| in one kernel
  } else {
    NF_TRY(plu_mask(params, pl, m, packed, kl, st));
    {  // PL = P @ Lh
      GemmP g;
      g.A0 = params + pl.slot_off[NF_P_P]; g.lda0 = m.D; g.K0 = m.D; g.sA0 = pl.block_stride;
      g.B0 = packed + kl.Lh; g.ldb0 = m.D; g.sB0 = kl.block_stride;
      g.C = packed + kl.PL; g.ldc = m.D; g.sC = kl.block_stride;
      g.M = m.D; g.N = m.D; g.batch = m.n;
      NF_TRY(simt_gemm(g, st));
      // W = PL @ Uh
      g.A0 = packed + kl.PL; g.sA0 = kl.block_stride;
      g.B0 = packed + kl.Uh;
      g.C = packed + kl.W;
      NF_TRY(simt_gemm(g, st));
      NF_TRY(transpose2d(packed + kl.WT, m.D, kl.block_stride, 0, packed + kl.W, m.D, kl.block_stride, 0, m.D, m.D, m.n, 1, st));
    }
  }
  NF_TRY(fold_ln_affine(params, pl, m, packed, kl, s1));
  // K-major B operands of the data-gradient GEMMs: W2f^T (H1,C) and W1^T (K1,H1), per net and block
  NF_TRY(transpose2d(packed + kl.net0 + kl.W2fT, m.C, kl.net_stride, kl.block_stride, packed + kl.net0 + kl.W2f,
                     m.H1, kl.net_stride, kl.block_stride, m.C, m.H1, 2, m.n, s1));
  NF_TRY(transpose2d(packed + kl.net0 + kl.W1T, m.H1, kl.net_stride, kl.block_stride,
                     params + pl.slot_off[NF_P_W1], m.K1, pl.net_stride, pl.block_stride, m.H1, m.K1, 2, m.n, s2));
  NF_TRY(pack_w1(params, pl, m, packed, kl, s2));
  if (what & NF_PREP_INV) {
    const PrepWs w = carve_prep(m, workspace);
    NF_TRY(plu_tri_inverse(packed, kl, m, w.uinv, w.linv, st));
    GemmP g;  // T = Uinv @ Linv ; Winv = T @ P_inv   (torch_fcnf.py:54)
    g.A0 = w.uinv; g.lda0 = m.D; g.K0 = m.D; g.sA0 = DD;
    g.B0 = w.linv; g.ldb0 = m.D; g.sB0 = DD;
    g.C = w.T; g.ldc = m.D; g.sC = DD;
    g.M = m.D; g.N = m.D; g.batch = m.n;
    NF_TRY(simt_gemm(g, st));
    g.A0 = w.T;
    g.B0 = params + pl.slot_off[NF_P_PINV]; g.sB0 = pl.block_stride;
    g.C = packed + kl.Winv; g.sC = kl.block_stride;
    NF_TRY(simt_gemm(g, st));
    NF_TRY(transpose2d(packed + kl.WinvT, m.D, kl.block_stride, 0, packed + kl.Winv, m.D, kl.block_stride, 0, m.D, m.D, m.n, 1, st));
  }
  if (ss) {
    NF_CUDA(cudaEventRecord(ss->done1[0], s1));
    NF_CUDA(cudaEventRecord(ss->done2[0], s2));
    NF_CUDA(cudaStreamWaitEvent(st, ss->done1[0], 0));
    NF_CUDA(cudaStreamWaitEvent(st, ss->done2[0], 0));
  }
  if (kl.hi_off) NF_TRY(split_tables(packed, kl.hi_off, kl.lo_off, (int64_t)m.n * kl.block_stride, st));
  return 0;
  };
  return run_graphed(key, user_st, body);
}

int nf_flow_forward(const NfFlowDesc* desc, const float* params, const void* packed_v, const float* x,
                    const float* y, int64_t B, float* z, float* logdet, float* outputs, void* stash_v,
                    void* workspace, int64_t ws_bytes, void* stream) {
  NF_TRY(check_desc(desc));
  NF_CHECK_ARG(B >= 0 && B < (int64_t(1) << 30), NF_E_BADARG, "B=%lld out of range", (long long)B);
  NF_TRY(check_ptr(params, "params", true));
  NF_TRY(check_ptr(packed_v, "packed", true));
  if (B == 0) return 0;
  NF_TRY(check_ptr(x, "x", true));
  NF_TRY(check_ptr(y, "y", false));     // NULL: y == 0 (fast path, see conditioners())
  NF_TRY(check_ptr(z, "z", true));
  NF_TRY(check_ptr(logdet, "logdet", true));
  NF_TRY(check_ptr(outputs, "outputs", false));
  NF_TRY(check_ptr(stash_v, "stash", false));
  NF_TRY(check_ptr(workspace, "workspace", true));
  const Dims m = dims_of(desc);
  NfParamLayout pl;
  param_layout(m, &pl);
  const PackedLayout kl = packed_layout(m);
  const FwdWs w = carve_fwd(m, B, workspace);
  NF_CHECK_ARG(ws_bytes >= w.total * (int64_t)sizeof(float), NF_E_WORKSPACE, "workspace too small: %lld < %lld",
               (long long)ws_bytes, (long long)(w.total * sizeof(float)));
  cudaStream_t user_st = reinterpret_cast<cudaStream_t>(stream);
  const float* packed = reinterpret_cast<const float*>(packed_v);
  float* stash = reinterpret_cast<float*>(stash_v);
  const StashLayout sl = stash_layout(m, B);
  const int64_t BD = B * m.D;

  GraphKey key;
  key.op = 2; key.desc = *desc; key.B = B; key.ws_bytes = ws_bytes;
  key.ptr[0] = params; key.ptr[1] = packed_v; key.ptr[7] = stash_v; key.ptr[8] = workspace;
  key.flags = (outputs ? 1 : 0) | (stash ? 2 : 0) | (y ? 4 : 0);
  // inputs -> persistent staging (the stash keeps x and y for backward anyway)
  float* x_in = stash ? stash + sl.xs : w.io_x;
  float* y_in = stash ? stash + sl.y : w.io_y;
  float* const logdet_user = logdet;
  float* const z_user = z;
  float* const outputs_user = outputs;
  {
    CopySegs c;
    c.add(x_in, x, BD);
    c.add(y_in, y, B * m.R);
    NF_TRY(copy_segments(c, user_st));
  }
  if (y) y = y_in;
  logdet = w.io_ld;
  z = w.io_z;
  if (outputs && !stash) outputs = w.io_outs;
  const float* z_src = nullptr;
  auto body = [&](cudaStream_t st) -> int {
  NF_CUDA(cudaMemsetAsync(logdet, 0, B * sizeof(float), st));
  const float* cur = x_in;
  const bool fused = use_row_fused(m);
  StackP sp;
  if (fused && stack_params(m, pl, kl, params, packed, B, &sp)) {
    sp.inverse = 0;
    sp.cur = w.xp; sp.cur_w = w.xp; sp.ld_cur = w.Dp; sp.y = y;
    sp.Wmat = packed + kl.W;
    if (stash) {
      float* sb = stash + sl.blocks;
      sp.xh1 = sb + sl.xh1; sp.xh2 = sb + sl.xh2; sp.xh_net = sl.net_stride; sp.xh_blk = sl.block_stride;
      sp.store_xh2 = 1;
      sp.rstd1 = sb + sl.r1; sp.rstd2 = sb + sl.r2;
      sp.mask1 = reinterpret_cast<uint32_t*>(sb + sl.m1); sp.mask2 = reinterpret_cast<uint32_t*>(sb + sl.m2);
      sp.mw = sl.mw1;
      sp.out = stash + sl.xs + BD; sp.out_blk = BD;
      sp.s_out = stash + sl.ss; sp.s_blk = B * m.dt;
    } else {
      sp.xh1 = w.xh1; sp.xh2 = w.xh2; sp.xh_net = B * m.C; sp.xh_blk = 0;
      if (outputs) { sp.out = outputs; sp.out_blk = BD; }
      else { sp.out = z; sp.out_blk = 0; }
    }
    sp.logdet = logdet;
    if (flow_stack_supported(sp)) {
      GemmP g;   // xp = x @ W_0   (torch_fcnf.py:40); later blocks: the previous block's tail
      g.A0 = cur; g.lda0 = m.D; g.K0 = m.D;
      g.B0 = packed + kl.W; g.ldb0 = m.D;
      g.C = w.xp; g.ldc = w.Dp; g.M = (int)B; g.N = m.D;
      NF_TRY(simt_gemm(g, st));
      return flow_stack(sp, st);
    }
  }
  for (int i = 0; i < m.n; ++i) {
    const float* pb = params + (int64_t)i * pl.block_stride;
    const float* kb = packed + (int64_t)i * kl.block_stride;
    if (!fused || i == 0)   // xp = cur @ W   (torch_fcnf.py:40); fused: done by the previous block's tail
      NF_TRY(plu_matmul(m, kl, cur, m.D, kb + kl.WT, w.xp, w.Dp, B, st));
    float *xh1 = w.xh1, *xh2 = w.xh2, *r1 = nullptr, *r2 = nullptr;
    uint32_t *m1 = nullptr, *m2 = nullptr;
    int64_t xh1_net = B * m.H1, xh2_net = B * m.C, r_net = 0, m_net = 0;
    if (stash) {
      float* sb = stash + sl.blocks + (int64_t)i * sl.block_stride;
      xh1 = sb + sl.xh1; xh2 = sb + sl.xh2; r1 = sb + sl.r1; r2 = sb + sl.r2;
      m1 = reinterpret_cast<uint32_t*>(sb + sl.m1); m2 = reinterpret_cast<uint32_t*>(sb + sl.m2);
      xh1_net = xh2_net = r_net = m_net = sl.net_stride;
    }
    float* next;
    if (stash) next = stash + sl.xs + (int64_t)(i + 1) * BD;
    else if (outputs) next = outputs + (int64_t)i * BD;
    else if (i == m.n - 1) next = z;
    else next = (i & 1) ? w.pp1 : w.pp0;   // dense (B,D): only the PLU matmul reads it
    float* s_out = stash ? stash + sl.ss + (int64_t)i * B * m.dt : nullptr;
    TailP t{};
    t.xh2 = xh2; t.xh2_net = xh2_net;
    t.W3f = kb + kl.net0 + kl.W3f; t.w_net = kl.net_stride;
    t.b3f = kb + kl.net0 + kl.b3f; t.b_net = kl.net_stride;
    t.in = w.xp; t.ld_in = w.Dp;
    t.cdev = packed + kl.logdet_c + i;
    t.Wmat = (i + 1 < m.n) ? kb + kl.block_stride + kl.W : nullptr;   // next block's W
    t.out = next; t.ld_out = m.D;
    t.out2 = w.xp; t.ld_out2 = w.Dp;                                   // in place: a warp owns its row
    t.logdet = logdet; t.s_out = s_out;
    t.B = B; t.D = m.D; t.C = m.C; t.inverse = 0;
    bool tail_done = false;
    NF_TRY(conditioners(m, pl, kl, pb, kb, w.xp, w.Dp, y, B, w, xh1, xh1_net, xh2, xh2_net, r1, r2, r_net, m1, m2,
                        m_net, sl.mw1, sl.mw2, !fused, st, fused ? &t : nullptr, &tail_done, stash != nullptr));
    if (fused) {
      if (!tail_done) NF_TRY(row_v2_supported(m.D, m.C, 4) ? row_tail2(t, st) : row_tail(t, st));
    } else {
      NF_TRY(affine_fwd(w.xp, w.Dp, w.o + B * m.dt, w.o, kb + kl.net0 + kl.net_stride + kl.b3f,
                        kb + kl.net0 + kl.b3f, packed + kl.logdet_c + i, 0.f, B, m.D, next, m.D, logdet, s_out, st));
    }
    cur = next;
  }
  return 0;
  };
  NF_TRY(run_graphed(key, user_st, body));
  // outputs -> the caller's buffers (the last block wrote z into the stash, the outputs array or io_z)
  if (stash) z_src = stash + sl.xs + (int64_t)m.n * BD;
  else if (outputs) z_src = outputs + (int64_t)(m.n - 1) * BD;
  else z_src = z;
  {
    CopySegs c;
    c.add(z_user, z_src, BD);
    c.add(logdet_user, logdet, B);
    if (outputs_user) c.add(outputs_user, stash ? stash + sl.xs + BD : outputs, (int64_t)m.n * BD);
    NF_TRY(copy_segments(c, user_st));
  }
  return 0;
}


int nf_flow_inverse(const NfFlowDesc* desc, const float* params, const void* packed_v, const float* z,
                    const float* y, int64_t B, float* x, float* logdet, float* seq, void* workspace,
                    int64_t ws_bytes, void* stream) {
  NF_TRY(check_desc(desc));
  NF_CHECK_ARG(B >= 0 && B < (int64_t(1) << 30), NF_E_BADARG, "B=%lld out of range", (long long)B);
  NF_TRY(check_ptr(params, "params", true));
  NF_TRY(check_ptr(packed_v, "packed", true));
  if (B == 0) return 0;
  NF_TRY(check_ptr(z, "z", true));
  NF_TRY(check_ptr(y, "y", false));     // NULL: y == 0
  NF_TRY(check_ptr(x, "x", true));
  NF_TRY(check_ptr(logdet, "logdet", false));
  NF_TRY(check_ptr(seq, "seq", false));
  NF_TRY(check_ptr(workspace, "workspace", true));
  const Dims m = dims_of(desc);
  NfParamLayout pl;
  param_layout(m, &pl);
  const PackedLayout kl = packed_layout(m);
  const FwdWs w = carve_fwd(m, B, workspace);
  NF_CHECK_ARG(ws_bytes >= w.total * (int64_t)sizeof(float), NF_E_WORKSPACE, "workspace too small: %lld < %lld",
               (long long)ws_bytes, (long long)(w.total * sizeof(float)));
  cudaStream_t user_st = reinterpret_cast<cudaStream_t>(stream);
  const float* packed = reinterpret_cast<const float*>(packed_v);
  const int64_t BD = B * m.D;

  GraphKey key;
  key.op = 3; key.desc = *desc; key.B = B; key.ws_bytes = ws_bytes;
  key.ptr[0] = params; key.ptr[1] = packed_v; key.ptr[6] = seq; key.ptr[7] = workspace;
  key.flags = (logdet ? 1 : 0) | (y ? 2 : 0);
  float* const x_user = x;
  float* const logdet_user = logdet;
  {
    CopySegs c;
    c.add(w.io_x, z, BD);
    c.add(w.io_y, y, B * m.R);
    NF_TRY(copy_segments(c, user_st));
  }
  z = w.io_x; if (y) y = w.io_y; x = w.io_z;
  if (logdet) logdet = w.io_ld;
  auto body = [&](cudaStream_t st) -> int {
  if (logdet) NF_CUDA(cudaMemsetAsync(logdet, 0, B * sizeof(float), st));
  if (seq) NF_CUDA(cudaMemcpyAsync(seq, z, BD * sizeof(float), cudaMemcpyDeviceToDevice, st));
  // running buffer with a 16-byte row pitch (x_cond is a TMA operand of the first Linear)
  const float* cur = z;
  int ldcur = m.D;
  if (w.Dp != m.D) {
    NF_TRY(copy2d(w.pp1, w.Dp, z, m.D, B, m.D, st));
    cur = w.pp1; ldcur = w.Dp;
  }
  int step = 0;
  const bool fused = use_row_fused(m);
  StackP sp;
  if (fused && !seq && stack_params(m, pl, kl, params, packed, B, &sp)) {
    sp.inverse = 1;
    sp.cur = cur; sp.cur_w = const_cast<float*>(cur); sp.ld_cur = ldcur; sp.y = y;   // (a workspace copy of z: updated in place)
    sp.Wmat = packed + kl.Winv;
    sp.xh1 = w.xh1; sp.xh2 = w.xh2; sp.xh_net = B * m.C; sp.xh_blk = 0;
    sp.logdet = logdet;
    sp.final_out = x; sp.ld_final = m.D;
    if (flow_stack_supported(sp)) return flow_stack(sp, st);
  }
  for (int i = m.n - 1; i >= 0; --i, ++step) {
    const float* pb = params + (int64_t)i * pl.block_stride;
    const float* kb = packed + (int64_t)i * kl.block_stride;
    float* next;
    int ldnext;
    if (i == 0 && !seq) { next = x; ldnext = m.D; }
    else if (seq && w.Dp == m.D) { next = seq + (int64_t)(step + 1) * BD; ldnext = m.D; }
    else { next = (step & 1) ? w.pp1 : w.pp0; ldnext = w.Dp; }
    // affine first (torch_fcnf.py:112-113), then x = xp @ W^-1 (:115 -> :55)
    TailP t{};
    t.xh2 = w.xh2; t.xh2_net = B * m.C;
    t.W3f = kb + kl.net0 + kl.W3f; t.w_net = kl.net_stride;
    t.b3f = kb + kl.net0 + kl.b3f; t.b_net = kl.net_stride;
    t.in = cur; t.ld_in = ldcur;
    t.cdev = packed + kl.logdet_c + i;
    t.Wmat = kb + kl.Winv;
    t.out2 = next; t.ld_out2 = ldnext;
    t.logdet = logdet;
    t.B = B; t.D = m.D; t.C = m.C; t.inverse = 1;
    bool tail_done = false;
    NF_TRY(conditioners(m, pl, kl, pb, kb, cur, ldcur, y, B, w, w.xh1, B * m.H1, w.xh2, B * m.C, nullptr, nullptr, 0,
                        nullptr, nullptr, 0, 0, 0, !fused, st, fused ? &t : nullptr, &tail_done, false));
    if (fused) {
      if (!tail_done) NF_TRY(row_v2_supported(m.D, m.C, 4) ? row_tail2(t, st) : row_tail(t, st));
    } else {
      NF_TRY(affine_inv(cur, ldcur, w.o + B * m.dt, w.o, kb + kl.net0 + kl.net_stride + kl.b3f, kb + kl.net0 + kl.b3f,
                        packed + kl.logdet_c + i, 0.f, B, m.D, w.xp, w.Dp, logdet, st));
      NF_TRY(plu_matmul(m, kl, w.xp, w.Dp, kb + kl.WinvT, next, ldnext, B, st));
    }
    if (seq && w.Dp != m.D) NF_TRY(copy2d(seq + (int64_t)(step + 1) * BD, m.D, next, ldnext, B, m.D, st));
    cur = next; ldcur = ldnext;
  }
  if (seq) NF_CUDA(cudaMemcpyAsync(x, seq + (int64_t)m.n * BD, BD * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return 0;
  };
  NF_TRY(run_graphed(key, user_st, body));
  {
    CopySegs c;
    c.add(x_user, x, BD);
    c.add(logdet_user, logdet, B);
    NF_TRY(copy_segments(c, user_st));
  }
  return 0;
}

// dp_bucket > 0: data-parallel backward -- the call's block range is processed in sub-ranges of dp_bucket blocks from
// the top block down and the finished gradient slice of every sub-range is averaged over the ranks (NCCL, the
// library's communicator) on a third side stream while the next sub-range computes; everything is part of the one
// launch graph of the call.
static int backward_impl(const NfFlowDesc* desc, const float* params, const void* packed_v, const void* stash_v,
                         const float* y, int64_t B, const float* dz, const float* dlogdet, float* dx, float* dy,
                         float* dparams, int32_t block_begin, int32_t block_end, int32_t dp_bucket, void* workspace,
                         int64_t ws_bytes, void* stream) {
  NF_TRY(check_desc(desc));
  NF_CHECK_ARG(block_begin >= 0 && block_begin < block_end && block_end <= desc->n_blocks, NF_E_BADARG,
               "block range [%d, %d) outside [0, %d)", block_begin, block_end, desc->n_blocks);
  const bool first_range = block_end == desc->n_blocks, last_range = block_begin == 0;
  const int nr_all = block_end - block_begin;
  if (dp_bucket > 0) NF_CHECK_ARG(dp_world() > 1 && dparams != nullptr, NF_E_BADARG,
                                  "data-parallel backward needs nf_dp_init (world > 1) and a gradient arena");
  NF_CHECK_ARG(B >= 0 && B < (int64_t(1) << 30), NF_E_BADARG, "B=%lld out of range", (long long)B);
  NF_TRY(check_ptr(params, "params", true));
  NF_TRY(check_ptr(packed_v, "packed", true));
  NF_TRY(check_ptr(dparams, "dparams", false));
  const Dims m = dims_of(desc);
  NfParamLayout pl;
  param_layout(m, &pl);
  cudaStream_t user_st = reinterpret_cast<cudaStream_t>(stream);
  if (B == 0) {
    if (dparams) NF_TRY(fill_zero(dparams + (int64_t)block_begin * pl.block_stride, (int64_t)nr_all * pl.block_stride, user_st));
    // (an empty shard still takes part in the collective)
    if (dp_bucket > 0) NF_TRY(dp_allreduce(dparams + (int64_t)block_begin * pl.block_stride, (int64_t)nr_all * pl.block_stride, true, user_st));
    return 0;
  }
  NF_TRY(check_ptr(stash_v, "stash", true));
  NF_TRY(check_ptr(y, "y", false));       // NULL: the forward ran with y == 0 (no dy, no y columns in dW1)
  NF_TRY(check_ptr(dz, "dz", false));
  NF_TRY(check_ptr(dlogdet, "dlogdet", false));
  NF_TRY(check_ptr(dx, "dx", false));
  NF_TRY(check_ptr(dy, "dy", false));
  NF_TRY(check_ptr(workspace, "workspace", true));
  const PackedLayout kl = packed_layout(m);
  const StashLayout sl = stash_layout(m, B);
  const BwdWs w = carve_bwd(m, B, workspace);
  NF_CHECK_ARG(ws_bytes >= w.total * (int64_t)sizeof(float), NF_E_WORKSPACE, "workspace too small: %lld < %lld",
               (long long)ws_bytes, (long long)(w.total * sizeof(float)));
  const float* packed = reinterpret_cast<const float*>(packed_v);
  const float* stash = reinterpret_cast<const float*>(stash_v);
  const int64_t BD = B * m.D, Hm = std::max(m.H1, m.C), DD = (int64_t)m.D * m.D;

  GraphKey key;
  key.op = 4; key.desc = *desc; key.B = B; key.ws_bytes = ws_bytes;
  key.ptr[0] = params; key.ptr[1] = packed_v; key.ptr[2] = stash_v; key.ptr[8] = dparams; key.ptr[9] = workspace;
  const bool has_y = y != nullptr && m.R > 0;
  if (!has_y && dy) {   // y == 0 was not an input: its gradient is reported as zero
    NF_TRY(fill_zero(dy, B * m.R, user_st));
    dy = nullptr;
  }
  key.flags = (dz ? 1 : 0) | (dlogdet ? 2 : 0) | (dx ? 4 : 0) | (dy ? 8 : 0) | (has_y ? 16 : 0) | (block_begin << 8) | (block_end << 20);
  key.ptr[10] = reinterpret_cast<const void*>((uintptr_t)(dp_bucket > 0 ? dp_bucket : 0));
  // cotangents in, dx / dy out through the workspace; y is the copy forward left in the stash
  float* const dx_user = dx;
  float* const dy_user = dy;
  {
    CopySegs c;
    if (first_range) c.add(w.io_dz, dz, BD);       // later ranges continue from the workspace
    c.add(w.io_dld, dlogdet, B);
    NF_TRY(copy_segments(c, user_st));
  }
  if (dz) dz = w.io_dz;
  if (dlogdet) dlogdet = w.io_dld;
  if (dx) dx = w.io_dx;
  if (dy) dy = w.io_dy;
  y = stash + sl.y;
  auto body = [&](cudaStream_t st) -> int {
  const bool fused = use_row_fused(m);
  const bool v2 = fused && row_v2_supported(m.D, m.C, m.H1);
  const int whatif = option(NF_OPT_WHATIF);
  SideStreams* ss = (v2 && (dparams || dy)) ? side_streams() : nullptr;
  SideStreams* sar = dp_bucket > 0 ? side_streams(true) : nullptr;     // all-reduce stream
  NF_CHECK_ARG(dp_bucket <= 0 || sar != nullptr, NF_E_UNSUPPORTED, "could not create the all-reduce stream");
  cudaStream_t s1 = ss ? ss->s1 : st, s2 = ss ? ss->s2 : st;
  if (ss) {   // what only the side streams consume is prepared there, off the chain: dy (written by the dy GEMMs on s2) and
              // the sum of the log-det cotangents (read by the PLU gradient kernel at the very end)
    NF_CUDA(cudaEventRecord(ss->evB, st));
    NF_CUDA(cudaStreamWaitEvent(s1, ss->evB, 0));
    NF_CUDA(cudaStreamWaitEvent(s2, ss->evB, 0));
  }
  if (dparams) NF_TRY(fill_zero(dparams + (int64_t)block_begin * pl.block_stride, (int64_t)nr_all * pl.block_stride, st));
  if (dy && first_range) NF_TRY(fill_zero(dy, B * m.R, s2));
  if (dparams) {
    NF_TRY(fill_zero(w.dWs + (int64_t)block_begin * DD, nr_all * DD, st));
    if (dlogdet) NF_TRY(sum_vector(dlogdet, B, w.sumdld, s1));
    else NF_TRY(fill_zero(w.sumdld, 4, s1));
  }
  // block i < n-1 reads the gradient block i+1 left in d0 / d1 (alternating from the top block down), also across
  // range calls; the top block reads dz (NULL means zero, handled inside the kernels)
  int pp = (m.n - block_end) & 1;
  const float* gcur = first_range ? dz : (pp ? w.d0 : w.d1);
  struct { float *dxp, *dout, *gA, *gB; } w2[2] = {{w.dxp, w.dout, w.gA, w.gB}, {w.dxp2, w.dout2, w.gA2, w.gB2}};
  // deferred, batched conditioner weight gradients (see defer_fits): both GEMMs must be the TS kernels that take the
  // bias / x_cond reductions along
  bool defer = false, ln1_fused_any = false;
  // auto: only where it was measured to win -- the data-gradient chain fills the machine about once per kernel (C2a:
  // 128 CTAs), so per-block weight gradients on the side streams compete with it: train step 0.658 -> 0.628 ms.  Smaller
  // grids leave idle SMs that absorb the side work for free (C1a, B = 256: 1.152 -> 1.202 ms when deferred) and on
  // multi-wave grids the per-block launches fill the tails of the waves (C3, B = 16384: 2.218 -> 2.260 ms).
  const int64_t chain_ctas = cdiv(B, 128) * 2 * cdiv(m.H1, 128);
  const int dw = option(NF_OPT_DEFER_WGRAD);
  if (v2 && dparams && has_y && w.gAall && m.dc <= 16 && (dw == 1 || (dw < 0 && chain_ctas >= 96 && chain_ctas <= 192))) {
    TcGemmTnP e, h;
    e.A = w.gAall; e.lda = m.C; e.sA = B * Hm; e.B = stash + sl.blocks + sl.xh1; e.ldb = m.H1; e.sB = sl.net_stride;
    e.C = dparams; e.ldc = m.H1; e.M = m.C; e.N = m.H1; e.K = B; e.batch = 2;
    h = e; h.A = w.gBall; h.lda = m.H1; h.B = y; h.ldb = m.R; h.sB = 0; h.ldc = m.K1; h.M = m.H1; h.N = m.R;
    defer = rides_ok(m.prec, e) && rides_ok(m.prec, h);
  }
  // with du1 of every block kept, the dy GEMMs of the blocks (M = B, N = R, K = 2 H1 each; C2a: 6 launches of 11 us beside
  // the chain) are ONE launch after the loop whose batch entries accumulate into dy with red.add
  bool defer_dy = false;
  if (defer && dy && has_y && ss && prec_tc(m.prec) && m.R >= 16 && option(NF_OPT_DEFER_DY)) {
    TcGemmP t;
    t.A0 = w.gBall; t.lda0 = m.H1; t.K0 = m.H1; t.A1 = w.gBall + B * Hm; t.lda1 = m.H1; t.K1 = m.H1;
    t.B0 = packed + kl.net0 + kl.W1T + (int64_t)m.dc * m.H1; t.ldb0 = m.H1; t.B1 = t.B0 + kl.net_stride; t.ldb1 = m.H1;
    t.C = dy; t.ldc = m.R; t.M = (int)B; t.N = m.R; t.accumulate = 1;
    defer_dy = tc_gemm_supported(t);
  }
  // ... and what is left of the column reductions (G3 = dout^T x_hat2 with colsum(dout), dW = x_in^T dxp): with dout / dxp
  // of every block kept they are one launch each after the loop (C2a: 12 launches of 5-11 us beside the chain -> 2)
  const bool defer_cols = defer && ss && w.doutAll && option(NF_OPT_DEFER_COLS) != 0;
  std::vector<char> cols_later((size_t)m.n, 0);
  bool any_ar = false;
  for (int sub_end = block_end; sub_end > block_begin;) {
  const int sub_begin = dp_bucket > 0 ? std::max(block_begin, sub_end - dp_bucket) : block_begin;
  const int nr = sub_end - sub_begin;
  bool used1[2] = {false, false}, used2[2] = {false, false};
  for (int i = sub_end - 1; i >= sub_begin; --i) {
    const int par = ss ? ((m.n - 1 - i) & 1) : 0;
    auto wb = w2[par];          // this block's intermediate buffers
    if (defer_cols) { wb.dout = w.doutAll + (int64_t)i * 2 * B * m.dt; wb.dxp = w.dxpAll + (int64_t)i * BD; }
    float* const gA = defer ? w.gAall + (int64_t)i * 2 * B * Hm : wb.gA;   // du2 (deferred weight gradients: kept per block)
    float* const gB = defer ? w.gBall + (int64_t)i * 2 * B * Hm : wb.gB;   // du1
    if (ss) {   // the side work of block i+2 read this buffer set
      if (used1[par]) NF_CUDA(cudaStreamWaitEvent(st, ss->done1[par], 0));
      if (used2[par]) NF_CUDA(cudaStreamWaitEvent(st, ss->done2[par], 0));
    }
    const float* kb = packed + (int64_t)i * kl.block_stride;
    float* dn = dparams ? dparams + (int64_t)i * pl.block_stride : nullptr;
    const float* xin = stash + sl.xs + (int64_t)i * BD;
    const float* zi = stash + sl.xs + (int64_t)(i + 1) * BD;
    const float* si = stash + sl.ss + (int64_t)i * B * m.dt;
    const float* sb = stash + sl.blocks + (int64_t)i * sl.block_stride;
    const uint32_t* m1 = reinterpret_cast<const uint32_t*>(sb + sl.m1);
    const uint32_t* m2 = reinterpret_cast<const uint32_t*>(sb + sl.m2);

    if (fused && v2) {
      // a-d in one row-local kernel: affine backward, data gradient through the last Linear and through
      // LayerNorm + LeakyReLU of layer 2; the skinny parameter gradients are reduced at the end of the block
      BwdHead2P h{};
      h.dz = gcur; h.ld_dz = m.D; h.z = zi; h.s = si; h.dld = dlogdet;
      h.xh2 = sb + sl.xh2; h.xh2_net = sl.net_stride;
      h.rstd = sb + sl.r2; h.rstd_net = sl.net_stride;
      h.mask = m2; h.mask_net = sl.net_stride; h.mw = sl.mw2;
      h.W3f = kb + kl.net0 + kl.W3f; h.w_net = kl.net_stride;
      h.dxp = wb.dxp; h.dout = wb.dout; h.dout_net = B * m.dt;
      h.du2 = gA; h.du2_net = B * Hm;
      h.B = B; h.D = m.D; h.C = m.C;
      NF_TRY(row_bwd_head2(h, st));
      if (ss) { NF_CUDA(cudaEventRecord(ss->evA, st)); NF_CUDA(cudaStreamWaitEvent(s1, ss->evA, 0)); }
    } else if (fused) {
      // a-d in one kernel: affine backward, last-Linear gradients, data gradient through the last
      // Linear and through LayerNorm + LeakyReLU of layer 2
      BwdHeadP h{};
      h.dz = gcur; h.ld_dz = m.D; h.z = zi; h.s = si; h.dld = dlogdet;
      h.xh2 = sb + sl.xh2; h.xh2_net = sl.net_stride;
      h.rstd = sb + sl.r2; h.rstd_net = sl.net_stride;
      h.mask = m2; h.mask_net = sl.net_stride; h.mw = sl.mw2;
      h.W3f = kb + kl.net0 + kl.W3f; h.w_net = kl.net_stride;
      h.dxp = wb.dxp; h.du2 = gA; h.du2_net = B * Hm;
      h.G3 = dn ? dn + pl.slot_off[NF_P_W3] : nullptr;
      h.g3 = dn ? dn + pl.slot_off[NF_P_B3] : nullptr;
      h.gb2 = dn ? dn + pl.slot_off[NF_P_B2] : nullptr;
      h.g_net = pl.net_stride;
      h.B = B; h.D = m.D; h.C = m.C;
      NF_TRY(row_bwd_head(h, st));
    } else {
      // a. affine coupling backward: dout[0] = dt, dout[1] = ds
      NF_TRY(affine_bwd(gcur, zi, si, dlogdet, B, m.D, wb.dxp, wb.dout + B * m.dt, wb.dout,
                        dn ? dn + pl.net_stride + pl.slot_off[NF_P_B3] : nullptr,
                        dn ? dn + pl.slot_off[NF_P_B3] : nullptr, st));
      if (dn) {  // b. G3 = dout^T x_hat2  -> slot W3 (unfolded at the end)
        TcGemmTnP g;
        g.A = wb.dout; g.lda = m.dt; g.sA = B * m.dt;
        g.B = sb + sl.xh2; g.ldb = m.C; g.sB = sl.net_stride;
        g.C = dn + pl.slot_off[NF_P_W3]; g.ldc = m.C; g.sC = pl.net_stride;
        g.M = m.dt; g.N = m.C; g.K = B; g.batch = 2;
        NF_TRY(wgrad_tn(m.prec, g, st));
      }
      {  // c. d x_hat2 = dout W3f
        bool done = false;
        if (prec_tc(m.prec) && m.dt >= 16 && (m.dt & 3) == 0) {
          // tensor-core form: the K-major B operand is W3f^T (C, dt), transposed per block into the workspace (the
          // FFMA GEMM below took 44 us per block on C5 for what is a 67 MB store)
          float* w3t = w.w3t[par];
          NF_TRY(transpose2d(w3t, m.dt, (int64_t)m.C * m.dt, 0, kb + kl.net0 + kl.W3f, m.C, kl.net_stride, 0, m.dt, m.C, 2, 1, st));
          TcGemmP t;
          t.A0 = wb.dout; t.lda0 = m.dt; t.sA0 = B * m.dt; t.K0 = m.dt;
          t.B0 = w3t; t.ldb0 = m.dt; t.sB0 = (int64_t)m.C * m.dt;
          t.C = gA; t.ldc = m.C; t.sC = B * Hm;
          t.M = (int)B; t.N = m.C; t.batch = 2;
          if (m.prec == NF_PREC_TF32) { t.passes = 1; t.raw_hi = 1; }
          if (tc_gemm_supported(t)) { NF_TRY(tc_gemm(t, st)); done = true; }
        }
        if (!done) {
          GemmP g;
          g.A0 = wb.dout; g.lda0 = m.dt; g.sA0 = B * m.dt; g.K0 = m.dt;
          g.B0 = kb + kl.net0 + kl.W3f; g.ldb0 = m.C; g.sB0 = kl.net_stride;
          g.C = gA; g.ldc = m.C; g.sC = B * Hm;
          g.M = (int)B; g.N = m.C; g.batch = 2;
          NF_TRY(simt_gemm(g, st));
        }
      }
      {  // d. through LayerNorm + LeakyReLU (layer 2); column sums -> slot b2
        LnBwdP p{};
        p.g = gA; p.g_net = B * Hm;
        p.xh = sb + sl.xh2; p.xh_net = sl.net_stride;
        p.rstd = sb + sl.r2; p.rstd_net = sl.net_stride;
        p.mask = m2; p.mask_net = sl.net_stride; p.mw = sl.mw2;
        p.colsum = dn ? dn + pl.slot_off[NF_P_B2] : nullptr; p.colsum_net = pl.net_stride;
        p.N = m.C; p.nets = 2; p.B = B;
        NF_TRY(ln_lrelu_bwd(p, st));
      }
    }
    bool fold_e = false, fold_h = false;   // batch-row reductions taken by the weight-gradient GEMMs (tc_gemm.cuh)
    if (dn) {  // e. G2 = du2^T x_hat1 -> slot W2
      TcGemmTnP g;
      g.A = gA; g.lda = m.C; g.sA = B * Hm;
      g.B = sb + sl.xh1; g.ldb = m.H1; g.sB = sl.net_stride;
      g.C = dn + pl.slot_off[NF_P_W2]; g.ldc = m.H1; g.sC = pl.net_stride;
      g.M = m.C; g.N = m.H1; g.K = B; g.batch = 2;
      if (fused && v2 && rides_ok(m.prec, g)) {   // gb2 = colsum(du2) rides on this GEMM's A operand
        g.colsum = dn + pl.slot_off[NF_P_B2]; g.sColsum = pl.net_stride;
        fold_e = true;
      }
      if (!(whatif & 2) && !defer) NF_TRY(wgrad_tn(m.prec, g, s1));
    }
    bool ln1_fused = false;
    {  // f. d x_hat1 = du2 W2f   (K-major B operand = W2f^T)
      TcGemmP g;
      g.A0 = gA; g.lda0 = m.C; g.sA0 = B * Hm; g.K0 = m.C;
      g.B0 = kb + kl.net0 + kl.W2fT; g.ldb0 = m.C; g.sB0 = kl.net_stride;
      g.C = gB; g.ldc = m.H1; g.sC = B * Hm;
      g.M = (int)B; g.N = m.H1; g.batch = 2;
      if (kl.hi_off) { g.B0hi = g.B0 + kl.hi_off; g.B0lo = g.B0 + kl.lo_off; }
      if (v2 && prec_tc(m.prec) && m.C >= 32) {
        // f + g in one kernel: LayerNorm-1 / LeakyReLU-1 backward as the GEMM's row-wise epilogue; the bias
        // gradient (column sums of du1) is taken by the block's column-reduction launch
        TcGemmP f = g;
        if (f.K0 >= 512) f.raw_hi = 0;
        if (m.prec == NF_PREC_TF32) { f.passes = 1; f.raw_hi = 1; f.B0hi = f.B0lo = nullptr; }
        f.epi.kind = TC_EPI_LN_BWD;
        f.epi.xh = sb + sl.xh1; f.epi.ldxh = m.H1; f.epi.sXh = sl.net_stride;
        f.epi.rstd = const_cast<float*>(sb + sl.r1); f.epi.sRstd = sl.net_stride;
        f.epi.mask = const_cast<uint32_t*>(m1); f.epi.sMask = sl.net_stride; f.epi.mw = sl.mw1;
        if (tc_gemm_epi_supported(f)) {
          NF_TRY(tc_gemm(f, st));
          ln1_fused = true;
        }
      }
      if (!ln1_fused) NF_TRY(linear_nt(m.prec, g, st));
    }
    ln1_fused_any = ln1_fused;
    if (!ln1_fused) {  // g. through LayerNorm + LeakyReLU (layer 1); column sums -> slot b1
      LnBwdP p{};
      p.g = gB; p.g_net = B * Hm;
      p.xh = sb + sl.xh1; p.xh_net = sl.net_stride;
      p.rstd = sb + sl.r1; p.rstd_net = sl.net_stride;
      p.mask = m1; p.mask_net = sl.net_stride; p.mw = sl.mw1;
      p.colsum = dn ? dn + pl.slot_off[NF_P_B1] : nullptr; p.colsum_net = pl.net_stride;
      p.N = m.H1; p.nets = 2; p.B = B;
      NF_TRY(ln_lrelu_bwd(p, st));
    }
    if (ss) { NF_CUDA(cudaEventRecord(ss->evB, st)); NF_CUDA(cudaStreamWaitEvent(s2, ss->evB, 0)); }
    if (dn) {  // h. dW1 = du1^T [x_cond | y]   (fused path: the x_cond columns come from the tail kernel)
      TcGemmTnP g;
      g.A = gB; g.lda = m.H1; g.sA = B * Hm;
      g.B = zi; g.ldb = m.D; g.sB = 0;
      g.C = dn + pl.slot_off[NF_P_W1]; g.ldc = m.K1; g.sC = pl.net_stride;
      g.M = m.H1; g.N = m.dc; g.K = B; g.batch = 2;
      if (!fused) NF_TRY(wgrad_tn(m.prec, g, st));
      if (has_y) {
        g.B = y; g.ldb = m.R;
        g.C = dn + pl.slot_off[NF_P_W1] + m.dc; g.N = m.R;
        if (fused && v2 && m.dc <= 16 && rides_ok(m.prec, g)) {
          // G1[:, :dc] = du1^T x_cond and gb1 = colsum(du1) ride on this GEMM's A operand (du1)
          g.xc = zi; g.ldxc = m.D; g.nxc = m.dc;
          g.skinny = dn + pl.slot_off[NF_P_W1]; g.ldsk = m.K1; g.sSk = pl.net_stride;
          if (ln1_fused) { g.colsum = dn + pl.slot_off[NF_P_B1]; g.sColsum = pl.net_stride; }
          fold_h = true;
        }
        if (!(whatif & 2) && !defer) NF_TRY(wgrad_tn(m.prec, g, s2));
      }
    }
    {  // i. d h0 = du1_t W1_t + du1_s W1_s : first dc columns -> dxp, the rest -> dy  (B operand = W1^T)
      TcGemmP g;
      g.A0 = gB; g.lda0 = m.H1; g.K0 = m.H1;
      g.A1 = gB + B * Hm; g.lda1 = m.H1; g.K1 = m.H1;
      g.B0 = kb + kl.net0 + kl.W1T; g.ldb0 = m.H1;
      g.B1 = g.B0 + kl.net_stride; g.ldb1 = m.H1;
      g.C = wb.dxp; g.ldc = m.D; g.M = (int)B; g.N = m.dc; g.accumulate = 1;
      if (!fused) NF_TRY(linear_nt(m.prec, g, st));
      if (dy && has_y) {
        g.B0 += (int64_t)m.dc * m.H1; g.B1 += (int64_t)m.dc * m.H1;
        if (kl.hi_off) { g.B0hi = g.B0 + kl.hi_off; g.B0lo = g.B0 + kl.lo_off; g.B1hi = g.B1 + kl.hi_off; g.B1lo = g.B1 + kl.lo_off; }
        g.C = dy; g.ldc = m.R; g.N = m.R;
        if (!(whatif & 4) && !defer_dy) NF_TRY(linear_nt(m.prec, g, s2));
      }
    }
    float* dst = (i > 0 || dx) ? ((i == 0) ? dx : (pp ? w.d1 : w.d0)) : nullptr;
    if (fused && v2) {
      // i (x_cond part) and k per row, then every reduction over the batch rows of this block in one launch
      BwdTail2P t{};
      t.du1 = gB; t.du1_net = B * Hm;
      t.W1T = kb + kl.net0 + kl.W1T; t.w_net = kl.net_stride;
      t.W = kb + kl.W;
      t.dxp = wb.dxp; t.dxin = dst; t.ld_dxin = m.D;
      t.B = B; t.D = m.D; t.H1 = m.H1;
      NF_TRY(row_bwd_tail2(t, st));
      if (dn) {
        ColJob jobs[3];
        // G3 = dout^T x_hat2 (folded last Linear), gb2 = colsum(du2), g3 = colsum(dout)
        jobs[0].A = sb + sl.xh2; jobs[0].lda = m.C; jobs[0].sA = sl.net_stride; jobs[0].N = m.C;
        jobs[0].coef = wb.dout; jobs[0].ldcf = m.dt; jobs[0].sCf = B * m.dt; jobs[0].m = m.dt;
        jobs[0].out = dn + pl.slot_off[NF_P_W3]; jobs[0].so_i = m.C; jobs[0].so_c = 1; jobs[0].sOut = pl.net_stride;
        if (!fold_e) {
          jobs[0].A2 = gA; jobs[0].lda2 = m.C; jobs[0].sA2 = B * Hm;
          jobs[0].out2 = dn + pl.slot_off[NF_P_B2]; jobs[0].sOut2 = pl.net_stride;
        }
        jobs[0].out3 = dn + pl.slot_off[NF_P_B3]; jobs[0].sOut3 = pl.net_stride;
        jobs[0].batch = 2;
        // G1[:, :dc] = du1^T x_cond   (x_cond = z[:, :dc])
        jobs[1].A = gB; jobs[1].lda = m.H1; jobs[1].sA = B * Hm; jobs[1].N = m.H1;
        jobs[1].coef = zi; jobs[1].ldcf = m.D; jobs[1].sCf = 0; jobs[1].m = m.dc;
        jobs[1].out = dn + pl.slot_off[NF_P_W1]; jobs[1].so_i = 1; jobs[1].so_c = m.K1; jobs[1].sOut = pl.net_stride;
        if (ln1_fused) {   // gb1 = colsum(du1): same array as A, loaded once
          jobs[1].A2 = gB; jobs[1].lda2 = m.H1; jobs[1].sA2 = B * Hm;
          jobs[1].out2 = dn + pl.slot_off[NF_P_B1]; jobs[1].sOut2 = pl.net_stride;
        }
        jobs[1].batch = 2;
        // dW = x_in^T dxp
        jobs[2].A = wb.dxp; jobs[2].lda = m.D; jobs[2].N = m.D;
        jobs[2].coef = xin; jobs[2].ldcf = m.D; jobs[2].m = m.D;
        jobs[2].out = w.dWs + (int64_t)i * DD; jobs[2].so_i = m.D; jobs[2].so_c = 1;
        jobs[2].batch = 1;
        if (whatif & 1) {
        } else if (defer_cols && fold_e && fold_h) {
          cols_later[i] = 1;
        } else if (ss) {
          NF_CUDA(cudaEventRecord(ss->evC, st));
          NF_TRY(col_reduce(&jobs[0], 1, B, s1));
          NF_CUDA(cudaStreamWaitEvent(s1, ss->evC, 0));   // dxp is final after the tail kernel
          NF_TRY(xtg_small(xin, wb.dxp, w.dWs + (int64_t)i * DD, B, m.D, s1));
          if (!fold_h) NF_TRY(col_reduce(&jobs[1], 1, B, s2));
        } else {
          NF_TRY(col_reduce(jobs, fold_h ? 1 : 2, B, st));
          NF_TRY(xtg_small(xin, wb.dxp, w.dWs + (int64_t)i * DD, B, m.D, st));
        }
      }
      if (ss) {
        NF_CUDA(cudaEventRecord(ss->done1[par], s1));
        NF_CUDA(cudaEventRecord(ss->done2[par], s2));
        used1[par] = used2[par] = true;
      }
    } else if (fused) {
      // i (x_cond part), j, k in one kernel
      BwdTailP t{};
      t.du1 = gB; t.du1_net = B * Hm;
      t.W1T = kb + kl.net0 + kl.W1T; t.w_net = kl.net_stride;
      t.z = zi; t.xin = xin; t.W = kb + kl.W;
      t.dxp = wb.dxp; t.dxin = dst; t.ld_dxin = m.D;
      t.G1 = dn ? dn + pl.slot_off[NF_P_W1] : nullptr; t.g_net = pl.net_stride; t.ldg = m.K1;
      t.dW = dn ? w.dWs + (int64_t)i * DD : nullptr;
      t.B = B; t.D = m.D; t.H1 = m.H1;
      NF_TRY(row_bwd_tail(t, st));
    } else {
      if (dn) {  // j. dW = x_in^T dxp   (tensor cores for wide flows; dWs was zero-filled)
        TcGemmTnP g;
        g.A = xin; g.lda = m.D; g.B = wb.dxp; g.ldb = m.D;
        g.C = w.dWs + (int64_t)i * DD; g.ldc = m.D; g.M = m.D; g.N = m.D; g.K = B;
        NF_TRY(wgrad_tn(m.D >= 32 ? m.prec : NF_PREC_FP32_SIMT, g, st));
      }
      if (dst)   // k. d x_in = dxp W^T
        NF_TRY(plu_matmul(m, kl, wb.dxp, m.D, kb + kl.W, dst, m.D, B, st));
    }
    if (dst) { gcur = dst; pp ^= 1; }
  }
  if (ss) {   // join the side streams
    for (int par = 0; par < 2; ++par) {
      if (used1[par]) NF_CUDA(cudaStreamWaitEvent(st, ss->done1[par], 0));
      if (used2[par]) NF_CUDA(cudaStreamWaitEvent(st, ss->done2[par], 0));
    }
  }
  cudaStream_t tail_st = st;
  bool skinny_cols = false;
  if (defer_cols) {   // on s1, beside the batched weight-gradient GEMMs; the un-folding kernel (s1 as well) comes after them
    NF_CUDA(cudaEventRecord(ss->evC, st));
    NF_CUDA(cudaStreamWaitEvent(s1, ss->evC, 0));
    ColJob jobs[kColMaxJobs];
    int nj = 0, first = -1, last = -1;
    bool all_later = true;
    for (int i = sub_begin; i < sub_end; ++i) all_later = all_later && cols_later[i];
    skinny_cols = all_later && option(NF_OPT_DEFER_SKINNY) != 0 && !(whatif & 2);
    for (int i = sub_begin; i < sub_end; ++i) {
      if (!cols_later[i]) continue;
      if (first < 0) first = i;
      last = i;
      float* dn = dparams + (int64_t)i * pl.block_stride;
      const float* sb = stash + sl.blocks + (int64_t)i * sl.block_stride;
      ColJob& j = jobs[nj++];
      j = ColJob();
      j.A = sb + sl.xh2; j.lda = m.C; j.sA = sl.net_stride; j.N = m.C;
      j.coef = w.doutAll + (int64_t)i * 2 * B * m.dt; j.ldcf = m.dt; j.sCf = B * m.dt; j.m = m.dt;
      j.out = dn + pl.slot_off[NF_P_W3]; j.so_i = m.C; j.so_c = 1; j.sOut = pl.net_stride;
      j.out3 = dn + pl.slot_off[NF_P_B3]; j.sOut3 = pl.net_stride;
      j.batch = 2;
      if (nj == kColMaxJobs) { NF_TRY(col_reduce(jobs, nj, B, s1)); nj = 0; }
      if (skinny_cols) {   // G1[:, :dc] = du1^T x_cond here instead of on the dW1 GEMM's splitter threads
        ColJob& k = jobs[nj++];
        k = ColJob();
        k.A = w.gBall + (int64_t)i * 2 * B * Hm; k.lda = m.H1; k.sA = B * Hm; k.N = m.H1;
        k.coef = stash + sl.xs + (int64_t)(i + 1) * BD; k.ldcf = m.D; k.sCf = 0; k.m = m.dc;
        k.out = dn + pl.slot_off[NF_P_W1]; k.so_i = 1; k.so_c = m.K1; k.sOut = pl.net_stride;
        k.batch = 2;
        if (nj == kColMaxJobs) { NF_TRY(col_reduce(jobs, nj, B, s1)); nj = 0; }
      }
    }
    if (nj) NF_TRY(col_reduce(jobs, nj, B, s1));
    if (first >= 0) {   // (every block of the range takes the same path: same shapes, same options)
      NF_CHECK_ARG(last - first + 1 == nr, NF_E_UNSUPPORTED, "deferred column reductions: mixed blocks");
      NF_TRY(xtg_small(stash + sl.xs + (int64_t)first * BD, w.dxpAll + (int64_t)first * BD, w.dWs + (int64_t)first * DD, B, m.D,
                       s1, nr, BD, BD, DD));
    }
    NF_CUDA(cudaEventRecord(ss->done1[0], s1));
    used1[0] = true;
  }
  if (defer_dy && !(whatif & 4)) {
    NF_CUDA(cudaEventRecord(ss->evA, st));
    NF_CUDA(cudaStreamWaitEvent(s2, ss->evA, 0));
    TcGemmP g;
    g.A0 = w.gBall + (int64_t)sub_begin * 2 * B * Hm; g.lda0 = m.H1; g.K0 = m.H1; g.sA0 = 2 * B * Hm;
    g.A1 = g.A0 + B * Hm; g.lda1 = m.H1; g.K1 = m.H1; g.sA1 = 2 * B * Hm;
    g.B0 = packed + (int64_t)sub_begin * kl.block_stride + kl.net0 + kl.W1T + (int64_t)m.dc * m.H1; g.ldb0 = m.H1;
    g.B1 = g.B0 + kl.net_stride; g.ldb1 = m.H1; g.sB0 = g.sB1 = kl.block_stride;
    if (kl.hi_off) { g.B0hi = g.B0 + kl.hi_off; g.B0lo = g.B0 + kl.lo_off; g.B1hi = g.B1 + kl.hi_off; g.B1lo = g.B1 + kl.lo_off; }
    g.C = dy; g.ldc = m.R; g.sC = 0; g.M = (int)B; g.N = m.R; g.batch = nr; g.accumulate = 1;
    NF_TRY(linear_nt(m.prec, g, s2));
    NF_CUDA(cudaEventRecord(ss->done2[0], s2));
  }
  if (defer && !(whatif & 2)) {
    float* dpr = dparams + (int64_t)sub_begin * pl.block_stride;
    TcGemmTnP g;   // e. G2 = du2^T x_hat1 and gb2 = colsum(du2) of every block and net of the range
    g.A = w.gAall + (int64_t)sub_begin * 2 * B * Hm; g.lda = m.C; g.sA = B * Hm;
    g.B = stash + sl.blocks + (int64_t)sub_begin * sl.block_stride + sl.xh1; g.ldb = m.H1; g.sB = sl.net_stride;
    g.C = dpr + pl.slot_off[NF_P_W2]; g.ldc = m.H1; g.sC = pl.net_stride; g.sC2 = pl.block_stride;
    g.M = m.C; g.N = m.H1; g.K = B; g.batch = 2 * nr; g.inner = 2;
    g.colsum = dpr + pl.slot_off[NF_P_B2]; g.sColsum = pl.net_stride; g.sColsum2 = pl.block_stride;
    NF_TRY(wgrad_tn(m.prec, g, st));
    if (ss && dparams) {   // the parameter-sized tail (PLU gradients, un-folding) needs nothing of dW1: it runs beside it
      NF_CUDA(cudaEventRecord(ss->evC, st));
      NF_CUDA(cudaStreamWaitEvent(s1, ss->evC, 0));
      tail_st = s1;
    }
    TcGemmTnP h;   // h. dW1 = du1^T [x_cond | y] and gb1 = colsum(du1)
    h.A = w.gBall + (int64_t)sub_begin * 2 * B * Hm; h.lda = m.H1; h.sA = B * Hm;
    h.B = y; h.ldb = m.R; h.sB = 0;
    h.C = dpr + pl.slot_off[NF_P_W1] + m.dc; h.ldc = m.K1; h.sC = pl.net_stride; h.sC2 = pl.block_stride;
    h.M = m.H1; h.N = m.R; h.K = B; h.batch = 2 * nr; h.inner = 2;
    if (!skinny_cols) {
      h.xc = stash + sl.xs + (int64_t)(sub_begin + 1) * BD; h.ldxc = m.D; h.nxc = m.dc; h.sXc2 = BD;
      h.skinny = dpr + pl.slot_off[NF_P_W1]; h.ldsk = m.K1; h.sSk = pl.net_stride; h.sSk2 = pl.block_stride;
    }
    if (ln1_fused_any) { h.colsum = dpr + pl.slot_off[NF_P_B1]; h.sColsum = pl.net_stride; h.sColsum2 = pl.block_stride; }
    NF_TRY(wgrad_tn(m.prec, h, st));
  }
  if (dparams) {
    const float* prm = params + (int64_t)sub_begin * pl.block_stride;
    const float* pck = packed + (int64_t)sub_begin * kl.block_stride;
    float* dWr = w.dWs + (int64_t)sub_begin * DD;
    Dims mr = m;
    mr.n = nr;   // the parameter-sized kernels below run over this range's blocks only
    float* dpr = dparams + (int64_t)sub_begin * pl.block_stride;
    if (m.D <= 32) {
      NF_TRY(plu_grad_small(prm, pl, mr, pck, kl, dWr, w.sumdld, dpr, tail_st));
    } else {
    GemmP g;  // T1 = P^T dW
    g.transA = 1; g.A0 = prm + pl.slot_off[NF_P_P]; g.lda0 = m.D; g.sA0 = pl.block_stride; g.K0 = m.D;
    g.B0 = dWr; g.ldb0 = m.D; g.sB0 = DD;
    g.C = w.T1; g.ldc = m.D; g.sC = DD; g.M = m.D; g.N = m.D; g.batch = nr;
    NF_TRY(simt_gemm(g, tail_st));
    GemmP h;  // dLh = T1 Uh^T
    h.A0 = w.T1; h.lda0 = m.D; h.sA0 = DD; h.K0 = m.D;
    h.B0 = pck + kl.Uh; h.ldb0 = m.D; h.sB0 = kl.block_stride; h.transB = 1;
    h.C = w.dLh; h.ldc = m.D; h.sC = DD; h.M = m.D; h.N = m.D; h.batch = nr;
    NF_TRY(simt_gemm(h, tail_st));
    GemmP u;  // dUh = (P Lh)^T dW
    u.transA = 1; u.A0 = pck + kl.PL; u.lda0 = m.D; u.sA0 = kl.block_stride; u.K0 = m.D;
    u.B0 = dWr; u.ldb0 = m.D; u.sB0 = DD;
    u.C = w.dUh; u.ldc = m.D; u.sC = DD; u.M = m.D; u.N = m.D; u.batch = nr;
    NF_TRY(simt_gemm(u, tail_st));
    NF_TRY(plu_grad_finish(prm, pl, mr, w.dLh, w.dUh, w.sumdld, dpr, tail_st));
    }
    NF_TRY(unfold_grads(prm, pl, mr, dpr, tail_st));
    if (tail_st != st) {
      NF_CUDA(cudaEventRecord(ss->done1[0], tail_st));
      NF_CUDA(cudaStreamWaitEvent(st, ss->done1[0], 0));
    }
    if (sar) {   // this sub-range's gradients are final: average them over the ranks while the next sub-range computes
      NF_CUDA(cudaEventRecord(sar->evD, st));
      NF_CUDA(cudaStreamWaitEvent(sar->s3, sar->evD, 0));
      NF_TRY(dp_allreduce(dpr, (int64_t)nr * pl.block_stride, true, sar->s3));
      any_ar = true;
    }
  }
  if (defer_dy && !(whatif & 4)) NF_CUDA(cudaStreamWaitEvent(st, ss->done2[0], 0));
  sub_end = sub_begin;
  }   // sub-ranges
  if (any_ar) {
    NF_CUDA(cudaEventRecord(sar->done3, sar->s3));
    NF_CUDA(cudaStreamWaitEvent(st, sar->done3, 0));
  }
  return 0;
  };
  NF_TRY(run_graphed(key, user_st, body));
  if (last_range) {
    CopySegs c;
    c.add(dx_user, dx, BD);
    c.add(dy_user, dy, B * m.R);
    NF_TRY(copy_segments(c, user_st));
  }
  return 0;
}

int nf_flow_backward_range(const NfFlowDesc* desc, const float* params, const void* packed_v, const void* stash_v,
                           const float* y, int64_t B, const float* dz, const float* dlogdet, float* dx, float* dy,
                           float* dparams, int32_t block_begin, int32_t block_end, void* workspace, int64_t ws_bytes,
                           void* stream) {
  return backward_impl(desc, params, packed_v, stash_v, y, B, dz, dlogdet, dx, dy, dparams, block_begin, block_end, 0,
                       workspace, ws_bytes, stream);
}

int nf_flow_backward(const NfFlowDesc* desc, const float* params, const void* packed_v, const void* stash_v,
                     const float* y, int64_t B, const float* dz, const float* dlogdet, float* dx, float* dy,
                     float* dparams, void* workspace, int64_t ws_bytes, void* stream) {
  NF_TRY(check_desc(desc));
  return backward_impl(desc, params, packed_v, stash_v, y, B, dz, dlogdet, dx, dy, dparams, 0, desc->n_blocks, 0,
                       workspace, ws_bytes, stream);
}

int nf_flow_backward_dp(const NfFlowDesc* desc, const float* params, const void* packed_v, const void* stash_v,
                        const float* y, int64_t B, const float* dz, const float* dlogdet, float* dx, float* dy,
                        float* dparams, int32_t bucket_blocks, void* workspace, int64_t ws_bytes, void* stream) {
  NF_TRY(check_desc(desc));
  NF_CHECK_ARG(bucket_blocks >= 1, NF_E_BADARG, "bucket_blocks must be >= 1");
  return backward_impl(desc, params, packed_v, stash_v, y, B, dz, dlogdet, dx, dy, dparams, 0, desc->n_blocks,
                       bucket_blocks, workspace, ws_bytes, stream);
}

// Backward of the REVERSE direction (x = reverse(z, y), optionally with its log-det; flax nf.py:131-148, used
// differentiably by offline-rl/agents/vinf.py:50,68).  The intermediate values of reverse(z) are those of forward(x)
// (block i's forward input is its reverse output), so the stash is the one nf_flow_forward(x = reverse output) leaves.
// Per block, in forward block order i = 0 .. n-1 (the reverse pass ran n-1 .. 0), with g the cotangent of the block's
// reverse output x_i:
//   g_xp = g W_i^-T ;  dW_i = -x_i^T g_xp          (x_i = xp W^-1  =>  dW^-1 = xp^T g,  dW = -W^-T dW^-1 W^-T)
//   dz_t = g_xp_t e^s ;  ds = g_xp_t z_t e^s + dlogdet ;  dt = g_xp_t ;  dz_c = g_xp_c + (conditioner input gradient)
// and the conditioner backward (last Linear, LayerNorm / LeakyReLU, the two wide Linears, dy) is the forward
// direction's.  This is the general path (plain GEMM + row kernels), not a tuned one.
int nf_flow_inverse_backward(const NfFlowDesc* desc, const float* params, const void* packed_v, const void* stash_v,
                             const float* y, int64_t B, const float* dx, const float* dlogdet, float* dz, float* dy,
                             float* dparams, void* workspace, int64_t ws_bytes, void* stream) {
  NF_TRY(check_desc(desc));
  NF_CHECK_ARG(B >= 0 && B < (int64_t(1) << 30), NF_E_BADARG, "B=%lld out of range", (long long)B);
  NF_TRY(check_ptr(params, "params", true));
  NF_TRY(check_ptr(packed_v, "packed", true));
  NF_TRY(check_ptr(dparams, "dparams", false));
  const Dims m = dims_of(desc);
  NfParamLayout pl;
  param_layout(m, &pl);
  cudaStream_t user_st = reinterpret_cast<cudaStream_t>(stream);
  if (B == 0) {
    if (dparams) NF_TRY(fill_zero(dparams, pl.total, user_st));
    return 0;
  }
  NF_TRY(check_ptr(stash_v, "stash", true));
  NF_TRY(check_ptr(y, "y", false));
  NF_TRY(check_ptr(dx, "dx", false)); NF_TRY(check_ptr(dlogdet, "dlogdet", false));
  NF_TRY(check_ptr(dz, "dz", false)); NF_TRY(check_ptr(dy, "dy", false));
  NF_TRY(check_ptr(workspace, "workspace", true));
  const PackedLayout kl = packed_layout(m);
  const StashLayout sl = stash_layout(m, B);
  const BwdWs w = carve_bwd(m, B, workspace);
  NF_CHECK_ARG(ws_bytes >= w.total * (int64_t)sizeof(float), NF_E_WORKSPACE, "workspace too small: %lld < %lld",
               (long long)ws_bytes, (long long)(w.total * sizeof(float)));
  const float* packed = reinterpret_cast<const float*>(packed_v);
  const float* stash = reinterpret_cast<const float*>(stash_v);
  const int64_t BD = B * m.D, Hm = std::max(m.H1, m.C), DD = (int64_t)m.D * m.D;
  const bool has_y = y != nullptr && m.R > 0;

  GraphKey key;
  key.op = 6; key.desc = *desc; key.B = B; key.ws_bytes = ws_bytes;
  key.ptr[0] = params; key.ptr[1] = packed_v; key.ptr[2] = stash_v; key.ptr[8] = dparams; key.ptr[9] = workspace;
  key.flags = (dx ? 1 : 0) | (dlogdet ? 2 : 0) | (dz ? 4 : 0) | (dy ? 8 : 0) | (has_y ? 16 : 0);
  float* const dz_user = dz;
  float* const dy_user = dy;
  {
    CopySegs c;
    c.add(w.io_dz, dx, BD);
    c.add(w.io_dld, dlogdet, B);
    NF_TRY(copy_segments(c, user_st));
  }
  if (dx) dx = w.io_dz;
  if (dlogdet) dlogdet = w.io_dld;
  if (dz) dz = w.io_dx;
  if (dy) dy = w.io_dy;
  y = has_y ? stash + sl.y : nullptr;
  auto body = [&](cudaStream_t st) -> int {
  if (dparams) {
    NF_TRY(fill_zero(dparams, pl.total, st));
    NF_TRY(fill_zero(w.dWs, m.n * DD, st));
    if (dlogdet) NF_TRY(sum_vector(dlogdet, B, w.sumdld, st));
    else NF_TRY(fill_zero(w.sumdld, 4, st));
  }
  if (dy) NF_TRY(fill_zero(dy, B * m.R, st));
  if (!dx) NF_TRY(fill_zero(w.io_dz, BD, st));
  const float* gcur = dx ? dx : w.io_dz;
  struct { float *dxp, *dout, *gA, *gB; } w2[2] = {{w.dxp, w.dout, w.gA, w.gB}, {w.dxp2, w.dout2, w.gA2, w.gB2}};
  for (int i = 0; i < m.n; ++i) {
    const auto& wb = w2[i & 1];
    const float* kb = packed + (int64_t)i * kl.block_stride;
    float* dn = dparams ? dparams + (int64_t)i * pl.block_stride : nullptr;
    const float* xin = stash + sl.xs + (int64_t)i * BD;          // reverse output of block i
    const float* zi = stash + sl.xs + (int64_t)(i + 1) * BD;     // reverse input of block i
    const float* si = stash + sl.ss + (int64_t)i * B * m.dt;
    const float* sb = stash + sl.blocks + (int64_t)i * sl.block_stride;
    const uint32_t* m1 = reinterpret_cast<const uint32_t*>(sb + sl.m1);
    const uint32_t* m2 = reinterpret_cast<const uint32_t*>(sb + sl.m2);
    NF_TRY(plu_matmul(m, kl, gcur, m.D, kb + kl.Winv, w.d0, m.D, B, st));   // g_xp = g W^-T  -> d0 (scratch)
    if (dn) {  // dW = -x_in^T g_xp
      GemmP g;
      g.transA = 1; g.A0 = xin; g.lda0 = m.D; g.K0 = (int)B;
      g.B0 = w.d0; g.ldb0 = m.D;
      g.C = w.dWs + (int64_t)i * DD; g.ldc = m.D; g.M = m.D; g.N = m.D;
      g.splitk = pick_splitk(g.M, g.N, B, 1);
      g.alpha = -1.0f;
      NF_TRY(simt_gemm(g, st));
    }
    // affine coupling, reverse direction: dxp := [g_xp_c, g_xp_t e^s] (= dz of this block before the conditioner term)
    NF_TRY(affine_bwd(w.d0, zi, si, dlogdet, B, m.D, wb.dxp, wb.dout + B * m.dt, wb.dout,
                      dn ? dn + pl.net_stride + pl.slot_off[NF_P_B3] : nullptr,
                      dn ? dn + pl.slot_off[NF_P_B3] : nullptr, st, 1));
    if (dn) {  // G3 = dout^T x_hat2
      TcGemmTnP g;
      g.A = wb.dout; g.lda = m.dt; g.sA = B * m.dt;
      g.B = sb + sl.xh2; g.ldb = m.C; g.sB = sl.net_stride;
      g.C = dn + pl.slot_off[NF_P_W3]; g.ldc = m.C; g.sC = pl.net_stride;
      g.M = m.dt; g.N = m.C; g.K = B; g.batch = 2;
      NF_TRY(wgrad_tn(m.prec, g, st));
    }
    {  // d x_hat2 = dout W3f
      GemmP g;
      g.A0 = wb.dout; g.lda0 = m.dt; g.sA0 = B * m.dt; g.K0 = m.dt;
      g.B0 = kb + kl.net0 + kl.W3f; g.ldb0 = m.C; g.sB0 = kl.net_stride;
      g.C = wb.gA; g.ldc = m.C; g.sC = B * Hm;
      g.M = (int)B; g.N = m.C; g.batch = 2;
      NF_TRY(simt_gemm(g, st));
    }
    {  // through LayerNorm + LeakyReLU (layer 2); column sums -> slot b2
      LnBwdP p{};
      p.g = wb.gA; p.g_net = B * Hm;
      p.xh = sb + sl.xh2; p.xh_net = sl.net_stride;
      p.rstd = sb + sl.r2; p.rstd_net = sl.net_stride;
      p.mask = m2; p.mask_net = sl.net_stride; p.mw = sl.mw2;
      p.colsum = dn ? dn + pl.slot_off[NF_P_B2] : nullptr; p.colsum_net = pl.net_stride;
      p.N = m.C; p.nets = 2; p.B = B;
      NF_TRY(ln_lrelu_bwd(p, st));
    }
    if (dn) {  // G2 = du2^T x_hat1
      TcGemmTnP g;
      g.A = wb.gA; g.lda = m.C; g.sA = B * Hm;
      g.B = sb + sl.xh1; g.ldb = m.H1; g.sB = sl.net_stride;
      g.C = dn + pl.slot_off[NF_P_W2]; g.ldc = m.H1; g.sC = pl.net_stride;
      g.M = m.C; g.N = m.H1; g.K = B; g.batch = 2;
      NF_TRY(wgrad_tn(m.prec, g, st));
    }
    {  // d x_hat1 = du2 W2f
      TcGemmP g;
      g.A0 = wb.gA; g.lda0 = m.C; g.sA0 = B * Hm; g.K0 = m.C;
      g.B0 = kb + kl.net0 + kl.W2fT; g.ldb0 = m.C; g.sB0 = kl.net_stride;
      g.C = wb.gB; g.ldc = m.H1; g.sC = B * Hm;
      g.M = (int)B; g.N = m.H1; g.batch = 2;
      if (kl.hi_off) { g.B0hi = g.B0 + kl.hi_off; g.B0lo = g.B0 + kl.lo_off; }
      NF_TRY(linear_nt(m.prec, g, st));
    }
    {  // through LayerNorm + LeakyReLU (layer 1); column sums -> slot b1
      LnBwdP p{};
      p.g = wb.gB; p.g_net = B * Hm;
      p.xh = sb + sl.xh1; p.xh_net = sl.net_stride;
      p.rstd = sb + sl.r1; p.rstd_net = sl.net_stride;
      p.mask = m1; p.mask_net = sl.net_stride; p.mw = sl.mw1;
      p.colsum = dn ? dn + pl.slot_off[NF_P_B1] : nullptr; p.colsum_net = pl.net_stride;
      p.N = m.H1; p.nets = 2; p.B = B;
      NF_TRY(ln_lrelu_bwd(p, st));
    }
    if (dn) {  // dW1 = du1^T [z_c | y]
      TcGemmTnP g;
      g.A = wb.gB; g.lda = m.H1; g.sA = B * Hm;
      g.B = zi; g.ldb = m.D; g.sB = 0;
      g.C = dn + pl.slot_off[NF_P_W1]; g.ldc = m.K1; g.sC = pl.net_stride;
      g.M = m.H1; g.N = m.dc; g.K = B; g.batch = 2;
      NF_TRY(wgrad_tn(m.prec, g, st));
      if (has_y) {
        g.B = y; g.ldb = m.R;
        g.C = dn + pl.slot_off[NF_P_W1] + m.dc; g.N = m.R;
        NF_TRY(wgrad_tn(m.prec, g, st));
      }
    }
    {  // d h0 = du1_t W1_t + du1_s W1_s : first dc columns -> dz_c, the rest -> dy
      TcGemmP g;
      g.A0 = wb.gB; g.lda0 = m.H1; g.K0 = m.H1;
      g.A1 = wb.gB + B * Hm; g.lda1 = m.H1; g.K1 = m.H1;
      g.B0 = kb + kl.net0 + kl.W1T; g.ldb0 = m.H1;
      g.B1 = g.B0 + kl.net_stride; g.ldb1 = m.H1;
      g.C = wb.dxp; g.ldc = m.D; g.M = (int)B; g.N = m.dc; g.accumulate = 1;
      NF_TRY(linear_nt(m.prec, g, st));
      if (dy && has_y) {
        g.B0 += (int64_t)m.dc * m.H1; g.B1 += (int64_t)m.dc * m.H1;
        if (kl.hi_off) { g.B0hi = g.B0 + kl.hi_off; g.B0lo = g.B0 + kl.lo_off; g.B1hi = g.B1 + kl.hi_off; g.B1lo = g.B1 + kl.lo_off; }
        g.C = dy; g.ldc = m.R; g.N = m.R;
        NF_TRY(linear_nt(m.prec, g, st));
      }
    }
    gcur = wb.dxp;
  }
  if (dz) NF_CUDA(cudaMemcpyAsync(dz, gcur, BD * sizeof(float), cudaMemcpyDeviceToDevice, st));
  if (dparams) {
    if (m.D <= 32) {
      NF_TRY(plu_grad_small(params, pl, m, packed, kl, w.dWs, w.sumdld, dparams, st, -1.0f));
    } else {
    GemmP g;  // T1 = P^T dW
    g.transA = 1; g.A0 = params + pl.slot_off[NF_P_P]; g.lda0 = m.D; g.sA0 = pl.block_stride; g.K0 = m.D;
    g.B0 = w.dWs; g.ldb0 = m.D; g.sB0 = DD;
    g.C = w.T1; g.ldc = m.D; g.sC = DD; g.M = m.D; g.N = m.D; g.batch = m.n;
    NF_TRY(simt_gemm(g, st));
    GemmP h;  // dLh = T1 Uh^T
    h.A0 = w.T1; h.lda0 = m.D; h.sA0 = DD; h.K0 = m.D;
    h.B0 = packed + kl.Uh; h.ldb0 = m.D; h.sB0 = kl.block_stride; h.transB = 1;
    h.C = w.dLh; h.ldc = m.D; h.sC = DD; h.M = m.D; h.N = m.D; h.batch = m.n;
    NF_TRY(simt_gemm(h, st));
    GemmP u;  // dUh = (P Lh)^T dW
    u.transA = 1; u.A0 = packed + kl.PL; u.lda0 = m.D; u.sA0 = kl.block_stride; u.K0 = m.D;
    u.B0 = w.dWs; u.ldb0 = m.D; u.sB0 = DD;
    u.C = w.dUh; u.ldc = m.D; u.sC = DD; u.M = m.D; u.N = m.D; u.batch = m.n;
    NF_TRY(simt_gemm(u, st));
    NF_TRY(plu_grad_finish(params, pl, m, w.dLh, w.dUh, w.sumdld, dparams, st, -1.0f));
    }
    NF_TRY(unfold_grads(params, pl, m, dparams, st));
  }
  return 0;
  };
  NF_TRY(run_graphed(key, user_st, body));
  {
    CopySegs c;
    c.add(dz_user, dz, BD);
    c.add(dy_user, dy, B * m.R);
    NF_TRY(copy_segments(c, user_st));
  }
  return 0;
}

int nf_linear(const float* A, int64_t lda, const float* W, int64_t ldw, const float* bias, float* C, int64_t ldc,
              int64_t M, int32_t N, int32_t K, int32_t engine, void* stream) {
  NF_CHECK_ARG(M >= 0 && N >= 1 && K >= 1 && M < (int64_t(1) << 30), NF_E_BADARG, "bad shape M=%lld N=%d K=%d",
               (long long)M, N, K);
  NF_CHECK_ARG(lda >= K && ldw >= K && ldc >= N, NF_E_BADARG, "row pitch smaller than the row");
  NF_TRY(check_ptr(A, "A", true)); NF_TRY(check_ptr(W, "W", true)); NF_TRY(check_ptr(C, "C", true));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (engine == NF_PREC_FP32_SIMT) {
    GemmP g;
    g.A0 = A; g.lda0 = (int)lda; g.K0 = K; g.B0 = W; g.ldb0 = (int)ldw; g.transB = 1;
    g.C = C; g.ldc = (int)ldc; g.bias = bias; g.M = (int)M; g.N = N;
    return simt_gemm(g, st);
  }
  NF_CHECK_ARG(engine == NF_PREC_TF32X3 || (engine >= 3 && engine <= 7), NF_E_UNSUPPORTED, "unknown engine %d", engine);
  TcGemmP t;
  t.A0 = A; t.lda0 = lda; t.K0 = K; t.B0 = W; t.ldb0 = ldw; t.C = C; t.ldc = ldc; t.bias = bias;
  t.M = (int)M; t.N = N; t.passes = (engine == 3 || engine == 7) ? 1 : 3;
  t.raw_hi = (engine == 4 || engine == 6 || engine == 3 || engine == 7);
  t.bk = (engine == 5 || engine == 6 || engine == 7) ? 32 : 16;
  return tc_gemm(t, st);
}

int nf_linear_wgrad(const float* dY, int64_t lddy, const float* X, int64_t ldx, float* dW, int64_t lddw, int64_t B,
                    int32_t n_out, int32_t n_in, int32_t engine, void* stream) {
  NF_CHECK_ARG(B >= 0 && n_out >= 1 && n_in >= 1, NF_E_BADARG, "bad shape");
  NF_TRY(check_ptr(dY, "dY", true)); NF_TRY(check_ptr(X, "X", true)); NF_TRY(check_ptr(dW, "dW", true));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (engine == NF_PREC_FP32_SIMT) {
    GemmP g;
    g.transA = 1; g.A0 = dY; g.lda0 = (int)lddy; g.K0 = (int)B;
    g.B0 = X; g.ldb0 = (int)ldx;
    g.C = dW; g.ldc = (int)lddw; g.M = n_out; g.N = n_in;
    g.splitk = pick_splitk(n_out, n_in, B, 1);
    if (g.splitk == 1) g.accumulate = 1;
    return simt_gemm(g, st);
  }
  TcGemmTnP t;
  t.A = dY; t.lda = lddy; t.B = X; t.ldb = ldx; t.C = dW; t.ldc = lddw;
  t.M = n_out; t.N = n_in; t.K = B; t.passes = engine == 3 ? 1 : 3;
  t.bk = engine == 5 ? 32 : 16;
  return tc_gemm_tn(t, st);
}

int nf_affine_logdet_fwd(const float* xp, const float* s, const float* t, const float* bias_s, const float* bias_t,
                         float logdet_const, int64_t B, int32_t D, float* z, float* logdet, float* s_out,
                         void* stream) {
  NF_CHECK_ARG(D >= 2 && B >= 0, NF_E_BADARG, "bad shape B=%lld D=%d", (long long)B, D);
  NF_TRY(check_ptr(xp, "xp", true)); NF_TRY(check_ptr(s, "s", true)); NF_TRY(check_ptr(t, "t", true));
  NF_TRY(check_ptr(z, "z", true)); NF_TRY(check_ptr(logdet, "logdet", false)); NF_TRY(check_ptr(s_out, "s_out", false));
  return affine_fwd(xp, D, s, t, bias_s, bias_t, nullptr, logdet_const, B, D, z, D, logdet, s_out,
                    reinterpret_cast<cudaStream_t>(stream));
}

int nf_affine_logdet_inv(const float* z, const float* s, const float* t, const float* bias_s, const float* bias_t,
                         float logdet_const, int64_t B, int32_t D, float* xp, float* logdet, void* stream) {
  NF_CHECK_ARG(D >= 2 && B >= 0, NF_E_BADARG, "bad shape B=%lld D=%d", (long long)B, D);
  NF_TRY(check_ptr(z, "z", true)); NF_TRY(check_ptr(s, "s", true)); NF_TRY(check_ptr(t, "t", true));
  NF_TRY(check_ptr(xp, "xp", true)); NF_TRY(check_ptr(logdet, "logdet", false));
  return affine_inv(z, D, s, t, bias_s, bias_t, nullptr, logdet_const, B, D, xp, D, logdet,
                    reinterpret_cast<cudaStream_t>(stream));
}

int nf_affine_logdet_bwd(const float* dz, const float* z, const float* s_full, const float* dlogdet, int64_t B,
                         int32_t D, float* dxp, float* ds, float* dt, void* stream) {
  NF_CHECK_ARG(D >= 2 && B >= 0, NF_E_BADARG, "bad shape B=%lld D=%d", (long long)B, D);
  NF_TRY(check_ptr(z, "z", true)); NF_TRY(check_ptr(s_full, "s_full", true));
  NF_TRY(check_ptr(dxp, "dxp", true)); NF_TRY(check_ptr(ds, "ds", true)); NF_TRY(check_ptr(dt, "dt", true));
  return affine_bwd(dz, z, s_full, dlogdet, B, D, dxp, ds, dt, nullptr, nullptr,
                    reinterpret_cast<cudaStream_t>(stream));
}

int nf_nll_forward(const float* z, const float* logdet, int64_t B, int32_t D, float* loss, void* stream) {
  NF_CHECK_ARG(D >= 1 && B >= 0, NF_E_BADARG, "bad shape B=%lld D=%d", (long long)B, D);
  NF_TRY(check_ptr(z, "z", B > 0)); NF_TRY(check_ptr(logdet, "logdet", B > 0)); NF_TRY(check_ptr(loss, "loss", true));
  return nll_forward(z, logdet, B, D, loss, reinterpret_cast<cudaStream_t>(stream));
}

int nf_nll_cotangents(const float* z, const float* g, int64_t B, int32_t D, float* dz, float* dld, void* stream) {
  NF_CHECK_ARG(D >= 1 && B >= 0, NF_E_BADARG, "bad shape B=%lld D=%d", (long long)B, D);
  NF_TRY(check_ptr(z, "z", B > 0)); NF_TRY(check_ptr(dz, "dz", B > 0)); NF_TRY(check_ptr(dld, "dld", B > 0));
  return nll_cotangents(z, g, B, D, dz, dld, reinterpret_cast<cudaStream_t>(stream));
}

int nf_stdnormal_logprob(const float* z, int64_t B, int32_t D, float* lp, void* stream) {
  NF_CHECK_ARG(D >= 1 && B >= 0, NF_E_BADARG, "bad shape B=%lld D=%d", (long long)B, D);
  NF_TRY(check_ptr(z, "z", B > 0)); NF_TRY(check_ptr(lp, "lp", B > 0));
  return stdnormal_logprob(z, B, D, lp, reinterpret_cast<cudaStream_t>(stream));
}

int nf_stdnormal_logprob_bwd(const float* z, const float* g, int64_t B, int32_t D, float* dz, void* stream) {
  NF_CHECK_ARG(D >= 1 && B >= 0, NF_E_BADARG, "bad shape B=%lld D=%d", (long long)B, D);
  NF_TRY(check_ptr(z, "z", B > 0)); NF_TRY(check_ptr(g, "g", B > 0)); NF_TRY(check_ptr(dz, "dz", B > 0));
  return stdnormal_logprob_bwd(z, g, B, D, dz, reinterpret_cast<cudaStream_t>(stream));
}

int nf_logprob_group_argmin(const float* z, const float* logdet, const float* x, int64_t G, int64_t N, int32_t D,
                            float* out_x, int64_t* out_idx, float* out_lp, void* stream) {
  NF_CHECK_ARG(D >= 1 && G >= 0 && N >= 1 && G <= 0x7fffffff, NF_E_BADARG, "bad shape G=%lld N=%lld D=%d", (long long)G,
               (long long)N, D);
  NF_TRY(check_ptr(z, "z", G > 0)); NF_TRY(check_ptr(logdet, "logdet", G > 0)); NF_TRY(check_ptr(x, "x", out_x != nullptr));
  static_assert(sizeof(long long) == sizeof(int64_t), "index type");
  return group_argmin_logp(z, logdet, x, G, N, D, out_x, reinterpret_cast<long long*>(out_idx), out_lp,
                           reinterpret_cast<cudaStream_t>(stream));
}

int nf_noise_augment(const float* x, int64_t n, float std_, uint64_t seed, uint64_t offset, float* out, void* stream) {
  NF_CHECK_ARG(n >= 0, NF_E_BADARG, "n=%lld out of range", (long long)n);
  NF_TRY(check_ptr(x, "x", n > 0)); NF_TRY(check_ptr(out, "out", n > 0));
  return noise_augment(x, n, std_, seed, offset, out, reinterpret_cast<cudaStream_t>(stream));
}

int nf_adamw_step(const NfFlowDesc* desc, float* params, const float* grads, float* exp_avg, float* exp_avg_sq,
                  const unsigned char* slot_mask, float lr, float beta1, float beta2, float eps, float weight_decay,
                  int64_t step, void* stream) {
  NF_TRY(check_desc(desc));
  NF_TRY(check_ptr(params, "params", true)); NF_TRY(check_ptr(grads, "grads", true));
  NF_TRY(check_ptr(exp_avg, "exp_avg", true)); NF_TRY(check_ptr(exp_avg_sq, "exp_avg_sq", true));
  NF_CHECK_ARG(slot_mask != nullptr && step >= 1, NF_E_BADARG, "slot_mask is NULL or step < 1");
  const Dims m = dims_of(desc);
  NfParamLayout pl;
  param_layout(m, &pl);
  AdamWP p{};
  p.params = params; p.grads = grads; p.exp_avg = exp_avg; p.exp_avg_sq = exp_avg_sq; p.slot_mask = slot_mask;
  const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
  p.lr_wd_factor = (float)(1.0 - (double)lr * (double)weight_decay);
  p.one_minus_beta1 = (float)(1.0 - (double)beta1); p.beta2 = beta2; p.one_minus_beta2 = (float)(1.0 - (double)beta2);
  p.step_size = (float)((double)lr / bc1); p.bc2_sqrt = (float)sqrt(bc2); p.eps = eps;
  return adamw_step(p, pl, m.n, reinterpret_cast<cudaStream_t>(stream));
}

int nf_plu_build(const float* s, const float* U, const float* L, const float* P, const float* P_inv, int32_t D,
                 float* W, float* W_inv, float* logdet, void* workspace, int64_t ws_bytes, void* stream) {
  // One-block flow with no conditioner: reuse the packed-table machinery on a temporary arena
  // carved from the workspace.
  NF_CHECK_ARG(D >= 2 && D <= 4096, NF_E_BADARG, "D=%d out of range", D);
  NF_TRY(check_ptr(workspace, "workspace", true));
  NfFlowDesc d{};
  d.D = D; d.C = 1; d.H1 = 1; d.R = 0; d.n_blocks = 1; d.ln_eps = 1e-5f; d.precision = 0;
  const Dims m = dims_of(&d);
  NfParamLayout pl;
  param_layout(m, &pl);
  const PackedLayout kl = packed_layout(m);
  const PrepWs pw = carve_prep(m, nullptr);
  const int64_t need = (pad32(pl.total) + pad32(kl.total) + pw.total) * (int64_t)sizeof(float);
  NF_CHECK_ARG(ws_bytes >= need, NF_E_WORKSPACE, "workspace too small: %lld < %lld", (long long)ws_bytes,
               (long long)need);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  float* arena = reinterpret_cast<float*>(workspace);
  float* packed = arena + pad32(pl.total);
  float* rest = packed + pad32(kl.total);
  const int64_t DD = (int64_t)D * D;
  NF_CUDA(cudaMemsetAsync(arena, 0, pl.total * sizeof(float), st));
  NF_CUDA(cudaMemcpyAsync(arena + pl.slot_off[NF_P_S], s, D * sizeof(float), cudaMemcpyDeviceToDevice, st));
  NF_CUDA(cudaMemcpyAsync(arena + pl.slot_off[NF_P_U], U, DD * sizeof(float), cudaMemcpyDeviceToDevice, st));
  NF_CUDA(cudaMemcpyAsync(arena + pl.slot_off[NF_P_L], L, DD * sizeof(float), cudaMemcpyDeviceToDevice, st));
  NF_CUDA(cudaMemcpyAsync(arena + pl.slot_off[NF_P_P], P, DD * sizeof(float), cudaMemcpyDeviceToDevice, st));
  if (P_inv)
    NF_CUDA(cudaMemcpyAsync(arena + pl.slot_off[NF_P_PINV], P_inv, DD * sizeof(float), cudaMemcpyDeviceToDevice, st));
  NF_TRY(nf_flow_prepare(&d, arena, packed, NF_PREP_FWD | (W_inv ? NF_PREP_INV : 0), rest,
                         pw.total * (int64_t)sizeof(float), stream));
  if (W) NF_CUDA(cudaMemcpyAsync(W, packed + kl.W, DD * sizeof(float), cudaMemcpyDeviceToDevice, st));
  if (W_inv) NF_CUDA(cudaMemcpyAsync(W_inv, packed + kl.Winv, DD * sizeof(float), cudaMemcpyDeviceToDevice, st));
  if (logdet) NF_CUDA(cudaMemcpyAsync(logdet, packed + kl.logdet_c, sizeof(float), cudaMemcpyDeviceToDevice, st));
  return 0;
}

}  // extern "C"
