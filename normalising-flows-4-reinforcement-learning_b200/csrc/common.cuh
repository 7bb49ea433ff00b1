// Shared host/device helpers for the nf_b200 C-ABI library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/nf_b200.h"

namespace nf {

constexpr float kLeakySlope = 0.01f;  // nn.LeakyReLU() default (torch_fcnf.py:75)

// ---- error plumbing (thread-local message, never throws) --------------------------------
char* err_buf();
int set_err(int code, const char* fmt, ...);
int64_t& launch_counter();

#define NF_CHECK_ARG(cond, code, ...)                    \
  do {                                                   \
    if (!(cond)) return nf::set_err((code), __VA_ARGS__); \
  } while (0)

#define NF_CUDA(expr)                                                                      \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess)                                                                 \
      return nf::set_err((int)_e, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),  \
                         __FILE__, __LINE__);                                              \
  } while (0)

// counts the launch and converts a launch failure to a return code
#define NF_LAUNCHED()                    \
  do {                                   \
    nf::launch_counter()++;              \
    NF_CUDA(cudaPeekAtLastError());      \
  } while (0)

#define NF_TRY(expr)            \
  do {                          \
    int _r = (expr);            \
    if (_r != 0) return _r;     \
  } while (0)

// the C ABI's pointer contract: 16-byte aligned, NULL only where the argument is optional
static inline int check_ptr(const void* p, const char* name, bool required) {
  if (!p) {
    NF_CHECK_ARG(!required, NF_E_BADARG, "%s is NULL", name);
    return 0;
  }
  NF_CHECK_ARG((reinterpret_cast<uintptr_t>(p) & 15u) == 0, NF_E_ALIGN, "%s is not 16-byte aligned", name);
  return 0;
}

// ---- run-time options (nf_set_option / NF_B200_<NAME> environment variables) ------------------------------
enum NfOpt {
  NF_OPT_PDL = 0,        // programmatic dependent launch on every kernel                       (default 1)
  NF_OPT_STREAMS,        // backward on three streams                                           (default 1)
  NF_OPT_TMA_STORE,      // GEMM epilogues write through bulk tensor stores                     (default 1)
  NF_OPT_NACC,           // main accumulators in rotation, 0 = kernel default (3)               (default 0)
  NF_OPT_FUSED_LN,       // LayerNorm fused into the GEMM epilogue: -1 auto, 0 never, 1 always  (default -1)
  NF_OPT_FUSED_TAIL,     // flow tail fused into the layer-2 GEMM                               (default 0)
  NF_OPT_BN256,          // 256-wide GEMM tiles: 0 auto (K <= 256), 1 always, 2 never           (default 0)
  NF_OPT_TS,             // A operand of the K-major GEMMs from tensor memory                   (default 1)
  NF_OPT_PRESPLIT,       // weight operands split into hi / lo once per parameter update        (default 1)
  NF_OPT_STACK,          // persistent whole-stack forward / inverse kernel: -1 auto (small grids), 0 never, 1 always (default -1)
  NF_OPT_STACK_NSG,      // splitter groups of the whole-stack kernel: 1 or 2 (default 2)
  NF_OPT_WHATIF,         // EXPERIMENTS ONLY (wrong gradients): bit 0 skip column reductions, 1 skip weight-gradient
                         // GEMMs, 2 skip the dy GEMM -- upper bounds for what removing that work could buy (default 0)
  NF_OPT_RIDE,           // bias / x_cond weight gradients taken along by the weight-gradient GEMMs          (default 1)
  NF_OPT_DEFER_WGRAD,    // conditioner weight gradients of all blocks in two batched launches after the block loop: -1 auto (small flows), 0, 1 (default -1)
  NF_OPT_MCAST,          // weight tiles of the large K-major GEMMs multicast over CTA pairs: -1 auto, 0 never, 1 whenever possible (default -1)
  NF_OPT_BN64,           // 64-wide tiles for K-major GEMMs whose 128-wide grid fills at most half the SMs (default 1)
  NF_OPT_TF_AR,          // transformer flow: persistent autoregressive-inverse kernel for B <= 128: -1 auto, 0 never (default -1)
  NF_OPT_PREP_FUSED,     // D <= 32: forward-direction tables of nf_flow_prepare in one kernel (default 1)
  NF_OPT_DEFER_DY,       // with deferred weight gradients: the dy GEMMs of all blocks as one batched launch after the loop (default 1)
  NF_OPT_DEFER_COLS,     // with deferred weight gradients: the remaining column reductions of all blocks as one launch each after the loop (default 1)
  NF_OPT_DEFER_SKINNY,   // with deferred column reductions: the x_cond columns of dW1 as column-reduction jobs instead of riding on the dW1 GEMM (default 0: measured 7 us slower on C2a)
  NF_OPT_COUNT
};
int option(int id);

// ---- programmatic dependent launch (PDL) -------------------------------------------------------
// Every kernel of the library is launched with cudaLaunchAttributeProgrammaticStreamSerialization and
// starts with NF_PDL_ENTRY(): it blocks until the kernel before it in the stream has completed and
// flushed (griddepcontrol.wait), then lets the kernel after it be scheduled (launch_dependents), so
// that launch latency and prologue of kernel N+1 overlap the execution of kernel N -- the flow is a
// chain of 60-500 short dependent kernels.  Because every kernel waits before its first global
// access, completion is transitive and no RAW/WAR hazard is exposed.  NF_B200_PDL=0 turns it off.
#if defined(__CUDACC__)
#define NF_PDL_ENTRY()                                          \
  do {                                                          \
    asm volatile("griddepcontrol.wait;" ::: "memory");          \
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); \
  } while (0)

inline bool pdl_enabled() { return option(NF_OPT_PDL) != 0; }
template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#define NF_LAUNCH(kernel, grid, block, smem, st, ...) \
  (void)nf::launch_pdl(kernel, dim3(grid), dim3(block), (size_t)(smem), st, __VA_ARGS__)
#endif

// cudaFuncSetAttribute is per device: a call site remembers which devices it has configured (the Python layer
// runs modules on non-current devices too, so one process-wide flag would leave the second GPU without the
// > 48 KB dynamic shared memory opt-in).
struct PerDeviceOnce {
  bool done[64] = {};
  bool first() {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64) return true;
    if (done[d]) return false;
    done[d] = true;
    return true;
  }
};

static inline int64_t pad4(int64_t v) { return (v + 3) & ~int64_t(3); }
static inline int64_t pad32(int64_t v) { return (v + 31) & ~int64_t(31); }
static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

struct Dims {
  int D, C, H1, R, n, dc, dt, K1;
  float eps;
  int prec;
  int fastvar;   // NF_FLAG_LN_FAST_VARIANCE
};

static inline Dims dims_of(const NfFlowDesc* d) {
  Dims m;
  m.D = d->D; m.C = d->C; m.H1 = d->H1; m.R = d->R; m.n = d->n_blocks;
  m.dc = (d->D + 1) / 2;  // torch.tensor_split(x, 2, dim=1): first chunk is the larger one
  m.dt = d->D / 2;
  m.K1 = m.dc + d->R;
  m.eps = d->ln_eps;
  m.prec = d->precision;
  m.fastvar = (d->reserved & NF_FLAG_LN_FAST_VARIANCE) ? 1 : 0;
  return m;
}

// precision modes that run the contractions on the tensor cores; the reduced mode issues one raw TF32 pass
static inline bool prec_tc(int prec) { return prec == NF_PREC_TF32X3 || prec == NF_PREC_TF32; }

int check_desc(const NfFlowDesc* d);
void param_layout(const Dims& m, NfParamLayout* out);

// ---- packed (derived) parameter tables -----------------------------------------------------
// Per block:  W (D*D), Winv (D*D), Lh (D*D), Uh (D*D), PL (D*D) [= P @ Lh], then per net (t, s):
//             W2f (C*H1), b2f (C), W3f (dt*C), b3f (dt), W2fT (H1*C), W1T (K1*H1), W1p (H1*K1p)
//             (the transposes are the K-major B operands of the data-gradient GEMMs).
// After all blocks: logdet_c (n floats, sum log|s| per block) and logdet_total (1 float).
struct PackedLayout {
  int64_t W, Winv, Lh, Uh, PL;
  int64_t WT, WinvT;     // transposes of W / W^-1: the K-major B operands of x @ W and xp @ W^-1 on the tensor cores
  int64_t net0;          // offset of net t tables inside a block
  int64_t W2f, b2f, W3f, b3f, W2fT, W1T, W1p;  // offsets inside a net
  int dcp, K1p;  // W1p: (H1, K1p) copy of W1 whose y columns start at the 16-byte aligned column dcp
  int64_t net_stride, block_stride;
  int64_t logdet_c, logdet_total, total;
  // tensor-core precision only: rounded-to-tf32 `hi` and residual `lo` copies of every table above, at
  // table + hi_off / table + lo_off (the GEMMs' B operands arrive pre-split; 0 = absent)
  int64_t hi_off, lo_off;
};
PackedLayout packed_layout(const Dims& m);

// ---- stash written by forward for backward ---------------------------------------------------
// xs: (n+1, B, D) block inputs (xs[0] = x, xs[i+1] = output of block i)
// ss: (n, B, dt)  log-scales s of every block ; y: (B, R) conditioning input
// then per block, per net (t, s):  xh1 (B*H1), xh2 (B*C), r1 (B), r2 (B), m1 (B*W1w), m2 (B*W2w)
struct StashLayout {
  int64_t xs, ss, y, blocks;   // y: (B, R) copy of the conditioning input (backward reads it from here)
  int64_t xh1, xh2, r1, r2, m1, m2;   // offsets inside a net
  int64_t net_stride, block_stride, total;
  int mw1, mw2;                        // mask words per row
};
StashLayout stash_layout(const Dims& m, int64_t B);

}  // namespace nf
