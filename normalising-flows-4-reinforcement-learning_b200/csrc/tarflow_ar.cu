// Persistent autoregressive-inverse kernel of the causal-transformer flow for evaluation-sized batches (B <= 128).
//
// Sampling one block is A dependent steps (dimension t needs x_0 .. x_{t-1}), each a pass of ONE token per sample through
// the block's layers: with B = 20 environments (utils/torch_evaluations.py:79-82) that is 8 x 4 x 8 launches of kernels whose
// fixed cost (TMEM allocation, TMA round trips, epilogue) is ~8 us each -- the per-token launch sequence takes 8.8 ms per
// sampling call whatever the batch.  Here ONE cooperative launch per block runs the whole loop:
//   * CTA g owns `cp` output columns of every Linear of every layer (3 cp of qkv, E cp of fc1) and keeps those weight rows in
//     shared memory for the whole loop (12.6 MB of block weights / 128 CTAs = 98 KB each at the script's shape), so a step
//     re-reads no weights from L2;
//   * a phase = [activations of a 24-row chunk from L2 into shared memory -> LayerNorm there -> this CTA's slice of the
//     Linear with FFMA (warp = output column, lanes = reduction, weight row in registers) -> slice to L2], phases are
//     separated by a grid barrier (one atomic + spin per CTA): 1 + 5 L barriers per step;
//   * K / V rows go to the same (B, T, 3C) cache the per-token path uses; attention of one (sample, head) per warp;
//   * the tail (output projection, clamp, x_t = z_t exp(xa) + xb) is computed redundantly by every CTA, so the next token's
//     embedding needs no exchange.
// Exact fp32 arithmetic (FFMA), bit-identical across CTAs where replicated.
// Measured (B = 20, the script's shape, one B200): sampling call 8.76 ms (per-token launches) -> 4.27 ms; per block 2.07 M
// cycles of which 33 % grid barriers (176 x ~2 us including the wait for the slowest CTA), 21 % LN1 + qkv, 22 % LN2 + fc1 +
// GELU, 12 % fc2, 8 % proj (tools/tarflow_ar_timeline.py; a grid barrier alone costs 1.2 us, tools/micro/grid_barrier_bench.cu).
#include "tarflow_ar.cuh"

namespace nf { long long* tc_dbg_buffer(); }   // tc_gemm.cu: the buffer set by nf_debug_tc_timeline

#include <algorithm>

namespace nf {
namespace {

constexpr int kThreads = 256, kWarps = 8, kRB = 24;     // rows per chunk (<= 32: one lane per row in the slice epilogue;
                                                         // 24 = the evaluation batch of 20 environments in ONE chunk)
constexpr int kMaxK32 = 32;                              // reduction length <= 1024 (weight row in registers)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float gelu_f(float u) { return 0.5f * u * (1.0f + erff(u * 0.70710678118654752f)); }

// all CTAs of the (co-resident, cooperative) grid; `target` advances by gridDim.x per barrier, the counter never resets
__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned& target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    target += gridDim.x;
    __threadfence();
    atomicAdd(counter, 1u);
    while (*reinterpret_cast<volatile unsigned*>(counter) < target) { }
    __threadfence();
  }
  __syncthreads();
}

// n4 float4 from global to shared memory with eight loads per thread in flight (a plain copy loop exposes one L2 round trip
// per iteration: 16 iterations for a 16 x 1024 chunk were 6 us of a phase).  CG = true: other CTAs wrote the data, bypass L1.
template <bool CG>
__device__ __forceinline__ void copy_f4(float4* dst, const float4* src, int n4) {
  for (int base = 0; base < n4; base += kThreads * 8) {
    float4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = base + u * kThreads + threadIdx.x;
      if (i < n4) v[u] = CG ? __ldcg(src + i) : __ldg(src + i);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = base + u * kThreads + threadIdx.x;
      if (i < n4) dst[i] = v[u];
    }
  }
}
// rows x K activations (dense rows) from L2 into shared memory
__device__ __forceinline__ void load_rows(float* dst, const float* src, long long ld, int rows, int K) {
  (void)ld;     // exchange buffers are dense: ld == K
  copy_f4<true>(reinterpret_cast<float4*>(dst), reinterpret_cast<const float4*>(src), rows * (K >> 2));
}

// in place over rows x C in shared memory: (x - mean) rstd gamma + beta   (warp per row)
__device__ __forceinline__ void layer_norm(float* a, int rows, int C, const float* gamma, const float* beta, float eps) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int r = w; r < rows; r += kWarps) {
    float* x = a + r * C;
    float s = 0.f;
    for (int k = lane; k < C; k += 32) s += x[k];
    const float mean = warp_sum(s) / (float)C;
    float m2 = 0.f;
    for (int k = lane; k < C; k += 32) { const float d = x[k] - mean; m2 = fmaf(d, d, m2); }
    const float rs = 1.0f / sqrtf(warp_sum(m2) / (float)C + eps);
    for (int k = lane; k < C; k += 32) x[k] = fmaf((x[k] - mean) * rs, __ldg(gamma + k), __ldg(beta + k));
  }
}

// out(b, j) = sum_k act[b][k] W[j][k], j < ncols.  A warp takes one column (weight row in registers, lanes = reduction) and
// one group of rows: with few columns per CTA (cp = 2 for proj / fc2) the rows are split over the otherwise idle warps.
// Rows go four at a time so the four shuffle reductions overlap; lane b keeps the result of row b and the epilogue
// (residual / bias loads from L2, the store) runs once per column with all rows in flight.
template <class F>
__device__ __forceinline__ void slice_gemm(const float* act, int rows, int K, const float* W, int ncols, F&& emit) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int nk = K >> 5;
  const int groups = ncols >= kWarps ? 1 : kWarps / ncols;      // row groups per column
  const int rpg = (rows + groups - 1) / groups;
  for (int item = w; item < ncols * groups; item += kWarps) {
    const int j = item % ncols, b0 = (item / ncols) * rpg, b1 = min(rows, b0 + rpg);
    if (b0 >= b1) continue;
    float wr[kMaxK32];
#pragma unroll
    for (int i = 0; i < kMaxK32; ++i) wr[i] = i < nk ? W[(long long)j * K + lane + 32 * i] : 0.f;
    float mine = 0.f;
    for (int b = b0; b < b1; b += 4) {
      float s[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float* ar = act + min(b + u, b1 - 1) * K + lane;
#pragma unroll
        for (int i = 0; i < kMaxK32; ++i)
          if (i < nk) s[u] = fmaf(ar[32 * i], wr[i], s[u]);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int u = 0; u < 4; ++u) s[u] += __shfl_xor_sync(0xffffffffu, s[u], o);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (b + u < b1 && lane == b + u) mine = s[u];
    }
    if (lane >= b0 && lane < b1) emit(lane, j, mine);
  }
}

__global__ void __launch_bounds__(kThreads, 1) k_tf_ar_block(const ArP p) {
  extern __shared__ __align__(16) float sm[];
  NF_PDL_ENTRY();
  const int C = p.C, EC = p.EC, L = p.L, A = p.A, B = p.B, cp = p.cp, E = EC / C;
  const int g = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int per_layer = cp * C * (4 + 2 * E);            // floats of this CTA's weight slices per layer
  float* wsm = sm;                                        // [L][qkv 3cp x C | proj cp x C | fc1 E cp x C | fc2 cp x EC]
  float* act = sm + (p.w_smem ? (size_t)L * per_layer : 0);   // [kRB][max(C, EC)]
  float* abuf = act + (size_t)kRB * EC;                   // [kRB][C]
  float* xprev = abuf + (size_t)kRB * C;                  // [B] the dimension produced by the previous step
  unsigned target = 0;
  long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tlast = clock64();   // (debug timeline, p.dbg)
#define AR_T(i) do { if (p.dbg && threadIdx.x == 0) { const long long n_ = clock64(); tacc[i] += n_ - tlast; tlast = n_; } } while (0)

  // this CTA's weight rows: shared-memory resident for the whole loop when they fit, else read from L2 through the same pointers
  auto wptr = [&](int l, int which) -> const float* {     // which: 0 qkv, 1 proj, 2 fc1, 3 fc2
    const float* lp = p.layers + (long long)l * p.layer_stride;
    if (p.w_smem) {
      const float* base = wsm + (size_t)l * per_layer;
      return which == 0 ? base : which == 1 ? base + 3 * cp * C : which == 2 ? base + 4 * cp * C : base + (4 + E) * cp * C;
    }
    return which == 0 ? lp + p.o_wqkv + (long long)g * 3 * cp * C : which == 1 ? lp + p.o_wproj + (long long)g * cp * C
         : which == 2 ? lp + p.o_w1 + (long long)g * E * cp * C : lp + p.o_w2 + (long long)g * cp * EC;
  };
  if (p.w_smem) {
    for (int l = 0; l < L; ++l) {
      const float* lp = p.layers + (long long)l * p.layer_stride;
      float* base = wsm + (size_t)l * per_layer;
      const float* src[4] = {lp + p.o_wqkv + (long long)g * 3 * cp * C, lp + p.o_wproj + (long long)g * cp * C,
                             lp + p.o_w1 + (long long)g * E * cp * C, lp + p.o_w2 + (long long)g * cp * EC};
      const int n[4] = {3 * cp * C, cp * C, E * cp * C, cp * EC};
      float* dst = base;
      for (int q = 0; q < 4; ++q) {
        copy_f4<false>(reinterpret_cast<float4*>(dst), reinterpret_cast<const float4*>(src[q]), n[q] / 4);
        dst += n[q];
      }
    }
    __syncthreads();
  }

  const float scale = 1.0f / sqrtf((float)p.hd);
  const int HDV = p.hd >> 5;
  for (int t = 0; t < A; ++t) {
    // ---- embedding of token t: this CTA's columns of h                (k_tf_embed of the per-token path)
    for (int i = threadIdx.x; i < B * cp; i += kThreads) {
      const int b = i / cp, n = g * cp + i % cp;
      float v = __ldg(p.pos + (long long)t * C + n);
      v += t == 0 ? __ldg(p.cproj + (long long)b * C + n) : fmaf(xprev[b], __ldg(p.win + n), __ldg(p.bin + n));
      p.h[(long long)b * C + n] = v;
    }
    AR_T(0);
    grid_barrier(p.bar, target);
    AR_T(7);
    for (int l = 0; l < L; ++l) {
      const float* lp = p.layers + (long long)l * p.layer_stride;
      float* qkv = p.qkv + (long long)l * B * A * 3 * C;
      // ---- P1: a = LN1(h); this CTA's 3 cp columns of qkv -> cache row (b, t)
      for (int r0 = 0; r0 < B; r0 += kRB) {
        const int rows = min(kRB, B - r0);
        load_rows(abuf, p.h + (long long)r0 * C, C, rows, C);
        __syncthreads();
        layer_norm(abuf, rows, C, lp + p.o_ln1w, lp + p.o_ln1b, p.eps);
        __syncthreads();
        slice_gemm(abuf, rows, C, wptr(l, 0), 3 * cp, [&](int b, int j, float s) {
          const int n = g * 3 * cp + j;
          qkv[((long long)(r0 + b) * A + t) * 3 * C + n] = s + __ldg(lp + p.o_bqkv + n);
        });
        __syncthreads();
      }
      AR_T(1);
      grid_barrier(p.bar, target);
      AR_T(7);
      // ---- P2: causal attention of (sample, head) pairs over the cached rows 0 .. t, one warp per pair
      for (int pr = g * kWarps + warp; pr < B * p.H; pr += gridDim.x * kWarps) {
        const int b = pr / p.H, head = pr % p.H;
        const float* base = qkv + (long long)b * A * 3 * C + head * p.hd + lane;
        float q[4], acc[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { q[i] = i < HDV ? __ldcg(base + (long long)t * 3 * C + 32 * i) : 0.f; acc[i] = 0.f; }
        float sj = -INFINITY;
        for (int j0 = 0; j0 <= t; j0 += 8) {          // eight keys per batch, their loads issued before the first use
          float kv[8][4];
#pragma unroll
          for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int i = 0; i < 4; ++i)
              kv[u][i] = (j0 + u <= t && i < HDV) ? __ldcg(base + (long long)(j0 + u) * 3 * C + C + 32 * i) : 0.f;
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            if (j0 + u <= t) {
              float part = 0.f;
#pragma unroll
              for (int i = 0; i < 4; ++i) part = fmaf(q[i], kv[u][i], part);
              const float tot = warp_sum(part) * scale;
              if (lane == j0 + u) sj = tot;
            }
          }
        }
        const float mx = warp_max(sj);
        const float e = lane <= t ? expf(sj - mx) : 0.f;
        const float pj_all = e / warp_sum(e);
        for (int j0 = 0; j0 <= t; j0 += 8) {
          float vv[8][4];
#pragma unroll
          for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int i = 0; i < 4; ++i)
              vv[u][i] = (j0 + u <= t && i < HDV) ? __ldcg(base + (long long)(j0 + u) * 3 * C + 2 * C + 32 * i) : 0.f;
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            if (j0 + u <= t) {
              const float pj = __shfl_sync(0xffffffffu, pj_all, j0 + u);
#pragma unroll
              for (int i = 0; i < 4; ++i) acc[i] = fmaf(pj, vv[u][i], acc[i]);
            }
          }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) if (i < HDV) p.o[(long long)b * C + head * p.hd + lane + 32 * i] = acc[i];
      }
      AR_T(2);
      grid_barrier(p.bar, target);
      AR_T(7);
      // ---- P3: h2 = h + proj(o): this CTA's cp columns
      for (int r0 = 0; r0 < B; r0 += kRB) {
        const int rows = min(kRB, B - r0);
        load_rows(abuf, p.o + (long long)r0 * C, C, rows, C);
        __syncthreads();
        slice_gemm(abuf, rows, C, wptr(l, 1), cp, [&](int b, int j, float s) {
          const int n = g * cp + j;
          const long long idx = (long long)(r0 + b) * C + n;
          p.h2[idx] = __ldcg(p.h + idx) + s + __ldg(lp + p.o_bproj + n);
        });
        __syncthreads();
      }
      AR_T(3);
      grid_barrier(p.bar, target);
      AR_T(7);
      // ---- P4: m = LN2(h2); this CTA's E cp columns of GELU(fc1(m))
      for (int r0 = 0; r0 < B; r0 += kRB) {
        const int rows = min(kRB, B - r0);
        load_rows(abuf, p.h2 + (long long)r0 * C, C, rows, C);
        __syncthreads();
        layer_norm(abuf, rows, C, lp + p.o_ln2w, lp + p.o_ln2b, p.eps);
        __syncthreads();
        slice_gemm(abuf, rows, C, wptr(l, 2), E * cp, [&](int b, int j, float s) {
          const int n = g * E * cp + j;
          p.g[(long long)(r0 + b) * EC + n] = gelu_f(s + __ldg(lp + p.o_b1 + n));
        });
        __syncthreads();
      }
      AR_T(4);
      grid_barrier(p.bar, target);
      AR_T(7);
      // ---- P5: h = h2 + fc2(g): this CTA's cp columns
      for (int r0 = 0; r0 < B; r0 += kRB) {
        const int rows = min(kRB, B - r0);
        load_rows(act, p.g + (long long)r0 * EC, EC, rows, EC);
        __syncthreads();
        slice_gemm(act, rows, EC, wptr(l, 3), cp, [&](int b, int j, float s) {
          const int n = g * cp + j;
          const long long idx = (long long)(r0 + b) * C + n;
          p.h[idx] = __ldcg(p.h2 + idx) + s + __ldg(lp + p.o_b2 + n);
        });
        __syncthreads();
      }
      AR_T(5);
      grid_barrier(p.bar, target);
      AR_T(7);
    }
    // ---- tail, replicated in every CTA: (raw, xb) = h . wout + bout; x_t = z_t exp(xa) + xb   (k_tf_tail, inverse)
    const int col = p.flip ? A - 1 - t : t;
    for (int r0 = 0; r0 < B; r0 += kRB) {
      const int rows = min(kRB, B - r0);
      load_rows(abuf, p.h + (long long)r0 * C, C, rows, C);
      __syncthreads();
      for (int r = warp; r < rows; r += kWarps) {
        const float* hr = abuf + r * C;
        float s0 = 0.f, s1 = 0.f;
        for (int k = lane; k < C; k += 32) { const float v = hr[k]; s0 = fmaf(v, __ldg(p.wout + k), s0); s1 = fmaf(v, __ldg(p.wout + C + k), s1); }
        const float raw = warp_sum(s0) + __ldg(p.bout), xb = warp_sum(s1) + __ldg(p.bout + 1);
        const float xa = p.clamp > 0.f ? p.clamp * tanhf(raw / p.clamp) : raw;
        const long long b = r0 + r;
        const float res = fmaf(__ldg(p.zin + b * A + col), expf(xa), xb);
        if (lane == 0) {
          xprev[b] = res;
          if (g == 0) {
            p.xout[b * A + col] = res;
            if (p.logdet) p.logdet[b] = ((p.ld_first && t == 0) ? 0.f : p.logdet[b]) + xa;
          }
        }
      }
      __syncthreads();
    }
    AR_T(6);
    grid_barrier(p.bar, target);   // every CTA has read h before the next token's embedding overwrites it
    AR_T(7);
  }
  if (p.dbg && threadIdx.x == 0)
    for (int i = 0; i < 8; ++i) p.dbg[(long long)blockIdx.x * 8 + i] = tacc[i];
#undef AR_T
}

int g_sm_count[64] = {};
int sm_count() {
  int d = 0;
  if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64) return 0;
  if (!g_sm_count[d]) {
    int n = 0, coop = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, d);
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, d);
    g_sm_count[d] = coop ? n : -1;
  }
  return g_sm_count[d];
}

size_t ar_smem(const ArShape& s, int cp, bool w_smem) {
  const int E = s.EC / s.C;
  const size_t w = w_smem ? (size_t)s.L * cp * s.C * (4 + 2 * E) : 0;
  return (w + (size_t)kRB * s.EC + (size_t)kRB * s.C + (size_t)((s.B + 3) & ~3)) * sizeof(float);
}

}  // namespace

bool tf_ar_plan(const ArShape& s, int* cp_out, int* w_smem_out) {
  const int sms = sm_count();
  if (sms <= 0 || s.B < 1 || s.B > 128 || s.C % 32 != 0 || s.EC % 32 != 0 || s.EC > 32 * kMaxK32 || s.C > 32 * kMaxK32 ||
      s.EC % s.C != 0 || s.hd > 128 || s.A > 32)
    return false;
  int cp = 0;
  for (int c = 1; c <= 32; c <<= 1)
    if (s.C % c == 0 && s.C / c <= sms) { cp = c; break; }
  if (!cp) return false;
  constexpr size_t kLimit = 220 * 1024;
  const bool w = ar_smem(s, cp, true) <= kLimit;
  if (ar_smem(s, cp, w) > kLimit) return false;
  if (cp_out) *cp_out = cp;
  if (w_smem_out) *w_smem_out = w ? 1 : 0;
  return true;
}

int tf_ar_block(ArP p, cudaStream_t st) {
  ArShape s{p.B, p.A, p.C, p.EC, p.L, p.hd};
  int cp = 0, w = 0;
  NF_CHECK_ARG(tf_ar_plan(s, &cp, &w), NF_E_UNSUPPORTED, "tf_ar_block: shape not supported (B=%d C=%d EC=%d)", p.B, p.C, p.EC);
  p.cp = cp; p.w_smem = w;
  p.dbg = tc_dbg_buffer();
  const size_t smem = ar_smem(s, cp, w != 0);
  {
    static PerDeviceOnce once;
    if (once.first()) NF_CUDA(cudaFuncSetAttribute(k_tf_ar_block, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(p.C / cp));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;      // every CTA co-resident: the grid barrier cannot deadlock
  attr[0].val.cooperative = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  NF_CUDA(cudaLaunchKernelEx(&cfg, k_tf_ar_block, p));
  NF_LAUNCHED();
  return 0;
}

}  // namespace nf
