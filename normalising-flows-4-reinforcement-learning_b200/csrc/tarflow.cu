// Causal-transformer flow: the model the goal-conditioned BC script builds with
// RealNVP(input_dims, channels, rep_dim, head_dim, blocks, layers_per_block, prior) (il-and-cil/torch_nfgcbc.py:27,119),
// calls as model(actions (B, A, 1), cond (B, 1, rep)) -> (z, outputs, logdets) (:136-142) and samples from with
// model.reverse(z, cond) (utils/torch_evaluations.py:79-82).  Its source is absent upstream (PARITY UNPINNED); the
// arithmetic is the published TarFlow meta block as restated in oracle/torch_tarflow.py.
//
// Layout: activations are (B * T, C) row-major, T = A tokens per sample (conditioning token, then action dimensions
// 0 .. A-2).  Every Linear, data gradient and weight gradient goes through the conditioner-layer dispatch of the
// coupling flow (tcgen05 3xTF32 / FFMA, linear.cuh); bias gradients ride on the weight-gradient GEMMs.  Row kernels here:
//   k_tf_embed        tokens -> h0 = [cond_proj(y) | x_t w_in + b_in] + pos_embed
//   k_tf_add_ln       h = h + r (residual), a = LayerNorm(h) gamma + beta, (mean, rstd) kept for backward
//   k_tf_attn_fwd     causal softmax attention over the <= 32 tokens of one (sample, head): one warp, lane = head
//                     dimension for the dot products, lane = key for the softmax
//   k_tf_gelu(+bwd)   exact (erf) GELU
//   k_tf_tail         output projection (2 x C), log-scale clamp, affine transform, log-det; forward, one inverse step
//                     (which also embeds the dimension it has just produced as the next token) and backward
//   k_tf_ln_bwd, k_tf_attn_bwd, k_tf_embed_bwd, k_tf_colsum   the backward row kernels
// Inverse: the autoregressive loop of A steps per block; step t runs token t (B rows) through the layers, writes its
// K / V rows into the per-layer cache (the qkv buffer itself, row pitch T * 3C) and attends over rows <= t.  All of it
// is one cached launch graph per (desc, B).
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "elementwise.cuh"
#include "graphs.cuh"
#include "linear.cuh"
#include "prof.cuh"
#include "tarflow_ar.cuh"

namespace nf {
namespace {

constexpr int kTfMaxV = 8;     // float4 per lane: C <= 1024
constexpr int kTfMaxT = 32;    // tokens: lane = key in the softmax

struct TfDims {
  int A, T, C, R, hd, H, nb, L, E, EC, prec;
  float eps, clamp;
};

int check_tf(const NfTarflowDesc* d) {
  NF_CHECK_ARG(d != nullptr, NF_E_BADARG, "tarflow desc is NULL");
  NF_CHECK_ARG(d->A >= 1 && d->A <= kTfMaxT, NF_E_UNSUPPORTED, "tarflow: A=%d must be 1..%d", d->A, kTfMaxT);
  NF_CHECK_ARG(d->C >= 32 && d->C % 4 == 0 && d->C <= 128 * kTfMaxV, NF_E_UNSUPPORTED,
               "tarflow: channels=%d must be a multiple of 4 in 32..%d", d->C, 128 * kTfMaxV);
  NF_CHECK_ARG(d->head_dim == 32 || d->head_dim == 64 || d->head_dim == 128, NF_E_UNSUPPORTED,
               "tarflow: head_dim=%d must be 32, 64 or 128", d->head_dim);
  NF_CHECK_ARG(d->C % d->head_dim == 0, NF_E_BADARG, "tarflow: head_dim=%d does not divide channels=%d", d->head_dim, d->C);
  NF_CHECK_ARG(d->R >= 1 && d->n_blocks >= 1 && d->n_blocks <= 64 && d->n_layers >= 1 && d->n_layers <= 64 &&
               d->expansion >= 1 && d->expansion <= 8, NF_E_BADARG, "tarflow: R=%d blocks=%d layers=%d expansion=%d",
               d->R, d->n_blocks, d->n_layers, d->expansion);
  NF_CHECK_ARG(d->ln_eps > 0.f && d->logscale_clamp >= 0.f, NF_E_BADARG, "tarflow: ln_eps / logscale_clamp out of range");
  NF_CHECK_ARG(d->precision == NF_PREC_FP32_SIMT || d->precision == NF_PREC_TF32X3, NF_E_UNSUPPORTED,
               "tarflow: precision %d is not built", d->precision);
  return 0;
}

TfDims tf_dims(const NfTarflowDesc* d) {
  TfDims m;
  m.A = d->A; m.T = d->A; m.C = d->C; m.R = d->R; m.hd = d->head_dim; m.H = d->C / d->head_dim;
  m.nb = d->n_blocks; m.L = d->n_layers; m.E = d->expansion; m.EC = d->expansion * d->C; m.prec = d->precision;
  m.eps = d->ln_eps; m.clamp = d->logscale_clamp;
  return m;
}

// ---- flat parameter arena ------------------------------------------------------------------------------------------
struct TfLayout {
  int64_t pos, win, bin, wc, bc, layers, layer_stride, wout, bout, block_stride, total;
  int64_t ln1w, ln1b, wqkv, bqkv, wproj, bproj, ln2w, ln2b, w1, b1, w2, b2;   // inside a layer
};
TfLayout tf_layout(const TfDims& m) {
  TfLayout l{};
  int64_t off = 0;
  auto take = [&](int64_t n) { int64_t o = off; off += pad4(n); return o; };
  const int64_t C = m.C;
  l.ln1w = take(C); l.ln1b = take(C); l.wqkv = take(3 * C * C); l.bqkv = take(3 * C);
  l.wproj = take(C * C); l.bproj = take(C); l.ln2w = take(C); l.ln2b = take(C);
  l.w1 = take((int64_t)m.EC * C); l.b1 = take(m.EC); l.w2 = take(C * (int64_t)m.EC); l.b2 = take(C);
  l.layer_stride = off;
  off = 0;
  l.pos = take((int64_t)m.T * C); l.win = take(C); l.bin = take(C); l.wc = take(C * (int64_t)m.R); l.bc = take(C);
  l.layers = off; off += l.layer_stride * m.L;
  l.wout = take(2 * C); l.bout = take(2);
  l.block_stride = off;
  l.total = off * m.nb;
  return l;
}

// ---- stash (forward -> backward) -------------------------------------------------------------------------------------
// y (B, R); xs (nb + 1, B, A): xs[0] = x, xs[i + 1] = output of block i; xa (nb, B, A) clamped log-scales in token order;
// per block: per layer {h_in, a, qkv, o, h2, mm, u, g, st1, st2, p}, then h_final
struct TfStash {
  int64_t y, xs, xa, blocks, block_stride, total;
  int64_t h_in, a, qkv, o, h2, mm, u, g, st1, st2, p, layer_stride;   // inside a layer
  int64_t hf;                                                          // inside a block, after the layers
};
TfStash tf_stash(const TfDims& m, int64_t B) {
  TfStash s{};
  const int64_t M = B * m.T, MC = M * m.C;
  int64_t off = 0;
  auto take = [&](int64_t n) { int64_t o = off; off += pad4(n); return o; };
  s.h_in = take(MC); s.a = take(MC); s.qkv = take(3 * MC); s.o = take(MC); s.h2 = take(MC); s.mm = take(MC);
  s.u = take(M * m.EC); s.g = take(M * m.EC); s.st1 = take(2 * M); s.st2 = take(2 * M);
  s.p = take(B * m.H * m.T * m.T);
  s.layer_stride = off;
  s.hf = s.layer_stride * m.L;
  s.block_stride = s.hf + pad4(MC);
  off = 0;
  s.y = take(B * m.R); s.xs = take((int64_t)(m.nb + 1) * B * m.A); s.xa = take((int64_t)m.nb * B * m.A);
  s.blocks = off;
  s.total = off + s.block_stride * m.nb;
  return s;
}

// ---- workspace ---------------------------------------------------------------------------------------------------------
struct TfWs {
  int64_t io_x, io_y, io_z, io_ld, io_dz, io_dld, io_dx, io_dy, x0, x1, xa, cproj;
  int64_t tmp, bufA, bufB, bufO, big, big2, qkv, kv, wT, bars, total;
  int64_t t_qkv, t_proj, t_w1, t_w2, t_layer_stride, t_wc, t_block_stride;   // transposed weights inside wT
};
TfWs tf_ws(const TfDims& m, int64_t B) {
  TfWs w{};
  const int64_t M = B * m.T, MC = M * m.C, C = m.C;
  int64_t off = 0;
  auto take = [&](int64_t n) { int64_t o = off; off += pad4(n); return o; };
  w.io_x = take(B * m.A); w.io_y = take(B * m.R); w.io_z = take(B * m.A); w.io_ld = take(B);
  w.io_dz = take(B * m.A); w.io_dld = take(B); w.io_dx = take(B * m.A); w.io_dy = take(B * m.R);
  w.x0 = take(B * m.A); w.x1 = take(B * m.A); w.xa = take(B * m.A); w.cproj = take(B * C);
  w.tmp = take(MC); w.bufA = take(MC); w.bufB = take(MC); w.bufO = take(MC);
  w.big = take(M * m.EC); w.big2 = take(4 * C + 16);   // big2: sink for unwanted parameter gradients
  w.qkv = take(3 * MC);
  w.kv = take(3 * MC * m.L);          // inverse: K / V rows of every layer (the qkv rows of the tokens seen so far)
  w.t_qkv = 0; w.t_proj = 3 * C * C; w.t_w1 = w.t_proj + C * C; w.t_w2 = w.t_w1 + C * (int64_t)m.EC;
  w.t_layer_stride = w.t_w2 + C * (int64_t)m.EC;
  w.t_wc = w.t_layer_stride * m.L;
  w.t_block_stride = w.t_wc + pad4(C * (int64_t)m.R);
  w.wT = take(w.t_block_stride * m.nb);
  w.bars = take(64);                    // grid-barrier counters of the persistent inverse kernel (one per block)
  w.total = off;
  return w;
}

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
__device__ __forceinline__ float dot4(const float4& a, const float4& b) {
  return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w)));
}

// h[b, t - t0, :] = pos[t, :] + (t == 0 ? cproj[b, :] : x[b, perm(t - 1)] w_in + b_in), t in [t0, t1)
__global__ void __launch_bounds__(256) k_tf_embed(const float* __restrict__ x, const float* __restrict__ cproj,
                                                  const float* __restrict__ win, const float* __restrict__ bin,
                                                  const float* __restrict__ pos, float* __restrict__ h, long long B,
                                                  int A, int C, int t0, int t1, int flip) {
  NF_PDL_ENTRY();
  const int C4 = C >> 2, nt = t1 - t0;
  const long long n = B * nt * C4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % C4);
    const long long r = i / C4;
    const int t = t0 + (int)(r % nt);
    const long long b = r / nt;
    float4 v = __ldg(reinterpret_cast<const float4*>(pos + (long long)t * C) + c4);
    if (t == 0) {
      const float4 c = __ldg(reinterpret_cast<const float4*>(cproj + b * C) + c4);
      v.x += c.x; v.y += c.y; v.z += c.z; v.w += c.w;
    } else {
      const float xv = __ldg(x + b * A + (flip ? A - t : t - 1));
      const float4 w = __ldg(reinterpret_cast<const float4*>(win) + c4);
      const float4 bb = __ldg(reinterpret_cast<const float4*>(bin) + c4);
      v.x += fmaf(xv, w.x, bb.x); v.y += fmaf(xv, w.y, bb.y); v.z += fmaf(xv, w.z, bb.z); v.w += fmaf(xv, w.w, bb.w);
    }
    reinterpret_cast<float4*>(h)[i] = v;
  }
}

// h_out = h_in (+ r); a = LayerNorm(h_out) gamma + beta; stats[row] = (mean, rstd).  Warp per row, row in registers.
// gamma == NULL: only the residual add.  h_out may alias h_in; rows are dense (pitch C).
template <int NV>
__global__ void __launch_bounds__(256) k_tf_add_ln(const float* h_in, const float* __restrict__ r,
                                                   float* h_out, float* __restrict__ a,
                                                   float2* __restrict__ stats, const float* __restrict__ gamma,
                                                   const float* __restrict__ beta, long long rows, int C, float eps) {
  NF_PDL_ENTRY();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long row = (long long)blockIdx.x * 8 + warp;
  if (row >= rows) return;
  const float4* hr = reinterpret_cast<const float4*>(h_in + row * C);
  const float4* rr = r ? reinterpret_cast<const float4*>(r + row * C) : nullptr;
  float4 v[NV];
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int c4 = lane + 32 * k;
    v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (4 * c4 < C) {
      v[k] = hr[c4];
      if (rr) { const float4 q = rr[c4]; v[k].x += q.x; v[k].y += q.y; v[k].z += q.z; v[k].w += q.w; }
    }
    sum += (v[k].x + v[k].y) + (v[k].z + v[k].w);
  }
  if (h_out && (rr || h_out != h_in)) {
    float4* ho = reinterpret_cast<float4*>(h_out + row * C);
#pragma unroll
    for (int k = 0; k < NV; ++k) if (4 * (lane + 32 * k) < C) ho[lane + 32 * k] = v[k];
  }
  if (!gamma) return;
  const float mean = warp_sum(sum) / (float)C;
  float m2 = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    if (4 * (lane + 32 * k) < C) {
      const float p = v[k].x - mean, q = v[k].y - mean, s = v[k].z - mean, u = v[k].w - mean;
      m2 += (p * p + q * q) + (s * s + u * u);
    }
  }
  const float rs = 1.0f / sqrtf(warp_sum(m2) / (float)C + eps);
  float4* ar = reinterpret_cast<float4*>(a + row * C);
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int c4 = lane + 32 * k;
    if (4 * c4 < C) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c4);
      const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + c4);
      float4 o;
      o.x = fmaf((v[k].x - mean) * rs, g.x, b.x); o.y = fmaf((v[k].y - mean) * rs, g.y, b.y);
      o.z = fmaf((v[k].z - mean) * rs, g.z, b.z); o.w = fmaf((v[k].w - mean) * rs, g.w, b.w);
      ar[c4] = o;
    }
  }
  if (lane == 0 && stats) stats[row] = make_float2(mean, rs);
}

// dh_out = (dres) + rstd (da gamma - mean(da gamma) - x_hat mean(da gamma x_hat)), x_hat = (h - mean) rstd;
// dgamma += colsum(da x_hat), dbeta += colsum(da).  Warp per row over the CTA's rows, partial sums through shared memory.
template <int NV>
__global__ void __launch_bounds__(256, 2) k_tf_ln_bwd(const float* __restrict__ da, const float* __restrict__ h,
                                                      const float2* __restrict__ stats, const float* __restrict__ gamma,
                                                      const float* __restrict__ dres, float* __restrict__ dh_out,
                                                      float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                      long long rows, int C, int rows_per_cta) {
  extern __shared__ float smem_f[];   // gamma [C], then the flush area [8 warps][2][C]
  NF_PDL_ENTRY();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* sg = smem_f;
  float* part = smem_f + C;
  for (int i = threadIdx.x; i < C; i += blockDim.x) sg[i] = __ldg(gamma + i);
  __syncthreads();
  float4 ag[NV], ab[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) ag[k] = ab[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  const long long r0 = (long long)blockIdx.x * rows_per_cta, r1 = min(rows, r0 + rows_per_cta);
  const float invC = 1.0f / (float)C;
  for (long long row = r0 + warp; row < r1; row += 8) {
    const float4* dr = reinterpret_cast<const float4*>(da + row * C);
    const float4* hr = reinterpret_cast<const float4*>(h + row * C);
    const float2 st = stats[row];
    float4 d[NV], x[NV];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c4 = lane + 32 * k;
      const bool ok = 4 * c4 < C;
      d[k] = ok ? dr[c4] : make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 hv = ok ? hr[c4] : make_float4(st.x, st.x, st.x, st.x);
      x[k].x = (hv.x - st.x) * st.y; x[k].y = (hv.y - st.x) * st.y; x[k].z = (hv.z - st.x) * st.y; x[k].w = (hv.w - st.x) * st.y;
      ag[k].x = fmaf(d[k].x, x[k].x, ag[k].x); ag[k].y = fmaf(d[k].y, x[k].y, ag[k].y);
      ag[k].z = fmaf(d[k].z, x[k].z, ag[k].z); ag[k].w = fmaf(d[k].w, x[k].w, ag[k].w);
      ab[k].x += d[k].x; ab[k].y += d[k].y; ab[k].z += d[k].z; ab[k].w += d[k].w;
      const float4 gm = ok ? *reinterpret_cast<const float4*>(sg + 4 * c4) : make_float4(0.f, 0.f, 0.f, 0.f);
      d[k].x *= gm.x; d[k].y *= gm.y; d[k].z *= gm.z; d[k].w *= gm.w;
      s1 += (d[k].x + d[k].y) + (d[k].z + d[k].w);
      s2 += dot4(d[k], x[k]);
    }
    const float m1 = warp_sum(s1) * invC, m2 = warp_sum(s2) * invC;
    const float4* rr = dres ? reinterpret_cast<const float4*>(dres + row * C) : nullptr;
    float4* out = reinterpret_cast<float4*>(dh_out + row * C);
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c4 = lane + 32 * k;
      if (4 * c4 < C) {
        float4 o;
        o.x = st.y * (d[k].x - m1 - x[k].x * m2); o.y = st.y * (d[k].y - m1 - x[k].y * m2);
        o.z = st.y * (d[k].z - m1 - x[k].z * m2); o.w = st.y * (d[k].w - m1 - x[k].w * m2);
        if (rr) { const float4 q = rr[c4]; o.x += q.x; o.y += q.y; o.z += q.z; o.w += q.w; }
        out[c4] = o;
      }
    }
  }
  float* mine = part + (size_t)warp * 2 * C;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int c = 4 * (lane + 32 * k);
    if (c < C) {
      *reinterpret_cast<float4*>(mine + c) = ag[k];
      *reinterpret_cast<float4*>(mine + C + c) = ab[k];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) v += part[(size_t)w * 2 * C + i];
    atomicAdd(i < C ? dgamma + i : dbeta + (i - C), v);
  }
}

// Causal attention of one (sample, head) per warp.  qkv: (B, T, 3C) rows [q | k | v], heads of hd = 32 HDV contiguous
// columns.  Queries t in [t0, t1); o: (B, t1 - t0, C); P (optional): (B, H, T, T) softmax rows, zero above the diagonal.
template <int HDV>
__global__ void __launch_bounds__(128) k_tf_attn_fwd(const float* __restrict__ qkv, float* __restrict__ o,
                                                     float* __restrict__ P, long long B, int T, int C, int H, int t0,
                                                     int t1, float scale) {
  NF_PDL_ENTRY();
  const int lane = threadIdx.x & 31;
  const long long w = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
  if (w >= B * H) return;
  const long long b = w / H;
  const int head = (int)(w % H);
  const float* base = qkv + b * T * 3 * C + head * (32 * HDV) + lane;
  for (int t = t0; t < t1; ++t) {
    float q[HDV];
#pragma unroll
    for (int i = 0; i < HDV; ++i) q[i] = base[(long long)t * 3 * C + 32 * i];
    float sj = -INFINITY;
    for (int j = 0; j <= t; ++j) {
      const float* kr = base + (long long)j * 3 * C + C;
      float part = 0.f;
#pragma unroll
      for (int i = 0; i < HDV; ++i) part = fmaf(q[i], kr[32 * i], part);
      const float tot = warp_sum(part) * scale;
      if (lane == j) sj = tot;
    }
    const float mx = warp_max(sj);
    const float e = lane <= t ? expf(sj - mx) : 0.f;
    const float p = e / warp_sum(e);
    if (P && lane < T) P[(w * T + t) * T + lane] = p;
    float acc[HDV];
#pragma unroll
    for (int i = 0; i < HDV; ++i) acc[i] = 0.f;
    for (int j = 0; j <= t; ++j) {
      const float pj = __shfl_sync(0xffffffffu, p, j);
      const float* vr = base + (long long)j * 3 * C + 2 * C;
#pragma unroll
      for (int i = 0; i < HDV; ++i) acc[i] = fmaf(pj, vr[32 * i], acc[i]);
    }
    float* orow = o + (b * (t1 - t0) + (t - t0)) * C + head * (32 * HDV) + lane;
#pragma unroll
    for (int i = 0; i < HDV; ++i) orow[32 * i] = acc[i];
  }
}

// Backward of the above over all T queries: dqkv (B, T, 3C) from do (B, T, C), the stashed qkv / o / P.
//   delta_t = do_t . o_t;  dS[t, j] = P[t, j] (do_t . v_j - delta_t) scale;  dq_t = sum_j dS[t, j] k_j;
//   dk_j = sum_t dS[t, j] q_t;  dv_j = sum_t P[t, j] do_t.   Lane j keeps column j of dS and P (TMAX registers each).
template <int HDV, int TMAX>
__global__ void __launch_bounds__(128) k_tf_attn_bwd(const float* __restrict__ qkv, const float* __restrict__ o,
                                                     const float* __restrict__ P, const float* __restrict__ d_o,
                                                     float* __restrict__ dqkv, long long B, int T, int C, int H,
                                                     float scale) {
  NF_PDL_ENTRY();
  const int lane = threadIdx.x & 31;
  const long long w = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
  if (w >= B * H) return;
  const long long b = w / H;
  const int head = (int)(w % H);
  const int hoff = head * (32 * HDV) + lane;
  const float* base = qkv + b * T * 3 * C + hoff;
  const float* ob = o + b * T * C + hoff;
  const float* dob = d_o + b * T * C + hoff;
  float* dbase = dqkv + b * T * 3 * C + hoff;
  float ds[TMAX], pc[TMAX];
#pragma unroll
  for (int t = 0; t < TMAX; ++t) {
    ds[t] = 0.f; pc[t] = 0.f;
    if (t < T) {
      float dov[HDV], part = 0.f;
#pragma unroll
      for (int i = 0; i < HDV; ++i) { dov[i] = dob[(long long)t * C + 32 * i]; part = fmaf(dov[i], ob[(long long)t * C + 32 * i], part); }
      const float delta = warp_sum(part);
      const float pl = (lane < T) ? P[(w * T + t) * T + lane] : 0.f;     // zero above the diagonal
      float mine = 0.f;
      for (int j = 0; j <= t; ++j) {
        const float* vr = base + (long long)j * 3 * C + 2 * C;
        float pp = 0.f;
#pragma unroll
        for (int i = 0; i < HDV; ++i) pp = fmaf(dov[i], vr[32 * i], pp);
        const float dp = warp_sum(pp);
        if (lane == j) mine = dp;
      }
      pc[t] = pl;
      ds[t] = pl * (mine - delta) * scale;
      // dq_t = sum_j dS[t, j] k_j
      float acc[HDV];
#pragma unroll
      for (int i = 0; i < HDV; ++i) acc[i] = 0.f;
      for (int j = 0; j <= t; ++j) {
        const float dsj = __shfl_sync(0xffffffffu, ds[t], j);
        const float* kr = base + (long long)j * 3 * C + C;
#pragma unroll
        for (int i = 0; i < HDV; ++i) acc[i] = fmaf(dsj, kr[32 * i], acc[i]);
      }
#pragma unroll
      for (int i = 0; i < HDV; ++i) dbase[(long long)t * 3 * C + 32 * i] = acc[i];
    }
  }
  for (int j = 0; j < T; ++j) {
    float dk[HDV], dv[HDV];
#pragma unroll
    for (int i = 0; i < HDV; ++i) dk[i] = dv[i] = 0.f;
#pragma unroll
    for (int t = 0; t < TMAX; ++t) {
      if (t < T && t >= j) {
        const float dsj = __shfl_sync(0xffffffffu, ds[t], j);
        const float pj = __shfl_sync(0xffffffffu, pc[t], j);
#pragma unroll
        for (int i = 0; i < HDV; ++i) {
          dk[i] = fmaf(dsj, base[(long long)t * 3 * C + 32 * i], dk[i]);
          dv[i] = fmaf(pj, dob[(long long)t * C + 32 * i], dv[i]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < HDV; ++i) {
      dbase[(long long)j * 3 * C + C + 32 * i] = dk[i];
      dbase[(long long)j * 3 * C + 2 * C + 32 * i] = dv[i];
    }
  }
}

__device__ __forceinline__ float gelu_f(float u) { return 0.5f * u * (1.0f + erff(u * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_d(float u) {
  return 0.5f * (1.0f + erff(u * 0.70710678118654752f)) + u * 0.39894228040143268f * expf(-0.5f * u * u);
}
__global__ void __launch_bounds__(256) k_tf_gelu(const float4* u, float4* g, long long n4) {
  NF_PDL_ENTRY();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = u[i];
    g[i] = make_float4(gelu_f(v.x), gelu_f(v.y), gelu_f(v.z), gelu_f(v.w));
  }
}
// dg <- dg * GELU'(u)
__global__ void __launch_bounds__(256) k_tf_gelu_bwd(float4* __restrict__ dg, const float4* __restrict__ u, long long n4) {
  NF_PDL_ENTRY();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = u[i];
    float4 d = dg[i];
    d.x *= gelu_d(v.x); d.y *= gelu_d(v.y); d.z *= gelu_d(v.z); d.w *= gelu_d(v.w);
    dg[i] = d;
  }
}

// Output projection + affine transform of the tokens t in [t0, t1) of every sample (warp per sample).
//   (raw, xb) = (h_t (+ r_t)) . wout[0 / 1] + bout;  xa = clamp ? clamp tanh(raw / clamp) : raw;  column = perm(t)
//   forward: out[col] = (in[col] - xb) exp(-xa), logdet -= xa;   inverse: out[col] = in[col] exp(xa) + xb, logdet += xa
// h / r rows: (B, t1 - t0, C).  inverse additionally embeds the produced dimension as token t1 (hnext (B, C)) when t1 < T.
template <int NV>
__global__ void __launch_bounds__(256) k_tf_tail(const float* h, const float* __restrict__ r,
                                                 const float* __restrict__ wout, const float* __restrict__ bout,
                                                 const float* __restrict__ in, float* __restrict__ out,
                                                 float* __restrict__ logdet, float* __restrict__ xa_out, long long B, int A,
                                                 int C, int t0, int t1, int flip, float clamp, int inverse, int ld_first,
                                                 const float* __restrict__ win, const float* __restrict__ bin,
                                                 const float* __restrict__ pos, float* hnext) {
  NF_PDL_ENTRY();
  const int lane = threadIdx.x & 31;
  const long long b = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (b >= B) return;
  float4 w0[NV], w1[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int c4 = lane + 32 * k;
    const bool ok = 4 * c4 < C;
    w0[k] = ok ? __ldg(reinterpret_cast<const float4*>(wout) + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
    w1[k] = ok ? __ldg(reinterpret_cast<const float4*>(wout + C) + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const float b0 = __ldg(bout), b1 = __ldg(bout + 1);
  float ld = 0.f, last = 0.f;
  const int nt = t1 - t0;
  for (int t = t0; t < t1; ++t) {
    const long long row = b * nt + (t - t0);
    const float4* hr = reinterpret_cast<const float4*>(h + row * C);
    const float4* rr = r ? reinterpret_cast<const float4*>(r + row * C) : nullptr;
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c4 = lane + 32 * k;
      if (4 * c4 < C) {
        float4 v = hr[c4];
        if (rr) { const float4 q = rr[c4]; v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w; }
        s0 += dot4(v, w0[k]);
        s1 += dot4(v, w1[k]);
      }
    }
    const float raw = warp_sum(s0) + b0, xb = warp_sum(s1) + b1;
    const float xa = clamp > 0.f ? clamp * tanhf(raw / clamp) : raw;
    const int col = flip ? A - 1 - t : t;
    const float v = in[b * A + col];
    const float res = inverse ? fmaf(v, expf(xa), xb) : (v - xb) * expf(-xa);
    ld += inverse ? xa : -xa;
    last = res;
    if (lane == 0) {
      out[b * A + col] = res;
      if (xa_out) xa_out[b * A + t] = xa;
    }
  }
  if (lane == 0 && logdet) logdet[b] = (ld_first ? 0.f : logdet[b]) + ld;
  if (inverse && hnext && t1 < A) {      // token t1 = the dimension just produced
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c4 = lane + 32 * k;
      if (4 * c4 < C) {
        const float4 w = __ldg(reinterpret_cast<const float4*>(win) + c4);
        const float4 bb = __ldg(reinterpret_cast<const float4*>(bin) + c4);
        const float4 p = __ldg(reinterpret_cast<const float4*>(pos + (long long)t1 * C) + c4);
        reinterpret_cast<float4*>(hnext + b * C)[c4] =
            make_float4(fmaf(last, w.x, bb.x) + p.x, fmaf(last, w.y, bb.y) + p.y, fmaf(last, w.z, bb.z) + p.z,
                        fmaf(last, w.w, bb.w) + p.w);
      }
    }
  }
}

// Backward of the forward tail for all A tokens: dh (B, T, C), dx_direct (B, A), dwout (2, C) +=, dbout (2) +=.
//   g = dz[col];  dx[col] = g e^-xa;  dxb = -g e^-xa;  dxa = -g z[col] - dlogdet;  draw = dxa (1 - (xa / clamp)^2)
template <int NV>
__global__ void __launch_bounds__(256) k_tf_tail_bwd(const float* __restrict__ dz, const float* __restrict__ dld,
                                                     const float* __restrict__ z, const float* __restrict__ xa_st,
                                                     const float* __restrict__ hf, const float* __restrict__ wout,
                                                     float* __restrict__ dh, float* __restrict__ dx,
                                                     float* __restrict__ dwout, float* __restrict__ dbout, long long B,
                                                     int A, int C, int flip, float clamp, int rows_per_cta) {
  extern __shared__ float smem_f[];   // [8 warps][2][C] + [8][2]
  NF_PDL_ENTRY();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4 w0[NV], w1[NV], a0[NV], a1[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int c4 = lane + 32 * k;
    const bool ok = 4 * c4 < C;
    w0[k] = ok ? __ldg(reinterpret_cast<const float4*>(wout) + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
    w1[k] = ok ? __ldg(reinterpret_cast<const float4*>(wout + C) + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
    a0[k] = a1[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  float sb0 = 0.f, sb1 = 0.f;
  const long long r0 = (long long)blockIdx.x * rows_per_cta, r1 = min(B, r0 + rows_per_cta);
  for (long long b = r0 + warp; b < r1; b += 8) {
    const float gl = dld ? dld[b] : 0.f;
    for (int t = 0; t < A; ++t) {
      const int col = flip ? A - 1 - t : t;
      const float xa = xa_st[b * A + t];
      const float g = dz ? dz[b * A + col] : 0.f;
      const float e = expf(-xa);
      const float dxb = -g * e;
      float draw = -g * z[b * A + col] - gl;
      if (clamp > 0.f) { const float q = xa / clamp; draw *= 1.0f - q * q; }
      if (lane == 0) dx[b * A + col] = g * e;
      const float4* hr = reinterpret_cast<const float4*>(hf + (b * A + t) * C);
      float4* dr = reinterpret_cast<float4*>(dh + (b * A + t) * C);
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const int c4 = lane + 32 * k;
        if (4 * c4 < C) {
          const float4 hv = hr[c4];
          a0[k].x = fmaf(draw, hv.x, a0[k].x); a0[k].y = fmaf(draw, hv.y, a0[k].y);
          a0[k].z = fmaf(draw, hv.z, a0[k].z); a0[k].w = fmaf(draw, hv.w, a0[k].w);
          a1[k].x = fmaf(dxb, hv.x, a1[k].x); a1[k].y = fmaf(dxb, hv.y, a1[k].y);
          a1[k].z = fmaf(dxb, hv.z, a1[k].z); a1[k].w = fmaf(dxb, hv.w, a1[k].w);
          dr[c4] = make_float4(fmaf(draw, w0[k].x, dxb * w1[k].x), fmaf(draw, w0[k].y, dxb * w1[k].y),
                               fmaf(draw, w0[k].z, dxb * w1[k].z), fmaf(draw, w0[k].w, dxb * w1[k].w));
        }
      }
      sb0 += draw; sb1 += dxb;
    }
  }
  float* mine = smem_f + (size_t)warp * 2 * C;
  float* sbias = smem_f + (size_t)8 * 2 * C;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int c = 4 * (lane + 32 * k);
    if (c < C) {
      *reinterpret_cast<float4*>(mine + c) = a0[k];
      *reinterpret_cast<float4*>(mine + C + c) = a1[k];
    }
  }
  if (lane == 0) { sbias[2 * warp] = sb0; sbias[2 * warp + 1] = sb1; }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) v += smem_f[(size_t)w * 2 * C + i];
    atomicAdd(dwout + i, v);
  }
  if (threadIdx.x < 2) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) v += sbias[2 * w + threadIdx.x];
    atomicAdd(dbout + threadIdx.x, v);
  }
}

// Backward of the embedding: dx[b, perm(t - 1)] += dh0[b, t] . w_in;  dw_in += sum_{b, t >= 1} dh0[b, t] x[b, perm(t - 1)]
// (d pos_embed and d b_in are plain column sums: k_tf_colsum / k_tf_bin_from_pos)
template <int NV>
__global__ void __launch_bounds__(256) k_tf_embed_bwd(const float* __restrict__ dh0, const float* __restrict__ x,
                                                      const float* __restrict__ win, float* __restrict__ dx,
                                                      float* __restrict__ dwin, long long B, int A, int C, int flip,
                                                      int rows_per_cta) {
  extern __shared__ float smem_f[];   // [8 warps][C]
  NF_PDL_ENTRY();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4 wv[NV], acc[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int c4 = lane + 32 * k;
    wv[k] = 4 * c4 < C ? __ldg(reinterpret_cast<const float4*>(win) + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
    acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const long long r0 = (long long)blockIdx.x * rows_per_cta, r1 = min(B, r0 + rows_per_cta);
  for (long long b = r0 + warp; b < r1; b += 8) {
    for (int t = 1; t < A; ++t) {
      const int col = flip ? A - t : t - 1;
      const float xv = x[b * A + col];
      const float4* dr = reinterpret_cast<const float4*>(dh0 + (b * A + t) * C);
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const int c4 = lane + 32 * k;
        if (4 * c4 < C) {
          const float4 d = dr[c4];
          s += dot4(d, wv[k]);
          acc[k].x = fmaf(d.x, xv, acc[k].x); acc[k].y = fmaf(d.y, xv, acc[k].y);
          acc[k].z = fmaf(d.z, xv, acc[k].z); acc[k].w = fmaf(d.w, xv, acc[k].w);
        }
      }
      s = warp_sum(s);
      if (lane == 0) dx[b * A + col] += s;
    }
  }
  float* mine = smem_f + (size_t)warp * C;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int c = 4 * (lane + 32 * k);
    if (c < C) *reinterpret_cast<float4*>(mine + c) = acc[k];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) v += smem_f[(size_t)w * C + i];
    atomicAdd(dwin + i, v);
  }
}

// out[c] += sum over the CTA's rows of X[row * ld + c]
__global__ void __launch_bounds__(128) k_tf_colsum(const float* __restrict__ X, long long ld, long long rows, int N,
                                                   float* __restrict__ out, int rows_per_cta) {
  NF_PDL_ENTRY();
  const int c = blockIdx.y * 128 + threadIdx.x;
  if (c >= N) return;
  const long long r0 = (long long)blockIdx.x * rows_per_cta, r1 = min(rows, r0 + rows_per_cta);
  float acc = 0.f;
  for (long long r = r0; r < r1; ++r) acc += X[r * ld + c];
  atomicAdd(&out[c], acc);
}

// d b_in[c] = sum_{t >= 1} d pos[t, c]   (both are sums of dh0 over the samples)
__global__ void __launch_bounds__(256) k_tf_bin_from_pos(const float* __restrict__ dpos, float* __restrict__ dbin, int T,
                                                         int C) {
  NF_PDL_ENTRY();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float acc = 0.f;
  for (int t = 1; t < T; ++t) acc += dpos[(long long)t * C + c];
  dbin[c] += acc;
}

// ---- launch helpers ------------------------------------------------------------------------------------------------
int nv_of(int C) { const int nv = (int)cdiv(C, 128); return nv <= 1 ? 1 : (nv <= 2 ? 2 : (nv <= 4 ? 4 : kTfMaxV)); }

#define TF_NV_SWITCH(C, CALL)                  \
  switch (nv_of(C)) {                          \
    case 1: { constexpr int NV = 1; CALL; } break; \
    case 2: { constexpr int NV = 2; CALL; } break; \
    case 4: { constexpr int NV = 4; CALL; } break; \
    default: { constexpr int NV = kTfMaxV; CALL; } break; \
  }

int rows_per_cta_for(long long rows) {
  long long rpc = cdiv(rows, 296);
  return (int)std::max<long long>(32, cdiv(rpc, 8) * 8);
}

int tf_embed(const float* x, const float* cproj, const float* win, const float* bin, const float* pos, float* h,
             long long B, const TfDims& m, int t0, int t1, int flip, cudaStream_t st) {
  ProfScope prof(st, PROF_TF_ROW, 4.0 * B * (t1 - t0) * m.C);
  const long long n = B * (t1 - t0) * (m.C / 4);
  if (n <= 0) return 0;
  NF_LAUNCH((k_tf_embed), (unsigned)std::min<long long>(cdiv(n, 256), 148 * 8), 256, 0, st, x, cproj, win, bin, pos, h, B,
            m.A, m.C, t0, t1, flip);
  NF_LAUNCHED();
  return 0;
}

int tf_add_ln(const float* h_in, const float* r, float* h_out, float* a, float* stats, const float* gamma,
              const float* beta, long long rows, const TfDims& m, cudaStream_t st) {
  ProfScope prof(st, PROF_TF_ROW, 4.0 * rows * m.C * (r ? 4 : 3));
  if (rows <= 0) return 0;
  const unsigned grid = (unsigned)cdiv(rows, 8);
  TF_NV_SWITCH(m.C, NF_LAUNCH((k_tf_add_ln<NV>), grid, 256, 0, st, h_in, r, h_out, a, reinterpret_cast<float2*>(stats),
                              gamma, beta, rows, m.C, m.eps));
  NF_LAUNCHED();
  return 0;
}

int tf_ln_bwd(const float* da, const float* h, const float* stats, const float* gamma, const float* dres, float* dh_out,
              float* dgamma, float* dbeta, long long rows, const TfDims& m, cudaStream_t st) {
  ProfScope prof(st, PROF_TF_ROW, 4.0 * rows * m.C * (dres ? 4 : 3));
  if (rows <= 0) return 0;
  const int rpc = rows_per_cta_for(rows);
  const unsigned grid = (unsigned)cdiv(rows, rpc);
  const size_t smem = (size_t)(1 + 8 * 2) * m.C * sizeof(float);
  {
    static PerDeviceOnce once;
    if (once.first()) {
      constexpr int kMax = (1 + 8 * 2) * 128 * kTfMaxV * (int)sizeof(float);
      NF_CUDA(cudaFuncSetAttribute(k_tf_ln_bwd<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMax));
      NF_CUDA(cudaFuncSetAttribute(k_tf_ln_bwd<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMax));
      NF_CUDA(cudaFuncSetAttribute(k_tf_ln_bwd<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMax));
      NF_CUDA(cudaFuncSetAttribute(k_tf_ln_bwd<kTfMaxV>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMax));
    }
  }
  TF_NV_SWITCH(m.C, NF_LAUNCH((k_tf_ln_bwd<NV>), grid, 256, smem, st, da, h, reinterpret_cast<const float2*>(stats), gamma,
                              dres, dh_out, dgamma, dbeta, rows, m.C, rpc));
  NF_LAUNCHED();
  return 0;
}

int tf_attn_fwd(const float* qkv, float* o, float* P, long long B, const TfDims& m, int t0, int t1, cudaStream_t st) {
  ProfScope prof(st, PROF_TF_ATTN, 4.0 * B * m.C * 0.5 * (t1 - t0) * (t0 + t1 + 1));
  if (B <= 0) return 0;
  const unsigned grid = (unsigned)cdiv(B * m.H, 4);
  const float scale = 1.0f / sqrtf((float)m.hd);
  switch (m.hd / 32) {
    case 1: NF_LAUNCH((k_tf_attn_fwd<1>), grid, 128, 0, st, qkv, o, P, B, m.T, m.C, m.H, t0, t1, scale); break;
    case 2: NF_LAUNCH((k_tf_attn_fwd<2>), grid, 128, 0, st, qkv, o, P, B, m.T, m.C, m.H, t0, t1, scale); break;
    default: NF_LAUNCH((k_tf_attn_fwd<4>), grid, 128, 0, st, qkv, o, P, B, m.T, m.C, m.H, t0, t1, scale); break;
  }
  NF_LAUNCHED();
  return 0;
}

template <int HDV>
int tf_attn_bwd_t(const float* qkv, const float* o, const float* P, const float* d_o, float* dqkv, long long B,
                  const TfDims& m, cudaStream_t st) {
  const unsigned grid = (unsigned)cdiv(B * m.H, 4);
  const float scale = 1.0f / sqrtf((float)m.hd);
  if (m.T <= 8) NF_LAUNCH((k_tf_attn_bwd<HDV, 8>), grid, 128, 0, st, qkv, o, P, d_o, dqkv, B, m.T, m.C, m.H, scale);
  else if (m.T <= 16) NF_LAUNCH((k_tf_attn_bwd<HDV, 16>), grid, 128, 0, st, qkv, o, P, d_o, dqkv, B, m.T, m.C, m.H, scale);
  else NF_LAUNCH((k_tf_attn_bwd<HDV, 32>), grid, 128, 0, st, qkv, o, P, d_o, dqkv, B, m.T, m.C, m.H, scale);
  NF_LAUNCHED();
  return 0;
}
int tf_attn_bwd(const float* qkv, const float* o, const float* P, const float* d_o, float* dqkv, long long B,
                const TfDims& m, cudaStream_t st) {
  ProfScope prof(st, PROF_TF_ATTN, 8.0 * B * m.C * 0.5 * m.T * (m.T + 1));
  if (B <= 0) return 0;
  switch (m.hd / 32) {
    case 1: return tf_attn_bwd_t<1>(qkv, o, P, d_o, dqkv, B, m, st);
    case 2: return tf_attn_bwd_t<2>(qkv, o, P, d_o, dqkv, B, m, st);
    default: return tf_attn_bwd_t<4>(qkv, o, P, d_o, dqkv, B, m, st);
  }
}

int tf_gelu(const float* u, float* g, long long n, cudaStream_t st) {
  ProfScope prof(st, PROF_TF_ROW, 8.0 * n);
  if (n <= 0) return 0;
  NF_LAUNCH((k_tf_gelu), (unsigned)std::min<long long>(cdiv(n / 4, 256), 148 * 16), 256, 0, st,
            reinterpret_cast<const float4*>(u), reinterpret_cast<float4*>(g), n / 4);
  NF_LAUNCHED();
  return 0;
}
int tf_gelu_bwd(float* dg, const float* u, long long n, cudaStream_t st) {
  ProfScope prof(st, PROF_TF_ROW, 12.0 * n);
  if (n <= 0) return 0;
  NF_LAUNCH((k_tf_gelu_bwd), (unsigned)std::min<long long>(cdiv(n / 4, 256), 148 * 16), 256, 0, st,
            reinterpret_cast<float4*>(dg), reinterpret_cast<const float4*>(u), n / 4);
  NF_LAUNCHED();
  return 0;
}

struct TailArgs {
  const float* h; const float* r; const float* wout; const float* bout; const float* in; float* out; float* logdet;
  float* xa_out; int t0, t1, flip, inverse, ld_first;
  const float* win = nullptr; const float* bin = nullptr; const float* pos = nullptr; float* hnext = nullptr;
};
int tf_tail(const TailArgs& a, long long B, const TfDims& m, cudaStream_t st) {
  ProfScope prof(st, PROF_TF_ROW, 4.0 * B * (a.t1 - a.t0) * m.C * (a.r ? 2 : 1));
  if (B <= 0) return 0;
  const unsigned grid = (unsigned)cdiv(B, 8);
  TF_NV_SWITCH(m.C, NF_LAUNCH((k_tf_tail<NV>), grid, 256, 0, st, a.h, a.r, a.wout, a.bout, a.in, a.out, a.logdet, a.xa_out, B,
                              m.A, m.C, a.t0, a.t1, a.flip, m.clamp, a.inverse, a.ld_first, a.win, a.bin, a.pos, a.hnext));
  NF_LAUNCHED();
  return 0;
}

int tf_tail_bwd(const float* dz, const float* dld, const float* z, const float* xa, const float* hf, const float* wout,
                float* dh, float* dx, float* dwout, float* dbout, long long B, const TfDims& m, int flip, cudaStream_t st) {
  ProfScope prof(st, PROF_TF_ROW, 8.0 * B * m.A * m.C);
  if (B <= 0) return 0;
  const int rpc = rows_per_cta_for(B);
  const size_t smem = ((size_t)8 * 2 * m.C + 16) * sizeof(float);
  {
    static PerDeviceOnce once;
    if (once.first()) {
      constexpr int kMax = (8 * 2 * 128 * kTfMaxV + 16) * (int)sizeof(float);
      NF_CUDA(cudaFuncSetAttribute(k_tf_tail_bwd<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMax));
      NF_CUDA(cudaFuncSetAttribute(k_tf_tail_bwd<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMax));
      NF_CUDA(cudaFuncSetAttribute(k_tf_tail_bwd<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMax));
      NF_CUDA(cudaFuncSetAttribute(k_tf_tail_bwd<kTfMaxV>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMax));
    }
  }
  TF_NV_SWITCH(m.C, NF_LAUNCH((k_tf_tail_bwd<NV>), (unsigned)cdiv(B, rpc), 256, smem, st, dz, dld, z, xa, hf, wout, dh, dx,
                              dwout, dbout, B, m.A, m.C, flip, m.clamp, rpc));
  NF_LAUNCHED();
  return 0;
}

int tf_embed_bwd(const float* dh0, const float* x, const float* win, float* dx, float* dwin, long long B, const TfDims& m,
                 int flip, cudaStream_t st) {
  ProfScope prof(st, PROF_TF_ROW, 4.0 * B * m.A * m.C);
  if (B <= 0 || m.A < 2) return 0;
  const int rpc = rows_per_cta_for(B);
  const size_t smem = (size_t)8 * m.C * sizeof(float);
  TF_NV_SWITCH(m.C, NF_LAUNCH((k_tf_embed_bwd<NV>), (unsigned)cdiv(B, rpc), 256, smem, st, dh0, x, win, dx, dwin, B, m.A, m.C,
                              flip, rpc));
  NF_LAUNCHED();
  return 0;
}

int tf_colsum(const float* X, long long ld, long long rows, int N, float* out, cudaStream_t st) {
  ProfScope prof(st, PROF_TF_ROW, 4.0 * rows * N);
  if (rows <= 0 || N <= 0) return 0;
  const int rpc = rows_per_cta_for(rows);
  NF_LAUNCH((k_tf_colsum), dim3((unsigned)cdiv(rows, rpc), (unsigned)cdiv(N, 128)), 128, 0, st, X, ld, rows, N, out, rpc);
  NF_LAUNCHED();
  return 0;
}

// C (rows, N) = A (rows, K; pitch lda) @ W (N, K)^T + bias     (pitch ldc)
int tf_linear(int prec, const float* A, long long lda, const float* W, int K, const float* bias, float* C, long long ldc,
              int N, long long rows, int accumulate, cudaStream_t st) {
  TcGemmP t;
  t.A0 = A; t.lda0 = lda; t.K0 = K; t.B0 = W; t.ldb0 = K; t.C = C; t.ldc = ldc; t.bias = bias;
  t.M = (int)rows; t.N = N; t.accumulate = accumulate;
  return linear_nt(prec, t, st);
}

// dW (n_out, n_in) += dY (rows, n_out)^T X (rows, n_in);  db (n_out) += colsum(dY) -- on the GEMM's A operand when it can
int tf_wgrad(int prec, const float* dY, long long lddy, const float* X, long long ldx, float* dW, float* db, int n_out,
             int n_in, long long rows, cudaStream_t st) {
  TcGemmTnP w;
  w.A = dY; w.lda = lddy; w.B = X; w.ldb = ldx; w.C = dW; w.ldc = n_in; w.M = n_out; w.N = n_in; w.K = rows;
  if (db && rides_ok(prec, w)) {
    w.colsum = db;
    return wgrad_tn(prec, w, st);
  }
  NF_TRY(wgrad_tn(prec, w, st));
  return db ? tf_colsum(dY, lddy, rows, n_out, db, st) : 0;
}

GraphKey tf_key(int op, const NfTarflowDesc* d, int64_t B, int64_t ws_bytes) {
  GraphKey k;
  k.op = op; k.B = B; k.ws_bytes = ws_bytes;
  k.desc.D = d->A; k.desc.C = d->C; k.desc.H1 = d->n_layers * 4096 + d->head_dim * 8 + d->expansion; k.desc.R = d->R;
  k.desc.n_blocks = d->n_blocks; k.desc.ln_eps = d->ln_eps; k.desc.precision = d->precision;
  memcpy(&k.desc.reserved, &d->logscale_clamp, sizeof(float));
  return k;
}

// One pass of tokens [t0, t1) of every sample through the layers of block `blk` (rows = B (t1 - t0) dense rows, except
// the qkv rows, which always live at their (b, t) position of a (B, T, 3C) buffer so attention can see earlier tokens).
// Training (sb != NULL, all tokens): every intermediate goes to the stash.  Returns with the last residual pending:
// the block output is *h_last + *r_last.
struct PassBufs {
  float* h;      // in: embedded tokens (rows, C); updated in place when there is no stash
  float* a;      // LayerNorm output
  float* o;      // attention output
  float* tmp;    // residual branch output
  float* big;    // MLP hidden (rows, EC)
  float* qkv0;   // (B, T, 3C) per layer, stride qkv_layer
  int64_t qkv_layer;
};
int tf_layers(const TfDims& m, const TfLayout& pl, const float* bp, long long B, int t0, int t1, const PassBufs& bf,
              float* sb, const TfStash& sl, const float** h_last, const float** r_last, cudaStream_t st) {
  const int nt = t1 - t0;
  const long long rows = B * nt;
  const float* h = bf.h;
  const float* r = nullptr;
  for (int l = 0; l < m.L; ++l) {
    const float* lp = bp + pl.layers + (int64_t)l * pl.layer_stride;
    float* ls = sb ? sb + (int64_t)l * sl.layer_stride : nullptr;
    float* h_in = ls ? ls + sl.h_in : bf.h;
    float* a = ls ? ls + sl.a : bf.a;
    float* qkv = ls ? ls + sl.qkv : bf.qkv0 + (int64_t)l * bf.qkv_layer;
    float* o = ls ? ls + sl.o : bf.o;
    float* h2 = ls ? ls + sl.h2 : bf.h;
    float* mm = ls ? ls + sl.mm : bf.a;
    float* u = ls ? ls + sl.u : bf.big;
    float* g = ls ? ls + sl.g : bf.big;
    // x = x + attn(x)
    NF_TRY(tf_add_ln(h, r, h_in, a, ls ? ls + sl.st1 : nullptr, lp + pl.ln1w, lp + pl.ln1b, rows, m, st));
    NF_TRY(tf_linear(m.prec, a, m.C, lp + pl.wqkv, m.C, lp + pl.bqkv, qkv + (int64_t)t0 * 3 * m.C,
                     (nt == m.T ? 3 : 3 * m.T) * (long long)m.C, 3 * m.C, rows, 0, st));
    NF_TRY(tf_attn_fwd(qkv, o, ls ? ls + sl.p : nullptr, B, m, t0, t1, st));
    NF_TRY(tf_linear(m.prec, o, m.C, lp + pl.wproj, m.C, lp + pl.bproj, bf.tmp, m.C, m.C, rows, 0, st));
    // x = x + mlp(x)
    NF_TRY(tf_add_ln(h_in, bf.tmp, h2, mm, ls ? ls + sl.st2 : nullptr, lp + pl.ln2w, lp + pl.ln2b, rows, m, st));
    NF_TRY(tf_linear(m.prec, mm, m.C, lp + pl.w1, m.C, lp + pl.b1, u, m.EC, m.EC, rows, 0, st));
    NF_TRY(tf_gelu(u, g, rows * m.EC, st));
    NF_TRY(tf_linear(m.prec, g, m.EC, lp + pl.w2, m.EC, lp + pl.b2, bf.tmp, m.C, m.C, rows, 0, st));
    h = h2; r = bf.tmp;
  }
  *h_last = h; *r_last = r;
  return 0;
}

}  // namespace
}  // namespace nf

using namespace nf;

extern "C" {

NF_API int64_t nf_tarflow_param_layout(const NfTarflowDesc* desc, int64_t* offs) {
  if (check_tf(desc) != 0) return -1;
  const TfLayout l = tf_layout(tf_dims(desc));
  if (offs) {
    const int64_t v[NF_TARFLOW_LAYOUT_SLOTS] = {l.block_stride, l.pos, l.win, l.bin, l.wc, l.bc, l.layers, l.layer_stride,
                                                l.ln1w, l.ln1b, l.wqkv, l.bqkv, l.wproj, l.bproj, l.ln2w, l.ln2b,
                                                l.w1, l.b1, l.w2, l.b2, l.wout, l.bout};
    for (int i = 0; i < NF_TARFLOW_LAYOUT_SLOTS; ++i) offs[i] = v[i];
  }
  return l.total;
}

NF_API int64_t nf_tarflow_stash_bytes(const NfTarflowDesc* desc, int64_t B) {
  if (check_tf(desc) != 0 || B < 0) return -1;
  return std::max<int64_t>(16, tf_stash(tf_dims(desc), B).total * 4);
}

NF_API int64_t nf_tarflow_workspace_bytes(const NfTarflowDesc* desc, int64_t B) {
  if (check_tf(desc) != 0 || B < 0) return -1;
  return std::max<int64_t>(16, tf_ws(tf_dims(desc), B).total * 4);
}

NF_API int nf_tarflow_forward(const NfTarflowDesc* desc, const float* params, const float* x, const float* y, int64_t B,
                              float* z, float* logdet, float* outputs, void* stash, void* workspace, int64_t ws_bytes,
                              void* stream) {
  NF_TRY(check_tf(desc));
  const TfDims m = tf_dims(desc);
  NF_CHECK_ARG(B >= 0 && B * m.T < (int64_t(1) << 30), NF_E_BADARG, "B=%lld out of range", (long long)B);
  if (B == 0) return 0;
  const TfLayout pl = tf_layout(m);
  const TfStash sl = tf_stash(m, B);
  const TfWs wl = tf_ws(m, B);
  NF_TRY(check_ptr(params, "params", true)); NF_TRY(check_ptr(x, "x", true)); NF_TRY(check_ptr(y, "y", true));
  NF_TRY(check_ptr(z, "z", true)); NF_TRY(check_ptr(logdet, "logdet", false)); NF_TRY(check_ptr(outputs, "outputs", false));
  NF_TRY(check_ptr(stash, "stash", false)); NF_TRY(check_ptr(workspace, "workspace", true));
  NF_CHECK_ARG(ws_bytes >= wl.total * 4, NF_E_WORKSPACE, "workspace too small: %lld < %lld bytes", (long long)ws_bytes,
               (long long)wl.total * 4);
  cudaStream_t user_st = static_cast<cudaStream_t>(stream);
  float* ws = static_cast<float*>(workspace);
  float* sbase = static_cast<float*>(stash);
  const int64_t BA = B * m.A;
  float* x_st = sbase ? sbase + sl.xs : ws + wl.io_x;
  float* y_st = sbase ? sbase + sl.y : ws + wl.io_y;
  {
    CopySegs c;
    c.add(x_st, x, BA); c.add(y_st, y, B * m.R);
    NF_TRY(copy_segments(c, user_st));
  }
  GraphKey key = tf_key(30, desc, B, ws_bytes);
  key.ptr[0] = params; key.ptr[1] = stash; key.ptr[2] = workspace; key.ptr[3] = outputs;
  key.flags = (logdet ? 1 : 0);
  auto body = [&](cudaStream_t st) -> int {
    const float* cur = x_st;
    for (int i = 0; i < m.nb; ++i) {
      const float* bp = params + (int64_t)i * pl.block_stride;
      float* sb = sbase ? sbase + sl.blocks + (int64_t)i * sl.block_stride : nullptr;
      const int flip = i & 1;
      float* nxt = sbase ? sbase + sl.xs + (int64_t)(i + 1) * BA
                         : (outputs ? outputs + (int64_t)i * BA : (i == m.nb - 1 ? ws + wl.io_z : ws + (i & 1 ? wl.x1 : wl.x0)));
      NF_TRY(tf_linear(m.prec, y_st, m.R, bp + pl.wc, m.R, bp + pl.bc, ws + wl.cproj, m.C, m.C, B, 0, st));
      PassBufs bf{};
      bf.h = ws + wl.bufA; bf.a = ws + wl.bufB; bf.o = ws + wl.bufO; bf.tmp = ws + wl.tmp; bf.big = ws + wl.big;
      bf.qkv0 = ws + wl.qkv; bf.qkv_layer = 0;
      NF_TRY(tf_embed(cur, ws + wl.cproj, bp + pl.win, bp + pl.bin, bp + pl.pos, bf.h, B, m, 0, m.T, flip, st));
      const float *hl, *rl;
      NF_TRY(tf_layers(m, pl, bp, B, 0, m.T, bf, sb, sl, &hl, &rl, st));
      if (sb) {   // backward needs the block's final activations (d proj_out.weight)
        NF_TRY(tf_add_ln(hl, rl, sb + sl.hf, nullptr, nullptr, nullptr, nullptr, B * m.T, m, st));
        hl = sb + sl.hf; rl = nullptr;
      }
      TailArgs t{};
      t.h = hl; t.r = rl; t.wout = bp + pl.wout; t.bout = bp + pl.bout; t.in = cur; t.out = nxt;
      t.logdet = logdet ? ws + wl.io_ld : nullptr; t.xa_out = sbase ? sbase + sl.xa + (int64_t)i * BA : nullptr;
      t.t0 = 0; t.t1 = m.A; t.flip = flip; t.inverse = 0; t.ld_first = i == 0;
      NF_TRY(tf_tail(t, B, m, st));
      cur = nxt;
    }
    return 0;
  };
  NF_TRY(run_graphed(key, user_st, body));
  CopySegs c;
  const float* last = sbase ? sbase + sl.xs + (int64_t)m.nb * BA : (outputs ? outputs + (int64_t)(m.nb - 1) * BA : ws + wl.io_z);
  c.add(z, last, BA);
  c.add(logdet, ws + wl.io_ld, B);
  if (sbase && outputs) c.add(outputs, sbase + sl.xs + BA, (int64_t)m.nb * BA);
  return copy_segments(c, user_st);
}

NF_API int nf_tarflow_inverse(const NfTarflowDesc* desc, const float* params, const float* z, const float* y, int64_t B,
                              float* x, float* logdet, void* workspace, int64_t ws_bytes, void* stream) {
  NF_TRY(check_tf(desc));
  const TfDims m = tf_dims(desc);
  NF_CHECK_ARG(B >= 0 && B * m.T < (int64_t(1) << 30), NF_E_BADARG, "B=%lld out of range", (long long)B);
  if (B == 0) return 0;
  const TfLayout pl = tf_layout(m);
  const TfStash sl = tf_stash(m, B);
  const TfWs wl = tf_ws(m, B);
  NF_TRY(check_ptr(params, "params", true)); NF_TRY(check_ptr(z, "z", true)); NF_TRY(check_ptr(y, "y", true));
  NF_TRY(check_ptr(x, "x", true)); NF_TRY(check_ptr(logdet, "logdet", false)); NF_TRY(check_ptr(workspace, "workspace", true));
  NF_CHECK_ARG(ws_bytes >= wl.total * 4, NF_E_WORKSPACE, "workspace too small: %lld < %lld bytes", (long long)ws_bytes,
               (long long)wl.total * 4);
  cudaStream_t user_st = static_cast<cudaStream_t>(stream);
  float* ws = static_cast<float*>(workspace);
  const int64_t BA = B * m.A;
  {
    CopySegs c;
    c.add(ws + wl.io_z, z, BA); c.add(ws + wl.io_y, y, B * m.R);
    NF_TRY(copy_segments(c, user_st));
  }
  GraphKey key = tf_key(31, desc, B, ws_bytes);
  key.ptr[0] = params; key.ptr[2] = workspace;
  key.flags = (logdet ? 1 : 0);
  // evaluation-sized batches: one persistent cooperative kernel per block runs the whole autoregressive loop
  // (tarflow_ar.cu; option "tf_ar": -1 auto = whenever the shape is supported, 0 never)
  const ArShape shape{(int)std::min<int64_t>(B, 1 << 20), m.A, m.C, m.EC, m.L, m.hd};
  const bool persistent = option(NF_OPT_TF_AR) != 0 && m.nb <= 64 && tf_ar_plan(shape, nullptr, nullptr);
  key.flags |= persistent ? 2 : 0;
  auto body = [&](cudaStream_t st) -> int {
    const float* cur = ws + wl.io_z;
    if (persistent) NF_TRY(fill_zero(ws + wl.bars, 64, st));
    for (int i = m.nb - 1; i >= 0; --i) {
      const float* bp = params + (int64_t)i * pl.block_stride;
      const int flip = i & 1;
      float* nxt = i == 0 ? ws + wl.io_x : ws + ((i & 1) ? wl.x1 : wl.x0);
      NF_TRY(tf_linear(m.prec, ws + wl.io_y, m.R, bp + pl.wc, m.R, bp + pl.bc, ws + wl.cproj, m.C, m.C, B, 0, st));
      if (persistent) {
        ArP a{};
        a.pos = bp + pl.pos; a.win = bp + pl.win; a.bin = bp + pl.bin; a.wout = bp + pl.wout; a.bout = bp + pl.bout;
        a.layers = bp + pl.layers; a.layer_stride = pl.layer_stride;
        a.o_ln1w = pl.ln1w; a.o_ln1b = pl.ln1b; a.o_wqkv = pl.wqkv; a.o_bqkv = pl.bqkv; a.o_wproj = pl.wproj; a.o_bproj = pl.bproj;
        a.o_ln2w = pl.ln2w; a.o_ln2b = pl.ln2b; a.o_w1 = pl.w1; a.o_b1 = pl.b1; a.o_w2 = pl.w2; a.o_b2 = pl.b2;
        a.A = m.A; a.C = m.C; a.H = m.H; a.hd = m.hd; a.L = m.L; a.EC = m.EC; a.B = (int)B; a.flip = flip;
        a.ld_first = i == m.nb - 1; a.eps = m.eps; a.clamp = m.clamp;
        a.zin = cur; a.xout = nxt; a.logdet = logdet ? ws + wl.io_ld : nullptr; a.cproj = ws + wl.cproj;
        a.qkv = ws + wl.kv; a.o = ws + wl.bufO; a.h = ws + wl.bufA; a.h2 = ws + wl.bufB; a.g = ws + wl.big;
        a.bar = reinterpret_cast<unsigned*>(ws + wl.bars) + i;
        NF_TRY(tf_ar_block(a, st));
        cur = nxt;
        continue;
      }
      PassBufs bf{};
      bf.h = ws + wl.bufA; bf.a = ws + wl.bufB; bf.o = ws + wl.bufO; bf.tmp = ws + wl.tmp; bf.big = ws + wl.big;
      bf.qkv0 = ws + wl.kv; bf.qkv_layer = 3 * B * m.T * m.C;
      // token 0 = the conditioning token
      NF_TRY(tf_embed(nullptr, ws + wl.cproj, bp + pl.win, bp + pl.bin, bp + pl.pos, bf.h, B, m, 0, 1, flip, st));
      for (int t = 0; t < m.A; ++t) {     // the autoregressive loop: dimension t from the tokens 0 .. t
        const float *hl, *rl;
        NF_TRY(tf_layers(m, pl, bp, B, t, t + 1, bf, nullptr, sl, &hl, &rl, st));
        TailArgs ta{};
        ta.h = hl; ta.r = rl; ta.wout = bp + pl.wout; ta.bout = bp + pl.bout; ta.in = cur; ta.out = nxt;
        ta.logdet = logdet ? ws + wl.io_ld : nullptr; ta.xa_out = nullptr;
        ta.t0 = t; ta.t1 = t + 1; ta.flip = flip; ta.inverse = 1; ta.ld_first = (i == m.nb - 1 && t == 0);
        ta.win = bp + pl.win; ta.bin = bp + pl.bin; ta.pos = bp + pl.pos; ta.hnext = bf.h;
        NF_TRY(tf_tail(ta, B, m, st));
      }
      cur = nxt;
    }
    return 0;
  };
  NF_TRY(run_graphed(key, user_st, body));
  CopySegs c;
  c.add(x, ws + wl.io_x, BA);
  c.add(logdet, ws + wl.io_ld, B);
  return copy_segments(c, user_st);
}

NF_API int nf_tarflow_backward(const NfTarflowDesc* desc, const float* params, const void* stash, int64_t B,
                               const float* dz, const float* dlogdet, float* dx, float* dy, float* dparams,
                               void* workspace, int64_t ws_bytes, void* stream) {
  NF_TRY(check_tf(desc));
  const TfDims m = tf_dims(desc);
  NF_CHECK_ARG(B >= 0 && B * m.T < (int64_t(1) << 30), NF_E_BADARG, "B=%lld out of range", (long long)B);
  const TfLayout pl = tf_layout(m);
  const TfStash sl = tf_stash(m, B);
  const TfWs wl = tf_ws(m, B);
  NF_TRY(check_ptr(params, "params", true));
  cudaStream_t user_st = static_cast<cudaStream_t>(stream);
  if (B == 0) return dparams ? fill_zero(dparams, pl.total, user_st) : 0;
  NF_TRY(check_ptr(stash, "stash", true)); NF_TRY(check_ptr(dz, "dz", false)); NF_TRY(check_ptr(dlogdet, "dlogdet", false));
  NF_TRY(check_ptr(dx, "dx", false)); NF_TRY(check_ptr(dy, "dy", false)); NF_TRY(check_ptr(dparams, "dparams", false));
  NF_TRY(check_ptr(workspace, "workspace", true));
  NF_CHECK_ARG(ws_bytes >= wl.total * 4, NF_E_WORKSPACE, "workspace too small: %lld < %lld bytes", (long long)ws_bytes,
               (long long)wl.total * 4);
  float* ws = static_cast<float*>(workspace);
  const float* sbase = static_cast<const float*>(stash);
  const int64_t BA = B * m.A, M = B * m.T;
  {
    CopySegs c;
    c.add(ws + wl.io_dz, dz, BA); c.add(ws + wl.io_dld, dlogdet, B);
    NF_TRY(copy_segments(c, user_st));
  }
  GraphKey key = tf_key(32, desc, B, ws_bytes);
  key.ptr[0] = params; key.ptr[1] = stash; key.ptr[2] = workspace; key.ptr[3] = dparams;
  key.flags = (dz ? 1 : 0) | (dlogdet ? 2 : 0) | (dx ? 4 : 0) | (dy ? 8 : 0);
  auto body = [&](cudaStream_t st) -> int {
    if (dparams) NF_TRY(fill_zero(dparams, pl.total, st));
    if (dy) NF_TRY(fill_zero(ws + wl.io_dy, B * m.R, st));
    const float* y_st = sbase + sl.y;
    const float* gz = dz ? ws + wl.io_dz : nullptr;
    const float* gld = dlogdet ? ws + wl.io_dld : nullptr;
    float* scratch_dp = ws + wl.big2;     // gradient sink when the caller wants no parameter gradients
    for (int i = m.nb - 1; i >= 0; --i) {
      const float* bp = params + (int64_t)i * pl.block_stride;
      float* dp = dparams ? dparams + (int64_t)i * pl.block_stride : nullptr;
      const float* sb = sbase + sl.blocks + (int64_t)i * sl.block_stride;
      float* wt = ws + wl.wT + (int64_t)i * wl.t_block_stride;
      const int flip = i & 1;
      const float* xin = sbase + sl.xs + (int64_t)i * BA;
      const float* zi = sbase + sl.xs + (int64_t)(i + 1) * BA;
      float* dxi = i == 0 ? ws + wl.io_dx : ws + ((i & 1) ? wl.x1 : wl.x0);
      float* dh = ws + wl.bufA;         // gradient of the residual stream
      float* dh2 = ws + wl.bufB;
      // tail: d proj_out, dh_final, the direct part of dx
      NF_TRY(tf_tail_bwd(gz, gld, zi, sbase + sl.xa + (int64_t)i * BA, sb + sl.hf, bp + pl.wout, dh, dxi,
                         dp ? dp + pl.wout : scratch_dp, dp ? dp + pl.bout : scratch_dp + 2 * m.C, B, m, flip, st));
      for (int l = m.L - 1; l >= 0; --l) {
        const float* lp = bp + pl.layers + (int64_t)l * pl.layer_stride;
        float* dl = dp ? dp + pl.layers + (int64_t)l * pl.layer_stride : nullptr;
        const float* ls = sb + (int64_t)l * sl.layer_stride;
        float* wtl = wt + (int64_t)l * wl.t_layer_stride;
        float* dbig = ws + wl.big;
        // h_out = h2 + fc2(g):  d fc2, dg = dh W2, du = dg GELU'(u), d fc1, dm = du W1, dh2 = dh + LN2'(dm)
        if (dl) NF_TRY(tf_wgrad(m.prec, dh, m.C, ls + sl.g, m.EC, dl + pl.w2, dl + pl.b2, m.C, m.EC, M, st));
        NF_TRY(transpose2d(wtl + wl.t_w2, m.C, 0, 0, lp + pl.w2, m.EC, 0, 0, m.C, m.EC, 1, 1, st));      // (C, EC) -> (EC, C)
        NF_TRY(tf_linear(m.prec, dh, m.C, wtl + wl.t_w2, m.C, nullptr, dbig, m.EC, m.EC, M, 0, st));
        NF_TRY(tf_gelu_bwd(dbig, ls + sl.u, M * m.EC, st));
        if (dl) NF_TRY(tf_wgrad(m.prec, dbig, m.EC, ls + sl.mm, m.C, dl + pl.w1, dl + pl.b1, m.EC, m.C, M, st));
        NF_TRY(transpose2d(wtl + wl.t_w1, m.EC, 0, 0, lp + pl.w1, m.C, 0, 0, m.EC, m.C, 1, 1, st));      // (EC, C) -> (C, EC)
        NF_TRY(tf_linear(m.prec, dbig, m.EC, wtl + wl.t_w1, m.EC, nullptr, ws + wl.tmp, m.C, m.C, M, 0, st));
        NF_TRY(tf_ln_bwd(ws + wl.tmp, ls + sl.h2, ls + sl.st2, lp + pl.ln2w, dh, dh2, dl ? dl + pl.ln2w : scratch_dp,
                         dl ? dl + pl.ln2b : scratch_dp + m.C, M, m, st));
        // h2 = h_in + proj(o):  d proj, do = dh2 Wp, attention backward, d qkv, da = dqkv Wqkv, dh_in = dh2 + LN1'(da)
        if (dl) NF_TRY(tf_wgrad(m.prec, dh2, m.C, ls + sl.o, m.C, dl + pl.wproj, dl + pl.bproj, m.C, m.C, M, st));
        NF_TRY(transpose2d(wtl + wl.t_proj, m.C, 0, 0, lp + pl.wproj, m.C, 0, 0, m.C, m.C, 1, 1, st));
        NF_TRY(tf_linear(m.prec, dh2, m.C, wtl + wl.t_proj, m.C, nullptr, ws + wl.bufO, m.C, m.C, M, 0, st));
        float* dqkv = ws + wl.qkv;
        NF_TRY(tf_attn_bwd(ls + sl.qkv, ls + sl.o, ls + sl.p, ws + wl.bufO, dqkv, B, m, st));
        if (dl) NF_TRY(tf_wgrad(m.prec, dqkv, 3 * m.C, ls + sl.a, m.C, dl + pl.wqkv, dl + pl.bqkv, 3 * m.C, m.C, M, st));
        NF_TRY(transpose2d(wtl + wl.t_qkv, 3 * m.C, 0, 0, lp + pl.wqkv, m.C, 0, 0, 3 * m.C, m.C, 1, 1, st));   // (3C, C) -> (C, 3C)
        NF_TRY(tf_linear(m.prec, dqkv, 3 * m.C, wtl + wl.t_qkv, 3 * m.C, nullptr, ws + wl.tmp, m.C, m.C, M, 0, st));
        NF_TRY(tf_ln_bwd(ws + wl.tmp, ls + sl.h_in, ls + sl.st1, lp + pl.ln1w, dh2, dh, dl ? dl + pl.ln1w : scratch_dp,
                         dl ? dl + pl.ln1b : scratch_dp + m.C, M, m, st));
      }
      // embedding: dh = d h0 (B, T, C)
      if (dp) {
        NF_TRY(tf_colsum(dh, (long long)m.T * m.C, B, m.T * m.C, dp + pl.pos, st));
        NF_LAUNCH((k_tf_bin_from_pos), (unsigned)cdiv(m.C, 256), 256, 0, st, dp + pl.pos, dp + pl.bin, m.T, m.C);
        NF_LAUNCHED();
      }
      if (i > 0 || dx || dp)
        NF_TRY(tf_embed_bwd(dh, xin, bp + pl.win, dxi, dp ? dp + pl.win : scratch_dp, B, m, flip, st));
      // conditioning token: d cond_proj, dy += dh0[:, 0, :] Wc
      if (dp) NF_TRY(tf_wgrad(m.prec, dh, (long long)m.T * m.C, y_st, m.R, dp + pl.wc, dp + pl.bc, m.C, m.R, B, st));
      if (dy) {
        NF_TRY(transpose2d(wt + wl.t_wc, m.C, 0, 0, bp + pl.wc, m.R, 0, 0, m.C, m.R, 1, 1, st));          // (C, R) -> (R, C)
        NF_TRY(tf_linear(m.prec, dh, (long long)m.T * m.C, wt + wl.t_wc, m.C, nullptr, ws + wl.io_dy, m.R, m.R, B, 1, st));
      }
      gz = dxi; gld = dlogdet ? ws + wl.io_dld : nullptr;
    }
    return 0;
  };
  NF_TRY(run_graphed(key, user_st, body));
  CopySegs c;
  c.add(dx, ws + wl.io_dx, BA);
  c.add(dy, ws + wl.io_dy, B * m.R);
  return copy_segments(c, user_st);
}

}  // extern "C"
