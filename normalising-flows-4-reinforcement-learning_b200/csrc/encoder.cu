// Conditioning encoder: n_hidden x [Linear -> LayerNorm(affine) -> SiLU] -> Linear, the
// module that produces the flow's y in every reference caller (GEncoder, torch_fcnf.py:154-190; flax twin
// nf.py:188-212 with eps 1e-6).  The Linears and their data / weight gradients go through the same dispatch as
// the flow's conditioner layers (tcgen05 3xTF32, FFMA for odd pitches); the row-wise part is two kernels here:
//   k_enc_ln_silu_fwd  x_hat = LN_noaffine(u), h = SiLU(gamma x_hat + beta)            (warp per row, 128-bit accesses)
//   k_enc_ln_silu_bwd  du from dh, plus the column sums d gamma, d beta, d bias        (registers -> smem -> atomics)
// SiLU follows the affine LayerNorm, so -- unlike the flow's LeakyReLU -> LN -> Linear -- gamma / beta cannot be
// folded into the next Linear.
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "elementwise.cuh"
#include "graphs.cuh"
#include "linear.cuh"

namespace nf {
namespace {

constexpr int kEncMaxV = 8;   // float4 per lane: hidden <= 1024

struct EncDims {
  int in, H, out, L, prec;
  float eps;
};

int check_enc(const NfEncoderDesc* d) {
  NF_CHECK_ARG(d != nullptr, NF_E_BADARG, "encoder desc is NULL");
  NF_CHECK_ARG(d->in_features >= 1 && d->out_features >= 1, NF_E_BADARG, "encoder: in_features=%d out_features=%d",
               d->in_features, d->out_features);
  NF_CHECK_ARG(d->hidden >= 4 && d->hidden % 4 == 0 && d->hidden <= 128 * kEncMaxV, NF_E_UNSUPPORTED,
               "encoder: hidden=%d must be a multiple of 4, at most %d", d->hidden, 128 * kEncMaxV);
  NF_CHECK_ARG(d->n_hidden >= 1 && d->n_hidden <= NF_ENC_MAX_HIDDEN_LAYERS, NF_E_UNSUPPORTED,
               "encoder: n_hidden=%d out of range", d->n_hidden);
  NF_CHECK_ARG(d->ln_eps > 0.f, NF_E_BADARG, "encoder: ln_eps must be positive");
  NF_CHECK_ARG(d->precision >= 0 && d->precision <= 1, NF_E_BADARG, "encoder: unknown precision %d", d->precision);
  return 0;
}

EncDims enc_dims(const NfEncoderDesc* d) {
  return EncDims{d->in_features, d->hidden, d->out_features, d->n_hidden, d->precision, d->ln_eps};
}

struct EncLayout {
  int64_t w[NF_ENC_MAX_HIDDEN_LAYERS], b[NF_ENC_MAX_HIDDEN_LAYERS], lw[NF_ENC_MAX_HIDDEN_LAYERS],
      lb[NF_ENC_MAX_HIDDEN_LAYERS];
  int64_t wo, bo, total;
};

EncLayout enc_layout(const EncDims& m) {
  EncLayout l{};
  int64_t off = 0;
  auto take = [&](int64_t n) { int64_t o = off; off += pad4(n); return o; };
  for (int i = 0; i < m.L; ++i) {
    l.w[i] = take((int64_t)m.H * (i == 0 ? m.in : m.H));
    l.b[i] = take(m.H);
    l.lw[i] = take(m.H);
    l.lb[i] = take(m.H);
  }
  l.wo = take((int64_t)m.out * m.H);
  l.bo = take(m.out);
  l.total = off;
  return l;
}

// stash: a copy of the input g (B, in) -- forward's staging buffer, read again by the first layer's weight gradient --
// then per hidden layer x_hat (B,H), h (B,H), rstd (B)
struct EncStash {
  int64_t g, layers, xh, h, rstd, layer_stride, total;
};
EncStash enc_stash(const EncDims& m, int64_t B) {
  EncStash s{};
  s.g = 0; s.layers = pad4(B * m.in);
  s.xh = 0; s.h = B * m.H; s.rstd = 2 * B * m.H;
  s.layer_stride = 2 * B * m.H + pad4(B);
  s.total = s.layers + s.layer_stride * m.L;
  return s;
}

// workspace: two (B,H) activation buffers, one transposed weight, and the I/O staging buffers that keep the cached
// launch graphs independent of the caller's input / output pointers (graphs.cuh)
struct EncWs {
  int64_t a, b, wt, io_g, io_out, io_dg, total;
};
EncWs enc_ws(const EncDims& m, int64_t B) {
  EncWs w{};
  int64_t off = 0;
  auto take = [&](int64_t n) { int64_t o = off; off += pad4(n); return o; };
  w.a = take(B * m.H); w.b = take(B * m.H);
  w.wt = take((int64_t)m.H * std::max(std::max(m.H, m.in), m.out));
  w.io_g = take(B * m.in); w.io_out = take(B * m.out); w.io_dg = take(B * m.in);
  w.total = off;
  return w;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// x_hat = (u - mean) / sqrt(var + eps) in place, h = SiLU(gamma x_hat + beta)     (torch_fcnf.py:172-174)
template <int NV>
__global__ void __launch_bounds__(256) k_enc_ln_silu_fwd(float* __restrict__ xh, float* __restrict__ h,
                                                         float* __restrict__ rstd, const float* __restrict__ gamma,
                                                         const float* __restrict__ beta, long long B, int H, float eps) {
  NF_PDL_ENTRY();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long row = (long long)blockIdx.x * 8 + warp;
  if (row >= B) return;
  float4* xr = reinterpret_cast<float4*>(xh + row * H);
  float4 v[NV];
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int c4 = lane + 32 * k;
    v[k] = 4 * c4 < H ? xr[c4] : make_float4(0.f, 0.f, 0.f, 0.f);
    sum += (v[k].x + v[k].y) + (v[k].z + v[k].w);
  }
  const float mean = warp_sum(sum) / (float)H;
  float m2 = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    if (4 * (lane + 32 * k) < H) {
      const float a = v[k].x - mean, b = v[k].y - mean, c = v[k].z - mean, d = v[k].w - mean;
      m2 += (a * a + b * b) + (c * c + d * d);
    }
  }
  const float rs = 1.0f / sqrtf(warp_sum(m2) / (float)H + eps);
  float4* hr = reinterpret_cast<float4*>(h + row * H);
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int c4 = lane + 32 * k;
    if (4 * c4 < H) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c4);
      const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + c4);
      float4 x, o;
      x.x = (v[k].x - mean) * rs; x.y = (v[k].y - mean) * rs; x.z = (v[k].z - mean) * rs; x.w = (v[k].w - mean) * rs;
      const float a0 = fmaf(x.x, g.x, b.x), a1 = fmaf(x.y, g.y, b.y), a2 = fmaf(x.z, g.z, b.z), a3 = fmaf(x.w, g.w, b.w);
      o.x = a0 / (1.0f + expf(-a0)); o.y = a1 / (1.0f + expf(-a1));
      o.z = a2 / (1.0f + expf(-a2)); o.w = a3 / (1.0f + expf(-a3));
      xr[c4] = x;
      hr[c4] = o;
    }
  }
  if (lane == 0 && rstd) rstd[row] = rs;
}

// g: dh in, du out (in place).  a = gamma x_hat + beta, da = dh SiLU'(a), d x_hat = da gamma,
// du = rstd (d x_hat - mean(d x_hat) - x_hat mean(d x_hat x_hat));  column sums over the CTA's rows:
// dgamma += da x_hat, dbeta += da, dbias += du.
// Warp per row, lane = NV float4 column groups.  Registers decide the occupancy here (the first version kept gamma,
// beta and da live: 174 registers, one CTA per SM, 27 us for 8192 x 512 = 27 % of the HBM rate): gamma / beta are read
// from shared memory, the column sums are accumulated as soon as their terms exist, and the eight warps' partial sums
// are combined through shared memory (no shared atomics) before one global atomic per column per CTA.
template <int NV>
__global__ void __launch_bounds__(256, 2) k_enc_ln_silu_bwd(float* __restrict__ g, const float* __restrict__ xh,
                                                            const float* __restrict__ rstd,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                            float* __restrict__ dbias, long long B, int H, int rows_per_cta) {
  extern __shared__ float smem_f[];   // gamma [H], beta [H], then the flush area [8 warps][3][H]
  NF_PDL_ENTRY();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* sg = smem_f;
  float* sb = smem_f + H;
  float* part = smem_f + 2 * H;
  for (int i = threadIdx.x; i < H; i += blockDim.x) { sg[i] = __ldg(gamma + i); sb[i] = __ldg(beta + i); }
  __syncthreads();
  float4 ag[NV], ab[NV], au[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) ag[k] = ab[k] = au[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  const long long r0 = (long long)blockIdx.x * rows_per_cta;
  const long long r1 = min(B, r0 + rows_per_cta);
  const float invH = 1.0f / (float)H;
  for (long long row = r0 + warp; row < r1; row += 8) {
    float4* gr = reinterpret_cast<float4*>(g + row * H);
    const float4* xr = reinterpret_cast<const float4*>(xh + row * H);
    float4 dh[NV], x[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c4 = lane + 32 * k;
      const bool ok = 4 * c4 < H;
      dh[k] = ok ? gr[c4] : make_float4(0.f, 0.f, 0.f, 0.f);
      x[k] = ok ? __ldg(xr + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const float rs = rstd[row];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c4 = lane + 32 * k;
      const bool ok = 4 * c4 < H;
      const float4 gm = ok ? *reinterpret_cast<const float4*>(sg + 4 * c4) : make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 bt = ok ? *reinterpret_cast<const float4*>(sb + 4 * c4) : make_float4(0.f, 0.f, 0.f, 0.f);
      // returns d x_hat; accumulates da x_hat and da
      auto one = [&](float d, float xx, float gg, float bb, float& accg, float& accb) {
        const float a = fmaf(xx, gg, bb);
        const float sgm = 1.0f / (1.0f + expf(-a));
        const float da = d * sgm * (1.0f + a * (1.0f - sgm));
        accg = fmaf(da, xx, accg);
        accb += da;
        return da * gg;
      };
      dh[k].x = one(dh[k].x, x[k].x, gm.x, bt.x, ag[k].x, ab[k].x);
      dh[k].y = one(dh[k].y, x[k].y, gm.y, bt.y, ag[k].y, ab[k].y);
      dh[k].z = one(dh[k].z, x[k].z, gm.z, bt.z, ag[k].z, ab[k].z);
      dh[k].w = one(dh[k].w, x[k].w, gm.w, bt.w, ag[k].w, ab[k].w);
      s1 += (dh[k].x + dh[k].y) + (dh[k].z + dh[k].w);
      s2 = fmaf(dh[k].x, x[k].x, fmaf(dh[k].y, x[k].y, fmaf(dh[k].z, x[k].z, fmaf(dh[k].w, x[k].w, s2))));
    }
    const float m1 = warp_sum(s1) * invH, m2 = warp_sum(s2) * invH;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c4 = lane + 32 * k;
      float4 du;
      du.x = rs * (dh[k].x - m1 - x[k].x * m2); du.y = rs * (dh[k].y - m1 - x[k].y * m2);
      du.z = rs * (dh[k].z - m1 - x[k].z * m2); du.w = rs * (dh[k].w - m1 - x[k].w * m2);
      if (4 * c4 < H) gr[c4] = du;
      au[k].x += du.x; au[k].y += du.y; au[k].z += du.z; au[k].w += du.w;
    }
  }
  // flush: warp w writes its partial sums to part[w][3][H]; then every thread adds the eight partials of its outputs
  float* mine = part + (size_t)warp * 3 * H;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int c = 4 * (lane + 32 * k);
    if (c < H) {
      *reinterpret_cast<float4*>(mine + c) = ag[k];
      *reinterpret_cast<float4*>(mine + H + c) = ab[k];
      *reinterpret_cast<float4*>(mine + 2 * H + c) = au[k];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 3 * H; i += blockDim.x) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) v += part[(size_t)w * 3 * H + i];
    float* dst = i < H ? dgamma + i : (i < 2 * H ? dbeta + (i - H) : dbias + (i - 2 * H));
    atomicAdd(dst, v);
  }
}

// out[c] += sum over the CTA's rows of A[row][c]      (bias gradient of the output Linear)
__global__ void __launch_bounds__(128) k_enc_colsum(const float* __restrict__ A, long long lda, long long B, int N,
                                                    float* __restrict__ out, int rows_per_cta) {
  NF_PDL_ENTRY();
  const int c = blockIdx.y * 128 + threadIdx.x;
  if (c >= N) return;
  const long long r0 = (long long)blockIdx.x * rows_per_cta, r1 = min(B, r0 + rows_per_cta);
  float acc = 0.f;
  for (long long r = r0; r < r1; ++r) acc += A[r * lda + c];
  atomicAdd(&out[c], acc);
}

int rows_per_cta_for(long long B) {
  // enough CTAs for two waves on 148 SMs, at least 32 rows each (fewer global atomics per row; measured on the kitchen
  // encoder, batch 8192: 16 rows 0.803 ms, 32 rows 0.790 ms, 64 rows 0.830 ms per forward + backward)
  long long rpc = cdiv(B, 296);
  rpc = std::max<long long>(32, cdiv(rpc, 8) * 8);
  return (int)rpc;
}

int ln_silu_fwd(float* xh, float* h, float* rstd, const float* gamma, const float* beta, long long B, int H, float eps,
                cudaStream_t st) {
  if (B <= 0) return 0;
  const unsigned grid = (unsigned)cdiv(B, 8);
  const int nv = (int)cdiv(H, 128);
  switch (nv) {
    case 1: NF_LAUNCH((k_enc_ln_silu_fwd<1>), grid, 256, 0, st, xh, h, rstd, gamma, beta, B, H, eps); break;
    case 2: NF_LAUNCH((k_enc_ln_silu_fwd<2>), grid, 256, 0, st, xh, h, rstd, gamma, beta, B, H, eps); break;
    case 3: case 4: NF_LAUNCH((k_enc_ln_silu_fwd<4>), grid, 256, 0, st, xh, h, rstd, gamma, beta, B, H, eps); break;
    default: NF_LAUNCH((k_enc_ln_silu_fwd<kEncMaxV>), grid, 256, 0, st, xh, h, rstd, gamma, beta, B, H, eps); break;
  }
  NF_LAUNCHED();
  return 0;
}

int ln_silu_bwd(float* g, const float* xh, const float* rstd, const float* gamma, const float* beta, float* dgamma,
                float* dbeta, float* dbias, long long B, int H, cudaStream_t st) {
  if (B <= 0) return 0;
  const int rpc = rows_per_cta_for(B);
  const unsigned grid = (unsigned)cdiv(B, rpc);
  const size_t smem = (size_t)(2 + 8 * 3) * H * sizeof(float);   // gamma, beta, 8 warps x 3 partial-sum rows
  const int nv = (int)cdiv(H, 128);
  {
    static PerDeviceOnce once;   // more than 48 KB of dynamic shared memory from hidden = 512 on
    if (once.first()) {
      constexpr int kMax = (2 + 8 * 3) * 128 * kEncMaxV * (int)sizeof(float);
      NF_CUDA(cudaFuncSetAttribute(k_enc_ln_silu_bwd<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMax));
      NF_CUDA(cudaFuncSetAttribute(k_enc_ln_silu_bwd<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMax));
      NF_CUDA(cudaFuncSetAttribute(k_enc_ln_silu_bwd<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMax));
      NF_CUDA(cudaFuncSetAttribute(k_enc_ln_silu_bwd<kEncMaxV>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMax));
    }
  }
  switch (nv) {
    case 1: NF_LAUNCH((k_enc_ln_silu_bwd<1>), grid, 256, smem, st, g, xh, rstd, gamma, beta, dgamma, dbeta, dbias, B, H, rpc); break;
    case 2: NF_LAUNCH((k_enc_ln_silu_bwd<2>), grid, 256, smem, st, g, xh, rstd, gamma, beta, dgamma, dbeta, dbias, B, H, rpc); break;
    case 3: case 4: NF_LAUNCH((k_enc_ln_silu_bwd<4>), grid, 256, smem, st, g, xh, rstd, gamma, beta, dgamma, dbeta, dbias, B, H, rpc); break;
    default: NF_LAUNCH((k_enc_ln_silu_bwd<kEncMaxV>), grid, 256, smem, st, g, xh, rstd, gamma, beta, dgamma, dbeta, dbias, B, H, rpc); break;
  }
  NF_LAUNCHED();
  return 0;
}

int colsum(const float* A, long long lda, long long B, int N, float* out, cudaStream_t st) {
  if (B <= 0 || N <= 0) return 0;
  const int rpc = rows_per_cta_for(B);
  NF_LAUNCH((k_enc_colsum), dim3((unsigned)cdiv(B, rpc), (unsigned)cdiv(N, 128)), 128, 0, st, A, lda, B, N, out, rpc);
  NF_LAUNCHED();
  return 0;
}

// launch-sequence cache key (graphs.cuh): the encoder's descriptor folded into the flow descriptor's fields
GraphKey enc_key(int op, const NfEncoderDesc* d, int64_t B, int64_t ws_bytes) {
  GraphKey k;
  k.op = op; k.B = B; k.ws_bytes = ws_bytes;
  k.desc.D = d->in_features; k.desc.C = d->hidden; k.desc.H1 = d->n_hidden; k.desc.R = d->out_features;
  k.desc.n_blocks = 0; k.desc.ln_eps = d->ln_eps; k.desc.precision = d->precision;
  return k;
}

// C[r][c] = bias ? bias[c] : 0   (base value of a split-K accumulation)
__global__ void __launch_bounds__(256) k_enc_fill_rows(float* __restrict__ C, long long n, int N,
                                                       const float* __restrict__ bias) {
  NF_PDL_ENTRY();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    C[i] = bias ? __ldg(bias + (int)(i % N)) : 0.f;
}

// C (B, N) = A (B, K) @ W (N, K)^T + bias
int enc_linear(int prec, const float* A, int K, const float* W, const float* bias, float* C, int N, int64_t B,
               cudaStream_t st) {
  TcGemmP t;
  t.A0 = A; t.lda0 = K; t.B0 = W; t.ldb0 = K; t.C = C; t.ldc = N; t.bias = bias;
  t.M = (int)B; t.N = N; t.K0 = K;
  // Few output tiles and a long reduction (the ant encoder's first layer at batch 256: 8 tiles, K = 4100): start C
  // at the bias and accumulate, which lets the GEMM split the reduction over the idle SMs.
  if (prec_tc(prec) && K >= 512 && cdiv(B, 128) * cdiv(N, 128) * 2 <= 148 && tc_gemm_supported(t)) {
    const long long n = B * N;
    NF_LAUNCH((k_enc_fill_rows), (unsigned)std::min<long long>(cdiv(n, 256), 148 * 4), 256, 0, st, C, n, N, bias);
    NF_LAUNCHED();
    t.bias = nullptr; t.accumulate = 1;
  }
  return linear_nt(prec, t, st);
}

}  // namespace
}  // namespace nf

using namespace nf;

extern "C" {

NF_API int64_t nf_encoder_param_layout(const NfEncoderDesc* desc, int64_t* offs) {
  if (check_enc(desc) != 0) return -1;
  const EncDims m = enc_dims(desc);
  const EncLayout l = enc_layout(m);
  if (offs) {
    for (int i = 0; i < m.L; ++i) {
      offs[4 * i + 0] = l.w[i]; offs[4 * i + 1] = l.b[i]; offs[4 * i + 2] = l.lw[i]; offs[4 * i + 3] = l.lb[i];
    }
    offs[4 * m.L] = l.wo; offs[4 * m.L + 1] = l.bo;
  }
  return l.total;
}

NF_API int64_t nf_encoder_stash_bytes(const NfEncoderDesc* desc, int64_t B) {
  if (check_enc(desc) != 0 || B < 0) return -1;
  return std::max<int64_t>(16, enc_stash(enc_dims(desc), B).total * 4);
}

NF_API int64_t nf_encoder_workspace_bytes(const NfEncoderDesc* desc, int64_t B) {
  if (check_enc(desc) != 0 || B < 0) return -1;
  return std::max<int64_t>(16, enc_ws(enc_dims(desc), B).total * 4);
}

NF_API int nf_encoder_forward(const NfEncoderDesc* desc, const float* params, const float* g, int64_t B, float* out,
                              void* stash, void* workspace, int64_t ws_bytes, void* stream) {
  NF_TRY(check_enc(desc));
  NF_CHECK_ARG(B >= 0 && B < (int64_t(1) << 30), NF_E_BADARG, "B=%lld out of range", (long long)B);
  const EncDims m = enc_dims(desc);
  const EncLayout l = enc_layout(m);
  const EncStash sl = enc_stash(m, B);
  const EncWs wl = enc_ws(m, B);
  NF_TRY(check_ptr(params, "params", true));
  if (B == 0) return 0;
  NF_TRY(check_ptr(g, "g", true)); NF_TRY(check_ptr(out, "out", true));
  NF_TRY(check_ptr(stash, "stash", false)); NF_TRY(check_ptr(workspace, "workspace", true));
  NF_CHECK_ARG(ws_bytes >= wl.total * 4, NF_E_WORKSPACE, "workspace too small: %lld < %lld bytes", (long long)ws_bytes,
               (long long)wl.total * 4);
  cudaStream_t user_st = static_cast<cudaStream_t>(stream);
  float* ws = static_cast<float*>(workspace);
  float* sb = static_cast<float*>(stash);
  float* g_in = sb ? sb + sl.g : ws + wl.io_g;
  float* out_st = ws + wl.io_out;
  {
    CopySegs c;
    c.add(g_in, g, B * m.in);
    NF_TRY(copy_segments(c, user_st));
  }
  GraphKey key = enc_key(20, desc, B, ws_bytes);
  key.ptr[0] = params; key.ptr[1] = stash; key.ptr[2] = workspace;
  auto body = [&](cudaStream_t st) -> int {
    const float* hin = g_in;
    int K = m.in;
    for (int i = 0; i < m.L; ++i) {
      // with a stash every layer keeps its x_hat / h / rstd; without one the two workspace buffers alternate roles
      float* xh = sb ? sb + sl.layers + i * sl.layer_stride + sl.xh : ws + wl.a;
      float* h = sb ? sb + sl.layers + i * sl.layer_stride + sl.h : ws + wl.b;
      float* rs = sb ? sb + sl.layers + i * sl.layer_stride + sl.rstd : nullptr;
      NF_TRY(enc_linear(m.prec, hin, K, params + l.w[i], params + l.b[i], xh, m.H, B, st));        // fc_i   (:170)
      NF_TRY(ln_silu_fwd(xh, h, rs, params + l.lw[i], params + l.lb[i], B, m.H, m.eps, st));       // norm_i, silu (:171-172)
      hin = h; K = m.H;
    }
    return enc_linear(m.prec, hin, K, params + l.wo, params + l.bo, out_st, m.out, B, st);         // fc_out (:186)
  };
  NF_TRY(run_graphed(key, user_st, body));
  CopySegs c;
  c.add(out, out_st, B * m.out);
  return copy_segments(c, user_st);
}

NF_API int nf_encoder_backward(const NfEncoderDesc* desc, const float* params, const void* stash, int64_t B,
                               const float* dout, float* dg, float* dparams, void* workspace, int64_t ws_bytes,
                               void* stream) {
  NF_TRY(check_enc(desc));
  NF_CHECK_ARG(B >= 0 && B < (int64_t(1) << 30), NF_E_BADARG, "B=%lld out of range", (long long)B);
  const EncDims m = enc_dims(desc);
  const EncLayout l = enc_layout(m);
  const EncStash sl = enc_stash(m, B);
  const EncWs wl = enc_ws(m, B);
  NF_TRY(check_ptr(params, "params", true)); NF_TRY(check_ptr(dparams, "dparams", true));
  cudaStream_t user_st = static_cast<cudaStream_t>(stream);
  if (B == 0) return fill_zero(dparams, l.total, user_st);
  NF_TRY(check_ptr(stash, "stash", true)); NF_TRY(check_ptr(dout, "dout", true));
  NF_TRY(check_ptr(dg, "dg", false)); NF_TRY(check_ptr(workspace, "workspace", true));
  NF_CHECK_ARG(ws_bytes >= wl.total * 4, NF_E_WORKSPACE, "workspace too small: %lld < %lld bytes", (long long)ws_bytes,
               (long long)wl.total * 4);
  float* ws = static_cast<float*>(workspace);
  const float* sb = static_cast<const float*>(stash);
  const float* g_in = sb + sl.g;
  float* dout_st = ws + wl.io_out;      // forward's output staging is free by now
  float* dg_st = ws + wl.io_dg;
  {
    CopySegs c;
    c.add(dout_st, dout, B * m.out);
    NF_TRY(copy_segments(c, user_st));
  }
  GraphKey key = enc_key(21, desc, B, ws_bytes);
  key.ptr[0] = params; key.ptr[1] = stash; key.ptr[2] = dparams; key.ptr[3] = workspace;
  key.flags = dg ? 1 : 0;
  auto body = [&](cudaStream_t st) -> int {
    NF_TRY(fill_zero(dparams, l.total, st));
    float* wt = ws + wl.wt;
    auto layer = [&](int i, int64_t off) { return sb + sl.layers + i * sl.layer_stride + off; };
    // output Linear: dW_out += dout^T h_L, db_out = colsum(dout), dh_L = dout W_out
    {
      TcGemmTnP w;
      w.A = dout_st; w.lda = m.out; w.B = layer(m.L - 1, sl.h); w.ldb = m.H; w.C = dparams + l.wo; w.ldc = m.H;
      w.M = m.out; w.N = m.H; w.K = B;
      NF_TRY(wgrad_tn(m.prec, w, st));
      NF_TRY(colsum(dout_st, m.out, B, m.out, dparams + l.bo, st));
      NF_TRY(transpose2d(wt, m.out, 0, 0, params + l.wo, m.H, 0, 0, m.out, m.H, 1, 1, st));   // (out, H) -> (H, out)
      NF_TRY(enc_linear(m.prec, dout_st, m.out, wt, nullptr, ws + wl.a, m.H, B, st));
    }
    float* cur = ws + wl.a;     // dh_i, then du_i in place
    float* other = ws + wl.b;
    for (int i = m.L - 1; i >= 0; --i) {
      const int K = i == 0 ? m.in : m.H;
      NF_TRY(ln_silu_bwd(cur, layer(i, sl.xh), layer(i, sl.rstd), params + l.lw[i], params + l.lb[i],
                         dparams + l.lw[i], dparams + l.lb[i], dparams + l.b[i], B, m.H, st));
      TcGemmTnP w;   // dW_i += du_i^T h_{i-1}
      w.A = cur; w.lda = m.H; w.B = i == 0 ? g_in : layer(i - 1, sl.h); w.ldb = K; w.C = dparams + l.w[i]; w.ldc = K;
      w.M = m.H; w.N = K; w.K = B;
      NF_TRY(wgrad_tn(m.prec, w, st));
      if (i > 0 || dg) {
        // dh_{i-1} = du_i W_i : the K-major B operand is W_i^T (K, H)
        NF_TRY(transpose2d(wt, m.H, 0, 0, params + l.w[i], K, 0, 0, m.H, K, 1, 1, st));
        float* dst = i == 0 ? dg_st : other;
        NF_TRY(enc_linear(m.prec, cur, m.H, wt, nullptr, dst, K, B, st));
        std::swap(cur, other);
      }
    }
    return 0;
  };
  NF_TRY(run_graphed(key, user_st, body));
  CopySegs c;
  c.add(dg, dg_st, B * m.in);
  return copy_segments(c, user_st);
}

}  // extern "C"
