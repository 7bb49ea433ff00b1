// Row-wise and parameter-sized kernels of the flow (sm_100a).
#include "elementwise.cuh"

#include <algorithm>
#include "prof.cuh"

namespace nf {

namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------------------------------
// LeakyReLU + LayerNorm (no affine: gamma/beta are folded into the next Linear, see fold_ln_affine)
// one warp per row; three passes over the row (mean, centred variance, normalise) like
// nn.LayerNorm's two-pass statistics (torch_fcnf.py:75-76).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_lrelu_ln_fwd(LnFwdP p) {
  NF_PDL_ENTRY();
  const int net = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= p.B) return;
  const float* u = p.u + net * p.u_net + row * p.N;
  const float* bias = p.bias ? p.bias + net * p.bias_net : nullptr;
  float* xh = p.xh + net * p.xh_net + row * p.N;
  float* xlo = p.xh_lo ? p.xh_lo + net * p.xh_lo_net + row * p.N : nullptr;
  const int N = p.N;

  float sum = 0.f;
  for (int c = lane; c < N; c += 32) {
    float v = u[c] + (bias ? bias[c] : 0.f);
    v = v > 0.f ? v : kLeakySlope * v;
    sum += v;
  }
  const float mean = warp_sum(sum) / (float)N;
  float sq = 0.f;
  for (int c = lane; c < N; c += 32) {
    float v = u[c] + (bias ? bias[c] : 0.f);
    v = v > 0.f ? v : kLeakySlope * v;
    const float d = p.fast_var ? v : v - mean;
    sq = fmaf(d, d, sq);
  }
  float var = warp_sum(sq) / (float)N;
  if (p.fast_var) var = fmaxf(var - mean * mean, 0.f);   // flax: E[x^2] - E[x]^2, clipped at 0
  const float rstd = rsqrtf(var + p.eps);
  uint32_t* mask = p.mask ? p.mask + net * p.mask_net + row * p.mw : nullptr;
  for (int c0 = 0; c0 < N; c0 += 32) {
    const int c = c0 + lane;
    float pre = 0.f;
    if (c < N) {
      pre = u[c] + (bias ? bias[c] : 0.f);
      const float v = pre > 0.f ? pre : kLeakySlope * pre;
      const float o = (v - mean) * rstd;
      if (xlo) {
        const float hi = __uint_as_float(__float_as_uint(o) & 0xffffe000u);
        xh[c] = hi;
        xlo[c] = o - hi;
      } else {
        xh[c] = o;
      }
    }
    const uint32_t bits = __ballot_sync(0xffffffffu, c < N && pre > 0.f);
    if (mask && lane == 0) mask[c0 >> 5] = bits;
  }
  if (p.rstd && lane == 0) p.rstd[net * p.rstd_net + row] = rstd;
}

// Vectorised single-pass versions of the two kernels around this comment (N % 4 == 0, N <= 1024, 16-byte aligned
// rows): a warp keeps its row in registers as NV float4 per lane, so u / g / x_hat are read exactly once with 128-bit
// accesses (the scalar kernels read u three times and g, x_hat twice; measured on C3, B = 16384, H1 = 320:
// backward 119 us, forward 28 us for 126 / 84 MB of algorithmic traffic).
template <int NV>
__global__ void __launch_bounds__(256) k_lrelu_ln_fwd_v4(LnFwdP p) {
  NF_PDL_ENTRY();
  const int net = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= p.B) return;
  const int N = p.N;
  const float4* u = reinterpret_cast<const float4*>(p.u + net * p.u_net + row * N);
  const float4* bias = p.bias ? reinterpret_cast<const float4*>(p.bias + net * p.bias_net) : nullptr;
  float4* xh = reinterpret_cast<float4*>(p.xh + net * p.xh_net + row * N);
  uint32_t* mask = p.mask ? p.mask + net * p.mask_net + row * p.mw : nullptr;
  float4 v[NV];
  uint32_t nib[NV];
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int c4 = lane + 32 * k;
    nib[k] = 0u;
    v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (4 * c4 < N) {
      float4 t = u[c4];
      if (bias) { const float4 b = __ldg(bias + c4); t.x += b.x; t.y += b.y; t.z += b.z; t.w += b.w; }
      nib[k] = (t.x > 0.f ? 1u : 0u) | (t.y > 0.f ? 2u : 0u) | (t.z > 0.f ? 4u : 0u) | (t.w > 0.f ? 8u : 0u);
      t.x = t.x > 0.f ? t.x : kLeakySlope * t.x; t.y = t.y > 0.f ? t.y : kLeakySlope * t.y;
      t.z = t.z > 0.f ? t.z : kLeakySlope * t.z; t.w = t.w > 0.f ? t.w : kLeakySlope * t.w;
      v[k] = t;
      sum += (t.x + t.y) + (t.z + t.w);
    }
  }
  const float mean = warp_sum(sum) / (float)N;
  float sq = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    if (4 * (lane + 32 * k) < N) {
      const float sh = p.fast_var ? 0.f : mean;
      const float a = v[k].x - sh, b = v[k].y - sh, c = v[k].z - sh, d = v[k].w - sh;
      sq += (a * a + b * b) + (c * c + d * d);
    }
  }
  float var = warp_sum(sq) / (float)N;
  if (p.fast_var) var = fmaxf(var - mean * mean, 0.f);   // flax: E[x^2] - E[x]^2, clipped at 0
  const float rstd = rsqrtf(var + p.eps);
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int c4 = lane + 32 * k;
    if (4 * c4 < N)
      xh[c4] = make_float4((v[k].x - mean) * rstd, (v[k].y - mean) * rstd, (v[k].z - mean) * rstd, (v[k].w - mean) * rstd);
    // sign mask: 8 consecutive lanes hold the 32 bits of word (lane >> 3) + 4 k
    uint32_t w = nib[k] << ((lane & 7) * 4);
    w |= __shfl_xor_sync(0xffffffffu, w, 1);
    w |= __shfl_xor_sync(0xffffffffu, w, 2);
    w |= __shfl_xor_sync(0xffffffffu, w, 4);
    const int word = (lane >> 3) + 4 * k;
    if (mask && (lane & 7) == 0 && 32 * word < N) mask[word] = w;
  }
  if (p.rstd && lane == 0) p.rstd[net * p.rstd_net + row] = rstd;
}

template <int NV>
__global__ void __launch_bounds__(256, 2) k_ln_lrelu_bwd_v4(LnBwdP p, int rows_per_cta) {
  extern __shared__ float s_part[];   // [8 warps][N] column sums
  NF_PDL_ENTRY();
  const int net = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int N = p.N;
  float4 cs[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) cs[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  const long long r0 = (long long)blockIdx.x * rows_per_cta;
  const long long r1 = min(p.B, r0 + rows_per_cta);
  const float invN = 1.0f / (float)N;
  for (long long row = r0 + warp; row < r1; row += 8) {
    float4* g = reinterpret_cast<float4*>(p.g + net * p.g_net + row * N);
    const float4* xh = reinterpret_cast<const float4*>(p.xh + net * p.xh_net + row * N);
    const uint32_t* mask = p.mask + net * p.mask_net + row * p.mw;
    float4 gv[NV], xv[NV];
    uint32_t mw[NV];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c4 = lane + 32 * k;
      const bool ok = 4 * c4 < N;
      gv[k] = ok ? g[c4] : make_float4(0.f, 0.f, 0.f, 0.f);
      xv[k] = ok ? __ldg(xh + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
      mw[k] = ok ? __ldg(mask + (lane >> 3) + 4 * k) : 0u;
    }
    const float rstd = p.rstd[net * p.rstd_net + row];
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      s1 += (gv[k].x + gv[k].y) + (gv[k].z + gv[k].w);
      s2 = fmaf(gv[k].x, xv[k].x, fmaf(gv[k].y, xv[k].y, fmaf(gv[k].z, xv[k].z, fmaf(gv[k].w, xv[k].w, s2))));
    }
    const float m1 = warp_sum(s1) * invN, m2 = warp_sum(s2) * invN;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c4 = lane + 32 * k;
      const uint32_t bits = mw[k] >> ((lane & 7) * 4);
      float4 d;
      d.x = rstd * (gv[k].x - m1 - xv[k].x * m2) * ((bits & 1u) ? 1.0f : kLeakySlope);
      d.y = rstd * (gv[k].y - m1 - xv[k].y * m2) * ((bits & 2u) ? 1.0f : kLeakySlope);
      d.z = rstd * (gv[k].z - m1 - xv[k].z * m2) * ((bits & 4u) ? 1.0f : kLeakySlope);
      d.w = rstd * (gv[k].w - m1 - xv[k].w * m2) * ((bits & 8u) ? 1.0f : kLeakySlope);
      if (4 * c4 < N) {
        g[c4] = d;
        cs[k].x += d.x; cs[k].y += d.y; cs[k].z += d.z; cs[k].w += d.w;
      }
    }
  }
  if (p.colsum) {   // uniform across the CTA
    float* mine = s_part + (size_t)warp * N;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c = 4 * (lane + 32 * k);
      if (c < N) *reinterpret_cast<float4*>(mine + c) = cs[k];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < N; c += blockDim.x) {
      float v = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) v += s_part[(size_t)w * N + c];
      if (v != 0.f) atomicAdd(&p.colsum[net * p.colsum_net + c], v);
    }
  }
}

// d u = LeakyReLU'(u) * rstd * (g - mean(g) - x_hat * mean(g * x_hat)),  g = d x_hat
// One warp owns kBwdRowsPerWarp rows: a statistics pass, then per 256-column chunk a pass that
// writes d u and keeps the column sums (bias gradient) in registers; the warps of a CTA are
// summed through shared memory and added to global memory with one atomic per column and CTA.
constexpr int kBwdRowsPerWarp = 8;
constexpr int kBwdWarps = 8;
__global__ void __launch_bounds__(kBwdWarps * 32) k_ln_lrelu_bwd(LnBwdP p, int rpw) {
  NF_PDL_ENTRY();
  __shared__ float s_col[kBwdWarps][256];
  const int net = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int N = p.N;
  const long long row0 = ((long long)blockIdx.x * kBwdWarps + warp) * rpw;
  float m1[kBwdRowsPerWarp], m2[kBwdRowsPerWarp];
#pragma unroll
  for (int r = 0; r < kBwdRowsPerWarp; ++r) {
    const long long row = row0 + r;
    float s1 = 0.f, s2 = 0.f;
    if (r < rpw && row < p.B) {
      const float* g = p.g + net * p.g_net + row * N;
      const float* xh = p.xh + net * p.xh_net + row * N;
      for (int c = lane; c < N; c += 32) {
        const float gv = g[c];
        s1 += gv;
        s2 = fmaf(gv, xh[c], s2);
      }
    }
    m1[r] = warp_sum(s1) / (float)N;
    m2[r] = warp_sum(s2) / (float)N;
  }
  for (int ch = 0; ch * 256 < N; ++ch) {
    float cs[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) cs[k] = 0.f;
#pragma unroll
    for (int r = 0; r < kBwdRowsPerWarp; ++r) {
      const long long row = row0 + r;
      if (r < rpw && row < p.B) {
        float* g = p.g + net * p.g_net + row * N;
        const float* xh = p.xh + net * p.xh_net + row * N;
        const uint32_t* mask = p.mask + net * p.mask_net + row * p.mw;
        const float rstd = p.rstd[net * p.rstd_net + row];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int c = ch * 256 + k * 32 + lane;
          if (c < N) {
            float dv = rstd * (g[c] - m1[r] - xh[c] * m2[r]);
            dv *= ((mask[c >> 5] >> lane) & 1u) ? 1.0f : kLeakySlope;
            g[c] = dv;
            cs[k] += dv;
          }
        }
      }
    }
    if (p.colsum) {   // uniform across the CTA
#pragma unroll
      for (int k = 0; k < 8; ++k) s_col[warp][k * 32 + lane] = cs[k];
      __syncthreads();
      const int c = ch * 256 + threadIdx.x;
      if (threadIdx.x < 256 && c < N) {
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < kBwdWarps; ++w) v += s_col[w][threadIdx.x];
        if (v != 0.f) atomicAdd(&p.colsum[net * p.colsum_net + c], v);
      }
      __syncthreads();
    }
  }
}

// ------------------------------------------------------------------------------------------
// Affine coupling + log-det (torch_fcnf.py:101-104 forward, :112-113 inverse).
// Small D: one thread per row (adjacent threads touch adjacent rows -> coalesced, 128-bit when
// the halves are multiples of 4).  Large D: one warp per row.
// ------------------------------------------------------------------------------------------
template <int VEC, bool INVERSE>
__global__ void __launch_bounds__(256) k_affine_rowthread(const float* __restrict__ in, const float* __restrict__ s,
                                                          const float* __restrict__ t, const float* __restrict__ bs,
                                                          const float* __restrict__ bt, const float* cdev, float chost,
                                                          long long B, int D, int ldi, int ldo, float* __restrict__ out,
                                                          float* __restrict__ logdet, float* __restrict__ s_out) {
  NF_PDL_ENTRY();
  const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= B) return;
  const int dc = (D + 1) / 2, dt = D / 2;
  const float* xi = in + row * ldi;
  float* xo = out + row * ldo;
  const float* sr = s + row * dt;
  const float* tr = t + row * dt;
  float ssum = 0.f;
  if (VEC == 4) {
    for (int j = 0; j < dc; j += 4) *reinterpret_cast<float4*>(xo + j) = *reinterpret_cast<const float4*>(xi + j);
    for (int j = 0; j < dt; j += 4) {
      float4 sv = *reinterpret_cast<const float4*>(sr + j);
      float4 tv = *reinterpret_cast<const float4*>(tr + j);
      const float4 xv = *reinterpret_cast<const float4*>(xi + dc + j);
      if (bs) { sv.x += bs[j]; sv.y += bs[j + 1]; sv.z += bs[j + 2]; sv.w += bs[j + 3]; }
      if (bt) { tv.x += bt[j]; tv.y += bt[j + 1]; tv.z += bt[j + 2]; tv.w += bt[j + 3]; }
      float4 o;
      if (!INVERSE) {
        o.x = (xv.x - tv.x) * expf(-sv.x); o.y = (xv.y - tv.y) * expf(-sv.y);
        o.z = (xv.z - tv.z) * expf(-sv.z); o.w = (xv.w - tv.w) * expf(-sv.w);
      } else {
        o.x = fmaf(xv.x, expf(sv.x), tv.x); o.y = fmaf(xv.y, expf(sv.y), tv.y);
        o.z = fmaf(xv.z, expf(sv.z), tv.z); o.w = fmaf(xv.w, expf(sv.w), tv.w);
      }
      ssum += (sv.x + sv.y) + (sv.z + sv.w);
      *reinterpret_cast<float4*>(xo + dc + j) = o;
      if (s_out) *reinterpret_cast<float4*>(s_out + row * dt + j) = sv;
    }
  } else {
    for (int j = 0; j < dc; ++j) xo[j] = xi[j];
    for (int j = 0; j < dt; ++j) {
      const float sv = sr[j] + (bs ? bs[j] : 0.f);
      const float tv = tr[j] + (bt ? bt[j] : 0.f);
      const float xv = xi[dc + j];
      xo[dc + j] = INVERSE ? fmaf(xv, expf(sv), tv) : (xv - tv) * expf(-sv);
      ssum += sv;
      if (s_out) s_out[row * dt + j] = sv;
    }
  }
  if (logdet) {
    const float c = chost + (cdev ? *cdev : 0.f);
    logdet[row] += INVERSE ? (ssum - c) : (c - ssum);
  }
}

template <bool INVERSE>
__global__ void __launch_bounds__(256) k_affine_rowwarp(const float* __restrict__ in, const float* __restrict__ s,
                                                        const float* __restrict__ t, const float* __restrict__ bs,
                                                        const float* __restrict__ bt, const float* cdev, float chost,
                                                        long long B, int D, int ldi, int ldo, float* __restrict__ out,
                                                        float* __restrict__ logdet, float* __restrict__ s_out) {
  NF_PDL_ENTRY();
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= B) return;
  const int dc = (D + 1) / 2, dt = D / 2;
  const float* xi = in + row * ldi;
  float* xo = out + row * ldo;
  for (int j = lane; j < dc; j += 32) xo[j] = xi[j];
  float ssum = 0.f;
  for (int j = lane; j < dt; j += 32) {
    const float sv = s[row * dt + j] + (bs ? bs[j] : 0.f);
    const float tv = t[row * dt + j] + (bt ? bt[j] : 0.f);
    const float xv = xi[dc + j];
    xo[dc + j] = INVERSE ? fmaf(xv, expf(sv), tv) : (xv - tv) * expf(-sv);
    ssum += sv;
    if (s_out) s_out[row * dt + j] = sv;
  }
  ssum = warp_sum(ssum);
  if (logdet && lane == 0) {
    const float c = chost + (cdev ? *cdev : 0.f);
    logdet[row] += INVERSE ? (ssum - c) : (c - ssum);
  }
}

// Vectorised form for D % 8 == 0 (both halves are whole float4 chunks): G = 2^k lanes share a row (G >= chunks per half,
// capped at 32), a warp covers 32 / G consecutive rows, every access is a 128-bit load / store that is contiguous across the
// lanes of a row (and across rows for s / t), the row sum of s is an xor-shuffle reduction inside the lane group.
// Replaces row-per-thread (D <= 64: a warp instruction touched 32 different lines) and scalar warp-per-row (D > 64).
template <int G, bool INVERSE>
__global__ void __launch_bounds__(256) k_affine_vec(const float* __restrict__ in, const float* __restrict__ s,
                                                    const float* __restrict__ t, const float* __restrict__ bs,
                                                    const float* __restrict__ bt, const float* cdev, float chost,
                                                    long long B, int D, int ldi, int ldo, float* __restrict__ out,
                                                    float* __restrict__ logdet, float* __restrict__ s_out) {
  NF_PDL_ENTRY();
  constexpr int RPW = 32 / G;                     // rows per warp
  const int lane = threadIdx.x & 31, g = lane % G;
  const long long row = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * RPW + lane / G;
  const bool ok = row < B;
  const int h4 = D >> 3;                          // float4 chunks per half (dc == dt)
  float ssum = 0.f;
  if (ok) {
    const float4* xi = reinterpret_cast<const float4*>(in + row * ldi);
    float4* xo = reinterpret_cast<float4*>(out + row * ldo);
    const float4* sr = reinterpret_cast<const float4*>(s + row * (D >> 1));
    const float4* tr = reinterpret_cast<const float4*>(t + row * (D >> 1));
    float4* so = s_out ? reinterpret_cast<float4*>(s_out + row * (D >> 1)) : nullptr;
    // two chunks per trip, all eight loads issued before the first use (D = 400: 50 chunks per half over 32 lanes)
    for (int c = g; c < h4; c += 2 * G) {
      const int c2 = c + G;
      const bool two = c2 < h4;
      const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 xc0 = xi[c], xv0 = xi[h4 + c];
      float4 sv0 = sr[c], tv0 = tr[c];
      const float4 xc1 = two ? xi[c2] : zero, xv1 = two ? xi[h4 + c2] : zero;
      float4 sv1 = two ? sr[c2] : zero, tv1 = two ? tr[c2] : zero;
      auto one = [&](int cc, const float4& xc, const float4& xv, float4& sv, float4& tv) {
        if (bs) { const float4 b = __ldg(reinterpret_cast<const float4*>(bs) + cc); sv.x += b.x; sv.y += b.y; sv.z += b.z; sv.w += b.w; }
        if (bt) { const float4 b = __ldg(reinterpret_cast<const float4*>(bt) + cc); tv.x += b.x; tv.y += b.y; tv.z += b.z; tv.w += b.w; }
        float4 o;
        if (!INVERSE) {
          o.x = (xv.x - tv.x) * expf(-sv.x); o.y = (xv.y - tv.y) * expf(-sv.y);
          o.z = (xv.z - tv.z) * expf(-sv.z); o.w = (xv.w - tv.w) * expf(-sv.w);
        } else {
          o.x = fmaf(xv.x, expf(sv.x), tv.x); o.y = fmaf(xv.y, expf(sv.y), tv.y);
          o.z = fmaf(xv.z, expf(sv.z), tv.z); o.w = fmaf(xv.w, expf(sv.w), tv.w);
        }
        ssum += (sv.x + sv.y) + (sv.z + sv.w);
        xo[cc] = xc;
        xo[h4 + cc] = o;
        if (so) so[cc] = sv;
      };
      one(c, xc0, xv0, sv0, tv0);
      if (two) one(c2, xc1, xv1, sv1, tv1);
    }
  }
#pragma unroll
  for (int o = G >> 1; o > 0; o >>= 1) ssum += __shfl_xor_sync(0xffffffffu, ssum, o);
  if (ok && g == 0 && logdet) {
    const float c = chost + (cdev ? *cdev : 0.f);
    logdet[row] += INVERSE ? (ssum - c) : (c - ssum);
  }
}

// backward, same lane grouping (no column sums): dxp_c = g_c; forward: dxp_t = g e^-s, ds = -g z_t - dld, dt = -g e^-s;
// reverse direction: dxp_t = g e^s, ds = g z_t e^s + dld, dt = g
template <int G>
__global__ void __launch_bounds__(256) k_affine_bwd_vec(const float* __restrict__ dz, const float* __restrict__ z,
                                                        const float* __restrict__ sfull, const float* __restrict__ dld,
                                                        long long B, int D, float* __restrict__ dxp,
                                                        float* __restrict__ ds, float* __restrict__ dtt, int inverse) {
  NF_PDL_ENTRY();
  constexpr int RPW = 32 / G;
  const int lane = threadIdx.x & 31, g = lane % G;
  const long long row = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * RPW + lane / G;
  if (row >= B) return;
  const int h4 = D >> 3;
  const float4* gr = dz ? reinterpret_cast<const float4*>(dz + row * D) : nullptr;
  const float4* zr = reinterpret_cast<const float4*>(z + row * D);
  const float4* sr = reinterpret_cast<const float4*>(sfull + row * (D >> 1));
  float4* xo = reinterpret_cast<float4*>(dxp + row * D);
  float4* dso = reinterpret_cast<float4*>(ds + row * (D >> 1));
  float4* dto = reinterpret_cast<float4*>(dtt + row * (D >> 1));
  const float gl = dld ? dld[row] : 0.f;
#pragma unroll 2
  for (int c = g; c < h4; c += G) {
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 gc = gr ? gr[c] : zero, gt = gr ? gr[h4 + c] : zero;
    const float4 zt = zr[h4 + c], sv = sr[c];
    float4 ox, od, ot;
    if (inverse) {
      ox.x = gt.x * expf(sv.x); ox.y = gt.y * expf(sv.y); ox.z = gt.z * expf(sv.z); ox.w = gt.w * expf(sv.w);
      od.x = fmaf(ox.x, zt.x, gl); od.y = fmaf(ox.y, zt.y, gl); od.z = fmaf(ox.z, zt.z, gl); od.w = fmaf(ox.w, zt.w, gl);
      ot = gt;
    } else {
      ox.x = gt.x * expf(-sv.x); ox.y = gt.y * expf(-sv.y); ox.z = gt.z * expf(-sv.z); ox.w = gt.w * expf(-sv.w);
      od.x = -gt.x * zt.x - gl; od.y = -gt.y * zt.y - gl; od.z = -gt.z * zt.z - gl; od.w = -gt.w * zt.w - gl;
      ot.x = -ox.x; ot.y = -ox.y; ot.z = -ox.z; ot.w = -ox.w;
    }
    xo[c] = gc;
    xo[h4 + c] = ox;
    dso[c] = od;
    dto[c] = ot;
  }
}

// backward of the affine coupling: per element, plus column sums of ds/dt (last-layer bias grads)
__global__ void __launch_bounds__(256) k_affine_bwd(const float* __restrict__ dz, const float* __restrict__ z,
                                                    const float* __restrict__ sfull, const float* __restrict__ dld,
                                                    long long B, int D, float* __restrict__ dxp,
                                                    float* __restrict__ ds, float* __restrict__ dtt,
                                                    float* g3s, float* g3t, int inverse) {
  NF_PDL_ENTRY();
  extern __shared__ float sh[];  // 2*dt partial column sums
  const int dc = (D + 1) / 2, dt = D / 2;
  for (int i = threadIdx.x; i < 2 * dt; i += blockDim.x) sh[i] = 0.f;
  __syncthreads();
  const long long total = B * (long long)D;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long row = idx / D;
    const int j = (int)(idx - row * D);
    const float g = dz ? dz[idx] : 0.f;
    if (j < dc) {
      dxp[idx] = g;
    } else {
      const int jt = j - dc;
      const float sv = sfull[row * dt + jt];
      float dsv, dtv;
      if (inverse) {
        // reverse direction (torch_fcnf.py:112-113): xp_t = z_t e^s + t, logdet += sum s.  `dz` is the cotangent of
        // xp, `z` the block's z:  dz_t = g e^s, ds = g z_t e^s + dlogdet, dt = g
        const float ge = g * expf(sv);
        dxp[idx] = ge;
        dsv = ge * z[idx] + (dld ? dld[row] : 0.f);
        dtv = g;
      } else {
        const float ge = g * expf(-sv);
        dxp[idx] = ge;
        dsv = -g * z[idx] - (dld ? dld[row] : 0.f);
        dtv = -ge;
      }
      ds[row * dt + jt] = dsv;
      dtt[row * dt + jt] = dtv;
      if (g3s) {
        atomicAdd(&sh[jt], dsv);
        atomicAdd(&sh[dt + jt], dtv);
      }
    }
  }
  __syncthreads();
  if (g3s) {
    for (int i = threadIdx.x; i < dt; i += blockDim.x) {
      atomicAdd(&g3s[i], sh[i]);
      atomicAdd(&g3t[i], sh[dt + i]);
    }
  }
}

// ------------------------------------------------------------------------------------------
// parameter-sized kernels
// ------------------------------------------------------------------------------------------
// Lh = tril(L,-1)+I, Uh = triu(U,1)+diag(s)  (torch_fcnf.py:34-38), logdet_c = sum log|s| (:41)
__global__ void k_plu_mask(const float* params, NfParamLayout pl, int D, float* packed, PackedLayout kl) {
  NF_PDL_ENTRY();
  const int b = blockIdx.y;
  const float* pb = params + (long long)b * pl.block_stride;
  float* kb = packed + (long long)b * kl.block_stride;
  const float* L = pb + pl.slot_off[NF_P_L];
  const float* U = pb + pl.slot_off[NF_P_U];
  const float* s = pb + pl.slot_off[NF_P_S];
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < D * D; idx += gridDim.x * blockDim.x) {
    const int i = idx / D, j = idx % D;
    kb[kl.Lh + idx] = i > j ? L[idx] : (i == j ? 1.f : 0.f);
    kb[kl.Uh + idx] = i < j ? U[idx] : (i == j ? s[i] : 0.f);
  }
  if (blockIdx.x == 0 && threadIdx.x < 32) {
    float acc = 0.f;
    for (int i = threadIdx.x; i < D; i += 32) acc += logf(fabsf(s[i]));
    acc = warp_sum(acc);
    if (threadIdx.x == 0) packed[kl.logdet_c + b] = acc;
  }
}

__global__ void k_logdet_total(float* packed, PackedLayout kl, int n) {
  NF_PDL_ENTRY();
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    float acc = 0.f;
    for (int i = 0; i < n; ++i) acc += packed[kl.logdet_c + i];
    packed[kl.logdet_total] = acc;
  }
}

// Small flow dimensions (D <= 32): the whole chain  mask -> PL = P Lh -> W = PL Uh -> W^T  plus the block's log|det| in ONE
// kernel, one CTA per block with the five D x D matrices in shared memory (as five launches this chain was the critical
// path of nf_flow_prepare: ~20 us per training step on C2a, where the parameters change every step).
__global__ void __launch_bounds__(256) k_plu_build_small(const float* __restrict__ params, NfParamLayout pl, int D,
                                                         float* __restrict__ packed, PackedLayout kl) {
  NF_PDL_ENTRY();
  __shared__ float sP[1024], sLh[1024], sUh[1024], sPL[1024];
  const int b = blockIdx.x, DD = D * D;
  const float* pb = params + (long long)b * pl.block_stride;
  float* kb = packed + (long long)b * kl.block_stride;
  const float* L = pb + pl.slot_off[NF_P_L];
  const float* U = pb + pl.slot_off[NF_P_U];
  const float* s = pb + pl.slot_off[NF_P_S];
  const float* P = pb + pl.slot_off[NF_P_P];
  for (int idx = threadIdx.x; idx < DD; idx += blockDim.x) {
    const int i = idx / D, j = idx % D;
    const float lh = i > j ? L[idx] : (i == j ? 1.f : 0.f);
    const float uh = i < j ? U[idx] : (i == j ? s[i] : 0.f);
    sP[idx] = P[idx]; sLh[idx] = lh; sUh[idx] = uh;
    kb[kl.Lh + idx] = lh; kb[kl.Uh + idx] = uh;
  }
  if (threadIdx.x < 32) {
    float acc = 0.f;
    for (int i = threadIdx.x; i < D; i += 32) acc += logf(fabsf(s[i]));
    acc = warp_sum(acc);
    if (threadIdx.x == 0) packed[kl.logdet_c + b] = acc;
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < DD; idx += blockDim.x) {
    const int i = idx / D, j = idx % D;
    float acc = 0.f;
    for (int k = 0; k < D; ++k) acc = fmaf(sP[i * D + k], sLh[k * D + j], acc);
    sPL[idx] = acc;
    kb[kl.PL + idx] = acc;
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < DD; idx += blockDim.x) {
    const int i = idx / D, j = idx % D;
    float acc = 0.f;
    for (int k = 0; k < D; ++k) acc = fmaf(sPL[i * D + k], sUh[k * D + j], acc);
    kb[kl.W + idx] = acc;
    kb[kl.WT + j * D + i] = acc;
  }
}

// ---- all derived tables of the forward direction in ONE kernel (D <= 32) ---------------------------------------
// nf_flow_prepare(NF_PREP_FWD) was a graph of 9 kernels in three branches joined by the hi/lo split (critical path
// fold -> transpose -> split, ~24 us on C2a, in front of every training step because the parameters change every
// step).  Every table is a pure function of the parameters, so each CTA of this kernel produces its tables directly --
// value, rounded-to-tf32 `hi` and residual `lo` in the same store -- and nothing waits for another CTA:
//   blockIdx.z = block * 2 + net, blockIdx.x = job -- every job is one load -> store pass, no loop-carried latency:
//     32 x 32 tile of W2:   W2f = W2 diag(gamma1) and its transpose W2fT
//     8 rows of W2:         b2f = b2 + W2 beta1 (one warp per row);   one more job: W3f = W3 diag(gamma2), b3f = b3 + W3 beta2
//     32 x 32 tile of W1:   W1T (transpose)
//     8 rows of W1:         W1p (y columns moved to the aligned column dcp)
//     last job, net 0 only: the PLU tables of the block (k_plu_build_small's body) and, in block 0, the total log|det|
__device__ __forceinline__ void put3(float* packed, long long off, float v, long long hi_off, long long lo_off) {
  packed[off] = v;
  if (hi_off) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    const float h = __uint_as_float(r);
    packed[off + hi_off] = h;
    packed[off + lo_off] = v - h;
  }
}

struct PrepJobs { int tF, rB, tW, rP; };   // job counts: W2 tiles, W2 row groups (+1: W3), W1 tiles, W1 row groups (+1: PLU)

__global__ void __launch_bounds__(256) k_prepare_fwd(const float* __restrict__ params, NfParamLayout pl, Dims m,
                                                     float* __restrict__ packed, PackedLayout kl, PrepJobs J) {
  NF_PDL_ENTRY();
  __shared__ float tile[32][33];
  __shared__ float sP[1024], sLh[1024], sUh[1024], sPL[1024];
  const int b = blockIdx.z >> 1, net = blockIdx.z & 1;
  int job = blockIdx.x;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  const float* pn = params + (long long)b * pl.block_stride + net * pl.net_stride;
  const long long kn = (long long)b * kl.block_stride + kl.net0 + net * kl.net_stride;   // offset of this net's tables
  const long long HI = kl.hi_off, LO = kl.lo_off;
  if (job < J.tF) {
    const float* W = pn + pl.slot_off[NF_P_W2];
    const int tk = (m.H1 + 31) / 32, r0 = (job / tk) * 32, k0 = (job % tk) * 32, K = m.H1;
    const int k = k0 + tx;
    const float g = k < K ? pn[pl.slot_off[NF_P_G1] + k] : 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int r = r0 + ty + 8 * q;
      float v = 0.f;
      if (r < m.C && k < K) {
        v = W[(long long)r * K + k] * g;
        put3(packed, kn + kl.W2f + (long long)r * K + k, v, HI, LO);
      }
      tile[ty + 8 * q][tx] = v;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 4; ++q) {                  // W2fT[k, r]: row k0 + ty + 8q, column r0 + tx
      const int kk = k0 + ty + 8 * q, r = r0 + tx;
      if (kk < K && r < m.C) put3(packed, kn + kl.W2fT + (long long)kk * m.C + r, tile[tx][ty + 8 * q], HI, LO);
    }
    return;
  }
  job -= J.tF;
  if (job < J.rB) {                                // b2f: one warp per row, lane-strided like k_fold (same summation order)
    const int r = job * 8 + ty, K = m.H1;
    if (r >= m.C) return;
    const float* W = pn + pl.slot_off[NF_P_W2] + (long long)r * K;
    const float* bet = pn + pl.slot_off[NF_P_BE1];
    float acc = 0.f;
    for (int k = tx; k < K; k += 32) acc = fmaf(W[k], bet[k], acc);
    acc = warp_sum(acc);
    if (tx == 0) put3(packed, kn + kl.b2f + r, pn[pl.slot_off[NF_P_B2] + r] + acc, HI, LO);
    return;
  }
  job -= J.rB;
  if (job == 0) {
    const float* W = pn + pl.slot_off[NF_P_W3];
    const float* gam = pn + pl.slot_off[NF_P_G2];
    const float* bet = pn + pl.slot_off[NF_P_BE2];
    for (int r = ty; r < m.dt; r += 8) {
      float acc = 0.f;
      for (int k = tx; k < m.C; k += 32) {
        const float w = W[(long long)r * m.C + k];
        put3(packed, kn + kl.W3f + (long long)r * m.C + k, w * gam[k], HI, LO);
        acc = fmaf(w, bet[k], acc);
      }
      acc = warp_sum(acc);
      if (tx == 0) put3(packed, kn + kl.b3f + r, pn[pl.slot_off[NF_P_B3] + r] + acc, HI, LO);
    }
    return;
  }
  job -= 1;
  if (job < J.tW) {                                // W1T[c, r] = W1[r, c]
    const float* W = pn + pl.slot_off[NF_P_W1];
    const int tc = (m.K1 + 31) / 32, r0 = (job / tc) * 32, c0 = (job % tc) * 32;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int r = r0 + ty + 8 * q, c = c0 + tx;
      tile[ty + 8 * q][tx] = (r < m.H1 && c < m.K1) ? W[(long long)r * m.K1 + c] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int c = c0 + ty + 8 * q, r = r0 + tx;
      if (c < m.K1 && r < m.H1) put3(packed, kn + kl.W1T + (long long)c * m.H1 + r, tile[tx][ty + 8 * q], HI, LO);
    }
    return;
  }
  job -= J.tW;
  if (job < J.rP) {                                // W1p rows
    const float* W = pn + pl.slot_off[NF_P_W1];
    const int r = job * 8 + ty;
    if (r >= m.H1) return;
    for (int c = tx; c < kl.K1p; c += 32) {
      float v = 0.f;
      if (c < m.dc) v = W[(long long)r * m.K1 + c];
      else if (c >= kl.dcp && c - kl.dcp < m.R) v = W[(long long)r * m.K1 + m.dc + c - kl.dcp];
      put3(packed, kn + kl.W1p + (long long)r * kl.K1p + c, v, HI, LO);
    }
    return;
  }
  if (net == 0) {
    const int D = m.D, DD = D * D;
    const float* pb = params + (long long)b * pl.block_stride;
    const long long kb = (long long)b * kl.block_stride;
    const float* L = pb + pl.slot_off[NF_P_L];
    const float* U = pb + pl.slot_off[NF_P_U];
    const float* sv = pb + pl.slot_off[NF_P_S];
    const float* P = pb + pl.slot_off[NF_P_P];
    for (int idx = threadIdx.x; idx < DD; idx += blockDim.x) {
      const int i = idx / D, j = idx % D;
      const float lh = i > j ? L[idx] : (i == j ? 1.f : 0.f);
      const float uh = i < j ? U[idx] : (i == j ? sv[i] : 0.f);
      sP[idx] = P[idx]; sLh[idx] = lh; sUh[idx] = uh;
      put3(packed, kb + kl.Lh + idx, lh, HI, LO);
      put3(packed, kb + kl.Uh + idx, uh, HI, LO);
    }
    if (threadIdx.x < 32) {                           // log|det| of this block; block 0 also writes the total over the blocks
      float acc = 0.f;
      for (int i = threadIdx.x; i < D; i += 32) acc += logf(fabsf(sv[i]));
      acc = warp_sum(acc);
      if (threadIdx.x == 0) packed[kl.logdet_c + b] = acc;
      if (b == 0) {
        float tot = 0.f;
        for (int bb = 0; bb < m.n; ++bb) {            // same order and arithmetic as k_plu_mask + k_logdet_total
          const float* sb = params + (long long)bb * pl.block_stride + pl.slot_off[NF_P_S];
          float a = 0.f;
          for (int i = threadIdx.x; i < D; i += 32) a += logf(fabsf(sb[i]));
          a = warp_sum(a);
          tot += a;
        }
        if (threadIdx.x == 0) packed[kl.logdet_total] = tot;
      }
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < DD; idx += blockDim.x) {
      const int i = idx / D, j = idx % D;
      float acc = 0.f;
      for (int k = 0; k < D; ++k) acc = fmaf(sP[i * D + k], sLh[k * D + j], acc);
      sPL[idx] = acc;
      put3(packed, kb + kl.PL + idx, acc, HI, LO);
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < DD; idx += blockDim.x) {
      const int i = idx / D, j = idx % D;
      float acc = 0.f;
      for (int k = 0; k < D; ++k) acc = fmaf(sPL[i * D + k], sUh[k * D + j], acc);
      put3(packed, kb + kl.W + idx, acc, HI, LO);
      put3(packed, kb + kl.WT + j * D + i, acc, HI, LO);
    }
  }
}

// Small flow dimensions: T1 = P^T dW, dLh = T1 Uh^T, dUh = (P Lh)^T dW and the masking / diagonal terms of
// k_plu_grad_finish in one kernel, one CTA per block (three FFMA GEMM launches + the finish kernel before).
__global__ void __launch_bounds__(256) k_plu_grad_small(const float* __restrict__ params, NfParamLayout pl, int D,
                                                        const float* __restrict__ packed, PackedLayout kl,
                                                        const float* __restrict__ dW, const float* __restrict__ sum_dld,
                                                        float* __restrict__ dparams, float ld_sign) {
  NF_PDL_ENTRY();
  __shared__ float sP[1024], sUh[1024], sPL[1024], sdW[1024], sT1[1024];
  const int b = blockIdx.x, DD = D * D;
  const float* pb = params + (long long)b * pl.block_stride;
  const float* kb = packed + (long long)b * kl.block_stride;
  float* db = dparams + (long long)b * pl.block_stride;
  for (int idx = threadIdx.x; idx < DD; idx += blockDim.x) {
    sP[idx] = pb[pl.slot_off[NF_P_P] + idx];
    sUh[idx] = kb[kl.Uh + idx];
    sPL[idx] = kb[kl.PL + idx];
    sdW[idx] = dW[(long long)b * DD + idx];
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < DD; idx += blockDim.x) {     // T1 = P^T dW
    const int i = idx / D, j = idx % D;
    float acc = 0.f;
    for (int k = 0; k < D; ++k) acc = fmaf(sP[k * D + i], sdW[k * D + j], acc);
    sT1[idx] = acc;
  }
  __syncthreads();
  const float sd = sum_dld ? ld_sign * *sum_dld : 0.f;
  for (int idx = threadIdx.x; idx < DD; idx += blockDim.x) {
    const int i = idx / D, j = idx % D;
    float dl = 0.f, du = 0.f;
    if (i > j) for (int k = 0; k < D; ++k) dl = fmaf(sT1[i * D + k], sUh[j * D + k], dl);       // dLh = T1 Uh^T
    if (i <= j) for (int k = 0; k < D; ++k) du = fmaf(sPL[k * D + i], sdW[k * D + j], du);      // dUh = PL^T dW
    db[pl.slot_off[NF_P_L] + idx] = i > j ? dl : 0.f;
    db[pl.slot_off[NF_P_U] + idx] = i < j ? du : 0.f;
    if (i == j) db[pl.slot_off[NF_P_S] + i] = du + sd / pb[pl.slot_off[NF_P_S] + i];
  }
}

// Triangular inverses by substitution against the identity, one thread per column
// (torch.linalg.solve_triangular(..., eye), torch_fcnf.py:49-52).
__global__ void k_tri_inverse(const float* packed, PackedLayout kl, int D, float* uinv, float* linv) {
  NF_PDL_ENTRY();
  const int b = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;  // column of the inverse
  if (j >= D) return;
  const float* Uh = packed + (long long)b * kl.block_stride + kl.Uh;
  const float* Lh = packed + (long long)b * kl.block_stride + kl.Lh;
  float* X = uinv + (long long)b * D * D;
  float* Y = linv + (long long)b * D * D;
  // upper: U X = I  -> back substitution; X[i][j] = 0 for i > j
  for (int i = D - 1; i >= 0; --i) {
    float acc = (i == j) ? 1.f : 0.f;
    if (i <= j) {
      for (int k = i + 1; k <= j; ++k) acc = fmaf(-Uh[i * D + k], X[k * D + j], acc);
      X[i * D + j] = acc / Uh[i * D + i];
    } else {
      X[i * D + j] = 0.f;
    }
  }
  // unit lower: L Y = I -> forward substitution; Y[i][j] = 0 for i < j
  for (int i = 0; i < D; ++i) {
    if (i < j) { Y[i * D + j] = 0.f; continue; }
    float acc = (i == j) ? 1.f : 0.f;
    for (int k = j; k < i; ++k) acc = fmaf(-Lh[i * D + k], Y[k * D + j], acc);
    Y[i * D + j] = acc;
  }
}

// W2f = W2 diag(gamma1), b2f = b2 + W2 beta1 ; W3f = W3 diag(gamma2), b3f = b3 + W3 beta2.
// LayerNorm's affine part commutes into the next Linear, so the forward only stores x_hat.
__global__ void k_fold(const float* params, NfParamLayout pl, Dims m, float* packed, PackedLayout kl) {
  NF_PDL_ENTRY();
  const int b = blockIdx.z, net = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);  // over C + dt output rows
  const float* pn = params + (long long)b * pl.block_stride + net * pl.net_stride;
  float* kn = packed + (long long)b * kl.block_stride + kl.net0 + net * kl.net_stride;
  const float *W, *bias, *gam, *bet;
  float *Wf, *bf;
  int K, r;
  if (row < m.C) {
    r = row; K = m.H1;
    W = pn + pl.slot_off[NF_P_W2]; bias = pn + pl.slot_off[NF_P_B2];
    gam = pn + pl.slot_off[NF_P_G1]; bet = pn + pl.slot_off[NF_P_BE1];
    Wf = kn + kl.W2f; bf = kn + kl.b2f;
  } else if (row < m.C + m.dt) {
    r = row - m.C; K = m.C;
    W = pn + pl.slot_off[NF_P_W3]; bias = pn + pl.slot_off[NF_P_B3];
    gam = pn + pl.slot_off[NF_P_G2]; bet = pn + pl.slot_off[NF_P_BE2];
    Wf = kn + kl.W3f; bf = kn + kl.b3f;
  } else {
    return;
  }
  float acc = 0.f;
  for (int k = lane; k < K; k += 32) {
    const float w = W[(long long)r * K + k];
    Wf[(long long)r * K + k] = w * gam[k];
    acc = fmaf(w, bet[k], acc);
  }
  acc = warp_sum(acc);
  if (lane == 0) bf[r] = bias[r] + acc;
}

// W1p[r][c] = W1[r][c] for the x_cond columns c < dc, W1p[r][dcp + j] = W1[r][dc + j] for the y
// columns, zero elsewhere.
__global__ void k_pack_w1(const float* params, NfParamLayout pl, Dims m, float* packed, PackedLayout kl) {
  NF_PDL_ENTRY();
  const int b = blockIdx.z, net = blockIdx.y, r = blockIdx.x;
  const float* W = params + (long long)b * pl.block_stride + net * pl.net_stride + pl.slot_off[NF_P_W1] +
                   (long long)r * m.K1;
  float* Wp = packed + (long long)b * kl.block_stride + kl.net0 + net * kl.net_stride + kl.W1p + (long long)r * kl.K1p;
  for (int c = threadIdx.x; c < kl.K1p; c += blockDim.x) {
    float v = 0.f;
    if (c < m.dc) v = W[c];
    else if (c >= kl.dcp && c - kl.dcp < m.R) v = W[m.dc + c - kl.dcp];
    Wp[c] = v;
  }
}

// Inverse of the folding for gradients (in place on the gradient arena):
//   slot W{2,3} holds G = du^T x_hat, slot b{2,3} holds g = colsum(du)
//   dW[n,k] = G[n,k] gamma[k] + g[n] beta[k];  dgamma[k] = sum_n G[n,k] W[n,k];  dbeta[k] = sum_n g[n] W[n,k]
constexpr int kUnfoldRows = 16;
__global__ void k_unfold(const float* params, NfParamLayout pl, Dims m, float* dparams) {
  NF_PDL_ENTRY();
  const int layer = blockIdx.z & 1, net = (blockIdx.z >> 1) & 1, b = blockIdx.z >> 2;
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  const int rows = layer == 0 ? m.C : m.dt;
  const int K = layer == 0 ? m.H1 : m.C;
  const int n0 = blockIdx.y * kUnfoldRows;
  if (col >= K || n0 >= rows) return;
  const float* pn = params + (long long)b * pl.block_stride + net * pl.net_stride;
  float* dn = dparams + (long long)b * pl.block_stride + net * pl.net_stride;
  const int sw = layer == 0 ? NF_P_W2 : NF_P_W3, sb = layer == 0 ? NF_P_B2 : NF_P_B3;
  const int sg = layer == 0 ? NF_P_G1 : NF_P_G2, sbe = layer == 0 ? NF_P_BE1 : NF_P_BE2;
  const float* __restrict__ W = pn + pl.slot_off[sw];
  float* __restrict__ G = dn + pl.slot_off[sw];
  const float* __restrict__ g = dn + pl.slot_off[sb];
  const float gam = pn[pl.slot_off[sg] + col], bet = pn[pl.slot_off[sbe] + col];
  float dgam = 0.f, dbet = 0.f;
  const int n1 = min(rows, n0 + kUnfoldRows);
  // all loads of the strip first (the in-place store to G made every iteration wait for the one before it: this
  // kernel is the tail of the backward pass, 12.8 us on C2a for 4 MB of gradients)
  float Gv[kUnfoldRows], wv[kUnfoldRows], gn[kUnfoldRows];
#pragma unroll
  for (int i = 0; i < kUnfoldRows; ++i) {
    const int n = n0 + i;
    const bool ok = n < n1;
    const long long idx = (long long)n * K + col;
    Gv[i] = ok ? G[idx] : 0.f;
    wv[i] = ok ? W[idx] : 0.f;
    gn[i] = ok ? g[n] : 0.f;
  }
#pragma unroll
  for (int i = 0; i < kUnfoldRows; ++i) {
    const int n = n0 + i;
    dgam = fmaf(Gv[i], wv[i], dgam);
    dbet = fmaf(gn[i], wv[i], dbet);
    if (n < n1) G[(long long)n * K + col] = fmaf(Gv[i], gam, gn[i] * bet);
  }
  atomicAdd(&dn[pl.slot_off[sg] + col], dgam);    // the gamma / beta slots are zero on entry
  atomicAdd(&dn[pl.slot_off[sbe] + col], dbet);
}

// dL = tril(dLh,-1), dU = triu(dUh,1), ds = diag(dUh) + sum(dlogdet)/s
__global__ void k_plu_grad_finish(const float* params, NfParamLayout pl, int D, const float* dLh, const float* dUh,
                                  const float* sum_dld, float* dparams, float ld_sign) {
  NF_PDL_ENTRY();
  const int b = blockIdx.y;
  const float* pb = params + (long long)b * pl.block_stride;
  float* db = dparams + (long long)b * pl.block_stride;
  const float* dl = dLh + (long long)b * D * D;
  const float* du = dUh + (long long)b * D * D;
  const float sd = sum_dld ? ld_sign * *sum_dld : 0.f;   // reverse direction: log-det = -sum log|s| per block
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < D * D; idx += gridDim.x * blockDim.x) {
    const int i = idx / D, j = idx % D;
    db[pl.slot_off[NF_P_L] + idx] = i > j ? dl[idx] : 0.f;
    db[pl.slot_off[NF_P_U] + idx] = i < j ? du[idx] : 0.f;
    if (i == j) db[pl.slot_off[NF_P_S] + i] = du[idx] + sd / pb[pl.slot_off[NF_P_S] + i];
  }
}

__global__ void k_sum_vector(const float* v, long long n, float* out) {
  NF_PDL_ENTRY();
  __shared__ float sh[32];
  float acc = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    acc += v[i];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float a = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.f;
    a = warp_sum(a);
    if (threadIdx.x == 0) atomicAdd(out, a);
  }
}

__global__ void k_axpy(float* dst, const float* src, long long n) {
  NF_PDL_ENTRY();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dst[i] += src[i];
}

__global__ void k_copy2d(float* dst, long long ldd, const float* src, long long lds, long long rows, int cols) {
  NF_PDL_ENTRY();
  const long long total = rows * cols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols;
    const int c = (int)(i - r * cols);
    dst[r * ldd + c] = src[r * lds + c];
  }
}

__global__ void k_transpose2d(float* dst, long long ldd, long long sd0, long long sd1, const float* src, long long lds,
                              long long ss0, long long ss1, int rows, int cols, int batch0) {
  NF_PDL_ENTRY();
  __shared__ float tile[32][33];
  const int b0 = blockIdx.z % batch0, b1 = blockIdx.z / batch0;
  const float* s = src + b0 * ss0 + b1 * ss1;
  float* d = dst + b0 * sd0 + b1 * sd1;
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < rows && c < cols) ? s[(long long)r * lds + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (r < rows && c < cols) d[(long long)c * ldd + r] = tile[threadIdx.x][i];
  }
}

}  // namespace

// ------------------------------------------------------------------------------------------
// host wrappers
// ------------------------------------------------------------------------------------------
static bool al16(const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; }

int lrelu_ln_fwd(const LnFwdP& p, cudaStream_t st) {
  if (p.B <= 0) return 0;
  dim3 grid((unsigned)cdiv(p.B, 8), (unsigned)p.nets);
  ProfScope prof(st, PROF_LN_FWD, 8.0 * p.B * p.N * p.nets);   // read u, write x_hat
  const bool v4 = !p.xh_lo && p.N % 4 == 0 && p.N <= 1024 && al16(p.u) && al16(p.xh) && al16(p.bias) &&
                  p.u_net % 4 == 0 && p.xh_net % 4 == 0 && p.bias_net % 4 == 0;
  if (v4) {
    const int nv = (int)cdiv(p.N, 128);
    if (nv <= 1) NF_LAUNCH((k_lrelu_ln_fwd_v4<1>), grid, 256, 0, st, p);
    else if (nv <= 2) NF_LAUNCH((k_lrelu_ln_fwd_v4<2>), grid, 256, 0, st, p);
    else if (nv <= 4) NF_LAUNCH((k_lrelu_ln_fwd_v4<4>), grid, 256, 0, st, p);
    else NF_LAUNCH((k_lrelu_ln_fwd_v4<8>), grid, 256, 0, st, p);
  } else {
    NF_LAUNCH((k_lrelu_ln_fwd), grid, 256, 0, st, p);
  }
  NF_LAUNCHED();
  return 0;
}

int ln_lrelu_bwd(const LnBwdP& p, cudaStream_t st) {
  if (p.B <= 0) return 0;
  // latency-bound: fill the GPU with warps first, then give each warp several rows
  long long rpw = p.B * p.nets / (148 * 24);
  if (rpw < 1) rpw = 1;
  if (rpw > kBwdRowsPerWarp) rpw = kBwdRowsPerWarp;
  dim3 grid((unsigned)cdiv(p.B, kBwdWarps * rpw), (unsigned)p.nets);
  ProfScope prof(st, PROF_LN_BWD, 12.0 * p.B * p.N * p.nets);  // read g, x_hat; write du
  const bool v4 = p.N % 4 == 0 && p.N <= 1024 && al16(p.g) && al16(p.xh) && p.g_net % 4 == 0 && p.xh_net % 4 == 0;
  if (v4) {
    // two CTAs per SM, two waves; at least 32 rows per CTA (one global atomic per column and CTA)
    long long rpc = cdiv(p.B * p.nets, 2 * 2 * 148);
    rpc = std::max<long long>(32, cdiv(rpc, 8) * 8);
    dim3 g2((unsigned)cdiv(p.B, rpc), (unsigned)p.nets);
    const size_t smem = (size_t)8 * p.N * sizeof(float);
    const int nv = (int)cdiv(p.N, 128);
    if (nv <= 1) NF_LAUNCH((k_ln_lrelu_bwd_v4<1>), g2, 256, smem, st, p, (int)rpc);
    else if (nv <= 2) NF_LAUNCH((k_ln_lrelu_bwd_v4<2>), g2, 256, smem, st, p, (int)rpc);
    else if (nv <= 4) NF_LAUNCH((k_ln_lrelu_bwd_v4<4>), g2, 256, smem, st, p, (int)rpc);
    else NF_LAUNCH((k_ln_lrelu_bwd_v4<8>), g2, 256, smem, st, p, (int)rpc);
  } else {
    NF_LAUNCH((k_ln_lrelu_bwd), grid, kBwdWarps * 32, 0, st, p, (int)rpw);
  }
  NF_LAUNCHED();
  return 0;
}

template <bool INV>
static int affine_launch(const float* in, int ldi, const float* s, const float* t, const float* bs, const float* bt,
                         const float* cdev, float chost, long long B, int D, float* out, int ldo, float* logdet,
                         float* s_out, cudaStream_t st) {
  if (B <= 0) return 0;
  // algorithmic bytes: read xp (D), s, t (D/2 each), logdet; write z (D), logdet  
  ProfScope prof(st, PROF_AFFINE, 4.0 * B * (2.0 * D + 2.0 * (D / 2) + 2.0));
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
  if (D % 8 == 0 && ldi % 4 == 0 && ldo % 4 == 0 && al16(in) && al16(out) && al16(s) && al16(t) && al16(s_out) &&
      al16(bs) && al16(bt) && in != out) {
    const int h4 = D / 8;
#define NF_AFF(G)                                                                                                   \
    NF_LAUNCH((k_affine_vec<G, INV>), (unsigned)cdiv(B, 8 * (32 / G)), 256, 0, st, in, s, t, bs, bt, cdev, chost, B, D, ldi, \
              ldo, out, logdet, s_out)
    if (h4 <= 1) NF_AFF(1); else if (h4 <= 2) NF_AFF(2); else if (h4 <= 4) NF_AFF(4); else if (h4 <= 8) NF_AFF(8);
    else if (h4 <= 16) NF_AFF(16); else NF_AFF(32);
#undef NF_AFF
    NF_LAUNCHED();
    return 0;
  }
  if (D <= 64) {
    const bool vec = (D % 8 == 0) && (ldi % 4 == 0) && (ldo % 4 == 0);
    dim3 grid((unsigned)cdiv(B, 256));
    if (vec)
      NF_LAUNCH((k_affine_rowthread<4, INV>), grid, 256, 0, st, in, s, t, bs, bt, cdev, chost, B, D, ldi, ldo, out, logdet, s_out);
    else
      NF_LAUNCH((k_affine_rowthread<1, INV>), grid, 256, 0, st, in, s, t, bs, bt, cdev, chost, B, D, ldi, ldo, out, logdet, s_out);
  } else {
    dim3 grid((unsigned)cdiv(B, 8));
    NF_LAUNCH((k_affine_rowwarp<INV>), grid, 256, 0, st, in, s, t, bs, bt, cdev, chost, B, D, ldi, ldo, out, logdet, s_out);
  }
  NF_LAUNCHED();
  return 0;
}

int affine_fwd(const float* xp, int ld_in, const float* s, const float* t, const float* bias_s, const float* bias_t,
               const float* cdev, float chost, long long B, int D, float* z, int ld_out, float* logdet, float* s_out,
               cudaStream_t st) {
  return affine_launch<false>(xp, ld_in, s, t, bias_s, bias_t, cdev, chost, B, D, z, ld_out, logdet, s_out, st);
}
int affine_inv(const float* z, int ld_in, const float* s, const float* t, const float* bias_s, const float* bias_t,
               const float* cdev, float chost, long long B, int D, float* xp, int ld_out, float* logdet,
               cudaStream_t st) {
  return affine_launch<true>(z, ld_in, s, t, bias_s, bias_t, cdev, chost, B, D, xp, ld_out, logdet, nullptr, st);
}

int affine_bwd(const float* dz, const float* z, const float* s_full, const float* dlogdet, long long B, int D,
               float* dxp, float* ds, float* dt, float* g3s, float* g3t, cudaStream_t st, int inverse) {
  if (B <= 0) return 0;
  const long long total = B * D;
  long long blocks = cdiv(total, 256 * 4);
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  ProfScope prof(st, PROF_AFFINE, 4.0 * B * (3.0 * D + 3.0 * (D / 2) + 1.0));
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
  if (!g3s && !g3t && D % 8 == 0 && al16(dz) && al16(z) && al16(s_full) && al16(dxp) && al16(ds) && al16(dt)) {
    const int h4 = D / 8;
#define NF_AFFB(G) \
    NF_LAUNCH((k_affine_bwd_vec<G>), (unsigned)cdiv(B, 8 * (32 / G)), 256, 0, st, dz, z, s_full, dlogdet, B, D, dxp, ds, dt, inverse)
    if (h4 <= 1) NF_AFFB(1); else if (h4 <= 2) NF_AFFB(2); else if (h4 <= 4) NF_AFFB(4); else if (h4 <= 8) NF_AFFB(8);
    else if (h4 <= 16) NF_AFFB(16); else NF_AFFB(32);
#undef NF_AFFB
    NF_LAUNCHED();
    return 0;
  }
  NF_LAUNCH((k_affine_bwd), (unsigned)blocks, 256, 2 * (D / 2) * sizeof(float) + 16, st, dz, z, s_full, dlogdet, B, D, dxp, ds, dt, g3s, g3t, inverse);
  NF_LAUNCHED();
  return 0;
}

int plu_mask(const float* params, const NfParamLayout& pl, const Dims& m, float* packed, const PackedLayout& kl,
             cudaStream_t st) {
  dim3 grid((unsigned)std::min<int64_t>(cdiv((int64_t)m.D * m.D, 256), 64), (unsigned)m.n);
  NF_LAUNCH((k_plu_mask), grid, 256, 0, st, params, pl, m.D, packed, kl);
  NF_LAUNCHED();
  NF_LAUNCH((k_logdet_total), 1, 32, 0, st, packed, kl, m.n);
  NF_LAUNCHED();
  return 0;
}

int plu_build_small(const float* params, const NfParamLayout& pl, const Dims& m, float* packed, const PackedLayout& kl,
                    cudaStream_t st) {
  NF_CHECK_ARG(m.D <= 32, NF_E_UNSUPPORTED, "plu_build_small: D=%d", m.D);
  NF_LAUNCH((k_plu_build_small), (unsigned)m.n, 256, 0, st, params, pl, m.D, packed, kl);
  NF_LAUNCHED();
  NF_LAUNCH((k_logdet_total), 1, 32, 0, st, packed, kl, m.n);
  NF_LAUNCHED();
  return 0;
}

int prepare_fwd_fused(const float* params, const NfParamLayout& pl, const Dims& m, float* packed, const PackedLayout& kl,
                      cudaStream_t st) {
  NF_CHECK_ARG(m.D <= 32, NF_E_UNSUPPORTED, "prepare_fwd_fused: D=%d", m.D);
  PrepJobs J;
  J.tF = (int)(cdiv(m.C, 32) * cdiv(m.H1, 32)); J.rB = (int)cdiv(m.C, 8);
  J.tW = (int)(cdiv(m.H1, 32) * cdiv(m.K1, 32)); J.rP = (int)cdiv(m.H1, 8);
  dim3 grid((unsigned)(J.tF + J.rB + 1 + J.tW + J.rP + 1), 1, (unsigned)(2 * m.n));
  ProfScope prof(st, PROF_PARAM, 0.0);
  NF_LAUNCH((k_prepare_fwd), grid, 256, 0, st, params, pl, m, packed, kl, J);
  NF_LAUNCHED();
  return 0;
}

int plu_grad_small(const float* params, const NfParamLayout& pl, const Dims& m, const float* packed, const PackedLayout& kl,
                   const float* dW, const float* sum_dld, float* dparams, cudaStream_t st, float ld_sign) {
  NF_CHECK_ARG(m.D <= 32, NF_E_UNSUPPORTED, "plu_grad_small: D=%d", m.D);
  if (m.n <= 0) return 0;
  NF_LAUNCH((k_plu_grad_small), (unsigned)m.n, 256, 0, st, params, pl, m.D, packed, kl, dW, sum_dld, dparams, ld_sign);
  NF_LAUNCHED();
  return 0;
}

int plu_tri_inverse(const float* packed, const PackedLayout& kl, const Dims& m, float* uinv, float* linv,
                    cudaStream_t st) {
  dim3 grid((unsigned)cdiv(m.D, 64), (unsigned)m.n);
  NF_LAUNCH((k_tri_inverse), grid, 64, 0, st, packed, kl, m.D, uinv, linv);
  NF_LAUNCHED();
  return 0;
}

int fold_ln_affine(const float* params, const NfParamLayout& pl, const Dims& m, float* packed, const PackedLayout& kl,
                   cudaStream_t st) {
  dim3 grid((unsigned)cdiv(m.C + m.dt, 8), 2, (unsigned)m.n);
  NF_LAUNCH((k_fold), grid, 256, 0, st, params, pl, m, packed, kl);
  NF_LAUNCHED();
  return 0;
}

int pack_w1(const float* params, const NfParamLayout& pl, const Dims& m, float* packed, const PackedLayout& kl,
            cudaStream_t st) {
  dim3 grid((unsigned)m.H1, 2, (unsigned)m.n);
  NF_LAUNCH((k_pack_w1), grid, 128, 0, st, params, pl, m, packed, kl);
  NF_LAUNCHED();
  return 0;
}

int unfold_grads(const float* params, const NfParamLayout& pl, const Dims& m, float* dparams, cudaStream_t st) {
  const int K = m.H1 > m.C ? m.H1 : m.C;
  const int rows = m.C > m.dt ? m.C : m.dt;
  dim3 grid((unsigned)cdiv(K, 128), (unsigned)cdiv(rows, kUnfoldRows), (unsigned)(4 * m.n));
  ProfScope prof(st, PROF_PARAM, 0.0);
  NF_LAUNCH((k_unfold), grid, 128, 0, st, params, pl, m, dparams);
  NF_LAUNCHED();
  return 0;
}

int plu_grad_finish(const float* params, const NfParamLayout& pl, const Dims& m, const float* dLh, const float* dUh,
                    const float* sum_dld, float* dparams, cudaStream_t st, float ld_sign) {
  dim3 grid((unsigned)std::min<int64_t>(cdiv((int64_t)m.D * m.D, 256), 64), (unsigned)m.n);
  NF_LAUNCH((k_plu_grad_finish), grid, 256, 0, st, params, pl, m.D, dLh, dUh, sum_dld, dparams, ld_sign);
  NF_LAUNCHED();
  return 0;
}

int sum_vector(const float* v, long long n, float* out, cudaStream_t st) {
  NF_CUDA(cudaMemsetAsync(out, 0, sizeof(float), st));
  if (n <= 0) return 0;
  long long blocks = cdiv(n, 256 * 8);
  if (blocks > 296) blocks = 296;
  NF_LAUNCH((k_sum_vector), (unsigned)blocks, 256, 0, st, v, n, out);
  NF_LAUNCHED();
  return 0;
}

int fill_zero(float* p, long long n, cudaStream_t st) {
  if (n <= 0) return 0;
  NF_CUDA(cudaMemsetAsync(p, 0, n * sizeof(float), st));
  return 0;
}

int copy2d(float* dst, long long ldd, const float* src, long long lds, long long rows, int cols, cudaStream_t st) {
  if (rows <= 0 || cols <= 0) return 0;
  long long blocks = cdiv(rows * cols, 256 * 4);
  if (blocks > 148 * 8) blocks = 148 * 8;
  NF_LAUNCH((k_copy2d), (unsigned)blocks, 256, 0, st, dst, ldd, src, lds, rows, cols);
  NF_LAUNCHED();
  return 0;
}

int transpose2d(float* dst, long long ldd, long long sdst0, long long sdst1, const float* src, long long lds,
                long long ssrc0, long long ssrc1, int rows, int cols, int batch0, int batch1, cudaStream_t st) {
  if (rows <= 0 || cols <= 0) return 0;
  dim3 grid((unsigned)cdiv(cols, 32), (unsigned)cdiv(rows, 32), (unsigned)(batch0 * batch1));
  NF_LAUNCH((k_transpose2d), grid, dim3(32, 8), 0, st, dst, ldd, sdst0, sdst1, src, lds, ssrc0, ssrc1, rows, cols, batch0);
  NF_LAUNCHED();
  return 0;
}

int axpy_rows(float* dst, const float* src, long long n, cudaStream_t st) {
  if (n <= 0) return 0;
  long long blocks = cdiv(n, 256 * 4);
  if (blocks > 148 * 8) blocks = 148 * 8;
  NF_LAUNCH((k_axpy), (unsigned)blocks, 256, 0, st, dst, src, n);
  NF_LAUNCHED();
  return 0;
}

namespace {
// ------------------------------------------------------------------------------------------
// fused negative log-likelihood and its cotangents; fused AdamW over the flat arena
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_nll_fwd(const float* __restrict__ z, const float* __restrict__ logdet,
                                                 long long B, int D, float* loss) {
  NF_PDL_ENTRY();
  __shared__ float s_part[8];
  const float kHalfLog2Pi = 0.918938533204672742f;
  float acc = 0.f;
  for (long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x; row < B; row += (long long)gridDim.x * blockDim.x) {
    float sq = 0.f;
    for (int j = 0; j < D; ++j) { const float v = z[row * D + j]; sq = fmaf(v, v, sq); }
    acc += 0.5f * sq + (float)D * kHalfLog2Pi - logdet[row];
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 8) {
    float v = s_part[threadIdx.x];
    for (int o = 4; o > 0; o >>= 1) v += __shfl_xor_sync(0xffu, v, o);
    if (threadIdx.x == 0) atomicAdd(loss, v / (float)B);
  }
}

__global__ void __launch_bounds__(256) k_nll_cot(const float* __restrict__ z, const float* __restrict__ g, long long B,
                                                 int D, float* __restrict__ dz, float* __restrict__ dld) {
  NF_PDL_ENTRY();
  const float scale = (g ? g[0] : 1.0f) / (float)B;
  const long long n = B * D;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    dz[i] = scale * z[i];
    if (i < B) dld[i] = -scale;
  }
}

__global__ void __launch_bounds__(256) k_adamw(AdamWP p, NfParamLayout pl) {
  NF_PDL_ENTRY();
  const int blk = blockIdx.y / NF_P_SLOTS, slot = blockIdx.y % NF_P_SLOTS;
  if (!p.slot_mask[blockIdx.y]) return;
  const long long base = (long long)blk * pl.block_stride + pl.slot_off[slot];
  const long long n = pl.slot_numel[slot];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long e = base + i;
    const float g = p.grads[e];
    float w = p.params[e] * p.lr_wd_factor;                       // param.mul_(1 - lr * weight_decay)
    float m = p.exp_avg[e];
    m = fmaf(p.one_minus_beta1, g - m, m);                        // exp_avg.lerp_(grad, 1 - beta1)
    float v = p.exp_avg_sq[e] * p.beta2;                          // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    v = fmaf(p.one_minus_beta2 * g, g, v);
    const float denom = sqrtf(v) / p.bc2_sqrt + p.eps;
    w -= p.step_size * (m / denom);                               // param.addcdiv_(exp_avg, denom, value=-step_size)
    p.params[e] = w; p.exp_avg[e] = m; p.exp_avg_sq[e] = v;
  }
}
}  // namespace

namespace {
__global__ void __launch_bounds__(256) k_split_tables(const float4* __restrict__ t, float4* __restrict__ hi,
                                                      float4* __restrict__ lo, long long n4) {
  NF_PDL_ENTRY();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = t[i];
    float4 h, l;
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v.x)); h.x = __uint_as_float(r);
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v.y)); h.y = __uint_as_float(r);
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v.z)); h.z = __uint_as_float(r);
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v.w)); h.w = __uint_as_float(r);
    l.x = v.x - h.x; l.y = v.y - h.y; l.z = v.z - h.z; l.w = v.w - h.w;
    hi[i] = h; lo[i] = l;
  }
}
}  // namespace

int split_tables(float* tables, long long hi_off, long long lo_off, long long n, cudaStream_t st) {
  const long long n4 = (n + 3) / 4;
  const long long blocks = std::min<long long>(std::max<long long>(cdiv(n4, 256), 1), 148 * 8);
  ProfScope prof(st, PROF_PARAM, 0.0);
  NF_LAUNCH((k_split_tables), (unsigned)blocks, 256, 0, st, reinterpret_cast<const float4*>(tables),
            reinterpret_cast<float4*>(tables + hi_off), reinterpret_cast<float4*>(tables + lo_off), n4);
  NF_LAUNCHED();
  return 0;
}

int nll_forward(const float* z, const float* logdet, long long B, int D, float* loss, cudaStream_t st) {
  NF_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), st));
  if (B <= 0) return 0;
  const long long blocks = std::min<long long>(cdiv(B, 256), 148 * 4);
  NF_LAUNCH((k_nll_fwd), (unsigned)blocks, 256, 0, st, z, logdet, B, D, loss);
  NF_LAUNCHED();
  return 0;
}

int nll_cotangents(const float* z, const float* g, long long B, int D, float* dz, float* dld, cudaStream_t st) {
  if (B <= 0) return 0;
  const long long blocks = std::min<long long>(cdiv(B * D, 256), 148 * 8);
  NF_LAUNCH((k_nll_cot), (unsigned)blocks, 256, 0, st, z, g, B, D, dz, dld);
  NF_LAUNCHED();
  return 0;
}

// ------------------------------------------------------------------------------------------
// log-density of the standard-normal prior, the `model.prior.log_prob(z)` of the reference's loss line
// (torch_nfbc_ant.py:146,269: MultivariateNormal(0, I), evaluated there through a Cholesky / Mahalanobis path of ~10 kernels):
//   lp[b] = -0.5 |z_b|^2 - D/2 log(2 pi);   backward  dz[b, j] = -g[b] z[b, j]
// One warp per row group: G = 2^k lanes share a row (D / 4 chunks when D % 4 == 0, scalar otherwise), like k_affine_vec.
// ------------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(256) k_stdnormal_logprob(const float* __restrict__ z, long long B, int D,
                                                           float* __restrict__ lp) {
  NF_PDL_ENTRY();
  const int lane = threadIdx.x & 31;
  const long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const float c = -0.5f * (float)D * 1.8378770664093453f;
  if (D <= 32) {                      // a warp takes 32 / G rows, G lanes per row
    int G = 1;
    while (G < D) G <<= 1;
    const int rpw = 32 / G, g = lane % G;
    const long long row = w * rpw + lane / G;
    float s = 0.f;
    if (row < B && g < D) { const float v = z[row * D + g]; s = v * v; }
    for (int o = G >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (row < B && g == 0) lp[row] = fmaf(-0.5f, s, c);
  } else {                            // warp per row
    if (w >= B) return;
    float s = 0.f;
    for (int j = lane; j < D; j += 32) { const float v = z[w * D + j]; s = fmaf(v, v, s); }
    s = warp_sum(s);
    if (lane == 0) lp[w] = fmaf(-0.5f, s, c);
  }
}
__global__ void __launch_bounds__(256) k_stdnormal_logprob_bwd(const float* __restrict__ z, const float* __restrict__ g,
                                                               long long B, int D, float* __restrict__ dz) {
  NF_PDL_ENTRY();
  const long long n = B * D;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dz[i] = -g[i / D] * z[i];
}
}  // namespace

int stdnormal_logprob(const float* z, long long B, int D, float* lp, cudaStream_t st) {
  if (B <= 0) return 0;
  int G = 1;
  while (G < D && G < 32) G <<= 1;
  const long long rows_per_warp = D <= 32 ? 32 / G : 1;
  NF_LAUNCH((k_stdnormal_logprob), (unsigned)cdiv(B, 8 * rows_per_warp), 256, 0, st, z, B, D, lp);
  NF_LAUNCHED();
  return 0;
}
int stdnormal_logprob_bwd(const float* z, const float* g, long long B, int D, float* dz, cudaStream_t st) {
  if (B <= 0) return 0;
  NF_LAUNCH((k_stdnormal_logprob_bwd), (unsigned)std::min<long long>(cdiv(B * D, 256), 148 * 8), 256, 0, st, z, g, B, D, dz);
  NF_LAUNCHED();
  return 0;
}

// ------------------------------------------------------------------------------------------
// Group-wise argmin of the log-probability (unsupservised-gcrl/train_auto_NF_rl.py:328-333: for every environment the
// candidate goal with the LOWEST log N(z) + logdet among its N candidates is the next exploration goal): one CTA per group
// computes lp[n] = -0.5 |z_n|^2 - D/2 log(2 pi) + logdet[n], reduces (value, index) with ties to the smaller index
// (jnp.argmin), and copies the winning candidate row.  Replaces log_prob -> argmin -> gather (and the (G, N) round trip).
// ------------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(256) k_group_argmin_logp(const float* __restrict__ z, const float* __restrict__ logdet,
                                                           const float* __restrict__ x, long long N, int D,
                                                           float* __restrict__ out_x, long long* __restrict__ out_idx,
                                                           float* __restrict__ out_lp) {
  NF_PDL_ENTRY();
  __shared__ float s_v[8];
  __shared__ long long s_i[8];
  const long long g = blockIdx.x;
  const float c = -0.5f * (float)D * 1.8378770664093453f;
  float best = INFINITY;
  long long bi = 0x7fffffffffffffffll;
  for (long long n = threadIdx.x; n < N; n += blockDim.x) {
    const float* zr = z + (g * N + n) * D;
    float ss = 0.f;
    for (int j = 0; j < D; ++j) ss = fmaf(zr[j], zr[j], ss);
    const float lp = fmaf(-0.5f, ss, c) + logdet[g * N + n];
    if (lp < best || (lp == best && n < bi)) { best = lp; bi = n; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov < best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
  if ((threadIdx.x & 31) == 0) { s_v[threadIdx.x >> 5] = best; s_i[threadIdx.x >> 5] = bi; }
  __syncthreads();
  if (threadIdx.x < 32) {
    best = threadIdx.x < (blockDim.x >> 5) ? s_v[threadIdx.x] : INFINITY;
    bi = threadIdx.x < (blockDim.x >> 5) ? s_i[threadIdx.x] : 0x7fffffffffffffffll;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov < best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if (threadIdx.x == 0) { s_v[0] = best; s_i[0] = bi; }
  }
  __syncthreads();
  bi = s_i[0];
  if (bi >= N) bi = 0;      // every candidate NaN / +inf: fall back to the first one
  if (threadIdx.x == 0) {
    if (out_idx) out_idx[g] = bi;
    if (out_lp) out_lp[g] = s_v[0];
  }
  if (out_x && x)
    for (int j = threadIdx.x; j < D; j += blockDim.x) out_x[g * D + j] = x[(g * N + bi) * D + j];
}
}  // namespace

int group_argmin_logp(const float* z, const float* logdet, const float* x, long long G, long long N, int D, float* out_x,
                      long long* out_idx, float* out_lp, cudaStream_t st) {
  if (G <= 0 || N <= 0) return 0;
  NF_LAUNCH((k_group_argmin_logp), (unsigned)G, 256, 0, st, z, logdet, x, N, D, out_x, out_idx, out_lp);
  NF_LAUNCHED();
  return 0;
}

// ------------------------------------------------------------------------------------------
// noise augmentation  out = x + std * N(0, 1)   (torch_nfbc_ant.py:264-265: naction + 0.1 * randn_like(naction))
// ------------------------------------------------------------------------------------------
// Counter-based Philox4x32-10 (Salmon et al. 2011; the generator behind torch.randn on CUDA): element group g = i / 4
// uses counter (offset + g) and key `seed`, so the stream is a pure function of (seed, offset, index) -- independent of
// the grid -- and a training loop advances `offset` by the number of groups it consumed.  Four uniforms -> two
// Box-Muller pairs -> four normals, one 128-bit load and store per thread.
namespace {
__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
    const uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
    c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
    k0 += W0; k1 += W1;
  }
}
__device__ __forceinline__ float u01(uint32_t r) { return (float)(r >> 8) * (1.0f / 16777216.0f) + (0.5f / 16777216.0f); }

__global__ void __launch_bounds__(256) k_noise_augment(const float* __restrict__ x, long long n, float std_,
                                                       unsigned long long seed, unsigned long long offset,
                                                       float* __restrict__ out) {
  NF_PDL_ENTRY();
  const long long groups = (n + 3) >> 2;
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += (long long)gridDim.x * blockDim.x) {
    const unsigned long long ctr = offset + (unsigned long long)g;
    uint32_t c[4] = {(uint32_t)ctr, (uint32_t)(ctr >> 32), 0u, 0u};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    float z[4];
    {
      const float r0 = sqrtf(-2.0f * __logf(u01(c[0]))), r1 = sqrtf(-2.0f * __logf(u01(c[2])));
      float s0, c0, s1, c1;
      __sincosf(6.2831853071795865f * u01(c[1]), &s0, &c0);
      __sincosf(6.2831853071795865f * u01(c[3]), &s1, &c1);
      z[0] = r0 * c0; z[1] = r0 * s0; z[2] = r1 * c1; z[3] = r1 * s1;
    }
    const long long i = g << 2;
    if (i + 3 < n) {
      const float4 v = *reinterpret_cast<const float4*>(x + i);
      *reinterpret_cast<float4*>(out + i) = make_float4(fmaf(std_, z[0], v.x), fmaf(std_, z[1], v.y), fmaf(std_, z[2], v.z),
                                                        fmaf(std_, z[3], v.w));
    } else {
      for (int k = 0; k < 4 && i + k < n; ++k) out[i + k] = fmaf(std_, z[k], x[i + k]);
    }
  }
}
}  // namespace

int noise_augment(const float* x, long long n, float std_, unsigned long long seed, unsigned long long offset, float* out,
                  cudaStream_t st) {
  if (n <= 0) return 0;
  const long long blocks = std::min<long long>(cdiv(cdiv(n, 4), 256), 148 * 8);
  ProfScope prof(st, PROF_AFFINE, 8.0 * n);
  NF_LAUNCH((k_noise_augment), (unsigned)blocks, 256, 0, st, x, n, std_, seed, offset, out);
  NF_LAUNCHED();
  return 0;
}

int adamw_step(const AdamWP& p, const NfParamLayout& pl, int n_blocks, cudaStream_t st) {
  long long mx = 0;
  for (int s = 0; s < NF_P_SLOTS; ++s) mx = std::max<long long>(mx, pl.slot_numel[s]);
  const long long bx = std::min<long long>(std::max<long long>(cdiv(mx, 256 * 4), 1), 64);
  NF_LAUNCH((k_adamw), dim3((unsigned)bx, (unsigned)(n_blocks * NF_P_SLOTS)), 256, 0, st, p, pl);
  NF_LAUNCHED();
  return 0;
}

namespace {
__global__ void __launch_bounds__(256) k_copy_segments(CopySegs c) {
  NF_PDL_ENTRY();
  const int seg = blockIdx.y;
  const long long n = c.n[seg];
  const float* __restrict__ src = c.src[seg];
  float* __restrict__ dst = c.dst[seg];
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15u) == 0) {
    const long long n4 = n >> 2;
    for (long long k = i; k < n4; k += stride) reinterpret_cast<float4*>(dst)[k] = __ldg(reinterpret_cast<const float4*>(src) + k);
    for (long long k = (n4 << 2) + i; k < n; k += stride) dst[k] = src[k];
  } else {
    for (long long k = i; k < n; k += stride) dst[k] = src[k];
  }
}
}  // namespace

int copy_segments(const CopySegs& c, cudaStream_t st) {
  if (c.count == 0) return 0;
  long long mx = 0;
  for (int i = 0; i < c.count; ++i) mx = std::max(mx, c.n[i]);
  const long long blocks = std::min<long long>(std::max<long long>(cdiv(mx, 4 * 256), 1), 148 * 4);
  NF_LAUNCH((k_copy_segments), dim3((unsigned)blocks, (unsigned)c.count), 256, 0, st, c);
  NF_LAUNCHED();
  return 0;
}

}  // namespace nf
