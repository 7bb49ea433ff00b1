// Row-fused small-D kernels (see rowfused.cuh).  One warp owns a row at a time; lane j holds
// element j of the D-wide flow vectors and columns lane + 32k of the C-wide activations.
#include "rowfused.cuh"

#include <algorithm>

#include "prof.cuh"

namespace nf {
namespace {

constexpr int kWarps = 8;          // warps per CTA
constexpr int kChunkK = 8;         // 8 x 32 = 256 activation columns per register-accumulator chunk

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---------------------------------------------------------------------------------------------
// forward / inverse tail: o = x_hat2 W3f^T + b3f ; affine coupling ; log-det ; neighbour PLU matmul
// (torch_fcnf.py:77/88 last Linear, :101-104 forward, :112-115 inverse, :40 / :55 PLU matmul)
// ---------------------------------------------------------------------------------------------
template <int MAXDT>
__global__ void __launch_bounds__(kWarps * 32) k_row_tail(TailP p) {
  NF_PDL_ENTRY();
  __shared__ float sW[kRowFusedMaxD * kRowFusedMaxD];
  const int D = p.D, C = p.C, dc = (D + 1) / 2, dt = D / 2;
  if (p.Wmat)
    for (int i = threadIdx.x; i < D * D; i += blockDim.x) sW[i] = p.Wmat[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long row = (long long)blockIdx.x * kWarps + warp;
  if (row >= p.B) return;

  const float* xt = p.xh2 + row * C;               // net 0 = t
  const float* xs = p.xh2 + p.xh2_net + row * C;   // net 1 = s
  const float* Wt = p.W3f;
  const float* Ws = p.W3f + p.w_net;
  float at[MAXDT], as[MAXDT];
#pragma unroll
  for (int i = 0; i < MAXDT; ++i) { at[i] = 0.f; as[i] = 0.f; }
  for (int c = lane; c < C; c += 32) {
    const float vt = xt[c], vs = xs[c];
#pragma unroll
    for (int i = 0; i < MAXDT; ++i)
      if (i < dt) {
        at[i] = fmaf(vt, __ldg(Wt + (long long)i * C + c), at[i]);
        as[i] = fmaf(vs, __ldg(Ws + (long long)i * C + c), as[i]);
      }
  }
  float myS = 0.f, myT = 0.f;   // lane dc + i keeps s_i, t_i
#pragma unroll
  for (int i = 0; i < MAXDT; ++i)
    if (i < dt) {
      const float st = warp_sum(at[i]), ss = warp_sum(as[i]);
      if (lane == dc + i) { myT = st + p.b3f[i]; myS = ss + p.b3f[p.b_net + i]; }
    }
  const float xv = lane < D ? p.in[row * p.ld_in + lane] : 0.f;
  float zv = xv;
  if (lane >= dc && lane < D) zv = p.inverse ? fmaf(xv, expf(myS), myT) : (xv - myT) * expf(-myS);
  const float ssum = warp_sum((lane >= dc && lane < D) ? myS : 0.f);
  if (!p.inverse) {
    if (lane < D) p.out[row * p.ld_out + lane] = zv;
    if (p.s_out && lane >= dc && lane < D) p.s_out[row * dt + (lane - dc)] = myS;
  }
  if (p.logdet && lane == 0) {
    const float c = p.cdev ? *p.cdev : 0.f;
    p.logdet[row] += p.inverse ? (ssum - c) : (c - ssum);
  }
  if (p.Wmat) {
    float o = 0.f;
    for (int j = 0; j < D; ++j) {
      const float zj = __shfl_sync(0xffffffffu, zv, j);
      if (lane < D) o = fmaf(zj, sW[j * D + lane], o);
    }
    if (lane < D) p.out2[row * p.ld_out2 + lane] = o;
  }
}

// ---------------------------------------------------------------------------------------------
// backward head
// ---------------------------------------------------------------------------------------------
// Per net, per 256-column chunk, every warp accumulates over its rows in registers
//   cs[k]      = sum_rows du2[row][c]                      (bias gradient of layer 2)
//   G[i][k]    = sum_rows dout[row][i] * x_hat2[row][c]    (folded last-Linear weight gradient)
// for c = chunk*256 + lane + 32k; the 8 warps of the CTA are then summed through shared memory and
// added to global memory with one atomic per element and CTA.
template <int MAXDT>
__global__ void __launch_bounds__(kWarps * 32) k_row_bwd_head(BwdHeadP p, int rows_per_warp) {
  NF_PDL_ENTRY();
  extern __shared__ float smem[];   // [kWarps][(1 + MAXDT) * 256] reduction slices, then [kWarps][rows_per_warp][2] stats
  const int D = p.D, C = p.C, dc = (D + 1) / 2, dt = D / 2;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int SLICE = (1 + MAXDT) * kChunkK * 32;
  float* slice = smem + warp * SLICE;
  float* stats = smem + kWarps * SLICE + warp * rows_per_warp * 2;
  const long long row0 = ((long long)blockIdx.x * kWarps + warp) * rows_per_warp;
  const bool want_param = p.G3 != nullptr;
  float acc_g3t = 0.f, acc_g3s = 0.f;   // lane dc + i: column sums of dt_i, ds_i over this warp's rows

  for (int net = 0; net < 2; ++net) {
    const float* W3 = p.W3f + net * p.w_net;
    // ---- stats pass (and, for net 0, the affine backward itself)
    for (int r = 0; r < rows_per_warp; ++r) {
      const long long row = row0 + r;
      if (row >= p.B) break;
      const float g = (p.dz && lane < D) ? p.dz[row * p.ld_dz + lane] : 0.f;
      const float zv = lane < D ? p.z[row * D + lane] : 0.f;
      float ds_ = 0.f, dt_ = 0.f, dxpv = g;
      if (lane >= dc && lane < D) {
        const float e = expf(-p.s[row * dt + (lane - dc)]);
        dxpv = g * e;
        dt_ = -dxpv;
        ds_ = -g * zv - (p.dld ? p.dld[row] : 0.f);
      }
      if (net == 0) {
        if (lane < D) p.dxp[row * D + lane] = dxpv;
        acc_g3t += dt_;
        acc_g3s += ds_;
      }
      const float mine = net == 0 ? dt_ : ds_;
      float dout[MAXDT];
#pragma unroll
      for (int i = 0; i < MAXDT; ++i) dout[i] = i < dt ? __shfl_sync(0xffffffffu, mine, dc + i) : 0.f;
      const float* xh = p.xh2 + net * p.xh2_net + row * C;
      float s1 = 0.f, s2 = 0.f;
      for (int c = lane; c < C; c += 32) {
        float gc = 0.f;
#pragma unroll
        for (int i = 0; i < MAXDT; ++i)
          if (i < dt) gc = fmaf(dout[i], __ldg(W3 + (long long)i * C + c), gc);
        s1 += gc;
        s2 = fmaf(gc, xh[c], s2);
      }
      s1 = warp_sum(s1) / (float)C;
      s2 = warp_sum(s2) / (float)C;
      if (lane == 0) { stats[r * 2] = s1; stats[r * 2 + 1] = s2; }
    }
    __syncwarp();
    // ---- chunk passes: du2 and the register accumulators
    for (int ch = 0; ch * kChunkK * 32 < C; ++ch) {
      float cs[kChunkK], G[MAXDT][kChunkK];
#pragma unroll
      for (int k = 0; k < kChunkK; ++k) {
        cs[k] = 0.f;
#pragma unroll
        for (int i = 0; i < MAXDT; ++i) G[i][k] = 0.f;
      }
      for (int r = 0; r < rows_per_warp; ++r) {
        const long long row = row0 + r;
        if (row >= p.B) break;
        const float g = (p.dz && lane < D) ? p.dz[row * p.ld_dz + lane] : 0.f;
        float mine = 0.f;
        if (lane >= dc && lane < D) {
          if (net == 0) mine = -g * expf(-p.s[row * dt + (lane - dc)]);
          else mine = -g * p.z[row * D + lane] - (p.dld ? p.dld[row] : 0.f);
        }
        float dout[MAXDT];
#pragma unroll
        for (int i = 0; i < MAXDT; ++i) dout[i] = i < dt ? __shfl_sync(0xffffffffu, mine, dc + i) : 0.f;
        const float* xh = p.xh2 + net * p.xh2_net + row * C;
        const uint32_t* mask = p.mask + net * p.mask_net + row * p.mw;
        float* du = p.du2 + net * p.du2_net + row * C;
        const float rs = p.rstd[net * p.rstd_net + row];
        const float m1 = stats[r * 2], m2 = stats[r * 2 + 1];
#pragma unroll
        for (int k = 0; k < kChunkK; ++k) {
          const int c = (ch * kChunkK + k) * 32 + lane;
          if (c < C) {
            float gc = 0.f;
#pragma unroll
            for (int i = 0; i < MAXDT; ++i)
              if (i < dt) gc = fmaf(dout[i], __ldg(W3 + (long long)i * C + c), gc);
            const float x = xh[c];
            float dv = rs * (gc - m1 - x * m2);
            dv *= ((mask[c >> 5] >> lane) & 1u) ? 1.0f : kLeakySlope;
            du[c] = dv;
            cs[k] += dv;
#pragma unroll
            for (int i = 0; i < MAXDT; ++i)
              if (i < dt) G[i][k] = fmaf(dout[i], x, G[i][k]);
          }
        }
      }
      if (want_param) {   // uniform across the CTA
#pragma unroll
        for (int k = 0; k < kChunkK; ++k) {
          slice[k * 32 + lane] = cs[k];
#pragma unroll
          for (int i = 0; i < MAXDT; ++i) slice[(1 + i) * kChunkK * 32 + k * 32 + lane] = G[i][k];
        }
        __syncthreads();
        for (int e = threadIdx.x; e < (1 + dt) * kChunkK * 32; e += blockDim.x) {
          float v = 0.f;
#pragma unroll
          for (int w = 0; w < kWarps; ++w) v += smem[w * SLICE + e];
          const int which = e / (kChunkK * 32), c = ch * kChunkK * 32 + e % (kChunkK * 32);
          if (c < C && v != 0.f) {
            if (which == 0) atomicAdd(&p.gb2[net * p.g_net + c], v);
            else atomicAdd(&p.G3[net * p.g_net + (long long)(which - 1) * C + c], v);
          }
        }
        __syncthreads();
      }
    }
  }
  if (want_param) {   // last-Linear bias gradients: one atomic per CTA and element
    slice[lane] = acc_g3t;
    slice[32 + lane] = acc_g3s;
    __syncthreads();
    if (threadIdx.x < 64) {
      const int which = threadIdx.x >> 5, j = threadIdx.x & 31;   // 0: t, 1: s
      if (j >= dc && j < D) {
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) v += smem[w * SLICE + which * 32 + j];
        atomicAdd(&p.g3[which * p.g_net + (j - dc)], v);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// backward tail
// ---------------------------------------------------------------------------------------------
template <int MAXDC>
__global__ void __launch_bounds__(kWarps * 32) k_row_bwd_tail(BwdTailP p, int rows_per_warp) {
  NF_PDL_ENTRY();
  extern __shared__ float smem[];   // [kWarps][max(D*D, MAXDC*256)] reduction slices, then sW[D*D]
  const int D = p.D, H1 = p.H1, dc = (D + 1) / 2;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int SLICE = MAXDC * kChunkK * 32 > kRowFusedMaxD * kRowFusedMaxD ? MAXDC * kChunkK * 32
                                                                             : kRowFusedMaxD * kRowFusedMaxD;
  float* slice = smem + warp * SLICE;
  float* sW = smem + kWarps * SLICE;
  for (int i = threadIdx.x; i < D * D; i += blockDim.x) sW[i] = p.W[i];
  __syncthreads();
  const long long row0 = ((long long)blockIdx.x * kWarps + warp) * rows_per_warp;

  // ---- phase A: dh0 (x_cond columns), dxp, dx_in = dxp W^T, dW += x_in^T dxp
  float accW[kRowFusedMaxD];   // lane b keeps column b of x_in^T dxp
#pragma unroll
  for (int a = 0; a < kRowFusedMaxD; ++a) accW[a] = 0.f;
  for (int r = 0; r < rows_per_warp; ++r) {
    const long long row = row0 + r;
    if (row >= p.B) break;
    float dh[MAXDC];
#pragma unroll
    for (int j = 0; j < MAXDC; ++j) dh[j] = 0.f;
    for (int net = 0; net < 2; ++net) {
      const float* du = p.du1 + net * p.du1_net + row * H1;
      const float* WT = p.W1T + net * p.w_net;
      for (int h = lane; h < H1; h += 32) {
        const float v = du[h];
#pragma unroll
        for (int j = 0; j < MAXDC; ++j)
          if (j < dc) dh[j] = fmaf(v, __ldg(WT + (long long)j * H1 + h), dh[j]);
      }
    }
    float add = 0.f;
#pragma unroll
    for (int j = 0; j < MAXDC; ++j)
      if (j < dc) {
        const float t = warp_sum(dh[j]);
        if (lane == j) add = t;
      }
    float dxpv = 0.f, xinv = 0.f;
    if (lane < D) {
      dxpv = p.dxp[row * D + lane] + add;
      p.dxp[row * D + lane] = dxpv;
      xinv = p.xin[row * D + lane];
    }
    if (p.dxin) {
      float o = 0.f;
      for (int j = 0; j < D; ++j) {
        const float v = __shfl_sync(0xffffffffu, dxpv, j);
        if (lane < D) o = fmaf(v, sW[lane * D + j], o);
      }
      if (lane < D) p.dxin[row * p.ld_dxin + lane] = o;
    }
    if (p.dW) {
#pragma unroll
      for (int a = 0; a < kRowFusedMaxD; ++a)
        if (a < D) accW[a] = fmaf(__shfl_sync(0xffffffffu, xinv, a), dxpv, accW[a]);
    }
  }
  if (p.dW) {
#pragma unroll
    for (int a = 0; a < kRowFusedMaxD; ++a)
      if (a < D && lane < D) slice[a * D + lane] = accW[a];
    __syncthreads();
    for (int e = threadIdx.x; e < D * D; e += blockDim.x) {
      float v = 0.f;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) v += smem[w * SLICE + e];
      if (v != 0.f) atomicAdd(&p.dW[e], v);
    }
    __syncthreads();
  }

  // ---- phase B: G1[:, :dc] += du1^T x_cond, per net and 256-column chunk
  if (p.G1) {
    for (int net = 0; net < 2; ++net) {
      for (int ch = 0; ch * kChunkK * 32 < H1; ++ch) {
        float G[MAXDC][kChunkK];
#pragma unroll
        for (int j = 0; j < MAXDC; ++j)
#pragma unroll
          for (int k = 0; k < kChunkK; ++k) G[j][k] = 0.f;
        for (int r = 0; r < rows_per_warp; ++r) {
          const long long row = row0 + r;
          if (row >= p.B) break;
          const float zc = lane < dc ? p.z[row * D + lane] : 0.f;
          float xc[MAXDC];
#pragma unroll
          for (int j = 0; j < MAXDC; ++j) xc[j] = j < dc ? __shfl_sync(0xffffffffu, zc, j) : 0.f;
          const float* du = p.du1 + net * p.du1_net + row * H1;
#pragma unroll
          for (int k = 0; k < kChunkK; ++k) {
            const int h = (ch * kChunkK + k) * 32 + lane;
            if (h < H1) {
              const float v = du[h];
#pragma unroll
              for (int j = 0; j < MAXDC; ++j)
                if (j < dc) G[j][k] = fmaf(v, xc[j], G[j][k]);
            }
          }
        }
#pragma unroll
        for (int j = 0; j < MAXDC; ++j)
#pragma unroll
          for (int k = 0; k < kChunkK; ++k) slice[j * kChunkK * 32 + k * 32 + lane] = G[j][k];
        __syncthreads();
        for (int e = threadIdx.x; e < dc * kChunkK * 32; e += blockDim.x) {
          float v = 0.f;
#pragma unroll
          for (int w = 0; w < kWarps; ++w) v += smem[w * SLICE + e];
          const int j = e / (kChunkK * 32), h = ch * kChunkK * 32 + e % (kChunkK * 32);
          if (h < H1 && v != 0.f) atomicAdd(&p.G1[net * p.g_net + (long long)h * p.ldg + j], v);
        }
        __syncthreads();
      }
    }
  }
}

int rows_per_warp_for(long long B) {
  // these kernels are latency-bound: first fill the GPU with warps (>= 24 per SM), only then give a
  // warp several rows to amortise the cross-warp reduction
  long long rpw = B / (148 * 24);
  if (rpw < 1) rpw = 1;
  if (rpw > 32) rpw = 32;
  return (int)rpw;
}

}  // namespace

int row_tail(const TailP& p, cudaStream_t st) {
  if (p.B <= 0) return 0;
  NF_CHECK_ARG(p.D >= 2 && p.D <= kRowFusedMaxD, NF_E_UNSUPPORTED, "row_tail: D=%d", p.D);
  const int dt = p.D / 2;
  dim3 grid((unsigned)cdiv(p.B, kWarps));
  ProfScope prof(st, PROF_ROW_TAIL, 4.0 * p.B * (2.0 * p.C + 3.0 * p.D));
  if (dt <= 4) NF_LAUNCH((k_row_tail<4>), grid, kWarps * 32, 0, st, p);
  else NF_LAUNCH((k_row_tail<16>), grid, kWarps * 32, 0, st, p);
  NF_LAUNCHED();
  return 0;
}

size_t row_bwd_head_smem(int D, int C) {
  const int maxdt = D / 2 <= 4 ? 4 : 16;
  return (size_t)kWarps * (1 + maxdt) * kChunkK * 32 * 4 + (size_t)kWarps * 32 * 2 * 4;
}

int row_bwd_head(const BwdHeadP& p, cudaStream_t st) {
  if (p.B <= 0) return 0;
  NF_CHECK_ARG(p.D >= 2 && p.D <= kRowFusedMaxD, NF_E_UNSUPPORTED, "row_bwd_head: D=%d", p.D);
  const int dt = p.D / 2;
  const int rpw = rows_per_warp_for(p.B);
  dim3 grid((unsigned)cdiv(p.B, (long long)kWarps * rpw));
  const size_t smem = row_bwd_head_smem(p.D, p.C);
  ProfScope prof(st, PROF_ROW_HEAD, 4.0 * p.B * (4.0 * p.C + 4.0 * p.D));
  if (dt <= 4) {
    static PerDeviceOnce once;
    if (once.first()) { NF_CUDA(cudaFuncSetAttribute(k_row_bwd_head<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRowFusedSmemLimit)); }
    NF_LAUNCH((k_row_bwd_head<4>), grid, kWarps * 32, smem, st, p, rpw);
  } else {
    static PerDeviceOnce once;
    if (once.first()) { NF_CUDA(cudaFuncSetAttribute(k_row_bwd_head<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRowFusedSmemLimit)); }
    NF_LAUNCH((k_row_bwd_head<16>), grid, kWarps * 32, smem, st, p, rpw);
  }
  NF_LAUNCHED();
  return 0;
}

size_t row_bwd_tail_smem(int D, int H1) {
  const int maxdc = (D + 1) / 2 <= 8 ? 8 : 16;
  const size_t slice = (size_t)std::max(maxdc * kChunkK * 32, kRowFusedMaxD * kRowFusedMaxD);
  return (kWarps * slice + kRowFusedMaxD * kRowFusedMaxD) * 4;
}

int row_bwd_tail(const BwdTailP& p, cudaStream_t st) {
  if (p.B <= 0) return 0;
  NF_CHECK_ARG(p.D >= 2 && p.D <= kRowFusedMaxD, NF_E_UNSUPPORTED, "row_bwd_tail: D=%d", p.D);
  const int dc = (p.D + 1) / 2;
  const int rpw = rows_per_warp_for(p.B);
  dim3 grid((unsigned)cdiv(p.B, (long long)kWarps * rpw));
  const size_t smem = row_bwd_tail_smem(p.D, p.H1);
  ProfScope prof(st, PROF_ROW_BTAIL, 4.0 * p.B * (4.0 * p.H1 + 4.0 * p.D));
  if (dc <= 8) {
    static PerDeviceOnce once;
    if (once.first()) { NF_CUDA(cudaFuncSetAttribute(k_row_bwd_tail<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRowFusedSmemLimit)); }
    NF_LAUNCH((k_row_bwd_tail<8>), grid, kWarps * 32, smem, st, p, rpw);
  } else {
    static PerDeviceOnce once;
    if (once.first()) { NF_CUDA(cudaFuncSetAttribute(k_row_bwd_tail<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRowFusedSmemLimit)); }
    NF_LAUNCH((k_row_bwd_tail<16>), grid, kWarps * 32, smem, st, p, rpw);
  }
  NF_LAUNCHED();
  return 0;
}

}  // namespace nf
