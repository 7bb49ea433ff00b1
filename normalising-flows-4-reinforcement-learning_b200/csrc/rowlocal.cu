// Row-local backward kernels and the multi-job column reduction (see rowlocal.cuh).
#include "rowlocal.cuh"
#include "rowfused.cuh"

#include <algorithm>

#include "prof.cuh"

namespace nf {
namespace {

constexpr int kWarps = 8;
constexpr int kMaxWSmem = 40 * 1024;   // stage the skinny weight matrix in shared memory up to this size

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float dot4(const float4 a, const float4 b) {
  return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w)));
}

// ---------------------------------------------------------------------------------------------
// backward head: one warp per row, lane l owns columns 4*(l + 32k) .. +3 of the C-wide vectors
// ---------------------------------------------------------------------------------------------
template <int NV>
__global__ void __launch_bounds__(kWarps * 32) k_row_bwd_head2(BwdHead2P p, int w_smem, int rpw) {
  NF_PDL_ENTRY();
  extern __shared__ float4 sW4[];   // [2][dt][C/4] when w_smem
  const int D = p.D, C = p.C, dc = (D + 1) / 2, dt = D / 2, C4 = C >> 2;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (w_smem) {
    for (int e = threadIdx.x; e < 2 * dt * C4; e += blockDim.x) {
      const int net = e / (dt * C4), rem = e - net * dt * C4;
      sW4[e] = __ldg(reinterpret_cast<const float4*>(p.W3f + net * p.w_net) + rem);
    }
    __syncthreads();
  }
  const float invC = 1.0f / (float)C;
  for (int r = 0; r < rpw; ++r) {
    const long long row = ((long long)blockIdx.x * kWarps + warp) * rpw + r;
    if (row >= p.B) return;
    // ---- every global load of this row is issued before the first use: one exposed L2/DRAM round
    // trip per row instead of four dependent ones (ncu: long-scoreboard stalls dominated the kernel)
    const float g = (p.dz && lane < D) ? p.dz[row * p.ld_dz + lane] : 0.f;
    const float zv = lane < D ? p.z[row * D + lane] : 0.f;
    const bool is_t = lane >= dc && lane < D;
    const float sv = is_t ? p.s[row * dt + (lane - dc)] : 0.f;
    const float dldv = p.dld ? p.dld[row] : 0.f;
    float4 xv[2][NV];
    uint32_t mwv[2][NV];
    float rs[2];
#pragma unroll
    for (int net = 0; net < 2; ++net) {
      const float4* xh4 = reinterpret_cast<const float4*>(p.xh2 + net * p.xh2_net + row * C);
      const uint32_t* mask = p.mask + net * p.mask_net + row * p.mw;
      rs[net] = p.rstd[net * p.rstd_net + row];
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const int c4 = lane + 32 * k;
        xv[net][k] = c4 < C4 ? xh4[c4] : make_float4(0.f, 0.f, 0.f, 0.f);
        mwv[net][k] = c4 < C4 ? mask[c4 >> 3] : 0u;
      }
    }
    // ---- affine coupling backward (lane j = flow dimension j)
    float ds_ = 0.f, dt_ = 0.f, dxpv = g;
    if (is_t) {
      const float e = expf(-sv);
      dxpv = g * e;
      dt_ = -dxpv;
      ds_ = -g * zv - dldv;
      p.dout[row * dt + (lane - dc)] = dt_;
      p.dout[p.dout_net + row * dt + (lane - dc)] = ds_;
    }
    if (lane < D) p.dxp[row * D + lane] = dxpv;

#pragma unroll
    for (int net = 0; net < 2; ++net) {
      const float mine = net == 0 ? dt_ : ds_;
      float4 gc[NV];
#pragma unroll
      for (int k = 0; k < NV; ++k) gc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      const float4* Wg = reinterpret_cast<const float4*>(p.W3f + net * p.w_net);
      for (int i = 0; i < dt; ++i) {
        const float d = __shfl_sync(0xffffffffu, mine, dc + i);
#pragma unroll
        for (int k = 0; k < NV; ++k) {
          const int c4 = lane + 32 * k;
          if (c4 < C4) {
            const float4 w = w_smem ? sW4[(net * dt + i) * C4 + c4] : __ldg(Wg + (long long)i * C4 + c4);
            gc[k].x = fmaf(d, w.x, gc[k].x); gc[k].y = fmaf(d, w.y, gc[k].y);
            gc[k].z = fmaf(d, w.z, gc[k].z); gc[k].w = fmaf(d, w.w, gc[k].w);
          }
        }
      }
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        s1 += (gc[k].x + gc[k].y) + (gc[k].z + gc[k].w);
        s2 += dot4(gc[k], xv[net][k]);
      }
      s1 = warp_sum(s1) * invC;
      s2 = warp_sum(s2) * invC;
      float4* du4 = reinterpret_cast<float4*>(p.du2 + net * p.du2_net + row * C);
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const int c4 = lane + 32 * k;
        if (c4 < C4) {
          const uint32_t bits = mwv[net][k] >> (4 * (c4 & 7));
          const float4 x = xv[net][k];
          float4 o;
          o.x = rs[net] * (gc[k].x - s1 - x.x * s2) * ((bits & 1u) ? 1.0f : kLeakySlope);
          o.y = rs[net] * (gc[k].y - s1 - x.y * s2) * ((bits & 2u) ? 1.0f : kLeakySlope);
          o.z = rs[net] * (gc[k].z - s1 - x.z * s2) * ((bits & 4u) ? 1.0f : kLeakySlope);
          o.w = rs[net] * (gc[k].w - s1 - x.w * s2) * ((bits & 8u) ? 1.0f : kLeakySlope);
          du4[c4] = o;
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// backward tail
// ---------------------------------------------------------------------------------------------
template <int NV, int MAXDC>
__global__ void __launch_bounds__(kWarps * 32) k_row_bwd_tail2(BwdTail2P p, int w_smem, int rpw) {
  NF_PDL_ENTRY();
  extern __shared__ float4 sW4[];   // [2][dc][H1/4] when w_smem
  __shared__ float sW[32 * 32];
  const int D = p.D, H1 = p.H1, dc = (D + 1) / 2, H4 = H1 >> 2;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < D * D; i += blockDim.x) sW[i] = p.W[i];
  if (w_smem) {
    for (int e = threadIdx.x; e < 2 * dc * H4; e += blockDim.x) {
      const int net = e / (dc * H4), rem = e - net * dc * H4;
      sW4[e] = __ldg(reinterpret_cast<const float4*>(p.W1T + net * p.w_net) + rem);
    }
  }
  __syncthreads();
  for (int r = 0; r < rpw; ++r) {
    const long long row = ((long long)blockIdx.x * kWarps + warp) * rpw + r;
    if (row >= p.B) return;
    // all global loads of the row first
    float4 dv[2][NV];
#pragma unroll
    for (int net = 0; net < 2; ++net) {
      const float4* du4 = reinterpret_cast<const float4*>(p.du1 + net * p.du1_net + row * H1);
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const int c4 = lane + 32 * k;
        dv[net][k] = c4 < H4 ? du4[c4] : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    const float dxp_in = lane < D ? p.dxp[row * D + lane] : 0.f;
    float dh[MAXDC];
#pragma unroll
    for (int j = 0; j < MAXDC; ++j) dh[j] = 0.f;
#pragma unroll
    for (int net = 0; net < 2; ++net) {
      const float4* Wg = reinterpret_cast<const float4*>(p.W1T + net * p.w_net);
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const int c4 = lane + 32 * k;
        if (c4 < H4) {
#pragma unroll
          for (int j = 0; j < MAXDC; ++j)
            if (j < dc) {
              const float4 w = w_smem ? sW4[(net * dc + j) * H4 + c4] : __ldg(Wg + (long long)j * H4 + c4);
              dh[j] += dot4(dv[net][k], w);
            }
        }
      }
    }
    float add = 0.f;
#pragma unroll
    for (int j = 0; j < MAXDC; ++j)
      if (j < dc) {
        const float t = warp_sum(dh[j]);
        if (lane == j) add = t;
      }
    const float dxpv = dxp_in + add;
    if (lane < dc) p.dxp[row * D + lane] = dxpv;
    if (p.dxin) {   // dx_in[a] = sum_j dxp[j] W[a][j]
      float o = 0.f;
      for (int j = 0; j < D; ++j) {
        const float v = __shfl_sync(0xffffffffu, dxpv, j);
        if (lane < D) o = fmaf(v, sW[lane * D + j], o);
      }
      if (lane < D) p.dxin[row * p.ld_dxin + lane] = o;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// forward / inverse tail (same contract as k_row_tail in rowfused.cu), 128-bit loads issued up front
// ---------------------------------------------------------------------------------------------
template <int NV, int MAXDT>
__global__ void __launch_bounds__(kWarps * 32) k_row_tail2(TailP p, int w_smem, int rpw) {
  NF_PDL_ENTRY();
  extern __shared__ float4 sW4[];   // [2][dt][C/4] when w_smem
  __shared__ float sW[32 * 32];
  const int D = p.D, C = p.C, dc = (D + 1) / 2, dt = D / 2, C4 = C >> 2;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (p.Wmat)
    for (int i = threadIdx.x; i < D * D; i += blockDim.x) sW[i] = p.Wmat[i];
  if (w_smem) {
    for (int e = threadIdx.x; e < 2 * dt * C4; e += blockDim.x) {
      const int net = e / (dt * C4), rem = e - net * dt * C4;
      sW4[e] = __ldg(reinterpret_cast<const float4*>(p.W3f + net * p.w_net) + rem);
    }
  }
  __syncthreads();
  const float cst = p.cdev ? *p.cdev : 0.f;
  for (int r = 0; r < rpw; ++r) {
    const long long row = ((long long)blockIdx.x * kWarps + warp) * rpw + r;
    if (row >= p.B) return;
    float4 xv[2][NV];
#pragma unroll
    for (int net = 0; net < 2; ++net) {
      const float4* xh4 = reinterpret_cast<const float4*>(p.xh2 + net * p.xh2_net + row * C);
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const int c4 = lane + 32 * k;
        xv[net][k] = c4 < C4 ? xh4[c4] : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    const float xin = lane < D ? p.in[row * p.ld_in + lane] : 0.f;
    const float ld_in = (p.logdet && lane == 0) ? p.logdet[row] : 0.f;
    float myS = 0.f, myT = 0.f;   // lane dc + i keeps s_i, t_i
#pragma unroll
    for (int i = 0; i < MAXDT; ++i) {
      if (i < dt) {
        float at = 0.f, as = 0.f;
        const float4* Wt = reinterpret_cast<const float4*>(p.W3f) + (long long)i * C4;
        const float4* Ws = reinterpret_cast<const float4*>(p.W3f + p.w_net) + (long long)i * C4;
#pragma unroll
        for (int k = 0; k < NV; ++k) {
          const int c4 = lane + 32 * k;
          if (c4 < C4) {
            const float4 wt = w_smem ? sW4[i * C4 + c4] : __ldg(Wt + c4);
            const float4 ws = w_smem ? sW4[(dt + i) * C4 + c4] : __ldg(Ws + c4);
            at += dot4(xv[0][k], wt);
            as += dot4(xv[1][k], ws);
          }
        }
        at = warp_sum(at); as = warp_sum(as);
        if (lane == dc + i) { myT = at + p.b3f[i]; myS = as + p.b3f[p.b_net + i]; }
      }
    }
    float zv = xin;
    const bool is_t = lane >= dc && lane < D;
    if (is_t) zv = p.inverse ? fmaf(xin, expf(myS), myT) : (xin - myT) * expf(-myS);
    const float ssum = warp_sum(is_t ? myS : 0.f);
    if (!p.inverse) {
      if (lane < D) p.out[row * p.ld_out + lane] = zv;
      if (p.s_out && is_t) p.s_out[row * dt + (lane - dc)] = myS;
    }
    if (p.logdet && lane == 0) p.logdet[row] = ld_in + (p.inverse ? (ssum - cst) : (cst - ssum));
    if (p.Wmat) {
      float o = 0.f;
      for (int j = 0; j < D; ++j) {
        const float zj = __shfl_sync(0xffffffffu, zv, j);
        if (lane < D) o = fmaf(zj, sW[j * D + lane], o);
      }
      if (lane < D) p.out2[row * p.ld_out2 + lane] = o;
    }
  }
}

int rows_per_warp_for(long long B) {
  long long rpw = B / (kWarps * 148LL * 8);   // one resident wave of CTAs first, then several rows per warp
  return (int)std::min<long long>(std::max<long long>(rpw, 1), 8);
}

// ---------------------------------------------------------------------------------------------
// multi-job column reduction
// ---------------------------------------------------------------------------------------------
// A warp covers 128 columns of one row (lane = 4 adjacent columns, one 128-bit load), the 8 warps of a
// CTA take rows w, w + 8, ... of the CTA's row chunk; coefficients of the chunk sit in shared memory
// (zero-padded to MAXM so the FMA loop needs no predicates).  Loads of kColU rows are issued before
// the first use.  History (ncu, C2a): one column per lane + one row per dependent load = 40 us; batched
// loads = 31 us, instruction-bound (11 M warp instructions); this layout issues ~3x fewer.
constexpr int kColU = 4;       // rows per warp and load batch (8 rows / 64-row CTAs measured slower: twice the atomics)
constexpr int kColSub = 128;   // rows per shared-memory coefficient sub-chunk
constexpr int kColTile = 128;  // columns per CTA
struct ColArgs {
  ColJob j[kColMaxJobs];
  int njobs;
  long long B;
  int rows_per_cta;
};

__device__ __forceinline__ float4 ld4(const float* base, long long off, int nvalid, bool vec) {
  if (nvalid <= 0) return make_float4(0.f, 0.f, 0.f, 0.f);
  if (vec) return __ldg(reinterpret_cast<const float4*>(base + off));
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  v.x = __ldg(base + off);
  if (nvalid > 1) v.y = __ldg(base + off + 1);
  if (nvalid > 2) v.z = __ldg(base + off + 2);
  if (nvalid > 3) v.w = __ldg(base + off + 3);
  return v;
}

template <int MAXM>
__device__ __forceinline__ void col_body(const ColJob& J, int z, long long B, int rows_per_cta, float* s_coef,
                                         float* s_acc) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int c = blockIdx.y * kColTile + lane * 4, m = J.m;
  const int nvalid = min(4, J.N - c);
  const float* A = J.out ? J.A + z * J.sA : nullptr;
  const float* A2 = (J.out2 && J.A2) ? J.A2 + z * J.sA2 : nullptr;
  const float* coef = m ? J.coef + z * J.sCf : nullptr;
  const bool same2 = A2 != nullptr && A2 == A && J.lda2 == J.lda;   // column sums of the weighted matrix itself
  const bool vecA = A && (J.lda & 3) == 0 && (reinterpret_cast<uintptr_t>(A) & 15u) == 0 && nvalid == 4;
  const bool vecA2 = A2 && (J.lda2 & 3) == 0 && (reinterpret_cast<uintptr_t>(A2) & 15u) == 0 && nvalid == 4;
  const bool do3 = J.out3 != nullptr && blockIdx.y == 0 && m > 0;
  const long long row_begin = (long long)blockIdx.x * rows_per_cta;
  const long long row_end = min(B, row_begin + (long long)rows_per_cta);
  float4 acc[MAXM], acc2 = make_float4(0.f, 0.f, 0.f, 0.f);
  float acc3 = 0.f;
#pragma unroll
  for (int i = 0; i < MAXM; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int e = threadIdx.x; e < (MAXM + 2) * kColTile; e += blockDim.x) s_acc[e] = 0.f;

  for (long long sub = row_begin; sub < row_end; sub += kColSub) {
    const int nsub = (int)min((long long)kColSub, row_end - sub);
    if (m) {
#pragma unroll 4
      for (int e = threadIdx.x; e < nsub * MAXM; e += blockDim.x) {
        const int r = e / MAXM, i = e % MAXM;
        s_coef[e] = i < m ? __ldg(coef + (sub + r) * J.ldcf + i) : 0.f;
      }
    }
    __syncthreads();
    for (int r0 = w; r0 < nsub; r0 += 8 * kColU) {
      float4 av[kColU], a2[kColU];
#pragma unroll
      for (int u = 0; u < kColU; ++u) {
        const int r = r0 + 8 * u;
        const int nv = r < nsub ? nvalid : 0;
        av[u] = A ? ld4(A, (sub + r) * J.lda + c, nv, vecA) : make_float4(0.f, 0.f, 0.f, 0.f);
        a2[u] = A2 ? (same2 ? av[u] : ld4(A2, (sub + r) * J.lda2 + c, nv, vecA2)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < kColU; ++u) {
        const int r = r0 + 8 * u;
        if (r < nsub) {
          if (A) {
#pragma unroll
            for (int i4 = 0; i4 < MAXM; i4 += 4) {
              const float4 cf = *reinterpret_cast<const float4*>(s_coef + r * MAXM + i4);
              const float cfv[4] = {cf.x, cf.y, cf.z, cf.w};
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                acc[i4 + k].x = fmaf(cfv[k], av[u].x, acc[i4 + k].x); acc[i4 + k].y = fmaf(cfv[k], av[u].y, acc[i4 + k].y);
                acc[i4 + k].z = fmaf(cfv[k], av[u].z, acc[i4 + k].z); acc[i4 + k].w = fmaf(cfv[k], av[u].w, acc[i4 + k].w);
              }
            }
          }
          acc2.x += a2[u].x; acc2.y += a2[u].y; acc2.z += a2[u].z; acc2.w += a2[u].w;
          if (do3 && lane < m) acc3 += s_coef[r * MAXM + lane];
        }
      }
    }
    __syncthreads();
  }
  // combine the 8 warps through shared memory, then one global atomic per element and CTA
  if (nvalid > 0) {
    if (A) {
#pragma unroll
      for (int i = 0; i < MAXM; ++i)
        if (i < m) {
          float* d = s_acc + i * kColTile + lane * 4;
          atomicAdd(d, acc[i].x); atomicAdd(d + 1, acc[i].y); atomicAdd(d + 2, acc[i].z); atomicAdd(d + 3, acc[i].w);
        }
    }
    if (A2) {
      float* d = s_acc + MAXM * kColTile + lane * 4;
      atomicAdd(d, acc2.x); atomicAdd(d + 1, acc2.y); atomicAdd(d + 2, acc2.z); atomicAdd(d + 3, acc2.w);
    }
  }
  if (do3 && lane < m) atomicAdd(&s_acc[(MAXM + 1) * kColTile + lane], acc3);
  __syncthreads();
  for (int e = threadIdx.x; e < (MAXM + 2) * kColTile; e += blockDim.x) {
    const int i = e / kColTile, t = e % kColTile, cc = blockIdx.y * kColTile + t;
    const float v = s_acc[e];
    if (v == 0.f) continue;
    if (i < MAXM) {
      if (A && i < m && cc < J.N) atomicAdd(&J.out[z * J.sOut + i * J.so_i + cc * J.so_c], v);
    } else if (i == MAXM) {
      if (A2 && cc < J.N) atomicAdd(&J.out2[z * J.sOut2 + cc], v);
    } else {
      if (do3 && t < m) atomicAdd(&J.out3[z * J.sOut3 + t], v);
    }
  }
}

// KMAX = largest m over the jobs of the launch (bounds the register budget of the kernel)
template <int KMAX>
__global__ void __launch_bounds__(256) k_col_reduce(const __grid_constant__ ColArgs a) {
  NF_PDL_ENTRY();
  __shared__ __align__(16) float s_coef[kColSub * KMAX];
  __shared__ float s_acc[(KMAX + 2) * kColTile];
  int job = 0, z = blockIdx.z;
  while (job < a.njobs - 1 && z >= a.j[job].batch) { z -= a.j[job].batch; ++job; }
  const ColJob& J = a.j[job];
  if ((int)blockIdx.y * kColTile >= J.N) return;   // CTA-uniform
  if (J.m <= 4) { col_body<4>(J, z, a.B, a.rows_per_cta, s_coef, s_acc); return; }
  if constexpr (KMAX >= 8) {
    if (J.m <= 8) { col_body<8>(J, z, a.B, a.rows_per_cta, s_coef, s_acc); return; }
  }
  if constexpr (KMAX >= 16) {
    if (J.m <= 16) { col_body<16>(J, z, a.B, a.rows_per_cta, s_coef, s_acc); return; }
  }
  if constexpr (KMAX >= 32) col_body<32>(J, z, a.B, a.rows_per_cta, s_coef, s_acc);
}

// dW (D, D) += X (B, D)^T G (B, D) for the flow dimension D <= 32 (the invertible-linear weight gradient x_in^T dxp):
// both operands of the CTA's 64-row chunk are staged in shared memory with coalesced loads, thread = output element(s).
// As a col_reduce job this took 12.9 us on C2a (32 CTAs, 2 useful lanes per warp); this kernel is bound by its launch.
constexpr int kXtgRows = 64;
__global__ void __launch_bounds__(256) k_xtg_small(const float* __restrict__ X, const float* __restrict__ G,
                                                   float* __restrict__ dW, long long B, int D, int rows_per_cta,
                                                   long long sX, long long sG, long long sW) {
  NF_PDL_ENTRY();
  X += blockIdx.y * sX; G += blockIdx.y * sG; dW += blockIdx.y * sW;   // blockIdx.y: flow block (batched launch)
  __shared__ float sx[kXtgRows * 32], sg[kXtgRows * 32];
  const int DD = D * D;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  const long long r0 = (long long)blockIdx.x * rows_per_cta, r1 = min(B, r0 + rows_per_cta);
  for (long long sub = r0; sub < r1; sub += kXtgRows) {
    const int n = (int)min((long long)kXtgRows, r1 - sub) * D;
    for (int e = threadIdx.x; e < n; e += blockDim.x) { sx[e] = __ldg(X + sub * D + e); sg[e] = __ldg(G + sub * D + e); }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int e = threadIdx.x + 256 * q;
      if (e < DD) {
        const int i = e / D, j = e % D;
        float a = acc[q];
        for (int r = 0; r < n; r += D) a = fmaf(sx[r + i], sg[r + j], a);
        acc[q] = a;
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int e = threadIdx.x + 256 * q;
    if (e < DD) atomicAdd(dW + e, acc[q]);
  }
}

}  // namespace

int xtg_small(const float* X, const float* G, float* dW, long long B, int D, cudaStream_t st, int batch, long long sX,
              long long sG, long long sW) {
  if (B <= 0 || batch <= 0) return 0;
  NF_CHECK_ARG(D >= 1 && D <= 32, NF_E_UNSUPPORTED, "xtg_small: D=%d", D);
  long long rpc = cdiv(cdiv(B, 296), kXtgRows) * kXtgRows;
  if (rpc < kXtgRows) rpc = kXtgRows;
  ProfScope prof(st, PROF_COL_REDUCE, 8.0 * B * D * batch);
  NF_LAUNCH((k_xtg_small), dim3((unsigned)cdiv(B, rpc), (unsigned)batch), 256, 0, st, X, G, dW, B, D, (int)rpc, sX, sG, sW);
  NF_LAUNCHED();
  return 0;
}

bool row_v2_supported(int D, int C, int H1) {
  return D >= 2 && D <= 32 && (C & 3) == 0 && C >= 4 && C <= 1024 && (H1 & 3) == 0 && H1 >= 4 && H1 <= 1024;
}

int row_bwd_head2(const BwdHead2P& p, cudaStream_t st) {
  if (p.B <= 0) return 0;
  NF_CHECK_ARG(row_v2_supported(p.D, p.C, 4), NF_E_UNSUPPORTED, "row_bwd_head2: D=%d C=%d", p.D, p.C);
  const int dt = p.D / 2;
  const size_t wbytes = (size_t)2 * dt * p.C * 4;
  const int w_smem = wbytes <= (size_t)kMaxWSmem;
  const int rpw = rows_per_warp_for(p.B);
  dim3 grid((unsigned)cdiv(p.B, (long long)kWarps * rpw));
  const size_t smem = w_smem ? wbytes : 0;
  ProfScope prof(st, PROF_ROW_HEAD, 4.0 * p.B * (4.0 * p.C + 4.0 * p.D));
  if (p.C <= 256) NF_LAUNCH((k_row_bwd_head2<2>), grid, kWarps * 32, smem, st, p, w_smem, rpw);
  else if (p.C <= 512) NF_LAUNCH((k_row_bwd_head2<4>), grid, kWarps * 32, smem, st, p, w_smem, rpw);
  else NF_LAUNCH((k_row_bwd_head2<8>), grid, kWarps * 32, smem, st, p, w_smem, rpw);
  NF_LAUNCHED();
  return 0;
}

int row_bwd_tail2(const BwdTail2P& p, cudaStream_t st) {
  if (p.B <= 0) return 0;
  NF_CHECK_ARG(row_v2_supported(p.D, 4, p.H1), NF_E_UNSUPPORTED, "row_bwd_tail2: D=%d H1=%d", p.D, p.H1);
  const int dc = (p.D + 1) / 2;
  const size_t wbytes = (size_t)2 * dc * p.H1 * 4;
  const int w_smem = wbytes <= (size_t)kMaxWSmem;
  const int rpw = rows_per_warp_for(p.B);
  dim3 grid((unsigned)cdiv(p.B, (long long)kWarps * rpw));
  const size_t smem = w_smem ? wbytes : 0;
  ProfScope prof(st, PROF_ROW_BTAIL, 4.0 * p.B * (2.0 * p.H1 + 3.0 * p.D));
#define NF_TAIL2(NV)                                                                              \
  do {                                                                                            \
    if (dc <= 8) NF_LAUNCH((k_row_bwd_tail2<NV, 8>), grid, kWarps * 32, smem, st, p, w_smem, rpw);         \
    else NF_LAUNCH((k_row_bwd_tail2<NV, 16>), grid, kWarps * 32, smem, st, p, w_smem, rpw);                \
  } while (0)
  if (p.H1 <= 256) NF_TAIL2(2);
  else if (p.H1 <= 512) NF_TAIL2(4);
  else NF_TAIL2(8);
#undef NF_TAIL2
  NF_LAUNCHED();
  return 0;
}

int row_tail2(const TailP& p, cudaStream_t st) {
  if (p.B <= 0) return 0;
  NF_CHECK_ARG(row_v2_supported(p.D, p.C, 4), NF_E_UNSUPPORTED, "row_tail2: D=%d C=%d", p.D, p.C);
  const int dt = p.D / 2;
  const size_t wbytes = (size_t)2 * dt * p.C * 4;
  const int w_smem = wbytes <= (size_t)kMaxWSmem;
  const int rpw = rows_per_warp_for(p.B);
  dim3 grid((unsigned)cdiv(p.B, (long long)kWarps * rpw));
  const size_t smem = w_smem ? wbytes : 0;
  ProfScope prof(st, PROF_ROW_TAIL, 4.0 * p.B * (2.0 * p.C + 3.0 * p.D));
#define NF_TAILF(NV)                                                                           \
  do {                                                                                         \
    if (dt <= 4) NF_LAUNCH((k_row_tail2<NV, 4>), grid, kWarps * 32, smem, st, p, w_smem, rpw);          \
    else NF_LAUNCH((k_row_tail2<NV, 16>), grid, kWarps * 32, smem, st, p, w_smem, rpw);                 \
  } while (0)
  if (p.C <= 256) NF_TAILF(2);
  else if (p.C <= 512) NF_TAILF(4);
  else NF_TAILF(8);
#undef NF_TAILF
  NF_LAUNCHED();
  return 0;
}

int col_reduce(const ColJob* jobs, int njobs, long long B, cudaStream_t st) {
  if (B <= 0 || njobs <= 0) return 0;
  NF_CHECK_ARG(njobs <= kColMaxJobs, NF_E_BADARG, "col_reduce: too many jobs (%d)", njobs);
  ColArgs a{};
  int tiles = 0, zs = 0, maxm = 0;
  double bytes = 0;
  for (int j = 0; j < njobs; ++j) {
    a.j[j] = jobs[j];
    NF_CHECK_ARG(jobs[j].m >= 0 && jobs[j].m <= 32 && jobs[j].N >= 1 && jobs[j].batch >= 1, NF_E_BADARG,
                 "col_reduce: job %d has m=%d N=%d", j, jobs[j].m, jobs[j].N);
    tiles = std::max(tiles, (int)cdiv(jobs[j].N, kColTile));
    maxm = std::max(maxm, jobs[j].m);
    zs += jobs[j].batch;
    bytes += 4.0 * B * jobs[j].batch * ((jobs[j].out ? jobs[j].N : 0) + (jobs[j].out2 ? jobs[j].N : 0) + jobs[j].m);
  }
  a.njobs = njobs;
  a.B = B;
  a.rows_per_cta = B <= 16384 ? 128 : 1024;
  dim3 grid((unsigned)cdiv(B, a.rows_per_cta), (unsigned)tiles, (unsigned)zs);
  ProfScope prof(st, PROF_COL_REDUCE, bytes);
  if (maxm <= 4) NF_LAUNCH((k_col_reduce<4>), grid, 256, 0, st, a);
  else if (maxm <= 8) NF_LAUNCH((k_col_reduce<8>), grid, 256, 0, st, a);
  else if (maxm <= 16) NF_LAUNCH((k_col_reduce<16>), grid, 256, 0, st, a);
  else NF_LAUNCH((k_col_reduce<32>), grid, 256, 0, st, a);
  NF_LAUNCHED();
  return 0;
}

}  // namespace nf
