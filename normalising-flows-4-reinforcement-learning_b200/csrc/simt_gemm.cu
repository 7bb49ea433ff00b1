// Exact-fp32 FFMA GEMM (64x64x16 tiles, 256 threads, 4x4 register tiles).
#include "prof.cuh"
#include "simt_gemm.cuh"

namespace nf {

namespace {
constexpr int BM = 64, BN = 64, BK = 16, PADM = 4;

template <bool TA, bool TB>
__global__ void __launch_bounds__(256) k_gemm(GemmP p) {
  NF_PDL_ENTRY();
  __shared__ __align__(16) float As[BK][BM + PADM];
  __shared__ __align__(16) float Bs[BK][BN + PADM];

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int tiles_n = (p.N + BN - 1) / BN;
  const int tile_m = blockIdx.x / tiles_n, tile_n = blockIdx.x % tiles_n;
  const int m0 = tile_m * BM, n0 = tile_n * BN;
  const int z = blockIdx.z;

  const float* A0 = p.A0 + (long long)z * p.sA0;
  const float* B0 = p.B0 + (long long)z * p.sB0;
  const float* A1 = p.A1 ? p.A1 + (long long)z * p.sA1 : nullptr;
  const float* B1 = p.B1 ? p.B1 + (long long)z * p.sB1 : nullptr;
  float* C = p.C + (long long)z * p.sC;

  // k-tiles of segment 0 then segment 1, split across blockIdx.y
  const int kt0 = (p.K0 + BK - 1) / BK;
  const int kt1 = (p.K1 + BK - 1) / BK;
  const int kt = kt0 + kt1;
  const int per = (kt + p.splitk - 1) / p.splitk;
  const int kt_begin = blockIdx.y * per;
  const int kt_end = min(kt, kt_begin + per);

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j] = 0.f;

  for (int t = kt_begin; t < kt_end; ++t) {
    const bool seg1 = t >= kt0;
    const float* A = seg1 ? A1 : A0;
    const float* B = seg1 ? B1 : B0;
    const int lda = seg1 ? p.lda1 : p.lda0;
    const int ldb = seg1 ? p.ldb1 : p.ldb0;
    const int K = seg1 ? p.K1 : p.K0;
    const int k0 = (seg1 ? t - kt0 : t) * BK;

    // ---- load A tile into As[k][m]
    if (!TA) {
      const int k = tid & 15;
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const int m = (tid >> 4) + 16 * i;
        const int gm = m0 + m, gk = k0 + k;
        As[k][m] = (gm < p.M && gk < K) ? A[(long long)gm * lda + gk] : 0.f;
      }
    } else {
      const int m = tid & 63;
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const int k = (tid >> 6) + 4 * i;
        const int gm = m0 + m, gk = k0 + k;
        As[k][m] = (gm < p.M && gk < K) ? A[(long long)gk * lda + gm] : 0.f;
      }
    }
    // ---- load B tile into Bs[k][n]
    if (!TB) {
      const int n = tid & 63;
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const int k = (tid >> 6) + 4 * i;
        const int gn = n0 + n, gk = k0 + k;
        Bs[k][n] = (gn < p.N && gk < K) ? B[(long long)gk * ldb + gn] : 0.f;
      }
    } else {
      const int k = tid & 15;
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const int n = (tid >> 4) + 16 * i;
        const int gn = n0 + n, gk = k0 + k;
        Bs[k][n] = (gn < p.N && gk < K) ? B[(long long)gn * ldb + gk] : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; k++) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

  const float* bias = p.bias ? p.bias + (long long)z * p.sBias : nullptr;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int gn = n0 + tx * 4 + j;
      if (gn >= p.N) continue;
      float v = p.alpha * acc[i][j];
      float* dst = C + (long long)gm * p.ldc + gn;
      if (p.splitk > 1) {
        if (bias && blockIdx.y == 0) v += bias[gn];
        atomicAdd(dst, v);
      } else {
        if (bias) v += bias[gn];
        *dst = p.accumulate ? (*dst + v) : v;
      }
    }
  }
}
}  // namespace

int pick_splitk(int M, int N, int64_t K, int batch) {
  const int64_t tiles = cdiv(M, BM) * cdiv(N, BN) * batch;
  const int64_t ktiles = cdiv(K, BK);
  int64_t want = cdiv(148 * 4, tiles);          // ~4 CTAs per SM
  int64_t maxsplit = ktiles / 8 > 0 ? ktiles / 8 : 1;  // at least 8 k-tiles per split
  int64_t s = want < maxsplit ? want : maxsplit;
  if (s < 1) s = 1;
  if (s > 512) s = 512;
  return (int)s;
}

int simt_gemm(const GemmP& p, cudaStream_t st) {
  if (p.M <= 0 || p.N <= 0 || p.batch <= 0) return 0;
  NF_CHECK_ARG(p.K1 == 0 || (p.A1 && p.B1), NF_E_BADARG, "simt_gemm: segment 1 without pointers");
  dim3 grid((unsigned)(cdiv(p.M, BM) * cdiv(p.N, BN)), (unsigned)p.splitk, (unsigned)p.batch);
  NF_CHECK_ARG(p.splitk >= 1 && p.splitk <= 65535 && p.batch <= 65535, NF_E_BADARG, "simt_gemm: grid");
  ProfScope prof(st, PROF_SIMT_GEMM, 2.0 * p.M * p.N * (double)(p.K0 + p.K1) * p.batch);
  if (!p.transA && !p.transB) NF_LAUNCH((k_gemm<false, false>), grid, 256, 0, st, p);
  else if (!p.transA && p.transB) NF_LAUNCH((k_gemm<false, true>), grid, 256, 0, st, p);
  else if (p.transA && !p.transB) NF_LAUNCH((k_gemm<true, false>), grid, 256, 0, st, p);
  else NF_LAUNCH((k_gemm<true, true>), grid, 256, 0, st, p);
  NF_LAUNCHED();
  return 0;
}

}  // namespace nf
