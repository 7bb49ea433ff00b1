// tcgen05 (5th-gen tensor core) GEMM for the conditioner layers, sm_100a only.
//
//   C[z][m, n] (+)= sum_k A[z][m, k] * B[z][n, k]   (+ bias[z][n])
//
// Both operands are K-major fp32 in global memory.  fp32 parity is obtained with the 3xTF32
// split: every operand tile is split inside the kernel into hi = rna_tf32(x) and lo = x - hi and
// three kind::tf32 MMAs (lo*hi, hi*lo, hi*hi) accumulate in fp32 in TMEM.
#pragma once
#include "common.cuh"

namespace nf {

// Row-wise epilogues fused into the GEMM (the cluster of CTAs along N spans the whole output row):
//   TC_EPI_LN_FWD: C = LayerNorm_noaffine(LeakyReLU(acc + bias)), rstd (M) and sign mask (M, mw words) written
//   TC_EPI_LN_BWD: C = LeakyReLU'(mask) rstd (acc - mean(acc) - xh mean(acc xh))   (xh, rstd, mask read)
enum { TC_EPI_NONE = 0, TC_EPI_LN_FWD = 1, TC_EPI_LN_BWD = 2 };
struct TcEpi {
  int kind = TC_EPI_NONE;
  float* rstd = nullptr; long long sRstd = 0;
  uint32_t* mask = nullptr; long long sMask = 0; int mw = 0;
  const float* xh = nullptr; long long ldxh = 0, sXh = 0;
  float eps = 1e-5f;
  int fast_var = 0;   // TC_EPI_LN_FWD: flax variance E[x^2] - E[x]^2 (see LnFwdP)
  // TC_EPI_LN_FWD only: finish the flow block inside the same kernel (see tc_gemm_tail_supported).  Same
  // contract as TailP / row_tail2 (rowfused.cuh); store_c = 0 skips writing C (x_hat) altogether.
  struct Tail {
    int on = 0, store_c = 1, D = 0, inverse = 0, ld_in = 0, ld_out = 0, ld_out2 = 0;
    const float* W3f = nullptr; long long w_net = 0;
    const float* b3f = nullptr; long long b_net = 0;
    const float* in = nullptr; const float* cdev = nullptr; const float* Wmat = nullptr;
    float* out = nullptr; float* out2 = nullptr; float* logdet = nullptr; float* s_out = nullptr;
  } tail;
};

struct TcGemmP {
  // K-major operands; the reduction may be two segments (A0|A1) x (B0|B1)
  const float* A0 = nullptr; const float* A1 = nullptr;
  const float* B0 = nullptr; const float* B1 = nullptr;
  // optional pre-split copies of the B segments (hi = rna_tf32(B), lo = B - hi, same pitches and strides): the
  // TS kernels then load both with the TMA and the splitter threads only handle the A rows
  const float* B0hi = nullptr; const float* B0lo = nullptr; const float* B1hi = nullptr; const float* B1lo = nullptr;
  float* C = nullptr;
  const float* bias = nullptr;
  int M = 0, N = 0, K0 = 0, K1 = 0;
  long long lda0 = 0, lda1 = 0, ldb0 = 0, ldb1 = 0, ldc = 0;     // row pitches in floats (multiples of 4)
  long long sA0 = 0, sA1 = 0, sB0 = 0, sB1 = 0, sC = 0, sBias = 0;  // batch strides in floats (0 = shared)
  int batch = 1;
  int accumulate = 0;
  int passes = 3;       // 3 = 3xTF32 (fp32 parity), 1 = single TF32 pass
  int raw_hi = 1;       // feed the raw fp32 tile as the hi operand (the tensor core ignores the low 13 bits):
                        // lo = x - trunc_tf32(x); saves one shared-memory store per element (measured +10%)
  int bk = 16;          // reduction elements per pipeline stage: 16 or 32
  TcEpi epi;            // optional fused row-wise epilogue (see tc_gemm_epi_supported)
};

// Weight-gradient form: C[z][m, n] += sum_k A[z][k, m] * B[z][k, n]  (both operands MN-major: the
// reduction runs over the batch rows).  The reduction is split over `splitk` CTAs whose partial
// tiles are combined with red.global.add, so C must hold the base value (zeros) on entry.
struct TcGemmTnP {
  const float* A = nullptr; const float* B = nullptr; float* C = nullptr;
  int M = 0, N = 0; long long K = 0;
  long long lda = 0, ldb = 0, ldc = 0;       // pitches in floats (lda, ldb multiples of 4)
  long long sA = 0, sB = 0, sC = 0;          // batch strides in floats (0 = shared operand)
  int batch = 1;
  int splitk = 0;                            // 0 = choose
  int passes = 3;
  int bk = 16;
  // reductions over the batch rows that share the A operand, taken by the kernel's splitter threads (TS kernels only,
  // see tc_gemm_tn_rides_supported; outputs are accumulated with atomics and must hold the base value on entry):
  //   colsum[z][m] += sum_k A[z][k, m];   skinny[z][m, j] += sum_k A[z][k, m] * xc[k, j], j < nxc <= 16 (xc shared by z)
  float* colsum = nullptr; long long sColsum = 0;
  const float* xc = nullptr; long long ldxc = 0; int nxc = 0;
  float* skinny = nullptr; long long ldsk = 0, sSk = 0;
  // two-level batch (inner > 0): z = zo * inner + zi.  A and B still advance by sA / sB per z; the outputs live at
  // zo * s?2 + zi * s? and xc at zo * sXc2 (one launch for the weight gradients of several blocks x 2 nets)
  int inner = 0; long long sC2 = 0, sColsum2 = 0, sSk2 = 0, sXc2 = 0;
};
bool tc_gemm_tn_rides_supported(const TcGemmTnP& p);
int tc_gemm_tn(const TcGemmTnP& p, cudaStream_t st);
bool tc_gemm_tn_supported(const TcGemmTnP& p);

// returns 0, or NF_E_UNSUPPORTED when the shape/alignment cannot be handled by TMA (caller decides)
int tc_gemm(const TcGemmP& p, cudaStream_t st);
bool tc_gemm_supported(const TcGemmP& p);
// the fused epilogues need N = ncl * BN with BN in {64, 128} and ncl <= 8 (one cluster per 128-row tile)
bool tc_gemm_epi_supported(const TcGemmP& p);
// fused tail: both nets in one cluster (batch == 2, 2 * N / BN <= 8 CTAs) and D / 2 <= 4
bool tc_gemm_tail_supported(const TcGemmP& p, int D);

}  // namespace nf
