// tcgen05 / TMEM / TMA GEMM for the conditioner layers (sm_100a).  See tc_gemm.cuh.
//
// One 128 x BN output tile per CTA; fp32 parity through the 3xTF32 split (lo*hi, hi*lo, hi*hi).
//   warp 0        TMA producer: operand tiles -> shared-memory ring (SWIZZLE_64B / 128B); pre-split weight
//                 tables arrive as two tiles (hi, lo)
//   warp 1        MMA issuer: the whole warp runs the loop with warp-uniform operands, one elected lane issues
//                 tcgen05.mma kind::tf32 and the commits; owns tcgen05.alloc / dealloc
//   warps 2..5    splitter group 0 (and 6..9 = group 1 in the TS kernels, alternating stages): thread = tile
//                 row = TMEM lane.  TS form: reads its row of the A tile, writes hi / lo to a 4-slot TMEM ring
//                 (tcgen05.st), splits B in shared memory only when it is not pre-split.  SS form: splits both
//                 tiles in shared memory (hi in place, lo to a twin).  Then the epilogue.
// Barriers: full[s] (TMA -> splitters), split[s] (splitters -> MMA), empty[s] (MMA commit -> producer),
// aempty[i] (MMA commit -> splitters, TMEM A slot i), accbar (MMA commit -> epilogue), xbar / tbar (cluster).
// Accumulators in TMEM: the hi*hi products rotate over NACC `main` accumulators, the cross terms go to `corr`
// (the tensor core truncates when it accumulates: fewer sequential adds per accumulator = less drift); the
// epilogue adds them with round-to-nearest.
// Epilogues (the splitter groups share the tile's columns): plain (bias / accumulate / split-K reduce-add through bulk tensor stores), EPI_LN_FWD and
// EPI_LN_BWD (LeakyReLU + LayerNorm forward / backward over a CTA cluster spanning the row, optional fused
// flow tail).  MNMAJOR = true is the weight-gradient form (reduction over the batch rows, split-K).
#include <cuda.h>
#include <stdlib.h>

#include <algorithm>

#include "prof.cuh"
#include "tc_common.cuh"
#include "tc_gemm.cuh"

namespace nf {
namespace {

constexpr int BM = 128;
constexpr int NTHREADS = 192;
// BK = fp32 elements of the reduction per pipeline stage: 32 (128 B rows, SWIZZLE_128B) or 16 (64 B rows,
// SWIZZLE_64B; twice the stages in the same shared memory, which hides the TMA + split latency better).

struct KArgs {
  float* C;
  const float* bias;
  int M, N, K0, K1;
  long long ldc, sC, sBias;
  int zA0, zA1, zB0, zB1;  // 1: operand is batched over blockIdx.z, 0: shared
  int accumulate, passes, raw_hi;
  int nacc;                       // main accumulators in rotation, 0 = all NACC (option "nacc")
  int splitk, chunks_per_split;   // MN-major form only
  int shared_c;                   // K-major form: the batch entries accumulate into ONE output (red.add), e.g. dy over blocks
  // MN-major TS form, blockIdx.y == 0 CTAs only: the splitter threads already hold every element of their A column
  // (thread = output row m, elements = the batch rows of the stage), so the reductions over the batch rows that share
  // this operand ride along: colsum[z][m] += sum_k A[k, m] (the bias gradient) and skinny[z][m, j] += sum_k A[k, m] *
  // xc[k, j], j < nxc <= 16 (the x_cond columns of dW1) -- instead of a separate column-reduction pass over A
  float* colsum; long long sColsum;
  const float* xc; long long ldxc; int nxc;
  int xc_vec;                            // xc rows may be read as float4 chunks (aligned, 4 ceil(nxc / 4) <= row length), nxc <= 8
  int xc_al;                             // the same without the column limit (prefetched form, up to 16 columns)
  float* skinny; long long ldsk, sSk;
  // MN-major form, two-level batch (zin > 0): z = zo * zin + zi; A / B keep one uniform stride (TMA coordinate z), the
  // outputs live at zo * s?2 + zi * s? (the per-block / per-net slots of the gradient arena) and xc at zo * sXc2
  int zin; long long sC2, sColsum2, sSk2, sXc2;
  // row-wise epilogues over the cluster (EPI_LN_FWD / EPI_LN_BWD)
  float* rstd; long long sRstd;          // fwd: out (row), bwd: in
  uint32_t* mask; long long sMask; int mw;   // fwd: out, bwd: in; mw words per row
  const float* xh; long long ldxh, sXh;  // bwd: normalised activations of the layer (stash)
  float eps;
  int fast_var;                          // LayerNorm variance as E[x^2] - E[x]^2 (flax)
  int ncl;                               // CTAs per cluster = N / BN (the cluster spans the whole row)
  long long* dbg;                        // optional: 8 clock64 timestamps per CTA (nf_debug_tc_timeline)
  int tma_store;                         // mC is valid: write the output tile with bulk tensor stores
  int b_presplit;                        // TS form: mB0 / mB1 are the hi tables and mB0l / mB1l the lo tables of B
  // fused flow tail (EPI_LN_FWD on layer 2, D/2 <= 4, cluster (1, ncl, 2) = both nets): last Linear, affine
  // coupling, log-det and the neighbouring PLU matmul finished by the cluster's leader CTA (rowfused.cuh TailP)
  int tail, store_c;                     // store_c = 0: x_hat is not written (inference: nothing else reads it)
  int t_D, t_inverse, t_ld_in, t_ld_out, t_ld_out2;
  const float* t_W3f; long long t_w_net;
  const float* t_b3f; long long t_b_net;
  const float* t_in; const float* t_cdev; const float* t_Wmat;
  float* t_out; float* t_out2; float* t_logdet; float* t_s_out;
};
constexpr int EPI_NONE = 0, EPI_LN_FWD = 1, EPI_LN_BWD = 2;

// NS = true ("no split"): single TF32 pass on the raw fp32 tiles, SS form -- a stage is just [A | B], so twice as many
// stages are in flight in the same shared memory (the loop is bound by bytes in flight / round-trip latency).
template <int BN, int STAGES, int BK, bool TS = false, bool NS = false>
struct Smem {
  static constexpr int A_BYTES = BM * BK * 4;
  static constexpr int B_BYTES = BN * BK * 4;
  static constexpr int HALF = A_BYTES + B_BYTES;   // raw/hi tiles; the lo twins follow
  // SS form: [A | B | A lo | B lo]; TS form: [A | B | B lo] (A's halves live in tensor memory), which buys two more
  // stages of TMA prefetch in the same shared memory.  The lo twin of tile byte o sits at LO + o in both.
  static constexpr int LO = NS ? 0 : (TS ? B_BYTES : HALF);
  static constexpr int STAGE = NS ? HALF : (TS ? A_BYTES + 2 * B_BYTES : 2 * HALF);
  static constexpr int BARS = STAGES * STAGE;      // byte offset of the barrier block
  static constexpr int XCHG = BARS + 512;          // [kMaxCluster][128] float2 partial row statistics (epilogue kernels)
  static constexpr int TOTAL = BARS + 512 + 1024;  // + alignment slack
  static constexpr int TPART = XCHG + 8 * 128 * 8;                 // [8 ranks][128] float4 partial last-Linear outputs (leader CTA)
  static constexpr int TOTAL_EPI = TPART + 8 * 128 * 16 + 1024;
};

// TS = true: the A operand (the 128 activation rows of the tile) is fed to the MMAs from TENSOR MEMORY instead of
// shared memory.  The main loop of the SS form is bound by shared-memory bandwidth (96 KB of traffic per
// 16-deep stage): each k8 step reads the A tile three times.  Here the splitter threads
// (thread = tile row, the TMEM lane it may access) read their row of the TMA tile once, split it and write hi / lo
// with tcgen05.st into a 4-slot ring of 32 TMEM columns; the MMAs read A from there ([a_tmem] operand form) and
// only the B operand comes from shared memory (64 KB per stage).  The ring takes 128 TMEM columns, so 128-wide
// tiles rotate over two main accumulators instead of three (all golden cases, C5 with K = 1024 included, stay
// inside the parity contract).
// NSG = splitter groups of four warps.  One group needs ~1 100 cycles per 16-deep stage (wait for the tile 120, load +
// split 320, wait for the TMEM slot 300, tcgen05.st + wait + fences + arrive 340 -- a chain of fixed latencies, measured
// with clock64) against 384 cycles of MMAs, so the TS kernels run two groups on alternating stages; both groups
// share the epilogue (plain: alternating 32-column chunks; row-wise: each group owns half of the tile's columns, which
// also halves the registers a thread needs for its row -- 320 threads leave 200 per thread).
// MC = true (K-major TS form with pre-split weights, plain epilogue): the two CTAs of a (2, 1, 1) cluster compute
// neighbouring row tiles of the same output columns, i.e. they consume the SAME weight tiles.  Each loads half of every B
// tile (hi and lo) and the TMA multicasts it into both CTAs' stage, so a weight tile leaves L2 once per CTA pair (16 KB ->
// 8 KB per stage and CTA, section 8 of DESIGN.md).  A stage may be refilled only when BOTH CTAs' MMAs have read it: the
// MMA commits arrive on the empty barrier of both CTAs (count 2).
template <int BN, int STAGES, bool MNMAJOR, int BK, int EPI = EPI_NONE, bool TS = false, int NSG = 1, bool NS = false, bool MC = false>
__global__ void __launch_bounds__(64 + 128 * NSG, 1)
k_tc_gemm(const __grid_constant__ CUtensorMap mA0, const __grid_constant__ CUtensorMap mA1,
          const __grid_constant__ CUtensorMap mB0, const __grid_constant__ CUtensorMap mB1,
          const __grid_constant__ CUtensorMap mB0l, const __grid_constant__ CUtensorMap mB1l,
          const __grid_constant__ CUtensorMap mC, const KArgs a) {
  using S = Smem<BN, STAGES, BK, TS, NS>;
  static_assert(!NS || !TS, "the no-split form feeds both operands from shared memory");
  static_assert(!MC || (TS && !MNMAJOR && EPI == EPI_NONE), "weight-tile multicast: K-major TS form, plain epilogue");
  constexpr int A_BYTES = S::A_BYTES;
  constexpr int BOX = 32 * BK * 4;   // bytes of one 32-column MN-major TMA box
  extern __shared__ uint8_t smem_dyn[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::BARS);
  uint64_t* full = bars;
  uint64_t* split = bars + STAGES;
  uint64_t* empty = bars + 2 * STAGES;
  uint64_t* accbar = bars + 3 * STAGES;
  uint64_t* xbar = bars + 3 * STAGES + 1;      // cluster exchange barrier (epilogue kernels)
  uint64_t* tbar = bars + 3 * STAGES + 2;      // fused tail: partials have arrived at the leader CTA
  uint64_t* aempty = bars + 3 * STAGES + 3;    // TS: the MMAs have read TMEM A slot i (4 slots)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * STAGES + 11);   // (up to 8 aempty barriers)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  long long* dbg = a.dbg ? a.dbg + 8ll * (blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * (long long)blockIdx.z)) : nullptr;
#define NF_TS(i) do { if (dbg) dbg[i] = clock64(); } while (0)
  if (threadIdx.x == 64) NF_TS(0);
  const int z = blockIdx.z / a.splitk;
  // output offsets of this batch entry (two-level in the batched weight-gradient launches)
  const int z_o = (MNMAJOR && a.zin > 0) ? z / a.zin : 0, z_i = (MNMAJOR && a.zin > 0) ? z % a.zin : z;
  const long long c_off = (long long)z_o * a.sC2 + (long long)z_i * a.sC;
  const int nk0 = (a.K0 + BK - 1) / BK, nk1 = (a.K1 + BK - 1) / BK;
  // chunk range of this CTA: everything (K-major form) or one split of the batch rows (MN-major form)
  // both forms can split the reduction over CTAs (partial tiles are combined with red.global.add)
  const int kc_begin = a.splitk > 1 ? (blockIdx.z % a.splitk) * a.chunks_per_split : 0;
  const int nk = a.splitk > 1 ? min((MNMAJOR ? nk0 : nk0 + nk1) - kc_begin, a.chunks_per_split)
                              : (MNMAJOR ? nk0 : nk0 + nk1);
  // The tensor core TRUNCATES (toward zero) when it adds products into the fp32 accumulator (measured with
  // tools/tc_precision.py: all-positive operands lose 5.4e-8 of the sum per k8 step, mirrored for negative
  // ones; random-sign data sees zero-mean noise that also grows linearly with the reduction length -- a
  // multiplicative correction therefore does not help, it was tried).  The hi*hi products therefore rotate over NACC independent
  // accumulators that are summed with round-to-nearest adds in the epilogue (bias / NACC), and the
  // weight-gradient form additionally caps the rows per CTA (launch_tn).
  static_assert(!TS || (BK == 16 && BN <= 128), "TS form: 16-deep stages, BN <= 128");
  constexpr int NACC = TS ? (BN <= 64 ? 3 : 2) : (BN <= 128 ? 3 : 1);
#ifndef NF_ASLOTS64
#define NF_ASLOTS64 8
#endif
  // TS: TMEM A ring (hi 16 + lo 16 columns per slot).  The ring depth bounds the main loop: a slot is reusable only
  // after  MMAs complete -> commit -> splitter tcgen05.st -> split barrier -> MMA issue, a fixed ~2 500-cycle round
  // trip, so the loop runs at (slots / round trip) stages per cycle whatever the tile width or the number of MMA
  // passes (measured: 640 cycles per stage with 4 slots for BN = 64 and 128, one pass and three).  128-wide tiles
  // have 128 TMEM columns left for 4 slots; 64-wide tiles have room for 8.
  constexpr int kASlots = BN <= 64 ? NF_ASLOTS64 : 4;
  constexpr uint32_t A_BASE = (NACC + 1) * BN;                        // first TMEM column of the A ring
  constexpr uint32_t TMEM_NEED = (NACC + 1) * BN + (TS ? kASlots * 2 * BK : 0);
  constexpr uint32_t TMEM_COLS = TMEM_NEED <= 256 ? 256 : 512;        // NACC main + 1 corr accumulator (+ A ring)
  // k8 steps this CTA issues (the epilogue must know how many main accumulators were written)
  int total_k8;
  if (MNMAJOR) {
    const long long len = min((long long)a.K0, (long long)(kc_begin + nk) * BK) - (long long)kc_begin * BK;
    total_k8 = (int)((len + 7) / 8);
  } else {
    total_k8 = 0;
    for (int g = kc_begin; g < kc_begin + nk; ++g) {
      const int kleft = g >= nk0 ? a.K1 - (g - nk0) * BK : a.K0 - g * BK;
      total_k8 += kleft >= BK ? BK / 8 : (kleft + 7) / 8;
    }
  }
  // Rotating over all NACC accumulators is also FASTER than using fewer (measured: 840 vs 940-1100 cycles per
  // 16-deep stage): back-to-back MMAs into one accumulator serialise on it.  The "nacc" option
  // (tests, experiments) can lower the count; the default is the full rotation with a separate corr accumulator.
  const int nacc_rt = a.nacc < 1 ? NACC : (a.nacc > NACC ? NACC : a.nacc);
  const int nacc_used = total_k8 < nacc_rt ? total_k8 : nacc_rt;
  const bool with_corr = !NS && a.passes == 3;

  if (threadIdx.x == 0) {
    prefetch_tmap(&mA0); prefetch_tmap(&mB0);
    if (a.tma_store) prefetch_tmap(&mC);
    if (TS && a.b_presplit) { prefetch_tmap(&mB0l); if (nk1) prefetch_tmap(&mB1l); }
    if (nk1) { prefetch_tmap(&mA1); prefetch_tmap(&mB1); }
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&split[s], 4); mbar_init(&empty[s], MC ? 2 : 1); }
    mbar_init(accbar, 1);
    if (TS) for (int i = 0; i < kASlots; ++i) mbar_init(&aempty[i], 1);
    if (EPI != EPI_NONE) mbar_init(xbar, (uint32_t)a.ncl * 128u * NSG);
    if (EPI == EPI_LN_FWD) mbar_init(tbar, 2u * (uint32_t)a.ncl * 128u);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (EPI != EPI_NONE || MC) cluster_sync_all();   // every CTA's barriers are initialised before a peer arrives on them
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  // everything above (barrier init, TMEM allocation, tensor-map prefetch) overlaps the previous kernel
  NF_PDL_ENTRY();
  if (threadIdx.x == 64) NF_TS(1);   // prologue done

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    {
      for (int it = 0; it < nk; ++it) {
        const int s = it % STAGES, use = it / STAGES;
        if (lane == 0) mbar_wait(&empty[s], (use & 1) ^ 1);   // one poller per warp: 32 spinning lanes steal smem cycles
        __syncwarp();
        uint8_t* st = smem + s * S::STAGE;
        if (elect_one()) {
          mbar_expect_tx(&full[s], S::HALF + ((TS && !MNMAJOR && a.b_presplit) ? S::B_BYTES : 0));
          if (MNMAJOR) {
            // operands are (K rows, M|N columns): one 32 x 32 box per 32-column group
            const int k0 = (kc_begin + it) * BK;
#pragma unroll
            for (int g = 0; g < BM / 32; ++g) tma_load_3d(st + g * BOX, &mA0, &full[s], m0 + g * 32, k0, a.zA0 ? z : 0);
#pragma unroll
            for (int g = 0; g < BN / 32; ++g)
              tma_load_3d(st + A_BYTES + g * BOX, &mB0, &full[s], n0 + g * 32, k0, a.zB0 ? z : 0);
          } else {
            const int g = kc_begin + it;
            const bool seg1 = g >= nk0;
            const int kk = (seg1 ? g - nk0 : g) * BK;
            tma_load_3d(st, seg1 ? &mA1 : &mA0, &full[s], kk, m0, (seg1 ? a.zA1 : a.zA0) ? z : 0);
            if constexpr (MC) {
              // this CTA's half of the weight tile (BN / 2 rows of hi and of lo), delivered to both CTAs of the pair
              const int half = (int)cluster_ctarank() * (BN / 2);
              const int zb = (seg1 ? a.zB1 : a.zB0) ? z : 0;
              tma_load_3d_mc(st + A_BYTES + half * BK * 4, seg1 ? &mB1 : &mB0, &full[s], kk, n0 + half, zb, (uint16_t)3);
              tma_load_3d_mc(st + S::LO + A_BYTES + half * BK * 4, seg1 ? &mB1l : &mB0l, &full[s], kk, n0 + half, zb, (uint16_t)3);
            } else {
            tma_load_3d(st + A_BYTES, seg1 ? &mB1 : &mB0, &full[s], kk, n0, (seg1 ? a.zB1 : a.zB0) ? z : 0);
            if (TS && a.b_presplit)   // the lo tile lands where the in-kernel split would have put it
              tma_load_3d(st + S::LO + A_BYTES, seg1 ? &mB1l : &mB0l, &full[s], kk, n0, (seg1 ? a.zB1 : a.zB0) ? z : 0);
            }
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    {
      // kind::tf32, fp32 accumulate, M = 128, N = BN; bits 15/16 = MN-major A/B (cute::UMMA::InstrDescriptor)
      // (an A operand in tensor memory is always rows x K: only B keeps the MN-major bit in the TS weight-gradient form)
      constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (MNMAJOR ? ((TS ? 2u : 3u) << 15) : 0u) |
                                 (uint32_t(BN >> 3) << 17) | (uint32_t(BM >> 4) << 24);
      // One thread issues every MMA, so the instruction latency of THIS loop paces the tensor pipe (a
      // runtime modulo for the accumulator rotation alone cost 30 % of the mainloop): descriptors are a
      // constant high word plus a 14-bit address in the low word, the rotation is a wrapping counter.
      const uint32_t tmem_corr = tmem_base + NACC * BN;
      const uint64_t d0 = MNMAJOR ? make_sdesc_mn<BK>(0u) : make_sdesc<BK>(0u);
      const uint32_t dhi = static_cast<uint32_t>(d0 >> 32), dlo = static_cast<uint32_t>(d0);
      constexpr uint32_t kStep = MNMAJOR ? (1024u >> 4) : (32u >> 4);   // next k8 block inside a stage
      const uint32_t rot_end = tmem_base + (uint32_t)nacc_rt * BN;
      uint32_t tmem_main = tmem_base, acc_c = 0, fresh = (uint32_t)nacc_rt;   // `fresh` mains not written yet
      const bool three = !NS && a.passes == 3;
      // single TF32 pass on the raw fp32 tiles (the tensor core ignores the low 13 mantissa bits): nothing to split, the
      // MMAs consume the TMA tiles directly -- TMA -> full barrier -> MMA, no thread in between
      uint64_t* ready = (NS || (!TS && a.passes == 1 && a.raw_hi)) ? full : split;
      // The tensor pipe accepts MMAs about as fast as it executes them (measured: 6 issues take 360-390 cycles, and
      // with every wait removed a stage still took 690 cycles of which 330 were this loop's own bookkeeping), so
      // whatever this warp does between two issue bursts is tensor-pipe idle time.  Everything is strength-reduced:
      // stage index / phase / descriptor word / TMEM slot are wrapping counters, the K tail is a running remainder,
      // the issuing lane is elected once.
      const bool leader = elect_one();
      const uint32_t lo0 = dlo + ((smem_u32(smem) & 0x3FFFFu) >> 4);          // descriptor low word of stage 0's A tile
      uint32_t st_lo = lo0;
      int s = 0;
      uint32_t ph = 0;
      uint32_t ta_slot = tmem_base + A_BASE;                                     // TS: first column of the current A slot
      const uint32_t ta_end = tmem_base + A_BASE + kASlots * 2 * BK;
      int aslot = 0;
      int g = kc_begin;                                                           // global chunk index (K-major: over both segments)
      int kleft = MNMAJOR ? (int)min((long long)a.K0 - (long long)kc_begin * BK, (long long)0x7fffffff)
                          : (g >= nk0 ? a.K1 - (g - nk0) * BK : a.K0 - g * BK);
      // one stage worth of MMAs (3 per k8 step): lo*hi and hi*lo into corr, hi*hi into the rotating main accumulator
      auto issue_stage = [&](uint32_t lo, uint32_t ta, int nks) {
        uint32_t la_hi = lo, lb_hi = lo + (A_BYTES >> 4), la_lo = lo + (S::LO >> 4), lb_lo = lo + ((A_BYTES + S::LO) >> 4);
#pragma unroll
        for (int ks = 0; ks < BK / 8; ++ks) {
          if (ks < nks) {
            const uint64_t a_hi = (static_cast<uint64_t>(dhi) << 32) | la_hi, b_hi = (static_cast<uint64_t>(dhi) << 32) | lb_hi;
            const uint64_t a_lo = (static_cast<uint64_t>(dhi) << 32) | la_lo, b_lo = (static_cast<uint64_t>(dhi) << 32) | lb_lo;
            uint32_t acc_m = fresh == 0 ? 1u : 0u;
            if constexpr (TS) {
              const uint32_t ta_hi = ta + (uint32_t)ks * 8, ta_lo = ta_hi + BK;
              if (leader) {
                if (three) {
                  umma_tf32_ts(tmem_corr, ta_lo, b_hi, idesc, acc_c);
                  umma_tf32_ts(tmem_corr, ta_hi, b_lo, idesc, 1u);
                }
                umma_tf32_ts(tmem_main, ta_hi, b_hi, idesc, acc_m);
              }
            } else
            if (leader) {
              if (three) {
                umma_tf32(tmem_corr, a_lo, b_hi, idesc, acc_c);
                umma_tf32(tmem_corr, a_hi, b_lo, idesc, 1u);
              }
              umma_tf32(tmem_main, a_hi, b_hi, idesc, acc_m);
            }
            acc_c = 1;
            if (fresh) --fresh;
            tmem_main += BN;
            if (tmem_main == rot_end) tmem_main = tmem_base;
            la_hi += kStep; lb_hi += kStep; la_lo += kStep; lb_lo += kStep;
          }
        }
      };
      // two stages per loop trip: both barriers are awaited first, then 12 MMAs go out back to back, so the fixed
      // cost of a trip (barrier poll, warp sync, proxy fence, commits, counters) is paid once per 768 MMA cycles
      for (int it = 0; it < nk; it += 2) {
        const bool two = it + 1 < nk;
        const int nks0 = kleft >= BK ? BK / 8 : (kleft + 7) >> 3;
        int kleft1 = kleft - BK;
        if (!MNMAJOR && g + 1 == nk0) kleft1 = a.K1;
        const int nks1 = kleft1 >= BK ? BK / 8 : (kleft1 + 7) >> 3;
        const bool wrap = s + 1 == STAGES;
        const int s1 = wrap ? 0 : s + 1;
        const uint32_t ph1 = wrap ? ph ^ 1u : ph;
        const uint32_t lo1 = wrap ? lo0 : st_lo + (S::STAGE >> 4);
        const bool awrap = aslot + 1 == kASlots;
        const int aslot1 = awrap ? 0 : aslot + 1;
        const uint32_t ta1 = awrap ? tmem_base + A_BASE : ta_slot + 2 * BK;
        if (lane == 0) {
          mbar_wait(&ready[s], ph);
          if (two) mbar_wait(&ready[s1], ph1);
        }
        __syncwarp();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        issue_stage(st_lo, ta_slot, nks0);
        if (leader) {
          if constexpr (MC) umma_commit_mc(&empty[s], (uint16_t)3);   // both CTAs of the pair refill this stage's weights
          else umma_commit(&empty[s]);   // frees the stage once the MMAs above have read it
          if (TS) umma_commit(&aempty[aslot]);
        }
        if (two) {
          issue_stage(lo1, ta1, nks1);
          if (leader) {
            if constexpr (MC) umma_commit_mc(&empty[s1], (uint16_t)3);
            else umma_commit(&empty[s1]);
            if (TS) umma_commit(&aempty[aslot1]);
          }
        }
        // advance two stages
        kleft = kleft1 - BK;
        g += 2;
        if (!MNMAJOR && g == nk0) kleft = a.K1;
        s = s1; ph = ph1; st_lo = lo1; aslot = aslot1; ta_slot = ta1;
        {
          const bool w2 = s + 1 == STAGES;
          st_lo = w2 ? lo0 : st_lo + (S::STAGE >> 4);
          if (w2) ph ^= 1u;
          s = w2 ? 0 : s + 1;
          const bool aw2 = aslot + 1 == kASlots;
          ta_slot = aw2 ? tmem_base + A_BASE : ta_slot + 2 * BK;
          aslot = aw2 ? 0 : aslot + 1;
        }
      }
      (void)ta_end;
      if (elect_one()) umma_commit(accbar);        // accumulator complete
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------------ splitters, then epilogue
    const int t = (threadIdx.x - 64) & 127;       // 0..127 inside the splitter group
    const int grp = (threadIdx.x - 64) >> 7;      // splitter group: takes the stages with it % NSG == grp
    const bool nosplit = NS || (!TS && a.passes == 1 && a.raw_hi);
    constexpr int kMaxXc = 16, kVecXc = 8;   // (the 128-bit path keeps 4 rows x kVecXc columns in registers)
    const bool ride = MNMAJOR && TS && blockIdx.y == 0 && (a.colsum != nullptr || a.skinny != nullptr);
    float csum = 0.f, sk[kMaxXc];
#pragma unroll
    for (int j = 0; j < kMaxXc; ++j) sk[j] = 0.f;
    // riding skinny reduction: lane l fetches float4 (l & 1) of row (l >> 1) of the stage's xc tile one stage ahead
    // (any xc with at most 8 columns: 128-bit loads when its rows allow them, else up to four scalar loads per lane)
    const bool xc_pre = ride && a.skinny != nullptr && a.nxc <= 16 && BK == 16;
    const bool xc_wide = xc_pre && a.nxc > 8;     // columns 8..15: a second float4 per lane (C5: 16 x_cond columns)
    float4 xc_next = make_float4(0.f, 0.f, 0.f, 0.f), xc_cur = xc_next, xc_next2 = xc_next, xc_cur2 = xc_next, av_r[4];
    auto xc_fetch = [&](int it_, int half) -> float4 {
      const long long kr = min((long long)(kc_begin + it_) * BK + (lane >> 1), (long long)a.K0 - 1);
      const int j0 = 8 * half + 4 * (lane & 1);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (j0 >= a.nxc) return v;
      const float* xr = a.xc + (long long)z_o * a.sXc2 + kr * a.ldxc + j0;
      if (a.xc_al) return __ldg(reinterpret_cast<const float4*>(xr));
      v.x = __ldg(xr);
      if (j0 + 1 < a.nxc) v.y = __ldg(xr + 1);
      if (j0 + 2 < a.nxc) v.z = __ldg(xr + 2);
      if (j0 + 3 < a.nxc) v.w = __ldg(xr + 3);
      return v;
    };
    if (xc_pre && grp < nk) {
      xc_next = xc_fetch(grp, 0);
      if (xc_wide) xc_next2 = xc_fetch(grp, 1);
    }
    for (int it = grp; it < nk && !nosplit; it += NSG) {
      const int s = it % STAGES, use = it / STAGES;
      if (xc_pre) {
        xc_cur = xc_next; xc_cur2 = xc_next2;
        if (it + NSG < nk) {                               // in flight across the wait below
          xc_next = xc_fetch(it + NSG, 0);
          if (xc_wide) xc_next2 = xc_fetch(it + NSG, 1);
        }
      }
      mbar_wait(&full[s], use & 1);
      if (threadIdx.x == 64 && it == 0) NF_TS(2);   // first operand tile landed
      float4* raw = reinterpret_cast<float4*>(smem + s * S::STAGE);
      float4* lo = reinterpret_cast<float4*>(smem + s * S::STAGE + S::LO);
      if constexpr (TS) {
        // B tile: split in shared memory as in the SS form; A tile: this thread's row -> tensor memory
        constexpr int A4 = A_BYTES / 16, PERB = S::B_BYTES / 16 / 128;
        const int r = (warp & 3) * 32 + lane;                        // tile row = TMEM lane of this thread
        float4 av[4], bv[PERB];
        if constexpr (MNMAJOR) {
          // weight-gradient form: the A tile is four unswizzled [BK rows][32 columns] boxes; this thread's tile row is
          // column `lane` of box (warp & 3): element k at box + k * 32 + lane (consecutive lanes, no bank conflicts)
          const float* boxf = reinterpret_cast<const float*>(raw) + (warp & 3) * (BOX / 4) + lane;
#pragma unroll
          for (int c = 0; c < 4; ++c)
            av[c] = make_float4(boxf[(4 * c) * 32], boxf[(4 * c + 1) * 32], boxf[(4 * c + 2) * 32], boxf[(4 * c + 3) * 32]);
          if (ride) {   // rows past K are zero-filled by the TMA, so they add nothing
            csum += ((av[0].x + av[0].y) + (av[0].z + av[0].w)) + ((av[1].x + av[1].y) + (av[1].z + av[1].w)) +
                    ((av[2].x + av[2].y) + (av[2].z + av[2].w)) + ((av[3].x + av[3].y) + (av[3].z + av[3].w));
            if (a.skinny) {
              const long long k0 = (long long)(kc_begin + it) * BK;
              const float* xbase = a.xc + (long long)z_o * a.sXc2;
              if (xc_pre) {
                // (done after this stage's operands have been handed to the MMA warp, see the end of the iteration)
#pragma unroll
                for (int c = 0; c < 4; ++c) av_r[c] = av[c];
              } else if (a.xc_vec) {
                // rows of xc as 128-bit loads (same address in every lane: one broadcast transaction each), four rows
                // in flight at a time -- as scalar loads this rode at 3 100 cycles per stage (ncu, C2a batched dW1)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                  const float e4[4] = {av[c].x, av[c].y, av[c].z, av[c].w};
                  float4 xv[4][kVecXc / 4];
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    const long long kr = min(k0 + 4 * c + e, (long long)a.K0 - 1);
                    const float4* xr = reinterpret_cast<const float4*>(xbase + kr * a.ldxc);
#pragma unroll
                    for (int j4 = 0; j4 < kVecXc / 4; ++j4)
                      xv[e][j4] = 4 * j4 < a.nxc ? __ldg(xr + j4) : make_float4(0.f, 0.f, 0.f, 0.f);
                  }
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
#pragma unroll
                    for (int j4 = 0; j4 < kVecXc / 4; ++j4) {
                      if (4 * j4 < a.nxc) {
                        sk[4 * j4] = fmaf(e4[e], xv[e][j4].x, sk[4 * j4]); sk[4 * j4 + 1] = fmaf(e4[e], xv[e][j4].y, sk[4 * j4 + 1]);
                        sk[4 * j4 + 2] = fmaf(e4[e], xv[e][j4].z, sk[4 * j4 + 2]); sk[4 * j4 + 3] = fmaf(e4[e], xv[e][j4].w, sk[4 * j4 + 3]);
                      }
                    }
                  }
                }
              } else {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                  const float e4[4] = {av[c].x, av[c].y, av[c].z, av[c].w};
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    const long long kr = min(k0 + 4 * c + e, (long long)a.K0 - 1);
                    const float* xr = xbase + kr * a.ldxc;      // same address in every lane: one broadcast transaction
#pragma unroll
                    for (int j = 0; j < kMaxXc; ++j)
                      if (j < a.nxc) sk[j] = fmaf(e4[e], __ldg(xr + j), sk[j]);
                  }
                }
              }
            }
          }
        } else {
          const int sw = (r >> 1) & 3;                               // SWIZZLE_64B: 16-byte chunk c sits at c ^ ((row >> 1) & 3)
#pragma unroll
          for (int c = 0; c < 4; ++c) av[c] = raw[r * 4 + (c ^ sw)];
        }
        const bool split_b = MNMAJOR || !a.b_presplit;              // pre-split weights arrive as two TMA tiles
#pragma unroll
        for (int k = 0; k < PERB; ++k) bv[k] = split_b ? raw[A4 + t + 128 * k] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < PERB; ++k) {
          if (!split_b) break;
          const int i = A4 + t + 128 * k;
          float4 h, l;
          if (a.raw_hi) {
            h.x = __uint_as_float(__float_as_uint(bv[k].x) & 0xffffe000u); h.y = __uint_as_float(__float_as_uint(bv[k].y) & 0xffffe000u);
            h.z = __uint_as_float(__float_as_uint(bv[k].z) & 0xffffe000u); h.w = __uint_as_float(__float_as_uint(bv[k].w) & 0xffffe000u);
          } else {
            h.x = rna_tf32(bv[k].x); h.y = rna_tf32(bv[k].y); h.z = rna_tf32(bv[k].z); h.w = rna_tf32(bv[k].w);
            raw[i] = h;
          }
          l.x = bv[k].x - h.x; l.y = bv[k].y - h.y; l.z = bv[k].z - h.z; l.w = bv[k].w - h.w;
          lo[i] = l;
        }
        uint32_t hi16[16], lo16[16];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float xs[4] = {av[c].x, av[c].y, av[c].z, av[c].w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float h = a.raw_hi ? __uint_as_float(__float_as_uint(xs[e]) & 0xffffe000u) : rna_tf32(xs[e]);
            hi16[4 * c + e] = __float_as_uint(h);
            lo16[4 * c + e] = __float_as_uint(xs[e] - h);
          }
        }
        const int slot = it % kASlots;
        if (it >= kASlots) {                                         // the MMAs of iteration it - 4 have read this slot
          if (lane == 0) mbar_wait(&aempty[slot], ((it / kASlots) & 1) ^ 1);
          __syncwarp();
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        const uint32_t ta = tmem_base + A_BASE + (uint32_t)slot * 2 * BK + (static_cast<uint32_t>((warp & 3) * 32) << 16);
        tmem_st16(ta, hi16);
        tmem_st16(ta + BK, lo16);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        if (split_b) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> UMMA reads
      } else {
      // every load of the stage is issued before the first store: the in-place/aliased stores otherwise
      // serialise load -> store -> load and the splitters pace the MMAs (measured 800 cycles per stage
      // against a 384-cycle MMA floor, tools/tc_timeline.py)
      constexpr int PER = S::HALF / 16 / 128;
      static_assert(S::HALF % (16 * 128) == 0, "tile bytes must be a multiple of 2 KB");
      float4 v[PER];
#pragma unroll
      for (int k = 0; k < PER; ++k) v[k] = raw[t + 128 * k];
#pragma unroll
      for (int k = 0; k < PER; ++k) {
        const int i = t + 128 * k;
        float4 h, l;
        if (a.raw_hi) {
          h.x = __uint_as_float(__float_as_uint(v[k].x) & 0xffffe000u); h.y = __uint_as_float(__float_as_uint(v[k].y) & 0xffffe000u);
          h.z = __uint_as_float(__float_as_uint(v[k].z) & 0xffffe000u); h.w = __uint_as_float(__float_as_uint(v[k].w) & 0xffffe000u);
        } else {
          h.x = rna_tf32(v[k].x); h.y = rna_tf32(v[k].y); h.z = rna_tf32(v[k].z); h.w = rna_tf32(v[k].w);
          raw[i] = h;
        }
        l.x = v[k].x - h.x; l.y = v[k].y - h.y; l.z = v[k].z - h.z; l.w = v[k].w - h.w;
        lo[i] = l;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> UMMA reads
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&split[s]);
      if constexpr (TS && MNMAJOR) {
        if (xc_pre) {   // riding skinny reduction, in the shadow of this stage's MMAs (the operands are already handed over)
        // The stage's xc tile (BK rows x <= 8 columns = 32 float4) was fetched one float4 per lane BEFORE the
        // wait for the operand tile (xc_cur); the warp passes it round through its 512 B (1 KB for 9..16 columns) of shared memory and
        // every lane reads the 16 rows back as broadcasts.  (Loading the rows after the wait -- 16 dependent 128-bit
        // loads in four rounds -- made the main loop of C2a's batched dW1 GEMM 64.6 K cycles per CTA; this form: 46.9 K;
        // without the riding reduction: 27 K.  What-if runs: the prefetch load alone costs 6 K, the FMAs 9 K, the
        // hand-over through shared memory 9 K -- the two splitter groups are a latency chain, every instruction added
        // to an iteration delays the group's next stage.  A dedicated rider warp is the next step, not taken.)
        float4* xs = reinterpret_cast<float4*>(smem + S::BARS + 512) + (warp - 2) * 64;
        float4* xs2 = xs + 32;                          // columns 8..15
        xs[lane] = xc_cur;
        if (xc_wide) xs2[lane] = xc_cur2;
        __syncwarp();
        // (two batches of eight broadcast loads, all issued before the FMAs that use them)
        const float ar[16] = {av_r[0].x, av_r[0].y, av_r[0].z, av_r[0].w, av_r[1].x, av_r[1].y, av_r[1].z, av_r[1].w,
                              av_r[2].x, av_r[2].y, av_r[2].z, av_r[2].w, av_r[3].x, av_r[3].y, av_r[3].z, av_r[3].w};
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float4 v[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] = xs[2 * (8 * h + e)];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            sk[0] = fmaf(ar[8 * h + e], v[e].x, sk[0]); sk[1] = fmaf(ar[8 * h + e], v[e].y, sk[1]);
            sk[2] = fmaf(ar[8 * h + e], v[e].z, sk[2]); sk[3] = fmaf(ar[8 * h + e], v[e].w, sk[3]);
          }
          if (a.nxc > 4) {
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = xs[2 * (8 * h + e) + 1];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              sk[4] = fmaf(ar[8 * h + e], v[e].x, sk[4]); sk[5] = fmaf(ar[8 * h + e], v[e].y, sk[5]);
              sk[6] = fmaf(ar[8 * h + e], v[e].z, sk[6]); sk[7] = fmaf(ar[8 * h + e], v[e].w, sk[7]);
            }
          }
          if (xc_wide) {
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = xs2[2 * (8 * h + e)];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              sk[8] = fmaf(ar[8 * h + e], v[e].x, sk[8]); sk[9] = fmaf(ar[8 * h + e], v[e].y, sk[9]);
              sk[10] = fmaf(ar[8 * h + e], v[e].z, sk[10]); sk[11] = fmaf(ar[8 * h + e], v[e].w, sk[11]);
            }
            if (a.nxc > 12) {
#pragma unroll
              for (int e = 0; e < 8; ++e) v[e] = xs2[2 * (8 * h + e) + 1];
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                sk[12] = fmaf(ar[8 * h + e], v[e].x, sk[12]); sk[13] = fmaf(ar[8 * h + e], v[e].y, sk[13]);
                sk[14] = fmaf(ar[8 * h + e], v[e].z, sk[14]); sk[15] = fmaf(ar[8 * h + e], v[e].w, sk[15]);
              }
            }
          }
        }
        __syncwarp();
        }
      }
    }
    if (ride) {   // one atomic per output element, CTA and splitter group
      const int mrow = m0 + (warp & 3) * 32 + lane;
      if (mrow < a.M) {
        if (a.colsum) atomicAdd(a.colsum + (long long)z_o * a.sColsum2 + (long long)z_i * a.sColsum + mrow, csum);
        if (a.skinny) {
          float* o = a.skinny + (long long)z_o * a.sSk2 + (long long)z_i * a.sSk + (long long)mrow * a.ldsk;
#pragma unroll
          for (int j = 0; j < kMaxXc; ++j)
            if (j < a.nxc) atomicAdd(o + j, sk[j]);
        }
      }
    }
    // epilogue: this warp may only touch TMEM lanes 32*(warp%4) .. +31; thread = one output row
    if (threadIdx.x == 64) NF_TS(3);   // last tile split
    mbar_wait(accbar, 0);
    if (threadIdx.x == 64) NF_TS(4);   // accumulators complete
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int q = warp & 3;
    const int row = m0 + q * 32 + lane;
    if constexpr (EPI != EPI_NONE) {
      // ------------------------------------------------------------------ row-wise epilogue over the cluster
      // The cluster (1, ncl, 1) spans the whole output row: this thread holds its row's BN columns in
      // registers, exchanges two partial statistics with the ncl - 1 peer CTAs through distributed
      // shared memory (push + remote mbarrier arrive, one round) and finishes the row-wise operation:
      //   EPI_LN_FWD: x_hat = LayerNorm_noaffine(LeakyReLU(acc + bias)), rstd, sign mask   (torch_fcnf.py:75-76)
      //   EPI_LN_BWD: du = LeakyReLU'(mask) rstd (g - mean(g) - x_hat mean(g x_hat)),  g = acc
      const int rit = q * 32 + lane;
      const bool rowok = row < a.M;
      // the NSG splitter groups share the tile's columns: this thread owns columns [cg0, cg0 + CW) of its row (half
      // the registers per thread, which is what lets the 320-thread kernel keep the row in registers without spills)
      constexpr int CW = BN / NSG;
      const int gq = NSG > 1 ? grp : 0;
      const int cg0 = gq * CW;
      const bool tail = NSG == 1 && a.tail;
      float v[CW];
#pragma unroll
      for (int c0 = 0; c0 < CW; c0 += 32) {
        float r[32];
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(cg0 + c0);
        tmem_ld_sum<BN, NACC>(taddr, nacc_used, with_corr, r);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[c0 + j] = r[j];
      }
      if (threadIdx.x == 64) NF_TS(5);   // row in registers
      float2* xchg = reinterpret_cast<float2*>(smem + S::XCHG);   // [rank][row in tile]
      const int ncl = a.ncl;
      // cluster (1, ncl, 1): rank = N tile; with the fused tail (1, ncl, 2): rank = N tile + ncl * net, and the
      // LayerNorm exchange stays inside the net's ncl CTAs
      const uint32_t rank_base = (EPI == EPI_LN_FWD && a.tail) ? (uint32_t)z * (uint32_t)ncl : 0u;
      const uint32_t my_rank = cluster_ctarank() - rank_base;
      float* sxh = reinterpret_cast<float*>(smem) + (gq * 4 + q) * (32 * (CW + 4));   // bwd: this warp's x_hat tile
      float* sx = reinterpret_cast<float*>(smem) + NSG * 4 * 32 * (CW + 4) + (gq * 4 + q) * kXWarpFloats;   // 32 x 32 transposition buffer
      float* cwarp = a.C + c_off + (long long)(m0 + q * 32) * a.ldc + n0 + cg0;
      const int rows_valid_w = a.M - (m0 + q * 32);
      // TMA-store boxes (4 KB each, 1024-byte aligned): after the x_hat tiles / transposition buffers
      uint8_t* boxes = smem + ((NSG * 4 * 32 * (CW + 4) * 4 + NSG * 4 * kXWarpFloats * 4 + 1023) & ~1023) +
                       (q * (BN / 32) + (cg0 >> 5)) * 4096;
      float p0, p1;            // the two partial statistics of this thread's CW columns
      uint32_t mbits[CW / 32];
      if constexpr (EPI == EPI_LN_FWD) {
        const float* bias = a.bias ? a.bias + (long long)z * a.sBias + n0 + cg0 : nullptr;
        float sum = 0.f;
#pragma unroll
        for (int c = 0; c < CW; ++c) {
          const float pre = v[c] + (bias ? __ldg(bias + c) : 0.f);
          if ((c & 31) == 0) mbits[c >> 5] = 0u;
          if (pre > 0.f) mbits[c >> 5] |= 1u << (c & 31);
          v[c] = pre > 0.f ? pre : kLeakySlope * pre;
          sum += v[c];
        }
        p0 = sum * (1.0f / CW);                     // local mean
        float m2 = 0.f;
        const float sh = a.fast_var ? 0.f : p0;
#pragma unroll
        for (int c = 0; c < CW; ++c) { const float d = v[c] - sh; m2 = fmaf(d, d, m2); }
        p1 = m2;                                    // local centred sum of squares (fast_var: plain sum of squares)
      } else {
        // x_hat tile of this warp's 32 rows -> shared memory with coalesced 128-bit loads (the operand
        // stages are free now); both passes then read the lane's own row from there
        const float* xw = a.xh + (long long)z * a.sXh + (long long)(m0 + q * 32) * a.ldxh + n0 + cg0;
        const int rows_valid = a.M - (m0 + q * 32);
#pragma unroll 8
        for (int idx = lane; idx < 32 * (CW / 4); idx += 32) {
          const int r = idx / (CW / 4), c4 = idx % (CW / 4);
          const float4 t = r < rows_valid ? __ldg(reinterpret_cast<const float4*>(xw + (long long)r * a.ldxh) + c4)
                                          : make_float4(0.f, 0.f, 0.f, 0.f);
          *reinterpret_cast<float4*>(sxh + r * (CW + 4) + 4 * c4) = t;
        }
        __syncwarp();
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int c = 0; c < CW; c += 4) {
          const float4 x = *reinterpret_cast<const float4*>(sxh + lane * (CW + 4) + c);
          s1 += (v[c] + v[c + 1]) + (v[c + 2] + v[c + 3]);
          s2 = fmaf(v[c], x.x, fmaf(v[c + 1], x.y, fmaf(v[c + 2], x.z, fmaf(v[c + 3], x.w, s2))));
        }
        p0 = s1; p1 = s2;
      }
      // push to every CTA of the cluster (own included), then wait for everyone's partials
      const uint32_t slot = smem_u32(&xchg[(my_rank * NSG + gq) * 128 + rit]);   // [rank][group][row in tile]
      const uint32_t xb = smem_u32(xbar);
      for (int pr = 0; pr < ncl; ++pr) {
        st_cluster_v2(mapa_u32(slot, rank_base + (uint32_t)pr), p0, p1);
        mbar_arrive_remote(mapa_u32(xb, rank_base + (uint32_t)pr));
      }
      mbar_wait_cluster(xbar, 0);
      if (threadIdx.x == 64) NF_TS(6);   // statistics exchanged
      float* crow = a.C + c_off + (long long)row * a.ldc + n0 + cg0;
      if constexpr (EPI == EPI_LN_FWD) {
        // combine equal-sized groups (Chan et al.): mean = avg(mean_i), M2 = sum M2_i + BN sum (mean_i - mean)^2
        float mean = 0.f;
        for (int pr = 0; pr < ncl * NSG; ++pr) mean += xchg[pr * 128 + rit].x;
        mean /= (float)(ncl * NSG);
        float m2 = 0.f;
        for (int pr = 0; pr < ncl * NSG; ++pr) {
          const float2 s = xchg[pr * 128 + rit];
          const float d = s.x - mean;
          m2 += a.fast_var ? s.y : s.y + (float)CW * d * d;
        }
        float var = m2 / (float)(ncl * BN);
        if (a.fast_var) var = fmaxf(var - mean * mean, 0.f);
        const float rstd = rsqrtf(var + a.eps);
        // fused tail: this CTA's slice of the folded last Linear, W3f[net][i][n0 .. n0 + BN), in shared memory
        const int dt = a.t_D >> 1;
        float* sW3 = reinterpret_cast<float*>(smem + 160 * 1024);          // [4][BN]
        float* sWm = sW3 + 4 * BN;                                          // [D][D] (leader)
        float part[4] = {0.f, 0.f, 0.f, 0.f};
        if (tail) {
          const int te = threadIdx.x - 64;
          for (int idx = te; idx < 4 * BN; idx += 128) {
            const int i = idx / BN, c = idx - i * BN;
            sW3[idx] = i < dt ? __ldg(a.t_W3f + (long long)z * a.t_w_net + (long long)i * a.N + n0 + c) : 0.f;
          }
          if (cluster_ctarank() == 0 && a.t_Wmat)
            for (int idx = te; idx < a.t_D * a.t_D; idx += 128) sWm[idx] = __ldg(a.t_Wmat + idx);
          asm volatile("bar.sync 1, 128;" ::: "memory");   // the four epilogue warps
        }
#pragma unroll
        for (int c0 = 0; c0 < CW; c0 += 32) {
          float o[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) o[j] = (v[c0 + j] - mean) * rstd;
          if (tail) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 w4 = *reinterpret_cast<const float4*>(sW3 + i * BN + c0 + j);
                part[i] = fmaf(o[j], w4.x, fmaf(o[j + 1], w4.y, fmaf(o[j + 2], w4.z, fmaf(o[j + 3], w4.w, part[i]))));
              }
            }
          }
          if (!a.store_c) continue;
          if (a.tma_store) {
            uint8_t* box = boxes + (c0 >> 5) * 4096;
            box_put(box, lane, o);
            // one store per 32-column chunk: the TMA drains the box while the next chunk is computed (issuing
            // all stores of the tile at the end behind a single proxy fence measured 20 % slower)
            if (elect_one()) { tma_store_3d(&mC, box, n0 + cg0 + c0, m0 + q * 32, z); tma_store_commit(); }
            __syncwarp();
          } else {
            xpose_put(sx, lane, o);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int rr = (lane >> 3) + 4 * i;
              if (rr < rows_valid_w)
                *reinterpret_cast<float4*>(cwarp + (long long)rr * a.ldc + c0 + 4 * (lane & 7)) = xpose_get(sx, lane, i);
            }
            __syncwarp();
          }
        }
        if (rowok) {
          if (a.mask) {
            uint32_t* mk = a.mask + (long long)z * a.sMask + (long long)row * a.mw + ((n0 + cg0) >> 5);
#pragma unroll
            for (int w = 0; w < CW / 32; ++w) mk[w] = mbits[w];
          }
          if (a.rstd && my_rank == 0 && cg0 == 0) a.rstd[(long long)z * a.sRstd + row] = rstd;
        }
        if (tail) {
          // partial last-Linear outputs of this CTA's columns -> the cluster's leader CTA (rank 0)
          float4* tpart = reinterpret_cast<float4*>(smem + S::TPART);     // [rank][row in tile]
          const uint32_t crank = cluster_ctarank();
          const uint32_t dst = mapa_u32(smem_u32(&tpart[crank * 128 + rit]), 0u);
          asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "f"(part[0]), "f"(part[1]),
                       "f"(part[2]), "f"(part[3]) : "memory");
          mbar_arrive_remote(mapa_u32(smem_u32(tbar), 0u));
          if (crank == 0) {
            mbar_wait_cluster(tbar, 0);
            // t = net 0 (ranks 0 .. ncl-1), s = net 1 (ranks ncl .. 2 ncl - 1)     (torch_fcnf.py:99-104 / :110-115)
            float tv[4] = {0.f, 0.f, 0.f, 0.f}, sv[4] = {0.f, 0.f, 0.f, 0.f};
            for (int pr = 0; pr < ncl; ++pr) {
              const float4 pt = tpart[pr * 128 + rit], ps = tpart[(ncl + pr) * 128 + rit];
              tv[0] += pt.x; tv[1] += pt.y; tv[2] += pt.z; tv[3] += pt.w;
              sv[0] += ps.x; sv[1] += ps.y; sv[2] += ps.z; sv[3] += ps.w;
            }
            if (rowok) {
              const int D = a.t_D, dc = (D + 1) >> 1;
              constexpr int MAXD = 9;
              float xin[MAXD], zv[MAXD];
#pragma unroll
              for (int j = 0; j < MAXD; ++j) xin[j] = j < D ? a.t_in[(long long)row * a.t_ld_in + j] : 0.f;
              float ssum = 0.f;
#pragma unroll
              for (int j = 0; j < MAXD; ++j) zv[j] = xin[j];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                if (i < dt) {
                  const float tt = tv[i] + a.t_b3f[i], ss = sv[i] + a.t_b3f[a.t_b_net + i];
                  sv[i] = ss;
                  ssum += ss;
#pragma unroll
                  for (int j = 0; j < MAXD; ++j)
                    if (j == dc + i) zv[j] = a.t_inverse ? fmaf(xin[j], expf(ss), tt) : (xin[j] - tt) * expf(-ss);
                }
              }
              if (!a.t_inverse) {
#pragma unroll
                for (int j = 0; j < MAXD; ++j)
                  if (j < D) a.t_out[(long long)row * a.t_ld_out + j] = zv[j];
                if (a.t_s_out) {
#pragma unroll
                  for (int i = 0; i < 4; ++i)
                    if (i < dt) a.t_s_out[(long long)row * dt + i] = sv[i];
                }
              }
              if (a.t_logdet) {
                const float cst = a.t_cdev ? *a.t_cdev : 0.f;
                a.t_logdet[row] += a.t_inverse ? (ssum - cst) : (cst - ssum);
              }
              if (a.t_Wmat) {
#pragma unroll
                for (int j = 0; j < MAXD; ++j) {
                  if (j < D) {
                    float o2 = 0.f;
#pragma unroll
                    for (int k = 0; k < MAXD; ++k)
                      if (k < D) o2 = fmaf(zv[k], sWm[k * D + j], o2);
                    a.t_out2[(long long)row * a.t_ld_out2 + j] = o2;
                  }
                }
              }
            }
          }
        }
      } else {
        float s1 = 0.f, s2 = 0.f;
        for (int pr = 0; pr < ncl * NSG; ++pr) { const float2 s = xchg[pr * 128 + rit]; s1 += s.x; s2 += s.y; }
        const float inv = 1.0f / (float)(ncl * BN);
        s1 *= inv; s2 *= inv;
        float rs_row = 0.f;
        if (rowok) {
          rs_row = a.rstd[(long long)z * a.sRstd + row];
          const uint32_t* mk = a.mask + (long long)z * a.sMask + (long long)row * a.mw + ((n0 + cg0) >> 5);
#pragma unroll
          for (int w = 0; w < CW / 32; ++w) mbits[w] = mk[w];
        }
#pragma unroll
        for (int c0 = 0; c0 < CW; c0 += 32) {
          float o[32];
          const uint32_t bits = mbits[c0 >> 5];
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 x = *reinterpret_cast<const float4*>(sxh + lane * (CW + 4) + c0 + j);
            o[j] = rs_row * (v[c0 + j] - s1 - x.x * s2) * (((bits >> j) & 1u) ? 1.0f : kLeakySlope);
            o[j + 1] = rs_row * (v[c0 + j + 1] - s1 - x.y * s2) * (((bits >> (j + 1)) & 1u) ? 1.0f : kLeakySlope);
            o[j + 2] = rs_row * (v[c0 + j + 2] - s1 - x.z * s2) * (((bits >> (j + 2)) & 1u) ? 1.0f : kLeakySlope);
            o[j + 3] = rs_row * (v[c0 + j + 3] - s1 - x.w * s2) * (((bits >> (j + 3)) & 1u) ? 1.0f : kLeakySlope);
          }
          if (a.tma_store) {
            uint8_t* box = boxes + (c0 >> 5) * 4096;
            box_put(box, lane, o);
            // one store per 32-column chunk: the TMA drains the box while the next chunk is computed (issuing
            // all stores of the tile at the end behind a single proxy fence measured 20 % slower)
            if (elect_one()) { tma_store_3d(&mC, box, n0 + cg0 + c0, m0 + q * 32, z); tma_store_commit(); }
            __syncwarp();
          } else {
            xpose_put(sx, lane, o);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int rr = (lane >> 3) + 4 * i;
              if (rr < rows_valid_w)
                *reinterpret_cast<float4*>(cwarp + (long long)rr * a.ldc + c0 + 4 * (lane & 7)) = xpose_get(sx, lane, i);
            }
            __syncwarp();
          }
        }
      }
      if (a.tma_store) tma_store_wait_all();   // (only the electing lanes have groups in flight)
    } else {
    const bool red = MNMAJOR || a.splitk > 1 || a.shared_c;   // partial tile: combine with red.global.add
    const float* bias = (a.bias && kc_begin == 0) ? a.bias + (long long)z * a.sBias : nullptr;
    const bool vec_ok = (a.ldc & 3) == 0 && ((reinterpret_cast<uintptr_t>(a.C) + (size_t)c_off * 4) & 15) == 0;
    float* crow = a.C + c_off + (long long)row * a.ldc;
    // (a shared-memory transposition for fully coalesced stores was measured SLOWER here: this epilogue is
    // bound by the 64 B/clk TMEM read path, not by the LSU; the fused epilogues keep the transposition)
    for (int c0 = 0; c0 < BN; c0 += 32) {
      if (n0 + c0 >= a.N) break;   // warp-uniform
      if (NSG > 1 && ((c0 >> 5) % NSG) != grp) continue;   // the splitter groups share the epilogue chunk by chunk
      float rf[32];
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(c0);
      tmem_ld_sum<BN, NACC>(taddr, nacc_used, with_corr, rf);
      if (threadIdx.x == 64 && c0 == 0) NF_TS(5);
      if (a.tma_store) {
        if (bias) {
#pragma unroll
          for (int j = 0; j < 32; ++j) rf[j] += (n0 + c0 + j < a.N) ? __ldg(bias + n0 + c0 + j) : 0.f;
        }
        uint8_t* box = smem + (q * (BN / 32) + (c0 >> 5)) * 4096;
        box_put(box, lane, rf);
        if (elect_one()) {
          if (red || a.accumulate) tma_reduce_add_3d(&mC, box, n0 + c0, m0 + q * 32, a.shared_c ? 0 : z);
          else tma_store_3d(&mC, box, n0 + c0, m0 + q * 32, z);
          tma_store_commit();
        }
        __syncwarp();
        if (threadIdx.x == 64 && c0 == 0) NF_TS(6);
        continue;
      }
      if (row < a.M) {
        const int cbase = n0 + c0;
        if (cbase + 32 <= a.N && vec_ok) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 v = make_float4(rf[j], rf[j + 1], rf[j + 2], rf[j + 3]);
            float4* dst = reinterpret_cast<float4*>(crow + cbase + j);
            if (red) {
              if (!MNMAJOR && bias) { v.x += bias[cbase + j]; v.y += bias[cbase + j + 1]; v.z += bias[cbase + j + 2]; v.w += bias[cbase + j + 3]; }
              red_add_v4(crow + cbase + j, v);
            } else {
              if (bias) { v.x += bias[cbase + j]; v.y += bias[cbase + j + 1]; v.z += bias[cbase + j + 2]; v.w += bias[cbase + j + 3]; }
              if (a.accumulate) { const float4 o = *dst; v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
              *dst = v;
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (cbase + j < a.N) {
              if (red) {
                atomicAdd(crow + cbase + j, rf[j] + ((!MNMAJOR && bias) ? bias[cbase + j] : 0.f));
              } else {
                float v = rf[j] + (bias ? bias[cbase + j] : 0.f);
                if (a.accumulate) v += crow[cbase + j];
                crow[cbase + j] = v;
              }
            }
          }
        }
      }
      if (threadIdx.x == 64 && c0 == 0) NF_TS(6);
    }
    if (a.tma_store) tma_store_wait_all();
    }  // plain epilogue
  }
  if (threadIdx.x == 64) NF_TS(7);     // outputs stored
#undef NF_TS
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (MC) cluster_sync_all();   // the peer may still multicast into this CTA's stages / arrive on its barriers
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// host side: tensor maps
// ---------------------------------------------------------------------------------------------
long long* g_tc_dbg = nullptr;   // set by nf_debug_tc_timeline

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// fp32 matrix with `inner` contiguous floats per row and `rows` rows (row pitch `ld` floats), optionally
// batched; box = 32 x box_rows x 1, SWIZZLE_128B, zero fill out of bounds.
int make_map(CUtensorMap* map, const float* base, long long inner, long long rows, long long ld, long long batch,
             long long bstride, int box_inner, int box_rows, CUtensorMapSwizzle swz) {
  EncodeTiledFn fn = encode_fn();
  NF_CHECK_ARG(fn != nullptr, NF_E_UNSUPPORTED, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t dims[3] = {(cuuint64_t)inner, (cuuint64_t)rows, (cuuint64_t)(bstride ? batch : 1)};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 4, (cuuint64_t)(bstride ? bstride : ld * rows) * 4};
  cuuint32_t box[3] = {(cuuint32_t)box_inner, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  NF_CHECK_ARG(r == CUDA_SUCCESS, NF_E_BADARG, "cuTensorMapEncodeTiled failed (%d): inner=%lld rows=%lld ld=%lld",
               (int)r, inner, rows, ld);
  return 0;
}

// output tensor map for the TMA-store epilogue: 32 x 32 fp32 boxes, SWIZZLE_128B; returns false when TMA cannot
// address C (the direct-store epilogue is used then)
bool make_c_map(CUtensorMap* map, float* C, long long N, long long M, long long ldc, long long batch, long long sC) {
  if (!option(NF_OPT_TMA_STORE) || !C || (reinterpret_cast<uintptr_t>(C) & 15u) != 0 || (ldc & 3) != 0 || (sC & 3) != 0) return false;
  return make_map(map, C, N, M, ldc, batch, sC, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B) == 0;
}

bool ok_operand(const float* p, long long ld, long long bstride) {
  return p && (reinterpret_cast<uintptr_t>(p) & 15u) == 0 && (ld & 3) == 0 && (bstride & 3) == 0;
}

template <int BN, int STAGES, bool MN, int BK>
int set_smem() {
  static PerDeviceOnce once;
  if (once.first()) {
    NF_CUDA(cudaFuncSetAttribute(k_tc_gemm<BN, STAGES, MN, BK>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 Smem<BN, STAGES, BK>::TOTAL));
  }
  return 0;
}

int pick_nacc() { return option(NF_OPT_NACC); }   // 0 = kernel default (all NACC accumulators)

template <int BN, int STAGES, int BK, bool TS = false, bool NS = false, int NSG_ = 0, bool MC = false>
int launch(const TcGemmP& p, cudaStream_t st) {
  using S = Smem<BN, STAGES, BK, TS, NS>;
  constexpr int NSG = NSG_ ? NSG_ : (TS ? 2 : 1);
  int r;
  {
    static PerDeviceOnce once;
    if (once.first()) {
      NF_CUDA(cudaFuncSetAttribute((k_tc_gemm<BN, STAGES, false, BK, EPI_NONE, TS, NSG, NS, MC>),
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
    }
  }
  CUtensorMap mA0, mA1, mB0, mB1, mB0l, mB1l;
  const CUtensorMapSwizzle swz = BK == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  // pre-split weight tables (hi / lo) feed the TS form directly
  const bool pre = TS && option(NF_OPT_PRESPLIT) != 0 && p.B0hi && p.B0lo && (p.K1 <= 0 || (p.B1hi && p.B1lo));
  if ((r = make_map(&mA0, p.A0, p.K0, p.M, p.lda0, p.batch, p.sA0, BK, BM, swz))) return r;
  constexpr int BROWS = MC ? BN / 2 : BN;     // rows of one weight box (multicast: each CTA of the pair loads half a tile)
  NF_CHECK_ARG(!MC || pre, NF_E_UNSUPPORTED, "tc_gemm: the multicast form needs pre-split weights");
  if ((r = make_map(&mB0, pre ? p.B0hi : p.B0, p.K0, p.N, p.ldb0, p.batch, p.sB0, BK, BROWS, swz))) return r;
  mB0l = mB0;
  if (pre && (r = make_map(&mB0l, p.B0lo, p.K0, p.N, p.ldb0, p.batch, p.sB0, BK, BROWS, swz))) return r;
  if (p.K1 > 0) {
    if ((r = make_map(&mA1, p.A1, p.K1, p.M, p.lda1, p.batch, p.sA1, BK, BM, swz))) return r;
    if ((r = make_map(&mB1, pre ? p.B1hi : p.B1, p.K1, p.N, p.ldb1, p.batch, p.sB1, BK, BROWS, swz))) return r;
    mB1l = mB1;
    if (pre && (r = make_map(&mB1l, p.B1lo, p.K1, p.N, p.ldb1, p.batch, p.sB1, BK, BROWS, swz))) return r;
  } else {
    mA1 = mA0; mB1 = mB0; mB1l = mB0l;
  }
  KArgs a{};
  a.C = p.C; a.bias = p.bias; a.M = p.M; a.N = p.N; a.K0 = p.K0; a.K1 = p.K1;
  a.ldc = p.ldc; a.sC = p.sC; a.sBias = p.sBias;
  a.zA0 = p.sA0 != 0; a.zA1 = p.sA1 != 0; a.zB0 = p.sB0 != 0; a.zB1 = p.sB1 != 0;
  a.accumulate = p.accumulate; a.passes = p.passes; a.raw_hi = p.raw_hi;
  // split the reduction over CTAs when the output is accumulated anyway (C already holds the base value) and
  // the tile grid leaves most of the 148 SMs idle: the dy GEMM of C2a (M = 4096, N = 64, K = 512) ran on 32
  // CTAs for 37 us
  const long long nchunks = cdiv(p.K0, BK) + cdiv(p.K1, BK);
  const long long tiles = cdiv(p.M, BM) * cdiv(p.N, BN) * p.batch;
  long long splitk = 1;
  if (p.accumulate && tiles * 2 <= 148 && nchunks >= 8) {
    splitk = std::min<long long>(148 / tiles, nchunks / 4);
    if (splitk < 1) splitk = 1;
  }
  const long long per = cdiv(nchunks, splitk);
  splitk = cdiv(nchunks, per);
  a.splitk = (int)splitk; a.chunks_per_split = (int)per;
  a.dbg = g_tc_dbg;
  a.nacc = pick_nacc();
  dim3 grid((unsigned)cdiv(p.M, BM), (unsigned)cdiv(p.N, BN), (unsigned)(p.batch * splitk));
  ProfScope prof(st, PROF_TC_GEMM, 2.0 * p.M * p.N * (double)(p.K0 + p.K1) * p.batch);
  CUtensorMap mC = mA0;
  a.shared_c = (p.batch > 1 && p.sC == 0) ? 1 : 0;
  NF_CHECK_ARG(!a.shared_c || (p.accumulate && p.epi.kind == TC_EPI_NONE && !p.bias), NF_E_BADARG,
               "tc_gemm: a batch that shares its output must accumulate (no bias, no fused epilogue)");
  a.tma_store = make_c_map(&mC, p.C, p.N, p.M, p.ldc, a.shared_c ? 1 : p.batch, p.sC) ? 1 : 0;
  a.b_presplit = pre ? 1 : 0;
  if constexpr (MC) {
    NF_CHECK_ARG(splitk == 1, NF_E_UNSUPPORTED, "tc_gemm: the multicast form does not split the reduction");
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((grid.x + 1u) & ~1u, grid.y, grid.z);      // whole pairs: an odd last row tile gets an idle partner
    cfg.blockDim = dim3(64 + 128 * NSG);
    cfg.dynamicSmemBytes = S::TOTAL;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 2 : 1;
    NF_CUDA(cudaLaunchKernelEx(&cfg, k_tc_gemm<BN, STAGES, false, BK, EPI_NONE, TS, NSG, NS, MC>, mA0, mA1, mB0, mB1, mB0l, mB1l, mC, a));
  } else {
  NF_LAUNCH((k_tc_gemm<BN, STAGES, false, BK, EPI_NONE, TS, NSG, NS, MC>), grid, 64 + 128 * NSG, S::TOTAL, st, mA0, mA1, mB0, mB1, mB0l, mB1l, mC, a);
  }
  NF_LAUNCHED();
  return 0;
}

// NSG = 2 (TS only): two splitter groups on alternating stages; the row-wise epilogue then splits the tile's columns
// between the groups (the fused flow tail needs the whole row in one thread and keeps NSG = 1)
template <int BN, int STAGES, int BK, int EPI, bool TS = false, int NSG = 1, bool NS = false>
int launch_epi(const TcGemmP& p, cudaStream_t st) {
  using S = Smem<BN, STAGES, BK, TS, NS>;
  static_assert(NSG == 1 || (TS && (BN / NSG) % 32 == 0), "two splitter groups need the TS form and 32-column slices");
  static PerDeviceOnce once;
  if (once.first()) {
    NF_CUDA(cudaFuncSetAttribute((k_tc_gemm<BN, STAGES, false, BK, EPI, TS, NSG, NS>), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 S::TOTAL_EPI));
  }
  int r;
  constexpr bool MC = false;   // (the fused-epilogue kernels keep their own cluster along N: no weight-tile multicast)
  CUtensorMap mA0, mA1, mB0, mB1, mB0l, mB1l;
  const CUtensorMapSwizzle swz = BK == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  // pre-split weight tables (hi / lo) feed the TS form directly
  const bool pre = TS && option(NF_OPT_PRESPLIT) != 0 && p.B0hi && p.B0lo && (p.K1 <= 0 || (p.B1hi && p.B1lo));
  if ((r = make_map(&mA0, p.A0, p.K0, p.M, p.lda0, p.batch, p.sA0, BK, BM, swz))) return r;
  constexpr int BROWS = MC ? BN / 2 : BN;     // rows of one weight box (multicast: each CTA of the pair loads half a tile)
  NF_CHECK_ARG(!MC || pre, NF_E_UNSUPPORTED, "tc_gemm: the multicast form needs pre-split weights");
  if ((r = make_map(&mB0, pre ? p.B0hi : p.B0, p.K0, p.N, p.ldb0, p.batch, p.sB0, BK, BROWS, swz))) return r;
  mB0l = mB0;
  if (pre && (r = make_map(&mB0l, p.B0lo, p.K0, p.N, p.ldb0, p.batch, p.sB0, BK, BROWS, swz))) return r;
  if (p.K1 > 0) {
    if ((r = make_map(&mA1, p.A1, p.K1, p.M, p.lda1, p.batch, p.sA1, BK, BM, swz))) return r;
    if ((r = make_map(&mB1, pre ? p.B1hi : p.B1, p.K1, p.N, p.ldb1, p.batch, p.sB1, BK, BROWS, swz))) return r;
    mB1l = mB1;
    if (pre && (r = make_map(&mB1l, p.B1lo, p.K1, p.N, p.ldb1, p.batch, p.sB1, BK, BROWS, swz))) return r;
  } else {
    mA1 = mA0; mB1 = mB0; mB1l = mB0l;
  }
  KArgs a{};
  a.C = p.C; a.bias = p.bias; a.M = p.M; a.N = p.N; a.K0 = p.K0; a.K1 = p.K1;
  a.ldc = p.ldc; a.sC = p.sC; a.sBias = p.sBias;
  a.zA0 = p.sA0 != 0; a.zA1 = p.sA1 != 0; a.zB0 = p.sB0 != 0; a.zB1 = p.sB1 != 0;
  a.accumulate = 0; a.passes = p.passes; a.raw_hi = p.raw_hi;
  a.splitk = 1; a.chunks_per_split = 0;
  a.dbg = g_tc_dbg;
  a.nacc = pick_nacc();
  a.rstd = p.epi.rstd; a.sRstd = p.epi.sRstd; a.mask = p.epi.mask; a.sMask = p.epi.sMask; a.mw = p.epi.mw;
  a.xh = p.epi.xh; a.ldxh = p.epi.ldxh; a.sXh = p.epi.sXh; a.eps = p.epi.eps; a.fast_var = p.epi.fast_var;
  a.ncl = p.N / BN;
  const TcEpi::Tail& t = p.epi.tail;
  a.tail = (EPI == EPI_LN_FWD && t.on) ? 1 : 0;
  a.store_c = a.tail ? t.store_c : 1;
  a.t_D = t.D; a.t_inverse = t.inverse; a.t_ld_in = t.ld_in; a.t_ld_out = t.ld_out; a.t_ld_out2 = t.ld_out2;
  a.t_W3f = t.W3f; a.t_w_net = t.w_net; a.t_b3f = t.b3f; a.t_b_net = t.b_net;
  a.t_in = t.in; a.t_cdev = t.cdev; a.t_Wmat = t.Wmat;
  a.t_out = t.out; a.t_out2 = t.out2; a.t_logdet = t.logdet; a.t_s_out = t.s_out;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)cdiv(p.M, BM), (unsigned)a.ncl, (unsigned)p.batch);
  cfg.blockDim = dim3(64 + 128 * NSG);
  cfg.dynamicSmemBytes = S::TOTAL_EPI;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1; attr[0].val.clusterDim.y = (unsigned)a.ncl; attr[0].val.clusterDim.z = a.tail ? 2u : 1u;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 2 : 1;
  ProfScope prof(st, PROF_TC_GEMM, 2.0 * p.M * p.N * (double)(p.K0 + p.K1) * p.batch);
  CUtensorMap mC = mA0;
  a.tma_store = make_c_map(&mC, p.C, p.N, p.M, p.ldc, p.batch, p.sC) ? 1 : 0;
  a.b_presplit = pre ? 1 : 0;
  NF_CUDA(cudaLaunchKernelEx(&cfg, k_tc_gemm<BN, STAGES, false, BK, EPI, TS, NSG, NS>, mA0, mA1, mB0, mB1, mB0l, mB1l, mC, a));
  NF_LAUNCHED();
  return 0;
}

template <int EPI>
int launch_epi_pick(const TcGemmP& p, cudaStream_t st) {
  if (p.passes == 1 && p.raw_hi && !(EPI == EPI_LN_FWD && p.epi.tail.on)) {   // single raw TF32 pass: no-split form
    if (p.N % 128 == 0 && p.N / 128 <= kMaxCluster) return launch_epi<128, 11, 16, EPI, false, 1, true>(p, st);
    return launch_epi<64, 14, 16, EPI, false, 1, true>(p, st);
  }
  const bool ts = option(NF_OPT_TS) != 0;
  const bool tail = EPI == EPI_LN_FWD && p.epi.tail.on;
  if (p.N % 128 == 0 && p.N / 128 <= kMaxCluster) {
    if (ts) {
      if constexpr (EPI == EPI_LN_FWD) { if (tail) return launch_epi<128, 8, 16, EPI, true, 1>(p, st); }
      return launch_epi<128, 8, 16, EPI, true, 2>(p, st);
    }
    return launch_epi<128, 6, 16, EPI>(p, st);
  }
  if (ts) {
    if constexpr (EPI == EPI_LN_FWD) { if (tail) return launch_epi<64, 12, 16, EPI, true, 1>(p, st); }
    return launch_epi<64, 12, 16, EPI, true, 2>(p, st);
  }
  return launch_epi<64, 8, 16, EPI>(p, st);
}

template <int BN, int STAGES, int BK, bool TS = false, bool NS = false, int NSG_ = 0>
int launch_tn(const TcGemmTnP& p, cudaStream_t st) {
  using S = Smem<BN, STAGES, BK, TS, NS>;
  constexpr int NSG = NSG_ ? NSG_ : (TS ? 2 : 1);
  constexpr int kXcBytes = TS ? 4 * NSG * 1024 : 0;   // per splitter warp: the stage's xc tile (riding skinny reduction), after the barrier block
  int r;
  {
    static PerDeviceOnce once;
    if (once.first()) {
      NF_CUDA(cudaFuncSetAttribute((k_tc_gemm<BN, STAGES, true, BK, EPI_NONE, TS, NSG, NS>),
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL + kXcBytes));
    }
  }
  CUtensorMap mA, mB;
  // operands are (K rows, M|N contiguous): inner = M|N, rows = K, 32 x BK boxes (TS: the A boxes are read by the
  // splitter threads, not by the tensor core, and stay unswizzled)
  if ((r = make_map(&mA, p.A, p.M, p.K, p.lda, p.batch, p.sA, 32, BK,
                    TS ? CU_TENSOR_MAP_SWIZZLE_NONE : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))) return r;
  if ((r = make_map(&mB, p.B, p.N, p.K, p.ldb, p.batch, p.sB, 32, BK, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))) return r;
  const long long nk = cdiv(p.K, BK);
  const long long tiles = cdiv(p.M, BM) * cdiv(p.N, BN) * p.batch;
  // one wave: as many splits as fit on the 148 SMs without spilling into a second, nearly empty wave
  long long splitk = p.splitk > 0 ? p.splitk : (tiles >= 148 ? 1 : 148 / tiles);
  // accumulate-truncation bound (see the kernel): at most kMaxRowsPerSplit batch rows per CTA
  constexpr long long kMaxRowsPerSplit = TS ? 1024 : 1536;   // two rotating accumulators in the TS form
  if (p.splitk <= 0 && splitk < cdiv(p.K, kMaxRowsPerSplit)) splitk = cdiv(p.K, kMaxRowsPerSplit);
  if (p.splitk <= 0 && p.inner > 0 && tiles < 148) {
    // batched launch over several blocks: pick the split with the fewest (waves x stages per CTA), a fixed ~8 stages of
    // prologue / epilogue per CTA included, among the splits that respect the rows-per-CTA bound
    const long long smin = std::max<long long>(1, cdiv(p.K, kMaxRowsPerSplit));
    long long best = smin, best_cost = -1;
    for (long long s_ = smin; s_ <= std::max<long long>(smin, 24); ++s_) {
      const long long cost = cdiv(tiles * s_, 148) * (cdiv(nk, s_) + 8);
      if (best_cost < 0 || cost < best_cost) { best = s_; best_cost = cost; }
    }
    splitk = best;
  }
  if (splitk > nk) splitk = nk;
  if (splitk < 1) splitk = 1;
  const long long per = cdiv(nk, splitk);
  splitk = cdiv(nk, per);                       // no empty split
  NF_CHECK_ARG(splitk * p.batch <= 65535, NF_E_BADARG, "tc_gemm_tn: grid too large");
  KArgs a{};
  a.C = p.C; a.bias = nullptr; a.M = p.M; a.N = p.N;
  a.K0 = (int)p.K; a.K1 = 0;
  a.ldc = p.ldc; a.sC = p.sC; a.sBias = 0;
  a.zA0 = p.sA != 0; a.zB0 = p.sB != 0; a.zA1 = a.zB1 = 0;
  a.accumulate = 1; a.passes = p.passes; a.raw_hi = NS ? 1 : 0;
  a.splitk = (int)splitk; a.chunks_per_split = (int)per;
  a.dbg = g_tc_dbg;
  a.nacc = pick_nacc();
  if (p.colsum || p.skinny) {
    NF_CHECK_ARG(TS, NF_E_UNSUPPORTED, "tc_gemm_tn: column sums ride only on the TS weight-gradient kernels");
    NF_CHECK_ARG(!p.skinny || (p.xc && p.nxc >= 1 && p.nxc <= 16), NF_E_BADARG, "tc_gemm_tn: 1..16 skinny columns, got %d", p.nxc);
    a.colsum = p.colsum; a.sColsum = p.sColsum;
    a.xc = p.xc; a.ldxc = p.ldxc; a.nxc = p.skinny ? p.nxc : 0;
    a.xc_al = p.skinny && (p.ldxc & 3) == 0 && (reinterpret_cast<uintptr_t>(p.xc) & 15u) == 0 && (p.sXc2 & 3) == 0 &&
              4 * ((p.nxc + 3) / 4) <= p.ldxc;
    a.xc_vec = a.xc_al && p.nxc <= 8;
    a.skinny = p.skinny; a.ldsk = p.ldsk; a.sSk = p.sSk;
  }
  if (p.inner > 0) {
    NF_CHECK_ARG(p.batch % p.inner == 0, NF_E_BADARG, "tc_gemm_tn: batch %d is not a multiple of the inner batch %d", p.batch, p.inner);
    a.zin = p.inner; a.sC2 = p.sC2; a.sColsum2 = p.sColsum2; a.sSk2 = p.sSk2; a.sXc2 = p.sXc2;
  }
  dim3 grid((unsigned)cdiv(p.M, BM), (unsigned)cdiv(p.N, BN), (unsigned)(splitk * p.batch));
  ProfScope prof(st, PROF_TC_WGRAD, 2.0 * p.M * p.N * (double)p.K * p.batch);
  CUtensorMap mC = mA;
  // (the bulk-tensor-store epilogue addresses C through a 3-d map with one batch stride: two-level batches store directly)
  a.tma_store = (p.inner <= 0 && make_c_map(&mC, p.C, p.N, p.M, p.ldc, p.batch, p.sC)) ? 1 : 0;
  NF_LAUNCH((k_tc_gemm<BN, STAGES, true, BK, EPI_NONE, TS, NSG, NS>), grid, 64 + 128 * NSG, S::TOTAL + kXcBytes, st, mA, mA, mB, mB, mB, mB, mC, a);
  NF_LAUNCHED();
  return 0;
}

}  // namespace

bool tc_encode_available() { return encode_fn() != nullptr; }

int tc_make_map4(CUtensorMap* map, const float* base, long long inner, long long rows, long long ld, long long batch,
                 long long bstride, long long blocks, long long blkstride, int box_inner, int box_rows,
                 CUtensorMapSwizzle swz) {
  EncodeTiledFn fn = encode_fn();
  NF_CHECK_ARG(fn != nullptr, NF_E_UNSUPPORTED, "cuTensorMapEncodeTiled is not available from this driver");
  if (batch < 1 || bstride == 0) { batch = 1; bstride = ld * rows; }
  if (blocks < 1 || blkstride == 0) { blocks = 1; blkstride = bstride * batch; }
  cuuint64_t dims[4] = {(cuuint64_t)inner, (cuuint64_t)rows, (cuuint64_t)batch, (cuuint64_t)blocks};
  cuuint64_t strides[3] = {(cuuint64_t)ld * 4, (cuuint64_t)bstride * 4, (cuuint64_t)blkstride * 4};
  cuuint32_t box[4] = {(cuuint32_t)box_inner, (cuuint32_t)box_rows, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  NF_CHECK_ARG(r == CUDA_SUCCESS, NF_E_BADARG, "cuTensorMapEncodeTiled (4d) failed (%d): inner=%lld rows=%lld ld=%lld "
               "batch=%lld/%lld blocks=%lld/%lld", (int)r, inner, rows, ld, batch, bstride, blocks, blkstride);
  return 0;
}

void g_tc_dbg_set(long long* p) { g_tc_dbg = p; }
long long* tc_dbg_buffer() { return g_tc_dbg; }

bool tc_gemm_supported(const TcGemmP& p) {
  if (p.M <= 0 || p.N <= 0 || p.K0 <= 0 || p.batch <= 0 || p.batch > 65535) return false;
  if (!ok_operand(p.A0, p.lda0, p.sA0) || !ok_operand(p.B0, p.ldb0, p.sB0)) return false;
  if (p.K1 > 0 && (!ok_operand(p.A1, p.lda1, p.sA1) || !ok_operand(p.B1, p.ldb1, p.sB1))) return false;
  if (!p.C || cdiv(p.N, 64) > 65535) return false;
  return encode_fn() != nullptr;
}

bool tc_gemm_epi_supported(const TcGemmP& p) {
  if (!tc_gemm_supported(p) || p.accumulate) return false;
  // One output tile per CTA: the row-wise epilogue is not overlapped with another tile's MMAs, so it only
  // pays while the launch is latency-bound (at most ~2 waves of CTAs); measured on C3/C4/C5 at B >= 8192 the
  // separate row kernel is 15-35 % faster, on C2a (B = 4096) the fused one wins.
  const int force = option(NF_OPT_FUSED_LN);
  if (force == 0) return false;
  if (force != 1 && cdiv(p.M, BM) * cdiv(p.N, 128) * p.batch > 2 * 148) return false;
  const bool n128 = p.N % 128 == 0 && p.N / 128 <= kMaxCluster;
  const bool n64 = p.N % 64 == 0 && p.N / 64 <= kMaxCluster;
  if (!n128 && !n64) return false;
  if ((p.ldc & 3) != 0 || (reinterpret_cast<uintptr_t>(p.C) & 15u) != 0 || (p.sC & 3) != 0) return false;
  if (p.epi.kind == TC_EPI_LN_FWD) return p.epi.mask == nullptr || p.epi.mw * 32 >= p.N;
  if (p.epi.kind == TC_EPI_LN_BWD)
    return p.epi.xh && p.epi.rstd && p.epi.mask && (p.epi.ldxh & 3) == 0 && (p.epi.sXh & 3) == 0 &&
           (reinterpret_cast<uintptr_t>(p.epi.xh) & 15u) == 0 && p.epi.mw * 32 >= p.N;
  return false;
}

bool tc_gemm_tail_supported(const TcGemmP& p, int D) {
  // Off by default: measured on C2a (B = 4096) the fused kernel saves the 6 tail launches of a forward pass
  // but not time (0.227 vs 0.221 ms): with graph replay + PDL a launch costs little, and the leader CTA's
  // serial finish is no faster than 512 CTAs of k_row_tail2.  NF_B200_FUSED_TAIL=1 turns it on.
  if (option(NF_OPT_FUSED_TAIL) != 1 || p.epi.kind != TC_EPI_LN_FWD || p.batch != 2 || D < 2 || D > 9 || !tc_gemm_epi_supported(p)) return false;
  const int bn = (p.N % 128 == 0 && p.N / 128 <= kMaxCluster) ? 128 : 64;
  return 2 * (p.N / bn) <= kMaxCluster;
}

int tc_gemm(const TcGemmP& p, cudaStream_t st) {
  if (p.M <= 0 || p.N <= 0 || p.batch <= 0) return 0;
  if (p.epi.kind != TC_EPI_NONE) {
    NF_CHECK_ARG(tc_gemm_epi_supported(p), NF_E_UNSUPPORTED, "tc_gemm: fused row epilogue needs N = k * 64 (k <= 8), N=%d", p.N);
    return p.epi.kind == TC_EPI_LN_FWD ? launch_epi_pick<EPI_LN_FWD>(p, st) : launch_epi_pick<EPI_LN_BWD>(p, st);
  }
  NF_CHECK_ARG(tc_gemm_supported(p), NF_E_UNSUPPORTED,
               "tc_gemm: operands must be 16-byte aligned with row pitches that are multiples of 4 floats");
  // small problems: prefer more, smaller CTAs (3 pipeline stages) over filling TMEM
  // BN = 256 fills TMEM with one main + one corr accumulator; BN <= 128 leaves room for the three rotating
  // main accumulators that keep long reductions inside fp32 parity, so 256 is only used while one accumulator is enough (K <= 256, pick_nacc).
  if (p.passes == 1 && p.raw_hi && p.bk != 32)   // single raw TF32 pass: no-split form, 12 stages in flight
    return p.N > 64 ? launch<128, 12, 16, false, true>(p, st) : launch<64, 16, 16, false, true>(p, st);
  const int wide_ok = option(NF_OPT_BN256);
  const long long ctas256 = (wide_ok == 1 || (wide_ok == 0 && p.K0 + p.K1 <= 256))
                                ? cdiv(p.M, BM) * cdiv(p.N, 256) * p.batch : 0;
  if (p.bk == 32) {
    if (p.N > 128 && ctas256 >= 120) return launch<256, 2, 32>(p, st);
    if (p.N > 64) return launch<128, 3, 32>(p, st);
    return launch<64, 4, 32>(p, st);
  }
  if (p.N > 128 && ctas256 >= 120) return launch<256, 4, 16>(p, st);
  const bool ts = option(NF_OPT_TS) != 0;
  // under-filled grids of 128-wide tiles (the per-token GEMMs of the autoregressive inverse: M = 4096, N = 256 is 64 CTAs on
  // 148 SMs): 64-wide tiles double the CTAs; the launch is latency-bound, the narrower MMA does not matter (option "bn64": 0 off)
  if (ts && option(NF_OPT_BN64) != 0 && p.N > 64 && p.N % 64 == 0 && !p.accumulate &&
      cdiv(p.M, BM) * cdiv(p.N, 128) * p.batch <= 74)
    return launch<64, 12, 16, true>(p, st);
  if (p.N > 64) {
    if (option(NF_OPT_TS) == 2) return launch<128, 4, 16, true>(p, st);   // experiment: half the stage ring
    if (ts) {
      // weight-tile multicast over CTA pairs: large grids only (several waves of tiles: the L2 fabric is the shared
      // resource there), pre-split weights, no split-K (option "mcast": -1 auto, 0 never, 1 whenever possible)
      const int mc = option(NF_OPT_MCAST);
      const bool pre = option(NF_OPT_PRESPLIT) != 0 && p.B0hi && p.B0lo && (p.K1 <= 0 || (p.B1hi && p.B1lo));
      const long long tiles = cdiv(p.M, BM) * cdiv(p.N, 128) * p.batch;
      const bool nosplit = !(p.accumulate && tiles * 2 <= 148);
      // measured (tools/quick_time.py, B = 8192 / 16384): C4 (N = 512) forward 3.169 -> 3.001 ms, C3 (N = 256 / 320) 0.793 ->
      // 0.777 ms, but C5 (N = 1024, K = 1024) 9.48 -> 9.78 ms -- pairing two CTAs couples their pipelines, and with eight
      // column tiles per row tile the A tiles already share L2 well; auto therefore stops at N = 512
      if (mc != 0 && pre && nosplit && p.M > BM && p.N % 64 == 0 && (mc == 1 || (tiles >= 2 * 148 && p.N <= 512)))
        return launch<128, 8, 16, true, false, 0, true>(p, st);
      return launch<128, 8, 16, true>(p, st);
    }
    return launch<128, 6, 16>(p, st);
  }
  return ts ? launch<64, 12, 16, true>(p, st) : launch<64, 8, 16>(p, st);
}

bool tc_gemm_tn_supported(const TcGemmTnP& p) {
  if (p.M <= 0 || p.N <= 0 || p.K <= 0 || p.K >= (1ll << 31) || p.batch <= 0) return false;
  if (!ok_operand(p.A, p.lda, p.sA) || !ok_operand(p.B, p.ldb, p.sB) || !p.C) return false;
  return encode_fn() != nullptr;
}

bool tc_gemm_tn_rides_supported(const TcGemmTnP& p) {
  // mirrors the dispatch below: the TS kernels (3 passes, 16-deep stages) are the ones with splitter threads on A
  return tc_gemm_tn_supported(p) && p.passes == 3 && p.bk != 32 && option(NF_OPT_BN256) != 1 && option(NF_OPT_TS) != 0;
}

int tc_gemm_tn(const TcGemmTnP& p, cudaStream_t st) {
  if (p.M <= 0 || p.N <= 0 || p.batch <= 0 || p.K <= 0) return 0;
  NF_CHECK_ARG(tc_gemm_tn_supported(p), NF_E_UNSUPPORTED,
               "tc_gemm_tn: operands must be 16-byte aligned with row pitches that are multiples of 4 floats");
  if (p.passes == 1 && p.bk != 32)   // single raw TF32 pass: no-split form
    return p.N > 64 ? launch_tn<128, 12, 16, false, true>(p, st) : launch_tn<64, 16, 16, false, true>(p, st);
  const int wide_ok = option(NF_OPT_BN256);
  if (wide_ok != 1) {
    if (p.bk == 32) return p.N > 64 ? launch_tn<128, 3, 32>(p, st) : launch_tn<64, 4, 32>(p, st);
    if (option(NF_OPT_TS) != 0) return p.N > 64 ? launch_tn<128, 8, 16, true>(p, st) : launch_tn<64, 12, 16, true>(p, st);
    return p.N > 64 ? launch_tn<128, 6, 16>(p, st) : launch_tn<64, 8, 16>(p, st);
  }
  if (p.bk == 32) {
    if (p.N > 128) return launch_tn<256, 2, 32>(p, st);
    if (p.N > 64) return launch_tn<128, 3, 32>(p, st);
    return launch_tn<64, 4, 32>(p, st);
  }
  if (p.N > 128) return launch_tn<256, 4, 16>(p, st);
  if (p.N > 64) return launch_tn<128, 6, 16>(p, st);
  return launch_tn<64, 8, 16>(p, st);
}

}  // namespace nf

// Debug aid: when `buf` is non-null every following tcgen05 GEMM launch writes 8 clock64 timestamps per
// CTA into it (start, prologue done, first tile landed, last tile split, accumulators complete, row in
// registers, statistics exchanged, outputs stored); tools/tc_timeline.py prints the phase durations.
extern "C" __attribute__((visibility("default"))) int nf_debug_tc_timeline(long long* buf) {
  nf::g_tc_dbg_set(buf);
  return 0;
}
