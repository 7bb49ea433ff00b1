// Persistent whole-stack kernel of the coupling flow: see flow_stack.cuh.
//
// Warp roles as in tc_gemm.cu (TS form: A operand through tensor memory, pre-split weights):
//   warp 0       TMA producer            warp 1   MMA issuer (tcgen05.mma kind::tf32 x3 per k8 step), owns TMEM
//   warps 2..5   splitter group 0 (+ 6..9 = group 1 when NSG = 2), then the row-wise epilogue
// Pipeline state (stage ring, TMEM A ring, barrier parities) runs on across the 2 n phases of the stack; every
// phase ends with a cluster barrier executed by ALL warps, after the phase's global writes (bulk tensor stores of
// x_hat, the leader's generic stores of the running buffer) have completed and been fenced for the async proxy.
#include <cuda.h>

#include <algorithm>

#include "flow_stack.cuh"
#include "prof.cuh"
#include "tc_common.cuh"

namespace nf {
namespace {

constexpr int MAXD = 9, MAXD_SQ_BYTES = 9 * 9 * 4;
constexpr int BM = 128, BN = 128, BK = 16, STAGES = 7, NACC = 2, kASlots = 4;
constexpr int A_BYTES = BM * BK * 4, B_BYTES = BN * BK * 4;
constexpr int LO = B_BYTES;                       // the lo twin of a B tile byte sits LO bytes after it
constexpr int STAGE = A_BYTES + 2 * B_BYTES;      // [A | B hi | B lo]
constexpr int BARS = STAGES * STAGE;
constexpr int XCHG = BARS + 512;                  // [rank][group][128] float2 partial row statistics
constexpr int TPART = XCHG + 8 * 128 * 8;         // [cluster rank (<= 8)][group (<= 2)][128] float4 partial last-Linear outputs (leader CTA)
constexpr int SBIAS = TPART + 16 * 128 * 16;      // [BN] bias slice of the current phase
constexpr int TOTAL = SBIAS + BN * 4 + 1024;
constexpr int kW3Off = 150 * 1024;                // tail scratch inside the (idle) operand stages: [4][BN] + [D][D]
static_assert(kW3Off + 4 * BN * 4 + MAXD_SQ_BYTES <= STAGES * STAGE && kW3Off >= 64 * 1024, "tail scratch must sit inside the stage ring, after the store boxes");
constexpr uint32_t A_BASE = (NACC + 1) * BN;      // first TMEM column of the A ring
constexpr uint32_t TMEM_COLS = 512;

struct StackArgs {
  int M, D, dc, dt, H, R, n, inverse, raw_hi, ncl, xh_slots, fast_var;
  float eps;
  const float* b1; long long b1_net, b1_blk;
  const float* b2f; const float* W3f; const float* b3f; const float* Wmat; long long w_net, w_blk;
  const float* cdev;
  const float* cur; float* cur_w; int ld_cur;
  float* rstd1; float* rstd2; uint32_t* mask1; uint32_t* mask2; long long st_net, st_blk; int mw;
  int store_xh2;
  float* out; long long out_blk; float* s_out; long long s_blk; float* logdet;
  float* final_out; int ld_final;
  long long* dbg;   // optional: 12 clock64 timestamps per CTA and phase (nf_debug_tc_timeline)
};

__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
// all bulk stores of this thread have COMPLETED (their global writes are performed), not merely read their source
__device__ __forceinline__ void tma_store_wait_complete() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
// Remote shared-memory store that signals the TARGET CTA's mbarrier with the bytes it wrote (no separate release-arrive:
// the per-thread st.shared::cluster + mbarrier.arrive.release.cluster pair measured ~3 000 cycles per exchange)
__device__ __forceinline__ void st_async_v2(uint32_t raddr, uint32_t rbar, float a, float b) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f32 [%0], {%2, %3}, [%1];"
               ::"r"(raddr), "r"(rbar), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void st_async_v4(uint32_t raddr, uint32_t rbar, float a, float b, float c, float d) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%2, %3, %4, %5}, [%1];"
               ::"r"(raddr), "r"(rbar), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
// lane = row of the box; v = its 32 columns (no proxy fence: the caller fences once for all boxes)
__device__ __forceinline__ void box_put_nofence(uint8_t* box, int lane, const float (&v)[32]) {
#pragma unroll
  for (int j = 0; j < 8; ++j)
    *reinterpret_cast<float4*>(box + lane * 128 + ((j ^ (lane & 7)) << 4)) =
        make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
}

// Maps: mX (running buffer, x_cond columns), mY, W1 x_cond / y column blocks (hi, lo), x_hat1 as load source,
// W2f (hi, lo), x_hat1 / x_hat2 as store targets.
template <int NSG>
__global__ void __launch_bounds__(64 + 128 * NSG, 1)
k_flow_stack(const __grid_constant__ CUtensorMap mX, const __grid_constant__ CUtensorMap mY,
             const __grid_constant__ CUtensorMap mW1a_h, const __grid_constant__ CUtensorMap mW1a_l,
             const __grid_constant__ CUtensorMap mW1b_h, const __grid_constant__ CUtensorMap mW1b_l,
             const __grid_constant__ CUtensorMap mXh1_ld, const __grid_constant__ CUtensorMap mW2_h,
             const __grid_constant__ CUtensorMap mW2_l, const __grid_constant__ CUtensorMap mXh1_st,
             const __grid_constant__ CUtensorMap mXh2_st, const StackArgs a) {
  extern __shared__ uint8_t smem_dyn[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BARS);
  uint64_t* full = bars;
  uint64_t* split = bars + STAGES;
  uint64_t* empty = bars + 2 * STAGES;
  uint64_t* accbar = bars + 3 * STAGES;
  uint64_t* xbar = bars + 3 * STAGES + 1;      // cluster exchange of the LayerNorm partial statistics
  uint64_t* tbar = bars + 3 * STAGES + 2;      // tail partials have arrived at the leader CTA
  uint64_t* aempty = bars + 3 * STAGES + 3;    // the MMAs have read TMEM A slot i (4 slots)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * STAGES + 7);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN, z = blockIdx.z;
  const int ncl = a.ncl;
  // reduction chunks of the two phases: phase 0 = [x_cond (dc) | y (R)], phase 1 = x_hat1 (H)
  const int nk0a = (a.dc + BK - 1) / BK, nk0b = (a.R + BK - 1) / BK, nkp0 = nk0a + nk0b, nkp1 = a.H / BK;

  if (threadIdx.x == 0) {
    prefetch_tmap(&mX); prefetch_tmap(&mY); prefetch_tmap(&mW1a_h); prefetch_tmap(&mW1a_l);
    prefetch_tmap(&mW1b_h); prefetch_tmap(&mW1b_l); prefetch_tmap(&mXh1_ld); prefetch_tmap(&mW2_h);
    prefetch_tmap(&mW2_l); prefetch_tmap(&mXh1_st); prefetch_tmap(&mXh2_st);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&split[s], 4); mbar_init(&empty[s], 1); }
    mbar_init(accbar, 1);
    for (int i = 0; i < kASlots; ++i) mbar_init(&aempty[i], 1);
    // one remote arrive per WARP (after a warp barrier), not per thread: 256-512 remote arrivals per barrier
    // measured 4 700 cycles per exchange (tools/stack_timeline.py)
    mbar_init(xbar, 1);   // one local arrive.expect_tx per phase; the peers' st.async complete the bytes
    mbar_init(tbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();   // every CTA's barriers are initialised before a peer arrives on them
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  NF_PDL_ENTRY();

  const int nphases = 2 * a.n;
  long long* dbg = a.dbg ? a.dbg + 12ll * nphases * (blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * (long long)blockIdx.z)) : nullptr;
#define NF_TS(i) do { if (dbg && threadIdx.x == 64) dbg[12 * p + (i)] = clock64(); } while (0)
  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    int s = 0;
    uint32_t ph = 0;   // parity of the current use of stage s
    for (int p = 0; p < nphases; ++p) {
      const int step = p >> 1, phase = p & 1;
      const int blk = a.inverse ? a.n - 1 - step : step;
      const int xslot = a.xh_slots > 1 ? blk : 0;
      const int nk = phase ? nkp1 : nkp0;
      for (int c = 0; c < nk; ++c) {
        if (lane == 0) mbar_wait(&empty[s], ph ^ 1u);
        __syncwarp();
        uint8_t* st = smem + s * STAGE;
        if (elect_one()) {
          mbar_expect_tx(&full[s], A_BYTES + 2 * B_BYTES);
          if (phase == 0) {
            const bool seg1 = c >= nk0a;
            const int kk = (seg1 ? c - nk0a : c) * BK;
            tma_load_4d(st, seg1 ? &mY : &mX, &full[s], kk, m0, 0, 0);
            tma_load_4d(st + A_BYTES, seg1 ? &mW1b_h : &mW1a_h, &full[s], kk, n0, z, blk);
            tma_load_4d(st + A_BYTES + LO, seg1 ? &mW1b_l : &mW1a_l, &full[s], kk, n0, z, blk);
          } else {
            const int kk = c * BK;
            tma_load_4d(st, &mXh1_ld, &full[s], kk, m0, z, xslot);
            tma_load_4d(st + A_BYTES, &mW2_h, &full[s], kk, n0, z, blk);
            tma_load_4d(st + A_BYTES + LO, &mW2_l, &full[s], kk, n0, z, blk);
          }
        }
        __syncwarp();
        if (++s == STAGES) { s = 0; ph ^= 1u; }
      }
      cluster_sync_all();   // end of phase (see the epilogue warps)
      fence_proxy_async_all();
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(BN >> 3) << 17) | (uint32_t(BM >> 4) << 24);
    const uint32_t tmem_corr = tmem_base + NACC * BN;
    const uint64_t d0 = make_sdesc<BK>(0u);
    const uint32_t dhi = static_cast<uint32_t>(d0 >> 32), dlo = static_cast<uint32_t>(d0);
    const bool leader = elect_one();
    const uint32_t lo0 = dlo + ((smem_u32(smem) & 0x3FFFFu) >> 4);
    int s = 0, aslot = 0;
    uint32_t ph = 0;
    for (int p = 0; p < nphases; ++p) {
      const int phase = p & 1;
      const int nk = phase ? nkp1 : nkp0;
      uint32_t tmem_main = tmem_base, acc_c = 0, fresh = NACC;
      auto kleft_of = [&](int c) { return phase ? a.H - c * BK : (c < nk0a ? a.dc - c * BK : a.R - (c - nk0a) * BK); };
      auto issue_stage = [&](int sidx, int slot, int nks) {
        const uint32_t st_lo = lo0 + (uint32_t)sidx * (STAGE >> 4);
        const uint32_t ta = tmem_base + A_BASE + (uint32_t)slot * 2 * BK;
        uint32_t lb_hi = st_lo + (A_BYTES >> 4), lb_lo = st_lo + ((A_BYTES + LO) >> 4);
#pragma unroll
        for (int ks = 0; ks < BK / 8; ++ks) {
          if (ks < nks) {
            const uint64_t b_hi = (static_cast<uint64_t>(dhi) << 32) | lb_hi, b_lo = (static_cast<uint64_t>(dhi) << 32) | lb_lo;
            const uint32_t acc_m = fresh == 0 ? 1u : 0u;
            const uint32_t ta_hi = ta + (uint32_t)ks * 8, ta_lo = ta_hi + BK;
            if (leader) {
              umma_tf32_ts(tmem_corr, ta_lo, b_hi, idesc, acc_c);
              umma_tf32_ts(tmem_corr, ta_hi, b_lo, idesc, 1u);
              umma_tf32_ts(tmem_main, ta_hi, b_hi, idesc, acc_m);
            }
            acc_c = 1;
            if (fresh) --fresh;
            tmem_main += BN;
            if (tmem_main == tmem_base + NACC * BN) tmem_main = tmem_base;
            lb_hi += 32u >> 4; lb_lo += 32u >> 4;
          }
        }
      };
      // two stages per loop trip: both barriers are awaited first, then up to 12 MMAs go out back to back
      for (int c = 0; c < nk; c += 2) {
        const bool two = c + 1 < nk;
        const int kl0 = kleft_of(c), kl1 = two ? kleft_of(c + 1) : 0;
        const int nks0 = kl0 >= BK ? BK / 8 : (kl0 + 7) >> 3, nks1 = kl1 >= BK ? BK / 8 : (kl1 + 7) >> 3;
        const bool wrap = s + 1 == STAGES;
        const int s1 = wrap ? 0 : s + 1;
        const uint32_t ph1 = wrap ? ph ^ 1u : ph;
        const int aslot1 = aslot + 1 == kASlots ? 0 : aslot + 1;
        if (lane == 0) {
          mbar_wait(&split[s], ph);
          if (two) mbar_wait(&split[s1], ph1);
        }
        __syncwarp();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        issue_stage(s, aslot, nks0);
        if (leader) { umma_commit(&empty[s]); umma_commit(&aempty[aslot]); }
        if (two) {
          issue_stage(s1, aslot1, nks1);
          if (leader) { umma_commit(&empty[s1]); umma_commit(&aempty[aslot1]); }
        }
        __syncwarp();
        const int adv = two ? 2 : 1;
        for (int k = 0; k < adv; ++k) {
          if (++s == STAGES) { s = 0; ph ^= 1u; }
          if (++aslot == kASlots) aslot = 0;
        }
      }
      if (leader) umma_commit(accbar);
      __syncwarp();
      cluster_sync_all();   // end of phase: the epilogue has drained the accumulators
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
  } else {
    // ------------------------------------------------------------------ splitters, then the row-wise epilogue
    const int t = (threadIdx.x - 64) & 127;
    const int grp = (threadIdx.x - 64) >> 7;
    const int q = warp & 3;
    const int rit = q * 32 + lane;                // row in the tile = TMEM lane of this thread
    const int row = m0 + rit;
    const bool rowok = row < a.M;
    constexpr int CW = BN / NSG;
    const int gq = NSG > 1 ? grp : 0;
    const int cg0 = gq * CW;
    const int dt = a.dt, D = a.D, dc = a.dc;
    const uint32_t rank_base = (uint32_t)z * (uint32_t)ncl;
    const uint32_t my_rank = cluster_ctarank() - rank_base;
    const uint32_t crank = cluster_ctarank();
    float2* xchg = reinterpret_cast<float2*>(smem + XCHG);
    float4* tpart = reinterpret_cast<float4*>(smem + TPART);
    uint8_t* boxes = smem + (q * (BN / 32) + (cg0 >> 5)) * 4096;   // 16 boxes of 4 KB at the start of the stage ring
    float* sW3 = reinterpret_cast<float*>(smem + kW3Off);          // [4][BN]
    float* sWm = sW3 + 4 * BN;                                     // [D][D] (leader)
    float* sBias = reinterpret_cast<float*>(smem + SBIAS);         // [BN]
    int gi = 0;     // global stage counter (all phases)
    for (int p = 0; p < nphases; ++p) {
      const int step = p >> 1, phase = p & 1;
      const int blk = a.inverse ? a.n - 1 - step : step;
      const int xslot = a.xh_slots > 1 ? blk : 0;
      const int nk = phase ? nkp1 : nkp0;
      NF_TS(0);
      if (threadIdx.x == 64) {
        mbar_expect_tx(xbar, (uint32_t)ncl * NSG * 128u * 8u);
        if (phase == 1 && crank == 0) mbar_expect_tx(tbar, 2u * (uint32_t)ncl * NSG * 128u * 16u);
      }
      {   // this phase's bias slice -> shared memory (the global latency hides behind the main loop)
        const float* bsrc = phase ? a.b2f + (long long)z * a.w_net + (long long)blk * a.w_blk + n0
                                  : a.b1 + (long long)z * a.b1_net + (long long)blk * a.b1_blk + n0;
        const int te = threadIdx.x - 64;
        if (te < BN) sBias[te] = __ldg(bsrc + te);
      }
      for (int c = 0; c < nk; ++c, ++gi) {
        if (NSG > 1 && (gi % NSG) != grp) continue;
        const int s = gi % STAGES;
        mbar_wait(&full[s], (uint32_t)(gi / STAGES) & 1u);
        if (c < NSG) NF_TS(1);
        const float4* raw = reinterpret_cast<const float4*>(smem + s * STAGE);
        const int sw = (rit >> 1) & 3;             // SWIZZLE_64B: 16-byte chunk c sits at c ^ ((row >> 1) & 3)
        float4 av[4];
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) av[cc] = raw[rit * 4 + (cc ^ sw)];
        uint32_t hi16[16], lo16[16];
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          const float xs[4] = {av[cc].x, av[cc].y, av[cc].z, av[cc].w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float h = a.raw_hi ? __uint_as_float(__float_as_uint(xs[e]) & 0xffffe000u) : rna_tf32(xs[e]);
            hi16[4 * cc + e] = __float_as_uint(h);
            lo16[4 * cc + e] = __float_as_uint(xs[e] - h);
          }
        }
        const int slot = gi % kASlots;
        if (gi >= kASlots) {
          if (lane == 0) mbar_wait(&aempty[slot], ((uint32_t)(gi / kASlots) & 1u) ^ 1u);
          __syncwarp();
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        const uint32_t ta = tmem_base + A_BASE + (uint32_t)slot * 2 * BK + (static_cast<uint32_t>(q * 32) << 16);
        tmem_st16(ta, hi16);
        tmem_st16(ta + BK, lo16);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&split[s]);
      }
      // ---------------------------------------------------------------- epilogue of this phase
      NF_TS(2);
      mbar_wait(accbar, (uint32_t)p & 1u);
      NF_TS(3);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      // how many main accumulators were written (k8 steps of this phase)
      int total_k8;
      if (phase) total_k8 = a.H / 8;
      else total_k8 = (a.dc + 7) / 8 + (a.R / BK) * 2 + (((a.R % BK) + 7) >> 3);
      const int nacc_used = total_k8 < NACC ? total_k8 : NACC;
      float v[CW];
#pragma unroll
      for (int c0 = 0; c0 < CW; c0 += 32) {
        float r[32];
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(cg0 + c0);
        tmem_ld_sum<BN, NACC>(taddr, nacc_used, true, r);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[c0 + j] = r[j];
      }
      NF_TS(4);
      if (NSG == 1) asm volatile("bar.sync 1, 128;" ::: "memory");   // sBias is complete
      else asm volatile("bar.sync 1, 256;" ::: "memory");
      const float* bias = sBias + cg0;
      uint32_t mbits[CW / 32];
      float sum = 0.f;
#pragma unroll
      for (int c = 0; c < CW; ++c) {
        const float pre = v[c] + bias[c];
        if ((c & 31) == 0) mbits[c >> 5] = 0u;
        if (pre > 0.f) mbits[c >> 5] |= 1u << (c & 31);
        v[c] = pre > 0.f ? pre : kLeakySlope * pre;
        sum += v[c];
      }
      const float p0 = sum * (1.0f / CW);
      float p1 = 0.f;
      const float sh = a.fast_var ? 0.f : p0;
#pragma unroll
      for (int c = 0; c < CW; ++c) { const float d = v[c] - sh; p1 = fmaf(d, d, p1); }
      {
        const uint32_t slot = smem_u32(&xchg[(my_rank * NSG + gq) * 128 + rit]);
        const uint32_t xb = smem_u32(xbar);
        for (int pr = 0; pr < ncl; ++pr)
          st_async_v2(mapa_u32(slot, rank_base + (uint32_t)pr), mapa_u32(xb, rank_base + (uint32_t)pr), p0, p1);
      }
      mbar_wait_cluster(xbar, (uint32_t)p & 1u);
      NF_TS(5);
      float mean = 0.f;
      for (int pr = 0; pr < ncl * NSG; ++pr) mean += xchg[pr * 128 + rit].x;
      mean /= (float)(ncl * NSG);
      float m2 = 0.f;
      for (int pr = 0; pr < ncl * NSG; ++pr) {
        const float2 sx = xchg[pr * 128 + rit];
        const float d = sx.x - mean;
        m2 += a.fast_var ? sx.y : sx.y + (float)CW * d * d;
      }
      float var = m2 / (float)(ncl * BN);
      if (a.fast_var) var = fmaxf(var - mean * mean, 0.f);
      const float rstd = rsqrtf(var + a.eps);
      const bool tail = phase == 1;
      const bool store_c = phase == 0 || a.store_xh2;
      float part[4] = {0.f, 0.f, 0.f, 0.f};
      const float* Wm = nullptr;
      if (tail) {
        // this CTA's slice of the folded last Linear, W3f[net][i][n0 .. n0 + BN), and (leader) the PLU matrix
        const int te = threadIdx.x - 64;
        const float* W3 = a.W3f + (long long)z * a.w_net + (long long)blk * a.w_blk;
        for (int idx = te; idx < 4 * BN; idx += 128 * NSG) {
          const int i = idx / BN, c = idx - i * BN;
          sW3[idx] = i < dt ? __ldg(W3 + (long long)i * a.H + n0 + c) : 0.f;
        }
        if (a.inverse) Wm = a.Wmat + (long long)blk * a.w_blk;
        else if (blk + 1 < a.n) Wm = a.Wmat + (long long)(blk + 1) * a.w_blk;
        if (crank == 0 && Wm)
          for (int idx = te; idx < D * D; idx += 128 * NSG) sWm[idx] = __ldg(Wm + idx);
        if (NSG == 1) asm volatile("bar.sync 1, 128;" ::: "memory");
        else asm volatile("bar.sync 1, 256;" ::: "memory");
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
              const float4 w4 = *reinterpret_cast<const float4*>(sW3 + i * BN + cg0 + c0 + j);
              part[i] = fmaf(o[j], w4.x, fmaf(o[j + 1], w4.y, fmaf(o[j + 2], w4.z, fmaf(o[j + 3], w4.w, part[i]))));
            }
          }
        }
        if (store_c) box_put_nofence(boxes + (c0 >> 5) * 4096, lane, o);
      }
      if (store_c) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (elect_one()) {
#pragma unroll
          for (int c0 = 0; c0 < CW; c0 += 32)
            tma_store_4d(phase ? &mXh2_st : &mXh1_st, boxes + (c0 >> 5) * 4096, n0 + cg0 + c0, m0 + q * 32, z, xslot);
          tma_store_commit();
        }
        __syncwarp();
      }
      NF_TS(6);
      if (rowok) {
        uint32_t* mask = phase ? a.mask2 : a.mask1;
        float* rs = phase ? a.rstd2 : a.rstd1;
        const long long so = (long long)z * a.st_net + (long long)blk * a.st_blk;
        if (mask) {
          uint32_t* mk = mask + so + (long long)row * a.mw + ((n0 + cg0) >> 5);
#pragma unroll
          for (int w = 0; w < CW / 32; ++w) mk[w] = mbits[w];
        }
        if (rs && my_rank == 0 && cg0 == 0) rs[so + row] = rstd;
      }
      if (tail) {
        // partial last-Linear outputs of this thread's columns -> the cluster's leader CTA (rank 0)
        const uint32_t dst = mapa_u32(smem_u32(&tpart[(crank * NSG + gq) * 128 + rit]), 0u);
        st_async_v4(dst, mapa_u32(smem_u32(tbar), 0u), part[0], part[1], part[2], part[3]);
        if (crank == 0 && gq == 0) {
          // this row of the running buffer: loaded before the wait so that the global latency overlaps it
          float xin[MAXD];
#pragma unroll
          for (int j = 0; j < MAXD; ++j) xin[j] = (rowok && j < D) ? a.cur[(long long)row * a.ld_cur + j] : 0.f;
          const float ld_old = (rowok && a.logdet) ? a.logdet[row] : 0.f;
          const float* b3 = a.b3f + (long long)blk * a.w_blk;
          float b3t[4], b3s[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) { b3t[i] = i < dt ? __ldg(b3 + i) : 0.f; b3s[i] = i < dt ? __ldg(b3 + a.w_net + i) : 0.f; }
          const float cst = __ldg(a.cdev + blk);
          mbar_wait_cluster(tbar, (uint32_t)step & 1u);
          NF_TS(11);
          // t = net 0 (ranks 0 .. ncl-1), s = net 1 (ranks ncl .. 2 ncl - 1)     (torch_fcnf.py:99-104 / :110-115)
          float tv[4] = {0.f, 0.f, 0.f, 0.f}, sv[4] = {0.f, 0.f, 0.f, 0.f};
          for (int pr = 0; pr < ncl * NSG; ++pr) {
            const float4 pt = tpart[pr * 128 + rit], ps = tpart[(ncl * NSG + pr) * 128 + rit];
            tv[0] += pt.x; tv[1] += pt.y; tv[2] += pt.z; tv[3] += pt.w;
            sv[0] += ps.x; sv[1] += ps.y; sv[2] += ps.z; sv[3] += ps.w;
          }
          if (rowok) {
            float zv[MAXD];
            float ssum = 0.f;
#pragma unroll
            for (int j = 0; j < MAXD; ++j) zv[j] = xin[j];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              if (i < dt) {
                const float tt = tv[i] + b3t[i], ss = sv[i] + b3s[i];
                sv[i] = ss;
                ssum += ss;
#pragma unroll
                for (int j = 0; j < MAXD; ++j)
                  if (j == dc + i) zv[j] = a.inverse ? fmaf(xin[j], expf(ss), tt) : (xin[j] - tt) * expf(-ss);
              }
            }
            if (!a.inverse) {
              float* outp = a.out_blk ? a.out + (long long)blk * a.out_blk : (blk == a.n - 1 ? a.out : nullptr);
              if (outp) {
#pragma unroll
                for (int j = 0; j < MAXD; ++j)
                  if (j < D) outp[(long long)row * D + j] = zv[j];
              }
              if (a.s_out) {
                float* sp = a.s_out + (long long)blk * a.s_blk;
#pragma unroll
                for (int i = 0; i < 4; ++i)
                  if (i < dt) sp[(long long)row * dt + i] = sv[i];
              }
            }
            if (a.logdet) {
              a.logdet[row] = ld_old + (a.inverse ? (ssum - cst) : (cst - ssum));
            }
            if (Wm) {
              const bool last = a.inverse && step == a.n - 1 && a.final_out;
              float* o2 = last ? a.final_out : a.cur_w;
              const int ld2 = last ? a.ld_final : a.ld_cur;
#pragma unroll
              for (int j = 0; j < MAXD; ++j) {
                if (j < D) {
                  float acc = 0.f;
#pragma unroll
                  for (int k = 0; k < MAXD; ++k)
                    if (k < D) acc = fmaf(zv[k], sWm[k * D + j], acc);
                  o2[(long long)row * ld2 + j] = acc;
                }
              }
            }
          }
        }
      }
      // end of phase: this thread's bulk stores are complete, its generic global stores are ordered before the
      // peers' next TMA loads, then the whole cluster meets (every warp of every CTA executes this barrier)
      NF_TS(7);
      tma_store_wait_complete();
      NF_TS(8);
      // generic-proxy global stores (the leader's running-buffer rows) -> the peers' next TMA loads (async proxy);
      // the cluster barrier below releases / acquires at cluster scope
      if (tail && crank == 0) fence_proxy_async_all();
      NF_TS(9);
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      cluster_sync_all();
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      NF_TS(10);
    }
  }
#undef NF_TS
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

template <int NSG>
int launch_stack(const StackP& p, cudaStream_t st) {
  static PerDeviceOnce once;
  if (once.first())
    NF_CUDA(cudaFuncSetAttribute(k_flow_stack<NSG>, cudaFuncAttributeMaxDynamicSharedMemorySize, TOTAL));
  const int dc = (p.D + 1) / 2, dt = p.D / 2, H = p.H;
  const long long B = p.B;
  const int slots = p.xh_blk ? p.n : 1;
  const CUtensorMapSwizzle k64 = CU_TENSOR_MAP_SWIZZLE_64B, s128 = CU_TENSOR_MAP_SWIZZLE_128B;
  CUtensorMap mX, mY, mW1a_h, mW1a_l, mW1b_h, mW1b_l, mXh1_ld, mW2_h, mW2_l, mXh1_st, mXh2_st;
  NF_TRY(tc_make_map4(&mX, p.cur, dc, B, p.ld_cur, 1, 0, 1, 0, BK, BM, k64));
  NF_TRY(tc_make_map4(&mY, p.y, p.R, B, p.R, 1, 0, 1, 0, BK, BM, k64));
  NF_TRY(tc_make_map4(&mW1a_h, p.W1p + p.hi_off, dc, H, p.K1p, 2, p.w_net, p.n, p.w_blk, BK, BN, k64));
  NF_TRY(tc_make_map4(&mW1a_l, p.W1p + p.lo_off, dc, H, p.K1p, 2, p.w_net, p.n, p.w_blk, BK, BN, k64));
  NF_TRY(tc_make_map4(&mW1b_h, p.W1p + p.hi_off + p.dcp, p.R, H, p.K1p, 2, p.w_net, p.n, p.w_blk, BK, BN, k64));
  NF_TRY(tc_make_map4(&mW1b_l, p.W1p + p.lo_off + p.dcp, p.R, H, p.K1p, 2, p.w_net, p.n, p.w_blk, BK, BN, k64));
  NF_TRY(tc_make_map4(&mXh1_ld, p.xh1, H, B, H, 2, p.xh_net, slots, p.xh_blk, BK, BM, k64));
  NF_TRY(tc_make_map4(&mW2_h, p.W2f + p.hi_off, H, H, H, 2, p.w_net, p.n, p.w_blk, BK, BN, k64));
  NF_TRY(tc_make_map4(&mW2_l, p.W2f + p.lo_off, H, H, H, 2, p.w_net, p.n, p.w_blk, BK, BN, k64));
  NF_TRY(tc_make_map4(&mXh1_st, p.xh1, H, B, H, 2, p.xh_net, slots, p.xh_blk, 32, 32, s128));
  NF_TRY(tc_make_map4(&mXh2_st, p.xh2, H, B, H, 2, p.xh_net, slots, p.xh_blk, 32, 32, s128));
  StackArgs a{};
  a.M = (int)B; a.D = p.D; a.dc = dc; a.dt = dt; a.H = H; a.R = p.R; a.n = p.n; a.inverse = p.inverse;
  a.raw_hi = p.raw_hi; a.ncl = H / BN; a.xh_slots = slots; a.eps = p.eps; a.fast_var = p.fast_var;
  a.b1 = p.b1; a.b1_net = p.b1_net; a.b1_blk = p.b1_blk;
  a.b2f = p.b2f; a.W3f = p.W3f; a.b3f = p.b3f; a.Wmat = p.Wmat; a.w_net = p.w_net; a.w_blk = p.w_blk;
  a.cdev = p.cdev; a.cur = p.cur; a.cur_w = p.cur_w; a.ld_cur = p.ld_cur;
  a.rstd1 = p.rstd1; a.rstd2 = p.rstd2; a.mask1 = p.mask1; a.mask2 = p.mask2;
  a.st_net = p.xh_net; a.st_blk = p.xh_blk; a.mw = p.mw; a.store_xh2 = p.store_xh2;
  a.out = p.out; a.out_blk = p.out_blk; a.s_out = p.s_out; a.s_blk = p.s_blk; a.logdet = p.logdet;
  a.final_out = p.final_out; a.ld_final = p.ld_final;
  a.dbg = tc_dbg_buffer();
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)cdiv(B, BM), (unsigned)a.ncl, 2u);
  cfg.blockDim = dim3(64 + 128 * NSG);
  cfg.dynamicSmemBytes = TOTAL;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1; attr[0].val.clusterDim.y = (unsigned)a.ncl; attr[0].val.clusterDim.z = 2u;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 2 : 1;
  const double flops = 2.0 * B * 2.0 * ((double)(dc + p.R) * H + (double)H * H) * p.n;
  ProfScope prof(st, PROF_TC_GEMM, flops);
  NF_CUDA(cudaLaunchKernelEx(&cfg, k_flow_stack<NSG>, mX, mY, mW1a_h, mW1a_l, mW1b_h, mW1b_l, mXh1_ld, mW2_h, mW2_l,
                             mXh1_st, mXh2_st, a));
  NF_LAUNCHED();
  return 0;
}

}  // namespace

bool flow_stack_supported(const StackP& p) {
  if (p.B <= 0 || p.n < 1 || p.D < 2 || p.D > MAXD || p.D / 2 > 4) return false;
  if (p.H % BN != 0 || 2 * (p.H / BN) > kMaxCluster) return false;
  if (p.R < 4 || (p.R & 3) != 0 || (p.ld_cur & 3) != 0 || (p.K1p & 3) != 0 || (p.dcp & 3) != 0) return false;
  if ((p.w_net & 3) || (p.w_blk & 3) || (p.xh_net & 3) || (p.xh_blk & 3) || (p.hi_off & 3) || (p.lo_off & 3)) return false;
  if (!p.hi_off || !p.lo_off || !p.cur || !p.y || !p.xh1 || !p.xh2) return false;
  if (p.mask1 && p.mw * 32 < p.H) return false;
  return tc_encode_available();
}

int flow_stack(const StackP& p, cudaStream_t st) {
  NF_CHECK_ARG(flow_stack_supported(p), NF_E_UNSUPPORTED, "flow_stack: unsupported shape (D=%d H=%d R=%d)", p.D, p.H, p.R);
  return option(NF_OPT_STACK_NSG) == 1 ? launch_stack<1>(p, st) : launch_stack<2>(p, st);
}

}  // namespace nf
