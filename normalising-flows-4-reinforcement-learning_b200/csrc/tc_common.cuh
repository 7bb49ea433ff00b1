// Device-side building blocks shared by the tcgen05 kernels (tc_gemm.cu, flow_stack.cu): mbarrier / TMA / tcgen05
// wrappers, descriptor builders, the warp transposition helpers.  sm_100a only.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace nf {

// host: 4-D fp32 tensor map (inner, rows, batch, block); strides in floats, box = box_inner x box_rows x 1 x 1
int tc_make_map4(CUtensorMap* map, const float* base, long long inner, long long rows, long long ld, long long batch,
                 long long bstride, long long blocks, long long blkstride, int box_inner, int box_rows,
                 CUtensorMapSwizzle swz);
bool tc_encode_available();
long long* tc_dbg_buffer();   // nf_debug_tc_timeline: optional clock64 timeline buffer (null = off)

namespace {

constexpr int kMaxCluster = 8;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P1;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must end in a trap (reported as a launch failure), never a hang.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  for (uint32_t spins = 0; !mbar_try(bar, parity); ++spins) {
    if (spins > (1u << 24)) {
      printf("nf_b200 tc_gemm: mbarrier wait timed out (block %d,%d,%d thread %d)\n", blockIdx.x, blockIdx.y,
             blockIdx.z, threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// the same load delivered to the CTAs of `mask` in the cluster (same shared-memory offset and mbarrier offset in each)
__device__ __forceinline__ void tma_load_3d_mc(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                                               uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4, %5}], [%2], %6;"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "h"(mask)
      : "memory");
}
// Output tile -> global memory through the TMA: the epilogue thread (one row of the TMEM layout) writes its
// 32 columns into a 32 x 32 fp32 box in shared memory (128-byte rows, SWIZZLE_128B so the row-per-lane
// 128-bit stores are conflict-free), one elected lane issues a bulk tensor store or reduce-add.  Rows and
// columns outside the matrix are clipped by the tensor map.  Replaces row-per-lane global stores that
// were LSU-bound (1 000 cycles per 32-column chunk, tools/tc_timeline.py).
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_reduce_add_3d(const CUtensorMap* map, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// the shared-memory source has been read (it may be released); the global writes complete with the grid
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// lane = row of the box; v = its 32 columns
__device__ __forceinline__ void box_put(uint8_t* box, int lane, const float (&v)[32]) {
#pragma unroll
  for (int j = 0; j < 8; ++j)
    *reinterpret_cast<float4*>(box + lane * 128 + ((j ^ (lane & 7)) << 4)) =
        make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
// K-major operand descriptor: rows of BK*4 bytes (128 B -> SWIZZLE_128B, 64 B -> SWIZZLE_64B), 8-row
// groups SBO bytes apart, descriptor version 1 (Blackwell); layout codes of cute::UMMA::LayoutType.
template <int BK>
__device__ __forceinline__ uint64_t make_sdesc(uint32_t saddr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;                       // LBO (unused for swizzled K-major), 16 B units
  d |= static_cast<uint64_t>((8 * BK * 4) >> 4) << 32;       // SBO: 8 rows
  d |= static_cast<uint64_t>(1) << 46;                       // version
  d |= static_cast<uint64_t>(BK == 32 ? 2 : 4) << 61;        // SWIZZLE_128B : SWIZZLE_64B
  return d;
}
// MN-major fp32/tf32 operands must use the 128B-swizzle-with-32B-atom layout (cute
// Layout_MN_SW128_32B_Atom, UMMA layout type 1, TMA CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): canonical
// form ((T,8,m),(4,k)):((1,T,LBO),(8T,SBO)) in elements, T = 4 fp32 -- 32 fp32 contiguous along M/N
// per 128 B row, 4 k-rows per 512 B swizzle atom.  Every 32 x BK TMA box is BK*128 B, so
// LBO (distance between 32-element M/N groups) = BK*128 B and SBO (distance between 4-row k groups)
// = 512 B.
template <int BK>
__device__ __forceinline__ uint64_t make_sdesc_mn(uint32_t saddr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((BK * 128) >> 4) << 16;         // LBO
  d |= static_cast<uint64_t>(512 >> 4) << 32;                // SBO
  d |= static_cast<uint64_t>(1) << 46;                       // version
  d |= static_cast<uint64_t>(1) << 61;                       // SWIZZLE_128B_BASE32B
  return d;
}
__device__ __forceinline__ void red_add_v4(float* dst, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// A operand from tensor memory (M x K, one 32-bit element per column, row = lane)
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// the arrive lands on the barrier at the same offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// Sum of the NACC main accumulators (only `nacc_used` were written) and the corr accumulator of one 32-column
// chunk: all TMEM loads are issued before the single wait (one TMEM round trip per chunk instead of four).
template <int BN, int NACC>
__device__ __forceinline__ void tmem_ld_sum(uint32_t taddr, int nacc_used, bool with_corr, float (&out)[32]) {
  uint32_t r[NACC + 1][32];
  tmem_ld32(taddr, r[0]);
#pragma unroll
  for (int ac = 1; ac <= NACC; ++ac)
    if (ac < NACC ? ac < nacc_used : with_corr) tmem_ld32(taddr + ac * BN, r[ac]);
  tmem_ld_wait();
#pragma unroll
  for (int j = 0; j < 32; ++j) out[j] = __uint_as_float(r[0][j]);
#pragma unroll
  for (int ac = 1; ac <= NACC; ++ac)
    if (ac < NACC ? ac < nacc_used : with_corr) {
#pragma unroll
      for (int j = 0; j < 32; ++j) out[j] += __uint_as_float(r[ac][j]);
    }
}
// One elected lane of a fully converged warp.  The producer and MMA warps run their loops with ALL lanes
// (so every operand stays warp-uniform and lives in uniform registers) and only the instruction issue is
// elected; under `if (lane == 0)` the compiler wrapped every UTCHMMA / UTMALDG in an ELECT + R2UR.BROADCAST
// + BRA.U.ANY waterfall loop, which made the single issuing thread the bottleneck of the tensor pipe.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_v2(uint32_t raddr, float a, float b) {
  asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(raddr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t raddr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  for (uint32_t spins = 0;; ++spins) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, P1;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (ok) return;
    if (spins > (1u << 24)) {
      printf("nf_b200 tc_gemm: cluster exchange timed out (block %d,%d,%d thread %d)\n", blockIdx.x, blockIdx.y,
             blockIdx.z, threadIdx.x);
      __trap();
    }
  }
}
// Warp-level 32 x 32 transposition through shared memory (pitch 36 floats: conflict-free for the 128-bit
// row-per-lane accesses and for the 128-byte row segments).  In the TMEM layout a lane owns a ROW; global
// memory wants a warp on one row segment.  A row-per-lane float4 store touches 32 different 128-byte
// lines (measured: the epilogue was LSU-bound at 4800 cycles per 128 x 128 tile); after the transposition
// every store instruction writes four full lines.
constexpr int kXPitch = 36;
constexpr int kXWarpFloats = 32 * kXPitch;
__device__ __forceinline__ void xpose_put(float* sbuf, int lane, const float (&v)[32]) {
#pragma unroll
  for (int j = 0; j < 8; ++j)
    *reinterpret_cast<float4*>(sbuf + lane * kXPitch + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  __syncwarp();
}
// i-th of 8 pieces: rows (lane >> 3) + 4 i, columns 4 (lane & 7) .. +3
__device__ __forceinline__ float4 xpose_get(const float* sbuf, int lane, int i) {
  return *reinterpret_cast<const float4*>(sbuf + ((lane >> 3) + 4 * i) * kXPitch + 4 * (lane & 7));
}
// the reverse direction: coalesced pieces in, row-per-lane out
__device__ __forceinline__ void xpose_put_piece(float* sbuf, int lane, int i, float4 v) {
  *reinterpret_cast<float4*>(sbuf + ((lane >> 3) + 4 * i) * kXPitch + 4 * (lane & 7)) = v;
}
__device__ __forceinline__ void xpose_get_row(const float* sbuf, int lane, float (&v)[32]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float4 t = *reinterpret_cast<const float4*>(sbuf + lane * kXPitch + 4 * j);
    v[4 * j] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
  }
}
__device__ __forceinline__ float rna_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}


}  // namespace
}  // namespace nf
