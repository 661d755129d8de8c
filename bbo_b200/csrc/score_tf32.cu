// TF32 scoring mode: the n^2 posterior-variance panel on 5th-gen tensor cores.
//
//   |L^-1 k*_c|^2 = sum_i ( sum_{j<=i} Ks[c,j] * Linv[i,j] )^2          (gp.py:80-81)
//
// Operands of the panel: TF32's 10-bit mantissa in binary16 containers -- Ks holds correlations K*/s2 in (0, 1],
// Linv a power-of-two multiple of L^-1 (k_make_f16) -- multiplied by tcgen05 kind::f16 into float32 accumulators;
// the epilogue's sums of squares are scaled back by s2^2 2^-2e (panel_factor).  Same rounding as kind::tf32 at
// twice the rate and half the bytes.
//
// Kernels (one default per shape; the A/B generations of round 1 were removed in round 2):
//   K* producers (correlations of a chunk rounded to binary16, fused with mu = c + K* alpha from the float32 values)
//     k_kstar_umma       d <= 126     tcgen05 kind::tf32, 3xTF32 contraction by pairing the k-steps of [hi | lo]
//                                     operand rows; candidate tile resident, observation boxes in a TMA ring, TMEM
//                                     accumulator ring, TMA tensor stores
//     k_kstar_mma2       (mma.sync 3xTF32 with the two augmentation columns; A/B reference for the one above)
//     k_kstar_f32_fast   126 < d <= 192 CUDA cores, expanded form;   k_kstar_f32  any d, exact-difference form
//   n^2 panel (V = L^-1 k* is never written: the epilogue squares and row-sums the accumulators)
//     k_vnorm_tf32_2sm<true>   persistent CTA pairs walking a work list (k_panel_schedule): default while the
//                              binary16 L^-1 fits in L2 (n <= ~6300)
//     k_vnorm_tf32_2sm<false>  one CTA pair per candidate tile pair, lock-step over the row pairs (n = 8192)
//   (chunks are padded to 256 candidates, so every chunk runs on CTA pairs; the single-CTA forms are gone)
//   warp roles of the panel kernels: warp 0 TMA producer (cp.async.bulk.tensor.2d, 64-byte swizzle, mbarrier
//   ring), warp 1 one thread issuing tcgen05.mma kind::f16 (M = 128 per CTA x N = 256 x K = 16, accumulators in
//   TMEM, two row blocks per K* stage), warps 2-5 epilogue (tcgen05.ld.32x32b.x32 -> square + row-sum).
//   Triangular: row block nb only contracts k < (nb+1)*256.
// SASS evidence: UTCHMMA(.2CTA) / UTMALDG / UTMASTG / LDTM / SYNCS (see profiles/).
#include <cuda.h>
#include <cuda_fp16.h>
#include <algorithm>
#include <mutex>
#include <vector>

#include <cstdio>
#include <cstdlib>
#include <functional>

#include "score.cuh"

namespace bbo {

namespace tf {
constexpr int BM = 128;       // candidates per CTA tile (UMMA M)
constexpr int BN = 256;       // rows of L^-1 per accumulator (UMMA N)
constexpr int BK = 32;        // floats per stage row = 128 bytes = one swizzle atom
constexpr int UK = 8;         // UMMA K for tf32
constexpr int STAGES = 4;
constexpr int A_BYTES = BM * BK * 4;   // 16 KB
constexpr int B_BYTES = BN * BK * 4;   // 32 KB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024;   // + slack for 1024-byte alignment
constexpr int THREADS = 192;  // 6 warps
constexpr int TMEM_COLS = 512;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// bounded wait: a protocol bug traps (-> CUDA error) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t ok = 0;
  uint64_t t0 = 0;
  for (uint32_t spin = 0;; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    if (ok) return;
    if ((spin & 1023u) == 1023u) {   // time-based bound: 2 s without progress -> trap
      uint64_t t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      if (t0 == 0) t0 = t;
      else if (t - t0 > 2000000000ull) asm volatile("trap;");
    }
  }
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// K-major operand tile, rows of 128 bytes, 128-byte swizzle (what TMA SWIZZLE_128B writes):
// start address >> 4, LBO = 1 (unused for swizzled K-major), SBO = 1024 B (8 rows), version 1, layout 2
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= uint64_t((saddr & 0x3FFFFu) >> 4);
  d |= uint64_t(1) << 16;
  d |= uint64_t(1024 >> 4) << 32;
  d |= uint64_t(1) << 46;
  d |= uint64_t(2) << 61;
  return d;
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
// split form: issue the load, wait later (the wait names the registers so that no use can be scheduled above it)
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&r)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]),
                 "+r"(r[15]), "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]),
                 "+r"(r[22]), "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]),
                 "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :
               : "memory");
}
}  // namespace tf

// ------------------------------------------------------------------------------------------
// vpart[split, c] = sum over this split's 256-row blocks of (Ks[c,:] . Linv[i,:])^2
// grid (qc_pad/128, nsplit), 192 threads, 1 CTA / SM.
// ------------------------------------------------------------------------------------------
// ------------------------------------------------------------------------------------------
// Paired variant: each K* stage feeds TWO row blocks (both TMEM accumulators are live at once), so
// the K* strip of a candidate tile is read from HBM/L2 once per PAIR of row blocks instead of once
// per block (half the DRAM traffic of k_vnorm_tf32, 17 % less L2->SM traffic).  Stages are
// 128x16 (K*) + 2 x 256x16 (L^-1) floats with 64-byte swizzle = 40 KB, 5 stages.
// Blocks of CTA y: y, y+S, y+2S, ... (S = gridDim.y) taken two at a time; block b only contracts
// k < 256(b+1).  Accumulation slots as in k_vnorm_tf32 (bit-identical results).
// ------------------------------------------------------------------------------------------
namespace tfp {
constexpr int BM = 128, BN = 256, BK = 32, UK = 16, STAGES = 5;   // BK binary16 values = 64-byte rows
constexpr int A_BYTES = BM * BK * 2;      // 8 KB
constexpr int B_BYTES = BN * BK * 2;      // 16 KB
constexpr int STAGE_BYTES = A_BYTES + 2 * B_BYTES;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024;
// K-major tile, rows of 64 bytes, 64-byte swizzle: SBO = 512 B (8 rows), layout type 4
__device__ __forceinline__ uint64_t make_desc64(uint32_t saddr) {
  uint64_t d = 0;
  d |= uint64_t((saddr & 0x3FFFFu) >> 4);
  d |= uint64_t(1) << 16;
  d |= uint64_t(512 >> 4) << 32;
  d |= uint64_t(1) << 46;
  d |= uint64_t(4) << 61;
  return d;
}
}  // namespace tfp

// The panel contracts correlations (K* / sigma^2, in (0, 1]) with 2^e L^-1 (k_make_f16): its sums of squares are
// 2^2e / sigma^4 times |L^-1 k*|^2.  lscale[0] = 2^-2e (a power of two: the product below is exact in float64).
__device__ __forceinline__ double panel_factor(const double* __restrict__ hyp, int d, const float* __restrict__ lscale) {
  const double s2 = hyp[2 * d];
  return s2 * s2 * double(__ldg(lscale));
}

// ------------------------------------------------------------------------------------------
// 2-SM variant (tcgen05 cta_group::2): a CTA pair computes M = 256 candidates x N = 256 rows per
// instruction.  Each CTA stages its own 128-candidate K* tile and only HALF (128 rows) of every
// L^-1 tile; the tensor cores read the other half from the peer SM, so operand bytes ingested per
// SM and per flop drop by a third versus the single-CTA form of round 1 (that ingest, ~14 TB/s chip-wide, is
// what bounds the 1-SM kernels).  Stages are 3 x 8 KB = 24 KB, 8 stages.
//   * both CTAs run a TMA producer; all loads signal the LEADER's full barrier (count 2, each
//     producer arrives with its own expect_tx, the peer remotely through mapa)
//   * only the leader's warp 1 issues tcgen05.mma.cta_group::2; tcgen05.commit.cta_group::2
//     multicasts the "stage free" / "accumulator full" arrivals to both CTAs
//   * each CTA's epilogue drains its own TMEM half and arrives on the leader's "accumulator empty"
//     barrier (count 256)
// Two row blocks share every K* stage (both TMEM accumulators are live at once); block b only contracts
// k < 256 (b + 1); partial sums go to fixed slots (bit-identical results for every split and chunking).
// ------------------------------------------------------------------------------------------
namespace tf2 {
constexpr int BM = 128, BN = 256, BK = 32, UK = 16, STAGES = 8;   // BK binary16 values = 64-byte rows
constexpr int A_BYTES = BM * BK * 2;            // 8 KB
constexpr int BH_BYTES = (BN / 2) * BK * 2;     // 8 KB: this CTA's half of one L^-1 tile
constexpr int STAGE_BYTES = A_BYTES + 2 * BH_BYTES;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024;
__device__ __forceinline__ uint32_t map_to_cta(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_expect_tx_cluster(uint32_t cluster_addr, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cluster.b64 _, [%0], %1;" ::"r"(cluster_addr), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(void* dst, const CUtensorMap* map, uint32_t leader_bar, int c0,
                                                int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      ::"r"(tf::smem_u32(dst)), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void mma_f16_2sm(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t p;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(p));
  return p != 0;
}
// the same as tf::mbar_wait / commit_2sm with the barrier's shared address computed by the caller (once, outside its
// loop: the conversion of a __shared__ pointer costs three uniform-datapath instructions every time)
__device__ __forceinline__ void mbar_wait_addr(uint32_t addr, uint32_t parity) {
  uint32_t ok = 0;
  uint64_t t0 = 0;
  for (uint32_t spin = 0;; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    if (ok) return;
    if ((spin & 1023u) == 1023u) {   // time-based bound: 2 s without progress -> trap
      uint64_t t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      if (t0 == 0) t0 = t;
      else if (t - t0 > 2000000000ull) asm volatile("trap;");
    }
  }
}
__device__ __forceinline__ void commit_2sm_addr(uint32_t addr) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(addr), "h"(uint16_t(3))
               : "memory");
}
__device__ __forceinline__ void commit_2sm(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(tf::smem_u32(bar)), "h"(uint16_t(3))
               : "memory");
}
}  // namespace tf2

template <bool PERSIST>
__global__ void __launch_bounds__(tf::THREADS, 1)
k_vnorm_tf32_2sm(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, int np,
                 int64_t qc_pad, double* __restrict__ vpart, int shrink, const int* __restrict__ sched, int maxj,
                 const double* __restrict__ hyp, int d, const float* __restrict__ lscale) {
  using namespace tf;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[tf2::STAGES], empty_bar[tf2::STAGES], tfull_bar[2], tempty_bar[2];
  __shared__ uint32_t tmem_base_sh;
  uint32_t cta_rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta_rank));
  const bool leader = cta_rank == 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));
  const int nblk = (np + tf2::BN - 1) / tf2::BN;
  const int S = gridDim.y;
  // Work of this CTA pair.  Legacy form: the candidate tile blockIdx.x and the row-block pairs (b0, b0 + S),
  // b0 = blockIdx.y, blockIdx.y + 2S, ...  Persistent form: item j of the pair's list in `sched`
  // (k_panel_schedule): item = 64 * tile pair + row pair r, row blocks (2r, 2r + 1); -1 ends the list.
  auto item_at = [&](int j, int& b0, int& b1, int& c0) -> bool {
    if (PERSIST) {
      const int it = __ldg(sched + int64_t(blockIdx.x >> 1) * maxj + j);
      if (it < 0) return false;
      c0 = ((it >> 6) * 2 + int(cta_rank)) * tf2::BM;
      b0 = (it & 63) * 2;
      b1 = b0 + 1;
      return true;
    }
    b0 = int(blockIdx.y) + 2 * S * j;
    b1 = b0 + S;
    c0 = int(blockIdx.x) * tf2::BM;
    return b0 < nblk;
  };

  if (threadIdx.x == 0) {
    for (int s = 0; s < tf2::STAGES; ++s) {
      mbar_init(&full_bar[s], 2);      // one arrive(+expect_tx) from each CTA's producer (used in the leader)
      mbar_init(&empty_bar[s], 1);     // one multicast commit per phase
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], 256);  // epilogue threads of both CTAs (used in the leader)
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {   // same warp id in both CTAs
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_sh)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_sh;

  if (warp == 0) {
    {   // ===== TMA producer (both CTAs): the whole warp walks the loop, an elect.sync lane issues =====
      const bool el = tf2::elect_one();
      const uint32_t empty0 = smem_u32(&empty_bar[0]);
      const uint32_t lfull0 = tf2::map_to_cta(smem_u32(&full_bar[0]), 0);   // the LEADER's full barriers
      const int half = int(cta_rank) * (tf2::BN / 2);
      int stage = 0;
      uint32_t phase = 0;
      uint8_t* sa = sgen;
      // one stage: K* tile of this CTA's candidates + this CTA's half (128 rows, starting at r0 / r1) of the L^-1
      // tiles of the active row blocks
      auto do_stage = [&](int k, int c0, bool act0, bool act1, int r0, int r1) {
        tf2::mbar_wait_addr(empty0 + uint32_t(stage) * 8u, phase ^ 1);
        if (el) {
          const uint32_t lbar = lfull0 + uint32_t(stage) * 8u;
          tf2::mbar_expect_tx_cluster(lbar, tf2::A_BYTES + (act0 ? tf2::BH_BYTES : 0) + (act1 ? tf2::BH_BYTES : 0));
          tf2::tma_load_2d_2sm(sa, &mapA, lbar, k, c0);
          if (act0) tf2::tma_load_2d_2sm(sa + tf2::A_BYTES, &mapB, lbar, k, r0);
          if (act1) tf2::tma_load_2d_2sm(sa + tf2::A_BYTES + tf2::BH_BYTES, &mapB, lbar, k, r1);
        }
        sa += tf2::STAGE_BYTES;
        if (++stage == tf2::STAGES) {
          stage = 0;
          phase ^= 1;
          sa = sgen;
        }
      };
      for (int j = 0;; ++j) {
        int b0, b1, c0;
        if (!item_at(j, b0, b1, c0)) break;
        const bool has1 = b1 < nblk;
        const int kend0 = min((b0 + 1) * tf2::BN, np);
        const int kend = has1 ? min((b1 + 1) * tf2::BN, np) : kend0;
        const int kd0 = b0 * tf2::BN, kd1 = has1 ? min(b1 * tf2::BN, kend) : kend;
        const int r0 = b0 * tf2::BN + half, r1 = b1 * tf2::BN + half;
        // diagonal block, kk columns in: the MMA takes N = 256 - kk, i.e. the first 128 - kk/2 rows of either
        // CTA's tile, and accumulator column c must stay physical row c: the second CTA's box starts kk/2 rows
        // earlier (its first rows are then physical rows 128 - kk/2 ..)
        const int dsh = (shrink && cta_rank != 0) ? tf2::BK / 2 : 0;   // rows the box moves up per diagonal stage
        int k = 0;
        for (; k < kd0; k += tf2::BK) do_stage(k, c0, true, has1, r0, r1);
        for (int sh = 0; k < kend0; k += tf2::BK, sh += dsh) do_stage(k, c0, true, has1, r0 - sh, r1);
        if (has1) {
          for (; k < kd1; k += tf2::BK) do_stage(k, c0, false, true, 0, r1);
          for (int sh = 0; k < kend; k += tf2::BK, sh += dsh) do_stage(k, c0, false, true, 0, r1 - sh);
        }
      }
    }
  } else if (warp == 1) {
    if (leader) {   // ===== MMA issuer (leader CTA only) =====
      // D=f32, A=B=f16, K-major, N=256, M=256 (two SMs).  In the diagonal 256 x 256 block of a row block the rows
      // above the stage's first column are zero and sit at the end of the block (k_make_f16): N shrinks by 32
      // per stage there, so the diagonal blocks cost 17/32 of a full block instead of 1 (BBO_TF32_FULL_DIAG=1: off).
      //
      // ONE thread feeds the tensor cores of both SMs, and it used to be what bounded the kernel: ~155 SASS
      // instructions per stage (shared-window address conversions, descriptor construction, the diagonal-block test,
      // a vote loop around every uniform-datapath instruction because the compiler could not see a single-lane
      // region) against 512 tensor-pipe cycles of work -- the pipe was busy 65 % of the time with TF32 and with
      // binary16 operands alike.  Now the whole warp runs the loop and an elect.sync lane issues; barrier addresses
      // and the stage-0 descriptor are computed once, a stage advances them by constants, and an item is walked in
      // three phases (both blocks full / first block diagonal / second block alone) so that no per-stage test of the
      // diagonal remains.
      const bool el = tf2::elect_one();
      const uint32_t idesc_base = (1u << 4) | (uint32_t(256 >> 4) << 24);   // D = f32, A = B = f16 (format 0)
      const uint32_t idesc_full = idesc_base | (uint32_t(tf2::BN >> 3) << 17);
      constexpr uint32_t ID_STEP = uint32_t(tf2::BK >> 3) << 17;            // N shrinks by BK per diagonal stage
      const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
      const uint32_t tfull0 = smem_u32(&tfull_bar[0]), tempty0 = smem_u32(&tempty_bar[0]);
      const uint64_t desc0 = tfp::make_desc64(sbase);
      constexpr uint64_t STAGE_D = uint64_t(tf2::STAGE_BYTES >> 4), A_D = uint64_t(tf2::A_BYTES >> 4),
                         BH_D = uint64_t(tf2::BH_BYTES >> 4);
      int stage = 0;
      uint32_t phase = 0, acc_phase = 0, acc1_phase = 0;   // the second accumulator idles in an item without b1
      uint64_t adesc = desc0;
      const uint32_t tm0 = tmem_base, tm1 = tmem_base + uint32_t(tf2::BN);
      // one stage: wait for the operands, 2 MMAs (K = 2 x 16) per active accumulator, release the stage
      auto do_stage = [&](bool act0, bool act1, uint32_t id0, uint32_t id1, uint32_t accum) {
        tf2::mbar_wait_addr(full0 + uint32_t(stage) * 8u, phase);
        tc_fence_after();
        if (el) {
          if (act0) {
            tf2::mma_f16_2sm(tm0, adesc, adesc + A_D, id0, accum);
            tf2::mma_f16_2sm(tm0, adesc + 2, adesc + A_D + 2, id0, 1u);
          }
          if (act1) {
            tf2::mma_f16_2sm(tm1, adesc, adesc + A_D + BH_D, id1, accum);
            tf2::mma_f16_2sm(tm1, adesc + 2, adesc + A_D + BH_D + 2, id1, 1u);
          }
          tf2::commit_2sm_addr(empty0 + uint32_t(stage) * 8u);       // frees the stage in both CTAs
        }
        adesc += STAGE_D;
        if (++stage == tf2::STAGES) {
          stage = 0;
          phase ^= 1;
          adesc = desc0;
        }
      };
      for (int j = 0;; ++j) {
        int b0, b1, c0;
        if (!item_at(j, b0, b1, c0)) break;
        const bool has1 = b1 < nblk;
        const int kend0 = min((b0 + 1) * tf2::BN, np);
        const int kend = has1 ? min((b1 + 1) * tf2::BN, np) : kend0;
        const int kd0 = b0 * tf2::BN, kd1 = has1 ? min(b1 * tf2::BN, kend) : kend;   // starts of the diagonal blocks
        tf2::mbar_wait_addr(tempty0, acc_phase ^ 1);
        if (has1) tf2::mbar_wait_addr(tempty0 + 8u, acc1_phase ^ 1);
        tc_fence_after();
        int k = 0;
        // both blocks below their diagonal
        for (; k < kd0; k += tf2::BK) do_stage(true, has1, idesc_full, idesc_full, k != 0 ? 1u : 0u);
        // block b0 on its diagonal (N shrinks stage by stage), b1 still full
        uint32_t id0 = idesc_full;
        for (; k < kend0; k += tf2::BK) {
          do_stage(true, has1, id0, idesc_full, k != 0 ? 1u : 0u);
          if (shrink) id0 -= ID_STEP;
        }
        if (el) tf2::commit_2sm_addr(tfull0);                        // block b0 complete in both halves
        if (has1) {
          // b1 alone: full below its diagonal (legacy form only: b1 > b0 + 1), then its diagonal
          for (; k < kd1; k += tf2::BK) do_stage(false, true, idesc_full, idesc_full, 1u);
          uint32_t id1 = idesc_full;
          for (; k < kend; k += tf2::BK) {
            do_stage(false, true, idesc_full, id1, 1u);
            if (shrink) id1 -= ID_STEP;
          }
          if (el) tf2::commit_2sm_addr(tfull0 + 8u);
          acc1_phase ^= 1;
        }
        acc_phase ^= 1;
      }
    }
  } else {   // ===== epilogue (warps 2..5 of both CTAs) =====
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    double tot0 = 0.0, tot1 = 0.0, tot2 = 0.0, tot3 = 0.0;
    const double fac = panel_factor(hyp, d, lscale);
    uint32_t acc_ph[2] = {0, 0};
    for (int j = 0;; ++j) {
      int b0, b1, c0;
      if (!item_at(j, b0, b1, c0)) break;
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        const int nb = a == 0 ? b0 : b1;
        if (nb >= nblk) break;
        mbar_wait(&tfull_bar[a], acc_ph[a]);
        acc_ph[a] ^= 1;
        tc_fence_after();
        const uint32_t taddr = tmem_base + (uint32_t(quarter * 32) << 16) + uint32_t(a * tf2::BN);
        float part = 0.f;
#pragma unroll 2
        for (int c = 0; c < tf2::BN; c += 32) {
          float v[32];
          tmem_ld32(taddr + uint32_t(c), v);
          float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            s0 = fmaf(v[i], v[i], s0);
            s1 = fmaf(v[i + 1], v[i + 1], s1);
            s2 = fmaf(v[i + 2], v[i + 2], s2);
            s3 = fmaf(v[i + 3], v[i + 3], s3);
          }
          part += (s0 + s1) + (s2 + s3);
        }
        if (PERSIST) {   // one slot per row block; k_acq / k_acq_topk add them with the four-slot association
          vpart[int64_t(nb) * qc_pad + c0 + row] = double(part) * fac;
        } else {
          const int slot = nb & 3;
          const double dp = double(part) * fac;
          if (slot == 0) tot0 += dp;
          else if (slot == 1) tot1 += dp;
          else if (slot == 2) tot2 += dp;
          else tot3 += dp;
        }
        tc_fence_before();
        tf2::mbar_arrive_cluster(tf2::map_to_cta(smem_u32(&tempty_bar[a]), 0));   // leader's barrier
      }
    }
    if (!PERSIST) {
      const int c0 = int(blockIdx.x) * tf2::BM;
      const double tots[4] = {tot0, tot1, tot2, tot3};
#pragma unroll
      for (int g4 = 0; g4 < 4; ++g4)
        if (g4 % S == int(blockIdx.y)) vpart[int64_t(g4) * qc_pad + c0 + row] = tots[g4];
    }
  }
  tc_fence_before();
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------------------------------
// Work list of the persistent panel (one warp).  Items (tile pair tp, row pair r) cost ~(r + 1) units (row blocks
// 2r and 2r + 1 contract 512 (r + 1) columns of K*).  They are taken in an order that keeps the pairs that are in
// flight at any time on FEW K* strips -- groups of `gs` tile pairs (default: strips of ~4 MB, i.e. one tile pair at
// n = 4096; measured 32.65 / 32.83 / 32.86 / 34.4 ms per step for 4 / 9 / 13 / 36 MB), inside a group the long row
// pairs first -- and
// dealt by list scheduling: every item goes to the CTA pair with the least work so far (lowest index on ties).
// That is what a run-time work queue would do under this cost model, without any in-kernel hand-over.  The K*
// strip of a tile pair (256 x np floats) is read by all of its row pairs while it is still in L2, instead of
// 4.5 times from HBM.  sched[c * maxj + j] = 64 tp + r, -1 terminated; a list holds at most maxj - 1 items and a full
// list is passed over (nc * (maxj - 1) >= ntp * nrp is the caller's side of the contract).
// ------------------------------------------------------------------------------------------
__global__ void k_panel_schedule(int ntp, int nrp, int nc, int gs, int maxj, int* __restrict__ sched) {
  const int lane = threadIdx.x;
  int load[3] = {0, 0, 0}, cnt[3] = {0, 0, 0};   // clusters lane, lane + 32, lane + 64
  for (int g0 = 0; g0 < ntp; g0 += gs) {
    const int ge = min(g0 + gs, ntp);
    for (int r = nrp - 1; r >= 0; --r) {
      for (int tp = g0; tp < ge; ++tp) {
        // arg-min of the loads (lowest cluster index on ties)
        int best = 0x7fffffff, bc = 0;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const int c = lane + 32 * i;
          if (c < nc && cnt[i] < maxj - 1 && load[i] < best) {   // a full list takes no more items
            best = load[i];
            bc = c;
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const int ob = __shfl_xor_sync(0xffffffffu, best, o), oc = __shfl_xor_sync(0xffffffffu, bc, o);
          if (ob < best || (ob == best && oc < bc)) {
            best = ob;
            bc = oc;
          }
        }
        if ((bc & 31) == lane) {
          const int i = bc >> 5;
#pragma unroll
          for (int ii = 0; ii < 3; ++ii)
            if (ii == i) {
              sched[int64_t(bc) * maxj + cnt[ii]] = tp * 64 + r;
              cnt[ii] += 1;
              load[ii] += 4 * (r + 1) + 1;
            }
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int c = lane + 32 * i;
    if (c < nc) sched[int64_t(c) * maxj + min(cnt[i], maxj - 1)] = -1;
  }
}

// ------------------------------------------------------------------------------------------
// K* (qc_pad, np) binary16 correlations and mu (CUDA cores, any d).  grid (qc_pad/64), 256 threads.
// ------------------------------------------------------------------------------------------
constexpr int kFC = 32, kFS = kFC + 1;

template <typename T1>
__global__ void __launch_bounds__(256)
k_kstar_f32(bbo_gp_hyper h, const T1* __restrict__ Xq, int64_t qc, const double* __restrict__ X, int64_t n,
            int64_t np, const double* __restrict__ hyp, const double* __restrict__ alpha, __half* __restrict__ Ks,
            double* __restrict__ mu) {
  __shared__ float x1s[64 * kFS], x2s[64 * kFS], ils[kFC];
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int64_t c0 = int64_t(blockIdx.x) * 64;
  const int d = h.d;
  const bool matern = h.kind == BBO_KERNEL_MATERN52;
  const float eps = matern ? float(h.pairwise_eps) : 0.f;
  const float s2 = float(hyp[2 * d]);
  float macc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int64_t j0 = 0; j0 < np; j0 += 64) {
    float s[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) s[a][b] = 0.f;
    for (int kc = 0; kc < d; kc += kFC) {
      const int kw = min(kFC, d - kc);
      __syncthreads();
      for (int idx = tid; idx < 64 * kFC; idx += 256) {
        int r = idx / kFC, k = idx % kFC;
        float v1 = 0.f, v2 = 0.f;
        if (k < kw) {
          if (c0 + r < qc) v1 = float(Xq[(c0 + r) * d + kc + k]);
          if (j0 + r < n) v2 = float(X[(j0 + r) * d + kc + k]);
        }
        x1s[r * kFS + k] = v1;
        x2s[r * kFS + k] = v2;
      }
      if (tid < kFC) ils[tid] = tid < kw ? float(hyp[d + kc + tid]) : 0.f;
      __syncthreads();
      for (int k = 0; k < kw; ++k) {
        const float il = ils[k];
        float av[4], bv[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) av[a] = x1s[(ty * 4 + a) * kFS + k];
#pragma unroll
        for (int b = 0; b < 4; ++b) bv[b] = x2s[(tx + 16 * b) * kFS + k];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            float de = fmaf(bv[b] - av[a], il, eps);
            s[a][b] = fmaf(de, de, s[a][b]);
          }
      }
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int64_t j = j0 + tx + 16 * b;
        float v = 0.f;
        if (j < n) {
          if (matern) {
            float r = 2.2360679775f * sqrtf(s[a][b]);
            v = (1.f + r + r * r * (1.f / 3.f)) * __expf(-r);
          } else {
            v = __expf(-0.5f * s[a][b]);
          }
        }
        macc[a] = fmaf(v, float(alpha[j]), macc[a]);
        Ks[(c0 + ty * 4 + a) * np + j] = __float2half_rn(v);      // correlation; sigma^2 enters mu and the panel factor
      }
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    float v = macc[a];
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (tx == 0) mu[c0 + ty * 4 + a] = hyp[2 * d + 1] + double(s2) * double(v);
  }
}

// ------------------------------------------------------------------------------------------
// Pre-scaled float32 training set for the fast K* kernel: Xs[j,k] = float(X[j,k] / l_k) (scaled in
// float64, rounded once), zero rows on the padding; bnorm[j] = |Xs[j,:]|^2; alpha32 = float(alpha).
// grid (np/8), 256 threads: one warp per training row.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_prescale(const double* __restrict__ X, int64_t n, int64_t np, int d, const double* __restrict__ hyp,
           const double* __restrict__ alpha, float* __restrict__ Xs, float* __restrict__ bnorm,
           float* __restrict__ alpha32) {
  const int64_t j = int64_t(blockIdx.x) * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (j >= np) return;
  float acc = 0.f;
  for (int k = lane; k < d; k += 32) {
    const float v = j < n ? float(X[j * d + k] * hyp[d + k]) : 0.f;
    Xs[j * d + k] = v;
    acc = fmaf(v, v, acc);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) {
    bnorm[j] = acc;
    alpha32[j] = float(alpha[j]);
  }
}

// ------------------------------------------------------------------------------------------
// Fast K* for d <= kFastMaxD (TF32 mode only; tolerance 1e-3):
//   s = |a'|^2 + |b|^2 - 2 a'.b,  a' = xq/l - eps, b = x/l   ( == sum ((x - xq)/l + eps)^2 up to
//   ~5e-6 absolute float32 cancellation error, far inside this mode's gate; the FP64 mode keeps
//   the exact-difference form).  One FFMA per (candidate, observation, dimension).
// 128 candidates per CTA, 8x4 register tile per thread, float4 shared loads along the feature
// axis (row stride = smallest 4*odd >= d+4 floats), candidate tile staged once, training tiles loaded into a
// double-buffered smem tile (one barrier per tile).  grid (qc_pad/128), 256 threads, dynamic smem.
// ------------------------------------------------------------------------------------------
template <typename T1>
__global__ void __launch_bounds__(256)
k_kstar_f32_fast(bbo_gp_hyper h, const T1* __restrict__ Xq, int64_t qc, const float* __restrict__ Xs,
                 const float* __restrict__ bnorm, int64_t n, int64_t np, const double* __restrict__ hyp,
                 const float* __restrict__ alpha32, __half* __restrict__ Ks, double* __restrict__ mu) {
  extern __shared__ __align__(16) float ksm[];
  const int d = h.d;
  const int d4 = (d + 3) >> 2;
  const int LS = 4 * ((d4 + 1) | 1);      // row stride: LS/4 odd -> float4 reads conflict free per quarter warp
  float* x1s = ksm;                       // 128 x LS
  float* x2s0 = x1s + 128 * LS;           // 2 x 64 x LS
  float* als = x2s0 + 2 * 64 * LS;        // 2 x 64
  float* bns = als + 128;                 // 2 x 64
  float* ans = bns + 128;                 // 128
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int64_t c0 = int64_t(blockIdx.x) * 128;
  const bool matern = h.kind == BBO_KERNEL_MATERN52;
  const float eps = matern ? float(h.pairwise_eps) : 0.f;
  const float s2 = float(hyp[2 * d]);
  const int dp = 4 * d4;                  // padded feature count (zeros beyond d)
  for (int idx = tid; idx < 128 * dp; idx += 256) {
    const int r = idx / dp, k = idx % dp;
    float v = 0.f;
    if (k < d) {
      const double x = (c0 + r < qc) ? double(Xq[(c0 + r) * d + k]) : 0.0;
      v = float(x * hyp[d + k]) - eps;
    }
    x1s[r * LS + k] = v;
  }
  // training tile: 64 x dp floats straight into the idle buffer (visible after the tile barrier)
  auto tload = [&](int64_t j0, int buf) {
    float* xb = x2s0 + buf * 64 * LS;
    for (int idx = tid; idx < 64 * dp; idx += 256) {
      const int r = idx / dp, k = idx % dp;
      xb[r * LS + k] = (k < d) ? Xs[(j0 + r) * d + k] : 0.f;
    }
    if (tid < 64) {
      als[buf * 64 + tid] = alpha32[j0 + tid];
      bns[buf * 64 + tid] = bnorm[j0 + tid];
    }
  };
  tload(0, 0);
  __syncthreads();
  if (tid < 128) {
    float acc = 0.f;
    for (int k = 0; k < d; ++k) acc = fmaf(x1s[tid * LS + k], x1s[tid * LS + k], acc);
    ans[tid] = acc;
  }
  __syncthreads();
  float an[8], macc[8];
#pragma unroll
  for (int a = 0; a < 8; ++a) {
    an[a] = ans[ty + 16 * a];
    macc[a] = 0.f;
  }
  const int ntile = int(np / 64);
  for (int t = 0; t < ntile; ++t) {
    const int buf = t & 1;
    const int64_t j0 = int64_t(t) * 64;
    if (t + 1 < ntile) tload(j0 + 64, buf ^ 1);
    float acc[8][4];
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
    const float* xb = x2s0 + buf * 64 * LS;
    for (int k4 = 0; k4 < d4; ++k4) {
      float4 av[8], bv[4];
#pragma unroll
      for (int a = 0; a < 8; ++a) av[a] = *reinterpret_cast<const float4*>(&x1s[(ty + 16 * a) * LS + 4 * k4]);
#pragma unroll
      for (int b = 0; b < 4; ++b) bv[b] = *reinterpret_cast<const float4*>(&xb[(tx + 16 * b) * LS + 4 * k4]);
#pragma unroll
      for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          acc[a][b] = fmaf(av[a].x, bv[b].x, acc[a][b]);
          acc[a][b] = fmaf(av[a].y, bv[b].y, acc[a][b]);
          acc[a][b] = fmaf(av[a].z, bv[b].z, acc[a][b]);
          acc[a][b] = fmaf(av[a].w, bv[b].w, acc[a][b]);
        }
    }
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int64_t j = j0 + tx + 16 * b;
      const float al = als[buf * 64 + tx + 16 * b], bn = bns[buf * 64 + tx + 16 * b];
      const bool valid = j < n;
#pragma unroll
      for (int a = 0; a < 8; ++a) {
        float v = 0.f;
        if (valid) {
          const float sq = fmaxf(fmaf(-2.f, acc[a][b], an[a] + bn), 0.f);
          if (matern) {
            const float r = 2.2360679775f * sqrtf(sq);
            v = (1.f + r + r * r * (1.f / 3.f)) * __expf(-r);
          } else {
            v = __expf(-0.5f * sq);
          }
        }
        macc[a] = fmaf(v, al, macc[a]);
        Ks[(c0 + ty + 16 * a) * np + j] = __float2half_rn(v);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int a = 0; a < 8; ++a) {
    float v = macc[a];
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (tx == 0) mu[c0 + ty + 16 * a] = hyp[2 * d + 1] + double(s2) * double(v);
  }
}

// ---- tensor-core K* producers: shared helpers
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x));
  const float r = x - __uint_as_float(hi);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(r));
}
// two floats -> one word of binary16 (round to nearest even), `lo` at the lower address
__device__ __forceinline__ uint32_t pack_half2(float lo, float hi) {
  uint32_t u;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(u) : "f"(hi), "f"(lo));
  return u;
}
// cvt.rna.tf32.f32 of a finite value with two integer-pipe instructions (round to nearest, ties away:
// add half an ulp of the 13 dropped bits, clear them).  The conversion instruction runs on the XU pipe,
// which the K* producers already load with one MUFU.EX2 per value.
__device__ __forceinline__ uint32_t round_tf32_bits(float x) {
  return (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
}
__device__ __forceinline__ void mma_tf32_16x8x8(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// ------------------------------------------------------------------------------------------
// K* producer on mma.sync (default for 30 < d <= 126; d <= 30 uses the tcgen05 producer k_kstar_umma below).
// The scaled dot products a'.b are computed with error-compensated 3xTF32 (a_hi b_hi + a_hi b_lo + a_lo b_hi,
// relative error ~2^-21) on mma.sync.m16n8k8.tf32 (SASS HMMA.1688.F32.TF32); the contraction is
// AUGMENTED by two columns so that the tensor core returns the finished exponent / distance:
//   RBF     a = [log2e * a', A', 1, 0..]   b = [b, 1, B', 0..]   a.b = log2(s2) - log2e/2 |a'-b|^2
//           with A' = log2(s2) - log2e/2 |a'|^2, B' = -log2e/2 |b|^2        ->  k = ex2(a.b)
//   Matern  a = [-2 a', |a'|^2, 1, 0..]    b = [b, 1, |b|^2, 0..] a.b = |a'-b|^2
// For d = 20 the two columns live in the zero padding of the 8-wide k-steps, so they are free.
// Padded observations carry B' = -1e4 (|b|^2 = 1e8): their K* is exactly 0 without a predicate.
// B fragments come from ldmatrix.x4 (hi/lo x two k-halves per instruction; the b16 view of a row of
// four 32-bit words hands lane (g,t) word t of row g, which is the tf32 B fragment), ex2 is the bare
// MUFU.  Per 16x8 output tile: 3*KS HMMA + KS LDSM + 1 LDS.64 + 4 MUFU + 4 CVT + 4 FFMA + 2 STG.64
// ------------------------------------------------------------------------------------------
constexpr int kMma2MaxD = 126;
__global__ void __launch_bounds__(256)
k_prescale_aug(const double* __restrict__ X, int64_t n, int64_t np, int d, int dk, int matern,
               const double* __restrict__ hyp, const double* __restrict__ alpha, uint32_t* __restrict__ Xhi,
               uint32_t* __restrict__ Xlo, float* __restrict__ alpha32) {
  const int64_t j = int64_t(blockIdx.x) * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (j >= np) return;
  float acc = 0.f;
  for (int k = lane; k < d; k += 32) {
    const float v = (j < n) ? float(X[j * d + k] * hyp[d + k]) : 0.f;
    uint32_t hi, lo;
    split_tf32(v, hi, lo);
    Xhi[j * dk + k] = hi;
    Xlo[j * dk + k] = lo;
    acc = fmaf(v, v, acc);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  for (int k = d + lane; k < dk; k += 32) {
    float v = 0.f;
    if (k == d) v = 1.f;
    else if (k == d + 1) v = (j < n) ? (matern ? acc : -0.5f * 1.4426950408889634f * acc) : (matern ? 1e8f : -1e4f);
    uint32_t hi, lo;
    split_tf32(v, hi, lo);
    Xhi[j * dk + k] = hi;
    Xlo[j * dk + k] = lo;
  }
  if (lane == 0) alpha32[j] = (j < n) ? float(alpha[j]) : 0.f;
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void ldsm_x4(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, unsigned addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}

template <typename T1, int KS, bool MATERN>   // KS = ceil((d+2)/8) k-steps
__global__ void __launch_bounds__(256)
k_kstar_mma2(bbo_gp_hyper h, const T1* __restrict__ Xq, int64_t qc, const uint32_t* __restrict__ Xhi,
             const uint32_t* __restrict__ Xlo, int64_t np, const double* __restrict__ hyp,
             const float* __restrict__ alpha32, __half* __restrict__ Ks, double* __restrict__ mu) {
  constexpr int DK = KS * 8;
  constexpr int LS = 4 * ((DK / 4 + 1) | 1);   // LS/4 odd: ldmatrix rows hit distinct 16-byte bank groups
  extern __shared__ __align__(16) uint32_t msm2[];
  uint32_t* bhi = msm2;                  // 2 x 64 x LS
  uint32_t* blo = bhi + 2 * 64 * LS;     // 2 x 64 x LS
  float* als = reinterpret_cast<float*>(blo + 2 * 64 * LS);   // 2 x 64
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int d = h.d;
  const float eps = MATERN ? float(h.pairwise_eps) : 0.f;
  const float s2 = float(hyp[2 * d]);
  const float kLog2e = 1.4426950408889634f;
  const int64_t row0 = int64_t(blockIdx.x) * 128 + warp * 16;   // this warp's 16 candidates

  // A fragments: a0 (g, t), a1 (g+8, t), a2 (g, t+4), a3 (g+8, t+4) per k-step; a' = xq/l - eps
  uint32_t ahi[KS][4], alo[KS][4];
  {
    float av[KS][4];
    float an0 = 0.f, an1 = 0.f;   // |a'|^2 of rows g and g+8
#pragma unroll
    for (int ks = 0; ks < KS; ++ks)
#pragma unroll
      for (int f = 0; f < 4; ++f) {
        const int64_t r = row0 + g + ((f & 1) ? 8 : 0);
        const int k = ks * 8 + t + ((f & 2) ? 4 : 0);
        float v = 0.f;
        if (k < d) {
          const double x = (r < qc) ? double(Xq[r * d + k]) : 0.0;
          v = float(x * hyp[d + k]) - eps;
        }
        av[ks][f] = v;
        if (f & 1) an1 = fmaf(v, v, an1);
        else an0 = fmaf(v, v, an0);
      }
    an0 += __shfl_xor_sync(0xffffffffu, an0, 1);
    an0 += __shfl_xor_sync(0xffffffffu, an0, 2);
    an1 += __shfl_xor_sync(0xffffffffu, an1, 1);
    an1 += __shfl_xor_sync(0xffffffffu, an1, 2);
    const float R0 = MATERN ? an0 : -0.5f * kLog2e * an0;      // correlations: no sigma^2 in the exponent
    const float R1 = MATERN ? an1 : -0.5f * kLog2e * an1;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks)
#pragma unroll
      for (int f = 0; f < 4; ++f) {
        const int k = ks * 8 + t + ((f & 2) ? 4 : 0);
        float u = 0.f;
        if (k < d) u = (MATERN ? -2.f : kLog2e) * av[ks][f];
        else if (k == d) u = (f & 1) ? R1 : R0;
        else if (k == d + 1) u = 1.f;
        split_tf32(u, ahi[ks][f], alo[ks][f]);
      }
  }

  auto tload = [&](int64_t j0, int buf) {
    constexpr int CH = DK / 4;   // 16-byte chunks per row
    for (int c = tid; c < 64 * CH * 2; c += 256) {
      const int arr = c / (64 * CH), rem = c % (64 * CH), r = rem / CH, ch = rem % CH;
      const uint32_t* src = (arr ? Xlo : Xhi) + (j0 + r) * DK + ch * 4;
      uint32_t* dst = (arr ? blo : bhi) + (buf * 64 + r) * LS + ch * 4;
      const unsigned sa = static_cast<unsigned>(__cvta_generic_to_shared(dst));
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(src));
    }
    if (tid < 64) als[buf * 64 + tid] = alpha32[j0 + tid];
    asm volatile("cp.async.commit_group;");
  };
  tload(0, 0);
  // ldmatrix lane address: matrices 0/1 = hi (k 0-3, 4-7), 2/3 = lo; lane supplies row (lane & 7)
  const unsigned lane_addr =
      static_cast<unsigned>(__cvta_generic_to_shared((lane & 16) ? blo : bhi)) +
      unsigned(((lane & 7) * LS + ((lane >> 3) & 1) * 4) * 4);
  float m0 = 0.f, m1 = 0.f;   // mu partials of rows g, g+8
  const int ntile = int(np / 64);
  __half* o0 = Ks + (row0 + g) * np + 2 * t;
  __half* o1 = Ks + (row0 + g + 8) * np + 2 * t;
  for (int tl = 0; tl < ntile; ++tl) {
    const int buf = tl & 1;
    asm volatile("cp.async.wait_group 0;");
    __syncthreads();                                   // tile tl visible; everyone is done with tile tl-1
    if (tl + 1 < ntile) tload(int64_t(tl + 1) * 64, buf ^ 1);
    const unsigned tb = lane_addr + unsigned(buf * 64 * LS * 4);
    const float* ap = als + buf * 64 + 2 * t;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      float c[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        uint32_t h0, h1, l0, l1;
        ldsm_x4(h0, h1, l0, l1, tb + unsigned((nt * 8 * LS + ks * 8) * 4));
        mma_tf32_16x8x8(c, alo[ks], h0, h1);
        mma_tf32_16x8x8(c, ahi[ks], l0, l1);
        mma_tf32_16x8x8(c, ahi[ks], h0, h1);
      }
      float v[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if (!MATERN) {
          v[e] = ex2_approx(fminf(c[e], 0.f));
        } else {
          const float r = 2.2360679775f * sqrtf(fmaxf(c[e], 0.f));
          v[e] = fmaf(r, fmaf(r, 1.f / 3.f, 1.f), 1.f) * ex2_approx(-kLog2e * r);
        }
      }
      const float2 al = *reinterpret_cast<const float2*>(ap + nt * 8);
      m0 = fmaf(v[0], al.x, fmaf(v[1], al.y, m0));
      m1 = fmaf(v[2], al.x, fmaf(v[3], al.y, m1));
      *reinterpret_cast<uint32_t*>(o0 + nt * 8) = pack_half2(v[0], v[1]);
      *reinterpret_cast<uint32_t*>(o1 + nt * 8) = pack_half2(v[2], v[3]);
    }
    o0 += 64;
    o1 += 64;
  }
  m0 += __shfl_xor_sync(0xffffffffu, m0, 1);
  m0 += __shfl_xor_sync(0xffffffffu, m0, 2);
  m1 += __shfl_xor_sync(0xffffffffu, m1, 1);
  m1 += __shfl_xor_sync(0xffffffffu, m1, 2);
  if (t == 0) {
    mu[row0 + g] = hyp[2 * d + 1] + double(s2) * double(m0);
    mu[row0 + g + 8] = hyp[2 * d + 1] + double(s2) * double(m1);
  }
}

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(tf::smem_u32(dst)), "l"(src), "r"(bytes), "r"(tf::smem_u32(bar))
               : "memory");
}

// ------------------------------------------------------------------------------------------
// K* producer on tcgen05 (d <= 126): the augmented 3xTF32 contraction of k_kstar_mma2 on the 5th-generation tensor
// cores.  Operand rows hold the TF32 hi and lo parts of the augmented vectors side by side,
//     candidates   Aop[c] = [ hi(u) | lo(u) | 0.. ]      u = the augmented candidate row of k_kstar_mma2
//     observations Bop[j] = [ hi(b) | lo(b) | 0.. ]      b as in k_prescale_aug
// (2 dk words, padded to 32-word boxes) and the three partial products u_hi b_hi + u_lo b_hi + u_hi b_lo -- the
// finished exponent (RBF) / squared distance (Matern) -- are formed by PAIRING k-steps: a `hi` step of an observation
// box meets the candidate's hi and lo steps, a `lo` step its hi step.  Persistent CTAs (one per SM), 576 threads:
//   warp 0     TMA producer: the 128-candidate A tile once per tile (two tiles resident when they fit: the next one
//              loads under this one's MMAs), observation boxes (SWIZZLE_128B, 128 rows x 32 words) through a ring
//   warp 1     one thread issues tcgen05.mma kind::tf32, M = 128 x N = 128 x K = 8, 3 dk / 8 instructions per
//              128 x 128 block, fully unrolled on the template parameter DKS = dk / 8 (every descriptor = base +
//              constant; with run-time bounds this thread was the bottleneck); four 128-column TMEM accumulators in a
//              ring, so the tensor core runs up to four blocks ahead of the epilogue
//   warps 2-17 epilogue, two groups of eight warps on alternate blocks (accumulators), a warp = one TMEM lane quarter
//              x 64 columns of its block, slab by slab: tcgen05.ld 32 lanes x 32 columns ->
//              ex2 (Matern: sqrt, cubic, ex2) -> mu += K* alpha (alpha staged one block ahead by cp.async) ->
//              cvt.rn.f16x2.f32 -> st.shared.v4 into one of two 64-byte-swizzled 32 x 32 staging tiles ->
//              TMA tensor store (cp.async.bulk.tensor ... bulk_group)
// so every K* byte leaves the SM as part of a full line (the mma.sync forms scatter 8-byte stores).
// mu = c + s2 x the four slab partials in a fixed order (deterministic).  History and counters:
// profiles/kstar_umma_r02_ncu_full.md (72 us per 37888 x 4096 chunk; 122 us in round 1 with float32 output).
// ------------------------------------------------------------------------------------------
namespace ku {
constexpr int BM = 128, BN = 128, BOXK = 32;
constexpr int BOX_BYTES = 128 * BOXK * 4;          // 16 KB: 128 rows x 128 bytes
constexpr int MAXBOX = 8;                          // operand rows [hi | lo]: 2 dk <= 256 words, dk <= 128
constexpr int MAXBOX_STREAM = 12;                  // (workspace sizing of Aop / Bop: rows of up to 384 words)
constexpr int MAXRING = 6, ACCS = 4;               // B boxes in flight, TMEM accumulators
constexpr int EPI_WARPS = 16;
constexpr int THREADS = 64 + 32 * EPI_WARPS;       // 576
constexpr int OUT_TILE = 32 * 64;                  // staging tile: 32 candidates x 32 binary16 values
constexpr int OUT_BYTES = 2 * OUT_TILE;            // two per epilogue warp: the TMA store of one drains under the next
constexpr int kMaxD = 30;                          // two A tiles resident (the next one loads under this one's MMAs)
constexpr int kMaxDStream = 126;                   // dk = 8 ceil((d+2)/8) <= 128
// Shared-memory plan for `nbox` boxes per operand row: A tiles (two when they fit), ring of B boxes, staging tiles
// (two per epilogue warp when they fit).  214 KB of dynamic shared memory + 1 KB alignment slack + 11.5 KB static.
struct Plan {
  int na, ring, stg2;
  size_t bytes;
};
static Plan plan(int nbox) {
  const int budget = 214 * 1024;
  Plan p;
  p.na = (2 * nbox * BOX_BYTES + 4 * BOX_BYTES + EPI_WARPS * OUT_BYTES <= budget) ? 2 : 1;
  p.stg2 = (p.na * nbox * BOX_BYTES + 4 * BOX_BYTES + EPI_WARPS * OUT_BYTES <= budget) ? 1 : 0;
  const int stg = EPI_WARPS * (p.stg2 ? OUT_BYTES : OUT_TILE);
  p.ring = (budget - p.na * nbox * BOX_BYTES - stg) / BOX_BYTES;
  if (p.ring > MAXRING) p.ring = MAXRING;
  p.bytes = size_t(p.na * nbox + p.ring) * BOX_BYTES + size_t(stg) + 1024;
  return p;
}
// descriptor offset of the k-step that starts at word w of an operand row stored in 32-word boxes (128-byte swizzle):
// box w / 32 (BOX_BYTES >> 4 per box in the 16-byte address field), 32 bytes per k-step inside the swizzle atom
__host__ __device__ constexpr uint64_t desc_off(int w) {
  return uint64_t((w >> 5) * (BOX_BYTES >> 4) + ((w & 31) >> 2));
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t saddr, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(map), "r"(saddr), "r"(c0), "r"(c1)
               : "memory");
}
}  // namespace ku

// rows of Bop: [hi | lo | 0..] of the augmented observation vector (see k_prescale_aug), ktp words per row
__global__ void __launch_bounds__(256)
k_prescale_aug_umma(const double* __restrict__ X, int64_t n, int64_t np, int d, int dk, int ktp, int matern,
                    const double* __restrict__ hyp, const double* __restrict__ alpha, uint32_t* __restrict__ Bop,
                    float* __restrict__ alpha32) {
  const int64_t j = int64_t(blockIdx.x) * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (j >= np) return;
  float acc = 0.f;
  for (int k = lane; k < d; k += 32) {
    const float v = (j < n) ? float(X[j * d + k] * hyp[d + k]) : 0.f;
    acc = fmaf(v, v, acc);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  uint32_t* row = Bop + j * ktp;
  for (int k = lane; k < dk; k += 32) {
    float v = 0.f;
    if (k < d) v = (j < n) ? float(X[j * d + k] * hyp[d + k]) : 0.f;
    else if (k == d) v = 1.f;
    else if (k == d + 1) v = (j < n) ? (matern ? acc : -0.5f * 1.4426950408889634f * acc) : (matern ? 1e8f : -1e4f);
    uint32_t hi, lo;
    split_tf32(v, hi, lo);
    row[k] = hi;
    row[dk + k] = lo;
  }
  for (int k = 2 * dk + lane; k < ktp; k += 32) row[k] = 0u;
  if (lane == 0) alpha32[j] = (j < n) ? float(alpha[j]) : 0.f;
}

// rows of Aop: [hi | lo | 0..] of the augmented candidate vector (see k_prescale_cand)
template <typename T1>
__global__ void __launch_bounds__(256)
k_prescale_cand_umma(bbo_gp_hyper h, const T1* __restrict__ Xq, int64_t qc, int64_t qc_pad, int dk, int ktp,
                     const double* __restrict__ hyp, uint32_t* __restrict__ Aop) {
  const int64_t c = int64_t(blockIdx.x) * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (c >= qc_pad) return;
  const int d = h.d;
  const bool matern = h.kind == BBO_KERNEL_MATERN52;
  const float eps = matern ? float(h.pairwise_eps) : 0.f;
  const float kLog2e = 1.4426950408889634f;
  float an = 0.f;
  for (int k = lane; k < d; k += 32) {
    const double x = (c < qc) ? double(Xq[c * d + k]) : 0.0;
    const float v = float(x * hyp[d + k]) - eps;
    an = fmaf(v, v, an);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) an += __shfl_xor_sync(0xffffffffu, an, o);
  const float R = matern ? an : -0.5f * kLog2e * an;      // the producers emit correlations (no sigma^2)
  uint32_t* row = Aop + c * ktp;
  for (int k = lane; k < dk; k += 32) {
    float u = 0.f;
    if (k < d) {
      const double x = (c < qc) ? double(Xq[c * d + k]) : 0.0;
      u = (matern ? -2.f : kLog2e) * (float(x * hyp[d + k]) - eps);
    } else if (k == d) {
      u = R;
    } else if (k == d + 1) {
      u = 1.f;
    }
    uint32_t hi, lo;
    split_tf32(u, hi, lo);
    row[k] = hi;
    row[dk + k] = lo;
  }
  for (int k = 2 * dk + lane; k < ktp; k += 32) row[k] = 0u;
}

template <bool MATERN, int DKS>     // DKS = dk / 8 k-steps per half row
__global__ void __launch_bounds__(ku::THREADS, 1)
k_kstar_umma(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
             const __grid_constant__ CUtensorMap mapO, int np, int tiles, int nbox, int dk, int d, int na, int ring,
             int stg2, const double* __restrict__ hyp, const float* __restrict__ alpha32, double* __restrict__ mu) {
  using namespace tf;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t a_full[2], a_empty[2], b_full[ku::MAXRING], b_empty[ku::MAXRING], t_full[ku::ACCS],
      t_empty[ku::ACCS];
  __shared__ uint32_t tmem_base_sh;
  __shared__ float mu_sh[2][3][ku::BM];
  __shared__ __align__(16) float al_sh[ku::EPI_WARPS][2][64];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));
  const int a_bytes = nbox * ku::BOX_BYTES;        // one candidate tile: 128 rows x nbox boxes
  const int ring_off = na * a_bytes;               // B boxes follow the A tiles
  const int stg_off = ring_off + ring * ku::BOX_BYTES;
  const int nblk = np / ku::BN;

  if (threadIdx.x == 0) {
    for (int a = 0; a < 2; ++a) {
      mbar_init(&a_full[a], 1);
      mbar_init(&a_empty[a], 1);
    }
    for (int s = 0; s < ku::MAXRING; ++s) {
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], 1);
    }
    for (int a = 0; a < ku::ACCS; ++a) {
      mbar_init(&t_full[a], 1);
      mbar_init(&t_empty[a], ku::EPI_WARPS * 32 / 2);   // one group of eight epilogue warps drains an accumulator
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_sh)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_sh;

  if (warp == 0) {
    if (lane == 0) {   // ===== TMA producer =====
      uint32_t a_phase0 = 0, a_phase1 = 0, phase = 0;
      int stage = 0, abuf = 0;
      auto load_a = [&](int tile, int buf) {
        uint32_t& ph = buf == 0 ? a_phase0 : a_phase1;
        mbar_wait(&a_empty[buf], ph ^ 1);   // every MMA that read this buffer has completed
        ph ^= 1;
        mbar_expect_tx(&a_full[buf], uint32_t(a_bytes));
        for (int b = 0; b < nbox; ++b)
          tma_load_2d(sgen + buf * a_bytes + b * ku::BOX_BYTES, &mapA, &a_full[buf], b * ku::BOXK, tile * ku::BM);
      };
      if (int(blockIdx.x) < tiles) load_a(blockIdx.x, 0);
      for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int next = tile + int(gridDim.x);
        for (int nb = 0; nb < nblk; ++nb) {
          // two A tiles: the next one is requested one block before this tile ends (its buffer was released when
          // the previous tile's MMAs completed, long ago)
          if (na == 2 && nb == nblk - 1 && next < tiles) load_a(next, abuf ^ 1);
          for (int b = 0; b < nbox; ++b) {
            mbar_wait(&b_empty[stage], phase ^ 1);
            mbar_expect_tx(&b_full[stage], uint32_t(ku::BOX_BYTES));
            tma_load_2d(sgen + ring_off + stage * ku::BOX_BYTES, &mapB, &b_full[stage], b * ku::BOXK, nb * ku::BN);
            if (++stage == ring) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
        if (na == 2) abuf ^= 1;
        else if (next < tiles) load_a(next, 0);       // one A tile: reloaded when this tile's MMAs are done
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {   // ===== MMA issuer =====
      // D=f32, A=B=tf32, both K-major, N=128, M=128.  Operand rows are [hi | lo] (2 dk words): a box of the B row
      // holds up to four 8-word k-steps; a `hi` step meets the candidate's hi AND lo step, a `lo` step its hi step
      // (hi hi + lo hi + hi lo: the 3xTF32 product), so every B box crosses L2 -> SM once per 128 x 128 block and the
      // candidate tile not at all.  This one thread feeds the tensor core of the SM: with run-time loop bounds it
      // spent ~300 instructions of descriptor arithmetic per 9-MMA block and the epilogue warps waited for it (ncu:
      // 29 % of all stall samples on their accumulator wait); DKS = dk / 8 is a template parameter so that every
      // descriptor is the tile's / the stage's base plus a compile-time constant.
      constexpr int DK = 8 * DKS, NBOX = (2 * DK + 31) / 32;
      const uint32_t idesc =
          (1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(ku::BN >> 3) << 17) | (uint32_t(ku::BM >> 4) << 24);
      uint32_t a_phase0 = 0, a_phase1 = 0, phase = 0, acc_phase = 0;
      int stage = 0, acc = 0, abuf = 0;
      for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        if (abuf == 0) {
          mbar_wait(&a_full[0], a_phase0);
          a_phase0 ^= 1;
        } else {
          mbar_wait(&a_full[1], a_phase1);
          a_phase1 ^= 1;
        }
        const uint64_t a0 = make_desc(sbase + uint32_t(abuf * a_bytes));
        for (int nb = 0; nb < nblk; ++nb) {
          mbar_wait(&t_empty[acc], acc_phase ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + uint32_t(acc * ku::BN);
#pragma unroll
          for (int b = 0; b < NBOX; ++b) {
            mbar_wait(&b_full[stage], phase);
            tc_fence_after();
            const uint64_t bbase = make_desc(sbase + uint32_t(ring_off + stage * ku::BOX_BYTES));
#pragma unroll
            for (int s4 = 0; s4 < 4; ++s4) {
              const int w = 32 * b + 8 * s4;          // compile-time after unrolling
              if (w < 2 * DK) {
                const uint64_t bdesc = bbase + uint64_t(s4 * 2);
                if (w < DK) {
                  tc_mma_tf32(d_tmem, a0 + ku::desc_off(w), bdesc, idesc, w != 0 ? 1u : 0u);
                  tc_mma_tf32(d_tmem, a0 + ku::desc_off(DK + w), bdesc, idesc, 1u);
                } else {
                  tc_mma_tf32(d_tmem, a0 + ku::desc_off(w - DK), bdesc, idesc, 1u);
                }
              }
            }
            tc_commit(&b_empty[stage]);
            if (++stage == ring) {
              stage = 0;
              phase ^= 1;
            }
          }
          tc_commit(&t_full[acc]);
          if (++acc == ku::ACCS) {
            acc = 0;
            acc_phase ^= 1;
          }
        }
        tc_commit(&a_empty[abuf]);
        if (na == 2) abuf ^= 1;
      }
    }
  } else {   // ===== epilogue (warps 2..17) =====
    // Two groups of eight warps take ALTERNATE 128 x 128 blocks (group g the blocks with blk % 2 == g, i.e. the
    // accumulators g and g + 2 of the ring); inside a block a warp owns one TMEM lane quarter and 64 columns (two
    // 32-column slabs).  While one group waits for its accumulator, loads it and hands its slabs to the TMA store, the
    // other is in its MUFU phase: with all sixteen warps on the same block the XU pipe idled through every load /
    // store phase (57 % busy).
    const int ew = warp - 2, quarter = warp & 3, grp = (ew >> 2) & 1, half = ew >> 3;
    const int row = quarter * 32 + lane;
    const int stg_bytes = stg2 ? ku::OUT_BYTES : ku::OUT_TILE;
    const uint32_t stg = sbase + uint32_t(stg_off + ew * stg_bytes);
    const float s2 = float(hyp[2 * d]);
    const float kLog2e = 1.4426950408889634f;
    const double cmean = hyp[2 * d + 1];
    const uint32_t tlane = tmem_base + (uint32_t(quarter * 32) << 16) + uint32_t(half * 64);
    const float* abase = alpha32 + half * 64;
    float m = 0.f;
    // alpha of this warp's 64 columns, one own block ahead (cp.async, per-warp double buffer)
    float* alw = &al_sh[ew][0][0];
    auto alpha_fetch = [&](int nb_, int buf_) {
      if (lane < 16) {
        const unsigned sa = smem_u32(alw + buf_ * 64 + lane * 4);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(abase + nb_ * ku::BN + lane * 4)
                     : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    const int my_tiles = (tiles - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x);
    const int nblocks = my_tiles * nblk;     // 128 x 128 blocks this CTA produces, accumulators in ring order
    if (grp < nblocks) alpha_fetch(grp % nblk, 0);          // block index grp is this group's first
    int mine = 0;                                           // own blocks done: parity selects the alpha buffer
    int tile = blockIdx.x, tcount = 0;
    for (int blk0 = 0; blk0 < nblocks; blk0 += nblk) {      // one candidate tile
      for (int nb = 0; nb < nblk; ++nb) {
        const int blk = blk0 + nb;
        if ((blk & 1) != grp) continue;
        const int acc = blk & 3;
        const uint32_t acc_phase = uint32_t(blk >> 2) & 1u;
        const float* al = alw + (mine & 1) * 64;
        {
          const int nxt = blk + 2;                           // this group's next block (harmless after the last)
          alpha_fetch(nxt < nblocks ? nxt % nblk : 0, (mine + 1) & 1);
        }
        mbar_wait(&t_full[acc], acc_phase);
        tc_fence_after();
#pragma unroll
        for (int sl = 0; sl < 2; ++sl) {
          uint32_t r[32];
          tmem_ld32_issue(tlane + uint32_t(acc * ku::BN + sl * 32), r);
          if (sl == 0) asm volatile("cp.async.wait_group 1;" ::: "memory");      // this block's alpha has landed
          if (lane == 0) {   // the staging tile about to be written is free again
            if (stg2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          }
          __syncwarp();
          tmem_ld_wait(r);
          if (sl == 1) {                                     // both slabs are in registers: the accumulator is free
            tc_fence_before();
            mbar_arrive(&t_empty[acc]);
          }
          // values -> correlations, mu partial, binary16 pairs into the 64-byte-swizzled staging tile, TMA store
          const uint32_t stile = stg + uint32_t(stg2 ? sl * ku::OUT_TILE : 0);
          const uint32_t sdst = stile + uint32_t(lane * 64);
#pragma unroll
          for (int c8 = 0; c8 < 4; ++c8) {
            uint32_t u[4];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const float4 av = *reinterpret_cast<const float4*>(al + sl * 32 + c8 * 8 + h * 4);
              const float a4[4] = {av.x, av.y, av.z, av.w};
              float cv[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float x = __uint_as_float(r[c8 * 8 + h * 4 + e]);
                if (!MATERN) {
                  cv[e] = ex2_approx(fminf(x, 0.f));
                } else {
                  const float rr = 2.2360679775f * sqrtf(fmaxf(x, 0.f));
                  cv[e] = fmaf(rr, fmaf(rr, 1.f / 3.f, 1.f), 1.f) * ex2_approx(-kLog2e * rr);
                }
                m = fmaf(cv[e], a4[e], m);
              }
              u[2 * h] = pack_half2(cv[0], cv[1]);
              u[2 * h + 1] = pack_half2(cv[2], cv[3]);
            }
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sdst + uint32_t((c8 ^ ((lane >> 1) & 3)) * 16)),
                         "r"(u[0]), "r"(u[1]), "r"(u[2]), "r"(u[3])
                         : "memory");
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();                          // staging tile complete
          if (lane == 0) {
            ku::tma_store_2d(&mapO, stile, nb * ku::BN + half * 64 + sl * 32, tile * ku::BM + quarter * 32);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
        ++mine;
      }
      // end of the tile: mu = c + s2 x four partial sums in a fixed order.  A partial is keyed by the PARITY of the
      // block index inside the tile (and the column half), not by the group: which group gets the even blocks of a
      // tile depends on the tile's position in this CTA's sequence when nblk is odd, and mu must not depend on that
      // (bit-identical results for every chunking and sharding).
      const int part = 2 * (grp ^ (blk0 & 1)) + half;
      if (part > 0) mu_sh[tcount & 1][part - 1][row] = m;
      asm volatile("bar.sync 1, %0;" ::"n"(ku::EPI_WARPS * 32) : "memory");
      if (part == 0)
        mu[int64_t(tile) * ku::BM + row] =
            cmean + double(s2) * double(((m + mu_sh[tcount & 1][0][row]) + mu_sh[tcount & 1][1][row]) +
                                        mu_sh[tcount & 1][2][row]);
      m = 0.f;
      tile += gridDim.x;
      ++tcount;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

template <int KS>
static size_t kstar_mma2_smem() {
  constexpr int DK = KS * 8, LS = 4 * ((DK / 4 + 1) | 1);
  return sizeof(uint32_t) * (size_t(4) * 64 * LS) + sizeof(float) * 128;
}
constexpr int kFastMaxD = 192;
static size_t kstar_fast_smem(int d) {
  const int LS = 4 * ((((d + 3) / 4) + 1) | 1);
  return sizeof(float) * (size_t(128) * LS + 2 * 64 * LS + 128 + 128 + 128);
}

// out (rows256, np): the panel's copy of linv (np, np); rows >= np are zero so that every 256-row TMA box of the
// B operand is fully in bounds.  Inside each 256-row block the rows are stored in REVERSE order (physical row c holds
// logical row 255 - c).  The panel only sums squares over the rows of a block, so their order is free, and this order
// puts the rows that are structurally zero in the diagonal block (logical row < column) at the END: for the
// 32-column stage starting kk columns into the diagonal block only the first 256 - kk physical rows are non-zero and
// the tensor core is given N = 256 - kk instead of 256 (k_vnorm_tf32_2sm: accumulator column c = physical row c in
// every stage).
// Values: binary16 copy of L^-1: out[prow, col] = half(2^e linv[lrow, col]) with the rows of every 256-row block
// in reverse order (see the panel's diagonal blocks), zero rows up to tf32_rows(n), and e chosen so that the largest
// |entry| lands in [2^13, 2^14): the 10-bit mantissa is TF32's, the scaling keeps every entry that matters in the
// normal range (an entry below 2^-28 of the largest is rounded with an absolute error of 2^-39 of the largest).
// The scale lives behind the data: float words [0] = 2^-2e (what the panel's sums of squares are multiplied by),
// [1] = 2^e, [2] = bits of max |entry| (scratch of the reduction).
__global__ void k_linv_absmax(int64_t total, const double* __restrict__ in, uint32_t* __restrict__ out_bits) {
  float m = 0.f;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x)
    m = fmaxf(m, fabsf(static_cast<float>(in[i])));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(out_bits, __float_as_uint(m));   // non-negative floats order as uints
}
__global__ void k_linv_scale(float* __restrict__ sc) {
  const float m = __uint_as_float(reinterpret_cast<const uint32_t*>(sc)[2]);
  int e = 0;
  if (m > 0.f && isfinite(m)) {
    int ex;
    frexpf(m, &ex);          // m = f 2^ex, f in [0.5, 1)
    e = 14 - ex;             // 2^e m in [2^13, 2^14)
    e = e > 50 ? 50 : (e < -50 ? -50 : e);   // keep 2^-2e a normal float (|L^-1| beyond 2^+-36 is not a GP)
  }
  sc[0] = exp2f(float(-2 * e));
  sc[1] = exp2f(float(e));
}
__global__ void k_make_f16(int64_t np, int64_t total, const double* __restrict__ in, const float* __restrict__ sc,
                           __half* __restrict__ out) {
  int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < total) {
    const int64_t prow = i / np, col = i - prow * np;
    const int64_t lrow = (prow >> 8) * 256 + 255 - (prow & 255);
    const double f = lrow < np ? in[lrow * np + col] * double(sc[1]) : 0.0;
    out[i] = __double2half(f);
  }
}

int64_t tf32_rows(int64_t n) { return ceil_div(padded(n), tf::BN) * tf::BN; }

// the scale words of the binary16 L^-1 copy (see k_make_f16): right behind the rows x np halves
static const float* linv_scale_ptr(const bbo_gp_state& s) {
  return s.linv_tf32 + tf32_rows(s.n) * padded(s.n) / 2;
}

int make_tf32(int64_t n, const double* linv, float* out, cudaStream_t st) {
  const int64_t np = padded(n), total = tf32_rows(n) * np;
  float* sc = out + total / 2;
  BBO_CUDA_CHECK(cudaMemsetAsync(sc, 0, 4 * sizeof(float), st));
  k_linv_absmax<<<unsigned(std::min<int64_t>(ceil_div(np * np, 256), 1184)), 256, 0, st>>>(
      np * np, linv, reinterpret_cast<uint32_t*>(sc) + 2);
  BBO_LAUNCH_CHECK();
  k_linv_scale<<<1, 1, 0, st>>>(sc);
  BBO_LAUNCH_CHECK();
  k_make_f16<<<unsigned(ceil_div(total, 256)), 256, 0, st>>>(np, total, linv, sc, reinterpret_cast<__half*>(out));
  BBO_LAUNCH_CHECK();
  return BBO_OK;
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
// TF32 scratch layout inside ScoreWorkspace::tf32:
//   Ks[2]  (qp, np) binary16 double-buffered K* chunk      mu1 (qp) float64  second mu buffer
//   Xs     (np, d)  float32  pre-scaled training inputs    alpha32 (np) float32
struct Tf32Scratch {
  __half* Ks[2];
  double* mu[3];     // K* producers write mu of chunk i into mu[i % 3] (selection of chunk i runs up to two chunks later)
  float* Xs;
  float* bnorm;
  float* alpha32;
  uint32_t* Xhi;
  uint32_t* Xlo;
  uint32_t* Bfrag;   // candidate vectors of the current chunk in B-fragment order (third K* form)
  uint32_t* Aop;     // (qp, 96) candidate operand rows of the current chunk (fourth K* form, tcgen05)
  uint32_t* Bop;     // (np, 96) observation operand rows
  double* vp32[2];   // (32, qp) |v|^2 partial sums, one slot per 256-row block (persistent panel), double-buffered
};
constexpr int64_t kSchedInts = 16384;   // work lists of the persistent panel (library-owned, see Tf32Aux)
static size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

size_t score_tf32_extra_bytes(int64_t n, int64_t d, int64_t qp) {
  const int64_t np = padded(n);
  return 2 * align256(size_t(qp) * np * sizeof(__half)) + 2 * align256(size_t(qp) * sizeof(double)) +
         align256(size_t(np) * d * sizeof(float)) + 2 * align256(size_t(np) * sizeof(float)) +
         2 * align256(size_t(np) * (d + 16) * sizeof(uint32_t)) +
         align256(size_t(qp) * 2 * (d + 16) * sizeof(uint32_t)) +
         align256(size_t(qp) * 32 * ku::MAXBOX_STREAM * sizeof(uint32_t)) +
         align256(size_t(np) * 32 * ku::MAXBOX_STREAM * sizeof(uint32_t)) + 2 * align256(size_t(32) * qp * sizeof(double)) + 256;
}

static Tf32Scratch carve(const ScoreWorkspace& w, int64_t d) {
  Tf32Scratch t;
  char* p = static_cast<char*>(w.tf32);
  const size_t kb = align256(size_t(w.qp) * w.np * sizeof(__half));
  t.Ks[0] = reinterpret_cast<__half*>(p);
  p += kb;
  t.Ks[1] = reinterpret_cast<__half*>(p);
  p += kb;
  t.mu[0] = w.mu;
  t.mu[1] = reinterpret_cast<double*>(p);
  p += align256(size_t(w.qp) * sizeof(double));
  t.mu[2] = reinterpret_cast<double*>(p);
  p += align256(size_t(w.qp) * sizeof(double));
  t.Xs = reinterpret_cast<float*>(p);
  p += align256(size_t(w.np) * d * sizeof(float));
  t.bnorm = reinterpret_cast<float*>(p);
  p += align256(size_t(w.np) * sizeof(float));
  t.alpha32 = reinterpret_cast<float*>(p);
  p += align256(size_t(w.np) * sizeof(float));
  t.Xhi = reinterpret_cast<uint32_t*>(p);
  p += align256(size_t(w.np) * (d + 16) * sizeof(uint32_t));
  t.Xlo = reinterpret_cast<uint32_t*>(p);
  p += align256(size_t(w.np) * (d + 16) * sizeof(uint32_t));
  t.Bfrag = reinterpret_cast<uint32_t*>(p);
  p += align256(size_t(w.qp) * 2 * (d + 16) * sizeof(uint32_t));
  t.Aop = reinterpret_cast<uint32_t*>(p);
  p += align256(size_t(w.qp) * 32 * ku::MAXBOX_STREAM * sizeof(uint32_t));
  t.Bop = reinterpret_cast<uint32_t*>(p);
  p += align256(size_t(w.np) * 32 * ku::MAXBOX_STREAM * sizeof(uint32_t));
  for (int i = 0; i < 2; ++i) {
    t.vp32[i] = reinterpret_cast<double*>(p);
    p += align256(size_t(32) * w.qp * sizeof(double));
  }
  return t;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

// 2-D float32 row-major (rows, cols) tensor, box = (box_rows, 32 floats), 128-byte swizzle
// row-major (rows, cols) tensor of 4-byte words or binary16 values; boxes of box_rows x box_k elements whose rows
// are 128 bytes (SWIZZLE_128B) or 64 bytes (SWIZZLE_64B)
static int make_map_bytes(CUtensorMap* m, const void* base, CUtensorMapDataType dt, int esize, int64_t rows,
                          int64_t cols, int box_rows, int box_k) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return cuda_fail(cudaErrorNotSupported, "cudaGetDriverEntryPoint(cuTensorMapEncodeTiled)");
  if (box_k * esize != 128 && box_k * esize != 64) return BBO_ERR_BAD_ARG;
  cuuint64_t gdim[2] = {cuuint64_t(cols), cuuint64_t(rows)};
  cuuint64_t gstr[1] = {cuuint64_t(cols) * cuuint64_t(esize)};
  cuuint32_t box[2] = {cuuint32_t(box_k), cuuint32_t(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, dt, 2, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   box_k * esize == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return cuda_fail(cudaErrorInvalidValue, "cuTensorMapEncodeTiled");
  return BBO_OK;
}
static int make_map(CUtensorMap* m, const float* base, int64_t rows, int64_t cols, int box_rows, int box_k = tf::BK) {
  return make_map_bytes(m, base, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, rows, cols, box_rows, box_k);
}
static int make_map(CUtensorMap* m, const __half* base, int64_t rows, int64_t cols, int box_rows, int box_k) {
  return make_map_bytes(m, base, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, rows, cols, box_rows, box_k);
}

// auxiliary stream + events of the K* / panel software pipeline (created once per device)
constexpr int kSchedSlots = 4;
struct Tf32Aux {
  cudaStream_t sk = nullptr;   // K* producer
  cudaStream_t ss = nullptr;   // acquisition + selection
  cudaEvent_t entry = nullptr, kdone[2] = {nullptr, nullptr}, pdone[2] = {nullptr, nullptr},
              sdone[2] = {nullptr, nullptr};
  // work lists of the persistent panel: kSchedSlots cached shapes (a pool has full chunks and one last chunk),
  // 64 KB of library-owned device memory each, rebuilt by k_panel_schedule only when the shape changes.  The
  // cache is shared by every stream of the device, so each slot carries two events: `built` (recorded behind
  // k_panel_schedule; a stream that hits the cache waits for it before its panel reads the list) and `used`
  // (recorded behind the last panel that read the slot; a rebuild waits for it before overwriting the list).
  int* sched[kSchedSlots] = {};
  int64_t sched_key[kSchedSlots] = {};
  cudaEvent_t built[kSchedSlots] = {}, used[kSchedSlots] = {};
  uint64_t stamp[kSchedSlots] = {};
  uint64_t clock = 0;
};
static Tf32Aux g_aux[64];
static std::mutex g_aux_mutex;
static int get_aux(Tf32Aux** out) {
  int dev = 0;
  BBO_CUDA_CHECK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return BBO_ERR_BAD_ARG;
  std::lock_guard<std::mutex> lock(g_aux_mutex);
  Tf32Aux& a = g_aux[dev];
  if (!a.sk) {
    BBO_CUDA_CHECK(cudaStreamCreateWithFlags(&a.sk, cudaStreamNonBlocking));
    BBO_CUDA_CHECK(cudaStreamCreateWithFlags(&a.ss, cudaStreamNonBlocking));
    BBO_CUDA_CHECK(cudaEventCreateWithFlags(&a.entry, cudaEventDisableTiming));
    for (int i = 0; i < 2; ++i) {
      BBO_CUDA_CHECK(cudaEventCreateWithFlags(&a.kdone[i], cudaEventDisableTiming));
      BBO_CUDA_CHECK(cudaEventCreateWithFlags(&a.pdone[i], cudaEventDisableTiming));
      BBO_CUDA_CHECK(cudaEventCreateWithFlags(&a.sdone[i], cudaEventDisableTiming));
    }
    for (int i = 0; i < kSchedSlots; ++i) {
      BBO_CUDA_CHECK(cudaEventCreateWithFlags(&a.built[i], cudaEventDisableTiming));
      BBO_CUDA_CHECK(cudaEventCreateWithFlags(&a.used[i], cudaEventDisableTiming));
      BBO_CUDA_CHECK(cudaMalloc(&a.sched[i], size_t(kSchedInts) * sizeof(int)));
    }
  }
  *out = &a;
  return BBO_OK;
}

// tcgen05 K* producer: default for d <= 30.  BBO_KSTAR_NO_UMMA=1 sends those shapes through k_kstar_mma2 (the
// default for 30 < d <= 126) -- used by the test that checks both producers give the same selection.
static bool use_kstar_umma(int d) {
  static const bool off = getenv("BBO_KSTAR_NO_UMMA") != nullptr;
  // 30 < d <= 126: the streaming tcgen05 form (round 2); BBO_KSTAR_NO_UMMA_STREAM=1 keeps k_kstar_mma2 there
  static const bool no_stream = getenv("BBO_KSTAR_NO_UMMA_STREAM") != nullptr;
  return !off && (d <= ku::kMaxD || (!no_stream && d <= ku::kMaxDStream));
}
static int sm_count() {
  static int sms[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (sms[dev] == 0 && cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
    sms[dev] = 148;
  return sms[dev];
}

// `side_sms` > 0: the producer runs BESIDE a panel that leaves that many SMs free (see score_tf32_pipeline): its grid
// is capped there and launched as clusters of two CTAs so that it occupies whole TPCs (the panel's CTA pairs need both
// SMs of a TPC).
static int launch_kstar(const bbo_gp_state& s, const ScoreWorkspace& w, const Tf32Scratch& t, int buf,
                        const void* Xq, int xq_is_f32, int64_t qc, int64_t qc_pad, double* mu_chunk, cudaStream_t st,
                        int side_sms = 0) {
  const int64_t np = w.np;
  prof_begin(kProfKstar, st);
  if (use_kstar_umma(s.h.d)) {
    const int dk = 8 * ((s.h.d + 2 + 7) / 8), ktp = 32 * ((2 * dk + 31) / 32), nbox = ktp / 32;
    const bool matern = s.h.kind == BBO_KERNEL_MATERN52;
    if (xq_is_f32)
      k_prescale_cand_umma<float><<<unsigned(qc_pad / 8), 256, 0, st>>>(s.h, static_cast<const float*>(Xq), qc, qc_pad,
                                                                       dk, ktp, s.hyp, t.Aop);
    else
      k_prescale_cand_umma<double><<<unsigned(qc_pad / 8), 256, 0, st>>>(s.h, static_cast<const double*>(Xq), qc,
                                                                        qc_pad, dk, ktp, s.hyp, t.Aop);
    BBO_LAUNCH_CHECK();
    CUtensorMap mapA, mapB, mapO;
    int rc = make_map(&mapA, reinterpret_cast<const float*>(t.Aop), qc_pad, ktp, ku::BM, ku::BOXK);
    if (rc != BBO_OK) return rc;
    rc = make_map(&mapB, reinterpret_cast<const float*>(t.Bop), np, ktp, ku::BN, ku::BOXK);
    if (rc != BBO_OK) return rc;
    rc = make_map(&mapO, t.Ks[buf], qc_pad, np, 32, 32);        // 32 candidates x 32 binary16 values (64-byte rows)
    if (rc != BBO_OK) return rc;
    const int tiles = int(qc_pad / ku::BM);
    int grid = tiles < sm_count() ? tiles : sm_count();
    const bool side = side_sms >= 2 && grid > side_sms;
    if (side) grid = side_sms & ~1;
    const double* hyp = s.hyp;
    ku::Plan pl = ku::plan(nbox);
    if (const char* e = getenv("BBO_KSTAR_NA")) pl.na = atoi(e);         // tuning aids (the plan stays within the budget
    if (const char* e = getenv("BBO_KSTAR_RING")) pl.ring = atoi(e);     // only for values not above the planned ones)
    if (const char* e = getenv("BBO_KSTAR_STG2")) pl.stg2 = atoi(e);
    if (pl.ring < 2 || nbox > ku::MAXBOX) return BBO_ERR_BAD_SHAPE;
    size_t most = 0;
    for (int b = 1; b <= ku::MAXBOX; ++b) most = std::max(most, ku::plan(b).bytes);
#define BBO_KSTAR_UMMA(DKS_)                                                                                        \
  case DKS_: {                                                                                                      \
    static DeviceOnce once;                                                                                         \
    if (once.first()) {                                                                                             \
      BBO_CUDA_CHECK(cudaFuncSetAttribute(k_kstar_umma<false, DKS_>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                          int(most)));                                                              \
      BBO_CUDA_CHECK(cudaFuncSetAttribute(k_kstar_umma<true, DKS_>, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                                          int(most)));                                                              \
    }                                                                                                               \
    cudaLaunchConfig_t kcfg = {};                                                                                   \
    kcfg.gridDim = dim3(unsigned(grid));                                                                            \
    kcfg.blockDim = dim3(ku::THREADS);                                                                              \
    kcfg.dynamicSmemBytes = pl.bytes;                                                                               \
    kcfg.stream = st;                                                                                               \
    cudaLaunchAttribute kat[1];                                                                                     \
    kat[0].id = cudaLaunchAttributeClusterDimension;                                                                \
    kat[0].val.clusterDim.x = side ? 2 : 1;                                                                         \
    kat[0].val.clusterDim.y = 1;                                                                                    \
    kat[0].val.clusterDim.z = 1;                                                                                    \
    kcfg.attrs = kat;                                                                                               \
    kcfg.numAttrs = 1;                                                                                              \
    const int np_k = int(np), d_k = s.h.d;                                                                          \
    const float* al_k = t.alpha32;                                                                                  \
    if (matern)                                                                                                     \
      BBO_CUDA_CHECK(cudaLaunchKernelEx(&kcfg, k_kstar_umma<true, DKS_>, mapA, mapB, mapO, np_k, tiles, nbox, dk,   \
                                        d_k, pl.na, pl.ring, pl.stg2, hyp, al_k, mu_chunk));                        \
    else                                                                                                            \
      BBO_CUDA_CHECK(cudaLaunchKernelEx(&kcfg, k_kstar_umma<false, DKS_>, mapA, mapB, mapO, np_k, tiles, nbox, dk,  \
                                        d_k, pl.na, pl.ring, pl.stg2, hyp, al_k, mu_chunk));                        \
  } break;
    switch (dk / 8) {
      BBO_KSTAR_UMMA(1)
      BBO_KSTAR_UMMA(2)
      BBO_KSTAR_UMMA(3)
      BBO_KSTAR_UMMA(4)
      BBO_KSTAR_UMMA(5)
      BBO_KSTAR_UMMA(6)
      BBO_KSTAR_UMMA(7)
      BBO_KSTAR_UMMA(8)
      BBO_KSTAR_UMMA(9)
      BBO_KSTAR_UMMA(10)
      BBO_KSTAR_UMMA(11)
      BBO_KSTAR_UMMA(12)
      BBO_KSTAR_UMMA(13)
      BBO_KSTAR_UMMA(14)
      BBO_KSTAR_UMMA(15)
      BBO_KSTAR_UMMA(16)
      default:
        return BBO_ERR_BAD_ARG;
    }
#undef BBO_KSTAR_UMMA
  } else if (s.h.d <= kMma2MaxD) {
    const int ks = (s.h.d + 2 + 7) / 8;
    const unsigned grid = unsigned(qc_pad / 128);
    const bool matern = s.h.kind == BBO_KERNEL_MATERN52;
#define BBO_KSTAR_MMA2_L(T_, KS_, M_)                                                                          \
  k_kstar_mma2<T_, KS_, M_><<<grid, 256, sm, st>>>(s.h, static_cast<const T_*>(Xq), qc, t.Xhi, t.Xlo, np, s.hyp, \
                                                   t.alpha32, t.Ks[buf], mu_chunk)
#define BBO_KSTAR_MMA2(KS_)                                                                                    \
  case KS_: {                                                                                                  \
    const size_t sm = kstar_mma2_smem<KS_>();                                                                  \
    static DeviceOnce attr_once;                                                                               \
    if (sm > 48 * 1024 && attr_once.first()) {                                                                 \
      BBO_CUDA_CHECK(cudaFuncSetAttribute(k_kstar_mma2<float, KS_, false>,                                     \
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, int(sm)));              \
      BBO_CUDA_CHECK(cudaFuncSetAttribute(k_kstar_mma2<float, KS_, true>,                                      \
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, int(sm)));              \
      BBO_CUDA_CHECK(cudaFuncSetAttribute(k_kstar_mma2<double, KS_, false>,                                    \
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, int(sm)));              \
      BBO_CUDA_CHECK(cudaFuncSetAttribute(k_kstar_mma2<double, KS_, true>,                                     \
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, int(sm)));              \
    }                                                                                                          \
    if (xq_is_f32) {                                                                                           \
      if (matern) BBO_KSTAR_MMA2_L(float, KS_, true);                                                          \
      else BBO_KSTAR_MMA2_L(float, KS_, false);                                                                \
    } else {                                                                                                   \
      if (matern) BBO_KSTAR_MMA2_L(double, KS_, true);                                                         \
      else BBO_KSTAR_MMA2_L(double, KS_, false);                                                               \
    }                                                                                                          \
  } break;
    switch (ks) {
      BBO_KSTAR_MMA2(1)
      BBO_KSTAR_MMA2(2)
      BBO_KSTAR_MMA2(3)
      BBO_KSTAR_MMA2(4)
      BBO_KSTAR_MMA2(5)
      BBO_KSTAR_MMA2(6)
      BBO_KSTAR_MMA2(7)
      BBO_KSTAR_MMA2(8)
      BBO_KSTAR_MMA2(9)
      BBO_KSTAR_MMA2(10)
      BBO_KSTAR_MMA2(11)
      BBO_KSTAR_MMA2(12)
      BBO_KSTAR_MMA2(13)
      BBO_KSTAR_MMA2(14)
      BBO_KSTAR_MMA2(15)
      BBO_KSTAR_MMA2(16)
      default:
        return BBO_ERR_BAD_ARG;
    }
#undef BBO_KSTAR_MMA2
#undef BBO_KSTAR_MMA2_L
  } else if (s.h.d <= kFastMaxD) {
    static DeviceOnce fast_attr_once;
    const size_t ksm = kstar_fast_smem(s.h.d);
    if (fast_attr_once.first()) {
      BBO_CUDA_CHECK(cudaFuncSetAttribute(k_kstar_f32_fast<float>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          int(kstar_fast_smem(kFastMaxD))));
      BBO_CUDA_CHECK(cudaFuncSetAttribute(k_kstar_f32_fast<double>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          int(kstar_fast_smem(kFastMaxD))));
    }
    if (xq_is_f32)
      k_kstar_f32_fast<float><<<unsigned(qc_pad / 128), 256, ksm, st>>>(s.h, static_cast<const float*>(Xq), qc, t.Xs,
                                                                     t.bnorm, s.n, np, s.hyp, t.alpha32, t.Ks[buf],
                                                                     mu_chunk);
    else
      k_kstar_f32_fast<double><<<unsigned(qc_pad / 128), 256, ksm, st>>>(s.h, static_cast<const double*>(Xq), qc, t.Xs,
                                                                      t.bnorm, s.n, np, s.hyp, t.alpha32, t.Ks[buf],
                                                                      mu_chunk);
  } else {
    if (xq_is_f32)
      k_kstar_f32<float><<<unsigned(qc_pad / 64), 256, 0, st>>>(s.h, static_cast<const float*>(Xq), qc, s.X, s.n, np,
                                                               s.hyp, s.alpha, t.Ks[buf], mu_chunk);
    else
      k_kstar_f32<double><<<unsigned(qc_pad / 64), 256, 0, st>>>(s.h, static_cast<const double*>(Xq), qc, s.X, s.n,
                                                                np, s.hyp, s.alpha, t.Ks[buf], mu_chunk);
  }
  BBO_LAUNCH_CHECK();
  prof_end(kProfKstar, st);
  return BBO_OK;
}

// Does a chunk of qc_pad candidates run on the persistent form of the panel (launch_panel's own conditions)?
static bool panel_persists(int64_t np, int64_t qc_pad, int reserve_sms) {
  static const bool no_persist = getenv("BBO_TF32_NO_PERSIST") != nullptr;
  static const int64_t l2_mb = getenv("BBO_TF32_PERSIST_MAX_MB") ? atoll(getenv("BBO_TF32_PERSIST_MAX_MB")) : 80;
  const int64_t tiles = qc_pad / tf::BM, nblk = ceil_div(np, tf::BN);
  if (no_persist || tiles % 2 != 0) return false;
  const int ntp = int(tiles / 2), nrp = int((nblk + 1) / 2), nitems = ntp * nrp;
  const int pairs = (sm_count() - reserve_sms) / 2;
  const int nc = nitems < pairs ? nitems : pairs;
  const int maxj = (ntp * nrp + nc - 1) / nc * 4 + 8;
  return np * np * int64_t(sizeof(__half)) <= (l2_mb << 20) && nblk >= 2 && nblk <= 32 && nc <= 96 &&
         int64_t(nc) * maxj <= kSchedInts;
}

// *vpart_out / *nsplit_out: where the partial sums of this chunk are and how k_acq / k_acq_topk add them (4 = the
// four fixed slots of `vpart`; -nblk = one slot per row block in the persistent panel's own buffer).
// qc_pad is a multiple of 256 (ScoreWorkspace pads chunks so): every chunk runs on CTA pairs.
// `reserve_sms`: SMs the persistent form leaves to the K* producer of the next chunk (0 = none).
static int launch_panel(const bbo_gp_state& s, const ScoreWorkspace& w, const Tf32Scratch& t, int buf,
                        int64_t qc_pad, double* vpart, int* nsplit_out, double** vpart_out, cudaStream_t st,
                        int reserve_sms = 0) {
  const int64_t np = w.np;
  if (qc_pad % (2 * tf::BM) != 0) return BBO_ERR_BAD_SHAPE;
  constexpr bool two_sm = true;
  static DeviceOnce attr2_once;
  if (attr2_once.first()) {
    BBO_CUDA_CHECK(cudaFuncSetAttribute(k_vnorm_tf32_2sm<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        tf2::SMEM_BYTES));
    BBO_CUDA_CHECK(cudaFuncSetAttribute(k_vnorm_tf32_2sm<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        tf2::SMEM_BYTES));
  }
  CUtensorMap mapA, mapB;
  int rc = make_map(&mapA, t.Ks[buf], qc_pad, np, tf::BM, tfp::BK);
  if (rc != BBO_OK) return rc;
  rc = make_map(&mapB, reinterpret_cast<const __half*>(s.linv_tf32), tf32_rows(s.n), np, tf::BN / 2, tfp::BK);
  if (rc != BBO_OK) return rc;
  const int64_t tiles = qc_pad / tf::BM;
  const int64_t nblk = ceil_div(np, tf::BN);
  // the 256-row blocks of a candidate tile are dealt round-robin to 1, 2 or 4 CTAs, chosen by
  // occupancy; partial sums always go to four fixed slots (block index mod 4), so the result is
  // bit-identical for every split, chunking and sharding of the pool.
  int64_t nsplit = tiles >= 148 ? 1 : (2 * tiles >= 148 ? 2 : 4);
  if (nblk < 2) nsplit = 1;
  *nsplit_out = 4;   // k_acq always sums the four slots
  *vpart_out = vpart;
  const double* hyp = s.hyp;
  const int d_i = s.h.d;
  const float* lscale = linv_scale_ptr(s);
  prof_begin(kProfTf32, st);
  {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(unsigned(tiles), unsigned(nsplit));
    cfg.blockDim = dim3(tf::THREADS);
    cfg.dynamicSmemBytes = tf2::SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    const int np_i = int(np);
    const int64_t qp_i = w.qp;
    double* vp = vpart;
    static const int shrink = getenv("BBO_TF32_FULL_DIAG") == nullptr ? 1 : 0;
    static const bool no_persist = getenv("BBO_TF32_NO_PERSIST") != nullptr;
    const int ntp = int(tiles / 2), nrp = int((nblk + 1) / 2);
    // CTA pairs: one per SM pair, or one per item when the chunk has fewer items than that (a small pool spreads the
    // row pairs of its few tile pairs over the chip)
    const int nitems = ntp * nrp;
    const int pairs = (sm_count() - reserve_sms) / 2;
    const int nc = nitems < pairs ? nitems : pairs;
    // list capacity: four times the mean item count (the lists hold ntp * nrp items in total, so with full lists
    // skipped by the scheduler every item finds a place)
    const int maxj = (ntp * nrp + nc - 1) / nc * 4 + 8;
    // The persistent form trades the lock-step walk over the row pairs (which keeps the L^-1 blocks in flight few)
    // for K* strip reuse: it needs the whole TF32 L^-1 resident in L2 next to the strips (n = 4096: 67 of 126 MB).
    // At n = 8192 (268 MB) it was 9 % slower than the lock-step form, which therefore keeps the large problems.
    static const int64_t l2_mb = getenv("BBO_TF32_PERSIST_MAX_MB") ? atoll(getenv("BBO_TF32_PERSIST_MAX_MB")) : 80;
    const bool linv_fits_l2 = np * np * int64_t(sizeof(__half)) <= (l2_mb << 20);
    if (two_sm && !no_persist && linv_fits_l2 && nblk >= 2 && nblk <= 32 && nc <= 96 &&
        int64_t(nc) * maxj <= kSchedInts) {
      // persistent form: one CTA pair per SM pair walks a work list (k_panel_schedule); one slot per row block
      static const int64_t group_mb = getenv("BBO_TF32_PANEL_GROUP_MB") ? atoll(getenv("BBO_TF32_PANEL_GROUP_MB")) : 4;
      const int gs_raw = int((group_mb << 20) / (int64_t(256) * np * 4));
      const int gs = gs_raw < 1 ? 1 : gs_raw;
      Tf32Aux* ax = nullptr;
      rc = get_aux(&ax);
      if (rc != BBO_OK) return rc;
      const int64_t key = (int64_t(ntp) << 40) | (int64_t(nblk) << 24) | (int64_t(nc) << 12) | int64_t(gs);
      int slot = -1;
      {
        std::lock_guard<std::mutex> lock(g_aux_mutex);
        for (int i = 0; i < kSchedSlots; ++i)
          if (ax->sched_key[i] == key) slot = i;
        if (slot >= 0) {
          // built on some stream earlier: this stream's panel must not read the list before that kernel ran
          BBO_CUDA_CHECK(cudaStreamWaitEvent(st, ax->built[slot], 0));
        } else {
          slot = 0;   // least recently used slot
          for (int i = 1; i < kSchedSlots; ++i)
            if (ax->stamp[i] < ax->stamp[slot]) slot = i;
          if (ax->sched_key[slot] != 0) BBO_CUDA_CHECK(cudaStreamWaitEvent(st, ax->used[slot], 0));
          k_panel_schedule<<<1, 32, 0, st>>>(ntp, nrp, nc, gs, maxj, ax->sched[slot]);
          BBO_LAUNCH_CHECK();
          BBO_CUDA_CHECK(cudaEventRecord(ax->built[slot], st));
          ax->sched_key[slot] = key;
        }
        ax->stamp[slot] = ++ax->clock;
      }
      cfg.gridDim = dim3(unsigned(2 * nc), 1);
      const int* sc = ax->sched[slot];
      double* vp32 = t.vp32[buf];
      BBO_CUDA_CHECK(cudaLaunchKernelEx(&cfg, k_vnorm_tf32_2sm<true>, mapA, mapB, np_i, qp_i, vp32, shrink, sc, maxj,
                                        hyp, d_i, lscale));
      {
        std::lock_guard<std::mutex> lock(g_aux_mutex);
        BBO_CUDA_CHECK(cudaEventRecord(ax->used[slot], st));
      }
      *nsplit_out = -int(nblk);
      *vpart_out = vp32;
    } else
      BBO_CUDA_CHECK(cudaLaunchKernelEx(&cfg, k_vnorm_tf32_2sm<false>, mapA, mapB, np_i, qp_i, vp, shrink,
                                        static_cast<const int*>(nullptr), 0, hyp, d_i, lscale));
  }
  BBO_LAUNCH_CHECK();
  prof_end(kProfTf32, st);
  return BBO_OK;
}

// Software pipeline over the chunks of one pool.  Three streams when the tcgen05 K* producer is in use:
//   sk   K*(i+1)                      queued behind K*(i); its CTAs fill the tail of panel(i)
//   st   panel(i)                     back to back: nothing else sits between two panels
//   ss   acquisition + top-k of (i)   small kernels that share the SMs with panel(i+1)
// Buffers are double: Ks / mu (written by K*), the |v|^2 partial sums (written by the panel; slots 0-3 and 4-7 of
// w.vpart).  K*(i+2) waits for panel(i) (Ks) and selection(i) (mu); panel(i+2) follows K*(i+2), hence also
// selection(i) (partial sums).  `consume(c0, qc, mu, nsplit, vpart, stream)` enqueues the acquisition / top-k work.
int score_tf32_pipeline(const bbo_gp_state& s, const ScoreWorkspace& w, const void* Xq, int xq_is_f32, int64_t q,
                        const ConsumeFn& consume, cudaStream_t st, bool short_first) {
  Tf32Aux* aux = nullptr;
  int rc = get_aux(&aux);
  if (rc != BBO_OK) return rc;
  // Two streams: K* of chunk i+1 is queued behind K* of chunk i on the auxiliary stream while the panel of chunk i
  // runs on `st`.  Neither the tcgen05 K* producer (214 KB shared memory, 512 TMEM columns) nor the panel fits on an
  // SM next to the other, so the K* CTAs only take the SMs the panel's last wave has left: they fill its tail
  // (measured 35.4 -> 35.0 ms/step, twice, same box).  With the mma.sync producers, which do share SMs with the
  // panel, the overlap gained nothing (40.7 vs 40.6 ms/step: both saturate shared-memory bandwidth), so they stay
  // on one stream.  BBO_TF32_OVERLAP=1 / BBO_TF32_NO_OVERLAP=1 force either schedule.
  static const bool force_on = getenv("BBO_TF32_OVERLAP") != nullptr;
  static const bool force_off = getenv("BBO_TF32_NO_OVERLAP") != nullptr;
  // (the per-kernel CUDA-event regions of bbo_profile_* are only meaningful on one stream: profiling forces it)
  const bool no_overlap = force_off || prof_enabled() || (!force_on && !use_kstar_umma(s.h.d));
  Tf32Aux serial;
  if (no_overlap) {
    serial = *aux;
    serial.sk = st;
    aux = &serial;
  }
  const Tf32Scratch t = carve(w, s.h.d);
  const size_t esz = xq_is_f32 ? sizeof(float) : sizeof(double);
  const int d = s.h.d;
  if (use_kstar_umma(d)) {
    const int dk = 8 * ((d + 2 + 7) / 8), ktp = 32 * ((2 * dk + 31) / 32);
    k_prescale_aug_umma<<<unsigned(ceil_div(w.np, 8)), 256, 0, st>>>(s.X, s.n, w.np, d, dk, ktp,
                                                                    s.h.kind == BBO_KERNEL_MATERN52 ? 1 : 0, s.hyp,
                                                                    s.alpha, t.Bop, t.alpha32);
    BBO_LAUNCH_CHECK();
  } else if (d <= kMma2MaxD) {
    const int dk = 8 * ((d + 2 + 7) / 8);
    k_prescale_aug<<<unsigned(ceil_div(w.np, 8)), 256, 0, st>>>(s.X, s.n, w.np, d, dk,
                                                               s.h.kind == BBO_KERNEL_MATERN52 ? 1 : 0, s.hyp, s.alpha,
                                                               t.Xhi, t.Xlo, t.alpha32);
    BBO_LAUNCH_CHECK();
  } else if (d <= kFastMaxD) {
    k_prescale<<<unsigned(ceil_div(w.np, 8)), 256, 0, st>>>(s.X, s.n, w.np, d, s.hyp, s.alpha, t.Xs, t.bnorm,
                                                           t.alpha32);
    BBO_LAUNCH_CHECK();
  }
  // one stream (default): stream order already is the dependency chain, no events needed
  const bool two_streams = aux->sk != st;
  cudaStream_t ssel = two_streams ? aux->ss : st;
  if (two_streams) {
    BBO_CUDA_CHECK(cudaEventRecord(aux->entry, st));
    BBO_CUDA_CHECK(cudaStreamWaitEvent(aux->sk, aux->entry, 0));
    BBO_CUDA_CHECK(cudaStreamWaitEvent(aux->ss, aux->entry, 0));
  }
  // `short_first`: the selection starts from an empty list, so EVERY candidate of the first chunk passes its
  // threshold filter and the chunk is reduced by arg-max rounds (0.5 ms for a 740 k-candidate chunk at n = 256, on
  // the critical path of a 1.3 ms pool).  A first chunk of one wave (one 128-candidate tile per SM) establishes the
  // threshold cheaply; from the second chunk on only a fraction of a per cent pass.
  const int64_t wave = int64_t(sm_count()) * 128;
  const int64_t q0 = (short_first && w.qp >= 2 * wave && q > 2 * wave && (wave / 128) % 2 == 0) ? wave : w.qp;
  const int64_t nchunk = q <= q0 ? 1 : 1 + ceil_div(q - q0, w.qp);
  // BBO_TF32_TIMELINE=1 (debug aid): start / end of every K*, panel and selection of the call, in us after the first
  // launch, on stderr (synchronises at the end of the call)
  static const bool timeline = getenv("BBO_TF32_TIMELINE") != nullptr;
  std::vector<cudaEvent_t> tl;
  auto mark = [&](cudaStream_t sx) {
    if (!timeline) return;
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    cudaEventRecord(e, sx);
    tl.push_back(e);
  };
  auto chunk_start = [&](int64_t i) { return i == 0 ? int64_t(0) : q0 + (i - 1) * w.qp; };
  auto chunk_len = [&](int64_t i) {
    const int64_t full = i == 0 ? q0 : w.qp, left = q - chunk_start(i);
    return left < full ? left : full;
  };
  auto chunk_ptr = [&](int64_t i) { return static_cast<const char*>(Xq) + size_t(chunk_start(i)) * d * esz; };
  // Side-by-side mode (BBO_TF32_KSTAR_SMS = SMs for the producer): while panel(i) runs on a persistent grid that
  // leaves those SMs free, K*(i+1) runs on them, instead of on the whole chip in the gap between two panels.
  static const int kstar_sms = getenv("BBO_TF32_KSTAR_SMS") ? atoi(getenv("BBO_TF32_KSTAR_SMS")) : 0;
  auto side_for = [&](int64_t i) -> int {   // SMs reserved while panel(i) and K*(i+1) run
    if (kstar_sms < 2 || !two_streams || !use_kstar_umma(d)) return 0;
    const int64_t qc = chunk_len(i);
    return panel_persists(w.np, ceil_div(qc, 256) * 256, kstar_sms) ? kstar_sms : 0;
  };
  {
    const int64_t qc = chunk_len(0);
    mark(aux->sk);
    rc = launch_kstar(s, w, t, 0, chunk_ptr(0), xq_is_f32, qc, ceil_div(qc, 256) * 256, t.mu[0], aux->sk);
    if (rc != BBO_OK) return rc;
    mark(aux->sk);
    if (two_streams) BBO_CUDA_CHECK(cudaEventRecord(aux->kdone[0], aux->sk));
  }
  for (int64_t i = 0; i < nchunk; ++i) {
    const int buf = int(i & 1);
    if (i + 1 < nchunk) {
      const int nb = buf ^ 1;
      // chunk i-1 released K* buffer nb when its panel finished.  mu is triple-buffered, so K*(i+1) never waits for
      // the selection of chunk i-1 (small kernels that start when panel(i-1) ends, as this K* would like to);
      // mu[(i+1) % 3] was last read by the selection of chunk i-2, finished a whole panel ago (sdone[buf] still holds
      // that record here: chunk i's own is made further down).  Steady state (BBO_TF32_TIMELINE=1): panel(i) and
      // K*(i+1) become ready together, the panel takes the SMs, K*(i+1) runs when it ends (~95 us), panel(i+1)
      // follows 6 us later: a chunk costs panel + K*, selection runs beside the K*.
      if (two_streams && i >= 1) BBO_CUDA_CHECK(cudaStreamWaitEvent(aux->sk, aux->pdone[nb], 0));
      if (two_streams && i >= 2) BBO_CUDA_CHECK(cudaStreamWaitEvent(aux->sk, aux->sdone[buf], 0));
      const int64_t qn = chunk_len(i + 1);
      mark(aux->sk);
      rc = launch_kstar(s, w, t, nb, chunk_ptr(i + 1), xq_is_f32, qn, ceil_div(qn, 256) * 256, t.mu[(i + 1) % 3],
                        aux->sk, side_for(i));
      if (rc != BBO_OK) return rc;
      mark(aux->sk);
      if (two_streams) BBO_CUDA_CHECK(cudaEventRecord(aux->kdone[nb], aux->sk));
    }
    const int64_t qc = chunk_len(i);
    if (two_streams) {
      BBO_CUDA_CHECK(cudaStreamWaitEvent(st, aux->kdone[buf], 0));
      // the partial-sum buffer `buf` was read by the selection of chunk i-2 (long finished)
      if (i >= 2) BBO_CUDA_CHECK(cudaStreamWaitEvent(st, aux->sdone[buf], 0));
    }
    int nsplit = 1;
    double* vpart = nullptr;
    mark(st);
    rc = launch_panel(s, w, t, buf, ceil_div(qc, 256) * 256, w.vpart + int64_t(buf) * 4 * w.qp, &nsplit, &vpart, st,
                      i + 1 < nchunk ? side_for(i) : 0);
    if (rc != BBO_OK) return rc;
    mark(st);
    if (two_streams) {
      BBO_CUDA_CHECK(cudaEventRecord(aux->pdone[buf], st));
      BBO_CUDA_CHECK(cudaStreamWaitEvent(ssel, aux->pdone[buf], 0));
    }
    mark(ssel);
    rc = consume(chunk_start(i), qc, t.mu[i % 3], nsplit, vpart, ssel);
    if (rc != BBO_OK) return rc;
    mark(ssel);
    if (two_streams) BBO_CUDA_CHECK(cudaEventRecord(aux->sdone[buf], ssel));
  }
  if (two_streams) {   // the caller's stream sees the finished selection
    BBO_CUDA_CHECK(cudaStreamWaitEvent(st, aux->sdone[int((nchunk - 1) & 1)], 0));
  }
  if (timeline) {
    // order of marks: K*(0) begin/end, then per chunk i: [K*(i+1) begin/end,] panel(i) begin/end, selection(i) begin/end
    cudaDeviceSynchronize();
    size_t m = 2;
    for (int64_t i = 0; i < nchunk && i < 8; ++i) {
      float kb = -1.f, ke = -1.f, pb = -1.f, pe = -1.f, sb = -1.f, se = -1.f;
      if (i + 1 < nchunk) {
        cudaEventElapsedTime(&kb, tl[0], tl[m]);
        cudaEventElapsedTime(&ke, tl[0], tl[m + 1]);
        m += 2;
      }
      cudaEventElapsedTime(&pb, tl[0], tl[m]);
      cudaEventElapsedTime(&pe, tl[0], tl[m + 1]);
      cudaEventElapsedTime(&sb, tl[0], tl[m + 2]);
      cudaEventElapsedTime(&se, tl[0], tl[m + 3]);
      m += 4;
      fprintf(stderr, "tf32 timeline chunk %lld: panel %.1f -> %.1f us   K*(next) %.1f -> %.1f us   select %.1f -> %.1f us\n",
              (long long)i, pb * 1e3f, pe * 1e3f, kb * 1e3f, ke * 1e3f, sb * 1e3f, se * 1e3f);
    }
    for (cudaEvent_t e : tl) cudaEventDestroy(e);
  }
  return BBO_OK;
}

}  // namespace bbo
