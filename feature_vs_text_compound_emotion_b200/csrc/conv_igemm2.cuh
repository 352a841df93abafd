// CTA-pair variant of the implicit-GEMM convolution: tcgen05.mma.cta_group::2, M = 256 per pair.
//
// Why: one SM can pull only ~80 B/cycle through the TMA path (tools/tma_probe.cu) while the
// single-CTA 128x256 tile needs 96 B/cycle (16 KB of A + 32 KB of B per 512-cycle k-step).  In a
// pair each CTA loads its own 128 rows of A but only HALF of the B tile (BN/2 output channels); the
// pair's MMA reads A and B from both CTAs' shared memory, so the per-SM operand traffic drops to
// 62.5 B/cycle and one issuing thread feeds two tensor cores.
//
// Protocol (cluster of 2 along M, rank 0 = leader):
//   * full[s] lives in the leader.  The leader's producers arrive with expect_tx covering BOTH CTAs'
//     bytes; both CTAs issue `cp.async.bulk.tensor...cta_group::2` loads whose mbarrier operand has
//     the peer bit cleared, i.e. points at the leader's barrier.
//   * empty[s] and tfull[a] exist in both CTAs; the leader's MMA warp signals them with
//     `tcgen05.commit.cta_group::2...multicast::cluster` (mask 0b11).
//   * tempty[a] lives in the leader and counts one arrival per epilogue warp of both CTAs
//     (remote `mbarrier.arrive.relaxed.cluster` from the peer, issued as soon as the accumulator
//     is in registers -- before the epilogue's global stores, see conv_epilogue_core).
//   * TMEM is allocated with `tcgen05.alloc.cta_group::2` by the same warp in both CTAs; each CTA
//     keeps its own 128 rows x BN columns of the accumulator and runs the ordinary epilogue on them.
// Requires ksteps % STAGES == 0 (stage-aligned unrolled loops, as in the ALIGNED 1-CTA variant).
#pragma once
#include "conv_igemm.cuh"

namespace cer {

constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;      // clears the CTA-rank bit of a shared::cluster address (pair)

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma2_load_2d(const CUtensorMap* m, uint32_t leader_bar, uint32_t dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma2_load_im2col_4d(const CUtensorMap* m, uint32_t leader_bar, uint32_t dst, int c, int w,
                                                    int h, int n, uint16_t off_w, uint16_t off_h) {
  asm volatile(
      "cp.async.bulk.tensor.4d.im2col.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(leader_bar), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w), "h"(off_h)
      : "memory");
}
__device__ __forceinline__ void umma2_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the barrier at the same shared-memory offset in both CTAs of the pair
__device__ __forceinline__ void umma2_commit_mc(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3)
               : "memory");
}

template <int BN, int STAGES>
struct Conv2Smem {
  static constexpr int kBBytes = (BN / 2) * 128;             // this CTA's half of the B tile
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBarOffset = STAGES * kStageBytes;
  static constexpr int kNumBars = 2 * STAGES + 4;            // full, empty, tfull[2], tempty[2]
  static constexpr int kTableOffset = (kBarOffset + kNumBars * 8 + 16 + 15) & ~15;
  static constexpr int kTableFloats = BN == 256 ? 8192 : 10 * BN;
  static constexpr int kTotal = kTableOffset + kTableFloats * 4 + 1024;
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(kConvThreads, 1) conv_igemm2_kernel(const __grid_constant__ ConvKernelParams p) {
  using L = Conv2Smem<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  float* s_bias = reinterpret_cast<float*>(smem + L::kTableOffset);
  float* s_alpha = s_bias + p.bias_classes * p.Cout;

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();                     // 0 = leader
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const int num_pair_m = (p.num_m_tiles + 1) >> 1;
  const int total_ptiles = num_pair_m * p.num_n_tiles;
  const int ksteps = p.ksteps_main + p.ksteps2;
  const int rounds = ksteps / STAGES;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 2);       // leader's A and B producers (each expect_tx for both CTAs); unused in the peer
      mbar_init(&empty_bar[s], 1);      // one multicast commit
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);                  // one multicast commit
      mbar_init(&tempty_bar[a], 2 * kEpiWarps);     // one arrival per epilogue warp of both CTAs (leader's copy is used)
    }
    fence_barrier_init();
  }
  if (warp == kProdWarp0 && lane == 0) tma_prefetch_desc(&p.tmap_a);
  if (warp == kProdWarp0 + 2 && lane == 0) tma_prefetch_desc(&p.tmap_b);
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(2 * BN)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  if (warp < kEpiWarps) {
    for (int i = threadIdx.x; i < p.bias_classes * p.Cout; i += kEpiWarps * 32) s_bias[i] = __ldg(p.bias + i);
    if (p.alpha != nullptr)
      for (int i = threadIdx.x; i < p.Cout; i += kEpiWarps * 32) s_alpha[i] = __ldg(p.alpha + i);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                     // both CTAs' barriers are initialised before any remote signal
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();

  const uint32_t smem_base = smem_u32(smem);
  const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar);
  const uint32_t tfull0 = smem_u32(tfull_bar), tempty0 = smem_u32(tempty_bar);

  if (warp >= kProdWarp0) {
    // ---------------- producers: warps 9,10 = A (stages pa, pa+2, ...), warps 11,12 = B ----------------
    const int pw = warp - kProdWarp0;
    const bool a_role = pw < 2;
    const int pa = a_role ? pw : pw - 2;
    const int hw = p.Hout * p.Wout;
    uint32_t phase = 0;
    const uint32_t sa0 = smem_base + pa * L::kStageBytes + (a_role ? 0 : kABytes);
    const uint32_t eb0 = empty0 + pa * 8;
    const uint32_t fb0_leader = (full0 + pa * 8) & kPeerBitMask;      // the leader CTA's full barrier
    const uint32_t tx_bytes = 2 * (a_role ? kABytes : L::kBBytes);    // both CTAs' loads land on the leader's barrier
    for (int pt = pair; pt < total_ptiles; pt += num_pairs) {
      const int pm = pt / p.num_n_tiles;
      const int n_tile = pt - pm * p.num_n_tiles;
      const int m_tile = 2 * pm + rank;
      int cw = 0, ch = 0, n_img = 0;
      if (a_role) {
        const int m0 = m_tile * kBlockM;
        n_img = m0 / hw;
        const int rem = m0 - n_img * hw;
        const int oh = rem / p.Wout;
        const int ow = rem - oh * p.Wout;
        cw = ow * p.stride - p.pad; ch = oh * p.stride - p.pad;
      }
      const int n0 = n_tile * BN + rank * (BN / 2);
      for (int r = 0; r < rounds; ++r) {
#pragma unroll
        for (int j = 0; j < STAGES; j += 2) {
          const int ks = r * STAGES + pa + j;
          mbar_wait_a(eb0 + j * 8, phase ^ 1);
          if (a_role) {
            int tap = 0, chunk = ks;
            if (p.ksize == 3) { tap = ks >> p.cin_shift; chunk = ks & (p.cin_chunks - 1); }
            const int rr = (tap * 11) >> 5;
            const int ss = tap - rr * 3;
            if (elect_one()) {
              if (rank == 0) mbar_expect_tx_a(fb0_leader + j * 8, tx_bytes);
              tma2_load_im2col_4d(&p.tmap_a, fb0_leader + j * 8, sa0 + j * L::kStageBytes, chunk * kBlockK, cw, ch, n_img,
                                  (uint16_t)ss, (uint16_t)rr);
            }
          } else {
            if (elect_one()) {
              if (rank == 0) mbar_expect_tx_a(fb0_leader + j * 8, tx_bytes);
              tma2_load_2d(&p.tmap_b, fb0_leader + j * 8, sa0 + j * L::kStageBytes, ks * kBlockK, n0);
            }
          }
          __syncwarp();
        }
        phase ^= 1;
      }
    }
  } else if (warp == kMmaWarp) {
    // ---------------- MMA issuer: leader CTA only ----------------
    if (rank == 0) {
      constexpr uint32_t idesc = umma_idesc(2 * kBlockM, BN, /*bf16*/ 1);
      constexpr uint32_t kStageLo = L::kStageBytes >> 4;
      const uint32_t a_lo0 = umma_desc_lo(smem_base);
      uint32_t phase = 0;
      int it = 0;
      for (int pt = pair; pt < total_ptiles; pt += num_pairs, ++it) {
        const uint32_t acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait_a(tempty0 + acc * 8, acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int r = 0; r < rounds; ++r) {
#pragma unroll
          for (int sidx = 0; sidx < STAGES; ++sidx) {
            mbar_wait_a(full0 + sidx * 8, phase);
            tc_fence_after();
            if (elect_one()) {
              const uint32_t a_lo = a_lo0 + sidx * kStageLo;
              const uint32_t b_lo = a_lo + (kABytes >> 4);
              umma2_f16(tmem_d, umma_desc_from_lo(a_lo), umma_desc_from_lo(b_lo), idesc, (r | sidx) != 0 ? 1u : 0u);
              umma2_f16(tmem_d, umma_desc_from_lo(a_lo + 2), umma_desc_from_lo(b_lo + 2), idesc, 1u);
              umma2_f16(tmem_d, umma_desc_from_lo(a_lo + 4), umma_desc_from_lo(b_lo + 4), idesc, 1u);
              umma2_f16(tmem_d, umma_desc_from_lo(a_lo + 6), umma_desc_from_lo(b_lo + 6), idesc, 1u);
              umma2_commit_mc(empty0 + sidx * 8);
              if (sidx == STAGES - 1 && r == rounds - 1) umma2_commit_mc(tfull0 + acc * 8);
            }
            __syncwarp();
          }
          phase ^= 1;
        }
      }
    }
  } else {
    // ---------------- epilogue: every CTA drains its own 128 accumulator rows ----------------
    const uint32_t tempty_leader = tempty0 & kPeerBitMask;
    int it = 0;
    for (int pt = pair; pt < total_ptiles; pt += num_pairs, ++it) {
      const uint32_t acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int pm = pt / p.num_n_tiles;
      const int n_tile = pt - pm * p.num_n_tiles;
      conv_epilogue_tile<BN>(p, s_bias, s_alpha, tmem_base + acc * BN, 2 * pm + (int)rank, n_tile, warp, lane,
                             tfull0 + acc * 8, acc_phase, tempty_leader + acc * 8);
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                     // nobody leaves while the peer may still signal / read
  if (warp == kMmaWarp) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * BN) : "memory");
  }
}

// ------------------------------------------------------------------------------------------
// Weights-resident CTA-pair variant (IR-50 stage 2: 128 -> 128 @20x20, 18 k-steps, one n-tile).
// The streaming pair kernel needs 16 KB of A + 8 KB of B per 256-cycle k-step = 96 B/cycle/SM, above
// the ~80 B/cycle the TMA path delivers (tools/tma_probe.cu; ncu: tensor pipe 58 %).  Here each CTA
// keeps ITS HALF of the whole weight matrix (ksteps x 64 rows x 128 B = 144 KB) in shared memory for
// the life of the CTA and the ring carries only A (64 B/cycle/SM).  Four A stages fit beside it, so
// the ring is walked with run-time stage indices (a k-step is 256 tensor cycles at N = 128: the few
// extra issue instructions do not matter).
// ------------------------------------------------------------------------------------------
template <int BN, int STAGES, int KSTEPS>
struct Conv2BresSmem {
  static constexpr int kBBytes = (BN / 2) * 128;             // this CTA's half of one k-step of B
  static constexpr int kBOffset = STAGES * kABytes;
  static constexpr int kBarOffset = kBOffset + KSTEPS * kBBytes;
  static constexpr int kNumBars = 2 * STAGES + 5;            // full, empty, tfull[2], tempty[2], bres
  static constexpr int kTableOffset = (kBarOffset + kNumBars * 8 + 16 + 15) & ~15;
  static constexpr int kTableFloats = 10 * BN;
  static constexpr int kTotal = kTableOffset + kTableFloats * 4 + 1024;
};

template <int BN, int STAGES, int KSTEPS>
__global__ void __launch_bounds__(kConvThreads, 1) conv_igemm2_bres_kernel(const __grid_constant__ ConvKernelParams p) {
  using L = Conv2BresSmem<BN, STAGES, KSTEPS>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* bres_bar = tempty_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bres_bar + 1);
  float* s_bias = reinterpret_cast<float*>(smem + L::kTableOffset);
  float* s_alpha = s_bias + p.bias_classes * p.Cout;

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const int total_ptiles = (p.num_m_tiles + 1) >> 1;          // one n-tile

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], 2 * kEpiWarps); }
    mbar_init(bres_bar, 1);
    fence_barrier_init();
  }
  if (warp == kProdWarp0 && lane == 0) {
    tma_prefetch_desc(&p.tmap_a);
    if (p.ksteps2 > 0) tma_prefetch_desc(&p.tmap_a2);
  }
  if (warp == kProdWarp0 + 2 && lane == 0) tma_prefetch_desc(&p.tmap_b);
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(2 * BN)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  if (warp < kEpiWarps) {
    for (int i = threadIdx.x; i < p.bias_classes * p.Cout; i += kEpiWarps * 32) s_bias[i] = __ldg(p.bias + i);
    if (p.alpha != nullptr)
      for (int i = threadIdx.x; i < p.Cout; i += kEpiWarps * 32) s_alpha[i] = __ldg(p.alpha + i);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();

  const uint32_t smem_base = smem_u32(smem);
  const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar);
  const uint32_t tfull0 = smem_u32(tfull_bar), tempty0 = smem_u32(tempty_bar);

  if (warp == kProdWarp0 + 2) {
    // ---- resident weights: every CTA loads its 64 output channels of all k-steps, once
    if (elect_one()) {
      const uint32_t bar = smem_u32(bres_bar) & kPeerBitMask;            // the leader's barrier
      if (rank == 0) mbar_expect_tx_a(bar, 2 * KSTEPS * L::kBBytes);
      for (int ks = 0; ks < KSTEPS; ++ks)
        tma2_load_2d(&p.tmap_b, bar, smem_base + L::kBOffset + ks * L::kBBytes, ks * kBlockK, (int)rank * (BN / 2));
    }
    __syncwarp();
  } else if (warp == kProdWarp0 || warp == kProdWarp0 + 1) {
    // ---- A producers: warp q takes the k-steps ks = q (mod 2) of every tile; g = g0 + ks counts the k-steps of this
    // pair across tiles (KSTEPS may be odd: 9 taps x 2 chunks + one k-step of a fused 1x1 projection)
    const int q = warp - kProdWarp0;
    const int hw = p.Hout * p.Wout;
    uint32_t g0 = 0;
    for (int pt = pair; pt < total_ptiles; pt += num_pairs, g0 += KSTEPS) {
      const int m0 = (2 * pt + (int)rank) * kBlockM;
      const int n_img = m0 / hw;
      const int rem = m0 - n_img * hw;
      const int oh = rem / p.Wout, ow = rem - oh * p.Wout;
      const int cw = ow * p.stride - p.pad, ch = oh * p.stride - p.pad;
      for (int ks = q; ks < KSTEPS; ks += 2) {
        const uint32_t g = g0 + ks;
        const uint32_t stage = g % STAGES, phase = (g / STAGES) & 1;
        mbar_wait_a(empty0 + stage * 8, phase ^ 1);
        int tap = 0, chunk = ks;
        if (p.ksize == 3) { tap = ks >> p.cin_shift; chunk = ks & (p.cin_chunks - 1); }
        const int rr = (tap * 11) >> 5, ss = tap - rr * 3;
        if (elect_one()) {
          const uint32_t fb = (full0 + stage * 8) & kPeerBitMask;
          if (rank == 0) mbar_expect_tx_a(fb, 2 * kABytes);
          if (ks < p.ksteps_main)
            tma2_load_im2col_4d(&p.tmap_a, fb, smem_base + stage * kABytes, chunk * kBlockK, cw, ch, n_img, (uint16_t)ss,
                                (uint16_t)rr);
          else                                                   // fused 1x1 / stride-2 projection of the unit's input
            tma2_load_im2col_4d(&p.tmap_a2, fb, smem_base + stage * kABytes, (ks - p.ksteps_main) * kBlockK, ow * p.stride2,
                                oh * p.stride2, n_img, 0, 0);
        }
        __syncwarp();
      }
    }
  } else if (warp == kMmaWarp) {
    if (rank == 0) {
      constexpr uint32_t idesc = umma_idesc(2 * kBlockM, BN, /*bf16*/ 1);
      const uint32_t a_lo0 = umma_desc_lo(smem_base);
      const uint32_t b_lo0 = umma_desc_lo(smem_base + L::kBOffset);
      mbar_wait_a(smem_u32(bres_bar), 0);
      uint32_t g = 0;
      int it = 0;
      for (int pt = pair; pt < total_ptiles; pt += num_pairs, ++it) {
        const uint32_t acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait_a(tempty0 + acc * 8, acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int ks = 0; ks < KSTEPS; ++ks, ++g) {
          const uint32_t stage = g % STAGES, phase = (g / STAGES) & 1;
          mbar_wait_a(full0 + stage * 8, phase);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t a_lo = a_lo0 + stage * (kABytes >> 4);
            const uint32_t b_lo = b_lo0 + ks * (L::kBBytes >> 4);
            umma2_f16(tmem_d, umma_desc_from_lo(a_lo), umma_desc_from_lo(b_lo), idesc, ks != 0 ? 1u : 0u);
            umma2_f16(tmem_d, umma_desc_from_lo(a_lo + 2), umma_desc_from_lo(b_lo + 2), idesc, 1u);
            umma2_f16(tmem_d, umma_desc_from_lo(a_lo + 4), umma_desc_from_lo(b_lo + 4), idesc, 1u);
            umma2_f16(tmem_d, umma_desc_from_lo(a_lo + 6), umma_desc_from_lo(b_lo + 6), idesc, 1u);
            umma2_commit_mc(empty0 + stage * 8);
            if (ks == KSTEPS - 1) umma2_commit_mc(tfull0 + acc * 8);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp < kEpiWarps) {
    const uint32_t tempty_leader = tempty0 & kPeerBitMask;
    int it = 0;
    for (int pt = pair; pt < total_ptiles; pt += num_pairs, ++it) {
      const uint32_t acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      conv_epilogue_tile<BN>(p, s_bias, s_alpha, tmem_base + acc * BN, 2 * pt + (int)rank, 0, warp, lane, tfull0 + acc * 8,
                             acc_phase, tempty_leader + acc * 8);
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == kMmaWarp) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * BN) : "memory");
  }
}

}  // namespace cer
