// 3x3 / stride 1 / pad 1 convolution with Cin = 64 as a "halo" implicit GEMM (IR-50 stage 1,
// the 64->128 widening conv, VGGish conv2).
//
// Why a second kernel: with Cin = 64 a k-step (one tap) is only 128 tensor-core cycles at N = 64,
// but the im2col kernel re-fetches the 16 KB A tile for every one of the 9 taps: 9x read
// amplification through L2 (10.7 TB/s of L2->SM traffic at 2400 frames, ncu: L2 58-65 % busy, tensor
// pipe 35 %; profiles/r01_conv_v6_s1_metrics.txt).  Here one TMA *tiled* load brings the
// (16+2) x (8+2) pixel halo of a 16 x 8 output tile (180 rows of 128 B = 22.5 KB, OOB = zero
// padding) and all 9 taps are MMAs over shifted views of the same shared-memory slab:
//   pixel (y, x) of the tile, tap (r, s)  ->  slab row (y + r) * 10 + (x + s)
// The 128 A rows of an MMA are 16 groups of 8 consecutive slab rows (x = 0..7), one group per
// output row y: group stride (descriptor SBO) = 10 rows = 1280 B, start = (r*10 + s) * 128 B.
// The 128-byte swizzle is a function of the absolute shared-memory address (bits [4,7) ^= bits
// [7,10)), so a start address that is not a multiple of 1024 B reads back exactly what TMA wrote
// with the descriptor's base-offset field left at 0 (measured: setting it to (addr >> 7) & 7 gives
// wrong results; profiles/r01_halo_conv.txt).
// L2->SM traffic drops from 9 x 16 KB to 22.5 KB per tile; weights stay resident in smem.
//
// What bounds it (round 2, tools/umma_probe.cu + profiles/r02_halo64_conv1.txt): with both operands in shared
// memory an M128 x N64 x K16 MMA cannot retire faster than one per 48 cycles (6 KB of operands at the 128 B/cycle
// the tensor core reads shared memory; N = 128: 64 cycles = its tensor floor), against a 32-cycle tensor floor.
// The kernel issues one per 58 cycles: the MMA warp is never idle (its stall samples sit ON the UTCHMMA
// instructions), the halo producer waits 83 % of its time for a free slot, the accumulator is free when needed.
// So the N = 64 layers are bound by the tensor core's shared-memory operand path, not by HBM (35 %), L2 or
// the epilogue -- which is also why (a) fusing conv1 -> conv2 of a unit would not help (same MMAs, plus
// halo recompute) and (b) an epilogue that transposes through shared memory to coalesce its stores made the
// kernel 4 % SLOWER (it takes shared-memory bandwidth from the MMAs; tried, reverted).  What is left on the
// table is the half-empty third band of a 40-row map (36000 instead of 30000 tiles per 2400 frames).
//
// Warps: 0-7 epilogue (shared with conv_igemm.cuh), 8 MMA issuer, 9 halo producer, 10 weight loader.
#pragma once
#include "conv_igemm.cuh"

namespace cer {

constexpr int kHaloThreads = 352;
constexpr int kHaloTileH = 16, kHaloTileW = 8;
constexpr int kHaloRows = (kHaloTileH + 2) * (kHaloTileW + 2);       // 180 slab rows of 128 B
constexpr int kHaloBytes = kHaloRows * 128;                          // 23040: the TMA box
constexpr int kHaloStageBytes = 23 * 1024;                           // padded to the 1024 B swizzle atom

template <int BN>
struct HaloSmem {
  static constexpr int kStages = BN == 64 ? 4 : 3;
  static constexpr int kBBytes = BN * 128;                           // one tap of the weights
  static constexpr int kBOffset = kStages * kHaloStageBytes;
  static constexpr int kBarOffset = kBOffset + 9 * kBBytes;
  static constexpr int kNumBars = 2 * kStages + 5;
  static constexpr int kTableOffset = (kBarOffset + kNumBars * 8 + 16 + 15) & ~15;
  static constexpr int kTableFloats = 10 * BN;
  static constexpr int kTotal = kTableOffset + kTableFloats * 4 + 1024;
};

__device__ __forceinline__ void tma_load_tile_4d(const CUtensorMap* m, uint32_t bar, uint32_t dst, int c, int w, int h, int n) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(n)
      : "memory");
}

template <int BN>
__global__ void __launch_bounds__(kHaloThreads, 1) conv_halo_kernel(const __grid_constant__ ConvKernelParams p) {
  using L = HaloSmem<BN>;
  constexpr int S = L::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
  uint64_t* empty_bar = full_bar + S;
  uint64_t* tfull_bar = empty_bar + S;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* bres_bar = tempty_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bres_bar + 1);
  float* s_bias = reinterpret_cast<float*>(smem + L::kTableOffset);
  float* s_alpha = s_bias + p.bias_classes * p.Cout;

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int tiles_per_frame = p.halo_bands * p.halo_cts;
  const int total_tiles = p.halo_frames * tiles_per_frame;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], kEpiWarps); }
    mbar_init(bres_bar, 1);
    fence_barrier_init();
  }
  if (warp == 9 && lane == 0) tma_prefetch_desc(&p.tmap_a);
  if (warp == 10 && lane == 0) tma_prefetch_desc(&p.tmap_b);
  if (warp == kMmaWarp) { tmem_alloc(tmem_slot, 2 * BN); tmem_relinquish(); }
  if (warp < kEpiWarps) {
    for (int i = threadIdx.x; i < p.bias_classes * p.Cout; i += kEpiWarps * 32) s_bias[i] = __ldg(p.bias + i);
    if (p.alpha != nullptr)
      for (int i = threadIdx.x; i < p.Cout; i += kEpiWarps * 32) s_alpha[i] = __ldg(p.alpha + i);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar);
  const uint32_t tfull0 = smem_u32(tfull_bar), tempty0 = smem_u32(tempty_bar);

  if (warp == 10) {
    // ---- weights: 9 taps x [BN][64] once, resident for the life of the CTA
    if (elect_one()) {
      const uint32_t bar = smem_u32(bres_bar);
      mbar_expect_tx_a(bar, 9 * L::kBBytes);
      for (int t = 0; t < 9; ++t) tma_load_2d_a(&p.tmap_b, bar, smem_base + L::kBOffset + t * L::kBBytes, t * kBlockK, 0);
    }
    __syncwarp();
  } else if (warp == 9) {
    // ---- halo producer: one tiled TMA per output tile
    uint32_t stage = 0, phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int n = tile / tiles_per_frame;
      const int rem = tile - n * tiles_per_frame;
      const int band = rem / p.halo_cts, ct = rem - band * p.halo_cts;
      mbar_wait_a(empty0 + stage * 8, phase ^ 1);
      if (elect_one()) {
        mbar_expect_tx_a(full0 + stage * 8, kHaloBytes);
        tma_load_tile_4d(&p.tmap_a, full0 + stage * 8, smem_base + stage * kHaloStageBytes, 0, ct * kHaloTileW - 1,
                         band * kHaloTileH - 1, n);
      }
      __syncwarp();
      if (++stage == S) { stage = 0; phase ^= 1; }
    }
  } else if (warp == kMmaWarp) {
    // ---- MMA issuer: 9 taps x 4 K-slices over shifted views of the slab
    constexpr uint32_t idesc = umma_idesc(kBlockM, BN, /*bf16*/ 1);
    constexpr uint32_t kHiB = kUmmaDescHiSw128;                                   // weights: dense 8-row atoms
    constexpr uint32_t kHiA = ((kHaloTileW + 2) * 128u >> 4) | (1u << 14) | (2u << 29);   // SBO = 10 rows
    const uint32_t b_lo0 = umma_desc_lo(smem_base + L::kBOffset);
    const uint32_t a_lo0 = umma_desc_lo(smem_base);
    mbar_wait_a(smem_u32(bres_bar), 0);
    uint32_t stage = 0, phase = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const uint32_t acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait_a(tempty0 + acc * 8, acc_phase ^ 1);
      mbar_wait_a(full0 + stage * 8, phase);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * BN;
      const uint32_t a_stage = a_lo0 + stage * (kHaloStageBytes >> 4);
      if (elect_one()) {
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const int rho = (tap / 3) * (kHaloTileW + 2) + (tap % 3);               // first slab row of this tap
          const uint32_t a_lo = a_stage + rho * 8;                                // 128 B = 8 x 16 B
          const uint32_t b_lo = b_lo0 + tap * (L::kBBytes >> 4);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t ad = (static_cast<uint64_t>(kHiA) << 32) | (a_lo + 2 * k);
            const uint64_t bd = (static_cast<uint64_t>(kHiB) << 32) | (b_lo + 2 * k);
            umma_f16(tmem_d, ad, bd, idesc, (tap | k) != 0 ? 1u : 0u);
          }
        }
        umma_commit_a(empty0 + stage * 8);
        umma_commit_a(tfull0 + acc * 8);
      }
      __syncwarp();
      if (++stage == S) { stage = 0; phase ^= 1; }
    }
  } else if (warp < kEpiWarps) {
    // ---- epilogue: TMEM lane l of the tile is pixel (y, x) = (l / 8, l % 8)
    int it = 0;
    const int row = (warp & 3) * 32 + lane;
    const int ty = row >> 3, tx = row & 7;
    const int n0 = (warp >> 2) * (BN / 2);
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const uint32_t acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int n = tile / tiles_per_frame;
      const int rem = tile - n * tiles_per_frame;
      const int band = rem / p.halo_cts, ct = rem - band * p.halo_cts;
      const int oh = band * kHaloTileH + ty, ow = ct * kHaloTileW + tx;
      const bool valid = oh < p.Hout && ow < p.Wout;
      int cls = 0;
      if (p.bias_classes == 9)
        cls = (oh == 0 ? 0 : (oh == p.Hout - 1 ? 2 : 1)) * 3 + (ow == 0 ? 0 : (ow == p.Wout - 1 ? 2 : 1));
      const size_t m = (static_cast<size_t>(n) * p.Hout + oh) * p.Wout + ow;
      if (p.res != nullptr) {
        // the residual read is the epilogue's only long-latency load: pull the NEXT tile's line into L2 now,
        // a whole tile ahead of the __ldg that needs it
        const int nt = tile + gridDim.x;
        if (nt < total_tiles) {
          const int n2 = nt / tiles_per_frame;
          const int rem2 = nt - n2 * tiles_per_frame;
          const int band2 = rem2 / p.halo_cts, ct2 = rem2 - band2 * p.halo_cts;
          const int oh2 = band2 * kHaloTileH + ty, ow2 = ct2 * kHaloTileW + tx;
          if (oh2 < p.Hout) {
            const __nv_bfloat16* rp = p.res + ((static_cast<size_t>(n2) * p.Hout + oh2) * p.Wout + ow2) * p.Cout + n0;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(rp));
          }
        }
      }
      const bool pool_store = p.pool_xor && valid && !(oh & 1) && !(ow & 1);
      const size_t pool_off = ((static_cast<size_t>(n) * (p.Hout >> 1) + (oh >> 1)) * (p.Wout >> 1) + (ow >> 1)) * p.Cout + n0;
      conv_epilogue_core<BN>(p, s_bias, s_alpha, tmem_base + acc * BN, n0, warp, tfull0 + acc * 8, acc_phase, valid, cls,
                             m * p.Cout + n0, tempty0 + acc * 8, pool_store, pool_off);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * BN);
  }
}

}  // namespace cer
