// 3x3 / stride 1 / pad 1 convolution with Cin = 64 -> Cout = 64 as a "strip" implicit GEMM: the halo kernel
// (conv_halo.cuh) without its half-empty tiles.
//
// conv_halo tiles a map into 16-row x 8-column tiles, so a 40-row map costs three bands of which the third is
// half empty: 36000 instead of 30000 MMA tiles per 2400 frames, on a layer that is bound by the MMA itself
// (N = 64 operands cannot leave shared memory faster than one MMA per 48 cycles, tools/umma_probe.cu).
// Here the map is cut into 8-pixel-wide column STRIPS and all strips of all frames are chained into one long
// virtual column of rows; an M tile is ANY 16 consecutive virtual rows (16 groups of 8 pixels), so tiles run
// across strip and frame boundaries and none is padded (H % 8 == 0: every 8-row half tile lies in one strip).
// The price is the vertical halo: a tile that crosses a boundary has two unrelated 10-row halos, which cannot
// share one slab with a uniform group stride.  So the slab is split by filter ROW: for r = 0, 1, 2 a slab of
// 16 groups x 10 pixels holds the rows (y + r - 1) of the tile, each loaded as two 8-row TMA boxes (one per
// half tile, 10 KB each, out-of-bounds = the conv's zero padding).  Tap (r, s) is then an MMA over slab r
// with start offset s pixels and the same 1280-byte group stride as in the halo kernel.  L2->SM traffic is
// 60 KB per tile (halo: 22.5 KB, im2col: 144 KB); shared memory: 2 stages x 60 KB + 72 KB of resident weights.
//
// Warps: 0-7 epilogue, 8 MMA issuer, 9 producer of half tile 0, 10 weight loader, then producer of half tile 1.
#pragma once
#include "conv_halo.cuh"

namespace cer {

constexpr int kStripBoxRows = 8;
constexpr int kStripHalfBytes = kStripBoxRows * (kHaloTileW + 2) * 128;      // 10240: one TMA box
constexpr int kStripSlabBytes = 2 * kStripHalfBytes;                         // 20480: one filter row, 16 groups
constexpr int kStripStageBytes = 3 * kStripSlabBytes;                        // 61440 = 60 KB
constexpr int kStripStages = 2;

struct StripSmem {
  static constexpr int kBBytes = 64 * 128;                                   // one tap of the weights (BN = 64)
  static constexpr int kBOffset = kStripStages * kStripStageBytes;
  static constexpr int kBarOffset = kBOffset + 9 * kBBytes;
  static constexpr int kNumBars = 2 * kStripStages + 5;
  static constexpr int kTableOffset = (kBarOffset + kNumBars * 8 + 16 + 15) & ~15;
  static constexpr int kTableFloats = 10 * 64;
  static constexpr int kTotal = kTableOffset + kTableFloats * 4 + 1024;
};

// p.tmap_a: tiled 4-D map of the NHWC input with box {64 ch, 10 px, 8 rows, 1 frame}; p.halo_frames = frames,
// p.halo_cts = strips per frame (W / 8).  M tiles: ceil(frames * strips * H / 16).
__global__ void __launch_bounds__(kHaloThreads, 1) conv_strip_kernel(const __grid_constant__ ConvKernelParams p) {
  using L = StripSmem;
  constexpr int S = kStripStages;
  constexpr int BN = 64;
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
  const int H = p.Hout;
  const long long vrows = static_cast<long long>(p.halo_frames) * p.halo_cts * H;     // virtual rows
  const int total_tiles = static_cast<int>((vrows + 15) >> 4);

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(&full_bar[s], 2); mbar_init(&empty_bar[s], 1); }     // two producers
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

  if (warp == 9 || warp == 10) {
    const int half = warp - 9;
    if (half == 1) {
      // ---- weights: 9 taps x [64][64] once, resident for the life of the CTA
      if (elect_one()) {
        const uint32_t bar = smem_u32(bres_bar);
        mbar_expect_tx_a(bar, 9 * L::kBBytes);
        for (int t = 0; t < 9; ++t) tma_load_2d_a(&p.tmap_b, bar, smem_base + L::kBOffset + t * L::kBBytes, t * kBlockK, 0);
      }
      __syncwarp();
    }
    // ---- producer of half tile `half`: three 8-row boxes (filter rows r = 0, 1, 2) per tile
    uint32_t stage = 0, phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const long long v0 = static_cast<long long>(tile) * 16 + half * 8;              // first virtual row of this half
      const bool live = v0 < vrows;
      const int sidx = static_cast<int>(v0 / H);
      const int y0 = static_cast<int>(v0 - static_cast<long long>(sidx) * H);
      const int n = sidx / p.halo_cts, ct = sidx - n * p.halo_cts;
      mbar_wait_a(empty0 + stage * 8, phase ^ 1);
      if (elect_one()) {
        mbar_expect_tx_a(full0 + stage * 8, live ? 3 * kStripHalfBytes : 0);
        if (live) {
          const uint32_t dst = smem_base + stage * kStripStageBytes + half * kStripHalfBytes;
#pragma unroll
          for (int r = 0; r < 3; ++r)
            tma_load_tile_4d(&p.tmap_a, full0 + stage * 8, dst + r * kStripSlabBytes, 0, ct * kHaloTileW - 1, y0 + r - 1, n);
        }
      }
      __syncwarp();
      if (++stage == S) { stage = 0; phase ^= 1; }
    }
  } else if (warp == kMmaWarp) {
    // ---- MMA issuer: tap (r, s) = slab r, start offset s pixels, 4 K-slices
    constexpr uint32_t idesc = umma_idesc(kBlockM, BN, /*bf16*/ 1);
    constexpr uint32_t kHiB = kUmmaDescHiSw128;
    constexpr uint32_t kHiA = ((kHaloTileW + 2) * 128u >> 4) | (1u << 14) | (2u << 29);     // group stride = 10 pixels
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
      const uint32_t a_stage = a_lo0 + stage * (kStripStageBytes >> 4);
      if (elect_one()) {
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const uint32_t a_lo = a_stage + (tap / 3) * (kStripSlabBytes >> 4) + (tap % 3) * 8;     // + s pixels of 128 B
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
    // ---- epilogue: TMEM lane l of the tile is (virtual row 16 * tile + l / 8, column l % 8 of its strip)
    int it = 0;
    const int row = (warp & 3) * 32 + lane;
    const int g = row >> 3, tx = row & 7;
    const int n0 = (warp >> 2) * (BN / 2);
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const uint32_t acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const long long v = static_cast<long long>(tile) * 16 + g;
      const bool valid = v < vrows;
      const long long vc = valid ? v : 0;
      const int sidx = static_cast<int>(vc / H);
      const int oh = static_cast<int>(vc - static_cast<long long>(sidx) * H);
      const int n = sidx / p.halo_cts, ct = sidx - n * p.halo_cts;
      const int ow = ct * kHaloTileW + tx;
      int cls = 0;
      if (p.bias_classes == 9)
        cls = (oh == 0 ? 0 : (oh == p.Hout - 1 ? 2 : 1)) * 3 + (ow == 0 ? 0 : (ow == p.Wout - 1 ? 2 : 1));
      const size_t m = (static_cast<size_t>(n) * p.Hout + oh) * p.Wout + ow;
      conv_epilogue_core<BN>(p, s_bias, s_alpha, tmem_base + acc * BN, n0, warp, tfull0 + acc * 8, acc_phase, valid, cls,
                             m * p.Cout + n0, tempty0 + acc * 8, false, 0);
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
