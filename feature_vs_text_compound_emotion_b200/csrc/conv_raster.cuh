// 3x3 / stride 1 / pad 1 convolution over a "padded raster" activation, CTA pair, weights resident
// (IR-50 stage 2: 128 -> 128 @20x20, six launches per forward).
//
// Why: the im2col pair kernel (conv_igemm2_bres_kernel) fetches a 16 KB A tile for every one of the
// 18 k-steps: 288 KB per 128-pixel tile through L2 = 62.5 B/cycle/SM at the tensor floor, more than
// L2 delivers to 148 SMs at once (ncu: tensor pipe 58 %, L2 -> SM 13.7 TB/s; profiles/r01_full_pair128.txt).
//
// Padded raster: a stage-2 activation is stored as [frames][H+1][W+1][C] with one zero column after every
// row and one zero row after every frame, i.e. position q = n*P + y*Wp + x (Wp = W+1, P = (H+1)*Wp).  Then
//   input(q, tap r s) = act[q + (r-1)*Wp + (s-1)]
// for EVERY output position, frame borders included: x-1 = -1 lands on the previous row's zero column,
// y-1 = -1 on the previous frame's zero row (or before the tensor: TMA fills zeros), y+1 = H on the
// frame's own zero row.  An M tile is any 128 consecutive positions, its operand for ALL nine taps of one
// 64-channel half is ONE 2-D TMA box of 128 + 2*(Wp+1) rows x 128 B (172 rows = 21.5 KB at 20x20), and a tap
// is an MMA over the box shifted by r*Wp + s rows: dense 8-row swizzle atoms (SBO = 1024 B), start address
// any multiple of 128 B (the swizzle is a function of the absolute address, see conv_halo.cuh).
// L2 -> SM traffic per tile drops from 288 KB to 43 KB; 400 of 441 positions are real pixels (9.3 % of
// the MMAs compute pad positions, stored as zeros so that the next layer can read the tensor as is).
//
// K order: channel half 0 (taps 0..8), then channel half 1; weights [Cout][tap][Cin] stay resident per CTA
// (this CTA's 64 output channels of all 18 k-steps = 144 KB), three 22 KB box slots ring beside them.
// Protocol, TMEM, epilogue: as conv_igemm2_bres_kernel.
#pragma once
#include "conv_igemm2.cuh"

namespace cer {

// Instantiated as <128, 3, 18, 176>: 128 -> 128 maps up to 22 wide (box rows 128 + 2 * (W + 2) <= 176).
// Tried and dropped: <64, 6, 9, 216> for IR-50 stage 1 (64 -> 64 @40x40, pair MMAs of N = 64, stem storing the
// padded raster): 2.21 ms per 2400 frames against 2.07 ms on conv_strip_kernel.
template <int BN, int SLOTS, int KSTEPS, int BOX_ROWS_MAX>
struct RasterSmem {
  static constexpr int kSlotBytes = BOX_ROWS_MAX * 128;        // a multiple of 1024: slots stay swizzle-atom aligned
  static_assert(kSlotBytes % 1024 == 0, "slot size must keep the 1024 B swizzle atoms aligned");
  static constexpr int kBBytes = (BN / 2) * 128;               // this CTA's half of one k-step of B
  static constexpr int kBOffset = SLOTS * kSlotBytes;
  static constexpr int kBarOffset = kBOffset + KSTEPS * kBBytes;
  static constexpr int kNumBars = 2 * SLOTS + 5;               // full, empty, tfull[2], tempty[2], bres
  static constexpr int kTableOffset = (kBarOffset + kNumBars * 8 + 16 + 15) & ~15;
  static constexpr int kTableFloats = 10 * BN;
  static constexpr int kTotal = kTableOffset + kTableFloats * 4 + 1024;
};

template <int BN, int SLOTS, int KSTEPS, int BOX_ROWS_MAX>
__global__ void __launch_bounds__(kConvThreads, 1) conv_raster2_kernel(const __grid_constant__ ConvKernelParams p) {
  using L = RasterSmem<BN, SLOTS, KSTEPS, BOX_ROWS_MAX>;
  constexpr int kRasterSlotBytes = L::kSlotBytes;
  constexpr int kHalves = KSTEPS / 9;                          // 64-channel halves of Cin
  static_assert(KSTEPS % 9 == 0, "3x3 taps");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
  uint64_t* empty_bar = full_bar + SLOTS;
  uint64_t* tfull_bar = empty_bar + SLOTS;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* bres_bar = tempty_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bres_bar + 1);
  float* s_bias = reinterpret_cast<float*>(smem + L::kTableOffset);
  float* s_alpha = s_bias + p.bias_classes * p.Cout;

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const int total_pos = p.rs_frames * p.rs_P;                  // positions that hold (or pad) real frames
  const int total_ptiles = (total_pos + 2 * kBlockM - 1) / (2 * kBlockM);
  const int wp = p.rs_wp;
  const int box_bytes = (kBlockM + 2 * (wp + 1)) * 128;

  if (threadIdx.x == 0) {
    for (int s = 0; s < SLOTS; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], 2 * kEpiWarps); }
    mbar_init(bres_bar, 1);
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
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();

  const uint32_t smem_base = smem_u32(smem);
  const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar);
  const uint32_t tfull0 = smem_u32(tfull_bar), tempty0 = smem_u32(tempty_bar);

  if (warp == kProdWarp0 + 2) {
    // ---- resident weights: every CTA loads its BN/2 output channels of all k-steps, once
    if (elect_one()) {
      const uint32_t bar = smem_u32(bres_bar) & kPeerBitMask;            // the leader's barrier
      if (rank == 0) mbar_expect_tx_a(bar, 2 * KSTEPS * L::kBBytes);
      for (int ks = 0; ks < KSTEPS; ++ks)
        tma2_load_2d(&p.tmap_b, bar, smem_base + L::kBOffset + ks * L::kBBytes, ks * kBlockK, (int)rank * (BN / 2));
    }
    __syncwarp();
  } else if (warp == kProdWarp0) {
    // ---- box producer: one 2-D box per (tile, channel half); rows before / after the tensor read as zero
    uint32_t g = 0;
    for (int pt = pair; pt < total_ptiles; pt += num_pairs) {
      const int q0 = (2 * pt + (int)rank) * kBlockM;
      for (int c = 0; c < kHalves; ++c, ++g) {
        const uint32_t slot = g % SLOTS, phase = (g / SLOTS) & 1;
        mbar_wait_a(empty0 + slot * 8, phase ^ 1);
        if (elect_one()) {
          const uint32_t fb = (full0 + slot * 8) & kPeerBitMask;
          if (rank == 0) mbar_expect_tx_a(fb, 2 * box_bytes);
          tma2_load_2d(&p.tmap_a, fb, smem_base + slot * kRasterSlotBytes, c * kBlockK, q0 - (wp + 1));
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
        for (int c = 0; c < kHalves; ++c, ++g) {
          const uint32_t slot = g % SLOTS, phase = (g / SLOTS) & 1;
          mbar_wait_a(full0 + slot * 8, phase);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t a_slot = a_lo0 + slot * (kRasterSlotBytes >> 4);
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
              const uint32_t a_lo = a_slot + ((tap / 3) * wp + (tap % 3)) * 8;       // shifted by r*Wp + s rows of 128 B
              const uint32_t b_lo = b_lo0 + (tap * kHalves + c) * (L::kBBytes >> 4);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma2_f16(tmem_d, umma_desc_from_lo(a_lo + 2 * k), umma_desc_from_lo(b_lo + 2 * k), idesc,
                          (c | tap | k) != 0 ? 1u : 0u);
            }
            umma2_commit_mc(empty0 + slot * 8);
            if (c == kHalves - 1) umma2_commit_mc(tfull0 + acc * 8);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp < kEpiWarps) {
    // ---- epilogue: TMEM lane l of this CTA's tile is position q0 + l
    const uint32_t tempty_leader = tempty0 & kPeerBitMask;
    const int row = (warp & 3) * 32 + lane;
    const int n0 = (warp >> 2) * (BN / 2);
    int it = 0;
    for (int pt = pair; pt < total_ptiles; pt += num_pairs, ++it) {
      const uint32_t acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int q = (2 * pt + (int)rank) * kBlockM + row;
      const int n = q / p.rs_P;
      const int rem = q - n * p.rs_P;
      const int oh = rem / wp, ow = rem - oh * wp;
      const bool in_range = q < total_pos;
      const bool valid = in_range && oh < p.Hout && ow < p.Wout;
      int cls = 0;
      if (p.bias_classes == 9)
        cls = (oh == 0 ? 0 : (oh >= p.Hout - 1 ? 2 : 1)) * 3 + (ow == 0 ? 0 : (ow >= p.Wout - 1 ? 2 : 1));
      const size_t q_off = static_cast<size_t>(q) * p.Cout + n0;                   // padded raster: linear in q
      const size_t m_off = ((static_cast<size_t>(n) * p.Hout + oh) * p.Wout + ow) * p.Cout + n0;
      const bool padded_out = p.out_wp != 0;
      conv_epilogue_core<BN>(p, s_bias, s_alpha, tmem_base + acc * BN, n0, warp, tfull0 + acc * 8, acc_phase, valid, cls,
                             padded_out ? q_off : m_off, tempty_leader + acc * 8, false, 0, /*res_off*/ q_off,
                             /*zero_store*/ padded_out && in_range);
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

// Zero the pad positions (one column per row, one row per frame) of a padded-raster tensor written by a
// kernel that only stores real pixels (the stride-2 conv in front of the stage).
__global__ void raster_zero_pads_kernel(__nv_bfloat16* __restrict__ act, int frames, int H, int W, int C) {
  const int wp = W + 1, per_frame = H + wp;                   // H pad columns + one pad row of Wp positions
  const int vec = C / 8;                                      // uint4 per position
  const long long total = static_cast<long long>(frames) * per_frame * vec;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(i % vec);
    const long long pi = i / vec;
    const int n = static_cast<int>(pi / per_frame);
    const int k = static_cast<int>(pi - static_cast<long long>(n) * per_frame);
    const int pos = k < H ? k * wp + W : H * wp + (k - H);
    reinterpret_cast<uint4*>(act + (static_cast<size_t>(n) * (H + 1) * wp + pos) * C)[v] = make_uint4(0, 0, 0, 0);
  }
}

// Padded raster -> dense NHWC (tests / cer_ir50_debug_activation).
__global__ void raster_unpad_kernel(const __nv_bfloat16* __restrict__ src, __nv_bfloat16* __restrict__ dst, int frames, int H,
                                    int W, int C) {
  const int wp = W + 1, vec = C / 8;
  const long long total = static_cast<long long>(frames) * H * W * vec;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(i % vec);
    const long long pix = i / vec;
    const int x = static_cast<int>(pix % W);
    const long long t = pix / W;
    const int y = static_cast<int>(t % H);
    const long long n = t / H;
    reinterpret_cast<uint4*>(dst + static_cast<size_t>(pix) * C)[v] =
        reinterpret_cast<const uint4*>(src + (static_cast<size_t>(n) * (H + 1) * wp + y * wp + x) * C)[v];
  }
}

}  // namespace cer
