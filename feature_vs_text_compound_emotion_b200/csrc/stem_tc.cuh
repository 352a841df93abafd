// IR-50 stem on the tensor cores: conv3x3(3 -> 64, pad 1) + folded BN + PReLU, fp32 NCHW in -> bf16 NHWC out.
// Reference: Backbone.input_layer, models/arcface_model.py:130-132.
//
// K = 27 is far too small for an im2col TMA (3 channels x 4 B = 12 B per pixel), and on CUDA cores the layer
// is bound by fp32 FMA issue (1728 FMAs per pixel, 380 us per 2400 frames against 82 us of HBM time).  Here
// the im2col row of a pixel -- 27 taps padded to K = 32 fp32 = exactly one 128-byte swizzle row -- is
// gathered by producer threads straight into the tcgen05 operand layout, and the contraction is four
// kind::tf32 MMAs (M128 x N64 x K8) per 128-pixel tile: 128 tensor cycles per tile.  tf32 keeps 10 mantissa
// bits of the input pixels and weights (inputs and weights are rounded to nearest, not truncated), i.e.
// more than the bf16 every later layer works in.  What is left is the gather, the epilogue and the
// 128 B/pixel output stream: the kernel is HBM-bound (19.2 KB in + 204.8 KB out per frame).
//
// Output: a tile is 128 pixels x 128 B = 16 KB of CONTIGUOUS global memory.  The epilogue writes its bf16 rows
// into a 128-byte-swizzled staging tile in shared memory and one thread hands the tile to the TMA unit
// (cp.async.bulk.tensor store): with one pixel per lane a warp-wide STG.128 would touch 32 different lines
// (16 B in each), and ncu showed the epilogue warps stalled on exactly those stores
// (profiles/r02_stem_tc_v1.txt: 279 us with direct stores).
//
// Warps: 0-7 epilogue (+ bias -> PReLU -> bf16 -> staging -> TMA store), 8 MMA issuer,
// 9-16 two producer groups of 128 threads (one pixel per thread; group g takes this CTA's tiles it = g mod 2).
#pragma once
#include "conv_igemm.cuh"

namespace cer {

constexpr int kStemThreads = 17 * 32;
constexpr int kStemStages = 4;
constexpr int kStemK = 32;             // 27 taps * channels, zero padded

struct StemSmem {
  static constexpr int kBOffset = kStemStages * kABytes;            // A stages: 128 rows x 128 B
  static constexpr int kBBytes = 64 * 128;                          // weights [64][32] fp32
  static constexpr int kBarOffset = kBOffset + kBBytes;
  static constexpr int kNumBars = 2 * kStemStages + 4;
  static constexpr int kTableOffset = (kBarOffset + kNumBars * 8 + 16 + 15) & ~15;
  static constexpr int kOutOffset = (kTableOffset + 2 * 64 * 4 + 1023) & ~1023;      // 2 staging tiles of 128 x 128 B
  static constexpr int kTotal = kOutOffset + 2 * kABytes + 1024;
};

__device__ __forceinline__ float to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// p: Hout/Wout/Cout(=64)/M/bias/alpha/out as for the conv kernels (tensor maps unused); x: fp32 [N][3][H][W];
// w: fp32 [27][64], k = (r*3 + s)*3 + c, BN scale folded.
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void named_barrier(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// tmap_out: the output as a 2-D tensor [pixels][64] bf16, box 64 x 128 rows, 128-byte swizzle.
__global__ void __launch_bounds__(kStemThreads, 1) stem_tc_kernel(const __grid_constant__ ConvKernelParams p,
                                                                  const __grid_constant__ CUtensorMap tmap_out,
                                                                  const float* __restrict__ x, const float* __restrict__ w) {
  using L = StemSmem;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
  uint64_t* empty_bar = full_bar + kStemStages;
  uint64_t* tfull_bar = empty_bar + kStemStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  float* s_bias = reinterpret_cast<float*>(smem + L::kTableOffset);
  float* s_alpha = s_bias + 64;

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.num_m_tiles;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStemStages; ++s) { mbar_init(&full_bar[s], 128); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], kEpiWarps); }
    fence_barrier_init();
  }
  if (warp == kMmaWarp) { tmem_alloc(tmem_slot, 128); tmem_relinquish(); }
  if (threadIdx.x == 0) tma_prefetch_desc(&tmap_out);
  if (warp < kEpiWarps) {
    if (threadIdx.x < 64) { s_bias[threadIdx.x] = __ldg(p.bias + threadIdx.x); s_alpha[threadIdx.x] = __ldg(p.alpha + threadIdx.x); }
  } else if (warp > kMmaWarp) {
    // weights -> B operand [co][k], 128-byte swizzle: 16-byte chunk j of row co lives at chunk j ^ (co & 7)
    float* sB = reinterpret_cast<float*>(smem + L::kBOffset);
    for (int i = threadIdx.x - (kMmaWarp + 1) * 32; i < 64 * kStemK; i += 8 * 32) {
      const int co = i & 63, k = i >> 6;
      const float v = k < 27 ? to_tf32(__ldg(w + k * 64 + co)) : 0.f;
      sB[co * 32 + (((k >> 2) ^ (co & 7)) << 2) + (k & 3)] = v;
    }
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar);
  const uint32_t tfull0 = smem_u32(tfull_bar), tempty0 = smem_u32(tempty_bar);

  if (warp > kMmaWarp) {
    // ===================== producers: one pixel per thread, one im2col row of 32 fp32 =====================
    const int g = (warp - (kMmaWarp + 1)) >> 2;                       // producer group 0 / 1
    const int t = threadIdx.x - (kMmaWarp + 1 + 4 * g) * 32;          // row of the tile
    const int H = p.Hout, W = p.Wout, hw = H * W;
    int it = g;
    for (int tile = blockIdx.x + g * gridDim.x; tile < num_tiles; tile += 2 * gridDim.x, it += 2) {
      const int m = tile * kBlockM + t;
      float v[kStemK];
#pragma unroll
      for (int k = 27; k < kStemK; ++k) v[k] = 0.f;
      {
        // all 27 loads are issued unconditionally from clamped addresses (one batch in flight), padding and
        // the ragged last tile are applied afterwards by selection
        const int mm = m < p.M ? m : p.M - 1;
        const int n = mm / hw;
        const int rem = mm - n * hw;
        const int oh = rem / W, ow = rem - oh * W;
        const float* xn = x + static_cast<size_t>(n) * 3 * hw;
        int roff[3], coff[3];
        bool okr[3], okc[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          const int ih = oh + r - 1, iw = ow + r - 1;
          okr[r] = ih >= 0 && ih < H && m < p.M;
          okc[r] = iw >= 0 && iw < W;
          roff[r] = min(max(ih, 0), H - 1) * W;
          coff[r] = min(max(iw, 0), W - 1);
        }
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int s = 0; s < 3; ++s)
#pragma unroll
            for (int c = 0; c < 3; ++c) v[(r * 3 + s) * 3 + c] = __ldg(xn + c * hw + roff[r] + coff[s]);
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int s = 0; s < 3; ++s)
#pragma unroll
            for (int c = 0; c < 3; ++c) v[(r * 3 + s) * 3 + c] = (okr[r] && okc[s]) ? v[(r * 3 + s) * 3 + c] : 0.f;
      }
      const int stage = it % kStemStages;
      const uint32_t phase = (it / kStemStages) & 1;
      mbar_wait_a(empty0 + stage * 8, phase ^ 1);
      float4* row = reinterpret_cast<float4*>(smem + stage * kABytes + t * 128);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        row[j ^ (t & 7)] = make_float4(to_tf32(v[4 * j]), to_tf32(v[4 * j + 1]), to_tf32(v[4 * j + 2]), to_tf32(v[4 * j + 3]));
      fence_proxy_async();                       // generic-proxy writes -> visible to the tensor core's async proxy
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(full0 + stage * 8) : "memory");
    }
  } else if (warp == kMmaWarp) {
    // ===================== MMA issuer: 4 x (M128 x N64 x K8) tf32 per tile =====================
    constexpr uint32_t idesc = umma_idesc(kBlockM, 64, /*tf32*/ 2);
    const uint32_t a_lo0 = umma_desc_lo(smem_base);
    const uint32_t b_lo = umma_desc_lo(smem_base + L::kBOffset);
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
      const int stage = it % kStemStages;
      const uint32_t phase = (it / kStemStages) & 1;
      mbar_wait_a(tempty0 + acc * 8, acc_phase ^ 1);
      mbar_wait_a(full0 + stage * 8, phase);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t a_lo = a_lo0 + stage * (kABytes >> 4);
        const uint32_t tmem_d = tmem_base + acc * 64;
#pragma unroll
        for (int k = 0; k < 4; ++k)               // 8 tf32 = 32 B along K = +2 in the descriptor address field
          umma_tf32(tmem_d, umma_desc_from_lo(a_lo + 2 * k), umma_desc_from_lo(b_lo + 2 * k), idesc, k != 0 ? 1u : 0u);
        umma_commit_a(empty0 + stage * 8);
        umma_commit_a(tfull0 + acc * 8);
      }
      __syncwarp();
    }
  } else {
    // ===================== epilogue (warps 0-7): + bias -> PReLU -> bf16 -> staging tile -> TMA store =====================
    const int quarter = warp & 3, half = warp >> 2;
    const int row = quarter * 32 + lane;                       // TMEM lane == pixel of the tile
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    const float* sb = s_bias + half * 32;
    const float* sal = s_alpha + half * 32;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
      mbar_wait_a(tfull0 + acc * 8, acc_phase);
      tc_fence_after();
      uint32_t v[32];
      tmem_ld_32x32(tmem_base + acc * 64 + lane_addr + half * 32, v);
      tmem_ld_wait();
      tc_fence_before();
      if (lane == 0) mbar_arrive_relaxed_cluster(tempty0 + acc * 8);
      uint32_t o[16];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 b = *reinterpret_cast<const float4*>(sb + 4 * j);
        const float4 a = *reinterpret_cast<const float4*>(sal + 4 * j);
        float f0 = __uint_as_float(v[4 * j + 0]) + b.x, f1 = __uint_as_float(v[4 * j + 1]) + b.y;
        float f2 = __uint_as_float(v[4 * j + 2]) + b.z, f3 = __uint_as_float(v[4 * j + 3]) + b.w;
        f0 = f0 >= 0.f ? f0 : f0 * a.x; f1 = f1 >= 0.f ? f1 : f1 * a.y;
        f2 = f2 >= 0.f ? f2 : f2 * a.z; f3 = f3 >= 0.f ? f3 : f3 * a.w;
        o[2 * j] = pack_bf16x2(f0, f1);
        o[2 * j + 1] = pack_bf16x2(f2, f3);
      }
      // staging[it & 1] was the source of the store issued two tiles ago
      if (threadIdx.x == 0) tma_store_wait_read<1>();
      named_barrier(1, kEpiWarps * 32);
      uint4* dst = reinterpret_cast<uint4*>(smem + L::kOutOffset + (it & 1) * kABytes + row * 128);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        dst[(half * 4 + j) ^ (row & 7)] = make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
      fence_proxy_async();
      named_barrier(2, kEpiWarps * 32);
      if (threadIdx.x == 0) {                                   // rows past M are clipped by the tensor map
        tma_store_2d(&tmap_out, smem_base + L::kOutOffset + (it & 1) * kABytes, 0, tile * kBlockM);
        tma_store_commit();
      }
    }
    if (threadIdx.x == 0) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 128);
  }
}

}  // namespace cer
