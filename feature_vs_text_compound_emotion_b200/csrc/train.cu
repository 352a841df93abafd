// Fusion-head training step (BASELINE config 4): forward in training mode, mean cross-entropy,
// backward through TCN x M / BatchNorm1d / cross-modal attention / LayerNorm / classifier, and the
// fused SGD / Adam / AdamW update.  C-ABI: cer_head_train_*, cer_ce_loss, cer_optimizer_step
// (include/cer_b200.h).
//
// Reference semantics: trainer.py:365-391 (step), experiment.py:133 (CrossEntropyLoss mean),
// models/model.py:511-526 under model.train(): Dropout after both LeakyReLUs of every
// TemporalBlock (temporal_convolutional_model.py:28,34) and on the attention output
// (transformer.py:194), BatchNorm1d with batch statistics + running-stat update (model.py:475,515),
// weight_norm re-parameterisation w = g v/||v|| (temporal_convolutional_model.py:24,30).
//
// Everything here is exact fp32 (CUDA cores): the step is 30 MFLOP/frame, three orders of
// magnitude below the IR-50 pass, and gradient parity with the reference at 1e-4 matters more than
// tensor-core throughput.  All activations are time-major rows [R = B*T][C].
//
// The dropout masks come from a counter hash (dropout_keep) that oracle/lfan_oracle.py restates,
// so forward AND backward are checkable element by element with dropout on.
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdint>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/cer_b200.h"
#include "common.h"
#include "row_gemm.h"

namespace cer {

constexpr float kLeaky = 0.01f;
constexpr float kBnEps = 1e-5f;
constexpr float kLnEps = 1e-5f;

__device__ __forceinline__ float drop_factor(const Drop& d, uint32_t idx) {
  if (d.thr == 0) return 1.f;
  return fmix32(idx * 0x9E3779B1u + d.key) >= d.thr ? d.scale : 0.f;
}
__device__ __forceinline__ float lrelu(float x) { return x > 0.f ? x : x * kLeaky; }
__device__ __forceinline__ float lrelu_grad(float saved) { return saved > 0.f ? 1.f : kLeaky; }

// ------------------------------------------------------------------------------------------
// Row GEMM with taps:  C[r, n] (+)= sum_j sum_k A[r + shift_j, k] * B_j(k, n)   (+ epilogue)
//   rows are grouped in windows of T (one per batch element); a shifted row outside its window
//   reads as zero (causal padding forward, anti-causal in dgrad).
//   B_KN = false: B_j is [N][K] row-major (torch Linear / conv weight: forward);
//   B_KN = true : B_j is [K][N] row-major (the same storage read transposed: dgrad).
// 64x64 tile, BK = 16, 256 threads, 4x4 outputs per thread.
// ------------------------------------------------------------------------------------------
// ---- shared micro-kernel: 128 x 64 outputs per CTA, 256 threads, 8 x 4 per thread, packed FP32 FMAs ----
// Blackwell issues two fp32 FMAs per instruction (fma.rn.f32x2).  The "m" operand comes from shared
// memory as natural pairs (rows 2i, 2i+1); the "n" operand is stored DUPLICATED ({b, b}) so that no
// register shuffling is needed: per k, 4 x LDS.128 feed 16 x FFMA2 (= 32 FMAs).
__device__ __forceinline__ void ffma2(unsigned long long& d, unsigned long long a, unsigned long long b) {
  asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
}
struct alignas(16) Pair2 { unsigned long long x, y; };      // two packed float2 = one LDS.128

constexpr int kGM = 128, kGN = 64, kGK = 16;

template <int BM>
__device__ __forceinline__ void micro_step(const float (*Ms)[BM + 4], const float2 (*Ns)[kGN], int kk, int ty, int tx,
                                           unsigned long long (&acc)[4][4]) {
  const Pair2 a01 = *reinterpret_cast<const Pair2*>(&Ms[kk][ty * 8]);          // rows (0,1) (2,3)
  const Pair2 a23 = *reinterpret_cast<const Pair2*>(&Ms[kk][ty * 8 + 4]);      // rows (4,5) (6,7)
  const Pair2 b01 = *reinterpret_cast<const Pair2*>(&Ns[kk][tx * 4]);          // {b0,b0} {b1,b1}
  const Pair2 b23 = *reinterpret_cast<const Pair2*>(&Ns[kk][tx * 4 + 2]);      // {b2,b2} {b3,b3}
  const unsigned long long a[4] = {a01.x, a01.y, a23.x, a23.y};
  const unsigned long long b[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) ffma2(acc[i][j], a[i], b[j]);
}
__device__ __forceinline__ float2 unpack2(unsigned long long v) {
  float2 r;
  asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
  return r;
}

// BM = 128 (256 threads) or 64 (128 threads): the smaller tile doubles the CTA count when the
// problem has too few tiles to fill 148 SMs.
template <bool B_KN, int BM>
__global__ void __launch_bounds__(2 * BM) row_gemm_kernel(const RowGemm g) {
  constexpr int NT = 2 * BM;                             // threads
  __shared__ __align__(16) float As[kGK][BM + 4];        // [k][row]
  __shared__ __align__(16) float2 Bs[kGK][kGN];          // [k][col] duplicated
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;                // thread -> rows ty*8.., cols tx*4..
  const int row0 = blockIdx.y * BM, col0 = blockIdx.x * kGN;
  unsigned long long acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0ull;

  // loader mapping.  A tile: BM rows x 16 k, two float4 per thread (rows a_r and a_r + BM/2).
  constexpr int AH = BM / 2;
  const int a_r = tid >> 2, a_k = (tid & 3) * 4;
  const bool a_vec = (g.lda & 3) == 0 && (g.K & 3) == 0 && (reinterpret_cast<uintptr_t>(g.A) & 15) == 0;
  int a_t[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) { const int r = row0 + a_r + AH * h; a_t[h] = r < g.R ? r % g.T : -(1 << 30); }
  // B tile: 16 k x 64 n = 256 float4; threads beyond 256 float4 slots (none) / fewer threads loop twice (BM = 64).
  // [K][N] storage: slot -> one k, 4 consecutive n;  [N][K] storage: one n, 4 consecutive k.
  constexpr int BP = 256 / NT;                           // B float4 slots per thread (1 or 2)
  const int b_ld = g.ldb ? g.ldb : (B_KN ? g.N : g.K);
  const bool b_vec = (b_ld & 3) == 0 && (reinterpret_cast<uintptr_t>(g.B) & 15) == 0 && (g.b_tap_stride & 3) == 0;

  const int ksteps = (g.K + kGK - 1) / kGK;
  const int total = g.taps * ksteps;
  float4 ra[2], rbv[BP];
  auto fetch = [&](int it) {
    const int j = it / ksteps, k0 = (it - j * ksteps) * kGK;
    const int shift = g.shift0 + j * g.shift_step;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = row0 + a_r + AH * h;
      const bool ok = (a_t[h] + shift) >= 0 && (a_t[h] + shift) < g.T;      // a_t < 0 for rows >= R
      const float* ap = g.A + (long long)(r + shift) * g.lda + k0 + a_k;
      if (ok && a_vec && k0 + a_k + 3 < g.K) ra[h] = __ldg(reinterpret_cast<const float4*>(ap));
      else {
        ra[h].x = (ok && k0 + a_k + 0 < g.K) ? __ldg(ap + 0) : 0.f;
        ra[h].y = (ok && k0 + a_k + 1 < g.K) ? __ldg(ap + 1) : 0.f;
        ra[h].z = (ok && k0 + a_k + 2 < g.K) ? __ldg(ap + 2) : 0.f;
        ra[h].w = (ok && k0 + a_k + 3 < g.K) ? __ldg(ap + 3) : 0.f;
      }
    }
    const float* bp = g.B + j * g.b_tap_stride;
#pragma unroll
    for (int u = 0; u < BP; ++u) {
      const int slot = tid + u * NT;
      const int b_a = B_KN ? (slot >> 4) : (slot >> 2);        // k (B_KN) or n (!B_KN)
      const int b_b = B_KN ? (slot & 15) * 4 : (slot & 3) * 4; // n (B_KN) or k (!B_KN)
      float4& rb = rbv[u];
      if (B_KN) {
        const int k = k0 + b_a, n = col0 + b_b;
        const float* q = bp + (long long)k * (g.ldb ? g.ldb : g.N) + n;
        if (k < g.K && b_vec && n + 3 < g.N) rb = __ldg(reinterpret_cast<const float4*>(q));
        else {
          rb.x = (k < g.K && n + 0 < g.N) ? __ldg(q + 0) : 0.f;
          rb.y = (k < g.K && n + 1 < g.N) ? __ldg(q + 1) : 0.f;
          rb.z = (k < g.K && n + 2 < g.N) ? __ldg(q + 2) : 0.f;
          rb.w = (k < g.K && n + 3 < g.N) ? __ldg(q + 3) : 0.f;
        }
      } else {
        const int n = col0 + b_a, k = k0 + b_b;
        const float* q = bp + (long long)n * (g.ldb ? g.ldb : g.K) + k;
        if (n < g.N && b_vec && k + 3 < g.K) rb = __ldg(reinterpret_cast<const float4*>(q));
        else {
          rb.x = (n < g.N && k + 0 < g.K) ? __ldg(q + 0) : 0.f;
          rb.y = (n < g.N && k + 1 < g.K) ? __ldg(q + 1) : 0.f;
          rb.z = (n < g.N && k + 2 < g.K) ? __ldg(q + 2) : 0.f;
          rb.w = (n < g.N && k + 3 < g.K) ? __ldg(q + 3) : 0.f;
        }
      }
    }
  };
  auto stash = [&]() {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      As[a_k + 0][a_r + AH * h] = ra[h].x; As[a_k + 1][a_r + AH * h] = ra[h].y;
      As[a_k + 2][a_r + AH * h] = ra[h].z; As[a_k + 3][a_r + AH * h] = ra[h].w;
    }
#pragma unroll
    for (int u = 0; u < BP; ++u) {
      const int slot = tid + u * NT;
      const int b_a = B_KN ? (slot >> 4) : (slot >> 2);
      const int b_b = B_KN ? (slot & 15) * 4 : (slot & 3) * 4;
      const float4 rb = rbv[u];
      if (B_KN) {
        Bs[b_a][b_b + 0] = make_float2(rb.x, rb.x); Bs[b_a][b_b + 1] = make_float2(rb.y, rb.y);
        Bs[b_a][b_b + 2] = make_float2(rb.z, rb.z); Bs[b_a][b_b + 3] = make_float2(rb.w, rb.w);
      } else {
        Bs[b_b + 0][b_a] = make_float2(rb.x, rb.x); Bs[b_b + 1][b_a] = make_float2(rb.y, rb.y);
        Bs[b_b + 2][b_a] = make_float2(rb.z, rb.z); Bs[b_b + 3][b_a] = make_float2(rb.w, rb.w);
      }
    }
  };
  fetch(0);
  for (int it = 0; it < total; ++it) {
    stash();
    __syncthreads();
    if (it + 1 < total) fetch(it + 1);           // global loads of the next tile fly during the FMAs
#pragma unroll
    for (int kk = 0; kk < kGK; ++kk) micro_step<BM>(As, Bs, kk, ty, tx, acc);
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int r = row0 + ty * 8 + 2 * i + half;
      if (r >= g.R) continue;
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const int n = col0 + tx * 4 + jj;
        if (n >= g.N) continue;
        const float2 pr = unpack2(acc[i][jj]);
        float v = half ? pr.y : pr.x;
        if (g.bias) v += __ldg(g.bias + n);
        const uint32_t idx = (uint32_t)r * (uint32_t)g.N + (uint32_t)n;
        float* cp = g.C + (long long)r * g.ldc + n;
        if (g.epi == EPI_LINEAR) {
          if (g.addend) v += g.addend[(long long)r * g.ld_add + n];
          if (g.accumulate) v += *cp;
          *cp = v;
        } else if (g.epi == EPI_RELU) {
          *cp = fmaxf(v, 0.f);
        } else if (g.epi == EPI_LRELU_DROP) {
          *cp = lrelu(v) * drop_factor(g.drop, idx);
        } else if (g.epi == EPI_BLOCK_OUT) {
          const float h2d = lrelu(v) * drop_factor(g.drop, idx);
          g.aux[(long long)r * g.ld_aux + n] = h2d;
          *cp = lrelu(h2d + g.addend[(long long)r * g.ld_add + n]);
        } else {   // EPI_DGRAD_ACT: gradient w.r.t. the pre-activation of a LReLU+dropout site
          if (g.addend) v += g.addend[(long long)r * g.ld_add + n];
          const float saved = g.aux[(long long)r * g.ld_aux + n];
          *cp = v * drop_factor(g.drop, idx) * lrelu_grad(saved);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// TF32 tensor-core variants (precision = "tf32", the default of the training plan).  Same tiles, same
// loaders (register prefetch of the next tile), same epilogues; the inner product runs on warp-level
// mma.sync.m16n8k8 TF32 MMAs with fp32 accumulate: each warp owns 32 x 32 outputs = 2 x 4 MMA tiles.
// Operands are rounded to TF32 (cvt.rna) when they are written to shared memory; tiles are stored
// [k][m] / [k][n] with row pitches of BM + 8 / 72 floats, which makes every fragment load
// (address = (k0 + t) * pitch + m0 + g) bank-conflict free.  The step is ~30 MFLOP per frame at 600-4800
// rows: latency- and launch-bound, so the descriptor-free warp-level MMA is used here rather than tcgen05;
// it removes the fp32 FMA floor (20 TFLOP/s) without a second set of epilogues.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t to_tf32_bits(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void mma_16x8x8_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
constexpr int kTcLdn = kGN + 8;      // 72

// one 16-deep k tile: Ms [16][LDM] (m contiguous), Ns [16][72]; this warp's 32 x 32 block at (m0, n0)
template <int LDM>
__device__ __forceinline__ void mma_tile_step(const uint32_t* Ms, const uint32_t* Ns, int m0, int n0, int g, int t,
                                              float (&acc)[2][4][4]) {
#pragma unroll
  for (int k8 = 0; k8 < kGK; k8 += 8) {
    uint32_t a[2][4], b[4][2];
#pragma unroll
    for (int mi = 0; mi < 2; ++mi) {
      const uint32_t* p = Ms + (k8 + t) * LDM + m0 + mi * 16 + g;
      a[mi][0] = p[0]; a[mi][1] = p[8]; a[mi][2] = p[4 * LDM]; a[mi][3] = p[4 * LDM + 8];
    }
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) {
      const uint32_t* p = Ns + (k8 + t) * kTcLdn + n0 + ni * 8 + g;
      b[ni][0] = p[0]; b[ni][1] = p[4 * kTcLdn];
    }
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) mma_16x8x8_tf32(acc[mi][ni], a[mi], b[ni]);
  }
}

template <bool B_KN, int BM>
__global__ void __launch_bounds__(2 * BM) row_gemm_tf32_kernel(const RowGemm g) {
  constexpr int NT = 2 * BM;
  constexpr int LDM = BM + 8;
  __shared__ __align__(16) uint32_t As[kGK * LDM];        // [k][row]
  __shared__ __align__(16) uint32_t Bs[kGK * kTcLdn];     // [k][col]
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int gq = lane >> 2, tq = lane & 3;
  const int wm = (warp >> 1) * 32, wn = (warp & 1) * 32;
  const int row0 = blockIdx.y * BM, col0 = blockIdx.x * kGN;
  float acc[2][4][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { acc[i][j][0] = acc[i][j][1] = acc[i][j][2] = acc[i][j][3] = 0.f; }

  constexpr int AH = BM / 2;
  const int a_r = tid >> 2, a_k = (tid & 3) * 4;
  const bool a_vec = (g.lda & 3) == 0 && (g.K & 3) == 0 && (reinterpret_cast<uintptr_t>(g.A) & 15) == 0;
  int a_t[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) { const int r = row0 + a_r + AH * h; a_t[h] = r < g.R ? r % g.T : -(1 << 30); }
  constexpr int BP = 256 / NT;
  const int b_ld = g.ldb ? g.ldb : (B_KN ? g.N : g.K);
  const bool b_vec = (b_ld & 3) == 0 && (reinterpret_cast<uintptr_t>(g.B) & 15) == 0 && (g.b_tap_stride & 3) == 0;
  const int ksteps = (g.K + kGK - 1) / kGK;
  const int total = g.taps * ksteps;
  float4 ra[2], rbv[BP];
  auto fetch = [&](int it) {
    const int j = it / ksteps, k0 = (it - j * ksteps) * kGK;
    const int shift = g.shift0 + j * g.shift_step;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = row0 + a_r + AH * h;
      const bool ok = (a_t[h] + shift) >= 0 && (a_t[h] + shift) < g.T;
      const float* ap = g.A + (long long)(r + shift) * g.lda + k0 + a_k;
      if (ok && a_vec && k0 + a_k + 3 < g.K) ra[h] = __ldg(reinterpret_cast<const float4*>(ap));
      else {
        ra[h].x = (ok && k0 + a_k + 0 < g.K) ? __ldg(ap + 0) : 0.f;
        ra[h].y = (ok && k0 + a_k + 1 < g.K) ? __ldg(ap + 1) : 0.f;
        ra[h].z = (ok && k0 + a_k + 2 < g.K) ? __ldg(ap + 2) : 0.f;
        ra[h].w = (ok && k0 + a_k + 3 < g.K) ? __ldg(ap + 3) : 0.f;
      }
    }
    const float* bp = g.B + j * g.b_tap_stride;
#pragma unroll
    for (int u = 0; u < BP; ++u) {
      const int slot = tid + u * NT;
      const int b_a = B_KN ? (slot >> 4) : (slot >> 2);
      const int b_b = B_KN ? (slot & 15) * 4 : (slot & 3) * 4;
      float4& rb = rbv[u];
      if (B_KN) {
        const int k = k0 + b_a, n = col0 + b_b;
        const float* q = bp + (long long)k * (g.ldb ? g.ldb : g.N) + n;
        if (k < g.K && b_vec && n + 3 < g.N) rb = __ldg(reinterpret_cast<const float4*>(q));
        else {
          rb.x = (k < g.K && n + 0 < g.N) ? __ldg(q + 0) : 0.f;
          rb.y = (k < g.K && n + 1 < g.N) ? __ldg(q + 1) : 0.f;
          rb.z = (k < g.K && n + 2 < g.N) ? __ldg(q + 2) : 0.f;
          rb.w = (k < g.K && n + 3 < g.N) ? __ldg(q + 3) : 0.f;
        }
      } else {
        const int n = col0 + b_a, k = k0 + b_b;
        const float* q = bp + (long long)n * (g.ldb ? g.ldb : g.K) + k;
        if (n < g.N && b_vec && k + 3 < g.K) rb = __ldg(reinterpret_cast<const float4*>(q));
        else {
          rb.x = (n < g.N && k + 0 < g.K) ? __ldg(q + 0) : 0.f;
          rb.y = (n < g.N && k + 1 < g.K) ? __ldg(q + 1) : 0.f;
          rb.z = (n < g.N && k + 2 < g.K) ? __ldg(q + 2) : 0.f;
          rb.w = (n < g.N && k + 3 < g.K) ? __ldg(q + 3) : 0.f;
        }
      }
    }
  };
  auto stash = [&]() {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      uint32_t* p = As + a_k * LDM + a_r + AH * h;
      p[0] = to_tf32_bits(ra[h].x); p[LDM] = to_tf32_bits(ra[h].y); p[2 * LDM] = to_tf32_bits(ra[h].z); p[3 * LDM] = to_tf32_bits(ra[h].w);
    }
#pragma unroll
    for (int u = 0; u < BP; ++u) {
      const int slot = tid + u * NT;
      const int b_a = B_KN ? (slot >> 4) : (slot >> 2);
      const int b_b = B_KN ? (slot & 15) * 4 : (slot & 3) * 4;
      const float4 rb = rbv[u];
      if (B_KN) {
        *reinterpret_cast<uint4*>(Bs + b_a * kTcLdn + b_b) = make_uint4(to_tf32_bits(rb.x), to_tf32_bits(rb.y), to_tf32_bits(rb.z), to_tf32_bits(rb.w));
      } else {
        uint32_t* p = Bs + b_b * kTcLdn + b_a;
        p[0] = to_tf32_bits(rb.x); p[kTcLdn] = to_tf32_bits(rb.y); p[2 * kTcLdn] = to_tf32_bits(rb.z); p[3 * kTcLdn] = to_tf32_bits(rb.w);
      }
    }
  };
  fetch(0);
  for (int it = 0; it < total; ++it) {
    stash();
    __syncthreads();
    if (it + 1 < total) fetch(it + 1);
    mma_tile_step<LDM>(As, Bs, wm, wn, gq, tq, acc);
    __syncthreads();
  }
#pragma unroll
  for (int mi = 0; mi < 2; ++mi) {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int r = row0 + wm + mi * 16 + gq + 8 * half;
      if (r >= g.R) continue;
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int n = col0 + wn + ni * 8 + 2 * tq + c;
          if (n >= g.N) continue;
          float v = acc[mi][ni][2 * half + c];
          if (g.bias) v += __ldg(g.bias + n);
          const uint32_t idx = (uint32_t)r * (uint32_t)g.N + (uint32_t)n;
          float* cp = g.C + (long long)r * g.ldc + n;
          if (g.epi == EPI_LINEAR) {
            if (g.addend) v += g.addend[(long long)r * g.ld_add + n];
            if (g.accumulate) v += *cp;
            *cp = v;
          } else if (g.epi == EPI_RELU) {
            *cp = fmaxf(v, 0.f);
          } else if (g.epi == EPI_LRELU_DROP) {
            *cp = lrelu(v) * drop_factor(g.drop, idx);
          } else if (g.epi == EPI_BLOCK_OUT) {
            const float h2d = lrelu(v) * drop_factor(g.drop, idx);
            g.aux[(long long)r * g.ld_aux + n] = h2d;
            *cp = lrelu(h2d + g.addend[(long long)r * g.ld_add + n]);
          } else {   // EPI_DGRAD_ACT
            if (g.addend) v += g.addend[(long long)r * g.ld_add + n];
            const float saved = g.aux[(long long)r * g.ld_aux + n];
            *cp = v * drop_factor(g.drop, idx) * lrelu_grad(saved);
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// Intra-CTA split-K variant of the row GEMM for SMALL grids.  The head's GEMMs have only 75-300
// 64x64 output tiles: with one 4-warp CTA per tile an SM holds 4-8 warps and every LDS->FFMA2
// dependency is exposed (ncu: short_scoreboard stalls, sm active 65 %).  Here a CTA is four groups
// of 128 threads; group g runs the same 64x64 micro-kernel over the (tap, k-tile) iterations
// it = g (mod 4) on its own shared-memory tiles and its own named barrier, the four partial
// accumulators are summed through shared memory, and group 0 runs the epilogue: 16 warps per SM
// with the same register blocking.  Summation order differs from the 1-group kernel only in the
// final 4-way add.
// ------------------------------------------------------------------------------------------
constexpr int kSplitGroupBytes = kGK * (64 + 4) * 4 + kGK * kGN * 8;                 // 4352 + 8192 per K-group

__device__ __forceinline__ void group_sync(int g) { asm volatile("bar.sync %0, 128;" ::"r"(g + 1) : "memory"); }

template <bool B_KN, int kSplitG>
__global__ void __launch_bounds__(128 * kSplitG) row_gemm_splitk_kernel(const RowGemm g) {
  constexpr int BM = 64, NT = 128;
  extern __shared__ __align__(16) uint8_t sk_smem[];
  const int grp = threadIdx.x >> 7;
  const int tid = threadIdx.x & 127;
  float (*As)[BM + 4] = reinterpret_cast<float (*)[BM + 4]>(sk_smem + grp * (kGK * (BM + 4) * 4 + kGK * kGN * 8));
  float2 (*Bs)[kGN] = reinterpret_cast<float2 (*)[kGN]>(reinterpret_cast<uint8_t*>(As) + kGK * (BM + 4) * 4);
  const int tx = tid & 15, ty = tid >> 4;
  const int row0 = blockIdx.y * BM, col0 = blockIdx.x * kGN;
  unsigned long long acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0ull;

  constexpr int AH = BM / 2;
  const int a_r = tid >> 2, a_k = (tid & 3) * 4;
  const bool a_vec = (g.lda & 3) == 0 && (g.K & 3) == 0 && (reinterpret_cast<uintptr_t>(g.A) & 15) == 0;
  int a_t[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) { const int r = row0 + a_r + AH * h; a_t[h] = r < g.R ? r % g.T : -(1 << 30); }
  constexpr int BP = 256 / NT;
  const int b_ld = g.ldb ? g.ldb : (B_KN ? g.N : g.K);
  const bool b_vec = (b_ld & 3) == 0 && (reinterpret_cast<uintptr_t>(g.B) & 15) == 0 && (g.b_tap_stride & 3) == 0;
  const int ksteps = (g.K + kGK - 1) / kGK;
  const int total = g.taps * ksteps;
  float4 ra[2], rbv[BP];
  auto fetch = [&](int it) {
    const int j = it / ksteps, k0 = (it - j * ksteps) * kGK;
    const int shift = g.shift0 + j * g.shift_step;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = row0 + a_r + AH * h;
      const bool ok = (a_t[h] + shift) >= 0 && (a_t[h] + shift) < g.T;
      const float* ap = g.A + (long long)(r + shift) * g.lda + k0 + a_k;
      if (ok && a_vec && k0 + a_k + 3 < g.K) ra[h] = __ldg(reinterpret_cast<const float4*>(ap));
      else {
        ra[h].x = (ok && k0 + a_k + 0 < g.K) ? __ldg(ap + 0) : 0.f;
        ra[h].y = (ok && k0 + a_k + 1 < g.K) ? __ldg(ap + 1) : 0.f;
        ra[h].z = (ok && k0 + a_k + 2 < g.K) ? __ldg(ap + 2) : 0.f;
        ra[h].w = (ok && k0 + a_k + 3 < g.K) ? __ldg(ap + 3) : 0.f;
      }
    }
    const float* bp = g.B + j * g.b_tap_stride;
#pragma unroll
    for (int u = 0; u < BP; ++u) {
      const int slot = tid + u * NT;
      const int b_a = B_KN ? (slot >> 4) : (slot >> 2);
      const int b_b = B_KN ? (slot & 15) * 4 : (slot & 3) * 4;
      float4& rb = rbv[u];
      if (B_KN) {
        const int k = k0 + b_a, n = col0 + b_b;
        const float* q = bp + (long long)k * (g.ldb ? g.ldb : g.N) + n;
        if (k < g.K && b_vec && n + 3 < g.N) rb = __ldg(reinterpret_cast<const float4*>(q));
        else {
          rb.x = (k < g.K && n + 0 < g.N) ? __ldg(q + 0) : 0.f;
          rb.y = (k < g.K && n + 1 < g.N) ? __ldg(q + 1) : 0.f;
          rb.z = (k < g.K && n + 2 < g.N) ? __ldg(q + 2) : 0.f;
          rb.w = (k < g.K && n + 3 < g.N) ? __ldg(q + 3) : 0.f;
        }
      } else {
        const int n = col0 + b_a, k = k0 + b_b;
        const float* q = bp + (long long)n * (g.ldb ? g.ldb : g.K) + k;
        if (n < g.N && b_vec && k + 3 < g.K) rb = __ldg(reinterpret_cast<const float4*>(q));
        else {
          rb.x = (n < g.N && k + 0 < g.K) ? __ldg(q + 0) : 0.f;
          rb.y = (n < g.N && k + 1 < g.K) ? __ldg(q + 1) : 0.f;
          rb.z = (n < g.N && k + 2 < g.K) ? __ldg(q + 2) : 0.f;
          rb.w = (n < g.N && k + 3 < g.K) ? __ldg(q + 3) : 0.f;
        }
      }
    }
  };
  auto stash = [&]() {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      As[a_k + 0][a_r + AH * h] = ra[h].x; As[a_k + 1][a_r + AH * h] = ra[h].y;
      As[a_k + 2][a_r + AH * h] = ra[h].z; As[a_k + 3][a_r + AH * h] = ra[h].w;
    }
#pragma unroll
    for (int u = 0; u < BP; ++u) {
      const int slot = tid + u * NT;
      const int b_a = B_KN ? (slot >> 4) : (slot >> 2);
      const int b_b = B_KN ? (slot & 15) * 4 : (slot & 3) * 4;
      const float4 rb = rbv[u];
      if (B_KN) {
        Bs[b_a][b_b + 0] = make_float2(rb.x, rb.x); Bs[b_a][b_b + 1] = make_float2(rb.y, rb.y);
        Bs[b_a][b_b + 2] = make_float2(rb.z, rb.z); Bs[b_a][b_b + 3] = make_float2(rb.w, rb.w);
      } else {
        Bs[b_b + 0][b_a] = make_float2(rb.x, rb.x); Bs[b_b + 1][b_a] = make_float2(rb.y, rb.y);
        Bs[b_b + 2][b_a] = make_float2(rb.z, rb.z); Bs[b_b + 3][b_a] = make_float2(rb.w, rb.w);
      }
    }
  };
  if (grp < total) fetch(grp);
  for (int it = grp; it < total; it += kSplitG) {
    stash();
    group_sync(grp);
    if (it + kSplitG < total) fetch(it + kSplitG);
#pragma unroll
    for (int kk = 0; kk < kGK; ++kk) micro_step<BM>(As, Bs, kk, ty, tx, acc);
    group_sync(grp);
  }
  // ---- sum the four partial tiles through shared memory (the tiles are dead now)
  __syncthreads();
  float* red = reinterpret_cast<float*>(sk_smem);                 // [kSplitG - 1][128 threads][32] = 48 KB
  if (grp > 0) {
    float* dst = red + ((grp - 1) * 128 + tid) * 32;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const float2 pr = unpack2(acc[i][jj]);
        dst[(i * 4 + jj) * 2 + 0] = pr.x;
        dst[(i * 4 + jj) * 2 + 1] = pr.y;
      }
  }
  __syncthreads();
  if (grp != 0) return;
  float vals[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const float2 pr = unpack2(acc[i][jj]);
      float lo = pr.x, hi = pr.y;
#pragma unroll
      for (int q = 0; q < kSplitG - 1; ++q) {
        const float* src = red + (q * 128 + tid) * 32 + (i * 4 + jj) * 2;
        lo += src[0]; hi += src[1];
      }
      vals[i][jj][0] = lo; vals[i][jj][1] = hi;
    }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int r = row0 + ty * 8 + 2 * i + half;
      if (r >= g.R) continue;
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const int n = col0 + tx * 4 + jj;
        if (n >= g.N) continue;
        float v = vals[i][jj][half];
        if (g.bias) v += __ldg(g.bias + n);
        const uint32_t idx = (uint32_t)r * (uint32_t)g.N + (uint32_t)n;
        float* cp = g.C + (long long)r * g.ldc + n;
        if (g.epi == EPI_LINEAR) {
          if (g.addend) v += g.addend[(long long)r * g.ld_add + n];
          if (g.accumulate) v += *cp;
          *cp = v;
        } else if (g.epi == EPI_RELU) {
          *cp = fmaxf(v, 0.f);
        } else if (g.epi == EPI_LRELU_DROP) {
          *cp = lrelu(v) * drop_factor(g.drop, idx);
        } else if (g.epi == EPI_BLOCK_OUT) {
          const float h2d = lrelu(v) * drop_factor(g.drop, idx);
          g.aux[(long long)r * g.ld_aux + n] = h2d;
          *cp = lrelu(h2d + g.addend[(long long)r * g.ld_add + n]);
        } else {
          if (g.addend) v += g.addend[(long long)r * g.ld_add + n];
          const float saved = g.aux[(long long)r * g.ld_aux + n];
          *cp = v * drop_factor(g.drop, idx) * lrelu_grad(saved);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// Weight gradient with taps:  dW_j[n][k] += sum_r G[r, n] * A[r + shift_j, k]
// CTA tile: 128 n x 64 k for one tap, a slice of the rows; grid = (ceil(K/64) * taps, ceil(N/128),
// row splits); fp32 atomicAdd into a zeroed buffer.  Same micro-kernel (m = n of G, n = k of A).
// ------------------------------------------------------------------------------------------
struct WGrad {
  const float* G; int ldg;
  const float* A; int lda;
  float* dW; long long w_tap_stride;
  int R, T, N, K;
  int taps, shift0, shift_step;
  int k_tiles;
  int rows_per_split;
  int tf32;                                 // 1: TF32 tensor-core kernel
  int ldw;                                  // row pitch of dW (0 = dense, K)
};

__global__ void __launch_bounds__(256) wgrad_kernel(const WGrad g) {
  __shared__ __align__(16) float Gs[kGK][kGM + 4];       // [row][n]
  __shared__ __align__(16) float2 As[kGK][kGN];          // [row][k] duplicated
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;                // thread -> n = ty*8.., k = tx*4..
  const int j = blockIdx.x / g.k_tiles;
  const int k0 = (blockIdx.x - j * g.k_tiles) * kGN;
  const int n0 = blockIdx.y * kGM;
  const int shift = g.shift0 + j * g.shift_step;
  const int r_begin = blockIdx.z * g.rows_per_split;
  const int r_end = min(g.R, r_begin + g.rows_per_split);
  unsigned long long acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[i][q] = 0ull;
  // loaders: G tile 16 rows x 128 n -> 2 float4 per thread; A tile 16 rows x 64 k -> 1 float4 per thread
  const int g_r = tid >> 4, g_c = (tid & 15) * 4;        // + 64 for the second half
  const int a_r = tid >> 4, a_c = (tid & 15) * 4;
  const bool g_vec = (g.ldg & 3) == 0 && (reinterpret_cast<uintptr_t>(g.G) & 15) == 0;
  const bool a_vec = (g.lda & 3) == 0 && (reinterpret_cast<uintptr_t>(g.A) & 15) == 0;
  float4 rg[2], ra;
  auto fetch = [&](int r0) {
    const int r = r0 + g_r;
    const bool r_ok = r < r_end;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int n = n0 + g_c + 64 * h;
      const float* q = g.G + (long long)r * g.ldg + n;
      if (r_ok && g_vec && n + 3 < g.N) rg[h] = __ldg(reinterpret_cast<const float4*>(q));
      else {
        rg[h].x = (r_ok && n + 0 < g.N) ? __ldg(q + 0) : 0.f;
        rg[h].y = (r_ok && n + 1 < g.N) ? __ldg(q + 1) : 0.f;
        rg[h].z = (r_ok && n + 2 < g.N) ? __ldg(q + 2) : 0.f;
        rg[h].w = (r_ok && n + 3 < g.N) ? __ldg(q + 3) : 0.f;
      }
    }
    const int t = r_ok ? r % g.T : -(1 << 30);
    const bool ok = (t + shift) >= 0 && (t + shift) < g.T;
    const int k = k0 + a_c;
    const float* q = g.A + (long long)(r + shift) * g.lda + k;
    if (ok && a_vec && k + 3 < g.K) ra = __ldg(reinterpret_cast<const float4*>(q));
    else {
      ra.x = (ok && k + 0 < g.K) ? __ldg(q + 0) : 0.f;
      ra.y = (ok && k + 1 < g.K) ? __ldg(q + 1) : 0.f;
      ra.z = (ok && k + 2 < g.K) ? __ldg(q + 2) : 0.f;
      ra.w = (ok && k + 3 < g.K) ? __ldg(q + 3) : 0.f;
    }
  };
  if (r_begin < r_end) fetch(r_begin);
  for (int r0 = r_begin; r0 < r_end; r0 += kGK) {
#pragma unroll
    for (int h = 0; h < 2; ++h) *reinterpret_cast<float4*>(&Gs[g_r][g_c + 64 * h]) = rg[h];
    As[a_r][a_c + 0] = make_float2(ra.x, ra.x); As[a_r][a_c + 1] = make_float2(ra.y, ra.y);
    As[a_r][a_c + 2] = make_float2(ra.z, ra.z); As[a_r][a_c + 3] = make_float2(ra.w, ra.w);
    __syncthreads();
    if (r0 + kGK < r_end) fetch(r0 + kGK);
#pragma unroll
    for (int rr = 0; rr < kGK; ++rr) micro_step<kGM>(Gs, As, rr, ty, tx, acc);
    __syncthreads();
  }
  float* wp = g.dW + j * g.w_tap_stride;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int n = n0 + ty * 8 + 2 * i + half;
      if (n >= g.N) continue;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int k = k0 + tx * 4 + q;
        const float2 pr = unpack2(acc[i][q]);
        if (k < g.K) atomicAdd(wp + (long long)n * (g.ldw ? g.ldw : g.K) + k, half ? pr.y : pr.x);
      }
    }
  }
}

// TF32 tensor-core weight gradient: same tiling (128 n x 64 k per tap, a slice of the rows, fp32 atomics);
// the reduction dimension (rows) is the MMA's k.  Gs [16 rows][136] is the "m" operand, As [16 rows][72] the "n".
__global__ void __launch_bounds__(256) wgrad_tf32_kernel(const WGrad g) {
  constexpr int LDM = kGM + 8;
  __shared__ __align__(16) uint32_t Gs[kGK * LDM];
  __shared__ __align__(16) uint32_t As[kGK * kTcLdn];
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int gq = lane >> 2, tq = lane & 3;
  const int wm = (warp >> 1) * 32, wn = (warp & 1) * 32;
  const int j = blockIdx.x / g.k_tiles;
  const int k0 = (blockIdx.x - j * g.k_tiles) * kGN;
  const int n0 = blockIdx.y * kGM;
  const int shift = g.shift0 + j * g.shift_step;
  const int r_begin = blockIdx.z * g.rows_per_split;
  const int r_end = min(g.R, r_begin + g.rows_per_split);
  float acc[2][4][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int q = 0; q < 4; ++q) { acc[i][q][0] = acc[i][q][1] = acc[i][q][2] = acc[i][q][3] = 0.f; }
  const int g_r = tid >> 4, g_c = (tid & 15) * 4;
  const int a_r = tid >> 4, a_c = (tid & 15) * 4;
  const bool g_vec = (g.ldg & 3) == 0 && (reinterpret_cast<uintptr_t>(g.G) & 15) == 0;
  const bool a_vec = (g.lda & 3) == 0 && (reinterpret_cast<uintptr_t>(g.A) & 15) == 0;
  float4 rg[2], ra;
  auto fetch = [&](int r0) {
    const int r = r0 + g_r;
    const bool r_ok = r < r_end;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int n = n0 + g_c + 64 * h;
      const float* q = g.G + (long long)r * g.ldg + n;
      if (r_ok && g_vec && n + 3 < g.N) rg[h] = __ldg(reinterpret_cast<const float4*>(q));
      else {
        rg[h].x = (r_ok && n + 0 < g.N) ? __ldg(q + 0) : 0.f;
        rg[h].y = (r_ok && n + 1 < g.N) ? __ldg(q + 1) : 0.f;
        rg[h].z = (r_ok && n + 2 < g.N) ? __ldg(q + 2) : 0.f;
        rg[h].w = (r_ok && n + 3 < g.N) ? __ldg(q + 3) : 0.f;
      }
    }
    const int t = r_ok ? r % g.T : -(1 << 30);
    const bool ok = (t + shift) >= 0 && (t + shift) < g.T;
    const int k = k0 + a_c;
    const float* q = g.A + (long long)(r + shift) * g.lda + k;
    if (ok && a_vec && k + 3 < g.K) ra = __ldg(reinterpret_cast<const float4*>(q));
    else {
      ra.x = (ok && k + 0 < g.K) ? __ldg(q + 0) : 0.f;
      ra.y = (ok && k + 1 < g.K) ? __ldg(q + 1) : 0.f;
      ra.z = (ok && k + 2 < g.K) ? __ldg(q + 2) : 0.f;
      ra.w = (ok && k + 3 < g.K) ? __ldg(q + 3) : 0.f;
    }
  };
  if (r_begin < r_end) fetch(r_begin);
  for (int r0 = r_begin; r0 < r_end; r0 += kGK) {
#pragma unroll
    for (int h = 0; h < 2; ++h)
      *reinterpret_cast<uint4*>(Gs + g_r * LDM + g_c + 64 * h) =
          make_uint4(to_tf32_bits(rg[h].x), to_tf32_bits(rg[h].y), to_tf32_bits(rg[h].z), to_tf32_bits(rg[h].w));
    *reinterpret_cast<uint4*>(As + a_r * kTcLdn + a_c) = make_uint4(to_tf32_bits(ra.x), to_tf32_bits(ra.y), to_tf32_bits(ra.z), to_tf32_bits(ra.w));
    __syncthreads();
    if (r0 + kGK < r_end) fetch(r0 + kGK);
    mma_tile_step<LDM>(Gs, As, wm, wn, gq, tq, acc);
    __syncthreads();
  }
  float* wp = g.dW + j * g.w_tap_stride;
#pragma unroll
  for (int mi = 0; mi < 2; ++mi)
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int n = n0 + wm + mi * 16 + gq + 8 * half;
      if (n >= g.N) continue;
#pragma unroll
      for (int ni = 0; ni < 4; ++ni)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int k = k0 + wn + ni * 8 + 2 * tq + c;
          if (k < g.K) atomicAdd(wp + (long long)n * (g.ldw ? g.ldw : g.K) + k, acc[mi][ni][2 * half + c]);
        }
    }
}

// Column sums  out[n] += sum_r X[r, n] (bias gradients).  grid = (ceil(N/32), row splits), block 32x8.
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ X, int ldx, int R, int N,
                                                     float* __restrict__ out, int rows_per_split) {
  __shared__ float red[8][33];
  const int n = blockIdx.x * 32 + threadIdx.x;
  const int r_begin = blockIdx.y * rows_per_split, r_end = min(R, r_begin + rows_per_split);
  float s = 0.f;
  if (n < N)
    for (int r = r_begin + threadIdx.y; r < r_end; r += 8) s += X[(long long)r * ldx + n];
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && n < N) {
#pragma unroll
    for (int i = 1; i < 8; ++i) s += red[i][threadIdx.x];
    atomicAdd(out + n, s);
  }
}

// ------------------------------------------------------------------------------------------
// weight_norm: W_eff[j][co][ci] = g[co] * v[co][ci][j] / ||v[co]||   (one block per co)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float block_sum(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  float s = 0.f;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += red[i];
  return s;
}

__global__ void __launch_bounds__(256) wn_fwd_kernel(const float* __restrict__ gparam, const float* __restrict__ v,
                                                     float* __restrict__ w_eff, float* __restrict__ inv_norm, int cout,
                                                     int cin, int k) {
  __shared__ float red[8];
  const int co = blockIdx.x;
  const int n = cin * k;
  const float* vp = v + (long long)co * n;
  float ss = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) ss += vp[i] * vp[i];
  ss = block_sum(ss, red);
  const float inv = 1.f / sqrtf(ss);
  const float s = gparam[co] * inv;
  if (threadIdx.x == 0) inv_norm[co] = inv;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int ci = i / k, j = i - ci * k;
    w_eff[((long long)j * cout + co) * cin + ci] = s * vp[i];
  }
}

// dg[co] = <dW, v>/||v||;  dv = g/||v|| * (dW - <dW, v>/||v||^2 * v)
__global__ void __launch_bounds__(256) wn_bwd_kernel(const float* __restrict__ gparam, const float* __restrict__ v,
                                                     const float* __restrict__ dw_eff, const float* __restrict__ inv_norm,
                                                     float* __restrict__ dg, float* __restrict__ dv, int cout, int cin, int k) {
  __shared__ float red[8];
  const int co = blockIdx.x;
  const int n = cin * k;
  const float* vp = v + (long long)co * n;
  float dot = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int ci = i / k, j = i - ci * k;
    dot += dw_eff[((long long)j * cout + co) * cin + ci] * vp[i];
  }
  dot = block_sum(dot, red);
  const float inv = inv_norm[co];
  if (threadIdx.x == 0) dg[co] = dot * inv;
  const float s = gparam[co] * inv, c = dot * inv * inv;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int ci = i / k, j = i - ci * k;
    dv[(long long)co * n + i] = s * (dw_eff[((long long)j * cout + co) * cin + ci] - c * vp[i]);
  }
}

// ------------------------------------------------------------------------------------------
// Block backward, first elementwise stage: from gy (grad of the block output y):
//   gpre = gy * LReLU'(y)                       (grad of h2d + res; also the residual-branch grad)
//   ga2  = gpre * drop2 * LReLU'(h2d)           (grad of conv2's pre-activation)
// ------------------------------------------------------------------------------------------
__global__ void block_bwd_pre_kernel(const float* __restrict__ gy, const float* __restrict__ y,
                                     const float* __restrict__ h2d, float* __restrict__ gpre, float* __restrict__ ga2,
                                     long long total, Drop drop) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const float gp = gy[i] * lrelu_grad(y[i]);
    gpre[i] = gp;
    ga2[i] = gp * drop_factor(drop, (uint32_t)i) * lrelu_grad(h2d[i]);
  }
}

// ------------------------------------------------------------------------------------------
// BatchNorm1d, training mode.  One block per 32 channels, 32x8 threads, three passes over the
// (L2-resident) column slab: mean, centred variance, normalise.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bn_train_fwd_kernel(const float* __restrict__ x, int ldx, int R, int C,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta,
                                                           float* __restrict__ z, int ldz, float* __restrict__ save_mean,
                                                           float* __restrict__ save_invstd, float* __restrict__ run_mean,
                                                           float* __restrict__ run_var, float momentum) {
  __shared__ float red[8][33];
  __shared__ float s_mean[32], s_inv[32];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const bool ok = c < C;
  float s = 0.f;
  if (ok) for (int r = threadIdx.y; r < R; r += 8) s += x[(long long)r * ldx + c];
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0) {
    for (int i = 1; i < 8; ++i) s += red[i][threadIdx.x];
    s_mean[threadIdx.x] = s / R;
  }
  __syncthreads();
  const float mean = s_mean[threadIdx.x];
  float v = 0.f;
  if (ok) for (int r = threadIdx.y; r < R; r += 8) { const float d = x[(long long)r * ldx + c] - mean; v += d * d; }
  red[threadIdx.y][threadIdx.x] = v;
  __syncthreads();
  if (threadIdx.y == 0) {
    for (int i = 1; i < 8; ++i) v += red[i][threadIdx.x];
    const float var = v / R;
    s_inv[threadIdx.x] = 1.f / sqrtf(var + kBnEps);
    if (ok) {
      save_mean[c] = mean;
      save_invstd[c] = s_inv[threadIdx.x];
      run_mean[c] = (1.f - momentum) * run_mean[c] + momentum * mean;
      run_var[c] = (1.f - momentum) * run_var[c] + momentum * (R > 1 ? v / (R - 1) : var);
    }
  }
  __syncthreads();
  if (!ok) return;
  const float inv = s_inv[threadIdx.x], ga = gamma[c], be = beta[c];
  for (int r = threadIdx.y; r < R; r += 8) z[(long long)r * ldz + c] = (x[(long long)r * ldx + c] - mean) * inv * ga + be;
}

__global__ void __launch_bounds__(256) bn_train_bwd_kernel(const float* __restrict__ gz, int ldgz, const float* __restrict__ x,
                                                           int ldx, int R, int C, const float* __restrict__ gamma,
                                                           const float* __restrict__ save_mean,
                                                           const float* __restrict__ save_invstd, float* __restrict__ gx,
                                                           int ldgx, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  __shared__ float red[2][8][33];
  __shared__ float s_sum[2][32];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const bool ok = c < C;
  const float mean = ok ? save_mean[c] : 0.f, inv = ok ? save_invstd[c] : 0.f;
  float sb = 0.f, sg = 0.f;
  if (ok)
    for (int r = threadIdx.y; r < R; r += 8) {
      const float g = gz[(long long)r * ldgz + c];
      sb += g;
      sg += g * (x[(long long)r * ldx + c] - mean) * inv;
    }
  red[0][threadIdx.y][threadIdx.x] = sb;
  red[1][threadIdx.y][threadIdx.x] = sg;
  __syncthreads();
  if (threadIdx.y == 0) {
    for (int i = 1; i < 8; ++i) { sb += red[0][i][threadIdx.x]; sg += red[1][i][threadIdx.x]; }
    s_sum[0][threadIdx.x] = sb;
    s_sum[1][threadIdx.x] = sg;
    if (ok) { dbeta[c] = sb; dgamma[c] = sg; }
  }
  __syncthreads();
  if (!ok) return;
  sb = s_sum[0][threadIdx.x]; sg = s_sum[1][threadIdx.x];
  const float k = gamma[c] * inv, invR = 1.f / R;
  for (int r = threadIdx.y; r < R; r += 8) {
    const float xh = (x[(long long)r * ldx + c] - mean) * inv;
    gx[(long long)r * ldgx + c] = k * (gz[(long long)r * ldgz + c] - sb * invR - xh * sg * invR);
  }
}

// ------------------------------------------------------------------------------------------
// Cross-modal attention over the M modality tokens (transformer.py:132-161).  qkv[m]: [R][3*md],
// per head h the slice [h*3*hd, (h+1)*3*hd) is q|k|v.  vals[r][(h*M + m)*hd + d].
// One thread per (row, head); M <= 4, hd <= 32.
// ------------------------------------------------------------------------------------------
struct AttnArgs {
  const float* qkv[CER_MAX_MODALS];
  float* gqkv[CER_MAX_MODALS];
  float* vals;          // fwd out [R][M*md]
  const float* gvals;   // bwd in
  int R, M, H, hd;
};

__device__ __forceinline__ void attn_probs(const AttnArgs& a, int r, int h, float att[CER_MAX_MODALS][CER_MAX_MODALS]) {
  const int md3 = 3 * a.H * a.hd;
  const float scale = rsqrtf((float)a.hd);
  for (int m = 0; m < a.M; ++m) {
    const float* q = a.qkv[m] + (long long)r * md3 + h * 3 * a.hd;
    float mx = -1e30f;
    for (int n = 0; n < a.M; ++n) {
      const float* k = a.qkv[n] + (long long)r * md3 + h * 3 * a.hd + a.hd;
      float s = 0.f;
      for (int d = 0; d < a.hd; ++d) s = fmaf(q[d], k[d], s);
      att[m][n] = s * scale;
      mx = fmaxf(mx, att[m][n]);
    }
    float den = 0.f;
    for (int n = 0; n < a.M; ++n) { att[m][n] = __expf(att[m][n] - mx); den += att[m][n]; }
    const float inv = 1.f / den;
    for (int n = 0; n < a.M; ++n) att[m][n] *= inv;
  }
}

__global__ void attn_fwd_kernel(const AttnArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.R * a.H) return;
  const int r = i / a.H, h = i - r * a.H;
  const int md3 = 3 * a.H * a.hd, E = a.M * a.H * a.hd;
  float att[CER_MAX_MODALS][CER_MAX_MODALS];
  attn_probs(a, r, h, att);
  for (int m = 0; m < a.M; ++m) {
    float* o = a.vals + (long long)r * E + (h * a.M + m) * a.hd;
    const float* vm = a.qkv[m] + (long long)r * md3 + h * 3 * a.hd + 2 * a.hd;
    for (int d = 0; d < a.hd; ++d) {
      float s = vm[d];                                     // "+ V" residual (transformer.py:157)
      for (int n = 0; n < a.M; ++n) s = fmaf(att[m][n], a.qkv[n][(long long)r * md3 + h * 3 * a.hd + 2 * a.hd + d], s);
      o[d] = s;
    }
  }
}

// attention probabilities only: maps[r][h][m][n] (MultimodalMultiheadAttention(return_attention=True))
__global__ void attn_maps_kernel(const AttnArgs a, float* __restrict__ maps) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.R * a.H) return;
  const int r = i / a.H, h = i - r * a.H;
  float att[CER_MAX_MODALS][CER_MAX_MODALS];
  attn_probs(a, r, h, att);
  float* o = maps + (long long)i * a.M * a.M;
  for (int m = 0; m < a.M; ++m)
    for (int n = 0; n < a.M; ++n) o[m * a.M + n] = att[m][n];
}

__global__ void attn_bwd_kernel(const AttnArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.R * a.H) return;
  const int r = i / a.H, h = i - r * a.H;
  const int md3 = 3 * a.H * a.hd, E = a.M * a.H * a.hd;
  const float scale = rsqrtf((float)a.hd);
  float att[CER_MAX_MODALS][CER_MAX_MODALS], gl[CER_MAX_MODALS][CER_MAX_MODALS];
  attn_probs(a, r, h, att);
  const long long base = (long long)r * md3 + h * 3 * a.hd;
  for (int m = 0; m < a.M; ++m) {
    const float* gv = a.gvals + (long long)r * E + (h * a.M + m) * a.hd;
    float dotsum = 0.f;
    for (int n = 0; n < a.M; ++n) {
      const float* vn = a.qkv[n] + base + 2 * a.hd;
      float s = 0.f;
      for (int d = 0; d < a.hd; ++d) s = fmaf(gv[d], vn[d], s);
      gl[m][n] = s;                       // d loss / d att[m][n]
      dotsum = fmaf(att[m][n], s, dotsum);
    }
    for (int n = 0; n < a.M; ++n) gl[m][n] = att[m][n] * (gl[m][n] - dotsum) * scale;   // d loss / d (q.k)
  }
  for (int m = 0; m < a.M; ++m) {
    float* gq = a.gqkv[m] + base;
    float* gk = gq + a.hd;
    float* gvv = gq + 2 * a.hd;
    const float* gvm = a.gvals + (long long)r * E + (h * a.M + m) * a.hd;
    for (int d = 0; d < a.hd; ++d) {
      float sq = 0.f, sk = 0.f, sv = gvm[d];
      for (int n = 0; n < a.M; ++n) {
        sq = fmaf(gl[m][n], a.qkv[n][base + a.hd + d], sq);            // dq_m = sum_n gl[m][n] k_n
        sk = fmaf(gl[n][m], a.qkv[n][base + d], sk);                   // dk_m = sum_n gl[n][m] q_n
        sv = fmaf(att[n][m], a.gvals[(long long)r * E + (h * a.M + n) * a.hd + d], sv);   // dv_m = sum_n att[n][m] gvals_n + gvals_m
      }
      gq[d] = sq; gk[d] = sk; gvv[d] = sv;
    }
  }
}

// ------------------------------------------------------------------------------------------
// Dropout + LayerNorm over E <= 128 features, one warp per row (transformer.py:194-196).
// fwd: od = drop(o) (in place), f = LN(od) written with leading dimension ldf; saves mean, rstd.
// bwd: gf (ld ldg) -> go = dLN * drop;  dgamma/dbeta via per-block smem partials + atomicAdd.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ln_drop_fwd_kernel(float* __restrict__ o, int R, int E, const float* __restrict__ gamma,
                                                          const float* __restrict__ beta, float* __restrict__ f, int ldf,
                                                          float* __restrict__ save_mean, float* __restrict__ save_rstd,
                                                          Drop drop) {
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= R) return;
  float v[4];
  float s = 0.f;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int e = lane + 32 * q;
    v[q] = 0.f;
    if (e < E) {
      v[q] = o[(long long)r * E + e] * drop_factor(drop, (uint32_t)r * (uint32_t)E + (uint32_t)e);
      o[(long long)r * E + e] = v[q];
      s += v[q];
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  const float mean = s / E;
  float ss = 0.f;
#pragma unroll
  for (int q = 0; q < 4; ++q) if (lane + 32 * q < E) { const float d = v[q] - mean; ss += d * d; }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
  const float rstd = 1.f / sqrtf(ss / E + kLnEps);
  if (lane == 0) { save_mean[r] = mean; save_rstd[r] = rstd; }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int e = lane + 32 * q;
    if (e < E) f[(long long)r * ldf + e] = (v[q] - mean) * rstd * gamma[e] + beta[e];
  }
}

__global__ void __launch_bounds__(256) ln_drop_bwd_kernel(const float* __restrict__ gf, int ldg, const float* __restrict__ od,
                                                          int R, int E, const float* __restrict__ gamma,
                                                          const float* __restrict__ save_mean, const float* __restrict__ save_rstd,
                                                          float* __restrict__ go, float* __restrict__ dgamma,
                                                          float* __restrict__ dbeta, Drop drop, int rows_per_block) {
  __shared__ float s_dg[128], s_db[128];
  if (threadIdx.x < 128) { s_dg[threadIdx.x] = 0.f; s_db[threadIdx.x] = 0.f; }
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float pdg[4] = {0.f, 0.f, 0.f, 0.f}, pdb[4] = {0.f, 0.f, 0.f, 0.f};
  const int r_begin = blockIdx.x * rows_per_block, r_end = min(R, r_begin + rows_per_block);
  for (int r = r_begin + w; r < r_end; r += 8) {
    const float mean = save_mean[r], rstd = save_rstd[r];
    float xh[4], gg[4];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int e = lane + 32 * q;
      xh[q] = 0.f; gg[q] = 0.f;
      if (e < E) {
        const float g = gf[(long long)r * ldg + e];
        xh[q] = (od[(long long)r * E + e] - mean) * rstd;
        gg[q] = g * gamma[e];
        pdg[q] += g * xh[q];
        pdb[q] += g;
        s1 += gg[q];
        s2 += gg[q] * xh[q];
      }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, off); s2 += __shfl_xor_sync(0xffffffffu, s2, off); }
    s1 /= E; s2 /= E;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int e = lane + 32 * q;
      if (e < E) go[(long long)r * E + e] = rstd * (gg[q] - s1 - xh[q] * s2) * drop_factor(drop, (uint32_t)r * (uint32_t)E + (uint32_t)e);
    }
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int e = lane + 32 * q;
    if (e < E) { atomicAdd(&s_dg[e], pdg[q]); atomicAdd(&s_db[e], pdb[q]); }
  }
  __syncthreads();
  if (threadIdx.x < E) { atomicAdd(dgamma + threadIdx.x, s_dg[threadIdx.x]); atomicAdd(dbeta + threadIdx.x, s_db[threadIdx.x]); }
}

// Mean cross-entropy over rows + its gradient (experiment.py:133; trainer.py:372-381).
__global__ void __launch_bounds__(256) ce_loss_kernel(const float* __restrict__ logits, const long long* __restrict__ labels,
                                                      int R, int n_cls, float* __restrict__ loss, float* __restrict__ dlogits) {
  __shared__ float red[8];
  // nn.CrossEntropyLoss(reduction="mean") averages over the rows whose label is not ignore_index (-100).
  // Every block counts the valid labels itself (R is a few thousand rows on this path), so the kernel
  // needs no scratch memory and no second launch.  Labels outside [0, n_cls) are treated as ignored:
  // they contribute neither loss nor gradient and are never used as an index.
  float cnt = 0.f;
  for (int i = threadIdx.x; i < R; i += blockDim.x) {
    const long long yi = labels[i];
    cnt += (yi >= 0 && yi < n_cls) ? 1.f : 0.f;
  }
  cnt = block_sum(cnt, red);
  __syncthreads();
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  float l = 0.f;
  if (r < R) {
    const float* p = logits + (long long)r * n_cls;
    const long long y = labels[r];
    const bool ok = y >= 0 && y < n_cls;
    const float inv = 1.f / cnt;
    if (ok) {
      float mx = p[0];
      for (int c = 1; c < n_cls; ++c) mx = fmaxf(mx, p[c]);
      float den = 0.f;
      for (int c = 0; c < n_cls; ++c) den += expf(p[c] - mx);
      const float lse = logf(den) + mx;
      l = (lse - p[y]) * inv;
      if (dlogits)
        for (int c = 0; c < n_cls; ++c) dlogits[(long long)r * n_cls + c] = (expf(p[c] - lse) - (c == (int)y ? 1.f : 0.f)) * inv;
    } else if (dlogits) {
      for (int c = 0; c < n_cls; ++c) dlogits[(long long)r * n_cls + c] = 0.f;
    }
  }
  l = block_sum(l, red);
  if (threadIdx.x == 0) atomicAdd(loss, cnt > 0.f ? l : (blockIdx.x == 0 ? __int_as_float(0x7fc00000) : 0.f));   // no valid row: nan, as torch
}

// Fused optimizer update on a flat fp32 buffer (torch.optim.{SGD,Adam,AdamW} arithmetic;
// instantiators.py:60-100).  kind 0: SGD(momentum, dampening, nesterov); 1: Adam; 2: AdamW.
__global__ void optimizer_kernel(int kind, float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                 float* __restrict__ v, long long n, float lr, float wd, float b1, float b2, float eps,
                                 int nesterov, int step, float bc1, float bc2_sqrt, float grad_scale) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float w = p[i], gr = g[i] * grad_scale;
    if (kind == 0) {
      gr = fmaf(wd, w, gr);
      if (b1 != 0.f) {
        const float buf = step == 1 ? gr : m[i] * b1 + (1.f - b2) * gr;      // b2 carries the dampening
        m[i] = buf;
        gr = nesterov ? gr + b1 * buf : buf;
      }
      p[i] = w - lr * gr;
    } else {
      if (kind == 2) w *= (1.f - lr * wd); else gr = fmaf(wd, w, gr);
      const float mm = m[i] * b1 + (1.f - b1) * gr;
      const float vv = v[i] * b2 + (1.f - b2) * gr * gr;
      m[i] = mm; v[i] = vv;
      const float denom = sqrtf(vv) / bc2_sqrt + eps;
      p[i] = w - (lr / bc1) * mm / denom;
    }
  }
}

}  // namespace cer

using namespace cer;

// ==========================================================================================
// Host side
// ==========================================================================================
namespace {

struct ConvBuf {          // per weight-normed conv
  float* w_eff;           // [k][cout][cin]
  float* inv_norm;        // [cout]
  float* dw_eff;          // [k][cout][cin] (zeroed per backward)
};
struct BlockBuf {
  ConvBuf c1, c2;
  float *h1d, *h2d, *y, *res;     // [R][cout]; res only with a downsample
};
struct ModalBuf {
  std::vector<BlockBuf> blk;
  float *bn_mean, *bn_invstd;     // [C]
  float* z;                       // [R][C] (modality 0: inside `cat`, leading dimension ld_z)
  int ld_z;
  float *qkv, *gqkv;              // [R][3*md]
  float *ga, *gb, *gc, *gd;       // backward ping-pong [R][max C of this modality] (modalities run concurrently)
};

inline size_t align_up(size_t x) { return (x + 255) & ~size_t(255); }

}  // namespace
int cer::launch_row_gemm(const RowGemm& g, bool b_kn, cudaStream_t st) {
  const int nt = (g.N + kGN - 1) / kGN;
  const bool big = (long long)((g.R + 127) / 128) * nt >= 2 * 148;      // enough 128-row tiles for two waves
  if (g.tf32) {
    // tensor-core path: 128-row tiles when they fill the SMs twice, else 64-row tiles (twice the CTAs)
    if (big) {
      dim3 grid(nt, (g.R + 127) / 128);
      if (b_kn) row_gemm_tf32_kernel<true, 128><<<grid, 256, 0, st>>>(g);
      else row_gemm_tf32_kernel<false, 128><<<grid, 256, 0, st>>>(g);
    } else {
      dim3 grid(nt, (g.R + 63) / 64);
      if (b_kn) row_gemm_tf32_kernel<true, 64><<<grid, 128, 0, st>>>(g);
      else row_gemm_tf32_kernel<false, 64><<<grid, 128, 0, st>>>(g);
    }
    CER_CUDA(cudaGetLastError());
    return CER_OK;
  }
  if (big) {
    dim3 grid(nt, (g.R + 127) / 128);
    if (b_kn) row_gemm_kernel<true, 128><<<grid, 256, 0, st>>>(g);
    else row_gemm_kernel<false, 128><<<grid, 256, 0, st>>>(g);
  } else {
    dim3 grid(nt, (g.R + 63) / 64);
    static int split = -1;
    if (split < 0) {
      const char* e = getenv("CER_GEMM_SPLITK");
      split = (e && e[0] == '0') ? 0 : 1;
    }
    static unsigned long long configured = 0;
    if (split && first_use_on_device(&configured)) {
      CER_CUDA(cudaFuncSetAttribute(row_gemm_splitk_kernel<true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * kSplitGroupBytes));
      CER_CUDA(cudaFuncSetAttribute(row_gemm_splitk_kernel<false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * kSplitGroupBytes));
    }
    const long long iters = (long long)g.taps * ((g.K + kGK - 1) / kGK);
    const long long ctas = (long long)grid.x * grid.y;
    if (split && ctas <= 148 && iters >= 8) {
      // at most one tile per SM and a long K: four K-groups per CTA (16 warps per SM instead of 4)
      if (b_kn) row_gemm_splitk_kernel<true, 4><<<grid, 512, 4 * kSplitGroupBytes, st>>>(g);
      else row_gemm_splitk_kernel<false, 4><<<grid, 512, 4 * kSplitGroupBytes, st>>>(g);
    } else if (split && ctas <= 2 * 148 && iters >= 4) {
      // up to two tiles per SM: two K-groups (256 threads, two CTAs per SM by registers)
      if (b_kn) row_gemm_splitk_kernel<true, 2><<<grid, 256, 2 * kSplitGroupBytes, st>>>(g);
      else row_gemm_splitk_kernel<false, 2><<<grid, 256, 2 * kSplitGroupBytes, st>>>(g);
    } else if (b_kn) row_gemm_kernel<true, 64><<<grid, 128, 0, st>>>(g);
    else row_gemm_kernel<false, 64><<<grid, 128, 0, st>>>(g);
  }
  CER_CUDA(cudaGetLastError());
  return CER_OK;
}
namespace {

int launch_wgrad(WGrad g, int num_sms, cudaStream_t st) {
  g.k_tiles = (g.K + kGN - 1) / kGN;
  const int tiles = g.k_tiles * g.taps * ((g.N + kGM - 1) / kGM);
  int splits = std::max(1, std::min((3 * num_sms + tiles - 1) / tiles, (g.R + 255) / 256));
  g.rows_per_split = (((g.R + splits - 1) / splits) + 15) / 16 * 16;
  splits = (g.R + g.rows_per_split - 1) / g.rows_per_split;
  dim3 grid(g.k_tiles * g.taps, (g.N + kGM - 1) / kGM, splits);
  if (g.tf32) wgrad_tf32_kernel<<<grid, 256, 0, st>>>(g);
  else wgrad_kernel<<<grid, 256, 0, st>>>(g);
  CER_CUDA(cudaGetLastError());
  return CER_OK;
}

int launch_colsum(const float* X, int ldx, int R, int N, float* out, cudaStream_t st) {
  const int splits = std::max(1, std::min(32, R / 256));
  const int rps = (R + splits - 1) / splits;
  dim3 grid((N + 31) / 32, (R + rps - 1) / rps), block(32, 8);
  colsum_kernel<<<grid, block, 0, st>>>(X, ldx, R, N, out, rps);
  CER_CUDA(cudaGetLastError());
  return CER_OK;
}

Drop make_drop(double p, uint32_t seed, uint32_t stream) {
  Drop d;
  d.key = seed + stream * 0x85EBCA6Bu;
  if (p <= 0.0) { d.thr = 0; d.scale = 1.f; return d; }
  const double t = p * 4294967296.0;
  d.thr = t >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)(long long)t;
  d.scale = 1.0f / (1.0f - (float)p);
  return d;
}

}  // namespace

struct cer_head_train {
  cer_head_train_spec s;
  int B, T, R, E, md3, num_sms;
  int tf32;                         // spec.precision: 1 = TF32 tensor-core GEMMs for the TCN convolutions (forward, dgrad, wgrad:
                                    // 99 % of the step's FLOPs), 0 = exact fp32.  The Linear layers of the fusion head
                                    // (qkv_proj, o_proj, regressor) stay fp32 in both modes, like torch's defaults
                                    // (cudnn.allow_tf32 = True, cuda.matmul.allow_tf32 = False)
  std::vector<ModalBuf> mod;
  float *vals, *o, *ln_mean, *ln_rstd, *cat, *gcat, *go, *gvals;
  cudaStream_t side[CER_MAX_MODALS];      // modality m > 0 runs on side[m]; modality 0 on the caller's stream
  cudaEvent_t ev_fork, ev_join[CER_MAX_MODALS];
  float* scratch_zero; size_t scratch_zero_bytes;    // all dw_eff buffers, zeroed per backward
  uint32_t seed;                  // seed of the last forward (backward re-derives the masks)
  int ld_cat;
};

static int spec_check(const cer_head_train_spec* s) {
  if (!s || s->n_modals < 1 || s->n_modals > CER_MAX_MODALS || s->kernel_size < 1 || s->modal_dim % s->num_heads ||
      s->modal_dim / s->num_heads > 32 || s->modal_dim * s->n_modals > 128 || s->n_out < 1 || !s->grad_flat)
    return set_error(CER_ERR_INVALID, "cer_head_train: bad spec");
  for (int m = 0; m < s->n_modals; ++m)
    if (s->modal[m].n_blocks < 1 || s->modal[m].n_blocks > CER_MAX_TCN_BLOCKS) return set_error(CER_ERR_INVALID, "cer_head_train: n_blocks");
  return CER_OK;
}

static size_t plan_layout(const cer_head_train_spec* s, int64_t B, int64_t T, cer_head_train* p, uint8_t* base) {
  // walks the workspace; with p == nullptr only the size is computed
  size_t off = 0;
  auto take = [&](size_t floats) { float* q = base ? reinterpret_cast<float*>(base + off) : nullptr; off += align_up(floats * 4); return q; };
  const size_t R = (size_t)B * T;
  const int k = s->kernel_size;
  const int E = s->modal_dim * s->n_modals, md3 = 3 * s->modal_dim;
  const int c0 = s->modal[0].blocks[s->modal[0].n_blocks - 1].c_out;
  int maxc = std::max(E, c0 + E);
  // zeroed scratch first (contiguous)
  size_t zero_begin = off;
  std::vector<std::vector<std::pair<float*, float*>>> dws(s->n_modals);
  for (int m = 0; m < s->n_modals; ++m)
    for (int i = 0; i < s->modal[m].n_blocks; ++i) {
      const cer_train_block& b = s->modal[m].blocks[i];
      float* d1 = take((size_t)k * b.c_out * b.c_in);
      float* d2 = take((size_t)k * b.c_out * b.c_out);
      dws[m].push_back({d1, d2});
      maxc = std::max(maxc, std::max(b.c_in, b.c_out));
    }
  size_t zero_end = off;
  if (p) { p->scratch_zero = reinterpret_cast<float*>(base + zero_begin); p->scratch_zero_bytes = zero_end - zero_begin; }
  float* cat = take(R * (c0 + E));
  if (p) { p->cat = cat; p->ld_cat = c0 + E; p->mod.resize(s->n_modals); }
  for (int m = 0; m < s->n_modals; ++m) {
    ModalBuf mb;
    for (int i = 0; i < s->modal[m].n_blocks; ++i) {
      const cer_train_block& b = s->modal[m].blocks[i];
      BlockBuf bb;
      bb.c1.w_eff = take((size_t)k * b.c_out * b.c_in); bb.c1.inv_norm = take(b.c_out); bb.c1.dw_eff = dws[m][i].first;
      bb.c2.w_eff = take((size_t)k * b.c_out * b.c_out); bb.c2.inv_norm = take(b.c_out); bb.c2.dw_eff = dws[m][i].second;
      bb.h1d = take(R * b.c_out); bb.h2d = take(R * b.c_out); bb.y = take(R * b.c_out);
      bb.res = b.wd ? take(R * b.c_out) : nullptr;
      mb.blk.push_back(bb);
    }
    const int C = s->modal[m].blocks[s->modal[m].n_blocks - 1].c_out;
    mb.bn_mean = take(C); mb.bn_invstd = take(C);
    if (m == 0) { mb.z = cat; mb.ld_z = c0 + E; } else { mb.z = take(R * C); mb.ld_z = C; }
    mb.qkv = take(R * md3); mb.gqkv = take(R * md3);
    int mc = md3;
    for (int i = 0; i < s->modal[m].n_blocks; ++i) mc = std::max(mc, std::max(s->modal[m].blocks[i].c_in, s->modal[m].blocks[i].c_out));
    mb.ga = take(R * mc); mb.gb = take(R * mc); mb.gc = take(R * mc); mb.gd = take(R * mc);
    if (p) p->mod[m] = mb;
  }
  float* vals = take(R * E); float* o = take(R * E); float* lm = take(R); float* lr = take(R);
  float* gcat = take(R * (c0 + E)); float* go = take(R * E); float* gvals = take(R * E);
  (void)maxc;
  if (p) { p->vals = vals; p->o = o; p->ln_mean = lm; p->ln_rstd = lr; p->gcat = gcat; p->go = go; p->gvals = gvals; }
  return off + 256;
}

extern "C" size_t cer_head_train_workspace_bytes(const cer_head_train_spec* s, int64_t batch, int64_t length) {
  if (spec_check(s) || batch <= 0 || length <= 0) return 0;
  return plan_layout(s, batch, length, nullptr, nullptr);
}

extern "C" int cer_head_train_create(cer_head_train** out, const cer_head_train_spec* s, int64_t batch, int64_t length,
                                     void* workspace_dev, size_t workspace_bytes) {
  int rc = spec_check(s);
  if (rc) return rc;
  if (!out || !workspace_dev || batch <= 0 || length <= 0 || batch * length > (1 << 22))
    return set_error(CER_ERR_INVALID, "cer_head_train_create: bad argument");
  rc = cer_check_device();
  if (rc) return rc;
  if (workspace_bytes < plan_layout(s, batch, length, nullptr, nullptr)) return set_error(CER_ERR_WORKSPACE, "cer_head_train_create: workspace too small");
  cer_head_train* p = new cer_head_train();
  p->s = *s;
  p->B = (int)batch; p->T = (int)length; p->R = (int)(batch * length);
  p->E = s->modal_dim * s->n_modals; p->md3 = 3 * s->modal_dim;
  p->tf32 = s->precision == 1 ? 1 : 0;
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&p->num_sms, cudaDevAttrMultiProcessorCount, dev);
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace_dev) + 255) & ~uintptr_t(255));
  plan_layout(s, batch, length, p, base);
  p->seed = 0;
  for (int m = 0; m < CER_MAX_MODALS; ++m) { p->side[m] = nullptr; p->ev_join[m] = nullptr; }
  p->ev_fork = nullptr;
  CER_CUDA(cudaEventCreateWithFlags(&p->ev_fork, cudaEventDisableTiming));
  for (int m = 1; m < s->n_modals; ++m) {
    CER_CUDA(cudaStreamCreateWithFlags(&p->side[m], cudaStreamNonBlocking));
    CER_CUDA(cudaEventCreateWithFlags(&p->ev_join[m], cudaEventDisableTiming));
  }
  *out = p;
  return CER_OK;
}

extern "C" void cer_head_train_destroy(cer_head_train* p) {
  if (!p) return;
  for (int m = 1; m < CER_MAX_MODALS; ++m) {
    if (p->side[m]) cudaStreamDestroy(p->side[m]);
    if (p->ev_join[m]) cudaEventDestroy(p->ev_join[m]);
  }
  if (p->ev_fork) cudaEventDestroy(p->ev_fork);
  delete p;
}

#define RC(x) do { int _rc = (x); if (_rc) return _rc; } while (0)

static int head_train_forward_impl(cer_head_train* p, const float* const* feats, uint32_t seed, float* logits, float* const* z_out,
                                   void* stream);
static int head_train_backward_impl(cer_head_train* p, const float* const* feats, const float* dlogits, const float* const* dz,
                                    void* stream);

extern "C" int cer_head_train_forward(cer_head_train* p, const float* const* feats, uint32_t seed, float* logits, void* stream) {
  if (!logits) return set_error(CER_ERR_INVALID, "cer_head_train_forward: bad argument");
  return head_train_forward_impl(p, feats, seed, logits, nullptr, stream);
}

// TemporalConvNet + BatchNorm1d of every modality only (what CAN / JMT / MT share with LFAN, models/model.py:672-676,
// :1155-1159): z_out[m] receives fp32 [batch*length][c_last[m]] (dense).
extern "C" int cer_head_train_tcn_forward(cer_head_train* p, const float* const* feats, uint32_t seed, float* const* z_out,
                                          void* stream) {
  if (!z_out) return set_error(CER_ERR_INVALID, "cer_head_train_tcn_forward: bad argument");
  return head_train_forward_impl(p, feats, seed, nullptr, z_out, stream);
}

extern "C" int cer_head_train_backward(cer_head_train* p, const float* const* feats, const float* dlogits, void* stream) {
  if (!dlogits) return set_error(CER_ERR_INVALID, "cer_head_train_backward: bad argument");
  return head_train_backward_impl(p, feats, dlogits, nullptr, stream);
}

// Backward of cer_head_train_tcn_forward: dz[m] = d loss / d z[m].  Does NOT clear the flat gradient buffer (the
// caller zeroes it once per step and other blocks add their gradients to it); bias / downsample gradients are
// accumulated, weight_g / weight_v / BatchNorm gradients are written.
extern "C" int cer_head_train_tcn_backward(cer_head_train* p, const float* const* feats, const float* const* dz, void* stream) {
  if (!dz) return set_error(CER_ERR_INVALID, "cer_head_train_tcn_backward: bad argument");
  return head_train_backward_impl(p, feats, nullptr, dz, stream);
}

static int head_train_forward_impl(cer_head_train* p, const float* const* feats, uint32_t seed, float* logits, float* const* z_out,
                                   void* stream) {
  if (!p || !feats) return set_error(CER_ERR_INVALID, "cer_head_train_forward: bad argument");
  const bool tcn_only = z_out != nullptr;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const cer_head_train_spec& s = p->s;
  const int R = p->R, T = p->T, k = s.kernel_size;
  p->seed = seed;
  cudaStream_t main_st = st;
  // the modalities are independent until the attention: run them concurrently (each TCN alone has too
  // few tiles to fill 148 SMs), modality 0 on the caller's stream, the others on side streams
  CER_CUDA(cudaEventRecord(p->ev_fork, main_st));
  for (int m = 0; m < s.n_modals; ++m) {
    const cer_train_modal& M = s.modal[m];
    ModalBuf& mb = p->mod[m];
    const float* x = feats[m];
    if (!x) return set_error(CER_ERR_INVALID, "cer_head_train_forward: null feature pointer");
    st = m == 0 ? main_st : p->side[m];
    if (m > 0) CER_CUDA(cudaStreamWaitEvent(st, p->ev_fork, 0));
    for (int i = 0; i < M.n_blocks; ++i) {
      const cer_train_block& b = M.blocks[i];
      BlockBuf& bb = mb.blk[i];
      wn_fwd_kernel<<<b.c_out, 256, 0, st>>>(b.conv1.g, b.conv1.v, bb.c1.w_eff, bb.c1.inv_norm, b.c_out, b.c_in, k);
      wn_fwd_kernel<<<b.c_out, 256, 0, st>>>(b.conv2.g, b.conv2.v, bb.c2.w_eff, bb.c2.inv_norm, b.c_out, b.c_out, k);
      CER_CUDA(cudaGetLastError());
      RowGemm g{}; g.tf32 = p->tf32;
      g.R = R; g.T = T; g.taps = k; g.shift0 = -(k - 1) * b.dilation; g.shift_step = b.dilation;
      // conv1: h1d = drop(LReLU(conv1(x) + b1))
      g.A = x; g.lda = b.c_in; g.K = b.c_in; g.B = bb.c1.w_eff; g.b_tap_stride = (long long)b.c_out * b.c_in;
      g.C = bb.h1d; g.ldc = b.c_out; g.N = b.c_out; g.bias = b.conv1.bias; g.epi = EPI_LRELU_DROP;
      g.drop = make_drop(s.p_tcn, seed, m * 16 + i * 2 + 0);
      RC(launch_row_gemm(g, false, st));
      // residual branch
      const float* res = x; int ld_res = b.c_in;
      if (b.wd) {
        RowGemm d{}; d.tf32 = p->tf32;
        d.R = R; d.T = T; d.taps = 1; d.A = x; d.lda = b.c_in; d.K = b.c_in; d.B = b.wd; d.C = bb.res; d.ldc = b.c_out;
        d.N = b.c_out; d.bias = b.bd; d.epi = EPI_LINEAR;
        RC(launch_row_gemm(d, false, st));
        res = bb.res; ld_res = b.c_out;
      }
      // conv2: h2d = drop(LReLU(conv2(h1d) + b2)); y = LReLU(h2d + res)
      g.A = bb.h1d; g.lda = b.c_out; g.K = b.c_out; g.B = bb.c2.w_eff; g.b_tap_stride = (long long)b.c_out * b.c_out;
      g.C = bb.y; g.bias = b.conv2.bias; g.epi = EPI_BLOCK_OUT; g.aux = bb.h2d; g.ld_aux = b.c_out;
      g.addend = res; g.ld_add = ld_res;
      g.drop = make_drop(s.p_tcn, seed, m * 16 + i * 2 + 1);
      RC(launch_row_gemm(g, false, st));
      x = bb.y;
    }
    const int C = M.blocks[M.n_blocks - 1].c_out;
    if (tcn_only && !z_out[m]) return set_error(CER_ERR_INVALID, "cer_head_train_tcn_forward: null output pointer");
    bn_train_fwd_kernel<<<(C + 31) / 32, dim3(32, 8), 0, st>>>(x, C, R, C, M.bn_w, M.bn_b, tcn_only ? z_out[m] : mb.z,
                                                                tcn_only ? C : mb.ld_z, mb.bn_mean, mb.bn_invstd, M.bn_mean,
                                                                M.bn_var, (float)s.bn_momentum);
    CER_CUDA(cudaGetLastError());
    if (tcn_only) {
      if (m > 0) CER_CUDA(cudaEventRecord(p->ev_join[m], st));
      continue;
    }
    RowGemm q{};
    q.R = R; q.T = T; q.taps = 1; q.A = mb.z; q.lda = mb.ld_z; q.K = C; q.B = M.wqkv; q.C = mb.qkv; q.ldc = p->md3;
    q.N = p->md3; q.bias = M.bqkv; q.epi = EPI_LINEAR;
    RC(launch_row_gemm(q, false, st));
    if (m > 0) CER_CUDA(cudaEventRecord(p->ev_join[m], st));
  }
  st = main_st;
  for (int m = 1; m < s.n_modals; ++m) CER_CUDA(cudaStreamWaitEvent(st, p->ev_join[m], 0));
  if (tcn_only) return CER_OK;
  AttnArgs a{};
  for (int m = 0; m < s.n_modals; ++m) a.qkv[m] = p->mod[m].qkv;
  a.vals = p->vals; a.R = R; a.M = s.n_modals; a.H = s.num_heads; a.hd = s.modal_dim / s.num_heads;
  attn_fwd_kernel<<<(R * a.H + 127) / 128, 128, 0, st>>>(a);
  CER_CUDA(cudaGetLastError());
  const int E = p->E, c0 = p->ld_cat - E;
  RowGemm o{};
  o.R = R; o.T = T; o.taps = 1; o.A = p->vals; o.lda = E; o.K = E; o.B = s.wo; o.C = p->o; o.ldc = E; o.N = E; o.bias = s.bo;
  o.epi = EPI_LINEAR;
  RC(launch_row_gemm(o, false, st));
  ln_drop_fwd_kernel<<<(R + 7) / 8, 256, 0, st>>>(p->o, R, E, s.ln_g, s.ln_b, p->cat + c0, p->ld_cat, p->ln_mean, p->ln_rstd,
                                                  make_drop(s.p_fusion, seed, 4096));
  CER_CUDA(cudaGetLastError());
  RowGemm r{};
  r.R = R; r.T = T; r.taps = 1; r.A = p->cat; r.lda = p->ld_cat; r.K = p->ld_cat; r.B = s.wr; r.C = logits; r.ldc = s.n_out;
  r.N = s.n_out; r.bias = s.br; r.epi = EPI_LINEAR;
  RC(launch_row_gemm(r, false, st));
  return CER_OK;
}

static int head_train_backward_impl(cer_head_train* p, const float* const* feats, const float* dlogits, const float* const* dz,
                                    void* stream) {
  if (!p || !feats) return set_error(CER_ERR_INVALID, "cer_head_train_backward: bad argument");
  const bool tcn_only = dz != nullptr;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const cer_head_train_spec& s = p->s;
  const int R = p->R, T = p->T, k = s.kernel_size, E = p->E, c0 = p->ld_cat - E, sms = p->num_sms;
  const uint32_t seed = p->seed;
  if (!tcn_only) CER_CUDA(cudaMemsetAsync(s.grad_flat, 0, (size_t)s.grad_count * 4, st));
  CER_CUDA(cudaMemsetAsync(p->scratch_zero, 0, p->scratch_zero_bytes, st));
  if (!tcn_only) {

  // classifier: logits = cat Wr^T + br
  { WGrad w{}; w.G = dlogits; w.ldg = s.n_out; w.A = p->cat; w.lda = p->ld_cat; w.dW = s.dwr; w.R = R; w.T = T; w.N = s.n_out;
    w.K = p->ld_cat; w.taps = 1; RC(launch_wgrad(w, sms, st)); }
  RC(launch_colsum(dlogits, s.n_out, R, s.n_out, s.dbr, st));
  { RowGemm g{}; g.R = R; g.T = T; g.taps = 1; g.A = dlogits; g.lda = s.n_out; g.K = s.n_out; g.B = s.wr; g.C = p->gcat;
    g.ldc = p->ld_cat; g.N = p->ld_cat; g.epi = EPI_LINEAR; RC(launch_row_gemm(g, true, st)); }
  // LayerNorm + dropout
  { const int rpb = std::max(8, (R + 4 * sms - 1) / (4 * sms));
    ln_drop_bwd_kernel<<<(R + rpb - 1) / rpb, 256, 0, st>>>(p->gcat + c0, p->ld_cat, p->o, R, E, s.ln_g, p->ln_mean, p->ln_rstd,
                                                             p->go, s.dln_g, s.dln_b, make_drop(s.p_fusion, seed, 4096), rpb);
    CER_CUDA(cudaGetLastError()); }
  // o_proj
  { WGrad w{}; w.G = p->go; w.ldg = E; w.A = p->vals; w.lda = E; w.dW = s.dwo; w.R = R; w.T = T; w.N = E; w.K = E; w.taps = 1;
    RC(launch_wgrad(w, sms, st)); }
  RC(launch_colsum(p->go, E, R, E, s.dbo, st));
  { RowGemm g{}; g.R = R; g.T = T; g.taps = 1; g.A = p->go; g.lda = E; g.K = E; g.B = s.wo; g.C = p->gvals; g.ldc = E; g.N = E;
    g.epi = EPI_LINEAR; RC(launch_row_gemm(g, true, st)); }
  // attention
  { AttnArgs a{};
    for (int m = 0; m < s.n_modals; ++m) { a.qkv[m] = p->mod[m].qkv; a.gqkv[m] = p->mod[m].gqkv; }
    a.gvals = p->gvals; a.R = R; a.M = s.n_modals; a.H = s.num_heads; a.hd = s.modal_dim / s.num_heads;
    attn_bwd_kernel<<<(R * a.H + 127) / 128, 128, 0, st>>>(a);
    CER_CUDA(cudaGetLastError()); }
  }   // !tcn_only

  cudaStream_t main_st = st;
  CER_CUDA(cudaEventRecord(p->ev_fork, main_st));
  for (int m = 0; m < s.n_modals; ++m) {
    const cer_train_modal& M = s.modal[m];
    ModalBuf& mb = p->mod[m];
    const int C = M.blocks[M.n_blocks - 1].c_out;
    st = m == 0 ? main_st : p->side[m];
    if (m > 0) CER_CUDA(cudaStreamWaitEvent(st, p->ev_fork, 0));
    const float* gz = mb.ga;        // [R][C]
    if (tcn_only) {
      if (!dz[m]) return set_error(CER_ERR_INVALID, "cer_head_train_tcn_backward: null gradient pointer");
      gz = dz[m];
    } else {
    // qkv projection
    { WGrad w{}; w.G = mb.gqkv; w.ldg = p->md3; w.A = mb.z; w.lda = mb.ld_z; w.dW = M.dwqkv; w.R = R; w.T = T; w.N = p->md3;
      w.K = C; w.taps = 1; RC(launch_wgrad(w, sms, st)); }
    RC(launch_colsum(mb.gqkv, p->md3, R, p->md3, M.dbqkv, st));
    { RowGemm g{}; g.R = R; g.T = T; g.taps = 1; g.A = mb.gqkv; g.lda = p->md3; g.K = p->md3; g.B = M.wqkv; g.C = mb.ga; g.ldc = C;
      g.N = C; g.epi = EPI_LINEAR;
      if (m == 0) { g.addend = p->gcat; g.ld_add = p->ld_cat; }        // the leader also feeds the classifier directly
      RC(launch_row_gemm(g, true, st)); }
    }
    // BatchNorm1d
    float* gy = mb.gb;
    bn_train_bwd_kernel<<<(C + 31) / 32, dim3(32, 8), 0, st>>>(gz, C, mb.blk[M.n_blocks - 1].y, C, R, C, M.bn_w, mb.bn_mean,
                                                                mb.bn_invstd, gy, C, M.dbn_w, M.dbn_b);
    CER_CUDA(cudaGetLastError());
    float* spare = mb.ga;     // gz is dead from here on
    for (int i = M.n_blocks - 1; i >= 0; --i) {
      const cer_train_block& b = M.blocks[i];
      BlockBuf& bb = mb.blk[i];
      const float* x = i == 0 ? feats[m] : mb.blk[i - 1].y;
      float* gpre = mb.gc;
      float* ga2 = mb.gd;
      const long long tot = (long long)R * b.c_out;
      block_bwd_pre_kernel<<<(int)std::min<long long>((tot + 255) / 256, 8LL * sms), 256, 0, st>>>(
          gy, bb.y, bb.h2d, gpre, ga2, tot, make_drop(s.p_tcn, seed, m * 16 + i * 2 + 1));
      CER_CUDA(cudaGetLastError());
      // conv2: bias, weight, input gradients
      RC(launch_colsum(ga2, b.c_out, R, b.c_out, b.conv2.dbias, st));
      { WGrad w{}; w.tf32 = p->tf32; w.G = ga2; w.ldg = b.c_out; w.A = bb.h1d; w.lda = b.c_out; w.dW = bb.c2.dw_eff;
        w.w_tap_stride = (long long)b.c_out * b.c_out; w.R = R; w.T = T; w.N = b.c_out; w.K = b.c_out; w.taps = k;
        w.shift0 = -(k - 1) * b.dilation; w.shift_step = b.dilation; RC(launch_wgrad(w, sms, st)); }
      wn_bwd_kernel<<<b.c_out, 256, 0, st>>>(b.conv2.g, b.conv2.v, bb.c2.dw_eff, bb.c2.inv_norm, b.conv2.dg, b.conv2.dv,
                                             b.c_out, b.c_out, k);
      CER_CUDA(cudaGetLastError());
      float* ga1 = gy;         // gy is dead once gpre/ga2 exist
      { RowGemm g{}; g.tf32 = p->tf32; g.R = R; g.T = T; g.taps = k; g.shift0 = (k - 1) * b.dilation; g.shift_step = -b.dilation;
        g.A = ga2; g.lda = b.c_out; g.K = b.c_out; g.B = bb.c2.w_eff; g.b_tap_stride = (long long)b.c_out * b.c_out;
        g.C = ga1; g.ldc = b.c_out; g.N = b.c_out; g.epi = EPI_DGRAD_ACT; g.aux = bb.h1d; g.ld_aux = b.c_out;
        g.drop = make_drop(s.p_tcn, seed, m * 16 + i * 2 + 0);
        RC(launch_row_gemm(g, true, st)); }
      // conv1: bias, weight gradients
      RC(launch_colsum(ga1, b.c_out, R, b.c_out, b.conv1.dbias, st));
      { WGrad w{}; w.tf32 = p->tf32; w.G = ga1; w.ldg = b.c_out; w.A = x; w.lda = b.c_in; w.dW = bb.c1.dw_eff;
        w.w_tap_stride = (long long)b.c_out * b.c_in; w.R = R; w.T = T; w.N = b.c_out; w.K = b.c_in; w.taps = k;
        w.shift0 = -(k - 1) * b.dilation; w.shift_step = b.dilation; RC(launch_wgrad(w, sms, st)); }
      wn_bwd_kernel<<<b.c_out, 256, 0, st>>>(b.conv1.g, b.conv1.v, bb.c1.dw_eff, bb.c1.inv_norm, b.conv1.dg, b.conv1.dv,
                                             b.c_out, b.c_in, k);
      CER_CUDA(cudaGetLastError());
      // downsample parameters
      if (b.wd) {
        RC(launch_colsum(gpre, b.c_out, R, b.c_out, b.dbd, st));
        WGrad w{}; w.tf32 = p->tf32; w.G = gpre; w.ldg = b.c_out; w.A = x; w.lda = b.c_in; w.dW = b.dwd; w.R = R; w.T = T; w.N = b.c_out;
        w.K = b.c_in; w.taps = 1; RC(launch_wgrad(w, sms, st));
      }
      if (i == 0) break;       // the input features need no gradient
      // gx = dgrad(conv1)(ga1) + residual-branch gradient
      float* gx = spare;
      { RowGemm g{}; g.tf32 = p->tf32; g.R = R; g.T = T; g.taps = k; g.shift0 = (k - 1) * b.dilation; g.shift_step = -b.dilation;
        g.A = ga1; g.lda = b.c_out; g.K = b.c_out; g.B = bb.c1.w_eff; g.b_tap_stride = (long long)b.c_out * b.c_in;
        g.C = gx; g.ldc = b.c_in; g.N = b.c_in; g.epi = EPI_LINEAR;
        if (!b.wd) { g.addend = gpre; g.ld_add = b.c_out; }
        RC(launch_row_gemm(g, true, st)); }
      if (b.wd) {
        RowGemm g{}; g.tf32 = p->tf32; g.R = R; g.T = T; g.taps = 1; g.A = gpre; g.lda = b.c_out; g.K = b.c_out; g.B = b.wd; g.C = gx; g.ldc = b.c_in;
        g.N = b.c_in; g.epi = EPI_LINEAR; g.accumulate = 1;
        RC(launch_row_gemm(g, true, st));
      }
      spare = gy;              // ga1's buffer is free again after the next block's pre-stage reads gx
      gy = gx;
    }
    if (m > 0) CER_CUDA(cudaEventRecord(p->ev_join[m], st));
  }
  for (int m = 1; m < s.n_modals; ++m) CER_CUDA(cudaStreamWaitEvent(main_st, p->ev_join[m], 0));
  return CER_OK;
}

extern "C" int cer_ce_loss(const float* logits, const int64_t* labels, int64_t rows, int32_t n_cls, float* loss_out,
                           float* dlogits_out, void* stream) {
  if (!logits || !labels || !loss_out || rows <= 0 || n_cls <= 0 || rows > (1 << 20))
    return set_error(CER_ERR_INVALID, "cer_ce_loss: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CER_CUDA(cudaMemsetAsync(loss_out, 0, 4, st));
  ce_loss_kernel<<<(int)((rows + 255) / 256), 256, 0, st>>>(logits, reinterpret_cast<const long long*>(labels), (int)rows, n_cls,
                                                            loss_out, dlogits_out);
  CER_CUDA(cudaGetLastError());
  return CER_OK;
}

extern "C" int cer_optimizer_step(int32_t kind, float* params, const float* grads, float* state_m, float* state_v, int64_t n,
                                  float lr, float weight_decay, float beta1_or_momentum, float beta2_or_dampening, float eps,
                                  int32_t nesterov, int32_t step, float grad_scale, void* stream) {
  if (!params || !grads || n <= 0 || kind < 0 || kind > 2 || step < 1) return set_error(CER_ERR_INVALID, "cer_optimizer_step: bad argument");
  if ((kind == 0 && beta1_or_momentum != 0.f && !state_m) || (kind > 0 && (!state_m || !state_v)))
    return set_error(CER_ERR_INVALID, "cer_optimizer_step: missing optimizer state");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float bc1 = 1.f, bc2s = 1.f;
  if (kind > 0) {
    bc1 = (float)(1.0 - pow((double)beta1_or_momentum, (double)step));
    bc2s = (float)sqrt(1.0 - pow((double)beta2_or_dampening, (double)step));
  }
  const int blocks = (int)std::min<int64_t>((n + 255) / 256, 148 * 16);
  optimizer_kernel<<<blocks, 256, 0, st>>>(kind, params, grads, state_m, state_v, (long long)n, lr, weight_decay,
                                           beta1_or_momentum, beta2_or_dampening, eps, nesterov, step, bc1, bc2s, grad_scale);
  CER_CUDA(cudaGetLastError());
  return CER_OK;
}

extern "C" int cer_modal_attention_maps(const float* const* qkv_dev, int64_t rows, int32_t n_modals, int32_t num_heads,
                                        int32_t head_dim, float* maps_out_dev, void* stream) {
  if (!qkv_dev || !maps_out_dev || rows < 0 || n_modals < 1 || n_modals > CER_MAX_MODALS || num_heads < 1 || head_dim < 1 ||
      rows > (1 << 24))
    return set_error(CER_ERR_INVALID, "cer_modal_attention_maps: bad argument");
  if (rows == 0) return CER_OK;
  AttnArgs a{};
  for (int m = 0; m < n_modals; ++m) {
    if (!qkv_dev[m]) return set_error(CER_ERR_INVALID, "cer_modal_attention_maps: null qkv pointer");
    a.qkv[m] = qkv_dev[m];
  }
  a.R = (int)rows; a.M = n_modals; a.H = num_heads; a.hd = head_dim;
  attn_maps_kernel<<<(int)((rows * num_heads + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(a, maps_out_dev);
  CER_CUDA(cudaGetLastError());
  return CER_OK;
}

extern "C" int cer_modal_attention_forward(const float* const* qkv_dev, int64_t rows, int32_t n_modals, int32_t num_heads,
                                           int32_t head_dim, float* vals_out_dev, void* stream) {
  if (!qkv_dev || !vals_out_dev || rows < 0 || n_modals < 1 || n_modals > CER_MAX_MODALS || num_heads < 1 || head_dim < 1 ||
      rows > (1 << 24))
    return set_error(CER_ERR_INVALID, "cer_modal_attention_forward: bad argument");
  if (rows == 0) return CER_OK;
  AttnArgs a{};
  for (int m = 0; m < n_modals; ++m) {
    if (!qkv_dev[m]) return set_error(CER_ERR_INVALID, "cer_modal_attention_forward: null qkv pointer");
    a.qkv[m] = qkv_dev[m];
  }
  a.vals = vals_out_dev; a.R = (int)rows; a.M = n_modals; a.H = num_heads; a.hd = head_dim;
  attn_fwd_kernel<<<(int)((rows * num_heads + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(a);
  CER_CUDA(cudaGetLastError());
  return CER_OK;
}

// ==========================================================================================
// Building blocks WITH backward for training the alternative heads CAN / JMT / MT
// (models/model.py:529-684, :709-750, :895-1167; the reference trains them through the same loop as LFAN,
// experiment.py:317-347).  The host side (heads_training.py) composes them into forward / backward; every GEMM
// runs in the kernels above (row GEMM, weight gradient with fp32 atomics, column sums).
// Gradients of PARAMETERS are accumulated (+=) into buffers the caller zeroes once per step; gradients of
// ACTIVATIONS are written unless stated otherwise.
// ==========================================================================================
namespace {

// dx = dy * act'(.) where the sign of the saved OUTPUT decides (LeakyReLU and ReLU keep the sign of their input)
__global__ void act_bwd_kernel(int act, const float* __restrict__ y, const float* __restrict__ dy, long long n, float* __restrict__ dx) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dx[i] = dy[i] * (y[i] > 0.f ? 1.f : (act == 1 ? kLeaky : 0.f));
}
__global__ void leaky_fwd_kernel(const float* __restrict__ x, long long n, float* __restrict__ y) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) y[i] = lrelu(x[i]);
}

// y = softmax(gate) * feat  =>  dfeat = dy * s,  dgate = s * (t - sum(t * s)),  t = dy * feat   (one warp per row)
__global__ void __launch_bounds__(256) softmax_gate_bwd_kernel(const float* __restrict__ gate, const float* __restrict__ feat,
                                                               const float* __restrict__ dy, int rows, int dim,
                                                               float* __restrict__ dgate, float* __restrict__ dfeat) {
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= rows) return;
  const long long o = (long long)r * dim;
  float mx = -3.4e38f;
  for (int i = lane; i < dim; i += 32) mx = fmaxf(mx, gate[o + i]);
#pragma unroll
  for (int k = 16; k > 0; k >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, k));
  float den = 0.f;
  for (int i = lane; i < dim; i += 32) den += expf(gate[o + i] - mx);
#pragma unroll
  for (int k = 16; k > 0; k >>= 1) den += __shfl_xor_sync(0xffffffffu, den, k);
  const float inv = 1.f / den;
  float dot = 0.f;
  for (int i = lane; i < dim; i += 32) dot += dy[o + i] * feat[o + i] * expf(gate[o + i] - mx) * inv;
#pragma unroll
  for (int k = 16; k > 0; k >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, k);
  for (int i = lane; i < dim; i += 32) {
    const float sm = expf(gate[o + i] - mx) * inv;
    dfeat[o + i] = dy[o + i] * sm;
    dgate[o + i] = sm * (dy[o + i] * feat[o + i] - dot);
  }
}

// y = LayerNorm(x + res) * gamma + beta  =>  dx (= dres), dgamma +=, dbeta +=.  One warp per row, lanes own the columns
// lane, lane + 32, ... (dim <= 512); a block walks `rpb` rows and adds its column sums with one atomic per column.
constexpr int kLnMaxCols = 16;
__global__ void __launch_bounds__(256) add_ln_bwd_kernel(const float* __restrict__ x, const float* __restrict__ res, int rows, int dim,
                                                         const float* __restrict__ gamma, float eps, const float* __restrict__ dy,
                                                         float* __restrict__ dx, float* __restrict__ dgamma,
                                                         float* __restrict__ dbeta, int rpb) {
  __shared__ float s_g[512], s_b[512];
  for (int i = threadIdx.x; i < dim; i += blockDim.x) { s_g[i] = 0.f; s_b[i] = 0.f; }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float ag[kLnMaxCols], ab[kLnMaxCols];
#pragma unroll
  for (int j = 0; j < kLnMaxCols; ++j) { ag[j] = 0.f; ab[j] = 0.f; }
  const int r_end = min(rows, (blockIdx.x + 1) * rpb);
  for (int r = blockIdx.x * rpb + warp; r < r_end; r += 8) {
    const long long o = (long long)r * dim;
    float h[kLnMaxCols], g[kLnMaxCols];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < kLnMaxCols; ++j) {
      const int c = lane + 32 * j;
      h[j] = c < dim ? x[o + c] + (res ? res[o + c] : 0.f) : 0.f;
      s += h[j];
    }
#pragma unroll
    for (int k = 16; k > 0; k >>= 1) s += __shfl_xor_sync(0xffffffffu, s, k);
    const float mean = s / dim;
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < kLnMaxCols; ++j) { const int c = lane + 32 * j; if (c < dim) { const float d = h[j] - mean; ss += d * d; } }
#pragma unroll
    for (int k = 16; k > 0; k >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, k);
    const float rstd = 1.f / sqrtf(ss / dim + eps);
    float sg = 0.f, sgx = 0.f;
#pragma unroll
    for (int j = 0; j < kLnMaxCols; ++j) {
      const int c = lane + 32 * j;
      if (c < dim) {
        const float xh = (h[j] - mean) * rstd, d = dy[o + c];
        h[j] = xh;
        g[j] = d * gamma[c];
        sg += g[j]; sgx += g[j] * xh;
        ag[j] += d * xh; ab[j] += d;
      } else { g[j] = 0.f; }
    }
#pragma unroll
    for (int k = 16; k > 0; k >>= 1) { sg += __shfl_xor_sync(0xffffffffu, sg, k); sgx += __shfl_xor_sync(0xffffffffu, sgx, k); }
    const float mg = sg / dim, mgx = sgx / dim;
#pragma unroll
    for (int j = 0; j < kLnMaxCols; ++j) { const int c = lane + 32 * j; if (c < dim) dx[o + c] = rstd * (g[j] - mg - h[j] * mgx); }
  }
#pragma unroll
  for (int j = 0; j < kLnMaxCols; ++j) { const int c = lane + 32 * j; if (c < dim) { atomicAdd(&s_g[c], ag[j]); atomicAdd(&s_b[c], ab[j]); } }
  __syncthreads();
  for (int i = threadIdx.x; i < dim; i += blockDim.x) { atomicAdd(dgamma + i, s_g[i]); atomicAdd(dbeta + i, s_b[i]); }
}

// P = softmax(S * scale) row-wise, in place; one warp per row
__global__ void __launch_bounds__(256) softmax_rows_kernel(float* __restrict__ S, int rows, int cols, int ld, float scale) {
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= rows) return;
  float* p = S + (long long)r * ld;
  float mx = -3.4e38f;
  for (int i = lane; i < cols; i += 32) mx = fmaxf(mx, p[i] * scale);
#pragma unroll
  for (int k = 16; k > 0; k >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, k));
  float den = 0.f;
  for (int i = lane; i < cols; i += 32) { const float e = expf(p[i] * scale - mx); p[i] = e; den += e; }
#pragma unroll
  for (int k = 16; k > 0; k >>= 1) den += __shfl_xor_sync(0xffffffffu, den, k);
  const float inv = 1.f / den;
  for (int i = lane; i < cols; i += 32) p[i] *= inv;
}
// dS = P * (dP - sum(dP * P)) * scale row-wise, written over dP
__global__ void __launch_bounds__(256) softmax_bwd_rows_kernel(const float* __restrict__ P, float* __restrict__ dP, int rows, int cols,
                                                               int ld, float scale) {
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= rows) return;
  const float* p = P + (long long)r * ld;
  float* d = dP + (long long)r * ld;
  float dot = 0.f;
  for (int i = lane; i < cols; i += 32) dot += d[i] * p[i];
#pragma unroll
  for (int k = 16; k > 0; k >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, k);
  for (int i = lane; i < cols; i += 32) d[i] = p[i] * (d[i] - dot) * scale;
}

int block_sms() {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms;
}

}  // namespace

extern "C" int cer_linear_backward(const float* x, int64_t rows, int32_t in_dim, int32_t ldx, const float* w, const float* dy,
                                   int32_t out_dim, int32_t lddy, float* dx, int32_t lddx, int32_t dx_accumulate, float* dw, float* db,
                                   void* stream) {
  if (!dy || rows <= 0 || in_dim <= 0 || out_dim <= 0 || lddy < out_dim || rows > (1 << 24) || (dx && (!w || lddx < in_dim)) ||
      (dw && (!x || ldx < in_dim)))
    return set_error(CER_ERR_INVALID, "cer_linear_backward: bad argument");
  int rc = cer_check_device();
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dx) {      // dx = dy W  (W [out][in] read as [K = out][N = in])
    RowGemm g{}; g.R = (int)rows; g.T = (int)rows; g.taps = 1; g.A = dy; g.lda = lddy; g.K = out_dim; g.B = w; g.C = dx; g.ldc = lddx;
    g.N = in_dim; g.epi = EPI_LINEAR; g.accumulate = dx_accumulate ? 1 : 0;
    RC(launch_row_gemm(g, true, st));
  }
  if (dw) {      // dw[out][in] += sum_r dy[r][out] x[r][in]
    WGrad wg{}; wg.G = dy; wg.ldg = lddy; wg.A = x; wg.lda = ldx; wg.dW = dw; wg.R = (int)rows; wg.T = (int)rows; wg.N = out_dim;
    wg.K = in_dim; wg.taps = 1;
    RC(launch_wgrad(wg, block_sms(), st));
  }
  if (db) RC(launch_colsum(dy, lddy, (int)rows, out_dim, db, st));
  return CER_OK;
}

extern "C" int cer_act_backward(int32_t act, const float* y, const float* dy, int64_t n, float* dx, void* stream) {
  if (!y || !dy || !dx || n < 0 || (act != 1 && act != 2)) return set_error(CER_ERR_INVALID, "cer_act_backward: bad argument");
  if (n == 0) return CER_OK;
  act_bwd_kernel<<<(int)std::min<int64_t>((n + 255) / 256, 148 * 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(act, y, dy, n, dx);
  CER_CUDA(cudaGetLastError());
  return CER_OK;
}

extern "C" int cer_leaky_relu_forward(const float* x, int64_t n, float* y, void* stream) {
  if (!x || !y || n < 0) return set_error(CER_ERR_INVALID, "cer_leaky_relu_forward: bad argument");
  if (n == 0) return CER_OK;
  leaky_fwd_kernel<<<(int)std::min<int64_t>((n + 255) / 256, 148 * 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, n, y);
  CER_CUDA(cudaGetLastError());
  return CER_OK;
}

extern "C" int cer_softmax_gate_backward(const float* gate, const float* feat, const float* dy, int64_t rows, int32_t dim, float* dgate,
                                         float* dfeat, void* stream) {
  if (!gate || !feat || !dy || !dgate || !dfeat || rows < 0 || dim <= 0 || rows > (1 << 30))
    return set_error(CER_ERR_INVALID, "cer_softmax_gate_backward: bad argument");
  if (rows == 0) return CER_OK;
  softmax_gate_bwd_kernel<<<(int)((rows + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(gate, feat, dy, (int)rows, dim, dgate, dfeat);
  CER_CUDA(cudaGetLastError());
  return CER_OK;
}

extern "C" int cer_add_layernorm_backward(const float* x, const float* res, int64_t rows, int32_t dim, const float* gamma, float eps,
                                          const float* dy, float* dx, float* dgamma, float* dbeta, void* stream) {
  if (!x || !gamma || !dy || !dx || !dgamma || !dbeta || rows <= 0 || dim <= 0 || dim > 32 * kLnMaxCols || rows > (1 << 30))
    return set_error(CER_ERR_INVALID, "cer_add_layernorm_backward: bad argument (dim <= 512)");
  const int sms = block_sms();
  const int rpb = (int)std::max<int64_t>(8, (rows + 4 * sms - 1) / (4 * sms));
  add_ln_bwd_kernel<<<(int)((rows + rpb - 1) / rpb), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, res, (int)rows, dim, gamma, eps, dy, dx,
                                                                                             dgamma, dbeta, rpb);
  CER_CUDA(cudaGetLastError());
  return CER_OK;
}

extern "C" int cer_bn1d_train_forward(const float* x, int64_t rows, int32_t c, const float* w, const float* b, float* y, float* save_mean,
                                      float* save_invstd, float* running_mean, float* running_var, float momentum, void* stream) {
  if (!x || !w || !b || !y || !save_mean || !save_invstd || !running_mean || !running_var || rows <= 0 || c <= 0 || rows > (1 << 24))
    return set_error(CER_ERR_INVALID, "cer_bn1d_train_forward: bad argument");
  bn_train_fwd_kernel<<<(c + 31) / 32, dim3(32, 8), 0, static_cast<cudaStream_t>(stream)>>>(x, c, (int)rows, c, w, b, y, c, save_mean,
                                                                                        save_invstd, running_mean, running_var, momentum);
  CER_CUDA(cudaGetLastError());
  return CER_OK;
}

extern "C" int cer_bn1d_train_backward(const float* dy, const float* x, int64_t rows, int32_t c, const float* w, const float* save_mean,
                                       const float* save_invstd, float* dx, float* dw, float* db, void* stream) {
  if (!dy || !x || !w || !save_mean || !save_invstd || !dx || !dw || !db || rows <= 0 || c <= 0 || rows > (1 << 24))
    return set_error(CER_ERR_INVALID, "cer_bn1d_train_backward: bad argument");
  bn_train_bwd_kernel<<<(c + 31) / 32, dim3(32, 8), 0, static_cast<cudaStream_t>(stream)>>>(dy, c, x, c, (int)rows, c, w, save_mean, save_invstd,
                                                                                        dx, c, dw, db);
  CER_CUDA(cudaGetLastError());
  return CER_OK;
}

// Single-head attention with the probabilities kept for backward (exact fp32 GEMMs):
//   probs[b] = softmax(q[b] k[b]^T / sqrt(dim))  [len_q][len_k],   out[b] = probs[b] v[b].
extern "C" int cer_sdpa_train_forward(const float* q, int32_t ldq, const float* k, int32_t ldk, const float* v, int32_t ldv, int32_t batch,
                                      int32_t len_q, int32_t len_k, int32_t dim, float* out, int32_t ldo, float* probs, void* stream) {
  if (!q || !k || !v || !out || !probs || batch <= 0 || len_q <= 0 || len_k <= 0 || dim <= 0 || ldq < dim || ldk < dim || ldv < dim || ldo < dim)
    return set_error(CER_ERR_INVALID, "cer_sdpa_train_forward: bad argument");
  int rc = cer_check_device();
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const float scale = 1.f / sqrtf((float)dim);
  for (int b = 0; b < batch; ++b) {
    float* P = probs + (size_t)b * len_q * len_k;
    { RowGemm g{}; g.R = len_q; g.T = len_q; g.taps = 1; g.A = q + (size_t)b * len_q * ldq; g.lda = ldq; g.K = dim;
      g.B = k + (size_t)b * len_k * ldk; g.ldb = ldk; g.C = P; g.ldc = len_k; g.N = len_k; g.epi = EPI_LINEAR;
      RC(launch_row_gemm(g, false, st)); }
    softmax_rows_kernel<<<(len_q + 7) / 8, 256, 0, st>>>(P, len_q, len_k, len_k, scale);
    CER_CUDA(cudaGetLastError());
    { RowGemm g{}; g.R = len_q; g.T = len_q; g.taps = 1; g.A = P; g.lda = len_k; g.K = len_k; g.B = v + (size_t)b * len_k * ldv;
      g.ldb = ldv; g.C = out + (size_t)b * len_q * ldo; g.ldc = ldo; g.N = dim; g.epi = EPI_LINEAR;
      RC(launch_row_gemm(g, true, st)); }
  }
  return CER_OK;
}

// Backward of cer_sdpa_train_forward.  dq is written; dk / dv are ACCUMULATED (zero them first, e.g. as part of a packed
// d(qkv) buffer); scratch: [batch][len_q][len_k] floats.
extern "C" int cer_sdpa_backward(const float* q, int32_t ldq, const float* k, int32_t ldk, const float* v, int32_t ldv, const float* probs,
                                 const float* dout, int32_t lddo, int32_t batch, int32_t len_q, int32_t len_k, int32_t dim, float* dq,
                                 int32_t lddq, float* dk, int32_t lddk, float* dv, int32_t lddv, float* scratch, void* stream) {
  if (!q || !k || !v || !probs || !dout || !dq || !dk || !dv || !scratch || batch <= 0 || len_q <= 0 || len_k <= 0 || dim <= 0)
    return set_error(CER_ERR_INVALID, "cer_sdpa_backward: bad argument");
  int rc = cer_check_device();
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int sms = block_sms();
  const float scale = 1.f / sqrtf((float)dim);
  for (int b = 0; b < batch; ++b) {
    const float* P = probs + (size_t)b * len_q * len_k;
    float* dS = scratch + (size_t)b * len_q * len_k;
    const float* qb = q + (size_t)b * len_q * ldq;
    const float* kb = k + (size_t)b * len_k * ldk;
    const float* vb = v + (size_t)b * len_k * ldv;
    const float* dob = dout + (size_t)b * len_q * lddo;
    // dv[key][e] += sum_q P[q][key] dout[q][e]
    { WGrad w{}; w.G = P; w.ldg = len_k; w.A = dob; w.lda = lddo; w.dW = dv + (size_t)b * len_k * lddv; w.ldw = lddv; w.R = len_q;
      w.T = len_q; w.N = len_k; w.K = dim; w.taps = 1; RC(launch_wgrad(w, sms, st)); }
    // dP = dout v^T
    { RowGemm g{}; g.R = len_q; g.T = len_q; g.taps = 1; g.A = dob; g.lda = lddo; g.K = dim; g.B = vb; g.ldb = ldv; g.C = dS;
      g.ldc = len_k; g.N = len_k; g.epi = EPI_LINEAR; RC(launch_row_gemm(g, false, st)); }
    softmax_bwd_rows_kernel<<<(len_q + 7) / 8, 256, 0, st>>>(P, dS, len_q, len_k, len_k, scale);
    CER_CUDA(cudaGetLastError());
    // dq = dS k
    { RowGemm g{}; g.R = len_q; g.T = len_q; g.taps = 1; g.A = dS; g.lda = len_k; g.K = len_k; g.B = kb; g.ldb = ldk;
      g.C = dq + (size_t)b * len_q * lddq; g.ldc = lddq; g.N = dim; g.epi = EPI_LINEAR; RC(launch_row_gemm(g, true, st)); }
    // dk[key][e] += sum_q dS[q][key] q[q][e]
    { WGrad w{}; w.G = dS; w.ldg = len_k; w.A = qb; w.lda = ldq; w.dW = dk + (size_t)b * len_k * lddk; w.ldw = lddk; w.R = len_q;
      w.T = len_q; w.N = len_k; w.K = dim; w.taps = 1; RC(launch_wgrad(w, sms, st)); }
  }
  return CER_OK;
}
