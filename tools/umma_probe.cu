// Microbenchmark: cycles per tcgen05.mma (kind::f16, bf16 operands, both from shared memory, 128-byte
// swizzle) as a function of N, cta_group and the A descriptor's group stride -- what the shared-memory
// operand path costs for the narrow (N = 64) layers of IR-50 stage 1.
//   cg = 1: M = 128 per CTA.   cg = 2: M = 256 per CTA pair, each CTA holds its 128 rows of A and N/2 rows of B.
// Every SM runs the same loop (148 CTAs / 74 pairs) so the numbers include chip-level effects.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/umma_probe tools/umma_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../feature_vs_text_compound_emotion_b200/csrc/conv_igemm2.cuh"

using namespace cer;

struct Params {
  int n;          // MMA N
  int iters;      // k-steps (each = 4 MMAs of K = 16)
  int stages;     // distinct A/B smem slots walked round-robin
  int sbo;        // A descriptor group stride in bytes (1024 dense, 1280 halo slab)
  int a_step;     // bytes between the A starts of consecutive k-steps inside a slot (halo taps: 128)
  int layout;     // descriptor swizzle mode: 2 = 128 B rows, 4 = 64 B rows, 6 = 32 B rows (K-major)
  int nacc;       // accumulators the MMAs rotate over (1 = every MMA depends on the previous one's accumulator)
  long long* cycles;
};

template <int CG>
__global__ void __launch_bounds__(128, 1) probe(const Params p) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t done_bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0;
  if (threadIdx.x == 0) { mbar_init(&done_bar, 1); fence_barrier_init(); }
  // operands: finite bf16 garbage
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + (i & 0xff);
  fence_proxy_async();
  if (warp == 0) {
    if (CG == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      tmem_alloc(&tmem_slot, 512);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t base = smem_u32(smem);
  constexpr int kSlot = 24 * 1024;                      // A slot (a 23 KB halo slab fits)
  const uint32_t b_base = base + p.stages * kSlot;
  const int b_bytes = (p.n / CG) * 128;
  long long t0 = 0, t1 = 0;
  if (warp == 0 && rank == 0) {
    const uint32_t idesc = umma_idesc(128 * CG, p.n, 1);
    // row bytes of the swizzle mode; a K = 16 slice is 32 B: slices_per_row of them share a row, the next
    // group of slices starts a new [rows x row_bytes] tile
    const int row_bytes = p.layout == 2 ? 128 : (p.layout == 4 ? 64 : 32);
    const int spr = row_bytes / 32;
    const uint32_t sbo = p.sbo != 1024 ? (uint32_t)p.sbo : 8u * row_bytes;
    const uint32_t hi_a = (sbo >> 4) | (1u << 14) | ((uint32_t)p.layout << 29);
    const uint32_t hi_b = ((8u * row_bytes) >> 4) | (1u << 14) | ((uint32_t)p.layout << 29);
    const uint32_t a_tile = (128 * row_bytes) >> 4, b_tile = ((p.n / CG) * row_bytes) >> 4;
    // all descriptors of one round (4 stages x 4 k-slices) are built BEFORE the timed loop: the loop body is
    // nothing but 16 tcgen05.mma with register operands (an issue loop with address arithmetic in it measured
    // 84-111 cycles per MMA whatever N was: the issuing thread, not the tensor pipe)
    uint64_t ad[16], bd[16];
    uint32_t td[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int s = (i >> 2) % p.stages, k = i & 3;
      const uint32_t a_lo = umma_desc_lo(base + s * kSlot + (i >> 2) * p.a_step);
      const uint32_t b_lo = umma_desc_lo(b_base + (s & 1) * b_bytes);
      ad[i] = (static_cast<uint64_t>(hi_a) << 32) | (a_lo + (k / spr) * a_tile + (k % spr) * 2);
      bd[i] = (static_cast<uint64_t>(hi_b) << 32) | (b_lo + (k / spr) * b_tile + (k % spr) * 2);
      td[i] = tmem + (i % p.nacc) * p.n;
    }
    t0 = clock64();
    for (int it = 0; it < p.iters; it += 4) {
      if (elect_one()) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          if (CG == 2) umma2_f16(td[i], ad[i], bd[i], idesc, 1u); else umma_f16(td[i], ad[i], bd[i], idesc, 1u);
        }
      }
      __syncwarp();
    }
    if (elect_one()) {
      if (CG == 2) umma2_commit_mc(smem_u32(&done_bar)); else umma_commit_a(smem_u32(&done_bar));
    }
    __syncwarp();
  }
  if (warp == 0) {            // both CTAs of a pair wait: the peer's shared memory is read until the last MMA retires
    mbar_wait_a(smem_u32(&done_bar), 0);
    t1 = clock64();
    if (rank == 0 && (threadIdx.x & 31) == 0) p.cycles[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  if (warp == 0) {
    tc_fence_after();
    if (CG == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    else tmem_dealloc(tmem, 512);
  }
}

template <int CG>
static double run(int n, int stages, int sbo, int a_step, int nacc, int layout, int iters, int sms) {
  long long* d;
  cudaMalloc(&d, sizeof(long long) * sms);
  cudaMemset(d, 0, sizeof(long long) * sms);
  Params p{n, iters, stages, sbo, a_step, layout, nacc, d};
  const size_t smem = 205 * 1024;
  cudaFuncSetAttribute(probe<CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(CG == 2 ? (sms / 2) * 2 : sms);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CG; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  for (int rep = 0; rep < 2; ++rep) {
    cudaError_t e = cudaLaunchKernelEx(&cfg, probe<CG>, p);
    if (e != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(e)); return -1; }
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return -1; }
  }
  std::vector<long long> h(sms);
  cudaMemcpy(h.data(), d, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
  cudaFree(d);
  double mx = 0;
  for (int i = 0; i < (int)cfg.gridDim.x; i += CG) mx = h[i] > mx ? (double)h[i] : mx;
  return mx / (iters * 4.0);
}

int main() {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int iters = 4096;
  printf("cycles per tcgen05.mma (K = 16, bf16, SS mode), max over CTAs, %d k-steps of 4 MMAs\n", iters);
  printf("%-4s %-5s %-7s %-6s %-5s %-7s %10s %12s %12s\n", "cg", "N", "stages", "sbo", "nacc", "swizzle", "cyc/MMA", "floor", "B/clk/SM");
  for (int cg = 1; cg <= 2; ++cg)
    for (int n : {64, 128, 256})
      for (int mode = 0; mode < 6; ++mode) {
        const int stages = mode == 0 ? 1 : 4;
        const int sbo = mode == 2 ? 1280 : 1024, a_step = mode == 2 ? 128 : 0;
        const int nacc = mode == 3 ? 2 : 1;
        const int layout = mode == 4 ? 4 : (mode == 5 ? 6 : 2);
        if (nacc * n > 512) continue;
        const double c = cg == 1 ? run<1>(n, stages, sbo, a_step, nacc, layout, iters, sms) : run<2>(n, stages, sbo, a_step, nacc, layout, iters, sms);
        const double floor_c = 128.0 * n / 256.0;            // tensor-pipe floor per MMA: max(M,128) * N / (256 * cg) cycles for M = 128 * cg
        const double bytes = 128 * 32 + (n / cg) * 32;       // smem bytes one SM reads per MMA
        printf("%-4d %-5d %-7d %-6d %-5d %-7d %10.1f %12.1f %12.1f\n", cg, n, stages, sbo, nacc, layout == 2 ? 128 : (layout == 4 ? 64 : 32), c, floor_c, bytes / c);
      }
  return 0;
}
