// Microbenchmark: cycles per TMA load for the A-operand fetch patterns considered for the conv
// kernel (im2col 128px x 64ch, tiled 3-D row boxes, tiled 2-D) -- no MMA, just producer/consumer.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tma_probe tools/tma_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../feature_vs_text_compound_emotion_b200/csrc/ptx.cuh"

using namespace cer;

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*PFN_encodeIm2col)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                     const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t,
                                     const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ void tma_load_3d(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

struct alignas(64) Params {
  CUtensorMap map;
  int mode;       // 0 im2col, 1 tiled3d rows, 2 tiled2d
  int iters;
  int bytes;      // per load
  int W, H, rows; // geometry
  int nprod;      // producer threads (1 or 2)
  int loads_per_stage;
  long long* cycles;
};

constexpr int STAGES = 6;

__global__ void __launch_bounds__(256, 1) probe(const __grid_constant__ Params p) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * 16384 * 2);
  uint64_t* empty = full + STAGES;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    fence_barrier_init();
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long t0 = clock64();
  if (warp < p.nprod && lane == 0) {
    // producer `warp` handles stages s with s % nprod == warp
    int cw = warp, cn = warp * 5; const int nbase = (blockIdx.x & 3) * 64;
    for (int i = warp; i < p.iters; i += p.nprod) {
      const int stage = i % STAGES;
      const uint32_t phase = (i / STAGES) & 1;
      mbar_wait(&empty[stage], phase ^ 1);
      mbar_expect_tx(&full[stage], p.bytes * p.loads_per_stage);
      for (int l = 0; l < p.loads_per_stage; ++l) {
        uint8_t* dst = smem + stage * 32768 + l * 16384;
        // cheap address arithmetic only: adds and masks
        cw += 1; if (cw >= 3) cw = 0;
        cn = (cn + 1) & 63;
        if (p.mode == 0) {
          tma_load_im2col_4d(&p.map, &full[stage], dst, 0, cw - 1, cw, cn + nbase, (uint16_t)cw, (uint16_t)(cn & 1));
        } else if (p.mode == 1) {
          tma_load_3d(&p.map, &full[stage], dst, 0, cw - 1, cn * p.rows + nbase * p.rows);
        } else {
          tma_load_2d(&p.map, &full[stage], dst, 0, cn * 128 + nbase * 128);
        }
      }
    }
  } else if (warp == 7 && lane == 0) {
    for (int i = 0; i < p.iters; ++i) {
      const int stage = i % STAGES;
      const uint32_t phase = (i / STAGES) & 1;
      mbar_wait(&full[stage], phase);
      mbar_arrive(&empty[stage]);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) p.cycles[blockIdx.x] = clock64() - t0;
}

int main() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaFree(0);
  cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &fn, 12000, cudaEnableDefault, &q);
  PFN_encodeTiled enc_tiled = (PFN_encodeTiled)fn;
  cudaGetDriverEntryPointByVersion("cuTensorMapEncodeIm2col", &fn, 12000, cudaEnableDefault, &q);
  PFN_encodeIm2col enc_im2col = (PFN_encodeIm2col)fn;

  const size_t bytes = 64ull << 20;   // 64 MB tensor: L2 resident after first touch
  void* buf;
  cudaMalloc(&buf, bytes);
  cudaMemset(buf, 0, bytes);
  long long* dcyc;
  cudaMalloc(&dcyc, 148 * 8);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, STAGES * 32768 + 1024 + 256);

  struct Case { const char* name; int mode, W, H, C, rows, nprod, lps; };
  std::vector<Case> cases = {
      {"im2col 128px x 64ch, 10x10x256, 1 producer", 0, 10, 10, 256, 0, 1, 1},
      {"im2col 128px x 64ch, 10x10x256, 2 producers", 0, 10, 10, 256, 0, 2, 1},
      {"im2col 128px x 64ch, 10x10x256, 4 producers", 0, 10, 10, 256, 0, 4, 1},
      {"im2col 128px x 64ch, 40x40x64,  1 producer", 0, 40, 40, 64, 0, 1, 1},
      {"im2col 2 loads/stage, 10x10x256", 0, 10, 10, 256, 0, 1, 2},
      {"tiled3d box[64,10,12] (120 rows), 1 producer", 1, 10, 0, 256, 12, 1, 1},
      {"tiled3d box[64,10,12] (120 rows), 2 producers", 1, 10, 0, 256, 12, 2, 1},
      {"tiled3d box[64,40,3] (120 rows), 1 producer", 1, 40, 0, 64, 3, 1, 1},
      {"tiled2d box[64,128], 1 producer", 2, 0, 0, 256, 0, 1, 1},
      {"tiled2d box[64,128], 2 producers", 2, 0, 0, 256, 0, 2, 1},
      {"tiled2d box[64,128], 4 producers", 2, 0, 0, 256, 0, 4, 1},
      {"tiled2d box[64,128], 2 loads/stage", 2, 0, 0, 256, 0, 1, 2},
  };
  for (auto& c : cases) {
    Params p{};
    p.mode = c.mode; p.iters = 4000; p.W = c.W; p.H = c.H; p.rows = c.rows; p.nprod = c.nprod; p.cycles = dcyc;
    p.loads_per_stage = c.lps;
    CUresult r;
    if (c.mode == 0) {
      const int N = 256;
      cuuint64_t dims[4] = {(cuuint64_t)c.C, (cuuint64_t)c.W, (cuuint64_t)c.H, (cuuint64_t)N};
      cuuint64_t strides[3] = {(cuuint64_t)c.C * 2, (cuuint64_t)c.W * c.C * 2, (cuuint64_t)c.H * c.W * c.C * 2};
      int lower[2] = {-1, -1}, upper[2] = {-1, -1};
      cuuint32_t estr[4] = {1, 1, 1, 1};
      r = enc_im2col(&p.map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, buf, dims, strides, lower, upper, 64, 128, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      p.bytes = 16384;
    } else if (c.mode == 1) {
      const int R = 4096;
      cuuint64_t dims[3] = {(cuuint64_t)c.C, (cuuint64_t)c.W, (cuuint64_t)R};
      cuuint64_t strides[2] = {(cuuint64_t)c.C * 2, (cuuint64_t)c.W * c.C * 2};
      cuuint32_t box[3] = {64, (cuuint32_t)c.W, (cuuint32_t)c.rows};
      cuuint32_t estr[3] = {1, 1, 1};
      r = enc_tiled(&p.map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, buf, dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      p.bytes = 64 * c.W * c.rows * 2;
    } else {
      cuuint64_t dims[2] = {(cuuint64_t)c.C, (cuuint64_t)(bytes / (c.C * 2))};
      cuuint64_t strides[1] = {(cuuint64_t)c.C * 2};
      cuuint32_t box[2] = {64, 128};
      cuuint32_t estr[2] = {1, 1};
      r = enc_tiled(&p.map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      p.bytes = 16384;
    }
    if (r != CUDA_SUCCESS) { printf("%s: encode failed %d\n", c.name, (int)r); continue; }
    for (int grid : {1, 148}) {
      probe<<<grid, 256, STAGES * 32768 + 1024 + 256>>>(p);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%s: %s\n", c.name, cudaGetErrorString(e)); return 1; }
      probe<<<grid, 256, STAGES * 32768 + 1024 + 256>>>(p);
      cudaDeviceSynchronize();
      std::vector<long long> h(148);
      cudaMemcpy(h.data(), dcyc, grid * 8, cudaMemcpyDeviceToHost);
      long long mx = 0;
      for (int i = 0; i < grid; ++i) mx = std::max(mx, h[i]);
      const double per = (double)mx / (p.iters * c.lps);
      printf("%-52s grid %3d: %7.1f cycles/load  %6.2f B/cycle/SM\n", c.name, grid, per, p.bytes / per);
    }
  }
  return 0;
}
