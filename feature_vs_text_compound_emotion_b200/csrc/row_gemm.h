// Row GEMM with taps (fp32, CUDA cores, packed FFMA2) shared by the training plan (train.cu) and
// the generic linear entry points (heads.cu).  Kernel and launcher are defined in train.cu.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace cer {

__host__ __device__ __forceinline__ uint32_t fmix32(uint32_t x) {
  x ^= x >> 16; x *= 0x85EBCA6Bu; x ^= x >> 13; x *= 0xC2B2AE35u; x ^= x >> 16;
  return x;
}
struct Drop {            // p == 0 <=> thr == 0 (everything kept, scale 1)
  uint32_t key;          // seed + stream * 0x85EBCA6B
  uint32_t thr;          // keep iff hash >= thr
  float scale;           // 1 / (1 - p)
};
enum Epi { EPI_LINEAR = 0, EPI_LRELU_DROP = 1, EPI_BLOCK_OUT = 2, EPI_DGRAD_ACT = 3, EPI_RELU = 4 };

struct RowGemm {
  const float* A; int lda;
  const float* B; long long b_tap_stride;   // elements between consecutive taps' matrices
  float* C; int ldc;
  int R, T, N, K;
  int taps, shift0, shift_step;             // shift_j = shift0 + j*shift_step (rows)
  const float* bias;                        // [N] or null
  const float* addend; int ld_add;          // [R][N] added before the activation, or null
  int accumulate;                           // C += result (EPI_LINEAR only)
  int epi;
  float* aux; int ld_aux;                   // BLOCK_OUT: h2d out; DGRAD_ACT: saved activation in
  Drop drop;
  int tf32;                                 // 1: TF32 tensor-core kernel (mma.sync, fp32 accumulate); 0: exact fp32 FFMA2
  int ldb;                                  // row pitch of B (0 = dense: K for [N][K] storage, N for [K][N])
};


// b_kn: B_j stored [K][N] (true) or [N][K] (false, nn.Linear / conv weight layout)
int launch_row_gemm(const RowGemm& g, bool b_kn, cudaStream_t st);

}  // namespace cer
