// Log-mel front end of the inline-VGGish modality on the device.  C-ABI: cer_logmel_forward,
// cer_frame_examples (include/cer_b200.h).
// Reference: abaw5_pre_processing/base/vggish/mel_features.py:92-236 (stft_magnitude, periodic_hann,
// spectrogram_to_mel_matrix, log_mel_spectrogram) and my_frame (:21-46) / vggish_input.py:37-82.
//
// The reference computes in float64 (numpy).  The whole front end is 0.44 MFLOP per 10 ms frame, so
// it is done in fp64 here as well (B200: ~40 TFLOP/s fp64, the cost is negligible next to VGGish's
// 1.7 GFLOP per example) and agrees with numpy to ~1e-12 before the final cast to fp32.
//   DFT by definition: X[k] = sum_n x[n] w[n] (cos, -sin)(2 pi k n / fft), twiddles from a table of
//   fft entries indexed (k*n) mod fft; |X| -> mel matrix [bins][mel] -> log(. + offset).
// One CTA per 4 STFT frames (each twiddle read feeds 4 frames); thread k owns spectrogram bin k,
// then one thread per (frame, mel band).
#include <cuda_runtime.h>
#include <cstdint>

#include "../../include/cer_b200.h"
#include "common.h"

namespace cer {

constexpr int kLmFrames = 4;      // STFT frames per CTA: every twiddle read from smem feeds 4 frames

__global__ void __launch_bounds__(288) logmel_kernel(const float* __restrict__ wave, long long n_samples,
                                                     const double* __restrict__ tables, int win, int hop, int fft, int n_mel,
                                                     double log_offset, float* __restrict__ out, int n_frames) {
  extern __shared__ __align__(16) double sm_d[];
  double* s_x = sm_d;                            // [win][kLmFrames] windowed samples, frame-minor
  double* s_cos = s_x + win * kLmFrames;         // [fft]
  double* s_sin = s_cos + fft;                   // [fft]
  double* s_mag = s_sin + fft;                   // [kLmFrames][fft/2 + 1]
  const int bins = fft / 2 + 1;
  const double* hann = tables;        // [win]
  const double* tcos = tables + win;  // [fft]
  const double* tsin = tcos + fft;    // [fft]
  const double* mel = tsin + fft;     // [bins][n_mel]
  const int f0 = blockIdx.x * kLmFrames;
  for (int i = threadIdx.x; i < win * kLmFrames; i += blockDim.x) {
    const int n = i / kLmFrames, q = i - n * kLmFrames;
    const int f = f0 + q;
    s_x[i] = f < n_frames ? (double)wave[(long long)f * hop + n] * hann[n] : 0.0;
  }
  for (int i = threadIdx.x; i < fft; i += blockDim.x) { s_cos[i] = tcos[i]; s_sin[i] = tsin[i]; }
  __syncthreads();
  for (int k = threadIdx.x; k < bins; k += blockDim.x) {
    double re[kLmFrames], im[kLmFrames];
#pragma unroll
    for (int q = 0; q < kLmFrames; ++q) { re[q] = 0.0; im[q] = 0.0; }
    int idx = 0;                                   // (k * n) mod fft, fft is a power of two
    for (int n = 0; n < win; ++n) {
      const double c = s_cos[idx], sn = s_sin[idx];
      const double2 x01 = *reinterpret_cast<const double2*>(&s_x[n * kLmFrames]);
      const double2 x23 = *reinterpret_cast<const double2*>(&s_x[n * kLmFrames + 2]);
      re[0] = fma(x01.x, c, re[0]); im[0] = fma(x01.x, sn, im[0]);
      re[1] = fma(x01.y, c, re[1]); im[1] = fma(x01.y, sn, im[1]);
      re[2] = fma(x23.x, c, re[2]); im[2] = fma(x23.x, sn, im[2]);
      re[3] = fma(x23.y, c, re[3]); im[3] = fma(x23.y, sn, im[3]);
      idx = (idx + k) & (fft - 1);
    }
#pragma unroll
    for (int q = 0; q < kLmFrames; ++q) s_mag[q * bins + k] = sqrt(re[q] * re[q] + im[q] * im[q]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n_mel * kLmFrames; i += blockDim.x) {
    const int q = i / n_mel, j = i - q * n_mel;
    if (f0 + q >= n_frames) continue;
    double acc = 0.0;
    const double* mg = s_mag + q * bins;
    for (int k = 0; k < bins; ++k) acc = fma(mg[k], mel[(long long)k * n_mel + j], acc);
    out[(long long)(f0 + q) * n_mel + j] = (float)log(acc + log_offset);
  }
}

// examples[e][t][:] = logmel[starts[e] + t][:]   (my_frame, mel_features.py:21-46)
__global__ void frame_examples_kernel(const float* __restrict__ logmel, const int* __restrict__ starts, int n_examples,
                                      int frames_per_example, int n_mel, float* __restrict__ out) {
  const long long per = (long long)frames_per_example * n_mel;
  const long long total = (long long)n_examples * per;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int e = (int)(i / per);
    const long long r = i - (long long)e * per;
    out[i] = logmel[(long long)starts[e] * n_mel + r];
  }
}

}  // namespace cer

using namespace cer;

extern "C" int64_t cer_logmel_num_frames(int64_t n_samples, int32_t win, int32_t hop) {
  if (n_samples < win || win <= 0 || hop <= 0) return 0;
  return 1 + (n_samples - win) / hop;
}

extern "C" int cer_logmel_forward(const float* wave_dev, int64_t n_samples, const double* tables_dev, int32_t win, int32_t hop,
                                  int32_t fft, int32_t n_mel, double log_offset, float* logmel_out_dev, void* stream) {
  if (!wave_dev || !tables_dev || !logmel_out_dev || win <= 0 || hop <= 0 || fft < win || (fft & (fft - 1)) || n_mel <= 0 ||
      fft > 4096)
    return set_error(CER_ERR_INVALID, "cer_logmel_forward: bad argument (fft must be a power of two >= win)");
  const int64_t n_frames = cer_logmel_num_frames(n_samples, win, hop);
  if (n_frames <= 0) return CER_OK;
  if (n_frames > (1ll << 30)) return set_error(CER_ERR_INVALID, "cer_logmel_forward: too many frames");
  const size_t smem = (size_t)(win * kLmFrames + 2 * fft + kLmFrames * (fft / 2 + 1)) * sizeof(double);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(logmel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_error(CER_ERR_CUDA, std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e));
  }
  logmel_kernel<<<(int)((n_frames + kLmFrames - 1) / kLmFrames), 288, smem, static_cast<cudaStream_t>(stream)>>>(wave_dev, n_samples, tables_dev, win, hop, fft,
                                                                                n_mel, log_offset, logmel_out_dev, (int)n_frames);
  CER_CUDA(cudaGetLastError());
  return CER_OK;
}

extern "C" int cer_frame_examples(const float* logmel_dev, const int32_t* starts_dev, int32_t n_examples,
                                  int32_t frames_per_example, int32_t n_mel, float* out_dev, void* stream) {
  if (!logmel_dev || !starts_dev || !out_dev || n_examples < 0 || frames_per_example <= 0 || n_mel <= 0)
    return set_error(CER_ERR_INVALID, "cer_frame_examples: bad argument");
  if (n_examples == 0) return CER_OK;
  const long long total = (long long)n_examples * frames_per_example * n_mel;
  const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  frame_examples_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(logmel_dev, starts_dev, n_examples,
                                                                               frames_per_example, n_mel, out_dev);
  CER_CUDA(cudaGetLastError());
  return CER_OK;
}
