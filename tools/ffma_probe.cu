// FP32 FMA throughput probe: 3-register FFMA vs packed FFMA2 (fma.rn.f32x2), 16 independent
// accumulators per thread, 8 warps per CTA, 4 CTAs per SM.  Prints FMA lanes per clock per SM.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_ffma(float* out, float a, float b, int iters) {
  float acc[16];
  for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 0.001f + i;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = fmaf(acc[i], a, b + i);
  float s = 0;
  for (int i = 0; i < 16; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ffma2(float* out, float a, float b, int iters) {
  unsigned long long acc[16], av, bv;
  asm("mov.b64 %0, {%1, %1};" : "=l"(av) : "f"(a));
  asm("mov.b64 %0, {%1, %1};" : "=l"(bv) : "f"(b));
  for (int i = 0; i < 16; ++i) { float x = threadIdx.x * 0.001f + i; asm("mov.b64 %0, {%1, %1};" : "=l"(acc[i]) : "f"(x)); }
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < 16; ++i) asm("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc[i]) : "l"(av), "l"(bv));
  float s = 0;
  for (int i = 0; i < 16; ++i) { float x, y; asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(acc[i])); s += x + y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  float* out; cudaMalloc(&out, sms * 4 * 256 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  for (int mode = 0; mode < 2; ++mode) {
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      if (mode == 0) k_ffma<<<sms * 4, 256>>>(out, 1.0001f, 0.5f, iters); else k_ffma2<<<sms * 4, 256>>>(out, 1.0001f, 0.5f, iters);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double fma = (double)sms * 4 * 256 * 16 * iters * (mode ? 2 : 1);
      if (rep == 2) printf("%s: %.3f ms  %.1f TFLOP/s  %.1f FMA/clk/SM at max clock %d MHz\n", mode ? "FFMA2" : "FFMA ", ms,
                           2 * fma / ms / 1e9, fma / (ms * 1e-3) / sms / (khz * 1e3), khz / 1000);
    }
  }
  return 0;
}
