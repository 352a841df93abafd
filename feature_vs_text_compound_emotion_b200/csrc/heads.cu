// fp32 building blocks for the alternative fusion heads of the reference (CAN, JMT, MT;
// models/model.py:529-684, :709-750, :895-1167): nn.Linear, the AttentionFusion gate, single-head
// scaled-dot-product attention over a sequence (nn.MultiheadAttention(E, 1)), residual + LayerNorm.
// C-ABI: cer_linear_forward, cer_softmax_gate, cer_sdpa_forward, cer_add_layernorm.
// These heads are ~0.1 % of the path's FLOPs and latency bound; they run exact fp32 on CUDA cores
// (the linear layers share the packed-FFMA2 row GEMM of the training plan).
#include <cuda_runtime.h>
#include <cstdint>

#include "../../include/cer_b200.h"
#include "common.h"
#include "row_gemm.h"

namespace cer {

// out[r, :] = softmax(gate[r, :]) * feat[r, :]    (AttentionFusion.forward, model.py:563-567)
__global__ void __launch_bounds__(256) softmax_gate_kernel(const float* __restrict__ gate, const float* __restrict__ feat,
                                                           int rows, int dim, float* __restrict__ out) {
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= rows) return;
  const float* g = gate + (long long)r * dim;
  float mx = -3.4e38f;
  for (int i = lane; i < dim; i += 32) mx = fmaxf(mx, g[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float s = 0.f;
  for (int i = lane; i < dim; i += 32) s += expf(g[i] - mx);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float inv = 1.f / s;
  for (int i = lane; i < dim; i += 32) out[(long long)r * dim + i] = expf(g[i] - mx) * inv * feat[(long long)r * dim + i];
}

// Single-head attention: out[b, i, :] = softmax_j(q[b,i,:] . k[b,j,:] / sqrt(E)) v[b,j,:].
// One warp per query row: lanes stride over the keys for the scores (kept in shared memory), then
// over the E output features.  E <= 128 * 4, len_k <= 4096 (scores in smem).
__global__ void __launch_bounds__(256) sdpa_kernel(const float* __restrict__ q, int ldq, const float* __restrict__ k, int ldk,
                                                   const float* __restrict__ v, int ldv, int len_q, int len_k, int E,
                                                   float* __restrict__ out, int ldo) {
  extern __shared__ float sm_f[];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* s_q = sm_f + w * (E + ((len_k + 3) & ~3));       // [E], 16-byte aligned per warp
  float* s_p = s_q + E;                      // [len_k]
  const int b = blockIdx.y;
  const int i = blockIdx.x * 8 + w;
  if (i >= len_q) return;
  const float* qp = q + ((long long)b * len_q + i) * ldq;
  for (int d = lane; d < E; d += 32) s_q[d] = qp[d];
  __syncwarp();
  const float scale = rsqrtf((float)E);
  const float* kb = k + (long long)b * len_k * ldk;
  float mx = -3.4e38f;
  for (int j = lane; j < len_k; j += 32) {
    const float4* kp = reinterpret_cast<const float4*>(kb + (long long)j * ldk);
    float s = 0.f;
    for (int d = 0; d < E / 4; ++d) {
      const float4 kk = __ldg(kp + d);
      const float4 qq = *reinterpret_cast<const float4*>(s_q + 4 * d);
      s = fmaf(qq.x, kk.x, s); s = fmaf(qq.y, kk.y, s); s = fmaf(qq.z, kk.z, s); s = fmaf(qq.w, kk.w, s);
    }
    s *= scale;
    s_p[j] = s;
    mx = fmaxf(mx, s);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float den = 0.f;
  for (int j = lane; j < len_k; j += 32) { const float e = expf(s_p[j] - mx); s_p[j] = e; den += e; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) den += __shfl_xor_sync(0xffffffffu, den, o);
  __syncwarp();
  const float inv = 1.f / den;
  const float* vb = v + (long long)b * len_k * ldv;
  float* op = out + ((long long)b * len_q + i) * ldo;
  for (int d = lane; d < E; d += 32) {
    float acc = 0.f;
    for (int j = 0; j < len_k; ++j) acc = fmaf(s_p[j], __ldg(vb + (long long)j * ldv + d), acc);
    op[d] = acc * inv;
  }
}

// out[r, :] = LayerNorm(x[r, :] + res[r, :]) * gamma + beta   (res may be null), one warp per row
__global__ void __launch_bounds__(256) add_layernorm_kernel(const float* __restrict__ x, const float* __restrict__ res,
                                                            int rows, int dim, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, float eps,
                                                            float* __restrict__ out) {
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= rows) return;
  const float* xp = x + (long long)r * dim;
  const float* rp = res ? res + (long long)r * dim : nullptr;
  float s = 0.f;
  for (int i = lane; i < dim; i += 32) s += xp[i] + (rp ? rp[i] : 0.f);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / dim;
  float ss = 0.f;
  for (int i = lane; i < dim; i += 32) { const float d = xp[i] + (rp ? rp[i] : 0.f) - mean; ss += d * d; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float rstd = 1.f / sqrtf(ss / dim + eps);
  for (int i = lane; i < dim; i += 32)
    out[(long long)r * dim + i] = (xp[i] + (rp ? rp[i] : 0.f) - mean) * rstd * gamma[i] + beta[i];
}

// y = tanh(y) in place (the REGRESSION task's output squashing, models/model.py:523, :682, :1165)
__global__ void tanh_inplace_kernel(float* __restrict__ y, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = tanhf(y[i]);
}

}  // namespace cer

using namespace cer;

extern "C" int cer_linear_forward(const float* x_dev, int64_t rows, int32_t in_dim, int32_t ldx, const float* w_dev,
                                  const float* bias_dev, int32_t out_dim, int32_t act, float* y_dev, int32_t ldy, void* stream) {
  if (!x_dev || !w_dev || !y_dev || rows < 0 || in_dim <= 0 || out_dim <= 0 || ldx < in_dim || ldy < out_dim || act < 0 || act > 2 ||
      rows > (1 << 24))
    return set_error(CER_ERR_INVALID, "cer_linear_forward: bad argument");
  if (rows == 0) return CER_OK;
  int rc = cer_check_device();
  if (rc) return rc;
  RowGemm g{};
  g.A = x_dev; g.lda = ldx; g.B = w_dev; g.C = y_dev; g.ldc = ldy;
  g.R = (int)rows; g.T = (int)rows; g.N = out_dim; g.K = in_dim; g.taps = 1;
  g.bias = bias_dev;
  g.epi = act == 0 ? EPI_LINEAR : (act == 1 ? EPI_LRELU_DROP : EPI_RELU);     // LeakyReLU = LRELU_DROP with p = 0
  g.drop.key = 0; g.drop.thr = 0; g.drop.scale = 1.f;
  return launch_row_gemm(g, false, static_cast<cudaStream_t>(stream));
}

extern "C" int cer_softmax_gate(const float* gate_dev, const float* feat_dev, int64_t rows, int32_t dim, float* out_dev,
                                void* stream) {
  if (!gate_dev || !feat_dev || !out_dev || rows < 0 || dim <= 0 || rows > (1 << 30))
    return set_error(CER_ERR_INVALID, "cer_softmax_gate: bad argument");
  if (rows == 0) return CER_OK;
  softmax_gate_kernel<<<(int)((rows + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(gate_dev, feat_dev, (int)rows, dim,
                                                                                           out_dev);
  CER_CUDA(cudaGetLastError());
  return CER_OK;
}

extern "C" int cer_sdpa_forward(const float* q_dev, int32_t ldq, const float* k_dev, int32_t ldk, const float* v_dev,
                                int32_t ldv, int32_t batch, int32_t len_q, int32_t len_k, int32_t dim, float* out_dev,
                                int32_t ldo, void* stream) {
  if (!q_dev || !k_dev || !v_dev || !out_dev || batch <= 0 || len_q <= 0 || len_k <= 0 || dim <= 0 || dim % 4 || ldq < dim ||
      ldk < dim || ldv < dim || ldo < dim || ldk % 4 || batch > 65535)
    return set_error(CER_ERR_INVALID, "cer_sdpa_forward: bad argument (dim and ldk must be multiples of 4)");
  if ((reinterpret_cast<uintptr_t>(k_dev) & 15) != 0) return set_error(CER_ERR_INVALID, "cer_sdpa_forward: k must be 16B aligned");
  const size_t smem = (size_t)8 * (dim + ((len_k + 3) & ~3)) * sizeof(float);
  if (smem > 200 * 1024) return set_error(CER_ERR_INVALID, "cer_sdpa_forward: sequence too long for the shared-memory score rows");
  if (smem > 48 * 1024) {
    static unsigned long long configured = 0;
    if (first_use_on_device(&configured))
      CER_CUDA(cudaFuncSetAttribute(sdpa_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  }
  dim3 grid((len_q + 7) / 8, batch);
  sdpa_kernel<<<grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(q_dev, ldq, k_dev, ldk, v_dev, ldv, len_q, len_k, dim,
                                                                     out_dev, ldo);
  CER_CUDA(cudaGetLastError());
  return CER_OK;
}

extern "C" int cer_add_layernorm(const float* x_dev, const float* res_dev, int64_t rows, int32_t dim, const float* gamma_dev,
                                 const float* beta_dev, float eps, float* out_dev, void* stream) {
  if (!x_dev || !gamma_dev || !beta_dev || !out_dev || rows < 0 || dim <= 0 || rows > (1 << 30))
    return set_error(CER_ERR_INVALID, "cer_add_layernorm: bad argument");
  if (rows == 0) return CER_OK;
  add_layernorm_kernel<<<(int)((rows + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(x_dev, res_dev, (int)rows, dim,
                                                                                            gamma_dev, beta_dev, eps, out_dev);
  CER_CUDA(cudaGetLastError());
  return CER_OK;
}

extern "C" int cer_tanh_inplace(float* y_dev, int64_t n, void* stream) {
  if (!y_dev || n < 0) return set_error(CER_ERR_INVALID, "cer_tanh_inplace: bad argument");
  if (n == 0) return CER_OK;
  const int blocks = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
  tanh_inplace_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(y_dev, (long long)n);
  CER_CUDA(cudaGetLastError());
  return CER_OK;
}
