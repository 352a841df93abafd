// Eval-time input pipeline on the device: stored uint8 face crops [N,H,W,3] -> the fp32 NCHW
// [N,3,40,40] tensor in [-1,1] that IR-50 consumes.  C-ABI: cer_preproc_* (include/cer_b200.h).
//
// Reference semantics (base/dataset.py:503-510, base/transforms3D.py:15-144):
//   GroupNumpyToPILImage -> GroupScale(48) [PIL antialiased BILINEAR resize of the smaller edge]
//   -> GroupCenterCrop(40) -> Stack -> ToTorchFormatTensor (/255) -> GroupNormalize(.5, .5).
// The resize is Pillow's two-pass 8-bit resampler (libImaging/Resample.c): triangle filter whose
// support grows with the down-scale factor, coefficients in 22-bit fixed point, the horizontal
// pass rounded to uint8 before the vertical pass.  Integer arithmetic throughout => bit-exact.
// Only the 40x40 crop window is computed: the rows/columns the crop never reads are skipped.
//
// One CTA per frame: tables -> smem, horizontal pass of the needed input rows into a uint8 smem
// slab [rows][crop][3], vertical pass + normalisation straight to global.  HBM bound: ~0.55 of
// the 196 KB frame is read once (the crop window's footprint), 19.2 KB written.
#include <cuda_runtime.h>
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../include/cer_b200.h"
#include "common.h"

namespace cer {

constexpr int kPrecisionBits = 32 - 8 - 2;     // Pillow PRECISION_BITS for 8-bit channels

struct PreprocParams {
  int H, W;                 // stored frame size
  int crop_h, crop_w;       // 40 x 40
  int y0, rows;             // first input row the crop needs, number of rows
  int ks_x, ks_y;           // taps per output column / row (table row pitch)
  const int* bx;            // [crop_w][2]  (xmin, count)
  const int* kx;            // [crop_w][ks_x]
  const int* by;            // [crop_h][2]  (ymin, count)
  const int* ky;            // [crop_h][ks_y]
};

__device__ __forceinline__ int clip8(int v) { v >>= kPrecisionBits; return v < 0 ? 0 : (v > 255 ? 255 : v); }

__global__ void __launch_bounds__(256) preprocess_kernel(const uint8_t* __restrict__ frames, float* __restrict__ out,
                                                         const PreprocParams p) {
  extern __shared__ __align__(16) uint8_t sm[];
  int* s_bx = reinterpret_cast<int*>(sm);
  int* s_kx = s_bx + p.crop_w * 2;
  int* s_by = s_kx + p.crop_w * p.ks_x;
  int* s_ky = s_by + p.crop_h * 2;
  uint8_t* inter = reinterpret_cast<uint8_t*>(s_ky + p.crop_h * p.ks_y);       // [rows][crop_w][3]
  for (int i = threadIdx.x; i < p.crop_w * 2; i += blockDim.x) s_bx[i] = p.bx[i];
  for (int i = threadIdx.x; i < p.crop_w * p.ks_x; i += blockDim.x) s_kx[i] = p.kx[i];
  for (int i = threadIdx.x; i < p.crop_h * 2; i += blockDim.x) s_by[i] = p.by[i];
  for (int i = threadIdx.x; i < p.crop_h * p.ks_y; i += blockDim.x) s_ky[i] = p.ky[i];
  __syncthreads();
  const uint8_t* f = frames + static_cast<size_t>(blockIdx.x) * p.H * p.W * 3;
  // horizontal pass (ImagingResampleHorizontal_8bpc): consecutive threads -> consecutive output columns
  const int tasks = p.rows * p.crop_w;
  for (int t = threadIdx.x; t < tasks; t += blockDim.x) {
    const int r = t / p.crop_w, xo = t - r * p.crop_w;
    const int xmin = s_bx[2 * xo], cnt = s_bx[2 * xo + 1];
    const uint8_t* src = f + (static_cast<size_t>(p.y0 + r) * p.W + xmin) * 3;
    const int* k = s_kx + xo * p.ks_x;
    int a0 = 1 << (kPrecisionBits - 1), a1 = a0, a2 = a0;
    for (int j = 0; j < cnt; ++j) {
      const int kk = k[j];
      a0 += __ldg(src + 3 * j) * kk;
      a1 += __ldg(src + 3 * j + 1) * kk;
      a2 += __ldg(src + 3 * j + 2) * kk;
    }
    uint8_t* d = inter + t * 3;
    d[0] = static_cast<uint8_t>(clip8(a0));
    d[1] = static_cast<uint8_t>(clip8(a1));
    d[2] = static_cast<uint8_t>(clip8(a2));
  }
  __syncthreads();
  // vertical pass (ImagingResampleVertical_8bpc) + /255 + (x - .5)/.5, written NCHW
  const int plane = p.crop_h * p.crop_w;
  float* o = out + static_cast<size_t>(blockIdx.x) * 3 * plane;
  for (int t = threadIdx.x; t < 3 * plane; t += blockDim.x) {
    const int c = t / plane, rem = t - c * plane;
    const int yo = rem / p.crop_w, xo = rem - yo * p.crop_w;
    const int ymin = s_by[2 * yo] - p.y0, cnt = s_by[2 * yo + 1];
    const int* k = s_ky + yo * p.ks_y;
    int a = 1 << (kPrecisionBits - 1);
    for (int j = 0; j < cnt; ++j) a += inter[((ymin + j) * p.crop_w + xo) * 3 + c] * k[j];
    const float v = __fdiv_rn(static_cast<float>(clip8(a)), 255.f);       // ToTorchFormatTensor
    o[t] = (v - 0.5f) * 2.0f;                                             // Normalize(.5, .5): (x - .5) / .5
  }
}

// precompute_coeffs + normalize_coeffs_8bpc for the triangle (BILINEAR) filter over a whole axis
static void pil_coeffs(int in_size, int out_size, std::vector<int>& bounds, std::vector<int>& kk, int& ksize) {
  const double scale = (double)in_size / out_size;
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = 1.0 * filterscale;
  ksize = (int)ceil(support) * 2 + 1;
  bounds.assign((size_t)out_size * 2, 0);
  kk.assign((size_t)out_size * ksize, 0);
  const double ss = 1.0 / filterscale;
  std::vector<double> k(ksize);
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = (xx + 0.5) * scale;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    double ww = 0.0;
    for (int x = 0; x < xmax; ++x) {
      double a = (x + xmin - center + 0.5) * ss;
      if (a < 0) a = -a;
      const double w = a < 1.0 ? 1.0 - a : 0.0;
      k[x] = w;
      ww += w;
    }
    for (int x = 0; x < xmax; ++x) {
      if (ww != 0.0) k[x] /= ww;
      const double pk = k[x] * (double)(1 << kPrecisionBits);
      kk[(size_t)xx * ksize + x] = k[x] < 0 ? (int)(-0.5 + pk) : (int)(0.5 + pk);
    }
    bounds[2 * xx] = xmin;
    bounds[2 * xx + 1] = xmax;
  }
}

}  // namespace cer

using namespace cer;

struct cer_preproc {
  PreprocParams p;
  size_t smem;
};

static const size_t kPreprocTableBytes = 64 * 1024;

extern "C" size_t cer_preproc_workspace_bytes(void) { return kPreprocTableBytes; }

extern "C" int cer_preproc_create(cer_preproc** out, int32_t in_h, int32_t in_w, int32_t resize, int32_t crop,
                                  void* workspace_dev, size_t workspace_bytes) {
  if (!out || !workspace_dev || in_h <= 0 || in_w <= 0 || resize <= 0 || crop <= 0 || crop > resize)
    return set_error(CER_ERR_INVALID, "cer_preproc_create: bad argument");
  if (workspace_bytes < kPreprocTableBytes) return set_error(CER_ERR_WORKSPACE, "cer_preproc_create: workspace too small");
  int rc = cer_check_device();
  if (rc) return rc;
  // torchvision Resize(int): the smaller edge becomes `resize` (GroupScale, transforms3D.py:103-117)
  int oh, ow;
  if (in_w <= in_h) { ow = resize; oh = (int)((double)resize * in_h / in_w); }
  else { oh = resize; ow = (int)((double)resize * in_w / in_h); }
  // torchvision CenterCrop: round-half-even of (size - crop) / 2
  const int top = (int)nearbyint((oh - crop) / 2.0), left = (int)nearbyint((ow - crop) / 2.0);
  std::vector<int> bxa, kxa, bya, kya;
  int ksx = 0, ksy = 0;
  pil_coeffs(in_w, ow, bxa, kxa, ksx);
  pil_coeffs(in_h, oh, bya, kya, ksy);
  std::vector<int> host;
  host.insert(host.end(), bxa.begin() + 2 * left, bxa.begin() + 2 * (left + crop));
  host.insert(host.end(), kxa.begin() + (size_t)left * ksx, kxa.begin() + (size_t)(left + crop) * ksx);
  host.insert(host.end(), bya.begin() + 2 * top, bya.begin() + 2 * (top + crop));
  host.insert(host.end(), kya.begin() + (size_t)top * ksy, kya.begin() + (size_t)(top + crop) * ksy);
  if (host.size() * 4 > kPreprocTableBytes) return set_error(CER_ERR_INVALID, "cer_preproc_create: down-scale factor too large for the table");
  const int y0 = bya[2 * top];
  const int y1 = bya[2 * (top + crop - 1)] + bya[2 * (top + crop - 1) + 1];
  cer_preproc* q = new cer_preproc();
  PreprocParams& p = q->p;
  p.H = in_h; p.W = in_w; p.crop_h = crop; p.crop_w = crop; p.y0 = y0; p.rows = y1 - y0; p.ks_x = ksx; p.ks_y = ksy;
  int* base = static_cast<int*>(workspace_dev);
  p.bx = base;
  p.kx = p.bx + 2 * crop;
  p.by = p.kx + (size_t)crop * ksx;
  p.ky = p.by + 2 * crop;
  q->smem = host.size() * 4 + (size_t)p.rows * crop * 3;
  if (q->smem > 200 * 1024) { delete q; return set_error(CER_ERR_INVALID, "cer_preproc_create: frame too large for the shared-memory slab"); }
  cudaError_t e = cudaMemcpy(workspace_dev, host.data(), host.size() * 4, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { delete q; return set_error(CER_ERR_CUDA, std::string("cudaMemcpy: ") + cudaGetErrorString(e)); }
  e = cudaFuncSetAttribute(preprocess_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);   // the cap above, for every plan
  if (e != cudaSuccess) { delete q; return set_error(CER_ERR_CUDA, std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e)); }
  *out = q;
  return CER_OK;
}

extern "C" int cer_preproc_forward(cer_preproc* q, const uint8_t* frames_dev, int64_t n_frames, float* out_dev, void* stream) {
  if (!q || n_frames < 0 || (n_frames > 0 && (!frames_dev || !out_dev)) || n_frames > (1ll << 30))
    return set_error(CER_ERR_INVALID, "cer_preproc_forward: bad argument");
  if (n_frames == 0) return CER_OK;
  preprocess_kernel<<<(int)n_frames, 256, q->smem, static_cast<cudaStream_t>(stream)>>>(frames_dev, out_dev, q->p);
  CER_CUDA(cudaGetLastError());
  return CER_OK;
}

extern "C" void cer_preproc_destroy(cer_preproc* q) { delete q; }
