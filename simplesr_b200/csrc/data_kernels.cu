// data_kernels.cu — device-side data preparation and evaluation metrics: what the reference does
// on the host with tf.data / tf.image before and after the hot path, so that eight GPUs are not fed by one CPU pipeline.
//   ssr_resize_bicubic   tf.image.resize(method="bicubic", antialias=True) for the LR synthesis
//                             (simple_sr/data_pipeline/data_pipeline.py:318-330): TensorFlow's ScaleAndTranslate with the
//                             Keys cubic kernel (a = -0.5), half-pixel centres, kernel support scaled by the down-scaling
//                             factor, per-pixel normalised weights; separable (columns, then rows).
//   ssr_augment               flip_along_x / flip_along_y / rotate90 (image_transforms.py:157-173, 320-345) and crops
//                             (:50-80) as ONE gather: out[n, y, x] = in[src_n, f(y, x)] - exact copies.
//   ssr_psnr_y                metrics.psnr_on_y (metrics.py:18-44): tf.image.rgb_to_yuv luma, then tf.image.psnr.
//   ssr_ssim                  metrics.ssim (metrics.py:47-59): tf.image.ssim, 11x11 Gaussian (sigma 1.5), k1 0.01, k2 0.03,
//                             VALID windows, mean over positions then over channels.
// All bandwidth / CUDA-core kernels; reductions are two-stage in a fixed order (deterministic).
#include <cuda_runtime.h>

#include <cstdint>

#include "internal.h"

namespace ssr {

static inline int grid_cap(int64_t work, int block, int waves = 16) {
  int64_t g = (work + block - 1) / block;
  if (g > 148 * waves) g = 148 * waves;
  return static_cast<int>(g < 1 ? 1 : g);
}

// ---------------------------------------------------------------- bicubic (Keys a = -0.5) with antialiasing
__device__ __forceinline__ float keys_cubic(float x) {
  x = fabsf(x);
  if (x >= 2.f) return 0.f;
  if (x >= 1.f) return ((-0.5f * x + 2.5f) * x - 4.f) * x + 2.f;
  return ((1.5f * x - 2.5f) * x) * x + 1.f;
}
// span and normalised weights of output index o (TensorFlow ComputeSpansCore): returns the first input index, writes up
// to kMaxSpan weights
constexpr int kMaxSpan = 4 * 8 + 2;  // radius 2 * kernel_scale (<= 8) * 2 + slack
__device__ __forceinline__ int resize_span(int o, int in_size, float inv_scale, float kernel_scale, float* w, int* count) {
  const float sample_f = (o + 0.5f) * inv_scale;
  const float radius = 2.f * kernel_scale;
  int lo = static_cast<int>(ceilf(sample_f - radius - 0.5f));
  int hi = static_cast<int>(floorf(sample_f + radius - 0.5f));
  lo = max(lo, 0);
  hi = min(hi, in_size - 1);
  const float inv_ks = 1.f / kernel_scale;
  float total = 0.f;
  int n = 0;
  for (int j = lo; j <= hi && n < kMaxSpan; ++j, ++n) {
    const float kw = keys_cubic((j + 0.5f - sample_f) * inv_ks);
    w[n] = kw;
    total += kw;
  }
  if (fabsf(total) > 1000.f * 1.17549435e-38f) {
    const float inv = 1.f / total;
    for (int k = 0; k < n; ++k) w[k] *= inv;
  }
  *count = n;
  return lo;
}
// pass 1: columns (x); pass 2: rows (y).  One thread per output element of the pass.
__global__ void resize_cols_kernel(const float* __restrict__ in, float* __restrict__ tmp, int n, int h, int w, int c, int ow,
                                   float inv_scale, float kernel_scale) {
  const int64_t total = static_cast<int64_t>(n) * h * ow * c;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int ch = static_cast<int>(i % c);
    int64_t q = i / c;
    const int ox = static_cast<int>(q % ow);
    q /= ow;  // q = n * h + y
    float wt[kMaxSpan];
    int cnt;
    const int lo = resize_span(ox, w, inv_scale, kernel_scale, wt, &cnt);
    const float* row = in + q * static_cast<int64_t>(w) * c + ch;
    float acc = 0.f;
    for (int k = 0; k < cnt; ++k) acc += wt[k] * __ldg(row + static_cast<int64_t>(lo + k) * c);
    tmp[i] = acc;
  }
}
__global__ void resize_rows_kernel(const float* __restrict__ tmp, float* __restrict__ out, int n, int h, int ow, int c, int oh,
                                   float inv_scale, float kernel_scale) {
  const int64_t total = static_cast<int64_t>(n) * oh * ow * c;
  const int64_t rowc = static_cast<int64_t>(ow) * c;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t xc = i % rowc;
    int64_t q = i / rowc;
    const int oy = static_cast<int>(q % oh);
    const int img = static_cast<int>(q / oh);
    float wt[kMaxSpan];
    int cnt;
    const int lo = resize_span(oy, h, inv_scale, kernel_scale, wt, &cnt);
    const float* col = tmp + static_cast<int64_t>(img) * h * rowc + xc;
    float acc = 0.f;
    for (int k = 0; k < cnt; ++k) acc += wt[k] * __ldg(col + static_cast<int64_t>(lo + k) * rowc);
    out[i] = acc;
  }
}

// ---------------------------------------------------------------- augmentation gather
// mode bits: 1 = flip along x (left-right), 2 = flip along y (up-down), bits 2-3 = rot90 count k (counter-clockwise, as
// tf.image.rot90).  Crop: the source window of output image i starts at (oy[i], ox[i]) of source image src[i]
// (NULL: image i, offset 0).  Order: crop, then rot90, then flips (applied to the cropped patch).
__global__ void augment_kernel(const float* __restrict__ in, float* __restrict__ out, int n_out, int ih, int iw, int c,
                               int oh, int ow, int mode, const int* __restrict__ src, const int* __restrict__ oy,
                               const int* __restrict__ ox) {
  const int k = (mode >> 2) & 3;
  const int ph = (k & 1) ? ow : oh, pw = (k & 1) ? oh : ow;  // patch size before the rotation
  const int64_t total = static_cast<int64_t>(n_out) * oh * ow * c;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int ch = static_cast<int>(i % c);
    int64_t q = i / c;
    int x = static_cast<int>(q % ow);
    q /= ow;
    int y = static_cast<int>(q % oh);
    const int img = static_cast<int>(q / oh);
    if (mode & 1) x = ow - 1 - x;
    if (mode & 2) y = oh - 1 - y;
    int py, px;  // position in the un-rotated patch: rot90(k) maps patch[py, px] -> out[y, x]
    switch (k) {
      case 1: py = x; px = pw - 1 - y; break;            // out[y, x] = patch[x, W-1-y]
      case 2: py = ph - 1 - y; px = pw - 1 - x; break;
      case 3: py = ph - 1 - x; px = y; break;
      default: py = y; px = x; break;
    }
    const int s = src ? src[img] : img;
    const int sy = py + (oy ? oy[img] : 0), sx = px + (ox ? ox[img] : 0);
    out[i] = __ldg(in + ((static_cast<int64_t>(s) * ih + sy) * iw + sx) * c + ch);
  }
}

// ---------------------------------------------------------------- PSNR on Y
constexpr int kMetBlocks = 32;
constexpr int kMetThreads = 256;
__device__ __forceinline__ float block_sum(float v, float* sm) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  if (threadIdx.x == 0)
    for (int k = 0; k < kMetThreads / 32; ++k) t += sm[k];
  __syncthreads();
  return t;  // valid in thread 0
}
__global__ void __launch_bounds__(kMetThreads) psnr_y_partial_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                                     int64_t pixels, float* __restrict__ partial) {
  __shared__ float sm[kMetThreads / 32];
  const int img = blockIdx.y;
  const int64_t chunk = (pixels + kMetBlocks - 1) / kMetBlocks;
  const int64_t lo = blockIdx.x * chunk, hi = min(pixels, lo + chunk);
  const float* pa = a + img * pixels * 3;
  const float* pb = b + img * pixels * 3;
  float acc = 0.f;
  for (int64_t p = lo + threadIdx.x; p < hi; p += kMetThreads) {
    // tf.image.rgb_to_yuv: Y = 0.299 R + 0.587 G + 0.114 B
    const float ya = 0.299f * pa[3 * p] + 0.587f * pa[3 * p + 1] + 0.114f * pa[3 * p + 2];
    const float yb = 0.299f * pb[3 * p] + 0.587f * pb[3 * p + 1] + 0.114f * pb[3 * p + 2];
    const float d = ya - yb;
    acc += d * d;
  }
  const float t = block_sum(acc, sm);
  if (threadIdx.x == 0) partial[img * kMetBlocks + blockIdx.x] = t;
}
__global__ void psnr_final_kernel(const float* __restrict__ partial, int n, int64_t count, float max_val,
                                  float* __restrict__ out) {
  for (int img = threadIdx.x; img < n; img += blockDim.x) {
    double s = 0.0;
    for (int k = 0; k < kMetBlocks; ++k) s += partial[img * kMetBlocks + k];
    const double mse = s / static_cast<double>(count);
    out[img] = static_cast<float>(20.0 * log10(static_cast<double>(max_val)) - 10.0 * log10(mse));
  }
}

// ---------------------------------------------------------------- SSIM
struct SsimWeights {
  float w[11];
};
// one thread per (window position, channel); the window sums use the separable Gaussian directly (121 taps)
__global__ void __launch_bounds__(kMetThreads) ssim_partial_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                                   int h, int w, int c, float c1, float c2,
                                                                   const SsimWeights g, float* __restrict__ partial) {
  __shared__ float sm[kMetThreads / 32];
  const int img = blockIdx.y;
  const int vh = h - 10, vw = w - 10;
  const int64_t total = static_cast<int64_t>(vh) * vw * c;
  const int64_t chunk = (total + kMetBlocks - 1) / kMetBlocks;
  const int64_t lo = blockIdx.x * chunk, hi = min(total, lo + chunk);
  const float* pa = a + static_cast<int64_t>(img) * h * w * c;
  const float* pb = b + static_cast<int64_t>(img) * h * w * c;
  float acc = 0.f;
  for (int64_t i = lo + threadIdx.x; i < hi; i += kMetThreads) {
    const int ch = static_cast<int>(i % c);
    const int64_t q = i / c;
    const int x = static_cast<int>(q % vw), y = static_cast<int>(q / vw);
    float m0 = 0.f, m1 = 0.f, s01 = 0.f, s00 = 0.f;
    for (int dy = 0; dy < 11; ++dy) {
      float r0 = 0.f, r1 = 0.f, r01 = 0.f, r00 = 0.f;
      const int64_t base = (static_cast<int64_t>(y + dy) * w + x) * c + ch;
#pragma unroll
      for (int dx = 0; dx < 11; ++dx) {
        const float va = __ldg(pa + base + static_cast<int64_t>(dx) * c), vb = __ldg(pb + base + static_cast<int64_t>(dx) * c);
        r0 += g.w[dx] * va;
        r1 += g.w[dx] * vb;
        r01 += g.w[dx] * va * vb;
        r00 += g.w[dx] * (va * va + vb * vb);
      }
      m0 += g.w[dy] * r0;
      m1 += g.w[dy] * r1;
      s01 += g.w[dy] * r01;
      s00 += g.w[dy] * r00;
    }
    // tf.image ops _ssim_helper: luminance * contrast-structure
    const float num0 = m0 * m1 * 2.f, den0 = m0 * m0 + m1 * m1;
    const float lum = (num0 + c1) / (den0 + c1);
    const float num1 = s01 * 2.f, den1 = s00;
    const float cs = (num1 - num0 + c2) / (den1 - den0 + c2);
    acc += lum * cs;
  }
  const float t = block_sum(acc, sm);
  if (threadIdx.x == 0) partial[img * kMetBlocks + blockIdx.x] = t;
}
__global__ void mean_final_kernel(const float* __restrict__ partial, int n, int64_t count, float* __restrict__ out) {
  for (int img = threadIdx.x; img < n; img += blockDim.x) {
    double s = 0.0;
    for (int k = 0; k < kMetBlocks; ++k) s += partial[img * kMetBlocks + k];
    out[img] = static_cast<float>(s / static_cast<double>(count));
  }
}

}  // namespace ssr

using namespace ssr;

#define SSR_CHECK_LAUNCH(name)                                                          \
  do {                                                                                  \
    cudaError_t e__ = cudaGetLastError();                                               \
    if (e__ != cudaSuccess) return set_error(SSR_ERR_CUDA, name ": %s", cudaGetErrorString(e__)); \
  } while (0)

extern "C" size_t ssr_resize_workspace_bytes(int n, int h, int w, int c, int scale) {
  if (n <= 0 || h <= 0 || w <= 0 || c <= 0 || scale <= 0) return 0;
  return static_cast<size_t>(n) * h * (w / scale) * c * sizeof(float);
}

extern "C" int ssr_resize_bicubic(const float* x, float* y, int n, int h, int w, int c, int scale, int antialias,
                                       void* workspace, void* stream) {
  if (!x || !y || !workspace || n <= 0 || h <= 0 || w <= 0 || c <= 0 || scale < 1 || scale > 8 || h % scale || w % scale)
    return set_error(SSR_ERR_INVALID, "resize_bicubic: sizes must be positive multiples of scale (1..8)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int oh = h / scale, ow = w / scale;
  const float inv_scale = static_cast<float>(scale);                       // input pixels per output pixel
  const float kernel_scale = antialias ? static_cast<float>(scale) : 1.f;  // support widened when down-scaling
  float* tmp = static_cast<float*>(workspace);
  resize_cols_kernel<<<grid_cap(static_cast<int64_t>(n) * h * ow * c, 256), 256, 0, st>>>(x, tmp, n, h, w, c, ow, inv_scale,
                                                                                       kernel_scale);
  SSR_CHECK_LAUNCH("resize_cols");
  resize_rows_kernel<<<grid_cap(static_cast<int64_t>(n) * oh * ow * c, 256), 256, 0, st>>>(tmp, y, n, h, ow, c, oh, inv_scale,
                                                                                        kernel_scale);
  SSR_CHECK_LAUNCH("resize_rows");
  return SSR_OK;
}

extern "C" int ssr_augment(const float* x, float* y, int n_out, int in_h, int in_w, int c, int out_h, int out_w, int mode,
                           const int* src_index, const int* off_y, const int* off_x, void* stream) {
  if (!x || !y || n_out <= 0 || in_h <= 0 || in_w <= 0 || c <= 0 || out_h <= 0 || out_w <= 0 || mode < 0 || mode > 15)
    return set_error(SSR_ERR_INVALID, "augment: bad argument");
  const int k = (mode >> 2) & 3;
  const int ph = (k & 1) ? out_w : out_h, pw = (k & 1) ? out_h : out_w;
  if (ph > in_h || pw > in_w) return set_error(SSR_ERR_INVALID, "augment: patch larger than the source image");
  const int64_t total = static_cast<int64_t>(n_out) * out_h * out_w * c;
  augment_kernel<<<grid_cap(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, y, n_out, in_h, in_w, c, out_h, out_w,
                                                                                   mode, src_index, off_y, off_x);
  SSR_CHECK_LAUNCH("augment");
  return SSR_OK;
}

extern "C" size_t ssr_metric_workspace_bytes(int n) { return static_cast<size_t>(n > 0 ? n : 1) * kMetBlocks * sizeof(float); }

extern "C" int ssr_psnr_y(const float* a, const float* b, int n, int h, int w, float max_val, void* workspace, float* out,
                          void* stream) {
  if (!a || !b || !workspace || !out || n <= 0 || h <= 0 || w <= 0 || max_val <= 0)
    return set_error(SSR_ERR_INVALID, "psnr_y: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t pixels = static_cast<int64_t>(h) * w;
  psnr_y_partial_kernel<<<dim3(kMetBlocks, n), kMetThreads, 0, st>>>(a, b, pixels, static_cast<float*>(workspace));
  SSR_CHECK_LAUNCH("psnr_y_partial");
  psnr_final_kernel<<<1, 128, 0, st>>>(static_cast<const float*>(workspace), n, pixels, max_val, out);
  SSR_CHECK_LAUNCH("psnr_y_final");
  return SSR_OK;
}

extern "C" int ssr_ssim(const float* a, const float* b, int n, int h, int w, int c, float max_val, void* workspace, float* out,
                        void* stream) {
  if (!a || !b || !workspace || !out || n <= 0 || h < 11 || w < 11 || c <= 0 || max_val <= 0)
    return set_error(SSR_ERR_INVALID, "ssim: images must be at least 11 x 11 (tf.image.ssim filter size)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SsimWeights g;     // _fspecial_gauss(11, 1.5): softmax of -x^2 / (2 sigma^2) along one axis (the 2-D kernel is its outer product)
  double tot = 0.0, e[11];
  for (int i = 0; i < 11; ++i) {
    const double d = i - 5.0;
    e[i] = exp(-0.5 * d * d / (1.5 * 1.5));
    tot += e[i];
  }
  for (int i = 0; i < 11; ++i) g.w[i] = static_cast<float>(e[i] / tot);
  const float c1 = (0.01f * max_val) * (0.01f * max_val), c2 = (0.03f * max_val) * (0.03f * max_val);
  ssim_partial_kernel<<<dim3(kMetBlocks, n), kMetThreads, 0, st>>>(a, b, h, w, c, c1, c2, g, static_cast<float*>(workspace));
  SSR_CHECK_LAUNCH("ssim_partial");
  mean_final_kernel<<<1, 128, 0, st>>>(static_cast<const float*>(workspace), n, static_cast<int64_t>(h - 10) * (w - 10) * c, out);
  SSR_CHECK_LAUNCH("ssim_final");
  return SSR_OK;
}
