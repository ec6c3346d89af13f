// train_kernels.cu — HBM-bound kernels of the training step: pixel losses (+ gradient, + per-image PSNR),
// fused multi-tensor Adam, per-channel reductions (bias / PReLU-slope gradients), dtype edge of the backward pass.
// Reference: mean_squared_error.py:57-58, mean_absolute_error.py:57-58, metrics.py:4-15 (tf.image.psnr),
// sr_model.py:436-441 (tape.gradient + Adam.apply_gradients), model_builder.py:118,281,314 (PReLU).
// All reductions are two-stage with a fixed summation order (deterministic, no float atomics).
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>

#include "internal.h"
#include "ptx_sm100.cuh"

namespace ssr {

constexpr int kLossBlocksPerImage = 32;
constexpr int kLossThreads = 256;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// stage 1: block (b, img) reduces a fixed contiguous slice of image img; optionally writes d(loss)/d(sr).
// 16-byte loads / stores, two vectors in flight per thread (the slice boundaries are multiples of 4 elements).
__device__ __forceinline__ void pixel_loss_elem(float h, float s, float g_mse, float g_mae, float& sq, float& ab, float& g) {
  const float d = s - h;
  sq += d * d;
  ab += fabsf(d);
  g = g_mse * d + g_mae * (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f));
}
__global__ void __launch_bounds__(kLossThreads)
    pixel_loss_partial_kernel(const float* __restrict__ hr, const float* __restrict__ sr, int64_t per_image, float g_mse,
                              float g_mae, float* __restrict__ grad, float2* __restrict__ partial) {
  const int img = blockIdx.y, b = blockIdx.x;
  int64_t chunk = (per_image + kLossBlocksPerImage - 1) / kLossBlocksPerImage;
  chunk = (chunk + 3) & ~static_cast<int64_t>(3);
  const int64_t lo = min(per_image, static_cast<int64_t>(b) * chunk), hi = min(per_image, lo + chunk);
  const float* h = hr + img * per_image;
  const float* s = sr + img * per_image;
  float* gr = grad ? grad + img * per_image : nullptr;
  float sq = 0.f, ab = 0.f;
  const bool vec = ((reinterpret_cast<uintptr_t>(h) | reinterpret_cast<uintptr_t>(s) | reinterpret_cast<uintptr_t>(gr)) & 15) == 0;
  int64_t i0 = lo;
  if (vec) {
    const int64_t n4 = (hi - lo) >> 2;
    const float4* h4 = reinterpret_cast<const float4*>(h + lo);
    const float4* s4 = reinterpret_cast<const float4*>(s + lo);
    float4* g4 = gr ? reinterpret_cast<float4*>(gr + lo) : nullptr;
    for (int64_t i = threadIdx.x; i < n4; i += 2 * kLossThreads) {
      const int64_t j = i + kLossThreads;
      const bool two = j < n4;
      const float4 ha = __ldg(h4 + i), sa = __ldg(s4 + i);
      float4 hb = make_float4(0.f, 0.f, 0.f, 0.f), sb = hb;
      if (two) {
        hb = __ldg(h4 + j);
        sb = __ldg(s4 + j);
      }
      float4 ga, gb;
      pixel_loss_elem(ha.x, sa.x, g_mse, g_mae, sq, ab, ga.x);
      pixel_loss_elem(ha.y, sa.y, g_mse, g_mae, sq, ab, ga.y);
      pixel_loss_elem(ha.z, sa.z, g_mse, g_mae, sq, ab, ga.z);
      pixel_loss_elem(ha.w, sa.w, g_mse, g_mae, sq, ab, ga.w);
      if (g4) g4[i] = ga;
      if (two) {
        pixel_loss_elem(hb.x, sb.x, g_mse, g_mae, sq, ab, gb.x);
        pixel_loss_elem(hb.y, sb.y, g_mse, g_mae, sq, ab, gb.y);
        pixel_loss_elem(hb.z, sb.z, g_mse, g_mae, sq, ab, gb.z);
        pixel_loss_elem(hb.w, sb.w, g_mse, g_mae, sq, ab, gb.w);
        if (g4) g4[j] = gb;
      }
    }
    i0 = lo + (n4 << 2);
  }
  for (int64_t i = i0 + threadIdx.x; i < hi; i += kLossThreads) {
    float g;
    pixel_loss_elem(h[i], s[i], g_mse, g_mae, sq, ab, g);
    if (gr) gr[i] = g;
  }
  __shared__ float s_sq[kLossThreads / 32], s_ab[kLossThreads / 32];
  sq = warp_sum(sq);
  ab = warp_sum(ab);
  if ((threadIdx.x & 31) == 0) {
    s_sq[threadIdx.x >> 5] = sq;
    s_ab[threadIdx.x >> 5] = ab;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, c = 0.f;
    for (int w = 0; w < kLossThreads / 32; ++w) {
      a += s_sq[w];
      c += s_ab[w];
    }
    partial[img * kLossBlocksPerImage + b] = make_float2(a, c);
  }
}

// stage 2 (one block): out[0] = mse, out[1] = mae (global means), out[2 + i] = PSNR of image i
__global__ void pixel_loss_final_kernel(const float2* __restrict__ partial, int n, int64_t per_image, float max_val,
                                        float* __restrict__ out) {
  for (int img = threadIdx.x; img < n; img += blockDim.x) {
    double sq = 0.0;
    for (int b = 0; b < kLossBlocksPerImage; ++b) sq += partial[img * kLossBlocksPerImage + b].x;
    const double mse_i = sq / static_cast<double>(per_image);
    out[2 + img] = static_cast<float>(20.0 * log10(static_cast<double>(max_val)) - 10.0 * log10(mse_i));
  }
  if (threadIdx.x == 0) {  // global sums in a fixed order
    double sq = 0.0, ab = 0.0;
    for (int i = 0; i < n * kLossBlocksPerImage; ++i) {
      sq += partial[i].x;
      ab += partial[i].y;
    }
    const double tot = static_cast<double>(per_image) * n;
    out[0] = static_cast<float>(sq / tot);
    out[1] = static_cast<float>(ab / tot);
  }
}

// ---------------------------------------------------------------- total variation (vgg_loss.py:166-169)
// tf.image.total_variation(x) = sum |x[i+1,j]-x[i,j]| + sum |x[i,j+1]-x[i,j]| per image (anisotropic L1); the reference
// adds weight * reduce_sum over the batch of it, on the de-normalised image 127.5 * (sr + 1): value_scale = 127.5.
// grad (optional, accumulated): d/d sr of weight * value_scale * TV(sr).  Two-stage fixed-order sum (block partials).
// A block walks whole image rows (blockIdx.x, + gridDim.x, ...): the neighbours of element e of a row are e +- c in
// the same row and e in the rows above / below - no index division per element; the three rows stay in L1 / L2.
constexpr int kTvBlocks = 148 * 8;
__global__ void __launch_bounds__(256) total_variation_partial_kernel(const float* __restrict__ x, int n, int h, int w, int c,
                                                                       float gscale, float* __restrict__ grad,
                                                                       float* __restrict__ partial) {
  const int row = w * c;
  const int rows = n * h;
  float acc = 0.f;
  auto sgn = [](float d) { return d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f); };
  for (int r = blockIdx.x; r < rows; r += gridDim.x) {
    const int yy = r % h;
    const float* xr = x + static_cast<int64_t>(r) * row;
    float* gr = grad ? grad + static_cast<int64_t>(r) * row : nullptr;
    const bool up = yy > 0, down = yy + 1 < h;
    for (int e = threadIdx.x; e < row; e += blockDim.x) {
      const float v = xr[e];
      float g = 0.f;
      if (down) {
        const float d = __ldg(xr + row + e) - v;
        acc += fabsf(d);
        g -= sgn(d);
      }
      if (e + c < row) {
        const float d = xr[e + c] - v;
        acc += fabsf(d);
        g -= sgn(d);
      }
      if (up) g += sgn(v - __ldg(xr - row + e));
      if (e >= c) g += sgn(v - xr[e - c]);
      if (gr) gr[e] += gscale * g;
    }
  }
  __shared__ float s_acc[8];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) s_acc[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int k = 0; k < 8; ++k) t += s_acc[k];
    partial[blockIdx.x] = t;
  }
}
// one warp: lane l sums partials [l * per, (l + 1) * per) in order (double), then a fixed shuffle tree
__global__ void total_variation_final_kernel(const float* __restrict__ partial, int nblocks, float scale,
                                             float* __restrict__ out) {
  const int per = (nblocks + 31) / 32;
  double t = 0.0;
  for (int b = threadIdx.x * per; b < min(nblocks, (threadIdx.x + 1) * per); ++b) t += partial[b];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  if (threadIdx.x == 0) out[0] = static_cast<float>(t * scale);
}

// ---------------------------------------------------------------- Adam (Keras OptimizerV2 semantics)
// theta -= lr_t * m / (sqrt(v) + eps), lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t) computed on the host.
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, int64_t count, float lr_t, float b1, float b2, float eps,
                            float grad_scale) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < count;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float gi = g[i] * grad_scale;
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] = p[i] - lr_t * mi / (sqrtf(vi) + eps);
  }
}

// ---------------------------------------------------------------- per-channel sums over pixels
// out[c] = sum_p x[p, coff + c] (bias gradient) or sum_p x[p,c] * min(0, z[p,c]) (PReLU slope gradient; z != nullptr).
// Two-stage, fixed order: block b sums pixels [b*chunk, (b+1)*chunk) for all channels, then one block adds partials.
constexpr int kChanBlocks = 148;
__global__ void channel_sum_partial_kernel(const __nv_bfloat16* __restrict__ x, int xcs, int xoff,
                                           const __nv_bfloat16* __restrict__ z, int zcs, int zoff, int64_t pixels, int c,
                                           float* __restrict__ partial) {
  const int64_t chunk = (pixels + gridDim.x - 1) / gridDim.x;
  const int64_t lo = blockIdx.x * chunk, hi = min(pixels, lo + chunk);
  // thread t owns channel t % c and every (blockDim.x / c)-th pixel of the block's range
  const int ch = threadIdx.x % c;
  const int lanes = blockDim.x / c;
  const int pl = threadIdx.x / c;
  float acc = 0.f;
  if (pl < lanes) {
    for (int64_t p = lo + pl; p < hi; p += lanes) {
      float v = __bfloat162float(x[p * xcs + xoff + ch]);
      if (z != nullptr) v *= fminf(__bfloat162float(z[p * zcs + zoff + ch]), 0.f);
      acc += v;
    }
  }
  extern __shared__ float s_acc[];
  s_acc[threadIdx.x] = (pl < lanes) ? acc : 0.f;
  __syncthreads();
  if (threadIdx.x < c) {
    float t = 0.f;
    for (int l = 0; l < lanes; ++l) t += s_acc[l * c + threadIdx.x];
    partial[blockIdx.x * c + threadIdx.x] = t;
  }
}
// 16-byte variant (c, strides and offsets multiples of 8): a thread owns 8 consecutive channels and every lanes-th pixel
__global__ void channel_sum_partial_v8_kernel(const __nv_bfloat16* __restrict__ x, int xcs, int xoff,
                                              const __nv_bfloat16* __restrict__ z, int zcs, int zoff, int64_t pixels, int c,
                                              float* __restrict__ partial) {
  const int64_t chunk = (pixels + gridDim.x - 1) / gridDim.x;
  const int64_t lo = blockIdx.x * chunk, hi = min(pixels, lo + chunk);
  const int groups = c >> 3;
  const int cg = threadIdx.x % groups, lanes = blockDim.x / groups, pl = threadIdx.x / groups;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  if (pl < lanes) {
    for (int64_t p = lo + pl; p < hi; p += lanes) {
      const uint4 qx = __ldg(reinterpret_cast<const uint4*>(x + p * xcs + xoff) + cg);
      const uint32_t wx[4] = {qx.x, qx.y, qx.z, qx.w};
      if (z != nullptr) {
        const uint4 qz = __ldg(reinterpret_cast<const uint4*>(z + p * zcs + zoff) + cg);
        const uint32_t wz[4] = {qz.x, qz.y, qz.z, qz.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          acc[2 * k] += __uint_as_float(wx[k] << 16) * fminf(__uint_as_float(wz[k] << 16), 0.f);
          acc[2 * k + 1] += __uint_as_float(wx[k] & 0xFFFF0000u) * fminf(__uint_as_float(wz[k] & 0xFFFF0000u), 0.f);
        }
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          acc[2 * k] += __uint_as_float(wx[k] << 16);
          acc[2 * k + 1] += __uint_as_float(wx[k] & 0xFFFF0000u);
        }
      }
    }
  }
  extern __shared__ float s_acc[];  // [lanes][c]
  if (pl < lanes) {
#pragma unroll
    for (int j = 0; j < 8; ++j) s_acc[pl * c + cg * 8 + j] = acc[j];
  }
  __syncthreads();
  for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
    float t = 0.f;
    for (int l = 0; l < lanes; ++l) t += s_acc[l * c + ch];
    partial[blockIdx.x * c + ch] = t;
  }
}
// block partials of a channel: 8 lanes (block b -> lane b % 8), lane sums added in lane order (fixed order)
__global__ void __launch_bounds__(256) channel_sum_final_kernel(const float* __restrict__ partial, int nblocks, int c,
                                                                float scale, float* __restrict__ out, int accumulate) {
  __shared__ double sm[256];
  const int ch = blockIdx.x * 32 + (threadIdx.x >> 3), lane8 = threadIdx.x & 7;
  double t = 0.0;
  if (ch < c)
    for (int b = lane8; b < nblocks; b += 8) t += partial[b * c + ch];
  sm[threadIdx.x] = t;
  __syncthreads();
  if (lane8 == 0 && ch < c) {
    double tot = 0.0;
    for (int l = 0; l < 8; ++l) tot += sm[threadIdx.x + l];
    const float r = static_cast<float>(tot) * scale;
    out[ch] = accumulate ? out[ch] + r : r;
  }
}

// ---------------------------------------------------------------- f32 [pixels,c] -> bf16 slice (gradient of the model output)
__global__ void f32_to_bf16_slice_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, int ycs, int yoff,
                                         int64_t pixels, int c) {
  const int64_t total = pixels * c;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t p = i / c;
    const int ch = static_cast<int>(i % c);
    y[p * ycs + yoff + ch] = __float2bfloat16_rn(x[i]);
  }
}

// ---------------------------------------------------------------- activation backward on bf16 slices (8 channels / thread)
// dz = dy * (z > 0 ? 1 : slope[c]); z is the forward PRE-activation for PReLU (its slopes start at 0, so the sign of the
// output does not determine min(0, x)) and may be the post-activation for LeakyReLU (same sign).
__global__ void act_bwd_kernel(const __nv_bfloat16* __restrict__ dy, int dcs, int doff,
                               const __nv_bfloat16* __restrict__ z, int zcs, int zoff, const float* __restrict__ alpha,
                               float alpha_s, __nv_bfloat16* __restrict__ dz, int ocs, int ooff, int64_t pixels, int c) {
  // PDL (see axpby_bf16_kernel): wait for the producer before any memory access, let the next conv's prologue start
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int groups = c / 8;
  const int64_t total = pixels * groups;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t pix = i / groups;
    const int g = static_cast<int>(i % groups);
    const uint4 qd = *reinterpret_cast<const uint4*>(dy + pix * dcs + doff + g * 8);
    const uint4 qz = *reinterpret_cast<const uint4*>(z + pix * zcs + zoff + g * 8);
    const uint32_t wd[4] = {qd.x, qd.y, qd.z, qd.w};
    const uint32_t wz[4] = {qz.x, qz.y, qz.z, qz.w};
    uint32_t wo[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float a0 = alpha ? alpha[g * 8 + 2 * k] : alpha_s, a1 = alpha ? alpha[g * 8 + 2 * k + 1] : alpha_s;
      const float lo = bf16_lo(wd[k]) * (bf16_lo(wz[k]) > 0.f ? 1.f : a0);
      const float hi = bf16_hi(wd[k]) * (bf16_hi(wz[k]) > 0.f ? 1.f : a1);
      wo[k] = pack_bf16x2(lo, hi);
    }
    *reinterpret_cast<uint4*>(dz + pix * ocs + ooff + g * 8) = make_uint4(wo[0], wo[1], wo[2], wo[3]);
  }
}

// y = z > 0 ? z : slope[c] * z  (PReLU / LeakyReLU forward on a stored pre-activation; training keeps z for the backward)
__global__ void act_fwd_kernel(const __nv_bfloat16* __restrict__ z, int zcs, int zoff, const float* __restrict__ alpha,
                               float alpha_s, __nv_bfloat16* __restrict__ y, int ocs, int ooff, int64_t pixels, int c) {
  const int groups = c / 8;
  const int64_t total = pixels * groups;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t pix = i / groups;
    const int g = static_cast<int>(i % groups);
    const uint4 qz = *reinterpret_cast<const uint4*>(z + pix * zcs + zoff + g * 8);
    const uint32_t wz[4] = {qz.x, qz.y, qz.z, qz.w};
    uint32_t wo[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float a0 = alpha ? alpha[g * 8 + 2 * k] : alpha_s, a1 = alpha ? alpha[g * 8 + 2 * k + 1] : alpha_s;
      const float lo = bf16_lo(wz[k]), hi = bf16_hi(wz[k]);
      wo[k] = pack_bf16x2(lo > 0.f ? lo : a0 * lo, hi > 0.f ? hi : a1 * hi);
    }
    *reinterpret_cast<uint4*>(y + pix * ocs + ooff + g * 8) = make_uint4(wo[0], wo[1], wo[2], wo[3]);
  }
}

// ---------------------------------------------------------------- space_to_depth(2) (gradient of depth_to_space), 16-byte vectors
// y[n,h,w,(2i+j)*C + c] = x[n,2h+i,2w+j,c]
__global__ void s2d2_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int n, int h, int w, int cv) {
  const int64_t total = static_cast<int64_t>(n) * h * w * 4 * cv;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % cv);
    int64_t q = i / cv;
    const int sub = static_cast<int>(q % 4);
    q /= 4;
    const int iw = static_cast<int>(q % w);
    q /= w;
    const int ih = static_cast<int>(q % h);
    const int nn = static_cast<int>(q / h);
    const int64_t src = ((static_cast<int64_t>(nn) * 2 * h + 2 * ih + (sub >> 1)) * (2 * w) + 2 * iw + (sub & 1)) * cv + c;
    y[i] = __ldg(x + src);
  }
}

// dz[i] = g[i] * (1 - y[i]^2): backward through the tanh output activation, fp32 (model_builder.py:93,133)
__global__ void tanh_bwd_f32_kernel(const float* __restrict__ g, const float* __restrict__ y, float* __restrict__ dz,
                                    int64_t count) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < count;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    dz[i] = g[i] * (1.f - y[i] * y[i]);
}

// ---------------------------------------------------------------- VGG19 perceptual loss: edge kernels (vgg_loss.py:144-148)
// y[p, 0..2] = ((x[p, 2-c] + 1) * 127.5 - mean_bgr[c]) : denormalise [-1,1] -> [0,255], RGB -> BGR, subtract the ImageNet
// means (keras.applications.vgg19.preprocess_input, caffe mode); bf16, padded to 16 channels (zeros).
__global__ void vgg_preprocess_kernel(const float* __restrict__ x, uint4* __restrict__ y, int64_t pixels) {
  for (int64_t p = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; p < pixels;
       p += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float r = x[3 * p], g = x[3 * p + 1], b = x[3 * p + 2];
    const float v0 = (b + 1.f) * 127.5f - 103.939f, v1 = (g + 1.f) * 127.5f - 116.779f, v2 = (r + 1.f) * 127.5f - 123.68f;
    y[2 * p] = make_uint4(pack_bf16x2(v0, v1), pack_bf16x2(v2, 0.f), 0u, 0u);
    y[2 * p + 1] = make_uint4(0u, 0u, 0u, 0u);
  }
}
// g[p, c] (+)= scale * 127.5 * dy[p, 2-c]   (dy: fp32 [pixels, 3] gradient of the preprocessed BGR image)
__global__ void vgg_preprocess_bwd_kernel(const float* __restrict__ dy, float* __restrict__ g, int64_t pixels,
                                          float scale, int accumulate) {
  const int64_t total = pixels * 3;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t p = i / 3;
    const int c = static_cast<int>(i % 3);
    const float v = scale * 127.5f * dy[3 * p + (2 - c)];
    g[i] = accumulate ? g[i] + v : v;
  }
}

// ---------------------------------------------------------------- MaxPooling2D(2, 2) VALID, bf16 NHWC, 8 channels / thread
__global__ void maxpool2_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, int n, int h, int w,
                                int c) {
  const int oh = h >> 1, ow = w >> 1, groups = c / 8;
  const int64_t total = static_cast<int64_t>(n) * oh * ow * groups;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int g = static_cast<int>(i % groups);
    int64_t q = i / groups;
    const int ox = static_cast<int>(q % ow);
    q /= ow;
    const int oy = static_cast<int>(q % oh);
    const int nn = static_cast<int>(q / oh);
    const __nv_bfloat16* base = x + ((static_cast<int64_t>(nn) * h + 2 * oy) * w + 2 * ox) * c + g * 8;
    const uint4 a = *reinterpret_cast<const uint4*>(base), b = *reinterpret_cast<const uint4*>(base + c);
    const uint4 d = *reinterpret_cast<const uint4*>(base + static_cast<int64_t>(w) * c),
                e = *reinterpret_cast<const uint4*>(base + static_cast<int64_t>(w) * c + c);
    const uint32_t wa[4] = {a.x, a.y, a.z, a.w}, wb[4] = {b.x, b.y, b.z, b.w}, wd[4] = {d.x, d.y, d.z, d.w},
                   we[4] = {e.x, e.y, e.z, e.w};
    uint32_t wo[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float lo = fmaxf(fmaxf(bf16_lo(wa[k]), bf16_lo(wb[k])), fmaxf(bf16_lo(wd[k]), bf16_lo(we[k])));
      const float hi = fmaxf(fmaxf(bf16_hi(wa[k]), bf16_hi(wb[k])), fmaxf(bf16_hi(wd[k]), bf16_hi(we[k])));
      wo[k] = pack_bf16x2(lo, hi);
    }
    *reinterpret_cast<uint4*>(y + i * 8) = make_uint4(wo[0], wo[1], wo[2], wo[3]);
  }
}
// MaxPoolGrad: dx[window] = dy at the FIRST maximum of the window (row-major), 0 elsewhere; one thread per (window, channel)
__global__ void maxpool2_bwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy,
                                    __nv_bfloat16* __restrict__ dx, int n, int h, int w, int c) {
  const int oh = h >> 1, ow = w >> 1;
  const int64_t total = static_cast<int64_t>(n) * oh * ow * c;
  const __nv_bfloat16 zero = __float2bfloat16_rn(0.f);
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int ch = static_cast<int>(i % c);
    int64_t q = i / c;
    const int ox = static_cast<int>(q % ow);
    q /= ow;
    const int oy = static_cast<int>(q % oh);
    const int nn = static_cast<int>(q / oh);
    const int64_t b00 = ((static_cast<int64_t>(nn) * h + 2 * oy) * w + 2 * ox) * c + ch;
    const int64_t idx[4] = {b00, b00 + c, b00 + static_cast<int64_t>(w) * c, b00 + static_cast<int64_t>(w) * c + c};
    int best = 0;
    float bv = __bfloat162float(x[idx[0]]);
#pragma unroll
    for (int k = 1; k < 4; ++k) {
      const float v = __bfloat162float(x[idx[k]]);
      if (v > bv) {
        bv = v;
        best = k;
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) dx[idx[k]] = (k == best) ? dy[i] : zero;
  }
}

// y[i] += a * x[i], fp32
__global__ void axpy_f32_kernel(const float* __restrict__ x, float* __restrict__ y, float a, int64_t count) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < count;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    y[i] += a * x[i];
}

}  // namespace ssr

using namespace ssr;

#define SSR_CHECK_LAUNCH(name)                                                          \
  do {                                                                                  \
    cudaError_t e__ = cudaGetLastError();                                               \
    if (e__ != cudaSuccess) return set_error(SSR_ERR_CUDA, name ": %s", cudaGetErrorString(e__)); \
  } while (0)

extern "C" size_t ssr_pixel_loss_workspace_bytes(int n) {
  return static_cast<size_t>(n > 0 ? n : 1) * kLossBlocksPerImage * sizeof(float2);
}

extern "C" int ssr_pixel_loss(const float* hr, const float* sr, int n, int64_t per_image, float w_mse, float w_mae,
                              float max_val, float* grad, void* workspace, float* out, void* stream) {
  if (!hr || !sr || !workspace || !out || n <= 0 || per_image <= 0)
    return set_error(SSR_ERR_INVALID, "pixel_loss: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const double tot = static_cast<double>(per_image) * n;
  const float g_mse = static_cast<float>(2.0 * w_mse / tot), g_mae = static_cast<float>(w_mae / tot);
  pixel_loss_partial_kernel<<<dim3(kLossBlocksPerImage, n), kLossThreads, 0, st>>>(
      hr, sr, per_image, g_mse, g_mae, grad, static_cast<float2*>(workspace));
  SSR_CHECK_LAUNCH("pixel_loss_partial");
  pixel_loss_final_kernel<<<1, 128, 0, st>>>(static_cast<const float2*>(workspace), n, per_image, max_val, out);
  SSR_CHECK_LAUNCH("pixel_loss_final");
  return SSR_OK;
}

extern "C" size_t ssr_total_variation_workspace_bytes(void) { return kTvBlocks * sizeof(float); }

extern "C" int ssr_total_variation(const float* x, int n, int h, int w, int c, float value_scale, float weight, float* grad,
                                   void* workspace, float* out1, void* stream) {
  if (!x || !workspace || !out1 || n <= 0 || h <= 0 || w <= 0 || c <= 0)
    return set_error(SSR_ERR_INVALID, "total_variation: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const float sc = value_scale * weight;
  total_variation_partial_kernel<<<kTvBlocks, 256, 0, st>>>(x, n, h, w, c, sc, grad, static_cast<float*>(workspace));
  SSR_CHECK_LAUNCH("total_variation_partial");
  total_variation_final_kernel<<<1, 32, 0, st>>>(static_cast<const float*>(workspace), kTvBlocks, sc, out1);
  SSR_CHECK_LAUNCH("total_variation_final");
  return SSR_OK;
}

extern "C" int ssr_adam_step(float* param, const float* grad, float* m, float* v, int64_t count, float lr_t,
                             float beta1, float beta2, float eps, float grad_scale, void* stream) {
  if (!param || !grad || !m || !v || count < 0) return set_error(SSR_ERR_INVALID, "adam_step: bad argument");
  if (count == 0) return SSR_OK;
  const int block = 256;
  int64_t g = (count + block - 1) / block;
  if (g > 148 * 8) g = 148 * 8;
  adam_kernel<<<static_cast<int>(g), block, 0, static_cast<cudaStream_t>(stream)>>>(param, grad, m, v, count, lr_t,
                                                                                   beta1, beta2, eps, grad_scale);
  SSR_CHECK_LAUNCH("adam_step");
  return SSR_OK;
}

extern "C" size_t ssr_channel_sum_workspace_bytes(int c) {
  return static_cast<size_t>(kChanBlocks) * (c > 0 ? c : 1) * sizeof(float);
}

extern "C" int ssr_channel_sum_bf16(const void* x, int x_cstride, int x_coff, const void* z, int z_cstride, int z_coff,
                                    int64_t pixels, int c, float scale, int accumulate, void* workspace, float* out,
                                    void* stream) {
  if (!x || !workspace || !out || pixels <= 0 || c <= 0 || c > 512)
    return set_error(SSR_ERR_INVALID, "channel_sum: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int block = (c <= 256) ? 256 : 512;
  const int nblocks = static_cast<int>(pixels < kChanBlocks ? pixels : kChanBlocks);
  const bool v8 = (c % 8 == 0) && (x_cstride % 8 == 0) && (x_coff % 8 == 0) && (!z || (z_cstride % 8 == 0 && z_coff % 8 == 0)) &&
                  (reinterpret_cast<uintptr_t>(x) % 16 == 0) && (!z || reinterpret_cast<uintptr_t>(z) % 16 == 0);
  if (v8) {
    const int lanes = 256 / (c / 8) > 0 ? 256 / (c / 8) : 1;
    channel_sum_partial_v8_kernel<<<nblocks, (c / 8) * lanes, static_cast<size_t>(lanes) * c * sizeof(float), st>>>(
        static_cast<const __nv_bfloat16*>(x), x_cstride, x_coff, static_cast<const __nv_bfloat16*>(z), z_cstride, z_coff,
        pixels, c, static_cast<float*>(workspace));
  } else {
    channel_sum_partial_kernel<<<nblocks, block, block * sizeof(float), st>>>(
        static_cast<const __nv_bfloat16*>(x), x_cstride, x_coff, static_cast<const __nv_bfloat16*>(z), z_cstride, z_coff,
        pixels, c, static_cast<float*>(workspace));
  }
  SSR_CHECK_LAUNCH("channel_sum_partial");
  channel_sum_final_kernel<<<(c + 31) / 32, 256, 0, st>>>(static_cast<const float*>(workspace), nblocks, c, scale, out,
                                                          accumulate);
  SSR_CHECK_LAUNCH("channel_sum_final");
  return SSR_OK;
}

extern "C" int ssr_f32_to_bf16_slice(const float* x, void* y, int y_cstride, int y_coff, int64_t pixels, int c,
                                     void* stream) {
  if (!x || !y || pixels < 0 || c <= 0 || y_cstride < y_coff + c)
    return set_error(SSR_ERR_INVALID, "f32_to_bf16_slice: bad shape");
  if (pixels == 0) return SSR_OK;
  const int block = 256;
  int64_t g = (pixels * c + block - 1) / block;
  if (g > 148 * 8) g = 148 * 8;
  f32_to_bf16_slice_kernel<<<static_cast<int>(g), block, 0, static_cast<cudaStream_t>(stream)>>>(
      x, static_cast<__nv_bfloat16*>(y), y_cstride, y_coff, pixels, c);
  SSR_CHECK_LAUNCH("f32_to_bf16_slice");
  return SSR_OK;
}

extern "C" int ssr_act_bwd_bf16(const void* dy, int dy_cstride, int dy_coff, const void* z, int z_cstride, int z_coff,
                                const float* alpha, float alpha_scalar, void* dz, int dz_cstride, int dz_coff,
                                int64_t pixels, int c, void* stream) {
  if (!dy || !z || !dz || pixels < 0 || c <= 0 || c % 8 || dy_cstride % 8 || dy_coff % 8 || z_cstride % 8 || z_coff % 8 ||
      dz_cstride % 8 || dz_coff % 8)
    return set_error(SSR_ERR_INVALID, "act_bwd: channel counts/offsets must be multiples of 8");
  if (pixels == 0) return SSR_OK;
  const int block = 256;
  int64_t g = (pixels * (c / 8) + block - 1) / block;
  if (g > 148 * 8) g = 148 * 8;
  cudaError_t le = launch_pdl(act_bwd_kernel, dim3(static_cast<int>(g)), dim3(block), 0, static_cast<cudaStream_t>(stream),
                              static_cast<const __nv_bfloat16*>(dy), dy_cstride, dy_coff,
                              static_cast<const __nv_bfloat16*>(z), z_cstride, z_coff, alpha, alpha_scalar,
                              static_cast<__nv_bfloat16*>(dz), dz_cstride, dz_coff, pixels, c);
  if (le != cudaSuccess) return set_error(SSR_ERR_CUDA, "act_bwd launch: %s", cudaGetErrorString(le));
  return SSR_OK;
}

extern "C" int ssr_space_to_depth2(const void* x, void* y, int n, int h, int w, int c, int elem_bytes, void* stream) {
  if (!x || !y || n < 0 || h < 0 || w < 0 || c <= 0 || (c * elem_bytes) % 16)
    return set_error(SSR_ERR_INVALID, "space_to_depth2: c * elem_bytes must be a multiple of 16");
  const int cv = c * elem_bytes / 16;
  const int64_t total = static_cast<int64_t>(n) * h * w * 4 * cv;
  if (total == 0) return SSR_OK;
  const int block = 256;
  int64_t g = (total + block - 1) / block;
  if (g > 148 * 16) g = 148 * 16;
  s2d2_kernel<<<static_cast<int>(g), block, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(x), static_cast<uint4*>(y), n, h, w, cv);
  SSR_CHECK_LAUNCH("space_to_depth2");
  return SSR_OK;
}

extern "C" int ssr_tanh_bwd_f32(const float* g, const float* y, float* dz, int64_t count, void* stream) {
  if (!g || !y || !dz || count < 0) return set_error(SSR_ERR_INVALID, "tanh_bwd: bad argument");
  if (count == 0) return SSR_OK;
  const int block = 256;
  int64_t gr = (count + block - 1) / block;
  if (gr > 148 * 8) gr = 148 * 8;
  tanh_bwd_f32_kernel<<<static_cast<int>(gr), block, 0, static_cast<cudaStream_t>(stream)>>>(g, y, dz, count);
  SSR_CHECK_LAUNCH("tanh_bwd");
  return SSR_OK;
}

extern "C" int ssr_act_fwd_bf16(const void* z, int z_cstride, int z_coff, const float* alpha, float alpha_scalar, void* y,
                                int y_cstride, int y_coff, int64_t pixels, int c, void* stream) {
  if (!z || !y || pixels < 0 || c <= 0 || c % 8 || z_cstride % 8 || z_coff % 8 || y_cstride % 8 || y_coff % 8)
    return set_error(SSR_ERR_INVALID, "act_fwd: channel counts/offsets must be multiples of 8");
  if (pixels == 0) return SSR_OK;
  const int block = 256;
  int64_t g = (pixels * (c / 8) + block - 1) / block;
  if (g > 148 * 8) g = 148 * 8;
  act_fwd_kernel<<<static_cast<int>(g), block, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(z), z_cstride, z_coff, alpha, alpha_scalar, static_cast<__nv_bfloat16*>(y),
      y_cstride, y_coff, pixels, c);
  SSR_CHECK_LAUNCH("act_fwd");
  return SSR_OK;
}

static inline int grid1d(int64_t work, int block, int waves = 8) {
  int64_t g = (work + block - 1) / block;
  if (g > 148 * waves) g = 148 * waves;
  return static_cast<int>(g < 1 ? 1 : g);
}

extern "C" int ssr_vgg_preprocess(const float* x, void* y_bf16_c16, int64_t pixels, void* stream) {
  if (!x || !y_bf16_c16 || pixels < 0) return set_error(SSR_ERR_INVALID, "vgg_preprocess: bad argument");
  if (pixels == 0) return SSR_OK;
  vgg_preprocess_kernel<<<grid1d(pixels, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, static_cast<uint4*>(y_bf16_c16), pixels);
  SSR_CHECK_LAUNCH("vgg_preprocess");
  return SSR_OK;
}

extern "C" int ssr_vgg_preprocess_bwd(const float* dy_bgr, float* g_rgb, int64_t pixels, float scale, int accumulate,
                                      void* stream) {
  if (!dy_bgr || !g_rgb || pixels < 0) return set_error(SSR_ERR_INVALID, "vgg_preprocess_bwd: bad argument");
  if (pixels == 0) return SSR_OK;
  vgg_preprocess_bwd_kernel<<<grid1d(pixels * 3, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(dy_bgr, g_rgb, pixels,
                                                                                                scale, accumulate);
  SSR_CHECK_LAUNCH("vgg_preprocess_bwd");
  return SSR_OK;
}

extern "C" int ssr_maxpool2_bf16(const void* x, void* y, int n, int h, int w, int c, void* stream) {
  if (!x || !y || n < 0 || h < 2 || w < 2 || c <= 0 || c % 8) return set_error(SSR_ERR_INVALID, "maxpool2: bad shape");
  const int64_t total = static_cast<int64_t>(n) * (h / 2) * (w / 2) * (c / 8);
  if (total == 0) return SSR_OK;
  maxpool2_kernel<<<grid1d(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(y), n, h, w, c);
  SSR_CHECK_LAUNCH("maxpool2");
  return SSR_OK;
}

extern "C" int ssr_maxpool2_bwd_bf16(const void* x, const void* dy, void* dx, int n, int h, int w, int c, void* stream) {
  if (!x || !dy || !dx || n < 0 || h < 2 || w < 2 || (h & 1) || (w & 1) || c <= 0)
    return set_error(SSR_ERR_INVALID, "maxpool2_bwd: bad shape (even h, w required)");
  const int64_t total = static_cast<int64_t>(n) * (h / 2) * (w / 2) * c;
  if (total == 0) return SSR_OK;
  maxpool2_bwd_kernel<<<grid1d(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(dy), static_cast<__nv_bfloat16*>(dx), n, h,
      w, c);
  SSR_CHECK_LAUNCH("maxpool2_bwd");
  return SSR_OK;
}

extern "C" int ssr_axpy_f32(const float* x, float* y, float a, int64_t count, void* stream) {
  if (!x || !y || count < 0) return set_error(SSR_ERR_INVALID, "axpy_f32: bad argument");
  if (count == 0) return SSR_OK;
  axpy_f32_kernel<<<grid1d(count, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, y, a, count);
  SSR_CHECK_LAUNCH("axpy_f32");
  return SSR_OK;
}
