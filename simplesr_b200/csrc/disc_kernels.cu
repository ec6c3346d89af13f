// disc_kernels.cu — bandwidth-bound pieces of the ESRGAN discriminator step (model_builder.build_discriminator
// :137-198, discriminator.py:147-172, ra_adversarial_loss.py:59-70, ra_discriminator_loss.py:55-66):
// stride-2 decimation / zero insertion around the stride-1 conv kernels, BatchNormalization (training statistics)
// fused with LeakyReLU forward and backward, the two Dense layers (batch <= 32: HBM-bound on the 134 MB weight matrix, so
// they run on CUDA cores with coalesced weight streaming, not on tensor cores), and the relativistic-average losses.
// Every reduction is two-stage with a fixed order (deterministic).
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstring>

#include "comm_dev.cuh"
#include "internal.h"
#include "ptx_sm100.cuh"

namespace ssr {

static inline int grid1(int64_t work, int block, int waves = 8) {
  int64_t g = (work + block - 1) / block;
  if (g > 148 * waves) g = 148 * waves;
  return static_cast<int>(g < 1 ? 1 : g);
}

// ---------------------------------------------------------------- stride 2 through the stride-1 kernels
// Conv2D(strides=2, padding="same") on even sizes pads (0, 1): out[y, x] = full[2y + 1, 2x + 1] of the stride-1 SAME conv.
__global__ void subsample2_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int n, int oh, int ow, int cv) {
  const int64_t total = static_cast<int64_t>(n) * oh * ow * cv;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % cv);
    int64_t q = i / cv;
    const int ox = static_cast<int>(q % ow);
    q /= ow;
    const int oy = static_cast<int>(q % oh);
    const int nn = static_cast<int>(q / oh);
    y[i] = __ldg(x + ((static_cast<int64_t>(nn) * 2 * oh + 2 * oy + 1) * (2 * ow) + 2 * ox + 1) * cv + c);
  }
}
// its adjoint: dx[2y+1, 2x+1] = dy[y, x], zero elsewhere (one thread per full-resolution vector)
__global__ void zero_insert2_kernel(const uint4* __restrict__ dy, uint4* __restrict__ dx, int n, int oh, int ow, int cv) {
  const int64_t total = static_cast<int64_t>(n) * 2 * oh * 2 * ow * cv;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % cv);
    int64_t q = i / cv;
    const int x = static_cast<int>(q % (2 * ow));
    q /= 2 * ow;
    const int y = static_cast<int>(q % (2 * oh));
    const int nn = static_cast<int>(q / (2 * oh));
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if ((x & 1) && (y & 1)) v = __ldg(dy + ((static_cast<int64_t>(nn) * oh + (y >> 1)) * ow + (x >> 1)) * cv + c);
    dx[i] = v;
  }
}

// ---------------------------------------------------------------- BatchNormalization (training=True) + LeakyReLU
constexpr int kBnBlocks = 148;
// partial[b][0][c] = sum x, partial[b][1][c] = sum x^2 over the block's pixel range (stats), or
// sum d', sum d' * xhat with d' = dy * lrelu'(y) (backward reduce; mean/istd given)
template <bool BWD>
__global__ void bn_reduce_partial_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy,
                                         const __nv_bfloat16* __restrict__ y, const float* __restrict__ mean,
                                         const float* __restrict__ istd, float alpha, int64_t pixels, int c,
                                         float* __restrict__ partial) {
  const int64_t chunk = (pixels + gridDim.x - 1) / gridDim.x;
  const int64_t lo = blockIdx.x * chunk, hi = min(pixels, lo + chunk);
  const int ch = threadIdx.x % c, lanes = blockDim.x / c, pl = threadIdx.x / c;
  float a0 = 0.f, a1 = 0.f;
  if (pl < lanes) {
    const float mu = BWD ? mean[ch] : 0.f, is = BWD ? istd[ch] : 0.f;
    for (int64_t p = lo + pl; p < hi; p += lanes) {
      const float xv = __bfloat162float(x[p * c + ch]);
      if (BWD) {
        const float d = __bfloat162float(dy[p * c + ch]) * (__bfloat162float(y[p * c + ch]) > 0.f ? 1.f : alpha);
        a0 += d;
        a1 += d * (xv - mu) * is;
      } else {
        a0 += xv;
        a1 += xv * xv;
      }
    }
  }
  extern __shared__ float sm[];
  sm[threadIdx.x] = (pl < lanes) ? a0 : 0.f;
  sm[blockDim.x + threadIdx.x] = (pl < lanes) ? a1 : 0.f;
  __syncthreads();
  if (threadIdx.x < c) {
    float t0 = 0.f, t1 = 0.f;
    for (int l = 0; l < lanes; ++l) {
      t0 += sm[l * c + threadIdx.x];
      t1 += sm[blockDim.x + l * c + threadIdx.x];
    }
    partial[(blockIdx.x * 2) * c + threadIdx.x] = t0;
    partial[(blockIdx.x * 2 + 1) * c + threadIdx.x] = t1;
  }
}
// Same for c % 8 == 0: a thread owns 8 consecutive channels (one 16-byte load per tensor) and every lanes-th pixel of the
// block's range; the lanes are added in a fixed order through shared memory.
template <bool BWD>
__global__ void bn_reduce_partial_v8_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy,
                                            const __nv_bfloat16* __restrict__ y, const float* __restrict__ mean,
                                            const float* __restrict__ istd, float alpha, int64_t pixels, int c,
                                            float* __restrict__ partial) {
  const int64_t chunk = (pixels + gridDim.x - 1) / gridDim.x;
  const int64_t lo = blockIdx.x * chunk, hi = min(pixels, lo + chunk);
  const int groups = c >> 3;
  const int cg = threadIdx.x % groups, lanes = blockDim.x / groups, pl = threadIdx.x / groups;
  float a0[8], a1[8], mu[8], is[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    a0[j] = a1[j] = 0.f;
    mu[j] = BWD ? mean[cg * 8 + j] : 0.f;
    is[j] = BWD ? istd[cg * 8 + j] : 0.f;
  }
  if (pl < lanes) {
    for (int64_t p = lo + pl; p < hi; p += lanes) {
      const uint4 qx = __ldg(reinterpret_cast<const uint4*>(x + p * c) + cg);
      const uint32_t wx[4] = {qx.x, qx.y, qx.z, qx.w};
      if (BWD) {
        const uint4 qd = __ldg(reinterpret_cast<const uint4*>(dy + p * c) + cg);
        const uint4 qy = __ldg(reinterpret_cast<const uint4*>(y + p * c) + cg);
        const uint32_t wd[4] = {qd.x, qd.y, qd.z, qd.w}, wy[4] = {qy.x, qy.y, qy.z, qy.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float d0 = __uint_as_float(wd[k] << 16) * (__uint_as_float(wy[k] << 16) > 0.f ? 1.f : alpha);
          const float d1 = __uint_as_float(wd[k] & 0xFFFF0000u) * (__uint_as_float(wy[k] & 0xFFFF0000u) > 0.f ? 1.f : alpha);
          a0[2 * k] += d0;
          a0[2 * k + 1] += d1;
          a1[2 * k] += d0 * (__uint_as_float(wx[k] << 16) - mu[2 * k]) * is[2 * k];
          a1[2 * k + 1] += d1 * (__uint_as_float(wx[k] & 0xFFFF0000u) - mu[2 * k + 1]) * is[2 * k + 1];
        }
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float x0 = __uint_as_float(wx[k] << 16), x1 = __uint_as_float(wx[k] & 0xFFFF0000u);
          a0[2 * k] += x0;
          a0[2 * k + 1] += x1;
          a1[2 * k] += x0 * x0;
          a1[2 * k + 1] += x1 * x1;
        }
      }
    }
  }
  extern __shared__ float sm[];  // [2][lanes][c]
  if (pl < lanes) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      sm[pl * c + cg * 8 + j] = a0[j];
      sm[(lanes + pl) * c + cg * 8 + j] = a1[j];
    }
  }
  __syncthreads();
  for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
    float t0 = 0.f, t1 = 0.f;
    for (int l = 0; l < lanes; ++l) {
      t0 += sm[l * c + ch];
      t1 += sm[(lanes + l) * c + ch];
    }
    partial[(blockIdx.x * 2) * c + ch] = t0;
    partial[(blockIdx.x * 2 + 1) * c + ch] = t1;
  }
}
// stats: mean, istd = 1/sqrt(var_biased + eps); moving stats (momentum m): mv = mv*m + batch*(1-m), variance unbiased
// The per-block partials of one channel are added by 8 lanes (block b -> lane b % 8) and the 8 lane sums in lane order:
// fixed order, 8x shorter dependent chain than one thread per channel.  blockDim = 256 = 32 channels x 8 lanes.
__device__ __forceinline__ void bn_final_sums(const float* __restrict__ partial, int nblocks, int c, int ch, int lane8,
                                              double* sm_s, double* sm_ss, double& s, double& ss) {
  double a = 0.0, b = 0.0;
  if (ch < c) {
    for (int k = lane8; k < nblocks; k += 8) {
      a += partial[(k * 2) * c + ch];
      b += partial[(k * 2 + 1) * c + ch];
    }
  }
  sm_s[threadIdx.x] = a;
  sm_ss[threadIdx.x] = b;
  __syncthreads();
  s = ss = 0.0;
  if (lane8 == 0) {
    for (int l = 0; l < 8; ++l) {
      s += sm_s[threadIdx.x + l];
      ss += sm_ss[threadIdx.x + l];
    }
  }
}
__global__ void __launch_bounds__(256) bn_stats_final_kernel(const float* __restrict__ partial, int nblocks, int c,
                                                             int64_t pixels, float eps, float momentum,
                                                             float* __restrict__ mean, float* __restrict__ istd,
                                                             float* __restrict__ moving_mean, float* __restrict__ moving_var) {
  __shared__ double sm_s[256], sm_ss[256];
  const int ch = blockIdx.x * 32 + (threadIdx.x >> 3), lane8 = threadIdx.x & 7;
  double s, ss;
  bn_final_sums(partial, nblocks, c, ch, lane8, sm_s, sm_ss, s, ss);
  if (lane8 == 0 && ch < c) {
    const double mu = s / pixels;
    double var = ss / pixels - mu * mu;
    if (var < 0.0) var = 0.0;
    mean[ch] = static_cast<float>(mu);
    istd[ch] = static_cast<float>(1.0 / sqrt(var + eps));
    if (moving_mean != nullptr) {
      const double unb = pixels > 1 ? var * pixels / (pixels - 1) : var;
      moving_mean[ch] = moving_mean[ch] * momentum + static_cast<float>(mu) * (1.f - momentum);
      moving_var[ch] = moving_var[ch] * momentum + static_cast<float>(unb) * (1.f - momentum);
    }
  }
}
// backward final: dbeta = sum d', dgamma = sum d' xhat  -> out2[0..c) = dgamma, out2[c..2c) = dbeta (accumulated into
// the gradient buffer if accumulate), plus the per-channel means the apply kernel needs
__global__ void __launch_bounds__(256) bn_bwd_final_kernel(const float* __restrict__ partial, int nblocks, int c,
                                                           float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                           int accumulate, float* __restrict__ sums) {
  __shared__ double sm_s[256], sm_ss[256];
  const int ch = blockIdx.x * 32 + (threadIdx.x >> 3), lane8 = threadIdx.x & 7;
  double s, sx;
  bn_final_sums(partial, nblocks, c, ch, lane8, sm_s, sm_ss, s, sx);
  if (lane8 == 0 && ch < c) {
    sums[ch] = static_cast<float>(s);
    sums[c + ch] = static_cast<float>(sx);
    if (dgamma != nullptr) {
      dgamma[ch] = accumulate ? dgamma[ch] + static_cast<float>(sx) : static_cast<float>(sx);
      dbeta[ch] = accumulate ? dbeta[ch] + static_cast<float>(s) : static_cast<float>(s);
    }
  }
}
// ---- sync-BatchNorm (data-parallel): every rank reduces its own block partials, publishes the 2c double sums in its
// heap (region alternating with the barrier epoch), meets the other ranks, and adds all ranks' sums in rank order - the
// batch statistics of the GLOBAL batch, bit-identical on every rank (model_builder.py:291-292 on one device).
__device__ __forceinline__ double ld_peer_f64(const double* p) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void bn_dp_exchange(const CommDev& cm, int slot, size_t sums_off, int c, int ch, bool owner,
                                               double& s, double& ss) {
  const uint32_t e = cm.counters[slot] + 1;
  const size_t off = sums_off + static_cast<size_t>(e & 1) * 2 * c * sizeof(double);
  if (owner) {
    double* mine = reinterpret_cast<double*>(cm.heap[cm.rank] + off);
    mine[ch] = s;
    mine[c + ch] = ss;
  }
  __threadfence_system();
  comm_barrier(cm, slot);
  if (owner) {
    double a = 0.0, b = 0.0;
    for (int p = 0; p < cm.world; ++p) {
      const double* theirs = reinterpret_cast<const double*>(cm.heap[p] + off);
      a += ld_peer_f64(theirs + ch);
      b += ld_peer_f64(theirs + c + ch);
    }
    s = a;
    ss = b;
  }
}
struct BnStatsDpArgs {
  int slot0;
  size_t sums_off;
  const float* partial;
  int nblocks, c;
  int64_t pixels_global;
  float eps, momentum;
  float *mean, *istd, *moving_mean, *moving_var;
};
__device__ __forceinline__ void bn_stats_final_dp_body(const CommDev& cm, const BnStatsDpArgs& a) {
  __shared__ double sm_s[256], sm_ss[256];
  const int c = a.c;
  const int ch = blockIdx.x * 32 + (threadIdx.x >> 3), lane8 = threadIdx.x & 7;
  double s, ss;
  bn_final_sums(a.partial, a.nblocks, c, ch, lane8, sm_s, sm_ss, s, ss);
  const bool owner = lane8 == 0 && ch < c;
  bn_dp_exchange(cm, a.slot0 + blockIdx.x, a.sums_off, c, ch, owner, s, ss);
  if (owner) {
    const double mu = s / a.pixels_global;
    double var = ss / a.pixels_global - mu * mu;
    if (var < 0.0) var = 0.0;
    a.mean[ch] = static_cast<float>(mu);
    a.istd[ch] = static_cast<float>(1.0 / sqrt(var + a.eps));
    if (a.moving_mean != nullptr) {
      const double unb = a.pixels_global > 1 ? var * a.pixels_global / (a.pixels_global - 1) : var;
      a.moving_mean[ch] = a.moving_mean[ch] * a.momentum + static_cast<float>(mu) * (1.f - a.momentum);
      a.moving_var[ch] = a.moving_var[ch] * a.momentum + static_cast<float>(unb) * (1.f - a.momentum);
    }
  }
}
__global__ void __launch_bounds__(256) bn_stats_final_dp_kernel(const CommDev cm, const BnStatsDpArgs a) {
  bn_stats_final_dp_body(cm, a);
}
// backward: dgamma / dbeta keep THIS rank's sums (the gradient exchange averages them like every other gradient); the
// sums the dz formula needs are over the global batch
struct BnBwdDpArgs {
  int slot0;
  size_t sums_off;
  const float* partial;
  int nblocks, c;
  float *dgamma, *dbeta;
  int accumulate;
  float* sums;
};
__device__ __forceinline__ void bn_bwd_final_dp_body(const CommDev& cm, const BnBwdDpArgs& a) {
  __shared__ double sm_s[256], sm_ss[256];
  const int c = a.c;
  const int ch = blockIdx.x * 32 + (threadIdx.x >> 3), lane8 = threadIdx.x & 7;
  double s, sx;
  bn_final_sums(a.partial, a.nblocks, c, ch, lane8, sm_s, sm_ss, s, sx);
  const bool owner = lane8 == 0 && ch < c;
  if (owner && a.dgamma != nullptr) {
    a.dgamma[ch] = a.accumulate ? a.dgamma[ch] + static_cast<float>(sx) : static_cast<float>(sx);
    a.dbeta[ch] = a.accumulate ? a.dbeta[ch] + static_cast<float>(s) : static_cast<float>(s);
  }
  bn_dp_exchange(cm, a.slot0 + blockIdx.x, a.sums_off, c, ch, owner, s, sx);
  if (owner) {
    a.sums[ch] = static_cast<float>(s);
    a.sums[c + ch] = static_cast<float>(sx);
  }
}
__global__ void __launch_bounds__(256) bn_bwd_final_dp_kernel(const CommDev cm, const BnBwdDpArgs a) {
  bn_bwd_final_dp_body(cm, a);
}
// emulated ranks (ssr_comm_open_local): all ranks' kernels as one cooperative launch, blockIdx.y = rank
template <typename Args>
struct MultiParams {
  CommDev dev[kCommMaxLocal];
  Args arg[kCommMaxLocal];
};
__global__ void __launch_bounds__(256) bn_stats_final_dp_multi_kernel(const __grid_constant__ MultiParams<BnStatsDpArgs> p) {
  bn_stats_final_dp_body(p.dev[blockIdx.y], p.arg[blockIdx.y]);
}
__global__ void __launch_bounds__(256) bn_bwd_final_dp_multi_kernel(const __grid_constant__ MultiParams<BnBwdDpArgs> p) {
  bn_bwd_final_dp_body(p.dev[blockIdx.y], p.arg[blockIdx.y]);
}
// y = lrelu(gamma * (x - mean) * istd + beta)
__global__ void bn_lrelu_fwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ mean,
                                    const float* __restrict__ istd, const float* __restrict__ gamma,
                                    const float* __restrict__ beta, float alpha, __nv_bfloat16* __restrict__ y,
                                    int64_t pixels, int c) {
  const int64_t total = pixels * c;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int ch = static_cast<int>(i % c);
    const float v = gamma[ch] * (__bfloat162float(x[i]) - mean[ch]) * istd[ch] + beta[ch];
    y[i] = __float2bfloat16_rn(v > 0.f ? v : alpha * v);
  }
}
// the same, 8 consecutive channels per thread (16-byte loads / stores); c % 8 == 0
__global__ void bn_lrelu_fwd_v8_kernel(const uint4* __restrict__ x, const float* __restrict__ mean,
                                       const float* __restrict__ istd, const float* __restrict__ gamma,
                                       const float* __restrict__ beta, float alpha, uint4* __restrict__ y, int64_t total8,
                                       int c8) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total8;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int ch = static_cast<int>(i % c8) * 8;
    const uint4 q = __ldg(x + i);
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int c0 = ch + 2 * k;
      float v0 = gamma[c0] * (__uint_as_float(w[k] << 16) - mean[c0]) * istd[c0] + beta[c0];
      float v1 = gamma[c0 + 1] * (__uint_as_float(w[k] & 0xFFFF0000u) - mean[c0 + 1]) * istd[c0 + 1] + beta[c0 + 1];
      v0 = v0 > 0.f ? v0 : alpha * v0;
      v1 = v1 > 0.f ? v1 : alpha * v1;
      o[k] = pack_bf16x2(v0, v1);
    }
    y[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}
__global__ void bn_lrelu_bwd_apply_v8_kernel(const uint4* __restrict__ x, const uint4* __restrict__ dy,
                                             const uint4* __restrict__ y, const float* __restrict__ mean,
                                             const float* __restrict__ istd, const float* __restrict__ gamma,
                                             const float* __restrict__ sums, float alpha, uint4* __restrict__ dz,
                                             int64_t total8, int c, float inv_m) {
  const int c8 = c >> 3;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total8;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int ch = static_cast<int>(i % c8) * 8;
    const uint4 qx = __ldg(x + i), qd = __ldg(dy + i), qy = __ldg(y + i);
    const uint32_t wx[4] = {qx.x, qx.y, qx.z, qx.w}, wd[4] = {qd.x, qd.y, qd.z, qd.w}, wy[4] = {qy.x, qy.y, qy.z, qy.w};
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float r[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int cc = ch + 2 * k + e;
        const float xv = e ? __uint_as_float(wx[k] & 0xFFFF0000u) : __uint_as_float(wx[k] << 16);
        const float dv = e ? __uint_as_float(wd[k] & 0xFFFF0000u) : __uint_as_float(wd[k] << 16);
        const float yv = e ? __uint_as_float(wy[k] & 0xFFFF0000u) : __uint_as_float(wy[k] << 16);
        const float d = dv * (yv > 0.f ? 1.f : alpha);
        const float xh = (xv - mean[cc]) * istd[cc];
        r[e] = gamma[cc] * istd[cc] * (d - sums[cc] * inv_m - xh * sums[c + cc] * inv_m);
      }
      o[k] = pack_bf16x2(r[0], r[1]);
    }
    dz[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}
// dz = gamma * istd * (d' - sum_d'/m - xhat * sum_d'xhat/m)
__global__ void bn_lrelu_bwd_apply_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy,
                                          const __nv_bfloat16* __restrict__ y, const float* __restrict__ mean,
                                          const float* __restrict__ istd, const float* __restrict__ gamma,
                                          const float* __restrict__ sums, float alpha, __nv_bfloat16* __restrict__ dz,
                                          int64_t pixels, int c, int64_t pixels_norm) {
  const int64_t total = pixels * c;
  const float inv_m = 1.f / static_cast<float>(pixels_norm);  // sync-BN: the statistics are over all ranks' pixels
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int ch = static_cast<int>(i % c);
    const float d = __bfloat162float(dy[i]) * (__bfloat162float(y[i]) > 0.f ? 1.f : alpha);
    const float xh = (__bfloat162float(x[i]) - mean[ch]) * istd[ch];
    dz[i] = __float2bfloat16_rn(gamma[ch] * istd[ch] * (d - sums[ch] * inv_m - xh * sums[c + ch] * inv_m));
  }
}

// ---------------------------------------------------------------- Dense (batch <= 32), fp32
static bool g_dense_no_tc = getenv("SSR_DENSE_NO_TC") != nullptr;   // development: CUDA-core kernels everywhere
constexpr int kDenseMaxN = 32;
constexpr int kDenseKSplit = 128;
// partial[s][n][o] = sum_{k in split s} x[n][k] W[k][o]; a thread owns FOUR output columns (one 16-byte weight load per
// k): the batch values are shared-memory broadcasts, and with one column per thread those broadcasts - 16 per 4 bytes
// of weights - capped the kernel at ~2 TB/s; four columns per broadcast put the 134 MB weight stream back on HBM.
template <int NMAX>
__global__ void __launch_bounds__(128) dense_fwd_partial_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                                int n, int K, int O, float* __restrict__ partial) {
  const int o = (blockIdx.x * 128 + threadIdx.x) * 4;
  const int s = blockIdx.y;
  const int kchunk = (K + kDenseKSplit - 1) / kDenseKSplit;
  const int k0 = s * kchunk, k1 = min(K, k0 + kchunk);
  __shared__ float xs[NMAX][64];
  float acc[NMAX][4];
#pragma unroll
  for (int i = 0; i < NMAX; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
  const bool vec = (O & 3) == 0 && o + 3 < O;
  for (int kb = k0; kb < k1; kb += 64) {
    const int kn = min(64, k1 - kb);
    __syncthreads();
    for (int i = threadIdx.x; i < n * 64; i += 128) {
      const int r = i / 64, cidx = i % 64;
      xs[r][cidx] = (cidx < kn) ? x[static_cast<int64_t>(r) * K + kb + cidx] : 0.f;
    }
    __syncthreads();
    if (o < O) {
      for (int kk0 = 0; kk0 < kn; kk0 += 4) {
        float4 wv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          wv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (kk0 + u < kn) {
            const float* wr = w + static_cast<int64_t>(kb + kk0 + u) * O + o;
            if (vec) {
              wv[u] = __ldg(reinterpret_cast<const float4*>(wr));
            } else {
              wv[u].x = __ldg(wr);
              if (o + 1 < O) wv[u].y = __ldg(wr + 1);
              if (o + 2 < O) wv[u].z = __ldg(wr + 2);
              if (o + 3 < O) wv[u].w = __ldg(wr + 3);
            }
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
          for (int i = 0; i < NMAX; ++i) {
            if (i < n) {
              const float xv = xs[i][kk0 + u];
              acc[i][0] += xv * wv[u].x;
              acc[i][1] += xv * wv[u].y;
              acc[i][2] += xv * wv[u].z;
              acc[i][3] += xv * wv[u].w;
            }
          }
        }
      }
    }
  }
  if (o < O) {
#pragma unroll
    for (int i = 0; i < NMAX; ++i) {
      if (i < n) {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (o + e < O) partial[(static_cast<int64_t>(s) * n + i) * O + o + e] = acc[i][e];
      }
    }
  }
}
// y[n][o] = act(b[o] + sum_s partial[s][n][o]); optionally also the pre-activation h
__global__ void dense_fwd_final_kernel(const float* __restrict__ partial, const float* __restrict__ b, int n, int O,
                                       int lrelu, float alpha, float* __restrict__ h, float* __restrict__ y) {
  const int64_t total = static_cast<int64_t>(n) * O;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float t = b[i % O];
    for (int s = 0; s < kDenseKSplit; ++s) t += partial[static_cast<int64_t>(s) * total + i];
    if (h != nullptr) h[i] = t;
    y[i] = (lrelu && t < 0.f) ? alpha * t : t;
  }
}
// dW[k][o] (+)= sum_n x[n][k] dy[n][o]: block = 128 output columns x a strip of 64 k; a thread keeps its dy column
// in registers, the x tile sits in shared memory as [k][n] (broadcast float4 reads), stores are coalesced over o.
constexpr int kDenseWgK = 64;
__global__ void __launch_bounds__(128) dense_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy, int n,
                                                          int K, int O, float* __restrict__ dw, int accumulate) {
  __shared__ __align__(16) float xs[kDenseWgK][kDenseMaxN];
  const int o = blockIdx.x * 128 + threadIdx.x;
  const int64_t k0 = static_cast<int64_t>(blockIdx.y) * kDenseWgK;
  for (int i = threadIdx.x; i < kDenseWgK * kDenseMaxN; i += 128) {
    const int kk = i % kDenseWgK, r = i / kDenseWgK;  // consecutive threads read consecutive k of one batch row
    xs[kk][r] = (r < n && k0 + kk < K) ? x[static_cast<int64_t>(r) * K + k0 + kk] : 0.f;
  }
  float d[kDenseMaxN];
#pragma unroll
  for (int i = 0; i < kDenseMaxN; ++i) d[i] = (i < n && o < O) ? dy[static_cast<int64_t>(i) * O + o] : 0.f;
  __syncthreads();
  if (o >= O) return;
  const int nq = (n + 3) / 4;
  for (int kk = 0; kk < kDenseWgK && k0 + kk < K; ++kk) {
    float t = 0.f;
#pragma unroll
    for (int q = 0; q < kDenseMaxN / 4; ++q) {
      if (q < nq) {
        const float4 xv = *reinterpret_cast<const float4*>(&xs[kk][4 * q]);
        t += xv.x * d[4 * q] + xv.y * d[4 * q + 1] + xv.z * d[4 * q + 2] + xv.w * d[4 * q + 3];
      }
    }
    float* out = dw + (k0 + kk) * O + o;
    *out = accumulate ? *out + t : t;
  }
}
// dx[n][k] = sum_o dy[n][o] W[k][o]: dy (n x O fp32) sits in shared memory, one warp handles FOUR k rows at a time:
// every dy value read from shared memory meets four weight rows (one k row per read kept the kernel at ~1 TB/s on
// shared-memory bandwidth, 16 LDS.128 per 16 bytes of weights), fixed-order shuffle tree at the end.
constexpr int kDenseDgRows = 4;
template <int NMAX>
__global__ void dense_dgrad_kernel(const float* __restrict__ dy, const float* __restrict__ w, int n, int K, int O,
                                   float* __restrict__ dx) {
  extern __shared__ __align__(16) float dys[];  // [n][O]
  for (int i = threadIdx.x; i < n * O; i += blockDim.x) dys[i] = dy[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const int64_t groups = (static_cast<int64_t>(K) + kDenseDgRows - 1) / kDenseDgRows;
  for (int64_t gq = blockIdx.x * static_cast<int64_t>(warps_per_block) + (threadIdx.x >> 5); gq < groups;
       gq += static_cast<int64_t>(gridDim.x) * warps_per_block) {
    const int64_t k = gq * kDenseDgRows;
    float acc[NMAX][kDenseDgRows];
#pragma unroll
    for (int i = 0; i < NMAX; ++i)
#pragma unroll
      for (int r = 0; r < kDenseDgRows; ++r) acc[i][r] = 0.f;
    if ((O & 3) == 0) {
      for (int o = 4 * lane; o < O; o += 128) {
        float4 wv[kDenseDgRows];
#pragma unroll
        for (int r = 0; r < kDenseDgRows; ++r)
          wv[r] = (k + r < K) ? __ldg(reinterpret_cast<const float4*>(w + (k + r) * O + o)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int i = 0; i < NMAX; ++i) {
          if (i < n) {
            const float4 dv = *reinterpret_cast<const float4*>(dys + i * O + o);
#pragma unroll
            for (int r = 0; r < kDenseDgRows; ++r)
              acc[i][r] += dv.x * wv[r].x + dv.y * wv[r].y + dv.z * wv[r].z + dv.w * wv[r].w;
          }
        }
      }
    } else {
      for (int o = lane; o < O; o += 32) {
        float wv[kDenseDgRows];
#pragma unroll
        for (int r = 0; r < kDenseDgRows; ++r) wv[r] = (k + r < K) ? __ldg(w + (k + r) * O + o) : 0.f;
#pragma unroll
        for (int i = 0; i < NMAX; ++i)
          if (i < n) {
            const float dv = dys[i * O + o];
#pragma unroll
            for (int r = 0; r < kDenseDgRows; ++r) acc[i][r] += dv * wv[r];
          }
      }
    }
#pragma unroll
    for (int i = 0; i < NMAX; ++i) {
      if (i < n) {
#pragma unroll
        for (int r = 0; r < kDenseDgRows; ++r) {
          float v = acc[i][r];
#pragma unroll
          for (int sft = 16; sft > 0; sft >>= 1) v += __shfl_xor_sync(0xffffffffu, v, sft);
          if (lane == 0 && k + r < K) dx[static_cast<int64_t>(i) * K + k + r] = v;
        }
      }
    }
  }
}
// ---------------------------------------------------------------- Dense on tensor cores (batch <= 16, wide layers)
// The CUDA-core kernels above are FMA-bound at batch 16 (16 FMAs per weight = ~54 TFLOP/s of fp32 at the HBM rate), not
// bandwidth-bound.  These three run the same sums as warp-level mma.sync m16n8k8 TF32 with the 3xTF32 split
// (v = hi + lo, a*b ~ a_hi*b_lo + a_lo*b_hi + a_hi*b_hi: ~21 significant bits, fp32 accumulation), the batch padded to the
// 16 rows of the A (forward, dgrad) operand or used as K (wgrad).  The weight matrix streams from HBM exactly once,
// straight into the B fragments (no shared-memory staging: every lane's loads are whole 32-byte sectors).  A 128 x N x 16
// tcgen05 tile would waste 7/8 of its rows here and needs the operands in shared memory; the work is 1 GFLOP per pass.
// Fragment maps: tools/dense_tc_fragment_check.py restates them on the CPU.
__device__ __forceinline__ void split_tf32(float v, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(v));
  const float r = v - __uint_as_float(hi);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(r));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void mma_3xtf32(float (&c)[4], const float (&a)[4], const float (&b)[2]) {
  uint32_t ah[4], al[4], bh[2], bl[2];
#pragma unroll
  for (int i = 0; i < 4; ++i) split_tf32(a[i], ah[i], al[i]);
#pragma unroll
  for (int i = 0; i < 2; ++i) split_tf32(b[i], bh[i], bl[i]);
  mma_tf32(c, ah, bl);
  mma_tf32(c, al, bh);
  mma_tf32(c, ah, bh);
}

// partial[s][n][o] = sum_{k in split s} x[n][k] W[k][o].  Block = 4 warps x 32 output columns, K range of split s in
// chunks of 64 (x chunk in shared memory, rows >= n are zero).  Warp: 4 n-tiles of 8 columns.
constexpr int kDenseTcKc = 64;
__global__ void __launch_bounds__(128) dense_fwd_tc_kernel(const float* __restrict__ x, const float* __restrict__ w, int n,
                                                           int K, int O, float* __restrict__ partial) {
  __shared__ float xs[16][kDenseTcKc + 4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int o0 = blockIdx.x * 128 + warp * 32;
  const int s = blockIdx.y;
  const int kchunk = ((K + kDenseKSplit - 1) / kDenseKSplit + 7) & ~7;
  const int k0 = min(K, s * kchunk), k1 = min(K, k0 + kchunk);
  float acc[4][4];
#pragma unroll
  for (int j = 0; j < 4; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
  for (int kb = k0; kb < k1; kb += kDenseTcKc) {
    __syncthreads();
    for (int i = threadIdx.x; i < 16 * kDenseTcKc; i += 128) {
      const int r = i / kDenseTcKc, cidx = i % kDenseTcKc;
      xs[r][cidx] = (r < n && kb + cidx < k1) ? x[static_cast<int64_t>(r) * K + kb + cidx] : 0.f;
    }
    __syncthreads();
    if (o0 >= O) continue;
    const int ksteps = (min(kDenseTcKc, k1 - kb) + 7) >> 3;
    for (int ks = 0; ks < ksteps; ++ks) {
      const int kk = 8 * ks;
      const float a[4] = {xs[g][kk + t], xs[g + 8][kk + t], xs[g][kk + t + 4], xs[g + 8][kk + t + 4]};
      const int64_t r0 = static_cast<int64_t>(kb + kk + t) * O, r1 = static_cast<int64_t>(kb + kk + t + 4) * O;
      const bool v0 = kb + kk + t < k1, v1 = kb + kk + t + 4 < k1;
      float b[4][2];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int o = o0 + 8 * j + g;
        b[j][0] = (v0 && o < O) ? __ldg(w + r0 + o) : 0.f;
        b[j][1] = (v1 && o < O) ? __ldg(w + r1 + o) : 0.f;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) mma_3xtf32(acc[j], a, b[j]);
    }
  }
  if (o0 >= O) return;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int o = o0 + 8 * j + 2 * t;
    float* p0 = partial + (static_cast<int64_t>(s) * n + g) * O + o;
    float* p1 = partial + (static_cast<int64_t>(s) * n + g + 8) * O + o;
    if (g < n) {
      if (o < O) p0[0] = acc[j][0];
      if (o + 1 < O) p0[1] = acc[j][1];
    }
    if (g + 8 < n) {
      if (o < O) p1[0] = acc[j][2];
      if (o + 1 < O) p1[1] = acc[j][3];
    }
  }
}

// dx[n][k] = sum_o dy[n][o] W[k][o]  (O % 32 == 0, K % 8 == 0).  dy (16 x O, rows >= n zero) sits in shared memory; a warp
// owns 8 weight rows at a time and walks O in groups of 32: lane (g, t) loads W[k0 + g][32 grp + 8 t .. + 8) as two
// 16-byte loads and uses o = 32 grp + 8 t + 2 s (slot t) and o + 1 (slot t + 4) in K-step s - the sum over o does not care
// about the order, and the A fragment reads dy with the same map.
template <int kWarps>
__global__ void __launch_bounds__(kWarps * 32) dense_dgrad_tc_kernel(const float* __restrict__ dy, const float* __restrict__ w,
                                                                     int n, int K, int O, float* __restrict__ dx) {
  extern __shared__ __align__(16) float dys[];  // [16][O + 8]
  const int ld = O + 8;
  for (int i = threadIdx.x; i < 16 * O; i += blockDim.x) {
    const int r = i / O, c = i % O;
    dys[r * ld + c] = (r < n) ? dy[static_cast<int64_t>(r) * O + c] : 0.f;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int tiles = K / 8;
  for (int tile = blockIdx.x * kWarps + warp; tile < tiles; tile += gridDim.x * kWarps) {
    const int kin0 = tile * 8;
    const float* wr = w + static_cast<int64_t>(kin0 + g) * O + 8 * t;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 2
    for (int grp = 0; grp < O / 32; ++grp) {
      const float4 w0 = __ldg(reinterpret_cast<const float4*>(wr + 32 * grp));
      const float4 w1 = __ldg(reinterpret_cast<const float4*>(wr + 32 * grp + 4));
      const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
      const float* d0 = dys + g * ld + 32 * grp + 8 * t;
      const float* d1 = d0 + 8 * ld;
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        const float2 e0 = *reinterpret_cast<const float2*>(d0 + 2 * s), e1 = *reinterpret_cast<const float2*>(d1 + 2 * s);
        const float a[4] = {e0.x, e1.x, e0.y, e1.y};
        const float b[2] = {wv[2 * s], wv[2 * s + 1]};
        mma_3xtf32(acc, a, b);
      }
    }
    const int k = kin0 + 2 * t;
    if (g < n) *reinterpret_cast<float2*>(dx + static_cast<int64_t>(g) * K + k) = make_float2(acc[0], acc[1]);
    if (g + 8 < n) *reinterpret_cast<float2*>(dx + static_cast<int64_t>(g + 8) * K + k) = make_float2(acc[2], acc[3]);
  }
}

// dW[k][o] (+)= sum_n x[n][k] dy[n][o]  (O % 8 == 0, K % 16 == 0).  The batch is the MMA's K (two steps of 8, rows >= n
// zero); a warp keeps the A fragments of its 16 weight rows in registers and walks all O / 8 column tiles, dy in shared
// memory; every lane stores (and, accumulating, first loads) 8-byte pairs: whole 32-byte sectors per row.
template <int kWarps>
__global__ void __launch_bounds__(kWarps * 32) dense_wgrad_tc_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                                     int n, int K, int O, float* __restrict__ dw,
                                                                     int accumulate) {
  extern __shared__ __align__(16) float dys[];  // [16][O + 8]
  const int ld = O + 8;
  for (int i = threadIdx.x; i < 16 * O; i += blockDim.x) {
    const int r = i / O, c = i % O;
    dys[r * ld + c] = (r < n) ? dy[static_cast<int64_t>(r) * O + c] : 0.f;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int tiles = K / 16;
  for (int tile = blockIdx.x * kWarps + warp; tile < tiles; tile += gridDim.x * kWarps) {
    const int kin0 = tile * 16;
    float a[2][4];
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      const int bt = 8 * ks + t;
      a[ks][0] = (bt < n) ? __ldg(x + static_cast<int64_t>(bt) * K + kin0 + g) : 0.f;
      a[ks][1] = (bt < n) ? __ldg(x + static_cast<int64_t>(bt) * K + kin0 + g + 8) : 0.f;
      a[ks][2] = (bt + 4 < n) ? __ldg(x + static_cast<int64_t>(bt + 4) * K + kin0 + g) : 0.f;
      a[ks][3] = (bt + 4 < n) ? __ldg(x + static_cast<int64_t>(bt + 4) * K + kin0 + g + 8) : 0.f;
    }
    float* out0 = dw + static_cast<int64_t>(kin0 + g) * O + 2 * t;
    float* out1 = out0 + static_cast<int64_t>(8) * O;
#pragma unroll 4
    for (int o0 = 0; o0 < O; o0 += 8) {
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      if (accumulate) {
        const float2 p0 = *reinterpret_cast<const float2*>(out0 + o0), p1 = *reinterpret_cast<const float2*>(out1 + o0);
        acc[0] = p0.x; acc[1] = p0.y; acc[2] = p1.x; acc[3] = p1.y;
      }
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        const float b[2] = {dys[(8 * ks + t) * ld + o0 + g], dys[(8 * ks + t + 4) * ld + o0 + g]};
        mma_3xtf32(acc, a[ks], b);
      }
      *reinterpret_cast<float2*>(out0 + o0) = make_float2(acc[0], acc[1]);
      *reinterpret_cast<float2*>(out1 + o0) = make_float2(acc[2], acc[3]);
    }
  }
}

// db[o] (+)= sum_n dy[n][o];  dh = dy * lrelu'(h)  (elementwise helpers)
__global__ void dense_bias_grad_kernel(const float* __restrict__ dy, int n, int O, float* __restrict__ db, int accumulate) {
  for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < O; o += gridDim.x * blockDim.x) {
    float t = 0.f;
    for (int r = 0; r < n; ++r) t += dy[static_cast<int64_t>(r) * O + o];
    db[o] = accumulate ? db[o] + t : t;
  }
}
__global__ void lrelu_bwd_f32_kernel(const float* __restrict__ dy, const float* __restrict__ h, float alpha,
                                     float* __restrict__ dh, int64_t count) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < count;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    dh[i] = dy[i] * (h[i] > 0.f ? 1.f : alpha);
}

// ---------------------------------------------------------------- relativistic-average GAN losses (one block, n <= 1024)
// out[0] = generator loss, out[1] = discriminator loss; g_dsr / d_dsr / d_dhr = d loss / d critic (fp32 [n_local]).
// Labels: scalars, or per-sample arrays (label smoothing, discriminator.py:240-254).  Data-parallel (use_comm): the
// relativistic means run over the GLOBAL batch (ra_adversarial_loss.py:59-67, ra_discriminator_loss.py:55-66), so every
// rank publishes its critics and labels in its heap, gathers all ranks' in rank order, evaluates the global losses (the
// same numbers on every rank) and keeps the gradients of its own samples times grad_scale (= world: the ranks' gradients
// are averaged afterwards).
constexpr int kRaganMax = 1024;
struct RaganArgs {
  int use_comm, slot;
  size_t stage_off;
  const float *hc, *sc;
  int n_local;
  float hr_label, sr_label;
  const float *hr_labels, *sr_labels;
  float grad_scale;
  float *out, *g_dsr, *d_dsr, *d_dhr;
  int standard;  // 1: the non-relativistic critic (sigmoid + Keras BinaryCrossentropy on probabilities)
};
__device__ __forceinline__ void ragan_body(const CommDev& cm, const RaganArgs& ra) {
  const int use_comm = ra.use_comm, slot = ra.slot, n_local = ra.n_local;
  const size_t stage_off = ra.stage_off;
  const float *hc = ra.hc, *sc = ra.sc, *hr_labels = ra.hr_labels, *sr_labels = ra.sr_labels;
  const float hr_label = ra.hr_label, sr_label = ra.sr_label, grad_scale = ra.grad_scale;
  float *out = ra.out, *g_dsr = ra.g_dsr, *d_dsr = ra.d_dsr, *d_dhr = ra.d_dhr;
  __shared__ float s_hc[kRaganMax], s_sc[kRaganMax], s_lh[kRaganMax], s_ls[kRaganMax];
  const int world = use_comm ? cm.world : 1, rank = use_comm ? cm.rank : 0;
  const int n = n_local * world;
  if (use_comm) {
    const uint32_t e = cm.counters[slot] + 1;
    const size_t off = stage_off + static_cast<size_t>(e & 1) * 4 * n_local * sizeof(float);
    float* mine = reinterpret_cast<float*>(cm.heap[cm.rank] + off);
    for (int i = threadIdx.x; i < n_local; i += blockDim.x) {
      mine[i] = hc[i];
      mine[n_local + i] = sc[i];
      mine[2 * n_local + i] = hr_labels ? hr_labels[i] : hr_label;
      mine[3 * n_local + i] = sr_labels ? sr_labels[i] : sr_label;
    }
    __threadfence_system();
    comm_barrier(cm, slot);
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const float* theirs = reinterpret_cast<const float*>(cm.heap[i / n_local] + off);
      const int j = i % n_local;
      s_hc[i] = ld_peer_f(theirs + j);
      s_sc[i] = ld_peer_f(theirs + n_local + j);
      s_lh[i] = ld_peer_f(theirs + 2 * n_local + j);
      s_ls[i] = ld_peer_f(theirs + 3 * n_local + j);
    }
  } else {
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      s_hc[i] = hc[i];
      s_sc[i] = sc[i];
      s_lh[i] = hr_labels ? hr_labels[i] : hr_label;
      s_ls[i] = sr_labels ? sr_labels[i] : sr_label;
    }
  }
  __syncthreads();
  if (threadIdx.x != 0) return;  // n is the batch size (16): a serial, fixed-order evaluation in double
  double mh = 0.0, ms = 0.0;
  for (int i = 0; i < n; ++i) {
    mh += s_hc[i];
    ms += s_sc[i];
  }
  mh /= n;
  ms /= n;
  auto softplus = [](double z) { return fmax(z, 0.0) + log1p(exp(-fabs(z))); };
  auto sigmoid = [](double z) { return 1.0 / (1.0 + exp(-z)); };
  if (ra.standard) {
    // Standard GAN critic (model_builder.py:194-196: Dense(1, sigmoid)): the critics arrive as logits z, p = sigmoid(z).
    // AdversarialLoss (adversarial_loss.py:58): BCE(1, p_sr); DiscriminatorLoss (discriminator_loss.py:56-59):
    // BCE(sr_labels, p_sr) + BCE(hr_labels, p_hr), Keras BinaryCrossentropy on probabilities: p clipped to
    // [eps, 1 - eps], eps = 1e-7, -(y log(p + eps) + (1 - y) log(1 - p + eps)), mean over the (global) batch.
    const double eps = 1e-7;
    auto bce = [&](double y, double z, double& dz) {
      const double pr = sigmoid(z);
      const double pc = fmin(fmax(pr, eps), 1.0 - eps);
      const double inside = (pr > eps && pr < 1.0 - eps) ? pr * (1.0 - pr) : 0.0;  // d clip(sigmoid(z)) / dz
      dz = (-y / (pc + eps) + (1.0 - y) / (1.0 - pc + eps)) * inside;
      return -(y * log(pc + eps) + (1.0 - y) * log(1.0 - pc + eps));
    };
    double gl = 0.0, dl = 0.0, t;
    for (int i = 0; i < n; ++i) {
      gl += bce(1.0, s_sc[i], t);
      dl += bce(s_ls[i], s_sc[i], t) + bce(s_lh[i], s_hc[i], t);
    }
    out[0] = static_cast<float>(gl / n);
    out[1] = static_cast<float>(dl / n);
    for (int j = 0; j < n_local; ++j) {
      const int i = rank * n_local + j;
      double dg, ds, dh;
      bce(1.0, s_sc[i], dg);
      bce(s_ls[i], s_sc[i], ds);
      bce(s_lh[i], s_hc[i], dh);
      g_dsr[j] = static_cast<float>(grad_scale * dg / n);
      d_dsr[j] = static_cast<float>(grad_scale * ds / n);
      d_dhr[j] = static_cast<float>(grad_scale * dh / n);
    }
    return;
  }
  double gl = 0.0, dl = 0.0, sga_g = 0.0, sgb_g = 0.0, sga_d = 0.0, sgb_d = 0.0;
  for (int i = 0; i < n; ++i) {
    const double a = s_hc[i] - ms, b = s_sc[i] - mh;
    gl += softplus(a) - 0.0 * a + softplus(b) - 1.0 * b;
    dl += softplus(a) - s_lh[i] * a + softplus(b) - s_ls[i] * b;
    sga_g += (sigmoid(a) - 0.0) / n;
    sgb_g += (sigmoid(b) - 1.0) / n;
    sga_d += (sigmoid(a) - s_lh[i]) / n;
    sgb_d += (sigmoid(b) - s_ls[i]) / n;
  }
  out[0] = static_cast<float>(0.5 * gl / n);
  out[1] = static_cast<float>(0.5 * dl / n);
  (void)sgb_g;
  for (int j = 0; j < n_local; ++j) {
    const int i = rank * n_local + j;
    const double a = s_hc[i] - ms, b = s_sc[i] - mh;
    g_dsr[j] = static_cast<float>(grad_scale * 0.5 * ((sigmoid(b) - 1.0) / n - sga_g / n));
    d_dsr[j] = static_cast<float>(grad_scale * 0.5 * ((sigmoid(b) - s_ls[i]) / n - sga_d / n));
    d_dhr[j] = static_cast<float>(grad_scale * 0.5 * ((sigmoid(a) - s_lh[i]) / n - sgb_d / n));
  }
}
__global__ void __launch_bounds__(32) ragan_kernel(const CommDev cm, const RaganArgs a) { ragan_body(cm, a); }
__global__ void __launch_bounds__(32) ragan_multi_kernel(const __grid_constant__ MultiParams<RaganArgs> p) {
  ragan_body(p.dev[blockIdx.y], p.arg[blockIdx.y]);
}

template <typename Args, typename Kernel>
static cudaError_t launch_multi(Kernel kernel, const CommDev* dev, const unsigned char* blobs, size_t stride, int world,
                                dim3 grid, dim3 block, cudaStream_t st) {
  MultiParams<Args> p;
  memset(&p, 0, sizeof(p));
  for (int r = 0; r < world; ++r) {
    p.dev[r] = dev[r];
    memcpy(&p.arg[r], blobs + r * stride, sizeof(Args));
  }
  void* params[] = {&p};
  grid.y = world;
  return cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(kernel), grid, block, params, 0, st);
}
static cudaError_t launch_bn_stats_multi(const CommDev* dev, const unsigned char* blobs, size_t stride, int world,
                                         cudaStream_t st) {
  BnStatsDpArgs a0;
  memcpy(&a0, blobs, sizeof(a0));
  return launch_multi<BnStatsDpArgs>(bn_stats_final_dp_multi_kernel, dev, blobs, stride, world, dim3((a0.c + 31) / 32),
                                     dim3(256), st);
}
static cudaError_t launch_bn_bwd_multi(const CommDev* dev, const unsigned char* blobs, size_t stride, int world,
                                       cudaStream_t st) {
  BnBwdDpArgs a0;
  memcpy(&a0, blobs, sizeof(a0));
  return launch_multi<BnBwdDpArgs>(bn_bwd_final_dp_multi_kernel, dev, blobs, stride, world, dim3((a0.c + 31) / 32),
                                   dim3(256), st);
}
static cudaError_t launch_ragan_multi(const CommDev* dev, const unsigned char* blobs, size_t stride, int world,
                                      cudaStream_t st) {
  return launch_multi<RaganArgs>(ragan_multi_kernel, dev, blobs, stride, world, dim3(1), dim3(32), st);
}

}  // namespace ssr

using namespace ssr;

#define SSR_CHECK_LAUNCH(name)                                                          \
  do {                                                                                  \
    cudaError_t e__ = cudaGetLastError();                                               \
    if (e__ != cudaSuccess) return set_error(SSR_ERR_CUDA, name ": %s", cudaGetErrorString(e__)); \
  } while (0)

extern "C" int ssr_subsample2(const void* x, void* y, int n, int oh, int ow, int c, int elem_bytes, void* stream) {
  if (!x || !y || n < 0 || oh <= 0 || ow <= 0 || c <= 0 || (c * elem_bytes) % 16)
    return set_error(SSR_ERR_INVALID, "subsample2: c * elem_bytes must be a multiple of 16");
  const int cv = c * elem_bytes / 16;
  const int64_t total = static_cast<int64_t>(n) * oh * ow * cv;
  if (total == 0) return SSR_OK;
  subsample2_kernel<<<grid1(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(x), static_cast<uint4*>(y), n, oh, ow, cv);
  SSR_CHECK_LAUNCH("subsample2");
  return SSR_OK;
}

extern "C" int ssr_zero_insert2(const void* dy, void* dx, int n, int oh, int ow, int c, int elem_bytes, void* stream) {
  if (!dy || !dx || n < 0 || oh <= 0 || ow <= 0 || c <= 0 || (c * elem_bytes) % 16)
    return set_error(SSR_ERR_INVALID, "zero_insert2: c * elem_bytes must be a multiple of 16");
  const int cv = c * elem_bytes / 16;
  const int64_t total = static_cast<int64_t>(n) * 4 * oh * ow * cv;
  if (total == 0) return SSR_OK;
  zero_insert2_kernel<<<grid1(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(dy), static_cast<uint4*>(dx), n, oh, ow, cv);
  SSR_CHECK_LAUNCH("zero_insert2");
  return SSR_OK;
}

extern "C" size_t ssr_bn_workspace_bytes(int c) { return static_cast<size_t>(kBnBlocks) * 2 * (c > 0 ? c : 1) * sizeof(float); }

static int bn_block(int c) { return c <= 256 ? 256 : 512; }

static int bn_bwd_apply_launch(const void* x, const void* dy, const void* y, const float* mean, const float* istd,
                               const float* gamma, const float* sums, float alpha, void* dz, int64_t pixels, int c,
                               int64_t pixels_norm, cudaStream_t st) {
  const uintptr_t al = reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(y) |
                       reinterpret_cast<uintptr_t>(dz);
  if (c % 8 == 0 && (al & 15) == 0) {
    const int64_t total8 = pixels * (c / 8);
    bn_lrelu_bwd_apply_v8_kernel<<<grid1(total8, 256), 256, 0, st>>>(
        static_cast<const uint4*>(x), static_cast<const uint4*>(dy), static_cast<const uint4*>(y), mean, istd, gamma, sums,
        alpha, static_cast<uint4*>(dz), total8, c, 1.f / static_cast<float>(pixels_norm));
  } else {
    bn_lrelu_bwd_apply_kernel<<<grid1(pixels * c, 256), 256, 0, st>>>(
        static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(dy), static_cast<const __nv_bfloat16*>(y),
        mean, istd, gamma, sums, alpha, static_cast<__nv_bfloat16*>(dz), pixels, c, pixels_norm);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(SSR_ERR_CUDA, "bn_bwd_apply: %s", cudaGetErrorString(e));
  return SSR_OK;
}

extern "C" int ssr_bn_stats_bf16(const void* x, int64_t pixels, int c, float eps, float momentum, void* workspace,
                                 float* mean, float* istd, float* moving_mean, float* moving_var, void* stream) {
  if (!x || !workspace || !mean || !istd || pixels <= 0 || c <= 0 || c > 512)
    return set_error(SSR_ERR_INVALID, "bn_stats: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int block = bn_block(c);
  const int nb = static_cast<int>(pixels < kBnBlocks ? pixels : kBnBlocks);
  if (c % 8 == 0) {
    const int lanes = 256 / (c / 8) > 0 ? 256 / (c / 8) : 1;
    const int vb = (c / 8) * lanes;  // <= 256 threads for c <= 2048
    bn_reduce_partial_v8_kernel<false><<<nb, vb, 2 * static_cast<size_t>(lanes) * c * sizeof(float), st>>>(
      static_cast<const __nv_bfloat16*>(x), nullptr, nullptr, nullptr, nullptr, 0.f, pixels, c,
      static_cast<float*>(workspace));
  } else {
    bn_reduce_partial_kernel<false><<<nb, block, 2 * block * sizeof(float), st>>>(
      static_cast<const __nv_bfloat16*>(x), nullptr, nullptr, nullptr, nullptr, 0.f, pixels, c,
      static_cast<float*>(workspace));
  }
  SSR_CHECK_LAUNCH("bn_stats_partial");
  bn_stats_final_kernel<<<(c + 31) / 32, 256, 0, st>>>(static_cast<const float*>(workspace), nb, c, pixels, eps, momentum, mean, istd,
                                           moving_mean, moving_var);
  SSR_CHECK_LAUNCH("bn_stats_final");
  return SSR_OK;
}

extern "C" int ssr_bn_lrelu_fwd_bf16(const void* x, const float* mean, const float* istd, const float* gamma,
                                     const float* beta, float alpha, void* y, int64_t pixels, int c, void* stream) {
  if (!x || !y || !mean || !istd || !gamma || !beta || pixels < 0 || c <= 0)
    return set_error(SSR_ERR_INVALID, "bn_lrelu_fwd: bad argument");
  if (pixels == 0) return SSR_OK;
  if (c % 8 == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0) {
    const int64_t total8 = pixels * (c / 8);
    bn_lrelu_fwd_v8_kernel<<<grid1(total8, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const uint4*>(x), mean, istd, gamma, beta, alpha, static_cast<uint4*>(y), total8, c / 8);
    SSR_CHECK_LAUNCH("bn_lrelu_fwd_v8");
    return SSR_OK;
  }
  bn_lrelu_fwd_kernel<<<grid1(pixels * c, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), mean, istd, gamma, beta, alpha, static_cast<__nv_bfloat16*>(y), pixels, c);
  SSR_CHECK_LAUNCH("bn_lrelu_fwd");
  return SSR_OK;
}

extern "C" int ssr_bn_lrelu_bwd_bf16(const void* x, const void* dy, const void* y, const float* mean, const float* istd,
                                     const float* gamma, float alpha, int64_t pixels, int c, void* workspace,
                                     float* sums_2c, float* dgamma, float* dbeta, int accumulate, void* dz,
                                     void* stream) {
  if (!x || !dy || !y || !mean || !istd || !gamma || !workspace || !sums_2c || !dz || pixels <= 0 || c <= 0 || c > 512)
    return set_error(SSR_ERR_INVALID, "bn_lrelu_bwd: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int block = bn_block(c);
  const int nb = static_cast<int>(pixels < kBnBlocks ? pixels : kBnBlocks);
  if (c % 8 == 0) {
    const int lanes = 256 / (c / 8) > 0 ? 256 / (c / 8) : 1;
    const int vb = (c / 8) * lanes;
    bn_reduce_partial_v8_kernel<true><<<nb, vb, 2 * static_cast<size_t>(lanes) * c * sizeof(float), st>>>(
      static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(dy), static_cast<const __nv_bfloat16*>(y),
      mean, istd, alpha, pixels, c, static_cast<float*>(workspace));
  } else {
    bn_reduce_partial_kernel<true><<<nb, block, 2 * block * sizeof(float), st>>>(
      static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(dy), static_cast<const __nv_bfloat16*>(y),
      mean, istd, alpha, pixels, c, static_cast<float*>(workspace));
  }
  SSR_CHECK_LAUNCH("bn_bwd_partial");
  bn_bwd_final_kernel<<<(c + 31) / 32, 256, 0, st>>>(static_cast<const float*>(workspace), nb, c, dgamma, dbeta, accumulate, sums_2c);
  SSR_CHECK_LAUNCH("bn_bwd_final");
  return bn_bwd_apply_launch(x, dy, y, mean, istd, gamma, sums_2c, alpha, dz, pixels, c, pixels, st);
}

extern "C" size_t ssr_dense_workspace_bytes(int n, int out_features) {
  return static_cast<size_t>(kDenseKSplit) * (n > 0 ? n : 1) * (out_features > 0 ? out_features : 1) * sizeof(float);
}

extern "C" int ssr_dense_fwd_f32(const float* x, const float* w, const float* b, int n, int in_features, int out_features,
                                 int lrelu, float alpha, void* workspace, float* pre_act, float* y, void* stream) {
  if (!x || !w || !b || !workspace || !y || n <= 0 || n > kDenseMaxN || in_features <= 0 || out_features <= 0)
    return set_error(SSR_ERR_INVALID, "dense_fwd: bad argument (batch <= 32)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (n <= 16 && out_features >= 64 && !g_dense_no_tc)   // wide layer: tensor cores (3xTF32), see dense_fwd_tc_kernel
    dense_fwd_tc_kernel<<<dim3((out_features + 127) / 128, kDenseKSplit), 128, 0, st>>>(
        x, w, n, in_features, out_features, static_cast<float*>(workspace));
  else if (n <= 16)
    dense_fwd_partial_kernel<16><<<dim3((out_features + 511) / 512, kDenseKSplit), 128, 0, st>>>(
        x, w, n, in_features, out_features, static_cast<float*>(workspace));
  else
    dense_fwd_partial_kernel<kDenseMaxN><<<dim3((out_features + 511) / 512, kDenseKSplit), 128, 0, st>>>(
        x, w, n, in_features, out_features, static_cast<float*>(workspace));
  SSR_CHECK_LAUNCH("dense_fwd_partial");
  dense_fwd_final_kernel<<<grid1(static_cast<int64_t>(n) * out_features, 256), 256, 0, st>>>(
      static_cast<const float*>(workspace), b, n, out_features, lrelu, alpha, pre_act, y);
  SSR_CHECK_LAUNCH("dense_fwd_final");
  return SSR_OK;
}

extern "C" int ssr_dense_bwd_f32(const float* x, const float* w, const float* dy, int n, int in_features,
                                 int out_features, float* dx, float* dw, float* db, int accumulate, void* stream) {
  if (!x || !w || !dy || n <= 0 || n > kDenseMaxN || in_features <= 0 || out_features <= 0)
    return set_error(SSR_ERR_INVALID, "dense_bwd: bad argument (batch <= 32)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t tc_smem = static_cast<size_t>(16) * (out_features + 8) * sizeof(float);
  const bool tc = n <= 16 && out_features >= 64 && tc_smem <= 200 * 1024 && !g_dense_no_tc;
  if (dw != nullptr) {
    if (tc && out_features % 8 == 0 && in_features % 16 == 0 && (reinterpret_cast<uintptr_t>(dw) & 7) == 0) {
      if (int rc = opt_in_dynamic_smem(reinterpret_cast<const void*>(dense_wgrad_tc_kernel<8>), 200 * 1024, "dense_wgrad_tc_kernel")) return rc;
      const int tiles = in_features / 16;
      dense_wgrad_tc_kernel<8><<<std::min((tiles + 7) / 8, 148 * 3), 256, tc_smem, st>>>(x, dy, n, in_features, out_features, dw,
                                                                                       accumulate);
    } else {
      dense_wgrad_kernel<<<dim3((out_features + 127) / 128, (in_features + kDenseWgK - 1) / kDenseWgK), 128, 0, st>>>(
          x, dy, n, in_features, out_features, dw, accumulate);
    }
    SSR_CHECK_LAUNCH("dense_wgrad");
  }
  if (db != nullptr) {
    dense_bias_grad_kernel<<<grid1(out_features, 128), 128, 0, st>>>(dy, n, out_features, db, accumulate);
    SSR_CHECK_LAUNCH("dense_bias_grad");
  }
  if (dx != nullptr) {
    const size_t dsm = static_cast<size_t>(n) * out_features * sizeof(float);
    if (dsm > 200 * 1024)
      return set_error(SSR_ERR_UNSUPPORTED, "dense_bwd: batch * out_features * 4 must be <= 200 KB");
    const int gdg = grid1(static_cast<int64_t>((in_features + kDenseDgRows - 1) / kDenseDgRows) * 32, 256, 4);
    if (tc && out_features % 32 == 0 && in_features % 8 == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0 &&
        (reinterpret_cast<uintptr_t>(dx) & 7) == 0) {
      if (int rc = opt_in_dynamic_smem(reinterpret_cast<const void*>(dense_dgrad_tc_kernel<8>), 200 * 1024, "dense_dgrad_tc_kernel")) return rc;
      const int tiles = in_features / 8;
      dense_dgrad_tc_kernel<8><<<std::min((tiles + 7) / 8, 148 * 3), 256, tc_smem, st>>>(dy, w, n, in_features, out_features, dx);
    } else if (n <= 16) {
      if (int rc = opt_in_dynamic_smem(reinterpret_cast<const void*>(dense_dgrad_kernel<16>), 200 * 1024, "dense_dgrad_kernel")) return rc;
      dense_dgrad_kernel<16><<<gdg, 256, dsm, st>>>(dy, w, n, in_features, out_features, dx);
    } else {
      if (int rc = opt_in_dynamic_smem(reinterpret_cast<const void*>(dense_dgrad_kernel<kDenseMaxN>), 200 * 1024, "dense_dgrad_kernel")) return rc;
      dense_dgrad_kernel<kDenseMaxN><<<gdg, 256, dsm, st>>>(dy, w, n, in_features, out_features, dx);
    }
    SSR_CHECK_LAUNCH("dense_dgrad");
  }
  return SSR_OK;
}

extern "C" int ssr_lrelu_bwd_f32(const float* dy, const float* h, float alpha, float* dh, int64_t count, void* stream) {
  if (!dy || !h || !dh || count < 0) return set_error(SSR_ERR_INVALID, "lrelu_bwd_f32: bad argument");
  if (count == 0) return SSR_OK;
  lrelu_bwd_f32_kernel<<<grid1(count, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(dy, h, alpha, dh, count);
  SSR_CHECK_LAUNCH("lrelu_bwd_f32");
  return SSR_OK;
}

extern "C" int ssr_ragan_losses(const float* hr_critic, const float* sr_critic, int n, float hr_label, float sr_label,
                                float* out2, float* g_dsr, float* d_dsr, float* d_dhr, void* stream) {
  if (!hr_critic || !sr_critic || !out2 || !g_dsr || !d_dsr || !d_dhr || n <= 0 || n > 1024)
    return set_error(SSR_ERR_INVALID, "ragan_losses: bad argument");
  CommDev none;
  memset(&none, 0, sizeof(none));
  RaganArgs a{0, 0, 0, hr_critic, sr_critic, n, hr_label, sr_label, nullptr, nullptr, 1.f, out2, g_dsr, d_dsr, d_dhr, 0};
  ragan_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(none, a);
  SSR_CHECK_LAUNCH("ragan_losses");
  return SSR_OK;
}

static int gan_losses_impl(ssr_comm* comm, int slot, size_t stage_off, const float* hr_critic, const float* sr_critic,
                           int n_local, float hr_label, float sr_label, const float* hr_labels, const float* sr_labels,
                           float* out2, float* g_dsr, float* d_dsr, float* d_dhr, int standard, void* stream) {
  if (!hr_critic || !sr_critic || !out2 || !g_dsr || !d_dsr || !d_dhr || n_local <= 0)
    return set_error(SSR_ERR_INVALID, "ragan_losses_ex: bad argument");
  CommDev cm;
  memset(&cm, 0, sizeof(cm));
  int world = 1;
  if (comm != nullptr) {
    const CommDev* d = comm_dev(comm);
    if (!d) return set_error(SSR_ERR_INVALID, "ragan_losses_ex: comm peers not opened");
    cm = *d;
    world = cm.world;
    if (slot < 0 || slot >= kCommMaxSlots || stage_off < kCommDataOffset || stage_off % 16 ||
        stage_off + 8 * static_cast<size_t>(n_local) * sizeof(float) > comm_heap_bytes(comm))
      return set_error(SSR_ERR_INVALID, "ragan_losses_ex: staging of 8 * n_local floats must lie inside the heap");
  }
  if (n_local * world > kRaganMax) return set_error(SSR_ERR_INVALID, "ragan_losses_ex: global batch > %d", kRaganMax);
  RaganArgs a{comm != nullptr, slot, stage_off, hr_critic, sr_critic, n_local, hr_label, sr_label, hr_labels, sr_labels,
              static_cast<float>(world), out2, g_dsr, d_dsr, d_dhr, standard};
  if (comm_is_group(comm))
    return comm_group_collective(comm, static_cast<cudaStream_t>(stream), reinterpret_cast<const void*>(&launch_ragan_multi),
                                 &a, sizeof(a), launch_ragan_multi);
  ragan_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(cm, a);
  SSR_CHECK_LAUNCH("ragan_losses_ex");
  return SSR_OK;
}

extern "C" int ssr_ragan_losses_ex(ssr_comm* comm, int slot, size_t stage_off, const float* hr_critic,
                                   const float* sr_critic, int n_local, float hr_label, float sr_label,
                                   const float* hr_labels, const float* sr_labels, float* out2, float* g_dsr, float* d_dsr,
                                   float* d_dhr, void* stream) {
  return gan_losses_impl(comm, slot, stage_off, hr_critic, sr_critic, n_local, hr_label, sr_label, hr_labels, sr_labels, out2,
                         g_dsr, d_dsr, d_dhr, 0, stream);
}

extern "C" int ssr_gan_losses_ex(ssr_comm* comm, int slot, size_t stage_off, const float* hr_critic,
                                 const float* sr_critic, int n_local, float hr_label, float sr_label,
                                 const float* hr_labels, const float* sr_labels, float* out2, float* g_dsr, float* d_dsr,
                                 float* d_dhr, void* stream) {
  return gan_losses_impl(comm, slot, stage_off, hr_critic, sr_critic, n_local, hr_label, sr_label, hr_labels, sr_labels, out2,
                         g_dsr, d_dsr, d_dhr, 1, stream);
}

// ---------------------------------------------------------------- sync-BatchNorm entry points (data-parallel training)
static int bn_dp_check(ssr_comm* comm, int slot0, size_t sums_off, int c, const char* what) {
  if (!comm_dev(comm)) return set_error(SSR_ERR_INVALID, "%s: comm is NULL or its peers are not opened", what);
  if (slot0 < 0 || slot0 + (c + 31) / 32 > kCommMaxSlots || sums_off < kCommDataOffset || sums_off % 16 ||
      sums_off + 4 * static_cast<size_t>(c) * sizeof(double) > comm_heap_bytes(comm))
    return set_error(SSR_ERR_INVALID, "%s: needs (c+31)/32 slots and 32*c bytes of heap at sums_off", what);
  return SSR_OK;
}

static int bn_partial_launch(bool bwd, const void* x, const void* dy, const void* y, const float* mean, const float* istd,
                             float alpha, int64_t pixels, int c, void* workspace, cudaStream_t st, int* nb_out) {
  const int block = bn_block(c);
  const int nb = static_cast<int>(pixels < kBnBlocks ? pixels : kBnBlocks);
  *nb_out = nb;
  const __nv_bfloat16 *xb = static_cast<const __nv_bfloat16*>(x), *db = static_cast<const __nv_bfloat16*>(dy),
                      *yb = static_cast<const __nv_bfloat16*>(y);
  float* ws = static_cast<float*>(workspace);
  if (c % 8 == 0) {
    const int lanes = 256 / (c / 8) > 0 ? 256 / (c / 8) : 1;
    const int vb = (c / 8) * lanes;
    const size_t sm = 2 * static_cast<size_t>(lanes) * c * sizeof(float);
    if (bwd) bn_reduce_partial_v8_kernel<true><<<nb, vb, sm, st>>>(xb, db, yb, mean, istd, alpha, pixels, c, ws);
    else bn_reduce_partial_v8_kernel<false><<<nb, vb, sm, st>>>(xb, nullptr, nullptr, nullptr, nullptr, 0.f, pixels, c, ws);
  } else {
    if (bwd) bn_reduce_partial_kernel<true><<<nb, block, 2 * block * sizeof(float), st>>>(xb, db, yb, mean, istd, alpha, pixels, c, ws);
    else bn_reduce_partial_kernel<false><<<nb, block, 2 * block * sizeof(float), st>>>(xb, nullptr, nullptr, nullptr, nullptr, 0.f, pixels, c, ws);
  }
  SSR_CHECK_LAUNCH("bn_partial");
  return SSR_OK;
}

extern "C" int ssr_bn_stats_bf16_dp(ssr_comm* comm, int slot0, size_t sums_off, const void* x, int64_t pixels_local, int c,
                                    float eps, float momentum, void* workspace, float* mean, float* istd,
                                    float* moving_mean, float* moving_var, void* stream) {
  if (!x || !workspace || !mean || !istd || pixels_local <= 0 || c <= 0 || c > 512)
    return set_error(SSR_ERR_INVALID, "bn_stats_dp: bad argument");
  if (int rc = bn_dp_check(comm, slot0, sums_off, c, "bn_stats_dp")) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int nb = 0;
  if (int rc = bn_partial_launch(false, x, nullptr, nullptr, nullptr, nullptr, 0.f, pixels_local, c, workspace, st, &nb)) return rc;
  const CommDev* cm = comm_dev(comm);
  BnStatsDpArgs a{slot0, sums_off, static_cast<const float*>(workspace), nb, c, pixels_local * cm->world, eps, momentum,
                  mean, istd, moving_mean, moving_var};
  if (comm_is_group(comm))
    return comm_group_collective(comm, st, reinterpret_cast<const void*>(&launch_bn_stats_multi), &a, sizeof(a),
                                 launch_bn_stats_multi);
  bn_stats_final_dp_kernel<<<(c + 31) / 32, 256, 0, st>>>(*cm, a);
  SSR_CHECK_LAUNCH("bn_stats_final_dp");
  return SSR_OK;
}

extern "C" int ssr_bn_lrelu_bwd_bf16_dp(ssr_comm* comm, int slot0, size_t sums_off, const void* x, const void* dy,
                                        const void* y, const float* mean, const float* istd, const float* gamma,
                                        float alpha, int64_t pixels_local, int c, void* workspace, float* sums_2c,
                                        float* dgamma, float* dbeta, int accumulate, void* dz, void* stream) {
  if (!x || !dy || !y || !mean || !istd || !gamma || !workspace || !sums_2c || !dz || pixels_local <= 0 || c <= 0 || c > 512)
    return set_error(SSR_ERR_INVALID, "bn_lrelu_bwd_dp: bad argument");
  if (int rc = bn_dp_check(comm, slot0, sums_off, c, "bn_lrelu_bwd_dp")) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int nb = 0;
  if (int rc = bn_partial_launch(true, x, dy, y, mean, istd, alpha, pixels_local, c, workspace, st, &nb)) return rc;
  const CommDev* cm = comm_dev(comm);
  BnBwdDpArgs a{slot0, sums_off, static_cast<const float*>(workspace), nb, c, dgamma, dbeta, accumulate, sums_2c};
  if (comm_is_group(comm)) {
    if (int rc = comm_group_collective(comm, st, reinterpret_cast<const void*>(&launch_bn_bwd_multi), &a, sizeof(a),
                                       launch_bn_bwd_multi))
      return rc;
  } else {
    bn_bwd_final_dp_kernel<<<(c + 31) / 32, 256, 0, st>>>(*cm, a);
    SSR_CHECK_LAUNCH("bn_bwd_final_dp");
  }
  return bn_bwd_apply_launch(x, dy, y, mean, istd, gamma, sums_2c, alpha, dz, pixels_local, c, pixels_local * cm->world, st);
}
