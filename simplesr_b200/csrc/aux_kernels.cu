// aux_kernels.cu — HBM-bound helpers around the convolutions: dtype edges, residual axpby,
// depth_to_space (standalone, bit-exact), overlapping-tile segmentation and stitching.
// All are pure data movement / elementwise: coalesced, 16-byte vectorised where alignment allows,
// grid sized as a multiple of the SM count with grid-stride loops.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <vector>

#include "internal.h"
#include "ptx_sm100.cuh"

namespace ssr {

static inline int grid_for(size_t work_items, int block, int sms = 148, int waves = 8) {
  size_t g = (work_items + block - 1) / block;
  size_t cap = static_cast<size_t>(sms) * waves;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

// ---------------------------------------------------------------- fp32 [pixels,c] -> bf16 [pixels,cpad]
// one thread per 8 output channels (16-byte store)
__global__ void f32_to_bf16_pad_kernel(const float* __restrict__ x, uint4* __restrict__ y, int64_t pixels, int c,
                                       int cpad) {
  const int groups = cpad / 8;
  const int64_t total = pixels * groups;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t pix = i / groups;
    const int g = static_cast<int>(i % groups);
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int ch = g * 8 + k;
      v[k] = (ch < c) ? __ldg(x + pix * c + ch) : 0.f;
    }
    uint4 q;
    q.x = pack_bf16x2(v[0], v[1]);
    q.y = pack_bf16x2(v[2], v[3]);
    q.z = pack_bf16x2(v[4], v[5]);
    q.w = pack_bf16x2(v[6], v[7]);
    y[i] = q;
  }
}

// y[n,h,w, dx*c + ch] = x[n,h,w+dx-kw/2,ch]; one thread per 8 output channels
__global__ void im2col_x_kernel(const float* __restrict__ x, uint4* __restrict__ y, int64_t rows, int w, int c, int kw,
                                int cpad) {
  const int groups = cpad / 8;
  const int64_t total = rows * w * groups;
  const int half = kw >> 1;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int g = static_cast<int>(i % groups);
    const int64_t pix = i / groups;
    const int xw = static_cast<int>(pix % w);
    const int64_t row = pix / w;
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int ch = g * 8 + k;
      const int dx = ch / c, cc = ch - dx * c;
      const int sx = xw + dx - half;
      v[k] = (dx < kw && sx >= 0 && sx < w) ? __ldg(x + (row * w + sx) * c + cc) : 0.f;
    }
    uint4 q;
    q.x = pack_bf16x2(v[0], v[1]);
    q.y = pack_bf16x2(v[2], v[3]);
    q.z = pack_bf16x2(v[4], v[5]);
    q.w = pack_bf16x2(v[6], v[7]);
    y[i] = q;
  }
}

__global__ void bf16_to_f32_kernel(const __nv_bfloat16* __restrict__ x, int cs, int coff, float* __restrict__ y,
                                   int64_t pixels, int c) {
  const int64_t total = pixels * c;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t pix = i / c;
    const int ch = static_cast<int>(i % c);
    y[i] = __bfloat162float(x[pix * cs + coff + ch]);
  }
}

// ---------------------------------------------------------------- out = a + beta*b (bf16 slices, 8 ch / thread)
__global__ void axpby_bf16_kernel(const __nv_bfloat16* __restrict__ a, int acs, int aoff,
                                  const __nv_bfloat16* __restrict__ b, int bcs, int boff, float beta,
                                  __nv_bfloat16* __restrict__ out, int ocs, int ooff, int64_t pixels, int c) {
  // PDL: resident early, nothing touched before the previous kernel of the stream has completed; the next kernel (a
  // conv) may in turn load its weights while this one runs
  grid_dep_wait();
  grid_dep_launch();
  const int groups = c / 8;
  const int64_t total = pixels * groups;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t pix = i / groups;
    const int g = static_cast<int>(i % groups);
    const uint4 qa = *reinterpret_cast<const uint4*>(a + pix * acs + aoff + g * 8);
    const uint4 qb = *reinterpret_cast<const uint4*>(b + pix * bcs + boff + g * 8);
    const uint32_t wa[4] = {qa.x, qa.y, qa.z, qa.w};
    const uint32_t wb[4] = {qb.x, qb.y, qb.z, qb.w};
    uint32_t wo[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float lo = bf16_lo(wa[k]) + beta * bf16_lo(wb[k]);
      const float hi = bf16_hi(wa[k]) + beta * bf16_hi(wb[k]);
      wo[k] = pack_bf16x2(lo, hi);
    }
    *reinterpret_cast<uint4*>(out + pix * ocs + ooff + g * 8) = make_uint4(wo[0], wo[1], wo[2], wo[3]);
  }
}

// ---------------------------------------------------------------- depth_to_space(2), NHWC, DCR
// out[n, 2h+i, 2w+j, c] = in[n, h, w, (2i+j)*C + c].  VEC-byte vectors; one thread per output vector so that
// stores are perfectly coalesced and loads are contiguous runs of C*elem bytes.
template <typename V>
__global__ void d2s2_kernel(const V* __restrict__ x, V* __restrict__ y, int n, int h, int w, int cv /* vectors per C */) {
  const int64_t total = static_cast<int64_t>(n) * (2 * h) * (2 * w) * cv;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % cv);
    int64_t q = i / cv;
    const int ox = static_cast<int>(q % (2 * w));
    q /= (2 * w);
    const int oy = static_cast<int>(q % (2 * h));
    const int nn = static_cast<int>(q / (2 * h));
    const int ih = oy >> 1, si = oy & 1, iw = ox >> 1, sj = ox & 1;
    const int64_t src = ((static_cast<int64_t>(nn) * h + ih) * w + iw) * (4 * cv) + (2 * si + sj) * cv + c;
    y[i] = __ldg(x + src);
  }
}

// ---------------------------------------------------------------- fp32-class activations as bf16 pairs
// The "fp32" precision mode of the SRResNet path (BASELINE.json configs[0] is an fp32 configuration): an fp32 activation
// a is carried as two bf16 tensors hi = bf16(a), lo = bf16(a - hi), a weight likewise, and a convolution becomes three
// tcgen05 passes accumulated in fp32, a_hi*w_hi + a_lo*w_hi + a_hi*w_lo (the dropped lo*lo term is ~2^-18 relative).
// This kernel is the elementwise tail between two such convolutions: v = act(z) (+ residual), optionally through the
// depth_to_space(2) permutation (model_builder.py:279), stored as fp32 and / or as the (hi | lo) channel pair.
__global__ void act_split_kernel(const float* __restrict__ z, int64_t total, int oh, int ow, int c, int up, int act,
                                 float act_alpha, const float* __restrict__ alpha, const float* __restrict__ res,
                                 float* __restrict__ y32, __nv_bfloat16* __restrict__ hl, int hl_cs, int hi_off, int lo_off) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int ch = static_cast<int>(i % c);
    const int64_t pix = i / c;   // output pixel (n, y, x)
    int64_t src = i;
    if (up == 2) {
      const int x = static_cast<int>(pix % ow);
      const int64_t q = pix / ow;
      const int y = static_cast<int>(q % oh);
      const int64_t n = q / oh;
      const int64_t spix = (n * (oh >> 1) + (y >> 1)) * (ow >> 1) + (x >> 1);
      src = spix * (4 * c) + ((y & 1) * 2 + (x & 1)) * c + ch;   // TF NHWC depth_to_space (DCR)
    }
    float v = z[src];
    if (act == SSR_ACT_LRELU) v = v > 0.f ? v : act_alpha * v;
    else if (act == SSR_ACT_PRELU) v = v > 0.f ? v : alpha[ch] * v;
    else if (act == SSR_ACT_RELU) v = fmaxf(v, 0.f);
    else if (act == SSR_ACT_TANH) v = tanhf(v);
    if (res != nullptr) v += res[i];
    if (y32 != nullptr) y32[i] = v;
    if (hl != nullptr) {
      const __nv_bfloat16 hi = __float2bfloat16_rn(v);
      hl[pix * hl_cs + hi_off + ch] = hi;
      hl[pix * hl_cs + lo_off + ch] = __float2bfloat16_rn(v - __bfloat162float(hi));
    }
  }
}
// y = x - float(bf16(x)): the low part of an fp32 tensor (input image of the x-unrolled first convolution)
__global__ void bf16_residual_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t count) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < count;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    y[i] = x[i] - __bfloat162float(__float2bfloat16_rn(x[i]));
}

// ---------------------------------------------------------------- overlapping tiles
// tiles[t, ty, tx, ch] = img[r*ph + ty - ov, cidx*pw + tx - ov, ch] (0 outside the image), t = r*cols + cidx.
// `img` holds image rows [row0, row0 + nrows) only (a rank's band of a sharded tiled inference; the whole image: 0, h).
// Both kernels move whole pixel ROWS (contiguous runs of floats in the image and in the tile): blockIdx.x walks
// (tile, tile row) jobs, the threads of a block copy one run with 16-byte accesses when source and destination are
// aligned alike - no per-element index arithmetic.  Exact copies (bit-exact against image_utils.py:124-148, 167-184).
__device__ __forceinline__ void copy_run(const float* __restrict__ src, float* __restrict__ dst, int len) {
  if (((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0) {
    const int n4 = len >> 2;
    const float4* s4 = reinterpret_cast<const float4*>(src);
    float4* d4 = reinterpret_cast<float4*>(dst);
    for (int i = threadIdx.x; i < n4; i += blockDim.x) d4[i] = __ldg(s4 + i);
    for (int i = (n4 << 2) + threadIdx.x; i < len; i += blockDim.x) dst[i] = __ldg(src + i);
  } else {
    for (int i = threadIdx.x; i < len; i += blockDim.x) dst[i] = __ldg(src + i);
  }
}
__device__ __forceinline__ void zero_run(float* __restrict__ dst, int len) {
  for (int i = threadIdx.x; i < len; i += blockDim.x) dst[i] = 0.f;
}

__global__ void __launch_bounds__(256) segment_tiles_kernel(const float* __restrict__ img, int h, int w, int c, int ph, int pw,
                                                            int ov, int cols, int tile_begin, int tile_count, int row0,
                                                            int nrows, float* __restrict__ tiles) {
  const int tsy = ph + 2 * ov, tsx = pw + 2 * ov;
  const int jobs = tile_count * tsy;
  for (int job = blockIdx.x; job < jobs; job += gridDim.x) {
    const int tl = job / tsy, ty = job - tl * tsy;
    const int t = tl + tile_begin;
    const int y = (t / cols) * ph + ty - ov;
    const int xs = (t % cols) * pw - ov;                       // image column of tile column 0
    float* dst = tiles + (static_cast<int64_t>(tl) * tsy + ty) * tsx * c;
    if (y < 0 || y >= h || y < row0 || y >= row0 + nrows) {
      zero_run(dst, tsx * c);
      continue;
    }
    const int lo = max(0, -xs), hi = min(tsx, w - xs);         // tile columns [lo, hi) lie inside the image
    if (lo > 0) zero_run(dst, min(lo, tsx) * c);
    if (hi > lo) copy_run(img + (static_cast<int64_t>(y - row0) * w + xs + lo) * c, dst + static_cast<int64_t>(lo) * c, (hi - lo) * c);
    if (hi < tsx) zero_run(dst + static_cast<int64_t>(max(hi, 0)) * c, (tsx - max(hi, 0)) * c);
  }
}

// out[y, x, ch] = tiles[(y/psy)*cols + x/psx][ov*s + y%psy, ov*s + x%psx, ch],  psy = patch_h*scale, psx = patch_w*scale.
// `out` holds output rows [row0, row0 + nrows) only (a rank's band); rows outside it are skipped.
__global__ void __launch_bounds__(256) stitch_tiles_kernel(const float* __restrict__ tiles, int H, int W, int c, int psy, int psx,
                                                           int ovs, int cols, int tile_begin, int tile_count, int row0,
                                                           int nrows, float* __restrict__ out) {
  const int tsy = psy + 2 * ovs, tsx = psx + 2 * ovs;
  const int jobs = tile_count * psy;
  for (int job = blockIdx.x; job < jobs; job += gridDim.x) {
    const int tl = job / psy, py = job - tl * psy;
    const int t = tl + tile_begin;
    const int y = (t / cols) * psy + py;
    const int x0 = (t % cols) * psx;
    if (y >= H || y < row0 || y >= row0 + nrows || x0 >= W) continue;
    const int len = min(psx, W - x0) * c;
    copy_run(tiles + ((static_cast<int64_t>(tl) * tsy + (ovs + py)) * tsx + ovs) * c,
             out + (static_cast<int64_t>(y - row0) * W + x0) * c, len);
  }
}

// ---------------------------------------------------------------- tcgen05 issue-rate microbenchmark
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int m, int n, int a_swz, int iters, int a_shift_rows, long long* cycles_out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  // operands: A 128 rows x 128 B at base, B n rows x 128 B at base + 16 KB; contents irrelevant but finite
  for (uint32_t i = threadIdx.x; i < (1024 + 32768 + 32768) / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(smem_raw)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    fence_mbar_init();
  }
  fence_proxy_async();
  if (warp == 0) tmem_alloc(smem_u32(&tmem_slot), 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp == 0) {
    const uint32_t leader = elect_one() ? 1u : 0u;
    const uint32_t idesc = umma_idesc_bf16(m, n);
    // A rows are 128 / 64 / 32 bytes for swizzle 128B / 64B / 32B; SBO = 8 rows
    const uint32_t a_row = (a_swz == 2) ? 128u : (a_swz == 4 ? 64u : 32u);
    const uint32_t hi_a = umma_desc_hi(8 * a_row, a_swz), hi_b = umma_desc_hi(1024, 2);
    const uint32_t a_lo = umma_desc_lo(base + a_shift_rows * a_row) | (1u << 16);
    const uint32_t b_lo = umma_desc_lo(base + 32768) | (1u << 16);
    const uint32_t kmask = (a_row >> 4) - 1;
    // warm-up
    umma_bf16_pred(tmem, desc64(hi_a, a_lo), desc64(hi_b, b_lo), idesc, 0, leader);
    umma_commit_pred(smem_u32(&bar), leader);
    mbar_wait(smem_u32(&bar), 0);
    const long long t0 = clock64();
    for (int i = 0; i < iters; i += 4) {
#pragma unroll
      for (int k = 0; k < 4; ++k)  // walk the K-steps of the row
        umma_bf16_pred(tmem, desc64(hi_a, a_lo + ((2 * k) & kmask)), desc64(hi_b, b_lo + 2 * k), idesc, 1, leader);
    }
    const long long t_issue = clock64();
    umma_commit_pred(smem_u32(&bar), leader);
    mbar_wait(smem_u32(&bar), 1);
    const long long t1 = clock64();
    if (leader) {
      cycles_out[blockIdx.x] = t1 - t0;
      cycles_out[gridDim.x + blockIdx.x] = t_issue - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 256);
  }
}

// same for CTA pairs: M = 256 over two SMs (cta_group::2), issued by the leader CTA of each cluster
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
    mma_rate2_kernel(int n, int iters, long long* cycles_out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  const uint32_t rank = cluster_ctarank();
  for (uint32_t i = threadIdx.x; i < (1024 + 32768 + 32768) / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(smem_raw)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    fence_mbar_init();
  }
  fence_proxy_async();
  cluster_sync_all();
  if (warp == 0) tmem_alloc2(smem_u32(&tmem_slot), 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  cluster_sync_all();
  if (rank == 0 && warp == 0) {
    const uint32_t leader = elect_one() ? 1u : 0u;
    const uint32_t idesc = umma_idesc_bf16(256, n);
    const uint32_t hi = umma_desc_hi(1024, 2);
    const uint32_t a_lo = umma_desc_lo(base) | (1u << 16);
    const uint32_t b_lo = umma_desc_lo(base + 32768) | (1u << 16);
    umma2_bf16_pred(tmem, desc64(hi, a_lo), desc64(hi, b_lo), idesc, 0, leader);
    umma2_commit_mc_pred(smem_u32(&bar), 1, leader);
    mbar_wait(smem_u32(&bar), 0);
    const long long t0 = clock64();
    for (int i = 0; i < iters; i += 4) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma2_bf16_pred(tmem, desc64(hi, a_lo + 2 * k), desc64(hi, b_lo + 2 * k), idesc, 1, leader);
    }
    const long long t_issue = clock64();
    umma2_commit_mc_pred(smem_u32(&bar), 1, leader);
    mbar_wait(smem_u32(&bar), 1);
    const long long t1 = clock64();
    if (leader) {
      cycles_out[blockIdx.x >> 1] = t1 - t0;
      cycles_out[(gridDim.x >> 1) + (blockIdx.x >> 1)] = t_issue - t0;
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc2(tmem, 256);
  }
}

int diag_mma_rate2(ssr_ctx* ctx, int n, int iters, float* host_cycles_per_mma) {
  if (n < 32 || n > 256 || n % 32 != 0 || iters <= 0 || iters % 4 != 0) return set_error(SSR_ERR_INVALID, "diag_mma_rate2: bad n/iters");
  long long* d = nullptr;
  const int pairs = ctx->sm_count / 2;
  if (cudaMalloc(&d, sizeof(long long) * pairs * 2) != cudaSuccess) return set_error(SSR_ERR_NOMEM, "cudaMalloc");
  const int smem = 1024 + 32768 + 32768;
  cudaFuncSetAttribute(mma_rate2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  mma_rate2_kernel<<<2 * pairs, 128, smem>>>(n, iters, d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    cudaFree(d);
    return set_error(SSR_ERR_CUDA, "mma_rate2_kernel: %s", cudaGetErrorString(e));
  }
  std::vector<long long> h(pairs * 2);
  cudaMemcpy(h.data(), d, sizeof(long long) * pairs * 2, cudaMemcpyDeviceToHost);
  cudaFree(d);
  long long mx = 0, mi = 0;
  for (int i = 0; i < pairs; ++i) mx = std::max(mx, h[i]), mi = std::max(mi, h[pairs + i]);
  host_cycles_per_mma[0] = static_cast<float>(mx) / iters;
  host_cycles_per_mma[1] = static_cast<float>(mi) / iters;
  ctx->launches++;
  return SSR_OK;
}

int diag_mma_rate(ssr_ctx* ctx, int m, int n, int a_swz, int iters, int a_shift_rows, float* host_cycles_per_mma) {
  if (!(m == 64 || m == 128) || !(a_swz == 2 || a_swz == 4 || a_swz == 6) || n < 8 || n > 256 || n % (m == 64 ? 8 : 16) != 0 || iters % 4 != 0 || iters <= 0 || a_shift_rows < 0 || a_shift_rows > 128) return set_error(SSR_ERR_INVALID, "diag_mma_rate: bad n/iters");
  long long* d = nullptr;
  const int grid = ctx->sm_count;
  if (cudaMalloc(&d, sizeof(long long) * grid * 2) != cudaSuccess) return set_error(SSR_ERR_NOMEM, "cudaMalloc");
  const int smem = 1024 + 32768 + 32768;
  cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  mma_rate_kernel<<<grid, 128, smem>>>(m, n, a_swz, iters, a_shift_rows, d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    cudaFree(d);
    return set_error(SSR_ERR_CUDA, "mma_rate_kernel: %s", cudaGetErrorString(e));
  }
  std::vector<long long> h(grid * 2);
  cudaMemcpy(h.data(), d, sizeof(long long) * grid * 2, cudaMemcpyDeviceToHost);
  cudaFree(d);
  long long mx = 0, mi = 0;
  for (int i = 0; i < grid; ++i) mx = std::max(mx, h[i]), mi = std::max(mi, h[grid + i]);
  host_cycles_per_mma[0] = static_cast<float>(mx) / iters;  // until the last MMA completed
  host_cycles_per_mma[1] = static_cast<float>(mi) / iters;  // until the last MMA was issued
  ctx->launches++;
  return SSR_OK;
}

}  // namespace ssr

// ================================================================= C ABI (bandwidth kernels)
using namespace ssr;

#define SSR_CHECK_LAUNCH(name)                                                          \
  do {                                                                                  \
    cudaError_t e__ = cudaGetLastError();                                               \
    if (e__ != cudaSuccess) return set_error(SSR_ERR_CUDA, name ": %s", cudaGetErrorString(e__)); \
  } while (0)

extern "C" int ssr_f32_to_bf16_pad(const float* x, void* y, int64_t pixels, int c, int cpad, void* stream) {
  if (pixels < 0 || c <= 0 || cpad < c || cpad % 8 != 0) return set_error(SSR_ERR_INVALID, "f32_to_bf16_pad: bad shape");
  if (pixels == 0) return SSR_OK;
  const int block = 256;
  f32_to_bf16_pad_kernel<<<grid_for(pixels * (cpad / 8), block), block, 0, static_cast<cudaStream_t>(stream)>>>(
      x, static_cast<uint4*>(y), pixels, c, cpad);
  SSR_CHECK_LAUNCH("f32_to_bf16_pad");
  return SSR_OK;
}

extern "C" int ssr_im2col_x_f32_to_bf16(const float* x, void* y, int n, int h, int w, int c, int kw, int cpad,
                                        void* stream) {
  if (n < 0 || h < 0 || w < 0 || c <= 0 || kw < 1 || !(kw & 1) || cpad < kw * c || cpad % 16 != 0)
    return set_error(SSR_ERR_INVALID, "im2col_x: bad shape");
  const int64_t rows = static_cast<int64_t>(n) * h;
  if (rows * w == 0) return SSR_OK;
  const int block = 256;
  im2col_x_kernel<<<grid_for(rows * w * (cpad / 8), block), block, 0, static_cast<cudaStream_t>(stream)>>>(
      x, static_cast<uint4*>(y), rows, w, c, kw, cpad);
  SSR_CHECK_LAUNCH("im2col_x");
  return SSR_OK;
}

extern "C" int ssr_bf16_to_f32(const void* x, int x_cstride, int x_coff, float* y, int64_t pixels, int c,
                               void* stream) {
  if (pixels < 0 || c <= 0 || x_cstride < x_coff + c) return set_error(SSR_ERR_INVALID, "bf16_to_f32: bad shape");
  if (pixels == 0) return SSR_OK;
  const int block = 256;
  bf16_to_f32_kernel<<<grid_for(pixels * c, block), block, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), x_cstride, x_coff, y, pixels, c);
  SSR_CHECK_LAUNCH("bf16_to_f32");
  return SSR_OK;
}

extern "C" int ssr_axpby_bf16(const void* a, int a_cstride, int a_coff, const void* b, int b_cstride, int b_coff,
                              float beta, void* out, int out_cstride, int out_coff, int64_t pixels, int c,
                              void* stream) {
  if (pixels < 0 || c <= 0 || c % 8 || a_cstride % 8 || a_coff % 8 || b_cstride % 8 || b_coff % 8 || out_cstride % 8 ||
      out_coff % 8)
    return set_error(SSR_ERR_INVALID, "axpby_bf16: channel counts/offsets must be multiples of 8");
  if (pixels == 0) return SSR_OK;
  const int block = 256;
  cudaError_t le = launch_pdl(axpby_bf16_kernel, dim3(grid_for(pixels * (c / 8), block)), dim3(block), 0,
                              static_cast<cudaStream_t>(stream), static_cast<const __nv_bfloat16*>(a), a_cstride, a_coff,
                              static_cast<const __nv_bfloat16*>(b), b_cstride, b_coff, beta,
                              static_cast<__nv_bfloat16*>(out), out_cstride, out_coff, pixels, c);
  if (le != cudaSuccess) return set_error(SSR_ERR_CUDA, "axpby_bf16 launch: %s", cudaGetErrorString(le));
  return SSR_OK;
}

extern "C" int ssr_depth_to_space2(const void* x, void* y, int n, int h, int w, int c, int elem_bytes, void* stream) {
  if (n < 0 || h < 0 || w < 0 || c <= 0 || !(elem_bytes == 2 || elem_bytes == 4))
    return set_error(SSR_ERR_INVALID, "depth_to_space2: bad shape");
  const int64_t out_elems = static_cast<int64_t>(n) * h * w * 4 * c;
  if (out_elems == 0) return SSR_OK;
  const int row_bytes = c * elem_bytes;  // contiguous run per (pixel, sub-position)
  const int block = 256;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool al16 = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
  if (row_bytes % 16 == 0 && al16) {
    const int cv = row_bytes / 16;
    d2s2_kernel<uint4><<<grid_for(out_elems * elem_bytes / 16, block, 148, 16), block, 0, st>>>(
        static_cast<const uint4*>(x), static_cast<uint4*>(y), n, h, w, cv);
  } else if (row_bytes % 4 == 0) {
    const int cv = row_bytes / 4;
    d2s2_kernel<uint32_t><<<grid_for(out_elems * elem_bytes / 4, block, 148, 16), block, 0, st>>>(
        static_cast<const uint32_t*>(x), static_cast<uint32_t*>(y), n, h, w, cv);
  } else {
    d2s2_kernel<uint16_t><<<grid_for(out_elems, block, 148, 16), block, 0, st>>>(
        static_cast<const uint16_t*>(x), static_cast<uint16_t*>(y), n, h, w, c);
  }
  SSR_CHECK_LAUNCH("depth_to_space2");
  return SSR_OK;
}

extern "C" int ssr_segment_tiles_ex(const float* img, int h, int w, int c, int patch_h, int patch_w, int overlap,
                                    int tile_begin, int tile_count, int src_row0, int src_rows, float* tiles,
                                    void* stream) {
  if (!img || !tiles || h <= 0 || w <= 0 || c <= 0 || patch_h <= 0 || patch_w <= 0 || overlap < 0)
    return set_error(SSR_ERR_INVALID, "segment_tiles: bad shape");
  if (h < patch_h || w < patch_w)
    return set_error(SSR_ERR_INVALID, "Patch dimensions are larger than image size");  // image_utils.py:115-116
  if (src_row0 < 0 || src_rows <= 0 || src_row0 + src_rows > h)
    return set_error(SSR_ERR_INVALID, "segment_tiles: source band outside the image");
  const int cols = (w + patch_w - 1) / patch_w, rows = (h + patch_h - 1) / patch_h;
  if (tile_begin < 0 || tile_count < 0 || tile_begin + tile_count > rows * cols)
    return set_error(SSR_ERR_INVALID, "segment_tiles: tile range out of bounds");
  if (tile_count == 0) return SSR_OK;
  // the band must hold every image row the selected tiles touch
  const int r_first = tile_begin / cols, r_last = (tile_begin + tile_count - 1) / cols;
  const int need_lo = std::max(0, r_first * patch_h - overlap), need_hi = std::min(h, (r_last + 1) * patch_h + overlap);
  if (src_row0 > need_lo || src_row0 + src_rows < need_hi)
    return set_error(SSR_ERR_INVALID, "segment_tiles: source band [%d, %d) does not cover rows [%d, %d) of the tiles",
                     src_row0, src_row0 + src_rows, need_lo, need_hi);
  const int64_t jobs = static_cast<int64_t>(tile_count) * (patch_h + 2 * overlap);
  const int block = (patch_w + 2 * overlap) * c >= 1024 ? 256 : 128;
  segment_tiles_kernel<<<static_cast<int>(std::min<int64_t>(jobs, 148 * 16)), block, 0, static_cast<cudaStream_t>(stream)>>>(
      img, h, w, c, patch_h, patch_w, overlap, cols, tile_begin, tile_count, src_row0, src_rows, tiles);
  SSR_CHECK_LAUNCH("segment_tiles");
  return SSR_OK;
}

extern "C" int ssr_segment_tiles(const float* img, int h, int w, int c, int patch, int overlap, int tile_begin,
                                 int tile_count, float* tiles, void* stream) {
  return ssr_segment_tiles_ex(img, h, w, c, patch, patch, overlap, tile_begin, tile_count, 0, h, tiles, stream);
}

extern "C" int ssr_stitch_tiles_ex(const float* tiles, int h, int w, int c, int patch_h, int patch_w, int overlap,
                                   int scale, int tile_begin, int tile_count, int out_row0, int out_rows, float* out,
                                   void* stream) {
  if (!tiles || !out || h <= 0 || w <= 0 || c <= 0 || patch_h <= 0 || patch_w <= 0 || overlap < 0 || scale <= 0)
    return set_error(SSR_ERR_INVALID, "stitch_tiles: bad shape");
  if (out_row0 < 0 || out_rows <= 0 || out_row0 + out_rows > h * scale)
    return set_error(SSR_ERR_INVALID, "stitch_tiles: output band outside the image");
  const int cols = (w + patch_w - 1) / patch_w, rows = (h + patch_h - 1) / patch_h;
  if (tile_begin < 0 || tile_count < 0 || tile_begin + tile_count > rows * cols)
    return set_error(SSR_ERR_INVALID, "stitch_tiles: tile range out of bounds");
  if (tile_count == 0) return SSR_OK;
  const int psy = patch_h * scale, psx = patch_w * scale;
  const int64_t jobs = static_cast<int64_t>(tile_count) * psy;
  const int block = psx * c >= 1024 ? 256 : 128;
  stitch_tiles_kernel<<<static_cast<int>(std::min<int64_t>(jobs, 148 * 16)), block, 0, static_cast<cudaStream_t>(stream)>>>(
      tiles, h * scale, w * scale, c, psy, psx, overlap * scale, cols, tile_begin, tile_count, out_row0, out_rows, out);
  SSR_CHECK_LAUNCH("stitch_tiles");
  return SSR_OK;
}

extern "C" int ssr_stitch_tiles(const float* tiles, int h, int w, int c, int patch, int overlap, int scale,
                                int tile_begin, int tile_count, float* out, void* stream) {
  return ssr_stitch_tiles_ex(tiles, h, w, c, patch, patch, overlap, scale, tile_begin, tile_count, 0, h * scale, out, stream);
}

extern "C" int ssr_act_split_f32(const float* z, int n, int h, int w, int c_out, int up, int act, float act_alpha,
                                 const float* alpha, const float* res32, float* y32, void* hi_lo_bf16, int hl_cstride,
                                 int hi_off, int lo_off, void* stream) {
  if (!z || n <= 0 || h <= 0 || w <= 0 || c_out <= 0 || !(up == 1 || up == 2) || act < 0 || act > 4 ||
      (act == SSR_ACT_PRELU && !alpha) || (!y32 && !hi_lo_bf16))
    return set_error(SSR_ERR_INVALID, "act_split_f32: bad argument");
  if (hi_lo_bf16 && (hl_cstride < c_out || hi_off < 0 || lo_off < 0 || hi_off + c_out > hl_cstride || lo_off + c_out > hl_cstride))
    return set_error(SSR_ERR_INVALID, "act_split_f32: hi / lo slices outside the channel stride");
  const int64_t total = static_cast<int64_t>(n) * h * up * w * up * c_out;
  const int block = 256;
  act_split_kernel<<<grid_for(total, block, 148, 16), block, 0, static_cast<cudaStream_t>(stream)>>>(
      z, total, h * up, w * up, c_out, up, act, act_alpha, alpha, res32, y32, static_cast<__nv_bfloat16*>(hi_lo_bf16),
      hl_cstride, hi_off, lo_off);
  SSR_CHECK_LAUNCH("act_split_f32");
  return SSR_OK;
}

extern "C" int ssr_bf16_residual_f32(const float* x, float* y, int64_t count, void* stream) {
  if (!x || !y || count < 0) return set_error(SSR_ERR_INVALID, "bf16_residual_f32: bad argument");
  if (count == 0) return SSR_OK;
  bf16_residual_kernel<<<grid_for(count, 256, 148, 16), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, y, count);
  SSR_CHECK_LAUNCH("bf16_residual_f32");
  return SSR_OK;
}
