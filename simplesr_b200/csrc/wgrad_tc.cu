// wgrad_tc.cu — weight gradient of a stride-1 SAME convolution as a split-K tcgen05 GEMM (Conv2DBackpropFilter of the
// reference's tape backward, sr_model.py:436-438 through model_builder.py:285-293).
//
//   dW[dy,dx,ci,co] = sum_{n,y,x} X[n, y+dy-ph, x+dx-pw, ci] * dZ[n, y, x, co]
//
// GEMM view per tap: M = ci (64-channel chunk), N = co (64-channel chunk), K = pixels.  Both operands are "MN-major"
// for the tensor core: a pixel's 64 channels are the 128 contiguous bytes of one shared-memory row, consecutive pixels
// are consecutive K.  That is exactly the layout the forward kernel's TMA boxes produce, so:
//   * per pixel tile (Hb rows x Wb columns, pitch P = Wb + kw - 1) the producer loads ONE halo box of X
//     {64 ch, P, Hb + kh - 1} and ONE box of dZ {64 ch, P, Hb}.  dZ is addressed through a 5-D tensor map that splits W
//     into (W / Wb, Wb), so the P - Wb pitch columns of the dZ box are out of bounds in that dimension and arrive as
//     ZEROS: the scratch positions of the pitched tile contribute nothing.  Rows beyond H are zero-filled the same way.
//   * tap (dy,dx) is a row offset (dy*P + dx) of the A descriptor.  Two taps share one M = 128 MMA: rows 0-63 are the
//     64 input channels of tap 2a, rows 64-127 those of tap 2a+1, the descriptor's leading byte offset being the
//     distance between the two taps' rows.  Up to 8 such accumulators (16 taps x 64 ci x 64 co fp32) live in TMEM for
//     the whole kernel: the CTA walks its share of the pixel tiles (split-K) and only then drains TMEM once.
//   * the bias gradient (column sums of dZ) rides along: units with ci == 0 run one more accumulator whose A operand is
//     a constant tile of ones (2 KB, K-step invariant), parked in tap slots 14 / 15 of the partials (3x3 kernels use 9).
//   * partial sums go to a workspace [unit][cta][accumulator][float4 column group][128 rows] (row-minor, so that a warp
//     of the drain - one TMEM lane = one row per thread - writes 512 contiguous bytes per instruction); a second kernel
//     adds them in a fixed order (deterministic) into the HWIO fp32 gradient.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstring>

#include "internal.h"
#include "ptx_sm100.cuh"

namespace ssr {

constexpr int kWgThreads = 192;     // warps 0-3: final TMEM drain, warp 4: TMA producer, warp 5: MMA issuer
constexpr int kWgTapsPerGroup = 16; // 8 accumulators x 2 taps
constexpr int kWgSmem = 232448;
constexpr int kWgBiasAcc = 7;       // TMEM accumulator (tap slots 14, 15) of the bias gradient
constexpr int kWgOnesBytes = 2048;  // one K step of an all-ones MN-major A tile: 16 pixel rows x 128 B

constexpr int kWgMaxItems = 8;      // convolutions batched into one launch (same n, h, w, kh, kw)
constexpr int kWgMaxUnits = 64;     // (item, co chunk, ci chunk, tap group) work units of one launch

struct WgradParams {
  CUtensorMap tmap[2 * kWgMaxItems];  // item i: [2i] = X {C, W, H, N}, [2i+1] = dZ {C, Wb, W/Wb, H, N}
  uint32_t unit_desc[kWgMaxUnits];    // item | ci << 8 | co << 16 | tap group << 24
  uint32_t item_bias;                 // bit i: item i also accumulates the column sums of dZ (bias gradient)
  float* partial;       // [unit][cta][16 taps][64][64]
  int kh, kw, P, Hb, Wb, R;
  int tiles_x, tiles_y, n_img, tiles_total;
  int n_ci, n_co, n_groups, ctas_per_unit;
  int ksteps;           // Hb * P / 16
  int stage_bytes, z_offset, tx_bytes, stages;  // z_offset: dZ box inside a stage; tx_bytes: bytes TMA writes per stage
  int xbox_bytes;       // bytes of the X box: the kw-1 pixel rows after it are read by the last K step and must be zero
};

// K-advance and tap offsets are plain address arithmetic: the 128-byte swizzle is a function of the absolute smem
// address, which is how TMA wrote the rows.
__device__ __forceinline__ uint64_t mn_desc(uint32_t saddr, uint32_t lbo_bytes) {
  const uint32_t lo = ((saddr >> 4) & 0x3FFFu) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
  const uint32_t hi = ((1024u >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29);  // SBO = 8 K-rows, version 1, SWIZZLE_128B
  return (static_cast<uint64_t>(hi) << 32) | lo;
}

__global__ void __launch_bounds__(kWgThreads, 1) wgrad_tc_kernel(const __grid_constant__ WgradParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t ctrl = smem_base + p.stages * p.stage_bytes;
  uint8_t* ctrl_gen = smem_gen + p.stages * p.stage_bytes;
  auto bar_full = [&](int s) { return ctrl + 8u * s; };
  auto bar_empty = [&](int s) { return ctrl + 8u * (4 + s); };
  const uint32_t bar_done = ctrl + 8u * 8;
  const uint32_t tmem_slot = ctrl + 8u * 9;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(ctrl_gen + 8u * 9);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int unit = blockIdx.x / p.ctas_per_unit, r = blockIdx.x % p.ctas_per_unit;
  const uint32_t ud = p.unit_desc[unit];
  const int item = ud & 0xFF, ci = (ud >> 8) & 0xFF, co = (ud >> 16) & 0xFF, g = ud >> 24;
  const CUtensorMap* const tmap_x = &p.tmap[2 * item];
  const CUtensorMap* const tmap_z = &p.tmap[2 * item + 1];
  const int taps_total = p.kh * p.kw;
  const int tap0 = g * kWgTapsPerGroup;
  const int ntaps = min(kWgTapsPerGroup, taps_total - tap0);
  const int naccs = (ntaps + 1) >> 1;
  const bool bias_unit = ((p.item_bias >> item) & 1u) && ci == 0 && g == 0;  // this unit also sums dZ over the pixels

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(bar_full(s), 1);
      mbar_init(bar_empty(s), 1);
    }
    mbar_init(bar_done, 1);
    fence_mbar_init();
  }
  if (warp == 5) tmem_alloc(tmem_slot, 512);
  // The last taps' A window runs kw-1 pixel rows past the X box.  Their partner dZ rows are zero (pitch columns), but
  // 0 x garbage must not be NaN: clear the slack between the X box and the dZ box of every stage once (TMA never writes it).
  for (int s = 0; s < p.stages; ++s) {
    uint32_t* slack = reinterpret_cast<uint32_t*>(smem_gen + s * p.stage_bytes + p.xbox_bytes);
    for (int i = threadIdx.x; i < (p.z_offset - p.xbox_bytes) / 4; i += blockDim.x) slack[i] = 0u;
  }
  if (bias_unit) {
    uint32_t* ones = reinterpret_cast<uint32_t*>(ctrl_gen + 2048);
    for (int i = threadIdx.x; i < kWgOnesBytes / 4; i += blockDim.x) ones[i] = 0x3F803F80u;  // bf16 1.0 pairs
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;
  const int my_tiles = (p.tiles_total - r + p.ctas_per_unit - 1) / p.ctas_per_unit;

  if (warp == 4) {
    // ------------------------------------------------------------ TMA producer
    if (elect_one()) {
      prefetch_tmap(tmap_x);
      prefetch_tmap(tmap_z);
    }
    __syncwarp();
    grid_dep_wait();
    int s = 0;
    uint32_t ph = 0;
    const int txy = p.tiles_x * p.tiles_y;
    for (int tile = r; tile < p.tiles_total; tile += p.ctas_per_unit) {
      const int n = tile / txy, rem = tile - n * txy;
      const int ty = rem / p.tiles_x, tx = rem % p.tiles_x;
      mbar_wait_sleep(bar_empty(s), ph ^ 1, 100);
      if (elect_one()) {
        const uint32_t dst = smem_base + s * p.stage_bytes;
        mbar_expect_tx(bar_full(s), p.tx_bytes);
        tma_load_4d(dst, tmap_x, bar_full(s), ci * 64, tx * p.Wb - (p.kw >> 1), ty * p.Hb - (p.kh >> 1), n);
        tma_load_5d(dst + p.z_offset, tmap_z, bar_full(s), co * 64, 0, tx, ty * p.Hb, n);
      }
      __syncwarp();
      if (++s == p.stages) {
        s = 0;
        ph ^= 1;
      }
    }
  } else if (warp == 5) {
    // ------------------------------------------------------------ MMA issuer
    // idesc: D = f32, A = B = bf16, both MN-major (bits 15, 16), N = 64, M = 128
    const uint32_t idesc = umma_idesc_bf16(128, 64) | (1u << 15) | (1u << 16);
    int s = 0;
    uint32_t ph = 0;
    for (int i = 0; i < my_tiles; ++i) {
      mbar_wait(bar_full(s), ph);
      tc_fence_after();
      const uint32_t xs = smem_base + s * p.stage_bytes, zs = xs + p.z_offset;
      if (elect_one()) {
        for (int a = 0; a < naccs; ++a) {
          const int t1 = tap0 + 2 * a, t2 = min(t1 + 1, taps_total - 1);
          const int off1 = (t1 / p.kw) * p.P + (t1 % p.kw), off2 = (t2 / p.kw) * p.P + (t2 % p.kw);
          const uint32_t a_addr = xs + off1 * 128;
          const uint32_t lbo = static_cast<uint32_t>(off2 - off1) * 128u;
          for (int k = 0; k < p.ksteps; ++k) {
            umma_bf16(tmem_base + a * 64, mn_desc(a_addr + k * 2048, lbo), mn_desc(zs + k * 2048, 0), idesc,
                      (i | k) != 0);
          }
        }
        if (bias_unit) {
          // every row of this accumulator = sum over the tile's pixels of dZ[p, co] (pitch columns / rows beyond H are 0)
          for (int k = 0; k < p.ksteps; ++k)
            umma_bf16(tmem_base + kWgBiasAcc * 64, mn_desc(ctrl + 2048, 0), mn_desc(zs + k * 2048, 0), idesc, (i | k) != 0);
        }
        umma_commit(bar_empty(s));
        if (i == my_tiles - 1) umma_commit(bar_done);
      }
      __syncwarp();
      if (++s == p.stages) {
        s = 0;
        ph ^= 1;
      }
    }
  } else {
    // ------------------------------------------------------------ final drain: TMEM -> fp32 partials
    float4* dst = reinterpret_cast<float4*>(p.partial) + static_cast<size_t>(blockIdx.x) * (kWgTapsPerGroup / 2) * 16 * 128;
    if (my_tiles > 0) {
      mbar_wait(bar_done, 0);
      tc_fence_after();
    }
    const int row = warp * 32 + lane;            // accumulator row: (tap parity, ci)
    for (int a = 0; a < kWgTapsPerGroup / 2; ++a) {
      if (!(a < naccs || (bias_unit && a == kWgBiasAcc))) continue;  // slots the reduction never reads
      for (int c0 = 0; c0 < 64; c0 += 16) {
        uint32_t v[16];
        if (my_tiles > 0) {
          tmem_ld16(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + a * 64 + c0, v);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = 0u;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
          dst[(static_cast<size_t>(a) * 16 + (c0 >> 2) + j) * 128 + row] =
              make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                          __uint_as_float(v[4 * j + 3]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// dW[tap][ci][co] (HWIO fp32) = scale * sum over the CTAs of a unit, fixed order.  One thread per (tap, 4 output
// channels, input channel): consecutive threads = consecutive rows of the partial layout = contiguous float4 reads.
// blockIdx.y = item of the batched launch.
struct WgradReduceItem {
  float* dw;
  float* dbias;
  int cin_real, cout, n_ci, n_co, unit_base;
  float scale, bias_scale;
  int accumulate, bias_accumulate;
};
struct WgradReduceParams {
  WgradReduceItem item[kWgMaxItems];
  int taps, n_groups, ctas_per_unit;
};
__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, const __grid_constant__ WgradReduceParams rp) {
  const WgradReduceItem& it = rp.item[blockIdx.y];
  const int taps = rp.taps, cin_real = it.cin_real, cout = it.cout, n_ci = it.n_ci, n_groups = rp.n_groups;
  const int ctas_per_unit = rp.ctas_per_unit;
  float* const dw = it.dw;
  float* const dbias = it.dbias;
  const int nq = (cout + 3) / 4;
  const int64_t total = static_cast<int64_t>(taps) * nq * cin_real;
  const float4* P = reinterpret_cast<const float4*>(partial);
  const size_t cta_stride = static_cast<size_t>(kWgTapsPerGroup / 2) * 16 * 128;  // float4 per CTA
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total + (dbias ? nq : 0);
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const bool is_bias = i >= total;
    int t = 0, ci_g = 0, qg;
    if (is_bias) {
      qg = static_cast<int>(i - total);
    } else {
      ci_g = static_cast<int>(i % cin_real);
      qg = static_cast<int>((i / cin_real) % nq);
      t = static_cast<int>(i / (static_cast<int64_t>(cin_real) * nq));
    }
    const int co0 = 4 * qg;
    const int g = t / kWgTapsPerGroup, tl = t % kWgTapsPerGroup;
    const int unit = it.unit_base + (co0 / 64 * n_ci + (is_bias ? 0 : ci_g / 64)) * n_groups + (is_bias ? 0 : g);
    const int a = is_bias ? kWgBiasAcc : (tl >> 1);
    const int row = is_bias ? 0 : ((tl & 1) * 64 + (ci_g & 63));
    const float4* src = P + static_cast<size_t>(unit) * ctas_per_unit * cta_stride + (static_cast<size_t>(a) * 16 + ((co0 & 63) >> 2)) * 128 + row;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int c = 0; c < ctas_per_unit; ++c) {
      const float4 v = src[static_cast<size_t>(c) * cta_stride];
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    const float sc = is_bias ? it.bias_scale : it.scale;
    const float r[4] = {acc.x * sc, acc.y * sc, acc.z * sc, acc.w * sc};
    float* o = is_bias ? dbias + co0 : dw + (static_cast<int64_t>(t) * cin_real + ci_g) * cout + co0;
    const int acc_flag = is_bias ? it.bias_accumulate : it.accumulate;
#pragma unroll
    for (int e = 0; e < 4; ++e)
      if (co0 + e < cout) o[e] = acc_flag ? o[e] + r[e] : r[e];
  }
}

static int round_up_i(int a, int b) { return (a + b - 1) / b * b; }
static int gcd_i(int a, int b) { return b ? gcd_i(b, a % b) : a; }

struct WgradPlan {
  int Wb, Hb, P, R, ksteps, stage_bytes, xbox, zbox, stages;
  int n_ci, n_co, n_groups, units, ctas_per_unit;
  size_t ws_bytes;
};

// `units_override` > 0: the launch batches several convolutions, units = their total (n_ci / n_co are then per item)
static bool wgrad_plan(int sm_count, int n, int h, int w, int cin, int cout, int kh, int kw, WgradPlan* pl,
                       int units_override = 0) {
  if (kh < 1 || kw < 1 || kh > 9 || kw > 9 || !(kh & 1) || !(kw & 1)) return false;
  pl->n_ci = (cin + 63) / 64;
  pl->n_co = (cout + 63) / 64;
  pl->n_groups = (kh * kw + kWgTapsPerGroup - 1) / kWgTapsPerGroup;
  pl->units = units_override > 0 ? units_override : pl->n_ci * pl->n_co * pl->n_groups;
  if (pl->units > kWgMaxUnits || pl->units > sm_count) return false;
  // Wb: the largest divisor of W whose two-stage tile fits shared memory (Hb then makes Hb*P a multiple of 16)
  int best = 0;
  for (int wb = 1; wb <= w; ++wb) {
    if (w % wb) continue;
    const int P = wb + kw - 1;
    if (P > 256) break;
    const int hb = 16 / gcd_i(P, 16);
    if (hb + kh - 1 > 256) continue;
    const int stage = round_up_i(((hb + kh - 1) * P + kw - 1) * 128, 1024) + round_up_i(hb * P * 128, 1024);
    if (2 * stage + 2048 + kWgOnesBytes > kWgSmem - 1024) continue;
    best = wb;
  }
  if (!best) return false;
  pl->Wb = best;
  pl->P = best + kw - 1;
  pl->Hb = 16 / gcd_i(pl->P, 16);
  // Hb: a multiple of hb0 (keeps Hb*P % 16 == 0) such that two stages fit; among those, the one with the shortest
  // critical path = (tiles per CTA) x (K steps per tile + ~6 steps' worth of per-tile latency).  Small training patches
  // have fewer tiles than CTAs: a smaller tile then spreads the pixels over more CTAs.
  const int hb0 = pl->Hb;
  const int ctas = std::max(1, sm_count / pl->units);
  long long best_cost = -1;
  int best_hb = hb0;
  for (int hb = hb0; hb <= std::max(h, hb0); hb += hb0) {
    const int stage = round_up_i(((hb + kh - 1) * pl->P + kw - 1) * 128, 1024) + round_up_i(hb * pl->P * 128, 1024);
    if (hb > hb0 && (hb + kh - 1 > 256 || 2 * stage + 2048 + kWgOnesBytes > kWgSmem - 1024 || hb * pl->P > 1024)) break;
    const long long tiles = static_cast<long long>(w / pl->Wb) * ((h + hb - 1) / hb) * std::max(n, 1);
    const long long cost = ((tiles + ctas - 1) / ctas) * (hb * pl->P / 16 + 6);
    if (best_cost < 0 || cost < best_cost) {
      best_cost = cost;
      best_hb = hb;
    }
  }
  pl->Hb = best_hb;
  pl->R = pl->Hb + kh - 1;
  pl->ksteps = pl->Hb * pl->P / 16;
  pl->xbox = pl->R * pl->P * 128;
  pl->zbox = pl->Hb * pl->P * 128;
  pl->stage_bytes = round_up_i(pl->xbox + (kw - 1) * 128, 1024) + round_up_i(pl->zbox, 1024);
  pl->stages = std::min(4, (kWgSmem - 1024 - 2048 - kWgOnesBytes) / pl->stage_bytes);
  pl->ctas_per_unit = std::max(1, sm_count / pl->units);
  if (n > 0) {  // never more CTAs than pixel tiles (their partials would be zeros for the reduction to read)
    const long long tiles = static_cast<long long>(w / pl->Wb) * ((h + pl->Hb - 1) / pl->Hb) * n;
    pl->ctas_per_unit = static_cast<int>(std::min<long long>(pl->ctas_per_unit, std::max<long long>(tiles, 1)));
  }
  pl->ws_bytes = static_cast<size_t>(pl->units) * pl->ctas_per_unit * kWgTapsPerGroup * 64 * 64 * sizeof(float);
  return pl->stages >= 2;
}

}  // namespace ssr

using namespace ssr;

extern "C" size_t ssr_conv2d_wgrad_workspace_bytes(ssr_ctx* ctx, int h, int w, int cin, int cout, int kh, int kw) {
  WgradPlan pl;
  if (!ctx || !wgrad_plan(ctx->sm_count, 0, h, w, cin, cout, kh, kw, &pl)) {
    set_error(SSR_ERR_UNSUPPORTED, "conv2d_wgrad: unsupported shape (h=%d w=%d cin=%d cout=%d k=%dx%d)", h, w, cin, cout,
              kh, kw);
    return 0;
  }
  return pl.ws_bytes;
}

static int wgrad_launch(ssr_ctx* ctx, const void* x, int x_cstride, int x_coff, int cin_real, const void* dz,
                        int dz_cstride, int dz_coff, int cout, int n, int h, int w, int kh, int kw, float scale,
                        int accumulate, void* workspace, float* dw_hwio, float* dbias, float bias_scale,
                        int bias_accumulate, void* stream);

extern "C" int ssr_conv2d_wgrad(ssr_ctx* ctx, const void* x, int x_cstride, int x_coff, int cin_real, const void* dz,
                                int dz_cstride, int dz_coff, int cout, int n, int h, int w, int kh, int kw, float scale,
                                int accumulate, void* workspace, float* dw_hwio, void* stream) {
  return wgrad_launch(ctx, x, x_cstride, x_coff, cin_real, dz, dz_cstride, dz_coff, cout, n, h, w, kh, kw, scale, accumulate,
                      workspace, dw_hwio, nullptr, 0.f, 0, stream);
}

extern "C" int ssr_conv2d_wgrad_bias(ssr_ctx* ctx, const void* x, int x_cstride, int x_coff, int cin_real, const void* dz,
                                     int dz_cstride, int dz_coff, int cout, int n, int h, int w, int kh, int kw, float scale,
                                     int accumulate, void* workspace, float* dw_hwio, float* dbias, float bias_scale,
                                     int bias_accumulate, void* stream) {
  if (!dbias) return set_error(SSR_ERR_INVALID, "conv2d_wgrad_bias: dbias is NULL");
  if (kh * kw > 2 * kWgBiasAcc)
    return set_error(SSR_ERR_UNSUPPORTED, "conv2d_wgrad_bias: at most %d taps (use ssr_channel_sum_bf16)", 2 * kWgBiasAcc);
  return wgrad_launch(ctx, x, x_cstride, x_coff, cin_real, dz, dz_cstride, dz_coff, cout, n, h, w, kh, kw, scale, accumulate,
                      workspace, dw_hwio, dbias, bias_scale, bias_accumulate, stream);
}

static int items_units(const ssr_wgrad_item* items, int count, int kh, int kw) {
  const int n_groups = (kh * kw + kWgTapsPerGroup - 1) / kWgTapsPerGroup;
  int units = 0;
  for (int i = 0; i < count; ++i) units += ((items[i].cin_real + 63) / 64) * ((items[i].cout + 63) / 64) * n_groups;
  return units;
}

extern "C" size_t ssr_conv2d_wgrad_multi_workspace_bytes(ssr_ctx* ctx, const ssr_wgrad_item* items, int count, int h,
                                                         int w, int kh, int kw) {
  WgradPlan pl;
  if (!ctx || !items || count < 1 || count > kWgMaxItems ||
      !wgrad_plan(ctx->sm_count, 0, h, w, 64, 64, kh, kw, &pl, items_units(items, count, kh, kw))) {
    set_error(SSR_ERR_UNSUPPORTED, "conv2d_wgrad_multi: unsupported batch (count=%d h=%d w=%d k=%dx%d)", count, h, w, kh, kw);
    return 0;
  }
  return pl.ws_bytes;
}

extern "C" int ssr_conv2d_wgrad_multi(ssr_ctx* ctx, const ssr_wgrad_item* items, int count, int n, int h, int w, int kh,
                                      int kw, void* workspace, void* stream) {
  if (!ctx || !items || !workspace) return set_error(SSR_ERR_INVALID, "conv2d_wgrad_multi: NULL argument");
  if (count < 1 || count > kWgMaxItems) return set_error(SSR_ERR_INVALID, "conv2d_wgrad_multi: 1 <= count <= %d", kWgMaxItems);
  if (n <= 0 || h <= 0 || w <= 0) return set_error(SSR_ERR_INVALID, "conv2d_wgrad_multi: empty");
  const int units = items_units(items, count, kh, kw);
  WgradPlan pl;
  if (!wgrad_plan(ctx->sm_count, n, h, w, 64, 64, kh, kw, &pl, units))
    return set_error(SSR_ERR_UNSUPPORTED, "conv2d_wgrad_multi: unsupported batch (units=%d h=%d w=%d k=%dx%d)", units, h, w,
                     kh, kw);
  WgradParams p;
  memset(&p, 0, sizeof(p));
  WgradReduceParams rp;
  memset(&rp, 0, sizeof(rp));
  int unit = 0;
  int64_t max_total = 0;
  for (int i = 0; i < count; ++i) {
    const ssr_wgrad_item& it = items[i];
    if (!it.x || !it.dz || !it.dw_hwio || it.cin_real <= 0 || it.cout <= 0)
      return set_error(SSR_ERR_INVALID, "conv2d_wgrad_multi: item %d has a NULL / empty field", i);
    if (it.x_cstride % 8 || it.x_coff % 8 || it.dz_cstride % 8 || it.dz_coff % 8)
      return set_error(SSR_ERR_INVALID, "conv2d_wgrad: channel strides / offsets must be multiples of 8");
    if (it.dbias && kh * kw > 2 * kWgBiasAcc)
      return set_error(SSR_ERR_UNSUPPORTED, "conv2d_wgrad: bias gradient rides along for at most %d taps", 2 * kWgBiasAcc);
    // X: dims {C, W, H, N}; the channel extent is what the slice really holds (rounded to 8 for the 16-byte rule):
    // channels beyond it are zero-filled by TMA, so they add nothing to rows >= cin of the accumulator
    const int cx = std::min(round_up_i(it.cin_real, 8), it.x_cstride - it.x_coff);
    const int cz = std::min(round_up_i(it.cout, 8), it.dz_cstride - it.dz_coff);
    {
      cuuint64_t gdim[4] = {static_cast<cuuint64_t>(cx), static_cast<cuuint64_t>(w), static_cast<cuuint64_t>(h),
                            static_cast<cuuint64_t>(n)};
      cuuint64_t gstr[3] = {static_cast<cuuint64_t>(it.x_cstride) * 2, static_cast<cuuint64_t>(it.x_cstride) * 2 * w,
                            static_cast<cuuint64_t>(it.x_cstride) * 2 * w * h};
      cuuint32_t box[4] = {64, static_cast<cuuint32_t>(pl.P), static_cast<cuuint32_t>(pl.R), 1};
      cuuint32_t es[4] = {1, 1, 1, 1};
      CUresult cr = ctx->encode_tiled(&p.tmap[2 * i], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4,
                                      const_cast<uint8_t*>(static_cast<const uint8_t*>(it.x)) + it.x_coff * 2, gdim, gstr,
                                      box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (cr != CUDA_SUCCESS) return set_error(SSR_ERR_CUDA, "wgrad: cuTensorMapEncodeTiled(x) failed (%d)", (int)cr);
    }
    {
      // dZ: dims {C, Wb, W/Wb, H, N}: the box is P wide in dimension 1, so its last kw-1 columns are out of bounds = 0
      cuuint64_t gdim[5] = {static_cast<cuuint64_t>(cz), static_cast<cuuint64_t>(pl.Wb),
                            static_cast<cuuint64_t>(w / pl.Wb), static_cast<cuuint64_t>(h), static_cast<cuuint64_t>(n)};
      cuuint64_t gstr[4] = {static_cast<cuuint64_t>(it.dz_cstride) * 2, static_cast<cuuint64_t>(it.dz_cstride) * 2 * pl.Wb,
                            static_cast<cuuint64_t>(it.dz_cstride) * 2 * w,
                            static_cast<cuuint64_t>(it.dz_cstride) * 2 * w * h};
      cuuint32_t box[5] = {64, static_cast<cuuint32_t>(pl.P), 1, static_cast<cuuint32_t>(pl.Hb), 1};
      cuuint32_t es[5] = {1, 1, 1, 1, 1};
      CUresult cr = ctx->encode_tiled(&p.tmap[2 * i + 1], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5,
                                      const_cast<uint8_t*>(static_cast<const uint8_t*>(it.dz)) + it.dz_coff * 2, gdim, gstr,
                                      box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (cr != CUDA_SUCCESS) return set_error(SSR_ERR_CUDA, "wgrad: cuTensorMapEncodeTiled(dz) failed (%d)", (int)cr);
    }
    const int n_ci = (it.cin_real + 63) / 64, n_co = (it.cout + 63) / 64;
    WgradReduceItem& ri = rp.item[i];
    ri.dw = it.dw_hwio;
    ri.dbias = it.dbias;
    ri.cin_real = it.cin_real;
    ri.cout = it.cout;
    ri.n_ci = n_ci;
    ri.n_co = n_co;
    ri.unit_base = unit;
    ri.scale = it.scale;
    ri.bias_scale = it.bias_scale;
    ri.accumulate = it.accumulate;
    ri.bias_accumulate = it.bias_accumulate;
    if (it.dbias) p.item_bias |= 1u << i;
    for (int co = 0; co < n_co; ++co)        // unit = base + (co * n_ci + ci) * n_groups + g  (the reduction's formula)
      for (int ci = 0; ci < n_ci; ++ci)
        for (int g = 0; g < pl.n_groups; ++g) p.unit_desc[unit++] = i | (ci << 8) | (co << 16) | (g << 24);
    max_total = std::max<int64_t>(max_total, (static_cast<int64_t>(kh) * kw * it.cin_real + (it.dbias ? 1 : 0)) *
                                                 ((it.cout + 3) / 4));
  }
  p.partial = static_cast<float*>(workspace);
  p.kh = kh;
  p.kw = kw;
  p.P = pl.P;
  p.Hb = pl.Hb;
  p.Wb = pl.Wb;
  p.R = pl.R;
  p.tiles_x = w / pl.Wb;
  p.tiles_y = (h + pl.Hb - 1) / pl.Hb;
  p.n_img = n;
  p.tiles_total = p.tiles_x * p.tiles_y * n;
  p.n_groups = pl.n_groups;
  p.ctas_per_unit = pl.ctas_per_unit;
  p.ksteps = pl.ksteps;
  p.z_offset = round_up_i(pl.xbox + (kw - 1) * 128, 1024);  // the dZ box starts on a swizzle-atom boundary
  p.xbox_bytes = pl.xbox;
  p.tx_bytes = pl.xbox + pl.zbox;
  p.stage_bytes = pl.stage_bytes;
  p.stages = pl.stages;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (int rc = opt_in_dynamic_smem(reinterpret_cast<const void*>(wgrad_tc_kernel), kWgSmem, "wgrad_tc_kernel")) return rc;
  const int smem = 1024 + p.stages * p.stage_bytes + 2048 + kWgOnesBytes;
  wgrad_tc_kernel<<<units * pl.ctas_per_unit, kWgThreads, smem, st>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(SSR_ERR_CUDA, "wgrad_tc_kernel launch: %s", cudaGetErrorString(e));
  rp.taps = kh * kw;
  rp.n_groups = pl.n_groups;
  rp.ctas_per_unit = pl.ctas_per_unit;
  const int block = 256;
  const int gx = static_cast<int>(std::min<int64_t>((max_total + block - 1) / block, std::max(1, 148 * 8 / count)));
  wgrad_reduce_kernel<<<dim3(gx, count), block, 0, st>>>(static_cast<const float*>(workspace), rp);
  e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(SSR_ERR_CUDA, "wgrad_reduce launch: %s", cudaGetErrorString(e));
  ctx->launches += 2;
  return SSR_OK;
}

static int wgrad_launch(ssr_ctx* ctx, const void* x, int x_cstride, int x_coff, int cin_real, const void* dz,
                        int dz_cstride, int dz_coff, int cout, int n, int h, int w, int kh, int kw, float scale,
                        int accumulate, void* workspace, float* dw_hwio, float* dbias, float bias_scale,
                        int bias_accumulate, void* stream) {
  if (!ctx || !x || !dz || !workspace || !dw_hwio) return set_error(SSR_ERR_INVALID, "conv2d_wgrad: NULL argument");
  if (n <= 0 || h <= 0 || w <= 0 || cin_real <= 0 || cout <= 0) return set_error(SSR_ERR_INVALID, "conv2d_wgrad: empty");
  ssr_wgrad_item it;
  memset(&it, 0, sizeof(it));
  it.x = x; it.x_cstride = x_cstride; it.x_coff = x_coff; it.cin_real = cin_real;
  it.dz = dz; it.dz_cstride = dz_cstride; it.dz_coff = dz_coff; it.cout = cout;
  it.scale = scale; it.accumulate = accumulate; it.dw_hwio = dw_hwio;
  it.dbias = dbias; it.bias_scale = bias_scale; it.bias_accumulate = bias_accumulate;
  return ssr_conv2d_wgrad_multi(ctx, &it, 1, n, h, w, kh, kw, workspace, stream);
}
