// conv_tc.cu — weight-stationary implicit-GEMM convolution for sm_100a (tcgen05 + TMEM + TMA).
//
// Replaces Keras Conv2D(padding="same", strides=1)+BiasAdd and its elementwise tail
// (reference: simple_sr/utils/models/model_builder.py:285-293 with :85,:90,:335,:338,:349-350,:279,:91-94).
//
// Data flow per CTA (persistent, one CTA per SM):
//   * prologue: the CTA's weight slab  W[tap][chunk][n_slab x 64ch]  (bf16, K-major, pre-swizzled 128B)
//     is bulk-copied into shared memory once and stays there ("weight stationary").
//   * per pixel tile (Hb x Wb outputs) and per 64-channel chunk, ONE TMA box
//     {64ch, P = Wb+ks-1, R = Hb+ks-1} brings the halo tile into a pipeline stage (zero fill outside
//     the image = SAME padding).  The stage is a linear array of R*P "pixel rows" of 128 bytes.
//   * the MMA thread issues, for every tap (dy,dx), tcgen05.mma M=128 x N=n_slab x K=16 whose A
//     descriptor starts (dy*P+dx) pixel rows into the stage: output row m <-> tile pixel
//     (m / P, m % P).  Rows with m % P >= Wb are scratch and are dropped by the epilogue.
//     The 3x3 (or 9x9) window therefore re-reads shared memory, not L2.
//   * accumulators are double-buffered in TMEM; four epilogue warps read them with tcgen05.ld,
//     apply bias / LeakyReLU / PReLU / tanh / residual and store bf16 or fp32 into a channel slice
//     (dense-block concat) or through the depth_to_space(2) address map.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <initializer_list>
#include <type_traits>

#include <cstdlib>
#include <mutex>
#include <set>

#include "internal.h"
#include "ptx_sm100.cuh"

namespace ssr {

constexpr int kEpiWarps = 8;   // warps 0-7: epilogue (TMEM lane quadrant = warp & 3, column half = warp >> 2)
constexpr int kTmaWarp = 8;    // warp 8: TMA producer
constexpr int kMmaWarp0 = 9;   // warps 9..11: MMA issuers (up to 3, each walking its own tiles)
constexpr int kMaxMmaWarps = 3;
constexpr int kThreads = (kEpiWarps + 1 + kMaxMmaWarps) * 32;
constexpr int kAccs = 2 * kMaxMmaWarps;  // TMEM accumulators (two per issuing warp)
constexpr int kMaxStages = 8;
// Carry format of the growth-conv pairs: 32 partial sums per accumulator row, as fp16 (round to nearest, saturating) or
// raw fp32.  The tails are L2-bandwidth-bound and the carry is half of their bytes; fp16 keeps 11 significant bits, four
// times finer than the bf16 rounding every stored activation gets anyway.  (|partial sum| > 65504 saturates: the dense
// blocks' activations are O(1).)
constexpr bool kCarryF16 = true;
constexpr int kCarryQ = kCarryF16 ? 4 : 8;             // 16-byte groups per accumulator row
constexpr int kCarryTileBytes = kCarryQ * 128 * 16;    // one pixel tile: [16-byte group][128 rows]
constexpr int kCarrySlots = 4;  // carry tiles in shared memory per epilogue group (carry_in kernels)
constexpr int kMaxNSlab = 128;
constexpr int kSmemBytes = 232448;  // 227 KB opt-in maximum
constexpr int kSmemCtrlBytes = 2048;  // barriers + bias/alpha staging

struct ConvKParams {
  CUtensorMap tmap;
  const uint8_t* wpack;
  const float* bias;
  const float* alpha;
  void* out;
  void* out2;
  const void* res;
  int out_dtype, res_dtype;
  int out_cstride, out_coff, out2_cstride, out2_coff, res_cstride, res_coff;
  int OH, OW, up;
  int n_img, H, W;
  int nchunks, ksteps_last;
  int n_slab, n_slabs, n_store;  // n_slab: MMA N = TMEM columns per accumulator = epilogue columns
  int w_rows;                    // weight rows per (tap, chunk) tile in THIS CTA's smem (n_slab, or n_slab/2 for CTA pairs)
  int kh, kw, Wb, Hb, P;
  int tiles_x, tiles_y, tiles_total, ctas_per_slab;
  int stages, stage_bytes, box_bytes, w_bytes;
  // Second tile geometry for a ragged bottom strip (H % Hb != 0): the strip's few rows are covered by WIDE tiles
  // (Wb2 x Hb2, pitch P2) instead of a row of mostly empty Wb x Hb tiles.  Tiles [0, nA_total) are geometry A
  // (tiles_x x tiles_yA per image, image-major), tiles [nA_total, tiles_total) geometry B (tiles_x2 per image, rows
  // [y2, H)).  128x128 images: 8 x 18 + 4 = 148 tiles per image instead of 152 - at batch 16 exactly 16 rounds over 148
  // SMs instead of 17.  nA_total == tiles_total: unused.
  CUtensorMap tmap2;
  int nA_total, tiles_yA, Wb2, Hb2, P2, tiles_x2, y2, box_bytes2;
  int act;
  float act_alpha, res_beta;
  int dbg_flags;
  uint32_t tmem_cols;
  int mma_warps;     // 1..3 issuing warps (see the parity note in the kernel)
  // growth-conv pairing (EPI bits 16 / 32): the trailing n_slab - n_act accumulator columns are the partial sums of the
  // NEXT conv over the same input and leave as raw fp32 (carry_out [pixels, n_slab - n_act]); the next conv adds them
  // to its own accumulator before bias + activation (carry_in [pixels, n_slab])
  const float* carry_in;
  float* carry_out;
  int n_act;
  // fused activation backward (EPI bit 128): output columns [mask_lo, mask_lo + mask_n) of this conv, times
  // (mask_z > 0 ? 1 : mask_alpha), also go to mask_out [pixels, mask_out_cstride] - the dZ of the layer whose
  // gradient slice this dgrad conv completes (training.py, dense blocks)
  const void* mask_z;
  void* mask_out;
  int mask_cstride, mask_coff, mask_out_cstride, mask_lo, mask_n;
  float mask_alpha;
  int tile_rev;         // walk the pixel tiles from the last to the first (desc.tile_order): layers alternate, so each one
                        // reads first what its producer wrote last and is still in L2
  int row16;            // 16-byte units per pixel row of a stage / weight row: 8 (64 channels, 128B swizzle) or 4 (cin <= 32: 64B swizzle)
  int epi_stage_bytes;  // per-warp staging of the specialised epilogue (32 rows x min(n_slab, 64) bf16), after the control block
  long long* trace;  // debug: CTA 0 records clock64() per role/event (3 x 512 entries)
  // Tile-level dependencies between consecutive launches of a stream (ssr_conv_chain_*, see ChainState below): chain[0] =
  // epoch base, chain[1] = waits that gave up, flag arrays from chain + kChainHdr.  chain_pub_off != 0: after a pixel
  // tile's stores the epilogue writes epoch + chain_pub_ord into flag[tile].  chain_dep_off != 0: this launch does NOT
  // wait for the whole previous grid (griddepcontrol.wait); instead every tile waits until the 3x3 neighbourhood of
  // tiles of the previous launch (same tile decomposition, checked by the host) carries epoch + chain_dep_ord.
  // chain[kChainCnt + ord] counts the published tiles of launch ord (zeroed by the epoch bump): once it reaches
  // tiles_total the whole previous grid is done and the per-tile polling stops.
  uint32_t* chain;
  int chain_pub_off, chain_dep_off;
  uint32_t chain_pub_ord, chain_dep_ord;
};
constexpr uint32_t kChainStride = 4096;  // epoch bump per ssr_conv_chain_begin: > launches of one sequence
constexpr int kChainCnt = 16;            // chain buffer: 16 header words, kChainStride per-launch tile counters, flag arrays
constexpr int kChainHdr = kChainCnt + static_cast<int>(kChainStride);
constexpr uint32_t kChainSpinLimit = 1u << 22;  // ~1 s of polling, then the wait gives up (counted in chain[1])

// Development instrumentation (per-role clock64 timeline, launch-boundary stamps, the A/B flags of the epilogue wait and
// the carry's L2 policies) is only compiled with `make clean all EXTRA=-DSSR_DEV`: measured on one box, the timeline
// stamps alone cost C2 2.1 % and the ESRGAN step 2.5 % while switched off (the three warp roles share a small
// instruction cache), the boundary stamps another ~1 %.  tools/gpu_trace*.py and tools/gpu_boundary.py need that build.
#ifdef SSR_DEV
constexpr bool kDev = true;
#else
constexpr bool kDev = false;
#endif
#ifndef SSR_DEV
#define SSR_TRACE(role, idx) do { } while (0)
#else
#define SSR_TRACE(role, idx)                                                        \
  do {                                                                            \
    if (p.trace != nullptr && blockIdx.x == 0 && (idx) < 512) p.trace[(role)*512 + (idx)] = clock64(); \
  } while (0)
#endif

// wall-clock (globaltimer, ns) stamps of CTA 0 at the launch boundaries: entries 500.. of role 0 (SSR_DEV builds)
constexpr bool kTraceBoundary = kDev;
#define SSR_TRACE_G(idx)                                                                         \
  do {                                                                                           \
    if (kTraceBoundary && p.trace != nullptr && blockIdx.x == 0) {                                                 \
      unsigned long long t__;                                                                    \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t__));                                   \
      p.trace[500 + (idx)] = static_cast<long long>(t__);                                        \
    }                                                                                            \
  } while (0)

#ifdef SSR_WATCHDOG
// debug build: a wait that does not complete records where it was (p.trace must be host-pinned memory) and traps
#define SSR_WD_WAIT(bar, par)                                                              \
  do {                                                                                     \
    if (!mbar_wait_bounded((bar), (par))) {                                                \
      if (p.trace != nullptr && atomicAdd(reinterpret_cast<unsigned long long*>(p.trace), 1ull) < 15) { \
        long long* t__ = p.trace + 8 * (1 + (threadIdx.x >> 5) % 12 + 12 * (blockIdx.x != 0));          \
        t__[0] = __LINE__; t__[1] = blockIdx.x; t__[2] = threadIdx.x;                       \
        t__[3] = static_cast<long long>((bar) - ctrl_smem) / 8; t__[4] = (par);               \
        __threadfence_system();                                                              \
      }                                                                                      \
      __nanosleep(2000000);                                                                  \
      __trap();                                                                              \
    }                                                                                        \
  } while (0)
#undef SSR_TRACE
#define SSR_TRACE(role, idx) do { } while (0)
#define mbar_wait(bar, par) SSR_WD_WAIT(bar, par)
#define mbar_wait_sleep(bar, par, ns) SSR_WD_WAIT(bar, par)
#endif

// v[i] = act(acc[i] + bias[i]) for 16 consecutive channels.  The activation is a template parameter so that the
// (warp-uniform) dispatch happens once per 16 channels; bias / PReLU slopes come in registers (loaded from shared
// memory while the TMEM load is in flight).
template <int ACT>
__device__ __forceinline__ void bias_act16(const uint32_t* __restrict__ r, const float4* __restrict__ b4,
                                           const float* __restrict__ s_alpha16, float lrelu_alpha, float (&v)[16]) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float bb[4] = {b4[q].x, b4[q].y, b4[q].z, b4[q].w};
    float aa[4] = {lrelu_alpha, lrelu_alpha, lrelu_alpha, lrelu_alpha};
    if (ACT == SSR_ACT_PRELU) {  // per-channel slopes (SRResNet only): read from shared memory here
      const float4 a = *reinterpret_cast<const float4*>(s_alpha16 + 4 * q);
      aa[0] = a.x; aa[1] = a.y; aa[2] = a.z; aa[3] = a.w;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float x = __uint_as_float(r[4 * q + j]) + bb[j];
      if (ACT == SSR_ACT_LRELU || ACT == SSR_ACT_PRELU) x = x > 0.f ? x : aa[j] * x;
      if (ACT == SSR_ACT_TANH) x = tanhf(x);
      if (ACT == SSR_ACT_RELU) x = fmaxf(x, 0.f);
      v[4 * q + j] = x;
    }
  }
}

// EPI < 0: generic epilogue (any dtype / channel count, runtime dispatch).  EPI >= 0: specialised bf16 epilogue with
// activation EPI & 7 and a bf16 residual iff EPI & 8 (n_store == n_slab, 16-byte aligned slices), EPI & 16: split
// accumulator (columns >= n_act leave as fp32 carry), EPI & 32: fp32 carry added before bias / activation;
// keeping it small matters: the three warp roles share a tiny instruction cache.
// PAIR: two CTAs of a cluster form one M = 256 MMA (cta_group::2): each loads its own 128-pixel tile and HALF of the
// weight rows, the leader CTA issues for both.  Used when a full-N weight slab does not fit one SM (192 -> 64).
// CHAIN: the tile-dependency code (ssr_conv_chain_*) is only compiled into the twins of the kernels a dense-block chain
// uses; measured, it costs the hot kernels 2-4 % even when idle (registers, instruction cache).
template <int KS, int EPI, bool PAIR, bool CHAIN = false>
__global__ void __launch_bounds__(kThreads, 1) conv_tc_kernel(const __grid_constant__ ConvKParams p) {
  constexpr bool kChain = CHAIN;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // Dynamic shared memory is only guaranteed 16B aligned: realign to 1024 (swizzle-128B atoms).
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  const uint32_t w_smem = smem_base;
  const uint32_t stage_smem = w_smem + p.w_bytes;
  const uint32_t ctrl_smem = stage_smem + p.stages * p.stage_bytes;
  uint8_t* ctrl_gen = smem_gen + p.w_bytes + p.stages * p.stage_bytes;

  // control block layout (8-byte barriers first)
  // Two "full" barriers per stage, used by alternate passes over the ring: several MMA warps wait on the ring out of
  // order, TMA boxes land out of order, and a parity wait is only meaningful within one phase of its barrier.  With
  // the passes split by parity a wait for pass r needs pass r - 2 of that stage to have landed, which the in-order
  // producer guarantees once the warp's previous chunk was issued (host: (nw - 1) * nchunks + 1 <= stages).
  auto bar_full = [&](int s, uint32_t pass) { return ctrl_smem + 8u * (1 + 2 * s + (pass & 1)); };
  auto bar_empty = [&](int s) { return ctrl_smem + 8u * (1 + 2 * kMaxStages + s); };
  auto bar_tfull = [&](int a) { return ctrl_smem + 8u * (1 + 3 * kMaxStages + a); };
  auto bar_tempty = [&](int a) { return ctrl_smem + 8u * (1 + 3 * kMaxStages + kAccs + a); };
  const uint32_t tmem_slot = ctrl_smem + 8u * (1 + 3 * kMaxStages + 2 * kAccs);
  volatile uint32_t* tmem_slot_gen =
      reinterpret_cast<volatile uint32_t*>(ctrl_gen + 8u * (1 + 3 * kMaxStages + 2 * kAccs));
  float* s_bias = reinterpret_cast<float*>(ctrl_gen + 1024);
  float* s_alpha = reinterpret_cast<float*>(ctrl_gen + 1024 + kMaxNSlab * 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) SSR_TRACE_G(0);  // kernel entry

  // work decomposition: this CTA handles tiles tile_first + it * tile_step; in PAIR mode the two CTAs of a cluster take
  // tiles 2v and 2v + 1 (crank) of pair-tile v and run the same number of iterations (the odd one out is a dummy tile
  // whose TMA box is entirely out of bounds = zeros and whose pixels are never stored).
  // PAIR with several output-channel slabs (deep layers, cin >= 256: only 16 or 32 weight rows fit one SM): pair pv
  // works on slab pv / ctas_per_slab - weight slabs 2 * slab and 2 * slab + 1 of the ordinary multi-slab image, one per
  // CTA - so the MMA runs N = 2 x (rows per SM) wide and each activation box is fetched once per TWO slabs.
  const int crank = PAIR ? static_cast<int>(cluster_ctarank()) : 0;
  const int pv = blockIdx.x >> 1;                                     // PAIR: pair index
  const int slab = PAIR ? pv / p.ctas_per_slab : blockIdx.x / p.ctas_per_slab;  // output-channel slab of the epilogue
  const int wslab = PAIR ? 2 * slab + crank : slab;                   // weight slab resident in this CTA
  const int tile_first = PAIR ? 2 * (pv % p.ctas_per_slab) + crank : blockIdx.x % p.ctas_per_slab;
  const int tile_step = PAIR ? 2 * p.ctas_per_slab : p.ctas_per_slab;
  const int tile_end = p.tiles_total + crank;                         // loop bound: tile - crank < tiles_total
  // The weight slab arrives in groups, in the order the MMAs of a CTA's FIRST tile use it: chunk 0 one kernel row at a
  // time (groups 0-2; KS != 3: everything in group 0), then one group per further 64-channel chunk (group 2 + ch).  The
  // first tile starts after a ninth (a third for one-chunk layers) of the slab instead of all of it - at one pixel
  // tile per SM (training patches) the slab load was the longest serial piece of a launch.
  // PAIR: bar_wpg(g) on the leader = the peer's group g has landed (forwarded by an idle warp of the peer).
  constexpr int kWGroups = 3 + 7;  // <= 8 chunks
  constexpr int kWgBase = 3 + 3 * kMaxStages + 2 * kAccs + 2 * kCarrySlots;
  auto bar_wg = [&](int g) { return ctrl_smem + 8u * (kWgBase + g); };
  auto bar_wpg = [&](int g) { return ctrl_smem + 8u * (kWgBase + kWGroups + g); };
  const int n_wgroups = (KS == 3) ? 2 + p.nchunks : 1;
  // carry_in kernels: each epilogue group keeps the fp32 carry tiles of its next kCarrySlots tiles in shared memory
  // (16 KB contiguous per tile in the tile-major layout, one bulk copy each, issued three tiles ahead).  Loads issued at
  // the top of a tile exposed the loaded-DRAM latency, and one tile of lookahead is only 32 KB in flight per SM
  // (~3 TB/s over the chip by Little's law).  Slot index: kCarrySlots * group + iteration % kCarrySlots.
  auto bar_cfull = [&](int a) { return ctrl_smem + 8u * (3 + 3 * kMaxStages + 2 * kAccs + a); };
  const uint32_t carry_smem = ctrl_smem + kSmemCtrlBytes + kEpiWarps * p.epi_stage_bytes;
  const int pad_y = (KS == 3) ? 1 : (p.kh >> 1), pad_x = (KS == 3) ? 1 : (p.kw >> 1);  // KS == 0: runtime kh x kw

  // the ~64 barriers are initialised by three threads of different warps side by side (the prologue is on the critical
  // path of every launch: the next layer's CTA only becomes resident when this SM's previous CTA has exited)
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(bar_full(s, 0), 1);
      mbar_init(bar_full(s, 1), 1);
      mbar_init(bar_empty(s), 1);
    }
    fence_mbar_init();
  } else if (threadIdx.x == 32) {
    for (int a = 0; a < kAccs; ++a) {
      mbar_init(bar_tfull(a), 1);
      mbar_init(bar_tempty(a), PAIR ? kEpiWarps : kEpiWarps / 2);  // PAIR: the peer's epilogue warps arrive remotely
    }
    for (int a = 0; a < 2 * kCarrySlots; ++a) mbar_init(bar_cfull(a), 1);
    fence_mbar_init();
  } else if (threadIdx.x == 64) {
    for (int g = 0; g < kWGroups; ++g) {
      mbar_init(bar_wg(g), 1);
      mbar_init(bar_wpg(g), 1);
    }
    fence_mbar_init();
  }
  if (PAIR) cluster_sync_all();  // both CTAs' barriers are initialised before anyone signals across the pair
  if (warp == kMmaWarp0) {
    if (PAIR) tmem_alloc2(tmem_slot, p.tmem_cols); else tmem_alloc(tmem_slot, p.tmem_cols);
  }
  // bias / PReLU slopes: the global loads are issued here and land in shared memory after the block-wide barrier, behind
  // a barrier of the epilogue warps only - the producer and the MMA warps do not wait for them (n_slab <= 128 < 256)
  float bias_r = 0.f, alpha_r = p.act_alpha;
  if (warp < kEpiWarps && static_cast<int>(threadIdx.x) < p.n_slab) {
    if (p.bias) bias_r = __ldg(p.bias + slab * p.n_slab + threadIdx.x);
    if (p.alpha) alpha_r = __ldg(p.alpha + (p.up == 2 ? 0 : slab * p.n_slab) + threadIdx.x);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;
  if (threadIdx.x == 0) SSR_TRACE_G(1);  // barriers initialised, TMEM allocated
  // PDL: let the next layer's CTAs be scheduled as soon as an SM frees up (they prefetch their weights and then block in
  // griddepcontrol.wait until this whole grid has completed), which hides launch latency and the tile-count imbalance.
  // (the head of a chain triggers later, once the epoch base its dependents read is known to be written)
  const bool chain_dep = kChain && p.chain_dep_off != 0;
  const bool chain_pub = kChain && p.chain_pub_off != 0;
  const bool chain_head = chain_pub && !chain_dep;
  if (!chain_head) grid_dep_launch();

  // Tile it of this CTA (tile index rank + it * ctas_per_slab) is issued by MMA warp it % nw into TMEM accumulator
  // acc(it) = it % nw + nw * ((it / nw) & 1): every accumulator is owned by one issuing warp.
  const int nw = p.mma_warps;
  const int txyA = p.tiles_x * p.tiles_yA;
  // tile index (image order) -> geometry and position.  Out-of-range indices (the dummy tile of an odd CTA pair) land in
  // image n_img: the TMA box is entirely out of bounds = zeros, and nothing is stored.
  auto tile_is_b = [&](int tq) { return tq >= p.nA_total && tq < p.tiles_total; };
  auto tile_pos = [&](int tq, bool gb, int& n, int& y0, int& x0) {
    if (tq < 0 || tq >= p.tiles_total) {
      n = p.n_img;
      y0 = x0 = 0;
    } else if (gb) {
      const int t2 = tq - p.nA_total;
      n = t2 / p.tiles_x2;
      y0 = p.y2;
      x0 = (t2 - n * p.tiles_x2) * p.Wb2;
    } else {
      n = tq / txyA;
      const int rem = tq - n * txyA;
      y0 = (rem / p.tiles_x) * p.Hb;
      x0 = (rem % p.tiles_x) * p.Wb;
    }
  };
  // the geometry of a loop tile: in PAIR mode both CTAs (and the dummy) follow the pair's even tile - the host only uses
  // geometry B when no pair straddles the A / B boundary in either tile order
  auto loop_tile_q = [&](int tile) { return p.tile_rev ? p.tiles_total - 1 - tile : tile; };
  auto loop_tile_is_b = [&](int tile) { return tile_is_b(loop_tile_q(PAIR ? (tile & ~1) : tile)); };

  if (warp == kTmaWarp) {
    // ------------------------------------------------------------ TMA producer (warp-uniform loop, one elected lane issues)
    if (elect_one()) {
      prefetch_tmap(&p.tmap);
      if (p.nA_total < p.tiles_total) prefetch_tmap(&p.tmap2);
      // weight slab: contiguous, pre-swizzled image (does not depend on the previous layer)
      const uint8_t* wsrc = p.wpack + static_cast<size_t>(wslab) * p.w_bytes;
      const uint32_t tile_b = p.w_rows * p.row16 * 16;        // one (tap, chunk) weight tile
      const int taps = (KS == 3) ? 9 : p.kh * p.kw;
      if (KS == 3) {
        for (int ch = 0; ch < p.nchunks; ++ch) {
          for (int t = 0; t < 9; ++t) {
            const int g = (ch == 0) ? t / 3 : 2 + ch;
            if (t == 0 || (ch == 0 && t % 3 == 0)) mbar_expect_tx(bar_wg(g), (ch == 0 ? 3 : 9) * tile_b);
            const uint32_t off = (t * p.nchunks + ch) * tile_b;
            bulk_load(w_smem + off, wsrc + off, tile_b, bar_wg(g));
          }
        }
      } else {
        mbar_expect_tx(bar_wg(0), p.w_bytes);
        for (int off = 0; off < p.w_bytes; off += 32768) {
          const int sz = min(32768, p.w_bytes - off);
          bulk_load(w_smem + off, wsrc + off, sz, bar_wg(0));
        }
      }
      (void)taps;
    }
    __syncwarp();
    uint32_t dep_target = 0;
    bool dep_dead = false;
    if (chain_dep) {
      dep_target = ld_acquire_gpu_u32(p.chain) + p.chain_dep_ord;
    } else {
      grid_dep_wait();  // PDL: activations of the previous layer are complete and visible from here on
      if (chain_head) grid_dep_launch();
    }
    if (lane == 0) SSR_TRACE_G(2);  // the previous grid has completed
    const uint32_t* const dep_flags = p.chain + p.chain_dep_off;
    const uint32_t* const dep_cnt = p.chain + kChainCnt + p.chain_dep_ord;
    // chained launch: the halo box of tile (n, y0, x0) reads the previous launch's tiles around it - lane l polls one of
    // them, lane 31 the previous launch's tile counter: once that grid is complete (dep_dead) nothing is polled any more
    auto dep_wait_tile = [&](bool gb, int n, int y0, int x0) {
      int idx = -1;
      const bool has_b = p.nA_total < p.tiles_total;
      if (!gb) {
        const int ty = y0 / p.Hb, tx = x0 / p.Wb;
        if (lane < 9) {
          const int yy = ty + lane / 3 - 1, xx = tx + lane % 3 - 1;
          if (yy >= 0 && yy < p.tiles_yA && xx >= 0 && xx < p.tiles_x) idx = n * txyA + yy * p.tiles_x + xx;
        } else if (has_b && ty == p.tiles_yA - 1) {
          const int lo = max(0, (x0 - 1) / p.Wb2), hi = min(p.tiles_x2 - 1, (x0 + p.Wb) / p.Wb2);
          const int t2 = lo + (lane - 9);
          if (t2 <= hi) idx = p.nA_total + n * p.tiles_x2 + t2;
        }
      } else {
        const int t2 = x0 / p.Wb2;
        if (lane < 3) {
          const int xx = t2 + lane - 1;
          if (xx >= 0 && xx < p.tiles_x2) idx = p.nA_total + n * p.tiles_x2 + xx;
        } else {
          const int lo = max(0, (x0 - 1) / p.Wb), hi = min(p.tiles_x - 1, (x0 + p.Wb2) / p.Wb);
          const int xx = lo + (lane - 3);
          if (xx <= hi) idx = n * txyA + (p.tiles_yA - 1) * p.tiles_x + xx;
        }
      }
      bool ok = idx < 0;
      for (uint32_t spins = 0;; ++spins) {
        bool all_done = false;
        if (lane == 31) all_done = ld_acquire_gpu_u32(dep_cnt) >= static_cast<uint32_t>(p.tiles_total);
        else if (!ok) ok = static_cast<int>(ld_acquire_gpu_u32(dep_flags + idx) - dep_target) >= 0;
        if (__any_sync(0xffffffffu, all_done)) {
          dep_dead = true;
          break;
        }
        if (__all_sync(0xffffffffu, ok)) break;
        if (spins > kChainSpinLimit) {
          if (lane == 0) atomicAdd(p.chain + 1, 1u);
          dep_dead = true;
          break;
        }
        __nanosleep(64);
      }
      fence_proxy_async_all();
    };
    int tr_i = 0;
    int s = 0;
    uint32_t pass = 0;  // passes over the ring
    for (int tile = tile_first; tile < tile_end; tile += tile_step) {
      const bool gb = loop_tile_is_b(tile);
      int n, y0, x0;
      tile_pos(loop_tile_q(tile), gb, n, y0, x0);
      if (chain_dep && !dep_dead && n < p.n_img) dep_wait_tile(gb, n, y0, x0);
      y0 -= pad_y;
      x0 -= pad_x;
      const CUtensorMap* const tm = gb ? &p.tmap2 : &p.tmap;
      const uint32_t bytes = gb ? p.box_bytes2 : p.box_bytes;
      for (int ch = 0; ch < p.nchunks; ++ch) {
        mbar_wait_sleep(bar_empty(s), (pass & 1) ^ 1, 100);
        const uint32_t bfull = bar_full(s, pass);
        if (elect_one()) {
          if (PAIR) {
            // both boxes complete on the LEADER's barrier, which expects the bytes of the pair
            if (crank == 0) mbar_expect_tx(bfull, 2 * bytes);
            tma2_load_4d(stage_smem + s * p.stage_bytes, tm, bfull, ch * 64, x0, y0, n);
          } else {
            mbar_expect_tx(bfull, bytes);
            tma_load_4d(stage_smem + s * p.stage_bytes, tm, bfull, ch * 64, x0, y0, n);
          }
          SSR_TRACE(0, tr_i);
          if (tr_i == 0) SSR_TRACE_G(3);  // first activation box issued
        }
        __syncwarp();
        ++tr_i;
        if (++s == p.stages) {
          s = 0;
          ++pass;
        }
      }
    }
  } else if (warp >= kMmaWarp0) {
    // ------------------------------------------------------------ MMA issuers
    // A single thread cannot issue tcgen05.mma faster than ~43-46 clk/instruction, above the tensor-pipe time of the
    // N <= 64 shapes (32-40 clk: shared-memory operand reads), and it pays ~600 clk of barrier waits per tile.  So up to
    // three warps issue concurrently, each walking its own tiles (it = w, w + nw, ...) over the shared stage ring.
    // Each warp walks its own tiles over the shared stage ring; its consecutive waits are (nw - 1) * nchunks + 1 stages
    // apart, which the host keeps <= stages when it picks nw (see the note at the barrier layout).
    const int mw = warp - kMmaWarp0;
    const uint32_t idesc = umma_idesc_bf16(PAIR ? 256 : 128, p.n_slab);
    const uint32_t r16 = p.row16;                           // 16 B units per operand row (8: 128B swizzle, 4: 64B swizzle)
    const uint32_t b_tile16 = p.w_rows * r16;               // one [w_rows x row] weight tile, in 16 B units
    const uint32_t b_tap16 = p.nchunks * b_tile16;          // weight tiles of consecutive taps
    const uint32_t desc_hi = umma_desc_hi(128 * r16, r16 == 8 ? 2 : 4);  // SBO = 8 rows; SWIZZLE_128B / SWIZZLE_64B
    const uint32_t P8a = p.P * r16, P8b = p.P2 * r16;       // one tile row of pixels in 16 B units (geometry A / B)
    const uint32_t b_lo0 = umma_desc_lo(w_smem) | (1u << 16);  // low descriptor word: (addr >> 4) | LBO field = 1
    const bool issuer = mw < nw && crank == 0;   // PAIR: only the leader CTA issues (for both SMs)
    if (PAIR && crank == 1 && mw == 0) {
      // peer CTA: this warp has no MMAs to issue - it tells the leader which weight groups have landed here
      for (int g = 0; g < n_wgroups; ++g) {
        mbar_wait(bar_wg(g), 0);
        if (elect_one()) mbar_arrive_cluster(leader_cta_addr(bar_wpg(g)));
        __syncwarp();
      }
    }
    auto wait_wgroup = [&](int g) {  // first tile of a warp only: weight group g is resident (in both CTAs of a pair)
      mbar_wait(bar_wg(g), 0);
      if (PAIR) mbar_wait(bar_wpg(g), 0);
    };
    int s = 0;
    uint32_t pass = 0;
    auto advance = [&](int nst) {
      s += nst;
      while (s >= p.stages) {
        s -= p.stages;
        ++pass;
      }
    };
    advance(mw * p.nchunks);
    int u = 0;  // tiles this warp has issued
    for (int it = mw, tile = tile_first + mw * tile_step; issuer && tile < tile_end; it += nw, tile += nw * tile_step, ++u) {
      const int acc = mw + nw * (u & 1);
      const uint32_t P8 = loop_tile_is_b(tile) ? P8b : P8a;
      // waiting for the epilogue: back off (a tight probe loop takes issue slots from the epilogue warps on this scheduler)
      mbar_wait_sleep(bar_tempty(acc), ((u >> 1) & 1) ^ 1, 64);
      const uint32_t d_tmem = tmem_base + acc * p.n_slab;
      for (int ch = 0; ch < p.nchunks; ++ch) {
        mbar_wait(bar_full(s, pass), (pass >> 1) & 1);
        tc_fence_after();
        const uint32_t a_desc = umma_desc_lo(stage_smem + s * p.stage_bytes) | (1u << 16);
        const uint32_t b_desc = b_lo0 + ch * b_tile16;
        const int ksteps = (ch == p.nchunks - 1) ? p.ksteps_last : 4;
        const bool first = (u == 0);
        if (first && (KS != 3 || ch > 0)) wait_wgroup(KS == 3 ? 2 + ch : 0);
        // elect.sync directly at the branch: ptxas then knows a single lane runs the block and keeps the descriptor
        // arithmetic on the uniform datapath (2-3 instructions per MMA instead of ~11 with R2UR round trips).
        if (elect_one()) {
          if (ch == 0) { SSR_TRACE(1, 4 * it); if (it == 0) SSR_TRACE_G(4); }  // first MMA issued
          if (KS == 3) {
            // hot path: 9 * ksteps back-to-back MMAs, fully unrolled per K-step count (a partial last chunk - 16, 32 or 48
            // channels - is as common as a full one: cin = 96, 160, the 32-channel tails of the paired growth convs)
            auto issue = [&](auto KSTEPS) {
#pragma unroll
              for (int t = 0; t < 9; ++t) {
                if (t % 3 == 0 && first && ch == 0) wait_wgroup(t / 3);  // uniform; false after the warp's first tile
                const uint32_t a_tap = a_desc + ((t / 3) * P8 + (t % 3) * r16);
                const uint32_t b_tap = b_desc + t * b_tap16;
#pragma unroll
                for (int k = 0; k < decltype(KSTEPS)::value; ++k) {
                  if (PAIR)
                    umma2_bf16(d_tmem, desc64(desc_hi, a_tap + 2 * k), desc64(desc_hi, b_tap + 2 * k), idesc,
                               (ch | t | k) != 0);
                  else
                    umma_bf16(d_tmem, desc64(desc_hi, a_tap + 2 * k), desc64(desc_hi, b_tap + 2 * k), idesc,
                              (ch | t | k) != 0);
                }
              }
            };
            switch (ksteps) {
              case 4: issue(std::integral_constant<int, 4>{}); break;
              case 3: issue(std::integral_constant<int, 3>{}); break;
              case 2: issue(std::integral_constant<int, 2>{}); break;
              default: issue(std::integral_constant<int, 1>{}); break;
            }
          } else {
            const int kh = (KS == 3) ? 3 : p.kh, kw = (KS == 3) ? 3 : p.kw;
            for (int dy = 0; dy < kh; ++dy) {
              for (int dx = 0; dx < kw; ++dx) {
                const int t = dy * kw + dx;
                const uint32_t a_tap = a_desc + (dy * P8 + dx * r16);
                const uint32_t b_tap = b_desc + t * b_tap16;
                for (int k = 0; k < ksteps; ++k) {
                  if (PAIR)
                    umma2_bf16(d_tmem, desc64(desc_hi, a_tap + 2 * k), desc64(desc_hi, b_tap + 2 * k), idesc,
                               (ch | t | k) != 0);
                  else
                    umma_bf16(d_tmem, desc64(desc_hi, a_tap + 2 * k), desc64(desc_hi, b_tap + 2 * k), idesc,
                              (ch | t | k) != 0);
                }
              }
            }
          }
          // stage reusable once these MMAs have read it; accumulator complete after the last chunk.
          // PAIR: the arrivals are multicast to the barrier at the same offset in both CTAs.
          if (PAIR) umma2_commit_mc(bar_empty(s), 3); else umma_commit(bar_empty(s));
          if (ch == p.nchunks - 1) {
            if (PAIR) umma2_commit_mc(bar_tfull(acc), 3); else umma_commit(bar_tfull(acc));
            SSR_TRACE(1, 4 * it + 1);
          }
        }
        __syncwarp();
        advance(1);
      }
      advance((nw - 1) * p.nchunks);  // the other MMA warps consume the next tiles' stages
    }
  } else {
    // ------------------------------------------------------------ epilogue
    // Two groups of four warps handle alternate tiles (group g: it = g, g + 2, ...), so that two tiles' worth of
    // barrier-wait / TMEM-load / store latency is in flight.  Within a group warp e covers TMEM lanes 32*(e & 3)..
    // (hardware rule: lane quadrant = warp index % 4) and all n_slab columns of its 32 pixels.
    const int quad = warp & 3, eg = warp >> 2;
    const int m = quad * 32 + lane;
    if (static_cast<int>(threadIdx.x) < p.n_slab) {
      s_bias[threadIdx.x] = bias_r;
      s_alpha[threadIdx.x] = alpha_r;
    }
    named_bar_sync(5, kEpiWarps * 32);
    const int lyA = m / p.P, lxA = m % p.P;                    // this lane's pixel inside a tile of geometry A / B
    const int lyB = m / max(p.P2, 1), lxB = m % max(p.P2, 1);
    const int sub_y = (p.up == 2) ? (slab >> 1) : 0, sub_x = (p.up == 2) ? (slab & 1) : 0;
    const int ch_base = (p.up == 2) ? 0 : slab * p.n_slab;  // first channel of this slab in the output slice
    const int cps2 = 2 * tile_step;
    const int tile0 = tile_first + eg * tile_step;
    int ar = eg % nw, ac = eg / nw;  // it % nw and it / nw, advanced incrementally (it += 2)
    if (!chain_dep) grid_dep_wait();  // the residual / carry may be produced by the previous layer
    const uint32_t chain_base = (chain_pub || chain_dep) ? ld_acquire_gpu_u32(p.chain) : 0u;
    const uint32_t* const dep_flags = p.chain + p.chain_dep_off;
    const uint32_t dep_target = chain_base + p.chain_dep_ord;
    bool dep_dead = false;
    // chained launch: tile q of the previous launch is complete (hence, transitively, everything older this tile reads)
    const uint32_t* const dep_cnt = p.chain + kChainCnt + p.chain_dep_ord;
    // warp-uniform: odd lanes watch the previous launch's tile counter (whole grid done: stop polling for good), even
    // lanes the flag of this tile
    auto dep_wait_own = [&](int tile_) {
      if (dep_dead || tile_ >= p.tiles_total) return;
      const uint32_t* f = dep_flags + (p.tile_rev ? p.tiles_total - 1 - tile_ : tile_);
      const bool cnt_lane = (lane & 1) != 0;
      for (uint32_t spins = 0;; ++spins) {
        const uint32_t v = ld_acquire_gpu_u32(cnt_lane ? dep_cnt : f);
        const bool hit = cnt_lane ? v >= static_cast<uint32_t>(p.tiles_total) : static_cast<int>(v - dep_target) >= 0;
        if (__any_sync(0xffffffffu, hit && cnt_lane)) {
          dep_dead = true;
          break;
        }
        if (__any_sync(0xffffffffu, hit)) break;
        if (spins > kChainSpinLimit) {
          if (lane == 0) atomicAdd(p.chain + 1, 1u);
          dep_dead = true;
          break;
        }
        __nanosleep(64);
      }
    };
    auto dep_wait_thread = [&](int tile_) {   // one thread (the carry fetch of a tile three iterations ahead)
      if (dep_dead || tile_ >= p.tiles_total) return;
      const uint32_t* f = dep_flags + (p.tile_rev ? p.tiles_total - 1 - tile_ : tile_);
      for (uint32_t spins = 0;; ++spins) {
        const uint32_t c = ld_acquire_gpu_u32(dep_cnt), v = ld_acquire_gpu_u32(f);
        if (c >= static_cast<uint32_t>(p.tiles_total) || static_cast<int>(v - dep_target) >= 0) break;
        if (spins > kChainSpinLimit) {
          atomicAdd(p.chain + 1, 1u);
          break;
        }
        __nanosleep(64);
      }
    };
    // publishing is deferred by one tile of the group: by the time the leader fences, those stores are a tile old and the
    // fence returns at once; the group's last tile is published after the loop
    int pub_pending = -1;
    auto chain_publish = [&](int tile_) {   // the group's leader thread
      __threadfence();
      *reinterpret_cast<volatile uint32_t*>(p.chain + p.chain_pub_off + loop_tile_q(tile_)) = chain_base + p.chain_pub_ord;
      atomicAdd(p.chain + kChainCnt + p.chain_pub_ord, 1u);
    };
    constexpr bool kCarryIn = (EPI >= 0) && ((EPI & 32) != 0);
    auto carry_fetch = [&](int tile_, int slot) {  // one elected thread of the group
      if (chain_dep) {
        dep_wait_thread(tile_);
        fence_proxy_async_all();
      }
      mbar_expect_tx(bar_cfull(slot), kCarryTileBytes);
      const uint8_t* csrc = reinterpret_cast<const uint8_t*>(p.carry_in) +
                            static_cast<size_t>(p.tile_rev ? p.tiles_total - 1 - tile_ : tile_) * kCarryTileBytes;
      if (kDev && (p.dbg_flags & 16))
        bulk_load(carry_smem + slot * kCarryTileBytes, csrc, kCarryTileBytes, bar_cfull(slot));
      else  // the carry is read exactly once: do not let it push live activations out of L2
        bulk_load_hint(carry_smem + slot * kCarryTileBytes, csrc, kCarryTileBytes, bar_cfull(slot), l2_policy_evict_first());
    };
    if (kCarryIn && quad == 0 && lane == 0) {
      for (int j = 0; j < kCarrySlots - 1; ++j)
        if (tile0 + j * cps2 < tile_end) carry_fetch(tile0 + j * cps2, kCarrySlots * eg + j);
    }
    int gi = 0;  // iterations of this group
    for (int it = eg, tile = tile0; tile < tile_end; it += 2, tile += cps2, ++gi) {
      const int acc = ar + nw * (ac & 1);
      const uint32_t par = (ac >> 1) & 1;
      const bool gb = loop_tile_is_b(tile);
      int n, ty0, tx0;
      tile_pos(loop_tile_q(tile), gb, n, ty0, tx0);
      if (chain_dep) dep_wait_own(tile);  // before the residual prefetch (warp-uniform: every lane polls the same word)
      const int ly = gb ? lyB : lyA, lx = gb ? lxB : lxA;
      const bool in_tile = gb ? (lx < p.Wb2 && ly < p.Hb2) : (lx < p.Wb && ly < p.Hb);
      const int y = ty0 + ly, x = tx0 + lx;
      const bool valid = in_tile && (y < p.H) && (x < p.W) && (tile < p.tiles_total);
      const size_t opix = (static_cast<size_t>(n) * p.OH + (y * p.up + sub_y)) * p.OW + (x * p.up + sub_x);
      const size_t rpix = (static_cast<size_t>(n) * p.H + y) * p.W + x;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * p.n_slab;
      if constexpr (EPI >= 0) {
        constexpr int ACT = EPI & 7;
        constexpr bool HAS_RES = (EPI & 8) != 0;
        constexpr bool CARRY_OUT = (EPI & 16) != 0, CARRY_IN = (EPI & 32) != 0;
        constexpr bool STAGED = (EPI & 64) != 0;  // stores (and residual loads) go through the per-warp transposition
        constexpr bool MASKED = (EPI & 128) != 0; // fused activation backward into mask_out
        constexpr bool RES_MASK = (EPI & 256) != 0;  // the "residual" is a LeakyReLU output: out = acc * lrelu'(res)
        // Global traffic goes through a per-warp transposition in shared memory: a lane owns one pixel (TMEM lane), but a
        // store instruction in which every lane writes 16 B of a different pixel costs 32 L1 wavefronts and 32 partial-
        // sector L2 requests.  Staged, eight (four) consecutive lanes cover the 128 (64) contiguous bytes of one pixel.
        // Rows are XOR-swizzled so that both the pixel-major and the transposed accesses are bank-conflict free.
        __nv_bfloat16* const out_base = reinterpret_cast<__nv_bfloat16*>(p.out) + p.out_coff + ch_base;
        __nv_bfloat16* const out2_base =
            p.out2 ? reinterpret_cast<__nv_bfloat16*>(p.out2) + p.out2_coff + ch_base : nullptr;
        uint4* const stg = reinterpret_cast<uint4*>(ctrl_gen + kSmemCtrlBytes + warp * p.epi_stage_bytes);
        const int opix_i = valid ? static_cast<int>(opix) : -1;  // host guarantees the pixel count fits 31 bits
        uint4 rq[8];  // HAS_RES kernels have n_slab <= 64 (host dispatch): one pass
        if (HAS_RES && !STAGED) {
          if (valid) {
            const uint4* rp = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.res) +
                                                             static_cast<size_t>(opix_i) * p.res_cstride + p.res_coff + ch_base);
#pragma unroll
            for (int q = 0; q < 8; ++q)
              if (8 * q < p.n_slab) rq[q] = ld_cg_v4(rp + q);
          }
        }
        // staged residual: the rows of one 64-channel pass are loaded line-wide (transposed) into rq, parked in the
        // staging buffer at the start of the pass, and the next pass is prefetched while this one is computed
        auto res_prefetch = [&](int pass) {
          const int cpr = min(64, p.n_slab - 64 * pass) >> 3;  // 16-byte chunks per staged row: 2, 4 or 8
          const int lg = (cpr == 8) ? 3 : (cpr == 4) ? 2 : 1;
          const __nv_bfloat16* const res_base =
              reinterpret_cast<const __nv_bfloat16*>(p.res) + p.res_coff + ch_base + 64 * pass;
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            if (k < cpr) {
              const int rr = (k << (5 - lg)) + (lane >> lg);
              const int px = __shfl_sync(0xffffffffu, opix_i, rr);  // up == 1: residual pixel == output pixel
              rq[k] = make_uint4(0u, 0u, 0u, 0u);
              if (px >= 0)
                rq[k] = ld_cg_v4(reinterpret_cast<const uint4*>(res_base + static_cast<size_t>(px) * p.res_cstride) + (lane & (cpr - 1)));
            }
          }
        };
        if (HAS_RES && STAGED) res_prefetch(0);
        uint4 cq[kCarryQ];  // CARRY_IN kernels have n_slab == 32 (host dispatch): the 32 partial sums of this pixel
        // carry layout: [tile][float4 index q][accumulator row m]: a warp instruction touches 512 contiguous bytes (4 lines).
        // Producer and consumer use the same tile decomposition (host), so (tile, m) names the same pixel in both.
        if (threadIdx.x == 0) { SSR_TRACE(2, 4 * it); }
        if (CARRY_IN) {
          // every warp of the group has read the other slot (previous iteration): refill it with the next tile's carry
          named_bar_sync(1 + eg, 128);
          if (quad == 0 && lane == 0 && tile + (kCarrySlots - 1) * cps2 < tile_end)
            carry_fetch(tile + (kCarrySlots - 1) * cps2, kCarrySlots * eg + ((gi + kCarrySlots - 1) % kCarrySlots));
          const int slot = kCarrySlots * eg + (gi % kCarrySlots);
          mbar_wait(bar_cfull(slot), (gi / kCarrySlots) & 1);
          const uint4* cp = reinterpret_cast<const uint4*>(ctrl_gen + kSmemCtrlBytes + kEpiWarps * p.epi_stage_bytes +
                                                          slot * kCarryTileBytes) + m;
#pragma unroll
          for (int q = 0; q < kCarryQ; ++q) cq[q] = cp[q * 128];
        }
        // the epilogue is ahead of the MMAs most of the time: probe with a back-off (eight spinning warps cost issue slots
        // and, under the board's power cap, clock)
        if (kDev && (p.dbg_flags & 1)) mbar_wait(bar_tfull(acc), par); else mbar_wait_sleep(bar_tfull(acc), par, 40);
        tc_fence_after();
        if (kChain && pub_pending >= 0 && pub_pending != tile) {
          if (quad == 0 && lane == 0) chain_publish(pub_pending);
          pub_pending = -1;
        }
        if (threadIdx.x == 0) { SSR_TRACE(2, 4 * it + 1); }
        uint32_t r32[32];
#pragma unroll
        for (int g = 0; g < kMaxNSlab / 16; ++g) {
          // leave the unrolled chain with ONE jump: every skipped 16-column block is a far branch of its own otherwise,
          // and each lands on a cold instruction-cache line (N = 32 layers skipped six of them per tile)
          if (16 * g >= p.n_slab) break;
          {
            const int pass = g >> 2;                                  // 64 output channels per staging pass
            const int n_out = CARRY_OUT ? p.n_act : p.n_slab;         // columns that leave as bf16 activations
            const int cpr = min(64, n_out - 64 * pass) >> 3;          // 16-byte chunks per staged row: 8, 4 or 2
            const int lg = (cpr == 8) ? 3 : (cpr == 4) ? 2 : 1;
            const int sw_own = (lane >> (3 - lg)) & (cpr - 1);
            uint4* const sp = stg + lane * cpr;
            const int c = (g & 3) * 2;
            if (HAS_RES && STAGED && (g & 3) == 0) {
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                if (k < cpr) {
                  const int rr = (k << (5 - lg)) + (lane >> lg);
                  const int sw = (rr >> (3 - lg)) & (cpr - 1);
                  stg[rr * cpr + ((lane & (cpr - 1)) ^ sw)] = rq[k];
                }
              }
              __syncwarp();
              if (64 * (pass + 1) < p.n_slab) res_prefetch(pass + 1);
            }
            // TMEM is read 32 columns at a time (one round trip and one wait per pair of 16-column groups)
            if ((g & 1) == 0) {
              if (16 * (g + 1) < p.n_slab) tmem_ld32(taddr + 16 * g, r32); else tmem_ld16_lo(taddr + 16 * g, r32);
            }
            float4 b4[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) b4[q] = *reinterpret_cast<const float4*>(s_bias + 16 * g + 4 * q);
            if ((g & 1) == 0) {
              tmem_ld_wait();
              if (16 * (g + 2) >= p.n_slab) {
                // all TMEM reads of this warp are done: hand the accumulator back before the global stores
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                  if (PAIR && crank == 1) mbar_arrive_cluster(leader_cta_addr(bar_tempty(acc))); else mbar_arrive(bar_tempty(acc));
                }
                if (threadIdx.x == 0) { SSR_TRACE(2, 4 * it + 2); }
              }
            }
            uint32_t* const r = r32 + 16 * (g & 1);
            if (CARRY_OUT && 16 * g >= p.n_act) {
              // partial sums of the next conv: raw fp32, no bias / activation (scratch rows write their own slots)
              if (PAIR && tile >= p.tiles_total) continue;  // the odd tile out of a CTA pair is a dummy
              const int tq = p.tile_rev ? p.tiles_total - 1 - tile : tile;  // the carry is indexed by the image-order tile
              uint4* cp = reinterpret_cast<uint4*>(p.carry_out) +
                          (static_cast<size_t>(tq) * kCarryQ + (16 * g - p.n_act) / (kCarryF16 ? 8 : 4)) * 128 + m;
              // the next launch reads the carry back: ask L2 to hold on to it
              const uint64_t keep = l2_policy_evict_last();
#pragma unroll
              for (int q = 0; q < (kCarryF16 ? 2 : 4); ++q) {
                uint4 cv;
                if (kCarryF16)
                  cv = make_uint4(pack_f16x2_sat(r[8 * q], r[8 * q + 1]), pack_f16x2_sat(r[8 * q + 2], r[8 * q + 3]),
                                  pack_f16x2_sat(r[8 * q + 4], r[8 * q + 5]), pack_f16x2_sat(r[8 * q + 6], r[8 * q + 7]));
                else
                  cv = make_uint4(r[4 * q], r[4 * q + 1], r[4 * q + 2], r[4 * q + 3]);
                if (kDev && (p.dbg_flags & 16)) cp[q * 128] = cv; else st_global_v4_hint(cp + q * 128, cv, keep);
              }
            } else {
              if (CARRY_IN && g < 2) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  float c0, c1, c2, c3;
                  if (kCarryF16) {
                    const uint4 c4 = cq[2 * (g & 1) + (q >> 1)];
                    unpack_f16x2((q & 1) ? c4.z : c4.x, c0, c1);
                    unpack_f16x2((q & 1) ? c4.w : c4.y, c2, c3);
                  } else {
                    const uint4 c4 = cq[(4 * (g & 1) + q) % kCarryQ];
                    c0 = __uint_as_float(c4.x); c1 = __uint_as_float(c4.y);
                    c2 = __uint_as_float(c4.z); c3 = __uint_as_float(c4.w);
                  }
                  r[4 * q] = __float_as_uint(__uint_as_float(r[4 * q]) + c0);
                  r[4 * q + 1] = __float_as_uint(__uint_as_float(r[4 * q + 1]) + c1);
                  r[4 * q + 2] = __float_as_uint(__uint_as_float(r[4 * q + 2]) + c2);
                  r[4 * q + 3] = __float_as_uint(__uint_as_float(r[4 * q + 3]) + c3);
                }
              }
              float v[16];
              bias_act16<ACT>(r, b4, s_alpha + 16 * g, p.act_alpha, v);
              if (HAS_RES && (STAGED || g < 4)) {
                const uint4 qa = STAGED ? sp[c ^ sw_own] : rq[2 * (g & 3)], qb = STAGED ? sp[(c + 1) ^ sw_own] : rq[2 * (g & 3) + 1];
                const uint32_t w[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
                if (RES_MASK) {
                  // `res` is the forward activation y of a LeakyReLU layer: out = acc * lrelu'(y)  (its backward)
#pragma unroll
                  for (int i = 0; i < 8; ++i) {
                    v[2 * i] *= bf16_lo(w[i]) > 0.f ? 1.f : p.act_alpha;
                    v[2 * i + 1] *= bf16_hi(w[i]) > 0.f ? 1.f : p.act_alpha;
                  }
                } else {
#pragma unroll
                  for (int i = 0; i < 8; ++i) {
                    v[2 * i] = bf16_lo(w[i]) + p.res_beta * v[2 * i];
                    v[2 * i + 1] = bf16_hi(w[i]) + p.res_beta * v[2 * i + 1];
                  }
                }
              }
              if (threadIdx.x == 0 && g < 2) { SSR_TRACE(2, 256 + 4 * it + 2 * g); }
              uint4 q0, q1;
              q0.x = pack_bf16x2(v[0], v[1]);   q0.y = pack_bf16x2(v[2], v[3]);
              q0.z = pack_bf16x2(v[4], v[5]);   q0.w = pack_bf16x2(v[6], v[7]);
              q1.x = pack_bf16x2(v[8], v[9]);   q1.y = pack_bf16x2(v[10], v[11]);
              q1.z = pack_bf16x2(v[12], v[13]); q1.w = pack_bf16x2(v[14], v[15]);
              if (MASKED) {
                const int mc = ch_base + 16 * g - p.mask_lo;   // column inside the masked slice (warp-uniform)
                if (mc >= 0 && mc < p.mask_n && valid) {
                  // dz = bf16(out) * act'(z): same values the stand-alone ssr_act_bwd_bf16 would read back
                  const uint4* zp = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.mask_z) +
                                                                  static_cast<size_t>(opix_i) * p.mask_cstride + p.mask_coff + mc);
                  const uint4 z0 = ld_cg_v4(zp), z1 = ld_cg_v4(zp + 1);
                  const uint32_t zw[8] = {z0.x, z0.y, z0.z, z0.w, z1.x, z1.y, z1.z, z1.w};
                  const uint32_t ow[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
                  uint32_t dw[8];
#pragma unroll
                  for (int i = 0; i < 8; ++i) {
                    const float lo = bf16_lo(ow[i]) * (bf16_lo(zw[i]) > 0.f ? 1.f : p.mask_alpha);
                    const float hi = bf16_hi(ow[i]) * (bf16_hi(zw[i]) > 0.f ? 1.f : p.mask_alpha);
                    dw[i] = pack_bf16x2(lo, hi);
                  }
                  uint4* mp = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.mask_out) +
                                                       static_cast<size_t>(opix_i) * p.mask_out_cstride + mc);
                  mp[0] = make_uint4(dw[0], dw[1], dw[2], dw[3]);
                  mp[1] = make_uint4(dw[4], dw[5], dw[6], dw[7]);
                }
              }
              if (STAGED) {
                sp[c ^ sw_own] = q0;
                sp[(c + 1) ^ sw_own] = q1;
              } else if (valid) {
                uint4* op = reinterpret_cast<uint4*>(out_base + static_cast<size_t>(opix_i) * p.out_cstride + 16 * g);
                op[0] = q0;
                op[1] = q1;
                if (out2_base != nullptr) {
                  uint4* op2 = reinterpret_cast<uint4*>(out2_base + static_cast<size_t>(opix_i) * p.out2_cstride + 16 * g);
                  op2[0] = q0;
                  op2[1] = q1;
                }
              }
            }
            if (threadIdx.x == 0 && g < 2) { SSR_TRACE(2, 256 + 4 * it + 2 * g + 1); }
            // end of a staging pass: write the 32 staged rows out, whole pixels per group of lanes
            const bool pass_end = ((g & 3) == 3) || (16 * (g + 1) >= n_out);
            if (STAGED && pass_end && !(CARRY_OUT && 16 * g >= p.n_act)) {
              __syncwarp();
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                if (k < cpr) {
                  const int rr = (k << (5 - lg)) + (lane >> lg);
                  const int cc = lane & (cpr - 1);
                  const int sw = (rr >> (3 - lg)) & (cpr - 1);
                  const int px = __shfl_sync(0xffffffffu, opix_i, rr);
                  const uint4 val = stg[rr * cpr + (cc ^ sw)];
                  if (px >= 0) {
                    reinterpret_cast<uint4*>(out_base + static_cast<size_t>(px) * p.out_cstride + 64 * pass)[cc] = val;
                    if (out2_base != nullptr)
                      reinterpret_cast<uint4*>(out2_base + static_cast<size_t>(px) * p.out2_cstride + 64 * pass)[cc] = val;
                  }
                }
              }
              __syncwarp();
            }
          }
        }
      } else {
        // generic path: any output dtype / channel count / residual dtype, runtime dispatch (edge layers only)
        mbar_wait(bar_tfull(acc), par);
        tc_fence_after();
        if (kChain && pub_pending >= 0 && pub_pending != tile) {
          if (quad == 0 && lane == 0) chain_publish(pub_pending);
          pub_pending = -1;
        }
        for (int c0 = 0; c0 < p.n_slab; c0 += 16) {
          uint32_t r[16];
          tmem_ld16(taddr + c0, r);
          float4 b4[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) b4[q] = *reinterpret_cast<const float4*>(s_bias + c0 + 4 * q);
          tmem_ld_wait();
          if (c0 + 16 >= p.n_slab) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if (PAIR && crank == 1) mbar_arrive_cluster(leader_cta_addr(bar_tempty(acc))); else mbar_arrive(bar_tempty(acc));
            }
          }
          if (!valid || c0 >= p.n_store) continue;
          float v[16];
          switch (p.act) {
            case SSR_ACT_LRELU: bias_act16<SSR_ACT_LRELU>(r, b4, s_alpha + c0, p.act_alpha, v); break;
            case SSR_ACT_PRELU: bias_act16<SSR_ACT_PRELU>(r, b4, s_alpha + c0, p.act_alpha, v); break;
            case SSR_ACT_TANH: bias_act16<SSR_ACT_TANH>(r, b4, s_alpha + c0, p.act_alpha, v); break;
            case SSR_ACT_RELU: bias_act16<SSR_ACT_RELU>(r, b4, s_alpha + c0, p.act_alpha, v); break;
            default: bias_act16<SSR_ACT_NONE>(r, b4, s_alpha + c0, p.act_alpha, v); break;
          }
          const int nst = min(16, p.n_store - c0);
          if (p.res_dtype == SSR_BF16) {
            const __nv_bfloat16* rp =
                reinterpret_cast<const __nv_bfloat16*>(p.res) + rpix * p.res_cstride + p.res_coff + ch_base + c0;
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (i < nst) v[i] = __bfloat162float(rp[i]) + p.res_beta * v[i];
          } else if (p.res_dtype == SSR_F32) {
            const float* rp = reinterpret_cast<const float*>(p.res) + rpix * p.res_cstride + p.res_coff + ch_base + c0;
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (i < nst) v[i] = rp[i] + p.res_beta * v[i];
          }
          if (p.out_dtype == SSR_BF16) {
            __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.out) + opix * p.out_cstride + p.out_coff + ch_base + c0;
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (i < nst) op[i] = __float2bfloat16_rn(v[i]);
          } else {
            float* op = reinterpret_cast<float*>(p.out) + opix * p.out_cstride + p.out_coff + ch_base + c0;
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (i < nst) op[i] = v[i];
          }
          if (p.out2 != nullptr) {
            __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.out2) + opix * p.out2_cstride + p.out2_coff + ch_base + c0;
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (i < nst) op[i] = __float2bfloat16_rn(v[i]);
          }
        }
      }
      if (threadIdx.x == 0) { SSR_TRACE(2, 4 * it + 3); }
      if (chain_pub) {
        // the group's stores of this tile are issued (bar.sync orders them before the leader's later fence)
        named_bar_sync(3 + eg, 128);
        if (tile < p.tiles_total) pub_pending = tile;
      }
      // next tile of this group: (ar, ac) += 2 in radix nw
      ar += 2;
      while (ar >= nw) {
        ar -= nw;
        ++ac;
      }
    }
    if (kChain && pub_pending >= 0 && quad == 0 && lane == 0) chain_publish(pub_pending);  // the group's last tile
    if (quad == 0 && lane == 0) SSR_TRACE_G(5 + eg);  // this group's last tile is stored
  }

  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();  // PAIR: the peer's smem / barriers stay alive until both are done
  if (warp == kMmaWarp0) {
    tc_fence_after();
    if (PAIR) tmem_dealloc2(tmem_base, p.tmem_cols); else tmem_dealloc(tmem_base, p.tmem_cols);
    if (lane == 0) SSR_TRACE_G(7);  // TMEM released: the CTA exits
  }
}

// ------------------------------------------------------------------------------------------------
// weight packing: HWIO fp32 -> [slab][tap][chunk][n_slab rows x 128 B] bf16, 128B-swizzled rows
// ------------------------------------------------------------------------------------------------
// mode 0: forward.  B rows = output channels, K = input channels, tap order as stored.
// mode 1: dgrad.  The packed conv maps dZ (cin_real = cout_fwd channels) to dX (cout = cin_fwd channels) with the
//         kernel rotated by 180 degrees: value = w[taps-1-t][co (= ci_fwd)][ci (= co_fwd)].
// mode 2: dgrad over an x-unrolled dZ (ssr_im2col_x_f32_to_bf16): taps = kh rows, K index = dx * cout_fwd + co_fwd.
//         fwd_kw / fwd_cout describe the forward kernel [kh, fwd_kw, cout(=cin_fwd), fwd_cout].
__device__ __forceinline__ void pack_weight_element(size_t i, const float* __restrict__ w, uint8_t* __restrict__ packed,
                                                    int taps, int cin_real, int nchunks, int cout, int n_slab, int mode,
                                                    int fwd_kw, int fwd_cout, int row_bytes) {
  // 32-bit index arithmetic: an image has < 2^31 elements, and the six 64-bit divisions per element were the whole cost
  // of the per-step re-pack (rc is 64 or 32: a shift)
  const int rc = row_bytes >> 1;  // channels per packed row: 64 (128B swizzle) or 32 (64B swizzle, cin <= 32)
  const uint32_t i32 = static_cast<uint32_t>(i);
  const int rc_shift = (rc == 64) ? 6 : 5;
  const int c = static_cast<int>(i32 & static_cast<uint32_t>(rc - 1));
  uint32_t q = i32 >> rc_shift;
  const int r = static_cast<int>(q % static_cast<uint32_t>(n_slab));
  q /= static_cast<uint32_t>(n_slab);
  const int ch = static_cast<int>(q % static_cast<uint32_t>(nchunks));
  q /= static_cast<uint32_t>(nchunks);
  const int t = static_cast<int>(q % static_cast<uint32_t>(taps));
  const int slab = static_cast<int>(q / static_cast<uint32_t>(taps));
  const int ci = ch * rc + c;
  const int co = slab * n_slab + r;
  float v = 0.f;
  if (ci < cin_real && co < cout) {
    if (mode == 0) {
      v = w[(static_cast<size_t>(t) * cin_real + ci) * cout + co];
    } else if (mode == 1) {
      v = w[(static_cast<size_t>(taps - 1 - t) * cout + co) * cin_real + ci];
    } else {
      const int dx = ci / fwd_cout, cf = ci - dx * fwd_cout;
      v = w[((static_cast<size_t>(taps - 1 - t) * fwd_kw + (fwd_kw - 1 - dx)) * cout + co) * fwd_cout + cf];
    }
  }
  // byte offset inside the [n_slab x 128B] tile, Swizzle<3,4,3>
  const int chunk16 = c >> 3;
  const size_t tile = ((static_cast<size_t>(slab) * taps + t) * nchunks + ch) * (static_cast<size_t>(n_slab) * row_bytes);
  const int sw = (row_bytes == 128) ? (r & 7) : ((r >> 1) & 3);  // XOR of the 16-byte chunk index: address bits [7,10) / [7,9)
  const size_t off = tile + (r >> 3) * (8 * row_bytes) + (r & 7) * row_bytes + ((chunk16 ^ sw) << 4) + (c & 7) * 2;
  *reinterpret_cast<__nv_bfloat16*>(packed + off) = __float2bfloat16_rn(v);
}

__global__ void pack_weights_kernel(const float* __restrict__ w, uint8_t* __restrict__ packed, int taps, int cin_real,
                                    int nchunks, int cout, int n_slab, int n_slabs, int mode, int fwd_kw, int fwd_cout,
                                    int row_bytes) {
  const size_t total = static_cast<size_t>(n_slabs) * taps * nchunks * n_slab * (row_bytes >> 1);
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x)
    pack_weight_element(i, w, packed, taps, cin_real, nchunks, cout, n_slab, mode, fwd_kw, fwd_cout, row_bytes);
}

// All weight images of a network in ONE launch (the training step re-packs ~700 images after every Adam update):
// blockIdx.y = table entry.
struct PackEntry {
  const float* w;
  uint8_t* packed;
  int taps, cin_real, nchunks, cout, n_slab, n_slabs, mode, fwd_kw, fwd_cout, row_bytes;
  // mode 3 (one source group of a composed dgrad image): destination K range [k0, k0 + kn), destination row r <-> input
  // channel row0 + r of the source conv [kh, kw, src_cin, src_cout], value scaled
  int k0, kn, row0, src_cin, src_cout;
  float scale;
  int pad_[12];
};
static_assert(sizeof(PackEntry) == SSR_PACK_ENTRY_BYTES, "PackEntry is the device table entry of ssr_conv2d_pack_batch");

// mode 3: value(t, r, ci) = scale * w_src[taps-1-t][row0 + r][ci - k0]  (180-degree rotated, transposed: the dgrad form),
// written at K index ci of the destination image.  Only the group's own K range is touched.
__device__ __forceinline__ void pack_slice_group(const PackEntry& e, size_t i) {
  const uint32_t i32 = static_cast<uint32_t>(i);
  const int kk = static_cast<int>(i32 % static_cast<uint32_t>(e.kn));
  const uint32_t q = i32 / static_cast<uint32_t>(e.kn);
  const int co = static_cast<int>(q % static_cast<uint32_t>(e.cout));
  const int t = static_cast<int>(q / static_cast<uint32_t>(e.cout));
  const int ci = e.k0 + kk;
  const int rc = e.row_bytes >> 1;
  const int ch = ci / rc, c = ci % rc;
  const int slab = co / e.n_slab, r = co % e.n_slab;
  const float v = e.scale * e.w[(static_cast<size_t>(e.taps - 1 - t) * e.src_cin + (e.row0 + co)) * e.src_cout + kk];
  const int chunk16 = c >> 3;
  const size_t tile = ((static_cast<size_t>(slab) * e.taps + t) * e.nchunks + ch) * (static_cast<size_t>(e.n_slab) * e.row_bytes);
  const int sw = (e.row_bytes == 128) ? (r & 7) : ((r >> 1) & 3);
  const size_t off = tile + (r >> 3) * (8 * e.row_bytes) + (r & 7) * e.row_bytes + ((chunk16 ^ sw) << 4) + (c & 7) * 2;
  *reinterpret_cast<__nv_bfloat16*>(e.packed + off) = __float2bfloat16_rn(v);
}

__global__ void pack_weights_batch_kernel(const PackEntry* __restrict__ table) {
  const PackEntry e = table[blockIdx.y];
  if (e.mode == 3) {
    const size_t total = static_cast<size_t>(e.taps) * e.cout * e.kn;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<size_t>(gridDim.x) * blockDim.x)
      pack_slice_group(e, i);
    return;
  }
  const size_t total = static_cast<size_t>(e.n_slabs) * e.taps * e.nchunks * e.n_slab * (e.row_bytes >> 1);
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x)
    pack_weight_element(i, e.w, e.packed, e.taps, e.cin_real, e.nchunks, e.cout, e.n_slab, e.mode, e.fwd_kw, e.fwd_cout,
                        e.row_bytes);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static int round_up(int a, int b) { return (a + b - 1) / b * b; }

bool g_force_rows128 = getenv("SSR_ROWS128") != nullptr;  // debug: SSR_ROWS128 in the environment keeps 128-byte rows everywhere

bool conv_plan(int kh, int kw, int cin, int cout, int up, ConvPlan* pl, int split) {
  if (kh < 1 || kw < 1 || kh > 9 || kw > 9 || !(kh & 1) || !(kw & 1)) return false;  // odd sizes up to 9 (SAME, stride 1)
  if (cin <= 0 || cin % 16 != 0 || cout <= 0) return false;
  pl->nchunks = (cin + 63) / 64;
  pl->ksteps_last = (cin - (pl->nchunks - 1) * 64) / 16;
  // <= 32 input channels: operand rows of 64 bytes (64B swizzle) - half the shared memory, L2 and DRAM traffic of a
  // 128-byte row that would be half zeros (growth-conv tails, the dgrad of the growth convs, the RGB edge conv)
  pl->row_bytes = (cin <= 32 && !g_force_rows128) ? 64 : 128;
  const int taps = kh * kw;
  if (up == 2) {
    if (cout % 64 != 0) return false;  // cout/4 must be a multiple of 16
    pl->n_slabs = 4;
    pl->n_slab = cout / 4;
  } else if (up == 1 && split == 2) {
    // caller-requested split of the N rows over a CTA pair (ssr_conv2d_pack_weights_pair / desc.w_split)
    if (cout % 32 != 0 || cout > kMaxNSlab || kh != 3 || kw != 3) return false;
    pl->n_slabs = 2;
    pl->n_slab = cout / 2;
  } else if (up == 1) {
    int n_slab = round_up(cout, 16);
    int n_slabs = 1;
    // weight slab must leave room for >= 3 pipeline stages
    const int w_max = kSmemBytes - 1024 - kSmemCtrlBytes - 3 * 24 * 1024;
    while ((n_slab > kMaxNSlab || static_cast<long long>(taps) * pl->nchunks * n_slab * pl->row_bytes > w_max) &&
           n_slab % 32 == 0) {
      n_slab /= 2;
      n_slabs *= 2;
    }
    pl->n_slab = n_slab;
    pl->n_slabs = n_slabs;
  } else {
    return false;
  }
  if (pl->n_slab > kMaxNSlab || pl->n_slab % 16 != 0) return false;
  pl->w_bytes = static_cast<long long>(taps) * pl->nchunks * pl->n_slab * pl->row_bytes;
  if (pl->w_bytes > kSmemBytes - 1024 - kSmemCtrlBytes - 2 * 24 * 1024) return false;
  return true;
}

// Pick the output tile (Wb x Hb, with (Hb-1)*P + Wb <= 128) that minimises a per-chunk time model:
// tiles x max(MMA issue time, L2->smem time of the halo box).
static void pick_tile(int kh, int kw, int H, int W, int n_slab, int max_stage_bytes, int* Wb_out, int* Hb_out,
                      int row_bytes = 128) {
  long long best = -1;
  int bw = 0, bh = 0;
  const int mma_cyc = std::max(n_slab / 2, 46);  // measured issue floor: ~46 clk per M=128 x K=16 MMA for N <= 64
  for (int Wb = 4; Wb <= 128; Wb += 2) {
    const int P = Wb + kw - 1;
    if (P > 256) break;
    const int Hb = (128 - Wb) / P + 1;
    if (Hb < 1 || Hb + kh - 1 > 256) continue;
    const int rows = std::max((Hb + kh - 1) * P, 128 + (kh - 1) * P + (kw - 1));
    if (rows * row_bytes > max_stage_bytes) continue;
    const long long tiles = static_cast<long long>((W + Wb - 1) / Wb) * ((H + Hb - 1) / Hb);
    const long long t_mma = static_cast<long long>(kh) * kw * 4 * mma_cyc;
    const long long t_load = static_cast<long long>(P) * (Hb + kh - 1) * row_bytes / 24;  // ~24 B/clk/SM from L2 with all SMs pulling
    const long long cost = tiles * std::max(t_mma, t_load);
    if (best < 0 || cost < best) {
      best = cost;
      bw = Wb;
      bh = Hb;
    }
  }
  *Wb_out = bw;
  *Hb_out = bh;
}

// Tile layout of one launch: geometry A everywhere, or A for the full tile rows and a second, wide geometry B for a ragged
// bottom strip when that saves tiles (see ConvKParams).  B is only used when neither tile order lets a CTA pair straddle
// the A / B boundary (both tile counts even), so that every launch over the same tensor - paired or not - tiles alike.
struct TileLayout {
  int Wb, Hb, tiles_x, tiles_y, tiles_yA;
  int Wb2, Hb2, tiles_x2, y2;   // Wb2 == 0: geometry B unused
  int tiles_total, nA_total;
};
static bool g_no_geo_b = getenv("SSR_NO_GEO_B") != nullptr;

static TileLayout tile_layout(int kh, int kw, int n, int H, int W, int Wb, int Hb, int row_bytes, int max_stage_bytes,
                              int sms, bool allow_b) {
  TileLayout t;
  memset(&t, 0, sizeof(t));
  t.Wb = Wb;
  t.Hb = Hb;
  t.tiles_x = (W + Wb - 1) / Wb;
  t.tiles_y = (H + Hb - 1) / Hb;
  t.tiles_yA = t.tiles_y;
  t.tiles_total = t.nA_total = t.tiles_x * t.tiles_y * n;
  const int rem = H - (t.tiles_y - 1) * Hb;  // rows of the last tile row
  if (!allow_b || g_no_geo_b || rem >= Hb || t.tiles_y < 2 || sms < 1) return t;
  // Strip tiles are `rem` rows high and as wide as 128 accumulator rows and the stage budget allow.  The strip geometry
  // is only used when it saves a whole round of tiles over the SMs; among the widths that do, the smallest stage wins.
  auto rounds = [&](int tiles) { return (tiles + sms - 1) / sms; };
  const int nA = t.tiles_x * (t.tiles_y - 1) * n;
  const int rowsA = std::max((Hb + kh - 1) * (Wb + kw - 1), 128 + (kh - 1) * (Wb + kw - 1) + (kw - 1));
  int best_tx2 = 0, best_rows = 0, best_rounds = rounds(t.tiles_total);  // must beat geometry A alone
  for (int tx2 = 1; tx2 < t.tiles_x; ++tx2) {
    const int wb = (W + tx2 - 1) / tx2, P = wb + kw - 1;
    if ((rem - 1) * P + wb > 128 || P > 256) continue;
    const int rows = std::max(rowsA, std::max((rem + kh - 1) * P, 128 + (kh - 1) * P + (kw - 1)));
    if (rows * row_bytes > max_stage_bytes) continue;
    const int nB = tx2 * n;
    if ((nA & 1) || (nB & 1)) continue;
    const int r = rounds(nA + nB);
    if (r < best_rounds || (best_tx2 && r == best_rounds && rows < best_rows)) {
      best_tx2 = tx2;
      best_rows = rows;
      best_rounds = r;
    }
  }
  if (!best_tx2) return t;
  t.tiles_yA = t.tiles_y - 1;
  t.Wb2 = (W + best_tx2 - 1) / best_tx2;
  t.Hb2 = rem;
  t.tiles_x2 = best_tx2;
  t.y2 = t.tiles_yA * Hb;
  t.nA_total = nA;
  t.tiles_total = nA + best_tx2 * n;
  return t;
}

// ------------------------------------------------------------------------------------------------
// Tile-level dependencies between consecutive conv launches (ssr_conv_chain_*)
// ------------------------------------------------------------------------------------------------
// A launch normally starts its loads after griddepcontrol.wait, i.e. after the WHOLE previous grid has completed and
// flushed: the drain of layer k (last tile's epilogue on every SM) and the fill of layer k+1 (first box, first MMAs)
// never overlap, ~7 us per launch at C2's size and most of a launch at training sizes.  Inside a chain the epilogue of
// layer k publishes a flag per pixel tile and layer k+1 (same tile decomposition) waits, per tile, for the 3x3
// neighbourhood of tiles its halo box reads.  Completion of a tile thereby implies completion of everything it depends
// on in all earlier chained launches (the dependency cone), which also orders every write-after-read on the ping-pong
// buffers.  Deadlock-free: a dependent grid only becomes resident after every CTA of its predecessor has triggered
// launch_dependents, i.e. is resident itself.
// Host state is per calling thread (plans are captured by one thread each; emulated ranks use several).
struct ChainState {
  uint32_t* buf = nullptr;
  int max_tiles = 0;
  uint32_t ord = 0;
  // the most recent publishing launch
  bool last_valid = false;
  cudaStream_t last_stream = nullptr;
  uint32_t last_ord = 0;
  int last_geo[14] = {0};
  long long chained = 0, published = 0;
};
static thread_local ChainState g_chain;

// one block: new epoch, per-launch tile counters back to zero (everything earlier in the stream has completed)
__global__ void chain_bump_kernel(uint32_t* chain) {
  if (threadIdx.x == 0) chain[0] += kChainStride;
  for (uint32_t i = threadIdx.x; i < kChainStride; i += blockDim.x) chain[kChainCnt + i] = 0u;
}

size_t conv_chain_bytes(int n, int h, int w) {
  const size_t tiles = static_cast<size_t>(n) * ((h + 3) / 4) * ((w + 3) / 4) + 1024;  // tiles are >= 4 pixels wide
  return (kChainHdr + 2 * tiles) * sizeof(uint32_t);
}
int conv_chain_begin(ssr_ctx* ctx, void* buf, size_t bytes, cudaStream_t stream) {
  ChainState& c = g_chain;
  c = ChainState();
  if (buf == nullptr) return SSR_OK;
  if (bytes < (kChainHdr + 2 * 64) * sizeof(uint32_t) || (reinterpret_cast<uintptr_t>(buf) & 15))
    return set_error(SSR_ERR_INVALID, "conv_chain_begin: buffer too small or misaligned");
  c.buf = static_cast<uint32_t*>(buf);
  c.max_tiles = static_cast<int>((bytes / sizeof(uint32_t) - kChainHdr) / 2);
  chain_bump_kernel<<<1, 1024, 0, stream>>>(c.buf);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(SSR_ERR_CUDA, "chain_bump launch: %s", cudaGetErrorString(e));
  ctx->launches++;
  return SSR_OK;
}
void conv_chain_end(long long* chained, long long* published) {
  if (chained) *chained = g_chain.chained;
  if (published) *published = g_chain.published;
  g_chain = ChainState();
}
void conv_chain_break() { g_chain.last_valid = false; }

size_t conv2d_carry_tiles(int n, int h, int w) {
  // upper bound (the strip geometry only ever removes tiles): the carry is indexed by tile, so a larger buffer is fine
  int Wb = 0, Hb = 0;
  pick_tile(3, 3, h, w, 32, 40 * 1024, &Wb, &Hb);
  if (Wb == 0) return 0;
  return static_cast<size_t>((w + Wb - 1) / Wb) * ((h + Hb - 1) / Hb) * n;
}

int conv2d_fwd_launch(ssr_ctx* ctx, const ssr_conv_desc* d, const void* x, const void* w_packed, const float* bias,
                      const float* alpha, const void* res, void* out, void* out2, cudaStream_t stream,
                      const float* carry_in, float* carry_out, int carry_out_cols, const ConvMask* mask) {
  ConvPlan pl;
  const int kh = d->ksize, kw = d->ksize_w > 0 ? d->ksize_w : d->ksize;
  if (d->w_split != 0 && d->w_split != 2) return set_error(SSR_ERR_INVALID, "conv2d: w_split must be 0 or 2");
  if (!conv_plan(kh, kw, d->cin, d->cout, d->up, &pl, d->w_split))
    return set_error(SSR_ERR_UNSUPPORTED, "conv2d: unsupported (ksize=%d cin=%d cout=%d up=%d)", d->ksize, d->cin,
                     d->cout, d->up);
  if (d->n <= 0 || d->h <= 0 || d->w <= 0) return set_error(SSR_ERR_INVALID, "conv2d: empty input");
  if (d->in_cstride % 8 != 0 || d->in_cstride < d->cin)
    return set_error(SSR_ERR_INVALID, "conv2d: in_cstride must be a multiple of 8 and >= cin");
  if (d->in_cvalid != 0 && (d->in_cvalid < d->cin || d->in_cvalid > d->in_cstride || d->in_cvalid % 8 != 0))
    return set_error(SSR_ERR_INVALID, "conv2d: in_cvalid must be a multiple of 8 in [cin, in_cstride]");
  if ((reinterpret_cast<uintptr_t>(x) & 15) != 0) return set_error(SSR_ERR_INVALID, "conv2d: x must be 16B aligned");

  ConvKParams p;
  memset(&p, 0, sizeof(p));
  // CTA pairs: a layer whose full-N weight slab was split in two halves (192 -> 64) runs as ONE N = 2 * n_slab MMA
  // over two SMs, each holding one half of the weight rows (debug bit6 disables it)
  // ... and so does every layer with an even number of slabs (cin >= 256: 16 or 32 rows per SM): slabs 2j, 2j + 1 pair up
  const bool pair = (pl.n_slabs >= 2 && pl.n_slabs % 2 == 0 && d->up == 1 && 2 * pl.n_slab <= kMaxNSlab && kh == 3 && kw == 3 &&
                     d->out_dtype == SSR_BF16 && d->cout == pl.n_slabs * pl.n_slab &&
                     (d->act == SSR_ACT_NONE || d->act == SSR_ACT_LRELU || d->act == SSR_ACT_RELU) &&
                     (res == nullptr || d->res_dtype == SSR_BF16) && !(ctx->debug_flags & 64) && ctx->sm_count >= 2 &&
                     ((pl.n_slabs == 2 && d->act != SSR_ACT_RELU) || !(ctx->debug_flags & 0x10000)));  // round-1 pairings only
  if (d->w_split == 2 && !pair)
    return set_error(SSR_ERR_UNSUPPORTED, "conv2d: w_split = 2 needs the CTA-pair form (3x3, bf16 out, act none / LeakyReLU)");
  const int n_mma = pair ? 2 * pl.n_slab : pl.n_slab;
  const int n_slabs = pair ? pl.n_slabs / 2 : pl.n_slabs;
  const int n_store = (d->up == 2) ? n_mma : std::min(n_mma, d->cout);  // cout < n_slab only when n_slabs == 1
  const int res_dtype = (res == nullptr) ? SSR_NONE : d->res_dtype;
  // specialised epilogue when the slice is bf16, 16-byte aligned and covers whole 32-column halves
  int epi = -1;
  const bool k33 = (kh == 3 && kw == 3);
  const long long out_pixels = static_cast<long long>(d->n) * d->h * d->up * d->w * d->up;
  if (k33 && d->out_dtype == SSR_BF16 && (res_dtype == SSR_NONE || (res_dtype == SSR_BF16 && d->up == 1)) &&
      n_store == n_mma && n_mma % 16 == 0 && d->act >= 0 && d->act <= 4 && out_pixels < (1ll << 31) && !(ctx->debug_flags & 32))
    epi = d->act + (res_dtype == SSR_BF16 ? 8 : 0);
  const bool res_mask = d->act == SSR_ACT_LRELU_MASK;
  if (res_mask) {
    // out = (acc + bias) * lrelu'(res): the direct bf16 epilogue with the residual as the mask (n <= 64, one slab)
    if (!(k33 && d->out_dtype == SSR_BF16 && res_dtype == SSR_BF16 && d->up == 1 && n_store == n_mma && n_mma % 16 == 0 &&
          n_mma <= 64 && n_slabs == 1 && !pair && out2 == nullptr && !carry_in && !carry_out && mask == nullptr &&
          out_pixels < (1ll << 31)))
      return set_error(SSR_ERR_UNSUPPORTED, "conv2d: SSR_ACT_LRELU_MASK needs a 3x3 conv, bf16 out and res, cout <= 64");
    epi = 8 + 256;
  }
  const int n_act = n_mma - (carry_out ? carry_out_cols : 0);
  // Stores of >= 64 channels per pixel go through per-warp transposition buffers (32 rows x min(columns, 64) bf16)
  // so that each store instruction writes whole 128-byte lines; narrower slices are stored directly.
  bool staged = epi >= 0 && !res_mask && n_act >= 64 && (n_act % 64 == 0 || n_act % 64 == 16 || n_act % 64 == 32) && !(ctx->debug_flags & 128);
  int epi_stage = staged ? 32 * 64 * 2 : 0;
  const int carry_ring = carry_in ? 2 * kCarrySlots * kCarryTileBytes : 0;  // carry tiles: kCarrySlots per epilogue group
  int smem_free = kSmemBytes - 1024 - kSmemCtrlBytes - static_cast<int>(pl.w_bytes) - kEpiWarps * epi_stage - carry_ring;
  if (staged && smem_free < 2 * 24 * 1024) {
    // big weight slab: no room for the transposition buffers next to two pipeline stages
    staged = false;
    epi_stage = 0;
    smem_free = kSmemBytes - 1024 - kSmemCtrlBytes - static_cast<int>(pl.w_bytes) - carry_ring;
  }
  if (!staged && res_dtype == SSR_BF16 && n_mma > 64) epi = -1;  // the direct epilogue prefetches <= 64 residual channels
  p.epi_stage_bytes = epi_stage;
  int Wb = 0, Hb = 0;
  const int stage_budget = smem_free / 2;
  if (carry_in != nullptr || carry_out != nullptr)
    pick_tile(kh, kw, d->h, d->w, 32, 40 * 1024, &Wb, &Hb);  // carry producer and consumer must tile identically
  else
    pick_tile(kh, kw, d->h, d->w, pl.n_slab, stage_budget, &Wb, &Hb, pl.row_bytes);
  if (Wb == 0) return set_error(SSR_ERR_UNSUPPORTED, "conv2d: no tile shape fits shared memory");
  if (ctx->force_wb > 0) {
    Wb = ctx->force_wb;
    Hb = (128 - Wb) / (Wb + kw - 1) + 1;
  }
  const bool carry_launch = carry_in != nullptr || carry_out != nullptr;
  // carry producer and consumer tile alike: the layout of a carry launch is decided as for 128-byte rows (the producer's)
  const TileLayout tl = tile_layout(kh, kw, d->n, d->h, d->w, Wb, Hb, carry_launch ? 128 : pl.row_bytes,
                                    carry_launch ? 40 * 1024 : stage_budget, ctx->sm_count,
                                    ctx->force_wb == 0 && !(ctx->debug_flags & 8) && d->up >= 1);
  p.kh = kh;
  p.kw = kw;
  p.Wb = Wb;
  p.Hb = Hb;
  p.P = Wb + kw - 1;
  const int R = Hb + kh - 1;
  int rows_needed = std::max(R * p.P, 128 + (kh - 1) * p.P + (kw - 1));
  p.row16 = pl.row_bytes / 16;
  p.box_bytes = R * p.P * pl.row_bytes;
  p.Wb2 = tl.Wb2;
  p.Hb2 = tl.Hb2;
  p.P2 = tl.Wb2 ? tl.Wb2 + kw - 1 : 0;
  p.tiles_x2 = tl.tiles_x2;
  p.y2 = tl.y2;
  p.tiles_yA = tl.tiles_yA;
  const int R2 = tl.Hb2 + kh - 1;
  if (tl.Wb2) {
    rows_needed = std::max(rows_needed, std::max(R2 * p.P2, 128 + (kh - 1) * p.P2 + (kw - 1)));
    p.box_bytes2 = R2 * p.P2 * pl.row_bytes;
  }
  p.stage_bytes = round_up(rows_needed * pl.row_bytes, 1024);
  p.w_bytes = static_cast<int>(pl.w_bytes);
  p.stages = std::min(kMaxStages, smem_free / p.stage_bytes);
  if (p.stages < 2) return set_error(SSR_ERR_UNSUPPORTED, "conv2d: not enough shared memory for 2 stages");

  // input tensor map: dims {C, W, H, N}
  // channel extent of the tensor map: whole 64-channel rows when the caller says they are readable (the extra channels
  // meet zero weights or are never touched by an MMA), else cin with TMA zero fill
  const int row_ch = pl.row_bytes / 2;  // channels per stage row
  const int c_ext = std::min(std::max(d->cin, d->in_cvalid), (d->cin + row_ch - 1) / row_ch * row_ch);
  cuuint64_t gdim[4] = {static_cast<cuuint64_t>(c_ext), static_cast<cuuint64_t>(d->w), static_cast<cuuint64_t>(d->h),
                        static_cast<cuuint64_t>(d->n)};
  cuuint64_t gstr[3] = {static_cast<cuuint64_t>(d->in_cstride) * 2, static_cast<cuuint64_t>(d->in_cstride) * 2 * d->w,
                        static_cast<cuuint64_t>(d->in_cstride) * 2 * d->w * d->h};
  cuuint32_t box[4] = {static_cast<cuuint32_t>(row_ch), static_cast<cuuint32_t>(p.P), static_cast<cuuint32_t>(R), 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult cr = ctx->encode_tiled(&p.tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), gdim, gstr, box,
                                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                  pl.row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                                  pl.row_bytes == 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_64B,
                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) return set_error(SSR_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", static_cast<int>(cr));
  if (tl.Wb2) {
    cuuint32_t box2[4] = {static_cast<cuuint32_t>(row_ch), static_cast<cuuint32_t>(p.P2), static_cast<cuuint32_t>(R2), 1};
    cr = ctx->encode_tiled(&p.tmap2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), gdim, gstr, box2, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE,
                           pl.row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                           pl.row_bytes == 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_64B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return set_error(SSR_ERR_CUDA, "cuTensorMapEncodeTiled (strip) failed (%d)", static_cast<int>(cr));
  }

  p.wpack = static_cast<const uint8_t*>(w_packed);
  p.bias = bias;
  p.alpha = (d->act == SSR_ACT_PRELU) ? alpha : nullptr;
  p.out = out;
  p.out2 = out2;
  p.res = (d->res_dtype == SSR_NONE) ? nullptr : res;
  p.out_dtype = d->out_dtype;
  p.res_dtype = (res == nullptr) ? SSR_NONE : d->res_dtype;
  p.out_cstride = d->out_cstride;
  p.out_coff = d->out_coff;
  p.out2_cstride = d->out2_cstride;
  p.out2_coff = d->out2_coff;
  p.res_cstride = d->res_cstride;
  p.res_coff = d->res_coff;
  p.up = d->up;
  p.OH = d->h * d->up;
  p.OW = d->w * d->up;
  p.n_img = d->n;
  p.H = d->h;
  p.W = d->w;
  p.nchunks = pl.nchunks;
  p.ksteps_last = pl.ksteps_last;
  p.n_slab = n_mma;
  p.w_rows = pl.n_slab;
  p.n_slabs = n_slabs;
  p.n_store = n_store;
  p.carry_in = carry_in;
  p.tile_rev = (d->tile_order != 0 && !(ctx->debug_flags & 1)) ? 1 : 0;
  p.carry_out = carry_out;
  p.n_act = n_act;
  p.tiles_x = tl.tiles_x;
  p.tiles_y = tl.tiles_y;
  p.tiles_total = tl.tiles_total;
  p.nA_total = tl.nA_total;
  p.ctas_per_slab = std::max(1, std::min(p.tiles_total, ctx->sm_count / n_slabs));
  if (pair) p.ctas_per_slab = std::max(1, std::min((p.tiles_total + 1) / 2, (ctx->sm_count / 2) / n_slabs));  // PAIRS per slab
  p.act = d->act;
  p.act_alpha = d->act_alpha;
  p.res_beta = d->res_beta;
  p.dbg_flags = ctx->debug_flags;
  p.trace = ctx->trace;
  if (ctx->trace != nullptr && ctx->trace_slots > 1) p.trace = ctx->trace + static_cast<size_t>(ctx->trace_next++ % ctx->trace_slots) * 1536;
  int nw = kMaxMmaWarps;
  while (nw > 1 && ((nw - 1) * pl.nchunks + 1 > p.stages || 2 * nw * n_mma > 512)) --nw;
  if (ctx->debug_flags & 4) nw = 1;  // debug bit2: single issuer
  p.mma_warps = nw;
  uint32_t cols = 32;
  while (cols < static_cast<uint32_t>(2 * nw * n_mma)) cols <<= 1;
  p.tmem_cols = cols;

  // vector stores need 16B-aligned channel slices
  if (d->out_dtype == SSR_BF16 && p.n_store % 16 == 0) {
    if (d->out_cstride % 8 != 0 || d->out_coff % 8 != 0 || (reinterpret_cast<uintptr_t>(out) & 15) != 0)
      return set_error(SSR_ERR_INVALID, "conv2d: bf16 out slice must be 16B aligned (cstride, coff multiples of 8)");
  }
  if (p.res_dtype == SSR_BF16 && p.n_store % 16 == 0) {
    if (d->res_cstride % 8 != 0 || d->res_coff % 8 != 0 || (reinterpret_cast<uintptr_t>(res) & 15) != 0)
      return set_error(SSR_ERR_INVALID, "conv2d: bf16 res slice must be 16B aligned");
  }

  const int smem = 1024 + p.w_bytes + p.stages * p.stage_bytes + kSmemCtrlBytes + kEpiWarps * epi_stage + carry_ring;
  void (*kern)(ConvKParams) = nullptr;
#define SSR_EPI_CASE(E, PAIRED) \
  case E: kern = staged ? conv_tc_kernel<3, (E) + 64, PAIRED> : conv_tc_kernel<3, E, PAIRED>; break;
  if (mask != nullptr) {
    // fused activation backward: the staged residual epilogue without activation, no CTA pairs
    if (!(epi == 8 && staged && !pair && !carry_in && !carry_out && out2 == nullptr))
      return set_error(SSR_ERR_UNSUPPORTED, "conv2d: masked output needs the staged bf16 residual epilogue (act none)");
    if (!mask->z || !mask->out || mask->n <= 0 || mask->n % 16 || mask->lo % 16 || mask->lo + mask->n > d->cout ||
        mask->z_cstride % 8 || mask->z_coff % 8 || mask->out_cstride % 8 || pl.n_slab % 16 ||
        (reinterpret_cast<uintptr_t>(mask->z) & 15) || (reinterpret_cast<uintptr_t>(mask->out) & 15))
      return set_error(SSR_ERR_INVALID, "conv2d: bad mask description (16-channel granularity, 16-byte alignment)");
    p.mask_z = mask->z;
    p.mask_out = mask->out;
    p.mask_cstride = mask->z_cstride;
    p.mask_coff = mask->z_coff;
    p.mask_out_cstride = mask->out_cstride;
    p.mask_lo = mask->lo;
    p.mask_n = mask->n;
    p.mask_alpha = mask->alpha;
    kern = conv_tc_kernel<3, 8 + 64 + 128, false>;
  } else if (carry_in != nullptr || carry_out != nullptr) {
    // growth-conv pairing: LeakyReLU, bf16 out, one slab; carry_out: N = 64 with 32 activated + 32 carried columns;
    // carry_in: N = 32
    const bool ok = epi == SSR_ACT_LRELU && n_slabs == 1 && !(carry_in && carry_out) && (!pair || carry_out) &&
                    (carry_out ? (n_mma == 64 && carry_out_cols == 32 && out2 == nullptr) : n_mma == 32);
    if (!ok) return set_error(SSR_ERR_UNSUPPORTED, "conv2d: unsupported carry configuration");
    kern = carry_out ? (pair ? conv_tc_kernel<3, 17, true> : conv_tc_kernel<3, 17, false>) : conv_tc_kernel<3, 33, false>;
  } else if (res_mask) {
    kern = conv_tc_kernel<3, 8 + 256, false>;
  } else if (pair) {
    switch (epi) {
      SSR_EPI_CASE(0, true)
      SSR_EPI_CASE(1, true)
      SSR_EPI_CASE(4, true)
      SSR_EPI_CASE(8, true)
      default: return set_error(SSR_ERR_UNSUPPORTED, "conv2d: pair mode epilogue %d", epi);
    }
  } else {
    switch (epi) {
      SSR_EPI_CASE(0, false)
      SSR_EPI_CASE(1, false)
      SSR_EPI_CASE(2, false)
      SSR_EPI_CASE(3, false)
      SSR_EPI_CASE(4, false)
      SSR_EPI_CASE(8, false)
      SSR_EPI_CASE(9, false)
      SSR_EPI_CASE(10, false)
      SSR_EPI_CASE(11, false)
      SSR_EPI_CASE(12, false)
      default:
        kern = k33 ? conv_tc_kernel<3, -1, false> : conv_tc_kernel<0, -1, false>;
    }
  }
#undef SSR_EPI_CASE
  // tile-level dependencies (ssr_conv_chain_*): publish when every tile is stored by exactly one CTA at the input
  // resolution; wait on flags instead of the grid when the previous launch of this stream published the same tiling.
  // Only the kernels a dense-block chain uses have a twin with the dependency code compiled in.
  {
    using KernFn = void (*)(ConvKParams);
    static const KernFn twins[][2] = {
        {conv_tc_kernel<3, 1, false>, conv_tc_kernel<3, 1, false, true>},
        {conv_tc_kernel<3, 17, false>, conv_tc_kernel<3, 17, false, true>},
        {conv_tc_kernel<3, 17, true>, conv_tc_kernel<3, 17, true, true>},
        {conv_tc_kernel<3, 33, false>, conv_tc_kernel<3, 33, false, true>},
        {conv_tc_kernel<3, 72, false>, conv_tc_kernel<3, 72, false, true>},
        {conv_tc_kernel<3, 72, true>, conv_tc_kernel<3, 72, true, true>},
        {conv_tc_kernel<3, 64, false>, conv_tc_kernel<3, 64, false, true>},
    };
    KernFn twin = nullptr;
    for (const auto& t : twins)
      if (t[0] == kern) twin = t[1];
    ChainState& cs = g_chain;
    const int geo[14] = {d->n, d->h, d->w, kh, kw, Wb, Hb, tl.tiles_x, tl.tiles_yA, tl.Wb2, tl.Hb2, tl.tiles_x2, tl.y2,
                         tl.tiles_total};
    const bool dep_ok = twin != nullptr && d->chain == 2 && cs.buf != nullptr && cs.last_valid && cs.last_stream == stream &&
                        mask == nullptr && memcmp(geo, cs.last_geo, sizeof(geo)) == 0 && k33;
    const bool pub_ok = twin != nullptr && d->chain >= 1 && cs.buf != nullptr && d->up == 1 && k33 && n_slabs == 1 &&
                        tl.tiles_total <= cs.max_tiles && cs.ord + 2 < kChainStride &&
                        (tl.Wb2 == 0 || tl.Wb2 / Wb + 4 <= 29);
    if (dep_ok) {
      p.chain = cs.buf;
      p.chain_dep_off = kChainHdr + static_cast<int>(cs.last_ord & 1) * cs.max_tiles;
      p.chain_dep_ord = cs.last_ord;
      cs.chained++;
    }
    if (pub_ok) {
      const uint32_t ord = ++cs.ord;
      p.chain = cs.buf;
      p.chain_pub_off = kChainHdr + static_cast<int>(ord & 1) * cs.max_tiles;
      p.chain_pub_ord = ord;
      cs.last_valid = true;
      cs.last_stream = stream;
      cs.last_ord = ord;
      memcpy(cs.last_geo, geo, sizeof(geo));
      cs.published++;
    } else if (cs.last_stream == stream) {
      cs.last_valid = false;
    }
    if (dep_ok || pub_ok) kern = twin;
  }
  if (int rc = opt_in_dynamic_smem(reinterpret_cast<const void*>(kern), kSmemBytes, "conv_tc_kernel")) return rc;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = pair ? dim3(2 * p.ctas_per_slab * n_slabs) : dim3(p.ctas_per_slab * n_slabs);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (!(ctx->debug_flags & 2)) {  // debug bit1: plain stream-ordered launches (no PDL)
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (pair) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 2;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, p);
  if (e != cudaSuccess) return set_error(SSR_ERR_CUDA, "conv_tc_kernel launch: %s", cudaGetErrorString(e));
  ctx->launches++;
  ctx->last_conv_tiles = p.tiles_total;
  return SSR_OK;
}

int conv2d_pack_launch(ssr_ctx* ctx, const float* w, int kh, int kw, int cin_real, int cin, int cout, int up,
                       void* packed, cudaStream_t stream, int mode, int fwd_kw, int fwd_cout, int split) {
  ConvPlan pl;
  if (!conv_plan(kh, kw, cin, cout, up, &pl, split))
    return set_error(SSR_ERR_UNSUPPORTED, "pack_weights: unsupported (k=%dx%d cin=%d cout=%d up=%d)", kh, kw, cin, cout, up);
  if (cin_real > cin || cin_real <= 0) return set_error(SSR_ERR_INVALID, "pack_weights: cin_real out of range");
  const size_t total = static_cast<size_t>(pl.n_slabs) * kh * kw * pl.nchunks * pl.n_slab * (pl.row_bytes / 2);
  const int block = 256;
  const int grid = static_cast<int>(std::min<size_t>((total + block - 1) / block, 148 * 8));
  pack_weights_kernel<<<grid, block, 0, stream>>>(w, static_cast<uint8_t*>(packed), kh * kw, cin_real, pl.nchunks, cout,
                                                  pl.n_slab, pl.n_slabs, mode, fwd_kw, fwd_cout, pl.row_bytes);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(SSR_ERR_CUDA, "pack_weights launch: %s", cudaGetErrorString(e));
  ctx->launches++;
  return SSR_OK;
}

int conv2d_pack_batch_entry(const ssr_pack_item* it, void* entry64) {
  int kh = it->kh, kw = it->kw, cin_real = it->cin_real, cin = it->cin, cout = it->cout, up = it->up;
  int fwd_kw = it->kw, fwd_cout = it->cout;
  if (it->mode == 1) {          // dgrad image: maps dZ (cout_fwd channels) to dX (cin_fwd channels)
    cin_real = it->cout;
    cin = (it->cout + 15) / 16 * 16;
    cout = it->cin_real;
    up = 1;
  } else if (it->mode == 2) {   // dgrad over the x-unrolled dZ
    cin_real = it->kw * it->cout;
    cin = (cin_real + 15) / 16 * 16;
    cout = it->cin_real;
    kw = 1;
    up = 1;
  } else if (it->mode == 3) {   // composed dgrad image: kh, kw, cin (K, multiple of 16), cout (rows) describe the DESTINATION conv
    cin_real = it->cin;
    up = 1;
    if (it->kn <= 0 || it->k0 < 0 || it->k0 + it->kn > it->cin || it->kn > it->src_cout || it->row0 < 0 ||
        it->row0 + it->cout > it->src_cin)
      return set_error(SSR_ERR_INVALID, "pack_batch: mode 3 group outside the destination / source conv");
  } else if (it->mode != 0) {
    return set_error(SSR_ERR_INVALID, "pack_batch: mode must be 0, 1, 2 or 3");
  }
  ConvPlan pl;
  if (!it->w_hwio || !it->packed || !conv_plan(kh, kw, cin, cout, up, &pl) || cin_real > cin || cin_real <= 0)
    return set_error(SSR_ERR_UNSUPPORTED, "pack_batch: unsupported item (k=%dx%d cin=%d cout=%d up=%d mode=%d)", it->kh,
                     it->kw, it->cin_real, it->cout, it->up, it->mode);
  PackEntry e;
  memset(&e, 0, sizeof(e));
  e.w = it->w_hwio;
  e.packed = static_cast<uint8_t*>(it->packed);
  e.taps = kh * kw;
  e.cin_real = cin_real;
  e.nchunks = pl.nchunks;
  e.cout = cout;
  e.n_slab = pl.n_slab;
  e.n_slabs = pl.n_slabs;
  e.mode = it->mode;
  e.fwd_kw = fwd_kw;
  e.fwd_cout = fwd_cout;
  e.row_bytes = pl.row_bytes;
  e.k0 = it->k0;
  e.kn = it->kn;
  e.row0 = it->row0;
  e.src_cin = it->src_cin;
  e.src_cout = it->src_cout;
  e.scale = it->scale;
  memcpy(entry64, &e, sizeof(e));
  return SSR_OK;
}

int conv2d_pack_batch_launch(ssr_ctx* ctx, const void* table_dev, int count, cudaStream_t stream) {
  pack_weights_batch_kernel<<<dim3(16, count), 256, 0, stream>>>(static_cast<const PackEntry*>(table_dev));
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(SSR_ERR_CUDA, "pack_weights_batch launch: %s", cudaGetErrorString(e));
  ctx->launches++;
  return SSR_OK;
}

}  // namespace ssr
