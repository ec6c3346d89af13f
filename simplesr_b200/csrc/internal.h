// internal.h — shared between the translation units of libssr_b200.so (not part of the ABI).
#pragma once
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstring>

#include "../../include/ssr_b200.h"

struct ssr_ctx {
  int device = 0;
  int sm_count = 0;
  int debug_flags = 0;
  int force_wb = 0;  // debug: force the conv output-tile width
  long long launches = 0;
  int last_conv_tiles = 0;  // debug: pixel tiles of the most recent conv launch
  long long* trace = nullptr;  // debug: device buffer of 3*512 int64 timestamps (conv kernel CTA 0)
  int trace_slots = 1;         // debug: consecutive conv launches write consecutive 1536-entry slots of `trace`
  int trace_next = 0;
  PFN_cuTensorMapEncodeTiled_v12000 encode_tiled = nullptr;
};

namespace ssr {

int set_error(int code, const char* fmt, ...);
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (device, kernel); thread-safe.  Returns SSR_OK or an error code.
int opt_in_dynamic_smem(const void* kernel, int bytes, const char* what);

// Launch with programmatic dependent launch allowed: the kernel may become resident while the previous kernel of the
// stream is still running.  It MUST execute griddepcontrol.wait before its first global memory access.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

struct ConvPlan {
  int nchunks;      // 64-channel K chunks
  int ksteps_last;  // K=16 steps in the last chunk
  int n_slab;       // UMMA N per CTA
  int n_slabs;      // output-channel slabs (separate CTAs)
  long long w_bytes;  // bytes of one packed weight slab
  int row_bytes;    // bytes per operand row: 128 (64 channels, 128B swizzle) or 64 (cin <= 32, 64B swizzle)
};
extern bool g_force_rows128;
bool conv_plan(int kh, int kw, int cin, int cout, int up, ConvPlan* pl, int split = 0);

struct ConvMask {   // fused activation backward of ssr_conv2d_fwd_mask
  const void* z;
  void* out;
  int z_cstride, z_coff, out_cstride, lo, n;
  float alpha;
};
int conv2d_fwd_launch(ssr_ctx* ctx, const ssr_conv_desc* d, const void* x, const void* w_packed, const float* bias,
                      const float* alpha, const void* res, void* out, void* out2, cudaStream_t stream,
                      const float* carry_in = nullptr, float* carry_out = nullptr, int carry_out_cols = 0,
                      const ConvMask* mask = nullptr);
int conv2d_pack_launch(ssr_ctx* ctx, const float* w, int kh, int kw, int cin_real, int cin, int cout, int up,
                       void* packed, cudaStream_t stream, int mode = 0, int fwd_kw = 0, int fwd_cout = 0, int split = 0);
int diag_mma_rate(ssr_ctx* ctx, int m, int n, int a_swz, int iters, int a_shift_rows, float* host_cycles_per_mma);

size_t conv2d_carry_tiles(int n, int h, int w);
size_t conv_chain_bytes(int n, int h, int w);
int conv_chain_begin(ssr_ctx* ctx, void* buf, size_t bytes, cudaStream_t stream);
void conv_chain_end(long long* chained, long long* published);
void conv_chain_break();
struct CommDev;
const CommDev* comm_dev(const ssr_comm* c);   // device view of an opened peer fabric (comm.cu), NULL if not opened
size_t comm_heap_bytes(const ssr_comm* c);
int conv2d_pack_batch_entry(const ssr_pack_item* it, void* entry64);
int conv2d_pack_batch_launch(ssr_ctx* ctx, const void* table_dev, int count, cudaStream_t stream);
int diag_mma_rate2(ssr_ctx* ctx, int n, int iters, float* host_cycles_per_mma);

}  // namespace ssr
