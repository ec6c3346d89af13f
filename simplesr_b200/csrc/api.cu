// api.cu — context, error reporting and memory helpers of the C ABI (include/ssr_b200.h).
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>

#include <mutex>
#include <set>
#include <vector>

#include "internal.h"

namespace ssr {

static thread_local char g_err[512] = "";

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

}  // namespace ssr

using namespace ssr;

extern "C" const char* ssr_last_error(void) { return g_err; }
extern "C" const char* ssr_version(void) { return "ssr_b200 0.1 (sm_100a)"; }

extern "C" int ssr_ctx_create(int device, ssr_ctx** out) {
  if (out == nullptr) return set_error(SSR_ERR_INVALID, "ctx_create: out is NULL");
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return set_error(SSR_ERR_CUDA, "ctx_create: no CUDA device (%s) - this library has no CPU fallback",
                     cudaGetErrorString(e));
  if (device < 0 || device >= count) return set_error(SSR_ERR_INVALID, "ctx_create: device %d out of range", device);
  e = cudaSetDevice(device);
  if (e != cudaSuccess) return set_error(SSR_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) return set_error(SSR_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
  if (prop.major != 10)
    return set_error(SSR_ERR_UNSUPPORTED, "ctx_create: device is sm_%d%d, this library is built for sm_100a only",
                     prop.major, prop.minor);
  ssr_ctx* ctx = new ssr_ctx();
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || fn == nullptr || qres != cudaDriverEntryPointSuccess) {
    delete ctx;
    return set_error(SSR_ERR_CUDA, "ctx_create: cuTensorMapEncodeTiled not available from the driver");
  }
  ctx->encode_tiled = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  *out = ctx;
  return SSR_OK;
}

extern "C" int ssr_ctx_destroy(ssr_ctx* ctx) {
  delete ctx;
  return SSR_OK;
}

extern "C" int ssr_ctx_sm_count(const ssr_ctx* ctx) { return ctx ? ctx->sm_count : 0; }

extern "C" int ssr_debug_set(ssr_ctx* ctx, int flags) {
  if (!ctx) return set_error(SSR_ERR_INVALID, "debug_set: ctx is NULL");
  ctx->debug_flags = (flags & 0xFF) | (flags & 0x7FFF0000);   // bits 16..: further flags
  ctx->force_wb = (flags >> 8) & 0xFF;  // bits 8..15: forced conv tile width (0 = automatic)
  return SSR_OK;
}

extern "C" int ssr_debug_last_conv_tiles(const ssr_ctx* ctx) { return ctx ? ctx->last_conv_tiles : 0; }

extern "C" int ssr_debug_trace(ssr_ctx* ctx, void* dev_int64_1536) {
  if (!ctx) return set_error(SSR_ERR_INVALID, "debug_trace: ctx is NULL");
  ctx->trace = static_cast<long long*>(dev_int64_1536);
  ctx->trace_slots = 1;
  ctx->trace_next = 0;
  return SSR_OK;
}
extern "C" int ssr_debug_trace_ring(ssr_ctx* ctx, void* dev_int64, int slots) {
  if (!ctx || slots < 1) return set_error(SSR_ERR_INVALID, "debug_trace_ring: bad argument");
  ctx->trace = static_cast<long long*>(dev_int64);
  ctx->trace_slots = slots;
  ctx->trace_next = 0;
  return SSR_OK;
}

#define SSR_CUDA(call, what)                                                                 \
  do {                                                                                       \
    cudaError_t e__ = (call);                                                                \
    if (e__ != cudaSuccess) return set_error(SSR_ERR_CUDA, what ": %s", cudaGetErrorString(e__)); \
  } while (0)

extern "C" int ssr_malloc(void** dptr, size_t bytes) {
  if (!dptr) return set_error(SSR_ERR_INVALID, "malloc: dptr is NULL");
  cudaError_t e = cudaMalloc(dptr, bytes ? bytes : 16);
  if (e != cudaSuccess) return set_error(SSR_ERR_NOMEM, "cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e));
  return SSR_OK;
}
extern "C" int ssr_free(void* dptr) {
  SSR_CUDA(cudaFree(dptr), "cudaFree");
  return SSR_OK;
}
extern "C" int ssr_memset(void* dptr, int value, size_t bytes, void* stream) {
  SSR_CUDA(cudaMemsetAsync(dptr, value, bytes, static_cast<cudaStream_t>(stream)), "cudaMemsetAsync");
  return SSR_OK;
}
extern "C" int ssr_memcpy_h2d(void* dst, const void* host_src, size_t bytes, void* stream) {
  SSR_CUDA(cudaMemcpyAsync(dst, host_src, bytes, cudaMemcpyHostToDevice, static_cast<cudaStream_t>(stream)),
           "cudaMemcpyAsync(h2d)");
  return SSR_OK;
}
extern "C" int ssr_memcpy_d2h(void* host_dst, const void* src, size_t bytes, void* stream) {
  SSR_CUDA(cudaMemcpyAsync(host_dst, src, bytes, cudaMemcpyDeviceToHost, static_cast<cudaStream_t>(stream)),
           "cudaMemcpyAsync(d2h)");
  return SSR_OK;
}
extern "C" int ssr_memcpy2d_d2h(void* host_dst, size_t dst_pitch, const void* src, size_t src_pitch, size_t width_bytes,
                               size_t rows, void* stream) {
  SSR_CUDA(cudaMemcpy2DAsync(host_dst, dst_pitch, src, src_pitch, width_bytes, rows, cudaMemcpyDeviceToHost,
                             static_cast<cudaStream_t>(stream)),
           "cudaMemcpy2DAsync(d2h)");
  return SSR_OK;
}
extern "C" int ssr_memcpy_d2d(void* dst, const void* src, size_t bytes, void* stream) {
  SSR_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream)),
           "cudaMemcpyAsync(d2d)");
  return SSR_OK;
}
extern "C" int ssr_stream_sync(void* stream) {
  SSR_CUDA(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)), "cudaStreamSynchronize");
  return SSR_OK;
}

extern "C" int ssr_host_alloc(void** hptr, size_t bytes) {
  if (!hptr) return set_error(SSR_ERR_INVALID, "host_alloc: hptr is NULL");
  cudaError_t e = cudaMallocHost(hptr, bytes ? bytes : 16);
  if (e != cudaSuccess) return set_error(SSR_ERR_NOMEM, "cudaMallocHost(%zu): %s", bytes, cudaGetErrorString(e));
  return SSR_OK;
}
extern "C" int ssr_host_free(void* hptr) {
  SSR_CUDA(cudaFreeHost(hptr), "cudaFreeHost");
  return SSR_OK;
}
extern "C" int ssr_stream_create(void** stream) {
  if (!stream) return set_error(SSR_ERR_INVALID, "stream_create: NULL");
  cudaStream_t s;
  SSR_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking), "cudaStreamCreate");
  *stream = s;
  return SSR_OK;
}
extern "C" int ssr_stream_destroy(void* stream) {
  SSR_CUDA(cudaStreamDestroy(static_cast<cudaStream_t>(stream)), "cudaStreamDestroy");
  return SSR_OK;
}
extern "C" int ssr_event_create(void** event) {
  if (!event) return set_error(SSR_ERR_INVALID, "event_create: NULL");
  cudaEvent_t e;
  SSR_CUDA(cudaEventCreate(&e), "cudaEventCreate");
  *event = e;
  return SSR_OK;
}
extern "C" int ssr_event_destroy(void* event) {
  SSR_CUDA(cudaEventDestroy(static_cast<cudaEvent_t>(event)), "cudaEventDestroy");
  return SSR_OK;
}
extern "C" int ssr_event_record(void* event, void* stream) {
  SSR_CUDA(cudaEventRecord(static_cast<cudaEvent_t>(event), static_cast<cudaStream_t>(stream)), "cudaEventRecord");
  return SSR_OK;
}
extern "C" int ssr_stream_wait_event(void* stream, void* event) {
  SSR_CUDA(cudaStreamWaitEvent(static_cast<cudaStream_t>(stream), static_cast<cudaEvent_t>(event), 0), "cudaStreamWaitEvent");
  return SSR_OK;
}
extern "C" int ssr_event_sync(void* event) {
  SSR_CUDA(cudaEventSynchronize(static_cast<cudaEvent_t>(event)), "cudaEventSynchronize");
  return SSR_OK;
}
extern "C" int ssr_event_elapsed_ms(void* start, void* stop, float* host_ms) {
  SSR_CUDA(cudaEventElapsedTime(host_ms, static_cast<cudaEvent_t>(start), static_cast<cudaEvent_t>(stop)),
           "cudaEventElapsedTime");
  return SSR_OK;
}
static thread_local int g_last_graph_kernels = -1;
extern "C" int ssr_graph_last_kernel_count(void) { return g_last_graph_kernels; }
extern "C" int ssr_graph_begin(void* stream) {
  SSR_CUDA(cudaStreamBeginCapture(static_cast<cudaStream_t>(stream), cudaStreamCaptureModeThreadLocal),
           "cudaStreamBeginCapture");
  return SSR_OK;
}
extern "C" int ssr_graph_end(void* stream, void** graph_exec) {
  if (!graph_exec) return set_error(SSR_ERR_INVALID, "graph_end: NULL");
  cudaGraph_t g = nullptr;
  SSR_CUDA(cudaStreamEndCapture(static_cast<cudaStream_t>(stream), &g), "cudaStreamEndCapture");
  {  // kernel nodes of the captured graph = kernels one launch of it runs (ssr_graph_last_kernel_count)
    size_t n = 0;
    g_last_graph_kernels = -1;
    if (cudaGraphGetNodes(g, nullptr, &n) == cudaSuccess && n > 0) {
      std::vector<cudaGraphNode_t> nodes(n);
      if (cudaGraphGetNodes(g, nodes.data(), &n) == cudaSuccess) {
        int k = 0;
        for (size_t i = 0; i < n; ++i) {
          cudaGraphNodeType t;
          if (cudaGraphNodeGetType(nodes[i], &t) == cudaSuccess && t == cudaGraphNodeTypeKernel) ++k;
        }
        g_last_graph_kernels = k;
      }
    }
  }
  cudaGraphExec_t ge = nullptr;
  cudaError_t e = cudaGraphInstantiate(&ge, g, 0);
  cudaGraphDestroy(g);
  if (e != cudaSuccess) return set_error(SSR_ERR_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(e));
  *graph_exec = ge;
  return SSR_OK;
}
extern "C" int ssr_graph_launch(void* graph_exec, void* stream) {
  SSR_CUDA(cudaGraphLaunch(static_cast<cudaGraphExec_t>(graph_exec), static_cast<cudaStream_t>(stream)),
           "cudaGraphLaunch");
  return SSR_OK;
}
extern "C" int ssr_graph_destroy(void* graph_exec) {
  SSR_CUDA(cudaGraphExecDestroy(static_cast<cudaGraphExec_t>(graph_exec)), "cudaGraphExecDestroy");
  return SSR_OK;
}
extern "C" int64_t ssr_ctx_launch_count(const ssr_ctx* ctx) { return ctx ? ctx->launches : 0; }

// ---------------------------------------------------------------- conv2d
extern "C" size_t ssr_conv2d_packed_bytes_hw(int kh, int kw, int cin, int cout, int up) {
  ConvPlan pl;
  if (!conv_plan(kh, kw, cin, cout, up, &pl)) {
    set_error(SSR_ERR_UNSUPPORTED, "conv2d_packed_bytes: unsupported (k=%dx%d cin=%d cout=%d up=%d)", kh, kw, cin, cout,
              up);
    return 0;
  }
  return static_cast<size_t>(pl.w_bytes) * pl.n_slabs;
}
extern "C" size_t ssr_conv2d_packed_bytes(int ksize, int cin, int cout, int up) {
  if (!(ksize == 1 || ksize == 3 || ksize == 9)) {
    set_error(SSR_ERR_UNSUPPORTED, "conv2d_packed_bytes: unsupported (ksize=%d cin=%d cout=%d up=%d)", ksize, cin, cout,
              up);
    return 0;
  }
  return ssr_conv2d_packed_bytes_hw(ksize, ksize, cin, cout, up);
}

extern "C" int ssr_conv2d_pack_weights_hw(ssr_ctx* ctx, const float* w_hwio, int kh, int kw, int cin_real, int cin,
                                          int cout, int up, void* packed, void* stream) {
  if (!ctx || !w_hwio || !packed) return set_error(SSR_ERR_INVALID, "conv2d_pack_weights: NULL argument");
  return conv2d_pack_launch(ctx, w_hwio, kh, kw, cin_real, cin, cout, up, packed, static_cast<cudaStream_t>(stream));
}
extern "C" int ssr_conv2d_pack_weights(ssr_ctx* ctx, const float* w_hwio, int ksize, int cin_real, int cin, int cout,
                                       int up, void* packed, void* stream) {
  return ssr_conv2d_pack_weights_hw(ctx, w_hwio, ksize, ksize, cin_real, cin, cout, up, packed, stream);
}

extern "C" int ssr_conv2d_pack_weights_pair(ssr_ctx* ctx, const float* w_hwio, int cin_real, int cin, int cout,
                                            void* packed, void* stream) {
  if (!ctx || !w_hwio || !packed) return set_error(SSR_ERR_INVALID, "conv2d_pack_weights_pair: NULL argument");
  return conv2d_pack_launch(ctx, w_hwio, 3, 3, cin_real, cin, cout, 1, packed, static_cast<cudaStream_t>(stream), 0, 0, 0, 2);
}

// dgrad: dX = conv(dZ, rot180(W)^T) runs through ssr_conv2d_fwd with this packed image (cin = cout_fwd, cout = cin_fwd)
extern "C" int ssr_conv2d_pack_weights_dgrad(ssr_ctx* ctx, const float* w_hwio, int kh, int kw, int cin_fwd, int cout_fwd,
                                             int unroll_x, void* packed, void* stream) {
  if (!ctx || !w_hwio || !packed) return set_error(SSR_ERR_INVALID, "conv2d_pack_weights_dgrad: NULL argument");
  if (unroll_x) {
    const int cin_real = kw * cout_fwd, cin = (cin_real + 15) / 16 * 16;
    return conv2d_pack_launch(ctx, w_hwio, kh, 1, cin_real, cin, cin_fwd, 1, packed, static_cast<cudaStream_t>(stream), 2,
                              kw, cout_fwd);
  }
  const int cin = (cout_fwd + 15) / 16 * 16;
  return conv2d_pack_launch(ctx, w_hwio, kh, kw, cout_fwd, cin, cin_fwd, 1, packed, static_cast<cudaStream_t>(stream), 1,
                            kw, cout_fwd);
}

extern "C" int ssr_conv2d_fwd(ssr_ctx* ctx, const ssr_conv_desc* d, const void* x, const void* w_packed,
                              const float* bias, const float* prelu_alpha, const void* res, void* out, void* out2,
                              void* stream) {
  if (!ctx || !d || !x || !w_packed || !out) return set_error(SSR_ERR_INVALID, "conv2d_fwd: NULL argument");
  if (d->act == SSR_ACT_PRELU && !prelu_alpha) return set_error(SSR_ERR_INVALID, "conv2d_fwd: PReLU needs alpha");
  if (d->res_dtype != SSR_NONE && !res) return set_error(SSR_ERR_INVALID, "conv2d_fwd: res_dtype set but res is NULL");
  return conv2d_fwd_launch(ctx, d, x, w_packed, bias, prelu_alpha, res, out, out2, static_cast<cudaStream_t>(stream));
}

extern "C" size_t ssr_conv_chain_bytes(int n, int h, int w) {
  if (n <= 0 || h <= 0 || w <= 0) return 0;
  return ssr::conv_chain_bytes(n, h, w);
}
extern "C" int ssr_conv_chain_begin(ssr_ctx* ctx, void* buf, size_t bytes, void* stream) {
  if (!ctx) return ssr::set_error(SSR_ERR_INVALID, "conv_chain_begin: ctx is NULL");
  return ssr::conv_chain_begin(ctx, buf, bytes, static_cast<cudaStream_t>(stream));
}
extern "C" int ssr_conv_chain_end(ssr_ctx* ctx, int64_t* published, int64_t* chained) {
  (void)ctx;
  long long c = 0, pb = 0;
  ssr::conv_chain_end(&c, &pb);
  if (published) *published = pb;
  if (chained) *chained = c;
  return SSR_OK;
}

extern "C" int ssr_diag_mma_rate(ssr_ctx* ctx, int n, int iters, int a_shift_rows, float* host_cycles_per_mma) {
  if (!ctx || !host_cycles_per_mma) return set_error(SSR_ERR_INVALID, "diag_mma_rate: NULL argument");
  return diag_mma_rate(ctx, 128, n, 2, iters, a_shift_rows, host_cycles_per_mma);
}

extern "C" int ssr_diag_mma_rate_ex(ssr_ctx* ctx, int m, int n, int a_swizzle, int iters, float* host_cycles_per_mma) {
  if (!ctx || !host_cycles_per_mma) return set_error(SSR_ERR_INVALID, "diag_mma_rate_ex: NULL argument");
  return diag_mma_rate(ctx, m, n, a_swizzle, iters, 0, host_cycles_per_mma);
}

extern "C" int ssr_diag_mma_rate_pair(ssr_ctx* ctx, int n, int iters, float* host_cycles_per_mma) {
  if (!ctx || !host_cycles_per_mma) return set_error(SSR_ERR_INVALID, "diag_mma_rate_pair: NULL argument");
  return diag_mma_rate2(ctx, n, iters, host_cycles_per_mma);
}

extern "C" int ssr_conv2d_fwd_carry(ssr_ctx* ctx, const ssr_conv_desc* d, const void* x, const void* w_packed,
                                    const float* bias, const void* res, void* out, const float* carry_in, float* carry_out,
                                    int carry_out_cols, void* stream) {
  if (!ctx || !d || !x || !w_packed || !out) return set_error(SSR_ERR_INVALID, "conv2d_fwd_carry: NULL argument");
  if (!carry_in && !carry_out) return set_error(SSR_ERR_INVALID, "conv2d_fwd_carry: no carry buffer");
  return conv2d_fwd_launch(ctx, d, x, w_packed, bias, nullptr, res, out, nullptr, static_cast<cudaStream_t>(stream),
                           carry_in, carry_out, carry_out_cols);
}

extern "C" int ssr_conv2d_fwd_mask(ssr_ctx* ctx, const ssr_conv_desc* d, const void* x, const void* w_packed,
                                   const float* bias, const void* res, void* out, const void* mask_z, int mask_z_cstride,
                                   int mask_z_coff, int mask_lo, int mask_n, float mask_alpha, void* mask_out,
                                   int mask_out_cstride, void* stream) {
  if (!ctx || !d || !x || !w_packed || !out || !res) return set_error(SSR_ERR_INVALID, "conv2d_fwd_mask: NULL argument");
  ConvMask m;
  m.z = mask_z;
  m.out = mask_out;
  m.z_cstride = mask_z_cstride;
  m.z_coff = mask_z_coff;
  m.out_cstride = mask_out_cstride;
  m.lo = mask_lo;
  m.n = mask_n;
  m.alpha = mask_alpha;
  return conv2d_fwd_launch(ctx, d, x, w_packed, bias, nullptr, res, out, nullptr, static_cast<cudaStream_t>(stream),
                           nullptr, nullptr, 0, &m);
}

// fp32 elements of a carry buffer for an [n,h,w] tensor (tile-major layout, see conv_tc.cu)
extern "C" size_t ssr_conv2d_carry_elems(ssr_ctx* ctx, int n, int h, int w) {
  (void)ctx;
  return conv2d_carry_tiles(n, h, w) * static_cast<size_t>(8 * 128 * 4);
}

// ---------------------------------------------------------------- batched weight packing
extern "C" int ssr_conv2d_pack_batch_prepare(ssr_ctx* ctx, const ssr_pack_item* items, int count, void* table_dev,
                                             void* stream) {
  if (!ctx || !items || !table_dev || count <= 0 || count > 65535)
    return set_error(SSR_ERR_INVALID, "pack_batch_prepare: bad argument (1 <= count <= 65535)");
  std::vector<uint8_t> host(static_cast<size_t>(count) * SSR_PACK_ENTRY_BYTES);
  for (int i = 0; i < count; ++i) {
    int rc = conv2d_pack_batch_entry(items + i, host.data() + static_cast<size_t>(i) * SSR_PACK_ENTRY_BYTES);
    if (rc != SSR_OK) return rc;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SSR_CUDA(cudaMemcpyAsync(table_dev, host.data(), host.size(), cudaMemcpyHostToDevice, st), "pack_batch_prepare: copy");
  SSR_CUDA(cudaStreamSynchronize(st), "pack_batch_prepare: sync");  // `host` dies with this call
  return SSR_OK;
}

extern "C" int ssr_conv2d_pack_batch(ssr_ctx* ctx, const void* table_dev, int count, void* stream) {
  if (!ctx || !table_dev || count <= 0 || count > 65535)
    return set_error(SSR_ERR_INVALID, "pack_batch: bad argument (1 <= count <= 65535)");
  return conv2d_pack_batch_launch(ctx, table_dev, count, static_cast<cudaStream_t>(stream));
}

// ---------------------------------------------------------------- shared-memory opt-in, once per (device, kernel)
namespace ssr {
int opt_in_dynamic_smem(const void* kernel, int bytes, const char* what) {
  static std::mutex mu;
  static std::set<std::pair<int, const void*>> done;
  std::lock_guard<std::mutex> lk(mu);
  int dev = 0;
  cudaGetDevice(&dev);
  if (done.count({dev, kernel})) return SSR_OK;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) return set_error(SSR_ERR_CUDA, "cudaFuncSetAttribute(%s): %s", what, cudaGetErrorString(e));
  done.insert({dev, kernel});
  return SSR_OK;
}
}  // namespace ssr
