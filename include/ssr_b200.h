/*
 * ssr_b200.h — C ABI of libssr_b200.so: the B200 (sm_100a) hot path of SimpleSR.
 *
 * The reference (bw0248/SimpleSR) is pure Python on TensorFlow 2.2 and has no FFI of its own; the
 * arithmetic on its hot path is a chain of TensorFlow ops.  Each entry point below names the
 * reference call site(s) (path:line under the reference repo) whose TF ops it replaces.  See
 * INTEGRATION.md for the reference-side binding (ctypes stub / tf.load_op_library shim).
 *
 * Conventions
 *  - plain C types only; every function returns 0 on success or a negative ssr_status;
 *    ssr_last_error() returns a thread-local message for the last failure on this thread.
 *  - the caller owns every buffer; all pointers are DEVICE pointers unless the name says "host".
 *  - every launch is asynchronous on the given cudaStream_t (passed as void*); no hidden syncs,
 *    no hidden allocations, no host<->device copies inside compute calls.
 *  - activations are NHWC.  "bf16" buffers may carry more channels per pixel than a layer reads or
 *    writes (cstride = channels per pixel of the buffer, coff = first channel of the slice): this is
 *    how the dense-block concatenation (model_builder.py:338) is never materialised.
 */
#ifndef SSR_B200_H_
#define SSR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ssr_ctx ssr_ctx;

enum ssr_status {
  SSR_OK = 0,
  SSR_ERR_INVALID = -1,   /* bad argument (maps to the reference's ValueError) */
  SSR_ERR_CUDA = -2,      /* CUDA runtime / driver failure */
  SSR_ERR_UNSUPPORTED = -3,
  SSR_ERR_NOMEM = -4
};

enum ssr_dtype { SSR_BF16 = 0, SSR_F32 = 1, SSR_NONE = -1 };

/* epilogue activation: applied to (acc + bias) before the residual */
enum ssr_act {
  SSR_ACT_NONE = 0,
  SSR_ACT_LRELU = 1, /* LeakyReLU(alpha)          model_builder.py:85,90,335  */
  SSR_ACT_PRELU = 2, /* PReLU(shared_axes=[1,2])  model_builder.py:118,281,314 */
  SSR_ACT_TANH = 3,  /* activation="tanh"         model_builder.py:93,133      */
  SSR_ACT_RELU = 4   /* ReLU()                    model_builder.py:265         */
};

/* ------------------------------------------------------------------ context / errors / memory */
const char* ssr_last_error(void);
const char* ssr_version(void);
int ssr_ctx_create(int device, ssr_ctx** out);
int ssr_ctx_destroy(ssr_ctx* ctx);
int ssr_ctx_sm_count(const ssr_ctx* ctx);

/* thin cudaMalloc / cudaMemcpyAsync wrappers so that a host without TensorFlow/PyTorch can drive the ABI */
int ssr_malloc(void** dptr, size_t bytes);
int ssr_free(void* dptr);
int ssr_memset(void* dptr, int value, size_t bytes, void* stream);
int ssr_memcpy_h2d(void* dst, const void* host_src, size_t bytes, void* stream);
int ssr_memcpy_d2h(void* host_dst, const void* src, size_t bytes, void* stream);
int ssr_memcpy_d2d(void* dst, const void* src, size_t bytes, void* stream);
int ssr_stream_sync(void* stream);
int ssr_host_alloc(void** hptr, size_t bytes);  /* pinned host memory (cudaMallocHost) */
int ssr_host_free(void* hptr);

/* streams, events (device-side timing on the launching stream) and CUDA graphs (launch-bound layer chains) */
int ssr_stream_create(void** stream);
int ssr_stream_destroy(void* stream);
int ssr_event_create(void** event);
int ssr_event_destroy(void* event);
int ssr_event_record(void* event, void* stream);
int ssr_event_sync(void* event);
int ssr_event_elapsed_ms(void* start, void* stop, float* host_ms);
int ssr_graph_begin(void* stream);                       /* cudaStreamBeginCapture (thread-local mode) */
int ssr_graph_end(void* stream, void** graph_exec);      /* end capture + instantiate */
int ssr_graph_launch(void* graph_exec, void* stream);
int ssr_graph_destroy(void* graph_exec);
/* number of kernels this context has launched (captured launches count once, at capture) */
int64_t ssr_ctx_launch_count(const ssr_ctx* ctx);

/* ------------------------------------------------------------------ conv2d (implicit GEMM, tcgen05)
 * Replaces tf.keras.layers.Conv2D(padding="same", strides=1) + BiasAdd and the elementwise ops the
 * reference chains behind it:  _build_conv_layer (model_builder.py:285-293), LeakyReLU (:85,:90,:335),
 * Lambda(x*0.2)+Add (:349-350,:363-364,:79), Concatenate (:338, via channel slices),
 * tf.nn.depth_to_space (:279, via up=2) and tanh (:91-94).
 *
 *   v   = act(bias[o] + sum_{i,j,c} x[n, h+i-p, w+j-p, c] * k[i,j,c,o])        (cross-correlation, SAME)
 *   y   = res_dtype==NONE ? v : res + res_beta * v
 *   up==1: out[n,h,w,out_coff+o] = y ;
 *   up==2: out[n,2h+i,2w+j,out_coff+c] = y for o = (2i+j)*(cout/4) + c   (TF NHWC depth_to_space, DCR)
 */
typedef struct ssr_conv_desc {
  int32_t n, h, w;       /* input batch, height, width                                            */
  int32_t cin;           /* input channels read (multiple of 16)                                  */
  int32_t in_cstride;    /* channels per pixel of the input buffer (bf16, multiple of 8)          */
  int32_t cout;          /* output channels of the convolution                                    */
  int32_t ksize;         /* 1, 3 or 9 (square, stride 1, SAME)                                    */
  int32_t act;           /* ssr_act                                                               */
  float act_alpha;       /* LeakyReLU slope                                                       */
  float res_beta;        /* y = res + res_beta * v                                                */
  int32_t up;            /* 1, or 2 = fused depth_to_space(block 2)                               */
  int32_t out_dtype;     /* ssr_dtype of out                                                      */
  int32_t out_cstride;   /* channels per pixel of the out buffer                                  */
  int32_t out_coff;      /* first channel written                                                 */
  int32_t res_dtype;     /* ssr_dtype of res, or SSR_NONE                                         */
  int32_t res_cstride;
  int32_t res_coff;
  int32_t out2_cstride;  /* optional second bf16 copy of y (out2 != NULL)                         */
  int32_t out2_coff;
  int32_t ksize_w;       /* kernel width if different from ksize (height); 0 = square             */
} ssr_conv_desc;

/* bytes of the packed (bf16, UMMA-ready, pre-swizzled) weight image for a layer */
size_t ssr_conv2d_packed_bytes(int ksize, int cin, int cout, int up);
/* HWIO fp32 [k,k,cin_real,cout] (Keras kernel layout, model_builder.py:287) -> packed image.
 * cin_real <= cin: input channels beyond cin_real are zero-filled (e.g. RGB 3 -> 16). */
int ssr_conv2d_pack_weights(ssr_ctx* ctx, const float* w_hwio, int ksize, int cin_real, int cin, int cout, int up,
                            void* packed, void* stream);
/* rectangular kernels (odd kh, kw <= 9): used for the 9x9x3 input convolution of SRResNet (model_builder.py:117),
 * which runs as a 9x1 convolution over an x-unrolled input (ssr_im2col_x_f32_to_bf16) */
size_t ssr_conv2d_packed_bytes_hw(int kh, int kw, int cin, int cout, int up);
int ssr_conv2d_pack_weights_hw(ssr_ctx* ctx, const float* w_hwio, int kh, int kw, int cin_real, int cin, int cout,
                               int up, void* packed, void* stream);
int ssr_conv2d_fwd(ssr_ctx* ctx, const ssr_conv_desc* d, const void* x, const void* w_packed, const float* bias,
                   const float* prelu_alpha, const void* res, void* out, void* out2, void* stream);

/* ------------------------------------------------------------------ bandwidth-bound kernels */
/* fp32 NHWC [n,h,w,c] -> bf16 NHWC with cpad >= c channels per pixel (extra channels zero).
 * Edge of the model: the reference feeds fp32 LR images in [0,1] (data_pipeline.py:318-330). */
int ssr_f32_to_bf16_pad(const float* x, void* y, int64_t pixels, int c, int cpad, void* stream);
/* fp32 NHWC [n,h,w,c] -> bf16 [n,h,w,cpad] with y[..., dx*c + ch] = x[n, h, w + dx - kw/2, ch] (0 outside the row):
 * unrolls the kernel width into channels so that a kw-wide convolution over few channels becomes a 1-wide one over
 * kw*c channels (cpad >= kw*c, multiple of 16; extra channels zero). */
int ssr_im2col_x_f32_to_bf16(const float* x, void* y, int n, int h, int w, int c, int kw, int cpad, void* stream);
/* bf16 slice -> fp32 dense [pixels, c] */
int ssr_bf16_to_f32(const void* x, int x_cstride, int x_coff, float* y, int64_t pixels, int c, void* stream);
/* out = a + beta * b on bf16 channel slices (Lambda*0.2 + Add, model_builder.py:363-364) */
int ssr_axpby_bf16(const void* a, int a_cstride, int a_coff, const void* b, int b_cstride, int b_coff, float beta,
                   void* out, int out_cstride, int out_coff, int64_t pixels, int c, void* stream);
/* tf.nn.depth_to_space(x, 2) NHWC, DCR order (model_builder.py:279): standalone, bit-exact.
 * elem_bytes in {2,4}; x: [n,h,w,4*c] -> y: [n,2h,2w,c] */
int ssr_depth_to_space2(const void* x, void* y, int n, int h, int w, int c, int elem_bytes, void* stream);

/* Overlapping-tile segmentation: image_utils._segment_with_overlap (image_utils.py:124-148).
 * img: fp32 [h,w,c]; tiles: fp32 [T, patch+2*overlap, patch+2*overlap, c], row-major tile order,
 * zero padding outside the image.  T = ceil(h/patch)*ceil(w/patch).  tile_begin/tile_count select a
 * contiguous range of tiles (multi-GPU sharding); tiles points at the first selected tile. */
int ssr_segment_tiles(const float* img, int h, int w, int c, int patch, int overlap, int tile_begin,
                      int tile_count, float* tiles, void* stream);
/* Stitch: image_utils.reconstruct_from_overlapping_patches + _reconstruct (image_utils.py:40-61,167-184).
 * tiles: fp32 [T, (patch+2*overlap)*scale, same, c] (range as above); out: fp32 [h*scale, w*scale, c]. */
int ssr_stitch_tiles(const float* tiles, int h, int w, int c, int patch, int overlap, int scale, int tile_begin,
                     int tile_count, float* out, void* stream);

/* ------------------------------------------------------------------ diagnostics */
/* tcgen05 issue-rate microbenchmark: iters back-to-back M=128 x N x K=16 MMAs per CTA on every SM, the A
 * operand starting a_shift_rows 128-byte rows into a swizzle-128B tile (0 = atom aligned).
 * Writes cycles per MMA to host_cycles_per_mma[0] (until the last MMA completed) and [1] (until the last one was
 * issued): the caller provides float[2]. Synchronous. */
int ssr_diag_mma_rate(ssr_ctx* ctx, int n, int iters, int a_shift_rows, float* host_cycles_per_mma);
/* same with M in {64,128} and the A-operand swizzle mode (2 = 128B, 4 = 64B, 6 = 32B rows) */
int ssr_diag_mma_rate_ex(ssr_ctx* ctx, int m, int n, int a_swizzle, int iters, float* host_cycles_per_mma);
/* same for CTA pairs: M = 256 x N x K = 16 (cta_group::2) issued by the leader CTA of every 2-CTA cluster */
int ssr_diag_mma_rate_pair(ssr_ctx* ctx, int n, int iters, float* host_cycles_per_mma);
/* debug knobs (0 = default): bit0 -> put (start>>7)&7 into the UMMA descriptor base_offset field */
int ssr_debug_set(ssr_ctx* ctx, int flags);
/* conv kernel timeline: CTA 0 writes clock64() stamps (3 roles x 512 events, int64) into the device buffer; NULL = off */
int ssr_debug_trace(ssr_ctx* ctx, void* dev_int64_1536);

#ifdef __cplusplus
}
#endif
#endif /* SSR_B200_H_ */
