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
typedef struct ssr_comm ssr_comm;

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
  SSR_ACT_RELU = 4,  /* ReLU()                    model_builder.py:265         */
  SSR_ACT_LRELU_MASK = 5 /* backward of LeakyReLU: y = (acc + bias) * (res > 0 ? 1 : act_alpha), res = the layer's
                            forward output (bf16); 3x3, cout <= 64.  Used by the training backward pass only */
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
int ssr_memcpy2d_d2h(void* host_dst, size_t dst_pitch, const void* src, size_t src_pitch, size_t width_bytes, size_t rows,
                    void* stream);   /* a rectangle of a pitched image (a rank's part of a tile row) */
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
int ssr_stream_wait_event(void* stream, void* event); /* fork / join between streams (also inside a graph capture) */
int ssr_event_sync(void* event);
int ssr_event_elapsed_ms(void* start, void* stop, float* host_ms);
int ssr_graph_begin(void* stream);                       /* cudaStreamBeginCapture (thread-local mode) */
int ssr_graph_end(void* stream, void** graph_exec);      /* end capture + instantiate */
int ssr_graph_last_kernel_count(void);                   /* kernel nodes of the graph this thread captured last (-1: unknown) */
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
  int32_t in_cvalid;     /* channels that may be READ per pixel starting at x (>= cin; 0 = cin): lets the TMA
                            box cover whole 64-channel rows instead of zero-filling a partial one            */
  int32_t w_split;       /* 0 = automatic; 2 = the weight image was packed by ssr_conv2d_pack_weights_pair: two
                            slabs of cout/2 rows, the conv runs on CTA pairs (cta_group::2, M = 256)          */
  int32_t tile_order;    /* 0 = pixel tiles first to last, 1 = last to first.  Alternating it from layer to layer makes
                            every layer read first what the previous one wrote last (still in L2)              */
  int32_t chain;         /* tile-level dependencies inside ssr_conv_chain_begin / _end (0 outside): 1 = publish a
                            completion flag per pixel tile; 2 = also start each tile as soon as the tiles of the PREVIOUS
                            launch of this stream that its halo reads are complete, instead of waiting for that whole
                            grid.  2 is only legal when the previous launch of the stream is a conv launched with
                            chain >= 1 that produced this launch's input (and everything else this launch reads or
                            overwrites was produced / last read by launches of the same chain or before its head);
                            it silently degrades to 1 when the two launches do not tile alike               */
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
/* 3x3, up = 1, cout a multiple of 32: the image split in two slabs of cout/2 rows for the CTA-pair form of the kernel
 * (desc.w_split = 2); same size as ssr_conv2d_packed_bytes(3, cin, cout, 1). */
int ssr_conv2d_pack_weights_pair(ssr_ctx* ctx, const float* w_hwio, int cin_real, int cin, int cout, void* packed,
                                 void* stream);
/* Conv2DBackpropInput (tape backward, sr_model.py:436-438): dX = conv(dZ, W rotated by 180 degrees, in/out swapped)
 * is ssr_conv2d_fwd with this packed image: cin = round16(cout_fwd) (cin_real = cout_fwd), cout = cin_fwd, same kh x kw.
 * unroll_x != 0: the dZ operand is x-unrolled (ssr_im2col_x_f32_to_bf16 with c = cout_fwd), the conv is kh x 1 over
 * round16(kw*cout_fwd) channels - used for the 9x9x64->3 output conv, whose plain dgrad weights exceed shared memory.
 * Size: ssr_conv2d_packed_bytes_hw(kh, unroll_x ? 1 : kw, <cin as above>, cin_fwd, 1). */
int ssr_conv2d_pack_weights_dgrad(ssr_ctx* ctx, const float* w_hwio, int kh, int kw, int cin_fwd, int cout_fwd,
                                  int unroll_x, void* packed, void* stream);
int ssr_conv2d_fwd(ssr_ctx* ctx, const ssr_conv_desc* d, const void* x, const void* w_packed, const float* bias,
                   const float* prelu_alpha, const void* res, void* out, void* out2, void* stream);

/* Tile-level dependencies between the consecutive convolutions of a layer chain (the 345 dense-block convs of
 * build_enhanced_resnet, model_builder.py:328-365, are one such chain).  Between ssr_conv_chain_begin and _end the calling
 * thread's conv launches honour desc.chain (see ssr_conv_desc).  `buf` (ssr_conv_chain_bytes for the largest [n,h,w]
 * tensor of the chain, zero-initialised once by the caller) holds an epoch counter and two flag arrays; _begin enqueues
 * the epoch bump on `stream`, so a captured launch sequence can be replayed.  A wait that does not complete within ~1 s
 * gives up and is counted in word 1 of `buf` (results are then undefined; 0 in a healthy run).
 * _end reports how many launches published flags / ran with tile-level dependencies (either may be NULL). */
size_t ssr_conv_chain_bytes(int n, int h, int w);
int ssr_conv_chain_begin(ssr_ctx* ctx, void* buf, size_t bytes, void* stream);
int ssr_conv_chain_end(ssr_ctx* ctx, int64_t* published, int64_t* chained);

/* Dense-block growth convs in pairs (model_builder.py:333-338).  The tensor core needs as many clocks for a
 * 128 x 32 x 16 MMA as for 128 x 64 x 16 (operand reads from shared memory dominate), so conv k is launched with
 * cout = 64: columns [0,32) are conv k itself (bias, LeakyReLU, bf16 slice store as usual) and columns [32,64) are the
 * partial sums of conv k+1 over the SAME input channels, handed over through carry_out [pixel tile][32 sums x 128 rows]
 * (carry_out_cols = 32; w_packed holds the two kernels side by side).  Conv k+1 then only convolves the 32 new channels
 * and adds carry_in to its fp32 accumulator before bias + LeakyReLU.  Exactly one of carry_in / carry_out.
 * The carry is private to the two launches (opaque tile-major layout; the sums travel as fp16, round-to-nearest,
 * saturating - finer than the bf16 rounding of the stored activations; the buffer is sized by ssr_conv2d_carry_elems). */
/* All weight images of a network in one launch (the training step re-packs every image after the optimizer update).
 * `items` (host) use the FORWARD conv geometry of ssr_conv2d_pack_weights_hw; mode 0 = forward image, 1 = dgrad image,
 * 2 = dgrad image over the x-unrolled dZ (ssr_conv2d_pack_weights_dgrad with unroll_x).  prepare() turns them into a
 * device table (count * SSR_PACK_ENTRY_BYTES bytes at table_dev, synchronous); ssr_conv2d_pack_batch is the launch. */
#define SSR_PACK_ENTRY_BYTES 128
typedef struct ssr_pack_item {
  const float* w_hwio; /* device, fp32 HWIO master */
  void* packed;        /* device, destination image */
  int32_t kh, kw, cin_real, cin, cout, up;
  int32_t mode;
  int32_t reserved;
  /* mode 3 - one source group of a COMPOSED dgrad image (the backward pass of a dense block evaluated slice by slice,
   * simplesr_b200/training.py): kh, kw, cin (K), cout (rows) describe the destination conv; the group fills K range
   * [k0, k0 + kn) with scale * rot180(w_hwio)[row0 + r][k - k0], w_hwio being a conv [kh, kw, src_cin, src_cout]. */
  int32_t k0, kn, row0, src_cin, src_cout;
  float scale;
} ssr_pack_item;
int ssr_conv2d_pack_batch_prepare(ssr_ctx* ctx, const ssr_pack_item* items, int count, void* table_dev, void* stream);
int ssr_conv2d_pack_batch(ssr_ctx* ctx, const void* table_dev, int count, void* stream);
/* size of a carry buffer for an [n,h,w] tensor in 4-byte units (an upper bound; private layout shared by the pair) */
size_t ssr_conv2d_carry_elems(ssr_ctx* ctx, int n, int h, int w);
int ssr_conv2d_fwd_carry(ssr_ctx* ctx, const ssr_conv_desc* d, const void* x, const void* w_packed, const float* bias,
                         const void* res, void* out, const float* carry_in, float* carry_out, int carry_out_cols,
                         void* stream);

/* ssr_conv2d_fwd with a bf16 residual and no activation (the dgrad form `out = res + beta * conv`), plus a fused
 * activation backward: output channels [mask_lo, mask_lo + mask_n) - after rounding to bf16 - times
 * (mask_z[pixel, mask_z_coff + c] > 0 ? 1 : mask_alpha) are also written to mask_out [pixels, mask_out_cstride] (bf16).
 * In the backward pass of a dense block (model_builder.py:328-341) the dgrad of conv k+1 completes the gradient of conv
 * k's output slice; this hands the next dgrad / wgrad its dZ without a separate LeakyReLU-backward launch.
 * mask_lo, mask_n: multiples of 16; cout >= 64 (staged epilogue). */
int ssr_conv2d_fwd_mask(ssr_ctx* ctx, const ssr_conv_desc* d, const void* x, const void* w_packed, const float* bias,
                        const void* res, void* out, const void* mask_z, int mask_z_cstride, int mask_z_coff, int mask_lo,
                        int mask_n, float mask_alpha, void* mask_out, int mask_out_cstride, void* stream);

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
/* General forms: rectangular patches (patch_h x patch_w; the reference's own tests pin (3,1), (1,3), (2,3), (3,2) with
 * overlap 0, tests/utils/image/test_image_utils.py:10,44-67) and BANDS for tile-sharded inference: `img` holds image
 * rows [src_row0, src_row0 + src_rows) only (it must cover every row the selected tiles read), `out` holds output rows
 * [out_row0, out_row0 + out_rows) only (pixels of the selected tiles outside the band are skipped). */
int ssr_segment_tiles_ex(const float* img, int h, int w, int c, int patch_h, int patch_w, int overlap, int tile_begin,
                         int tile_count, int src_row0, int src_rows, float* tiles, void* stream);
int ssr_stitch_tiles_ex(const float* tiles, int h, int w, int c, int patch_h, int patch_w, int overlap, int scale,
                        int tile_begin, int tile_count, int out_row0, int out_rows, float* out, void* stream);

/* ------------------------------------------------------------------ training step (bandwidth-bound parts) */
/* Pixel losses of the generator (mean_squared_error.py:57-58, mean_absolute_error.py:57-58: Keras global means) and the
 * PSNR metric of train_step (sr_model.py:453 -> metrics.py:4-15, tf.image.psnr per image), in one pass over hr / sr
 * (fp32 [n, per_image]).  out[0] = MSE, out[1] = MAE, out[2+i] = PSNR(image i, max_val).  If grad != NULL it receives
 * d(w_mse*MSE + w_mae*MAE)/d(sr).  Deterministic two-stage reduction; workspace from ssr_pixel_loss_workspace_bytes. */
size_t ssr_pixel_loss_workspace_bytes(int n);
int ssr_pixel_loss(const float* hr, const float* sr, int n, int64_t per_image, float w_mse, float w_mae, float max_val,
                   float* grad, void* workspace, float* out, void* stream);
/* Keras Adam.apply_gradients over ONE flat fp32 buffer holding every variable (sr_model.py:439-441):
 * m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2; p -= lr_t m / (sqrt(v) + eps), with g = grad * grad_scale and
 * lr_t = lr sqrt(1-b2^t)/(1-b1^t) computed by the caller. */
/* VGGLoss' optional total-variation term (vgg_loss.py:166-169): out1[0] = weight * value_scale * sum over the batch of
 * tf.image.total_variation(x) (x fp32 [n,h,w,c]; value_scale = 127.5 for sr in [-1,1]); its gradient is ADDED to grad
 * (may be NULL).  Deterministic two-stage sum. */
size_t ssr_total_variation_workspace_bytes(void);
int ssr_total_variation(const float* x, int n, int h, int w, int c, float value_scale, float weight, float* grad,
                        void* workspace, float* out1, void* stream);
int ssr_adam_step(float* param, const float* grad, float* m, float* v, int64_t count, float lr_t, float beta1,
                  float beta2, float eps, float grad_scale, void* stream);
/* out[c] (+)= scale * sum_p x[p, x_coff + c] * (z ? min(0, z[p, z_coff + c]) : 1): BiasAddGrad, and the PReLU slope
 * gradient d alpha_c = sum dy * min(0, x) (model_builder.py:118,281,314).  bf16 slices, fixed summation order. */
size_t ssr_channel_sum_workspace_bytes(int c);
int ssr_channel_sum_bf16(const void* x, int x_cstride, int x_coff, const void* z, int z_cstride, int z_coff,
                         int64_t pixels, int c, float scale, int accumulate, void* workspace, float* out, void* stream);
/* Conv2DBackpropFilter (tape backward of model_builder.py:285-293, sr_model.py:436-438) as a split-K tcgen05 GEMM:
 *   dw[dy,dx,ci,co] (HWIO fp32) (+)= scale * sum_{n,y,x} x[n, y+dy-kh/2, x+dx-kw/2, x_coff+ci] * dz[n, y, x, dz_coff+co]
 * x, dz: bf16 NHWC slices of [n,h,w,*] buffers (stride 1, SAME).  Deterministic (fixed-order reduction of per-CTA partials
 * held in `workspace`, size from ssr_conv2d_wgrad_workspace_bytes). */
size_t ssr_conv2d_wgrad_workspace_bytes(ssr_ctx* ctx, int h, int w, int cin, int cout, int kh, int kw);
int ssr_conv2d_wgrad(ssr_ctx* ctx, const void* x, int x_cstride, int x_coff, int cin_real, const void* dz, int dz_cstride,
                     int dz_coff, int cout, int n, int h, int w, int kh, int kw, float scale, int accumulate,
                     void* workspace, float* dw_hwio, void* stream);
/* Same, plus BiasAddGrad in the same launch: dbias[co] (+)= bias_scale * sum_{n,y,x} dz[n, y, x, dz_coff+co] (one more
 * accumulator whose A operand is a tile of ones).  kh*kw <= 14. */
int ssr_conv2d_wgrad_bias(ssr_ctx* ctx, const void* x, int x_cstride, int x_coff, int cin_real, const void* dz,
                          int dz_cstride, int dz_coff, int cout, int n, int h, int w, int kh, int kw, float scale,
                          int accumulate, void* workspace, float* dw_hwio, float* dbias, float bias_scale,
                          int bias_accumulate, void* stream);
/* Several weight gradients in ONE launch (+ one reduction): convolutions over tensors of the same n, h, w and kernel size,
 * e.g. the five convolutions of a dense block, which all read the block's 192-channel buffer (model_builder.py:328-341).
 * At training-patch sizes a single wgrad launch cannot fill the machine with K = pixels alone and spends most of its
 * time draining per-CTA partials; batched, the work units of all items share the CTAs.  count <= 8, at most 64 units
 * ((cin/64) x (cout/64) x tap groups summed over the items).  Item fields as in ssr_conv2d_wgrad_bias (dbias may be NULL). */
typedef struct ssr_wgrad_item {
  const void* x;
  int32_t x_cstride, x_coff, cin_real;
  const void* dz;
  int32_t dz_cstride, dz_coff, cout;
  float scale;
  int32_t accumulate;
  float* dw_hwio;
  float* dbias;
  float bias_scale;
  int32_t bias_accumulate;
} ssr_wgrad_item;
size_t ssr_conv2d_wgrad_multi_workspace_bytes(ssr_ctx* ctx, const ssr_wgrad_item* items, int count, int h, int w, int kh,
                                              int kw);
int ssr_conv2d_wgrad_multi(ssr_ctx* ctx, const ssr_wgrad_item* items, int count, int n, int h, int w, int kh, int kw,
                           void* workspace, void* stream);
/* y = z > 0 ? z : slope * z on bf16 slices (slope = alpha[c], or alpha_scalar when alpha == NULL): the training forward
 * stores the pre-activation z of every PReLU layer (model_builder.py:118,281,314) and activates it with this kernel */
int ssr_act_fwd_bf16(const void* z, int z_cstride, int z_coff, const float* alpha, float alpha_scalar, void* y,
                     int y_cstride, int y_coff, int64_t pixels, int c, void* stream);
/* dz = dy * (z > 0 ? 1 : slope): backward of PReLU (slope = alpha[c], z = forward pre-activation) or LeakyReLU
 * (alpha == NULL, slope = alpha_scalar, z = pre- or post-activation) on bf16 slices */
int ssr_act_bwd_bf16(const void* dy, int dy_cstride, int dy_coff, const void* z, int z_cstride, int z_coff,
                     const float* alpha, float alpha_scalar, void* dz, int dz_cstride, int dz_coff, int64_t pixels, int c,
                     void* stream);
/* gradient of tf.nn.depth_to_space(x, 2): y[n,h,w,(2i+j)*c + k] = x[n,2h+i,2w+j,k]; x: [n,2h,2w,c], bit-exact */
int ssr_space_to_depth2(const void* x, void* y, int n, int h, int w, int c, int elem_bytes, void* stream);
/* dz = g * (1 - y^2), fp32: backward through the tanh output activation (model_builder.py:93,133) */
int ssr_tanh_bwd_f32(const float* g, const float* y, float* dz, int64_t count, void* stream);
/* fp32 dense [pixels, c] -> bf16 slice (the loss gradient enters the backward pass in bf16) */
int ssr_f32_to_bf16_slice(const float* x, void* y, int y_cstride, int y_coff, int64_t pixels, int c, void* stream);

/* ------------------------------------------------------------------ VGG19 perceptual loss (vgg_loss.py:115-180) */
/* (x + 1) * 127.5 (vgg_loss.py:144-146) followed by keras vgg19.preprocess_input (:147-148: RGB -> BGR, minus
 * [103.939, 116.779, 123.68]).  x: fp32 [pixels,3] in [-1,1]; y: bf16 [pixels,16] (channels 3..15 zero). */
int ssr_vgg_preprocess(const float* x, void* y_bf16_c16, int64_t pixels, void* stream);
/* its backward: g_rgb[p,c] (+)= scale * 127.5 * dy_bgr[p,2-c] (fp32) */
int ssr_vgg_preprocess_bwd(const float* dy_bgr, float* g_rgb, int64_t pixels, float scale, int accumulate, void* stream);
/* MaxPooling2D(2,2) of the VGG copy (model_builder.py:267-269), bf16 NHWC, c % 8 == 0; odd trailing rows/cols dropped */
int ssr_maxpool2_bf16(const void* x, void* y, int n, int h, int w, int c, void* stream);
/* MaxPoolGrad: the gradient goes to the first maximum of each 2x2 window (row-major); h, w even */
int ssr_maxpool2_bwd_bf16(const void* x, const void* dy, void* dx, int n, int h, int w, int c, void* stream);
/* y += a * x (fp32): sums the gradients of several loss terms (generator.py:220-228) */
int ssr_axpy_f32(const float* x, float* y, float a, int64_t count, void* stream);

/* ------------------------------------------------------------------ discriminator step (model_builder.py:137-198) */
/* Conv2D(strides=2, padding="same") on even sizes = the stride-1 SAME conv sampled at odd positions (TF pads (0,1)):
 * y[n,i,j,:] = x[n,2i+1,2j+1,:] (x: [n,2oh,2ow,c]); ssr_zero_insert2 is its adjoint (gradient). Bit-exact copies. */
int ssr_subsample2(const void* x, void* y, int n, int oh, int ow, int c, int elem_bytes, void* stream);
int ssr_zero_insert2(const void* dy, void* dx, int n, int oh, int ow, int c, int elem_bytes, void* stream);
/* BatchNormalization(momentum, epsilon=1e-3) in training mode (model_builder.py:291-292): batch mean and biased variance
 * over (N,H,W) of x (bf16 dense [pixels,c]); istd = 1/sqrt(var+eps); the moving statistics (may be NULL) are updated
 * as m*momentum + batch*(1-momentum) with the unbiased variance. */
size_t ssr_bn_workspace_bytes(int c);
int ssr_bn_stats_bf16(const void* x, int64_t pixels, int c, float eps, float momentum, void* workspace, float* mean,
                      float* istd, float* moving_mean, float* moving_var, void* stream);
/* y = LeakyReLU_alpha(gamma * (x - mean) * istd + beta)   (model_builder.py:170-171, 295-306) */
int ssr_bn_lrelu_fwd_bf16(const void* x, const float* mean, const float* istd, const float* gamma, const float* beta,
                          float alpha, void* y, int64_t pixels, int c, void* stream);
/* backward of the pair: d' = dy * lrelu'(y); dbeta (+)= sum d'; dgamma (+)= sum d' xhat;
 * dz = gamma*istd*(d' - mean(d') - xhat*mean(d' xhat)).  sums_2c: scratch float[2c]; dgamma/dbeta may be NULL. */
int ssr_bn_lrelu_bwd_bf16(const void* x, const void* dy, const void* y, const float* mean, const float* istd,
                          const float* gamma, float alpha, int64_t pixels, int c, void* workspace, float* sums_2c,
                          float* dgamma, float* dbeta, int accumulate, void* dz, void* stream);
/* Dense (model_builder.py:191-193), fp32, batch <= 32: y = act(x W + b), W: [in, out] (Keras layout), optional pre-act */
size_t ssr_dense_workspace_bytes(int n, int out_features);
int ssr_dense_fwd_f32(const float* x, const float* w, const float* b, int n, int in_features, int out_features, int lrelu,
                      float alpha, void* workspace, float* pre_act, float* y, void* stream);
/* dx = dy W^T, dw (+)= x^T dy, db (+)= sum dy (any of dx/dw/db may be NULL) */
int ssr_dense_bwd_f32(const float* x, const float* w, const float* dy, int n, int in_features, int out_features, float* dx,
                      float* dw, float* db, int accumulate, void* stream);
int ssr_lrelu_bwd_f32(const float* dy, const float* h, float alpha, float* dh, int64_t count, void* stream);
/* RaAdversarialLoss (ra_adversarial_loss.py:59-69) and RaDiscriminatorLoss (ra_discriminator_loss.py:55-65) from the two
 * critics [n]: out2[0] = generator loss, out2[1] = discriminator loss (labels hr_label / sr_label); gradients w.r.t. the
 * critics: g_dsr (generator loss, SR critic), d_dsr / d_dhr (discriminator loss). */
int ssr_ragan_losses(const float* hr_critic, const float* sr_critic, int n, float hr_label, float sr_label, float* out2,
                     float* g_dsr, float* d_dsr, float* d_dhr, void* stream);

/* same with per-sample label arrays (label smoothing, discriminator.py:240-254; NULL = the scalar) and, with comm != NULL,
 * the relativistic means taken over the GLOBAL batch of a data-parallel step: every rank publishes its n_local critics and
 * labels in its heap (staging: 8 * n_local floats at stage_off, one barrier slot), gathers all ranks' values and writes
 * the global losses and the gradients of its own samples times world (the ranks' gradients are averaged afterwards). */
int ssr_ragan_losses_ex(ssr_comm* comm, int slot, size_t stage_off, const float* hr_critic, const float* sr_critic,
                        int n_local, float hr_label, float sr_label, const float* hr_labels, const float* sr_labels,
                        float* out2, float* g_dsr, float* d_dsr, float* d_dhr, void* stream);
/* The non-relativistic (SRGAN) critic, same arguments: the critics are the LOGITS of Dense(1, sigmoid)
 * (model_builder.py:194-196); out2[0] = AdversarialLoss = BCE(1, sigmoid(sr)) (adversarial_loss.py:58), out2[1] =
 * DiscriminatorLoss = BCE(sr_labels, sigmoid(sr)) + BCE(hr_labels, sigmoid(hr)) (discriminator_loss.py:56-59), Keras
 * BinaryCrossentropy on probabilities (clipped to [1e-7, 1 - 1e-7]); gradients w.r.t. the logits. */
int ssr_gan_losses_ex(ssr_comm* comm, int slot, size_t stage_off, const float* hr_critic, const float* sr_critic,
                      int n_local, float hr_label, float sr_label, const float* hr_labels, const float* sr_labels,
                      float* out2, float* g_dsr, float* d_dsr, float* d_dhr, void* stream);

/* ------------------------------------------------------------------ peer-memory fabric (data-parallel training)
 * New work required by BASELINE.json (the reference is single-device): one process per GPU; every rank
 * owns a HEAP (cudaMalloc) that all ranks map through CUDA IPC over NVLink / NVSwitch.  The collectives are kernels over
 * those mappings, synchronised by flags in each other's heap, and therefore part of the captured step graph.
 *   heap layout: [0, ssr_comm_data_offset()) barrier flags, then caller-managed data.  Buffers that take part in a
 *   collective sit at the SAME offset in every rank's heap (the host allocates them in the same order on all ranks).
 *   slots: every barrier site owns slot numbers < ssr_comm_max_slots() that no concurrently running kernel shares.
 * A wait that does not complete within the spin limit (default ~20 s) gives up and counts in ssr_comm_status instead of
 * hanging the GPU. */
#define SSR_COMM_HANDLE_BYTES 64
size_t ssr_comm_data_offset(void);
int ssr_comm_max_slots(void);
int ssr_comm_adam_slots(void);                       /* slots one ssr_comm_adam_step call site needs */
int ssr_comm_create(int device, int rank, int world, size_t heap_bytes, ssr_comm** out);
int ssr_comm_destroy(ssr_comm* comm);
void* ssr_comm_heap(ssr_comm* comm);                 /* local heap base (device pointer) */
size_t ssr_comm_heap_size(ssr_comm* comm);
int ssr_comm_ipc_handle(ssr_comm* comm, void* handle_out_64);          /* cudaIpcMemHandle_t of the local heap */
int ssr_comm_open_ipc(ssr_comm* comm, const void* handles_world_x_64); /* all ranks' handles, in rank order */
/* "Ranks" living in ONE process on ONE device (tests; world <= 4): comms[r] = the unopened rank r.  Kernels that wait
 * for each other must never be separate launches on one GPU, so an emulated group runs every collective as ONE
 * cooperative launch over all ranks: each rank calls the collective from its own host thread (eager launches, no graph
 * capture), the calls rendezvous on the host, and the last one launches the multi-rank kernel between the ranks' streams. */
int ssr_comm_open_local(ssr_comm* const* comms, int world);
int ssr_comm_set_spin_limit(ssr_comm* comm, double seconds);
int ssr_comm_status(ssr_comm* comm, unsigned long long* host_timeouts); /* synchronous: barrier waits that gave up */
int ssr_comm_barrier(ssr_comm* comm, int slot, void* stream);
/* out[i] = scale * sum over ranks of in[i] (count <= 65536; staging: 2 * count floats of heap at stage_off) */
int ssr_comm_allreduce_f32(ssr_comm* comm, int slot, size_t stage_off, const float* in, float* out, int count, float scale,
                           void* stream);
/* Device-resident optimizer clock (64 bytes): ssr_opt_prepare advances Keras' `iterations`, evaluates the learning-rate
 * schedule (none: base_lr; else tf.keras PiecewiseConstantDecay, examples/training/example_without_yaml.py:287-297) and
 * the bias-corrected Adam step size (sr_model.py:121-131), so that the whole update is capturable in the step graph. */
size_t ssr_opt_state_bytes(void);
int ssr_opt_state_set(void* state, int64_t iterations, void* stream);
int ssr_opt_state_get(const void* state, int64_t* host_iterations, float* host_lr, void* stream);
int ssr_opt_prepare(void* state, float base_lr, float beta1, float beta2, const int64_t* boundaries_dev,
                    const float* values_dev, int n_boundaries, void* stream);
/* ssr_adam_step with the step size read from the optimizer clock */
int ssr_adam_step_dev(float* param, const float* grad, float* m, float* v, int64_t count, const void* opt_state,
                      float beta1, float beta2, float eps, float grad_scale, void* stream);
/* Gradient reduce-scatter + Adam + parameter all-gather in one kernel over peer memory, elements [lo, hi) of the flat
 * buffers at grad_off / param_off of every heap (16-byte aligned, padded to a multiple of 4 floats; lo % 4 == 0): rank r
 * averages its shard of all ranks' gradients in rank order, updates m, v (local, full size) and the parameters, and
 * stores the new parameters into every rank's buffer.  Uses ssr_comm_adam_slots() slots from slot0. */
int ssr_comm_adam_step(ssr_comm* comm, int slot0, size_t grad_off, size_t param_off, float* m, float* v, int64_t lo,
                       int64_t hi, const void* opt_state, float beta1, float beta2, float eps, void* stream);
/* BatchNormalization over the global batch (sync-BN): as ssr_bn_stats_bf16 / ssr_bn_lrelu_bwd_bf16, the per-channel sums
 * exchanged through the heap (32 * c bytes at sums_off, (c + 31) / 32 slots from slot0).  dgamma / dbeta receive this
 * rank's sums (they are averaged with the other gradients), dz uses the global ones (model_builder.py:291-292). */
int ssr_bn_stats_bf16_dp(ssr_comm* comm, int slot0, size_t sums_off, const void* x, int64_t pixels_local, int c, float eps,
                         float momentum, void* workspace, float* mean, float* istd, float* moving_mean, float* moving_var,
                         void* stream);
int ssr_bn_lrelu_bwd_bf16_dp(ssr_comm* comm, int slot0, size_t sums_off, const void* x, const void* dy, const void* y,
                             const float* mean, const float* istd, const float* gamma, float alpha, int64_t pixels_local,
                             int c, void* workspace, float* sums_2c, float* dgamma, float* dbeta, int accumulate, void* dz,
                             void* stream);

/* ------------------------------------------------------------------ fp32-class precision (SRResNet, configs[0])
 * An fp32 activation is carried as two bf16 tensors (hi = bf16(a), lo = bf16(a - hi)) and a convolution runs as three
 * tcgen05 passes accumulated in fp32 through the conv kernel's fp32 residual input (a_hi*w_hi + a_lo*w_hi + a_hi*w_lo).
 * ssr_act_split_f32 is the elementwise tail between two such convolutions: z fp32 [n,h,w,c_out*up*up] ->
 * v = act(z) (+ res32), through depth_to_space(up) when up == 2 (model_builder.py:279), stored as fp32 (y32) and / or as the
 * (hi | lo) pair in channel slices [hi_off, +c_out) and [lo_off, +c_out) of a bf16 buffer with hl_cstride channels. */
int ssr_act_split_f32(const float* z, int n, int h, int w, int c_out, int up, int act, float act_alpha, const float* alpha,
                      const float* res32, float* y32, void* hi_lo_bf16, int hl_cstride, int hi_off, int lo_off,
                      void* stream);
int ssr_bf16_residual_f32(const float* x, float* y, int64_t count, void* stream);   /* y = x - float(bf16(x)) */

/* ------------------------------------------------------------------ device-side data preparation and eval metrics
 * (SURVEY.md §8f row 4).  fp32 NHWC images. */
/* tf.image.resize(x, [h/scale, w/scale], method="bicubic", antialias=...) as _prepare_img_pairs synthesises the LR batch
 * (data_pipeline.py:318-330): TensorFlow's ScaleAndTranslate with the Keys cubic kernel (a = -0.5), half-pixel centres,
 * support widened by the down-scaling factor when antialias, weights normalised per output pixel; columns then rows.
 * workspace: ssr_resize_workspace_bytes(n, h, w, c, scale). */
size_t ssr_resize_workspace_bytes(int n, int h, int w, int c, int scale);
int ssr_resize_bicubic(const float* x, float* y, int n, int h, int w, int c, int scale, int antialias, void* workspace,
                       void* stream);
/* Augmentations as one exact gather (image_transforms.py:50-80 crops, :157-173 rotate90, :320-345 flips): output image i is
 * the out_h x out_w window of source image src_index[i] (NULL: i) at (off_y[i], off_x[i]) (NULL: 0), rotated by
 * k = (mode >> 2) & 3 quarter turns counter-clockwise (tf.image.rot90; the window is taken before the rotation), then
 * flipped left-right (mode & 1, tf.image.flip_left_right) and / or up-down (mode & 2). */
int ssr_augment(const float* x, float* y, int n_out, int in_h, int in_w, int c, int out_h, int out_w, int mode,
                const int* src_index, const int* off_y, const int* off_x, void* stream);
/* metrics.psnr_on_y (metrics.py:18-44): tf.image.rgb_to_yuv luma of both RGB batches, then tf.image.psnr per image;
 * metrics.ssim (metrics.py:47-59): tf.image.ssim (11x11 Gaussian sigma 1.5, k1 0.01, k2 0.03, VALID windows) per image.
 * workspace: ssr_metric_workspace_bytes(n); out: n floats. */
size_t ssr_metric_workspace_bytes(int n);
int ssr_psnr_y(const float* a, const float* b, int n, int h, int w, float max_val, void* workspace, float* out, void* stream);
int ssr_ssim(const float* a, const float* b, int n, int h, int w, int c, float max_val, void* workspace, float* out,
             void* stream);

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
/* Development knobs of the conv kernel, for A/B measurements (0 = default).  Low byte = flags:
 *   1  tight epilogue wait and first-to-last tile order everywhere      2  no programmatic dependent launch
 *   4  one MMA-issuing warp                                             16 no L2 cache policies on the carry
 *   32 generic epilogue                                                 64 no CTA pairs
 *   128 no staged (line-wide) stores                                   8  no second tile geometry for a ragged bottom strip
 * bits 8..15: forced tile width in pixels (0 = tile picker).
 * 0x10000  CTA pairs only where round 1 used them (two slabs, no ReLU): the deep layers (cin >= 256) run one slab per CTA */
int ssr_debug_set(ssr_ctx* ctx, int flags);
/* pixel tiles of the most recent ssr_conv2d_* launch of this context (tests: which tile layout was chosen) */
int ssr_debug_last_conv_tiles(const ssr_ctx* ctx);
/* conv kernel timeline: CTA 0 writes clock64() stamps (3 roles x 512 events, int64) into the device buffer; NULL = off.
 * Only libraries built with -DSSR_DEV carry the stamps (they cost the hot kernels 2-3 % even when off); in the default
 * build the buffer stays untouched. */
int ssr_debug_trace(ssr_ctx* ctx, void* dev_int64_1536);
/* the same over `slots` consecutive conv launches (slot = launch % slots, 1536 int64 each); entries 500..507 of a slot
 * are globaltimer (ns) stamps of the launch boundaries: kernel entry, prologue done, previous grid complete, first
 * activation box, first MMA, last tile stored (epilogue group 0 / 1), TMEM released */
int ssr_debug_trace_ring(ssr_ctx* ctx, void* dev_int64, int slots);

#ifdef __cplusplus
}
#endif
#endif /* SSR_B200_H_ */
