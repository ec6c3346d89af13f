"""ctypes binding of libssr_b200.so (C ABI in include/ssr_b200.h).

This is the only place the shared library is loaded.  There is no CPU fallback: if the library is
missing the import of any compute entry point raises, and if no B200 is present ``Context()`` raises.
Host arrays are numpy; device memory is owned by :class:`DeviceBuffer` (cudaMalloc through the ABI),
so the product path needs neither TensorFlow nor PyTorch.
"""
import ctypes as C
import gc
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libssr_b200.so")

SSR_BF16, SSR_F32, SSR_NONE = 0, 1, -1
ACT_NONE, ACT_LRELU, ACT_PRELU, ACT_TANH, ACT_RELU, ACT_LRELU_MASK = 0, 1, 2, 3, 4, 5


class SsrError(RuntimeError):
    pass


class ConvDesc(C.Structure):
    """Mirror of ``ssr_conv_desc`` (include/ssr_b200.h)."""

    _fields_ = [
        ("n", C.c_int32), ("h", C.c_int32), ("w", C.c_int32),
        ("cin", C.c_int32), ("in_cstride", C.c_int32), ("cout", C.c_int32), ("ksize", C.c_int32),
        ("act", C.c_int32), ("act_alpha", C.c_float), ("res_beta", C.c_float), ("up", C.c_int32),
        ("out_dtype", C.c_int32), ("out_cstride", C.c_int32), ("out_coff", C.c_int32),
        ("res_dtype", C.c_int32), ("res_cstride", C.c_int32), ("res_coff", C.c_int32),
        ("out2_cstride", C.c_int32), ("out2_coff", C.c_int32), ("ksize_w", C.c_int32), ("in_cvalid", C.c_int32),
        ("w_split", C.c_int32), ("tile_order", C.c_int32), ("chain", C.c_int32),
    ]


_SIGNATURES = {
    "ssr_last_error": (C.c_char_p, []),
    "ssr_version": (C.c_char_p, []),
    "ssr_ctx_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "ssr_ctx_destroy": (C.c_int, [C.c_void_p]),
    "ssr_ctx_sm_count": (C.c_int, [C.c_void_p]),
    "ssr_malloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_size_t]),
    "ssr_free": (C.c_int, [C.c_void_p]),
    "ssr_memset": (C.c_int, [C.c_void_p, C.c_int, C.c_size_t, C.c_void_p]),
    "ssr_memcpy_h2d": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "ssr_memcpy_d2h": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "ssr_memcpy_d2d": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "ssr_stream_sync": (C.c_int, [C.c_void_p]),
    "ssr_host_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_size_t]),
    "ssr_host_free": (C.c_int, [C.c_void_p]),
    "ssr_stream_create": (C.c_int, [C.POINTER(C.c_void_p)]),
    "ssr_stream_destroy": (C.c_int, [C.c_void_p]),
    "ssr_event_create": (C.c_int, [C.POINTER(C.c_void_p)]),
    "ssr_event_destroy": (C.c_int, [C.c_void_p]),
    "ssr_event_record": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ssr_stream_wait_event": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ssr_event_sync": (C.c_int, [C.c_void_p]),
    "ssr_event_elapsed_ms": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_float)]),
    "ssr_graph_begin": (C.c_int, [C.c_void_p]),
    "ssr_graph_end": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "ssr_graph_launch": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ssr_graph_last_kernel_count": (C.c_int, []),
    "ssr_graph_destroy": (C.c_int, [C.c_void_p]),
    "ssr_ctx_launch_count": (C.c_int64, [C.c_void_p]),
    "ssr_conv2d_packed_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "ssr_conv2d_pack_weights": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                          C.c_void_p, C.c_void_p]),
    "ssr_conv2d_packed_bytes_hw": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "ssr_conv2d_pack_weights_hw": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                             C.c_int, C.c_void_p, C.c_void_p]),
    "ssr_im2col_x_f32_to_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                           C.c_int, C.c_void_p]),
    "ssr_conv2d_fwd": (C.c_int, [C.c_void_p, C.POINTER(ConvDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ssr_conv2d_fwd_mask": (C.c_int, [C.c_void_p, C.POINTER(ConvDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p,
                                      C.c_int, C.c_void_p]),
    "ssr_conv2d_carry_elems": (C.c_size_t, [C.c_void_p, C.c_int, C.c_int, C.c_int]),
    "ssr_conv_chain_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "ssr_conv_chain_begin": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "ssr_conv_chain_end": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "ssr_conv2d_fwd_carry": (C.c_int, [C.c_void_p, C.POINTER(ConvDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "ssr_f32_to_bf16_pad": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p]),
    "ssr_bf16_to_f32": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_int, C.c_void_p]),
    "ssr_axpby_bf16": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_float,
                                 C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_void_p]),
    "ssr_depth_to_space2": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                      C.c_void_p]),
    "ssr_segment_tiles": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                    C.c_void_p, C.c_void_p]),
    "ssr_stitch_tiles": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                   C.c_int, C.c_void_p, C.c_void_p]),
    "ssr_segment_tiles_ex": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                       C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "ssr_stitch_tiles_ex": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                      C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "ssr_memcpy2d_d2h": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_void_p]),
    "ssr_pixel_loss_workspace_bytes": (C.c_size_t, [C.c_int]),
    "ssr_pixel_loss": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_float, C.c_float,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ssr_total_variation_workspace_bytes": (C.c_size_t, []),
    "ssr_total_variation": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_void_p]),
    "ssr_adam_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_float,
                                C.c_float, C.c_float, C.c_float, C.c_void_p]),
    "ssr_channel_sum_workspace_bytes": (C.c_size_t, [C.c_int]),
    "ssr_channel_sum_bf16": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_int,
                                       C.c_float, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ssr_conv2d_wgrad_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "ssr_conv2d_wgrad": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                   C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_void_p,
                                   C.c_void_p, C.c_void_p]),
    "ssr_conv2d_wgrad_bias": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                        C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_void_p]),
    "ssr_conv2d_wgrad_multi_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                                            C.c_int]),
    "ssr_conv2d_wgrad_multi": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                         C.c_void_p, C.c_void_p]),
    "ssr_conv2d_pack_weights_pair": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "ssr_conv2d_pack_batch_prepare": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "ssr_conv2d_pack_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "ssr_conv2d_pack_weights_dgrad": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                                C.c_void_p, C.c_void_p]),
    "ssr_act_fwd_bf16": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_float, C.c_void_p, C.c_int, C.c_int,
                                   C.c_int64, C.c_int, C.c_void_p]),
    "ssr_act_bwd_bf16": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_float,
                                   C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_void_p]),
    "ssr_space_to_depth2": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "ssr_tanh_bwd_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "ssr_f32_to_bf16_slice": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_void_p]),
    "ssr_vgg_preprocess": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "ssr_vgg_preprocess_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_int, C.c_void_p]),
    "ssr_maxpool2_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "ssr_maxpool2_bwd_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.c_void_p]),
    "ssr_axpy_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_float, C.c_int64, C.c_void_p]),
    "ssr_subsample2": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "ssr_zero_insert2": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "ssr_bn_workspace_bytes": (C.c_size_t, [C.c_int]),
    "ssr_bn_stats_bf16": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ssr_bn_lrelu_fwd_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float,
                                        C.c_void_p, C.c_int64, C.c_int, C.c_void_p]),
    "ssr_bn_lrelu_bwd_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_float, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_int, C.c_void_p, C.c_void_p]),
    "ssr_dense_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "ssr_dense_fwd_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ssr_dense_bwd_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "ssr_lrelu_bwd_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_int64, C.c_void_p]),
    "ssr_ragan_losses": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_void_p]),
    "ssr_ragan_losses_ex": (C.c_int, [C.c_void_p, C.c_int, C.c_size_t, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ssr_gan_losses_ex": (C.c_int, [C.c_void_p, C.c_int, C.c_size_t, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ssr_comm_data_offset": (C.c_size_t, []),
    "ssr_comm_max_slots": (C.c_int, []),
    "ssr_comm_adam_slots": (C.c_int, []),
    "ssr_comm_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_size_t, C.POINTER(C.c_void_p)]),
    "ssr_comm_destroy": (C.c_int, [C.c_void_p]),
    "ssr_comm_heap": (C.c_void_p, [C.c_void_p]),
    "ssr_comm_heap_size": (C.c_size_t, [C.c_void_p]),
    "ssr_comm_ipc_handle": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ssr_comm_open_ipc": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ssr_comm_open_local": (C.c_int, [C.POINTER(C.c_void_p), C.c_int]),
    "ssr_comm_set_spin_limit": (C.c_int, [C.c_void_p, C.c_double]),
    "ssr_comm_status": (C.c_int, [C.c_void_p, C.POINTER(C.c_ulonglong)]),
    "ssr_comm_barrier": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "ssr_comm_allreduce_f32": (C.c_int, [C.c_void_p, C.c_int, C.c_size_t, C.c_void_p, C.c_void_p, C.c_int, C.c_float,
                                         C.c_void_p]),
    "ssr_opt_state_bytes": (C.c_size_t, []),
    "ssr_opt_state_set": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p]),
    "ssr_opt_state_get": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_float), C.c_void_p]),
    "ssr_opt_prepare": (C.c_int, [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_int,
                                  C.c_void_p]),
    "ssr_adam_step_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_float,
                                    C.c_float, C.c_float, C.c_float, C.c_void_p]),
    "ssr_comm_adam_step": (C.c_int, [C.c_void_p, C.c_int, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p, C.c_int64,
                                     C.c_int64, C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_void_p]),
    "ssr_bn_stats_bf16_dp": (C.c_int, [C.c_void_p, C.c_int, C.c_size_t, C.c_void_p, C.c_int64, C.c_int, C.c_float,
                                       C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ssr_bn_lrelu_bwd_bf16_dp": (C.c_int, [C.c_void_p, C.c_int, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.c_void_p, C.c_float, C.c_int64, C.c_int, C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "ssr_act_split_f32": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "ssr_bf16_residual_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "ssr_resize_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "ssr_resize_bicubic": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.c_void_p, C.c_void_p]),
    "ssr_augment": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                              C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ssr_metric_workspace_bytes": (C.c_size_t, [C.c_int]),
    "ssr_psnr_y": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p,
                             C.c_void_p]),
    "ssr_ssim": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p,
                           C.c_void_p]),
    "ssr_diag_mma_rate": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float)]),
    "ssr_diag_mma_rate_ex": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float)]),
    "ssr_diag_mma_rate_pair": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_float)]),
    "ssr_debug_trace": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ssr_debug_trace_ring": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "ssr_debug_set": (C.c_int, [C.c_void_p, C.c_int]),
    "ssr_debug_last_conv_tiles": (C.c_int, [C.c_void_p]),
}

_lib = None


def load():
    """Load the shared library once; fail loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SsrError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)")
        lib = C.CDLL(os.environ.get("SSR_LIB_PATH") or LIB_PATH)     # SSR_LIB_PATH: development A/B of two builds
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def exported_symbols():
    return sorted(_SIGNATURES)


def check(rc):
    if rc == 0:
        return
    msg = load().ssr_last_error().decode("utf-8", "replace")
    if rc == -1:
        raise ValueError(msg)  # the reference raises ValueError for bad arguments
    if rc == -4:
        raise MemoryError(msg)
    raise SsrError(f"ssr error {rc}: {msg}")


def _ptr(x):
    if x is None:
        return None
    if isinstance(x, (DeviceBuffer, DeviceView)):
        return x.ptr
    return int(x)


class DeviceView:
    """A non-owning window into a DeviceBuffer (e.g. one variable inside the flat parameter buffer)."""

    def __init__(self, base, offset, nbytes):
        self.base, self.offset, self.nbytes = base, int(offset), int(nbytes)

    @property
    def ptr(self):
        return self.base.ptr + self.offset

    def upload(self, arr, stream=None, offset=0):
        arr = np.ascontiguousarray(arr)
        assert offset + arr.nbytes <= self.nbytes
        check(load().ssr_memcpy_h2d(self.ptr + offset, arr.ctypes.data, arr.nbytes, stream))
        check(load().ssr_stream_sync(stream))

    def download(self, shape, dtype, stream=None, offset=0):
        out = np.empty(shape, dtype=dtype)
        assert offset + out.nbytes <= self.nbytes
        check(load().ssr_memcpy_d2h(out.ctypes.data, self.ptr + offset, out.nbytes, stream))
        check(load().ssr_stream_sync(stream))
        return out

    def zero(self, stream=None):
        check(load().ssr_memset(self.ptr, 0, self.nbytes, stream))

    def free(self):
        pass


PACK_ENTRY_BYTES = 128


class WgradItem(C.Structure):
    """ssr_wgrad_item (include/ssr_b200.h): one convolution of a batched weight-gradient launch."""
    _fields_ = [("x", C.c_void_p), ("x_cstride", C.c_int32), ("x_coff", C.c_int32), ("cin_real", C.c_int32),
                ("dz", C.c_void_p), ("dz_cstride", C.c_int32), ("dz_coff", C.c_int32), ("cout", C.c_int32),
                ("scale", C.c_float), ("accumulate", C.c_int32), ("dw_hwio", C.c_void_p), ("dbias", C.c_void_p),
                ("bias_scale", C.c_float), ("bias_accumulate", C.c_int32)]


class PackItem(C.Structure):
    """ssr_pack_item (include/ssr_b200.h): one weight image of a batched re-pack, forward conv geometry."""
    _fields_ = [("w_hwio", C.c_void_p), ("packed", C.c_void_p), ("kh", C.c_int32), ("kw", C.c_int32),
                ("cin_real", C.c_int32), ("cin", C.c_int32), ("cout", C.c_int32), ("up", C.c_int32),
                ("mode", C.c_int32), ("reserved", C.c_int32), ("k0", C.c_int32), ("kn", C.c_int32), ("row0", C.c_int32),
                ("src_cin", C.c_int32), ("src_cout", C.c_int32), ("scale", C.c_float)]


class DeviceBuffer:
    """A cudaMalloc'd region.  ``ptr`` is the raw device address."""

    def __init__(self, nbytes):
        self.nbytes = int(nbytes)
        p = C.c_void_p()
        check(load().ssr_malloc(C.byref(p), self.nbytes))
        self.ptr = p.value
        self._owned = True

    @classmethod
    def from_numpy(cls, arr, stream=None):
        arr = np.ascontiguousarray(arr)
        buf = cls(arr.nbytes)
        buf.upload(arr, stream)
        return buf

    def upload(self, arr, stream=None, offset=0):
        arr = np.ascontiguousarray(arr)
        assert offset + arr.nbytes <= self.nbytes
        check(load().ssr_memcpy_h2d(self.ptr + offset, arr.ctypes.data, arr.nbytes, stream))
        check(load().ssr_stream_sync(stream))

    def download(self, shape, dtype, stream=None, offset=0):
        out = np.empty(shape, dtype=dtype)
        assert offset + out.nbytes <= self.nbytes
        check(load().ssr_memcpy_d2h(out.ctypes.data, self.ptr + offset, out.nbytes, stream))
        check(load().ssr_stream_sync(stream))
        return out

    def zero(self, stream=None):
        check(load().ssr_memset(self.ptr, 0, self.nbytes, stream))

    def free(self):
        if getattr(self, "_owned", False) and self.ptr:
            if _capture.depth:
                _capture.deferred.append(self.ptr)    # cudaFree inside a stream capture would invalidate the capture
            else:
                load().ssr_free(self.ptr)
            self.ptr = None
            self._owned = False

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class _CaptureState(threading.local):
    """Per host thread (captures are cudaStreamCaptureModeThreadLocal): nesting depth of Graph recordings and the
    device pointers whose release was requested meanwhile."""
    depth = 0

    def __init__(self):
        self.deferred = []


_capture = _CaptureState()


class Context:
    """Per-device context (SM count, driver entry points).  Raises when no sm_100 GPU is visible."""

    def __init__(self, device=0):
        self.lib = load()
        h = C.c_void_p()
        check(self.lib.ssr_ctx_create(int(device), C.byref(h)))
        self.handle = h.value
        self.device = int(device)

    @property
    def sm_count(self):
        return self.lib.ssr_ctx_sm_count(self.handle)

    @property
    def launch_count(self):
        return self.lib.ssr_ctx_launch_count(self.handle)

    @property
    def last_conv_tiles(self):
        return self.lib.ssr_debug_last_conv_tiles(self.handle)

    def debug_set(self, flags=0, force_wb=0):
        check(self.lib.ssr_debug_set(self.handle, (flags & 0x7FFF00FF) | ((force_wb & 0xFF) << 8)))

    def conv2d_fwd_mask(self, desc, x, w_packed, bias, res, out, mask_z, mask_z_cstride, mask_z_coff, mask_lo, mask_n,
                        mask_alpha, mask_out, mask_out_cstride, stream=None):
        check(self.lib.ssr_conv2d_fwd_mask(self.handle, C.byref(desc), _ptr(x), _ptr(w_packed), _ptr(bias), _ptr(res),
                                           _ptr(out), _ptr(mask_z), mask_z_cstride, mask_z_coff, mask_lo, mask_n,
                                           mask_alpha, _ptr(mask_out), mask_out_cstride, stream))

    def conv2d_fwd_carry(self, desc, x, w_packed, bias, out, carry_in=None, carry_out=None, carry_out_cols=0, res=None,
                         stream=None):
        check(self.lib.ssr_conv2d_fwd_carry(self.handle, C.byref(desc), _ptr(x), _ptr(w_packed), _ptr(bias), _ptr(res),
                                            _ptr(out), _ptr(carry_in), _ptr(carry_out), carry_out_cols, stream))

    def conv_chain_begin(self, buf, stream=None):
        """Conv launches of this thread honour ``desc.chain`` until :meth:`conv_chain_end` (ssr_conv_chain_begin)."""
        check(self.lib.ssr_conv_chain_begin(self.handle, _ptr(buf), buf.nbytes if buf is not None else 0, stream))

    def conv_chain_end(self):
        """-> (launches that published tile flags, launches that ran with tile-level dependencies)."""
        pub, dep = C.c_int64(0), C.c_int64(0)
        check(self.lib.ssr_conv_chain_end(self.handle, C.byref(pub), C.byref(dep)))
        return pub.value, dep.value

    def diag_mma_rate_pair(self, n, iters=4096):
        v = (C.c_float * 2)()
        check(self.lib.ssr_diag_mma_rate_pair(self.handle, n, iters, v))
        return v[0], v[1]

    def debug_trace(self, buf):
        check(self.lib.ssr_debug_trace(self.handle, _ptr(buf)))

    def debug_trace_ring(self, buf, slots):
        check(self.lib.ssr_debug_trace_ring(self.handle, _ptr(buf), slots))

    def close(self):
        if self.handle:
            self.lib.ssr_ctx_destroy(self.handle)
            self.handle = None

    # ---- conv2d
    def conv_packed_bytes(self, ksize, cin, cout, up=1, ksize_w=None):
        if ksize_w is None:
            n = self.lib.ssr_conv2d_packed_bytes(ksize, cin, cout, up)
        else:
            n = self.lib.ssr_conv2d_packed_bytes_hw(ksize, ksize_w, cin, cout, up)
        if n == 0:
            raise ValueError(self.lib.ssr_last_error().decode())
        return n

    def conv_pack_weights(self, w_hwio_dev, ksize, cin_real, cin, cout, up, packed_dev, stream=None, ksize_w=None):
        check(self.lib.ssr_conv2d_pack_weights_hw(self.handle, _ptr(w_hwio_dev), ksize, ksize_w or ksize, cin_real, cin,
                                                  cout, up, _ptr(packed_dev), stream))

    def conv2d_fwd(self, desc, x, w_packed, bias, out, alpha=None, res=None, out2=None, stream=None):
        check(self.lib.ssr_conv2d_fwd(self.handle, C.byref(desc), _ptr(x), _ptr(w_packed), _ptr(bias), _ptr(alpha),
                                      _ptr(res), _ptr(out), _ptr(out2), stream))

    def conv_pack_weights_pair(self, w_hwio_dev, cin_real, cin, cout, packed_dev, stream=None):
        check(self.lib.ssr_conv2d_pack_weights_pair(self.handle, _ptr(w_hwio_dev), cin_real, cin, cout, _ptr(packed_dev),
                                                    stream))

    def pack_batch_prepare(self, items, stream=None):
        """items: list of PackItem -> device table (DeviceBuffer) for pack_batch."""
        arr = (PackItem * len(items))(*items)
        table = DeviceBuffer(len(items) * PACK_ENTRY_BYTES)
        check(self.lib.ssr_conv2d_pack_batch_prepare(self.handle, C.byref(arr), len(items), table.ptr, stream))
        return table

    def pack_batch(self, table, count, stream=None):
        check(self.lib.ssr_conv2d_pack_batch(self.handle, _ptr(table), count, stream))

    def conv_pack_weights_dgrad(self, w_hwio_dev, kh, kw, cin_fwd, cout_fwd, packed_dev, unroll_x=False, stream=None):
        check(self.lib.ssr_conv2d_pack_weights_dgrad(self.handle, _ptr(w_hwio_dev), kh, kw, cin_fwd, cout_fwd,
                                                     int(unroll_x), _ptr(packed_dev), stream))

    def conv_wgrad_workspace_bytes(self, h, w, cin, cout, kh, kw):
        n = self.lib.ssr_conv2d_wgrad_workspace_bytes(self.handle, h, w, cin, cout, kh, kw)
        if n == 0:
            raise ValueError(self.lib.ssr_last_error().decode())
        return n

    def conv2d_wgrad(self, x, x_cs, x_off, cin_real, dz, dz_cs, dz_off, cout, n, h, w, kh, kw, workspace, dw,
                     scale=1.0, accumulate=False, stream=None, dbias=None, bias_scale=1.0, bias_accumulate=False):
        if dbias is not None:
            check(self.lib.ssr_conv2d_wgrad_bias(self.handle, _ptr(x), x_cs, x_off, cin_real, _ptr(dz), dz_cs, dz_off,
                                                 cout, n, h, w, kh, kw, scale, int(accumulate), _ptr(workspace),
                                                 _ptr(dw), _ptr(dbias), bias_scale, int(bias_accumulate), stream))
            return
        check(self.lib.ssr_conv2d_wgrad(self.handle, _ptr(x), x_cs, x_off, cin_real, _ptr(dz), dz_cs, dz_off, cout, n, h,
                                        w, kh, kw, scale, int(accumulate), _ptr(workspace), _ptr(dw), stream))

    @staticmethod
    def wgrad_item(x, x_cs, x_off, cin_real, dz, dz_cs, dz_off, cout, dw, scale=1.0, accumulate=False, dbias=None,
                   bias_scale=1.0, bias_accumulate=False):
        return WgradItem(_ptr(x), x_cs, x_off, cin_real, _ptr(dz), dz_cs, dz_off, cout, scale, int(accumulate), _ptr(dw),
                         _ptr(dbias), bias_scale, int(bias_accumulate))

    def conv_wgrad_multi_workspace_bytes(self, items, h, w, kh, kw):
        arr = (WgradItem * len(items))(*items)
        n = self.lib.ssr_conv2d_wgrad_multi_workspace_bytes(self.handle, C.byref(arr), len(items), h, w, kh, kw)
        if n == 0:
            raise ValueError(self.lib.ssr_last_error().decode())
        return n

    def conv2d_wgrad_multi(self, items, n, h, w, kh, kw, workspace, stream=None):
        """Up to 8 weight gradients (same n, h, w, kernel size) in one launch + one reduction."""
        arr = (WgradItem * len(items))(*items)
        check(self.lib.ssr_conv2d_wgrad_multi(self.handle, C.byref(arr), len(items), n, h, w, kh, kw, _ptr(workspace),
                                              stream))

    def diag_mma_rate(self, n, iters=4096, a_shift_rows=0):
        v = (C.c_float * 2)()
        check(self.lib.ssr_diag_mma_rate(self.handle, n, iters, a_shift_rows, v))
        return v[0]


    def diag_mma_rate_ex(self, m, n, a_swizzle=2, iters=4096):
        v = (C.c_float * 2)()
        check(self.lib.ssr_diag_mma_rate_ex(self.handle, m, n, a_swizzle, iters, v))
        return v[0], v[1]


class Stream:
    def __init__(self):
        p = C.c_void_p()
        check(load().ssr_stream_create(C.byref(p)))
        self.ptr = p.value

    def sync(self):
        check(load().ssr_stream_sync(self.ptr))

    def wait_event(self, event):
        check(load().ssr_stream_wait_event(self.ptr, event.ptr))

    def destroy(self):
        if self.ptr:
            load().ssr_stream_destroy(self.ptr)
            self.ptr = None


def stream_wait_event(stream_ptr, event):
    """cudaStreamWaitEvent on a raw stream pointer (None = the default stream)."""
    check(load().ssr_stream_wait_event(stream_ptr, event.ptr))


class Event:
    def __init__(self):
        p = C.c_void_p()
        check(load().ssr_event_create(C.byref(p)))
        self.ptr = p.value

    def record(self, stream=None):
        check(load().ssr_event_record(self.ptr, stream))

    def sync(self):
        check(load().ssr_event_sync(self.ptr))

    def elapsed_ms(self, stop):
        v = C.c_float()
        check(load().ssr_event_elapsed_ms(self.ptr, stop.ptr, C.byref(v)))
        return v.value

    def destroy(self):
        if self.ptr:
            load().ssr_event_destroy(self.ptr)
            self.ptr = None


class OpsView:
    """List-like view of a plan's launch list (callables taking a stream pointer) that can redirect the launches
    appended through it to another stream: ``view.redirect = Stream`` ... ``view.redirect = None``."""

    def __init__(self, real):
        self.real, self.redirect = real, None

    def append(self, fn):
        if self.redirect is None:
            self.real.append(fn)
        else:
            r = self.redirect
            self.real.append(lambda s, fn=fn, r=r: fn(r.ptr))


class Graph:
    """A captured + instantiated CUDA graph of ABI launches on one stream."""

    def __init__(self, stream_ptr, record_fn):
        lib = load()
        # nothing may allocate or free device memory on this thread while the stream records: the cyclic collector is
        # paused (a leaked buffer finalised here would call cudaFree) and DeviceBuffer.free defers until the capture ends
        gc_was_on = gc.isenabled()
        gc.disable()
        check(lib.ssr_graph_begin(stream_ptr))
        _capture.depth += 1
        try:
            record_fn()
        finally:
            g = C.c_void_p()
            rc = lib.ssr_graph_end(stream_ptr, C.byref(g))
            _capture.depth -= 1
            if not _capture.depth:
                for ptr in _capture.deferred:
                    lib.ssr_free(ptr)
                del _capture.deferred[:]
            if gc_was_on:
                gc.enable()
        check(rc)
        self.ptr = g.value
        self.kernels = lib.ssr_graph_last_kernel_count()   # kernel nodes = kernels per launch of this graph

    def launch(self, stream_ptr):
        check(load().ssr_graph_launch(self.ptr, stream_ptr))

    def destroy(self):
        if self.ptr:
            load().ssr_graph_destroy(self.ptr)
            self.ptr = None


class PinnedArray:
    """numpy view over cudaMallocHost memory (for end-to-end timing with true async copies)."""

    def __init__(self, shape, dtype):
        self.shape = tuple(shape)
        self.dtype = np.dtype(dtype)
        self.nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        p = C.c_void_p()
        check(load().ssr_host_alloc(C.byref(p), self.nbytes))
        self.ptr = p.value
        buf = (C.c_char * self.nbytes).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=self.dtype).reshape(self.shape)

    def free(self):
        if self.ptr:
            self.array = None
            load().ssr_host_free(self.ptr)
            self.ptr = None


# ---- stateless bandwidth kernels
def f32_to_bf16_pad(x, y, pixels, c, cpad, stream=None):
    check(load().ssr_f32_to_bf16_pad(_ptr(x), _ptr(y), pixels, c, cpad, stream))


def im2col_x_f32_to_bf16(x, y, n, h, w, c, kw, cpad, stream=None):
    check(load().ssr_im2col_x_f32_to_bf16(_ptr(x), _ptr(y), n, h, w, c, kw, cpad, stream))


def bf16_to_f32(x, x_cstride, x_coff, y, pixels, c, stream=None):
    check(load().ssr_bf16_to_f32(_ptr(x), x_cstride, x_coff, _ptr(y), pixels, c, stream))


def axpby_bf16(a, a_cs, a_off, b, b_cs, b_off, beta, out, o_cs, o_off, pixels, c, stream=None):
    check(load().ssr_axpby_bf16(_ptr(a), a_cs, a_off, _ptr(b), b_cs, b_off, beta, _ptr(out), o_cs, o_off, pixels, c,
                                stream))


def depth_to_space2(x, y, n, h, w, c, elem_bytes, stream=None):
    check(load().ssr_depth_to_space2(_ptr(x), _ptr(y), n, h, w, c, elem_bytes, stream))


def segment_tiles(img, h, w, c, patch, overlap, tile_begin, tile_count, tiles, stream=None):
    check(load().ssr_segment_tiles(_ptr(img), h, w, c, patch, overlap, tile_begin, tile_count, _ptr(tiles), stream))


def stitch_tiles(tiles, h, w, c, patch, overlap, scale, tile_begin, tile_count, out, stream=None):
    check(load().ssr_stitch_tiles(_ptr(tiles), h, w, c, patch, overlap, scale, tile_begin, tile_count, _ptr(out),
                                  stream))


def segment_tiles_ex(img, h, w, c, patch_h, patch_w, overlap, tile_begin, tile_count, src_row0, src_rows, tiles, stream=None):
    check(load().ssr_segment_tiles_ex(_ptr(img), h, w, c, patch_h, patch_w, overlap, tile_begin, tile_count, src_row0,
                                      src_rows, _ptr(tiles), stream))


def stitch_tiles_ex(tiles, h, w, c, patch_h, patch_w, overlap, scale, tile_begin, tile_count, out_row0, out_rows, out,
                    stream=None):
    check(load().ssr_stitch_tiles_ex(_ptr(tiles), h, w, c, patch_h, patch_w, overlap, scale, tile_begin, tile_count,
                                     out_row0, out_rows, _ptr(out), stream))


def pixel_loss(hr, sr, n, per_image, w_mse, w_mae, max_val, grad, workspace, out, stream=None):
    check(load().ssr_pixel_loss(_ptr(hr), _ptr(sr), n, per_image, w_mse, w_mae, max_val, _ptr(grad), _ptr(workspace),
                                _ptr(out), stream))


def total_variation(x, n, h, w, c, value_scale, weight, grad, workspace, out1, stream=None):
    check(load().ssr_total_variation(_ptr(x), n, h, w, c, value_scale, weight, _ptr(grad), _ptr(workspace), _ptr(out1),
                                     stream))


def adam_step(param, grad, m, v, count, lr_t, beta1, beta2, eps, grad_scale=1.0, stream=None):
    check(load().ssr_adam_step(_ptr(param), _ptr(grad), _ptr(m), _ptr(v), count, lr_t, beta1, beta2, eps, grad_scale,
                               stream))


def channel_sum_bf16(x, x_cs, x_off, z, z_cs, z_off, pixels, c, scale, accumulate, workspace, out, stream=None):
    check(load().ssr_channel_sum_bf16(_ptr(x), x_cs, x_off, _ptr(z), z_cs, z_off, pixels, c, scale, int(accumulate),
                                      _ptr(workspace), _ptr(out), stream))


def act_fwd_bf16(z, z_cs, z_off, alpha, alpha_scalar, y, y_cs, y_off, pixels, c, stream=None):
    check(load().ssr_act_fwd_bf16(_ptr(z), z_cs, z_off, _ptr(alpha), alpha_scalar, _ptr(y), y_cs, y_off, pixels, c,
                                  stream))


def act_bwd_bf16(dy, dy_cs, dy_off, z, z_cs, z_off, alpha, alpha_scalar, dz, dz_cs, dz_off, pixels, c, stream=None):
    check(load().ssr_act_bwd_bf16(_ptr(dy), dy_cs, dy_off, _ptr(z), z_cs, z_off, _ptr(alpha), alpha_scalar, _ptr(dz),
                                  dz_cs, dz_off, pixels, c, stream))


def space_to_depth2(x, y, n, h, w, c, elem_bytes, stream=None):
    check(load().ssr_space_to_depth2(_ptr(x), _ptr(y), n, h, w, c, elem_bytes, stream))


def tanh_bwd_f32(g, y, dz, count, stream=None):
    check(load().ssr_tanh_bwd_f32(_ptr(g), _ptr(y), _ptr(dz), count, stream))


def f32_to_bf16_slice(x, y, y_cs, y_off, pixels, c, stream=None):
    check(load().ssr_f32_to_bf16_slice(_ptr(x), _ptr(y), y_cs, y_off, pixels, c, stream))


def vgg_preprocess(x, y, pixels, stream=None):
    check(load().ssr_vgg_preprocess(_ptr(x), _ptr(y), pixels, stream))


def vgg_preprocess_bwd(dy, g, pixels, scale, accumulate, stream=None):
    check(load().ssr_vgg_preprocess_bwd(_ptr(dy), _ptr(g), pixels, scale, int(accumulate), stream))


def maxpool2_bf16(x, y, n, h, w, c, stream=None):
    check(load().ssr_maxpool2_bf16(_ptr(x), _ptr(y), n, h, w, c, stream))


def maxpool2_bwd_bf16(x, dy, dx, n, h, w, c, stream=None):
    check(load().ssr_maxpool2_bwd_bf16(_ptr(x), _ptr(dy), _ptr(dx), n, h, w, c, stream))


def axpy_f32(x, y, a, count, stream=None):
    check(load().ssr_axpy_f32(_ptr(x), _ptr(y), a, count, stream))


def subsample2(x, y, n, oh, ow, c, elem_bytes, stream=None):
    check(load().ssr_subsample2(_ptr(x), _ptr(y), n, oh, ow, c, elem_bytes, stream))


def zero_insert2(dy, dx, n, oh, ow, c, elem_bytes, stream=None):
    check(load().ssr_zero_insert2(_ptr(dy), _ptr(dx), n, oh, ow, c, elem_bytes, stream))


def bn_stats_bf16(x, pixels, c, eps, momentum, ws, mean, istd, mmean, mvar, stream=None, site=None):
    """``site`` = (comm handle, slot0, heap offset) from ``PeerComm.bn_site``: statistics over all ranks' pixels."""
    if site is not None:
        check(load().ssr_bn_stats_bf16_dp(site[0], site[1], site[2], _ptr(x), pixels, c, eps, momentum, _ptr(ws),
                                          _ptr(mean), _ptr(istd), _ptr(mmean), _ptr(mvar), stream))
        return
    check(load().ssr_bn_stats_bf16(_ptr(x), pixels, c, eps, momentum, _ptr(ws), _ptr(mean), _ptr(istd), _ptr(mmean),
                                   _ptr(mvar), stream))


def bn_lrelu_fwd_bf16(x, mean, istd, gamma, beta, alpha, y, pixels, c, stream=None):
    check(load().ssr_bn_lrelu_fwd_bf16(_ptr(x), _ptr(mean), _ptr(istd), _ptr(gamma), _ptr(beta), alpha, _ptr(y), pixels,
                                       c, stream))


def bn_lrelu_bwd_bf16(x, dy, y, mean, istd, gamma, alpha, pixels, c, ws, sums, dgamma, dbeta, accumulate, dz,
                      stream=None, site=None):
    if site is not None:
        check(load().ssr_bn_lrelu_bwd_bf16_dp(site[0], site[1], site[2], _ptr(x), _ptr(dy), _ptr(y), _ptr(mean), _ptr(istd),
                                              _ptr(gamma), alpha, pixels, c, _ptr(ws), _ptr(sums), _ptr(dgamma),
                                              _ptr(dbeta), int(accumulate), _ptr(dz), stream))
        return
    check(load().ssr_bn_lrelu_bwd_bf16(_ptr(x), _ptr(dy), _ptr(y), _ptr(mean), _ptr(istd), _ptr(gamma), alpha, pixels, c,
                                       _ptr(ws), _ptr(sums), _ptr(dgamma), _ptr(dbeta), int(accumulate), _ptr(dz),
                                       stream))


def dense_fwd_f32(x, w, b, n, fin, fout, lrelu, alpha, ws, pre_act, y, stream=None):
    check(load().ssr_dense_fwd_f32(_ptr(x), _ptr(w), _ptr(b), n, fin, fout, int(lrelu), alpha, _ptr(ws), _ptr(pre_act),
                                   _ptr(y), stream))


def dense_bwd_f32(x, w, dy, n, fin, fout, dx, dw, db, accumulate, stream=None):
    check(load().ssr_dense_bwd_f32(_ptr(x), _ptr(w), _ptr(dy), n, fin, fout, _ptr(dx), _ptr(dw), _ptr(db),
                                   int(accumulate), stream))


def lrelu_bwd_f32(dy, h, alpha, dh, count, stream=None):
    check(load().ssr_lrelu_bwd_f32(_ptr(dy), _ptr(h), alpha, _ptr(dh), count, stream))


def ragan_losses(hc, sc, n, hr_label, sr_label, out2, g_dsr, d_dsr, d_dhr, stream=None):
    check(load().ssr_ragan_losses(_ptr(hc), _ptr(sc), n, hr_label, sr_label, _ptr(out2), _ptr(g_dsr), _ptr(d_dsr),
                                  _ptr(d_dhr), stream))


def ragan_losses_ex(hc, sc, n_local, hr_label, sr_label, hr_labels, sr_labels, out2, g_dsr, d_dsr, d_dhr, stream=None,
                    site=None, relativistic=True):
    """Per-sample label arrays (or None) and, with ``site`` = (comm handle, slot, heap offset), global-batch means.
    ``relativistic=False``: the standard (sigmoid + BinaryCrossentropy) critic, ssr_gan_losses_ex."""
    comm, slot, off = site if site is not None else (None, 0, 0)
    fn = load().ssr_ragan_losses_ex if relativistic else load().ssr_gan_losses_ex
    check(fn(comm, slot, off, _ptr(hc), _ptr(sc), n_local, hr_label, sr_label, _ptr(hr_labels), _ptr(sr_labels), _ptr(out2),
             _ptr(g_dsr), _ptr(d_dsr), _ptr(d_dhr), stream))


def opt_prepare(state, base_lr, b1, b2, boundaries=None, values=None, n_boundaries=0, stream=None):
    check(load().ssr_opt_prepare(_ptr(state), base_lr, b1, b2, _ptr(boundaries), _ptr(values), n_boundaries, stream))


def adam_step_dev(param, grad, m, v, count, opt_state, beta1, beta2, eps, grad_scale=1.0, stream=None):
    check(load().ssr_adam_step_dev(_ptr(param), _ptr(grad), _ptr(m), _ptr(v), count, _ptr(opt_state), beta1, beta2, eps,
                                   grad_scale, stream))


def act_split_f32(z, n, h, w, c_out, up, act, act_alpha, alpha, res32, y32, hi_lo, hl_cstride, hi_off, lo_off, stream=None):
    check(load().ssr_act_split_f32(_ptr(z), n, h, w, c_out, up, act, act_alpha, _ptr(alpha), _ptr(res32), _ptr(y32),
                                   _ptr(hi_lo), hl_cstride, hi_off, lo_off, stream))


def stream_sync(stream=None):
    check(load().ssr_stream_sync(stream))


# ---- bf16 <-> fp32 on the host (numpy has no bfloat16): bit patterns in uint16
def f32_to_bf16_bits(a):
    """Round-to-nearest-even fp32 -> bf16 bit pattern (uint16), same as cvt.rn.bf16.f32."""
    u = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)
    rounded = u + 0x7FFF + ((u >> 16) & 1)
    return (rounded >> 16).astype(np.uint16)


def bf16_bits_to_f32(b):
    return (np.ascontiguousarray(b, dtype=np.uint16).astype(np.uint32) << 16).view(np.float32)


def bf16_round(a):
    return bf16_bits_to_f32(f32_to_bf16_bits(a))
