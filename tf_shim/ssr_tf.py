"""Python side of the TensorFlow custom-op shim (tf_shim/ssr_tf_ops.cc): loads the op library and registers the
gradients, so that ``tape.gradient(loss, model.trainable_variables)`` (simple_sr/models/sr_model.py:419-447) traverses
layers built from the ops.  UNTESTED in this repository's image (TensorFlow is not installable there);
build with ``python tf_shim/build.py`` where ``import tensorflow`` works.

    from tf_shim import ssr_tf
    y = ssr_tf.conv2d_same(x_bf16, kernel_hwio_f32, bias_f32, act=ssr_tf.ACT_LRELU)      # differentiable
"""
import os

import tensorflow as tf

_HERE = os.path.dirname(os.path.abspath(__file__))
_ops = tf.load_op_library(os.path.join(_HERE, "libssr_tf_ops.so"))

ACT_NONE, ACT_LRELU, ACT_PRELU, ACT_TANH, ACT_RELU = 0, 1, 2, 3, 4
_EMPTY_F32 = lambda: tf.zeros([0], tf.float32)
_EMPTY_BF16 = lambda: tf.zeros([0], tf.bfloat16)


def depth_to_space2(x):
    return _ops.ssr_depth_to_space2(x)


@tf.RegisterGradient("SsrDepthToSpace2")
def _d2s_grad(op, dy):
    return tf.nn.space_to_depth(dy, 2)      # the adjoint permutation (ssr_space_to_depth2 behind the ABI)


def conv2d_same(x, kernel, bias, act=ACT_NONE, act_alpha=0.2):
    """Conv2D(padding="same", strides=1) + BiasAdd + activation on bf16 NHWC activations, fp32 HWIO kernel.  The packed
    operand images are produced by ops too, so the whole expression lives in the TF graph."""
    kh, kw, cin, cout = kernel.shape
    cin_p = -(-cin // 16) * 16

    @tf.custom_gradient
    def f(x, kernel, bias):
        packed = _ops.ssr_pack_weights(kernel, cin_padded=cin_p)
        z = _ops.ssr_conv2d(x, packed, bias, _EMPTY_F32(), _EMPTY_BF16(), cin=cin_p, cout=cout, ksize=kh, act=ACT_NONE)
        y = z if act == ACT_NONE else (tf.nn.leaky_relu(z, act_alpha) if act == ACT_LRELU else
                                       tf.nn.relu(z) if act == ACT_RELU else tf.tanh(z))

        def grad(dy):
            if act == ACT_LRELU:
                dz = dy * tf.cast(tf.where(z > 0, 1.0, act_alpha), dy.dtype)
            elif act == ACT_RELU:
                dz = dy * tf.cast(z > 0, dy.dtype)
            elif act == ACT_TANH:
                dz = dy * (1 - tf.square(tf.tanh(z)))
            else:
                dz = dy
            dpacked = _ops.ssr_pack_weights(kernel, cin_padded=cin_p, dgrad=True)
            cout_p = -(-cout // 16) * 16
            dx = _ops.ssr_conv2d(dz, dpacked, tf.zeros([cin], tf.float32), _EMPTY_F32(), _EMPTY_BF16(), cin=cout_p,
                                 cout=cin, ksize=kh, act=ACT_NONE)
            dk, db = _ops.ssr_conv2d_grad_filter(x, dz, cin=cin, cout=cout, ksize=kh)
            return dx, dk, db

        return y, grad

    return f(x, kernel, bias)
