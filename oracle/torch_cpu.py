"""TEST / BENCH INFRASTRUCTURE - not part of the product path (nothing under simplesr_b200/ imports this).

The reference's CPU path is TensorFlow 2.2 (Eigen / oneDNN convolutions on the host cores); TensorFlow cannot be
installed in this image, so the CPU baseline arm of bench.py times this restatement of
``model_builder.build_enhanced_resnet`` (simple_sr/utils/models/model_builder.py:42-96, 328-365) on
``torch.nn.functional.conv2d`` - oneDNN, the same class of CPU convolution backend - with every host thread.
``tests/test_oracle_ops.py`` pins it to the numpy oracle (oracle/ssr_oracle.py) on a small case.
"""
import math
import os

import numpy as np


def set_threads(torch, threads=None):
    """torchrun exports OMP_NUM_THREADS=1: ask for all host cores explicitly."""
    n = threads or os.cpu_count() or 1
    torch.set_num_threads(n)
    return torch.get_num_threads()


def _conv(torch, x, k, b):
    """TF Conv2D(padding="same", strides=1) on NCHW tensors; HWIO kernel."""
    w = torch.from_numpy(np.ascontiguousarray(np.transpose(k, (3, 2, 0, 1))))
    return torch.nn.functional.conv2d(x, w, torch.from_numpy(b), padding=(k.shape[0] // 2, k.shape[1] // 2))


def _d2s(torch, x, block=2):
    """tf.nn.depth_to_space NHWC "DCR" on an NCHW tensor: out[n, c, 2h+i, 2w+j] = in[n, (2i+j)*C + c, h, w]."""
    n, c4, h, w = x.shape
    c = c4 // (block * block)
    x = x.reshape(n, block, block, c, h, w).permute(0, 3, 4, 1, 5, 2)
    return x.reshape(n, c, h * block, w * block)


def rrdb_forward(params, x_nhwc, upsample_factor=4, num_rrdb_blocks=16, num_dense_blocks=3, num_convs=4,
                 residual_scaling=0.2):
    """Same graph as oracle.ssr_oracle.rrdb_forward, fp32, NHWC numpy in / out."""
    import torch
    lrelu = lambda t: torch.nn.functional.leaky_relu(t, 0.2)
    with torch.no_grad():
        x = torch.from_numpy(np.ascontiguousarray(np.transpose(x_nhwc, (0, 3, 1, 2))))
        fea = _conv(torch, x, *params["fea"][:2])
        t = fea
        for b in range(num_rrdb_blocks):
            for d in range(num_dense_blocks):
                pre = f"rrdb{b}_db{d}"
                cat = t
                for k in range(num_convs):
                    cat = torch.cat([cat, lrelu(_conv(torch, cat, *params[f"{pre}_conv{k}"][:2]))], dim=1)
                t = t + residual_scaling * _conv(torch, cat, *params[f"{pre}_out"][:2])
        t = fea + residual_scaling * t
        t = fea + _conv(torch, t, *params["trunk"][:2])
        for i in range(int(math.log(upsample_factor, 2))):
            t = lrelu(_d2s(torch, _conv(torch, t, *params[f"up{i}"][:2])))
        t = lrelu(_conv(torch, t, *params["hr"][:2]))
        t = torch.tanh(_conv(torch, t, *params["last"][:2]))
        return np.ascontiguousarray(t.permute(0, 2, 3, 1).numpy())
