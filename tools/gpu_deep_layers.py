"""Deep layers (cin >= 256) with and without slab pairing: time per launch at the VGG / discriminator shapes of the
ESRGAN step (batch 16).  usage: python tools/gpu_deep_layers.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simplesr_b200 import _lib as L  # noqa: E402
from simplesr_b200 import model_builder as MB  # noqa: E402

ctx = MB.get_context(0)
s = L.Stream()
shapes = [("vgg b2 128->128 @64", 16, 64, 64, 128, 128), ("vgg b3 256->256 @32", 16, 32, 32, 256, 256),
          ("vgg b4 256->512 @16", 16, 16, 16, 256, 512), ("vgg b4 512->512 @16", 16, 16, 16, 512, 512),
          ("vgg b5 512->512 @8", 16, 8, 8, 512, 512), ("disc 256->256 @32", 16, 32, 32, 256, 256),
          ("disc 512->512 @16", 16, 16, 16, 512, 512)]
rng = np.random.default_rng(0)
for name, n, h, w, cin, cout in shapes:
    x = L.DeviceBuffer(n * h * w * cin * 2)
    x.zero()
    wts = L.DeviceBuffer.from_numpy((rng.standard_normal((3, 3, cin, cout)) / 50).astype(np.float32))
    packed = L.DeviceBuffer(ctx.conv_packed_bytes(3, cin, cout, 1))
    ctx.conv_pack_weights(wts, 3, cin, cin, cout, 1, packed)
    out = L.DeviceBuffer(n * h * w * cout * 2)
    d = L.ConvDesc(n=n, h=h, w=w, cin=cin, in_cstride=cin, cout=cout, ksize=3, ksize_w=3, act=L.ACT_RELU, act_alpha=0.0,
                   res_beta=0.0, up=1, out_dtype=L.SSR_BF16, out_cstride=cout, out_coff=0, res_dtype=L.SSR_NONE)
    res = {}
    for tag, flag in (("one slab per CTA", 0x10000), ("slab pairs", 0)):
        ctx.debug_set(flag)
        for _ in range(3):
            ctx.conv2d_fwd(d, x, packed, None, out, stream=s.ptr)
        e0, e1 = L.Event(), L.Event()
        e0.record(s.ptr)
        for _ in range(20):
            ctx.conv2d_fwd(d, x, packed, None, out, stream=s.ptr)
        e1.record(s.ptr)
        s.sync()
        res[tag] = e0.elapsed_ms(e1) / 20 * 1e3
    ctx.debug_set(0)
    flops = 2.0 * 9 * cin * cout * n * h * w
    print(f"{name:24s} " + "  ".join(f"{k}: {v:7.1f} us ({flops / v / 1e6:6.0f} TF/s)" for k, v in res.items()), flush=True)
