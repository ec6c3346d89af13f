"""Dev tool (run under gpurun): time the dense-block conv shapes for several forced tile widths / batch sizes."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simplesr_b200 import _lib as L  # noqa: E402
from simplesr_b200 import model_builder as MB  # noqa: E402

model = MB.build_enhanced_resnet(upsample_factor=4, num_rrdb_blocks=1, seed=1)
model.sync_weights()
ctx, s = model.ctx, model.stream.ptr
wbs = [int(a) for a in sys.argv[1].split(",")] if len(sys.argv) > 1 else [0]
batches = [int(a) for a in sys.argv[2].split(",")] if len(sys.argv) > 2 else [16]
for n in batches:
    h = w = 128
    px = n * h * w
    src = L.DeviceBuffer(px * 192 * 2)
    dst = L.DeviceBuffer(px * 192 * 2)
    src.zero(s)
    for wb in wbs:
        ctx.debug_set(0, wb)
        row = []
        for k, name in enumerate(["rrdb0_db0_conv0", "rrdb0_db0_conv1", "rrdb0_db0_conv2", "rrdb0_db0_conv3",
                                  "rrdb0_db0_out"]):
            c = model.convs[name]
            last = name.endswith("out")
            d = L.ConvDesc(n=n, h=h, w=w, cin=c.cin, in_cstride=192, cout=c.cout, ksize=3,
                           act=(L.ACT_NONE if last else L.ACT_LRELU), act_alpha=0.2, res_beta=0.2, up=1,
                           out_dtype=L.SSR_BF16, out_cstride=192, out_coff=(0 if last else 64 + 32 * k),
                           res_dtype=(L.SSR_BF16 if last else L.SSR_NONE), res_cstride=192, res_coff=0,
                           out2_cstride=0, out2_coff=0)
            tgt = dst if last else src
            run = lambda: ctx.conv2d_fwd(d, src, c.d_packed, c.d_bias, tgt, res=(src if last else None), stream=s)
            for _ in range(3):
                run()
            e0, e1 = L.Event(), L.Event()
            e0.record(s)
            for _ in range(20):
                run()
            e1.record(s)
            e1.sync()
            ms = e0.elapsed_ms(e1) / 20
            fl = 2.0 * 9 * c.cin_real * c.cout * px
            row.append(f"{c.cin_real}->{c.cout}: {ms*1e3:7.1f}us {fl/ms/1e9:6.0f}TF")
        print(f"n={n} wb={wb}: " + " | ".join(row), flush=True)
    src.free()
    dst.free()
