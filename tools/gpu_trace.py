"""[needs a development build: make -C simplesr_b200/csrc clean all EXTRA=-DSSR_DEV]
Dev tool (run under gpurun): per-role timeline of CTA 0 of the conv kernel for the dense-block shapes."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simplesr_b200 import _lib as L  # noqa: E402
from simplesr_b200 import model_builder as MB  # noqa: E402

model = MB.build_enhanced_resnet(upsample_factor=4, num_rrdb_blocks=1, seed=1)
model.sync_weights()
ctx, s = model.ctx, model.stream.ptr
n, h, w = 16, 128, 128
px = n * h * w
src = L.DeviceBuffer(px * 192 * 2)
dst = L.DeviceBuffer(px * 192 * 2)
src.zero(s)
tr = L.DeviceBuffer(3 * 512 * 8)
flags = int(sys.argv[2]) if len(sys.argv) > 2 else 0
ctx.debug_set(flags)
names = sys.argv[1].split(",") if len(sys.argv) > 1 else ["rrdb0_db0_conv0", "rrdb0_db0_out"]
for name in names:
    c = model.convs[name]
    last = name.endswith("out")
    k = 0 if last else int(name[-1])
    d = L.ConvDesc(n=n, h=h, w=w, cin=c.cin, in_cstride=192, cout=c.cout, ksize=3,
                   act=(L.ACT_NONE if last else L.ACT_LRELU), act_alpha=0.2, res_beta=0.2, up=1,
                   out_dtype=L.SSR_BF16, out_cstride=192, out_coff=(0 if last else 64 + 32 * k),
                   res_dtype=(L.SSR_BF16 if last else L.SSR_NONE), res_cstride=192, res_coff=0,
                   out2_cstride=0, out2_coff=0)
    tgt = dst if last else src
    for _ in range(2):
        ctx.conv2d_fwd(d, src, c.d_packed, c.d_bias, tgt, res=(src if last else None), stream=s)
    tr.zero(s)
    ctx.debug_trace(tr)
    ctx.conv2d_fwd(d, src, c.d_packed, c.d_bias, tgt, res=(src if last else None), stream=s)
    ctx.debug_trace(None)
    L.stream_sync(s)
    t = tr.download((3, 512), np.int64, s)
    t0 = t[t > 0].min()
    rel = np.where(t > 0, t - t0, -1)
    print(f"== {name} (cycles since first event)")
    print("TMA issue  :", rel[0][:24].tolist())
    for it in range(8):
        m = rel[1][4 * it:4 * it + 4].tolist()
        e = rel[2][4 * it:4 * it + 4].tolist()
        print(f"tile {it}: MMA issue start/end {m} | EPI start/tfull_ok/tmem_done/stores_done {e}")
