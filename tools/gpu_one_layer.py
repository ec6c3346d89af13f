"""Dev tool: run each dense-block conv shape a few times (target for ncu)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simplesr_b200 import _lib as L  # noqa: E402
from simplesr_b200 import model_builder as MB  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
model = MB.build_enhanced_resnet(upsample_factor=4, num_rrdb_blocks=1, seed=1)
model.sync_weights()
ctx, s = model.ctx, model.stream.ptr
h = w = 128
px = n * h * w
src = L.DeviceBuffer(px * 192 * 2)
dst = L.DeviceBuffer(px * 192 * 2)
src.zero(s)
for k, name in enumerate(["rrdb0_db0_conv0", "rrdb0_db0_conv1", "rrdb0_db0_conv2", "rrdb0_db0_conv3", "rrdb0_db0_out"]):
    c = model.convs[name]
    last = name.endswith("out")
    d = L.ConvDesc(n=n, h=h, w=w, cin=c.cin, in_cstride=192, cout=c.cout, ksize=3,
                   act=(L.ACT_NONE if last else L.ACT_LRELU), act_alpha=0.2, res_beta=0.2, up=1,
                   out_dtype=L.SSR_BF16, out_cstride=192, out_coff=(0 if last else 64 + 32 * k),
                   res_dtype=(L.SSR_BF16 if last else L.SSR_NONE), res_cstride=192, res_coff=0,
                   out2_cstride=0, out2_coff=0)
    for _ in range(reps):
        ctx.conv2d_fwd(d, src, c.d_packed, c.d_bias, dst if last else src, res=(src if last else None), stream=s)
L.stream_sync(s)
print("done")
