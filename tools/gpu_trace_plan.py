"""[needs a development build: make -C simplesr_b200/csrc clean all EXTRA=-DSSR_DEV]
Per-role timeline (clock64 of CTA 0) of chosen launches of the C2 plan: usage
python tools/gpu_trace_plan.py <op index in the plan> [<op index> ...]   (2 = pair0, 3 = tail1, 4 = pair2, 5 = tail3, 6 = 192->64)"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simplesr_b200 import _lib as L  # noqa: E402
from simplesr_b200 import model_builder as MB  # noqa: E402

model = MB.build_enhanced_resnet(upsample_factor=4, num_rrdb_blocks=1, seed=1)
plan = model.plan(16, 128, 128)
s = model.stream.ptr
for _ in range(2):
    for op in plan.ops:
        op(s)
L.stream_sync(s)
tr = L.DeviceBuffer(3 * 512 * 8)
for idx in [int(a) for a in sys.argv[1:]] or [3]:
    for op in plan.ops[:idx]:
        op(s)
    tr.zero(s)
    model.ctx.debug_trace(tr)
    plan.ops[idx](s)
    model.ctx.debug_trace(None)
    L.stream_sync(s)
    t = tr.download((3, 512), np.int64, s)
    ev = t[:, :500]
    t0 = ev[ev > 0].min()
    rel = np.where(ev > 0, ev - t0, -1)
    print(f"== plan op {idx} (cycles since first event)")
    print("TMA issue  :", rel[0][:20].tolist())
    for it in range(16):
        m = rel[1][4 * it:4 * it + 2].tolist()
        e = rel[2][4 * it:4 * it + 4].tolist()
        print(f"tile {it:2d}: MMA issue start/end {m} | EPI start/tfull_ok/tmem_done/stores_done {e}")
