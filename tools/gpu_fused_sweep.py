"""Dev tool: time the paired growth-conv launches against the plain ones (batch 16 x 128 x 128)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simplesr_b200 import _lib as L, model_builder as MB
m = MB.build_enhanced_resnet(upsample_factor=4, num_rrdb_blocks=1, seed=1)
plan = m.plan(16, 128, 128)
s = m.stream.ptr
ops = plan.ops
def time_ops(ops, names):
    tot = 0.0
    for i, nm in enumerate(names):
        op = ops[i]
        g = L.Graph(s, lambda: [op(s) for _ in range(20)])
        for _ in range(2): g.launch(s)
        e0, e1 = L.Event(), L.Event()
        e0.record(s)
        for _ in range(5): g.launch(s)
        e1.record(s); e1.sync()
        us = e0.elapsed_ms(e1) / 100 * 1e3
        tot += us
        print(f"{nm}: {us:.1f} us", flush=True)
        g.destroy()
    print(f"  sum {tot:.1f} us")
names = ["f32->bf16", "fea", "pair0", "tail1", "pair2", "tail3", "out"]
time_ops(ops, names)
m2 = MB.build_enhanced_resnet(upsample_factor=4, num_rrdb_blocks=1, seed=1)
m2.fuse_growth = False
plan2 = m2.plan(16, 128, 128)
time_ops(plan2.ops, ["f32->bf16", "fea", "g0", "g1", "g2", "g3", "out"])
tr = L.DeviceBuffer(3 * 512 * 8)
for idx in (6, 3):
    tr.zero(s)
    m.ctx.debug_trace(tr)
    ops[idx](s)
    m.ctx.debug_trace(None)
    L.stream_sync(s)
    t = tr.download((3, 512), np.int64, s)
    t0 = t[t > 0].min()
    rel = np.where(t > 0, t - t0, -1)
    print(f"== {names[idx]}: TMA issue", rel[0][:20].tolist())
    for it in range(10):
        print(f"tile {it}: MMA start/end {rel[1][4*it:4*it+2].tolist()} | EPI start/tfull/tmem_done/done {rel[2][4*it:4*it+4].tolist()} | g0 math/store g1 math/store {rel[2][256+4*it:256+4*it+4].tolist()}")
