"""Dev tool: per-op time of the full C2 plan (batch 16 x 128 x 128, 23 RRDB), each op timed in a 10-repeat graph."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simplesr_b200 import _lib as L, model_builder as MB
m = MB.build_enhanced_resnet(upsample_factor=4, num_rrdb_blocks=23, seed=1)
if len(sys.argv) > 1: m.fuse_growth = bool(int(sys.argv[1]))
plan = m.plan(16, 128, 128)
s = m.stream.ptr
plan.run(s) if hasattr(plan, "run") else [op(s) for op in plan.ops]
L.stream_sync(s)
ts = []
for i, op in enumerate(plan.ops):
    g = L.Graph(s, lambda: [op(s) for _ in range(10)])
    g.launch(s)
    e0, e1 = L.Event(), L.Event()
    e0.record(s)
    for _ in range(3): g.launch(s)
    e1.record(s); e1.sync()
    ts.append(e0.elapsed_ms(e1) / 30 * 1e3)
    g.destroy()
n = len(ts)
print("ops", n, "sum %.1f us" % sum(ts))
print("head", [round(t, 1) for t in ts[:8]])
print("rdb part sum %.1f us" % sum(ts[2:2 + 69 * 5]))
print("tail", [round(t, 1) for t in ts[2 + 69 * 5:]])
