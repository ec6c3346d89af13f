"""Where does a conv launch's fixed cost go?  globaltimer stamps of CTA 0 over consecutive launches of the RRDB trunk
(eager launches with PDL, as in the captured graph): exit of launch k -> entry of k+1 -> prologue -> previous grid
complete -> first box -> first MMA -> last tile stored -> exit.  Needs a library built with the boundary stamps compiled in:
make -C simplesr_b200/csrc clean all EXTRA=-DSSR_DEV   (the stamps cost the hot kernels 2-3 % even when idle, so the
default build leaves them out).   usage: python tools/gpu_boundary.py [n h w [fuse]]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simplesr_b200 import _lib as L  # noqa: E402
from simplesr_b200 import model_builder as MB  # noqa: E402

shape = tuple(int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (16, 128, 128)
fuse = int(sys.argv[4]) if len(sys.argv) > 4 else 1
model = MB.build_enhanced_resnet(upsample_factor=4, num_rrdb_blocks=2, seed=1)
model.fuse_growth = bool(fuse)
plan = model.plan(*shape)
s = model.stream.ptr
SLOTS = 24
tr = L.DeviceBuffer(SLOTS * 1536 * 8)
for _ in range(3):
    for op in plan.ops:
        op(s)
L.stream_sync(s)
tr.zero(s)
model.ctx.debug_trace_ring(tr, SLOTS)
for op in plan.ops[:2 + 2 * 3 * 5]:       # pad kernel, fea conv, then the dense blocks
    op(s)
model.ctx.debug_trace(None)
L.stream_sync(s)
t = tr.download((SLOTS, 3, 512), np.int64, s)
g = t[:, 0, 500:508].astype(np.float64)      # [slot][event] ns
names = ["entry", "prologue", "dep_done", "first_box", "first_mma", "stored_g0", "stored_g1", "exit"]
print("launch | gap from previous exit | " + " | ".join(names[1:]) + "   (us after entry)")
prev_exit = None
for k in range(SLOTS):
    if g[k, 0] == 0:
        continue
    gap = (g[k, 0] - prev_exit) / 1e3 if prev_exit else float("nan")
    rel = [(g[k, j] - g[k, 0]) / 1e3 if g[k, j] > 0 else float("nan") for j in range(1, 8)]
    print(f"{k:3d} | {gap:7.2f} | " + " | ".join(f"{v:7.2f}" for v in rel))
    prev_exit = g[k, 7]
