"""Profiling target (run under ncu through gpurun): the launches of one RRDB dense block as the C2 plan issues them at
batch 16 x 128x128 - fea, growth pair 0 (64->32 | carry), tail 1, growth pair 2 (128->32 | carry), tail 3,
192->64 + residual (CTA pair) - three rounds, in network order."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simplesr_b200 import _lib as L  # noqa: E402
from simplesr_b200 import model_builder as MB  # noqa: E402

model = MB.build_enhanced_resnet(upsample_factor=4, num_rrdb_blocks=1, seed=1)
if len(sys.argv) > 1:
    model.fuse_growth = bool(int(sys.argv[1]))
plan = model.plan(16, 128, 128)
s = model.stream.ptr
for op in plan.ops:          # whole 1-block network once (fills every buffer)
    op(s)
for _ in range(3):
    for op in plan.ops[1:7]:
        op(s)
L.stream_sync(s)
print("done")
