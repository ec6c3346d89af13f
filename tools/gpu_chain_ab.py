"""A/B of tile-level dependencies (ssr_conv_chain_*) on the C2 step and at training size: ms per forward pass, chained vs not.
usage: python tools/gpu_chain_ab.py [reps]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simplesr_b200 import _lib as L  # noqa: E402
from simplesr_b200 import model_builder as MB  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
for shape in ((16, 128, 128), (16, 32, 32), (2, 32, 32)):
    for tag, opts in (("chain", dict(chain_deps=True)), ("grid", dict(chain_deps=False)),
                      ("chain+snake", dict(chain_deps=True, chain_snake=True)), ("chain", dict(chain_deps=True)),
                      ("grid", dict(chain_deps=False))):
        m = MB.build_enhanced_resnet(upsample_factor=4, num_rrdb_blocks=23, seed=1)
        for k, v in opts.items():
            setattr(m, k, v)
        plan = m.plan(*shape)
        s = m.stream.ptr
        x = np.random.default_rng(0).uniform(0, 1, size=(*shape, 3)).astype(np.float32)
        L.check(m.ctx.lib.ssr_memcpy_h2d(plan.buffers["in_f32"].ptr, x.ctypes.data, x.nbytes, s))
        for _ in range(5):
            plan.run(s)
        e0, e1 = L.Event(), L.Event()
        e0.record(s)
        for _ in range(reps):
            plan.run(s)
        e1.record(s)
        m.stream.sync()
        ms = e0.elapsed_ms(e1) / reps
        print(f"{shape} {tag:12s} {ms:8.4f} ms  chain_stats={plan.chain_stats} timeouts={plan.chain_timeouts()}", flush=True)
        m.release()
