"""Does splitting a latency-bound batch over concurrent streams pay?  RRDB-23 forward at training size (32x32 LR):
one plan of batch 16 (160 pixel tiles = 2 waves over 148 SMs per layer) against k plans of batch 16/k launched as
CUDA graphs on k streams.  usage: python tools/gpu_split_streams.py [reps]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simplesr_b200 import _lib as L  # noqa: E402
from simplesr_b200 import model_builder as MB  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
for hw in (32, 24):
    for parts in (1, 2, 4):
        n = 16 // parts
        models = [MB.build_enhanced_resnet(upsample_factor=4, num_rrdb_blocks=23, seed=1) for _ in range(parts)]
        for m in models:
            m.fuse_growth = False          # what the trainers run
        plans = [m.plan(n, hw, hw) for m in models]
        x = np.random.default_rng(0).uniform(0, 1, size=(n, hw, hw, 3)).astype(np.float32)
        for m, p in zip(models, plans):
            L.check(m.ctx.lib.ssr_memcpy_h2d(p.buffers["in_f32"].ptr, x.ctypes.data, x.nbytes, m.stream.ptr))
            for _ in range(3):
                p.run(m.stream.ptr)
            m.stream.sync()
        main = models[0].stream
        e0, e1 = L.Event(), L.Event()
        joins = [L.Event() for _ in models]
        e0.record(main.ptr)
        for _ in range(reps):
            fork = L.Event()
            fork.record(main.ptr)
            for m, p, j in zip(models, plans, joins):
                if m is not models[0]:
                    m.stream.wait_event(fork)
                p.run(m.stream.ptr)
                if m is not models[0]:
                    j.record(m.stream.ptr)
                    main.wait_event(j)
        e1.record(main.ptr)
        main.sync()
        print(f"LR {hw}x{hw}: {parts} stream(s) x batch {n}: {e0.elapsed_ms(e1) / reps:.4f} ms per 16 images", flush=True)
        for m in models:
            m.release()
