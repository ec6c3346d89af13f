import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simplesr_b200 import _lib as L, model_builder as MB, evaluation as EV
m = MB.build_enhanced_resnet(upsample_factor=4, num_rrdb_blocks=23, seed=1)
s = m.stream.ptr
for nb in (16, 8):
    plan = m.plan(nb, 192, 192)
    for _ in range(2): plan.run(s)
    m.stream.sync()
    e0, e1 = L.Event(), L.Event()
    e0.record(s)
    for _ in range(3): plan.run(s)
    e1.record(s); e1.sync()
    print(f"plan({nb},192,192): {e0.elapsed_ms(e1)/3:.2f} ms per batch", flush=True)
img = np.random.default_rng(0).uniform(0, 1, size=(1024, 1024, 3)).astype(np.float32)
pin = L.PinnedArray((4096, 4096, 3), np.float32)
for i in range(3):
    t0 = time.perf_counter()
    EV.upscale_tiled(m, img, tile_batch=16, out=pin.array)
    print(f"upscale_tiled call {i}: {(time.perf_counter()-t0)*1e3:.1f} ms", flush=True)
t0 = time.perf_counter(); b = L.DeviceBuffer(201326592); b.zero(s); m.stream.sync(); print("malloc+zero", (time.perf_counter()-t0)*1e3)
t0 = time.perf_counter(); L.check(m.ctx.lib.ssr_memcpy_d2h(pin.ptr, b.ptr, 201326592, s)); m.stream.sync(); print("d2h pinned", (time.perf_counter()-t0)*1e3)
out = np.empty((4096, 4096, 3), np.float32)
t0 = time.perf_counter(); L.check(m.ctx.lib.ssr_memcpy_d2h(out.ctypes.data, b.ptr, 201326592, s)); m.stream.sync(); print("d2h pageable", (time.perf_counter()-t0)*1e3)
t0 = time.perf_counter(); b.free(); print("free", (time.perf_counter()-t0)*1e3)
