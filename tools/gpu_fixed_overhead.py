"""Dev tool: per-launch fixed cost of the conv layers = intercept of time vs tiles per CTA (batch 4/8/16/32)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simplesr_b200 import _lib as L, model_builder as MB
names = ["f32->bf16", "fea", "g0", "g1", "g2", "g3", "out"]
res = {}
for fuse in (False, True):
    for b in (4, 8, 16, 32):
        m = MB.build_enhanced_resnet(upsample_factor=4, num_rrdb_blocks=1, seed=1)
        m.fuse_growth = fuse
        m.ctx.debug_set(int(os.environ.get('SSR_DBG', '0')))
        plan = m.plan(b, 128, 128)
        s = m.stream.ptr
        for i, nm in enumerate(names):
            op = plan.ops[i]
            g = L.Graph(s, lambda: [op(s) for _ in range(20)])
            for _ in range(2): g.launch(s)
            e0, e1 = L.Event(), L.Event()
            e0.record(s)
            for _ in range(5): g.launch(s)
            e1.record(s); e1.sync()
            res[(fuse, nm, b)] = e0.elapsed_ms(e1) / 100 * 1e3
            if os.environ.get('VERBOSE'): print(fuse, b, nm, res[(fuse, nm, b)], flush=True)
            g.destroy()
        m.release()
    print("fused" if fuse else "plain")
    for nm in names[1:]:
        t = [res[(fuse, nm, b)] for b in (4, 8, 16, 32)]
        slope = (t[3] - t[1]) / 24.0
        print(f"  {nm}: " + " ".join(f"{x:6.1f}" for x in t) + f"  us | per image {slope:.3f} us, intercept {t[1] - 8 * slope:.1f} us")
