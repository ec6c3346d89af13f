"""Dev tool (run under gpurun): MMA issue-rate probes (M, N, A swizzle)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simplesr_b200 import _lib as L  # noqa: E402

ctx = L.Context(0)
out = {}
for m in (128, 64):
    for swz in (2, 4, 6):
        for n in (16, 32, 64, 96, 128, 192, 256):
            try:
                v = ctx.diag_mma_rate_ex(m, n, swz, 8192)
            except Exception as e:  # noqa: BLE001
                v = str(e)
            out[f"m{m}_swz{swz}_n{n}"] = v
            print(f"M={m} swz={swz} N={n}: {v}", flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/probe2.json", "w"), indent=1)
