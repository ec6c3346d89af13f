"""Dev tool (run under gpurun): MMA issue-rate probes (M, N; CTA pairs)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simplesr_b200 import _lib as L  # noqa: E402

ctx = L.Context(0)
out = {}
for n in (32, 64, 96, 128, 192, 256):
    try:
        v = ctx.diag_mma_rate_pair(n, 8192)
    except Exception as e:  # noqa: BLE001
        v = (str(e), None)
    out[f"pair_n{n}"] = v
    print(f"2-CTA M=256 N={n}: complete {v[0]} issue {v[1]} cycles/MMA", flush=True)
for m in (128, 64):
    for n in (16, 32, 64, 96, 128, 192, 256):
        v = ctx.diag_mma_rate_ex(m, n, 2, 8192)
        out[f"m{m}_n{n}"] = v
        print(f"1-CTA M={m} N={n}: complete {v[0]:.2f} issue {v[1]:.2f} cycles/MMA", flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/probe2.json", "w"), indent=1)
