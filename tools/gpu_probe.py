"""First-contact GPU probe (run under gpurun): tcgen05 issue rate vs N and conv parity on a few shapes,
with both UMMA-descriptor base_offset conventions.  Not a pytest file."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.helpers import L, conv_case, rel_err  # noqa: E402

out = {}
ctx = L.Context(0)
print("SMs", ctx.sm_count, flush=True)
for n in (16, 32, 64, 128, 256):
    out[f"mma_cycles_n{n}"] = ctx.diag_mma_rate(n, 8192)
    print("mma cycles/instr N=%d: %.2f" % (n, out[f"mma_cycles_n{n}"]), flush=True)

cases = [
    dict(n=1, h=16, w=16, cin_real=64, cout=32),
    dict(n=2, h=20, w=37, cin_real=96, cout=32, act=L.ACT_LRELU),
    dict(n=1, h=33, w=18, cin_real=192, cout=64, res=True),
    dict(n=1, h=12, w=12, cin_real=64, cout=256, up=2, act=L.ACT_LRELU),
    dict(n=1, h=24, w=24, cin_real=64, cout=3, act=L.ACT_TANH, out_dtype=L.SSR_F32),
    dict(n=1, h=16, w=16, cin_real=3, cout=64),
]
for flags in (0, 1):
    for wb in (0, 8):
        ctx.debug_set(flags, wb)
        for i, c in enumerate(cases):
            t0 = time.time()
            try:
                got, ref, _ = conv_case(ctx, **c)
                e = rel_err(got, ref)
            except Exception as ex:  # noqa: BLE001
                e = f"EXC {ex}"
            out[f"flags{flags}_wb{wb}_case{i}"] = e
            print(f"flags={flags} wb={wb} case{i} {c}: rel_err={e}  ({time.time()-t0:.2f}s)", flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/probe.json", "w"), indent=1)
