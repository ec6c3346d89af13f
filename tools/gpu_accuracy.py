"""Dev tool (run under gpurun): end-to-end accuracy of the full-depth generators against the fp32 / bf16 oracles."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ssr_oracle as O  # noqa: E402
from simplesr_b200 import model_builder as MB  # noqa: E402

rng = np.random.default_rng(0)
x = rng.uniform(0, 1, size=(1, 32, 32, 3)).astype(np.float32)
for nb in (3, 23):
    params = O.init_rrdb_params(seed=1, bias_std=0.05, upsample_factor=4, num_rrdb_blocks=nb)
    m = MB.build_enhanced_resnet(upsample_factor=4, num_rrdb_blocks=nb, seed=0)
    w = []
    for name, _, _ in O.rrdb_layer_specs(upsample_factor=4, num_rrdb_blocks=nb):
        w.extend(params[name])
    m.set_weights(w)
    got = m(x)
    r32 = O.rrdb_forward(params, x, upsample_factor=4, num_rrdb_blocks=nb)
    r16 = O.rrdb_forward(params, x, upsample_factor=4, num_rrdb_blocks=nb, act_dtype="bf16")
    print(f"RRDB-{nb}: psnr(gpu,f32)={float(O.psnr(got, r32).min()):.2f} psnr(gpu,bf16orc)={float(O.psnr(got, r16).min()):.2f} "
          f"psnr(bf16orc,f32)={float(O.psnr(r16, r32).min()):.2f} rel={np.abs(got-r32).max()/np.abs(r32).max():.2e}", flush=True)
    m.release()
for nb in (2, 16):
    params = O.init_srresnet_params(seed=1, bias_std=0.05, alpha_std=0.15, upsample_factor=4, num_res_blocks=nb)
    m = MB.build_resnet(upsample_factor=4, num_res_blocks=nb, seed=0)
    w = []
    for name, *_ in O.srresnet_layer_specs(upsample_factor=4, num_res_blocks=nb):
        k, b, a = params[name]
        w.extend([k, b] + ([a] if a is not None else []))
    m.set_weights(w)
    got = m(x)
    r32 = O.srresnet_forward(params, x, upsample_factor=4, num_res_blocks=nb)
    r16 = O.srresnet_forward(params, x, upsample_factor=4, num_res_blocks=nb, act_dtype="bf16")
    print(f"SRResNet-{nb}: psnr(gpu,f32)={float(O.psnr(got, r32).min()):.2f} psnr(gpu,bf16orc)={float(O.psnr(got, r16).min()):.2f} "
          f"psnr(bf16orc,f32)={float(O.psnr(r16, r32).min()):.2f} rel={np.abs(got-r32).max()/np.abs(r32).max():.2e}", flush=True)
    m.release()
