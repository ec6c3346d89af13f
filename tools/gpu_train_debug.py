import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.helpers import O, rel_err
from tests.test_gpu_train_step import _setup
from simplesr_b200.training import SRResNetTrainer
nb, sf = 2, 2
m, params = _setup(nb, sf)
rng = np.random.default_rng(0)
n, h, w = 2, 12, 10
lr = rng.uniform(0, 1, size=(n, h, w, 3)).astype(np.float32)
hr = rng.uniform(-1, 1, size=(n, h * sf, w * sf, 3)).astype(np.float32)
tr = SRResNetTrainer(m, loss=("mse", 1.0), learning_rate=0.0)
out = tr.train_step(lr, hr, use_graph=False)
loss32, sr32, g32 = O.srresnet_loss_and_grads(params, lr, hr, upsample_factor=sf, num_res_blocks=nb)
loss16, sr16, g16 = O.srresnet_loss_and_grads(params, lr, hr, upsample_factor=sf, num_res_blocks=nb, act_dtype="bf16")
print(out, loss32, loss16)
got = tr.gradients()
for name in g32:
    for i, kind in enumerate(("kernel", "bias", "alpha")):
        if g32[name][i] is None: continue
        print(f"{name:12s} {kind:6s} e32={rel_err(got[name][i], g32[name][i]):.4f} e16={rel_err(got[name][i], g16[name][i]):.4f} o16v32={rel_err(g16[name][i], g32[name][i]):.4f} max={np.abs(g32[name][i]).max():.3e}")
