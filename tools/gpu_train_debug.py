import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.helpers import O, rel_err
from tests.test_gpu_train_step import _setup_rrdb
from simplesr_b200.training import RRDBTrainer
nb, sf = 1, 2
m, params = _setup_rrdb(nb, sf)
rng = np.random.default_rng(0)
n, h, w = 2, 12, 10
lr = rng.uniform(0, 1, size=(n, h, w, 3)).astype(np.float32)
hr = rng.uniform(-1, 1, size=(n, h * sf, w * sf, 3)).astype(np.float32)
tr = RRDBTrainer(m, loss=[("mse", 1.0), ("mae", 0.1)], learning_rate=0.0)
out = tr.train_step(lr, hr, use_graph=False)
kw = dict(upsample_factor=sf, num_rrdb_blocks=nb, w_mse=1.0, w_mae=0.1)
loss32, sr32, g32 = O.rrdb_loss_and_grads(params, lr, hr, **kw)
loss16, sr16, g16 = O.rrdb_loss_and_grads(params, lr, hr, act_dtype="bf16", **kw)
print(out, loss32, loss16)
got = tr.gradients()
for name in g32:
    for i, kind in enumerate(("kernel", "bias")):
        print(f"{name:16s} {kind:6s} e32={rel_err(got[name][i], g32[name][i]):.4f} e16={rel_err(got[name][i], g16[name][i]):.4f} o16v32={rel_err(g16[name][i], g32[name][i]):.4f} max_ref={np.abs(g32[name][i]).max():.3e} max_got={np.abs(got[name][i]).max():.3e}")
