import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.helpers import O, rel_err
from tests.test_gpu_vgg import _vgg_pair
from simplesr_b200 import vgg as V
model, params = _vgg_pair()
rng = np.random.default_rng(1)
for layer, (n, h, w) in [("block2_conv2", (2, 32, 48)), ("block3_conv4", (2, 32, 32)), ("block4_conv4", (2, 32, 32)), ("block5_conv4", (2, 32, 32))]:
    hr = rng.uniform(-1, 1, size=(n, h, w, 3)).astype(np.float32)
    sr = np.clip(hr + rng.normal(0, 0.2, size=hr.shape), -1, 1).astype(np.float32)
    lf = V.VGGLoss(output_layers=layer, loss_weight=0.5, after_activation=False, vgg=model)
    loss, g = lf.loss_and_grad(hr, sr)
    l32, g32 = O.vgg_loss_and_grad(params, hr, sr, output_layer=layer, loss_weight=0.5)
    l16, g16 = O.vgg_loss_and_grad(params, hr, sr, output_layer=layer, loss_weight=0.5, act_dtype="bf16")
    cos = lambda a, b: float((a * b).sum() / (np.linalg.norm(a) * np.linalg.norm(b)))
    print(f"{layer}: loss {loss:.5g} / {l32:.5g} / {l16:.5g} | maxrel gpu-16 {rel_err(g, g16):.3f} gpu-32 {rel_err(g, g32):.3f} 16-32 {rel_err(g16, g32):.3f} | cos gpu-16 {cos(g, g16):.4f} gpu-32 {cos(g, g32):.4f} 16-32 {cos(g16, g32):.4f}", flush=True)
    lf.release()
