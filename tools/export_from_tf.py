"""Pins conv / generator parity against the REAL reference: run this where TensorFlow 2.x and bw0248/SimpleSR are
importable (neither is in the build image - DESIGN.md section 2 says "parity unpinned" for that reason).

    python tools/export_from_tf.py /path/to/SimpleSR  [out.npz]

For each of the two generators it builds the reference's own Keras model through its own builder
(simple_sr/utils/models/model_builder.py: build_resnet :99, build_enhanced_resnet :42) at a small size, fills the
weights with seeded values, runs ONE forward pass on a seeded input and stores

    <tag>/input, <tag>/output                      NHWC float32
    <tag>/weights/<index>|<variable name>          model.get_weights() in Keras order
    <tag>/config                                   json of the builder arguments

plus tf.nn.depth_to_space and the reference's tiling round trip on a seeded image.  tests/test_tf_golden.py loads the
file from tests/golden/tf_reference_vectors.npz when it exists and checks the numpy oracle (CPU) and the CUDA path (GPU)
against it; without the file those tests are skipped.
"""
import json
import os
import sys

import numpy as np


def main():
    if len(sys.argv) < 2:
        sys.exit(__doc__)
    ref_root = os.path.abspath(sys.argv[1])
    out_path = sys.argv[2] if len(sys.argv) > 2 else os.path.join(
        os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "tf_reference_vectors.npz")
    sys.path.insert(0, ref_root)
    import tensorflow as tf
    from simple_sr.utils.models import model_builder as MB
    from simple_sr.utils.image import image_utils as IU

    rng = np.random.default_rng(20240)
    out = {}

    def dump(tag, model, cfg, x):
        ws = []
        for var, w in zip(model.weights, model.get_weights()):
            # keep the initialiser's scale (so activations stay O(1)) but make the values seed-reproducible
            std = float(np.std(w)) if np.std(w) > 0 else 0.05
            v = (rng.standard_normal(w.shape) * std).astype(np.float32)
            if "moving_variance" in var.name:
                v = (np.abs(v) + 0.5).astype(np.float32)
            elif "gamma" in var.name:
                v = (1.0 + v).astype(np.float32)
            ws.append(v)
        model.set_weights(ws)
        y = model(tf.constant(x), training=False).numpy()
        out[f"{tag}/input"], out[f"{tag}/output"] = x, y.astype(np.float32)
        out[f"{tag}/config"] = np.frombuffer(json.dumps(cfg).encode(), np.uint8)
        for i, (var, w) in enumerate(zip(model.weights, ws)):
            out[f"{tag}/weights/{i:04d}|{var.name}"] = w

    cfg = dict(upsample_factor=4, num_filters=64, num_res_blocks=2, batch_normalization=False)
    dump("srresnet", MB.build_resnet(**cfg), cfg, rng.uniform(0, 1, (1, 24, 20, 3)).astype(np.float32))
    cfg = dict(upsample_factor=4, num_filters=64, num_res_blocks=2, batch_normalization=True)
    dump("srresnet_bn", MB.build_resnet(**cfg), cfg, rng.uniform(0, 1, (1, 24, 20, 3)).astype(np.float32))
    cfg = dict(upsample_factor=4, num_filters=64, num_rrdb_blocks=1, num_dense_blocks=3, num_convs=4, kernel_size=3,
               residual_scaling_factor=0.2)
    dump("rrdb", MB.build_enhanced_resnet(**cfg), cfg, rng.uniform(0, 1, (2, 17, 33, 3)).astype(np.float32))

    x = rng.standard_normal((2, 5, 7, 16)).astype(np.float32)
    out["d2s/input"], out["d2s/output"] = x, tf.nn.depth_to_space(x, 2).numpy()

    img = rng.uniform(0, 1, (150, 201, 3)).astype(np.float32)
    ov = 16
    patches, padding = IU.segment_into_patches(tf.constant(img), patch_width=64, patch_height=64, pixel_overlap=ov)
    patches = tf.convert_to_tensor(patches)
    # same call as operations/evaluation.py:269-274 (scale 1)
    recon = IU.reconstruct_from_overlapping_patches(patches, image_height=img.shape[0], image_width=img.shape[1],
                                                    pixel_overlap=ov, horizontal_padding=padding[0][1] - ov,
                                                    vertical_padding=padding[1][1] - ov)
    out["tiling/input"] = img
    out["tiling/patches"] = np.asarray(patches, np.float32)
    out["tiling/padding"] = np.asarray(padding, np.int64)
    out["tiling/recon"] = np.asarray(recon, np.float32)

    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    np.savez_compressed(out_path, **out)
    print("wrote", out_path, os.path.getsize(out_path), "bytes; tensorflow", tf.__version__)


if __name__ == "__main__":
    main()
