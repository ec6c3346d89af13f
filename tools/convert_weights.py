"""Inspect / convert generator weight files on any machine (no GPU, no TensorFlow, no h5py).

    python tools/convert_weights.py --info rrdb_gen_120.h5          # layer table + the architecture the shapes imply
    python tools/convert_weights.py rrdb_gen_120.h5 rrdb_gen_120.npz  # Keras HDF5 (sr_model.py:244)  ->  self-describing .npz
    python tools/convert_weights.py gen.npz gen.h5                   # .npz of GeneratorModel.save    ->  Keras HDF5 layout

Both formats load through ``build_or_load_generator_model(pretrained_model_path=...)`` /
``GeneratorModel.load_weights``; the HDF5 side is ``simplesr_b200/h5lite.py`` + ``keras_h5.py``.
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simplesr_b200 import h5lite, keras_h5  # noqa: E402


def _npz_config(arch, kw, residual_scaling=0.2):
    if arch == "rrdb":
        return dict(num_filters=kw["num_filters"], num_rrdb_blocks=kw["num_rrdb_blocks"],
                    num_dense_blocks=kw["num_dense_blocks"], num_convs=kw["num_convs"],
                    residual_scaling_factor=residual_scaling, input_dims=[None, None])
    return dict(num_filters=kw["num_filters"], num_res_blocks=kw["num_res_blocks"], batch_norm=kw["batch_normalization"],
                momentum=kw["momentum"], input_dims=[None, None])


def info(path, out=sys.stdout):
    layers, meta = h5lite.load_keras_weights(path)
    for k in ("keras_version", "backend"):
        if k in meta:
            print(f"{k}: {meta[k]}", file=out)
    n_params = 0
    for lname, ws in layers:
        for wname, arr in ws:
            n_params += arr.size
            print(f"  {lname:32s} {wname:48s} {tuple(arr.shape)}", file=out)
    print(f"{len(layers)} layers, {sum(1 for _, ws in layers if ws)} with weights, {n_params} parameters", file=out)
    try:
        arch, kw = keras_h5.infer_generator(layers, meta.get("model_config"))
        print(f"architecture: {arch} {kw}", file=out)
    except ValueError as e:
        print(f"architecture: not a SimpleSR generator ({e})", file=out)


def h5_to_npz(src, dst, residual_scaling=0.2):
    arch, kw, trainable, moving = keras_h5.read_generator_file(src)
    layers, _ = h5lite.load_keras_weights(src)
    names = [w for _, ws in layers for w, _ in ws if keras_h5._kind(w) not in keras_h5._MOVING] + \
            [w for _, ws in layers for w, _ in ws if keras_h5._kind(w) in keras_h5._MOVING]
    arrays = {f"{i:04d}|{n}": a for i, (n, a) in enumerate(zip(names, trainable + moving))}
    with open(dst, "wb") as f:
        np.savez(f, __architecture__=arch, __upsample_factor__=kw["upsample_factor"],
                 __config__=json.dumps(_npz_config(arch, kw, residual_scaling)), **arrays)
    return arch, kw


def npz_to_h5(src, dst):
    with np.load(src) as z:
        keys = sorted(k for k in z.files if "|" in k)
        named = [(k.split("|", 1)[1], z[k]) for k in keys]
    h5lite.save_keras_weights(dst, keras_h5.variables_to_layers(named), under_model_weights=True)


def main(argv):
    if len(argv) == 2 and argv[0] == "--info":
        info(argv[1])
        return 0
    if len(argv) != 2:
        print(__doc__)
        return 2
    src, dst = argv
    if keras_h5.is_hdf5(src) and dst.endswith(".npz"):
        arch, kw = h5_to_npz(src, dst)
        print(f"{src} -> {dst}: {arch} {kw}")
    elif src.endswith(".npz") and dst.endswith((".h5", ".hdf5")):
        npz_to_h5(src, dst)
        print(f"{src} -> {dst}")
    else:
        print("give one Keras .h5 and one .npz path")
        return 2
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
