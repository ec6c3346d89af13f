"""Keras HDF5 files <-> the variable lists of this package (host logic only, no device).

The reference's weight files are Keras HDF5: the generator as ``model.save("<type>_gen_<epoch>.h5")``
(simple_sr/models/sr_model.py:233-244), read back with ``tf.keras.models.load_model(path)``
(simple_sr/utils/models/model_builder.py:17-19); VGG19 through ``load_weights(custom_weights_path)`` on a ``.h5``
(model_builder.py:222,249).  Both hold one group per Keras layer in ``model.layers`` order with that layer's weights
in creation order.  For the two generator graphs ``model.layers`` order is creation order (every weighted layer sits on
the one chain from input to output, so Keras' depth sort keeps it), which is the order of
``GeneratorModel.trainable_variables`` here: per conv ``[kernel, bias]``, then BatchNormalization
``[gamma, beta, moving_mean, moving_variance]``, then PReLU ``[alpha]``.

Weights are matched by POSITION and shape, never by name: Keras auto-names the layers (``conv2d_17``,
``p_re_lu_3``, ...), old VGG19 files name weights ``block1_conv1_W_1:0``.
"""
import json
import math

import numpy as np

from . import h5lite

_MOVING = ("moving_mean", "moving_variance")


def _kind(weight_name):
    """'kernel' / 'bias' / 'gamma' / ... from 'conv2d_3/kernel:0'."""
    return weight_name.rsplit("/", 1)[-1].split(":")[0]


def variables_to_layers(named_arrays):
    """[(variable name, array)] in ``model.variables`` order (trainable first, then the moving statistics) ->
    Keras layers [(layer name, [(weight name, array)])] in creation order.  A layer is the part of the variable name
    before '/'; BatchNormalization layers get their moving statistics appended after gamma / beta as Keras stores them."""
    layers, index = [], {}
    for name, arr in named_arrays:
        lname = name.split("/", 1)[0]
        if lname not in index:
            index[lname] = len(layers)
            layers.append((lname, []))
        arr = np.asarray(arr, np.float32)
        if _kind(name) == "alpha" and arr.ndim == 1:
            arr = arr.reshape(1, 1, -1)                      # PReLU(shared_axes=[1, 2]) keeps (1, 1, C) in Keras
        layers[index[lname]][1].append((name, arr))
    return layers


def layers_to_weight_list(layers):
    """Keras layers -> (trainable arrays in creation order, moving statistics in creation order): the argument order of
    ``GeneratorModel.set_weights(trainable + moving)``."""
    trainable, moving = [], []
    for _, ws in layers:
        for wname, arr in ws:
            arr = np.asarray(arr, np.float32)
            if _kind(wname) == "alpha":
                arr = arr.reshape(-1)                        # (1, 1, C) in Keras, one slope per channel here
            (moving if _kind(wname) in _MOVING else trainable).append(arr)
    return trainable, moving


def _batchnorm_momentum(model_config, default=0.8):
    """momentum of the first BatchNormalization layer in a Keras ``model_config`` JSON (the weights do not carry it)."""
    if not model_config:
        return default
    try:
        cfg = json.loads(model_config)
        for layer in cfg["config"]["layers"]:
            if layer.get("class_name") == "BatchNormalization":
                return float(layer["config"]["momentum"])
    except (ValueError, KeyError, TypeError):
        pass
    return default


def infer_generator(layers, model_config=None):
    """Recover the builder arguments of a generator from its weights.

    Returns ``(architecture, kwargs)`` for ``build_resnet`` / ``build_enhanced_resnet``.  PReLU slopes mark the SRResNet
    graph (model_builder.py:99-134; the RRDB graph uses LeakyReLU, :42-96).  In the RRDB graph only the TOTAL number of
    dense blocks matters - ``_rrdb_block`` adds no residual of its own (:344-351) - so the split into
    ``num_rrdb_blocks x num_dense_blocks`` is reported as ``total / 3 x 3`` when divisible (the reference default) and
    ``total x 1`` otherwise; ``residual_scaling_factor`` lives in a Lambda layer and is not in the file.
    """
    kinds = [[_kind(w) for w, _ in ws] for _, ws in layers]
    kernels = [a for (_, ws) in layers for (w, a) in ws if _kind(w) == "kernel" and np.ndim(a) == 4]
    if len(kernels) < 4:
        raise ValueError(f"not a generator weight file: {len(kernels)} convolution kernels")
    has_alpha = any("alpha" in k for k in kinds)
    nf = int(kernels[0].shape[3])
    if kernels[0].shape[2] != 3 or kernels[-1].shape[3] != 3:
        raise ValueError("not a generator weight file: first / last convolution do not map 3 <-> num_filters channels")
    n_up = sum(1 for k in kernels if k.shape[3] == 4 * nf and k.shape[2] == nf)
    if n_up not in (1, 2, 3):
        raise ValueError(f"cannot infer the upsample factor: {n_up} sub-pixel convolutions")
    sf = 2 ** n_up
    if has_alpha:
        if kernels[0].shape[0] != 9:
            raise ValueError("SRResNet weight file expected a 9x9 first convolution")
        body = len(kernels) - 3 - n_up                       # first, trunk, last
        if body < 0 or body % 2:
            raise ValueError(f"cannot infer num_res_blocks from {len(kernels)} convolutions")
        bn = any("gamma" in k for k in kinds)
        return "srresnet", dict(upsample_factor=sf, num_filters=nf, num_res_blocks=body // 2, batch_normalization=bn,
                                momentum=_batchnorm_momentum(model_config))
    gc = int(kernels[1].shape[3])
    if 2 * gc != nf:
        raise ValueError(f"RRDB growth channels {gc} != num_filters // 2 = {nf // 2}")
    num_convs = 0
    while 1 + num_convs < len(kernels) and kernels[1 + num_convs].shape[3] == gc \
            and kernels[1 + num_convs].shape[2] == nf + num_convs * gc:
        num_convs += 1
    body = len(kernels) - 4 - n_up                           # fea, trunk, hr, last
    if num_convs == 0 or body <= 0 or body % (num_convs + 1):
        raise ValueError(f"cannot infer the dense-block structure from {len(kernels)} convolutions")
    total = body // (num_convs + 1)
    ndb = 3 if total % 3 == 0 else 1
    return "rrdb", dict(upsample_factor=sf, num_filters=nf, num_rrdb_blocks=total // ndb, num_dense_blocks=ndb,
                        num_convs=num_convs, kernel_size=int(kernels[1].shape[0]))


def read_generator_file(path):
    """``(architecture, builder kwargs, trainable arrays, moving statistics)`` of a Keras generator ``.h5``."""
    layers, meta = h5lite.load_keras_weights(path)
    arch, kwargs = infer_generator(layers, meta.get("model_config"))
    trainable, moving = layers_to_weight_list(layers)
    return arch, kwargs, trainable, moving


def write_model_file(path, variables, model_config=None):
    """``model.save(path)`` in the Keras HDF5 weight layout (weights only: the graph itself is rebuilt from the shapes by
    :func:`infer_generator`; a Keras ``model_config`` is written only when the caller supplies one).  ``variables``:
    objects with ``.name`` and ``.numpy()`` in ``model.variables`` order."""
    layers = variables_to_layers([(v.name, v.numpy()) for v in variables])
    # Keras order inside a BatchNormalization layer: gamma, beta, moving_mean, moving_variance - already the case since
    # the moving statistics come last in model.variables and are appended to their layer's list
    h5lite.save_keras_weights(path, layers, model_config=model_config, under_model_weights=True)
    return path


def read_vgg19_file(path, conv_names):
    """The 32 arrays ``[kernel, bias] x 16`` of a Keras VGG19 weight file (``vgg19_weights_tf_dim_ordering_tf_kernels
    [_notop].h5`` or a ``save_weights`` of the custom copy), matched by layer name ``blockN_convM`` where the file has
    those names and by position otherwise; the classifier head (``fc1`` ...) is ignored."""
    layers, _ = h5lite.load_keras_weights(path)
    by_name = {ln: ws for ln, ws in layers}
    if all(n in by_name for n in conv_names):
        picked = [by_name[n] for n in conv_names]
    else:
        picked = [ws for _, ws in layers if len(ws) == 2 and np.ndim(ws[0][1]) == 4][:len(conv_names)]
    if len(picked) != len(conv_names):
        raise ValueError(f"{path}: {len(picked)} convolution layers, VGG19 has {len(conv_names)}")
    out = []
    for name, ws in zip(conv_names, picked):
        if len(ws) != 2:
            raise ValueError(f"{path}: layer {name} holds {len(ws)} weights, expected kernel and bias")
        k, b = (np.asarray(a, np.float32) for _, a in ws)
        if k.ndim != 4:
            k, b = b, k
        out += [k, b]
    return out


def is_hdf5(path):
    """True when ``path`` names an HDF5 file (by suffix, or by signature when the file exists)."""
    path = str(path)
    if path.endswith((".h5", ".hdf5", ".keras.h5")):
        return True
    try:
        with open(path, "rb") as fh:
            return fh.read(8) == h5lite.SIGNATURE
    except OSError:
        return False
