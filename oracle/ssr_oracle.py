"""CPU oracle for the SimpleSR hot path — TEST INFRASTRUCTURE ONLY.

This module is a numpy restatement of the arithmetic the reference (bw0248/SimpleSR, TensorFlow 2.2)
executes on its generator / tiled-inference path.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it; the product package
``simplesr_b200`` never does.

Parity status
  * tiling / stitching (``segment_into_patches``, ``reconstruct_from_overlapping_patches``,
    ``reconstruct_from_patches``): PINNED against the reference's own known-answer fixtures
    (tests/utils/image/test_image_utils.py:16-30,69-111: baboon/comic/lena PNGs and the 3x3 / 5x3
    matrices) in tests/test_oracle_tiling.py.
  * depth_to_space: pinned by TensorFlow's documented NHWC (DCR) definition and the hand example in
    tests/test_oracle_ops.py.
  * conv / generator graphs / losses: PARITY UNPINNED — the reference holds no golden tensors or saved
    weights for them and TensorFlow cannot be installed here.  They follow the
    TensorFlow semantics listed in SURVEY.md §9 and are cross-checked against torch.nn.functional on
    CPU (an independent implementation) in tests/test_oracle_ops.py.

Every function cites the reference file:line it follows (paths relative to the reference repo).
"""
import math

import numpy as np

# ----------------------------------------------------------------------------------------------
# bf16 helpers (the CUDA path stores activations / weights in bf16; the oracle can emulate that)
# ----------------------------------------------------------------------------------------------


def bf16_round(a):
    """fp32 -> bf16 (round to nearest even) -> fp32."""
    u = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16).astype(np.uint32) << 16
    return r.view(np.float32).reshape(np.shape(a))


def _q(a, act_dtype):
    return bf16_round(a) if act_dtype == "bf16" else np.asarray(a, dtype=np.float32)


# ----------------------------------------------------------------------------------------------
# TensorFlow op semantics
# ----------------------------------------------------------------------------------------------


def same_padding(in_size, k, s):
    """TF 'SAME': out = ceil(in/s); pad_total = max((out-1)*s + k - in, 0); begin = pad_total // 2."""
    out = -(-in_size // s)
    total = max((out - 1) * s + k - in_size, 0)
    return out, total // 2, total - total // 2


def conv2d_same(x, kernel, bias=None, stride=1):
    """tf.keras.layers.Conv2D(padding='same') + BiasAdd — model_builder.py:285-290.

    x: [N,H,W,Cin] fp32, kernel: HWIO [kh,kw,Cin,Cout], cross-correlation.
    """
    x = np.asarray(x, dtype=np.float32)
    kernel = np.asarray(kernel, dtype=np.float32)
    n, h, w, cin = x.shape
    kh, kw, kcin, cout = kernel.shape
    assert kcin == cin, (kcin, cin)
    oh, pt, pb = same_padding(h, kh, stride)
    ow, pl, pr = same_padding(w, kw, stride)
    xp = np.zeros((n, h + pt + pb, w + pl + pr, cin), dtype=np.float32)
    xp[:, pt:pt + h, pl:pl + w, :] = x
    y = np.zeros((n * oh * ow, cout), dtype=np.float32)
    for i in range(kh):
        for j in range(kw):
            xs = xp[:, i:i + (oh - 1) * stride + 1:stride, j:j + (ow - 1) * stride + 1:stride, :]
            y += xs.reshape(-1, cin) @ kernel[i, j]
    y = y.reshape(n, oh, ow, cout)
    if bias is not None:
        y = y + np.asarray(bias, dtype=np.float32)
    return y


def leaky_relu(x, alpha=0.2):
    """tf.keras.layers.LeakyReLU(alpha) — model_builder.py:85,90,335."""
    return np.where(x > 0, x, np.float32(alpha) * x).astype(np.float32)


def prelu(x, alpha):
    """PReLU(shared_axes=[1,2]): max(0,x) + alpha_c*min(0,x) — model_builder.py:118,281,314."""
    return (np.maximum(x, 0) + np.asarray(alpha, dtype=np.float32) * np.minimum(x, 0)).astype(np.float32)


def depth_to_space(x, block=2):
    """tf.nn.depth_to_space NHWC (DCR): out[n,h*b+i,w*b+j,c] = in[n,h,w,(i*b+j)*C+c] — model_builder.py:279."""
    n, h, w, c4 = x.shape
    c = c4 // (block * block)
    y = x.reshape(n, h, w, block, block, c).transpose(0, 1, 3, 2, 4, 5)
    return np.ascontiguousarray(y.reshape(n, h * block, w * block, c))


# ----------------------------------------------------------------------------------------------
# initialisers — RNG streams cannot match TF's; parity is always same-weights
# ----------------------------------------------------------------------------------------------


def he_normal_scaled(rng, shape, scale=0.2):
    """he_normal() with .scale overwritten to 0.2 (model_builder.py:60-61): VarianceScaling(scale, fan_in,
    truncated_normal): stddev = sqrt(scale/fan_in)/.87962566103423978, samples truncated at 2 stddev."""
    fan_in = int(np.prod(shape[:-1]))
    std = math.sqrt(scale / fan_in) / 0.87962566103423978
    out = rng.standard_normal(size=shape)
    bad = np.abs(out) > 2.0
    while bad.any():
        out[bad] = rng.standard_normal(size=int(bad.sum()))
        bad = np.abs(out) > 2.0
    return (out * std).astype(np.float32)


def glorot_uniform(rng, shape):
    """Keras default kernel initializer (SRResNet convs, model_builder.py:287 with initializer=None)."""
    rf = int(np.prod(shape[:-2]))
    fan_in, fan_out = shape[-2] * rf, shape[-1] * rf
    limit = math.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-limit, limit, size=shape).astype(np.float32)


# ----------------------------------------------------------------------------------------------
# RRDB generator — model_builder.build_enhanced_resnet (model_builder.py:42-96)
# ----------------------------------------------------------------------------------------------


def rrdb_layer_specs(upsample_factor=4, num_filters=64, num_rrdb_blocks=16, num_dense_blocks=3, num_convs=4):
    """(name, cin, cout) for every conv in Keras creation order == trainable_variables order [kernel, bias]."""
    if upsample_factor not in (2, 4, 8):
        raise ValueError("upsample factor not supported - please choose either 2, 4 or 8")  # model_builder.py:63-64
    nf, gc = num_filters, num_filters // 2
    specs = [("fea", 3, nf)]
    for b in range(num_rrdb_blocks):
        for d in range(num_dense_blocks):
            for k in range(num_convs):
                specs.append((f"rrdb{b}_db{d}_conv{k}", nf + k * gc, gc))
            specs.append((f"rrdb{b}_db{d}_out", nf + num_convs * gc, nf))
    specs.append(("trunk", nf, nf))
    for u in range(int(math.log(upsample_factor, 2))):
        specs.append((f"up{u}", nf, nf * 4))
    specs.append(("hr", nf, nf))
    specs.append(("last", nf, 3))
    return specs


def init_rrdb_params(seed=1, bias_std=0.0, **kw):
    """Synthetic weights with the reference's initialiser (he_normal, scale 0.2; biases zero unless bias_std)."""
    rng = np.random.default_rng(seed)
    params = {}
    for name, cin, cout in rrdb_layer_specs(**kw):
        k = he_normal_scaled(rng, (3, 3, cin, cout))
        b = (rng.standard_normal(cout) * bias_std).astype(np.float32) if bias_std else np.zeros(cout, np.float32)
        params[name] = (k, b)
    return params


def rrdb_forward(params, x, upsample_factor=4, num_rrdb_blocks=16, num_dense_blocks=3, num_convs=4,
                 residual_scaling=0.2, act_dtype="f32", taps=None):
    """Forward pass of build_enhanced_resnet (model_builder.py:42-96, 328-365).

    act_dtype="bf16" rounds weights and every stored activation to bf16 (what the CUDA path stores);
    "f32" is the reference's arithmetic.  ``taps`` (dict) receives named intermediate activations.
    """
    q = lambda a: _q(a, act_dtype)

    def conv(name, t):
        k, b = params[name]
        return conv2d_same(t, q(k), b)

    def tap(name, t):
        if taps is not None:
            taps[name] = t
        return t

    beta = np.float32(residual_scaling)
    x = q(x)
    fea = tap("fea", q(conv("fea", x)))                                        # :67-68
    r = fea
    for b in range(num_rrdb_blocks):                                           # :356-362
        for d in range(num_dense_blocks):                                      # :347-350
            prev = [r]
            t = r
            for k in range(num_convs):                                         # :333-338
                c = q(leaky_relu(conv(f"rrdb{b}_db{d}_conv{k}", t), 0.2))
                tap(f"rrdb{b}_db{d}_conv{k}", c)
                prev.append(c)
                t = np.concatenate(prev, axis=3)
            dd = conv(f"rrdb{b}_db{d}_out", t)                                 # :340
            r = tap(f"rrdb{b}_db{d}_out", q(r + beta * dd))                    # :349-350
    t2 = tap("trunk_in", q(fea + beta * r))                                    # :363-364
    u = tap("trunk", q(fea + conv("trunk", t2)))                               # :76-79
    for i in range(int(math.log(upsample_factor, 2))):                         # :81-85
        u = tap(f"up{i}", q(leaky_relu(depth_to_space(conv(f"up{i}", u), 2), 0.2)))
    u = tap("hr", q(leaky_relu(conv("hr", u), 0.2)))                           # :87-90
    return tap("last", np.tanh(conv("last", u)).astype(np.float32))            # :91-94


# ----------------------------------------------------------------------------------------------
# SRResNet generator — model_builder.build_resnet (model_builder.py:99-134), batch_norm=False
# ----------------------------------------------------------------------------------------------


def srresnet_layer_specs(upsample_factor=4, num_filters=64, num_res_blocks=16):
    if upsample_factor not in (2, 4, 8):
        raise ValueError("upsample factor not supported - please choose either 2, 4 or 8")  # model_builder.py:113-114
    nf = num_filters
    specs = [("first", 9, 3, nf, True)]                      # (name, ksize, cin, cout, has_prelu)
    for b in range(num_res_blocks):
        specs.append((f"res{b}_conv0", 3, nf, nf, True))
        specs.append((f"res{b}_conv1", 3, nf, nf, False))
    specs.append(("trunk", 3, nf, nf, False))
    for u in range(int(math.log(upsample_factor, 2))):
        specs.append((f"up{u}", 3, nf, nf * 4, True))
    specs.append(("last", 9, nf, 3, False))
    return specs


def init_srresnet_params(seed=1, bias_std=0.0, alpha_std=0.0, **kw):
    """glorot_uniform kernels, zero biases, PReLU alpha zeros (Keras defaults); *_std > 0 randomises them so the
    bias / alpha paths are exercised."""
    rng = np.random.default_rng(seed)
    params = {}
    for name, ks, cin, cout, has_prelu in srresnet_layer_specs(**kw):
        k = glorot_uniform(rng, (ks, ks, cin, cout))
        b = (rng.standard_normal(cout) * bias_std).astype(np.float32) if bias_std else np.zeros(cout, np.float32)
        a = None
        if has_prelu:
            ac = cout // 4 if name.startswith("up") else cout  # PReLU sits after depth_to_space (model_builder.py:279-281)
            a = (rng.standard_normal(ac) * alpha_std).astype(np.float32) if alpha_std else np.zeros(ac, np.float32)
        params[name] = (k, b, a)
    return params


BN_EPS = 1e-3   # tf.keras.layers.BatchNormalization default epsilon (model_builder.py:291-292 passes only momentum)


def init_srresnet_bn(seed=1, num_res_blocks=16, num_filters=64, randomize=False):
    """BatchNormalization variables of build_resnet(batch_normalization=True): one layer after each res-block conv and
    after the trunk conv (model_builder.py:309-319, 123-125).  Keras initial state: gamma 1, beta 0, moving_mean 0,
    moving_variance 1; ``randomize`` draws trained-looking values so that the inference fold is exercised."""
    rng = np.random.default_rng(seed + 977)
    names = [f"res{b}_conv{j}" for b in range(num_res_blocks) for j in (0, 1)] + ["trunk"]
    bn = {}
    for n in names:
        if randomize:
            bn[n] = dict(gamma=(1.0 + 0.2 * rng.standard_normal(num_filters)).astype(np.float32),
                         beta=(0.1 * rng.standard_normal(num_filters)).astype(np.float32),
                         mean=(0.1 * rng.standard_normal(num_filters)).astype(np.float32),
                         var=rng.uniform(0.5, 1.5, num_filters).astype(np.float32))
        else:
            bn[n] = dict(gamma=np.ones(num_filters, np.float32), beta=np.zeros(num_filters, np.float32),
                         mean=np.zeros(num_filters, np.float32), var=np.ones(num_filters, np.float32))
    return bn


def batch_norm_inference(t, bn):
    """BatchNormalization(training=False): gamma * (x - moving_mean) / sqrt(moving_variance + eps) + beta."""
    return (bn["gamma"] * (t - bn["mean"]) / np.sqrt(bn["var"] + BN_EPS) + bn["beta"]).astype(np.float32)


def srresnet_forward(params, x, upsample_factor=4, num_res_blocks=16, act_dtype="f32", taps=None, bn=None):
    """Forward pass of build_resnet (model_builder.py:99-134, 309-325); ``bn`` = init_srresnet_bn(...) for
    batch_normalization=True at inference (moving statistics)."""
    q = lambda a: _q(a, act_dtype)

    def conv(name, t):
        k, b, _ = params[name]
        y = conv2d_same(t, q(k), b)
        if bn is not None and name in bn:
            y = batch_norm_inference(y, bn[name])                                # :291-292
        return y

    def tap(name, t):
        if taps is not None:
            taps[name] = t
        return t

    x = q(x)
    t = tap("first", q(prelu(conv("first", x), params["first"][2])))            # :117-118
    skip = t
    for b in range(num_res_blocks):                                             # :309-319
        u = q(prelu(conv(f"res{b}_conv0", t), params[f"res{b}_conv0"][2]))
        t = tap(f"res{b}", q(t + conv(f"res{b}_conv1", u)))
    t = tap("trunk", q(conv("trunk", t) + skip))                                # :123-126
    for i in range(int(math.log(upsample_factor, 2))):                          # :128-131, 275-282
        t = tap(f"up{i}", q(prelu(depth_to_space(conv(f"up{i}", t), 2), params[f"up{i}"][2])))
    return tap("last", np.tanh(conv("last", t)).astype(np.float32))             # :133


# ----------------------------------------------------------------------------------------------
# tiling / stitching — simple_sr/utils/image/image_utils.py
# ----------------------------------------------------------------------------------------------


def segment_into_patches(tensor, patch_width=32, patch_height=32, pixel_overlap=0):
    """image_utils.segment_into_patches (image_utils.py:85-121). tensor: [H,W,C] or [1,H,W,C].
    Returns (patches [T,ph(+2ov),pw(+2ov),C], padding [[top,bottom],[left,right]])."""
    t = np.asarray(tensor)
    if t.ndim != 3 and not (t.ndim == 4 and t.shape[0] == 1):
        raise ValueError("Tensor must be of rank 3")
    if t.ndim == 4:
        t = t[0]
    if t.shape[0] < patch_height or t.shape[1] < patch_width:
        raise ValueError("Patch dimensions are larger than image size")          # :115-116
    if pixel_overlap != 0:
        return _segment_with_overlap(t, patch_width, patch_height, pixel_overlap)
    return _segment(t, patch_width, patch_height)


def _segment_with_overlap(t, patch_width, patch_height, pixel_overlap):
    """image_utils._segment_with_overlap (image_utils.py:124-148), including its swapped loop steps."""
    hp = [pixel_overlap, pixel_overlap]
    vp = [pixel_overlap, pixel_overlap]
    if t.shape[0] % patch_height != 0:
        hp[1] += (patch_height - t.shape[0]) % patch_height
    if t.shape[1] % patch_width != 0:
        vp[1] += (patch_width - t.shape[1]) % patch_width
    padded = np.pad(t, [hp, vp, [0, 0]], mode="constant", constant_values=0)
    patches = []
    for row in range(pixel_overlap, padded.shape[0] - pixel_overlap, patch_width):      # sic: steps by patch_width
        for col in range(pixel_overlap, padded.shape[1] - pixel_overlap, patch_height):  # sic: steps by patch_height
            x0, x1 = col - pixel_overlap, col + patch_width + pixel_overlap
            y0, y1 = row - pixel_overlap, row + patch_height + pixel_overlap
            patches.append(padded[y0:y1, x0:x1, :])
    return np.stack(patches), [hp, vp]


def _segment(t, patch_width, patch_height):
    """image_utils._segment (image_utils.py:151-164): space_to_batch + split + stack + reshape, i.e. row-major
    non-overlapping patches of the bottom/right zero-padded image."""
    hp = [0, 0]
    vp = [0, 0]
    if t.shape[0] % patch_height != 0:
        hp = [0, (patch_height - t.shape[0]) % patch_height]
    if t.shape[1] % patch_width != 0:
        vp = [0, (patch_width - t.shape[1]) % patch_width]
    padded = np.pad(t, [hp, vp, [0, 0]], mode="constant", constant_values=0)
    H, W, C = padded.shape
    gh, gw = H // patch_height, W // patch_width
    # space_to_batch([t], [ph,pw]): out[(i*pw+j), gy, gx, c] = padded[gy*ph+i, gx*pw+j, c]
    s2b = padded.reshape(gh, patch_height, gw, patch_width, C).transpose(1, 3, 0, 2, 4).reshape(
        patch_height * patch_width, gh, gw, C)
    # split along batch into ph*pw pieces of [1,gh,gw,C]; stack on axis 3 -> [1,gh,gw,ph*pw,C]; reshape
    stacked = np.stack([s2b[i:i + 1] for i in range(patch_height * patch_width)], axis=3)
    return stacked.reshape(-1, patch_height, patch_width, C), [hp, vp]


def _reconstruct(patches, original_height, original_width, padded_height, padded_width):
    """image_utils._reconstruct (image_utils.py:167-184): reshape / split / stack / batch_to_space / crop."""
    ph, pw, pc = patches.shape[1], patches.shape[2], patches.shape[3]
    gh, gw = padded_height // ph, padded_width // pw
    r = patches.reshape(1, gh, gw, ph * pw, pc)
    r = np.stack([r[:, :, :, i:i + 1, :] for i in range(ph * pw)], axis=0)      # tf.split(.., 3) + tf.stack(.., 0)
    r = r.reshape(ph * pw, gh, gw, pc)
    # batch_to_space(r, [ph,pw]): out[0, gy*ph+i, gx*pw+j, c] = r[i*pw+j, gy, gx, c]
    out = r.reshape(ph, pw, gh, gw, pc).transpose(2, 0, 3, 1, 4).reshape(gh * ph, gw * pw, pc)
    return out[0:original_height, 0:original_width, :]


def reconstruct_from_overlapping_patches(patches, image_height, image_width, pixel_overlap, horizontal_padding,
                                         vertical_padding):
    """image_utils.reconstruct_from_overlapping_patches (image_utils.py:40-61)."""
    patches = np.asarray(patches)
    if patches.ndim != 4:
        raise ValueError("Tensor with patches needs to be of rank 4")
    p = patches[:, pixel_overlap:-pixel_overlap, pixel_overlap:-pixel_overlap, :]
    return _reconstruct(p, image_height, image_width, image_height + horizontal_padding,
                        image_width + vertical_padding)


def reconstruct_from_patches(patches, original_height, original_width, horizontal_padding=0, vertical_padding=0):
    """image_utils.reconstruct_from_patches (image_utils.py:64-82)."""
    patches = np.asarray(patches)
    if patches.ndim != 4:
        raise ValueError("Tensor with patches needs to be of rank 4")
    if horizontal_padding < 0 or vertical_padding < 0:
        raise ValueError("Padding can't be negative")
    return _reconstruct(patches, original_height, original_width, original_height + horizontal_padding,
                        original_width + vertical_padding)


def eligible_efficient_inference(shape, min_width=1000, min_height=1000):
    """evaluation._eligible_efficient_inference (evaluation.py:340-348) on a shape tuple.
    (The reference names the pair (width, height) but compares both against both thresholds symmetrically.)"""
    if len(shape) not in (3, 4):
        return False
    if len(shape) == 4 and shape[0] != 1:
        return False
    a, b = (shape[1], shape[2]) if len(shape) == 4 else (shape[0], shape[1])
    return bool(a > min_width and b > min_height)


def tiled_upscale(model_fn, lr, scale, patch=128, pixel_overlap=32):
    """evaluate_on_testdata's tiled branch (evaluation.py:253-277) + _upscale (evaluation.py:351-359):
    segment -> model on every tile with batch 1 -> stitch.  lr: [1,H,W,3] or [H,W,3]."""
    lr = np.asarray(lr, dtype=np.float32)
    h, w = (lr.shape[1], lr.shape[2]) if lr.ndim == 4 else (lr.shape[0], lr.shape[1])
    tiles, padding = segment_into_patches(lr, patch_width=patch, patch_height=patch, pixel_overlap=pixel_overlap)
    sr = np.concatenate([model_fn(tiles[i:i + 1]) for i in range(tiles.shape[0])], axis=0)
    ovs = pixel_overlap * scale
    return reconstruct_from_overlapping_patches(
        sr, image_height=h * scale, image_width=w * scale, pixel_overlap=ovs,
        horizontal_padding=padding[0][1] * scale - ovs, vertical_padding=padding[1][1] * scale - ovs)


# ----------------------------------------------------------------------------------------------
# data preparation: tf.image.resize(bicubic, antialias) and the augmentations - simple_sr/data_pipeline/data_pipeline.py
# :318-330, simple_sr/utils/image/image_transforms.py:50-80, 157-173, 320-345
# ----------------------------------------------------------------------------------------------


def _keys_cubic(x):
    """Keys cubic convolution kernel, a = -0.5 (TensorFlow's "keyscubic" ScaleAndTranslate kernel)."""
    x = np.abs(np.asarray(x, np.float32))
    near = ((np.float32(1.5) * x - np.float32(2.5)) * x) * x + np.float32(1)
    far = ((np.float32(-0.5) * x + np.float32(2.5)) * x - np.float32(4)) * x + np.float32(2)
    return np.where(x >= 2, np.float32(0), np.where(x >= 1, far, near)).astype(np.float32)


def _resize_weights(in_size, out_size, antialias):
    """Dense [out_size, in_size] weight matrix of tf.image.resize(method="bicubic") along one axis: TensorFlow's
    ScaleAndTranslate spans (ComputeSpansCore) - sample position (o + 0.5) * in/out, support radius 2 * kernel_scale with
    kernel_scale = max(in/out, 1) when antialias else 1, indices clamped to the image, weights normalised to sum 1."""
    inv_scale = np.float32(in_size) / np.float32(out_size)
    ks = np.float32(max(float(inv_scale), 1.0)) if antialias else np.float32(1)
    mat = np.zeros((out_size, in_size), np.float32)
    for o in range(out_size):
        sample = (np.float32(o) + np.float32(0.5)) * inv_scale
        lo = max(int(np.ceil(sample - np.float32(2) * ks - np.float32(0.5))), 0)
        hi = min(int(np.floor(sample + np.float32(2) * ks - np.float32(0.5))), in_size - 1)
        j = np.arange(lo, hi + 1)
        w = _keys_cubic((j.astype(np.float32) + np.float32(0.5) - sample) / ks)
        tot = w.sum(dtype=np.float32)
        if abs(tot) > 1000 * np.finfo(np.float32).tiny:
            w = w / tot
        mat[o, lo:hi + 1] = w
    return mat


def resize_bicubic(x, scale, antialias=True):
    """tf.image.resize(x, (h / scale, w / scale), method="bicubic", antialias=antialias) on NHWC fp32 - what
    _prepare_img_pairs uses to synthesise the LR image (data_pipeline.py:318-330).  Columns first, then rows."""
    x = np.asarray(x, np.float32)
    n, h, w, c = x.shape
    wx, wy = _resize_weights(w, w // scale, antialias), _resize_weights(h, h // scale, antialias)
    tmp = np.einsum("ow,nhwc->nhoc", wx, x).astype(np.float32)
    return np.einsum("ph,nhoc->npoc", wy, tmp).astype(np.float32)


def prepare_img_pairs(hr_uint8, scale, antialias=True):
    """DataPipeline._prepare_img_pairs (data_pipeline.py:318-330, bicubic filter, no JPEG noise): LR in [0,1], HR in [-1,1]."""
    hr = np.asarray(hr_uint8, np.float32)
    return resize_bicubic(hr / np.float32(255), scale, antialias), hr / np.float32(127.5) - np.float32(1)


def flip_along_x(x):
    """image_transforms.flip_along_x (:320-331) = tf.image.flip_up_down."""
    return np.asarray(x)[..., ::-1, :, :]


def flip_along_y(x):
    """image_transforms.flip_along_y (:334-345) = tf.image.flip_left_right."""
    return np.asarray(x)[..., :, ::-1, :]


def rotate90(x, rotations):
    """image_transforms.rotate90 (:157-173) = tf.image.rot90: counter-clockwise quarter turns of the (H, W) axes."""
    x = np.asarray(x)
    return np.rot90(x, k=rotations, axes=(x.ndim - 3, x.ndim - 2))


def rgb_to_y(x):
    """Luma of tf.image.rgb_to_yuv."""
    x = np.asarray(x, np.float32)
    return (np.float32(0.299) * x[..., 0] + np.float32(0.587) * x[..., 1] + np.float32(0.114) * x[..., 2]).astype(np.float32)


def psnr_on_y(a, b, max_val=2.0):
    """metrics.psnr_on_y (metrics.py:18-44): PSNR of the Y channels, per image."""
    return psnr(rgb_to_y(a)[..., None], rgb_to_y(b)[..., None], max_val=max_val)


def ssim(a, b, max_val=2.0):
    """metrics.ssim (metrics.py:47-59) = tf.image.ssim: 11x11 Gaussian window (sigma 1.5, VALID), k1 = 0.01, k2 = 0.03,
    luminance * contrast-structure averaged over the window positions, then over the channels; per image."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    if a.ndim == 3:
        a, b = a[None], b[None]
    g = np.exp(-0.5 * (np.arange(11) - 5.0) ** 2 / 1.5 ** 2)
    g /= g.sum()

    def blur(t):
        n, h, w, c = t.shape
        rows = sum(g[k] * t[:, :, k:w - 10 + k] for k in range(11))
        return sum(g[k] * rows[:, k:h - 10 + k] for k in range(11))

    c1, c2 = (0.01 * max_val) ** 2, (0.03 * max_val) ** 2
    m0, m1 = blur(a), blur(b)
    num0, den0 = 2 * m0 * m1, m0 * m0 + m1 * m1
    lum = (num0 + c1) / (den0 + c1)
    cs = (2 * blur(a * b) - num0 + c2) / (blur(a * a + b * b) - den0 + c2)
    return np.mean(lum * cs, axis=(1, 2, 3)).astype(np.float32)


# ----------------------------------------------------------------------------------------------
# pixel losses / metrics
# ----------------------------------------------------------------------------------------------


def mean_squared_error(hr, sr):
    """tf.keras.losses.MeanSquaredError() — mean_squared_error.py:57 (global mean for equal shapes)."""
    d = np.asarray(hr, np.float32) - np.asarray(sr, np.float32)
    return np.float32(np.mean(d.astype(np.float64) ** 2))


def mean_absolute_error(hr, sr):
    """tf.keras.losses.MeanAbsoluteError() — mean_absolute_error.py:57."""
    d = np.asarray(hr, np.float32) - np.asarray(sr, np.float32)
    return np.float32(np.mean(np.abs(d.astype(np.float64))))


def psnr(a, b, max_val=2.0):
    """tf.image.psnr per image — metrics.py:4-15: 20 log10(max) - 10 log10(mean((a-b)^2)) over H,W,C."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    mse = np.mean((a - b) ** 2, axis=(-3, -2, -1))
    with np.errstate(divide="ignore"):
        return (20.0 * np.log10(max_val) - 10.0 * np.log10(mse)).astype(np.float32)


# ----------------------------------------------------------------------------------------------
# backward pass (what tf.GradientTape computes in SRModel.train_step, sr_model.py:419-441) and Keras Adam
# ----------------------------------------------------------------------------------------------


def conv2d_same_backward(x, kernel, dy):
    """Gradients of conv2d_same (stride 1): returns (dx, dkernel, dbias) — Conv2DBackpropInput / Conv2DBackpropFilter /
    BiasAddGrad of the tape backward through model_builder.py:285-290."""
    x = np.asarray(x, np.float32)
    kernel = np.asarray(kernel, np.float32)
    dy = np.asarray(dy, np.float32)
    n, h, w, cin = x.shape
    kh, kw, _, cout = kernel.shape
    _, pt, pb = same_padding(h, kh, 1)
    _, pl, pr = same_padding(w, kw, 1)
    xp = np.zeros((n, h + pt + pb, w + pl + pr, cin), np.float32)
    xp[:, pt:pt + h, pl:pl + w] = x
    dxp = np.zeros_like(xp)
    dk = np.zeros_like(kernel)
    dyf = dy.reshape(-1, cout)
    for i in range(kh):
        for j in range(kw):
            xs = xp[:, i:i + h, j:j + w, :].reshape(-1, cin)
            dk[i, j] = xs.T @ dyf
            dxp[:, i:i + h, j:j + w, :] += (dyf @ kernel[i, j].T).reshape(n, h, w, cin)
    return dxp[:, pt:pt + h, pl:pl + w], dk, dyf.sum(axis=0)


def space_to_depth(x, block=2):
    """Inverse of depth_to_space (its gradient): out[n,h,w,(i*b+j)*C+c] = in[n,h*b+i,w*b+j,c]."""
    n, hb, wb, c = x.shape
    h, w = hb // block, wb // block
    y = x.reshape(n, h, block, w, block, c).transpose(0, 1, 3, 2, 4, 5)
    return np.ascontiguousarray(y.reshape(n, h, w, block * block * c))


def adam_update(param, grad, m, v, t, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-7):
    """Keras OptimizerV2 Adam: returns (param, m, v) after step t (1-based)."""
    lr_t = np.float32(lr * math.sqrt(1.0 - beta2 ** t) / (1.0 - beta1 ** t))
    # Keras evaluates (1 - beta) in the variable dtype: float32(1) - float32(0.999) = 0.0010000467, not 0.001
    m = np.float32(beta1) * m + (np.float32(1) - np.float32(beta1)) * grad
    v = np.float32(beta2) * v + (np.float32(1) - np.float32(beta2)) * grad * grad
    return (param - lr_t * m / (np.sqrt(v) + np.float32(eps))).astype(np.float32), m, v


def srresnet_loss_and_grads(params, lr_batch, hr_batch, upsample_factor=4, num_res_blocks=16, act_dtype="f32", bn=None,
                            stats_out=None):
    """Forward + MSE loss + gradients of every variable for build_resnet: what train_step's generator tape yields with
    loss_functions=[MeanSquaredError()] (sr_model.py:419-441, generator.py:220-228).
    Returns (loss, sr, grads) with grads[name] = (dkernel, dbias, dalpha or None).
    ``bn`` (init_srresnet_bn) switches on batch_normalization=True in TRAINING mode: batch statistics (biased variance)
    after both convs of every res block and the trunk conv (model_builder.py:291-292, 309-319, 123-125); then
    grads[name + "_bn"] = (dgamma, dbeta) and ``stats_out[name] = (batch mean, biased batch variance)``.
    act_dtype="bf16" rounds weights and stored activations like the CUDA path (gradients stay fp32)."""
    q = lambda a: _q(a, act_dtype)
    nb = num_res_blocks
    nup = int(math.log(upsample_factor, 2))
    K = {n_: q(p[0]) for n_, p in params.items()}
    cache = {}

    def conv(name, t):
        cache[name + "/x"] = t
        z = conv2d_same(t, K[name], params[name][1])
        if bn is not None and name in bn:                                       # BatchNormalization(training=True)
            z = q(z)
            mu = z.mean(axis=(0, 1, 2), dtype=np.float64).astype(np.float32)
            var = z.var(axis=(0, 1, 2), dtype=np.float64).astype(np.float32)
            istd = (1.0 / np.sqrt(var + BN_EPS)).astype(np.float32)
            xhat = ((z - mu) * istd).astype(np.float32)
            cache[name + "/xhat"], cache[name + "/istd"] = xhat, istd
            if stats_out is not None:
                stats_out[name] = (mu, var)
            z = (bn[name]["gamma"] * xhat + bn[name]["beta"]).astype(np.float32)
        return z

    def prelu_f(name, z):
        cache[name + "/z"] = z
        return prelu(z, params[name][2])

    x = q(lr_batch)
    t = q(prelu_f("first", q(conv("first", x))))
    skip = t
    for b in range(nb):
        u = q(prelu_f(f"res{b}_conv0", q(conv(f"res{b}_conv0", t))))
        t = q(t + (q(conv(f"res{b}_conv1", u)) if bn is not None else conv(f"res{b}_conv1", u)))
    t = q((q(conv("trunk", t)) if bn is not None else conv("trunk", t)) + skip)
    for i in range(nup):
        z = q(depth_to_space(conv(f"up{i}", t), 2))
        t = q(prelu_f(f"up{i}", z))
    sr = np.tanh(conv("last", t)).astype(np.float32)
    diff = sr - np.asarray(hr_batch, np.float32)
    loss = np.float32(np.mean(diff.astype(np.float64) ** 2))

    grads = {}

    def conv_b(name, dz):
        if bn is not None and name in bn:
            xhat, istd = cache[name + "/xhat"], cache[name + "/istd"]
            m = dz.shape[0] * dz.shape[1] * dz.shape[2]
            dgamma = (dz * xhat).reshape(-1, dz.shape[-1]).sum(0)
            dbeta = dz.reshape(-1, dz.shape[-1]).sum(0)
            grads[name + "_bn"] = (dgamma.astype(np.float32), dbeta.astype(np.float32))
            dz = ((bn[name]["gamma"] * istd) * (dz - dbeta / m - xhat * (dgamma / m))).astype(np.float32)
        dx, dk, db = conv2d_same_backward(cache[name + "/x"], K[name], dz)
        grads[name] = [dk, db, None]
        return dx

    def prelu_b(name, dy):
        z = cache[name + "/z"]
        a = np.asarray(params[name][2], np.float32)
        grads[name][2] = (dy * np.minimum(z, 0)).reshape(-1, z.shape[-1]).sum(axis=0)
        return (dy * np.where(z > 0, np.float32(1), a)).astype(np.float32)

    d = (2.0 * diff / diff.size).astype(np.float32) * (1.0 - sr * sr)          # through MSE and tanh
    d = conv_b("last", d)
    for i in reversed(range(nup)):
        grads[f"up{i}"] = [None, None, None]
        dz = prelu_b(f"up{i}", d)
        dz = space_to_depth(dz, 2)
        a_keep = grads[f"up{i}"][2]
        d = conv_b(f"up{i}", dz)
        grads[f"up{i}"][2] = a_keep
    dskip = d
    d = conv_b("trunk", d)
    for b in reversed(range(nb)):
        du = conv_b(f"res{b}_conv1", d)
        grads[f"res{b}_conv0"] = [None, None, None]
        dz = prelu_b(f"res{b}_conv0", du)
        a_keep = grads[f"res{b}_conv0"][2]
        d = d + conv_b(f"res{b}_conv0", dz)
        grads[f"res{b}_conv0"][2] = a_keep
    d = d + dskip
    grads["first"] = [None, None, None]
    dz = prelu_b("first", d)
    a_keep = grads["first"][2]
    conv_b("first", dz)
    grads["first"][2] = a_keep
    return loss, sr, {k_: tuple(v_) for k_, v_ in grads.items()}


def rrdb_loss_and_grads(params, lr_batch, hr_batch, upsample_factor=4, num_rrdb_blocks=16, num_dense_blocks=3,
                        num_convs=4, residual_scaling=0.2, w_mse=1.0, w_mae=0.0, act_dtype="f32", extra_loss=None):
    """Forward + pixel loss (w_mse * MSE + w_mae * MAE) + gradients of every variable of build_enhanced_resnet
    (model_builder.py:42-96, 328-365): the generator tape of train_step (sr_model.py:419-441) for the "resnet" model
    type.  ``extra_loss(sr) -> (value, dvalue/dsr)`` adds further loss functors (generator.py:220-228), e.g. the VGG
    loss.  Returns (loss, sr, grads) with grads[name] = (dkernel, dbias)."""
    q = lambda a: _q(a, act_dtype)
    beta = np.float32(residual_scaling)
    nup = int(math.log(upsample_factor, 2))
    K = {n_: q(p[0]) for n_, p in params.items()}
    cache, grads = {}, {}

    def conv(name, t):
        cache[name] = t
        return conv2d_same(t, K[name], params[name][1])

    def conv_b(name, dz):
        dx, dk, db = conv2d_same_backward(cache[name], K[name], dz)
        grads[name] = (dk, db)
        return dx

    x = q(lr_batch)
    fea = q(conv("fea", x))
    r = fea
    blocks = []
    for b in range(num_rrdb_blocks):
        for d in range(num_dense_blocks):
            pre = f"rrdb{b}_db{d}"
            feats = [r]
            for k in range(num_convs):
                feats.append(q(leaky_relu(conv(f"{pre}_conv{k}", np.concatenate(feats, axis=3)), 0.2)))
            dd = conv(f"{pre}_out", np.concatenate(feats, axis=3))
            blocks.append((pre, feats))
            r = q(r + beta * dd)
    t_in = q(fea + beta * r)
    u = q(fea + conv("trunk", t_in))
    ups = []
    for i in range(nup):
        u = q(leaky_relu(depth_to_space(conv(f"up{i}", u), 2), 0.2))
        ups.append(u)
    hr_y = q(leaky_relu(conv("hr", u), 0.2))
    sr = np.tanh(conv("last", hr_y)).astype(np.float32)
    diff = sr - np.asarray(hr_batch, np.float32)
    loss = np.float32(w_mse * np.mean(diff.astype(np.float64) ** 2) + w_mae * np.mean(np.abs(diff.astype(np.float64))))

    lrelu_b = lambda dy, y: (dy * np.where(y > 0, np.float32(1), np.float32(0.2))).astype(np.float32)
    d = ((w_mse * 2.0 * diff + w_mae * np.sign(diff)) / diff.size).astype(np.float32)
    if extra_loss is not None:                               # further loss functors: (value, d value / d sr)
        ev, eg = extra_loss(sr)
        loss = np.float32(loss + ev)
        d = d + eg
    d = d * (1.0 - sr * sr)
    d = conv_b("last", d)
    d = conv_b("hr", lrelu_b(d, hr_y))
    for i in reversed(range(nup)):
        d = conv_b(f"up{i}", space_to_depth(lrelu_b(d, ups[i]), 2))
    g_fea = d.copy()
    d_ti = conv_b("trunk", d)
    g_fea += d_ti
    G = beta * d_ti
    nf = fea.shape[-1]
    for pre, feats in reversed(blocks):
        gbuf = beta * conv_b(f"{pre}_out", G)            # gradient of the concatenated [x | c1..c4]
        gbuf[..., :nf] += G
        for k in reversed(range(num_convs)):
            lo = nf + k * (feats[1].shape[-1])
            dz = lrelu_b(gbuf[..., lo:lo + feats[k + 1].shape[-1]], feats[k + 1])
            gbuf[..., :lo] += conv_b(f"{pre}_conv{k}", dz)
        # the out conv's weight/bias gradients carry the residual scaling
        dk, db = grads[f"{pre}_out"]
        grads[f"{pre}_out"] = (beta * dk, beta * db)
        G = gbuf[..., :nf]
    conv_b("fea", g_fea + G)
    return loss, sr, grads


# ----------------------------------------------------------------------------------------------
# VGG19 perceptual loss — model_builder.build_vgg_19 / _custom_vgg (:201-272) + vgg_loss.VGGLoss (:59-180)
# ----------------------------------------------------------------------------------------------

VGG19_LAYERS = [("block1_conv1", 3, 64), ("block1_conv2", 64, 64), ("block1_pool",),
                ("block2_conv1", 64, 128), ("block2_conv2", 128, 128), ("block2_pool",),
                ("block3_conv1", 128, 256), ("block3_conv2", 256, 256), ("block3_conv3", 256, 256),
                ("block3_conv4", 256, 256), ("block3_pool",),
                ("block4_conv1", 256, 512), ("block4_conv2", 512, 512), ("block4_conv3", 512, 512),
                ("block4_conv4", 512, 512), ("block4_pool",),
                ("block5_conv1", 512, 512), ("block5_conv2", 512, 512), ("block5_conv3", 512, 512),
                ("block5_conv4", 512, 512), ("block5_pool",)]
VGG_MEAN_BGR = np.array([103.939, 116.779, 123.68], np.float32)


def init_vgg19_params(seed=2, bias_std=0.05):
    """Synthetic VGG19 weights at the real shapes (He-normal; the ImageNet weights Keras downloads are not available
    offline - parity is always same-weights)."""
    rng = np.random.default_rng(seed)
    params = {}
    for layer in VGG19_LAYERS:
        if len(layer) == 3:
            name, cin, cout = layer
            k = (rng.standard_normal((3, 3, cin, cout)) * math.sqrt(2.0 / (9 * cin))).astype(np.float32)
            params[name] = (k, (rng.standard_normal(cout) * bias_std).astype(np.float32))
    return params


def vgg_preprocess(x):
    """(x + 1) * 127.5 then caffe-mode preprocess_input: RGB -> BGR, minus the ImageNet means (vgg_loss.py:144-148)."""
    x = (np.asarray(x, np.float32) + np.float32(1)) * np.float32(127.5)
    return (x[..., ::-1] - VGG_MEAN_BGR).astype(np.float32)


def maxpool2(x):
    n, h, w, c = x.shape
    return x[:, :h // 2 * 2, :w // 2 * 2].reshape(n, h // 2, 2, w // 2, 2, c).max(axis=(2, 4))


def maxpool2_backward(x, dy):
    """Gradient to the first maximum of every 2x2 window (row-major order)."""
    n, h, w, c = x.shape
    win = x[:, :h // 2 * 2, :w // 2 * 2].reshape(n, h // 2, 2, w // 2, 2, c).transpose(0, 1, 3, 5, 2, 4).reshape(
        n, h // 2, w // 2, c, 4)
    first = np.argmax(win, axis=-1)
    dwin = (np.arange(4) == first[..., None]) * dy[..., None]
    dx = np.zeros_like(x)
    dx[:, :h // 2 * 2, :w // 2 * 2] = dwin.reshape(n, h // 2, w // 2, c, 2, 2).transpose(0, 1, 4, 2, 5, 3).reshape(
        n, h // 2 * 2, w // 2 * 2, c)
    return dx


def vgg19_features(params, x_pre, output_layer="block5_conv4", after_activation=False, act_dtype="f32", cache=None):
    """Truncated custom VGG19 (vgg_loss.py:106-110): features of `output_layer`, before its ReLU unless
    after_activation.  x_pre: preprocessed BGR image."""
    q = lambda a: _q(a, act_dtype)
    t = q(x_pre)
    for layer in VGG19_LAYERS:
        name = layer[0]
        if len(layer) == 3:
            if cache is not None:
                cache[name + "/x"] = t
            z = conv2d_same(t, q(params[name][0]), params[name][1])
            if name == output_layer and not after_activation:
                return z.astype(np.float32)
            t = q(np.maximum(z, 0))
            if cache is not None:
                cache[name + "/y"] = t
            if name == output_layer:
                return t
        else:
            if cache is not None:
                cache[name + "/x"] = t
            t = maxpool2(t)
    raise ValueError(f"unknown layer {output_layer}")


def vgg_loss_and_grad(params, hr, sr, output_layer="block5_conv4", feature_scale=1.0, loss_weight=1.0,
                      act_dtype="f32"):
    """VGGLoss.__call__ (vgg_loss.py:115-180, denormalize=True, no TV term) and its gradient w.r.t. sr.
    Returns (loss, dsr)."""
    q = lambda a: _q(a, act_dtype)
    f_hr = vgg19_features(params, vgg_preprocess(hr), output_layer, act_dtype=act_dtype) * np.float32(feature_scale)
    cache = {}
    f_sr = vgg19_features(params, vgg_preprocess(sr), output_layer, act_dtype=act_dtype, cache=cache) * np.float32(
        feature_scale)
    diff = f_sr - f_hr
    loss = np.float32(np.mean(diff.astype(np.float64) ** 2) * loss_weight)
    d = (2.0 * loss_weight * feature_scale * diff / diff.size).astype(np.float32)
    started = False
    for layer in reversed(VGG19_LAYERS):
        name = layer[0]
        if not started:
            if name != output_layer:
                continue
            started = True
            d, _, _ = conv2d_same_backward(cache[name + "/x"], q(params[name][0]), d)
            continue
        if len(layer) == 3:
            d = d * (cache[name + "/y"] > 0)
            d, _, _ = conv2d_same_backward(cache[name + "/x"], q(params[name][0]), d)
        else:
            d = maxpool2_backward(cache[name + "/x"], d)
    dsr = (np.float32(127.5) * d[..., ::-1]).astype(np.float32)
    return loss, dsr


def total_variation(x):
    """tf.image.total_variation on a batch: per image, sum |x[i+1,j]-x[i,j]| + sum |x[i,j+1]-x[i,j]| (vgg_loss.py:167)."""
    x = np.asarray(x, np.float64)
    return np.abs(x[:, 1:] - x[:, :-1]).sum(axis=(1, 2, 3)) + np.abs(x[:, :, 1:] - x[:, :, :-1]).sum(axis=(1, 2, 3))


def total_variation_backward(x):
    """d sum_batch(total_variation(x)) / d x."""
    x = np.asarray(x, np.float64)
    g = np.zeros_like(x)
    sy, sx = np.sign(x[:, 1:] - x[:, :-1]), np.sign(x[:, :, 1:] - x[:, :, :-1])
    g[:, 1:] += sy
    g[:, :-1] -= sy
    g[:, :, 1:] += sx
    g[:, :, :-1] -= sx
    return g


def vgg_loss_general(params, hr, sr, output_layers, feature_scale=1.0, loss_weight=1.0, after_activation=True,
                     total_variation_loss=False, total_variation_weight=2 * 10e-8, act_dtype="f32"):
    """VGGLoss.__call__ in full (vgg_loss.py:115-180, denormalize=True): the sum over ``output_layers`` of
    MSE(features) * loss_weight (:162-164), features after the ReLU when ``after_activation`` (the stock Keras VGG19,
    :80-86) else before it (the custom network), plus total_variation_weight * reduce_sum(tf.image.total_variation(
    127.5 * (sr + 1))) (:166-169).  Returns (loss, d loss / d sr)."""
    q = lambda a: _q(a, act_dtype)
    layers = output_layers if isinstance(output_layers, list) else [output_layers]
    names = [l[0] for l in VGG19_LAYERS]
    deepest = max(layers, key=names.index)

    def forward(img, cache):
        t = q(vgg_preprocess(img))
        feats = {}
        for layer in VGG19_LAYERS:
            name = layer[0]
            if len(layer) == 3:
                cache[name + "/x"] = t
                z = conv2d_same(t, q(params[name][0]), params[name][1])
                if name in layers and not after_activation:
                    feats[name] = (z.astype(np.float32) if name == deepest else q(z))
                t = q(np.maximum(z, 0))
                cache[name + "/y"] = t
                if name in layers and after_activation:
                    feats[name] = t
                if name == deepest:
                    return feats
            else:
                cache[name + "/x"] = t
                t = maxpool2(t)
        raise ValueError(deepest)

    f_hr, cache = forward(hr, {}), {}
    f_sr = forward(sr, cache)
    loss, dfeat = 0.0, {}
    for name in layers:
        diff = (f_sr[name] - f_hr[name]) * np.float32(feature_scale)
        loss += float(np.mean(diff.astype(np.float64) ** 2) * loss_weight)
        dfeat[name] = (2.0 * loss_weight * feature_scale * diff / diff.size).astype(np.float32)
    d, started = None, False
    for layer in reversed(VGG19_LAYERS):
        name = layer[0]
        if not started:
            if name != deepest:
                continue
            started = True
        if len(layer) == 3:
            if name in layers and after_activation:
                d = dfeat[name] if d is None else d + dfeat[name]
            if d is not None and not (name == deepest and not after_activation):
                d = d * (cache[name + "/y"] > 0)
            if name in layers and not after_activation:
                d = dfeat[name] if d is None else d + dfeat[name]
            d, _, _ = conv2d_same_backward(cache[name + "/x"], q(params[name][0]), d)
        else:
            d = maxpool2_backward(cache[name + "/x"], d)
    dsr = (np.float32(127.5) * d[..., ::-1]).astype(np.float32)
    if total_variation_loss:
        den = (np.asarray(sr, np.float64) + 1.0) * 127.5
        loss += float(total_variation_weight * total_variation(den).sum())
        dsr = dsr + (total_variation_weight * 127.5 * total_variation_backward(den)).astype(np.float32)
    return np.float32(loss), dsr


# ----------------------------------------------------------------------------------------------
# Discriminator (model_builder.build_discriminator :137-198) and the relativistic-average GAN losses
# (ra_adversarial_loss.py:59-70, ra_discriminator_loss.py:55-66); train_step's GAN branch (sr_model.py:419-451)
# ----------------------------------------------------------------------------------------------

DISC_CONVS = [("d_conv0", 3, 64, 1, False), ("d_conv1", 64, 64, 2, True), ("d_conv2", 64, 128, 1, True),
              ("d_conv3", 128, 128, 2, True), ("d_conv4", 128, 256, 1, True), ("d_conv5", 256, 256, 2, True),
              ("d_conv6", 256, 512, 1, True), ("d_conv7", 512, 512, 2, True)]   # (name, cin, cout, stride, batch_norm)
BN_EPS = np.float32(1e-3)     # Keras BatchNormalization default epsilon


def init_discriminator_params(seed=3, input_hw=(128, 128), bias_std=0.0):
    """he_normal with scale 0.2 (model_builder.py:155-157); BN gamma = 1, beta = 0; Dense layers use the same
    initialiser, zero biases.  Flatten is NHWC row-major."""
    rng = np.random.default_rng(seed)
    p = {}
    for name, cin, cout, _, bn in DISC_CONVS:
        b = (rng.standard_normal(cout) * bias_std).astype(np.float32) if bias_std else np.zeros(cout, np.float32)
        p[name] = [he_normal_scaled(rng, (3, 3, cin, cout)), b]
        if bn:
            p[name + "_bn"] = [np.ones(cout, np.float32), np.zeros(cout, np.float32)]
    flat = (input_hw[0] // 16) * (input_hw[1] // 16) * 512
    p["d_dense0"] = [he_normal_scaled(rng, (flat, 1024)), np.zeros(1024, np.float32)]
    p["d_dense1"] = [he_normal_scaled(rng, (1024, 1)), np.zeros(1, np.float32)]
    return p


def discriminator_forward(params, x, cache=None, alpha=0.2, act_dtype="f32"):
    """Critic of a batch (training=True: BatchNormalization uses the batch statistics, biased variance)."""
    q = lambda a: _q(a, act_dtype)
    t = q(x)
    for name, _, _, stride, bn in DISC_CONVS:
        z = conv2d_same(t, q(params[name][0]), params[name][1], stride=stride)
        if cache is not None:
            cache[name + "/x"] = t
        if bn:
            z = q(z)
            mu = z.mean(axis=(0, 1, 2), dtype=np.float64).astype(np.float32)
            var = z.var(axis=(0, 1, 2), dtype=np.float64).astype(np.float32)
            xhat = (z - mu) / np.sqrt(var + BN_EPS)
            g, b = params[name + "_bn"]
            if cache is not None:
                cache[name + "/xhat"], cache[name + "/istd"] = xhat, 1.0 / np.sqrt(var + BN_EPS)
            z = g * xhat + b
        t = q(leaky_relu(z, alpha))
        if cache is not None:
            cache[name + "/y"] = t
    f = t.reshape(t.shape[0], -1)
    h = f @ params["d_dense0"][0] + params["d_dense0"][1]
    a = leaky_relu(h, alpha)
    c = a @ params["d_dense1"][0] + params["d_dense1"][1]
    if cache is not None:
        cache["flat"], cache["h"], cache["a"], cache["shape"] = f, h, a, t.shape
    return c.astype(np.float32)


def _conv_backward_strided(x, kernel, dz, stride):
    if stride == 1:
        return conv2d_same_backward(x, kernel, dz)
    # stride 2 on even sizes: out = full[1::2, 1::2] of the stride-1 SAME conv (TF pads (0, 1))
    n, h, w, _ = x.shape
    full = np.zeros((n, h, w, dz.shape[-1]), np.float32)
    full[:, 1::2, 1::2] = dz
    return conv2d_same_backward(x, kernel, full)


def discriminator_backward(params, cache, dcritic, alpha=0.2, act_dtype="f32"):
    """Gradients of sum(critic * dcritic): returns (dx, grads) with grads[name] = [dkernel, dbias] / [dgamma, dbeta]."""
    q = lambda a: _q(a, act_dtype)
    grads = {}
    a, h, f = cache["a"], cache["h"], cache["flat"]
    grads["d_dense1"] = [a.T @ dcritic, dcritic.sum(0)]
    da = dcritic @ params["d_dense1"][0].T
    dh = da * np.where(h > 0, np.float32(1), np.float32(alpha))
    grads["d_dense0"] = [f.T @ dh, dh.sum(0)]
    d = (dh @ params["d_dense0"][0].T).reshape(cache["shape"]).astype(np.float32)
    for name, _, _, stride, bn in reversed(DISC_CONVS):
        y = cache[name + "/y"]
        d = d * np.where(y > 0, np.float32(1), np.float32(alpha))
        if bn:
            xhat, istd = cache[name + "/xhat"], cache[name + "/istd"]
            g = params[name + "_bn"][0]
            m = d.shape[0] * d.shape[1] * d.shape[2]
            dgamma = (d * xhat).reshape(-1, d.shape[-1]).sum(0)
            dbeta = d.reshape(-1, d.shape[-1]).sum(0)
            grads[name + "_bn"] = [dgamma, dbeta]
            d = (g * istd) * (d - dbeta / m - xhat * (dgamma / m))
        d, dk, db = _conv_backward_strided(cache[name + "/x"], q(params[name][0]), d.astype(np.float32), stride)
        grads[name] = [dk, db]
    return d, grads


def _softplus(z):
    return np.maximum(z, 0) + np.log1p(np.exp(-np.abs(z)))


def _sigmoid(z):
    return 1.0 / (1.0 + np.exp(-z))


def ragan_losses(hr_critic, sr_critic, hr_label=1.0, sr_label=0.0):
    """RaAdversarialLoss (generator) and RaDiscriminatorLoss with their gradients w.r.t. both critics.
    BCE(y, z) from logits = mean(softplus(z) - y z).  Returns dict(g_loss, d_loss, g_dsr, d_dsr, d_dhr)."""
    hc = np.asarray(hr_critic, np.float64).ravel()
    sc = np.asarray(sr_critic, np.float64).ravel()
    n = hc.size
    a, b = hc - sc.mean(), sc - hc.mean()
    bce = lambda y, z: np.mean(_softplus(z) - y * z)
    dbce = lambda y, z: (_sigmoid(z) - y) / n          # d mean-BCE / d z_i

    def grads(y_h, y_s):
        ga, gb = dbce(y_h, a), dbce(y_s, b)            # w.r.t. a_i, b_i
        dhr = 0.5 * (ga - gb.sum() / n)                # a_i = hr_i - mean(sr);  b_j = sr_j - mean(hr)
        dsr = 0.5 * (gb - ga.sum() / n)
        return dhr.astype(np.float32).reshape(-1, 1), dsr.astype(np.float32).reshape(-1, 1)

    g_loss = 0.5 * (bce(0.0, a) + bce(1.0, b))          # ra_adversarial_loss.py:59-69
    d_loss = 0.5 * (bce(hr_label, a) + bce(sr_label, b))  # ra_discriminator_loss.py:55-65
    g_dhr, g_dsr = grads(0.0, 1.0)
    d_dhr, d_dsr = grads(hr_label, sr_label)
    return dict(g_loss=np.float32(g_loss), d_loss=np.float32(d_loss), g_dsr=g_dsr, g_dhr=g_dhr, d_dsr=d_dsr, d_dhr=d_dhr)


def gan_losses(hr_logit, sr_logit, hr_label=1.0, sr_label=0.0):
    """The standard (non-relativistic) critic: Dense(1, sigmoid) (model_builder.py:194-196), AdversarialLoss =
    BCE(ones, p_sr) (adversarial_loss.py:58), DiscriminatorLoss = BCE(sr_labels, p_sr) + BCE(hr_labels, p_hr)
    (discriminator_loss.py:56-59).  tf.keras.losses.BinaryCrossentropy() on probabilities: p clipped to [eps, 1-eps],
    -(y log(p+eps) + (1-y) log(1-p+eps)), eps = 1e-7, mean over the batch.  Gradients are w.r.t. the LOGITS.
    Returns dict(g_loss, d_loss, g_dsr, d_dsr, d_dhr)."""
    zh = np.asarray(hr_logit, np.float64).ravel()
    zs = np.asarray(sr_logit, np.float64).ravel()
    n, eps = zh.size, 1e-7

    def bce(y, z):
        p = _sigmoid(z)
        pc = np.clip(p, eps, 1.0 - eps)
        inside = np.where((p > eps) & (p < 1.0 - eps), p * (1.0 - p), 0.0)
        loss = np.mean(-(y * np.log(pc + eps) + (1.0 - y) * np.log(1.0 - pc + eps)))
        dz = (-y / (pc + eps) + (1.0 - y) / (1.0 - pc + eps)) * inside / n
        return loss, dz.astype(np.float32).reshape(-1, 1)

    g_loss, g_dsr = bce(1.0, zs)
    ls, d_dsr = bce(np.broadcast_to(np.asarray(sr_label, np.float64).ravel(), zs.shape), zs)
    lh, d_dhr = bce(np.broadcast_to(np.asarray(hr_label, np.float64).ravel(), zh.shape), zh)
    return dict(g_loss=np.float32(g_loss), d_loss=np.float32(ls + lh), g_dsr=g_dsr, d_dsr=d_dsr, d_dhr=d_dhr)
