"""Host-side mirror of the data preparation the reference runs in tf.data on the CPU
(``simple_sr/utils/image/image_transforms.py``: crop_naive :50-80, rotate90 :157-173, flip_along_x / flip_along_y :320-345,
resize :348-368; ``DataPipeline._prepare_img_pairs`` data_pipeline.py:318-330), evaluated on the B200 so that the training
GPUs are not fed by one host pipeline.  numpy NHWC (or HWC) in and out; the flips, rotations and crops are exact copies.
"""
import numpy as np

from . import _lib as L


def _batch(img):
    x = np.ascontiguousarray(img, np.float32)
    if x.ndim not in (3, 4):
        raise ValueError("expected a rank-3 or rank-4 image tensor")
    return (x[None], True) if x.ndim == 3 else (x, False)


def _gather(x, out_h, out_w, mode, src=None, oy=None, ox=None, n_out=None):
    n, h, w, c = x.shape
    n_out = n if n_out is None else n_out
    d_x, d_y = L.DeviceBuffer.from_numpy(x), L.DeviceBuffer(n_out * out_h * out_w * c * 4)
    idx = [L.DeviceBuffer.from_numpy(np.ascontiguousarray(a, np.int32)) if a is not None else None for a in (src, oy, ox)]
    L.check(L.load().ssr_augment(d_x.ptr, d_y.ptr, n_out, h, w, c, out_h, out_w, mode, *[L._ptr(i) for i in idx], None))
    out = d_y.download((n_out, out_h, out_w, c), np.float32)
    for b in [d_x, d_y] + [i for i in idx if i is not None]:
        b.free()
    return out


def flip_along_x(img_tensor):
    """:320-331 - tf.image.flip_up_down."""
    x, single = _batch(img_tensor)
    y = _gather(x, x.shape[1], x.shape[2], 2)
    return y[0] if single else y


def flip_along_y(img_tensor):
    """:334-345 - tf.image.flip_left_right."""
    x, single = _batch(img_tensor)
    y = _gather(x, x.shape[1], x.shape[2], 1)
    return y[0] if single else y


def rotate90(img_tensor, rotations=None, rng=None):
    """:157-173 - tf.image.rot90 by ``rotations`` quarter turns (None: drawn from [1, 3) like the reference)."""
    x, single = _batch(img_tensor)
    if rotations is None:
        rotations = int((rng or np.random.default_rng()).integers(1, 3))
    k = int(rotations) % 4
    oh, ow = (x.shape[2], x.shape[1]) if k & 1 else (x.shape[1], x.shape[2])
    y = _gather(x, oh, ow, k << 2)
    return y[0] if single else y


def crop_naive(img_tensor, num_crops, patch_dims, random_seed=None):
    """:50-80 - ``num_crops`` random windows of ``patch_dims`` = (h, w, c) from one image (tf.image.random_crop draws the
    offsets uniformly; numpy's generator stands in for TensorFlow's stream).  Returns [num_crops, h, w, c]."""
    x, _ = _batch(img_tensor)
    if x.shape[0] != 1:
        raise ValueError("crop_naive crops one image at a time")
    ph, pw = int(patch_dims[0]), int(patch_dims[1])
    if ph > x.shape[1] or pw > x.shape[2]:
        raise ValueError("crop dimensions are larger than the image")
    rng = np.random.default_rng(random_seed)
    oy = rng.integers(0, x.shape[1] - ph + 1, size=num_crops)
    ox = rng.integers(0, x.shape[2] - pw + 1, size=num_crops)
    return _gather(x, ph, pw, 0, src=np.zeros(num_crops, np.int32), oy=oy, ox=ox, n_out=num_crops)


def resize_bicubic(img_tensor, scale, antialias=True):
    """tf.image.resize(img, (h / scale, w / scale), method="bicubic", antialias=antialias) for an integer down-scaling
    factor (:348-368 with the bicubic filter the SR pipelines use)."""
    x, single = _batch(img_tensor)
    n, h, w, c = x.shape
    if h % scale or w % scale:
        raise ValueError("image sides must be multiples of the scale")
    d_x, d_y = L.DeviceBuffer.from_numpy(x), L.DeviceBuffer(n * (h // scale) * (w // scale) * c * 4)
    ws = L.DeviceBuffer(L.load().ssr_resize_workspace_bytes(n, h, w, c, scale))
    L.check(L.load().ssr_resize_bicubic(d_x.ptr, d_y.ptr, n, h, w, c, scale, int(bool(antialias)), ws.ptr, None))
    y = d_y.download((n, h // scale, w // scale, c), np.float32)
    for b in (d_x, d_y, ws):
        b.free()
    return y[0] if single else y


def prepare_img_pairs(hr_img, scale, antialias=True):
    """DataPipeline._prepare_img_pairs (data_pipeline.py:318-330; bicubic, no JPEG noise): HR pixels in [0, 255] ->
    (LR in [0, 1] down-scaled by ``scale``, HR in [-1, 1])."""
    hr = np.ascontiguousarray(hr_img, np.float32)
    return resize_bicubic(hr / np.float32(255), scale, antialias), hr / np.float32(127.5) - np.float32(1)
