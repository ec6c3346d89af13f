"""Host-side mirror of ``simple_sr/utils/image/metrics.py`` (psnr :4-15, psnr_on_y :18-44, ssim :47-59): same names and
arguments, evaluated on the B200 (``ssr_pixel_loss`` / ``ssr_psnr_y`` / ``ssr_ssim``).  numpy NHWC in, per-image values
out (a rank-3 image gives a scalar, as ``tf.image.psnr`` / ``tf.image.ssim`` do)."""
import numpy as np

from . import _lib as L


def _pair(tensor1, tensor2):
    a = np.ascontiguousarray(tensor1, np.float32)
    b = np.ascontiguousarray(tensor2, np.float32)
    if a.shape != b.shape:
        raise ValueError("tensors need to have the same shape")                            # metrics.py:29-30
    if a.ndim > 4 or a.ndim < 3:
        raise ValueError("tensors need to be either of rank 4 or rank 3")                  # metrics.py:31-32
    single = a.ndim == 3
    if single:
        a, b = a[None], b[None]
    return a, b, single


def _run(kind, tensor1, tensor2, max_val):
    a, b, single = _pair(tensor1, tensor2)
    n, h, w, c = a.shape
    d_a, d_b = L.DeviceBuffer.from_numpy(a), L.DeviceBuffer.from_numpy(b)
    ws = L.DeviceBuffer(max(L.load().ssr_metric_workspace_bytes(n), L.load().ssr_pixel_loss_workspace_bytes(n)))
    out = L.DeviceBuffer((2 + n) * 4)
    lib = L.load()
    if kind == "psnr":
        L.pixel_loss(d_a, d_b, n, h * w * c, 0.0, 0.0, float(max_val), None, ws, out)
        vals = out.download((2 + n,), np.float32)[2:]
    elif kind == "psnr_y":
        if c != 3:
            raise ValueError("psnr_on_y needs RGB images")
        L.check(lib.ssr_psnr_y(d_a.ptr, d_b.ptr, n, h, w, float(max_val), ws.ptr, out.ptr, None))
        vals = out.download((n,), np.float32)
    else:
        L.check(lib.ssr_ssim(d_a.ptr, d_b.ptr, n, h, w, c, float(max_val), ws.ptr, out.ptr, None))
        vals = out.download((n,), np.float32)
    for buf in (d_a, d_b, ws, out):
        buf.free()
    return vals[0] if single else vals.copy()


def psnr(tensor1, tensor2, max_val=2.0):
    """metrics.psnr (:4-15) = tf.image.psnr."""
    return _run("psnr", tensor1, tensor2, max_val)


def psnr_on_y(tensor1, tensor2, max_val=2.0):
    """metrics.psnr_on_y (:18-44): PSNR between the luma channels (tf.image.rgb_to_yuv)."""
    return _run("psnr_y", tensor1, tensor2, max_val)


def ssim(tensor1, tensor2, max_val=2.0):
    """metrics.ssim (:47-59) = tf.image.ssim."""
    return _run("ssim", tensor1, tensor2, max_val)
