"""Host-side mirror of the tiling half of ``simple_sr/utils/image/image_utils.py`` (lines 40-184).

Same function names, argument meaning and errors as the reference; the copies themselves run on the B200 through
``ssr_segment_tiles`` / ``ssr_stitch_tiles`` (bit-exact data movement, include/ssr_b200.h).  numpy in, numpy out;
the device-resident path used by the tiled upscaler lives in :mod:`simplesr_b200.evaluation`.

Rectangular patches (``patch_width != patch_height``) are supported without overlap, which is what the reference pins
(tests/utils/image/test_image_utils.py:10,44-67: (3,1), (1,3), (2,3), (3,2) through ``_segment`` / ``_reconstruct``).
With ``pixel_overlap > 0`` the reference's loops step rows by ``patch_width`` and columns by ``patch_height``
(image_utils.py:139-140), which for non-square patches yields ragged tiles that ``tf.convert_to_tensor`` rejects, so
that combination raises ValueError here as well.
"""
import numpy as np

from . import _lib as L


def _as_hwc(tensor):
    t = np.asarray(tensor)
    if t.ndim != 3 and not (t.ndim == 4 and t.shape[0] == 1):
        raise ValueError("Tensor must be of rank 3")                       # image_utils.py:108-109
    return t[0] if t.ndim == 4 else t


def segment_into_patches(tensor, patch_width=32, patch_height=32, pixel_overlap=0, stream=None):
    """image_utils.segment_into_patches (:85-121).  Returns (patches [T,ph+2ov,pw+2ov,C], padding [[t,b],[l,r]])."""
    t = _as_hwc(tensor)
    if t.shape[0] < patch_height or t.shape[1] < patch_width:
        raise ValueError("Patch dimensions are larger than image size")    # :115-116
    pw, ph, ov = int(patch_width), int(patch_height), int(pixel_overlap)
    if pw != ph and ov != 0:
        raise ValueError("overlapping patches must be square (the reference's own loops produce ragged tiles otherwise, "
                         "image_utils.py:139-147)")
    h, w, c = t.shape
    rows, cols = -(-h // ph), -(-w // pw)
    padding = [[ov, ov + (ph - h) % ph], [ov, ov + (pw - w) % pw]]          # :126-133 / :152-157
    src = np.ascontiguousarray(t, dtype=np.float32)
    if not np.array_equal(src.astype(t.dtype), t):
        raise ValueError("image values are not exactly representable in float32")
    d_img = L.DeviceBuffer.from_numpy(src, stream)
    tsy, tsx = ph + 2 * ov, pw + 2 * ov
    d_tiles = L.DeviceBuffer(rows * cols * tsy * tsx * c * 4)
    L.segment_tiles_ex(d_img, h, w, c, ph, pw, ov, 0, rows * cols, 0, h, d_tiles, stream)
    out = d_tiles.download((rows * cols, tsy, tsx, c), np.float32, stream)
    d_img.free()
    d_tiles.free()
    return out.astype(t.dtype), padding


def _stitch(patches, image_height, image_width, pixel_overlap, padded_height, padded_width, stream=None):
    patches = np.asarray(patches)
    if patches.ndim != 4:
        raise ValueError("Tensor with patches needs to be of rank 4")      # :54-55, :76-77
    ov = int(pixel_overlap)
    ph, pw = patches.shape[1] - 2 * ov, patches.shape[2] - 2 * ov
    if ph <= 0 or pw <= 0:
        raise ValueError("pixel_overlap leaves no patch interior")
    if ph != pw and ov != 0:
        raise ValueError("overlapping patches must be square")
    rows, cols = -(-image_height // ph), -(-image_width // pw)
    if padded_height != rows * ph or padded_width != cols * pw or patches.shape[0] != rows * cols:
        raise ValueError("padding / patch count do not describe a tiling of the image")
    c = patches.shape[3]
    src = np.ascontiguousarray(patches, dtype=np.float32)
    d_tiles = L.DeviceBuffer.from_numpy(src, stream)
    d_out = L.DeviceBuffer(image_height * image_width * c * 4)
    L.stitch_tiles_ex(d_tiles, image_height, image_width, c, ph, pw, ov, 1, 0, rows * cols, 0, image_height, d_out, stream)
    out = d_out.download((image_height, image_width, c), np.float32, stream)
    d_tiles.free()
    d_out.free()
    return out.astype(patches.dtype)


def reconstruct_from_overlapping_patches(patches, image_height, image_width, pixel_overlap, horizontal_padding,
                                         vertical_padding, stream=None):
    """image_utils.reconstruct_from_overlapping_patches (:40-61)."""
    return _stitch(patches, image_height, image_width, pixel_overlap, image_height + horizontal_padding,
                   image_width + vertical_padding, stream)


def reconstruct_from_patches(patches, original_height, original_width, horizontal_padding=0, vertical_padding=0,
                             stream=None):
    """image_utils.reconstruct_from_patches (:64-82)."""
    if np.asarray(patches).ndim != 4:
        raise ValueError("Tensor with patches needs to be of rank 4")
    if horizontal_padding < 0 or vertical_padding < 0:
        raise ValueError("Padding can't be negative")                      # :79-80
    return _stitch(patches, original_height, original_width, 0, original_height + horizontal_padding,
                   original_width + vertical_padding, stream)
