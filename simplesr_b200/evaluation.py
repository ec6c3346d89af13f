"""Host-side mirror of the inference half of ``simple_sr/operations/evaluation.py``.

``_eligible_efficient_inference`` (:340-348) and ``_upscale`` (:351-359) keep the reference's names and meaning.
``upscale_tiled`` is the tiled branch of ``evaluate_on_testdata`` (:253-277) fused into one device-resident pass:
segment (zero-padded halo tiles) -> generator on batches of tiles -> stitch, with the tiles optionally sharded across
ranks (one process per GPU, no collective: every rank stitches its own contiguous range of tiles).
"""
import numpy as np

from . import _lib as L


def _get_tensor_height_width(tensor):
    shape = np.shape(tensor)
    if len(shape) == 4:
        return shape[1], shape[2]
    if len(shape) == 3:
        return shape[0], shape[1]
    raise ValueError(f"Received tensor with unexpected rank: {len(shape)}")   # evaluation.py:336-337


def _eligible_efficient_inference(tensor, min_width=1000, min_height=1000):
    """evaluation._eligible_efficient_inference (:340-348)."""
    shape = np.shape(tensor)
    if len(shape) != 3 and len(shape) != 4:
        return False
    if len(shape) == 4 and shape[0] != 1:
        return False
    batch_width, batch_height = _get_tensor_height_width(tensor)
    return bool(batch_width > min_width and batch_height > min_height)


def _upscale(model, lr_batch, tile_batch=16):
    """evaluation._upscale (:351-359): every element of ``lr_batch`` goes through the model on its own (batch 1 in the
    reference); here ``tile_batch`` elements share a launch, which gives the same values (images of a batch are
    independent in every kernel, pinned by tests/test_gpu_rrdb.py::test_rrdb_batch_independence)."""
    lr = np.asarray(lr_batch, dtype=np.float32)
    if lr.ndim == 3:
        lr = lr[None]
    outs = [model(lr[i:i + tile_batch], training=False) for i in range(0, lr.shape[0], tile_batch)]
    return np.concatenate(outs, axis=0)


def tile_range(num_tiles, rank=0, world_size=1):
    """Contiguous block of row-major tile indices owned by ``rank`` (image_utils.py:139-147 order)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, extra = divmod(num_tiles, world_size)
    begin = rank * base + min(rank, extra)
    return begin, base + (1 if rank < extra else 0)


def upscale_tiled(model, lr_image, patch=128, pixel_overlap=32, tile_batch=16, rank=0, world_size=1, out=None):
    """Memory-efficient x``scale`` inference of one large image (evaluation.py:253-277).

    lr_image: [H,W,3] or [1,H,W,3] float32 in [0,1].  Returns the SR image [H*s,W*s,3] float32; with
    ``world_size > 1`` only the pixels of this rank's tiles are written (the rest of ``out`` is left untouched, zeros
    if allocated here), so summing / overlaying the ranks' outputs gives the single-GPU result bit for bit.
    """
    lr = np.asarray(lr_image, dtype=np.float32)
    if lr.ndim == 4:
        if lr.shape[0] != 1:
            raise ValueError("Tensor must be of rank 3")
        lr = lr[0]
    if lr.ndim != 3 or lr.shape[2] != 3:
        raise ValueError("expected an [H,W,3] image")
    h, w, c = lr.shape
    if h < patch or w < patch:
        raise ValueError("Patch dimensions are larger than image size")
    sf = model.upsample_factor
    rows, cols = -(-h // patch), -(-w // patch)
    begin, count = tile_range(rows * cols, rank, world_size)
    ts = patch + 2 * pixel_overlap
    s = model.stream.ptr
    lr = np.ascontiguousarray(lr)
    # the LR image and the stitched SR image live in device buffers cached on the model (cudaMalloc / cudaFree of the
    # 0.8 GB output of a 2048x2048 input would cost more than the D2H copy)
    cache = model.__dict__.setdefault("_tiled_buffers", {})
    if (h, w) not in cache:
        for old in cache.values():
            for b in old:
                b.free()
        cache.clear()
        cache[(h, w)] = (L.DeviceBuffer(lr.nbytes), L.DeviceBuffer(h * sf * w * sf * c * 4))
    d_img, d_out = cache[(h, w)]
    L.check(model.ctx.lib.ssr_memcpy_h2d(d_img.ptr, lr.ctypes.data, lr.nbytes, s))
    d_out.zero(s)
    done = 0
    while done < count:
        nb = min(tile_batch, count - done)
        plan = model.plan(nb, ts, ts)
        L.segment_tiles(d_img, h, w, c, patch, pixel_overlap, begin + done, nb, plan.buffers["in_f32"], s)
        plan.run(s, model.use_graph)
        L.stitch_tiles(plan.buffers["out_f32"], h, w, c, patch, pixel_overlap, sf, begin + done, nb, d_out, s)
        done += nb
    if out is None:
        out = np.empty((h * sf, w * sf, c), dtype=np.float32)
    elif out.shape != (h * sf, w * sf, c) or out.dtype != np.float32 or not out.flags.c_contiguous:
        raise ValueError("out must be a C-contiguous float32 array of shape [H*s, W*s, 3]")
    L.check(model.ctx.lib.ssr_memcpy_d2h(out.ctypes.data, d_out.ptr, out.nbytes, s))
    model.stream.sync()
    return out
