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


def _load_model(model_path):
    """evaluation._load_model (:320-327): the generator stored at ``model_path`` - a Keras ``.h5`` as the reference's
    training writes it (sr_model.py:244) or the package's ``.npz``; a missing file prints and exits like the reference."""
    import sys
    from . import model_builder
    try:
        return model_builder.build_or_load_generator_model(None, None, None, None, None, None, None, None, (None, None),
                                                           pretrained_model_path=model_path)
    except OSError:
        print(f"Error could not locate model at path: {model_path}, exiting")
        sys.exit(1)


def upscale(model, lr_batch, segmentation_min_width=1000, segmentation_min_height=1000, **tiled):
    """The per-batch body of ``evaluate_on_testdata`` (:253-277): a single image with both sides above the thresholds
    takes the memory-efficient path (128 x 128 patches, 32 pixels of overlap, stitched: ``upscale_tiled``; ``tiled`` may
    carry ``rank`` / ``world_size`` / ``tile_batch`` / ``out``), anything else goes straight through the model.
    Returns ``[N, s*H, s*W, 3]`` float32 (:276-277 puts the batch axis back on the stitched image)."""
    if _eligible_efficient_inference(lr_batch, min_width=segmentation_min_width, min_height=segmentation_min_height):
        return upscale_tiled(model, lr_batch, patch=128, pixel_overlap=32, **tiled)[None]
    return _upscale(model, lr_batch)


def tile_range(num_tiles, rank=0, world_size=1):
    """Contiguous block of row-major tile indices owned by ``rank`` (image_utils.py:139-147 order)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, extra = divmod(num_tiles, world_size)
    begin = rank * base + min(rank, extra)
    return begin, base + (1 if rank < extra else 0)


def tile_band(h, w, patch, pixel_overlap, begin, count):
    """LR rows a rank reads and LR rows it writes for tiles [begin, begin+count): ((src_row0, src_rows), (out_row0,
    out_rows)) - the tile rows it touches, plus the halo on the source side, clipped to the image."""
    cols = -(-w // patch)
    r0, r1 = begin // cols, (begin + count - 1) // cols
    src0, src1 = max(0, r0 * patch - pixel_overlap), min(h, (r1 + 1) * patch + pixel_overlap)
    out0, out1 = r0 * patch, min(h, (r1 + 1) * patch)
    return (src0, src1 - src0), (out0, out1 - out0)


def upscale_tiled(model, lr_image, patch=128, pixel_overlap=32, tile_batch=16, rank=0, world_size=1, out=None):
    """Memory-efficient x``scale`` inference of one large image (evaluation.py:253-277).

    lr_image: [H,W,3] or [1,H,W,3] float32 in [0,1].  Returns the SR image [H*s,W*s,3] float32; with
    ``world_size > 1`` only the pixels of this rank's tiles are written (the rest of ``out`` is left untouched, zeros
    if allocated here), so summing / overlaying the ranks' outputs gives the single-GPU result bit for bit.

    A rank touches only its BAND: it uploads the LR rows its tiles read (tile rows + halo), keeps an SR buffer of its
    tile rows only, and copies each finished batch of tiles back to the host while the next batch computes (second
    stream).  No collective: the halo is re-read from the source image.
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
    lib = model.ctx.lib
    lr = np.ascontiguousarray(lr)
    if out is None:
        out = np.zeros((h * sf, w * sf, c), dtype=np.float32) if world_size > 1 else np.empty((h * sf, w * sf, c), np.float32)
    elif out.shape != (h * sf, w * sf, c) or out.dtype != np.float32 or not out.flags.c_contiguous:
        raise ValueError("out must be a C-contiguous float32 array of shape [H*s, W*s, 3]")
    if count == 0:
        return out
    (src0, src_rows), (out0, out_rows) = tile_band(h, w, patch, pixel_overlap, begin, count)
    # the band buffers live on the model (cudaMalloc / cudaFree of hundreds of MB would cost more than the copies)
    cache = model.__dict__.setdefault("_tiled_buffers", {})
    key = (h, w, patch, pixel_overlap, begin, count)
    if key not in cache:
        for old in cache.values():
            for b in old[:2]:
                b.free()
            old[2].destroy()
            for ev in old[3]:
                ev.destroy()
        cache.clear()
        cache[key] = (L.DeviceBuffer(src_rows * w * c * 4), L.DeviceBuffer(out_rows * sf * w * sf * c * 4), L.Stream(),
                      [L.Event() for _ in range(2)])
    d_img, d_out, copy_stream, events = cache[key]
    L.check(lib.ssr_memcpy_h2d(d_img.ptr, lr[src0:src0 + src_rows].ctypes.data, src_rows * w * c * 4, s))
    row_bytes = w * sf * c * 4
    ps = patch * sf

    def copy_back(t0, t1, ev):
        """Tiles [t0, t1) are stitched (event ev on the compute stream): copy their pixels to the host on the copy
        stream - whole tile rows in one piece, a partial tile row as a rectangle."""
        copy_stream.wait_event(ev)
        t = t0
        while t < t1:
            r, c0 = divmod(t, cols)
            c1 = min(cols, c0 + (t1 - t))
            y0, y1 = r * ps, min(h * sf, (r + 1) * ps)
            if c0 == 0 and c1 == cols:
                # as many complete tile rows as the range holds
                nr = (t1 - t) // cols
                y1 = min(h * sf, (r + nr) * ps)
                L.check(lib.ssr_memcpy_d2h(out.ctypes.data + y0 * row_bytes, d_out.ptr + (y0 - out0 * sf) * row_bytes,
                                           (y1 - y0) * row_bytes, copy_stream.ptr))
                t += nr * cols
            else:
                x0, x1 = c0 * ps, min(w * sf, c1 * ps)
                L.check(lib.ssr_memcpy2d_d2h(out.ctypes.data + y0 * row_bytes + x0 * c * 4, row_bytes,
                                             d_out.ptr + (y0 - out0 * sf) * row_bytes + x0 * c * 4, row_bytes,
                                             (x1 - x0) * c * 4, y1 - y0, copy_stream.ptr))
                t += c1 - c0

    done = 0
    k = 0
    while done < count:
        nb = min(tile_batch, count - done)
        plan = model.plan(nb, ts, ts)
        L.segment_tiles_ex(d_img, h, w, c, patch, patch, pixel_overlap, begin + done, nb, src0, src_rows,
                           plan.buffers["in_f32"], s)
        plan.run(s, model.use_graph)
        L.stitch_tiles_ex(plan.buffers["out_f32"], h, w, c, patch, patch, pixel_overlap, sf, begin + done, nb, out0 * sf,
                          out_rows * sf, d_out, s)
        ev = events[k & 1]
        ev.record(s)
        copy_back(begin + done, begin + done + nb, ev)
        done += nb
        k += 1
    copy_stream.sync()
    model.stream.sync()
    return out
