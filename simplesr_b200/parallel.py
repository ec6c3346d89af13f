"""Data-parallel plumbing: one process per GPU, ``torch.distributed`` (NCCL over NVLink/NVSwitch) for the one exchange
step of the training path - the all-reduce of the flat gradient buffer between backward and Adam.

The reference has no parallelism at all; this is new work required by BASELINE.json.  The product's
device buffers are plain cudaMalloc regions behind the C ABI, so the collective runs on a zero-copy torch view of the
buffer (``__cuda_array_interface__``) on the trainer's own CUDA stream: no staging copy, no extra synchronisation.
"""
import numpy as np


class _CudaArray:
    """Minimal ``__cuda_array_interface__`` carrier for a raw device pointer (fp32, 1-D)."""

    def __init__(self, ptr, count):
        self.__cuda_array_interface__ = {"shape": (int(count),), "typestr": "<f4", "data": (int(ptr), False),
                                         "version": 2, "strides": None}


def shard_batch(global_batch, rank, world_size):
    """Images [begin, begin+count) of the global batch owned by ``rank``; equal shares are required so that the mean of
    the ranks' gradients equals the gradient of the global-batch mean loss (sr_model.py:419-441 on one device)."""
    if global_batch % world_size != 0:
        raise ValueError(f"global batch {global_batch} is not divisible by {world_size} ranks")
    per = global_batch // world_size
    return rank * per, per


def mean_over_ranks_numpy(arrays):
    """What the all-reduce leaves in every rank's buffer (host restatement used by the CPU tests)."""
    return np.mean(np.stack([np.asarray(a, np.float32) for a in arrays]), axis=0).astype(np.float32)


def make_grad_allreduce(dist, torch):
    """Returns ``allreduce(grad_buffer, count, stream_ptr)`` for :class:`simplesr_b200.training.SRResNetTrainer`:
    in-place MEAN over the ranks of the flat fp32 gradient, ordered on ``stream_ptr``."""
    world = dist.get_world_size()

    def allreduce(buf, count, stream_ptr):
        t = torch.as_tensor(_CudaArray(buf.ptr, count), device=torch.device("cuda", torch.cuda.current_device()))
        ext = torch.cuda.ExternalStream(stream_ptr)
        with torch.cuda.stream(ext):
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            t.mul_(1.0 / world)

    return allreduce
