"""Data-parallel plumbing: one process per GPU.

The reference has no parallelism at all; this is new work required by BASELINE.json.  Two mechanisms:

* :class:`PeerComm` - the peer-memory fabric (``csrc/comm.cu``): every rank owns a heap that all ranks map through CUDA
  IPC over NVLink / NVSwitch; the collectives of a training step (gradient reduce-scatter + Adam + parameter all-gather,
  sync-BatchNorm sums, the RaGAN critic gather, the loss metrics) are kernels over those mappings and live inside the
  captured step graph.  ``torch.distributed`` only carries the 64-byte IPC handles (``PeerComm.connect``).
* :func:`make_grad_allreduce` - the fallback when peer mapping is unavailable: a NCCL all-reduce (through
  ``torch.distributed``) of the flat gradient buffer between backward and Adam, outside the graph; BatchNorm statistics
  and the RaGAN means then stay per rank (a known deviation from the single-device step).
"""
import ctypes as C

import numpy as np

from . import _lib as L


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
    """What the gradient exchange leaves in every rank's buffer (host restatement used by the CPU tests)."""
    return np.mean(np.stack([np.asarray(a, np.float32) for a in arrays]), axis=0).astype(np.float32)


def adam_shard_range(lo, hi, rank, world):
    """Elements of [lo, hi) whose Adam update ``rank`` owns in ``ssr_comm_adam_step`` (units of 4 floats, contiguous,
    the last rank takes the remainder) - host restatement of the kernel's split, pinned by tests/test_host_logic.py."""
    n4 = (hi - lo + 3) // 4
    per = -(-n4 // world)
    a, b = rank * per, min(n4, (rank + 1) * per)
    return lo + 4 * a, lo + 4 * max(a, b)


class _HeapBase:
    def __init__(self, ptr):
        self.ptr = ptr


class HeapView(L.DeviceView):
    """A window of the local heap; ``heap_off`` is its offset, identical on every rank."""

    def __init__(self, base, offset, nbytes):
        super().__init__(base, offset, nbytes)
        self.heap_off = int(offset)


class PeerComm:
    """One rank's end of the peer-memory fabric.  Allocation order (``alloc`` / ``slots``) must be the same on all ranks:
    collectives address peers' buffers by heap offset and meet at slot numbers."""

    def __init__(self, rank, world, device=0, heap_bytes=64 << 20):
        self.lib = L.load()
        self.rank, self.world, self.device = int(rank), int(world), int(device)
        h = C.c_void_p()
        L.check(self.lib.ssr_comm_create(self.device, self.rank, self.world, int(heap_bytes), C.byref(h)))
        self.handle = h.value
        self.heap_bytes = int(heap_bytes)
        self._base = _HeapBase(self.lib.ssr_comm_heap(self.handle))
        self._bump = self.lib.ssr_comm_data_offset()
        self._slot = 0
        self.mode = "single" if world == 1 else None

    # ---- wiring ------------------------------------------------------------------------------------------------
    def ipc_handle(self):
        buf = (C.c_ubyte * 64)()
        L.check(self.lib.ssr_comm_ipc_handle(self.handle, buf))
        return bytes(buf)

    def open_ipc(self, handles):
        blob = b"".join(handles)
        assert len(blob) == 64 * self.world
        L.check(self.lib.ssr_comm_open_ipc(self.handle, blob))
        self.mode = "ipc"

    @classmethod
    def connect(cls, dist, device, heap_bytes):
        """All ranks of an initialised ``torch.distributed`` group: create the heaps and map each other's."""
        rank, world = dist.get_rank(), dist.get_world_size()
        comm = cls(rank, world, device, heap_bytes)
        if world > 1:
            handles = [None] * world
            dist.all_gather_object(handles, comm.ipc_handle())
            comm.open_ipc(handles)
            dist.barrier()
        return comm

    @classmethod
    def local_group(cls, world, device=0, heap_bytes=64 << 20, spin_seconds=None):
        """``world`` (<= 4) ranks inside ONE process on ONE device, for tests on a single GPU.  Kernels that wait for each
        other must never be separate launches on one GPU, so in this mode every collective of the group runs as ONE
        cooperative launch over all ranks (same device code, blockIdx.y = rank): each rank calls it from its own host
        thread, the calls rendezvous on the host (``run_ranks``), and the steps run eagerly (no graph capture)."""
        comms = [cls(r, world, device, heap_bytes) for r in range(world)]
        if world > 1:
            arr = (C.c_void_p * world)(*[c.handle for c in comms])
            L.check(comms[0].lib.ssr_comm_open_local(arr, world))
            for c in comms:
                c.mode = "local"
        if spin_seconds:
            for c in comms:
                L.check(c.lib.ssr_comm_set_spin_limit(c.handle, float(spin_seconds)))
        return comms

    # ---- heap / slot allocation (deterministic: same calls in the same order on every rank) -------------------------
    def alloc(self, nbytes, align=256):
        off = -(-self._bump // align) * align
        if off + nbytes > self.heap_bytes:
            raise MemoryError(f"peer heap exhausted: {off + nbytes} > {self.heap_bytes} bytes (pass a larger heap_bytes)")
        self._bump = off + int(nbytes)
        return HeapView(self._base, off, nbytes)

    def reset(self):
        """Forget every allocation (heap and slots): the next user starts from an empty fabric.  Call it on ALL ranks at
        the same point, after the previous user's last step has completed everywhere (the slot epochs stay in step)."""
        self._bump = self.lib.ssr_comm_data_offset()
        self._slot = 0

    def slots(self, n):
        s = self._slot
        if s + n > self.lib.ssr_comm_max_slots():
            raise MemoryError("peer fabric: out of barrier slots")
        self._slot += int(n)
        return s

    def bn_site(self, c):
        """(handle, slot0, heap offset) for one sync-BatchNorm call site over ``c`` channels."""
        return (self.handle, self.slots((c + 31) // 32), self.alloc(32 * c).heap_off)

    def ragan_site(self, n_local):
        return (self.handle, self.slots(1), self.alloc(8 * n_local * 4).heap_off)

    def adam_site(self):
        return self.slots(self.lib.ssr_comm_adam_slots())

    def allreduce_site(self, count):
        return (self.slots(1), self.alloc(2 * count * 4).heap_off)

    # ---- collectives ------------------------------------------------------------------------------------------------
    def allreduce_f32(self, site, src, dst, count, scale, stream=None):
        L.check(self.lib.ssr_comm_allreduce_f32(self.handle, site[0], site[1], L._ptr(src), L._ptr(dst), count, scale,
                                                stream))

    def adam_step(self, slot0, grad, param, m, v, lo, hi, opt_state, b1, b2, eps, stream=None):
        L.check(self.lib.ssr_comm_adam_step(self.handle, slot0, grad.heap_off, param.heap_off, L._ptr(m), L._ptr(v), lo, hi,
                                            L._ptr(opt_state), b1, b2, eps, stream))

    def barrier(self, slot, stream=None):
        L.check(self.lib.ssr_comm_barrier(self.handle, slot, stream))

    def timeouts(self):
        """Barrier waits that gave up (synchronous).  Non-zero means a rank went missing: results are invalid."""
        v = C.c_ulonglong()
        L.check(self.lib.ssr_comm_status(self.handle, C.byref(v)))
        return int(v.value)

    def check(self):
        t = self.timeouts()
        if t:
            raise L.SsrError(f"peer fabric: {t} barrier wait(s) timed out on rank {self.rank} - a rank is missing or slow")

    def destroy(self):
        if self.handle:
            self.lib.ssr_comm_destroy(self.handle)
            self.handle = None


def run_ranks(fns):
    """Run one callable per emulated rank, each in its own host thread (the collectives of a ``local_group`` rendezvous
    on the host: every rank must be inside the same collective at the same time).  Returns the results in rank order;
    re-raises the first exception."""
    import threading
    results, errors = [None] * len(fns), [None] * len(fns)

    def work(i):
        try:
            results[i] = fns[i]()
        except BaseException as e:   # noqa: BLE001 - reported to the caller below
            errors[i] = e

    threads = [threading.Thread(target=work, args=(i,)) for i in range(len(fns))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for e in errors:
        if e is not None:
            raise e
    return results


def make_grad_allreduce(dist, torch):
    """Fallback exchange: returns ``allreduce(grad_buffer, count, stream_ptr)`` leaving the MEAN over the ranks of the flat
    fp32 gradient in place (NCCL through torch.distributed, ordered on ``stream_ptr``, outside the step graph)."""
    world = dist.get_world_size()

    def allreduce(buf, count, stream_ptr):
        t = torch.as_tensor(_CudaArray(buf.ptr, count), device=torch.device("cuda", torch.cuda.current_device()))
        ext = torch.cuda.ExternalStream(stream_ptr)
        with torch.cuda.stream(ext):
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            t.mul_(1.0 / world)

    return allreduce
