#!/usr/bin/env python
"""bench.py — benchmark of the SimpleSR hot path on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload ...]

ONE JSON line (rank 0).  The headline is BASELINE.json configs[1]: RRDB x4 generator (23 RRDB blocks, nf=64, gc=32)
inference on a batch of 16 synthetic 128x128 LR images per GPU, bf16 storage / fp32 accumulation; one "step" = one
forward pass of the batch; metric = output megapixels per second of the whole job (N > 1: replicas, no collective).

The same run also measures the rest of BASELINE.json's metric and north_star partitions and puts them under "extra":
  extra.esrgan_train   configs[3], ESRGAN training step (RRDB-23 + MAE + VGG19 + RaGAN + both Adam updates): img/s.
                       N > 1: data-parallel over the peer-memory fabric - "strong" (global batch 16 split over the ranks)
                       and "weak" (16 images per GPU), sync-BN and global RaGAN means included
  extra.srresnet_train configs[2], SRResNet x4 training step (MSE, global batch 16 of 96x96 HR), data-parallel
  extra.tiled_infer    configs[4], RRDB x4 tiled inference of one 2048x2048 LR image, tiles sharded over the ranks
  extra.bandwidth_kernels (N = 1) achieved HBM GB/s of the bandwidth-bound kernels at the C2 / C4 shapes
See DESIGN.md "Measurement" for the definition of every key.
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "RRDB x4 output Mpix/s infer"
UNIT = "Mpix/s"
BATCH, LR, SCALE, NB = 16, 128, 4, 23
OUT_MPIX = BATCH * (LR * SCALE) ** 2 / 1e6

# dram__bytes_read.sum + dram__bytes_write.sum of one step from the ncu --set full capture of the launches of a dense
# block as the plan issues them (ncu flushes the caches before every kernel) x 69 blocks + the edge launches (the two
# up-sampling convs and the HR convs write the 0.5 GB HR tensors).  Not measured in the run that prints it:
# "traffic_source" names the committed capture.
DRAM_TRAFFIC = {"bytes_per_step": int((69 * (35.96 + 36.87 + 73.05 + 36.91 + 115.44) + 2433.7) * 1e6),
                "source": "profiles/r02_conv_ncu_full.csv (ncu --set full: the five launches of a dense block x 69 + the "
                          "trunk / up-sampling / output convs; not measured in this run)"}


def rrdb_macs_per_lr_pixel(nb=NB, nf=64, gc=32, scale=SCALE):
    """Algorithmic MACs per LR pixel of build_enhanced_resnet: 17,926,848 for RRDB-23 x4."""
    dense = sum(9 * (nf + k * gc) * gc for k in range(4)) + 9 * (nf + 4 * gc) * nf
    macs = 9 * 3 * nf + nb * 3 * dense + 9 * nf * nf
    res = 1
    for _ in range(int(np.log2(scale))):
        macs += res * 9 * nf * 4 * nf
        res *= 4
    macs += res * 9 * nf * nf + res * 9 * nf * 3
    return macs


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm_gbs=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sustained=p["bf16_tflops_sustained"],
                    source="measured")
    return dict(hbm_gbs=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi sampler for the clocks line of B200_PROFILING.md.

    nvidia-smi needs a few hundred ms to come up, which is longer than a short timed region (20 steps of 12 ms), so the
    sampler is started BEFORE the warm-up and every sample carries nvidia-smi's own timestamp; ``begin()`` / ``end()``
    mark the window of the load and ``stop()`` keeps the samples inside it.  A caller whose timed region is shorter than
    ``MIN_WINDOW_S`` keeps the same load running (untimed) until the window is that long (``extend_until``)."""

    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    MIN_WINDOW_S = 0.6

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.t_begin = self.t_end = None

    def start(self):
        if self.proc is not None:
            return
        self.t_begin = self.t_end = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50", "-i",
                 str(self.device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def begin(self):
        import datetime
        self.t_begin = datetime.datetime.now()
        self.t_end = None

    def end(self):
        import datetime
        self.t_end = datetime.datetime.now()

    def extend_until(self, run_more):
        """Call ``run_more()`` (one more untimed step of the same load, synchronised) until the window since
        ``begin()`` is MIN_WINDOW_S long; returns the number of extra steps."""
        import datetime
        extra = 0
        if self.proc is None or self.t_begin is None:
            return extra
        while (datetime.datetime.now() - self.t_begin).total_seconds() < self.MIN_WINDOW_S and extra < 100000:
            run_more()
            extra += 1
        self.end()
        return extra

    @staticmethod
    def _stamp(text):
        import datetime
        for fmt in ("%Y/%m/%d %H:%M:%S.%f", "%Y/%m/%d %H:%M:%S"):
            try:
                return datetime.datetime.strptime(text, fmt)
            except ValueError:
                pass
        return None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        if self.t_begin is not None and self.t_end is None:
            self.end()
        proc, self.proc = self.proc, None
        proc.terminate()
        try:
            out, _ = proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            proc.kill()
            out, _ = proc.communicate()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = []
        for line in (out or "").strip().splitlines():
            f = [t.strip() for t in line.split(",")]
            if len(f) < 9:
                continue
            try:
                rows.append((self._stamp(f[0]), float(f[1]), float(f[2]), f))
            except ValueError:
                continue
        window = "all samples since the warm-up (no usable timestamps)"
        if self.t_begin is not None and any(r[0] is not None for r in rows):
            inside = [r for r in rows if r[0] is not None and self.t_begin <= r[0] <= self.t_end]
            if inside:
                rows = inside
                window = f"{(self.t_end - self.t_begin).total_seconds() * 1e3:.0f} ms of the measured load"
            else:
                window = "all samples since the warm-up (none inside the marked window)"
        elif self.t_begin is None:
            window = "start() to stop()"
        sm, mx, reasons = [], [], set()
        for _, a, b, f in rows:
            sm.append(a)
            mx.append(b)
            hit = False
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
                    hit = True
            # any other clock-event reason (e.g. sw power scaling / idle ramp) is reported by its raw bit mask
            if not hit and f[4].lower() not in ("0", "0x0", "0x0000000000000000", "[n/a]", "n/a", ""):
                reasons.add(f"mask:{f[4]}")
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": window}


class Dist:
    """torch.distributed (NCCL) for the plumbing: barrier, max over ranks, exchange of the IPC handles."""

    def __init__(self, rank, local_rank, world):
        self.rank, self.local_rank, self.world = rank, local_rank, world
        self.torch = self.dist = None
        if world > 1:
            import torch
            import torch.distributed as dist
            torch.cuda.set_device(local_rank)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            self.torch, self.dist = torch, dist

    def barrier(self, stream=None):
        if stream is not None:
            stream.sync()
        if self.dist is not None:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def reduce_max(self, v):
        if self.dist is None:
            return v
        t = self.torch.tensor([v], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def peer_comm(self, heap_bytes):
        """The peer-memory fabric of this job, or (None, reason) when the heaps cannot be mapped (then the training
        extras fall back to the NCCL all-reduce hook and say so)."""
        if self.world == 1:
            return None, "single GPU"
        from simplesr_b200 import parallel as P
        ok, comm, why = 1, None, ""
        try:
            comm = P.PeerComm.connect(self.dist, self.local_rank, heap_bytes)
        except Exception as e:   # noqa: BLE001 - any mapping failure selects the fallback on ALL ranks
            ok, why = 0, f"{type(e).__name__}: {e}"
        t = self.torch.tensor([ok], dtype=self.torch.int32, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
        if int(t.item()) == 0:
            if comm is not None:
                comm.destroy()
            return None, why or "a peer rank could not map the heaps"
        return comm, "cuda-ipc"

    def close(self):
        if self.dist is not None:
            self.dist.barrier()
            self.dist.destroy_process_group()


def build_model(nb=NB, seed=1, device=0):
    from simplesr_b200 import model_builder as MB
    return MB.build_enhanced_resnet(upsample_factor=SCALE, num_rrdb_blocks=nb, seed=seed, device=device)


def time_layer_shapes(model, peaks, reps=20):
    """CUDA-event timing of the five launches of one dense block, each alone in a graph of `reps` launches
    (roofline.per_layer).  The plan pairs the growth convs: launch k computes conv k and, over the same input, the
    partial sums of conv k + 1 (N = 64), the tail adds the newest 32 input channels (model_builder._fused_growth_ops)."""
    from simplesr_b200 import _lib as L
    s = model.stream.ptr
    plan = model.plan(BATCH, LR, LR)
    px = BATCH * LR * LR
    if getattr(model, "fuse_growth", False):
        shapes = [("growth pair 0: conv 64->32 + 64->32 partial", 64 * 64), ("tail 1: conv 32->32 + carry", 32 * 32),
                  ("growth pair 2: conv 128->32 + 128->32 partial", 128 * 64), ("tail 3: conv 32->32 + carry", 32 * 32),
                  ("conv 192->64 + residual (CTA pair)", 192 * 64)]
    else:
        shapes = [("conv 64->32", 64 * 32), ("conv 96->32", 96 * 32), ("conv 128->32", 128 * 32),
                  ("conv 160->32", 160 * 32), ("conv 192->64 + residual (CTA pair)", 192 * 64)]
    out = []
    for (name, kn), op in zip(shapes, plan.ops[2:7]):
        g = L.Graph(s, lambda: [op(s) for _ in range(reps)])
        g.launch(s)
        e0, e1 = L.Event(), L.Event()
        e0.record(s)
        for _ in range(3):
            g.launch(s)
        e1.record(s)
        e1.sync()
        g.destroy()
        ms = e0.elapsed_ms(e1) / (3 * reps)
        flops = 2.0 * 9 * kn * px
        out.append({"layer": name, "ms": round(ms, 4), "tflops": round(flops / ms / 1e9, 1),
                    "frac": round(flops / ms / 1e9 / peaks["tf_burst"], 3)})
    return out


# ---------------------------------------------------------------------------------------------------------------------
# CPU arm: the reference's CPU path restated on torch-CPU / oneDNN (TensorFlow cannot be installed here)
# ---------------------------------------------------------------------------------------------------------------------
def cpu_rrdb_images_per_second(images, repeats=1, threads=None):
    """Times oracle.torch_cpu.rrdb_forward (RRDB-23 x4, fp32, 128x128 LR) on `images` images of the batch with all host
    threads.  Returns (seconds per call, threads)."""
    import torch
    from oracle import ssr_oracle as O
    from oracle import torch_cpu as T
    n_thr = T.set_threads(torch, threads)
    params = O.init_rrdb_params(seed=1, upsample_factor=SCALE, num_rrdb_blocks=NB)
    x = np.random.default_rng(0).uniform(0, 1, size=(images, LR, LR, 3)).astype(np.float32)
    t0 = time.perf_counter()
    for _ in range(repeats):
        T.rrdb_forward(params, x, upsample_factor=SCALE, num_rrdb_blocks=NB)
    return (time.perf_counter() - t0) / repeats, n_thr


def cpu_baseline_sample():
    """Bounded sample for the "cpu_baseline" key of our own line: ONE step of the workload - the whole batch of 16 in one
    call, as the GPU arm runs it - after a one-image warm-up (about 10 s on the GPU box's host cores)."""
    cpu_rrdb_images_per_second(1)
    dt, thr = cpu_rrdb_images_per_second(BATCH)
    mpix = BATCH * (LR * SCALE) ** 2 / 1e6
    return {"value": round(mpix / dt, 4), "unit": UNIT, "cores": thr, "kind": "port",
            "sample": f"one step = all {BATCH} images in one call (RRDB-{NB} x4, 128x128 LR, fp32, torch-CPU/oneDNN "
                      f"restatement of the reference graph - not TensorFlow), {dt:.1f} s"}


def run_reference(args, rank):
    """--impl reference: the reference's CPU path on the host cores.  TensorFlow 2.2 cannot be installed here, so this
    times the torch-CPU/oneDNN restatement (oracle/torch_cpu.py, pinned to the numpy oracle by tests) with every host
    thread (set explicitly: torchrun exports OMP_NUM_THREADS=1).  Each step = a bounded sample of the batch of 16, sized
    from one warm-up image so that the whole run ends within a few minutes; steps and warmup are the requested ones."""
    if rank != 0:
        return
    t_img, thr = cpu_rrdb_images_per_second(1)               # also the first warm-up
    budget = 150.0 / max(1, args.steps + args.warmup)        # seconds per step
    images = int(max(1, min(BATCH, budget // max(t_img, 1e-3))))
    for _ in range(max(0, args.warmup - 1)):
        cpu_rrdb_images_per_second(images)
    dt, thr = cpu_rrdb_images_per_second(images, repeats=args.steps)
    val = images * (LR * SCALE) ** 2 / 1e6 / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": round(val, 4), "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt * 1e3, 2), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"RRDB-{NB} x4 inference, batch {BATCH} of {LR}x{LR} LR per GPU (configs[1])",
                   "note": f"reference arm = CPU restatement of the reference graph on torch-CPU/oneDNN, {thr} threads "
                           f"(TensorFlow unavailable offline); each step = {images} of the {BATCH} images"},
        "cpu_baseline": {"value": round(val, 4), "unit": UNIT, "cores": thr, "kind": "port",
                         "sample": f"{images} of {BATCH} images per step, {args.steps} steps"},
        "e2e": {"value": round(val, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------------
# training workloads (configs[2], configs[3])
# ---------------------------------------------------------------------------------------------------------------------
TRAIN_METRICS = {"srresnet_train": "SRResNet x4 train img/s", "rrdb_train": "RRDB x4 generator train img/s",
                 "esrgan_g_train": "ESRGAN generator (MAE+VGG) train img/s", "esrgan_train": "ESRGAN train img/s"}


def train_flops(workload, gb, hrs):
    """fwd + dgrad + wgrad of every conv (the input conv has no dgrad): ~3x the forward MACs."""
    if workload == "srresnet_train":
        return 0.1227e12 * gb / 16
    flops = 3.0 * 2.0 * rrdb_macs_per_lr_pixel() * gb * (hrs // 4) ** 2
    if workload in ("esrgan_g_train", "esrgan_train"):
        flops += 3.0 * 2.0 * 6.370e9 * gb     # VGG19 to block5_conv4: 2 forwards + 1 dgrad
    if workload == "esrgan_train":
        flops += 8.0 * 2.0 * 1.572e9 * gb     # discriminator: 2 fwd + 3 backward passes (upper bound, §8a a13)
    return flops


def bench_train(workload, D, steps, warmup, per_gpu_batch=None, comm=None, comm_why="", sampler=None):
    """One training workload.  global batch 16 split over the ranks ("strong"), or per_gpu_batch images on every rank
    ("weak").  The timed loop IS the public train_step: pinned host batches in, metrics out (one step late)."""
    from simplesr_b200 import _lib as L
    from simplesr_b200 import model_builder as MB
    from simplesr_b200 import parallel as P
    from simplesr_b200.training import RRDBTrainer, SRResNetTrainer
    rank, world, dev = D.rank, D.world, D.local_rank
    srres = workload == "srresnet_train"
    hrs = 96 if srres else 128
    if per_gpu_batch is None:
        gb = 16
        begin, per = P.shard_batch(gb, rank, world)
        scaling = "strong"
    else:
        per, gb, begin = per_gpu_batch, per_gpu_batch * world, rank * per_gpu_batch
        scaling = "weak"
    hook = None
    if comm is not None:
        comm.reset()          # the previous workload's buffers are dead (barrier at its end): reuse the heap
    if world > 1 and comm is None:
        hook = P.make_grad_allreduce(D.dist, D.torch)
    if srres:
        model = MB.build_resnet(upsample_factor=4, num_res_blocks=16, batch_normalization=False, seed=1, device=dev)
        tr = SRResNetTrainer(model, loss=("mse", 1.0), learning_rate=1e-4, allreduce=hook, comm=comm)
        extra = []
    else:
        model = MB.build_enhanced_resnet(upsample_factor=4, num_rrdb_blocks=NB, seed=1, device=dev)
        extra = []
        if workload in ("esrgan_g_train", "esrgan_train"):
            from simplesr_b200 import vgg as V
            extra = [V.VGGLoss(output_layers="block5_conv4", loss_weight=1.0, after_activation=False, seed=2, device=dev)]
        if workload == "esrgan_train":
            from simplesr_b200 import discriminator as DM
            disc = DM.build_discriminator(input_dims=(hrs, hrs), relativistic=True, seed=3, device=dev)
            extra.append(DM.RaGANLoss(disc, loss_weight=5e-3, learning_rate=1e-4, allreduce=hook))
        tr = RRDBTrainer(model, loss=("mae", 1e-2 if extra else 1.0), learning_rate=1e-4, allreduce=hook, comm=comm,
                         extra_losses=extra)
    rng = np.random.default_rng(0)
    lr_all = rng.uniform(0, 1, size=(gb, hrs // 4, hrs // 4, 3)).astype(np.float32)
    hr_all = rng.uniform(-1, 1, size=(gb, hrs, hrs, 3)).astype(np.float32)
    pin_lr, pin_hr = L.PinnedArray((per, hrs // 4, hrs // 4, 3), np.float32), L.PinnedArray((per, hrs, hrs, 3), np.float32)
    pin_lr.array[...] = lr_all[begin:begin + per]
    pin_hr.array[...] = hr_all[begin:begin + per]
    tr.prepare(per, hrs // 4, hrs // 4)
    D.barrier(model.stream)          # all ranks enter the first step together (in-graph barriers have a spin limit)
    for _ in range(warmup):
        tr.train_step(pin_lr.array, pin_hr.array, lag=1)
    tr.flush()
    D.barrier(model.stream)
    if sampler is not None:
        sampler.start()
    e0, e1 = L.Event(), L.Event()
    s = model.stream.ptr
    t0 = time.perf_counter()
    e0.record(s)
    out = None
    for _ in range(steps):
        out = tr.train_step(pin_lr.array, pin_hr.array, lag=1)
    e1.record(s)
    tr.flush()
    wall_ms = (time.perf_counter() - t0) * 1e3
    dev_ms = e0.elapsed_ms(e1)
    D.barrier(model.stream)
    ms = D.reduce_max(wall_ms) / steps
    ms_dev = D.reduce_max(dev_ms) / steps
    clocks = sampler.stop() if sampler is not None else None
    timeouts = comm.timeouts() if comm is not None else 0
    launches = tr.launches_per_step(per, hrs // 4, hrs // 4)
    res = None
    if rank == 0:
        flops = train_flops(workload, gb, hrs)
        peak = load_peaks()["tf_sustained"] * world
        exchange = ("none (single GPU)" if world == 1 else
                    "peer-memory fabric over NVLink (CUDA IPC): reduce-scatter + Adam + all-gather in one kernel per "
                    "bucket inside the step graph; sync-BN and global RaGAN means" if comm is not None else
                    f"FALLBACK: NCCL all-reduce hook outside the graph, per-rank BN statistics ({comm_why})")
        res = {"metric": TRAIN_METRICS[workload], "value": round(gb / (ms_dev * 1e-3), 1), "unit": "img/s",
               "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": round(ms_dev, 4), "scaling": scaling,
               "global_batch": gb, "per_gpu_batch": per, "dtype": "bf16",
               "e2e": {"value": round(gb / (ms * 1e-3), 1), "unit": "img/s", "ms_per_step": round(ms, 4),
                       "h2d_bytes_per_step": int(pin_lr.nbytes + pin_hr.nbytes),
                       "d2h_bytes_per_step": int(tr._plan(per, hrs // 4, hrs // 4)["metrics_count"] * 4),
                       "note": "public train_step: pinned host batches in, metrics out (read one step late)"},
               "gpu_launches_per_step": launches,
               "roofline": {"bound": "tensor", "achieved": round(flops / (ms_dev * 1e-3) / 1e12, 2), "peak": peak,
                            "unit": "TFLOP/s", "frac": round(flops / (ms_dev * 1e-3) / 1e12 / peak, 5),
                            "kernel": f"whole step ({flops / 1e12:.4f} TFLOP algorithmic)"},
               "exchange": exchange, "comm_timeouts": timeouts, "last_metrics": tr.last_metrics()}
        if clocks is not None:
            res["clocks"] = clocks
    tr.release()
    for el in extra:
        if hasattr(el, "release"):
            el.release()
    model.release()
    pin_lr.free()
    pin_hr.free()
    return res


# ---------------------------------------------------------------------------------------------------------------------
# tiled inference (configs[4])
# ---------------------------------------------------------------------------------------------------------------------
def bench_tiled(D, steps, warmup, side, sampler=None, model=None):
    from simplesr_b200 import _lib as L
    from simplesr_b200 import evaluation as EV
    rank, world = D.rank, D.world
    own = model is None
    if own:
        model = build_model(device=D.local_rank)
    rng = np.random.default_rng(0)
    pin_in = L.PinnedArray((side, side, 3), np.float32)
    pin_in.array[...] = rng.uniform(0, 1, size=(side, side, 3)).astype(np.float32)
    pin = L.PinnedArray((side * SCALE, side * SCALE, 3), np.float32)
    tiles = (-(-side // 128)) ** 2
    begin, count = EV.tile_range(tiles, rank, world)
    for _ in range(warmup):
        EV.upscale_tiled(model, pin_in.array, tile_batch=16, rank=rank, world_size=world, out=pin.array)
    D.barrier(model.stream)
    if sampler is not None:
        sampler.start()
    t0 = time.perf_counter()
    for _ in range(steps):
        EV.upscale_tiled(model, pin_in.array, tile_batch=16, rank=rank, world_size=world, out=pin.array)
    ms = D.reduce_max((time.perf_counter() - t0) * 1e3) / steps
    D.barrier(model.stream)
    clocks = sampler.stop() if sampler is not None else None
    (src0, src_rows), (out0, out_rows) = EV.tile_band(side, side, 128, 32, begin, count)
    res = None
    if rank == 0:
        mpix = (side * SCALE) ** 2 / 1e6
        flops = 2.0 * rrdb_macs_per_lr_pixel() * tiles * 192 * 192
        peak = load_peaks()["tf_sustained"] * world
        res = {"metric": "RRDB x4 tiled output Mpix/s infer", "value": round(mpix / (ms * 1e-3), 2), "unit": "Mpix/s",
               "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": round(ms, 3), "scaling": "strong",
               "dtype": "bf16",
               "config": {"workload": f"RRDB-23 x4 memory-efficient inference of one {side}x{side} LR image: {tiles} "
                                      "tiles of 192x192 (128 + 2x32 halo) (configs[4])",
                          "parallelism": f"tiles sharded over {world} ranks, no collective; each rank uploads, "
                                         "stitches and copies back its own band"},
               "e2e": {"value": round(mpix / (ms * 1e-3), 2), "unit": "Mpix/s",
                       "h2d_bytes_per_step": int(src_rows * side * 12), "d2h_bytes_per_step": int(out_rows * SCALE * side * SCALE * 12),
                       "note": "the timed loop IS the public upscale_tiled call (pinned host image in, host image out); "
                               "bytes are rank 0's band"},
               "roofline": {"bound": "tensor", "achieved": round(flops / (ms * 1e-3) / 1e12, 1), "peak": peak,
                            "unit": "TFLOP/s", "frac": round(flops / (ms * 1e-3) / 1e12 / peak, 4),
                            "kernel": "conv_tc_kernel over all tiles (halo recompute counted as the reference does: "
                                      "2.25x an untiled pass)"}}
        if clocks is not None:
            res["clocks"] = clocks
    if own:
        model.release()
    pin.free()
    pin_in.free()
    return res


# ---------------------------------------------------------------------------------------------------------------------
# bandwidth-bound kernels: achieved HBM GB/s at the C2 / C4 shapes (north_star: "achieved HBM GB/s ... for pixel-shuffle
# and elementwise work")
# ---------------------------------------------------------------------------------------------------------------------
def bench_bandwidth_kernels(reps=10):
    from simplesr_b200 import _lib as L
    peaks = load_peaks()
    s = L.Stream()
    flush = L.DeviceBuffer(256 << 20)        # > 126 MB L2: written between timed launches

    def timed(fn):
        for _ in range(2):
            fn(s.ptr)
        tot = 0.0
        for _ in range(reps):
            flush.zero(s.ptr)
            e0, e1 = L.Event(), L.Event()
            e0.record(s.ptr)
            fn(s.ptr)
            e1.record(s.ptr)
            e1.sync()
            tot += e0.elapsed_ms(e1)
            e0.destroy()
            e1.destroy()
        return tot / reps

    out = []

    def add(name, nbytes, fn):
        ms = timed(fn)
        gbs = nbytes / ms / 1e6
        out.append({"kernel": name, "bytes": int(nbytes), "ms": round(ms, 4), "gbs": round(gbs, 1),
                    "frac": round(gbs / peaks["hbm_gbs"], 3)})

    # depth_to_space(2), C2 second up-conv: [16,256,256,256] bf16 -> [16,512,512,64]
    n, h, w, c = 16, 256, 256, 256
    a, b = L.DeviceBuffer(n * h * w * c * 2), L.DeviceBuffer(n * h * w * c * 2)
    add("ssr_depth_to_space2 [16,256,256,256] bf16", 2 * a.nbytes, lambda st: L.depth_to_space2(a, b, n, h, w, c // 4, 2, st))
    add("ssr_space_to_depth2 [16,512,512,64] bf16", 2 * a.nbytes, lambda st: L.space_to_depth2(b, a, n, h, w, c // 4, 2, st))
    a.free()
    b.free()
    # pixel loss + gradient, C2-sized SR batch [16,512,512,3] fp32: hr + sr read, gradient written
    cnt = 16 * 512 * 512 * 3
    hr, sr, g = (L.DeviceBuffer(cnt * 4) for _ in range(3))
    ws, o = L.DeviceBuffer(L.load().ssr_pixel_loss_workspace_bytes(16)), L.DeviceBuffer(18 * 4)
    add("ssr_pixel_loss [16,512,512,3] fp32 (+grad)", 3 * cnt * 4,
        lambda st: L.pixel_loss(hr, sr, 16, cnt // 16, 1.0, 0.0, 2.0, g, ws, o, st))
    tvw = L.DeviceBuffer(L.load().ssr_total_variation_workspace_bytes())
    add("ssr_total_variation [16,512,512,3] fp32 (+grad)", 3 * cnt * 4,
        lambda st: L.total_variation(sr, 16, 512, 512, 3, 127.5, 2e-7, g, tvw, o, st))
    for x in (hr, sr, g):
        x.free()
    # Adam over the RRDB-23 parameters (16.9 M fp32): p, g, m, v read; p, m, v written
    cnt = 16919555
    p, gg, m, v = (L.DeviceBuffer(cnt * 4) for _ in range(4))
    for x in (p, gg, m, v):
        x.zero(s.ptr)
    add("ssr_adam_step 16.9 M fp32", 7 * cnt * 4, lambda st: L.adam_step(p, gg, m, v, cnt, 1e-4, 0.9, 0.999, 1e-7, 1.0, st))
    for x in (p, gg, m, v):
        x.free()
    # elementwise residual merge / activation backward / max-pool at C4 HR-feature shapes [16,128,128,64] bf16
    px, ch = 16 * 128 * 128, 64
    x1, x2, x3 = (L.DeviceBuffer(px * ch * 2) for _ in range(3))
    add("ssr_axpby_bf16 [16,128,128,64]", 3 * px * ch * 2, lambda st: L.axpby_bf16(x1, ch, 0, x2, ch, 0, 0.2, x3, ch, 0, px, ch, st))
    add("ssr_act_bwd_bf16 [16,128,128,64]", 3 * px * ch * 2,
        lambda st: L.act_bwd_bf16(x1, ch, 0, x2, ch, 0, None, 0.2, x3, ch, 0, px, ch, st))
    y = L.DeviceBuffer(px * ch * 2 // 4)
    add("ssr_maxpool2_bf16 [16,128,128,64]", px * ch * 2 * 5 // 4, lambda st: L.maxpool2_bf16(x1, y, 16, 128, 128, ch, st))
    # BatchNorm statistics + affine/LeakyReLU, discriminator layer 1 output [16,64,64,64] bf16
    pxb = 16 * 64 * 64
    bws = L.DeviceBuffer(L.load().ssr_bn_workspace_bytes(ch))
    mean, istd, gam, bet = (L.DeviceBuffer(ch * 4) for _ in range(4))
    add("ssr_bn_stats_bf16 [16,64,64,64]", pxb * ch * 2, lambda st: L.bn_stats_bf16(x1, pxb, ch, 1e-3, 0.8, bws, mean, istd, None, None, st))
    add("ssr_bn_lrelu_fwd_bf16 [16,64,64,64]", 2 * pxb * ch * 2,
        lambda st: L.bn_lrelu_fwd_bf16(x1, mean, istd, gam, bet, 0.2, x2, pxb, ch, st))
    for x in (x1, x2, x3, y):
        x.free()
    # tile gather / scatter of the tiled path: 16 tiles of a 2048^2 image
    side = 2048
    img, tiles = L.DeviceBuffer(side * side * 12), L.DeviceBuffer(16 * 192 * 192 * 12)
    add("ssr_segment_tiles 16 x 192^2 x 3 fp32", 2 * tiles.nbytes, lambda st: L.segment_tiles(img, side, side, 3, 128, 32, 0, 16, tiles, st))
    sro, big = L.DeviceBuffer(16 * 768 * 768 * 12), L.DeviceBuffer(512 * 8192 * 12)
    add("ssr_stitch_tiles 16 x 512^2 x 3 fp32", 2 * 16 * 512 * 512 * 12,
        lambda st: L.stitch_tiles_ex(sro, side, side, 3, 128, 128, 32, 4, 0, 16, 0, 512, big, st))
    # discriminator Dense(32768 -> 1024), batch 16: the 134 MB fp32 weight matrix is the traffic
    K, O_ = 32768, 1024
    xw, wd, bd = L.DeviceBuffer(16 * K * 4), L.DeviceBuffer(K * O_ * 4), L.DeviceBuffer(O_ * 4)
    dws = L.DeviceBuffer(L.load().ssr_dense_workspace_bytes(16, O_))
    hh, yy = L.DeviceBuffer(16 * O_ * 4), L.DeviceBuffer(16 * O_ * 4)
    add("ssr_dense_fwd_f32 16 x 32768 -> 1024", K * O_ * 4, lambda st: L.dense_fwd_f32(xw, wd, bd, 16, K, O_, True, 0.2, dws, hh, yy, st))
    gx, gw, gb_ = L.DeviceBuffer(16 * K * 4), L.DeviceBuffer(K * O_ * 4), L.DeviceBuffer(O_ * 4)
    add("ssr_dense_bwd_f32 16 x 32768 -> 1024 (dx, dW, db)", 3 * K * O_ * 4,
        lambda st: L.dense_bwd_f32(xw, wd, yy, 16, K, O_, gx, gw, gb_, False, st))
    return {"peak_gbs": peaks["hbm_gbs"], "peak_source": peaks["source"], "l2": "256 MB written between timed launches",
            "kernels": out}


# ---------------------------------------------------------------------------------------------------------------------
def bench_rrdb_infer(args, D, sampler):
    from simplesr_b200 import _lib as L
    rank, world = D.rank, D.world
    if rank == 0:
        sampler.start()          # nvidia-smi comes up (NVML init can take a second) while the model is built
    peaks = load_peaks()
    model = build_model(device=D.local_rank)
    ctx, stream = model.ctx, model.stream
    plan = model.plan(BATCH, LR, LR)
    rng = np.random.default_rng(rank)
    pin_in = L.PinnedArray((BATCH, LR, LR, 3), np.float32)
    pin_out = L.PinnedArray((BATCH, LR * SCALE, LR * SCALE, 3), np.float32)
    pin_in.array[...] = rng.uniform(0, 1, size=pin_in.shape).astype(np.float32)
    L.check(ctx.lib.ssr_memcpy_h2d(plan.buffers["in_f32"].ptr, pin_in.ptr, pin_in.nbytes, stream.ptr))

    # ---- device-resident throughput ("value")
    for _ in range(args.warmup):
        plan.run(stream.ptr)
    D.barrier(stream)
    if rank == 0:
        sampler.begin()
    e0, e1 = L.Event(), L.Event()
    e0.record(stream.ptr)
    for _ in range(args.steps):
        plan.run(stream.ptr)
    e1.record(stream.ptr)
    e1.sync()
    clock_extra = 0
    if rank == 0:
        sampler.end()
        # 20 steps are 0.25 s, less than a handful of nvidia-smi periods: the same step keeps running, untimed, until the
        # clocks have been sampled over MIN_WINDOW_S of this load
        def one_more():
            plan.run(stream.ptr)
            stream.sync()
        clock_extra = sampler.extend_until(one_more)
    D.barrier(stream)
    ms_total = D.reduce_max(e0.elapsed_ms(e1))
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None:
        clocks["untimed_steps_after_the_timed_region"] = clock_extra
    ms_step = ms_total / args.steps

    # ---- end to end through the public API (host buffers, copies inside the timed region)
    for _ in range(2):
        model(pin_in.array, out=pin_out.array)
    D.barrier(stream)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        model(pin_in.array, out=pin_out.array)
    e2e_ms = D.reduce_max((time.perf_counter() - t0) * 1e3) / args.steps
    D.barrier(stream)

    line = None
    if rank == 0:
        flops_step = 2.0 * rrdb_macs_per_lr_pixel() * BATCH * LR * LR
        ach = flops_step / (ms_step * 1e-3) / 1e12
        line = {
            "metric": METRIC, "value": round(world * OUT_MPIX / (ms_step * 1e-3), 2), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_step, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"RRDB-{NB} x4 inference, batch {BATCH} of {LR}x{LR} LR per GPU (configs[1])",
                       "parallelism": f"replicas x{world} (no collective)",
                       "l2": "no flush: per-step working set ~1.4 GB > 126 MB L2",
                       "parity": "conv numerics: oracle restates TF semantics, unpinned (TensorFlow absent); tiling / "
                                 "depth_to_space pinned to the reference's fixtures"},
            "e2e": {"value": round(world * OUT_MPIX / (e2e_ms * 1e-3), 2), "unit": UNIT,
                    "h2d_bytes_per_step": int(pin_in.nbytes), "d2h_bytes_per_step": int(pin_out.nbytes),
                    "ms_per_step": round(e2e_ms, 4)},
            "gpu_launches": int(plan.launches * args.steps),
            "clocks": clocks,
            "roofline": {"bound": "tensor", "achieved": round(ach, 1), "peak": peaks["tf_sustained"],
                         "unit": "TFLOP/s", "frac": round(ach / peaks["tf_sustained"], 4),
                         "traffic": DRAM_TRAFFIC["bytes_per_step"], "traffic_source": DRAM_TRAFFIC["source"],
                         "kernel": "conv_tc_kernel (all 351 conv launches of the step; algorithmic FLOPs / step time)",
                         "peak_source": peaks["source"] + " sustained (kernel timed inside a long step)"},
        }
        if not args.no_per_layer:
            line["roofline"]["per_layer"] = time_layer_shapes(model, peaks)
    pin_in.free()
    pin_out.free()
    return line, model


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="all",
                    choices=["all", "rrdb_infer", "srresnet_train", "rrdb_train", "esrgan_g_train", "esrgan_train",
                             "tiled_infer", "bandwidth"],
                    help="all = the headline (rrdb_infer, BASELINE configs[1]) with the other configs under 'extra' "
                         "(default); a single name runs only that workload and prints its own line")
    ap.add_argument("--tiled-lr", type=int, default=2048, help="LR image side of the tiled workload")
    ap.add_argument("--weak-batch", type=int, default=0, help="training workloads: images per GPU (weak scaling)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-per-layer", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="headline only")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    D = Dist(rank, local_rank, world)
    sampler = ClockSampler(local_rank)
    heap = 736 << 20     # ESRGAN: G 2 x 67.7 MB + D 2 x 153 MB of parameters / gradients + small exchange regions

    if args.workload in TRAIN_METRICS:
        comm, why = D.peer_comm(heap)
        res = bench_train(args.workload, D, args.steps, args.warmup, per_gpu_batch=(args.weak_batch or None), comm=comm,
                          comm_why=why, sampler=(sampler if rank == 0 else None))
        if rank == 0:
            res.update({"higher_is_better": True, "vs_baseline": None, "data": "synthetic"})
            print(json.dumps(res), flush=True)
    elif args.workload == "tiled_infer":
        res = bench_tiled(D, args.steps, args.warmup, args.tiled_lr, sampler=(sampler if rank == 0 else None))
        if rank == 0:
            res.update({"higher_is_better": True, "vs_baseline": None, "data": "synthetic"})
            print(json.dumps(res), flush=True)
    elif args.workload == "bandwidth":
        if rank == 0:
            print(json.dumps(bench_bandwidth_kernels()), flush=True)
    else:
        line, model = bench_rrdb_infer(args, D, sampler)
        extra = {}
        if args.workload == "all" and not args.no_extra:
            # each extra is guarded: a failure is reported in the line instead of losing the headline
            def guarded(name, fn):
                try:
                    r = fn()
                    if rank == 0:
                        extra[name] = r
                except Exception as e:   # noqa: BLE001
                    if rank == 0:
                        extra[name] = {"error": f"{type(e).__name__}: {e}"}
            guarded("tiled_infer", lambda: bench_tiled(D, max(2, min(args.steps, 3)), 3, args.tiled_lr, model=model))
            model.release()
            comm, why = D.peer_comm(heap)
            tsteps = max(3, min(args.steps, 10))
            guarded("esrgan_train", lambda: bench_train("esrgan_train", D, tsteps, 3, comm=comm, comm_why=why))
            if world > 1:
                guarded("esrgan_train_weak", lambda: bench_train("esrgan_train", D, tsteps, 3, per_gpu_batch=16, comm=comm,
                                                                  comm_why=why))
            guarded("srresnet_train", lambda: bench_train("srresnet_train", D, tsteps, 3, comm=comm, comm_why=why))
            if world == 1:
                guarded("bandwidth_kernels", bench_bandwidth_kernels)
        else:
            model.release()
        if rank == 0:
            if extra:
                line["extra"] = extra
            if world == 1 and not args.no_cpu_baseline:
                try:
                    line["cpu_baseline"] = cpu_baseline_sample()
                except Exception as e:   # noqa: BLE001 - the CPU arm must not cost the GPU line
                    line["cpu_baseline"] = {"error": f"{type(e).__name__}: {e}"}
            print(json.dumps(line), flush=True)
    D.close()


if __name__ == "__main__":
    main()
