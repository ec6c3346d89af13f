"""One training iteration of the SRResNet generator on the B200: the non-GAN branch of ``SRModel.train_step``
(simple_sr/models/sr_model.py:403-453) with pixel losses (``MeanSquaredError`` / ``MeanAbsoluteError``,
loss_functions/mean_squared_error.py:57-58, mean_absolute_error.py:57-58), the PSNR metric of
``_update_metrics`` (sr_model.py:630-634) and Keras Adam (sr_model.py:439-441).

What TensorFlow does with a GradientTape is laid out here as an explicit launch list over the C ABI:

  forward   : the inference launch list, except that every layer output stays resident and PReLU layers store their
              pre-activation z (Keras PReLU slopes start at 0, so min(0, z) cannot be recovered from the output)
  loss      : ssr_pixel_loss  -> MSE, MAE, per-image PSNR and d(loss)/d(sr) in one pass; tanh' applied in fp32
  backward  : per layer  dZ = dY * act'(z)            ssr_act_bwd_bf16
                         dbias, dalpha                ssr_channel_sum_bf16 (fixed-order reductions)
                         dW = X^T * dZ                ssr_conv2d_wgrad   (split-K tcgen05)
                         dX = dZ * rot180(W)^T        ssr_conv2d_fwd with dgrad-packed weights (+ skip gradient as `res`)
              depth_to_space backward = ssr_space_to_depth2
  update    : the optimizer clock (Keras `iterations`, learning-rate schedule, bias-corrected step size) lives on the
              device (ssr_opt_prepare), so the update is part of the step graph: Adam runs per BUCKET of layers, in the
              order the backward pass completes them, on a stream of its own next to the remaining dgrad chain.
              Data-parallel (``comm=PeerComm``): the bucket kernel is ssr_comm_adam_step - gradient reduce-scatter +
              Adam + parameter all-gather over peer memory (NVLink) in one launch; BatchNorm statistics are exchanged the
              same way (sync-BN), so a data-parallel step computes the single-device step of the global batch.
              Then ONE launch re-packs the bf16 weight images (forward + dgrad) from the updated masters.
  metrics   : copied to pinned host memory after the graph; ``train_step(..., lag=1)`` returns the previous step's, so
              the host never waits for the device (the reference syncs every step, sr_model.py:526-529).

All variables live in one flat fp32 device buffer in Keras order ([kernel, bias, (alpha)] per conv);
gradients, Adam m and v mirror that layout.  The whole iteration is captured in one CUDA graph per input shape.
"""
import math

import numpy as np

from . import _lib as L
from .model_builder import GeneratorModel


class PiecewiseConstantDecay:
    """``tf.keras.optimizers.schedules.PiecewiseConstantDecay(boundaries, values)`` as the ESRGAN recipes use it
    (examples/training/example_without_yaml.py:287-297): ``values[0]`` while step <= boundaries[0], ``values[i]`` while
    boundaries[i-1] < step <= boundaries[i], ``values[-1]`` afterwards.  Evaluated on the device by ssr_opt_prepare."""

    def __init__(self, boundaries, values):
        if len(values) != len(boundaries) + 1:
            raise ValueError("The length of boundaries should be 1 less than the length of values")
        self.boundaries = [int(b) for b in boundaries]
        self.values = [float(v) for v in values]

    def __call__(self, step):
        i = 0
        while i < len(self.boundaries) and step > self.boundaries[i]:
            i += 1
        return self.values[i]


class FlatAdam:
    """Keras Adam (sr_model.py:121-131) over one flat fp32 parameter buffer, clock and schedule on the device.

    ``comm`` (a :class:`simplesr_b200.parallel.PeerComm` of world > 1) puts parameters and gradients into the peer heap
    and turns every update into ssr_comm_adam_step (reduce-scatter + Adam + all-gather over NVLink)."""

    def __init__(self, host_flat, learning_rate, beta_1, beta_2, epsilon, stream_ptr, comm=None):
        self.count = int(host_flat.size)
        self.padded = -(-self.count // 4) * 4
        self.comm = comm if (comm is not None and comm.world > 1) else None
        flat = np.zeros(self.padded, np.float32)
        flat[:self.count] = host_flat
        nbytes = flat.nbytes
        if self.comm is not None:
            self.d_param, self.d_grad = self.comm.alloc(nbytes), self.comm.alloc(nbytes)
            self.site = None          # slots are taken per call site (emit)
        else:
            self.d_param, self.d_grad = L.DeviceBuffer(nbytes), L.DeviceBuffer(nbytes)
        self.d_param.upload(flat, stream_ptr)
        self.d_m, self.d_v = L.DeviceBuffer(nbytes), L.DeviceBuffer(nbytes)
        for b in (self.d_grad, self.d_m, self.d_v):
            b.zero(stream_ptr)
        self.b1, self.b2, self.eps = float(beta_1), float(beta_2), float(epsilon)
        self.state = L.DeviceBuffer(L.load().ssr_opt_state_bytes())
        L.check(L.load().ssr_opt_state_set(self.state.ptr, 0, stream_ptr))
        self.set_learning_rate(learning_rate, stream_ptr)
        self._sites = {}

    def set_learning_rate(self, learning_rate, stream_ptr=None):
        self.schedule = learning_rate if isinstance(learning_rate, PiecewiseConstantDecay) else None
        self.lr = float(learning_rate.values[0]) if self.schedule else float(learning_rate)
        if self.schedule:
            self.d_bounds = L.DeviceBuffer.from_numpy(np.asarray(self.schedule.boundaries + [0], np.int64), stream_ptr)
            self.d_values = L.DeviceBuffer.from_numpy(np.asarray(self.schedule.values, np.float32), stream_ptr)

    def learning_rate_at(self, iterations):
        return self.schedule(iterations) if self.schedule else self.lr

    def set_iterations(self, iterations, stream_ptr=None):
        L.check(L.load().ssr_opt_state_set(self.state.ptr, int(iterations), stream_ptr))

    def prepare(self, stream_ptr):
        """Advance the clock: once per step, before any ``update`` of that step."""
        if self.schedule:
            L.opt_prepare(self.state, self.lr, self.b1, self.b2, self.d_bounds, self.d_values,
                          len(self.schedule.boundaries), stream_ptr)
        else:
            L.opt_prepare(self.state, self.lr, self.b1, self.b2, None, None, 0, stream_ptr)

    def reserve(self, key):
        if self.comm is not None and key not in self._sites:
            self._sites[key] = self.comm.adam_site()

    def update(self, lo, hi, stream_ptr, key=None):
        """Adam over elements [lo, hi) (lo a multiple of 4).  ``key`` names the call site (its barrier slots)."""
        if self.comm is not None:
            k = key if key is not None else (lo, hi)
            if k not in self._sites:
                self._sites[k] = self.comm.adam_site()
            self.comm.adam_step(self._sites[k], self.d_grad, self.d_param, self.d_m, self.d_v, lo, hi, self.state, self.b1,
                                self.b2, self.eps, stream_ptr)
        else:
            off = lo * 4
            view = lambda b: L.DeviceView(b, off, (hi - lo) * 4)
            L.adam_step_dev(view(self.d_param), view(self.d_grad), view(self.d_m), view(self.d_v), hi - lo, self.state,
                            self.b1, self.b2, self.eps, 1.0, stream_ptr)

    def iterations(self, stream_ptr=None):
        it = L.C.c_int64()
        L.check(L.load().ssr_opt_state_get(self.state.ptr, L.C.byref(it), None, stream_ptr))
        return int(it.value)


class _PlanBuilder:
    """Collects the device buffers and the launch list of one training iteration (shared by the two generators)."""

    def __init__(self, trainer):
        self.tr, self.ctx = trainer, trainer.ctx
        self.B, self.ops = {}, []
        self.wg_ws_bytes = 0
        self.fuse_bias_grad = getattr(trainer, "fuse_bias_grad", True)
        self._wg_pending = None
        # weight gradients on a second stream: they only feed the optimizer, so they overlap the dgrad chain
        # (fork = the main stream's dZ is ready; joins are placed by the trainer where buffers are reused)
        self.side = trainer.side_stream() if getattr(trainer, "overlap_wgrad", False) else None
        self.events = []
        # optimizer updates: in the graph on a third stream (per bucket), or collected for eager execution after the
        # fallback all-reduce hook
        self.in_graph_update = trainer.allreduce is None
        self.update_stream = trainer.update_stream() if self.in_graph_update else None
        self.update_ops = []
        self._updated = 0
        if self.in_graph_update:
            opt = trainer.opt
            self.ops.append(lambda s: opt.prepare(s))

    def _event(self):
        ev = L.Event()
        self.events.append(ev)
        return ev

    def side_mark(self):
        """Event on the side stream after everything queued there so far (None without a side stream)."""
        self.flush_wgrad()
        if self.side is None:
            return None
        ev, side = self._event(), self.side
        self.ops.append(lambda s: ev.record(side.ptr))
        return ev

    def main_wait(self, ev):
        if ev is not None:
            self.ops.append(lambda s: L.stream_wait_event(s, ev))

    def join(self):
        self.main_wait(self.side_mark())

    def bucket_ready(self, names):
        """Every gradient of the convs ``names`` (contiguous in Keras order) has been queued (main + side stream) and their
        own dgrad launches are behind us: the Adam update of that flat range (with the gradient exchange, data-parallel)
        and the re-pack of their weight images start on the update stream as soon as those kernels finish."""
        lo, hi = self.tr.range_of(names)
        if hi <= lo:
            return
        self.flush_wgrad()
        self._updated += hi - lo
        tr, opt = self.tr, self.tr.opt
        opt.reserve((lo, hi))     # barrier slots in plan-building order: the same on every rank
        if not self.in_graph_update:
            self.update_ops.append(lambda s: opt.update(lo, hi, s, key=(lo, hi)))
            return
        ev_main, ev_side, up = self._event(), (self._event() if self.side is not None else None), self.update_stream
        side = self.side

        def launch(s):
            ev_main.record(s)
            up.wait_event(ev_main)
            if side is not None:
                ev_side.record(side.ptr)
                up.wait_event(ev_side)
            opt.update(lo, hi, up.ptr, key=(lo, hi))
            tr._repack(up.ptr, names)
        self.ops.append(launch)

    def finish_update(self):
        """All buckets are queued: join the update stream, re-pack the weight images, let the loss functors apply their
        own updates (the discriminator's, sr_model.py:444-451)."""
        tr = self.tr
        assert self._updated == tr.opt.padded, f"buckets cover {self._updated} of {tr.opt.padded} parameters"
        target = self.ops if self.in_graph_update else self.update_ops
        if self.in_graph_update:
            done, up = self._event(), self.update_stream
            target.append(lambda s: (done.record(up.ptr), L.stream_wait_event(s, done)))   # images re-packed per bucket
        else:
            target.append(lambda s: tr._repack(s))
        for el in tr.extra_losses:
            if hasattr(el, "emit_update"):
                el.emit_update(target)

    def buf(self, name, nbytes):
        self.B[name] = L.DeviceBuffer(nbytes)
        return self.B[name]

    def add(self, fn):
        self.ops.append(fn)

    def conv(self, conv_, n_, h_, w_, x, xcs, out, ocs, ocoff=0, packed=None, cin=None, cout=None, kh=None, kw=None,
             up=None, res=None, res_cs=None, res_coff=0, res_beta=1.0, bias=True, act=L.ACT_NONE, act_alpha=0.0,
             mask=None, out_dtype=L.SSR_BF16):
        # mask = (z, z_cstride, z_coff, lo, n, alpha, dz_out, dz_cstride): fused activation backward (ssr_conv2d_fwd_mask)
        ctx = self.ctx
        d = L.ConvDesc(n=n_, h=h_, w=w_, cin=cin or conv_.cin, in_cstride=xcs, cout=cout or conv_.cout,
                       ksize=kh or conv_.kh, ksize_w=(kw if kw is not None else conv_.kw), act=act,
                       act_alpha=act_alpha, res_beta=res_beta, up=(up if up is not None else conv_.up),
                       out_dtype=out_dtype, out_cstride=ocs, out_coff=ocoff,
                       res_dtype=(L.SSR_BF16 if res is not None else L.SSR_NONE), res_cstride=(res_cs or ocs),
                       res_coff=res_coff, out2_cstride=0, out2_coff=0)
        # always the trainer's own images / bias views: they follow the flat masters after every update
        pk = packed or self.tr.fwd_packed.get(conv_.name, conv_.d_packed)
        bs = self.tr.fwd_bias.get(conv_.name, conv_.d_bias) if bias else None
        if mask is not None:
            mz, mzcs, mzoff, mlo, mn, malpha, mout, mocs = mask
            self.ops.append(lambda s: ctx.conv2d_fwd_mask(d, x, pk, bs, res, out, mz, mzcs, mzoff, mlo, mn, malpha, mout, mocs,
                                                          stream=s))
            return
        self.ops.append(lambda s: ctx.conv2d_fwd(d, x, pk, bs, out, res=res, stream=s))

    # Weight gradients are BATCHED: requests over tensors of the same n, h, w and kernel size collect in a pending group
    # (up to 8 convolutions / 64 work units) and go out as ONE ssr_conv2d_wgrad_multi launch + one reduction when the
    # group is flushed - explicitly (the five convolutions of a dense block share the block's input buffer), or before any
    # event that later work waits on.  At training-patch sizes a lone wgrad launch spends most of its ~27 us draining
    # per-CTA split-K partials; batched, the units of all convolutions share the CTAs and the drain happens once.
    def wgrad(self, name, x, xcs, cin_real, dz, zcs, cout, n_, h_, w_, kh, kw, scale=1.0, xoff=0, zoff=0):
        geom = (n_, h_, w_, kh, kw)
        units = -(-cin_real // 64) * -(-cout // 64) * -(-(kh * kw) // 16)
        pend = self._wg_pending
        if pend and (pend["geom"] != geom or len(pend["items"]) >= 8 or pend["units"] + units > 64
                     or not getattr(self.tr, "batch_wgrad", True)):
            self.flush_wgrad()
            pend = None
        if not pend:
            pend = self._wg_pending = dict(geom=geom, items=[], units=0)
        dw = self.tr._view(self.tr.layout[name]["k"], self.tr.d_grad)
        pend["items"].append(dict(name=name, x=x, xcs=xcs, xoff=xoff, cin=cin_real, dz=dz, zcs=zcs, zoff=zoff, cout=cout,
                                  dw=dw, scale=scale, db=None, bscale=1.0))
        pend["units"] += units

    def bias_grad(self, name, dz, zcs, cout, pixels, scale=1.0, zoff=0):
        B = self.B
        db = self.tr._view(self.tr.layout[name]["b"], self.tr.d_grad)
        pend = self._wg_pending
        it = pend["items"][-1] if pend and pend["items"] else None
        if (self.fuse_bias_grad and it is not None and it["name"] == name and it["dz"] is dz and it["zcs"] == zcs
                and it["zoff"] == zoff and it["cout"] == cout and pend["geom"][3] * pend["geom"][4] <= 14):
            # BiasAddGrad rides in the wgrad kernel (one more accumulator over a tile of ones): two launches fewer
            it["db"], it["bscale"] = db, scale
            return
        self.ops.append(lambda s: L.channel_sum_bf16(dz, zcs, zoff, None, 0, 0, pixels, cout, scale, False, B["cs_ws"], db, s))

    def flush_wgrad(self):
        """Emit the pending group of weight gradients: fork from the main stream (their dZ are final), one batched launch."""
        pend, self._wg_pending = self._wg_pending, None
        if not pend or not pend["items"]:
            return
        ctx, B, side = self.ctx, self.B, self.side
        n_, h_, w_, kh, kw = pend["geom"]
        items = [ctx.wgrad_item(i["x"], i["xcs"], i["xoff"], i["cin"], i["dz"], i["zcs"], i["zoff"], i["cout"], i["dw"],
                                scale=i["scale"], dbias=i["db"], bias_scale=i["bscale"]) for i in pend["items"]]
        self.wg_ws_bytes = max(self.wg_ws_bytes, ctx.conv_wgrad_multi_workspace_bytes(items, h_, w_, kh, kw))
        if side is not None:
            ev = self._event()
            self.ops.append(lambda s: (ev.record(s), side.wait_event(ev)))
        self.ops.append(lambda s: ctx.conv2d_wgrad_multi(items, n_, h_, w_, kh, kw, B["wg_ws"],
                                                         stream=(side.ptr if side else s)))

    def prelu_bwd(self, name, dy, z, ch, pixels, dz_out):
        """dalpha = sum dy * min(0, z); dz = dy * prelu'(z)."""
        B = self.B
        da = self.tr._view(self.tr.layout[name]["a"], self.tr.d_grad)
        al = self.tr.model.convs[name].d_alpha
        self.ops.append(lambda s: L.channel_sum_bf16(dy, ch, 0, z, ch, 0, pixels, ch, 1.0, False, B["cs_ws"], da, s))
        self.ops.append(lambda s: L.act_bwd_bf16(dy, ch, 0, z, ch, 0, al, 0.0, dz_out, ch, 0, pixels, ch, s))

    def finish(self, n, H, W):
        self.flush_wgrad()
        self.buf("wg_ws", max(self.wg_ws_bytes, 16))
        return dict(buffers=self.B, ops=self.ops, graph=None, n=n, H=H, W=W, events=self.events,
                    update_ops=self.update_ops)


class _TrainerBase:
    """Owns the flat parameter / gradient / optimizer-state buffers of a generator and runs train steps.

    ``loss`` is ``("mse" | "mae", weight)`` or a list of those; ``learning_rate`` a float or a
    :class:`PiecewiseConstantDecay`.  Data parallel: ``comm`` (a :class:`simplesr_b200.parallel.PeerComm`) runs the
    gradient exchange, sync-BatchNorm and the metric means inside the step graph over peer memory; ``allreduce``
    (fallback, no peer mapping) is called as ``allreduce(grad_buffer, count, stream_ptr)`` between backward and Adam,
    outside the graph, and must leave the MEAN over ranks in the buffer.
    """

    def __init__(self, model, loss=("mse", 1.0), learning_rate=1e-3, beta_1=0.9, beta_2=0.999, epsilon=1e-7,
                 allreduce=None, extra_losses=(), comm=None, buckets=4):
        # extra_losses: device-side loss functors with emit(ops, B, prefix, n, H, W, hr, sr, g_sr) -> out buffer and a
        # .name, e.g. simplesr_b200.vgg.VGGLoss (generator.py:220-228 sums the functors; so do the gradients)
        self.has_bn = bool(model.non_trainable_variables)
        if self.has_bn and not getattr(self, "SUPPORTS_BN", False):
            raise NotImplementedError("training this generator with batch normalisation is not built; build it with "
                                      "batch_normalization=False")
        self.extra_losses = list(extra_losses)
        if not isinstance(model, GeneratorModel) or model.architecture != self.ARCH:
            raise ValueError(f"{type(self).__name__} needs a {self.ARCH} model from simplesr_b200.model_builder")
        if comm is not None and allreduce is not None:
            raise ValueError("pass either comm (peer-memory fabric) or allreduce (fallback hook), not both")
        self.model = model
        model.fuse_growth = False      # the paired inference weights derive from host variables, stale under training
        model.release()
        self.ctx, self.stream = model.ctx, model.stream
        losses = [loss] if isinstance(loss[0], str) else list(loss)
        self.w_mse = float(sum(w for k, w in losses if k == "mse"))
        self.w_mae = float(sum(w for k, w in losses if k == "mae"))
        if any(k not in ("mse", "mae") for k, _ in losses):
            raise ValueError("supported pixel losses: 'mse', 'mae'")
        self.b1, self.b2, self.eps = float(beta_1), float(beta_2), float(epsilon)
        self.allreduce = allreduce
        self.comm = comm if (comm is not None and comm.world > 1) else None
        self.num_buckets = max(1, int(buckets))
        self.iterations = 0
        self._plans = {}
        self._pending = None        # (pinned slot, event, plan) of the step whose metrics have not been read yet
        self._last_shape = None     # (n, h, w) of the most recent LR batch
        self._metric_slots = None
        for el in self.extra_losses:
            if hasattr(el, "attach"):
                el.attach(self)
        self._build_flat(learning_rate)

    @property
    def lr(self):
        return self.opt.learning_rate_at(self.iterations)

    # ---- flat parameter buffer ----------------------------------------------------------------------------------
    def _build_flat(self, learning_rate):
        m = self.model
        self.layout = {}          # conv name -> dict(k=(off,size), b=(off,size), a=(off,size)|None)
        off = 0
        host = []
        for name, c in m.convs.items():
            ent = {}
            for key, var in (("k", c.kernel), ("b", c.bias), ("a", c.alpha),
                             ("g", c.bn["gamma"] if c.bn else None), ("be", c.bn["beta"] if c.bn else None)):
                if var is None:
                    ent[key] = None
                    continue
                arr = var.numpy().ravel()
                ent[key] = (off, arr.size)
                host.append(arr)
                off += arr.size
            self.layout[name] = ent
        self.count = off
        flat = np.concatenate(host).astype(np.float32)
        s = self.stream.ptr
        self.opt = FlatAdam(flat, learning_rate, self.b1, self.b2, self.eps, s, comm=self.comm)
        self.d_param, self.d_grad, self.d_m, self.d_v = self.opt.d_param, self.opt.d_grad, self.opt.d_m, self.opt.d_v
        # dgrad weight images + redirect the forward convs to the flat masters
        self.dgrad_packed = {}
        self.fwd_packed, self.fwd_bias, self.bn_moving = {}, {}, {}
        for name, c in m.convs.items():
            c.sync(self.ctx, s)                       # allocates c.d_packed
            ent = self.layout[name]
            if self.has_bn:
                # batch norm: inference folds the moving statistics into c.d_packed / c.d_bias (model_builder._Conv.sync);
                # the training forward needs the raw conv, so it gets images of its own and the model's stay untouched
                self.fwd_packed[name] = L.DeviceBuffer(self.ctx.conv_packed_bytes(c.kh, c.cin, c.cout, c.up, ksize_w=c.kw))
                self.fwd_bias[name] = self._view(ent["b"])
                if c.bn is not None:
                    self.bn_moving[name] = (L.DeviceBuffer.from_numpy(c.bn["moving_mean"].numpy(), s),
                                            L.DeviceBuffer.from_numpy(c.bn["moving_variance"].numpy(), s))
            else:
                c.d_bias = self._view(ent["b"])
            if ent["a"] is not None:
                c.d_alpha = self._view(ent["a"])
            if name in self.NO_DGRAD or self._slice_dgrad(name):
                continue                              # the network input needs no gradient / composed slice images instead
            if name in self.UNROLLED_DGRAD:
                nbytes = self.ctx.conv_packed_bytes(c.kh, 32, c.cin_real, 1, ksize_w=1)   # x-unrolled 27 -> 32
            else:
                nbytes = self.ctx.conv_packed_bytes(c.kh, -(-c.cout // 16) * 16, c.cin_real, 1, ksize_w=c.kw)
            self.dgrad_packed[name] = L.DeviceBuffer(nbytes)
        self._build_slice_images()
        self._repack(s)
        self.stream.sync()
        self._install_hooks()

    def _slice_dgrad(self, name):
        """True for convolutions whose input gradient is produced by composed slice images (RRDB dense blocks)."""
        return False

    def _build_slice_images(self):
        self.slice_packed, self._slice_items = {}, {}

    def side_stream(self):
        """Second stream for work that only feeds the optimizer (weight gradients); created on first use."""
        if getattr(self, "_side", None) is None:
            self._side = L.Stream()
        return self._side

    def update_stream(self):
        """Third stream: the per-bucket optimizer updates (and, data-parallel, the gradient exchange)."""
        if getattr(self, "_update", None) is None:
            self._update = L.Stream()
        return self._update

    def _view(self, ent, buf=None):
        off, size = ent
        return L.DeviceView(buf or self.d_param, off * 4, size * 4)

    def range_of(self, names):
        """Flat [lo, hi) covered by the variables of convs ``names`` (contiguous in Keras order); the last conv of the
        model extends to the 4-float padding of the buffer."""
        ents = [e for n in names for e in self.layout[n].values() if e is not None]
        lo, hi = min(e[0] for e in ents), max(e[0] + e[1] for e in ents)
        if hi == self.count:
            hi = self.opt.padded
        if lo % 4:
            raise ValueError(f"bucket boundary {lo} is not a multiple of 4 floats")
        return lo, hi

    def _pack_items(self):
        """The device table of every weight image, in Keras conv order; ``self._pack_range[name]`` = [first, last) entries
        that derive from the variables of conv ``name`` (a block's composed slice images hang on its last conv)."""
        if getattr(self, "_pack_table", None) is None:
            items, self._pack_range = [], {}
            for name, c in self.model.convs.items():
                first = len(items)
                k = self._view(self.layout[name]["k"])
                fwd = self.fwd_packed.get(name, c.d_packed)
                items.append(L.PackItem(k.ptr, fwd.ptr, c.kh, c.kw, c.cin_real, c.cin, c.cout, c.up, 0, 0))
                if name in self.dgrad_packed:
                    mode = 2 if name in self.UNROLLED_DGRAD else 1
                    items.append(L.PackItem(k.ptr, self.dgrad_packed[name].ptr, c.ksize, c.ksize, c.cin_real, c.cin,
                                            c.cout, 1, mode, 0))
                items.extend(self._slice_items.get(name, []))
                self._pack_range[name] = (first, len(items))
            self._pack_count = len(items)
            self._pack_table = self.ctx.pack_batch_prepare(items, self.stream.ptr)
        return self._pack_table

    def _repack(self, s, names=None):
        """bf16 weight images (forward, dgrad, composed slices) from the fp32 masters in the flat buffer: ONE launch over
        the device table - all of it, or the entries of the convs ``names`` (contiguous in Keras order: one optimizer
        bucket, re-packed on the update stream right behind its Adam while the backward pass continues)."""
        table = self._pack_items()
        if names is None:
            self.ctx.pack_batch(table, self._pack_count, s)
            return
        lo = min(self._pack_range[n][0] for n in names)
        hi = max(self._pack_range[n][1] for n in names)
        self.ctx.pack_batch(L.DeviceView(table, lo * L.PACK_ENTRY_BYTES, (hi - lo) * L.PACK_ENTRY_BYTES), hi - lo, s)

    def _install_hooks(self):
        """``variable.numpy()`` on the model reads the trained values back from the device; ``variable.assign`` (and so
        ``set_weights`` / ``load_weights``: resuming from a checkpoint after the trainer exists) writes into the flat
        master buffer and refreshes the packed images."""
        trainer = self

        def make_pull(var, ent):
            def pull():
                off, size = ent
                var._value = trainer.d_param.download((size,), np.float32, trainer.stream.ptr, offset=off * 4).reshape(
                    var._value.shape)
            return pull

        def make_push(ent):
            def push(value):
                trainer.flush()
                trainer.d_param.upload(np.ascontiguousarray(value, np.float32).ravel(), trainer.stream.ptr, offset=ent[0] * 4)
                trainer._repack(trainer.stream.ptr)
                trainer.stream.sync()
            return push

        for name, c in self.model.convs.items():
            ent = self.layout[name]
            for key, var in (("k", c.kernel), ("b", c.bias), ("a", c.alpha),
                             ("g", c.bn["gamma"] if c.bn else None), ("be", c.bn["beta"] if c.bn else None)):
                if var is not None:
                    var._pull = make_pull(var, ent[key])
                    var._push = make_push(ent[key])
            if name in self.bn_moving:
                def make_pull_moving(var, dbuf):
                    def pull():
                        var._value = dbuf.download(var._value.shape, np.float32, trainer.stream.ptr)
                    return pull

                def make_push_moving(dbuf):
                    def push(value):
                        trainer.flush()
                        dbuf.upload(np.ascontiguousarray(value, np.float32), trainer.stream.ptr)
                    return push
                for var, dbuf in zip((c.bn["moving_mean"], c.bn["moving_variance"]), self.bn_moving[name]):
                    var._pull = make_pull_moving(var, dbuf)
                    var._push = make_push_moving(dbuf)

    def gradients(self):
        """Host copies of the last step's gradients, ``{conv name: (dkernel, dbias, dalpha|None)}``.  Data-parallel:
        this rank's gradients (the exchange averages them inside the Adam kernel, not in place)."""
        self.flush()
        flat = self.d_grad.download((self.count,), np.float32, self.stream.ptr)
        out = {}
        for name, c in self.model.convs.items():
            ent = self.layout[name]
            get = lambda e, shape: None if e is None else flat[e[0]:e[0] + e[1]].reshape(shape)
            out[name] = (get(ent["k"], c.kernel.shape), get(ent["b"], c.bias.shape),
                         get(ent["a"], c.alpha.shape if c.alpha is not None else None))
            if ent.get("g") is not None:
                out[name + "_bn"] = (get(ent["g"], (c.cout,)), get(ent["be"], (c.cout,)))
        return out

    # ---- public ---------------------------------------------------------------------------------------------------------
    def train_step(self, lr_batch, hr_batch, use_graph=True, lag=0):
        """One iteration: forward, loss, backward, gradient exchange, Adam, weight re-pack - all on the device.
        Returns ``{"loss", "mse", "mae", "psnr", <functor names>}`` (psnr = batch mean of
        tf.image.psnr(hr, sr, max_val=2.0)); data-parallel: means over all ranks.  ``lag=1`` returns the metrics of the
        PREVIOUS step (None on the first call) so the host never waits for the device; ``last_metrics()`` reads the
        newest."""
        lr = np.ascontiguousarray(lr_batch, dtype=np.float32)
        hr = np.ascontiguousarray(hr_batch, dtype=np.float32)
        n, h, w, _ = lr.shape
        sf = self.model.upsample_factor
        if hr.shape != (n, h * sf, w * sf, 3):
            raise ValueError(f"hr batch shape {hr.shape} does not match lr batch {lr.shape} at scale {sf}")
        plan = self._plan(n, h, w)
        self._last_shape = (n, h, w)
        s = self.stream.ptr
        B = plan["buffers"]
        L.check(self.ctx.lib.ssr_memcpy_h2d(B["in_f32"].ptr, lr.ctypes.data, lr.nbytes, s))
        L.check(self.ctx.lib.ssr_memcpy_h2d(B["hr_f32"].ptr, hr.ctypes.data, hr.nbytes, s))
        for el in self.extra_losses:
            if hasattr(el, "pre_step"):
                el.pre_step(s)
        self._run(plan, s, use_graph)
        if self.allreduce is not None:
            # fallback exchange (NCCL through torch.distributed), then the same update launches, eagerly
            self.allreduce(self.d_grad, self.count, s)
            self.opt.prepare(s)
            for op in plan["update_ops"]:
                op(s)
        self.iterations += 1
        if self.has_bn:
            for c in self.model.convs.values():
                c.dirty = True     # the model's folded inference images are stale after every update
        for el in self.extra_losses:
            if hasattr(el, "post_step"):
                el.post_step(s)
        # metrics: one small device buffer -> pinned host slot; read now (lag 0) or at the next call (lag 1)
        slots = self._metric_slots
        k = self.iterations & 1
        L.check(self.ctx.lib.ssr_memcpy_d2h(slots[k].ptr, plan["metrics"].ptr, plan["metrics_count"] * 4, s))
        self._metric_events[k].record(s)
        prev, self._pending = self._pending, (k, plan)
        if lag:
            return self._read_metrics(*prev) if prev is not None else None
        return self._read_metrics(k, plan)

    def last_metrics(self):
        """Metrics of the most recent step (waits for it)."""
        return self._read_metrics(*self._pending) if self._pending is not None else None

    def last_sr(self):
        """The generated batch ``[n, sf*h, sf*w, 3]`` fp32 of the most recent step (waits for it) - what the reference
        hands to its image-metric functions after the update (sr_model.py:453)."""
        if self._pending is None or self._last_shape is None:
            return None
        _, plan = self._pending
        n, h, w = self._last_shape
        sf = self.model.upsample_factor
        self.stream.sync()
        return plan["buffers"]["out_f32"].download((n, h * sf, w * sf, 3), np.float32, self.stream.ptr)

    def prepare(self, n, h, w):
        """Build the launch list for an [n,h,w,3] LR batch and capture its graph without running it.  Data-parallel
        ranks call this before their first step so that they enter it together (the in-graph barriers wait for the
        slowest rank, but only for the spin limit of the fabric)."""
        plan = self._plan(n, h, w)
        if plan["graph"] is None and not (self.comm is not None and self.comm.mode == "local"):
            s = self.stream.ptr
            plan["graph"] = L.Graph(s, lambda: [op(s) for op in plan["ops"]])
            plan["kernel_launches"] = plan["graph"].kernels + (len(plan["update_ops"]) if self.allreduce else 0)
        return plan

    def _read_metrics(self, k, plan):
        self._metric_events[k].sync()
        out = self._metric_slots[k].array
        n = plan["n"]
        res = {"loss": float(self.w_mse * out[0] + self.w_mae * out[1]), "mse": float(out[0]), "mae": float(out[1]),
               "psnr": float(np.mean(out[2:2 + n]))}
        for el, off in zip(self.extra_losses, plan["extra_off"]):
            v = float(out[off]) * getattr(el, "metric_scale", el.loss_weight * el.feature_scale ** 2)
            res[el.name] = v
            res["loss"] += v
            if hasattr(el, "metric_names"):
                for j, name in enumerate(el.metric_names):
                    if name != el.name:
                        res[name] = float(out[off + j])
        return res

    def flush(self):
        """Wait for everything queued on the trainer's stream (steps run asynchronously with ``lag=1``)."""
        self.stream.sync()
        if self.comm is not None:
            self.comm.check()

    overlap_losses = True   # independent loss functors (VGG, RaGAN) run on their own streams, joined before the backward

    def _metrics_layout(self, pb, n):
        """One device buffer for every metric of the step: [mse, mae, psnr x n | functor 0 | functor 1 ...]; the
        functors write their slices directly.  Data-parallel: a second buffer receives the mean over the ranks."""
        sizes = [2 + n] + [getattr(el, "metric_floats", lambda n_: 2 + n_)(n) for el in self.extra_losses]
        offs = np.cumsum([0] + sizes)
        raw = pb.buf("metrics_raw", int(offs[-1]) * 4)
        views = [L.DeviceView(raw, int(o) * 4, int(sz) * 4) for o, sz in zip(offs[:-1], sizes)]
        return raw, views, [int(o) for o in offs[1:-1]], int(offs[-1])

    def _finish_metrics(self, pb, plan, raw, extra_off, count):
        plan["extra_off"], plan["metrics_count"] = extra_off, count
        # pinned landing slots for the metrics, allocated here (not inside a step: allocations may synchronise the device)
        if self._metric_slots is None or self._metric_slots[0].shape[0] < count:
            self._metric_slots = [L.PinnedArray((max(64, count),), np.float32) for _ in range(2)]
            self._metric_events = [L.Event(), L.Event()]
        if self.comm is None:
            plan["metrics"] = raw
            return
        avg = pb.buf("metrics_avg", count * 4)
        site, comm = self.comm.allreduce_site(count), self.comm
        # appended after the update: the last launches of the graph
        plan["ops"].append(lambda s: comm.allreduce_f32(site, raw, avg, count, 1.0 / comm.world, s))
        plan["metrics"] = avg

    def _emit_extra(self, pb, n, H, W, hr_f32, sr, g_sr, outs_views):
        """Launches of the extra loss functors (generator.py:220-228 sums them).  Functor 0 runs on the main stream and
        adds its image gradient to ``g_sr`` directly; with ``overlap_losses`` every further functor runs concurrently on
        a stream of its own into a private gradient buffer that is added to ``g_sr`` after the join."""
        if not self.overlap_losses or len(self.extra_losses) < 2:
            return [el.emit(pb.ops, pb.B, f"extra{i}_", n, H, W, hr_f32, sr, g_sr, accumulate=True, out=outs_views[i])
                    for i, el in enumerate(self.extra_losses)]
        ops = L.OpsView(pb.ops)
        count = n * H * W * 3
        fork = pb._event()
        ops.append(lambda s: fork.record(s))
        pending, outs = [], [None] * len(self.extra_losses)
        for i, el in enumerate(self.extra_losses):
            if i == 0:
                continue
            st = self._loss_stream(i)
            g_i = pb.buf(f"extra{i}_g_sr", count * 4)
            done = pb._event()
            ops.append(lambda s, st=st: st.wait_event(fork))
            ops.redirect = st
            ops.append(lambda s, g_i=g_i: L.check(L.load().ssr_memset(g_i.ptr, 0, g_i.nbytes, s)))
            outs[i] = el.emit(ops, pb.B, f"extra{i}_", n, H, W, hr_f32, sr, g_i, accumulate=True, out=outs_views[i])
            ops.redirect = None
            ops.append(lambda s, st=st, done=done: done.record(st.ptr))
            pending.append((g_i, done))
        outs[0] = self.extra_losses[0].emit(ops, pb.B, "extra0_", n, H, W, hr_f32, sr, g_sr, accumulate=True,
                                            out=outs_views[0])
        for g_i, done in pending:
            ops.append(lambda s, done=done: L.stream_wait_event(s, done))
            ops.append(lambda s, g_i=g_i: L.axpy_f32(g_i, g_sr, 1.0, count, s))
        return outs

    def _loss_stream(self, i):
        if not hasattr(self, "_loss_streams"):
            self._loss_streams = {}
        if i not in self._loss_streams:
            self._loss_streams[i] = L.Stream()
        return self._loss_streams[i]

    def _run(self, plan, s, use_graph):
        if use_graph and not (self.comm is not None and self.comm.mode == "local"):   # emulated ranks: eager only
            if plan["graph"] is None:
                plan["graph"] = L.Graph(s, lambda: [op(s) for op in plan["ops"]])
                plan["kernel_launches"] = plan["graph"].kernels + (len(plan["update_ops"]) if self.allreduce else 0)
            plan["graph"].launch(s)
        else:
            for op in plan["ops"]:
                op(s)

    def launches_per_step(self, n, h, w):
        """Kernels of one step (counted by the context while the plan's launch list runs eagerly once is not possible
        without side effects, so: launch-list entries that are kernels; events / waits are not counted)."""
        plan = self._plan(n, h, w)
        return int(plan.get("kernel_launches") or len(plan["ops"]))

    def release(self):
        self.stream.sync()
        for plan in self._plans.values():
            if plan["graph"] is not None:
                plan["graph"].destroy()
            for b in plan["buffers"].values():
                b.free()
            for ev in plan.get("events", []):
                ev.destroy()
        self._plans = {}
        for attr in ("_side", "_update"):
            if getattr(self, attr, None) is not None:
                getattr(self, attr).destroy()
                setattr(self, attr, None)
        for st in getattr(self, "_loss_streams", {}).values():
            st.destroy()
        self._loss_streams = {}


class SRResNetTrainer(_TrainerBase):
    """Training iterations of a ``build_resnet`` model (model_builder.py:99-134), with or without batch normalisation."""

    ARCH = "srresnet"
    SUPPORTS_BN = True      # batch_normalization=True: batch statistics in the forward pass, full BN backward
    overlap_wgrad = True    # every gradient buffer of the backward pass is written once: no reuse hazards
    NO_DGRAD = ("first",)
    UNROLLED_DGRAD = ("last",)   # 9x9x64->3: dgrad over the x-unrolled dZ

    # ---- launch list ------------------------------------------------------------------------------------------------
    def _plan(self, n, h, w):
        key = (n, h, w)
        if key in self._plans:
            return self._plans[key]
        m, ctx = self.model, self.ctx
        c = m.convs
        nf, nb, sf = m.config["num_filters"], m.config["num_res_blocks"], m.upsample_factor
        nup = int(math.log(sf, 2))
        px = n * h * w
        pb = _PlanBuilder(self)
        B, ops = pb.B, pb.ops
        buf, conv, wgrad, bias_grad, prelu_bwd = pb.buf, pb.conv, pb.wgrad, pb.bias_grad, pb.prelu_bwd

        # ------------------------------------------------------------------ forward (all activations stay resident)
        in_f32 = buf("in_f32", px * 3 * 4)
        hr_f32 = buf("hr_f32", px * sf * sf * 3 * 4)
        x32 = buf("x_unrolled", px * 32 * 2)
        ops.append(lambda s: L.im2col_x_f32_to_bf16(in_f32, x32, n, h, w, 3, 9, 32, s))
        z_first, y_first = buf("z_first", px * nf * 2), buf("y_first", px * nf * 2)
        conv(c["first"], n, h, w, x32, 32, z_first, nf)
        ops.append(lambda s: L.act_fwd_bf16(z_first, nf, 0, c["first"].d_alpha, 0.0, y_first, nf, 0, px, nf, s))
        # BatchNormalization(training=True) after a conv (model_builder.py:291-292): batch statistics + moving-average
        # update (ssr_bn_stats_bf16), then the affine map (ssr_bn_lrelu_fwd_bf16 with slope 1 = no activation).  The conv
        # output z (pre-normalisation) and the BN output y are both kept for the backward pass.
        bn_state = {}
        if self.has_bn:
            buf("bn_ws", L.load().ssr_bn_workspace_bytes(nf))
            buf("bn_sums", 2 * nf * 4)

        comm = self.comm

        def bn_fwd(name, z, pixels):
            ent = self.layout[name]
            y, mean, istd = buf(f"bn_y_{name}", pixels * nf * 2), buf(f"bn_mean_{name}", nf * 4), buf(f"bn_istd_{name}", nf * 4)
            g, be = self._view(ent["g"]), self._view(ent["be"])
            mm, mv = self.bn_moving[name]
            mom, eps = c[name].bn_momentum, c[name].bn_eps
            site = comm.bn_site(nf) if comm is not None else None      # sync-BN: statistics of the global batch
            ops.append(lambda s: L.bn_stats_bf16(z, pixels, nf, eps, mom, B["bn_ws"], mean, istd, mm, mv, s, site=site))
            ops.append(lambda s: L.bn_lrelu_fwd_bf16(z, mean, istd, g, be, 1.0, y, pixels, nf, s))
            bn_state[name] = (z, y, mean, istd)
            return y

        def bn_bwd(name, dy, pixels):
            """dz of the conv output from the gradient dy of the BN output; dgamma / dbeta into the gradient buffer."""
            z, y, mean, istd = bn_state[name]
            ent = self.layout[name]
            g = self._view(ent["g"])
            dg, dbe = self._view(ent["g"], self.d_grad), self._view(ent["be"], self.d_grad)
            dz = buf(f"bn_dz_{name}", pixels * nf * 2)
            site = comm.bn_site(nf) if comm is not None else None
            ops.append(lambda s: L.bn_lrelu_bwd_bf16(z, dy, y, mean, istd, g, 1.0, pixels, nf, B["bn_ws"], B["bn_sums"], dg,
                                                     dbe, False, dz, s, site=site))
            return dz

        t = y_first
        t_in, z0s, us = [], [], []
        for b in range(nb):
            n0, n1 = f"res{b}_conv0", f"res{b}_conv1"
            z0, u, t_out = buf(f"z0_{b}", px * nf * 2), buf(f"u_{b}", px * nf * 2), buf(f"t_{b}", px * nf * 2)
            conv(c[n0], n, h, w, t, nf, z0, nf)
            pre0 = bn_fwd(n0, z0, px) if c[n0].bn is not None else z0          # what PReLU sees
            al = c[n0].d_alpha
            ops.append(lambda s, pre0=pre0, u=u, al=al: L.act_fwd_bf16(pre0, nf, 0, al, 0.0, u, nf, 0, px, nf, s))
            if c[n1].bn is not None:
                z1 = buf(f"z1_{b}", px * nf * 2)
                conv(c[n1], n, h, w, u, nf, z1, nf)
                y1 = bn_fwd(n1, z1, px)
                ops.append(lambda s, t=t, y1=y1, t_out=t_out: L.axpby_bf16(t, nf, 0, y1, nf, 0, 1.0, t_out, nf, 0, px, nf, s))
            else:
                conv(c[n1], n, h, w, u, nf, t_out, nf, res=t)
            t_in.append(t)
            z0s.append(pre0)
            us.append(u)
            t = t_out
        t_last = t
        trunk = buf("trunk", px * nf * 2)
        if c["trunk"].bn is not None:
            zt = buf("z_trunk", px * nf * 2)
            conv(c["trunk"], n, h, w, t_last, nf, zt, nf)
            yt = bn_fwd("trunk", zt, px)
            ops.append(lambda s: L.axpby_bf16(y_first, nf, 0, yt, nf, 0, 1.0, trunk, nf, 0, px, nf, s))
        else:
            conv(c["trunk"], n, h, w, t_last, nf, trunk, nf, res=y_first)
        up_in, up_z, up_y = [], [], []
        cur, hh, ww = trunk, h, w
        for i in range(nup):
            pxo = n * 4 * hh * ww
            z, y = buf(f"upz_{i}", pxo * nf * 2), buf(f"upy_{i}", pxo * nf * 2)
            conv(c[f"up{i}"], n, hh, ww, cur, nf, z, nf)                      # up=2: depth_to_space in the store
            al = c[f"up{i}"].d_alpha
            ops.append(lambda s, z=z, y=y, al=al, pxo=pxo: L.act_fwd_bf16(z, nf, 0, al, 0.0, y, nf, 0, pxo, nf, s))
            up_in.append((cur, hh, ww))
            up_z.append(z)
            up_y.append(y)
            cur, hh, ww = y, 2 * hh, 2 * ww
        H, W = hh, ww
        pxh = n * H * W
        sr = buf("out_f32", pxh * 3 * 4)
        lc = c["last"]
        # through pb.conv like every other layer: the trainer's own weight image and bias view (with batch norm the
        # model's d_packed / d_bias are the folded INFERENCE images and do not follow the updates)
        conv(lc, n, H, W, cur, nf, sr, 3, act=L.ACT_TANH, out_dtype=L.SSR_F32)

        # ------------------------------------------------------------------ loss, metric and d(loss)/d(pre-tanh)
        g_sr, dz_f32 = buf("g_sr", pxh * 3 * 4), buf("dz_last_f32", pxh * 3 * 4)
        buf("loss_ws", L.load().ssr_pixel_loss_workspace_bytes(n))
        metrics_raw, mviews, extra_off, mcount = self._metrics_layout(pb, n)
        loss_out = mviews[0]
        buf("cs_ws", L.load().ssr_channel_sum_workspace_bytes(256))
        ops.append(lambda s: L.pixel_loss(hr_f32, sr, n, H * W * 3, self.w_mse, self.w_mae, 2.0, g_sr, B["loss_ws"],
                                          loss_out, s))
        extra_out = self._emit_extra(pb, n, H, W, hr_f32, sr, g_sr, mviews[1:])
        ops.append(lambda s: L.tanh_bwd_f32(g_sr, sr, dz_f32, pxh * 3, s))

        # ------------------------------------------------------------------ backward
        dz16 = buf("dz_last_bf16", pxh * 16 * 2)          # 3 -> 16 channels (16-byte rule of the TMA tensor map)
        dzu = buf("dz_last_unrolled", pxh * 32 * 2)
        ops.append(lambda s: L.check(L.load().ssr_memset(dz16.ptr, 0, dz16.nbytes, s)))
        ops.append(lambda s: L.f32_to_bf16_slice(dz_f32, dz16, 16, 0, pxh, 3, s))
        ops.append(lambda s: L.im2col_x_f32_to_bf16(dz_f32, dzu, n, H, W, 3, 9, 32, s))
        wgrad("last", cur, nf, nf, dz16, 16, 3, n, H, W, 9, 9)
        bias_grad("last", dz16, 16, 3, pxh)
        d_hr = buf("d_hr", pxh * nf * 2)
        conv(lc, n, H, W, dzu, 32, d_hr, nf, packed=self.dgrad_packed["last"], cin=32, cout=nf, kh=9, kw=1, up=1,
             bias=False)
        d = d_hr
        for i in reversed(range(nup)):
            x_in, hi, wi = up_in[i]
            pxo, pxi = n * 4 * hi * wi, n * hi * wi
            dzh = buf(f"dzh_{i}", pxo * nf * 2)
            prelu_bwd(f"up{i}", d, up_z[i], nf, pxo, dzh)
            dzl = buf(f"dzl_{i}", pxi * 4 * nf * 2)
            ops.append(lambda s, dzh=dzh, dzl=dzl, hi=hi, wi=wi: L.space_to_depth2(dzh, dzl, n, hi, wi, nf, 2, s))
            wgrad(f"up{i}", x_in, nf, nf, dzl, 4 * nf, 4 * nf, n, hi, wi, 3, 3)
            bias_grad(f"up{i}", dzl, 4 * nf, 4 * nf, pxi)
            dn = buf(f"d_up_in_{i}", pxi * nf * 2)
            conv(c[f"up{i}"], n, hi, wi, dzl, 4 * nf, dn, nf, packed=self.dgrad_packed[f"up{i}"], cin=4 * nf, cout=nf,
                 up=1, bias=False)
            d = dn
        d_trunk = d                                        # gradient of the trunk output (also flows into the skip)
        dz_trunk = bn_bwd("trunk", d_trunk, px) if c["trunk"].bn is not None else d_trunk
        wgrad("trunk", t_last, nf, nf, dz_trunk, nf, nf, n, h, w, 3, 3)
        bias_grad("trunk", dz_trunk, nf, nf, px)
        d = buf("d_t_last", px * nf * 2)
        conv(c["trunk"], n, h, w, dz_trunk, nf, d, nf, packed=self.dgrad_packed["trunk"], cin=nf, cout=nf, bias=False)
        # (after the bucket's last dgrad launch: its weight images are re-packed right behind its Adam update)
        pb.bucket_ready(["trunk"] + [f"up{i}" for i in range(nup)] + ["last"])
        # the res blocks complete from the last to the first: one bucket per group of blocks
        groups = np.array_split(np.arange(nb), min(max(1, self.num_buckets - 1), nb))
        group_first = {int(g[0]): [int(j) for j in g] for g in groups if len(g)}
        for b in reversed(range(nb)):
            n0, n1 = f"res{b}_conv0", f"res{b}_conv1"
            dz1 = bn_bwd(n1, d, px) if c[n1].bn is not None else d
            wgrad(n1, us[b], nf, nf, dz1, nf, nf, n, h, w, 3, 3)
            bias_grad(n1, dz1, nf, nf, px)
            du = buf(f"du_{b}", px * nf * 2)
            conv(c[n1], n, h, w, dz1, nf, du, nf, packed=self.dgrad_packed[n1], cin=nf, cout=nf, bias=False)
            dz0 = buf(f"dz0_{b}", px * nf * 2)
            prelu_bwd(n0, du, z0s[b], nf, px, dz0)
            if c[n0].bn is not None:
                dz0 = bn_bwd(n0, dz0, px)
            wgrad(n0, t_in[b], nf, nf, dz0, nf, nf, n, h, w, 3, 3)
            bias_grad(n0, dz0, nf, nf, px)
            dprev = buf(f"d_t_{b}", px * nf * 2)
            conv(c[n0], n, h, w, dz0, nf, dprev, nf, packed=self.dgrad_packed[n0], cin=nf, cout=nf, res=d, bias=False)
            d = dprev
            if b in group_first and b != 0:
                pb.bucket_ready([f"res{j}_conv{k}" for j in group_first[b] for k in (0, 1)])
        d_first = buf("d_y_first", px * nf * 2)           # block chain + long skip
        ops.append(lambda s: L.axpby_bf16(d, nf, 0, d_trunk, nf, 0, 1.0, d_first, nf, 0, px, nf, s))
        dzf = buf("dz_first", px * nf * 2)
        prelu_bwd("first", d_first, z_first, nf, px, dzf)
        wgrad("first", x32, 32, 27, dzf, nf, nf, n, h, w, 9, 1)
        bias_grad("first", dzf, nf, nf, px)
        pb.bucket_ready(["first"] + [f"res{j}_conv{k}" for j in group_first[0] for k in (0, 1)])
        for el in self.extra_losses:                       # side-stream work of the loss functors joins here
            if hasattr(el, "emit_join"):
                el.emit_join(ops)
        pb.join()
        pb.finish_update()
        plan = pb.finish(n, H, W)
        self._finish_metrics(pb, plan, metrics_raw, extra_off, mcount)
        self._plans[key] = plan
        return plan



class RRDBTrainer(_TrainerBase):
    """Training iterations of a ``build_enhanced_resnet`` model (model_builder.py:42-96, 328-365) with pixel losses:
    the PSNR-oriented RRDB pre-training that precedes ESRGAN's GAN phase (examples/training: rrdb recipe).

    Every dense block keeps its own [N,H,W,192] activation buffer [x | c1..c4] for the backward pass.  The backward pass
    of a block mirrors the forward one: the gradient of the buffer is produced slice by slice, each slice by ONE
    convolution over a channel prefix of a gradient-source buffer [G | dZ_3 | dZ_2 | dZ_1 | dZ_0] with a weight image
    composed of the rotated kernels of every layer that read that slice (see ``_plan``), the LeakyReLU backward fused
    into the epilogue - five launches per block, nothing accumulated in place."""

    ARCH = "rrdb"
    NO_DGRAD = ("fea",)
    UNROLLED_DGRAD = ()

    def _slice_dgrad(self, name):
        return name.startswith("rrdb")

    def _build_slice_images(self):
        """Composed weight images of the slice-by-slice dense-block backward (see _plan): image (block, j) maps the
        prefix [G | dZ_3 .. dZ_(4-j)] of the gradient-source buffer to buffer slice 4-j (j = 4: channels [0, 64))."""
        m, ctx = self.model, self.ctx
        cfg = m.config
        nf, gc, nc = cfg["num_filters"], cfg["num_filters"] // 2, cfg["num_convs"]
        beta = float(cfg["residual_scaling_factor"])
        cw = nf + nc * gc
        self.slice_packed, self._slice_items = {}, {}
        for b in range(cfg["num_rrdb_blocks"]):
            for d in range(cfg["num_dense_blocks"]):
                pre = f"rrdb{b}_db{d}"
                k_out = self._view(self.layout[f"{pre}_out"]["k"])
                items = []
                for j in range(nc + 1):
                    cin_j, rows = nf + j * gc, (gc if j < nc else nf)
                    row0 = nf + (nc - 1 - j) * gc if j < nc else 0
                    img = L.DeviceBuffer(ctx.conv_packed_bytes(3, cin_j, rows, 1))
                    img.zero(self.stream.ptr)          # the unused half of a partial last K chunk is never read; keep it clean
                    self.slice_packed[(pre, j)] = img
                    # the G group: beta * rot180(W_out)[rows of the slice]
                    items.append(L.PackItem(k_out.ptr, img.ptr, 3, 3, cin_j, cin_j, rows, 1, 3, 0, 0, nf, row0, cw, nf, beta))
                    for q in range(j):                       # dZ_(nc-1-q): the growth convolutions that read this slice
                        mconv = nc - 1 - q
                        k_m = self._view(self.layout[f"{pre}_conv{mconv}"]["k"])
                        items.append(L.PackItem(k_m.ptr, img.ptr, 3, 3, cin_j, cin_j, rows, 1, 3, 0, nf + q * gc, gc, row0,
                                                nf + mconv * gc, gc, 1.0))
                self._slice_items[f"{pre}_out"] = items

    overlap_wgrad = True    # wgrad kernels run on a side stream next to the dgrad chain

    def _plan(self, n, h, w):
        key = (n, h, w)
        if key in self._plans:
            return self._plans[key]
        m = self.model
        c = m.convs
        cfg = m.config
        nf, gc = cfg["num_filters"], cfg["num_filters"] // 2
        nblk, ndb, nc = cfg["num_rrdb_blocks"], cfg["num_dense_blocks"], cfg["num_convs"]
        beta = float(cfg["residual_scaling_factor"])
        cw = nf + nc * gc
        sf = m.upsample_factor
        nup = int(math.log(sf, 2))
        px = n * h * w
        pb = _PlanBuilder(self)
        B, ops = pb.B, pb.ops
        buf, conv, wgrad, bias_grad = pb.buf, pb.conv, pb.wgrad, pb.bias_grad
        names = [(f"rrdb{b}_db{d}", b, d) for b in range(nblk) for d in range(ndb)]
        ND = len(names)

        def lrelu_bwd(dy, dcs, doff, y, ycs, yoff, dz, zcs, pixels, ch):
            ops.append(lambda s: L.act_bwd_bf16(dy, dcs, doff, y, ycs, yoff, None, 0.2, dz, zcs, 0, pixels, ch, s))

        # ------------------------------------------------------------------ forward
        in_f32 = buf("in_f32", px * 3 * 4)
        hr_f32 = buf("hr_f32", px * sf * sf * 3 * 4)
        x16 = buf("x16", px * 16 * 2)
        ops.append(lambda s: L.f32_to_bf16_pad(in_f32, x16, px, 3, 16, s))
        D = [buf(f"dense_{i}", px * cw * 2) for i in range(ND + 1)]
        conv(c["fea"], n, h, w, x16, 16, D[0], cw)                       # fea lives in channels [0,64) of block 0's buffer
        for i, (pre, _, _) in enumerate(names):
            for k in range(nc):
                conv(c[f"{pre}_conv{k}"], n, h, w, D[i], cw, D[i], cw, ocoff=nf + k * gc, act=L.ACT_LRELU, act_alpha=0.2)
            conv(c[f"{pre}_out"], n, h, w, D[i], cw, D[i + 1], cw, res=D[i], res_cs=cw, res_beta=beta)
        t_in = buf("trunk_in", px * nf * 2)
        ops.append(lambda s: L.axpby_bf16(D[0], cw, 0, D[ND], cw, 0, beta, t_in, nf, 0, px, nf, s))   # fea + beta * r
        u0 = buf("u0", px * nf * 2)
        conv(c["trunk"], n, h, w, t_in, nf, u0, nf, res=D[0], res_cs=cw)
        up_in, up_y = [], []
        cur, hh, ww = u0, h, w
        for i in range(nup):
            y = buf(f"upy_{i}", n * 4 * hh * ww * nf * 2)
            conv(c[f"up{i}"], n, hh, ww, cur, nf, y, nf, act=L.ACT_LRELU, act_alpha=0.2)   # d2s in the store, then lrelu
            up_in.append((cur, hh, ww))
            up_y.append(y)
            cur, hh, ww = y, 2 * hh, 2 * ww
        H, W = hh, ww
        pxh = n * H * W
        hr_y = buf("hr_y", pxh * nf * 2)
        conv(c["hr"], n, H, W, cur, nf, hr_y, nf, act=L.ACT_LRELU, act_alpha=0.2)
        sr = buf("out_f32", pxh * 3 * 4)
        lc = c["last"]
        conv(lc, n, H, W, hr_y, nf, sr, 3, act=L.ACT_TANH, out_dtype=L.SSR_F32)

        # ------------------------------------------------------------------ loss
        g_sr, dz_f32 = buf("g_sr", pxh * 3 * 4), buf("dz_last_f32", pxh * 3 * 4)
        buf("loss_ws", L.load().ssr_pixel_loss_workspace_bytes(n))
        metrics_raw, mviews, extra_off, mcount = self._metrics_layout(pb, n)
        loss_out = mviews[0]
        buf("cs_ws", L.load().ssr_channel_sum_workspace_bytes(256))
        ops.append(lambda s: L.pixel_loss(hr_f32, sr, n, H * W * 3, self.w_mse, self.w_mae, 2.0, g_sr, B["loss_ws"],
                                          loss_out, s))
        extra_out = self._emit_extra(pb, n, H, W, hr_f32, sr, g_sr, mviews[1:])
        ops.append(lambda s: L.tanh_bwd_f32(g_sr, sr, dz_f32, pxh * 3, s))

        # ------------------------------------------------------------------ backward: HR tail
        dz16 = buf("dz_last_bf16", pxh * 16 * 2)
        ops.append(lambda s: L.check(L.load().ssr_memset(dz16.ptr, 0, dz16.nbytes, s)))
        ops.append(lambda s: L.f32_to_bf16_slice(dz_f32, dz16, 16, 0, pxh, 3, s))
        wgrad("last", hr_y, nf, nf, dz16, 16, 3, n, H, W, 3, 3)
        bias_grad("last", dz16, 16, 3, pxh)
        d_hry = buf("d_hr_y", pxh * nf * 2)
        conv(lc, n, H, W, dz16, 16, d_hry, nf, packed=self.dgrad_packed["last"], cin=16, cout=nf, bias=False)
        dz_hr = buf("dz_hr", pxh * nf * 2)
        lrelu_bwd(d_hry, nf, 0, hr_y, nf, 0, dz_hr, nf, pxh, nf)
        wgrad("hr", cur, nf, nf, dz_hr, nf, nf, n, H, W, 3, 3)
        bias_grad("hr", dz_hr, nf, nf, pxh)
        d = buf("d_up_last", pxh * nf * 2)
        conv(c["hr"], n, H, W, dz_hr, nf, d, nf, packed=self.dgrad_packed["hr"], cin=nf, cout=nf, bias=False)
        for i in reversed(range(nup)):
            x_in, hi, wi = up_in[i]
            pxo, pxi = n * 4 * hi * wi, n * hi * wi
            dzh = buf(f"dzh_{i}", pxo * nf * 2)
            lrelu_bwd(d, nf, 0, up_y[i], nf, 0, dzh, nf, pxo, nf)
            dzl = buf(f"dzl_{i}", pxi * 4 * nf * 2)
            ops.append(lambda s, dzh=dzh, dzl=dzl, hi=hi, wi=wi: L.space_to_depth2(dzh, dzl, n, hi, wi, nf, 2, s))
            wgrad(f"up{i}", x_in, nf, nf, dzl, 4 * nf, 4 * nf, n, hi, wi, 3, 3)
            bias_grad(f"up{i}", dzl, 4 * nf, 4 * nf, pxi)
            dn = buf(f"d_up_in_{i}", pxi * nf * 2)
            conv(c[f"up{i}"], n, hi, wi, dzl, 4 * nf, dn, nf, packed=self.dgrad_packed[f"up{i}"], cin=4 * nf, cout=nf,
                 up=1, bias=False)
            d = dn
        # ------------------------------------------------------------------ backward: trunk
        d_u0 = d                                           # u0 = fea + conv(trunk, trunk_in)
        wgrad("trunk", t_in, nf, nf, d_u0, nf, nf, n, h, w, 3, 3)
        bias_grad("trunk", d_u0, nf, nf, px)
        d_ti = buf("d_trunk_in", px * nf * 2)
        conv(c["trunk"], n, h, w, d_u0, nf, d_ti, nf, packed=self.dgrad_packed["trunk"], cin=nf, cout=nf, bias=False)
        # (after the bucket's last dgrad launch: its weight images are re-packed right behind its Adam update)
        pb.bucket_ready(["trunk"] + [f"up{i}" for i in range(nup)] + ["hr", "last"])
        g_fea = buf("g_fea", px * nf * 2)                  # fea feeds u0 (identity) and trunk_in (identity)
        ops.append(lambda s: L.axpby_bf16(d_u0, nf, 0, d_ti, nf, 0, 1.0, g_fea, nf, 0, px, nf, s))
        # ------------------------------------------------------------------ backward: dense blocks, slice by slice
        # The mirror image of the forward pass.  Forward, conv k reads the channel PREFIX [x | c1..ck] of the block's
        # buffer and writes the next 32-channel slice; backward, the gradient of the buffer is produced one slice at a
        # time, last slice first, by ONE convolution over the prefix of a gradient-source buffer
        #     SRC = [ G (64: dL/dx_next) | dZ_3 | dZ_2 | dZ_1 | dZ_0 ]      (32 channels each)
        # with a weight image composed of the rotated kernels of every later layer that reads that slice (pack mode 3):
        #     j = 0..3:  dZ_(3-j) = lrelu'(c_(4-j)) * conv(SRC[:, :64+32j])        -> SRC slice j+1   (SSR_ACT_LRELU_MASK)
        #     j = 4   :  dL/dx    = G + conv(SRC[:, :192])                            -> next block's SRC[:, :64]
        # Five launches per block (was ten: an N = 192 dgrad, an axpby, four LeakyReLU-backward kernels and four
        # read-modify-write dgrads over up to 160 channels), every output written once.  beta rides in the G group's
        # weights.  SRC rotates over three buffers so that a block only overwrites what the side stream's weight-gradient
        # launch of the block two steps later has finished reading (block_done).
        SRC = [buf("gsrc_a", px * cw * 2), buf("gsrc_b", px * cw * 2), buf("gsrc_c", px * cw * 2)]
        zero = buf("zero", px * nf * 2)
        ops.append(lambda s: L.check(L.load().ssr_memset(zero.ptr, 0, zero.nbytes, s)))
        first_src = SRC[(ND - 1) % 3]                      # trunk_in = fea + beta * r  ->  G of the last block
        ops.append(lambda s: L.axpby_bf16(zero, nf, 0, d_ti, nf, 0, beta, first_src, cw, 0, px, nf, s))
        block_done = {}
        # dense blocks complete from the last to the first: one optimizer bucket per group of blocks
        groups = np.array_split(np.arange(ND), min(max(1, self.num_buckets - 1), ND))
        group_first = {int(g[0]): [int(j) for j in g] for g in groups if len(g)}
        block_convs = lambda j: [f"{names[j][0]}_conv{k}" for k in range(nc)] + [f"{names[j][0]}_out"]
        for i in reversed(range(ND)):
            pre = names[i][0]
            src, nxt = SRC[i % 3], SRC[(i - 1) % 3]
            pb.main_wait(block_done.get(i + 2))
            for j in range(nc + 1):
                cin_j = nf + j * gc
                pk = self.slice_packed[(pre, j)]
                if j < nc:
                    yoff = nf + (nc - 1 - j) * gc          # forward activation c_(nc-j): channels of D[i] the mask reads
                    conv(c[f"{pre}_conv0"], n, h, w, src, cw, src, cw, ocoff=nf + j * gc, packed=pk, cin=cin_j, cout=gc,
                         res=D[i], res_cs=cw, res_coff=yoff, bias=False, act=L.ACT_LRELU_MASK, act_alpha=0.2)
                else:
                    conv(c[f"{pre}_out"], n, h, w, src, cw, nxt, cw, packed=pk, cin=cw, cout=nf, res=src, res_cs=cw,
                         res_beta=1.0, bias=False)
            # weight gradients of the block's five convolutions: one batched launch on the side stream
            wgrad(f"{pre}_out", D[i], cw, cw, src, cw, nf, n, h, w, 3, 3, scale=beta, zoff=0)
            bias_grad(f"{pre}_out", src, cw, nf, px, scale=beta, zoff=0)
            for k in reversed(range(nc)):
                zoff = nf + (nc - 1 - k) * gc              # dZ_k sits behind G and the dZ of the later convolutions
                wgrad(f"{pre}_conv{k}", D[i], cw, nf + k * gc, src, cw, gc, n, h, w, 3, 3, zoff=zoff)
                bias_grad(f"{pre}_conv{k}", src, cw, gc, px, zoff=zoff)
            block_done[i] = pb.side_mark()
            if i in group_first and i != 0:
                pb.bucket_ready([nm for j in group_first[i] for nm in block_convs(j)])
        G, Gcs = SRC[(0 - 1) % 3], cw                      # dL/dx of block 0 = gradient of fea through the trunk
        g_fea_t = buf("g_fea_total", px * nf * 2)
        ops.append(lambda s, G=G, Gcs=Gcs: L.axpby_bf16(g_fea, nf, 0, G, Gcs, 0, 1.0, g_fea_t, nf, 0, px, nf, s))
        wgrad("fea", x16, 16, 3, g_fea_t, nf, nf, n, h, w, 3, 3)
        bias_grad("fea", g_fea_t, nf, nf, px)
        pb.bucket_ready(["fea"] + [nm for j in group_first[0] for nm in block_convs(j)])
        for el in self.extra_losses:                       # side-stream work of the loss functors joins here
            if hasattr(el, "emit_join"):
                el.emit_join(ops)
        pb.join()                                          # every weight gradient is in before the fallback all-reduce
        pb.finish_update()
        plan = pb.finish(n, H, W)
        self._finish_metrics(pb, plan, metrics_raw, extra_off, mcount)
        self._plans[key] = plan
        return plan
