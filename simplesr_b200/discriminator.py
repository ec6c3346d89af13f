"""ESRGAN discriminator and the relativistic-average GAN branch of ``SRModel.train_step`` on the B200.

Mirrors ``model_builder.build_discriminator`` (simple_sr/utils/models/model_builder.py:137-198),
``Discriminator.critic_train_batch`` (simple_sr/models/discriminator.py:147-172: the critic is called on the SR batch and
on the HR batch, both with training=True, i.e. separate BatchNormalization batch statistics per call),
``RaAdversarialLoss`` (ra_adversarial_loss.py:59-70) for the generator and ``RaDiscriminatorLoss``
(ra_discriminator_loss.py:55-66) for the discriminator, and the two ``apply_gradients`` of sr_model.py:436-451.

``RaGANLoss`` is a loss functor for the generator trainers (``extra_losses=[...]``): inside the training graph it runs
D(sr) and D(hr), evaluates both relativistic losses, adds ``loss_weight * d(L_G)/d(sr)`` to the generator's image
gradient (dgrad through D, BatchNormalization backward included), and accumulates the discriminator's own gradients from
both critic passes, and applies Keras-Adam to D - all inside the trainer's step graph (sr_model.py:444-451).

Convolutions run on the tcgen05 conv kernel; the stride-2 layers as the stride-1 convolution sampled at odd positions
(``ssr_subsample2`` / ``ssr_zero_insert2``: exact, 4x the MACs of those four layers - a true strided implicit GEMM is
future work); BatchNorm, the two Dense layers and the losses are bandwidth-bound CUDA-core kernels (disc_kernels.cu).
Label smoothing (discriminator.py:240-254) is an option of ``RaGANLoss`` (per-sample random target labels).
"""
import math

import numpy as np

from . import _lib as L
from .model_builder import Variable, _he_normal_scaled, get_context

DISC_CONVS = [("d_conv0", 3, 64, 1, False), ("d_conv1", 64, 64, 2, True), ("d_conv2", 64, 128, 1, True),
              ("d_conv3", 128, 128, 2, True), ("d_conv4", 128, 256, 1, True), ("d_conv5", 256, 256, 2, True),
              ("d_conv6", 256, 512, 1, True), ("d_conv7", 512, 512, 2, True)]
BN_EPS = 1e-3


def smoothed_labels(rng, sr_shape, hr_shape, label_smoothing, smoothing_offset):
    """``Discriminator._get_labels`` (discriminator.py:236-254): float64 target labels ``(sr_labels, hr_labels)``.
    Without smoothing 0 and 1 - offset (the constructor zeroes the offset then, :68-70); with it
    ``sr = U(0,1) * offset`` and ``hr = 1 - offset + U(0, 0.5)``."""
    noise_sr = noise_hr = 0.0
    if label_smoothing:
        noise_hr = rng.uniform(0.0, 0.5, size=hr_shape)
        noise_sr = rng.uniform(0.0, 1.0, size=sr_shape) * smoothing_offset
    sr_labels = np.zeros(sr_shape, np.float64) + noise_sr
    hr_labels = np.ones(hr_shape, np.float64) - smoothing_offset + noise_hr
    return sr_labels, hr_labels


class DiscriminatorModel:
    """Variables in Keras creation order: per conv [kernel, bias, (gamma, beta)], then the two Dense layers."""

    def __init__(self, input_dims=(128, 128), num_filters=64, alpha=0.2, momentum=0.8, seed=None, device=0,
                 relativistic=True):
        self.relativistic = bool(relativistic)
        if num_filters != 64:
            raise ValueError("only num_filters=64 is supported by the sm_100a discriminator")
        if input_dims[0] is None or input_dims[0] % 16 or input_dims[1] % 16:
            raise ValueError("input_dims must be multiples of 16 (Flatten fixes the input size, model_builder.py:188-189)")
        rng = np.random.default_rng(seed)
        self.input_dims, self.alpha, self.momentum = tuple(input_dims), float(alpha), float(momentum)
        self.ctx = get_context(device)
        self.vars = {}            # name -> list of Variables
        for name, cin, cout, _, bn in DISC_CONVS:
            self.vars[name] = [Variable(f"{name}/kernel:0", _he_normal_scaled(rng, (3, 3, cin, cout))),
                               Variable(f"{name}/bias:0", np.zeros(cout, np.float32))]
            if bn:
                self.vars[name + "_bn"] = [Variable(f"{name}_bn/gamma:0", np.ones(cout, np.float32)),
                                           Variable(f"{name}_bn/beta:0", np.zeros(cout, np.float32))]
        self.flat_features = (input_dims[0] // 16) * (input_dims[1] // 16) * 512
        self.vars["d_dense0"] = [Variable("d_dense0/kernel:0", _he_normal_scaled(rng, (self.flat_features, 1024))),
                                 Variable("d_dense0/bias:0", np.zeros(1024, np.float32))]
        self.vars["d_dense1"] = [Variable("d_dense1/kernel:0", _he_normal_scaled(rng, (1024, 1))),
                                 Variable("d_dense1/bias:0", np.zeros(1, np.float32))]

    @property
    def trainable_variables(self):
        out = []
        for vs in self.vars.values():
            out.extend(vs)
        return out

    def set_params(self, params):
        """params: {name: [array, array]} as produced by the oracle's init_discriminator_params."""
        for name, arrs in params.items():
            for v, a in zip(self.vars[name], arrs):
                v.assign(a)

    def count_params(self):
        return int(sum(v.numpy().size for v in self.trainable_variables))


def build_discriminator(input_dims=(None, None), num_filters=64, alpha=0.2, kernel_size=3, momentum=0.8,
                        relativistic=False, initializer=None, seed=None, device=0):
    """model_builder.build_discriminator (:137-139).  ``relativistic=False`` is the standard critic whose last layer is
    ``Dense(1, activation="sigmoid")`` (:194-196): same variables; the sigmoid is evaluated with the losses
    (``ssr_gan_losses_ex`` takes the logits), so both forms share every kernel up to the critic."""
    if kernel_size != 3:
        raise ValueError("only kernel_size=3 is supported")
    return DiscriminatorModel(input_dims=input_dims, num_filters=num_filters, alpha=alpha, momentum=momentum, seed=seed,
                              device=device, relativistic=relativistic)


class RaGANLoss:
    """Relativistic-average adversarial term of the ESRGAN step + the discriminator's own update (``relativistic=False``:
    the standard-GAN pair of the SRGAN recipe, ``GANLoss``).

    ``label_smoothing`` / ``smoothing_offset``: the random target labels of ``Discriminator._get_labels``
    (discriminator.py:240-254): SR labels U(0,1)*offset, HR labels 1 - offset + U(0,0.5); drawn on the host every step
    (numpy generator ``seed``; TensorFlow's stream cannot be reproduced) and uploaded before the step graph runs.
    Data-parallel (the trainer's ``comm``): BatchNormalization uses the statistics of the global batch in both critic
    passes (sync-BN), the relativistic means run over the global batch, and the discriminator's gradients are exchanged
    by the same reduce-scatter + Adam + all-gather kernel as the generator's."""

    def __init__(self, discriminator, loss_weight=5e-3, learning_rate=1e-4, beta_1=0.9, beta_2=0.999, epsilon=1e-7,
                 allreduce=None, label_smoothing=False, smoothing_offset=0.3, seed=None, relativistic=None):
        # relativistic=None: as the discriminator was built.  False = the standard GAN pair AdversarialLoss
        # (adversarial_loss.py:58) / DiscriminatorLoss (discriminator_loss.py:56-59) on sigmoid critics.
        self.relativistic = bool(getattr(discriminator, "relativistic", True) if relativistic is None else relativistic)
        self.name = "ra_adversarial_loss" if self.relativistic else "adversarial_loss"
        self.metric_names = (["ra_adversarial_loss", "ra_discriminator_loss"] if self.relativistic
                             else ["adversarial_loss", "discriminator_loss"])
        self.D = discriminator
        self.ctx = discriminator.ctx
        self.loss_weight, self.feature_scale = float(loss_weight), 1.0
        self.weighted = self.loss_weight != 1.0
        self._opt_args = (learning_rate, float(beta_1), float(beta_2), float(epsilon))
        self.allreduce = allreduce
        self.label_smoothing = bool(label_smoothing)
        self.smoothing_offset = float(smoothing_offset) if label_smoothing else 0.0     # discriminator.py:68-70
        self._rng = np.random.default_rng(seed)
        self.iterations = 0
        self.comm = None
        self._out = None
        self._labels = None
        self.overlap_update = True   # weight-gradient passes of the critic on a side stream (see emit)
        self.opt = None

    @property
    def metric_scale(self):
        return self.loss_weight

    def metric_floats(self, n):
        return 2

    def attach(self, trainer):
        """Called by the generator trainer: share its peer fabric, then build the flat buffers (in its heap)."""
        self.comm = getattr(trainer, "comm", None)
        if self.comm is not None and self.allreduce is not None:
            raise ValueError("RaGANLoss: pass either the trainer's comm or an allreduce hook, not both")
        if self.opt is None:
            self._build_flat()

    # ---- flat parameter buffer of the discriminator ------------------------------------------------------------------
    def _build_flat(self):
        from .training import FlatAdam
        self.layout, off, host = {}, 0, []
        for name, vs in self.D.vars.items():
            ent = []
            for v in vs:
                a = v.numpy().ravel()
                ent.append((off, a.size))
                host.append(a)
                off += a.size
            self.layout[name] = ent
        self.count = off
        flat = np.concatenate(host).astype(np.float32)
        lr, b1, b2, eps = self._opt_args
        self.opt = FlatAdam(flat, lr, b1, b2, eps, None, comm=self.comm)
        self.d_param, self.d_grad, self.d_m, self.d_v = self.opt.d_param, self.opt.d_grad, self.opt.d_m, self.opt.d_v
        self.packed, self.dpacked = {}, {}
        items = []
        for name, cin, cout, _, _ in DISC_CONVS:
            cin_p, cout_p = -(-cin // 16) * 16, -(-cout // 16) * 16
            self.packed[name] = L.DeviceBuffer(self.ctx.conv_packed_bytes(3, cin_p, cout, 1))
            self.dpacked[name] = L.DeviceBuffer(self.ctx.conv_packed_bytes(3, cout_p, cin, 1))
            k = self._p(name, 0)
            items.append(L.PackItem(k.ptr, self.packed[name].ptr, 3, 3, cin, cin_p, cout, 1, 0, 0))
            items.append(L.PackItem(k.ptr, self.dpacked[name].ptr, 3, 3, cin, cin_p, cout, 1, 1, 0))
        self._pack_count = len(items)
        self._pack_table = self.ctx.pack_batch_prepare(items, None)
        self._repack(None)
        L.stream_sync(None)
        trainer = self
        for name, vs in self.D.vars.items():
            for v, (o, sz) in zip(vs, self.layout[name]):
                v._pull = (lambda v=v, o=o, sz=sz: setattr(
                    v, "_value", trainer.d_param.download((sz,), np.float32, None, offset=o * 4).reshape(v._value.shape)))
                v._push = (lambda value, o=o: (trainer.d_param.upload(np.ascontiguousarray(value, np.float32).ravel(), None,
                                                                     offset=o * 4), trainer._repack(None),
                                               L.stream_sync(None)))

    def _p(self, name, i, buf=None):
        off, size = self.layout[name][i]
        return L.DeviceView(buf or self.d_param, off * 4, size * 4)

    def _repack(self, s):
        """All 16 weight images (forward + dgrad of the 8 convs) from the fp32 masters in ONE launch."""
        self.ctx.pack_batch(self._pack_table, self._pack_count, s)

    def gradients(self):
        flat = self.d_grad.download((self.count,), np.float32)
        return {name: [flat[o:o + sz].reshape(v.shape) for v, (o, sz) in zip(self.D.vars[name], self.layout[name])]
                for name in self.D.vars}

    # ---- launch list -------------------------------------------------------------------------------------------------
    def emit(self, ops, B, prefix, n, H, W, hr_f32, sr_f32, g_sr, accumulate=True, out=None):
        if (H, W) != self.D.input_dims:
            raise ValueError(f"discriminator was built for {self.D.input_dims} inputs, got {(H, W)}")
        if self.opt is None:
            self._build_flat()           # stand-alone use (no trainer attached)
        ctx, alpha = self.ctx, self.D.alpha
        comm = self.comm
        bn_site = (lambda c_: comm.bn_site(c_)) if comm is not None else (lambda c_: None)
        wg_bytes = 0
        # The two backward passes that only produce the discriminator's own weight gradients do not feed the generator's
        # backward pass: they run on a second stream next to it (fork after the losses, join in emit_join, which the
        # trainer appends after the generator's backward).  `ops` redirects launches to that stream while it is set.
        ops = L.OpsView(ops)
        side = self._side_stream() if self.overlap_update else None

        def buf(name, nbytes):
            B[prefix + name] = L.DeviceBuffer(nbytes)
            return B[prefix + name]

        def conv(x, xcs, cin, out, cout, packed, bias, h, w, act=L.ACT_NONE, out_dtype=L.SSR_BF16, ocs=None):
            d = L.ConvDesc(n=n, h=h, w=w, cin=cin, in_cstride=xcs, cout=cout, ksize=3, ksize_w=3, act=act,
                           act_alpha=alpha, res_beta=0.0, up=1, out_dtype=out_dtype, out_cstride=(ocs or cout),
                           out_coff=0, res_dtype=L.SSR_NONE, res_cstride=0, res_coff=0, out2_cstride=0, out2_coff=0)
            ops.append(lambda s: ctx.conv2d_fwd(d, x, packed, bias, out, stream=s))

        bn_ws = buf("bn_ws", L.load().ssr_bn_workspace_bytes(512))
        cs_ws = buf("cs_ws", L.load().ssr_channel_sum_workspace_bytes(512))
        bn_sums = buf("bn_sums", 2 * 512 * 4)
        # scratch of the passes on the side stream (they run concurrently with the "gsr" pass on the main stream)
        bn_ws2 = buf("bn_ws2", L.load().ssr_bn_workspace_bytes(512))
        cs_ws2 = buf("cs_ws2", L.load().ssr_channel_sum_workspace_bytes(512))
        bn_sums2 = buf("bn_sums2", 2 * 512 * 4)
        dn_ws = buf("dense_ws", L.load().ssr_dense_workspace_bytes(n, 1024))
        dn_ws2 = buf("dense_ws2", L.load().ssr_dense_workspace_bytes(n, 1024))
        F = self.D.flat_features

        # ------------------------------------------------------------------ forward of one critic pass
        def forward(tag, img, bn_ws=bn_ws, dn_ws=dn_ws):
            c = {}
            x16 = buf(f"{tag}_x16", n * H * W * 16 * 2)
            ops.append(lambda s: L.f32_to_bf16_pad(img, x16, n * H * W, 3, 16, s))
            t, tcs, h, w = x16, 16, H, W
            for name, cin, cout, stride, bn in DISC_CONVS:
                cin_p = -(-cin // 16) * 16
                rec = dict(x=t, xcs=tcs, h=h, w=w)
                if not bn:
                    y = buf(f"{tag}_{name}_y", n * h * w * cout * 2)
                    conv(t, tcs, cin_p, y, cout, self.packed[name], self._p(name, 1), h, w, act=L.ACT_LRELU)
                    rec.update(y=y, oh=h, ow=w)
                else:
                    zf = buf(f"{tag}_{name}_zf", n * h * w * cout * 2)
                    conv(t, tcs, cin_p, zf, cout, self.packed[name], self._p(name, 1), h, w)
                    oh, ow = h // stride, w // stride
                    if stride == 2:
                        z = buf(f"{tag}_{name}_z", n * oh * ow * cout * 2)
                        ops.append(lambda s, zf=zf, z=z, oh=oh, ow=ow, cout=cout: L.subsample2(zf, z, n, oh, ow, cout, 2, s))
                    else:
                        z = zf
                    mean, istd = buf(f"{tag}_{name}_mean", cout * 4), buf(f"{tag}_{name}_istd", cout * 4)
                    y = buf(f"{tag}_{name}_y", n * oh * ow * cout * 2)
                    g, b = self._p(name + "_bn", 0), self._p(name + "_bn", 1)
                    px = n * oh * ow
                    site = bn_site(cout)      # sync-BN: batch statistics of the global batch (separately per critic call)
                    ops.append(lambda s, z=z, px=px, cout=cout, mean=mean, istd=istd, site=site:
                               L.bn_stats_bf16(z, px, cout, BN_EPS, self.D.momentum, bn_ws, mean, istd, None, None, s,
                                               site=site))
                    ops.append(lambda s, z=z, px=px, cout=cout, mean=mean, istd=istd, g=g, b=b, y=y:
                               L.bn_lrelu_fwd_bf16(z, mean, istd, g, b, alpha, y, px, cout, s))
                    rec.update(z=z, y=y, mean=mean, istd=istd, oh=oh, ow=ow)
                    h, w = oh, ow
                c[name] = rec
                t, tcs = rec["y"], cout
            flat = buf(f"{tag}_flat", n * F * 4)
            ops.append(lambda s, t=t: L.bf16_to_f32(t, 512, 0, flat, n * F // 512, 512, s))
            hpre, act = buf(f"{tag}_h", n * 1024 * 4), buf(f"{tag}_a", n * 1024 * 4)
            ops.append(lambda s: L.dense_fwd_f32(flat, self._p("d_dense0", 0), self._p("d_dense0", 1), n, F, 1024, True,
                                                 alpha, dn_ws, hpre, act, s))
            critic = buf(f"{tag}_critic", n * 4)
            ops.append(lambda s: L.dense_fwd_f32(act, self._p("d_dense1", 0), self._p("d_dense1", 1), n, 1024, 1, False,
                                                 alpha, dn_ws, None, critic, s))
            c.update(flat=flat, h=hpre, a=act, critic=critic)
            return c

        # ------------------------------------------------------------------ backward of one critic pass
        def backward(tag, c, dcritic, want_w, acc_w, g_img=None, g_scale=1.0, scratch=None):
            nonlocal wg_bytes
            G = self.d_grad
            bn_ws_, bn_sums_, cs_ws_ = scratch or (bn_ws, bn_sums, cs_ws)
            da, dh = buf(f"{tag}_da", n * 1024 * 4), buf(f"{tag}_dh", n * 1024 * 4)
            dflat = buf(f"{tag}_dflat", n * F * 4)
            ops.append(lambda s: L.dense_bwd_f32(c["a"], self._p("d_dense1", 0), dcritic, n, 1024, 1, da,
                                                 self._p("d_dense1", 0, G) if want_w else None,
                                                 self._p("d_dense1", 1, G) if want_w else None, acc_w, s))
            ops.append(lambda s: L.lrelu_bwd_f32(da, c["h"], alpha, dh, n * 1024, s))
            ops.append(lambda s: L.dense_bwd_f32(c["flat"], self._p("d_dense0", 0), dh, n, F, 1024, dflat,
                                                 self._p("d_dense0", 0, G) if want_w else None,
                                                 self._p("d_dense0", 1, G) if want_w else None, acc_w, s))
            d = buf(f"{tag}_d_top", n * F * 2)
            ops.append(lambda s, d=d: L.f32_to_bf16_slice(dflat, d, 512, 0, n * F // 512, 512, s))
            for name, cin, cout, stride, bn in reversed(DISC_CONVS):
                r = c[name]
                h, w, oh, ow = r["h"], r["w"], r["oh"], r["ow"]
                pxo = n * oh * ow
                dz = buf(f"{tag}_{name}_dz", pxo * cout * 2)
                if bn:
                    dg = self._p(name + "_bn", 0, G) if want_w else None
                    db_ = self._p(name + "_bn", 1, G) if want_w else None
                    site = bn_site(cout)
                    ops.append(lambda s, r=r, d=d, dz=dz, pxo=pxo, cout=cout, dg=dg, db_=db_, name=name, site=site:
                               L.bn_lrelu_bwd_bf16(r["z"], d, r["y"], r["mean"], r["istd"], self._p(name + "_bn", 0), alpha,
                                                   pxo, cout, bn_ws_, bn_sums_, dg, db_, acc_w, dz, s, site=site))
                else:
                    ops.append(lambda s, r=r, d=d, dz=dz, pxo=pxo, cout=cout:
                               L.act_bwd_bf16(d, cout, 0, r["y"], cout, 0, None, alpha, dz, cout, 0, pxo, cout, s))
                if stride == 2:
                    dzf = buf(f"{tag}_{name}_dzf", n * h * w * cout * 2)
                    ops.append(lambda s, dz=dz, dzf=dzf, oh=oh, ow=ow, cout=cout: L.zero_insert2(dz, dzf, n, oh, ow, cout, 2, s))
                else:
                    dzf = dz
                if want_w:
                    wg_bytes = max(wg_bytes, ctx.conv_wgrad_workspace_bytes(h, w, cin, cout, 3, 3))
                    dw, dbias = self._p(name, 0, G), self._p(name, 1, G)
                    ops.append(lambda s, r=r, dzf=dzf, cin=cin, cout=cout, h=h, w=w, dw=dw:
                               ctx.conv2d_wgrad(r["x"], r["xcs"], 0, cin, dzf, cout, 0, cout, n, h, w, 3, 3,
                                                B[prefix + "wg_ws"], dw, accumulate=acc_w, stream=s))
                    ops.append(lambda s, dz=dz, cout=cout, pxo=pxo, dbias=dbias:
                               L.channel_sum_bf16(dz, cout, 0, None, 0, 0, pxo, cout, 1.0, acc_w, cs_ws_, dbias, s))
                if name == "d_conv0":
                    if g_img is not None:
                        dimg = buf(f"{tag}_dimg", n * h * w * 3 * 4)
                        conv(dzf, cout, cout, dimg, 3, self.dpacked[name], None, h, w, out_dtype=L.SSR_F32, ocs=3)
                        ops.append(lambda s, dimg=dimg: L.axpy_f32(dimg, g_img, g_scale, n * H * W * 3, s))
                else:
                    dn = buf(f"{tag}_{name}_dx", n * h * w * cin * 2)
                    conv(dzf, cout, cout, dn, cin, self.dpacked[name], None, h, w)
                    d = dn

        # D(hr) does not depend on D(sr): second stream, own scratch, joined before the losses
        if side is not None:
            f0, f1 = L.Event(), L.Event()
            self._fwd_events = [f0, f1]
            ops.append(lambda s: (f0.record(s), side.wait_event(f0)))
            ops.redirect = side
            c_hr = forward("hr", hr_f32, bn_ws=bn_ws2, dn_ws=dn_ws2)
            ops.redirect = None
            ops.append(lambda s: f1.record(side.ptr))
            c_sr = forward("sr", sr_f32)
            ops.append(lambda s: L.stream_wait_event(s, f1))
        else:
            c_sr = forward("sr", sr_f32)
            c_hr = forward("hr", hr_f32)
        if out is None:
            out = buf("out", 2 * 4)
        g_dsr, d_dsr, d_dhr = buf("g_dsr", n * 4), buf("d_dsr", n * 4), buf("d_dhr", n * 4)
        # target labels: [hr labels | sr labels], refreshed by pre_step when label smoothing is on
        labels = buf("labels", 2 * n * 4)
        labels.upload(np.concatenate([np.ones(n, np.float32), np.zeros(n, np.float32)]))
        self._labels, self._labels_n = labels, n
        lab_hr, lab_sr = L.DeviceView(labels, 0, n * 4), L.DeviceView(labels, n * 4, n * 4)
        rsite = comm.ragan_site(n) if comm is not None else None
        ops.append(lambda s: L.ragan_losses_ex(c_hr["critic"], c_sr["critic"], n, 1.0, 0.0, lab_hr, lab_sr, out, g_dsr,
                                               d_dsr, d_dhr, s, site=rsite, relativistic=self.relativistic))
        # generator: d(loss_weight * L_G)/d(sr) through D(sr) (L_G's dependence on D(hr) does not reach the generator)
        backward("gsr", c_sr, g_dsr, want_w=False, acc_w=False, g_img=g_sr, g_scale=self.loss_weight)
        # discriminator: weight gradients through both critic passes
        if side is not None:
            fork, self._side_done = L.Event(), L.Event()
            self._events = [fork, self._side_done]
            # fork after the losses AND after the "gsr" launches were queued: the side passes read the same activations
            ops.append(lambda s: (fork.record(s), side.wait_event(fork)))
            ops.redirect = side
        sc = (bn_ws2, bn_sums2, cs_ws2) if side is not None else None
        backward("dsr", c_sr, d_dsr, want_w=True, acc_w=False, scratch=sc)
        backward("dhr", c_hr, d_dhr, want_w=True, acc_w=True, scratch=sc)
        if self.allreduce is None:
            # the discriminator's apply_gradients (sr_model.py:444-451) inside the step graph, right behind its last
            # weight-gradient kernel: clock, then Adam (data-parallel: with the gradient exchange over peer memory).
            # The re-pack of the weight images waits for the generator's backward pass (emit_join): it still reads them.
            opt = self.opt
            opt.reserve("all")
            ops.append(lambda s: opt.prepare(s))
            ops.append(lambda s: opt.update(0, opt.padded, s, key="all"))
            # new weight images right away, on the same stream: the generator-side pass through D ("gsr") finished before
            # this stream forked, nothing reads the old images any more in this step
            ops.append(lambda s: self._repack(s))
        if side is not None:
            ops.redirect = None
            done = self._side_done
            ops.append(lambda s: done.record(side.ptr))
        buf("wg_ws", max(wg_bytes, 16))
        self._out = out
        return out

    def emit_join(self, ops):
        """Appended by the trainer after the generator's backward pass: the main stream waits for the discriminator's
        weight-gradient passes (and its Adam update) on the side stream, inside the step graph."""
        if self.overlap_update and getattr(self, "_side_done", None) is not None:
            done = self._side_done
            ops.append(lambda s: L.stream_wait_event(s, done))

    def emit_update(self, ops):
        """Appended by the trainer after the generator's own update.  In the step graph the discriminator's Adam and re-pack
        already ran on its side stream (``emit``); with the fallback all-reduce hook the whole update runs here, eagerly."""
        if self.allreduce is not None:
            opt, hook, this = self.opt, self.allreduce, self
            ops.append(lambda s: hook(this.d_grad, this.count, s))
            ops.append(lambda s: opt.prepare(s))
            ops.append(lambda s: opt.update(0, opt.padded, s, key="all"))
            ops.append(lambda s: self._repack(s))

    def _side_stream(self):
        if getattr(self, "_side", None) is None:
            self._side = L.Stream()
        return self._side

    def pre_step(self, stream_ptr):
        """Before the step graph: this step's random target labels (discriminator.py:240-254)."""
        if not self.label_smoothing or self._labels is None:
            return
        n, off = self._labels_n, self.smoothing_offset
        sr, hr = smoothed_labels(self._rng, (n,), (n,), True, off)
        if getattr(self, "_labels_host", None) is None or self._labels_host.shape[0] != 2 * n:
            self._labels_host = L.PinnedArray((2 * n,), np.float32)
        self._labels_host.array[:n] = hr
        self._labels_host.array[n:] = sr
        L.check(L.load().ssr_memcpy_h2d(self._labels.ptr, self._labels_host.ptr, 2 * n * 4, stream_ptr))

    def post_step(self, stream_ptr):
        self.iterations += 1

    def read_losses(self, stream_ptr=None):
        o = self._out.download((2,), np.float32, stream_ptr)
        return {self.metric_names[0]: float(o[0]), self.metric_names[1]: float(o[1])}


class GANLoss(RaGANLoss):
    """The non-relativistic critic step (Discriminator.initialize_standard, discriminator.py:306-361): AdversarialLoss for
    the generator, DiscriminatorLoss for the critic, sigmoid critics."""

    def __init__(self, discriminator, loss_weight=1e-3, **kw):
        kw.setdefault("relativistic", False)
        super().__init__(discriminator, loss_weight=loss_weight, **kw)
