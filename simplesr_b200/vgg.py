"""VGG19 feature extraction and the perceptual loss on the B200.

Mirrors ``model_builder.build_vgg_19`` / ``_custom_vgg`` (simple_sr/utils/models/model_builder.py:201-272: VGG19 convs
WITHOUT fused activation followed by separate ReLU layers, so that ``block5_conv4`` can be read before activation) and
``VGGLoss`` (simple_sr/utils/models/loss_functions/vgg_loss.py:59-180): denormalise [-1,1] -> [0,255], caffe-mode
``preprocess_input``, truncated VGG on HR and SR, ``MSE(features) * loss_weight``.

The network is frozen (``vgg.trainable = False``, vgg_loss.py:104): only the gradient w.r.t. the SR image is needed, i.e.
dgrad through the 16 convolutions, ReLU masks and max-pool routing - no wgrad.  Keras downloads the ImageNet weights;
there is no network here, so ``build_vgg_19`` starts from seeded He-normal weights at the real shapes and
``set_weights`` accepts the real ones (Keras order: kernel, bias per conv layer).
"""
import math

import numpy as np

from . import _lib as L
from .model_builder import Variable, get_context

VGG19_LAYERS = [("block1_conv1", 3, 64), ("block1_conv2", 64, 64), ("block1_pool",),
                ("block2_conv1", 64, 128), ("block2_conv2", 128, 128), ("block2_pool",),
                ("block3_conv1", 128, 256), ("block3_conv2", 256, 256), ("block3_conv3", 256, 256),
                ("block3_conv4", 256, 256), ("block3_pool",),
                ("block4_conv1", 256, 512), ("block4_conv2", 512, 512), ("block4_conv3", 512, 512),
                ("block4_conv4", 512, 512), ("block4_pool",),
                ("block5_conv1", 512, 512), ("block5_conv2", 512, 512), ("block5_conv3", 512, 512),
                ("block5_conv4", 512, 512), ("block5_pool",)]


class VGG19Model:
    """The custom VGG19 copy: host fp32 weights + packed device images (forward and dgrad)."""

    def __init__(self, seed=None, device=0):
        rng = np.random.default_rng(seed)
        self.ctx = get_context(device)
        self.kernels, self.biases = {}, {}
        for layer in VGG19_LAYERS:
            if len(layer) == 3:
                name, cin, cout = layer
                k = (rng.standard_normal((3, 3, cin, cout)) * math.sqrt(2.0 / (9 * cin))).astype(np.float32)
                self.kernels[name] = Variable(f"{name}/kernel:0", k, self._dirty)
                self.biases[name] = Variable(f"{name}/bias:0", np.zeros(cout, np.float32), self._dirty)
        self.trainable = False
        self._dev = None

    def _dirty(self):
        self._dev = None

    @property
    def weights(self):
        out = []
        for layer in VGG19_LAYERS:
            if len(layer) == 3:
                out += [self.kernels[layer[0]], self.biases[layer[0]]]
        return out

    def set_weights(self, weights):
        ws = self.weights
        if len(weights) != len(ws):
            raise ValueError(f"expected {len(ws)} weight arrays, got {len(weights)}")
        for v, w in zip(ws, weights):
            v.assign(w)

    def get_layer_names(self):
        return [l[0] for l in VGG19_LAYERS]

    def device_weights(self, stream=None):
        """{name: (packed_fwd, bias, packed_dgrad)} - packed once (the network is frozen)."""
        if self._dev is None:
            ctx, dev = self.ctx, {}
            for layer in VGG19_LAYERS:
                if len(layer) != 3:
                    continue
                name, cin, cout = layer
                cin_p, cout_p = -(-cin // 16) * 16, -(-cout // 16) * 16
                d_k = L.DeviceBuffer.from_numpy(self.kernels[name].numpy(), stream)
                fwd = L.DeviceBuffer(ctx.conv_packed_bytes(3, cin_p, cout, 1))
                ctx.conv_pack_weights(d_k, 3, cin, cin_p, cout, 1, fwd, stream)
                bwd = L.DeviceBuffer(ctx.conv_packed_bytes(3, cout_p, cin, 1))
                ctx.conv_pack_weights_dgrad(d_k, 3, 3, cin, cout, bwd, stream=stream)
                bias = L.DeviceBuffer.from_numpy(self.biases[name].numpy(), stream)
                L.stream_sync(stream)
                d_k.free()
                dev[name] = (fwd, bias, bwd)
            self._dev = dev
        return self._dev


def build_vgg_19(input_shape=(None, None), load_custom_weights=False, custom_weights_path=None, seed=2, device=0):
    """model_builder.build_vgg_19 (:201-225).  ``custom_weights_path``: a Keras ``.h5`` weight file as the reference
    passes to ``load_weights`` (:222) - e.g. the stock ``vgg19_weights_tf_dim_ordering_tf_kernels_notop.h5`` - or an
    .npz with the 32 arrays in Keras order."""
    model = VGG19Model(seed=seed, device=device)
    if load_custom_weights:
        if custom_weights_path is None:
            raise ValueError("no path for custom weights supplied")             # :217-218
        import os
        if not os.path.isfile(custom_weights_path):
            raise ValueError("can't locate custom weights in supplied path")   # :219-220
        from . import keras_h5
        if keras_h5.is_hdf5(custom_weights_path):
            model.set_weights(keras_h5.read_vgg19_file(custom_weights_path,
                                                       [l[0] for l in VGG19_LAYERS if len(l) == 3]))
        else:
            with np.load(custom_weights_path) as z:
                model.set_weights([z[k] for k in sorted(z.files)])
    return model


class VGGLoss:
    """vgg_loss.VGGLoss: features of one or several ``output_layers`` (summed, :162-164), taken before the ReLU
    (``after_activation=False``: the custom network, ESRGAN preset generator.py:436-437) or after it
    (``after_activation=True``, the reference's default :61 - the stock Keras VGG19 whose convs carry the ReLU), plus the
    optional total-variation term (:166-169).

    ``__call__(hr, sr, hr_critic, sr_critic, batch_metrics, epoch_metrics)`` keeps the loss-functor signature
    (generator.py:220-228) and returns the scalar; ``loss_and_grad`` also returns d(loss)/d(sr).  Inside a trainer the
    same launch list is emitted into the training graph (``emit``)."""

    metric_scale = 1.0          # out[0] already is the weighted sum the reference returns

    def __init__(self, output_layers="block5_conv4", feature_scale=1.0, loss_weight=1.0, total_variation_loss=False,
                 total_varation_weight=2 * 10e-8, after_activation=True, track_metrics=True, vgg=None, seed=2, device=0):
        self.name = "vgg_loss"
        self.metric_names = [self.name]
        self.feature_scale, self.loss_weight = float(feature_scale), float(loss_weight)
        self.weighted = self.loss_weight != 1.0                                  # vgg_loss.py:71-73
        self.total_variation_loss = bool(total_variation_loss)
        self.total_variation_weight = float(total_varation_weight)               # (sic) the reference's spelling, :60
        self.after_activation = bool(after_activation)
        self.track_metrics = track_metrics
        self.output_layers = output_layers if isinstance(output_layers, list) else [output_layers]
        self.vgg = vgg or build_vgg_19(seed=seed, device=device)
        names = self.vgg.get_layer_names()
        for l in self.output_layers:
            if l not in names or "pool" in l:
                raise ValueError(f"No such layer: {l}")
        order = [l for l in names if l in self.output_layers]
        self.layer = order[-1]                                                   # deepest requested layer
        self.ctx = self.vgg.ctx
        self._plans = {}
        self.loss = 0.0
        self.world = 1                 # data-parallel: set by attach(); the TV term is a SUM over the batch
        self.overlap_branches = True   # HR branch on a second stream (emit)
        self._side = None

    def attach(self, trainer):
        self.world = trainer.comm.world if getattr(trainer, "comm", None) is not None else 1

    def metric_floats(self, n):
        return 2 + n

    def _side_stream(self):
        if self._side is None:
            self._side = L.Stream()
        return self._side

    # ---- launch list ----------------------------------------------------------------------------------------------
    def emit(self, ops, B, prefix, n, H, W, hr_f32, sr_f32, g_sr, accumulate=True, out=None):
        """Appends to ``ops`` the launches computing the loss (into out[0]) and adding its gradient w.r.t. the SR image
        (fp32 [n,H,W,3] in [-1,1]) into ``g_sr``.  Buffers go into dict ``B``."""
        ctx = self.ctx
        dev = self.vgg.device_weights()
        want = set(self.output_layers)
        pre_act = not self.after_activation
        layers = []
        for layer in VGG19_LAYERS:
            layers.append(layer)
            if layer[0] == self.layer:
                break

        def buf(name, nbytes):
            B[prefix + name] = L.DeviceBuffer(nbytes)
            return B[prefix + name]

        def conv(x, xcs, cin, out_, cout, packed, bias, h, w, act, out_dtype=L.SSR_BF16, ocs=None):
            d = L.ConvDesc(n=n, h=h, w=w, cin=cin, in_cstride=xcs, cout=cout, ksize=3, ksize_w=3, act=act,
                           act_alpha=0.0, res_beta=0.0, up=1, out_dtype=out_dtype, out_cstride=(ocs or cout),
                           out_coff=0, res_dtype=L.SSR_NONE, res_cstride=0, res_coff=0, out2_cstride=0, out2_coff=0)
            ops.append(lambda s: ctx.conv2d_fwd(d, x, packed, bias, out_, stream=s))

        feats = {"hr": {}, "sr": {}}      # branch -> layer -> (fp32 feature buffer, elements, channels)
        saved = None
        # The HR branch does not depend on the SR branch: it runs on a second stream (fork here, join before the loss),
        # which fills the SMs the deep, small-spatial VGG layers leave idle.
        ops = L.OpsView(ops)
        side = self._side_stream() if self.overlap_branches else None
        hr_done = None
        if side is not None:
            fork, hr_done = L.Event(), L.Event()
            self._events = [fork, hr_done]
            ops.append(lambda s: (fork.record(s), side.wait_event(fork)))
        for branch, src in (("hr", hr_f32), ("sr", sr_f32)):
            ops.redirect = side if branch == "hr" else None
            px = n * H * W
            t = buf(f"{branch}_pre", px * 16 * 2)
            ops.append(lambda s, src=src, t=t, px=px: L.vgg_preprocess(src, t, px, s))
            tcs, h, w = 16, H, W
            trace = []   # ("conv", name, input, relu output | None, h, w, cin, cout) / ("pool", ...)
            for layer in layers:
                name = layer[0]
                if len(layer) == 3:
                    _, cin, cout = layer
                    fwd, bias, _ = dev[name]
                    cin_p, cnt = -(-cin // 16) * 16, n * h * w * cout
                    deepest = name == self.layer
                    if name in want and pre_act and deepest:
                        # the deepest pre-activation feature leaves the conv as fp32 directly (nothing follows it)
                        f32 = buf(f"{branch}_{name}_f32", cnt * 4)
                        conv(t, tcs, cin_p, f32, cout, fwd, bias, h, w, L.ACT_NONE, out_dtype=L.SSR_F32)
                        feats[branch][name] = (f32, cnt, cout)
                        y = None
                    elif name in want and pre_act:
                        # pre-activation feature in the middle of the chain: keep z, continue with relu(z)
                        z = buf(f"{branch}_{name}_z", cnt * 2)
                        conv(t, tcs, cin_p, z, cout, fwd, bias, h, w, L.ACT_NONE)
                        f32 = buf(f"{branch}_{name}_f32", cnt * 4)
                        ops.append(lambda s, z=z, f32=f32, cnt=cnt, cout=cout: L.bf16_to_f32(z, cout, 0, f32, cnt // cout, cout, s))
                        feats[branch][name] = (f32, cnt, cout)
                        y = buf(f"{branch}_{name}", cnt * 2)
                        ops.append(lambda s, z=z, y=y, cnt=cnt, cout=cout:
                                   L.act_fwd_bf16(z, cout, 0, None, 0.0, y, cout, 0, cnt // cout, cout, s))
                    else:
                        y = buf(f"{branch}_{name}", cnt * 2)
                        conv(t, tcs, cin_p, y, cout, fwd, bias, h, w, L.ACT_RELU)
                        if name in want:     # after_activation: the feature is the ReLU output
                            f32 = buf(f"{branch}_{name}_f32", cnt * 4)
                            ops.append(lambda s, y=y, f32=f32, cnt=cnt, cout=cout:
                                       L.bf16_to_f32(y, cout, 0, f32, cnt // cout, cout, s))
                            feats[branch][name] = (f32, cnt, cout)
                    trace.append(("conv", name, t, y, h, w, cin, cout))
                    t, tcs = y, cout
                else:
                    y = buf(f"{branch}_{name}", n * (h // 2) * (w // 2) * tcs * 2)
                    ops.append(lambda s, t=t, y=y, h=h, w=w, c=tcs: L.maxpool2_bf16(t, y, n, h, w, c, s))
                    trace.append(("pool", name, t, y, h, w, tcs, tcs))
                    t, h, w = y, h // 2, w // 2
            if branch == "hr" and side is not None:
                ops.redirect = None
                ops.append(lambda s: hr_done.record(side.ptr))
            if branch == "sr":
                saved = trace
        ops.redirect = None
        if side is not None:
            ops.append(lambda s: L.stream_wait_event(s, hr_done))
        # loss = sum_l loss_weight * mean((s f_sr - s f_hr)^2) = sum_l (loss_weight s^2) * MSE_l       vgg_loss.py:154-164
        wgt = self.loss_weight * self.feature_scale ** 2
        if out is None:
            out = buf("out", (2 + n) * 4)
        ops.append(lambda s: L.check(L.load().ssr_memset(out.ptr, 0, (2 + n) * 4, s)))
        ws = buf("ws", L.load().ssr_pixel_loss_workspace_bytes(n))
        d_feat16 = {}
        for name in self.output_layers:
            f_hr, cnt, fc = feats["hr"][name]
            f_sr = feats["sr"][name][0]
            tmp, d_feat = buf(f"out_{name}", (2 + n) * 4), buf(f"d_feat_{name}", cnt * 4)
            ops.append(lambda s, f_hr=f_hr, f_sr=f_sr, cnt=cnt, d_feat=d_feat, tmp=tmp:
                       L.pixel_loss(f_hr, f_sr, n, cnt // n, wgt, 0.0, 1.0, d_feat, ws, tmp, s))
            ops.append(lambda s, tmp=tmp: L.axpy_f32(tmp, out, wgt, 1, s))
            d16 = buf(f"d_feat_bf16_{name}", cnt * 2)
            ops.append(lambda s, d_feat=d_feat, d16=d16, cnt=cnt, fc=fc: L.f32_to_bf16_slice(d_feat, d16, fc, 0, cnt // fc, fc, s))
            d_feat16[name] = d16
        # ---- backward through the SR branch: dgrad convs, ReLU masks, max-pool routing; the feature gradients enter at
        # their layers (after the ReLU for post-activation features, before it for pre-activation ones)
        d, dcs = None, 0
        add = lambda a, b_, px_, c_: ops.append(lambda s: L.axpby_bf16(a, c_, 0, b_, c_, 0, 1.0, a, c_, 0, px_, c_, s))
        for i in reversed(range(len(saved))):
            kind, name, x_in, y, h, w, cin, cout = saved[i]
            if kind == "conv":
                pxl = n * h * w
                if name in want and not pre_act:
                    if d is None:
                        d, dcs = d_feat16[name], cout
                    else:
                        add(d, d_feat16[name], pxl, cout)
                if d is not None and y is not None:
                    dz = buf(f"dz_{name}", pxl * cout * 2)
                    ops.append(lambda s, d=d, dcs=dcs, y=y, dz=dz, pxl=pxl, cout=cout:
                               L.act_bwd_bf16(d, dcs, 0, y, cout, 0, None, 0.0, dz, cout, 0, pxl, cout, s))
                    d, dcs = dz, cout
                if name in want and pre_act:
                    if d is None:
                        d, dcs = d_feat16[name], cout
                    else:
                        add(d, d_feat16[name], pxl, cout)
                _, _, bwd = dev[name]
                if i == 0:
                    dx = buf("d_pre_f32", pxl * 3 * 4)
                    conv(d, dcs, cout, dx, 3, bwd, None, h, w, L.ACT_NONE, out_dtype=L.SSR_F32, ocs=3)
                    ops.append(lambda s, dx=dx, pxl=pxl: L.vgg_preprocess_bwd(dx, g_sr, pxl, 1.0, accumulate, s))
                else:
                    dx = buf(f"dx_{name}", pxl * cin * 2)
                    conv(d, dcs, cout, dx, cin, bwd, None, h, w, L.ACT_NONE)
                    d, dcs = dx, cin
            else:
                dx = buf(f"dx_{name}", n * h * w * cin * 2)
                ops.append(lambda s, x_in=x_in, d=d, dx=dx, h=h, w=w, cin=cin: L.maxpool2_bwd_bf16(x_in, d, dx, n, h, w,
                                                                                                 cin, s))
                d, dcs = dx, cin
        if self.total_variation_loss:
            # weight * reduce_sum(tf.image.total_variation(127.5 * (sr + 1))): a SUM over the batch, so a data-parallel
            # rank scales its share by the number of ranks (the ranks' losses and gradients are averaged afterwards)
            tv_ws = buf("tv_ws", L.load().ssr_total_variation_workspace_bytes())
            tv_out = buf("tv_out", 4)
            wtv = self.total_variation_weight * self.world
            ops.append(lambda s: L.total_variation(sr_f32, n, H, W, 3, 127.5, wtv, g_sr, tv_ws, tv_out, s))
            ops.append(lambda s: L.axpy_f32(tv_out, out, 1.0, 1, s))
        return out

    # ---- standalone use (numpy in, scalar out) -----------------------------------------------------------------------------
    def loss_and_grad(self, hr_batch, sr_batch):
        hr = np.ascontiguousarray(hr_batch, np.float32)
        sr = np.ascontiguousarray(sr_batch, np.float32)
        if hr.shape != sr.shape or hr.ndim != 4 or hr.shape[3] != 3:
            raise ValueError("hr and sr batches must be NHWC with 3 channels and equal shapes")
        n, H, W, _ = hr.shape
        if H % 16 or W % 16:
            raise ValueError("image sides must be multiples of 16 (four 2x2 poolings before block5)")
        key = (n, H, W)
        if key not in self._plans:
            B, ops = {}, []
            B["hr"], B["sr"], B["g"] = (L.DeviceBuffer(hr.nbytes) for _ in range(3))
            out = self.emit(ops, B, "vgg_", n, H, W, B["hr"], B["sr"], B["g"], accumulate=False)
            self._plans[key] = (B, ops, out)
        B, ops, out = self._plans[key]
        B["hr"].upload(hr)
        B["sr"].upload(sr)
        for op in ops:
            op(None)
        L.stream_sync(None)
        for st in (self._side,):
            if st is not None:
                st.sync()
        loss = float(out.download((2 + n,), np.float32)[0])
        return loss, B["g"].download(hr.shape, np.float32)

    def __call__(self, hr_batch, sr_batch, hr_critic=None, sr_critic=None, batch_metrics=None, epoch_metrics=None,
                 denormalize=True):
        if not denormalize:
            raise NotImplementedError("inputs must be in [-1, 1] (denormalize=True), as every reference preset uses")
        self.loss, _ = self.loss_and_grad(hr_batch, sr_batch)
        if self.track_metrics and batch_metrics is not None:
            batch_metrics[self.name](self.loss)
            epoch_metrics[self.name](self.loss)
        return self.loss

    def release(self):
        for B, _, _ in self._plans.values():
            for b in B.values():
                b.free()
        self._plans = {}
