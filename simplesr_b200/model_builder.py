"""Host-side mirror of ``simple_sr/utils/models/model_builder.py`` for the generator hot path.

The functions keep the reference's names, arguments and error behaviour
(``build_enhanced_resnet`` model_builder.py:42-96, ``build_resnet`` :99-134,
``build_or_load_generator_model`` :13-39) but return a :class:`GeneratorModel` whose forward pass is a
chain of ``ssr_conv2d_fwd`` launches (hand-written sm_100a kernels behind the C ABI) instead of a Keras
graph.  The object honours the duck-typed protocol the reference relies on
(``model(lr_batch, training=False)``, ``.trainable_variables``, ``.save(path)``), so it plugs into
``Generator(architecture=callable)`` / ``pretrained_model=`` (generator.py:124-137).

There is no CPU path: constructing a model needs a B200 and the built ``libssr_b200.so``.
"""
import math
import os

import numpy as np

from . import _lib as L

_CONTEXTS = {}


def get_context(device=0):
    if device not in _CONTEXTS:
        _CONTEXTS[device] = L.Context(device)
    return _CONTEXTS[device]


# ------------------------------------------------------------------------------------------------
# initialisers — same distributions as the reference; RNG streams differ from TF's
# ------------------------------------------------------------------------------------------------


def _he_normal_scaled(rng, shape, scale=0.2):
    """he_normal() with ``.scale = 0.2`` (model_builder.py:60-61): truncated normal, stddev
    sqrt(scale / fan_in) / 0.8796..., resampled outside two standard deviations."""
    fan_in = int(np.prod(shape[:-1]))
    std = math.sqrt(scale / fan_in) / 0.87962566103423978
    out = rng.standard_normal(size=shape)
    bad = np.abs(out) > 2.0
    while bad.any():
        out[bad] = rng.standard_normal(size=int(bad.sum()))
        bad = np.abs(out) > 2.0
    return (out * std).astype(np.float32)


def _glorot_uniform(rng, shape):
    rf = int(np.prod(shape[:-2]))
    limit = math.sqrt(6.0 / (shape[-2] * rf + shape[-1] * rf))
    return rng.uniform(-limit, limit, size=shape).astype(np.float32)


def npz_path(path):
    """The file ``GeneratorModel.save(path)`` writes / ``load_weights`` and the loader read: ``path`` with ``.npz``."""
    path = str(path)
    return path if path.endswith(".npz") else path + ".npz"


class Variable:
    """Minimal stand-in for a ``tf.Variable`` in ``model.trainable_variables``."""

    def __init__(self, name, value, on_assign=None):
        self.name = name
        self._value = np.ascontiguousarray(value, dtype=np.float32)
        self._on_assign = on_assign
        self._pull = None   # set by a trainer: refreshes _value from the device-resident master copy
        self._push = None   # set by a trainer: writes an assigned value into the device-resident master copy

    @property
    def shape(self):
        return self._value.shape

    def numpy(self):
        if self._pull is not None:
            self._pull()
        return self._value

    def assign(self, value):
        value = np.asarray(value, dtype=np.float32)
        if value.shape != self._value.shape:
            raise ValueError(f"shape mismatch assigning {self.name}: {value.shape} vs {self._value.shape}")
        self._value = np.ascontiguousarray(value)
        if self._push is not None:
            self._push(self._value)
        if self._on_assign:
            self._on_assign()


class _Conv:
    """One Conv2D(+fused tail) of the graph: host master weights + device copies."""

    def __init__(self, name, ksize, cin, cout, kernel, bias, up=1, alpha=None, unroll_x=False, bn=None):
        self.name, self.ksize, self.cout, self.up = name, ksize, cout, up
        # unroll_x: the kernel width is folded into the channels (kh x 1 conv over kw*cin channels of an x-unrolled
        # input, ssr_im2col_x_f32_to_bf16); HWIO [kh,kw,cin,cout] row-major IS [kh,1,kw*cin,cout], no repacking.
        self.kh = ksize
        self.kw = 1 if unroll_x else ksize
        self.cin_real = cin * ksize if unroll_x else cin
        self.cin = -(-self.cin_real // 16) * 16
        self.kernel = Variable(f"{name}/kernel:0", kernel, self._dirty)
        self.bias = Variable(f"{name}/bias:0", bias, self._dirty)
        self.alpha = Variable(f"{name}_prelu/alpha:0", alpha, self._dirty) if alpha is not None else None
        # BatchNormalization after the conv (model_builder.py:291-292).  At inference it is an affine map per output
        # channel and is folded into the packed weights / bias: no extra pass over the activations.
        self.bn = None
        if bn is not None:
            self.bn = {k: Variable(f"{name}_bn/{k}:0", bn[k], self._dirty)
                       for k in ("gamma", "beta", "moving_mean", "moving_variance")}
            self.bn_eps = float(bn.get("epsilon", 1e-3))
            self.bn_momentum = float(bn.get("momentum", 0.8))
        self.d_packed = self.d_bias = self.d_alpha = None
        self.d_packed_hi = self.d_packed_lo = None
        self.split_precision = False
        self.dirty = True

    def _dirty(self):
        self.dirty = True

    def variables(self):
        """Trainable variables in Keras layer order: Conv2D (kernel, bias), BatchNormalization (gamma, beta), PReLU."""
        v = [self.kernel, self.bias]
        if self.bn is not None:
            v += [self.bn["gamma"], self.bn["beta"]]
        if self.alpha is not None:
            v.append(self.alpha)
        return v

    def non_trainable_variables(self):
        return [self.bn["moving_mean"], self.bn["moving_variance"]] if self.bn is not None else []

    def effective_kernel_bias(self):
        """(HWIO kernel, bias) the device sees: the conv's own, with the inference-mode batch norm folded in."""
        k, b = self.kernel.numpy(), self.bias.numpy()
        if self.bn is None:
            return k, b
        s = self.bn["gamma"].numpy() / np.sqrt(self.bn["moving_variance"].numpy() + self.bn_eps)
        return (k * s).astype(np.float32), ((b - self.bn["moving_mean"].numpy()) * s + self.bn["beta"].numpy()).astype(np.float32)

    def sync(self, ctx, stream=None):
        """(Re)upload and repack the weights if they changed."""
        if not self.dirty:
            return
        if self.d_packed is None:
            self.d_packed = L.DeviceBuffer(ctx.conv_packed_bytes(self.kh, self.cin, self.cout, self.up,
                                                                 ksize_w=self.kw))
            self.d_bias = L.DeviceBuffer(self.cout * 4)
            if self.alpha is not None:
                self.d_alpha = L.DeviceBuffer(self.alpha.numpy().size * 4)
        k_eff, b_eff = self.effective_kernel_bias()
        d_w = L.DeviceBuffer.from_numpy(k_eff, stream)
        ctx.conv_pack_weights(d_w, self.kh, self.cin_real, self.cin, self.cout, self.up, self.d_packed, stream,
                              ksize_w=self.kw)
        if self.split_precision:
            # "fp32" precision mode: w = w_hi + w_lo as two bf16 images (always packed with up = 1: the depth_to_space
            # permutation is applied by the elementwise tail, ssr_act_split_f32)
            if self.d_packed_hi is None:
                nbytes = ctx.conv_packed_bytes(self.kh, self.cin, self.cout, 1, ksize_w=self.kw)
                self.d_packed_hi, self.d_packed_lo = L.DeviceBuffer(nbytes), L.DeviceBuffer(nbytes)
            k_lo = (k_eff - L.bf16_round(k_eff)).astype(np.float32)
            d_lo = L.DeviceBuffer.from_numpy(k_lo, stream)
            ctx.conv_pack_weights(d_w, self.kh, self.cin_real, self.cin, self.cout, 1, self.d_packed_hi, stream, ksize_w=self.kw)
            ctx.conv_pack_weights(d_lo, self.kh, self.cin_real, self.cin, self.cout, 1, self.d_packed_lo, stream, ksize_w=self.kw)
            L.stream_sync(stream)
            d_lo.free()
        self.d_bias.upload(b_eff, stream)
        if self.alpha is not None and hasattr(self.d_alpha, "upload"):
            self.d_alpha.upload(self.alpha.numpy(), stream)   # (a trainer's view into its flat buffer IS the master copy)
        L.stream_sync(stream)
        d_w.free()
        self.dirty = False


class _DerivedConv:
    """A conv whose packed weights are a function of the model's variables (growth-conv pairing): ``make()`` returns
    (HWIO kernel, bias) from the current host values."""

    def __init__(self, make, cin, cout, pair_split=False):
        self.make, self.cin_real, self.cout = make, cin, cout
        self.pair_split = pair_split        # image packed for the CTA-pair form of the kernel (desc.w_split = 2)
        self.cin = -(-cin // 16) * 16
        self.d_packed = self.d_bias = None
        self.dirty = True

    def sync(self, ctx, stream=None):
        if not self.dirty:
            return
        k, b = self.make()
        if self.d_packed is None:
            self.d_packed = L.DeviceBuffer(ctx.conv_packed_bytes(3, self.cin, self.cout, 1))
            self.d_bias = L.DeviceBuffer(self.cout * 4)
        d_w = L.DeviceBuffer.from_numpy(np.ascontiguousarray(k, np.float32), stream)
        if self.pair_split:
            ctx.conv_pack_weights_pair(d_w, self.cin_real, self.cin, self.cout, self.d_packed, stream)
        else:
            ctx.conv_pack_weights(d_w, 3, self.cin_real, self.cin, self.cout, 1, self.d_packed, stream)
        self.d_bias.upload(np.ascontiguousarray(b, np.float32), stream)
        L.stream_sync(stream)
        d_w.free()
        self.dirty = False


class _Plan:
    """Device buffers + recorded launch list of one forward pass for a fixed input shape."""

    def __init__(self, model, n, h, w):
        self.model, self.n, self.h, self.w = model, n, h, w
        self.buffers = {}
        self.ops = []          # list of zero-arg callables taking the stream pointer
        self.graph = None
        self.launches = 0
        self._order = 0
        # tile-level dependencies between consecutive conv launches (ssr_conv_chain_*): the plan opens a chain, every
        # conv asks for chain = 2 when the op right before it was a chained conv and 1 otherwise
        self.chain_buf = None
        self._prev_conv = False
        self.chain_stats = None      # (published, chained) of the most recent pass over the ops

    def open_chain(self):
        """First op of a plan whose model has ``chain_deps``: bump the epoch, arm desc.chain for this thread."""
        if not getattr(self.model, "chain_deps", False):
            return
        self.chain_buf = self.buf("conv_chain", L.load().ssr_conv_chain_bytes(self.n, self.h * self.model.upsample_factor,
                                                                              self.w * self.model.upsample_factor))
        self.chain_buf.zero()
        L.stream_sync(None)
        ctx, cb = self.model.ctx, self.chain_buf
        self.add(lambda s: ctx.conv_chain_begin(cb, s))

    def close_chain(self):
        if self.chain_buf is not None:
            ctx = self.model.ctx
            self.add(lambda s: setattr(self, "chain_stats", ctx.conv_chain_end()))
            self.launches -= 1      # bookkeeping only, no kernel

    def chain_flag(self):
        """desc.chain of the next conv launch (which must be added with ``conv=True``)."""
        if self.chain_buf is None:
            return 0
        flag = 2 if self._prev_conv else 1
        return flag

    def chain_timeouts(self):
        """Tile waits that gave up (word 1 of the chain buffer): 0 in a healthy run."""
        if self.chain_buf is None:
            return 0
        return int(self.chain_buf.download((2,), np.uint32)[1])

    def next_order(self):
        """desc.tile_order of the next conv launch: the layers alternate between walking their pixel tiles first to
        last and last to first, so each one starts with what its producer wrote last (the working set of a batch of
        16 exceeds the 126 MB L2; the tail of the previous layer's output is still resident)."""
        if getattr(self.model, "snake_order", "all") != "all" or (self.chain_buf is not None and not self.model.chain_snake):
            return 0       # chained launches walk their tiles in the same order: layer k+1 follows right behind layer k
        self._order ^= 1
        return self._order

    def buf(self, name, nbytes):
        if name not in self.buffers:
            self.buffers[name] = L.DeviceBuffer(nbytes)
        return self.buffers[name]

    def add(self, fn, conv=False):
        self.ops.append(fn)
        self.launches += 1
        self._prev_conv = bool(conv) and self.chain_buf is not None

    def run(self, stream_ptr, use_graph=True):
        if use_graph and stream_ptr is not None:
            if self.graph is None:
                self.graph = L.Graph(stream_ptr, lambda: [op(stream_ptr) for op in self.ops])
            self.graph.launch(stream_ptr)
        else:
            for op in self.ops:
                op(stream_ptr)

    def free(self):
        if self.graph is not None:
            self.graph.destroy()
            self.graph = None
        for b in self.buffers.values():
            b.free()
        self.buffers = {}


class GeneratorModel:
    """Keras-``Model``-like object produced by :func:`build_enhanced_resnet` / :func:`build_resnet`."""

    def __init__(self, architecture, upsample_factor, convs, config, device=0):
        self.architecture = architecture
        self.upsample_factor = upsample_factor
        self.config = dict(config)
        self.convs = convs                         # ordered dict name -> _Conv (Keras creation order)
        self.ctx = get_context(device)
        if os.environ.get("SSR_DBG"):            # development: kernel debug flags (ssr_debug_set)
            self.ctx.debug_set(int(os.environ["SSR_DBG"]))
        self.stream = L.Stream()
        self._plans = {}
        self.use_graph = True
        # RRDB inference: run the growth convs of a dense block in pairs (ssr_conv2d_fwd_carry); a trainer switches this
        # off because the paired weight images are derived from the host variables
        self.fuse_growth = True
        # the N = 64 pair launches on CTA pairs (cta_group::2): half of the weight rows per SM, M = 256 MMAs
        self.pair_growth = True
        # "all": consecutive conv launches walk their pixel tiles in opposite directions (_Plan.next_order);
        # "tails": only the carry consumers run last-to-first; "off"
        self.snake_order = os.environ.get("SSR_SNAKE", "all")
        # tile-level dependencies between the consecutive convs of the RRDB trunk (ssr_conv_chain_*): layer k+1 starts a
        # pixel tile as soon as the tiles of layer k its halo reads are stored, instead of waiting for the whole grid.
        # Bit-identical, but measured SLOWER on B200 (C2: 14.1 vs 12.6 ms): one CTA per SM means layer
        # k+1 cannot be resident before layer k's CTA has drained, so there is nothing to overlap, and the hardware's grid
        # dependency resolves faster than flag polling.  Off by default; kept as a tested option (SSR_CHAIN=1).
        self.chain_deps = os.environ.get("SSR_CHAIN", "0") == "1"
        self.chain_snake = False     # development: keep the alternating tile order inside a chain
        self._fused = {}
        self.precision = "bf16"

    def set_precision(self, precision):
        """"bf16" (default): bf16 operands, fp32 accumulation.  "fp32" (SRResNet inference only; BASELINE.json configs[0]
        is an fp32 configuration): activations and weights carried as (hi, lo) bf16 pairs, every convolution three
        tcgen05 passes accumulated in fp32 - products exact to ~2^-17, at three times the tensor work."""
        if precision not in ("bf16", "fp32"):
            raise ValueError("precision must be 'bf16' or 'fp32'")
        if precision == "fp32" and self.architecture != "srresnet":
            raise NotImplementedError("the fp32 precision mode is built for the SRResNet generator")
        if precision != self.precision:
            self.release()
            self.precision = precision
            for c in self.convs.values():
                c.split_precision = precision == "fp32"
                c.dirty = True
        return self

    # ---- Keras-like surface -------------------------------------------------------------------
    @property
    def trainable_variables(self):
        out = []
        for c in self.convs.values():
            out.extend(c.variables())
        return out

    @property
    def non_trainable_variables(self):
        """BatchNormalization moving statistics (empty without batch norm)."""
        out = []
        for c in self.convs.values():
            out.extend(c.non_trainable_variables())
        return out

    @property
    def variables(self):
        return self.trainable_variables + self.non_trainable_variables

    def get_weights(self):
        return [v.numpy() for v in self.variables]

    def set_weights(self, weights):
        tv = self.variables
        if len(weights) != len(tv):
            raise ValueError(f"expected {len(tv)} weight arrays, got {len(weights)}")
        for v, w in zip(tv, weights):
            v.assign(w)

    def count_params(self):
        return int(sum(v.numpy().size for v in self.trainable_variables))

    def save(self, path):
        """Keras ``model.save(path)`` stand-in (sr_model.py:244): writes ``<path>.npz`` (the suffix is added unless
        present) with the variables in order and the builder arguments, so that the file rebuilds the model on its own
        like the reference's ``load_model`` (model_builder.py:17-19)."""
        import json
        if str(path).endswith((".h5", ".hdf5")):
            # the reference's own format (sr_model.py:244 writes "<type>_gen_<epoch>.h5"): Keras HDF5 weight layout,
            # one group per layer in creation order; read back by load_weights / build_or_load_generator_model here and
            # by keras load_weights(by_name=False) there
            from . import keras_h5
            os.makedirs(os.path.dirname(os.path.abspath(str(path))) or ".", exist_ok=True)
            return keras_h5.write_model_file(str(path), self.variables)
        path = npz_path(path)
        arrays = {f"{i:04d}|{v.name}": v.numpy() for i, v in enumerate(self.variables)}
        os.makedirs(os.path.dirname(os.path.abspath(path)) or ".", exist_ok=True)
        cfg = {k: (list(v) if isinstance(v, tuple) else v) for k, v in self.config.items()}
        with open(path, "wb") as f:       # a file object: np.savez appends no suffix of its own
            np.savez(f, __architecture__=self.architecture, __upsample_factor__=self.upsample_factor,
                     __config__=json.dumps(cfg), **arrays)
        return path

    def load_weights(self, path):
        from . import keras_h5
        if keras_h5.is_hdf5(path):
            # Keras HDF5 (model.save / save_weights of the reference): matched by position, shapes checked by assign
            _, _, trainable, moving = keras_h5.read_generator_file(str(path))
            self.set_weights(trainable + moving)
            return
        with np.load(npz_path(path)) as z:
            keys = sorted(k for k in z.files if "|" in k)
            self.set_weights([z[k] for k in keys])

    def __call__(self, lr_batch, training=False, out=None):
        """``model(lr_batch, training=...)`` — generator.py:200, evaluation.py:357.  numpy in, numpy out.
        ``out`` optionally receives the result (e.g. a pinned host array) instead of a fresh allocation."""
        x = np.ascontiguousarray(lr_batch, dtype=np.float32)
        if x.ndim != 4 or x.shape[3] != 3:
            raise ValueError(f"expected NHWC input with 3 channels, got shape {x.shape}")
        if training and self.non_trainable_variables:
            # BatchNormalization(training=True): batch statistics + moving-average update (generator.py:200 passes
            # training= through; Generator.srresnet() defaults to batch_norm=True, generator.py:285)
            return _forward_srresnet_bn_training(self, x, out)
        n, h, w, _ = x.shape
        plan = self.plan(n, h, w)
        s = self.stream.ptr
        L.check(self.ctx.lib.ssr_memcpy_h2d(plan.buffers["in_f32"].ptr, x.ctypes.data, x.nbytes, s))
        plan.run(s, self.use_graph)
        sf = self.upsample_factor
        if out is None:
            out = np.empty((n, h * sf, w * sf, 3), dtype=np.float32)
        elif out.shape != (n, h * sf, w * sf, 3) or out.dtype != np.float32 or not out.flags.c_contiguous:
            raise ValueError("out must be a C-contiguous float32 array of shape [n, s*h, s*w, 3]")
        L.check(self.ctx.lib.ssr_memcpy_d2h(out.ctypes.data, plan.buffers["out_f32"].ptr, out.nbytes, s))
        self.stream.sync()
        return out

    # ---- device-level surface (bench / tiled inference keep data resident) ---------------------
    def sync_weights(self):
        changed = any(c.dirty for c in self.convs.values())
        for c in self.convs.values():
            c.sync(self.ctx, self.stream.ptr)
        if changed:
            for f in self._fused.values():
                f.dirty = True
        for f in self._fused.values():
            f.sync(self.ctx, self.stream.ptr)

    max_plans = 4   # cached input shapes (each holds a full activation set and a CUDA graph); least recently used goes

    def plan(self, n, h, w):
        self.sync_weights()
        key = (n, h, w)
        if key in self._plans:
            self._plans[key] = self._plans.pop(key)       # most recently used last
            return self._plans[key]
        while len(self._plans) >= max(1, self.max_plans):
            old = next(iter(self._plans))
            self.stream.sync()
            self._plans.pop(old).free()
        if self.architecture == "rrdb":
            self._plans[key] = _plan_rrdb(self, n, h, w)
        elif self.architecture == "srresnet":
            self._plans[key] = (_plan_srresnet_fp32 if self.precision == "fp32" else _plan_srresnet)(self, n, h, w)
        else:
            raise NotImplementedError(self.architecture)
        return self._plans[key]

    def release(self):
        for p in self._plans.values():
            p.free()
        self._plans = {}
        for d_img, d_out, copy_stream, events in self.__dict__.pop("_tiled_buffers", {}).values():
            d_img.free()
            d_out.free()
            copy_stream.destroy()
            for ev in events:
                ev.destroy()


def _conv_op(plan, ctx, conv, n, h, w, x, in_cstride, out, out_cstride, out_coff, act=L.ACT_NONE, act_alpha=0.2,
             res=None, res_cstride=0, res_coff=0, res_beta=1.0, out_dtype=L.SSR_BF16, out2=None, out2_cstride=0,
             out2_coff=0, cin=None):
    d = L.ConvDesc(n=n, h=h, w=w, cin=cin or conv.cin, in_cstride=in_cstride, in_cvalid=in_cstride, cout=conv.cout,
                   ksize=conv.kh, ksize_w=conv.kw,
                   act=act, act_alpha=act_alpha, res_beta=res_beta, up=conv.up, out_dtype=out_dtype,
                   out_cstride=out_cstride, out_coff=out_coff,
                   res_dtype=(L.SSR_BF16 if res is not None else L.SSR_NONE), res_cstride=res_cstride,
                   res_coff=res_coff, out2_cstride=out2_cstride, out2_coff=out2_coff, tile_order=plan.next_order(),
                   chain=plan.chain_flag())
    plan.add(lambda s, d=d: ctx.conv2d_fwd(d, x, conv.d_packed, conv.d_bias, out, alpha=conv.d_alpha, res=res,
                                           out2=out2, stream=s), conv=True)


def _fused_growth_ops(p, m, pre, n, h, w, src, cw):
    """The four growth convs of a dense block (model_builder.py:333-338) as two pairs.  An MMA over N = 32 output
    channels costs the tensor core as many clocks as one over N = 64 (operand reads dominate), so conv k (k = 0, 2) is
    launched with the kernels of conv k and conv k+1 (restricted to conv k's input channels) side by side: 32 activated
    channels + 32 fp32 partial sums (``carry``); conv k+1 then only convolves the 32 channels conv k just produced and adds
    the carry before bias + LeakyReLU.  28 -> 20 K-steps of A-operand reads per pixel tile and dense block."""
    ctx, c = m.ctx, m.convs
    px = n * h * w
    carry = p.buf("growth_carry", ctx.lib.ssr_conv2d_carry_elems(ctx.handle, n, h, w) * 4)
    for k in (0, 2):
        ca, cb = c[f"{pre}_conv{k}"], c[f"{pre}_conv{k + 1}"]
        cin_a = 64 + 32 * k
        key_a, key_b = f"{pre}_pair{k}", f"{pre}_tail{k + 1}"
        if key_a not in m._fused:
            m._fused[key_a] = _DerivedConv(
                lambda ca=ca, cb=cb, cin_a=cin_a: (np.concatenate([ca.kernel.numpy(), cb.kernel.numpy()[:, :, :cin_a, :]],
                                                                  axis=3),
                                                   np.concatenate([ca.bias.numpy(), np.zeros(32, np.float32)])),
                cin_a, 64, pair_split=(m.pair_growth and cin_a >= 128))   # K = 64 is too short to amortise the pair's barriers
            m._fused[key_b] = _DerivedConv(
                lambda cb=cb, cin_a=cin_a: (cb.kernel.numpy()[:, :, cin_a:cin_a + 32, :], cb.bias.numpy()), 32, 32)
            m._fused[key_a].sync(ctx, m.stream.ptr)
            m._fused[key_b].sync(ctx, m.stream.ptr)
        fa, fb = m._fused[key_a], m._fused[key_b]
        da = L.ConvDesc(n=n, h=h, w=w, cin=cin_a, in_cstride=cw, in_cvalid=cw, cout=64, ksize=3, ksize_w=3,
                        act=L.ACT_LRELU,
                        act_alpha=0.2, res_beta=0.0, up=1, out_dtype=L.SSR_BF16, out_cstride=cw, out_coff=cin_a,
                        res_dtype=L.SSR_NONE, res_cstride=0, res_coff=0, out2_cstride=0, out2_coff=0,
                        w_split=(2 if fa.pair_split else 0), tile_order=p.next_order(), chain=p.chain_flag())
        p.add(lambda s, da=da, fa=fa: ctx.conv2d_fwd_carry(da, src, fa.d_packed, fa.d_bias, src, carry_out=carry,
                                                            carry_out_cols=32, stream=s), conv=True)
        xb = L.DeviceView(src, cin_a * 2, src.nbytes - cin_a * 2)      # the 32 channels conv k just wrote
        db = L.ConvDesc(n=n, h=h, w=w, cin=32, in_cstride=cw, in_cvalid=cw - cin_a, cout=32, ksize=3, ksize_w=3,
                        act=L.ACT_LRELU,
                        act_alpha=0.2, res_beta=0.0, up=1, out_dtype=L.SSR_BF16, out_cstride=cw, out_coff=cin_a + 32,
                        res_dtype=L.SSR_NONE, res_cstride=0, res_coff=0, out2_cstride=0, out2_coff=0,
                        tile_order=(1 if m.snake_order == "tails" and p.chain_buf is None else p.next_order()),
                        chain=p.chain_flag())
        p.add(lambda s, db=db, fb=fb, xb=xb: ctx.conv2d_fwd_carry(db, xb, fb.d_packed, fb.d_bias, src, carry_in=carry,
                                                                   stream=s), conv=True)


def _plan_rrdb(m, n, h, w):
    """Launch list of build_enhanced_resnet (model_builder.py:42-96, 328-365).

    Buffers: two ping-pong [n,h,w,192] bf16 tensors hold [x | c1 | c2 | c3 | c4] of the current dense
    block, so Concatenate (:338) is a channel offset; the 192->64 conv writes ``res + 0.2*conv`` (:349-350)
    into channels [0,64) of the other buffer.
    """
    cfg = m.config
    nf, gc = cfg["num_filters"], cfg["num_filters"] // 2
    nb, ndb, nc = cfg["num_rrdb_blocks"], cfg["num_dense_blocks"], cfg["num_convs"]
    beta = cfg["residual_scaling_factor"]
    cw = nf + nc * gc                      # 192
    sf = m.upsample_factor
    ctx = m.ctx
    px = n * h * w
    p = _Plan(m, n, h, w)
    in_f32 = p.buf("in_f32", px * 3 * 4)
    x16 = p.buf("x16", px * 16 * 2)
    fea = p.buf("fea", px * nf * 2)
    bufs = [p.buf("dense_a", px * cw * 2), p.buf("dense_b", px * cw * 2)]
    c = m.convs

    p.open_chain()
    p.add(lambda s: L.f32_to_bf16_pad(in_f32, x16, px, 3, 16, s))
    _conv_op(p, ctx, c["fea"], n, h, w, x16, 16, fea, nf, 0, out2=bufs[0], out2_cstride=cw, out2_coff=0)
    cur = 0
    for b in range(nb):
        for d in range(ndb):
            src, dst = bufs[cur], bufs[1 - cur]
            if m.fuse_growth and nc == 4 and nf == 64 and gc == 32:
                _fused_growth_ops(p, m, f"rrdb{b}_db{d}", n, h, w, src, cw)
            else:
                for k in range(nc):
                    _conv_op(p, ctx, c[f"rrdb{b}_db{d}_conv{k}"], n, h, w, src, cw, src, cw, nf + k * gc,
                             act=L.ACT_LRELU, act_alpha=0.2)
            _conv_op(p, ctx, c[f"rrdb{b}_db{d}_out"], n, h, w, src, cw, dst, cw, 0, res=src, res_cstride=cw,
                     res_coff=0, res_beta=beta)
            cur = 1 - cur
    src, dst = bufs[cur], bufs[1 - cur]
    # trunk_in = fea + beta * r  (model_builder.py:363-364)
    p.add(lambda s: L.axpby_bf16(fea, nf, 0, src, cw, 0, beta, dst, cw, 0, px, nf, s))
    u = p.buf("u0", px * nf * 2)
    _conv_op(p, ctx, c["trunk"], n, h, w, dst, cw, u, nf, 0, res=fea, res_cstride=nf, res_coff=0, res_beta=1.0)
    hh, ww = h, w
    for i in range(int(math.log(sf, 2))):
        nxt = p.buf(f"u{i + 1}", n * (2 * hh) * (2 * ww) * nf * 2)
        _conv_op(p, ctx, c[f"up{i}"], n, hh, ww, u, nf, nxt, nf, 0, act=L.ACT_LRELU, act_alpha=0.2)
        u, hh, ww = nxt, 2 * hh, 2 * ww
    v = p.buf("hr", n * hh * ww * nf * 2)
    _conv_op(p, ctx, c["hr"], n, hh, ww, u, nf, v, nf, 0, act=L.ACT_LRELU, act_alpha=0.2)
    out = p.buf("out_f32", n * hh * ww * 3 * 4)
    _conv_op(p, ctx, c["last"], n, hh, ww, v, nf, out, 3, 0, act=L.ACT_TANH, out_dtype=L.SSR_F32)
    p.close_chain()
    return p


def _plan_srresnet(m, n, h, w):
    """Launch list of build_resnet without batch norm (model_builder.py:99-134, 309-325).

    The 9x9x3 input convolution (:117) runs as a 9x1 convolution over the x-unrolled image (27 -> 32 channels);
    PReLU (:118,:314,:281) and the skip additions (:318,:126) are conv epilogues; depth_to_space (:279) is the
    store-address map of the up-convolutions (PReLU commutes with the permutation, its slopes are per output channel).
    """
    cfg = m.config
    nf, nb = cfg["num_filters"], cfg["num_res_blocks"]
    sf = m.upsample_factor
    ctx = m.ctx
    px = n * h * w
    p = _Plan(m, n, h, w)
    c = m.convs
    in_f32 = p.buf("in_f32", px * 3 * 4)
    x32 = p.buf("x_unrolled", px * c["first"].cin * 2)
    skip = p.buf("skip", px * nf * 2)
    ta, tb, u = p.buf("t_a", px * nf * 2), p.buf("t_b", px * nf * 2), p.buf("u", px * nf * 2)

    p.add(lambda s: L.im2col_x_f32_to_bf16(in_f32, x32, n, h, w, 3, 9, c["first"].cin, s))
    _conv_op(p, ctx, c["first"], n, h, w, x32, c["first"].cin, skip, nf, 0, act=L.ACT_PRELU)
    cur, nxt = skip, ta
    for b in range(nb):
        _conv_op(p, ctx, c[f"res{b}_conv0"], n, h, w, cur, nf, u, nf, 0, act=L.ACT_PRELU)
        _conv_op(p, ctx, c[f"res{b}_conv1"], n, h, w, u, nf, nxt, nf, 0, res=cur, res_cstride=nf, res_coff=0,
                 res_beta=1.0)
        cur, nxt = nxt, (tb if nxt is ta else ta)
    _conv_op(p, ctx, c["trunk"], n, h, w, cur, nf, nxt, nf, 0, res=skip, res_cstride=nf, res_coff=0, res_beta=1.0)
    t, hh, ww = nxt, h, w
    for i in range(int(math.log(sf, 2))):
        up = p.buf(f"up{i}", n * (2 * hh) * (2 * ww) * nf * 2)
        _conv_op(p, ctx, c[f"up{i}"], n, hh, ww, t, nf, up, nf, 0, act=L.ACT_PRELU)
        t, hh, ww = up, 2 * hh, 2 * ww
    out = p.buf("out_f32", n * hh * ww * 3 * 4)
    _conv_op(p, ctx, c["last"], n, hh, ww, t, nf, out, 3, 0, act=L.ACT_TANH, out_dtype=L.SSR_F32)
    return p


def _plan_srresnet_fp32(m, n, h, w):
    """build_resnet (model_builder.py:99-134) in the "fp32" precision mode: every activation is an fp32 tensor carried
    as a (hi | lo) bf16 channel pair, every convolution = three tcgen05 passes accumulated in fp32 through the conv
    kernel's fp32 residual input (a_hi*w_hi + a_lo*w_hi + a_hi*w_lo), the residual / skip stream stays fp32, and
    ssr_act_split_f32 applies PReLU / tanh / the skip additions / depth_to_space between convolutions."""
    cfg = m.config
    nf, nb = cfg["num_filters"], cfg["num_res_blocks"]
    sf = m.upsample_factor
    ctx = m.ctx
    px = n * h * w
    p = _Plan(m, n, h, w)
    c = m.convs
    cin0 = c["first"].cin
    in_f32, in_lo = p.buf("in_f32", px * 3 * 4), p.buf("in_lo_f32", px * 3 * 4)
    x_hi, x_lo = p.buf("x_unrolled_hi", px * cin0 * 2), p.buf("x_unrolled_lo", px * cin0 * 2)
    p.add(lambda s: L.check(L.load().ssr_bf16_residual_f32(in_f32.ptr, in_lo.ptr, px * 3, s)))
    p.add(lambda s: L.im2col_x_f32_to_bf16(in_f32, x_hi, n, h, w, 3, 9, cin0, s))
    p.add(lambda s: L.im2col_x_f32_to_bf16(in_lo, x_lo, n, h, w, 3, 9, cin0, s))

    def conv3(cv, hh, ww, a_hi, a_lo, cs, z):
        """z (fp32 [.., cout]) = conv(a; w) + bias, a = a_hi + a_lo, w = w_hi + w_lo."""
        for k, (x, packed) in enumerate(((a_hi, cv.d_packed_hi), (a_lo, cv.d_packed_hi), (a_hi, cv.d_packed_lo))):
            d = L.ConvDesc(n=n, h=hh, w=ww, cin=cv.cin, in_cstride=cs, in_cvalid=cv.cin, cout=cv.cout, ksize=cv.kh,
                           ksize_w=cv.kw, act=L.ACT_NONE, act_alpha=0.0, res_beta=1.0, up=1, out_dtype=L.SSR_F32,
                           out_cstride=cv.cout, out_coff=0, res_dtype=(L.SSR_F32 if k else L.SSR_NONE),
                           res_cstride=cv.cout, res_coff=0, out2_cstride=0, out2_coff=0)
            bias = cv.d_bias if k == 0 else None
            p.add(lambda s, d=d, x=x, packed=packed, bias=bias, k=k: ctx.conv2d_fwd(d, x, packed, bias, z,
                                                                                   res=(z if k else None), stream=s))

    def halves(buf):                      # (hi, lo) views of a [.., 2 nf] bf16 buffer
        return buf, L.DeviceView(buf, nf * 2, buf.nbytes - nf * 2)

    def tail(z, hh, ww, cout, up, act, alpha, res32, y32, hl):
        p.add(lambda s: L.act_split_f32(z, n, hh, ww, cout, up, act, 0.0, alpha, res32, y32, hl, 2 * nf, 0, nf, s))

    z = p.buf("z_f32", px * nf * 4)
    skip32 = p.buf("skip_f32", px * nf * 4)
    hl = [p.buf("hl_a", px * 2 * nf * 2), p.buf("hl_b", px * 2 * nf * 2)]
    hl_u = p.buf("hl_u", px * 2 * nf * 2)
    t32 = [skip32, p.buf("t32_a", px * nf * 4), p.buf("t32_b", px * nf * 4)]
    conv3(c["first"], h, w, x_hi, x_lo, cin0, z)
    tail(z, h, w, nf, 1, L.ACT_PRELU, c["first"].d_alpha, None, skip32, hl[0])
    cur, cur32 = 0, 0
    for b in range(nb):
        c0, c1 = c[f"res{b}_conv0"], c[f"res{b}_conv1"]
        conv3(c0, h, w, *halves(hl[cur]), 2 * nf, z)
        tail(z, h, w, nf, 1, L.ACT_PRELU, c0.d_alpha, None, None, hl_u)
        conv3(c1, h, w, *halves(hl_u), 2 * nf, z)
        nxt32 = 1 if cur32 != 1 else 2
        tail(z, h, w, nf, 1, L.ACT_NONE, None, t32[cur32], t32[nxt32], hl[1 - cur])     # x_in + conv1(...), :318
        cur, cur32 = 1 - cur, nxt32
    conv3(c["trunk"], h, w, *halves(hl[cur]), 2 * nf, z)
    tail(z, h, w, nf, 1, L.ACT_NONE, None, skip32, None, hl[1 - cur])                    # + skip, :126
    src, hh, ww = hl[1 - cur], h, w
    for i in range(int(math.log(sf, 2))):
        cv = c[f"up{i}"]
        zu = p.buf(f"zu{i}_f32", n * hh * ww * cv.cout * 4)
        conv3(cv, hh, ww, *halves(src), 2 * nf, zu)
        up_hl = p.buf(f"hl_up{i}", n * 4 * hh * ww * 2 * nf * 2)
        p.add(lambda s, zu=zu, hh=hh, ww=ww, cv=cv, up_hl=up_hl:
              L.act_split_f32(zu, n, hh, ww, nf, 2, L.ACT_PRELU, 0.0, cv.d_alpha, None, None, up_hl, 2 * nf, 0, nf, s))
        src, hh, ww = up_hl, 2 * hh, 2 * ww
    zl = p.buf("z_last_f32", n * hh * ww * 3 * 4)
    conv3(c["last"], hh, ww, *halves(src), 2 * nf, zl)
    out = p.buf("out_f32", n * hh * ww * 3 * 4)
    p.add(lambda s: L.act_split_f32(zl, n, hh, ww, 3, 1, L.ACT_TANH, 0.0, None, None, out, None, 0, 0, 0, s))
    return p


def _forward_srresnet_bn_training(m, x, out=None):
    """``model(lr_batch, training=True)`` of build_resnet WITH batch normalisation (model_builder.py:291-292, 309-319):
    every BatchNormalization layer normalises with the statistics of this batch and moves its moving mean / variance
    (momentum, unbiased variance) - what Keras does whenever the layer is called with training=True.  Eager launches
    (this is not the inference hot path): raw conv -> ssr_bn_stats_bf16 -> affine -> PReLU / skip."""
    if m.architecture != "srresnet":
        raise NotImplementedError("batch normalisation only exists in the SRResNet generator")
    ctx, c, cfg = m.ctx, m.convs, m.config
    nf, nb, sf = cfg["num_filters"], cfg["num_res_blocks"], m.upsample_factor
    n, h, w, _ = x.shape
    px = n * h * w
    s = m.stream.ptr
    m.sync_weights()
    keep = []

    def buf(nbytes):
        keep.append(L.DeviceBuffer(nbytes))
        return keep[-1]

    def conv(cv, hh, ww, src, src_cs, dst, packed=None, bias=None, act=L.ACT_NONE, res=None, out_dtype=L.SSR_BF16, ocs=None):
        d = L.ConvDesc(n=n, h=hh, w=ww, cin=cv.cin, in_cstride=src_cs, in_cvalid=src_cs, cout=cv.cout, ksize=cv.kh,
                       ksize_w=cv.kw, act=act, act_alpha=0.0, res_beta=1.0, up=cv.up, out_dtype=out_dtype,
                       out_cstride=(ocs or nf), out_coff=0, res_dtype=(L.SSR_BF16 if res is not None else L.SSR_NONE),
                       res_cstride=nf, res_coff=0, out2_cstride=0, out2_coff=0)
        ctx.conv2d_fwd(d, src, packed or cv.d_packed, bias or cv.d_bias, dst, alpha=cv.d_alpha, res=res, stream=s)

    bn_ws = buf(L.load().ssr_bn_workspace_bytes(nf))
    mean, istd = buf(nf * 4), buf(nf * 4)

    def conv_bn(name, src):
        """raw conv (the model's own image has the inference statistics folded in) + batch-statistics normalisation"""
        cv = c[name]
        d_w = L.DeviceBuffer.from_numpy(cv.kernel.numpy(), s)
        raw = buf(ctx.conv_packed_bytes(cv.kh, cv.cin, cv.cout, cv.up, ksize_w=cv.kw))
        ctx.conv_pack_weights(d_w, cv.kh, cv.cin_real, cv.cin, cv.cout, cv.up, raw, s, ksize_w=cv.kw)
        keep.append(d_w)
        bias = L.DeviceBuffer.from_numpy(cv.bias.numpy(), s)
        keep.append(bias)
        z, y = buf(px * nf * 2), buf(px * nf * 2)
        conv(cv, h, w, src, nf, z, packed=raw, bias=bias)
        mm = L.DeviceBuffer.from_numpy(cv.bn["moving_mean"].numpy(), s)
        mv = L.DeviceBuffer.from_numpy(cv.bn["moving_variance"].numpy(), s)
        g = L.DeviceBuffer.from_numpy(cv.bn["gamma"].numpy(), s)
        be = L.DeviceBuffer.from_numpy(cv.bn["beta"].numpy(), s)
        keep.extend([mm, mv, g, be])
        L.bn_stats_bf16(z, px, nf, cv.bn_eps, cv.bn_momentum, bn_ws, mean, istd, mm, mv, s)
        L.bn_lrelu_fwd_bf16(z, mean, istd, g, be, 1.0, y, px, nf, s)
        cv.bn["moving_mean"].assign(mm.download((nf,), np.float32, s))
        cv.bn["moving_variance"].assign(mv.download((nf,), np.float32, s))
        return y

    in_f32 = buf(x.nbytes)
    in_f32.upload(x, s)
    x32 = buf(px * c["first"].cin * 2)
    L.im2col_x_f32_to_bf16(in_f32, x32, n, h, w, 3, 9, c["first"].cin, s)
    skip = buf(px * nf * 2)
    conv(c["first"], h, w, x32, c["first"].cin, skip, act=L.ACT_PRELU)
    t = skip
    for b in range(nb):
        y0 = conv_bn(f"res{b}_conv0", t)
        u = buf(px * nf * 2)
        L.act_fwd_bf16(y0, nf, 0, c[f"res{b}_conv0"].d_alpha, 0.0, u, nf, 0, px, nf, s)
        y1 = conv_bn(f"res{b}_conv1", u)
        t_out = buf(px * nf * 2)
        L.axpby_bf16(t, nf, 0, y1, nf, 0, 1.0, t_out, nf, 0, px, nf, s)
        t = t_out
    yt = conv_bn("trunk", t)
    cur = buf(px * nf * 2)
    L.axpby_bf16(skip, nf, 0, yt, nf, 0, 1.0, cur, nf, 0, px, nf, s)
    hh, ww = h, w
    for i in range(int(math.log(sf, 2))):
        up = buf(n * 4 * hh * ww * nf * 2)
        conv(c[f"up{i}"], hh, ww, cur, nf, up, act=L.ACT_PRELU)
        cur, hh, ww = up, 2 * hh, 2 * ww
    d_out = buf(n * hh * ww * 3 * 4)
    conv(c["last"], hh, ww, cur, nf, d_out, act=L.ACT_TANH, out_dtype=L.SSR_F32, ocs=3)
    if out is None:
        out = np.empty((n, hh, ww, 3), np.float32)
    elif out.shape != (n, hh, ww, 3) or out.dtype != np.float32 or not out.flags.c_contiguous:
        raise ValueError("out must be a C-contiguous float32 array of shape [n, s*h, s*w, 3]")
    L.check(ctx.lib.ssr_memcpy_d2h(out.ctypes.data, d_out.ptr, out.nbytes, s))
    m.stream.sync()
    for b in keep:
        b.free()
    return out


# ------------------------------------------------------------------------------------------------
# builders (reference names)
# ------------------------------------------------------------------------------------------------


def build_enhanced_resnet(upsample_factor=2, num_filters=64, num_rrdb_blocks=16, num_dense_blocks=3, num_convs=4,
                          kernel_size=3, residual_scaling_factor=0.2, input_dims=(None, None), seed=None, device=0):
    """RRDB generator — same signature and defaults as model_builder.build_enhanced_resnet (:42-44)."""
    if upsample_factor not in [2, 4, 8]:
        raise ValueError("upsample factor not supported - please choose either 2, 4 or 8")
    if kernel_size != 3:
        raise ValueError("only kernel_size=3 is supported by the sm_100a RRDB path")
    if num_filters % 32 != 0:
        raise ValueError("num_filters must be a multiple of 32 (growth channels are num_filters//2)")
    rng = np.random.default_rng(seed)
    nf, gc = num_filters, num_filters // 2
    convs = {}

    def add(name, cin, cout, up=1):
        convs[name] = _Conv(name, 3, cin, cout, _he_normal_scaled(rng, (3, 3, cin, cout)),
                            np.zeros(cout, np.float32), up=up)

    add("fea", 3, nf)
    for b in range(num_rrdb_blocks):
        for d in range(num_dense_blocks):
            for k in range(num_convs):
                add(f"rrdb{b}_db{d}_conv{k}", nf + k * gc, gc)
            add(f"rrdb{b}_db{d}_out", nf + num_convs * gc, nf)
    add("trunk", nf, nf)
    for u in range(int(math.log(upsample_factor, 2))):
        add(f"up{u}", nf, nf * 4, up=2)
    add("hr", nf, nf)
    add("last", nf, 3)
    cfg = dict(num_filters=num_filters, num_rrdb_blocks=num_rrdb_blocks, num_dense_blocks=num_dense_blocks,
               num_convs=num_convs, residual_scaling_factor=residual_scaling_factor, input_dims=input_dims)
    return GeneratorModel("rrdb", upsample_factor, convs, cfg, device=device)


def build_resnet(upsample_factor=2, num_filters=64, num_res_blocks=16, momentum=0.8, input_dims=(None, None),
                 batch_normalization=True, seed=None, device=0, precision="bf16"):
    """SRResNet generator - same signature and defaults as model_builder.build_resnet (:99-100) plus seed / device;
    Keras default initialisers (glorot_uniform kernels, zero biases, zero PReLU slopes, batch norm gamma 1 / beta 0 /
    moving mean 0 / moving variance 1).  ``batch_normalization`` puts a BatchNormalization(momentum) after both convs
    of every res block and after the trunk conv (:309-319, :123-125); inference folds it into the convs, training
    (simplesr_b200.training.SRResNetTrainer) runs it with batch statistics and updates the moving averages."""
    if upsample_factor not in [2, 4, 8]:
        raise ValueError("upsample factor not supported - please choose either 2, 4 or 8")   # :113-114
    rng = np.random.default_rng(seed)
    nf = num_filters
    convs = {}

    def add(name, ks, cin, cout, prelu, up=1, unroll_x=False, bn=False):
        ac = cout // 4 if up == 2 else cout
        bn_init = None
        if bn and batch_normalization:
            bn_init = dict(gamma=np.ones(cout, np.float32), beta=np.zeros(cout, np.float32),
                           moving_mean=np.zeros(cout, np.float32), moving_variance=np.ones(cout, np.float32),
                           momentum=momentum, epsilon=1e-3)
        convs[name] = _Conv(name, ks, cin, cout, _glorot_uniform(rng, (ks, ks, cin, cout)), np.zeros(cout, np.float32),
                            up=up, alpha=(np.zeros(ac, np.float32) if prelu else None), unroll_x=unroll_x, bn=bn_init)

    add("first", 9, 3, nf, True, unroll_x=True)
    for b in range(num_res_blocks):
        add(f"res{b}_conv0", 3, nf, nf, True, bn=True)
        add(f"res{b}_conv1", 3, nf, nf, False, bn=True)
    add("trunk", 3, nf, nf, False, bn=True)
    for u in range(int(math.log(upsample_factor, 2))):
        add(f"up{u}", 3, nf, nf * 4, True, up=2)
    add("last", 9, nf, 3, False)
    cfg = dict(num_filters=num_filters, num_res_blocks=num_res_blocks, batch_norm=bool(batch_normalization),
               momentum=momentum, input_dims=input_dims)
    return GeneratorModel("srresnet", upsample_factor, convs, cfg, device=device).set_precision(precision)


def build_or_load_generator_model(upsample_factor, architecture, num_blocks, num_filters, kernel_size,
                                  residual_scaling, kernel_initializer, batch_norm, input_dims, num_convs=4,
                                  num_dense_blocks=3, pretrained_model_path=None):
    """Same dispatch as model_builder.build_or_load_generator_model (:13-39)."""
    if pretrained_model_path is not None:
        import json
        from . import keras_h5
        if keras_h5.is_hdf5(pretrained_model_path):
            # the reference's generator files (sr_model.py:244): the graph is rebuilt from the weight shapes, the
            # arguments of the call are ignored as with keras load_model (:17-19) - except residual_scaling, which
            # Keras keeps in a Lambda layer and the weights cannot tell
            arch, kw, trainable, moving = keras_h5.read_generator_file(str(pretrained_model_path))
            if arch == "rrdb":
                model = build_enhanced_resnet(residual_scaling_factor=0.2 if residual_scaling is None
                                              else residual_scaling, input_dims=input_dims or (None, None), **kw)
            else:
                model = build_resnet(input_dims=input_dims or (None, None), **kw)
            model.set_weights(trainable + moving)
            return model
        with np.load(npz_path(pretrained_model_path)) as z:
            arch = str(z["__architecture__"])
            sf = int(z["__upsample_factor__"])
            saved = json.loads(str(z["__config__"])) if "__config__" in z.files else None
        if saved is not None:
            # self-describing file: the arguments of the call are ignored, as with keras load_model (:17-19)
            dims = tuple(saved.get("input_dims") or (None, None))
            if arch == "rrdb":
                model = build_enhanced_resnet(upsample_factor=sf, num_filters=saved["num_filters"],
                                              num_rrdb_blocks=saved["num_rrdb_blocks"],
                                              num_dense_blocks=saved["num_dense_blocks"], num_convs=saved["num_convs"],
                                              residual_scaling_factor=saved["residual_scaling_factor"], input_dims=dims)
            elif arch == "srresnet":
                model = build_resnet(upsample_factor=sf, num_filters=saved["num_filters"],
                                     num_res_blocks=saved["num_res_blocks"], momentum=saved.get("momentum", 0.8),
                                     input_dims=dims, batch_normalization=saved["batch_norm"])
            else:
                raise ValueError("architecture not recognized")
        elif arch == "rrdb":
            model = build_enhanced_resnet(upsample_factor=sf, num_filters=num_filters, num_rrdb_blocks=num_blocks,
                                          num_dense_blocks=num_dense_blocks, num_convs=num_convs,
                                          kernel_size=kernel_size, residual_scaling_factor=residual_scaling,
                                          input_dims=input_dims)
        elif arch == "srresnet":
            model = build_resnet(upsample_factor=sf, num_filters=num_filters, num_res_blocks=num_blocks,
                                 input_dims=input_dims, batch_normalization=batch_norm)
        else:
            raise ValueError("architecture not recognized")
        model.load_weights(pretrained_model_path)
        return model
    if type(architecture) is str and architecture == "rrdb":
        return build_enhanced_resnet(upsample_factor=upsample_factor, num_filters=num_filters,
                                     num_rrdb_blocks=num_blocks, num_dense_blocks=num_dense_blocks,
                                     num_convs=num_convs, kernel_size=kernel_size,
                                     residual_scaling_factor=residual_scaling, input_dims=input_dims)
    elif type(architecture) is str and architecture == "srresnet":
        return build_resnet(upsample_factor=upsample_factor, num_filters=num_filters, num_res_blocks=num_blocks,
                            input_dims=input_dims, batch_normalization=batch_norm)          # :29-34
    elif callable(architecture):
        return architecture()
    else:
        raise ValueError("architecture not recognized")
