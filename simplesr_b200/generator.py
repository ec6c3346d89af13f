"""Host-side mirrors of ``simple_sr/models/generator.py`` (``Generator``), ``simple_sr/models/discriminator.py``
(``Discriminator``) and the pixel / adversarial loss functors of ``simple_sr/utils/models/loss_functions``.

Same constructor arguments, presets and protocol as the reference, so code written against it (``run_training``,
``Generator.from_yaml``-style construction, custom loss lambdas) finds the same names; the arithmetic behind them
runs on the B200 through the C ABI.  Inside ``SRModel.train_step`` the functors are not called one by one: the
trainer reads their *specification* (kind, weight) and emits the whole step as one device graph; called directly
(``calculate_train_loss`` / ``calculate_validation_loss``, generator.py:202-257) they evaluate on the device and
return the scalar, updating the metric dictionaries exactly like the reference's functors.
"""
import numpy as np

from . import _lib as L
from . import model_builder


class Mean:
    """``tf.keras.metrics.Mean`` stand-in: ``m(value)`` accumulates, ``result()`` is the running mean."""

    def __init__(self):
        self.reset_states()

    def __call__(self, value):
        self.total += float(np.mean(value))
        self.count += 1

    def result(self):
        return self.total / self.count if self.count else 0.0

    def reset_states(self):
        self.total, self.count = 0.0, 0


def _track(functor, value, weighted_value, batch_metrics, epoch_metrics):
    if functor.track_metrics and batch_metrics is not None:
        batch_metrics[functor.name](value)
        epoch_metrics[functor.name](value)
        if functor.weighted:
            batch_metrics[f"weighted_{functor.name}"](weighted_value)
            epoch_metrics[f"weighted_{functor.name}"](weighted_value)


def pixel_metrics(hr_batch, sr_batch, max_val=2.0):
    """(mse, mae, psnr per image) of two host batches, evaluated by ssr_pixel_loss on the device."""
    hr = np.ascontiguousarray(hr_batch, np.float32)
    sr = np.ascontiguousarray(sr_batch, np.float32)
    if hr.shape != sr.shape:
        raise ValueError("hr and sr batches must have equal shapes")
    n = hr.shape[0]
    d_hr, d_sr = L.DeviceBuffer.from_numpy(hr), L.DeviceBuffer.from_numpy(sr)
    ws = L.DeviceBuffer(L.load().ssr_pixel_loss_workspace_bytes(n))
    out = L.DeviceBuffer((2 + n) * 4)
    L.pixel_loss(d_hr, d_sr, n, hr.size // n, 0.0, 0.0, max_val, None, ws, out)
    vals = out.download((2 + n,), np.float32)
    for b in (d_hr, d_sr, ws, out):
        b.free()
    return float(vals[0]), float(vals[1]), vals[2:].copy()


class _PixelLoss:
    """mean_squared_error.py:40-66 / mean_absolute_error.py:40-66: Keras loss object (global mean) times loss_weight."""

    kind = None

    def __init__(self, weighted=False, loss_weight=1.0, track_metrics=True):
        self.weighted, self.loss_weight, self.track_metrics = bool(weighted), float(loss_weight), track_metrics
        self.loss = self.weighted_loss = 0.0

    def __call__(self, hr_batch, sr_batch, hr_critic=None, sr_critic=None, batch_metrics=None, epoch_metrics=None):
        mse, mae, _ = pixel_metrics(hr_batch, sr_batch)
        self.loss = mse if self.kind == "mse" else mae
        self.weighted_loss = self.loss * self.loss_weight
        _track(self, self.loss, self.weighted_loss, batch_metrics, epoch_metrics)
        return self.weighted_loss


class MeanSquaredError(_PixelLoss):
    kind, name = "mse", "mean_squared_error"


class MeanAbsoluteError(_PixelLoss):
    kind, name = "mae", "mean_absolute_error"


class RaAdversarialLoss:
    """ra_adversarial_loss.py:42-78 - the generator's relativistic-average term.  Inside a training step it is realised
    by :class:`simplesr_b200.discriminator.RaGANLoss` (which also owns the discriminator's update); called directly with
    critics it evaluates the loss value on the device (ssr_ragan_losses)."""

    name = "ra_adversarial_loss"

    def __init__(self, weighted=False, loss_weight=1.0, track_metrics=True):
        self.weighted, self.loss_weight, self.track_metrics = bool(weighted), float(loss_weight), track_metrics
        self.total_loss = self.weighted_loss = 0.0

    def __call__(self, hr_batch, sr_batch, hr_critic, sr_critic, batch_metrics=None, epoch_metrics=None):
        hc = np.ascontiguousarray(hr_critic, np.float32).ravel()
        sc = np.ascontiguousarray(sr_critic, np.float32).ravel()
        n = hc.size
        bufs = [L.DeviceBuffer.from_numpy(hc), L.DeviceBuffer.from_numpy(sc)] + [L.DeviceBuffer(max(8, n * 4)) for _ in range(4)]
        L.ragan_losses(bufs[0], bufs[1], n, 1.0, 0.0, bufs[2], bufs[3], bufs[4], bufs[5])
        self.total_loss = float(bufs[2].download((2,), np.float32)[0])
        for b in bufs:
            b.free()
        self.weighted_loss = self.total_loss * self.loss_weight
        _track(self, self.total_loss, self.weighted_loss, batch_metrics, epoch_metrics)
        return self.weighted_loss


class RaDiscriminatorLoss:
    """ra_discriminator_loss.py:42-74 (marker + weight; evaluated inside RaGANLoss)."""

    name = "ra_discriminator_loss"

    def __init__(self, weighted=False, loss_weight=1.0, track_metrics=True):
        self.weighted, self.loss_weight, self.track_metrics = bool(weighted), float(loss_weight), track_metrics


class AdversarialLoss(RaAdversarialLoss):
    """adversarial_loss.py:26-66 - the standard-GAN generator term BCE(1, sigmoid critic of SR).  Inside a training step
    it is realised by :class:`simplesr_b200.discriminator.GANLoss`; called directly it takes the critic PROBABILITIES
    (the reference's Dense(1, sigmoid) output) like the reference's functor."""

    name = "adversarial_loss"

    def __call__(self, hr_batch, sr_batch, hr_critic, sr_critic, batch_metrics=None, epoch_metrics=None):
        p = np.clip(np.asarray(sr_critic, np.float64).ravel(), 1e-7, 1.0 - 1e-7)     # Keras BinaryCrossentropy
        self.total_loss = float(np.mean(-np.log(p + 1e-7)))
        self.weighted_loss = self.total_loss * self.loss_weight
        _track(self, self.total_loss, self.weighted_loss, batch_metrics, epoch_metrics)
        return self.weighted_loss


class DiscriminatorLoss(RaDiscriminatorLoss):
    """discriminator_loss.py:26-72 (marker + weight; evaluated inside GANLoss)."""

    name = "discriminator_loss"


def load_yaml(config_yaml):
    """utils/config/yaml_helper.py:73-80: a path is read (PyYAML here, ruamel.yaml there; ``!!python/tuple`` as in
    examples/training/minimal_example.yaml is understood), an already loaded dict is returned as is."""
    if isinstance(config_yaml, dict):
        return config_yaml
    import yaml
    with open(config_yaml) as f:
        return yaml.load(f, Loader=yaml.FullLoader)


def init_loss_functions_from_yaml(config_yaml):
    """utils/config/yaml_helper.py:44-51: ``loss_functions: [{loss_function: <class name>, <ctor kwargs>...}]`` ->
    functor objects.  The names are the reference's classes (yaml_helper.py:4-10)."""
    from .vgg import VGGLoss
    known = {c.__name__: c for c in (MeanAbsoluteError, MeanSquaredError, AdversarialLoss, RaAdversarialLoss,
                                     DiscriminatorLoss, RaDiscriminatorLoss, VGGLoss)}
    out = []
    for spec in config_yaml["loss_functions"]:
        name = spec["loss_function"]
        if name not in known:
            raise AttributeError(f"module 'yaml_helper' has no attribute {name!r}")     # what getattr raises there
        out.append(known[name](**{k: v for k, v in spec.items() if k != "loss_function"}))
    return out


class Generator:
    """generator.py:17-137 - same constructor, presets and methods."""

    def __init__(self, upsample_factor, architecture, loss_functions, num_blocks=16, num_dense_blocks=3, num_filters=64,
                 num_convs=4, kernel_size=3, residual_scaling=0.2, kernel_initializer=None, batch_norm=False,
                 input_dims=(None, None), pretrained_model_path=None, pretrained_model=None):
        self._architecture, self._upsample_factor = architecture, upsample_factor
        if loss_functions is None or (isinstance(loss_functions, list) and len(loss_functions) == 0):
            raise ValueError("no loss function for generator supplied")                     # generator.py:82-83
        if not isinstance(loss_functions, list):
            loss_functions = [loss_functions]
        self._loss_functions = loss_functions
        self._batch_metrics, self._epoch_metrics_train, self._epoch_metrics_valid = {}, {}, {}
        for idx, f in enumerate(self._loss_functions):                                     # generator.py:91-106
            name = getattr(f, "name", f"loss_function_{idx}")
            for d in (self._batch_metrics, self._epoch_metrics_train, self._epoch_metrics_valid):
                d[name] = Mean()
                if getattr(f, "weighted", False):
                    d[f"weighted_{name}"] = Mean()
        for d in (self._batch_metrics, self._epoch_metrics_train, self._epoch_metrics_valid):
            d["generator_loss"] = Mean()
        if pretrained_model is not None:
            self._model = pretrained_model                                                  # generator.py:124-126
        else:
            self._model = model_builder.build_or_load_generator_model(
                upsample_factor=upsample_factor, architecture=architecture, num_blocks=num_blocks,
                num_dense_blocks=num_dense_blocks, num_filters=num_filters, num_convs=num_convs, kernel_size=kernel_size,
                residual_scaling=residual_scaling, kernel_initializer=kernel_initializer, batch_norm=batch_norm,
                input_dims=input_dims, pretrained_model_path=pretrained_model_path)

    def model(self):
        return self._model

    def set_model(self, model):
        self._model = model

    def loss_functions(self):
        return self._loss_functions

    def batch_metrics(self):
        return self._batch_metrics

    def epoch_metrics(self, train=True):
        return self._epoch_metrics_train if train else self._epoch_metrics_valid

    def reset_epoch_metrics(self):
        for m in list(self._epoch_metrics_train.values()) + list(self._epoch_metrics_valid.values()):
            m.reset_states()

    def reset_batch_metrics(self):
        for m in self._batch_metrics.values():
            m.reset_states()

    def formatted_epoch_metrics(self, train=True):
        metrics = self.epoch_metrics(train)
        info = f"\ttotal loss: {metrics['generator_loss'].result():.5f}\n"
        for name, m in metrics.items():
            if name != "generator_loss":
                info += f"\t{name}: {m.result():.5f}\n"
        return info

    def generate(self, lr_batch, training=True):
        return self._model(lr_batch, training=training)                                    # generator.py:200

    def _total(self, sr_batch, hr_batch, sr_critic, hr_critic, epoch_metrics):
        total = 0
        for f in self._loss_functions:
            total += f(hr_batch, sr_batch, hr_critic, sr_critic, self._batch_metrics, epoch_metrics)
        self._batch_metrics["generator_loss"](total)
        epoch_metrics["generator_loss"](total)
        return total

    def calculate_train_loss(self, sr_batch, hr_batch, sr_critic, hr_critic):
        return self._total(sr_batch, hr_batch, sr_critic, hr_critic, self._epoch_metrics_train)   # generator.py:220-228

    def calculate_validation_loss(self, sr_batch, hr_batch, sr_critic, hr_critic):
        return self._total(sr_batch, hr_batch, sr_critic, hr_critic, self._epoch_metrics_valid)

    # ---- presets (generator.py:279-450) --------------------------------------------------------------------------
    @staticmethod
    def srresnet(upsample_factor, loss_function=None, num_blocks=16, num_filters=64, kernel_size=3, batch_norm=True,
                 input_dims=(None, None), pretrained_model_path=None, pretrained_model=None):
        if loss_function is None:
            loss_function = [MeanSquaredError(weighted=False, loss_weight=1.0)]
        return Generator(upsample_factor=upsample_factor, architecture="srresnet", loss_functions=loss_function,
                         num_blocks=num_blocks, num_filters=num_filters, kernel_size=kernel_size, batch_norm=batch_norm,
                         input_dims=input_dims, pretrained_model_path=pretrained_model_path,
                         pretrained_model=pretrained_model)

    @staticmethod
    def rrdb(upsample_factor, loss_functions=MeanAbsoluteError, loss_weight=1.0, num_blocks=16, num_dense_blocks=3,
             num_filters=64, num_convs=4, kernel_size=3, residual_scaling=0.2, kernel_initializer=None, batch_norm=False,
             input_dims=(None, None), pretrained_model_path=None, pretrained_model=None):
        weighted = loss_weight != 1.0
        return Generator(upsample_factor=upsample_factor, architecture="rrdb",
                         loss_functions=[loss_functions(weighted=weighted, loss_weight=loss_weight)], num_blocks=num_blocks,
                         num_dense_blocks=num_dense_blocks, num_filters=num_filters, num_convs=num_convs,
                         kernel_size=kernel_size, residual_scaling=residual_scaling, kernel_initializer=kernel_initializer,
                         batch_norm=batch_norm, input_dims=input_dims, pretrained_model_path=pretrained_model_path,
                         pretrained_model=pretrained_model)

    @staticmethod
    def esrgan_generator(upsample_factor, vgg_layer="block5_conv4", vgg_feature_scaling=1.0, vgg_loss_weight=1.0,
                         adversarial_loss_weight=5e-3, l1_loss_weight=1e-2, num_blocks=16, num_dense_blocks=3,
                         num_filters=64, num_convs=4, kernel_size=3, input_dims=(None, None), pretrained_model_path=None,
                         pretrained_model=None, vgg=None):
        from .vgg import VGGLoss
        return Generator(
            upsample_factor=upsample_factor, architecture="rrdb",
            loss_functions=[MeanAbsoluteError(weighted=True, loss_weight=l1_loss_weight),
                            RaAdversarialLoss(weighted=True, loss_weight=adversarial_loss_weight),
                            VGGLoss(output_layers=vgg_layer, feature_scale=vgg_feature_scaling,
                                    loss_weight=vgg_loss_weight, after_activation=False, vgg=vgg)],
            num_blocks=num_blocks, num_dense_blocks=num_dense_blocks, num_filters=num_filters, num_convs=num_convs,
            kernel_size=kernel_size, input_dims=input_dims, pretrained_model_path=pretrained_model_path,
            pretrained_model=pretrained_model)


    @staticmethod
    def srgan_generator(upsample_factor, vgg_loss, vgg_layer, vgg_feature_scaling=(1 / 12.75), vgg_loss_weight=1.0,
                        adversarial_loss_weight=1e-3, num_blocks=16, num_filters=64, kernel_size=3, batch_norm=True,
                        input_dims=(None, None), pretrained_model_path=None, pretrained_model=None, vgg=None):
        """generator.py:357-403: SRResNet in adversarial mode - VGG loss (post-activation features) or MSE, plus the
        standard adversarial loss."""
        from .vgg import VGGLoss
        if vgg_loss:
            loss_functions = [VGGLoss(vgg_layer, feature_scale=vgg_feature_scaling, loss_weight=vgg_loss_weight,
                                      after_activation=True, vgg=vgg)]
        else:
            loss_functions = [MeanSquaredError(weighted=False, loss_weight=1.0)]
        if adversarial_loss_weight != 1.0:
            loss_functions.append(AdversarialLoss(weighted=True, loss_weight=adversarial_loss_weight))
        else:
            loss_functions.append(AdversarialLoss(weighted=False, loss_weight=1.0))
        return Generator(upsample_factor=upsample_factor, architecture="srresnet", loss_functions=loss_functions,
                         num_blocks=num_blocks, num_filters=num_filters, kernel_size=kernel_size, batch_norm=batch_norm,
                         input_dims=input_dims, pretrained_model_path=pretrained_model_path,
                         pretrained_model=pretrained_model)

    @staticmethod
    def from_yaml(config_yaml):
        """generator.py:452-472: ``model.generator`` of a training YAML (a path or the loaded dict) -> Generator."""
        conf = dict(load_yaml(config_yaml)["model"]["generator"])
        conf["loss_functions"] = init_loss_functions_from_yaml(conf)
        if isinstance(conf.get("input_dims"), list):
            conf["input_dims"] = tuple(conf["input_dims"])
        return Generator(**conf)


class Discriminator:
    """discriminator.py:17-110 - the relativistic critic with its loss functor and label-smoothing settings."""

    def __init__(self, loss_function, relativistic, label_smoothing=False, smoothing_offset=0.3, num_filters=64, alpha=0.2,
                 kernel_size=3, momentum=0.8, initializer=None, input_dims=(None, None), seed=None, device=0):
        from .discriminator import build_discriminator
        self._model = build_discriminator(input_dims=input_dims, num_filters=num_filters, alpha=alpha,
                                          kernel_size=kernel_size, momentum=momentum, relativistic=relativistic,
                                          initializer=initializer, seed=seed, device=device)
        self._relativistic = relativistic
        self._label_smoothing = label_smoothing
        self._smoothing_offset = smoothing_offset if label_smoothing else 0.0          # discriminator.py:68-70
        self._loss_function = loss_function
        self._label_rng = np.random.default_rng(seed)
        self._batch_metrics, self._epoch_metrics_train, self._epoch_metrics_valid = {}, {}, {}
        for d in (self._batch_metrics, self._epoch_metrics_train, self._epoch_metrics_valid):
            d[loss_function.name] = Mean()
            if loss_function.weighted:
                d[f"weighted_{loss_function.name}"] = Mean()

    @staticmethod
    def initialize_relativistic(weighted_loss=False, loss_weight=1.0, num_filters=64, alpha=0.2, kernel_size=3,
                                momentum=0.8, initializer=None, input_dims=(None, None), label_smoothing=False,
                                smoothing_offset=0.3, seed=None, device=0):
        """discriminator.py:264-303 (plus the label-smoothing switches of the constructor, and seed / device)."""
        return Discriminator(loss_function=RaDiscriminatorLoss(weighted=weighted_loss, loss_weight=loss_weight),
                             relativistic=True, label_smoothing=label_smoothing, smoothing_offset=smoothing_offset,
                             num_filters=num_filters, alpha=alpha, kernel_size=kernel_size, momentum=momentum,
                             initializer=initializer, input_dims=input_dims, seed=seed, device=device)

    @staticmethod
    def initialize_standard(weighted_loss=False, loss_weight=1.0, label_smoothing=False, smoothing_offset=0.3,
                            num_filters=64, alpha=0.2, kernel_size=3, momentum=0.8, initializer=None,
                            input_dims=(None, None), seed=None, device=0):
        """discriminator.py:306-361: the sigmoid critic with DiscriminatorLoss."""
        return Discriminator(loss_function=DiscriminatorLoss(weighted=weighted_loss, loss_weight=loss_weight),
                             relativistic=False, label_smoothing=label_smoothing, smoothing_offset=smoothing_offset,
                             num_filters=num_filters, alpha=alpha, kernel_size=kernel_size, momentum=momentum,
                             initializer=initializer, input_dims=input_dims, seed=seed, device=device)

    @staticmethod
    def from_yaml(config_yaml):
        """discriminator.py:363-383: ``model.discriminator`` of a training YAML -> Discriminator.  The YAML lists
        ``loss_functions`` like the generator's section; the critic has one (the constructor's ``loss_function``)."""
        conf = dict(load_yaml(config_yaml)["model"]["discriminator"])
        funcs = init_loss_functions_from_yaml(conf) if "loss_functions" in conf else [conf.pop("loss_function")]
        conf.pop("loss_functions", None)
        if len(funcs) != 1:
            raise ValueError("the discriminator takes exactly one loss function")
        if isinstance(conf.get("input_dims"), list):
            conf["input_dims"] = tuple(conf["input_dims"])
        conf.setdefault("relativistic", type(funcs[0]) is RaDiscriminatorLoss)
        return Discriminator(loss_function=funcs[0], **conf)

    def _get_labels(self, sr_critic, hr_critic):
        """discriminator.py:236-254: the target labels of one critic step (inside ``SRModel.train_step`` the same rule
        feeds the device-side losses, ``discriminator.RaGANLoss.pre_step``)."""
        from .discriminator import smoothed_labels
        return smoothed_labels(self._label_rng, np.shape(sr_critic), np.shape(hr_critic), self._label_smoothing,
                               self._smoothing_offset)

    def model(self):
        return self._model

    def loss_function(self):
        return self._loss_function

    def batch_metrics(self):
        return self._batch_metrics

    def epoch_metrics(self, train=True):
        return self._epoch_metrics_train if train else self._epoch_metrics_valid

    def reset_epoch_metrics(self):
        for m in list(self._epoch_metrics_train.values()) + list(self._epoch_metrics_valid.values()):
            m.reset_states()

    def reset_batch_metrics(self):
        for m in self._batch_metrics.values():
            m.reset_states()

    def formatted_epoch_metrics(self, train=True):
        metrics = self.epoch_metrics(train)
        name = self._loss_function.name
        info = f"\t{name}: {metrics[name].result():.5f}\n"
        for k, m in metrics.items():
            if k != name:
                info += f"\t{k}: {m.result():.5f}\n"
        return info
