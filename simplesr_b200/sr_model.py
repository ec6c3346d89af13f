"""Host-side mirror of the training driver ``simple_sr/models/sr_model.py``: ``SRModel`` for both model types
("resnet": generator + pixel / perceptual losses; "gan": + relativistic discriminator), with the protocol
``run_training`` drives (simple_sr/operations/training.py:36-112): ``generator_optimizer().iterations`` /
``_decayed_lr``, ``stop_early``, ``before_epoch``, ``train_step``, ``after_train_batch``, ``validation_step``,
``after_validation_batch``, ``formatted_epoch_metrics``, ``after_epoch``, ``after_training``, ``epoch_history``,
``batch_history``, ``name``, ``generator()``, plus ``save_model`` (:233-244) and checkpoint save / restore
(:172-192: step, tracked metric, generator + optimizer slots, discriminator + optimizer slots).

What differs from the reference on purpose: ``train_step`` is ONE captured device graph (forward, losses, backward,
both Adam updates, data-parallel exchange) instead of two GradientTapes; TensorBoard writers, plotting
(``test_and_plot``) stay in the reference's Python (out of scope) and are accepted but ignored here.
"""
import json
import os

import numpy as np

from . import metrics as image_metric_fns
from .generator import (Discriminator, Generator, Mean, MeanAbsoluteError, MeanSquaredError, RaAdversarialLoss,
                        pixel_metrics)
from .training import PiecewiseConstantDecay, RRDBTrainer, SRResNetTrainer


class Adam:
    """Configuration stand-in for ``tf.keras.optimizers.Adam`` (the only optimizer the reference's recipes use):
    ``learning_rate`` a float or a :class:`PiecewiseConstantDecay`."""

    def __init__(self, learning_rate=1e-3, beta_1=0.9, beta_2=0.999, epsilon=1e-7):
        self.learning_rate, self.beta_1, self.beta_2, self.epsilon = learning_rate, beta_1, beta_2, epsilon

    @classmethod
    def from_config(cls, config):
        """Keras optimizer config (sr_model.py:121-131 passes ``*_optimizer_config`` to ``optimizer.from_config``): the
        learning rate is a number or a serialised schedule ``{"class_name": "PiecewiseConstantDecay", "config":
        {"boundaries": [...], "values": [...]}}`` as in the reference's tests/models/test_learnrate_scheduling.py."""
        kw = {k: config[k] for k in ("learning_rate", "beta_1", "beta_2", "epsilon") if k in config}
        lr = kw.get("learning_rate")
        if isinstance(lr, dict):
            name, inner = lr.get("class_name", "PiecewiseConstantDecay"), lr.get("config", lr)
            if name != "PiecewiseConstantDecay":
                raise ValueError(f"learning-rate schedule {name!r} is not supported (PiecewiseConstantDecay only)")
            kw["learning_rate"] = PiecewiseConstantDecay(list(inner["boundaries"]), list(inner["values"]))
        return cls(**kw)

    def get_config(self):
        lr = self.learning_rate
        if isinstance(lr, PiecewiseConstantDecay):
            lr = {"boundaries": lr.boundaries, "values": lr.values}
        return dict(name="Adam", learning_rate=lr, beta_1=self.beta_1, beta_2=self.beta_2, epsilon=self.epsilon)


class _Scalar:
    def __init__(self, v):
        self._v = v

    def numpy(self):
        return self._v


class _OptimizerView:
    """What ``run_training`` reads from the optimizer (training.py:67-69): ``iterations.numpy()``, ``_decayed_lr()``."""

    def __init__(self, trainer_like, config):
        self._t, self._config = trainer_like, config

    @property
    def iterations(self):
        return _Scalar(self._t.iterations)

    def _decayed_lr(self, _dtype=None):
        return _Scalar(np.float32(self._t.opt.learning_rate_at(self._t.iterations)))

    @property
    def learning_rate(self):
        return self._t.opt.learning_rate_at(self._t.iterations)

    def _get_hyper(self, name):
        """Keras ``OptimizerV2._get_hyper``: the schedule object (``.boundaries`` / ``.values``) for a scheduled learning
        rate, otherwise a scalar with ``.numpy()``."""
        v = getattr(self._config, name)
        return v if isinstance(v, PiecewiseConstantDecay) else _Scalar(np.float32(v))

    def get_config(self):
        return self._config.get_config()


class EarlyStopping:
    """utils/models/early_stopping.py:3-37, same rule."""

    def __init__(self, metric_key, patience):
        self._metric_key, self._patience = metric_key, patience
        self._epochs_without_improvement = self._num_epochs_after_best = 0
        self._early_stop = False
        self._current_best_val = float("-inf")

    def evaluate_stop_criterion(self, metric_history):
        this = metric_history[-1]
        last = metric_history[-2] if len(metric_history) > 1 else float("-inf")
        if this > self._current_best_val:
            self._epochs_without_improvement = self._num_epochs_after_best = 0
            self._current_best_val = this
        else:
            self._num_epochs_after_best += 1
            if this < last:
                self._epochs_without_improvement += 1
        if self._epochs_without_improvement >= self._patience:
            self._early_stop = True

    def stop_early(self):
        return self._early_stop

    def num_epochs_after_best(self):
        return self._num_epochs_after_best


def _make_optimizer(opt, config):
    if opt is None:
        return None
    if config is not None and hasattr(opt, "from_config"):
        return opt.from_config(config)
    return opt() if callable(opt) and not isinstance(opt, Adam) else opt


class SRModel:
    """``SRModel(model_type, generator, generator_optimizer, ...)`` - sr_model.py:67-213."""

    def __init__(self, model_type, generator, generator_optimizer=None, generator_optimizer_config=None,
                 discriminator=None, discriminator_optimizer=None, discriminator_optimizer_config=None,
                 image_metrics=None, early_stop_metric="psnr", early_stop_patience=100, epoch_train_summary_writer=None,
                 batch_train_summary_writer=None, epoch_validation_summary_writer=None,
                 batch_validation_summary_writer=None, resnet_checkpoint=None, config=None, comm=None, allreduce=None,
                 metric_lag=0):
        if not isinstance(model_type, str) or model_type.lower() not in ["gan", "resnet"]:
            raise ValueError("model type not recognized")                                        # :84-85
        if generator is None:
            raise ValueError("no generator was supplied")
        if generator_optimizer is None and resnet_checkpoint is None:
            raise ValueError("no generator optimizer was supplied")
        model_type = model_type.lower()
        if model_type == "gan" and discriminator is None:
            raise ValueError("model type is GAN but no discriminator supplied")
        if model_type == "gan" and discriminator_optimizer is None:
            raise ValueError("model type is GAN but no discriminator optimizer supplied")
        if model_type == "resnet" and discriminator is not None:
            raise ValueError("model type is Resnet but discriminator was supplied")              # :94-95
        if not isinstance(generator, Generator):
            raise ValueError("generator must be a simplesr_b200.generator.Generator")
        self._model_type = self.name = model_type
        self._generator, self._discriminator = generator, discriminator
        self._epochs = self._iterations = 0
        self._metric_lag = int(metric_lag)
        self._model_dir, self._checkpoint_dir = "./models", "./checkpoints"
        self._config = config
        if config is not None:
            self._model_dir = getattr(config, "model_dir", None) or self._model_dir
            self._checkpoint_dir = getattr(config, "checkpoint_dir", None) or self._checkpoint_dir
        g_opt = _make_optimizer(generator_optimizer, generator_optimizer_config) or Adam()
        d_opt = _make_optimizer(discriminator_optimizer, discriminator_optimizer_config)
        self._generator_optimizer_config, self._discriminator_optimizer_config = g_opt, d_opt

        # ---- the loss functors of the Generator become ONE device step (generator.py:220-228 sums them)
        from .vgg import VGGLoss
        pixel, extra, self._ragan = [], [], None
        for f in generator.loss_functions():
            if isinstance(f, (MeanSquaredError, MeanAbsoluteError)):
                pixel.append((f.kind, f.loss_weight))
            elif isinstance(f, VGGLoss):
                extra.append(f)
            elif isinstance(f, RaAdversarialLoss):
                if model_type != "gan":
                    raise ValueError(f"{type(f).__name__} needs model_type 'gan' and a discriminator")
                from .discriminator import RaGANLoss
                relativistic = f.name == "ra_adversarial_loss"
                if relativistic != bool(discriminator.model().relativistic):
                    raise ValueError(f"{type(f).__name__} does not match the discriminator (relativistic="
                                     f"{discriminator.model().relativistic})")
                self._ragan = RaGANLoss(discriminator.model(), loss_weight=f.loss_weight, learning_rate=d_opt.learning_rate,
                                        beta_1=d_opt.beta_1, beta_2=d_opt.beta_2, epsilon=d_opt.epsilon,
                                        label_smoothing=discriminator._label_smoothing,
                                        smoothing_offset=discriminator._smoothing_offset or 0.3,
                                        relativistic=relativistic)
                extra.append(self._ragan)
            else:
                raise ValueError(f"loss functor {getattr(f, 'name', f)!r} has no device implementation: supported are "
                                 "MeanSquaredError, MeanAbsoluteError, VGGLoss, RaAdversarialLoss")
        if model_type == "gan" and self._ragan is None:
            raise ValueError("model type is GAN but the generator has no RaAdversarialLoss")
        model = generator.model()
        cls = {"srresnet": SRResNetTrainer, "rrdb": RRDBTrainer}.get(getattr(model, "architecture", None))
        if cls is None:
            raise ValueError("generator model must come from simplesr_b200.model_builder (srresnet or rrdb)")
        self._trainer = cls(model, loss=pixel or [("mae", 0.0)], learning_rate=g_opt.learning_rate, beta_1=g_opt.beta_1,
                            beta_2=g_opt.beta_2, epsilon=g_opt.epsilon, extra_losses=extra, comm=comm, allreduce=allreduce)

        # ---- metrics (sr_model.py:194-208)
        # {name: func(hr_batch, sr_batch)} like the reference (default dict(psnr=metrics.psnr), sr_model.py:198).  A key
        # bound to metrics.psnr (or None) is produced by the fused step on the device; any other callable - psnr_on_y,
        # ssim, a user lambda - is evaluated after the step on the generated batch (one extra device -> host copy of the
        # SR batch per step, as the reference's eager call implies), which also fixes the metric lag to 0
        self._image_metrics = dict(image_metrics) if image_metrics is not None else dict(psnr=image_metric_fns.psnr)
        self._fused_image_metrics = [k for k, f in self._image_metrics.items() if f is None or f is image_metric_fns.psnr]
        self._host_image_metrics = [k for k in self._image_metrics if k not in self._fused_image_metrics]
        if self._host_image_metrics:
            self._metric_lag = 0
        self._train_epoch_metrics = {k: Mean() for k in self._image_metrics}
        self._valid_epoch_metrics = {k: Mean() for k in self._image_metrics}
        self._batch_metrics = {k: Mean() for k in self._image_metrics}
        init = lambda train: {k: [] for k in self._combined_epoch_metrics(train)}
        self._train_batch_history, self._train_epoch_history = init(True), init(True)
        self._valid_batch_history, self._valid_epoch_history = init(False), init(False)
        self._early_stop_metric = early_stop_metric
        self._early_stopping_util = EarlyStopping(early_stop_metric, early_stop_patience)
        self._checkpoint_metric = -1.0
        self._saved_checkpoints = []
        if resnet_checkpoint is not None:
            self.restore_checkpoint(resnet_checkpoint, generator_only=(model_type == "gan"))

    # ---- accessors ------------------------------------------------------------------------------------------------
    def iterations(self):
        return self._iterations

    def generator(self):
        return self._generator.model()

    def discriminator(self):
        return self._discriminator.model() if self._discriminator is not None else None

    def generator_optimizer(self):
        return _OptimizerView(self._trainer, self._generator_optimizer_config)

    def discriminator_optimizer(self):
        if self._ragan is None:
            return None
        return _OptimizerView(self._ragan, self._discriminator_optimizer_config)

    def stop_early(self):
        return self._early_stopping_util.stop_early()

    def _combined_epoch_metrics(self, train=True):
        out = dict(self._train_epoch_metrics if train else self._valid_epoch_metrics)
        out.update(self._generator.epoch_metrics(train))
        if self._model_type == "gan":
            out.update(self._discriminator.epoch_metrics(train))
        return out

    def _combined_batch_metrics(self):
        out = dict(self._batch_metrics)
        out.update(self._generator.batch_metrics())
        if self._model_type == "gan":
            out.update(self._discriminator.batch_metrics())
        return out

    def epoch_metrics(self, train=True):
        return self._combined_epoch_metrics(train)

    def batch_metrics(self):
        return self._combined_batch_metrics()

    def epoch_history(self, train=True):
        return self._train_epoch_history if train else self._valid_epoch_history

    def batch_history(self, train=True):
        return self._train_batch_history if train else self._valid_batch_history

    # ---- the step -------------------------------------------------------------------------------------------------
    def _record(self, m, epoch_g, epoch_img, epoch_d):
        """Feed one step's device metrics into the metric dictionaries the way the functors do (generator.py:220-228,
        ra_discriminator_loss.py:68-73, sr_model.py:630-634)."""
        bg = self._generator.batch_metrics()
        for f in self._generator.loss_functions():
            if isinstance(f, MeanSquaredError):
                raw = m["mse"]
            elif isinstance(f, MeanAbsoluteError):
                raw = m["mae"]
            elif isinstance(f, RaAdversarialLoss):
                raw = m[f.name] / f.loss_weight if f.loss_weight else 0.0
            else:
                raw = m[f.name]      # VGGLoss tracks the value it returns (vgg_loss.py:171-174)
            for d in (bg, epoch_g):
                d[f.name](raw)
                if getattr(f, "weighted", False) and f"weighted_{f.name}" in d:
                    d[f"weighted_{f.name}"](raw * f.loss_weight)
        for d in (bg, epoch_g):
            d["generator_loss"](m["loss"])
        if self._model_type == "gan" and self._ragan is not None and self._ragan.metric_names[1] in m:
            lf, dl = self._discriminator.loss_function(), m[self._ragan.metric_names[1]]
            for d in (self._discriminator.batch_metrics(), epoch_d):
                d[lf.name](dl)
                if lf.weighted:
                    d[f"weighted_{lf.name}"](dl * lf.loss_weight)
        for key in self._fused_image_metrics:                 # tf.image.psnr(hr, sr, max_val=2.0), batch mean (:453)
            epoch_img[key](m["psnr"])
            self._batch_metrics[key](m["psnr"])

    def _update_metrics(self, hr_batch, sr_batch, epoch_metrics, only=None):
        """sr_model.py:630-634: every image metric on (hr, sr), recorded in the epoch and the batch dictionaries."""
        for key, func in self._image_metrics.items():
            if only is not None and key not in only:
                continue
            res = (func or image_metric_fns.psnr)(hr_batch, sr_batch)
            epoch_metrics[key](res)
            self._batch_metrics[key](res)

    def train_step(self, lr_batch, hr_batch):
        """sr_model.py:403-453 as one device graph; returns the step's metrics (the reference returns None)."""
        m = self._trainer.train_step(lr_batch, hr_batch, lag=self._metric_lag)
        if m is None:      # metric_lag = 1, first step: nothing to record yet
            return None
        self._record(m, self._generator.epoch_metrics(True), self._train_epoch_metrics,
                     self._discriminator.epoch_metrics(True) if self._discriminator else None)
        if self._host_image_metrics:
            self._update_metrics(np.ascontiguousarray(hr_batch, np.float32), self._trainer.last_sr(),
                                 self._train_epoch_metrics, only=self._host_image_metrics)
        out = {"generator_loss": m["loss"], "mean_squared_error": m["mse"], "mean_absolute_error": m["mae"],
               "psnr": m["psnr"]}
        out.update({k: v for k, v in m.items() if k not in ("loss", "mse", "mae", "psnr")})
        return out

    def validation_step(self, lr_batch, hr_batch):
        """sr_model.py:455-480: forward with training=False, pixel / perceptual losses and image metrics; no update.
        The adversarial terms are not evaluated here (the critic's inference mode - moving statistics - is not built)."""
        from .vgg import VGGLoss
        sr = self._generator.generate(lr_batch, training=False)
        hr = np.ascontiguousarray(hr_batch, np.float32)
        total = 0.0
        ge = self._generator.epoch_metrics(False)
        for f in self._generator.loss_functions():
            if isinstance(f, (MeanSquaredError, MeanAbsoluteError, VGGLoss)):
                total += f(hr, sr, None, None, self._generator.batch_metrics(), ge)
        self._generator.batch_metrics()["generator_loss"](total)
        ge["generator_loss"](total)
        self._update_metrics(hr, sr, self._valid_epoch_metrics)                                # :480
        psnr = float(np.mean(pixel_metrics(hr, sr)[2]))       # tf.image.psnr(max_val=2.0) per image, on the device
        return {"generator_loss": total, "psnr": psnr}

    def test_and_plot(self, *args, **kwargs):
        """Plotting is out of scope; kept so that ``run_training`` can call it."""

    def after_train_batch(self):
        self._iterations = self._trainer.iterations                                            # :526
        self._update_history(self._combined_batch_metrics(), self._train_batch_history)
        self._reset_batch_metrics()

    def after_validation_batch(self):
        self._update_history(self._combined_batch_metrics(), self._valid_batch_history)
        self._reset_batch_metrics()

    def before_epoch(self):
        self._reset_epoch_metrics()
        self._epochs += 1

    def after_epoch(self):
        """sr_model.py:563-599: save the generator, update the histories, early stopping, checkpoint on a new best."""
        self._trainer.flush()
        self.save_model(self._model_dir)
        self._update_history(self._combined_epoch_metrics(True), self._train_epoch_history)
        self._update_history(self._combined_epoch_metrics(False), self._valid_epoch_history)
        self._checkpoint_metric = self._valid_epoch_metrics[self._early_stop_metric].result()
        self._early_stopping_util.evaluate_stop_criterion(self._valid_epoch_history[self._early_stop_metric])
        if self.stop_early() and self._saved_checkpoints:
            self.restore_checkpoint(self._saved_checkpoints[-1])
        if self._early_stopping_util.num_epochs_after_best() == 0:
            self._saved_checkpoints.append(self.save_checkpoint())
            for old in self._saved_checkpoints[:-5]:                                           # max_to_keep=5 (:191)
                if os.path.exists(old):
                    os.remove(old)
            self._saved_checkpoints = self._saved_checkpoints[-5:]

    def after_training(self):
        if self._saved_checkpoints:
            self.restore_checkpoint(self._saved_checkpoints[-1])
        self.save_model(self._model_dir, postfix="best")
        self._reset_epoch_metrics()

    def formatted_epoch_metrics(self):
        return self._format_metrics(self._train_epoch_metrics, "Training") + \
            self._format_metrics(self._valid_epoch_metrics, "Validation")

    def _format_metrics(self, metrics, header):
        train = header == "Training"
        img = "".join(f"{k}: {metrics[k].result():.5f}\n" for k in self._image_metrics)
        gen = self._generator.formatted_epoch_metrics(train=train)
        if self._model_type == "gan":
            return f"{header}\n{img}Generator\n{gen}Discriminator\n{self._discriminator.formatted_epoch_metrics(train=train)}"
        return f"{header}\n " + "".join(f"{n}: {m.result():.4f}, " for n, m in metrics.items()) + "\n" + gen + "\n"

    def _update_history(self, metrics, history):
        for name, metric in metrics.items():
            history.setdefault(name, []).append(metric.result())

    def _reset_epoch_metrics(self):
        for m in list(self._train_epoch_metrics.values()) + list(self._valid_epoch_metrics.values()):
            m.reset_states()
        self._generator.reset_epoch_metrics()
        if self._model_type == "gan":
            self._discriminator.reset_epoch_metrics()

    def _reset_batch_metrics(self):
        for m in self._batch_metrics.values():
            m.reset_states()
        self._generator.reset_batch_metrics()
        if self._model_type == "gan":
            self._discriminator.reset_batch_metrics()

    # ---- construction from a config object / a YAML file ---------------------------------------------------------
    @staticmethod
    def init(config, generator, generator_optimizer, generator_optimizer_config=None, discriminator=None,
             discriminator_optimizer=None, discriminator_optimizer_config=None, image_metrics=None, **kw):
        """sr_model.py:704-739: model type inferred from the presence of a discriminator; early stopping, save
        directories (and, in the reference, the TensorBoard writers) come from the config object."""
        return SRModel(model_type="resnet" if discriminator is None else "gan", generator=generator,
                       generator_optimizer=generator_optimizer, generator_optimizer_config=generator_optimizer_config,
                       discriminator=discriminator, discriminator_optimizer=discriminator_optimizer,
                       discriminator_optimizer_config=discriminator_optimizer_config, image_metrics=image_metrics,
                       early_stop_metric=getattr(config, "early_stop_metric", "psnr"),
                       early_stop_patience=getattr(config, "early_stop_patience", 100), config=config, **kw)

    @staticmethod
    def from_yaml(config_yaml, config=None, **kw):
        """The ``model:`` section of a training YAML -> SRModel, as ``ConfigUtil.from_yaml`` assembles it
        (utils/config/config_util.py:311-331): ``generator`` / ``discriminator`` through their ``from_yaml``,
        ``generator_optimizer`` / ``discriminator_optimizer`` by name with the optional ``*_optimizer_config``.  The
        ``general:`` section (data pipeline, save paths) stays with the reference's ConfigUtil; its object may be passed
        as ``config``."""
        from .generator import load_yaml
        conf = load_yaml(config_yaml)
        model = conf["model"]

        def optimizer(key):
            name = model[key]
            if name != "Adam":
                raise ValueError(f"optimizer {name!r} is not supported by the device step (Adam only, the optimizer of "
                                 f"every recipe in the reference's examples)")
            return Adam

        generator = Generator.from_yaml(conf)
        discriminator = d_opt = d_cfg = None
        if "discriminator" in model:
            discriminator = Discriminator.from_yaml(conf)
            d_opt = optimizer("discriminator_optimizer")
            d_cfg = model.get("discriminator_optimizer_config")
        return SRModel.init(config, generator, optimizer("generator_optimizer"), model.get("generator_optimizer_config"),
                            discriminator, d_opt, d_cfg, **kw)

    # ---- persistence ----------------------------------------------------------------------------------------------
    def save_model(self, save_path, postfix=None):
        """sr_model.py:233-244: ``<save_path>/<type>_gen_<postfix>.h5`` in the Keras HDF5 weight layout, the file name
        and format the reference writes (``GeneratorModel.save`` -> keras_h5 / h5lite; no h5py needed)."""
        if postfix is None:
            postfix = self._epochs
        self._trainer.flush()
        return self._generator.model().save(f"{save_path}/{self._model_type}_gen_{postfix}.h5")

    def save_checkpoint(self, path=None):
        """The content of the reference's tf.train.Checkpoint (sr_model.py:172-187): step, tracked metric, generator
        variables + Adam slots + iterations, and the same for the discriminator in GAN mode."""
        tr = self._trainer
        tr.flush()
        if path is None:
            os.makedirs(f"{self._checkpoint_dir}/{self._model_type}", exist_ok=True)
            path = f"{self._checkpoint_dir}/{self._model_type}/ckpt-{self._epochs}.npz"
        dl = lambda opt, b: b.download((opt.count,), np.float32, tr.stream.ptr)
        arrays = dict(step=np.int64(self._epochs), metric=np.float64(self._checkpoint_metric),
                      meta=json.dumps(dict(model_type=self._model_type, architecture=self.generator().architecture)),
                      g_param=dl(tr.opt, tr.d_param), g_m=dl(tr.opt, tr.d_m), g_v=dl(tr.opt, tr.d_v),
                      g_iterations=np.int64(tr.iterations))
        for i, v in enumerate(self.generator().non_trainable_variables):
            arrays[f"g_state_{i:04d}"] = v.numpy()
        if self._ragan is not None:
            r = self._ragan
            arrays.update(d_param=dl(r.opt, r.d_param), d_m=dl(r.opt, r.d_m), d_v=dl(r.opt, r.d_v),
                          d_iterations=np.int64(r.iterations))
        with open(path, "wb") as f:
            np.savez(f, **arrays)
        return path

    def restore_checkpoint(self, path, generator_only=False):
        tr = self._trainer
        tr.flush()
        with np.load(path) as z:
            if z["g_param"].size != tr.opt.count:
                raise ValueError("checkpoint does not match the generator (parameter count differs)")
            s = tr.stream.ptr
            for key, buf in (("g_param", tr.d_param), ("g_m", tr.d_m), ("g_v", tr.d_v)):
                buf.upload(z[key], s)
            tr.iterations = int(z["g_iterations"])
            tr.opt.set_iterations(tr.iterations, s)
            tr._repack(s)
            for i, v in enumerate(self.generator().non_trainable_variables):
                v.assign(z[f"g_state_{i:04d}"])
            for c in self.generator().convs.values():
                c.dirty = True
            if self._ragan is not None and not generator_only and "d_param" in z.files:
                r = self._ragan
                for key, buf in (("d_param", r.d_param), ("d_m", r.d_m), ("d_v", r.d_v)):
                    buf.upload(z[key], None)
                r.iterations = int(z["d_iterations"])
                r.opt.set_iterations(r.iterations, None)
                r._repack(None)
            if not generator_only:
                self._epochs = int(z["step"])
                self._checkpoint_metric = float(z["metric"])
            self._iterations = tr.iterations
        tr.flush()

    def __str__(self):
        d = self._discriminator_optimizer_config.get_config() if self._discriminator_optimizer_config else None
        return (f"# SR Model\nmodel type: {self._model_type}\n"
                f"generator optimizer: {self._generator_optimizer_config.get_config()}\n"
                f"discriminator optimizer: {d}\nimage metrics: {list(self._image_metrics)}\n"
                f"early stop metric: {self._early_stop_metric}\n")


__all__ = ["SRModel", "Adam", "PiecewiseConstantDecay", "EarlyStopping", "Generator", "Discriminator"]
