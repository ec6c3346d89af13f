"""Host-side mirror of the training driver ``simple_sr/models/sr_model.py`` for the path that is built: the
``"resnet"`` model type (generator only, pixel losses), i.e. ``SRModel.train_step`` (:403-453) without the GAN branch.

Only the step itself is mirrored (``train_step``, ``generator()``, ``generator_optimizer().iterations``): epoch
bookkeeping, TensorBoard, checkpoints and early stopping stay in the reference's Python (out of scope).
"""
from .training import SRResNetTrainer


class _OptimizerView:
    def __init__(self, trainer):
        self._t = trainer

    @property
    def iterations(self):
        return self._t.iterations

    @property
    def learning_rate(self):
        return self._t.lr


class SRModel:
    """``SRModel(generator_model, ...)``: ``train_step(lr_batch, hr_batch)`` runs one iteration entirely on the device
    and returns the batch metrics the reference accumulates (``generator_loss``, per-loss values, ``psnr``)."""

    def __init__(self, generator_model, loss=("mse", 1.0), learning_rate=1e-3, beta_1=0.9, beta_2=0.999, epsilon=1e-7,
                 allreduce=None):
        self._generator = generator_model
        self._model_type = "resnet"
        self._trainer = SRResNetTrainer(generator_model, loss=loss, learning_rate=learning_rate, beta_1=beta_1,
                                        beta_2=beta_2, epsilon=epsilon, allreduce=allreduce)

    def generator(self):
        return self._generator

    def generator_optimizer(self):
        return _OptimizerView(self._trainer)

    def train_step(self, lr_batch, hr_batch):
        m = self._trainer.train_step(lr_batch, hr_batch)
        return {"generator_loss": m["loss"], "mean_squared_error": m["mse"], "mean_absolute_error": m["mae"],
                "psnr": m["psnr"]}
