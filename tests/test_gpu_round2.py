"""Round-2 surface of the hot path on the B200 against the oracle: the optimizer inside the step graph (two steps with
batch norm, learning-rate schedule, resume after the trainer exists), VGGLoss' remaining options (post-activation
features, several layers, total variation), rectangular patches, ``model(x, training=True)`` with batch norm, label
smoothing, the SRModel / Generator facade."""
import os
import tempfile

import numpy as np
import pytest

from tests.helpers import L, O, rel_err
from tests.test_gpu_train_step import _setup, _setup_bn

pytestmark = pytest.mark.gpu


def _oracle_adam_all(params, bn, grads, state, t, lr):
    """One Keras-Adam step of every SRResNet variable (kernel, bias, alpha, gamma, beta) on the host."""
    for name in list(params):
        arrs = list(params[name])
        for i in range(3):
            if arrs[i] is None:
                continue
            key = (name, i)
            m, v = state.get(key, (np.zeros_like(arrs[i]), np.zeros_like(arrs[i])))
            arrs[i], m, v = O.adam_update(arrs[i], grads[name][i], m, v, t, lr=lr)
            state[key] = (m, v)
        params[name] = tuple(arrs)
        if bn is not None and name in bn:
            for j, k in enumerate(("gamma", "beta")):
                key = (name, k)
                m, v = state.get(key, (np.zeros_like(bn[name][k]), np.zeros_like(bn[name][k])))
                bn[name][k], m, v = O.adam_update(bn[name][k], grads[name + "_bn"][j], m, v, t, lr=lr)
                state[key] = (m, v)


def test_two_steps_with_batch_norm_follow_the_oracle():
    """lr > 0, two iterations, batch norm on: the second forward pass must see the weights of the first update in EVERY
    layer (the 'last' conv read the model's inference image in round 1) - loss and gradients of step 2 vs the oracle."""
    from simplesr_b200.training import SRResNetTrainer
    nb, sf = 2, 2
    m, params, bn = _setup_bn(nb, sf)
    rng = np.random.default_rng(3)
    lr = rng.uniform(0, 1, size=(2, 12, 10, 3)).astype(np.float32)
    hr = rng.uniform(-1, 1, size=(2, 24, 20, 3)).astype(np.float32)
    step = 2e-3      # Adam's first updates move every weight by ~step: a stale layer shows up clearly in the loss
    tr = SRResNetTrainer(m, loss=("mse", 1.0), learning_rate=step)
    kw = dict(upsample_factor=sf, num_res_blocks=nb)
    p, b_, state = dict(params), {k: dict(v) for k, v in bn.items()}, {}
    for t in (1, 2):
        out = tr.train_step(lr, hr)
        loss, _, g = O.srresnet_loss_and_grads(p, lr, hr, bn=b_, **kw)
        assert abs(out["loss"] - loss) <= 3e-2 * abs(loss), (t, out["loss"], loss)
        if t == 2:
            # the two trajectories differ wherever a tiny gradient changed sign in step 1 (Adam moves every weight by
            # +-step): the step-2 gradients agree in direction, not element by element
            got = tr.gradients()
            cos = lambda a, b: float((a * b).sum() / (np.linalg.norm(a) * np.linalg.norm(b) + 1e-30))
            for name in ("last", "up0", "trunk"):
                assert cos(got[name][0], g[name][0]) >= 0.9, (name, cos(got[name][0], g[name][0]))
        if t == 2:
            # the test is only worth something if a stale 'last' layer (round 1's bug) would have been seen: on the
            # oracle the step-2 loss with the step-1 'last' weights is 0.32 against 1.00 with the updated ones
            loss_stale, _, _ = O.srresnet_loss_and_grads({**p, "last": params["last"]}, lr, hr, bn=b_, **kw)
            assert abs(loss_stale - loss) > 0.3 * abs(loss), (loss_stale, loss)
        _oracle_adam_all(p, b_, g, state, t, step)
    assert tr.iterations == 2 and tr.opt.iterations() == 2
    tr.release()
    m.release()


def test_learning_rate_schedule_on_the_device():
    """PiecewiseConstantDecay evaluated by ssr_opt_prepare inside the step graph: the weights move by the scheduled
    step size (Adam's first steps move every weight by ~lr)."""
    from simplesr_b200.training import PiecewiseConstantDecay, SRResNetTrainer
    m, params = _setup(1, 2)
    sch = PiecewiseConstantDecay([1], [1e-3, 1e-5])       # iterations 0, 1 -> 1e-3; from iteration 2 -> 1e-5
    tr = SRResNetTrainer(m, loss=("mse", 1.0), learning_rate=sch)
    rng = np.random.default_rng(0)
    lr = rng.uniform(0, 1, size=(2, 8, 8, 3)).astype(np.float32)
    hr = rng.uniform(-1, 1, size=(2, 16, 16, 3)).astype(np.float32)
    k0 = m.convs["last"].kernel.numpy().copy()
    moves = []
    for _ in range(4):
        tr.train_step(lr, hr)
        k1 = m.convs["last"].kernel.numpy().copy()
        moves.append(float(np.abs(k1 - k0).max()))
        k0 = k1
    assert 5e-4 < moves[0] <= 1.05e-3 and 5e-4 < moves[1] <= 2e-3, moves
    assert moves[2] <= 2e-5 and moves[3] <= 2e-5, moves
    assert tr.lr == 1e-5
    tr.release()


def test_resume_after_the_trainer_exists():
    """load_weights / assign AFTER constructing the trainer reach the device masters: training continues
    from the restored values and inference sees them."""
    from simplesr_b200 import model_builder as MB
    from simplesr_b200.training import SRResNetTrainer
    m, params = _setup(1, 2)
    rng = np.random.default_rng(0)
    lr = rng.uniform(0, 1, size=(2, 8, 8, 3)).astype(np.float32)
    hr = rng.uniform(-1, 1, size=(2, 16, 16, 3)).astype(np.float32)
    with tempfile.TemporaryDirectory() as d:
        path = m.save(os.path.join(d, "gen"))
        assert path.endswith("gen.npz") and os.path.exists(path)
        y0 = m(lr, training=False)
        tr = SRResNetTrainer(m, loss=("mse", 1.0), learning_rate=1e-2)
        for _ in range(3):
            tr.train_step(lr, hr)
        assert np.abs(m(lr, training=False) - y0).max() > 1e-3      # training moved the model
        m.load_weights(os.path.join(d, "gen"))                        # no suffix: the same file
        np.testing.assert_array_equal(m(lr, training=False), y0)     # inference sees the restored weights
        out = tr.train_step(lr, hr, use_graph=True)
        loss0, _, _ = O.srresnet_loss_and_grads(params, lr, hr, upsample_factor=2, num_res_blocks=1)
        assert abs(out["loss"] - loss0) <= 2e-3 * abs(loss0)         # and so does the training forward
        # the file describes the model: the loader rebuilds it without being told the block count
        m2 = MB.build_or_load_generator_model(2, "srresnet", 99, 64, 3, 0.2, None, True, (None, None),
                                              pretrained_model_path=os.path.join(d, "gen"))
        assert m2.config["num_res_blocks"] == 1 and m2.config["batch_norm"] is False
        np.testing.assert_array_equal(m2(lr, training=False), y0)
        m2.release()
    tr.release()
    m.release()


@pytest.mark.parametrize("arch", ["srresnet_bn", "rrdb"])
def test_keras_h5_generator_files(arch):
    """The reference's generator files are Keras HDF5 (sr_model.py:244 writes "<type>_gen_<epoch>.h5", model_builder.py:
    17-19 loads them): ``save("x.h5")`` writes that layout, ``build_or_load_generator_model(pretrained_model_path=)``
    rebuilds the graph from the weight shapes alone, ``load_weights`` restores into an existing model."""
    from simplesr_b200 import h5lite, model_builder as MB
    rng = np.random.default_rng(5)
    if arch == "rrdb":
        m = MB.build_enhanced_resnet(upsample_factor=2, num_rrdb_blocks=2, num_dense_blocks=3, seed=3)
    else:
        m = MB.build_resnet(upsample_factor=4, num_res_blocks=2, batch_normalization=True, momentum=0.8, seed=3)
    for v in m.variables:                                   # biases, slopes, BN statistics away from their zero / one init
        base = v.numpy()
        if base.ndim == 1 or "alpha" in v.name:
            lo = 0.5 if v.name.endswith(("gamma:0", "moving_variance:0")) else -0.2
            v.assign(rng.uniform(lo, lo + 1.0 if lo > 0 else 0.2, size=base.shape).astype(np.float32))
    x = rng.uniform(0, 1, size=(1, 12, 10, 3)).astype(np.float32)
    y0 = m(x, training=False)
    with tempfile.TemporaryDirectory() as d:
        path = m.save(os.path.join(d, f"{arch}_gen_7.h5"))
        layers, meta = h5lite.load_keras_weights(path)
        assert meta["backend"] == "tensorflow" and len(layers) == len({v.name.split("/")[0] for v in m.variables})
        m2 = MB.build_or_load_generator_model(8, "callable-ignored", 1, 32, 3, 0.2, None, False, (None, None),
                                              pretrained_model_path=path)
        assert m2.architecture == m.architecture and m2.upsample_factor == m.upsample_factor
        if arch == "rrdb":
            assert m2.config["num_rrdb_blocks"] * m2.config["num_dense_blocks"] == 6
        else:
            assert m2.config["num_res_blocks"] == 2 and m2.config["batch_norm"] is True
        np.testing.assert_array_equal(m2(x, training=False), y0)
        for v in m2.variables:
            v.assign(np.zeros(v.shape, np.float32))
        m2.load_weights(path)
        np.testing.assert_array_equal(m2(x, training=False), y0)
        m2.release()
    m.release()


class _CustomLossFunctor:
    """tests/models/test_generator.py:60-83 (CustomLossFunctionTest), numpy in place of TensorFlow."""

    def __init__(self, weighted=False, loss_weight=1.0, track_metrics=True):
        self.name = "custom_loss_func_test"
        self.track_metrics, self.weighted, self.loss_weight = track_metrics, weighted, loss_weight

    def __call__(self, hr_batch, sr_batch, hr_critic, sr_critic, batch_metrics, epoch_metrics):
        loss = float(np.max(4 * hr_batch + sr_batch))
        if self.track_metrics:
            batch_metrics[self.name](loss)
            epoch_metrics[self.name](loss)
        return loss * self.loss_weight


def test_custom_loss_functions_like_the_reference_tests():
    """The reference's tests/models/test_generator.py: a lambda and a functor class next to MeanSquaredError in
    ``Generator(loss_functions=[...])`` - ``calculate_train_loss`` is their sum, ``generator_loss`` and the per-functor
    metrics record it (generator.py:91-106, 220-228)."""
    from simplesr_b200.generator import Generator, MeanSquaredError
    from simplesr_b200 import model_builder as MB
    rng = np.random.default_rng(0)
    b1 = rng.uniform(-1, 1, size=(2, 16, 16, 3)).astype(np.float32)
    b2 = rng.uniform(-1, 1, size=(2, 16, 16, 3)).astype(np.float32)
    tiny = MB.build_resnet(upsample_factor=2, num_res_blocks=1, batch_normalization=False, seed=0)
    want_mse = float(np.mean((b1.astype(np.float64) - b2) ** 2))
    # 1. test_custom_loss_function_as_lambda
    mse = MeanSquaredError(track_metrics=False)
    g = Generator(upsample_factor=2, architecture="srresnet", pretrained_model=tiny,
                  loss_functions=[mse, lambda hr, sr, x, xx, xxx, xxxx: float(np.max(hr) + np.min(sr))])
    assert set(g.batch_metrics()) == {"mean_squared_error", "loss_function_1", "generator_loss"}
    _mse = mse(b1, b2, None, None, None, None)
    assert abs(_mse - want_mse) <= 2e-6 * want_mse
    custom = float(np.max(b2) + np.min(b1))
    loss = g.calculate_train_loss(b1, b2, None, None)                  # (sr_batch, hr_batch, ...) as generator.py:202
    assert loss == _mse + custom
    assert g.batch_metrics()["generator_loss"].result() == loss
    assert g.epoch_metrics()["generator_loss"].result() == loss
    assert g.epoch_metrics(train=False)["generator_loss"].result() == 0.0
    # 2. test_custom_loss_function_as_class
    mse = MeanSquaredError(track_metrics=True)
    fn = _CustomLossFunctor(track_metrics=True)
    g = Generator(upsample_factor=2, architecture="srresnet", pretrained_model=tiny, loss_functions=[mse, fn])
    custom = float(np.max(4 * b2 + b1))
    loss = g.calculate_train_loss(b1, b2, None, None)
    assert abs(loss - (want_mse + custom)) <= 1e-5
    for metrics in (g.batch_metrics(), g.epoch_metrics()):
        assert metrics["generator_loss"].result() == loss
        assert abs(metrics[mse.name].result() - want_mse) <= 2e-6 * want_mse
        assert metrics[fn.name].result() == custom
    v = g.calculate_validation_loss(b1, b2, None, None)
    assert v == loss and g.epoch_metrics(train=False)["generator_loss"].result() == loss
    with pytest.raises(ValueError):
        Generator(upsample_factor=2, architecture="srresnet", loss_functions=None)
    tiny.release()


def test_vgg19_weights_from_keras_h5():
    """build_vgg_19(load_custom_weights=True, custom_weights_path="...h5") as the reference calls it
    (model_builder.py:217-222): the stock Keras file layout with weights named block1_conv1_W_1:0."""
    from simplesr_b200 import h5lite, vgg as V
    src = V.build_vgg_19(seed=11)
    layers = [("input_1", [])]
    for layer in V.VGG19_LAYERS:
        name = layer[0]
        layers.append((name, [(f"{name}_W_1:0", src.kernels[name].numpy()), (f"{name}_b_1:0", src.biases[name].numpy())]
                       if len(layer) == 3 else []))
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "vgg19_weights_tf_dim_ordering_tf_kernels_notop.h5")
        h5lite.save_keras_weights(path, layers)
        got = V.build_vgg_19(load_custom_weights=True, custom_weights_path=path, seed=12)
        for a, b in zip(src.weights, got.weights):
            np.testing.assert_array_equal(a.numpy(), b.numpy())
        with pytest.raises(ValueError):
            V.build_vgg_19(load_custom_weights=True, custom_weights_path=os.path.join(d, "missing.h5"))


@pytest.mark.parametrize("layers,after,tv", [("block2_conv2", True, False), (["block1_conv2", "block2_conv2"], False, True),
                                             (["block2_conv2", "block3_conv4"], True, True)])
def test_vgg_loss_options(layers, after, tv):
    """after_activation=True (the reference default, vgg_loss.py:61), several output layers (:106-110, 162-164), total
    variation (:166-169) against the oracle's general VGG loss (itself checked against torch autograd)."""
    from simplesr_b200 import vgg as V
    from tests.test_gpu_vgg import _vgg_pair
    model, params = _vgg_pair()
    rng = np.random.default_rng(1)
    hr = rng.uniform(-1, 1, size=(2, 32, 48, 3)).astype(np.float32)
    sr = np.clip(hr + rng.normal(0, 0.2, size=hr.shape), -1, 1).astype(np.float32)
    kw = dict(feature_scale=1.0 / 12.75, loss_weight=0.5, after_activation=after, total_variation_loss=tv)
    fn = V.VGGLoss(output_layers=layers, vgg=model, total_varation_weight=2e-6, **kw)
    loss, grad = fn.loss_and_grad(hr, sr)
    ref_loss, ref_grad = O.vgg_loss_general(params, hr, sr, layers, total_variation_weight=2e-6, **kw)
    r16_loss, r16_grad = O.vgg_loss_general(params, hr, sr, layers, total_variation_weight=2e-6, act_dtype="bf16", **kw)
    assert abs(loss - ref_loss) <= 3e-2 * abs(ref_loss), (loss, ref_loss, r16_loss)
    assert abs(loss - r16_loss) <= 1.5e-2 * abs(r16_loss), (loss, r16_loss)
    cos = lambda a, b: float((a * b).sum() / (np.linalg.norm(a) * np.linalg.norm(b)))
    assert cos(grad, r16_grad) >= 0.98, cos(grad, r16_grad)
    assert cos(grad, ref_grad) >= cos(r16_grad, ref_grad) - 0.01
    if tv:
        only = V.VGGLoss(output_layers=layers, vgg=model, total_varation_weight=2e-6, **{**kw, "total_variation_loss": False})
        l0, g0 = only.loss_and_grad(hr, sr)
        den = (sr.astype(np.float64) + 1) * 127.5
        np.testing.assert_allclose(loss - l0, 2e-6 * O.total_variation(den).sum(), rtol=2e-3)
        np.testing.assert_allclose(grad - g0, 2e-6 * 127.5 * O.total_variation_backward(den), rtol=1e-3, atol=1e-6)
        only.release()
    fn.release()


def test_total_variation_kernel(ctx):
    rng = np.random.default_rng(0)
    x = rng.uniform(-1, 1, size=(3, 17, 23, 3)).astype(np.float32)
    d_x, d_g = L.DeviceBuffer.from_numpy(x), L.DeviceBuffer.from_numpy(np.ones_like(x))
    ws, out = L.DeviceBuffer(L.load().ssr_total_variation_workspace_bytes()), L.DeviceBuffer(4)
    L.total_variation(d_x, 3, 17, 23, 3, 127.5, 0.25, d_g, ws, out)
    L.stream_sync()
    den = (x.astype(np.float64) + 1) * 127.5
    np.testing.assert_allclose(out.download((1,), np.float32)[0], 0.25 * O.total_variation(den).sum(), rtol=1e-5)
    np.testing.assert_allclose(d_g.download(x.shape, np.float32), 1 + 0.25 * 127.5 * O.total_variation_backward(den),
                               rtol=1e-6, atol=1e-6)


PATCH_DIMS = [(1, 1), (2, 2), (3, 3), (3, 1), (1, 3), (2, 3), (3, 2)]      # the reference's own list, test_image_utils.py:10


def test_rectangular_patches_like_the_reference_tests(ctx):
    """tests/utils/image/test_image_utils.py:44-67 restated: the 3x3 and 5x3 matrices, every patch shape of PATCH_DIMS
    (width, height), segment and reconstruct - bit-exact against the oracle's restatement of _segment / _reconstruct."""
    from simplesr_b200 import image_utils as IU
    m33 = np.arange(1, 10, dtype=np.float32).reshape(3, 3, 1).repeat(3, axis=2)
    m53 = np.arange(1, 16, dtype=np.float32).reshape(3, 5, 1).repeat(3, axis=2)
    for pw, ph in PATCH_DIMS:
        for mat in (m33, m53):
            got, pad = IU.segment_into_patches(mat, patch_width=pw, patch_height=ph)
            ref, rpad = O.segment_into_patches(mat, patch_width=pw, patch_height=ph)
            assert got.shape[1] == ph and got.shape[2] == pw
            np.testing.assert_array_equal(got, np.asarray(ref))
            assert [list(map(int, p)) for p in pad] == [list(map(int, p)) for p in rpad]
            rec = IU.reconstruct_from_patches(got, original_height=mat.shape[0], original_width=mat.shape[1],
                                              horizontal_padding=pad[0][1], vertical_padding=pad[1][1])
            np.testing.assert_array_equal(rec, mat)
    with pytest.raises(ValueError):
        IU.segment_into_patches(m53, patch_width=2, patch_height=3, pixel_overlap=1)
    with pytest.raises(ValueError):
        IU.segment_into_patches(m33, patch_width=4, patch_height=1)


def test_model_call_training_true_with_batch_norm():
    """generator.generate(lr, training=True) on Generator.srresnet() (batch_norm=True, generator.py:285,200): batch
    statistics in the forward pass and the moving averages move (momentum 0.8)."""
    nb, sf = 2, 2
    m, params, bn = _setup_bn(nb, sf)
    rng = np.random.default_rng(3)
    x = rng.uniform(0, 1, size=(2, 12, 10, 3)).astype(np.float32)
    hr = np.zeros((2, 24, 20, 3), np.float32)
    y = m(x, training=True)
    stats = {}
    _, sr32, _ = O.srresnet_loss_and_grads(params, x, hr, upsample_factor=sf, num_res_blocks=nb, bn=bn, stats_out=stats)
    # batch-statistics normalisation divides by the std of bf16-rounded conv outputs: the bf16-storage oracle itself sits
    # at 52.4 dB / 2.2e-2 from the fp32 oracle on this network
    assert float(O.psnr(y, sr32, max_val=2.0).min()) > 49.0
    assert rel_err(y, sr32) <= 4e-2
    cnt = x.shape[0] * x.shape[1] * x.shape[2]
    mv = {v.name: v.numpy() for v in m.non_trainable_variables}
    for name in bn:
        mu, var = stats[name]
        np.testing.assert_allclose(mv[f"{name}_bn/moving_mean:0"], 0.8 * bn[name]["mean"] + 0.2 * mu, rtol=2e-2, atol=2e-3)
        np.testing.assert_allclose(mv[f"{name}_bn/moving_variance:0"], 0.8 * bn[name]["var"] + 0.2 * var * cnt / (cnt - 1),
                                   rtol=2e-2, atol=2e-3)
    y_inf = m(x, training=False)                 # inference folds the (moved) moving statistics: a different function
    assert np.abs(y_inf - y).max() > 1e-3
    m.release()


def test_label_smoothing_targets(ctx):
    """discriminator.py:240-254: SR labels U(0,1) * offset, HR labels 1 - offset + U(0, 0.5), new every step; the
    discriminator loss kernel uses them per sample."""
    from simplesr_b200 import discriminator as DM
    d = DM.build_discriminator(input_dims=(64, 64), relativistic=True, seed=3)
    loss = DM.RaGANLoss(d, loss_weight=5e-3, learning_rate=0.0, label_smoothing=True, smoothing_offset=0.3, seed=11)
    n = 4
    rng = np.random.default_rng(0)
    hr = rng.uniform(-1, 1, size=(n, 64, 64, 3)).astype(np.float32)
    sr = np.clip(hr + rng.normal(0, 0.3, size=hr.shape), -1, 1).astype(np.float32)
    B, ops = {}, []
    dh, ds, g = L.DeviceBuffer.from_numpy(hr), L.DeviceBuffer.from_numpy(sr), L.DeviceBuffer(sr.nbytes)
    g.zero()
    out = loss.emit(ops, B, "t_", n, 64, 64, dh, ds, g)
    seen = []
    for _ in range(2):
        loss.pre_step(None)
        for op in ops:
            op(None)
        L.stream_sync(None)
        loss._side_stream().sync()
        lab = B["t_labels"].download((2 * n,), np.float32)
        seen.append(lab.copy())
        assert np.all(lab[:n] >= 0.7 - 1e-6) and np.all(lab[:n] <= 1.2 + 1e-6)        # 1 - 0.3 + U(0, 0.5)
        assert np.all(lab[n:] >= 0.0) and np.all(lab[n:] <= 0.3 + 1e-6)
        hc, sc = B["t_hr_critic"].download((n,), np.float32), B["t_sr_critic"].download((n,), np.float32)
        R = O.ragan_losses(hc, sc, hr_label=lab[:n].astype(np.float64), sr_label=lab[n:].astype(np.float64))
        o = out.download((2,), np.float32)
        np.testing.assert_allclose(o, [R["g_loss"], R["d_loss"]], rtol=1e-5)
        np.testing.assert_allclose(B["t_d_dhr"].download((n, 1), np.float32), R["d_dhr"], rtol=1e-4, atol=1e-7)
    assert not np.array_equal(seen[0], seen[1])


def test_srmodel_facade_runs_the_reference_loop():
    """The protocol run_training drives (operations/training.py:36-112) on the 'gan' model type: two epochs of two
    batches, validation, histories, early-stopping bookkeeping, save_model and checkpoint restore."""
    from simplesr_b200 import vgg as V
    from simplesr_b200.generator import Discriminator, Generator
    from simplesr_b200.sr_model import Adam, PiecewiseConstantDecay, SRModel
    gen = Generator.esrgan_generator(upsample_factor=4, num_blocks=1, vgg=V.build_vgg_19(seed=2))
    disc = Discriminator.initialize_relativistic(input_dims=(64, 64), seed=3)
    sched = PiecewiseConstantDecay([2], [1e-4, 5e-5])
    with tempfile.TemporaryDirectory() as d:
        class Cfg:
            model_dir, checkpoint_dir = os.path.join(d, "models"), os.path.join(d, "ckpt")
        srm = SRModel("gan", gen, generator_optimizer=Adam(learning_rate=sched), discriminator=disc,
                      discriminator_optimizer=lambda: Adam(learning_rate=1e-4), early_stop_metric="psnr",
                      early_stop_patience=5, config=Cfg)
        rng = np.random.default_rng(0)
        batches = [(rng.uniform(0, 1, size=(2, 16, 16, 3)).astype(np.float32),
                    rng.uniform(-1, 1, size=(2, 64, 64, 3)).astype(np.float32)) for _ in range(2)]
        for epoch in range(2):
            assert srm.generator_optimizer().iterations.numpy() == 2 * epoch
            assert srm.generator_optimizer()._decayed_lr(None).numpy() == pytest.approx(1e-4 if epoch == 0 else 1e-4 if 2 * epoch <= 2 else 5e-5)
            assert not srm.stop_early()
            srm.before_epoch()
            for lr_b, hr_b in batches:
                out = srm.train_step(lr_b, hr_b)
                srm.after_train_batch()
                assert {"generator_loss", "psnr", "vgg_loss", "ra_adversarial_loss", "ra_discriminator_loss"} <= set(out)
            for lr_b, hr_b in batches[:1]:
                srm.validation_step(lr_b, hr_b)
                srm.after_validation_batch()
            assert "Generator" in srm.formatted_epoch_metrics()
            srm.after_epoch()
        assert srm.iterations() == 4
        hist = srm.epoch_history(train=True)
        assert len(hist["generator_loss"]) == 2 and len(hist["psnr"]) == 2 and len(hist["ra_discriminator_loss"]) == 2
        assert len(srm.batch_history(train=True)["generator_loss"]) == 4
        assert os.path.exists(os.path.join(Cfg.model_dir, "gan_gen_2.h5"))
        # checkpoint round trip: weights, Adam slots and the optimizer clock come back
        ck = srm.save_checkpoint()
        w = srm.generator().get_weights()
        srm.train_step(*batches[0])
        assert any(np.abs(a - b).max() > 0 for a, b in zip(srm.generator().get_weights(), w))
        srm.restore_checkpoint(ck)
        for a, b in zip(srm.generator().get_weights(), w):
            np.testing.assert_array_equal(a, b)
        assert srm.generator_optimizer().iterations.numpy() == 4
        srm.after_training()
        assert os.path.exists(os.path.join(Cfg.model_dir, "gan_gen_best.h5"))
    with pytest.raises(ValueError):
        SRModel("resnet", gen, generator_optimizer=Adam(), discriminator=disc)
    with pytest.raises(ValueError):
        SRModel("vae", gen, generator_optimizer=Adam())


def test_srmodel_standard_gan_recipe():
    """SRGAN recipe through the facade: SRResNet generator + MSE + AdversarialLoss against the sigmoid critic
    (Discriminator.initialize_standard, discriminator.py:306-361), label smoothing on."""
    from simplesr_b200.generator import AdversarialLoss, Discriminator, Generator, MeanSquaredError
    from simplesr_b200.sr_model import Adam, SRModel
    gen = Generator(upsample_factor=4, architecture="srresnet", num_blocks=2, batch_norm=False,
                    loss_functions=[MeanSquaredError(), AdversarialLoss(weighted=True, loss_weight=1e-3)])
    disc = Discriminator.initialize_standard(input_dims=(64, 64), label_smoothing=True, seed=3)
    srm = SRModel("gan", gen, generator_optimizer=Adam(learning_rate=1e-4), discriminator=disc,
                  discriminator_optimizer=Adam(learning_rate=1e-4))
    rng = np.random.default_rng(0)
    lr_b = rng.uniform(0, 1, size=(2, 16, 16, 3)).astype(np.float32)
    hr_b = rng.uniform(-1, 1, size=(2, 64, 64, 3)).astype(np.float32)
    srm.before_epoch()
    for _ in range(3):
        out = srm.train_step(lr_b, hr_b)
        srm.after_train_batch()
    assert {"generator_loss", "mean_squared_error", "adversarial_loss", "discriminator_loss"} <= set(out), out
    assert np.isfinite(out["adversarial_loss"]) and out["adversarial_loss"] > 0
    assert np.isfinite(out["discriminator_loss"]) and out["discriminator_loss"] > 0
    # the relativistic functor does not pair with the sigmoid critic
    from simplesr_b200.generator import RaAdversarialLoss
    gen2 = Generator(upsample_factor=4, architecture="srresnet", num_blocks=1, batch_norm=False,
                     loss_functions=[MeanSquaredError(), RaAdversarialLoss()])
    with pytest.raises(ValueError):
        SRModel("gan", gen2, generator_optimizer=Adam(), discriminator=disc, discriminator_optimizer=Adam())


def test_learnrate_scheduling_like_the_reference_test():
    """The reference's tests/models/test_learnrate_scheduling.py: an optimizer CONFIG with a serialised
    PiecewiseConstantDecay (boundaries [2, 5]) and beta_1 / beta_2; the decayed learning rate read before each of seven
    one-step epochs follows 3e-4 x3, 2e-5 x3, 3e-6, and the hyper-parameters are readable through ``_get_hyper``."""
    from simplesr_b200.generator import Generator
    from simplesr_b200.sr_model import Adam, SRModel
    boundaries, rates = [2, 5], [3e-4, 2e-5, 3e-6]
    expected = [3e-4, 3e-4, 3e-4, 2e-5, 2e-5, 2e-5, 3e-6]
    cfg = {"learning_rate": {"class_name": "PiecewiseConstantDecay", "config": {"boundaries": boundaries, "values": rates}},
           "beta_1": 0.5, "beta_2": 0.8}
    srm = SRModel(model_type="resnet", generator=Generator.srresnet(upsample_factor=2, num_blocks=1),
                  generator_optimizer=Adam, generator_optimizer_config=cfg)
    opt = srm.generator_optimizer()
    assert opt._get_hyper("learning_rate").boundaries == boundaries
    np.testing.assert_array_almost_equal(rates, opt._get_hyper("learning_rate").values, decimal=6)
    assert abs(opt._get_hyper("beta_1").numpy() - 0.5) <= 1e-5 and abs(opt._get_hyper("beta_2").numpy() - 0.8) <= 1e-5
    rng = np.random.default_rng(0)
    lr_b = rng.uniform(0, 1, size=(2, 16, 16, 3)).astype(np.float32)
    hr_b = rng.uniform(-1, 1, size=(2, 32, 32, 3)).astype(np.float32)
    w_prev = [v.numpy().copy() for v in srm.generator().trainable_variables]   # not the BN moving statistics
    for i in range(7):
        assert abs(srm.generator_optimizer()._decayed_lr(np.float32).numpy() - expected[i]) <= 1e-9, i
        srm.train_step(lr_b, hr_b)
        w = [v.numpy().copy() for v in srm.generator().trainable_variables]
        # Adam moves a weight by at most ~lr per step (|m_hat / sqrt(v_hat)| <= 1 up to the bias-correction ratio, which
        # is largest on the first steps with beta_1 0.5 / beta_2 0.8): the device applied THIS step's scheduled rate
        step = max(float(np.abs(a - b).max()) for a, b in zip(w, w_prev))
        assert 0.2 * expected[i] <= step <= 3.0 * expected[i], (i, step, expected[i])
        w_prev = w


def test_discriminator_labels_like_the_reference_test():
    """The reference's tests/models/test_discriminator.py on the facade object (standard critic, 80 x 80 input)."""
    from simplesr_b200.generator import Discriminator, DiscriminatorLoss
    off = 0.3
    disc = Discriminator(loss_function=DiscriminatorLoss(weighted=False), relativistic=False, label_smoothing=True,
                         smoothing_offset=off, input_dims=(80, 80))
    sr_critic, hr_critic = np.ones((50,), np.float64), np.zeros((50,), np.float64)
    sr_labels, hr_labels = disc._get_labels(sr_critic, hr_critic)
    assert len(sr_labels) == 50 and len(hr_labels) == 50 and sr_labels.dtype == np.float64
    assert sr_labels.min() >= 0 and sr_labels.max() <= off and sr_labels.std() > 0
    assert hr_labels.min() >= 1 - off and hr_labels.max() <= 1 + off and hr_labels.std() > 0
    disc = Discriminator(loss_function=DiscriminatorLoss(weighted=False), relativistic=False, label_smoothing=False,
                         input_dims=(80, 80))
    sr_labels, hr_labels = disc._get_labels(sr_critic, hr_critic)
    assert len(sr_labels) == 50 and len(hr_labels) == 50
    assert sr_labels.min() == 0 and sr_labels.max() == 0 and sr_labels.std() == 0
    assert hr_labels.min() == 1 and hr_labels.max() == 1 and hr_labels.std() == 0


def test_image_metrics_like_the_reference_test():
    """The reference's tests/models/test_srmodel.py::test_metrics: default ``dict(psnr=metrics.psnr)``, every function of
    simple_sr.utils.image.metrics, a user lambda, and a PSNR on non-normalised images; ``_update_metrics`` feeds epoch and
    batch dictionaries, ``after_train_batch`` moves the batch value into the history and resets it."""
    from simplesr_b200 import metrics
    from simplesr_b200.generator import Generator, MeanSquaredError
    from simplesr_b200.sr_model import Adam, SRModel
    rng = np.random.default_rng(0)
    b1 = rng.uniform(-1, 1, size=(2, 24, 24, 3)).astype(np.float32)
    b2 = np.clip(b1 + rng.normal(0, 0.1, size=b1.shape), -1, 1).astype(np.float32)

    def load(image_metrics=None):
        gen = Generator(upsample_factor=2, architecture="srresnet", num_blocks=1, batch_norm=False,
                        loss_functions=MeanSquaredError())
        return SRModel("resnet", gen, generator_optimizer=Adam, image_metrics=image_metrics)

    def initialised(model, expected):
        assert len(model._image_metrics) == len(expected)
        for key, func in expected.items():
            assert model._image_metrics[key] is func
            for d in (model._train_epoch_metrics, model._valid_epoch_metrics, model._batch_metrics):
                assert key in d and d[key].result() == 0.0

    m = load()
    initialised(m, dict(psnr=metrics.psnr))
    m._update_metrics(b1, b2, m._train_epoch_metrics)
    m._update_metrics(b1, b2, m._valid_epoch_metrics)
    want = float(np.mean(metrics.psnr(b1, b2, max_val=2.0)))
    assert abs(want - float(np.mean(O.psnr(b1, b2, max_val=2.0)))) <= 1e-3
    for d in (m._train_epoch_metrics, m._valid_epoch_metrics, m._batch_metrics):
        assert d["psnr"].result() == pytest.approx(want, rel=1e-6)
    assert len(m._train_batch_history["psnr"]) == 0
    m.after_train_batch()
    assert m._train_batch_history["psnr"] == [pytest.approx(want, rel=1e-6)]
    assert m._batch_metrics["psnr"].result() == 0.0 and m._train_epoch_metrics["psnr"].result() == pytest.approx(want)
    # every metric of the module
    expected = dict(psnr=metrics.psnr, PSNR_Y=metrics.psnr_on_y, SSIM=metrics.ssim)
    m2 = load(expected)
    initialised(m2, expected)
    m2._update_metrics(b1, b2, m2._train_epoch_metrics)
    vals = {k: float(np.mean(f(b1, b2, max_val=2.0))) for k, f in expected.items()}
    for k, v in vals.items():
        assert m2._train_epoch_metrics[k].result() == pytest.approx(v, rel=1e-6)
        assert m2._batch_metrics[k].result() == pytest.approx(v, rel=1e-6)
        assert m2._valid_epoch_metrics[k].result() == 0.0
    # a custom metric, and a PSNR on images in [0, 255]
    m3 = load(dict(PSNR_Y=metrics.psnr_on_y, ADD=lambda img1, img2: img1 + img2,
                   PSNR=lambda img1, img2: metrics.psnr(img1 * 127.5 + 127.5, img2 * 127.5 + 127.5, max_val=255)))
    m3._update_metrics(b1, b2, m3._train_epoch_metrics)
    assert m3._train_epoch_metrics["ADD"].result() == pytest.approx(float(np.mean(b1 + b2)), abs=1e-6)
    assert m3._train_epoch_metrics["PSNR"].result() == pytest.approx(want, abs=2e-3)      # same images, rescaled
    # inside train_step: the fused psnr comes from the device, the others are evaluated on the generated batch
    m4 = load(dict(psnr=metrics.psnr, ADD=lambda hr, sr: hr + sr, SSIM=metrics.ssim))
    lr_b = rng.uniform(0, 1, size=(2, 12, 12, 3)).astype(np.float32)
    out = m4.train_step(lr_b, b1)
    sr = m4._trainer.last_sr()
    assert sr.shape == b1.shape
    assert m4._train_epoch_metrics["ADD"].result() == pytest.approx(float(np.mean(b1 + sr)), abs=1e-6)
    assert m4._train_epoch_metrics["SSIM"].result() == pytest.approx(float(np.mean(metrics.ssim(b1, sr))), rel=1e-5)
    assert m4._train_epoch_metrics["psnr"].result() == pytest.approx(out["psnr"], rel=1e-6)
    assert out["psnr"] == pytest.approx(float(np.mean(metrics.psnr(b1, sr))), abs=2e-3)
    v = m4.validation_step(lr_b, b1)
    assert m4._valid_epoch_metrics["psnr"].result() == pytest.approx(v["psnr"], rel=1e-6)
    assert m4._valid_epoch_metrics["ADD"].result() != 0.0


def test_models_from_the_reference_yaml_schema():
    """``Generator.from_yaml`` / ``Discriminator.from_yaml`` / the ``model:`` section as ConfigUtil.from_yaml assembles it
    (generator.py:452-472, discriminator.py:363-383, config_util.py:311-331): the reference's minimal example and an
    ESRGAN recipe, trained for a few steps."""
    from simplesr_b200.generator import Discriminator, Generator, MeanSquaredError
    from simplesr_b200.sr_model import SRModel
    from simplesr_b200.training import PiecewiseConstantDecay
    gold = os.path.join(os.path.dirname(__file__), "golden")
    rng = np.random.default_rng(0)
    # 1. examples/training/minimal_example.yaml: SRResNet x2 (16 blocks, batch norm off by the constructor default), MSE
    gen = Generator.from_yaml(os.path.join(gold, "minimal_example.yaml"))
    assert gen.model().architecture == "srresnet" and gen.model().upsample_factor == 2
    assert gen.model().config["num_res_blocks"] == 16 and gen.model().config["batch_norm"] is False
    assert [type(f) for f in gen.loss_functions()] == [MeanSquaredError]
    gen.model().release()
    srm = SRModel.from_yaml(os.path.join(gold, "minimal_example.yaml"))
    assert srm.name == "resnet" and srm.discriminator() is None
    lr_b = rng.uniform(0, 1, size=(2, 12, 12, 3)).astype(np.float32)
    hr_b = rng.uniform(-1, 1, size=(2, 24, 24, 3)).astype(np.float32)
    w0 = srm.generator().convs["last"].kernel.numpy().copy()
    for _ in range(3):
        out = srm.train_step(lr_b, hr_b)
    assert np.isfinite(out["generator_loss"]) and out["generator_loss"] == pytest.approx(out["mean_squared_error"])
    assert srm.generator_optimizer().iterations.numpy() == 3
    assert np.abs(srm.generator().convs["last"].kernel.numpy() - w0).max() > 1e-4    # Adam at the Keras default rate
    # 2. an ESRGAN recipe: RRDB + MAE + RaGAN + VGG, relativistic critic, scheduled generator learning rate
    path = os.path.join(gold, "esrgan_example.yaml")
    disc = Discriminator.from_yaml(path)
    assert disc.model().relativistic and disc.loss_function().name == "ra_discriminator_loss"
    srm = SRModel.from_yaml(path)
    assert srm.name == "gan"
    g = srm.generator()
    assert g.architecture == "rrdb" and g.upsample_factor == 4
    assert g.config["num_rrdb_blocks"] == 1 and g.config["num_dense_blocks"] == 2
    sch = srm.generator_optimizer()._get_hyper("learning_rate")
    assert isinstance(sch, PiecewiseConstantDecay) and sch.boundaries == [2, 4]
    assert abs(srm.generator_optimizer()._get_hyper("beta_2").numpy() - 0.99) < 1e-6
    lr_b = rng.uniform(0, 1, size=(2, 16, 16, 3)).astype(np.float32)
    hr_b = rng.uniform(-1, 1, size=(2, 64, 64, 3)).astype(np.float32)
    for i in range(4):
        want = [1e-4, 1e-4, 1e-4, 5e-5][i]
        assert srm.generator_optimizer()._decayed_lr(None).numpy() == pytest.approx(want)
        out = srm.train_step(lr_b, hr_b)
    assert {"generator_loss", "mean_absolute_error", "vgg_loss", "ra_adversarial_loss", "ra_discriminator_loss"} <= set(out)
    assert all(np.isfinite(v) for v in out.values())
