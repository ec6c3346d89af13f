"""ESRGAN discriminator + relativistic-average GAN losses on the B200 vs the oracle (which is itself checked against
torch autograd in tests/test_oracle_backward.py): critics, both losses, the generator-side gradient w.r.t. the SR image
through D (BatchNorm backward included), and every discriminator gradient accumulated over the two critic passes."""
import numpy as np
import pytest

from tests.helpers import L, O, rel_err

pytestmark = pytest.mark.gpu
HW = (64, 64)   # the last BatchNorm then sees 4 x 4 x 4 samples per channel (32x32 would leave 16: ill-conditioned)


def _disc_pair(seed=3):
    from simplesr_b200 import discriminator as DM
    params = O.init_discriminator_params(seed=seed, input_hw=HW, bias_std=0.05)
    rng = np.random.default_rng(7)
    for k in list(params):
        if k.endswith("_bn"):
            params[k] = [(1 + 0.1 * rng.standard_normal(params[k][0].shape)).astype(np.float32),
                         (0.1 * rng.standard_normal(params[k][1].shape)).astype(np.float32)]
    d = DM.build_discriminator(input_dims=HW, relativistic=True, seed=0)
    d.set_params(params)
    return d, params


def test_discriminator_structure():
    from simplesr_b200 import discriminator as DM
    d = DM.build_discriminator(input_dims=(128, 128), relativistic=True, seed=0)
    assert len(d.trainable_variables) == 34                     # 8 convs x2 + 7 BN x2 + 2 dense x2
    # SURVEY.md §8a counts 38,241,857 conv + dense parameters; the 7 BatchNorm layers add gamma and beta
    assert d.count_params() == 38241857 + 2 * (64 + 128 * 2 + 256 * 2 + 512 * 2)
    std = DM.build_discriminator(input_dims=(128, 128), relativistic=False, seed=0)     # same variables, sigmoid critic
    assert not std.relativistic and std.count_params() == d.count_params()


def test_standard_gan_losses_kernel(ctx):
    """ssr_gan_losses_ex (AdversarialLoss / DiscriminatorLoss on sigmoid critics, label smoothing) vs the oracle."""
    rng = np.random.default_rng(2)
    n = 16
    zh = rng.normal(1.0, 3.0, n).astype(np.float32)
    zs = rng.normal(-1.0, 3.0, n).astype(np.float32)
    zs[0], zh[0] = 30.0, -30.0                       # saturated: the probability clip switches the gradient off
    lh = (0.7 + 0.5 * rng.uniform(size=n)).astype(np.float32)
    ls = (0.3 * rng.uniform(size=n)).astype(np.float32)
    b = [L.DeviceBuffer.from_numpy(a) for a in (zh, zs, lh, ls)] + [L.DeviceBuffer(max(8, n * 4)) for _ in range(4)]
    for labels in (False, True):
        L.ragan_losses_ex(b[0], b[1], n, 1.0, 0.0, b[2] if labels else None, b[3] if labels else None, b[4], b[5], b[6],
                          b[7], None, relativistic=False)
        R = O.gan_losses(zh, zs, hr_label=lh if labels else 1.0, sr_label=ls if labels else 0.0)
        o = b[4].download((2,), np.float32)
        np.testing.assert_allclose(o, [R["g_loss"], R["d_loss"]], rtol=1e-5)
        for buf, key in ((b[5], "g_dsr"), (b[6], "d_dsr"), (b[7], "d_dhr")):
            np.testing.assert_allclose(buf.download((n, 1), np.float32), R[key], rtol=1e-4, atol=1e-9)


def test_standard_gan_step_against_oracle(ctx):
    """GANLoss inside a step: sigmoid critics, both BCE losses, generator-side gradient and D's weight gradients."""
    from simplesr_b200 import discriminator as DM
    d, params = _disc_pair()
    d.relativistic = False
    loss = DM.GANLoss(d, loss_weight=1e-3, learning_rate=0.0)
    assert loss.metric_names == ["adversarial_loss", "discriminator_loss"]
    n = 4
    rng = np.random.default_rng(0)
    hr = rng.uniform(-1, 1, size=(n, *HW, 3)).astype(np.float32)
    sr = np.clip(hr + rng.normal(0, 0.3, size=hr.shape), -1, 1).astype(np.float32)
    B, ops = {}, []
    dh, ds = L.DeviceBuffer.from_numpy(hr), L.DeviceBuffer.from_numpy(sr)
    g = L.DeviceBuffer(sr.nbytes)
    g.zero()
    out = loss.emit(ops, B, "t_", n, HW[0], HW[1], dh, ds, g)
    for op in ops:
        op(None)
    L.stream_sync(None)
    ch, cs = {}, {}
    hc = O.discriminator_forward(params, hr, cache=ch)
    sc = O.discriminator_forward(params, sr, cache=cs)
    R = O.gan_losses(hc, sc)
    o = out.download((2,), np.float32)
    assert abs(o[0] - R["g_loss"]) <= 2e-2 * abs(R["g_loss"]) and abs(o[1] - R["d_loss"]) <= 2e-2 * abs(R["d_loss"])
    got_sc = B["t_sr_critic"].download((n, 1), np.float32)
    got_hc = B["t_hr_critic"].download((n, 1), np.float32)
    R2 = O.gan_losses(got_hc, got_sc)
    np.testing.assert_allclose(o, [R2["g_loss"], R2["d_loss"]], rtol=1e-5)
    cos = lambda a, b: float((a * b).sum() / (np.linalg.norm(a) * np.linalg.norm(b) + 1e-30))
    dx, _ = O.discriminator_backward(params, cs, R["g_dsr"])
    got_g = g.download(sr.shape, np.float32)
    assert cos(got_g, 1e-3 * dx) >= 0.97, cos(got_g, 1e-3 * dx)
    _, gs = O.discriminator_backward(params, cs, R["d_dsr"])
    _, gh = O.discriminator_backward(params, ch, R["d_dhr"])
    got = loss.gradients()
    for name in ("d_dense0", "d_dense1", "d_conv7", "d_conv0"):
        ref = (gs[name][0] + gh[name][0]).reshape(got[name][0].shape)
        assert cos(got[name][0], ref) >= 0.97, (name, cos(got[name][0], ref))


def test_ragan_step_against_oracle(ctx):
    from simplesr_b200 import discriminator as DM
    d, params = _disc_pair()
    loss = DM.RaGANLoss(d, loss_weight=5e-3, learning_rate=0.0)
    n = 4
    rng = np.random.default_rng(0)
    hr = rng.uniform(-1, 1, size=(n, *HW, 3)).astype(np.float32)
    sr = np.clip(hr + rng.normal(0, 0.3, size=hr.shape), -1, 1).astype(np.float32)
    B, ops = {}, []
    dh, ds = L.DeviceBuffer.from_numpy(hr), L.DeviceBuffer.from_numpy(sr)
    g = L.DeviceBuffer(sr.nbytes)
    g.zero()
    out = loss.emit(ops, B, "t_", n, HW[0], HW[1], dh, ds, g)
    for op in ops:
        op(None)
    L.stream_sync(None)
    # ---- oracle
    ch, cs = {}, {}
    hc = O.discriminator_forward(params, hr, cache=ch)
    sc = O.discriminator_forward(params, sr, cache=cs)
    hc16 = O.discriminator_forward(params, hr, act_dtype="bf16")
    R = O.ragan_losses(hc, sc)
    got_sc = B["t_sr_critic"].download((n, 1), np.float32)
    got_hc = B["t_hr_critic"].download((n, 1), np.float32)
    tol_c = 3 * float(np.abs(hc16 - hc).max()) + 1e-3 * float(np.abs(hc).max())
    assert np.abs(got_hc - hc).max() <= tol_c and np.abs(got_sc - sc).max() <= tol_c, (got_hc.ravel(), hc.ravel())
    o = out.download((2,), np.float32)
    assert abs(o[0] - R["g_loss"]) <= 2e-2 * abs(R["g_loss"]) and abs(o[1] - R["d_loss"]) <= 2e-2 * abs(R["d_loss"])
    # exact check of the loss kernel itself on the GPU's own critics
    R2 = O.ragan_losses(got_hc, got_sc)
    np.testing.assert_allclose(o, [R2["g_loss"], R2["d_loss"]], rtol=1e-5)
    np.testing.assert_allclose(B["t_g_dsr"].download((n, 1), np.float32), R2["g_dsr"], rtol=1e-4, atol=1e-7)
    np.testing.assert_allclose(B["t_d_dsr"].download((n, 1), np.float32), R2["d_dsr"], rtol=1e-4, atol=1e-7)
    np.testing.assert_allclose(B["t_d_dhr"].download((n, 1), np.float32), R2["d_dhr"], rtol=1e-4, atol=1e-7)
    # ---- generator-side gradient through D(sr)
    dx, _ = O.discriminator_backward(params, cs, R["g_dsr"])
    cos = lambda a, b: float((a * b).sum() / (np.linalg.norm(a) * np.linalg.norm(b) + 1e-30))
    got_g = g.download(sr.shape, np.float32)
    assert cos(got_g, 5e-3 * dx) >= 0.97, cos(got_g, 5e-3 * dx)
    np.testing.assert_allclose(np.linalg.norm(got_g), np.linalg.norm(5e-3 * dx), rtol=0.1)
    # ---- discriminator gradients (sum over both critic passes)
    _, gs = O.discriminator_backward(params, cs, R["d_dsr"])
    _, gh = O.discriminator_backward(params, ch, R["d_dhr"])
    got = loss.gradients()
    for name in gs:
        for i in range(2):
            if i == 1 and name.startswith("d_conv") and not name.endswith("_bn") and name != "d_conv0":
                continue      # bias in front of BatchNorm: exactly zero gradient, only rounding noise on both sides
            ref = (gs[name][i] + gh[name][i]).reshape(got[name][i].shape)
            if np.abs(ref).max() < 1e-6:
                continue      # e.g. the last bias: the relativistic gradients of the two critic passes cancel exactly
            assert cos(got[name][i], ref) >= 0.97, (name, i, cos(got[name][i], ref))
            # (no element-wise bound: one LeakyReLU sign flip of a hidden unit with |h| below bf16 resolution rescales a whole
            #  column of the dense gradient by 5x; direction and magnitude are what matters)
            ratio = float(np.linalg.norm(got[name][i]) / np.linalg.norm(ref))
            assert 0.85 <= ratio <= 1.15, (name, i, ratio)


def test_bn_dense_subsample_kernels(ctx):
    rng = np.random.default_rng(1)
    px, c = 300, 64
    x = O.bf16_round(rng.standard_normal((px, c)).astype(np.float32) * 2 + 0.5)
    dx_ = L.DeviceBuffer.from_numpy(L.f32_to_bf16_bits(x))
    ws = L.DeviceBuffer(L.load().ssr_bn_workspace_bytes(c))
    mean, istd = L.DeviceBuffer(c * 4), L.DeviceBuffer(c * 4)
    mm, mv = L.DeviceBuffer.from_numpy(np.zeros(c, np.float32)), L.DeviceBuffer.from_numpy(np.ones(c, np.float32))
    L.bn_stats_bf16(dx_, px, c, 1e-3, 0.8, ws, mean, istd, mm, mv)
    np.testing.assert_allclose(mean.download((c,), np.float32), x.mean(0), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(istd.download((c,), np.float32), 1 / np.sqrt(x.var(0) + 1e-3), rtol=1e-5)
    np.testing.assert_allclose(mm.download((c,), np.float32), 0.2 * x.mean(0), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(mv.download((c,), np.float32), 0.8 + 0.2 * x.var(0, ddof=1), rtol=1e-5)
    # dense forward / backward
    n, K, Oo = 5, 700, 130
    xx = rng.standard_normal((n, K)).astype(np.float32)
    w = rng.standard_normal((K, Oo)).astype(np.float32) / 20
    b = rng.standard_normal(Oo).astype(np.float32)
    dxx, dw, db = L.DeviceBuffer.from_numpy(xx), L.DeviceBuffer.from_numpy(w), L.DeviceBuffer.from_numpy(b)
    dws = L.DeviceBuffer(L.load().ssr_dense_workspace_bytes(n, Oo))
    h, y = L.DeviceBuffer(n * Oo * 4), L.DeviceBuffer(n * Oo * 4)
    L.dense_fwd_f32(dxx, dw, db, n, K, Oo, True, 0.2, dws, h, y)
    ref_h = xx @ w + b
    np.testing.assert_allclose(h.download((n, Oo), np.float32), ref_h, rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(y.download((n, Oo), np.float32), O.leaky_relu(ref_h, 0.2), rtol=1e-4, atol=1e-4)
    dy = rng.standard_normal((n, Oo)).astype(np.float32)
    ddy = L.DeviceBuffer.from_numpy(dy)
    gx, gw, gb = L.DeviceBuffer(xx.nbytes), L.DeviceBuffer(w.nbytes), L.DeviceBuffer(b.nbytes)
    L.dense_bwd_f32(dxx, dw, ddy, n, K, Oo, gx, gw, gb, False)
    np.testing.assert_allclose(gx.download(xx.shape, np.float32), dy @ w.T, rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(gw.download(w.shape, np.float32), xx.T @ dy, rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(gb.download(b.shape, np.float32), dy.sum(0), rtol=1e-4, atol=1e-5)
    # stride-2 decimation and its adjoint are exact copies
    a = rng.integers(0, 65536, size=(2, 6, 8, 8)).astype(np.uint16)
    da, ds = L.DeviceBuffer.from_numpy(a), L.DeviceBuffer(2 * 3 * 4 * 8 * 2)
    L.subsample2(da, ds, 2, 3, 4, 8, 2)
    sub = ds.download((2, 3, 4, 8), np.uint16)
    np.testing.assert_array_equal(sub, a[:, 1::2, 1::2])
    dz = L.DeviceBuffer(a.nbytes)
    L.zero_insert2(ds, dz, 2, 3, 4, 8, 2)
    z = np.zeros_like(a)
    z[:, 1::2, 1::2] = sub
    np.testing.assert_array_equal(dz.download(a.shape, np.uint16), z)


@pytest.mark.parametrize("n,K,Oo", [(16, 2048, 256), (5, 1024, 1024), (3, 96, 64)])
def test_dense_tensor_core_kernels(ctx, n, K, Oo):
    """Wide Dense layers at batch <= 16 run as mma.sync TF32 with the 3xTF32 split (fp32-class accuracy): forward,
    input gradient, weight gradient (fresh and accumulated) against numpy in float64."""
    rng = np.random.default_rng(4)
    xx = rng.standard_normal((n, K)).astype(np.float32)
    w = (rng.standard_normal((K, Oo)) / np.sqrt(K)).astype(np.float32)
    b = rng.standard_normal(Oo).astype(np.float32)
    dy = rng.standard_normal((n, Oo)).astype(np.float32)
    dxx, dw, db, ddy = (L.DeviceBuffer.from_numpy(a) for a in (xx, w, b, dy))
    dws = L.DeviceBuffer(L.load().ssr_dense_workspace_bytes(n, Oo))
    h, y = L.DeviceBuffer(n * Oo * 4), L.DeviceBuffer(n * Oo * 4)
    L.dense_fwd_f32(dxx, dw, db, n, K, Oo, True, 0.2, dws, h, y)
    ref_h = xx.astype(np.float64) @ w.astype(np.float64) + b
    np.testing.assert_allclose(h.download((n, Oo), np.float32), ref_h, rtol=2e-5, atol=2e-5)
    gx, gw, gb = L.DeviceBuffer(xx.nbytes), L.DeviceBuffer(w.nbytes), L.DeviceBuffer(b.nbytes)
    L.dense_bwd_f32(dxx, dw, ddy, n, K, Oo, gx, gw, gb, False)
    ref_gx = dy.astype(np.float64) @ w.astype(np.float64).T
    ref_gw = xx.astype(np.float64).T @ dy.astype(np.float64)
    np.testing.assert_allclose(gx.download(xx.shape, np.float32), ref_gx, rtol=2e-5, atol=2e-5)
    np.testing.assert_allclose(gw.download(w.shape, np.float32), ref_gw, rtol=2e-5, atol=2e-5)
    L.dense_bwd_f32(dxx, dw, ddy, n, K, Oo, None, gw, gb, True)          # accumulate a second pass
    np.testing.assert_allclose(gw.download(w.shape, np.float32), 2 * ref_gw, rtol=2e-5, atol=4e-5)
    np.testing.assert_allclose(gb.download(b.shape, np.float32), 2 * dy.sum(0), rtol=1e-4, atol=1e-5)


def test_full_esrgan_step_against_oracle():
    """BASELINE configs[3] in miniature: RRDB generator + MAE*1e-2 + VGG19(block5_conv4, pre-activation)*1.0 +
    RaGAN*5e-3 (generator.py:433-438) and the discriminator update, one iteration; generator gradients against the
    oracle's composite loss, both adversarial losses against the oracle on the oracle's own SR image."""
    from simplesr_b200 import discriminator as DM
    from simplesr_b200 import model_builder as MB
    from simplesr_b200 import vgg as V
    from simplesr_b200.training import RRDBTrainer
    from tests.test_gpu_vgg import _vgg_pair
    nb, sf, n, lrs = 1, 4, 2, 16
    params = O.init_rrdb_params(seed=1, bias_std=0.05, upsample_factor=sf, num_rrdb_blocks=nb)
    m = MB.build_enhanced_resnet(upsample_factor=sf, num_rrdb_blocks=nb, seed=0)
    weights = []
    for name, _, _ in O.rrdb_layer_specs(upsample_factor=sf, num_rrdb_blocks=nb):
        weights.extend(params[name])
    m.set_weights(weights)
    vgg_model, vparams = _vgg_pair()
    d, dparams = _disc_pair()
    vl = V.VGGLoss(output_layers="block5_conv4", loss_weight=1.0, after_activation=False, vgg=vgg_model)
    gl = DM.RaGANLoss(d, loss_weight=5e-3, learning_rate=0.0)
    tr = RRDBTrainer(m, loss=("mae", 1e-2), learning_rate=0.0, extra_losses=[vl, gl])
    rng = np.random.default_rng(0)
    lr = rng.uniform(0, 1, size=(n, lrs, lrs, 3)).astype(np.float32)
    hr = rng.uniform(-1, 1, size=(n, lrs * sf, lrs * sf, 3)).astype(np.float32)
    out = tr.train_step(lr, hr)
    got = tr.gradients()
    adv = {}

    def extra(sr):
        lv, gv = O.vgg_loss_and_grad(vparams, hr, sr, output_layer="block5_conv4", loss_weight=1.0)
        ch, cs = {}, {}
        hc = O.discriminator_forward(dparams, hr, cache=ch)
        sc = O.discriminator_forward(dparams, sr, cache=cs)
        R = O.ragan_losses(hc, sc)
        adv.update(R)
        dx, _ = O.discriminator_backward(dparams, cs, R["g_dsr"])
        return lv + 5e-3 * R["g_loss"], gv + np.float32(5e-3) * dx

    loss32, sr32, g32 = O.rrdb_loss_and_grads(params, lr, hr, upsample_factor=sf, num_rrdb_blocks=nb, w_mse=0.0,
                                              w_mae=1e-2, extra_loss=extra)
    assert abs(out["loss"] - loss32) <= 3e-2 * abs(loss32), (out, loss32)
    assert abs(out["ra_adversarial_loss"] - 5e-3 * adv["g_loss"]) <= 5e-2 * 5e-3 * adv["g_loss"]
    assert abs(out["ra_discriminator_loss"] - adv["d_loss"]) <= 5e-2 * adv["d_loss"]
    cos = lambda a, b: float((a * b).sum() / (np.linalg.norm(a) * np.linalg.norm(b) + 1e-30))
    for name in ("last", "hr", "up1", "trunk", "rrdb0_db2_out", "rrdb0_db0_conv0", "fea"):
        assert cos(got[name][0], g32[name][0]) >= 0.95, (name, cos(got[name][0], g32[name][0]))
    dg = gl.gradients()
    assert all(np.isfinite(a).all() for v in dg.values() for a in v)
    assert np.abs(dg["d_dense0"][0]).max() > 0
    tr.release()
