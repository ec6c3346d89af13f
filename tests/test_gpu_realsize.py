"""BASELINE.json configs at their REAL sizes against the oracle (round-1 verdict: only C1 and C3 were).

C2  RRDB-23 x4, batch 16 of 128x128 LR through the default plan (paired growth convs, CTA pairs, carry, snake order,
    2432 pixel tiles over 148 SMs = 17 ring passes per CTA): two images against the oracle, batch independence for all
    16, 30 back-to-back launches (the regime that once hung).
C4  one ESRGAN step at batch 16 of 128x128 HR: every loss term and the gradients nearest the loss against the oracle.
C5  tiled inference of a 1024x1024 LR image (64 tiles of 192x192) against the oracle's stitch, and sharded.

The fp32 oracle of the C2 / C5 forward passes is evaluated through oracle/torch_cpu.py (the same graph on oneDNN; pinned
to the numpy oracle by tests/test_oracle_ops.py) - the numpy restatement needs ~5 s per 128x128 image.
"""
import numpy as np
import pytest

from tests.helpers import L, O, rel_err

pytestmark = pytest.mark.gpu


def _rrdb(nb, seed=1, bias_std=0.05):
    from simplesr_b200 import model_builder as MB
    params = O.init_rrdb_params(seed=seed, bias_std=bias_std, upsample_factor=4, num_rrdb_blocks=nb)
    m = MB.build_enhanced_resnet(upsample_factor=4, num_rrdb_blocks=nb, seed=0)
    weights = []
    for name, _, _ in O.rrdb_layer_specs(upsample_factor=4, num_rrdb_blocks=nb):
        weights.extend(params[name])
    m.set_weights(weights)
    return m, params


def test_c2_rrdb23_batch16_default_plan():
    import torch
    from oracle import torch_cpu as T
    T.set_threads(torch)
    m, params = _rrdb(23)
    rng = np.random.default_rng(0)
    x = rng.uniform(0, 1, size=(16, 128, 128, 3)).astype(np.float32)
    got = m(x, training=False)
    assert got.shape == (16, 512, 512, 3) and np.isfinite(got).all()
    plan = m.plan(16, 128, 128)
    assert plan.launches == 353                      # 351 convs (138 of them growth pairs + tails) + 2 elementwise
    for i in (0, 9):
        ref = T.rrdb_forward(params, x[i:i + 1], upsample_factor=4, num_rrdb_blocks=23)
        psnr = float(O.psnr(got[i:i + 1], ref, max_val=2.0).min())
        assert psnr > 50.0, (i, psnr)
        assert rel_err(got[i:i + 1], ref) <= 1e-2, (i, rel_err(got[i:i + 1], ref))
    # batch independence at full size: the same images in another batch order give the same pixels, bit for bit
    perm = np.roll(np.arange(16), 5)
    got_p = m(x[perm], training=False)
    for j, i in enumerate(perm):
        assert np.array_equal(got_p[j], got[i]), (j, i)
    # 30 back-to-back launches of the captured step, no host sync in between
    s = m.stream.ptr
    L.check(m.ctx.lib.ssr_memcpy_h2d(plan.buffers["in_f32"].ptr, x.ctypes.data, x.nbytes, s))
    for _ in range(30):
        plan.run(s)
    again = np.empty_like(got)
    L.check(m.ctx.lib.ssr_memcpy_d2h(again.ctypes.data, plan.buffers["out_f32"].ptr, again.nbytes, s))
    m.stream.sync()
    assert np.array_equal(again, got)
    m.release()
    # the same step with tile-level dependencies between the layers (ssr_conv_chain_*, an option): the dense-block convs
    # follow their predecessor tile by tile (all but the first, whose producer - the RGB conv - tiles differently), no
    # wait gives up, bit-identical
    m2, _ = _rrdb(23)
    m2.chain_deps = True
    got2 = m2(x, training=False)
    plan2 = m2.plan(16, 128, 128)
    assert plan2.chain_stats is not None and plan2.chain_stats[1] >= 344, plan2.chain_stats
    assert plan2.chain_timeouts() == 0
    assert np.array_equal(got2, got)
    m2.release()


def test_c4_esrgan_step_batch16_hr128():
    """configs[3] at size: RRDB-23 + MAE*1e-2 + VGG19 block5_conv4 pre-activation + RaGAN*5e-3, discriminator update."""
    from simplesr_b200 import discriminator as DM
    from simplesr_b200 import vgg as V
    from simplesr_b200.training import RRDBTrainer
    from tests.test_gpu_vgg import _vgg_pair
    nb, sf, n, lrs = 23, 4, 16, 32
    m, params = _rrdb(nb)
    vgg_model, vparams = _vgg_pair()
    dparams = O.init_discriminator_params(seed=3, input_hw=(128, 128), bias_std=0.05)
    d = DM.build_discriminator(input_dims=(128, 128), relativistic=True, seed=0)
    d.set_params(dparams)
    vl = V.VGGLoss(output_layers="block5_conv4", loss_weight=1.0, after_activation=False, vgg=vgg_model)
    gl = DM.RaGANLoss(d, loss_weight=5e-3, learning_rate=0.0)
    tr = RRDBTrainer(m, loss=("mae", 1e-2), learning_rate=0.0, extra_losses=[vl, gl])
    rng = np.random.default_rng(0)
    lr = rng.uniform(0, 1, size=(n, lrs, lrs, 3)).astype(np.float32)
    hr = rng.uniform(-1, 1, size=(n, lrs * sf, lrs * sf, 3)).astype(np.float32)
    out = tr.train_step(lr, hr)
    got = tr.gradients()
    adv, parts = {}, {}

    def extra(sr):
        lv, gv = O.vgg_loss_and_grad(vparams, hr, sr, output_layer="block5_conv4", loss_weight=1.0)
        cs = {}
        hc = O.discriminator_forward(dparams, hr)
        sc = O.discriminator_forward(dparams, sr, cache=cs)
        R = O.ragan_losses(hc, sc)
        adv.update(R)
        parts["vgg"] = float(lv)
        dx, _ = O.discriminator_backward(dparams, cs, R["g_dsr"])
        return lv + 5e-3 * R["g_loss"], gv + np.float32(5e-3) * dx

    loss32, sr32, g32 = O.rrdb_loss_and_grads(params, lr, hr, upsample_factor=sf, num_rrdb_blocks=nb, w_mse=0.0,
                                              w_mae=1e-2, extra_loss=extra)
    assert abs(out["loss"] - loss32) <= 3e-2 * abs(loss32), (out, loss32)
    assert abs(out["mae"] - float(O.mean_absolute_error(hr, sr32))) <= 2e-3 * out["mae"]
    assert abs(out["vgg_loss"] - parts["vgg"]) <= 3e-2 * parts["vgg"], (out["vgg_loss"], parts["vgg"])
    assert abs(out["ra_adversarial_loss"] - 5e-3 * adv["g_loss"]) <= 5e-2 * 5e-3 * adv["g_loss"]
    assert abs(out["ra_discriminator_loss"] - adv["d_loss"]) <= 5e-2 * adv["d_loss"]
    np.testing.assert_allclose(out["psnr"], float(np.mean(O.psnr(hr, sr32, 2.0))), rtol=2e-3)
    cos = lambda a, b: float((a * b).sum() / (np.linalg.norm(a) * np.linalg.norm(b) + 1e-30))
    for name in ("last", "hr", "up1", "up0", "trunk"):
        assert cos(got[name][0], g32[name][0]) >= 0.95, (name, cos(got[name][0], g32[name][0]))
    assert all(np.isfinite(g[0]).all() for g in got.values())
    dg = gl.gradients()
    assert all(np.isfinite(a).all() for v in dg.values() for a in v)
    tr.release()
    m.release()


def test_c5_tiled_1024_against_the_oracle_stitch_and_sharded():
    """64 tiles of 192x192 (patch 128, overlap 32) of a 1024x1024 LR image through upscale_tiled on a 1-block model:
    equal to the oracle's segment -> forward -> stitch within the conv tolerance; sharded over 2 and 8 ranks the union of
    the bands is the single-rank image bit for bit."""
    import torch
    from oracle import torch_cpu as T
    from simplesr_b200 import evaluation as EV
    T.set_threads(torch)
    m, params = _rrdb(1)
    rng = np.random.default_rng(5)
    lr = rng.uniform(0, 1, size=(1024, 1024, 3)).astype(np.float32)
    full = EV.upscale_tiled(m, lr, patch=128, pixel_overlap=32, tile_batch=16)
    assert full.shape == (4096, 4096, 3)
    fwd = lambda t: np.concatenate([T.rrdb_forward(params, t[i:i + 8], upsample_factor=4, num_rrdb_blocks=1)
                                    for i in range(0, t.shape[0], 8)])
    orc = O.tiled_upscale(fwd, lr, 4, patch=128, pixel_overlap=32)
    assert float(O.psnr(full, orc, max_val=2.0)) > 50.0
    assert rel_err(full, orc) <= 1e-2
    for world in (2, 8):
        acc = np.zeros_like(full)
        for r in range(world):
            part = EV.upscale_tiled(m, lr, patch=128, pixel_overlap=32, tile_batch=16, rank=r, world_size=world)
            b, c = EV.tile_range(64, r, world)
            (_, _), (o0, on) = EV.tile_band(1024, 1024, 128, 32, b, c)
            assert not part[:o0 * 4].any() and not part[(o0 + on) * 4:].any()      # only the rank's band is written
            acc += part
        np.testing.assert_array_equal(acc, full)
    m.release()
