"""SRResNet training iteration (the non-GAN branch of SRModel.train_step, sr_model.py:403-453) against the oracle:
loss, PSNR metric, every gradient, and the weights after Adam steps.

Tolerances.  Single ops (dgrad, wgrad, reductions) meet the per-layer bound max|err|/max|ref| <= 1e-2 in
tests/test_gpu_train_ops.py.  End to end, a gradient tensor inherits the bf16 rounding of every stored activation it
depends on (and PReLU kinks flip for pre-activations within one bf16 ulp of 0): the oracle itself, run with bf16 storage
but fp32 gradients, sits 2-5 % from the all-fp32 oracle on these random-weight networks.  The CUDA step must be (a)
within 8e-2 of the same-storage oracle (gradients themselves are stored in bf16 between layers), (b) never more than twice as far from the fp32 oracle as the same-storage
oracle is (+1e-2).  Loss within 1e-3 relative, PSNR within 1e-3."""
import numpy as np
import pytest

from tests.helpers import O, rel_err

pytestmark = pytest.mark.gpu


def _setup(nb, sf, seed=1):
    from simplesr_b200 import model_builder as MB
    params = O.init_srresnet_params(seed=seed, bias_std=0.05, alpha_std=0.15, upsample_factor=sf, num_res_blocks=nb)
    m = MB.build_resnet(upsample_factor=sf, num_res_blocks=nb, batch_normalization=False, seed=0)
    weights = []
    for name, *_ in O.srresnet_layer_specs(upsample_factor=sf, num_res_blocks=nb):
        k, b, a = params[name]
        weights.extend([k, b] + ([a] if a is not None else []))
    m.set_weights(weights)
    return m, params


@pytest.mark.parametrize("nb,sf,shape", [(2, 2, (2, 12, 10)), (3, 4, (2, 8, 8))])
def test_gradients_match_oracle(nb, sf, shape):
    from simplesr_b200.training import SRResNetTrainer
    m, params = _setup(nb, sf)
    rng = np.random.default_rng(0)
    n, h, w = shape
    lr = rng.uniform(0, 1, size=(n, h, w, 3)).astype(np.float32)
    hr = rng.uniform(-1, 1, size=(n, h * sf, w * sf, 3)).astype(np.float32)
    tr = SRResNetTrainer(m, loss=("mse", 1.0), learning_rate=0.0)        # lr 0: inspect gradients, keep the weights
    out = tr.train_step(lr, hr)
    loss32, sr32, g32 = O.srresnet_loss_and_grads(params, lr, hr, upsample_factor=sf, num_res_blocks=nb)
    loss16, sr16, g16 = O.srresnet_loss_and_grads(params, lr, hr, upsample_factor=sf, num_res_blocks=nb,
                                                  act_dtype="bf16")
    assert abs(out["loss"] - loss32) <= 1e-3 * abs(loss32)
    np.testing.assert_allclose(out["psnr"], float(np.mean(O.psnr(hr, sr32, 2.0))), rtol=1e-3)
    got = tr.gradients()
    for name in g32:
        for i, kind in enumerate(("kernel", "bias", "alpha")):
            if g32[name][i] is None:
                continue
            e32, e16 = rel_err(got[name][i], g32[name][i]), rel_err(got[name][i], g16[name][i])
            inherent = rel_err(g16[name][i], g32[name][i])
            assert e16 <= 8e-2, (name, kind, e16)
            assert e32 <= 2 * inherent + 1e-2, (name, kind, e32, inherent)
    # eager launches == captured graph, bit for bit
    tr2 = SRResNetTrainer(_setup(nb, sf)[0], loss=("mse", 1.0), learning_rate=0.0)
    tr2.train_step(lr, hr, use_graph=False)
    g2 = tr2.gradients()
    for name in got:
        for a, b in zip(got[name], g2[name]):
            if a is not None:
                assert np.array_equal(a, b), name
    tr.release()
    tr2.release()


def test_three_adam_steps_follow_the_oracle():
    """Weights after 3 iterations with Keras Adam semantics (lr 1e-3): every variable moves like the oracle's."""
    from simplesr_b200.training import SRResNetTrainer
    nb, sf = 2, 2
    m, params = _setup(nb, sf)
    rng = np.random.default_rng(1)
    lr = rng.uniform(0, 1, size=(2, 10, 10, 3)).astype(np.float32)
    hr = rng.uniform(-1, 1, size=(2, 20, 20, 3)).astype(np.float32)
    tr = SRResNetTrainer(m, loss=[("mse", 1.0), ("mae", 0.01)], learning_rate=1e-3)
    p = {k: [np.array(v[0]), np.array(v[1]), None if v[2] is None else np.array(v[2])] for k, v in params.items()}
    mom = {k: [np.zeros_like(a) if a is not None else None for a in v] for k, v in p.items()}
    vel = {k: [np.zeros_like(a) if a is not None else None for a in v] for k, v in p.items()}
    losses = []
    for t in (1, 2, 3):
        out = tr.train_step(lr, hr)
        cur = {k: tuple(v) for k, v in p.items()}
        loss, sr, g = O.srresnet_loss_and_grads(cur, lr, hr, upsample_factor=sf, num_res_blocks=nb)
        mae = O.mean_absolute_error(hr, sr)
        # add the MAE term's gradient through a second backward with a sign-loss: reuse linearity via finite structure
        losses.append((out["loss"], float(loss + 0.01 * mae)))
        g_mae = _mae_grads(cur, lr, hr, sf, nb)
        for k in p:
            for i in range(3):
                if p[k][i] is None:
                    continue
                gi = g[k][i] + 0.01 * g_mae[k][i]
                p[k][i], mom[k][i], vel[k][i] = O.adam_update(p[k][i], gi, mom[k][i], vel[k][i], t, lr=1e-3)
    for a, b in losses:
        assert abs(a - b) <= 2e-3 * abs(b), losses
    assert losses[-1][0] < losses[0][0]                      # the loss goes down
    tv = {v.name: v.numpy() for v in m.trainable_variables}
    for k in p:
        moved = p[k][0] - params[k][0]
        got_moved = tv[f"{k}/kernel:0"] - params[k][0]
        # 3 Adam steps move every weight by about 3e-3 * sign(gradient): the displacement fields must line up
        # (weights whose tiny gradient changes sign under bf16 rounding are the only disagreement)
        cos = float((moved * got_moved).sum() / (np.linalg.norm(moved) * np.linalg.norm(got_moved) + 1e-30))
        assert cos > 0.9, (k, cos)
        assert np.abs(got_moved).max() <= 3.5e-3
    tr.release()


def _mae_grads(params, lr, hr, sf, nb):
    """Gradients of mean|sr - hr| through the oracle: backprop sign(sr-hr)/N by reusing the MSE path with a target
    shifted so that 2*(sr-hr')/N equals sign(sr-hr)/N."""
    _, sr, _ = O.srresnet_loss_and_grads(params, lr, hr, upsample_factor=sf, num_res_blocks=nb)
    hr2 = sr - 0.5 * np.sign(sr - hr)
    _, _, g = O.srresnet_loss_and_grads(params, lr, hr2.astype(np.float32), upsample_factor=sf, num_res_blocks=nb)
    return g


def _setup_rrdb(nb, sf, seed=1):
    from simplesr_b200 import model_builder as MB
    params = O.init_rrdb_params(seed=seed, bias_std=0.05, upsample_factor=sf, num_rrdb_blocks=nb)
    m = MB.build_enhanced_resnet(upsample_factor=sf, num_rrdb_blocks=nb, seed=0)
    weights = []
    for name, _, _ in O.rrdb_layer_specs(upsample_factor=sf, num_rrdb_blocks=nb):
        weights.extend(params[name])
    m.set_weights(weights)
    return m, params


@pytest.mark.parametrize("nb,sf,shape", [(1, 2, (2, 12, 10)), (2, 4, (1, 8, 8))])
def test_rrdb_gradients_match_oracle(nb, sf, shape):
    """RRDB generator with MSE + 0.1 MAE: every kernel / bias gradient of the dense blocks (in-place 192-channel
    gradient assembly), trunk, up-convs and tail against the oracle (same tolerance scheme as SRResNet)."""
    from simplesr_b200.training import RRDBTrainer
    m, params = _setup_rrdb(nb, sf)
    rng = np.random.default_rng(0)
    n, h, w = shape
    lr = rng.uniform(0, 1, size=(n, h, w, 3)).astype(np.float32)
    hr = rng.uniform(-1, 1, size=(n, h * sf, w * sf, 3)).astype(np.float32)
    tr = RRDBTrainer(m, loss=[("mse", 1.0), ("mae", 0.1)], learning_rate=0.0)
    out = tr.train_step(lr, hr)
    kw = dict(upsample_factor=sf, num_rrdb_blocks=nb, w_mse=1.0, w_mae=0.1)
    loss32, sr32, g32 = O.rrdb_loss_and_grads(params, lr, hr, **kw)
    loss16, sr16, g16 = O.rrdb_loss_and_grads(params, lr, hr, act_dtype="bf16", **kw)
    assert abs(out["loss"] - loss32) <= 1e-3 * abs(loss32)
    got = tr.gradients()
    worst = 0.0
    for name in g32:
        for i, kind in enumerate(("kernel", "bias")):
            e32, e16 = rel_err(got[name][i], g32[name][i]), rel_err(got[name][i], g16[name][i])
            inherent = rel_err(g16[name][i], g32[name][i])
            worst = max(worst, e16)
            assert e16 <= 8e-2, (name, kind, e16)
            assert e32 <= 2 * inherent + 1e-2, (name, kind, e32, inherent)
    tr.release()


def test_rrdb_training_reduces_the_loss():
    from simplesr_b200.training import RRDBTrainer
    m, _ = _setup_rrdb(1, 2)
    rng = np.random.default_rng(2)
    lr = rng.uniform(0, 1, size=(2, 12, 12, 3)).astype(np.float32)
    hr = np.clip(np.repeat(np.repeat(lr, 2, 1), 2, 2) * 2 - 1, -1, 1).astype(np.float32)   # a learnable target
    tr = RRDBTrainer(m, loss=("mse", 1.0), learning_rate=1e-3)
    losses = [tr.train_step(lr, hr)["loss"] for _ in range(12)]
    assert losses[-1] < 0.8 * losses[0], losses
    tr.release()


def test_rrdb_fused_activation_backward_is_bit_identical():
    """ssr_conv2d_fwd_mask: LeakyReLU backward of the growth convs fused into the dgrad epilogues (RRDBTrainer.fuse_act_bwd)
    must give the same gradients, bit for bit, as the separate ssr_act_bwd_bf16 launches (same bf16 values are masked),
    with and without the side-stream overlap of the weight gradients."""
    from simplesr_b200.training import RRDBTrainer
    rng = np.random.default_rng(5)
    lr = rng.uniform(0, 1, size=(2, 12, 10, 3)).astype(np.float32)
    hr = rng.uniform(-1, 1, size=(2, 24, 20, 3)).astype(np.float32)
    grads = []
    for fuse, overlap in ((False, False), (True, True), (False, True)):
        m, _ = _setup_rrdb(1, 2)
        tr = RRDBTrainer(m, loss=[("mse", 1.0), ("mae", 0.1)], learning_rate=0.0)
        tr.fuse_act_bwd, tr.overlap_wgrad = fuse, overlap
        tr.train_step(lr, hr)
        grads.append(tr.gradients())
        tr.release()
        m.release()
    for other in grads[1:]:
        for name in grads[0]:
            for a, b in zip(grads[0][name], other[name]):
                if a is not None:
                    assert np.array_equal(a, b), name


def _setup_bn(nb, sf, seed=7):
    from simplesr_b200 import model_builder as MB
    params = O.init_srresnet_params(seed=seed, bias_std=0.05, alpha_std=0.15, upsample_factor=sf, num_res_blocks=nb)
    bn = O.init_srresnet_bn(seed=seed, num_res_blocks=nb, randomize=True)
    m = MB.build_resnet(upsample_factor=sf, num_res_blocks=nb, seed=0)             # batch_normalization=True (default)
    weights, moving = [], []
    for name, *_ in O.srresnet_layer_specs(upsample_factor=sf, num_res_blocks=nb):
        k, b, a = params[name]
        weights.extend([k, b])
        if name in bn:
            weights.extend([bn[name]["gamma"], bn[name]["beta"]])
            moving.extend([bn[name]["mean"], bn[name]["var"]])
        if a is not None:
            weights.append(a)
    m.set_weights(weights + moving)
    return m, params, bn


def test_srresnet_batch_norm_training_step():
    """build_resnet(batch_normalization=True) trained on the GPU: batch statistics in the forward pass, the full
    BatchNormalization backward (dgamma, dbeta, dz), the moving-average update (momentum 0.8, unbiased variance as TF's
    fused batch norm), and inference afterwards with the updated moving statistics folded into the convs.
    Same tolerance scheme as the batch-norm-free test above."""
    from simplesr_b200.training import SRResNetTrainer
    nb, sf = 2, 2
    m, params, bn = _setup_bn(nb, sf)
    rng = np.random.default_rng(3)
    n, h, w = 2, 12, 10
    lr = rng.uniform(0, 1, size=(n, h, w, 3)).astype(np.float32)
    hr = rng.uniform(-1, 1, size=(n, h * sf, w * sf, 3)).astype(np.float32)
    tr = SRResNetTrainer(m, loss=("mse", 1.0), learning_rate=0.0)
    out = tr.train_step(lr, hr)
    stats = {}
    kw = dict(upsample_factor=sf, num_res_blocks=nb, bn=bn)
    loss32, sr32, g32 = O.srresnet_loss_and_grads(params, lr, hr, stats_out=stats, **kw)
    loss16, sr16, g16 = O.srresnet_loss_and_grads(params, lr, hr, act_dtype="bf16", **kw)
    assert abs(out["loss"] - loss32) <= 2e-3 * abs(loss32)
    got = tr.gradients()
    for name in g32:
        for i, kind in enumerate(("kernel", "bias", "alpha") if not name.endswith("_bn") else ("gamma", "beta")):
            if g32[name][i] is None or (kind == "bias" and name in bn):
                continue          # a bias in front of a batch norm has a zero gradient (pure rounding noise)
            e32, e16 = rel_err(got[name][i], g32[name][i]), rel_err(got[name][i], g16[name][i])
            inherent = rel_err(g16[name][i], g32[name][i])
            # batch norm rescales by 1/std of bf16-rounded conv outputs: on this random-weight network the storage-induced
            # distance itself reaches 20 % on one layer (up0), so the same-storage bound scales with it there
            assert e16 <= max(8e-2, 0.6 * inherent), (name, kind, e16, inherent)
            assert e32 <= 2 * inherent + 1e-2, (name, kind, e32, inherent)
    # moving statistics: new = 0.8 * old + 0.2 * batch (variance with Bessel's correction)
    cnt = n * h * w
    mv = {v.name: v.numpy() for v in m.non_trainable_variables}
    for name in bn:
        mu, var = stats[name]
        np.testing.assert_allclose(mv[f"{name}_bn/moving_mean:0"], 0.8 * bn[name]["mean"] + 0.2 * mu, rtol=2e-2, atol=2e-3)
        np.testing.assert_allclose(mv[f"{name}_bn/moving_variance:0"], 0.8 * bn[name]["var"] + 0.2 * var * cnt / (cnt - 1),
                                   rtol=2e-2, atol=2e-3)
    # inference after the step uses the updated moving statistics (folded into the convs)
    bn2 = {k: dict(gamma=v["gamma"], beta=v["beta"], mean=mv[f"{k}_bn/moving_mean:0"], var=mv[f"{k}_bn/moving_variance:0"])
           for k, v in bn.items()}
    x = rng.uniform(0, 1, size=(1, 16, 12, 3)).astype(np.float32)
    got_y = m(x, training=False)
    ref_y = O.srresnet_forward(params, x, upsample_factor=sf, num_res_blocks=nb, bn=bn2)
    assert float(O.psnr(got_y, ref_y, max_val=2.0).min()) > 50.0
    tr.release()
    m.release()


def test_c3_configuration_step():
    """BASELINE.json configs[2] at its real size: SRResNet-16 x4, MSE, batch 16 of 96x96 HR (24x24 LR).  Loss, PSNR metric
    and the gradients nearest to the loss (last / up convs) against the oracle; the deep-layer gradients of an untrained
    16-block network amplify bf16 rounding beyond any useful bound (see test_gpu_srresnet.py) and are covered at depth 2-3
    above."""
    from simplesr_b200.training import SRResNetTrainer
    nb, sf = 16, 4
    m, params = _setup(nb, sf, seed=11)
    for b in range(nb):                       # small residual updates, as in a trained network
        k, bb, a = params[f"res{b}_conv1"]
        params[f"res{b}_conv1"] = (k * np.float32(0.25), bb, a)
    weights = []
    for name, *_ in O.srresnet_layer_specs(upsample_factor=sf, num_res_blocks=nb):
        k, b, a = params[name]
        weights.extend([k, b] + ([a] if a is not None else []))
    m.set_weights(weights)
    rng = np.random.default_rng(4)
    lr = rng.uniform(0, 1, size=(16, 24, 24, 3)).astype(np.float32)
    hr = rng.uniform(-1, 1, size=(16, 96, 96, 3)).astype(np.float32)
    tr = SRResNetTrainer(m, loss=("mse", 1.0), learning_rate=0.0)
    out = tr.train_step(lr, hr)
    loss32, sr32, g32 = O.srresnet_loss_and_grads(params, lr, hr, upsample_factor=sf, num_res_blocks=nb)
    assert abs(out["loss"] - loss32) <= 2e-3 * abs(loss32)
    np.testing.assert_allclose(out["psnr"], float(np.mean(O.psnr(hr, sr32, 2.0))), rtol=2e-3)
    got = tr.gradients()
    for name in ("last", "up1", "up0"):
        assert rel_err(got[name][0], g32[name][0]) <= 5e-2, (name, rel_err(got[name][0], g32[name][0]))
    assert all(np.isfinite(g[0]).all() for g in got.values())
    tr.release()
