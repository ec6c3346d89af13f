"""The oracle's backward pass against torch autograd (an independent implementation), CPU only."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import ssr_oracle as O


def test_conv_backward_matches_autograd():
    rng = np.random.default_rng(0)
    x = rng.standard_normal((2, 7, 9, 5)).astype(np.float32)
    k = rng.standard_normal((3, 3, 5, 4)).astype(np.float32)
    dy = rng.standard_normal((2, 7, 9, 4)).astype(np.float32)
    dx, dk, db = O.conv2d_same_backward(x, k, dy)
    xt = torch.from_numpy(x).permute(0, 3, 1, 2).requires_grad_(True)
    kt = torch.from_numpy(k).permute(3, 2, 0, 1).requires_grad_(True)
    bt = torch.zeros(4, requires_grad=True)
    y = F.conv2d(xt, kt, bt, padding=1)
    y.backward(torch.from_numpy(dy).permute(0, 3, 1, 2))
    np.testing.assert_allclose(dx, xt.grad.permute(0, 2, 3, 1).numpy(), rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(dk, kt.grad.permute(2, 3, 1, 0).numpy(), rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(db, bt.grad.numpy(), rtol=1e-4, atol=1e-4)


def test_space_to_depth_inverts_depth_to_space():
    x = np.random.default_rng(1).standard_normal((2, 3, 4, 8)).astype(np.float32)
    np.testing.assert_array_equal(O.space_to_depth(O.depth_to_space(x, 2), 2), x)


def _torch_srresnet(params, x, nb, sf, bn=None):
    """Independent torch restatement (NCHW, autograd) of build_resnet; ``bn``: training-mode batch norm after the res
    block convs and the trunk conv (F.batch_norm with batch statistics, eps 1e-3)."""
    P = {}
    BN = {}
    if bn is not None:
        for name, v in bn.items():
            BN[name] = (torch.tensor(v["gamma"]).clone().requires_grad_(True),
                        torch.tensor(v["beta"]).clone().requires_grad_(True))
    for name, (k, b, a) in params.items():
        P[name] = (torch.tensor(k).permute(3, 2, 0, 1).clone().requires_grad_(True),
                   torch.tensor(b).clone().requires_grad_(True),
                   torch.tensor(a).clone().requires_grad_(True) if a is not None else None)

    def conv(name, t):
        k, b, _ = P[name]
        y = F.conv2d(t, k, b, padding=k.shape[-1] // 2)
        if name in BN:
            y = F.batch_norm(y, None, None, BN[name][0], BN[name][1], training=True, eps=1e-3)
        return y

    def prelu(name, t):
        return F.prelu(t, P[name][2])

    def d2s(t):   # TF DCR order expressed with torch ops
        n, c4, h, w = t.shape
        c = c4 // 4
        return t.reshape(n, 2, 2, c, h, w).permute(0, 3, 4, 1, 5, 2).reshape(n, c, 2 * h, 2 * w)

    t = prelu("first", conv("first", x))
    skip = t
    for b in range(nb):
        u = prelu(f"res{b}_conv0", conv(f"res{b}_conv0", t))
        t = t + conv(f"res{b}_conv1", u)
    t = conv("trunk", t) + skip
    for i in range(int(np.log2(sf))):
        t = prelu(f"up{i}", d2s(conv(f"up{i}", t)))
    if bn is not None:
        return torch.tanh(conv("last", t)), P, BN
    return torch.tanh(conv("last", t)), P


def test_srresnet_gradients_match_autograd():
    nb, sf = 2, 2
    params = O.init_srresnet_params(seed=3, bias_std=0.05, alpha_std=0.2, upsample_factor=sf, num_res_blocks=nb)
    rng = np.random.default_rng(0)
    lr = rng.uniform(0, 1, size=(2, 6, 5, 3)).astype(np.float32)
    hr = rng.uniform(-1, 1, size=(2, 12, 10, 3)).astype(np.float32)
    loss, sr, grads = O.srresnet_loss_and_grads(params, lr, hr, upsample_factor=sf, num_res_blocks=nb)
    srt, P = _torch_srresnet(params, torch.tensor(lr).permute(0, 3, 1, 2), nb, sf)
    lt = F.mse_loss(srt, torch.tensor(hr).permute(0, 3, 1, 2))
    lt.backward()
    np.testing.assert_allclose(loss, lt.item(), rtol=1e-5)
    np.testing.assert_allclose(sr, srt.detach().permute(0, 2, 3, 1).numpy(), rtol=1e-4, atol=1e-5)
    for name, (dk, db, da) in grads.items():
        k, b, a = P[name]
        scale = max(float(k.grad.abs().max()), 1e-8)
        np.testing.assert_allclose(dk, k.grad.permute(2, 3, 1, 0).numpy(), rtol=2e-3, atol=2e-4 * scale, err_msg=name)
        np.testing.assert_allclose(db, b.grad.numpy(), rtol=2e-3, atol=1e-6, err_msg=name)
        if a is not None:
            np.testing.assert_allclose(da, a.grad.numpy(), rtol=2e-3, atol=1e-6, err_msg=name)


def test_adam_matches_torch_adam_with_keras_epsilon_placement():
    """Keras: p -= lr_t * m / (sqrt(v) + eps) with lr_t = lr sqrt(1-b2^t)/(1-b1^t) (epsilon is NOT bias corrected)."""
    rng = np.random.default_rng(0)
    p = rng.standard_normal(50).astype(np.float32)
    g = rng.standard_normal(50).astype(np.float32)
    m = np.zeros_like(p)
    v = np.zeros_like(p)
    p1, m1, v1 = O.adam_update(p, g, m, v, 1, lr=1e-3)
    np.testing.assert_allclose(m1, 0.1 * g, rtol=1e-5)
    np.testing.assert_allclose(v1, 0.001 * g * g, rtol=1e-4)
    lr_t = 1e-3 * np.sqrt(1 - 0.999) / (1 - 0.9)
    np.testing.assert_allclose(p1, p - lr_t * m1 / (np.sqrt(v1) + 1e-7), rtol=1e-5)
    # first step moves every weight by ~lr in the direction of -sign(g)
    np.testing.assert_allclose(p1 - p, -1e-3 * np.sign(g), rtol=1e-3)


def _torch_rrdb(params, x, nb, sf):
    P = {n_: (torch.tensor(k).permute(3, 2, 0, 1).clone().requires_grad_(True), torch.tensor(b).clone().requires_grad_(True))
         for n_, (k, b) in params.items()}
    conv = lambda name, t: F.conv2d(t, P[name][0], P[name][1], padding=1)
    lre = lambda t: F.leaky_relu(t, 0.2)

    def d2s(t):
        n, c4, h, w = t.shape
        c = c4 // 4
        return t.reshape(n, 2, 2, c, h, w).permute(0, 3, 4, 1, 5, 2).reshape(n, c, 2 * h, 2 * w)

    fea = conv("fea", x)
    r = fea
    for b in range(nb):
        for d in range(3):
            feats = [r]
            for k in range(4):
                feats.append(lre(conv(f"rrdb{b}_db{d}_conv{k}", torch.cat(feats, 1))))
            r = r + 0.2 * conv(f"rrdb{b}_db{d}_out", torch.cat(feats, 1))
    u = fea + conv("trunk", fea + 0.2 * r)
    for i in range(int(np.log2(sf))):
        u = lre(d2s(conv(f"up{i}", u)))
    return torch.tanh(conv("last", lre(conv("hr", u)))), P


def test_rrdb_gradients_match_autograd():
    nb, sf = 1, 2
    params = O.init_rrdb_params(seed=3, bias_std=0.05, upsample_factor=sf, num_rrdb_blocks=nb)
    rng = np.random.default_rng(0)
    lr = rng.uniform(0, 1, size=(2, 6, 5, 3)).astype(np.float32)
    hr = rng.uniform(-1, 1, size=(2, 12, 10, 3)).astype(np.float32)
    loss, sr, grads = O.rrdb_loss_and_grads(params, lr, hr, upsample_factor=sf, num_rrdb_blocks=nb, w_mse=1.0, w_mae=0.1)
    srt, P = _torch_rrdb(params, torch.tensor(lr).permute(0, 3, 1, 2), nb, sf)
    hrt = torch.tensor(hr).permute(0, 3, 1, 2)
    lt = F.mse_loss(srt, hrt) + 0.1 * F.l1_loss(srt, hrt)
    lt.backward()
    np.testing.assert_allclose(loss, lt.item(), rtol=1e-5)
    for name, (dk, db) in grads.items():
        k, b = P[name]
        scale = max(float(k.grad.abs().max()), 1e-10)
        np.testing.assert_allclose(dk, k.grad.permute(2, 3, 1, 0).numpy(), rtol=2e-3, atol=2e-4 * scale, err_msg=name)
        np.testing.assert_allclose(db, b.grad.numpy(), rtol=2e-3, atol=2e-4 * max(float(b.grad.abs().max()), 1e-10),
                                   err_msg=name)


def test_vgg_loss_and_gradient_match_autograd():
    """Truncated VGG19 (to block2_conv2 to keep it small), pre-activation features, caffe preprocessing."""
    params = O.init_vgg19_params(seed=2)
    rng = np.random.default_rng(0)
    hr = rng.uniform(-1, 1, size=(2, 12, 16, 3)).astype(np.float32)
    sr = rng.uniform(-1, 1, size=(2, 12, 16, 3)).astype(np.float32)
    layer = "block2_conv2"
    loss, dsr = O.vgg_loss_and_grad(params, hr, sr, output_layer=layer, loss_weight=0.7)

    def feats(x):
        t = (x + 1) * 127.5
        t = t[..., [2, 1, 0]] - torch.tensor(O.VGG_MEAN_BGR)
        t = t.permute(0, 3, 1, 2)
        for lay in O.VGG19_LAYERS:
            name = lay[0]
            if len(lay) == 3:
                k, b = params[name]
                z = F.conv2d(t, torch.tensor(k).permute(3, 2, 0, 1), torch.tensor(b), padding=1)
                if name == layer:
                    return z
                t = F.relu(z)
            else:
                t = F.max_pool2d(t, 2)

    srt = torch.tensor(sr, requires_grad=True)
    lt = 0.7 * F.mse_loss(feats(srt), feats(torch.tensor(hr)))
    lt.backward()
    np.testing.assert_allclose(loss, lt.item(), rtol=1e-4)
    np.testing.assert_allclose(dsr, srt.grad.numpy(), rtol=2e-3, atol=2e-4 * float(srt.grad.abs().max()))
    assert O.vgg_preprocess(np.zeros((1, 1, 1, 3), np.float32)).ravel().tolist() == pytest.approx(
        [127.5 - 103.939, 127.5 - 116.779, 127.5 - 123.68], rel=1e-6)


@pytest.mark.parametrize("after_activation,layers,tv", [(True, ["block2_conv2"], False),
                                                         (True, ["block1_conv2", "block2_conv2"], True),
                                                         (False, ["block1_conv2", "block2_conv1"], True)])
def test_vgg_loss_general_matches_autograd(after_activation, layers, tv):
    """vgg_loss.py:150-169 in full: several output layers summed, features after / before the ReLU, and the
    total-variation term on the de-normalised SR batch (an independent torch restatement with autograd)."""
    params = O.init_vgg19_params(seed=2)
    rng = np.random.default_rng(0)
    hr = rng.uniform(-1, 1, size=(2, 12, 16, 3)).astype(np.float32)
    sr = rng.uniform(-1, 1, size=(2, 12, 16, 3)).astype(np.float32)
    loss, dsr = O.vgg_loss_general(params, hr, sr, layers, feature_scale=0.5, loss_weight=0.7,
                                   after_activation=after_activation, total_variation_loss=tv,
                                   total_variation_weight=1e-4)

    def feats(x):
        t = (x + 1) * 127.5
        t = t[..., [2, 1, 0]] - torch.tensor(O.VGG_MEAN_BGR)
        t = t.permute(0, 3, 1, 2)
        out = []
        for lay in O.VGG19_LAYERS:
            name = lay[0]
            if len(lay) == 3:
                k, b = params[name]
                z = F.conv2d(t, torch.tensor(k).permute(3, 2, 0, 1), torch.tensor(b), padding=1)
                t = F.relu(z)
                if name in layers:
                    out.append(t if after_activation else z)
                if len(out) == len(layers):
                    return out
            else:
                t = F.max_pool2d(t, 2)

    srt = torch.tensor(sr, requires_grad=True)
    lt = sum(0.7 * F.mse_loss(a * 0.5, b * 0.5) for a, b in zip(feats(srt), feats(torch.tensor(hr))))
    if tv:
        den = (srt + 1) * 127.5
        lt = lt + 1e-4 * ((den[:, 1:] - den[:, :-1]).abs().sum() + (den[:, :, 1:] - den[:, :, :-1]).abs().sum())
    lt.backward()
    np.testing.assert_allclose(loss, lt.item(), rtol=1e-4)
    np.testing.assert_allclose(dsr, srt.grad.numpy(), rtol=2e-3, atol=2e-4 * float(srt.grad.abs().max()))
    if len(layers) == 1 and not tv and not after_activation:
        ref = O.vgg_loss_and_grad(params, hr, sr, output_layer=layers[0], feature_scale=0.5, loss_weight=0.7)
        np.testing.assert_allclose(loss, ref[0], rtol=1e-6)


def _torch_disc(params, x, hw):
    P = {k: [torch.tensor(v[0]).clone().requires_grad_(True), torch.tensor(v[1]).clone().requires_grad_(True)]
         for k, v in params.items()}
    t = x.permute(0, 3, 1, 2)
    for name, _, _, stride, bn in O.DISC_CONVS:
        k, b = P[name]
        if stride == 2:
            t = F.pad(t, (0, 1, 0, 1))      # TF SAME, even size, k=3, s=2: pad bottom/right only
            t = F.conv2d(t, k.permute(3, 2, 0, 1), b, stride=2)
        else:
            t = F.conv2d(t, k.permute(3, 2, 0, 1), b, padding=1)
        if bn:
            g, be = P[name + "_bn"]
            t = F.batch_norm(t, None, None, g, be, training=True, eps=1e-3)
        t = F.leaky_relu(t, 0.2)
    f = t.permute(0, 2, 3, 1).reshape(t.shape[0], -1)
    h = F.leaky_relu(f @ P["d_dense0"][0] + P["d_dense0"][1], 0.2)
    return h @ P["d_dense1"][0] + P["d_dense1"][1], P


def test_discriminator_and_ragan_match_autograd():
    hw = (32, 32)
    params = O.init_discriminator_params(seed=3, input_hw=hw, bias_std=0.05)
    for k in list(params):
        if k.endswith("_bn"):
            rng = np.random.default_rng(len(k))
            params[k] = [(1 + 0.1 * rng.standard_normal(params[k][0].shape)).astype(np.float32),
                         (0.1 * rng.standard_normal(params[k][1].shape)).astype(np.float32)]
    rng = np.random.default_rng(0)
    hr = rng.uniform(-1, 1, size=(4, *hw, 3)).astype(np.float32)
    sr = rng.uniform(-1, 1, size=(4, *hw, 3)).astype(np.float32)
    ch, cs = {}, {}
    hc = O.discriminator_forward(params, hr, cache=ch)
    sc = O.discriminator_forward(params, sr, cache=cs)
    L = O.ragan_losses(hc, sc)
    srt = torch.tensor(sr, requires_grad=True)
    sct, P = _torch_disc(params, srt, hw)
    hct, P2 = _torch_disc(params, torch.tensor(hr), hw)
    np.testing.assert_allclose(sc, sct.detach().numpy(), rtol=2e-3, atol=2e-5)
    bce = torch.nn.BCEWithLogitsLoss()
    g_loss = 0.5 * (bce(hct - sct.mean(), torch.zeros_like(hct)) + bce(sct - hct.mean(), torch.ones_like(sct)))
    np.testing.assert_allclose(L["g_loss"], g_loss.item(), rtol=1e-5)
    # generator side: gradient w.r.t. the SR image through D(sr)
    g_loss.backward(retain_graph=True)
    dx, _ = O.discriminator_backward(params, cs, L["g_dsr"])
    np.testing.assert_allclose(dx, srt.grad.numpy(), rtol=5e-3, atol=5e-3 * float(srt.grad.abs().max()))
    # discriminator side: weight gradients through both critic passes (shared weights: two torch copies summed)
    for pp in (P, P2):
        for v in pp.values():
            for t_ in v:
                t_.grad = None
    d_loss = 0.5 * (bce(hct - sct.mean(), torch.ones_like(hct)) + bce(sct - hct.mean(), torch.zeros_like(sct)))
    np.testing.assert_allclose(L["d_loss"], d_loss.item(), rtol=1e-5)
    d_loss.backward()
    _, gs = O.discriminator_backward(params, cs, L["d_dsr"])
    _, gh = O.discriminator_backward(params, ch, L["d_dhr"])
    for name in gs:
        for i in range(2):
            ref = P[name][i].grad.numpy() + P2[name][i].grad.numpy()
            if i == 1 and name != "d_conv0" and name.startswith("d_conv") and not name.endswith("_bn"):
                continue     # the bias of a conv that feeds BatchNormalization has an exactly-zero gradient (noise only)
            got = gs[name][i] + gh[name][i]
            # (the bias of a conv that feeds BatchNormalization has an exactly-zero gradient: only rounding noise)
            np.testing.assert_allclose(got.reshape(ref.shape), ref, rtol=5e-3, atol=5e-3 * max(float(np.abs(ref).max()), 1e-6),
                                       err_msg=f"{name}[{i}]")


def test_srresnet_batch_norm_gradients_match_autograd():
    """build_resnet(batch_normalization=True) in training mode: the oracle's forward / backward through the 2*nb+1
    BatchNormalization layers (batch statistics, biased variance, eps 1e-3) against torch autograd."""
    nb, sf = 2, 2
    params = O.init_srresnet_params(seed=5, bias_std=0.05, alpha_std=0.2, upsample_factor=sf, num_res_blocks=nb)
    bn = O.init_srresnet_bn(seed=5, num_res_blocks=nb, randomize=True)
    rng = np.random.default_rng(1)
    lr = rng.uniform(0, 1, size=(2, 6, 5, 3)).astype(np.float32)
    hr = rng.uniform(-1, 1, size=(2, 12, 10, 3)).astype(np.float32)
    stats = {}
    loss, sr, grads = O.srresnet_loss_and_grads(params, lr, hr, upsample_factor=sf, num_res_blocks=nb, bn=bn, stats_out=stats)
    srt, P, BN = _torch_srresnet(params, torch.tensor(lr).permute(0, 3, 1, 2), nb, sf, bn=bn)
    lt = F.mse_loss(srt, torch.tensor(hr).permute(0, 3, 1, 2))
    lt.backward()
    np.testing.assert_allclose(loss, lt.item(), rtol=1e-5)
    np.testing.assert_allclose(sr, srt.detach().permute(0, 2, 3, 1).numpy(), rtol=1e-4, atol=1e-5)
    assert set(stats) == set(bn)
    for name in bn:
        dgamma, dbeta = grads[name + "_bn"]
        np.testing.assert_allclose(dgamma, BN[name][0].grad.numpy(), rtol=2e-3, atol=2e-6)
        np.testing.assert_allclose(dbeta, BN[name][1].grad.numpy(), rtol=2e-3, atol=2e-6)
    for name, g in grads.items():
        if name.endswith("_bn"):
            continue
        dk, db, da = g
        k, b, a = P[name]
        np.testing.assert_allclose(dk, k.grad.permute(2, 3, 1, 0).numpy(), rtol=2e-3, atol=2e-6)
        if name not in bn:   # a bias in front of a batch norm has zero gradient (rounding noise on both sides)
            np.testing.assert_allclose(db, b.grad.numpy(), rtol=2e-3, atol=2e-6)


def test_standard_gan_losses_match_autograd():
    """oracle.gan_losses (sigmoid critic + Keras BinaryCrossentropy on probabilities) against torch autograd."""
    import torch
    rng = np.random.default_rng(5)
    n = 6
    zh, zs = rng.normal(0.5, 2.0, n), rng.normal(-0.5, 2.0, n)
    lh, ls = 0.7 + 0.5 * rng.uniform(size=n), 0.3 * rng.uniform(size=n)
    R = O.gan_losses(zh, zs, hr_label=lh, sr_label=ls)
    th, ts = torch.tensor(zh, requires_grad=True), torch.tensor(zs, requires_grad=True)
    eps = 1e-7

    def bce(y, z):
        p = torch.clamp(torch.sigmoid(z), eps, 1 - eps)
        return torch.mean(-(y * torch.log(p + eps) + (1 - y) * torch.log(1 - p + eps)))

    g = bce(torch.ones(n, dtype=torch.float64), ts)
    d = bce(torch.tensor(ls), ts) + bce(torch.tensor(lh), th)
    np.testing.assert_allclose(R["g_loss"], g.item(), rtol=1e-6)
    np.testing.assert_allclose(R["d_loss"], d.item(), rtol=1e-6)
    (g_dsr,) = torch.autograd.grad(g, ts, retain_graph=True)
    d_dhr, d_dsr = torch.autograd.grad(d, [th, ts])
    np.testing.assert_allclose(R["g_dsr"].ravel(), g_dsr.numpy(), rtol=1e-5, atol=1e-9)
    np.testing.assert_allclose(R["d_dsr"].ravel(), d_dsr.numpy(), rtol=1e-5, atol=1e-9)
    np.testing.assert_allclose(R["d_dhr"].ravel(), d_dhr.numpy(), rtol=1e-5, atol=1e-9)
    # away from saturation it is the logits form TensorFlow's graph mode substitutes for a Sigmoid producer
    sp = lambda z: np.maximum(z, 0) + np.log1p(np.exp(-np.abs(z)))
    np.testing.assert_allclose(R["g_loss"], np.mean(sp(zs) - zs), rtol=1e-5)
