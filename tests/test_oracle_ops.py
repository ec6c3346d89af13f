"""The oracle's op semantics against independent implementations (CPU only).

The reference holds no golden tensors for conv / generator numerics (parity unpinned, see oracle header), so the
numpy restatement is cross-checked against torch.nn.functional (a different code base) and hand examples.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import ssr_oracle as O


def _torch_conv_same(x, k, b, stride):
    """TF SAME padding reproduced with explicit asymmetric padding, then a VALID torch conv."""
    n, h, w, c = x.shape
    kh, kw = k.shape[:2]
    _, pt, pb = O.same_padding(h, kh, stride)
    _, pl, pr = O.same_padding(w, kw, stride)
    xt = torch.from_numpy(x).permute(0, 3, 1, 2)
    xt = F.pad(xt, (pl, pr, pt, pb))
    wt = torch.from_numpy(k).permute(3, 2, 0, 1)
    y = F.conv2d(xt, wt, torch.from_numpy(b), stride=stride)
    return y.permute(0, 2, 3, 1).numpy()


@pytest.mark.parametrize("ks,stride,h,w", [(3, 1, 9, 11), (9, 1, 12, 10), (3, 2, 8, 8), (3, 2, 7, 9), (1, 1, 5, 5)])
def test_conv2d_same_matches_torch(ks, stride, h, w):
    rng = np.random.default_rng(ks * 10 + stride)
    x = rng.standard_normal((2, h, w, 5)).astype(np.float32)
    k = rng.standard_normal((ks, ks, 5, 7)).astype(np.float32)
    b = rng.standard_normal(7).astype(np.float32)
    got = O.conv2d_same(x, k, b, stride=stride)
    ref = _torch_conv_same(x, k, b, stride)
    assert got.shape == ref.shape == (2, -(-h // stride), -(-w // stride), 7)
    np.testing.assert_allclose(got, ref, rtol=1e-4, atol=1e-4)


def test_same_padding_stride2_even_input_pads_bottom_right_only():
    """TF SAME for k=3, s=2, even size: pad (0, 1) - differs from PyTorch padding=1."""
    assert O.same_padding(8, 3, 2) == (4, 0, 1)
    assert O.same_padding(7, 3, 2) == (4, 1, 1)
    assert O.same_padding(128, 3, 1) == (128, 1, 1)
    assert O.same_padding(128, 9, 1) == (128, 4, 4)


def test_depth_to_space_is_tf_dcr_order():
    """tf.nn.depth_to_space NHWC: out[n,2h+i,2w+j,c] = in[n,h,w,(2i+j)*C+c]; TF's documented example."""
    x = np.arange(1, 17, dtype=np.float32).reshape(1, 2, 2, 4)
    # TensorFlow docs: [[[[1,2,3,4],[5,6,7,8]],[[9,10,11,12],[13,14,15,16]]]] -> 4x4x1
    ref = np.array([[1, 2, 5, 6], [3, 4, 7, 8], [9, 10, 13, 14], [11, 12, 15, 16]], np.float32).reshape(1, 4, 4, 1)
    np.testing.assert_array_equal(O.depth_to_space(x, 2), ref)
    rng = np.random.default_rng(0)
    x = rng.standard_normal((2, 3, 5, 12)).astype(np.float32)
    y = O.depth_to_space(x, 2)
    for i in range(2):
        for j in range(2):
            np.testing.assert_array_equal(y[:, i::2, j::2, :], x[..., (2 * i + j) * 3:(2 * i + j + 1) * 3])
    # NOT PyTorch pixel_shuffle order (channel = c*4 + 2i + j)
    ps = F.pixel_shuffle(torch.from_numpy(x).permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1).numpy()
    assert not np.array_equal(y, ps)


def test_activations():
    x = np.array([-2.0, -0.5, 0.0, 0.5, 3.0], np.float32)
    np.testing.assert_allclose(O.leaky_relu(x, 0.2), [-0.4, -0.1, 0.0, 0.5, 3.0], rtol=1e-6)
    np.testing.assert_allclose(O.prelu(x, np.float32(0.25)), [-0.5, -0.125, 0.0, 0.5, 3.0], rtol=1e-6)
    np.testing.assert_allclose(O.leaky_relu(x, 0.2), F.leaky_relu(torch.from_numpy(x), 0.2).numpy(), rtol=1e-6)


def test_losses_and_psnr():
    """Keras MSE/MAE = global mean for equal shapes (§9.9); tf.image.psnr per image (§9.10, metrics.py:4-15)."""
    rng = np.random.default_rng(1)
    a = rng.uniform(-1, 1, size=(3, 8, 8, 3)).astype(np.float32)
    b = rng.uniform(-1, 1, size=(3, 8, 8, 3)).astype(np.float32)
    np.testing.assert_allclose(O.mean_squared_error(a, b), F.mse_loss(torch.from_numpy(a), torch.from_numpy(b)).item(),
                               rtol=1e-6)
    np.testing.assert_allclose(O.mean_absolute_error(a, b), F.l1_loss(torch.from_numpy(a), torch.from_numpy(b)).item(),
                               rtol=1e-6)
    p = O.psnr(a, b, max_val=2.0)
    assert p.shape == (3,)
    for i in range(3):
        mse = np.mean((a[i].astype(np.float64) - b[i]) ** 2)
        np.testing.assert_allclose(p[i], 20 * np.log10(2.0) - 10 * np.log10(mse), rtol=1e-6)
    assert np.isinf(O.psnr(a, a)).all()   # test_metrics.py: identical images -> inf


def test_initialisers_have_the_reference_statistics():
    """he_normal with scale 0.2 (model_builder.py:60-61): truncated at 2 sigma, std sqrt(0.2/fan_in) after correction."""
    rng = np.random.default_rng(0)
    w = O.he_normal_scaled(rng, (3, 3, 64, 256))
    fan_in = 3 * 3 * 64
    sigma = np.sqrt(0.2 / fan_in) / 0.87962566103423978
    assert np.abs(w).max() <= 2 * sigma + 1e-7
    np.testing.assert_allclose(w.std(), np.sqrt(0.2 / fan_in), rtol=2e-2)
    g = O.glorot_uniform(rng, (3, 3, 64, 64))
    assert np.abs(g).max() <= np.sqrt(6.0 / (2 * 9 * 64))


def test_rrdb_structure_counts():
    """RRDB-23 x4 has 351 convs and 16,919,555 parameters; SRResNet x4 37 convs."""
    specs = O.rrdb_layer_specs(upsample_factor=4, num_rrdb_blocks=23)
    assert len(specs) == 351
    assert sum(9 * cin * cout + cout for _, cin, cout in specs) == 16919555
    assert len(O.srresnet_layer_specs(upsample_factor=4)) == 37
    with pytest.raises(ValueError):
        O.rrdb_layer_specs(upsample_factor=3)


def test_rrdb_forward_shapes_and_dtype_modes():
    params = O.init_rrdb_params(seed=1, bias_std=0.05, upsample_factor=2, num_rrdb_blocks=1)
    x = np.random.default_rng(0).uniform(0, 1, size=(1, 6, 7, 3)).astype(np.float32)
    y32 = O.rrdb_forward(params, x, upsample_factor=2, num_rrdb_blocks=1)
    y16 = O.rrdb_forward(params, x, upsample_factor=2, num_rrdb_blocks=1, act_dtype="bf16")
    assert y32.shape == y16.shape == (1, 12, 14, 3)
    assert np.abs(y32).max() <= 1.0
    assert np.abs(y32 - y16).max() < 2e-2
    # no outer per-RRDB residual (model_builder.py:344-351): with zero weights in every dense block the trunk is fea
    zp = {k: (np.zeros_like(v[0]), np.zeros_like(v[1])) if k.startswith("rrdb") else v for k, v in params.items()}
    taps = {}
    O.rrdb_forward(zp, x, upsample_factor=2, num_rrdb_blocks=1, taps=taps)
    np.testing.assert_allclose(taps["trunk_in"], taps["fea"] * 1.2, rtol=1e-6)


def test_srresnet_forward_shapes():
    params = O.init_srresnet_params(seed=1, bias_std=0.05, alpha_std=0.1, upsample_factor=4, num_res_blocks=2)
    x = np.random.default_rng(0).uniform(0, 1, size=(1, 10, 9, 3)).astype(np.float32)
    y = O.srresnet_forward(params, x, upsample_factor=4, num_res_blocks=2)
    assert y.shape == (1, 40, 36, 3) and np.abs(y).max() <= 1.0


def test_batch_norm_inference_matches_torch():
    """oracle.batch_norm_inference == torch.nn.functional.batch_norm(training=False, eps=1e-3) (Keras default epsilon)."""
    import torch
    from oracle import ssr_oracle as O
    rng = np.random.default_rng(11)
    x = rng.standard_normal((2, 5, 7, 8)).astype(np.float32)
    bn = dict(gamma=rng.uniform(0.5, 1.5, 8).astype(np.float32), beta=rng.standard_normal(8).astype(np.float32),
              mean=rng.standard_normal(8).astype(np.float32), var=rng.uniform(0.5, 2.0, 8).astype(np.float32))
    got = O.batch_norm_inference(x, bn)
    t = torch.nn.functional.batch_norm(torch.from_numpy(x).permute(0, 3, 1, 2), torch.from_numpy(bn["mean"]),
                                       torch.from_numpy(bn["var"]), torch.from_numpy(bn["gamma"]),
                                       torch.from_numpy(bn["beta"]), training=False, eps=1e-3)
    np.testing.assert_allclose(got, t.permute(0, 2, 3, 1).numpy(), rtol=1e-5, atol=1e-6)
    # Keras initial state: identity up to 1/sqrt(1 + eps)
    ident = O.init_srresnet_bn(num_res_blocks=1, num_filters=8)["trunk"]
    np.testing.assert_allclose(O.batch_norm_inference(x, ident), x / np.sqrt(1.0 + 1e-3), rtol=1e-6)


def test_torch_cpu_baseline_restatement_matches_the_oracle():
    """bench.py's CPU arm times oracle/torch_cpu.py (torch-CPU / oneDNN restatement of build_enhanced_resnet): it must
    compute what the numpy oracle computes."""
    from oracle import torch_cpu as T
    p = O.init_rrdb_params(seed=1, bias_std=0.05, upsample_factor=4, num_rrdb_blocks=2)
    x = np.random.default_rng(0).uniform(0, 1, size=(2, 20, 24, 3)).astype(np.float32)
    a = O.rrdb_forward(p, x, upsample_factor=4, num_rrdb_blocks=2)
    b = T.rrdb_forward(p, x, upsample_factor=4, num_rrdb_blocks=2)
    assert a.shape == b.shape == (2, 80, 96, 3)
    np.testing.assert_allclose(b, a, rtol=0, atol=1e-5)


def test_total_variation_matches_definition():
    """tf.image.total_variation (vgg_loss.py:167): anisotropic L1 of neighbour differences, per image."""
    x = np.arange(2 * 3 * 4 * 2, dtype=np.float64).reshape(2, 3, 4, 2) ** 1.5
    tv = O.total_variation(x)
    for i in range(2):
        ref = sum(abs(x[i, r + 1, c, k] - x[i, r, c, k]) for r in range(2) for c in range(4) for k in range(2)) + \
            sum(abs(x[i, r, c + 1, k] - x[i, r, c, k]) for r in range(3) for c in range(3) for k in range(2))
        assert tv[i] == pytest.approx(ref)


def test_bicubic_antialias_restatement_matches_an_independent_implementation():
    """tf.image.resize(bicubic, antialias=True) restated from TensorFlow's ScaleAndTranslate (oracle.resize_bicubic)
    against torch's antialiased bicubic interpolation, which implements the same algorithm (Keys a = -0.5, half-pixel
    centres, support scaled by the down-scaling factor, normalised weights)."""
    import torch
    rng = np.random.default_rng(0)
    x = rng.uniform(0, 1, size=(2, 32, 48, 3)).astype(np.float32)
    for scale in (2, 4):
        a = O.resize_bicubic(x, scale, antialias=True)
        t = torch.nn.functional.interpolate(torch.from_numpy(x).permute(0, 3, 1, 2), size=(32 // scale, 48 // scale),
                                            mode="bicubic", antialias=True, align_corners=False)
        np.testing.assert_allclose(a, t.permute(0, 2, 3, 1).numpy(), rtol=0, atol=1e-6)


def test_ssim_and_psnr_y_restatements():
    """tf.image.ssim on constant images has a closed form; identical images give exactly 1; Y of grey is the grey."""
    a = np.full((1, 16, 16, 3), 0.25, np.float32)
    b = np.full((1, 16, 16, 3), 0.5, np.float32)
    c1 = (0.01 * 1.0) ** 2
    assert O.ssim(a, b, max_val=1.0)[0] == pytest.approx((2 * 0.25 * 0.5 + c1) / (0.25 ** 2 + 0.5 ** 2 + c1), rel=1e-6)
    x = np.random.default_rng(1).uniform(0, 1, size=(2, 20, 24, 3)).astype(np.float32)
    np.testing.assert_allclose(O.ssim(x, x, max_val=1.0), 1.0, rtol=1e-7)
    np.testing.assert_allclose(O.rgb_to_y(a), 0.25, rtol=1e-6)
    assert O.psnr_on_y(a, b, max_val=1.0)[0] == pytest.approx(-10 * np.log10(0.0625), rel=1e-5)
    np.testing.assert_array_equal(O.rotate90(x, 1)[0], np.rot90(x[0], 1))
