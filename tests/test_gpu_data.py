"""Device-side data preparation and evaluation metrics against the oracle: bicubic-antialias LR
synthesis (tf.image.resize), flips / rot90 / crops (exact copies), PSNR-Y and SSIM (tf.image)."""
import numpy as np
import pytest

from tests.helpers import O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("scale,shape", [(4, (2, 96, 128)), (2, (1, 50, 34)), (8, (1, 64, 64))])
def test_bicubic_antialias_downscale(ctx, scale, shape):
    from simplesr_b200 import image_transforms as IT
    rng = np.random.default_rng(0)
    hr = rng.integers(0, 256, size=(*shape, 3)).astype(np.float32)
    lr, hr11 = IT.prepare_img_pairs(hr, scale)
    ref_lr, ref_hr = O.prepare_img_pairs(hr, scale)
    assert lr.shape == (shape[0], shape[1] // scale, shape[2] // scale, 3)
    np.testing.assert_allclose(lr, ref_lr, rtol=0, atol=2e-6)          # same weights; fp32 summation order differs
    np.testing.assert_array_equal(hr11, ref_hr)
    np.testing.assert_allclose(IT.resize_bicubic(hr[0], scale, antialias=False),
                               O.resize_bicubic(hr[:1], scale, antialias=False)[0], rtol=0, atol=3e-4)


def test_augmentations_are_exact_copies(ctx):
    from simplesr_b200 import image_transforms as IT
    rng = np.random.default_rng(1)
    x = rng.standard_normal((3, 10, 14, 3)).astype(np.float32)
    np.testing.assert_array_equal(IT.flip_along_x(x), O.flip_along_x(x))
    np.testing.assert_array_equal(IT.flip_along_y(x), O.flip_along_y(x))
    np.testing.assert_array_equal(IT.flip_along_x(x[0]), O.flip_along_x(x[0]))
    for k in range(5):
        np.testing.assert_array_equal(IT.rotate90(x, k), O.rotate90(x, k))
    crops = IT.crop_naive(x[1], 5, (6, 7, 3), random_seed=3)
    assert crops.shape == (5, 6, 7, 3)
    r = np.random.default_rng(3)
    oy, ox = r.integers(0, 10 - 6 + 1, size=5), r.integers(0, 14 - 7 + 1, size=5)
    for i in range(5):
        np.testing.assert_array_equal(crops[i], x[1, oy[i]:oy[i] + 6, ox[i]:ox[i] + 7])
    with pytest.raises(ValueError):
        IT.crop_naive(x[0], 1, (11, 4, 3))


def test_psnr_y_and_ssim(ctx):
    from simplesr_b200 import metrics as M
    rng = np.random.default_rng(2)
    a = rng.uniform(-1, 1, size=(3, 40, 52, 3)).astype(np.float32)
    b = np.clip(a + rng.normal(0, 0.1, size=a.shape), -1, 1).astype(np.float32)
    np.testing.assert_allclose(M.psnr(a, b), O.psnr(a, b, max_val=2.0), rtol=1e-5)
    np.testing.assert_allclose(M.psnr_on_y(a, b), O.psnr_on_y(a, b, max_val=2.0), rtol=1e-5)
    np.testing.assert_allclose(M.ssim(a, b), O.ssim(a, b, max_val=2.0), rtol=2e-5)
    np.testing.assert_allclose(M.ssim(a, b, max_val=255), O.ssim(a, b, max_val=255), rtol=2e-5)
    assert M.ssim(a[0], a[0]) == pytest.approx(1.0, abs=1e-6)          # tests/utils/image/test_metrics.py:47-49
    assert np.isinf(M.psnr_on_y(a[0], a[0]))                            # :35-36
    # the reference's own check of psnr_on_y (test_metrics.py:39-45)
    y1, y2 = O.rgb_to_y(a[0]), O.rgb_to_y(b[0])
    ref = -10 * np.log10(np.mean((y1 - y2) ** 2))
    assert M.psnr_on_y(a[0], b[0], max_val=1.0) == pytest.approx(ref, abs=1e-4)
    with pytest.raises(ValueError):
        M.psnr_on_y(a, b[:2])
