"""Pins the oracle's tiling / stitching against the reference's own known-answer fixtures (CPU only).

Reference tests restated here: tests/utils/image/test_image_utils.py:44-111 (segmentation shapes, segment +
reconstruct round trips on the 3x3 / 5x3 matrices and on the PNG fixtures, with and without overlap) and the
outputs of the reference's own run (tests/data/reconstructed/recon2_*.png) through their sha256 in
tests/golden/tiling_fixtures.npz (made by tools/make_golden.py).
"""
import hashlib
import os

import numpy as np
import pytest

from oracle import ssr_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tiling_fixtures.npz")
PATCH_DIMS = [(1, 1), (2, 2), (3, 3), (3, 1), (1, 3), (2, 3), (3, 2)]   # test_image_utils.py:10

MAT_3X3 = np.array([[[1] * 3, [2] * 3, [3] * 3], [[4] * 3, [5] * 3, [6] * 3], [[7] * 3, [8] * 3, [9] * 3]])
MAT_5X3 = np.arange(1, 16).reshape(3, 5, 1).repeat(3, axis=2)


@pytest.fixture(scope="module")
def golden():
    with np.load(GOLDEN) as z:
        return {k: z[k] for k in z.files}


@pytest.mark.parametrize("patch_dim", PATCH_DIMS)
@pytest.mark.parametrize("matrix", [MAT_3X3, MAT_5X3], ids=["3x3", "5x3"])
def test_segmentation_shapes_and_round_trip(matrix, patch_dim):
    """test_image_utils.py:44-67."""
    patches, padding = O.segment_into_patches(matrix, patch_width=patch_dim[0], patch_height=patch_dim[1])
    assert patches.ndim == 4
    assert patches.shape[2] == patch_dim[0] and patches.shape[1] == patch_dim[1]
    rec = O.reconstruct_from_patches(patches, original_height=matrix.shape[0], original_width=matrix.shape[1],
                                     horizontal_padding=padding[0][1], vertical_padding=padding[1][1])
    np.testing.assert_array_equal(matrix, rec)


@pytest.mark.parametrize("name", ["comic", "baboon_crop", "lena_crop"])
@pytest.mark.parametrize("ps", [32, 64, 128])
def test_overlap_round_trip_on_reference_images(golden, name, ps):
    """test_image_utils.py:69-90: patch in {32,64,128}, overlap = patch // 4, images that are not multiples."""
    img = golden[name]
    ov = ps // 4
    patches, padding = O.segment_into_patches(img[None], patch_width=ps, patch_height=ps, pixel_overlap=ov)
    assert patches.shape[1:] == (ps + 2 * ov, ps + 2 * ov, 3)
    rec = O.reconstruct_from_overlapping_patches(patches, image_height=img.shape[0], image_width=img.shape[1],
                                                 pixel_overlap=ov, horizontal_padding=padding[0][1] - ov,
                                                 vertical_padding=padding[1][1] - ov)
    assert rec.dtype == img.dtype
    np.testing.assert_array_equal(img, rec)
    if name == "comic":  # the reference's own output for this input, bit for bit
        digest = hashlib.sha256(np.ascontiguousarray(rec).tobytes()).digest()
        assert digest == golden[f"ref_recon_comic_sha256_{ps}"].tobytes()


@pytest.mark.parametrize("name", ["comic", "baboon_crop", "lena_crop"])
@pytest.mark.parametrize("ps", [32, 64, 128])
def test_plain_round_trip_on_reference_images(golden, name, ps):
    """test_image_utils.py:92-111 (no overlap: space_to_batch / batch_to_space formulation)."""
    img = golden[name]
    patches, padding = O.segment_into_patches(img, patch_width=ps, patch_height=ps)
    rec = O.reconstruct_from_patches(patches, original_height=img.shape[0], original_width=img.shape[1],
                                     horizontal_padding=padding[0][1], vertical_padding=padding[1][1])
    np.testing.assert_array_equal(img, rec)


def test_tile_order_and_zero_padding():
    """image_utils.py:124-148: row-major tiles, zero padding of `overlap` on every side plus the ragged remainder."""
    img = np.arange(1, 5 * 7 * 1 + 1, dtype=np.float32).reshape(5, 7, 1)
    patches, padding = O.segment_into_patches(img, patch_width=4, patch_height=4, pixel_overlap=1)
    assert padding == [[1, 1 + 3], [1, 1 + 1]]
    assert patches.shape == (4, 6, 6, 1)
    padded = np.pad(img, [(1, 4), (1, 2), (0, 0)])
    np.testing.assert_array_equal(patches[0], padded[0:6, 0:6])
    np.testing.assert_array_equal(patches[1], padded[0:6, 4:10])
    np.testing.assert_array_equal(patches[2], padded[4:10, 0:6])
    np.testing.assert_array_equal(patches[3], padded[4:10, 4:10])


def test_errors_match_reference():
    """image_utils.py:108-116 and :54-55,:76-80."""
    with pytest.raises(ValueError, match="larger than image"):
        O.segment_into_patches(np.zeros((16, 16, 3)), patch_width=32, patch_height=32)
    with pytest.raises(ValueError, match="rank 3"):
        O.segment_into_patches(np.zeros((2, 16, 16, 3)), patch_width=8, patch_height=8)
    with pytest.raises(ValueError, match="rank 4"):
        O.reconstruct_from_overlapping_patches(np.zeros((16, 16, 3)), 16, 16, 2, 0, 0)
    with pytest.raises(ValueError, match="negative"):
        O.reconstruct_from_patches(np.zeros((1, 4, 4, 3)), 4, 4, horizontal_padding=-1)


def test_tiled_upscale_with_identity_model_is_exact():
    """evaluation.py:253-277 + :351-359: with a nearest-neighbour 'model' the stitched result equals the direct
    upscale bit for bit (tiling is pure data movement)."""
    rng = np.random.default_rng(0)
    lr = rng.uniform(0, 1, size=(1, 150, 170, 3)).astype(np.float32)
    up = lambda t: t.repeat(4, axis=1).repeat(4, axis=2)
    got = O.tiled_upscale(up, lr, scale=4, patch=64, pixel_overlap=16)
    np.testing.assert_array_equal(got, up(lr)[0])


def test_eligible_efficient_inference():
    """evaluation.py:340-348."""
    assert O.eligible_efficient_inference((1, 2048, 2048, 3))
    assert O.eligible_efficient_inference((1001, 1001, 3))
    assert not O.eligible_efficient_inference((2, 2048, 2048, 3))
    assert not O.eligible_efficient_inference((1, 1000, 2048, 3))
    assert not O.eligible_efficient_inference((2048, 2048))
