"""Builds tests/golden/tiling_fixtures.npz from the reference's own test fixtures (run HERE, where /root/reference exists).

The reference pins its tiling path with known-answer round trips on three images
(/root/reference/tests/utils/image/test_image_utils.py:69-111; fixtures tests/data/{baboon,comic,lena}.png) and ships
the outputs of its own run under tests/data/reconstructed/recon{idx}_{p}x{p}.png.  Those dumps went through
array_to_img(scale=True) (image_utils.py:28-37), i.e. a min-max stretch: for comic.png (range 0..255) the dump is the
stitched tensor itself; for baboon / lena it is round(255 * (x - min) / (max - min)).

Stored (uint8, to keep the fixture small):
  comic            [361,250,3]  full image (non-multiple of every patch size)
  baboon_crop      [203,187,3]  rows 100.., cols 200.. of baboon.png (ragged size)
  lena_crop        [256,192,3]  rows 128.., cols 160.. of lena.png (multiple of 32/64, ragged for 128 in width)
  ref_recon_comic_sha256_{32,64,128}   sha256 of the raw bytes of the reference's own stitched output for comic.png
                                       (its tests/data/reconstructed/recon2_*.png, decoded) - the known answer
"""
import hashlib
import os
import sys

import numpy as np
from PIL import Image

REF = "/root/reference/tests/data"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                   "tiling_fixtures.npz")


def load(name):
    return np.array(Image.open(os.path.join(REF, name)))


def main():
    if not os.path.isdir(REF):
        sys.exit("reference fixtures not found (this script only runs in the build container)")
    baboon, comic, lena = load("baboon.png"), load("comic.png"), load("lena.png")
    out = {
        "comic": comic,
        "baboon_crop": baboon[100:303, 200:387],
        "lena_crop": lena[128:384, 160:352],
    }
    for ps in (32, 64, 128):
        ref = np.ascontiguousarray(load(f"reconstructed/recon2_{ps}x{ps}.png"))
        out[f"ref_recon_comic_sha256_{ps}"] = np.frombuffer(hashlib.sha256(ref.tobytes()).digest(), np.uint8)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
