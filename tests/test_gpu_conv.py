"""Parity of ssr_conv2d_fwd (tcgen05 implicit GEMM) against the oracle, through the C ABI.

Tolerance: inputs/weights are bf16-exact on both sides, accumulation is fp32; the only differences are the
fp32 summation order and the final bf16 rounding of the output (2^-9 relative) -> per-layer
max|err| / max|ref| <= 1e-2 (BASELINE.json), in practice ~3e-3.
"""
import numpy as np
import pytest

from tests.helpers import L, conv_case, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-2

CASES = {
    "rrdb_growth0": dict(n=1, h=16, w=16, cin_real=64, cout=32, act=L.ACT_LRELU),
    "rrdb_growth1_ragged": dict(n=2, h=20, w=37, cin_real=96, cout=32, act=L.ACT_LRELU),
    "rrdb_growth2": dict(n=1, h=9, w=50, cin_real=128, cout=32, act=L.ACT_LRELU),
    "rrdb_growth3_slice": dict(n=1, h=31, w=17, cin_real=160, cout=32, act=L.ACT_LRELU, in_cstride=192,
                               out_cstride=192, out_coff=160),
    "rrdb_out_residual": dict(n=1, h=33, w=18, cin_real=192, cout=64, res=True, res_beta=0.2),
    "trunk_residual": dict(n=2, h=16, w=16, cin_real=64, cout=64, res=True, res_beta=1.0),
    "upconv_d2s": dict(n=1, h=12, w=13, cin_real=64, cout=256, up=2, act=L.ACT_LRELU),
    "last_tanh_f32": dict(n=1, h=24, w=24, cin_real=64, cout=3, act=L.ACT_TANH, out_dtype=L.SSR_F32),
    "first_rgb": dict(n=2, h=16, w=19, cin_real=3, cout=64),
    "tiny_image": dict(n=3, h=3, w=2, cin_real=64, cout=32),
    "single_pixel": dict(n=1, h=1, w=1, cin_real=64, cout=64),
    "wide_row": dict(n=1, h=2, w=300, cin_real=64, cout=32),
    "prelu": dict(n=1, h=16, w=16, cin_real=64, cout=64, act=L.ACT_PRELU),
    "relu_vgg_like": dict(n=1, h=16, w=16, cin_real=64, cout=128, act=L.ACT_RELU),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_conv_parity(ctx, name):
    got, ref, untouched = conv_case(ctx, **CASES[name])
    assert got.shape == ref.shape
    assert np.isfinite(got).all()
    assert rel_err(got, ref) <= TOL, f"{name}: rel_err {rel_err(got, ref)}"
    # channels before out_coff belong to other layers of the dense block and must not be written
    assert not untouched.any()


def test_conv_rejects_bad_arguments(ctx):
    with pytest.raises(ValueError):
        ctx.conv_packed_bytes(5, 64, 64, 1)       # kernel size not in {1,3,9}
    with pytest.raises(ValueError):
        ctx.conv_packed_bytes(3, 20, 64, 1)       # cin not a multiple of 16
    with pytest.raises(ValueError):
        ctx.conv_packed_bytes(3, 64, 100, 2)      # depth_to_space needs cout % 64 == 0


def test_conv_is_deterministic(ctx):
    a, _, _ = conv_case(ctx, n=2, h=20, w=37, cin_real=160, cout=32, act=L.ACT_LRELU, seed=5)
    b, _, _ = conv_case(ctx, n=2, h=20, w=37, cin_real=160, cout=32, act=L.ACT_LRELU, seed=5)
    assert np.array_equal(a, b)


@pytest.mark.parametrize("kw", [
    dict(n=2, h=37, w=29, cin_real=64, cout=32, act=L.ACT_LRELU),                 # direct epilogue
    dict(n=1, h=40, w=48, cin_real=192, cout=64, res=True),                       # CTA pair, staged residual
    dict(n=1, h=21, w=16, cin_real=192, cout=64, res=True),                       # CTA pair with an odd tile count (dummy tile)
    dict(n=2, h=19, w=23, cin_real=64, cout=256, up=2, act=L.ACT_LRELU),          # depth_to_space store map
    dict(n=3, h=16, w=16, cin_real=32, cout=96, res=True, res_beta=1.0),          # 64-byte rows, multi-pass residual
])
def test_tile_order_does_not_change_the_result(ctx, kw):
    """desc.tile_order only changes the order in which a CTA walks its pixel tiles: bit-identical outputs."""
    a, ref, _ = conv_case(ctx, tile_order=0, **kw)
    b, _, _ = conv_case(ctx, tile_order=1, **kw)
    assert np.array_equal(a, b)
    assert rel_err(a, ref) <= 1e-2


@pytest.mark.parametrize("kw,tiles", [
    (dict(n=2, h=128, w=128, cin_real=64, cout=32, act=L.ACT_LRELU), (304, 296)),           # direct epilogue
    (dict(n=2, h=128, w=128, cin_real=192, cout=64, res=True), (304, 296)),                 # CTA pair, staged residual
    (dict(n=2, h=128, w=128, cin_real=64, cout=256, up=2, act=L.ACT_LRELU), (304, 296)),    # depth_to_space store map
    (dict(n=2, h=128, w=128, cin_real=32, cout=32, act=L.ACT_LRELU), None),                 # 64-byte operand rows
    (dict(n=4, h=100, w=75, cin_real=64, cout=64, res=True, res_beta=1.0), None),           # ragged in x as well
])
@pytest.mark.parametrize("order", [0, 1])
def test_strip_tile_geometry_does_not_change_the_result(ctx, kw, tiles, order):
    """A ragged bottom row of pixel tiles is covered by fewer, wider tiles when that saves a round over the SMs (148
    tiles per 128x128 image instead of 152).  Every pixel still sees the same MMA sequence: bit-identical outputs."""
    try:
        ctx.debug_set(8)                     # geometry A only
        a, ref, _ = conv_case(ctx, tile_order=order, **kw)
        ta = ctx.last_conv_tiles
    finally:
        ctx.debug_set(0)
    b, _, _ = conv_case(ctx, tile_order=order, **kw)
    tb = ctx.last_conv_tiles
    if tiles is not None and ctx.sm_count == 148:
        assert (ta, tb) == tiles
    assert tb <= ta
    assert np.array_equal(a, b)
    assert rel_err(a, ref) <= 1e-2


@pytest.mark.parametrize("kw", [
    dict(n=2, h=16, w=16, cin_real=256, cout=256, act=L.ACT_RELU),       # VGG block 3: 8 slabs of 32 rows -> 4 slab pairs
    dict(n=3, h=8, w=8, cin_real=512, cout=512),                         # VGG block 5 (pre-activation): 32 slabs of 16 rows
    dict(n=1, h=17, w=9, cin_real=256, cout=512, act=L.ACT_LRELU),       # discriminator 256 -> 512, odd tile count
    dict(n=2, h=24, w=24, cin_real=128, cout=128, act=L.ACT_RELU),       # VGG block 2: two slabs of 64 rows, N = 128 pair
    dict(n=1, h=12, w=40, cin_real=512, cout=256, res=True, res_beta=1.0),  # dgrad shape of 256 -> 512, with residual
])
def test_deep_layers_run_on_cta_pairs(ctx, kw):
    """Layers whose weight slab only fits 16 or 32 output rows per SM (cin >= 256) pair their slabs up: two CTAs hold one
    slab each and the leader issues M = 256 MMAs over both, N twice as wide.  Same MMA sequence per output element:
    bit-identical to one slab per CTA, and within tolerance of the oracle."""
    try:
        ctx.debug_set(0x10000)               # round-1 pairings only: these layers run one slab per CTA
        a, ref, _ = conv_case(ctx, **kw)
    finally:
        ctx.debug_set(0)
    b, _, _ = conv_case(ctx, **kw)
    assert np.array_equal(a, b)
    assert rel_err(b, ref) <= TOL, rel_err(b, ref)
