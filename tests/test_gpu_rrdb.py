"""RRDB generator forward (build_enhanced_resnet) through the Keras-like model object vs the oracle.

Tolerances (BASELINE.json): end-to-end PSNR of the output difference > 50 dB (peak-to-peak 2.0) and
max|err| / max|ref| <= 1e-2 against the fp32 oracle; against the bf16-emulating oracle (same storage
precision, fp32 accumulation) the bound is 10x tighter in practice and catches indexing mistakes.
"""
import numpy as np
import pytest

from tests.helpers import O, rel_err

pytestmark = pytest.mark.gpu


def _model_and_params(nb, sf, seed=1, bias_std=0.05):
    from simplesr_b200 import model_builder as MB
    params = O.init_rrdb_params(seed=seed, bias_std=bias_std, upsample_factor=sf, num_rrdb_blocks=nb)
    m = MB.build_enhanced_resnet(upsample_factor=sf, num_rrdb_blocks=nb, seed=0)
    weights = []
    for name, _, _ in O.rrdb_layer_specs(upsample_factor=sf, num_rrdb_blocks=nb):
        weights.extend(params[name])
    m.set_weights(weights)
    return m, params


@pytest.mark.parametrize("nb,sf,shape", [(2, 4, (2, 24, 20)), (1, 2, (1, 17, 33)), (3, 4, (1, 40, 40))])
def test_rrdb_forward_parity(nb, sf, shape):
    m, params = _model_and_params(nb, sf)
    rng = np.random.default_rng(0)
    x = rng.uniform(0, 1, size=(*shape, 3)).astype(np.float32)
    got = m(x, training=False)
    ref32 = O.rrdb_forward(params, x, upsample_factor=sf, num_rrdb_blocks=nb, act_dtype="f32")
    ref16 = O.rrdb_forward(params, x, upsample_factor=sf, num_rrdb_blocks=nb, act_dtype="bf16")
    assert got.shape == ref32.shape == (shape[0], shape[1] * sf, shape[2] * sf, 3)
    assert np.isfinite(got).all()
    psnr = float(O.psnr(got, ref32, max_val=2.0).min())
    assert psnr > 50.0, psnr
    assert rel_err(got, ref32) <= 1e-2, rel_err(got, ref32)
    assert rel_err(got, ref16) <= 5e-3, rel_err(got, ref16)
    # eager launches and the captured CUDA graph must agree bit for bit
    m.use_graph = False
    again = m(x, training=False)
    assert np.array_equal(got, again)
    m.release()


def test_rrdb_batch_independence():
    """Each image of a batch must equal the same image run alone (tile->CTA mapping must not leak)."""
    m, _ = _model_and_params(1, 4)
    rng = np.random.default_rng(3)
    x = rng.uniform(0, 1, size=(3, 20, 28, 3)).astype(np.float32)
    full = m(x)
    for i in range(3):
        assert np.array_equal(full[i:i + 1], m(x[i:i + 1]))
    m.release()


def test_rrdb_rejects_bad_scale():
    from simplesr_b200 import model_builder as MB
    with pytest.raises(ValueError):
        MB.build_enhanced_resnet(upsample_factor=3)
    with pytest.raises(ValueError):
        MB.build_or_load_generator_model(4, "unknown", 1, 64, 3, 0.2, None, False, (None, None))


def test_plan_variants_agree():
    """The launch-plan options of the RRDB forward (paired growth convs, CTA-pair form of the K = 128 pair, snake tile
    order) change scheduling and the split of the fp32 accumulation, not the arithmetic: snake order and the CTA-pair
    form are bit-identical, pairing the growth convs stays within bf16 output rounding of the plain plan."""
    from simplesr_b200 import model_builder as MB
    x = np.random.default_rng(3).uniform(0, 1, size=(2, 37, 41, 3)).astype(np.float32)
    outs = {}
    for tag, opts in (("default", {}), ("no_snake", dict(snake_order="off")), ("no_cta_pair", dict(pair_growth=False)),
                      ("plain", dict(fuse_growth=False))):
        m = MB.build_enhanced_resnet(upsample_factor=2, num_rrdb_blocks=2, seed=4)
        for k, v in opts.items():
            setattr(m, k, v)
        outs[tag] = m(x, training=False)
        m.release()
    assert np.array_equal(outs["default"], outs["no_snake"])
    assert np.array_equal(outs["default"], outs["no_cta_pair"])
    err = float(np.abs(outs["default"] - outs["plain"]).max() / np.abs(outs["plain"]).max())
    assert err <= 1e-2, err


@pytest.mark.parametrize("nb,sf,shape,fuse", [(2, 2, (2, 37, 41), True), (2, 4, (4, 64, 64), True), (1, 4, (16, 32, 32), False),
                                              (3, 2, (1, 130, 70), True), (1, 4, (3, 8, 8), True)])
def test_chained_launches_are_bit_identical(nb, sf, shape, fuse):
    """Tile-level dependencies between consecutive conv launches (ssr_conv_chain_*) only change WHEN a pixel tile may
    start: the result equals the plan with whole-grid dependencies bit for bit - small grids (several launches resident
    at once), grids of one CTA per SM, ragged images with the strip geometry, the paired and the plain growth convs -
    also over repeated graph launches, and no tile wait ever gives up."""
    from simplesr_b200 import model_builder as MB
    from tests.helpers import L
    x = np.random.default_rng(5).uniform(0, 1, size=(*shape, 3)).astype(np.float32)
    outs = {}
    for chained in (True, False):
        m = MB.build_enhanced_resnet(upsample_factor=sf, num_rrdb_blocks=nb, seed=4)
        m.chain_deps, m.fuse_growth = chained, fuse
        outs[chained] = m(x, training=False)
        plan = m.plan(*shape)
        if chained:
            assert plan.chain_stats is not None and plan.chain_stats[1] >= nb * 3 * 2, plan.chain_stats
            print("chain stats", shape, plan.chain_stats)
            s = m.stream.ptr
            for _ in range(10):
                plan.run(s)
            again = np.empty_like(outs[True])
            L.check(m.ctx.lib.ssr_memcpy_d2h(again.ctypes.data, plan.buffers["out_f32"].ptr, again.nbytes, s))
            m.stream.sync()
            assert np.array_equal(again, outs[True])
            assert plan.chain_timeouts() == 0
        else:
            assert plan.chain_stats is None
        m.release()
    assert np.array_equal(outs[True], outs[False])
