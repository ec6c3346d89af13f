"""GPU parity of the training-step kernels against the oracle, through the C ABI: pixel losses + PSNR + gradient,
Adam, per-channel reductions, activation backward, space_to_depth, dgrad (conv kernel with rotated weights) and the
split-K tcgen05 wgrad.  Tolerances: fp32 elementwise 1e-6 relative; bf16-input GEMMs max|err|/max|ref| <= 1e-2."""
import numpy as np
import pytest

from tests.helpers import L, O, rel_err

pytestmark = pytest.mark.gpu


def _dev(a):
    return L.DeviceBuffer.from_numpy(np.ascontiguousarray(a))


def test_pixel_loss_and_gradient(ctx):
    rng = np.random.default_rng(0)
    n, per = 5, 40 * 36 * 3
    hr = rng.uniform(-1, 1, size=(n, per)).astype(np.float32)
    sr = rng.uniform(-1, 1, size=(n, per)).astype(np.float32)
    sr[3, 7] = hr[3, 7]   # a zero difference: sign(0) = 0 in the MAE gradient
    dh, ds = _dev(hr), _dev(sr)
    dg = L.DeviceBuffer(hr.nbytes)
    ws = L.DeviceBuffer(L.load().ssr_pixel_loss_workspace_bytes(n))
    out = L.DeviceBuffer((2 + n) * 4)
    w_mse, w_mae = 0.7, 0.01
    L.pixel_loss(dh, ds, n, per, w_mse, w_mae, 2.0, dg, ws, out)
    got = out.download((2 + n,), np.float32)
    np.testing.assert_allclose(got[0], O.mean_squared_error(hr, sr), rtol=2e-6)
    np.testing.assert_allclose(got[1], O.mean_absolute_error(hr, sr), rtol=2e-6)
    np.testing.assert_allclose(got[2:], O.psnr(hr.reshape(n, 40, 36, 3), sr.reshape(n, 40, 36, 3), 2.0), rtol=2e-6)
    d = sr - hr
    ref_g = (w_mse * 2.0 * d + w_mae * np.sign(d)) / d.size
    np.testing.assert_allclose(dg.download(hr.shape, np.float32), ref_g, rtol=1e-5, atol=1e-12)
    # deterministic
    L.pixel_loss(dh, ds, n, per, w_mse, w_mae, 2.0, dg, ws, out)
    assert np.array_equal(out.download((2 + n,), np.float32), got)


def test_adam_step_matches_keras_semantics(ctx):
    rng = np.random.default_rng(1)
    cnt = 100003
    p = rng.standard_normal(cnt).astype(np.float32)
    m = np.zeros(cnt, np.float32)
    v = np.zeros(cnt, np.float32)
    dp, dm, dv = _dev(p), _dev(m), _dev(v)
    for t in (1, 2, 3):
        g = rng.standard_normal(cnt).astype(np.float32) * 0.1
        dg = _dev(g)
        lr_t = 1e-3 * np.sqrt(1 - 0.999 ** t) / (1 - 0.9 ** t)
        L.adam_step(dp, dg, dm, dv, cnt, float(lr_t), 0.9, 0.999, 1e-7)
        p, m, v = O.adam_update(p, g, m, v, t, lr=1e-3)
        dg.free()
    np.testing.assert_allclose(dp.download((cnt,), np.float32), p, rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(dv.download((cnt,), np.float32), v, rtol=1e-5)


def test_channel_sums_act_bwd_s2d(ctx):
    rng = np.random.default_rng(2)
    px, c = 777, 64
    dy = O.bf16_round(rng.standard_normal((px, c)).astype(np.float32))
    z = O.bf16_round(rng.standard_normal((px, c)).astype(np.float32))
    alpha = rng.uniform(0, 0.5, c).astype(np.float32)
    ddy, dz_, da = _dev(L.f32_to_bf16_bits(dy)), _dev(L.f32_to_bf16_bits(z)), _dev(alpha)
    ws = L.DeviceBuffer(L.load().ssr_channel_sum_workspace_bytes(c))
    out = L.DeviceBuffer(c * 4)
    L.channel_sum_bf16(ddy, c, 0, None, 0, 0, px, c, 1.0, False, ws, out)
    np.testing.assert_allclose(out.download((c,), np.float32), dy.sum(0), rtol=1e-4, atol=1e-4)
    L.channel_sum_bf16(ddy, c, 0, dz_, c, 0, px, c, 1.0, False, ws, out)
    np.testing.assert_allclose(out.download((c,), np.float32), (dy * np.minimum(z, 0)).sum(0), rtol=1e-4, atol=1e-4)
    dout = L.DeviceBuffer(px * c * 2)
    L.act_bwd_bf16(ddy, c, 0, dz_, c, 0, da, 0.0, dout, c, 0, px, c)
    got = L.bf16_bits_to_f32(dout.download((px, c), np.uint16))
    np.testing.assert_array_equal(got, O.bf16_round(dy * np.where(z > 0, 1.0, alpha).astype(np.float32)))
    x = rng.integers(0, 65536, size=(2, 6, 10, 16)).astype(np.uint16)
    dx, dy2 = _dev(x), L.DeviceBuffer(x.nbytes)
    L.space_to_depth2(dx, dy2, 2, 3, 5, 16, 2)
    np.testing.assert_array_equal(dy2.download((2, 3, 5, 64), np.uint16), O.space_to_depth(x, 2))


def _conv_with_packed(ctx, x, packed, cin, cout, ks, kw=None, res=None):
    n, h, w, cs = x.shape
    dx = _dev(L.f32_to_bf16_bits(x))
    dout = L.DeviceBuffer(n * h * w * cout * 2)
    dres = _dev(L.f32_to_bf16_bits(res)) if res is not None else None
    d = L.ConvDesc(n=n, h=h, w=w, cin=cin, in_cstride=cs, cout=cout, ksize=ks, ksize_w=(kw or 0), act=L.ACT_NONE,
                   act_alpha=0.0, res_beta=1.0, up=1, out_dtype=L.SSR_BF16, out_cstride=cout, out_coff=0,
                   res_dtype=(L.SSR_BF16 if res is not None else L.SSR_NONE), res_cstride=cout, res_coff=0,
                   out2_cstride=0, out2_coff=0)
    ctx.conv2d_fwd(d, dx, packed, None, dout, res=dres)
    L.stream_sync()
    return L.bf16_bits_to_f32(dout.download((n, h, w, cout), np.uint16))


@pytest.mark.parametrize("cin,cout,h,w", [(64, 64, 13, 17), (64, 256, 8, 9), (192, 32, 10, 12)])
def test_dgrad_3x3(ctx, cin, cout, h, w):
    """dX of a 3x3 conv = ssr_conv2d_fwd over dZ with the rotated / transposed packed weights."""
    rng = np.random.default_rng(3)
    k = O.bf16_round(rng.standard_normal((3, 3, cin, cout)).astype(np.float32) / np.sqrt(9 * cout))
    dz = O.bf16_round(rng.standard_normal((2, h, w, cout)).astype(np.float32))
    ref, _, _ = O.conv2d_same_backward(np.zeros((2, h, w, cin), np.float32), k, dz)
    dk = _dev(k)
    cin_d = -(-cout // 16) * 16
    packed = L.DeviceBuffer(ctx.conv_packed_bytes(3, cin_d, cin, 1))
    ctx.conv_pack_weights_dgrad(dk, 3, 3, cin, cout, packed)
    got = _conv_with_packed(ctx, dz, packed, cin_d, cin, 3)
    assert rel_err(got, ref) <= 1e-2


def test_dgrad_9x9_unrolled(ctx):
    """The 9x9x64->3 output conv: dgrad over the x-unrolled fp32 dZ (27 -> 32 channels, 9x1 taps)."""
    rng = np.random.default_rng(4)
    n, h, w = 1, 20, 22
    k = O.bf16_round(rng.standard_normal((9, 9, 64, 3)).astype(np.float32) / 15.0)
    dz = O.bf16_round(rng.standard_normal((n, h, w, 3)).astype(np.float32))
    ref, _, _ = O.conv2d_same_backward(np.zeros((n, h, w, 64), np.float32), k, dz)
    ddz = _dev(dz)
    dzu = L.DeviceBuffer(n * h * w * 32 * 2)
    L.im2col_x_f32_to_bf16(ddz, dzu, n, h, w, 3, 9, 32)
    packed = L.DeviceBuffer(ctx.conv_packed_bytes(9, 32, 64, 1, ksize_w=1))
    ctx.conv_pack_weights_dgrad(_dev(k), 9, 9, 64, 3, packed, unroll_x=True)
    dout = L.DeviceBuffer(n * h * w * 64 * 2)
    d = L.ConvDesc(n=n, h=h, w=w, cin=32, in_cstride=32, cout=64, ksize=9, ksize_w=1, act=L.ACT_NONE, act_alpha=0.0,
                   res_beta=0.0, up=1, out_dtype=L.SSR_BF16, out_cstride=64, out_coff=0, res_dtype=L.SSR_NONE,
                   res_cstride=0, res_coff=0, out2_cstride=0, out2_coff=0)
    ctx.conv2d_fwd(d, dzu, packed, None, dout)
    L.stream_sync()
    got = L.bf16_bits_to_f32(dout.download((n, h, w, 64), np.uint16))
    assert rel_err(got, ref) <= 1e-2


WGRAD_CASES = {
    "res_64_64": dict(n=2, h=24, w=24, cin=64, cout=64, kh=3, kw=3),
    "ragged_rows": dict(n=1, h=13, w=20, cin=64, cout=64, kh=3, kw=3),
    "up_64_256": dict(n=1, h=12, w=16, cin=64, cout=256, kh=3, kw=3),
    "dense_192_32": dict(n=1, h=16, w=16, cin=192, cout=32, kh=3, kw=3),
    "first_unrolled_9x1": dict(n=2, h=24, w=24, cin=27, cout=64, kh=9, kw=1, xcs=32),
    "last_9x9_to_3": dict(n=1, h=32, w=32, cin=64, cout=3, kh=9, kw=9, zcs=16),
    "hr_96": dict(n=1, h=96, w=96, cin=64, cout=64, kh=3, kw=3),
}


@pytest.mark.parametrize("name", sorted(WGRAD_CASES))
def test_wgrad_parity(ctx, name):
    c = dict(WGRAD_CASES[name])
    n, h, w, cin, cout, kh, kw = (c[k] for k in ("n", "h", "w", "cin", "cout", "kh", "kw"))
    xcs = c.get("xcs", -(-cin // 8) * 8)
    zcs = c.get("zcs", -(-cout // 8) * 8)
    rng = np.random.default_rng(5)
    x = np.zeros((n, h, w, xcs), np.float32)
    x[..., :cin] = O.bf16_round(rng.standard_normal((n, h, w, cin)).astype(np.float32))
    dz = np.zeros((n, h, w, zcs), np.float32)
    dz[..., :cout] = O.bf16_round(rng.standard_normal((n, h, w, cout)).astype(np.float32))
    _, ref, _ = O.conv2d_same_backward(x[..., :cin], np.zeros((kh, kw, cin, cout), np.float32), dz[..., :cout])
    dx, ddz = _dev(L.f32_to_bf16_bits(x)), _dev(L.f32_to_bf16_bits(dz))
    ws = L.DeviceBuffer(ctx.conv_wgrad_workspace_bytes(h, w, cin, cout, kh, kw))
    dw = L.DeviceBuffer(ref.nbytes)
    ctx.conv2d_wgrad(dx, xcs, 0, cin, ddz, zcs, 0, cout, n, h, w, kh, kw, ws, dw)
    L.stream_sync()
    got = dw.download(ref.shape, np.float32)
    assert np.isfinite(got).all()
    assert rel_err(got, ref) <= 1e-2, (name, rel_err(got, ref))
    # accumulate + scale, and determinism
    ctx.conv2d_wgrad(dx, xcs, 0, cin, ddz, zcs, 0, cout, n, h, w, kh, kw, ws, dw, scale=0.5, accumulate=True)
    L.stream_sync()
    got2 = dw.download(ref.shape, np.float32)
    np.testing.assert_allclose(got2, got * 1.5, rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("name", ["res_64_64", "ragged_rows", "up_64_256", "dense_192_32", "hr_96"])
def test_wgrad_with_bias_gradient(ctx, name):
    """ssr_conv2d_wgrad_bias: dW identical to ssr_conv2d_wgrad, dbias = BiasAddGrad of dZ (fp32 sum of the bf16 values:
    exact up to summation order, tolerance 1e-5 relative to max|ref|)."""
    c = dict(WGRAD_CASES[name])
    n, h, w, cin, cout, kh, kw = (c[k] for k in ("n", "h", "w", "cin", "cout", "kh", "kw"))
    xcs, zcs = -(-cin // 8) * 8, -(-cout // 8) * 8
    rng = np.random.default_rng(6)
    x = O.bf16_round(rng.standard_normal((n, h, w, xcs)).astype(np.float32))
    dz = O.bf16_round(rng.standard_normal((n, h, w, zcs)).astype(np.float32))
    dx, ddz = _dev(L.f32_to_bf16_bits(x)), _dev(L.f32_to_bf16_bits(dz))
    ws = L.DeviceBuffer(ctx.conv_wgrad_workspace_bytes(h, w, cin, cout, kh, kw))
    dw0, dw1 = L.DeviceBuffer(kh * kw * cin * cout * 4), L.DeviceBuffer(kh * kw * cin * cout * 4)
    db = L.DeviceBuffer.from_numpy(np.full(cout, 7.0, np.float32))
    ctx.conv2d_wgrad(dx, xcs, 0, cin, ddz, zcs, 0, cout, n, h, w, kh, kw, ws, dw0)
    ctx.conv2d_wgrad(dx, xcs, 0, cin, ddz, zcs, 0, cout, n, h, w, kh, kw, ws, dw1, dbias=db, bias_scale=0.5)
    L.stream_sync()
    a, b = dw0.download((kh, kw, cin, cout), np.float32), dw1.download((kh, kw, cin, cout), np.float32)
    np.testing.assert_array_equal(a, b)
    ref = 0.5 * dz[..., :cout].astype(np.float64).sum(axis=(0, 1, 2))
    got = db.download((cout,), np.float32)
    assert np.abs(got - ref).max() <= 1e-5 * max(1.0, np.abs(ref).max()), np.abs(got - ref).max()
    ctx.conv2d_wgrad(dx, xcs, 0, cin, ddz, zcs, 0, cout, n, h, w, kh, kw, ws, dw1, dbias=db, bias_scale=0.5,
                     bias_accumulate=True)
    L.stream_sync()
    np.testing.assert_allclose(db.download((cout,), np.float32), 2 * got, rtol=1e-6, atol=1e-6)


def test_pack_batch_matches_single_packs(ctx):
    """ssr_conv2d_pack_batch writes byte-identical images to the per-layer pack calls (forward, dgrad, x-unrolled dgrad)."""
    rng = np.random.default_rng(8)
    cases = [(3, 3, 64, 64, 32, 1, 0), (3, 3, 160, 160, 32, 1, 1), (3, 3, 64, 64, 256, 2, 0), (3, 3, 192, 192, 64, 1, 1),
             (9, 9, 64, 64, 3, 1, 2), (3, 3, 3, 16, 64, 1, 0)]
    items, singles, bufs = [], [], []
    for kh, kw, cin_real, cin, cout, up, mode in cases:
        k = _dev(rng.standard_normal((kh, kw, cin_real, cout)).astype(np.float32))
        if mode == 0:
            nbytes = ctx.conv_packed_bytes(kh, cin, cout, up, ksize_w=kw)
        elif mode == 1:
            nbytes = ctx.conv_packed_bytes(kh, -(-cout // 16) * 16, cin_real, 1, ksize_w=kw)
        else:
            nbytes = ctx.conv_packed_bytes(kh, 32, cin_real, 1, ksize_w=1)
        a, b = L.DeviceBuffer(nbytes), L.DeviceBuffer(nbytes)
        a.zero(); b.zero()
        if mode == 0:
            ctx.conv_pack_weights(k, kh, cin_real, cin, cout, up, a, ksize_w=kw)
        else:
            ctx.conv_pack_weights_dgrad(k, kh, kw, cin_real, cout, a, unroll_x=(mode == 2))
        items.append(L.PackItem(k.ptr, b.ptr, kh, kw, cin_real, cin, cout, up, mode, 0))
        singles.append((a, b, nbytes))
        bufs.append(k)
    table = ctx.pack_batch_prepare(items)
    ctx.pack_batch(table, len(items))
    L.stream_sync()
    for a, b, nbytes in singles:
        np.testing.assert_array_equal(a.download((nbytes,), np.uint8), b.download((nbytes,), np.uint8))


def test_wgrad_multi_equals_single_launches(ctx):
    """ssr_conv2d_wgrad_multi: the five convolutions of a dense block (one 192-channel input buffer, channel-prefix
    reads, 32-channel dZ slices and the 64-channel block gradient) in ONE launch == five single launches (same products,
    other split-K partition: fp32 summation order only), bias gradients included, scale / accumulate honoured."""
    rng = np.random.default_rng(9)
    n, h, w, cw = 3, 20, 16, 192
    x = O.bf16_round(rng.standard_normal((n, h, w, cw)).astype(np.float32))
    dx = _dev(L.f32_to_bf16_bits(x))
    specs = [(64, 32), (96, 32), (128, 32), (160, 32), (192, 64)]
    dzs = [O.bf16_round(rng.standard_normal((n, h, w, co)).astype(np.float32)) for _, co in specs]
    ddz = [_dev(L.f32_to_bf16_bits(z)) for z in dzs]
    single_w, single_b = [], []
    for (ci, co), dz in zip(specs, ddz):
        ws = L.DeviceBuffer(ctx.conv_wgrad_workspace_bytes(h, w, ci, co, 3, 3))
        dw, db = L.DeviceBuffer(9 * ci * co * 4), L.DeviceBuffer(co * 4)
        ctx.conv2d_wgrad(dx, cw, 0, ci, dz, co, 0, co, n, h, w, 3, 3, ws, dw, scale=0.2, dbias=db, bias_scale=0.2)
        L.stream_sync()
        single_w.append(dw.download((3, 3, ci, co), np.float32))
        single_b.append(db.download((co,), np.float32))
    mw = [L.DeviceBuffer.from_numpy(np.ones(9 * ci * co, np.float32)) for ci, co in specs]
    mb = [L.DeviceBuffer(co * 4) for _, co in specs]
    items = [ctx.wgrad_item(dx, cw, 0, ci, dz, co, 0, co, dw, scale=0.2, accumulate=(i == 4), dbias=(db if i != 2 else None),
                            bias_scale=0.2)
             for i, ((ci, co), dz, dw, db) in enumerate(zip(specs, ddz, mw, mb))]
    ws = L.DeviceBuffer(ctx.conv_wgrad_multi_workspace_bytes(items, h, w, 3, 3))
    ctx.conv2d_wgrad_multi(items, n, h, w, 3, 3, ws)
    L.stream_sync()
    for i, ((ci, co), ref_w, ref_b) in enumerate(zip(specs, single_w, single_b)):
        got = mw[i].download((3, 3, ci, co), np.float32)
        ref = ref_w + 1.0 if i == 4 else ref_w               # item 4 accumulates onto the ones it started from
        np.testing.assert_allclose(got, ref, rtol=1e-4, atol=1e-4 * np.abs(ref_w).max())
        if i != 2:
            np.testing.assert_allclose(mb[i].download((co,), np.float32), ref_b, rtol=1e-4, atol=1e-4 * np.abs(ref_b).max())
    # the oracle on one of them
    _, ref, _ = O.conv2d_same_backward(x[..., :160], np.zeros((3, 3, 160, 32), np.float32), dzs[3])
    assert rel_err(mw[3].download((3, 3, 160, 32), np.float32), 0.2 * ref) <= 1e-2
    with pytest.raises(Exception):
        ctx.conv2d_wgrad_multi(items * 2, n, h, w, 3, 3, ws)      # more than 8 items
