"""Shared helpers for the parity tests: drive the C ABI on seeded inputs and compare with the oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import ssr_oracle as O  # noqa: E402
from simplesr_b200 import _lib as L  # noqa: E402


def has_gpu():
    try:
        import ctypes
        lib = L.load()
        h = ctypes.c_void_p()
        rc = lib.ssr_ctx_create(0, ctypes.byref(h))
        if rc == 0:
            lib.ssr_ctx_destroy(h)
        return rc == 0
    except Exception:
        return False


def rel_err(got, ref):
    """max |got-ref| / max |ref| — the per-layer metric of BASELINE.json (<= 1e-2)."""
    ref = np.asarray(ref, np.float64)
    got = np.asarray(got, np.float64)
    return float(np.max(np.abs(got - ref)) / max(np.max(np.abs(ref)), 1e-30))


def conv_case(ctx, n, h, w, cin_real, cout, ksize=3, act=L.ACT_NONE, act_alpha=0.2, res=False, res_beta=0.2, up=1,
              out_dtype=L.SSR_BF16, in_cstride=None, out_cstride=None, out_coff=0, seed=0, bias_scale=0.5,
              x_scale=1.0, tile_order=0):
    """Run one conv through ssr_conv2d_fwd and through the oracle on identical (bf16-rounded) inputs.
    Returns (got fp32 [n,oh,ow,cout'], ref fp32)."""
    rng = np.random.default_rng(seed)
    cin = -(-cin_real // 16) * 16
    in_cstride = in_cstride or cin
    cout_store = cout // 4 if up == 2 else cout
    out_cstride = out_cstride or (cout_store + out_coff)
    x = O.bf16_round(rng.uniform(-1, 1, size=(n, h, w, in_cstride)).astype(np.float32) * x_scale)
    if cin_real < cin:
        x[..., cin_real:cin] = rng.uniform(-1, 1, size=(n, h, w, cin - cin_real))  # must be ignored (zero weights)
        x = O.bf16_round(x)
    k = O.bf16_round(rng.standard_normal((ksize, ksize, cin_real, cout)).astype(np.float32) / np.sqrt(ksize * ksize * cin_real))
    b = (rng.standard_normal(cout) * bias_scale).astype(np.float32)
    alpha = (rng.uniform(0.05, 0.5, size=cout_store)).astype(np.float32) if act == L.ACT_PRELU else None
    oh, ow = h * up, w * up
    r = O.bf16_round(rng.uniform(-1, 1, size=(n, h, w, cout)).astype(np.float32)) if res else None

    # ---- oracle
    y = O.conv2d_same(x[..., :cin_real], k, b)
    if act == L.ACT_LRELU:
        y = O.leaky_relu(y, act_alpha)
    elif act == L.ACT_TANH:
        y = np.tanh(y)
    elif act == L.ACT_RELU:
        y = np.maximum(y, 0)
    elif act == L.ACT_PRELU:
        a_full = np.tile(alpha, 4) if up == 2 else alpha
        y = O.prelu(y, a_full)
    if res:
        y = r + np.float32(res_beta) * y
    if up == 2:
        y = O.depth_to_space(y, 2)
    ref = y.astype(np.float32)

    # ---- device
    dx = L.DeviceBuffer.from_numpy(L.f32_to_bf16_bits(x))
    dw = L.DeviceBuffer.from_numpy(k)
    db = L.DeviceBuffer.from_numpy(b)
    da = L.DeviceBuffer.from_numpy(alpha) if alpha is not None else None
    dr = L.DeviceBuffer.from_numpy(L.f32_to_bf16_bits(r)) if res else None
    packed = L.DeviceBuffer(ctx.conv_packed_bytes(ksize, cin, cout, up))
    ctx.conv_pack_weights(dw, ksize, cin_real, cin, cout, up, packed)
    esz = 2 if out_dtype == L.SSR_BF16 else 4
    dout = L.DeviceBuffer(n * oh * ow * out_cstride * esz)
    dout.zero()
    d = L.ConvDesc(n=n, h=h, w=w, cin=cin, in_cstride=in_cstride, cout=cout, ksize=ksize, act=act,
                   act_alpha=act_alpha, res_beta=res_beta, up=up, out_dtype=out_dtype, out_cstride=out_cstride,
                   out_coff=out_coff, res_dtype=(L.SSR_BF16 if res else L.SSR_NONE), res_cstride=cout, res_coff=0,
                   out2_cstride=0, out2_coff=0, tile_order=tile_order)
    ctx.conv2d_fwd(d, dx, packed, db, dout, alpha=da, res=dr)
    L.stream_sync()
    if out_dtype == L.SSR_BF16:
        got = L.bf16_bits_to_f32(dout.download((n, oh, ow, out_cstride), np.uint16))
    else:
        got = dout.download((n, oh, ow, out_cstride), np.float32)
    untouched = got[..., :out_coff]
    got = got[..., out_coff:out_coff + cout_store]
    for buf in (dx, dw, db, da, dr, packed, dout):
        if buf is not None:
            buf.free()
    return got, ref, untouched
