"""Data-parallel training over the peer-memory fabric (csrc/comm.cu): the collectives against numpy, and the claim
BASELINE.json's north_star makes for the partition - a data-parallel step over the ranks computes the SAME function as
the single-device step on the global batch (sync-BatchNorm, global RaGAN means, gradient mean, one Adam update).

Ranks are emulated inside one process on one GPU (``PeerComm.local_group``): every rank is a trainer with its own
streams driven by its own host thread; each collective of the group runs as ONE cooperative launch over all ranks (the
same device code, blockIdx.y = rank) between the ranks' streams - kernels that wait for each other are never separate
launches on one GPU.  ``test_two_process_*`` runs the real thing (torchrun, CUDA IPC over NVLink, the collectives
inside the captured step graphs) when two GPUs are visible; tools/run_gpu_r2.sh dp2 logs it into profiles/.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

from tests.helpers import L, O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _group(world, heap=64 << 20):
    from simplesr_b200 import parallel as P
    return P.PeerComm.local_group(world, 0, heap, spin_seconds=2.0)


def _check(comms):
    for c in comms:
        assert c.timeouts() == 0


@pytest.mark.parametrize("world", [2, 4])
def test_small_allreduce_and_barrier(ctx, world):
    from simplesr_b200 import parallel as P
    comms = _group(world)
    streams = [L.Stream() for _ in range(world)]
    rng = np.random.default_rng(0)
    count = 1000
    vals = [rng.standard_normal(count).astype(np.float32) for _ in range(world)]
    srcs = [L.DeviceBuffer.from_numpy(v) for v in vals]
    dsts = [L.DeviceBuffer(count * 4) for _ in range(world)]
    sites = [c.allreduce_site(count) for c in comms]
    assert len(set(sites)) == 1                       # same slots / offsets on every rank
    def rank(r):
        for rep in range(3):                          # back-to-back calls reuse the site (epoch parity staging)
            comms[r].allreduce_f32(sites[0], srcs[r], dsts[r], count, 1.0 / world, streams[r].ptr)
        comms[r].barrier(comms[r].slots(1), streams[r].ptr)
        streams[r].sync()

    P.run_ranks([lambda r=r: rank(r) for r in range(world)])
    ref = np.zeros(count, np.float32)
    for v in vals:
        ref = ref + v                                 # rank order, like the kernel
    ref = ref * np.float32(1.0 / world)
    for dst in dsts:
        np.testing.assert_array_equal(dst.download((count,), np.float32), ref)
    _check(comms)
    for c in comms:
        c.destroy()


@pytest.mark.parametrize("world", [2, 4])
def test_comm_adam_step_equals_allreduce_then_adam(ctx, world):
    """ssr_comm_adam_step: gradient mean over the ranks (rank order), Keras Adam on the owner's shard, parameters
    identical on all ranks afterwards - against numpy; odd length (padding) and a range that starts inside the buffer."""
    from simplesr_b200 import parallel as P
    from simplesr_b200.training import FlatAdam
    comms = _group(world)
    streams = [L.Stream() for _ in range(world)]
    rng = np.random.default_rng(1)
    count = 10007
    p0 = rng.standard_normal(count).astype(np.float32)
    opts = [FlatAdam(p0, 1e-2, 0.9, 0.999, 1e-7, None, comm=c) for c in comms]
    grads = [rng.standard_normal(count).astype(np.float32) for _ in range(world)]
    m = np.zeros(count, np.float32)
    v = np.zeros(count, np.float32)
    p = p0.copy()
    for o in opts:                                    # slots in the same order on every rank, before the threads start
        o.reserve("a")
        o.reserve("b")

    def rank(r):
        o, st = opts[r], streams[r]
        o.prepare(st.ptr)
        o.update(0, 5000, st.ptr, key="a")
        o.update(5000, o.padded, st.ptr, key="b")
        st.sync()

    for t in (1, 2):
        for o, g in zip(opts, grads):
            o.d_grad.upload(g * np.float32(t))
        P.run_ranks([lambda r=r: rank(r) for r in range(world)])
        gsum = np.zeros(count, np.float32)
        for g in grads:
            gsum = gsum + g * np.float32(t)
        p, m, v = O.adam_update(p, gsum * np.float32(1.0 / world), m, v, t, lr=1e-2)
    got = [o.d_param.download((count,), np.float32) for o in opts]
    for gp in got[1:]:
        np.testing.assert_array_equal(gp, got[0])     # all ranks hold the same parameters, bit for bit
    np.testing.assert_allclose(got[0], p, rtol=2e-5, atol=2e-6)
    assert opts[0].iterations() == 2
    _check(comms)
    for c in comms:
        c.destroy()


def test_sync_batchnorm_statistics_and_ragan_means(ctx):
    """Sync-BN forward / backward sums and the relativistic losses over the global batch == the single-device kernels
    on the concatenated batch."""
    from simplesr_b200 import parallel as P
    world, px, c = 2, 300, 64
    comms = _group(world)
    streams = [L.Stream() for _ in range(world)]
    rng = np.random.default_rng(2)
    x = O.bf16_round(rng.standard_normal((world * px, c)).astype(np.float32) * 2 + 0.5)
    dy = O.bf16_round(rng.standard_normal((world * px, c)).astype(np.float32))
    gamma = (1 + 0.1 * rng.standard_normal(c)).astype(np.float32)
    beta = (0.1 * rng.standard_normal(c)).astype(np.float32)

    def run(xs, dys, sites_f, sites_b, sts):
        res = []
        bufs = []
        for i, (xx, dd) in enumerate(zip(xs, dys)):
            n_px = xx.shape[0]
            d = dict(x=L.DeviceBuffer.from_numpy(L.f32_to_bf16_bits(xx)), dy=L.DeviceBuffer.from_numpy(L.f32_to_bf16_bits(dd)),
                     ws=L.DeviceBuffer(L.load().ssr_bn_workspace_bytes(c)), mean=L.DeviceBuffer(c * 4), istd=L.DeviceBuffer(c * 4),
                     mm=L.DeviceBuffer.from_numpy(np.zeros(c, np.float32)), mv=L.DeviceBuffer.from_numpy(np.ones(c, np.float32)),
                     g=L.DeviceBuffer.from_numpy(gamma), b=L.DeviceBuffer.from_numpy(beta), y=L.DeviceBuffer(n_px * c * 2),
                     dz=L.DeviceBuffer(n_px * c * 2), sums=L.DeviceBuffer(2 * c * 4), dg=L.DeviceBuffer(c * 4), db=L.DeviceBuffer(c * 4),
                     n=n_px)
            bufs.append(d)
        def rank(i):
            d, s = bufs[i], sts[i].ptr
            L.bn_stats_bf16(d["x"], d["n"], c, 1e-3, 0.8, d["ws"], d["mean"], d["istd"], d["mm"], d["mv"], s,
                            site=sites_f[i])
            L.bn_lrelu_fwd_bf16(d["x"], d["mean"], d["istd"], d["g"], d["b"], 0.2, d["y"], d["n"], c, s)
            L.bn_lrelu_bwd_bf16(d["x"], d["dy"], d["y"], d["mean"], d["istd"], d["g"], 0.2, d["n"], c, d["ws"], d["sums"],
                                d["dg"], d["db"], False, d["dz"], s, site=sites_b[i])
            sts[i].sync()

        P.run_ranks([lambda i=i: rank(i) for i in range(len(bufs))])
        for d in bufs:
            res.append({k: d[k].download((c,), np.float32) for k in ("mean", "istd", "mm", "mv", "dg", "db")} |
                       {"y": d["y"].download((d["n"], c), np.uint16), "dz": d["dz"].download((d["n"], c), np.uint16)})
        return res

    single = run([x], [dy], [None], [None], streams[:1])[0]
    sf, sb = [cm.bn_site(c) for cm in comms], [cm.bn_site(c) for cm in comms]
    parts = run([x[:px], x[px:]], [dy[:px], dy[px:]], sf, sb, streams)
    for r in range(world):
        for k in ("mean", "istd", "mm", "mv"):
            np.testing.assert_allclose(parts[r][k], single[k], rtol=1e-6, atol=1e-7)
    y = np.concatenate([parts[0]["y"], parts[1]["y"]])
    dz = np.concatenate([parts[0]["dz"], parts[1]["dz"]])
    assert np.mean(y != single["y"]) < 1e-3          # identical up to one-ulp flips from the last bit of mean / istd
    d_a, d_b = L.bf16_bits_to_f32(dz), L.bf16_bits_to_f32(single["dz"])
    assert np.abs(d_a - d_b).max() <= 1e-2 * np.abs(d_b).max()
    np.testing.assert_allclose(parts[0]["dg"] + parts[1]["dg"], single["dg"], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(parts[0]["db"] + parts[1]["db"], single["db"], rtol=1e-4, atol=1e-4)
    # ---- RaGAN: global means, per-sample labels
    n = 3
    hc = rng.standard_normal(world * n).astype(np.float32)
    sc = rng.standard_normal(world * n).astype(np.float32)
    lh = (0.7 + 0.5 * rng.uniform(size=world * n)).astype(np.float32)
    ls = (0.3 * rng.uniform(size=world * n)).astype(np.float32)

    def ragan(hc_, sc_, lh_, ls_, site, st):
        k = hc_.size
        b = [L.DeviceBuffer.from_numpy(a) for a in (hc_, sc_, lh_, ls_)] + [L.DeviceBuffer(8)] + [L.DeviceBuffer(k * 4) for _ in range(3)]
        L.ragan_losses_ex(b[0], b[1], k, 1.0, 0.0, b[2], b[3], b[4], b[5], b[6], b[7], st.ptr, site=site)
        return b

    ref = ragan(hc, sc, lh, ls, None, streams[0])
    streams[0].sync()
    sites = [cm.ragan_site(n) for cm in comms]

    def ragan_rank(r):
        b = ragan(hc[r * n:(r + 1) * n], sc[r * n:(r + 1) * n], lh[r * n:(r + 1) * n], ls[r * n:(r + 1) * n], sites[r],
                  streams[r])
        streams[r].sync()
        return b

    outs = P.run_ranks([lambda r=r: ragan_rank(r) for r in range(world)])
    ref_o = ref[4].download((2,), np.float32)
    R = O.ragan_losses(hc, sc, hr_label=lh.astype(np.float64), sr_label=ls.astype(np.float64))
    np.testing.assert_allclose(ref_o, [R["g_loss"], R["d_loss"]], rtol=1e-5)
    for r in range(world):
        np.testing.assert_allclose(outs[r][4].download((2,), np.float32), ref_o, rtol=1e-6)
        for j in (5, 6, 7):   # own samples' gradients, times world (the ranks' gradients are averaged afterwards)
            np.testing.assert_allclose(outs[r][j].download((n,), np.float32),
                                       world * ref[j].download((world * n,), np.float32)[r * n:(r + 1) * n], rtol=1e-5,
                                       atol=1e-8)
    _check(comms)
    for cm in comms:
        cm.destroy()


def _srresnet(nb, sf, bn, seed=1):
    from simplesr_b200 import model_builder as MB
    params = O.init_srresnet_params(seed=seed, bias_std=0.05, alpha_std=0.15, upsample_factor=sf, num_res_blocks=nb)
    m = MB.build_resnet(upsample_factor=sf, num_res_blocks=nb, batch_normalization=bn, seed=0)
    by_name = {v.name: v for v in m.variables}
    for name, *_ in O.srresnet_layer_specs(upsample_factor=sf, num_res_blocks=nb):
        k, b, a = params[name]
        by_name[f"{name}/kernel:0"].assign(k)
        by_name[f"{name}/bias:0"].assign(b)
        if a is not None:
            by_name[f"{name}_prelu/alpha:0"].assign(a)
    if bn:
        rng = np.random.default_rng(5)
        for v in m.variables:
            if v.name.endswith("gamma:0"):
                v.assign((1 + 0.1 * rng.standard_normal(v.shape)).astype(np.float32))
            elif v.name.endswith("beta:0"):
                v.assign((0.1 * rng.standard_normal(v.shape)).astype(np.float32))
    return m


def _run_ranks(trainers, lr, hr, steps):
    """One host thread per emulated rank; the plans (heap offsets, barrier slots) are built first, rank by rank."""
    from simplesr_b200 import parallel as P
    world = len(trainers)
    per = lr.shape[0] // world
    for tr in trainers:
        tr.prepare(per, lr.shape[1], lr.shape[2])

    def rank(r):
        tr = trainers[r]
        for _ in range(steps):
            tr.train_step(lr[r * per:(r + 1) * per], hr[r * per:(r + 1) * per], lag=1)
        tr.flush()
        return tr.last_metrics()

    return P.run_ranks([lambda r=r: rank(r) for r in range(world)])


@pytest.mark.parametrize("bn", [False, True])
def test_dp_srresnet_step_equals_single_device_step(ctx, bn):
    """2 ranks x 2 images vs 1 device x 4 images, two Adam steps (lr > 0), with and without (sync-)BatchNorm."""
    from simplesr_b200.training import SRResNetTrainer
    nb, sf, world = 2, 2, 2
    rng = np.random.default_rng(0)
    lr = rng.uniform(0, 1, size=(4, 12, 12, 3)).astype(np.float32)
    hr = rng.uniform(-1, 1, size=(4, 24, 24, 3)).astype(np.float32)
    single = SRResNetTrainer(_srresnet(nb, sf, bn), loss=("mse", 1.0), learning_rate=1e-3)
    for _ in range(2):
        ms = single.train_step(lr, hr)
    comms = _group(world)
    ranks = [SRResNetTrainer(_srresnet(nb, sf, bn), loss=("mse", 1.0), learning_rate=1e-3, comm=c, buckets=3) for c in comms]
    md = _run_ranks(ranks, lr, hr, 2)
    _check(comms)
    for m in md:
        assert m == md[0]                                 # the metrics are means over the ranks: identical everywhere
    assert abs(md[0]["loss"] - ms["loss"]) <= 2e-3 * abs(ms["loss"]), (md[0], ms)
    assert abs(md[0]["psnr"] - ms["psnr"]) <= 2e-3 * abs(ms["psnr"])
    w1 = {v.name: v.numpy().copy() for v in single.model.variables}
    wr = [{v.name: v.numpy().copy() for v in tr.model.variables} for tr in ranks]
    start = {v.name: v.numpy().copy() for v in _srresnet(nb, sf, bn).variables}
    for name, ref in w1.items():
        np.testing.assert_array_equal(wr[0][name], wr[1][name], err_msg=name)        # replicas stay identical
        moved_ref, moved = ref - start[name], wr[0][name] - start[name]
        if np.abs(moved_ref).max() < 1e-6:
            continue
        # two Adam steps move every weight by ~2e-3 * sign(g): the displacement fields agree except where a tiny
        # gradient changes sign under a different fp32 summation order
        cos = float((moved * moved_ref).sum() / (np.linalg.norm(moved) * np.linalg.norm(moved_ref) + 1e-30))
        assert cos > 0.98, (name, cos)
        assert np.abs(wr[0][name] - ref).max() <= 4.1e-3, name
    for tr in ranks + [single]:
        tr.release()
    for c in comms:
        c.destroy()


def test_dp_esrgan_step_equals_single_device_step(ctx):
    """The full ESRGAN step (RRDB + MAE + VGG + RaGAN, discriminator update, BatchNorm in the critic) on 2 ranks x 2
    images vs 1 device x 4 images: losses, the generator's and the discriminator's weights after two steps."""
    from simplesr_b200 import discriminator as DM
    from simplesr_b200 import model_builder as MB
    from simplesr_b200 import vgg as V
    from simplesr_b200.training import RRDBTrainer
    sf, lrs, world = 4, 16, 2
    rng = np.random.default_rng(0)
    lr = rng.uniform(0, 1, size=(4, lrs, lrs, 3)).astype(np.float32)
    hr = rng.uniform(-1, 1, size=(4, lrs * sf, lrs * sf, 3)).astype(np.float32)
    vgg_model = V.build_vgg_19(seed=2)

    def make(comm=None):
        m = MB.build_enhanced_resnet(upsample_factor=sf, num_rrdb_blocks=1, seed=1)
        d = DM.build_discriminator(input_dims=(lrs * sf, lrs * sf), relativistic=True, seed=3)
        vl = V.VGGLoss(output_layers="block5_conv4", loss_weight=1.0, after_activation=False, vgg=vgg_model)
        gl = DM.RaGANLoss(d, loss_weight=5e-3, learning_rate=1e-3)
        return RRDBTrainer(m, loss=("mae", 1e-2), learning_rate=1e-3, extra_losses=[vl, gl], comm=comm, buckets=3), gl

    # a conv bias that feeds BatchNormalization has an exactly-zero gradient (the batch mean removes it): what the
    # kernels produce there is fp32 rounding noise, Adam normalises it to +-lr steps, and the direction of that noise
    # depends on the summation order - those eight biases are compared for replica identity only
    noise_only = {f"{name}/bias:0" for name, _, _, _, bn in DM.DISC_CONVS if bn}
    single, gl1 = make()
    comms, ranks = [], []
    try:
        for _ in range(2):
            ms = single.train_step(lr, hr)
        comms = _group(world, heap=256 << 20)
        made = [make(c) for c in comms]
        ranks = [t for t, _ in made]
        md = _run_ranks(ranks, lr, hr, 2)
        _check(comms)
        assert md[0] == md[1]
        for k in ("loss", "mae", "vgg_loss", "ra_adversarial_loss", "ra_discriminator_loss"):
            assert abs(md[0][k] - ms[k]) <= 3e-2 * abs(ms[k]) + 1e-6, (k, md[0][k], ms[k])
        g1 = {v.name: v.numpy().copy() for v in single.model.variables}
        g2 = [{v.name: v.numpy().copy() for v in t.model.variables} for t in ranks]
        d1 = {v.name: v.numpy().copy() for v in gl1.D.trainable_variables}
        d2 = [{v.name: v.numpy().copy() for v in gl.D.trainable_variables} for _, gl in made]
        start_g = {v.name: v.numpy() for v in MB.build_enhanced_resnet(upsample_factor=sf, num_rrdb_blocks=1,
                                                                       seed=1).variables}
        start_d = {v.name: v.numpy() for v in DM.build_discriminator(input_dims=(lrs * sf, lrs * sf), relativistic=True,
                                                                     seed=3).trainable_variables}
        for ref, got, start in ((g1, g2, start_g), (d1, d2, start_d)):
            for name, r in ref.items():
                np.testing.assert_array_equal(got[0][name], got[1][name], err_msg=name)
                a, b = got[0][name] - start[name], r - start[name]
                if np.abs(b).max() < 1e-6 or r.size < 64 or name in noise_only:
                    continue
                cos = float((a * b).sum() / (np.linalg.norm(a) * np.linalg.norm(b) + 1e-30))
                assert cos > 0.9, (name, cos)
    finally:
        # device state is released even when an assertion fires: a leaked trainer collected later would free device
        # memory in the middle of another test's graph capture
        for tr in ranks + [single]:
            tr.release()
        for c in comms:
            c.destroy()


def _gpu_count():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=30).stdout
        return sum(1 for line in out.splitlines() if line.startswith("GPU "))
    except Exception:   # noqa: BLE001
        return 0


@pytest.mark.skipif(_gpu_count() < 2, reason="needs two GPUs (the emulated-rank tests above cover the same kernels on one)")
def test_two_process_dp_over_cuda_ipc():
    """torchrun x 2: CUDA-IPC peer mapping over NVLink, DP SRResNet (sync-BN) == the single-device step."""
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29541",
                          os.path.join(ROOT, "tools", "gpu_dp_check.py")], capture_output=True, text=True, timeout=600,
                         cwd=ROOT)
    assert res.returncode == 0, (res.stdout[-2000:], res.stderr[-3000:])
    assert "DP_CHECK_OK" in res.stdout
