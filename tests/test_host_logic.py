"""Host-side logic that needs no GPU: tile sharding (also under a world_size-2 gloo group), eligibility policy."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from simplesr_b200 import evaluation as EV

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("tiles,world", [(256, 8), (256, 1), (10, 3), (3, 8), (0, 2)])
def test_tile_range_partitions_exactly(tiles, world):
    covered = []
    for r in range(world):
        b, c = EV.tile_range(tiles, r, world)
        covered.extend(range(b, b + c))
    assert covered == list(range(tiles))
    counts = [EV.tile_range(tiles, r, world)[1] for r in range(world)]
    assert max(counts) - min(counts) <= 1


def test_tile_range_rejects_bad_rank():
    with pytest.raises(ValueError):
        EV.tile_range(10, 2, 2)


def test_eligibility_policy_matches_reference():
    """evaluation.py:340-348."""
    assert EV._eligible_efficient_inference(np.zeros((1, 1001, 1001, 3), np.float32))
    assert not EV._eligible_efficient_inference(np.zeros((2, 1001, 1001, 3), np.float32))
    assert not EV._eligible_efficient_inference(np.zeros((1000, 1001, 3), np.float32))
    assert not EV._eligible_efficient_inference(np.zeros((5, 5), np.float32))


def test_sharded_stitch_world2_gloo():
    """Two processes (gloo) each stitch their tile range with the CPU oracle's placement rule; gathering the parts
    reproduces the single-process result bit for bit - the N>1 data path has no collective, only disjoint writes."""
    script = textwrap.dedent("""
        import os, sys
        import numpy as np
        import torch, torch.distributed as dist
        sys.path.insert(0, %r)
        from oracle import ssr_oracle as O
        from simplesr_b200 import evaluation as EV
        dist.init_process_group("gloo")
        rank, world = dist.get_rank(), dist.get_world_size()
        rng = np.random.default_rng(0)
        img = rng.integers(0, 255, size=(70, 90, 3)).astype(np.float32)
        p, ov = 32, 8
        tiles, padding = O.segment_into_patches(img, p, p, pixel_overlap=ov)
        b, c = EV.tile_range(tiles.shape[0], rank, world)
        cols = -(-90 // p)
        part = np.zeros_like(img)
        for t in range(b, b + c):
            r0, c0 = (t // cols) * p, (t %% cols) * p
            core = tiles[t, ov:ov + p, ov:ov + p]
            hh, ww = min(p, 70 - r0), min(p, 90 - c0)
            part[r0:r0 + hh, c0:c0 + ww] = core[:hh, :ww]
        tt = torch.from_numpy(part)
        dist.all_reduce(tt)   # test-side gather only (disjoint supports): the product path has no collective
        if rank == 0:
            assert np.array_equal(tt.numpy(), img)
            print("OK")
        dist.destroy_process_group()
    """ % ROOT)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533")
    # torchrun cannot take -c: write the script to a temp file
    import tempfile
    with tempfile.NamedTemporaryFile("w", suffix=".py", delete=False) as f:
        f.write(script)
        path = f.name
    try:
        res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                              "--master-addr", "127.0.0.1", "--master-port", "29533", path], env=env,
                             capture_output=True, text=True, timeout=240)
    finally:
        os.unlink(path)
    assert res.returncode == 0, res.stderr[-2000:]
    assert "OK" in res.stdout


def test_batch_sharding_and_gradient_mean():
    from simplesr_b200 import parallel as P
    assert [P.shard_batch(16, r, 8) for r in (0, 7)] == [(0, 2), (14, 2)]
    with pytest.raises(ValueError):
        P.shard_batch(16, 0, 3)
    rng = np.random.default_rng(0)
    g = [rng.standard_normal(7).astype(np.float32) for _ in range(4)]
    np.testing.assert_allclose(P.mean_over_ranks_numpy(g), np.mean(g, axis=0), rtol=1e-6)


def test_data_parallel_gradient_equals_global_batch_gradient():
    """The DP identity the training path relies on (oracle arithmetic, no GPU): the mean over ranks of the gradients of
    the per-rank mean losses equals the gradient of the global-batch mean loss when the shares are equal."""
    from oracle import ssr_oracle as O
    from simplesr_b200 import parallel as P
    params = O.init_srresnet_params(seed=2, bias_std=0.05, alpha_std=0.1, upsample_factor=2, num_res_blocks=1)
    rng = np.random.default_rng(0)
    lr = rng.uniform(0, 1, size=(4, 6, 6, 3)).astype(np.float32)
    hr = rng.uniform(-1, 1, size=(4, 12, 12, 3)).astype(np.float32)
    _, _, g_all = O.srresnet_loss_and_grads(params, lr, hr, upsample_factor=2, num_res_blocks=1)
    parts = []
    for r in range(2):
        b, c = P.shard_batch(4, r, 2)
        parts.append(O.srresnet_loss_and_grads(params, lr[b:b + c], hr[b:b + c], upsample_factor=2, num_res_blocks=1)[2])
    for name in g_all:
        for i in range(3):
            if g_all[name][i] is None:
                continue
            mean = P.mean_over_ranks_numpy([parts[0][name][i], parts[1][name][i]])
            np.testing.assert_allclose(mean, g_all[name][i], rtol=1e-4, atol=1e-7)


def test_ops_view_redirects_launches_to_another_stream():
    """_lib.OpsView: launches appended while ``redirect`` is set run on that stream, whatever stream the plan is
    replayed on (the fork / join structure of the training step graphs is built on this)."""
    from simplesr_b200._lib import OpsView

    class FakeStream:
        ptr = 4242

    real, seen = [], []
    ops = OpsView(real)
    ops.append(lambda s: seen.append(("main", s)))
    ops.redirect = FakeStream()
    ops.append(lambda s: seen.append(("side", s)))
    ops.redirect = None
    ops.append(lambda s: seen.append(("main2", s)))
    assert len(real) == 3
    for op in real:
        op(7)
    assert seen == [("main", 7), ("side", 4242), ("main2", 7)]


@pytest.mark.parametrize("lo,hi,world", [(0, 10008, 2), (5000, 10008, 4), (0, 16, 8), (0, 4, 3), (128, 1728 + 128, 8)])
def test_adam_shards_partition_the_range(lo, hi, world):
    """Host restatement of ssr_comm_adam_step's split: contiguous shards in units of 4 floats that tile [lo, hi)."""
    from simplesr_b200 import parallel as P
    covered = []
    for r in range(world):
        a, b = P.adam_shard_range(lo, hi, r, world)
        assert a % 4 == 0 and a <= b
        covered.extend(range(a, b))
    assert covered == list(range(lo, lo + 4 * ((hi - lo + 3) // 4)))


@pytest.mark.parametrize("h,w,world", [(2048, 2048, 8), (2048, 2048, 3), (300, 500, 2), (128, 128, 1)])
def test_tile_bands_cover_exactly_what_the_tiles_touch(h, w, world):
    """A rank uploads LR rows [src0, src0+n) and owns SR rows of its tile rows only (evaluation.tile_band)."""
    patch, ov = 128, 32
    rows, cols = -(-h // patch), -(-w // patch)
    for r in range(world):
        b, c = EV.tile_range(rows * cols, r, world)
        if c == 0:
            continue
        (s0, sn), (o0, on) = EV.tile_band(h, w, patch, ov, b, c)
        r0, r1 = b // cols, (b + c - 1) // cols
        assert s0 == max(0, r0 * patch - ov) and s0 + sn == min(h, (r1 + 1) * patch + ov)
        assert o0 == r0 * patch and o0 + on == min(h, (r1 + 1) * patch)
        assert 0 <= s0 <= o0 and o0 + on <= s0 + sn <= h


def test_piecewise_constant_decay_rule():
    """tf.keras PiecewiseConstantDecay as the ESRGAN recipe uses it (example_without_yaml.py:287-297): values[0] while
    step <= boundaries[0], the last value after the last boundary."""
    from simplesr_b200.training import PiecewiseConstantDecay
    sch = PiecewiseConstantDecay([50000, 100000, 200000, 300000], [1e-4, 5e-5, 2.5e-5, 1.25e-5, 6.25e-6])
    assert sch(0) == 1e-4 and sch(50000) == 1e-4 and sch(50001) == 5e-5 and sch(300000) == 1.25e-5
    assert sch(300001) == 6.25e-6 and sch(10 ** 9) == 6.25e-6
    with pytest.raises(ValueError):
        PiecewiseConstantDecay([1, 2], [0.1, 0.2])


def test_smoothed_labels_rule():
    """discriminator.py:236-254 (restated in simplesr_b200.discriminator.smoothed_labels): ranges and the no-smoothing
    constants the reference's tests/models/test_discriminator.py pins."""
    from simplesr_b200.discriminator import smoothed_labels
    rng = np.random.default_rng(0)
    sr, hr = smoothed_labels(rng, (50,), (50,), True, 0.3)
    assert sr.dtype == np.float64 and sr.shape == (50,) and hr.shape == (50,)
    assert 0 <= sr.min() and sr.max() <= 0.3 and sr.std() > 0
    assert 0.7 <= hr.min() and hr.max() <= 1.2 and hr.std() > 0          # 1 - offset + U(0, 0.5)
    sr, hr = smoothed_labels(rng, (50,), (7,), False, 0.0)
    assert (sr == 0).all() and (hr == 1).all() and hr.shape == (7,)


def test_adam_from_keras_config():
    """sr_model.py:121-131 builds the optimizer with ``from_config``; the learning rate may be a serialised schedule
    (tests/models/test_learnrate_scheduling.py)."""
    from simplesr_b200.sr_model import Adam
    from simplesr_b200.training import PiecewiseConstantDecay
    a = Adam.from_config({"learning_rate": {"class_name": "PiecewiseConstantDecay",
                                            "config": {"boundaries": [2, 5], "values": [3e-4, 2e-5, 3e-6]}},
                          "beta_1": 0.5, "beta_2": 0.8, "name": "Adam", "amsgrad": False})
    assert isinstance(a.learning_rate, PiecewiseConstantDecay) and a.learning_rate.boundaries == [2, 5]
    assert [a.learning_rate(s) for s in range(7)] == [3e-4, 3e-4, 3e-4, 2e-5, 2e-5, 2e-5, 3e-6]
    assert (a.beta_1, a.beta_2, a.epsilon) == (0.5, 0.8, 1e-7)
    assert Adam.from_config({"learning_rate": 1e-4}).learning_rate == 1e-4
    with pytest.raises(ValueError):
        Adam.from_config({"learning_rate": {"class_name": "ExponentialDecay", "config": {}}})


def test_loss_functions_from_yaml():
    """utils/config/yaml_helper.py:44-51: class names of the loss_functions table -> functor objects with their kwargs."""
    import os
    from simplesr_b200 import generator as G
    conf = G.load_yaml(os.path.join(os.path.dirname(__file__), "golden", "minimal_example.yaml"))
    assert conf["general"]["crop_size"] == (80, 80, 3)                         # !!python/tuple as in the reference file
    fns = G.init_loss_functions_from_yaml(conf["model"]["generator"])
    assert [type(f) for f in fns] == [G.MeanSquaredError] and fns[0].loss_weight == 1.0 and not fns[0].weighted
    fns = G.init_loss_functions_from_yaml({"loss_functions": [
        {"loss_function": "MeanAbsoluteError", "weighted": True, "loss_weight": 0.01},
        {"loss_function": "RaAdversarialLoss", "weighted": True, "loss_weight": 5e-3},
        {"loss_function": "AdversarialLoss"}, {"loss_function": "DiscriminatorLoss"},
        {"loss_function": "RaDiscriminatorLoss", "track_metrics": False}]})
    assert [f.name for f in fns] == ["mean_absolute_error", "ra_adversarial_loss", "adversarial_loss",
                                     "discriminator_loss", "ra_discriminator_loss"]
    assert fns[0].weighted and fns[0].loss_weight == 0.01 and fns[4].track_metrics is False
    with pytest.raises(AttributeError):
        G.init_loss_functions_from_yaml({"loss_functions": [{"loss_function": "HingeLoss"}]})
    assert G.load_yaml(conf) is conf


def test_clock_sampler_windows_and_extension(tmp_path, monkeypatch):
    """bench.ClockSampler against a stand-in nvidia-smi (same CSV line format, one line per 50 ms, slow to come up):
    samples are kept by their own timestamps inside the marked window, a window shorter than MIN_WINDOW_S is extended
    by running more untimed steps, and the throttle reasons are decoded."""
    import stat
    import sys
    import time
    fake = tmp_path / "nvidia-smi"
    marker = tmp_path / "under_load"
    fake.write_text(f"""#!{sys.executable}
import datetime, os, sys, time
time.sleep(0.25)                                  # NVML start-up
while True:
    now = datetime.datetime.now().strftime("%Y/%m/%d %H:%M:%S.%f")[:-3]
    load = os.path.exists({str(marker)!r})        # idle clocks until the test puts the GPU "under load"
    clock, cap = (1750, "Active") if load else (300, "Not Active")
    print(f"{{now}}, {{clock}}, 1965, 812.40, 0x0000000000000004, Not Active, Not Active, Not Active, {{cap}}", flush=True)
    time.sleep(0.05)
""")
    fake.chmod(fake.stat().st_mode | stat.S_IEXEC)
    monkeypatch.setenv("PATH", str(tmp_path) + os.pathsep + os.environ.get("PATH", ""))
    import bench
    sm = bench.ClockSampler(0)
    sm.start()
    time.sleep(1.0)                                 # "warm-up": nvidia-smi comes up, idle-clock samples land before begin()
    marker.write_text("1")
    time.sleep(0.12)
    sm.begin()
    time.sleep(0.1)                                 # a timed region shorter than the minimum window
    sm.end()
    extra = sm.extend_until(lambda: time.sleep(0.02))
    assert extra >= 5                               # ~0.5 s of additional load at >= 20 ms per step
    c = sm.stop()
    assert c["samples"] >= 2 and c["sm_mhz"] == 1750.0 and c["sm_max_mhz"] == 1965.0      # no idle sample inside
    assert c["reasons"] == ["sw_power_cap"] and "ms of the measured load" in c["window"]
    # without marks every sample counts (the other workloads' use); a missing binary is reported, not raised
    sm.start()
    time.sleep(1.0)
    c = sm.stop()
    assert c["samples"] >= 2 and c["window"] == "start() to stop()"
    monkeypatch.setenv("PATH", str(tmp_path / "nowhere"))
    sm = bench.ClockSampler(0)
    sm.start()
    assert sm.extend_until(lambda: None) == 0
    c = sm.stop()
    assert c["sm_mhz"] is None and c["samples"] == 0 and c["reasons"] == ["nvidia-smi unavailable"]


def test_evaluation_dispatch_and_model_loading(monkeypatch, tmp_path, capsys):
    """evaluate_on_testdata's per-batch rule (evaluation.py:253-277) and _load_model (:320-327) - host logic only: a
    stand-in model (nearest-neighbour x2) for the direct path, the tiled path's arguments checked through a stub."""
    from simplesr_b200 import evaluation as EV

    def model(x, training=False):
        assert training is False
        return np.repeat(np.repeat(np.asarray(x, np.float32), 2, axis=1), 2, axis=2)

    small = np.random.default_rng(0).uniform(0, 1, size=(3, 20, 24, 3)).astype(np.float32)
    out = EV.upscale(model, small)
    assert out.shape == (3, 40, 48, 3)
    np.testing.assert_array_equal(out, model(small))
    calls = []

    def fake_tiled(m, lr, patch, pixel_overlap, **kw):
        calls.append((np.shape(lr), patch, pixel_overlap, kw))
        h, w = np.shape(lr)[-3:-1]
        return np.zeros((2 * h, 2 * w, 3), np.float32)

    monkeypatch.setattr(EV, "upscale_tiled", fake_tiled)
    big = np.zeros((1, 1001, 1002, 3), np.float32)
    assert EV.upscale(model, big, rank=1, world_size=2).shape == (1, 2002, 2004, 3)
    assert calls == [((1, 1001, 1002, 3), 128, 32, {"rank": 1, "world_size": 2})]
    assert EV.upscale(model, np.zeros((1, 1000, 1002, 3), np.float32)).shape == (1, 2000, 2004, 3)     # not above 1000
    assert EV.upscale(model, np.zeros((2, 1001, 1002, 3), np.float32)).shape == (2, 2002, 2004, 3)     # a real batch
    assert len(calls) == 1
    assert EV.upscale(model, np.zeros((1, 300, 400, 3), np.float32), segmentation_min_width=200,
                      segmentation_min_height=200).shape == (1, 600, 800, 3)
    assert len(calls) == 2
    for name in ("missing_gen_3.h5", "missing_gen_3"):
        with pytest.raises(SystemExit) as e:
            EV._load_model(str(tmp_path / name))
        assert e.value.code == 1
        assert "could not locate model" in capsys.readouterr().out
