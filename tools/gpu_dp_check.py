"""torchrun --nproc-per-node N tools/gpu_dp_check.py - the data-parallel step over REAL CUDA-IPC peer mappings:
N ranks x (global batch / N) images against the single-device step on the global batch (computed by rank 0 on its own
GPU with a second, fabric-less trainer).  SRResNet with BatchNorm exercises sync-BN, the bucketed reduce-scatter + Adam
+ all-gather kernel and the metric all-reduce; a mini ESRGAN adds the discriminator's sync-BN, the RaGAN gather and its
own exchange.  Prints DP_CHECK_OK (rank 0) on success; used by tests/test_gpu_dp.py and tools/run_gpu_multi.sh."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from simplesr_b200 import discriminator as DM
    from simplesr_b200 import model_builder as MB
    from simplesr_b200 import parallel as P
    from simplesr_b200 import vgg as V
    from simplesr_b200.training import RRDBTrainer, SRResNetTrainer
    comm = P.PeerComm.connect(dist, local, 128 << 20)
    assert comm.mode == "ipc"
    gb = 2 * world
    rng = np.random.default_rng(0)
    ok = True

    def sync_all(model):
        model.stream.sync()
        dist.barrier()
        torch.cuda.synchronize()

    # ---------------- SRResNet + BatchNorm
    lr = rng.uniform(0, 1, size=(gb, 12, 12, 3)).astype(np.float32)
    hr = rng.uniform(-1, 1, size=(gb, 24, 24, 3)).astype(np.float32)
    mk = lambda: MB.build_resnet(upsample_factor=2, num_res_blocks=2, batch_normalization=True, seed=1, device=local)
    b, per = P.shard_batch(gb, rank, world)
    tr = SRResNetTrainer(mk(), loss=("mse", 1.0), learning_rate=1e-3, comm=comm, buckets=3)
    tr.prepare(per, 12, 12)
    sync_all(tr.model)
    for _ in range(3):
        tr.train_step(lr[b:b + per], hr[b:b + per], lag=1)
    tr.flush()
    md = tr.last_metrics()
    flat = tr.d_param.download((tr.count,), np.float32)
    gathered = [None] * world
    dist.all_gather_object(gathered, (md, flat))
    if rank == 0:
        ref = SRResNetTrainer(mk(), loss=("mse", 1.0), learning_rate=1e-3)
        for _ in range(3):
            ms = ref.train_step(lr, hr)
        pref = ref.d_param.download((ref.count,), np.float32)
        start = SRResNetTrainer(mk(), loss=("mse", 1.0), learning_rate=0.0).d_param.download((ref.count,), np.float32)
        for m_r, p_r in gathered:
            ok &= m_r == gathered[0][0] and np.array_equal(p_r, gathered[0][1])
        ok &= abs(md["loss"] - ms["loss"]) <= 2e-3 * abs(ms["loss"])
        a, c = flat - start, pref - start
        cos = float((a * c).sum() / (np.linalg.norm(a) * np.linalg.norm(c)))
        ok &= cos > 0.98
        print(f"srresnet+bn: loss dp {md['loss']:.6f} single {ms['loss']:.6f}, displacement cosine {cos:.4f}, "
              f"timeouts {comm.timeouts()}", flush=True)
    tr.release()
    sync_all(tr.model)
    comm.reset()

    # ---------------- mini ESRGAN
    lr = rng.uniform(0, 1, size=(gb, 16, 16, 3)).astype(np.float32)
    hr = rng.uniform(-1, 1, size=(gb, 64, 64, 3)).astype(np.float32)

    def make(cm):
        m = MB.build_enhanced_resnet(upsample_factor=4, num_rrdb_blocks=1, seed=1, device=local)
        d = DM.build_discriminator(input_dims=(64, 64), relativistic=True, seed=3, device=local)
        vl = V.VGGLoss(output_layers="block5_conv4", loss_weight=1.0, after_activation=False, seed=2, device=local)
        gl = DM.RaGANLoss(d, loss_weight=5e-3, learning_rate=1e-3)
        return RRDBTrainer(m, loss=("mae", 1e-2), learning_rate=1e-3, extra_losses=[vl, gl], comm=cm, buckets=3), gl

    tr, gl = make(comm)
    tr.prepare(per, 16, 16)
    sync_all(tr.model)
    for _ in range(2):
        tr.train_step(lr[b:b + per], hr[b:b + per], lag=1)
    tr.flush()
    md = tr.last_metrics()
    dflat = gl.d_param.download((gl.count,), np.float32)
    gathered = [None] * world
    dist.all_gather_object(gathered, (md, dflat))
    if rank == 0:
        ref, _ = make(None)
        for _ in range(2):
            ms = ref.train_step(lr, hr)
        for m_r, p_r in gathered:
            ok &= m_r == gathered[0][0] and np.array_equal(p_r, gathered[0][1])
        for k in ("loss", "mae", "vgg_loss", "ra_adversarial_loss", "ra_discriminator_loss"):
            ok &= abs(md[k] - ms[k]) <= 3e-2 * abs(ms[k]) + 1e-6
        print("esrgan dp    :", {k: round(v, 6) for k, v in md.items()}, flush=True)
        print("esrgan single:", {k: round(v, 6) for k, v in ms.items()}, flush=True)
    ok &= comm.timeouts() == 0
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("DP_CHECK_OK" if int(flag.item()) == 1 else "DP_CHECK_FAILED", flush=True)
    tr.release()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
