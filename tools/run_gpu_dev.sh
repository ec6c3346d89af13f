set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2>&1; cat gpurun_out/bench.log | cut -c1-400
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.log 2>&1
for w in srresnet_train rrdb_train esrgan_g_train esrgan_train; do python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/bench_$w.log 2>&1; done
python bench.py --workload tiled_infer --steps 2 --warmup 3 > gpurun_out/bench_tiled_infer.log 2>&1
python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-per-layer > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 1410 -c 353 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-per-layer > gpurun_out/ncu_launches.log 2>&1
python tools/gpu_profile_target.py > gpurun_out/plain_target.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:conv_tc -c 15 -o gpurun_out/prof_r01_dense python tools/gpu_profile_target.py > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
python tools/gpu_probe2.py > gpurun_out/probe2.log 2>&1
