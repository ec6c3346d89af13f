set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; cat gpurun_out/bench.log | cut -c1-600
python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-per-layer > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 1410 -c 353 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-per-layer > gpurun_out/ncu_launches.log 2>&1
python tools/gpu_profile_target.py > gpurun_out/plain_target.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:conv_tc -c 15 -o gpurun_out/prof_r01_dense python tools/gpu_profile_target.py > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
