# round-2 GPU session script: bash tools/run_gpu_r2.sh <stage> ...   (stages run in the order given)
#   conv     tests/test_gpu_conv.py only, stop everything if it fails
#   tests    pytest -m gpu (log to gpurun_out/pytest_gpu.log)
#   newtests only the round-2 test files
#   bench    the driver's command (headline + extras) and the reference arm
#   work     single workloads (train / tiled) for the DESIGN.md table
#   launches ncu launch list of the headline step (gpu__time_duration only)
#   full     ncu --set full of one dense block's launches
#   dp2      2-rank data-parallel check + bench --gpus 2 (needs gpurun --gpus 2)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for stage in "$@"; do
case $stage in
conv)
  # the conv kernel alone first, short timeouts: a change that hangs must not take the whole session with it
  timeout 400 python -m pytest tests/test_gpu_conv.py -m gpu -x -q --timeout 120 > gpurun_out/pytest_conv.log 2>&1; rc=$?
  echo "pytest exit=$rc" >> gpurun_out/pytest_conv.log; tail -8 gpurun_out/pytest_conv.log
  if [ $rc -ne 0 ]; then echo "conv tests failed: stopping"; exit 1; fi ;;
tests)
  timeout 1700 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$?" >> gpurun_out/pytest_gpu.log
  tail -15 gpurun_out/pytest_gpu.log ;;
newtests)
  timeout 1500 python -m pytest tests/test_gpu_dp.py tests/test_gpu_round2.py tests/test_gpu_realsize.py -m gpu -q > gpurun_out/pytest_new.log 2>&1; echo "pytest exit=$?" >> gpurun_out/pytest_new.log
  tail -40 gpurun_out/pytest_new.log ;;
smoke)
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit=$?" >> gpurun_out/smoke.log; tail -3 gpurun_out/smoke.log ;;
bench)
  timeout 800 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit=$?"; tail -c 6000 gpurun_out/bench.log; tail -5 gpurun_out/bench.err
  timeout 400 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_reference.log 2>&1; tail -c 1500 gpurun_out/bench_reference.log ;;
work)
  for w in srresnet_train rrdb_train esrgan_g_train esrgan_train; do timeout 300 python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/bench_$w.log 2>&1; tail -c 900 gpurun_out/bench_$w.log; done
  timeout 300 python bench.py --workload tiled_infer --steps 3 --warmup 3 > gpurun_out/bench_tiled_infer.log 2>&1; tail -c 900 gpurun_out/bench_tiled_infer.log ;;
launches)
  # whole process, graph nodes included; tools/launch_summary.py --last 353 keeps the final (timed) step
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-per-layer --no-extra > gpurun_out/plain_bench.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node -c 8000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-per-layer --no-extra > gpurun_out/ncu_launches.log 2>&1
  tail -2 gpurun_out/ncu_launches.log; wc -l gpurun_out/launches.csv ;;
launches_train)
  python bench.py --workload esrgan_train --steps 1 --warmup 3 > gpurun_out/plain_esrgan.log 2>&1 &&
  timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node -c 12000 --csv --log-file gpurun_out/launches_esrgan.csv python bench.py --workload esrgan_train --steps 1 --warmup 3 > gpurun_out/ncu_esrgan.log 2>&1
  tail -2 gpurun_out/ncu_esrgan.log; wc -l gpurun_out/launches_esrgan.csv ;;
full)
  python tools/gpu_profile_target.py > gpurun_out/plain_target.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 11 -c 18 -f -o gpurun_out/prof_r02_dense python tools/gpu_profile_target.py > gpurun_out/ncu_full.log 2>&1
  tail -2 gpurun_out/ncu_full.log
  ncu -i gpurun_out/prof_r02_dense.ncu-rep --page raw --csv > gpurun_out/prof_r02_dense_raw.csv 2>/dev/null
  python tools/ncu_summarise.py gpurun_out/prof_r02_dense_raw.csv gpurun_out/prof_r02_dense_summary.csv
  cut -c1-260 gpurun_out/prof_r02_dense_summary.csv ;;
full_wgrad)
  # ncu --set full of the weight-gradient kernels (and their reduction) inside the ESRGAN step graph
  python bench.py --workload esrgan_train --steps 1 --warmup 3 > gpurun_out/plain_esrgan.log 2>&1 &&
  ncu --set full --clock-control none --import-source on --graph-profiling node -k regex:wgrad -s 178 -c 8 -f -o gpurun_out/prof_r02_wgrad python bench.py --workload esrgan_train --steps 1 --warmup 3 > gpurun_out/ncu_full_wgrad.log 2>&1
  tail -2 gpurun_out/ncu_full_wgrad.log
  ncu -i gpurun_out/prof_r02_wgrad.ncu-rep --page raw --csv > gpurun_out/prof_r02_wgrad_raw.csv 2>/dev/null
  python tools/ncu_summarise.py gpurun_out/prof_r02_wgrad_raw.csv gpurun_out/prof_r02_wgrad_summary.csv
  cut -c1-260 gpurun_out/prof_r02_wgrad_summary.csv ;;
full_dense)
  # ncu --set full of the discriminator's Dense(32768 -> 1024) kernels (3xTF32 mma.sync) inside the ESRGAN step graph
  timeout 200 ncu --set full --clock-control none --import-source on --graph-profiling node -k regex:dense -s 14 -c 7 -f -o gpurun_out/prof_r02_densefc python bench.py --workload esrgan_train --steps 1 --warmup 3 > gpurun_out/ncu_full_densefc.log 2>&1
  tail -2 gpurun_out/ncu_full_densefc.log
  ncu -i gpurun_out/prof_r02_densefc.ncu-rep --page raw --csv > gpurun_out/prof_r02_densefc_raw.csv 2>/dev/null
  python tools/ncu_summarise.py gpurun_out/prof_r02_densefc_raw.csv gpurun_out/prof_r02_densefc_summary.csv
  cut -c1-260 gpurun_out/prof_r02_densefc_summary.csv ;;
dp2)
  N=${SSR_NGPU:-2}
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tools/gpu_dp_check.py > gpurun_out/dp_check_$N.log 2>&1; echo "dp_check exit=$?"; tail -8 gpurun_out/dp_check_$N.log
  timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err; echo "bench N=$N exit=$?"; tail -c 7000 gpurun_out/bench_n$N.log; tail -5 gpurun_out/bench_n$N.err ;;
*) echo "unknown stage $stage" ;;
esac
done
