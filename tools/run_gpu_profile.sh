# one ncu pass per gpurun call: $1 = launches | full
cd $GRAFT_REPO_ROOT
if [ "$1" = launches ]; then
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-per-layer > gpurun_out/plain_bench.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -s 1410 -c 353 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-per-layer > gpurun_out/ncu_launches.log 2>&1
  tail -2 gpurun_out/ncu_launches.log; wc -l gpurun_out/launches.csv
else
  python tools/gpu_profile_target.py > gpurun_out/plain_target.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 11 -c 18 -f -o gpurun_out/prof_r01_dense python tools/gpu_profile_target.py > gpurun_out/ncu_full.log 2>&1
  tail -2 gpurun_out/ncu_full.log
  ncu -i gpurun_out/prof_r01_dense.ncu-rep --page raw --csv > gpurun_out/prof_r01_dense_raw.csv 2>/dev/null
  python tools/ncu_summarise.py gpurun_out/prof_r01_dense_raw.csv gpurun_out/prof_r01_dense_summary.csv
  cut -c1-260 gpurun_out/prof_r01_dense_summary.csv
fi
