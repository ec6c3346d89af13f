set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python tools/gpu_probe2.py > gpurun_out/probe2.log 2>&1
timeout 300 python tests/gpu_layer_sweep.py 0,16,30,62,128 16 > gpurun_out/sweep.log 2>&1; cat gpurun_out/sweep.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench2.log 2>&1; cat gpurun_out/bench2.log
