timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_tq.log 2>&1; echo pytest rc=$?; tail -4 gpurun_out/pytest_gpu_tq.log
bash profiles/scripts/r01_quick.sh
