set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_v9.log 2>&1; echo pytest rc=$?; tail -4 gpurun_out/pytest_gpu_v9.log
( time timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err ) 2>&1 | tail -3; cat gpurun_out/bench_default.json; tail -3 gpurun_out/bench_default.err
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err ) 2>&1 | tail -3; cat gpurun_out/bench_reference.json; tail -3 gpurun_out/bench_reference.err
lscpu | head -20
