timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_cli.py -m gpu -x -q -k "halo or cli" > gpurun_out/pytest_e2e5.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/pytest_e2e5.log
timeout 600 python bench.py --workload cfg5 --no-cpu-baseline > gpurun_out/bench_cfg5_e2e.json 2> gpurun_out/bench_cfg5_e2e.err; echo rc=$?
python - <<PY
import json
d=json.load(open("gpurun_out/bench_cfg5_e2e.json")); print(d["value"], d["ms_per_step"], d["e2e"])
PY
