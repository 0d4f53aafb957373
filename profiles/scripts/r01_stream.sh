timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_cli.py tests/test_golden.py -m gpu -x -q -k "streamed or cli or golden or exchange" > gpurun_out/pytest_stream.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/pytest_stream.log
python profiles/scripts/dbg_e2e5.py 2>&1 | tail -8
for w in cfg5 cfg4 cfg2; do
timeout 600 python bench.py --workload $w --no-cpu-baseline > gpurun_out/bench_${w}_e2e.json 2> gpurun_out/bench_${w}_e2e.err; echo rc=$?
python - <<PY
import json
d=json.load(open("gpurun_out/bench_${w}_e2e.json")); print("$w", d["value"], d["ms_per_step"], d["e2e"])
PY
done
