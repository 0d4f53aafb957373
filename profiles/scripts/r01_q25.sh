timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -x -q -k "halo or golden or exchange or streamed" > gpurun_out/pytest_q.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/pytest_q.log
for w in cfg2 cfg2 cfg5 cfg1; do
timeout 400 python bench.py --workload $w --no-cpu-baseline --no-e2e > gpurun_out/q_$w.json 2> gpurun_out/q_$w.err
python - <<PY
import json
d=json.load(open("gpurun_out/q_$w.json"))
print("$w", "value",d["value"],"ms/step",d["ms_per_step"],"kernel_ms",d["roofline"]["kernel_ms"],"frac",d["roofline"]["frac"],"pass_ms",d["roofline"]["whole_pass"]["ms"])
PY
done
