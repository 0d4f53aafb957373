timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "halo" > gpurun_out/pytest_q5t.log 2>&1; echo pytest rc=$?; tail -2 gpurun_out/pytest_q5t.log
timeout 400 python bench.py --workload cfg5 --no-cpu-baseline --no-e2e > gpurun_out/q_cfg5.json 2> gpurun_out/q_cfg5.err
python - <<PY
import json
d=json.load(open("gpurun_out/q_cfg5.json"))
print("cfg5", "value",d["value"],"ms/step",d["ms_per_step"],"kernel_ms",d["roofline"]["kernel_ms"],"pass_ms",d["roofline"]["whole_pass"]["ms"])
PY
