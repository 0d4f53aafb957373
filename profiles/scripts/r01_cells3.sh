# cell-index kernel: halo parity tests with both CTA shapes + cfg5 bench A/B + profile of the default
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "halo or golden" > gpurun_out/pytest_cells3.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/pytest_cells3.log
CC_K2C_WARPS=24 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "halo" > gpurun_out/pytest_cells3_w24.log 2>&1; echo pytest24 rc=$?; tail -3 gpurun_out/pytest_cells3_w24.log
for v in 32 24; do
  CC_K2C_WARPS=$v timeout 400 python bench.py --workload cfg5 --no-cpu-baseline --no-e2e > gpurun_out/q_cfg5_w$v.json 2> gpurun_out/q_cfg5_w$v.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/q_cfg5_w$v.json"))
    print("cfg5 warps $v", "value",d["value"],"ms/step",d["ms_per_step"],"kernel_ms",d["roofline"]["kernel_ms"],"frac",d["roofline"]["frac"],"pass_ms",d["roofline"]["whole_pass"]["ms"])
except Exception as e:
    print("failed", e); print(open("gpurun_out/q_cfg5_w$v.err").read()[-1500:])
PY
done
bash profiles/scripts/r01_cfg5b.sh
