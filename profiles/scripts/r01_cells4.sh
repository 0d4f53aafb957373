# halo parity tests + cfg5/cfg2 bench + launch list of cfg5
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "halo or golden or exchange" > gpurun_out/pytest_cells4.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/pytest_cells4.log
for v in 24 20 16; do
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
CC_K2C_WARPS=24 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_cfg5_w24.csv python bench.py --workload cfg5 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_l.log 2>&1; echo ncu rc=$?
