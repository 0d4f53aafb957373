# cell-index kernel v2: halo parity tests + cfg5 / cfg2 quick bench
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "halo or golden" > gpurun_out/pytest_cells2.log 2>&1; echo pytest rc=$?; tail -5 gpurun_out/pytest_cells2.log
for w in cfg5 cfg2; do
  timeout 400 python bench.py --workload $w --no-cpu-baseline --no-e2e > gpurun_out/q_$w.json 2> gpurun_out/q_$w.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/q_$w.json"))
    print("$w", "value",d["value"],"ms/step",d["ms_per_step"],"kernel_ms",d["roofline"]["kernel_ms"],"frac",d["roofline"]["frac"],"pass_ms",d["roofline"]["whole_pass"]["ms"])
except Exception as e:
    print("$w failed", e); print(open("gpurun_out/q_$w.err").read()[-1500:])
PY
done
grep -i "kernel\|ms" gpurun_out/q_cfg5.err | tail -20
