for w in cfg5 cfg2; do
timeout 400 python bench.py --workload $w --no-cpu-baseline --no-e2e > gpurun_out/q_$w.json 2> gpurun_out/q_$w.err
python - <<PY
import json
d=json.load(open("gpurun_out/q_$w.json"))
print("$w", "value",d["value"],"ms/step",d["ms_per_step"],"kernel_ms",d["roofline"]["kernel_ms"],"frac",d["roofline"]["frac"],"pass_ms",d["roofline"]["whole_pass"]["ms"])
PY
done
