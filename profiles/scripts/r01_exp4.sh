run() { # name, env..., workload
  name=$1; shift; w=$1; shift
  env "$@" timeout 300 python bench.py --workload $w --no-cpu-baseline --no-e2e > gpurun_out/x_$name.json 2>/dev/null; python - <<PY
import json
d=json.load(open("gpurun_out/x_$name.json")); print("$name", "$w", "value",d["value"],"ms/step",d["ms_per_step"],"kernel_ms",d["roofline"]["kernel_ms"],"frac",d["roofline"]["frac"])
PY
}
for w in const1b cfg4; do
run base_$w $w A=1
run nocarve_$w $w CC_NO_CARVEOUT=1
run drain1_$w $w CC_K1_DRAIN=1
run drain1_nocarve_$w $w CC_K1_DRAIN=1 CC_NO_CARVEOUT=1
run drain8_$w $w CC_K1_DRAIN=8
run minb3_$w $w CC_K1_MINB=3
done
run cfg2_base cfg2 A=1
run cfg2_nocarve cfg2 CC_NO_CARVEOUT=1
