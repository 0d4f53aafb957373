bash profiles/scripts/r01_bench_all.sh v7
run() { name=$1; shift; w=$1; shift
  env "$@" timeout 300 python bench.py --workload $w --no-cpu-baseline --no-e2e > gpurun_out/x_$name.json 2>/dev/null; python - <<PY
import json
d=json.load(open("gpurun_out/x_$name.json")); print("$name", "$w", "value",d["value"],"ms/step",d["ms_per_step"],"kernel_ms",d["roofline"]["kernel_ms"],"frac",d["roofline"]["frac"])
PY
}
for w in const1b cfg4; do run drain1_$w $w CC_K1_DRAIN=1; run drain16_$w $w CC_K1_DRAIN=16; done
