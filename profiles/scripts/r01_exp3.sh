bash profiles/scripts/r01_bench_all.sh v6
for m in 3; do for w in const1b cfg4; do CC_K1_MINB=$m timeout 300 python bench.py --workload $w --no-cpu-baseline --no-e2e > gpurun_out/minb_${m}_$w.json 2>/dev/null; python - <<PY
import json
d=json.load(open("gpurun_out/minb_${m}_$w.json")); print("K1_MINB=$m $w value",d["value"],"ms/step",d["ms_per_step"],"kernel_ms",d["roofline"]["kernel_ms"],"frac",d["roofline"]["frac"])
PY
done; done
