bash profiles/scripts/r01_bench_all.sh v4
for g in 32 64 128; do CC_L2_FETCH=$g timeout 300 python bench.py --workload const1b --no-cpu-baseline --no-e2e > gpurun_out/l2f_$g.json 2>/dev/null; python - <<PY
import json
d=json.load(open("gpurun_out/l2f_$g.json")); print("L2_FETCH=$g const1b value",d["value"],"ms/step",d["ms_per_step"],"kernel_ms",d["roofline"]["kernel_ms"])
PY
CC_L2_FETCH=$g timeout 300 python bench.py --workload cfg4 --no-cpu-baseline --no-e2e > gpurun_out/l2f4_$g.json 2>/dev/null; python - <<PY
import json
d=json.load(open("gpurun_out/l2f4_$g.json")); print("L2_FETCH=$g cfg4 value",d["value"],"ms/step",d["ms_per_step"],"kernel_ms",d["roofline"]["kernel_ms"])
PY
done
CC_NO_BINS=1 timeout 300 python bench.py --workload cfg5 --no-cpu-baseline --no-e2e --steps 3 > gpurun_out/cfg5_nobins.json 2>/dev/null; python - <<PY
import json
d=json.load(open("gpurun_out/cfg5_nobins.json")); print("cfg5 no bins value",d["value"],"ms/step",d["ms_per_step"],"surv",d["config"]["survivors_total"])
PY
