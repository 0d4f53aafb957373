# usage (on the GPU box): bash profiles/scripts/r01_bench_all.sh [tag]
tag=${1:-run}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$tag.log 2>&1; echo pytest rc=$?; tail -5 gpurun_out/pytest_gpu_$tag.log
for w in cfg2 const1b cfg4 cfg1 cfg5; do
  timeout 400 python bench.py --workload $w --no-cpu-baseline > gpurun_out/bench_${w}_$tag.json 2> gpurun_out/bench_${w}_$tag.err; echo $w rc=$?
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_${w}_$tag.json"))
    print("$w", "value",d["value"],"ms/step",d["ms_per_step"],"roof",d["roofline"]["achieved"],d["roofline"]["frac"],"kernel_ms",d["roofline"]["kernel_ms"],"e2e",d["e2e"]["value"],"surv",d["config"]["survivors_total"],"launches",d["gpu_launches"])
except Exception as e:
    print("$w failed", e); print(open("gpurun_out/bench_${w}_$tag.err").read()[-1500:])
PY
done
