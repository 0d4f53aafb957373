CC_K2_TMA=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -x -q -k "halo or golden or streamed or large_resident" > gpurun_out/pytest_tma.log 2>&1; echo pytest-tma rc=$?; tail -4 gpurun_out/pytest_tma.log
for v in 0 1; do CC_K2_TMA=$v timeout 300 python bench.py --no-cpu-baseline --no-e2e --steps 50 > gpurun_out/tma_$v.json 2>/dev/null; python - <<PY
import json
d=json.load(open("gpurun_out/tma_$v.json")); print("CC_K2_TMA=$v cfg2 value",d["value"],"ms/step",d["ms_per_step"],"kernel_ms",d["roofline"]["kernel_ms"],"GB/s",d["roofline"]["achieved"])
PY
done
