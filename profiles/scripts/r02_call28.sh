# host spins on the flag k3_small raises instead of cudaStreamSynchronize: tests + same-box A/B on cfg2
set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "small_cutout or overflow or alternating or halo_single or nan_theta or slots or streamed" -x > gpurun_out/r02c28_pytest.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/r02c28_pytest.log
for rep in 1 2; do
for mode in flag sync; do
if [ $mode = sync ]; then export CC_NO_FLAG_WAIT=1; else unset CC_NO_FLAG_WAIT; fi
timeout 600 python bench.py --workload cfg2 --steps 200 --warmup 10 --no-cpu-baseline --no-e2e --no-parity --no-configs > gpurun_out/r02c28_bench_cfg2_${mode}_$rep.json 2> gpurun_out/r02c28_bench_cfg2_${mode}_$rep.err; echo bench $mode rc=$?
python - <<PY
import json
d=json.load(open("gpurun_out/r02c28_bench_cfg2_${mode}_$rep.json")); r=d["roofline"]
print("cfg2 $mode $rep", "value", d["value"], "ms", d["ms_per_step"], "kernel", r.get("kernel_ms"), "pass", r.get("whole_pass",{}).get("ms"))
PY
done
done
unset CC_NO_FLAG_WAIT
