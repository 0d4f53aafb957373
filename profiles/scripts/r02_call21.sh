# shared pass for several box lengths (processLC_boxes / cfg5_shared): CLI + at-size parity + bench
set -x
timeout 1500 python -m pytest tests/test_cli.py tests/test_gpu_parity.py -q -m gpu -k "box_lengths or many_halos or small_cutout" -x > gpurun_out/r02c21_pytest.log 2>&1; echo pytest rc=$?; tail -4 gpurun_out/r02c21_pytest.log
for w in cfg5_shared cfg5 cfg2; do
timeout 600 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-parity --no-configs > gpurun_out/r02c21_bench_$w.json 2> gpurun_out/r02c21_bench_$w.err; echo bench $w rc=$?
python - <<PY
import json
d=json.load(open("gpurun_out/r02c21_bench_$w.json")); r=d["roofline"]
print("$w", "value", d["value"], "ms", d["ms_per_step"], "kernel", r.get("kernel"), r.get("kernel_ms"), "pass", r.get("whole_pass",{}).get("ms"), "halos/s", d.get("halos_per_s"), "surv", d["config"]["survivors_total"])
PY
done
B="python bench.py --workload cfg5_shared --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-parity --no-configs"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r02c21_launches_cfg5_shared.csv $B > gpurun_out/r02c21_ncu.log 2>&1; echo launches rc=$?
