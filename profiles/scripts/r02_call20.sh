# k3_small counting path rewritten (owner loads hoisted, unrolled counting, rows staged in shared memory): suite + cfg2 A/B
set -x
timeout 1500 python -m pytest tests -q -m gpu -k "not all_floats" -x > gpurun_out/r02c20_pytest.log 2>&1; echo pytest rc=$?; tail -4 gpurun_out/r02c20_pytest.log
for w in cfg2 cfg1 cfg5; do
timeout 600 python bench.py --workload $w --steps 50 --warmup 5 --no-cpu-baseline --no-e2e --no-parity --no-configs > gpurun_out/r02c20_bench_$w.json 2> gpurun_out/r02c20_bench_$w.err; echo bench $w rc=$?
python - <<PY
import json
d=json.load(open("gpurun_out/r02c20_bench_$w.json")); r=d["roofline"]
print("$w", "value", d["value"], "ms", d["ms_per_step"], "kernel", r.get("kernel_ms"), "pass", r.get("whole_pass",{}).get("ms"))
PY
done
B="python bench.py --workload cfg2 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-parity --no-configs"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02c20_launches_cfg2.csv $B > gpurun_out/r02c20_ncu_cfg2.log 2>&1; echo launches rc=$?
grep k3_small gpurun_out/r02c20_launches_cfg2.csv | head -4
timeout 900 ncu --set full --cache-control none --clock-control none --import-source on -k regex:k3_small -s 3 -c 1 -o gpurun_out/r02c20_prof_k3s_warm -f $B > gpurun_out/r02c20_ncu_k3s_warm.log 2>&1; echo full k3s warm rc=$?
