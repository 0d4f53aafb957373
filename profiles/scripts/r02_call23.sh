# particle index of the resident step (k2p_*): parity tests, then cfg5 / cfg5_shared with and without it
set -x
timeout 1500 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "index or many_halos or cell_index" -x > gpurun_out/r02c23_pytest.log 2>&1; echo pytest rc=$?; tail -6 gpurun_out/r02c23_pytest.log
for w in cfg5_shared cfg5; do
for idx in auto; do
CC_INDEX=$idx timeout 600 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-parity --no-configs > gpurun_out/r02c23_bench_${w}_$idx.json 2> gpurun_out/r02c23_bench_${w}_$idx.err; echo bench $w $idx rc=$?
python - <<PY
import json
d=json.load(open("gpurun_out/r02c23_bench_${w}_$idx.json")); r=d["roofline"]
print("$w $idx", "value", d["value"], "ms", d["ms_per_step"], "kernel", r.get("kernel"), r.get("kernel_ms"), "pass", r.get("whole_pass",{}).get("ms"), "halos/s", d.get("halos_per_s"), "surv", d["config"]["survivors_total"], d.get("residency"))
PY
done
done
B="python bench.py --workload cfg5_shared --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-parity --no-configs"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 120 --csv --log-file gpurun_out/r02c23_launches_cfg5_shared.csv $B > gpurun_out/r02c23_ncu.log 2>&1; echo launches rc=$?
