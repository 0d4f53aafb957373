# k3_place per segment + row->record map for the gather: GPU suite (without the two exhaustive float sweeps) and cfg5 lines
set -x
timeout 1500 python -m pytest tests -q -m gpu -k "not all_floats" -x > gpurun_out/r02c25_pytest.log 2>&1; echo pytest rc=$?; tail -4 gpurun_out/r02c25_pytest.log
for w in cfg5_shared cfg5 cfg5_onepass; do
timeout 600 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-parity --no-configs > gpurun_out/r02c25_bench_$w.json 2> gpurun_out/r02c25_bench_$w.err; echo bench $w rc=$?
python - <<PY
import json
d=json.load(open("gpurun_out/r02c25_bench_$w.json")); r=d["roofline"]
print("$w", "value", d["value"], "ms", d["ms_per_step"], "kernel", r.get("kernel"), r.get("kernel_ms"), "pass", r.get("whole_pass",{}).get("ms"), "halos/s", d.get("halos_per_s"), "read", r.get("particles_read_per_launch"), "idx_ms", d["residency"].get("index_build_ms"))
PY
done
B="python bench.py --workload cfg5_shared --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-parity --no-configs"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 120 --csv --log-file gpurun_out/r02c25_launches_cfg5_shared.csv $B > gpurun_out/r02c25_ncu.log 2>&1; echo launches rc=$?
