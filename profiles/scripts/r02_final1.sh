# round 2 evidence, 1 GPU: the whole GPU suite, smoke, the default bench line + reference arm, ncu launch lists and full sets
set -x
timeout 2400 python -m pytest tests -q -m gpu --durations=6 > gpurun_out/r02f1_pytest.log 2>&1; echo pytest rc=$?; tail -12 gpurun_out/r02f1_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02f1_smoke.log 2>&1; echo smoke rc=$?; tail -2 gpurun_out/r02f1_smoke.log
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02f1_bench_default.json 2> gpurun_out/r02f1_bench_default.err ) 2>&1 | tail -3; echo bench rc=$?
tail -3 gpurun_out/r02f1_bench_default.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02f1_bench_default.json"))
r=d["roofline"]
print("cfg2 value",d["value"],"ms",d["ms_per_step"],"kernel_ms",r["kernel_ms"],"frac",r["frac"],"pass",r["whole_pass"]["ms"],"e2e",d["e2e"]["value"],d["e2e"].get("roofline",{}).get("frac"),"launches",d["gpu_launches"])
for k,c in d.get("configs",{}).items():
    if "error" in c: print(k,c); continue
    r=c["roofline"]
    print(k,"value",c["value"],"ms",c["ms_per_step"],"kernel_ms",r.get("kernel_ms"),"frac",r["frac"],"pass",r.get("whole_pass",{}).get("ms"),r.get("whole_pass",{}).get("frac"),"step_frac",r.get("whole_step",{}).get("frac"),"surv",c["config"]["survivors_total"],c.get("halos_per_s"),"first",c.get("residency",{}).get("first_cutout_ms"))
print("parity ok", d.get("parity_check",{}).get("ok"), "cpu", d.get("cpu_baseline",{}).get("value"))
PY
( time timeout 900 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/r02f1_bench_reference.json 2> gpurun_out/r02f1_bench_reference.err ) 2>&1 | tail -3; echo ref rc=$?; cut -c1-400 gpurun_out/r02f1_bench_reference.json
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active"
for w in cfg2 const1b cfg4 cfg5 cfg5_shared cfg1; do
  B="python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-parity --no-configs"
  timeout 600 ncu --metrics $M --clock-control none -c 160 --csv --log-file gpurun_out/r02f1_launches_$w.csv $B > gpurun_out/r02f1_ncu_$w.log 2>&1; echo launches $w rc=$?
done
B="python bench.py --workload cfg2 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-parity --no-configs"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k2_cutout_halo -s 3 -c 1 -o gpurun_out/r02f1_prof_cfg2 -f $B > gpurun_out/r02f1_ncu_full_cfg2.log 2>&1; echo full cfg2 rc=$?
B="python bench.py --workload const1b --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-parity --no-configs"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k1s_ -s 9 -c 3 -o gpurun_out/r02f1_prof_const1b -f $B > gpurun_out/r02f1_ncu_full_const1b.log 2>&1; echo full const1b rc=$?
B="python bench.py --workload cfg4 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-parity --no-configs"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k1s_ -s 12 -c 3 -o gpurun_out/r02f1_prof_cfg4 -f $B > gpurun_out/r02f1_ncu_full_cfg4.log 2>&1; echo full cfg4 rc=$?
B="python bench.py --workload cfg5 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-parity --no-configs"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k2p_query|k3_bucket|k3_rank|k3_gather_halo" -s 16 -c 4 -o gpurun_out/r02f1_prof_cfg5 -f $B > gpurun_out/r02f1_ncu_full_cfg5.log 2>&1; echo full cfg5 rc=$?
