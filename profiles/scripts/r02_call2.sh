# round 2, call 2: full GPU suite, the default bench line (all configs + parity leg), ncu --set full of the wide-field pass
set -x
nvidia-smi -L; nproc; free -g | head -2
timeout 1500 python -m pytest tests -x -q -m gpu --durations=8 > gpurun_out/r02c2_pytest.log 2>&1; echo pytest rc=$?; tail -15 gpurun_out/r02c2_pytest.log
( time timeout 900 python bench.py > gpurun_out/r02c2_bench_default.json 2> gpurun_out/r02c2_bench_default.err ) 2>&1 | tail -3; echo bench rc=$?
tail -5 gpurun_out/r02c2_bench_default.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02c2_bench_default.json"))
r=d["roofline"]
print("cfg2 value",d["value"],"ms",d["ms_per_step"],"kernel_ms",r["kernel_ms"],"frac",r["frac"],"pass",r["whole_pass"]["ms"],"e2e",d["e2e"]["value"],"resid",d["residency"])
for k,c in d.get("configs",{}).items():
    if "error" in c: print(k,c); continue
    r=c["roofline"]
    print(k,"value",c["value"],"ms",c["ms_per_step"],"kernel_ms",r.get("kernel_ms"),"frac",r["frac"],"pass",r.get("whole_pass",{}).get("ms"),r.get("whole_pass",{}).get("frac"),"step_frac",r.get("whole_step",{}).get("frac"),"surv",c["config"]["survivors_total"],"wall",c["bench_wall_s"],"resid",c.get("residency",{}).get("first_cutout_ms"),c.get("residency",{}).get("payload_build_ms"))
print("parity",d.get("parity_check"))
print("cpu",d.get("cpu_baseline"))
PY
B="python bench.py --workload cfg4 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-parity"
timeout 300 $B > gpurun_out/r02c2_plain_cfg4.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:k1s_ -s 12 -c 3 -o gpurun_out/r02c2_prof_cfg4 -f $B > gpurun_out/r02c2_ncu_cfg4.log 2>&1; echo ncu cfg4 rc=$?
