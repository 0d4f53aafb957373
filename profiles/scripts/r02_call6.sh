# round 2, call 6: launch-tail work (zero-copy results, self-cleaning state, counting sort in k3_small, no timing events)
set -x
timeout 1500 python -m pytest tests -q -m gpu -x --deselect tests/test_gpu_parity.py::test_device_libm_equals_host_restatement_on_all_floats --deselect tests/test_gpu_parity.py::test_device_arcsec_scaling_equals_specification_on_all_floats > gpurun_out/r02c6_pytest.log 2>&1; echo pytest rc=$?; tail -6 gpurun_out/r02c6_pytest.log
( time timeout 900 python bench.py > gpurun_out/r02c6_bench_default.json 2> gpurun_out/r02c6_bench_default.err ) 2>&1 | tail -3; echo bench rc=$?
tail -5 gpurun_out/r02c6_bench_default.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02c6_bench_default.json"))
r=d["roofline"]
print("cfg2 value",d["value"],"ms",d["ms_per_step"],"kernel_ms",r["kernel_ms"],"frac",r["frac"],"pass",r["whole_pass"]["ms"],"e2e",d["e2e"]["value"],"resid",d["residency"]["resident_ms"],d["residency"]["first_cutout_ms"])
for k,c in d.get("configs",{}).items():
    if "error" in c: print(k,c); continue
    r=c["roofline"]
    print(k,"value",c["value"],"ms",c["ms_per_step"],"kernel_ms",r.get("kernel_ms"),"frac",r["frac"],"pass",r.get("whole_pass",{}).get("ms"),r.get("whole_pass",{}).get("frac"),"step_frac",r.get("whole_step",{}).get("frac"),"surv",c["config"]["survivors_total"])
print("parity ok", d.get("parity_check",{}).get("ok"))
PY
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-parity --no-configs"
timeout 300 $B > gpurun_out/r02c6_plain_cfg2.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02c6_launches_cfg2.csv $B > gpurun_out/r02c6_ncu_cfg2.log 2>&1; echo launches rc=$?
grep -v "^==" gpurun_out/r02c6_launches_cfg2.csv | awk -F'","' 'NR>1{print $5, $NF}' | tail -12
