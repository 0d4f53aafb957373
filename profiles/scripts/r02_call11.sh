# round 2, call 11: sift append, K3 bucket/place changes: full suite, cfg5 + default bench, launch list
set -x
timeout 1500 python -m pytest tests -q -m gpu -x --deselect tests/test_gpu_parity.py::test_device_libm_equals_host_restatement_on_all_floats --deselect tests/test_gpu_parity.py::test_device_arcsec_scaling_equals_specification_on_all_floats > gpurun_out/r02c11_pytest.log 2>&1; echo pytest rc=$?; tail -4 gpurun_out/r02c11_pytest.log
( time timeout 900 python bench.py > gpurun_out/r02c11_bench_default.json 2> gpurun_out/r02c11_bench_default.err ) 2>&1 | tail -3; echo bench rc=$?
tail -3 gpurun_out/r02c11_bench_default.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02c11_bench_default.json"))
r=d["roofline"]
print("cfg2 value",d["value"],"ms",d["ms_per_step"],"kernel_ms",r["kernel_ms"],"frac",r["frac"],"pass",r["whole_pass"]["ms"],"e2e",d["e2e"]["value"],d["e2e"].get("roofline",{}).get("frac"))
for k,c in d.get("configs",{}).items():
    if "error" in c: print(k,c); continue
    r=c["roofline"]
    print(k,"value",c["value"],"ms",c["ms_per_step"],"kernel_ms",r.get("kernel_ms"),"frac",r["frac"],"pass",r.get("whole_pass",{}).get("ms"),r.get("whole_pass",{}).get("frac"),"step_frac",r.get("whole_step",{}).get("frac"),"surv",c["config"]["survivors_total"],c.get("halos_per_s"))
print("parity ok", d.get("parity_check",{}).get("ok"))
PY
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active"
B="python bench.py --workload cfg5 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-parity --no-configs"
timeout 900 ncu --metrics $M --clock-control none -c 140 --csv --log-file gpurun_out/r02c11_launches_cfg5.csv $B > gpurun_out/r02c11_ncu_cfg5.log 2>&1; echo launches rc=$?
python - <<'PY'
import csv
rows=[r for r in csv.reader(open("gpurun_out/r02c11_launches_cfg5.csv")) if len(r)>10]
hdr=rows[0]; ik=hdr.index("Kernel Name"); im=hdr.index("Metric Name"); iv=hdr.index("Metric Value"); iid=hdr.index("ID")
d={}
for r in rows[1:]:
    d.setdefault((int(r[iid]),r[ik]),{})[r[im]]=r[iv]
for (i,k),m in sorted(d.items())[-32:]:
    print(i,k[:34].ljust(34),"us",round(float(m.get("gpu__time_duration.sum",0))/1e3,1),"rdMB",round(float(m.get("dram__bytes_read.sum",0))/1e6,1),"wrMB",round(float(m.get("dram__bytes_write.sum",0))/1e6,1),"Minst",round(float(m.get("smsp__inst_executed.sum",0))/1e6,2),"issue%",m.get("smsp__issue_active.avg.pct_of_peak_sustained_active","")[:5])
PY
