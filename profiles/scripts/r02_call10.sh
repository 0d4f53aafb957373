# round 2, call 10: staged many-halo kernels v2 (warp-level match rounds, shared-memory counters in exact, 8192-record chunks)
set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "halo or many or small_cutout or nan or payload or alternating" > gpurun_out/r02c10_pytest.log 2>&1; echo pytest rc=$?; tail -4 gpurun_out/r02c10_pytest.log
timeout 300 python bench.py --workload cfg5 --no-cpu-baseline --no-e2e --no-parity --steps 10 --warmup 3 > gpurun_out/r02c10_cfg5.json 2> gpurun_out/r02c10_cfg5.err; echo cfg5 rc=$?
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02c10_cfg5.json")); r=d["roofline"]
print("cfg5 value",d["value"],"ms",d["ms_per_step"],"kernel_ms",r["kernel_ms"],"pass",r["whole_pass"]["ms"],"halos/s",d.get("halos_per_s"))
PY
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active"
B="python bench.py --workload cfg5 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-parity --no-configs"
timeout 900 ncu --metrics $M --clock-control none -c 120 --csv --log-file gpurun_out/r02c10_launches_cfg5.csv $B > gpurun_out/r02c10_ncu_cfg5.log 2>&1; echo launches rc=$?
python - <<'PY'
import csv
rows=[r for r in csv.reader(open("gpurun_out/r02c10_launches_cfg5.csv")) if len(r)>10]
hdr=rows[0]; ik=hdr.index("Kernel Name"); im=hdr.index("Metric Name"); iv=hdr.index("Metric Value"); iid=hdr.index("ID")
d={}
for r in rows[1:]:
    d.setdefault((int(r[iid]),r[ik]),{})[r[im]]=r[iv]
for (i,k),m in sorted(d.items())[-28:]:
    print(i,k[:34].ljust(34),"us",round(float(m.get("gpu__time_duration.sum",0))/1e3,1),"rdMB",round(float(m.get("dram__bytes_read.sum",0))/1e6,1),"Minst",round(float(m.get("smsp__inst_executed.sum",0))/1e6,2),"issue%",m.get("smsp__issue_active.avg.pct_of_peak_sustained_active","")[:5])
PY
