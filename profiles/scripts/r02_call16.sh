set -x
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active"
B="python bench.py --workload cfg5_onepass --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-parity --no-configs"
timeout 600 ncu --metrics $M --clock-control none -c 80 --csv --log-file gpurun_out/r02c16_launches_cfg5_onepass.csv $B > gpurun_out/r02c16_ncu.log 2>&1; echo launches rc=$?
python - <<'PY'
import csv
rows=[r for r in csv.reader(open("gpurun_out/r02c16_launches_cfg5_onepass.csv")) if len(r)>10]
hdr=rows[0]; ik=hdr.index("Kernel Name"); im=hdr.index("Metric Name"); iv=hdr.index("Metric Value"); iid=hdr.index("ID")
d={}
for r in rows[1:]:
    d.setdefault((int(r[iid]),r[ik]),{})[r[im]]=r[iv]
for (i,k),m in sorted(d.items())[-18:]:
    print(i,k[:34].ljust(34),"us",round(float(m.get("gpu__time_duration.sum",0))/1e3,1),"rdMB",round(float(m.get("dram__bytes_read.sum",0))/1e6,1),"wrMB",round(float(m.get("dram__bytes_write.sum",0))/1e6,1),"Minst",round(float(m.get("smsp__inst_executed.sum",0))/1e6,2),"issue%",m.get("smsp__issue_active.avg.pct_of_peak_sustained_active","")[:5])
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k2c_match -s 4 -c 1 -o gpurun_out/r02c16_prof_match -f $B > gpurun_out/r02c16_ncu_full.log 2>&1; echo full rc=$?
