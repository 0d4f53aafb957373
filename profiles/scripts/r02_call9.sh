# round 2, call 9: ncu evidence -- launch lists (durations, DRAM bytes, instructions) of cfg5 / cfg2 / cfg4, full sets of the new kernels
set -x
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active"
for w in cfg5 cfg2 cfg4; do
  B="python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-parity --no-configs"
  timeout 300 $B > gpurun_out/r02c9_plain_$w.log 2>&1 && timeout 900 ncu --metrics $M --clock-control none -c 120 --csv --log-file gpurun_out/r02c9_launches_$w.csv $B > gpurun_out/r02c9_ncu_$w.log 2>&1; echo launches $w rc=$?
done
B="python bench.py --workload cfg5 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-parity --no-configs"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k2c_ -s 33 -c 3 -o gpurun_out/r02c9_prof_cfg5 -f $B > gpurun_out/r02c9_ncu_cfg5_full.log 2>&1; echo ncu cfg5 full rc=$?
python - <<'PY'
import csv
for w in ("cfg5","cfg2","cfg4"):
    try:
        rows=[r for r in csv.reader(open(f"gpurun_out/r02c9_launches_{w}.csv")) if len(r)>10]
    except Exception as e:
        print(w,e); continue
    hdr=rows[0]; ik=hdr.index("Kernel Name"); im=hdr.index("Metric Name"); iv=hdr.index("Metric Value"); iid=hdr.index("ID")
    d={}
    for r in rows[1:]:
        d.setdefault((int(r[iid]),r[ik]),{})[r[im]]=r[iv]
    print("==",w)
    for (i,k),m in sorted(d.items())[-26:]:
        print(i,k[:34].ljust(34),"us",round(float(m.get("gpu__time_duration.sum",0))/1e3,1),"rdMB",round(float(m.get("dram__bytes_read.sum",0))/1e6,1),"wrMB",round(float(m.get("dram__bytes_write.sum",0))/1e6,1),"Minst",round(float(m.get("smsp__inst_executed.sum",0))/1e6,2),"issue%",m.get("smsp__issue_active.avg.pct_of_peak_sustained_active","")[:5],"warps%",m.get("sm__warps_active.avg.pct_of_peak_sustained_active","")[:5])
PY
