# round 2, call 17: many halos -- staged kernels vs the one-kernel scan (now with shared-memory counters), per pass; auto choice
set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "halo or many or small_cutout or nan or alternating" > gpurun_out/r02c17_pytest.log 2>&1; echo pytest rc=$?; tail -4 gpurun_out/r02c17_pytest.log
for w in cfg5 cfg5_onepass; do for m in staged mono auto; do
  CC_CELLS=$m timeout 300 python bench.py --workload $w --no-cpu-baseline --no-e2e --no-parity --no-configs --steps 10 --warmup 3 > gpurun_out/r02c17_${w}_$m.json 2> gpurun_out/r02c17_${w}_$m.err; echo $w $m rc=$?
done; done
python - <<'PY'
import json
for w in ("cfg5","cfg5_onepass"):
    for m in ("staged","mono","auto"):
        try:
            d=json.load(open(f"gpurun_out/r02c17_{w}_{m}.json")); r=d["roofline"]
            print(w,m,"ms/step",d["ms_per_step"],"scan-stage ms (last pass)",r["kernel_ms"],"pass ms",r["whole_pass"]["ms"],"halos/s",d.get("halos_per_s"))
        except Exception as e:
            print(w,m,"failed",e); print(open(f"gpurun_out/r02c17_{w}_{m}.err").read()[-800:])
PY
M="gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active"
B="python bench.py --workload cfg5 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-parity --no-configs"
CC_CELLS=mono timeout 600 ncu --metrics $M --clock-control none -c 100 --csv --log-file gpurun_out/r02c17_launches_cfg5_mono.csv $B > gpurun_out/r02c17_ncu.log 2>&1; echo launches rc=$?
python - <<'PY'
import csv
rows=[r for r in csv.reader(open("gpurun_out/r02c17_launches_cfg5_mono.csv")) if len(r)>10]
hdr=rows[0]; ik=hdr.index("Kernel Name"); im=hdr.index("Metric Name"); iv=hdr.index("Metric Value"); iid=hdr.index("ID")
d={}
for r in rows[1:]:
    d.setdefault((int(r[iid]),r[ik]),{})[r[im]]=r[iv]
for (i,k),m in sorted(d.items())[-24:]:
    if "cells" in k: print(i,k[:44].ljust(44),"us",round(float(m.get("gpu__time_duration.sum",0))/1e3,1),"Minst",round(float(m.get("smsp__inst_executed.sum",0))/1e6,2),"issue%",m.get("smsp__issue_active.avg.pct_of_peak_sustained_active","")[:5])
PY
