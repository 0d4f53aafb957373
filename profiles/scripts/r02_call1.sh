# round 2, call 1: segment-ordered K1 kernels -- parity, A/B against the tile-ordered kernels, gather microbenchmark
set -x
nvidia-smi -L; nproc; free -g | head -2
T="tests/test_gpu_parity.py"
K="const or streamed or large_resident or device_cols"
timeout 900 python -m pytest $T -x -q -m gpu -k "$K" > gpurun_out/r02c1_pytest_seg.log 2>&1; echo pytest seg rc=$?; tail -3 gpurun_out/r02c1_pytest_seg.log
CC_K1_RING_XYZ=1 timeout 900 python -m pytest $T -x -q -m gpu -k "$K" > gpurun_out/r02c1_pytest_seg_xyz.log 2>&1; echo pytest xyz rc=$?; tail -3 gpurun_out/r02c1_pytest_seg_xyz.log
CC_GATHER_BULK=1 timeout 900 python -m pytest $T -x -q -m gpu -k "$K" > gpurun_out/r02c1_pytest_seg_bulk.log 2>&1; echo pytest bulk rc=$?; tail -3 gpurun_out/r02c1_pytest_seg_bulk.log
run() { # tag, workload, env...
  tag=$1; w=$2; shift 2
  env "$@" timeout 400 python bench.py --workload $w --no-cpu-baseline --no-e2e --steps 10 --warmup 3 > gpurun_out/r02c1_${w}_$tag.json 2> gpurun_out/r02c1_${w}_$tag.err; echo $w $tag rc=$?
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02c1_${w}_$tag.json"))
    r=d["roofline"]
    print("$w $tag", "value",d["value"],"ms/step",d["ms_per_step"],"scan_ms",r["kernel_ms"],"scan_frac",r["frac"],"pass_ms",r["whole_pass"]["ms"],"pass_frac",r["whole_pass"]["frac"],"surv",d["config"]["survivors_total"])
except Exception as e:
    print("$w $tag failed", e); print(open("gpurun_out/r02c1_${w}_$tag.err").read()[-1500:])
PY
}
for w in cfg4 const1b cfg1; do
  run tiles $w CC_K1_TILES=1
  run seg $w A=1
  run seg_xyz $w CC_K1_RING_XYZ=1
  run seg_bulk $w CC_GATHER_BULK=1
  run seg_nopayload $w CC_NO_PAYLOAD=1
done
timeout 300 ./profiles/bin/mb_gather > gpurun_out/r02c1_mb_gather.txt 2>&1; echo mb rc=$?; cat gpurun_out/r02c1_mb_gather.txt
B="python bench.py --workload cfg4 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none -c 40 --csv --log-file gpurun_out/r02c1_launches_cfg4_seg.csv $B > gpurun_out/r02c1_ncu_cfg4.log 2>&1; echo launches rc=$?
grep -v "^==" gpurun_out/r02c1_launches_cfg4_seg.csv | awk -F'","' 'NR>1{print $5, $(NF-2), $NF}' | tail -16
