# round 2, call 15: single-kernel use case 1 with the two-level look-back: parity subset + A/B
set -x
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -q -m gpu -x -k "const or streamed or large_resident or golden or full_size or payload" > gpurun_out/r02c15_pytest.log 2>&1; echo pytest rc=$?; tail -4 gpurun_out/r02c15_pytest.log
run() { tag=$1; w=$2; shift 2
  env "$@" timeout 400 python bench.py --workload $w --no-cpu-baseline --no-e2e --no-parity --no-configs --steps 10 --warmup 3 > gpurun_out/r02c15_${w}_$tag.json 2> gpurun_out/r02c15_${w}_$tag.err; echo $w $tag rc=$?
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02c15_${w}_$tag.json")); r=d["roofline"]
    print("$w $tag", "value",d["value"],"ms/step",d["ms_per_step"],"kernel_ms",r["kernel_ms"],"frac",r["frac"],"pass_ms",r["whole_pass"]["ms"],"pass_frac",r["whole_pass"]["frac"],"surv",d["config"]["survivors_total"])
except Exception as e:
    print("$w $tag failed", e); print(open("gpurun_out/r02c15_${w}_$tag.err").read()[-1500:])
PY
}
for w in cfg4 const1b cfg1; do
  run fused $w CC_K1=fused
done
run staged cfg4 CC_K1=staged
B="python bench.py --workload cfg4 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-parity --no-configs"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k1f_ -s 4 -c 1 -o gpurun_out/r02c15_prof_cfg4_fused -f $B > gpurun_out/r02c15_ncu_cfg4.log 2>&1; echo ncu rc=$?
