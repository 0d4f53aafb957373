# round 2, call 12: single-kernel use case 1 (k1f_cutout): parity, A/B against the three-kernel form, ncu
set -x
timeout 1500 python -m pytest tests -q -m gpu -x --deselect tests/test_gpu_parity.py::test_device_libm_equals_host_restatement_on_all_floats --deselect tests/test_gpu_parity.py::test_device_arcsec_scaling_equals_specification_on_all_floats > gpurun_out/r02c12_pytest.log 2>&1; echo pytest rc=$?; tail -6 gpurun_out/r02c12_pytest.log
run() { tag=$1; w=$2; shift 2
  env "$@" timeout 400 python bench.py --workload $w --no-cpu-baseline --no-e2e --no-parity --no-configs --steps 10 --warmup 3 > gpurun_out/r02c12_${w}_$tag.json 2> gpurun_out/r02c12_${w}_$tag.err; echo $w $tag rc=$?
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02c12_${w}_$tag.json")); r=d["roofline"]
    print("$w $tag", "value",d["value"],"ms/step",d["ms_per_step"],"kernel_ms",r["kernel_ms"],"frac",r["frac"],"pass_ms",r["whole_pass"]["ms"],"pass_frac",r["whole_pass"]["frac"],"first",d["residency"]["first_cutout_ms"],"surv",d["config"]["survivors_total"])
except Exception as e:
    print("$w $tag failed", e); print(open("gpurun_out/r02c12_${w}_$tag.err").read()[-1500:])
PY
}
for w in cfg4 const1b cfg1; do
  run fused $w CC_K1=fused
  run staged $w CC_K1=staged
done
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active"
B="python bench.py --workload cfg4 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-parity --no-configs"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k1f_ -s 4 -c 1 -o gpurun_out/r02c12_prof_cfg4_fused -f $B > gpurun_out/r02c12_ncu_cfg4.log 2>&1; echo ncu rc=$?
