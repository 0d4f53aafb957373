# round 2, call 18: L2 prefetch of the warp's next tile in the scan kernels (A/B), host-side pick with software prefetch (e2e of the wide field)
set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "const_equals or halo_batch or many_halos_one or streamed_equals" > gpurun_out/r02c18_pytest.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/r02c18_pytest.log
CC_PREFETCH=1 timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "const_equals or halo_batch or many_halos_one or ragged" > gpurun_out/r02c18_pytest_pf.log 2>&1; echo pytest pf rc=$?; tail -3 gpurun_out/r02c18_pytest_pf.log
for w in cfg2 const1b cfg4 cfg5 cfg1; do for pf in 0 1; do
  CC_PREFETCH=$pf timeout 300 python bench.py --workload $w --no-cpu-baseline --no-e2e --no-parity --no-configs --steps 10 --warmup 3 > gpurun_out/r02c18_${w}_pf$pf.json 2> gpurun_out/r02c18_${w}_pf$pf.err; echo $w $pf rc=$?
done; done
timeout 300 python bench.py --workload cfg4 --no-cpu-baseline --no-parity --no-configs --steps 5 --warmup 3 > gpurun_out/r02c18_cfg4_e2e.json 2> gpurun_out/r02c18_cfg4_e2e.err; echo cfg4 e2e rc=$?
python - <<'PY'
import json
for w in ("cfg2","const1b","cfg4","cfg5","cfg1"):
    for pf in (0,1):
        try:
            d=json.load(open(f"gpurun_out/r02c18_{w}_pf{pf}.json")); r=d["roofline"]
            print(w,"prefetch",pf,"ms/step",d["ms_per_step"],"kernel_ms",r["kernel_ms"],"frac",r["frac"],"pass ms",r["whole_pass"]["ms"])
        except Exception as e:
            print(w,pf,"failed",e)
d=json.load(open("gpurun_out/r02c18_cfg4_e2e.json")); print("cfg4 e2e",d["e2e"])
PY
