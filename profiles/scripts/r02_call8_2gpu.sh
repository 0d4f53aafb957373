# round 2, call 8 (2 GPUs): eager exchange, staged many-halo kernels (A/B against the single-kernel variant), CLI fix
set -x
timeout 1200 python -m pytest tests -q -m gpu -x --deselect tests/test_gpu_parity.py::test_device_libm_equals_host_restatement_on_all_floats --deselect tests/test_gpu_parity.py::test_device_arcsec_scaling_equals_specification_on_all_floats > gpurun_out/r02c8_pytest.log 2>&1; echo pytest rc=$?; tail -8 gpurun_out/r02c8_pytest.log
CC_CELLS_MONO=1 timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "halo or many" > gpurun_out/r02c8_pytest_mono.log 2>&1; echo pytest mono rc=$?; tail -3 gpurun_out/r02c8_pytest_mono.log
for e in "A=1" "CC_CELLS_MONO=1"; do
  env $e CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --workload cfg5 --no-cpu-baseline --no-e2e --no-parity --steps 10 --warmup 3 > gpurun_out/r02c8_cfg5_$e.json 2> gpurun_out/r02c8_cfg5_$e.err; echo cfg5 $e rc=$?
done
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
( time timeout 900 $TR bench.py --gpus 2 > gpurun_out/r02c8_bench_2gpu.json 2> gpurun_out/r02c8_bench_2gpu.err ) 2>&1 | tail -3; echo bench2 rc=$?
tail -4 gpurun_out/r02c8_bench_2gpu.err
python - <<'PY'
import json
for f in ("r02c8_cfg5_A=1","r02c8_cfg5_CC_CELLS_MONO=1","r02c8_bench_2gpu"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json"))
    except Exception as e:
        print(f,"failed",e); print(open(f"gpurun_out/{f}.err").read()[-1500:]); continue
    r=d["roofline"]
    print(f,"value",d["value"],"ms",d["ms_per_step"],"kernel_ms",r["kernel_ms"],"frac",r["frac"],"pass",r["whole_pass"]["ms"],"step_frac",r["whole_step"]["frac"],"e2e",d.get("e2e",{}).get("value"),d.get("e2e",{}).get("roofline"))
    for k,c in d.get("configs",{}).items():
        if "error" in c: print(k,c); continue
        r=c["roofline"]
        print("  ",k,"value",c["value"],"ms",c["ms_per_step"],"kernel_ms",r.get("kernel_ms"),"frac",r["frac"],"pass",r.get("whole_pass",{}).get("ms"),"surv",c["config"]["survivors_total"])
    pc=d.get("parity_check") or {}
    print("   parity ok",pc.get("ok"),[ (c["what"],c["ok_rank0"],c["survivors_total"]) for c in pc.get("checks",[])])
PY
