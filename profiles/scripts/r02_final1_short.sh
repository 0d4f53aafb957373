# round 2 evidence, 1 GPU, without the ncu passes: the whole GPU suite, smoke, the default bench line + reference arm
set -x
timeout 2400 python -m pytest tests -q -m gpu --durations=6 > gpurun_out/r02f1_pytest.log 2>&1; echo pytest rc=$?; tail -4 gpurun_out/r02f1_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02f1_smoke.log 2>&1; echo smoke rc=$?; tail -2 gpurun_out/r02f1_smoke.log
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02f1_bench_default.json 2> gpurun_out/r02f1_bench_default.err ) 2>&1 | tail -3; echo bench rc=$?
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02f1_bench_default.json"))
r=d["roofline"]
print("cfg2 value",d["value"],"ms",d["ms_per_step"],"kernel_ms",r["kernel_ms"],"frac",r["frac"],"pass",r["whole_pass"]["ms"],"e2e",d["e2e"]["value"],d["e2e"].get("roofline",{}).get("frac"),"launches",d["gpu_launches"])
for k,c in d.get("configs",{}).items():
    if "error" in c: print(k,c); continue
    r=c["roofline"]
    print(k,"value",c["value"],"ms",c["ms_per_step"],"kernel",r.get("kernel"),r.get("kernel_ms"),"frac",r["frac"],"pass",r.get("whole_pass",{}).get("ms"),"surv",c["config"]["survivors_total"],c.get("halos_per_s"),"idx_ms",c.get("residency",{}).get("index_build_ms"))
print("parity ok", d.get("parity_check",{}).get("ok"), "cpu", d.get("cpu_baseline",{}).get("value"))
PY
( time timeout 900 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/r02f1_bench_reference.json 2> gpurun_out/r02f1_bench_reference.err ) 2>&1 | tail -3; echo ref rc=$?; cut -c1-200 gpurun_out/r02f1_bench_reference.json
