# round 2 evidence, 2 GPUs: multi-rank tests + CLI on 2 GPUs + default bench line at N=2
set -x
timeout 1200 python -m pytest tests/test_gpu_multirank.py tests/test_cli.py -q -m gpu > gpurun_out/r02f2_pytest_2gpu.log 2>&1; echo pytest rc=$?; tail -5 gpurun_out/r02f2_pytest_2gpu.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
( time timeout 900 $TR bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02f2_bench_2gpu.json 2> gpurun_out/r02f2_bench_2gpu.err ) 2>&1 | tail -3; echo bench2 rc=$?
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02f2_bench_2gpu.json")); r=d["roofline"]
print("cfg2 value",d["value"],"ms",d["ms_per_step"],"kernel_ms",r["kernel_ms"],"e2e",d["e2e"]["value"],d["e2e"].get("roofline",{}).get("frac"))
for k,c in d.get("configs",{}).items():
    if "error" in c: print(k,c); continue
    print("  ",k,"value",c["value"],"ms",c["ms_per_step"])
pc=d.get("parity_check") or {}
print("   parity ok",pc.get("ok"),len(pc.get("checks",[])))
PY
