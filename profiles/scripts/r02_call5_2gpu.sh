# round 2, call 5 (2 GPUs): multi-rank parity tests, CLI on 2 GPUs, default bench line at N=2 (peer + NCCL parity leg)
set -x
nvidia-smi topo -m 2>&1 | head -20
lscpu | grep -i "numa\|socket\|model name\|^CPU(s)"
timeout 900 python -m pytest tests/test_gpu_multirank.py tests/test_cli.py -q -m gpu > gpurun_out/r02c5_pytest_2gpu.log 2>&1; echo pytest rc=$?; tail -8 gpurun_out/r02c5_pytest_2gpu.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
( time timeout 900 $TR bench.py --gpus 2 > gpurun_out/r02c5_bench_2gpu.json 2> gpurun_out/r02c5_bench_2gpu.err ) 2>&1 | tail -3; echo bench2 rc=$?
tail -4 gpurun_out/r02c5_bench_2gpu.err
timeout 600 $TR bench.py --gpus 2 --nccl --no-configs --no-parity > gpurun_out/r02c5_bench_2gpu_nccl.json 2> gpurun_out/r02c5_bench_2gpu_nccl.err; echo bench2 nccl rc=$?
python - <<'PY'
import json
for f in ("r02c5_bench_2gpu","r02c5_bench_2gpu_nccl"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json"))
    except Exception as e:
        print(f,"failed",e); continue
    r=d["roofline"]
    print(f,"cfg2 value",d["value"],"ms",d["ms_per_step"],"kernel_ms",r["kernel_ms"],"frac",r["frac"],"e2e",d["e2e"]["value"],d["e2e"]["ms_per_step"],d["config"]["count_exchange"][:40])
    for k,c in d.get("configs",{}).items():
        if "error" in c: print(k,c); continue
        r=c["roofline"]
        print("  ",k,"value",c["value"],"ms",c["ms_per_step"],"kernel_ms",r.get("kernel_ms"),"frac",r["frac"],"pass",r.get("whole_pass",{}).get("ms"),"surv",c["config"]["survivors_total"],"wall",c["bench_wall_s"])
    print("   parity",json.dumps(d.get("parity_check"))[:1800])
PY
timeout 300 ./profiles/bin/mb_h2d > gpurun_out/r02c5_mb_h2d_2gpu.txt 2>&1; echo mb rc=$?; cat gpurun_out/r02c5_mb_h2d_2gpu.txt
