# round 2, call 7 (8 GPUs): host->device concurrency table, default bench line at N=8 (all configs + parity over peer/NCCL)
set -x
nvidia-smi topo -m 2>&1 | head -24
lscpu | grep -i "numa\|socket\|model name\|^CPU(s)"; free -g | head -2
timeout 300 ./profiles/bin/mb_h2d > gpurun_out/r02c7_mb_h2d_8gpu.txt 2>&1; echo mb rc=$?; cat gpurun_out/r02c7_mb_h2d_8gpu.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29711"
( time timeout 1200 $TR bench.py --gpus 8 > gpurun_out/r02c7_bench_8gpu.json 2> gpurun_out/r02c7_bench_8gpu.err ) 2>&1 | tail -3; echo bench8 rc=$?
tail -4 gpurun_out/r02c7_bench_8gpu.err
python - <<'PY'
import json
for f in ("r02c7_bench_8gpu",):
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
    pc=d.get("parity_check") or {}
    print("   parity ok",pc.get("ok"),[ (c["what"],c["ok_rank0"],c["survivors_total"]) for c in pc.get("checks",[])])
PY
