# round 2 evidence, 8 GPUs: default bench line at N=8 and N=4 (all configs + parity over eager / peer / NCCL exchange)
set -x
for n in ${NLIST:-8 4}; do
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2971$n"
( time timeout 1200 $TR bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r02f8_bench_${n}gpu.json 2> gpurun_out/r02f8_bench_${n}gpu.err ) 2>&1 | tail -3; echo bench$n rc=$?
python - <<PY
import json
d=json.load(open("gpurun_out/r02f8_bench_${n}gpu.json")); r=d["roofline"]
print("N=$n cfg2 value",d["value"],"ms",d["ms_per_step"],"kernel_ms",r["kernel_ms"],"e2e",d["e2e"]["value"],d["e2e"]["ms_per_step"],d["e2e"].get("roofline"))
for k,c in d.get("configs",{}).items():
    if "error" in c: print(k,c); continue
    print("  ",k,"value",c["value"],"ms",c["ms_per_step"],"surv",c["config"]["survivors_total"])
pc=d.get("parity_check") or {}
print("   parity ok",pc.get("ok"),len(pc.get("checks",[])))
PY
done
