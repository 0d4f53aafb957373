# usage: bash profiles/scripts/r01_ngpu_final.sh N [tests]   (run under gpurun --gpus N)
N=$1
if [ "$2" = tests ]; then
  timeout 900 python -m pytest tests/test_gpu_multirank.py tests/test_cli.py -m gpu -x -q > gpurun_out/pytest_${N}gpu.log 2>&1; echo pytest rc=$?; tail -4 gpurun_out/pytest_${N}gpu.log
fi
port=29540
for w in ${WORKLOADS:-cfg2 cfg5 cfg3}; do
  port=$((port+1))
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --steps 20 --warmup 5 --workload $w > gpurun_out/bench${N}_$w.json 2> gpurun_out/bench${N}_$w.err; echo $w rc=$?
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench${N}_$w.json")); print("${N}gpu $w", d["value"], d["unit"], "ms/step", d["ms_per_step"], "scaling", d["scaling"], "e2e", d.get("e2e",{}).get("value"), d["config"].get("count_exchange"))
except Exception as e:
    print("failed", e); print(open("gpurun_out/bench${N}_$w.err").read()[-1200:])
PY
done
