set -x
timeout 600 python -m pytest tests/test_gpu_multirank.py tests/test_cli.py -m gpu -x -q > gpurun_out/pytest_2gpu.log 2>&1; echo rc=$?; tail -6 gpurun_out/pytest_2gpu.log
for mode in "" "--no-peer"; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 20 --warmup 5 --no-e2e $mode > gpurun_out/bench2_cfg2$mode.json 2> gpurun_out/bench2_cfg2$mode.err; echo rc=$?
python - <<PY
import json
d=json.load(open("gpurun_out/bench2_cfg2$mode.json")); print("2gpu cfg2 [$mode]", d["value"], d["ms_per_step"], d["config"]["count_exchange"], d["roofline"]["kernel_ms"])
PY
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench2_full.json 2> gpurun_out/bench2_full.err; echo rc=$?; cat gpurun_out/bench2_full.json
