timeout 900 python -m pytest tests/test_gpu_multirank.py tests/test_cli.py -m gpu -x -q > gpurun_out/pytest_2gpu.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/pytest_2gpu.log
run() { name=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29570 + RANDOM % 100)) bench.py --gpus 2 --steps 20 --warmup 5 --workload $W --no-e2e > gpurun_out/m_$name.json 2> gpurun_out/m_$name.err
  python -c "
import json
try:
    d=json.load(open('gpurun_out/m_$name.json')); print('$name', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'])
except Exception as e: print('$name failed', e)"
}
W=cfg5 run cfg5_onecta A=1
W=cfg5 run cfg5_defaultctas CC_NCCL_DEFAULT_CTAS=1
W=cfg2 run cfg2_onecta A=1
W=cfg2 run cfg2_defaultctas CC_NCCL_DEFAULT_CTAS=1
