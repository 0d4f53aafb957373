N=${1:-2}
timeout 900 python -m pytest tests/test_gpu_multirank.py tests/test_cli.py tests/test_gpu_parity.py -m gpu -x -q -k "two_gpus or cli or exchange" > gpurun_out/pytest_xchg.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/pytest_xchg.log
run() { name=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29570 + RANDOM % 100)) bench.py --gpus $N --steps 20 --warmup 5 --workload $W --no-e2e $EXTRA > gpurun_out/x_$name.json 2> gpurun_out/x_$name.err
  python -c "
import json
try:
    d=json.load(open('gpurun_out/x_$name.json')); print('$name', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['whole_pass']['ms'])
except Exception as e: print('$name failed', e); print(open('gpurun_out/x_$name.err').read()[-800:])"
}
W=cfg2 EXTRA= run cfg2_nccl A=1
W=cfg2 EXTRA= run cfg2_nccl_b A=1
W=cfg2 EXTRA=--peer run cfg2_peer A=1
W=cfg5 EXTRA= run cfg5_nccl A=1
W=const1b EXTRA="--particles 250000000" run const_nccl A=1
