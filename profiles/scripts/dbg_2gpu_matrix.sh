run() { # name, env...
  name=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29570 + RANDOM % 100)) bench.py --gpus 2 --steps 20 --warmup 5 --workload cfg5 --no-e2e > gpurun_out/m_$name.json 2> gpurun_out/m_$name.err
  python -c "
import json
try:
    d=json.load(open('gpurun_out/m_$name.json')); print('$name', d['ms_per_step'], d['roofline']['kernel_ms'])
except Exception as e: print('$name failed', e)"
}
run default A=1
run sys_nccl CC_NCCL_LIB=/usr/lib/x86_64-linux-gnu/libnccl.so.2.27.3
run p2p_off NCCL_P2P_DISABLE=1
run nvls_off NCCL_NVLS_ENABLE=0
run cumem_off NCCL_CUMEM_ENABLE=0
run maxctas1 NCCL_MAX_CTAS=1 NCCL_MAX_NCHANNELS=1
