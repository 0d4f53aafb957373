# 2-GPU run: torchrun bench (NCCL count exchange) + CLI on 2 GPUs vs reference binary
set -x
nvidia-smi -L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench2_cfg2.json 2> gpurun_out/bench2_cfg2.err; echo rc=$?; cat gpurun_out/bench2_cfg2.json; tail -3 gpurun_out/bench2_cfg2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --workload const1b --particles 500000000 > gpurun_out/bench2_const.json 2> gpurun_out/bench2_const.err; echo rc=$?; cat gpurun_out/bench2_const.json; tail -3 gpurun_out/bench2_const.err
CC_TEST_NGPUS=2 timeout 600 python -m pytest tests/test_cli.py -m gpu -x -q > gpurun_out/pytest_cli_2gpu.log 2>&1; echo rc=$?; tail -5 gpurun_out/pytest_cli_2gpu.log
