set -x
for w in cfg2 cfg3; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 4 --steps 20 --warmup 5 --workload $w > gpurun_out/bench4_$w.json 2> gpurun_out/bench4_$w.err; echo $w rc=$?; cat gpurun_out/bench4_$w.json; tail -3 gpurun_out/bench4_$w.err
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 4 --steps 20 --warmup 5 --peer --no-e2e > gpurun_out/bench4_cfg2_peer.json 2> gpurun_out/bench4_cfg2_peer.err; echo peer rc=$?; cat gpurun_out/bench4_cfg2_peer.json
CC_NGPUS=4 timeout 600 python -m pytest tests/test_cli.py -m gpu -x -q -k "ngpus or single" > gpurun_out/pytest_cli_4gpu.log 2>&1; tail -3 gpurun_out/pytest_cli_4gpu.log
