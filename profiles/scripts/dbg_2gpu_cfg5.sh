timeout 300 python bench.py --workload cfg5 --no-cpu-baseline --no-e2e > gpurun_out/d2_single.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/d2_single.json')); print('single process on the 2-GPU box:', d['ms_per_step'], d['roofline']['kernel_ms'])"
OMP_NUM_THREADS=1 timeout 300 python bench.py --workload cfg5 --no-cpu-baseline --no-e2e > gpurun_out/d2_single_omp1.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/d2_single_omp1.json')); print('single process OMP=1:', d['ms_per_step'], d['roofline']['kernel_ms'])"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 20 --warmup 5 --workload cfg5 --no-e2e > gpurun_out/d2_two.json 2> gpurun_out/d2_two.err
python -c "
import json; d=json.load(open('gpurun_out/d2_two.json')); print('two ranks NCCL:', d['ms_per_step'], d['roofline']['kernel_ms'])"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus 2 --steps 20 --warmup 5 --workload cfg5 --no-e2e --peer > gpurun_out/d2_two_peer.json 2> gpurun_out/d2_two_peer.err
python -c "
import json; d=json.load(open('gpurun_out/d2_two_peer.json')); print('two ranks peer:', d['ms_per_step'], d['roofline']['kernel_ms'])"
