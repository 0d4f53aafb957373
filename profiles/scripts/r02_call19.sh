# where do k3_small's ~30 us go in cfg2 (266 survivors)?  full-set capture with source counters
set -x
B="python bench.py --workload cfg2 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-parity --no-configs"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k3_small -s 3 -c 1 -o gpurun_out/r02c19_prof_k3s -f $B > gpurun_out/r02c19_ncu_k3s.log 2>&1; echo full k3s rc=$?
timeout 900 ncu --set full --cache-control none --clock-control none --import-source on -k regex:k3_small -s 3 -c 1 -o gpurun_out/r02c19_prof_k3s_warm -f $B > gpurun_out/r02c19_ncu_k3s_warm.log 2>&1; echo full k3s warm rc=$?
