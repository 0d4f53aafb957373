# ncu evidence for round 1 (one ncu "use": launch lists + full captures of the two scan kernels)
set -x
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
timeout 300 $B > gpurun_out/plain_cfg2.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_cfg2.csv $B > gpurun_out/ncu_cfg2.log 2>&1; echo launches cfg2 rc=$?
timeout 300 $B --workload const1b > gpurun_out/plain_const1b.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_const1b.csv $B --workload const1b > gpurun_out/ncu_const1b.log 2>&1; echo launches const1b rc=$?
timeout 300 $B --workload cfg4 > gpurun_out/plain_cfg4.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_cfg4.csv $B --workload cfg4 > gpurun_out/ncu_cfg4.log 2>&1; echo launches cfg4 rc=$?
timeout 300 $B > gpurun_out/plain_cfg2b.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:k2_cutout_halo -s 3 -c 2 -o gpurun_out/prof_k2 -f $B > gpurun_out/ncu_k2.log 2>&1; echo k2 rc=$?
timeout 300 $B --workload const1b > gpurun_out/plain_k1.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:k1_ -s 9 -c 6 -o gpurun_out/prof_k1 -f $B --workload const1b > gpurun_out/ncu_k1.log 2>&1; echo k1 rc=$?
timeout 300 $B --workload cfg4 > gpurun_out/plain_k1d.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:k1_ -s 9 -c 3 -o gpurun_out/prof_k1_cfg4 -f $B --workload cfg4 > gpurun_out/ncu_k1d.log 2>&1; echo k1 cfg4 rc=$?
