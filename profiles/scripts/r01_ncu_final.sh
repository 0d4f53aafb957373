# round-1 ncu evidence for the final kernels (one ncu "use" per gpurun call)
set -x
for w in cfg2 const1b cfg4 cfg5; do
B="python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
timeout 300 $B > gpurun_out/plain_$w.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_$w.csv $B > gpurun_out/ncu_$w.log 2>&1; echo launches $w rc=$?
done
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
timeout 300 $B > gpurun_out/plain_k2.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:k2_cutout_halo -s 3 -c 1 -o gpurun_out/prof_k2_final -f $B > gpurun_out/ncu_k2.log 2>&1; echo k2 rc=$?
B="python bench.py --workload const1b --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
timeout 300 $B > gpurun_out/plain_k1.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:k1_ -s 9 -c 3 -o gpurun_out/prof_k1_final -f $B > gpurun_out/ncu_k1.log 2>&1; echo k1 rc=$?
B="python bench.py --workload cfg4 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
timeout 300 $B > gpurun_out/plain_k1d.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:k1_ -s 9 -c 3 -o gpurun_out/prof_k1_cfg4_final -f $B > gpurun_out/ncu_k1d.log 2>&1; echo k1 cfg4 rc=$?
