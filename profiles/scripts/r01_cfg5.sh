B="python bench.py --workload cfg5 --particles 20000000 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
timeout 300 $B > gpurun_out/plain_cfg5.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:k2_cutout_halo_binned -s 3 -c 1 -o gpurun_out/prof_k2b -f $B > gpurun_out/ncu_k2b.log 2>&1; echo rc=$?
cat gpurun_out/plain_cfg5.log | tail -2
