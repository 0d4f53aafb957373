B="python bench.py --workload cfg5 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
timeout 300 $B > gpurun_out/plain_cfg5.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:k2_cutout_halo_cells -s 3 -c 1 -o gpurun_out/prof_k2cells -f $B > gpurun_out/ncu_k2cells.log 2>&1; echo rc=$?
