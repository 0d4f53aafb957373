# launch lists (gpu__time_duration per launch) for the main workloads; one ncu "use"
for w in cfg2 const1b cfg4 cfg5 cfg1; do
B="python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
timeout 300 $B > gpurun_out/plain_$w.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_$w.csv $B > gpurun_out/ncu_$w.log 2>&1; echo launches $w rc=$?
done
