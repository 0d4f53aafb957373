set -x
B="python bench.py --workload cfg5_shared --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-parity --no-configs"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k2p_query|k3_bucket|k3_rank|k3_gather_halo" -s 16 -c 4 -o gpurun_out/r02c26_prof -f $B > gpurun_out/r02c26_ncu.log 2>&1; echo full rc=$?
