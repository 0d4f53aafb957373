# one ncu use per gpurun call.  usage: bash profiles/scripts/r01_ncu_one.sh launches|full <workload> [kernel regex] [skip] [count]
mode=$1; w=$2; k=${3:-.}; s=${4:-0}; c=${5:-1}
B="python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
timeout 300 $B > gpurun_out/plain_$w.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$w.log; exit 1; }
if [ "$mode" = launches ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_${w}.csv $B > gpurun_out/ncu_$w.log 2>&1; echo launches $w rc=$?
else
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c $c -o gpurun_out/prof_${w}_final -f $B > gpurun_out/ncu_full_$w.log 2>&1; echo full $w rc=$?
fi
