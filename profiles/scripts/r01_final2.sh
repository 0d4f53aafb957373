bash profiles/scripts/r01_bench_all.sh v13
timeout 900 python bench.py --workload cfg3 > gpurun_out/bench_cfg3_v13.json 2> gpurun_out/bench_cfg3_v13.err; echo cfg3 rc=$?; head -c 700 gpurun_out/bench_cfg3_v13.json; echo; tail -2 gpurun_out/bench_cfg3_v13.err
( time timeout 600 python bench.py > gpurun_out/bench_default_v13.json 2> gpurun_out/bench_default_v13.err ) 2>&1 | tail -3; cat gpurun_out/bench_default_v13.json
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_v13.json 2> gpurun_out/bench_reference_v13.err ) 2>&1 | tail -3; cat gpurun_out/bench_reference_v13.json; tail -2 gpurun_out/bench_reference_v13.err
