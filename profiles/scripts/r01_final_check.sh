# what the driver runs at round end, in order: gpu tests, smoke, reference arm, default bench
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/final_pytest.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/final_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final_smoke.log 2>&1; echo smoke rc=$?; tail -2 gpurun_out/final_smoke.log
( time timeout 900 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/final_reference.json 2> gpurun_out/final_reference.err ) 2>&1 | grep real; echo ref rc=$?; cat gpurun_out/final_reference.json
( time timeout 900 python bench.py > gpurun_out/final_default.json 2> gpurun_out/final_default.err ) 2>&1 | grep real; echo bench rc=$?; cat gpurun_out/final_default.json
