# last check of the round: smoke (now with the index walker) and the ordering tests
set -x
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02c27_smoke.log 2>&1; echo smoke rc=$?; tail -2 gpurun_out/r02c27_smoke.log
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "segment or small_cutout or overflow" -x > gpurun_out/r02c27_pytest.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/r02c27_pytest.log
