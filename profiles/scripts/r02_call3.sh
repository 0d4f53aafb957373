set -x
for e in "A=1" "CC_PAYLOAD=never" "CC_NO_BINS=1"; do
  env $e timeout 200 python profiles/scripts/dbg_cfg5_groups.py 125000000 > gpurun_out/r02c3_dbg_$e.log 2>&1; echo "$e rc=$?"; tail -12 gpurun_out/r02c3_dbg_$e.log | cut -c1-250
done
timeout 200 python profiles/scripts/dbg_cfg5_groups.py 125000000 30 > gpurun_out/r02c3_dbg_30only.log 2>&1; echo "30only rc=$?"; tail -4 gpurun_out/r02c3_dbg_30only.log | cut -c1-250
timeout 200 python profiles/scripts/dbg_cfg5_groups.py 20000000 > gpurun_out/r02c3_dbg_20m.log 2>&1; echo "20m rc=$?"; tail -4 gpurun_out/r02c3_dbg_20m.log | cut -c1-250
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "nan_theta or payload_policy or alternating" > gpurun_out/r02c3_pytest.log 2>&1; echo pytest rc=$?; tail -15 gpurun_out/r02c3_pytest.log
