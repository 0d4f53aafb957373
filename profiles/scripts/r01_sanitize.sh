# compute-sanitizer memcheck on the small parity cases (one sanitizer tool per gpurun call)
K='test_const_ragged_sizes or test_const_everything_survives or test_halo_batch_equals_oracle or test_halo_many_halos_one_pass or test_exchange_counts_single_rank or test_halo_empty_inputs or test_const_shards or test_halo_sort_keys'
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 66 --log-file gpurun_out/memcheck.log python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$K" > gpurun_out/memcheck_pytest.log 2>&1; echo memcheck rc=$?
tail -5 gpurun_out/memcheck_pytest.log; tail -8 gpurun_out/memcheck.log
