mkdir -p gpurun_out
timeout 400 python -m pytest "tests/test_gpu_shard.py::test_in_process_group_matches_reference_tables_and_folds" "tests/test_gpu_shard.py::test_in_process_group_on_a_benchmark_size_fold" "tests/test_gpu_parity.py::test_tuned_kernels_equal_generic_kernels" "tests/test_gpu_parity.py::test_beyond_the_tuned_range" "tests/test_gpu_parity.py::test_real_int16_wrap_matches_the_reference" "tests/test_gpu_parity.py::test_int16_negative_wrap_at_n213" -m gpu -x -q > gpurun_out/r2_gputest_s18.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r2_gputest_s18.log
( echo "lean:"; python profiles/shard_one.py 300 1; echo "old:"; CCJ_SHARD_LEAN=0 python profiles/shard_one.py 300 1 ) > gpurun_out/r2_shard_lean_speed.log 2>&1; cat gpurun_out/r2_shard_lean_speed.log
