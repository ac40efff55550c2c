mkdir -p gpurun_out
timeout 300 python -m pytest "tests/test_gpu_shard.py::test_in_process_group_matches_reference_tables_and_folds" "tests/test_gpu_shard.py::test_in_process_group_on_a_benchmark_size_fold" "tests/test_gpu_parity.py::test_tuned_kernels_equal_generic_kernels" "tests/test_gpu_parity.py::test_beyond_the_tuned_range" "tests/test_gpu_parity.py::test_real_int16_wrap_matches_the_reference" -m gpu -x -q > gpurun_out/r2_gputest_s21.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r2_gputest_s21.log
( echo "cooperative PR/PM windows, unroll 4:"; python profiles/shard_one.py 300 1
  for v in cw2 cw1 cw4b10; do echo "$v:"; CCJ_B200_LIB=$PWD/ccj_b200/variants/libccj_$v.so python profiles/shard_one.py 300 1 | tail -1; done ) > gpurun_out/r2_shard_lean_coop.log 2>&1; cat gpurun_out/r2_shard_lean_coop.log
